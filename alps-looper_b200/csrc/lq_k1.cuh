// lq_k1.cuh -- K1, the diagonal update (path_integral.C:403-425,484-537 / standalone/loop.C:93-115).
//
// The reference walks imaginary time carrying spins_c[s]; here every candidate is decided
// independently from the spin at the window start and the OFF-DIAGONAL operators of the window,
// which the diagonal update never changes.
//
// One PERSISTENT CTA per (tile, chunk of consecutive windows): the pages of one tile are
// consecutive in memory, so the CTA streams them one after the other.  Page w+1 is fetched by the
// bulk-copy engine (cp.async.bulk global -> shared, completion on an mbarrier) while page w is
// processed: the buffer is free again as soon as pass 1 has turned the page into the two things the
// rest of the step needs -- the kept (off-diagonal) operators, compacted in page order, and per
// K-site the times of the off-diagonal legs ("columns").  Old diagonal operators (62 % of a page on
// the Heisenberg workloads) are dropped right there (path_integral.C:519-521) and never staged.
//
// Per window:
//   0  per bucket: number of candidates K ~ Poisson(beta v dtau) (k1_poisson); CTA prefix sum ->
//      candidate slots; K-site spins at the window start; columns cleared
//   1  wait for the page; pass A counts the off-diagonal operators per warp slice, pass B writes
//      them to the kept list IN PAGE ORDER (= bucket-major, time-sorted) and appends their times to
//      the columns of their two K-sites; the halo buckets (foreign bonds that touch a K-site) are
//      read straight from global memory and only feed the columns
//      -> barrier; the next page is requested
//   2  per candidate (flat): uniform time, Philox4x32-10 keyed by (bond, window, step, i);
//      is_compatible (graph_impl.h:257) from the two spins at that time = start spins xor parity of
//      the earlier off-diagonal legs: FC branch-free compares per site on one row of the column
//      table; graph by the model's weights (graph_impl.h:679).  Site-graph candidates
//      (graph_impl.h:67-87) are always accepted.
//   3  per bucket: new size = kept + accepted -> CTA prefix sum -> new bucket offsets
//   4  per accepted candidate / kept operator (flat): rank inside the new bucket = kept operators
//      before it (the kept list is sorted) + accepted candidates before it (rejected ones carry a time
//      beyond the window and never count) -> scatter into the compacted new page
#pragma once
#include "lq_device.cuh"

namespace lq {

// 1/K for the Poisson inverse-CDF recursion p_K = p_{K-1} * mu / K
__constant__ double c_rcp[33] = {
    0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11,
    1.0 / 12, 1.0 / 13, 1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19, 1.0 / 20, 1.0 / 21,
    1.0 / 22, 1.0 / 23, 1.0 / 24, 1.0 / 25, 1.0 / 26, 1.0 / 27, 1.0 / 28, 1.0 / 29, 1.0 / 30, 1.0 / 31, 1.0 / 32};

// Number of candidates of one bond in one window: K ~ Poisson(mu) by inversion of the CDF with one
// 53-bit uniform u in (0,1] (replaces looper/poisson_distribution.h:44-113, whose product method
// spends K+1 uniforms, and the exponential gaps of path_integral.C:413-423).  emu = exp(-mu).
// K is capped at 32 (the tail beyond it is < 1e-21 for the means the windows are sized for; the
// host rejects window_ops above 8); *overflow is raised if the cap was hit.
__device__ __forceinline__ int k1_poisson(double u, double emu, double mu, bool* overflow) {
  int K = 0;
  double pk = emu, cdf = pk;
  while (u > cdf && K < 32) { ++K; pk *= mu * c_rcp[K]; cdf += pk; }
  *overflow = (K >= 32 && u > cdf);
  return K;
}

// test hook (tests/test_gpu_poisson.py against test/poisson_distribution.op): histogram of
// k1_poisson over `count` Philox draws
__global__ void k_debug_poisson(double mean, long long count, uint32_t key0, uint32_t key1,
                                unsigned long long* hist, int nbins) {
  const double emu = exp(-mean);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    const philox_t x = philox4x32_10((uint32_t)i, (uint32_t)(i >> 32), 0u, LQ_STREAM_CAND, key0, key1);
    bool ovf;
    const int K = k1_poisson(u53(x.x, x.y), emu, mean, &ovf);
    if (K < nbins) atomicAdd(hist + K, 1ull);
  }
}

// ---- bulk copy (TMA) + mbarrier, sm_90+ PTX -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// generic-proxy reads of a buffer must be ordered before the async proxy overwrites it
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- shared-memory layout ---------------------------------------------------------------------
struct K1Smem {
  double* ptime;     // [capA]  page buffer, bulk-copy destination (TMA variant only)
  uint32_t* pinfo;   // [capA]
  uint32_t* col;     // [nksp][FC+4] keys (k1_key: 32-bit window-relative times) of the off-diagonal legs per
                     // K-site, 0xffffffff-padded: one row per site; the row stride of FC+4 words spreads
                     // the rows over the banks (16-byte reads of 8 different rows hit 8 different bank
                     // groups; a stride of 128 bytes made every read an 8-way conflict and the kernel
                     // shared-memory bound, profiles/r02_k1.md)
  double* ctime;     // [ccap]  candidate times
  double* ktime;     // [kcap]  kept operators, page order
  uint32_t* kse;     // [nbmax] kept list range of a bucket: start | end << 16
  uint32_t* bs2;     // [nloc]  K-sites of the two ends of a local bucket: k0 | k1 << 16 (0xffff: none)
  int* ksite;        // [nksp]  global site of a K-site
  int* hpg;          // [hmax]  first page of the tile that owns a halo bucket
  int* colcnt;       // [nksp]  legs appended to a column (may exceed FC -> exact path)
  int* wsum;         // [34]    per-warp counts of pass A -> exclusive scan
  uint16_t* klb;     // [kcap]  local bucket | site-operator flag << 15
  uint16_t* cmeta;   // [ccap]  local bucket | accepted << 10 | graph << 11
  uint16_t* cbase;   // [nbmax+1] first candidate of a bucket
  uint16_t* noff;    // [nbmax+1] new bucket offsets
  uint16_t* hlb;     // [hmax]  local bucket index of a halo bucket inside its own tile
  uint8_t* kspin;    // [nksp]
};

__host__ __device__ inline int k1_nksp(int nksmax) { return (nksmax + 31) & ~31; }
__host__ __device__ inline int k1_capa(int cap) { return (cap + 3) & ~3; }

// Byte offsets of the arrays above inside the dynamic shared memory, computed ONCE on the host and
// passed as a kernel argument: the sizes are run-time values, and with the offsets derived from them
// inside the kernel every use of a list in divergent code paid ~10 instructions of address arithmetic
// again (8 % of the kernel's instructions, profiles/r02_k1.md); from the constant bank they are an
// operand of the add.
struct K1Layout {
  uint32_t ptime, pinfo, col, ctime, ktime, kse, bs2, ksite, hpg, colcnt, wsum, klb, cmeta, cbase, noff, hlb, kspin, total;
};

inline K1Layout k1_layout(bool tma, int fc, int cap, int ccap, int kcap, int nbmax, int hmax, int nksmax) {
  const size_t nloc = (size_t)nbmax + hmax, nksp = k1_nksp(nksmax), capA = k1_capa(cap);
  K1Layout L;
  size_t p = 0;
  L.ptime = (uint32_t)p; if (tma) p += capA * 8;
  L.pinfo = (uint32_t)p; if (tma) p += capA * 4;    // capA % 4 == 0: stays 16-byte aligned
  L.col = (uint32_t)p; p += nksp * (fc + 4) * 4;    // (nksp % 32 == 0: the next array stays 16-byte aligned)
  L.ctime = (uint32_t)p; p += (size_t)ccap * 8;
  L.ktime = (uint32_t)p; p += ((size_t)kcap * 8 + 15) & ~(size_t)15;
  L.kse = (uint32_t)p; p += (size_t)nbmax * 4;
  L.bs2 = (uint32_t)p; p += nloc * 4;
  L.ksite = (uint32_t)p; p += nksp * 4;
  L.hpg = (uint32_t)p; p += (size_t)hmax * 4;
  L.colcnt = (uint32_t)p; p += nksp * 4;
  L.wsum = (uint32_t)p; p += 34 * 4;
  L.klb = (uint32_t)p; p += (size_t)kcap * 2;
  L.cmeta = (uint32_t)p; p += (size_t)ccap * 2;
  L.cbase = (uint32_t)p; p += ((size_t)nbmax + 1) * 2;
  L.noff = (uint32_t)p; p += ((size_t)nbmax + 1) * 2;
  L.hlb = (uint32_t)p; p += (size_t)hmax * 2;
  L.kspin = (uint32_t)p; p += nksp;
  L.total = (uint32_t)p;
  return L;
}

inline size_t k1_smem_bytes(bool tma, int fc, int cap, int ccap, int kcap, int nbmax, int hmax, int nksmax) {
  return (size_t)k1_layout(tma, fc, cap, ccap, kcap, nbmax, hmax, nksmax).total + 64;
}

__device__ __forceinline__ void k1_carve(const K1Layout& L, unsigned char* smem, K1Smem& S) {
  S.ptime = (double*)(smem + L.ptime);
  S.pinfo = (uint32_t*)(smem + L.pinfo);
  S.col = (uint32_t*)(smem + L.col);
  S.ctime = (double*)(smem + L.ctime);
  S.ktime = (double*)(smem + L.ktime);
  S.kse = (uint32_t*)(smem + L.kse);
  S.bs2 = (uint32_t*)(smem + L.bs2);
  S.ksite = (int*)(smem + L.ksite);
  S.hpg = (int*)(smem + L.hpg);
  S.colcnt = (int*)(smem + L.colcnt);
  S.wsum = (int*)(smem + L.wsum);
  S.klb = (uint16_t*)(smem + L.klb);
  S.cmeta = (uint16_t*)(smem + L.cmeta);
  S.cbase = (uint16_t*)(smem + L.cbase);
  S.noff = (uint16_t*)(smem + L.noff);
  S.hlb = (uint16_t*)(smem + L.hlb);
  S.kspin = (uint8_t*)(smem + L.kspin);
}

// parity of the off-diagonal legs before tc on the two sites of a candidate's bond, straight from the
// pages in global memory (only used when a leg shares the candidate's 32-bit key or a site carries more
// than FC legs in one window).
// A leg at exactly the candidate's f64 time counts as earlier iff its bond comes first in bond_order_key:
// the order the walk gives such a pair (lq_device.cuh).  Legs of the candidate's own bond are seen from
// both of its sites and cancel under any rule.
// (plain arguments: passing the Dev struct by reference would force a copy of the kernel parameters
// into local memory and turn every access to them into a local load.  ONE call for both sites: with one
// call per site the kernel around it came out 2-3 % slower -- 64 registers either way, but 8-24 bytes of
// stack instead of none; K1 at 1024 x 1024, beta = 128: 6.19 -> 6.05 ms.)
__device__ __noinline__ int k1_parity_exact(const int* __restrict__ adj_off, const int* __restrict__ adj,
                                            const int* __restrict__ bond_tl, const uint32_t* __restrict__ tile_key,
                                            int Wl, int nbmax, int cap,
                                            const uint16_t* __restrict__ boff, const uint32_t* __restrict__ info,
                                            const double* __restrict__ time, int wl, int site0, int site1, double tc, int b) {
  const uint32_t kc = bond_order_key(tile_key, bond_tl, b);
  int par = 0;
#pragma unroll 1
  for (int side = 0; side < 2; ++side) {
    const int site = side ? site1 : site0;
    for (int a = adj_off[site]; a < adj_off[site + 1]; ++a) {
      const int b2 = adj[a] >> 1;
      const int tl = bond_tl[b2];
      const size_t p = (size_t)(tl >> 10) * Wl + wl;
      const uint16_t* bo = boff + p * (size_t)(nbmax + 1) + (tl & 1023);
      const size_t base = p * (size_t)cap;
      for (int j = bo[0]; j < bo[1]; ++j) {
        const double t2 = time[base + j];
        const bool before = t2 < tc || (t2 == tc && bond_order_key(tile_key, bond_tl, b2) < kc);
        par ^= (int)(info[base + j] & LQ_INFO_OFFDIAG) & (int)before;
      }
    }
  }
  return par;
}

// Window-relative 32-bit key of a time: monotone in t (floor of a monotone function, the same for
// legs and candidates), so keys order two times whenever they differ; equal keys (2^-32 of a window
// apart) are resolved on the exact path.  < 2^32 - 256, so 0xffffffff pads the columns.
// `shift` > 0 coarsens the keys (LQ_K1_KEYBITS, tests only: ties become frequent, the trajectory
// must not change).
__device__ __forceinline__ uint32_t k1_key(double t, double tlo, double kscale, int shift) {
  return __double2uint_rd((t - tlo) * kscale) >> shift;
}

#define LQ_K1_NONE 0xffffu
#define LQ_K1_QMAX 5   /* buckets per thread whose candidate counts travel in one packed register */

// NT threads; FC (8, 12, 16) = off-diagonal legs per K-site and window held in the column table;
// TMA: pages arrive through cp.async.bulk + mbarrier (else: plain loads, no page buffer);
// TQ: test hook LQ_K1_TIMEBITS (Dev::k1_timebits) -- candidate times cut to a few bits of a window; its own
// instantiation, so that the production kernels carry no trace of it
template <int NT, int FC, bool TMA, bool TQ = false>
__global__ void __launch_bounds__(NT, (NT <= 256 ? 3 : 2))
k_diag_update(Dev d, int src, const StepParams* __restrict__ sp, int chunk_len, const K1Layout L) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ int s_scanb[2][2][32];   // block_exscan2_1b
  int scan_par = 0;
  __shared__ int s_misc[4];                 // [0] error word at entry
  __shared__ __align__(8) uint64_t s_mbar;
  K1Smem S;
  k1_carve(L, s_raw, S);
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u;
  const int wid = tid >> 5;
  constexpr int NW = NT / 32;
  constexpr int CS = FC + 4;   // row stride of the column table (words)
  const int nchunks = (d.Wl + chunk_len - 1) / chunk_len;
  const int t = (int)(blockIdx.x / (unsigned)nchunks);
  const int w_begin = (int)(blockIdx.x - (unsigned)t * (unsigned)nchunks) * chunk_len;
  const int w_end = min(w_begin + chunk_len, d.Wl);
  const int dst = src ^ 1;
  const double beta = sp->beta;
  const uint32_t key0 = sp->key0, key1 = sp->key1, mcs = sp->mcs;

  // ---- per-tile constants -----------------------------------------------------------------------
  const int b0 = d.bond_base[t], nb = d.bond_base[t + 1] - b0;
  const uint32_t kbase = d.tile_key ? (d.tile_key[t] << 10) : (uint32_t)b0;   // Philox counter of bucket lb: kbase + lb
  const int h0 = d.halo_off[t], nh = d.halo_off[t + 1] - h0;
  const int cls = d.tile_class[t];
  const int nks = d.cls_nks[cls];
  const int ns = d.site_base[t + 1] - d.site_base[t];
  const int nksp = k1_nksp(d.nksmax);
  const int Q = (nb + NT - 1) / NT;
  {
    const int* bsx = d.bs + d.cls_bs[cls];
    for (int i = tid; i < nb + nh; i += NT) {
      const int k0 = bsx[2 * i], k1 = bsx[2 * i + 1];
      S.bs2[i] = (uint32_t)(k0 < 0 ? LQ_K1_NONE : k0) | ((uint32_t)(k1 < 0 ? LQ_K1_NONE : k1) << 16);
    }
    for (int k = tid; k < nks; k += NT) S.ksite[k] = k < ns ? d.site_base[t] + k : d.hsite[d.hsite_off[t] + k - ns];
    for (int h = tid; h < nh; h += NT) {
      const int tl = d.bond_tl[d.halo_bond[h0 + h]];
      S.hpg[h] = (tl >> 10) * d.Wl;
      S.hlb[h] = (uint16_t)(tl & 1023);
    }
    for (int i = tid; i < CS * nksp; i += NT) S.col[i] = 0xffffffffu;
    for (int k = tid; k < nksp; k += NT) S.colcnt[k] = 0;
    if (tid == 0) {
      s_misc[0] = *d.d_err;   // read ONCE per CTA: an earlier step of this batch overflowed -> leave both buffers alone
      if (TMA) mbar_init(&s_mbar, 1);
    }
  }
  __syncthreads();
  if (s_misc[0]) return;

  // request a page: bytes rounded up to 16 (the over-read stays inside the page: capacity % 4 == 0)
  auto request_page = [&](int wl) {
    const size_t p = (size_t)t * d.Wl + wl;
    const int n = d.pcount[src][p];
    const uint32_t bt = ((uint32_t)n * 8u + 15u) & ~15u, bi = ((uint32_t)n * 4u + 15u) & ~15u;
    mbar_arrive_expect_tx(&s_mbar, bt + bi);
    if (n > 0) {
      bulk_g2s(S.ptime, d.time[src] + p * (size_t)d.cap, bt, &s_mbar);
      bulk_g2s(S.pinfo, d.info[src] + p * (size_t)d.cap, bi, &s_mbar);
    }
  };
  if (TMA && tid == 0 && w_begin < w_end) request_page(w_begin);
  uint32_t mphase = 0;

  for (int wl = w_begin; wl < w_end; ++wl) {
    const int wg = d.w0 + wl;
    const size_t p = (size_t)t * d.Wl + wl;
    const double tlo = d.wlo[wg], thi = d.wlo[wg + 1], width = thi - tlo;   // = window_lo / window_hi (host table)
    const double kscale = d.wks[wg];   // 4294967040 / width (host table: no f64 division per thread and window)
    const int kshift = d.k1_keyshift;
    const int n_own = d.pcount[src][p];
    uint16_t* bo_new = d.boff[dst] + p * (size_t)(d.nbmax + 1);
    bool page_pending = TMA;   // a bulk copy into the page buffer is in flight / unconsumed

    // the step is lost: mark the page empty, drain the copy engine, leave (all threads take this path together)
    auto bail = [&](int err) {
      if (tid == 0) { atomicOr(d.d_err, err); d.pcount[dst][p] = 0; }
      for (int i = tid; i <= nb; i += NT) bo_new[i] = 0;
      if (TMA && page_pending) mbar_wait(&s_mbar, mphase);
    };

    // halo buckets (foreign bonds that touch a K-site): only their off-diagonal legs matter, and only
    // for the columns.  First thing of the window, so that the dependent loads (bucket extent ->
    // info words -> times, each batch in flight together) overlap the candidate counts below.
    for (int h = tid; h < nh; h += NT) {
      const size_t p2 = (size_t)S.hpg[h] + wl;
      const uint16_t* bo2 = d.boff[src] + p2 * (size_t)(d.nbmax + 1) + S.hlb[h];
      const int o0 = bo2[0], hn = bo2[1] - o0;
      const uint32_t* hi = d.info[src] + p2 * (size_t)d.cap + o0;
      const double* ht = d.time[src] + p2 * (size_t)d.cap + o0;
      const uint32_t kk = S.bs2[nb + h];
      const uint32_t k0 = kk & 0xffffu, k1 = kk >> 16;
      for (int j0 = 0; j0 < hn; j0 += 4) {
        uint32_t ii[4];
        double tt[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) ii[u] = (j0 + u < hn) ? hi[j0 + u] : 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) tt[u] = (ii[u] & LQ_INFO_OFFDIAG) ? ht[j0 + u] : 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (ii[u] & LQ_INFO_OFFDIAG) {
            const uint32_t key = k1_key(tt[u], tlo, kscale, kshift);
            if (k0 != LQ_K1_NONE) { const int f = atomicAdd(&S.colcnt[k0], 1); if (f < FC) S.col[k0 * CS + f] = key; }
            if (k1 != LQ_K1_NONE) { const int f = atomicAdd(&S.colcnt[k1], 1); if (f < FC) S.col[k1 * CS + f] = key; }
          }
      }
    }

    // ---- 0: candidates per bucket ---------------------------------------------------------------
    // (a thread owns Q CONSECUTIVE buckets, so that the prefix sums over threads run in bucket order)
    int ksum = 0;
    uint32_t kpack = 0;   // K of this thread's buckets, 6 bits each (nb <= LQ_K1_QMAX * NT, checked by the host)
    {
      for (int q = 0; q < Q; ++q) {
        const int lb = tid * Q + q;
        if (lb >= nb) break;
        const int b = b0 + lb;
        const double mu = beta * d.bond_rate[b] * width;
        int K = 0;
        if (mu > 0) {
          const philox_t x = philox4x32_10(kbase + (uint32_t)lb, (uint32_t)wg, mcs, LQ_STREAM_CAND, key0, key1);
          bool ovf;
          K = k1_poisson(u53(x.x, x.y), d.bond_emu[b], mu, &ovf);
          if (ovf) atomicOr(d.d_err, LQ_ERR_CAND_FULL);
        }
        kpack |= (uint32_t)K << (6 * q);
        ksum += K;
      }
    }
    // ---- 1: the page -> kept list (page order) + columns ----------------------------------------
    if (TMA) { mbar_wait(&s_mbar, mphase); mphase ^= 1u; page_pending = false; }
    const double* pt = TMA ? S.ptime : d.time[src] + p * (size_t)d.cap;
    const uint32_t* pi = TMA ? S.pinfo : d.info[src] + p * (size_t)d.cap;
    // every warp owns a contiguous slice of the page (a multiple of 32 operators); pass A counts its
    // off-diagonal operators, and ONE two-valued CTA prefix sum hands out the candidate slots of the
    // buckets and the kept-list base of every warp
    const int per_warp = ((n_own + NW * 32 - 1) / (NW * 32)) * 32;
    const int j_lo = min(wid * per_warp, n_own), j_hi = min(j_lo + per_warp, n_own);
    int wcnt = 0;
    for (int j = j_lo + (int)lane; j < j_hi; j += 32) wcnt += (int)(pi[j] & LQ_INFO_OFFDIAG);
    wcnt = __reduce_add_sync(0xffffffffu, wcnt);
    int C, nkept;
    int2 ex = block_exscan2_1b(make_int2(ksum, lane == 0 ? wcnt : 0), &C, &nkept, s_scanb, scan_par);
    // (every thread has left the previous window by now: its lists may be rewritten from here on)
    if (C > d.ccap || nkept > d.kcap) { bail(LQ_ERR_CAND_FULL); return; }
    for (int k = tid; k < nks; k += NT) S.kspin[k] = d.spinW[(size_t)wl * d.N + S.ksite[k]];
    for (int lb = tid; lb < nb; lb += NT) S.kse[lb] = 0u;
    {
      int cb = ex.x;
      for (int q = 0; q < Q; ++q) {
        const int lb = tid * Q + q;
        if (lb >= nb) break;
        const int K = (int)((kpack >> (6 * q)) & 63u);
        S.cbase[lb] = (uint16_t)cb;
        for (int i = 0; i < K; ++i) S.cmeta[cb + i] = (uint16_t)lb;
        cb += K;
      }
      if (tid == 0) S.cbase[nb] = (uint16_t)C;
    }
    {
      int kbase = __shfl_sync(0xffffffffu, ex.y, 0);   // pass B
      for (int j0 = j_lo; j0 < j_hi; j0 += 32) {
        const int j = j0 + (int)lane;
        const uint32_t inf = (j < j_hi) ? pi[j] : 0u;
        const bool offd = (inf & LQ_INFO_OFFDIAG) != 0;
        const unsigned m = __ballot_sync(0xffffffffu, offd);
        if (offd) {
          const double tt = pt[j];
          const int lb = (int)(inf >> LQ_INFO_LBSHIFT);
          const int slot = kbase + __popc(m & ((1u << lane) - 1u));
          S.ktime[slot] = tt;
          S.klb[slot] = (uint16_t)(lb | ((inf & LQ_INFO_SITE) ? 0x8000 : 0));
          const uint32_t kk = S.bs2[lb];
          const uint32_t k0 = kk & 0xffffu, k1 = kk >> 16;
          const uint32_t key = k1_key(tt, tlo, kscale, kshift);
          { const int f = atomicAdd(&S.colcnt[k0], 1); if (f < FC) S.col[k0 * CS + f] = key; }
          if (k1 != LQ_K1_NONE) { const int f = atomicAdd(&S.colcnt[k1], 1); if (f < FC) S.col[k1 * CS + f] = key; }
        }
        kbase += __popc(m);
      }
    }
    __syncthreads();
    // the page buffer is free: fetch the next page while this one is decided and written
    if (TMA && wl + 1 < w_end) {
      if (tid == 0) { fence_proxy_async(); request_page(wl + 1); }
      page_pending = true;
    }

    // kept-list range of every bucket (the list is bucket-major)
    for (int i = tid; i < nkept; i += NT) {
      const int lb = S.klb[i] & 0x3ff;
      if (i == 0 || (S.klb[i - 1] & 0x3ff) != lb) atomicOr(&S.kse[lb], (uint32_t)i);
      if (i + 1 == nkept || (S.klb[i + 1] & 0x3ff) != lb) atomicOr(&S.kse[lb], (uint32_t)(i + 1) << 16);
    }

    // ---- 2: time, acceptance and graph of every candidate ----------------------------------------
    for (int c = tid; c < C; c += NT) {
      const int lb = S.cmeta[c];
      const int i = c - S.cbase[lb];
      const int b = b0 + lb;
      const philox_t x = philox4x32_10(kbase + (uint32_t)lb, (uint32_t)wg, mcs, LQ_STREAM_CAND + 1u + (uint32_t)i, key0, key1);
      // 53 uniform bits, (x.x << 32 | x.y) >> 11, converted in two exact 32-bit halves
      const double frac = __uint2double_rn(x.x >> 11) * (1.0 / 2097152.0) +
                          __uint2double_rn((x.x << 21) | (x.y >> 11)) * (1.0 / 9007199254740992.0);
      double tc = tlo + frac * width;
      if constexpr (TQ) tc = tlo + floor(ldexp(frac, d.k1_timebits)) * ldexp(1.0, -d.k1_timebits) * width;
      if (!(tc < thi)) tc = tlo;
      const uint32_t kk = S.bs2[lb];
      const uint32_t k0 = kk & 0xffffu, k1 = kk >> 16;
      int g = 0;
      if (k1 != LQ_K1_NONE) {   // bond graph; none: site graph, compatible with any spin (graph_impl.h:69)
        int par = S.kspin[k0] ^ S.kspin[k1];   // operators on this bond sit in both columns and cancel
        const uint32_t kc = k1_key(tc, tlo, kscale, kshift);
        const uint4* c0 = (const uint4*)(S.col + k0 * CS);
        const uint4* c1 = (const uint4*)(S.col + k1 * CS);
        bool tie = false;
#pragma unroll
        for (int f = 0; f < FC / 4; ++f) {
          const uint4 a = c0[f], e = c1[f];
          par ^= (int)(a.x < kc) ^ (int)(a.y < kc) ^ (int)(a.z < kc) ^ (int)(a.w < kc) ^
                 (int)(e.x < kc) ^ (int)(e.y < kc) ^ (int)(e.z < kc) ^ (int)(e.w < kc);
          tie |= (a.x == kc) | (a.y == kc) | (a.z == kc) | (a.w == kc) | (e.x == kc) | (e.y == kc) | (e.z == kc) | (e.w == kc);
        }
        // a leg within one key of the candidate, or a column that overflowed: decide on the f64 times
        if (tie || S.colcnt[k0] > FC || S.colcnt[k1] > FC)
          par = (S.kspin[k0] ^ S.kspin[k1]) ^
                k1_parity_exact(d.adj_off, d.adj, d.bond_tl, d.tile_key, d.Wl, d.nbmax, d.cap, d.boff[src], d.info[src], d.time[src], wl,
                                S.ksite[k0], S.ksite[k1], tc, b);
        const float4 pr = d.bond_p[b];
        const float u = u24(x.z);
        g = -1;
        if (par) { if (u < pr.x) g = 0; else if (u < pr.y) g = 2; }
        else     { if (u < pr.z) g = 1; else if (u < pr.w) g = 3; }
      }
      S.ctime[c] = g >= 0 ? tc : 4.0;   // rejected: beyond every window, sinks to the end of its bucket in step 4
      S.cmeta[c] = (uint16_t)(lb | (g >= 0 ? (0x400 | (g << 11)) : 0));
    }
    __syncthreads();

    // ---- 3: new bucket sizes -> offsets ------------------------------------------------------------
    // (the columns are dead from here on: cleared for the next window)
    for (int i = tid; i < CS * nksp; i += NT) S.col[i] = 0xffffffffu;
    for (int k = tid; k < nksp; k += NT) S.colcnt[k] = 0;
    int cnt = 0;
    uint32_t npack = 0;   // accepted candidates of this thread's buckets, 6 bits each
    for (int q = 0; q < Q; ++q) {
      const int lb = tid * Q + q;
      if (lb >= nb) break;
      const int K = (int)((kpack >> (6 * q)) & 63u);
      const int c0 = S.cbase[lb];
      int n = 0;
      for (int i = 0; i < K; ++i) n += (S.cmeta[c0 + i] >> 10) & 1;
      npack |= (uint32_t)n << (6 * q);
      const uint32_t se = S.kse[lb];
      cnt += n + (int)(se >> 16) - (int)(se & 0xffffu);
    }
    int total, unused;
    int run = block_exscan2_1b(make_int2(cnt, 0), &total, &unused, s_scanb, scan_par).x;
    if (total > d.cap) { bail(LQ_ERR_PAGE_FULL); return; }
    if (tid == 0) { bo_new[nb] = (uint16_t)total; d.pcount[dst][p] = total; }

    for (int q = 0; q < Q; ++q) {
      const int lb = tid * Q + q;
      if (lb >= nb) break;
      const int K = (int)((kpack >> (6 * q)) & 63u), nacc = (int)((npack >> (6 * q)) & 63u);
      const uint32_t se = S.kse[lb];
      (void)K;
      S.noff[lb] = (uint16_t)run;
      bo_new[lb] = (uint16_t)run;
      run += nacc + (int)(se >> 16) - (int)(se & 0xffffu);
    }
    __syncthreads();

    // ---- 4: rank inside the new bucket -> scatter into the compacted new page ----------------------
    // flat, one element per thread, with three-instruction loop bodies: a rejected candidate carries
    // the time 4.0 and never counts.  (A sort + merge per bucket thread ran at 10 of 32 lanes and took
    // 29 % of the kernel's instructions, profiles/r02_k1.md.)
    double* wt = d.time[dst] + p * (size_t)d.cap;
    uint32_t* wi = d.info[dst] + p * (size_t)d.cap;
    for (int c = tid; c < C; c += NT) {
      const double tc = S.ctime[c];
      if (tc > 3.0) continue;
      const uint32_t mt = S.cmeta[c];
      const int lb = (int)(mt & 0x3ff);
      const int c0 = S.cbase[lb], c1 = S.cbase[lb + 1];
      const uint32_t se = S.kse[lb];
      int rank = 0;
      for (int k = c0; k < c; ++k) rank += (int)(S.ctime[k] <= tc);        // equal times: the earlier draw first
      for (int k = c + 1; k < c1; ++k) rank += (int)(S.ctime[k] < tc);
      for (int i = (int)(se & 0xffffu); i < (int)(se >> 16); ++i) rank += (int)(S.ktime[i] <= tc);   // kept operators first on ties
      const int pos = S.noff[lb] + rank;
      wt[pos] = tc;
      wi[pos] = ((uint32_t)lb << LQ_INFO_LBSHIFT) | (((mt >> 11) & 3u) << LQ_INFO_GSHIFT) |
                ((d.has_site && (S.bs2[lb] >> 16) == LQ_K1_NONE) ? LQ_INFO_SITE : 0u);
    }
    for (int i = tid; i < nkept; i += NT) {
      const uint32_t kl = S.klb[i];
      const int lb = (int)(kl & 0x3ff);
      const double tt = S.ktime[i];
      const int r0 = i - (int)(S.kse[lb] & 0xffffu);
      int rank = r0;
      const int c1 = S.cbase[lb + 1];
      for (int k = S.cbase[lb]; k < c1; ++k) rank += (int)(S.ctime[k] < tt);
      uint32_t g = 0;
      const int b = b0 + lb;
      const float q0 = d.bond_q[b];
      if (q0 < 1.0f) {  // graph_impl.h:324-327 choose_offdiagonal
        const philox_t x = philox4x32_10(kbase + (uint32_t)lb, (uint32_t)wg, mcs, LQ_STREAM_OFFD + (uint32_t)r0, key0, key1);
        g = (u24(x.x) < q0) ? 0u : 1u;
      }
      const int pos = S.noff[lb] + rank;
      wt[pos] = tt;
      wi[pos] = ((uint32_t)lb << LQ_INFO_LBSHIFT) | (g << LQ_INFO_GSHIFT) | LQ_INFO_OFFDIAG | ((kl & 0x8000u) ? LQ_INFO_SITE : 0u);
    }
    // (no barrier here: the next window touches the columns -- cleared before the prefix sum above --
    // and its own registers only until its first prefix sum, whose barriers every thread reaches
    // after it has finished this window)
  }
}

}  // namespace lq
