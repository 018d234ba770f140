// lq_kernels.cuh -- the sm_100a kernels of one loop-update Monte Carlo step.
//
//   K1 k_diag_update (lq_k1.cuh)   path_integral.C:403-425,484-537 / standalone/loop.C:93-115
//   K2 k_walk, k_carry_scan, k_union_local, k_union_global, k_close
//                      path_integral.C:539-566,584-588; graph_impl.h:67-87,168-177,277-295; union_find.h:242-284
//   K3 k_compress, k_rootflip, k_relabel   union_find.h:325-343 (set_id / copy_id); path_integral.C:796-799
//   K4 k_estimate, k_estimate_sites   path_integral.C:650-737; measurement.h:612-650; susceptibility.h:117-155;
//                      transmag.h:62-117; stiffness.h:82-133; operator flip path_integral.C:815-819
//   K5 k_collect       path_integral.C:774-777; susceptibility.h:182-198
//   K6 k_flip_spins    path_integral.C:820-823
//   k_mr_*             looper/parallel.h:1609-1809 (slab merge, one all-gather + one all-reduce)
//
// All kernels are HBM-bound integer/byte work: no tensor-core path.  Pages are processed one CTA
// per page; node- and cluster-indexed kernels are grid-stride free, sized for the arena capacity
// with a device-side bound so that a whole step needs no host synchronisation.
#pragma once
#include "lq_device.cuh"
#include "lq_k1.cuh"

namespace lq {

struct BucketRef {
  size_t base;  // first slot of the bucket in the page arrays
  int n;        // operators in the bucket
  int idx0;     // dense operator index of the first one
};

__device__ __forceinline__ BucketRef bucket_of(const Dev& d, int buf, int b, int wl) {
  const int tl = d.bond_tl[b];
  const int t = tl >> 10, lb = tl & 1023;
  const size_t p = (size_t)t * d.Wl + wl;
  const uint16_t* bo = d.boff[buf] + p * (size_t)(d.nbmax + 1) + lb;
  const int o0 = bo[0], o1 = bo[1];
  BucketRef r;
  r.base = p * (size_t)d.cap + o0;
  r.n = o1 - o0;
  r.idx0 = d.nbase[p] + o0;
  return r;
}

// flip decision of a cluster (path_integral.C:796-799): one bit per cluster id, packed; written by
// k_flipbits right after the ids are known (and overridden for open clusters on slab engines)
__device__ __forceinline__ uint32_t flip_of(const Dev& d, uint32_t cid) {
  return ((long long)cid < d.nccap) ? ((d.flipw[cid >> 5] >> (cid & 31u)) & 1u) : 0u;
}
// label word written by k_relabel: cluster id (| flip decision << 31 on serial engines)
#define LQ_CID(v) ((v) & 0x7fffffffu)
__device__ __forceinline__ uint32_t flip_of_label(const Dev& d, uint32_t v) {
  return d.fpack ? (v >> 31) : flip_of(d, v);
}

__device__ __forceinline__ node_t upper_node(const Dev& d, int idx, int side) {
  return (node_t)d.N + (d.npo == 2 ? (node_t)(2 * (size_t)idx + side) : (node_t)idx);
}

// ------------------------------------------------------------------------------------------
// Shared-memory stage of one page and its walk halo (the "time-slice tile"): the operators of the
// tile's own buckets are copied as one contiguous stream, the buckets of foreign bonds that touch
// an OWN site are gathered behind them.  All neighbour look-ups of the walk then run on shared
// memory through the static per-class stencils (local bucket ids).  (K1 has its own stage with the
// larger halo of the far-end sites, lq_k1.cuh.)
// The walk only needs the ORDER of the operators at a site and their off-diagonal bit, so an operator
// is staged as ONE 32-bit word: the window-relative time key of K1 (k1_key, monotone in the time)
// cut to 28 bits, the off-diagonal bit below it, and three zero bits that take the number of the merge
// head.  4 bytes per operator instead of 12 (three times as many resident CTAs), integer compares
// instead of f64 ones; two heads whose keys agree in the 28 time bits (2^-28 of a window apart, ~60
// pairs per step at the headline size) are ordered on their f64 times in global memory.
// ------------------------------------------------------------------------------------------
// shared memory through 32-bit shared-space addresses (per-thread tables in inner loops)
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}

#define LQ_WKEY_END 0xfffffff0u   /* exhausted merge head (real keys stay below 2^32 - 256) */
struct Stage {
  uint32_t* key;   // [scap+1] time key << 4 | offdiagonal << 3
  int* off;        // [nloc+1] first staged slot of each local bucket
  int* idx0;       // [nloc]   dense operator index of the first operator of the bucket
  int* gbond;      // [nloc]   global bond id (tie-break order)
  uint2* head;     // [zmax*blockDim] merge heads of the site walk, [head][thread]: {staged position | end << 16,
                   // dense index - staged position}
  int nb, nh;
};

__host__ __device__ inline size_t stage_bytes(int scap, int nbmax, int hmax, int zmax, int tpb) {
  const size_t nloc = (size_t)nbmax + hmax;
  return (((size_t)scap + 1) * 4 + (3 * nloc + 1) * 4 + 7) / 8 * 8 + (size_t)zmax * tpb * 8 + 64;
}

__device__ __forceinline__ uint32_t walk_key(double t, uint32_t inf, double tlo, double kscale, int shift) {
  return (((k1_key(t, tlo, kscale, 0) >> 4) >> shift) << 4) | ((inf & LQ_INFO_OFFDIAG) << 3);
}

__device__ __forceinline__ bool stage_page(const Dev& d, int buf, int t, int wl, unsigned char* smem,
                                           Stage& S, int* s_scan) {
  const int nloc_max = d.nbmax + d.hmax;
  S.key = (uint32_t*)smem;
  S.off = (int*)(S.key + d.scap + 1);
  S.idx0 = S.off + nloc_max + 1;
  S.gbond = S.idx0 + nloc_max;
  S.head = (uint2*)(((uintptr_t)(S.gbond + nloc_max) + 7) & ~(uintptr_t)7);
  const size_t p = (size_t)t * d.Wl + wl;
  const int b0 = d.bond_base[t];
  S.nb = d.bond_base[t + 1] - b0;
  const int h0 = d.halo_off[t];
  S.nh = d.whalo_cnt[t];  // the walk only needs the halo buckets at own sites (they come first)
  const int n_own = d.pcount[buf][p];
  const int base_idx = d.nbase[p];
  const uint16_t* bo = d.boff[buf] + p * (size_t)(d.nbmax + 1);
  const int tid = threadIdx.x;
  const int wg = d.w0 + wl;
  const double tlo = d.wlo[wg], kscale = d.wks[wg];
  const int kshift = d.k1_keyshift;
  // halo buckets: their extents are two dependent loads away (halo_tl -> bucket offsets and operator base).
  // The loads are issued HERE and first used after the own page has been copied, so that the chain runs
  // under that copy instead of behind it.
  int b2 = 0, h_o0 = 0, h_o1 = 0, h_nb = 0;
  size_t h_p2 = 0;
  if (tid < S.nh) {
    b2 = d.halo_bond[h0 + tid];
    const int tl = d.halo_tl[h0 + tid];
    h_p2 = (size_t)(tl >> 10) * d.Wl + wl;
    const uint16_t* bo2 = d.boff[buf] + h_p2 * (size_t)(d.nbmax + 1) + (tl & 1023);
    h_o0 = bo2[0]; h_o1 = bo2[1];
    h_nb = d.nbase[h_p2];
  }
  for (int i = tid; i < S.nb; i += blockDim.x) {
    const int o = bo[i];
    S.off[i] = o;
    S.idx0[i] = base_idx + o;
    S.gbond[i] = b0 + i;
  }
  const double* gt = d.time[buf] + p * (size_t)d.cap;
  const uint32_t* gi = d.info[buf] + p * (size_t)d.cap;
  {
    // four operators per thread and round: all loads are in flight before the first conversion
    const int step = 4 * (int)blockDim.x;
    for (int j0 = tid; j0 < n_own; j0 += step) {
      double tt[4];
      uint32_t ii[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * (int)blockDim.x;
        if (j < n_own) { tt[u] = gt[j]; ii[u] = gi[j]; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * (int)blockDim.x;
        if (j < n_own) S.key[j] = walk_key(tt[u], ii[u], tlo, kscale, kshift);
      }
    }
  }
  BucketRef r;
  r.base = h_p2 * (size_t)d.cap + h_o0;
  r.n = h_o1 - h_o0;
  r.idx0 = h_nb + h_o0;
  int total;
  const int hoff = block_exscan(r.n, &total, s_scan);
  const bool ok = (n_own + total <= d.scap);
  if (tid < S.nh) {
    S.off[S.nb + tid] = n_own + hoff;
    S.idx0[S.nb + tid] = r.idx0;
    S.gbond[S.nb + tid] = b2;
    if (ok)   // batches of four: all loads of a batch are in flight before the first shared-memory store
      for (int j0 = 0; j0 < r.n; j0 += 4) {
        double tt[4];
        uint32_t ii[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u < r.n) { tt[u] = d.time[buf][r.base + j0 + u]; ii[u] = d.info[buf][r.base + j0 + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u < r.n) S.key[n_own + hoff + j0 + u] = walk_key(tt[u], ii[u], tlo, kscale, kshift);
      }
  }
  if (tid == 0) S.off[S.nb + S.nh] = n_own + total;
  __syncthreads();
  return ok;
}

// f64 time of the operator in slot `slot` of the bucket of bond b (rare path of the walk: key ties)
__device__ __noinline__ double walk_exact_time(const int* __restrict__ bond_tl, const uint16_t* __restrict__ boff,
                                               const double* __restrict__ time, int Wl, int nbmax, int cap,
                                               int b, int wl, int slot) {
  const int tl = bond_tl[b];
  const size_t p = (size_t)(tl >> 10) * Wl + wl;
  return time[p * (size_t)cap + boff[p * (size_t)(nbmax + 1) + (tl & 1023)] + slot];
}

// ------------------------------------------------------------------------------------------
// generic exclusive scan of uint32 arrays (bucket bases, root ranks): 3 small kernels
// ------------------------------------------------------------------------------------------
#define LQ_SCAN_ITEMS 4
#define LQ_SCAN_THREADS 1024
#define LQ_SCAN_CHUNK (LQ_SCAN_ITEMS * LQ_SCAN_THREADS)

__global__ void __launch_bounds__(LQ_SCAN_THREADS)
k_scan_blocksum(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ bsum) {
  __shared__ int s_scan[34];
  const size_t base = (size_t)blockIdx.x * LQ_SCAN_CHUNK + (size_t)threadIdx.x * LQ_SCAN_ITEMS;
  int v = 0;
#pragma unroll
  for (int i = 0; i < LQ_SCAN_ITEMS; ++i) if (base + i < n) v += (int)in[base + i];
  int total;
  block_exscan(v, &total, s_scan);
  if (threadIdx.x == 0) bsum[blockIdx.x] = (uint32_t)total;
}

// single CTA: exclusive scan of the block sums in place, total to *total_out (and total_out2)
__global__ void __launch_bounds__(LQ_SCAN_THREADS)
k_scan_top(uint32_t* bsum, size_t nblk, uint32_t* total_out, int* total_out2) {
  __shared__ int s_scan[34];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (size_t start = 0; start < nblk; start += LQ_SCAN_THREADS) {
    const size_t i = start + threadIdx.x;
    const int v = (i < nblk) ? (int)bsum[i] : 0;
    int total;
    const int ex = block_exscan(v, &total, s_scan);
    const uint32_t carry = s_carry;
    if (i < nblk) bsum[i] = carry + (uint32_t)ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + (uint32_t)total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (total_out) *total_out = s_carry;
    if (total_out2) *total_out2 = (int)s_carry;
  }
}

__global__ void __launch_bounds__(LQ_SCAN_THREADS)
k_scan_final(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n,
             const uint32_t* __restrict__ bsum) {
  __shared__ int s_scan[34];
  const size_t base = (size_t)blockIdx.x * LQ_SCAN_CHUNK + (size_t)threadIdx.x * LQ_SCAN_ITEMS;
  uint32_t x[LQ_SCAN_ITEMS];
  int v = 0;
#pragma unroll
  for (int i = 0; i < LQ_SCAN_ITEMS; ++i) { x[i] = (base + i < n) ? in[base + i] : 0u; v += (int)x[i]; }
  int total;
  uint32_t run = bsum[blockIdx.x] + (uint32_t)block_exscan(v, &total, s_scan);
#pragma unroll
  for (int i = 0; i < LQ_SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = run;
    run += x[i];
  }
}

// ------------------------------------------------------------------------------------------
// K2a: parent[x] = x for every node of this step (sites + operator legs)
// ------------------------------------------------------------------------------------------
__global__ void k_init_nodes(Dev d) {  // site nodes; operator nodes are written by k_union_local
  const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x < (size_t)d.N) d.parent[x] = (x < (size_t)d.Nown) ? (node_t)x : NODE_JUNK;   // (ghost sites: lq_space.cuh)
}

// ------------------------------------------------------------------------------------------
// K2b: per site, the node of the world-line segment that crosses the start of every window
// (the reference's current[s], path_integral.C:452, carried through imaginary time).
// ------------------------------------------------------------------------------------------
// The walk below does not know which node enters a window from below (that depends on all earlier
// windows), so it leaves the FIRST leg of every site open -- firstW[wl][s] = operator | side << 31 --
// and reports the node leaving the window at the top in curW[wl+1][s] (NODE_NONE if the site has no
// leg in the window).  This kernel closes the chain: one thread per site runs over the windows,
// carries the crossing node forward and patches the open first legs.  All its loads are coalesced
// over the sites and independent of the carried value.  (It replaces a kernel that looked up the
// last operator of every bucket of every site and window before the walk: 3.3 ms and 12 GB.)
__global__ void k_carry_scan(Dev d) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.Nwalk) return;
  node_t cur = (node_t)s;
  d.curW[s] = cur;
  for (int wl = 0; wl < d.Wl; ++wl) {
    const size_t i = (size_t)wl * d.N + s;
    const uint32_t f = d.firstW[i];
    if (f != NODE_NONE && (long long)(f & 0x7fffffffu) < d.ncap)
      ((f >> 31) ? d.low1 : d.low0)[f & 0x7fffffffu] = cur | ((uint32_t)d.spinW[i] << 31);
    const node_t nxt = d.curW[i + d.N];
    if (nxt == NODE_NONE) d.curW[i + d.N] = cur;
    else cur = nxt;
  }
}

// ------------------------------------------------------------------------------------------
// K2c: world-line walk.  One CTA per page (staged with its halo), one thread per SITE of the
// tile: a z-way merge of the time-sorted buckets incident to the site replays the reference's
// sequential sweep for this site and window (path_integral.C:539-566 with current[s] and
// spins_c[s] in registers): every leg gets the node arriving from below and the spin on it.
// Each leg of the page is visited exactly once -- O(legs x z) instead of one neighbourhood scan
// per operator.  low0/low1[idx] = node below on the source/target side | spin << 31.
// ------------------------------------------------------------------------------------------
// Z = compile-time bound on the coordination number: the merge heads live in registers -- per head the
// key of its next unread operator with the head's number in the low three bits (so the minimum over
// the heads names the winner), its staged position and end packed in one word, and the offset from
// staged position to dense operator index; every step is an integer minimum over Z registers and
// reloads only the head that advanced.  Z == 0: generic version with the head positions in shared
// memory.  Heads that agree in the 28 time bits of the key are ordered on the f64 times, then by
// bond_order_key: two operators of one site at exactly the same time come in the order the diagonal update
// assumed when it accepted the later one (k1_parity_exact; lq_device.cuh).
#ifndef LQ_WALK_MINB
#define LQ_WALK_MINB 5   /* resident 256-thread CTAs per SM the square-lattice walk is compiled for (48 registers;
                            6 = 40 registers spills in the prologue and measured 3 % slower) */
#endif
template <int MAXT, int Z, int NPO>   // NPO: graph nodes per operator (Dev::npo)
__global__ void __launch_bounds__(MAXT, (MAXT <= 256 && Z > 2 && Z <= 4) ? LQ_WALK_MINB : 1)
k_walk(Dev d, int buf) {
  extern __shared__ __align__(16) unsigned char s_stage[];
  __shared__ int s_scan[34];
  const int t = (int)(blockIdx.x / (unsigned)d.Wl), wl = (int)(blockIdx.x - (unsigned)t * (unsigned)d.Wl);
  Stage S;
  const bool staged = stage_page(d, buf, t, wl, s_stage, S, s_scan);
  const int tid = threadIdx.x;
  const int sb = d.site_base[t];
  if (tid >= d.site_base[t + 1] - sb) return;
  const int s = sb + tid;
  if (!staged) {   // the step is lost (sticky error word); leave nothing for k_carry_scan to patch
    if (tid == 0) atomicOr(d.d_err, LQ_ERR_PAGE_FULL);
    d.firstW[(size_t)wl * d.N + s] = NODE_NONE;
    d.curW[(size_t)(wl + 1) * d.N + s] = NODE_NONE;
    return;
  }
  const int cls = d.tile_class[t];
  const int* sso = d.sst_off + d.cls_sso[cls];
  const int* sse = d.sst + d.cls_sst[cls] + sso[tid];
  const int z = sso[tid + 1] - sso[tid];
  node_t cur = 0;          // the node entering from below is patched in by k_carry_scan
  bool have = false;       // a leg of this site has been seen in this window
  uint32_t spin = d.spinW[(size_t)wl * d.N + s];
  // the merge heads of this thread: [head][thread] in shared memory, addressed by 32-bit shared offsets
  // (through generic pointers every access recomputed the shared window: 6 instructions per load)
  const int hs = blockDim.x;
  uint2* const hx = S.head + tid;
  const uint32_t hx_s = smem_u32(hx), key_s = smem_u32(S.key), hstep = (uint32_t)hs * 8u;
  if (Z > 0) {
    constexpr int ZZ = Z > 0 ? Z : 1;
    uint32_t tk[ZZ];           // key | head number of the next unread operator of every head
    uint32_t sdm = 0;          // side of the site on the bond of head k, one bit each
#pragma unroll
    for (int k = 0; k < Z; ++k) {
      tk[k] = LQ_WKEY_END | (uint32_t)k;
      if (k < z) {
        const int ent = sse[k], lid = ent >> 1;
        const int h = S.off[lid], e = S.off[lid + 1];
        sts64(hx_s + (uint32_t)k * hstep, (uint32_t)h | ((uint32_t)e << 16), (uint32_t)(S.idx0[lid] - h));
        sdm |= (uint32_t)(ent & 1) << k;
        if (h < e) tk[k] = S.key[h] | (uint32_t)k;
      }
    }
    // one leg: FIRST = the first leg of the site in this window (its lower end stays open for
    // k_carry_scan); peeled so that the loop below carries no test for it.  Returns false at the end.
    auto step = [&](auto first) -> bool {
      uint32_t m = tk[0];
#pragma unroll
      for (int k = 1; k < Z; ++k) m = min(m, tk[k]);
      if (m >= LQ_WKEY_END) return false;
      // another head within 16 of the minimum shares its time key (or sits on the next one): rare
      uint32_t dmin = 0xffffffffu;
      const uint32_t nm = ~m;   // tk - m - 1: the winner itself wraps to 0xffffffff
#pragma unroll
      for (int k = 0; k < Z; ++k) dmin = min(dmin, tk[k] + nm);
      if (dmin < 15u) {
        const uint32_t mh = m >> 4;
        double bt = 4.0;
        uint32_t bb = 0xffffffffu;
#pragma unroll
        for (int k = 0; k < Z; ++k)
          if ((tk[k] >> 4) == mh) {
            const int lid = sse[k] >> 1, gb = S.gbond[lid];
            const double t2 = walk_exact_time(d.bond_tl, d.boff[buf], d.time[buf], d.Wl, d.nbmax, d.cap, gb, wl,
                                              (int)(lds32(hx_s + (uint32_t)k * hstep) & 0xffffu) - S.off[lid]);
            const uint32_t ok = bond_order_key(d.tile_key, d.bond_tl, gb);   // equal f64 times: the order K1 assumed
            if (t2 < bt || (t2 == bt && ok < bb)) { bt = t2; bb = ok; m = tk[k]; }
          }
      }
      const uint32_t best = m & 7u;
      // the winner's state: one 64-bit shared-memory load, its position goes back incremented
      const uint32_t ha = hx_s + best * hstep;
      const uint2 st = lds64(ha);
      const uint32_t bh = st.x & 0xffffu, ben = st.x >> 16;
      const uint32_t kn = lds32(key_s + 4u * bh + 4u);   // (the stage has one spare word behind the last operator)
      sts32(ha, st.x + 1u);
      const uint32_t tn = ((bh + 1u < ben) ? kn : LQ_WKEY_END) | best;
#pragma unroll
      for (int k = 0; k < Z; ++k) tk[k] = ((uint32_t)k == best) ? tn : tk[k];
      const uint32_t bsd = (sdm >> best) & 1u;
      const uint32_t idx = st.y + bh;
      if (decltype(first)::value) d.firstW[(size_t)wl * d.N + s] = idx | (bsd << 31);
      else d.low0[idx + bsd * d.lowstride] = cur | (spin << 31);   // (one array: low1[i] = low0[lowstride + i])
      spin ^= (m >> 3) & 1u;
      cur = (node_t)d.N + (NPO == 2 ? 2u * idx + bsd : idx);   // = upper_node(d, idx, bsd)
      return true;
    };
    have = step(std::true_type());
    if (have)
      while (step(std::false_type())) {}
  } else {
    for (int k = 0; k < z; ++k) hx[k * hs].x = (uint32_t)S.off[sse[k] >> 1];
    for (;;) {
      int best = -1, bh = 0, bent = 0;
      uint32_t bk = 0;
      bool tie = false;
      for (int k = 0; k < z; ++k) {
        const int ent = sse[k];
        const int lid = ent >> 1;
        const int h = (int)hx[k * hs].x;
        if (h < S.off[lid + 1]) {
          const uint32_t k2 = S.key[h];
          if (best < 0 || (k2 >> 4) < (bk >> 4)) { best = k; bk = k2; bh = h; bent = ent; tie = false; }
          else if ((k2 >> 4) == (bk >> 4)) tie = true;
        }
      }
      if (best < 0) break;
      if (tie) {   // heads with the same time key: f64 times, then bond ids
        const uint32_t mh = bk >> 4;
        double bt = 4.0;
        uint32_t bb = 0xffffffffu;
        for (int k = 0; k < z; ++k) {
          const int ent = sse[k];
          const int lid = ent >> 1;
          const int h = (int)hx[k * hs].x;
          if (h < S.off[lid + 1] && (S.key[h] >> 4) == mh) {
            const int gb = S.gbond[lid];
            const double t2 = walk_exact_time(d.bond_tl, d.boff[buf], d.time[buf], d.Wl, d.nbmax, d.cap, gb, wl, h - S.off[lid]);
            const uint32_t ok = bond_order_key(d.tile_key, d.bond_tl, gb);
            if (t2 < bt || (t2 == bt && ok < bb)) { bt = t2; bb = ok; best = k; bk = S.key[h]; bh = h; bent = ent; }
          }
        }
      }
      hx[best * hs].x = (uint32_t)(bh + 1);
      const int lid = bent >> 1, side = bent & 1;
      const int idx = S.idx0[lid] + (bh - S.off[lid]);
      if (have) (side ? d.low1 : d.low0)[idx] = cur | (spin << 31);
      else d.firstW[(size_t)wl * d.N + s] = (uint32_t)idx | ((uint32_t)side << 31);
      have = true;
      spin ^= (bk >> 3) & 1u;
      cur = upper_node(d, idx, side);
    }
  }
  if (!have) d.firstW[(size_t)wl * d.N + s] = NODE_NONE;
  d.curW[(size_t)(wl + 1) * d.N + s] = have ? cur : NODE_NONE;
}

// ------------------------------------------------------------------------------------------
// K2d: unions, in two levels (the chunk idea of looper/parallel.h applied inside one GPU: resolve
// what is local to a time-slice tile in shared memory, send only its boundary to HBM).
//  k_union_local   one CTA per group of pages.  Edges whose two ends are operator nodes of THIS group are
//                  unified in a shared-memory union-find (same min-index hooking); every node of
//                  the page is then written to parent[] already pointing at its page-local root,
//                  so no separate initialisation pass and shorter global chains.
//  k_union_global  flat over the operators: the remaining edges (an end outside the page: carry
//                  nodes, other windows, other tiles) go to the lock-free union-find in HBM.
// Graph rules (graph_impl.h:277-295):
//   g = 0      unify(below0, below1); the upper legs are the operator's own new node
//   g = 1      cross: upper0 ~ below1, upper1 ~ below0            (needs npo == 2)
//   g = 2, 3   freeze: all four legs in one cluster
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sm_find(uint32_t* par, uint32_t x) {
  uint32_t p = ((volatile uint32_t*)par)[x];
  while (p != x) {
    const uint32_t gp = ((volatile uint32_t*)par)[p];
    if (gp != p) par[x] = gp;
    x = p;
    p = gp;
  }
  return x;
}
__device__ __forceinline__ void sm_union(uint32_t* par, uint32_t a, uint32_t b) {
  uint32_t ra = sm_find(par, a), rb = sm_find(par, b);
  while (ra != rb) {
    if (ra < rb) { const uint32_t t = ra; ra = rb; rb = t; }
    const uint32_t old = atomicCAS(par + ra, ra, rb);
    if (old == ra) return;
    ra = sm_find(par, old);
    rb = sm_find(par, rb);
  }
}

#define LQ_XCAP 1536  /* external edges per union group kept in the list (12 KB of shared memory) */
// edges of one operator (graph rules above); site graphs (graph_impl.h:79-86) cut the world line:
// a new node and no union
template <class F>
__device__ __forceinline__ void op_edges(const Dev& d, uint32_t inf, int idx, node_t p0, node_t p1, F edge) {
  const node_t u0 = upper_node(d, idx, 0);
  if (inf & LQ_INFO_SITE) {   // (with two nodes per operator the spare one joins the new segment)
    if (d.npo == 2) edge(u0, upper_node(d, idx, 1));
    return;
  }
  const int g = (inf >> LQ_INFO_GSHIFT) & 3;
  if (d.npo == 2) {
    const node_t u1 = upper_node(d, idx, 1);
    if (g == 0) { edge(p0, p1); edge(u0, u1); }
    else if (g == 1) { edge(u0, p1); edge(u1, p0); }
    else { edge(p0, p1); edge(u0, p0); edge(u1, p0); }
  } else {
    edge(p0, p1);
    if (g & 2) edge(u0, p0);
  }
}

// Both kernels run one CTA per GROUP of d.ug consecutive windows of one tile: the pages of a group
// are consecutive, so their operator nodes form one contiguous range [lo, hi).  Grouping windows
// makes the links that cross a window boundary (one per site and window) local as well.
__global__ void __launch_bounds__(256)
k_union_local(Dev d, int buf) {
  extern __shared__ uint32_t s_par[];
  const int ngw = (d.Wl + d.ug - 1) / d.ug;
  const int t = blockIdx.x / ngw, w_first = (blockIdx.x % ngw) * d.ug;
  const int w_last = min(w_first + d.ug, d.Wl);
  const size_t p_first = (size_t)t * d.Wl + w_first, p_end = (size_t)t * d.Wl + w_last;
  const int idx_lo = d.nbase[p_first], idx_hi = d.nbase[p_end];
  const int nn = d.npo * (idx_hi - idx_lo);
  const node_t lo = upper_node(d, idx_lo, 0), hi = lo + (node_t)nn;
  // The edges that leave the group (0.12 per operator) are handed to k_union_global as a compact
  // list, so that kernel does not read the operators (12 bytes each) a second time.
  __shared__ uint2 s_x[LQ_XCAP];
  __shared__ int s_xn;
  if (threadIdx.x == 0) s_xn = 0;
  for (int i = threadIdx.x; i < nn; i += blockDim.x) s_par[i] = (uint32_t)i;
  __syncthreads();
  for (size_t p = p_first; p < p_end; ++p) {
    const int n = d.pcount[buf][p];
    const int idx0 = d.nbase[p];
    const uint32_t* gi = d.info[buf] + p * (size_t)d.cap;
    auto edge = [&](node_t a, node_t b) {
      if (a == b) return;   // both legs arrive from the same node (consecutive operators on one bond)
      if (a >= lo && a < hi && b >= lo && b < hi) sm_union(s_par, a - lo, b - lo);
      else { const int slot = atomicAdd(&s_xn, 1); if (slot < d.xcap) s_x[slot] = make_uint2(a, b); }
    };
    // two operators per thread and round: the six loads are in flight before the first (divergent,
    // shared-memory bound) union
    for (int j = threadIdx.x; j < n; j += 2 * blockDim.x) {
      const int j2 = j + blockDim.x;
      const bool two = j2 < n;
      const int idx = idx0 + j, idx2 = idx0 + (two ? j2 : j);
      const node_t a0 = d.low0[idx], a1 = d.low1[idx], b0 = d.low0[idx2], b1 = d.low1[idx2];
      const uint32_t ia = gi[j], ib = gi[two ? j2 : j];
      op_edges(d, ia, idx, a0 & 0x7fffffffu, a1 & 0x7fffffffu, edge);
      if (two) op_edges(d, ib, idx2, b0 & 0x7fffffffu, b1 & 0x7fffffffu, edge);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nn; i += blockDim.x) {
    uint32_t r = (uint32_t)i, pr = s_par[r];
    while (pr != r) { r = pr; pr = s_par[r]; }
    d.parent[lo + i] = lo + r;
  }
  const int xn = s_xn;
  if (threadIdx.x == 0) d.xcount[blockIdx.x] = (xn <= d.xcap) ? xn : -1;   // -1: list overflowed, rescan the group
  if (xn <= d.xcap) {
    uint2* xe = d.xedge + (size_t)blockIdx.x * LQ_XCAP;
    for (int i = threadIdx.x; i < xn; i += blockDim.x) xe[i] = s_x[i];
  }
}

// Only ~0.12 edges per operator leave their group, and every one of them is a chain of dependent
// random loads (1.5 tree hops, one CAS).  Chasing them where they are found leaves 7 of 8 lanes
// idle during the latency-bound part, so k_union_local compacts the external edges of its group
// into a list (it reads the operators anyway) and k_union_global drains the lists with every lane
// holding an edge.
__global__ void __launch_bounds__(256)
k_union_global(Dev d, int buf) {
  const int xn = d.xcount[blockIdx.x];
  if (xn >= 0) {
    // drain the group's list: every lane holds an edge, and the first hop of the next entry is
    // requested while the current one is chased
    const uint2* xe = d.xedge + (size_t)blockIdx.x * LQ_XCAP;
    uint2 e = make_uint2(0u, 0u);
    int i = threadIdx.x;
    if (i < xn) e = xe[i];
    for (; i < xn; i += blockDim.x) {
      const int in = i + blockDim.x;
      uint2 en = make_uint2(0u, 0u);
      if (in < xn) {
        en = xe[in];
        asm volatile("prefetch.global.L2 [%0];" ::"l"(d.parent + en.x));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(d.parent + en.y));
      }
      if (d.dbg & 1) uf_union_count(d.parent, e.x, e.y, d.dbgc);
      else uf_union(d.parent, e.x, e.y);
      e = en;
    }
    return;
  }
  // the list overflowed (more than LQ_XCAP external edges in one group): scan the operators again
  if ((d.dbg & 1) && threadIdx.x == 0) atomicAdd(d.dbgc + 6, 1ull);   // (counted for the tests)
  const int ngw = (d.Wl + d.ug - 1) / d.ug;
  const int t = blockIdx.x / ngw, w_first = (blockIdx.x % ngw) * d.ug;
  const int w_last = min(w_first + d.ug, d.Wl);
  const size_t p_first = (size_t)t * d.Wl + w_first, p_end = (size_t)t * d.Wl + w_last;
  const node_t lo = upper_node(d, d.nbase[p_first], 0), hi = upper_node(d, d.nbase[p_end], 0);
  for (size_t p = p_first; p < p_end; ++p) {
    const int n = d.pcount[buf][p];
    const int idx0 = d.nbase[p];
    const uint32_t* gi = d.info[buf] + p * (size_t)d.cap;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const int idx = idx0 + j;
      const node_t p0 = d.low0[idx] & 0x7fffffffu, p1 = d.low1[idx] & 0x7fffffffu;
      op_edges(d, gi[j], idx, p0, p1, [&](node_t a, node_t b) {
        if (a != b && !(a >= lo && a < hi && b >= lo && b < hi)) uf_union(d.parent, a, b);
      });
    }
  }
}

// close the world lines in imaginary time (path_integral.C:584-588); serial engine only --
// with several slabs the top boundary is merged by the exchange step instead.
__global__ void k_close(Dev d) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.Nown) return;
  uf_union(d.parent, (node_t)s, d.curW[(size_t)d.Wl * d.N + s]);
}

// ------------------------------------------------------------------------------------------
// K3: cluster ids.  set_id numbers the roots in array order (union_find.h:325-330); here the root
// flags of 32 consecutive nodes are one ballot word, ranks come from a scan over the words, and
// copy_id (:338-343) becomes one gather per node.  parent[] ends up holding the cluster id.
// ------------------------------------------------------------------------------------------
#define LQ_NPT 4  /* nodes per thread: independent pointer chases in flight */
// SPACE: engines of the spatial cut hold ghost nodes marked NODE_JUNK (lq_space.cuh); a template parameter so
// that the serial kernel does not carry the test
template <bool SPACE, bool COUNT>   // COUNT: hop statistics (LQ_DBG & 1, experiments only)
__global__ void __launch_bounds__(256)
k_compress(Dev d, size_t nwords_cap) {
  const size_t nn = (size_t)d.N + (size_t)d.npo * (size_t)(*d.d_ntotal);
  const size_t nwords = (nn + 31) >> 5;
  const size_t base = (size_t)blockIdx.x * (256 * LQ_NPT) + threadIdx.x;
  // all unions are done: plain (L1-allocating) loads are safe here and neighbouring nodes mostly
  // chase to the same few roots
  node_t r[LQ_NPT], pr[LQ_NPT];
  unsigned junk = 0;
#pragma unroll
  for (int k = 0; k < LQ_NPT; ++k) {
    const size_t x = base + (size_t)k * 256;
    r[k] = (node_t)x;
    pr[k] = (x < nn) ? d.parent[x] : (node_t)x;
    if (SPACE && pr[k] == NODE_JUNK) { pr[k] = (node_t)x; junk |= 1u << k; }   // unreferenced ghost node
  }
  bool any = true;
  unsigned hops = 0;
  while (any) {
    any = false;
#pragma unroll
    for (int k = 0; k < LQ_NPT; ++k)
      if (pr[k] != r[k]) { r[k] = pr[k]; pr[k] = d.parent[r[k]]; any = true; if (COUNT) ++hops; }
  }
  if (COUNT) { atomicAdd(d.dbgc + 4, (unsigned long long)LQ_NPT); atomicAdd(d.dbgc + 5, (unsigned long long)hops); }
  // root flags of 32 consecutive nodes = one ballot word; the LQ_NPT words of a warp are 8 words apart
  // and stored by its first LQ_NPT lanes
  const unsigned lane = threadIdx.x & 31u;
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < LQ_NPT; ++k) {
    const size_t x = base + (size_t)k * 256;
    bool isroot = false;
    if (x < nn) {
      if (r[k] != (node_t)x) d.parent[x] = r[k];
      isroot = (r[k] == (node_t)x) && !(SPACE && ((junk >> k) & 1u));
    }
    const uint32_t word = __ballot_sync(0xffffffffu, isroot);
    if (lane == (unsigned)k) mine = word;
  }
  if (lane < LQ_NPT) {
    const size_t w = ((base - lane) >> 5) + (size_t)lane * 8;   // (base - lane: node of lane 0, a multiple of 32)
    if (w < nwords_cap) {   // words beyond the live nodes are cleared: the scan input stays clean
      d.bitmap[w] = (w < nwords) ? mine : 0u;
      d.wcount[w] = (w < nwords) ? (uint32_t)__popc(mine) : 0u;
    }
  }
}

__device__ __forceinline__ uint32_t cid_of_root(const Dev& d, node_t r) {
  return d.wbase[r >> 5] + (uint32_t)__popc(d.bitmap[r >> 5] & ((1u << (r & 31)) - 1u));
}

// Serial engines decide the flips per ROOT before the relabelling: one Philox call yields the 32
// flip bits of a bitmap word (path_integral.C:796-799, Bernoulli(1/2) per cluster), packed with the
// word's id base and root flags into one 16-byte record.  k_relabel then needs ONE gather per node
// and stores cluster id | flip << 31, so the estimator and the spin flip read the decision with the
// id instead of chasing a second table.  (Slab engines keep the id-indexed table: the flips of open
// clusters are only known after the exchange.)
__global__ void __launch_bounds__(256)
k_rootflip(Dev d, const StepParams* __restrict__ sp) {
  const size_t nn = (size_t)d.N + (size_t)d.npo * (size_t)(*d.d_ntotal);
  const size_t nwords = (nn + 31) >> 5;
  const uint32_t key0 = sp->key0, key1 = sp->key1, mcs = sp->mcs;
  for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < nwords; w += (size_t)gridDim.x * blockDim.x) {
    const uint32_t bm = d.bitmap[w];
    uint32_t fl = 0;
    if (bm) fl = philox4x32_10((uint32_t)w, (uint32_t)(w >> 32), mcs, LQ_STREAM_FLIP, key0, key1).x & bm;
    d.rootw[w] = make_uint4(d.wbase[w], bm, fl, 0u);
    if (d.has_site && (long long)(w << 5) < d.nccap + 32) d.openw[w] = 0u;   // see k_flipbits
  }
}

__global__ void __launch_bounds__(256)
k_relabel(Dev d) {
  const size_t nn = (size_t)d.N + (size_t)d.npo * (size_t)(*d.d_ntotal);
  const size_t base = (size_t)blockIdx.x * (256 * LQ_NPT) + threadIdx.x;
  if (base == 0 && d.rootflip) {   // (slab engines: k_set_ncs, before the exchange)
    // clusters rooted at a site node come first (site ids are the smallest node ids)
    d.d_nc[1] = cid_of_root(d, (node_t)d.N);
    if ((long long)d.d_nc[0] > d.nccap) atomicOr(d.d_err, LQ_ERR_CLUSTER_FULL);
  }
  node_t r[LQ_NPT];
#pragma unroll
  for (int k = 0; k < LQ_NPT; ++k) {
    const size_t x = base + (size_t)k * 256;
    r[k] = (x < nn) ? d.parent[x] : 0u;
  }
  if (d.rootflip) {
    uint4 q[LQ_NPT];
#pragma unroll
    for (int k = 0; k < LQ_NPT; ++k) q[k] = d.rootw[r[k] >> 5];
#pragma unroll
    for (int k = 0; k < LQ_NPT; ++k) {
      const size_t x = base + (size_t)k * 256;
      if (x < nn)
        d.parent[x] = (q[k].x + (uint32_t)__popc(q[k].y & ((1u << (r[k] & 31)) - 1u))) | (((q[k].z >> (r[k] & 31)) & 1u) << 31);
    }
  } else {
    uint32_t wb[LQ_NPT], bm[LQ_NPT];
#pragma unroll
    for (int k = 0; k < LQ_NPT; ++k) {
      if (r[k] == NODE_JUNK) r[k] = 0u;   // (its label is never read)
      wb[k] = d.wbase[r[k] >> 5]; bm[k] = d.bitmap[r[k] >> 5];
    }
#pragma unroll
    for (int k = 0; k < LQ_NPT; ++k) {
      const size_t x = base + (size_t)k * 256;
      if (x < nn) {   // slab engines relabel AFTER the exchange: the flip table is complete (open clusters included)
        const uint32_t c = wb[k] + (uint32_t)__popc(bm[k] & ((1u << (r[k] & 31)) - 1u));
        d.parent[x] = c | (flip_of(d, c) << 31);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// K4: per-cluster sums of the improved estimators (susceptibility.h:117-155).  Every leg that
// ends (+t) or begins (-t) at an operator contributes to the cluster it belongs to:
//   usize += t/2, umag += t(1/2-c), ssize += g t/2, smag += g t(1/2-c)
// accumulated in integer half-units of 2^-40, so sums do not depend on the order of the atomics.
// One CTA per page; contributions are first merged in a shared-memory hash keyed by cluster id
// (at low temperature most legs of a page belong to a handful of long loops), then flushed with
// one global atomic per distinct key and field.
// ------------------------------------------------------------------------------------------
// Clusters are numbered in the order of their roots, so the clusters rooted inside a page have
// CONSECUTIVE ids [cbase, cend): their sums live in a direct-mapped table (slot = id - cbase; no
// key, no probing, and the flush is one coalesced run of REDs).  Ids outside that range belong to
// clusters rooted in other pages -- few distinct ones, many legs each -- and go through a small hash.
#define LQ_LT 1024   /* direct-mapped slots (clusters rooted in this page)          */
#define LQ_FH 256    /* hashed slots (clusters rooted elsewhere); power of two      */
#define LQ_HASH (LQ_LT + LQ_FH)
struct EstHash {
  uint32_t key[LQ_FH];
  uint32_t lo[4][LQ_HASH];  // 64-bit sums as (lo, hi) pairs: native 32-bit shared atomics with an
  uint32_t hi[4][LQ_HASH];  // explicit carry instead of the CAS loop a 64-bit shared atomicAdd compiles to
};

__device__ __forceinline__ void smem_add64(uint32_t* lo, uint32_t* hi, long long v) {
  const uint32_t vlo = (uint32_t)v, vhi = (uint32_t)((unsigned long long)v >> 32);
  const uint32_t old = atomicAdd(lo, vlo);
  const uint32_t add = vhi + ((uint32_t)(old + vlo) < old ? 1u : 0u);
  if (add) atomicAdd(hi, add);
}

__device__ __forceinline__ void est_global_add(const Dev& d, uint32_t cid, long long a, long long b,
                                               long long c, long long e) {
  if ((long long)cid >= d.nccap) return;  // flagged by k_relabel
  unsigned long long* est = (unsigned long long*)d.est;
  // one array per field: the four sums of a cluster in ONE 32-byte sector were measured slower
  // (k_estimate +15 %, k_collect x2: the L2 atomic units serialise per sector)
  if (a) atomicAdd(est + 0 * d.nccap + cid, (unsigned long long)a);
  if (b) atomicAdd(est + 1 * d.nccap + cid, (unsigned long long)b);
  if (c) atomicAdd(est + 2 * d.nccap + cid, (unsigned long long)c);
  if (e) atomicAdd(est + 3 * d.nccap + cid, (unsigned long long)e);
}

// slot of a cluster in the page's table, -1 if the hashed part is crowded
__device__ __forceinline__ int est_slot(EstHash* h, uint32_t cid, uint32_t cbase, uint32_t nloc) {
  const uint32_t li = cid - cbase;
  if (li < nloc) return (int)li;
  uint32_t s = (cid * 2654435761u) >> 24;  // 8 bits
  for (int probe = 0; probe < 8; ++probe) {
    const uint32_t k = atomicCAS(&h->key[s], 0xffffffffu, cid);
    if (k == 0xffffffffu || k == cid) return LQ_LT + (int)s;
    s = (s + 1) & (LQ_FH - 1);
  }
  return -1;
}

__device__ __forceinline__ void est_hash_add(const Dev& d, EstHash* h, uint32_t cbase, uint32_t nloc, uint32_t cid,
                                             long long a, long long b, long long c, long long e) {
  const int slot = est_slot(h, cid, cbase, nloc);
  if (slot < 0) { est_global_add(d, cid, a, b, c, e); return; }  // table crowded: go straight to HBM
  if (a) smem_add64(&h->lo[0][slot], &h->hi[0][slot], a);
  if (b) smem_add64(&h->lo[1][slot], &h->hi[1][slot], b);
  if (c) smem_add64(&h->lo[2][slot], &h->hi[2][slot], c);
  if (e) smem_add64(&h->lo[3][slot], &h->hi[3][slot], e);
}

// winding-number legs (stiffness.h:93-104, source side only): three more 32-bit fields per slot of
// the same table, one global atomic per distinct cluster and dimension at the end
struct WindHash { int w[3][LQ_HASH]; };

__device__ __forceinline__ void wind_add(const Dev& d, EstHash* h, WindHash* wh, uint32_t cbase, uint32_t nloc,
                                         uint32_t cid, int sgn, const short* vec) {
  const int slot = est_slot(h, cid, cbase, nloc);
  if (slot >= 0) {
    for (int x = 0; x < d.sdim; ++x)
      if (vec[x]) atomicAdd(&wh->w[x][slot], sgn * (int)vec[x]);
  } else if ((long long)cid < d.nccap) {
    for (int x = 0; x < d.sdim; ++x)
      if (vec[x]) atomicAdd(d.wind + (size_t)x * d.nccap + cid, sgn * (int)vec[x]);
  }
}

// Bernoulli(1/2) per cluster (path_integral.C:796-799): Philox4x32-10 keyed by (cluster id, rank,
// step); 32 clusters per thread, one packed word each.  Persistent grid.
__global__ void __launch_bounds__(256)
k_flipbits(Dev d, const StepParams* __restrict__ sp) {
  const uint32_t nc = d.d_nc[0];
  const size_t nw = ((size_t)nc + 31) >> 5;
  const uint32_t key0 = sp->key0, key1 = sp->key1, mcs = sp->mcs;
  for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < nw && (long long)(w << 5) < d.nccap;
       w += (size_t)gridDim.x * blockDim.x) {
    uint32_t word = 0;
    for (int i = 0; i < 32; i += 4) {   // one Philox call yields four bits' worth of words
      const philox_t x = philox4x32_10((uint32_t)(w * 8 + (i >> 2)), (uint32_t)d.rank, mcs, LQ_STREAM_FLIP, key0, key1);
      word |= (x.x & 1u) << i | (x.y & 1u) << (i + 1) | (x.z & 1u) << (i + 2) | (x.w & 1u) << (i + 3);
    }
    d.flipw[w] = word;
    if (d.has_site) d.openw[w] = 0u;   // "cut by a site operator" flags of this step (transmag.h:72-81)
  }
}

// FLIP: also apply the cluster flip to the operator (K6, path_integral.C:815-819): its type changes
// iff the cluster arriving from below on the source side and the one leaving upwards are flipped
// differently -- the two cluster ids are already in registers here.
#define LQ_EST_U 1  /* operators per thread and iteration: independent gathers in flight */
// NPO2: two nodes per operator (models with the cross graph); a template parameter so that the
// usual one-node instantiation does not carry the four-leg registers
template <bool FLIP, bool STIFF, bool NPO2>
__global__ void __launch_bounds__(256)
k_estimate(Dev d, int buf) {
  extern __shared__ unsigned char s_raw[];
  EstHash* h = (EstHash*)s_raw;
  WindHash* wh = (WindHash*)(h + 1);           // (STIFF only)
  short* s_vec = (short*)(wh + 1);             // [3*nbmax] relative vectors of the own bonds (STIFF only)
  signed char* s_gg = STIFF ? (signed char*)(s_vec + 3 * d.nbmax) : (signed char*)(h + 1);   // [2*nbmax] gauge of the two ends of every own bond
  for (int i = threadIdx.x; i < LQ_HASH; i += blockDim.x) {
    if (i < LQ_FH) h->key[i] = 0xffffffffu;
#pragma unroll
    for (int f = 0; f < 4; ++f) { h->lo[f][i] = 0u; h->hi[f][i] = 0u; }
    if (STIFF) { wh->w[0][i] = 0; wh->w[1][i] = 0; wh->w[2][i] = 0; }
  }
  const size_t p = blockIdx.x;
  const int t = (int)(blockIdx.x / (unsigned)d.Wl);
  const int b0 = d.bond_base[t], nb = d.bond_base[t + 1] - b0;
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    s_gg[2 * i] = d.gauge[d.bond_s0[b0 + i]];
    const int s1 = d.bond_s1[b0 + i];
    s_gg[2 * i + 1] = s1 >= 0 ? d.gauge[s1] : (signed char)0;
    if (STIFF) { s_vec[3 * i] = d.bond_vec[3 * (b0 + i)]; s_vec[3 * i + 1] = d.bond_vec[3 * (b0 + i) + 1]; s_vec[3 * i + 2] = d.bond_vec[3 * (b0 + i) + 2]; }
  }
  __syncthreads();
  const int n = d.pcount[buf][p];
  const int idx0 = d.nbase[p];
  // an arena overflowed in this batch: the page buffers hold the configuration to rewind to (or
  // scratch) -- no operator may change type any more (lq_engine.cu sweep_many)
  const bool dead = FLIP && (*d.d_err != 0);
  // ids of the clusters rooted in this page: [cbase, cbase + nloc) (capped at the table size)
  const node_t nlo = upper_node(d, idx0, 0), nhi = nlo + (node_t)(d.npo * n);
  const uint32_t cbase = cid_of_root(d, nlo);
  const uint32_t nloc = min(cid_of_root(d, nhi) - cbase, (uint32_t)LQ_LT);
  // (staging the cluster ids of the page's own nodes in shared memory was measured 10 % slower than
  // gathering them: the gathers already hit L1/L2)
  uint32_t* ginfo = d.info[buf] + p * (size_t)d.cap;
  const double* gtime = d.time[buf] + p * (size_t)d.cap;
  for (int j0 = threadIdx.x; j0 < n; j0 += blockDim.x * LQ_EST_U) {
    uint32_t inf[LQ_EST_U], l0[LQ_EST_U], l1[LQ_EST_U], cl0[LQ_EST_U], cu0[LQ_EST_U], cl1[LQ_EST_U], cu1[LQ_EST_U];
    double tt[LQ_EST_U];
    bool act[LQ_EST_U];
    // all loads of the LQ_EST_U operators are issued before the first use
#pragma unroll
    for (int u = 0; u < LQ_EST_U; ++u) {
      const int j = j0 + u * (int)blockDim.x;
      act[u] = j < n;
      const int jj = act[u] ? j : j0;
      inf[u] = ginfo[jj];
      tt[u] = gtime[jj];
      l0[u] = d.low0[idx0 + jj];
      if (NPO2) l1[u] = d.low1[idx0 + jj];   // (one node per operator: only its spin bit would be used, see below)
    }
#pragma unroll
    for (int u = 0; u < LQ_EST_U; ++u) {
      const int idx = idx0 + (act[u] ? j0 + u * (int)blockDim.x : j0);
      cl0[u] = d.parent[l0[u] & 0x7fffffffu];
      cu0[u] = d.parent[upper_node(d, idx, 0)];
      if (NPO2 && !(inf[u] & LQ_INFO_SITE)) {   // (low1 of a site operator is never written)
        cl1[u] = d.parent[l1[u] & 0x7fffffffu];
        cu1[u] = d.parent[upper_node(d, idx, 1)];
      }
    }
#pragma unroll
    for (int u = 0; u < LQ_EST_U; ++u) {
      if (!act[u]) continue;
      const int g = (inf[u] >> LQ_INFO_GSHIFT) & 3;
      if (g & 2) continue;  // frozen graphs: skipped by the estimators (path_integral.C:692), never flip
      const int j = j0 + u * (int)blockDim.x;
      // path integral: imaginary time in 2^-40 fixed point; SSE: position in the operator string (sse.C:358-361)
      const long long q = d.sse ? (long long)d.spos[idx0 + j] : time_to_fx(tt[u]);
      const int lb = (int)(inf[u] >> LQ_INFO_LBSHIFT);
      const int g0 = s_gg[2 * lb], g1 = s_gg[2 * lb + 1];
      // spins below (written by the walk).  Without the cross graph the spin on the target side follows
      // from the graph: graphs 0 and 2 sit on antiparallel spins, 3 on parallel ones (graph_impl.h:257) --
      // four bytes per operator that need not be read
      const int c0 = (int)(l0[u] >> 31), c1 = NPO2 ? (int)(l1[u] >> 31) : (c0 ^ 1 ^ (g & 1));
      const int off = (int)(inf[u] & LQ_INFO_OFFDIAG);
      const int m0 = 1 - 2 * c0, m1 = 1 - 2 * c1;              // 2(1/2-c) below
      const int n0 = 1 - 2 * (c0 ^ off), n1 = 1 - 2 * (c1 ^ off);  // above
      if (FLIP && !dead && ((flip_of_label(d, cl0[u]) ^ flip_of_label(d, cu0[u])) & 1u)) ginfo[j] = inf[u] ^ LQ_INFO_OFFDIAG;
      cl0[u] = LQ_CID(cl0[u]); cu0[u] = LQ_CID(cu0[u]);
      if (NPO2) { cl1[u] = LQ_CID(cl1[u]); cu1[u] = LQ_CID(cu1[u]); }
      if (inf[u] & LQ_INFO_SITE) {
        // site operator: end_s below / begin_s above on its one site (path_integral.C:718-726);
        // both clusters are cut open for the transverse magnetisation (transmag.h:72-81)
        est_hash_add(d, h, cbase, nloc, cl0[u], q, q * m0, q * g0, q * g0 * m0);
        est_hash_add(d, h, cbase, nloc, cu0[u], -q, -q * n0, -q * g0, -q * g0 * n0);
        if ((long long)cl0[u] < d.nccap) atomicOr(d.openw + (cl0[u] >> 5), 1u << (cl0[u] & 31u));
        if ((long long)cu0[u] < d.nccap) atomicOr(d.openw + (cu0[u] >> 5), 1u << (cu0[u] & 31u));
        continue;
      }
      if (STIFF) {   // stiffness.h:93-104: end_bs adds (1-2c) vr below, begin_bs subtracts it above
        wind_add(d, h, wh, cbase, nloc, cl0[u], m0, s_vec + 3 * lb);
        wind_add(d, h, wh, cbase, nloc, cu0[u], -n0, s_vec + 3 * lb);
      }
      if (!NPO2) {
        // l0 = l1 = cl0, u0 = u1 = cu0 (graph 0).  When the cluster below is the cluster above (most
        // legs of the long loops) the two contributions are merged: a diagonal operator then adds
        // nothing at all, an off-diagonal one only its change of magnetisation
        if (cl0[u] == cu0[u]) {
          const long long b = q * ((m0 + m1) - (n0 + n1)), e = q * ((g0 * m0 + g1 * m1) - (g0 * n0 + g1 * n1));
          if (b | e) est_hash_add(d, h, cbase, nloc, cl0[u], 0, b, 0, e);
        } else {
          est_hash_add(d, h, cbase, nloc, cl0[u], 2 * q, q * (m0 + m1), q * (g0 + g1), q * (g0 * m0 + g1 * m1));
          est_hash_add(d, h, cbase, nloc, cu0[u], -2 * q, -q * (n0 + n1), -q * (g0 + g1), -q * (g0 * n0 + g1 * n1));
        }
      } else {
        // four legs; legs of the same cluster are merged before they reach the table
        uint32_t id[4] = {cl0[u], cl1[u], cu0[u], cu1[u]};
        long long A[4] = {q, q, -q, -q}, B[4] = {q * m0, q * m1, -q * n0, -q * n1};
        long long C[4] = {q * g0, q * g1, -q * g0, -q * g1}, E[4] = {q * g0 * m0, q * g1 * m1, -q * g0 * n0, -q * g1 * n1};
#pragma unroll
        for (int i = 1; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < i; ++j)
            if (id[i] == id[j]) { A[j] += A[i]; B[j] += B[i]; C[j] += C[i]; E[j] += E[i]; A[i] = B[i] = C[i] = E[i] = 0; }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (A[i] | B[i] | C[i] | E[i]) est_hash_add(d, h, cbase, nloc, id[i], A[i], B[i], C[i], E[i]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < LQ_HASH; i += blockDim.x) {
    const uint32_t k = (i < LQ_LT) ? ((uint32_t)i < nloc ? cbase + (uint32_t)i : 0xffffffffu) : h->key[i - LQ_LT];
    if (k != 0xffffffffu) {
      long long v[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) v[f] = (long long)(((unsigned long long)h->hi[f][i] << 32) | h->lo[f][i]);
      est_global_add(d, k, v[0], v[1], v[2], v[3]);
      if (STIFF && (long long)k < d.nccap)
        for (int x = 0; x < d.sdim; ++x)
          if (wh->w[x][i]) atomicAdd(d.wind + (size_t)x * d.nccap + k, wh->w[x][i]);
    }
  }
}

// start_bottom / stop_top of every world line (path_integral.C:666-669,729-733;
// susceptibility.h:139-155); for a slab [tau0,tau1) they are start(tau0) / stop(tau1)
// (path_integral_mpi.C), and only rank 0 owns the tau = 0 magnetisations.
// At low temperature most world lines belong to a few long loops, so the lanes of a warp are first
// grouped by cluster id (match.any) and reduced with the integer warp reduction; one atomic per
// group and field reaches L2 instead of one per site.
__device__ __forceinline__ void site_group_add(const Dev& d, bool valid, uint32_t cid, int a, int b, int c, int e,
                                               long long q, bool to_est0) {
  const uint32_t key = valid ? cid : 0xffffffffu;
  const unsigned grp = __match_any_sync(0xffffffffu, key);
  const int sa = __reduce_add_sync(grp, a), sb = __reduce_add_sync(grp, b);
  const int sc = __reduce_add_sync(grp, c), se = __reduce_add_sync(grp, e);
  const unsigned lane = threadIdx.x & 31u;
  if (!valid || lane != (unsigned)(__ffs(grp) - 1)) return;
  if (to_est0) {
    if (sa) atomicAdd(d.est0 + 0 * (size_t)d.N + cid, sa);
    if (sb) atomicAdd(d.est0 + 1 * (size_t)d.N + cid, sb);
    if (sc) atomicAdd(d.est0 + 2 * (size_t)d.N + cid, sc);
    if (se) atomicAdd(d.est0 + 3 * (size_t)d.N + cid, se);
  } else {
    est_global_add(d, cid, q * sa, q * sb, q * sc, q * se);
  }
}

__global__ void __launch_bounds__(128)
k_estimate_sites(Dev d) {
  const int s0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = s0 < d.Nown;
  const int s = valid ? s0 : 0;
  const int g = d.gauge[s];
  const int c = d.spinW[s];
  const int m = 1 - 2 * c;
  const uint32_t cb = LQ_CID(d.parent[s]);
  const uint32_t ct = LQ_CID(d.parent[d.curW[(size_t)d.Wl * d.N + s]]);
  // (SSE: the world lines start at string position 0 and stop at the string length, sse.C:200,283)
  const long long qlo = d.sse ? 0ll : time_to_fx(window_lo(d.w0, d.W));
  const long long qhi = d.sse ? (long long)(*d.d_ntotal)
                              : ((d.w0 + d.Wl >= d.W) ? (1ll << 40) : time_to_fx(window_hi(d.w0 + d.Wl - 1, d.W)));
  if (d.rank == 0 || d.space) site_group_add(d, valid, cb, 1, m, g, g * m, 0, true);   // (tau = 0 lies in rank 0's slab)
  if (qlo) site_group_add(d, valid, cb, 1, m, g, g * m, -qlo, false);
  // periodic in imaginary time: the spin at the top of the slab stack equals the one at tau = 0;
  // inside a slab it is the spin at the start of the next slab = spinW[Wl]
  const int ctop = d.spinW[(size_t)d.Wl * d.N + s];
  const int mt = 1 - 2 * ctop;
  site_group_add(d, valid, ct, 1, mt, g, g * mt, qhi, false);
}

// ------------------------------------------------------------------------------------------
// K5: collector += estimate over clusters (susceptibility.h:182-198) and the Bernoulli(1/2) flip
// decision per cluster (path_integral.C:796-799).  Deterministic two-stage reduction
// (warp shuffles -> per-CTA partials -> one CTA).  The cluster arena is zeroed on the way.
// ------------------------------------------------------------------------------------------
#define LQ_NSUM 16   /* 14 susceptibility sums (susceptibility.h:158-160), transmag length (transmag.h:98), stiffness w2 */
#define LQ_NSUS 14
#define LQ_GEST_MAX 10   /* int64 fields per global open cluster: 4 sums, the 4 tau=0 sums packed in 2, [site-leg count],
                            [windings]; d.gstride = 6 + has_site + sdim of them travel in the all-reduce */
// two int32 sums in one int64 field, additive: x = hi * 2^32 + lo as INTEGERS (not bit fields), so the
// sum of packed values is the packed pair of sums as long as both stay inside int32
__device__ __forceinline__ unsigned long long pack2_i32(int hi, int lo) {
  return (unsigned long long)((long long)hi * 4294967296ll + (long long)lo);
}
__device__ __forceinline__ int unpack2_lo(long long x) { return (int)(unsigned)(unsigned long long)x; }
__device__ __forceinline__ int unpack2_hi(long long x, int lo) { return (int)((x - (long long)lo) >> 32); }
__global__ void __launch_bounds__(256)
k_collect(Dev d, double* partial) {
  __shared__ double s_red[8][LQ_NSUM];
  const uint32_t nc = d.d_nc[0], ncs = d.d_nc[1];
  double v[LQ_NSUM];
#pragma unroll
  for (int i = 0; i < LQ_NSUM; ++i) v[i] = 0;
  // persistent grid: every thread strides over the clusters and keeps its 14 running sums in
  // registers, so the warp-shuffle reduction runs once per thread instead of once per cluster
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nc && (long long)c < d.nccap;
       c += (size_t)gridDim.x * blockDim.x) {
    const double sc = d.sse ? 0.5 : 0.5 / LQ_FX;   // half units of 2^-40 (path integral) / of one string position (SSE)
    // read-and-clear with one L2 atomic per field: the sums were built by RED atomics, and on this
    // part plain loads of lines last touched by atomics are an order of magnitude slower than
    // atomics on them (measured: 0.11 ms -> 0.03 ms for 1.5e6 clusters, profiles/)
    unsigned long long* e = (unsigned long long*)d.est;
    // (zero_umag / zero_ssize: sums that the model keeps at zero for every cluster without a site node --
    // antiparallel legs only, bonds between opposite sublattices only -- are not read at all)
    const bool site_rooted = c < ncs;
    const double usize = sc * i64_to_f64((long long)atomicExch(e + 0 * d.nccap + c, 0ull));
    const double umag = (d.zero_umag && !site_rooted) ? 0.0 : sc * i64_to_f64((long long)atomicExch(e + 1 * d.nccap + c, 0ull));
    const double ssize = (d.zero_ssize && !site_rooted) ? 0.0 : sc * i64_to_f64((long long)atomicExch(e + 2 * d.nccap + c, 0ull));
    const double smag = sc * i64_to_f64((long long)atomicExch(e + 3 * d.nccap + c, 0ull));
    double usize0 = 0, umag0 = 0, ssize0 = 0, smag0 = 0;
    if (c < ncs) {
      usize0 = 0.5 * atomicExch(d.est0 + 0 * (size_t)d.N + c, 0); umag0 = 0.5 * atomicExch(d.est0 + 1 * (size_t)d.N + c, 0);
      ssize0 = 0.5 * atomicExch(d.est0 + 2 * (size_t)d.N + c, 0); smag0 = 0.5 * atomicExch(d.est0 + 3 * (size_t)d.N + c, 0);
    }
    // order = lq_collector: umag0 usize2 umag2 usize4 umag4 usize umag | smag0 ssize2 smag2 ssize4 smag4 ssize smag
    const double a2 = usize0 * usize0, b2 = umag0 * umag0, e2 = ssize0 * ssize0, g2 = smag0 * smag0;
    v[0] += umag0; v[1] += a2; v[2] += b2; v[3] += a2 * a2; v[4] += b2 * b2;
    v[5] += usize * usize; v[6] += umag * umag;
    v[7] += smag0; v[8] += e2; v[9] += g2; v[10] += e2 * e2; v[11] += g2 * g2;
    v[12] += ssize * ssize; v[13] += smag * smag;
    // transmag.h:98-101: only clusters cut by a site operator count, with their total length = 2 usize
    if (d.has_site && ((d.openw[c >> 5] >> (c & 31u)) & 1u)) v[14] += 2.0 * usize;
    for (int x = 0; x < d.sdim; ++x) {   // stiffness.h:125-128: w2 += (winding / 2)^2 per dimension
      const double w = d.wscale[x] * (double)atomicExch(d.wind + (size_t)x * d.nccap + c, 0);
      v[15] += w * w;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < LQ_NSUM; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[wid][i] = x;
  }
  __syncthreads();
  if (threadIdx.x < LQ_NSUM) {
    double x = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) x += s_red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * LQ_NSUM + threadIdx.x] = x;
  }
}

// one CTA: sum the per-CTA partials in a fixed order -> out[0..13]; out[14] = nc, out[15] = nop
__global__ void __launch_bounds__(256)
k_collect_final(Dev d, const double* partial, size_t nblk_cap, double* out) {
  __shared__ double s_red[8][LQ_NSUM];
  const uint32_t nc = d.d_nc[0];
  const size_t nblk = nblk_cap;
  double v[LQ_NSUM];
#pragma unroll
  for (int i = 0; i < LQ_NSUM; ++i) v[i] = 0;
  for (size_t k = threadIdx.x; k < nblk; k += blockDim.x)
#pragma unroll
    for (int i = 0; i < LQ_NSUM; ++i) v[i] += partial[k * LQ_NSUM + i];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < LQ_NSUM; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[wid][i] = x;
  }
  __syncthreads();
  if (threadIdx.x < LQ_NSUM) {
    double x = 0;
    for (int w = 0; w < 8; ++w) x += s_red[w][threadIdx.x];
    out[threadIdx.x < LQ_NSUS ? threadIdx.x : threadIdx.x + 4] = x;   // slot 18: transmag length, 19: stiffness w2
  }
  if (threadIdx.x == 0) {
    out[14] = (double)nc;
    out[15] = (double)d.nbase[d.Pown];   // operators this rank owns (all of them on a serial engine)
    out[16] = (double)(*d.d_err);
    out[17] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------
// SSE representation (sse.C:168-407): the operator string is the time-ordered list of the operators
// and the estimators use the POSITION of an operator in it as its time (sse.C:251-283: `t`,
// stop_top(operators.size())).  Positions come from one counting sort over (window, time bin):
// histogram -> exclusive scan -> scatter -> rank inside the bin by comparing (time, bond) with the
// handful of operators that share it.  Ties in time are ordered by the internal bond index, then by the
// position inside the bucket: the order lq_get_state exports.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sse_bin(const Dev& d, int wl, double t) {
  const int wg = d.w0 + wl;
  const double tlo = d.wlo[wg], width = d.wlo[wg + 1] - tlo;
  int b = (int)((t - tlo) / width * (double)d.nbin);
  b = b < 0 ? 0 : (b >= d.nbin ? d.nbin - 1 : b);
  return (uint32_t)wl * (uint32_t)d.nbin + (uint32_t)b;
}

__global__ void __launch_bounds__(256)
k_sse_hist(Dev d, int buf) {
  const size_t p = blockIdx.x;
  const int wl = (int)(blockIdx.x % (unsigned)d.Wl);
  const int n = d.pcount[buf][p];
  const double* gt = d.time[buf] + p * (size_t)d.cap;
  for (int j = threadIdx.x; j < n; j += blockDim.x) atomicAdd(d.bincnt + sse_bin(d, wl, gt[j]), 1u);
}

__global__ void __launch_bounds__(256)
k_sse_scatter(Dev d, int buf) {
  const size_t p = blockIdx.x;
  const int t = (int)(blockIdx.x / (unsigned)d.Wl), wl = (int)(blockIdx.x - (unsigned)t * (unsigned)d.Wl);
  const int n = d.pcount[buf][p];
  const int idx0 = d.nbase[p];
  const int b0 = d.bond_base[t];
  const double* gt = d.time[buf] + p * (size_t)d.cap;
  const uint32_t* gi = d.info[buf] + p * (size_t)d.cap;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double tt = gt[j];
    const uint32_t g = sse_bin(d, wl, tt);
    const uint32_t slot = d.binbase[g] + atomicAdd(d.binfill + g, 1u);
    d.sorted_time[slot] = tt;
    d.sorted_id[slot] = make_uint2((uint32_t)(b0 + (int)(gi[j] >> LQ_INFO_LBSHIFT)), (uint32_t)(idx0 + j));
  }
}

// one thread per slot of the (window, bin)-ordered list; bins hold a few operators each
__global__ void __launch_bounds__(256)
k_sse_rank(Dev d) {
  const size_t nbins = (size_t)d.Wl * d.nbin;
  const uint32_t n = (uint32_t)(*d.d_ntotal);
  // the bin of a slot: every thread walks the bins of its CTA's slot range once (binary search start)
  const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  size_t lo = 0, hi = nbins;   // last bin with binbase <= s
  while (hi - lo > 1) {
    const size_t mid = (lo + hi) >> 1;
    if (d.binbase[mid] <= s) lo = mid; else hi = mid;
  }
  const uint32_t beg = d.binbase[lo], end = d.binbase[lo + 1];
  const double tt = d.sorted_time[s];
  const uint2 me = d.sorted_id[s];
  uint32_t r = beg;
  for (uint32_t k = beg; k < end; ++k) {
    const double tk = d.sorted_time[k];
    bool before = tk < tt;
    if (tk == tt) { const uint2 o = d.sorted_id[k]; before = o.x < me.x || (o.x == me.x && o.y < me.y); }   // then bond, then bucket order
    r += before ? 1u : 0u;
  }
  d.spos[me.y] = r;
}

// ------------------------------------------------------------------------------------------
// Graph-launched steps (small systems, where 20 launches per step are a visible share of a step
// that takes a few hundred microseconds): the captured kernels read their inputs from ONE fixed
// record and write the collector to ONE fixed slot; this last node of the graph files the slot in
// the result ring of the batch and advances the step counter on the device.
// ------------------------------------------------------------------------------------------
struct StepCtl {
  StepParams sp;   // inputs of the step being run
  int slot;        // next free entry of the result ring
  int pad;
  double* ring;    // [slots][32] collectors of the batch
};

__global__ void k_step_advance(StepCtl* ctl, const double* fixed_out) {
  double* dst = ctl->ring + (size_t)ctl->slot * 32;
  dst[threadIdx.x] = fixed_out[threadIdx.x];
  __syncthreads();
  if (threadIdx.x == 0) { ctl->sp.mcs += 1u; ctl->slot += 1; }
}

// ------------------------------------------------------------------------------------------
// K6: flip (path_integral.C:815-823).  An operator changes between diagonal and off-diagonal iff
// the cluster arriving from below on the source side and the one leaving upwards there are
// flipped differently (loop_0 / loop_1 of graph_impl.h:277-295 in leg form).  The spin carried
// into every window flips with the cluster of the segment crossing the window start.
// ------------------------------------------------------------------------------------------
__global__ void k_flip_spins(Dev d) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)(d.Wl + 1) * d.N) return;
  if (d.space && (int)(i % (size_t)d.N) >= d.Nown) return;   // ghost sites: their owners send the spins
  if (*d.d_err) return;   // an arena overflowed in this step: keep the spins of the configuration it started from
  const uint32_t c = d.parent[d.curW[i]];
  d.spinW[i] ^= (uint8_t)flip_of_label(d, c);
}

// ------------------------------------------------------------------------------------------
// parity-test helper: cluster ids of the upper legs of every operator in page order
// ------------------------------------------------------------------------------------------
__global__ void k_export_labels(Dev d, int buf, uint32_t* out /* [2*ncap-ish], by dense idx */) {
  const size_t p = blockIdx.x;
  const int n = d.pcount[buf][p];
  const int idx0 = d.nbase[p];
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int idx = idx0 + j;
    out[2 * (size_t)idx] = LQ_CID(d.parent[upper_node(d, idx, 0)]);
    out[2 * (size_t)idx + 1] = LQ_CID(d.parent[upper_node(d, idx, 1)]);
  }
}

// ==========================================================================================
// Multi-GPU: imaginary-time slabs (path_integral_mpi.C:231-232).  Rank r owns the windows
// [w0, w0+Wl) of ALL sites; its bottom boundary nodes are the site nodes s (ids [0,N)), its top
// boundary nodes are curW[Wl][s] (path_integral_mpi.C:478-479,662).  After the local labelling
// every rank publishes, per site, the "open id" of the cluster touching its bottom and its top
// boundary (the 2N links of a chunk, looper/parallel.h:139-229); the P x 2N ids are all-gathered
// and EVERY rank unifies top(r) with bottom(r+1) redundantly (parallel.h:524-531,1739) -- on
// NVSwitch one all-gather replaces the staged ring shifts of parallel.h:1675-1789.  Partial
// estimates of open clusters are summed with one integer all-reduce indexed by global cluster id;
// flip bits of open clusters are Philox draws keyed by that id, identical on all ranks.
// ==========================================================================================
struct MrDev {
  uint32_t* topmin;   // [nccap] min boundary site of a local cluster (NONE if it touches none)
  uint32_t* sendb;    // [2N]  botoid[s], topoid[s]
  uint32_t* recvb;    // [P*2N]
  uint32_t* gparent;  // [P*2N]
  uint8_t* gused;     // [P*2N]
  uint32_t* gbitmap;  // [(P*2N)/32 + 1]
  uint32_t* gwcount;
  uint32_t* gwbase;
  long long* gest;    // [gcap][8]
  uint32_t* d_g;      // [0] ngc (global open clusters), [1] noc_local
  double* rankvec;    // [32] this rank's closed-cluster sums
  double* allvec;     // [P*32]
  double* gsum;       // [16] sums over global clusters
  uint32_t gcap;      // open-cluster slots that travel in this step's all-reduce (the host sizes it from the
                      // count of two steps ago, so that no step waits for a read-back; see merge_open_clusters)
  size_t gn;          // nodes of the gathered boundary forest (slabs: P*2N; spatial cut: P*stride)
  size_t gbase;       // first node of this rank in it
};

// Slab engines run the exchange BEFORE the relabelling (so that k_relabel can pack the flips of the
// open clusters, known only after the exchange, into the labels like a serial engine does per root):
// until then parent[] holds roots, and the cluster id of a node is computed from its root.
__device__ __forceinline__ uint32_t cid_pre(const Dev& d, size_t x) { return cid_of_root(d, d.parent[x]); }

__global__ void k_set_ncs(Dev d) {
  d.d_nc[1] = cid_of_root(d, (node_t)d.N);   // clusters rooted at a site node come first
  if ((long long)d.d_nc[0] > d.nccap) atomicOr(d.d_err, LQ_ERR_CLUSTER_FULL);
}

__global__ void k_mr_topmin(Dev d, MrDev m) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.N) return;
  const uint32_t cb = cid_pre(d, s);
  const uint32_t ct = cid_pre(d, d.curW[(size_t)d.Wl * d.N + s]);
  // bottom-touching clusters (cid < ncs) are represented by their smallest BOTTOM site,
  // clusters that only touch the top boundary by their smallest top site
  if ((long long)cb < d.nccap) atomicMin(m.topmin + cb, (uint32_t)s);
  if (ct >= d.d_nc[1] && (long long)ct < d.nccap) atomicMin(m.topmin + ct, (uint32_t)s);
}

// open id of a local cluster: its cid if it touches the bottom boundary (cid < ncs), else
// N + (smallest top site)
__device__ __forceinline__ uint32_t open_id(const Dev& d, const MrDev& m, uint32_t cid) {
  if (cid < d.d_nc[1]) return cid;
  return (uint32_t)d.N + (((long long)cid < d.nccap) ? m.topmin[cid] : 0u);
}

__global__ void k_mr_ids(Dev d, MrDev m) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.N) return;
  // bit 31 of entry 0 carries "an arena of this rank overflowed in this step": every rank learns it
  // with the boundary ids, before anything is flipped, so that all ranks rewind together
  m.sendb[s] = cid_pre(d, s) | ((s == 0 && *d.d_err) ? 0x80000000u : 0u);
  m.sendb[d.N + s] = open_id(d, m, cid_pre(d, d.curW[(size_t)d.Wl * d.N + s]));
}

__global__ void k_mr_ginit(Dev d, MrDev m) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)d.nranks * 2 * d.N) return;
  m.gparent[i] = (uint32_t)i;
  m.gused[i] = 0;
}

__global__ void k_mr_gunion(Dev d, MrDev m) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)d.nranks * d.N) return;
  const int r = (int)(i / d.N), s = (int)(i % d.N), rn = (r + 1) % d.nranks;
  const size_t N2 = 2 * (size_t)d.N;
  const uint32_t a = (uint32_t)(r * N2 + m.recvb[r * N2 + d.N + s]);   // top of slab r
  const uint32_t b = (uint32_t)(rn * N2 + (m.recvb[rn * N2 + s] & 0x7fffffffu));       // bottom of slab r+1
  if (s == 0 && (m.recvb[rn * N2] >> 31)) atomicOr(d.d_err, LQ_ERR_REMOTE);   // (see k_mr_ids)
  m.gused[a] = 1;
  m.gused[b] = 1;
  uf_union(m.gparent, a, b);
}

__global__ void k_mr_gcompress(Dev d, MrDev m) {
  const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nn = m.gn;
  if ((x >> 5) >= ((nn + 31) >> 5)) return;
  bool isroot = false;
  if (x < nn) {
    node_t r = (node_t)x, pr = uf_load(m.gparent + r);
    while (pr != r) { r = pr; pr = uf_load(m.gparent + r); }
    m.gparent[x] = r;
    isroot = (r == (node_t)x) && m.gused[x];
  }
  const uint32_t word = __ballot_sync(0xffffffffu, isroot);
  if ((threadIdx.x & 31) == 0) { m.gbitmap[x >> 5] = word; m.gwcount[x >> 5] = (uint32_t)__popc(word); }
}

__device__ __forceinline__ uint32_t global_cid(const Dev& d, const MrDev& m, uint32_t oid) {
  const uint32_t r = m.gparent[m.gbase + oid];
  return m.gwbase[r >> 5] + (uint32_t)__popc(m.gbitmap[r >> 5] & ((1u << (r & 31)) - 1u));
}

// representative threads move the partial sums of their open cluster into the global table
// (collect_estimates, parallel.h:415-427) and clear the local entry
// sum of a 64-bit value over the lanes of `grp` (two's complement, mod 2^64): three 24-bit chunks
// through the integer warp reduction
__device__ __forceinline__ unsigned long long group_sum64(unsigned grp, unsigned long long v) {
  const unsigned a = __reduce_add_sync(grp, (unsigned)(v & 0xffffffull));
  const unsigned b = __reduce_add_sync(grp, (unsigned)((v >> 24) & 0xffffffull));
  const unsigned c = __reduce_add_sync(grp, (unsigned)(v >> 48));
  return (unsigned long long)a + ((unsigned long long)b << 24) + ((unsigned long long)c << 48);
}

// A global loop passes through a slab many times: its local pieces are different local clusters with
// the SAME global id, and at low temperature one macroscopic loop owns ~1e5 of them per slab.  Their
// representatives are first combined per warp (match.any on the global id + warp reduction), so one
// atomic per warp, global cluster and field reaches the table instead of one per piece (the plain
// version spent 1.1 ms serialised on the few addresses of that loop, profiles/r02_slabs.md).
#define LQ_GATHER_SLOTS 64
__global__ void __launch_bounds__(1024)
k_mr_gather(Dev d, MrDev m) {
  // second level: the warp leaders of a CTA meet in a small shared-memory table keyed by the global
  // id (colliding keys go straight to the global table), flushed with one atomic per slot and field
  __shared__ uint32_t s_key[LQ_GATHER_SLOTS];
  __shared__ unsigned long long s_val[LQ_GATHER_SLOTS][LQ_GEST_MAX];
  for (int i = threadIdx.x; i < LQ_GATHER_SLOTS * LQ_GEST_MAX; i += blockDim.x) {
    if (i < LQ_GATHER_SLOTS) s_key[i] = 0xffffffffu;
    (&s_val[0][0])[i] = 0ull;
  }
  __syncthreads();
  const int s0 = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = s0 < d.N;
  const int s = valid ? s0 : 0;
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t ncs = d.d_nc[1];
  uint32_t cl[2];
  cl[0] = LQ_CID(d.parent[s]);   // (runs after k_relabel: labels = cluster id | flip << 31)
  cl[1] = LQ_CID(d.parent[d.curW[(size_t)d.Wl * d.N + s]]);
  unsigned nrep = 0;
  for (int k = 0; k < 2; ++k) {
    const uint32_t c = cl[k];
    // representative of its local open cluster?  (bottom-touching clusters go through their bottom site)
    const bool rep = valid && !(k == 1 && c == cl[0]) && (long long)c < d.nccap && m.topmin[c] == (uint32_t)s &&
                     !(k == 1 && c < ncs);
    uint32_t gid = 0xffffffffu;
    unsigned long long v[LQ_GEST_MAX];
#pragma unroll
    for (int f = 0; f < LQ_GEST_MAX; ++f) v[f] = 0ull;
    if (rep) {
      gid = global_cid(d, m, open_id(d, m, c));
#pragma unroll
      for (int f = 0; f < 4; ++f) v[f] = atomicExch((unsigned long long*)d.est + f * d.nccap + c, 0ull);
      if (c < ncs) {
        int z[4];
#pragma unroll
        for (int f = 0; f < 4; ++f) z[f] = atomicExch(d.est0 + f * (size_t)d.N + c, 0);
        v[4] = pack2_i32(z[0], z[1]);   // the four tau = 0 sums (|x| <= N) travel as two fields
        v[5] = pack2_i32(z[2], z[3]);
      }
      if (d.has_site && ((d.openw[c >> 5] >> (c & 31u)) & 1u)) v[6] = 1ull;
      for (int x = 0; x < d.sdim; ++x) v[6 + d.has_site + x] = (unsigned long long)(long long)atomicExch(d.wind + (size_t)x * d.nccap + c, 0);
      ++nrep;
      if (gid >= m.gcap) {   // more open clusters than the all-reduce of this step carries: the step is lost
        atomicOr(d.d_err, LQ_ERR_OPEN_FULL);
        gid = 0u;
#pragma unroll
        for (int f = 0; f < LQ_GEST_MAX; ++f) v[f] = 0ull;
      }
    }
    const unsigned grp = __match_any_sync(0xffffffffu, gid);
    const bool leader = rep && lane == (unsigned)(__ffs(grp) - 1);
    unsigned long long* ge = (unsigned long long*)m.gest + (size_t)(rep ? gid : 0u) * d.gstride;
    int slot = -1;
    if (leader) {
      const int h = (int)((gid * 2654435761u) >> 26);   // 6 bits
      const uint32_t old = atomicCAS(&s_key[h], 0xffffffffu, gid);
      if (old == 0xffffffffu || old == gid) slot = h;
    }
    for (int f = 0; f < d.gstride; ++f) {
      const unsigned long long t = group_sum64(grp, v[f]);
      if (leader && t) {
        if (slot >= 0) atomicAdd(&s_val[slot][f], t);
        else atomicAdd(ge + f, t);
      }
    }
  }
  nrep = __reduce_add_sync(0xffffffffu, nrep);
  if (lane == 0 && nrep) atomicAdd(m.d_g + 1, nrep);
  __syncthreads();
  for (int i = threadIdx.x; i < LQ_GATHER_SLOTS * d.gstride; i += blockDim.x) {
    const int slot = i / d.gstride, f = i - slot * d.gstride;
    const uint32_t gid = s_key[slot];
    const unsigned long long t = s_val[slot][f];
    if (gid != 0xffffffffu && t) atomicAdd((unsigned long long*)m.gest + (size_t)gid * d.gstride + f, t);
  }
}

// flip bits of open clusters: one Philox draw per GLOBAL cluster id (same on every rank)
__global__ void k_mr_openflips(Dev d, MrDev m, const StepParams* __restrict__ sp) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.N) return;
  const uint32_t ncs = d.d_nc[1];
  uint32_t cl[2];
  cl[0] = cid_pre(d, s);   // (runs before k_relabel)
  cl[1] = cid_pre(d, d.curW[(size_t)d.Wl * d.N + s]);
  for (int k = 0; k < 2; ++k) {
    const uint32_t c = cl[k];
    if (k == 1 && c == cl[0]) break;
    if ((long long)c >= d.nccap || m.topmin[c] != (uint32_t)s) continue;
    if (k == 1 && c < ncs) continue;
    const uint32_t gid = global_cid(d, m, open_id(d, m, c));
    philox_t x = philox4x32_10(gid, 0xffffffffu, sp->mcs, LQ_STREAM_FLIP, sp->key0, sp->key1);
    if (x.x & 1u) atomicOr(d.flipw + (c >> 5), 1u << (c & 31u));
    else atomicAnd(d.flipw + (c >> 5), ~(1u << (c & 31u)));
  }
}

__global__ void k_mr_reset_topmin(Dev d, MrDev m) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.N) return;
  const uint32_t cb = LQ_CID(d.parent[s]);   // (runs after k_relabel)
  const uint32_t ct = LQ_CID(d.parent[d.curW[(size_t)d.Wl * d.N + s]]);
  if ((long long)cb < d.nccap) m.topmin[cb] = 0xffffffffu;
  if ((long long)ct < d.nccap) m.topmin[ct] = 0xffffffffu;
}

// sums over the all-reduced global clusters: persistent grid -> per-CTA partials (fixed assignment
// of clusters to threads, so the result is reproducible), then k_mr_gsum; clears the table
__global__ void __launch_bounds__(256)
k_mr_gcollect(Dev d, MrDev m, double* partial) {
  __shared__ double s_red[8][LQ_NSUM];
  const uint32_t ngc = min(m.d_g[0], m.gcap);
  double v[LQ_NSUM];
#pragma unroll
  for (int i = 0; i < LQ_NSUM; ++i) v[i] = 0;
  const double sc = 0.5 / LQ_FX;
  for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < ngc; c += (size_t)gridDim.x * blockDim.x) {
    long long* ge = m.gest + c * d.gstride;
    const double usize = sc * i64_to_f64(ge[0]), umag = sc * i64_to_f64(ge[1]);
    const double ssize = sc * i64_to_f64(ge[2]), smag = sc * i64_to_f64(ge[3]);
    const int u0 = unpack2_lo(ge[4]), u1 = unpack2_lo(ge[5]);
    const double usize0 = 0.5 * (double)unpack2_hi(ge[4], u0), umag0 = 0.5 * (double)u0;
    const double ssize0 = 0.5 * (double)unpack2_hi(ge[5], u1), smag0 = 0.5 * (double)u1;
    if (d.has_site && ge[6] > 0) v[14] += 2.0 * usize;
    for (int x = 0; x < d.sdim; ++x) { const double w = d.wscale[x] * i64_to_f64(ge[6 + d.has_site + x]); v[15] += w * w; }
    for (int f = 0; f < d.gstride; ++f) ge[f] = 0;
    const double a = usize0 * usize0, b = umag0 * umag0, e = ssize0 * ssize0, g = smag0 * smag0;
    v[0] += umag0; v[1] += a; v[2] += b; v[3] += a * a; v[4] += b * b; v[5] += usize * usize; v[6] += umag * umag;
    v[7] += smag0; v[8] += e; v[9] += g; v[10] += e * e; v[11] += g * g; v[12] += ssize * ssize; v[13] += smag * smag;
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < LQ_NSUM; ++i) {
    double x = v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) s_red[wid][i] = x;
  }
  __syncthreads();
  if (threadIdx.x < LQ_NSUM) {
    double x = 0;
    for (int w = 0; w < 8; ++w) x += s_red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * LQ_NSUM + threadIdx.x] = x;
  }
}

__global__ void k_mr_gsum(MrDev m, const double* partial, int nblk) {   // LQ_NSUM warps: one sum each, fixed order
  const int i = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double x = 0;
  for (int k = lane; k < nblk; k += 32) x += partial[(size_t)k * LQ_NSUM + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  if (lane == 0) m.gsum[i] = x;
}

// The collectors of the closed clusters travel in the SAME all-reduce as the open-cluster sums: every
// rank writes its 32 doubles into its own slot of the tail of the table and zeros the other slots, so
// the integer sum of the bit patterns is the pattern itself (x + 0 + ... + 0) -- one collective less
// per step.  tail[r * 32 + i]; slot = this rank's k_collect_final output.
__global__ void k_mr_rankvec(Dev d, MrDev m, const double* slot, double* tail) {
  const int i = threadIdx.x;   // 32 threads
  for (int r = 0; r < d.nranks; ++r) {
    double x = 0;
    if (r == d.rank) {
      if (i < LQ_NSUS) x = slot[i];
      if (i == 18 || i == 19) x = slot[i];              // transmag length / stiffness w2 of the closed clusters
      if (i == 14) x = slot[14] - (double)m.d_g[1];     // closed clusters of this rank
      if (i == 15) x = slot[15];                        // operators of this slab
      if (i == 16) x = slot[16];                        // error flags
    }
    tail[r * 32 + i] = x;
  }
}

// after the all-reduce: out = sum over ranks (fixed order) + global-cluster sums.  The tail is
// zeroed again: the next step's open-cluster table may reach into it.
__global__ void k_mr_final(Dev d, MrDev m, double* tail, double* slot) {
  const int i = threadIdx.x;
  if (i >= 32) return;
  double x = 0;
  int err = 0;   // slot 16: error bits, OR-ed (not summed) over the ranks
  for (int r = 0; r < d.nranks; ++r) { x += tail[r * 32 + i]; err |= (int)tail[r * 32 + i]; tail[r * 32 + i] = 0.0; }
  if (i == 16) x = (double)err;
  if (i < LQ_NSUS) x += m.gsum[i];
  if (i == 18) x += m.gsum[14];
  if (i == 19) x += m.gsum[15];
  if (i == 14) x += (double)m.d_g[0];
  if (i == 17) x = (double)m.d_g[0];  // number of clusters that were open (diagnostic)
  slot[i] = x;
}

}  // namespace lq
