// lq_k1.cuh -- K1, the diagonal update (path_integral.C:403-425,484-537 / standalone/loop.C:93-115).
//
// One CTA per page (tile x window), staged with its halo in shared memory.  The reference walks
// imaginary time carrying spins_c[s]; here every candidate is decided independently from the
// spin at the window start and the OFF-DIAGONAL operators of the window, which the diagonal update
// never changes.  FLAT work mapping -- every phase gives a thread one bucket, one staged operator
// or one candidate (thread-per-bucket loops over data-dependent lists run with 4-7 of 32 lanes
// active on this workload, measured; see profiles/):
//   1  per bucket   halo bucket extent; number of candidates K ~ Poisson(beta v dtau) by inverse
//      CDF (poisson_distribution.h:60-75 + the exponential gaps of path_integral.C:413-423 in one
//      step); ONE packed CTA prefix sum -> halo slots and candidate slots
//   2  stage the page (coalesced) and its halo; per K-site spin at the window start
//   3  per staged operator: off-diagonal legs are appended to the fixed-width, 2.0-padded column
//      of their K-site in `flist` (one shared-memory atomic per leg); kept own operators are
//      ballot-compacted
//   4  per candidate: uniform time, Philox4x32-10 keyed by (bond, window, step, i); is_compatible
//      (graph_impl.h:257) from the two spins at that time = start spins xor parity of the earlier
//      off-diagonal legs -- FC branch-free compares per site; graph by the model's weights
//      (graph_impl.h:679).  Site-graph candidates (graph_impl.h:67-87) are always accepted.
//   5  per bucket   new size = kept + accepted -> CTA prefix sum -> new bucket offsets
//   6  per accepted candidate / kept operator: rank inside the new bucket -> scatter into the
//      compacted new page (old diagonal operators are dropped, path_integral.C:519-521)
#pragma once
#include "lq_device.cuh"

namespace lq {

// template parameter FC (8, 12, 16): off-diagonal legs per K-site and window held in the fast list

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

struct K1Smem {
  double* time;     // [scap]  staged operators (own page first, halo buckets behind)
  double* ctime;    // [ccap]  candidate times
  double* flist;    // [FC][nksp] off-diagonal leg times per K-site, 2.0-padded
  uint32_t* info;   // [scap]  (halo copies carry the LOCAL bucket id of this tile in the bond bits)
  int* off;         // [nloc+1] first staged slot of each local bucket
  int* cbase;       // [nbmax+1] first candidate of each own bucket
  int* noff;        // [nbmax+1] new bucket offsets
  int* fcnt;        // [nksp]  off-diagonal legs on a K-site (may exceed FC -> slow path)
  int* nkb;         // [nbmax] kept operators per own bucket
  uint16_t* clb;    // [ccap]  owning local bucket of a candidate
  uint16_t* klist;  // [cap]   staged slots of the kept own operators, compacted
  uint16_t* alist;  // [ccap]  accepted candidates, compacted
  uint8_t* cacc;    // [ccap]  accepted bit | graph << 1
  uint8_t* kspin;   // [nksp]
};

__host__ __device__ inline int k1_nksp(int nksmax) { return (nksmax + 31) & ~31; }

__host__ __device__ inline size_t k1_smem_bytes(int fc, int scap, int ccap, int cap, int nbmax, int hmax, int nksmax) {
  const size_t nloc = (size_t)nbmax + hmax, nksp = k1_nksp(nksmax);
  return ((size_t)scap + ccap + (size_t)fc * nksp) * 8 + (size_t)scap * 4 +
         (nloc + 1 + 3 * ((size_t)nbmax + 1) + nksp) * 4 + (2 * (size_t)ccap + cap) * 2 + (size_t)ccap + nksp + 64;
}

__device__ __forceinline__ void k1_carve(const Dev& d, int fc, unsigned char* smem, K1Smem& S) {
  const size_t nloc = (size_t)d.nbmax + d.hmax, nksp = k1_nksp(d.nksmax);
  S.time = (double*)smem;
  S.ctime = S.time + d.scap;
  S.flist = S.ctime + d.ccap;
  S.info = (uint32_t*)(S.flist + (size_t)fc * nksp);
  S.off = (int*)(S.info + d.scap);
  S.cbase = S.off + nloc + 1;
  S.noff = S.cbase + d.nbmax + 1;
  S.fcnt = S.noff + d.nbmax + 1;
  S.nkb = S.fcnt + nksp;
  S.clb = (uint16_t*)(S.nkb + d.nbmax + 1);
  S.klist = S.clb + d.ccap;
  S.alist = S.klist + d.cap;
  S.cacc = (uint8_t*)(S.alist + d.ccap);
  S.kspin = S.cacc + d.ccap;
}

// parity of the off-diagonal legs before tc on K-site k, straight from the staged buckets
// (only used when a site carries more than FC legs in one window)
__device__ __noinline__ int k1_parity_slow(const double* time, const uint32_t* info, const int* off,
                                           const int* sso, const int* sse, int k, double tc) {
  int par = 0;
  for (int e = sso[k]; e < sso[k + 1]; ++e) {
    const int lid = sse[e] >> 1;
    for (int j = off[lid]; j < off[lid + 1]; ++j)
      par ^= (int)(info[j] & LQ_INFO_OFFDIAG) & (int)(time[j] < tc);
  }
  return par;
}

template <int MAXT, int FC>
__global__ void __launch_bounds__(MAXT, (MAXT <= 192 ? 6 : (MAXT <= 320 ? 4 : (MAXT <= 576 ? 2 : 1))))
k_diag_update(Dev d, int src, const StepParams* __restrict__ sp) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ int s_scan[34];
  __shared__ int s_cnt[2];
  if (*d.d_err) return;   // an earlier step of this batch overflowed: leave both page buffers alone
  K1Smem S;
  k1_carve(d, FC, s_raw, S);
  const double beta = sp->beta;
  const uint32_t key0 = sp->key0, key1 = sp->key1, mcs = sp->mcs;
  const int dst = src ^ 1;
  const size_t p = blockIdx.x;
  const int t = (int)(blockIdx.x / (unsigned)d.Wl), wl = (int)(blockIdx.x - (unsigned)t * (unsigned)d.Wl), wg = d.w0 + wl;  // (32-bit division)
  const int tid = threadIdx.x;
  const unsigned lane = tid & 31u;
  const int b0 = d.bond_base[t];
  const int nb = d.bond_base[t + 1] - b0;
  const int h0 = d.halo_off[t];
  const int nh = d.halo_off[t + 1] - h0;
  const int nksp = k1_nksp(d.nksmax);
  uint16_t* bo_new = d.boff[dst] + p * (size_t)(d.nbmax + 1);
  const double tlo = d.wlo[wg], thi = d.wlo[wg + 1], width = thi - tlo;   // = window_lo / window_hi (host table)
  const int n_own = d.pcount[src][p];

  // ---- 1: halo bucket extents and candidate counts, one packed prefix sum ----------------------
  size_t hbase = 0;
  int hn = 0;
  if (tid < nh) {
    const int tl = d.bond_tl[d.halo_bond[h0 + tid]];
    const size_t p2 = (size_t)(tl >> 10) * d.Wl + wl;
    const uint16_t* bo2 = d.boff[src] + p2 * (size_t)(d.nbmax + 1) + (tl & 1023);
    const int o0 = bo2[0];
    hbase = p2 * (size_t)d.cap + o0;
    hn = bo2[1] - o0;
  }
  int K = 0;
  if (tid < nb) {
    const int b = b0 + tid;
    const double mu = beta * d.bond_rate[b] * width;
    if (mu > 0) {
      const philox_t x = philox4x32_10((uint32_t)b, (uint32_t)wg, mcs, LQ_STREAM_CAND, key0, key1);
      bool ovf;
      K = k1_poisson(u53(x.x, x.y), d.bond_emu[b], mu, &ovf);
      if (ovf) atomicOr(d.d_err, LQ_ERR_CAND_FULL);
    }
  }
  int packed_total;   // both sums stay below 2^16 (scap, ccap <= 65535, checked by the host)
  const int packed = block_exscan((K << 16) | hn, &packed_total, s_scan);
  const int hoff = packed & 0xffff, cb = packed >> 16;
  const int n_all = n_own + (packed_total & 0xffff), C = packed_total >> 16;
  if (n_all > d.scap || C > d.ccap) {
    if (tid == 0) { atomicOr(d.d_err, n_all > d.scap ? LQ_ERR_PAGE_FULL : LQ_ERR_CAND_FULL); d.pcount[dst][p] = 0; }
    if (tid < nb) bo_new[tid] = 0;
    if (tid == 0) bo_new[nb] = 0;
    return;
  }

  // ---- 2: stage the page and its halo; K-site spins; clear the lists ---------------------------
  const int cls = d.tile_class[t];
  const int nks = d.cls_nks[cls];
  const int ns = d.site_base[t + 1] - d.site_base[t];
  const int* sso = d.sst_off + d.cls_sso[cls];
  const int* sse = d.sst + d.cls_sst[cls];
  const int* bsx = d.bs + d.cls_bs[cls];
  {
    const uint16_t* bo = d.boff[src] + p * (size_t)(d.nbmax + 1);
    const double* gt = d.time[src] + p * (size_t)d.cap;
    const uint32_t* gi = d.info[src] + p * (size_t)d.cap;
    for (int j = tid; j < n_own; j += blockDim.x) { S.time[j] = gt[j]; S.info[j] = gi[j]; }
    if (tid < nb) { S.off[tid] = bo[tid]; S.nkb[tid] = 0; S.cbase[tid] = cb; }
    if (tid == 0) S.cbase[nb] = C;
    if (tid < nh) {
      S.off[nb + tid] = n_own + hoff;
      // batches of four: all loads of a batch are in flight before the first shared-memory store
      for (int j0 = 0; j0 < hn; j0 += 4) {
        double tt[4];
        uint32_t ii[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u < hn) { tt[u] = d.time[src][hbase + j0 + u]; ii[u] = d.info[src][hbase + j0 + u]; }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (j0 + u < hn) {
            S.time[n_own + hoff + j0 + u] = tt[u];
            S.info[n_own + hoff + j0 + u] = (ii[u] & ((1u << LQ_INFO_LBSHIFT) - 1u)) | ((uint32_t)(nb + tid) << LQ_INFO_LBSHIFT);
          }
      }
    }
    if (tid == 0) { S.off[nb + nh] = n_all; s_cnt[0] = 0; s_cnt[1] = 0; }
    if (tid < nks) {
      const int sg = tid < ns ? d.site_base[t] + tid : d.hsite[d.hsite_off[t] + tid - ns];
      S.kspin[tid] = d.spinW[(size_t)wl * d.N + sg];
      S.fcnt[tid] = 0;
    }
    for (int i = tid; i < FC * nksp; i += blockDim.x) S.flist[i] = 2.0;
    if (tid < nb)
      for (int i = 0; i < K; ++i) S.clb[cb + i] = (uint16_t)tid;
  }
  __syncthreads();

  // ---- 3: off-diagonal legs -> K-site columns; kept own operators compacted --------------------
  for (int j0 = 0; j0 < n_all; j0 += blockDim.x) {
    const int j = j0 + tid;
    const uint32_t inf = (j < n_all) ? S.info[j] : 0u;
    const bool offd = (inf & LQ_INFO_OFFDIAG) != 0;
    if (offd) {
      const int lid = (int)(inf >> LQ_INFO_LBSHIFT);
      const int k0 = bsx[2 * lid], k1 = bsx[2 * lid + 1];
      const double tt = S.time[j];
      if (k0 >= 0) { const int f = atomicAdd(&S.fcnt[k0], 1); if (f < FC) S.flist[f * nksp + k0] = tt; }
      if (k1 >= 0) { const int f = atomicAdd(&S.fcnt[k1], 1); if (f < FC) S.flist[f * nksp + k1] = tt; }
      if (j < n_own) atomicAdd(&S.nkb[lid], 1);
    }
    const bool keep = offd && j < n_own;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(&s_cnt[0], __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (keep) S.klist[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)j;
  }
  __syncthreads();

  // ---- 4: time, acceptance and graph of every candidate ----------------------------------------
  for (int c0 = 0; c0 < C; c0 += blockDim.x) {
    const int c = c0 + tid;
    bool accepted = false;
    if (c < C) {
      const int lb = S.clb[c];
      const int i = c - S.cbase[lb];
      const int b = b0 + lb;
      const philox_t x = philox4x32_10((uint32_t)b, (uint32_t)wg, mcs, LQ_STREAM_CAND + 1u + (uint32_t)i, key0, key1);
      double tc = tlo + (u53(x.x, x.y) - 1.0 / 9007199254740992.0) * width;
      if (!(tc < thi)) tc = tlo;
      const int k0 = bsx[2 * lb], k1 = bsx[2 * lb + 1];
      int g = 0;
      if (k1 >= 0) {   // bond graph; k1 < 0: site graph, compatible with any spin (graph_impl.h:69)
        int par = S.kspin[k0] ^ S.kspin[k1];   // operators on this bond sit in both lists and cancel
#pragma unroll
        for (int f = 0; f < FC; ++f)
          par ^= (int)(S.flist[f * nksp + k0] < tc) ^ (int)(S.flist[f * nksp + k1] < tc);
        if (S.fcnt[k0] > FC || S.fcnt[k1] > FC)
          par = (S.kspin[k0] ^ S.kspin[k1]) ^ k1_parity_slow(S.time, S.info, S.off, sso, sse, k0, tc) ^
                k1_parity_slow(S.time, S.info, S.off, sso, sse, k1, tc);
        const float4 pr = d.bond_p[b];
        const float u = u24(x.z);
        g = -1;
        if (par) { if (u < pr.x) g = 0; else if (u < pr.y) g = 2; }
        else     { if (u < pr.z) g = 1; else if (u < pr.w) g = 3; }
      }
      S.ctime[c] = tc;
      S.cacc[c] = (g >= 0) ? (uint8_t)(1 | (g << 1)) : (uint8_t)0;
      accepted = g >= 0;
    }
    const unsigned m = __ballot_sync(0xffffffffu, accepted);   // compact the accepted candidates
    int base = 0;
    if (lane == 0 && m) base = atomicAdd(&s_cnt[1], __popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (accepted) S.alist[base + __popc(m & ((1u << lane) - 1u))] = (uint16_t)c;
  }
  __syncthreads();

  // ---- 5: new bucket sizes -> offsets ------------------------------------------------------------
  int cnt = 0;
  if (tid < nb) {
    int nacc = 0;
    for (int i = 0; i < K; ++i) nacc += S.cacc[cb + i] & 1;
    cnt = S.nkb[tid] + nacc;
  }
  int total;
  const int noff = block_exscan(cnt, &total, s_scan);
  if (total > d.cap) {
    if (tid == 0) { atomicOr(d.d_err, LQ_ERR_PAGE_FULL); d.pcount[dst][p] = 0; }
    if (tid < nb) bo_new[tid] = 0;
    if (tid == 0) bo_new[nb] = 0;
    return;
  }
  if (tid < nb) { bo_new[tid] = (uint16_t)noff; S.noff[tid] = noff; }
  if (tid == 0) { bo_new[nb] = (uint16_t)total; d.pcount[dst][p] = total; }
  __syncthreads();

  // ---- 6: scatter into the compacted new page ----------------------------------------------------
  double* wt = d.time[dst] + p * (size_t)d.cap;
  uint32_t* wi = d.info[dst] + p * (size_t)d.cap;
  const int n_acc = s_cnt[1], n_keep = s_cnt[0];
  for (int ia = tid; ia < n_acc; ia += blockDim.x) {
    const int c = S.alist[ia];
    const uint32_t acc = S.cacc[c];
    const int lb = S.clb[c];
    const double tc = S.ctime[c];
    int rank = 0;
    const int o1 = S.off[lb + 1];
    for (int j = S.off[lb]; j < o1; ++j)      // kept operators come first on ties
      rank += (int)(S.info[j] & LQ_INFO_OFFDIAG) & (int)(S.time[j] <= tc);
    const int c1 = S.cbase[lb + 1];
    for (int k = S.cbase[lb]; k < c; ++k)        // equal times: the earlier draw comes first
      rank += (int)(S.cacc[k] & 1) & (int)(S.ctime[k] <= tc);
    for (int k = c + 1; k < c1; ++k)
      rank += (int)(S.cacc[k] & 1) & (int)(S.ctime[k] < tc);
    const int pos = S.noff[lb] + rank;
    wt[pos] = tc;
    wi[pos] = ((uint32_t)lb << LQ_INFO_LBSHIFT) | ((acc >> 1) << LQ_INFO_GSHIFT) |
              ((d.has_site && bsx[2 * lb + 1] < 0) ? LQ_INFO_SITE : 0u);
  }
  for (int ik = tid; ik < n_keep; ik += blockDim.x) {
    const int j = S.klist[ik];
    const uint32_t inf = S.info[j];
    const int lb = (int)(inf >> LQ_INFO_LBSHIFT);
    const double tt = S.time[j];
    const int o0 = S.off[lb];
    int rank = 0;
    for (int k = o0; k < j; ++k) rank += (int)(S.info[k] & LQ_INFO_OFFDIAG);
    const int c1 = S.cbase[lb + 1];
    for (int k = S.cbase[lb]; k < c1; ++k) rank += (int)(S.cacc[k] & 1) & (int)(S.ctime[k] < tt);
    uint32_t g = 0;
    const int b = b0 + lb;
    const float q0 = d.bond_q[b];
    if (q0 < 1.0f) {  // graph_impl.h:324-327 choose_offdiagonal
      const philox_t x = philox4x32_10((uint32_t)b, (uint32_t)wg, mcs, LQ_STREAM_OFFD + (uint32_t)(j - o0), key0, key1);
      g = (u24(x.x) < q0) ? 0u : 1u;
    }
    const int pos = S.noff[lb] + rank;
    wt[pos] = tt;
    wi[pos] = ((uint32_t)lb << LQ_INFO_LBSHIFT) | (g << LQ_INFO_GSHIFT) | LQ_INFO_OFFDIAG | (inf & LQ_INFO_SITE);
  }
}

}  // namespace lq
