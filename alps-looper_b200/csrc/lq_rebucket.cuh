// lq_rebucket.cuh -- the configuration moves into a new page layout ON THE DEVICE.
//
// lq_set_beta (path_integral.C:98-105; the annealing of temperature.h:39-80 calls it every step) and the
// rewind after an arena overflow change the number of windows and/or the page capacity.  Round 1 sent
// every operator through the host (lq_get_state -> sort -> lq_set_state: minutes and tens of GB of
// host memory at 1e9 operators).  The operators of one (tile, bond) column, read window
// after window, are already time-ordered, and imaginary times are stored as fractions of beta: a new
// layout only re-cuts every column at the new window boundaries.  One thread per column counts, one
// CTA per new page turns the sizes into bucket offsets, the same threads scatter; the spins at the new
// window starts are the spins at the first one xor the parity of the off-diagonal legs below.
#pragma once
#include "lq_device.cuh"

namespace lq {

struct OldPages {   // the layout the operators come from (the live buffer of the old arenas)
  const double* time;
  const uint32_t* info;
  const uint16_t* boff;
  int W, Wl, w0, cap;
};

// window of a time: the arithmetic of window_of() in lq_engine.cu (host bucketing of lq_set_state)
__device__ __forceinline__ int window_of_dev(double t, int W) {
  int w = (int)(t * (double)W);
  if (w >= W) w = W - 1;
  if (w < 0) w = 0;
  while (w > 0 && window_lo(w, W) > t) --w;
  while (w + 1 < W && window_hi(w, W) <= t) ++w;
  return w;
}

// PASS 0: sizes of the new buckets, stored one slot up in boff[0] (zeroed before); PASS 1: placement,
// boff[0] holds the offsets by then.  One thread per (tile, bond) column of the local tiles.
template <int PASS>
__global__ void __launch_bounds__(128)
k_rb_columns(Dev d, OldPages o, int nbonds_local) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbonds_local) return;
  const int tl = d.bond_tl[b];
  const int t = tl >> 10, lb = tl & 1023;
  const size_t nb1 = (size_t)d.nbmax + 1;
  int cw = -1, cnt = 0, lim = 0;   // new window being filled, operators placed in it so far, size of its bucket
  size_t dst = 0;
  for (int wo = 0; wo < o.Wl; ++wo) {
    const size_t po = (size_t)t * o.Wl + wo;
    const uint16_t* bo = o.boff + po * nb1 + lb;
    const int j0 = bo[0], j1 = bo[1];
    for (int j = j0; j < j1; ++j) {
      const double tt = o.time[po * (size_t)o.cap + j];
      const int wn = window_of_dev(tt, d.W) - d.w0;
      if (wn < 0 || wn >= d.Wl) { atomicOr(d.d_err, LQ_ERR_BOUNDARY); continue; }   // cannot happen: slab bounds are window bounds
      if (wn != cw) {
        const size_t pn = (size_t)t * d.Wl + wn;
        if (PASS == 0) {
          if (cw >= 0) d.boff[0][((size_t)t * d.Wl + cw) * nb1 + lb + 1] = (uint16_t)min(cnt, 65535);
        } else {
          const int o0 = d.boff[0][pn * nb1 + lb];
          lim = (int)d.boff[0][pn * nb1 + lb + 1] - o0;   // (0 on a page that overflowed: k_rb_offsets)
          dst = pn * (size_t)d.cap + o0;
        }
        cw = wn;
        cnt = 0;
      }
      if (PASS == 1 && cnt < lim) {
        d.time[0][dst + cnt] = tt;
        d.info[0][dst + cnt] = o.info[po * (size_t)o.cap + j];
      }
      ++cnt;
    }
  }
  if (PASS == 0 && cw >= 0) d.boff[0][((size_t)t * d.Wl + cw) * nb1 + lb + 1] = (uint16_t)min(cnt, 65535);
}

// one CTA per new page: bucket sizes (one slot up) -> offsets, page count, capacity check
__global__ void __launch_bounds__(128)
k_rb_offsets(Dev d) {
  __shared__ int s_scan[34];
  const size_t p = blockIdx.x;
  const int t = (int)(p / (size_t)d.Wl);
  const int nb = d.bond_base[t + 1] - d.bond_base[t];
  uint16_t* bo = d.boff[0] + p * (size_t)(d.nbmax + 1);
  const int per = (nb + (int)blockDim.x - 1) / (int)blockDim.x;
  const int l0 = min((int)threadIdx.x * per, nb), l1 = min(l0 + per, nb);
  int sum = 0;
  for (int lb = l0; lb < l1; ++lb) sum += bo[lb + 1];
  int total;
  int run = block_exscan(sum, &total, s_scan);
  if (total > d.cap) {   // the page overflows its capacity: the step that follows would be lost anyway -- flag, keep it empty
    if (threadIdx.x == 0) { atomicOr(d.d_err, LQ_ERR_PAGE_FULL); d.pcount[0][p] = 0; }
    for (int lb = l0; lb < l1; ++lb) bo[lb + 1] = 0;
    return;
  }
  for (int lb = l0; lb < l1; ++lb) { run += bo[lb + 1]; bo[lb + 1] = (uint16_t)run; }   // inclusive: offset of bucket lb + 1
  if (threadIdx.x == 0) { bo[0] = 0; d.pcount[0][p] = total; }
}

// parity of the off-diagonal legs of every (site, new window), left in the spin row of the NEXT window start
__global__ void k_rb_spin_parity(Dev d, int ntiles_local) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)d.Wl * d.N) return;
  const int wl = (int)(i / (size_t)d.N), s = (int)(i - (size_t)wl * d.N);
  int par = 0;
  for (int a = d.adj_off[s]; a < d.adj_off[s + 1]; ++a) {
    const int tl = d.bond_tl[d.adj[a] >> 1];
    if ((tl >> 10) >= ntiles_local) continue;   // (spatial cut: a tile this rank does not hold; the site's spin is never read)
    const size_t p = (size_t)(tl >> 10) * d.Wl + wl;
    const uint16_t* bo = d.boff[0] + p * (size_t)(d.nbmax + 1) + (tl & 1023);
    for (int j = bo[0]; j < bo[1]; ++j) par ^= (int)(d.info[0][p * (size_t)d.cap + j] & LQ_INFO_OFFDIAG);
  }
  d.spinW[(size_t)(wl + 1) * d.N + s] = (uint8_t)par;
}

__global__ void k_rb_spin_scan(Dev d, const uint8_t* __restrict__ spins0) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.N) return;
  uint8_t c = spins0[s];
  d.spinW[s] = c;
  for (int wl = 0; wl < d.Wl; ++wl) {
    c ^= d.spinW[(size_t)(wl + 1) * d.N + s];
    d.spinW[(size_t)(wl + 1) * d.N + s] = c;
  }
}

}  // namespace lq
