// lq_space.cuh -- multi-GPU, spatial cut (lq_options.cut = LQ_CUT_SPACE).
//
// Reference: the reference shares the LATTICE among OpenMP threads (bond ownership,
// looper/lattice.h:692-787, wavefront wait path_integral.C:494-502) and imaginary time among MPI
// ranks (looper/parallel.h); BASELINE config 3 asks for the lattice cut across GPUs.  Here rank q owns
// a contiguous range of tiles over the whole imaginary-time axis and keeps ghost copies of the
// neighbouring tiles (lq_engine.cu, SpacePlan).  K1, the walk, the unions and the estimators run
// unchanged on the owned tiles (the walk also on the W ghost tiles, so that every leg of an owned
// operator finds its lower node locally); every edge of the world-line graph is applied by exactly
// one rank.  What crosses a cut are the NODES both sides refer to: for every ordered pair
// (owner, user) a boundary segment lists the owner's sites that are far-end sites of the user and
// all operators on the owner's bonds that touch a K-site of the user -- dense, in (bond, window,
// slot) order, which both sides derive from identical copies of the pages.  From there on the merge
// is the one of the imaginary-time slabs (parallel.h:1609-1809): every entry publishes the smallest
// entry of its local cluster (pack_tree, union_find.h:352-400), one all-gather, every rank unifies
// entry i of the owner's segment with entry i of the user's redundantly, partial sums of the open
// clusters meet in one integer all-reduce indexed by the global cluster id, flips are Philox draws
// keyed by that id.
#pragma once
#include "lq_kernels.cuh"

namespace lq {

#define LQ_SP_HDR 4   /* header words of a rank's boundary buffer: [0] error bits of the rank */

struct SpSeg {          // a boundary segment this rank takes part in
  long long off;        // first entry in this rank's buffer
  long long opcap;      // operators the segment can hold
  int ns, nb;           // sites, bonds
  int site0, bond0;     // first entry in the site / bond tables
  int user_side;        // this rank holds the GHOST copies of the segment's nodes
  int pad;
  long long pk0;        // first operator slot of the segment in the packed halo stream (pk_time / pk_info)
  long long sp0;        // first byte of the segment in the packed spin stream
};
struct SpGSeg { int owner, user; long long off_owner, off_user, cap; };

struct SpDev {
  int nseg, nbt, nst, ngseg;
  long long cb;         // entries of this rank's buffer
  long long cbmax;      // max over ranks
  long long stride;     // LQ_SP_HDR + cbmax: words per rank in the gathered buffer
  const SpSeg* seg;
  const int* site;      // [nst] local site of a site entry
  const int* sseg;      // [nst] its segment
  const int* bond;      // [nbt] local bond of a bond entry
  const int* bseg;      // [nbt] its segment
  uint32_t* cnt;        // [nbt*Wl + 1] operators in (bond entry, window)
  uint32_t* base;       // [nbt*Wl + 2] exclusive scan
  node_t* bnode;        // [cbmax] local node of every entry (NODE_NONE: unused)
  const SpGSeg* gseg;   // [ngseg] all segments of the run
  // halo stream: the operators of the segments' bonds, dense in (bond, window, slot) order -- what the
  // owner sends instead of whole pages; the user rebuilds its ghost pages from it
  double* pk_time;      // [sum of opcap over this rank's segments]
  uint32_t* pk_info;
  uint8_t* pk_spin;     // [(Wl+1) * ns per segment] spins of the segments' sites at every window start
  const int* gt_off;    // [ghost tiles + 1] needed buckets of a ghost tile ...
  const int* gt_lb;     //   local bucket in the tile (ascending)
  const int* gt_j;      //   its bond entry (index into bond / bseg)
};

// Ghost nodes nobody here refers to must not become clusters: they start as NODE_JUNK (k_compress and
// k_relabel skip them) and k_sp_fill makes the boundary candidates among them proper singletons.
__global__ void k_sp_init_ghost(Dev d) {
  const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t lo = (size_t)d.N + (size_t)d.npo * (size_t)d.nbase[d.Pown];
  const size_t hi = (size_t)d.N + (size_t)d.npo * (size_t)(*d.d_ntotal);
  if (lo + x < hi) d.parent[lo + x] = NODE_JUNK;
}

__global__ void k_sp_count(Dev d, SpDev sp, int buf) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)sp.nbt * d.Wl) return;
  const int j = (int)(i / d.Wl), wl = (int)(i - (size_t)j * d.Wl);
  sp.cnt[i] = (uint32_t)bucket_of(d, buf, sp.bond[j], wl).n;
}

// entries of the operators: segment offset + ns + npo * (dense rank in (bond, window, slot) order) + side
__global__ void k_sp_fill(Dev d, SpDev sp, int buf) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)sp.nst) {   // site entries
    const SpSeg g = sp.seg[sp.sseg[i]];
    const int s = sp.site[i];
    sp.bnode[g.off + ((long long)i - g.site0)] = (node_t)s;
    if (g.user_side) d.parent[s] = (node_t)s;
  }
  if (i >= (size_t)sp.nbt * d.Wl) return;
  const int j = (int)(i / d.Wl), wl = (int)(i - (size_t)j * d.Wl);
  const SpSeg g = sp.seg[sp.bseg[j]];
  const long long rel = (long long)sp.base[i] - (long long)sp.base[(size_t)g.bond0 * d.Wl];
  const BucketRef r = bucket_of(d, buf, sp.bond[j], wl);
  // (a segment that overflows loses the step -- all ranks rewind and grow -- but its ghost nodes must
  // still become proper nodes: the unions of this step reach them)
  const bool fits = rel + r.n <= g.opcap;
  if (!fits) atomicOr(d.d_err, LQ_ERR_NODE_FULL);
  for (int k = 0; k < r.n; ++k)
    for (int side = 0; side < d.npo; ++side) {
      const node_t x = upper_node(d, r.idx0 + k, side);
      if (fits) sp.bnode[g.off + g.ns + (long long)d.npo * (rel + k) + side] = x;
      if (g.user_side) d.parent[x] = x;
    }
}

// owner side: the buckets of the segments' bonds -> dense stream (same order as the boundary entries)
__global__ void k_sp_pack(Dev d, SpDev sp, int buf, int with_spins) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (with_spins && i < (size_t)sp.nst * (size_t)(d.Wl + 1)) {   // spins of the segments' sites, row w = window start w
    const size_t e = i % (size_t)sp.nst, w = i / (size_t)sp.nst;
    const SpSeg g = sp.seg[sp.sseg[e]];
    if (!g.user_side) sp.pk_spin[g.sp0 + (long long)w * g.ns + ((long long)e - g.site0)] = d.spinW[w * (size_t)d.N + sp.site[e]];
  }
  if (i >= (size_t)sp.nbt * d.Wl) return;
  const int j = (int)(i / d.Wl), wl = (int)(i - (size_t)j * d.Wl);
  const SpSeg g = sp.seg[sp.bseg[j]];
  if (g.user_side) return;
  const long long rel = (long long)sp.base[i] - (long long)sp.base[(size_t)g.bond0 * d.Wl];
  const BucketRef r = bucket_of(d, buf, sp.bond[j], wl);
  if (rel + r.n > g.opcap) { atomicOr(d.d_err, LQ_ERR_NODE_FULL); return; }   // (the user sees the same counts and skips the bucket too)
  for (int k = 0; k < r.n; ++k) {
    sp.pk_time[g.pk0 + rel + k] = d.time[buf][r.base + k];
    sp.pk_info[g.pk0 + rel + k] = d.info[buf][r.base + k];
  }
}

// user side: one CTA per ghost page rebuilds it from the stream -- the needed buckets of the tile in
// ascending bucket order, all other buckets empty
__global__ void __launch_bounds__(256)
k_sp_unpack(Dev d, SpDev sp, int buf, int with_spins) {
  __shared__ int s_scan[34];
  __shared__ int s_carry;
  if (with_spins) {   // (grid-stride over the received spins; every CTA takes its share)
    const size_t total = (size_t)sp.nst * (size_t)(d.Wl + 1);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
      const size_t e = i % (size_t)sp.nst, w = i / (size_t)sp.nst;
      const SpSeg g = sp.seg[sp.sseg[e]];
      if (g.user_side) d.spinW[w * (size_t)d.N + sp.site[e]] = sp.pk_spin[g.sp0 + (long long)w * g.ns + ((long long)e - g.site0)];
    }
  }
  const int gt = (int)(blockIdx.x / (unsigned)d.Wl), wl = (int)(blockIdx.x - (unsigned)gt * (unsigned)d.Wl);
  const int t = d.Pown / d.Wl + gt;   // ghost tiles follow the owned ones
  const size_t p = (size_t)t * d.Wl + wl;
  const int nb = d.bond_base[t + 1] - d.bond_base[t];
  uint16_t* bo = d.boff[buf] + p * (size_t)(d.nbmax + 1);
  double* pt = d.time[buf] + p * (size_t)d.cap;
  uint32_t* pi = d.info[buf] + p * (size_t)d.cap;
  const int e0 = sp.gt_off[gt], ne = sp.gt_off[gt + 1] - e0;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  const int nchunks = max(1, (ne + (int)blockDim.x - 1) / (int)blockDim.x);
  for (int c = 0; c < nchunks; ++c) {
    const int e = c * (int)blockDim.x + (int)threadIdx.x;
    int n = 0, lb = 0;
    long long src = 0;
    if (e < ne) {
      lb = sp.gt_lb[e0 + e];
      const int j = sp.gt_j[e0 + e];
      const SpSeg g = sp.seg[sp.bseg[j]];
      const size_t ci = (size_t)j * d.Wl + wl;
      const long long rel = (long long)sp.base[ci] - (long long)sp.base[(size_t)g.bond0 * d.Wl];
      n = (int)sp.cnt[ci];
      if (rel + n > g.opcap) { atomicOr(d.d_err, LQ_ERR_NODE_FULL); n = 0; }   // (the owner did not pack it either)
      src = g.pk0 + rel;
    }
    int total;
    const int ex = block_exscan(n, &total, s_scan);
    const int carry = s_carry;
    int off = carry + ex;
    if (carry + total > d.cap) {   // cannot happen with identical page capacities; never write beyond the page
      if (threadIdx.x == 0) atomicOr(d.d_err, LQ_ERR_PAGE_FULL);
      n = 0;
      off = min(off, d.cap);
    }
    if (e < ne) {
      // the buckets after the previous needed one up to this one all start where this one starts
      const int lb_prev = (e == 0) ? -1 : sp.gt_lb[e0 + e - 1];
      for (int b = lb_prev + 1; b <= lb; ++b) bo[b] = (uint16_t)off;
      for (int k = 0; k < n; ++k) { pt[off + k] = sp.pk_time[src + k]; pi[off + k] = sp.pk_info[src + k]; }
    }
    __syncthreads();
    if (threadIdx.x == 0) s_carry = min(carry + total, d.cap);
    __syncthreads();
  }
  const int tot = s_carry;
  const int lb_last = ne > 0 ? sp.gt_lb[e0 + ne - 1] : -1;
  for (int b = lb_last + 1 + (int)threadIdx.x; b <= nb; b += blockDim.x) bo[b] = (uint16_t)tot;
  if (threadIdx.x == 0) d.pcount[buf][p] = tot;
}

__global__ void k_sp_topmin(Dev d, SpDev sp, MrDev m) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (size_t)sp.cb) return;
  const node_t x = sp.bnode[k];
  if (x == NODE_NONE) return;
  const uint32_t c = cid_pre(d, x);
  if ((long long)c < d.nccap) atomicMin(m.topmin + c, (uint32_t)k);
}

__global__ void k_sp_ids(Dev d, SpDev sp, MrDev m) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < LQ_SP_HDR) m.sendb[k] = (k == 0) ? (uint32_t)(*d.d_err) : 0u;
  if (k >= (size_t)sp.cbmax) return;
  uint32_t v = NODE_NONE;
  if (k < (size_t)sp.cb) {
    const node_t x = sp.bnode[k];
    if (x != NODE_NONE) {
      const uint32_t c = cid_pre(d, x);
      v = ((long long)c < d.nccap) ? m.topmin[c] : (uint32_t)k;   // (arena overflow: the step is lost anyway)
    }
  }
  m.sendb[LQ_SP_HDR + k] = v;
}

// gathered buffer -> forest over (rank, entry): every entry under the representative of its local cluster
__global__ void k_sp_ginit(Dev d, SpDev sp, MrDev m) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m.gn) return;
  const size_t r = i / (size_t)sp.stride, k = i - r * (size_t)sp.stride;
  uint32_t par = (uint32_t)i;
  uint8_t used = 0;
  if (k < LQ_SP_HDR) {
    if (k == 0 && m.recvb[i] != 0u && (int)r != d.rank) atomicOr(d.d_err, LQ_ERR_REMOTE);   // all ranks rewind together
  } else {
    const uint32_t v = m.recvb[i];
    if (v != NODE_NONE) { par = (uint32_t)(r * (size_t)sp.stride + LQ_SP_HDR + v); used = 1; }
  }
  m.gparent[i] = par;
  m.gused[i] = used;
}

// entry i of a segment is the same physical node on the owner and on the user
__global__ void k_sp_gunion(Dev d, SpDev sp, MrDev m) {
  const SpGSeg g = sp.gseg[blockIdx.y];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.cap) return;
  const size_t a = (size_t)g.owner * (size_t)sp.stride + LQ_SP_HDR + (size_t)(g.off_owner + i);
  const size_t b = (size_t)g.user * (size_t)sp.stride + LQ_SP_HDR + (size_t)(g.off_user + i);
  const bool ua = m.recvb[a] != NODE_NONE, ub = m.recvb[b] != NODE_NONE;
  if (ua && ub) uf_union(m.gparent, (node_t)a, (node_t)b);
  else if (ua != ub) atomicOr(d.d_err, LQ_ERR_BOUNDARY);   // the two copies of a page differ
}

__device__ __forceinline__ bool sp_is_rep(const Dev& d, const SpDev& sp, const MrDev& m, size_t k, bool relabelled, uint32_t* cid) {
  if (k >= (size_t)sp.cb) return false;
  const node_t x = sp.bnode[k];
  if (x == NODE_NONE) return false;
  const uint32_t c = relabelled ? LQ_CID(d.parent[x]) : cid_pre(d, x);
  *cid = c;
  return (long long)c < d.nccap && m.topmin[c] == (uint32_t)k;
}

__global__ void k_sp_openflips(Dev d, SpDev sp, MrDev m, const StepParams* __restrict__ spar) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t c;
  if (!sp_is_rep(d, sp, m, k, false, &c)) return;
  const uint32_t gid = global_cid(d, m, (uint32_t)k);
  const philox_t x = philox4x32_10(gid, 0xffffffffu, spar->mcs, LQ_STREAM_FLIP, spar->key0, spar->key1);
  if (x.x & 1u) atomicOr(d.flipw + (c >> 5), 1u << (c & 31u));
  else atomicAnd(d.flipw + (c >> 5), ~(1u << (c & 31u)));
}

// the representative entry of every open local cluster moves its partial sums into the global table
// (collect_estimates, parallel.h:415-427); warp- and CTA-level merging by global id as in k_mr_gather
__global__ void __launch_bounds__(1024)
k_sp_gather(Dev d, SpDev sp, MrDev m) {
  __shared__ uint32_t s_key[LQ_GATHER_SLOTS];
  __shared__ unsigned long long s_val[LQ_GATHER_SLOTS][LQ_GEST_MAX];
  for (int i = threadIdx.x; i < LQ_GATHER_SLOTS * LQ_GEST_MAX; i += blockDim.x) {
    if (i < LQ_GATHER_SLOTS) s_key[i] = 0xffffffffu;
    (&s_val[0][0])[i] = 0ull;
  }
  __syncthreads();
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = threadIdx.x & 31u;
  uint32_t c = 0;
  const bool rep = sp_is_rep(d, sp, m, k, true, &c);
  uint32_t gid = 0xffffffffu;
  unsigned long long v[LQ_GEST_MAX];
#pragma unroll
  for (int f = 0; f < LQ_GEST_MAX; ++f) v[f] = 0ull;
  if (rep) {
    gid = global_cid(d, m, (uint32_t)k);
#pragma unroll
    for (int f = 0; f < 4; ++f) v[f] = atomicExch((unsigned long long*)d.est + f * d.nccap + c, 0ull);
    if (c < d.d_nc[1]) {
      int z[4];
#pragma unroll
      for (int f = 0; f < 4; ++f) z[f] = atomicExch(d.est0 + f * (size_t)d.N + c, 0);
      v[4] = pack2_i32(z[0], z[1]);   // the four tau = 0 sums (|x| <= N) travel as two fields
      v[5] = pack2_i32(z[2], z[3]);
    }
    if (d.has_site && ((d.openw[c >> 5] >> (c & 31u)) & 1u)) v[6] = 1ull;
    for (int x = 0; x < d.sdim; ++x) v[6 + d.has_site + x] = (unsigned long long)(long long)atomicExch(d.wind + (size_t)x * d.nccap + c, 0);
    if (gid >= m.gcap) {   // (see k_mr_gather)
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
    const int h = (int)((gid * 2654435761u) >> 26);
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
  const unsigned nrep = __reduce_add_sync(0xffffffffu, rep ? 1u : 0u);
  if (lane == 0 && nrep) atomicAdd(m.d_g + 1, nrep);
  __syncthreads();
  for (int i = threadIdx.x; i < LQ_GATHER_SLOTS * d.gstride; i += blockDim.x) {
    const int sl = i / d.gstride, f = i - sl * d.gstride;
    const uint32_t g2 = s_key[sl];
    const unsigned long long t = s_val[sl][f];
    if (g2 != 0xffffffffu && t) atomicAdd((unsigned long long*)m.gest + (size_t)g2 * d.gstride + f, t);
  }
}

__global__ void k_sp_reset(Dev d, SpDev sp, MrDev m) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= (size_t)sp.cb) return;
  const node_t x = sp.bnode[k];
  if (x == NODE_NONE) return;
  const uint32_t c = LQ_CID(d.parent[x]);   // (runs after k_relabel)
  if ((long long)c < d.nccap) m.topmin[c] = 0xffffffffu;
}

}  // namespace lq
