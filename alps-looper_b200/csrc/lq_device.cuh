// lq_device.cuh -- device-side data layout, Philox, lock-free union-find, block primitives.
//
// Layout in HBM (see DESIGN.md "Data layout"):
//   space is cut into T tiles of neighbouring sites, imaginary time into W windows; the operators
//   of the bonds owned by tile t inside window w form one PAGE p = t*Wl + wl of fixed capacity
//   `cap`, stored SoA (time f64, info u32), grouped by bond and time-sorted inside each
//   (bond, window) BUCKET; boff[p*(nbmax+1) + lb] are the bucket offsets inside the page.
//   Node ids of the world-line graph: site s -> s, upper leg(s) of operator #idx -> N + NPO*idx(+side)
//   with idx = nbase[p] + j dense over the occupied slots.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <type_traits>

namespace lq {

typedef uint32_t node_t;
static const node_t NODE_NONE = 0xffffffffu;
static const node_t NODE_JUNK = 0xfffffffeu;   // spatial cut: ghost node nobody on this rank refers to (no cluster of its own)

// info word of an operator (mirrors looper/operator.h type_ in the low bits)
//   bit 0      offdiagonal            (local_operator_type::offdiagonal, operator.h:44)
//   bits 2..3  graph type g           (type_ >> 2, operator.h:76; graph_impl.h:93-103)
//   bit 4      site operator (the "bond" is a pseudo-bond with one end, see lq_engine.cu)
//   (the spins below the operator travel in bit 31 of low0 / low1, written by the walk)
//   bits 8..   local bond index inside the owning tile
#define LQ_INFO_OFFDIAG 1u
#define LQ_INFO_GSHIFT 2
#define LQ_INFO_SITE 16u   /* site operator (location_impl.h:37 is_site): cuts the world line */
#define LQ_INFO_LBSHIFT 8

// error bits raised by kernels (sticky, read back by the host after a sweep)
#define LQ_ERR_PAGE_FULL 1
#define LQ_ERR_CAND_FULL 2
#define LQ_ERR_NODE_FULL 4
#define LQ_ERR_CLUSTER_FULL 8
#define LQ_ERR_REMOTE 16   /* slab engines: another rank overflowed in this step (all ranks rewind together) */
#define LQ_ERR_OPEN_FULL 64 /* multi-rank: more open clusters than this step's all-reduce carries (replayed with the exact count) */
#define LQ_ERR_BOUNDARY 32 /* spatial cut: the two copies of a boundary page disagree (internal error) */

// fixed-point scale of imaginary time in the cluster sums (order-independent integer atomics)
#define LQ_FX 1099511627776.0 /* 2^40 */

// per-step inputs, copied host -> device from pinned memory before every step
struct StepParams {
  double beta;
  uint32_t key0, key1;  // Philox key (seed)
  uint32_t mcs;         // step counter = Philox counter word
  uint32_t pad;
};

struct Dev {
  // ---- lattice, internal (tile-contiguous) numbering ----
  int N, B, T, nbmax;
  int W;            // global number of windows
  int w0, Wl;       // this rank owns windows [w0, w0+Wl)
  const double* wlo;  // [W+1] window bounds: wlo[w] = window_lo(w, W), wlo[W] = 1 (two f64 divisions less per thread)
  const double* wks;  // [W]   scale of the 32-bit window-relative time keys (k1_key): 4294967040 / (wlo[w+1] - wlo[w])
  int cap;          // page capacity (operators)
  int npo;          // nodes per operator (1: graphs {0,2,3}; 2: cross graph present)
  int ug;           // windows per union group (k_union_local / k_union_global)
  int has_site;     // the model has site graphs: bonds with bond_s1 < 0 are site pseudo-bonds
  int zero_umag;    // every operator sits on antiparallel spins (graphs 0 and 2 only, no site graphs): clusters
                    // without a site node have umag == 0 identically
  int zero_ssize;   // every bond joins sites of opposite gauge (and no site graphs): their ssize == 0 identically
  int rank, nranks;
  // spatial cut (lq_space.cuh): the sites / pages this rank OWNS come first, then the ghost copies.
  // Serial and slab engines: Nown = Nwalk = N, Pown = all pages, space = 0.
  int space;        // LQ_CUT_SPACE engine
  int Nown;         // owned sites [0, Nown)
  int Nwalk;        // sites whose world lines are walked here [0, Nwalk): owned + W ghost tiles
  int Pown;         // owned pages [0, Pown)
  const uint32_t* tile_key;   // spatial cut: rank-independent id of a tile; the Philox counters of K1 use
                              // tile_key << 10 | local bond (the internal bond numbering differs from rank to
                              // rank: two ranks must not draw the same stream); NULL: the internal bond id
  const int* bond_s0;    // [B] source site
  const int* bond_s1;    // [B] target site
  const int* bond_tl;    // [B] owning tile << 10 | local bond index
  const double* bond_emu;  // [B] exp(-beta * rate / W): P(no candidate in a window)
  const int* whalo_cnt;  // [T] leading halo buckets that touch an own site (walk halo)
  const int* bond_base;  // [T+1] first bond of tile
  const int* adj_off;    // [N+1]
  const int* adj;        // [2B] (bond << 1 | side)
  const double* bond_rate;  // [B] sum_g v_g  (candidate rate / beta, graph_impl.h:694)
  const float4* bond_p;     // [B] {P(g=0|anti), P(accept|anti), P(g=1|par), P(accept|par)}
  const float* bond_q;      // [B] P(g=0 | offdiagonal)  (graph_impl.h:311-313)
  const signed char* gauge;  // [N] +1/-1/0
  // ---- static per-tile halo lists and per-class stencils (see Partition in lq_engine.cu) ----
  const int* site_base;   // [T+1] first site of a tile (sites are tile-contiguous)
  const int* halo_off;    // [T+1]
  const int* halo_bond;   // global bond ids of the halo buckets of a tile
  const int* halo_tl;     // bond_tl of the same buckets (owning tile << 10 | local bond index)
  const int* hsite_off;   // [T+1]
  const int* hsite;       // global ids of the halo K-sites of a tile
  const int* tile_class;  // [T]
  const int* cls_bs;      // [nclasses] start of the class in bs
  const int* cls_sso;     // [nclasses] start of the class in sst_off
  const int* cls_sst;     // [nclasses] start of the class in sst
  const int* cls_nks;     // [nclasses] K-sites of the class
  const int* bs;          // per class [2*(nb+H)] K-site of the two ends of a local bucket (-1: none)
  const int* sst_off;     // per class [nks+1]
  const int* sst;         // (local bucket << 1 | side) incident to a K-site
  int hmax;               // max halo buckets per tile
  int nksmax, zmax;       // max K-sites per tile, max coordination number
  int scap;               // operators that fit the shared-memory stage (own page + halo)
  int ccap;               // candidates per page that fit the stage (K1)
  int k1_keyshift;        // LQ_K1_KEYBITS (tests): the 32-bit time keys of K1 are coarsened by this shift
  int k1_timebits;        // LQ_K1_TIMEBITS (tests): candidate times are cut to this many bits of a window, so that
                          // EQUAL f64 times on neighbouring bonds -- once per ~10^4 steps at full resolution -- happen
                          // all the time (0: off)
  int kcap;               // kept (off-diagonal) operators per page that fit the kept list of K1
  // ---- pages (double buffered) ----
  double* time[2];
  uint32_t* info[2];
  uint16_t* boff[2];
  int* pcount[2];
  int* nbase;  // [P+1] exclusive scan of pcount (dense operator index base)
  // ---- per (window, site) carries ----
  uint8_t* spinW;  // [(Wl+1)*N] spin at the start of local window wl
  node_t* curW;    // [(Wl+1)*N] node of the world-line segment crossing the start of window wl
  uint32_t* firstW;  // [Wl*N] first leg of the site in the window (operator | side << 31), NODE_NONE if none
  // ---- union-find / labels ----
  node_t* parent;  // [N + npo*ncap]; after k_relabel holds the cluster id
  node_t* low0;    // [ncap] node below the operator on the source side
  node_t* low1;    // [ncap] node below on the target side; low1 == low0 + lowstride (one allocation)
  uint32_t lowstride;
  uint32_t* bitmap;  // root flags, one bit per node
  uint32_t* wcount;  // roots per bitmap word -> exclusive scan in wbase
  uint32_t* wbase;
  uint4* rootw;      // [words] {id base, root flags, flip bits of the roots, 0} (serial engines, k_rootflip)
  int fpack;         // labels carry the flip decision in bit 31 when the estimators run
  int rootflip;      // ... because k_relabel decided it per root (serial engines); slab engines pack it afterwards (k_pack_flips)
  uint2* xedge;      // [groups][LQ_XCAP] edges that leave their union group (k_union_local -> k_union_global)
  int* xcount;       // [groups] entries of the list, -1 if it overflowed
  int xcap;          // entries a list may hold (LQ_XCAP; the tests lower it through the LQ_XCAP environment variable)
  // ---- clusters ----
  long long* est;  // [4][nccap] usize, umag, ssize, smag in half units of LQ_FX
  int* est0;       // [4][N]     usize0, umag0, ssize0, smag0 in half units
  uint32_t* flipw; // [nccap/32 + 1] flip decision per cluster id, packed
  uint32_t* openw; // [nccap/32 + 1] cluster is cut by a site operator (has_site only), packed
  int sdim;              // dimensions of the winding-number estimator (0: stiffness not measured)
  int gstride;           // int64 fields per global open cluster in the slab exchange
  const short* bond_vec; // [3*B] relative bond vectors (stiffness.h:63-76) in units of the smallest component
  double wscale[3];      // half that unit per dimension: (winding / 2) = wscale * integer sum
  int* wind;             // [sdim][nccap] winding of every cluster in those units
  // ---- SSE representation (sse.C): position of every operator in the time-ordered string ----
  int sse;               // LQ_REPR_SSE: estimator times are string positions, the top is the string length
  uint32_t* spos;        // [ncap] string position of the operator with dense index idx
  int nbin;              // time bins per window of the counting sort that produces them
  uint32_t* bincnt;      // [Wl*nbin + 1] operators per (window, bin); exclusive scan in binbase
  uint32_t* binbase;
  uint32_t* binfill;     // [Wl*nbin] scatter cursors
  double* sorted_time;   // [ncap] operators in (window, bin) order
  uint2* sorted_id;      // [ncap] {internal bond, dense index}
  long long ncap;   // operator arena (= P*cap)
  long long nccap;  // cluster arena
  // ---- scalars on device ----
  int* d_ntotal;     // total operators after the update (nbase[P])
  uint32_t* d_nc;    // [0] number of clusters, [1] clusters rooted at a site node
  int* d_err;
  int dbg;           // LQ_DBG environment variable (experiments only)
  unsigned long long* dbgc;  // [8] experiment counters (LQ_DBG & 1): edges, find hops, CAS retries, ...
};

// ------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11) -- counter-based, so every (bond, window, draw) and every
// cluster gets its own stream without any state in memory.
// ------------------------------------------------------------------------------------------
struct philox_t { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ philox_t philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                           uint32_t c3, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
#else
    uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  philox_t r; r.x = c0; r.y = c1; r.z = c2; r.w = c3;
  return r;
}

// uniform in (0,1] with 53 bits
__host__ __device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
  uint64_t v = (((uint64_t)a << 32) | b) >> 11;
  return (double)(v + 1) * (1.0 / 9007199254740992.0);
}
// uniform in [0,1) with 24 bits (graph choice)
__host__ __device__ __forceinline__ float u24(uint32_t a) { return (float)(a >> 8) * (1.0f / 16777216.0f); }

// RNG stream ids (counter word 3)
#define LQ_STREAM_CAND 0x10000000u
#define LQ_STREAM_OFFD 0x20000000u
#define LQ_STREAM_FLIP 0x30000000u

// ------------------------------------------------------------------------------------------
// Lock-free union-find (replaces looper/union_find.h:242-284 + atomic.h CAS lock).
// Invariant: parent[x] <= x; a root has parent[x] == x; hooking always puts the LARGER root under
// the SMALLER one with one atomicCAS, so the final root of a cluster is its minimum node index --
// the reference's LOOPER_USE_DETERMINISTIC_UNIFY rule (union_find.h:229-233, 260-264) -- and the
// partition is independent of the order in which threads win.
// Reads may hit stale L1 lines: a stale parent is still a valid ancestor (links only ever move
// towards smaller indices) and every hook is validated by the atomicCAS at L2, so staleness costs
// at most extra hops, while hot roots of big clusters are served from L1.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ node_t uf_load(const node_t* p) { return *(const volatile node_t*)p; }

// Two operators of ONE site with exactly the same f64 time (the continuum has no such pair; in f64 a
// candidate meets an off-diagonal leg of a neighbouring bond at its own time about once per 10^4 steps
// at the headline size): ordered by bond, lower key first.  The diagonal update (k1_parity_exact) and
// the walk (its tie path) must agree on that order -- a candidate accepted against the spins "before"
// the leg and then walked "after" it would sit on parallel spins -- and on the spatial cut two ranks
// that hold the same site must agree as well, so the key is rank independent: the tile's id in the
// global tiling | local bond (= the internal bond id on serial and slab engines, same order).
__device__ __forceinline__ uint32_t bond_order_key(const uint32_t* __restrict__ tile_key, const int* __restrict__ bond_tl, int b) {
  if (!tile_key) return (uint32_t)b;
  const int tl = bond_tl[b];
  return (tile_key[tl >> 10] << 10) | (uint32_t)(tl & 1023);
}

__device__ __forceinline__ node_t uf_find(node_t* parent, node_t x) {
  node_t p = uf_load(parent + x);
  while (p != x) {
    node_t gp = uf_load(parent + p);
    if (gp != p) parent[x] = gp;  // path halving: only non-roots are written, always to an ancestor
    x = p;
    p = gp;
  }
  return x;
}

// instrumented copy for LQ_DBG=1 (hop statistics of the global union-find)
__device__ __forceinline__ node_t uf_find_count(node_t* parent, node_t x, unsigned& hops) {
  node_t p = uf_load(parent + x);
  while (p != x) {
    node_t gp = uf_load(parent + p);
    if (gp != p) parent[x] = gp;
    x = p;
    p = gp;
    ++hops;
  }
  return x;
}
__device__ __forceinline__ void uf_union_count(node_t* parent, node_t a, node_t b, unsigned long long* c) {
  unsigned hops = 0, retries = 0;
  node_t ra = uf_find_count(parent, a, hops);
  node_t rb = uf_find_count(parent, b, hops);
  while (ra != rb) {
    if (ra < rb) { node_t t = ra; ra = rb; rb = t; }
    node_t old = atomicCAS(parent + ra, ra, rb);
    if (old == ra) break;
    ++retries;
    ra = uf_find_count(parent, old, hops);
    rb = uf_find_count(parent, rb, hops);
  }
  atomicAdd(c + 0, 1ull); atomicAdd(c + 1, (unsigned long long)hops); atomicAdd(c + 2, (unsigned long long)retries);
  if (hops > 16) atomicAdd(c + 3, 1ull);
}

__device__ __forceinline__ void uf_union(node_t* parent, node_t a, node_t b) {
  node_t ra = uf_find(parent, a);
  node_t rb = uf_find(parent, b);
  while (ra != rb) {
    if (ra < rb) { node_t t = ra; ra = rb; rb = t; }  // ra > rb: hook ra under rb
    node_t old = atomicCAS(parent + ra, ra, rb);
    if (old == ra) return;
    ra = uf_find(parent, old);  // somebody else hooked ra first; continue from its new ancestor
    rb = uf_find(parent, rb);
  }
}

// ------------------------------------------------------------------------------------------
// block-wide exclusive scan of one int per thread (warp shuffles + one smem hop)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int block_exscan(int v, int* total, int* smem /* >= 33 ints */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) smem[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = (lane < nw) ? smem[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += n;
    }
    smem[lane] = winc - w;
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  int res = smem[wid] + inc - v;
  *total = smem[32];
  __syncthreads();
  return res;
}

// the same for two ints per thread at once
__device__ __forceinline__ int2 block_exscan2(int2 v, int* total_x, int* total_y, int* smem /* >= 66 ints */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int ix = v.x, iy = v.y;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int nx = __shfl_up_sync(0xffffffffu, ix, o), ny = __shfl_up_sync(0xffffffffu, iy, o);
    if (lane >= o) { ix += nx; iy += ny; }
  }
  if (lane == 31) { smem[wid] = ix; smem[33 + wid] = iy; }
  __syncthreads();
  if (wid == 0) {
    const int wx = (lane < nw) ? smem[lane] : 0, wy = (lane < nw) ? smem[33 + lane] : 0;
    int sx = wx, sy = wy;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int nx = __shfl_up_sync(0xffffffffu, sx, o), ny = __shfl_up_sync(0xffffffffu, sy, o);
      if (lane >= o) { sx += nx; sy += ny; }
    }
    smem[lane] = sx - wx;
    smem[33 + lane] = sy - wy;
    if (lane == 31) { smem[32] = sx; smem[65] = sy; }
  }
  __syncthreads();
  const int2 res = make_int2(smem[wid] + ix - v.x, smem[33 + wid] + iy - v.y);
  *total_x = smem[32];
  *total_y = smem[65];
  __syncthreads();
  return res;
}

// One-barrier variant for kernels that scan several times per iteration (K1): every warp scans the
// warp totals itself, and two shared-memory buffers alternate (`par`), so the buffer a call writes was
// last read two calls -- hence at least one barrier -- ago.
__device__ __forceinline__ int2 block_exscan2_1b(int2 v, int* total_x, int* total_y, int (*buf)[2][32], int& par) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int ix = v.x, iy = v.y;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int nx = __shfl_up_sync(0xffffffffu, ix, o), ny = __shfl_up_sync(0xffffffffu, iy, o);
    if (lane >= o) { ix += nx; iy += ny; }
  }
  if (lane == 31) { buf[par][0][wid] = ix; buf[par][1][wid] = iy; }
  __syncthreads();
  const int wx = (lane < nw) ? buf[par][0][lane] : 0, wy = (lane < nw) ? buf[par][1][lane] : 0;
  int sx = wx, sy = wy;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int nx = __shfl_up_sync(0xffffffffu, sx, o), ny = __shfl_up_sync(0xffffffffu, sy, o);
    if (lane >= o) { sx += nx; sy += ny; }
  }
  const int bx = __shfl_sync(0xffffffffu, sx - wx, wid), by = __shfl_sync(0xffffffffu, sy - wy, wid);
  *total_x = __shfl_sync(0xffffffffu, sx, 31);
  *total_y = __shfl_sync(0xffffffffu, sy, 31);
  par ^= 1;
  return make_int2(bx + ix - v.x, by + iy - v.y);
}

// 64-bit integer <-> double without the (very slow on sm_100) I2F.F64.S64 / F2I.S64.F64 paths
__device__ __forceinline__ double i64_to_f64(long long v) {
  return __int2double_rn((int)(v >> 32)) * 4294967296.0 + __uint2double_rn((unsigned)v);
}
// imaginary time in [0,1) -> 40-bit fixed point (truncated): mantissa bits of 1 + t
__device__ __forceinline__ long long time_to_fx(double t) {
  return (__double_as_longlong(t + 1.0) & 0x000fffffffffffffll) >> 12;
}

// time window bounds; identical arithmetic on host (bucketing in lq_set_state) and device
__host__ __device__ __forceinline__ double window_lo(int w, int W) { return (double)w / (double)W; }
__host__ __device__ __forceinline__ double window_hi(int w, int W) {
  return (w + 1 >= W) ? 1.0 : (double)(w + 1) / (double)W;
}

}  // namespace lq
