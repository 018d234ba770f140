// lq_engine.cu -- host side of the engine behind include/lq.h: spatial tiling, arenas, the
// per-step kernel schedule, state import/export, and the extern "C" entry points.
//
// Build: alps-looper_b200/csrc/Makefile  (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)
// There is deliberately no CPU path in this file: every entry point that computes needs a CUDA
// device and fails with LQ_E_CUDA otherwise.
#include <chrono>
#include "../../include/lq.h"
#include "lq_kernels.cuh"
#include "lq_space.cuh"
#include "lq_rebucket.cuh"

#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only: libnccl.so.2 is bound at run time (dlopen), so the
                    // library loads and the serial engine runs on a box without NCCL

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>
#include <string>
#include <vector>

namespace {

thread_local std::string g_err;

struct cuda_error { cudaError_t e; const char* what; const char* file; int line; };
#define CK(call)                                                              \
  do {                                                                        \
    cudaError_t e__ = (call);                                                 \
    if (e__ != cudaSuccess) throw cuda_error{e__, #call, __FILE__, __LINE__}; \
  } while (0)

struct lq_error { int code; std::string msg; };
[[noreturn]] void fail(int code, const std::string& m) { throw lq_error{code, m}; }

template <class T>
struct DBuf {  // device array
  T* p = nullptr;
  size_t n = 0;
  void alloc(size_t count, size_t* total) {
    release();
    n = count;
    if (count) {
      CK(cudaMalloc((void**)&p, count * sizeof(T)));
      if (total) *total += count * sizeof(T);
    }
  }
  void upload(const std::vector<T>& v, size_t* total) {
    alloc(v.size(), total);
    if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  void take(DBuf& o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
  ~DBuf() { release(); }
};

// ---------------------------------------------------------------------------------------------
// Spatial tiling: sites are renumbered so that every tile is a contiguous range, bonds so that
// the bonds owned by a tile (those whose source site lies in it) are contiguous.
// ---------------------------------------------------------------------------------------------
struct Partition {
  int N = 0, B = 0, T = 0, nbmax = 0;
  std::vector<int> site_e2i, site_i2e, bond_e2i, bond_i2e;
  std::vector<int> bond_s0, bond_s1, bond_tile, bond_base, adj_off, adj;
  // per-tile halo (bonds owned by other tiles that touch an end site of an owned bond) and the
  // per-bond stencils in LOCAL bucket ids (own bonds 0..nb-1, halo nb..nb+H-1); stencils are
  // shared by all tiles of the same shape ("class")
  int hmax = 0, whmax = 0, nclasses = 0, nksmax = 0, nsmax = 0, zmax = 0;
  std::vector<int> site_base;             // [T+1] first (internal) site of a tile
  std::vector<int> halo_off, halo_bond;   // [T+1], global bond ids of the halo buckets
  std::vector<int> whalo_cnt;             // [T] leading halo buckets that touch an OWN site (walk halo)
  std::vector<int> hsite_off, hsite;      // [T+1], global site ids of the halo K-sites
  std::vector<int> tile_class;            // [T]
  std::vector<int> cls_bs, cls_sso, cls_sst, cls_nks;  // [nclasses] starts in bs / sst_off / sst; K-sites
  std::vector<int> bs;                    // per class [2*(nb+H)] K-site of the two ends of a local bucket (-1: none)
  std::vector<int> sst_off;               // per class [nks+1] offsets into sst
  std::vector<int> sst;                   // (local bucket << 1 | side)
};

// src/dst: end sites of the B internal "bonds"; dst < 0 marks the pseudo-bond that carries the site
// operators of src (graph_impl.h:67-87): it has one end only.
// tile_relabel (spatial cut): a permutation of the tile ids -- every engine numbers its own tiles first,
// then its ghost tiles, so that the pages it holds are the first ones (see SpacePlan below)
void make_partition(const lq_lattice& L, int B, const int* Lsrc, const int* Ldst, int tile_sites, Partition& P,
                    const std::vector<int>* tile_relabel = nullptr) {
  const int N = L.num_sites;
  P.N = N;
  P.B = B;
  std::vector<int> tile_of(N);
  long long prod = 1;
  int nd = 0;
  for (int k = 0; k < 3; ++k)
    if (L.dims[k] > 0) { prod *= L.dims[k]; ++nd; }
  int T = 0;
  if (nd > 0 && prod == N) {
    int dim[3] = {std::max(1, L.dims[0]), std::max(1, L.dims[1]), std::max(1, L.dims[2])};
    int e[3] = {1, 1, 1};
    // distribute the factors of two of tile_sites round-robin over the extended directions
    int budget = std::max(1, tile_sites);
    bool grew = true;
    while (budget > 1 && grew) {
      grew = false;
      for (int k = 0; k < 3 && budget > 1; ++k)
        if (e[k] * 2 <= dim[k]) { e[k] *= 2; budget /= 2; grew = true; }
    }
    int nt[3];
    for (int k = 0; k < 3; ++k) nt[k] = (dim[k] + e[k] - 1) / e[k];
    T = nt[0] * nt[1] * nt[2];
    for (int s = 0; s < N; ++s) {
      int x = s % dim[0], y = (s / dim[0]) % dim[1], z = s / (dim[0] * dim[1]);
      tile_of[s] = (x / e[0]) + nt[0] * ((y / e[1]) + nt[1] * (z / e[2]));
    }
  } else {
    const int S = std::max(1, tile_sites);
    T = (N + S - 1) / S;
    for (int s = 0; s < N; ++s) tile_of[s] = s / S;
  }
  // drop empty tiles (ragged shapes) by compacting tile ids in ascending order
  {
    std::vector<int> tmap(T, 0);
    for (int s = 0; s < N; ++s) tmap[tile_of[s]] = 1;
    int Tc = 0;
    for (int t = 0; t < T; ++t) tmap[t] = tmap[t] ? Tc++ : -1;
    for (int s = 0; s < N; ++s) tile_of[s] = tmap[tile_of[s]];
    P.T = T = Tc;
  }
  if (tile_relabel) {
    if ((int)tile_relabel->size() != T) fail(LQ_E_INVALID, "tile relabelling of the wrong size (internal error)");
    for (int s = 0; s < N; ++s) tile_of[s] = (*tile_relabel)[tile_of[s]];
  }
  // sites
  P.site_i2e.resize(N);
  std::iota(P.site_i2e.begin(), P.site_i2e.end(), 0);
  std::stable_sort(P.site_i2e.begin(), P.site_i2e.end(),
                   [&](int a, int b) { return tile_of[a] < tile_of[b]; });
  P.site_e2i.resize(N);
  for (int i = 0; i < N; ++i) P.site_e2i[P.site_i2e[i]] = i;
  // bonds
  P.bond_i2e.resize(B);
  std::iota(P.bond_i2e.begin(), P.bond_i2e.end(), 0);
  std::stable_sort(P.bond_i2e.begin(), P.bond_i2e.end(),
                   [&](int a, int b) { return tile_of[Lsrc[a]] < tile_of[Lsrc[b]]; });
  P.bond_e2i.resize(B);
  P.bond_s0.resize(B);
  P.bond_s1.resize(B);
  P.bond_tile.resize(B);
  P.bond_base.assign(T + 1, 0);
  for (int i = 0; i < B; ++i) {
    const int e = P.bond_i2e[i];
    P.bond_e2i[e] = i;
    P.bond_s0[i] = P.site_e2i[Lsrc[e]];
    P.bond_s1[i] = Ldst[e] >= 0 ? P.site_e2i[Ldst[e]] : -1;
    P.bond_tile[i] = tile_of[Lsrc[e]];
    P.bond_base[P.bond_tile[i] + 1]++;
  }
  P.nbmax = 0;
  for (int t = 0; t < T; ++t) {
    P.nbmax = std::max(P.nbmax, P.bond_base[t + 1]);
    P.bond_base[t + 1] += P.bond_base[t];
  }
  // adjacency
  P.adj_off.assign(N + 1, 0);
  for (int i = 0; i < B; ++i) { P.adj_off[P.bond_s0[i] + 1]++; if (P.bond_s1[i] >= 0) P.adj_off[P.bond_s1[i] + 1]++; }
  for (int s = 0; s < N; ++s) P.adj_off[s + 1] += P.adj_off[s];
  P.adj.resize((size_t)P.adj_off[N]);
  std::vector<int> fill(P.adj_off.begin(), P.adj_off.end() - 1);
  for (int i = 0; i < B; ++i) {
    P.adj[fill[P.bond_s0[i]]++] = (i << 1) | 0;
    if (P.bond_s1[i] >= 0) P.adj[fill[P.bond_s1[i]]++] = (i << 1) | 1;
  }
  // ---- halos and stencils -------------------------------------------------------------------
  // "K-sites" of a tile = its own sites followed by the far end sites of its owned bonds that lie
  // in other tiles; bucket halo = foreign bonds touching any K-site.  Per tile CLASS (tiles of
  // identical shape share one copy): the two K-sites of every owned bond and, per K-site, the
  // list of incident local buckets (own bonds 0..nb-1, halo buckets nb..) with the side at which
  // the site sits on that bond.
  P.site_base.assign(T + 1, 0);
  for (int i = 0; i < N; ++i) P.site_base[tile_of[P.site_i2e[i]] + 1]++;
  for (int t = 0; t < T; ++t) P.site_base[t + 1] += P.site_base[t];
  P.halo_off.assign(T + 1, 0);
  P.hsite_off.assign(T + 1, 0);
  P.tile_class.assign(T, 0);
  std::map<std::vector<int>, int> classes;
  std::vector<int> local(B, -1), lsite(N, -1);
  for (int t = 0; t < T; ++t) {
    const int b0 = P.bond_base[t], nb = P.bond_base[t + 1] - b0;
    const int s0 = P.site_base[t], ns = P.site_base[t + 1] - s0;
    std::vector<int> halo, hsites, bs, offs(1, 0), ent;
    for (int ls = 0; ls < ns; ++ls) lsite[s0 + ls] = ls;
    for (int lb = 0; lb < nb; ++lb) local[b0 + lb] = lb;
    for (int lb = 0; lb < nb; ++lb)
      for (int side = 0; side < 2; ++side) {
        const int sg = side ? P.bond_s1[b0 + lb] : P.bond_s0[b0 + lb];
        if (sg < 0) { bs.push_back(-1); continue; }
        if (lsite[sg] < 0) { lsite[sg] = ns + (int)hsites.size(); hsites.push_back(sg); }
        bs.push_back(lsite[sg]);
      }
    const int nks = ns + (int)hsites.size();
    int nwalk = 0;
    for (int ks = 0; ks < nks; ++ks) {
      if (ks == ns) nwalk = (int)halo.size();
      const int sg = ks < ns ? s0 + ks : hsites[ks - ns];
      for (int a = P.adj_off[sg]; a < P.adj_off[sg + 1]; ++a) {
        const int b2 = P.adj[a] >> 1, side2 = P.adj[a] & 1;
        if (local[b2] < 0) { local[b2] = nb + (int)halo.size(); halo.push_back(b2); }
        ent.push_back((local[b2] << 1) | side2);
      }
      offs.push_back((int)ent.size());
      P.zmax = std::max(P.zmax, P.adj_off[sg + 1] - P.adj_off[sg]);
    }
    for (int h : halo) {  // K-sites at the two ends of the halo buckets (-1: not a K-site)
      bs.push_back(lsite[P.bond_s0[h]]);
      bs.push_back(P.bond_s1[h] >= 0 ? lsite[P.bond_s1[h]] : -1);
    }
    for (int lb = 0; lb < nb; ++lb) local[b0 + lb] = -1;
    for (int h : halo) local[h] = -1;
    for (int ls = 0; ls < ns; ++ls) lsite[s0 + ls] = -1;
    for (int h : hsites) lsite[h] = -1;
    std::vector<int> sig;
    sig.push_back(nb); sig.push_back(ns); sig.push_back(nks);
    sig.insert(sig.end(), bs.begin(), bs.end());
    sig.insert(sig.end(), offs.begin(), offs.end());
    sig.insert(sig.end(), ent.begin(), ent.end());
    auto it = classes.find(sig);
    if (it == classes.end()) {
      const int c = (int)classes.size();
      classes[sig] = c;
      P.tile_class[t] = c;
      P.cls_bs.push_back((int)P.bs.size());
      P.cls_sso.push_back((int)P.sst_off.size());
      P.cls_sst.push_back((int)P.sst.size());
      P.cls_nks.push_back(nks);
      P.bs.insert(P.bs.end(), bs.begin(), bs.end());
      P.sst_off.insert(P.sst_off.end(), offs.begin(), offs.end());
      P.sst.insert(P.sst.end(), ent.begin(), ent.end());
    } else {
      P.tile_class[t] = it->second;
    }
    if (nks == ns) nwalk = (int)halo.size();
    P.whalo_cnt.push_back(nwalk);
    P.whmax = std::max(P.whmax, nwalk);
    P.halo_off[t + 1] = P.halo_off[t] + (int)halo.size();
    P.halo_bond.insert(P.halo_bond.end(), halo.begin(), halo.end());
    P.hsite_off[t + 1] = P.hsite_off[t] + (int)hsites.size();
    P.hsite.insert(P.hsite.end(), hsites.begin(), hsites.end());
    P.hmax = std::max(P.hmax, (int)halo.size());
    P.nksmax = std::max(P.nksmax, nks);
    P.nsmax = std::max(P.nsmax, ns);
  }
  P.nclasses = (int)classes.size();
}

// ---------------------------------------------------------------------------------------------
// Spatial cut (lq_options.cut = LQ_CUT_SPACE): rank q owns the contiguous range of tiles
// [T q / P, T (q+1) / P) of the GLOBAL tiling G (the same on every rank) over the whole imaginary-time
// axis -- the bond ownership of looper/lattice.h:692-787 taken across GPUs.  It also holds ghost
// copies of the foreign tiles its kernels read:
//   W tiles  hold the far-end sites of its owned bonds; their world lines are walked here too, so
//            that every leg of an OWNED operator gets its lower node locally (no reverse exchange),
//   H tiles  own the foreign bonds that touch one of its K-sites (halo buckets of K1 and of the walk)
//            or a site of a W tile from outside (walk halo of the W tiles).
// A ghost page only holds the buckets of the bonds that touch a K-site of this rank (the bond lists of
// the boundary segments below): their owners send them as dense streams twice per step -- after K1
// (new operators) and before it (the flips of the previous step) -- and k_sp_unpack rebuilds the pages.
// Every engine renumbers the tiles [owned | W | H | rest], so that the pages it holds are the first
// ones and the kernels run unchanged on a prefix of the tiles.
// Clusters that cross a cut are merged through BOUNDARY SEGMENTS: for every ordered pair
// (owner q, user r) the sites of q that are far-end K-sites of r and the bonds of q that touch a
// K-site of r, in a canonical order both sides agree on.  The user holds exact copies of the pages of
// those bonds, so entry i of the segment is the same physical node on both sides.
// ---------------------------------------------------------------------------------------------
struct SpaceSeg {
  int owner = 0, user = 0;
  std::vector<int> sites, bonds;   // EXTERNAL site ids / internal-bond order key -> stored as external ids
  long long off_owner = 0, off_user = 0, cap = 0, opcap = 0;   // (sized with beta, size_space())
};
struct SpacePlan {
  int nranks = 1, rank = 0;
  int To = 0, Tw = 0, Tloc = 0;
  std::vector<int> relabel;                      // global tile -> this rank's tile id
  std::vector<int> rounds;                       // [delta] some rank q has a segment for rank (q+delta)%P: the round exists
  std::vector<SpaceSeg> segs;                    // all segments of the run, canonical order
};

// external ids of the two ends of an internal bond b of partition G: pseudo-bonds (site graphs) are
// identified by Breal + site
void plan_space(const Partition& G, int nranks, int rank, SpacePlan& S) {
  const int T = G.T, P = nranks;
  if (T < P) fail(LQ_E_INVALID, "spatial cut: fewer tiles than ranks (lower lq_options.tile_sites)");
  S.nranks = P; S.rank = rank;
  auto owner = [&](int t) { return (int)(((long long)t * P) / T); };
  std::vector<int> site_tile(G.N);
  for (int t = 0; t < T; ++t)
    for (int s = G.site_base[t]; s < G.site_base[t + 1]; ++s) site_tile[s] = t;
  // per rank: far-end sites F, halo bonds HB, tile sets
  std::vector<std::vector<int>> F(P), HB(P), local(P);
  std::vector<std::vector<int>> relabel(P, std::vector<int>(T, -1));
  std::vector<std::vector<char>> is_w(P, std::vector<char>(T, 0));
  std::vector<int> nown(P, 0), nw(P, 0), nloc(P, 0);
  for (int q = 0; q < P; ++q) {
    std::vector<char> in_w(T, 0), in_h(T, 0);
    for (int t = 0; t < T; ++t) {
      if (owner(t) != q) continue;
      for (int k = G.hsite_off[t]; k < G.hsite_off[t + 1]; ++k)
        if (owner(site_tile[G.hsite[k]]) != q) F[q].push_back(G.hsite[k]);
      for (int k = G.halo_off[t]; k < G.halo_off[t + 1]; ++k)
        if (owner(G.bond_tile[G.halo_bond[k]]) != q) HB[q].push_back(G.halo_bond[k]);
    }
    for (auto* v : {&F[q], &HB[q]}) { std::sort(v->begin(), v->end()); v->erase(std::unique(v->begin(), v->end()), v->end()); }
    for (int s : F[q]) in_w[site_tile[s]] = 1;
    for (int b : HB[q]) in_h[G.bond_tile[b]] = 1;
    for (int t = 0; t < T; ++t)   // walk halo of the W tiles: the leading whalo_cnt halo buckets
      if (in_w[t])
        for (int k = 0; k < G.whalo_cnt[t]; ++k) {
          const int t2 = G.bond_tile[G.halo_bond[G.halo_off[t] + k]];
          if (owner(t2) != q) in_h[t2] = 1;
        }
    int next = 0;
    for (int t = 0; t < T; ++t) if (owner(t) == q) relabel[q][t] = next++;
    nown[q] = next;
    if (next == 0) fail(LQ_E_INVALID, "spatial cut: a rank owns no tile");
    for (int t = 0; t < T; ++t) if (owner(t) != q && in_w[t]) { relabel[q][t] = next++; is_w[q][t] = 1; }
    nw[q] = next - nown[q];
    for (int t = 0; t < T; ++t) if (owner(t) != q && !in_w[t] && in_h[t]) relabel[q][t] = next++;
    nloc[q] = next;
    for (int t = 0; t < T; ++t) if (relabel[q][t] < 0) relabel[q][t] = next++;
  }
  S.To = nown[rank]; S.Tw = nw[rank]; S.Tloc = nloc[rank];
  S.relabel = relabel[rank];
  // boundary segments, canonical order (owner, user); ids leave as EXTERNAL site ids / external bond ids
  // (pseudo-bond of site s: Breal + s, the convention of Partition::bond_i2e)
  S.segs.clear();
  for (int q = 0; q < P; ++q)
    for (int r = 0; r < P; ++r) {
      if (q == r) continue;
      SpaceSeg g;
      g.owner = q; g.user = r;
      for (int s : F[r]) if (owner(site_tile[s]) == q) g.sites.push_back(G.site_i2e[s]);
      for (int b : HB[r]) if (owner(G.bond_tile[b]) == q) g.bonds.push_back(G.bond_i2e[b]);
      if (!g.sites.empty() || !g.bonds.empty()) S.segs.push_back(std::move(g));
    }
  // halo rounds: in round delta every rank sends the stream of its segment for rank+delta and receives
  // the one of rank-delta (every rank makes the same calls, MPI_Sendrecv style)
  S.rounds.assign(P, 0);
  for (const SpaceSeg& g : S.segs) S.rounds[(g.user - g.owner + P) % P] = 1;
}

inline int window_of(double t, int W) {
  int w = (int)(t * W);
  if (w >= W) w = W - 1;
  if (w < 0) w = 0;
  while (w > 0 && lq::window_lo(w, W) > t) --w;
  while (w + 1 < W && lq::window_hi(w, W) <= t) ++w;
  return w;
}

// ---------------------------------------------------------------------------------------------
// NCCL, bound at run time.  In a process that already holds NCCL (torch) dlopen returns that copy.
// ---------------------------------------------------------------------------------------------
struct NcclApi {
  void* lib = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
};
NcclApi& nccl_api() {
  static NcclApi api;
  if (api.lib) return api;
  const char* names[] = {getenv("LQ_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names)
    if (n && *n && (lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
  if (!lib) fail(LQ_E_COMM, std::string("cannot load libnccl.so.2 (set LQ_NCCL_LIB): ") + dlerror());
#define LQ_SYM(f) api.f = (decltype(api.f))dlsym(lib, "nccl" #f); \
  if (!api.f) fail(LQ_E_COMM, "libnccl lacks nccl" #f)
  LQ_SYM(GetUniqueId); LQ_SYM(CommInitRank); LQ_SYM(CommDestroy); LQ_SYM(AllGather); LQ_SYM(AllReduce);
  LQ_SYM(GetErrorString); LQ_SYM(GetVersion); LQ_SYM(Send); LQ_SYM(Recv); LQ_SYM(GroupStart); LQ_SYM(GroupEnd);
#undef LQ_SYM
  api.lib = lib;
  return api;
}
void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) fail(LQ_E_COMM, std::string(what) + ": " + nccl_api().GetErrorString(r));
}

const char* kTimerLabels[17] = {"", "", "", "dispatch", "init", "fill_times(K1 rng)", "init_fragments",
                                "insert/remove+reconnect", "halo exchange (spatial cut)", "close in tau", "boundary estimates", "assign ids",
                                "accumulate", "collect", "flip decision", "flip", "measurement"};

}  // namespace

struct lq_engine {
  // configuration
  Partition part;
  std::vector<double> weights;  // [B][4] external bond order
  std::vector<signed char> gauge_e;
  double energy_offset = 0, beta = 1;
  lq_options opt{};
  int W = 1, Wl = 1, w0 = 0, cap = 0, tpb = 32, npo = 1, ug = 1;
  int Breal = 0;          // bonds of the caller's lattice; internal bonds Breal.. are site pseudo-bonds
  bool has_site = false;
  bool zero_umag = false, zero_ssize = false;   // sums that stay zero for this model (see lq_device.cuh)
  // arena growth after a device-side overflow (the reference's vectors grow on demand,
  // path_integral.C:240-243 RESERVE_*): multipliers on reserve / candidate slots / cluster_reserve
  double grow_pages = 1, grow_cand = 1, grow_clusters = 1;
  int64_t regrows = 0;
  // spatial cut (LQ_CUT_SPACE): tiles / sites / pages this engine holds; serial and slab engines hold everything
  bool space = false;
  SpacePlan plan;
  int Tl = 0, To = 0, Tw = 0;       // local tiles (owned + ghosts), owned, walked ghost tiles
  int Ns = 0, Nown = 0, Nwalk = 0;  // local sites, owned, walked
  size_t Pown = 0;                  // owned pages
  lq::SpDev spd{};
  std::vector<lq::SpSeg> sp_segs_h;        // this rank's segments (host copy; offsets follow beta)
  std::vector<int> sp_seg_index;           // their indices in plan.segs
  DBuf<lq::SpSeg> sp_seg;
  DBuf<lq::SpGSeg> sp_gseg;
  DBuf<int> sp_site, sp_sseg, sp_bond, sp_bseg;
  DBuf<uint32_t> sp_cnt, sp_base, sp_bnode;
  DBuf<uint8_t> sp_pk_spin;
  DBuf<double> sp_pk_time;
  DBuf<uint32_t> sp_pk_info;
  DBuf<int> sp_gt_off, sp_gt_lb, sp_gt_j;
  DBuf<uint32_t> sp_bond_key;
  long long sp_maxcap = 0;
  int gstride() const { return 6 + (has_site ? 1 : 0) + sdim; }
  int sdim = 0;                      // dimensions of the winding estimator (0 = off)
  std::vector<short> bond_vec_e;     // [3 * external bond] components in units of wunit[x]
  double wunit[3] = {0, 0, 0};
  size_t P = 0;
  long long ncap = 0, nccap = 0;
  size_t nwords_cap = 0, device_bytes = 0, static_bytes = 0;
  int sm_count = 0, smem_optin = 0;
  uint32_t mcs = 0;
  int cur = 0;  // live page buffer
  int64_t launches = 0;
  bool timers_on = false;
  // device
  lq::Dev d{};
  cudaStream_t stream = nullptr;
  DBuf<int> bond_s0, bond_s1, bond_base, adj_off, adj, pcount[2], nbase, d_ntotal, d_err;
  DBuf<int> whalo_cnt, bond_tl, halo_tl;
  DBuf<double> bond_emu, wlo, wks;
  DBuf<int> site_base, halo_off, halo_bond, hsite_off, hsite, tile_class, cls_bs, cls_sso, cls_sst, cls_nks, bs, sst_off, sst;
  int scap = 0, ccap = 0;
  size_t stage_smem = 0, walk_smem = 0;
  int tpb_walk = 32;
  typedef void (*walk_fn_t)(lq::Dev, int);
  walk_fn_t walk_fn = nullptr;
  typedef void (*k1_fn_t)(lq::Dev, int, const lq::StepParams*, int, const lq::K1Layout);
  k1_fn_t k1_fn = nullptr;
  lq::K1Layout k1_lay = lq::K1Layout();
  int k1_fc = 12, k1_nt = 256, k1_chunk = 1, kcap = 0;
  bool k1_tma = true;
  bool k1_tq = false;   // LQ_K1_TIMEBITS: the test-hook instantiation of K1 (256 threads, 16-slot columns, plain loads)
  double grow_kept = 1;
  // diagonal update specialised on block size, on the width of the per-site columns and on how the
  // pages reach shared memory (bulk copy + mbarrier, or plain loads)
  k1_fn_t pick_k1() const {
#define LQ_PICK2(MT, F) (k1_tma ? lq::k_diag_update<MT, F, true> : lq::k_diag_update<MT, F, false>)
#define LQ_PICK1(MT) (k1_fc <= 8 ? LQ_PICK2(MT, 8) : k1_fc <= 12 ? LQ_PICK2(MT, 12) : LQ_PICK2(MT, 16))
    if (k1_tq) return lq::k_diag_update<256, 16, false, true>;   // LQ_K1_TIMEBITS (tests)
    if (k1_nt <= 128) return LQ_PICK1(128);
    if (k1_nt <= 256) return LQ_PICK1(256);
    if (k1_nt <= 384) return LQ_PICK1(384);
    return LQ_PICK1(512);
#undef LQ_PICK1
#undef LQ_PICK2
  }
  // world-line walk specialised on block size and coordination number
  walk_fn_t pick_walk() const {
    const int z = part.zmax;
#define LQ_PICKZ(MT, NP) (z <= 2 ? lq::k_walk<MT, 2, NP> : z <= 4 ? lq::k_walk<MT, 4, NP> : z <= 6 ? lq::k_walk<MT, 6, NP> : \
                          z <= 8 ? lq::k_walk<MT, 8, NP> : lq::k_walk<MT, 0, NP>)
#define LQ_PICK(MT) (npo == 2 ? LQ_PICKZ(MT, 2) : LQ_PICKZ(MT, 1))
    if (tpb_walk <= 256) return LQ_PICK(256);
    if (tpb_walk <= 640) return LQ_PICK(640);
    return LQ_PICK(1024);
#undef LQ_PICK
  }
  DBuf<double> bond_rate, time_[2], partial, d_out;
  DBuf<float4> bond_p;
  DBuf<float> bond_q;
  DBuf<signed char> gauge;
  DBuf<uint32_t> info[2], parent, low0, bitmap, wcount, wbase, scan_tmp, d_nc, curW, firstW, labels;
  DBuf<uint16_t> boff[2];
  DBuf<uint8_t> spinW;
  DBuf<uint32_t> flipw, openw;
  DBuf<uint32_t> sse_rank, sse_bincnt, sse_binbase, sse_binfill;   // SSE representation (k_sse_*)
  DBuf<double> sse_time;
  DBuf<uint2> sse_id;
  bool sse = false;
  int nbin = 1;
  DBuf<uint4> rootw;
  DBuf<uint2> xedge;
  DBuf<int> xcount;
  DBuf<short> bond_vec;
  DBuf<int> wind;
  DBuf<unsigned long long> dbgc;
  DBuf<long long> est;
  DBuf<int> est0;
  double* h_out = nullptr;  // pinned
  size_t h_out_n = 0;
  lq::StepParams* h_params = nullptr;  // pinned ring of per-step inputs
  DBuf<lq::StepParams> d_params;
  size_t params_n = 0;
  int64_t h2d_bytes = 0, d2h_bytes = 0;
  size_t nblk_collect = 0;
  lq_comm comm{};
  bool has_comm = false;
  ncclComm_t nccl = nullptr;   // lq_comm_init: the engine's own NCCL communicator (path_integral_mpi.C:75)
  // multi-rank (imaginary-time slabs)
  lq::MrDev mr{};
  DBuf<uint32_t> mr_topmin, mr_sendb, mr_recvb, mr_gparent, mr_gbitmap, mr_gwcount, mr_gwbase, mr_dg;
  DBuf<uint8_t> mr_gused;
  DBuf<long long> mr_gest;
  DBuf<double> mr_rankvec, mr_allvec, mr_gsum;
  uint32_t* h_mr = nullptr;  // pinned: ngc read-back, a ring of 4 steps x 4 words
  cudaEvent_t mr_ev[4] = {nullptr, nullptr, nullptr, nullptr};
  uint64_t mr_step = 0;      // labellings since the arenas were sized / a state was loaded
  size_t mr_gslots = 0;      // open-cluster slots the table can hold
  // timers
  struct TimerAcc { double sec = 0; int count = 0; } tacc[17];
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> tpending;

  ~lq_engine() {
    if (host_timers) {
      std::string m = "lq host enqueue ms per call, rank " + std::to_string(opt.rank) + ":";
      for (int i = 0; i < 17; ++i)
        if (hcnt[i] > 40) m += " [" + std::to_string(i) + "] " + std::to_string(1e3 * hacc[i] / (hcnt[i] - 40)) + " x" + std::to_string(hcnt[i]);
      std::fprintf(stderr, "%s\n", m.c_str());
    }
    if (h_out) cudaFreeHost(h_out);
    if (h_params) cudaFreeHost(h_params);
    if (h_mr) cudaFreeHost(h_mr);
    for (auto& e : mr_ev) if (e) cudaEventDestroy(e);
    if (h_ctl) cudaFreeHost(h_ctl);
    if (nccl) { cudaStreamSynchronize(stream); nccl_api().CommDestroy(nccl); }
    drop_graphs();
    for (auto& t : tpending) { cudaEventDestroy(t.second.first); cudaEventDestroy(t.second.second); }
    if (stream) cudaStreamDestroy(stream);
  }

  // -------------------------------------------------------------------------------------------
  void setup(const lq_lattice& L, const lq_model& M, double beta_, const lq_options& o) {
    if (L.num_sites <= 0 || L.num_bonds < 0 || !L.src || !L.dst) fail(LQ_E_INVALID, "empty lattice");
    if (!(beta_ > 0)) fail(LQ_E_INVALID, "beta must be positive");
    for (int b = 0; b < L.num_bonds; ++b) {
      if (L.src[b] < 0 || L.src[b] >= L.num_sites || L.dst[b] < 0 || L.dst[b] >= L.num_sites)
        fail(LQ_E_INVALID, "bond endpoint out of range");
      if (L.src[b] == L.dst[b]) fail(LQ_E_INVALID, "self-loop bond");
    }
    opt = o;
    if (opt.tile_sites <= 0) opt.tile_sites = 64;
    if (!(opt.window_ops > 0)) opt.window_ops = 3.0;
    if (opt.window_ops > 8.0) fail(LQ_E_INVALID, "window_ops above 8: the candidate draw of one bond and window is capped at 32");
    if (!(opt.reserve > 0)) opt.reserve = 1.7;
    if (!(opt.cluster_reserve > 0)) opt.cluster_reserve = 0.75;
    if (opt.nranks < 1) { opt.nranks = 1; opt.rank = 0; }
    if (opt.rank < 0 || opt.rank >= opt.nranks) fail(LQ_E_INVALID, "rank out of range");
    timers_on = (opt.flags & 1) != 0;
    if (opt.representation != LQ_REPR_PATH_INTEGRAL && opt.representation != LQ_REPR_SSE)
      fail(LQ_E_INVALID, "unknown lq_options.representation");
    sse = opt.representation == LQ_REPR_SSE;
    if (sse && opt.nranks > 1) fail(LQ_E_UNSUPPORTED, "the SSE representation runs on a serial engine (string positions are global)");
    if (opt.cut != LQ_CUT_TIME && opt.cut != LQ_CUT_SPACE) fail(LQ_E_INVALID, "unknown lq_options.cut");
    space = opt.cut == LQ_CUT_SPACE && opt.nranks > 1;
    beta = beta_;
    energy_offset = M.energy_offset;
    weights.resize(4 * (size_t)L.num_bonds);
    for (int b = 0; b < L.num_bonds; ++b)
      for (int g = 0; g < 4; ++g) {
        double v = M.bond_weights ? M.bond_weights[4 * (size_t)b + g] : M.uniform_weights[g];
        if (!(v >= 0)) fail(LQ_E_INVALID, "negative graph weight");
        weights[4 * (size_t)b + g] = v;
      }
    // site graphs (weight_impl.h:62-88: v0 = |Hx|/2): every site gets a one-ended pseudo-bond
    // B + s behind the real bonds; its candidates are always accepted (graph_impl.h:69)
    Breal = L.num_bonds;
    has_site = false;
    std::vector<int> xsrc(L.src, L.src + L.num_bonds), xdst(L.dst, L.dst + L.num_bonds);
    {
      std::vector<double> sw(L.num_sites, M.uniform_site_weight);
      if (M.site_weights) sw.assign(M.site_weights, M.site_weights + L.num_sites);
      for (double v : sw) {
        if (!(v >= 0)) fail(LQ_E_INVALID, "negative site graph weight");
        if (v > 0) has_site = true;
      }
      if (has_site)
        for (int s = 0; s < L.num_sites; ++s) {
          xsrc.push_back(s);
          xdst.push_back(-1);
          weights.push_back(sw[s]);
          weights.push_back(0); weights.push_back(0); weights.push_back(0);
        }
    }
    // relative bond vectors for the winding numbers (stiffness.h:63-76): every component is stored
    // as an integer multiple of the smallest non-zero |component| of its dimension (1/L for the ALPS
    // lattices, where the relative vector is the bond vector over the lattice extent), so the
    // per-cluster windings are exact integer sums
    sdim = 0;
    bond_vec_e.assign(3 * xsrc.size(), 0);
    wunit[0] = wunit[1] = wunit[2] = 0;
    if (L.bond_vectors && L.vector_dim > 0) {
      sdim = std::min(3, (int)L.vector_dim);   // stiffness.h:68-73 caps at MAX_DIM
      for (int x = 0; x < 3; ++x) {
        double u = 0;
        for (int b = 0; b < L.num_bonds; ++b) {
          const double a = std::fabs(L.bond_vectors[3 * (size_t)b + x]);
          if (a > 1e-12 && (u == 0 || a < u)) u = a;
        }
        wunit[x] = u;
        for (int b = 0; b < L.num_bonds && u > 0; ++b) {
          const double m = L.bond_vectors[3 * (size_t)b + x] / u;
          if (!(std::fabs(m) < 32000) || std::fabs(m - std::nearbyint(m)) > 1e-6)
            fail(LQ_E_UNSUPPORTED, "bond vectors of one dimension are not integer multiples of the smallest one");
          bond_vec_e[3 * (size_t)b + x] = (short)std::nearbyint(m);
        }
      }
    }
    gauge_e.assign(L.num_sites, 0);
    if (L.gauge)
      for (int s = 0; s < L.num_sites; ++s) gauge_e[s] = (signed char)(L.gauge[s] > 0 ? 1 : (L.gauge[s] < 0 ? -1 : 0));

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      fail(LQ_E_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                          " (this engine has no CPU fallback)");
    CK(cudaSetDevice(opt.device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, opt.device));
    sm_count = prop.multiProcessorCount;
    smem_optin = (int)prop.sharedMemPerBlockOptin;
    CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    graphs_off = getenv("LQ_NO_GRAPH") != nullptr;

    if (space) {
      // the global tiling (the same on every rank) decides who owns what; this engine then numbers
      // its own tiles first and its ghost tiles behind them
      Partition G;
      make_partition(L, (int)xsrc.size(), xsrc.data(), xdst.data(), opt.tile_sites, G);
      plan_space(G, opt.nranks, opt.rank, plan);
      make_partition(L, (int)xsrc.size(), xsrc.data(), xdst.data(), opt.tile_sites, part, &plan.relabel);
      Tl = plan.Tloc; To = plan.To; Tw = plan.Tw;
      Ns = part.site_base[Tl]; Nown = part.site_base[To]; Nwalk = part.site_base[To + Tw];
    } else {
      make_partition(L, (int)xsrc.size(), xsrc.data(), xdst.data(), opt.tile_sites, part);
      Tl = To = part.T; Tw = 0;
      Ns = Nown = Nwalk = part.N;
    }
    if (part.nbmax + 1 > 1024)
      fail(LQ_E_INVALID, "tile owns more than 1023 bonds: lower lq_options.tile_sites");
    if (part.hmax > 1024)
      fail(LQ_E_INVALID, "tile halo has more than 1024 buckets: lower lq_options.tile_sites");
    if (part.T >= (1 << 21)) fail(LQ_E_INVALID, "more than 2^21 tiles: raise lq_options.tile_sites");
    if (part.nksmax > 1024) fail(LQ_E_INVALID, "tile touches more than 1024 sites: lower lq_options.tile_sites");
    // K1: a thread per own bucket / halo bucket / K-site.  (nbmax + 1 keeps 17 warps for the usual
    // 512-bond tile; exactly 16 warps measured 3 % slower -- the flat phases lose a warp)
    tpb = ((std::max(std::max(part.nbmax + 1, part.hmax), part.nksmax) + 31) / 32) * 32;

    // static tables
    const int N = part.N, B = part.B;
    std::vector<double> rate(B);
    std::vector<float4> bp(B);
    std::vector<float> bq(B);
    npo = 1;
    for (int i = 0; i < B; ++i) {
      const double* v = &weights[4 * (size_t)part.bond_i2e[i]];
      const double sum = v[0] + v[1] + v[2] + v[3];
      rate[i] = sum;
      if (v[1] > 0) npo = 2;
      if (sum > 0) bp[i] = make_float4((float)(v[0] / sum), (float)((v[0] + v[2]) / sum),
                                       (float)(v[1] / sum), (float)((v[1] + v[3]) / sum));
      else bp[i] = make_float4(0, 0, 0, 0);
      bq[i] = (v[0] + v[1] > 0) ? (float)(v[0] / (v[0] + v[1])) : 1.0f;
      if (v[1] == 0) bq[i] = 1.0f;
    }
    std::vector<signed char> gi(N);
    for (int i = 0; i < N; ++i) gi[i] = gauge_e[part.site_i2e[i]];
    zero_umag = zero_ssize = !has_site;
    for (int i = 0; i < B; ++i) {
      const double* v = &weights[4 * (size_t)part.bond_i2e[i]];
      if (v[1] > 0 || v[3] > 0) zero_umag = false;                          // graphs on parallel spins
      if (v[1] > 0) zero_ssize = false;   // cross graph: the two lower legs belong to different clusters
      if (part.bond_s1[i] >= 0 && gi[part.bond_s0[i]] + gi[part.bond_s1[i]] != 0) zero_ssize = false;
    }
    bond_s0.upload(part.bond_s0, &device_bytes);
    bond_s1.upload(part.bond_s1, &device_bytes);
    bond_base.upload(part.bond_base, &device_bytes);
    adj_off.upload(part.adj_off, &device_bytes);
    adj.upload(part.adj, &device_bytes);
    bond_rate.upload(rate, &device_bytes);
    bond_p.upload(bp, &device_bytes);
    bond_q.upload(bq, &device_bytes);
    gauge.upload(gi, &device_bytes);
    {
      std::vector<short> bv(3 * (size_t)B, 0);
      for (int i = 0; i < B; ++i)
        for (int x = 0; x < 3; ++x) bv[3 * (size_t)i + x] = bond_vec_e[3 * (size_t)part.bond_i2e[i] + x];
      bond_vec.upload(bv, &device_bytes);
    }
    whalo_cnt.upload(part.whalo_cnt, &device_bytes);
    {
      std::vector<int> tl(B);
      for (int i = 0; i < B; ++i) tl[i] = (part.bond_tile[i] << 10) | (i - part.bond_base[part.bond_tile[i]]);
      bond_tl.upload(tl, &device_bytes);
      std::vector<int> htl(part.halo_bond.size());   // the same for every halo bucket: one dependent load less per stage
      for (size_t k = 0; k < htl.size(); ++k) htl[k] = tl[part.halo_bond[k]];
      halo_tl.upload(htl, &device_bytes);
    }
    site_base.upload(part.site_base, &device_bytes);
    halo_off.upload(part.halo_off, &device_bytes);
    halo_bond.upload(part.halo_bond, &device_bytes);
    hsite_off.upload(part.hsite_off, &device_bytes);
    hsite.upload(part.hsite, &device_bytes);
    tile_class.upload(part.tile_class, &device_bytes);
    cls_bs.upload(part.cls_bs, &device_bytes);
    cls_sso.upload(part.cls_sso, &device_bytes);
    cls_sst.upload(part.cls_sst, &device_bytes);
    cls_nks.upload(part.cls_nks, &device_bytes);
    bs.upload(part.bs, &device_bytes);
    sst_off.upload(part.sst_off, &device_bytes);
    sst.upload(part.sst, &device_bytes);
    if (space) {
      // boundary segments this rank takes part in: site / bond entries in the rank's own numbering
      std::vector<int> hs, hss, hb, hbs;
      for (size_t k = 0; k < plan.segs.size(); ++k) {
        const SpaceSeg& g = plan.segs[k];
        if (g.owner != opt.rank && g.user != opt.rank) continue;
        lq::SpSeg x{};
        x.ns = (int)g.sites.size(); x.nb = (int)g.bonds.size();
        x.site0 = (int)hs.size(); x.bond0 = (int)hb.size();
        x.user_side = g.user == opt.rank ? 1 : 0;
        for (int se : g.sites) { hs.push_back(part.site_e2i[se]); hss.push_back((int)sp_segs_h.size()); }
        for (int be : g.bonds) { hb.push_back(part.bond_e2i[be]); hbs.push_back((int)sp_segs_h.size()); }
        for (size_t i = (size_t)x.site0; i < hs.size(); ++i)
          if (hs[i] >= Ns) fail(LQ_E_INVALID, "spatial cut: boundary site outside the local tiles (internal error)");
        for (size_t i = (size_t)x.bond0; i < hb.size(); ++i)
          if (part.bond_tile[hb[i]] >= Tl) fail(LQ_E_INVALID, "spatial cut: boundary bond outside the local tiles (internal error)");
        sp_segs_h.push_back(x);
        sp_seg_index.push_back((int)k);
      }
      if (hs.empty()) hs.push_back(0), hss.push_back(0);
      if (hb.empty()) hb.push_back(0), hbs.push_back(0);
      spd.nseg = (int)sp_segs_h.size();
      spd.nst = 0; spd.nbt = 0;
      for (const auto& x : sp_segs_h) { spd.nst += x.ns; spd.nbt += x.nb; }
      sp_site.upload(hs, &device_bytes); sp_sseg.upload(hss, &device_bytes);
      sp_bond.upload(hb, &device_bytes); sp_bseg.upload(hbs, &device_bytes);
      {
        // per ghost tile: the buckets its owner sends (bonds of the user-side segments), ascending
        std::vector<std::vector<std::pair<int, int>>> need((size_t)std::max(0, Tl - To));
        for (size_t i = 0; i < sp_segs_h.size(); ++i) {
          if (!sp_segs_h[i].user_side) continue;
          for (int j = sp_segs_h[i].bond0; j < sp_segs_h[i].bond0 + sp_segs_h[i].nb; ++j) {
            const int t = part.bond_tile[hb[j]];
            if (t < To) fail(LQ_E_INVALID, "spatial cut: ghost bond in an owned tile (internal error)");
            need[t - To].push_back({hb[j] - part.bond_base[t], j});
          }
        }
        std::vector<int> go(1, 0), glb, gj;
        for (auto& v : need) {
          std::sort(v.begin(), v.end());
          for (auto& x : v) { glb.push_back(x.first); gj.push_back(x.second); }
          go.push_back((int)glb.size());
        }
        if (glb.empty()) { glb.push_back(0); gj.push_back(0); }
        sp_gt_off.upload(go, &device_bytes); sp_gt_lb.upload(glb, &device_bytes); sp_gt_j.upload(gj, &device_bytes);
      }
      // Philox counters of K1: a tile id every rank agrees on (two ranks must not draw the same
      // stream for their tiles number 0, 1, ...): the tile's id in the global tiling
      std::vector<uint32_t> key(part.T);
      for (int t = 0; t < part.T; ++t) key[plan.relabel[t]] = (uint32_t)t;
      sp_bond_key.upload(key, &device_bytes);
    }
    d_ntotal.alloc(1, &device_bytes);
    d_err.alloc(1, &device_bytes);
    d_nc.alloc(2, &device_bytes);
    CK(cudaMemset(d_err.p, 0, sizeof(int)));
    CK(cudaMemset(d_ntotal.p, 0, sizeof(int)));
    CK(cudaMemset(d_nc.p, 0, 2 * sizeof(uint32_t)));
    size_arenas();
    // initial state: all up, no operators (path_integral.C:225)
    clear_state();
  }

  // (re)size everything that depends on beta
  void size_arenas() {
    const int N = part.N, B = part.B, T = part.T;
    if (static_bytes == 0) static_bytes = device_bytes;   // lattice tables, sized once in setup()
    device_bytes = static_bytes;                          // the arenas below are counted afresh
    drop_graphs();   // captured kernels hold the old pointers and grid sizes
    double maxrate = 0;
    std::vector<double> tile_rate(T, 0.0);
    for (int i = 0; i < B; ++i) {
      const double* v = &weights[4 * (size_t)part.bond_i2e[i]];
      const double sum = v[0] + v[1] + v[2] + v[3];
      maxrate = std::max(maxrate, sum);
      tile_rate[part.bond_tile[i]] += sum;
    }
    W = (int)std::ceil(beta * maxrate / opt.window_ops);
    if (W < 1) W = 1;
    if (space) { Wl = W; w0 = 0; }   // every rank holds the whole imaginary-time axis of its tiles
    else {
      W = ((W + opt.nranks - 1) / opt.nranks) * opt.nranks;
      Wl = W / opt.nranks;
      w0 = opt.rank * Wl;
    }
    double mu = 0;
    for (int t = 0; t < T; ++t) mu = std::max(mu, tile_rate[t] * beta / W);
    {
      std::vector<double> emu(B);
      for (int i = 0; i < B; ++i) {
        const double* v = &weights[4 * (size_t)part.bond_i2e[i]];
        emu[i] = std::exp(-beta * (v[0] + v[1] + v[2] + v[3]) / W);
      }
      bond_emu.upload(emu, nullptr);
      std::vector<double> wl(W + 1);
      for (int w = 0; w < W; ++w) wl[w] = lq::window_lo(w, W);
      wl[W] = 1.0;   // == window_hi(W - 1, W); window_hi(w, W) == window_lo(w + 1, W) below it
      wlo.upload(wl, nullptr);
      std::vector<double> ks(W);
      for (int w = 0; w < W; ++w) ks[w] = 4294967040.0 / (wl[w + 1] - wl[w]);
      wks.upload(ks, nullptr);
    }
    const double m = opt.reserve * grow_pages * mu;
    long long c = (long long)std::ceil(m + 6.0 * std::sqrt(m) + 16.0);
    c = (c + 3) & ~3ll;   // pages start on 16-byte boundaries (bulk copies of K1)
    if (c > 65532) fail(LQ_E_INVALID, "page capacity exceeds 65532 operators: lower tile_sites or window_ops");
    cap = (int)c;
    {
      // shared-memory stage: own page + halo buckets (halo/own bucket ratio, 1.5x head room)
      const double ratio = (double)part.hmax / std::max(1, part.nbmax);
      const double hm = m * ratio;
      scap = cap + (int)std::ceil(hm + 6.0 * std::sqrt(hm) + 16.0);
      const double cm = mu;  // mean candidates per page
      ccap = (int)std::ceil(grow_cand * (cm + 8.0 * std::sqrt(cm) + 32.0));
      if (ccap > 32767 || scap > 65535 || cap > 65532)
        fail(LQ_E_INVALID, "too many candidates / staged operators per page: lower tile_sites or window_ops");
      tpb_walk = ((std::max(part.nsmax, part.whmax) + 31) / 32) * 32;
      // K1 (lq_k1.cuh).  Column width: mean off-diagonal legs per site and window ~ 0.6 z window_ops
      {
        const double ml = 0.45 * part.zmax * opt.window_ops;   // (measured: 5.3 on the square-lattice Heisenberg workloads)
        const double want = ml + 2.5 * std::sqrt(ml);
        // (a column that overflows sends every candidate on the site through the exact path, 400 warp
        // instructions at 4 lanes: one step wider than the estimate asks for -- 0.6 % of the sites
        // overflowed 12 slots at a mean of 5.3 legs, profiles/r02_k1.md)
        k1_fc = want <= 6 ? 8 : (want <= 9 ? 12 : 16);
        if (getenv("LQ_FC")) k1_fc = atoi(getenv("LQ_FC")) <= 8 ? 8 : (atoi(getenv("LQ_FC")) <= 12 ? 12 : 16);
        // kept list: the off-diagonal operators of a page (half a page to start with; slab engines,
        // which cannot rewind, take the whole page)
        kcap = (int)std::min<double>(cap, std::ceil((opt.nranks > 1 ? 1.0 : 0.55) * grow_kept * cap) + 16);
        // threads: a thread owns up to LQ_K1_QMAX buckets; one bucket per thread up to 512
        k1_nt = part.nbmax <= 128 ? 128 : (part.nbmax <= 256 ? 256 : 512);
        if (getenv("LQ_K1_NT")) k1_nt = atoi(getenv("LQ_K1_NT"));
        k1_nt = k1_nt <= 128 ? 128 : (k1_nt <= 256 ? 256 : (k1_nt <= 384 ? 384 : 512));
        while (part.nbmax > LQ_K1_QMAX * k1_nt && k1_nt < 512) k1_nt = k1_nt < 256 ? 256 : (k1_nt < 384 ? 384 : 512);
        if (part.nbmax > LQ_K1_QMAX * k1_nt) fail(LQ_E_INVALID, "tile owns too many bonds for K1: lower lq_options.tile_sites");
        k1_tma = !getenv("LQ_K1_NOTMA");
        k1_tq = getenv("LQ_K1_TIMEBITS") && atoi(getenv("LQ_K1_TIMEBITS")) > 0;
        if (k1_tq) {
          if (part.nbmax > LQ_K1_QMAX * 256) fail(LQ_E_INVALID, "LQ_K1_TIMEBITS (test hook) needs tiles of at most 1280 bonds");
          k1_nt = 256; k1_fc = 16; k1_tma = false;
        }
        // two resident CTAs per SM matter more than wide columns or the bulk-copied page buffer
        const size_t two_ctas = (size_t)(227 * 1024) / 2 - 1200;
        auto k1_bytes = [&]() { return lq::k1_smem_bytes(k1_tma, k1_fc, cap, ccap, kcap, part.nbmax, part.hmax, part.nksmax); };
        while (k1_fc > 8 && k1_bytes() > two_ctas && !getenv("LQ_FC") && !k1_tq) k1_fc -= 4;
        if (k1_bytes() > (size_t)smem_optin - 2048)
          k1_tma = false;   // page buffer does not fit beside the lists: stream the page with plain loads
        // windows per persistent CTA: enough CTAs for ~8 waves, at least 4 windows to amortise the tile set-up
        {
          const long long want_ctas = 8ll * sm_count * 3;
          const int Tk = To;   // (the tiles K1 runs on: the owned ones)
          long long nch = std::max<long long>(1, std::min<long long>(Wl, (want_ctas + Tk - 1) / Tk));
          k1_chunk = (int)((Wl + nch - 1) / nch);
          if (k1_chunk < 4) k1_chunk = std::min(4, Wl);
          if (getenv("LQ_K1_CHUNK")) k1_chunk = std::max(1, std::min(Wl, atoi(getenv("LQ_K1_CHUNK"))));
        }
      }
      stage_smem = lq::k1_smem_bytes(k1_tma, k1_fc, cap, ccap, kcap, part.nbmax, part.hmax, part.nksmax);
      k1_lay = lq::k1_layout(k1_tma, k1_fc, cap, ccap, kcap, part.nbmax, part.hmax, part.nksmax);
      walk_smem = lq::stage_bytes(scap, part.nbmax, part.hmax, part.zmax, tpb_walk);
      if (stage_smem > (size_t)smem_optin - 2048 || walk_smem > (size_t)smem_optin - 2048 ||
          (size_t)npo * cap * sizeof(uint32_t) > (size_t)smem_optin - 2048)
        fail(LQ_E_INVALID, "page + halo do not fit shared memory: lower tile_sites or window_ops");
      // The dynamic shared-memory limit is an attribute of the kernel, shared by every engine of the
      // process: always raise it to the device maximum (the launches pass their own sizes) so that
      // engines of different shapes can coexist.
      const int sm = smem_optin - 2048;   // (the limit covers static + dynamic shared memory)
      if (sdim > 0) {
        CK(cudaFuncSetAttribute(lq::k_estimate<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CK(cudaFuncSetAttribute(lq::k_estimate<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CK(cudaFuncSetAttribute(lq::k_estimate<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
        CK(cudaFuncSetAttribute(lq::k_estimate<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
      }
      k1_fn = pick_k1();
      CK(cudaFuncSetAttribute(k1_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
      // union groups: consecutive windows of a tile unified in one shared-memory union-find
      // (measured at 512x512 beta=128, ps per operator of walk + unions: 1 window 42.7, 2 windows
      // 39.6, 3 windows 40.0, 6 windows 46.4 -- deeper shared-memory trees and fewer CTAs eat the
      // gain; LQ_UG overrides for experiments)
      ug = 2;
      if (getenv("LQ_UG")) ug = std::max(1, atoi(getenv("LQ_UG")));
      ug = (int)std::min<size_t>((size_t)ug, std::max<size_t>(1, (96 * 1024) / ((size_t)npo * cap * sizeof(uint32_t))));
      ug = std::min(ug, Wl);
      CK(cudaFuncSetAttribute(lq::k_union_local, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              sm - (int)(LQ_XCAP * sizeof(uint2)) - 64));   // minus its static edge list
      walk_fn = pick_walk();
      CK(cudaFuncSetAttribute(walk_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    }
    P = (size_t)Tl * Wl;
    Pown = (size_t)To * Wl;
    ncap = (long long)P * cap;
    const long long nodes_cap = (long long)Ns + (long long)npo * ncap;
    if (nodes_cap >= 0x7ffffff0ll) fail(LQ_E_INVALID, "more than 2^31 graph nodes on one GPU: lower lq_options.reserve or split the run over more GPUs");
    nccap = std::min<long long>(nodes_cap, (long long)Ns + (long long)std::ceil(std::min(1.0, opt.cluster_reserve * grow_clusters) * (double)(npo * ncap)));
    nwords_cap = (size_t)((nodes_cap + 31) / 32);

    size_t* tb = &device_bytes;
    for (int k = 0; k < 2; ++k) {
      time_[k].alloc((size_t)ncap, tb);
      info[k].alloc((size_t)ncap, tb);
      boff[k].alloc(P * (size_t)(part.nbmax + 1), tb);
      pcount[k].alloc(P, tb);
    }
    nbase.alloc(P + 1, tb);
    spinW.alloc((size_t)(Wl + 1) * Ns, tb);
    curW.alloc((size_t)(Wl + 1) * Ns, tb);
    firstW.alloc((size_t)Wl * Ns, tb);
    parent.alloc((size_t)nodes_cap, tb);
    low0.alloc(2 * (size_t)ncap, tb);   // low1 = low0 + ncap: ONE array, so the walk selects the side with an index offset
    bitmap.alloc(nwords_cap + 1, tb);
    wcount.alloc(nwords_cap + 1, tb);
    wbase.alloc(nwords_cap + 1, tb);
    rootw.alloc(opt.nranks == 1 ? nwords_cap + 1 : 1, tb);
    {
      const size_t ngroups = (size_t)To * ((Wl + ug - 1) / ug);
      xedge.alloc(ngroups * LQ_XCAP, tb);
      xcount.alloc(ngroups, tb);
    }
    if (sse) {
      // ~4 operators per time bin at full pages
      nbin = 1;
      while ((long long)nbin * 4 < (long long)part.T * cap && nbin < (1 << 24)) nbin <<= 1;
      const size_t nb = (size_t)Wl * nbin;
      sse_rank.alloc((size_t)ncap, tb);
      sse_time.alloc((size_t)ncap, tb);
      sse_id.alloc((size_t)ncap, tb);
      sse_bincnt.alloc(nb + 1, tb); sse_binbase.alloc(nb + 2, tb); sse_binfill.alloc(nb + 1, tb);
    }
    if (space) size_space();
    const size_t scan_n = std::max(std::max(std::max(nwords_cap, P), sse ? (size_t)Wl * nbin : (size_t)0),
                                   space ? std::max((size_t)spd.nbt * Wl + 2, (size_t)(opt.nranks * spd.stride) / 32 + 2) : (size_t)0) + 1;
    scan_tmp.alloc((scan_n + LQ_SCAN_CHUNK - 1) / LQ_SCAN_CHUNK + 1, tb);
    est.alloc(4 * (size_t)nccap, tb);
    est0.alloc(4 * (size_t)Ns, tb);
    flipw.alloc((size_t)nccap / 32 + 2, tb);
    wind.alloc(sdim > 0 ? (size_t)sdim * (size_t)nccap : 1, tb);
    CK(cudaMemset(wind.p, 0, wind.n * sizeof(int)));
    openw.alloc(has_site ? (size_t)nccap / 32 + 2 : 1, tb);
    CK(cudaMemset(openw.p, 0, openw.n * sizeof(uint32_t)));
    nblk_collect = std::min<size_t>(((size_t)nccap + 255) / 256, (size_t)sm_count * 8);
    partial.alloc(nblk_collect * LQ_NSUM, tb);
    if (opt.nranks > 1) {
      // gathered boundary forest: slabs 2N nodes per rank (bottom + top of every world line); spatial
      // cut `stride` words per rank (header + the entries of its boundary segments)
      const size_t per_rank = space ? (size_t)spd.stride : 2 * (size_t)N;
      const size_t g2 = (size_t)opt.nranks * per_rank, gw = (g2 + 31) / 32 + 1;
      mr_topmin.alloc((size_t)nccap, tb);
      mr_sendb.alloc(per_rank, tb);
      mr_recvb.alloc(g2, tb);
      mr_gparent.alloc(g2, tb);
      mr_gused.alloc(g2, tb);
      mr_gbitmap.alloc(gw, tb); mr_gwcount.alloc(gw, tb); mr_gwbase.alloc(gw + 1, tb);
      // (spatial cut: every global open cluster has at least two entries, one per side of a cut)
      mr_gslots = space ? g2 / 2 + 1 : g2;
      mr_gest.alloc(mr_gslots * (size_t)gstride() + 32 * (size_t)opt.nranks, tb);   // the collectors' slots (k_mr_rankvec) come first
      mr_dg.alloc(4, tb);
      mr_rankvec.alloc(32, tb); mr_allvec.alloc(32 * (size_t)opt.nranks, tb); mr_gsum.alloc(16, tb);
      CK(cudaMemset(mr_topmin.p, 0xff, mr_topmin.n * sizeof(uint32_t)));
      CK(cudaMemset(mr_gest.p, 0, mr_gest.n * sizeof(long long)));
      CK(cudaMemset(mr_dg.p, 0, 4 * sizeof(uint32_t)));
      if (!h_mr) CK(cudaMallocHost((void**)&h_mr, 16 * sizeof(uint32_t)));
      for (auto& e : mr_ev) if (!e) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      mr_step = 0;
      mr.topmin = mr_topmin.p; mr.sendb = mr_sendb.p; mr.recvb = mr_recvb.p; mr.gparent = mr_gparent.p;
      mr.gused = mr_gused.p; mr.gbitmap = mr_gbitmap.p; mr.gwcount = mr_gwcount.p; mr.gwbase = mr_gwbase.p;
      mr.gest = mr_gest.p + 32 * (size_t)opt.nranks; mr.d_g = mr_dg.p; mr.rankvec = mr_rankvec.p; mr.allvec = mr_allvec.p;
      mr.gsum = mr_gsum.p;
      mr.gn = g2;
      mr.gbase = space ? (size_t)opt.rank * per_rank + LQ_SP_HDR : (size_t)opt.rank * per_rank;
      if (space) {
        sp_cnt.alloc((size_t)spd.nbt * Wl + 2, tb);
        sp_base.alloc((size_t)spd.nbt * Wl + 3, tb);
        sp_bnode.alloc((size_t)std::max<long long>(spd.cbmax, 1), tb);
        CK(cudaMemset(sp_cnt.p, 0, sp_cnt.n * sizeof(uint32_t)));
        spd.cnt = sp_cnt.p; spd.base = sp_base.p; spd.bnode = sp_bnode.p;
        spd.site = sp_site.p; spd.sseg = sp_sseg.p; spd.bond = sp_bond.p; spd.bseg = sp_bseg.p;
        spd.seg = sp_seg.p; spd.gseg = sp_gseg.p;
        // halo streams (k_sp_pack / k_sp_unpack): operators and spins of this rank's segments
        long long pk = 0, spb = 0;
        for (auto& x : sp_segs_h) { pk += x.opcap; spb += (long long)x.ns * (Wl + 1); }
        sp_pk_time.alloc((size_t)std::max<long long>(pk, 1), tb);
        sp_pk_info.alloc((size_t)std::max<long long>(pk, 1), tb);
        sp_pk_spin.alloc((size_t)std::max<long long>(spb, 1), tb);
        spd.pk_time = sp_pk_time.p; spd.pk_info = sp_pk_info.p; spd.pk_spin = sp_pk_spin.p;
        spd.gt_off = sp_gt_off.p; spd.gt_lb = sp_gt_lb.p; spd.gt_j = sp_gt_j.p;
      }
    }
    CK(cudaMemset(est.p, 0, est.n * sizeof(long long)));
    CK(cudaMemset(est0.p, 0, est0.n * sizeof(int)));
    CK(cudaMemset(bitmap.p, 0, bitmap.n * sizeof(uint32_t)));
    CK(cudaMemset(wcount.p, 0, wcount.n * sizeof(uint32_t)));
    fill_dev();
  }

  // boundary segments of the spatial cut: capacities follow beta (like the pages), offsets follow the
  // capacities; every rank computes the layout of ALL ranks' buffers (the merge is redundant)
  void size_space() {
    std::vector<long long> fill(opt.nranks, 0);
    std::vector<lq::SpGSeg> gs;
    sp_maxcap = 0;
    for (auto& g : plan.segs) {
      double rate = 0;
      for (int be : g.bonds) { const double* v = &weights[4 * (size_t)be]; rate += v[0] + v[1] + v[2] + v[3]; }
      const double m = opt.reserve * grow_pages * beta * rate;
      g.opcap = (long long)std::ceil(m + 6.0 * std::sqrt(m) + 16.0);
      g.cap = (long long)g.sites.size() + (long long)npo * g.opcap;
      g.off_owner = fill[g.owner]; fill[g.owner] += g.cap;
      g.off_user = fill[g.user]; fill[g.user] += g.cap;
      gs.push_back({g.owner, g.user, g.off_owner, g.off_user, g.cap});
      sp_maxcap = std::max(sp_maxcap, g.cap);
    }
    spd.cb = fill[opt.rank];
    spd.cbmax = *std::max_element(fill.begin(), fill.end());
    spd.stride = LQ_SP_HDR + spd.cbmax;
    if ((long long)opt.nranks * spd.stride >= 0x7ffffff0ll) fail(LQ_E_INVALID, "spatial cut: boundary too large");
    spd.ngseg = (int)gs.size();
    if (gs.empty()) gs.push_back({0, 0, 0, 0, 0});
    long long pk = 0, spb = 0;
    for (size_t i = 0; i < sp_segs_h.size(); ++i) {
      const SpaceSeg& g = plan.segs[sp_seg_index[i]];
      sp_segs_h[i].off = sp_segs_h[i].user_side ? g.off_user : g.off_owner;
      sp_segs_h[i].opcap = g.opcap;
      sp_segs_h[i].pk0 = pk; pk += g.opcap;
      sp_segs_h[i].sp0 = spb; spb += (long long)sp_segs_h[i].ns * (Wl + 1);
    }
    std::vector<lq::SpSeg> hseg = sp_segs_h;
    if (hseg.empty()) hseg.push_back(lq::SpSeg{});
    sp_seg.upload(hseg, nullptr);
    sp_gseg.upload(gs, nullptr);
  }

  void fill_dev() {
    d.N = Ns; d.B = part.B; d.T = Tl; d.nbmax = part.nbmax;
    d.tile_key = space ? sp_bond_key.p : nullptr;
    d.space = space ? 1 : 0; d.Nown = Nown; d.Nwalk = Nwalk; d.Pown = (int)Pown;
    d.W = W; d.w0 = w0; d.Wl = Wl; d.cap = cap; d.npo = npo; d.ug = ug; d.has_site = has_site ? 1 : 0;
    d.zero_umag = zero_umag ? 1 : 0; d.zero_ssize = zero_ssize ? 1 : 0;
    d.rank = opt.rank; d.nranks = opt.nranks;
    d.bond_s0 = bond_s0.p; d.bond_s1 = bond_s1.p; d.bond_base = bond_base.p;
    d.adj_off = adj_off.p; d.adj = adj.p; d.bond_rate = bond_rate.p; d.bond_p = bond_p.p;
    d.bond_q = bond_q.p; d.gauge = gauge.p;
    d.whalo_cnt = whalo_cnt.p; d.bond_tl = bond_tl.p; d.halo_tl = halo_tl.p; d.bond_emu = bond_emu.p; d.wlo = wlo.p; d.wks = wks.p;
    d.site_base = site_base.p; d.halo_off = halo_off.p; d.halo_bond = halo_bond.p;
    d.hsite_off = hsite_off.p; d.hsite = hsite.p; d.tile_class = tile_class.p;
    d.cls_bs = cls_bs.p; d.cls_sso = cls_sso.p; d.cls_sst = cls_sst.p; d.cls_nks = cls_nks.p;
    d.bs = bs.p; d.sst_off = sst_off.p; d.sst = sst.p;
    d.hmax = part.hmax; d.nksmax = part.nksmax; d.zmax = part.zmax; d.scap = scap; d.ccap = ccap; d.kcap = kcap;
    d.k1_keyshift = getenv("LQ_K1_KEYBITS") ? std::max(0, std::min(31, 32 - atoi(getenv("LQ_K1_KEYBITS")))) : 0;
    d.k1_timebits = getenv("LQ_K1_TIMEBITS") ? std::max(0, std::min(30, atoi(getenv("LQ_K1_TIMEBITS")))) : 0;
    for (int k = 0; k < 2; ++k) {
      d.time[k] = time_[k].p; d.info[k] = info[k].p; d.boff[k] = boff[k].p; d.pcount[k] = pcount[k].p;
    }
    d.nbase = nbase.p; d.spinW = spinW.p; d.curW = curW.p; d.firstW = firstW.p; d.parent = parent.p; d.low0 = low0.p;
    d.low1 = low0.p + (size_t)ncap; d.lowstride = (uint32_t)ncap; d.bitmap = bitmap.p; d.wcount = wcount.p; d.wbase = wbase.p;
    d.rootw = rootw.p; d.fpack = 1; d.rootflip = opt.nranks == 1 ? 1 : 0;
    d.xedge = xedge.p; d.xcount = xcount.p;
    d.xcap = getenv("LQ_XCAP") ? std::max(0, std::min(LQ_XCAP, atoi(getenv("LQ_XCAP")))) : LQ_XCAP;
    d.est = est.p; d.est0 = est0.p; d.flipw = flipw.p; d.openw = openw.p;
    d.sdim = sdim; d.bond_vec = bond_vec.p; d.wind = wind.p; d.gstride = gstride();
    for (int x = 0; x < 3; ++x) d.wscale[x] = 0.5 * wunit[x];   // stiffness.h:127: (winding / 2)^2
    d.sse = sse ? 1 : 0; d.spos = sse_rank.p; d.nbin = nbin; d.bincnt = sse_bincnt.p; d.binbase = sse_binbase.p;
    d.binfill = sse_binfill.p; d.sorted_time = sse_time.p; d.sorted_id = sse_id.p;
    d.ncap = ncap; d.nccap = nccap;
    d.d_ntotal = d_ntotal.p; d.d_nc = d_nc.p; d.d_err = d_err.p;
    d.dbg = getenv("LQ_DBG") ? atoi(getenv("LQ_DBG")) : 0;
    if (!dbgc.p) { dbgc.alloc(8, nullptr); CK(cudaMemset(dbgc.p, 0, 8 * sizeof(unsigned long long))); }
    d.dbgc = dbgc.p;
  }

  void clear_state() {
    for (int k = 0; k < 2; ++k) {
      CK(cudaMemset(boff[k].p, 0, boff[k].n * sizeof(uint16_t)));
      CK(cudaMemset(pcount[k].p, 0, pcount[k].n * sizeof(int)));
    }
    CK(cudaMemset(nbase.p, 0, nbase.n * sizeof(int)));
    CK(cudaMemset(spinW.p, 0, spinW.n));
    CK(cudaMemset(d_ntotal.p, 0, sizeof(int)));
    cur = 0;
  }

  // -------------------------------------------------------------------------------------------
  // timers (ids of path_integral.C:284-299); device time between two events on the stream
  struct Section {
    lq_engine* e; int id; cudaEvent_t a = nullptr, b = nullptr;
    std::chrono::steady_clock::time_point h0;
    Section(lq_engine* e_, int id_) : e(e_), id(id_) {
      if (e->timers_on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, e->stream); }
      if (e->host_timers) h0 = std::chrono::steady_clock::now();
    }
    ~Section() {
      if (e->timers_on) { cudaEventRecord(b, e->stream); e->tpending.push_back({id, {a, b}}); }
      if (e->host_timers) {   // (steady state: the first 40 calls -- connection set-up, first launches -- are left out)
        if (e->hcnt[id] >= 40) e->hacc[id] += std::chrono::duration<double>(std::chrono::steady_clock::now() - h0).count();
        e->hcnt[id]++;
      }
    }
  };
  // LQ_HOST_TIMERS=1 (diagnostic): host time spent ENQUEUEING every section, printed when the engine is destroyed
  bool host_timers = getenv("LQ_HOST_TIMERS") != nullptr;
  double hacc[17] = {0};
  long hcnt[17] = {0};
  void drain_timers() {
    for (auto& t : tpending) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, t.second.first, t.second.second) == cudaSuccess) {
        tacc[t.first].sec += 1e-3 * ms;
        tacc[t.first].count++;
      }
      cudaEventDestroy(t.second.first);
      cudaEventDestroy(t.second.second);
    }
    tpending.clear();
  }

  // -------------------------------------------------------------------------------------------
  void scan_u32(const uint32_t* in, uint32_t* out, size_t n, uint32_t* total_out, int* total_out2) {
    const size_t nblk = (n + LQ_SCAN_CHUNK - 1) / LQ_SCAN_CHUNK;
    lq::k_scan_blocksum<<<(unsigned)nblk, LQ_SCAN_THREADS, 0, stream>>>(in, n, scan_tmp.p);
    lq::k_scan_top<<<1, LQ_SCAN_THREADS, 0, stream>>>(scan_tmp.p, nblk, total_out, total_out2);
    lq::k_scan_final<<<(unsigned)nblk, LQ_SCAN_THREADS, 0, stream>>>(in, out, n, scan_tmp.p);
    launches += 3;
  }

  static unsigned grid_for(size_t n, int threads) { return (unsigned)((n + threads - 1) / threads); }

  // K2 + K3 on the live buffer (+ K4/K5 sums); used by the step and by lq_build_clusters
  void label_clusters(double* out_slot, const lq::StepParams* sp_in, bool flip) {
    const lq::StepParams* sp = sp_in;   // (`sp` the member is the spatial-cut table)
    const int N = Ns;
    const size_t nodes_cap = (size_t)N + (size_t)npo * (size_t)ncap;
    {
      Section s(this, 6);
      scan_u32((const uint32_t*)pcount[cur].p, (uint32_t*)nbase.p, P, (uint32_t*)(nbase.p + P), d_ntotal.p);
      lq::k_init_nodes<<<grid_for(N, 256), 256, 0, stream>>>(d);
      launches += 1;
      if (sse) {   // string positions of the operators (sse.C: the index t of the sweep over the string)
        const size_t nb = (size_t)Wl * nbin;
        CK(cudaMemsetAsync(sse_bincnt.p, 0, (nb + 1) * sizeof(uint32_t), stream));
        CK(cudaMemsetAsync(sse_binfill.p, 0, (nb + 1) * sizeof(uint32_t), stream));
        lq::k_sse_hist<<<(unsigned)P, 256, 0, stream>>>(d, cur);
        scan_u32(sse_bincnt.p, sse_binbase.p, nb + 1, nullptr, nullptr);
        lq::k_sse_scatter<<<(unsigned)P, 256, 0, stream>>>(d, cur);
        lq::k_sse_rank<<<grid_for((size_t)ncap, 256), 256, 0, stream>>>(d);
        launches += 3;
      }
      if (space) {
        // ghost nodes: junk unless they sit in a boundary segment; the segments' entries in (bond, window, slot) order
        const size_t nghost = (size_t)npo * (size_t)(P - Pown) * (size_t)cap;
        const size_t ncnt = (size_t)spd.nbt * Wl;
        if (nghost) lq::k_sp_init_ghost<<<grid_for(nghost, 256), 256, 0, stream>>>(d);
        // (bucket sizes and offsets of the segments: left by the halo exchange that precedes every labelling)
        CK(cudaMemsetAsync(sp_bnode.p, 0xff, sp_bnode.n * sizeof(uint32_t), stream));
        lq::k_sp_fill<<<grid_for(std::max<size_t>(std::max(ncnt, (size_t)spd.nst), 1), 256), 256, 0, stream>>>(d, spd, cur);
        launches += 2;
      }
    }
    {
      Section s(this, 7);
      walk_fn<<<(unsigned)((size_t)(To + Tw) * Wl), tpb_walk, walk_smem, stream>>>(d, cur);
      lq::k_carry_scan<<<grid_for(Nwalk, 128), 128, 0, stream>>>(d);   // closes the chain of windows per site
      launches += 1;
      const unsigned ngroups = (unsigned)((size_t)To * ((Wl + ug - 1) / ug));
      lq::k_union_local<<<ngroups, 256, (size_t)ug * npo * cap * sizeof(uint32_t), stream>>>(d, cur);
      lq::k_union_global<<<ngroups, 256, 0, stream>>>(d, cur);
      launches += 3;
    }
    {
      Section s(this, 9);   // (slab engines: the top boundary is merged by the exchange instead)
      if (opt.nranks == 1 || space) { lq::k_close<<<grid_for(Nown, 128), 128, 0, stream>>>(d); launches += 1; }
    }
    {
      Section s(this, 11);
      {
        const unsigned g = grid_for(nwords_cap * 32, 256 * LQ_NPT);
        if (d.dbg & 1) {
          if (space) lq::k_compress<true, true><<<g, 256, 0, stream>>>(d, nwords_cap);
          else lq::k_compress<false, true><<<g, 256, 0, stream>>>(d, nwords_cap);
        } else if (space) lq::k_compress<true, false><<<g, 256, 0, stream>>>(d, nwords_cap);
        else lq::k_compress<false, false><<<g, 256, 0, stream>>>(d, nwords_cap);
      }
      scan_u32(wcount.p, wbase.p, nwords_cap, wbase.p + nwords_cap, (int*)d_nc.p);
      if (opt.nranks == 1) {   // flip decision per root (path_integral.C:796-799), packed for k_relabel
        lq::k_rootflip<<<(unsigned)std::min<size_t>((nwords_cap + 255) / 256, (size_t)sm_count * 8), 256, 0, stream>>>(d, sp);
        launches += 1;
      }
      if (opt.nranks == 1) { lq::k_relabel<<<grid_for(nodes_cap, 256 * LQ_NPT), 256, 0, stream>>>(d); launches += 1; }
      launches += 1;
    }
    if (opt.nranks > 1) {
      // multi-rank engines: exchange first (on roots), flips of ALL clusters, then the relabelling packs them
      lq::k_set_ncs<<<1, 1, 0, stream>>>(d);
      merge_open_clusters();
      Section s(this, 14);
      lq::k_flipbits<<<(unsigned)std::min<size_t>(nblk_collect, (size_t)sm_count * 8), 256, 0, stream>>>(d, sp);
      if (space) lq::k_sp_openflips<<<grid_for(std::max<size_t>((size_t)spd.cb, 1), 256), 256, 0, stream>>>(d, spd, mr, sp);
      else lq::k_mr_openflips<<<grid_for(N, 128), 128, 0, stream>>>(d, mr, sp);
      lq::k_relabel<<<grid_for(nodes_cap, 256 * LQ_NPT), 256, 0, stream>>>(d);
      launches += 4;
    }
    {
      Section s(this, 12);
      const size_t est_smem = sizeof(lq::EstHash) + 2 * (size_t)part.nbmax + 16 +
                              (sdim > 0 ? sizeof(lq::WindHash) + 6 * (size_t)part.nbmax : 0);
      typedef void (*est_fn_t)(lq::Dev, int);
      static const est_fn_t est_fns[8] = {
          lq::k_estimate<false, false, false>, lq::k_estimate<false, false, true>, lq::k_estimate<false, true, false>,
          lq::k_estimate<false, true, true>,   lq::k_estimate<true, false, false>, lq::k_estimate<true, false, true>,
          lq::k_estimate<true, true, false>,   lq::k_estimate<true, true, true>};
      est_fns[(flip ? 4 : 0) | (sdim > 0 ? 2 : 0) | (npo == 2 ? 1 : 0)]<<<(unsigned)Pown, 256, est_smem, stream>>>(d, cur);
      launches += 1;
    }
    {
      Section s(this, 10);   // world-line ends at the slab boundaries; open-cluster sums into the exchange table
      lq::k_estimate_sites<<<grid_for(Nown, 128), 128, 0, stream>>>(d);
      launches += 1;
      if (space) {
        const size_t cb = std::max<size_t>((size_t)spd.cb, 1);
        lq::k_sp_gather<<<grid_for(cb, 1024), 1024, 0, stream>>>(d, spd, mr);
        lq::k_sp_reset<<<grid_for(cb, 256), 256, 0, stream>>>(d, spd, mr);
        launches += 2;
      } else if (opt.nranks > 1) {
        lq::k_mr_gather<<<grid_for(N, 1024), 1024, 0, stream>>>(d, mr);
        lq::k_mr_reset_topmin<<<grid_for(N, 128), 128, 0, stream>>>(d, mr);
        launches += 2;
      }
    }
    {
      Section s(this, 13);
      lq::k_collect<<<(unsigned)nblk_collect, 256, 0, stream>>>(d, partial.p);
      lq::k_collect_final<<<1, 256, 0, stream>>>(d, partial.p, nblk_collect, out_slot);
      launches += 2;
    }
    if (opt.nranks > 1) finish_open_clusters(out_slot, sp);
  }

  // ---- multi-rank exchange (looper/parallel.h:1609-1809 restated for one all-gather + one
  // all-reduce over NVLink; timer ids 21.. of parallel.h:1562-1583 are folded into id 13) -------
  void comm_check(int rc, const char* what) {
    if (rc != 0) fail(LQ_E_COMM, std::string(what) + " failed in the communicator callback");
  }

  void all_gather(const void* send, void* recv, size_t bytes, const char* what) {
    if (nccl) nccl_check(nccl_api().AllGather(send, recv, bytes, ncclChar, nccl, stream), what);
    else comm_check(comm.all_gather(comm.ctx, send, recv, (int64_t)bytes, stream), what);
  }
  void all_reduce_i64(void* buf, size_t count, const char* what) {
    if (nccl) nccl_check(nccl_api().AllReduce(buf, buf, count, ncclInt64, ncclSum, nccl, stream), what);
    else comm_check(comm.all_reduce_i64(comm.ctx, buf, (int64_t)count, stream), what);
  }

  void merge_open_clusters() {
    Section s(this, 13);
    if (space) {
      const size_t cb = std::max<size_t>((size_t)spd.cb, 1), g2 = mr.gn;
      lq::k_sp_topmin<<<grid_for(cb, 256), 256, 0, stream>>>(d, spd, mr);
      lq::k_sp_ids<<<grid_for(std::max<size_t>((size_t)spd.cbmax, LQ_SP_HDR), 256), 256, 0, stream>>>(d, spd, mr);
      launches += 2;
      all_gather(mr.sendb, mr.recvb, (size_t)spd.stride * sizeof(uint32_t), "all_gather(boundary entries)");
      CK(cudaMemsetAsync(mr.d_g, 0, 4 * sizeof(uint32_t), stream));
      lq::k_sp_ginit<<<grid_for(g2, 256), 256, 0, stream>>>(d, spd, mr);
      if (spd.ngseg > 0 && sp_maxcap > 0)
        lq::k_sp_gunion<<<dim3(grid_for((size_t)sp_maxcap, 256), (unsigned)spd.ngseg), 256, 0, stream>>>(d, spd, mr);
      lq::k_mr_gcompress<<<grid_for(((g2 + 31) / 32) * 32, 256), 256, 0, stream>>>(d, mr);
      launches += 3;
      scan_u32(mr.gwcount, mr.gwbase, (g2 + 31) / 32, mr.gwbase + (g2 + 31) / 32, (int*)mr.d_g);
      size_open_cluster_table();
      return;
    }
    const int N = part.N;
    const size_t g2 = (size_t)opt.nranks * 2 * N;
    lq::k_mr_topmin<<<grid_for(N, 128), 128, 0, stream>>>(d, mr);
    lq::k_mr_ids<<<grid_for(N, 128), 128, 0, stream>>>(d, mr);
    launches += 2;
    all_gather(mr.sendb, mr.recvb, 2 * (size_t)N * sizeof(uint32_t), "all_gather(boundary ids)");
    CK(cudaMemsetAsync(mr.d_g, 0, 4 * sizeof(uint32_t), stream));
    lq::k_mr_ginit<<<grid_for(g2, 256), 256, 0, stream>>>(d, mr);
    lq::k_mr_gunion<<<grid_for(g2 / 2, 256), 256, 0, stream>>>(d, mr);
    lq::k_mr_gcompress<<<grid_for(((g2 + 31) / 32) * 32, 256), 256, 0, stream>>>(d, mr);
    launches += 3;
    scan_u32(mr.gwcount, mr.gwbase, (g2 + 31) / 32, mr.gwbase + (g2 + 31) / 32, (int*)mr.d_g);
    size_open_cluster_table();
  }

  // The all-reduce of the open-cluster sums needs its length on the HOST.  Reading the number of
  // global open clusters back every step made the host wait for the GPU and then the GPU wait for the
  // host to enqueue the next step.  The count is therefore taken from TWO steps
  // ago (+25 %; the same number on every rank, which the collective needs), whose read-back has long
  // arrived; a step with more open clusters than that raises the error word and is replayed like any
  // other arena overflow.  The first two labellings after (re)sizing wait for the exact count.
  void size_open_cluster_table() {
    const int slot = (int)(mr_step & 3);
    CK(cudaMemcpyAsync(h_mr + 4 * slot, mr.d_g, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CK(cudaEventRecord(mr_ev[slot], stream));
    uint64_t want;
    if (mr_step >= 2 && !getenv("LQ_MR_SYNC")) {
      const int old = (int)((mr_step - 2) & 3);
      CK(cudaEventSynchronize(mr_ev[old]));
      const uint64_t n = h_mr[4 * old];
      want = n + std::max<uint64_t>(n / 16, 4096);   // (measured step-to-step changes: 1e-3 at 1e6 open clusters)
    } else {
      CK(cudaEventSynchronize(mr_ev[slot]));
      want = std::max<uint64_t>(1, h_mr[4 * slot]);
    }
    mr.gcap = (uint32_t)std::min<uint64_t>(want, mr_gslots);
    ++mr_step;
  }

  // ---- spatial cut: halo exchange (SpacePlan) ---------------------------------------------------
  // Every owner packs the buckets of its segments' bonds into a dense stream (k_sp_pack), the streams
  // travel by ncclSend/ncclRecv (one group: both neighbours at once) or lq_comm.send_recv, and the user
  // rebuilds its ghost pages from them (k_sp_unpack).  Twice per step: before K1 -- the flips of the
  // previous step changed operator types and spins -- and after it, for the new operators.
  struct XMsg { const void* s; size_t sb; void* r; size_t rb; };
  void xchg_round(int dl, const std::vector<XMsg>& msgs) {
    const int Pn = opt.nranks, dst = (opt.rank + dl) % Pn, src = (opt.rank - dl + Pn) % Pn;
    if (nccl) {
      NcclApi& a = nccl_api();
      for (const XMsg& m : msgs) {
        if (m.sb) nccl_check(a.Send(m.s, m.sb, ncclChar, dst, nccl, stream), "ncclSend(halo stream)");
        if (m.rb) nccl_check(a.Recv(m.r, m.rb, ncclChar, src, nccl, stream), "ncclRecv(halo stream)");
      }
    } else {
      if (!comm.send_recv) fail(LQ_E_COMM, "the spatial cut needs lq_comm.send_recv (or lq_comm_init)");
      for (const XMsg& m : msgs)
        comm_check(comm.send_recv(comm.ctx, m.s, (int64_t)m.sb, dst, m.r, (int64_t)m.rb, src, stream), "send_recv(halo stream)");
    }
  }
  void exchange_ghosts(int buf, bool spins) {
    Section s(this, 8);
    const size_t ncnt = (size_t)spd.nbt * Wl, nspin = spins ? (size_t)spd.nst * (size_t)(Wl + 1) : 0;
    // owner side: bucket sizes -> dense offsets -> stream
    if (ncnt) lq::k_sp_count<<<grid_for(ncnt, 256), 256, 0, stream>>>(d, spd, buf);
    scan_u32(sp_cnt.p, sp_base.p, ncnt + 1, nullptr, nullptr);
    lq::k_sp_pack<<<grid_for(std::max<size_t>(std::max(ncnt, nspin), 1), 256), 256, 0, stream>>>(d, spd, buf, spins ? 1 : 0);
    launches += 5;
    std::vector<XMsg> msgs;
    if (nccl) nccl_check(nccl_api().GroupStart(), "ncclGroupStart");
    for (int dl = 1; dl < opt.nranks; ++dl) {
      if (!plan.rounds[dl]) continue;
      const int dst = (opt.rank + dl) % opt.nranks, src = (opt.rank - dl + opt.nranks) % opt.nranks;
      const lq::SpSeg* a = nullptr; const lq::SpSeg* b = nullptr;   // my segment for dst / src's segment for me
      for (size_t i = 0; i < sp_segs_h.size(); ++i) {
        const SpaceSeg& g = plan.segs[sp_seg_index[i]];
        if (g.owner == opt.rank && g.user == dst) a = &sp_segs_h[i];
        if (g.user == opt.rank && g.owner == src) b = &sp_segs_h[i];
      }
      msgs.clear();
      auto add = [&](auto* base, auto off_of, auto len_of) {
        typedef std::remove_pointer_t<decltype(base)> T;
        msgs.push_back({a ? (const void*)(base + off_of(*a)) : nullptr, a ? len_of(*a) * sizeof(T) : 0,
                        b ? (void*)(base + off_of(*b)) : nullptr, b ? len_of(*b) * sizeof(T) : 0});
      };
      add(sp_cnt.p, [&](const lq::SpSeg& x) { return (size_t)x.bond0 * Wl; }, [&](const lq::SpSeg& x) { return (size_t)x.nb * Wl; });
      add(sp_pk_time.p, [&](const lq::SpSeg& x) { return (size_t)x.pk0; }, [&](const lq::SpSeg& x) { return (size_t)x.opcap; });
      add(sp_pk_info.p, [&](const lq::SpSeg& x) { return (size_t)x.pk0; }, [&](const lq::SpSeg& x) { return (size_t)x.opcap; });
      if (spins)
        add(sp_pk_spin.p, [&](const lq::SpSeg& x) { return (size_t)x.sp0; }, [&](const lq::SpSeg& x) { return (size_t)x.ns * (size_t)(Wl + 1); });
      xchg_round(dl, msgs);
    }
    if (nccl) nccl_check(nccl_api().GroupEnd(), "ncclGroupEnd");
    // user side: the received bucket sizes complete the count table -> offsets -> ghost pages
    scan_u32(sp_cnt.p, sp_base.p, ncnt + 1, nullptr, nullptr);
    const size_t nghost = (size_t)(Tl - To) * Wl;
    if (nghost) lq::k_sp_unpack<<<(unsigned)nghost, 256, 0, stream>>>(d, spd, buf, spins ? 1 : 0);
    launches += 4;
  }

  void finish_open_clusters(double* out_slot, const lq::StepParams* sp) {
    Section s(this, 13);
    // one all-reduce carries, in front, every rank's collector of its closed clusters and then the open-cluster sums
    double* tail = (double*)mr_gest.p;
    lq::k_mr_rankvec<<<1, 32, 0, stream>>>(d, mr, out_slot, tail);
    all_reduce_i64(mr_gest.p, 32 * (size_t)opt.nranks + (size_t)mr.gcap * gstride(), "all_reduce(collectors + open-cluster sums)");
    const unsigned gblk = (unsigned)std::min<size_t>(nblk_collect, (size_t)sm_count * 4);
    lq::k_mr_gcollect<<<gblk, 256, 0, stream>>>(d, mr, partial.p);   // (partial is free again after k_collect_final)
    lq::k_mr_gsum<<<1, 32 * LQ_NSUM, 0, stream>>>(mr, partial.p, (int)gblk);
    lq::k_mr_final<<<1, 32, 0, stream>>>(d, mr, tail, out_slot);
    launches += 4;
  }

  void ensure_out(size_t slots) {
    if (params_n < slots) {
      if (h_params) cudaFreeHost(h_params);
      CK(cudaMallocHost((void**)&h_params, slots * sizeof(lq::StepParams)));
      d_params.alloc(slots, &device_bytes);
      params_n = slots;
    }
    if (d_out.n < slots * 32) d_out.alloc(slots * 32, &device_bytes);
    if (h_out_n < slots * 32) {
      if (h_out) cudaFreeHost(h_out);
      CK(cudaMallocHost((void**)&h_out, slots * 32 * sizeof(double)));
      h_out_n = slots * 32;
    }
  }

  // stage the inputs of `count` steps in pinned memory and copy them to the device
  void stage_params(int count, bool advance) {
    for (int i = 0; i < count; ++i) {
      h_params[i].beta = beta;
      h_params[i].key0 = (uint32_t)opt.seed;
      h_params[i].key1 = (uint32_t)(opt.seed >> 32);
      h_params[i].mcs = advance ? mcs + (uint32_t)i : 0xffffffffu;
      h_params[i].pad = 0;
    }
    CK(cudaMemcpyAsync(d_params.p, h_params, (size_t)count * sizeof(lq::StepParams),
                       cudaMemcpyHostToDevice, stream));
    h2d_bytes += (int64_t)count * (int64_t)sizeof(lq::StepParams);
  }

  void enqueue_step(double* out_slot, const lq::StepParams* sp) {
    if (space) exchange_ghosts(cur, true);   // the flips of the previous step: types of the ghost operators, ghost spins
    {
      Section s(this, 5);
      const unsigned nch = (unsigned)((Wl + k1_chunk - 1) / k1_chunk);
      k1_fn<<<(unsigned)To * nch, k1_nt, stage_smem, stream>>>(d, cur, sp, k1_chunk, k1_lay);
      launches += 1;
      cur ^= 1;
    }
    if (space) exchange_ghosts(cur, false);   // the new operators on the halo bonds
    label_clusters(out_slot, sp, true);   // includes the flip of the operators (fused into K4)
    {
      Section s(this, 15);
      lq::k_flip_spins<<<grid_for((size_t)(Wl + 1) * Ns, 256), 256, 0, stream>>>(d);
      launches += 1;
    }
    ++mcs;
  }

  // ---- CUDA graphs for small systems ---------------------------------------------------------
  // One executable graph per page-buffer parity holds a whole Monte Carlo step; a batch of steps is
  // `count` graph launches.  Only serial engines with timers off, and only below LQ_GRAPH_PAGES
  // pages: above that the launches are hidden behind kernels that run for milliseconds.
  static constexpr size_t LQ_GRAPH_PAGES = 65536;
  cudaGraphExec_t gexec[2] = {nullptr, nullptr};
  int64_t glaunches[2] = {0, 0};
  lq::StepCtl* h_ctl = nullptr;   // pinned
  DBuf<lq::StepCtl> d_ctl;
  DBuf<double> d_gout;            // the fixed collector slot of the captured kernels
  bool graphs_off = false;

  void drop_graphs() {
    for (int k = 0; k < 2; ++k)
      if (gexec[k]) { cudaGraphExecDestroy(gexec[k]); gexec[k] = nullptr; }
  }
  bool use_graphs() const {
    return opt.nranks == 1 && !timers_on && !graphs_off && P <= LQ_GRAPH_PAGES;
  }
  void capture_graph(int parity) {
    if (!d_ctl.p) {
      d_ctl.alloc(1, &device_bytes);
      d_gout.alloc(32, &device_bytes);
      CK(cudaMallocHost((void**)&h_ctl, sizeof(lq::StepCtl)));
    }
    const int cur_save = cur;
    const uint32_t mcs_save = mcs;
    const int64_t l0 = launches;
    cudaGraph_t g = nullptr;
    cur = parity;
    CK(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
    enqueue_step(d_gout.p, &d_ctl.p->sp);
    lq::k_step_advance<<<1, 32, 0, stream>>>(d_ctl.p, d_gout.p);
    launches += 1;
    const cudaError_t ce = cudaStreamEndCapture(stream, &g);
    cur = cur_save;
    mcs = mcs_save;
    glaunches[parity] = launches - l0;
    launches = l0;
    if (ce != cudaSuccess) throw cuda_error{ce, "cudaStreamEndCapture", __FILE__, __LINE__};
    const cudaError_t ie = cudaGraphInstantiate(&gexec[parity], g, 0);
    cudaGraphDestroy(g);
    if (ie != cudaSuccess) throw cuda_error{ie, "cudaGraphInstantiate", __FILE__, __LINE__};
  }
  void enqueue_batch_graph(int count) {
    h_ctl->sp.beta = beta;
    h_ctl->sp.key0 = (uint32_t)opt.seed;
    h_ctl->sp.key1 = (uint32_t)(opt.seed >> 32);
    h_ctl->sp.mcs = mcs;
    h_ctl->sp.pad = 0;
    h_ctl->slot = 0;
    h_ctl->pad = 0;
    h_ctl->ring = d_out.p;
    CK(cudaMemcpyAsync(d_ctl.p, h_ctl, sizeof(lq::StepCtl), cudaMemcpyHostToDevice, stream));
    h2d_bytes += (int64_t)sizeof(lq::StepCtl);
    for (int i = 0; i < count; ++i) {
      CK(cudaGraphLaunch(gexec[cur], stream));
      launches += glaunches[cur];
      cur ^= 1;
      ++mcs;
    }
  }

  void to_collector(const double* o, lq_collector* c) const {
    c->umag0 = o[0]; c->usize2 = o[1]; c->umag2 = o[2]; c->usize4 = o[3]; c->umag4 = o[4];
    c->usize = o[5]; c->umag = o[6];
    c->smag0 = o[7]; c->ssize2 = o[8]; c->smag2 = o[9]; c->ssize4 = o[10]; c->smag4 = o[11];
    c->ssize = o[12]; c->smag = o[13];
    c->nc = o[14]; c->nop = o[15]; c->noc = o[17]; c->tlen = o[18]; c->w2 = o[19];
    c->ene = energy_offset - c->nop / beta;  // path_integral.C:851
  }

  void check_err(int err) {
    if (!err) return;
    CK(cudaMemsetAsync(d_err.p, 0, sizeof(int), stream));
    std::string m = "arena overflow:";
    if (err & LQ_ERR_PAGE_FULL) m += " page full (raise lq_options.reserve);";
    if (err & LQ_ERR_CAND_FULL) m += " too many candidates in one page or bucket (lower window_ops);";
    if (err & LQ_ERR_CLUSTER_FULL) m += " cluster arena full (raise cluster_reserve);";
    if (err & LQ_ERR_NODE_FULL) m += " node arena / boundary segment full;";
    if (err & LQ_ERR_OPEN_FULL) m += " open-cluster table of the exchange full;";
    if (err & LQ_ERR_BOUNDARY) m += " the two copies of a boundary page disagree (internal error);";
    fail(LQ_E_OVERFLOW, m);
  }

  // A step that overflows an arena leaves the configuration it started from intact: K1 writes the
  // other page buffer, the labelling works on scratch, and k_flip_spins / every later K1 do nothing
  // once the error word is set.  The host then rewinds to that step, enlarges the arena that
  // overflowed (operators travel through lq_get_state / lq_set_state) and runs the remaining steps
  // again.  Random numbers are keyed by (bond, window, step, draw) and by cluster id, not by slots,
  // so the trajectory does not depend on the capacities.
  void sweep_many(int count, lq_collector* out, int depth = 0) {
    if (count <= 0) return;
    if (opt.nranks > 1 && !has_comm) fail(LQ_E_COMM, "nranks > 1 but neither lq_comm_init nor lq_set_comm was called");
    ensure_out((size_t)count);
    const int cur0 = cur;
    const uint32_t mcs0 = mcs;
    if (use_graphs()) {
      for (int k = 0; k < 2; ++k)
        if (!gexec[k]) capture_graph(k);
      enqueue_batch_graph(count);
    } else {
      stage_params(count, true);
      for (int i = 0; i < count; ++i) enqueue_step(d_out.p + (size_t)i * 32, d_params.p + i);
    }
    CK(cudaMemcpyAsync(h_out, d_out.p, (size_t)count * 32 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    d2h_bytes += (int64_t)count * 32 * (int64_t)sizeof(double);
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    drain_timers();
    int err = 0, first_bad = count;
    for (int i = 0; i < count; ++i) {
      const int e = (int)h_out[(size_t)i * 32 + 16];
      if (e && first_bad == count) first_bad = i;
      err |= e;
      if (out && first_bad == count) to_collector(h_out + (size_t)i * 32, out + i);
    }
    if (!err) return;
    // (slab engines: every rank sees the OR of all ranks' error bits in the collector and has kept the
    // configuration of the failing step -- LQ_ERR_REMOTE, k_mr_ids -- so all ranks rewind together)
    // (a boundary mismatch next to an overflow is a consequence of the lost step, alone it is a bug)
    if (depth >= 8 || ((err & LQ_ERR_BOUNDARY) && !(err & (LQ_ERR_PAGE_FULL | LQ_ERR_CAND_FULL | LQ_ERR_NODE_FULL | LQ_ERR_CLUSTER_FULL))))
      check_err(err);
    CK(cudaMemsetAsync(d_err.p, 0, sizeof(int), stream));
    cur = cur0 ^ (first_bad & 1);
    mcs = mcs0 + (uint32_t)first_bad;
    if (!(err & ~(LQ_ERR_OPEN_FULL | LQ_ERR_REMOTE))) {
      // only the open-cluster table of the all-reduce was short (size_open_cluster_table): nothing to
      // grow, the configuration the step started from is intact -- run it again with the exact count
      mr_step = 0;
    } else {
      if (err & (LQ_ERR_PAGE_FULL | LQ_ERR_NODE_FULL)) grow_pages *= 1.5;   // (boundary segments are sized like the pages)
      if (err & LQ_ERR_CAND_FULL) { grow_cand *= 1.5; grow_kept *= 1.6; }
      if (err & LQ_ERR_CLUSTER_FULL) grow_clusters *= 1.5;
      ++regrows;
      rebucket();
    }
    sweep_many(count - first_bad, out ? out + first_bad : nullptr, depth + 1);
  }

  // operators and spins of this engine (of its slab, on a slab engine) through new arenas: after a
  // change of beta or of an arena size
  // Everything size_arenas() allocates, freed up front so that a re-size peaks at (old live pages + new
  // arenas) instead of (all old + all new).
  void release_arenas() {
    for (int k = 0; k < 2; ++k) { time_[k].release(); info[k].release(); boff[k].release(); pcount[k].release(); }
    nbase.release(); spinW.release(); curW.release(); firstW.release(); parent.release(); low0.release();
    bitmap.release(); wcount.release(); wbase.release(); rootw.release(); xedge.release(); xcount.release();
    sse_rank.release(); sse_time.release(); sse_id.release(); sse_bincnt.release(); sse_binbase.release(); sse_binfill.release();
    scan_tmp.release(); est.release(); est0.release(); flipw.release(); wind.release(); openw.release(); partial.release();
    mr_topmin.release(); mr_sendb.release(); mr_recvb.release(); mr_gparent.release(); mr_gused.release();
    mr_gbitmap.release(); mr_gwcount.release(); mr_gwbase.release(); mr_gest.release();
    sp_cnt.release(); sp_base.release(); sp_bnode.release(); sp_pk_time.release(); sp_pk_info.release(); sp_pk_spin.release();
  }

  // operators and spins of this engine (of its slab / its tiles on a multi-rank engine) into new arenas:
  // after a change of beta or of an arena size.  On the device (lq_rebucket.cuh): every (tile, bond)
  // column is already time-ordered and only has to be re-cut at the new window boundaries.
  void rebucket() {
    if (getenv("LQ_REBUCKET_HOST")) { rebucket_host(); return; }
    // (spatial cut: all ranks come here together -- lq_set_beta is collective, and an overflow makes every
    // rank rewind.  The ghost pages still carry the operator types of before the last flip: refresh them,
    // the spins at the new window starts are recomputed from the off-diagonal legs of owned AND ghost operators.)
    if (space) {
      if (!has_comm) fail(LQ_E_COMM, "nranks > 1 but neither lq_comm_init nor lq_set_comm was called");
      exchange_ghosts(cur, true);
    }
    CK(cudaStreamSynchronize(stream));
    DBuf<double> otime;
    DBuf<uint32_t> oinfo;
    DBuf<uint16_t> oboff;
    DBuf<uint8_t> ospin;
    otime.take(time_[cur]); oinfo.take(info[cur]); oboff.take(boff[cur]);
    ospin.alloc((size_t)Ns, nullptr);
    CK(cudaMemcpy(ospin.p, spinW.p, (size_t)Ns, cudaMemcpyDeviceToDevice));
    const lq::OldPages o{otime.p, oinfo.p, oboff.p, W, Wl, w0, cap};
    release_arenas();
    const int nbl = part.bond_base[Tl];
    for (int attempt = 0;; ++attempt) {
      size_arenas();
      clear_state();
      CK(cudaDeviceSynchronize());   // (the memsets above ran on the default stream)
      CK(cudaMemsetAsync(d_err.p, 0, sizeof(int), stream));
      if (nbl) lq::k_rb_columns<0><<<grid_for(nbl, 128), 128, 0, stream>>>(d, o, nbl);
      lq::k_rb_offsets<<<(unsigned)P, 128, 0, stream>>>(d);
      int err = 0;
      CK(cudaMemcpyAsync(&err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      launches += 2;
      if (!err) break;
      CK(cudaMemsetAsync(d_err.p, 0, sizeof(int), stream));
      if ((err & LQ_ERR_BOUNDARY) || space || attempt >= 6)   // (spatial cut: the page capacity is part of the exchange format)
        fail(LQ_E_OVERFLOW, "the configuration does not fit the new pages (raise lq_options.reserve)");
      grow_pages *= 1.5;   // a page of the new layout is fuller than its capacity allows: larger pages, again
      release_arenas();
    }
    if (nbl) lq::k_rb_columns<1><<<grid_for(nbl, 128), 128, 0, stream>>>(d, o, nbl);
    lq::k_rb_spin_parity<<<grid_for((size_t)Wl * Ns, 256), 256, 0, stream>>>(d, Tl);
    lq::k_rb_spin_scan<<<grid_for(Ns, 128), 128, 0, stream>>>(d, ospin.p);
    launches += 3;
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
  }

  // the same through the host (lq_get_state -> lq_set_state); kept for LQ_REBUCKET_HOST=1 as a cross-check
  void rebucket_host() {
    if (space) {
      if (!has_comm) fail(LQ_E_COMM, "nranks > 1 but neither lq_comm_init nor lq_set_comm was called");
      exchange_ghosts(cur, true);
    }
    int64_t n = 0;
    get_state(nullptr, nullptr, &n, true);
    std::vector<int32_t> spins(part.N);
    std::vector<lq_op> ops((size_t)n);
    get_state(spins.data(), ops.data(), &n, true);
    size_arenas();
    clear_state();
    set_state(spins.data(), ops.data(), n, opt.nranks > 1);
  }

  // -------------------------------------------------------------------------------------------
  // state import / export (host side, test and checkpoint path)
  // -------------------------------------------------------------------------------------------
  // local = true (slab engines re-bucketing their own slab): `spins` is the state at the start of this
  // rank's slab and `ops` holds the operators of the slab only
  void set_state(const int32_t* spins, const lq_op* ops, int64_t n, bool local = false) {
    const int N = part.N;
    // two passes over the (time-sorted) operators: bucket sizes -> offsets, then placement with the
    // offsets as cursors -- flat arrays only (a vector per bucket would be 3.6e8 vectors at 1024^2, beta 1024)
    const size_t nb1 = (size_t)part.nbmax + 1;
    std::vector<uint16_t> hb(P * nb1, 0);
    std::vector<int> hc(P, 0);
    auto locate = [&](const lq_op& o, size_t* pg, int* lbo, uint32_t* info) -> bool {
      const bool is_site = !(o.loc & 1);
      const int pos = o.loc >> 1;
      const int bi = part.bond_e2i[is_site ? Breal + pos : pos];
      const int w = window_of(o.time, W);
      if (w < w0 || w >= w0 + Wl) return false;
      const int tl = part.bond_tile[bi];
      if (tl >= Tl) return false;   // spatial cut: a tile this rank neither owns nor mirrors
      *lbo = bi - part.bond_base[tl];
      *pg = (size_t)tl * Wl + (w - w0);
      *info = ((uint32_t)*lbo << LQ_INFO_LBSHIFT) | ((uint32_t)((o.type >> 2) & 3) << LQ_INFO_GSHIFT) |
              (uint32_t)(o.type & 1) | (is_site ? LQ_INFO_SITE : 0u);
      return true;
    };
    // spin at the start of every global window
    std::vector<uint8_t> par((size_t)(W + 1) * N, 0);
    double tprev = -1;
    for (int64_t k = 0; k < n; ++k) {
      const bool is_site = !(ops[k].loc & 1);   // location_impl.h:37: loc = pos << 1 | is_bond
      if (is_site && !has_site) fail(LQ_E_INVALID, "site operator on a model without site graph weights");
      const int pos = ops[k].loc >> 1;
      if (pos < 0 || pos >= (is_site ? N : Breal)) fail(LQ_E_INVALID, "operator position out of range");
      const int be = is_site ? Breal + pos : pos;
      const double t = ops[k].time;
      if (!(t >= 0 && t < 1)) fail(LQ_E_INVALID, "operator time outside [0,1)");
      if (t < tprev) fail(LQ_E_INVALID, "operators must be sorted by time");
      tprev = t;
      const int g = (ops[k].type >> 2) & 3;
      if (g == 1 && npo != 2) fail(LQ_E_INVALID, "cross graph on a model without v[1] weight");
      if (is_site && g != 0) fail(LQ_E_INVALID, "site operators carry graph 0 (graph_impl.h:68)");
      const int bi = part.bond_e2i[be];
      const int w = window_of(t, W);
      if (ops[k].type & 1) {
        par[(size_t)(w + 1) * N + part.bond_s0[bi]] ^= 1;
        if (!is_site) par[(size_t)(w + 1) * N + part.bond_s1[bi]] ^= 1;
      }
      size_t pg; int lb; uint32_t inf;
      if (!locate(ops[k], &pg, &lb, &inf)) continue;
      if (++hc[pg] > cap) fail(LQ_E_OVERFLOW, "page full while loading state (raise lq_options.reserve)");
      hb[pg * nb1 + lb + 1]++;
    }
    for (size_t p = 0; p < P; ++p)   // sizes (stored one slot up) -> offsets
      for (int lb = 0; lb < part.nbmax; ++lb) hb[p * nb1 + lb + 1] = (uint16_t)(hb[p * nb1 + lb + 1] + hb[p * nb1 + lb]);
    std::vector<double> ht((size_t)ncap, 0.0);
    std::vector<uint32_t> hi((size_t)ncap, 0u);
    for (int64_t k = 0; k < n; ++k) {   // placement: hb[lb] is the cursor of bucket lb (ops arrive time-sorted)
      size_t pg; int lb; uint32_t inf;
      if (!locate(ops[k], &pg, &lb, &inf)) continue;
      const size_t at = pg * (size_t)cap + hb[pg * nb1 + lb]++;
      ht[at] = ops[k].time;
      hi[at] = inf;
    }
    for (size_t p = 0; p < P; ++p) {    // every cursor now holds the END of its bucket: shift back to the starts
      for (int lb = part.nbmax; lb > 0; --lb) hb[p * nb1 + lb] = hb[p * nb1 + lb - 1];
      hb[p * nb1] = 0;
    }
    // (spatial cut: the device rows hold the Ns local sites, which come first in the numbering; the
    // spins of the sites beyond the walked ones are never read)
    std::vector<uint8_t> sw((size_t)(Wl + 1) * Ns);
    {
      std::vector<uint8_t> c(N);
      for (int i = 0; i < N; ++i) c[i] = (uint8_t)(spins[part.site_i2e[i]] & 1);
      for (int w = 0; w <= W; ++w) {
        for (int i = 0; i < N; ++i) c[i] ^= par[(size_t)w * N + i];
        if (w >= w0 && w <= w0 + Wl)
          std::memcpy(&sw[(size_t)(w - w0) * Ns], c.data(), Ns);
      }
      for (int i = 0; i < (space ? Nown : N) && !local; ++i)
        if (c[i] != (uint8_t)(spins[part.site_i2e[i]] & 1))
          fail(LQ_E_INVALID, "operator string is not periodic in imaginary time");
    }
    CK(cudaStreamSynchronize(stream));
    cur = 0;
    mr_step = 0;   // (the open-cluster count of the old configuration says nothing about the new one)
    CK(cudaMemcpy(time_[0].p, ht.data(), ht.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(info[0].p, hi.data(), hi.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(boff[0].p, hb.data(), hb.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(pcount[0].p, hc.data(), hc.size() * sizeof(int), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(spinW.p, sw.data(), sw.size(), cudaMemcpyHostToDevice));
  }

  struct HostOp { double time; int bi; uint32_t info; int idx; };

  // all local operators sorted by (time, internal bond); idx = dense device index
  // (spatial cut: the owned pages only, unless `with_ghosts`; idx counts the pages that are listed)
  void fetch_ops(std::vector<HostOp>& out, bool with_ghosts = false) {
    CK(cudaStreamSynchronize(stream));
    const size_t P = with_ghosts ? this->P : Pown;
    std::vector<int> hc(P);
    CK(cudaMemcpy(hc.data(), pcount[cur].p, P * sizeof(int), cudaMemcpyDeviceToHost));
    out.clear();
    {
      size_t total = 0;
      for (size_t p = 0; p < P; ++p) total += (size_t)hc[p];
      out.reserve(total);
    }
    // contiguous runs of pages per copy (two copies per 64 MiB of operators instead of two per page)
    const size_t chunk = std::max<size_t>(1, ((size_t)64 << 20) / ((size_t)cap * 12));
    std::vector<double> ht(chunk * (size_t)cap);
    std::vector<uint32_t> hi(chunk * (size_t)cap);
    int idx = 0;
    for (size_t p0 = 0; p0 < P; p0 += chunk) {
      const size_t p1 = std::min(P, p0 + chunk);
      CK(cudaMemcpy(ht.data(), time_[cur].p + p0 * (size_t)cap, (p1 - p0) * (size_t)cap * sizeof(double), cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(hi.data(), info[cur].p + p0 * (size_t)cap, (p1 - p0) * (size_t)cap * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      for (size_t p = p0; p < p1; ++p) {
        const int n = hc[p];
        const int tl = (int)(p / Wl);
        const size_t o = (p - p0) * (size_t)cap;
        for (int j = 0; j < n; ++j)
          out.push_back({ht[o + j], part.bond_base[tl] + (int)(hi[o + j] >> LQ_INFO_LBSHIFT), hi[o + j], idx + j});
        idx += n;
      }
    }
    // (equal times: by bond like the walk -- bond_order_key -- and, on one bond, in bucket order)
    std::sort(out.begin(), out.end(), [](const HostOp& a, const HostOp& b) {
      return a.time < b.time || (a.time == b.time && (a.bi < b.bi || (a.bi == b.bi && a.idx < b.idx)));
    });
  }

  // (spatial cut: the operators this rank owns and the spins of its own sites, -1 elsewhere; with
  // `with_ghosts` also the ghost copies and the spins of the walked ghost sites -- rebucket())
  void get_state(int32_t* spins, lq_op* ops, int64_t* n, bool with_ghosts = false) {
    std::vector<HostOp> v;
    fetch_ops(v, with_ghosts);
    if (n) *n = (int64_t)v.size();
    if (ops)
      for (size_t k = 0; k < v.size(); ++k) {
        ops[k].time = v[k].time;
        const int be = part.bond_i2e[v[k].bi];
        ops[k].loc = be < Breal ? ((be << 1) | 1) : ((be - Breal) << 1);
        ops[k].type = (int32_t)(v[k].info & 0xf) & ~2;  // offdiag bit + graph bits
      }
    if (spins) {
      std::vector<uint8_t> sw(Ns);
      CK(cudaMemcpy(sw.data(), spinW.p, Ns, cudaMemcpyDeviceToHost));
      const int nvalid = with_ghosts ? Nwalk : Nown;
      for (int i = 0; i < part.N; ++i) spins[part.site_i2e[i]] = i < nvalid ? (int32_t)sw[i] : -1;
    }
  }

  void build_clusters(int32_t* labels_out, int64_t* nc_out, lq_collector* coll_out) {
    if (opt.nranks > 1 && labels_out) fail(LQ_E_UNSUPPORTED, "labels are only exported by a serial engine");
    if (opt.nranks > 1 && !has_comm) fail(LQ_E_COMM, "nranks > 1 but neither lq_comm_init nor lq_set_comm was called");
    ensure_out(1);
    stage_params(1, false);
    if (space) exchange_ghosts(cur, true);
    label_clusters(d_out.p, d_params.p, false);
    labels.alloc(2 * (size_t)std::max<long long>(ncap, 1), nullptr);
    lq::k_export_labels<<<(unsigned)P, 256, 0, stream>>>(d, cur, labels.p);
    launches += 1;
    CK(cudaMemcpyAsync(h_out, d_out.p, 32 * sizeof(double), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    CK(cudaGetLastError());
    drain_timers();
    check_err((int)h_out[16]);
    if (coll_out) to_collector(h_out, coll_out);
    if (nc_out) *nc_out = (int64_t)h_out[14];
    if (labels_out) {
      std::vector<HostOp> v;
      fetch_ops(v);
      const int N = part.N;
      const size_t n = v.size();
      std::vector<uint32_t> sl(N), ol(2 * std::max<size_t>(n, 1));
      CK(cudaMemcpy(sl.data(), parent.p, N * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      if (n) CK(cudaMemcpy(ol.data(), labels.p, 2 * n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
      // canonical min-index labels over: site s (external) -> s ; legs of operator k -> N+2k, N+2k+1
      const size_t ncl = (size_t)h_out[14];
      std::vector<int32_t> minidx(ncl + 1, -1);
      auto touch = [&](uint32_t cid, int32_t idx) {
        if (cid >= minidx.size()) fail(LQ_E_INVALID, "cluster id out of range (internal error)");
        if (minidx[cid] < 0) minidx[cid] = idx;
      };
      for (auto& v : sl) v &= 0x7fffffffu;   // bit 31 of a label is the flip decision
      for (int se = 0; se < N; ++se) touch(sl[part.site_e2i[se]], se);
      for (size_t k = 0; k < n; ++k) {
        touch(ol[2 * (size_t)v[k].idx], (int32_t)(N + 2 * k));
        touch(ol[2 * (size_t)v[k].idx + 1], (int32_t)(N + 2 * k + 1));
      }
      for (int se = 0; se < N; ++se) labels_out[se] = minidx[sl[part.site_e2i[se]]];
      for (size_t k = 0; k < n; ++k) {
        labels_out[N + 2 * k] = minidx[ol[2 * (size_t)v[k].idx]];
        labels_out[N + 2 * k + 1] = minidx[ol[2 * (size_t)v[k].idx + 1]];
      }
    }
    labels.release();
  }
};

// ---------------------------------------------------------------------------------------------
// extern "C"
// ---------------------------------------------------------------------------------------------
#define LQ_TRY(body)                                                                      \
  try { body; return LQ_OK; }                                                             \
  catch (const lq_error& e) { g_err = e.msg; return e.code; }                             \
  catch (const cuda_error& e) {                                                           \
    g_err = std::string(cudaGetErrorString(e.e)) + " in " + e.what + " at " + e.file + ":" + \
            std::to_string(e.line);                                                       \
    return LQ_E_CUDA;                                                                     \
  }                                                                                       \
  catch (const std::bad_alloc&) { g_err = "host out of memory"; return LQ_E_NOMEM; }      \
  catch (const std::exception& e) { g_err = e.what(); return LQ_E_INVALID; }

extern "C" {

int lq_create(lq_handle* out, const lq_lattice* lat, const lq_model* model, double beta,
              const lq_options* opt) {
  if (!out || !lat || !model) { g_err = "null argument"; return LQ_E_INVALID; }
  *out = nullptr;
  lq_options o{};
  if (opt) o = *opt;
  lq_engine* e = nullptr;
  try {
    e = new lq_engine();
    e->setup(*lat, *model, beta, o);
    *out = e;
    return LQ_OK;
  } catch (const lq_error& x) { g_err = x.msg; delete e; return x.code; }
  catch (const cuda_error& x) {
    g_err = std::string(cudaGetErrorString(x.e)) + " in " + x.what + " at " + x.file + ":" + std::to_string(x.line);
    delete e;
    return (x.e == cudaErrorMemoryAllocation) ? LQ_E_NOMEM : LQ_E_CUDA;
  } catch (const std::exception& x) { g_err = x.what(); delete e; return LQ_E_INVALID; }
}

int lq_destroy(lq_handle h) {
  if (!h) return LQ_OK;
  cudaSetDevice(h->opt.device);
  delete h;
  return LQ_OK;
}

int lq_set_beta(lq_handle h, double beta) {
  if (!h) { g_err = "null handle"; return LQ_E_INVALID; }
  LQ_TRY({
    if (!(beta > 0)) fail(LQ_E_INVALID, "beta must be positive");
    CK(cudaSetDevice(h->opt.device));
    // (slab engines: the slab boundaries r / nranks do not move with beta, so every rank re-buckets
    // its own operators; path_integral_mpi.C anneals the same way, temperature.h:39-80)
    h->beta = beta;
    h->rebucket();
  })
}

int lq_set_state(lq_handle h, const int32_t* spins, const lq_op* ops, int64_t n) {
  if (!h || !spins || (n > 0 && !ops) || n < 0) { g_err = "bad argument"; return LQ_E_INVALID; }
  LQ_TRY({ CK(cudaSetDevice(h->opt.device)); h->set_state(spins, ops, n); })
}

int lq_get_state(lq_handle h, int32_t* spins, lq_op* ops, int64_t* n) {
  if (!h) { g_err = "null handle"; return LQ_E_INVALID; }
  LQ_TRY({ CK(cudaSetDevice(h->opt.device)); h->get_state(spins, ops, n); })
}

uint32_t lq_get_step(lq_handle h) { return h ? h->mcs : 0; }
int lq_set_step(lq_handle h, uint32_t step) {
  if (!h) { g_err = "null handle"; return LQ_E_INVALID; }
  h->mcs = step;
  return LQ_OK;
}

int lq_sweep(lq_handle h, lq_collector* out) {
  if (!h) { g_err = "null handle"; return LQ_E_INVALID; }
  LQ_TRY({ CK(cudaSetDevice(h->opt.device)); h->sweep_many(1, out); })
}

int lq_sweep_many(lq_handle h, int32_t count, lq_collector* out) {
  if (!h || count < 0) { g_err = "bad argument"; return LQ_E_INVALID; }
  LQ_TRY({
    CK(cudaSetDevice(h->opt.device));
    int done = 0;
    while (done < count) {  // bounded batches keep the pinned result buffer small
      const int m = std::min(count - done, 4096);
      h->sweep_many(m, out ? out + done : nullptr);
      done += m;
    }
  })
}

int lq_build_clusters(lq_handle h, int32_t* labels_out, int64_t* nc_out, lq_collector* coll_out) {
  if (!h) { g_err = "null handle"; return LQ_E_INVALID; }
  LQ_TRY({ CK(cudaSetDevice(h->opt.device)); h->build_clusters(labels_out, nc_out, coll_out); })
}

int lq_timers(lq_handle h, lq_timer* out, int32_t* count) {
  if (!h || !count) { g_err = "bad argument"; return LQ_E_INVALID; }
  int n = 0;
  for (int id = 3; id <= 16; ++id) {
    if (h->tacc[id].count == 0) continue;
    if (out && n < *count) {
      out[n].id = id;
      out[n].count = h->tacc[id].count;
      out[n].seconds = h->tacc[id].sec;
      std::snprintf(out[n].label, sizeof out[n].label, "%s", kTimerLabels[id]);
    }
    ++n;
  }
  *count = n;
  return LQ_OK;
}

static void tiling_info_impl(const lq_lattice* lat, int32_t tile_sites, int32_t with_sites, lq_tiling* out) {
  for (int b = 0; b < lat->num_bonds; ++b)
    if (lat->src[b] < 0 || lat->src[b] >= lat->num_sites || lat->dst[b] < 0 || lat->dst[b] >= lat->num_sites ||
        lat->src[b] == lat->dst[b])
      fail(LQ_E_INVALID, "bond endpoint out of range");
  std::vector<int> xsrc(lat->src, lat->src + lat->num_bonds);
  std::vector<int> xdst(lat->dst, lat->dst + lat->num_bonds);
  if (with_sites)
    for (int s = 0; s < lat->num_sites; ++s) { xsrc.push_back(s); xdst.push_back(-1); }
  Partition P;
  make_partition(*lat, (int)xsrc.size(), xsrc.data(), xdst.data(), tile_sites > 0 ? tile_sites : 64, P);
  out->num_tiles = P.T; out->num_classes = P.nclasses;
  out->max_bonds = P.nbmax; out->max_sites = P.nsmax;
  out->max_halo_buckets = P.hmax; out->max_walk_halo = P.whmax; out->max_ksites = P.nksmax; out->max_degree = P.zmax;
  out->owned_bonds = P.bond_base[P.T];
  out->halo_buckets = P.halo_off[P.T];
}

int lq_tiling_info(const lq_lattice* lat, int32_t tile_sites, int32_t with_sites, lq_tiling* out) {
  if (!lat || !out || lat->num_sites <= 0 || lat->num_bonds < 0 || (lat->num_bonds > 0 && (!lat->src || !lat->dst))) {
    g_err = "bad argument";
    return LQ_E_INVALID;
  }
  LQ_TRY(tiling_info_impl(lat, tile_sites, with_sites, out))
}

// peer >= 0: the checksums cover only the segments shared with that rank (so that the two sides can be compared)
static void space_plan_impl(const lq_lattice* lat, int32_t tile_sites, int32_t with_sites, int32_t nranks, int32_t rank,
                            int32_t peer, lq_space_plan* out) {
  if (nranks < 2 || rank < 0 || rank >= nranks || peer >= nranks) fail(LQ_E_INVALID, "bad rank / nranks");
  for (int b = 0; b < lat->num_bonds; ++b)
    if (lat->src[b] < 0 || lat->src[b] >= lat->num_sites || lat->dst[b] < 0 || lat->dst[b] >= lat->num_sites ||
        lat->src[b] == lat->dst[b])
      fail(LQ_E_INVALID, "bond endpoint out of range");
  std::vector<int> xsrc(lat->src, lat->src + lat->num_bonds);
  std::vector<int> xdst(lat->dst, lat->dst + lat->num_bonds);
  if (with_sites)
    for (int s = 0; s < lat->num_sites; ++s) { xsrc.push_back(s); xdst.push_back(-1); }
  const int ts = tile_sites > 0 ? tile_sites : 64;
  Partition G, L;
  make_partition(*lat, (int)xsrc.size(), xsrc.data(), xdst.data(), ts, G);
  SpacePlan S;
  plan_space(G, nranks, rank, S);
  make_partition(*lat, (int)xsrc.size(), xsrc.data(), xdst.data(), ts, L, &S.relabel);
  *out = lq_space_plan{};
  out->owned_tiles = S.To; out->walked_ghost_tiles = S.Tw; out->ghost_tiles = S.Tloc - S.To;
  out->owned_sites = L.site_base[S.To]; out->walked_sites = L.site_base[S.To + S.Tw]; out->local_sites = L.site_base[S.Tloc];
  std::vector<char> nb(nranks, 0);
  auto mix = [](int64_t h, int64_t v) { return (int64_t)(((uint64_t)h * 1099511628211ull) ^ (uint64_t)(v + 0x9e3779b97f4a7c15ull)); };
  for (const SpaceSeg& g : S.segs) {
    if (g.owner != rank && g.user != rank) continue;
    const int other = g.owner == rank ? g.user : g.owner;
    out->segments++;
    nb[other] = 1;
    const bool counted = peer < 0 || other == peer;
    int64_t& sum = g.owner == rank ? out->checksum_owner : out->checksum_user;
    if (g.owner == rank) { out->owner_bonds += (int64_t)g.bonds.size(); out->owner_sites += (int64_t)g.sites.size(); }
    else { out->user_bonds += (int64_t)g.bonds.size(); out->user_sites += (int64_t)g.sites.size(); }
    if (counted) {
      for (int s2 : g.sites) sum = mix(sum, s2);
      for (int b2 : g.bonds) sum = mix(sum, (int64_t)b2 + ((int64_t)1 << 40));
    }
    // every listed bond / site must lie in a tile this rank holds
    for (int s2 : g.sites) if (L.site_e2i[s2] >= out->local_sites) fail(LQ_E_INVALID, "boundary site outside the local tiles (internal error)");
    for (int b2 : g.bonds) if (L.bond_tile[L.bond_e2i[b2]] >= S.Tloc) fail(LQ_E_INVALID, "boundary bond outside the local tiles (internal error)");
  }
  for (int r = 0; r < nranks; ++r) out->neighbours += nb[r];
}

int lq_space_plan_info(const lq_lattice* lat, int32_t tile_sites, int32_t with_sites, int32_t nranks, int32_t rank,
                       int32_t peer, lq_space_plan* out) {
  if (!lat || !out || lat->num_sites <= 0 || lat->num_bonds < 0 || (lat->num_bonds > 0 && (!lat->src || !lat->dst))) {
    g_err = "bad argument";
    return LQ_E_INVALID;
  }
  LQ_TRY(space_plan_impl(lat, tile_sites, with_sites, nranks, rank, peer, out))
}

int lq_get_info(lq_handle h, lq_info* out) {
  if (!h || !out) { g_err = "bad argument"; return LQ_E_INVALID; }
  out->num_tiles = h->part.T;
  out->num_windows = h->W;
  out->page_capacity = h->cap;
  out->threads_per_page = h->k1_nt;
  out->op_capacity = h->ncap;
  out->cluster_capacity = h->nccap;
  out->device_bytes = (int64_t)h->device_bytes;
  out->sm_count = h->sm_count;
  out->nodes_per_op = h->npo;
  return LQ_OK;
}

int lq_enable_timers(lq_handle h, int on) {
  if (!h) { g_err = "null handle"; return LQ_E_INVALID; }
  h->timers_on = on != 0;
  return LQ_OK;
}

int64_t lq_kernel_launches(lq_handle h) { return h ? h->launches : 0; }
int64_t lq_regrow_count(lq_handle h) { return h ? h->regrows : 0; }
int64_t lq_h2d_bytes(lq_handle h) { return h ? h->h2d_bytes : 0; }
int64_t lq_d2h_bytes(lq_handle h) { return h ? h->d2h_bytes : 0; }

int lq_set_comm(lq_handle h, const lq_comm* comm) {
  if (!h || !comm) { g_err = "bad argument"; return LQ_E_INVALID; }
  h->comm = *comm;
  h->has_comm = true;
  return LQ_OK;
}

int lq_comm_unique_id(void* id_out) {
  if (!id_out) { g_err = "bad argument"; return LQ_E_INVALID; }
  LQ_TRY({
    static_assert(sizeof(ncclUniqueId) == LQ_NCCL_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    nccl_check(nccl_api().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(id_out, &id, sizeof id);
  })
}

int lq_comm_init(lq_handle h, const void* nccl_unique_id, int32_t rank, int32_t nranks) {
  if (!h || !nccl_unique_id) { g_err = "bad argument"; return LQ_E_INVALID; }
  LQ_TRY({
    if (rank != h->opt.rank || nranks != h->opt.nranks)
      fail(LQ_E_INVALID, "lq_comm_init: rank / nranks differ from lq_options of this engine");
    if (nranks < 2) fail(LQ_E_INVALID, "lq_comm_init on a serial engine");
    if (h->nccl) fail(LQ_E_INVALID, "lq_comm_init called twice");
    CK(cudaSetDevice(h->opt.device));
    ncclUniqueId id;
    std::memcpy(&id, nccl_unique_id, sizeof id);
    nccl_check(nccl_api().CommInitRank(&h->nccl, nranks, id, rank), "ncclCommInitRank");
    h->has_comm = true;
  })
}

void* lq_stream(lq_handle h) { return h ? (void*)h->stream : nullptr; }

// experiment counters (LQ_DBG=1; not part of the public header): reads and clears 8 values
int lq_debug_counters(lq_handle h, unsigned long long* out) {
  if (!h || !out) return LQ_E_INVALID;
  cudaSetDevice(h->opt.device);
  cudaStreamSynchronize(h->stream);
  cudaMemcpy(out, h->dbgc.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaMemset(h->dbgc.p, 0, 8 * sizeof(unsigned long long));
  return LQ_OK;
}

// test hook (not part of the public header): histogram of the K1 Poisson draw, see lq_k1.cuh
static void debug_poisson_impl(double mean, long long count, unsigned long long seed, unsigned long long* hist, int nbins) {
  unsigned long long* d = nullptr;
  CK(cudaMalloc((void**)&d, nbins * sizeof(unsigned long long)));
  CK(cudaMemset(d, 0, nbins * sizeof(unsigned long long)));
  lq::k_debug_poisson<<<592, 256>>>(mean, count, (uint32_t)seed, (uint32_t)(seed >> 32), d, nbins);
  CK(cudaMemcpy(hist, d, nbins * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  CK(cudaFree(d));
}
int lq_debug_poisson(double mean, long long count, unsigned long long seed, unsigned long long* hist, int nbins) {
  if (!hist || nbins <= 0 || !(mean > 0)) return LQ_E_INVALID;
  LQ_TRY(debug_poisson_impl(mean, count, seed, hist, nbins))
}

const char* lq_last_error(void) { return g_err.c_str(); }
const char* lq_version(void) { return "alps-looper_b200 0.1 (sm_100a)"; }

}  // extern "C"
