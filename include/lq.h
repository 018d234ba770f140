/*
 * lq.h -- C ABI of the B200-native loop-update engine ("lq" = loop QMC).
 *
 * This is the drop-in boundary for ONE hot path of ALPS/looper: the body of
 * loop_worker::dispatch<FIELD=false,SIGN=false,IMPROVE=true>() (path_integral.C:355-864,
 * standalone/loop.C:80-180): diagonal update -> union-find cluster labelling -> cluster flip +
 * improved estimators.  The reference has no FFI; its seam is the C++ worker class
 * (path_integral.C:57-125).  Each entry point below names the reference code it replaces.
 * The host-side mirror of that class (looper::loop_worker in alps-looper_b200/looper/) is a thin
 * C++ layer over exactly these calls.
 *
 * Conventions: opaque handle; int return (0 = ok, negative = LQ_E_*; text via lq_last_error);
 * no exceptions cross the ABI; all pointers are caller-owned HOST memory; one host thread per
 * handle; one CUDA stream per handle.  There is NO CPU fallback: if no CUDA device / kernel image
 * is usable, lq_create fails with LQ_E_CUDA.
 */
#ifndef LQ_H
#define LQ_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LQ_OK            0
#define LQ_E_INVALID    -1   /* bad argument / inconsistent state                                */
#define LQ_E_CUDA       -2   /* CUDA runtime error (message has the cudaError string)            */
#define LQ_E_OVERFLOW   -3   /* a page / node / cluster arena is full: raise lq_options.reserve  */
#define LQ_E_NOMEM      -4
#define LQ_E_COMM       -5   /* multi-GPU exchange failed                                        */
#define LQ_E_UNSUPPORTED -6

typedef struct lq_engine* lq_handle;

/* Lattice = the reference's virtual graph for S=1/2 (looper/lattice.h:576-675): bond b joins
 * source(b)=src[b] and target(b)=dst[b] (lattice.h:49-62); gauge[s] = +1/-1 staggering sign
 * (lattice.h:85, read at susceptibility.h:56,129), all 0 if not bipartite.  dims is an optional
 * hint for the spatial tiling (site index = x + dims[0]*(y + dims[1]*z)); 0,0,0 = unknown. */
typedef struct lq_lattice {
  int32_t        num_sites;
  int32_t        num_bonds;
  const int32_t* src;
  const int32_t* dst;
  const double*  gauge;      /* may be NULL (= not bipartite)                                    */
  int32_t        dims[3];
  /* optional, for the spin stiffness (looper/stiffness.h:63-76): relative lattice vector of every
   * bond (ALPS bond_vector_relative of the real graph: the bond vector over the lattice extent, so
   * that a world line wrapping once has winding 1), 3 doubles per bond; within one dimension all
   * components must be integer multiples of the smallest non-zero one.  vector_dim = spatial
   * dimension to measure (stiffness.h MAX_DIM = 3); NULL / 0 = not measured. */
  int32_t        vector_dim;
  const double*  bond_vectors;
} lq_lattice;

/* Model = graph weights per bond, the output of weight_helper (looper/weight_impl.h:349-423):
 * v[4*b + g] for the XXZ graphs g = 0..3 (graph_impl.h:93-103); S=1/2 HAF J=1 is {0.5,0,0,0}
 * (test/weight.op "Jxy = 1, Jz = 1").  bond_weights == NULL means uniform_weights for all bonds.
 * energy_offset = model.energy_offset() (model.h:84) = sum of bond offsets. */
typedef struct lq_model {
  const double* bond_weights;     /* 4 * num_bonds, or NULL                                      */
  double        uniform_weights[4];
  double        energy_offset;
  /* site graph weight v0 = |Hx|/2 of site_weight_helper (weight_impl.h:62-88; transverse field):
   * per site, or NULL = uniform_site_weight for all sites.  0 = no site operators. */
  const double* site_weights;     /* num_sites, or NULL                                          */
  double        uniform_site_weight;
} lq_model;

typedef struct lq_options {
  uint64_t seed;             /* Philox key; the reference seeds mt19937 (loop.C:50)              */
  int32_t  device;           /* CUDA device ordinal                                              */
  int32_t  tile_sites;       /* spatial tile size in sites (0 = default 64)                      */
  double   window_ops;       /* target mean candidate count per bond and time window (0 = 3.0)   */
  double   reserve;          /* page capacity / mean candidate count (0 = default 1.7); the
                                analogue of RESERVE_OPERATORS (path_integral.C:240-243)          */
  double   cluster_reserve;  /* cluster arena / operator arena (0 = default 0.75)                */
  int32_t  rank, nranks;     /* imaginary-time slab of this engine (path_integral_mpi.C:231-232);
                                nranks <= 1 = serial                                             */
  int32_t  flags;            /* bit 0: section timers on (ALPS_ENABLE_TIMER)                     */
  int32_t  representation;   /* LQ_REPR_PATH_INTEGRAL ("loop; path integral", path_integral.C:873) or
                                LQ_REPR_SSE ("loop; sse", sse.C:411): the improved estimators are
                                those of the fixed-length operator string -- the "time" of an
                                operator is its position in the string and the string length n is the
                                top (sse.C:251-283,358-361) -- and the collector's sums are in those
                                units (commit: susceptibility.h:213-215, beta (x/n + x0) / (n+1) / V) */
  int32_t  cut;              /* nranks > 1: how the configuration is shared among the engines.
                                LQ_CUT_TIME: imaginary-time slabs (path_integral_mpi.C:231-232,
                                looper/parallel.h); LQ_CUT_SPACE: every engine owns a contiguous range of
                                spatial tiles over the whole imaginary-time axis (the bond ownership of
                                looper/lattice.h:692-787 taken across GPUs), keeps ghost copies of the
                                neighbouring tiles and merges the clusters that cross a cut through the
                                same all-gather + all-reduce pair                                   */
} lq_options;
#define LQ_REPR_PATH_INTEGRAL 0
#define LQ_REPR_SSE           1
#define LQ_CUT_TIME           0
#define LQ_CUT_SPACE          1

/* looper/operator.h:120-143: {time_, loc_, type_}; loc = pos<<1 | is_bond
 * (location_impl.h:37); type bit0 = offdiagonal, bits >= 2 = graph type (operator.h:62,76). */
typedef struct lq_op {
  double  time;
  int32_t loc;
  int32_t type;
} lq_op;

/* basic_measurement::collector nop_/nc_/noc_ (looper/measurement.h:366-372), energy ene_
 * (energy.h:56), susceptibility improved collector (susceptibility.h:158-160), transmag. */
typedef struct lq_collector {
  double nop, nc, noc;
  double ene;
  double umag0, usize2, umag2, usize4, umag4, usize, umag;
  double smag0, ssize2, smag2, ssize4, smag4, ssize, smag;
  double tlen;   /* transverse_magnetization collector length (transmag.h:95-101): total length of
                    the clusters cut by a site operator; "Transverse Magnetization" = tlen / 2   */
  double w2;     /* stiffness collector (stiffness.h:118-133): sum over clusters and dimensions of
                    (winding / 2)^2; "Stiffness" = w2 / (beta * vector_dim)                      */
} lq_collector;

/* Section timers; ids mirror path_integral.C:284-299 (4 init .. 16 measurement). */
typedef struct lq_timer {
  int32_t id;
  int32_t count;
  double  seconds;           /* device time (CUDA events)                                        */
  char    label[40];
} lq_timer;

typedef struct lq_info {
  int32_t num_tiles, num_windows, page_capacity, threads_per_page;
  int64_t op_capacity, cluster_capacity;
  int64_t device_bytes;
  int32_t sm_count, nodes_per_op;
} lq_info;

/* Host-side spatial tiling of a lattice (no GPU needed; the part of lq_create that replaces the
 * OpenMP bond ownership of looper/lattice.h:692-787): tiles of neighbouring sites, every bond owned
 * by the tile of its source site, per-tile halo of the foreign bonds that touch a site of an owned
 * bond.  with_sites != 0 adds the one-ended pseudo-bond of every site (site graphs).  For tests
 * and for choosing lq_options.tile_sites. */
typedef struct lq_tiling {
  int32_t num_tiles, num_classes;      /* tiles; distinct tile shapes (stencils are shared per shape) */
  int32_t max_bonds, max_sites;        /* owned by one tile                                        */
  int32_t max_halo_buckets, max_walk_halo, max_ksites, max_degree;
  int64_t owned_bonds, halo_buckets;   /* sums over the tiles (owned_bonds == bonds incl. pseudo)  */
} lq_tiling;
int lq_tiling_info(const lq_lattice* lat, int32_t tile_sites, int32_t with_sites, lq_tiling* out);

/* Host-side plan of the spatial cut (lq_options.cut = LQ_CUT_SPACE; no GPU needed): what rank `rank` of
 * `nranks` owns and mirrors.  The part of lq_create that takes the lattice sharing of
 * looper/lattice.h:692-787 across GPUs.  For tests and for sizing. */
typedef struct lq_space_plan {
  int32_t owned_tiles, walked_ghost_tiles, ghost_tiles;   /* ghost_tiles includes the walked ones          */
  int32_t owned_sites, walked_sites, local_sites;         /* local = owned + sites of all ghost tiles     */
  int32_t segments;                                       /* boundary segments this rank takes part in    */
  int32_t neighbours;                                     /* ranks it exchanges halo streams with         */
  int64_t owner_bonds, owner_sites;                       /* entries of the segments it OWNS (sends)      */
  int64_t user_bonds, user_sites;                         /* ... of the segments it uses (receives)       */
  int64_t checksum_owner, checksum_user;                  /* order-dependent hash of the owner / user side
                                                             lists (external ids): the user side of rank r
                                                             for owner q equals q's owner side for r      */
} lq_space_plan;
int lq_space_plan_info(const lq_lattice* lat, int32_t tile_sites, int32_t with_sites, int32_t nranks,
                       int32_t rank, int32_t peer, lq_space_plan* out);

/* Replaces loop_worker::loop_worker (path_integral.C:202-305): builds lattice tables, graph
 * chooser tables, sizes the arenas.  Initial state: all spins up, no operators (:225). */
int lq_create(lq_handle* out, const lq_lattice* lat, const lq_model* model, double beta,
              const lq_options* opt);
int lq_destroy(lq_handle h);

/* set_beta (path_integral.C:98-105).  Re-tiles imaginary time; operators are kept. */
int lq_set_beta(lq_handle h, double beta);

/* load()/save() payload (path_integral.C:111-124): spins at tau=0 (0 up / 1 down) and the
 * operator string sorted by time.  lq_get_state with ops == NULL only returns *n. */
int lq_set_state(lq_handle h, const int32_t* spins, const lq_op* ops, int64_t n);
int lq_get_state(lq_handle h, int32_t* spins, lq_op* ops, int64_t* n);
/* The random numbers of a step are Philox(seed; bond, window, STEP, draw): the step counter is the
 * whole generator state (the reference checkpoints its generator with the worker, alps::rng_helper).
 * Restore it together with the state, the same seed and the same lq_options.tile_sites /
 * window_ops (the keys use the engine's internal numbering) to continue the same Markov chain. */
uint32_t lq_get_step(lq_handle h);
int lq_set_step(lq_handle h, uint32_t step);

/* One Monte Carlo step = the whole of dispatch() (path_integral.C:355-864).  The collector is
 * that of THIS step's clusters; ene = energy_offset - nop/beta (:851). */
int lq_sweep(lq_handle h, lq_collector* out);
/* `count` steps back to back; out[i] for step i (out may be NULL). One device sync at the end. */
int lq_sweep_many(lq_handle h, int32_t count, lq_collector* out);

/* Cluster construction only (path_integral.C:539-588 + union_find.h:325-343) on the CURRENT
 * state, no diagonal update, no flip.  Operators are numbered k = 0..n-1 in the order
 * lq_get_state returns them.  labels_out has num_sites + 2n entries: cluster id of site s, then
 * of the two legs leaving operator k upwards (source side, target side).  Ids are canonical
 * min-index labels over that numbering (looper LOOPER_USE_DETERMINISTIC_UNIFY convention,
 * union_find.h:229-233), so they compare bit-for-bit with the reference partition. */
int lq_build_clusters(lq_handle h, int32_t* labels_out, int64_t* nc_out, lq_collector* coll_out);

int lq_timers(lq_handle h, lq_timer* out, int32_t* count);   /* timer.summarize (timer.hpp:184) */
/* ALPS_ENABLE_TIMER at run time (CMakeLists.txt:34-48): CUDA events around every phase */
int lq_enable_timers(lq_handle h, int on);
int lq_get_info(lq_handle h, lq_info* out);
/* number of kernel launches issued by this handle so far */
int64_t lq_kernel_launches(lq_handle h);
/* The arenas are sized at lq_create (lq_options.reserve, cluster_reserve).  When a step overflows
 * one of them, a serial engine rewinds to the configuration the step started from, enlarges that
 * arena by 1.5x and repeats the step -- the analogue of the reference's growing vectors
 * (RESERVE_* only pre-size them, path_integral.C:240-243).  The trajectory does not depend on the
 * capacities.  This counts those events; LQ_E_OVERFLOW is returned after 8 of them in one call and
 * by slab engines (nranks > 1), whose configuration is then undefined. */
int64_t lq_regrow_count(lq_handle h);
/* bytes copied host->device (per-step inputs, from pinned memory) and device->host (collectors)
 * by lq_sweep / lq_sweep_many so far */
int64_t lq_h2d_bytes(lq_handle h);
int64_t lq_d2h_bytes(lq_handle h);

/* Multi-GPU (imaginary-time slabs or spatial strips, one engine per GPU; replaces
 * parallel_cluster_unifier::unify, looper/parallel.h:1609-1809): the engine calls
 * exchange(ctx, ...) between the local labelling and the flip.  See INTEGRATION.md. */
typedef struct lq_comm {
  void* ctx;
  /* all-gather `bytes` bytes per rank (device pointers, on the engine's stream) */
  int (*all_gather)(void* ctx, const void* send_dev, void* recv_dev, int64_t bytes, void* stream);
  /* in-place sum all-reduce of `count` int64 values (device pointer) */
  int (*all_reduce_i64)(void* ctx, void* buf_dev, int64_t count, void* stream);
  /* LQ_CUT_SPACE only (may be NULL otherwise): send `send_bytes` to rank dst and receive `recv_bytes`
   * from rank src in one step (MPI_Sendrecv; the halo pages and spins of the ghost tiles).  Every rank
   * makes the same sequence of calls; either size may be 0. */
  int (*send_recv)(void* ctx, const void* send_dev, int64_t send_bytes, int32_t dst,
                   void* recv_dev, int64_t recv_bytes, int32_t src, void* stream);
} lq_comm;
int lq_set_comm(lq_handle h, const lq_comm* comm);
/* The engine's own data plane: an NCCL communicator over the ranks of the run (the reference hands
 * the worker a boost::mpi::communicator, path_integral_mpi.C:75,1012, and parallel_cluster_unifier
 * talks MPI itself, looper/parallel.h:1430-1600).  One rank calls lq_comm_unique_id, the host
 * distributes the LQ_NCCL_ID_BYTES bytes by whatever means it has (MPI_Bcast, a file, torch), every
 * rank calls lq_comm_init with its lq_options.rank / nranks; the all-gather and the all-reduce of
 * every step then run as ncclAllGather / ncclAllReduce on the engine's stream.  libnccl.so.2 is
 * bound at run time (LQ_NCCL_LIB overrides the name).  lq_set_comm remains for hosts that bring
 * their own transport. */
#define LQ_NCCL_ID_BYTES 128
int lq_comm_unique_id(void* id_out /* LQ_NCCL_ID_BYTES */);
int lq_comm_init(lq_handle h, const void* nccl_unique_id, int32_t rank, int32_t nranks);
void* lq_stream(lq_handle h);                                 /* cudaStream_t of the handle      */

const char* lq_last_error(void);
const char* lq_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LQ_H */
