/*
 * oracle.h -- CPU restatement of the ALPS/looper loop update (TEST INFRASTRUCTURE ONLY).
 *
 * This is the checker, never the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (alps-looper_b200/csrc, include/lq.h) never links, imports or calls anything here.
 *
 * Parity status: PINNED.  The restatement reproduces
 *   - standalone/loop.op           (bit-for-bit, `oracle_loop` binary, tests/test_oracle.py)
 *   - test/union_find.op           (bit-for-bit, orc_union_find_replay)
 *   - test/weight.op XXZ rows      (orc_xxz_weights)
 *   - test/poisson_distribution.op (bit-for-bit, orc_poisson_replay)
 *   - the reference binary itself  (oracle/_ref/loop, built from /root/reference/standalone/loop.C
 *                                   by oracle/Makefile when the reference tree is present)
 * The generic-model functions (orc_model_*: XXZ bond graphs + site graphs; orc_stiffness) restate
 * path_integral.C / graph_impl.h / transmag.h / stiffness.h, which do not compile here (ALPS) and
 * whose golden outputs depend on the ALPS generator: they are pinned statistically, to
 *   - the reference's OWN Monte Carlo results for every S = 1/2 task of loop.op and extras/{transmag,gap,
 *     corrlen,top,localsus}/*.op (tests/golden/ref_runs.json, tests/test_oracle_refruns.py): energy,
 *     magnetisations, susceptibilities and what only the algorithm defines -- "Number of Clusters", the
 *     generalised magnetisations, "Stiffness", "Transverse Magnetization" -- path integral and SSE,
 *     Heisenberg / Ising (frozen graphs) / transverse field (site graphs) / simple cubic,
 *   - exact diagonalisation (tests/golden/ed_*.json, whose generator reproduces the reference's
 *     'diagonalization' tasks to the printed digits; tests/test_oracle_model.py; the stiffness to
 *     <W^2> = beta F''(twist = 0), tests/golden/ed_stiffness.json, both estimators)
 * -- "parity unpinned" at the bit level for that part (no RNG-free golden exists for it).
 *
 * Every function cites the reference file:line it follows (paths relative to the
 * reference root, wistaria/alps-looper).
 */
#ifndef LOOPER_ORACLE_H
#define LOOPER_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* looper/operator.h:42-46,120-143 -- {time_, loc_, type_}; loc = pos<<1 | is_bond
 * (looper/location_impl.h:37); type bit0 = offdiagonal, bits>=2 = graph type. */
typedef struct orc_op {
  double  time;
  int32_t loc;
  int32_t type;
} orc_op;

/* basic_measurement::collector (looper/measurement.h:366-372) + energy (looper/energy.h:56)
 * + susceptibility::improved_estimator::collector (looper/susceptibility.h:158-160)
 * + the three standalone sums (standalone/common.h:67-72). */
typedef struct orc_collector {
  double nop, nc, noc;
  double ene;
  double umag0, usize2, umag2, usize4, umag4, usize, umag;
  double smag0, ssize2, smag2, ssize4, smag4, ssize, smag;
  /* standalone/common.h collector_t */
  double sa_usus, sa_smag, sa_ssus;
  /* transverse_magnetization collector (looper/transmag.h:95-101) */
  double tlen;
} orc_collector;

typedef struct orc_sim orc_sim;

/* Generic lattice: bond b joins src[b]--dst[b]; gauge[s] = +-1 (looper/lattice.h:85). */
orc_sim* orc_create(int nsites, int nbonds, const int32_t* src, const int32_t* dst,
                    const double* gauge, double beta, uint32_t seed);
void     orc_destroy(orc_sim*);
/* on = 0: orc_sweep runs the statements of standalone/loop.C only (3 cluster sums, common.h:67-72)
 * and leaves the 14 looper sums at zero -- the timing configuration of bench.py's CPU legs; the
 * random numbers drawn do not depend on it.  Default on. */
void     orc_set_looper_estimators(orc_sim*, int on);

/* One Monte Carlo step, statement-for-statement standalone/loop.C:87-167 (uniform bond choice,
 * weight 1/2 per bond).  Fills the collector from the clusters of this step. */
void     orc_sweep(orc_sim*, orc_collector* out);
int64_t  orc_num_ops(const orc_sim*);
/* State after the last sweep: spins at tau=0 and the operator string (types after the flip). */
void     orc_get_state(const orc_sim*, int32_t* spins, orc_op* ops);
void     orc_set_state(orc_sim*, const int32_t* spins, const orc_op* ops, int64_t n);
/* Graph of the last sweep (before the flip): spins before the update, operators as built
 * (pre-flip types), reference cluster ids (fragments[x].id()) of lower/upper cluster per operator
 * and per site, and the flip decision per cluster id. */
void     orc_get_last_graph(const orc_sim*, int32_t* spins_before, orc_op* ops_built,
                            int32_t* lower_id, int32_t* upper_id, int32_t* site_id,
                            int32_t* nc, int32_t* flip /* size nc */);

/* Cluster construction only (path_integral.C:539-588 / standalone/loop.C:117-128 restated for a
 * GIVEN configuration): spins at tau=0 and operators sorted by time (HAF graph 0 / XXZ graphs
 * 0..3 in type>>2; site operators, loc bit 0 clear, with the site graph of graph_impl.h:67-87).
 * Outputs canonical min-index labels for the N + n nodes (node N+k = k-th operator in the given order; for graphs without a new fragment the label is that of the
 * cluster passing above source site), the number of clusters, and the collector of
 * looper/susceptibility.h:97-198 and standalone/loop.C:141-157 for that configuration.
 * Returns 0, or -1 if the operator string is inconsistent with the spins. */
int      orc_build_clusters(int nsites, int nbonds, const int32_t* src, const int32_t* dst,
                            const double* gauge, const int32_t* spins, const orc_op* ops,
                            int64_t n, int32_t* labels_out, int64_t* nc_out,
                            orc_collector* coll_out);

/* Generic model, one Monte Carlo step of path_integral.C:403-864 (serial): per-bond XXZ graph
 * weights v[4*b+g] (graph_impl.h:255-328) and per-site weights (graph_impl.h:67-87,
 * weight_impl.h:62-88).  Returns 0, or the error of orc_build_clusters. */
typedef struct orc_model_sim orc_model_sim;
orc_model_sim* orc_model_create(int nsites, int nbonds, const int32_t* src, const int32_t* dst,
                                const double* gauge, const double* bond_weights,
                                const double* site_weights, double beta, uint32_t seed);
void     orc_model_destroy(orc_model_sim*);
int      orc_model_sweep(orc_model_sim*, orc_collector* out);
int64_t  orc_model_num_ops(const orc_model_sim*);
void     orc_model_get_state(const orc_model_sim*, int32_t* spins, orc_op* ops);

/* SSE worker, sse.C:168-407 restated on the same model object (orc_model_create): one Monte Carlo
 * step of the grand-canonical diagonal update (:207-249: insert with (nop+1) u < beta W, remove with
 * beta W u < nop, choose_diagonal / choose_offdiagonal), cluster construction and flip.  The string
 * is exported by orc_model_get_state with times (k + 1/2) / n.  The collector carries the SSE sums:
 * operator k has time k, the top is n (sse.C:251-283,358-361).  Pinned to exact diagonalisation and,
 * statistically, to the SSE blocks of loop.op (tests/test_oracle_sse.py); "parity unpinned" at the
 * bit level like the rest of the generic model (ALPS generator). */
int      orc_sse_sweep(orc_model_sim*, orc_collector* out);
/* SSE collector of a GIVEN string: spins at the bottom and the operators in string order (their
 * `time` fields are ignored). */
int      orc_sse_collect(int nsites, int nbonds, const int32_t* src, const int32_t* dst,
                         const double* gauge, const int32_t* spins, const orc_op* ops, int64_t n,
                         orc_collector* out);

/* looper/stiffness.h:82-133 on a given configuration (bond_vectors: 3 doubles per bond, relative
 * lattice vectors): returns the improved-estimator collector w2 = sum_c sum_i (winding_i / 2)^2
 * ("Stiffness" = w2 / (beta dim)) and, in *w2_normal, the normal estimator (:137-170); -1 on an
 * illegal configuration. */
double   orc_stiffness(int nsites, int nbonds, const int32_t* src, const int32_t* dst,
                       const double* bond_vectors, int dim, const int32_t* spins, const orc_op* ops,
                       int64_t n, double* w2_normal);

/* looper/union_find.h:57-82,145-172,242-284 replayed as test/union_find.C:40-74 does;
 * writes the exact text of test/union_find.op into buf (returns length needed). */
int      orc_union_find_replay(char* buf, int buflen);

/* looper/weight_impl.h:165-188 (standard/ergodic XXZ solution); a = FORCE_SCATTER. */
void     orc_xxz_weights(double jxy, double jz, double a, double v[4], double* offset, int* sign);

/* standalone/loop.C main() end to end on a chain (golden: standalone/loop.op):
 * out = {nc, nc_err, ene, ene_err, usus, usus_err, smag, smag_err, ssus, ssus_err}. */
void     orc_run_chain(int length, double temperature, unsigned sweeps, unsigned therm,
                       double out[10]);

/* looper/poisson_distribution.h:44-113 replayed as test/poisson_distribution.C:33-69 does (MEAN,
 * COUNT from test/poisson_distribution.ip); writes the exact text of test/poisson_distribution.op
 * into buf (returns the length needed) and the raw histogram into bins_out. */
int      orc_poisson_replay(double mean, int count, char* buf, int buflen, int64_t* bins_out,
                            int nbins_out);

#ifdef __cplusplus
}
#endif
#endif
