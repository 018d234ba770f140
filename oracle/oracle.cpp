// oracle.cpp -- CPU restatement of the ALPS/looper loop update.  TEST INFRASTRUCTURE ONLY:
// see oracle.h for the rules (never linked, imported or called by the product path) and for the
// golden vectors that pin it.  Citations are file:line under the reference root.
//
// Build: see oracle/Makefile (g++ -O2 -shared -fPIC; also the `oracle_loop` CLI that prints the
// text of standalone/loop.op).

#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <vector>

namespace {

// ---------------------------------------------------------------------------------------------
// standalone/union_find.h:41-54 (node), :69-72 (add), :74-77 (root_index), :103-125
// (update_link, unify_weight).  parent_ < 0 : root holding -weight.
// ---------------------------------------------------------------------------------------------
struct sa_node {
  int parent = -1;
  int id = 0;
  bool is_root() const { return parent < 0; }
  int weight() const { return -parent; }
};

inline int sa_add(std::vector<sa_node>& v) {
  v.push_back(sa_node());
  return int(v.size()) - 1;
}
inline int sa_root_index(std::vector<sa_node> const& v, int g) {
  while (!v[g].is_root()) g = v[g].parent;
  return g;
}
inline void sa_update_link(std::vector<sa_node>& v, int g, int r) {
  while (g != r) {
    int p = v[g].parent;
    v[g].parent = r;
    g = p;
  }
}
// standalone/union_find.h:112-125: heavier root wins, ties keep r0.
inline int sa_unify(std::vector<sa_node>& v, int g0, int g1) {
  int r0 = sa_root_index(v, g0);
  int r1 = sa_root_index(v, g1);
  if (r0 != r1) {
    if (v[r0].weight() < v[r1].weight()) std::swap(r0, r1);
    v[r0].parent += v[r1].parent;
    v[r1].parent = r0;
  }
  sa_update_link(v, g0, r0);
  sa_update_link(v, g1, r0);
  return r0;
}

// ---------------------------------------------------------------------------------------------
// looper/union_find.h:57-82 (node: parent_ <= 0 root with -weight, else parent+1),
// :111-136 (node_noweight), :158-172 (root_index_ph), :242-284 (unify_pathhalving, serial branch)
// ---------------------------------------------------------------------------------------------
struct lp_node {
  int parent_ = -1;
  bool is_root() const { return parent_ <= 0; }
  void set_parent(int p) { parent_ = p + 1; }
  int parent() const { return parent_ - 1; }
  void set_weight(int w) { parent_ = -w; }
  int weight() const { return -parent_; }
};
struct lp_node_noweight {
  int parent_ = -1;  // id_mask: root with id 0
  bool is_root() const { return parent_ <= 0; }
  void set_parent(int p) { parent_ = p + 1; }
  int parent() const { return parent_ - 1; }
  void set_weight(int) { parent_ = 0 ^ -1; }
  int weight() const { return 0; }
};
template <class T>
int lp_root_index(std::vector<T> const& v, int g) {
  while (!v[g].is_root()) g = v[g].parent();
  return g;
}
template <class T>
int lp_root_index_ph(std::vector<T>& v, int g) {
  if (v[g].is_root()) return g;
  while (true) {
    int p = v[g].parent();
    if (v[p].is_root()) return p;
    v[g].set_parent(v[p].parent());
    g = p;
  }
}
template <class T>
int lp_unify(std::vector<T>& v, int g0, int g1) {
  int r0 = lp_root_index_ph(v, g0);
  int r1 = lp_root_index_ph(v, g1);
  if (r0 != r1) {
    if (v[r0].weight() < v[r1].weight()) std::swap(r0, r1);
    v[r0].set_weight(v[r0].weight() + v[r1].weight());
    v[r1].set_parent(r0);
  }
  return r0;
}

// standalone/observable.h:29-41
struct observable {
  double sum = 0, esq = 0;
  unsigned count = 0;
  void operator<<(double x) { sum += x; esq += x * x; ++count; }
  double mean() const { return count > 0 ? sum / count : 0.; }
  double error() const {
    return count > 1 ? std::sqrt((esq / count - mean() * mean()) / (count - 1)) : 0.;
  }
};

inline double power2(double x) { return x * x; }
inline double power4(double x) { return power2(power2(x)); }

// looper/susceptibility.h:97-105 estimate (8 doubles)
struct lp_estimate {
  double usize0 = 0, umag0 = 0, usize = 0, umag = 0;
  double ssize0 = 0, smag0 = 0, ssize = 0, smag = 0;
  // susceptibility.h:126-132 end_s ; :117-119 begin_s = end_s(-t)
  void end_s(double gg, double t, int c) {
    usize += t * 0.5;
    umag += t * (0.5 - c);
    ssize += gg * t * 0.5;
    smag += gg * t * (0.5 - c);
  }
  void begin_s(double gg, double t, int c) { end_s(gg, -t, c); }
  // susceptibility.h:139-146 start_bottom
  void start_bottom(double gg, double t, int c) {
    begin_s(gg, t, c);
    usize0 += 0.5;
    umag0 += (0.5 - c);
    ssize0 += gg * 0.5;
    smag0 += gg * (0.5 - c);
  }
};

// susceptibility.h:182-198 collector += estimate ; standalone/common.h:67-72
inline void collect(orc_collector& c, lp_estimate const& e) {
  c.umag0 += e.umag0;
  c.usize2 += power2(e.usize0);
  c.umag2 += power2(e.umag0);
  c.usize4 += power4(e.usize0);
  c.umag4 += power4(e.umag0);
  c.usize += power2(e.usize);
  c.umag += power2(e.umag);
  c.smag0 += e.smag0;
  c.ssize2 += power2(e.ssize0);
  c.smag2 += power2(e.smag0);
  c.ssize4 += power4(e.ssize0);
  c.smag4 += power4(e.smag0);
  c.ssize += power2(e.ssize);
  c.smag += power2(e.smag);
}

struct sa_estimate {  // standalone/common.h:44-56
  double mag = 0, size = 0, length = 0;
};

struct op_t {  // standalone/common.h:29-37
  int type;    // 0 diagonal, 1 offdiagonal (bit0) | graph<<2
  int bond;
  int upper_cluster, lower_cluster;
  double time;
};

}  // namespace

struct orc_sim {
  int nsites, nbonds;
  std::vector<int> src, dst;
  std::vector<double> gauge;
  double beta;
  std::mt19937 eng;
  std::uniform_real_distribution<> d_uniform;
  std::exponential_distribution<> d_time;
  std::vector<op_t> operators, operators_p;
  std::vector<int> spins, current, spins_before;
  std::vector<sa_node> fragments;
  std::vector<char> to_flip;
  std::vector<int> built_type;  // operator types as built (pre flip)
  int nc = 0;
  bool looper_estimators = true;   // orc_set_looper_estimators
  std::vector<sa_estimate> estimates;   // kept across sweeps like loop.C:64 (resize(0); resize(nc))
  orc_sim(int ns, int nb, const int32_t* s, const int32_t* d, const double* g, double b,
          uint32_t seed)
      : nsites(ns), nbonds(nb), src(s, s + nb), dst(d, d + nb), gauge(ns, 0.0), beta(b),
        eng(seed), d_time(b * nb / 2), spins(ns, 0), current(ns) {
    if (g) gauge.assign(g, g + ns);
  }
};

extern "C" {

orc_sim* orc_create(int nsites, int nbonds, const int32_t* src, const int32_t* dst,
                    const double* gauge, double beta, uint32_t seed) {
  return new orc_sim(nsites, nbonds, src, dst, gauge, beta, seed);
}
void orc_destroy(orc_sim* s) { delete s; }
void orc_set_looper_estimators(orc_sim* s, int on) { s->looper_estimators = on != 0; }

// standalone/loop.C:87-167, with left()/right() (common.h:92-93) read from the bond table.
void orc_sweep(orc_sim* S, orc_collector* out) {
  const int nsites = S->nsites, nbonds = S->nbonds;
  std::vector<op_t>&operators = S->operators, &operators_p = S->operators_p;
  std::vector<int>&spins = S->spins, &current = S->current;
  std::vector<sa_node>& fragments = S->fragments;
  // (test hook for orc_get_last_graph; not a statement of loop.C -- skipped on the timed CPU legs)
  if (S->looper_estimators) S->spins_before = spins;

  // loop.C:87
  std::swap(operators, operators_p);
  operators.resize(0);
  // loop.C:90-91
  fragments.resize(0);
  fragments.resize(nsites);
  for (int s = 0; s < nsites; ++s) current[s] = s;

  // loop.C:93-126
  double t = S->d_time(S->eng);
  for (std::vector<op_t>::iterator opi = operators_p.begin();
       t < 1 || opi != operators_p.end();) {
    if (opi == operators_p.end() || t < opi->time) {
      const int b = static_cast<int>(nbonds * S->d_uniform(S->eng));
      if (spins[S->src[b]] != spins[S->dst[b]]) {
        op_t o;
        o.type = 0;
        o.bond = b;
        o.time = t;
        operators.push_back(o);
        t += S->d_time(S->eng);
      } else {
        t += S->d_time(S->eng);
        continue;
      }
    } else {
      if (opi->type == 0) {
        ++opi;
        continue;
      } else {
        operators.push_back(*opi);
        ++opi;
      }
    }
    op_t& oi = operators.back();
    const int s0 = S->src[oi.bond];
    const int s1 = S->dst[oi.bond];
    if (oi.type == 1) {
      spins[s0] ^= 1;
      spins[s1] ^= 1;
    }
    oi.lower_cluster = sa_unify(fragments, current[s0], current[s1]);
    oi.upper_cluster = current[s0] = current[s1] = sa_add(fragments);
  }
  // loop.C:128
  for (int s = 0; s < nsites; ++s) sa_unify(fragments, s, current[s]);

  // loop.C:135-137
  int nc = 0;
  for (auto& f : fragments)
    if (f.is_root()) f.id = nc++;
  for (auto& f : fragments) f.id = fragments[sa_root_index(fragments, int(&f - &fragments[0]))].id;
  S->nc = nc;
  S->to_flip.assign(nc, 0);
  std::vector<sa_estimate>& estimates = S->estimates;
  estimates.resize(0);
  estimates.resize(nc);
  std::vector<lp_estimate> lest(S->looper_estimators ? nc : 0);

  // loop.C:141-151
  for (auto& op : operators) {
    const double tt = op.time;
    estimates[fragments[op.lower_cluster].id].length += 2 * tt;
    estimates[fragments[op.upper_cluster].id].length -= 2 * tt;
  }
  for (int s = 0; s < nsites; ++s) {
    const int id = fragments[s].id;
    estimates[id].mag += 1 - 2 * spins[s];
    estimates[id].size += 1;
    estimates[id].length += 1;
  }
  // looper estimator on the same graph: path_integral.C:682-734 (accum_i), HAF graph 0:
  // loop_l0 = loop_l1 = loop0, loop_u0 = loop_u1 = loop1 (graph_impl.h:489-494)
  if (S->looper_estimators) {
    std::vector<int> sc(spins);  // spins at tau = 0 (== spins after the walk, periodic)
    for (int s = 0; s < nsites; ++s)
      lest[fragments[s].id].start_bottom(S->gauge[s], 0.0, sc[s]);
    for (auto& op : operators) {
      const int s0 = S->src[op.bond], s1 = S->dst[op.bond];
      lp_estimate& lo = lest[fragments[op.lower_cluster].id];
      lo.end_s(S->gauge[s0], op.time, sc[s0]);
      lo.end_s(S->gauge[s1], op.time, sc[s1]);
      if (op.type & 1) {
        sc[s0] ^= 1;
        sc[s1] ^= 1;
      }
      lp_estimate& up = lest[fragments[op.upper_cluster].id];
      up.begin_s(S->gauge[s0], op.time, sc[s0]);
      up.begin_s(S->gauge[s1], op.time, sc[s1]);
    }
    for (int s = 0; s < nsites; ++s)
      lest[fragments[current[s]].id].end_s(S->gauge[s], 1.0, sc[s]);
  }

  // loop.C:154-157 and path_integral.C:774-777
  orc_collector coll;
  std::memset(&coll, 0, sizeof(coll));
  coll.nop = double(operators.size());
  coll.nc = nc;
  for (auto& est : estimates) {
    coll.sa_usus += est.mag * est.mag;
    coll.sa_smag += est.size * est.size;
    coll.sa_ssus += est.length * est.length;
  }
  for (auto& e : lest) collect(coll, e);
  // path_integral.C:850-851 with energy_offset = sum of bond offsets = nbonds/4 (weight_impl.h:187)
  coll.ene = 0.25 * nbonds - coll.nop / S->beta;

  if (S->looper_estimators) {   // test hook (orc_get_last_graph), skipped on the timed CPU legs
    S->built_type.resize(operators.size());
    for (size_t i = 0; i < operators.size(); ++i) S->built_type[i] = operators[i].type;
  }

  // loop.C:160
  for (int c = 0; c < nc; ++c) S->to_flip[c] = (S->d_uniform(S->eng) < 0.5);
  // loop.C:163-167
  for (auto& op : operators)
    if (S->to_flip[fragments[op.lower_cluster].id] ^ S->to_flip[fragments[op.upper_cluster].id])
      op.type ^= 1;
  for (int s = 0; s < nsites; ++s)
    if (S->to_flip[fragments[s].id]) spins[s] ^= 1;

  if (out) *out = coll;
}

int64_t orc_num_ops(const orc_sim* S) { return int64_t(S->operators.size()); }

void orc_get_state(const orc_sim* S, int32_t* spins, orc_op* ops) {
  for (int s = 0; s < S->nsites; ++s) spins[s] = S->spins[s];
  for (size_t i = 0; i < S->operators.size(); ++i) {
    ops[i].time = S->operators[i].time;
    ops[i].loc = (S->operators[i].bond << 1) | 1;
    ops[i].type = S->operators[i].type;
  }
}

void orc_set_state(orc_sim* S, const int32_t* spins, const orc_op* ops, int64_t n) {
  for (int s = 0; s < S->nsites; ++s) S->spins[s] = spins[s];
  S->operators.resize(size_t(n));
  for (int64_t i = 0; i < n; ++i) {
    S->operators[i].time = ops[i].time;
    S->operators[i].bond = ops[i].loc >> 1;
    S->operators[i].type = ops[i].type & 1;
    S->operators[i].lower_cluster = S->operators[i].upper_cluster = 0;
  }
}

void orc_get_last_graph(const orc_sim* S, int32_t* spins_before, orc_op* ops_built,
                        int32_t* lower_id, int32_t* upper_id, int32_t* site_id, int32_t* nc,
                        int32_t* flip) {
  for (int s = 0; s < S->nsites; ++s) {
    if (spins_before) spins_before[s] = S->looper_estimators ? S->spins_before[s] : -1;   // hooks are off on the timed legs
    if (site_id) site_id[s] = S->fragments[s].id;
  }
  for (size_t i = 0; i < S->operators.size(); ++i) {
    if (ops_built) {
      ops_built[i].time = S->operators[i].time;
      ops_built[i].loc = (S->operators[i].bond << 1) | 1;
      ops_built[i].type = S->looper_estimators ? S->built_type[i] : -1;
    }
    if (lower_id) lower_id[i] = S->fragments[S->operators[i].lower_cluster].id;
    if (upper_id) upper_id[i] = S->fragments[S->operators[i].upper_cluster].id;
  }
  if (nc) *nc = S->nc;
  if (flip)
    for (int c = 0; c < S->nc; ++c) flip[c] = S->to_flip[c];
}

// path_integral.C:539-566 (walk) + graph_impl.h:277-295 (xxz reconnect; g=0 is the HAF case
// :168-177) + graph_impl.h:79-86 (site reconnect) + path_integral.C:584-588 (close) +
// union_find.h:325-343 (set_id/copy_id): the cluster graph of a GIVEN configuration.
struct cluster_graph {
  std::vector<int> l0, l1, u0, u1;  // fragment of the four legs of every operator
  std::vector<int> current;         // fragment crossing tau = 1 on every site
  std::vector<int> id;              // cluster id of every fragment
  int nc = 0;
};

static int build_graph(int nsites, int nbonds, const int32_t* src, const int32_t* dst,
                       const int32_t* spins, const orc_op* ops, int64_t n, cluster_graph& G) {
  std::vector<lp_node> fragments(nsites);
  fragments.reserve(size_t(nsites) + size_t(n));
  std::vector<int>& current = G.current;
  current.resize(nsites);
  std::vector<int> sc(spins, spins + nsites);
  for (int s = 0; s < nsites; ++s) current[s] = s;
  std::vector<int>&l0 = G.l0, &l1 = G.l1, &u0 = G.u0, &u1 = G.u1;
  l0.assign(n, 0); l1.assign(n, 0); u0.assign(n, 0); u1.assign(n, 0);
  double tprev = -1;
  for (int64_t k = 0; k < n; ++k) {
    if (ops[k].time < tprev) return -4;
    tprev = ops[k].time;
    if (!(ops[k].loc & 1)) {
      // site operator (path_integral.C:552-561): graph_impl.h:79-86 site_graph_type::reconnect --
      // the world line is cut, the part above is a new fragment; no compatibility condition (:69)
      const int s = ops[k].loc >> 1;
      if (s < 0 || s >= nsites || (ops[k].type >> 2) != 0) return -3;
      if (ops[k].type & 1) sc[s] ^= 1;
      fragments.push_back(lp_node());
      l0[k] = l1[k] = current[s];
      u0[k] = u1[k] = current[s] = int(fragments.size()) - 1;
      continue;
    }
    const int b = ops[k].loc >> 1;
    if (b < 0 || b >= nbonds) return -3;
    const int s0 = src[b], s1 = dst[b];
    const int g = ops[k].type >> 2;
    if (ops[k].type & 1) {
      // off-diagonal (S+S- / S-S+): antiparallel spins below; its graph comes from
      // choose_offdiagonal (graph_impl.h:324-327), i.e. 0 or 1, with no compatibility test
      if (sc[s0] == sc[s1] || (g & 2)) return -1;
    } else {
      // diagonal: graph_impl.h:257 is_compatible(g, c0, c1) = (g & 1) ^ c0 ^ c1 (path_integral.C:503)
      if (!((g & 1) ^ sc[s0] ^ sc[s1])) return -1;
    }
    if (ops[k].type & 1) {
      sc[s0] ^= 1;
      sc[s1] ^= 1;
    }
    int loop0, loop1;
    if ((g & 2) == 2) {  // frozen: graph_impl.h:281-282
      loop0 = loop1 = current[s0] = current[s1] = lp_unify(fragments, current[s0], current[s1]);
    } else if (g == 0) {  // graph_impl.h:283-287
      fragments.push_back(lp_node());
      loop0 = lp_unify(fragments, current[s0], current[s1]);
      loop1 = current[s0] = current[s1] = int(fragments.size()) - 1;
    } else {  // cross: graph_impl.h:288-292
      loop0 = current[s0];
      loop1 = current[s1];
      std::swap(current[s0], current[s1]);
    }
    // graph_impl.h:489-494
    l0[k] = loop0;
    l1[k] = (g == 0) ? loop0 : loop1;
    u0[k] = loop1;
    u1[k] = (g == 0) ? loop1 : loop0;
  }
  for (int s = 0; s < nsites; ++s)
    if (sc[s] != spins[s]) return -1;  // not periodic in imaginary time
  for (int s = 0; s < nsites; ++s) lp_unify(fragments, s, current[s]);

  // union_find.h:325-343
  const int nf = int(fragments.size());
  std::vector<int>& id = G.id;
  id.assign(nf, -1);
  int nc = 0;
  for (int i = 0; i < nf; ++i)
    if (fragments[i].is_root()) id[i] = nc++;
  for (int i = 0; i < nf; ++i) id[i] = id[lp_root_index(fragments, i)];
  G.nc = nc;
  return 0;
}

// orc_build_clusters: the graph above + path_integral.C:666-734 (improved accumulators) +
// :774-777 (collect).
int orc_build_clusters(int nsites, int nbonds, const int32_t* src, const int32_t* dst,
                       const double* gauge, const int32_t* spins, const orc_op* ops, int64_t n,
                       int32_t* labels_out, int64_t* nc_out, orc_collector* coll_out) {
  cluster_graph G;
  const int rc = build_graph(nsites, nbonds, src, dst, spins, ops, n, G);
  if (rc != 0) return rc;
  const std::vector<int>&l0 = G.l0, &l1 = G.l1, &u0 = G.u0, &u1 = G.u1, &current = G.current, &id = G.id;
  const int nc = G.nc;
  std::vector<int> sc(spins, spins + nsites);

  // canonical min-index labels over leg nodes: site s -> s, upper legs of op k -> N+2k, N+2k+1
  if (labels_out) {
    std::vector<int> minidx(nc, -1);
    auto touch = [&](int cid, int idx) {
      if (minidx[cid] < 0) minidx[cid] = idx;  // visited in increasing idx order
    };
    for (int s = 0; s < nsites; ++s) touch(id[s], s);
    for (int64_t k = 0; k < n; ++k) {
      touch(id[u0[k]], int(nsites + 2 * k));
      touch(id[u1[k]], int(nsites + 2 * k + 1));
    }
    for (int s = 0; s < nsites; ++s) labels_out[s] = minidx[id[s]];
    for (int64_t k = 0; k < n; ++k) {
      labels_out[nsites + 2 * k] = minidx[id[u0[k]]];
      labels_out[nsites + 2 * k + 1] = minidx[id[u1[k]]];
    }
  }
  if (nc_out) *nc_out = nc;

  if (coll_out) {
    std::vector<lp_estimate> lest(nc);
    std::vector<sa_estimate> est(nc);
    std::vector<char> cut(nc, 0);  // transmag.h:64-81: closed == false once a site leg touches it
    std::vector<double> gg(nsites, 0.0);
    if (gauge) gg.assign(gauge, gauge + nsites);
    sc.assign(spins, spins + nsites);
    for (int s = 0; s < nsites; ++s) {
      lest[id[s]].start_bottom(gg[s], 0.0, sc[s]);
      est[id[s]].mag += 1 - 2 * sc[s];
      est[id[s]].size += 1;
      est[id[s]].length += 1;
    }
    for (int64_t k = 0; k < n; ++k) {
      const double t = ops[k].time;
      if (!(ops[k].loc & 1)) {  // path_integral.C:716-726
        const int s = ops[k].loc >> 1;
        lest[id[l0[k]]].end_s(gg[s], t, sc[s]);
        est[id[l0[k]]].length += t;
        if (ops[k].type & 1) sc[s] ^= 1;
        lest[id[u0[k]]].begin_s(gg[s], t, sc[s]);
        est[id[u0[k]]].length -= t;
        cut[id[l0[k]]] = cut[id[u0[k]]] = 1;
        continue;
      }
      const int b = ops[k].loc >> 1;
      const int s0 = src[b], s1 = dst[b];
      const int g = ops[k].type >> 2;
      if ((g & 2) == 2) {  // frozen graphs are skipped by the accumulators (path_integral.C:692)
        if (ops[k].type & 1) { sc[s0] ^= 1; sc[s1] ^= 1; }
        continue;
      }
      lest[id[l0[k]]].end_s(gg[s0], t, sc[s0]);
      lest[id[l1[k]]].end_s(gg[s1], t, sc[s1]);
      est[id[l0[k]]].length += t;
      est[id[l1[k]]].length += t;
      if (ops[k].type & 1) { sc[s0] ^= 1; sc[s1] ^= 1; }
      lest[id[u0[k]]].begin_s(gg[s0], t, sc[s0]);
      lest[id[u1[k]]].begin_s(gg[s1], t, sc[s1]);
      est[id[u0[k]]].length -= t;
      est[id[u1[k]]].length -= t;
    }
    for (int s = 0; s < nsites; ++s) lest[id[current[s]]].end_s(gg[s], 1.0, sc[s]);
    orc_collector coll;
    std::memset(&coll, 0, sizeof(coll));
    coll.nop = double(n);
    coll.nc = nc;
    for (auto& e : lest) collect(coll, e);
    for (auto& e : est) {
      coll.sa_usus += e.mag * e.mag;
      coll.sa_smag += e.size * e.size;
      coll.sa_ssus += e.length * e.length;
    }
    // transmag.h:98-101: collector.length += closed ? 0 : estimate.length, where the estimate's
    // length (begin/end/start_bottom/stop_top, :72-92) is the total length of the cluster's legs
    // = 2 * usize of susceptibility.h:126-132 (every leg adds +-t there with weight 1/2)
    for (int c = 0; c < nc; ++c)
      if (cut[c]) coll.tlen += 2 * lest[c].usize;
    *coll_out = coll;
  }
  return 0;
}

// stiffness.h:82-133 improved estimator on the cluster graph of a given configuration: per cluster
// winding[i] += (1-2c) vr[i] where a leg ends on the SOURCE side of a bond operator (end_bs, :98-102),
// -= (1-2c) vr[i] where one begins there (begin_bs, :93-97); frozen bond graphs are skipped by the
// accumulators (path_integral.C:692); site operators do not contribute (:92,:103).
// Returns w2 = sum over clusters and dimensions of (winding / 2)^2 (:125-128); also the normal
// estimator (:137-170) of the same configuration in *w2_normal.
double orc_stiffness(int nsites, int nbonds, const int32_t* src, const int32_t* dst,
                     const double* bond_vectors, int dim, const int32_t* spins, const orc_op* ops,
                     int64_t n, double* w2_normal) {
  cluster_graph G;
  if (build_graph(nsites, nbonds, src, dst, spins, ops, n, G) != 0) return -1;
  std::vector<double> wind(size_t(G.nc) * 3, 0.0);
  double total[3] = {0, 0, 0};
  std::vector<int> sc(spins, spins + nsites);
  for (int64_t k = 0; k < n; ++k) {
    if (!(ops[k].loc & 1)) {
      if (ops[k].type & 1) sc[ops[k].loc >> 1] ^= 1;
      continue;
    }
    const int b = ops[k].loc >> 1, s0 = src[b], s1 = dst[b];
    const int g = ops[k].type >> 2;
    const double* vr = bond_vectors + 3 * size_t(b);
    const bool frozen = (g & 2) == 2;
    for (int i = 0; i < dim; ++i) {
      if (!frozen) wind[size_t(G.id[G.l0[k]]) * 3 + i] += (1 - 2 * sc[s0]) * vr[i];
      total[i] += (1 - 2 * sc[s0]) * vr[i];   // normal_estimator: every bond operator (path_integral.C:529-531)
    }
    if (ops[k].type & 1) { sc[s0] ^= 1; sc[s1] ^= 1; }
    for (int i = 0; i < dim; ++i) {
      if (!frozen) wind[size_t(G.id[G.u0[k]]) * 3 + i] -= (1 - 2 * sc[s0]) * vr[i];
      total[i] -= (1 - 2 * sc[s0]) * vr[i];
    }
  }
  double w2 = 0;
  for (double w : wind) w2 += power2(0.5 * w);
  if (w2_normal) {
    *w2_normal = 0;
    for (int i = 0; i < dim; ++i) *w2_normal += power2(0.5 * total[i]);
  }
  return w2;
}

// ---------------------------------------------------------------------------------------------
// Generic model: path_integral.C:403-864 restated (XXZ bond graphs 0..3 + site graphs), serial.
// RNG: std::mt19937 + std::exponential_distribution for the gaps (path_integral.C:413-423 uses
// boost's with the ALPS generator -- unpinned in the reference) and a cumulative-weight
// table in place of alps::random_choice (graph_impl.h:679; same distribution).
// ---------------------------------------------------------------------------------------------
struct orc_model_sim {
  int nsites, nbonds;
  std::vector<int> src, dst;
  std::vector<double> gauge, bw, sw;
  std::vector<double> cum;          // cumulative graph weights
  std::vector<int> cum_graph;       // (pos << 3 | g << 1 | is_bond)
  double beta, total;
  std::mt19937 eng;
  std::uniform_real_distribution<> uni;
  std::vector<int> spins;
  std::vector<orc_op> ops;
};

orc_model_sim* orc_model_create(int nsites, int nbonds, const int32_t* src, const int32_t* dst,
                                const double* gauge, const double* bond_weights,
                                const double* site_weights, double beta, uint32_t seed) {
  orc_model_sim* S = new orc_model_sim();
  S->nsites = nsites; S->nbonds = nbonds;
  S->src.assign(src, src + nbonds); S->dst.assign(dst, dst + nbonds);
  S->gauge.assign(nsites, 0.0);
  if (gauge) S->gauge.assign(gauge, gauge + nsites);
  S->bw.assign(bond_weights, bond_weights + 4 * size_t(nbonds));
  S->sw.assign(nsites, 0.0);
  if (site_weights) S->sw.assign(site_weights, site_weights + nsites);
  S->beta = beta;
  S->eng.seed(seed);
  S->spins.assign(nsites, 0);  // path_integral.C:225
  double acc = 0;
  for (int s = 0; s < nsites; ++s)     // graph_impl.h:560-600: site graphs, then bond graphs
    if (S->sw[s] > 0) { acc += S->sw[s]; S->cum.push_back(acc); S->cum_graph.push_back(s << 3); }
  for (int b = 0; b < nbonds; ++b)
    for (int g = 0; g < 4; ++g)
      if (S->bw[4 * size_t(b) + g] > 0) {
        acc += S->bw[4 * size_t(b) + g];
        S->cum.push_back(acc);
        S->cum_graph.push_back((b << 3) | (g << 1) | 1);
      }
  S->total = acc;
  return S;
}
void orc_model_destroy(orc_model_sim* S) { delete S; }
int64_t orc_model_num_ops(const orc_model_sim* S) { return int64_t(S->ops.size()); }
void orc_model_get_state(const orc_model_sim* S, int32_t* spins, orc_op* ops) {
  for (int s = 0; s < S->nsites; ++s) spins[s] = S->spins[s];
  for (size_t i = 0; i < S->ops.size(); ++i) ops[i] = S->ops[i];
}

int orc_model_sweep(orc_model_sim* S, orc_collector* out) {
  const int nsites = S->nsites;
  std::vector<orc_op> ops_p;
  std::swap(ops_p, S->ops);
  std::vector<int> sc(S->spins);
  // path_integral.C:403-425 fill times
  std::vector<double> times;
  if (S->total > 0) {
    std::exponential_distribution<> expdist(S->beta * S->total);
    double t = 0;
    while (t < 1) { t += expdist(S->eng); times.push_back(t); }
  } else {
    times.push_back(2.0);
  }
  // path_integral.C:484-566 diagonal update (the reconnect part is done by build_graph below)
  size_t tmi = 0, opi = 0;
  while (opi < ops_p.size() || times[tmi] < 1) {
    orc_op o;
    if (opi == ops_p.size() || times[tmi] < ops_p[opi].time) {
      const double r = S->uni(S->eng) * S->total;
      size_t k = size_t(std::upper_bound(S->cum.begin(), S->cum.end(), r) - S->cum.begin());
      if (k >= S->cum.size()) k = S->cum.size() - 1;
      const int cg = S->cum_graph[k];
      o.time = times[tmi++];
      if (cg & 1) {
        const int b = cg >> 3, g = (cg >> 1) & 3;
        // graph_impl.h:257 is_compatible(g, c0, c1) = (g & 1) ^ c0 ^ c1
        if (!((g & 1) ^ sc[S->src[b]] ^ sc[S->dst[b]])) continue;
        o.loc = (b << 1) | 1;
        o.type = g << 2;
      } else {
        o.loc = (cg >> 3) << 1;
        o.type = 0;
      }
    } else {
      o = ops_p[opi++];
      if (!(o.type & 1)) continue;  // diagonal operators are removed (path_integral.C:519-521)
      if (o.loc & 1) {
        const int b = o.loc >> 1;
        // graph_impl.h:311-327 choose_offdiagonal: g = 0 with probability v0 / (v0 + v1), else 1
        const double v0 = S->bw[4 * size_t(b)], v1 = S->bw[4 * size_t(b) + 1];
        const double pr = (v0 + v1 > 1e-10) ? v0 / (v0 + v1) : 1;
        const int g = (S->uni(S->eng) < pr) ? 0 : 1;
        o.type = (g << 2) | 1;
        sc[S->src[b]] ^= 1;
        sc[S->dst[b]] ^= 1;
      } else {
        sc[o.loc >> 1] ^= 1;
      }
    }
    S->ops.push_back(o);
  }
  cluster_graph G;
  const int rc = build_graph(nsites, S->nbonds, S->src.data(), S->dst.data(), S->spins.data(),
                             S->ops.data(), int64_t(S->ops.size()), G);
  if (rc != 0) return rc;
  if (out) {
    orc_build_clusters(nsites, S->nbonds, S->src.data(), S->dst.data(), S->gauge.data(),
                       S->spins.data(), S->ops.data(), int64_t(S->ops.size()), nullptr, nullptr, out);
    double off = 0;  // model.energy_offset(): bond offsets = weight/2 (weight_impl.h:187), site = v0 (:80)
    for (double v : S->bw) off += v / 2;
    for (double v : S->sw) off += v;
    out->ene = off - out->nop / S->beta;   // path_integral.C:850-851
  }
  // path_integral.C:796-823 flip
  std::vector<char> flip(G.nc);
  for (int c = 0; c < G.nc; ++c) flip[c] = (S->uni(S->eng) < 0.5);
  for (size_t k = 0; k < S->ops.size(); ++k) {
    // operator.h loop_0 / loop_1 in leg form: the cluster below and the one above on the source side
    // (frozen graphs never change type: both legs are one cluster)
    if (flip[G.id[G.l0[k]]] ^ flip[G.id[G.u0[k]]]) S->ops[k].type ^= 1;
  }
  for (int s = 0; s < nsites; ++s)
    if (flip[G.id[s]]) S->spins[s] ^= 1;
  return 0;
}

// test/union_find.C:40-74 with boost::mt19937(29833u) == std::mt19937(29833) and
// boost::uniform_real<>()(eng) == eng() / 2^32.
int orc_union_find_replay(char* buf, int buflen) {
  const int n = 100;
  std::mt19937 eng(29833u);
  auto rng = [&]() { return eng() / 4294967296.0; };
  std::string out;
  char line[256];
  out += "[[union find test]]\n";
  std::vector<lp_node> nodes(n);
  std::vector<lp_node_noweight> nodes_noweight(n);
  out += "\n[making tree]\n";
  for (int i = 0; i < n; i++) {
    int i0 = static_cast<int>(n * rng());
    int i1 = static_cast<int>(n * rng());
    std::snprintf(line, sizeof line, "connecting node %d to node %d\n", i0, i1);
    out += line;
    lp_unify(nodes, i0, i1);
    lp_unify(nodes_noweight, i0, i1);
  }
  out += "\n[results]\n";
  for (int i = 0; i < n; i++) {
    if (nodes[i].is_root())
      std::snprintf(line, sizeof line, "node %d is root and tree size is %d\n", i,
                    nodes[i].weight());
    else
      std::snprintf(line, sizeof line, "node %d's parent is %d and its root is %d\n", i,
                    nodes[i].parent(), lp_root_index(nodes, i));
    out += line;
  }
  for (int i = 0; i < n; i++) {  // test/union_find.C:68-74 prints `nodes` again
    if (nodes[i].is_root())
      std::snprintf(line, sizeof line, "node %d is root\n", i);
    else
      std::snprintf(line, sizeof line, "node %d's parent is %d and its root is %d\n", i,
                    nodes[i].parent(), lp_root_index(nodes, i));
    out += line;
  }
  if (buf && buflen > 0) {
    int m = std::min<int>(buflen - 1, int(out.size()));
    std::memcpy(buf, out.data(), m);
    buf[m] = 0;
  }
  return int(out.size());
}

// looper/weight_impl.h:165-188; crop_0/crop_01 from looper/crop.h.
void orc_xxz_weights(double jxy_in, double jz, double a, double v[4], double* offset, int* sign) {
  auto crop_0 = [](double x) { return x > 0 ? x : 0.0; };
  a = a < 0 ? 0 : (a > 1 ? 1 : a);
  if (sign) *sign = (jxy_in <= 0 ? 1 : -1);
  double jxy = std::abs(jxy_in);
  if (jxy + std::abs(jz) > 1e-10) {
    if (jxy - jz > 2 * a * jxy) {
      v[0] = crop_0(std::min(jxy / 2, (jxy + jz) / 4));
      v[1] = crop_0(std::min(jxy / 2, (jxy - jz) / 4));
      v[2] = crop_0(-(jxy - jz) / 2.0);
      v[3] = crop_0(-(jxy + jz) / 2);
    } else {
      v[0] = (1 - a) * jxy / 2;
      v[1] = a * jxy / 2;
      v[2] = -((1 - 2 * a) * jxy - jz) / 2;
      v[3] = 0;
    }
  } else {
    v[0] = v[1] = v[2] = v[3] = 0;
  }
  if (offset) *offset = (v[0] + v[1] + v[2] + v[3]) / 2;
}

// standalone/loop.C:39-195 end to end.
void orc_run_chain(int length, double temperature, unsigned sweeps, unsigned therm,
                   double out[10]) {
  std::vector<int32_t> src(length), dst(length);
  std::vector<double> gauge(length);
  for (int b = 0; b < length; ++b) {
    src[b] = b;                            // common.h:92
    dst[b] = (b == length - 1) ? 0 : b + 1;  // common.h:93
    gauge[b] = (b & 1) ? -1 : 1;
  }
  const double beta = 1 / temperature;
  orc_sim* S = orc_create(length, length, src.data(), dst.data(), gauge.data(), beta, 29833);
  S->looper_estimators = false;   // loop.C's own statements only (its three sums; no looper collector, no test hooks)
  observable num_clusters, energy, usus, smag, ssus;
  for (unsigned mcs = 0; mcs < therm + sweeps; ++mcs) {
    orc_collector coll;
    orc_sweep(S, &coll);
    if (mcs >= therm) {  // loop.C:173-179
      num_clusters << coll.nc;
      energy << (0.25 * length - coll.nop / beta) / length;
      usus << 0.25 * beta * coll.sa_usus / length;
      smag << 0.25 * coll.sa_smag;
      ssus << 0.25 * beta * coll.sa_ssus / length;
    }
  }
  orc_destroy(S);
  out[0] = num_clusters.mean(); out[1] = num_clusters.error();
  out[2] = energy.mean();       out[3] = energy.error();
  out[4] = usus.mean();         out[5] = usus.error();
  out[6] = smag.mean();         out[7] = smag.error();
  out[8] = ssus.mean();         out[9] = ssus.error();
}

// looper/poisson_distribution.h:44-113 (both branches) driven as test/poisson_distribution.C:33-69
// does: boost::mt19937 default seed (= std::mt19937 default, 5489), boost::uniform_real<> on a
// 32-bit engine = eng() / 2^32.  Writes the exact text of test/poisson_distribution.op
// (setprecision(3) rows "r poisson frequency error") into buf; returns the length needed.
int orc_poisson_replay(double mean, int count, char* buf, int buflen, int64_t* bins_out, int nbins_out) {
  std::mt19937 eng;
  auto rng = [&]() { return eng() / 4294967296.0; };
  const double exp_mean = std::exp(-mean);                       // poisson_distribution.h:63
  const bool big = mean >= 16;                                   // THRESHOLD, :45,64
  const double sqr = big ? std::sqrt(2 * mean) : 0, alxm = big ? std::log(mean) : 0;
  const double gm = big ? mean * alxm - std::lgamma(mean + 1) : 0;
  auto draw = [&]() -> int {
    if (!big) {                                                  // :81-88 O(mean) product method
      double product = 1;
      for (int m = 0;; ++m) {
        product *= rng();
        if (product <= exp_mean) return m;
      }
    }
    double em, y, t;                                             // :89-101 rejection method
    do {
      do {
        y = std::tan(M_PI * rng());
        em = sqr * y + mean;
      } while (em < 0.0);
      em = std::floor(em);
      t = 0.9 * (1 + y * y) * std::exp(em * alxm - std::lgamma(em + 1) - gm);
    } while (rng() > t);
    return int(em);
  };
  std::vector<int> bins(int(5 * mean), 0);                       // poisson_distribution.C:50-55
  for (int c = 0; c < count; ++c) {
    int r = draw();
    if (r < 5 * mean) ++bins[r];
  }
  // `std::cout << std::setprecision(3)` (:62-69) prints like printf("%.3g"); formatted with snprintf
  // because this library is loaded into Python processes that may already hold another libstdc++
  std::string out;
  double poi = std::exp(-mean);
  for (size_t r = 0; r < bins.size(); ++r) {
    char line[128];
    std::snprintf(line, sizeof line, "%zu %.3g %.3g %.3g\n", r, poi, (double)bins[r] / count,
                  std::sqrt((double)bins[r]) / count);
    out += line;
    poi *= mean / (r + 1);
  }
  if (buf && buflen > 0) {
    int m = std::min<int>(buflen - 1, int(out.size()));
    std::memcpy(buf, out.data(), m);
    buf[m] = 0;
  }
  if (bins_out)
    for (int r = 0; r < nbins_out; ++r) bins_out[r] = r < (int)bins.size() ? bins[r] : 0;
  return int(out.size());
}

// ---------------------------------------------------------------------------------------------
// SSE worker, sse.C:168-407 restated (serial, XXZ bond graphs 0..3 + site graphs), on the same
// model tables as orc_model_sim.  The operator string has no times: the "time" of operator k is its
// position (sse.C:251-283 `t`), the top is the string length (:283).  The string is exported with
// time = (k + 1/2) / n so that orc_model_get_state / orc_build_clusters see an ordered list; the
// cluster sums are linear in the time, so the SSE collector is the one of that list with
// times k/n, scaled by n (sums) and n^2 (squared sums) -- done in orc_sse_collect below.
// RNG as orc_model_sim (the reference's generator is ALPS's: unpinned).
// ---------------------------------------------------------------------------------------------
// collector of a GIVEN string (spins at the bottom + operators in string order), SSE times
int orc_sse_collect(int nsites, int nbonds, const int32_t* src, const int32_t* dst, const double* gauge,
                    const int32_t* spins, const orc_op* ops_in, int64_t n, orc_collector* out) {
  std::vector<orc_op> ops(ops_in, ops_in + n);
  for (int64_t k = 0; k < n; ++k) ops[k].time = n > 0 ? double(k) / double(n) : 0;   // exact scaling below
  const int rc = orc_build_clusters(nsites, nbonds, src, dst, gauge, spins, ops.data(), n, nullptr, nullptr, out);
  if (rc != 0) return rc;
  const double f = double(n), f2 = f * f;
  // the "0" sums (tau = 0 magnetisations) and counters do not carry a time
  out->usize *= f2; out->umag *= f2; out->ssize *= f2; out->smag *= f2;
  out->sa_ssus *= f2;
  out->tlen *= f;
  return 0;
}

int orc_sse_sweep(orc_model_sim* S, orc_collector* out) {
  const int nsites = S->nsites;
  std::vector<orc_op> ops_p;
  std::swap(ops_p, S->ops);
  std::vector<int> sc(S->spins);
  int nop = int(ops_p.size());                                  // sse.C:182
  const double bw = S->beta * S->total;                         // sse.C:205 beta * model.graph_weight()
  bool try_gap = true;
  size_t opi = 0;
  // sse.C:207-249 (the reconnect / accumulate part, :251-281, is done on the finished string below)
  while (try_gap || opi != ops_p.size()) {
    orc_op o;
    if (try_gap) {
      if ((nop + 1) * S->uni(S->eng) < bw) {                    // :209
        const double r = S->uni(S->eng) * S->total;             // :210 model.choose_graph
        size_t k = size_t(std::upper_bound(S->cum.begin(), S->cum.end(), r) - S->cum.begin());
        if (k >= S->cum.size()) k = S->cum.size() - 1;
        const int cg = S->cum_graph[k];
        bool ok;
        if (cg & 1) {
          const int b = cg >> 3, g = (cg >> 1) & 3;
          ok = ((g & 1) ^ sc[S->src[b]] ^ sc[S->dst[b]]) != 0;  // :211-213, graph_impl.h:257
          o.loc = (b << 1) | 1;
          o.type = g << 2;
        } else {
          ok = true;                                            // site graph, graph_impl.h:69
          o.loc = (cg >> 3) << 1;
          o.type = 0;
        }
        if (!ok) { try_gap = false; continue; }                 // :217-219
        ++nop;                                                  // :215
      } else {
        try_gap = false;                                        // :221-223
        continue;
      }
    } else {
      o = ops_p[opi];
      if (!(o.type & 1)) {                                      // diagonal
        if (bw * S->uni(S->eng) < nop) { --nop; ++opi; continue; }   // :226-229 remove
        if (o.loc & 1) {                                        // :234-237 choose_diagonal (graph_impl.h:307-311)
          const int b = o.loc >> 1;
          const double* v = &S->bw[4 * size_t(b)];
          const int c = 1 ^ sc[S->src[b]] ^ sc[S->dst[b]];      // 0 antiparallel, 1 parallel
          const double den = c ? v[1] + v[3] : v[0] + v[2];
          const double pr = den > 1e-10 ? (c ? v[1] : v[0]) / den : 1;
          o.type = (((S->uni(S->eng) < pr) ? 0 : 2) ^ c) << 2;
        } else {
          S->uni(S->eng);                                       // :231-233 (site: graph 0; one number is drawn)
        }
      } else if (o.loc & 1) {                                   // :240-244 choose_offdiagonal (graph_impl.h:324-327)
        const int b = o.loc >> 1;
        const double v0 = S->bw[4 * size_t(b)], v1 = S->bw[4 * size_t(b) + 1];
        const double pr = (v0 + v1 > 1e-10) ? v0 / (v0 + v1) : 1;
        o.type = (((S->uni(S->eng) < pr) ? 0 : 1) << 2) | 1;
      }
      ++opi;
      try_gap = true;
    }
    if (o.type & 1) {                                           // :257-261, :270-274 walk the spins
      if (o.loc & 1) { sc[S->src[o.loc >> 1]] ^= 1; sc[S->dst[o.loc >> 1]] ^= 1; }
      else sc[o.loc >> 1] ^= 1;
    }
    S->ops.push_back(o);
  }
  const int64_t n = int64_t(S->ops.size());
  for (int64_t k = 0; k < n; ++k) S->ops[k].time = (double(k) + 0.5) / double(n);
  cluster_graph G;
  const int rc = build_graph(nsites, S->nbonds, S->src.data(), S->dst.data(), S->spins.data(),
                             S->ops.data(), n, G);
  if (rc != 0) return rc;
  if (out) {
    orc_sse_collect(nsites, S->nbonds, S->src.data(), S->dst.data(), S->gauge.data(), S->spins.data(),
                    S->ops.data(), n, out);
    double off = 0;
    for (double v : S->bw) off += v / 2;
    for (double v : S->sw) off += v;
    out->ene = off - double(n) / S->beta;                       // sse.C:398-399
  }
  // sse.C:366-381 flip
  std::vector<char> flip(G.nc);
  for (int c = 0; c < G.nc; ++c) flip[c] = (2 * S->uni(S->eng) - 1) < 0;
  for (size_t k = 0; k < S->ops.size(); ++k)
    if (flip[G.id[G.l0[k]]] ^ flip[G.id[G.u0[k]]]) S->ops[k].type ^= 1;
  for (int s = 0; s < nsites; ++s)
    if (flip[G.id[s]]) S->spins[s] ^= 1;
  return 0;
}

}  // extern "C"
