// CPU unit test of the host-side looper:: mirror (no GPU needed, no liblq.so calls).
#include <cassert>
#include <cmath>
#include <iostream>
#include <random>
#include <sstream>
#include "../../alps-looper_b200/looper/lattice.h"
#include "../../alps-looper_b200/looper/measurement.h"
#include "../../alps-looper_b200/looper/model.h"
#include "../../alps-looper_b200/looper/montecarlo.h"
#include "../../alps-looper_b200/looper/parameters.h"
#include "../../alps-looper_b200/looper/union_find.h"

#define CHECK(x) do { if (!(x)) { std::cerr << "FAILED " #x " at line " << __LINE__ << "\n"; return 1; } } while (0)

int main() {
  using namespace looper;
  // --- union_find: the sequence of test/union_find.C:40-58; partition printed as min-index labels
  {
    const int n = 100;
    std::mt19937 eng(29833u);
    auto rng = [&]() { return eng() / 4294967296.0; };
    std::vector<union_find::node> nodes(n);
    std::vector<union_find::node_noweight> nw(n);
    for (int i = 0; i < n; ++i) {
      int i0 = int(n * rng()), i1 = int(n * rng());
      union_find::unify(nodes, i0, i1);
      union_find::unify(nw, i0, i1);
    }
    int nc = union_find::set_id(nodes, 0, n, 0);
    union_find::copy_id(nodes, 0, n);
    CHECK(nc == union_find::count_root(nodes, 0, n));
    std::cout << "uf";
    for (int i = 0; i < n; ++i) {
      CHECK(union_find::root_index(nodes, i) == union_find::root_index(nw, i));
      CHECK(union_find::root_index(nodes, i) <= i);  // min-index roots
      std::cout << ' ' << union_find::root_index(nodes, i);
    }
    std::cout << "\n";
    int wsum = 0;
    for (int i = 0; i < n; ++i) if (nodes[i].is_root()) wsum += nodes[i].weight();
    CHECK(wsum == n);
  }
  // --- weights: test/weight.op rows
  {
    xxz_bond_weight_helper w(bond_parameter_xxz(0, 1, 1));
    CHECK(w.v[0] == 0.5 && w.v[1] == 0 && w.v[2] == 0 && w.v[3] == 0 && w.offset == 0.25 && w.sign == -1);
    xxz_bond_weight_helper w2(bond_parameter_xxz(0, 1, 0.5));
    CHECK(w2.v[0] == 0.375 && w2.v[1] == 0.125 && w2.offset == 0.25);
    xxz_bond_weight_helper w3(bond_parameter_xxz(0, 1, 2));
    CHECK(w3.v[0] == 0.5 && w3.v[2] == 0.5 && w3.offset == 0.5);
    xxz_bond_weight_helper w4(bond_parameter_xxz(0, 1, 1), 0.1);
    CHECK(std::abs(w4.v[0] - 0.45) < 1e-15 && std::abs(w4.v[1] - 0.05) < 1e-15 && std::abs(w4.v[2] - 0.1) < 1e-15 && std::abs(w4.offset - 0.3) < 1e-15);
  }
  // --- parameters + lattice + model
  {
    Parameters p;
    std::istringstream in("LATTICE = \"square lattice\";\nL = 4; W = 6  // comment\nJ = 1;\nT = 0.5\nSWEEPS = 1024;\nT_START_0 = 2; T_DURATION_0 = 64;\n");
    p.parse(in);
    CHECK(p.value_or_default<int>("L", 0) == 4 && p.value_or_default<int>("W", 0) == 6);
    lattice_helper lat(p);
    CHECK(num_sites(lat.vg()) == 24 && num_bonds(lat.vg()) == 48 && lat.is_bipartite());
    for (int b = 0; b < num_bonds(lat.vg()); ++b)
      CHECK(gauge(source(b, lat.vg()), lat.vg()) * gauge(target(b, lat.vg()), lat.vg()) == -1);
    spinmodel_helper m(p, lat);
    CHECK(std::abs(m.graph_weight() - 24.0) < 1e-12);       // 48 bonds * 1/2
    CHECK(std::abs(m.energy_offset() - 12.0) < 1e-12);      // B/4 (standalone/loop.C:175)
    mc_steps mcs(p);
    CHECK(mcs.thermalization() == 128 && mcs.sweeps() == 1024);
    temperature t(p);
    CHECK(t.annealing_steps() == 64 && t(0) == 2.0 && std::abs(t(32) - 1.25) < 1e-12 && t(64) == 0.5 && t(1000) == 0.5);
    Parameters q; q["LATTICE"] = "chain lattice"; q.set("L", 8);
    lattice_helper ch(q);
    CHECK(target(7, ch.vg()) == 0 && source(3, ch.vg()) == 3);  // standalone/common.h:92-93
  }
  // --- observable binning: AR(1) series, error must exceed the naive one
  {
    std::mt19937 eng(1);
    std::normal_distribution<> g;
    observable o;
    double x = 0;
    for (int i = 0; i < (1 << 16); ++i) { x = 0.9 * x + g(eng); o << x; }
    CHECK(o.count() == (1u << 16));
    CHECK(o.error() > 2.5 * o.naive_error());
    CHECK(std::abs(o.mean()) < 6 * o.error());
    observable_set s;
    lq_collector c = lq_collector();
    c.nop = 100; c.ene = -5; c.umag2 = 2; c.umag4 = 3; c.usize = 7; c.smag = 7; c.smag2 = 4;
    energy::commit(s, c, 10.0, 16.0);
    susceptibility::commit(s, c, 10.0, 16.0, true);
    CHECK(s["Energy Density"].mean() == -5.0 / 16);
    CHECK(s["Energy^2"].mean() == 25 - 100 / 100.0);
    CHECK(s["Magnetization^4"].mean() == 3 * 4 - 2 * 3);
    CHECK(s["Staggered Susceptibility"].mean() == 10.0 * 7 / 16);
  }
  // --- transverse field (site graphs, weight_impl.h:62-88), transmag and stiffness commits
  {
    Parameters p; p["LATTICE"] = "square lattice"; p.set("L", 4); p.set("Jxy", -1.0); p.set("Jz", 0.5); p.set("Gamma", 0.6);
    lattice_helper lat(p);
    spinmodel_helper m(p, lat);
    CHECK(std::abs(m.site_weight() - 0.3) < 1e-15);
    CHECK(std::abs(m.graph_weight() - (32 * 0.5 + 16 * 0.3)) < 1e-12);     // bonds v0+v1 = 1/2, sites |Hx|/2
    CHECK(std::abs(m.energy_offset() - (32 * 0.25 + 16 * 0.3)) < 1e-12);   // weight_impl.h:80,187
    CHECK(lat.vg().dimension == 2 && lat.vg().bond_vector_relative.size() == 3 * 32);
    CHECK(lat.vg().bond_vector_relative[0] == 0.25 && lat.vg().bond_vector_relative[3 * 16 + 1] == 0.25);  // 1 / extent
    Parameters q = p; q.set("Jxy", 1.0);   // antiferromagnetic XY coupling + field: sign problem
    bool threw = false;
    try { spinmodel_helper bad(q, lat); } catch (const std::invalid_argument&) { threw = true; }
    CHECK(threw);
    observable_set s;
    lq_collector c = lq_collector();
    c.tlen = 6; c.w2 = 8;
    transverse_magnetization::commit(s, c, 16.0);
    stiffness::commit(s, c, 4.0, 2);
    CHECK(s["Transverse Magnetization"].mean() == 3.0 && s["Transverse Magnetization Density"].mean() == 3.0 / 16);
    CHECK(s["Stiffness"].mean() == 8.0 / (4.0 * 2));
  }
  // --- evaluated observables (energy.h:89-102 specific heat, susceptibility.h:340-376 Binder ratios): jackknife
  {
    std::mt19937 eng(7);
    std::normal_distribution<> g(1.0, 0.5);
    observable_set s;
    const int n = 40000;   // not a power of two: the last, incomplete bin must stay out of the jackknife
    for (int i = 0; i < n; ++i) {
      const double e = g(eng);
      s["Inverse Temperature"] << 2.0; s["Volume"] << 4.0;
      s["Energy"] << e; s["Energy^2"] << e * e;
      s["Magnetization^2"] << e * e; s["Magnetization^4"] << e * e * e * e;
    }
    CHECK(s["Energy"].bin_sums().size() >= 64 && s["Energy"].bin_sums().size() < 128);
    CHECK(s["Energy"].bin_sums().size() * s["Energy"].bin_size() <= (unsigned long)n);
    energy::evaluate(s);
    evaluate_susceptibility(s);
    CHECK(s.has_evaluated("Specific Heat") && s.has_evaluated("Binder Ratio of Magnetization"));
    CHECK(!s.has_evaluated("Binder Ratio of Staggered Magnetization"));          // operands missing: skipped like the reference's try/catch
    // beta^2 var(e) / vol = 4 * 0.25 / 4; jackknife error of a variance of n Gaussian samples: var * sqrt(2 / n)
    const std::pair<double, double> c = s.evaluated("Specific Heat");
    CHECK(std::abs(c.first - 0.25) < 5 * c.second && c.second > 0);
    CHECK(std::abs(c.second / (0.25 * std::sqrt(2.0 / n)) - 1) < 0.35);
    // <x^2>^2 / <x^4> of N(1, 1/2): (1 + 1/4)^2 / (1 + 6/4 + 3/16)
    const std::pair<double, double> b = s.evaluated("Binder Ratio of Magnetization");
    CHECK(std::abs(b.first - 1.5625 / 2.6875) < 5 * b.second && b.second > 0 && b.second < 0.01);
    // a plain mean through the same machinery: the jackknife error is the bin-level error of the mean
    CHECK(s.evaluate("mean", {"Energy"}, [](const std::vector<double>& x) { return x[0]; }));
    CHECK(std::abs(s.evaluated("mean").second / s["Energy"].naive_error() - 1) < 0.3);
    // the bins travel with the checkpoint
    std::stringstream ck;
    s.save(ck);
    observable_set t;
    t.load(ck);
    CHECK(bool(ck));
    energy::evaluate(t);
    CHECK(t.evaluated("Specific Heat") == c && t["Energy"].error() == s["Energy"].error());
    observable_set few;
    few["Energy"] << 1.0; few["Energy^2"] << 1.0; few["Inverse Temperature"] << 1.0; few["Volume"] << 1.0;
    energy::evaluate(few);
    CHECK(!few.has_evaluated("Specific Heat"));   // fewer than two bins: nothing to evaluate
  }
  // --- ALPS parameter-file conventions the reference's own inputs use (loop.ip, check/*, extras/*/*.ip):
  //     tasks in braces on top of the globals before them, ';' ',' and newline as separators, comments,
  //     quoted values with ';' inside, numeric expressions over other parameters
  {
    std::istringstream in(
        "LATTICE = \"chain lattice\"\nMODEL = \"spin\"\nlocal_S = 1/2; L = 4, Jxy = -1; Jz = -1  // ferromagnet\n"
        "ALGORITHM = \"loop; path integral\"\n# a comment { with = braces }\n"
        "{ T = 0.1 } { T = 0.2 }\n{ T = 1/L; ALGORITHM = \"loop; sse\" }\nSWEEPS = 2*512\n{ T = (1+1)/pi }\n");
    Parameters base; base.set("SWEEPS", 64);
    std::vector<Parameters> t = Parameters::parse_tasks(in, base);
    CHECK(t.size() == 4);
    CHECK(t[0].value_or_default<double>("T", 0) == 0.1 && t[1].value_or_default<double>("T", 0) == 0.2);
    CHECK(t[2].value_or_default<double>("T", 0) == 0.25 && t[2].get("ALGORITHM") == "loop; sse");
    CHECK(t[0].get("ALGORITHM") == "loop; path integral" && t[0].value_or_default<int>("L", 0) == 4);
    CHECK(t[0].value_or_default<double>("local_S", 0) == 0.5 && t[0].value_or_default<double>("Jz", 0) == -1);
    CHECK(t[0].value_or_default<int>("SWEEPS", 0) == 64 && t[3].value_or_default<int>("SWEEPS", 0) == 1024);   // globals apply to LATER tasks
    CHECK(std::abs(t[3].value_or_default<double>("T", 0) - 2 / 3.14159265358979323846) < 1e-15);
    CHECK(t[2].task_keys().size() == 2 && t[2].task_keys()[0] == "T");
    CHECK(t[0].value_or_default<std::string>("LATTICE", "") == "chain lattice");
    std::istringstream none("L = 6\nT = 0.5\n");
    CHECK(Parameters::parse_tasks(none, base).size() == 1);
    for (const char* bad : {"{ T = 1 ", "T = 1 }", "{ { T = 1 } }"}) {
      std::istringstream b(bad);
      bool threw = false;
      try { Parameters::parse_tasks(b); } catch (const std::invalid_argument&) { threw = true; }
      CHECK(threw);
    }
    Parameters q; q["T"] = "1/Lx"; q["N"] = "7/2"; q["A"] = "A+1";
    for (const char* k : {"T", "A"}) { bool threw = false; try { q.value_or_default<double>(k, 0); } catch (const std::invalid_argument&) { threw = true; } CHECK(threw); }
    { bool threw = false; try { q.value_or_default<int>("N", 0); } catch (const std::invalid_argument&) { threw = true; } CHECK(threw); }
    CHECK(q.value_or_default<double>("N", 0) == 3.5);
  }
  // --- a drop-in refuses what it does not implement (instead of quietly simulating S = 1/2 XXZ)
  {
    Parameters p; p["LATTICE"] = "chain lattice"; p.set("L", 4); p.set("Jxy", -1.0);
    lattice_helper lat(p);
    auto refused = [&](const char* k, const char* v) {
      Parameters q = p; q[k] = v;
      try { spinmodel_helper m(q, lat); } catch (const std::invalid_argument&) { return true; }
      return false;
    };
    CHECK(!refused("local_S", "1/2") && !refused("MODEL", "spin") && !refused("D", "0") && !refused("h", "0"));
    CHECK(refused("local_S", "1") && refused("local_S", "3/2") && refused("MODEL", "XYZ spin") && refused("h", "0.3"));
    CHECK(refused("D", "-0.2") && refused("Jx", "1") && refused("Jz1", "1") && refused("Gamma1", "-0.3") && refused("J'", "0.5"));
    CHECK(!refused("Jz0", "1") && refused("J0", "1") && refused("h0", "0.3") && !refused("h0", "0"));   // type 0 is every bond of a plain chain
    // type-dependent couplings on a lattice that has the types (extras/transmag, check/transmag-3): the
    // antiferromagnet in a STAGGERED transverse field has no sign problem, in a uniform one it has
    Parameters a; a["LATTICE"] = "alternating chain lattice"; a.set("L", 4);
    a.set("Jz0", 1.0); a.set("Jxy0", 1.0); a.set("Jz1", 1.0); a.set("Jxy1", 1.0); a.set("Gamma0", 0.3); a.set("Gamma1", -0.3);
    lattice_helper alt(a);
    CHECK(alt.vg().site_type[1] == 1 && alt.vg().site_type[2] == 0 && alt.vg().bond_type[0] == 0 && alt.vg().bond_type[1] == 1 && alt.vg().bond_type[3] == 1);
    spinmodel_helper ma(a, alt);
    CHECK(ma.uniform_site_weights() && std::abs(ma.site_weight() - 0.15) < 1e-15 && ma.has_site_weights());
    CHECK(std::abs(ma.graph_weight() - (4 * 0.5 + 4 * 0.15)) < 1e-12 && std::abs(ma.energy_offset() - (4 * 0.25 + 4 * 0.15)) < 1e-12);
    auto throws = [&](Parameters q) { try { spinmodel_helper m2(q, alt); } catch (const std::invalid_argument&) { return true; } return false; };
    Parameters u = a; u.set("Gamma1", 0.3);            // uniform field on the antiferromagnet
    CHECK(throws(u));
    Parameters f = a; f.set("Jxy0", -1.0); f.set("Jxy1", -1.0);   // ferromagnetic exchange: now the STAGGERED field is the problem
    CHECK(throws(f));
    f.set("Gamma1", 0.3);
    CHECK(!throws(f));
    Parameters d = a; d.set("Jxy1", 0.5); d.set("Jz1", 2.0); d.set("Gamma1", -0.1);   // dimerised, different fields
    spinmodel_helper md(d, alt);
    CHECK(!md.uniform_site_weights() && md.site_weight() == 0 && md.site_weights()[0] == 0.15 && md.site_weights()[1] == 0.05);
    CHECK(md.bond_weights()[4 * 0 + 0] == 0.5 && md.bond_weights()[4 * 1 + 0] == 0.25 && md.bond_weights()[4 * 1 + 2] == 0.75);   // Jxy = 1/2, Jz = 2: v0 = 1/4, v2 = 3/4
    Parameters o = a; o.set("Gamma1", 0.0);            // a field on one sublattice only: no constraint from the other
    spinmodel_helper mo(o, alt);
    CHECK(!mo.uniform_site_weights() && mo.has_site_weights() && mo.site_weights()[1] == 0);
    // a periodic direction of extent 2 is a double bond in ALPS (test/lattice.op: L = 2, W = 4 -> 16 bonds): refused
    for (const char* name : {"chain lattice", "square lattice", "simple cubic lattice"}) {
      Parameters two; two["LATTICE"] = name; two.set("L", 2); two.set("W", 4);
      bool th = false;
      try { lattice_helper l2(two); } catch (const std::invalid_argument&) { th = true; }
      CHECK(th);
    }
    { Parameters lad; lad["LATTICE"] = "ladder"; lad.set("L", 4); lattice_helper l(lad); CHECK(num_sites(l.vg()) == 8 && num_bonds(l.vg()) == 12); }
    Parameters tri; tri["LATTICE"] = "chain lattice"; tri.set("L", 3);   // odd ring, antiferromagnetic exchange: frustrated
    lattice_helper ring3(tri);
    { bool th = false; try { spinmodel_helper m3(tri, ring3); } catch (const std::invalid_argument&) { th = true; } CHECK(th); }
    tri.set("Jxy", -1.0);
    { spinmodel_helper m3(tri, ring3); CHECK(std::abs(m3.graph_weight() - 1.5) < 1e-12); }
    // LATTICE = "site" (check/site-*, extras/transmag): one spin in a transverse field, no bonds, no sign
    Parameters s; s["LATTICE"] = "site"; s.set("Gamma", 0.7);
    lattice_helper one(s);
    spinmodel_helper m(s, one);
    CHECK(num_sites(one.vg()) == 1 && num_bonds(one.vg()) == 0 && one.vg().dimension == 0);
    CHECK(std::abs(m.graph_weight() - 0.35) < 1e-15 && std::abs(m.energy_offset() - 0.35) < 1e-15);
  }
  std::cout << "host ok\n";
  return 0;
}
