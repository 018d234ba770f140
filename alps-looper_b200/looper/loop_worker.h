// loop_worker.h -- the reference's worker interface ("loop; path integral",
// path_integral.C:57-125,202-351,831-862) with its dispatch() body replaced by calls into the C ABI
// of include/lq.h.  Same member names, argument meaning and error behaviour (exceptions for bad
// parameters), minus the ALPS base classes: Parameters / observable_set are the stand-ins of
// parameters.h / measurement.h.
#pragma once
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>
#include <istream>
#include <ostream>
#include <stdexcept>
#include <string>
#include <vector>
#include "../../include/lq.h"
#include "lattice.h"
#include "measurement.h"
#include "model.h"
#include "montecarlo.h"
#include "parameters.h"

namespace looper {

// Stand-in for the boost::mpi::communicator the parallel worker receives (path_integral_mpi.C:75):
// rank / size of the imaginary-time slabs, one process and one GPU per rank.  The ranks only have to
// agree on 128 bytes once -- the NCCL unique id, created by rank 0 and passed through `id_file` (a
// path all ranks can read; an MPI host would MPI_Bcast it instead); every collective of a step is
// then issued by the engine itself (lq_comm_init, include/lq.h).
struct communicator {
  int rank_ = 0, size_ = 1;
  std::string id_file;
  int rank() const { return rank_; }
  int size() const { return size_; }
};

class loop_worker {
public:
  typedef double weight_parameter_type;

  // parallel flavour (path_integral_mpi.C:75, PARAPACK_REGISTER_PARALLEL_ALGORITHM :1012)
  loop_worker(const communicator& c, const Parameters& p) : loop_worker(p, c) {}

  explicit loop_worker(const Parameters& p, const communicator& c = communicator())
      : lattice(p), model(p, lattice), temp(p), mcs(p), comm_(c) {
    // ALGORITHM (loop.C:25-34; PARAPACK_REGISTER_ALGORITHM path_integral.C:873 "loop; path integral",
    // sse.C:411 "loop; sse"; plain "loop" is the path integral, loop.C).  Both workers sample the same
    // ensemble -- an SSE string is the time-ordered operator list of a world-line configuration -- and
    // the engine reports the estimators of the representation asked for (lq_options.representation).
    const std::string alg = p.value_or_default("ALGORITHM", "loop; path integral");
    if (alg == "loop" || alg == "loop; path integral") sse_ = false;
    else if (alg == "loop; sse") sse_ = true;
    else throw std::invalid_argument("unknown ALGORITHM '" + alg + "' (loop; path integral | loop; sse)");
    if (temp.annealing_steps() > mcs.thermalization())
      throw std::invalid_argument("longer annealing steps than thermalization");  // path_integral.C:213
    enable_improved_estimator = !p.defined("DISABLE_IMPROVED_ESTIMATOR");
    if (!enable_improved_estimator)
      throw std::invalid_argument("the accelerated path implements the improved estimators only");
    const virtual_graph& g = lattice.vg();
    lq_lattice L;
    L.num_sites = num_sites(g);
    L.num_bonds = num_bonds(g);
    static const int32_t no_bond = 0;   // LATTICE = "site": no bonds, but the C ABI wants non-null arrays
    L.src = L.num_bonds ? g.src.data() : &no_bond;
    L.dst = L.num_bonds ? g.dst.data() : &no_bond;
    L.gauge = lattice.is_bipartite() ? g.gauge.data() : nullptr;
    for (int k = 0; k < 3; ++k) L.dims[k] = g.dims[k];
    // MEASURE[Stiffness] (looper/stiffness.h): winding numbers need the relative bond vectors
    measure_stiffness = p.defined("MEASURE[Stiffness]") && g.dimension > 0;
    L.vector_dim = measure_stiffness ? g.dimension : 0;
    L.bond_vectors = measure_stiffness ? g.bond_vector_relative.data() : nullptr;
    lq_model M;
    static const double no_weight[4] = {0, 0, 0, 0};
    M.bond_weights = L.num_bonds ? model.bond_weights().data() : no_weight;
    for (int k = 0; k < 4; ++k) M.uniform_weights[k] = 0;
    M.energy_offset = model.energy_offset();
    // (the same |Hx|/2 on every site goes in as one number; per-site values only when they differ)
    M.site_weights = model.uniform_site_weights() ? nullptr : model.site_weights().data();
    M.uniform_site_weight = model.site_weight();
    lq_options o = lq_options();
    o.seed = p.value_or_default<unsigned long long>("WORKER_SEED", p.value_or_default<unsigned long long>("SEED", 29833ull));
    seed_ = o.seed;
    o.device = p.value_or_default<int>("DEVICE", comm_.size() > 1 ? comm_.rank() : 0);
    o.rank = comm_.rank();
    o.nranks = comm_.size();
    o.tile_sites = p.value_or_default<int>("TILE_SITES", 0);
    tile_sites_ = o.tile_sites;
    o.reserve = p.value_or_default<double>("RESERVE_OPERATORS_FACTOR", 0.0);  // cf. RESERVE_OPERATORS (:241)
    o.cluster_reserve = p.value_or_default<double>("RESERVE_ESTIMATES_FACTOR", 0.0);
    o.flags = p.defined("ENABLE_TIMER") ? 1 : 0;
    o.representation = sse_ ? LQ_REPR_SSE : LQ_REPR_PATH_INTEGRAL;
    {
      // how the ranks share the configuration: "time" = imaginary-time slabs (path_integral_mpi.C:231-232),
      // "space" = strips of tiles (the lattice sharing of looper/lattice.h:692-787 taken across GPUs)
      const std::string cut = p.value_or_default<std::string>("PARTITION", "time");
      if (cut != "time" && cut != "space") throw std::invalid_argument("PARTITION must be \"time\" or \"space\"");
      o.cut = cut == "space" ? LQ_CUT_SPACE : LQ_CUT_TIME;
    }
    beta_ = 1.0 / temp(0);
    check(lq_create(&h_, &L, &M, beta_, &o));
    if (comm_.size() > 1) connect();
  }
  ~loop_worker() { lq_destroy(h_); }
  loop_worker(const loop_worker&) = delete;
  loop_worker& operator=(const loop_worker&) = delete;

  void init_observables(const Parameters&, observable_set& obs) {  // path_integral.C:307-321
    obs["Temperature"]; obs["Inverse Temperature"]; obs["Volume"]; obs["Number of Sites"];
    obs["Number of Clusters"];
    energy::init_observables(obs);
  }
  bool is_thermalized() const { return mcs.is_thermalized(); }
  double progress() const { return mcs.progress(); }

  // one Monte Carlo step (path_integral.C:323-351): the update runs on the GPU, the observables
  // of the step are appended on the host once thermalised
  void run(observable_set& obs) {
    if (!mcs.can_work()) return;
    const double b = 1.0 / temp(mcs());
    if (b != beta_) { check(lq_set_beta(h_, b)); beta_ = b; }
    lq_collector coll;
    check(lq_sweep(h_, &coll));
    const bool measure = mcs.is_thermalized();   // decided BEFORE the increment: the sweep that completes the
    ++mcs;                                       // thermalisation is not measured (standalone/loop.C:173)
    if (!measure) return;
    const double vol = lattice.volume();
    obs["Temperature"] << 1 / beta_;            // path_integral.C:832-836
    obs["Inverse Temperature"] << beta_;
    obs["Volume"] << vol;
    obs["Number of Sites"] << double(num_sites(lattice.rg()));
    obs["Number of Clusters"] << coll.nc;
    energy::commit(obs, coll, beta_, vol);
    susceptibility::commit(obs, coll, beta_, vol, lattice.is_bipartite(), sse_);
    if (model.has_site_weights()) transverse_magnetization::commit(obs, coll, vol);
    if (measure_stiffness) stiffness::commit(obs, coll, beta_, lattice.vg().dimension);
    last_ = coll;
  }

  // exchange Monte Carlo hooks (path_integral.C:98-109)
  void set_beta(double beta) { temp.set_beta(beta); }
  weight_parameter_type weight_parameter() const {
    long long n = 0;
    check(lq_get_state(h_, nullptr, nullptr, (int64_t*)&n));
    return double(n);
  }
  static double log_weight(double gw, double beta) { return std::log(beta) * gw; }

  // checkpoint payload (path_integral.C:111-124): mcs, spins, operators {type_, loc_, time_}
  void save(unsigned& mcs_out, std::vector<int32_t>& spins, std::vector<lq_op>& ops) const {
    int64_t n = 0;
    check(lq_get_state(h_, nullptr, nullptr, &n));
    spins.resize(num_sites(lattice.vg()));
    ops.resize(size_t(n));
    check(lq_get_state(h_, spins.data(), ops.data(), &n));
    mcs_out = mcs();
  }
  void load(unsigned mcs_in, const std::vector<int32_t>& spins, const std::vector<lq_op>& ops) {
    check(lq_set_state(h_, spins.data(), ops.data(), int64_t(ops.size())));
    mcs.set(mcs_in);
  }

  // The same payload as a byte stream, in the reference's field order (path_integral.C:111-124:
  // `dp << mcs << spins << operators`; operator.h:107,138: type_, loc_, time_ per operator; vectors
  // as a 32-bit count followed by the elements).  Native little-endian behind an 8-byte tag -- the
  // XDR container of alps::ODump itself is ALPS's, not reproduced here.
  void save(std::ostream& os) const {
    unsigned m = 0;
    std::vector<int32_t> spins;
    std::vector<lq_op> ops;
    save(m, spins, ops);
    os.write("LQCKPT02", 8);
    // what the chain depends on besides the configuration: seed, tiling (the Philox keys use the engine's
    // internal numbering) and temperature -- validated on load, a resume with other settings would silently
    // continue a different chain
    lq_info info;
    check(lq_get_info(h_, &info));
    put(os, uint64_t(seed_));
    put(os, int32_t(tile_sites_));
    put(os, int32_t(info.num_tiles));
    put(os, double(beta_));
    put(os, uint32_t(lq_get_step(h_)));    // the generator state: Philox step counter (include/lq.h)
    put(os, uint32_t(m));
    put(os, uint32_t(spins.size()));
    os.write(reinterpret_cast<const char*>(spins.data()), std::streamsize(spins.size() * sizeof(int32_t)));
    put(os, uint32_t(ops.size() >> 32));   // operator strings beyond 2^32 entries: count as two words
    put(os, uint32_t(ops.size()));
    for (const lq_op& o : ops) { put(os, int32_t(o.type)); put(os, int32_t(o.loc)); put(os, double(o.time)); }
    if (!os) throw std::runtime_error("checkpoint: write failed");
  }
  void load(std::istream& is) {
    char tag[8];
    is.read(tag, 8);
    if (!is || std::memcmp(tag, "LQCKPT02", 8) != 0) throw std::runtime_error("checkpoint: bad tag");
    lq_info info;
    check(lq_get_info(h_, &info));
    const uint64_t seed = get<uint64_t>(is);
    const int32_t tile = get<int32_t>(is), ntiles = get<int32_t>(is);
    const double beta = get<double>(is);
    if (seed != seed_ || tile != tile_sites_ || ntiles != info.num_tiles)
      throw std::runtime_error("checkpoint: seed or tiling differ from this run (the chain would not continue)");
    const uint32_t step = get<uint32_t>(is);
    const uint32_t m = get<uint32_t>(is), ns = get<uint32_t>(is);
    if (ns != uint32_t(num_sites(lattice.vg()))) throw std::runtime_error("checkpoint: lattice size differs");
    std::vector<int32_t> spins(ns);
    is.read(reinterpret_cast<char*>(spins.data()), std::streamsize(ns * sizeof(int32_t)));
    const uint64_t hi = get<uint32_t>(is), lo = get<uint32_t>(is);
    std::vector<lq_op> ops(size_t((hi << 32) | lo));
    for (lq_op& o : ops) { o.type = get<int32_t>(is); o.loc = get<int32_t>(is); o.time = get<double>(is); }
    if (!is) throw std::runtime_error("checkpoint: truncated");
    // the pages are sized for the temperature: go to the checkpoint's before the operators come in
    // (a run that was annealing starts hot, with pages too small for a late configuration)
    if (beta != beta_) { check(lq_set_beta(h_, beta)); beta_ = beta; }
    load(m, spins, ops);
    check(lq_set_step(h_, step));
  }

  const lq_collector& last_collector() const { return last_; }
  lq_handle handle() const { return h_; }
  const lattice_helper& lat() const { return lattice; }

private:
  // NCCL rendezvous through a file: rank 0 writes the unique id (tmp + rename), the others wait for it
  void connect() {
    unsigned char id[LQ_NCCL_ID_BYTES];
    if (comm_.id_file.empty()) throw std::invalid_argument("communicator without an id_file");
    if (comm_.rank() == 0) {
      check(lq_comm_unique_id(id));
      const std::string tmp = comm_.id_file + ".tmp";
      { std::ofstream f(tmp, std::ios::binary | std::ios::trunc); f.write(reinterpret_cast<const char*>(id), sizeof id); }
      if (std::rename(tmp.c_str(), comm_.id_file.c_str()) != 0) throw std::runtime_error("cannot publish the NCCL id");
    } else {
      bool ok = false;
      for (int tries = 0; tries < 1200 && !ok; ++tries) {   // up to two minutes
        std::ifstream f(comm_.id_file, std::ios::binary);
        if (f && f.read(reinterpret_cast<char*>(id), sizeof id) && f.gcount() == (std::streamsize)sizeof id) ok = true;
        else std::this_thread::sleep_for(std::chrono::milliseconds(100));
      }
      if (!ok) throw std::runtime_error("timed out waiting for the NCCL id of rank 0");
    }
    check(lq_comm_init(h_, id, comm_.rank(), comm_.size()));
  }
  template <class T> static void put(std::ostream& os, T v) { os.write(reinterpret_cast<const char*>(&v), sizeof v); }
  template <class T> static T get(std::istream& is) { T v = T(); is.read(reinterpret_cast<char*>(&v), sizeof v); return v; }
  static void check(int rc) {
    if (rc != LQ_OK) throw std::runtime_error(std::string("lq: ") + lq_last_error());
  }
  lattice_helper lattice;
  spinmodel_helper model;
  temperature temp;
  mc_steps mcs;
  bool enable_improved_estimator = true;
  bool measure_stiffness = false;
  bool sse_ = false;
  communicator comm_;
  unsigned long long seed_ = 0;
  int tile_sites_ = 0;
  double beta_ = 1;
  lq_handle h_ = nullptr;
  lq_collector last_ = lq_collector();
};

}  // namespace looper
