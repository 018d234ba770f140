// loop.cpp -- minimal driver around looper::loop_worker (stand-in for loop.C:25-34 +
// alps::parapack::start): reads ALPS-style parameters from files or stdin ("KEY = value" statements, '{ ... }' blocks
// = one task each, expressions such as "T = 1/L": parameters.h), or the standalone kernel's flags -l/-t/-n
// (standalone/options.h:40-63), runs every task on the GPU one after the other and prints its observables.
// Usage: loop [-l L] [-t T] [-n sweeps] [--lattice "square lattice"]
//             [--checkpoint file] [--nranks P] [--dry-run] [params-file | -] ...
// The flags apply to every task, whatever their position on the command line.
// --checkpoint: resume from the file if it exists (path_integral.C:111-124 load), write it at the end
//               (several tasks: file.0, file.1, ...)
// --dry-run:    read and validate every task (lattice, model, schedule) without touching a GPU
// --nranks P:   one Markov chain over P GPUs (imaginary-time slabs, path_integral_mpi.C): the process forks
//               P - 1 children BEFORE touching CUDA, rank r drives GPU r, the NCCL id travels through a file
#include <cstring>
#include <fstream>
#include <iostream>
#include <sys/wait.h>
#include <unistd.h>
#include "loop_worker.h"

// measurements the reference would take and this path does not (correlations, structure factor, gap, ...:
// SURVEY section 2 rows marked OUT): say so instead of ignoring the request silently
static void warn_unmeasured(const looper::Parameters& p) {
  for (const auto& kv : p.items()) {
    const std::string& k = kv.first;
    if (k.compare(0, 7, "MEASURE") != 0 || k == "MEASURE[Stiffness]") continue;
    if (kv.second == "0" || kv.second == "false") continue;
    std::cerr << "warning: " << k << " is not measured by the accelerated path\n";
  }
}

// --dry-run: everything the worker's constructor checks before it creates the engine
static void check_task(const looper::Parameters& p) {
  warn_unmeasured(p);
  looper::lattice_helper lat(p);
  looper::spinmodel_helper model(p, lat);
  looper::temperature temp(p);
  looper::mc_steps mcs(p);
  const std::string alg = p.value_or_default("ALGORITHM", "loop; path integral");
  if (alg != "loop" && alg != "loop; path integral" && alg != "loop; sse")
    throw std::invalid_argument("unknown ALGORITHM '" + alg + "' (loop; path integral | loop; sse)");
  if (temp.annealing_steps() > mcs.thermalization()) throw std::invalid_argument("longer annealing steps than thermalization");
  if (p.defined("DISABLE_IMPROVED_ESTIMATOR")) throw std::invalid_argument("the accelerated path implements the improved estimators only");
  std::cout << "ok: " << alg << ", " << num_sites(lat.vg()) << " sites, " << num_bonds(lat.vg()) << " bonds, T = " << temp(mcs.thermalization())
            << ", graph weight " << model.graph_weight() << ", " << mcs.thermalization() << " + " << mcs.sweeps() << " sweeps\n";
}

// one task: a worker, its observables, the run, the report (rank 0)
static void run_task(const looper::Parameters& p, const looper::communicator& comm, const std::string& ckpt) {
  if (comm.rank() == 0) warn_unmeasured(p);
  looper::loop_worker w(comm, p);
  looper::observable_set obs;
  w.init_observables(p, obs);
  if (!ckpt.empty()) {
    std::ifstream in(ckpt, std::ios::binary);
    if (in) { w.load(in); obs.load(in); std::cout << "resumed at " << w.progress() << " of the run\n"; }
  }
  while (w.progress() < 1) w.run(obs);
  if (!ckpt.empty()) {
    std::ofstream out(ckpt, std::ios::binary | std::ios::trunc);
    w.save(out);
    obs.save(out);   // the binning state travels with the worker, like the scheduler's ObservableSet dump
  }
  if (comm.rank() != 0) return;   // every rank holds the same observables; rank 0 reports
  if (!obs.has("Temperature") || obs["Temperature"].count() == 0) {   // (e.g. a checkpoint of a finished run)
    std::cout << "no measurement was taken\n";
    return;
  }
  const double N = w.lat().volume(), beta = 1 / obs["Temperature"].mean();
  // the five lines of standalone/loop.C:186-195, from the looper-named observables
  std::cout << "Number of Clusters        = " << obs["Number of Clusters"].mean() << " +- " << obs["Number of Clusters"].error() << "\n"
            << "Energy Density            = " << obs["Energy Density"].mean() << " +- " << obs["Energy Density"].error() << "\n"
            << "Uniform Susceptibility    = " << beta * obs["Magnetization^2"].mean() / N << " +- " << beta * obs["Magnetization^2"].error() / N << "\n"
            << "Staggered Magnetization^2 = " << obs["Staggered Magnetization^2"].mean() << " +- " << obs["Staggered Magnetization^2"].error() << "\n"
            << "Staggered Susceptibility  = " << obs["Staggered Susceptibility"].mean() << " +- " << obs["Staggered Susceptibility"].error() << "\n";
  looper::energy::evaluate(obs);                 // alps::parapack evaluators (loop.C:29-33: --evaluate)
  looper::evaluate_susceptibility(obs);
  if (p.defined("VERBOSE")) obs.print(std::cout);
}

int main(int argc, char** argv) {
  looper::Parameters base, flags;
  base["LATTICE"] = "chain lattice";
  base.set("L", 8);
  base.set("T", 0.2);
  base.set("SWEEPS", 1u << 16);
  std::string ckpt;
  int nranks = 1;
  bool dry_run = false;
  looper::communicator comm;
  std::vector<pid_t> children;
  std::vector<looper::Parameters> tasks;
  int failed = 0;
  try {
    for (int i = 1; i < argc; ++i) {
      if (!std::strcmp(argv[i], "-l") && i + 1 < argc) flags["L"] = argv[++i];
      else if (!std::strcmp(argv[i], "-t") && i + 1 < argc) flags["T"] = argv[++i];
      else if (!std::strcmp(argv[i], "-n") && i + 1 < argc) flags["SWEEPS"] = argv[++i];
      else if (!std::strcmp(argv[i], "--lattice") && i + 1 < argc) flags["LATTICE"] = argv[++i];
      else if (!std::strcmp(argv[i], "--checkpoint") && i + 1 < argc) ckpt = argv[++i];
      else if (!std::strcmp(argv[i], "--nranks") && i + 1 < argc) nranks = std::atoi(argv[++i]);
      else if (!std::strcmp(argv[i], "--dry-run")) dry_run = true;
      else {
        std::vector<looper::Parameters> t;
        if (!std::strcmp(argv[i], "-")) t = looper::Parameters::parse_tasks(std::cin, base);
        else {
          std::ifstream f(argv[i]);
          if (!f) throw std::invalid_argument(std::string("cannot open ") + argv[i]);
          t = looper::Parameters::parse_tasks(f, base);
        }
        tasks.insert(tasks.end(), t.begin(), t.end());
      }
    }
    if (tasks.empty()) tasks.push_back(base);
    for (looper::Parameters& p : tasks)
      for (const auto& kv : flags.items()) p[kv.first] = kv.second;
    std::string id_base;
    if (dry_run) nranks = 1;
    if (nranks > 1) {
      if (!ckpt.empty()) throw std::invalid_argument("--checkpoint with --nranks is not supported");
      comm.size_ = nranks;
      id_base = "/tmp/lq_nccl_id." + std::to_string((long)getpid());
      for (size_t t = 0; t < tasks.size(); ++t)   // no stale id may be lying around when the children start
        std::remove((id_base + (tasks.size() > 1 ? "." + std::to_string(t) : "")).c_str());
      for (int r = 1; r < nranks; ++r) {   // fork before any CUDA / NCCL call
        const pid_t c = fork();
        if (c < 0) throw std::runtime_error("fork failed");
        if (c == 0) { comm.rank_ = r; children.clear(); break; }
        children.push_back(c);
      }
    }
    for (size_t t = 0; t < tasks.size(); ++t) {
      const looper::Parameters& p = tasks[t];
      const std::string suffix = tasks.size() > 1 ? "." + std::to_string(t) : "";
      if (nranks > 1) comm.id_file = id_base + suffix;
      if (tasks.size() > 1 && comm.rank() == 0) {   // which task: the block's own statements
        std::cout << "[task " << t + 1 << " of " << tasks.size() << "]";
        for (const std::string& k : p.task_keys()) std::cout << ' ' << k << " = " << p.get(k) << ';';
        std::cout << "\n";
      }
      try {
        if (dry_run) check_task(p);
        else run_task(p, comm, ckpt.empty() ? ckpt : ckpt + suffix);
      } catch (const std::exception& e) {
        if (tasks.size() == 1) throw;
        std::cerr << "error in task " << t + 1 << ": " << e.what() << "\n";   // the other tasks still run
        ++failed;
      }
    }
    if (comm.rank() != 0) return failed ? 1 : 0;
    for (pid_t c : children) { int st = 0; waitpid(c, &st, 0); }
    for (size_t t = 0; nranks > 1 && t < tasks.size(); ++t)   // (only now: every rank has read every id)
      std::remove((id_base + (tasks.size() > 1 ? "." + std::to_string(t) : "")).c_str());
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << "\n";
    return 1;
  }
  return failed ? 1 : 0;
}
