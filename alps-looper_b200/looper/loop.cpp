// loop.cpp -- minimal driver around looper::loop_worker (stand-in for loop.C:25-34 +
// alps::parapack::start): reads "KEY = value" parameters from a file or stdin, or the standalone
// kernel's flags -l/-t/-n (standalone/options.h:40-63), runs the worker on the GPU and prints the
// observables.  Usage: loop [-l L] [-t T] [-n sweeps] [--lattice "square lattice"]
//                    [--checkpoint file] [--nranks P] [params-file]
// --checkpoint: resume from the file if it exists (path_integral.C:111-124 load), write it at the end
// --nranks P:   one Markov chain over P GPUs (imaginary-time slabs, path_integral_mpi.C): the process forks
//               P - 1 children BEFORE touching CUDA, rank r drives GPU r, the NCCL id travels through a file
#include <cstring>
#include <fstream>
#include <iostream>
#include <sys/wait.h>
#include <unistd.h>
#include "loop_worker.h"

int main(int argc, char** argv) {
  looper::Parameters p;
  p["LATTICE"] = "chain lattice";
  p.set("L", 8);
  p.set("T", 0.2);
  p.set("SWEEPS", 1u << 16);
  std::string ckpt;
  int nranks = 1;
  looper::communicator comm;
  std::vector<pid_t> children;
  try {
    for (int i = 1; i < argc; ++i) {
      if (!std::strcmp(argv[i], "-l") && i + 1 < argc) p["L"] = argv[++i];
      else if (!std::strcmp(argv[i], "-t") && i + 1 < argc) p["T"] = argv[++i];
      else if (!std::strcmp(argv[i], "-n") && i + 1 < argc) p["SWEEPS"] = argv[++i];
      else if (!std::strcmp(argv[i], "--lattice") && i + 1 < argc) p["LATTICE"] = argv[++i];
      else if (!std::strcmp(argv[i], "--checkpoint") && i + 1 < argc) ckpt = argv[++i];
      else if (!std::strcmp(argv[i], "--nranks") && i + 1 < argc) nranks = std::atoi(argv[++i]);
      else if (!std::strcmp(argv[i], "-")) p.parse(std::cin);
      else { std::ifstream f(argv[i]); if (!f) throw std::invalid_argument(std::string("cannot open ") + argv[i]); p.parse(f); }
    }
    if (nranks > 1) {
      if (!ckpt.empty()) throw std::invalid_argument("--checkpoint with --nranks is not supported");
      comm.size_ = nranks;
      comm.id_file = "/tmp/lq_nccl_id." + std::to_string((long)getpid());
      std::remove(comm.id_file.c_str());
      for (int r = 1; r < nranks; ++r) {   // fork before any CUDA / NCCL call
        const pid_t c = fork();
        if (c < 0) throw std::runtime_error("fork failed");
        if (c == 0) { comm.rank_ = r; children.clear(); break; }
        children.push_back(c);
      }
    }
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
    if (comm.rank() != 0) return 0;   // every rank holds the same observables; rank 0 reports
    for (pid_t c : children) { int st = 0; waitpid(c, &st, 0); }
    if (nranks > 1) std::remove(comm.id_file.c_str());
    if (!obs.has("Temperature") || obs["Temperature"].count() == 0) {   // (e.g. a checkpoint of a finished run)
      std::cout << "no measurement was taken\n";
      return 0;
    }
    const double N = w.lat().volume(), beta = 1 / obs["Temperature"].mean();
    // the five lines of standalone/loop.C:186-195, from the looper-named observables
    std::cout << "Number of Clusters        = " << obs["Number of Clusters"].mean() << " +- " << obs["Number of Clusters"].error() << "\n"
              << "Energy Density            = " << obs["Energy Density"].mean() << " +- " << obs["Energy Density"].error() << "\n"
              << "Uniform Susceptibility    = " << beta * obs["Magnetization^2"].mean() / N << " +- " << beta * obs["Magnetization^2"].error() / N << "\n"
              << "Staggered Magnetization^2 = " << obs["Staggered Magnetization^2"].mean() << " +- " << obs["Staggered Magnetization^2"].error() << "\n"
              << "Staggered Susceptibility  = " << obs["Staggered Susceptibility"].mean() << " +- " << obs["Staggered Susceptibility"].error() << "\n";
    if (p.defined("VERBOSE")) obs.print(std::cout);
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
