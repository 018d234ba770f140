// loop_mpi_threads.cpp -- runs the UNMODIFIED standalone/loop_mpi.C as P threads over the
// thread-backed mpi.h in this directory:  _ref/loop_mpi P [loop_mpi options].
// The reference file is included from where it lies (path given by -DREF_LOOP_MPI).
// TEST INFRASTRUCTURE ONLY (oracle/_ref build).
#include <algorithm>
#include <cstdlib>
#include <thread>
#include <vector>
#include "mpi.h"
// `int main(int argc, char* argv[]) {...}` in the reference has no return statement, which is
// fine for main() but undefined behaviour for a renamed ordinary int function; so rename it into
// a *void* function:  int loop_dummy(); static void loop_main(int argc, char* argv[]) {...}
#define main(A, B) loop_dummy(); static void loop_main(A, B)
#include REF_LOOP_MPI
#undef main
int main(int argc, char** argv) {
  int P = (argc > 1) ? std::atoi(argv[1]) : 1;
  if (P < 1) P = 1;
  mpi_shim::W().size = P;
  std::vector<std::thread> th;
  for (int r = 0; r < P; ++r)
    th.emplace_back([=]() {
      mpi_shim::rank_ref() = r;
      std::vector<char*> av;
      av.push_back(argv[0]);
      for (int i = 2; i < argc; ++i) av.push_back(argv[i]);
      int ac = int(av.size());
      char** avp = av.data();
      loop_main(ac, avp);
    });
  for (auto& t : th) t.join();
  return 0;
}
