// mpi.h -- thread-backed stand-in for the handful of MPI calls standalone/loop_mpi.C and
// standalone/parallel.h make (Init/Finalize/Comm_size/Comm_rank/Barrier/Send/Recv), so that the
// UNMODIFIED reference multi-rank kernel runs as P threads of one process and can be diffed
// against standalone/loop_mpi.op-P.  TEST INFRASTRUCTURE ONLY (oracle/_ref build).
#ifndef ORACLE_MPI_SHIM_H
#define ORACLE_MPI_SHIM_H
#include <sys/uio.h>
#include <unistd.h>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

typedef int MPI_Comm;
typedef int MPI_Datatype;
struct MPI_Status { int MPI_SOURCE, MPI_TAG, MPI_ERROR; };
#define MPI_COMM_WORLD 0
#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_SUCCESS 0

namespace mpi_shim {
struct world {
  int size = 1;
  std::mutex m;
  std::condition_variable cv;
  std::map<std::tuple<int, int, int>, std::deque<std::vector<char> > > box;  // (src,dst,tag)
  int bar_count = 0, bar_gen = 0;
};
inline world& W() { static world w; return w; }
inline int& rank_ref() { static thread_local int r = 0; return r; }
}  // namespace mpi_shim

inline int MPI_Init(int*, char***) { return 0; }
inline int MPI_Finalize() { return 0; }
inline int MPI_Comm_size(MPI_Comm, int* n) { *n = mpi_shim::W().size; return 0; }
inline int MPI_Comm_rank(MPI_Comm, int* r) { *r = mpi_shim::rank_ref(); return 0; }
inline int MPI_Barrier(MPI_Comm) {
  mpi_shim::world& w = mpi_shim::W();
  std::unique_lock<std::mutex> lk(w.m);
  int gen = w.bar_gen;
  if (++w.bar_count == w.size) { w.bar_count = 0; ++w.bar_gen; w.cv.notify_all(); }
  else w.cv.wait(lk, [&] { return w.bar_gen != gen; });
  return 0;
}
inline int MPI_Send(const void* buf, int count, MPI_Datatype dt, int dest, int tag, MPI_Comm) {
  mpi_shim::world& w = mpi_shim::W();
  std::vector<char> msg((size_t)count * dt);
  // standalone/parallel.h:240 sends 2N estimates out of a vector that may hold fewer: it reads past
  // the end of its buffer, which a real MPI survives (the receiver ignores the tail) but a memcpy may
  // not when the tail crosses an unmapped page.  process_vm_readv on the own process copies what is
  // readable and stops at the first fault instead of raising SIGSEGV.
  if (!msg.empty()) {
    size_t done = 0;
    while (done < msg.size()) {
      struct iovec lo = {msg.data() + done, msg.size() - done};
      struct iovec re = {(char*)const_cast<void*>(buf) + done, msg.size() - done};
      const ssize_t r = process_vm_readv(getpid(), &lo, 1, &re, 1, 0);
      if (r <= 0) break;
      done += (size_t)r;
    }
    if (done == 0) std::memcpy(msg.data(), buf, msg.size() < 4096 ? msg.size() : 4096);
  }
  { std::lock_guard<std::mutex> lk(w.m);
    w.box[std::make_tuple(mpi_shim::rank_ref(), dest, tag)].push_back(std::move(msg)); }
  w.cv.notify_all();
  return 0;
}
inline int MPI_Recv(void* buf, int count, MPI_Datatype dt, int source, int tag, MPI_Comm,
                    MPI_Status* st) {
  mpi_shim::world& w = mpi_shim::W();
  std::unique_lock<std::mutex> lk(w.m);
  auto key = std::make_tuple(source, mpi_shim::rank_ref(), tag);
  w.cv.wait(lk, [&] { auto it = w.box.find(key); return it != w.box.end() && !it->second.empty(); });
  std::vector<char> msg = std::move(w.box[key].front());
  w.box[key].pop_front();
  size_t n = std::min(msg.size(), (size_t)count * dt);
  if (n) std::memcpy(buf, msg.data(), n);
  if (st) { st->MPI_SOURCE = source; st->MPI_TAG = tag; st->MPI_ERROR = 0; }
  return 0;
}
#endif
