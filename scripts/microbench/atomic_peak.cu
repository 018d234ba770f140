// atomic_peak.cu -- measured peak throughput of the global atomics the engine relies on (B200):
//   RED.64  (atomicAdd, result unused: k_estimate's cluster sums)
//   CAS.32  (atomicCAS: the hook of the lock-free union-find, k_union_global)
//   EXCH.64 (atomicExch read-and-clear: k_collect)
// to sequential and to random addresses of a table much larger than L2 (1 GiB) and of an L2-resident
// one (32 MiB).  Prints one JSON object; north_star: "atomic throughput against B200 peak".
// Build: make -C scripts/microbench   Run: scripts/microbench/atomic_peak
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int OP, bool RANDOM>
__global__ void k_atomic(unsigned long long* t64, uint32_t* t32, size_t n, int iters, unsigned long long* sink) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  unsigned long long acc = 0;
  for (int it = 0; it < iters; ++it) {
    const size_t lin = (tid + (size_t)it * stride) % n;
    const size_t i = RANDOM ? ((size_t)mix((uint32_t)lin) * 2654435761ull + (size_t)mix((uint32_t)(lin >> 32) + 17u)) % n : lin;
    if (OP == 0) atomicAdd(t64 + i, 1ull);                       // RED.E.ADD.64
    else if (OP == 1) acc += atomicCAS(t32 + i, 0xffffffffu, 1u);  // ATOMG.CAS (never succeeds: pure traffic)
    else acc += atomicExch(t64 + i, 0ull);                        // ATOMG.EXCH.64
  }
  if (acc == 0x123456789ull) *sink = acc;
}

template <int OP, bool RANDOM>
double run(unsigned long long* t64, uint32_t* t32, size_t n, unsigned long long* sink) {
  const int blocks = 148 * 16, threads = 256, iters = 64;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k_atomic<OP, RANDOM><<<blocks, threads>>>(t64, t32, n, 8, sink);   // warm-up
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(a);
    k_atomic<OP, RANDOM><<<blocks, threads>>>(t64, t32, n, iters, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return (double)blocks * threads * iters / (best * 1e-3);
}

int main() {
  const size_t big = (size_t)1 << 27, small = (size_t)1 << 22;   // 1 GiB / 32 MiB of 8-byte slots
  unsigned long long *t64, *sink;
  uint32_t* t32;
  if (cudaMalloc(&t64, big * 8) != cudaSuccess || cudaMalloc(&t32, big * 4) != cudaSuccess || cudaMalloc(&sink, 8) != cudaSuccess) {
    printf("{\"error\": \"cudaMalloc failed\"}\n");
    return 1;
  }
  cudaMemset(t64, 0, big * 8);
  cudaMemset(t32, 0, big * 4);
  printf("{");
  const char* names[3] = {"red64", "cas32", "exch64"};
  for (int pass = 0; pass < 2; ++pass) {
    const size_t n = pass ? small : big;
    const char* tag = pass ? "l2_resident_32MiB" : "hbm_1GiB";
    double v[6] = {run<0, false>(t64, t32, n, sink), run<0, true>(t64, t32, n, sink), run<1, false>(t64, t32, n, sink),
                   run<1, true>(t64, t32, n, sink),  run<2, false>(t64, t32, n, sink), run<2, true>(t64, t32, n, sink)};
    for (int k = 0; k < 3; ++k)
      printf("%s\"%s_%s_sequential_per_s\": %.4g, \"%s_%s_random_per_s\": %.4g", (pass || k) ? ", " : "", names[k], tag, v[2 * k],
             names[k], tag, v[2 * k + 1]);
  }
  printf("}\n");
  return 0;
}
