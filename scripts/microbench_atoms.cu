// Microbenchmark: shared-memory atomicMax (ATOMS.MAX) throughput on B200, one 1024-thread CTA per SM.
// Patterns: conflict-free (lane i -> bank i), random cells in a 2752-word plane, runs of R equal cells
// across adjacent lanes.  Prints cycles per warp-level ATOMS instruction (per SM).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_atoms microbench_atoms.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(1024, 1) k_atoms(const uint32_t* __restrict__ idx, int iters, int n_idx,
                                                   unsigned long long* cycles_out, uint32_t* sink) {
  __shared__ uint32_t plane[4 * 2752];
  for (int i = threadIdx.x; i < 4 * 2752; i += blockDim.x) plane[i] = 0;
  __syncthreads();
  uint32_t my[16];
  for (int j = 0; j < 16; ++j) my[j] = idx[(threadIdx.x + j * 1024) % n_idx];
  __syncthreads();
  unsigned long long t0 = clock64();
  uint32_t v = threadIdx.x * 2654435761u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v = v * 1664525u + 1013904223u;
      atomicMax(&plane[my[j]], v);
    }
  }
  __syncthreads();
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
  if (v == 12345u) sink[0] = plane[threadIdx.x];
}

// Same traffic, but warp-aggregated first: lanes that target the same cell elect a leader
// (__match_any_sync), reduce their keys with __reduce_max_sync and issue ONE atomic per distinct cell.
__global__ void __launch_bounds__(1024, 1) k_atoms_agg(const uint32_t* __restrict__ idx, int iters, int n_idx,
                                                       unsigned long long* cycles_out, uint32_t* sink) {
  __shared__ uint32_t plane[4 * 2752];
  for (int i = threadIdx.x; i < 4 * 2752; i += blockDim.x) plane[i] = 0;
  __syncthreads();
  uint32_t my[16];
  for (int j = 0; j < 16; ++j) my[j] = idx[(threadIdx.x + j * 1024) % n_idx];
  __syncthreads();
  unsigned long long t0 = clock64();
  uint32_t v = threadIdx.x * 2654435761u;
  const int lane = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      v = v * 1664525u + 1013904223u;
      unsigned peers = __match_any_sync(0xFFFFFFFFu, my[j]);
      uint32_t m = __reduce_max_sync(peers, v);
      if (lane == __ffs(peers) - 1) atomicMax(&plane[my[j]], m);
    }
  }
  __syncthreads();
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
  if (v == 12345u) sink[0] = plane[threadIdx.x];
}

static bool g_agg = false;
static void run(const char* name, const uint32_t* h_idx, int n_idx, int iters) {
  uint32_t* d_idx; unsigned long long* d_cyc; uint32_t* d_sink;
  cudaMalloc(&d_idx, n_idx * 4); cudaMalloc(&d_cyc, 148 * 8); cudaMalloc(&d_sink, 4);
  cudaMemcpy(d_idx, h_idx, n_idx * 4, cudaMemcpyHostToDevice);
  for (int rep = 0; rep < 2; ++rep) {
    if (g_agg) k_atoms_agg<<<148, 1024>>>(d_idx, iters, n_idx, d_cyc, d_sink);
    else k_atoms<<<148, 1024>>>(d_idx, iters, n_idx, d_cyc, d_sink);
  }
  cudaDeviceSynchronize();
  unsigned long long h[148]; cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  double warp_instr = 32.0 * 16 * iters;   // per SM
  printf("%-28s %8.2f cycles per warp-step (per SM), %6.3f lane-atomics/cycle/SM  [%s]\n", name, avg / warp_instr,
         32.0 * warp_instr / avg, cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_idx); cudaFree(d_cyc); cudaFree(d_sink);
}

static void all_patterns();
int main() {
  printf("== plain ATOMS.MAX (what k_fused issues)\n");
  g_agg = false; all_patterns();
  printf("== __match_any_sync + __reduce_max_sync, one atomic per distinct cell in the warp\n");
  g_agg = true; all_patterns();
  return 0;
}
static void all_patterns() {
  const int N = 16 * 1024, iters = 200;
  static uint32_t idx[N];
  for (int i = 0; i < N; ++i) idx[i] = (i % 32) + 32 * ((i / 32) % 80);         // lane -> own bank
  run("conflict-free", idx, N, iters);
  uint32_t s = 1; for (int i = 0; i < N; ++i) { s = s * 1103515245u + 12345u; idx[i] = (s >> 8) % 2752; }
  run("random cell (2752)", idx, N, iters);
  for (int R = 2; R <= 32; R *= 2) {
    s = 7; uint32_t cur = 0;
    for (int i = 0; i < N; ++i) { if (i % R == 0) { s = s * 1103515245u + 12345u; cur = (s >> 8) % 2752; } idx[i] = cur; }
    char nm[64]; snprintf(nm, sizeof nm, "runs of %d equal lanes", R);
    run(nm, idx, N, iters);
  }
  for (int i = 0; i < N; ++i) idx[i] = 5;
  run("single address", idx, N, iters);
}
