// Microbenchmark (round 2): how fast does TMA gather boxes of 16-byte cells out of an NHWC fp32 map [n,G,G,64]?
// One elected thread per CTA issues `iters` boxes {4 ch, BW cols, BH rows} into shared memory (double-buffered,
// mbarrier complete_tx), 1 or 2 CTAs per SM, windows spread over a map set larger than L2 or kept L2-hot.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_tma microbench_tma.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned saddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 2) k_tma(const __grid_constant__ CUtensorMap tm, int iters, int box_bytes, int n_maps, int hot, int depth,
                                                unsigned long long* cycles, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar[32];
  if (threadIdx.x == 0) {
    for (int i = 0; i < 32; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(saddr(&bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  unsigned long long t0 = clock64();
  float acc = 0.f;
  if (threadIdx.x == 0) {
    unsigned phases = 0;
    for (int it = 0; it < iters + depth - 1; ++it) {
      const int s = it % depth;
      if (it < iters) {
        const int m = hot ? (blockIdx.x % n_maps) : ((blockIdx.x * 7 + it * 131) % n_maps);
        const int u = hot ? 10 + (blockIdx.x * 37) % 100 : 10 + (it * 37) % 150, v = hot ? 10 + (blockIdx.x * 53) % 100 : 10 + (it * 53) % 150, c = 4 * ((blockIdx.x + it) & 15);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(saddr(&bar[s])), "r"(box_bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];\n"
                     ::"r"(saddr(smem + s * box_bytes)), "l"(&tm), "r"(c), "r"(v), "r"(u), "r"(m), "r"(saddr(&bar[s])) : "memory");
      }
      if (it >= depth - 1) {
        const int w = (it - (depth - 1)) % depth;
        unsigned done = 0;
        while (!done)
          asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                       : "=r"(done) : "r"(saddr(&bar[w])), "r"((phases >> w) & 1u) : "memory");
        phases ^= 1u << w;
        acc += reinterpret_cast<float*>(smem + w * box_bytes)[it & 63];
      }
    }
  }
  __syncthreads();
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) { cycles[blockIdx.x] = t1 - t0; if (acc == 12345.f) sink[0] = acc; }
}

int main(int argc, char** argv) {
  const int G = 240, C = 64, n_maps = 512;          // 512 maps x 14.7 MB = 7.5 GB >> L2
  float* maps; cudaMalloc(&maps, (size_t)n_maps * G * G * C * 4); cudaMemset(maps, 0, (size_t)n_maps * G * G * C * 4);
  unsigned long long* cyc; cudaMalloc(&cyc, 1024 * 8);
  float* sink; cudaMalloc(&sink, 4);
  typedef CUresult (*enc_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  enc_t enc = (enc_t)fp;
  cudaFuncSetAttribute(k_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
  const int shapes[][2] = {{104, 4}, {104, 2}, {40, 8}, {40, 40}};
  for (auto& sh : shapes) {
    const int BW = sh[0], BH = sh[1];
    CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)G, (cuuint64_t)G, (cuuint64_t)n_maps};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)G * C * 4, (cuuint64_t)G * G * C * 4};
    const cuuint32_t box[4] = {4, (cuuint32_t)BW, (cuuint32_t)BH, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, maps, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    const int box_bytes = BW * BH * 16;
    for (int depth : {2, 4, 8})
    for (int per_sm = 1; per_sm <= 2; ++per_sm)
      for (int hot = 0; hot <= 1; ++hot) {
        const int grid = 148 * per_sm, iters = 128;
        if (depth * box_bytes > 100 * 1024) continue;
        const int smem = per_sm == 1 ? 120 * 1024 : depth * box_bytes;          // 120 KB blocks a second CTA
        k_tma<<<grid, 128, smem>>>(tm, 32, box_bytes, hot ? 148 : n_maps, hot, depth, cyc, sink);   // warm
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        k_tma<<<grid, 128, smem>>>(tm, iters, box_bytes, hot ? 148 : n_maps, hot, depth, cyc, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        cudaError_t e = cudaGetLastError();
        unsigned long long h[1024]; cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
        const double cells = (double)BW * BH;
        printf("box %3dx%-2d cells depth %2d  %d CTA/SM  %s : %.0f cyc/box/CTA = %.2f cells/cyc/CTA, %.2f cells/cyc/SM; chip %.0f GB/s useful (%s)\n", BW, BH, depth, per_sm,
               hot ? "L2-hot " : "L2-cold", avg / iters, cells / (avg / iters), per_sm * cells / (avg / iters),
               (double)grid * iters * box_bytes / (ms * 1e6), cudaGetErrorString(e));
      }
  }
  return 0;
}
