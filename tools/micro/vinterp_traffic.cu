// Microbenchmark (not part of the library): the MEMORY ACCESS PATTERN of k_vinterp_shared at C5
// without its arithmetic.  A warp owns 32 adjacent columns: it reads 128 B from each of n rows that
// lie 34.5 MB apart and writes 32 x m outputs (column-major per column: 400 B each).  Variants tell
// which side limits the kernel's ~2.3 TB/s.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/vinterp_traffic tools/micro/vinterp_traffic.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int kWarps = 8;

// mode bit 0: read rows, bit 1: write outputs.  wide = columns per lane (1: 128 B per warp-row, 4: 512 B)
template <int WIDE, bool FULLTILE>
__global__ void __launch_bounds__(kWarps * 32, 4) k_traffic(const float* __restrict__ phi, float* __restrict__ out,
                                                            int n, long long inner, int m, long long ntiles, int mode) {
  extern __shared__ float smem[];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int TM = FULLTILE ? 100 : 32;  // targets buffered per store
  float (*tile)[TM + 1] = reinterpret_cast<float (*)[TM + 1]>(smem + (size_t)w * 32 * WIDE * (TM + 1));
  for (long long cg = (long long)blockIdx.x * kWarps + w; cg < ntiles; cg += (long long)gridDim.x * kWarps) {
    const long long col0 = cg * 32 * WIDE;
    float acc[WIDE];
#pragma unroll
    for (int q = 0; q < WIDE; ++q) acc[q] = 0.f;
    if (mode & 1) {
      const float* p = phi + col0 + lane * WIDE;
#pragma unroll 4
      for (int j = 0; j < n; ++j) {
        if (WIDE == 4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(p + (long long)j * inner));
          acc[0] += v.x; acc[1 % WIDE] += v.y; acc[2 % WIDE] += v.z; acc[3 % WIDE] += v.w;
        } else {
          acc[0] += __ldg(p + (long long)j * inner);
        }
      }
    }
    if (mode & 2) {
      for (int t0 = 0; t0 < m; t0 += TM) {
        const int nt = (m - t0 < TM) ? (m - t0) : TM;
#pragma unroll
        for (int q = 0; q < WIDE; ++q)
          for (int t = 0; t < nt; ++t) tile[lane * WIDE + q][t] = acc[q] + t;
        __syncwarp();
        // each column's nt targets are contiguous in memory: 32 lanes x 4 B pieces
        for (int cc = 0; cc < 32 * WIDE; ++cc)
          for (int t = lane; t < nt; t += 32) __stcs(out + (col0 + cc) * m + t0 + t, tile[cc][t]);
        __syncwarp();
      }
    } else if (acc[0] == 12345.678f) {
      out[col0] = acc[0];
    }
  }
}

template <int WIDE, bool FULLTILE>
float run(const float* phi, float* out, int n, long long inner, int m, int mode) {
  const long long ntiles = inner / (32 * WIDE);
  constexpr int TM = FULLTILE ? 100 : 32;
  const size_t smem = (size_t)kWarps * 32 * WIDE * (TM + 1) * sizeof(float);
  cudaFuncSetAttribute(k_traffic<WIDE, FULLTILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_traffic<WIDE, FULLTILE>, kWarps * 32, smem);
  const int blocks = 148 * per_sm;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e9f;
  for (int it = 0; it < 5; ++it) {
    cudaEventRecord(a);
    k_traffic<WIDE, FULLTILE><<<blocks, kWarps * 32, smem>>>(phi, out, n, inner, m, ntiles, mode);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (it && ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  printf("wide=%d fulltile=%d mode=%d blocks/SM=%d: %.3f ms", WIDE, (int)FULLTILE, mode, per_sm, best);
  return best;
}

int main() {
  const int n = 75, m = 100;
  const long long inner = 2400LL * 3600;
  float *phi, *out;
  cudaMalloc(&phi, sizeof(float) * n * inner);
  cudaMalloc(&out, sizeof(float) * m * inner);
  cudaMemset(phi, 0, sizeof(float) * n * inner);
  const double rb = 4.0 * n * inner, wb = 4.0 * m * inner;
  for (int mode = 1; mode <= 3; ++mode) {
    const double bytes = ((mode & 1) ? rb : 0) + ((mode & 2) ? wb : 0);
    float ms = run<1, false>(phi, out, n, inner, m, mode);
    printf("  -> %.0f GB/s\n", bytes / ms / 1e6);
    ms = run<4, false>(phi, out, n, inner, m, mode);
    printf("  -> %.0f GB/s\n", bytes / ms / 1e6);
    if (mode & 2) {
      ms = run<1, true>(phi, out, n, inner, m, mode);
      printf("  -> %.0f GB/s\n", bytes / ms / 1e6);
    }
  }
  return 0;
}
