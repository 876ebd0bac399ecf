// Standalone B200 micro-benchmarks that ground the kernel design (DESIGN.md §"measured ceilings"):
//   1. row-gather ceiling: 8 lanes x float4 per 128-byte row, indices from a hash (no index-load
//      dependency), table sizes from L2-resident (one botnet graph) to HBM-sized (25 graphs)
//   2. legacy tensor path (mma.sync m16n8k8 tf32) issue rate per SM
//   3. plain copy (HBM stream) for reference
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench scripts/ubench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}

// each 8-lane group sums `deg` gathered rows (window-local random) and writes one row
template <int U>
__global__ void __launch_bounds__(256) k_gather(const float* __restrict__ x, float* __restrict__ out,
                                                int64_t n_rows, int deg, uint32_t window) {
  const int lane = threadIdx.x & 31, sub = lane & 7;
  const int64_t G = (int64_t)gridDim.x * blockDim.x / 8;
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 8; row < n_rows; row += G) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int64_t base = row / window * window;
    const uint32_t span = (uint32_t)min((int64_t)window, n_rows - base);
    for (int k = 0; k < deg; k += U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t j = base + hash32((uint32_t)row * 131u + k + u) % span;
        v[u] = __ldg(reinterpret_cast<const float4*>(x + j * 32 + sub * 4));
      }
#pragma unroll
      for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    *reinterpret_cast<float4*>(out + row * 32 + sub * 4) = acc;
  }
}

// 4 lanes x 256-bit per row (sm_100 LDG.256): 8 rows per warp instruction
template <int U>
__global__ void __launch_bounds__(256) k_gather256(const float* __restrict__ x, float* __restrict__ out,
                                                   int64_t n_rows, int deg, uint32_t window) {
  const int lane = threadIdx.x & 31, sub = lane & 3;
  const int64_t G = (int64_t)gridDim.x * blockDim.x / 4;
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 4; row < n_rows; row += G) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const int64_t base = row / window * window;
    const uint32_t span = (uint32_t)min((int64_t)window, n_rows - base);
    for (int k = 0; k < deg; k += U) {
      float v[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t j = base + hash32((uint32_t)row * 131u + k + u) % span;
        const float* p = x + j * 32 + sub * 8;
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[u][0]), "=f"(v[u][1]), "=f"(v[u][2]), "=f"(v[u][3]), "=f"(v[u][4]), "=f"(v[u][5]),
                       "=f"(v[u][6]), "=f"(v[u][7]) : "l"(p));
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += v[u][q];
    }
    float4* o = reinterpret_cast<float4*>(out + row * 32 + sub * 8);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

__global__ void __launch_bounds__(256) k_copy(const float4* __restrict__ a, float4* __restrict__ b, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    b[i] = a[i];
}

__global__ void __launch_bounds__(256) k_mma(float* out, int iters) {
  float d[4][4];
  uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, 7};
#pragma unroll
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) d[j][q] = 0.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += d[j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) k_mma_bf16(float* out, int iters) {
  float d[4][4];
  uint32_t a[4] = {threadIdx.x, threadIdx.x + 1, threadIdx.x + 2, threadIdx.x + 3}, b[2] = {threadIdx.x * 3, 7};
#pragma unroll
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) d[j][q] = 0.f;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                   : "+f"(d[j][0]), "+f"(d[j][1]), "+f"(d[j][2]), "+f"(d[j][3])
                   : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
  }
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) for (int q = 0; q < 4; ++q) s += d[j][q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main() {
  const int64_t n_big = 143107LL * 25;
  float *x, *out;
  CK(cudaMalloc(&x, n_big * 32 * 4));
  CK(cudaMalloc(&out, n_big * 32 * 4));
  CK(cudaMemset(x, 0, n_big * 32 * 4));
  const int deg = 16;
  printf("== gather ceiling: N rows, deg=%d, 128-byte rows, window = rows sharing a random range ==\n", deg);
  const uint32_t windows[] = {1024, 16384, 143107, 143107 * 4, (uint32_t)n_big};
  for (uint32_t w : windows) {
    for (int occ = 1; occ <= 2; ++occ) {
      const int blocks = 148 * 8 * occ;
      float ms = time_ms([&] { k_gather<8><<<blocks, 256>>>(x, out, n_big, deg, w); }, 5);
      double gathered = (double)n_big * deg * 128;
      printf("window %9u rows (%7.1f MB) grid %5d U=8: %.3f ms  gather %.2f TB/s (+write %.2f TB/s)\n", w,
             w * 128.0 / 1e6, blocks, ms, gathered / ms / 1e9, (double)n_big * 128 / ms / 1e9);
    }
  }
  {
    float ms = time_ms([&] { k_gather<4><<<148 * 8, 256>>>(x, out, n_big, deg, 143107); }, 5);
    printf("window 143107 U=4: %.3f ms gather %.2f TB/s\n", ms, (double)n_big * deg * 128 / ms / 1e9);
    ms = time_ms([&] { k_gather<16><<<148 * 8, 256>>>(x, out, n_big, deg, 143107); }, 5);
    printf("window 143107 U=16: %.3f ms gather %.2f TB/s\n", ms, (double)n_big * deg * 128 / ms / 1e9);
  }
  printf("== 128-bit x 8 lanes vs 256-bit x 4 lanes, window 143107 (one graph), by CTAs/SM and loads in flight ==\n");
  for (int cps = 2; cps <= 8; cps *= 2) {
    float ms = time_ms([&] { k_gather<4><<<148 * cps, 256>>>(x, out, n_big, deg, 143107); }, 5);
    printf("LDG.128 U=4 %d CTA/SM: %.3f ms %.2f TB/s\n", cps, ms, (double)n_big * deg * 128 / ms / 1e9);
    ms = time_ms([&] { k_gather<8><<<148 * cps, 256>>>(x, out, n_big, deg, 143107); }, 5);
    printf("LDG.128 U=8 %d CTA/SM: %.3f ms %.2f TB/s\n", cps, ms, (double)n_big * deg * 128 / ms / 1e9);
    ms = time_ms([&] { k_gather256<2><<<148 * cps, 256>>>(x, out, n_big, deg, 143107); }, 5);
    printf("LDG.256 U=2 %d CTA/SM: %.3f ms %.2f TB/s\n", cps, ms, (double)n_big * deg * 128 / ms / 1e9);
    ms = time_ms([&] { k_gather256<4><<<148 * cps, 256>>>(x, out, n_big, deg, 143107); }, 5);
    printf("LDG.256 U=4 %d CTA/SM: %.3f ms %.2f TB/s\n", cps, ms, (double)n_big * deg * 128 / ms / 1e9);
    ms = time_ms([&] { k_gather256<8><<<148 * cps, 256>>>(x, out, n_big, deg, 143107); }, 5);
    printf("LDG.256 U=8 %d CTA/SM: %.3f ms %.2f TB/s\n", cps, ms, (double)n_big * deg * 128 / ms / 1e9);
  }
  {
    float ms = time_ms([&] { k_copy<<<148 * 8, 256>>>((const float4*)x, (float4*)out, n_big * 8); }, 10);
    printf("copy %.1f MB: %.3f ms  %.2f TB/s (read+write)\n", n_big * 128.0 / 1e6, ms, 2.0 * n_big * 128 / ms / 1e9);
  }
  {
    const int iters = 4096;
    for (int bps = 1; bps <= 4; bps *= 2) {
      float ms = time_ms([&] { k_mma<<<148 * bps, 256>>>(out, iters); }, 3);
      double mmas = 148.0 * bps * 8 * iters * 4;
      printf("mma.sync tf32 m16n8k8, %d CTA/SM x 8 warps: %.3f ms, %.2f mma/clk/SM @1.9GHz, %.1f TFLOP/s\n", bps, ms,
             mmas / 148 / (ms * 1e-3 * 1.9e9), mmas * 4096 / ms / 1e9);
      ms = time_ms([&] { k_mma_bf16<<<148 * bps, 256>>>(out, iters); }, 3);
      printf("mma.sync bf16 m16n8k16, %d CTA/SM x 8 warps: %.3f ms, %.2f mma/clk/SM, %.1f TFLOP/s\n", bps, ms,
             mmas / 148 / (ms * 1e-3 * 1.9e9), mmas * 8192 / ms / 1e9);
    }
  }
  return 0;
}
