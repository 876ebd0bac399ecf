// Micro-benchmark: TMA tile::gather4 row gathers (4 x 128-byte rows per instruction) into a
// shared-memory ring, against the LDG.128 gather of scripts/ubench.cu.  Decides whether the
// aggregation kernel stages neighbour rows with TMA (DESIGN.md "measured ceilings").
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_tma.bin scripts/ubench_tma_gather.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
      : "memory");
}

constexpr int kStageRows = 128;            // 32 lanes x 4 rows
constexpr int kStageBytes = kStageRows * 128;

// 1 producer warp + NC consumer warps; consumers sum the staged rows (8 lanes per row, float4 per lane)
template <int STAGES, int NC>
__global__ void __launch_bounds__(32 * (1 + NC)) k_tma_gather(const __grid_constant__ CUtensorMap map, float* __restrict__ out,
                                                             int64_t n_rows, int64_t chunks_total, uint32_t window) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* ring = reinterpret_cast<float*>(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * kStageBytes);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, NC); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    int s = 0; uint32_t ph = 0;
    for (int64_t c = blockIdx.x; c < chunks_total; c += gridDim.x) {
      mbar_wait(empty + s, ph ^ 1);
      if (lane == 0) mbar_expect_tx(full + s, kStageBytes);
      __syncwarp();
      const int64_t row = c * 32 + lane;   // pseudo "target row"; 4 pseudo neighbours
      const int64_t base = row / window * window;
      const uint32_t span = (uint32_t)min((int64_t)window, n_rows - base);
      const int r0 = (int)(base + hash32((uint32_t)row * 4u + 0) % span);
      const int r1 = (int)(base + hash32((uint32_t)row * 4u + 1) % span);
      const int r2 = (int)(base + hash32((uint32_t)row * 4u + 2) % span);
      const int r3 = (int)(base + hash32((uint32_t)row * 4u + 3) % span);
      tma_gather4(ring + (s * kStageRows + lane * 4) * 32, &map, 0, r0, r1, r2, r3, full + s);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
  } else {
    const int cw = warp - 1;
    const int sub = lane & 7, grp = lane >> 3;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int s = 0; uint32_t ph = 0;
    for (int64_t c = blockIdx.x; c < chunks_total; c += gridDim.x) {
      mbar_wait(full + s, ph);
      // consumer warp cw sums rows [cw*128/NC, (cw+1)*128/NC) of the stage, 4 rows per instruction
      const float* st = ring + s * kStageRows * 32;
#pragma unroll
      for (int r = cw * (kStageRows / NC) + grp; r < (cw + 1) * (kStageRows / NC); r += 4) {
        const float4 v = *reinterpret_cast<const float4*>(st + r * 32 + sub * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + s);
      if (++s == STAGES) { s = 0; ph ^= 1; }
    }
    *reinterpret_cast<float4*>(out + ((int64_t)blockIdx.x * NC * 32 + cw * 32 + lane) * 4) = acc;
  }
}

template <typename F>
float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 2; ++i) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int STAGES, int NC>
void run(const CUtensorMap& map, float* out, int64_t n_big, uint32_t window, int ctas_per_sm) {
  const int smem = STAGES * kStageBytes + 2 * STAGES * 8;
  CK(cudaFuncSetAttribute(k_tma_gather<STAGES, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int64_t chunks = n_big * 16 / 128;   // same gathered volume as ubench.cu (deg 16)
  float ms = time_ms([&] { k_tma_gather<STAGES, NC><<<148 * ctas_per_sm, 32 * (1 + NC), smem>>>(map, out, n_big, chunks, window); }, 5);
  CK(cudaGetLastError());
  printf("tma gather4: stages %d consumers %d ctas/SM %d window %9u: %.3f ms  %.2f TB/s\n", STAGES, NC, ctas_per_sm, window, ms,
         (double)chunks * kStageBytes / ms / 1e9);
}

int main() {
  const int64_t n_big = 143107LL * 25;
  float *x, *out;
  CK(cudaMalloc(&x, n_big * 32 * 4));
  CK(cudaMalloc(&out, 1 << 26));
  CK(cudaMemset(x, 0, n_big * 32 * 4));
  EncodeTiled encode = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
  if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  CUtensorMap map;
  cuuint64_t dims[2] = {32, (cuuint64_t)n_big};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {32, 1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const uint32_t windows[] = {16384, 143107, (uint32_t)n_big};
  for (uint32_t w : windows) {
    run<8, 4>(map, out, n_big, w, 1);
    run<4, 4>(map, out, n_big, w, 2);
    run<4, 4>(map, out, n_big, w, 3);
    run<12, 4>(map, out, n_big, w, 1);
    run<2, 2>(map, out, n_big, w, 6);
  }
  return 0;
}
