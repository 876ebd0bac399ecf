// Probe of the tcgen05 / TMEM building blocks used by csrc/gcn_layer_tc.cu, against a CPU product.
// One un-swizzled "interleaved" tile layout (8-row x 16-byte core matrices) serves BOTH operand majors:
//   offset(r, c) = (c%4)*4 + (r%8)*16 + (c/4)*128 + (r/8)*(32*C)      for a tile of C columns
//   (1) K-major  (row-local products):  D[m][n]  = sum_k W[m][32+k] * B[n][k]     W wide tile [128][128] (cols 32..63), B [32][32]
//   (2) MN-major (transposed products): D2[m][n] = sum_r W[r][m] * X[r][n]        M = 128 columns of W, X [128][32], K = 128 rows
// Findings recorded in DESIGN.md: with SWIZZLE_128B the K-major form works but the MN-major form of a
// 32-bit type needs SWIZZLE_128B_BASE32B (a different smem image), so a swizzled tile cannot serve both.
// Inputs are small integers / 8 so single-pass TF32 is exact; any descriptor mistake shows as a mismatch.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tc_probe.bin scripts/tc_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 64-bit shared-memory matrix descriptor, SWIZZLE_128B, 128-byte rows (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t ltype = 2) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
  d |= (uint64_t)ltype << 61;   // layout type: 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// instruction descriptor: tf32 x tf32 -> f32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, int accumulate) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tc_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// row-major [rows][C] fp32 global -> interleaved tile; lanes take (row%8 = lane%8, chunk = lane/8 + 4i): conflict-free
__device__ __forceinline__ void fill_tile(float* tile, const float* src, int rows, int C, int tid, int nthreads) {
  const int chunks = C / 4;
  for (int i = tid; i < rows * chunks; i += nthreads) {
    const int r8 = i & 7, q = (i >> 3) % chunks, rg = i / (8 * chunks);
    const int r = rg * 8 + r8;
    const float4 v = *reinterpret_cast<const float4*>(src + r * C + q * 4);
    *reinterpret_cast<float4*>(reinterpret_cast<char*>(tile) + rg * (32 * C) + q * 128 + r8 * 16) = v;
  }
}

__global__ void __launch_bounds__(128) k_probe(const float* Wd, const float* B, const float* X, float* D, float* D2) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* sW = reinterpret_cast<float*>(smem);                    // wide tile [128][128]: 64 KB
  float* sX = reinterpret_cast<float*>(smem + 65536);            // [128][32]: 16 KB
  float* sB = reinterpret_cast<float*>(smem + 65536 + 16384);    // [32][32]: 4 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536 + 16384 + 4096);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 65536 + 16384 + 4096 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  fill_tile(sW, Wd, 128, 128, tid, 128);
  fill_tile(sX, X, 128, 32, tid, 128);
  fill_tile(sB, B, 32, 32, tid, 128);
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    // (1) K-major, no swizzle: A = columns 32..63 of the wide tile (LBO = 128 between 16-byte k chunks,
    //     SBO = 4096 between 8-row groups), B = [32][32] tile (LBO 128, SBO 1024); k-step = 8 columns = 2 chunks = 256 B
    const uint32_t id1 = make_idesc(128, 32, 0, 0);
    for (int k = 0; k < 4; ++k)
      mma_tf32(tmem, make_desc(smem_u32(sW) + 8 * 128 + 256 * k, 128, 4096, 0), make_desc(smem_u32(sB) + 256 * k, 128, 1024, 0), id1,
               k > 0);
    // (2) MN-major, no swizzle: A = all 128 columns (SBO = 128 between groups of 4 columns, LBO = 4096 between 8-row
    //     k groups), B = X (SBO 128, LBO 1024); k-step = 8 rows = one k group
    const uint32_t id2 = make_idesc(128, 32, 1, 1);
    for (int k = 0; k < 16; ++k)
      mma_tf32(tmem + 32, make_desc(smem_u32(sW) + 4096 * k, 4096, 128, 0), make_desc(smem_u32(sX) + 1024 * k, 1024, 128, 0), id2,
               k > 0);
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float v[32];
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  tmem_ld32(tmem + lane_base, v);
  for (int n = 0; n < 32; ++n) D[tid * 32 + n] = v[n];
  tmem_ld32(tmem + lane_base + 32, v);
  for (int n = 0; n < 32; ++n) D2[tid * 32 + n] = v[n];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<float> W(128 * 128), B(32 * 32), X(128 * 32);
  auto rnd = [](int i) { return (float)((int)((((unsigned)i * 2654435761u) >> 27) % 17u) - 8) / 8.f; };
  for (size_t i = 0; i < W.size(); ++i) W[i] = rnd((int)i + 1);
  for (size_t i = 0; i < B.size(); ++i) B[i] = rnd((int)i + 7777);
  for (size_t i = 0; i < X.size(); ++i) X[i] = rnd((int)i + 99991);
  float *dW, *dB, *dX, *dD, *dD2;
  CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dX, X.size() * 4));
  CK(cudaMalloc(&dD, 128 * 32 * 4)); CK(cudaMalloc(&dD2, 128 * 32 * 4));
  CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  const int smem = 65536 + 16384 + 4096 + 64 + 1024;
  CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_probe<<<1, 128, smem>>>(dW, dB, dX, dD, dD2);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * 32), D2(128 * 32);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D2.data(), dD2, D2.size() * 4, cudaMemcpyDeviceToHost));
  int bad1 = 0, bad2 = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 32; ++n) {
      float r = 0.f;
      for (int k = 0; k < 32; ++k) r += W[m * 128 + 32 + k] * B[n * 32 + k];
      if (r != D[m * 32 + n]) { if (bad1 < 3) printf("K-major mismatch m=%d n=%d ref=%g got=%g\n", m, n, r, D[m * 32 + n]); ++bad1; }
      float r2 = 0.f;
      for (int q = 0; q < 128; ++q) r2 += W[q * 128 + m] * X[q * 32 + n];
      if (r2 != D2[m * 32 + n]) { if (bad2 < 3) printf("MN-major mismatch m=%d n=%d ref=%g got=%g\n", m, n, r2, D2[m * 32 + n]); ++bad2; }
    }
  printf("interleaved layout: K-major %d mismatches of 4096; MN-major %d mismatches of 4096\n", bad1, bad2);
  return (bad1 || bad2) ? 2 : 0;
}
