// Known-answer probe for the "A operand from tensor memory" form of tcgen05.mma (kind::tf32) fed by
// tcgen05.st.16x256b from a gather-style register layout — the building block DESIGN §3.4 names for taking the
// K-major operand images out of shared memory.
//   * lane (g = lane / 4, q = lane % 4) of warp w holds 8 consecutive columns [8q, 8q+8) of rows g and g + 8 of a
//     16-row block (what two gather passes leave in a 4-lane group's registers);
//   * one tcgen05.st.sync.aligned.16x256b.x4 writes the 16 x 32 block to TMEM lanes [32 (w % 4) + 16 h, +16),
//     columns [col0, col0 + 32): register 4 kb + {0,1} = row g, logical K positions 8 kb + 2 q + {0,1};
//     register 4 kb + {2,3} = row g + 8.  Logical K position kappa = 8 kb + 2 q + e holds physical column
//     c = 8 q + 2 kb + e: the contraction index is permuted, the B image is filled with the same permutation;
//   * tcgen05.mma.cta_group::1.kind::tf32 [d], [a_tmem], b_desc, idesc, p   (M = 128, N = 32, K = 8 per step);
//   * B: SWIZZLE_128B K-major image in shared memory.
// Inputs are small integers / 8 (exact in tf32), so any layout mistake shows as a mismatch against the CPU product.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/tc_probe_ts.bin scripts/tc_probe_ts.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t ltype) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)ltype << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
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

__global__ void __launch_bounds__(128) k_probe(const float* A, const float* B, float* D) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* sB = reinterpret_cast<float*>(smem);   // [32 rows n][128 bytes], SWIZZLE_128B, K permuted
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4096);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 4096 + 16);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // B'(n, kappa) = B[n][c(kappa)], kappa = 8 kb + 2 q + e  <->  c = 8 q + 2 kb + e
  for (int i = tid; i < 32 * 32; i += 128) {
    const int n = i >> 5, kappa = i & 31;
    const int kb = kappa >> 3, q = (kappa >> 1) & 3, e = kappa & 1;
    const int c = 8 * q + 2 * kb + e;
    const int chunk = kappa >> 2;
    sB[(n * 128 + ((chunk ^ (n & 7)) << 4)) / 4 + (kappa & 3)] = B[n * 32 + c];
  }
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
  // A -> TMEM columns [0, 32): two 16-row blocks per warp
  const int g = lane >> 2, q = lane & 3;
  for (int h = 0; h < 2; ++h) {
    const int r0 = 32 * warp + 16 * h + g, r1 = r0 + 8;
    uint32_t r[16];
    for (int kb = 0; kb < 4; ++kb) {
      r[4 * kb + 0] = __float_as_uint(A[r0 * 32 + 8 * q + 2 * kb]);
      r[4 * kb + 1] = __float_as_uint(A[r0 * 32 + 8 * q + 2 * kb + 1]);
      r[4 * kb + 2] = __float_as_uint(A[r1 * 32 + 8 * q + 2 * kb]);
      r[4 * kb + 3] = __float_as_uint(A[r1 * 32 + 8 * q + 2 * kb + 1]);
    }
    const uint32_t taddr = tmem + ((uint32_t)(32 * warp + 16 * h) << 16) + 0;
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t id = make_idesc(128, 32);
    const uint64_t bdesc = make_desc(smem_u32(sB), 16, 1024, 2);
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_t = tmem + 8 * k;        // 8 tf32 = 8 columns
      const uint64_t b_d = bdesc + 2 * k;       // 32 bytes inside the swizzle row
      const uint32_t acc = k > 0;
      asm volatile(
          "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
          " tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem + 32), "r"(a_t), "l"(b_d), "r"(id), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  mbar_wait(bar, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t v[32];
  const uint32_t la = tmem + 32 + ((uint32_t)(32 * warp) << 16);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(la));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int n = 0; n < 32; ++n) D[tid * 32 + n] = __uint_as_float(v[n]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}

int main() {
  std::vector<float> A(128 * 32), B(32 * 32);
  auto rnd = [](int i) { return (float)((int)((((unsigned)i * 2654435761u) >> 27) % 17u) - 8) / 8.f; };
  for (size_t i = 0; i < A.size(); ++i) A[i] = rnd((int)i + 1);
  for (size_t i = 0; i < B.size(); ++i) B[i] = rnd((int)i + 7777);
  float *dA, *dB, *dD;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, 128 * 32 * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  const int smem = 4096 + 64 + 1024;
  k_probe<<<1, 128, smem>>>(dA, dB, dD);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(128 * 32);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < 32; ++n) {
      float r = 0.f;
      for (int k = 0; k < 32; ++k) r += A[m * 32 + k] * B[n * 32 + k];
      if (r != D[m * 32 + n]) { if (bad < 6) printf("mismatch m=%d n=%d ref=%g got=%g\n", m, n, r, D[m * 32 + n]); ++bad; }
    }
  printf("A from tensor memory (tcgen05.st.16x256b.x4 + TS mma): %d mismatches of 4096\n", bad);
  return bad ? 2 : 0;
}
