// Warp-level 3xTF32 tile products on the legacy tensor path (mma.sync.m16n8k8.tf32) for the
// narrow (H = 32) dense transforms of the GCN stack.
//
// Why tensor cores at H = 32: measured on B200 (profiles/r1_v2_summary.md) the FMA-pipe transforms
// cost as much as the aggregation itself — 6 products of [N,32]x[32,32] per layer are 44 GFLOP,
// 0.5 TFLOP per step, i.e. >7 ms at the FP32 FMA peak against a 5 ms HBM roofline for the whole
// step.  fp32 parity (rtol 1e-5) rules out single-pass TF32 (10-bit mantissa), so every product is
// split:  x = hi + lo,  hi = x rounded to tf32, lo = x - hi (exact in fp32) rounded to tf32;
// A*B ~= A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, fp32 accumulate.  The dropped lo*lo term and the
// rounding of lo are ~2^-24 relative, the level of an fp32 rounding.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mgcn {

__device__ __forceinline__ void mma_tf32_16x8x8(float (&d)[4], const uint32_t (&a)[4],
                                                const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// x = hi + lo + eps with hi, lo exact tf32 values and |eps| <= 2^-24 |x|: both parts are rounded to
// nearest (ties away, what cvt.rna.tf32.f32 does) with integer arithmetic on the full-rate ALU pipe —
// the tensor core itself would truncate, which costs two more bits per operand (measured: the
// layer-0 weight gradient of the 12-layer model misses rtol 1e-5 with truncating splits).
__device__ __forceinline__ uint32_t round_tf32(uint32_t bits) { return (bits + 0x1000u) & 0xffffe000u; }

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = round_tf32(__float_as_uint(x));
  lo = round_tf32(__float_as_uint(x - __uint_as_float(hi)));
}

// Accumulation inside the tensor core is not round-to-nearest (the adder truncates; Ootomo &
// Yokota 2022 measure RZ), so long chains through the accumulator pick up a biased error.  Every
// product below therefore (1) starts its chains from zero, (2) keeps the two small correction terms
// in their own accumulator (2^-11 of the main term: their truncation is invisible), and (3) is added
// to whatever it contributes to with ordinary RN adds on the FP32 pipe by the caller.
//
// out[j] (16 x 8 tiles, j < N/8) = A[16 x K] * B[K x N]   (overwrites out).
// A: fp32 row-major in shared memory (leading dimension lda), split on the fly.
// B: pre-split hi / lo planes, row-major [K][ldb] in shared memory.
// Fragment coordinates (PTX ISA, m16n8k8 .tf32): g = lane/4, t = lane%4;
//   a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  b0 (k=t, n=g) b1 (k=t+4, n=g);
//   c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1).
template <int K, int N>
__device__ __forceinline__ void warp_gemm16(const float* __restrict__ As, int lda,
                                            const float* __restrict__ Bhi,
                                            const float* __restrict__ Blo, int ldb,
                                            float (&out)[N / 8][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
  float corr[N / 8][4];
#pragma unroll
  for (int j = 0; j < N / 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) out[j][q] = corr[j][q] = 0.f;
#pragma unroll
  for (int k0 = 0; k0 < K; k0 += 8) {
    uint32_t ahi[4], alo[4];
    split_tf32(As[g * lda + k0 + t], ahi[0], alo[0]);
    split_tf32(As[(g + 8) * lda + k0 + t], ahi[1], alo[1]);
    split_tf32(As[g * lda + k0 + t + 4], ahi[2], alo[2]);
    split_tf32(As[(g + 8) * lda + k0 + t + 4], ahi[3], alo[3]);
#pragma unroll
    for (int j = 0; j < N / 8; ++j) {
      uint32_t bhi[2], blo[2];
      bhi[0] = __float_as_uint(Bhi[(k0 + t) * ldb + 8 * j + g]);
      bhi[1] = __float_as_uint(Bhi[(k0 + t + 4) * ldb + 8 * j + g]);
      blo[0] = __float_as_uint(Blo[(k0 + t) * ldb + 8 * j + g]);
      blo[1] = __float_as_uint(Blo[(k0 + t + 4) * ldb + 8 * j + g]);
      mma_tf32_16x8x8(corr[j], alo, bhi);
      mma_tf32_16x8x8(corr[j], ahi, blo);
      mma_tf32_16x8x8(out[j], ahi, bhi);
    }
  }
#pragma unroll
  for (int j = 0; j < N / 8; ++j)
#pragma unroll
    for (int q = 0; q < 4; ++q) out[j][q] += corr[j][q];
}

// acc[i][j] (i < M/16, j < N/8) += A^T * B over R rows:  C[m][n] += sum_r A[r][m] * B[r][n].
// A: [R x M] and B: [R x N], both fp32 row-major in shared memory, both split on the fly.  The R-row
// contribution is formed from zero (main and correction chains) and added to acc with RN adds.
template <int R, int M, int N>
__device__ __forceinline__ void warp_gemm_tn(const float* __restrict__ As, int lda,
                                             const float* __restrict__ Bs, int ldb,
                                             float (&acc)[M / 16][N / 8][4], int lane) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < M / 16; ++i) {
    float mainc[N / 8][4], corr[N / 8][4];
#pragma unroll
    for (int j = 0; j < N / 8; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) mainc[j][q] = corr[j][q] = 0.f;
#pragma unroll
    for (int r0 = 0; r0 < R; r0 += 8) {
      uint32_t ahi[4], alo[4];
      split_tf32(As[(r0 + t) * lda + 16 * i + g], ahi[0], alo[0]);
      split_tf32(As[(r0 + t) * lda + 16 * i + g + 8], ahi[1], alo[1]);
      split_tf32(As[(r0 + t + 4) * lda + 16 * i + g], ahi[2], alo[2]);
      split_tf32(As[(r0 + t + 4) * lda + 16 * i + g + 8], ahi[3], alo[3]);
#pragma unroll
      for (int j = 0; j < N / 8; ++j) {
        uint32_t bhi[2], blo[2];
        split_tf32(Bs[(r0 + t) * ldb + 8 * j + g], bhi[0], blo[0]);
        split_tf32(Bs[(r0 + t + 4) * ldb + 8 * j + g], bhi[1], blo[1]);
        mma_tf32_16x8x8(corr[j], alo, bhi);
        mma_tf32_16x8x8(corr[j], ahi, blo);
        mma_tf32_16x8x8(mainc[j], ahi, bhi);
      }
    }
#pragma unroll
    for (int j = 0; j < N / 8; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] += mainc[j][q] + corr[j][q];
  }
}

}  // namespace mgcn
