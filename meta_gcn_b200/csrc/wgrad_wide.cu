// Weight gradient of a wide dense transform on the 5th-generation tensor cores:
//     dW[Hi x Ho] = x^T (g * (gmask > 0)),   db[Ho] = colsum(g * (gmask > 0))
// for SAGEConv / GCNConv `matmul(x, weight) + bias` at hidden >= 64 (kernel/graph_sage.py:10,13,
// kernel/gcn.py:10,13; autograd at kernel/train_eval.py:145).  The FMA kernel of dense.cu needs 73 ms for
// [2.45 M x 256]^T [2.45 M x 256] (4.4 TFLOP/s); this one is a transposed product on tcgen05.mma:
//     D[128 x Ho] (+)= X_chunk^T[128 x 8 rows] * G_chunk[8 rows x Ho]      both operands MN-major
// from SWIZZLE_128B_BASE32B images (the only layout a 32-bit MN-major operand is read correctly from,
// scripts/tc_probe.cu), 3xTF32 with the correction terms in their own accumulator:
//     TMEM cols [0, Ho) = x_hi^T g_hi,   [256, 256 + Ho) = x_lo^T g_hi + x_hi^T g_lo.
// CTA (mb, s): column block mb of x (128 of the Hi columns = the M dimension) and every S-th 32-row chunk.
// 16 warps: load one chunk ahead (registers) -> hi/lo split -> 12 conflict-free STS.128 per thread into one of
// two operand stages (96 KB each) -> one thread issues 12 tcgen05.mma per chunk; the tensor core works on chunk
// i while the CTA splits chunk i+1.  Chains through the TMEM accumulator are cut every kFlush chunks (the
// accumulator add is not round-to-nearest, mma_tile.cuh): the accumulators are then added (RN, fp32) to the
// CTA's partial in global memory and restarted from zero.  Partials are reduced in a fixed order:
// deterministic, no atomics.
#include <mutex>

#include "common.cuh"
#include "tc05.cuh"

namespace mgcn {

constexpr int kGRows = 32;                  // rows per chunk (4 k-steps of 8)
constexpr int kGImg = kGRows * 32 * 4;      // one image: 32 rows x 32 floats, 4 KB
constexpr int kGStage = 24 * kGImg;         // x hi/lo 4 atoms each, g hi/lo 8 atoms each: 96 KB
constexpr int kGOffXh = 0, kGOffXl = 4 * kGImg, kGOffGh = 8 * kGImg, kGOffGl = 16 * kGImg;
constexpr int kGThreads = 512;
constexpr int kFlush = 16;                  // chunks per accumulator chain: 64 k-steps
constexpr int kGSmem = 2 * kGStage + 256 + 1024;

struct WgradWideArgs {
  const float* x;
  const float* g;
  const float* gmask;
  float* partial;     // [S][Hi][Ho]
  float* partial_b;   // [S][Ho] or NULL
  int64_t n_rows;
  int Hi, Ho, S;
};

__device__ __forceinline__ int g_image_off(int r, int q) {     // bytes; SWIZZLE_128B_BASE32B, 16-byte chunk q of row r
  return (r << 7) + ((((q >> 1) ^ (r & 3)) << 5) | ((q & 1) << 4));
}
__device__ __forceinline__ void g_split4(const float4 v, float4& hi, float4& lo) {
  const float e[4] = {v.x, v.y, v.z, v.w};
  float h[4], l[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    h[t] = __uint_as_float(round_tf32_bits(__float_as_uint(e[t])));
    l[t] = __uint_as_float(round_tf32_bits(__float_as_uint(e[t] - h[t])));
  }
  hi = make_float4(h[0], h[1], h[2], h[3]);
  lo = make_float4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void g_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

struct Pair8 {
  float4 a, b;   // the two 16-byte chunks of a 32-byte pair
};

__global__ void __launch_bounds__(kGThreads, 1) k_wgrad_wide(const WgradWideArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_done = reinterpret_cast<uint64_t*>(smem + 2 * kGStage);   // [0,1] stage consumed, [2] chain complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 2 * kGStage + 64);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int n_mb = (a.Hi + 127) / 128;
  const int mb = blockIdx.x % n_mb, split = blockIdx.x / n_mb;
  if (tid == 0) {
    mbar_init(bar_done + 0, 1);
    mbar_init(bar_done + 1, 1);
    mbar_init(bar_done + 2, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint64_t pol = policy_evict_first();

  // lane 8 qq + j of warp w: row 8 (w & 3) + j of the chunk, atom w >> 2 (x) resp. w >> 2 and (w >> 2) + 4 (g), 32-byte
  // pair qq; the pair's two 16-byte chunks are stored in an order that depends on j >> 2 (conflict-free, gcn_layer_tc.cu)
  const int j = lane & 7, qq = lane >> 3, flip = j >> 2;
  const int r = 8 * (warp & 3) + j, atom = warp >> 2;
  const int qa = 2 * qq + flip, qb = 2 * qq + (flip ^ 1);
  const int off_a = g_image_off(r, qa), off_b = g_image_off(r, qb);
  const int n_gatoms = a.Ho >> 5;
  const int xcol = 128 * mb + 32 * atom + 8 * qq;               // first of this thread's 8 x columns
  const int64_t n_chunks = (a.n_rows + kGRows - 1) / kGRows;

  auto load8 = [&](const float* base, int64_t row, int ld, int col, int ncols, bool on) {
    Pair8 v;
    v.a = v.b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (on && row < a.n_rows) {
      if (col < ncols) v.a = ld_f4_hint(base + row * ld + col, pol);
      if (col + 4 < ncols) v.b = ld_f4_hint(base + row * ld + col + 4, pol);
    }
    return v;
  };
  auto mask8 = [](Pair8 v, const Pair8& m) {
    v.a.x = m.a.x > 0.f ? v.a.x : 0.f; v.a.y = m.a.y > 0.f ? v.a.y : 0.f;
    v.a.z = m.a.z > 0.f ? v.a.z : 0.f; v.a.w = m.a.w > 0.f ? v.a.w : 0.f;
    v.b.x = m.b.x > 0.f ? v.b.x : 0.f; v.b.y = m.b.y > 0.f ? v.b.y : 0.f;
    v.b.z = m.b.z > 0.f ? v.b.z : 0.f; v.b.w = m.b.w > 0.f ? v.b.w : 0.f;
    return v;
  };
  struct Chunk {
    Pair8 x, g0, g1;
  };
  auto load_chunk = [&](int64_t chunk) {
    Chunk c;
    const bool on = chunk < n_chunks;
    const int64_t row = chunk * kGRows + r;
    c.x = load8(a.x, row, a.Hi, xcol, a.Hi, on);
    const int gc0 = 32 * atom + 8 * qq, gc1 = 32 * (atom + 4) + 8 * qq;
    c.g0 = load8(a.g, row, a.Ho, gc0, a.Ho, on && atom < n_gatoms);
    c.g1 = load8(a.g, row, a.Ho, gc1, a.Ho, on && atom + 4 < n_gatoms);
    if (a.gmask) {
      c.g0 = mask8(c.g0, load8(a.gmask, row, a.Ho, gc0, a.Ho, on && atom < n_gatoms));
      c.g1 = mask8(c.g1, load8(a.gmask, row, a.Ho, gc1, a.Ho, on && atom + 4 < n_gatoms));
    }
    return c;
  };
  auto store_pair = [&](unsigned char* st, int off_hi, int off_lo, int atom_idx, const Pair8& v) {
    float4 hi, lo;
    g_split4(flip ? v.b : v.a, hi, lo);
    *reinterpret_cast<float4*>(st + off_hi + atom_idx * kGImg + off_a) = hi;
    *reinterpret_cast<float4*>(st + off_lo + atom_idx * kGImg + off_a) = lo;
    g_split4(flip ? v.a : v.b, hi, lo);
    *reinterpret_cast<float4*>(st + off_hi + atom_idx * kGImg + off_b) = hi;
    *reinterpret_cast<float4*>(st + off_lo + atom_idx * kGImg + off_b) = lo;
  };

  float db0[8], db1[8];   // column sums of this thread's g columns (CTAs of column block 0 only)
#pragma unroll
  for (int t = 0; t < 8; ++t) db0[t] = db1[t] = 0.f;

  const uint32_t idesc = umma_idesc_tf32(128, a.Ho, 1, 1);
  const uint64_t dM = umma_desc(smem_u32(smem), kGImg, 512, 1);
  // TMEM lanes of this thread for the flush: quarter warp & 3; columns [cq * Ho/4, (cq + 1) * Ho/4), cq = warp >> 2
  const int m_row = 128 * mb + 32 * (warp & 3) + lane;
  const int ncol_w = a.Ho >> 2, col_w0 = (warp >> 2) * ncol_w;
  float* prow = a.partial + ((int64_t)split * a.Hi + m_row) * a.Ho;
  const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);

  int64_t chunk = split;
  Chunk cur = load_chunk(chunk);
#ifdef MGCN_WGRAD_DEEP
  Chunk nxt = load_chunk(chunk + a.S);     // two chunks of look-ahead (measured: no gain, 4.0 ms either way)
#endif
  int it = 0, n_flushed = 0;
  uint32_t chain_phase = 0;
  for (; chunk < n_chunks; chunk += a.S, ++it) {
#ifdef MGCN_WGRAD_DEEP
    Chunk nxt2 = load_chunk(chunk + 2 * (int64_t)a.S);
#else
    Chunk nxt = load_chunk(chunk + a.S);
#endif
    const int s = it & 1;
    if (it >= 2) mbar_wait(bar_done + s, ((it >> 1) - 1) & 1);   // the tensor core has consumed this stage
    unsigned char* st = smem + s * kGStage;
    store_pair(st, kGOffXh, kGOffXl, atom, cur.x);
    if (atom < n_gatoms) store_pair(st, kGOffGh, kGOffGl, atom, cur.g0);
    if (atom + 4 < n_gatoms) store_pair(st, kGOffGh, kGOffGl, atom + 4, cur.g1);
    if (a.partial_b && mb == 0) {
      db0[0] += cur.g0.a.x; db0[1] += cur.g0.a.y; db0[2] += cur.g0.a.z; db0[3] += cur.g0.a.w;
      db0[4] += cur.g0.b.x; db0[5] += cur.g0.b.y; db0[6] += cur.g0.b.z; db0[7] += cur.g0.b.w;
      db1[0] += cur.g1.a.x; db1[1] += cur.g1.a.y; db1[2] += cur.g1.a.z; db1[3] += cur.g1.a.w;
      db1[4] += cur.g1.b.x; db1[5] += cur.g1.b.y; db1[6] += cur.g1.b.z; db1[7] += cur.g1.b.w;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
    __syncthreads();
    const bool first_of_chain = (it % kFlush) == 0;
    const bool last_of_chain = (it % kFlush) == kFlush - 1 || chunk + a.S >= n_chunks;
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t so = (uint32_t)(s * kGStage) >> 4;
#pragma unroll
      for (int k = 0; k < kGRows / 8; ++k) {
        const uint32_t ko = (1024 * k) >> 4;    // 8 rows = two 4-row k atoms
        const uint64_t xh = dM + (so + (kGOffXh >> 4) + ko), xl = dM + (so + (kGOffXl >> 4) + ko);
        const uint64_t gh = dM + (so + (kGOffGh >> 4) + ko), gl = dM + (so + (kGOffGl >> 4) + ko);
        const int acc = !(first_of_chain && k == 0);
        umma_tf32(tmem + 0, xh, gh, idesc, acc);      // main
        umma_tf32(tmem + 256, xl, gh, idesc, acc);    // corrections
        umma_tf32(tmem + 256, xh, gl, idesc, 1);
      }
      umma_commit(bar_done + s);
      if (last_of_chain) umma_commit(bar_done + 2);
    }
    if (last_of_chain) {
      // ---- flush: accumulators (+)-> this CTA's partial, RN adds; the next chain starts from zero ----
      mbar_wait(bar_done + 2, chain_phase);
      chain_phase ^= 1;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c0 = col_w0; c0 < col_w0 + ncol_w; c0 += 16) {
        uint32_t m[16], c[16];
        g_tmem_ld16(lane_addr + c0, m);
        g_tmem_ld16(lane_addr + 256 + c0, c);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (m_row < a.Hi) {
#pragma unroll
          for (int q = 0; q < 16; q += 4) {
            float4 v = make_float4(__uint_as_float(m[q]) + __uint_as_float(c[q]), __uint_as_float(m[q + 1]) + __uint_as_float(c[q + 1]),
                                   __uint_as_float(m[q + 2]) + __uint_as_float(c[q + 2]), __uint_as_float(m[q + 3]) + __uint_as_float(c[q + 3]));
            float4* p = reinterpret_cast<float4*>(prow + c0 + q);
            if (n_flushed > 0) {
              const float4 o = *p;
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *p = v;
          }
        }
      }
      ++n_flushed;
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncthreads();   // all TMEM reads done before the next chain's first MMA overwrites the accumulators
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    cur = nxt;
#ifdef MGCN_WGRAD_DEEP
    nxt = nxt2;
#endif
  }
  if (n_flushed == 0 && m_row < a.Hi) {   // a CTA without chunks still owns its slice of the partial
    for (int c0 = col_w0; c0 < col_w0 + ncol_w; c0 += 4) *reinterpret_cast<float4*>(prow + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // ---- db: rows of a warp are added by a fixed butterfly, the 4 row groups (warp & 3) in order ----
  __syncthreads();   // every MMA that read the stages has completed (last chain flushed): reuse stage 0 as scratch
  if (a.partial_b && mb == 0) {
    float* red = reinterpret_cast<float*>(smem);   // [4 row groups][256 columns]
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float v0 = db0[t], v1 = db1[t];
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
      }
      if (j == 0) {
        const int col0 = 32 * atom + 8 * qq + t, col1 = col0 + 128;
        if (col0 < a.Ho) red[(warp & 3) * 256 + col0] = v0;
        if (col1 < a.Ho) red[(warp & 3) * 256 + col1] = v1;
      }
    }
    __syncthreads();
    if (tid < a.Ho) a.partial_b[(int64_t)split * a.Ho + tid] = (red[tid] + red[256 + tid]) + (red[512 + tid] + red[768 + tid]);
  }
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

bool wide_wgrad_applies(int64_t N, int64_t Hi, int64_t Ho, const float* x, const float* g, const float* gmask) {
  return N >= 32768 && Hi >= 64 && Hi <= 256 && Hi % 4 == 0 && Ho >= 64 && Ho <= 256 && Ho % 32 == 0 && aligned16(x) &&
         aligned16(g) && (!gmask || aligned16(gmask));
}

int wide_wgrad_splits(int64_t Hi) {
  const int n_mb = (int)((Hi + 127) / 128);
  return kNumSMs / n_mb;
}

// partial: [splits][Hi][Ho], partial_b: [splits][Ho]
int launch_wide_wgrad(const float* x, int64_t N, int64_t Hi, const float* g, const float* gmask, int64_t Ho,
                      float* dw, int64_t dw_sk, int64_t dw_sc, float* db, float* partial, float* partial_b,
                      void* stream) {
  WgradWideArgs a{};
  a.x = x; a.g = g; a.gmask = gmask; a.partial = partial; a.partial_b = db ? partial_b : nullptr;
  a.n_rows = N; a.Hi = (int)Hi; a.Ho = (int)Ho; a.S = wide_wgrad_splits(Hi);
  const int n_mb = (int)((Hi + 127) / 128);
  // the attribute belongs to (function, device): set on every call (cheap), so a second GPU in the same process
  // gets it too and a failure is reported every time
  cudaError_t attr_err = cudaSuccess;
  {
    attr_err = cudaFuncSetAttribute(k_wgrad_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, kGSmem);
  }
  MGCN_CHECK_CUDA(attr_err);
  MGCN_LAUNCH(k_wgrad_wide, (unsigned)(a.S * n_mb), kGThreads, kGSmem, stream, a);
  int rc = launch_reduce_partials(partial, a.S, (int)(Hi * Ho), (int)Ho, dw, dw_sk, dw_sc, stream);
  if (rc != MGCN_OK) return rc;
  if (db) rc = launch_reduce_partials(partial_b, a.S, (int)Ho, (int)Ho, db, 0, 1, stream);
  return rc;
}

}  // namespace mgcn
