// Wide dense transform on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM):
//     y[N, Ho] = row_scale * act( x[N, Hi] W + bias + add ),      Ho in {64, 128, 192, 256}
// — SAGEConv's `matmul(aggr_out, weight) + bias` at hidden 256 (kernel/graph_sage.py:10,13,26-28; PyG 1.3
// SAGEConv.update), GCNConv's x W at hidden >= 64 (kernel/gcn.py:10,13) and their input gradients
// (the same call with the weight read through swapped strides).  The narrow botnet widths stay on
// the FMA / mma.sync paths (dense.cu, gcn_layer.cu): this is the "real dense contraction" of north_star.
//
// fp32 parity (rtol 1e-5) with TF32 inputs: every operand is split x = hi + lo (both rounded to
// nearest tf32) and   y = A_hi B_hi  +  (A_lo B_hi + A_hi B_lo)   with the main term and the two
// correction terms in SEPARATE TMEM accumulators (the accumulator add of the tensor core truncates;
// the small terms must not ride in the long chain), added with RN in the epilogue.
//
// One persistent CTA per SM, 256 threads, tile = 128 rows x Ho columns:
//   TMEM      512 columns: [0, Ho) main, [256, 256 + Ho) correction            (128 lanes = rows)
//   smem      2 stages x { A_hi, A_lo : 128 x 32 fp32;  B_hi, B_lo : Ho x 32 fp32 }   K chunk = 32
//             all in the un-swizzled K-major "interleaved" UMMA layout (8-row x 16-byte core matrices):
//             offset(r, c) = (c%4)*4 + (r%8)*16 + (c/4)*128 + (r/8)*1024      LBO = 128 B, SBO = 1024 B
//   W         pre-split once per call into exactly that smem image (k_wide_prep) and fetched per chunk
//             with one bulk async copy (cp.async.bulk, mbarrier transaction count) per plane
//   x         next chunk prefetched into registers with 128-bit loads while the current one is split
//             into A_hi / A_lo; tcgen05.mma issued by one thread; tcgen05.commit frees the stage
//   epilogue  tcgen05.ld 32 lanes x 32 columns per warp, + correction + bias (+ add) -> act -> store
// scripts/tc_probe.cu is the known-answer test of the descriptor / layout conventions used here.
#include <mutex>

#include "common.cuh"
#include "tc05.cuh"

namespace mgcn {

constexpr int kWM = 128;              // rows per tile
constexpr int kWK = 32;               // K chunk (fp32 elements): 128 bytes per row
constexpr int kWThreads = 256;
constexpr int kAChunkBytes = kWM * kWK * 4;   // 16 KB per plane

// W(k, c) = w[k*w_sk + c*w_sc] -> per K chunk two planes (hi, lo) of Ho x 32 floats in the smem image
__global__ void __launch_bounds__(256) k_wide_prep(const float* __restrict__ w, int64_t w_sk, int64_t w_sc, int Hi,
                                                   int Ho, int n_chunks, float* __restrict__ img) {
  const int per_chunk = Ho * kWK;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_chunks * per_chunk; i += gridDim.x * blockDim.x) {
    const int kc = i / per_chunk, rem = i % per_chunk;
    const int n = rem / kWK, kk = rem % kWK;
    const int k = kc * kWK + kk;
    const float v = k < Hi ? __ldg(w + (int64_t)k * w_sk + (int64_t)n * w_sc) : 0.f;
    const uint32_t hi = round_tf32_bits(__float_as_uint(v));
    const uint32_t lo = round_tf32_bits(__float_as_uint(v - __uint_as_float(hi)));
    float* base = img + (int64_t)kc * 2 * per_chunk;
    base[ileave_off(n, kk)] = __uint_as_float(hi);
    base[per_chunk + ileave_off(n, kk)] = __uint_as_float(lo);
  }
}

struct WideArgs {
  const float* x;
  const float* img;       // pre-split weight images
  const float* bias;
  const float* add;
  const float* row_scale;
  float* y;
  int64_t n_rows;
  int Hi, Ho, n_chunks, act;
};

constexpr int kWLdo = 36;   // floats per staged output row (144 bytes)

__global__ void __launch_bounds__(kWThreads, 1) k_linear_wide(const WideArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_plane_bytes = a.Ho * kWK * 4;
  const int stage_bytes = 2 * kAChunkBytes + 2 * b_plane_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes);   // [0,1] full_b, [2,3] mma_done, [4] tile_done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 5; ++i) mbar_init(bars + i, 1);
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
  const uint32_t idesc = umma_idesc_tf32(kWM, a.Ho);

  // this thread's share of an A chunk: rows (tid%8) + 8*(tid/32), 16-byte column chunk q = (tid/8)%4 + 4*i, i < 2
  // (lanes vary the row inside an 8-row core-matrix group fastest: conflict-free smem stores, 64-byte row segments in HBM)
  const int64_t n_tiles = (a.n_rows + kWM - 1) / kWM;
  const int r_in_tile = (lane & 7) + 8 * warp;          // rows r_in_tile + 64*j, j < 2
  const int q0 = (lane >> 3);                            // chunks q0 and q0 + 4
  float4 cur[4], nxt[4];

  auto load_chunk = [&](float4 (&dst)[4], int64_t tile, int kc) {
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int64_t gr = tile * kWM + r_in_tile + 64 * j;
        const int col = kc * kWK + 4 * (q0 + 4 * i);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tile < n_tiles && gr < a.n_rows && col < a.Hi) v = __ldg(reinterpret_cast<const float4*>(a.x + gr * a.Hi + col));
        dst[2 * j + i] = v;
      }
  };

  uint32_t it = 0, tile_phase = 0;
  int64_t tile = blockIdx.x;
  if (tile < n_tiles) load_chunk(cur, tile, 0);
  for (; tile < n_tiles; tile += gridDim.x) {
    for (int kc = 0; kc < a.n_chunks; ++kc, ++it) {
      const int s = it & 1;
      unsigned char* stage = smem + s * stage_bytes;
      float* a_hi = reinterpret_cast<float*>(stage);
      float* a_lo = reinterpret_cast<float*>(stage + kAChunkBytes);
      unsigned char* b_hi = stage + 2 * kAChunkBytes;
      // prefetch the next chunk of x (next k chunk, or the first chunk of this CTA's next tile)
      {
        const bool last = kc + 1 == a.n_chunks;
        load_chunk(nxt, last ? tile + gridDim.x : tile, last ? 0 : kc + 1);
      }
      // stage s is free once the MMAs issued two iterations ago have completed
      if (it >= 2) mbar_wait(bars + 2 + s, ((it >> 1) - 1) & 1);
      if (tid == 0) {
        mbar_expect_tx(bars + s, 2 * b_plane_bytes);
        bulk_g2s(b_hi, a.img + (int64_t)kc * 2 * a.Ho * kWK, 2 * b_plane_bytes, bars + s);   // hi and lo planes are adjacent
      }
      // split this thread's 4 x float4 into the hi / lo images
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float4 v = cur[2 * j + i];
          const float e[4] = {v.x, v.y, v.z, v.w};
          float h[4], l[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const uint32_t hb = round_tf32_bits(__float_as_uint(e[t]));
            h[t] = __uint_as_float(hb);
            l[t] = __uint_as_float(round_tf32_bits(__float_as_uint(e[t] - h[t])));
          }
          const int off = ileave_off(r_in_tile + 64 * j, 4 * (q0 + 4 * i));
          *reinterpret_cast<float4*>(a_hi + off) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(a_lo + off) = make_float4(l[0], l[1], l[2], l[3]);
        }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      __syncthreads();
      if (tid == 0) {
        mbar_wait(bars + s, (it >> 1) & 1);      // weight planes of this chunk have landed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa_hi = smem_u32(a_hi), sa_lo = smem_u32(a_lo);
        const uint32_t sb_hi = smem_u32(b_hi), sb_lo = sb_hi + b_plane_bytes;
#pragma unroll
        for (int k = 0; k < kWK / 8; ++k) {
          const uint32_t ko = 256 * k;   // 8 columns = two 16-byte chunks, 128 bytes apart
          const uint64_t dah = umma_desc(sa_hi + ko, 128, 1024), dal = umma_desc(sa_lo + ko, 128, 1024);
          const uint64_t dbh = umma_desc(sb_hi + ko, 128, 1024), dbl = umma_desc(sb_lo + ko, 128, 1024);
          umma_tf32(tmem, dah, dbh, idesc, (kc | k) != 0);           // main
          umma_tf32(tmem + 256, dal, dbh, idesc, (kc | k) != 0);     // corrections
          umma_tf32(tmem + 256, dah, dbl, idesc, 1);
        }
        umma_commit(bars + 2 + s);                       // stage reusable when these MMAs are done
        if (kc + 1 == a.n_chunks) umma_commit(bars + 4); // tile accumulators complete
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
    }
    // ---- epilogue: warps w and w+4 share TMEM lanes 32*(w%4).., and split the columns in halves ----
    mbar_wait(bars + 4, tile_phase);
    tile_phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      // A TMEM lane is an output row, so a thread holds 32 consecutive columns of ITS row: stored directly, a warp
      // instruction would touch 32 different 128-byte lines (measured in k_layer_bwd_tc: the LSU pays per line).
      // Each warp therefore transposes its 32 x 32 slab through a private staging tile (144-byte rows:
      // conflict-free both ways) and stores 4 whole 128-byte row pieces per instruction; `add` is read the same way.
      const int64_t row0 = tile * kWM + 32 * (warp & 3);
      const uint32_t lane_addr = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
      const int half = a.Ho / 2;
      // (the staging tiles alias operand stage 0: every MMA of the tile has completed, tile_done)
      float* stg = reinterpret_cast<float*>(smem) + warp * (32 * kWLdo);
      const int sr = lane >> 3, sq = lane & 7;          // phase 2: row sr + 4 i of the slab, 16-byte chunk sq
      for (int c0 = (warp >> 2) * half; c0 < (warp >> 2) * half + half; c0 += 32) {
        uint32_t m[32], c[32];
        tmem_ld32(lane_addr + c0, m);
        tmem_ld32(lane_addr + 256 + c0, c);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int q = 0; q < 32; q += 4) {
          float o[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) o[t] = __uint_as_float(m[q + t]) + __uint_as_float(c[q + t]);
          *reinterpret_cast<float4*>(stg + lane * kWLdo + q) = make_float4(o[0], o[1], o[2], o[3]);
        }
        __syncwarp();
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.bias) bv = __ldg(reinterpret_cast<const float4*>(a.bias + c0 + 4 * sq));
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = sr + 4 * i;
          const int64_t row = row0 + r;
          if (row < a.n_rows) {
            float4 v = *reinterpret_cast<const float4*>(stg + r * kWLdo + 4 * sq);
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            if (a.add) {
              const float4 av = __ldg(reinterpret_cast<const float4*>(a.add + row * a.Ho + c0 + 4 * sq));
              v.x += av.x; v.y += av.y; v.z += av.z; v.w += av.w;
            }
            if (a.act == 1) {
              v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y;
              v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w;
            }
            const float rs = a.row_scale ? __ldg(a.row_scale + row) : 1.f;
            *reinterpret_cast<float4*>(a.y + row * a.Ho + c0 + 4 * sq) = make_float4(v.x * rs, v.y * rs, v.z * rs, v.w * rs);
          }
        }
        __syncwarp();   // the staging tile is rewritten by the next slab
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // all TMEM reads done before the next tile's first MMA overwrites the accumulators
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace mgcn

using namespace mgcn;

extern "C" int mgcn_linear_wide(const float* x, int64_t N, int64_t Hi, const float* w, int64_t w_sk,
                                int64_t w_sc, int64_t Ho, const float* bias, const float* add, int act,
                                const float* row_scale, float* y, void* workspace,
                                size_t* workspace_bytes, void* stream) {
  MGCN_REQUIRE(workspace_bytes != nullptr, MGCN_ERR_NULL);
  MGCN_REQUIRE(N >= 0 && N < (int64_t(1) << 31), MGCN_ERR_RANGE);
  MGCN_REQUIRE(Ho >= 64 && Ho <= 256 && Ho % 64 == 0, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(Hi >= 4 && Hi <= 4096 && Hi % 4 == 0, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(act == 0 || act == 1, MGCN_ERR_SHAPE);
  const int n_chunks = (int)ceil_div(Hi, kWK);
  WorkspaceCarver ws(workspace);
  float* img = ws.take<float>((size_t)n_chunks * 2 * Ho * kWK);
  if (workspace == nullptr) {
    *workspace_bytes = ws.bytes();
    return MGCN_OK;
  }
  MGCN_REQUIRE(*workspace_bytes >= ws.bytes(), MGCN_ERR_WORKSPACE);
  if (N == 0) return MGCN_OK;
  MGCN_REQUIRE(x && w && y, MGCN_ERR_NULL);
  MGCN_REQUIRE(aligned16(x) && aligned16(y) && aligned16(img) && (!add || aligned16(add)), MGCN_ERR_ALIGN);
  MGCN_LAUNCH(k_wide_prep, 64, 256, 0, stream, w, w_sk, w_sc, (int)Hi, (int)Ho, n_chunks, img);
  WideArgs a{};
  a.x = x; a.img = img; a.bias = bias; a.add = add; a.row_scale = row_scale; a.y = y;
  a.n_rows = N; a.Hi = (int)Hi; a.Ho = (int)Ho; a.n_chunks = n_chunks; a.act = act;
  size_t smem = 2 * (2 * (size_t)kAChunkBytes + 2 * (size_t)Ho * kWK * 4) + 128 + 1024;
  if (smem < 120 * 1024) smem = 120 * 1024;   // one CTA per SM: each CTA allocates all 512 TMEM columns
  // the attribute belongs to (function, device): set on every call (cheap), so a second GPU in the same process
  // gets it too and a failure is reported every time
  cudaError_t attr_err = cudaSuccess;
  {
    attr_err = cudaFuncSetAttribute(k_linear_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  }
  MGCN_CHECK_CUDA(attr_err);
  int64_t tiles = ceil_div(N, kWM);
  const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);
  MGCN_LAUNCH(k_linear_wide, grid, kWThreads, smem, stream, a);
  return MGCN_OK;
}
