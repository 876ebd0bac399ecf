// Backward of one residual GCN layer at hidden 32 as ONE launch: the transposed aggregation and the row-local products
// (autograd of gcn_model.py:89-106 around NodeModelAdditive.forward, gcn_base_models.py:199-243; stored format of
// gcn_fwd_tc.cu):
//
//     dxw_j   = row_scale_j * sum_{e: row[e]=j} gs[col[e]]                 (A_hat^T of the masked gradient, gathered)
//     G_j     = dxw_j W^T + gy_j R                                         x_j = z_j / x_scale_j
//     dW      = x^T dxw;   dR = gy^T x;   dr = colsum(gy)
//     gy_prev = G * (x > 0);   gs_prev = post * gy_prev * bits(hmask_prev)
//
// k_agg_flat + k_layer_bwd_tc did this in two launches with the [N,32] array dxw written to HBM and read back; here the
// gather warps put their sums straight into the operand images of the tensor core, so a layer's backward reads gs
// (gathered), gy and z and writes gy_prev and gs_prev — 0.9 GB less DRAM traffic per layer at the botnet batch.
//
// One persistent CTA per SM, 24 warps at 80 registers, the skeleton of k_gcn_fwd_tc over 64-row tiles:
//   producers  20 warps; a pass = 8 consecutive tasks of the by-source work order, one per 4-lane group: gather-sum of
//              gs (gather.cuh), the row's own gy (registers, in flight during the gather) and z (cp.async into a
//              per-warp staging kilobyte: no registers while the gather runs).  A finished pass takes the next free
//              8-row slot of the CTA (tiles fill in COMPLETION order), splits dxw, gy and x into tf32 hi / lo and
//              stores ten images: SWIZZLE_128B K-major images of dxw and gy for the row-local products and
//              SWIZZLE_128B_BASE32B MN-major images of dxw, gy and x for the transposed ones (tc05.cuh: a 32-bit
//              MN-major operand is only read correctly from that image, a K-major one never).  Group g of a warp owns
//              slot row ((g & 3) << 1) | (g >> 2) and stores its two 16-byte chunks in the order (g & 1): the 8 lanes of
//              a quarter warp then hit 8 different 16-byte bank groups in BOTH image types.  Slots without a finished
//              row (hub segments, padding) zero their MN-major rows.  The warp whose pass is the 8th of a tile issues
//              its 32 tcgen05.mma from one lane:
//                G      [64 x 64|32]  K-major,  M = 64:  main = dxw_hi Wt_hi + gy_hi R_hi, corrections in their own columns
//                T      [128 x 64]    MN-major, M = 128: [dxw_hi|gy_hi|dxw_lo|gy_lo]^T [x_hi | x_lo], 8 steps of 8 rows
//                colsum [64 x 64]     ones[64 x 8] (K-major) x [gy_hi | gy_lo] (MN-major): every row = colsum of the tile
//   epilogue   4 warps: T and the column sums -> running fp32 registers (RN adds; chains through the TMEM accumulator
//              stay 8 steps long); G -> both masks, per-target factor -> two staged 144-byte rows -> two 128-byte
//              cp.async.bulk stores per row (lanes 0..15 of a warp carry the 16 rows of its TMEM quarter)
//   barriers   as k_gcn_fwd_tc: full[tile % 4] (8 pass arrivals), done[tile % 4] (tcgen05.commit), tfree[stage].
// Hub rows (longer than the hub threshold): their segments store partial sums in mode 0; a second launch (mode 1) sums
// the partials per hub row (fixed order) and runs the same tile path.  Per-CTA partials of dW / dR / dr are reduced in a
// fixed order by k_bwd_tc_reduce / k_reduce_partials.  No atomics on data: run-to-run identical.
#include "common.cuh"
#include "gather.cuh"
#include "tc05.cuh"

namespace mgcn {

constexpr int kBfRows = 64;                      // rows per tile
constexpr int kBfPasses = kBfRows / 8;           // 8-row slots per tile
constexpr int kBfImg = kBfRows * 128;            // one operand image: 64 rows x 128 bytes
constexpr int kBfKDh = 0 * kBfImg;               // K-major dxw_hi, dxw_lo, gy_hi, gy_lo
constexpr int kBfKDl = 1 * kBfImg;
constexpr int kBfKGh = 2 * kBfImg;
constexpr int kBfKGl = 3 * kBfImg;
constexpr int kBfMN = 4 * kBfImg;                // MN-major [dxw_hi | gy_hi | dxw_lo | gy_lo]
constexpr int kBfMXh = 8 * kBfImg;               // MN-major x_hi, x_lo
constexpr int kBfMXl = 9 * kBfImg;
constexpr int kBfStageB = 10 * kBfImg;           // 80 KB
constexpr int kBfStages = 2;
constexpr int kBfOffB1 = kBfStages * kBfStageB;  // [Wt_hi ; Wt_lo]: 64 rows x 128 bytes, SWIZZLE_128B K-major
constexpr int kBfOffB2 = kBfOffB1 + 8192;        // [R_hi ; R_lo]
constexpr int kBfOffOnes = kBfOffB2 + 8192;      // ones[64 x 8], interleaved K-major (2 KB)
constexpr int kBfLdo = 20;                       // floats per staged half row (64 + 16 bytes: conflict-free both ways)
constexpr int kBfOffOut = kBfOffOnes + 2048;     // [8 epilogue warps][gy_prev | gs_prev][16 rows][kBfLdo]
#ifndef MGCN_BF_PROD
#define MGCN_BF_PROD 16
#endif
constexpr int kBfProdWarps = MGCN_BF_PROD;
constexpr int kBfEpiWarps = 8;                   // TMEM quarter = warp & 3, column half = warp >> 2
constexpr int kBfThreads = 32 * (kBfEpiWarps + kBfProdWarps);
constexpr int kBfOffZst = kBfOffOut + kBfEpiWarps * 2 * 16 * kBfLdo * 4;   // per producer warp: 8 rows x 128 bytes of z
constexpr int kBfOffScal = kBfOffZst + kBfProdWarps * 1024;            // [tile % 4][64] {row id, x > 0 bits, hmask_prev, post}
constexpr int kBfOffMisc = kBfOffScal + 2 * kBfStages * kBfRows * 16;  // barriers, counters, tmem slot
constexpr int kBfSmem = kBfOffMisc + 256 + 1024;
constexpr int kBfTmemBuf = 192;                  // columns per accumulator buffer: G [0,64), T [64,128), colsum [128,192)
static_assert(kBfSmem <= 232448, "shared memory");

struct BwdFusedArgs {
  const int4* tasks;
  const int32_t* nbr_w;
  const int32_t* seg_count;
  const int32_t* hub_rows;
  const int32_t* hub_seg0;
  const int32_t* hub_count;
  const int32_t* rowptr;
  const float* gs;           // [N,32] gathered
  const float* gy;           // [N,32]
  const float* z;            // [N,32] stored layer input, x_scale (.) x
  const float* x_scale;      // [N] or NULL
  const float* row_scale;    // [N] or NULL: factor of the aggregated row (per-source degree factor)
  const float* w;            // weight_node (in j, out c)
  const float* res_w;        // residual weight (out c, in j)
  const uint32_t* hmask_prev;
  const float* post;         // [N] or NULL
  float* gy_prev;            // [N,32] or NULL
  float* gs_prev;
  float* partial;            // [seg_cap,32] hub segment sums
  float* part_t;             // [grid][128][32] of THIS launch
  float* part_b;             // [grid][2][32]
  int64_t n_rows;
  int64_t seg_cap;
  int64_t hub_cap;
  int hub_threshold;
  int static_slots;          // 1: pass g of a CTA always fills slot g (weight gradients identical from run to run)
};

__device__ __forceinline__ int bf_sw128_off(int r, int q) {   // bytes; 16-byte chunk q of row r, SWIZZLE_128B K-major
  return (r << 7) + ((q ^ (r & 7)) << 4);
}
__device__ __forceinline__ int bf_mn_off(int r, int q) {      // bytes; SWIZZLE_128B_BASE32B
  return (r << 7) + ((((q >> 1) ^ (r & 3)) << 5) | ((q & 1) << 4));
}

__device__ __forceinline__ void bf_split4(float a, float b, float c, float d, float4& hi, float4& lo) {
  const float e[4] = {a, b, c, d};
  float h[4], l[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    h[t] = __uint_as_float(round_tf32_bits(__float_as_uint(e[t])));
    l[t] = __uint_as_float(round_tf32_bits(__float_as_uint(e[t] - h[t])));
  }
  hi = make_float4(h[0], h[1], h[2], h[3]);
  lo = make_float4(l[0], l[1], l[2], l[3]);
}

// the lane's 8 columns of one row -> hi / lo images; first the chunk 2 sub + flip, then the other one
template <bool kWithK>
__device__ __forceinline__ void bf_store_row(unsigned char* k_hi, unsigned char* k_lo, unsigned char* m_hi,
                                             unsigned char* m_lo, int r, int sub, int flip, const Row8& v) {
  float4 hi, lo;
  const int ca = 2 * sub + flip, cb = 2 * sub + 1 - flip;
  bf_split4(flip ? v.v[4] : v.v[0], flip ? v.v[5] : v.v[1], flip ? v.v[6] : v.v[2], flip ? v.v[7] : v.v[3], hi, lo);
  if (kWithK) {
    *reinterpret_cast<float4*>(k_hi + bf_sw128_off(r, ca)) = hi;
    *reinterpret_cast<float4*>(k_lo + bf_sw128_off(r, ca)) = lo;
  }
  *reinterpret_cast<float4*>(m_hi + bf_mn_off(r, ca)) = hi;
  *reinterpret_cast<float4*>(m_lo + bf_mn_off(r, ca)) = lo;
  bf_split4(flip ? v.v[0] : v.v[4], flip ? v.v[1] : v.v[5], flip ? v.v[2] : v.v[6], flip ? v.v[3] : v.v[7], hi, lo);
  if (kWithK) {
    *reinterpret_cast<float4*>(k_hi + bf_sw128_off(r, cb)) = hi;
    *reinterpret_cast<float4*>(k_lo + bf_sw128_off(r, cb)) = lo;
  }
  *reinterpret_cast<float4*>(m_hi + bf_mn_off(r, cb)) = hi;
  *reinterpret_cast<float4*>(m_lo + bf_mn_off(r, cb)) = lo;
}

__device__ __forceinline__ Row8 bf_ld_row8_stream(const float* p, uint64_t pol) {   // one LDG.256, streamed
  Row8 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]),
                 "=f"(r.v[7])
               : "l"(p), "l"(pol));
  return r;
}

__device__ __forceinline__ void bf_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void bf_tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

template <int kMode>   // 0: tiles of the work order (rows + hub segments), 1: tiles of the hub list
__global__ void __launch_bounds__(kBfThreads, 1) k_gcn_bwd_fused(const BwdFusedArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_done = reinterpret_cast<uint64_t*>(smem + kBfOffMisc);         // [2 S]
  uint64_t* bar_tfree = bar_done + 2 * kBfStages;                               // [S]
  uint64_t* bar_full = bar_tfree + kBfStages;                                   // [2 S] 8 pass arrivals per tile
  uint32_t* arrivals = reinterpret_cast<uint32_t*>(bar_full + 2 * kBfStages);   // [S] passes stored, never reset
  uint32_t* next_slot = arrivals + kBfStages;                                   // passes finished by this CTA so far
  uint32_t* tmem_slot = next_slot + 1;
  uint4* scal = reinterpret_cast<uint4*>(smem + kBfOffScal);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const bool want_prev = a.gy_prev != nullptr;

  // weight images (SWIZZLE_128B K-major, 64 rows: hi then lo): B1(n, k) = W[n][k], B2(n, k) = R[k][n]
  for (int i = tid; i < 32 * 32; i += kBfThreads) {
    const int n = i >> 5, k = i & 31;
    const float w1 = __ldg(a.w + n * 32 + k), w2 = __ldg(a.res_w + k * 32 + n);
    const float h1 = __uint_as_float(round_tf32_bits(__float_as_uint(w1)));
    const float h2 = __uint_as_float(round_tf32_bits(__float_as_uint(w2)));
    float* b1 = reinterpret_cast<float*>(smem + kBfOffB1);
    float* b2 = reinterpret_cast<float*>(smem + kBfOffB2);
    const int o_hi = (bf_sw128_off(n, k >> 2) >> 2) + (k & 3), o_lo = (bf_sw128_off(n + 32, k >> 2) >> 2) + (k & 3);
    b1[o_hi] = h1;
    b1[o_lo] = __uint_as_float(round_tf32_bits(__float_as_uint(w1 - h1)));
    b2[o_hi] = h2;
    b2[o_lo] = __uint_as_float(round_tf32_bits(__float_as_uint(w2 - h2)));
  }
  for (int i = tid; i < 512; i += kBfThreads) reinterpret_cast<float*>(smem + kBfOffOnes)[i] = 1.f;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kBfStages; ++s) {
      arrivals[s] = 0;
      if (s == 0) *next_slot = 0;
      mbar_init(bar_done + s, 1);
      mbar_init(bar_done + kBfStages + s, 1);
      mbar_init(bar_tfree + s, kBfEpiWarps);
      mbar_init(bar_full + s, kBfPasses);
      mbar_init(bar_full + kBfStages + s, kBfPasses);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint64_t pol = policy_evict_first();

  int64_t limit;   // valid tile entries: tasks (mode 0) or hub rows (mode 1)
  if (kMode == 0) {
    int64_t nseg = 0;
    if (a.seg_count) {
      nseg = *a.seg_count;
      if (nseg > a.seg_cap) nseg = a.seg_cap;
    }
    limit = a.n_rows + nseg;
  } else {
    limit = *a.hub_count;
    if (limit > a.hub_cap) limit = a.hub_cap;
  }
  const int64_t n_tiles = (limit + kBfRows - 1) / kBfRows;

  if (warp >= kBfEpiWarps) {
    // ------------------------------- producers -------------------------------
    const int pw = warp - kBfEpiWarps;
    const uint32_t idG64 = umma_idesc_tf32(64, 64), idG32 = umma_idesc_tf32(64, 32);
    const uint32_t idT = umma_idesc_tf32(128, 64, 1, 1), idC = umma_idesc_tf32(64, 64, 0, 1);
    const uint64_t dK = umma_desc(smem_u32(smem), 16, 1024, 2);          // SWIZZLE_128B K-major: SBO = 8 rows x 128 bytes
    const uint64_t dM = umma_desc(smem_u32(smem), kBfImg, 512, 1);       // BASE32B MN-major, atoms one image apart
    const uint64_t dM2 = umma_desc(smem_u32(smem), 2 * kBfImg, 512, 1);  // atoms two images apart: gy_hi | gy_lo
    const uint64_t dOnes = umma_desc(smem_u32(smem + kBfOffOnes), 128, 256, 0);
    const int sub = lane & 3, grp = lane >> 2, grp_lane0 = grp * 4;
    const unsigned gmask = 0xfu << grp_lane0;
    const int col = sub * 8;
    const int rslot = ((grp & 3) << 1) | (grp >> 2), flip = grp & 1;
    // this lane's 32 bytes of its row's z: two 16-byte chunks, positions swapped in odd rows (the 8 lanes of a quarter
    // warp then cover 8 different bank groups)
    unsigned char* zst0 = smem + kBfOffZst + pw * 1024 + grp * 128 + (((2 * sub) ^ (grp & 1)) << 4);
    unsigned char* zst1 = smem + kBfOffZst + pw * 1024 + grp * 128 + (((2 * sub + 1) ^ (grp & 1)) << 4);
    auto load_desc = [&](int64_t g) {
      int4 d = make_int4(-1, 0, 0, 0);
      const int64_t tile = blockIdx.x + (g / kBfPasses) * (int64_t)gridDim.x;
      const int64_t e = tile * kBfRows + (g % kBfPasses) * 8 + lane;
      if (lane < 8 && tile < n_tiles && e < limit) {
        if (kMode == 0) {
          d = ld_i4_hint(a.tasks + e, pol);
        } else {
          const int row = __ldg(a.hub_rows + e);
          const int len = __ldg(a.rowptr + row + 1) - __ldg(a.rowptr + row);
          d = make_int4(row, __ldg(a.hub_seg0 + e), (len + a.hub_threshold - 1) / a.hub_threshold, 0);
        }
      }
      return d;
    };
    struct PassIdx {
      int beg, end, gi, gin;
    };
    auto load_idx = [&](const int4& d) {
      PassIdx p{0, 0, 0, 0};
      if (kMode == 0) {
        p.beg = __shfl_sync(0xffffffffu, d.y, grp);
        p.end = __shfl_sync(0xffffffffu, d.z, grp);
        if (p.beg + sub < p.end) p.gi = ld_i32_hint(a.nbr_w + p.beg + sub, pol);
        if (p.beg + 4 + sub < p.end) p.gin = ld_i32_hint(a.nbr_w + p.beg + 4 + sub, pol);
      }
      return p;
    };
    int4 d = load_desc(pw);
    int4 dn = load_desc(pw + kBfProdWarps);
    PassIdx pi = load_idx(d);
    const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
    for (int64_t g = pw; g < my_tiles * kBfPasses; g += kBfProdWarps) {
      const int rowp = __shfl_sync(0xffffffffu, d.x, grp);
      const int slot = __shfl_sync(0xffffffffu, d.w, grp);
      const bool finish = rowp >= 0 && slot == 0;   // this group completes a row
      const int64_t own = (int64_t)(finish ? rowp : 0) * kGH + col;
      // the row's own gy (registers) and z (cp.async: lands in shared memory while the gather runs); its scalars, one
      // per lane of the group: row_scale, x_scale, hmask_prev, post — looked at only when the row is stored
      const Row8 gyrow = bf_ld_row8_stream(a.gy + own, pol);
      cp_async16_hint(zst0, a.z + own, 16, pol);
      cp_async16_hint(zst1, a.z + own + 4, 16, pol);
      asm volatile("cp.async.commit_group;" ::: "memory");
      uint32_t sc = 0x3f800000u;   // 1.0f
      if (sub == 2) sc = 0;
      {
        const void* sp = sub == 0 ? (const void*)a.row_scale : sub == 1 ? (const void*)a.x_scale
                         : sub == 2 ? (const void*)a.hmask_prev : (const void*)a.post;
        if (finish && sp) sc = __ldg(reinterpret_cast<const uint32_t*>(sp) + rowp);
      }
      const PassIdx pc = pi;
      const int4 dc = d;
      d = dn;
      pi = load_idx(d);                               // next pass: index batches in flight during this gather
      dn = load_desc(g + 2 * kBfProdWarps);
      Row8 acc;
#pragma unroll
      for (int q = 0; q < 8; ++q) acc.v[q] = 0.f;
      if (kMode == 0) {
        if (rowp >= 0) {
          acc = gather_sum(a.gs, a.nbr_w, pc.beg, pc.end, pc.gi, pc.gin, sub, grp_lane0, gmask, col, pol);
          if (slot != 0) store_partial(a.partial, slot, col, acc);
        }
      } else {
        // the whole warp sums one hub row at a time: 8 contiguous runs of its segment partials in parallel, the run
        // sums added left to right (a fixed order)
#pragma unroll 1
        for (int i = 0; i < 8; ++i) {
          const int rowi = __shfl_sync(0xffffffffu, dc.x, i);
          if (rowi < 0) continue;
          const int s0 = __shfl_sync(0xffffffffu, dc.y, i), ns = __shfl_sync(0xffffffffu, dc.z, i);
          const int per = (ns + 7) >> 3;
          const Row8 run = hub_run_sum(a.partial, s0, grp * per, min(ns, grp * per + per), col);
          Row8 tot;
#pragma unroll
          for (int q = 0; q < 8; ++q) tot.v[q] = __shfl_sync(0xffffffffu, run.v[q], sub);
#pragma unroll
          for (int g2 = 1; g2 < 8; ++g2) {
            Row8 other;
#pragma unroll
            for (int q = 0; q < 8; ++q) other.v[q] = __shfl_sync(0xffffffffu, run.v[q], 4 * g2 + sub);
            if (g2 * per < ns) row8_add(tot, other);
          }
          if (grp == i) acc = tot;
        }
      }
      // next free 8-row slot of this CTA: tile tl (in completion order), rows 8 ps .. 8 ps + 7
      uint32_t my = (uint32_t)g;
      if (!a.static_slots) {
        if (lane == 0) asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(my) : "r"(smem_u32(next_slot)) : "memory");
        my = __shfl_sync(0xffffffffu, my, 0);
      }
      const int64_t tl = my / kBfPasses;
      const int ps = (int)(my % kBfPasses), stage = (int)(tl % kBfStages);
      // the tensor core has consumed this stage's previous tile (tile tl - S of this CTA)
      if (tl >= kBfStages) {
        const int64_t tp = tl - kBfStages;
        mbar_wait(bar_done + (int)(tp % (2 * kBfStages)), (uint32_t)(tp / (2 * kBfStages)) & 1u);
      }
      const float pre_v = __uint_as_float(__shfl_sync(0xffffffffu, sc, grp_lane0));
      const float xs_v = __uint_as_float(__shfl_sync(0xffffffffu, sc, grp_lane0 + 1));
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      uint32_t xb = 0;
      {
        const int r = ps * 8 + rslot;
        unsigned char* st = smem + stage * kBfStageB;
        if (finish) {
          if (a.row_scale) {
#pragma unroll
            for (int q = 0; q < 8; ++q) acc.v[q] *= pre_v;
          }
          bf_store_row<true>(st + kBfKDh, st + kBfKDl, st + kBfMN, st + kBfMN + 2 * kBfImg, r, sub, flip, acc);
          bf_store_row<true>(st + kBfKGh, st + kBfKGl, st + kBfMN + kBfImg, st + kBfMN + 3 * kBfImg, r, sub, flip, gyrow);
          Row8 xr;
          {
            const float4 z0 = *reinterpret_cast<const float4*>(zst0), z1 = *reinterpret_cast<const float4*>(zst1);
            xr.v[0] = z0.x; xr.v[1] = z0.y; xr.v[2] = z0.z; xr.v[3] = z0.w;
            xr.v[4] = z1.x; xr.v[5] = z1.y; xr.v[6] = z1.z; xr.v[7] = z1.w;
          }
          if (a.x_scale) {
            const float inv = __frcp_rn(xs_v);
#pragma unroll
            for (int q = 0; q < 8; ++q) xr.v[q] *= inv;
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) xb |= (xr.v[q] > 0.f ? 1u : 0u) << q;
          xb <<= 8 * sub;
          bf_store_row<false>(nullptr, nullptr, st + kBfMXh, st + kBfMXl, r, sub, flip, xr);
        } else {
          // no row in this slot: its MN-major rows must not contribute to the transposed products
          const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int im = 4; im < 10; ++im) {
            *reinterpret_cast<float4*>(st + im * kBfImg + bf_mn_off(r, 2 * sub + flip)) = zero;
            *reinterpret_cast<float4*>(st + im * kBfImg + bf_mn_off(r, 2 * sub + 1 - flip)) = zero;
          }
        }
        xb |= __shfl_xor_sync(0xffffffffu, xb, 1);
        xb |= __shfl_xor_sync(0xffffffffu, xb, 2);
        uint32_t sv = sc;                                  // sub 2: hmask_prev, sub 3: post
        if (sub == 0) sv = (uint32_t)(finish ? rowp : -1);
        if (sub == 1) sv = xb;
        reinterpret_cast<uint32_t*>(scal + (int)(tl % (2 * kBfStages)) * kBfRows + r)[sub] = sv;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this lane's image stores -> async proxy
      __syncwarp();
      uint32_t old = 0;
      if (lane == 0) {
        mbar_arrive(bar_full + (int)(tl % (2 * kBfStages)));
        asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(arrivals + stage)) : "memory");
      }
      old = __shfl_sync(0xffffffffu, old, 0);
      if ((old % kBfPasses) == kBfPasses - 1) {
        // last pass of the tile: wait for the other passes' arrivals (acquire), then the accumulator buffer
        const uint32_t use = (uint32_t)(tl / kBfStages);
        mbar_wait(bar_full + (int)(tl % (2 * kBfStages)), (uint32_t)(tl / (2 * kBfStages)) & 1u);
        if (use >= 1) mbar_wait(bar_tfree + stage, (use - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t tb = tmem + stage * kBfTmemBuf;
          const uint32_t so = (uint32_t)(stage * kBfStageB) >> 4;
          if (want_prev) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t ko = 2 * k;   // 8 tf32 = 32 bytes inside the 128-byte swizzle row
              const uint64_t b1 = dK + ((kBfOffB1 >> 4) + ko), b2 = dK + ((kBfOffB2 >> 4) + ko);
              umma_tf32(tb + 0, dK + (so + (kBfKDh >> 4) + ko), b1, idG64, k > 0);    // dxw_hi [Wt_hi | Wt_lo]
              umma_tf32(tb + 0, dK + (so + (kBfKGh >> 4) + ko), b2, idG64, 1);        // gy_hi  [R_hi | R_lo]
              umma_tf32(tb + 32, dK + (so + (kBfKDl >> 4) + ko), b1, idG32, 1);       // dxw_lo Wt_hi
              umma_tf32(tb + 32, dK + (so + (kBfKGl >> 4) + ko), b2, idG32, 1);       // gy_lo  R_hi
            }
          }
#pragma unroll
          for (int k = 0; k < kBfRows / 8; ++k) {
            const uint32_t ko = (1024 * k) >> 4;    // 8 rows = two 4-row atoms
#ifndef MGCN_BF_NO_T
            umma_tf32(tb + 64, dM + (so + (kBfMN >> 4) + ko), dM + (so + (kBfMXh >> 4) + ko), idT, k > 0);
#endif
#ifndef MGCN_BF_NO_COLSUM
            umma_tf32(tb + 128, dOnes, dM2 + (so + ((kBfMN + kBfImg) >> 4) + ko), idC, k > 0);
#endif
          }
          umma_commit(bar_done + (int)(tl % (2 * kBfStages)));
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------- epilogue -------------------------------
    // warp = (column half, TMEM quarter).  T: TMEM lane 32 quarter + lane = row of [dxw_hi | gy_hi | dxw_lo | gy_lo]^T x,
    // this warp's 16 columns; G: lanes 0..15 of the quarter carry tile rows 16 quarter .. + 15, this warp's 16 columns
    const int quarter = warp & 3, half = warp >> 2;
    float acc_t[16], acc_b[8];
#pragma unroll
    for (int t = 0; t < 16; ++t) acc_t[t] = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) acc_b[t] = 0.f;
    const int r = 16 * quarter + (lane & 15);
    float* stg_y = reinterpret_cast<float*>(smem + kBfOffOut) + warp * (2 * 16 * kBfLdo);   // [16][kBfLdo]
    float* stg_s = stg_y + 16 * kBfLdo;
    int tl = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      const int stage = tl % kBfStages;
      mbar_wait(bar_done + tl % (2 * kBfStages), (uint32_t)(tl / (2 * kBfStages)) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = tmem + stage * kBfTmemBuf + ((uint32_t)(32 * quarter) << 16);
      {
        uint32_t d1[16], d2[16], cs[8];
        bf_tmem_ld16(ta + 64 + 16 * half, d1);
        bf_tmem_ld16(ta + 96 + 16 * half, d2);
        bf_tmem_ld8(ta + 128 + 8 * warp, cs);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int t = 0; t < 16; ++t) acc_t[t] += __uint_as_float(d1[t]) + __uint_as_float(d2[t]);
#pragma unroll
        for (int t = 0; t < 8; ++t) acc_b[t] += __uint_as_float(cs[t]);
      }
      if (want_prev) {
        mbar_wait(bar_full + tl % (2 * kBfStages), (uint32_t)(tl / (2 * kBfStages)) & 1u);   // acquire the scalar ring
        const uint4* ring = scal + (tl % (2 * kBfStages)) * kBfRows;
        const uint4 sc4 = ring[r];
        const int row_a = (int)ring[16 * quarter + (lane & 7)].x, row_b = (int)ring[16 * quarter + 8 + (lane & 7)].x;
        const uint32_t xb = sc4.y >> (16 * half), hb = sc4.z >> (16 * half);
        const float postv = __uint_as_float(sc4.w);
        uint32_t m[16], c1[16];
        bf_tmem_ld16(ta + 0 + 16 * half, m);
        bf_tmem_ld16(ta + 32 + 16 * half, c1);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // the accumulator buffer is in registers: hand it back; the warp barrier also orders the previous tile's reads of
        // the staging rows before this tile's writes
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tfree + stage);
#pragma unroll
        for (int c0 = 0; c0 < 16; c0 += 4) {
          float g[4], sv[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const int c = c0 + t;
            const float gv = __uint_as_float(m[c]) + __uint_as_float(c1[c]);
            g[t] = ((xb >> c) & 1u) ? gv : 0.f;
            sv[t] = ((hb >> c) & 1u) ? g[t] * postv : 0.f;
          }
          if (lane < 16) {
            *reinterpret_cast<float4*>(stg_y + (lane & 15) * kBfLdo + c0) = make_float4(g[0], g[1], g[2], g[3]);
            *reinterpret_cast<float4*>(stg_s + (lane & 15) * kBfLdo + c0) = make_float4(sv[0], sv[1], sv[2], sv[3]);
          }
        }
        __syncwarp();
        // staged half rows -> global: lane j moves the 16-byte chunk j >> 3 of rows j & 7 and 8 + (j & 7)
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int lr = 8 * it + (lane & 7), ch = lane >> 3;
          const int row = it == 0 ? row_a : row_b;
          if (row >= 0) {
            const int64_t o = (int64_t)row * kGH + 16 * half + 4 * ch;
            st_f4_hint(a.gy_prev + o, *reinterpret_cast<const float4*>(stg_y + lr * kBfLdo + 4 * ch), pol);
            st_f4_hint(a.gs_prev + o, *reinterpret_cast<const float4*>(stg_s + lr * kBfLdo + 4 * ch), pol);
          }
        }
      } else {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tfree + stage);
      }
    }
    // per-CTA partials: row = TMEM lane of the transposed products; column sums of gy from lane 0 of every warp
    float* p = a.part_t + ((int64_t)blockIdx.x * 128 + 32 * quarter + lane) * 32 + 16 * half;
#pragma unroll
    for (int t = 0; t < 16; t += 4) *reinterpret_cast<float4*>(p + t) = make_float4(acc_t[t], acc_t[t + 1], acc_t[t + 2], acc_t[t + 3]);
    if (lane == 0) {
      float* pb = a.part_b + (int64_t)blockIdx.x * 64 + 8 * warp;   // [gy_hi 0..31 | gy_lo 0..31]
#pragma unroll
      for (int t = 0; t < 8; ++t) pb[t] = acc_b[t];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

}  // namespace mgcn

using namespace mgcn;

extern "C" int mgcn_gcn_layer_bwd_fused(const mgcn_csr_t* gt, const float* gs, const float* gy, const float* z,
                                        const float* x_scale, const float* row_scale, const float* w,
                                        const float* res_w, const uint32_t* hmask_prev, const float* post, int64_t H,
                                        int static_slots, float* gy_prev, float* gs_prev, float* dw, float* d_res_w, float* d_res_b,
                                        void* workspace, size_t* workspace_bytes, void* stream) {
  MGCN_REQUIRE(workspace_bytes != nullptr && gt != nullptr, MGCN_ERR_NULL);
  MGCN_REQUIRE(H == kGH, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(gt->n_rows >= 0 && gt->n_rows < (int64_t(1) << 31), MGCN_ERR_RANGE);
  const bool hubs = gt->hub_rows && gt->hub_seg0 && gt->hub_count && gt->seg_count && gt->hub_cap > 0 && gt->seg_cap > 0;
  int dev = 0, sms = 0;
  MGCN_CHECK_CUDA(cudaGetDevice(&dev));
  MGCN_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t seg_cap = hubs ? gt->seg_cap : 0, hub_cap = hubs ? gt->hub_cap : 0;
  int64_t t0 = ceil_div(gt->n_rows + seg_cap, kBfRows), t1 = ceil_div(hub_cap, kBfRows);
  const int P0 = (int)(t0 < sms ? (t0 > 0 ? t0 : 1) : sms), P1 = hubs ? (int)(t1 < sms ? t1 : sms) : 0;
  WorkspaceCarver ws(workspace);
  float* partial = ws.take<float>((size_t)seg_cap * kGH);
  float* part_t = ws.take<float>((size_t)(P0 + P1) * 128 * 32);
  float* part_b = ws.take<float>((size_t)(P0 + P1) * 2 * 32);
  if (workspace == nullptr) {
    *workspace_bytes = ws.bytes();
    return MGCN_OK;
  }
  MGCN_REQUIRE(*workspace_bytes >= ws.bytes(), MGCN_ERR_WORKSPACE);
  MGCN_REQUIRE(dw && d_res_w && d_res_b, MGCN_ERR_NULL);
  MGCN_REQUIRE((gy_prev == nullptr) == (gs_prev == nullptr), MGCN_ERR_NULL);
  MGCN_REQUIRE(!gy_prev || hmask_prev, MGCN_ERR_NULL);
  if (gt->n_rows == 0) {
    MGCN_CHECK_CUDA(cudaMemsetAsync(dw, 0, 32 * 32 * 4, static_cast<cudaStream_t>(stream)));
    MGCN_CHECK_CUDA(cudaMemsetAsync(d_res_w, 0, 32 * 32 * 4, static_cast<cudaStream_t>(stream)));
    MGCN_CHECK_CUDA(cudaMemsetAsync(d_res_b, 0, 32 * 4, static_cast<cudaStream_t>(stream)));
    return MGCN_OK;
  }
  MGCN_REQUIRE(gs && gy && z && w && res_w && gt->rowptr && gt->tasks, MGCN_ERR_NULL);
  MGCN_REQUIRE(gt->nnz_cap == 0 || gt->nbr_w, MGCN_ERR_NULL);
  MGCN_REQUIRE((reinterpret_cast<uintptr_t>(gs) & 31u) == 0 && (reinterpret_cast<uintptr_t>(gy) & 31u) == 0, MGCN_ERR_ALIGN);
  MGCN_REQUIRE(aligned16(gt->tasks) && aligned16(z) && aligned16(partial) && (!gy_prev || aligned16(gy_prev)) &&
                   (!gs_prev || aligned16(gs_prev)),
               MGCN_ERR_ALIGN);
  BwdFusedArgs a{};
  a.tasks = reinterpret_cast<const int4*>(gt->tasks);
  a.nbr_w = gt->nbr_w;
  a.seg_count = hubs ? gt->seg_count : nullptr;
  a.hub_rows = gt->hub_rows;
  a.hub_seg0 = gt->hub_seg0;
  a.hub_count = gt->hub_count;
  a.rowptr = gt->rowptr;
  a.hub_threshold = gt->hub_threshold;
  a.gs = gs; a.gy = gy; a.z = z; a.x_scale = x_scale; a.row_scale = row_scale; a.w = w; a.res_w = res_w;
  a.hmask_prev = hmask_prev; a.post = post; a.gy_prev = gy_prev; a.gs_prev = gs_prev;
  a.partial = partial;
  a.part_t = part_t;
  a.part_b = part_b;
  a.n_rows = gt->n_rows;
  a.seg_cap = seg_cap;
  a.hub_cap = hub_cap;
  a.static_slots = static_slots ? 1 : 0;
  MGCN_CHECK_CUDA(cudaFuncSetAttribute(k_gcn_bwd_fused<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBfSmem));
  MGCN_CHECK_CUDA(cudaFuncSetAttribute(k_gcn_bwd_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBfSmem));
  MGCN_LAUNCH(k_gcn_bwd_fused<0>, (unsigned)P0, kBfThreads, kBfSmem, stream, a);
  if (hubs) {
    a.part_t = part_t + (size_t)P0 * 128 * 32;
    a.part_b = part_b + (size_t)P0 * 2 * 32;
    MGCN_LAUNCH(k_gcn_bwd_fused<1>, (unsigned)P1, kBfThreads, kBfSmem, stream, a);
  }
  const int rc = launch_bwd_tc_reduce(part_t, P0 + P1, dw, d_res_w, stream);
  if (rc != MGCN_OK) return rc;
  return launch_reduce_partials(part_b, 2 * (P0 + P1), 32, 32, d_res_b, 0, 1, stream);
}
