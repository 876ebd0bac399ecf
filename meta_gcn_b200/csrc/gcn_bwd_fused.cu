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
// gathered sums go through a small ring in shared memory into the operand images of the tensor core, so a layer's
// backward reads gs (gathered), gy and z and writes gy_prev and gs_prev — 0.9 GB less DRAM traffic per layer at the
// botnet batch, and one launch less.
//
// One persistent CTA per SM, three roles over 64-row tiles of the by-source work order (tile t of a CTA = its passes
// 8 t .. 8 t + 7, a pass = 8 consecutive tasks; the assignment is STATIC, so every sum has a fixed order and the
// results are identical from run to run):
//   gather     kBfGather warps; warp w owns passes w, w + kBfGather, ...: gather-sum of gs (gather.cuh: 4 lanes x
//              LDG.256 per gathered row, sums in edge order), nothing else — the sums go as raw fp32 rows (swizzled
//              16-byte chunks) into slot pass % kBfRing of a ring, one mbarrier pair per slot.  The first version of
//              this kernel let the gather warps also split and store the ten operand images; a pass then spent a
//              third of its time outside the gather and waited for operand stages (2 x 8 passes for 20 warps):
//              1.04 ms.  With the ring a gather warp is never further than its own pass from the next gather.
//   images     8 warps, warp p owns pass p of every tile: the ring rows, the rows' own gy and z (LDG.256, one tile ahead
//              when the ring slot is already full) -> scale, tf32 hi / lo split -> ten images per stage: K-major
//              (interleaved) dxw and gy for the row-local products, SWIZZLE_128B_BASE32B MN-major dxw, gy and x for
//              the transposed ones (tc05.cuh: a 32-bit MN-major operand is only read correctly from that image, a
//              K-major one never); lane 8 qq + j owns row j and 32-byte chunk qq, chunk order (j >> 2), conflict-free
//              in both image types.  Tasks that finish no row (hub segments, padding) contribute zero rows.  Warp
//              (tile % 8) issues the tile's 24 tcgen05.mma from one lane:
//                G  [64 x 64|32]  K-major,  M = 64:  main = dxw_hi Wt_hi + gy_hi R_hi, corrections in their own columns
//                T  [128 x 64]    MN-major, M = 128: [dxw_hi|gy_hi|dxw_lo|gy_lo]^T [x_hi | x_lo], 8 steps of 8 rows
//              Column sums of gy (dr) stay in registers of these warps.
//   epilogue   4 warps: T -> running fp32 registers (RN adds; chains through the TMEM accumulator stay 8 steps long);
//              G (lanes 0..15 of a TMEM quarter carry 16 tile rows) -> both masks, per-target factor -> staged rows ->
//              whole 64-byte row pieces to HBM
//   barriers   ring_full / ring_free [slot]; full[tile % 4] (8 image-warp arrivals), done[tile % 4] (tcgen05.commit),
//              tfree[stage] (epilogue -> issuer).  No wait can be lapped: a ring slot's next phase needs this phase's
//              consumer, done(t + 4) needs done(t + 2) needs the stores that wait for done(t).
// Hub rows (longer than the hub threshold): their segments store partial sums in mode 0; a second launch (mode 1) sums
// the partials per hub row (fixed order) and runs the same tile path.  Per-CTA partials of dW / dR / dr are reduced in a
// fixed order by k_bwd_tc_reduce.  No atomics on data.
//
// MEASURED (B200, botnet batch, scripts/prof_layer.py, profiles/r2_bwd_fused.md): 1.45 ms per layer against 0.48 + 0.61 ms
// for k_agg_flat + k_layer_bwd_tc, with 2.59 GB of DRAM traffic against 3.47 GB.  Correct and deterministic, but not
// faster, so fused.BWD_FUSED is off by default.  Why: the per-row operand-image work (48 tf32 splits, 20 STS.128 per
// lane and pass, ~560 instructions per 8 rows) needs the 16 warps k_layer_bwd_tc gives it; here 8 image warps share
// the SM with 16 gather warps and 4 epilogue warps (28 warps x 72 registers is all the register file holds) and every
// tile waits for its slowest image warp.  The 2-role version before it (20 warps that gather AND build images, tiles
// filled in completion order) ran 1.04 ms but its weight gradients depended on the completion order.
#include "common.cuh"
#include "gather.cuh"
#include "tc05.cuh"

namespace mgcn {

constexpr int kBfRows = 64;                      // rows per tile
constexpr int kBfPasses = kBfRows / 8;           // passes (8 rows) per tile = image warps
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
constexpr int kBfOffB1 = kBfStages * kBfStageB;  // [Wt_hi ; Wt_lo]: 64 x 32, interleaved K-major
constexpr int kBfOffB2 = kBfOffB1 + 8192;        // [R_hi ; R_lo]
#ifndef MGCN_BF_RING
#define MGCN_BF_RING 24
#endif
constexpr int kBfRing = MGCN_BF_RING;            // ring slots (passes): 1 KB of rows + 8 row ids each
constexpr int kBfOffRing = kBfOffB2 + 8192;
constexpr int kBfOffMeta = kBfOffRing + kBfRing * 1024;
constexpr int kBfLdo = 36;                       // floats per staged output row (144 bytes: conflict-free)
constexpr int kBfOffOut = kBfOffMeta + kBfRing * 32;       // [4 epilogue warps][gy_prev | gs_prev][16 rows][kBfLdo]
#ifndef MGCN_BF_GATHER
#define MGCN_BF_GATHER 16
#endif
constexpr int kBfGather = MGCN_BF_GATHER;
constexpr int kBfImgWarps = kBfPasses;
constexpr int kBfEpiWarps = 4;
constexpr int kBfThreads = 32 * (kBfEpiWarps + kBfImgWarps + kBfGather);
constexpr int kBfOffScal = kBfOffOut + kBfEpiWarps * 2 * 16 * kBfLdo * 4;   // [tile % 4][64] {row id, x > 0 bits, hmask_prev, post}
constexpr int kBfOffDr = kBfOffScal + 2 * kBfStages * kBfRows * 16;         // [8 image warps][32]
constexpr int kBfOffMisc = kBfOffDr + kBfImgWarps * 32 * 4;                 // barriers, tmem slot
constexpr int kBfSmem = kBfOffMisc + 1024 + 1024;
constexpr int kBfTmemBuf = 128;                  // columns per accumulator buffer: G [0,64), T [64,128)
static_assert(kBfSmem <= 232448, "shared memory");
static_assert((10 + 2 * kBfRing) * 8 + 16 <= 1024, "barrier block");
static_assert(kBfRing >= 16, "the image warps look one tile ahead");

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
  float* part_b;             // [grid][32]
  int64_t n_rows;
  int64_t seg_cap;
  int64_t hub_cap;
  int hub_threshold;
};

__device__ __forceinline__ int bf_k_off(int r, int q) {       // bytes; 16-byte chunk q of row r, interleaved K-major
  return ((r >> 3) << 10) + (q << 7) + ((r & 7) << 4);
}
__device__ __forceinline__ int bf_mn_off(int r, int q) {      // bytes; SWIZZLE_128B_BASE32B
  return (r << 7) + ((((q >> 1) ^ (r & 3)) << 5) | ((q & 1) << 4));
}

__device__ __forceinline__ void bf_split4(const float4 v, float4& hi, float4& lo) {
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

struct BfF8 {
  float4 lo, hi;
};
__device__ __forceinline__ BfF8 bf_ld_f8_stream(const float* p, uint64_t pol) {   // one LDG.256, streamed
  BfF8 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
               : "=f"(r.lo.x), "=f"(r.lo.y), "=f"(r.lo.z), "=f"(r.lo.w), "=f"(r.hi.x), "=f"(r.hi.y), "=f"(r.hi.z), "=f"(r.hi.w)
               : "l"(p), "l"(pol));
  return r;
}

__device__ __forceinline__ bool bf_mbar_test(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  return ok != 0;
}

// mbarrier wait that does not hammer the issue slots and the shared-memory pipe while it waits: the first version of
// this kernel spent half of its executed instructions in try_wait loops of warps that had nothing to do (ncu: 190 M
// polls per launch).  try_wait with a suspend-time hint parks the warp in hardware; a failed attempt sleeps `ns` more.
#ifndef MGCN_BF_SLEEP_G
#define MGCN_BF_SLEEP_G 200   // gather warps waiting for a ring slot
#endif
#ifndef MGCN_BF_SLEEP_I
#define MGCN_BF_SLEEP_I 32    // image warps (stage, issuer)
#endif
#ifndef MGCN_BF_SLEEP_E
#define MGCN_BF_SLEEP_E 64    // epilogue
#endif
#ifndef MGCN_BF_HINT
#define MGCN_BF_HINT 2000
#endif
__device__ __forceinline__ void bf_wait(uint64_t* b, uint32_t parity, unsigned ns) {
  uint32_t ok;
  for (;;) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity), "r"((uint32_t)MGCN_BF_HINT) : "memory");
    if (ok) break;
    if (ns) __nanosleep(ns);
  }
}

__device__ __forceinline__ void bf_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

template <int kMode>   // 0: tiles of the work order (rows + hub segments), 1: tiles of the hub list
__global__ void __launch_bounds__(kBfThreads, 1) k_gcn_bwd_fused(const BwdFusedArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_done = reinterpret_cast<uint64_t*>(smem + kBfOffMisc);   // [4]
  uint64_t* bar_full = bar_done + 4;                                      // [4] 8 image-warp arrivals per tile
  uint64_t* bar_tfree = bar_full + 4;                                     // [2]
  uint64_t* ring_full = bar_tfree + 2;                                    // [kBfRing]
  uint64_t* ring_free = ring_full + kBfRing;                              // [kBfRing]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ring_free + kBfRing);
  uint4* scal = reinterpret_cast<uint4*>(smem + kBfOffScal);
  int32_t* meta = reinterpret_cast<int32_t*>(smem + kBfOffMeta);          // [kBfRing][8] row ids (-1: no row)
  float* dr_red = reinterpret_cast<float*>(smem + kBfOffDr);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const bool want_prev = a.gy_prev != nullptr;

  // weight images (interleaved K-major, 64 rows: hi then lo): B1(n, k) = W[n][k], B2(n, k) = R[k][n]
  for (int i = tid; i < 32 * 32; i += kBfThreads) {
    const int n = i >> 5, k = i & 31;
    const float w1 = __ldg(a.w + n * 32 + k), w2 = __ldg(a.res_w + k * 32 + n);
    const float h1 = __uint_as_float(round_tf32_bits(__float_as_uint(w1)));
    const float h2 = __uint_as_float(round_tf32_bits(__float_as_uint(w2)));
    float* b1 = reinterpret_cast<float*>(smem + kBfOffB1);
    float* b2 = reinterpret_cast<float*>(smem + kBfOffB2);
    const int o_hi = (bf_k_off(n, k >> 2) >> 2) + (k & 3), o_lo = (bf_k_off(n + 32, k >> 2) >> 2) + (k & 3);
    b1[o_hi] = h1;
    b1[o_lo] = __uint_as_float(round_tf32_bits(__float_as_uint(w1 - h1)));
    b2[o_hi] = h2;
    b2[o_lo] = __uint_as_float(round_tf32_bits(__float_as_uint(w2 - h2)));
  }
  if (tid == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(bar_done + s, 1);
      mbar_init(bar_full + s, kBfImgWarps);
    }
    for (int s = 0; s < kBfStages; ++s) mbar_init(bar_tfree + s, kBfEpiWarps);
    for (int s = 0; s < kBfRing; ++s) {
      mbar_init(ring_full + s, 1);
      mbar_init(ring_free + s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
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
  const int my_tiles = blockIdx.x < n_tiles ? (int)((n_tiles - 1 - blockIdx.x) / gridDim.x + 1) : 0;

  if (warp >= kBfEpiWarps + kBfImgWarps) {
    // ------------------------------- gather warps -------------------------------
    const int gw = warp - kBfEpiWarps - kBfImgWarps;
    const int sub = lane & 3, grp = lane >> 2, grp_lane0 = grp * 4;
    const unsigned gmask = 0xfu << grp_lane0;
    const int col = sub * 8;
    // ring row grp of a slot, 16-byte chunk c at position c ^ grp: conflict-free for these stores (a quarter warp = two
    // rows of opposite parity) and for the image warps' reads (8 rows, one chunk)
    const int ring_o0 = grp * 128 + (((2 * sub) ^ grp) << 4), ring_o1 = grp * 128 + (((2 * sub + 1) ^ grp) << 4);
    auto load_desc = [&](int g) {
      int4 d = make_int4(-1, 0, 0, 0);
      const int64_t tile = blockIdx.x + (int64_t)(g / kBfPasses) * gridDim.x;
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
    int4 d = load_desc(gw);
    int4 dn = load_desc(gw + kBfGather);
    PassIdx pi = load_idx(d);
    const int n_pass = my_tiles * kBfPasses;
    for (int g = gw; g < n_pass; g += kBfGather) {
      const int rowp = __shfl_sync(0xffffffffu, d.x, grp);
      const int slot = __shfl_sync(0xffffffffu, d.w, grp);
      const bool finish = rowp >= 0 && slot == 0;   // this group completes a row
      const PassIdx pc = pi;
      const int4 dc = d;
      d = dn;
      pi = load_idx(d);                               // next pass: index batches in flight during this gather
      dn = load_desc(g + 2 * kBfGather);
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
      // ring slot of this pass: free once the image warp has read the pass that used it before
      const int q = g % kBfRing, use = g / kBfRing;
      if (use >= 1) bf_wait(ring_free + q, (uint32_t)(use - 1) & 1u, MGCN_BF_SLEEP_G);
      unsigned char* slot_p = smem + kBfOffRing + q * 1024;
      *reinterpret_cast<float4*>(slot_p + ring_o0) = make_float4(acc.v[0], acc.v[1], acc.v[2], acc.v[3]);
      *reinterpret_cast<float4*>(slot_p + ring_o1) = make_float4(acc.v[4], acc.v[5], acc.v[6], acc.v[7]);
      if (sub == 0) meta[q * 8 + grp] = finish ? rowp : -1;
      __syncwarp();
      if (lane == 0) mbar_arrive(ring_full + q);
    }
  } else if (warp >= kBfEpiWarps) {
    // ------------------------------- image warps -------------------------------
    // lane 8 qq + j owns row j of the warp's pass and the 32-byte chunk qq; the two 16-byte chunks are stored in the
    // order (j >> 2): the 8 lanes of a quarter warp hit 8 different 16-byte bank groups in the K-major images (bank
    // group = row & 7) and in the swizzled MN-major ones (bank group = (((q >> 1) ^ (row & 3)) << 1) | (q & 1))
    const int pw = warp - kBfEpiWarps;
    const int j = lane & 7, qq = lane >> 3, flip = j >> 2;
    const int r0 = 8 * pw + j;
    const int qa = 2 * qq + flip, qb = 2 * qq + (flip ^ 1);
    const int ko_a = bf_k_off(r0, qa), ko_b = bf_k_off(r0, qb);
    const int mo_a = bf_mn_off(r0, qa), mo_b = bf_mn_off(r0, qb);
    const int ring_a = j * 128 + ((qa ^ j) << 4), ring_b = j * 128 + ((qb ^ j) << 4);
    const uint32_t idG64 = umma_idesc_tf32(64, 64), idG32 = umma_idesc_tf32(64, 32), idT = umma_idesc_tf32(128, 64, 1, 1);
    const uint64_t dK = umma_desc(smem_u32(smem), 128, 1024, 0), dM = umma_desc(smem_u32(smem), kBfImg, 512, 1);
    float acc_dr[2][4];   // column sums of gy over this thread's rows
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int t = 0; t < 4; ++t) acc_dr[i][t] = 0.f;
    // own rows of a pass: row id from the ring's meta words, gy and z chunks, one scalar per chunk lane
    // (qq = 0: row_scale, 1: x_scale, 2: hmask_prev, 3: post)
    struct Own {
      int row;
      uint32_t sc;
      BfF8 gy, z;
    };
    auto load_own = [&](int tl) {
      Own o;
      o.row = meta[((tl * kBfPasses + pw) % kBfRing) * 8 + j];
      const int64_t off = (int64_t)(o.row >= 0 ? o.row : 0) * kGH + 8 * qq;
      o.gy = bf_ld_f8_stream(a.gy + off, pol);
      o.z = bf_ld_f8_stream(a.z + off, pol);
      o.sc = qq == 2 ? 0u : 0x3f800000u;
      const void* sp = qq == 0 ? (const void*)a.row_scale : qq == 1 ? (const void*)a.x_scale
                       : qq == 2 ? (const void*)a.hmask_prev : (const void*)a.post;
      if (o.row >= 0 && sp) o.sc = __ldg(reinterpret_cast<const uint32_t*>(sp) + o.row);
      return o;
    };
    auto slot_parity = [&](int tl) { return (uint32_t)((tl * kBfPasses + pw) / kBfRing) & 1u; };
    Own cur, nxt;
    if (my_tiles > 0) {
      bf_wait(ring_full + (pw % kBfRing), 0, MGCN_BF_SLEEP_I);
      cur = load_own(0);
    }
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int gp = tl * kBfPasses + pw, q = gp % kBfRing;
      const int stage = tl & 1;
      // next tile's pass: its own rows are requested now if the gather warps are already there
      bool have_nxt = false;
      if (tl + 1 < my_tiles) {
        have_nxt = bf_mbar_test(ring_full + ((gp + kBfPasses) % kBfRing), slot_parity(tl + 1));
        if (have_nxt) nxt = load_own(tl + 1);
      }
      // the ring rows of this pass, then the slot goes back to the gather warps
      const unsigned char* slot_p = smem + kBfOffRing + q * 1024;
      float4 da = *reinterpret_cast<const float4*>(slot_p + ring_a), db = *reinterpret_cast<const float4*>(slot_p + ring_b);
      __syncwarp();
      if (lane == 0) mbar_arrive(ring_free + q);
      const bool valid = cur.row >= 0;
      const float pre_v = __uint_as_float(__shfl_sync(0xffffffffu, cur.sc, j));        // lane j: qq = 0
      const float xs_v = __uint_as_float(__shfl_sync(0xffffffffu, cur.sc, 8 + j));     // qq = 1
      const uint32_t hm_v = __shfl_sync(0xffffffffu, cur.sc, 16 + j);
      const uint32_t post_v = __shfl_sync(0xffffffffu, cur.sc, 24 + j);
      if (!valid) {
        da = db = make_float4(0.f, 0.f, 0.f, 0.f);
        cur.gy.lo = cur.gy.hi = cur.z.lo = cur.z.hi = da;
      } else if (a.row_scale) {
        da.x *= pre_v; da.y *= pre_v; da.z *= pre_v; da.w *= pre_v;
        db.x *= pre_v; db.y *= pre_v; db.z *= pre_v; db.w *= pre_v;
      }
      // the tensor core has consumed this stage's previous tile
      if (tl >= 2) bf_wait(bar_done + ((tl - 2) & 3), (uint32_t)((tl - 2) >> 2) & 1u, MGCN_BF_SLEEP_I);
      unsigned char* st = smem + stage * kBfStageB;
      float4 hi, lo;
      // dxw (da = chunk qa, db = chunk qb)
      bf_split4(da, hi, lo);
      *reinterpret_cast<float4*>(st + kBfKDh + ko_a) = hi;
      *reinterpret_cast<float4*>(st + kBfKDl + ko_a) = lo;
      *reinterpret_cast<float4*>(st + kBfMN + 0 * kBfImg + mo_a) = hi;
      *reinterpret_cast<float4*>(st + kBfMN + 2 * kBfImg + mo_a) = lo;
      bf_split4(db, hi, lo);
      *reinterpret_cast<float4*>(st + kBfKDh + ko_b) = hi;
      *reinterpret_cast<float4*>(st + kBfKDl + ko_b) = lo;
      *reinterpret_cast<float4*>(st + kBfMN + 0 * kBfImg + mo_b) = hi;
      *reinterpret_cast<float4*>(st + kBfMN + 2 * kBfImg + mo_b) = lo;
      // gy
      {
        const float4 g0 = cur.gy.lo, g1 = cur.gy.hi;
        acc_dr[0][0] += g0.x; acc_dr[0][1] += g0.y; acc_dr[0][2] += g0.z; acc_dr[0][3] += g0.w;
        acc_dr[1][0] += g1.x; acc_dr[1][1] += g1.y; acc_dr[1][2] += g1.z; acc_dr[1][3] += g1.w;
      }
      bf_split4(flip ? cur.gy.hi : cur.gy.lo, hi, lo);
      *reinterpret_cast<float4*>(st + kBfKGh + ko_a) = hi;
      *reinterpret_cast<float4*>(st + kBfKGl + ko_a) = lo;
      *reinterpret_cast<float4*>(st + kBfMN + 1 * kBfImg + mo_a) = hi;
      *reinterpret_cast<float4*>(st + kBfMN + 3 * kBfImg + mo_a) = lo;
      bf_split4(flip ? cur.gy.lo : cur.gy.hi, hi, lo);
      *reinterpret_cast<float4*>(st + kBfKGh + ko_b) = hi;
      *reinterpret_cast<float4*>(st + kBfKGl + ko_b) = lo;
      *reinterpret_cast<float4*>(st + kBfMN + 1 * kBfImg + mo_b) = hi;
      *reinterpret_cast<float4*>(st + kBfMN + 3 * kBfImg + mo_b) = lo;
      // x = z / x_scale
      if (a.x_scale) {
        const float inv = __frcp_rn(xs_v);
        BfF8& v = cur.z;
        v.lo.x *= inv; v.lo.y *= inv; v.lo.z *= inv; v.lo.w *= inv;
        v.hi.x *= inv; v.hi.y *= inv; v.hi.z *= inv; v.hi.w *= inv;
      }
      uint32_t xb;
      {
        const float4 x0 = cur.z.lo, x1 = cur.z.hi;
        xb = (x0.x > 0.f ? 1u : 0u) | (x0.y > 0.f ? 2u : 0u) | (x0.z > 0.f ? 4u : 0u) | (x0.w > 0.f ? 8u : 0u) |
             (x1.x > 0.f ? 16u : 0u) | (x1.y > 0.f ? 32u : 0u) | (x1.z > 0.f ? 64u : 0u) | (x1.w > 0.f ? 128u : 0u);
        xb <<= 8 * qq;
      }
      bf_split4(flip ? cur.z.hi : cur.z.lo, hi, lo);
      *reinterpret_cast<float4*>(st + kBfMXh + mo_a) = hi;
      *reinterpret_cast<float4*>(st + kBfMXl + mo_a) = lo;
      bf_split4(flip ? cur.z.lo : cur.z.hi, hi, lo);
      *reinterpret_cast<float4*>(st + kBfMXh + mo_b) = hi;
      *reinterpret_cast<float4*>(st + kBfMXl + mo_b) = lo;
      xb |= __shfl_xor_sync(0xffffffffu, xb, 8);
      xb |= __shfl_xor_sync(0xffffffffu, xb, 16);
      // scalar ring (slot tile % 4: its previous tile was drained by the epilogue before MMA(tl - 2) was issued)
      if (qq == 0) scal[(tl & 3) * kBfRows + r0] = make_uint4((uint32_t)cur.row, xb, hm_v, post_v);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + (tl & 3));
      if (pw == (tl & 7)) {
        // this warp issues the tile's 24 tcgen05.mma (lane 0) once all 8 image warps have arrived and the epilogue has
        // drained the accumulator buffer
        bf_wait(bar_full + (tl & 3), (uint32_t)(tl >> 2) & 1u, MGCN_BF_SLEEP_I);
        if (tl >= 2) bf_wait(bar_tfree + stage, (uint32_t)((tl >> 1) - 1) & 1u, MGCN_BF_SLEEP_I);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t tb = tmem + stage * kBfTmemBuf;
          const uint32_t so = (uint32_t)(stage * kBfStageB) >> 4;
          if (want_prev) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t ko = (256 * k) >> 4;   // 8 columns = two 16-byte chunks
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
            umma_tf32(tb + 64, dM + (so + (kBfMN >> 4) + ko), dM + (so + (kBfMXh >> 4) + ko), idT, k > 0);
          }
          umma_commit(bar_done + (tl & 3));
        }
        __syncwarp();
      }
      if (tl + 1 < my_tiles) {
        if (!have_nxt) {
          bf_wait(ring_full + ((gp + kBfPasses) % kBfRing), slot_parity(tl + 1), MGCN_BF_SLEEP_I);
          nxt = load_own(tl + 1);
        }
        cur = nxt;
      }
    }
    // dr[c]: this thread holds columns 8 qq + 4 i + t summed over its rows; rows of the warp are added in a fixed
    // butterfly order, the 8 warps by one thread per column at the end of the kernel
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float v = acc_dr[i][t];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (j == 0) dr_red[pw * 32 + 8 * qq + 4 * i + t] = v;
      }
  } else {
    // ------------------------------- epilogue -------------------------------
    // T: TMEM lane 32 warp + lane = row of [dxw_hi | gy_hi | dxw_lo | gy_lo]^T x; G: lanes 0..15 of quarter `warp` carry
    // tile rows 16 warp .. 16 warp + 15
    float acc_t[32];
#pragma unroll
    for (int t = 0; t < 32; ++t) acc_t[t] = 0.f;
    const int r = 16 * warp + (lane & 15);
    float* stg_y = reinterpret_cast<float*>(smem + kBfOffOut) + warp * (2 * 16 * kBfLdo);   // [16][kBfLdo]
    float* stg_s = stg_y + 16 * kBfLdo;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int stage = tl & 1;
      bf_wait(bar_done + (tl & 3), (uint32_t)(tl >> 2) & 1u, MGCN_BF_SLEEP_E);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t ta = tmem + stage * kBfTmemBuf + ((uint32_t)(32 * warp) << 16);
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 8) {
        uint32_t d1[8], d2[8];
        bf_tmem_ld8(ta + 64 + c0, d1);
        bf_tmem_ld8(ta + 96 + c0, d2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int t = 0; t < 8; ++t) acc_t[c0 + t] += __uint_as_float(d1[t]) + __uint_as_float(d2[t]);
      }
      if (want_prev) {
        mbar_wait(bar_full + (tl & 3), (uint32_t)(tl >> 2) & 1u);   // acquire the scalar ring
        const uint4* ring = scal + (tl & 3) * kBfRows;
        const uint4 sc4 = ring[r];
        const int row_a = (int)ring[16 * warp + (lane & 7)].x, row_b = (int)ring[16 * warp + 8 + (lane & 7)].x;
        const uint32_t xb = sc4.y, hb = sc4.z;
        const float postv = __uint_as_float(sc4.w);
        __syncwarp();   // the previous tile's staged rows have been read by every lane
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 8) {
          uint32_t m[8], c1[8];
          bf_tmem_ld8(ta + 0 + c0, m);
          bf_tmem_ld8(ta + 32 + c0, c1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float g[8], sv[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int c = c0 + t;
            const float gv = __uint_as_float(m[t]) + __uint_as_float(c1[t]);
            g[t] = ((xb >> c) & 1u) ? gv : 0.f;
            sv[t] = ((hb >> c) & 1u) ? g[t] * postv : 0.f;
          }
          if (lane < 16) {
            *reinterpret_cast<float4*>(stg_y + lane * kBfLdo + c0) = make_float4(g[0], g[1], g[2], g[3]);
            *reinterpret_cast<float4*>(stg_y + lane * kBfLdo + c0 + 4) = make_float4(g[4], g[5], g[6], g[7]);
            *reinterpret_cast<float4*>(stg_s + lane * kBfLdo + c0) = make_float4(sv[0], sv[1], sv[2], sv[3]);
            *reinterpret_cast<float4*>(stg_s + lane * kBfLdo + c0 + 4) = make_float4(sv[4], sv[5], sv[6], sv[7]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tfree + stage);
        // staged rows -> global: lane l moves the 16-byte chunks (l >> 3) and 4 + (l >> 3) of rows l & 7 and 8 + (l & 7)
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int lr = 8 * (it & 1) + (lane & 7), ch = (lane >> 3) + 4 * (it >> 1);
          const int row = (it & 1) ? row_b : row_a;
          if (row >= 0) {
            const int64_t o = (int64_t)row * kGH + 4 * ch;
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
    // per-CTA partials of the transposed products: row = TMEM lane
    float* p = a.part_t + ((int64_t)blockIdx.x * 128 + 32 * warp + lane) * 32;
#pragma unroll
    for (int t = 0; t < 32; t += 4) *reinterpret_cast<float4*>(p + t) = make_float4(acc_t[t], acc_t[t + 1], acc_t[t + 2], acc_t[t + 3]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) {
    float s = 0.f;
    if (my_tiles > 0)
      for (int w = 0; w < kBfImgWarps; ++w) s += dr_red[w * 32 + tid];
    a.part_b[(int64_t)blockIdx.x * 32 + tid] = s;
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
  }
}

}  // namespace mgcn

using namespace mgcn;

extern "C" int mgcn_gcn_layer_bwd_fused(const mgcn_csr_t* gt, const float* gs, const float* gy, const float* z,
                                        const float* x_scale, const float* row_scale, const float* w,
                                        const float* res_w, const uint32_t* hmask_prev, const float* post, int64_t H,
                                        float* gy_prev, float* gs_prev, float* dw, float* d_res_w, float* d_res_b,
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
  float* part_b = ws.take<float>((size_t)(P0 + P1) * 32);
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
  MGCN_REQUIRE((reinterpret_cast<uintptr_t>(gs) & 31u) == 0 && (reinterpret_cast<uintptr_t>(gy) & 31u) == 0 &&
                   (reinterpret_cast<uintptr_t>(z) & 31u) == 0,
               MGCN_ERR_ALIGN);   // 256-bit row loads
  MGCN_REQUIRE(aligned16(gt->tasks) && aligned16(partial) && (!gy_prev || aligned16(gy_prev)) &&
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
  MGCN_CHECK_CUDA(cudaFuncSetAttribute(k_gcn_bwd_fused<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBfSmem));
  MGCN_CHECK_CUDA(cudaFuncSetAttribute(k_gcn_bwd_fused<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBfSmem));
  MGCN_LAUNCH(k_gcn_bwd_fused<0>, (unsigned)P0, kBfThreads, kBfSmem, stream, a);
  if (hubs) {
    a.part_t = part_t + (size_t)P0 * 128 * 32;
    a.part_b = part_b + (size_t)P0 * 32;
    MGCN_LAUNCH(k_gcn_bwd_fused<1>, (unsigned)P1, kBfThreads, kBfSmem, stream, a);
  }
  return launch_bwd_tc_reduce(part_t, P0 + P1, dw, d_res_w, part_b, P0 + P1, d_res_b, stream);
}
