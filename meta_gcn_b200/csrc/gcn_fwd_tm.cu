// k_gcn_fwd_tm: the forward layer of gcn_fwd_tc.cu with the A operands of its products in TENSOR MEMORY.
//
// Same contract as k_gcn_fwd_tc (mgcn_gcn_layer_fwd_tc: aggregate-then-transform, z = in_scale (.) x stored):
//     s = post * sum_gathered z;  h = relu(s W + bias);  y = h + (z R^T)/in_scale + r;  z' = out_scale * act(y)
// What changes is how the gathered sums reach the tensor core.  The layer kernels are bound by the LSU data pipe
// (DESIGN §3.4); in k_gcn_fwd_tc every row costs 4 shared-memory image stores (s hi / lo, z hi / lo) on top of its
// gathers.  Here a 4-lane group keeps the sums of TWO passes (rows g and g + 8 of a 16-row block) in registers and
// one `tcgen05.st.sync.aligned.16x256b.x4` per operand writes the block straight into tensor memory; the products are
// issued as `tcgen05.mma [d], [a_tmem], b_desc` (A from TMEM, weights from shared memory).  No operand image touches
// shared memory, the LSU pipe keeps only the gathers, and the L1 keeps ~180 KB instead of ~60 KB.
//   * register layout: lane (g, q) holds columns [8q, 8q+8) of its rows; register 4 kb + {0,1} of the store = columns
//     8q + 2kb + {0,1} of row g, 4 kb + {2,3} the same of row g + 8.  The tensor core therefore sees the contraction
//     index permuted (position 8 kb + 2 q + e holds column 8 q + 2 kb + e); the weight images are written with the
//     same permutation.  Conventions verified by scripts/tc_probe_ts.cu (profiles/r2_tc_probe_ts.log).
//   * a warp may only touch the TMEM lanes of its quarter (warp % 4), so a tile's 128 rows = 4 quarters x 2 blocks of 16
//     lanes; producer warps take blocks of THEIR quarter from a per-quarter counter, in completion order (see
//     gcn_fwd_tc.cu for why), and the double passes of the CTA are dealt to the quarters round robin so that every
//     quarter fills exactly two blocks per tile.
//   * TMEM map per stage (256 columns, 2 stages = all 512): A operands s_hi 0, s_lo 32, z_hi 64, z_lo 96;
//     accumulators D1 128 (main) / 160 (corrections), D2 192 / 224.
#include "common.cuh"
#include "gather.cuh"
#include "tc05.cuh"

namespace mgcn {

constexpr int kTmRows = 128;
constexpr int kTmStages = 2;
constexpr int kTmOffB1 = 0;                      // B1(n, kappa) = W[c(kappa)][n]: rows 0..31 hi, 32..63 lo; 64 x 128 B
constexpr int kTmOffB2 = 8192;                   // B2(n, kappa) = R[n][c(kappa)]
constexpr int kTmLdo = 36;
constexpr int kTmOffOut = 16384;                 // [128][144 B] staged output rows
constexpr int kTmOffVec = kTmOffOut + kTmRows * kTmLdo * 4;
constexpr int kTmOffScal = kTmOffVec + 256;      // [tile % 2S][128] {post, 1 / in_scale, out_scale, row id}
constexpr int kTmOffMisc = kTmOffScal + 2 * kTmStages * kTmRows * 16;
constexpr int kTmSmem = kTmOffMisc + 256 + 1024;
constexpr int kTmEpiWarps = 4;
#ifndef MGCN_TM_PROD
#define MGCN_TM_PROD 20
#endif
constexpr int kTmProdWarps = MGCN_TM_PROD;       // a multiple of 4: kTmPerQ per TMEM quarter
constexpr int kTmPerQ = kTmProdWarps / 4;
static_assert(kTmProdWarps % 4 == 0, "producer warps per quarter");
constexpr int kTmThreads = 32 * (kTmEpiWarps + kTmProdWarps);
constexpr int kTmStageCols = 256;
#ifndef MGCN_TM_L1
#define MGCN_TM_L1 1   // gathered rows allocate in L1 (this kernel leaves it ~180 KB); 0: bypass as the other kernels do
#endif

struct FwdTmArgs {
  const int4* tasks;
  const int32_t* nbr_w;
  const int32_t* seg_count;
  const int32_t* hub_rows;
  const int32_t* hub_seg0;
  const int32_t* hub_count;
  const int32_t* rowptr;
  const float* z;
  const float* w;
  const float* res_w;
  const float* res_b;
  const float* bias;
  const float* in_scale;
  const float* post;
  const float* out_scale;
  float* z_next;
  uint32_t* hmask;
  float* partial;
  int64_t n_rows;
  int64_t seg_cap;
  int64_t hub_cap;
  int act_out;
  int hub_threshold;
};

__device__ __forceinline__ int tm_sw128_off(int r, int q) { return (r << 7) + ((q ^ (r & 7)) << 4); }
__device__ __forceinline__ int tm_col_of(int kappa) {   // physical column of contraction position kappa
  return 8 * ((kappa >> 1) & 3) + 2 * (kappa >> 3) + (kappa & 1);
}

__device__ __forceinline__ void tm_tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

// rows a (g) and b (g + 8) of a 16-row block, this lane's 8 columns each -> hi and lo operand blocks in TMEM
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
// (the values are turned into their lo parts in place after the hi block has been issued: 16 temporaries, not 32)
__device__ __forceinline__ void tm_store_pair(uint32_t t_hi, uint32_t t_lo, Row8& a, Row8& b) {
  uint32_t v[16];
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      v[4 * kb + e] = round_tf32_bits(__float_as_uint(a.v[2 * kb + e]));
      v[4 * kb + 2 + e] = round_tf32_bits(__float_as_uint(b.v[2 * kb + e]));
      a.v[2 * kb + e] -= __uint_as_float(v[4 * kb + e]);
      b.v[2 * kb + e] -= __uint_as_float(v[4 * kb + 2 + e]);
    }
  }
  tm_st16(t_hi, v);
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      v[4 * kb + e] = round_tf32_bits(__float_as_uint(a.v[2 * kb + e]));
      v[4 * kb + 2 + e] = round_tf32_bits(__float_as_uint(b.v[2 * kb + e]));
    }
  }
  tm_st16(t_lo, v);
}

__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, int acc) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
      " tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}

template <int kMode>   // 0: double passes over the work order (rows + hub segments), 1: over the hub list
__global__ void __launch_bounds__(kTmThreads, 1) k_gcn_fwd_tm(const FwdTmArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar_done = reinterpret_cast<uint64_t*>(smem + kTmOffMisc);      // [2 S]
  uint64_t* bar_tfree = bar_done + 2 * kTmStages;                            // [S]
  uint64_t* bar_full = bar_tfree + kTmStages;                                // [2 S] 8 block arrivals per tile
  uint32_t* arrivals = reinterpret_cast<uint32_t*>(bar_full + 2 * kTmStages);   // [S]
  uint32_t* next_block = arrivals + kTmStages;                               // [4] blocks taken per TMEM quarter
  uint32_t* tmem_slot = next_block + 4;
  float* vec = reinterpret_cast<float*>(smem + kTmOffVec);
  float4* scal = reinterpret_cast<float4*>(smem + kTmOffScal);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

  // weight images, contraction index permuted as the TMEM operands are (tm_col_of)
  for (int i = tid; i < 32 * 32; i += kTmThreads) {
    const int n = i >> 5, kappa = i & 31;
    const int c = tm_col_of(kappa);
    const float w1 = __ldg(a.w + c * 32 + n), w2 = __ldg(a.res_w + n * 32 + c);
    const float h1 = __uint_as_float(round_tf32_bits(__float_as_uint(w1)));
    const float h2 = __uint_as_float(round_tf32_bits(__float_as_uint(w2)));
    float* b1 = reinterpret_cast<float*>(smem + kTmOffB1);
    float* b2 = reinterpret_cast<float*>(smem + kTmOffB2);
    const int o_hi = (tm_sw128_off(n, kappa >> 2) >> 2) + (kappa & 3), o_lo = (tm_sw128_off(n + 32, kappa >> 2) >> 2) + (kappa & 3);
    b1[o_hi] = h1;
    b1[o_lo] = __uint_as_float(round_tf32_bits(__float_as_uint(w1 - h1)));
    b2[o_hi] = h2;
    b2[o_lo] = __uint_as_float(round_tf32_bits(__float_as_uint(w2 - h2)));
  }
  if (tid < 32) {
    vec[tid] = a.res_b ? __ldg(a.res_b + tid) : 0.f;
    vec[32 + tid] = a.bias ? __ldg(a.bias + tid) : 0.f;
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < kTmStages; ++s) {
      arrivals[s] = 0;
      mbar_init(bar_done + s, 1);
      mbar_init(bar_done + kTmStages + s, 1);
      mbar_init(bar_tfree + s, kTmEpiWarps);
      mbar_init(bar_full + s, 8);
      mbar_init(bar_full + kTmStages + s, 8);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) next_block[q] = 0;
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

  int64_t limit;
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
  const int n_tiles = (int)((limit + kTmRows - 1) / kTmRows);
  const int my_tiles = (int)blockIdx.x < n_tiles ? (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
  const int limit32 = (int)limit;

  if (warp >= kTmEpiWarps) {
    // ------------------------------- producers -------------------------------
    const int pw = warp - kTmEpiWarps;
    const int qtr = warp & 3;            // TMEM quarter of this warp (== pw & 3)
    const int qi = pw >> 2;              // 0..4: index among the quarter's producer warps
    const uint32_t id64 = umma_idesc_tf32(kTmRows, 64), id32 = umma_idesc_tf32(kTmRows, 32);
    const uint64_t dsc = umma_desc(smem_u32(smem), 16, 1024, 2);
    const int sub = lane & 3, grp = lane >> 2, grp_lane0 = grp * 4;
    const unsigned gmask = 0xfu << grp_lane0;
    const int col = sub * 8;
    // double pass number j of this quarter covers tasks [16 dp, 16 dp + 16) of the CTA's tiles, dp = 4 j + qtr
    // descriptors: lanes 0..15 hold one entry each (lanes 0..7 pass A, 8..15 pass B)
    auto load_desc = [&](int j) {
      int4 d = make_int4(-1, 0, 0, 0);
      const int dp = 4 * j + qtr;
      const int tile = (int)blockIdx.x + (dp >> 3) * (int)gridDim.x;
      const int e = tile * kTmRows + (dp & 7) * 16 + lane;       // < 2^31: N + segments < 2^31 - 2^20
      if (lane < 16 && j < 2 * my_tiles && e < limit32) {
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
    auto load_idx = [&](const int4& d, int half) {
      PassIdx p{0, 0, 0, 0};
      if (kMode == 0) {
        p.beg = __shfl_sync(0xffffffffu, d.y, 8 * half + grp);
        p.end = __shfl_sync(0xffffffffu, d.z, 8 * half + grp);
        if (p.beg + sub < p.end) p.gi = ld_i32_hint(a.nbr_w + p.beg + sub, pol);
        if (p.beg + 4 + sub < p.end) p.gin = ld_i32_hint(a.nbr_w + p.beg + 4 + sub, pol);
      }
      return p;
    };
    int4 d = load_desc(qi);
    for (int j = qi; j < 2 * my_tiles; j += kTmPerQ) {
      // live state during the gathers is kept to the sums themselves: the 80-register budget holds one gather (8 + 32
      // landing registers) plus the first pass's sum, not prefetched state of the next double pass
      int rowA = __shfl_sync(0xffffffffu, d.x, grp), rowB = __shfl_sync(0xffffffffu, d.x, 8 + grp);
      const int4 dc = d;
      const PassIdx pa = load_idx(d, 0);
      const PassIdx pb = load_idx(d, 1);
      d = load_desc(j + kTmPerQ);           // next double pass: in flight during the gathers
      Row8 accA, accB;
#pragma unroll
      for (int q = 0; q < 8; ++q) accA.v[q] = accB.v[q] = 0.f;
      // partial-sum slots of the two passes' tasks: read with the FULL mask before the groups diverge (the source lanes
      // grp / 8 + grp belong to other groups, which may not take the branches below)
      const int slotA = __shfl_sync(0xffffffffu, dc.w, grp), slotB = __shfl_sync(0xffffffffu, dc.w, 8 + grp);
      if (kMode == 0) {
        // a hub segment stores its partial sum at once and leaves its row of the block empty (row id -1)
        if (rowA >= 0) {
          accA = gather_sum<MGCN_TM_L1 != 0>(a.z, a.nbr_w, pa.beg, pa.end, pa.gi, pa.gin, sub, grp_lane0, gmask, col, pol);
          if (slotA != 0) {
            store_partial(a.partial, slotA, col, accA);
            rowA = -1;
          }
        }
        if (rowB >= 0) {
          accB = gather_sum<MGCN_TM_L1 != 0>(a.z, a.nbr_w, pb.beg, pb.end, pb.gi, pb.gin, sub, grp_lane0, gmask, col, pol);
          if (slotB != 0) {
            store_partial(a.partial, slotB, col, accB);
            rowB = -1;
          }
        }
      } else {
#pragma unroll 1
        for (int i = 0; i < 16; ++i) {
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
          if (i == grp) accA = tot;
          if (i == 8 + grp) accB = tot;
        }
      }
      // the rows' own z: issued after the gathers (16 more live registers during them spill at the 80-register
      // budget), in flight while the block is claimed and the sums are split and stored
      const bool finA = rowA >= 0, finB = rowB >= 0;
      Row8 zA = ld_row8(a.z + (int64_t)(finA ? rowA : 0) * kGH + col);
      Row8 zB = ld_row8(a.z + (int64_t)(finB ? rowB : 0) * kGH + col);
      float scA = 1.f, scB = 1.f;
      {
        const float* sp = sub == 0 ? a.post : sub == 1 ? a.in_scale : sub == 2 ? a.out_scale : nullptr;
        if (finA && sp) scA = __ldg(sp + rowA);
        if (finB && sp) scB = __ldg(sp + rowB);
      }
      // next free 16-lane block of this warp's quarter: tile tl (in completion order), half h
      uint32_t my = 0;
      if (lane == 0) asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(my) : "r"(smem_u32(next_block + qtr)) : "memory");
      my = __shfl_sync(0xffffffffu, my, 0);
      const int64_t tl = my >> 1;
      const int h = (int)(my & 1u), stage = (int)(tl % kTmStages);
      if (tl >= kTmStages) {   // the tensor core has consumed this stage's previous tile
        const int64_t tp = tl - kTmStages;
        mbar_wait(bar_done + (int)(tp % (2 * kTmStages)), (uint32_t)(tp / (2 * kTmStages)) & 1u);
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      {
        const uint32_t tb = tmem + stage * kTmStageCols + ((uint32_t)(32 * qtr + 16 * h) << 16);
        tm_store_pair(tb + 0, tb + 32, accA, accB);
        tm_store_pair(tb + 64, tb + 96, zA, zB);
        const int r = 32 * qtr + 16 * h + grp;
        if (sub == 1) {
          scA = __frcp_rn(scA);
          scB = __frcp_rn(scB);
        }
        if (sub == 3) {
          scA = __int_as_float(finA ? rowA : -1);
          scB = __int_as_float(finB ? rowB : -1);
        }
        float4* sc = scal + (int)(tl % (2 * kTmStages)) * kTmRows;
        reinterpret_cast<float*>(sc + r)[sub] = scA;
        reinterpret_cast<float*>(sc + r + 8)[sub] = scB;
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      uint32_t old = 0;
      if (lane == 0) {
        mbar_arrive(bar_full + (int)(tl % (2 * kTmStages)));
        asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(arrivals + stage)) : "memory");
      }
      old = __shfl_sync(0xffffffffu, old, 0);
      if ((old & 7u) == 7u) {
        // 8th block of the tile: wait for the other blocks (acquire), then for the accumulator buffer, then issue
        const uint32_t use = (uint32_t)(tl / kTmStages);
        mbar_wait(bar_full + (int)(tl % (2 * kTmStages)), (uint32_t)(tl / (2 * kTmStages)) & 1u);
        if (use >= 1) mbar_wait(bar_tfree + stage, (use - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t ta = tmem + stage * kTmStageCols, td = ta + 128;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ko = 2 * k;   // 8 tf32 = 32 bytes inside the 128-byte swizzle row of the weight images
            const uint64_t b1 = dsc + ((kTmOffB1 >> 4) + ko), b2 = dsc + ((kTmOffB2 >> 4) + ko);
            umma_tf32_ts(td + 0, ta + 0 + 8 * k, b1, id64, k > 0);        // s_hi [W_hi | W_lo]
            umma_tf32_ts(td + 32, ta + 32 + 8 * k, b1, id32, 1);          // s_lo W_hi
            umma_tf32_ts(td + 64, ta + 64 + 8 * k, b2, id64, k > 0);      // z_hi [R_hi | R_lo]
            umma_tf32_ts(td + 96, ta + 96 + 8 * k, b2, id32, 1);          // z_lo R_hi
          }
          umma_commit(bar_done + (int)(tl % (2 * kTmStages)));
        }
        __syncwarp();
      }
    }
  } else {
    // ------------------------------- epilogue: thread per row -------------------------------
    const int r = 32 * warp + lane;
    float* stg = reinterpret_cast<float*>(smem + kTmOffOut) + r * kTmLdo;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int stage = tl % kTmStages;
      mbar_wait(bar_done + tl % (2 * kTmStages), (uint32_t)(tl / (2 * kTmStages)) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      mbar_wait(bar_full + tl % (2 * kTmStages), (uint32_t)(tl / (2 * kTmStages)) & 1u);
      const float4 sc4 = scal[(tl % (2 * kTmStages)) * kTmRows + r];
      const float postv = sc4.x, inv = sc4.y, outs = sc4.z;
      const int row = __float_as_int(sc4.w);
      const uint32_t ta = tmem + stage * kTmStageCols + 128 + ((uint32_t)(32 * warp) << 16);
      uint32_t bits = 0;
#pragma unroll
      for (int c0 = 0; c0 < 32; c0 += 8) {
        uint32_t m1[8], c1[8], m2[8], c2[8];
        tm_tmem_ld8(ta + c0, m1);
        tm_tmem_ld8(ta + 32 + c0, c1);
        tm_tmem_ld8(ta + 64 + c0, m2);
        tm_tmem_ld8(ta + 96 + c0, c2);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int c = c0 + t;
          const float v1 = __uint_as_float(m1[t]) + __uint_as_float(c1[t]);
          const float v2 = __uint_as_float(m2[t]) + __uint_as_float(c2[t]);
          float hh = __fmul_rn(postv, v1) + vec[32 + c];
          hh = hh > 0.f ? hh : 0.f;
          bits |= (hh > 0.f ? 1u : 0u) << c;
          float y = hh + (__fmul_rn(inv, v2) + vec[c]);
          if (a.act_out == 1) y = y > 0.f ? y : 0.f;
          o[t] = y * outs;
        }
        *reinterpret_cast<float4*>(stg + c0) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(stg + c0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tfree + stage);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (row >= 0) {
        a.hmask[row] = bits;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], 128, %2;" ::"l"(
                         a.z_next + (int64_t)row * kGH),
                     "r"(smem_u32(stg)), "l"(pol)
                     : "memory");
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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

extern "C" int mgcn_gcn_layer_fwd_tm(const mgcn_csr_t* g, const float* z, int64_t n_in, const float* w,
                                     const float* res_w, const float* res_b, const float* bias,
                                     const float* in_scale, const float* post, const float* out_scale, int act_out,
                                     int64_t H, float* z_next, uint32_t* hmask, void* workspace,
                                     size_t* workspace_bytes, void* stream) {
  MGCN_REQUIRE(workspace_bytes != nullptr && g != nullptr, MGCN_ERR_NULL);
  MGCN_REQUIRE(H == kGH, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(act_out == 0 || act_out == 1, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(n_in >= 0 && g->n_rows >= 0, MGCN_ERR_RANGE);
  const bool hubs = g->hub_rows && g->hub_seg0 && g->hub_count && g->seg_count && g->hub_cap > 0 && g->seg_cap > 0;
  WorkspaceCarver ws(workspace);
  float* partial = ws.take<float>(hubs ? (size_t)g->seg_cap * kGH : 0);
  if (workspace == nullptr) {
    *workspace_bytes = ws.bytes();
    return MGCN_OK;
  }
  MGCN_REQUIRE(*workspace_bytes >= ws.bytes(), MGCN_ERR_WORKSPACE);
  if (g->n_rows == 0) return MGCN_OK;
  MGCN_REQUIRE(z && w && res_w && z_next && hmask && g->rowptr && g->tasks, MGCN_ERR_NULL);
  MGCN_REQUIRE(g->nnz_cap == 0 || g->nbr_w, MGCN_ERR_NULL);
  MGCN_REQUIRE((reinterpret_cast<uintptr_t>(z) & 31u) == 0, MGCN_ERR_ALIGN);
  MGCN_REQUIRE(aligned16(g->tasks) && aligned16(z_next) && aligned16(partial), MGCN_ERR_ALIGN);
  FwdTmArgs a{};
  a.tasks = reinterpret_cast<const int4*>(g->tasks);
  a.nbr_w = g->nbr_w;
  a.seg_count = hubs ? g->seg_count : nullptr;
  a.hub_rows = g->hub_rows;
  a.hub_seg0 = g->hub_seg0;
  a.hub_count = g->hub_count;
  a.rowptr = g->rowptr;
  a.hub_threshold = g->hub_threshold;
  a.z = z; a.w = w; a.res_w = res_w; a.res_b = res_b; a.bias = bias;
  a.in_scale = in_scale; a.post = post; a.out_scale = out_scale;
  a.z_next = z_next; a.hmask = hmask; a.partial = partial;
  a.n_rows = g->n_rows;
  a.seg_cap = hubs ? g->seg_cap : 0;
  a.hub_cap = hubs ? g->hub_cap : 0;
  a.act_out = act_out;
  const int sms = num_sms();
  MGCN_CHECK_CUDA(cudaFuncSetAttribute(k_gcn_fwd_tm<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmSmem));
  MGCN_CHECK_CUDA(cudaFuncSetAttribute(k_gcn_fwd_tm<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmSmem));
  int64_t tiles = ceil_div(a.n_rows + a.seg_cap, kTmRows);
  MGCN_LAUNCH(k_gcn_fwd_tm<0>, (unsigned)(tiles < sms ? tiles : sms), kTmThreads, kTmSmem, stream, a);
  if (hubs) {
    tiles = ceil_div(a.hub_cap, kTmRows);
    MGCN_LAUNCH(k_gcn_fwd_tm<1>, (unsigned)(tiles < sms ? tiles : sms), kTmThreads, kTmSmem, stream, a);
  }
  return MGCN_OK;
}
