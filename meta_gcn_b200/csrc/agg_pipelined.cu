// Software-pipelined aggregation for inputs without per-entry weights (pre-scaled GCN messages,
// GIN sums, SAGE means):   out_i = act( post[i] * sum_k x[nbr_k] (/ len_i) + bias + residual_i ).
//
// The plain row-owned kernel (spmm.cu) is bound by a chain of four dependent loads per row
// (order -> rowptr -> nbr -> x): ncu shows ~40 % issue utilisation, L2 at 25 %, long-scoreboard
// stalls on the index loads (profiles/r1_v2_summary.md).  Here a persistent lane group walks the
// work descriptors {row, beg, end, partial_slot} that mgcn_csr_build wrote in work order and keeps
// three tasks in flight: descriptor of task i+3, first index batch (and post scale) of task i+2
// ... i+1, gathers of task i — so only the feature gather itself is exposed.  Inside a task the
// next index batch is fetched before the current batch's gathers are consumed.
//
// Summation order inside a row / segment is the row order (= edge_index order), one rounded add
// per entry: bit-identical to the reference's CPU scatter_add for rows within hub_threshold.
// Hub rows arrive as segment tasks writing partial rows; k_agg_hub_combine adds them left to right.
#include "common.cuh"

namespace mgcn {

struct PlainArgs {
  const int4* tasks;
  const int32_t* nbr;
  const int32_t* rowptr;
  const float* x;
  const float* post_scale;
  const float* bias;
  const float* residual;
  float* out;
  float* partial;
  const int32_t* seg_count;
  const int32_t* hub_rows;
  const int32_t* hub_seg0;
  const int32_t* hub_count;
  int64_t n_rows;
  int64_t seg_cap;
  int64_t hub_cap;
  int reduce;
  int act;
  int hub_threshold;
};

__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z),
                     __fadd_rn(a.w, b.w));
}

template <int LPR>
__device__ __forceinline__ void finish_row(const PlainArgs& a, int64_t row, int len, int col,
                                           float4 acc, float ps, float4 res, uint64_t pol_stream) {
  constexpr int H = LPR * 4;
  if (a.post_scale) {
    acc.x = __fmul_rn(acc.x, ps); acc.y = __fmul_rn(acc.y, ps);
    acc.z = __fmul_rn(acc.z, ps); acc.w = __fmul_rn(acc.w, ps);
  }
  if (a.reduce == 1) {
    const float c = (float)max(len, 1);
    acc.x = __fdiv_rn(acc.x, c); acc.y = __fdiv_rn(acc.y, c);
    acc.z = __fdiv_rn(acc.z, c); acc.w = __fdiv_rn(acc.w, c);
  }
  if (a.bias) acc = f4_add(acc, __ldg(reinterpret_cast<const float4*>(a.bias + col)));
  if (a.residual) acc = f4_add(acc, res);
  if (a.act == 1) {
    acc.x = acc.x < 0.f ? 0.f : acc.x; acc.y = acc.y < 0.f ? 0.f : acc.y;
    acc.z = acc.z < 0.f ? 0.f : acc.z; acc.w = acc.w < 0.f ? 0.f : acc.w;
  }
  st_f4_hint(a.out + row * H + col, acc, pol_stream);
}

template <int LPR>
__global__ void __launch_bounds__(256, 3) k_agg_plain(const PlainArgs a) {
  constexpr int H = LPR * 4;
  constexpr int GPW = 32 / LPR;
  constexpr int U = 4;  // gathers issued back to back per lane
  static_assert(LPR >= 4, "LPR >= 4");
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int grp = lane / LPR;
  const int grp_lane0 = grp * LPR;
  const unsigned gmask = LPR == 32 ? 0xffffffffu : ((((1u << (LPR & 31)) - 1u)) << grp_lane0);
  const int col = sub * 4;
  const int64_t G = (int64_t)gridDim.x * (blockDim.x >> 5) * GPW;
  const int64_t gid = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + grp;
  int64_t n_tasks = a.n_rows;
  if (a.seg_count) {
    int64_t ns = *a.seg_count;
    if (ns > a.seg_cap) ns = a.seg_cap;
    n_tasks += ns;
  }
  const int4 kNone = make_int4(-1, 0, 0, 0);
  const uint64_t pol_stream = policy_evict_first();

#define MGCN_LOAD_DESC(s) ((s) < n_tasks ? ld_i4_hint(a.tasks + (s), pol_stream) : kNone)
#define MGCN_LOAD_IDX(d) (((d).y + sub < (d).z) ? ld_i32_hint(a.nbr + (d).y + sub, pol_stream) : 0)
#define MGCN_LOAD_PS(d) ((a.post_scale && (d).x >= 0 && (d).w == 0) ? __ldg(a.post_scale + (d).x) : 1.f)

  int4 d0 = MGCN_LOAD_DESC(gid);
  int4 d1 = MGCN_LOAD_DESC(gid + G);
  int4 d2 = MGCN_LOAD_DESC(gid + 2 * G);
  int gi0 = MGCN_LOAD_IDX(d0);
  int gi1 = MGCN_LOAD_IDX(d1);
  float ps0 = MGCN_LOAD_PS(d0);
  float ps1 = MGCN_LOAD_PS(d1);

  for (int64_t s = gid; s < n_tasks; s += G) {
    const int4 d3 = MGCN_LOAD_DESC(s + 3 * G);
    const int gi2 = MGCN_LOAD_IDX(d2);
    const float ps2 = MGCN_LOAD_PS(d2);

    if (d0.x >= 0) {
      const int64_t row = d0.x;
      const int end = d0.z;
      float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.residual && d0.w == 0) res = __ldg(reinterpret_cast<const float4*>(a.residual + row * H + col));
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int e = d0.y;
      int gi = gi0;
      while (true) {
        const int cnt = min(LPR, end - e);
        const int e_next = e + LPR;
        int g_next = 0;
        if (e_next + sub < end) g_next = ld_i32_hint(a.nbr + e_next + sub, pol_stream);  // next batch, before the gathers
#pragma unroll
        for (int t0 = 0; t0 < LPR; t0 += U) {
          if (t0 < cnt) {
            int j[U];
#pragma unroll
            for (int u = 0; u < U; ++u) j[u] = __shfl_sync(gmask, gi, grp_lane0 + t0 + u);
            float4 xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (t0 + u < cnt)
                xv[u] = __ldg(reinterpret_cast<const float4*>(a.x + (int64_t)j[u] * H + col));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
              if (t0 + u < cnt) acc = f4_add(acc, xv[u]);
            }
          }
        }
        if (e_next >= end) break;
        e = e_next;
        gi = g_next;
      }
      if (d0.w != 0) {
        *reinterpret_cast<float4*>(a.partial + (int64_t)(d0.w - 1) * H + col) = acc;
      } else {
        finish_row<LPR>(a, row, end - d0.y, col, acc, ps0, res, pol_stream);
      }
    }
    d0 = d1; d1 = d2; d2 = d3;
    gi0 = gi1; gi1 = gi2;
    ps0 = ps1; ps1 = ps2;
  }
#undef MGCN_LOAD_DESC
#undef MGCN_LOAD_IDX
#undef MGCN_LOAD_PS
}

// one lane group per hub row: its segments' partial rows added left to right, then the epilogue
template <int LPR>
__global__ void __launch_bounds__(256) k_agg_hub_combine(const PlainArgs a) {
  constexpr int H = LPR * 4;
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int grp = lane / LPR;
  const int col = sub * 4;
  const int64_t G = (int64_t)gridDim.x * (blockDim.x >> 5) * GPW;
  int64_t k = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + grp;
  int64_t nh = *a.hub_count;
  if (nh > a.hub_cap) nh = a.hub_cap;
  for (; k < nh; k += G) {
    const int64_t row = __ldg(a.hub_rows + k);
    const int64_t s0 = __ldg(a.hub_seg0 + k);
    const int len = __ldg(a.rowptr + row + 1) - __ldg(a.rowptr + row);
    const int nseg = (len + a.hub_threshold - 1) / a.hub_threshold;
    float4 tot = *reinterpret_cast<const float4*>(a.partial + s0 * H + col);
    for (int q = 1; q < nseg; ++q)
      tot = f4_add(tot, *reinterpret_cast<const float4*>(a.partial + (s0 + q) * H + col));
    float4 res = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a.residual) res = __ldg(reinterpret_cast<const float4*>(a.residual + row * H + col));
    const float ps = a.post_scale ? __ldg(a.post_scale + row) : 1.f;
    finish_row<LPR>(a, row, len, col, tot, ps, res, policy_evict_first());
  }
}

template <int LPR>
static int launch_plain(const PlainArgs& a, void* stream) {
  constexpr int GPW = 32 / LPR;
  const int tasks_per_block = 8 * GPW;
  const int64_t max_tasks = a.n_rows + a.seg_cap;
  int64_t blocks = ceil_div(max_tasks > 0 ? max_tasks : 1, tasks_per_block);
  const int64_t persistent = (int64_t)kNumSMs * 3;  // 3 CTAs/SM resident (launch bounds)
  if (blocks > persistent) blocks = persistent;
  MGCN_LAUNCH((k_agg_plain<LPR>), (unsigned)blocks, 256, 0, stream, a);
  if (a.hub_cap > 0 && a.seg_cap > 0) {
    int64_t hb = ceil_div(a.hub_cap, tasks_per_block);
    if (hb > (int64_t)kNumSMs * 2) hb = (int64_t)kNumSMs * 2;
    MGCN_LAUNCH((k_agg_hub_combine<LPR>), (unsigned)hb, 256, 0, stream, a);
  }
  return MGCN_OK;
}

}  // namespace mgcn

using namespace mgcn;

extern "C" int mgcn_aggregate_prescaled(const mgcn_csr_t* g, const float* x, int64_t n_in,
                                        int64_t H, const float* post_scale, int reduce,
                                        const float* bias, const float* residual, int act,
                                        float* out, void* workspace, size_t* workspace_bytes,
                                        void* stream) {
  MGCN_REQUIRE(g != nullptr && workspace_bytes != nullptr, MGCN_ERR_NULL);
  MGCN_REQUIRE(reduce == 0 || reduce == 1, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(act == 0 || act == 1, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(n_in >= 0 && g->n_rows >= 0, MGCN_ERR_RANGE);
  MGCN_REQUIRE(H == 16 || H == 32 || H == 64 || H == 128, MGCN_ERR_SHAPE);
  const bool hubs = g->hub_rows && g->hub_seg0 && g->hub_count && g->seg_count &&
                    g->hub_cap > 0 && g->seg_cap > 0;
  WorkspaceCarver ws(workspace);
  float* partial = ws.take<float>(hubs ? (size_t)g->seg_cap * H : 0);
  if (workspace == nullptr) {
    *workspace_bytes = ws.bytes();
    return MGCN_OK;
  }
  MGCN_REQUIRE(*workspace_bytes >= ws.bytes(), MGCN_ERR_WORKSPACE);
  if (g->n_rows == 0) return MGCN_OK;
  MGCN_REQUIRE(g->rowptr && g->tasks && out, MGCN_ERR_NULL);
  MGCN_REQUIRE(g->nnz_cap == 0 || (g->nbr_w && x), MGCN_ERR_NULL);
  MGCN_REQUIRE(aligned16(g->tasks) && aligned16(x) && aligned16(out) && aligned16(partial) &&
                   (!bias || aligned16(bias)) && (!residual || aligned16(residual)),
               MGCN_ERR_ALIGN);
  if (H == 32 && reduce == 0 && bias == nullptr && residual == nullptr &&
      (reinterpret_cast<uintptr_t>(x) & 31u) == 0)
    return launch_agg_flat32(g, x, post_scale, act, out, partial, hubs, stream);
  PlainArgs a{};
  a.tasks = reinterpret_cast<const int4*>(g->tasks);
  a.nbr = g->nbr_w;   // descriptors index the work-ordered stream
  a.rowptr = g->rowptr;
  a.x = x;
  a.post_scale = post_scale;
  a.bias = bias;
  a.residual = residual;
  a.out = out;
  a.partial = partial;
  a.seg_count = hubs ? g->seg_count : nullptr;
  a.hub_rows = g->hub_rows;
  a.hub_seg0 = g->hub_seg0;
  a.hub_count = g->hub_count;
  a.n_rows = g->n_rows;
  a.seg_cap = hubs ? g->seg_cap : 0;
  a.hub_cap = hubs ? g->hub_cap : 0;
  a.reduce = reduce;
  a.act = act;
  a.hub_threshold = g->hub_threshold;
  switch (H) {
    case 16: return launch_plain<4>(a, stream);
    case 32: return launch_plain<8>(a, stream);
    case 64: return launch_plain<16>(a, stream);
    default: return launch_plain<32>(a, stream);
  }
}
