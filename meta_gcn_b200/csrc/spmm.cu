// Row-owned aggregation: out_i = act( reduce_k  x[nbr_k] * w_k  + bias + residual_i ).
//
// Replaces index_select -> mul -> scatter_add (gcn_base_models.py:223-237; PyG propagate) without
// materialising [E,H] and without atomics.  A lane group of LPR lanes owns one row and walks its
// entries in row order (= edge_index order), each lane holding 4 (or 1) feature columns, so the
// fp32 sum of a row is the same sequence of rounded adds the reference's CPU scatter_add performs.
// Rows longer than hub_threshold are cut into segments of hub_threshold entries (built by
// mgcn_csr_build): every segment is summed by a lane group like an ordinary row into a partial
// row, and the partials of a hub are combined left to right (deterministic; a different
// association than one long chain).
//
// Work order: rows are visited in g->order (sorted by length inside 16384-row windows), so the
// rows sharing a warp have equal trip counts and stay converged while the gathers stay inside a
// window-sized neighbourhood of the block-diagonal batch.
//
// Two weight modes:
//   kExact  w_k = nbr_scale[nbr_k] * edge_val[k] * row_scale[i] formed per entry in the reference's
//           rounding order, message = x * w_k (mul and add rounded separately).
//   kPlain  no per-entry weight: inputs already carry the per-source factor (or the sum is
//           unweighted); post_scale[i] multiplies the finished sum.
//
// Bound: HBM for the index stream and one read + one write of [N,H]; the gathers are served by
// L1/L2 (B_agg = 4E + 4(N+1) + 4N + 8NH algorithmic bytes per pass).
#include "common.cuh"

namespace mgcn {

enum WeightMode { kExact = 0, kPlain = 1 };

struct SpmmArgs {
  const int32_t* rowptr;
  const int32_t* gather_idx;  // nbr, or perm when gathering per-edge messages
  const int32_t* nbr;
  const int32_t* order;
  const float* x;
  const float* edge_val;
  const float* nbr_scale;
  const float* row_scale;   // kExact: per-entry factor; kPlain: post scale
  const float* bias;
  const float* residual;
  float* out;
  float* partial;           // [seg_cap, H] hub segment sums
  const int32_t* hub_rows;
  const int32_t* hub_seg0;
  const int32_t* hub_count;
  const int32_t* seg_row;
  const int32_t* seg_beg;
  const int32_t* seg_count;
  int64_t n_rows;
  int64_t hub_cap;
  int64_t seg_cap;
  int H;
  int reduce;
  int act;
  int hub_threshold;
};

template <int LPR>
__device__ __forceinline__ unsigned group_mask(int grp_lane0) {
  if constexpr (LPR == 32) return 0xffffffffu;
  else return ((1u << LPR) - 1u) << grp_lane0;
}

template <bool VEC4>
struct Vals {
  static constexpr int V = VEC4 ? 4 : 1;
  float v[V];
};

template <bool VEC4>
__device__ __forceinline__ Vals<VEC4> load_row(const float* __restrict__ p) {
  Vals<VEC4> r;
  if constexpr (VEC4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    r.v[0] = __ldg(p);
  }
  return r;
}

template <bool VEC4>
__device__ __forceinline__ void store_row(float* __restrict__ p, const Vals<VEC4>& r) {
  if constexpr (VEC4) {
    *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
  } else {
    *p = r.v[0];
  }
}

// Sequentially accumulate entries [beg,end) of one row for the columns [col, col+V).
// All LPR lanes of the group call this together (shuffles use gmask).
template <int LPR, bool VEC4, int MODE>
__device__ __forceinline__ void accumulate_range(const SpmmArgs& a, int beg, int end, int sub,
                                                 int grp_lane0, unsigned gmask, int col,
                                                 bool col_ok, float rs, Vals<VEC4>& acc) {
  constexpr int V = Vals<VEC4>::V;
  constexpr int U = LPR < 8 ? LPR : 8;  // gathers in flight per lane
  if constexpr (LPR == 1 && MODE != kExact) {
    // one lane per row (H <= 4): four index loads, then four gathers in flight per lane instead of a chain of
    // dependent load pairs (the H_in = 1 aggregation of the botnet stack's first layer); same sums, same order
    int e = beg;
    for (; e + 4 <= end; e += 4) {
      int j[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) j[u] = __ldg(a.gather_idx + e + u);
      Vals<VEC4> xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (col_ok) xv[u] = load_row<VEC4>(a.x + (int64_t)j[u] * a.H + col);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (col_ok) {
#pragma unroll
          for (int i = 0; i < V; ++i) acc.v[i] = __fadd_rn(acc.v[i], xv[u].v[i]);
        }
    }
    for (; e < end; ++e) {
      const int j = __ldg(a.gather_idx + e);
      if (col_ok) {
        const Vals<VEC4> xv = load_row<VEC4>(a.x + (int64_t)j * a.H + col);
#pragma unroll
        for (int i = 0; i < V; ++i) acc.v[i] = __fadd_rn(acc.v[i], xv.v[i]);
      }
    }
  } else {
  for (int e = beg; e < end; e += LPR) {
    // cooperative, coalesced fetch of up to LPR entries: lane `sub` takes entry e+sub
    const int k = e + sub;
    int gi = 0;
    float w = 1.f;
    if (k < end) {
      gi = __ldg(a.gather_idx + k);
      if constexpr (MODE == kExact) {
        bool has = false;
        if (a.nbr_scale) {
          w = __ldg(a.nbr_scale + __ldg(a.nbr + k));
          has = true;
        }
        if (a.edge_val) {
          const float ev = __ldg(a.edge_val + k);
          w = has ? __fmul_rn(w, ev) : ev;
          has = true;
        }
        if (a.row_scale) w = has ? __fmul_rn(w, rs) : rs;
      }
    }
    const int cnt = min(LPR, end - e);
#pragma unroll
    for (int t0 = 0; t0 < LPR; t0 += U) {
      if (t0 < cnt) {
        int j[U];
        float wj[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          j[u] = __shfl_sync(gmask, gi, grp_lane0 + t0 + u);
          if constexpr (MODE == kExact) wj[u] = __shfl_sync(gmask, w, grp_lane0 + t0 + u);
        }
        Vals<VEC4> xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (col_ok && t0 + u < cnt) xv[u] = load_row<VEC4>(a.x + (int64_t)j[u] * a.H + col);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (col_ok && t0 + u < cnt) {
#pragma unroll
            for (int i = 0; i < V; ++i) {
              if constexpr (MODE == kExact)
                acc.v[i] = __fadd_rn(acc.v[i], __fmul_rn(xv[u].v[i], wj[u]));
              else
                acc.v[i] = __fadd_rn(acc.v[i], xv[u].v[i]);
            }
          }
        }
      }
    }
  }
  }
}

template <bool VEC4, int MODE>
__device__ __forceinline__ void epilogue_store(const SpmmArgs& a, int64_t row, int len, int col,
                                               Vals<VEC4> acc) {
  constexpr int V = Vals<VEC4>::V;
  if constexpr (MODE == kPlain) {
    if (a.row_scale) {
      const float ps = __ldg(a.row_scale + row);
#pragma unroll
      for (int i = 0; i < V; ++i) acc.v[i] = __fmul_rn(acc.v[i], ps);
    }
  }
  if (a.reduce == 1) {
    const float c = (float)max(len, 1);
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] = __fdiv_rn(acc.v[i], c);
  }
  if (a.bias) {
    const Vals<VEC4> b = load_row<VEC4>(a.bias + col);
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] = __fadd_rn(acc.v[i], b.v[i]);
  }
  if (a.residual) {
    const Vals<VEC4> r = load_row<VEC4>(a.residual + row * a.H + col);
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] = __fadd_rn(acc.v[i], r.v[i]);
  }
  if (a.act == 1) {
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] = acc.v[i] < 0.f ? 0.f : acc.v[i];
  }
  store_row<VEC4>(a.out + row * a.H + col, acc);
}

// SEGMENTS == false: one lane group per row (rows longer than hub_threshold are skipped).
// SEGMENTS == true : one lane group per hub segment, result written to a.partial[segment].
template <int LPR, bool VEC4, int MODE, bool SEGMENTS>
__global__ void __launch_bounds__(256) k_spmm(const SpmmArgs a) {
  constexpr int V = Vals<VEC4>::V;
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int grp = lane / LPR;
  const int grp_lane0 = grp * LPR;
  const unsigned gmask = group_mask<LPR>(grp_lane0);
  const int64_t groups_per_grid = (int64_t)gridDim.x * (blockDim.x >> 5) * GPW;
  int64_t slot = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + grp;
  int64_t n_tasks = a.n_rows;
  if constexpr (SEGMENTS) {
    n_tasks = *a.seg_count;
    if (n_tasks > a.seg_cap) n_tasks = a.seg_cap;
  }
  for (; slot < n_tasks; slot += groups_per_grid) {
    int64_t row;
    int beg, end, len;
    if constexpr (SEGMENTS) {
      row = __ldg(a.seg_row + slot);
      beg = __ldg(a.seg_beg + slot);
      const int row_end = __ldg(a.rowptr + row + 1);
      end = min(beg + a.hub_threshold, row_end);
      len = end - beg;
    } else {
      row = a.order ? (int64_t)__ldg(a.order + slot) : slot;
      beg = __ldg(a.rowptr + row);
      end = __ldg(a.rowptr + row + 1);
      len = end - beg;
      if (len > a.hub_threshold) continue;  // summed via segments
    }
    float rs = 1.f;
    if constexpr (MODE == kExact) rs = a.row_scale ? __ldg(a.row_scale + row) : 1.f;
    for (int c0 = 0; c0 < a.H; c0 += LPR * V) {
      const int col = c0 + sub * V;
      const bool col_ok = col < a.H;
      Vals<VEC4> acc;
#pragma unroll
      for (int i = 0; i < V; ++i) acc.v[i] = 0.f;
      accumulate_range<LPR, VEC4, MODE>(a, beg, end, sub, grp_lane0, gmask, col, col_ok, rs, acc);
      if (col_ok) {
        if constexpr (SEGMENTS) store_row<VEC4>(a.partial + slot * a.H + col, acc);
        else epilogue_store<VEC4, MODE>(a, row, len, col, acc);
      }
    }
  }
}

// one lane group per hub row: partial sums of its segments added left to right, then the epilogue
template <int LPR, bool VEC4, int MODE>
__global__ void __launch_bounds__(256) k_spmm_hub_combine(const SpmmArgs a) {
  constexpr int V = Vals<VEC4>::V;
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int grp = lane / LPR;
  const int64_t groups_per_grid = (int64_t)gridDim.x * (blockDim.x >> 5) * GPW;
  int64_t k = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + grp;
  int64_t nh = *a.hub_count;
  if (nh > a.hub_cap) nh = a.hub_cap;
  for (; k < nh; k += groups_per_grid) {
    const int64_t row = __ldg(a.hub_rows + k);
    const int64_t s0 = __ldg(a.hub_seg0 + k);
    const int len = __ldg(a.rowptr + row + 1) - __ldg(a.rowptr + row);
    const int nseg = (len + a.hub_threshold - 1) / a.hub_threshold;
    for (int c0 = 0; c0 < a.H; c0 += LPR * V) {
      const int col = c0 + sub * V;
      if (col >= a.H) continue;
      Vals<VEC4> tot;
      if constexpr (VEC4) {
        const float4 t = *reinterpret_cast<const float4*>(a.partial + s0 * a.H + col);
        tot.v[0] = t.x; tot.v[1] = t.y; tot.v[2] = t.z; tot.v[3] = t.w;
      } else {
        tot.v[0] = a.partial[s0 * a.H + col];
      }
      for (int q = 1; q < nseg; ++q) {
        const float* p = a.partial + (s0 + q) * a.H + col;
#pragma unroll
        for (int i = 0; i < V; ++i) tot.v[i] = __fadd_rn(tot.v[i], p[i]);
      }
      epilogue_store<VEC4, MODE>(a, row, len, col, tot);
    }
  }
}

template <int LPR, bool VEC4, int MODE>
static int launch_spmm(const SpmmArgs& a, void* stream) {
  constexpr int GPW = 32 / LPR;
  const int rows_per_block = 8 * GPW;
  const int64_t blocks = ceil_div(a.n_rows, rows_per_block);
  MGCN_LAUNCH((k_spmm<LPR, VEC4, MODE, false>), (unsigned)blocks, 256, 0, stream, a);
  if (a.hub_cap > 0 && a.seg_cap > 0) {
    int64_t sb = ceil_div(a.seg_cap, rows_per_block);
    if (sb > (int64_t)kNumSMs * 8) sb = (int64_t)kNumSMs * 8;
    MGCN_LAUNCH((k_spmm<LPR, VEC4, MODE, true>), (unsigned)sb, 256, 0, stream, a);
    int64_t hb = ceil_div(a.hub_cap, rows_per_block);
    if (hb > (int64_t)kNumSMs * 2) hb = (int64_t)kNumSMs * 2;
    MGCN_LAUNCH((k_spmm_hub_combine<LPR, VEC4, MODE>), (unsigned)hb, 256, 0, stream, a);
  }
  return MGCN_OK;
}

template <int MODE>
static int dispatch_spmm(const SpmmArgs& a, bool vec4, void* stream) {
  const int64_t lanes = vec4 ? a.H / 4 : a.H;
  if (vec4) {
    if (lanes <= 1) return launch_spmm<1, true, MODE>(a, stream);
    if (lanes <= 2) return launch_spmm<2, true, MODE>(a, stream);
    if (lanes <= 4) return launch_spmm<4, true, MODE>(a, stream);
    if (lanes <= 8) return launch_spmm<8, true, MODE>(a, stream);
    if (lanes <= 16) return launch_spmm<16, true, MODE>(a, stream);
    return launch_spmm<32, true, MODE>(a, stream);
  }
  if (lanes <= 1) return launch_spmm<1, false, MODE>(a, stream);
  if (lanes <= 2) return launch_spmm<2, false, MODE>(a, stream);
  if (lanes <= 4) return launch_spmm<4, false, MODE>(a, stream);
  if (lanes <= 8) return launch_spmm<8, false, MODE>(a, stream);
  if (lanes <= 16) return launch_spmm<16, false, MODE>(a, stream);
  return launch_spmm<32, false, MODE>(a, stream);
}

static int fill_args(const mgcn_csr_t* g, const float* x, int64_t H, int gather_perm, float* out,
                     void* workspace, size_t* workspace_bytes, SpmmArgs* a, bool* done) {
  *done = false;
  MGCN_REQUIRE(g != nullptr && workspace_bytes != nullptr, MGCN_ERR_NULL);
  MGCN_REQUIRE(H >= 1 && H <= 65536, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(g->n_rows >= 0, MGCN_ERR_RANGE);
  const bool hubs = g->hub_rows && g->hub_seg0 && g->hub_count && g->seg_row && g->seg_beg &&
                    g->seg_count && g->hub_cap > 0 && g->seg_cap > 0;
  WorkspaceCarver ws(workspace);
  float* partial = ws.take<float>(hubs ? (size_t)g->seg_cap * H : 0);
  if (workspace == nullptr) {
    *workspace_bytes = ws.bytes();
    *done = true;
    return MGCN_OK;
  }
  MGCN_REQUIRE(*workspace_bytes >= ws.bytes(), MGCN_ERR_WORKSPACE);
  if (g->n_rows == 0) {
    *done = true;
    return MGCN_OK;
  }
  MGCN_REQUIRE(g->rowptr && out, MGCN_ERR_NULL);
  MGCN_REQUIRE(g->nnz_cap == 0 || (g->nbr && x), MGCN_ERR_NULL);
  MGCN_REQUIRE(!gather_perm || g->perm, MGCN_ERR_NULL);
  a->rowptr = g->rowptr;
  a->gather_idx = gather_perm ? g->perm : g->nbr;
  a->nbr = g->nbr;
  a->order = g->order;
  a->x = x;
  a->out = out;
  a->partial = partial;
  a->hub_rows = g->hub_rows;
  a->hub_seg0 = g->hub_seg0;
  a->hub_count = g->hub_count;
  a->seg_row = g->seg_row;
  a->seg_beg = g->seg_beg;
  a->seg_count = g->seg_count;
  a->n_rows = g->n_rows;
  a->hub_cap = hubs ? g->hub_cap : 0;
  a->seg_cap = hubs ? g->seg_cap : 0;
  a->H = (int)H;
  a->hub_threshold = hubs ? g->hub_threshold : 0x7fffffff;
  return MGCN_OK;
}

}  // namespace mgcn

using namespace mgcn;

extern "C" int mgcn_spmm(const mgcn_csr_t* g, const float* x, int64_t n_in, int64_t H,
                         int gather_perm, const float* edge_val, const float* nbr_scale,
                         const float* row_scale, int reduce, const float* bias,
                         const float* residual, int act, float* out, void* workspace,
                         size_t* workspace_bytes, void* stream) {
  MGCN_REQUIRE(reduce == 0 || reduce == 1, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(act == 0 || act == 1, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(n_in >= 0, MGCN_ERR_RANGE);
  SpmmArgs a{};
  bool done = false;
  int rc = fill_args(g, x, H, gather_perm, out, workspace, workspace_bytes, &a, &done);
  if (rc != MGCN_OK || done) return rc;
  a.edge_val = edge_val;
  a.nbr_scale = nbr_scale;
  a.row_scale = row_scale;
  a.bias = bias;
  a.residual = residual;
  a.reduce = reduce;
  a.act = act;
  const bool vec4 = (H % 4 == 0) && aligned16(x) && aligned16(out) && aligned16(a.partial) &&
                    (!bias || aligned16(bias)) && (!residual || aligned16(residual));
  if (!edge_val && !nbr_scale && !row_scale) return dispatch_spmm<kPlain>(a, vec4, stream);
  return dispatch_spmm<kExact>(a, vec4, stream);
}
