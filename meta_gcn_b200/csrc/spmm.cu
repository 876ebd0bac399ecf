// Row-owned aggregation: out_i = act( reduce_k  x[nbr_k] * w_k  + bias + residual_i ).
//
// Replaces index_select -> mul -> scatter_add (gcn_base_models.py:223-237; PyG propagate) without
// materialising [E,H] and without atomics.  A lane group of LPR lanes owns one row and walks its
// entries in row order (= edge_index order), each lane holding 4 (or 1) feature columns, so the
// fp32 result of a non-hub row is the same sequence of rounded mul/add the reference's CPU
// scatter_add performs.  Rows longer than hub_threshold are summed by a whole CTA: contiguous
// chunks per lane group, then a fixed left-to-right combine (deterministic, not order-identical).
//
// HBM-bound: per pass 4 B/entry of indices + one read and one write of [N,H]; the gathers are
// served by L1/L2 (graphs in a batch are block-diagonal, so the live set is a few graphs wide).
#include "common.cuh"

namespace mgcn {

struct SpmmArgs {
  const int32_t* rowptr;
  const int32_t* gather_idx;  // nbr, or perm when gathering per-edge messages
  const int32_t* nbr;
  const float* x;
  const float* edge_val;
  const float* nbr_scale;
  const float* row_scale;
  const float* bias;
  const float* residual;
  float* out;
  const int32_t* hub_rows;
  const int32_t* hub_count;
  int64_t n_rows;
  int64_t hub_cap;
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

template <bool VEC4>
__device__ __forceinline__ void mul_add(Vals<VEC4>& acc, const Vals<VEC4>& x, float w) {
#pragma unroll
  for (int i = 0; i < Vals<VEC4>::V; ++i) acc.v[i] = __fadd_rn(acc.v[i], __fmul_rn(x.v[i], w));
}

// Sequentially accumulate entries [beg,end) of one row for the columns [col, col+V).
// All LPR lanes of the group call this together (shuffles use gmask).
template <int LPR, bool VEC4>
__device__ __forceinline__ void accumulate_range(const SpmmArgs& a, int beg, int end, int sub,
                                                 int grp_lane0, unsigned gmask, int col,
                                                 bool col_ok, float rs, Vals<VEC4>& acc) {
  for (int e = beg; e < end; e += LPR) {
    // cooperative, coalesced fetch of up to LPR entries: lane `sub` takes entry e+sub
    const int k = e + sub;
    int gi = 0;
    float w = 1.f;
    if (k < end) {
      gi = __ldg(a.gather_idx + k);
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
    const int cnt = min(LPR, end - e);
    int t = 0;
    for (; t + 4 <= cnt; t += 4) {
      int j[4];
      float wj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        j[u] = __shfl_sync(gmask, gi, grp_lane0 + t + u);
        wj[u] = __shfl_sync(gmask, w, grp_lane0 + t + u);
      }
      Vals<VEC4> xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (col_ok) xv[u] = load_row<VEC4>(a.x + (int64_t)j[u] * a.H + col);
      }
      if (col_ok) {
#pragma unroll
        for (int u = 0; u < 4; ++u) mul_add<VEC4>(acc, xv[u], wj[u]);
      }
    }
    for (; t < cnt; ++t) {
      const int j = __shfl_sync(gmask, gi, grp_lane0 + t);
      const float wj = __shfl_sync(gmask, w, grp_lane0 + t);
      if (col_ok) {
        const Vals<VEC4> xv = load_row<VEC4>(a.x + (int64_t)j * a.H + col);
        mul_add<VEC4>(acc, xv, wj);
      }
    }
  }
}

template <bool VEC4>
__device__ __forceinline__ void epilogue_store(const SpmmArgs& a, int64_t row, int len, int col,
                                               Vals<VEC4> acc) {
  constexpr int V = Vals<VEC4>::V;
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

template <int LPR, bool VEC4>
__global__ void __launch_bounds__(256) k_spmm_rows(const SpmmArgs a) {
  constexpr int V = Vals<VEC4>::V;
  constexpr int GPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int grp = lane / LPR;
  const int grp_lane0 = grp * LPR;
  const unsigned gmask = group_mask<LPR>(grp_lane0);
  const int64_t row =
      ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * GPW + grp;
  if (row >= a.n_rows) return;
  const int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
  const int len = end - beg;
  if (len > a.hub_threshold) return;  // summed by k_spmm_hubs
  const float rs = a.row_scale ? __ldg(a.row_scale + row) : 1.f;
  for (int c0 = 0; c0 < a.H; c0 += LPR * V) {
    const int col = c0 + sub * V;
    const bool col_ok = col < a.H;
    Vals<VEC4> acc;
#pragma unroll
    for (int i = 0; i < V; ++i) acc.v[i] = 0.f;
    accumulate_range<LPR, VEC4>(a, beg, end, sub, grp_lane0, gmask, col, col_ok, rs, acc);
    if (col_ok) epilogue_store<VEC4>(a, row, len, col, acc);
  }
}

template <int LPR, bool VEC4>
__global__ void __launch_bounds__(256) k_spmm_hubs(const SpmmArgs a) {
  constexpr int V = Vals<VEC4>::V;
  constexpr int NG = 256 / LPR;  // lane groups per CTA
  __shared__ float part[NG][LPR * V];
  const int lane = threadIdx.x & 31;
  const int sub = lane % LPR;
  const int grp_lane0 = (lane / LPR) * LPR;
  const unsigned gmask = group_mask<LPR>(grp_lane0);
  const int g = threadIdx.x / LPR;
  int nh = *a.hub_count;
  if (nh > a.hub_cap) nh = (int)a.hub_cap;
  for (int h = blockIdx.x; h < nh; h += gridDim.x) {
    const int64_t row = a.hub_rows[h];
    const int beg = __ldg(a.rowptr + row), end = __ldg(a.rowptr + row + 1);
    const int len = end - beg;
    const int chunk = (len + NG - 1) / NG;
    const int my_beg = min(beg + g * chunk, end);
    const int my_end = min(my_beg + chunk, end);
    const float rs = a.row_scale ? __ldg(a.row_scale + row) : 1.f;
    for (int c0 = 0; c0 < a.H; c0 += LPR * V) {
      const int col = c0 + sub * V;
      const bool col_ok = col < a.H;
      Vals<VEC4> acc;
#pragma unroll
      for (int i = 0; i < V; ++i) acc.v[i] = 0.f;
      accumulate_range<LPR, VEC4>(a, my_beg, my_end, sub, grp_lane0, gmask, col, col_ok, rs, acc);
#pragma unroll
      for (int i = 0; i < V; ++i) part[g][sub * V + i] = acc.v[i];
      __syncthreads();
      // fixed left-to-right combine of the NG partial sums, one lane group does the epilogue
      if (g == 0) {
        Vals<VEC4> tot;
#pragma unroll
        for (int i = 0; i < V; ++i) tot.v[i] = part[0][sub * V + i];
        for (int q = 1; q < NG; ++q) {
#pragma unroll
          for (int i = 0; i < V; ++i) tot.v[i] = __fadd_rn(tot.v[i], part[q][sub * V + i]);
        }
        if (col_ok) epilogue_store<VEC4>(a, row, len, col, tot);
      }
      __syncthreads();
    }
  }
}

template <int LPR, bool VEC4>
static int launch_spmm(const SpmmArgs& a, void* stream) {
  constexpr int GPW = 32 / LPR;
  const int rows_per_block = 8 * GPW;
  const int64_t blocks = ceil_div(a.n_rows, rows_per_block);
  MGCN_LAUNCH((k_spmm_rows<LPR, VEC4>), (unsigned)blocks, 256, 0, stream, a);
  if (a.hub_cap > 0 && a.hub_rows && a.hub_count) {
    int64_t hb = a.hub_cap < (int64_t)kNumSMs * 4 ? a.hub_cap : (int64_t)kNumSMs * 4;
    MGCN_LAUNCH((k_spmm_hubs<LPR, VEC4>), (unsigned)hb, 256, 0, stream, a);
  }
  return MGCN_OK;
}

}  // namespace mgcn

using namespace mgcn;

extern "C" int mgcn_spmm(const mgcn_csr_t* g, const float* x, int64_t n_in, int64_t H,
                         int gather_perm, const float* edge_val, const float* nbr_scale,
                         const float* row_scale, int reduce, const float* bias,
                         const float* residual, int act, float* out, void* stream) {
  MGCN_REQUIRE(g != nullptr, MGCN_ERR_NULL);
  MGCN_REQUIRE(H >= 1 && H <= 65536, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(reduce == 0 || reduce == 1, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(act == 0 || act == 1, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(n_in >= 0 && g->n_rows >= 0, MGCN_ERR_RANGE);
  if (g->n_rows == 0) return MGCN_OK;
  MGCN_REQUIRE(g->rowptr && out, MGCN_ERR_NULL);
  MGCN_REQUIRE(g->nnz_cap == 0 || (g->nbr && x), MGCN_ERR_NULL);
  MGCN_REQUIRE(!gather_perm || g->perm, MGCN_ERR_NULL);
  const bool vec4 = (H % 4 == 0) && aligned16(x) && aligned16(out) &&
                    (!bias || aligned16(bias)) && (!residual || aligned16(residual));

  SpmmArgs a;
  a.rowptr = g->rowptr;
  a.gather_idx = gather_perm ? g->perm : g->nbr;
  a.nbr = g->nbr;
  a.x = x;
  a.edge_val = edge_val;
  a.nbr_scale = nbr_scale;
  a.row_scale = row_scale;
  a.bias = bias;
  a.residual = residual;
  a.out = out;
  a.hub_rows = g->hub_rows;
  a.hub_count = g->hub_count;
  a.n_rows = g->n_rows;
  a.hub_cap = g->hub_rows && g->hub_count ? g->hub_cap : 0;
  a.H = (int)H;
  a.reduce = reduce;
  a.act = act;
  a.hub_threshold = a.hub_cap > 0 ? g->hub_threshold : 0x7fffffff;

  const int64_t lanes = vec4 ? H / 4 : H;
  if (vec4) {
    if (lanes <= 1) return launch_spmm<1, true>(a, stream);
    if (lanes <= 2) return launch_spmm<2, true>(a, stream);
    if (lanes <= 4) return launch_spmm<4, true>(a, stream);
    if (lanes <= 8) return launch_spmm<8, true>(a, stream);
    if (lanes <= 16) return launch_spmm<16, true>(a, stream);
    return launch_spmm<32, true>(a, stream);
  }
  if (lanes <= 1) return launch_spmm<1, false>(a, stream);
  if (lanes <= 2) return launch_spmm<2, false>(a, stream);
  if (lanes <= 4) return launch_spmm<4, false>(a, stream);
  if (lanes <= 8) return launch_spmm<8, false>(a, stream);
  if (lanes <= 16) return launch_spmm<16, false>(a, stream);
  return launch_spmm<32, false>(a, stream);
}
