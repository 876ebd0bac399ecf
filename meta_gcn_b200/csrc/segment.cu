// Contiguous segment reductions: global_mean_pool / global_add_pool (kernel/gcn.py:29,
// gin.py:44, graph_sage.py:29) and GCNModel's pred_on='graph' mean (gcn_model.py:112-123).
// `batch` is sorted ascending (Batch.from_data_list), so a graph is a contiguous row range and the
// scatter_('mean') of the reference becomes a row-range sum: no atomics, fixed summation order.
#include "common.cuh"

namespace mgcn {

__global__ void __launch_bounds__(256) k_batch_to_offsets(const int64_t* __restrict__ batch,
                                                          int64_t N, int64_t G,
                                                          int32_t* __restrict__ offsets) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g > G) return;
  // first position whose graph id is >= g
  int64_t lo = 0, hi = N;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (batch[mid] < g) lo = mid + 1; else hi = mid;
  }
  offsets[g] = (int32_t)lo;
}

// One CTA per segment.  Thread (rl, c): row lane rl of RL, column c of a CW-wide column window.
// A row lane sums a contiguous chunk of at least kSegMinChunk rows sequentially, so segments of up
// to kSegMinChunk rows (TU-sized graphs) are summed in exactly the reference's row order.
constexpr int kSegMinChunk = 64;

// S slices per segment (S = 1: TU-sized graphs, one CTA per graph, result written directly).  Large
// segments (a botnet graph of 143 k nodes under pred_on='graph', a products-sized graph under
// global_mean_pool) are cut into S contiguous slices summed by S CTAs into partial rows, which
// k_segment_finish adds in slice order: deterministic, and the machine is filled (one CTA per segment
// took 200 ms for a 2.45 M-row graph at H = 256).
__global__ void __launch_bounds__(256)
    k_segment_reduce(const float* __restrict__ x, int H, const int32_t* __restrict__ offsets,
                     int mode, int S, float* __restrict__ partial, float* __restrict__ out) {
  __shared__ float part[256];
  const int g = blockIdx.x / S, sl = blockIdx.x % S;
  const int seg_beg = offsets[g], seg_end = offsets[g + 1];
  const int seg_len = seg_end - seg_beg;
  const int per = (seg_len + S - 1) / S;
  const int beg = min(seg_beg + sl * per, seg_end), end = min(beg + per, seg_end);
  const int len = end - beg;
  const int cw = H < 256 ? H : 256;  // columns handled per pass
  int rl_count = 256 / cw;           // row lanes
  if (rl_count < 1) rl_count = 1;
  const int c_in = threadIdx.x % cw;
  const int rl = threadIdx.x / cw;
  int chunk = (len + rl_count - 1) / rl_count;
  if (chunk < kSegMinChunk) chunk = kSegMinChunk;
  const bool lane_ok = rl < rl_count;
  const int my_beg = min(beg + rl * chunk, end);
  const int my_end = min(my_beg + chunk, end);
  for (int c0 = 0; c0 < H; c0 += cw) {
    const int c = c0 + c_in;
    float s = 0.f;
    if (lane_ok && c < H) {
      for (int r = my_beg; r < my_end; ++r) s = __fadd_rn(s, __ldg(x + (int64_t)r * H + c));
    }
    part[threadIdx.x] = s;
    __syncthreads();
    if (rl == 0 && c < H) {
      float t = part[c_in];
      for (int q = 1; q < rl_count; ++q) t = __fadd_rn(t, part[q * cw + c_in]);
      if (S == 1) {
        if (mode == 1) t = __fdiv_rn(t, (float)max(seg_len, 1));
        out[(int64_t)g * H + c] = t;
      } else {
        partial[(int64_t)blockIdx.x * H + c] = t;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
    k_segment_finish(const float* __restrict__ partial, int H, int S, const int32_t* __restrict__ offsets,
                     int64_t G, int mode, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * H) return;
  const int64_t g = i / H;
  const int c = (int)(i - g * H);
  float t = 0.f;
  for (int sl = 0; sl < S; ++sl) t = __fadd_rn(t, partial[(g * S + sl) * H + c]);
  if (mode == 1) t = __fdiv_rn(t, (float)max(offsets[g + 1] - offsets[g], 1));
  out[i] = t;
}

__global__ void __launch_bounds__(256)
    k_segment_broadcast(const float* __restrict__ gout, int H, const int32_t* __restrict__ offsets,
                        int64_t G, int64_t N, int mode, float* __restrict__ dx) {
  const int64_t total = N * H;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t n = i / H;
    const int c = (int)(i - n * H);
    // segment of row n: last g with offsets[g] <= n
    int64_t lo = 0, hi = G;  // invariant: offsets[lo] <= n < offsets[hi]
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (offsets[mid] <= n) lo = mid; else hi = mid;
    }
    float v = __ldg(gout + lo * H + c);
    if (mode == 1) v = __fdiv_rn(v, (float)max(offsets[lo + 1] - offsets[lo], 1));
    dx[i] = v;
  }
}

}  // namespace mgcn

using namespace mgcn;

extern "C" int mgcn_batch_to_offsets(const int64_t* batch, int64_t N, int64_t G, int32_t* offsets,
                                     void* stream) {
  MGCN_REQUIRE(N >= 0 && G >= 0 && N < (int64_t(1) << 31) && G < (int64_t(1) << 31),
               MGCN_ERR_RANGE);
  MGCN_REQUIRE(offsets != nullptr, MGCN_ERR_NULL);
  MGCN_REQUIRE(N == 0 || batch != nullptr, MGCN_ERR_NULL);
  MGCN_LAUNCH(k_batch_to_offsets, (unsigned)ceil_div(G + 1, 256), 256, 0, stream, batch, N, G,
              offsets);
  return MGCN_OK;
}

extern "C" int mgcn_segment_reduce(const float* x, int64_t H, const int32_t* offsets, int64_t G,
                                   int64_t N, int mode, float* out, void* workspace,
                                   size_t* workspace_bytes, void* stream) {
  MGCN_REQUIRE(workspace_bytes != nullptr, MGCN_ERR_NULL);
  MGCN_REQUIRE(H >= 1 && H <= 65536, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(mode == 0 || mode == 1, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(G >= 0 && G < (int64_t(1) << 31) && N >= 0, MGCN_ERR_RANGE);
  // slices per segment: only when the average segment is long, enough CTAs to fill the machine
  int64_t S = 1;
  if (G > 0 && N / G > 4096) {
    S = ceil_div((int64_t)kNumSMs * 4, G);
    const int64_t max_s = ceil_div(N / G, 1024);
    if (S > max_s) S = max_s;
    if (S < 1) S = 1;
  }
  WorkspaceCarver ws(workspace);
  float* partial = ws.take<float>(S > 1 ? (size_t)(G * S * H) : 0);
  if (workspace == nullptr) {
    *workspace_bytes = ws.bytes();
    return MGCN_OK;
  }
  MGCN_REQUIRE(*workspace_bytes >= ws.bytes(), MGCN_ERR_WORKSPACE);
  if (G == 0) return MGCN_OK;
  MGCN_REQUIRE(offsets && out, MGCN_ERR_NULL);
  MGCN_LAUNCH(k_segment_reduce, (unsigned)(G * S), 256, 0, stream, x, (int)H, offsets, mode, (int)S, partial, out);
  if (S > 1) {
    MGCN_LAUNCH(k_segment_finish, (unsigned)ceil_div(G * H, 256), 256, 0, stream, partial, (int)H, (int)S, offsets,
                G, mode, out);
  }
  return MGCN_OK;
}

extern "C" int mgcn_segment_broadcast(const float* gout, int64_t H, const int32_t* offsets,
                                      int64_t G, int64_t N, int mode, float* dx, void* stream) {
  MGCN_REQUIRE(H >= 1 && H <= 65536, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(mode == 0 || mode == 1, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(G >= 0 && N >= 0, MGCN_ERR_RANGE);
  if (N == 0) return MGCN_OK;
  MGCN_REQUIRE(G >= 1, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(gout && offsets && dx, MGCN_ERR_NULL);
  int64_t blocks = ceil_div(N * H, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  MGCN_LAUNCH(k_segment_broadcast, (unsigned)blocks, 256, 0, stream, gout, (int)H, offsets, G, N,
              mode, dx);
  return MGCN_OK;
}
