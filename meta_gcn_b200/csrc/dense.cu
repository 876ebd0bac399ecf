// Narrow dense transform on the FMA pipes: y = act(x W + b + add), dW = x^T g, relu backward.
//
// torch.matmul(x, weight_node) (gcn_base_models.py:201), the residual / final nn.Linear
// (gcn_model.py:64,73,96,108) and their autograd.  Register-tiled SGEMM: 128x32 output tile per
// 128-thread CTA, 8x4 outputs per thread, operands staged in shared memory; per 128 FMAs a thread
// issues 12 LDS.128.  K is walked in ascending order with fp32 FMA accumulation.
#include "common.cuh"

namespace mgcn {

constexpr int kBM = 128;  // rows per CTA
constexpr int kBN = 32;   // output columns per CTA
constexpr int kBK = 32;   // K slab
constexpr int kXsLd = kBK + 4;

__global__ void __launch_bounds__(128)
    k_linear(const float* __restrict__ x, const float* __restrict__ xmask, int64_t N, int Hi,
             const float* __restrict__ w, int64_t w_sk, int64_t w_sc, int Ho,
             const float* __restrict__ bias, const float* __restrict__ add, int act,
             const float* __restrict__ row_scale, float* __restrict__ y, int x_vec4, int y_vec4) {
  __shared__ __align__(16) float Xs[kBM][kXsLd];
  __shared__ __align__(16) float Ws[kBK][kBN];
  const int tid = threadIdx.x;
  const int cg = tid & 7;   // column group: columns 4*cg .. 4*cg+3 of the tile
  const int rg = tid >> 3;  // row group: rows rg + 16*i
  const int64_t row0 = (int64_t)blockIdx.x * kBM;
  const int col0 = blockIdx.y * kBN;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < Hi; k0 += kBK) {
    // ---- stage X[row0:row0+128, k0:k0+32] ----
    if (x_vec4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rg + 16 * i;
        const int64_t gr = row0 + r;
        const int kk = k0 + 4 * cg;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < N && kk < Hi) {
          v = __ldg(reinterpret_cast<const float4*>(x + gr * Hi + kk));
          if (xmask) {  // relu backward folded into the operand load: x * (mask > 0)
            const float4 m = __ldg(reinterpret_cast<const float4*>(xmask + gr * Hi + kk));
            v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f;
            v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
          }
        }
        *reinterpret_cast<float4*>(&Xs[r][4 * cg]) = v;
      }
    } else {
      for (int idx = tid; idx < kBM * kBK; idx += 128) {
        const int r = idx / kBK, c = idx % kBK;
        const int64_t gr = row0 + r;
        const int kk = k0 + c;
        float v = 0.f;
        if (gr < N && kk < Hi) {
          v = __ldg(x + gr * Hi + kk);
          if (xmask && !(__ldg(xmask + gr * Hi + kk) > 0.f)) v = 0.f;
        }
        Xs[r][c] = v;
      }
    }
    // ---- stage W[k0:k0+32, col0:col0+32] ----
    for (int idx = tid; idx < kBK * kBN; idx += 128) {
      const int kr = idx / kBN, c = idx % kBN;
      const int kk = k0 + kr, cc = col0 + c;
      Ws[kr][c] = (kk < Hi && cc < Ho) ? __ldg(w + (int64_t)kk * w_sk + (int64_t)cc * w_sc) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k4 = 0; k4 < kBK; k4 += 4) {
      float4 xv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        xv[i] = *reinterpret_cast<const float4*>(&Xs[rg + 16 * i][k4]);
      float4 wv[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) wv[q] = *reinterpret_cast<const float4*>(&Ws[k4 + q][4 * cg]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float xs[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[i][0] = fmaf(xs[q], wv[q].x, acc[i][0]);
          acc[i][1] = fmaf(xs[q], wv[q].y, acc[i][1]);
          acc[i][2] = fmaf(xs[q], wv[q].z, acc[i][2]);
          acc[i][3] = fmaf(xs[q], wv[q].w, acc[i][3]);
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue ----
  const int cc = col0 + 4 * cg;
  float bv[4] = {0.f, 0.f, 0.f, 0.f};
  if (bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (cc + j < Ho) bv[j] = __ldg(bias + cc + j);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gr = row0 + rg + 16 * i;
    if (gr >= N) continue;
    const float rs = row_scale ? __ldg(row_scale + gr) : 1.f;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = acc[i][j] + bv[j];
    if (y_vec4 && cc + 3 < Ho) {
      if (add) {
        const float4 av = __ldg(reinterpret_cast<const float4*>(add + gr * Ho + cc));
        o[0] += av.x; o[1] += av.y; o[2] += av.z; o[3] += av.w;
      }
      if (act == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = o[j] < 0.f ? 0.f : o[j];
      }
      if (row_scale) {
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] *= rs;
      }
      *reinterpret_cast<float4*>(y + gr * Ho + cc) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (cc + j < Ho) {
          float v = o[j];
          if (add) v += __ldg(add + gr * Ho + cc + j);
          if (act == 1) v = v < 0.f ? 0.f : v;
          if (row_scale) v *= rs;
          y[gr * Ho + cc + j] = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient: partial[p][k][c] = sum over the slab's rows of x[n,k] g[n,c]; then fixed-order
// reduction over p.  CTA = 256 threads = 4 row lanes x 64 threads, each thread a 4x4 block of a
// 32x32 tile of dW; a slab is walked in 64-row stages.
// ------------------------------------------------------------------------------------------------
constexpr int kWgRows = 64;
constexpr int kWgLd = 36;

__global__ void __launch_bounds__(256)
    k_wgrad_partial(const float* __restrict__ x, int64_t N, int Hi, const float* __restrict__ g,
                    const float* __restrict__ gmask, int Ho, int64_t rows_per_slab,
                    float* __restrict__ partial, float* __restrict__ partial_b, int x_vec4,
                    int g_vec4) {
  __shared__ __align__(16) float Xs[kWgRows][kWgLd];
  __shared__ __align__(16) float Gs[kWgRows][kWgLd];
  __shared__ float red[4][32][33];
  const int tid = threadIdx.x;
  const int lane_r = tid >> 6;   // 0..3: takes rows lane_r, lane_r+4, ... of a stage
  const int t64 = tid & 63;
  const int a4 = t64 >> 3;       // dW rows 4*a4..4*a4+3 (k index)
  const int c4 = t64 & 7;        // dW cols 4*c4..4*c4+3
  const int kt0 = blockIdx.y * 32;
  const int ct0 = blockIdx.z * 32;
  const int64_t n0 = (int64_t)blockIdx.x * rows_per_slab;
  const int64_t n1 = n0 + rows_per_slab < N ? n0 + rows_per_slab : N;

  float acc[4][4];
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t s0 = n0; s0 < n1; s0 += kWgRows) {
    // stage 64 rows of x[:, kt0:kt0+32] and g[:, ct0:ct0+32]
    if (x_vec4) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = (tid >> 3) + 32 * i, q = tid & 7;
        const int64_t gr = s0 + r;
        const int kk = kt0 + 4 * q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < n1 && kk < Hi) v = __ldg(reinterpret_cast<const float4*>(x + gr * Hi + kk));
        *reinterpret_cast<float4*>(&Xs[r][4 * q]) = v;
      }
    } else {
      for (int idx = tid; idx < kWgRows * 32; idx += 256) {
        const int r = idx >> 5, c = idx & 31;
        const int64_t gr = s0 + r;
        const int kk = kt0 + c;
        Xs[r][c] = (gr < n1 && kk < Hi) ? __ldg(x + gr * Hi + kk) : 0.f;
      }
    }
    if (g_vec4) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = (tid >> 3) + 32 * i, q = tid & 7;
        const int64_t gr = s0 + r;
        const int cc = ct0 + 4 * q;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gr < n1 && cc < Ho) {
          v = __ldg(reinterpret_cast<const float4*>(g + gr * Ho + cc));
          if (gmask) {
            const float4 m = __ldg(reinterpret_cast<const float4*>(gmask + gr * Ho + cc));
            v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f;
            v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
          }
        }
        *reinterpret_cast<float4*>(&Gs[r][4 * q]) = v;
      }
    } else {
      for (int idx = tid; idx < kWgRows * 32; idx += 256) {
        const int r = idx >> 5, c = idx & 31;
        const int64_t gr = s0 + r;
        const int cc = ct0 + c;
        float v = 0.f;
        if (gr < n1 && cc < Ho) {
          v = __ldg(g + gr * Ho + cc);
          if (gmask && !(__ldg(gmask + gr * Ho + cc) > 0.f)) v = 0.f;
        }
        Gs[r][c] = v;
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int r = lane_r; r < kWgRows; r += 4) {
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[r][4 * a4]);
      const float4 gv = *reinterpret_cast<const float4*>(&Gs[r][4 * c4]);
      const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(xs[i], gv.x, acc[i][0]);
        acc[i][1] = fmaf(xs[i], gv.y, acc[i][1]);
        acc[i][2] = fmaf(xs[i], gv.z, acc[i][2]);
        acc[i][3] = fmaf(xs[i], gv.w, acc[i][3]);
      }
      if (a4 == 0) {
        bsum[0] += gv.x; bsum[1] += gv.y; bsum[2] += gv.z; bsum[3] += gv.w;
      }
    }
    __syncthreads();
  }
  // combine the 4 row lanes in fixed order
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[lane_r][4 * a4 + i][4 * c4 + j] = acc[i][j];
  __syncthreads();
  float* pout = partial + ((int64_t)blockIdx.x * Hi) * Ho;
  for (int idx = tid; idx < 32 * 32; idx += 256) {
    const int kr = idx >> 5, c = idx & 31;
    const float s = ((red[0][kr][c] + red[1][kr][c]) + red[2][kr][c]) + red[3][kr][c];
    const int kk = kt0 + kr, cc = ct0 + c;
    if (kk < Hi && cc < Ho) pout[(int64_t)kk * Ho + cc] = s;
  }
  if (partial_b != nullptr && blockIdx.y == 0) {
    __syncthreads();
    if (a4 == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) red[lane_r][0][4 * c4 + j] = bsum[j];
    }
    __syncthreads();
    if (tid < 32) {
      const float s = ((red[0][0][tid] + red[1][0][tid]) + red[2][0][tid]) + red[3][0][tid];
      const int cc = ct0 + tid;
      if (cc < Ho) partial_b[(int64_t)blockIdx.x * Ho + cc] = s;
    }
  }
}

// out(k,c) = sum_p partial[p][k][c], p ascending, 4 interleaved chains combined in fixed order
__global__ void __launch_bounds__(256)
    k_wgrad_reduce(const float* __restrict__ partial, int P, int64_t count, int Ho, float* dw,
                   int64_t dw_sk, int64_t dw_sc) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int p = 0;
  for (; p + 4 <= P; p += 4) {
    s0 += partial[(int64_t)(p + 0) * count + i];
    s1 += partial[(int64_t)(p + 1) * count + i];
    s2 += partial[(int64_t)(p + 2) * count + i];
    s3 += partial[(int64_t)(p + 3) * count + i];
  }
  for (; p < P; ++p) s0 += partial[(int64_t)p * count + i];
  const float s = (s0 + s1) + (s2 + s3);
  if (Ho > 0) {
    const int64_t k = i / Ho, c = i % Ho;
    dw[k * dw_sk + c * dw_sc] = s;
  } else {
    dw[i] = s;
  }
}

__global__ void __launch_bounds__(256) k_relu_backward(const float* __restrict__ g,
                                                       const float* __restrict__ y, int64_t count,
                                                       float* __restrict__ gin) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    gin[i] = y[i] > 0.f ? g[i] : 0.f;
}

// out[n,c] = rs[n] * g[n,c] * (m1[n,c] > 0) * (m2[n,c] > 0)   (either mask / the scale may be absent)
__global__ void __launch_bounds__(256)
    k_masked_scale(const float4* __restrict__ g, const float4* __restrict__ m1,
                   const float4* __restrict__ m2, const float* __restrict__ rs, int64_t count4,
                   int h4, float4* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count4; i += stride) {
    float4 v = __ldg(g + i);
    if (m1) {
      const float4 m = __ldg(m1 + i);
      v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f;
      v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
    }
    if (m2) {
      const float4 m = __ldg(m2 + i);
      v.x = m.x > 0.f ? v.x : 0.f; v.y = m.y > 0.f ? v.y : 0.f;
      v.z = m.z > 0.f ? v.z : 0.f; v.w = m.w > 0.f ? v.w : 0.f;
    }
    if (rs) {
      const float s = __ldg(rs + i / h4);
      v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    }
    out[i] = v;
  }
}

static int wgrad_slabs(int64_t N) {
  // enough slabs to fill the machine for a 32x32 tile, each at least one 64-row stage
  int64_t p = kNumSMs * 4;
  const int64_t max_p = ceil_div(N > 0 ? N : 1, kWgRows);
  if (p > max_p) p = max_p;
  return (int)p;
}

}  // namespace mgcn

using namespace mgcn;

extern "C" int mgcn_linear_ex(const float* x, const float* xmask, int64_t N, int64_t Hi,
                              const float* w, int64_t w_sk, int64_t w_sc, int64_t Ho,
                              const float* bias, const float* add, int act, const float* row_scale,
                              float* y, void* stream) {
  MGCN_REQUIRE(N >= 0 && N < (int64_t(1) << 31), MGCN_ERR_RANGE);
  MGCN_REQUIRE(Hi >= 1 && Ho >= 1 && Hi <= 65536 && Ho <= 65536, MGCN_ERR_SHAPE);
  MGCN_REQUIRE(act == 0 || act == 1, MGCN_ERR_SHAPE);
  if (N == 0) return MGCN_OK;
  MGCN_REQUIRE(x && w && y, MGCN_ERR_NULL);
  if (narrow_linear_applies(Hi, Ho, x, xmask, add, y))
    return launch_narrow_linear(x, xmask, N, Hi, w, w_sk, w_sc, Ho, bias, add, act, row_scale, y, stream);
  const int x_vec4 = (Hi % 4 == 0) && aligned16(x) && (!xmask || aligned16(xmask));
  const int y_vec4 = (Ho % 4 == 0) && aligned16(y) && (!add || aligned16(add));
  dim3 grid((unsigned)ceil_div(N, kBM), (unsigned)ceil_div(Ho, kBN));
  MGCN_LAUNCH(k_linear, grid, 128, 0, stream, x, xmask, N, (int)Hi, w, w_sk, w_sc, (int)Ho, bias,
              add, act, row_scale, y, x_vec4, y_vec4);
  return MGCN_OK;
}

extern "C" int mgcn_linear(const float* x, int64_t N, int64_t Hi, const float* w, int64_t w_sk,
                           int64_t w_sc, int64_t Ho, const float* bias, const float* add, int act,
                           float* y, void* stream) {
  return mgcn_linear_ex(x, nullptr, N, Hi, w, w_sk, w_sc, Ho, bias, add, act, nullptr, y, stream);
}

extern "C" int mgcn_linear_wgrad_ex(const float* x, int64_t N, int64_t Hi, const float* g,
                                    const float* gmask, int64_t Ho, float* dw, int64_t dw_sk,
                                    int64_t dw_sc, float* db, void* workspace,
                                    size_t* workspace_bytes, void* stream) {
  MGCN_REQUIRE(workspace_bytes != nullptr, MGCN_ERR_NULL);
  MGCN_REQUIRE(N >= 0 && N < (int64_t(1) << 31), MGCN_ERR_RANGE);
  MGCN_REQUIRE(Hi >= 1 && Ho >= 1 && Hi <= 65536 && Ho <= 65536, MGCN_ERR_SHAPE);
  const int P = wgrad_slabs(N);
  size_t p_max = (size_t)P > narrow_wgrad_blocks(N) ? (size_t)P : narrow_wgrad_blocks(N);
  if (Hi <= 256 && (size_t)wide_wgrad_splits(Hi) > p_max) p_max = (size_t)wide_wgrad_splits(Hi);
  WorkspaceCarver ws(workspace);
  float* partial = ws.take<float>(p_max * Hi * Ho);
  float* partial_b = ws.take<float>(p_max * Ho);
  if (workspace == nullptr) {
    *workspace_bytes = ws.bytes();
    return MGCN_OK;
  }
  MGCN_REQUIRE(*workspace_bytes >= ws.bytes(), MGCN_ERR_WORKSPACE);
  MGCN_REQUIRE(dw != nullptr, MGCN_ERR_NULL);
  MGCN_REQUIRE(N == 0 || (x && g), MGCN_ERR_NULL);
  if (N > 0 && narrow_wgrad_applies(Hi, Ho, x, g, gmask))
    return launch_narrow_wgrad(x, N, Hi, g, gmask, Ho, dw, dw_sk, dw_sc, db, partial, partial_b, stream);
  if (N > 0 && wide_wgrad_applies(N, Hi, Ho, x, g, gmask))   // tcgen05 transposed product (wgrad_wide.cu)
    return launch_wide_wgrad(x, N, Hi, g, gmask, Ho, dw, dw_sk, dw_sc, db, partial, partial_b, stream);
  const int64_t rows_per_slab = ceil_div(ceil_div(N > 0 ? N : 1, P), kWgRows) * kWgRows;
  const int x_vec4 = (Hi % 4 == 0) && aligned16(x);
  const int g_vec4 = (Ho % 4 == 0) && aligned16(g) && (!gmask || aligned16(gmask));
  dim3 grid((unsigned)P, (unsigned)ceil_div(Hi, 32), (unsigned)ceil_div(Ho, 32));
  MGCN_LAUNCH(k_wgrad_partial, grid, 256, 0, stream, x, N, (int)Hi, g, gmask, (int)Ho,
              rows_per_slab, partial, db ? partial_b : nullptr, x_vec4, g_vec4);
  const int64_t count = Hi * Ho;
  int rc = launch_reduce_partials(partial, P, (int)count, (int)Ho, dw, dw_sk, dw_sc, stream);
  if (rc != MGCN_OK) return rc;
  if (db) rc = launch_reduce_partials(partial_b, P, (int)Ho, (int)Ho, db, 0, 1, stream);
  return rc;
}

extern "C" int mgcn_linear_wgrad(const float* x, int64_t N, int64_t Hi, const float* g, int64_t Ho,
                                 float* dw, int64_t dw_sk, int64_t dw_sc, float* db,
                                 void* workspace, size_t* workspace_bytes, void* stream) {
  return mgcn_linear_wgrad_ex(x, N, Hi, g, nullptr, Ho, dw, dw_sk, dw_sc, db, workspace,
                              workspace_bytes, stream);
}

extern "C" int mgcn_masked_scale(const float* g, const float* m1, const float* m2,
                                 const float* row_scale, int64_t N, int64_t H, float* out,
                                 void* stream) {
  MGCN_REQUIRE(N >= 0 && H >= 4 && H % 4 == 0, MGCN_ERR_SHAPE);
  if (N == 0) return MGCN_OK;
  MGCN_REQUIRE(g && out, MGCN_ERR_NULL);
  MGCN_REQUIRE(aligned16(g) && aligned16(out) && (!m1 || aligned16(m1)) && (!m2 || aligned16(m2)),
               MGCN_ERR_ALIGN);
  const int64_t count4 = N * H / 4;
  int64_t blocks = ceil_div(count4, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  MGCN_LAUNCH(k_masked_scale, (unsigned)blocks, 256, 0, stream,
              reinterpret_cast<const float4*>(g), reinterpret_cast<const float4*>(m1),
              reinterpret_cast<const float4*>(m2), row_scale, count4, (int)(H / 4),
              reinterpret_cast<float4*>(out));
  return MGCN_OK;
}

extern "C" int mgcn_relu_backward(const float* g, const float* y, int64_t count, float* g_in,
                                  void* stream) {
  MGCN_REQUIRE(count >= 0, MGCN_ERR_RANGE);
  if (count == 0) return MGCN_OK;
  MGCN_REQUIRE(g && y && g_in, MGCN_ERR_NULL);
  int64_t blocks = ceil_div(count, 256);
  if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
  MGCN_LAUNCH(k_relu_backward, (unsigned)blocks, 256, 0, stream, g, y, count, g_in);
  return MGCN_OK;
}
