// Dense transforms with one narrow side (<= 4 columns): the first layer of the botnet model
// (x is [N,1]: gcn_base_models.py:201 and the residual nn.Linear(1,32), gcn_model.py:64,96), the
// final projection nn.Linear(32,2) (gcn_model.py:73,108) and their autograd.  These are pure
// streaming passes over the wide operand; the 128x32 register-tiled FMA kernel of dense.cu spends
// 0.45-0.53 ms on each at the botnet batch where the bytes allow 0.08-0.16 ms
// (profiles/r1_layer_summary.md).  All reductions are fixed-order (deterministic, no atomics).
#include "common.cuh"

namespace mgcn {

// y[n, c..c+3] = rs[n] * act( sum_{k<Hi} x[n,k]*(xmask[n,k]>0) * W(k,c..) + bias + add ), Hi <= 4.
// One thread per float4 of y.
__global__ void __launch_bounds__(256)
    k_linear_small_in(const float* __restrict__ x, const float* __restrict__ xmask, int64_t N, int Hi,
                      const float* __restrict__ w, int64_t w_sk, int64_t w_sc, int Ho,
                      const float* __restrict__ bias, const float* __restrict__ add, int act,
                      const float* __restrict__ row_scale, float* __restrict__ y) {
  const int q4 = Ho >> 2;
  const int64_t total = N * q4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / q4;
    const int c = (int)(i - n * q4) * 4;
    float o[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < Hi; ++k) {
      float xv = __ldg(x + n * Hi + k);
      if (xmask && !(__ldg(xmask + n * Hi + k) > 0.f)) xv = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = fmaf(xv, __ldg(w + k * w_sk + (c + j) * w_sc), o[j]);
    }
    if (bias) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] += __ldg(bias + c + j);
    }
    if (add) {
      const float4 av = __ldg(reinterpret_cast<const float4*>(add + n * Ho + c));
      o[0] += av.x; o[1] += av.y; o[2] += av.z; o[3] += av.w;
    }
    if (act == 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = o[j] < 0.f ? 0.f : o[j];
    }
    if (row_scale) {
      const float rs = __ldg(row_scale + n);
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] *= rs;
    }
    *reinterpret_cast<float4*>(y + n * Ho + c) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// Same contract when 256 % (Ho/4) == 0: a thread keeps one 4-column chunk for all its rows, so its Hi x 4
// weights and the bias sit in registers, there is no per-element division, and 4 rows are in flight per
// thread (the generic kernel above ran at 1.9 TB/s of stores at the botnet batch).
__global__ void __launch_bounds__(256)
    k_linear_small_in_fixed(const float* __restrict__ x, const float* __restrict__ xmask, int64_t N, int Hi,
                            const float* __restrict__ w, int64_t w_sk, int64_t w_sc, int Ho,
                            const float* __restrict__ bias, const float* __restrict__ add, int act,
                            const float* __restrict__ row_scale, float* __restrict__ y) {
  const int q4 = Ho >> 2, rpb = 256 / q4;
  const int c = (threadIdx.x % q4) * 4, rl = threadIdx.x / q4;
  float wv[4][4], bv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) wv[k][j] = k < Hi ? __ldg(w + k * w_sk + (c + j) * w_sc) : 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) bv[j] = bias ? __ldg(bias + c + j) : 0.f;
  const uint64_t pol = policy_evict_first();
  const int64_t step = (int64_t)gridDim.x * rpb;
  for (int64_t n0 = (int64_t)blockIdx.x * rpb + rl; n0 < N; n0 += 4 * step) {
    float xv[4][4], rs[4];
    float4 av[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t n = n0 + u * step;
      rs[u] = 1.f;
      av[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) xv[u][k] = 0.f;
      if (n < N) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < Hi) {
            xv[u][k] = __ldg(x + n * Hi + k);
            if (xmask && !(__ldg(xmask + n * Hi + k) > 0.f)) xv[u][k] = 0.f;
          }
        }
        if (add) av[u] = ld_f4_hint(add + n * Ho + c, pol);
        if (row_scale) rs[u] = __ldg(row_scale + n);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t n = n0 + u * step;
      if (n < N) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < Hi) {
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = fmaf(xv[u][k], wv[k][j], o[j]);
          }
        }
        if (bias) {
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] += bv[j];
        }
        if (add) { o[0] += av[u].x; o[1] += av[u].y; o[2] += av[u].z; o[3] += av[u].w; }
        if (act == 1) {
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] = o[j] < 0.f ? 0.f : o[j];
        }
        if (row_scale) {
#pragma unroll
          for (int j = 0; j < 4; ++j) o[j] *= rs[u];
        }
        st_f4_hint(y + n * Ho + c, make_float4(o[0], o[1], o[2], o[3]), pol);
      }
    }
  }
}

// y[n, c] = rs[n] * act( sum_k x[n,k]*(xmask>0) * W(k,c) + bias[c] + add[n,c] ), Ho <= 4, Hi = 4*L with L a
// power of two <= 32: L lanes per row, float4 each, butterfly sum (fixed order).
template <int L>
__global__ void __launch_bounds__(256)
    k_linear_small_out(const float* __restrict__ x, const float* __restrict__ xmask, int64_t N,
                       const float* __restrict__ w, int64_t w_sk, int64_t w_sc, int Ho,
                       const float* __restrict__ bias, const float* __restrict__ add, int act,
                       const float* __restrict__ row_scale, float* __restrict__ y) {
  constexpr int Hi = 4 * L;
  const int sub = threadIdx.x % L;
  const int64_t rows_per_iter = (int64_t)gridDim.x * (blockDim.x / L);
  float wv[4][4];  // [output c][k within this lane's float4]
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int q = 0; q < 4; ++q) wv[c][q] = c < Ho ? __ldg(w + (4 * sub + q) * w_sk + c * w_sc) : 0.f;
  const int64_t n_iters = (N + rows_per_iter - 1) / rows_per_iter;   // uniform trip count: shuffles stay converged
  for (int64_t it = 0; it < n_iters; ++it) {
    const int64_t n = it * rows_per_iter + (int64_t)blockIdx.x * (blockDim.x / L) + threadIdx.x / L;
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) {
      xv = __ldg(reinterpret_cast<const float4*>(x + n * Hi + 4 * sub));
      if (xmask) {
        const float4 m = __ldg(reinterpret_cast<const float4*>(xmask + n * Hi + 4 * sub));
        xv.x = m.x > 0.f ? xv.x : 0.f; xv.y = m.y > 0.f ? xv.y : 0.f;
        xv.z = m.z > 0.f ? xv.z : 0.f; xv.w = m.w > 0.f ? xv.w : 0.f;
      }
    }
    float o[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float s = fmaf(xv.x, wv[c][0], fmaf(xv.y, wv[c][1], fmaf(xv.z, wv[c][2], xv.w * wv[c][3])));
#pragma unroll
      for (int d = L / 2; d >= 1; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
      o[c] = s;
    }
    if (n < N && sub == 0) {
      const float rs = row_scale ? __ldg(row_scale + n) : 1.f;
      for (int c = 0; c < Ho; ++c) {
        float v = o[c];
        if (bias) v += __ldg(bias + c);
        if (add) v += __ldg(add + n * Ho + c);
        if (act == 1) v = v < 0.f ? 0.f : v;
        y[n * Ho + c] = v * rs;
      }
    }
  }
}

// partial[b][k][c] = sum over the rows of block b of a[n,k]*(amask>0) * b_[n,c]*(bmask>0),
// k < Ka <= 4 (narrow operand), c < Hb = 4*L (wide operand, L lanes per row, float4 each);
// partial_a[b][k] = column sums of the narrow operand, partial_b[b][c] of the wide one.
template <int L>
__global__ void __launch_bounds__(256)
    k_wgrad_narrow(const float* __restrict__ a, const float* __restrict__ amask, int Ka,
                   const float* __restrict__ bw, const float* __restrict__ bmask, int64_t N,
                   float* __restrict__ partial, float* __restrict__ partial_a,
                   float* __restrict__ partial_b) {
  constexpr int Hb = 4 * L;
  constexpr int RPB = 256 / L;  // rows per block iteration
  __shared__ float red[RPB][4][Hb + 1];
  __shared__ float red_b[RPB][Hb + 1];
  __shared__ float red_a[RPB][4];
  const int sub = threadIdx.x % L, rl = threadIdx.x / L;
  float acc[4][4];
  float bs[4] = {0.f, 0.f, 0.f, 0.f}, as[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[k][j] = 0.f;
  // 4 rows of this thread are fetched before the first is used (one float4 in flight per thread left the
  // pass at 2.5 TB/s); rows are still accumulated in increasing order
  const int64_t step = (int64_t)gridDim.x * RPB;
  for (int64_t n0 = (int64_t)blockIdx.x * RPB + rl; n0 < N; n0 += 4 * step) {
    float4 bvs[4], ms[4];
    float avs[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t n = n0 + u * step;
      bvs[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      ms[u] = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) avs[u][k] = 0.f;
      if (n < N) {
        bvs[u] = __ldg(reinterpret_cast<const float4*>(bw + n * Hb + 4 * sub));
        if (bmask) ms[u] = __ldg(reinterpret_cast<const float4*>(bmask + n * Hb + 4 * sub));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < Ka) {
            avs[u][k] = __ldg(a + n * Ka + k);
            if (amask && !(__ldg(amask + n * Ka + k) > 0.f)) avs[u][k] = 0.f;
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (n0 + u * step < N) {
        float4 bv = bvs[u];
        if (bmask) {
          const float4 m = ms[u];
          bv.x = m.x > 0.f ? bv.x : 0.f; bv.y = m.y > 0.f ? bv.y : 0.f;
          bv.z = m.z > 0.f ? bv.z : 0.f; bv.w = m.w > 0.f ? bv.w : 0.f;
        }
        bs[0] += bv.x; bs[1] += bv.y; bs[2] += bv.z; bs[3] += bv.w;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < Ka) {
            const float av = avs[u][k];
            acc[k][0] = fmaf(av, bv.x, acc[k][0]);
            acc[k][1] = fmaf(av, bv.y, acc[k][1]);
            acc[k][2] = fmaf(av, bv.z, acc[k][2]);
            acc[k][3] = fmaf(av, bv.w, acc[k][3]);
            if (sub == 0) as[k] += av;
          }
        }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[rl][k][4 * sub + j] = acc[k][j];
#pragma unroll
  for (int j = 0; j < 4; ++j) red_b[rl][4 * sub + j] = bs[j];
  if (sub == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) red_a[rl][k] = as[k];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < Ka * Hb; idx += 256) {
    const int k = idx / Hb, c = idx % Hb;
    float s = 0.f;
    for (int r = 0; r < RPB; ++r) s += red[r][k][c];
    partial[((int64_t)blockIdx.x * Ka + k) * Hb + c] = s;
  }
  if (partial_b) {
    for (int c = threadIdx.x; c < Hb; c += 256) {
      float s = 0.f;
      for (int r = 0; r < RPB; ++r) s += red_b[r][c];
      partial_b[(int64_t)blockIdx.x * Hb + c] = s;
    }
  }
  if (partial_a && threadIdx.x < Ka) {
    float s = 0.f;
    for (int r = 0; r < RPB; ++r) s += red_a[r][threadIdx.x];
    partial_a[(int64_t)blockIdx.x * Ka + threadIdx.x] = s;
  }
}

// out[f(i)] = sum_p partial[p*count + i]: 8 warps split p into 8 contiguous ranges (4 interleaved
// chains each), combined in fixed order.  f maps i = k*Hc + c to out[k*sk + c*sc].
__global__ void __launch_bounds__(256)
    k_reduce_partials(const float* __restrict__ partial, int P, int count, int Hc, float* __restrict__ out,
                      int64_t sk, int64_t sc) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const int per = (P + 7) / 8;
  const int p0 = warp * per, p1 = min(P, p0 + per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (i < count) {
    int p = p0;
    for (; p + 4 <= p1; p += 4) {
      s0 += partial[(int64_t)(p + 0) * count + i];
      s1 += partial[(int64_t)(p + 1) * count + i];
      s2 += partial[(int64_t)(p + 2) * count + i];
      s3 += partial[(int64_t)(p + 3) * count + i];
    }
    for (; p < p1; ++p) s0 += partial[(int64_t)p * count + i];
  }
  red[warp][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (warp == 0 && i < count) {
    float s = red[0][lane];
#pragma unroll
    for (int w = 1; w < 8; ++w) s += red[w][lane];
    const int k = i / Hc, c = i % Hc;
    out[k * sk + c * sc] = s;
  }
}

int launch_reduce_partials(const float* partial, int P, int count, int Hc, float* out, int64_t sk,
                           int64_t sc, void* stream) {
  MGCN_LAUNCH(k_reduce_partials, (count + 31) / 32, 256, 0, stream, partial, P, count, Hc, out, sk, sc);
  return MGCN_OK;
}

static int stream_grid(int64_t work_items, int per_block) {
  int64_t b = ceil_div(work_items > 0 ? work_items : 1, per_block);
  const int64_t cap = (int64_t)kNumSMs * 8;
  return (int)(b < cap ? b : cap);
}

// ---- dispatch helpers used by dense.cu ------------------------------------------------------------
bool narrow_linear_applies(int64_t Hi, int64_t Ho, const float* x, const float* xmask, const float* add,
                           const float* y) {
  if (Hi <= 4 && Ho % 4 == 0 && aligned16(y) && (!add || aligned16(add))) return true;
  if (Ho <= 4 && (Hi == 16 || Hi == 32 || Hi == 64 || Hi == 128) && aligned16(x) &&
      (!xmask || aligned16(xmask)))
    return true;
  return false;
}

int launch_narrow_linear(const float* x, const float* xmask, int64_t N, int64_t Hi, const float* w,
                         int64_t w_sk, int64_t w_sc, int64_t Ho, const float* bias, const float* add,
                         int act, const float* row_scale, float* y, void* stream) {
  if (Hi <= 4 && Ho % 4 == 0 && Ho <= 1024 && 256 % (Ho / 4) == 0) {
    const int rpb = 256 / (int)(Ho / 4);
    int64_t blocks = ceil_div(N > 0 ? N : 1, (int64_t)rpb * 4);
    if (blocks > (int64_t)kNumSMs * 8) blocks = (int64_t)kNumSMs * 8;
    MGCN_LAUNCH(k_linear_small_in_fixed, (unsigned)blocks, 256, 0, stream, x, xmask, N, (int)Hi, w, w_sk, w_sc,
                (int)Ho, bias, add, act, row_scale, y);
    return MGCN_OK;
  }
  if (Hi <= 4 && Ho % 4 == 0) {
    MGCN_LAUNCH(k_linear_small_in, stream_grid(N * (Ho / 4), 256), 256, 0, stream, x, xmask, N, (int)Hi,
                w, w_sk, w_sc, (int)Ho, bias, add, act, row_scale, y);
    return MGCN_OK;
  }
  switch (Hi) {
    case 16:
      MGCN_LAUNCH((k_linear_small_out<4>), stream_grid(N, 64), 256, 0, stream, x, xmask, N, w, w_sk, w_sc,
                  (int)Ho, bias, add, act, row_scale, y);
      break;
    case 32:
      MGCN_LAUNCH((k_linear_small_out<8>), stream_grid(N, 32), 256, 0, stream, x, xmask, N, w, w_sk, w_sc,
                  (int)Ho, bias, add, act, row_scale, y);
      break;
    case 64:
      MGCN_LAUNCH((k_linear_small_out<16>), stream_grid(N, 16), 256, 0, stream, x, xmask, N, w, w_sk, w_sc,
                  (int)Ho, bias, add, act, row_scale, y);
      break;
    default:
      MGCN_LAUNCH((k_linear_small_out<32>), stream_grid(N, 8), 256, 0, stream, x, xmask, N, w, w_sk, w_sc,
                  (int)Ho, bias, add, act, row_scale, y);
      break;
  }
  return MGCN_OK;
}

bool narrow_wgrad_applies(int64_t Hi, int64_t Ho, const float* x, const float* g, const float* gmask) {
  auto wide_ok = [](int64_t h) { return h == 16 || h == 32 || h == 64 || h == 128; };
  if (Hi <= 4 && wide_ok(Ho) && aligned16(g) && (!gmask || aligned16(gmask))) return true;
  if (Ho <= 4 && wide_ok(Hi) && aligned16(x)) return true;
  return false;
}

size_t narrow_wgrad_blocks(int64_t N) {
  int64_t b = ceil_div(N > 0 ? N : 1, 64);
  const int64_t cap = (int64_t)kNumSMs * 4;
  return (size_t)(b < cap ? b : cap);
}

// dW(k,c) = sum_n x[n,k] * (g*(gmask>0))[n,c] written at dw[k*dw_sk + c*dw_sc]; db[c] = sum_n g*(mask)
// ws: [P][Knarrow*Hwide] + [P][Ho]
int launch_narrow_wgrad(const float* x, int64_t N, int64_t Hi, const float* g, const float* gmask,
                        int64_t Ho, float* dw, int64_t dw_sk, int64_t dw_sc, float* db, float* part,
                        float* part_db, void* stream) {
  const int P = (int)narrow_wgrad_blocks(N);
  const bool x_narrow = Hi <= 4;
  const int Ka = (int)(x_narrow ? Hi : Ho), Hb = (int)(x_narrow ? Ho : Hi);
  const float* a = x_narrow ? x : g;
  const float* amask = x_narrow ? nullptr : gmask;
  const float* b = x_narrow ? g : x;
  const float* bmask = x_narrow ? gmask : nullptr;
  float* pa = (!x_narrow && db) ? part_db : nullptr;   // db is the column sum of g
  float* pb = (x_narrow && db) ? part_db : nullptr;
  switch (Hb) {
    case 16: MGCN_LAUNCH((k_wgrad_narrow<4>), P, 256, 0, stream, a, amask, Ka, b, bmask, N, part, pa, pb); break;
    case 32: MGCN_LAUNCH((k_wgrad_narrow<8>), P, 256, 0, stream, a, amask, Ka, b, bmask, N, part, pa, pb); break;
    case 64: MGCN_LAUNCH((k_wgrad_narrow<16>), P, 256, 0, stream, a, amask, Ka, b, bmask, N, part, pa, pb); break;
    default: MGCN_LAUNCH((k_wgrad_narrow<32>), P, 256, 0, stream, a, amask, Ka, b, bmask, N, part, pa, pb); break;
  }
  // partial index i = ka*Hb + cb.  x narrow: (k,c) = (ka, cb); g narrow: (k,c) = (cb, ka)
  int rc = launch_reduce_partials(part, P, Ka * Hb, Hb, dw, x_narrow ? dw_sk : dw_sc,
                                  x_narrow ? dw_sc : dw_sk, stream);
  if (rc != MGCN_OK) return rc;
  if (db) rc = launch_reduce_partials(part_db, P, (int)Ho, (int)Ho, db, 0, 1, stream);
  return rc;
}

}  // namespace mgcn
