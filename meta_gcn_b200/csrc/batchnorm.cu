// BatchNorm1d of the GIN convolution's MLP (kernel/gin.py:10-16: Linear -> ReLU -> Linear -> ReLU -> BatchNorm1d(hidden)),
// training and evaluation mode, forward and backward, as fixed-order reductions (deterministic, no atomics) instead of
// the cuDNN / ATen kernels the reference reaches through torch.nn.BatchNorm1d.
//
//   training forward:  mean_c = sum_n x[n,c] / N;   var_c = sum_n (x[n,c] - mean_c)^2 / N      (two passes: no
//                      cancellation);  y = (x - mean) * rstd * gamma + beta,  rstd = 1 / sqrt(var + eps);
//                      running_mean = (1 - m) running_mean + m mean;  running_var = (1 - m) running_var + m var N/(N-1)
//   eval forward:      the same affine map with the running statistics
//   backward:          dbeta_c = sum_n g;  dgamma_c = sum_n g xhat;
//                      dx = gamma rstd (g - dbeta / N - xhat dgamma / N)      (training);  dx = gamma rstd g   (eval)
//
// Column sums: grid (ceil(H / 32), S) CTAs of 32 x 8 threads; thread (c, r) sums rows r, r + 8, ... of its slice in
// order, the 8 row lanes are added in order in shared memory, the S slices in order by k_bn_finish.
#include "common.cuh"

namespace mgcn {

constexpr int kBnRows = 8;

// mode 0: f1 = x;  mode 1: f1 = (x - mean)^2;  mode 2: f1 = g, f2 = g * (x - mean) * rstd
template <int kModeT>
__global__ void __launch_bounds__(32 * kBnRows)
    k_bn_partial(const float* __restrict__ x, const float* __restrict__ g, int64_t N, int H,
                 const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ part) {
  __shared__ float red[2][kBnRows][33];
  const int c = blockIdx.x * 32 + threadIdx.x, r = threadIdx.y, S = gridDim.y;
  const int64_t per = (N + S - 1) / S;
  const int64_t beg = min((int64_t)blockIdx.y * per, N), end = min(beg + per, N);
  float s1 = 0.f, s2 = 0.f;
  if (c < H) {
    const float mu = kModeT >= 1 ? mean[c] : 0.f;
    const float rs = kModeT == 2 ? rstd[c] : 0.f;
    for (int64_t n = beg + r; n < end; n += kBnRows) {
      const float xv = __ldg(x + n * H + c);
      if (kModeT == 0) {
        s1 += xv;
      } else if (kModeT == 1) {
        const float d = xv - mu;
        s1 = fmaf(d, d, s1);
      } else {
        const float gv = __ldg(g + n * H + c);
        s1 += gv;
        s2 = fmaf(gv, (xv - mu) * rs, s2);
      }
    }
  }
  red[0][r][threadIdx.x] = s1;
  red[1][r][threadIdx.x] = s2;
  __syncthreads();
  if (r == 0 && c < H) {
    float a = red[0][0][threadIdx.x], b = red[1][0][threadIdx.x];
#pragma unroll
    for (int q = 1; q < kBnRows; ++q) {
      a += red[0][q][threadIdx.x];
      b += red[1][q][threadIdx.x];
    }
    part[((int64_t)blockIdx.y * 2 + 0) * H + c] = a;
    part[((int64_t)blockIdx.y * 2 + 1) * H + c] = b;
  }
}

// what 0: mean = s1 / N.  what 1: var = s1 / N -> rstd; running statistics.  what 2: dbeta = s1, dgamma = s2.
__global__ void __launch_bounds__(256)
    k_bn_finish(const float* __restrict__ part, int S, int H, int64_t N, int what, float eps, float momentum,
                float* __restrict__ out1, float* __restrict__ out2, const float* __restrict__ mean,
                float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= H) return;
  float s1 = 0.f, s2 = 0.f;
  for (int s = 0; s < S; ++s) {
    s1 += part[((int64_t)s * 2 + 0) * H + c];
    s2 += part[((int64_t)s * 2 + 1) * H + c];
  }
  const float inv_n = 1.f / (float)N;
  if (what == 0) {
    out1[c] = s1 * inv_n;
  } else if (what == 1) {
    const float var = s1 * inv_n;
    out1[c] = 1.f / sqrtf(var + eps);
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean[c];
    if (running_var) {
      const float unbiased = N > 1 ? s1 / (float)(N - 1) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
    }
  } else {
    out1[c] = s1;
    out2[c] = s2;
  }
}

// y = (x - mean) * rstd * gamma + beta
__global__ void __launch_bounds__(256)
    k_bn_apply(const float* __restrict__ x, int64_t total, int H, const float* __restrict__ mean,
               const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
               float* __restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % H);
    const float ga = gamma ? gamma[c] : 1.f, be = beta ? beta[c] : 0.f;
    y[i] = (x[i] - mean[c]) * rstd[c] * ga + be;
  }
}

// dx = gamma * rstd * (g - dbeta / N - xhat * dgamma / N)   (training) |  gamma * rstd * g   (eval: dbeta == NULL)
__global__ void __launch_bounds__(256)
    k_bn_bwd_apply(const float* __restrict__ x, const float* __restrict__ g, int64_t total, int H, int64_t N,
                   const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                   const float* __restrict__ dbeta, const float* __restrict__ dgamma, float* __restrict__ dx) {
  const float inv_n = 1.f / (float)N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % H);
    const float ga = gamma ? gamma[c] : 1.f;
    float v = g[i];
    if (dbeta) {
      const float xhat = (x[i] - mean[c]) * rstd[c];
      v = v - dbeta[c] * inv_n - xhat * dgamma[c] * inv_n;
    }
    dx[i] = ga * rstd[c] * v;
  }
}

static int bn_slices(int64_t N, int64_t H) {
  // enough CTAs to fill the machine, at least 256 rows per slice
  int64_t col_blocks = ceil_div(H, 32);
  int64_t s = ceil_div((int64_t)kNumSMs * 4, col_blocks);
  const int64_t by_rows = ceil_div(N > 0 ? N : 1, 256);
  if (s > by_rows) s = by_rows;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return (int)s;
}

}  // namespace mgcn

using namespace mgcn;

extern "C" int mgcn_batchnorm_fwd(const float* x, int64_t N, int64_t H, const float* gamma, const float* beta,
                                  float eps, float momentum, int training, float* running_mean, float* running_var,
                                  float* mean, float* rstd, float* y, void* workspace, size_t* workspace_bytes,
                                  void* stream) {
  MGCN_REQUIRE(workspace_bytes != nullptr, MGCN_ERR_NULL);
  MGCN_REQUIRE(N >= 0 && H >= 1 && H < (1 << 20) && N < (int64_t(1) << 40), MGCN_ERR_RANGE);
  const int S = bn_slices(N, H);
  WorkspaceCarver ws(workspace);
  float* part = ws.take<float>((size_t)S * 2 * H);
  if (workspace == nullptr) {
    *workspace_bytes = ws.bytes();
    return MGCN_OK;
  }
  MGCN_REQUIRE(*workspace_bytes >= ws.bytes(), MGCN_ERR_WORKSPACE);
  if (N == 0) return MGCN_OK;
  MGCN_REQUIRE(x && y && mean && rstd, MGCN_ERR_NULL);
  MGCN_REQUIRE(training || (running_mean && running_var), MGCN_ERR_NULL);
  const dim3 grid((unsigned)ceil_div(H, 32), (unsigned)S), block(32, kBnRows);
  const int fin_blocks = (int)ceil_div(H, 256);
  if (training) {
    MGCN_LAUNCH(k_bn_partial<0>, grid, block, 0, stream, x, nullptr, N, (int)H, nullptr, nullptr, part);
    MGCN_LAUNCH(k_bn_finish, fin_blocks, 256, 0, stream, part, S, (int)H, N, 0, eps, momentum, mean, nullptr, nullptr,
                nullptr, nullptr);
    MGCN_LAUNCH(k_bn_partial<1>, grid, block, 0, stream, x, nullptr, N, (int)H, mean, nullptr, part);
    MGCN_LAUNCH(k_bn_finish, fin_blocks, 256, 0, stream, part, S, (int)H, N, 1, eps, momentum, rstd, nullptr, mean,
                running_mean, running_var);
  } else {
    // mean = running_mean, rstd = 1 / sqrt(running_var + eps): computed by k_bn_finish from a one-slice "partial"
    MGCN_CHECK_CUDA(cudaMemcpyAsync(mean, running_mean, sizeof(float) * H, cudaMemcpyDeviceToDevice,
                                    static_cast<cudaStream_t>(stream)));
    // part[0][0][c] = running_var[c] * N  so that s1 / N = running_var
    MGCN_CHECK_CUDA(cudaMemcpyAsync(part, running_var, sizeof(float) * H, cudaMemcpyDeviceToDevice,
                                    static_cast<cudaStream_t>(stream)));
    MGCN_LAUNCH(k_bn_finish, fin_blocks, 256, 0, stream, part, 1, (int)H, (int64_t)1, 1, eps, 0.f, rstd, nullptr, mean,
                nullptr, nullptr);
  }
  const int64_t total = N * H;
  int64_t blocks = ceil_div(total, 256 * 4);
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  MGCN_LAUNCH(k_bn_apply, (unsigned)blocks, 256, 0, stream, x, total, (int)H, mean, rstd, gamma, beta, y);
  return MGCN_OK;
}

extern "C" int mgcn_batchnorm_bwd(const float* x, const float* g, int64_t N, int64_t H, const float* gamma,
                                  const float* mean, const float* rstd, int training, float* dx, float* dgamma,
                                  float* dbeta, void* workspace, size_t* workspace_bytes, void* stream) {
  MGCN_REQUIRE(workspace_bytes != nullptr, MGCN_ERR_NULL);
  MGCN_REQUIRE(N >= 0 && H >= 1 && H < (1 << 20) && N < (int64_t(1) << 40), MGCN_ERR_RANGE);
  const int S = bn_slices(N, H);
  WorkspaceCarver ws(workspace);
  float* part = ws.take<float>((size_t)S * 2 * H);
  if (workspace == nullptr) {
    *workspace_bytes = ws.bytes();
    return MGCN_OK;
  }
  MGCN_REQUIRE(*workspace_bytes >= ws.bytes(), MGCN_ERR_WORKSPACE);
  if (N == 0) return MGCN_OK;
  MGCN_REQUIRE(x && g && mean && rstd && dgamma && dbeta, MGCN_ERR_NULL);
  const dim3 grid((unsigned)ceil_div(H, 32), (unsigned)S), block(32, kBnRows);
  MGCN_LAUNCH(k_bn_partial<2>, grid, block, 0, stream, x, g, N, (int)H, mean, rstd, part);
  MGCN_LAUNCH(k_bn_finish, (int)ceil_div(H, 256), 256, 0, stream, part, S, (int)H, N, 2, 0.f, 0.f, dbeta, dgamma,
              nullptr, nullptr, nullptr);
  if (dx) {
    const int64_t total = N * H;
    int64_t blocks = ceil_div(total, 256 * 4);
    if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
    MGCN_LAUNCH(k_bn_bwd_apply, (unsigned)blocks, 256, 0, stream, x, g, total, (int)H, N, mean, rstd, gamma,
                training ? dbeta : nullptr, dgamma, dx);
  }
  return MGCN_OK;
}
