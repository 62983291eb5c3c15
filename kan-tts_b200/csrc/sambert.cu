// SAM-BERT (kantts/models/sambert) kernels that are not convolutions / GEMMs: LayerNorm, the masked
// multi-head scaled-dot-product attention with d_head <= 64 (probabilities materialised, as the reference
// returns them), the FSMN depthwise memory block and the LengthRegulator gather.  The GEMM-shaped work of
// the model (QKV / output projections, conv feed-forward, prenets, FSMN feed-forward) runs through the
// conv kernels (conv_tc.cu / conv_ffma.cu) as kernel-size-1/3 convolutions over the same (B, L, C) rows.
//
// Everything here is exact fp32 on CUDA cores: with d_head = 16 (sambert_24k.yaml) one attention head is a
// K=16 contraction -- a single tcgen05 k-step -- and the kernels are bound by the probability-matrix
// traffic (B*H*Lq*Lk*4 bytes written in forward, read twice in backward), not by math.
#include <atomic>

#include "common.cuh"

namespace kt {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim, one warp per row (sambert/__init__.py:64,131,197; kantts_sambert.py:58,129)
// ------------------------------------------------------------------------------------------------
template <int NC>  // columns per lane, c <= 32*NC
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, float* __restrict__ y, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out, int rows, int c, float eps) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int r = warp; r < rows; r += nwarps) {
    const float* xr = x + (long long)r * c;
    float v[NC];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int col = lane + 32 * i;
      v[i] = col < c ? __ldg(xr + col) : 0.f;
      s += v[i];
    }
    const float mean = warp_sum(s) / (float)c;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int col = lane + 32 * i;
      const float d = col < c ? v[i] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / (float)c + eps);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int col = lane + 32 * i;
      if (col < c) y[(long long)r * c + col] = fmaf((v[i] - mean) * rstd, __ldg(gamma + col), __ldg(beta + col));
    }
    if (lane == 0) {
      mean_out[r] = mean;
      rstd_out[r] = rstd;
    }
  }
}

// dx = rstd * (g*dy - mean_c(g*dy) - xhat * mean_c(g*dy*xhat));  per-CTA partial column sums of
// dy*xhat (dgamma) and dy (dbeta) go to `partial` [gridDim.x][2][c], reduced by colsum2_kernel.
template <int NC>
__global__ void layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                     const float* __restrict__ gamma, const float* __restrict__ mean,
                                     const float* __restrict__ rstd, float* __restrict__ dx, float* __restrict__ partial,
                                     int rows, int c) {
  extern __shared__ float sm[];  // [warps][2][c]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float g[NC], ag[NC], ab[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int col = lane + 32 * i;
    g[i] = col < c ? __ldg(gamma + col) : 0.f;
    ag[i] = ab[i] = 0.f;
  }
  for (int r = blockIdx.x * nw + w; r < rows; r += gridDim.x * nw) {
    const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
    float xh[NC], gd[NC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int col = lane + 32 * i;
      const float d = col < c ? __ldg(dy + (long long)r * c + col) : 0.f;
      xh[i] = col < c ? (__ldg(x + (long long)r * c + col) - mu) * rs : 0.f;
      gd[i] = g[i] * d;
      s1 += gd[i];
      s2 = fmaf(gd[i], xh[i], s2);
      ag[i] = fmaf(d, xh[i], ag[i]);
      ab[i] += d;
    }
    s1 = warp_sum(s1) / (float)c;
    s2 = warp_sum(s2) / (float)c;
#pragma unroll
    for (int i = 0; i < NC; ++i) {
      const int col = lane + 32 * i;
      if (col < c) dx[(long long)r * c + col] = rs * (gd[i] - s1 - xh[i] * s2);
    }
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) {
    const int col = lane + 32 * i;
    if (col < c) {
      sm[(w * 2 + 0) * c + col] = ag[i];
      sm[(w * 2 + 1) * c + col] = ab[i];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * c; i += blockDim.x) {
    float t = 0.f;
    for (int ww = 0; ww < nw; ++ww) t += sm[ww * 2 * c + i];
    partial[(long long)blockIdx.x * 2 * c + i] = t;
  }
}

// out[i] = sum_p partial[p][i]   (deterministic second stage of the column reductions)
__global__ void colsum_partials_kernel(const float* __restrict__ partial, float* __restrict__ out, int nparts, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = 0.f;
  for (int p = 0; p < nparts; ++p) t += __ldg(partial + (long long)p * n + i);
  out[i] = t;
}

// partial rows are [2][c]: dgamma then dbeta
__global__ void ln_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                 int nparts, int c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * c) return;
  float t = 0.f;
  for (int p = 0; p < nparts; ++p) t += __ldg(partial + (long long)p * 2 * c + i);
  if (i < c) dgamma[i] = t; else dbeta[i - c] = t;
}

static int ln_grid(int rows, int warps_per_block) {
  const int want = (rows + warps_per_block - 1) / warps_per_block;
  return want < 1 ? 1 : (want > 148 * 2 ? 148 * 2 : want);
}

long long layernorm_bwd_workspace(int rows, int c) { return (long long)ln_grid(rows, 8) * 2 * c; }

template <int NC>
static int ln_fwd_launch(const float* x, const float* g, const float* b, float* y, float* mean, float* rstd, int rows,
                         int c, float eps, cudaStream_t st) {
  const int blocks = (rows + 7) / 8;
  layernorm_fwd_kernel<NC><<<blocks > 148 * 8 ? 148 * 8 : blocks, 256, 0, st>>>(x, g, b, y, mean, rstd, rows, c, eps);
  KT_CHECK_CUDA(cudaGetLastError());
  return KT_OK;
}

int layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int rows,
                  int c, float eps, cudaStream_t st) {
  KT_REQUIRE(x && gamma && beta && y && mean && rstd, "layernorm_fwd: null pointer");
  KT_REQUIRE(rows >= 0 && c >= 1 && c <= 1024, "layernorm_fwd: need 1 <= C <= 1024 (got %d)", c);
  if (rows == 0) return KT_OK;
  if (c <= 32) return ln_fwd_launch<1>(x, gamma, beta, y, mean, rstd, rows, c, eps, st);
  if (c <= 64) return ln_fwd_launch<2>(x, gamma, beta, y, mean, rstd, rows, c, eps, st);
  if (c <= 128) return ln_fwd_launch<4>(x, gamma, beta, y, mean, rstd, rows, c, eps, st);
  if (c <= 256) return ln_fwd_launch<8>(x, gamma, beta, y, mean, rstd, rows, c, eps, st);
  if (c <= 512) return ln_fwd_launch<16>(x, gamma, beta, y, mean, rstd, rows, c, eps, st);
  return ln_fwd_launch<32>(x, gamma, beta, y, mean, rstd, rows, c, eps, st);
}

template <int NC>
static int ln_bwd_launch(const float* dy, const float* x, const float* g, const float* mean, const float* rstd,
                         float* dx, float* ws, int rows, int c, int grid, cudaStream_t st) {
  layernorm_bwd_kernel<NC><<<grid, 256, 8 * 2 * c * sizeof(float), st>>>(dy, x, g, mean, rstd, dx, ws, rows, c);
  KT_CHECK_CUDA(cudaGetLastError());
  return KT_OK;
}

int layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd, float* dx,
                  float* dgamma, float* dbeta, float* workspace, long long workspace_floats, int rows, int c,
                  cudaStream_t st) {
  KT_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta && workspace, "layernorm_bwd: null pointer");
  KT_REQUIRE(rows >= 1 && c >= 1 && c <= 1024, "layernorm_bwd: need rows >= 1, 1 <= C <= 1024");
  const int grid = ln_grid(rows, 8);
  if (workspace_floats < (long long)grid * 2 * c) {
    set_error("layernorm_bwd: workspace too small");
    return KT_ERR_WORKSPACE;
  }
  int rc;
  if (c <= 32) rc = ln_bwd_launch<1>(dy, x, gamma, mean, rstd, dx, workspace, rows, c, grid, st);
  else if (c <= 64) rc = ln_bwd_launch<2>(dy, x, gamma, mean, rstd, dx, workspace, rows, c, grid, st);
  else if (c <= 128) rc = ln_bwd_launch<4>(dy, x, gamma, mean, rstd, dx, workspace, rows, c, grid, st);
  else if (c <= 256) rc = ln_bwd_launch<8>(dy, x, gamma, mean, rstd, dx, workspace, rows, c, grid, st);
  else if (c <= 512) rc = ln_bwd_launch<16>(dy, x, gamma, mean, rstd, dx, workspace, rows, c, grid, st);
  else rc = ln_bwd_launch<32>(dy, x, gamma, mean, rstd, dx, workspace, rows, c, grid, st);
  if (rc) return rc;
  ln_reduce_kernel<<<(2 * c + 127) / 128, 128, 0, st>>>(workspace, dgamma, dbeta, grid, c);
  KT_CHECK_CUDA(cudaGetLastError());
  return KT_OK;
}

// ------------------------------------------------------------------------------------------------
// Attention.  q/k/v/out are (B, L, *) row tensors addressed with explicit row strides and a per-head
// column offset h*D, so the (n*b, L, d) permute copies of the reference (sambert/__init__.py:85-100)
// never exist.  probs is the reference's `attn`: (H*B, Lq, Lk), head-major.
// ------------------------------------------------------------------------------------------------
constexpr int kKeyTile = 256;

struct AttnArgs {
  const float *q, *k, *v;
  const unsigned char* mask;
  float *out, *probs;
  const unsigned char* keep;   // attention dropout keep mask [(h*B+b)][Lq][Lk] or null
  float* probs_dropped;        // optional copy of the dropped probabilities (what the reference returns)
  float keep_scale;
  int B, H, Lq, Lk;
  int q_stride, k_stride, v_stride, o_stride;
  long long mask_b_stride;
  int mask_q_stride;
  float scale;
  int rows_per_warp;
};

template <int D>
__global__ void __launch_bounds__(256) attn_fwd_kernel(AttnArgs a) {
  extern __shared__ float sm[];
  constexpr int DP = D + 1;
  const int QT = 8 * a.rows_per_warp;
  float* s_sc = sm;                         // [QT][Lk]  scores -> probabilities
  float* s_kv = sm + (size_t)QT * a.Lk;     // [kKeyTile][DP]
  float* s_q = s_kv + kKeyTile * DP;        // [QT][D]
  const int bh = blockIdx.y, h = bh / a.B, b = bh % a.B;
  const int q0 = blockIdx.x * QT;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const float* qb = a.q + (long long)b * a.Lq * a.q_stride + h * D;
  const float* kb = a.k + (long long)b * a.Lk * a.k_stride + h * D;
  const float* vb = a.v + (long long)b * a.Lk * a.v_stride + h * D;
  for (int i = threadIdx.x; i < QT * D; i += blockDim.x) {
    const int r = i / D, d = i % D;
    s_q[i] = (q0 + r < a.Lq) ? __ldg(qb + (long long)(q0 + r) * a.q_stride + d) * a.scale : 0.f;
  }
  // pass 1: raw scores
  for (int kt = 0; kt < a.Lk; kt += kKeyTile) {
    const int nk = min(kKeyTile, a.Lk - kt);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * D; i += blockDim.x) {
      const int j = i / D, d = i % D;
      s_kv[j * DP + d] = __ldg(kb + (long long)(kt + j) * a.k_stride + d);
    }
    __syncthreads();
    for (int rr = 0; rr < a.rows_per_warp; ++rr) {
      const int r = w * a.rows_per_warp + rr;
      if (q0 + r >= a.Lq) break;
      float qr[D];
#pragma unroll
      for (int d = 0; d < D; ++d) qr[d] = s_q[r * D + d];
      const unsigned char* mrow = a.mask ? a.mask + b * a.mask_b_stride + (long long)(q0 + r) * a.mask_q_stride : nullptr;
      for (int j = lane; j < nk; j += 32) {
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) acc = fmaf(qr[d], s_kv[j * DP + d], acc);
        if (mrow && mrow[kt + j]) acc = -INFINITY;
        s_sc[(size_t)r * a.Lk + kt + j] = acc;
      }
    }
  }
  __syncwarp();
  // softmax per row (each warp owns its rows), probabilities to HBM
  for (int rr = 0; rr < a.rows_per_warp; ++rr) {
    const int r = w * a.rows_per_warp + rr;
    if (q0 + r >= a.Lq) break;
    float* sc = s_sc + (size_t)r * a.Lk;
    float m = -INFINITY;
    for (int j = lane; j < a.Lk; j += 32) m = fmaxf(m, sc[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < a.Lk; j += 32) {
      const float e = expf(sc[j] - m);
      sc[j] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.f / s;
    const long long prow = ((long long)bh * a.Lq + q0 + r) * a.Lk;
    float* pr = a.probs + prow;
    for (int j = lane; j < a.Lk; j += 32) {
      float p = sc[j] * inv;
      pr[j] = p;
      if (a.keep) {
        p = a.keep[prow + j] ? p * a.keep_scale : 0.f;
        if (a.probs_dropped) a.probs_dropped[prow + j] = p;
      }
      sc[j] = p;
    }
  }
  // pass 2: out = P V.  Lane l owns dims (l % D)[+32..] of key subset l / D.
  constexpr int G = D >= 32 ? 1 : 32 / D;       // key subsets per warp
  constexpr int DPL = D > 32 ? D / 32 : 1;      // dims per lane
  const int dim0 = D >= 32 ? lane : lane % D;
  const int sub = D >= 32 ? 0 : lane / D;
  float acc[4][DPL];
#pragma unroll
  for (int rr = 0; rr < 4; ++rr)
#pragma unroll
    for (int e = 0; e < DPL; ++e) acc[rr][e] = 0.f;
  for (int kt = 0; kt < a.Lk; kt += kKeyTile) {
    const int nk = min(kKeyTile, a.Lk - kt);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * D; i += blockDim.x) {
      const int j = i / D, d = i % D;
      s_kv[j * DP + d] = __ldg(vb + (long long)(kt + j) * a.v_stride + d);
    }
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
      if (rr >= a.rows_per_warp) break;
      const int r = w * a.rows_per_warp + rr;
      if (q0 + r >= a.Lq) break;
      const float* sc = s_sc + (size_t)r * a.Lk + kt;
      for (int j = sub; j < nk; j += G) {
        const float p = sc[j];
#pragma unroll
        for (int e = 0; e < DPL; ++e) acc[rr][e] = fmaf(p, s_kv[j * DP + dim0 + 32 * e], acc[rr][e]);
      }
    }
  }
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    if (rr >= a.rows_per_warp) break;
    const int r = w * a.rows_per_warp + rr;
    if (q0 + r >= a.Lq) break;
#pragma unroll
    for (int e = 0; e < DPL; ++e) {
      float t = acc[rr][e];
      if (G > 1) {
#pragma unroll
        for (int o = D; o < 32; o <<= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      }
      if (sub == 0) a.out[((long long)b * a.Lq + q0 + r) * a.o_stride + h * D + dim0 + 32 * e] = t;
    }
  }
}

struct AttnBwdArgs {
  const float *q, *k, *v, *probs, *dout;
  const unsigned char* keep;
  float keep_scale;
  float *dq, *dk, *dv, *delta;
  int B, H, Lq, Lk;
  int q_stride, k_stride, v_stride, o_stride;   // strides of q/dq, k/dk, v/dv, dout
  float scale;
  int accum_dq;
};

// per query row: delta = sum_j P dP,  dQ = scale * sum_j P (dP - delta) K[j],  dP[j] = dO . V[j]
// one warp per row; lanes over keys, K/V tiles staged in shared memory.
template <int D>
__global__ void __launch_bounds__(256) attn_bwd_q_kernel(AttnBwdArgs a) {
  extern __shared__ float sm[];
  constexpr int DP = D + 1;
  float* s_k = sm;                      // [kKeyTile][DP]
  float* s_v = s_k + kKeyTile * DP;     // [kKeyTile][DP]
  float* s_dp = s_v + kKeyTile * DP;    // [8][Lk]   dP of the warp's current row
  const int bh = blockIdx.y, h = bh / a.B, b = bh % a.B;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r = blockIdx.x * 8 + w;
  const bool live = r < a.Lq;
  const float* kb = a.k + (long long)b * a.Lk * a.k_stride + h * D;
  const float* vb = a.v + (long long)b * a.Lk * a.v_stride + h * D;
  float dor[D];
#pragma unroll
  for (int d = 0; d < D; ++d)
    dor[d] = live ? __ldg(a.dout + ((long long)b * a.Lq + r) * a.o_stride + h * D + d) : 0.f;
  const long long prow = ((long long)bh * a.Lq + (live ? r : 0)) * a.Lk;
  const float* pr = a.probs + prow;
  const unsigned char* kr = a.keep ? a.keep + prow : nullptr;
  float* dp = s_dp + (size_t)w * a.Lk;
  float delta = 0.f;
  for (int kt = 0; kt < a.Lk; kt += kKeyTile) {
    const int nk = min(kKeyTile, a.Lk - kt);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * D; i += blockDim.x) {
      const int j = i / D, d = i % D;
      s_v[j * DP + d] = __ldg(vb + (long long)(kt + j) * a.v_stride + d);
    }
    __syncthreads();
    if (live)
      for (int j = lane; j < nk; j += 32) {
        float acc = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) acc = fmaf(dor[d], s_v[j * DP + d], acc);
        if (kr) acc = kr[kt + j] ? acc * a.keep_scale : 0.f;   // d/dP through the dropout
        dp[kt + j] = acc;
        delta = fmaf(__ldg(pr + kt + j), acc, delta);
      }
  }
  delta = warp_sum(delta);
  if (live && lane == 0) a.delta[(long long)bh * a.Lq + r] = delta;
  float dqr[D];
#pragma unroll
  for (int d = 0; d < D; ++d) dqr[d] = 0.f;
  for (int kt = 0; kt < a.Lk; kt += kKeyTile) {
    const int nk = min(kKeyTile, a.Lk - kt);
    __syncthreads();
    for (int i = threadIdx.x; i < nk * D; i += blockDim.x) {
      const int j = i / D, d = i % D;
      s_k[j * DP + d] = __ldg(kb + (long long)(kt + j) * a.k_stride + d);
    }
    __syncthreads();
    if (live)
      for (int j = lane; j < nk; j += 32) {
        const float ds = __ldg(pr + kt + j) * (dp[kt + j] - delta);
#pragma unroll
        for (int d = 0; d < D; ++d) dqr[d] = fmaf(ds, s_k[j * DP + d], dqr[d]);
      }
  }
#pragma unroll
  for (int d = 0; d < D; ++d) dqr[d] = warp_sum(dqr[d]) * a.scale;
  if (live && lane == 0) {
    float* o = a.dq + ((long long)b * a.Lq + r) * a.q_stride + h * D;
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = a.accum_dq ? o[d] + dqr[d] : dqr[d];
  }
}

// per key j (one thread each, 128 keys per CTA): dV[j] = sum_i P[i,j] dO[i],
// dK[j] = scale * sum_i P[i,j] (dO[i].V[j] - delta[i]) Q[i]; the query rows stream through shared memory.
constexpr int kBwdKeys = 128;
constexpr int kBwdQTile = 32;

template <int D>
__global__ void __launch_bounds__(kBwdKeys) attn_bwd_kv_kernel(AttnBwdArgs a) {
  __shared__ float s_q[kBwdQTile][D];
  __shared__ float s_do[kBwdQTile][D];
  __shared__ float s_delta[kBwdQTile];
  const int bh = blockIdx.y, h = bh / a.B, b = bh % a.B;
  const int j = blockIdx.x * kBwdKeys + threadIdx.x;
  const bool live = j < a.Lk;
  float vr[D], dkr[D], dvr[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    vr[d] = live ? __ldg(a.v + ((long long)b * a.Lk + j) * a.v_stride + h * D + d) : 0.f;
    dkr[d] = dvr[d] = 0.f;
  }
  for (int i0 = 0; i0 < a.Lq; i0 += kBwdQTile) {
    const int nq = min(kBwdQTile, a.Lq - i0);
    __syncthreads();
    for (int t = threadIdx.x; t < nq * D; t += blockDim.x) {
      const int i = t / D, d = t % D;
      s_q[i][d] = __ldg(a.q + ((long long)b * a.Lq + i0 + i) * a.q_stride + h * D + d);
      s_do[i][d] = __ldg(a.dout + ((long long)b * a.Lq + i0 + i) * a.o_stride + h * D + d);
    }
    if (threadIdx.x < nq) s_delta[threadIdx.x] = __ldg(a.delta + (long long)bh * a.Lq + i0 + threadIdx.x);
    __syncthreads();
    if (live) {
      const long long pofs = ((long long)bh * a.Lq + i0) * a.Lk + j;
      const float* pc = a.probs + pofs;
      const unsigned char* kc = a.keep ? a.keep + pofs : nullptr;
      for (int i = 0; i < nq; ++i) {
        const float p = __ldg(pc + (long long)i * a.Lk);
        const float ks = kc ? (kc[(long long)i * a.Lk] ? a.keep_scale : 0.f) : 1.f;
        const float pd = p * ks;
        float dpv = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          dpv = fmaf(s_do[i][d], vr[d], dpv);
          dvr[d] = fmaf(pd, s_do[i][d], dvr[d]);
        }
        const float ds = p * (dpv * ks - s_delta[i]);
#pragma unroll
        for (int d = 0; d < D; ++d) dkr[d] = fmaf(ds, s_q[i][d], dkr[d]);
      }
    }
  }
  if (live) {
    float* ok = a.dk + ((long long)b * a.Lk + j) * a.k_stride + h * D;
    float* ov = a.dv + ((long long)b * a.Lk + j) * a.v_stride + h * D;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      ok[d] = dkr[d] * a.scale;
      ov[d] = dvr[d];
    }
  }
}

static int attn_check(const KtAttnDesc* d) {
  KT_REQUIRE(d, "attention: null descriptor");
  KT_REQUIRE(d->batch >= 1 && d->heads >= 1 && d->lq >= 1 && d->lk >= 1, "attention: bad sizes");
  KT_REQUIRE(d->d_head == 8 || d->d_head == 16 || d->d_head == 32 || d->d_head == 64,
             "attention: d_head must be 8, 16, 32 or 64 (got %d)", d->d_head);
  KT_REQUIRE(d->lk <= 2048, "attention: Lk <= 2048 (got %d)", d->lk);
  const int hd = d->heads * d->d_head;
  KT_REQUIRE(d->q_stride >= hd && d->k_stride >= hd && d->v_stride >= hd && d->o_stride >= hd,
             "attention: row strides smaller than heads*d_head");
  KT_REQUIRE((long long)d->heads * d->batch <= 65535, "attention: heads*batch <= 65535");
  return KT_OK;
}

template <int D>
static int attn_fwd_launch(const KtAttnDesc* d, AttnArgs a, cudaStream_t st) {
  a.rows_per_warp = d->lk <= 512 ? 4 : (d->lk <= 1024 ? 2 : 1);
  const int QT = 8 * a.rows_per_warp;
  const size_t smem = ((size_t)QT * d->lk + (size_t)kKeyTile * (D + 1) + (size_t)QT * D) * sizeof(float);
  static std::atomic<bool> attr_set{false};
  if (!attr_set.exchange(true)) KT_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  KT_REQUIRE(smem <= (size_t)kMaxDynSmem, "attention: shared memory budget exceeded");
  dim3 grid((d->lq + QT - 1) / QT, d->heads * d->batch);
  attn_fwd_kernel<D><<<grid, 256, smem, st>>>(a);
  KT_CHECK_CUDA(cudaGetLastError());
  return KT_OK;
}

int attention_fwd(const KtAttnDesc* d, const float* q, const float* k, const float* v, const unsigned char* mask,
                  const unsigned char* keep, float* out, float* probs, float* probs_dropped, cudaStream_t st) {
  int rc = attn_check(d);
  if (rc) return rc;
  KT_REQUIRE(q && k && v && out && probs, "attention_fwd: null pointer");
  KT_REQUIRE(!keep || (d->keep_scale >= 1.f), "attention_fwd: keep mask given but keep_scale < 1");
  AttnArgs a{q, k, v, mask, out, probs, keep, probs_dropped, d->keep_scale, d->batch, d->heads, d->lq, d->lk, d->q_stride, d->k_stride, d->v_stride,
             d->o_stride, d->mask_b_stride, d->mask_q_stride, d->scale, 4};
  switch (d->d_head) {
    case 8: return attn_fwd_launch<8>(d, a, st);
    case 16: return attn_fwd_launch<16>(d, a, st);
    case 32: return attn_fwd_launch<32>(d, a, st);
    default: return attn_fwd_launch<64>(d, a, st);
  }
}

template <int D>
static int attn_bwd_launch(const KtAttnDesc* d, AttnBwdArgs a, cudaStream_t st) {
  const size_t smem = ((size_t)2 * kKeyTile * (D + 1) + (size_t)8 * d->lk) * sizeof(float);
  static std::atomic<bool> attr_set{false};
  if (!attr_set.exchange(true)) KT_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_q_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
  KT_REQUIRE(smem <= (size_t)kMaxDynSmem, "attention: shared memory budget exceeded");
  dim3 gq((d->lq + 7) / 8, d->heads * d->batch);
  attn_bwd_q_kernel<D><<<gq, 256, smem, st>>>(a);
  KT_CHECK_CUDA(cudaGetLastError());
  dim3 gk((d->lk + kBwdKeys - 1) / kBwdKeys, d->heads * d->batch);
  attn_bwd_kv_kernel<D><<<gk, kBwdKeys, 0, st>>>(a);
  KT_CHECK_CUDA(cudaGetLastError());
  return KT_OK;
}

int attention_bwd(const KtAttnDesc* d, const float* q, const float* k, const float* v, const float* probs,
                  const unsigned char* keep,
                  const float* dout, float* dq, float* dk, float* dv, float* delta, int accum_dq, cudaStream_t st) {
  int rc = attn_check(d);
  if (rc) return rc;
  KT_REQUIRE(q && k && v && probs && dout && dq && dk && dv && delta, "attention_bwd: null pointer");
  AttnBwdArgs a{q, k, v, probs, dout, keep, d->keep_scale, dq, dk, dv, delta, d->batch, d->heads, d->lq, d->lk, d->q_stride, d->k_stride,
                d->v_stride, d->o_stride, d->scale, accum_dq};
  switch (d->d_head) {
    case 8: return attn_bwd_launch<8>(d, a, st);
    case 16: return attn_bwd_launch<16>(d, a, st);
    case 32: return attn_bwd_launch<32>(d, a, st);
    default: return attn_bwd_launch<64>(d, a, st);
  }
}

// ------------------------------------------------------------------------------------------------
// FSMN memory block (fsmn.py:46-77): depthwise FIR over time with asymmetric zero padding + skip,
// padded frames zeroed on the way in and on the way out.
//   xm = x * keep;   y[b,t,c] = keep[b,t] * ( xm[b,t,c] + sum_j w[c][j] * xm[b, t + j - lp, c] )
// ------------------------------------------------------------------------------------------------
__global__ void fsmn_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                const unsigned char* __restrict__ mask, float* __restrict__ y, int B, int T, int C,
                                int K, int lp) {
  const long long total = (long long)B * T * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bt = i / C;
    const int t = (int)(bt % T), b = (int)(bt / T);
    if (mask && mask[bt]) {
      y[i] = 0.f;
      continue;
    }
    float acc = __ldg(x + i);
    const float* wc = w + (long long)c * K;
    for (int j = 0; j < K; ++j) {
      const int s = t + j - lp;
      if (s < 0 || s >= T) continue;
      if (mask && mask[(long long)b * T + s]) continue;
      acc = fmaf(__ldg(wc + j), __ldg(x + ((long long)b * T + s) * C + c), acc);
    }
    y[i] = acc;
  }
}

// dx[b,s,c] = keep[b,s] * ( dym[b,s,c] + sum_j w[c][j] * dym[b, s - j + lp, c] ),  dym = dy * keep
__global__ void fsmn_bwd_data_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                     const unsigned char* __restrict__ mask, float* __restrict__ dx, int B, int T, int C,
                                     int K, int lp) {
  const long long total = (long long)B * T * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bt = i / C;
    const int s = (int)(bt % T), b = (int)(bt / T);
    if (mask && mask[bt]) {
      dx[i] = 0.f;
      continue;
    }
    float acc = __ldg(dy + i);
    const float* wc = w + (long long)c * K;
    for (int j = 0; j < K; ++j) {
      const int t = s - j + lp;
      if (t < 0 || t >= T) continue;
      if (mask && mask[(long long)b * T + t]) continue;
      acc = fmaf(__ldg(wc + j), __ldg(dy + ((long long)b * T + t) * C + c), acc);
    }
    dx[i] = acc;
  }
}

// dw[c][j] = sum_{b,t} dym[b,t,c] * xm[b, t + j - lp, c].  CTA = 32 channels x 8 warps (taps strided over
// warps), one (batch item, time chunk) per blockIdx.y; partials [chunks][C*K] reduced by colsum_partials.
constexpr int kFsmnChunk = 256;
__global__ void fsmn_bwd_weight_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                       const unsigned char* __restrict__ mask, float* __restrict__ partial, int B, int T,
                                       int C, int K, int lp, int chunks_per_b) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int b = blockIdx.y / chunks_per_b, t0 = (blockIdx.y % chunks_per_b) * kFsmnChunk;
  const int t1 = min(T, t0 + kFsmnChunk);
  if (c >= C) return;
  for (int j = w; j < K; j += 8) {
    float acc = 0.f;
    for (int t = t0; t < t1; ++t) {
      const int s = t + j - lp;
      if (s < 0 || s >= T) continue;
      if (mask && (mask[(long long)b * T + t] || mask[(long long)b * T + s])) continue;
      acc = fmaf(__ldg(dy + ((long long)b * T + t) * C + c), __ldg(x + ((long long)b * T + s) * C + c), acc);
    }
    partial[(long long)blockIdx.y * C * K + (long long)c * K + j] = acc;
  }
}

long long fsmn_bwd_workspace(int B, int T, int C, int K) {
  return (long long)B * ((T + kFsmnChunk - 1) / kFsmnChunk) * C * K;
}

static int grid_for(long long n, int threads) {
  long long blocks = (n + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

int fsmn_fwd(const float* x, const float* w, const unsigned char* mask, float* y, int B, int T, int C, int K, int lp,
             cudaStream_t st) {
  KT_REQUIRE(x && w && y && B >= 1 && T >= 1 && C >= 1 && K >= 1 && lp >= 0, "fsmn_fwd: bad arguments");
  fsmn_fwd_kernel<<<grid_for((long long)B * T * C, 256), 256, 0, st>>>(x, w, mask, y, B, T, C, K, lp);
  KT_CHECK_CUDA(cudaGetLastError());
  return KT_OK;
}

int fsmn_bwd(const float* x, const float* dy, const float* w, const unsigned char* mask, float* dx, float* dw,
             float* workspace, long long workspace_floats, int B, int T, int C, int K, int lp, cudaStream_t st) {
  KT_REQUIRE(x && dy && w && B >= 1 && T >= 1 && C >= 1 && K >= 1 && lp >= 0, "fsmn_bwd: bad arguments");
  if (dx) {
    fsmn_bwd_data_kernel<<<grid_for((long long)B * T * C, 256), 256, 0, st>>>(dy, w, mask, dx, B, T, C, K, lp);
    KT_CHECK_CUDA(cudaGetLastError());
  }
  if (dw) {
    const int cpb = (T + kFsmnChunk - 1) / kFsmnChunk;
    if (!workspace || workspace_floats < fsmn_bwd_workspace(B, T, C, K)) {
      set_error("fsmn_bwd: workspace too small");
      return KT_ERR_WORKSPACE;
    }
    KT_REQUIRE(B * cpb <= 65535, "fsmn_bwd: too many chunks");
    fsmn_bwd_weight_kernel<<<dim3((C + 31) / 32, B * cpb), 256, 0, st>>>(x, dy, mask, workspace, B, T, C, K, lp, cpb);
    KT_CHECK_CUDA(cudaGetLastError());
    colsum_partials_kernel<<<(C * K + 127) / 128, 128, 0, st>>>(workspace, dw, B * cpb, C * K);
    KT_CHECK_CUDA(cudaGetLastError());
  }
  return KT_OK;
}

// ------------------------------------------------------------------------------------------------
// LengthRegulator (adaptors.py:15-37) as a row gather instead of the one-hot matmul:
//   out[b,t,:] = idx[b,t] >= 0 ? in[b, idx[b,t], :] : 0          (idx = -1: beyond the rounded durations,
//   masked output frame, or r-padding)
// backward: din[b,i,:] = sum over the token's contiguous span [start[b,i], start[b,i]+count[b,i]) of the
// frames whose idx is i (masked frames inside the span carry idx -1).
// ------------------------------------------------------------------------------------------------
__global__ void rows_gather_fwd_kernel(const float* __restrict__ in, const int* __restrict__ idx, float* __restrict__ out,
                                       int B, int T_out, int T_in, int C) {
  const long long total = (long long)B * T_out * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bt = i / C;
    const int b = (int)(bt / T_out);
    const int src = __ldg(idx + bt);
    out[i] = src >= 0 ? __ldg(in + ((long long)b * T_in + src) * C + c) : 0.f;
  }
}

__global__ void rows_gather_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ idx,
                                       const int* __restrict__ start, const int* __restrict__ count,
                                       float* __restrict__ din, int B, int T_out, int T_in, int C) {
  const long long total = (long long)B * T_in * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long bi = i / C;
    const int tok = (int)(bi % T_in), b = (int)(bi / T_in);
    const int s0 = __ldg(start + bi), n = __ldg(count + bi);
    float acc = 0.f;
    for (int t = s0; t < s0 + n && t < T_out; ++t)
      if (__ldg(idx + (long long)b * T_out + t) == tok) acc += __ldg(dout + ((long long)b * T_out + t) * C + c);
    din[i] = acc;
  }
}

int rows_gather_fwd(const float* in, const int* idx, float* out, int B, int T_out, int T_in, int C, cudaStream_t st) {
  KT_REQUIRE(in && idx && out && B >= 1 && T_out >= 1 && T_in >= 1 && C >= 1, "rows_gather_fwd: bad arguments");
  rows_gather_fwd_kernel<<<grid_for((long long)B * T_out * C, 256), 256, 0, st>>>(in, idx, out, B, T_out, T_in, C);
  KT_CHECK_CUDA(cudaGetLastError());
  return KT_OK;
}

int rows_gather_bwd(const float* dout, const int* idx, const int* start, const int* count, float* din, int B, int T_out,
                    int T_in, int C, cudaStream_t st) {
  KT_REQUIRE(dout && idx && start && count && din && B >= 1 && T_out >= 1 && T_in >= 1 && C >= 1,
             "rows_gather_bwd: bad arguments");
  rows_gather_bwd_kernel<<<grid_for((long long)B * T_in * C, 256), 256, 0, st>>>(dout, idx, start, count, din, B, T_out,
                                                                                 T_in, C);
  KT_CHECK_CUDA(cudaGetLastError());
  return KT_OK;
}

}  // namespace kt
