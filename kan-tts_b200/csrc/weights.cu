// Weight re-parametrisation (weight_norm / plain-with-scale) fused with the layout change from
// the reference parameter layout to the kernel layouts of conv_ffma.cu / conv_tc.cu.
// Replaces torch._weight_norm + its backward (kantts/models/hifigan/layers.py:29,67,105,139).
#include "common.cuh"

namespace kt {

// index of reference element (a, b, j) [shape (d0, d1, k)] in the two kernel layouts
struct WLayout {
  int d0, d1, k, transposed, groups;
  __device__ __forceinline__ void map(int a, int b, int j, long long& i_fwd, long long& i_bwd) const {
    if (!transposed) {  // a = co, b = ci_l
      const int cout = d0, cin_g = d1, cout_g = d0 / groups, cin = d1 * groups;
      const int gi = a / cout_g, co_l = a % cout_g;
      i_fwd = ((long long)j * cin_g + b) * cout + a;
      i_bwd = ((long long)j * cout_g + co_l) * cin + (long long)gi * cin_g + b;
    } else {  // a = ci, b = co
      const int cin = d0, cout = d1;
      i_fwd = ((long long)j * cin + a) * cout + b;
      i_bwd = ((long long)j * cout + b) * cin + a;
    }
  }
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

__global__ void weight_prepare_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                      const float* __restrict__ inv_sigma, int mode, WLayout L,
                                      float* __restrict__ w_fwd, float* __restrict__ w_bwd,
                                      float* __restrict__ norm_out, float* __restrict__ w_ref) {
  __shared__ float red[8];
  const int a = blockIdx.x;
  const int n = L.d1 * L.k;
  const float* vs = v + (long long)a * n;
  float scale;
  if (mode == 1) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) { const float t = vs[i]; ss = fmaf(t, t, ss); }
    const float nrm = sqrtf(block_sum(ss, red));
    if (threadIdx.x == 0 && norm_out) norm_out[a] = nrm;
    scale = g[a] / nrm;
  } else {
    scale = inv_sigma ? *inv_sigma : 1.f;
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int b = i / L.k, j = i % L.k;
    const float w = vs[i] * scale;
    long long i_fwd, i_bwd;
    L.map(a, b, j, i_fwd, i_bwd);
    if (w_fwd) w_fwd[i_fwd] = w;
    if (w_bwd) w_bwd[i_bwd] = w;
    if (w_ref) w_ref[(long long)a * n + i] = w;
  }
}

__global__ void weight_grad_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                   const float* __restrict__ g, const float* __restrict__ norm,
                                   const float* __restrict__ inv_sigma, int mode, WLayout L,
                                   float* __restrict__ dv, float* __restrict__ dg, int accumulate,
                                   const float* __restrict__ dbias_src, float* __restrict__ dbias_dst, int nbias) {
  __shared__ float red[8];
  const int a = blockIdx.x;
  const int n = L.d1 * L.k;
  if (dbias_dst) {   // bias gradient hand-over (+= into the parameter's .grad when accumulating)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nbias; i += gridDim.x * blockDim.x)
      dbias_dst[i] = accumulate ? dbias_dst[i] + dbias_src[i] : dbias_src[i];
  }
  const float* vs = v + (long long)a * n;
  float* dvs = dv + (long long)a * n;
  if (mode == 1) {
    float dot = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      long long i_fwd, i_bwd;
      L.map(a, i / L.k, i % L.k, i_fwd, i_bwd);
      dot = fmaf(dw[L.transposed ? i_bwd : i_fwd], vs[i], dot);
    }
    dot = block_sum(dot, red);
    const float nrm = norm[a], ga = g[a];
    if (threadIdx.x == 0) dg[a] = accumulate ? dg[a] + dot / nrm : dot / nrm;
    const float c1 = ga / nrm, c2 = ga * dot / (nrm * nrm * nrm);
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      long long i_fwd, i_bwd;
      L.map(a, i / L.k, i % L.k, i_fwd, i_bwd);
      const float r = c1 * dw[L.transposed ? i_bwd : i_fwd] - c2 * vs[i];
      dvs[i] = accumulate ? dvs[i] + r : r;
    }
  } else {
    const float scale = inv_sigma ? *inv_sigma : 1.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      long long i_fwd, i_bwd;
      L.map(a, i / L.k, i % L.k, i_fwd, i_bwd);
      const float r = dw[L.transposed ? i_bwd : i_fwd] * scale;
      dvs[i] = accumulate ? dvs[i] + r : r;
    }
  }
}

int weight_prepare(const float* v, const float* g, const float* inv_sigma, int mode, int d0, int d1, int k,
                   int transposed, int groups, float* w_fwd, float* w_bwd, float* norm_out, float* w_ref,
                   cudaStream_t st) {
  KT_REQUIRE(v && d0 > 0 && d1 > 0 && k > 0 && groups > 0, "weight_prepare: bad arguments");
  KT_REQUIRE(mode == 0 || (mode == 1 && g && norm_out), "weight_prepare: mode 1 needs g and norm_out");
  KT_REQUIRE(!transposed || groups == 1, "weight_prepare: transposed conv must have groups == 1");
  KT_REQUIRE(transposed || d0 % groups == 0, "weight_prepare: d0 %% groups != 0");
  WLayout L{d0, d1, k, transposed, groups};
  weight_prepare_kernel<<<d0, 256, 0, st>>>(v, g, inv_sigma, mode, L, w_fwd, w_bwd, norm_out, w_ref);
  KT_CHECK_CUDA(cudaGetLastError());
  return KT_OK;
}

int debug_flags();   // 2048: skip the weight-norm backward kernel (ablation)

int weight_grad(const float* dw, const float* v, const float* g, const float* norm, const float* inv_sigma,
                int mode, int d0, int d1, int k, int transposed, int groups, float* dv, float* dg, int accumulate,
                const float* dbias_src, float* dbias_dst, int nbias, cudaStream_t st) {
  KT_REQUIRE(dw && v && dv && d0 > 0 && d1 > 0 && k > 0 && groups > 0, "weight_grad: bad arguments");
  KT_REQUIRE(mode == 0 || (mode == 1 && g && norm && dg), "weight_grad: mode 1 needs g, norm, dg");
  WLayout L{d0, d1, k, transposed, groups};
  KT_REQUIRE((dbias_dst == nullptr) == (dbias_src == nullptr) && nbias >= 0, "weight_grad: dbias_src / dbias_dst must come together");
  if (!(debug_flags() & 2048)) weight_grad_kernel<<<d0, 256, 0, st>>>(dw, v, g, norm, inv_sigma, mode, L, dv, dg, accumulate, dbias_src, dbias_dst, nbias);
  KT_CHECK_CUDA(cudaGetLastError());
  return KT_OK;
}

}  // namespace kt
