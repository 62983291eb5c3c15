// tcgen05 implicit-GEMM Conv1d (forward / data-gradient) for the GEMM-shaped HiFi-GAN layers.
//
// Precision: "bf16x3".  Every fp32 operand is split x = hi + lo (two bf16, |x - hi - lo| <~ 2^-17 |x|)
// and the product is accumulated in fp32 TMEM as hi*hi + hi*lo + lo*hi (the dropped lo*lo term is
// ~2^-18 relative).  A single-pass TF32/BF16 MMA does not meet the path's tolerance (mel-L1 <= 1e-4,
// SURVEY.md "hard parts"); three bf16 MMAs do, at twice the rate of 3xTF32.
//
// Data flow per CTA (one 128-row output tile x NT output channels, one batch item):
//   warps 0-3  stage the channels-last fp32 activation tile (128 + halo rows x 64 channels per K chunk),
//              applying the fused pre-activation / output-activation derivative, split it into the hi /
//              lo bf16 planes and store them as SWIZZLE_128B shared-memory images (rows = time steps).
//              im2col-free: tap j of the conv is the SAME image read through a UMMA descriptor whose
//              start address is shifted by tap_ioff[j] rows (the 128-byte swizzle is a function of absolute smem address bits, so a row shift needs no re-phasing).
//   warp 4     streams the pre-swizzled bf16 weight tiles (hi + lo, one tap x 64 input channels) with
//              cp.async.bulk (TMA engine) into a ring of shared-memory stages, mbarrier complete_tx.
//   warp 5     one elected thread issues tcgen05.mma (M=128, N=NT, K=16) x 4 k-slices x 3 products per
//              (chunk, tap); tcgen05.commit releases weight stages / activation images / signals the epilogue.
//   warps 0-3  epilogue: tcgen05.ld the fp32 accumulators (thread = output row), bias / activation /
//              residual (or act' mask for the data gradient), 16-byte stores to the channels-last output.
#include <algorithm>
#include <atomic>
#include <vector>

#include "common.cuh"
#include "tc_common.cuh"

namespace kt {

using namespace tc;

constexpr int kTcM = 128;        // output rows per CTA
constexpr int kTcKC = 64;        // input channels per K chunk (one 128-byte swizzle row of bf16)
constexpr int kTcMaxRows = 256;  // image rows (128 + halo) upper bound

// ---------------------------------------------------------------------------------------------
// weight packing: fp32 W[taps][K][N] (kernel layout of conv_ffma.cu) -> bf16 hi/lo SWIZZLE_128B tiles
//   block (j, kc, nt) = [hi tile | lo tile], tile = NT rows (n) x 64 (k) bf16, row = 128 bytes
// ---------------------------------------------------------------------------------------------
__global__ void tc_pack_weights_kernel(const float* __restrict__ w, int taps, int K, int N, int NT,
                                       __nv_bfloat16* __restrict__ out) {
  const int kchunks = K / kTcKC, ntiles = N / NT;
  const long long total = (long long)taps * K * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const int k = (int)((i / N) % K);
    const int j = (int)(i / ((long long)N * K));
    const float x = w[i];
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const __nv_bfloat16 lo = __float2bfloat16_rn(x - __bfloat162float(hi));
    const int kc = k / kTcKC, c = k % kTcKC, nt = n / NT, r = n % NT;
    const long long block = ((long long)j * kchunks + kc) * ntiles + nt;
    const long long base = block * (2LL * NT * kTcKC);
    const uint32_t off = (sw128_offset((uint32_t)r, (uint32_t)(c >> 3)) >> 1) + (uint32_t)(c & 7);
    out[base + off] = hi;
    out[base + (long long)NT * kTcKC + off] = lo;
  }
}

// ---------------------------------------------------------------------------------------------
// conv kernel
// ---------------------------------------------------------------------------------------------
struct TcParams {
  Side in;
  const __nv_bfloat16* wimg;
  const float* bias;
  const float* resid;
  Side mask;
  float* out;
  int batch, t_in, t_out, c_in, c_out;
  int out_act;
  float out_slope;
  int NT, ntiles, kchunks, rows, nb_stages, tmem_cols;
  int flags;  // bit0: use matrix base offset
  Phase ph;
};

constexpr int kTcThreads = 192;

__global__ void __launch_bounds__(kTcThreads, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve-up (all image / tile bases 1024-byte aligned)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int img_bytes = p.rows * 128;                 // one plane of one activation stage
  const int a_stage_bytes = 2 * img_bytes;            // hi + lo
  const int b_stage_bytes = 2 * p.NT * 128;           // hi + lo weight tile
  uint8_t* a_base = smem;
  uint8_t* b_base = a_base + 2 * a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_base + (size_t)p.nb_stages * b_stage_bytes);
  uint64_t* full_a = bars;            // [2]
  uint64_t* empty_a = bars + 2;       // [2]
  uint64_t* full_b = bars + 4;        // [nb]
  uint64_t* empty_b = full_b + p.nb_stages;
  uint64_t* tmem_full = empty_b + p.nb_stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Phase& ph = p.ph;
  const int m0 = blockIdx.x * kTcM;
  const int nt = blockIdx.y;
  const int bb = blockIdx.z;
  const int row_lo = m0 + ph.min_ioff;  // input time index of image row 0 (i_step == 1)

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&full_a[s], 128); mbar_init(&empty_a[s], 1); }
    for (int s = 0; s < p.nb_stages; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
    mbar_init(tmem_full, 1);
    mbar_fence_init();
    fence_proxy_async();
  }
  if (warp == 4) tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp < 4) {
    // ===================== activation producers =====================
    const float* in_b = p.in.p + (long long)bb * p.t_in * p.c_in;
    const float* aux_b = p.in.aux ? p.in.aux + (long long)bb * p.t_in * p.c_in : nullptr;
    for (int c = 0; c < p.kchunks; ++c) {
      const int s = c & 1;
      mbar_wait(&empty_a[s], ((c >> 1) & 1) ^ 1);
      uint8_t* img_hi = a_base + s * a_stage_bytes;
      uint8_t* img_lo = img_hi + img_bytes;
      stage_rows<5>(img_hi, img_lo, p.in, in_b, aux_b, p.c_in, c * kTcKC, row_lo, 0, p.t_in, p.rows, tid);
      fence_proxy_async();
      mbar_arrive(&full_a[s]);
    }

    // ===================== epilogue =====================
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int m = m0 + warp * 32 + lane;
    const bool valid = m < ph.M;
    const int to = ph.o_off + ph.o_step * (valid ? m : 0);
    const long long obase = ((long long)bb * p.t_out + to) * p.c_out + (long long)nt * p.NT;
    const uint32_t t_lane = tmem_acc + ((uint32_t)(warp * 32) << 16);
    for (int n0 = 0; n0 < p.NT; n0 += 32) {
      uint32_t rr[32];
      if (p.NT - n0 >= 32) {
        tmem_ld32(t_lane + (uint32_t)n0, rr);
      } else {  // NT % 32 == 16
        uint32_t r16[16];
        tmem_ld16(t_lane + (uint32_t)n0, r16);
#pragma unroll
        for (int e = 0; e < 16; ++e) { rr[e] = r16[e]; rr[16 + e] = 0u; }
      }
      tmem_ld_wait();
      if (valid) {
        const int ncols = min(32, p.NT - n0);
        for (int e = 0; e < ncols; e += 4) {
          const long long o = obase + n0 + e;
          float v[4] = {__uint_as_float(rr[e]), __uint_as_float(rr[e + 1]), __uint_as_float(rr[e + 2]), __uint_as_float(rr[e + 3])};
          if (p.bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + nt * p.NT + n0 + e));
            v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
          }
          if (p.out_act == KT_ACT_LRELU) {
#pragma unroll
            for (int z = 0; z < 4; ++z) v[z] = v[z] > 0.f ? v[z] : v[z] * p.out_slope;
          } else if (p.out_act == KT_ACT_TANH) {
#pragma unroll
            for (int z = 0; z < 4; ++z) v[z] = tanhf(v[z]);
          }
          if (p.mask.p) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p.mask.p + o));
            v[0] = side_apply(v[0], a.x, p.mask.mode, p.mask.slope);
            v[1] = side_apply(v[1], a.y, p.mask.mode, p.mask.slope);
            v[2] = side_apply(v[2], a.z, p.mask.mode, p.mask.slope);
            v[3] = side_apply(v[3], a.w, p.mask.mode, p.mask.slope);
          }
          if (p.resid) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(p.resid + o));
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
          }
          if (ph.accumulate) {
            const float4 a = *reinterpret_cast<const float4*>(p.out + o);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
          }
          *reinterpret_cast<float4*>(p.out + o) = make_float4(v[0], v[1], v[2], v[3]);
        }
      }
    }
    tc_fence_before();
  } else if (warp == 4) {
    // ===================== weight stream (bulk async copies) =====================
    if (lane == 0) {
      int it = 0;
      for (int c = 0; c < p.kchunks; ++c) {
        for (int n = 0; n < ph.ntaps; ++n, ++it) {
          const int s = it % p.nb_stages;
          const uint32_t par = ((it / p.nb_stages) & 1) ^ 1;
          mbar_wait(&empty_b[s], par);
          const long long block = ((long long)ph.tap_j[n] * p.kchunks + c) * p.ntiles + nt;
          const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wimg) + block * (long long)b_stage_bytes;
          mbar_arrive_expect_tx(&full_b[s], (uint32_t)b_stage_bytes);
          bulk_g2s(b_base + (size_t)s * b_stage_bytes, src, (uint32_t)b_stage_bytes, &full_b[s]);
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kTcM, p.NT, 0, 0);
      const bool use_bo = (p.flags & 1) != 0;
      int it = 0;
      uint32_t acc = 0;
      for (int c = 0; c < p.kchunks; ++c) {
        const int sa = c & 1;
        mbar_wait(&full_a[sa], (c >> 1) & 1);
        tc_fence_after();
        const uint32_t a_hi = smem_u32(a_base + sa * a_stage_bytes);
        const uint32_t a_lo = a_hi + (uint32_t)img_bytes;
        for (int n = 0; n < ph.ntaps; ++n, ++it) {
          const int sb = it % p.nb_stages;
          mbar_wait(&full_b[sb], (it / p.nb_stages) & 1);
          tc_fence_after();
          const uint32_t b_hi = smem_u32(b_base + (size_t)sb * b_stage_bytes);
          const uint32_t b_lo = b_hi + (uint32_t)(p.NT * 128);
          const uint32_t shift = (uint32_t)(ph.tap_ioff[n] - ph.min_ioff) * 128u;
#pragma unroll
          for (int kk = 0; kk < kTcKC / 16; ++kk) {
            const uint32_t ko = (uint32_t)kk * 32u;
            const uint64_t da_hi = smem_desc_sw128(a_hi + shift + ko, 16, 1024, use_bo);
            const uint64_t da_lo = smem_desc_sw128(a_lo + shift + ko, 16, 1024, use_bo);
            const uint64_t db_hi = smem_desc_sw128(b_hi + ko, 16, 1024, use_bo);
            const uint64_t db_lo = smem_desc_sw128(b_lo + ko, 16, 1024, use_bo);
            umma_bf16(tmem_acc, da_lo, db_hi, idesc, acc);
            acc = 1;
            umma_bf16(tmem_acc, da_hi, db_lo, idesc, 1);
            umma_bf16(tmem_acc, da_hi, db_hi, idesc, 1);
          }
          umma_commit(&empty_b[sb]);
        }
        umma_commit(&empty_a[sa]);
      }
      umma_commit(tmem_full);
    }
    __syncwarp();
  }

  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, (uint32_t)p.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int pick_nt(int n) {
  if (n % 16 != 0) return 0;
  if (n <= 256) return n;
  if (n % 256 == 0) return 256;
  if (n % 128 == 0) return 128;
  return 0;
}

// Is (direction dir: 0 fwd, 1 bwd_data) of this layer runnable on the tcgen05 kernel?  -> N tile or 0
int tc_plan(const KtConv1dDesc* d, int dir) {
  if (d->groups != 1 || d->upsample != 1) return 0;
  const int cin = dir == 0 ? d->c_in : d->c_out;     // contraction channels
  const int cout = dir == 0 ? d->c_out : d->c_in;    // produced channels
  if (cin % kTcKC != 0) return 0;
  const int nt = pick_nt(cout);
  if (nt == 0) return 0;
  // gather phases must have unit input step: conv fwd / conv dgrad with stride 1, transposed fwd (any stride)
  const bool scatter = (dir == 0) == (d->transposed != 0);
  if (!scatter && d->stride != 1) return 0;
  if (scatter && d->nsub != 1 && d->stride != 1) return 0;
  if (d->nsub != 1 && d->stride != 1) return 0;
  const long long halo = (long long)(d->kernel - 1) * d->dilation * d->nsub;
  if (kTcM + halo > kTcMaxRows) return 0;
  return nt;
}

int tc_pack_weights(const float* w, int taps, int K, int N, int NT, void* out, cudaStream_t st) {
  KT_REQUIRE(w && out && taps > 0 && K % kTcKC == 0 && NT > 0 && N % NT == 0 && NT % 16 == 0 && NT <= 256,
             "tc_pack_weights: bad shape taps=%d K=%d N=%d NT=%d", taps, K, N, NT);
  const long long total = (long long)taps * K * N;
  const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 16);
  tc_pack_weights_kernel<<<blocks, 256, 0, st>>>(w, taps, K, N, NT, reinterpret_cast<__nv_bfloat16*>(out));
  KT_CHECK_CUDA(cudaGetLastError());
  return KT_OK;
}

std::vector<Phase> conv_phases(const KtConv1dDesc* d, int dir);  // conv_ffma.cu

static int run_tc(TcParams p, cudaStream_t st) {
  const Phase& ph = p.ph;
  if (ph.M <= 0) return KT_OK;
  KT_REQUIRE(ph.i_step == 1 && ph.up == 1, "conv_tc: phase must have unit input step");
  p.rows = (kTcM + (ph.max_ioff - ph.min_ioff) + 7) & ~7;
  KT_REQUIRE(p.rows <= kTcMaxRows, "conv_tc: halo too large (%d rows)", p.rows);
  p.kchunks = p.c_in / kTcKC;
  p.ntiles = p.c_out / p.NT;
  p.tmem_cols = 32;
  while (p.tmem_cols < p.NT) p.tmem_cols <<= 1;
  const int a_bytes = 2 * 2 * p.rows * 128;
  const int b_stage = 2 * p.NT * 128;
  const int budget = 227 * 1024 - 1024 /*align slack*/ - a_bytes - 256 /*barriers*/;
  p.nb_stages = std::min(6, budget / b_stage);
  KT_REQUIRE(p.nb_stages >= 2, "conv_tc: shared memory budget exceeded (rows=%d NT=%d)", p.rows, p.NT);
  const size_t smem = 1024 + a_bytes + (size_t)p.nb_stages * b_stage + 256;
  static std::atomic<bool> cfg{false};
  if (!cfg.load(std::memory_order_acquire)) {
    KT_CHECK_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxDynSmem));
    cfg.store(true, std::memory_order_release);
  }
  dim3 grid(ceil_div(ph.M, kTcM), p.ntiles, p.batch);
  conv_tc_kernel<<<grid, kTcThreads, smem, st>>>(p);
  KT_CHECK_CUDA(cudaGetLastError());
  return KT_OK;
}

static Side make_side_tc(const float* p, const float* aux, int act, float slope, bool derivative) {
  Side s{p, aux, SIDE_PLAIN, slope};
  if (act == KT_ACT_LRELU) s.mode = derivative ? SIDE_DLRELU : SIDE_LRELU;
  else if (act == KT_ACT_TANH) s.mode = derivative ? SIDE_DTANH : SIDE_PLAIN;
  if (s.mode < SIDE_DLRELU) s.aux = nullptr;
  return s;
}

static int tc_flags() {
  static int flags = -1;
  if (flags < 0) {
    // Measured on B200 (round 1, profiles/r01_notes.md): the SWIZZLE_128B XOR is applied to ABSOLUTE
    // shared-memory address bits, so a row-shifted descriptor needs matrix-base-offset = 0; setting it
    // to (addr >> 7) & 7 double-applies the phase and gives wrong results.  Env override kept for the record.
    const char* e = getenv("KT_TC_BASE_OFFSET");
    flags = (e && e[0] == '1') ? 1 : 0;
  }
  return flags;
}

// nsub > 1 with stride 1 folds into a plain sequence of t*nsub rows with dilation*nsub
static KtConv1dDesc fold_nsub(const KtConv1dDesc* d) {
  KtConv1dDesc f = *d;
  if (d->nsub > 1) {
    f.t_in = d->t_in * d->nsub; f.t_out = d->t_out * d->nsub;
    f.dilation = d->dilation * d->nsub; f.pad_left = d->pad_left * d->nsub; f.nsub = 1;
  }
  return f;
}

int conv1d_fwd_tc(const KtConv1dDesc* d0, const float* x, const void* wimg, const float* bias, const float* resid,
                  float* y, cudaStream_t st) {
  const int nt = tc_plan(d0, 0);
  KT_REQUIRE(nt > 0, "conv1d_fwd_tc: layer not supported by the tcgen05 path");
  const KtConv1dDesc f = fold_nsub(d0);
  TcParams p{};
  p.in = make_side_tc(x, nullptr, f.act_in, f.act_in_slope, false);
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.bias = bias; p.resid = resid; p.mask = Side{nullptr, nullptr, 0, 0.f}; p.out = y;
  p.batch = f.batch; p.t_in = f.t_in; p.t_out = f.t_out; p.c_in = f.c_in; p.c_out = f.c_out;
  p.out_act = f.act_out; p.out_slope = f.act_out_slope; p.NT = nt; p.flags = tc_flags();
  for (const Phase& ph : conv_phases(&f, 0)) {
    p.ph = ph;
    int rc = run_tc(p, st);
    if (rc) return rc;
  }
  return KT_OK;
}

int conv1d_bwd_data_tc(const KtConv1dDesc* d0, const float* dy, const float* y, const void* wimg, const float* x,
                       float* dx, cudaStream_t st) {
  const int nt = tc_plan(d0, 1);
  KT_REQUIRE(nt > 0, "conv1d_bwd_data_tc: layer not supported by the tcgen05 path");
  KT_REQUIRE(d0->act_out == KT_ACT_NONE || y != nullptr, "bwd_data: y required when act_out != NONE");
  KT_REQUIRE(d0->act_in == KT_ACT_NONE || x != nullptr, "bwd_data: x required when act_in != NONE");
  const KtConv1dDesc f = fold_nsub(d0);
  TcParams p{};
  p.in = make_side_tc(dy, y, f.act_out, f.act_out_slope, true);
  p.wimg = reinterpret_cast<const __nv_bfloat16*>(wimg);
  p.bias = nullptr; p.resid = nullptr; p.out = dx;
  p.mask = f.act_in == KT_ACT_LRELU ? Side{x, nullptr, SIDE_DLRELU, f.act_in_slope} : Side{nullptr, nullptr, 0, 0.f};
  p.batch = f.batch; p.t_in = f.t_out; p.t_out = f.t_in; p.c_in = f.c_out; p.c_out = f.c_in;
  p.out_act = KT_ACT_NONE; p.out_slope = 0.f; p.NT = nt; p.flags = tc_flags();
  for (const Phase& ph : conv_phases(&f, 1)) {
    p.ph = ph;
    int rc = run_tc(p, st);
    if (rc) return rc;
  }
  return KT_OK;
}

}  // namespace kt
