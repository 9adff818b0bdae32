// The bandwidth-bound middle of ConvFeedForward between the two tensor-core GEMMs
// (transformer.py:140-150):  causal depthwise conv k=3 (122-131) -> GEGLU with exact-erf GELU
// (134-137) -> LayerNorm over the inner dim (147) -> dropout (148), fused in ONE pass over u.
//
// Layout: u is [M, 2*Fp] bf16 with the GEGLU value half in columns [0, Fp) and the gate half in
// [Fp, 2*Fp) (Fp = inner dim F padded to a multiple of 64; padded weights are zero so padded
// channels are exactly 0 everywhere).  A CTA owns a slab of kT consecutive time steps of ONE batch
// element (the conv never crosses batch elements) and every thread owns 8 channels of both halves
// (16-byte accesses); the two conv history rows are re-read (L2 hits) instead of staged.
#include "common.cuh"
#include "../../include/omlm_b200.h"

namespace omlm {

constexpr int kT = 8;   // time steps per slab (forward)
constexpr int kTB = 4;  // time steps per slab (backward: three live values per element)
constexpr int kMidMaxThreads = 384;

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

__device__ __forceinline__ void load8(const __nv_bfloat16* p, bool ok, float (&f)[8]) {
  uint4 raw = make_uint4(0, 0, 0, 0);
  if (ok) raw = *reinterpret_cast<const uint4*>(p);
  float2 t;
  t = unpack_bf16x2(raw.x); f[0] = t.x; f[1] = t.y;
  t = unpack_bf16x2(raw.y); f[2] = t.x; f[3] = t.y;
  t = unpack_bf16x2(raw.z); f[4] = t.x; f[5] = t.y;
  t = unpack_bf16x2(raw.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 o;
  o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
  o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = o;
}

// Sum T per-row partials over the whole block; result broadcast to every thread.
template <int T>
__device__ __forceinline__ void block_sum_rows(float (&v)[T], float* red /*[T][32]*/) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
#pragma unroll
  for (int r = 0; r < T; ++r) v[r] = warp_sum(v[r]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < T; ++r) red[r * 32 + warp] = v[r];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < T; ++r) {
    float s = lane < nw ? red[r * 32 + lane] : 0.f;
    v[r] = warp_sum(s);
  }
}

// keep flags for 8 channels of (row, chunk): 16 random bits per channel.
__device__ __forceinline__ void dropout_keep8(unsigned long long seed, uint32_t layer, long long row, int chunk,
                                              uint32_t thresh16, bool (&keep)[8]) {
  const uint4 r = philox4x32(static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32), static_cast<uint32_t>(chunk), layer,
                             static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
  keep[0] = (r.x & 0xffffu) >= thresh16; keep[1] = (r.x >> 16) >= thresh16;
  keep[2] = (r.y & 0xffffu) >= thresh16; keep[3] = (r.y >> 16) >= thresh16;
  keep[4] = (r.z & 0xffffu) >= thresh16; keep[5] = (r.z >> 16) >= thresh16;
  keep[6] = (r.w & 0xffffu) >= thresh16; keep[7] = (r.w >> 16) >= thresh16;
}

struct MidArgs {
  const __nv_bfloat16* u;     // [M, 2Fp]
  const float* conv_w;        // [2Fp, 3] packed
  const float* gamma;         // [Fp] packed (zeros in the padding)
  int N, F, Fp;
  float drop_p;               // 0 -> no dropout
  const unsigned long long* seed;
  uint32_t layer;
};

// ------------------------------------------------------------------------------------------------
// forward: hn = dropout(LN(gelu(conv(u)_gate) * conv(u)_value)),  stats = (mean, rstd) per row
__global__ void __launch_bounds__(kMidMaxThreads, 1)
ffn_mid_fwd_kernel(const MidArgs a, __nv_bfloat16* __restrict__ hn, float2* __restrict__ stats) {
  __shared__ float red[kT * 32];
  const int slabs = (a.N + kT - 1) / kT;
  const int b = blockIdx.x / slabs, t0 = (blockIdx.x - b * slabs) * kT;
  const int chunk = threadIdx.x, c0 = chunk * 8;
  const bool live = c0 < a.Fp;
  const long long row_base = static_cast<long long>(b) * a.N;
  const long ld = 2L * a.Fp;
  float wa[8][3], wg[8][3];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      wa[i][k] = live ? a.conv_w[(c0 + i) * 3 + k] : 0.f;
      wg[i][k] = live ? a.conv_w[(a.Fp + c0 + i) * 3 + k] : 0.f;
    }
  float a2[8], a1[8], g2[8], g1[8];
  {
    const bool ok2 = live && t0 - 2 >= 0, ok1 = live && t0 - 1 >= 0;
    const __nv_bfloat16* p2 = a.u + (row_base + t0 - 2) * ld + c0;
    const __nv_bfloat16* p1 = a.u + (row_base + t0 - 1) * ld + c0;
    load8(p2, ok2, a2); load8(p2 + a.Fp, ok2, g2);
    load8(p1, ok1, a1); load8(p1 + a.Fp, ok1, g1);
  }
  float hv[kT][8];
  float s[kT];
#pragma unroll
  for (int r = 0; r < kT; ++r) {
    const bool ok = live && (t0 + r) < a.N;
    float ac[8], gc[8];
    const __nv_bfloat16* p = a.u + (row_base + t0 + r) * ld + c0;
    load8(p, ok, ac); load8(p + a.Fp, ok, gc);
    s[r] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float ya = wa[i][0] * a2[i] + wa[i][1] * a1[i] + wa[i][2] * ac[i];
      const float yg = wg[i][0] * g2[i] + wg[i][1] * g1[i] + wg[i][2] * gc[i];
      hv[r][i] = gelu_erf(yg) * ya;
      s[r] += hv[r][i];
      a2[i] = a1[i]; a1[i] = ac[i]; g2[i] = g1[i]; g1[i] = gc[i];
    }
  }
  block_sum_rows(s, red);
  float mean[kT], q[kT];
#pragma unroll
  for (int r = 0; r < kT; ++r) {
    mean[r] = s[r] / a.F;
    q[r] = 0.f;
    if (live) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = hv[r][i] - mean[r]; q[r] += d * d; }
    }
  }
  block_sum_rows(q, red);
  float gm[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) gm[i] = live ? a.gamma[c0 + i] : 0.f;
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  const uint32_t thresh = static_cast<uint32_t>(a.drop_p * 65536.f);
  const unsigned long long seed = (a.drop_p > 0.f) ? *a.seed : 0ull;
#pragma unroll
  for (int r = 0; r < kT; ++r) {
    if (t0 + r >= a.N) break;
    // padded channels are 0, each contributed (0-mean)^2 to q: remove them
    const float var = (q[r] - (a.Fp - a.F) * mean[r] * mean[r]) / a.F;
    const float rstd = rsqrtf(var + 1e-5f);
    const long long row = row_base + t0 + r;
    if (threadIdx.x == 0) stats[row] = make_float2(mean[r], rstd);
    if (live) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (hv[r][i] - mean[r]) * rstd * gm[i];
      if (a.drop_p > 0.f) {
        bool keep[8];
        dropout_keep8(seed, a.layer, row, chunk, thresh, keep);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = keep[i] ? o[i] * keep_scale : 0.f;
      }
      store8(hn + row * a.Fp + c0, o);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, step 1 (row-local): dhn -> (dropout, LN backward) -> dh -> GEGLU backward -> dy [M, 2Fp];
// dgamma[c] += sum_rows g * hhat.
__global__ void __launch_bounds__(kMidMaxThreads, 1)
ffn_mid_bwd_rows_kernel(const MidArgs a, const __nv_bfloat16* __restrict__ dhn,
                        const float2* __restrict__ stats, __nv_bfloat16* __restrict__ dy,
                        float* __restrict__ dgamma) {
  __shared__ float red[kTB * 32];
  const int slabs = (a.N + kTB - 1) / kTB;
  const int b = blockIdx.x / slabs, t0 = (blockIdx.x - b * slabs) * kTB;
  const int chunk = threadIdx.x, c0 = chunk * 8;
  const bool live = c0 < a.Fp;
  const long long row_base = static_cast<long long>(b) * a.N;
  const long ld = 2L * a.Fp;
  float wa[8][3], wg[8][3], gm[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      wa[i][k] = live ? a.conv_w[(c0 + i) * 3 + k] : 0.f;
      wg[i][k] = live ? a.conv_w[(a.Fp + c0 + i) * 3 + k] : 0.f;
    }
    gm[i] = live ? a.gamma[c0 + i] : 0.f;
  }
  float a2[8], a1[8], g2[8], g1[8];
  {
    const bool ok2 = live && t0 - 2 >= 0, ok1 = live && t0 - 1 >= 0;
    const __nv_bfloat16* p2 = a.u + (row_base + t0 - 2) * ld + c0;
    const __nv_bfloat16* p1 = a.u + (row_base + t0 - 1) * ld + c0;
    load8(p2, ok2, a2); load8(p2 + a.Fp, ok2, g2);
    load8(p1, ok1, a1); load8(p1 + a.Fp, ok1, g1);
  }
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  const uint32_t thresh = static_cast<uint32_t>(a.drop_p * 65536.f);
  const unsigned long long seed = (a.drop_p > 0.f) ? *a.seed : 0ull;
  // per row we need ya, yg (recomputed), hhat, gg = gamma * g.  Keep ya/yg as bf16-free floats: 2*8*kTB regs.
  float ya[kTB][8], yg[kTB][8], gg[kTB][8];
  float s1[kTB], s2[kTB], dgacc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) dgacc[i] = 0.f;
#pragma unroll
  for (int r = 0; r < kTB; ++r) {
    const bool ok = live && (t0 + r) < a.N;
    const long long row = row_base + t0 + r;
    float ac[8], gc[8], gin[8];
    const __nv_bfloat16* p = a.u + row * ld + c0;
    load8(p, ok, ac); load8(p + a.Fp, ok, gc);
    load8(dhn + row * a.Fp + c0, ok, gin);
    float2 st = make_float2(0.f, 0.f);
    if ((t0 + r) < a.N) st = stats[row];
    if (a.drop_p > 0.f && ok) {
      bool keep[8];
      dropout_keep8(seed, a.layer, row, chunk, thresh, keep);
#pragma unroll
      for (int i = 0; i < 8; ++i) gin[i] = keep[i] ? gin[i] * keep_scale : 0.f;
    }
    s1[r] = 0.f; s2[r] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      ya[r][i] = wa[i][0] * a2[i] + wa[i][1] * a1[i] + wa[i][2] * ac[i];
      yg[r][i] = wg[i][0] * g2[i] + wg[i][1] * g1[i] + wg[i][2] * gc[i];
      const float hval = gelu_erf(yg[r][i]) * ya[r][i];
      const float hhat = (c0 + i < a.F) ? (hval - st.x) * st.y : 0.f;
      dgacc[i] += gin[i] * hhat;
      gg[r][i] = gin[i] * gm[i];
      s1[r] += gg[r][i];
      s2[r] += gg[r][i] * hhat;
      a2[i] = a1[i]; a1[i] = ac[i]; g2[i] = g1[i]; g1[i] = gc[i];
    }
  }
  block_sum_rows(s1, red);
  block_sum_rows(s2, red);
#pragma unroll
  for (int r = 0; r < kTB; ++r) {
    if (t0 + r >= a.N) break;
    if (!live) continue;
    const long long row = row_base + t0 + r;
    const float2 st = stats[row];
    const float m1 = s1[r] / a.F, m2 = s2[r] / a.F;
    float da[8], dg[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float ge = gelu_erf(yg[r][i]);
      const float hval = ge * ya[r][i];
      const float hhat = (hval - st.x) * st.y;
      const float dh = (c0 + i < a.F) ? st.y * (gg[r][i] - m1 - hhat * m2) : 0.f;
      da[i] = dh * ge;
      dg[i] = dh * ya[r][i] * gelu_erf_grad(yg[r][i]);
    }
    store8(dy + row * ld + c0, da);
    store8(dy + row * ld + a.Fp + c0, dg);
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&dgamma[c0 + i], dgacc[i]);
  }
}

// ------------------------------------------------------------------------------------------------
// backward, step 2 (conv transpose): du[t,c] = sum_k w[c,k] * dy[t+2-k, c];  dw[c,k] += sum_t dy[t,c] u[t-2+k,c]
// One CTA = rows_per_cta consecutive time steps of one batch element; thread = 8 channels of [0, 2Fp).
__global__ void __launch_bounds__(128)
conv_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ u,
                const float* __restrict__ conv_w, __nv_bfloat16* __restrict__ du, float* __restrict__ dconv_w,
                int N, int C /*2Fp*/, int rows_per_cta) {
  const int slabs = (N + rows_per_cta - 1) / rows_per_cta;
  const int b = blockIdx.x / slabs, t0 = (blockIdx.x - b * slabs) * rows_per_cta;
  const int t1 = min(N, t0 + rows_per_cta);
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 8;
  if (c0 >= C) return;
  const long long row_base = static_cast<long long>(b) * N;
  float w[8][3], dw[8][3];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) { w[i][k] = conv_w[(c0 + i) * 3 + k]; dw[i][k] = 0.f; }
  // sliding windows: dy rows t, t+1, t+2 ; u rows t-2, t-1, t
  float d0[8], d1[8], d2[8], u2[8], u1[8], u0[8];
  load8(dy + (row_base + t0) * C + c0, t0 < N, d0);
  load8(dy + (row_base + t0 + 1) * C + c0, t0 + 1 < N, d1);
  load8(u + (row_base + t0 - 2) * C + c0, t0 - 2 >= 0, u2);
  load8(u + (row_base + t0 - 1) * C + c0, t0 - 1 >= 0, u1);
  for (int t = t0; t < t1; ++t) {
    load8(dy + (row_base + t + 2) * C + c0, t + 2 < N, d2);
    load8(u + (row_base + t) * C + c0, true, u0);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      // y[t] = w0 u[t-2] + w1 u[t-1] + w2 u[t]  =>  du[t] = w2 dy[t] + w1 dy[t+1] + w0 dy[t+2]
      o[i] = w[i][2] * d0[i] + w[i][1] * d1[i] + w[i][0] * d2[i];
      dw[i][0] += d0[i] * u2[i];
      dw[i][1] += d0[i] * u1[i];
      dw[i][2] += d0[i] * u0[i];
      d0[i] = d1[i]; d1[i] = d2[i]; u2[i] = u1[i]; u1[i] = u0[i];
    }
    store8(du + (row_base + t) * C + c0, o);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 3; ++k) atomicAdd(&dconv_w[(c0 + i) * 3 + k], dw[i][k]);
}

static int mid_threads(int Fp) { return ((Fp / 8 + 31) / 32) * 32; }

}  // namespace omlm

extern "C" {

int omlm_ffn_mid_fwd(const void* u, const float* conv_w, const float* gamma, void* hn, float* stats, int B,
                     int N, int F, int Fp, float drop_p, const unsigned long long* seed, int layer,
                     void* stream) {
  using namespace omlm;
  OMLM_CHECK_ARG(B > 0 && N > 0 && F > 0 && Fp >= F && Fp % 8 == 0 && Fp <= 8 * kMidMaxThreads, "ffn_mid_fwd: bad shape F=%d Fp=%d", F, Fp);
  OMLM_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed != nullptr), "ffn_mid_fwd: bad dropout args");
  MidArgs a{reinterpret_cast<const __nv_bfloat16*>(u), conv_w, gamma, N, F, Fp, drop_p, seed, static_cast<uint32_t>(layer)};
  const int slabs = (N + kT - 1) / kT;
  ffn_mid_fwd_kernel<<<B * slabs, mid_threads(Fp), 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      a, reinterpret_cast<__nv_bfloat16*>(hn), reinterpret_cast<float2*>(stats));
  OMLM_LAUNCH_CHECK();
  return 0;
}

int omlm_ffn_mid_bwd(const void* dhn, const void* u, const float* stats, const float* conv_w, const float* gamma,
                     void* dy_scratch, void* du, float* dgamma, float* dconv_w, int B, int N, int F, int Fp,
                     float drop_p, const unsigned long long* seed, int layer, void* stream) {
  using namespace omlm;
  OMLM_CHECK_ARG(B > 0 && N > 0 && F > 0 && Fp >= F && Fp % 8 == 0 && Fp <= 8 * kMidMaxThreads, "ffn_mid_bwd: bad shape F=%d Fp=%d", F, Fp);
  auto st = reinterpret_cast<cudaStream_t>(stream);
  MidArgs a{reinterpret_cast<const __nv_bfloat16*>(u), conv_w, gamma, N, F, Fp, drop_p, seed, static_cast<uint32_t>(layer)};
  const int slabs = (N + kTB - 1) / kTB;
  ffn_mid_bwd_rows_kernel<<<B * slabs, mid_threads(Fp), 0, st>>>(
      a, reinterpret_cast<const __nv_bfloat16*>(dhn), reinterpret_cast<const float2*>(stats),
      reinterpret_cast<__nv_bfloat16*>(dy_scratch), dgamma);
  OMLM_LAUNCH_CHECK();
  const int C = 2 * Fp;
  const int rows_per_cta = 32;
  const int threads = 128;
  dim3 grid(B * ((N + rows_per_cta - 1) / rows_per_cta), (C / 8 + threads - 1) / threads);
  conv_bwd_kernel<<<grid, threads, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(dy_scratch),
                                            reinterpret_cast<const __nv_bfloat16*>(u), conv_w,
                                            reinterpret_cast<__nv_bfloat16*>(du), dconv_w, N, C, rows_per_cta);
  OMLM_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
