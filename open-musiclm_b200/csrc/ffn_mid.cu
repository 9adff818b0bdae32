// The SIMT parts of ConvFeedForward that are NOT fused into a GEMM epilogue (transformer.py:140-150):
//   forward : LayerNorm(F) + dropout on the h tile that the FFN-up GEMM epilogue produced (gemm_ffn_up.cu does the
//             causal depthwise conv k=3 (122-131) and GEGLU with exact-erf GELU (134-137) under the MMA);
//   backward: dropout/LN backward, GEGLU backward, transposed causal conv and the conv / gamma weight gradients.
//
// Layout: u / du are [M, 2*Fp] bf16 in the INTERLEAVED GEGLU order: channels in groups of 128, each group stored as
// [128 value columns | 128 gate columns] (Fp = F padded to a multiple of 128; padded weights are zero so padded
// channels are exactly 0 everywhere).  h / hn / dhn are [M, Fp] in natural channel order.
#include "common.cuh"
#include "../../include/omlm_b200.h"

namespace omlm {

// Column of channel c's VALUE half in the interleaved u / W1 / conv layout: 128-channel groups stored as
// [128 value columns | 128 gate columns] so that one 256-wide GEMM tile holds both halves of its channels.
__device__ __forceinline__ int ileave(int c) { return ((c >> 7) << 8) + (c & 127); }

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
// forward, second half: the FFN-up GEMM epilogue (gemm_ffn_up.cu) already produced h = gelu(conv(u)_gate) * conv(u)_value
// as bf16 and the per-row sums (sum h, sum h^2) in fp32.  This kernel finishes LayerNorm(F) + dropout:
//   hn = dropout((h - mean) * rstd * gamma),   stats[row] = (mean, rstd)  (kept for the backward pass).
// One warp per row, 16-byte accesses; HBM-bound (2 * M * Fp * 2 bytes).
__device__ __forceinline__ void load4(const __nv_bfloat16* p, bool ok, float (&f)[4]) {
  uint2 raw = make_uint2(0, 0);
  if (ok) raw = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

__global__ void __launch_bounds__(256)
ffn_norm_fwd_kernel(const __nv_bfloat16* __restrict__ h, const float2* __restrict__ rowsum,
                    const float* __restrict__ gamma, __nv_bfloat16* __restrict__ hn, float2* __restrict__ stats,
                    long M, int F, int Fp, float drop_p, const unsigned long long* __restrict__ seed_ptr, uint32_t layer) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * 8 + warp;
  if (row >= M) return;
  const float2 rsum = rowsum[row];
  const float mean = rsum.x / F;
  const float var = fmaxf(rsum.y / F - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  if (lane == 0) stats[row] = make_float2(mean, rstd);
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t thresh = static_cast<uint32_t>(drop_p * 65536.f);
  const unsigned long long seed = (drop_p > 0.f) ? *seed_ptr : 0ull;
  for (int chunk = lane; chunk * 8 < Fp; chunk += 32) {
    float v[8];
    load8(h + row * Fp + chunk * 8, true, v);
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + chunk * 8);
    const float4 g1 = *reinterpret_cast<const float4*>(gamma + chunk * 8 + 4);
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    bool keep[8];
    if (drop_p > 0.f) dropout_keep8(seed, layer, row, chunk, thresh, keep);
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i] = (v[i] - mean) * rstd * gm[i];          // gamma is zero in the padding -> padded channels stay 0
      if (drop_p > 0.f) o[i] = keep[i] ? o[i] * keep_scale : 0.f;
    }
    store8(hn + row * Fp + chunk * 8, o);
  }
}

// ------------------------------------------------------------------------------------------------
// backward.
//
// LayerNorm backward needs two row sums before any element can be finished:
//   m1 = mean_c(gamma * g),  m2 = mean_c(gamma * g * hhat),   g = dropout-backward(dhn).
// Since hn = hhat * gamma * mask/(1-p) was saved by the forward pass, gamma*g*hhat == dhn * hn, so both
// sums come from one cheap pass over (dhn, hn)  [kernel 1, one warp per row].
// With m1/m2 known the rest is element-local in the channel dimension, so kernel 2 lets every thread own
// 4 channels (of both GEGLU halves) and WALK DOWN a slab of time steps with sliding windows: it recomputes
// conv/GEGLU/LN from u, forms dy (never written to memory), applies the transposed causal conv
// du[t] = w2 dy[t] + w1 dy[t+1] + w0 dy[t+2] and accumulates dconv_w / dgamma in registers.
__global__ void __launch_bounds__(256)
ffn_mid_bwd_stats_kernel(const __nv_bfloat16* __restrict__ dhn, const __nv_bfloat16* __restrict__ hn,
                         const float* __restrict__ gamma, float2* __restrict__ rowstat, long M, int F, int Fp,
                         float drop_p, const unsigned long long* __restrict__ seed_ptr, uint32_t layer) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * 8 + warp;
  if (row >= M) return;
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const uint32_t thresh = static_cast<uint32_t>(drop_p * 65536.f);
  const unsigned long long seed = (drop_p > 0.f) ? *seed_ptr : 0ull;
  float s1 = 0.f, s2 = 0.f;
  for (int chunk = lane; chunk * 8 < Fp; chunk += 32) {
    float d[8], hv[8];
    load8(dhn + row * Fp + chunk * 8, true, d);
    load8(hn + row * Fp + chunk * 8, true, hv);
    bool keep[8];
    if (drop_p > 0.f) dropout_keep8(seed, layer, row, chunk, thresh, keep);
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + chunk * 8);
    const float4 g1 = *reinterpret_cast<const float4*>(gamma + chunk * 8 + 4);
    const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float g = (drop_p > 0.f && !keep[i]) ? 0.f : d[i] * keep_scale;
      s1 += gm[i] * g;
      s2 += d[i] * hv[i];
    }
  }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  if (lane == 0) rowstat[row] = make_float2(s1 / F, s2 / F);
}

constexpr int kWalkThreads = 128;
constexpr int kWalkCh = 2;   // channels (of each GEGLU half) per thread: small footprint -> many resident warps

__device__ __forceinline__ void load2(const __nv_bfloat16* p, bool ok, float (&f)[2]) {
  uint32_t raw = 0;
  if (ok) raw = *reinterpret_cast<const uint32_t*>(p);
  const float2 a = unpack_bf16x2(raw);
  f[0] = a.x; f[1] = a.y;
}

__global__ void __launch_bounds__(kWalkThreads, 6)
ffn_mid_bwd_walk_kernel(const MidArgs a, const __nv_bfloat16* __restrict__ dhn, const float2* __restrict__ stats,
                        const float2* __restrict__ rowstat, __nv_bfloat16* __restrict__ du,
                        float* __restrict__ dgamma, float* __restrict__ dconv_w, int rows_per_cta) {
  constexpr int CH = kWalkCh;
  const int slabs = (a.N + rows_per_cta - 1) / rows_per_cta;
  const int b = blockIdx.x / slabs, t0 = (blockIdx.x - b * slabs) * rows_per_cta;
  const int t_end = min(a.N, t0 + rows_per_cta);
  const int c0raw = (blockIdx.y * kWalkThreads + threadIdx.x) * CH;
  const bool active = c0raw < a.Fp;                  // inactive lanes stay alive for the shuffles below
  const int c0 = active ? c0raw : 0;
  const int chunk8 = c0 >> 3, sub = threadIdx.x & 3;  // dropout bits are defined per 8-channel chunk (4 adjacent lanes)
  const int ca = ileave(c0);                          // column of this thread's value channels in u / du / conv_w (gate: +128)
  const long long row_base = static_cast<long long>(b) * a.N;
  const long ld = 2L * a.Fp;
  float wa[CH][3], wg[CH][3], gm[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { wa[i][k] = a.conv_w[(ca + i) * 3 + k]; wg[i][k] = a.conv_w[(ca + 128 + i) * 3 + k]; }
    gm[i] = a.gamma[c0 + i];
  }
  const float keep_scale = a.drop_p > 0.f ? 1.f / (1.f - a.drop_p) : 1.f;
  const uint32_t thresh = static_cast<uint32_t>(a.drop_p * 65536.f);
  const unsigned long long seed = (a.drop_p > 0.f) ? *a.seed : 0ull;
  float ua2[CH], ua1[CH], ug2[CH], ug1[CH];           // u rows t'-2, t'-1
  {
    const bool ok2 = active && t0 - 2 >= 0, ok1 = active && t0 - 1 >= 0;
    const __nv_bfloat16* p2 = a.u + (row_base + t0 - 2) * ld + ca;
    const __nv_bfloat16* p1 = a.u + (row_base + t0 - 1) * ld + ca;
    load2(p2, ok2, ua2); load2(p2 + 128, ok2, ug2);
    load2(p1, ok1, ua1); load2(p1 + 128, ok1, ug1);
  }
  float da2[CH], da1[CH], dg2[CH], dg1[CH];   // dy rows t'-2, t'-1
  float dwa[CH][3], dwg[CH][3], dgam[CH];
#pragma unroll
  for (int i = 0; i < CH; ++i) { da2[i] = da1[i] = dg2[i] = dg1[i] = 0.f; dgam[i] = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) { dwa[i][k] = 0.f; dwg[i][k] = 0.f; } }

  // software pipeline: the (independent) loads of row t'+1 are issued before row t' is processed, so each
  // iteration overlaps one L2/HBM round trip with the arithmetic of the previous row
  struct RowIn { uint32_t ua, ug, d; float2 st, rs; };
  auto fetch = [&](int tp) {
    RowIn r;
    r.ua = 0u; r.ug = 0u; r.d = 0u; r.st = make_float2(0.f, 0.f); r.rs = r.st;
    if (active && tp < a.N) {
      const long long row = row_base + tp;
      r.ua = *reinterpret_cast<const uint32_t*>(a.u + row * ld + ca);
      r.ug = *reinterpret_cast<const uint32_t*>(a.u + row * ld + ca + 128);
      r.d = *reinterpret_cast<const uint32_t*>(dhn + row * a.Fp + c0);
      r.st = stats[row]; r.rs = rowstat[row];
    }
    return r;
  };
  RowIn cur = fetch(t0);
  for (int tp = t0; tp < t_end + 2; ++tp) {
    const bool valid = active && tp < a.N;
    const bool own = tp < t_end;                 // rows >= t_end are the next slab's: recomputed here only for the conv halo
    const long long row = row_base + tp;
    const RowIn nxt = fetch(tp + 1 < t_end + 2 ? tp + 1 : a.N);
    float ua0[CH], ug0[CH], d[CH];
    { float2 x = unpack_bf16x2(cur.ua); ua0[0] = x.x; ua0[1] = x.y;
      x = unpack_bf16x2(cur.ug); ug0[0] = x.x; ug0[1] = x.y;
      x = unpack_bf16x2(cur.d); d[0] = x.x; d[1] = x.y; }
    const float2 st = cur.st, rs = cur.rs;
    if (a.drop_p > 0.f) {
      // one Philox call per 8-channel chunk: lane 0 of each group of 4 computes it, the others borrow their 32 bits
      uint4 rnd = make_uint4(0, 0, 0, 0);
      if (sub == 0) rnd = philox4x32(static_cast<uint32_t>(row), static_cast<uint32_t>(row >> 32), static_cast<uint32_t>(chunk8), a.layer,
                                     static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
      const int src = threadIdx.x & 28;
      const uint32_t x = __shfl_sync(0xffffffffu, rnd.x, src), y = __shfl_sync(0xffffffffu, rnd.y, src);
      const uint32_t z = __shfl_sync(0xffffffffu, rnd.z, src), w = __shfl_sync(0xffffffffu, rnd.w, src);
      const uint32_t bits = sub == 0 ? x : (sub == 1 ? y : (sub == 2 ? z : w));
      d[0] = (bits & 0xffffu) >= thresh ? d[0] * keep_scale : 0.f;
      d[1] = (bits >> 16) >= thresh ? d[1] * keep_scale : 0.f;
    }
    float da0[CH], dg0[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const float ya = wa[i][0] * ua2[i] + wa[i][1] * ua1[i] + wa[i][2] * ua0[i];
      const float yg = wg[i][0] * ug2[i] + wg[i][1] * ug1[i] + wg[i][2] * ug0[i];
      float phi, pdf;
      normal_cdf_pdf(yg, phi, pdf);
      const float ge = yg * phi;
      const float hhat = (ge * ya - st.x) * st.y;
      const bool real = valid && (c0 + i < a.F);
      const float dh = real ? st.y * (gm[i] * d[i] - rs.x - hhat * rs.y) : 0.f;
      da0[i] = dh * ge;
      dg0[i] = dh * ya * fmaf(yg, pdf, phi);
      if (own && real) {
        dgam[i] += d[i] * hhat;
        dwa[i][0] += da0[i] * ua2[i]; dwa[i][1] += da0[i] * ua1[i]; dwa[i][2] += da0[i] * ua0[i];
        dwg[i][0] += dg0[i] * ug2[i]; dwg[i][1] += dg0[i] * ug1[i]; dwg[i][2] += dg0[i] * ug0[i];
      }
    }
    if (active && tp - 2 >= t0) {  // du[t'-2] = w2 dy[t'-2] + w1 dy[t'-1] + w0 dy[t']
      float oa[CH], og[CH];
#pragma unroll
      for (int i = 0; i < CH; ++i) {
        oa[i] = wa[i][2] * da2[i] + wa[i][1] * da1[i] + wa[i][0] * da0[i];
        og[i] = wg[i][2] * dg2[i] + wg[i][1] * dg1[i] + wg[i][0] * dg0[i];
      }
      *reinterpret_cast<uint32_t*>(du + (row - 2) * ld + ca) = pack_bf16x2(oa[0], oa[1]);
      *reinterpret_cast<uint32_t*>(du + (row - 2) * ld + ca + 128) = pack_bf16x2(og[0], og[1]);
    }
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      ua2[i] = ua1[i]; ua1[i] = ua0[i]; ug2[i] = ug1[i]; ug1[i] = ug0[i];
      da2[i] = da1[i]; da1[i] = da0[i]; dg2[i] = dg1[i]; dg1[i] = dg0[i];
    }
    cur = nxt;
  }
  if (!active) return;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    atomicAdd(&dgamma[c0 + i], dgam[i]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      atomicAdd(&dconv_w[(ca + i) * 3 + k], dwa[i][k]);
      atomicAdd(&dconv_w[(ca + 128 + i) * 3 + k], dwg[i][k]);
    }
  }
}


}  // namespace omlm

extern "C" {

int omlm_ffn_norm_fwd(const void* h, const float* rowsum, const float* gamma, void* hn, float* stats, long M, int F,
                      int Fp, float drop_p, const unsigned long long* seed, int layer, void* stream) {
  using namespace omlm;
  OMLM_CHECK_ARG(M > 0 && F > 0 && Fp >= F && Fp % 128 == 0, "ffn_norm_fwd: bad shape F=%d Fp=%d", F, Fp);
  OMLM_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || seed != nullptr), "ffn_norm_fwd: bad dropout args");
  ffn_norm_fwd_kernel<<<static_cast<int>((M + 7) / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(h), reinterpret_cast<const float2*>(rowsum), gamma,
      reinterpret_cast<__nv_bfloat16*>(hn), reinterpret_cast<float2*>(stats), M, F, Fp, drop_p, seed,
      static_cast<uint32_t>(layer));
  OMLM_LAUNCH_CHECK();
  return 0;
}

int omlm_ffn_mid_bwd(const void* dhn, const void* hn, const void* u, const float* stats, const float* conv_w,
                     const float* gamma, float* rowstat_scratch, void* du, float* dgamma, float* dconv_w, int B, int N,
                     int F, int Fp, float drop_p, const unsigned long long* seed, int layer, void* stream) {
  using namespace omlm;
  OMLM_CHECK_ARG(B > 0 && N > 0 && F > 0 && Fp >= F && Fp % 128 == 0, "ffn_mid_bwd: bad shape F=%d Fp=%d", F, Fp);
  auto st = reinterpret_cast<cudaStream_t>(stream);
  MidArgs a{reinterpret_cast<const __nv_bfloat16*>(u), conv_w, gamma, N, F, Fp, drop_p, seed, static_cast<uint32_t>(layer)};
  const long M = static_cast<long>(B) * N;
  ffn_mid_bwd_stats_kernel<<<static_cast<int>((M + 7) / 8), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(dhn), reinterpret_cast<const __nv_bfloat16*>(hn), gamma,
      reinterpret_cast<float2*>(rowstat_scratch), M, F, Fp, drop_p, seed, static_cast<uint32_t>(layer));
  OMLM_LAUNCH_CHECK();
  const int rows_per_cta = 32;
  dim3 grid(B * ((N + rows_per_cta - 1) / rows_per_cta), (Fp / kWalkCh + kWalkThreads - 1) / kWalkThreads);
  ffn_mid_bwd_walk_kernel<<<grid, kWalkThreads, 0, st>>>(a, reinterpret_cast<const __nv_bfloat16*>(dhn),
                                                         reinterpret_cast<const float2*>(stats),
                                                         reinterpret_cast<const float2*>(rowstat_scratch),
                                                         reinterpret_cast<__nv_bfloat16*>(du), dgamma, dconv_w, rows_per_cta);
  OMLM_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
