// FFN up-projection as a persistent tcgen05 GEMM whose epilogue does the causal depthwise conv (k=3), GEGLU with
// exact-erf GELU and the LayerNorm row statistics — the SIMT work runs in the epilogue warps' otherwise idle issue
// slots while the tensor core computes the next tile (TMEM accumulators are double-buffered).
//
//   u = xn @ W1^T                  [M, 2Fp]  (bf16, kept for the backward pass)      transformer.py:144
//   y[t] = w0 u[t-2] + w1 u[t-1] + w2 u[t]   per channel, zero history at sequence start   transformer.py:122-131
//   h = gelu_erf(y_gate) * y_value  [M, Fp]  (bf16)                                         transformer.py:134-137
//   rowsum[m][n tile] = (sum_c h, sum_c h^2) over the tile's 128 channels -> LN(F) statistics  transformer.py:147
//
// Tiling: 128 x 256 x 64, W1 rows interleaved so that one 256-column tile = [128 value | 128 gate] columns of the same
// 128 channels.  The conv needs rows t-1, t-2, so M tiles overlap by two rows (tile i covers rows 126 i - 2 ..
// 126 i + 125 and emits rows 126 i .. 126 i + 125; +1.6 % MMA work, no inter-CTA exchange).  The epilogue first parks
// the bf16 u tile in shared memory (row pitch 528 B: conflict-free for one-row-per-thread 16-byte accesses), releases
// TMEM, then every thread reads its own row and the two rows above it.
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/omlm_b200.h"

namespace omlm {

constexpr int kFuBM = 128, kFuBN = 256, kFuBK = 64, kFuStages = 3, kFuRowsOut = 126;
constexpr int kFuA = kFuBM * kFuBK * 2, kFuB = kFuBN * kFuBK * 2, kFuStage = kFuA + kFuB;   // 16K + 32K
constexpr int kFuUPitch = 528;                                   // bytes per row of the parked u tile
constexpr int kFuOffU = kFuStages * kFuStage;                    // 147456
constexpr int kFuOffBar = kFuOffU + kFuBM * kFuUPitch;
constexpr int kFuSmem = kFuOffBar + 256 + 1024;
constexpr int kFuThreads = 320;   // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter / per SM sub-partition)

// F16: xn, W1 arrive as fp16 and u, h leave as fp16 (all bounded by construction: LayerNorm output x weights);
// otherwise everything is bf16.
template <bool F16>
__global__ void __launch_bounds__(kFuThreads, 1)
gemm_ffn_up_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   __nv_bfloat16* __restrict__ u_out, __nv_bfloat16* __restrict__ h_out, float* __restrict__ rowsum,
                   const float* __restrict__ conv_w, int M, int Nseq, int K, int Fp) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kFuOffBar);
  uint64_t* empty_bar = full_bar + kFuStages;
  uint64_t* tfull_bar = empty_bar + kFuStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* usm = smem + kFuOffU;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + kFuRowsOut - 1) / kFuRowsOut;
  const int n_tiles = (2 * Fp) / kFuBN;
  const int kb_total = (K + kFuBK - 1) / kFuBK;
  const int work_total = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kFuStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // private set-up done: from here on global memory written by the previous kernel is touched

  if (warp == 0) {
    if (lane == 0) {   // ---------------------------------------------------------------- TMA producer
      int stage = 0; uint32_t phase = 0;
      for (int w = blockIdx.x; w < work_total; w += gridDim.x) {
        const int n_blk = w % n_tiles, m_blk = w / n_tiles;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kFuStage;
          mbar_expect_tx(&full_bar[stage], kFuStage);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * kFuBK, m_blk * kFuRowsOut - 2);   // rows < 0 / >= M are zero-filled
          tma_load_2d(sa + kFuA, &tmB, &full_bar[stage], kb * kFuBK, n_blk * kFuBN);
          if (++stage == kFuStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ---------------------------------------------------------------- MMA issuer
      constexpr uint32_t idesc = make_idesc_bf16(kFuBM, kFuBN, 0, 0) & (F16 ? ~((7u << 7) | (7u << 10)) : ~0u);
      int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
      for (int w = blockIdx.x; w < work_total; w += gridDim.x) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kFuBN;
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * kFuStage), sb = sa + kFuA;
#pragma unroll
          for (int k = 0; k < kFuBK / 16; ++k)
            umma_bf16(tmem_d, make_smem_desc(sa + k * 32, 16, 1024), make_smem_desc(sb + k * 32, 16, 1024), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == kFuStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull_bar[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // -------------------------------------------------------------------------------------- epilogue warps (256 threads)
    const int quarter = warp & 3;
    const int t = quarter * 32 + lane;           // tile row owned by this thread (= TMEM lane)
    const int et = threadIdx.x - 64;             // 0..255 within the epilogue group
    const int half = et >> 7;                    // warps 2-5: first 64 channels of the tile, warps 6-9: last 64
    int acc = 0; uint32_t acc_phase = 0;
    bool first = true;
    for (int w = blockIdx.x; w < work_total; w += gridDim.x) {
      const int n_blk = w % n_tiles, m_blk = w / n_tiles;
      if (!first) asm volatile("bar.sync 2, 256;" ::: "memory");   // everyone is done reading the previous u tile
      first = false;
      // conv taps of this lane's 4 value + 4 gate channels (phase 2 mapping), fetched under the wait for the MMAs
      float2 wv[3][2], wg[3][2];                    // taps [k][channel pair]
      {
        const float4* wp = reinterpret_cast<const float4*>(conv_w + static_cast<long>(n_blk * kFuBN + lane * 4) * 3);
        const float4* gp = reinterpret_cast<const float4*>(conv_w + static_cast<long>(n_blk * kFuBN + 128 + lane * 4) * 3);
        const float4 a0 = __ldg(wp), a1 = __ldg(wp + 1), a2 = __ldg(wp + 2);
        const float4 b0 = __ldg(gp), b1 = __ldg(gp + 1), b2 = __ldg(gp + 2);
        const float fa[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
        const float fb[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            wv[k][q] = make_float2(fa[(2 * q) * 3 + k], fa[(2 * q + 1) * 3 + k]);
            wg[k][q] = make_float2(fb[(2 * q) * 3 + k], fb[(2 * q + 1) * 3 + k]);
          }
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      // ---- phase 1: accumulators -> bf16 -> parked u tile (this warp: its 32 rows x 128 of the 256 columns)
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kFuBN + half * 128;
      uint8_t* urow = usm + t * kFuUPitch;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q;
          q.x = pack16x2<F16>(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1]));
          q.y = pack16x2<F16>(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
          q.z = pack16x2<F16>(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
          q.w = pack16x2<F16>(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
          *reinterpret_cast<uint4*>(urow + half * 256 + c * 64 + j * 16) = q;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);   // TMEM buffer free: the next tile's MMAs proceed under phase 2
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      asm volatile("bar.sync 2, 256;" ::: "memory");  // u tile complete
      // ---- phase 2: conv + GEGLU + row statistics.  Warp ew owns tile rows 2+16 ew .. (16 rows), lanes span the 128
      // channels (4 value + 4 gate columns each): every smem read and global store is contiguous across the warp, the
      // two history rows slide through registers, and the row sums are reduced with one transposing butterfly.
      {
        const int ew = et >> 5;
        const int t0 = 2 + ew * 16;
        const long grow0 = static_cast<long>(m_blk) * kFuRowsOut - 2 + t0;
        const int nrows = min(min(16, kFuBM - t0), static_cast<int>(min(static_cast<long>(16), M - grow0)));
        int pos = static_cast<int>(grow0 % Nseq);
        const uint8_t* rp = usm + t0 * kFuUPitch;
        float2 xv1[2], xg1[2], xv2[2], xg2[2];        // fp32x2 pairs: channels (0,1) and (2,3) of this lane
        {
          const uint2 a = *reinterpret_cast<const uint2*>(rp - kFuUPitch + lane * 8);
          const uint2 g = *reinterpret_cast<const uint2*>(rp - kFuUPitch + 256 + lane * 8);
          xv1[0] = unpack16x2<F16>(a.x); xv1[1] = unpack16x2<F16>(a.y); xg1[0] = unpack16x2<F16>(g.x); xg1[1] = unpack16x2<F16>(g.y);
        }
        {
          const uint2 a = *reinterpret_cast<const uint2*>(rp - 2 * kFuUPitch + lane * 8);
          const uint2 g = *reinterpret_cast<const uint2*>(rp - 2 * kFuUPitch + 256 + lane * 8);
          xv2[0] = unpack16x2<F16>(a.x); xv2[1] = unpack16x2<F16>(a.y); xg2[0] = unpack16x2<F16>(g.x); xg2[1] = unpack16x2<F16>(g.y);
        }
        if (pos == 1) { xv2[0] = xv2[1] = xg2[0] = xg2[1] = make_float2(0.f, 0.f); }   // row t-2 belongs to the previous sequence
        float st[32];
        __nv_bfloat16* ug = u_out + grow0 * (2L * Fp) + n_blk * kFuBN + lane * 8;
        __nv_bfloat16* hg = h_out + grow0 * static_cast<long>(Fp) + n_blk * 128 + lane * 4;
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          float s1 = 0.f, s2 = 0.f;
          if (r < nrows) {
            const uint8_t* rr = rp + r * kFuUPitch;
            *reinterpret_cast<uint4*>(ug + static_cast<long>(r) * (2L * Fp)) = *reinterpret_cast<const uint4*>(rr + lane * 16);
            const uint2 a = *reinterpret_cast<const uint2*>(rr + lane * 8);
            const uint2 g = *reinterpret_cast<const uint2*>(rr + 256 + lane * 8);
            const float2 xv0[2] = {unpack16x2<F16>(a.x), unpack16x2<F16>(a.y)}, xg0[2] = {unpack16x2<F16>(g.x), unpack16x2<F16>(g.y)};
            if (pos == 0) {   // sequence start: no history (the zeros then slide into the t-2 slot for the next row)
              const float2 z = make_float2(0.f, 0.f);
              xv1[0] = xv1[1] = xg1[0] = xg1[1] = z;
              xv2[0] = xv2[1] = xg2[0] = xg2[1] = z;
            }
            float2 h[2], sq = make_float2(0.f, 0.f), sm = sq;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const float2 yv = fma2(wv[0][q], xv2[q], fma2(wv[1][q], xv1[q], mul2(wv[2][q], xv0[q])));
              const float2 yg = fma2(wg[0][q], xg2[q], fma2(wg[1][q], xg1[q], mul2(wg[2][q], xg0[q])));
              h[q] = mul2(gelu_erf2(yg), yv);
              sm = add2(sm, h[q]);
              sq = fma2(h[q], h[q], sq);
              xv2[q] = xv1[q]; xv1[q] = xv0[q];
              xg2[q] = xg1[q]; xg1[q] = xg0[q];
            }
            s1 = sm.x + sm.y; s2 = sq.x + sq.y;
            *reinterpret_cast<uint2*>(hg + static_cast<long>(r) * Fp) = make_uint2(pack16x2<F16>(h[0].x, h[0].y), pack16x2<F16>(h[1].x, h[1].y));
            if (++pos == Nseq) pos = 0;
          }
          st[2 * r] = s1;
          st[2 * r + 1] = s2;
        }
        // transposing butterfly: lane l ends up with the warp-wide sum of st[l]
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
          const bool hi = (lane & off) != 0;
#pragma unroll
          for (int i = 0; i < off; ++i) {
            const float send = hi ? st[i] : st[i + off];
            const float keep = hi ? st[i + off] : st[i];
            st[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
          }
        }
        // per-tile partial sums, one slot per (row, n tile): omlm_ffn_norm_fwd adds the n_tiles partials in a fixed
        // order, so the forward pass stays bit-reproducible (fp32 atomics would make the LayerNorm statistics, and
        // after six layers of bf16 rounding every logit, depend on CTA timing)
        if ((lane >> 1) < nrows) rowsum[(grow0 + (lane >> 1)) * (2L * n_tiles) + 2 * n_blk + (lane & 1)] = st[0];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace omlm

extern "C" int omlm_gemm_ffn_up(const void* xn, const void* w1_packed, const float* conv_w_packed, void* u_out, void* h_out,
                                float* rowsum, int M, int Nseq, int K, int Fp, int act_f16, int max_ctas, void* stream) {
  using namespace omlm;
  OMLM_CHECK_ARG(M > 0 && Nseq > 0 && K > 0 && K % 8 == 0 && Fp > 0 && Fp % 128 == 0, "gemm_ffn_up: bad shape M=%d K=%d Fp=%d", M, K, Fp);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, xn, static_cast<uint64_t>(K), static_cast<uint64_t>(M), static_cast<uint64_t>(K) * 2, 64, kFuBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, w1_packed, static_cast<uint64_t>(K), static_cast<uint64_t>(2 * Fp), static_cast<uint64_t>(K) * 2, 64, kFuBN);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    OMLM_CUDA(cudaFuncSetAttribute(gemm_ffn_up_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmem));
    OMLM_CUDA(cudaFuncSetAttribute(gemm_ffn_up_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmem));
    configured = true;
  }
  const int m_tiles = (M + kFuRowsOut - 1) / kFuRowsOut, n_tiles = (2 * Fp) / kFuBN;
  int grid = num_sms();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  if (m_tiles * n_tiles < grid) grid = m_tiles * n_tiles;
  auto kern = act_f16 ? gemm_ffn_up_kernel<true> : gemm_ffn_up_kernel<false>;
  OMLM_KLAUNCH((kern), grid, kFuThreads, kFuSmem, reinterpret_cast<cudaStream_t>(stream), 
      tmA, tmB, reinterpret_cast<__nv_bfloat16*>(u_out), reinterpret_cast<__nv_bfloat16*>(h_out), rowsum, conv_w_packed, M,
      Nseq, K, Fp);
  OMLM_LAUNCH_CHECK();
  return 0;
}
