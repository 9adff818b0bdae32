// FFN up-projection as a persistent tcgen05 GEMM whose epilogue does the causal depthwise conv (k=3), GEGLU with
// exact-erf GELU and the LayerNorm row statistics — the SIMT work runs in the epilogue warps' otherwise idle issue
// slots while the tensor core computes the next tile (TMEM accumulators are double-buffered).
//
//   u = xn @ W1^T                  [M, 2Fp]  (bf16, kept for the backward pass)      transformer.py:144
//   y[t] = w0 u[t-2] + w1 u[t-1] + w2 u[t]   per channel, zero history at sequence start   transformer.py:122-131
//   h = gelu_erf(y_gate) * y_value  [M, Fp]  (bf16)                                         transformer.py:134-137
//   rowsum[m] += (sum_c h, sum_c h^2)   -> LayerNorm(F) statistics for omlm_ffn_norm_fwd    transformer.py:147
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
constexpr int kFuOffW = kFuOffU + kFuBM * kFuUPitch;             // conv taps of this tile's 256 columns (float4 each)
constexpr int kFuOffBar = kFuOffW + kFuBN * 16;
constexpr int kFuSmem = kFuOffBar + 256 + 1024;
constexpr int kFuThreads = 192;

__global__ void __launch_bounds__(kFuThreads, 1)
gemm_ffn_up_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   __nv_bfloat16* __restrict__ u_out, __nv_bfloat16* __restrict__ h_out, float* __restrict__ rowsum,
                   const float* __restrict__ conv_w, int M, int Nseq, int K, int Fp) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kFuOffBar);
  uint64_t* empty_bar = full_bar + kFuStages;
  uint64_t* tfull_bar = empty_bar + kFuStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint8_t* usm = smem + kFuOffU;
  float4* wsm = reinterpret_cast<float4*>(smem + kFuOffW);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + kFuRowsOut - 1) / kFuRowsOut;
  const int n_tiles = (2 * Fp) / kFuBN;
  const int kb_total = (K + kFuBK - 1) / kFuBK;
  const int work_total = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kFuStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

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
      constexpr uint32_t idesc = make_idesc_bf16(kFuBM, kFuBN, 0, 0);
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
    // -------------------------------------------------------------------------------------- epilogue warps (128 threads)
    const int quarter = warp & 3;
    const int t = quarter * 32 + lane;           // tile row owned by this thread (= TMEM lane)
    const int et = threadIdx.x - 64;             // 0..127 within the epilogue group
    int acc = 0; uint32_t acc_phase = 0;
    bool first = true;
    for (int w = blockIdx.x; w < work_total; w += gridDim.x) {
      const int n_blk = w % n_tiles, m_blk = w / n_tiles;
      if (!first) asm volatile("bar.sync 2, 128;" ::: "memory");   // everyone is done reading the previous u tile / taps
      first = false;
      // conv taps of this tile's 256 columns -> smem (3 scalar loads per column, two columns per thread)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int col = et + i * 128;
        const float* wp = conv_w + static_cast<long>(n_blk * kFuBN + col) * 3;
        wsm[col] = make_float4(wp[0], wp[1], wp[2], 0.f);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      // ---- phase 1: accumulators -> bf16 -> parked u tile
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kFuBN;
      uint8_t* urow = usm + t * kFuUPitch;
#pragma unroll 1
      for (int c = 0; c < kFuBN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q;
          q.x = pack_bf16x2(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1]));
          q.y = pack_bf16x2(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
          q.z = pack_bf16x2(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
          q.w = pack_bf16x2(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
          *reinterpret_cast<uint4*>(urow + c * 64 + j * 16) = q;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);   // TMEM buffer free: the next tile's MMAs proceed under phase 2
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      asm volatile("bar.sync 2, 128;" ::: "memory");  // u tile + taps complete
      // ---- phase 2: conv + GEGLU + row statistics on rows t >= 2 of the tile
      const long grow = static_cast<long>(m_blk) * kFuRowsOut - 2 + t;
      if (t >= 2 && grow < M) {
        const int pos = static_cast<int>(grow % Nseq);
        const float k1 = pos >= 1 ? 1.f : 0.f, k2 = pos >= 2 ? 1.f : 0.f;   // no history across sequence starts
        const uint8_t* r0 = urow;
        const uint8_t* r1 = urow - kFuUPitch;
        const uint8_t* r2 = urow - 2 * kFuUPitch;
        __nv_bfloat16* ug = u_out + grow * (2L * Fp) + n_blk * kFuBN;
#pragma unroll 4
        for (int j = 0; j < 32; ++j) reinterpret_cast<uint4*>(ug)[j] = reinterpret_cast<const uint4*>(r0)[j];
        __nv_bfloat16* hg = h_out + grow * static_cast<long>(Fp) + n_blk * 128;
        float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
        for (int j = 0; j < 16; ++j) {   // 8 channels per step
          const uint4 a0 = reinterpret_cast<const uint4*>(r0)[j], a1 = reinterpret_cast<const uint4*>(r1)[j],
                      a2 = reinterpret_cast<const uint4*>(r2)[j];
          const uint4 g0 = reinterpret_cast<const uint4*>(r0 + 256)[j], g1 = reinterpret_cast<const uint4*>(r1 + 256)[j],
                      g2 = reinterpret_cast<const uint4*>(r2 + 256)[j];
          const uint32_t A0[4] = {a0.x, a0.y, a0.z, a0.w}, A1[4] = {a1.x, a1.y, a1.z, a1.w}, A2[4] = {a2.x, a2.y, a2.z, a2.w};
          const uint32_t G0[4] = {g0.x, g0.y, g0.z, g0.w}, G1[4] = {g1.x, g1.y, g1.z, g1.w}, G2[4] = {g2.x, g2.y, g2.z, g2.w};
          uint32_t hp[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 xa0 = unpack_bf16x2(A0[e]), xa1 = unpack_bf16x2(A1[e]), xa2 = unpack_bf16x2(A2[e]);
            const float2 xg0 = unpack_bf16x2(G0[e]), xg1 = unpack_bf16x2(G1[e]), xg2 = unpack_bf16x2(G2[e]);
            const float4 wa0 = wsm[j * 8 + 2 * e], wa1 = wsm[j * 8 + 2 * e + 1];
            const float4 wg0 = wsm[128 + j * 8 + 2 * e], wg1 = wsm[128 + j * 8 + 2 * e + 1];
            const float ya0 = wa0.x * (k2 * xa2.x) + wa0.y * (k1 * xa1.x) + wa0.z * xa0.x;
            const float ya1 = wa1.x * (k2 * xa2.y) + wa1.y * (k1 * xa1.y) + wa1.z * xa0.y;
            const float yg0 = wg0.x * (k2 * xg2.x) + wg0.y * (k1 * xg1.x) + wg0.z * xg0.x;
            const float yg1 = wg1.x * (k2 * xg2.y) + wg1.y * (k1 * xg1.y) + wg1.z * xg0.y;
            const float h0 = gelu_erf(yg0) * ya0, h1 = gelu_erf(yg1) * ya1;
            s1 += h0 + h1;
            s2 += h0 * h0 + h1 * h1;
            hp[e] = pack_bf16x2(h0, h1);
          }
          reinterpret_cast<uint4*>(hg)[j] = make_uint4(hp[0], hp[1], hp[2], hp[3]);
        }
        atomicAdd(rowsum + 2 * grow, s1);
        atomicAdd(rowsum + 2 * grow + 1, s2);
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
                                float* rowsum, int M, int Nseq, int K, int Fp, int max_ctas, void* stream) {
  using namespace omlm;
  OMLM_CHECK_ARG(M > 0 && Nseq > 0 && K > 0 && K % 8 == 0 && Fp > 0 && Fp % 128 == 0, "gemm_ffn_up: bad shape M=%d K=%d Fp=%d", M, K, Fp);
  CUtensorMap tmA, tmB;
  int rc = make_tmap_bf16_2d(&tmA, xn, static_cast<uint64_t>(K), static_cast<uint64_t>(M), static_cast<uint64_t>(K) * 2, 64, kFuBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmB, w1_packed, static_cast<uint64_t>(K), static_cast<uint64_t>(2 * Fp), static_cast<uint64_t>(K) * 2, 64, kFuBN);
  if (rc) return rc;
  static bool configured = false;
  if (!configured) {
    OMLM_CUDA(cudaFuncSetAttribute(gemm_ffn_up_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmem));
    configured = true;
  }
  const int m_tiles = (M + kFuRowsOut - 1) / kFuRowsOut, n_tiles = (2 * Fp) / kFuBN;
  int grid = num_sms();
  if (max_ctas > 0 && max_ctas < grid) grid = max_ctas;
  if (m_tiles * n_tiles < grid) grid = m_tiles * n_tiles;
  gemm_ffn_up_kernel<<<grid, kFuThreads, kFuSmem, reinterpret_cast<cudaStream_t>(stream)>>>(
      tmA, tmB, reinterpret_cast<__nv_bfloat16*>(u_out), reinterpret_cast<__nv_bfloat16*>(h_out), rowsum, conv_w_packed, M,
      Nseq, K, Fp);
  OMLM_LAUNCH_CHECK();
  return 0;
}
