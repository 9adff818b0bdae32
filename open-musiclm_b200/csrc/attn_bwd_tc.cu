// Fused causal multi-query cosine-sim attention, backward, on tcgen05 / TMEM / TMA.
//
// Autograd of transformer.py:304-331 (see attn_bwd.cu for the math).  Folded-row layout: R = N*h query rows
// per batch element share one K/V head.  Work unit = (batch, 128-key tile, chunk of 128-row query tiles).
//
//   warp 0       TMA producer: K/V tile once, Q/dO tiles through a 2-stage ring
//   warp 1       tcgen05.mma issuer, five GEMMs per row tile, all M = 128:
//                   S  = Q K^T          (TMEM cols   0..127)      dP = dO V^T      (128..255)
//                   dV += P^T dO        (384..447, accumulated over the row tiles, A = P read MN-major)
//                   dK += dS^T Q        (320..383, accumulated,                    A = dS read MN-major)
//                   dQ  = dS K          (256..319, per tile,                       A = dS read K-major)
//   warps 2,3    bias-slice (Toeplitz window, causal -inf folded in) + key-mask builders, one tile ahead
//   warps 4-7    one thread per query row: S/dP rows from TMEM, P = exp2(s - lse), dS = P (dP - D), both written
//                as bf16 into ONE 128B-swizzled [row][key] smem tile each that serves as K-major and MN-major operand;
//                the dS tile is also TMA-stored to a global scratch [B, R, Ns] for the bias-gradient pass
//   warps 8-11   drain dQ (TMEM -> red.global.add.v4.f32), and dK/dV at the end
//
// The bias gradient dTable[hh, i-j] = sum dS is a diagonal sum; it is done by a second, bandwidth-bound kernel
// over the dS scratch (one thread per (head, delta), coalesced along delta, no atomics inside the sums).
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/omlm_b200.h"

namespace omlm {

constexpr int kBtThreads = 384;
constexpr int kBtBQ = 128, kBtBK = 128;
constexpr float kBtL2e = 1.4426950408889634f;

constexpr int kBoK = 0, kBoV = 16384, kBoQ = 32768 /*2 stages x 16K*/, kBoDO = 65536 /*2 x 16K*/;
constexpr int kBoP = 98304 /*32K*/, kBoDS = 131072 /*32K*/, kBoKneg = 163840 /*512 B*/, kBoBar = 164352 /*256 B*/;
constexpr int kBoBias = 164608;   // 2 buffers x h*W floats

__device__ __forceinline__ void bt_tmem_ld32(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
        "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]),
        "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
        "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr));
}
__device__ __forceinline__ float bt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bt_red4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kBtThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                   const __grid_constant__ CUtensorMap tmKV, const __grid_constant__ CUtensorMap tmDS,
                   const float* __restrict__ lse2, const float* __restrict__ dsum, const float* __restrict__ table,
                   int table_ld, const unsigned char* __restrict__ key_mask, float* __restrict__ dqn,
                   float* __restrict__ dkvn, int N, int h, float scale, int W, int Wd, int tiles_per_chunk,
                   int units_per_batch) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBoBar);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;    // [2]
  uint64_t* qdo_empty = bars + 3;   // [2]
  uint64_t* sd_full = bars + 5;     // S, dP complete in TMEM
  uint64_t* sd_free = bars + 6;     // S, dP copied to registers (4 warps)
  uint64_t* pds_full = bars + 7;    // P, dS tiles in smem (4 warps)
  uint64_t* pds_empty = bars + 8;   // dV/dK/dQ MMAs finished reading P, dS
  uint64_t* dq_full = bars + 9;     // dQ tile complete in TMEM
  uint64_t* dq_free = bars + 10;    // dQ tile drained (4 warps)
  uint64_t* b_full = bars + 11;     // [2]
  uint64_t* b_empty = bars + 13;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* kneg = reinterpret_cast<float*>(smem + kBoKneg);
  float* bias = reinterpret_cast<float*>(smem + kBoBias);
  const int slice = h * W;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int R = N * h;
  const int n_row_tiles = (R + kBtBQ - 1) / kBtBQ;
  // ---- work unit: (b, key tile kt, chunk of row tiles)
  const int b = blockIdx.x / units_per_batch;
  int u = blockIdx.x - b * units_per_batch;
  int kt = 0, rt_begin = 0, rt_end = 0;
  for (;; ++kt) {
    const int first = (kt * kBtBK * h) / kBtBQ;
    const int chunks = (n_row_tiles - first + tiles_per_chunk - 1) / tiles_per_chunk;
    if (u < chunks) { rt_begin = first + u * tiles_per_chunk; rt_end = min(n_row_tiles, rt_begin + tiles_per_chunk); break; }
    u -= chunks;
  }
  const int j0 = kt * kBtBK;
  const int T = rt_end - rt_begin;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmDO); tma_prefetch_desc(&tmKV); tma_prefetch_desc(&tmDS);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1);
      mbar_init(&b_full[i], 2); mbar_init(&b_empty[i], 4);
    }
    mbar_init(sd_full, 1); mbar_init(sd_free, 4); mbar_init(pds_full, 4); mbar_init(pds_empty, 1);
    mbar_init(dq_full, 1); mbar_init(dq_free, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer
      if (lane == 0) {
        mbar_expect_tx(kv_full, 2 * 16384);
        tma_load_2d(smem + kBoK, &tmKV, kv_full, 0, b * N + j0);
        tma_load_2d(smem + kBoV, &tmKV, kv_full, 64, b * N + j0);
        for (int t = 0; t < T; ++t) {
          const int st = t & 1;
          mbar_wait(&qdo_empty[st], ((t >> 1) & 1) ^ 1);
          mbar_expect_tx(&qdo_full[st], 2 * 16384);
          tma_load_2d(smem + kBoQ + st * 16384, &tmQ, &qdo_full[st], 0, b * R + (rt_begin + t) * kBtBQ);
          tma_load_2d(smem + kBoDO + st * 16384, &tmDO, &qdo_full[st], 0, b * R + (rt_begin + t) * kBtBQ);
        }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer
      if (lane == 0) {
        constexpr uint32_t id_s = make_idesc_bf16(128, 128, 0, 0);   // S, dP: A and B K-major
        constexpr uint32_t id_kv = make_idesc_bf16(128, 64, 1, 1);   // dV, dK: A (P / dS) and B (dO / Q) MN-major
        constexpr uint32_t id_q = make_idesc_bf16(128, 64, 0, 1);    // dQ: A (dS) K-major, B (K) MN-major
        const uint32_t sk = smem_u32(smem + kBoK), sv = smem_u32(smem + kBoV), sq = smem_u32(smem + kBoQ),
                       sdo = smem_u32(smem + kBoDO), sp = smem_u32(smem + kBoP), sds = smem_u32(smem + kBoDS);
        auto issue_s_dp = [&](int st) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem_base + 0, make_smem_desc(sq + st * 16384 + ks * 32, 16, 1024),
                      make_smem_desc(sk + ks * 32, 16, 1024), id_s, ks > 0 ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            umma_bf16(tmem_base + 128, make_smem_desc(sdo + st * 16384 + ks * 32, 16, 1024),
                      make_smem_desc(sv + ks * 32, 16, 1024), id_s, ks > 0 ? 1u : 0u);
        };
        mbar_wait(kv_full, 0);
        mbar_wait(&qdo_full[0], 0);
        tc_fence_after();
        issue_s_dp(0);
        umma_commit(sd_full);
        for (int t = 0; t < T; ++t) {
          const int st = t & 1;
          if (t + 1 < T) {   // next tile's S / dP as soon as this tile's are in registers
            mbar_wait(sd_free, t & 1);
            mbar_wait(&qdo_full[(t + 1) & 1], ((t + 1) >> 1) & 1);
            tc_fence_after();
            issue_s_dp((t + 1) & 1);
            umma_commit(sd_full);
          }
          mbar_wait(pds_full, t & 1);
          if (t > 0) mbar_wait(dq_free, (t - 1) & 1);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)   // dV[key, d] += P^T dO   (k = 16 query rows per step)
            umma_bf16(tmem_base + 384, make_smem_desc(sp + ks * 2048, 16384, 1024),
                      make_smem_desc(sdo + st * 16384 + ks * 2048, 16384, 1024), id_kv, (t > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)   // dK[key, d] += dS^T Q
            umma_bf16(tmem_base + 320, make_smem_desc(sds + ks * 2048, 16384, 1024),
                      make_smem_desc(sq + st * 16384 + ks * 2048, 16384, 1024), id_kv, (t > 0 || ks > 0) ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)   // dQ[row, d] = dS K   (k = 16 keys per step)
            umma_bf16(tmem_base + 256, make_smem_desc(sds + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                      make_smem_desc(sk + ks * 2048, 16384, 1024), id_q, ks > 0 ? 1u : 0u);
          umma_commit(dq_full);
          umma_commit(&qdo_empty[st]);
          umma_commit(pds_empty);
        }
      }
    } else {
      // ------------------------------------------------------------------ bias-slice / key-mask builders (64 threads)
      const int tid = threadIdx.x - 64;
      for (int c = tid; c < kBtBK; c += 64) {
        const int j = j0 + c;
        const bool vis = (j < N) && (key_mask == nullptr || key_mask[static_cast<long long>(b) * N + j] != 0);
        kneg[c] = vis ? 0.f : -INFINITY;
      }
      for (int t = 0; t < T; ++t) {
        const int buf = t & 1;
        mbar_wait(&b_empty[buf], ((t >> 1) & 1) ^ 1);
        const int i_min = ((rt_begin + t) * kBtBQ) / h;
        const int delta_min = i_min - j0 - (kBtBK - 1);
        float* dst = bias + buf * slice;
        for (int hh0 = 0; hh0 < h; hh0 += 8) {
          float v[8][4];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float* trow = table + min(hh0 + k, h - 1) * table_ld;
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              const int w = tid + uu * 64;
              const int delta = delta_min + w;
              v[k][uu] = (w < Wd && delta >= 0) ? __ldg(trow + min(delta, N - 1)) : -INFINITY;
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (hh0 + k < h) {
#pragma unroll
              for (int uu = 0; uu < 4; ++uu) {
                const int w = tid + uu * 64;
                if (w < Wd) dst[(hh0 + k) * W + w] = v[k][uu] * kBtL2e;
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&b_full[buf]);
      }
    }
  } else if (warp < 8) {
    // -------------------------------------------------------------------- P / dS warpgroup: one thread per query row
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int quarter = warp & 3;
    const int row_local = quarter * 32 + lane;
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    uint8_t* prow = smem + kBoP + row_local * 128;
    uint8_t* dsrow = smem + kBoDS + row_local * 128;
    const int sw = row_local & 7;
    const float sc2 = scale * kBtL2e;
    for (int t = 0; t < T; ++t) {
      const int buf = t & 1;
      const int r0 = (rt_begin + t) * kBtBQ;
      const int r = r0 + row_local;
      const int rc = min(r, R - 1);
      const int i = rc / h, hh = rc - i * h;
      const int i_min = r0 / h;
      const float l2 = (r < R) ? lse2[static_cast<long long>(b) * R + r] : INFINITY;
      const float dsm = (r < R) ? dsum[static_cast<long long>(b) * R + r] : 0.f;
      mbar_wait(&b_full[buf], (t >> 1) & 1);
      mbar_wait(sd_full, t & 1);
      tc_fence_after();
      if (t > 0) {
        // the previous tile's MMAs and the TMA store of dS must be done with the P / dS buffers
        mbar_wait(pds_empty, (t - 1) & 1);
        if (threadIdx.x == 128) tma_store_wait_read<0>();
        asm volatile("bar.sync 1, 128;" ::: "memory");
      }
      const float* bp = bias + buf * slice + hh * W + (i - i_min) + (kBtBK - 1);
#pragma unroll 1
      for (int c4 = 0; c4 < 4; ++c4) {   // 32 keys per step
        float s[32], dp[32];
        bt_tmem_ld32(t_s + c4 * 32, s);
        bt_tmem_ld32(t_s + 128 + c4 * 32, dp);
        tmem_ld_wait();
        if (c4 == 3) {   // S and dP fully in registers: the tensor core may start the next tile's S / dP
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(sd_free);
        }
        uint32_t pp[16], dd[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          const int c = c4 * 32 + e;
          const float x0 = fmaf(s[e], sc2, bp[-c] + kneg[c]) - l2;
          const float x1 = fmaf(s[e + 1], sc2, bp[-c - 1] + kneg[c + 1]) - l2;
          const float p0 = bt_ex2(x0), p1 = bt_ex2(x1);
          pp[e >> 1] = pack_bf16x2(p0, p1);
          dd[e >> 1] = pack_bf16x2(p0 * (dp[e] - dsm), p1 * (dp[e + 1] - dsm));
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {   // four 16-byte chunks of 8 keys
          const int ch = c4 * 4 + q4;      // chunk 0..15 of the 128-key row
          const int off = (ch >> 3) * 16384 + (((ch & 7) ^ sw) << 4);
          *reinterpret_cast<uint4*>(prow + off) = make_uint4(pp[q4 * 4], pp[q4 * 4 + 1], pp[q4 * 4 + 2], pp[q4 * 4 + 3]);
          *reinterpret_cast<uint4*>(dsrow + off) = make_uint4(dd[q4 * 4], dd[q4 * 4 + 1], dd[q4 * 4 + 2], dd[q4 * 4 + 3]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_empty[buf]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
      // dS tile -> global scratch for the bias-gradient pass (needs the whole tile: all four warps)
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 128) {
        tma_store_3d(&tmDS, smem + kBoDS, j0, r0, b);
        tma_store_3d(&tmDS, smem + kBoDS + 16384, j0 + 64, r0, b);
        tma_store_commit();
      }
    }
    if (threadIdx.x == 128) tma_store_wait_all();   // the dS scratch must be complete before the grid retires
  } else {
    // -------------------------------------------------------------------- dQ / dK / dV drain warpgroup
    asm volatile("setmaxnreg.dec.sync.aligned.u32 136;");
    const int quarter = warp & 3;
    const int row_local = quarter * 32 + lane;
    const uint32_t t_q = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + 256;
    for (int t = 0; t < T; ++t) {
      const int r = (rt_begin + t) * kBtBQ + row_local;
      mbar_wait(dq_full, t & 1);
      tc_fence_after();
      float q[64];
      bt_tmem_ld32(t_q, q);
      bt_tmem_ld32(t_q + 32, q + 32);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_free);
      if (r < R) {
        float* dst = dqn + (static_cast<long long>(b) * R + r) * 64;
#pragma unroll
        for (int c = 0; c < 64; c += 4) bt_red4(dst + c, q[c] * scale, q[c + 1] * scale, q[c + 2] * scale, q[c + 3] * scale);
      }
    }
    // dq_full(T-1) also covers the last dV / dK MMAs (tcgen05.commit tracks everything issued before it)
    const int j = j0 + row_local;
    float a[64];
    bt_tmem_ld32(t_q + 64, a);         // dK: cols 320..383
    bt_tmem_ld32(t_q + 96, a + 32);
    tmem_ld_wait();
    if (j < N) {
      float* dst = dkvn + (static_cast<long long>(b) * N + j) * 128;
#pragma unroll
      for (int c = 0; c < 64; c += 4) bt_red4(dst + c, a[c] * scale, a[c + 1] * scale, a[c + 2] * scale, a[c + 3] * scale);
    }
    bt_tmem_ld32(t_q + 128, a);        // dV: cols 384..447
    bt_tmem_ld32(t_q + 160, a + 32);
    tmem_ld_wait();
    if (j < N) {
      float* dst = dkvn + (static_cast<long long>(b) * N + j) * 128 + 64;
#pragma unroll
      for (int c = 0; c < 64; c += 4) bt_red4(dst + c, a[c], a[c + 1], a[c + 2], a[c + 3]);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dTable[hh, delta] += sum_b sum_{i >= delta} dS[b, (i, hh), i - delta]
// grid: (ceil(N / 256), h, B * ichunks); thread = one delta; coalesced along delta (consecutive keys of one row).
__global__ void __launch_bounds__(256)
attn_dbias_kernel(const __nv_bfloat16* __restrict__ ds, long ld_row, long ld_batch, float* __restrict__ dtable,
                  int table_ld, int N, int h, int ichunk, int n_ichunks) {
  const int delta = blockIdx.x * 256 + threadIdx.x;
  const int hh = blockIdx.y;
  const int b = blockIdx.z / n_ichunks, ic = blockIdx.z - b * n_ichunks;
  if (delta >= N) return;
  const int i0 = max(delta, ic * ichunk), i1 = min(N, (ic + 1) * ichunk);
  float acc = 0.f;
  const __nv_bfloat16* base = ds + static_cast<long>(b) * ld_batch;
  for (int i = i0; i < i1; ++i) acc += __bfloat162float(base[static_cast<long>(i * h + hh) * ld_row + (i - delta)]);
  if (acc != 0.f) atomicAdd(&dtable[hh * table_ld + delta], acc);
}

// D[r] = sum_d dO[r, d] * O[r, d]
__global__ void __launch_bounds__(256)
attn_bwd_tc_dsum_kernel(const __nv_bfloat16* __restrict__ d_o, const __nv_bfloat16* __restrict__ o,
                        float* __restrict__ dsum, long rows) {
  const long r = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 3;
  const int sub = threadIdx.x & 7;
  float s = 0.f;
  if (r < rows) {
    const uint4 a = *reinterpret_cast<const uint4*>(d_o + r * 64 + sub * 8);
    const uint4 bq = *reinterpret_cast<const uint4*>(o + r * 64 + sub * 8);
    const uint32_t aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 x = unpack_bf16x2(aa[i]), y = unpack_bf16x2(bb[i]);
      s += x.x * y.x + x.y * y.y;
    }
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (r < rows && sub == 0) dsum[r] = s;
}

}  // namespace omlm

extern "C" int omlm_attn_bwd_tc(const void* qn, const void* kvn, const void* d_o, const void* o, const float* lse2,
                                const float* table, int table_ld, const unsigned char* key_mask, float* dsum_scratch,
                                void* ds_scratch, float* dqn, float* dkvn, float* dtable, int B, int N, int heads,
                                float scale, void* stream) {
  using namespace omlm;
  OMLM_CHECK_ARG(B > 0 && N > 0 && heads > 0, "attn_bwd_tc: bad shape");
  OMLM_CHECK_ARG(table_ld >= N, "attn_bwd_tc: bias table shorter than the sequence");
  auto st = reinterpret_cast<cudaStream_t>(stream);
  const long R = static_cast<long>(N) * heads;
  const long rows = static_cast<long>(B) * R;
  attn_bwd_tc_dsum_kernel<<<static_cast<int>((rows * 8 + 255) / 256), 256, 0, st>>>(
      reinterpret_cast<const __nv_bfloat16*>(d_o), reinterpret_cast<const __nv_bfloat16*>(o), dsum_scratch, rows);
  OMLM_LAUNCH_CHECK();
  const int Wd = (kBtBQ + heads - 1) / heads + 1 + (kBtBK - 1);
  int W = Wd;
  const int want = (32 % heads == 0) ? (32 / heads) % 32 : 1;
  while ((32 % heads == 0) ? (W % 32 != want) : (W % 2 == 0)) ++W;
  const int smem_bytes = kBoBias + 2 * heads * W * 4 + 1024;
  OMLM_CHECK_ARG(smem_bytes <= 232448, "attn_bwd_tc: too many heads (%d)", heads);
  const long Ns = (static_cast<long>(N) + 127) / 128 * 128;    // dS scratch row pitch (keys)
  CUtensorMap tmQ, tmDO, tmKV, tmDS;
  int rc = make_tmap_bf16_2d(&tmQ, qn, 64, static_cast<uint64_t>(rows), 128, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmDO, d_o, 64, static_cast<uint64_t>(rows), 128, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tmKV, kvn, 128, static_cast<uint64_t>(B) * N, 256, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tmDS, ds_scratch, static_cast<uint64_t>(Ns), static_cast<uint64_t>(R), static_cast<uint64_t>(B),
                         static_cast<uint64_t>(Ns) * 2, static_cast<uint64_t>(Ns) * 2 * R, 64, 128);
  if (rc) return rc;
  static int configured = 0;
  if (configured < smem_bytes) {
    OMLM_CUDA(cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    configured = smem_bytes;
  }
  const int n_row_tiles = static_cast<int>((R + kBtBQ - 1) / kBtBQ);
  const int n_key_tiles = (N + kBtBK - 1) / kBtBK;
  int tiles_per_chunk = n_row_tiles;
  for (int cand = 4; cand <= n_row_tiles; cand *= 2) {
    long units = 0;
    for (int kt = 0; kt < n_key_tiles; ++kt) {
      const int first = (kt * kBtBK * heads) / kBtBQ;
      units += (n_row_tiles - first + cand - 1) / cand;
    }
    if (units * B <= 4L * num_sms()) { tiles_per_chunk = cand; break; }
  }
  int units_per_batch = 0;
  for (int kt = 0; kt < n_key_tiles; ++kt) {
    const int first = (kt * kBtBK * heads) / kBtBQ;
    units_per_batch += (n_row_tiles - first + tiles_per_chunk - 1) / tiles_per_chunk;
  }
  attn_bwd_tc_kernel<<<B * units_per_batch, kBtThreads, smem_bytes, st>>>(
      tmQ, tmDO, tmKV, tmDS, lse2, dsum_scratch, table, table_ld, key_mask, dqn, dkvn, N, heads, scale, W, Wd,
      tiles_per_chunk, units_per_batch);
  OMLM_LAUNCH_CHECK();
  const int ichunk = 64;
  const int n_ichunks = (N + ichunk - 1) / ichunk;
  dim3 grid((N + 255) / 256, heads, B * n_ichunks);
  attn_dbias_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(ds_scratch), Ns, Ns * R, dtable, table_ld, N,
                                          heads, ichunk, n_ichunks);
  OMLM_LAUNCH_CHECK();
  return 0;
}
