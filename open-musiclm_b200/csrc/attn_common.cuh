// Shared pieces of the attention forward/backward kernels: swizzled smem tiles, ldmatrix, mma.sync.
//
// Layout trick used by both directions: multi-query attention with h heads sharing one K/V head is
// run as SINGLE-head attention over R = N*h "folded" query rows per batch element, row r = i*h + hh
// (position i, head hh).  That is exactly the memory order of q [B, N, h*64], so no data movement
// is needed; the causal rule becomes  key j visible  <=>  j <= r / h, a 128-row tile spans only
// 128/h positions (little causal waste), and the MQA reductions over heads in the backward pass
// (dK, dV) fall out of the ordinary sum over query rows.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace omlm {

// [rows x 64] bf16 tile, 128 bytes per row, 16-byte chunks XOR-swizzled by (row & 7).
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ void cp_async16(uint32_t smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// A-fragment (16 rows x 16 k) of a swizzled [rows x 64] tile: rows row0..row0+15, k chunk pair kstep.
__device__ __forceinline__ void load_a_frag(uint32_t tile_base, int row0, int kstep, int lane, uint32_t (&a)[4]) {
  const int r = row0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  const int c = kstep * 2 + (lane >> 4);
  ldsm_x4(tile_base + tile_off(r, c), a[0], a[1], a[2], a[3]);
}
// B-fragments for two adjacent n-tiles (n0..n0+15) at k-step kstep from a tile stored [n][k] (k contiguous):
// b[0],b[1] -> n-tile 0; b[2],b[3] -> n-tile 1.
__device__ __forceinline__ void load_b_frag_nk(uint32_t tile_base, int n0, int kstep, int lane, uint32_t (&b)[4]) {
  const int mi = lane >> 3;
  const int r = n0 + (lane & 7) + (mi >> 1) * 8;
  const int c = kstep * 2 + (mi & 1);
  ldsm_x4(tile_base + tile_off(r, c), b[0], b[1], b[2], b[3]);
}
// B-fragments for two adjacent n-tiles (n chunk pair) at k-step kstep from a tile stored [k][n] (n contiguous):
// uses ldmatrix.trans.  k rows k0..k0+15, n columns n0..n0+15.
__device__ __forceinline__ void load_b_frag_kn(uint32_t tile_base, int k0, int n0, int lane, uint32_t (&b)[4]) {
  const int mi = lane >> 3;
  const int r = k0 + (lane & 7) + (mi & 1) * 8;
  const int c = (n0 >> 3) + (mi >> 1);
  ldsm_x4_t(tile_base + tile_off(r, c), b[0], b[1], b[2], b[3]);
}
// A-fragment (16 m x 16 k) from a tile stored [k][m] (m contiguous), i.e. A = tile^T, via ldmatrix.trans.
// m rows m0..m0+15, k rows k0..k0+15 of the stored tile.
__device__ __forceinline__ void load_a_frag_t(uint32_t tile_base, int k0, int m0, int lane, uint32_t (&a)[4]) {
  // matrices: (m 0-7,k 0-7) (m 8-15,k 0-7) (m 0-7,k 8-15) (m 8-15,k 8-15); stored as [k][m] so each
  // 8x8 block is read transposed.
  const int mi = lane >> 3;
  const int r = k0 + (lane & 7) + (mi >> 1) * 8;
  const int c = (m0 >> 3) + (mi & 1);
  ldsm_x4_t(tile_base + tile_off(r, c), a[0], a[1], a[2], a[3]);
}

constexpr float kLog2e = 1.4426950408889634f;

}  // namespace omlm
