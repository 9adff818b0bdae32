/* libomlm_b200 — C ABI of the B200-native hot path of zhvng/open-musiclm.
 *
 * The reference has no native code and therefore no FFI of its own: its hot path is the chain of
 * torch calls inside open_musiclm/transformer.py and open_musiclm/open_musiclm.py.  Each entry point
 * below replaces one such call site (cited as file:line relative to the reference tree).  All
 * functions
 *   - take only PODs (device pointers, sizes, float scalars, a cudaStream_t passed as void*),
 *   - are asynchronous (enqueue-only on the given stream) and allocate no persistent memory,
 *   - return 0 on success, 1 on argument errors, 1000+cudaError_t on CUDA errors;
 *     omlm_last_error() returns a thread-local description.
 * There is no CPU fallback: without an sm_100a device every compute entry point fails.
 */
#ifndef OMLM_B200_H_
#define OMLM_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define OMLM_B200_ABI_VERSION 1

const char* omlm_last_error(void);
int omlm_abi_version(void);
/* 0 iff the current CUDA device is compute capability 10.x. */
int omlm_device_check(void);

/* bf16 GEMM on tcgen05 tensor cores:  out[m,n] = alpha * sum_k A(m,k) * B(n,k) (+ addend[m,n]).
 *   a_mn_major = 0: A is [M, lda] with k contiguous;  1: A is [K, lda] with m contiguous.
 *   b_mn_major = 0: B is [N, ldb] with k contiguous;  1: B is [K, ldb] with n contiguous.
 *   out_f32: 0 -> bf16 out, 1 -> fp32 out.  addend (fp32, may alias out) gives residual add /
 *   beta=1 accumulation.  splits > 1: split-K with fp32 atomic accumulation into out.
 *   row_split/row_valid: output-row compaction for the padded GEGLU weight layout (0 = off).
 *   n_valid: number of live output columns (<= N; 0 = N).  block_n in {128, 256}.
 * Replaces nn.Linear / einsum: transformer.py:144,149,254,333; open_musiclm.py:173,181 and their
 * autograd backward GEMMs. */
int omlm_gemm_bf16(const void* A, int a_mn_major, long lda, const void* B, int b_mn_major, long ldb,
                   int M, int N, int K, void* out, int out_f32, long ldo, const float* addend,
                   long ldadd, float alpha, int splits, int row_split, int row_valid, int n_valid,
                   int block_n, int max_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OMLM_B200_H_ */
