/*
 * cutter_vad_b200_dev.h -- development hooks of libcvad_b200_dev.so (csrc/cvad_dev.cu).  Not part of the drop-in
 * boundary and not in libcvad_b200.so: probes of the tcgen05 building blocks used by tests/test_gpu_tc_probe.py and
 * tools/tc_*.py.  All return 0 or a negative CVAD_E_* code (cutter_vad_b200.h); cvad_dev_last_error gives text.
 */
#ifndef CUTTER_VAD_B200_DEV_H
#define CUTTER_VAD_B200_DEV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char *cvad_dev_last_error(void);

/* Test hook: one 128x32x256 BF16 tensor-core GEMM (tcgen05, TMEM accumulator) on `device`;
   a_bf16[128][256], b_bf16[32][256] are raw bf16 bit patterns, d_out[128][32] = A * B^T in float32. */
int cvad_tc_probe(int device, const uint16_t *a_bf16, const uint16_t *b_bf16, float *d_out);

/* Measurement hook: issue reps x 4 BF16 MMAs of shape M x N x 16 (both operands in shared memory) on
   `grid` CTAs, cycling over n_acc TMEM accumulators; out2[0] = SM cycles from first issue to completion on CTA 0, out2[1] = MMAs issued. */
int cvad_tc_rate(int device, int M, int N, int reps, int a_tiles, int n_acc, int grid, long long *out2);

/* Test hook: B operand stored MN-major SWIZZLE_64B (N contiguous).  a_bf16[128][64], b_bf16[96][64] ->
   d_out[128][160]: columns 0..95 = A * B^T, columns 96..159 = A * B[32..95]^T (descriptor starting one N atom further). */
int cvad_tc_probe_mn(int device, const uint16_t *a_bf16, const uint16_t *b_bf16, float *d_out);

/* cvad_tc_rate with the B operand MN-major SWIZZLE_64B when b_mn != 0 (N % 32 == 0). */
int cvad_tc_rate_mn(int device, int M, int N, int reps, int a_tiles, int n_acc, int grid, int b_mn, long long *out2);

/* Measurement hook: one thread per CTA streams `tiles` cp.async.bulk copies of tile_bytes through a ring of `depth`
   slots from a src_bytes buffer (L2-resident on the measured run); out2[0] = SM cycles on CTA 0, out2[1] = tiles. */
int cvad_bulk_rate(int device, int tiles, int depth, int tile_bytes, int grid, size_t src_bytes, long long *out2);

#ifdef __cplusplus
}
#endif
#endif /* CUTTER_VAD_B200_DEV_H */
