/*
 * plaac_bench.h -- benchmark utilities exported by libplaac_cuda.so (NOT part of the
 * scoring path): on-device synthetic proteomes and a
 * register-resident FP64 issue-rate microbenchmark for the roofline denominator.
 * All pointers are device pointers on the current CUDA device; work is enqueued on
 * `stream` (a cudaStream_t passed as void*, NULL = default stream).
 */
#ifndef PLAAC_BENCH_H
#define PLAAC_BENCH_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Lengths ~ round(LogNormal(mu, sigma)) clamped to [min_len, max_len]; protein i of the
 * global proteome is keyed by (seed, first_index + i) with Philox4x32-10, so any shard of
 * any GPU count sees the same proteome. */
int plaac_bench_synth_lengths(void *stream, uint64_t seed, int64_t first_index, int64_t nprot, double mu, double sigma,
                              int32_t min_len, int32_t max_len, int64_t *d_lengths);

/* Residues iid from bg_freq (22 numbers, host pointer), X (code 0) at x_rate; with probability
 * prd_rate one segment of length U[60,300] (clipped to the protein) is overwritten with draws
 * from prd_freq (22 numbers, host pointer).  d_offsets = exclusive scan of the lengths (nprot+1). */
int plaac_bench_synth_residues(void *stream, uint64_t seed, int64_t first_index, int64_t nprot, const int64_t *d_offsets,
                               const double *bg_freq, const double *prd_freq, double prd_rate, double x_rate,
                               uint8_t *d_codes);

/* FP64 pipe peak: every SM runs independent register-resident DFMA (fma != 0) or DADD chains.
 * Returns instructions/s per lane summed over the GPU (lane-ops/s) in *ops_per_s. */
int plaac_bench_fp64_peak(int device, int use_fma, double *ops_per_s, float *ms);

#ifdef __cplusplus
}
#endif
#endif
