/* cnf_dev.h -- DEVELOPMENT entry points of libcnf_sm100_dev.so.  NOT part of the product ABI.
 *
 * The product library (libcnf_sm100.so, include/cnf.h) does not export any of these and does not compile the instrumented
 * kernel variants.  The same sources built with -DCNF_DEV give libcnf_sm100_dev.so = product ABI + the hooks below: hardware
 * self-tests that pin the tcgen05 descriptor encodings (used by tests/), issue-rate probes and per-stage clock stamps
 * (used by scripts/ while tuning; CNF_LIB_PATH=.../libcnf_sm100_dev.so makes the Python package load this build). */
#ifndef CNF_DEV_H_
#define CNF_DEV_H_

#include "cnf.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Development aid: device buffer of 256 int64 receiving SM-clock timestamps of CTA (0,0): [0..59] / [64..127] stage stamps of
 * the third tile (forward / backward role), [60..63] loop ends / flush end / kernel start, [128..191] / [192..255] tile-start
 * clocks of the first 64 tiles of each role, [256 + 8*cta ..] per-CTA %globaltimer stamps {entry, setup done, forward end,
 * backward end, flush done, tiles} (buffer of 256 + 8 * #CTAs int64).  NULL (default) selects the production kernel without any of it. */
int cnf_set_debug_buffer(void* dev_ptr);

/* tcgen05 self-test: D[128,N] = A[128,K] * B[N,K]^T with bf16 operands staged in shared memory in the
 * canonical no-swizzle UMMA layout; mode 0: A,B K-major; 1: B MN-major (dgrad view); 2: A,B MN-major
 * (wgrad view).  Used by tests to pin the descriptor encodings on hardware.
 * mode 3: A [128,K] staged in TENSOR MEMORY (packed bf16 pairs), B [N,K] K-major in shared memory.
 * mode 4/5: M = 64 weight-gradient view: A given [K,64], B given [K,N], both MN-major; D address lane offset 0 / 16;
 *           D receives ALL 128 lanes x N (sentinel -777 where untouched): row r lands in lane 32*(r/16) + r%16 (+16). */
int cnf_umma_selftest(int mode, int N, int Kdim, const float* A, const float* B, float* D, cnf_stream_t stream);

/* Hardware probe (development aid): where does an M=64 UMMA put its 64 accumulator rows in TMEM, and does the lane
 * field of the D address offset them?  D [128,N] receives all 128 lanes (sentinel -777 where untouched). */
int cnf_umma_m64_probe(int N, int Kdim, int lane_off, const float* A, const float* B, float* D, cnf_stream_t stream);

/* Development probe: one thread per CTA issues `reps` back-to-back tcgen05.mma (M x N x 16, bf16) into `ndst` rotating
 * accumulators and the CTA reports {issue cycles, cycles until the commit arrives} per CTA in out_cycles[2 * n_cta].
 * mode 0: A,B shared memory K-major; 1: A from TMEM; 2: A,B MN-major (weight-gradient view); 3: B MN-major. */
int cnf_umma_rate_probe(int mode, int M, int N, int reps, int ndst, int n_cta, long long* out_cycles, cnf_stream_t stream);

/* development aid: SM-clock stamps of CTA 0's epilogue thread 0 (64 int64 device buffer; NULL switches it off) */
int cnf_occmap_tc_set_debug(void* stamps);

#ifdef __cplusplus
}
#endif
#endif /* CNF_DEV_H_ */
