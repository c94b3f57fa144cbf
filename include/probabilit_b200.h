/* probabilit_b200 -- C ABI of the B200-native sampling hot path of tommyod/probabilit.
 *
 * The reference is pure Python; its "plugin seam" is duck-typed:
 *   correlator protocol   Correlator.set_target(C) -> self ; correlator(X) -> X'
 *                         (reference src/probabilit/correlation.py:162-179, :368-425;
 *                          called from src/probabilit/modeling.py:577-581)
 * Each entry point below names the reference interface it replaces.  All functions return a
 * pbl_status; pbl_last_error() gives the message of the last failure on the calling thread.
 * Pointers named *_dev are device pointers on the current CUDA device, everything else is host
 * memory.  Matrices are (n rows = observations, k columns = variables), fp64, addressed as
 * base[row * row_stride + col * col_stride] with strides in ELEMENTS (NumPy F-order: row_stride 1,
 * col_stride n; C-order: row_stride k, col_stride 1).  No call takes ownership of caller memory;
 * inputs are never modified (the reference writes into np.empty_like(X), correlation.py:418).
 */
#ifndef PROBABILIT_B200_H
#define PROBABILIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PBL_API __attribute__((visibility("default")))

typedef enum pbl_status {
  PBL_OK = 0,
  PBL_NOT_POSITIVE_DEFINITE = 1, /* ValueError "Rank data correlation not positive definite." correlation.py:399-403 */
  PBL_NON_FINITE = 2,            /* ValueError "array must not contain infs or NaNs" (scipy check_finite at :409) */
  PBL_BAD_SHAPE = 3,             /* ValueError from Correlator._validate_X, correlation.py:181-202 */
  PBL_CUDA_ERROR = 4,
  PBL_INTERNAL = 5,
  PBL_RETRY = 6 /* stage API only (pbl_ic_stage_status): the data are too dense for the 40-bit sort window;
                   the plan has switched to the exact 64-bit sort, repeat from pbl_ic_stage_begin.
                   pbl_ic_plan_run / pbl_iman_conover_f64 handle this internally. */
} pbl_status;

/* ---- library ---- */
PBL_API int pbl_version(void);
PBL_API const char* pbl_last_error(void);
PBL_API int pbl_device_count(void);
PBL_API int pbl_set_device(int device);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
PBL_API int64_t pbl_kernel_launches(void);

/* CUDA-event timing of the radix-sort digit passes (the dominant kernel), for bench.py's
 * roofline leg: enable, run, then read (#launches, their total ms, keys moved) and reset. */
PBL_API int pbl_sort_profile_enable(int on);
PBL_API int pbl_sort_profile_read(int64_t* launches, double* total_ms, int64_t* keys);

/* ---- memory helpers (so that a host language without a CUDA binding can stage data) ---- */
PBL_API int pbl_device_malloc(void** ptr_dev, uint64_t bytes);
PBL_API int pbl_device_free(void* ptr_dev);
PBL_API int pbl_host_malloc_pinned(void** ptr, uint64_t bytes);
PBL_API int pbl_host_free_pinned(void* ptr);
PBL_API int pbl_memcpy_h2d(void* dst_dev, const void* src, uint64_t bytes, void* stream);
PBL_API int pbl_memcpy_d2h(void* dst, const void* src_dev, uint64_t bytes, void* stream);
PBL_API int pbl_stream_synchronize(void* stream);

/* ---- Iman-Conover correlator: ImanConover().set_target(C)(X), correlation.py:288-425 ---- */
typedef struct pbl_ic_plan pbl_ic_plan;

/* Workspace for (n, k) problems on the current device.  col_batch <= 0: choose automatically
 * (columns sorted per launch batch; bounds the sort workspace). */
PBL_API int pbl_ic_plan_create(int64_t n, int32_t k, int32_t col_batch, pbl_ic_plan** plan);
/* flags bit 0 (PBL_IC_ROWS_ONLY): no sort workspace -- the plan serves the Gram / solve / transform
 * stages of a row shard in the multi-GPU driver (the sorts run in a second, column-shard plan). */
#define PBL_IC_ROWS_ONLY 1
PBL_API int pbl_ic_plan_create_ex(int64_t n, int32_t k, int32_t col_batch, int32_t flags, pbl_ic_plan** plan);
PBL_API int pbl_ic_plan_destroy(pbl_ic_plan* plan);
PBL_API uint64_t pbl_ic_plan_bytes(const pbl_ic_plan* plan);

/* Correlator.set_target (correlation.py:162-179): the k x k validation and P = cholesky(C) stay
 * on the host (NumPy); P_lower is that lower-triangular factor, row-major k*k doubles. */
PBL_API int pbl_ic_plan_set_target(pbl_ic_plan* plan, const double* P_lower);

/* ImanConover.__call__ (correlation.py:368-425) on device-resident X -> Y.  Synchronous: returns
 * after the stream has drained, with the status the reference would have raised. */
PBL_API int pbl_ic_plan_run(pbl_ic_plan* plan, const double* X_dev, int64_t x_row_stride,
                    int64_t x_col_stride, double* Y_dev, int64_t y_row_stride,
                    int64_t y_col_stride, void* stream);

/* Same call with HOST buffers (what a NumPy caller holds): copies X to the device, runs, copies Y
 * back.  X and Y must each be one contiguous block in C or F order. */
PBL_API int pbl_iman_conover_f64(const double* X, int64_t n, int32_t k, int64_t x_row_stride,
                         int64_t x_col_stride, const double* P_lower, double* Y,
                         int64_t y_row_stride, int64_t y_col_stride);

/* Stage-level entry points (asynchronous on `stream`): used by the parity tests and by the
 * multi-GPU host driver, which places its collectives between them.
 *   rank_scores : correlation.py:394-395 (+ np.sort of :423) for columns [col0, col0+ncols)
 *   gram        : the reduction inside np.corrcoef, :398
 *   solve       : corrcoef normalisation, cholesky, T = Q^-T P^T, :398-414 (n_total = global rows)
 *   transform   : :409-414 applied to the rows
 *   rank_gather : :419-423 for columns [col0, col0+ncols)
 *   status      : synchronise and report (PBL_OK / NOT_POSITIVE_DEFINITE / NON_FINITE / ...)  */
PBL_API int pbl_ic_stage_begin(pbl_ic_plan* plan, void* stream);
PBL_API int pbl_ic_stage_rank_scores(pbl_ic_plan* plan, const double* X_dev, int64_t row_stride,
                             int64_t col_stride, int32_t col0, int32_t ncols, void* stream);
PBL_API int pbl_ic_stage_gram(pbl_ic_plan* plan, void* stream);
PBL_API int pbl_ic_stage_solve(pbl_ic_plan* plan, int64_t n_total, void* stream);
PBL_API int pbl_ic_stage_transform(pbl_ic_plan* plan, void* stream);
PBL_API int pbl_ic_stage_rank_gather(pbl_ic_plan* plan, double* Y_dev, int64_t row_stride,
                             int64_t col_stride, int32_t col0, int32_t ncols, void* stream);
PBL_API int pbl_ic_stage_status(pbl_ic_plan* plan, void* stream);

/* Device pointers to the plan's intermediates (column-major [k][n] unless noted), for parity
 * tests and for collectives: what = 0 scores / correlated scores, 1 sortedX, 2 gram [k][k],
 * 3 colsum [k], 4 T [k][k] row-major, 5 work (R then Q) [k][k]. */
PBL_API int pbl_ic_plan_buffer(pbl_ic_plan* plan, int32_t what, void** ptr_dev, uint64_t* bytes);

/* ---- unit-cube generators: the `quantiles = ...` draw of Node.sample, src/probabilit/modeling.py:479-489.
 * out_dev is (n, d) fp64 addressed with element strides (column-major: row_stride 1, col_stride n).
 * Every point depends on its global row index only: a row shard passes its first row as
 * row0 / skip / start_index and needs no communication. ---- */

/* method=None: random_state.random((n, d)), modeling.py:485-486.  Philox4x32-10 counter stream
 * (statistical parity with NumPy's generators, not bit parity); row0 must be even. */
PBL_API int pbl_uniform_f64(uint64_t seed, uint64_t row0, int64_t n, int32_t d, double* out_dev,
                            int64_t row_stride, int64_t col_stride, void* stream);

/* method="sobol": scipy.stats.qmc.Sobol(d, rng).random(n), modeling.py:482,488-489.
 *   direction numbers: Joe-Kuo tables (poly[d], vinit[d][vinit_cols], int64) -> sv[d][bits]
 *   scramble: LMS + digital shift from the host generator's random bits (ltm_bits[d][bits][bits],
 *             shift_bits[d][bits], one byte per bit, drawn exactly like scipy draws them)
 *   points skip .. skip+n-1 of the sequence; bit-exact with scipy for the same sv / shift. */
PBL_API int pbl_sobol_direction_numbers(const int64_t* poly_dev, const int64_t* vinit_dev, int32_t vinit_cols,
                                        int32_t d, int32_t bits, uint64_t* sv_dev, void* stream);
PBL_API int pbl_sobol_scramble(const uint8_t* ltm_bits_dev, const uint8_t* shift_bits_dev, int32_t d,
                               int32_t bits, uint64_t* sv_dev, uint64_t* shift_dev, void* stream);
PBL_API int pbl_sobol_f64(const uint64_t* sv_dev, const uint64_t* shift_dev, int32_t d, int32_t bits,
                          uint64_t skip, int64_t n, double* out_dev, int64_t row_stride,
                          int64_t col_stride, void* stream);

/* method="halton": scipy.stats.qmc.Halton(d, rng).random(n), modeling.py:481.  bases[d] = first d
 * primes; perms (or NULL when unscrambled): per dimension perm_count[c] x bases[c] digit
 * permutations starting at perms[perm_off[c]].  Bit-exact with scipy for the same permutations. */
PBL_API int pbl_halton_f64(const int32_t* bases_dev, const int64_t* perms_dev, const int64_t* perm_off_dev,
                           const int32_t* perm_count_dev, int32_t d, uint64_t start_index, int64_t n,
                           double* out_dev, int64_t row_stride, int64_t col_stride, void* stream);

/* method="lhs": scipy.stats.qmc.LatinHypercube(d, rng).random(n), modeling.py:480 / README.md:113.
 * (perm + 1 - U) / n with a counter-based per-column permutation (statistical parity). */
PBL_API int pbl_lhs_f64(uint64_t seed, int64_t n, int32_t d, int32_t scramble, double* out_dev,
                        int64_t row_stride, int64_t col_stride, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PROBABILIT_B200_H */
