// Batched per-column LSD radix sort ("onesweep": one read + one write of every key per digit
// pass, chained-scan decoupled look-back between tiles) of 64-bit keys with 32-bit row-index
// payloads.  This is the engine under both ranking steps of Iman-Conover; it replaces NumPy's
// stable argsort inside scipy.stats.rankdata (reference call sites
// src/probabilit/correlation.py:394 and :422) and np.sort (:423).
//
// Layout in HBM (all SoA, one contiguous run per column):
//   keys  [ncols][n]  u64   order-preserving image of the fp64 values (common.cuh::flip_f64)
//   vals  [ncols][n]  u32   source row of each key
//   hist  [ncols][8][256] u32   digit histograms -> exclusive bin bases
//   status[ncols][ntiles][256] u32   look-back words: bit31 inclusive, bit30 partial, 30-bit count
// Columns are independent: blockIdx.y is the column, so one launch per digit pass covers the
// whole batch.  The first pass reads the caller's doubles in place (any row stride) and
// synthesises the payload (row index), so X is never copied or converted up front.
//
// Digit passes in which every key of a column has the same digit are skipped on the device
// (no host round trip): the scan kernel records, per column and pass, which ping-pong buffer
// is the source, and whether the pass runs at all.
#pragma once
#include "common.cuh"

namespace pbl {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kNumPasses = 64 / kRadixBits;

constexpr uint32_t kFlagInclusive = 0x80000000u;
constexpr uint32_t kFlagPartial = 0x40000000u;
constexpr uint32_t kValueMask = 0x3FFFFFFFu;
constexpr uint32_t kMaxSortN = 0x3FFFFFFFu;  // 30-bit counts in the look-back words

// Per (column, pass) routing decided on the device by sort_scan_kernel.
struct PassPlan {
  uint8_t run[kNumPasses];  // 1: the pass moves data, 0: skipped (constant digit)
  uint8_t src[kNumPasses];  // 0: raw input column, 1: buffer A, 2: buffer B
  uint8_t final_buf;        // buffer holding the sorted column after the last pass (0/1/2)
  uint8_t pad[7];
};

struct SortBuffers {
  uint64_t* keysA = nullptr;
  uint64_t* keysB = nullptr;
  uint32_t* valsA = nullptr;
  uint32_t* valsB = nullptr;
  uint32_t* hist = nullptr;      // [ncols][8][256]
  uint32_t* status = nullptr;    // [ncols][ntiles][256]
  uint32_t* tile_counter = nullptr;  // [8 passes + 1 scatter pass][ncols]
  PassPlan* plan = nullptr;      // [ncols]
  uint32_t* error_flag = nullptr;    // [4]: watchdog / NaN flags
};

// tile size of the partition kernel in use (PBL_SORT_CFG selects the instantiation)
int sort_tile_size();
size_t sort_status_bytes(int ncols, uint32_t n);

// Optional CUDA-event timing of every digit-pass launch (bench.py's roofline leg): when enabled,
// sort_columns_f64 brackets each onesweep_pass_kernel launch with events on the launching stream.
void sort_profile_enable(bool on);
// Drains the recorded events (synchronises them): number of pass launches, their total
// duration in ms and the number of keys they moved.
void sort_profile_read(int64_t* launches, double* total_ms, int64_t* keys);

// Sort `ncols` columns of n doubles each; column c starts at in + c*col_stride and its rows are
// row_stride elements apart.  On return (stream order) column c's sorted keys/rows are in
// buffer plan[c].final_buf (1 = A, 2 = B).  error_flag[1] is set if any input is NaN.
int sort_columns_f64(const double* in, int64_t row_stride, int64_t col_stride, uint32_t n,
                     int ncols, const SortBuffers& buf, bool use_lookback, cudaStream_t stream);

// "Scatter by row" that follows each sort: out[col][rows[p] * row_stride] = value[p], where
// rows[] is the sorted payload (vals[final]) and value[] was staged by the caller in keys[other].
// For long columns it runs as a partition pass into <= 256 L2-sized row windows followed by a
// window-local scatter (see sort.cu); short columns are scattered directly.
int scatter_by_row(uint32_t n, int ncols, const SortBuffers& buf, double* out, int64_t row_stride,
                   int64_t col_stride, bool use_lookback, cudaStream_t stream);

}  // namespace pbl
