// Batched per-column onesweep LSD radix sort (see sort.cuh for the layout).
#include "sort.cuh"

#include <vector>

namespace pbl {

namespace {

constexpr uint32_t kSpinLimit = 1u << 24;  // look-back watchdog: fail loudly instead of hanging

__device__ __forceinline__ uint32_t digit_of(uint64_t key, int shift) {
  return (uint32_t)(key >> shift) & (kRadix - 1);
}

// --------------------------------------------------------------------------------------
// Kernel 1: all 8 digit histograms of every column in one read of the input (8 B / key),
// plus the NaN check the reference gets from scipy's check_finite (correlation.py:409).
// Shared-memory atomics; the two most significant digits (sign/exponent bits: few distinct
// values for real data, i.e. same-address conflicts) are warp-aggregated with match.any.
// --------------------------------------------------------------------------------------
template <int BLOCK, int UNROLL>
__global__ void __launch_bounds__(BLOCK)
sort_hist_kernel(const double* __restrict__ raw, int64_t row_stride, int64_t col_stride,
                 uint32_t n, uint32_t* __restrict__ hist, uint32_t* __restrict__ error_flag) {
  __shared__ uint32_t sh[kNumPasses][kRadix];
  for (int i = threadIdx.x; i < kNumPasses * kRadix; i += BLOCK) (&sh[0][0])[i] = 0;
  __syncthreads();
  const int col = blockIdx.y;
  const double* colp = raw + (int64_t)col * col_stride;
  const uint32_t lane = lane_id();
  bool saw_nan = false;
  const uint64_t step = (uint64_t)gridDim.x * BLOCK * UNROLL;
  // the loop bound is warp-uniform so that match.any sees the full warp
  for (uint64_t base = (uint64_t)blockIdx.x * BLOCK * UNROLL; base < n; base += step) {
    uint64_t k[UNROLL];
    bool valid[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      uint64_t i = base + (uint64_t)u * BLOCK + threadIdx.x;
      valid[u] = i < n;
      double d = valid[u] ? ld_stream_f64(colp + (int64_t)i * row_stride) : 0.0;
      saw_nan |= (d != d);
      k[u] = flip_f64((uint64_t)__double_as_longlong(d));
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (valid[u]) {
#pragma unroll
        for (int p = 0; p < kNumPasses - 2; ++p) atomicAdd(&sh[p][digit_of(k[u], 8 * p)], 1u);
      }
      uint32_t top = valid[u] ? (uint32_t)(k[u] >> 48) : 0xFFFFFFFFu;
      uint32_t m = __match_any_sync(0xFFFFFFFFu, top);
      if (valid[u] && lane == (uint32_t)(__ffs(m) - 1)) {
        uint32_t c = __popc(m);
        atomicAdd(&sh[kNumPasses - 2][top & 255u], c);
        atomicAdd(&sh[kNumPasses - 1][top >> 8], c);
      }
    }
  }
  __syncthreads();
  uint32_t* h = hist + (size_t)col * kNumPasses * kRadix;
  for (int i = threadIdx.x; i < kNumPasses * kRadix; i += BLOCK) {
    uint32_t v = (&sh[0][0])[i];
    if (v) atomicAdd(&h[i], v);
  }
  if (saw_nan) error_flag[1] = 1u;
}

// block-wide exclusive scan over kRadix values held one per thread (blockDim.x == kRadix)
__device__ __forceinline__ uint32_t block_excl_scan_256(uint32_t v, uint32_t* s_wsum) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= (uint32_t)d) incl += t;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  uint32_t prefix = 0;
#pragma unroll
  for (int w = 0; w < kRadix / 32; ++w)
    if ((uint32_t)w < warp) prefix += s_wsum[w];
  __syncthreads();
  return prefix + incl - v;
}

// --------------------------------------------------------------------------------------
// Kernel 2: histograms -> exclusive bin bases, and the per-column pass plan (which digit
// passes are no-ops, which buffer each pass reads).  One block of 256 threads per column.
// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRadix)
sort_scan_kernel(uint32_t* __restrict__ hist, PassPlan* __restrict__ plan, uint32_t n) {
  __shared__ uint32_t s_wsum[kRadix / 32];
  __shared__ int s_const[kNumPasses];
  const int col = blockIdx.x;
  if (threadIdx.x < kNumPasses) s_const[threadIdx.x] = 0;
  __syncthreads();
  for (int p = 0; p < kNumPasses; ++p) {
    uint32_t* h = hist + ((size_t)col * kNumPasses + p) * kRadix;
    uint32_t c = h[threadIdx.x];
    if (c == n) s_const[p] = 1;
    uint32_t e = block_excl_scan_256(c, s_wsum);
    h[threadIdx.x] = e;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    PassPlan pp;
    int cur = 0;
    for (int p = 0; p < kNumPasses; ++p) {
      bool run = !s_const[p];
      // the data must leave the caller's array at least once: force the last pass if needed
      if (p == kNumPasses - 1 && cur == 0) run = true;
      pp.run[p] = run ? 1 : 0;
      pp.src[p] = (uint8_t)cur;
      if (run) cur = (cur == 1) ? 2 : 1;
    }
    pp.final_buf = (uint8_t)cur;
    for (int i = 0; i < 7; ++i) pp.pad[i] = 0;
    plan[col] = pp;
  }
}

// --------------------------------------------------------------------------------------
// Fallback (debug) path without look-back: per-tile digit counts + a serial scan over tiles.
// --------------------------------------------------------------------------------------
template <int BLOCK, int ITEMS>
__global__ void __launch_bounds__(BLOCK)
tile_hist_kernel(const double* __restrict__ raw, int64_t row_stride, int64_t col_stride,
                 const uint64_t* __restrict__ keysA, const uint64_t* __restrict__ keysB,
                 uint32_t* __restrict__ tile_counts, const PassPlan* __restrict__ plan,
                 uint32_t n, int pass, int ntiles) {
  constexpr int TILE = BLOCK * ITEMS;
  __shared__ uint32_t sh[kRadix];
  const int col = blockIdx.y;
  if (!plan[col].run[pass]) return;
  const int src = plan[col].src[pass];
  for (int i = threadIdx.x; i < kRadix; i += BLOCK) sh[i] = 0;
  __syncthreads();
  const uint32_t tile = blockIdx.x;
  const uint32_t start = tile * TILE;
  const uint32_t nvalid = min((uint32_t)TILE, n - start);
  for (uint32_t pos = threadIdx.x; pos < nvalid; pos += BLOCK) {
    uint64_t k;
    if (src == 0) {
      double d = raw[(int64_t)col * col_stride + (int64_t)(start + pos) * row_stride];
      k = flip_f64((uint64_t)__double_as_longlong(d));
    } else {
      k = (src == 1 ? keysA : keysB)[(size_t)col * n + start + pos];
    }
    atomicAdd(&sh[digit_of(k, 8 * pass)], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRadix; i += BLOCK)
    tile_counts[((size_t)col * ntiles + tile) * kRadix + i] = sh[i];
}

__global__ void tile_scan_kernel(uint32_t* __restrict__ tile_counts,
                                 const PassPlan* __restrict__ plan, int pass, int ntiles) {
  const int col = blockIdx.x;
  if (!plan[col].run[pass]) return;
  uint32_t run = 0;
  for (int t = 0; t < ntiles; ++t) {
    uint32_t* p = &tile_counts[((size_t)col * ntiles + t) * kRadix + threadIdx.x];
    uint32_t c = *p;
    *p = run;
    run += c;
  }
}

// --------------------------------------------------------------------------------------
// Kernel 3: one digit pass.  Tile = BLOCK*ITEMS keys, warp-striped so that
// (warp, item, lane) order == position order (each pass must be stable).
//   load -> match.any ranking into warp-private histograms -> scan over warps and bins
//   -> publish tile counts / look back for the exclusive tile prefix per bin
//   -> scatter keys+rows into shared memory in digit order -> coalesced runs to HBM.
// --------------------------------------------------------------------------------------
template <int BLOCK, int ITEMS>
__global__ void __launch_bounds__(BLOCK, 3)
onesweep_pass_kernel(const double* __restrict__ raw, int64_t row_stride, int64_t col_stride,
                     uint64_t* __restrict__ keysA, uint64_t* __restrict__ keysB,
                     uint32_t* __restrict__ valsA, uint32_t* __restrict__ valsB,
                     const uint32_t* __restrict__ bin_base_all, uint32_t* __restrict__ status,
                     uint32_t* __restrict__ tile_counter, const PassPlan* __restrict__ plan,
                     uint32_t* __restrict__ error_flag, uint32_t n, int pass, int ntiles,
                     int use_lookback) {
  constexpr int TILE = BLOCK * ITEMS;
  constexpr int NWARPS = BLOCK / 32;
  static_assert(BLOCK >= kRadix && BLOCK % 32 == 0, "one thread per bin is assumed");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem_raw);          // [TILE]
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys + TILE);     // [TILE]
  uint32_t* s_hist = s_vals + TILE;                                  // [NWARPS][kRadix]
  uint32_t* s_goff = s_hist + NWARPS * kRadix;                       // [kRadix]
  uint32_t* s_wsum = s_goff + kRadix;                                // [8]
  uint32_t* s_tile = s_wsum + 8;                                     // [1]

  const int col = blockIdx.y;
  const int tid = threadIdx.x;
  if (!plan[col].run[pass]) return;
  const int src = plan[col].src[pass];
  const int dst = (src == 1) ? 2 : 1;
  const int shift = pass * kRadixBits;

  if (tid == 0) *s_tile = atomicAdd(&tile_counter[col], 1u);
  for (int i = tid; i < NWARPS * kRadix; i += BLOCK) s_hist[i] = 0;
  __syncthreads();
  const uint32_t tile = *s_tile;
  const uint32_t tile_start = tile * (uint32_t)TILE;
  const uint32_t nvalid = min((uint32_t)TILE, n - tile_start);
  const uint32_t warp = tid >> 5, lane = tid & 31;
  const uint32_t pos0 = warp * (ITEMS * 32) + lane;

  // ---- load keys (padding = all-ones: last bin, after every real key) ----
  uint64_t key[ITEMS];
  if (src == 0) {
    const double* colp = raw + (int64_t)col * col_stride + (int64_t)tile_start * row_stride;
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
      uint32_t pos = pos0 + u * 32;
      key[u] = ~0ull;
      if (pos < nvalid) {
        double d = ld_stream_f64(colp + (int64_t)pos * row_stride);
        key[u] = flip_f64((uint64_t)__double_as_longlong(d));
      }
    }
  } else {
    const uint64_t* kin = (src == 1 ? keysA : keysB) + (size_t)col * n + tile_start;
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
      uint32_t pos = pos0 + u * 32;
      key[u] = (pos < nvalid) ? ld_stream_u64(kin + pos) : ~0ull;
    }
  }

  // ---- rank inside the warp ----
  uint32_t rank[ITEMS];
  uint32_t* wh = s_hist + warp * kRadix;
  const uint32_t lt = lanemask_lt();
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    uint32_t bin = digit_of(key[u], shift);
    uint32_t m = __match_any_sync(0xFFFFFFFFu, bin);
    uint32_t below = __popc(m & lt);
    int leader = 31 - __clz(m);  // highest lane of the group: below + 1 == group size
    uint32_t base = 0;
    if ((int)lane == leader) {
      base = wh[bin];
      wh[bin] = base + below + 1;
    }
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    rank[u] = base + below;
    __syncwarp();
  }
  __syncthreads();

  // ---- per bin: exclusive scan over warps, tile total, exclusive scan over bins ----
  uint32_t cnt = 0, bin_start = 0;
  if (tid < kRadix) {
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) {
      uint32_t c = s_hist[w * kRadix + tid];
      s_hist[w * kRadix + tid] = cnt;
      cnt += c;
    }
    if (tid == kRadix - 1) cnt -= (uint32_t)TILE - nvalid;  // drop the padding keys
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= (uint32_t)d) incl += t;
    }
    if (lane == 31) s_wsum[warp] = incl;
    bin_start = incl - cnt;
  }
  __syncthreads();
  if (tid < kRadix) {
#pragma unroll
    for (int w = 0; w < kRadix / 32; ++w)
      if ((uint32_t)w < warp) bin_start += s_wsum[w];

    // ---- exclusive prefix of this bin over all earlier tiles ----
    uint32_t excl = 0;
    uint32_t* st = status + (size_t)col * ntiles * kRadix;
    if (use_lookback) {
      if (tile == 0) {
        st_relaxed_u32(&st[tid], cnt | kFlagInclusive);
      } else {
        st_relaxed_u32(&st[(size_t)tile * kRadix + tid], cnt | kFlagPartial);
        int64_t t = (int64_t)tile - 1;
        uint32_t spins = 0;
        while (true) {
          uint32_t w = ld_relaxed_u32(&st[(size_t)t * kRadix + tid]);
          if ((w & (kFlagInclusive | kFlagPartial)) == 0) {
            if (++spins > kSpinLimit) {
              atomicExch(&error_flag[0], 1u);
              break;
            }
            continue;
          }
          excl += w & kValueMask;
          if (w & kFlagInclusive) break;
          --t;
        }
        st_relaxed_u32(&st[(size_t)tile * kRadix + tid], ((excl + cnt) & kValueMask) | kFlagInclusive);
      }
    } else {
      excl = st[(size_t)tile * kRadix + tid];
    }
    const uint32_t* bin_base = bin_base_all + ((size_t)col * kNumPasses + pass) * kRadix;
    s_goff[tid] = bin_base[tid] + excl - bin_start;  // + position in tile order = global slot
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) s_hist[w * kRadix + tid] += bin_start;
  }
  __syncthreads();

  // ---- scatter to shared memory in digit order ----
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    uint32_t bin = digit_of(key[u], shift);
    rank[u] += wh[bin];
    s_keys[rank[u]] = key[u];
  }
  if (src == 0) {
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) s_vals[rank[u]] = tile_start + pos0 + u * 32;
  } else {
    const uint32_t* vin = (src == 1 ? valsA : valsB) + (size_t)col * n + tile_start;
    uint32_t v[ITEMS];
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
      uint32_t pos = pos0 + u * 32;
      v[u] = (pos < nvalid) ? ld_stream_u32(vin + pos) : 0u;
    }
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) s_vals[rank[u]] = v[u];
  }
  __syncthreads();

  // ---- coalesced runs out to HBM ----
  uint64_t* kout = (dst == 1 ? keysA : keysB) + (size_t)col * n;
  uint32_t* vout = (dst == 1 ? valsA : valsB) + (size_t)col * n;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    uint32_t pos = j * BLOCK + tid;
    if (pos < nvalid) {
      uint64_t k = s_keys[pos];
      uint32_t g = s_goff[digit_of(k, shift)] + pos;
      kout[g] = k;
      vout[g] = s_vals[pos];
    }
  }
}

template <int BLOCK, int ITEMS>
constexpr size_t pass_smem_bytes() {
  return (size_t)BLOCK * ITEMS * 12 + (size_t)(BLOCK / 32) * kRadix * 4 + kRadix * 4 + 8 * 4 + 16;
}

}  // namespace

namespace {
struct PassEvent {
  cudaEvent_t start, stop;
  int64_t keys;
};
bool g_profile = false;
std::vector<PassEvent> g_events;
}  // namespace

void sort_profile_enable(bool on) { g_profile = on; }

void sort_profile_read(int64_t* launches, double* total_ms, int64_t* keys) {
  int64_t nl = 0, nk = 0;
  double ms = 0.0;
  for (auto& e : g_events) {
    float t = 0.f;
    if (cudaEventSynchronize(e.stop) == cudaSuccess &&
        cudaEventElapsedTime(&t, e.start, e.stop) == cudaSuccess) {
      ++nl;
      ms += t;
      nk += e.keys;
    }
    cudaEventDestroy(e.start);
    cudaEventDestroy(e.stop);
  }
  g_events.clear();
  if (launches) *launches = nl;
  if (total_ms) *total_ms = ms;
  if (keys) *keys = nk;
}

size_t sort_status_bytes(int ncols, uint32_t n) {
  size_t ntiles = ((size_t)n + kSortTile - 1) / kSortTile;
  return (size_t)ncols * ntiles * kRadix * sizeof(uint32_t);
}

int sort_columns_f64(const double* in, int64_t row_stride, int64_t col_stride, uint32_t n,
                     int ncols, const SortBuffers& buf, bool use_lookback, cudaStream_t stream) {
  if (n == 0 || ncols <= 0) return kOk;
  if (n > kMaxSortN) {
    set_last_error("sort_columns_f64: n exceeds 2^30-1 rows per column");
    return kBadShape;
  }
  const int ntiles = (int)(((size_t)n + kSortTile - 1) / kSortTile);
  PBL_CUDA_CHECK(cudaMemsetAsync(buf.hist, 0, (size_t)ncols * kNumPasses * kRadix * 4, stream));
  PBL_CUDA_CHECK(cudaMemsetAsync(buf.tile_counter, 0, (size_t)ncols * kNumPasses * 4, stream));

  {
    constexpr int HB = 512, HU = 4;
    int per_col = (int)(((size_t)n + HB * HU - 1) / (HB * HU));
    int want = (num_sms() * 4 + ncols - 1) / ncols;  // ~4 resident blocks per SM over the batch
    dim3 grid((unsigned)max(1, min(per_col, want)), (unsigned)ncols);
    sort_hist_kernel<HB, HU><<<grid, HB, 0, stream>>>(in, row_stride, col_stride, n, buf.hist,
                                                     buf.error_flag);
    PBL_LAUNCH_CHECK();
  }
  sort_scan_kernel<<<ncols, kRadix, 0, stream>>>(buf.hist, buf.plan, n);
  PBL_LAUNCH_CHECK();

  auto kern = onesweep_pass_kernel<kSortBlock, kSortItems>;
  constexpr size_t smem = pass_smem_bytes<kSortBlock, kSortItems>();
  static bool attr_set = false;
  if (!attr_set) {
    PBL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const size_t status_bytes = sort_status_bytes(ncols, n);
  for (int pass = 0; pass < kNumPasses; ++pass) {
    if (use_lookback) {
      PBL_CUDA_CHECK(cudaMemsetAsync(buf.status, 0, status_bytes, stream));
    } else {
      tile_hist_kernel<kSortBlock, kSortItems><<<dim3(ntiles, ncols), kSortBlock, 0, stream>>>(
          in, row_stride, col_stride, buf.keysA, buf.keysB, buf.status, buf.plan, n, pass, ntiles);
      PBL_LAUNCH_CHECK();
      tile_scan_kernel<<<ncols, kRadix, 0, stream>>>(buf.status, buf.plan, pass, ntiles);
      PBL_LAUNCH_CHECK();
    }
    PassEvent ev{};
    if (g_profile) {
      cudaEventCreate(&ev.start);
      cudaEventCreate(&ev.stop);
      ev.keys = (int64_t)n * ncols;
      cudaEventRecord(ev.start, stream);
    }
    kern<<<dim3(ntiles, ncols), kSortBlock, smem, stream>>>(
        in, row_stride, col_stride, buf.keysA, buf.keysB, buf.valsA, buf.valsB, buf.hist,
        buf.status, buf.tile_counter + (size_t)pass * ncols, buf.plan, buf.error_flag, n, pass,
        ntiles, use_lookback ? 1 : 0);
    if (g_profile) {
      cudaEventRecord(ev.stop, stream);
      g_events.push_back(ev);
    }
    PBL_LAUNCH_CHECK();
  }
  return kOk;
}

}  // namespace pbl
