// The onesweep digit pass as a PERSISTENT kernel with double-buffered bulk-async (TMA) tile loads.
// Included by sort.cu (inside namespace pbl, after PassArgs).
//
// A CTA of 8 warps loops over tiles of 4096 keys.  Each tile lives in one of two 48 KB shared-memory
// stages: thread 0 takes the ticket (col, tile) of the NEXT tile while the current one is being ranked,
// arms that stage's mbarrier with the byte count and issues cp.async.bulk.shared.global for the tile's
// keys (8 B each, or the caller's raw doubles in the first pass) and rows (4 B each) -- SASS: UBLKCP +
// SYNCS -- as soon as the first block barrier of the current tile has shown that the stage's previous
// occupant has been written out.  The warps wait on the mbarrier, pull their 16 keys + rows out of
// the stage into registers (conflict-free LDS), rank them (rank.cuh), and then REUSE the stage as the
// digit-ordered staging area from which coalesced runs go out to HBM.  Two CTAs per SM: while a tile
// is ranked and written out, the next one is already landing, and its ticket was taken a tile
// earlier, so that neither the global-load latency nor the ticket's L2 round trip is on the critical path.
//
// Look-back words: 64-bit, epoch-tagged (sort.cuh) -- the array is never cleared between passes.
//
// Global addresses of a tile are only 8 B (keys) / 4 B (rows) aligned when the column length is odd;
// a bulk copy needs 16 B: the copy covers the enclosing aligned range and the ticket carries the
// offset of the first element (the buffers have 16 B of slack at the end, see ic_plan_create).
#pragma once

constexpr int kTmaWorkers = 256;
constexpr int kTmaItems = 16;
constexpr int kTmaTile = kTmaWorkers * kTmaItems;
constexpr int kTmaThreads = kTmaWorkers;
constexpr int kTmaWarps = kTmaWorkers / 32;
constexpr uint32_t kTmaKeyBytes = (kTmaTile + 2) * 8;  // + one 16 B unit of alignment slack
constexpr uint32_t kTmaValBytes = (kTmaTile + 4) * 4;
constexpr uint32_t kTmaStageBytes = kTmaKeyBytes + kTmaValBytes;
constexpr size_t kTmaSmemBytes = 2 * (size_t)kTmaStageBytes + (size_t)kTmaWarps * kRadix * 4 + kRadix * 4 + 64 + 2 * 16 + 2 * 8;

enum TmaMode : uint32_t { kModeBuffers = 0, kModeRawBulk = 1, kModeRawDirect = 2, kModeEnd = 3 };
struct __align__(16) TmaTicket {
  uint32_t col, tile, nvalid;
  uint32_t mode_off;  // mode | key offset << 8 | row offset << 16   (offsets in elements inside the stage)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// try_wait suspends the thread in hardware until the phase completes or a time limit passes; the loop
// around it is bounded so that a lost transaction traps (CUDA error on the host) instead of hanging
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, uint32_t* error_flag) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
      atomicExch(&error_flag[kFlagWatchdog], 1u);
      __trap();
    }
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

struct TmaArgs {
  PassArgs p;
  const KeyMap* maps;       // [ncols], written by sort_scan_kernel
  uint64_t* status64;       // [ncols][ntiles][256]
  uint32_t* ticket;         // one counter for the launch
  uint32_t total_tiles;     // ncols * ntiles
  uint32_t epoch;           // 1 .. 2^30-1, different for every launch that uses status64
};


// Thread 0 only: decode ticket g (columns that skip this pass are stepped over with further tickets),
// publish it in s_tk and start the bulk copies of its tile into `stage` (completion on `full`).
__device__ __forceinline__ void tma_issue(const TmaArgs& a, uint32_t g, TmaTicket* s_tk, uint32_t stage, uint32_t full) {
  const int pass = a.p.pass;
  const uint32_t ntiles = (uint32_t)a.p.ntiles;
  TmaTicket tk;
  int src = 0;
  for (;;) {
    if (g >= a.total_tiles) {
      tk.col = tk.tile = tk.nvalid = 0;
      tk.mode_off = kModeEnd;
      *s_tk = tk;
      mbar_arrive(full);
      return;
    }
    tk.col = g / ntiles;
    tk.tile = g - tk.col * ntiles;
    if (a.p.plan[tk.col].run[pass]) break;
    g = atomicAdd(a.ticket, 1u);  // constant digit: the column sits this pass out
  }
  src = a.p.plan[tk.col].src[pass];
  tk.nvalid = min((uint32_t)kTmaTile, a.p.n - tk.tile * (uint32_t)kTmaTile);
  const size_t first = (size_t)tk.col * a.p.n + (size_t)tk.tile * kTmaTile;
  // the stage was last touched through the generic proxy (the previous occupant's staging); the barrier
  // that let this thread get here ordered those accesses, the fence hands the stage to the async proxy
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (src == 0) {
    const double* gp = a.p.raw + (int64_t)tk.col * a.p.col_stride + (int64_t)tk.tile * kTmaTile * a.p.row_stride;
    const bool bulk = a.p.row_stride == 1 && ((uintptr_t)gp & 15u) == 0 && (tk.nvalid & 1u) == 0;
    tk.mode_off = bulk ? kModeRawBulk : kModeRawDirect;
    *s_tk = tk;
    if (bulk) {
      mbar_arrive_expect_tx(full, tk.nvalid * 8u);
      bulk_g2s(stage, gp, tk.nvalid * 8u, full);
    } else {
      mbar_arrive(full);
    }
  } else {
    const uint64_t* gk = (src == 1 ? a.p.keysA : a.p.keysB) + first;
    const uint32_t* gv = (src == 1 ? a.p.valsA : a.p.valsB) + first;
    const uint32_t mk = (uint32_t)((uintptr_t)gk & 15u), mv = (uint32_t)((uintptr_t)gv & 15u);
    const uint32_t kb = (mk + tk.nvalid * 8u + 15u) & ~15u, vb = (mv + tk.nvalid * 4u + 15u) & ~15u;
    tk.mode_off = kModeBuffers | ((mk >> 3) << 8) | ((mv >> 2) << 16);
    *s_tk = tk;
    mbar_arrive_expect_tx(full, kb + vb);
    bulk_g2s(stage, reinterpret_cast<const unsigned char*>(gk) - mk, kb, full);
    bulk_g2s(stage + kTmaKeyBytes, reinterpret_cast<const unsigned char*>(gv) - mv, vb, full);
  }
}

struct TmaSmem {
  unsigned char* stage0;  // two stages, kTmaStageBytes apart
  uint32_t* hist;   // [8 warps][256]
  uint32_t* goff;   // [256]
  uint32_t* wsum;   // [8]
  TmaTicket* tk;    // [2]
  uint32_t full0;   // shared-space address of the two mbarriers (8 B apart)
  __device__ __forceinline__ unsigned char* stage(uint32_t s) const { return stage0 + s * kTmaStageBytes; }
  __device__ __forceinline__ uint32_t full(uint32_t s) const { return full0 + 8u * s; }
};

template <bool FULL>
__device__ __forceinline__ void tma_tile(const TmaArgs& a, const TmaSmem& sm, const uint32_t s, const TmaTicket tk,
                                         const KeyMap& map, const int dst, const uint32_t dshift,
                                         const uint32_t bin_base) {
  constexpr int ITEMS = kTmaItems;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t n = a.p.n;
  const uint32_t col = tk.col, tile = tk.tile, nvalid = tk.nvalid;
  const uint32_t mode = tk.mode_off & 255u, koff = (tk.mode_off >> 8) & 255u, voff = (tk.mode_off >> 16) & 255u;
  const uint32_t tile_start = tile * (uint32_t)kTmaTile;
  const uint32_t pos0 = warp * (ITEMS * 32) + lane;
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(sm.stage(s));
  uint32_t* s_vals = reinterpret_cast<uint32_t*>(sm.stage(s) + kTmaKeyBytes);

  // the ticket of the tile after this one: asked for now, needed after the first barrier
  uint32_t next_ticket = 0;
  if (tid == 0) next_ticket = atomicAdd(a.ticket, 1u);

  // the warp's private digit counters start at zero
  uint32_t* wh = sm.hist + warp * kRadix;
  {
    const uint4 z = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(wh)[lane] = z;
    reinterpret_cast<uint4*>(wh)[lane + 32] = z;
  }
  __syncwarp();

  // ---- keys and rows: stage -> registers ----
  uint64_t key[ITEMS];
  uint32_t negzero = 0;  // raw input: bit u set if item u was -0.0
  if (mode == kModeBuffers) {
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) key[u] = (FULL || pos0 + u * 32 < nvalid) ? s_keys[koff + pos0 + u * 32] : 0ull;
  } else {
    // first pass of a column: the caller's doubles (a bulk copy landed them in the key area, or -- strided /
    // misaligned input -- straight from global memory); the payload is the row index + the -0.0 flag
    const double* colp = a.p.raw + (int64_t)col * a.p.col_stride + (int64_t)(tile_start + pos0) * a.p.row_stride;
    const int64_t step = 32 * a.p.row_stride;
    const double* s_raw = reinterpret_cast<const double*>(sm.stage(s));
    double d[ITEMS];
    if (mode == kModeRawBulk) {
#pragma unroll
      for (int u = 0; u < ITEMS; ++u) d[u] = (FULL || pos0 + u * 32 < nvalid) ? s_raw[pos0 + u * 32] : 0.0;
    } else {
#pragma unroll
      for (int u = 0; u < ITEMS; ++u) d[u] = (FULL || pos0 + u * 32 < nvalid) ? ld_stream_f64(colp + u * step) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
      bool nz;
      key[u] = compact_key(key_of_double(d[u], &nz), map);
      negzero |= (nz ? 1u : 0u) << u;
    }
  }
  uint32_t dig[ITEMS];
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    dig[u] = (uint32_t)(key[u] >> dshift) & (uint32_t)(kRadix - 1);
    // padding of a partial tile goes to the last bin, behind every real key of the tile
    if (!FULL && pos0 + u * 32 >= nvalid) dig[u] = (uint32_t)(kRadix - 1);
  }

  uint32_t rank[ITEMS];
  warp_rank_digits<ITEMS>(dig, wh, rank, lane);
  __syncthreads();  // every key and row of the stage is in registers; all warps' counts are final
  // ... and every warp has finished writing out the PREVIOUS tile, whose stage can take the next one
  if (tid == 0) tma_issue(a, next_ticket, &sm.tk[s ^ 1u], smem_u32(sm.stage(s ^ 1u)), sm.full(s ^ 1u));

  // ---- per bin: the warps' counts -> tile total, published for the tiles behind us; exclusive scan over
  //      warps and bins; the counters become (bin start in the tile + keys of earlier warps in the bin) ----
  uint64_t* st = a.status64 + ((size_t)col * a.p.ntiles) * kRadix;
  const uint64_t tag = (uint64_t)a.epoch << 34;
  uint32_t cnt = 0;
#pragma unroll
  for (int w = 0; w < kTmaWarps; ++w) cnt += sm.hist[w * kRadix + tid];
  if (!FULL && tid == kRadix - 1) cnt -= (uint32_t)kTmaTile - nvalid;  // the padding keys
  st_relaxed_u64(&st[(size_t)tile * kRadix + tid], tag | (tile == 0 ? kStatusInclusive : kStatusPartial) | cnt);
  uint32_t incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
    if (lane >= (uint32_t)d) incl += t;
  }
  if (lane == 31) sm.wsum[warp] = incl;
  uint32_t bin_start = incl - cnt;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < kTmaWarps; ++w)
    if ((uint32_t)w < warp) bin_start += sm.wsum[w];
  {
    uint32_t run = bin_start;
#pragma unroll
    for (int w = 0; w < kTmaWarps; ++w) {
      const uint32_t c = sm.hist[w * kRadix + tid];
      sm.hist[w * kRadix + tid] = run;
      run += c;
    }
  }
  __syncthreads();  // offsets visible to everyone

  // ---- into digit order, in place: the keys now (every key of the stage has been in registers since the
  //      first barrier), the rows after one more barrier (they are only fetched here, to keep them out of
  //      the registers while the keys are ranked) ----
  uint32_t val[ITEMS];
  if (mode == kModeBuffers) {
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) val[u] = (FULL || pos0 + u * 32 < nvalid) ? s_vals[voff + pos0 + u * 32] : 0u;
  } else {
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) val[u] = (tile_start + pos0 + u * 32) | (((negzero >> u) & 1u) ? kNegZeroFlag : 0u);
  }
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    rank[u] += wh[dig[u]];
    s_keys[rank[u]] = key[u];
  }

  // ---- exclusive prefix of this bin over all earlier tiles of the column ----
  {
    uint32_t excl = 0;
    if (tile != 0) {
      constexpr int LB = 4;
      int64_t t = (int64_t)tile - 1;
      bool done = false;
      uint32_t spins = 0;
      while (!done) {
        uint64_t pre[LB];
#pragma unroll
        for (int i = 0; i < LB; ++i)
          pre[i] = (t - i >= 0) ? ld_relaxed_u64(&st[(size_t)(t - i) * kRadix + tid]) : (tag | kStatusInclusive);
#pragma unroll
        for (int i = 0; i < LB; ++i) {
          if (!done) {
            const uint64_t w = pre[i];
            if ((w >> 34) != (uint64_t)a.epoch || (w & (kStatusInclusive | kStatusPartial)) == 0) {  // not published yet
              if (++spins > kSpinLimit) {
                atomicExch(&a.p.error_flag[kFlagWatchdog], 1u);
                done = true;
              }
              break;
            }
            excl += (uint32_t)w;
            --t;
            if (w & kStatusInclusive) done = true;
          }
        }
      }
      st_relaxed_u64(&st[(size_t)tile * kRadix + tid], tag | kStatusInclusive | (uint32_t)(excl + cnt));
    }
    sm.goff[tid] = bin_base + excl - bin_start;  // + position in tile order = global slot
  }
  __syncthreads();  // every row has been fetched
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) s_vals[rank[u]] = val[u];
  __syncthreads();

  // ---- coalesced runs out to HBM ----
  uint64_t* out_keys = (dst == 1 ? a.p.keysA : a.p.keysB) + (size_t)col * n;
  uint32_t* out_vals = (dst == 1 ? a.p.valsA : a.p.valsB) + (size_t)col * n;
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const uint32_t pos = j * kTmaWorkers + tid;
    if (FULL || pos < nvalid) {
      const uint64_t k = s_keys[pos];
      const uint32_t d = (uint32_t)(k >> dshift) & (uint32_t)(kRadix - 1);
      const uint32_t g = sm.goff[d] + pos;
      st_u64_at(out_keys, g, k);
      st_u32_at(out_vals, g, s_vals[pos]);
    }
  }
  // this thread's accesses to the stage (generic proxy) are ordered before the bulk copy (async proxy) that
  // will refill it after the next block barrier
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__global__ void __launch_bounds__(kTmaThreads, 2) pass_tma_kernel(const TmaArgs a) {
  extern __shared__ __align__(128) unsigned char tsm[];
  TmaSmem sm;
  sm.stage0 = tsm;
  sm.hist = reinterpret_cast<uint32_t*>(tsm + 2 * kTmaStageBytes);
  sm.goff = sm.hist + kTmaWarps * kRadix;
  sm.wsum = sm.goff + kRadix;                               // [8] (+8 pad)
  sm.tk = reinterpret_cast<TmaTicket*>(sm.wsum + 16);       // [2]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(sm.tk + 2);  // full[2]
  sm.full0 = smem_u32(s_bar);
  const uint32_t tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(sm.full(0), 1);
    mbar_init(sm.full(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_issue(a, atomicAdd(a.ticket, 1u), &sm.tk[0], smem_u32(sm.stage(0)), sm.full(0));
  }
  __syncthreads();
  const int pass = a.p.pass;
  uint32_t cur_col = 0xFFFFFFFFu, dshift = 0, bin_base = 0;
  int dst = 1;
  KeyMap map;
  map.kmin = map.g0 = map.g = map.neg_al = 0;
  map.sh = 0;
  map.exact = true;
  for (uint32_t it = 0;; ++it) {
    const uint32_t s = it & 1u;
    mbar_wait(sm.full(s), (it >> 1) & 1u, a.p.error_flag);
    const TmaTicket tk = sm.tk[s];
    if ((tk.mode_off & 255u) == kModeEnd) break;
    if (tk.col != cur_col) {  // per-column constants (a CTA stays on a column for ~80 tiles)
      cur_col = tk.col;
      dst = (a.p.plan[cur_col].src[pass] == 1) ? 2 : 1;
      map = a.maps[cur_col];
      dshift = map.sh + (uint32_t)pass * kRadixBits;
      bin_base = a.p.bin_base_all[((size_t)cur_col * kMaxPasses + pass) * kRadix + tid];
    }
    if (tk.nvalid == (uint32_t)kTmaTile)
      tma_tile<true>(a, sm, s, tk, map, dst, dshift, bin_base);
    else
      tma_tile<false>(a, sm, s, tk, map, dst, dshift, bin_base);
  }
}
