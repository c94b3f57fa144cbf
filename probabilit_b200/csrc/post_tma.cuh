// The consumer of the windowed sort as a PERSISTENT kernel with double-buffered bulk-async (TMA) tile
// loads (same pipeline as the digit pass, tile_pipeline.cuh).  Included by ic.cu inside namespace pbl.
//
// Per column the sort leaves (compact key, row) pairs ordered by the 32-bit WINDOW value of the key
// (sort.cuh).  A tile of 4096 positions (+ 64 positions of halo on either side) is pulled into a
// shared-memory stage by cp.async.bulk; then, with every thread owning 16 warp-striped positions:
//  (1) ORDER COMPLETION.  A position whose window value equals a neighbour's is a "member" of a run
//      (a fifth of the positions at 0.2 keys per window value; everybody else does nothing).  A member
//      walks its run (<= kMaxRun keys, else the run is left as it is, which is fine if it is pure --
//      a tie run -- and raises kFlagWindowRetry otherwise), counts the keys that sort before it, and
//      posts its own slot at its destination; after a barrier every member slot PULLS the key and
//      row that belong there.  Halo slots do the same, so that the tile's neighbours are final too.
//  (2) TIE RUNS (only if the tile has two equal keys): head / tail flags of the runs as ballot words in
//      shared memory, start and end of every position's run by bit scans, runs that leave the tile
//      by binary search on the (monotone) window values of the column.
//  (3) the value each source row must receive:
//      MODE 0  ndtri(average_rank / (n+1)) -- vdw[p] when untied -- and sortedX[p] = the p-th smallest
//              input  (scipy.stats.rankdata + norm.ppf, reference correlation.py:394-395; np.sort, :423)
//      MODE 1  sortedX[run_start + (run_len-1)/2]  (rankdata(...).astype(int) - 1 + gather, :422-423)
//      MODE 2  the average rank itself (Spearman mode of CorrelationMatrix, :835-837)
//  (4) first half of the scatter by row: the tile's (row, value) pairs are grouped into <= 256 row
//      windows of L2 size by the shared multi-split (split_tile) and written out as coalesced runs;
//      scatter_rows_kernel then delivers value -> row inside L2.
#pragma once

constexpr int kPHalo = 64;  // >= 2 kMaxRun - 1: every member of a completable run sees the whole run
constexpr uint32_t kPKeyBytes = (kTile + 2 * kPHalo + 2) * 8;
constexpr uint32_t kPRowBytes = (kTile + 2 * kPHalo + 4) * 4;
constexpr uint32_t kPStageBytes = kPKeyBytes + kPRowBytes;
static_assert(kPKeyBytes % 16 == 0 && kPRowBytes % 16 == 0, "stage members are 16 B units");
static_assert(kPHalo >= 2 * kMaxRun - 1, "halo covers two runs");
constexpr int kPMaskWords = kTile / 32;
constexpr size_t kPostTmaSmemBytes =
    2 * (size_t)kPStageBytes + kSplitSmemBytes + 2 * kPHalo * 2 + 2 * kPMaskWords * 4 + 16 + 2 * 48 + 2 * 8;

struct __align__(16) PostTicket {
  uint32_t col, tile, nvalid;
  uint32_t kslot0, vslot0;  // stage slots (keys / rows) of the tile's first position
  int32_t qlo, qhi;         // tile-local positions present in the stage: [qlo, qhi)
  uint32_t end;             // 1: no more tiles
  uint32_t sh;              // window value = compact key >> sh
  uint32_t fb;              // ping-pong buffer that holds the sorted column (1 = A, 2 = B)
  uint32_t pad[2];
};

struct PostArgs {
  uint64_t* keysA;
  uint64_t* keysB;
  uint32_t* valsA;
  uint32_t* valsB;
  const PassPlan* plan;
  const KeyMap* maps;
  double* sortedX;        // [ncols][n] of this batch
  const double* vdw;      // [n]
  uint32_t* flags;
  uint64_t* status64;     // [ncols][ntiles][256]
  uint32_t* ticket;
  uint32_t n;
  uint32_t ntiles;
  uint32_t total_tiles;   // ncols * ntiles
  uint32_t epoch;
  int col_base;           // global index of the batch's first column
  int part_shift;         // row >> part_shift = row window; 32: no grouping (short columns)
  uint32_t ncols;
  uint32_t ncols_interleave;  // ticket order: width of the interleaved column groups (ticket_to_tile)
};

struct PostSmem {
  unsigned char* stage0;
  SplitSmem split;
  uint16_t* src_main;   // [kTile] destination slot -> source slot (aliases split.hist: idle until the multi-split)
  uint16_t* src_ext;    // [2 * kPHalo] the same for destinations beyond the first kTile stage positions
  uint32_t* head;       // [kPMaskWords] ballot words: position starts a tie run
  uint32_t* tail;       // [kPMaskWords] ... ends a tie run
  uint32_t* lohi;       // [2] start / end of the tie runs that leave the tile
  PostTicket* tk;       // [2]
  uint32_t full0;
  __device__ __forceinline__ unsigned char* stage(uint32_t s) const { return stage0 + s * kPStageBytes; }
  __device__ __forceinline__ uint32_t full(uint32_t s) const { return full0 + 8u * s; }
  __device__ __forceinline__ uint16_t& src(uint32_t i) const { return i < (uint32_t)kTile ? src_main[i] : src_ext[i - kTile]; }
};

// Thread 0 only: decode ticket g, publish it and start the bulk copies of its tile (+ halo).
__device__ __forceinline__ void post_issue(const PostArgs& a, uint32_t g, PostTicket* s_tk, uint32_t stage, uint32_t full) {
  PostTicket tk;
  tk.end = g >= a.total_tiles ? 1u : 0u;
  if (tk.end) {
    tk.col = tk.tile = tk.nvalid = tk.kslot0 = tk.vslot0 = tk.sh = tk.fb = tk.pad[0] = tk.pad[1] = 0;
    tk.qlo = tk.qhi = 0;
    *s_tk = tk;
    mbar_arrive(full);
    return;
  }
  ticket_to_tile(g, a.ncols, a.ntiles, a.ncols_interleave, &tk.col, &tk.tile);
  tk.sh = a.maps[tk.col].sh;
  tk.pad[0] = tk.pad[1] = 0;
  const uint32_t tile_start = tk.tile * (uint32_t)kTile;
  tk.nvalid = min((uint32_t)kTile, a.n - tile_start);
  tk.qlo = -(int32_t)min((uint32_t)kPHalo, tile_start);
  tk.qhi = (int32_t)min((uint32_t)(kTile + kPHalo), a.n - tile_start);
  const int fb = a.plan[tk.col].final_buf;
  tk.fb = (uint32_t)fb;
  const size_t first = (size_t)tk.col * a.n + tile_start + tk.qlo;
  const uint64_t* gk = (fb == 1 ? a.keysA : a.keysB) + first;
  const uint32_t* gv = (fb == 1 ? a.valsA : a.valsB) + first;
  const uint32_t cnt = (uint32_t)(tk.qhi - tk.qlo);
  const uint32_t mk = (uint32_t)((uintptr_t)gk & 15u), mv = (uint32_t)((uintptr_t)gv & 15u);
  const uint32_t kb = (mk + cnt * 8u + 15u) & ~15u, vb = (mv + cnt * 4u + 15u) & ~15u;
  tk.kslot0 = (mk >> 3) + (uint32_t)(-tk.qlo);
  tk.vslot0 = (mv >> 2) + (uint32_t)(-tk.qlo);
  *s_tk = tk;
  fence_proxy_async();
  mbar_arrive_expect_tx(full, kb + vb);
  bulk_g2s(stage, reinterpret_cast<const unsigned char*>(gk) - mk, kb, full);
  bulk_g2s(stage + kPKeyBytes, reinterpret_cast<const unsigned char*>(gv) - mv, vb, full);
}

// ndtri out of line: the tie path would otherwise carry 16 inlined copies
__device__ __noinline__ double ndtri_call(double q) { return ndtri(q); }

template <int MODE, bool FULL>
__device__ __forceinline__ void post_tile(const PostArgs& a, const PostSmem& sm, const uint32_t s, const PostTicket tk) {
  constexpr int ITEMS = kTileItems;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t n = a.n;
  const uint32_t col = tk.col, tile = tk.tile, nvalid = tk.nvalid;
  const int fb = (int)tk.fb;
  // only the shift of the window map is needed up front; the rest (MODE 0: decoding the keys) is asked for now
  KeyMap map = a.maps[col];
  map.sh = tk.sh;
  const uint32_t tile_start = tile * (uint32_t)kTile;
  const uint32_t pos0 = warp * (ITEMS * 32) + lane;
  const int qlo = tk.qlo, qhi = tk.qhi;
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(sm.stage(s));
  uint32_t* s_rows = reinterpret_cast<uint32_t*>(sm.stage(s) + kPKeyBytes);
  const uint64_t* kq = s_keys + tk.kslot0;  // indexed by tile-local position q in [qlo, qhi)
  const uint32_t* vq = s_rows + tk.vslot0;
  uint64_t* kqw = s_keys + tk.kslot0;
  uint32_t* vqw = s_rows + tk.vslot0;
  double* sx = a.sortedX + (size_t)col * n;
  const uint64_t* gkeys = (fb == 1 ? a.keysA : a.keysB) + (size_t)col * n;  // the column as the sort left it

  uint32_t next_ticket = 0;
  if (tid == 0) next_ticket = atomicAdd(a.ticket, 1u);

  // the position-indexed operand of step (3), asked for now so that its HBM latency is off the tile's
  // dependency chain: the van der Waerden score of the position (MODE 0) / the sorted marginal at the
  // position (MODE 1; used unless the position sits in a tie run, whose midpoint is fetched instead)
  double pre_val[ITEMS];
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    const uint32_t q = pos0 + u * 32;
    pre_val[u] = 0.0;
    if (MODE != 2 && (FULL || q < nvalid)) pre_val[u] = ld_stream_f64((MODE == 0 ? a.vdw : sx) + tile_start + q);
  }

  // ---- (1) order completion inside runs of equal window values ----
  // The column is sorted by window value, so the keys that share position q's window are exactly the
  // positions q - L .. q + R around it, and q's place among them is
  //     q - #{left neighbours of the run with a LARGER key} + #{right neighbours with a SMALLER key}
  // (equal keys keep their order).  A first, branch-free step looks at distance 1 and finds the MEMBERS of
  // runs (a fifth of the positions at 0.2 keys per window value, most of them at 1.5: columns of 8e8 rows);
  // a member then looks at distance d = 1, 2, ... on both sides at once until the window value changes --
  // ONE pass with two independent loads per step (it used to find the run's ends first and scan the run
  // again: 2 (L + R) + 3 dependent loads).
  uint64_t key[ITEMS];
  uint32_t members = 0;  // bit u: my position u shares its window value with a neighbour
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    const int q = (int)(pos0 + u * 32);
    key[u] = 0ull;
    if (FULL || (uint32_t)q < nvalid) {
      key[u] = kq[q];
      const bool fl = q - 1 >= qlo && same_window(key[u], kq[q - 1], map);
      const bool fr = q + 1 < qhi && same_window(key[u], kq[q + 1], map);
      members |= ((fl | fr) ? 1u : 0u) << u;
    }
  }
  // a halo slot for the first 128 threads (a run that straddles the tile's edge is completed by both tiles)
  int hq = 0x7FFFFFFF;
  bool hmember = false;
  if (tid < 2 * kPHalo) {
    const int q = tid < kPHalo ? (int)tid - kPHalo : (int)kTile + ((int)tid - kPHalo);
    if (q >= qlo && q < qhi && (q < 0 || q >= (int)nvalid)) {
      hq = q;
      const uint64_t my = kq[q];
      hmember = (q - 1 >= qlo && same_window(my, kq[q - 1], map)) || (q + 1 < qhi && same_window(my, kq[q + 1], map));
    }
  }
  int tie = 0, retry = 0;
  // destination of the member at q; posts "destination <- q" for the pull after the barrier
  auto resolve = [&](const int q) {
    const uint64_t my = kq[q];
    bool gl = q - 1 >= qlo, gr = q + 1 < qhi;
    int L = 0, R = 0, adj = 0, eq = 0;
    bool left_differs = false;
    for (int d = 1; d <= kMaxRun && (gl || gr); ++d) {
      if (gl) {
        const uint64_t l = kq[q - d];
        gl = same_window(l, my, map);
        if (d == 1) left_differs = l != my;
        if (gl) {
          ++L;
          adj -= l > my ? 1 : 0;
          eq |= l == my ? 1 : 0;
          gl = q - d - 1 >= qlo;
        }
      }
      if (gr) {
        const uint64_t r = kq[q + d];
        gr = same_window(r, my, map);
        if (gr) {
          ++R;
          adj += r < my ? 1 : 0;
          eq |= r == my ? 1 : 0;
          gr = q + d + 1 < qhi;
        }
      }
    }
    int dst = q;
    if (L + R + 1 <= kMaxRun) {
      dst = q + adj;
      tie |= eq;
    } else if (L > 0) {
      // a long run is left as it is: fine if it is pure (a tie run); any adjacent pair of DIFFERENT keys
      // inside it raises the retry flag in the tile that owns either key
      tie |= left_differs ? 0 : 1;
      if (left_differs && q >= 0 && q <= (int)nvalid) retry = 1;
    }
    sm.src((uint32_t)(dst - qlo)) = (uint16_t)(q - qlo);
  };
  {
    uint32_t mm = members;
    while (mm) {
      const int u = __ffs((int)mm) - 1;
      mm &= mm - 1;
      resolve((int)(pos0 + u * 32));
    }
    if (hmember) resolve(hq);
  }
  if (retry) a.flags[kFlagWindowRetry] = 1u;
  tie = __syncthreads_or(tie);
  // every warp is past the previous tile: its stage takes the tile after this one
  if (tid == 0) post_issue(a, next_ticket, &sm.tk[s ^ 1u], smem_u32(sm.stage(s ^ 1u)), sm.full(s ^ 1u));

  // pull: the key and row that belong at my member positions (rows of the other positions: as they are)
  uint32_t row[ITEMS];
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    const int q = (int)(pos0 + u * 32);
    row[u] = 0u;
    if (FULL || (uint32_t)q < nvalid) {
      int from = q;
      if ((members >> u) & 1u) {
        from = (int)sm.src((uint32_t)(q - qlo)) + qlo;
        key[u] = kq[from];
      }
      row[u] = vq[from];
    }
  }
  uint64_t hkey = 0;
  uint32_t hrow = 0;
  if (hmember) {
    const int from = (int)sm.src((uint32_t)(hq - qlo)) + qlo;
    hkey = kq[from];
    hrow = vq[from];
  }
  if (__syncthreads_or(members != 0u || hmember)) {  // (block-uniform) somebody moves: put the pulled pairs in place
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
      if ((members >> u) & 1u) {
        const int q = (int)(pos0 + u * 32);
        kqw[q] = key[u];
        vqw[q] = row[u];
      }
    }
    if (hmember) {
      kqw[hq] = hkey;
      vqw[hq] = hrow;
    }
    __syncthreads();
  }

  // ---- (2) + (3) tie runs and the value every source row must receive ----
  uint64_t big[ITEMS];
  if (tie) {
    // teq(q): position q holds the same value as position q - 1 (keys are canonical: -0.0 folded onto +0.0)
    auto teq = [&](const int q) -> bool {
      const int64_t g = (int64_t)tile_start + q;
      if (g <= 0 || g >= (int64_t)n) return false;
      return kq[q] == kq[q - 1];
    };
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
      const int q = (int)(pos0 + u * 32);
      const bool ok = FULL || (uint32_t)q < nvalid;
      const uint32_t hb = __ballot_sync(0xFFFFFFFFu, ok && !teq(q));
      const uint32_t tb = __ballot_sync(0xFFFFFFFFu, ok && !teq(q + 1));
      if (lane == 0) {
        sm.head[warp * ITEMS + u] = hb;
        sm.tail[warp * ITEMS + u] = tb;
      }
    }
    if (tid == 0) {
      // tie runs that cross the tile boundary: binary search on the window value in the column (valid
      // because such a run is pure: any impure long run has raised kFlagWindowRetry)
      uint32_t lo = tile_start, hi = tile_start + nvalid - 1;
      if (teq(0)) {
        const uint64_t w = window_value(kq[0], map);
        uint32_t x = 0, y = tile_start;  // first position in [0, tile_start) with window >= w
        while (x < y) {
          const uint32_t mid = x + (y - x) / 2;
          if (window_value(gkeys[mid], map) < w) x = mid + 1; else y = mid;
        }
        lo = x;
      }
      if (teq((int)nvalid)) {
        const uint64_t w = window_value(kq[(int)nvalid - 1], map);
        uint32_t x = tile_start + nvalid, y = n;  // first position with window > w
        while (x < y) {
          const uint32_t mid = x + (y - x) / 2;
          if (window_value(gkeys[mid], map) > w) y = mid; else x = mid + 1;
        }
        hi = x - 1;
      }
      sm.lohi[0] = lo;
      sm.lohi[1] = hi;
    }
    __syncthreads();
    const int nwords = (int)((nvalid + 31) / 32);
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
      const uint32_t q = pos0 + u * 32;
      big[u] = 0ull;
      if (FULL || q < nvalid) {
        const uint32_t g = tile_start + q;
        int c = (int)(warp * ITEMS + u);
        uint32_t m = sm.head[c] & (0xFFFFFFFFu >> (31 - lane));
        while (m == 0 && c > 0) m = sm.head[--c];
        const uint32_t st = m ? tile_start + (uint32_t)c * 32u + (31u - (uint32_t)__clz(m)) : sm.lohi[0];
        c = (int)(warp * ITEMS + u);
        m = sm.tail[c] & (0xFFFFFFFFu << lane);
        while (m == 0 && c + 1 < nwords) m = sm.tail[++c];
        const uint32_t en = m ? tile_start + (uint32_t)c * 32u + (uint32_t)(__ffs((int)m) - 1) : sm.lohi[1];
        double v;
        if (MODE == 2) {
          v = (double)((uint64_t)st + (uint64_t)en + 2ull) * 0.5;
        } else if (MODE == 0) {
          v = pre_val[u];  // untied: the score depends on the position only
          if (st != en) v = ndtri_call(__ddiv_rn((double)((uint64_t)st + (uint64_t)en + 2ull) * 0.5, (double)((uint64_t)n + 1ull)));
        } else {
          const uint32_t mid = st + (en - st) / 2;
          v = (mid == g) ? pre_val[u] : sx[mid];
        }
        big[u] = (uint64_t)__double_as_longlong(v);
      }
    }
  } else {
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
      const uint32_t q = pos0 + u * 32;
      double v = pre_val[u];
      if (MODE == 2) v = (double)((uint64_t)(tile_start + q) + 1ull);
      big[u] = (uint64_t)__double_as_longlong(v);
    }
  }
  if (MODE == 0) {
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
      const uint32_t q = pos0 + u * 32;
      if (FULL || q < nvalid) {
        const bool nz = (row[u] & kNegZeroFlag) != 0u;
        sx[tile_start + q] = nz ? -0.0 : key_to_double(expand_key(key[u], map));
        if (nz && (int)col + a.col_base == 0) a.flags[kFlagNegZeroCol0] = 1u;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) row[u] &= kRowMask;

  // ---- (4) (row, value) pairs out, grouped by row window ----
  uint64_t* out_val = (fb == 1 ? a.keysB : a.keysA) + (size_t)col * n;
  uint32_t* out_row = (fb == 1 ? a.valsB : a.valsA) + (size_t)col * n;
  if (a.part_shift >= 32) {  // short columns: scatter_rows_kernel delivers straight from sorted order
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) {
      const uint32_t q = pos0 + u * 32;
      if (FULL || q < nvalid) {
        out_row[tile_start + q] = row[u];
        out_val[tile_start + q] = big[u];
      }
    }
    __syncthreads();  // all reads of the stage are complete before the next tile's prefetch may refill it
    fence_proxy_async();
    return;
  }
  const uint32_t shift = (uint32_t)a.part_shift;
  uint32_t dig[ITEMS];
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    dig[u] = row[u] >> shift;
    if (!FULL && pos0 + u * 32 >= nvalid) dig[u] = (uint32_t)(kRadix - 1);
  }
  const uint64_t b = (uint64_t)tid << shift;  // rows are a permutation of 0..n-1: the windows' bases are known
  const uint32_t bin_base = (uint32_t)(b < n ? b : n);
  auto load_rows = [&](uint32_t (&small)[ITEMS]) {
#pragma unroll
    for (int u = 0; u < ITEMS; ++u) small[u] = row[u];
  };
  auto digit_at = [&](uint32_t pos) { return s_rows[pos] >> shift; };
  auto nothing = [] {};
  split_tile<FULL>(dig, big, load_rows, digit_at, nothing, sm.split, s_keys, s_rows,
                   a.status64 + ((size_t)col * a.ntiles) * kRadix, tile, nvalid, a.epoch, bin_base, out_val, out_row,
                   a.flags);
}

template <int MODE>
__global__ void __launch_bounds__(kTileThreads, 2) post_tma_kernel(const PostArgs a) {
  extern __shared__ __align__(128) unsigned char psm2[];
  PostSmem sm;
  sm.stage0 = psm2;
  sm.split.hist = reinterpret_cast<uint32_t*>(psm2 + 2 * kPStageBytes);
  sm.split.goff = sm.split.hist + kTileWarps * kRadix;
  sm.split.wsum = sm.split.goff + kRadix;  // [8] (+8 pad)
  sm.src_main = reinterpret_cast<uint16_t*>(sm.split.hist);
  sm.src_ext = reinterpret_cast<uint16_t*>(sm.split.wsum + 16);
  sm.head = reinterpret_cast<uint32_t*>(sm.src_ext + 2 * kPHalo);
  sm.tail = sm.head + kPMaskWords;
  sm.lohi = sm.tail + kPMaskWords;                             // [2] (+2 pad)
  sm.tk = reinterpret_cast<PostTicket*>(sm.lohi + 4);          // [2]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(sm.tk + 2);
  sm.full0 = smem_u32(s_bar);
  const uint32_t tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(sm.full(0), 1);
    mbar_init(sm.full(1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    post_issue(a, atomicAdd(a.ticket, 1u), &sm.tk[0], smem_u32(sm.stage(0)), sm.full(0));
  }
  __syncthreads();
  for (uint32_t it = 0;; ++it) {
    const uint32_t s = it & 1u;
    mbar_wait(sm.full(s), (it >> 1) & 1u, a.flags);
    const PostTicket tk = sm.tk[s];
    if (tk.end) break;
    if (tk.nvalid == (uint32_t)kTile)
      post_tile<MODE, true>(a, sm, s, tk);
    else
      post_tile<MODE, false>(a, sm, s, tk);
  }
}
