// Stable in-warp ranking of 8-bit digits against warp-private counters (the multi-split step of a
// onesweep digit pass and of the row-window partition that follows each sort).
//
// A warp holds ITEMS rows of 32 keys (row u, lane l <-> position u*32 + l).  For every key the
// function returns  (keys of the warp with the same digit at earlier positions) + counter[digit]
// as it stood before the call, and leaves counter[digit] += (keys of the warp with that digit).
// With counters that start at zero the ranks are 0-based inside the warp and the counters end up
// as the warp's digit histogram, so no separate counting phase is needed.
//
//   (1) the mask of lanes holding the same digit, from 8 ballots (one per digit bit).  MATCH.ANY is
//       avoided on purpose: measured 62 cycles per warp-op per SM on random digits (tools/micro/
//       match_bench.cu), 8 ballots cost 25.  The ballots are written in PTX so that ptxas emits one
//       R2P (digit bits -> predicates) per key and VOTE + SEL + LOP3 per bit: 3 ALU instructions per
//       digit bit instead of the 6-7 nvcc makes of `mm &= bit ? bal : ~bal`.
//   (2) the highest lane of every digit group reads and bumps the group's counter -- plain load /
//       store: the counters are private to the warp and a warp barrier orders item u's store before
//       item u+1's load (a returning shared atomic costs ~2 cycles per active lane on B200);
//   (3) broadcast of the base from that lane + the number of group members in lower lanes.
#pragma once
#include "common.cuh"

namespace pbl {

__device__ __forceinline__ uint32_t match_digit8(uint32_t d) {
  uint32_t mm = 0xFFFFFFFFu;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    // lut 0x60 = a & (b ^ c): bal ^ x is the ballot where my bit is set, its complement where it is not
    asm volatile(  // volatile: a vote must stay where it is (inline asm is not marked convergent)
        "{\n\t.reg .pred p;\n\t.reg .b32 t, bal, x;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 bal, p, 0xffffffff;\n\t"
        "selp.b32 x, 0, 0xffffffff, p;\n\t"
        "lop3.b32 %0, %0, bal, x, 0x60;\n\t}"
        : "+r"(mm)
        : "r"(d), "r"(1u << b));
  }
  return mm;
}

template <int ITEMS>
__device__ __forceinline__ void warp_rank_digits(const uint32_t (&dig)[ITEMS], uint32_t* __restrict__ counters,
                                                 uint32_t (&rank)[ITEMS], const uint32_t lane) {
  const uint32_t lt = lanemask_lt();
  uint32_t m[ITEMS];
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) m[u] = match_digit8(dig[u]);
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    uint32_t base = 0;
    if ((m[u] >> lane) == 1u) {  // highest lane of its group
      uint32_t* ctr = counters + dig[u];
      base = *ctr;
      *ctr = base + __popc(m[u]);
    }
    __syncwarp();
    rank[u] = base;
  }
#pragma unroll
  for (int u = 0; u < ITEMS; ++u)
    rank[u] = __shfl_sync(0xFFFFFFFFu, rank[u], 31 - __clz(m[u])) + __popc(m[u] & lt);
}

// Same, with positions that hold no key (the tail of a partial tile): they take part in the votes but
// match nobody, get no slot and leave the counters alone (their returned rank is meaningless).
template <int ITEMS>
__device__ __forceinline__ void warp_rank_digits_masked(const uint32_t (&dig)[ITEMS], const bool (&valid)[ITEMS],
                                                        uint32_t* __restrict__ counters, uint32_t (&rank)[ITEMS],
                                                        const uint32_t lane) {
  const uint32_t lt = lanemask_lt();
  uint32_t m[ITEMS];
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    const uint32_t mm = match_digit8(dig[u]);
    const uint32_t vb = __ballot_sync(0xFFFFFFFFu, valid[u]);
    m[u] = valid[u] ? (mm & vb) : 0u;
  }
#pragma unroll
  for (int u = 0; u < ITEMS; ++u) {
    uint32_t base = 0;
    if ((m[u] >> lane) == 1u) {
      uint32_t* ctr = counters + dig[u];
      base = *ctr;
      *ctr = base + __popc(m[u]);
    }
    __syncwarp();
    rank[u] = base;
  }
#pragma unroll
  for (int u = 0; u < ITEMS; ++u)
    rank[u] = __shfl_sync(0xFFFFFFFFu, rank[u], (31 - __clz(m[u])) & 31) + __popc(m[u] & lt);
}

}  // namespace pbl
