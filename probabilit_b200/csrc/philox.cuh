// Philox4x32-10 counter-based generator shared by the unit-cube kernels (qmc.cu) and the in-kernel
// uniform source of the graph evaluator (graph.cu): the value of (seed, global row, column) is the
// same whichever kernel produces it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pbl {

struct U4 {
  uint32_t x, y, z, w;
};
__device__ __forceinline__ U4 philox4x32_10(U4 ctr, uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    U4 n;
    n.x = hi1 ^ ctr.y ^ k0;
    n.y = lo1;
    n.z = hi0 ^ ctr.w ^ k1;
    n.w = lo0;
    ctr = n;
    k0 += W0;
    k1 += W1;
  }
  return ctr;
}
// 53-bit uniform in [0, 1) from two 32-bit words (the construction NumPy uses for its doubles)
__device__ __forceinline__ double u01_53(uint32_t a, uint32_t b) {
  return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}


// the uniform of global row g, column c of stream `seed` (pbl_uniform_f64 and PBL_OP_UNIFORM):
// one Philox block serves the row pair (g & ~1, g | 1)
__device__ __forceinline__ double philox_uniform_at(uint64_t seed, uint64_t g, uint32_t c) {
  U4 ctr = {(uint32_t)(g >> 1), (uint32_t)(g >> 33), c, 0x50424C31u};
  U4 o = philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
  return (g & 1) ? u01_53(o.z, o.w) : u01_53(o.x, o.y);
}

}  // namespace pbl
