// Unit-cube generators: (n, d) fp64 in [0, 1), written as N x d tiles with coalesced (and, for
// unit row stride, 128-bit) stores.  They replace the generator calls of the reference's
// Node.sample (src/probabilit/modeling.py:479-489):
//   method None      random_state.random((n, d))          -> philox_uniform_kernel (statistical parity)
//   method "sobol"   scipy.stats.qmc.Sobol(d, rng).random  -> sobol_* kernels        (bit-exact)
//   method "halton"  scipy.stats.qmc.Halton(d, rng).random -> halton_kernel         (bit-exact)
//   method "lhs"     scipy.stats.qmc.LatinHypercube        -> lhs_kernel            (statistical parity)
// Every point is a pure function of its row index, so row shards on several GPUs need no
// communication (skip-ahead = start the index at the shard's first row).
#include <type_traits>

#include "../../include/probabilit_b200.h"
#include "common.cuh"
#include "philox.cuh"

namespace pbl {
namespace {

// one thread: rows (2t, 2t+1) of one column
__global__ void __launch_bounds__(256)
philox_uniform_kernel(uint64_t seed, uint64_t row0, int64_t n, int32_t d, double* __restrict__ out,
                      int64_t row_stride, int64_t col_stride) {
  const int c = blockIdx.y;
  const int64_t pair = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t r = pair * 2;
  if (r >= n) return;
  const uint64_t g = row0 + (uint64_t)r;  // global row of the first of the two
  U4 ctr = {(uint32_t)(g >> 1), (uint32_t)(g >> 33), (uint32_t)c, 0x50424C31u};
  U4 o = philox4x32_10(ctr, (uint32_t)seed, (uint32_t)(seed >> 32));
  // the counter is per *pair of global rows*: shard boundaries must be even (the host checks)
  double a = u01_53(o.x, o.y), b = u01_53(o.z, o.w);
  double* p = out + (int64_t)c * col_stride + r * row_stride;
  if (row_stride == 1 && r + 1 < n && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    *reinterpret_cast<double2*>(p) = make_double2(a, b);
  } else {
    p[0] = a;
    if (r + 1 < n) p[row_stride] = b;
  }
}

// ------------------------------------------------------------------------------ Sobol'
// scipy.stats._sobol._initialize_v: Joe-Kuo direction numbers, one thread per dimension.
__global__ void sobol_direction_kernel(const int64_t* __restrict__ poly, const int64_t* __restrict__ vinit,
                                       int32_t vinit_cols, int32_t d, int32_t bits,
                                       uint64_t* __restrict__ sv) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= d) return;
  uint64_t v[64];
  if (i == 0) {
    for (int j = 0; j < bits; ++j) v[j] = 1;
  } else {
    const uint64_t p = (uint64_t)poly[i];
    const int m = 63 - __clzll((long long)p);
    for (int j = 0; j < bits; ++j) v[j] = 0;
    for (int j = 0; j < m && j < bits; ++j) v[j] = (uint64_t)vinit[(size_t)i * vinit_cols + j];
    for (int j = m; j < bits; ++j) {
      uint64_t nv = v[j - m];
      for (int k = 0; k < m; ++k)
        if ((p >> (m - 1 - k)) & 1ull) nv ^= (2ull << k) * v[j - k - 1];
      v[j] = nv;
    }
  }
  for (int j = 0; j < bits; ++j) sv[(size_t)i * bits + j] = v[j] << (bits - 1 - j);
}

// LMS scramble (scipy.stats._sobol._cscramble): sv[i][j] <- L_i * sv[i][j] over GF(2), bit
// vectors MSB first, L_i lower triangular with unit diagonal built from the host's random bits;
// shift[i] = sum_b shift_bits[i][b] << b.  One thread per (dimension, column j).
__global__ void sobol_scramble_kernel(const uint8_t* __restrict__ ltm_bits, const uint8_t* __restrict__ shift_bits,
                                      int32_t d, int32_t bits, uint64_t* __restrict__ sv,
                                      uint64_t* __restrict__ shift) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d * bits) return;
  const int i = t / bits, j = t % bits;
  const uint64_t vdj = sv[(size_t)i * bits + j];
  uint64_t t2 = 0;
  for (int p = 0; p < bits; ++p) {
    uint64_t row = 0;
    for (int k = 0; k <= p; ++k) {
      uint64_t bit = (k == p) ? 1ull : (uint64_t)(ltm_bits[((size_t)i * bits + p) * bits + k] & 1);
      row |= bit << (bits - 1 - k);
    }
    t2 |= (uint64_t)(__popcll(row & vdj) & 1) << (bits - 1 - p);
  }
  __syncthreads();  // all reads of sv by this block are done (blocks never share a (i, j))
  sv[(size_t)i * bits + j] = t2;
  if (j == 0) {
    uint64_t s = 0;
    for (int b = 0; b < bits; ++b) s |= (uint64_t)(shift_bits[(size_t)i * bits + b] & 1) << b;
    shift[i] = s;
  }
}

// point j = (shift ^ XOR_{b set in gray(j)} sv[b]) * 2^-bits.  One thread: the 8 consecutive points of an
// aligned group j0 = 8 g .. 8 g + 7 of one dimension.  gray(j0 + t) = gray(j0) ^ gray(t) for t < 8 (j0 has
// its three low bits clear), so the group shares B = shift ^ XOR_{b in gray(j0)} sv[b] -- one walk over the
// set bits per 8 points instead of per point -- and point t is B ^ C[t] with the 8 combinations C[t] of
// sv[0..2] staged once per block.  64 contiguous bytes per thread go out as four 128-bit stores.
// NARROW: bits <= 32, the whole computation in 32-bit registers.
template <bool NARROW>
__global__ void __launch_bounds__(256)
sobol_points_kernel(const uint64_t* __restrict__ sv, const uint64_t* __restrict__ shift, int32_t bits,
                    uint64_t skip, int64_t n, double* __restrict__ out, int64_t row_stride,
                    int64_t col_stride) {
  using W = typename std::conditional<NARROW, uint32_t, uint64_t>::type;
  __shared__ W s_sv[64];
  __shared__ W s_c[8];
  const int c = blockIdx.y;
  if (threadIdx.x < 64) s_sv[threadIdx.x] = threadIdx.x < bits ? (W)sv[(size_t)c * bits + threadIdx.x] : (W)0;
  __syncthreads();
  if (threadIdx.x < 8) {
    const uint32_t g = threadIdx.x ^ (threadIdx.x >> 1);
    s_c[threadIdx.x] = ((g & 1u) ? s_sv[0] : (W)0) ^ ((g & 2u) ? s_sv[1] : (W)0) ^ ((g & 4u) ? s_sv[2] : (W)0);
  }
  __syncthreads();
  // groups are aligned to multiples of 8 of the GLOBAL index; the first / last group of a shard may be partial
  const uint64_t group = (skip >> 3) + (uint64_t)blockIdx.x * 256 + threadIdx.x;
  const uint64_t j0 = group << 3;
  if (j0 >= skip + (uint64_t)n) return;
  uint64_t g = j0 ^ (j0 >> 1);
  W base = (W)shift[c];
  while (g) {
    const int b = __ffsll((long long)g) - 1;
    base ^= s_sv[b];
    g &= g - 1;
  }
  const double scale = __longlong_as_double((long long)(1023 - bits) << 52);  // 2^-bits
  double v[8];
#pragma unroll
  for (int t = 0; t < 8; ++t) {
    const W q = base ^ s_c[t];
    v[t] = (NARROW ? __uint2double_rn((uint32_t)q) : __ull2double_rn((uint64_t)q)) * scale;
  }
  const int64_t r0 = (int64_t)(j0 - skip);  // negative for the head of a partial first group
  double* colp = out + (int64_t)c * col_stride;
  if (r0 >= 0 && r0 + 8 <= n && row_stride == 1 && ((reinterpret_cast<uintptr_t>(colp + r0) & 15) == 0)) {
    double2* p = reinterpret_cast<double2*>(colp + r0);
#pragma unroll
    for (int t = 0; t < 4; ++t) p[t] = make_double2(v[2 * t], v[2 * t + 1]);
  } else {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      if (r0 + t >= 0 && r0 + t < n) colp[(r0 + t) * row_stride] = v[t];
  }
}

// ------------------------------------------------------------------------------ Halton
// scipy.stats._qmc_cy._cy_van_der_corput(_scrambled): radical inverse in base b with the same
// floating-point operation order (no FMA contraction), optional per-digit permutations.
__global__ void __launch_bounds__(256)
halton_kernel(const int32_t* __restrict__ bases, const int64_t* __restrict__ perms,
              const int64_t* __restrict__ perm_off, const int32_t* __restrict__ perm_count,
              uint64_t start_index, int64_t n, double* __restrict__ out, int64_t row_stride,
              int64_t col_stride) {
  const int c = blockIdx.y;
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  const uint32_t base = (uint32_t)bases[c];
  uint64_t quotient = start_index + (uint64_t)r;
  const double fb = (double)base;
  double b2r = __ddiv_rn(1.0, fb);
  double acc = 0.0;
  if (perms == nullptr) {
    while (__dadd_rn(1.0, -b2r) < 1.0) {
      uint64_t qn = quotient / base;
      uint32_t rem = (uint32_t)(quotient - qn * base);
      acc = __dadd_rn(acc, __dmul_rn((double)rem, b2r));
      b2r = __ddiv_rn(b2r, fb);
      quotient = qn;
    }
  } else {
    const int64_t* pc = perms + perm_off[c];
    const int count = perm_count[c];
    for (int j = 0; j < count; ++j) {
      uint64_t qn = quotient / base;
      uint32_t rem = (uint32_t)(quotient - qn * base);
      acc = __dadd_rn(acc, __dmul_rn((double)pc[(size_t)j * base + rem], b2r));
      b2r = __ddiv_rn(b2r, fb);
      quotient = qn;
    }
  }
  out[(int64_t)c * col_stride + r * row_stride] = acc;
}

// ------------------------------------------------------------------------------ Latin hypercube
// (perm_c(i) + 1 - u) / n  (scipy/stats/_qmc.py:1546-1559): per-column pseudo-random permutation
// of 0..n-1 as a keyed Feistel network over the next power of four with cycle walking (a
// bijection evaluated independently per row: no shuffle pass, no memory traffic), u from Philox.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint64_t feistel_perm(uint64_t i, uint64_t n, int half_bits, uint32_t k0, uint32_t k1) {
  const uint32_t mask = (half_bits >= 32) ? 0xFFFFFFFFu : ((1u << half_bits) - 1u);
  uint64_t x = i;
  do {
    uint32_t L = (uint32_t)(x >> half_bits) & mask, R = (uint32_t)x & mask;
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      uint32_t f = mix32(R ^ (k0 + 0x9E3779B9u * (uint32_t)r)) ^ mix32((R >> 7) + k1 + (uint32_t)r);
      uint32_t nl = R;
      R = (L ^ f) & mask;
      L = nl;
    }
    x = ((uint64_t)L << half_bits) | R;
  } while (x >= n);
  return x;
}

__global__ void __launch_bounds__(256)
lhs_kernel(uint64_t seed, int64_t n, int scramble, double* __restrict__ out, int64_t row_stride,
           int64_t col_stride) {
  const int c = blockIdx.y;
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= n) return;
  int bits = 64 - __clzll((long long)(n - 1 > 0 ? n - 1 : 1));
  int half = (bits + 1) / 2;
  if (half < 1) half = 1;
  U4 kk = philox4x32_10({(uint32_t)c, 0x4C485331u, 0u, 0u}, (uint32_t)seed, (uint32_t)(seed >> 32));
  uint64_t p = feistel_perm((uint64_t)r, (uint64_t)n, half, kk.x, kk.y);
  double u = 0.5;
  if (scramble) {
    U4 o = philox4x32_10({(uint32_t)r, (uint32_t)((uint64_t)r >> 32), (uint32_t)c, 0x4C485332u},
                         (uint32_t)seed, (uint32_t)(seed >> 32));
    u = u01_53(o.x, o.y);
  }
  out[(int64_t)c * col_stride + r * row_stride] = __ddiv_rn(__dadd_rn((double)(p + 1), -u), (double)n);
}

}  // namespace
}  // namespace pbl

using pbl::kBadShape;
using pbl::kOk;

extern "C" {

int pbl_uniform_f64(uint64_t seed, uint64_t row0, int64_t n, int32_t d, double* out_dev,
                    int64_t row_stride, int64_t col_stride, void* stream) {
  if (n < 0 || d < 0 || !out_dev || (row0 & 1)) {
    pbl::set_last_error("pbl_uniform_f64: bad arguments (row0 must be even)");
    return kBadShape;
  }
  if (n == 0 || d == 0) return kOk;
  dim3 grid((unsigned)(((n + 1) / 2 + 255) / 256), (unsigned)d);
  pbl::philox_uniform_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seed, row0, n, d, out_dev, row_stride,
                                                                   col_stride);
  PBL_LAUNCH_CHECK();
  return kOk;
}

int pbl_sobol_direction_numbers(const int64_t* poly_dev, const int64_t* vinit_dev, int32_t vinit_cols,
                                int32_t d, int32_t bits, uint64_t* sv_dev, void* stream) {
  if (d < 1 || bits < 1 || bits > 64 || !poly_dev || !vinit_dev || !sv_dev) return kBadShape;
  pbl::sobol_direction_kernel<<<(d + 63) / 64, 64, 0, (cudaStream_t)stream>>>(poly_dev, vinit_dev, vinit_cols,
                                                                            d, bits, sv_dev);
  PBL_LAUNCH_CHECK();
  return kOk;
}

int pbl_sobol_scramble(const uint8_t* ltm_bits_dev, const uint8_t* shift_bits_dev, int32_t d, int32_t bits,
                       uint64_t* sv_dev, uint64_t* shift_dev, void* stream) {
  if (d < 1 || bits < 1 || bits > 64 || !ltm_bits_dev || !shift_bits_dev || !sv_dev || !shift_dev)
    return kBadShape;
  // block = one dimension's `bits` columns (<= 64 threads): the in-place update is block-local
  pbl::sobol_scramble_kernel<<<d, bits, 0, (cudaStream_t)stream>>>(ltm_bits_dev, shift_bits_dev, d, bits,
                                                                 sv_dev, shift_dev);
  PBL_LAUNCH_CHECK();
  return kOk;
}

int pbl_sobol_f64(const uint64_t* sv_dev, const uint64_t* shift_dev, int32_t d, int32_t bits,
                  uint64_t skip, int64_t n, double* out_dev, int64_t row_stride, int64_t col_stride,
                  void* stream) {
  if (d < 0 || n < 0 || bits < 1 || bits > 64 || !sv_dev || !shift_dev || !out_dev) return kBadShape;
  if (n == 0 || d == 0) return kOk;
  const uint64_t ngroups = ((skip + (uint64_t)n + 7) >> 3) - (skip >> 3);
  dim3 grid((unsigned)((ngroups + 255) / 256), (unsigned)d);
  if (bits <= 32)
    pbl::sobol_points_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(sv_dev, shift_dev, bits, skip, n, out_dev,
                                                                         row_stride, col_stride);
  else
    pbl::sobol_points_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(sv_dev, shift_dev, bits, skip, n, out_dev,
                                                                          row_stride, col_stride);
  PBL_LAUNCH_CHECK();
  return kOk;
}

int pbl_halton_f64(const int32_t* bases_dev, const int64_t* perms_dev, const int64_t* perm_off_dev,
                   const int32_t* perm_count_dev, int32_t d, uint64_t start_index, int64_t n,
                   double* out_dev, int64_t row_stride, int64_t col_stride, void* stream) {
  if (d < 0 || n < 0 || !bases_dev || !out_dev) return kBadShape;
  if (n == 0 || d == 0) return kOk;
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)d);
  pbl::halton_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(bases_dev, perms_dev, perm_off_dev, perm_count_dev,
                                                           start_index, n, out_dev, row_stride, col_stride);
  PBL_LAUNCH_CHECK();
  return kOk;
}

int pbl_lhs_f64(uint64_t seed, int64_t n, int32_t d, int32_t scramble, double* out_dev, int64_t row_stride,
                int64_t col_stride, void* stream) {
  if (d < 0 || n < 0 || !out_dev) return kBadShape;
  if (n == 0 || d == 0) return kOk;
  dim3 grid((unsigned)((n + 255) / 256), (unsigned)d);
  pbl::lhs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seed, n, scramble, out_dev, row_stride, col_stride);
  PBL_LAUNCH_CHECK();
  return kOk;
}

}  // extern "C"
