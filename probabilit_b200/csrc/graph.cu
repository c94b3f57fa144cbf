// Fused evaluation of a probabilit modeling graph: ONE kernel evaluates every node of the graph
// for each sample instead of one NumPy temporary (n doubles through DRAM) per node.
//
// Replaces the per-node loop of Node.sample_from_quantiles (reference
// src/probabilit/modeling.py:586-612): Distribution._sample = scipy.stats ppf (:795-812),
// Constant._sample (:760-763), Transform._sample (:954-956, :986-990, :1007-1009, :1071-1072) and
// the per-node finite check (:600-606).  The host mirror (probabilit_b200/modeling.py) compiles
// the graph, in the reference's evaluation order, into the bytecode of include/probabilit_b200.h.
//
// Execution model: one thread per sample; node values live in shared-memory "slots"
// ([slot][thread], conflict-free), the program is staged in shared memory and decoded uniformly
// by the whole block (no divergence except inside the iterative inverse CDFs).  HBM traffic is
// 8 B per quantile column read + 8 B per retained node written; with PBL_OP_UNIFORM the quantiles
// are generated in-kernel and only retained nodes touch HBM.
#include <vector>

#include "../../include/probabilit_b200.h"
#include "common.cuh"
#include "philox.cuh"
#include "special.cuh"

namespace pbl {
namespace {

constexpr int kGraphBlock = 256;

__device__ __forceinline__ double rn_scale_loc(double v, double scale, double loc) {
  return __dadd_rn(__dmul_rn(v, scale), loc);  // `_ppf(q) * scale + loc` as two NumPy ufuncs (no FMA)
}

// scipy rv_continuous.ppf wrapper (scipy/stats/_distn_infrastructure.py:2305-2348); `core` is
// only evaluated for valid parameters and 0 < q < 1
template <typename F>
__device__ __forceinline__ double ppf_continuous(double q, bool args_ok, double a, double b, double loc,
                                                 double scale, F core) {
  const bool cond0 = args_ok && (scale > 0.0) && (loc == loc);
  if (!cond0) return PBL_NAN;
  if (q == 0.0) return rn_scale_loc(a, scale, loc);
  if (q == 1.0) return rn_scale_loc(b, scale, loc);
  if (q > 0.0 && q < 1.0) return rn_scale_loc(core(), scale, loc);
  return PBL_NAN;
}

// scipy rv_discrete.ppf wrapper (:3745-3792): q == 0 -> a - 1 + loc even for invalid parameters
template <typename F>
__device__ __forceinline__ double ppf_discrete(double q, bool args_ok, double a, double b, double loc, F core) {
  const bool cond0 = args_ok && (loc == loc);
  if (q == 0.0) return __dadd_rn(a - 1.0, loc);
  if (!cond0) return PBL_NAN;
  if (q == 1.0) return __dadd_rn(b, loc);
  if (q > 0.0 && q < 1.0) return __dadd_rn(core(), loc);
  return PBL_NAN;
}

__device__ __forceinline__ double triang_core(double q, double c) {
  // scipy/stats/_continuous_distns.py:10126  where(q < c, sqrt(c*q), 1 - sqrt((1-c)*(1-q)))
  if (q < c) return __dsqrt_rn(__dmul_rn(c, q));
  return __dadd_rn(1.0, -__dsqrt_rn(__dmul_rn(__dadd_rn(1.0, -c), __dadd_rn(1.0, -q))));
}

__device__ __noinline__ double eval_ppf(int op, double q, double p0, double p1, double p2) {
  switch (op) {
    case PBL_PPF_NORM:
      return ppf_continuous(q, true, -kInf, kInf, p0, p1, [&] { return ndtri(q); });
    case PBL_PPF_UNIFORM:
      return ppf_continuous(q, true, 0.0, 1.0, p0, p1, [&] { return q; });
    case PBL_PPF_EXPON:
      return ppf_continuous(q, true, 0.0, kInf, p0, p1, [&] { return -log1p(-q); });
    case PBL_PPF_TRIANG:
      return ppf_continuous(q, p0 >= 0.0 && p0 <= 1.0, 0.0, 1.0, p1, p2, [&] { return triang_core(q, p0); });
    case PBL_PPF_GAMMA:
      return ppf_continuous(q, p0 > 0.0, 0.0, kInf, p1, p2, [&] { return igami(p0, q); });
    case PBL_PPF_LOGNORM:
      return ppf_continuous(q, p0 > 0.0, 0.0, kInf, p1, p2, [&] { return exp(__dmul_rn(p0, ndtri(q))); });
    case PBL_PPF_POISSON:
      return ppf_discrete(q, p0 >= 0.0, 0.0, kInf, p1, [&] { return poisson_ppf_core(q, p0); });
    case PBL_PPF_BINOM:
      return ppf_discrete(q, p0 >= 0.0 && p1 >= 0.0 && p1 <= 1.0 && p0 == rint(p0), 0.0, p0, p2,
                          [&] { return binom_ppf_core(q, p0, p1); });
    case PBL_PPF_BERNOULLI:
      return ppf_discrete(q, p0 >= 0.0 && p0 <= 1.0, 0.0, 1.0, p1, [&] { return binom_ppf_core(q, 1.0, p0); });
  }
  return PBL_NAN;
}

// four-parameter distributions: (a, b, loc, scale)
__device__ __noinline__ double eval_ppf4(int op, double q, double a, double b, double loc, double scale) {
  if (op == PBL_PPF_BETA)
    return ppf_continuous(q, a > 0.0 && b > 0.0, 0.0, 1.0, loc, scale, [&] { return ibeta_inv(a, b, q); });
  // truncnorm: support [a, b] in standard units, scipy argcheck a < b
  return ppf_continuous(q, a < b, a, b, loc, scale, [&] { return truncnorm_ppf_core(q, a, b); });
}

// ---- table-lookup distributions (reference modeling.py:825-927) ----
// np.interp(x, xp, fp) (numpy compiled_base.c arr_interp): clamp outside, exact knots return fp[j]
__device__ __noinline__ double table_interp(double x, const double* __restrict__ t, int m) {
  const double* xp = t;
  const double* fp = t + m;
  if (x != x) return x;
  if (x >= xp[m - 1]) return fp[m - 1];
  if (x < xp[0]) return fp[0];
  int lo = 0, hi = m - 1;  // xp[lo] <= x < xp[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (x >= xp[mid]) lo = mid; else hi = mid;
  }
  if (xp[lo] == x) return fp[lo];
  const double slope = __ddiv_rn(__dsub_rn(fp[lo + 1], fp[lo]), __dsub_rn(xp[lo + 1], xp[lo]));
  double r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[lo])), fp[lo]);
  if (r != r) {
    r = __dadd_rn(__dmul_rn(slope, __dsub_rn(x, xp[lo + 1])), fp[lo + 1]);
    if (r != r && fp[lo] == fp[lo + 1]) r = fp[lo];
  }
  return r;
}
// np.searchsorted(cum, q, side="right"): the first index with cum[idx] > q
__device__ __noinline__ double table_search_right(double q, const double* __restrict__ cum, int m) {
  int lo = 0, hi = m;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (cum[mid] <= q) lo = mid + 1; else hi = mid;
  }
  return (double)lo;
}
// np.quantile(data, q, method=...) on sorted data (numpy/lib/_function_base_impl.py _quantile, _lerp)
__device__ __noinline__ double table_quantile(double q, const double* __restrict__ a, int m, int method) {
  if (q != q) return q;
  const double n1 = (double)(m - 1);
  const double vi = __dmul_rn(n1, q);
  if (method == 1 || method == 2 || method == 3) {
    double idx = method == 1 ? floor(vi) : (method == 2 ? ceil(vi) : rint(vi));
    if (idx < 0.0) idx = 0.0;
    if (idx > n1) idx = n1;
    return a[(int)idx];
  }
  if (method == 5) {  // closest_observation: discontinuous, index from n*q - 1.5, ties to odd
    const double index = __dadd_rn(__dadd_rn(__dmul_rn((double)m, q), -1.0), -0.5);
    const double prev = floor(index);
    const double gamma = index - prev;
    double res = (gamma == 0.0 && fmod(prev, 2.0) == 1.0) ? prev : prev + 1.0;
    if (res < 0.0) res = 0.0;
    if (res > n1) res = n1;
    return a[(int)res];
  }
  double index = vi;
  if (method == 4) index = __dmul_rn(0.5, __dadd_rn(floor(vi), ceil(vi)));
  double prev = floor(index), next = prev + 1.0;
  if (index >= n1) prev = next = n1;  // _get_indexes: at / above the last element both are the last
  if (index < 0.0) prev = next = 0.0;
  double gamma = __dsub_rn(index, prev);
  if (method == 4) gamma = (fmod(index, 1.0) == 0.0) ? 0.0 : 0.5;
  const double lo = a[(int)prev], hi = a[(int)next];
  const double diff = __dsub_rn(hi, lo);
  double r = __dadd_rn(lo, __dmul_rn(diff, gamma));
  if (gamma >= 0.5) r = __dsub_rn(hi, __dmul_rn(diff, __dsub_rn(1.0, gamma)));
  return r;
}

// numpy npy_divmod for doubles (numpy/_core/src/npymath/npy_math_internal.h.src)
__device__ __forceinline__ double np_divmod(double a, double b, double* modulus) {
  double mod = fmod(a, b);
  if (b == 0.0) {
    *modulus = mod;
    return a / b;
  }
  double div = (a - mod) / b;
  if (mod != 0.0) {
    if ((b < 0.0) != (mod < 0.0)) {
      mod += b;
      div -= 1.0;
    }
  } else {
    mod = copysign(0.0, b);
  }
  double floordiv;
  if (div != 0.0) {
    floordiv = floor(div);
    if (div - floordiv > 0.5) floordiv += 1.0;
  } else {
    floordiv = copysign(0.0, a / b);
  }
  *modulus = mod;
  return floordiv;
}

__device__ __forceinline__ double b2d(bool b) { return b ? 1.0 : 0.0; }

__device__ __noinline__ double eval_binary(int op, double a, double b) {
  switch (op) {
    case PBL_OP_POW: return pow(a, b);
    case PBL_OP_FLOORDIV: {
      double m;
      return np_divmod(a, b, &m);
    }
    case PBL_OP_MOD: {
      double m;
      np_divmod(a, b, &m);
      return m;
    }
    case PBL_OP_MAX: return (a >= b || a != a) ? a : b;   // np.maximum propagates NaN
    case PBL_OP_MIN: return (a <= b || a != a) ? a : b;
    case PBL_OP_ATAN2: return atan2(a, b);
    case PBL_OP_LT: return b2d(a < b);
    case PBL_OP_LE: return b2d(a <= b);
    case PBL_OP_GT: return b2d(a > b);
    case PBL_OP_GE: return b2d(a >= b);
    case PBL_OP_EQ: return b2d(a == b);
    case PBL_OP_NE: return b2d(a != b);
    case PBL_OP_AND: return b2d(a != 0.0 && b != 0.0);
    case PBL_OP_OR: return b2d(a != 0.0 || b != 0.0);
    case PBL_OP_ISCLOSE: {  // np.isclose(a, b): rtol 1e-5, atol 1e-8, equal_nan False
      if (isfinite(a) && isfinite(b)) return b2d(fabs(a - b) <= __dadd_rn(1e-8, __dmul_rn(1e-5, fabs(b))));
      return b2d(a == b);
    }
  }
  return PBL_NAN;
}

__device__ __noinline__ double eval_unary(int op, double a) {
  switch (op) {
    case PBL_OP_LOG: return log(a);
    case PBL_OP_EXP: return exp(a);
    case PBL_OP_SIGN: return a > 0.0 ? 1.0 : (a < 0.0 ? -1.0 : (a == 0.0 ? 0.0 : a));
    case PBL_OP_LOG10: return log10(a);
    case PBL_OP_SIN: return sin(a);
    case PBL_OP_COS: return cos(a);
    case PBL_OP_TAN: return tan(a);
    case PBL_OP_ASIN: return asin(a);
    case PBL_OP_ACOS: return acos(a);
    case PBL_OP_ATAN: return atan(a);
    case PBL_OP_SINH: return sinh(a);
    case PBL_OP_COSH: return cosh(a);
    case PBL_OP_TANH: return tanh(a);
    case PBL_OP_ASINH: return asinh(a);
    case PBL_OP_ACOSH: return acosh(a);
    case PBL_OP_ATANH: return atanh(a);
  }
  return PBL_NAN;
}

struct GraphArgs {
  const pbl_graph_instr* program;
  int n_instr, n_slots;
  int64_t n;
  uint64_t row0;
  const double* const* inputs;
  double* const* outputs;
  int* first_nonfinite;
};

__global__ void __launch_bounds__(kGraphBlock) graph_eval_kernel(const GraphArgs g) {
  extern __shared__ __align__(16) unsigned char gsm_raw[];
  double* slots = reinterpret_cast<double*>(gsm_raw);  // [n_slots][kGraphBlock]
  pbl_graph_instr* prog = reinterpret_cast<pbl_graph_instr*>(slots + (size_t)g.n_slots * kGraphBlock);
  {
    const int words = g.n_instr * (int)(sizeof(pbl_graph_instr) / 8);
    const uint64_t* src = reinterpret_cast<const uint64_t*>(g.program);
    uint64_t* dst = reinterpret_cast<uint64_t*>(prog);
    for (int i = threadIdx.x; i < words; i += kGraphBlock) dst[i] = src[i];
  }
  __syncthreads();
  const int tid = threadIdx.x;
  double* S = slots + tid;
#define SLOT(i) S[(size_t)(i) * kGraphBlock]
  for (int64_t base = (int64_t)blockIdx.x * kGraphBlock; base < g.n; base += (int64_t)gridDim.x * kGraphBlock) {
    const int64_t row = base + tid;
    const bool active = row < g.n;
    for (int pc = 0; pc < g.n_instr; ++pc) {
      const pbl_graph_instr& in = prog[pc];
      const int op = in.op & 0xFF, flags = in.op;
      const int dst = in.dst & 0xFF;
      const int s0 = in.src[0], s1 = in.src[1];
      if (op < 16) {
        if (op == PBL_OP_LOAD) {
          SLOT(dst) = active ? ld_stream_f64(g.inputs[s0] + row) : 0.5;
        } else if (op == PBL_OP_STORE) {
          if (active) __stcs(g.outputs[s1] + row, SLOT(s0));
        } else if (op == PBL_OP_CHECK) {
          const double v = SLOT(s0);
          if (active && !isfinite(v)) atomicMin(g.first_nonfinite, s1);
        } else if (op == PBL_OP_MOV) {
          const double v = s0 >= 0 ? SLOT(s0) : in.imm[0];
          SLOT(dst) = v;
          if ((flags & PBL_GRAPH_STORE) && active) __stcs(g.outputs[(uint32_t)in.dst >> 20] + row, v);
        } else if (op == PBL_OP_UNIFORM) {
          const uint64_t seed = (uint64_t)__double_as_longlong(in.imm[0]);
          SLOT(dst) = philox_uniform_at(seed, g.row0 + (uint64_t)(active ? row : 0), (uint32_t)s0);
        }
        continue;
      }
      double a;
      if (flags & PBL_GRAPH_Q_INPUT) {  // fused LOAD: the quantile comes straight from its column
        a = active ? ld_stream_f64(g.inputs[s0] + row) : 0.5;
      } else if (flags & PBL_GRAPH_Q_UNIFORM) {  // fused UNIFORM: generated in-kernel
        a = philox_uniform_at((uint64_t)__double_as_longlong(in.imm[0]), g.row0 + (uint64_t)(active ? row : 0),
                              (uint32_t)s0);
      } else {
        a = s0 >= 0 ? SLOT(s0) : in.imm[0];
      }
      double r;
      if (op >= 64) {  // unary
        switch (op) {
          case PBL_OP_NEG: r = -a; break;
          case PBL_OP_ABS: r = fabs(a); break;
          case PBL_OP_FLOOR: r = floor(a); break;
          case PBL_OP_CEIL: r = ceil(a); break;
          case PBL_OP_SQRT: r = __dsqrt_rn(a); break;
          case PBL_OP_SQUARE: r = __dmul_rn(a, a); break;
          case PBL_OP_NOT: r = b2d(a == 0.0); break;
          case PBL_OP_LOOKUP: {
            const int m = (int)in.imm[1];
            r = (a >= 0.0 && a < (double)m) ? g.inputs[s1][(int)a] : PBL_NAN;
            break;
          }
          default: r = eval_unary(op, a); break;
        }
      } else if (op >= 32) {  // binary
        const double b = s1 >= 0 ? SLOT(s1) : in.imm[1];
        switch (op) {
          case PBL_OP_ADD: r = __dadd_rn(a, b); break;
          case PBL_OP_MUL: r = __dmul_rn(a, b); break;
          case PBL_OP_SUB: r = __dsub_rn(a, b); break;
          case PBL_OP_DIV: r = __ddiv_rn(a, b); break;
          default: r = eval_binary(op, a, b); break;
        }
      } else {  // ppf: operand 0 = q, then up to three parameters
        if (op >= PBL_PPF_TABLE_INTERP && op <= PBL_PPF_TABLE_QUANTILE) {  // table lookups: src[1] names a device table
          const double* tab = g.inputs[s1];
          const int m = (int)in.imm[1];
          r = op == PBL_PPF_TABLE_INTERP ? table_interp(a, tab, m)
              : (op == PBL_PPF_TABLE_SEARCH ? table_search_right(a, tab, m)
                                            : table_quantile(a, tab, m, (int)in.imm[2]));
          SLOT(dst) = r;
          if (active) {
            if ((flags & PBL_GRAPH_CHECK) && !isfinite(r)) atomicMin(g.first_nonfinite, (in.dst >> 8) & 0xFFF);
            if (flags & PBL_GRAPH_STORE) __stcs(g.outputs[(uint32_t)in.dst >> 20] + row, r);
          }
          continue;
        }
        const int s2 = in.src[2], s3 = in.src[3];
        const double p0 = s1 >= 0 ? SLOT(s1) : in.imm[1];
        const double p1 = s2 >= 0 ? SLOT(s2) : in.imm[2];
        const double p2 = s3 >= 0 ? SLOT(s3) : in.imm[3];
        if (op == PBL_PPF_NORM) {  // the common case stays inline
          r = ppf_continuous(a, true, -kInf, kInf, p0, p1, [&] { return ndtri(a); });
        } else if (op >= PBL_PPF_BETA) {  // (q, a, b, scale); the host adds loc with a separate ADD
          r = eval_ppf4(op, a, p0, p1, 0.0, p2);
        } else {
          r = eval_ppf(op, a, p0, p1, p2);
        }
      }
      SLOT(dst) = r;
      if (active) {
        if ((flags & PBL_GRAPH_CHECK) && !isfinite(r)) atomicMin(g.first_nonfinite, (in.dst >> 8) & 0xFFF);
        if (flags & PBL_GRAPH_STORE) __stcs(g.outputs[(uint32_t)in.dst >> 20] + row, r);
      }
    }
  }
#undef SLOT
}

__global__ void __launch_bounds__(256) ppf_kernel(int what, const double* __restrict__ q, int64_t n, double p0,
                                                  double p1, double p2, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    out[i] = what >= PBL_PPF_BETA ? eval_ppf4(what, q[i], p0, p1, 0.0, p2) : eval_ppf(what, q[i], p0, p1, p2);
}

}  // namespace
}  // namespace pbl

using pbl::kBadShape;
using pbl::kOk;

extern "C" {

int pbl_graph_eval_f64(const pbl_graph_instr* program, int32_t n_instr, int32_t n_slots, int64_t n, uint64_t row0,
                       const double* const* inputs_dev, int32_t n_inputs, double* const* outputs_dev,
                       int32_t n_outputs, int32_t* first_nonfinite, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (first_nonfinite) *first_nonfinite = -1;
  if (!program || n_instr < 0 || n_instr > PBL_GRAPH_MAX_INSTR || n_slots < 0 || n_slots > PBL_GRAPH_MAX_SLOTS ||
      n < 0 || n_inputs < 0 || n_outputs < 0) {
    pbl::set_last_error("pbl_graph_eval_f64: bad arguments (program size / slot count out of range)");
    return kBadShape;
  }
  for (int i = 0; i < n_instr; ++i) {  // validate every operand before it reaches the device
    const pbl_graph_instr& in = program[i];
    const int op = in.op & 0xFF;
    bool ok = (in.dst & 0xFF) < std::max(n_slots, 1) && in.dst >= 0;
    const bool fused_q = (in.op & (PBL_GRAPH_Q_INPUT | PBL_GRAPH_Q_UNIFORM)) != 0;
    const int nsrc = op == PBL_OP_STORE || op == PBL_OP_CHECK || op == PBL_OP_LOAD || op == PBL_OP_UNIFORM ? 0 : 4;
    const bool table_op = (op >= PBL_PPF_TABLE_INTERP && op <= PBL_PPF_TABLE_QUANTILE) || op == PBL_OP_LOOKUP;
    for (int s = fused_q ? 1 : 0; s < nsrc; ++s) ok = ok && (in.src[s] < n_slots || (table_op && s == 1));
    if (table_op) ok = ok && in.src[1] >= 0 && in.src[1] < n_inputs && in.imm[1] >= 1.0 && in.imm[1] < 2147483647.0;
    if (op == PBL_OP_LOAD || (in.op & PBL_GRAPH_Q_INPUT)) ok = ok && in.src[0] >= 0 && in.src[0] < n_inputs;
    if (fused_q) ok = ok && op >= PBL_PPF_NORM && op <= PBL_PPF_TRUNCNORM && in.src[0] >= 0;
    if (in.op & PBL_GRAPH_STORE) ok = ok && (int)((uint32_t)in.dst >> 20) < n_outputs && (op >= 16 || op == PBL_OP_MOV);
    if (in.op & PBL_GRAPH_CHECK) ok = ok && op >= 16;
    if (op == PBL_OP_STORE) ok = ok && in.src[0] >= 0 && in.src[0] < n_slots && in.src[1] >= 0 && in.src[1] < n_outputs;
    if (op == PBL_OP_CHECK) ok = ok && in.src[0] >= 0 && in.src[0] < n_slots && in.src[1] >= 0;
    if (!ok) {
      pbl::set_last_error("pbl_graph_eval_f64: instruction " + std::to_string(i) + " has an operand out of range");
      return kBadShape;
    }
  }
  if (n == 0 || n_instr == 0) return kOk;

  // one staging buffer: [flag | program | input pointers | output pointers]
  const size_t prog_bytes = (size_t)n_instr * sizeof(pbl_graph_instr);
  const size_t off_prog = 16, off_in = off_prog + prog_bytes, off_out = off_in + (size_t)n_inputs * 8;
  const size_t total = off_out + (size_t)n_outputs * 8 + 16;
  std::vector<unsigned char> host(total, 0);
  *reinterpret_cast<int*>(host.data()) = 0x7FFFFFFF;
  memcpy(host.data() + off_prog, program, prog_bytes);
  if (n_inputs) memcpy(host.data() + off_in, inputs_dev, (size_t)n_inputs * 8);
  if (n_outputs) memcpy(host.data() + off_out, outputs_dev, (size_t)n_outputs * 8);
  unsigned char* dev = nullptr;
  PBL_CUDA_CHECK(cudaMalloc((void**)&dev, total));
  cudaError_t e = cudaMemcpyAsync(dev, host.data(), total, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) {
    pbl::GraphArgs a;
    a.program = reinterpret_cast<const pbl_graph_instr*>(dev + off_prog);
    a.n_instr = n_instr;
    a.n_slots = n_slots;
    a.n = n;
    a.row0 = row0;
    a.inputs = reinterpret_cast<const double* const*>(dev + off_in);
    a.outputs = reinterpret_cast<double* const*>(dev + off_out);
    a.first_nonfinite = reinterpret_cast<int*>(dev);
    const size_t smem = (size_t)n_slots * pbl::kGraphBlock * 8 + prog_bytes;
    e = cudaFuncSetAttribute(pbl::graph_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      const int64_t blocks_needed = (n + pbl::kGraphBlock - 1) / pbl::kGraphBlock;
      const int64_t cap = (int64_t)pbl::num_sms() * 64;
      pbl::graph_eval_kernel<<<(unsigned)std::min(blocks_needed, cap), pbl::kGraphBlock, smem, stream>>>(a);
      pbl::g_kernel_launches.fetch_add(1, std::memory_order_relaxed);
      e = cudaGetLastError();
    }
    int flag = 0x7FFFFFFF;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&flag, dev, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess && first_nonfinite) *first_nonfinite = (flag == 0x7FFFFFFF) ? -1 : flag;
  }
  cudaFree(dev);
  PBL_CUDA_CHECK(e);
  return kOk;
}

int pbl_ppf_f64(int32_t what, const double* q_dev, int64_t n, double p0, double p1, double p2, double* out_dev,
                void* stream) {
  if (n < 0 || (n > 0 && (!q_dev || !out_dev)) || what < PBL_PPF_NORM || what > PBL_PPF_TRUNCNORM ||
      (what >= PBL_PPF_TABLE_INTERP && what <= PBL_PPF_TABLE_QUANTILE)) {
    pbl::set_last_error("pbl_ppf_f64: bad arguments");
    return kBadShape;
  }
  if (n == 0) return kOk;
  const int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)pbl::num_sms() * 32);
  pbl::ppf_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(what, q_dev, n, p0, p1, p2, out_dev);
  PBL_LAUNCH_CHECK();
  return kOk;
}

}  // extern "C"
