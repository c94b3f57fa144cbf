// Device restatement of Cephes ndtri (what scipy.stats.norm.ppf evaluates for 0<q<1;
// reference call sites: src/probabilit/correlation.py:395, src/probabilit/modeling.py:805-812).
// Same operation order as oracle/ndtri.py::ndtri_scalar.  Every product/sum uses the
// explicit round-to-nearest intrinsics so that ptxas cannot contract them into FMAs
// (scipy's x86-64 build has none), keeping results bit-identical to scipy whenever
// log() agrees with glibc's (CUDA's log is <= 1 ulp; the central branch has no log).
#pragma once
#include <cuda_runtime.h>

namespace pbl {

__device__ __forceinline__ double nd_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double nd_add(double a, double b) { return __dadd_rn(a, b); }

template <int N>
__device__ __forceinline__ double nd_polevl(double x, const double (&c)[N]) {
  double r = c[0];
#pragma unroll
  for (int i = 1; i < N; ++i) r = nd_add(nd_mul(r, x), c[i]);
  return r;
}
template <int N>
__device__ __forceinline__ double nd_p1evl(double x, const double (&c)[N]) {
  double r = nd_add(x, c[0]);
#pragma unroll
  for (int i = 1; i < N; ++i) r = nd_add(nd_mul(r, x), c[i]);
  return r;
}

__device__ __forceinline__ double ndtri(double y0) {
  const double P0[5] = {-5.99633501014107895267E1, 9.80010754185999661536E1,
                        -5.66762857469070293439E1, 1.39312609387279679503E1,
                        -1.23916583867381258016E0};
  const double Q0[8] = {1.95448858338141759834E0,  4.67627912898881538453E0,
                        8.63602421390890590575E1,  -2.25462687854119370527E2,
                        2.00260212380060660359E2,  -8.20372256168333339912E1,
                        1.59056225126211695515E1,  -1.18331621121330003142E0};
  const double P1[9] = {4.05544892305962419923E0,  3.15251094599893866154E1,
                        5.71628192246421288162E1,  4.40805073893200834700E1,
                        1.46849561928858024014E1,  2.18663306850790267539E0,
                        -1.40256079171354495875E-1, -3.50424626827848203418E-2,
                        -8.57456785154685413611E-4};
  const double Q1[8] = {1.57799883256466749731E1,  4.53907635128879210584E1,
                        4.13172038254672030440E1,  1.50425385692907503408E1,
                        2.50464946208309415979E0,  -1.42182922854787788574E-1,
                        -3.80806407691578277194E-2, -9.33259480895457427372E-4};
  const double P2[9] = {3.23774891776946035970E0,  6.91522889068984211695E0,
                        3.93881025292474443415E0,  1.33303460815807542389E0,
                        2.01485389549179081538E-1, 1.23716634817820021358E-2,
                        3.01581553508235416007E-4, 2.65806974686737550832E-6,
                        6.23974539184983293730E-9};
  const double Q2[8] = {6.02427039364742014255E0,  3.67983563856160859403E0,
                        1.37702099489081330271E0,  2.16236993594496635890E-1,
                        1.34204006088543189037E-2, 3.28014464682127739104E-4,
                        2.89247864745380683936E-6, 6.79019408009981274425E-9};
  const double s2pi = 2.50662827463100050242E0;
  const double expm2 = 0.13533528323661269189;

  if (y0 == 0.0) return -__longlong_as_double(0x7FF0000000000000LL);
  if (y0 == 1.0) return __longlong_as_double(0x7FF0000000000000LL);
  if (!(y0 > 0.0 && y0 < 1.0)) return __longlong_as_double(0x7FF8000000000000LL);
  bool negate = true;
  double y = y0;
  if (y > nd_add(1.0, -expm2)) {
    y = nd_add(1.0, -y);
    negate = false;
  }
  if (y > expm2) {
    y = nd_add(y, -0.5);
    double y2 = nd_mul(y, y);
    double x = nd_add(y, nd_mul(y, __ddiv_rn(nd_mul(y2, nd_polevl(y2, P0)), nd_p1evl(y2, Q0))));
    return nd_mul(x, s2pi);
  }
  double x = __dsqrt_rn(nd_mul(-2.0, log(y)));
  double x0 = nd_add(x, -__ddiv_rn(log(x), x));
  double z = __ddiv_rn(1.0, x);
  double x1;
  if (x < 8.0)
    x1 = __ddiv_rn(nd_mul(z, nd_polevl(z, P1)), nd_p1evl(z, Q1));
  else
    x1 = __ddiv_rn(nd_mul(z, nd_polevl(z, P2)), nd_p1evl(z, Q2));
  x = nd_add(x0, -x1);
  return negate ? -x : x;
}

}  // namespace pbl
