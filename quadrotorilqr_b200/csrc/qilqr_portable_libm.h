// =============================================================================
// qilqr_portable_libm.h -- sin / cos / atan2 written with IEEE +, -, *, / only (no fused multiply-add,
// no table look-ups that depend on the platform), so that the SAME source gives the SAME bits when it is
// compiled by g++ -ffp-contract=off for the host and by nvcc -fmad=false for sm_100a.
//
// Why it exists: the reference calls the C library's sin / cos / atan2 (through manif / Eigen).  The CUDA
// math library and glibc round those functions differently in the last place, which is the one rounding
// difference between the STRICT build of this library (-DQILQR_STRICT: no FMA, true divisions) and the CPU
// oracle.  Building BOTH with -D..._PORTABLE_LIBM (libqilqr_b200_strict_plibm.so and
// oracle/libqilqr_oracle_plibm.so) removes it: the two sides must then agree bit for bit on every
// problem, which is what tools/full_batch_parity.py checks (profiles/r2_parity_*.json).  Neither build is
// used in production; the production build calls the CUDA math library.
//
// Algorithms: the classic fdlibm scheme -- Cody-Waite reduction by pi/2 in two pieces followed by the
// degree-13 / degree-14 minimax kernels for sin / cos on [-pi/4, pi/4]; atan by interval reduction and an
// odd degree-23 polynomial.  Accuracy <= 1 ulp on the argument ranges of this code (|x| <= a few pi), checked
// against numpy in tests/test_portable_libm.py.  Valid for |x| < 1e5 (two-piece reduction).
// =============================================================================
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define QPL __host__ __device__ inline
#else
#define QPL inline
#endif

namespace qilqr_plibm {

QPL double k_sin(double x, double y) {  // sin(x + y), |x| <= pi/4, y the tail of x
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
               S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  const double z = x * x;
  const double v = z * x;
  const double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}
QPL double k_cos(double x, double y) {  // cos(x + y), |x| <= pi/4
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
               C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  const double z = x * x;
  const double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  const double ax = fabs(x);
  if (ax < 0.3) return 1.0 - (0.5 * z - (z * r - x * y));
  const double qx = (ax > 0.78125) ? 0.28125 : 0.25 * ax;
  const double hz = 0.5 * z - qx;
  const double a = 1.0 - qx;
  return a - (hz - (z * r - x * y));
}
QPL void sincos(double x, double *s, double *c) {
  const double invpio2 = 6.36619772367581382433e-01;  // 2/pi
  const double pio2_1 = 1.57079632673412561417e+00;   // first 33 bits of pi/2
  const double pio2_1t = 6.07710050650619224932e-11;  // pi/2 - pio2_1
  double y0 = x, y1 = 0.0;
  int n = 0;
  if (fabs(x) > 0.78539816339744830962) {
    const double fn = rint(x * invpio2);
    n = int(fn);
    const double r = x - fn * pio2_1;  // exact: fn is small and pio2_1 has 33 significant bits
    const double w = fn * pio2_1t;
    y0 = r - w;
    y1 = (r - y0) - w;
  }
  const double sn = k_sin(y0, y1), cs = k_cos(y0, y1);
  switch (n & 3) {
    case 0: *s = sn; *c = cs; break;
    case 1: *s = cs; *c = -sn; break;
    case 2: *s = -sn; *c = -cs; break;
    default: *s = -cs; *c = sn; break;
  }
}
QPL double sin(double x) { double s, c; sincos(x, &s, &c); return s; }
QPL double cos(double x) { double s, c; sincos(x, &s, &c); return c; }

QPL double atan(double x0) {
  const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01, aT2 = 1.42857142725034663711e-01,
               aT3 = -1.11111104054623557880e-01, aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
               aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02, aT8 = 4.97687799461593236017e-02,
               aT9 = -3.65315727442169155270e-02, aT10 = 1.62858201153657823623e-02;
  const bool neg = x0 < 0.0;
  double x = fabs(x0);
  double hi = 0.0, lo = 0.0;
  int id = -1;
  if (x >= 0.4375) {
    if (x < 1.1875) {
      if (x < 0.6875) { id = 0; hi = 4.63647609000806093515e-01; lo = 2.26987774529616870924e-17; x = (2.0 * x - 1.0) / (2.0 + x); }
      else            { id = 1; hi = 7.85398163397448278999e-01; lo = 3.06161699786838301793e-17; x = (x - 1.0) / (x + 1.0); }
    } else {
      if (x < 2.4375) { id = 2; hi = 9.82793723247329054082e-01; lo = 1.39033110312309984516e-17; x = (x - 1.5) / (1.0 + 1.5 * x); }
      else            { id = 3; hi = 1.57079632679489655800e+00; lo = 6.12323399573676603587e-17; x = -1.0 / x; }
    }
  }
  const double z = x * x;
  const double w = z * z;
  const double s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
  const double s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
  if (id < 0) {
    const double r = x - x * (s1 + s2);
    return neg ? -r : r;
  }
  const double r = hi - ((x * (s1 + s2) - lo) - x);
  return neg ? -r : r;
}
QPL double atan2(double y, double x) {
  const double pi = 3.1415926535897931160e+00, pi_lo = 1.2246467991473531772e-16, pio2 = 1.5707963267948965580e+00;
  if (x != x || y != y) return x + y;
  if (y == 0.0) return (x < 0.0 || (x == 0.0 && 1.0 / x < 0.0)) ? ((1.0 / y < 0.0) ? -pi : pi) : y;
  if (x == 0.0) return (y < 0.0) ? -pio2 : pio2;
  const double q = fabs(y / x);
  double z;
  if (q > 1.8446744073709552e19) z = pio2 + 0.5 * pi_lo;   // |y/x| > 2^64
  else if (x < 0.0 && q < 5.421010862427522e-20) z = 0.0;  // |y/x| < 2^-64, x < 0
  else z = atan(q);
  if (x > 0.0) return (y < 0.0) ? -z : z;
  return (y < 0.0) ? (z - pi_lo) - pi : pi - (z - pi_lo);
}

}  // namespace qilqr_plibm
