// =============================================================================
// oracle/qilqr_oracle.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Dependency-free FP64 CPU restatement of the reference iLQR hot path
// (nitishthatte/QuadrotorILQR):  src/ilqr.hh, src/quadrotor_model.{hh,cc},
// src/cost.hh plus the pieces of manif (@ab560a3a, WORKSPACE:55-61) and
// Eigen 3.4.0 (WORKSPACE:39-45) those files call.  manif and Eigen are NOT
// vendored in the reference and are not available in this image, so the
// reference itself cannot be compiled here; their
// published formulas (Sola et al. "A micro Lie theory", Barfoot's SE(3)
// Jacobian, Eigen's LLT / pivoted LDLT) are restated below.
//
// PARITY STATUS: pinned only by the reference's own known-answer tests
// (ilqr_test.cc:102-190, quadrotor_model_test.cc:94-143), its finite-
// difference self-consistency tests (quadrotor_model_test.cc:145-447,
// cost_test.cc:27-151) and scipy expm/logm cross-checks of the Lie functions
// (tests/test_oracle_*.py).  No reference test pins an N>3 solve, so beyond
// those cases this oracle is "parity unpinned" -- it IS the golden source.
// Its formulas are additionally checked against an independent 50-digit
// restatement of one full iLQR iteration (matrix exponential / logarithm,
// numerically differentiated; tests/test_oracle_mpmath.py: agreement 1e-14),
// which bounds its rounding but is no substitute for the literal reference.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use anything under oracle/.
//
// Everything is written "as the reference writes it": dense 12x12 products
// including structural zeros, left-to-right association, no FMA contraction
// (build with -ffp-contract=off), no symmetrisation, no regularisation.
// The scalar type is a template parameter so that the same code can be run
// with a FLOP-counting scalar (oracle_capi.cc: qoracle_count_flops).
// =============================================================================
#pragma once
#ifdef QORACLE_PORTABLE_LIBM
#include "../quadrotorilqr_b200/csrc/qilqr_portable_libm.h"
#endif
#include <cmath>
#include <cfloat>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace qoracle {

// ----------------------------------------------------------------------------
// FLOP-counting scalar (adds, muls, divs and libm calls each count 1).
// ----------------------------------------------------------------------------
struct FlopCounter {
  static inline thread_local unsigned long long add = 0, mul = 0, div = 0, fn = 0;
  static void reset() { add = mul = div = fn = 0; }
  static unsigned long long total() { return add + mul + div + fn; }
};
struct CountedDouble {
  double v;
  CountedDouble() : v(0) {}
  CountedDouble(double x) : v(x) {}
  explicit operator double() const { return v; }
};
inline CountedDouble operator+(CountedDouble a, CountedDouble b) { ++FlopCounter::add; return {a.v + b.v}; }
inline CountedDouble operator-(CountedDouble a, CountedDouble b) { ++FlopCounter::add; return {a.v - b.v}; }
inline CountedDouble operator*(CountedDouble a, CountedDouble b) { ++FlopCounter::mul; return {a.v * b.v}; }
inline CountedDouble operator/(CountedDouble a, CountedDouble b) { ++FlopCounter::div; return {a.v / b.v}; }
inline CountedDouble operator-(CountedDouble a) { return {-a.v}; }
inline CountedDouble &operator+=(CountedDouble &a, CountedDouble b) { a = a + b; return a; }
inline CountedDouble &operator-=(CountedDouble &a, CountedDouble b) { a = a - b; return a; }
inline CountedDouble &operator*=(CountedDouble &a, CountedDouble b) { a = a * b; return a; }
inline CountedDouble &operator/=(CountedDouble &a, CountedDouble b) { a = a / b; return a; }
inline bool operator<(CountedDouble a, CountedDouble b) { return a.v < b.v; }
inline bool operator>(CountedDouble a, CountedDouble b) { return a.v > b.v; }
inline bool operator<=(CountedDouble a, CountedDouble b) { return a.v <= b.v; }
inline bool operator>=(CountedDouble a, CountedDouble b) { return a.v >= b.v; }
inline bool operator==(CountedDouble a, CountedDouble b) { return a.v == b.v; }
inline bool operator!=(CountedDouble a, CountedDouble b) { return a.v != b.v; }
inline CountedDouble sin(CountedDouble a) { ++FlopCounter::fn; return {std::sin(a.v)}; }
inline CountedDouble cos(CountedDouble a) { ++FlopCounter::fn; return {std::cos(a.v)}; }
inline CountedDouble sqrt(CountedDouble a) { ++FlopCounter::fn; return {std::sqrt(a.v)}; }
inline CountedDouble atan2(CountedDouble a, CountedDouble b) { ++FlopCounter::fn; return {std::atan2(a.v, b.v)}; }
inline CountedDouble abs(CountedDouble a) { return {std::fabs(a.v)}; }
inline double to_double(CountedDouble a) { return a.v; }
inline double to_double(double a) { return a; }

using std::abs;
using std::sqrt;
#ifdef QORACLE_PORTABLE_LIBM
// Bit-for-bit comparison builds only (liboracle ..._plibm.so): sin / cos / atan2 from the portable implementation
// that the STRICT CUDA build can be compiled with too (quadrotorilqr_b200/csrc/qilqr_portable_libm.h), instead of
// the C library's.  Everything else is unchanged.
using qilqr_plibm::atan2;
using qilqr_plibm::cos;
using qilqr_plibm::sin;
#else
using std::atan2;
using std::cos;
using std::sin;
#endif

// ----------------------------------------------------------------------------
// Minimal fixed-size dense matrix (row-major storage; storage order is not
// observable).  Products are coefficient-wise dot products, k ascending, as
// Eigen's lazy small-matrix product evaluates them.
// ----------------------------------------------------------------------------
template <class T, int R, int C>
struct Mat {
  T a[R * C];
  Mat() { for (int i = 0; i < R * C; ++i) a[i] = T(0.0); }
  T &operator()(int r, int c) { return a[r * C + c]; }
  const T &operator()(int r, int c) const { return a[r * C + c]; }
  T &operator[](int i) { return a[i]; }  // vectors
  const T &operator[](int i) const { return a[i]; }
  static Mat Zero() { return Mat(); }
  static Mat Identity() {
    Mat m;
    for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1.0);
    return m;
  }
  Mat<T, C, R> transpose() const {
    Mat<T, C, R> t;
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < C; ++c) t(c, r) = (*this)(r, c);
    return t;
  }
  template <int BR, int BC>
  Mat<T, BR, BC> block(int r0, int c0) const {
    Mat<T, BR, BC> b;
    for (int r = 0; r < BR; ++r)
      for (int c = 0; c < BC; ++c) b(r, c) = (*this)(r0 + r, c0 + c);
    return b;
  }
  template <int BR, int BC>
  void set_block(int r0, int c0, const Mat<T, BR, BC> &b) {
    for (int r = 0; r < BR; ++r)
      for (int c = 0; c < BC; ++c) (*this)(r0 + r, c0 + c) = b(r, c);
  }
};
template <class T, int R, int K, int C>
Mat<T, R, C> operator*(const Mat<T, R, K> &x, const Mat<T, K, C> &y) {
  Mat<T, R, C> z;
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) {
      T s = x(r, 0) * y(0, c);
      for (int k = 1; k < K; ++k) s = s + x(r, k) * y(k, c);
      z(r, c) = s;
    }
  return z;
}
template <class T, int R, int C>
Mat<T, R, C> operator+(const Mat<T, R, C> &x, const Mat<T, R, C> &y) {
  Mat<T, R, C> z;
  for (int i = 0; i < R * C; ++i) z.a[i] = x.a[i] + y.a[i];
  return z;
}
template <class T, int R, int C>
Mat<T, R, C> operator-(const Mat<T, R, C> &x, const Mat<T, R, C> &y) {
  Mat<T, R, C> z;
  for (int i = 0; i < R * C; ++i) z.a[i] = x.a[i] - y.a[i];
  return z;
}
template <class T, int R, int C>
Mat<T, R, C> operator-(const Mat<T, R, C> &x) {
  Mat<T, R, C> z;
  for (int i = 0; i < R * C; ++i) z.a[i] = -x.a[i];
  return z;
}
template <class T, int R, int C>
Mat<T, R, C> operator*(const T &s, const Mat<T, R, C> &x) {
  Mat<T, R, C> z;
  for (int i = 0; i < R * C; ++i) z.a[i] = s * x.a[i];
  return z;
}
template <class T, int R, int C>
Mat<T, R, C> operator*(const Mat<T, R, C> &x, const T &s) {
  Mat<T, R, C> z;
  for (int i = 0; i < R * C; ++i) z.a[i] = x.a[i] * s;
  return z;
}
template <class T, int R, int C>
Mat<T, R, C> operator/(const Mat<T, R, C> &x, const T &s) {
  Mat<T, R, C> z;
  for (int i = 0; i < R * C; ++i) z.a[i] = x.a[i] / s;
  return z;
}
template <class T, int N>
T squared_norm(const Mat<T, N, 1> &v) {
  T s = v[0] * v[0];
  for (int i = 1; i < N; ++i) s = s + v[i] * v[i];
  return s;
}
template <class T>
Mat<T, 3, 3> hat(const Mat<T, 3, 1> &w) {  // manif skew()
  Mat<T, 3, 3> W;
  W(0, 1) = -w[2]; W(0, 2) = w[1];
  W(1, 0) = w[2];  W(1, 2) = -w[0];
  W(2, 0) = -w[1]; W(2, 1) = w[0];
  return W;
}

template <class T> using Vec3 = Mat<T, 3, 1>;
template <class T> using Vec4 = Mat<T, 4, 1>;
template <class T> using Vec6 = Mat<T, 6, 1>;
template <class T> using Vec12 = Mat<T, 12, 1>;
template <class T> using Mat3 = Mat<T, 3, 3>;
template <class T> using Mat6 = Mat<T, 6, 6>;
template <class T> using Mat12 = Mat<T, 12, 12>;

constexpr double kManifEps = 1e-14;  // manif Constants<double>::eps [3P-RECALL]

// ----------------------------------------------------------------------------
// SO(3): unit quaternion stored (x, y, z, w) as Eigen/manif do.
// ----------------------------------------------------------------------------
template <class T>
struct Quat {
  T x, y, z, w;
};

// Eigen::QuaternionBase::toRotationMatrix  [3P-RECALL, SURVEY App. B]
template <class T>
Mat3<T> rotation_matrix(const Quat<T> &q) {
  const T tx = T(2.0) * q.x, ty = T(2.0) * q.y, tz = T(2.0) * q.z;
  const T twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const T txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const T tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  Mat3<T> R;
  R(0, 0) = T(1.0) - (tyy + tzz); R(0, 1) = txy - twz;            R(0, 2) = txz + twy;
  R(1, 0) = txy + twz;            R(1, 1) = T(1.0) - (txx + tzz); R(1, 2) = tyz - twx;
  R(2, 0) = txz - twy;            R(2, 1) = tyz + twx;            R(2, 2) = T(1.0) - (txx + tyy);
  return R;
}
// Eigen quaternion product
template <class T>
Quat<T> quat_mul(const Quat<T> &a, const Quat<T> &b) {
  Quat<T> r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
// manif SO3::compose: product, then first-order renormalisation if the squared
// norm drifted by more than eps.
template <class T>
Quat<T> so3_compose(const Quat<T> &a, const Quat<T> &b) {
  Quat<T> r = quat_mul(a, b);
  const T sq = r.x * r.x + r.y * r.y + r.z * r.z + r.w * r.w;
  if (abs(sq - T(1.0)) > T(kManifEps)) {
    const T scale = T(2.0) / (T(1.0) + sq);
    r.x = r.x * scale; r.y = r.y * scale; r.z = r.z * scale; r.w = r.w * scale;
  }
  return r;
}
// manif SO3Tangent::exp
template <class T>
Quat<T> so3_exp(const Vec3<T> &w) {
  const T theta_sq = squared_norm(w);
  if (theta_sq > T(kManifEps)) {
    const T theta = sqrt(theta_sq);
    // Eigen::AngleAxis(theta, w.normalized()) -> quaternion
    const Vec3<T> axis = w / theta;
    const T ha = T(0.5) * theta;
    const T s = sin(ha);
    return Quat<T>{s * axis[0], s * axis[1], s * axis[2], cos(ha)};
  }
  return Quat<T>{w[0] / T(2.0), w[1] / T(2.0), w[2] / T(2.0), T(1.0)};
}
// manif SO3::log
template <class T>
Vec3<T> so3_log(const Quat<T> &q) {
  const T sin_angle_sq = q.x * q.x + q.y * q.y + q.z * q.z;
  T log_coeff;
  if (sin_angle_sq > T(kManifEps)) {
    const T sin_angle = sqrt(sin_angle_sq);
    const T cos_angle = q.w;
    const T two_angle = T(2.0) * ((cos_angle < T(0.0)) ? atan2(-sin_angle, -cos_angle)
                                                      : atan2(sin_angle, cos_angle));
    log_coeff = two_angle / sin_angle;
  } else {
    log_coeff = T(2.0);
  }
  Vec3<T> w;
  w[0] = q.x * log_coeff; w[1] = q.y * log_coeff; w[2] = q.z * log_coeff;
  return w;
}
// manif SO3Tangent::ljac / rjac / ljacinv / rjacinv
template <class T>
Mat3<T> so3_ljac(const Vec3<T> &w) {
  const T theta_sq = squared_norm(w);
  const Mat3<T> W = hat(w);
  if (theta_sq <= T(kManifEps)) return Mat3<T>::Identity() + T(0.5) * W;
  const T theta = sqrt(theta_sq);
  const T s = sin(theta), c = cos(theta);
  return Mat3<T>::Identity() + ((T(1.0) - c) / theta_sq) * W +
         ((theta - s) / (theta_sq * theta)) * (W * W);
}
template <class T>
Mat3<T> so3_rjac(const Vec3<T> &w) {
  const T theta_sq = squared_norm(w);
  const Mat3<T> W = hat(w);
  if (theta_sq <= T(kManifEps)) return Mat3<T>::Identity() - T(0.5) * W;
  const T theta = sqrt(theta_sq);
  const T s = sin(theta), c = cos(theta);
  return Mat3<T>::Identity() - ((T(1.0) - c) / theta_sq) * W +
         ((theta - s) / (theta_sq * theta)) * (W * W);
}
template <class T>
Mat3<T> so3_ljacinv(const Vec3<T> &w) {
  const T theta_sq = squared_norm(w);
  const Mat3<T> W = hat(w);
  if (theta_sq <= T(kManifEps)) return Mat3<T>::Identity() - T(0.5) * W;
  const T theta = sqrt(theta_sq);
  return Mat3<T>::Identity() - T(0.5) * W +
         (T(1.0) / theta_sq - (T(1.0) + cos(theta)) / (T(2.0) * theta * sin(theta))) * (W * W);
}
template <class T>
Mat3<T> so3_rjacinv(const Vec3<T> &w) {
  const T theta_sq = squared_norm(w);
  const Mat3<T> W = hat(w);
  if (theta_sq <= T(kManifEps)) return Mat3<T>::Identity() + T(0.5) * W;
  const T theta = sqrt(theta_sq);
  return Mat3<T>::Identity() + T(0.5) * W +
         (T(1.0) / theta_sq - (T(1.0) + cos(theta)) / (T(2.0) * theta * sin(theta))) * (W * W);
}

// ----------------------------------------------------------------------------
// SE(3): translation + unit quaternion; tangent = [lin(3), ang(3)].
// ----------------------------------------------------------------------------
template <class T>
struct SE3 {
  Vec3<T> t;
  Quat<T> q;
  static SE3 Identity() {
    SE3 X;
    X.q = Quat<T>{T(0.0), T(0.0), T(0.0), T(1.0)};
    return X;
  }
};
template <class T> Vec3<T> lin(const Vec6<T> &tau) { Vec3<T> v; v[0] = tau[0]; v[1] = tau[1]; v[2] = tau[2]; return v; }
template <class T> Vec3<T> ang(const Vec6<T> &tau) { Vec3<T> v; v[0] = tau[3]; v[1] = tau[4]; v[2] = tau[5]; return v; }

// manif SE3Tangent::exp:  (Jl(w) * v, Exp(w))
template <class T>
SE3<T> se3_exp(const Vec6<T> &tau) {
  SE3<T> X;
  X.t = so3_ljac(ang(tau)) * lin(tau);
  X.q = so3_exp(ang(tau));
  return X;
}
// manif SE3::log:  (Jl^-1(w) * t, w),  w = Log(q)
template <class T>
Vec6<T> se3_log(const SE3<T> &X) {
  const Vec3<T> w = so3_log(X.q);
  const Vec3<T> v = so3_ljacinv(w) * X.t;
  Vec6<T> tau;
  for (int i = 0; i < 3; ++i) { tau[i] = v[i]; tau[3 + i] = w[i]; }
  return tau;
}
// manif SE3::compose: (R_a t_b + t_a, q_a (x) q_b)
template <class T>
SE3<T> se3_compose(const SE3<T> &A, const SE3<T> &B) {
  SE3<T> X;
  X.t = rotation_matrix(A.q) * B.t + A.t;
  X.q = so3_compose(A.q, B.q);
  return X;
}
// manif SE3::inverse: (-R^T t, q*)
template <class T>
SE3<T> se3_inverse(const SE3<T> &A) {
  SE3<T> X;
  X.t = -(rotation_matrix(A.q).transpose() * A.t);
  X.q = Quat<T>{-A.q.x, -A.q.y, -A.q.z, A.q.w};
  return X;
}
// manif SE3::adj: [[R, hat(t) R], [0, R]]
template <class T>
Mat6<T> se3_adj(const SE3<T> &A) {
  const Mat3<T> R = rotation_matrix(A.q);
  Mat6<T> Ad;
  Ad.template set_block<3, 3>(0, 0, R);
  Ad.template set_block<3, 3>(3, 3, R);
  Ad.template set_block<3, 3>(0, 3, hat(A.t) * R);
  return Ad;
}
// manif SE3Tangent::fillQ (Barfoot, eq. 7.86, in manif's rearrangement)
template <class T>
Mat3<T> se3_fillQ(const Vec6<T> &c) {
  const Vec3<T> v = lin(c), w = ang(c);
  const T theta_sq = squared_norm(w);
  const T A = T(0.5);
  T B, C, D;
  if (theta_sq <= T(kManifEps)) {
    B = T(1.0 / 6.0) + T(1.0 / 120.0) * theta_sq;
    C = -T(1.0 / 24.0) + T(1.0 / 720.0) * theta_sq;
    D = -T(1.0 / 60.0);
  } else {
    const T theta = sqrt(theta_sq);
    const T s = sin(theta), co = cos(theta);
    B = (theta - s) / (theta_sq * theta);
    C = (T(1.0) - theta_sq / T(2.0) - co) / (theta_sq * theta_sq);
    D = C - T(3.0) * (theta - s - theta_sq * theta / T(6.0)) / (theta_sq * theta_sq * theta);
  }
  const Mat3<T> V = hat(v), W = hat(w);
  const Mat3<T> VW = V * W;
  const Mat3<T> WV = VW.transpose();
  const Mat3<T> WVW = WV * W;
  const Mat3<T> VWW = VW * W;
  return A * V + B * (WV + VW + WVW) - C * (VWW - VWW.transpose() - T(3.0) * WVW) -
         D * (WVW * W);
}
// manif SE3Tangent::rjac: [[Jr(w), Q(-tau)], [0, Jr(w)]]
template <class T>
Mat6<T> se3_rjac(const Vec6<T> &tau) {
  Mat6<T> J;
  const Mat3<T> Jr = so3_rjac(ang(tau));
  J.template set_block<3, 3>(0, 0, Jr);
  J.template set_block<3, 3>(3, 3, Jr);
  J.template set_block<3, 3>(0, 3, se3_fillQ(-tau));
  return J;
}
template <class T>
Mat6<T> se3_ljac(const Vec6<T> &tau) {
  Mat6<T> J;
  const Mat3<T> Jl = so3_ljac(ang(tau));
  J.template set_block<3, 3>(0, 0, Jl);
  J.template set_block<3, 3>(3, 3, Jl);
  J.template set_block<3, 3>(0, 3, se3_fillQ(tau));
  return J;
}
// manif SE3Tangent::rjacinv: [[Jr^-1, -Jr^-1 Q(-tau) Jr^-1], [0, Jr^-1]]
template <class T>
Mat6<T> se3_rjacinv(const Vec6<T> &tau) {
  Mat6<T> J;
  const Mat3<T> Ji = so3_rjacinv(ang(tau));
  const Mat3<T> Q = se3_fillQ(-tau);
  J.template set_block<3, 3>(0, 0, Ji);
  J.template set_block<3, 3>(3, 3, Ji);
  J.template set_block<3, 3>(0, 3, -(Ji * Q * Ji));
  return J;
}
template <class T>
Mat6<T> se3_ljacinv(const Vec6<T> &tau) {
  Mat6<T> J;
  const Mat3<T> Ji = so3_ljacinv(ang(tau));
  const Mat3<T> Q = se3_fillQ(tau);
  J.template set_block<3, 3>(0, 0, Ji);
  J.template set_block<3, 3>(3, 3, Ji);
  J.template set_block<3, 3>(0, 3, -(Ji * Q * Ji));
  return J;
}
// manif LieGroupBase::rplus / plus:  X o Exp(tau);  d/dX = Ad(Exp(tau)^-1),
// d/dtau = Jr(tau)
template <class T>
SE3<T> se3_plus(const SE3<T> &X, const Vec6<T> &tau, Mat6<T> *J_X = nullptr,
                Mat6<T> *J_tau = nullptr) {
  const SE3<T> E = se3_exp(tau);
  if (J_tau) *J_tau = se3_rjac(tau);
  if (J_X) *J_X = se3_adj(se3_inverse(E));
  return se3_compose(X, E);
}
// manif LieGroupBase::rminus / minus:  Log(B^-1 o A); d/dA = Jr^-1(t),
// d/dB = -Jl^-1(t)
template <class T>
Vec6<T> se3_minus(const SE3<T> &A, const SE3<T> &B, Mat6<T> *J_A = nullptr,
                  Mat6<T> *J_B = nullptr) {
  const Vec6<T> t = se3_log(se3_compose(se3_inverse(B), A));
  if (J_A) *J_A = se3_rjacinv(t);
  if (J_B) *J_B = -se3_ljacinv(t);
  return t;
}

// ----------------------------------------------------------------------------
// Eigen::LLT<Matrix3d> (lower Cholesky) and Eigen::LDLT<Matrix4d> (pivoted,
// lower triangle only) [3P-RECALL, SURVEY App. B]
// ----------------------------------------------------------------------------
template <class T>
struct LLT3 {
  Mat3<T> L;
  bool ok = true;
  void compute(const Mat3<T> &A) {
    ok = true;
    L = Mat3<T>::Zero();
    for (int k = 0; k < 3; ++k) {
      T x = A(k, k);
      for (int j = 0; j < k; ++j) x = x - L(k, j) * L(k, j);
      if (!(x > T(0.0))) { ok = false; return; }
      x = sqrt(x);
      L(k, k) = x;
      for (int i = k + 1; i < 3; ++i) {
        T s = A(i, k);
        for (int j = 0; j < k; ++j) s = s - L(i, j) * L(k, j);
        L(i, k) = s / x;
      }
    }
  }
  template <int C>
  Mat<T, 3, C> solve(const Mat<T, 3, C> &b) const {
    Mat<T, 3, C> x = b;
    for (int c = 0; c < C; ++c) {
      for (int i = 0; i < 3; ++i) {  // L y = b
        T s = x(i, c);
        for (int j = 0; j < i; ++j) s = s - L(i, j) * x(j, c);
        x(i, c) = s / L(i, i);
      }
      for (int i = 2; i >= 0; --i) {  // L^T x = y
        T s = x(i, c);
        for (int j = i + 1; j < 3; ++j) s = s - L(j, i) * x(j, c);
        x(i, c) = s / L(i, i);
      }
    }
    return x;
  }
};

template <class T>
struct LDLT4 {
  static constexpr int N = 4;
  Mat<T, 4, 4> m;  // strictly-lower: L, diagonal: D
  int transpositions[4];
  void compute(const Mat<T, 4, 4> &A) {
    m = A;
    for (int k = 0; k < N; ++k) {
      // pivot: largest |diagonal| of the remaining block
      int p = k;
      T best = abs(m(k, k));
      for (int i = k + 1; i < N; ++i)
        if (abs(m(i, i)) > best) { best = abs(m(i, i)); p = i; }
      transpositions[k] = p;
      if (p != k) {
        const int s = N - p - 1;
        for (int j = 0; j < k; ++j) std::swap(m(k, j), m(p, j));
        for (int i = 0; i < s; ++i) std::swap(m(p + 1 + i, k), m(p + 1 + i, p));
        std::swap(m(k, k), m(p, p));
        for (int i = k + 1; i < p; ++i) std::swap(m(i, k), m(p, i));
      }
      const int rs = N - k - 1;
      if (k > 0) {
        T temp[4];
        for (int j = 0; j < k; ++j) temp[j] = m(j, j) * m(k, j);
        T acc = m(k, 0) * temp[0];
        for (int j = 1; j < k; ++j) acc = acc + m(k, j) * temp[j];
        m(k, k) = m(k, k) - acc;
        for (int i = 0; i < rs; ++i) {
          T a2 = m(k + 1 + i, 0) * temp[0];
          for (int j = 1; j < k; ++j) a2 = a2 + m(k + 1 + i, j) * temp[j];
          m(k + 1 + i, k) = m(k + 1 + i, k) - a2;
        }
      }
      const T akk = m(k, k);
      const bool pivot_valid = abs(akk) > T(0.0);
      if (k == 0 && !pivot_valid) {
        for (int j = 0; j < N; ++j) transpositions[j] = j;
        return;
      }
      if (rs > 0 && pivot_valid)
        for (int i = 0; i < rs; ++i) m(k + 1 + i, k) = m(k + 1 + i, k) / akk;
    }
  }
  template <int C>
  Mat<T, 4, C> solve(const Mat<T, 4, C> &b) const {
    Mat<T, 4, C> x = b;
    for (int c = 0; c < C; ++c) {
      for (int k = 0; k < N; ++k)  // P b
        if (transpositions[k] != k) std::swap(x(k, c), x(transpositions[k], c));
      for (int i = 0; i < N; ++i) {  // unit-lower solve
        T s = x(i, c);
        for (int j = 0; j < i; ++j) s = s - m(i, j) * x(j, c);
        x(i, c) = s;
      }
      for (int i = 0; i < N; ++i) {  // D pseudo-inverse
        if (abs(m(i, i)) > T(DBL_MIN)) x(i, c) = x(i, c) / m(i, i);
        else x(i, c) = T(0.0);
      }
      for (int i = N - 1; i >= 0; --i) {  // unit-upper solve
        T s = x(i, c);
        for (int j = i + 1; j < N; ++j) s = s - m(j, i) * x(j, c);
        x(i, c) = s;
      }
      for (int k = N - 1; k >= 0; --k)  // P^T
        if (transpositions[k] != k) std::swap(x(k, c), x(transpositions[k], c));
    }
    return x;
  }
};

// ----------------------------------------------------------------------------
// QuadrotorModel  (src/quadrotor_model.hh:7-67, src/quadrotor_model.cc)
// ----------------------------------------------------------------------------
constexpr int CONFIG_DIM = 6;
constexpr int STATE_DIM = 12;
constexpr int CONTROL_DIM = 4;

template <class T>
struct State {  // quadrotor_model.hh:11-14
  SE3<T> inertial_from_body;
  Vec6<T> body_velocity;
};
template <class T>
struct StateTangent {  // quadrotor_model.hh:18-28
  Vec6<T> body_velocity;
  Vec6<T> body_acceleration;
  Vec12<T> coeffs() const {  // quadrotor_model.cc:124-132
    Vec12<T> c;
    for (int i = 0; i < 6; ++i) { c[i] = body_velocity[i]; c[6 + i] = body_acceleration[i]; }
    return c;
  }
  static StateTangent from_coeffs(const Vec12<T> &c) {
    StateTangent t;
    for (int i = 0; i < 6; ++i) { t.body_velocity[i] = c[i]; t.body_acceleration[i] = c[6 + i]; }
    return t;
  }
};
template <class T> using StateJacobian = Mat<T, 12, 12>;
template <class T> using ControlJacobian = Mat<T, 12, 4>;
template <class T> using Control = Vec4<T>;
template <class T>
struct DynamicsDifferentials {  // quadrotor_model.hh:42-45
  StateJacobian<T> J_x;
  ControlJacobian<T> J_u;
};
template <class T>
struct BinaryStateFuncDiffs {  // quadrotor_model.hh:47-50
  StateJacobian<T> J_x_lhs;
  StateJacobian<T> J_x_rhs;
};

template <class T>
StateTangent<T> operator*(const T &s, const StateTangent<T> &t) {  // quadrotor_model.cc:148-152
  return {s * t.body_velocity, s * t.body_acceleration};
}
// quadrotor_model.cc:202-206
template <class T>
State<T> plus(const State<T> &x, const StateTangent<T> &t) {
  return {se3_plus(x.inertial_from_body, t.body_velocity), x.body_velocity + t.body_acceleration};
}
// quadrotor_model.cc:174-200
template <class T>
State<T> add(const State<T> &x, const StateTangent<T> &t, BinaryStateFuncDiffs<T> *diffs) {
  if (diffs) {
    Mat6<T> J_X, J_tau;
    State<T> added{se3_plus(x.inertial_from_body, t.body_velocity, &J_X, &J_tau),
                   x.body_velocity + t.body_acceleration};
    diffs->J_x_lhs = StateJacobian<T>::Identity();
    diffs->J_x_lhs.template set_block<6, 6>(0, 0, J_X);
    diffs->J_x_rhs = StateJacobian<T>::Identity();
    diffs->J_x_rhs.template set_block<6, 6>(0, 0, J_tau);
    return added;
  }
  return plus(x, t);
}
// quadrotor_model.cc:215-219
template <class T>
StateTangent<T> minus(const State<T> &lhs, const State<T> &rhs) {
  return {se3_minus(lhs.inertial_from_body, rhs.inertial_from_body),
          lhs.body_velocity - rhs.body_velocity};
}
// quadrotor_model.cc:221-250
template <class T>
StateTangent<T> minus(const State<T> &lhs, const State<T> &rhs, BinaryStateFuncDiffs<T> *diffs) {
  if (diffs) {
    Mat6<T> J_l, J_r;
    StateTangent<T> d{se3_minus(lhs.inertial_from_body, rhs.inertial_from_body, &J_l, &J_r),
                      lhs.body_velocity - rhs.body_velocity};
    diffs->J_x_lhs = StateJacobian<T>::Identity();
    diffs->J_x_lhs.template set_block<6, 6>(0, 0, J_l);
    diffs->J_x_rhs = -StateJacobian<T>::Identity();
    diffs->J_x_rhs.template set_block<6, 6>(0, 0, J_r);
    return d;
  }
  return minus(lhs, rhs);
}
// quadrotor_model.cc:266-276
template <class T>
State<T> euler_step(const State<T> &x, const StateTangent<T> &x_dot, const T &dt_s,
                    BinaryStateFuncDiffs<T> *diffs = nullptr) {
  if (diffs) {
    const State<T> x_next = add(x, dt_s * x_dot, diffs);
    diffs->J_x_rhs = diffs->J_x_rhs * dt_s;
    return x_next;
  }
  return plus(x, dt_s * x_dot);
}

template <class T>
struct QuadrotorModel {
  T mass_kg_;
  Mat3<T> inertia_;
  LLT3<T> inertia_llt_;
  T arm_length_m_;
  T torque_to_thrust_ratio_m_;
  Mat<T, 3, 4> moment_arms_;
  T g_mpss_;

  // quadrotor_model.cc:6-25
  QuadrotorModel(T mass_kg, const Mat3<T> &inertia, T arm_length_m, T torque_to_thrust_ratio_m,
                 T g_mpss)
      : mass_kg_(mass_kg), inertia_(inertia), arm_length_m_(arm_length_m),
        torque_to_thrust_ratio_m_(torque_to_thrust_ratio_m), g_mpss_(g_mpss) {
    const T a = arm_length_m_, r = torque_to_thrust_ratio_m_;
    const T rows[3][4] = {{T(0.0), -a, T(0.0), a}, {a, T(0.0), -a, T(0.0)}, {-r, r, -r, r}};
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j) moment_arms_(i, j) = rows[i][j];
    inertia_llt_.compute(inertia_);
    bool symmetric = true;  // Eigen isApprox(inertia^T): ||A-A^T|| <= 1e-12 min(||A||,||A^T||)
    {
      double d2 = 0, n2 = 0;
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          const double d = to_double(inertia_(i, j)) - to_double(inertia_(j, i));
          d2 += d * d;
          n2 += to_double(inertia_(i, j)) * to_double(inertia_(i, j));
        }
      symmetric = d2 <= 1e-12 * 1e-12 * n2;
    }
    if (!inertia_llt_.ok || !symmetric)
      throw std::runtime_error("Inertia matrix is not positive definite!");
  }

  // quadrotor_model.cc:65-122
  StateTangent<T> continuous_dynamics(const State<T> &x, const Control<T> &u,
                                      DynamicsDifferentials<T> *diffs = nullptr) const {
    StateTangent<T> xdot;
    xdot.body_velocity = x.body_velocity;
    Vec3<T> ez;
    ez[2] = T(1.0);
    const T usum = ((u[0] + u[1]) + u[2]) + u[3];
    const Mat3<T> Rt = rotation_matrix(x.inertial_from_body.q).transpose();
    // -g * R^T * e_z + sum(u) * e_z / m
    const Vec3<T> acc_lin = ((-g_mpss_) * Rt) * ez + (usum * ez) / mass_kg_;
    const Vec3<T> M_Nm = moment_arms_ * u;
    const Vec3<T> omega = ang(xdot.body_velocity);
    const Vec3<T> acc_ang = inertia_llt_.solve(M_Nm - (hat(omega) * inertia_) * omega);
    for (int i = 0; i < 3; ++i) {
      xdot.body_acceleration[i] = acc_lin[i];
      xdot.body_acceleration[3 + i] = acc_ang[i];
    }
    if (diffs) {
      diffs->J_x = StateJacobian<T>::Zero();
      diffs->J_x.template set_block<6, 6>(0, 6, Mat6<T>::Identity());
      const Vec3<T> RTez = Rt * ez;
      diffs->J_x.template set_block<3, 3>(6, 3, (-g_mpss_) * hat(RTez));
      const Vec3<T> Jomega = inertia_ * omega;
      const Mat3<T> Jomega_diff = hat(omega) * inertia_ - hat(Jomega);
      diffs->J_x.template set_block<3, 3>(9, 9, -inertia_llt_.solve(Jomega_diff));
      diffs->J_u = ControlJacobian<T>::Zero();
      for (int j = 0; j < 4; ++j) diffs->J_u(8, j) = T(1.0) / mass_kg_;
      diffs->J_u.template set_block<3, 4>(9, 0, inertia_llt_.solve(moment_arms_));
    }
    return xdot;
  }

  // quadrotor_model.cc:33-49 (leak at :38-39 removed)
  State<T> discrete_dynamics(const State<T> &x, const Control<T> &u, const T &dt_s,
                             DynamicsDifferentials<T> *diffs = nullptr) const {
    const StateTangent<T> x_dot = continuous_dynamics(x, u, diffs);
    BinaryStateFuncDiffs<T> euler_diffs;
    const State<T> x_next = euler_step(x, x_dot, dt_s, diffs ? &euler_diffs : nullptr);
    if (diffs) {
      diffs->J_x = euler_diffs.J_x_lhs + euler_diffs.J_x_rhs * diffs->J_x;
      diffs->J_u = euler_diffs.J_x_rhs * diffs->J_u;
    }
    return x_next;
  }
};

// ----------------------------------------------------------------------------
// A SECOND MODEL for the ModelT concept of ilqr.hh:25-44 -- NOT in the reference.
// ILQR<ModelT> only needs State/Control, discrete_dynamics(x, u, dt, diffs*) and the
// free minus(); this variant keeps the quadrotor's state manifold and changes the
// dynamics function, to check that the CUDA solver is not hard-wired to one of them
// (SURVEY.md section 8(f)-4):
//   integrator_ = 1  classical RK4 over continuous_dynamics, the scheme of the block the
//                    reference left commented out (quadrotor_model.cc:51-63): stage states
//                    x (+) dt_i k_{i-1} with dt_i = {0, dt/2, dt/2, dt}, weights
//                    {1/6, 2/6, 2/6, 1/6} (doubles here; the comment holds them as floats),
//                    then x+ = x (+) dt xdot.  The differentials are the chain rule through
//                    the four stages (the reference never wrote them).
//   coriolis_ = true adds the transport term -omega x v to the body-frame linear acceleration,
//                    which quadrotor_model.cc:65-78 leaves out.
// integrator_ = 0, coriolis_ = false reproduces QuadrotorModel exactly.
// There is no reference behaviour to pin this model against: the oracle IS its definition
// (finite differences check its Jacobians in tests/test_model_variants.py).
// ----------------------------------------------------------------------------
template <class T>
struct QuadrotorModelVariant {
  QuadrotorModel<T> base_;
  int integrator_ = 0;
  bool coriolis_ = false;

  StateTangent<T> continuous_dynamics(const State<T> &x, const Control<T> &u,
                                      DynamicsDifferentials<T> *diffs = nullptr) const {
    StateTangent<T> xdot = base_.continuous_dynamics(x, u, diffs);
    if (coriolis_) {
      const Vec3<T> v = lin(x.body_velocity), w = ang(x.body_velocity);
      const Vec3<T> transport = hat(v) * w;  // v x omega = -(omega x v)
      for (int i = 0; i < 3; ++i) xdot.body_acceleration[i] = xdot.body_acceleration[i] + transport[i];
      if (diffs) {
        diffs->J_x.template set_block<3, 3>(6, 6, -hat(w));
        diffs->J_x.template set_block<3, 3>(6, 9, hat(v));
      }
    }
    return xdot;
  }

  State<T> discrete_dynamics(const State<T> &x, const Control<T> &u, const T &dt_s,
                             DynamicsDifferentials<T> *diffs = nullptr) const {
    if (integrator_ == 0) {  // quadrotor_model.cc:33-49
      const StateTangent<T> x_dot = continuous_dynamics(x, u, diffs);
      BinaryStateFuncDiffs<T> euler_diffs;
      const State<T> x_next = euler_step(x, x_dot, dt_s, diffs ? &euler_diffs : nullptr);
      if (diffs) {
        diffs->J_x = euler_diffs.J_x_lhs + euler_diffs.J_x_rhs * diffs->J_x;
        diffs->J_u = euler_diffs.J_x_rhs * diffs->J_u;
      }
      return x_next;
    }
    const T coeffs[4] = {T(1.0) / T(6.0), T(2.0) / T(6.0), T(2.0) / T(6.0), T(1.0) / T(6.0)};
    const T half_dt_s = dt_s / T(2.0);
    const T dt_table[4] = {T(0.0), half_dt_s, half_dt_s, dt_s};
    StateTangent<T> k{Vec6<T>::Zero(), Vec6<T>::Zero()};
    Vec12<T> x_dot = Vec12<T>::Zero();
    StateJacobian<T> dk_dx = StateJacobian<T>::Zero(), dxdot_dx = StateJacobian<T>::Zero();
    ControlJacobian<T> dk_du = ControlJacobian<T>::Zero(), dxdot_du = ControlJacobian<T>::Zero();
    for (int i = 0; i < 4; ++i) {
      BinaryStateFuncDiffs<T> e;
      const State<T> xi = euler_step(x, k, dt_table[i], diffs ? &e : nullptr);
      DynamicsDifferentials<T> f;
      k = continuous_dynamics(xi, u, diffs ? &f : nullptr);
      x_dot = x_dot + coeffs[i] * k.coeffs();
      if (diffs) {
        dk_du = f.J_x * (e.J_x_rhs * dk_du) + f.J_u;
        dk_dx = f.J_x * (e.J_x_lhs + e.J_x_rhs * dk_dx);
        dxdot_dx = dxdot_dx + coeffs[i] * dk_dx;
        dxdot_du = dxdot_du + coeffs[i] * dk_du;
      }
    }
    BinaryStateFuncDiffs<T> e;
    const State<T> x_next =
        euler_step(x, StateTangent<T>::from_coeffs(x_dot), dt_s, diffs ? &e : nullptr);
    if (diffs) {
      diffs->J_x = e.J_x_lhs + e.J_x_rhs * dxdot_dx;
      diffs->J_u = e.J_x_rhs * dxdot_du;
    }
    return x_next;
  }
};

// ----------------------------------------------------------------------------
// Trajectory (src/trajectory.hh:9-24), CostFunction (src/cost.hh)
// ----------------------------------------------------------------------------
template <class T>
struct TrajectoryPoint {
  T time_s;
  State<T> state;
  Control<T> control;
};
template <class T> using Trajectory = std::vector<TrajectoryPoint<T>>;

template <class T>
struct CostDifferentials {  // cost.hh:22-28
  Vec12<T> x;
  Vec4<T> u;
  Mat12<T> xx;
  Mat<T, 4, 4> uu;
  Mat<T, 12, 4> xu;
};

template <class T>
struct CostFunction {
  Mat12<T> Q_;
  Mat<T, 4, 4> R_;
  Trajectory<T> desired_trajectory_;

  // cost.hh:36-61
  T operator()(const State<T> &x, const Control<T> &u, int i,
               CostDifferentials<T> *diffs = nullptr) const {
    const State<T> &x_d = desired_trajectory_.at(i).state;
    const Control<T> &u_d = desired_trajectory_.at(i).control;
    BinaryStateFuncDiffs<T> minus_diffs;
    const Vec12<T> dx = minus(x, x_d, &minus_diffs).coeffs();  // Jacobians always formed (:42-43)
    const StateJacobian<T> &J = minus_diffs.J_x_lhs;
    const Vec4<T> du = u - u_d;
    const T cost = ((dx.transpose() * Q_) * dx)(0, 0) + ((du.transpose() * R_) * du)(0, 0);
    if (diffs) {
      diffs->x = (((T(2.0) * dx.transpose()) * Q_) * J).transpose();
      diffs->xx = ((T(2.0) * J.transpose()) * Q_) * J;
      diffs->u = ((T(2.0) * du.transpose()) * R_).transpose();
      diffs->uu = T(2.0) * R_;
      diffs->xu = Mat<T, 12, 4>::Zero();
    }
    return cost;
  }
};

// ----------------------------------------------------------------------------
// ILQROptions (src/ilqr_options.hh:4-22) + batch-only extensions that default
// to the reference behaviour.
// ----------------------------------------------------------------------------
struct LineSearchParams { double step_update; double desired_reduction_frac; int max_iters; };
struct ConvergenceCriteria { double rtol; double atol; double max_iters; };
struct ILQROptions {
  LineSearchParams line_search_params;
  ConvergenceCriteria convergence_criteria;
  bool populate_debug;
  // extensions (not in the reference; defaults reproduce it)
  bool symmetrize_vxx = false;    // V_xx <- (V_xx + V_xx^T)/2 after ilqr.hh:133
  double quu_regularization = 0;  // Q_uu += mu I before ilqr.hh:126
};

enum Status {  // how solve() ended (ilqr.hh:53-87)
  STATUS_CONVERGED_EXPECTED = 1,  // exit A, ilqr.hh:66-68
  STATUS_CONVERGED_ACTUAL = 2,    // exit B, ilqr.hh:82-84
  STATUS_MAX_ITERS = 3,           // exit C, ilqr.hh:86
  STATUS_LINE_SEARCH_FAILED = 4,  // throw at ilqr.hh:191-193
  STATUS_NONFINITE = 5,           // the same throw, reached because the last candidate cost was NaN/Inf
};

struct CostReductionTerms { double QuTk = 0, kTQuuk = 0; };  // ilqr.hh:13-16
inline double calculate_cost_reduction(const CostReductionTerms &t, double step = 1.0) {  // :18-22
  return step * t.QuTk + step * step * t.kTQuuk / 2.0;
}

template <class T>
struct ControlUpdate {  // ilqr.hh:43-49
  Vec4<T> ff_update;
  Mat<T, 4, 12> feedback;
};
template <class T> using ControlUpdateTrajectory = std::vector<ControlUpdate<T>>;

template <class T>
struct IterDebug {  // ilqr_debug.hh:9-13
  Trajectory<T> trajectory;
  double cost;
};

struct LineSearchFailure : std::runtime_error {
  using std::runtime_error::runtime_error;
};

template <class T>
struct SolveResult {
  Trajectory<T> traj;
  std::vector<IterDebug<T>> debug;        // filled iff populate_debug
  std::vector<double> cost_history;       // new_cost after every completed iteration
  std::vector<double> step_history;       // accepted alpha of every completed iteration
  ControlUpdateTrajectory<T> last_update; // gains of the last backward pass
  int status = 0;
  int backward_passes = 0;
  int rollouts = 0;
  double final_cost = 0;
};

template <class T, class ModelT = QuadrotorModel<T>>
struct ILQR {
  ModelT model_;
  CostFunction<T> cost_function_;
  T dt_s_;
  ILQROptions options_;
  mutable int rollouts_ = 0;
  mutable double last_line_search_cost_ = 0.0;

  // ilqr.hh:89-95
  T cost_trajectory(const Trajectory<T> &traj) const {
    T cost = T(0.0);
    for (size_t i = 0; i < traj.size(); ++i)
      cost = cost + cost_function_(traj[i].state, traj[i].control, int(i));
    return cost;
  }

  // ilqr.hh:97-147
  std::pair<ControlUpdateTrajectory<T>, CostReductionTerms> backwards_pass(
      const Trajectory<T> &traj) const {
    ControlUpdateTrajectory<T> upd;
    const int num_pts = int(traj.size());
    upd.reserve(num_pts);
    Vec12<T> v_x;
    Mat12<T> v_xx;
    CostReductionTerms terms;
    T QuTk = T(0.0), kTQuuk = T(0.0);
    for (int i = num_pts - 1; i >= 0; --i) {
      DynamicsDifferentials<T> dd;
      model_.discrete_dynamics(traj[i].state, traj[i].control, dt_s_, &dd);
      const auto &J_x = dd.J_x;
      const auto &J_u = dd.J_u;
      CostDifferentials<T> C;
      cost_function_(traj[i].state, traj[i].control, i, &C);

      CostDifferentials<T> Q;
      Q.x = C.x + J_x.transpose() * v_x;
      Q.u = C.u + J_u.transpose() * v_x;
      Q.xx = C.xx + (J_x.transpose() * v_xx) * J_x;
      Q.uu = C.uu + (J_u.transpose() * v_xx) * J_u;
      Q.xu = C.xu + (J_x.transpose() * v_xx) * J_u;

      if (options_.quu_regularization != 0.0)
        for (int j = 0; j < 4; ++j) Q.uu(j, j) = Q.uu(j, j) + T(options_.quu_regularization);

      LDLT4<T> ldlt;
      ldlt.compute(Q.uu);
      const Mat<T, 4, 12> K = -ldlt.solve(Q.xu.transpose());
      const Vec4<T> k = -ldlt.solve(Q.u);
      upd.push_back(ControlUpdate<T>{k, K});

      v_x = Q.x - (K.transpose() * Q.uu) * k;
      v_xx = Q.xx - (K.transpose() * Q.uu) * K;
      if (options_.symmetrize_vxx) v_xx = T(0.5) * (v_xx + v_xx.transpose());

      QuTk = QuTk + (Q.u.transpose() * k)(0, 0);
      kTQuuk = kTQuuk + ((k.transpose() * Q.uu) * k)(0, 0);
    }
    // std::reverse (ilqr.hh:143)
    ControlUpdateTrajectory<T> rev(upd.rbegin(), upd.rend());
    terms.QuTk = to_double(QuTk);
    terms.kTQuuk = to_double(kTQuuk);
    return {std::move(rev), terms};
  }

  // ilqr.hh:149-172
  Trajectory<T> forward_sim(const Trajectory<T> &cur, const ControlUpdateTrajectory<T> &upd,
                            double alpha = 1.0) const {
    Trajectory<T> out;
    out.reserve(cur.size());
    ++rollouts_;
    State<T> state = cur.front().state;
    for (size_t i = 0; i < cur.size(); ++i) {
      const Vec12<T> dx = minus(state, cur[i].state).coeffs();
      const Control<T> control = cur[i].control + T(alpha) * upd[i].ff_update + upd[i].feedback * dx;
      out.push_back(TrajectoryPoint<T>{cur[i].time_s, state, control});
      state = model_.discrete_dynamics(state, control, dt_s_);
    }
    return out;
  }

  // ilqr.hh:174-194
  struct LineSearchResult { Trajectory<T> traj; double cost; double step; };
  LineSearchResult line_search(const Trajectory<T> &cur, double cur_cost,
                               const ControlUpdateTrajectory<T> &upd,
                               const CostReductionTerms &terms) const {
    double step = 1.0;
    for (int i = 0; i < options_.line_search_params.max_iters; ++i) {
      Trajectory<T> nt = forward_sim(cur, upd, step);
      const double nc = to_double(cost_trajectory(nt));
      const double desired = options_.line_search_params.desired_reduction_frac *
                             calculate_cost_reduction(terms, step);
      if (nc - cur_cost < desired) return {std::move(nt), nc, step};
      step *= options_.line_search_params.step_update;
      last_line_search_cost_ = nc;
    }
    throw LineSearchFailure("Reached maximum number of line search iterations, " +
                            std::to_string(options_.line_search_params.max_iters) + "\n");
  }

  // ilqr.hh:196-205
  bool is_converged(double cost, double new_cost) const {
    if (std::fabs(cost - new_cost) / std::fabs(cost) < options_.convergence_criteria.rtol) return true;
    if (std::fabs(cost - new_cost) < options_.convergence_criteria.atol) return true;
    return false;
  }

  // ilqr.hh:53-87.  Instead of propagating the line-search exception the
  // result carries STATUS_LINE_SEARCH_FAILED and the last accepted trajectory.
  SolveResult<T> solve(const Trajectory<T> &initial_traj) const {
    SolveResult<T> r;
    rollouts_ = 0;
    Trajectory<T> traj = initial_traj;
    double new_cost = to_double(cost_trajectory(traj));
    r.status = STATUS_MAX_ITERS;
    for (int i = 0; i < options_.convergence_criteria.max_iters; ++i) {
      auto [upd, terms] = backwards_pass(traj);
      ++r.backward_passes;
      r.last_update = upd;
      const double cost = new_cost;
      const double expected_new_cost = cost + calculate_cost_reduction(terms);
      if (i > 0 && is_converged(cost, expected_new_cost)) {
        r.status = STATUS_CONVERGED_EXPECTED;
        break;
      }
      double step = 1.0;
      if (i == 0) {
        traj = forward_sim(traj, upd, 1.0);
        new_cost = to_double(cost_trajectory(traj));
      } else {
        try {
          auto ls = line_search(traj, cost, upd, terms);
          traj = std::move(ls.traj);
          new_cost = ls.cost;
          step = ls.step;
        } catch (const LineSearchFailure &) {
          r.status = std::isfinite(last_line_search_cost_) ? STATUS_LINE_SEARCH_FAILED : STATUS_NONFINITE;
          break;
        }
      }
      if (options_.populate_debug) r.debug.push_back(IterDebug<T>{traj, new_cost});
      r.cost_history.push_back(new_cost);
      r.step_history.push_back(step);
      if (i > 0 && is_converged(cost, new_cost)) {
        r.status = STATUS_CONVERGED_ACTUAL;
        break;
      }
    }
    r.traj = std::move(traj);
    r.final_cost = new_cost;
    r.rollouts = rollouts_;
    return r;
  }
};

}  // namespace qoracle
