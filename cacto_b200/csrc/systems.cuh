// systems.cuh -- per-sample device functions for the six CACTO systems (dynamics step, Jacobians,
// end-effector position, reward).  Everything is templated on the system and on the scalar type so
// that each kernel instantiation carries only its own system's code and registers.
//
// Reference: environment.py (cited per function), robot_utils.py:399-405 (explicit Euler on
// M ddq = u - nle), urdf/*.urdf through cacto_chain.  Pinocchio's CRBA/RNEA/ABA-derivatives are
// replaced by (i) a closed form for the planar 3R arm written from the URDF link parameters and
// (ii) a 3-vector recursive Newton-Euler for the UR5 whose exact derivatives are taken in forward
// (tangent) mode.
#pragma once
#include <cuda_runtime.h>
#include "cacto_b200.h"

namespace cacto {

// ------------------------------------------------------------------------------------------ math
__device__ __forceinline__ void sincos_(float x, float& s, float& c) { sincosf(x, &s, &c); }
__device__ __forceinline__ void sincos_(double x, double& s, double& c) { sincos(x, &s, &c); }
__device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
__device__ __forceinline__ float exp_(float x) { return expf(x); }
__device__ __forceinline__ double exp_(double x) { return exp(x); }
__device__ __forceinline__ float log_(float x) { return logf(x); }
__device__ __forceinline__ double log_(double x) { return log(x); }
__device__ __forceinline__ float tan_(float x) { return tanf(x); }
__device__ __forceinline__ double tan_(double x) { return tan(x); }

// log(exp(z) + 1): the reference evaluates it literally in fp64 (environment.py:258-263) where
// exp(z) stays finite for z < 709; the branch keeps float32 finite for the same arguments.
template <typename T>
__device__ __forceinline__ T softplus_(T z) {
  return z > T(30) ? z + log_(T(1) + exp_(-z)) : log_(exp_(z) + T(1));
}

template <typename T>
__device__ __forceinline__ T pow10_(T x) {
  T x2 = x * x, x4 = x2 * x2, x8 = x4 * x4;
  return x8 * x2;
}

// ------------------------------------------------------------------------------------------ dual numbers
template <typename T>
struct Dual {
  T v, d;
  __device__ __forceinline__ Dual() {}
  __device__ __forceinline__ Dual(T a) : v(a), d(T(0)) {}
  __device__ __forceinline__ Dual(T a, T b) : v(a), d(b) {}
};
template <typename T> __device__ __forceinline__ Dual<T> operator+(Dual<T> a, Dual<T> b) { return Dual<T>(a.v + b.v, a.d + b.d); }
template <typename T> __device__ __forceinline__ Dual<T> operator-(Dual<T> a, Dual<T> b) { return Dual<T>(a.v - b.v, a.d - b.d); }
template <typename T> __device__ __forceinline__ Dual<T> operator-(Dual<T> a) { return Dual<T>(-a.v, -a.d); }
template <typename T> __device__ __forceinline__ Dual<T> operator*(Dual<T> a, Dual<T> b) { return Dual<T>(a.v * b.v, a.v * b.d + a.d * b.v); }
template <typename T> __device__ __forceinline__ void sincos_(Dual<T> x, Dual<T>& s, Dual<T>& c) {
  T sv, cv;
  sincos_(x.v, sv, cv);
  s = Dual<T>(sv, cv * x.d);
  c = Dual<T>(cv, -sv * x.d);
}
// sin / cos of x given sin / cos of its real part (one overload per number type; HDual: backward.cu)
__device__ __forceinline__ void sincos_from(float, float sv, float cv, float& s, float& c) { s = sv; c = cv; }
__device__ __forceinline__ void sincos_from(double, double sv, double cv, double& s, double& c) { s = sv; c = cv; }
template <typename T> __device__ __forceinline__ void sincos_from(Dual<T> x, T sv, T cv, Dual<T>& s, Dual<T>& c) {
  s = Dual<T>(sv, cv * x.d);
  c = Dual<T>(cv, -sv * x.d);
}
template <typename S> struct scalar_of { typedef S type; };
template <typename T> struct scalar_of<Dual<T>> { typedef T type; };

// ------------------------------------------------------------------------------------------ small vectors
template <typename S> __device__ __forceinline__ void cross3(const S* a, const S* b, S* o) {
  S x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}

// component access by a (warp-uniform) runtime axis without dynamic register indexing
template <typename S> __device__ __forceinline__ void axis_add(S* v, int ax, S x) {
  if (ax == 0) v[0] = v[0] + x; else if (ax == 1) v[1] = v[1] + x; else v[2] = v[2] + x;
}
template <typename S> __device__ __forceinline__ S axis_get(const S* v, int ax) { return ax == 0 ? v[0] : (ax == 1 ? v[1] : v[2]); }

// ------------------------------------------------------------------------------------------ serial chain (UR5)
// Rotation parent->child of joint i: R = Rfix * Rot(axis, q).  Apply R or R^T to a vector.
template <typename S>
struct JointRot {
  typedef typename scalar_of<S>::type T;
  T F[9];
  S s, c;
  int axis;
  __device__ __forceinline__ void rot_axis(const S* x, S* y) const {   // y = Rot(axis, q) x
    S a = x[0], b = x[1], d = x[2];
    if (axis == 0) { y[0] = a; y[1] = c * b - s * d; y[2] = s * b + c * d; }
    else if (axis == 1) { y[0] = c * a + s * d; y[1] = b; y[2] = c * d - s * a; }
    else { y[0] = c * a - s * b; y[1] = s * a + c * b; y[2] = d; }
  }
  __device__ __forceinline__ void rot_axis_T(const S* x, S* y) const { // y = Rot(axis, q)^T x
    S a = x[0], b = x[1], d = x[2];
    if (axis == 0) { y[0] = a; y[1] = c * b + s * d; y[2] = c * d - s * b; }
    else if (axis == 1) { y[0] = c * a - s * d; y[1] = b; y[2] = s * a + c * d; }
    else { y[0] = c * a + s * b; y[1] = c * b - s * a; y[2] = d; }
  }
  __device__ __forceinline__ void apply(const S* x, S* y) const {      // y = R x
    S t[3];
    rot_axis(x, t);
    for (int r = 0; r < 3; ++r) y[r] = S(F[3 * r]) * t[0] + S(F[3 * r + 1]) * t[1] + S(F[3 * r + 2]) * t[2];
  }
  __device__ __forceinline__ void apply_T(const S* x, S* y) const {    // y = R^T x
    S t[3];
    for (int r = 0; r < 3; ++r) t[r] = S(F[r]) * x[0] + S(F[3 + r]) * x[1] + S(F[6 + r]) * x[2];
    rot_axis_T(t, y);
  }
};

// tau = M(q) a + nle(q, v) for an all-revolute chain (recursive Newton-Euler in link frames).
// with_vel = false drops every velocity term (used for the columns of M).
// sc_pre: optional [2 N] = sin q_0 .. sin q_{N-1}, cos q_0 .. (the UR5 step and derivative call this seven times at the same q:
// the fp64 sincos of the six joints were 36 of the 42 evaluated per step).
template <typename S, int N, bool WITH_VEL>
__device__ void chain_rnea(const cacto_chain& ch, const S* q, const S* v, const S* a, typename scalar_of<S>::type grav, S* tau,
                           const S* sc_pre = nullptr) {
  typedef typename scalar_of<S>::type T;
  S sn[N], cs[N];
  S Fv[N][3], Nv[N][3];
  S w[3] = {S(T(0)), S(T(0)), S(T(0))}, wd[3] = {S(T(0)), S(T(0)), S(T(0))};
  S acc[3] = {S(T(0)), S(T(0)), S(grav)};
#pragma unroll
  for (int i = 0; i < N; ++i) {
    JointRot<S> J;
    for (int k = 0; k < 9; ++k) J.F[k] = T(ch.R[i][k]);
    J.axis = ch.axis[i];
    if (sc_pre != nullptr) { J.s = sc_pre[i]; J.c = sc_pre[N + i]; }
    else sincos_(q[i], J.s, J.c);
    sn[i] = J.s; cs[i] = J.c;
    S p[3] = {S(T(ch.p[i][0])), S(T(ch.p[i][1])), S(T(ch.p[i][2]))};
    S t0[3], t1[3], wp[3];
    // linear acceleration of the joint origin, expressed in the child frame
    cross3(wd, p, t0);
    if (WITH_VEL) { cross3(w, p, t1); cross3(w, t1, t1); for (int k = 0; k < 3; ++k) t0[k] = t0[k] + t1[k]; }
    for (int k = 0; k < 3; ++k) t0[k] = t0[k] + acc[k];
    J.apply_T(t0, acc);
    J.apply_T(wd, t0);
    const int ax = J.axis;
    if (WITH_VEL) {
      J.apply_T(w, wp);
      S ev[3] = {S(T(0)), S(T(0)), S(T(0))};
      axis_add(ev, ax, v[i]);
      cross3(wp, ev, t1);
      for (int k = 0; k < 3; ++k) { wd[k] = t0[k] + t1[k]; w[k] = wp[k]; }
      axis_add(w, ax, v[i]);
    } else {
      for (int k = 0; k < 3; ++k) wd[k] = t0[k];
    }
    axis_add(wd, ax, a[i]);
    // body wrench
    S c[3] = {S(T(ch.com[i][0])), S(T(ch.com[i][1])), S(T(ch.com[i][2]))};
    S ac[3];
    cross3(wd, c, ac);
    if (WITH_VEL) { cross3(w, c, t1); cross3(w, t1, t1); for (int k = 0; k < 3; ++k) ac[k] = ac[k] + t1[k]; }
    const T m = T(ch.mass[i]);
    for (int k = 0; k < 3; ++k) Fv[i][k] = S(m) * (ac[k] + acc[k]);
    const T ixx = T(ch.inertia[i][0]), iyy = T(ch.inertia[i][1]), izz = T(ch.inertia[i][2]);
    const T ixy = T(ch.inertia[i][3]), ixz = T(ch.inertia[i][4]), iyz = T(ch.inertia[i][5]);
    S Iwd[3] = {S(ixx) * wd[0] + S(ixy) * wd[1] + S(ixz) * wd[2], S(ixy) * wd[0] + S(iyy) * wd[1] + S(iyz) * wd[2],
                S(ixz) * wd[0] + S(iyz) * wd[1] + S(izz) * wd[2]};
    if (WITH_VEL) {
      S Iw[3] = {S(ixx) * w[0] + S(ixy) * w[1] + S(ixz) * w[2], S(ixy) * w[0] + S(iyy) * w[1] + S(iyz) * w[2],
                 S(ixz) * w[0] + S(iyz) * w[1] + S(izz) * w[2]};
      cross3(w, Iw, t1);
      for (int k = 0; k < 3; ++k) Iwd[k] = Iwd[k] + t1[k];
    }
    cross3(c, Fv[i], t1);
    for (int k = 0; k < 3; ++k) Nv[i][k] = Iwd[k] + t1[k];
  }
  S f[3] = {S(T(0)), S(T(0)), S(T(0))}, n[3] = {S(T(0)), S(T(0)), S(T(0))};
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    if (i < N - 1) {
      JointRot<S> J;
      for (int k = 0; k < 9; ++k) J.F[k] = T(ch.R[i + 1][k]);
      J.axis = ch.axis[i + 1];
      J.s = sn[i + 1]; J.c = cs[i + 1];
      S fc[3], nc[3], t1[3];
      J.apply(f, fc);
      J.apply(n, nc);
      S p[3] = {S(T(ch.p[i + 1][0])), S(T(ch.p[i + 1][1])), S(T(ch.p[i + 1][2]))};
      cross3(p, fc, t1);
      for (int k = 0; k < 3; ++k) { f[k] = fc[k]; n[k] = nc[k] + t1[k]; }
    }
    for (int k = 0; k < 3; ++k) { f[k] = f[k] + Fv[i][k]; n[k] = n[k] + Nv[i][k]; }
    tau[i] = axis_get(n, ch.axis[i]);
  }
}

// In-place Cholesky solve helpers for small SPD matrices (M is N x N row-major, overwritten by L).
template <typename T, int N>
__device__ __forceinline__ void chol_factor(T* M) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    T d = M[j * N + j];
    for (int k = 0; k < j; ++k) d -= M[j * N + k] * M[j * N + k];
    d = sqrt_(d);
    M[j * N + j] = d;
    T inv = T(1) / d;
    for (int i = j + 1; i < N; ++i) {
      T s = M[i * N + j];
      for (int k = 0; k < j; ++k) s -= M[i * N + k] * M[j * N + k];
      M[i * N + j] = s * inv;
    }
  }
}
template <typename T, int N>
__device__ __forceinline__ void chol_solve(const T* L, T* b) {   // b <- M^-1 b
#pragma unroll
  for (int i = 0; i < N; ++i) {
    T s = b[i];
    for (int k = 0; k < i; ++k) s -= L[i * N + k] * b[k];
    b[i] = s / L[i * N + i];
  }
#pragma unroll
  for (int i = N - 1; i >= 0; --i) {
    T s = b[i];
    for (int k = i + 1; k < N; ++k) s -= L[k * N + i] * b[k];
    b[i] = s / L[i * N + i];
  }
}

// ------------------------------------------------------------------------------------------ planar 3R closed form
// M(q) = M0 + Ma cos q2 + Mb cos q3 + Mc cos(q2+q3) with coefficients from the URDF link parameters
// (planar_manipulator_3dof.urdf:24-98).  Gravity is parallel to the joint axes: nle = Coriolis only.
template <typename T>
struct Planar3R {
  T a, b, c;          // a = m2 l1 r2 + m3 l1 l2, b = m3 l2 r3, c = m3 l1 r3
  T m0[6];            // constant part: 11 12 13 22 23 33
  __device__ __forceinline__ explicit Planar3R(const cacto_chain& ch) {
    const T m1 = T(ch.mass[0]), m2 = T(ch.mass[1]), m3 = T(ch.mass[2]);
    const T r1 = T(ch.com[0][0]), r2 = T(ch.com[1][0]), r3 = T(ch.com[2][0]);
    const T l1 = T(ch.p[1][0]), l2 = T(ch.p[2][0]);
    const T I1 = T(ch.inertia[0][2]), I2 = T(ch.inertia[1][2]), I3 = T(ch.inertia[2][2]);
    a = m2 * l1 * r2 + m3 * l1 * l2;
    b = m3 * l2 * r3;
    c = m3 * l1 * r3;
    m0[5] = I3 + m3 * r3 * r3;
    m0[4] = m0[5];
    m0[3] = I2 + m2 * r2 * r2 + m3 * l2 * l2 + m0[5];
    m0[2] = m0[5];
    m0[1] = m0[3];
    m0[0] = I1 + m1 * r1 * r1 + (m2 + m3) * l1 * l1 + m0[3];
  }
  // symmetric 3x3 stored as [11 12 13 22 23 33]; pattern * coefficient
  __device__ __forceinline__ void pat(T ka, T kb, T kc, T* o) const {
    // Ma = [[2a,a,0],[a,0,0],[0,0,0]], Mb = [[2b,2b,b],[2b,2b,b],[b,b,0]], Mc = [[2c,c,c],[c,0,0],[c,0,0]]
    o[0] = T(2) * (a * ka + b * kb + c * kc);
    o[1] = a * ka + T(2) * b * kb + c * kc;
    o[2] = b * kb + c * kc;
    o[3] = T(2) * b * kb;
    o[4] = b * kb;
    o[5] = T(0);
  }
};
template <typename T>
__device__ __forceinline__ void sym3_mul(const T* A, const T* x, T* y) {
  y[0] = A[0] * x[0] + A[1] * x[1] + A[2] * x[2];
  y[1] = A[1] * x[0] + A[3] * x[1] + A[4] * x[2];
  y[2] = A[2] * x[0] + A[4] * x[1] + A[5] * x[2];
}
template <typename T>
__device__ __forceinline__ void sym3_inv(const T* A, T* I) {
  T c00 = A[3] * A[5] - A[4] * A[4], c01 = A[2] * A[4] - A[1] * A[5], c02 = A[1] * A[4] - A[2] * A[3];
  T det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  T id = T(1) / det;
  I[0] = c00 * id; I[1] = c01 * id; I[2] = c02 * id;
  I[3] = (A[0] * A[5] - A[2] * A[2]) * id;
  I[4] = (A[1] * A[2] - A[0] * A[4]) * id;
  I[5] = (A[0] * A[3] - A[1] * A[1]) * id;
}

// Forward dynamics of the planar arm.  Optionally returns Minv (sym6) and the pieces the Jacobian needs.
template <typename T>
struct Planar3RState {
  T M[6], Mi[6], D2[6], D3[6], h[3], acc[3];
  T s2, c2, s3, c3, s23, c23;
};
template <typename T>
__device__ __forceinline__ void planar3r_forward(const Planar3R<T>& R, const T* q, const T* v, const T* u, Planar3RState<T>& st) {
  sincos_(q[1], st.s2, st.c2);
  sincos_(q[2], st.s3, st.c3);
  st.s23 = st.s2 * st.c3 + st.c2 * st.s3;
  st.c23 = st.c2 * st.c3 - st.s2 * st.s3;
  R.pat(st.c2, st.c3, st.c23, st.M);
  for (int k = 0; k < 6; ++k) st.M[k] += R.m0[k];
  R.pat(-st.s2, T(0), -st.s23, st.D2);      // dM/dq2
  R.pat(T(0), -st.s3, -st.s23, st.D3);      // dM/dq3
  // h = Mdot v - 0.5 [0, v'D2 v, v'D3 v],  Mdot = D2 v2 + D3 v3
  T d2v[3], d3v[3];
  sym3_mul(st.D2, v, d2v);
  sym3_mul(st.D3, v, d3v);
  st.h[0] = d2v[0] * v[1] + d3v[0] * v[2];
  st.h[1] = d2v[1] * v[1] + d3v[1] * v[2] - T(0.5) * (v[0] * d2v[0] + v[1] * d2v[1] + v[2] * d2v[2]);
  st.h[2] = d2v[2] * v[1] + d3v[2] * v[2] - T(0.5) * (v[0] * d3v[0] + v[1] * d3v[1] + v[2] * d3v[2]);
  sym3_inv(st.M, st.Mi);
  T rhs[3] = {u[0] - st.h[0], u[1] - st.h[1], u[2] - st.h[2]};
  sym3_mul(st.Mi, rhs, st.acc);
}

// ------------------------------------------------------------------------------------------ per-system API
template <int SYS> struct SysDims;
template <> struct SysDims<CACTO_SINGLE_INTEGRATOR> { static constexpr int NX = 2, NA = 2; };
template <> struct SysDims<CACTO_DOUBLE_INTEGRATOR> { static constexpr int NX = 4, NA = 2; };
template <> struct SysDims<CACTO_CAR> { static constexpr int NX = 5, NA = 2; };
template <> struct SysDims<CACTO_CAR_PARK> { static constexpr int NX = 5, NA = 2; };
template <> struct SysDims<CACTO_MANIPULATOR> { static constexpr int NX = 6, NA = 3; };
template <> struct SysDims<CACTO_UR5> { static constexpr int NX = 12, NA = 6; };

// UR5 helpers: mass matrix (row-major 6x6) and nle.
template <typename T>
__device__ __forceinline__ void ur5_sincos(const T* q, T* sc) {
  for (int i = 0; i < 6; ++i) sincos_(q[i], sc[i], sc[6 + i]);
}
// Composite-rigid-body algorithm in link frames (3-vector form, same frame conventions as chain_rnea): composite inertias
// (mass, first moment h = m c, rotational inertia about the frame origin) from the tip to the base, then for joint i the
// wrench (n, f) = (I_i e_i, e_i x h_i) of a unit joint acceleration, carried down the chain: M[i][j] = e_j . n in frame j.
// A third of the arithmetic of the six velocity-free RNEA passes it replaces (each of which re-derived every link's motion).
template <typename T>
__device__ void ur5_mass_matrix(const cacto_chain& ch, const T* q, T* M, const T* sc) {
  (void)q;
  T R[6][9];                      // R_i: frame i -> frame i - 1 (= Rfix_i Rot(axis_i, q_i))
  T cm[6], chh[6][3], cI[6][6];   // composite mass, first moment, inertia (xx, yy, zz, xy, xz, yz) about the frame origin
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    JointRot<T> J;
    for (int k = 0; k < 9; ++k) J.F[k] = T(ch.R[i][k]);
    J.axis = ch.axis[i];
    J.s = sc[i]; J.c = sc[6 + i];
    for (int r = 0; r < 3; ++r) J.rot_axis_T(&J.F[3 * r], &R[i][3 * r]);          // row r of Rfix times Rot
    const T m = T(ch.mass[i]), cx = T(ch.com[i][0]), cy = T(ch.com[i][1]), cz = T(ch.com[i][2]);
    cm[i] = m;
    chh[i][0] = m * cx; chh[i][1] = m * cy; chh[i][2] = m * cz;
    cI[i][0] = T(ch.inertia[i][0]) + m * (cy * cy + cz * cz);
    cI[i][1] = T(ch.inertia[i][1]) + m * (cx * cx + cz * cz);
    cI[i][2] = T(ch.inertia[i][2]) + m * (cx * cx + cy * cy);
    cI[i][3] = T(ch.inertia[i][3]) - m * cx * cy;
    cI[i][4] = T(ch.inertia[i][4]) - m * cx * cz;
    cI[i][5] = T(ch.inertia[i][5]) - m * cy * cz;
  }
#pragma unroll
  for (int i = 5; i >= 1; --i) {  // composite of link i (frame i) added to link i - 1 (frame i - 1)
    const T* Ri = R[i];
    const T p0 = T(ch.p[i][0]), p1 = T(ch.p[i][1]), p2 = T(ch.p[i][2]), m = cm[i];
    T h[3];
    for (int r = 0; r < 3; ++r) h[r] = Ri[3 * r] * chh[i][0] + Ri[3 * r + 1] * chh[i][1] + Ri[3 * r + 2] * chh[i][2];
    // A = I R^T (I symmetric), then Ir = R A
    const T ixx = cI[i][0], iyy = cI[i][1], izz = cI[i][2], ixy = cI[i][3], ixz = cI[i][4], iyz = cI[i][5];
    T A[3][3];
    for (int c = 0; c < 3; ++c) {
      A[0][c] = ixx * Ri[3 * c] + ixy * Ri[3 * c + 1] + ixz * Ri[3 * c + 2];
      A[1][c] = ixy * Ri[3 * c] + iyy * Ri[3 * c + 1] + iyz * Ri[3 * c + 2];
      A[2][c] = ixz * Ri[3 * c] + iyz * Ri[3 * c + 1] + izz * Ri[3 * c + 2];
    }
    auto rAr = [&](int r, int c) { return Ri[3 * r] * A[0][c] + Ri[3 * r + 1] * A[1][c] + Ri[3 * r + 2] * A[2][c]; };
    const T pp = p0 * p0 + p1 * p1 + p2 * p2, ph = p0 * h[0] + p1 * h[1] + p2 * h[2];
    const T dg = m * pp + T(2) * ph;                       // the isotropic part of the origin shift
    cm[i - 1] += m;
    chh[i - 1][0] += h[0] + m * p0; chh[i - 1][1] += h[1] + m * p1; chh[i - 1][2] += h[2] + m * p2;
    cI[i - 1][0] += rAr(0, 0) + dg - m * p0 * p0 - T(2) * p0 * h[0];
    cI[i - 1][1] += rAr(1, 1) + dg - m * p1 * p1 - T(2) * p1 * h[1];
    cI[i - 1][2] += rAr(2, 2) + dg - m * p2 * p2 - T(2) * p2 * h[2];
    cI[i - 1][3] += rAr(0, 1) - m * p0 * p1 - (p0 * h[1] + h[0] * p1);
    cI[i - 1][4] += rAr(0, 2) - m * p0 * p2 - (p0 * h[2] + h[0] * p2);
    cI[i - 1][5] += rAr(1, 2) - m * p1 * p2 - (p1 * h[2] + h[1] * p2);
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const int ax = ch.axis[i];
    T n[3], f[3];
    // n = I_i e, f = e x h_i
    if (ax == 0) { n[0] = cI[i][0]; n[1] = cI[i][3]; n[2] = cI[i][4]; f[0] = T(0); f[1] = -chh[i][2]; f[2] = chh[i][1]; }
    else if (ax == 1) { n[0] = cI[i][3]; n[1] = cI[i][1]; n[2] = cI[i][5]; f[0] = chh[i][2]; f[1] = T(0); f[2] = -chh[i][0]; }
    else { n[0] = cI[i][4]; n[1] = cI[i][5]; n[2] = cI[i][2]; f[0] = -chh[i][1]; f[1] = chh[i][0]; f[2] = T(0); }
    M[i * 6 + i] = axis_get(n, ax);
#pragma unroll
    for (int j = i; j >= 1; --j) {                         // (n, f) from frame j to frame j - 1
      const T* Rj = R[j];
      T fp[3], np[3], pj[3] = {T(ch.p[j][0]), T(ch.p[j][1]), T(ch.p[j][2])}, t[3];
      for (int r = 0; r < 3; ++r) {
        fp[r] = Rj[3 * r] * f[0] + Rj[3 * r + 1] * f[1] + Rj[3 * r + 2] * f[2];
        np[r] = Rj[3 * r] * n[0] + Rj[3 * r + 1] * n[1] + Rj[3 * r + 2] * n[2];
      }
      cross3(pj, fp, t);
      for (int r = 0; r < 3; ++r) { f[r] = fp[r]; n[r] = np[r] + t[r]; }
      const T mij = axis_get(n, (int)ch.axis[j - 1]);
      M[i * 6 + (j - 1)] = mij;
      M[(j - 1) * 6 + i] = mij;
    }
  }
}

// x' = f(x, u) without the time component.  environment.py:235-243 (SI), :80-91 + robot_utils.py:399-405
// (DI / manipulator / UR5), :437-448 (car), :584-595 (car_park).
template <int SYS, typename T>
__device__ __forceinline__ void sys_step(const cacto_sys_params& P, const T* x, const T* u, T* xn) {
  const T dt = T(P.dt);
  if (SYS == CACTO_SINGLE_INTEGRATOR) {
    xn[0] = x[0] + dt * u[0];
    xn[1] = x[1] + dt * u[1];
  } else if (SYS == CACTO_DOUBLE_INTEGRATOR) {
    const T m1 = T(P.chain.mass[0] + P.chain.mass[1]), m2 = T(P.chain.mass[1]);
    xn[0] = x[0] + x[2] * dt;
    xn[1] = x[1] + x[3] * dt;
    xn[2] = x[2] + (u[0] / m1) * dt;
    xn[3] = x[3] + (u[1] / m2) * dt;
  } else if (SYS == CACTO_CAR) {
    T s, c;
    sincos_(x[2], s, c);
    xn[0] = x[0] + dt * x[3] * c + dt * dt * x[4] * c / T(2);
    xn[1] = x[1] + dt * x[3] * s + dt * dt * x[4] * s / T(2);
    xn[2] = x[2] + dt * u[0];
    xn[3] = x[3] + dt * x[4];
    xn[4] = x[4] + dt * u[1];
  } else if (SYS == CACTO_CAR_PARK) {
    T s, c;
    sincos_(x[2], s, c);
    xn[0] = x[0] + dt * x[3] * c;
    xn[1] = x[1] + dt * x[3] * s;
    xn[2] = x[2] + dt * x[3] * tan_(x[4]) / T(P.L_delta);
    xn[3] = x[3] + dt * u[0];
    xn[4] = x[4] + dt * u[1] / T(P.tau_delta);
  } else if (SYS == CACTO_MANIPULATOR) {
    Planar3R<T> R(P.chain);
    Planar3RState<T> st;
    planar3r_forward(R, x, x + 3, u, st);
    for (int k = 0; k < 3; ++k) {
      xn[k] = x[k] + x[3 + k] * dt;
      xn[3 + k] = x[3 + k] + st.acc[k] * dt;
    }
  } else {
    T M[36], rhs[6], z[6] = {0, 0, 0, 0, 0, 0}, sc[12];
    ur5_sincos<T>(x, sc);
    ur5_mass_matrix<T>(P.chain, x, M, sc);
    chain_rnea<T, 6, true>(P.chain, x, x + 6, z, T(P.chain.gravity), rhs, sc);
    for (int k = 0; k < 6; ++k) rhs[k] = u[k] - rhs[k];
    chol_factor<T, 6>(M);
    chol_solve<T, 6>(M, rhs);
    for (int k = 0; k < 6; ++k) {
      xn[k] = x[k] + x[6 + k] * dt;
      xn[6 + k] = x[6 + k] + rhs[k] * dt;
    }
  }
}

// dt * d x'_{vel} / du  =  dt * Minv for the chains, the constant pattern for the analytic systems.
// Fu is (NX x NA) row-major, un-normalised, no time row.  environment.py:93-109,:209,:408,:555.
template <int SYS, typename T>
__device__ __forceinline__ void sys_Fu(const cacto_sys_params& P, const T* x, T* Fu) {
  constexpr int NX = SysDims<SYS>::NX, NA = SysDims<SYS>::NA;
  const T dt = T(P.dt);
  for (int k = 0; k < NX * NA; ++k) Fu[k] = T(0);
  if (SYS == CACTO_SINGLE_INTEGRATOR) {
    Fu[0 * NA + 0] = dt; Fu[1 * NA + 1] = dt;
  } else if (SYS == CACTO_DOUBLE_INTEGRATOR) {
    Fu[2 * NA + 0] = dt / T(P.chain.mass[0] + P.chain.mass[1]);
    Fu[3 * NA + 1] = dt / T(P.chain.mass[1]);
  } else if (SYS == CACTO_CAR) {
    Fu[2 * NA + 0] = dt; Fu[4 * NA + 1] = dt;
  } else if (SYS == CACTO_CAR_PARK) {
    Fu[3 * NA + 0] = dt; Fu[4 * NA + 1] = dt / T(P.tau_delta);
  } else if (SYS == CACTO_MANIPULATOR) {
    Planar3R<T> R(P.chain);
    T s2, c2, s3, c3;
    sincos_(x[1], s2, c2);
    sincos_(x[2], s3, c3);
    T M[6], Mi[6];
    R.pat(c2, c3, c2 * c3 - s2 * s3, M);
    for (int k = 0; k < 6; ++k) M[k] += R.m0[k];
    sym3_inv(M, Mi);
    const int ix[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Fu[(3 + i) * NA + j] = dt * Mi[ix[i][j]];
  } else {
    T M[36], sc[12];
    ur5_sincos<T>(x, sc);
    ur5_mass_matrix<T>(P.chain, x, M, sc);
    chol_factor<T, 6>(M);
#pragma unroll 1
    for (int j = 0; j < 6; ++j) {
      T e[6] = {0, 0, 0, 0, 0, 0};
      e[j] = T(1);
      chol_solve<T, 6>(M, e);
      for (int i = 0; i < 6; ++i) Fu[(6 + i) * NA + j] = dt * e[i];
    }
  }
}

// x' and Fu at the same (x, u) -- what the actor update needs per sample (NeuralNetwork.py:185-197).  For the UR5 both come from ONE
// mass matrix and ONE Cholesky factor (calling sys_step and sys_Fu rebuilt and refactored M); identical values either way.
template <int SYS, typename T>
__device__ __forceinline__ void sys_step_Fu(const cacto_sys_params& P, const T* x, const T* u, T* xn, T* Fu) {
  if (SYS == CACTO_MANIPULATOR) {                          // the forward dynamics already hold Minv (and the two sincos)
    constexpr int NX = SysDims<SYS>::NX, NA = SysDims<SYS>::NA;
    const T dt = T(P.dt);
    Planar3R<T> R(P.chain);
    Planar3RState<T> st;
    planar3r_forward(R, x, x + 3, u, st);
    for (int k = 0; k < 3; ++k) {
      xn[k] = x[k] + x[3 + k] * dt;
      xn[3 + k] = x[3 + k] + st.acc[k] * dt;
    }
    for (int k = 0; k < NX * NA; ++k) Fu[k] = T(0);
    const int ix[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Fu[(3 + i) * NA + j] = dt * st.Mi[ix[i][j]];
  } else if (SYS != CACTO_UR5) {
    sys_step<SYS, T>(P, x, u, xn);
    sys_Fu<SYS, T>(P, x, Fu);
  } else {
    constexpr int NX = SysDims<SYS>::NX, NA = SysDims<SYS>::NA;
    const T dt = T(P.dt);
    T M[36], rhs[6], z[6] = {0, 0, 0, 0, 0, 0}, sc[12];
    ur5_sincos<T>(x, sc);
    ur5_mass_matrix<T>(P.chain, x, M, sc);
    chain_rnea<T, 6, true>(P.chain, x, x + 6, z, T(P.chain.gravity), rhs, sc);
    for (int k = 0; k < 6; ++k) rhs[k] = u[k] - rhs[k];
    chol_factor<T, 6>(M);
    chol_solve<T, 6>(M, rhs);
    for (int k = 0; k < 6; ++k) {
      xn[k] = x[k] + x[6 + k] * dt;
      xn[6 + k] = x[6 + k] + rhs[k] * dt;
    }
    for (int k = 0; k < NX * NA; ++k) Fu[k] = T(0);
#pragma unroll 1
    for (int j = 0; j < 6; ++j) {
      T e[6] = {0, 0, 0, 0, 0, 0};
      e[j] = T(1);
      chol_solve<T, 6>(M, e);
      for (int i = 0; i < 6; ++i) Fu[(6 + i) * NA + j] = dt * e[i];
    }
  }
}

// Discrete-time Jacobians Fx (NX x NX), Fu (NX x NA).  environment.py:111-132,:221-233,:420-435,:567-582.
template <int SYS, typename T>
__device__ __forceinline__ void sys_jac(const cacto_sys_params& P, const T* x, const T* u, T* Fx, T* Fu) {
  constexpr int NX = SysDims<SYS>::NX, NA = SysDims<SYS>::NA;
  const T dt = T(P.dt);
  for (int k = 0; k < NX * NX; ++k) Fx[k] = T(0);
  for (int k = 0; k < NX; ++k) Fx[k * NX + k] = T(1);
  if (SYS == CACTO_SINGLE_INTEGRATOR) {
    sys_Fu<SYS, T>(P, x, Fu);
  } else if (SYS == CACTO_DOUBLE_INTEGRATOR) {
    sys_Fu<SYS, T>(P, x, Fu);
    Fx[0 * NX + 2] = dt; Fx[1 * NX + 3] = dt;
  } else if (SYS == CACTO_CAR) {
    sys_Fu<SYS, T>(P, x, Fu);
    T s, c;
    sincos_(x[2], s, c);
    Fx[0 * NX + 2] = -dt * x[3] * s - dt * dt * x[4] * s / T(2);
    Fx[0 * NX + 3] = dt * c;
    Fx[0 * NX + 4] = dt * dt * c / T(2);
    Fx[1 * NX + 2] = dt * x[3] * c + dt * dt * x[4] * c / T(2);
    Fx[1 * NX + 3] = dt * s;
    Fx[1 * NX + 4] = dt * dt * s / T(2);
    Fx[3 * NX + 4] = dt;
  } else if (SYS == CACTO_CAR_PARK) {
    sys_Fu<SYS, T>(P, x, Fu);
    T s, c, sd, cd;
    sincos_(x[2], s, c);
    sincos_(x[4], sd, cd);
    const T Ld = T(P.L_delta);
    Fx[0 * NX + 2] = -dt * x[3] * s;
    Fx[0 * NX + 3] = dt * c;
    Fx[1 * NX + 2] = dt * x[3] * c;
    Fx[1 * NX + 3] = dt * s;
    Fx[2 * NX + 3] = dt * (sd / cd) / Ld;
    Fx[2 * NX + 4] = dt * x[3] * (T(1) / (cd * cd)) / Ld;
  } else if (SYS == CACTO_MANIPULATOR) {
    // a = Minv (u - h);  da/dq_k = -Minv (D_k a + dh/dq_k),  da/dv = -Minv dh/dv
    Planar3R<T> R(P.chain);
    Planar3RState<T> st;
    const T* v = x + 3;
    planar3r_forward(R, x, v, u, st);
    const int ix[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
    T E22[6], E23[6], E33[6];
    R.pat(-st.c2, T(0), -st.c23, E22);
    R.pat(T(0), T(0), -st.c23, E23);
    R.pat(T(0), -st.c3, -st.c23, E33);
    T col[3], rhs[3], t2[3], t3[3];
    for (int k = 0; k < 3; ++k) Fx[k * NX + 3 + k] = dt;
    // d/dq1 = 0;  d/dq2, d/dq3
    for (int kq = 1; kq < 3; ++kq) {
      const T* Dk = kq == 1 ? st.D2 : st.D3;
      const T* E2k = kq == 1 ? E22 : E23;
      const T* E3k = kq == 1 ? E23 : E33;
      sym3_mul(E2k, v, t2);
      sym3_mul(E3k, v, t3);
      sym3_mul(Dk, st.acc, rhs);
      rhs[0] += t2[0] * v[1] + t3[0] * v[2];
      rhs[1] += t2[1] * v[1] + t3[1] * v[2] - T(0.5) * (v[0] * t2[0] + v[1] * t2[1] + v[2] * t2[2]);
      rhs[2] += t2[2] * v[1] + t3[2] * v[2] - T(0.5) * (v[0] * t3[0] + v[1] * t3[1] + v[2] * t3[2]);
      sym3_mul(st.Mi, rhs, col);
      for (int i = 0; i < 3; ++i) Fx[(3 + i) * NX + kq] = -dt * col[i];
    }
    // dh/dv_j: (D2 v2 + D3 v3)_{ij} + [j==1](D2 v)_i + [j==2](D3 v)_i - [i==1](D2 v)_j - [i==2](D3 v)_j
    sym3_mul(st.D2, v, t2);
    sym3_mul(st.D3, v, t3);
    for (int j = 0; j < 3; ++j) {
      for (int i = 0; i < 3; ++i) {
        T d = st.D2[ix[i][j]] * v[1] + st.D3[ix[i][j]] * v[2];
        if (j == 1) d += t2[i];
        if (j == 2) d += t3[i];
        if (i == 1) d -= t2[j];
        if (i == 2) d -= t3[j];
        rhs[i] = d;
      }
      sym3_mul(st.Mi, rhs, col);
      for (int i = 0; i < 3; ++i) Fx[(3 + i) * NX + 3 + j] += -dt * col[i];
    }
    for (int k = 0; k < NX * NA; ++k) Fu[k] = T(0);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Fu[(3 + i) * NA + j] = dt * st.Mi[ix[i][j]];
  } else {
    // UR5: a from the real pass, then tangent-mode RNEA at fixed a for d tau/dq_j and d tau/dv_j:
    // da/dz = -Minv d tau/dz.
    T L[36], acc[6], z[6] = {0, 0, 0, 0, 0, 0}, sc[12];
    const T g = T(P.chain.gravity);
    ur5_sincos<T>(x, sc);
    ur5_mass_matrix<T>(P.chain, x, L, sc);
    chain_rnea<T, 6, true>(P.chain, x, x + 6, z, g, acc, sc);
    for (int k = 0; k < 6; ++k) acc[k] = u[k] - acc[k];
    chol_factor<T, 6>(L);
    chol_solve<T, 6>(L, acc);
    for (int k = 0; k < 6; ++k) Fx[k * NX + 6 + k] = dt;
    typedef Dual<T> D;
#pragma unroll 1
    for (int j = 0; j < 12; ++j) {
      D qd[6], vd[6], ad[6], td[6];
      for (int k = 0; k < 6; ++k) { qd[k] = D(x[k]); vd[k] = D(x[6 + k]); ad[k] = D(acc[k]); }
      if (j < 6) qd[j].d = T(1); else vd[j - 6].d = T(1);
      D scd[12];                                   // sin / cos of the dual angles from the real ones: (s, c dq), (c, -s dq)
      for (int k = 0; k < 6; ++k) { scd[k] = D(sc[k], sc[6 + k] * qd[k].d); scd[6 + k] = D(sc[6 + k], -sc[k] * qd[k].d); }
      chain_rnea<D, 6, true>(P.chain, qd, vd, ad, g, td, scd);
      T col[6];
      for (int k = 0; k < 6; ++k) col[k] = td[k].d;
      chol_solve<T, 6>(L, col);
      for (int i = 0; i < 6; ++i) Fx[(6 + i) * NX + j] += -dt * col[i];
    }
    for (int k = 0; k < NX * NA; ++k) Fu[k] = T(0);
#pragma unroll 1
    for (int j = 0; j < 6; ++j) {
      T e[6] = {0, 0, 0, 0, 0, 0};
      e[j] = T(1);
      chol_solve<T, 6>(L, e);
      for (int i = 0; i < 6; ++i) Fu[(6 + i) * NA + j] = dt * e[i];
    }
  }
}

// End-effector position.  environment.py:146-156 (Pinocchio frame 'EE'), :245-250, :450-455, :597-602.
// sc_real (UR5 only, optional): sin q_0 .. sin q_5, cos q_0 .. cos q_5 of the REAL joint angles, for callers that evaluate the
// same configuration many times with different dual parts (the 78 hyper-dual reward evaluations per knot of the TO backward pass).
template <int SYS, typename T>
__device__ __forceinline__ void sys_ee(const cacto_sys_params& P, const T* x, T* p, const typename scalar_of<T>::type* sc_real = nullptr) {
  if (SYS == CACTO_SINGLE_INTEGRATOR || SYS == CACTO_CAR) {
    p[0] = x[0]; p[1] = x[1]; p[2] = T(0);
  } else if (SYS == CACTO_DOUBLE_INTEGRATOR) {
    p[0] = x[0] + T(P.chain.p[0][0] + P.chain.p[1][0] + P.chain.ee_p[0]);
    p[1] = x[1] + T(P.chain.p[0][1] + P.chain.p[1][1] + P.chain.ee_p[1]);
    p[2] = T(P.chain.p[0][2] + P.chain.p[1][2] + P.chain.ee_p[2]);
  } else if (SYS == CACTO_CAR_PARK) {
    T s, c;
    sincos_(x[2], s, c);
    p[0] = x[0] + c * T(P.L_delta / 2);
    p[1] = x[1] + s * T(P.L_delta / 2);
    p[2] = T(0);
  } else if (SYS == CACTO_MANIPULATOR) {
    T s1, c1, s12, c12, s123, c123;
    sincos_(x[0], s1, c1);
    sincos_(x[0] + x[1], s12, c12);
    sincos_(x[0] + x[1] + x[2], s123, c123);
    const T l1 = T(P.chain.p[1][0]), l2 = T(P.chain.p[2][0]), l3 = T(P.chain.ee_p[0]);
    p[0] = T(P.chain.p[0][0]) + l1 * c1 + l2 * c12 + l3 * c123;
    p[1] = T(P.chain.p[0][1]) + l1 * s1 + l2 * s12 + l3 * s123;
    p[2] = T(P.chain.p[0][2]);
  } else {
    // fold from the tip: p <- p_i + R_i p
    T v[3] = {T(P.chain.ee_p[0]), T(P.chain.ee_p[1]), T(P.chain.ee_p[2])};
#pragma unroll
    for (int i = 5; i >= 0; --i) {
      JointRot<T> J;
      for (int k = 0; k < 9; ++k) J.F[k] = typename scalar_of<T>::type(P.chain.R[i][k]);
      J.axis = P.chain.axis[i];
      if (sc_real != nullptr) sincos_from(x[i], sc_real[i], sc_real[6 + i], J.s, J.c);
      else sincos_(x[i], J.s, J.c);
      T t[3];
      J.apply(v, t);
      for (int k = 0; k < 3; ++k) v[k] = t[k] + T(P.chain.p[i][k]);
    }
    p[0] = v[0]; p[1] = v[1]; p[2] = v[2];
  }
}

// car_park rectangle cost (environment.py:604-613), written as the product of four smooth steps
// sigma(z) = 0.5 (1 + z / sqrt(1 + z^2)) -- algebraically identical to the reference's
// eight-factor expression.
template <typename T>
__device__ __forceinline__ T smooth_step(T z) { return T(0.5) * (T(1) + z / sqrt_(T(1) + z * z)); }
template <typename T>
__device__ __forceinline__ T park_obs(T x, T y, T xc, T yc, T Wx, T Wy, T k) {
  return smooth_step((y - yc + Wy / T(2)) * k) * (T(1) - smooth_step((y - yc - Wy / T(2)) * k)) *
         smooth_step((x - xc + Wx / T(2)) * k) * (T(1) - smooth_step((x - xc - Wx / T(2)) * k));
}

// Reward r(w, s, a).  environment.py:252-275 (SI), :329-351 (DI), :457-480 (car), :615-641 (car_park),
// :695-723 (manipulator), :780-805 (UR5); bound_control_cost :158-163.  `u` may be nullptr.
template <int SYS, typename T>
__device__ __forceinline__ T sys_reward(const cacto_sys_params& P, const double* w, const T* x, const T* u, bool plain_ucost,
                                        const typename scalar_of<T>::type* sc_real = nullptr) {
  constexpr int NX = SysDims<SYS>::NX, NA = SysDims<SYS>::NA;
  T p[3];
  sys_ee<SYS, T>(P, x, p, sc_real);
  const T alpha = T(P.alpha), alpha2 = T(P.alpha2);
  constexpr int DIMS = (SYS == CACTO_UR5) ? 3 : 2;
  T pk = T(0), dist = T(0);
  const T sq01 = sqrt_(T(0.1));
  for (int i = 0; i < DIMS; ++i) {
    T d = p[i] - T(P.target[i]);
    pk += sqrt_(d * d + T(0.1)) - sq01 - T(0.1);
    dist += d * d;
  }
  T peak = softplus_(-alpha2 * pk) / alpha2;
  T u_cost = T(0);
  if (u != nullptr) {
    for (int i = 0; i < NA; ++i) {
      u_cost += u[i] * u[i];
      if (!plain_ucost) u_cost += T(P.w_b) * pow10_(u[i] / T(P.u_max[i]));
    }
  }
  T r;
  if (SYS == CACTO_CAR_PARK) {
    T s, c;
    sincos_(x[2], s, c);
    T obs = T(0);
    for (int k = 0; k < 3; ++k) {
      const T xc = T(P.obs[2 * k]), yc = T(P.obs[2 * k + 1]), Wx = T(P.obs[6 + 2 * k]), Wy = T(P.obs[7 + 2 * k]);
      for (int j = 0; j < 10; ++j) {
        const T bx = T(P.check_points[2 * j]), by = T(P.check_points[2 * j + 1]);
        obs += park_obs<T>(c * bx - s * by + p[0], s * bx + c * by + p[1], xc, yc, Wx, Wy, T(P.k_db));
      }
    }
    r = -T(w[0]) * dist + T(w[1]) * peak - T(w[2]) * x[3] * x[3] - T(w[3]) * obs - T(w[6]) * u_cost + T(P.offset);
  } else {
    T ell[3];
    for (int k = 0; k < 3; ++k) {
      T e;
      if (SYS == CACTO_UR5) {
        T dx = p[0] - T(P.obs[3 * k]), dy = p[1] - T(P.obs[3 * k + 1]), dz = p[2] - T(P.obs[3 * k + 2]);
        T A = T(P.obs[9 + 3 * k]) / T(2), B = T(P.obs[10 + 3 * k]) / T(2), C = T(P.obs[11 + 3 * k]) / T(2);
        e = dx * dx / (A * A) + dy * dy / (B * B) + dz * dz / (C * C) - T(1);
      } else {
        T dx = p[0] - T(P.obs[2 * k]), dy = p[1] - T(P.obs[2 * k + 1]);
        T A = T(P.obs[6 + 2 * k]) / T(2), B = T(P.obs[7 + 2 * k]) / T(2);
        e = dx * dx / (A * A) + dy * dy / (B * B) - T(1);
      }
      ell[k] = softplus_(-alpha * e) / alpha;
    }
    T vel = T(0);
    if (SYS == CACTO_UR5 || (SYS == CACTO_MANIPULATOR && w[2] != 0.0)) {
      for (int i = NX / 2; i < NX; ++i) vel += x[i] * x[i];
    }
    r = -T(w[0]) * dist + T(w[1]) * peak - T(w[2]) * vel - T(w[3]) * ell[0] - T(w[4]) * ell[1] - T(w[5]) * ell[2] -
        T(w[6]) * u_cost + T(P.offset);
  }
  return T(P.scale) * r;
}

// d reward_batch / d action (NeuralNetwork.py:199-204 through environment.py:282-284):
// -scale * w6 * (2 a + 10 w_b a^9 / u_max^10)
template <int SYS, typename T>
__device__ __forceinline__ void sys_dr_da(const cacto_sys_params& P, T w6, const T* u, T* g) {
  constexpr int NA = SysDims<SYS>::NA;
  for (int i = 0; i < NA; ++i) {
    const T um = T(P.u_max[i]);
    const T r = u[i] / um;
    const T r2 = r * r, r4 = r2 * r2, r8 = r4 * r4;
    g[i] = -T(P.scale) * w6 * (T(2) * u[i] + T(10) * T(P.w_b) * (r8 * r) / um);
  }
}

}  // namespace cacto
