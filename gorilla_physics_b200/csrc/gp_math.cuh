// gp_math.cuh — 3-vectors, 3x3 rotations and spatial (Pluecker) transforms in f64.
// Device-side replacements for the nalgebra operations the reference leans on
// (SURVEY.md §8c table); everything is expressed in body coordinates.
#pragma once
#include <cmath>

#include "gp_topology.cuh"

namespace gp {

// ---- reciprocal and reciprocal square root without the library's special-case ladder ---------
// The hardware seed (MUFU.RCP64H / MUFU.RSQ64H through rcp/rsqrt.approx.ftz.f64, relative error
// about 2^-23) plus two Newton steps: ~1 ulp, 5-7 FP64 instructions, no branch and no out-of-line
// slow path (`1.0 / x` and sqrt(x) cost 12-28 instructions and a call each). Arguments here are
// positive and far from the ends of the exponent range (pivots of the mass matrix, penetration depths,
// squared sliding speeds); a zero, negative or non-finite argument still produces inf/NaN, which the
// callers flag like the division would.
GP_HD double gp_rcp(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  y = fma(y, fma(-x, y, 1.0), y);
  y = fma(y, fma(-x, y, 1.0), y);
  return y;
#else
  return 1.0 / x;
#endif
}
GP_HD double gp_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  y = fma(y, fma(-(h * y), y, 0.5), y);
  y = fma(y, fma(-(h * y), y, 0.5), y);
  return y;
#else
  return 1.0 / std::sqrt(x);
#endif
}

struct V3 {
  double x, y, z;
};
GP_HD V3 v3(double x, double y, double z) { return V3{x, y, z}; }
GP_HD V3 v3z() { return V3{0.0, 0.0, 0.0}; }
GP_HD V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
GP_HD V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
GP_HD V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
GP_HD V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
GP_HD V3 operator*(double s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
GP_HD V3& operator+=(V3& a, V3 b) { a.x += b.x; a.y += b.y; a.z += b.z; return a; }
GP_HD V3& operator-=(V3& a, V3 b) { a.x -= b.x; a.y -= b.y; a.z -= b.z; return a; }
GP_HD double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GP_HD V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
GP_HD V3 ld3(const double* p) { return {p[0], p[1], p[2]}; }
// acc + a x b as two fused multiply-adds per component (the plain `acc + cross(a, b)` costs a
// multiply, an FMA and an add: the compiler may not re-associate)
GP_HD V3 cross_add(V3 acc, V3 a, V3 b) {
  return {fma(a.y, b.z, fma(-a.z, b.y, acc.x)), fma(a.z, b.x, fma(-a.x, b.z, acc.y)),
          fma(a.x, b.y, fma(-a.y, b.x, acc.z))};
}
// acc - a x b
GP_HD V3 cross_sub(V3 acc, V3 a, V3 b) {
  return {fma(a.z, b.y, fma(-a.y, b.z, acc.x)), fma(a.x, b.z, fma(-a.z, b.x, acc.y)),
          fma(a.y, b.x, fma(-a.x, b.y, acc.z))};
}
// a * s + acc
GP_HD V3 fma3(V3 a, double s, V3 acc) { return {fma(a.x, s, acc.x), fma(a.y, s, acc.y), fma(a.z, s, acc.z)}; }
// acc + a . b
GP_HD double dot_add(double acc, V3 a, V3 b) { return fma(a.z, b.z, fma(a.y, b.y, fma(a.x, b.x, acc))); }

// rotation / general 3x3, row-major
struct M3 {
  double m[9];
};
GP_HD V3 mul(const M3& E, V3 v) {
  return {E.m[0] * v.x + E.m[1] * v.y + E.m[2] * v.z, E.m[3] * v.x + E.m[4] * v.y + E.m[5] * v.z,
          E.m[6] * v.x + E.m[7] * v.y + E.m[8] * v.z};
}
GP_HD V3 mulT(const M3& E, V3 v) {
  return {E.m[0] * v.x + E.m[3] * v.y + E.m[6] * v.z, E.m[1] * v.x + E.m[4] * v.y + E.m[7] * v.z,
          E.m[2] * v.x + E.m[5] * v.y + E.m[8] * v.z};
}
// acc + E v
GP_HD V3 mul_add(V3 acc, const M3& E, V3 v) {
  return {fma(E.m[2], v.z, fma(E.m[1], v.y, fma(E.m[0], v.x, acc.x))),
          fma(E.m[5], v.z, fma(E.m[4], v.y, fma(E.m[3], v.x, acc.y))),
          fma(E.m[8], v.z, fma(E.m[7], v.y, fma(E.m[6], v.x, acc.z)))};
}
GP_HD M3 mul(const M3& A, const M3& B) {
  M3 R;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      R.m[3 * i + j] = A.m[3 * i] * B.m[j] + A.m[3 * i + 1] * B.m[3 + j] + A.m[3 * i + 2] * B.m[6 + j];
  return R;
}
GP_HD M3 ldm3(const double* p) {
  M3 R;
#pragma unroll
  for (int i = 0; i < 9; ++i) R.m[i] = p[i];
  return R;
}
GP_HD M3 m3_identity() { return M3{{1, 0, 0, 0, 1, 0, 0, 0, 1}}; }

// unit quaternion (x,y,z,w) -> rotation matrix, the formula nalgebra's to_rotation_matrix uses
GP_HD M3 quat_to_mat(double i, double j, double k, double w) {
  double ww = w * w, ii = i * i, jj = j * j, kk = k * k;
  double ij = i * j * 2.0, wk = w * k * 2.0, wj = w * j * 2.0;
  double ik = i * k * 2.0, jk = j * k * 2.0, wi = w * i * 2.0;
  return M3{{ww + ii - jj - kk, ij - wk, wj + ik, wk + ij, ww - ii + jj - kk, jk - wi, ik - wj, wi + jk,
             ww - ii - jj + kk}};
}

// symmetric 3x3 stored xx,xy,xz,yy,yz,zz
struct S3 {
  double xx, xy, xz, yy, yz, zz;
};
GP_HD S3 lds3(const double* p) { return {p[0], p[1], p[2], p[3], p[4], p[5]}; }
GP_HD V3 mul(const S3& J, V3 v) {
  return {J.xx * v.x + J.xy * v.y + J.xz * v.z, J.xy * v.x + J.yy * v.y + J.yz * v.z,
          J.xz * v.x + J.yz * v.y + J.zz * v.z};
}
// acc + J v
GP_HD V3 mul_add(V3 acc, const S3& J, V3 v) {
  return {fma(J.xz, v.z, fma(J.xy, v.y, fma(J.xx, v.x, acc.x))), fma(J.yz, v.z, fma(J.yy, v.y, fma(J.xy, v.x, acc.y))),
          fma(J.zz, v.z, fma(J.yz, v.y, fma(J.xz, v.x, acc.z)))};
}
GP_HD S3 operator+(const S3& a, const S3& b) {
  return {a.xx + b.xx, a.xy + b.xy, a.xz + b.xz, a.yy + b.yy, a.yz + b.yz, a.zz + b.zz};
}

// E J E^T for symmetric J
GP_HD S3 rotate_sym(const M3& E, const S3& J) {
  // T = E J (rows of E times symmetric J)
  double t[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    double a = E.m[3 * i], b = E.m[3 * i + 1], c = E.m[3 * i + 2];
    t[3 * i + 0] = a * J.xx + b * J.xy + c * J.xz;
    t[3 * i + 1] = a * J.xy + b * J.yy + c * J.yz;
    t[3 * i + 2] = a * J.xz + b * J.yz + c * J.zz;
  }
  S3 R;
  R.xx = t[0] * E.m[0] + t[1] * E.m[1] + t[2] * E.m[2];
  R.xy = t[0] * E.m[3] + t[1] * E.m[4] + t[2] * E.m[5];
  R.xz = t[0] * E.m[6] + t[1] * E.m[7] + t[2] * E.m[8];
  R.yy = t[3] * E.m[3] + t[4] * E.m[4] + t[5] * E.m[5];
  R.yz = t[3] * E.m[6] + t[4] * E.m[7] + t[5] * E.m[8];
  R.zz = t[6] * E.m[6] + t[7] * E.m[7] + t[8] * E.m[8];
  return R;
}

// spatial vector: angular part, linear part
struct SV {
  V3 a, l;
};
GP_HD SV svz() { return SV{v3z(), v3z()}; }

// motion vector from predecessor to successor coordinates. E maps successor-frame vectors
// to the predecessor frame, r is the successor origin in predecessor coordinates.
GP_HD SV motion_to_child(const M3& E, V3 r, const SV& p) {
  return SV{mulT(E, p.a), mulT(E, cross_add(p.l, p.a, r))};
}
// the same for a motion vector whose angular part has no z component (p.a.z is a literal zero)
GP_HD SV motion_to_child_az0(const M3& E, V3 r, const SV& p) {
  const V3 a = V3{E.m[0] * p.a.x + E.m[3] * p.a.y, E.m[1] * p.a.x + E.m[4] * p.a.y, E.m[2] * p.a.x + E.m[5] * p.a.y};
  const V3 t = V3{fma(p.a.y, r.z, p.l.x), fma(-p.a.x, r.z, p.l.y), fma(p.a.x, r.y, fma(-p.a.y, r.x, p.l.z))};  // p.l + p.a x r
  return SV{a, mulT(E, t)};
}
// force vector from successor to predecessor coordinates
GP_HD SV force_to_parent(const M3& E, V3 r, const SV& f) {
  V3 fl = mul(E, f.l);
  return SV{mul_add(cross(r, fl), E, f.a), fl};
}
// the same for a force whose linear part has no z component (f.l.z is a literal zero)
GP_HD SV force_to_parent_lz0(const M3& E, V3 r, const SV& f) {
  V3 fl = V3{E.m[0] * f.l.x + E.m[1] * f.l.y, E.m[3] * f.l.x + E.m[4] * f.l.y, E.m[6] * f.l.x + E.m[7] * f.l.y};
  return SV{mul_add(cross(r, fl), E, f.a), fl};
}
// acc += (f expressed in predecessor coordinates)
GP_HD void force_acc_parent(SV& acc, const M3& E, V3 r, const SV& f) {
  V3 fl = mul(E, f.l);
  acc.a = mul_add(cross_add(acc.a, r, fl), E, f.a);
  acc.l += fl;
}

// rigid-body inertia about the frame origin: (J, c = m * com, m)
struct RBI {
  S3 J;
  V3 c;
  double m;
};
// I * (w; v) = (J w + c x v ; m v - c x w)        reference util.rs:18-28 mul_inertia
GP_HD SV mul(const RBI& I, const SV& v) {
  return SV{mul_add(cross(I.c, v.l), I.J, v.a), cross_sub(v.l * I.m, I.c, v.a)};
}
// express an inertia given in the successor frame in the predecessor frame
// (same algebra as reference inertia.rs:106-134, with Y = w r^T + r w^T, w = c' + (m/2) r)
GP_HD RBI inertia_to_parent(const M3& E, V3 r, const RBI& I) {
  V3 c1 = mul(E, I.c);
  S3 J1 = rotate_sym(E, I.J);
  V3 w = c1 + r * (0.5 * I.m);
  double tr = 2.0 * dot(w, r);
  RBI R;
  R.J.xx = J1.xx - 2.0 * w.x * r.x + tr;
  R.J.yy = J1.yy - 2.0 * w.y * r.y + tr;
  R.J.zz = J1.zz - 2.0 * w.z * r.z + tr;
  R.J.xy = J1.xy - (w.x * r.y + r.x * w.y);
  R.J.xz = J1.xz - (w.x * r.z + r.x * w.z);
  R.J.yz = J1.yz - (w.y * r.z + r.y * w.z);
  R.c = c1 + r * I.m;
  R.m = I.m;
  return R;
}


// acc += (composite inertia I of a child, expressed in the parent's coordinates).
// CONST_R (revolute / fixed joints: r and the subtree mass never change): the host has already folded
// the parallel-axis terms m (r.r 1 - r r^T) and m r into the parent's accumulator (MechParams::Jacc0,
// cacc0), r2 = 2 r comes from the constant bank, and what is left is
//   J += E J E^T - (c1 r^T + r c1^T) + 2 (c1.r) 1,   c += c1,   c1 = E c.
// Otherwise the same expression with w = c1 + (m/2) r in place of c1 covers the mass terms
// (inertia_to_parent above), and c += c1 + m r. The composite mass is a host constant either way.
template <bool CONST_R>
GP_HD void inertia_acc_parent(RBI& acc, const M3& E, V3 r, V3 r2, const RBI& I) {
  const V3 c1 = mul(E, I.c);
  V3 w = c1;
  if (!CONST_R) {
    w = V3{fma(r.x, 0.5 * I.m, c1.x), fma(r.y, 0.5 * I.m, c1.y), fma(r.z, 0.5 * I.m, c1.z)};
    r2 = r + r;
    acc.c = V3{fma(r.x, I.m, acc.c.x), fma(r.y, I.m, acc.c.y), fma(r.z, I.m, acc.c.z)};
  }
  acc.c += c1;
  double t[9];  // T = E J
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double a = E.m[3 * i], b = E.m[3 * i + 1], c = E.m[3 * i + 2];
    t[3 * i + 0] = a * I.J.xx + b * I.J.xy + c * I.J.xz;
    t[3 * i + 1] = a * I.J.xy + b * I.J.yy + c * I.J.yz;
    t[3 * i + 2] = a * I.J.xz + b * I.J.yz + c * I.J.zz;
  }
  const V3 e0 = V3{E.m[0], E.m[1], E.m[2]}, e1 = V3{E.m[3], E.m[4], E.m[5]}, e2 = V3{E.m[6], E.m[7], E.m[8]};
  const V3 t0 = V3{t[0], t[1], t[2]}, t1 = V3{t[3], t[4], t[5]}, t2 = V3{t[6], t[7], t[8]};
  // diagonal: -2 w_x r_x + 2 w.r = 2 (w_y r_y + w_z r_z)
  acc.J.xx = fma(w.z, r2.z, fma(w.y, r2.y, dot_add(acc.J.xx, t0, e0)));
  acc.J.yy = fma(w.z, r2.z, fma(w.x, r2.x, dot_add(acc.J.yy, t1, e1)));
  acc.J.zz = fma(w.y, r2.y, fma(w.x, r2.x, dot_add(acc.J.zz, t2, e2)));
  acc.J.xy = fma(-r.x, w.y, fma(-w.x, r.y, dot_add(acc.J.xy, t0, e1)));
  acc.J.xz = fma(-r.x, w.z, fma(-w.x, r.z, dot_add(acc.J.xz, t0, e2)));
  acc.J.yz = fma(-r.y, w.z, fma(-w.y, r.z, dot_add(acc.J.yz, t1, e2)));
}

}  // namespace gp
