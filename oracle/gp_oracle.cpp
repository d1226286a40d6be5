// gp_oracle.cpp — CPU f64 restatement of the gorilla-physics step path.
// TEST INFRASTRUCTURE ONLY (see gp_oracle.h). Never linked into the product library.
//
// Every function names the reference file:line it follows and keeps that code's
// operation order (world-frame quantities, world->body->world Coriolis round trip,
// duplicate inertia transform in newton_euler, partial-pivot LU) so that oracle-vs-Rust
// drift stays at rounding level. Compile with -ffp-contract=off: rustc never fuses a*b+c.
#include "gp_oracle.h"

#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

constexpr double GRAVITY = 9.81;  // lib.rs:39
constexpr double PI = 3.14159265358979323846;
constexpr double TWO_PI = 6.28318530717958647692;  // 2 PI (a literal: the counting build's scalar has no constexpr arithmetic)

// ---------------------------------------------------------------- small algebra (nalgebra 0.33.2)
struct V3 {
  double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }

struct M3 {
  double m[3][3];  // m[row][col]
};
inline M3 m3_zero() { return M3{{{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}}; }
inline M3 m3_identity() { return M3{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }
inline M3 operator+(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] + b.m[i][j];
  return r;
}
inline M3 operator-(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] - b.m[i][j];
  return r;
}
inline M3 operator*(const M3& a, double s) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[i][j] * s;
  return r;
}
inline M3 operator*(const M3& a, const M3& b) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
  return r;
}
inline V3 operator*(const M3& a, V3 v) {
  return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
          a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
          a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
inline M3 transpose(const M3& a) {
  M3 r;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) r.m[i][j] = a.m[j][i];
  return r;
}
inline M3 outer(V3 a, V3 b) {  // a * b^T
  return M3{{{a.x * b.x, a.x * b.y, a.x * b.z},
             {a.y * b.x, a.y * b.y, a.y * b.z},
             {a.z * b.x, a.z * b.y, a.z * b.z}}};
}
inline double trace(const M3& a) { return a.m[0][0] + a.m[1][1] + a.m[2][2]; }

struct Quat {
  double w, x, y, z;
};
// nalgebra Quaternion * Quaternion (Hamilton product)
inline Quat qmul(Quat a, Quat b) {
  return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z,
          a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
          a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x,
          a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w};
}
inline Quat qconj(Quat q) { return {q.w, -q.x, -q.y, -q.z}; }
// UnitQuaternion * Vector3: t = 2 (imag x v); v + w t + imag x t
inline V3 qrot(Quat q, V3 v) {
  V3 im{q.x, q.y, q.z};
  V3 t = cross(im, v) * 2.0;
  return v + t * q.w + cross(im, t);
}
// UnitQuaternion::to_rotation_matrix
inline M3 qmat(Quat q) {
  double i = q.x, j = q.y, k = q.z, w = q.w;
  double ww = w * w, ii = i * i, jj = j * j, kk = k * k;
  double ij = i * j * 2.0, wk = w * k * 2.0, wj = w * j * 2.0;
  double ik = i * k * 2.0, jk = j * k * 2.0, wi = w * i * 2.0;
  return M3{{{ww + ii - jj - kk, ij - wk, wj + ik},
             {wk + ij, ww - ii + jj - kk, jk - wi},
             {ik - wj, wi + jk, ww - ii - jj + kk}}};
}
// UnitQuaternion::from_axis_angle
inline Quat q_axis_angle(V3 axis, double angle) {
  double s = std::sin(angle / 2.0), c = std::cos(angle / 2.0);
  return {c, axis.x * s, axis.y * s, axis.z * s};
}
// UnitQuaternion::from_euler_angles(roll, pitch, yaw)
inline Quat q_euler(double roll, double pitch, double yaw) {
  double sr = std::sin(roll * 0.5), cr = std::cos(roll * 0.5);
  double sp = std::sin(pitch * 0.5), cp = std::cos(pitch * 0.5);
  double sy = std::sin(yaw * 0.5), cy = std::cos(yaw * 0.5);
  return {cr * cp * cy + sr * sp * sy, sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy,
          cr * cp * sy - sr * sp * cy};
}
// UnitQuaternion::from_scaled_axis / ::new(axisangle)
inline Quat q_scaled_axis(V3 aa) {
  double n = norm(aa);
  if (n == 0.0) return {1, 0, 0, 0};
  return q_axis_angle(aa / n, n);
}

struct Iso {
  Quat q;
  V3 t;
};
inline Iso iso_identity() { return {{1, 0, 0, 0}, {0, 0, 0}}; }
// Isometry3 * Isometry3
inline Iso iso_mul(const Iso& a, const Iso& b) { return {qmul(a.q, b.q), a.t + qrot(a.q, b.t)}; }
// Isometry3::inverse
inline Iso iso_inv(const Iso& a) {
  Quat qi = qconj(a.q);
  return {qi, qrot(qi, -a.t)};
}

// ---------------------------------------------------------------- mechanism
enum { J_FIXED = 0, J_REV = 1, J_PRIS = 2, J_FLOAT = 3 };

struct ContactPoint {  // contact.rs:17-21
  V3 location;
  double k;
};
struct HalfSpace {  // collision/halfspace.rs:6-11
  V3 point, normal;
  double alpha, mu;
};
struct Body {
  int parent;  // 0 = world
  int jtype;
  V3 axis;
  Iso init_iso;
  M3 moment;  // inertia.rs:32-37
  V3 cross_part;
  double mass;
  bool has_spring;
  double spring_k, spring_l;
  double armature;  // revolute.rs:29; only hybrid/articulated/mod.rs:247 reads it in the reference
  int qoff, voff, nq, nv;
  std::vector<ContactPoint> contact_points;  // rigid_body.rs:65
  int cp_index0;                             // index of this body's first point in the flat output
};

}  // namespace

namespace {
struct SpringContactDef {  // contact.rs:74-81
  int body;  // 1-based
  double l_rest;
  V3 direction;
  double k;
};
}  // namespace

struct gpo_mechanism {
  std::vector<SpringContactDef> spring_contacts;
  int nb, n_q, n_v, n_cp;
  std::vector<Body> bodies;  // bodies[i-1] = body i
  std::vector<HalfSpace> halfspaces;
  std::vector<std::vector<char>> supports;  // supports[j-1][i-1]
};

namespace {

struct SpatialInertia {  // world frame
  M3 moment;
  V3 cross_part;
  double mass;
};
struct SV {  // twist / wrench / spatial acceleration: angular, linear
  V3 ang, lin;
};
struct Jac {  // GeometricJacobian: up to 6 columns
  int k;
  V3 ang[6], lin[6];
};

// per-step scratch, fixed size (no heap in the hot loop)
struct Work {
  Iso joint_iso[GPO_MAX_BODIES];        // joint.transform.iso
  Iso b2r[GPO_MAX_BODIES + 1];          // bodies_to_root, [0] = world
  SV joint_twist[GPO_MAX_BODIES];       // body frame
  SV twist[GPO_MAX_BODIES + 1];         // world frame, [0] = world
  Jac S[GPO_MAX_BODIES + 1];            // motion subspaces in world frame
  SpatialInertia I[GPO_MAX_BODIES + 1];
  SpatialInertia Ic[GPO_MAX_BODIES + 1];
  SV contact_wrench[GPO_MAX_BODIES + 1];
  SV bias_accel[GPO_MAX_BODIES + 1];
  SV wrench[GPO_MAX_BODIES + 1];
  double M[GPO_MAX_NV * GPO_MAX_NV];
  double c[GPO_MAX_NV];
  double* sc_state = nullptr;  // spring contact state of this environment (8 doubles each), or null
  int sc_flags = 0;
};

// joint/revolute.rs:97-102, prismatic.rs:82-87, floating.rs:26-31 + pose.rs:30-33, fixed.rs
inline Iso joint_transform(const Body& b, const double* q) {
  switch (b.jtype) {
    case J_REV: {
      Iso r{q_axis_angle(b.axis, q[b.qoff]), {0, 0, 0}};
      return iso_mul(b.init_iso, r);
    }
    case J_PRIS: {
      Iso t{{1, 0, 0, 0}, b.axis * q[b.qoff]};
      return iso_mul(b.init_iso, t);
    }
    case J_FLOAT: {
      const double* p = q + b.qoff;  // x,y,z,w,tx,ty,tz (joint/mod.rs:216-218)
      Iso pose{{p[3], p[0], p[1], p[2]}, {p[4], p[5], p[6]}};
      return iso_mul(b.init_iso, pose);
    }
    default:
      return b.init_iso;
  }
}

// mechanism.rs:153-170
void bodies_to_root(const gpo_mechanism* m, const double* q, Work& w) {
  w.b2r[0] = iso_identity();
  for (int i = 1; i <= m->nb; ++i) {
    const Body& b = m->bodies[i - 1];
    w.joint_iso[i - 1] = joint_transform(b, q);
    w.b2r[i] = iso_mul(w.b2r[b.parent], w.joint_iso[i - 1]);
  }
}

// twist.rs:36-61
inline SV joint_twist(const Body& b, const double* v) {
  switch (b.jtype) {
    case J_REV:
      return {b.axis * v[b.voff], {0, 0, 0}};
    case J_PRIS:
      return {{0, 0, 0}, b.axis * v[b.voff]};
    case J_FLOAT: {
      const double* p = v + b.voff;
      return {{p[0], p[1], p[2]}, {p[3], p[4], p[5]}};
    }
    default:
      return {{0, 0, 0}, {0, 0, 0}};
  }
}

// Twist::transform, twist.rs:74-93 (quaternion rotate, then t x angular)
inline SV twist_transform(const SV& t, const Iso& iso) {
  V3 ang = qrot(iso.q, t.ang);
  V3 lin = qrot(iso.q, t.lin) + cross(iso.t, ang);
  return {ang, lin};
}
// SpatialAcceleration::transform, spatial_acceleration.rs:33-53 (rotation MATRIX)
inline SV accel_transform(const SV& a, const Iso& iso) {
  M3 rot = qmat(iso.q);
  V3 ang = rot * a.ang;
  V3 lin = rot * a.lin + cross(iso.t, ang);
  return {ang, lin};
}

// twist.rs:176-204
void body_twists(const gpo_mechanism* m, const double* v, Work& w) {
  w.twist[0] = {{0, 0, 0}, {0, 0, 0}};
  for (int i = 1; i <= m->nb; ++i) {
    const Body& b = m->bodies[i - 1];
    w.joint_twist[i - 1] = joint_twist(b, v);
    SV jt = twist_transform(w.joint_twist[i - 1], w.b2r[i]);
    w.twist[i] = {w.twist[b.parent].ang + jt.ang, w.twist[b.parent].lin + jt.lin};
  }
}

// *Joint::motion_subspace + GeometricJacobian::transform (geometric_jacobian.rs:87-107)
inline Jac motion_subspace_world(const Body& b, const Iso& body_to_root) {
  Jac body;
  body.k = b.nv;
  switch (b.jtype) {
    case J_REV:
      body.ang[0] = b.axis;
      body.lin[0] = {0, 0, 0};
      break;
    case J_PRIS:
      body.ang[0] = {0, 0, 0};
      body.lin[0] = b.axis;
      break;
    case J_FLOAT:
      for (int c = 0; c < 6; ++c) {
        body.ang[c] = {0, 0, 0};
        body.lin[c] = {0, 0, 0};
      }
      body.ang[0].x = 1; body.ang[1].y = 1; body.ang[2].z = 1;
      body.lin[3].x = 1; body.lin[4].y = 1; body.lin[5].z = 1;
      break;
    default:
      break;
  }
  M3 rot = qmat(body_to_root.q);
  V3 trans = body_to_root.t;
  Jac out;
  out.k = body.k;
  for (int c = 0; c < body.k; ++c) {
    out.ang[c] = rot * body.ang[c];
    out.lin[c] = rot * body.lin[c] + cross(trans, out.ang[c]);
  }
  return out;
}

// SpatialInertia::transform, inertia.rs:106-134
inline SpatialInertia inertia_transform(const Body& b, const Iso& body_to_root) {
  M3 R = qmat(body_to_root.q);
  V3 p = body_to_root.t;
  const M3& J = b.moment;
  V3 mc = b.cross_part;
  double mass = b.mass;

  V3 Rmc = R * mc;
  V3 mp = mass * p;
  V3 mcnew = Rmc + mp;
  M3 X = outer(Rmc, p);
  M3 Y = X + transpose(X) + outer(mp, p);
  M3 Jnew = (R * J) * transpose(R) - Y + m3_identity() * trace(Y);
  return {Jnew, mcnew, mass};
}

// util.rs:18-28
inline void mul_inertia(const M3& J, V3 c, double mass, V3 w, V3 v, V3& ang, V3& lin) {
  ang = J * w + cross(c, v);
  lin = mass * v - cross(c, w);
}

// mechanism.rs:637-696 (+ :592-625, momentum.rs:17-47)
// with_armature: Articulated::update_mass_matrix only (hybrid/articulated/mod.rs:199-269); MechanismState's
// mass_matrix never reads the armature
void mass_matrix(const gpo_mechanism* m, Work& w, bool with_armature = false) {
  const int nb = m->nb, n_v = m->n_v;
  for (int i = 0; i < n_v * n_v; ++i) w.M[i] = 0.0;
  for (int i = 1; i <= nb; ++i) {
    w.S[i] = motion_subspace_world(m->bodies[i - 1], w.b2r[i]);
    w.I[i] = inertia_transform(m->bodies[i - 1], w.b2r[i]);
  }
  // compute_crb_inertias: reverse body order, children added in ascending joint order
  for (int i = nb; i >= 1; --i) {
    SpatialInertia crb = w.I[i];
    for (int j = 1; j <= nb; ++j) {
      if (m->bodies[j - 1].parent == i) {
        crb.moment = crb.moment + w.Ic[j].moment;
        crb.cross_part = crb.cross_part + w.Ic[j].cross_part;
        crb.mass = crb.mass + w.Ic[j].mass;
      }
    }
    w.Ic[i] = crb;
  }
  for (int i = 1; i <= nb; ++i) {
    const Body& bi = m->bodies[i - 1];
    if (bi.jtype == J_FIXED) continue;
    const SpatialInertia& Ici = w.Ic[i];
    const Jac& Si = w.S[i];
    // Fi = MomentumMatrix::mul(Ici, Si)
    V3 Fang[6], Flin[6];
    for (int c = 0; c < Si.k; ++c) {
      Fang[c] = Ici.moment * Si.ang[c] + cross(Ici.cross_part, Si.lin[c]);
      Flin[c] = Ici.mass * Si.lin[c] - cross(Ici.cross_part, Si.ang[c]);
    }
    for (int j = 1; j <= i; ++j) {
      const Body& bj = m->bodies[j - 1];
      if (bj.jtype == J_FIXED) continue;
      if (!m->supports[j - 1][i - 1]) continue;
      const Jac& Sj = w.S[j];
      // Hij = Fi.angular^T * Sj.angular + Fi.linear^T * Sj.linear
      for (int r = 0; r < Si.k; ++r)
        for (int c = 0; c < Sj.k; ++c)
          w.M[(bi.voff + r) * n_v + (bj.voff + c)] = dot(Fang[r], Sj.ang[c]) + dot(Flin[r], Sj.lin[c]);
    }
  }
  // armature on the joint's own diagonal entry (hybrid/articulated/mod.rs:247; revolute joints only,
  // joint/mod.rs:96-108)
  if (with_armature) {
    for (int i = 1; i <= nb; ++i) {
      const Body& bi = m->bodies[i - 1];
      if (bi.jtype == J_REV && bi.armature != 0.0) w.M[bi.voff * n_v + bi.voff] += bi.armature;
    }
  }
  // mirror the lower triangle (mechanism.rs:688-693)
  for (int i = 0; i < n_v; ++i)
    for (int j = i + 1; j < n_v; ++j) w.M[i * n_v + j] = w.M[j * n_v + i];
}

// contact.rs:260-302
inline V3 calculate_contact_force(double penetration, V3 normal, V3 velocity, double k_A, double k_B,
                                  double alpha, double mu) {
  double z = penetration;
  double z_dot = -dot(velocity, normal);
  double zn = std::pow(z, 3.0 / 2.0);
  double k = k_A * k_B / (k_A + k_B);
  double a = alpha;
  double lambda = 3.0 / 2.0 * a * k;
  double pi_raw = lambda * zn * z_dot + k * zn;
  // f64::max(NaN, 0.0) == 0.0 in Rust: a NaN from powf of a slightly negative z gives no force
  double pi = (pi_raw > 0.0) ? pi_raw : 0.0;
  V3 f_normal = pi * normal;

  V3 v_t = velocity + z_dot * normal;
  double v_t_norm = norm(v_t);
  V3 f_friction{0, 0, 0};
  if (!(v_t_norm == 0.0)) {
    double v_s = 1e-3;
    double s = v_t_norm / v_s;
    double mu_eff = (s > 1.0) ? mu : mu * s;
    f_friction = (-mu_eff * pi) * (v_t / v_t_norm);
  }
  return f_normal + f_friction;
}

// contact.rs:97-128 (point-vs-halfspace branch) with contact.rs:41-68, twist.rs:119-132,
// halfspace.rs:39-44, contact.rs:321-338, wrench.rs:31-37
void contact_dynamics(const gpo_mechanism* m, Work& w, double* contact_forces) {
  for (int i = 1; i <= m->nb; ++i) {
    const Body& b = m->bodies[i - 1];
    SV wrench{{0, 0, 0}, {0, 0, 0}};
    const Iso& body_to_root = w.b2r[i];
    const SV& twist = w.twist[i];
    M3 rot = qmat(body_to_root.q);
    V3 trans = body_to_root.t;
    for (size_t c = 0; c < b.contact_points.size(); ++c) {
      const ContactPoint& cp = b.contact_points[c];
      V3 location = rot * cp.location + trans;              // ContactPoint::transform
      V3 velocity = twist.lin + cross(twist.ang, location);  // point_velocity
      V3 total{0, 0, 0};
      for (const HalfSpace& hs : m->halfspaces) {
        double margin = 1e-8;
        if (!(dot(location - hs.point, hs.normal) <= margin)) continue;  // has_inside
        double penetration = -dot(location - hs.point, hs.normal);       // compute_contact
        V3 f = calculate_contact_force(penetration, hs.normal, velocity, cp.k, 50e3, hs.alpha, hs.mu);
        wrench.ang = wrench.ang + cross(location, f);  // Wrench::from_force
        wrench.lin = wrench.lin + f;
        total = total + f;
      }
      if (contact_forces) {
        double* o = contact_forces + 3 * (b.cp_index0 + (int)c);
        o[0] = total.x; o[1] = total.y; o[2] = total.z;
      }
    }
    // spring contacts with halfspaces (contact.rs:133-186); state lives in w.sc_state
    if (w.sc_state) {
      for (size_t s = 0; s < m->spring_contacts.size(); ++s) {
        const SpringContactDef& sc = m->spring_contacts[s];
        if (sc.body != i) continue;
        double* st = w.sc_state + 8 * s;
        V3 body_location = trans;
        if (st[0] == 0.0) {
          V3 direction{st[4], st[5], st[6]};
          V3 spring_direction = rot * direction;
          V3 contact_location = body_location + spring_direction * st[7];
          for (size_t h = 0; h < m->halfspaces.size(); ++h) {
            const HalfSpace& hs = m->halfspaces[h];
            if (!(dot(contact_location - hs.point, hs.normal) <= 1e-8)) continue;
            st[0] = (double)(h + 1);
            st[1] = contact_location.x; st[2] = contact_location.y; st[3] = contact_location.z;
            break;
          }
        } else {
          V3 contact_location{st[1], st[2], st[3]};
          V3 dvec = contact_location - body_location;
          V3 spring_direction = dvec / norm(dvec);
          const HalfSpace& hs = m->halfspaces[(size_t)(long long)st[0] - 1];
          if (dot(spring_direction, hs.normal) > 0.0) w.sc_flags |= 4;  // panic!("Spring force is into the halfspace!")
          double direction_distance = norm(dvec);
          if (direction_distance < st[7]) {
            double spring_force = -sc.k * (direction_distance - st[7]);
            V3 force = -spring_direction * spring_force;
            wrench.ang = wrench.ang + cross(body_location, force);
            wrench.lin = wrench.lin + force;
          } else {
            st[0] = 0.0;
            V3 nd = spring_direction / norm(spring_direction);
            st[4] = nd.x; st[5] = nd.y; st[6] = nd.z;
          }
        }
      }
    }
    w.contact_wrench[i] = wrench;
  }
}

// dynamics.rs:143-224: coriolis bias (world -> body -> commutator -> world) + inv gravity
void bias_accelerations(const gpo_mechanism* m, Work& w) {
  SV coriolis[GPO_MAX_BODIES + 1];
  coriolis[0] = {{0, 0, 0}, {0, 0, 0}};
  for (int i = 1; i <= m->nb; ++i) {
    const Body& b = m->bodies[i - 1];
    const Iso& body_to_root = w.b2r[i];
    Iso root_to_body = iso_inv(body_to_root);
    SV body_twist = twist_transform(w.twist[i], root_to_body);
    const SV& jt = w.joint_twist[i - 1];
    // se3_commutator(body_twist, joint_twist), util.rs:44-53
    SV cb{cross(body_twist.ang, jt.ang), cross(body_twist.ang, jt.lin) + cross(body_twist.lin, jt.ang)};
    SV cw = accel_transform(cb, body_to_root);
    coriolis[i] = {coriolis[b.parent].ang + cw.ang, coriolis[b.parent].lin + cw.lin};
  }
  for (int i = 1; i <= m->nb; ++i) {
    V3 g{0.0, 0.0, GRAVITY};
    w.bias_accel[i] = {V3{0, 0, 0} + coriolis[i].ang, g + coriolis[i].lin};
  }
}

// dynamics.rs:41-100 (recomputes the world inertias, :48)
void newton_euler(const gpo_mechanism* m, Work& w) {
  w.wrench[0] = {{0, 0, 0}, {0, 0, 0}};
  for (int i = 1; i <= m->nb; ++i) {
    SpatialInertia I = inertia_transform(m->bodies[i - 1], w.b2r[i]);
    const SV& twist = w.twist[i];
    const SV& accel = w.bias_accel[i];
    V3 ang, lin, am, lm;
    mul_inertia(I.moment, I.cross_part, I.mass, accel.ang, accel.lin, ang, lin);
    mul_inertia(I.moment, I.cross_part, I.mass, twist.ang, twist.lin, am, lm);
    ang = ang + (cross(twist.ang, am) + cross(twist.lin, lm));
    lin = lin + cross(twist.ang, lm);
    w.wrench[i] = {ang, lin};
  }
}

// wrench.rs:96-126
void compute_torques(const gpo_mechanism* m, Work& w) {
  SV jw[GPO_MAX_BODIES + 1];
  for (int i = 0; i <= m->nb; ++i) jw[i] = w.wrench[i];
  for (int i = m->nb; i >= 1; --i) {
    const Body& b = m->bodies[i - 1];
    SV joint_wrench = jw[i];
    jw[b.parent].ang = jw[b.parent].ang + joint_wrench.ang;
    jw[b.parent].lin = jw[b.parent].lin + joint_wrench.lin;
    Jac S = motion_subspace_world(b, w.b2r[i]);
    for (int c = 0; c < S.k; ++c)
      w.c[b.voff + c] = dot(S.ang[c], joint_wrench.ang) + dot(S.lin[c], joint_wrench.lin);
  }
}

// nalgebra LU with partial (row) pivoting + solve; dynamics.rs:255-276
bool lu_solve(int n, const double* A_in, const double* b_in, double* x) {
  double A[GPO_MAX_NV * GPO_MAX_NV];
  int perm[GPO_MAX_NV];
  std::memcpy(A, A_in, sizeof(double) * n * n);
  for (int i = 0; i < n; ++i) x[i] = b_in[i];
  for (int i = 0; i < n; ++i) {
    int piv = i;
    double best = std::fabs(A[i * n + i]);
    for (int r = i + 1; r < n; ++r) {
      double a = std::fabs(A[r * n + i]);
      if (a > best) { best = a; piv = r; }
    }
    double diag = A[piv * n + i];
    if (diag == 0.0) return false;
    perm[i] = piv;
    if (piv != i) {
      for (int c = 0; c < n; ++c) std::swap(A[i * n + c], A[piv * n + c]);
    }
    double inv_diag = 1.0 / diag;
    for (int r = i + 1; r < n; ++r) A[r * n + i] *= inv_diag;
    for (int r = i + 1; r < n; ++r) {
      double l = A[r * n + i];
      for (int c = i + 1; c < n; ++c) A[r * n + c] -= l * A[i * n + c];
    }
  }
  for (int i = 0; i < n; ++i)
    if (perm[i] != i) std::swap(x[i], x[perm[i]]);
  for (int i = 0; i < n; ++i)  // unit lower
    for (int r = i + 1; r < n; ++r) x[r] -= A[r * n + i] * x[i];
  for (int i = n - 1; i >= 0; --i) {  // upper
    x[i] /= A[i * n + i];
    for (int r = 0; r < i; ++r) x[r] -= A[r * n + i] * x[i];
  }
  return true;
}

// dynamics.rs:322-364
int dynamics_continuous(const gpo_mechanism* m, const double* q, const double* v, const double* tau,
                        double* vdot, double* contact_forces, Work& w) {
  bodies_to_root(m, q, w);       // dynamics_quantities, dynamics.rs:280-295
  body_twists(m, v, w);
  mass_matrix(m, w);
  contact_dynamics(m, w, contact_forces);
  bias_accelerations(m, w);      // dynamics_bias, dynamics.rs:233-251
  newton_euler(m, w);
  for (int i = 1; i <= m->nb; ++i) {
    w.wrench[i].ang = w.wrench[i].ang - w.contact_wrench[i].ang;
    w.wrench[i].lin = w.wrench[i].lin - w.contact_wrench[i].lin;
  }
  compute_torques(m, w);
  // add_prismatic_joint_spring_force, dynamics.rs:298-315; spring_force = -k (l - l_rest)
  double rhs[GPO_MAX_NV];
  for (int i = 0; i < m->n_v; ++i) rhs[i] = tau ? tau[i] : 0.0;
  for (int i = 1; i <= m->nb; ++i) {
    const Body& b = m->bodies[i - 1];
    if (b.jtype == J_PRIS && b.has_spring) {
      double l = q[b.qoff];
      double f_spring = -b.spring_k * (l - b.spring_l);
      rhs[b.voff] = rhs[b.voff] + f_spring;
    }
  }
  if (m->n_v == 0) return 0;
  for (int i = 0; i < m->n_v; ++i) rhs[i] = rhs[i] - w.c[i];
  return lu_solve(m->n_v, w.M, rhs, vdot) ? 0 : 1;
}

// util.rs:83-102
inline Quat quaternion_derivative(Quat q, V3 omega) {
  double w = q.w, x = q.x, y = q.y, z = q.z;
  // rows of (Matrix4x3 / 2.0) * omega
  double r0 = (-x / 2.0) * omega.x + (-y / 2.0) * omega.y + (-z / 2.0) * omega.z;
  double r1 = (w / 2.0) * omega.x + (-z / 2.0) * omega.y + (y / 2.0) * omega.z;
  double r2 = (z / 2.0) * omega.x + (w / 2.0) * omega.y + (-x / 2.0) * omega.z;
  double r3 = (-y / 2.0) * omega.x + (x / 2.0) * omega.y + (w / 2.0) * omega.z;
  return {r0, r1, r2, r3};
}

// pose update shared by compute_new_q (integrators.rs:296-319) and euler_step (:230-271)
inline void integrate_pose(const double* qi, V3 ang, V3 lin, double dt, double* qo) {
  Quat rot{qi[3], qi[0], qi[1], qi[2]};
  V3 trans{qi[4], qi[5], qi[6]};
  Quat qd = quaternion_derivative(rot, ang);
  V3 translation_dot = qrot(rot, lin);
  V3 tn = trans + translation_dot * dt;
  Quat qn{rot.w + qd.w * dt, rot.x + qd.x * dt, rot.y + qd.y * dt, rot.z + qd.z * dt};
  double n = std::sqrt(qn.w * qn.w + qn.x * qn.x + qn.y * qn.y + qn.z * qn.z);
  qo[0] = qn.x / n; qo[1] = qn.y / n; qo[2] = qn.z / n; qo[3] = qn.w / n;
  qo[4] = tn.x; qo[5] = tn.y; qo[6] = tn.z;
}

// integrators.rs:276-319: v' = v + vdot dt ; q' = q (+) v' dt
void semi_implicit_euler_step(const gpo_mechanism* m, const double* q, const double* v,
                              const double* vdot, double dt, double* qn, double* vn) {
  for (int i = 0; i < m->n_v; ++i) vn[i] = v[i] + vdot[i] * dt;
  for (const Body& b : m->bodies) {
    if (b.jtype == J_REV || b.jtype == J_PRIS) {
      qn[b.qoff] = q[b.qoff] + vn[b.voff] * dt;
    } else if (b.jtype == J_FLOAT) {
      const double* s = vn + b.voff;
      integrate_pose(q + b.qoff, {s[0], s[1], s[2]}, {s[3], s[4], s[5]}, dt, qn + b.qoff);
    }
  }
}

// integrators.rs:230-271: v' = v + vdot dt ; q' = q (+) v dt
void euler_step(const gpo_mechanism* m, const double* q, const double* v, const double* vdot,
                double dt, double* qn, double* vn) {
  for (int i = 0; i < m->n_v; ++i) vn[i] = v[i] + vdot[i] * dt;
  for (const Body& b : m->bodies) {
    if (b.jtype == J_REV || b.jtype == J_PRIS) {
      qn[b.qoff] = q[b.qoff] + v[b.voff] * dt;
    } else if (b.jtype == J_FLOAT) {
      const double* s = v + b.voff;
      integrate_pose(q + b.qoff, {s[0], s[1], s[2]}, {s[3], s[4], s[5]}, dt, qn + b.qoff);
    }
  }
}

int step_impl(const gpo_mechanism* m, double* q, double* v, const double* tau, double dt,
              int integrator, Work& w) {
  const int nq = m->n_q, nv = m->n_v;
  double qn[GPO_MAX_NV + GPO_MAX_BODIES], vn[GPO_MAX_NV];
  int rc = 0;
  if (integrator == 0) {  // integrators.rs:25-39
    double vdot[GPO_MAX_NV];
    rc = dynamics_continuous(m, q, v, tau, vdot, nullptr, w);
    semi_implicit_euler_step(m, q, v, vdot, dt, qn, vn);
  } else if (integrator == 1) {  // runge_kutta_2, integrators.rs:177-192
    double f1[GPO_MAX_NV], f2[GPO_MAX_NV], q1[GPO_MAX_NV + GPO_MAX_BODIES], v1[GPO_MAX_NV];
    rc |= dynamics_continuous(m, q, v, tau, f1, nullptr, w);
    euler_step(m, q, v, f1, dt / 2.0, q1, v1);
    rc |= dynamics_continuous(m, q1, v1, tau, f2, nullptr, w);
    euler_step(m, q, v, f2, dt, qn, vn);
  } else if (integrator == 2) {  // runge_kutta_4, integrators.rs:195-225
    double f1[GPO_MAX_NV], f2[GPO_MAX_NV], f3[GPO_MAX_NV], f4[GPO_MAX_NV], ff[GPO_MAX_NV];
    double qs[GPO_MAX_NV + GPO_MAX_BODIES], vs[GPO_MAX_NV];
    rc |= dynamics_continuous(m, q, v, tau, f1, nullptr, w);
    euler_step(m, q, v, f1, dt / 2.0, qs, vs);
    rc |= dynamics_continuous(m, qs, vs, tau, f2, nullptr, w);
    euler_step(m, q, v, f2, dt / 2.0, qs, vs);
    rc |= dynamics_continuous(m, qs, vs, tau, f3, nullptr, w);
    euler_step(m, q, v, f3, dt, qs, vs);
    rc |= dynamics_continuous(m, qs, vs, tau, f4, nullptr, w);
    for (int i = 0; i < nv; ++i) ff[i] = (f1[i] + f2[i] * 2.0 + f3[i] * 2.0 + f4[i]) / 6.0;
    euler_step(m, q, v, ff, dt, qn, vn);
  } else {
    return 2;
  }
  std::memcpy(q, qn, sizeof(double) * nq);
  std::memcpy(v, vn, sizeof(double) * nv);
  return rc;
}

// inertia.rs:182-202
inline double kinetic_energy_body(const SpatialInertia& I, const SV& t) {
  V3 w = t.ang, v = t.lin;
  return (dot(w, I.moment * w) + dot(v, I.mass * v + 2.0 * cross(w, I.cross_part))) / 2.0;
}

// energy.rs:19-26
inline double double_pendulum_potential_energy2(const double* q, double m, double l) {
  double q1 = q[0], q2 = q[1];
  double h1 = l * std::sin(q1);
  double h2 = l * std::sin(q1) + l * std::sin(q1 + q2);
  return m * GRAVITY * (h1 + h2);
}

double kinetic_energy_impl(const gpo_mechanism* m, const double* q, const double* v, Work& w) {
  bodies_to_root(m, q, w);
  body_twists(m, v, w);
  double KE = 0.0;
  for (int i = 1; i <= m->nb; ++i) {
    SpatialInertia I = inertia_transform(m->bodies[i - 1], w.b2r[i]);
    KE += kinetic_energy_body(I, w.twist[i]);
  }
  return KE;
}

// rem_euclid for f64
inline double rem_euclid(double a, double b) {
  double r = std::fmod(a, b);
  return (r < 0.0) ? r + std::fabs(b) : r;
}

int control_impl(const gpo_mechanism* m, const double* q, const double* v, int controller,
                 const double* p, double* tau, Work& w) {
  for (int i = 0; i < m->n_v; ++i) tau[i] = 0.0;
  switch (controller) {
    case 0:
      return 0;
    case 1: {  // SO101PositionController, control/so101_control.rs:12-34
      double kp = p[0], kd = p[1], clamp = p[2];
      for (const Body& b : m->bodies) {
        if (b.jtype == J_FIXED) continue;
        if (b.jtype == J_FLOAT) return 1;  // q.float() would panic
        double t = kp * (0.0 - q[b.qoff]) - kd * v[b.voff];
        double sg = std::isnan(t) ? t : std::copysign(1.0, t);  // f64::signum
        tau[b.voff] = sg * std::fmin(std::fabs(t), clamp);
      }
      return 0;
    }
    case 2: {  // swingup_acrobot, control/swingup.rs:9-69
      double mm = p[0], l = p[1];
      double q1 = q[0], q2 = q[1], q1dot = v[0], q2dot = v[1];
      double KE = kinetic_energy_impl(m, q, v, w);
      double PE = double_pendulum_potential_energy2(q, mm, l);
      double E_target = mm * GRAVITY * (l + 2.0 * l);
      double dE = KE + PE - E_target;
      double k3 = 2.0;
      double u_bar = k3 * (dE * q1dot);
      double cap = 10.;
      if (u_bar > cap) u_bar = cap;
      else if (u_bar < -cap) u_bar = -cap;
      q2 = rem_euclid(q2, TWO_PI);
      if (q2 > PI) q2 -= TWO_PI;
      double k1 = 2.0, k2 = 2.0;
      double u_pd = -k1 * q2 - k2 * q2dot;
      double m1 = mm, m2 = mm, lc1 = l, lc2 = l, l1 = l, I1 = 0.0, I2 = 0.0;
      double c1 = std::cos(q1), s2 = std::sin(q2), c2 = std::cos(q2), c12 = std::cos(q1 + q2);
      double m11 = m1 * lc1 * lc1 + m2 * (l1 * l1 + lc2 * lc2 + 2. * l1 * lc2 * c2) + I1 + I2;
      double m22 = m2 * lc2 * lc2 + I2;
      double m12 = m2 * (lc2 * lc2 + l1 * lc2 * c2) + I2;
      double m21 = m12;
      double h1 = -m2 * l1 * lc2 * s2 * q2dot * q2dot - 2. * m2 * l1 * lc2 * s2 * q2dot * q1dot;
      double h2 = m2 * l1 * lc2 * s2 * q1dot * q1dot;
      double phi1 = (m1 * lc1 + m2 * l1) * GRAVITY * c1 + m2 * lc2 * GRAVITY * c12;
      double phi2 = m2 * lc2 * GRAVITY * c12;
      double m22_bar = m22 - m21 * m12 / m11;
      double h2_bar = h2 - m21 * h1 / m11;
      double phi2_bar = phi2 - m21 * phi1 / m11;
      tau[0] = 0.;
      tau[1] = m22_bar * (u_bar + u_pd) + h2_bar + phi2_bar;
      return 0;
    }
    case 3: {  // swingup_cart_pole, control/swingup.rs:76-110
      double m_c = p[0], m_p = p[1], l = p[2];
      double theta = q[1], theta_dot = v[1];
      double cos_theta = std::cos(theta), sin_theta = std::sin(theta);
      double KE = 0.5 * m_p * l * l * theta_dot * theta_dot;
      double PE = -m_p * GRAVITY * l * cos_theta;
      double E_target = m_p * GRAVITY * l;
      double dE = KE + PE - E_target;
      double K = 2.0;
      double u_bar = K * theta_dot * cos_theta * dE / (m_p * l);
      double x = q[0], x_dot = v[0];
      double Kp = 1.0, Kd = 1.0;
      double u_pd = -Kp * x - Kd * x_dot;
      double u = u_bar + u_pd;
      double f = (m_c + m_p * sin_theta * sin_theta) * u - m_p * GRAVITY * sin_theta * cos_theta -
                 m_p * l * sin_theta * theta_dot * theta_dot;
      tau[0] = f;
      tau[1] = 0.0;
      return 0;
    }
    case 5: case 6: case 7: {  // pendulum_gravity_inversion / _energy_shaping / _swing_up_and_balance, control/mod.rs:57-105
      if (m->nb != 1 || m->bodies[0].jtype != J_REV) return 1;
      const Body& b = m->bodies[0];
      double mass = b.mass;
      V3 com{b.cross_part.x / mass, b.cross_part.y / mass, b.cross_part.z / mass};  // SpatialInertia::center_of_mass
      double length_to_com = std::sqrt(com.x * com.x + com.y * com.y + com.z * com.z);
      double qq = q[0], vv = v[0];
      bool shaping = controller == 6 || (controller == 7 && std::fabs(qq - PI) > 0.15);
      if (shaping) {
        V3 omega{b.axis.x * vv, b.axis.y * vv, b.axis.z * vv};
        double E_desired = mass * GRAVITY * length_to_com;
        V3 Jw{b.moment.m[0][0] * omega.x + b.moment.m[0][1] * omega.y + b.moment.m[0][2] * omega.z,
              b.moment.m[1][0] * omega.x + b.moment.m[1][1] * omega.y + b.moment.m[1][2] * omega.z,
              b.moment.m[2][0] * omega.x + b.moment.m[2][1] * omega.y + b.moment.m[2][2] * omega.z};
        double KE = 0.5 * (omega.x * Jw.x + omega.y * Jw.y + omega.z * Jw.z);
        double PE = mass * GRAVITY * length_to_com * (-std::cos(qq));
        double E_diff = KE + PE - E_desired;
        tau[0] = -0.1 * vv * E_diff;
      } else {
        double gravity_inversion = 2.0 * mass * GRAVITY * length_to_com * std::sin(qq);
        double damping = -10.0 * vv;
        tau[0] = gravity_inversion + damping;
      }
      return 0;
    }
    default:
      return 2;
  }
}

// Hopper1DController::control, control/energy_control.rs:35-101. cstate = {leg_length_setpoint,
// v_vertical_prev}; p = {k_spring, h_setpoint, body_leg_length, leg_foot_length}
int control_hopper1d(const gpo_mechanism* m, const double* q, const double* v, const double* p, double* cstate,
                     double* tau, Work& w) {
  if (m->nb != 3 || m->bodies[0].jtype != J_FLOAT || m->bodies[1].jtype != J_PRIS || m->bodies[2].jtype != J_PRIS)
    return 1;
  const double k_spring = p[0], h_setpoint = p[1], body_leg_length = p[2], leg_foot_length = p[3];
  for (int i = 0; i < 6; ++i) tau[i] = 0.0;  // first floating joint unactuated
  double q1 = q[7], v1 = v[6], q_foot = q[8], v_foot = v[7];
  double v_vertical = v[5];
  double m_body = m->bodies[0].mass, m_leg = m->bodies[1].mass, m_foot = m->bodies[2].mass;
  if (cstate[1] < 0.0 && v_vertical > 0.0) {
    // hopper_energy, energy.rs:44-53
    double KE = kinetic_energy_impl(m, q, v, w);
    double PE = 0.0;
    for (int i = 1; i <= m->nb; ++i) PE += m->bodies[i - 1].mass * GRAVITY * w.b2r[i].t.z;
    double l_rest = 0.0;
    double EPE = 0.5 * k_spring * (q_foot - l_rest) * (q_foot - l_rest);
    double E = KE + PE + EPE;
    double E_target = GRAVITY * (m_body * h_setpoint + m_leg * (h_setpoint - body_leg_length) +
                                 m_foot * (h_setpoint - body_leg_length - leg_foot_length));
    double dE = E_target - E;
    cstate[0] = q_foot + std::sqrt(q_foot * q_foot + 2.0 * dE / k_spring);
  } else if (cstate[1] > 0.0 && v_vertical < 0.0) {
    cstate[0] = 0.0;
  }
  double kp = 2000.0, kd = 100.0;
  double p_term = kp * (cstate[0] - q1);
  double d_term = kd * (0.0 - v1);
  double tau1 = p_term + d_term;
  double l_rest = 0.0;
  double tau_foot;
  if (q_foot < l_rest) {
    tau_foot = -k_spring * (q_foot - l_rest);  // spring_force
  } else {
    double k_stop = 1e5, b_stop = 125.0;
    tau_foot = -k_stop * (q_foot - l_rest) - b_stop * v_foot;  // mechanical_stop
  }
  tau[6] = tau1 - tau_foot;
  tau[7] = tau_foot;
  cstate[1] = v_vertical;
  return 0;
}

int rollout_impl(const gpo_mechanism* m, double* q, double* v, const double* tau, double dt,
                 int64_t n_steps, int integrator, int controller, const double* params,
                 double* hq, double* hv, Work& w) {
  int rc = 0;
  double tau_c[GPO_MAX_NV];
  double cstate[2] = {0.0, 0.0};  // Hopper1DController starts from (0, 0), examples/1D_hopper.rs:105-113
  if (hq) std::memcpy(hq, q, sizeof(double) * m->n_q);
  if (hv) std::memcpy(hv, v, sizeof(double) * m->n_v);
  for (int64_t s = 0; s < n_steps; ++s) {
    const double* t = tau;
    if (controller == 4) {
      rc |= control_hopper1d(m, q, v, params, cstate, tau_c, w);
      t = tau_c;
    } else if (controller != 0) {
      rc |= control_impl(m, q, v, controller, params, tau_c, w);
      t = tau_c;
    }
    rc |= step_impl(m, q, v, t, dt, integrator, w);
    if (hq) std::memcpy(hq + (s + 1) * m->n_q, q, sizeof(double) * m->n_q);
    if (hv) std::memcpy(hv + (s + 1) * m->n_v, v, sizeof(double) * m->n_v);
  }
  return rc;
}

}  // namespace

// ---------------------------------------------------------------- C API
extern "C" {

int gpo_mechanism_create(const gpo_mechanism_desc* d, gpo_mechanism** out) {
  if (!d || !out || d->n_bodies < 0 || d->n_bodies > GPO_MAX_BODIES) return 1;
  gpo_mechanism* m = new gpo_mechanism();
  m->nb = d->n_bodies;
  int qoff = 0, voff = 0;
  for (int i = 0; i < m->nb; ++i) {
    Body b;
    b.parent = d->parent[i];
    if (b.parent < 0 || b.parent > i) { delete m; return 1; }  // parent must precede (mechanism.rs:98-125)
    b.jtype = d->joint_type[i];
    b.axis = {d->axis[3 * i], d->axis[3 * i + 1], d->axis[3 * i + 2]};
    const auto* s = d->init_iso + 7 * i;
    b.init_iso = {{s[3], s[0], s[1], s[2]}, {s[4], s[5], s[6]}};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) b.moment.m[r][c] = d->moment[9 * i + 3 * r + c];
    b.cross_part = {d->cross_part[3 * i], d->cross_part[3 * i + 1], d->cross_part[3 * i + 2]};
    b.mass = d->mass[i];
    b.has_spring = d->has_spring ? d->has_spring[i] != 0 : false;
    b.spring_k = (b.has_spring && d->spring_k) ? d->spring_k[i] : 0.0;
    b.spring_l = (b.has_spring && d->spring_l) ? d->spring_l[i] : 0.0;
    b.armature = d->armature ? d->armature[i] : 0.0;
    switch (b.jtype) {
      case J_REV: case J_PRIS: b.nq = 1; b.nv = 1; break;
      case J_FLOAT: b.nq = 7; b.nv = 6; break;
      case J_FIXED: b.nq = 0; b.nv = 0; break;
      default: delete m; return 1;
    }
    b.qoff = qoff; b.voff = voff;
    qoff += b.nq; voff += b.nv;
    b.cp_index0 = 0;
    m->bodies.push_back(b);
  }
  m->n_q = qoff; m->n_v = voff;
  if (m->n_v > GPO_MAX_NV) { delete m; return 1; }
  // add_contact_point (mechanism.rs:384): appended to the owning body's list, insertion order
  for (int c = 0; c < d->n_contact_points; ++c) {
    int body = d->cp_body[c];
    if (body < 1 || body > m->nb) { delete m; return 1; }
    m->bodies[body - 1].contact_points.push_back(
        {{d->cp_location[3 * c], d->cp_location[3 * c + 1], d->cp_location[3 * c + 2]}, d->cp_k[c]});
  }
  int idx = 0;
  for (Body& b : m->bodies) { b.cp_index0 = idx; idx += (int)b.contact_points.size(); }
  m->n_cp = idx;
  for (int s = 0; s < d->n_spring_contacts; ++s) {
    if (d->sc_body[s] < 1 || d->sc_body[s] > m->nb) { delete m; return 1; }
    m->spring_contacts.push_back({d->sc_body[s], d->sc_l_rest[s],
                                  {d->sc_direction[3 * s], d->sc_direction[3 * s + 1], d->sc_direction[3 * s + 2]},
                                  d->sc_k[s]});
  }
  for (int h = 0; h < d->n_halfspaces; ++h)
    m->halfspaces.push_back({{d->hs_point[3 * h], d->hs_point[3 * h + 1], d->hs_point[3 * h + 2]},
                             {d->hs_normal[3 * h], d->hs_normal[3 * h + 1], d->hs_normal[3 * h + 2]},
                             d->hs_alpha[h], d->hs_mu[h]});
  // supports (mechanism.rs:118-125)
  m->supports.assign(m->nb, std::vector<char>(m->nb, 0));
  for (int i = 1; i <= m->nb; ++i) {
    m->supports[i - 1][i - 1] = 1;
    int cur = i;
    while (m->bodies[cur - 1].parent != 0) {
      int p = m->bodies[cur - 1].parent;
      m->supports[p - 1][i - 1] = 1;
      cur = p;
    }
  }
  *out = m;
  return 0;
}

void gpo_mechanism_destroy(gpo_mechanism* m) { delete m; }
int gpo_n_q(const gpo_mechanism* m) { return m->n_q; }
int gpo_n_v(const gpo_mechanism* m) { return m->n_v; }
void gpo_supports(const gpo_mechanism* m, int32_t* out) {
  for (int j = 0; j < m->nb; ++j)
    for (int i = 0; i < m->nb; ++i) out[j * m->nb + i] = m->supports[j][i];
}

int gpo_dynamics(const gpo_mechanism* m, const double* q, const double* v, const double* tau,
                 double* vdot, double* contact_forces, double* mass_matrix_out, double* bias) {
  Work w;
  int rc = dynamics_continuous(m, q, v, tau, vdot, contact_forces, w);
  if (mass_matrix_out) std::memcpy(mass_matrix_out, w.M, sizeof(double) * m->n_v * m->n_v);
  if (bias) std::memcpy(bias, w.c, sizeof(double) * m->n_v);
  return rc;
}

int gpo_n_spring_contacts(const gpo_mechanism* m) { return (int)m->spring_contacts.size(); }

void gpo_spring_state_init(const gpo_mechanism* m, double* st) {
  for (size_t s = 0; s < m->spring_contacts.size(); ++s) {
    const SpringContactDef& sc = m->spring_contacts[s];
    double* o = st + 8 * s;
    o[0] = 0.0; o[1] = o[2] = o[3] = 0.0;
    o[4] = sc.direction.x; o[5] = sc.direction.y; o[6] = sc.direction.z;
    o[7] = sc.l_rest;
  }
}

int gpo_step_sc(const gpo_mechanism* m, double* q, double* v, const double* tau, double dt, double* sc_state) {
  Work w;
  w.sc_state = sc_state;
  int rc = step_impl(m, q, v, tau, dt, 0, w);
  return rc | w.sc_flags;
}

int gpo_dynamics_sc(const gpo_mechanism* m, const double* q, const double* v, const double* tau, double* vdot,
                    double* sc_state) {
  Work w;
  w.sc_state = sc_state;
  int rc = dynamics_continuous(m, q, v, tau, vdot, nullptr, w);
  return rc | w.sc_flags;
}

int gpo_step(const gpo_mechanism* m, double* q, double* v, const double* tau, double dt, int integrator) {
  Work w;
  return step_impl(m, q, v, tau, dt, integrator, w);
}

int gpo_control(const gpo_mechanism* m, const double* q, const double* v, int controller,
                const double* params, double* tau_out) {
  Work w;
  return control_impl(m, q, v, controller, params, tau_out, w);
}

int gpo_rollout(const gpo_mechanism* m, double* q, double* v, const double* tau, double dt,
                int64_t n_steps, int integrator, int controller, const double* params,
                double* history_q, double* history_v) {
  Work w;
  return rollout_impl(m, q, v, tau, dt, n_steps, integrator, controller, params, history_q, history_v, w);
}

int64_t gpo_simulate_step_count(double final_time, double dt) {
  double t = 0.0;
  int64_t n = 0;
  while (t < final_time) { t += dt; ++n; }
  return n;
}

int gpo_batch_rollout(const gpo_mechanism* m, double* q, double* v, const double* tau,
                      int64_t n_envs, double dt, int64_t n_steps, int integrator, int controller,
                      const double* params, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  std::vector<int> rcs(n_threads, 0);
  auto work = [&](int t) {
    Work w;
    int64_t lo = n_envs * t / n_threads, hi = n_envs * (t + 1) / n_threads;
    for (int64_t e = lo; e < hi; ++e)
      rcs[t] |= rollout_impl(m, q + e * m->n_q, v + e * m->n_v, tau ? tau + e * m->n_v : nullptr, dt,
                             n_steps, integrator, controller, params, nullptr, nullptr, w);
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  int rc = 0;
  for (int r : rcs) rc |= r;
  return rc;
}

int gpo_batch_dynamics(const gpo_mechanism* m, const double* q, const double* v, const double* tau,
                       int64_t n_envs, double* vdot, double* contact_forces, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  std::vector<int> rcs(n_threads, 0);
  auto work = [&](int t) {
    Work w;
    int64_t lo = n_envs * t / n_threads, hi = n_envs * (t + 1) / n_threads;
    for (int64_t e = lo; e < hi; ++e)
      rcs[t] |= dynamics_continuous(m, q + e * m->n_q, v + e * m->n_v, tau ? tau + e * m->n_v : nullptr,
                                    vdot + e * m->n_v,
                                    contact_forces ? contact_forces + e * 3 * m->n_cp : nullptr, w);
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work, t);
  work(0);
  for (auto& x : th) x.join();
  int rc = 0;
  for (int r : rcs) rc |= r;
  return rc;
}

// hybrid/articulated/mod.rs:124-197 (+ update_mass_matrix :199-269, nalgebra Cholesky)
int gpo_free_velocity(const gpo_mechanism* m, const double* q, const double* v, const double* tau, double dt,
                      int gravity_enabled, double* v_free) {
  Work w;
  bodies_to_root(m, q, w);
  body_twists(m, v, w);
  mass_matrix(m, w, true);
  const int nb = m->nb, n = m->n_v;
  // bias accelerations: world-frame commutator of body twist and joint twist, summed down the tree
  SV bias[GPO_MAX_BODIES + 1];
  for (int i = 1; i <= nb; ++i) {
    const Body& b = m->bodies[i - 1];
    SV jt = twist_transform(w.joint_twist[i - 1], w.b2r[i]);
    const SV& bt = w.twist[i];
    SV cor{cross(bt.ang, jt.ang), cross(bt.ang, jt.lin) + cross(bt.lin, jt.ang)};  // spatial_motion_cross
    if (b.parent == 0) {
      V3 g{0.0, 0.0, gravity_enabled ? GRAVITY : 0.0};
      bias[i] = {cor.ang, g + cor.lin};
    } else {
      bias[i] = {bias[b.parent].ang + cor.ang, bias[b.parent].lin + cor.lin};
    }
  }
  SV wr[GPO_MAX_BODIES + 1];
  for (int i = 1; i <= nb; ++i) {
    SpatialInertia I = inertia_transform(m->bodies[i - 1], w.b2r[i]);
    V3 ia, il, ha, hl;
    mul_inertia(I.moment, I.cross_part, I.mass, bias[i].ang, bias[i].lin, ia, il);
    mul_inertia(I.moment, I.cross_part, I.mass, w.twist[i].ang, w.twist[i].lin, ha, hl);
    // spatial_force_cross(v, Iv), util.rs:62-66
    V3 fa = cross(w.twist[i].ang, ha) + cross(w.twist[i].lin, hl);
    V3 fl = cross(w.twist[i].ang, hl);
    wr[i] = {ia + fa, il + fl};
  }
  for (int i = nb; i >= 1; --i) {
    int p = m->bodies[i - 1].parent;
    if (p != 0) wr[p] = {wr[p].ang + wr[i].ang, wr[p].lin + wr[i].lin};
  }
  double c[GPO_MAX_NV];
  for (int i = 1; i <= nb; ++i) {
    const Body& b = m->bodies[i - 1];
    Jac S = motion_subspace_world(b, w.b2r[i]);
    for (int k = 0; k < S.k; ++k) c[b.voff + k] = dot(S.ang[k], wr[i].ang) + dot(S.lin[k], wr[i].lin);
  }
  // Cholesky (nalgebra: lower L, column by column), then two triangular solves
  double L[GPO_MAX_NV * GPO_MAX_NV];
  std::memcpy(L, w.M, sizeof(double) * n * n);
  for (int j = 0; j < n; ++j) {
    for (int k = 0; k < j; ++k) {
      double f = L[j * n + k];
      for (int r = j; r < n; ++r) L[r * n + j] -= f * L[r * n + k];
    }
    double d = L[j * n + j];
    if (!(d > 0.0)) return 1;
    d = std::sqrt(d);
    L[j * n + j] = d;
    for (int r = j + 1; r < n; ++r) L[r * n + j] /= d;
  }
  double x[GPO_MAX_NV];
  for (int i = 0; i < n; ++i) x[i] = (tau ? tau[i] : 0.0) - c[i];
  for (int i = 0; i < n; ++i) {
    for (int k = 0; k < i; ++k) x[i] -= L[i * n + k] * x[k];
    x[i] /= L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    for (int k = i + 1; k < n; ++k) x[i] -= L[k * n + i] * x[k];
    x[i] /= L[i * n + i];
  }
  for (int i = 0; i < n; ++i) v_free[i] = v[i] + x[i] * dt;
  return 0;
}

double gpo_kinetic_energy(const gpo_mechanism* m, const double* q, const double* v) {
  Work w;
  return kinetic_energy_impl(m, q, v, w);
}

// mechanism.rs:352-362 — height of the FRAME ORIGIN, not of the centre of mass
double gpo_gravitational_energy(const gpo_mechanism* m, const double* q) {
  Work w;
  bodies_to_root(m, q, w);
  double PE = 0.0;
  for (int i = 1; i <= m->nb; ++i) PE += m->bodies[i - 1].mass * GRAVITY * w.b2r[i].t.z;
  return PE;
}

// mechanism.rs:365-377, energy.rs:3-5
double gpo_spring_energy(const gpo_mechanism* m, const double* q) {
  double E = 0.0;
  for (const Body& b : m->bodies)
    if (b.jtype == J_PRIS && b.has_spring) {
      double l = q[b.qoff];
      E += 0.5 * b.spring_k * (l - b.spring_l) * (l - b.spring_l);
    }
  return E;
}

void gpo_poses(const gpo_mechanism* m, const double* q, double* poses) {
  Work w;
  bodies_to_root(m, q, w);
  for (int i = 1; i <= m->nb; ++i) {
    double* o = poses + 7 * (i - 1);
    const Iso& t = w.b2r[i];
    o[0] = t.q.x; o[1] = t.q.y; o[2] = t.q.z; o[3] = t.q.w;
    o[4] = t.t.x; o[5] = t.t.y; o[6] = t.t.z;
  }
}

void gpo_body_twists(const gpo_mechanism* m, const double* q, const double* v, double* twists) {
  Work w;
  bodies_to_root(m, q, w);
  body_twists(m, v, w);
  for (int i = 1; i <= m->nb; ++i) {
    double* o = twists + 6 * (i - 1);
    o[0] = w.twist[i].ang.x; o[1] = w.twist[i].ang.y; o[2] = w.twist[i].ang.z;
    o[3] = w.twist[i].lin.x; o[4] = w.twist[i].lin.y; o[5] = w.twist[i].lin.z;
  }
}

// double_pendulum.rs:14-57 (2x2 LU with partial pivoting via lu_solve)
void gpo_simple_double_pendulum(double m1, double m2, double l1, double l2, double q1, double q2,
                                double q1dot, double q2dot, double vdot_out[2]) {
  double s1 = std::sin(q1), s2 = std::sin(q2), s12 = std::sin(q1 + q2), c2 = std::cos(q2);
  double I2 = m2 * l2 * l2;
  double m12 = I2 + m2 * l1 * l2 * c2;
  double M[4] = {(m1 + m2) * l1 * l1 + I2 + 2. * m2 * l1 * l2 * c2, m12, m12, I2};
  double C[4] = {0.0, -m2 * l1 * l2 * (2. * q1dot + q2dot) * s2,
                 0.5 * m2 * l1 * l2 * (2. * q1dot + q2dot) * s2, -0.5 * m2 * l1 * l2 * q1dot * s2};
  double tau_g[2] = {-GRAVITY * ((m1 + m2) * l1 * s1 + m2 * l2 * s12), -GRAVITY * (m2 * l2 * s12)};
  double bias[2] = {C[0] * q1dot + C[1] * q2dot - tau_g[0], C[2] * q1dot + C[3] * q2dot - tau_g[1]};
  double rhs[2] = {-bias[0], -bias[1]};
  lu_solve(2, M, rhs, vdot_out);
}

void gpo_quat_from_euler(double roll, double pitch, double yaw, double o[4]) {
  Quat q = q_euler(roll, pitch, yaw);
  o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
}
void gpo_quat_from_axis_angle(const double axis[3], double angle, double o[4]) {
  Quat q = q_axis_angle({axis[0], axis[1], axis[2]}, angle);
  o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
}
void gpo_quat_from_scaled_axis(const double aa[3], double o[4]) {
  Quat q = q_scaled_axis({aa[0], aa[1], aa[2]});
  o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w;
}
void gpo_twist_transform(const double iso[7], const double in[6], double out[6]) {
  Iso t{{iso[3], iso[0], iso[1], iso[2]}, {iso[4], iso[5], iso[6]}};
  SV r = twist_transform({{in[0], in[1], in[2]}, {in[3], in[4], in[5]}}, t);
  out[0] = r.ang.x; out[1] = r.ang.y; out[2] = r.ang.z;
  out[3] = r.lin.x; out[4] = r.lin.y; out[5] = r.lin.z;
}

}  // extern "C"
