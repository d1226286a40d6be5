// gorilla_b200.hpp — header-only C++ facade over the C ABI (gorilla_b200.h) with the reference's
// names and semantics, for n_envs = 1 drop-in use and for writing tests that read like the
// reference's own (tests/cpp/test_reference_suite.cpp).
//
//   reference (Rust)                                       here (C++17)
//   MechanismState::new(treejoints, bodies)    mechanism.rs:62    MechanismState::new_(treejoints, bodies)
//   state.update(&q, &v)                       mechanism.rs:209   state.update(q, v)
//   state.set_joint_q / set_joint_v            mechanism.rs:291   state.set_joint_q / set_joint_v (1-based ids)
//   state.add_halfspace / add_contact_point    mechanism.rs:379   same
//   state.kinetic_energy() / gravitational_energy() / spring_energy() / poses()
//   step(&mut state, dt, &tau, &integrator)    simulate.rs:20     step(state, dt, tau, integrator)
//   simulate(&mut state, T, dt, control_fn, &integrator)  simulate.rs:87   simulate(state, T, dt, control_fn, integrator)
//   dynamics_continuous(&mut state, &tau)      dynamics.rs:322    dynamics_continuous(state, tau)
//   enum Integrator, JointPosition/Velocity/Torque/Acceleration (joint/mod.rs:186-390)
//   helpers::build_* / builders::build_so101 / build_navbot       build_* (via gp_model_create)
//
// Errors: the reference panics; here every failure throws gorilla::Error carrying the status code
// of the C ABI. There is no CPU fallback: without a CUDA device construction of a state throws.
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "gorilla_b200.h"

namespace gorilla {

using Float = double;  // reference src/types.rs:3
constexpr Float GRAVITY = 9.81;
constexpr Float PI = 3.14159265358979323846;
constexpr Float TWO_PI = 2.0 * PI;
inline const std::string WORLD_FRAME = "world";

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
  if (rc != GP_OK) {
    char buf[1024];
    gp_last_error(buf, sizeof(buf));
    throw Error(rc, buf);
  }
}

// ---------------------------------------------------------------- small algebra (nalgebra stand-ins)
struct Vector3 {
  Float x = 0, y = 0, z = 0;
  static Vector3 x_axis() { return {1, 0, 0}; }
  static Vector3 y_axis() { return {0, 1, 0}; }
  static Vector3 z_axis() { return {0, 0, 1}; }
  static Vector3 zeros() { return {0, 0, 0}; }
  Vector3 operator-() const { return {-x, -y, -z}; }
  Vector3 operator+(const Vector3& o) const { return {x + o.x, y + o.y, z + o.z}; }
  Vector3 operator-(const Vector3& o) const { return {x - o.x, y - o.y, z - o.z}; }
  Vector3 operator*(Float s) const { return {x * s, y * s, z * s}; }
  Vector3 scale(Float s) const { return *this * s; }
  Float dot(const Vector3& o) const { return x * o.x + y * o.y + z * o.z; }
  Float norm() const { return std::sqrt(dot(*this)); }
  Float norm_squared() const { return dot(*this); }
};
inline Vector3 vector(Float x, Float y, Float z) { return {x, y, z}; }

struct Matrix3 {
  std::array<Float, 9> m{};  // row-major
  static Matrix3 zeros() { return {}; }
  static Matrix3 identity() { return from_diagonal({1, 1, 1}); }
  static Matrix3 from_diagonal(const Vector3& d) {
    Matrix3 r;
    r.m[0] = d.x; r.m[4] = d.y; r.m[8] = d.z;
    return r;
  }
  static Matrix3 from_diagonal_element(Float v) { return from_diagonal({v, v, v}); }
  static Matrix3 new_(Float a, Float b, Float c, Float d, Float e, Float f, Float g, Float h, Float i) {
    Matrix3 r;
    r.m = {a, b, c, d, e, f, g, h, i};
    return r;
  }
  Matrix3 operator+(const Matrix3& o) const { Matrix3 r; for (int k = 0; k < 9; ++k) r.m[k] = m[k] + o.m[k]; return r; }
  Matrix3 operator*(Float s) const { Matrix3 r; for (int k = 0; k < 9; ++k) r.m[k] = m[k] * s; return r; }
};
// moment about the frame origin from a COM-frame moment: J + m (|c|^2 1 - c c^T)  (rigid_body.rs:135-137)
inline Matrix3 parallel_axis(const Matrix3& moment_com, Float mass, const Vector3& com) {
  const Float n2 = com.norm_squared();
  const Float c[3] = {com.x, com.y, com.z};
  Matrix3 r = moment_com;
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) r.m[3 * a + b] += mass * ((a == b ? n2 : 0.0) - c[a] * c[b]);
  return r;
}

struct UnitQuaternion {
  Float x = 0, y = 0, z = 0, w = 1;
  static UnitQuaternion identity() { return {}; }
  static UnitQuaternion from_euler_angles(Float roll, Float pitch, Float yaw) {  // R = Rz Ry Rx
    const Float sr = std::sin(roll * 0.5), cr = std::cos(roll * 0.5);
    const Float sp = std::sin(pitch * 0.5), cp = std::cos(pitch * 0.5);
    const Float sy = std::sin(yaw * 0.5), cy = std::cos(yaw * 0.5);
    return {sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
            cr * cp * cy + sr * sp * sy};
  }
  static UnitQuaternion from_axis_angle(const Vector3& axis, Float angle) {
    const Float s = std::sin(angle / 2.0), c = std::cos(angle / 2.0);
    return {axis.x * s, axis.y * s, axis.z * s, c};
  }
  static UnitQuaternion from_scaled_axis(const Vector3& aa) {
    const Float n = aa.norm();
    return n == 0.0 ? identity() : from_axis_angle(aa * (1.0 / n), n);
  }
  UnitQuaternion inverse() const { return {-x, -y, -z, w}; }
  UnitQuaternion operator*(const UnitQuaternion& b) const {
    return {w * b.x + x * b.w + y * b.z - z * b.y, w * b.y - x * b.z + y * b.w + z * b.x,
            w * b.z + x * b.y - y * b.x + z * b.w, w * b.w - x * b.x - y * b.y - z * b.z};
  }
  Vector3 operator*(const Vector3& v) const {
    const Vector3 im{x, y, z};
    const Vector3 t = Vector3{im.y * v.z - im.z * v.y, im.z * v.x - im.x * v.z, im.x * v.y - im.y * v.x} * 2.0;
    const Vector3 c{im.y * t.z - im.z * t.y, im.z * t.x - im.x * t.z, im.x * t.y - im.y * t.x};
    return v + t * w + c;
  }
};

struct Isometry3 {
  UnitQuaternion rotation;
  Vector3 translation_;
  static Isometry3 identity() { return {}; }
  static Isometry3 translation(Float x, Float y, Float z) { return {UnitQuaternion::identity(), {x, y, z}}; }
  static Isometry3 rotation_(const Vector3& axisangle) { return {UnitQuaternion::from_scaled_axis(axisangle), {}}; }
  static Isometry3 new_(const Vector3& t, const Vector3& axisangle) {
    return {UnitQuaternion::from_scaled_axis(axisangle), t};
  }
  static Isometry3 from_parts(const Vector3& t, const UnitQuaternion& r) { return {r, t}; }
  Isometry3 operator*(const Isometry3& b) const { return {rotation * b.rotation, translation_ + rotation * b.translation_}; }
};

// ---------------------------------------------------------------- frames, inertia, bodies, joints
struct Transform3D {  // spatial/transform.rs:54-58
  std::string from, to;
  Isometry3 iso;
  static Transform3D new_(const std::string& from, const std::string& to, const Isometry3& iso) { return {from, to, iso}; }
  static Transform3D identity(const std::string& from, const std::string& to) { return {from, to, Isometry3::identity()}; }
  static Transform3D move_x(const std::string& from, const std::string& to, Float a) { return {from, to, Isometry3::translation(a, 0, 0)}; }
  static Transform3D move_z(const std::string& from, const std::string& to, Float a) { return {from, to, Isometry3::translation(0, 0, a)}; }
  static Transform3D move_xyz(const std::string& from, const std::string& to, Float x, Float y, Float z) {
    return {from, to, Isometry3::translation(x, y, z)};
  }
  static Transform3D new_xyz_rpy(const std::string& from, const std::string& to, const std::vector<Float>& xyz,
                                 const std::vector<Float>& rpy) {
    return {from, to, Isometry3::from_parts({xyz[0], xyz[1], xyz[2]}, UnitQuaternion::from_euler_angles(rpy[0], rpy[1], rpy[2]))};
  }
};

struct SpatialInertia {  // inertia.rs:32-37
  std::string frame;
  Matrix3 moment;
  Vector3 cross_part;
  Float mass = 0;
  static SpatialInertia new_(const Matrix3& moment, const Vector3& cross_part, Float mass, const std::string& frame) {
    return {frame, moment, cross_part, mass};
  }
};

struct ContactPoint {  // contact.rs:17-38
  std::string frame;
  Vector3 location;
  Float k = 50e3;
  static ContactPoint new_(const std::string& frame, const Vector3& location) { return {frame, location, 50e3}; }
  static ContactPoint new_with_k(const std::string& frame, const Vector3& location, Float k) { return {frame, location, k}; }
};

struct HalfSpace {  // collision/halfspace.rs:6-37
  Vector3 point, normal;
  Float alpha = 0.9, mu = 0.5;
  static HalfSpace new_(const Vector3& normal, Float distance) { return {normal * distance, normal, 0.9, 0.5}; }
  static HalfSpace new_with_params(const Vector3& normal, Float distance, Float alpha, Float mu) {
    return {normal * distance, normal, alpha, mu};
  }
};

struct RigidBody {  // rigid_body.rs:63-69 (colliders / visual meshes are out of scope)
  SpatialInertia inertia;
  std::vector<ContactPoint> contact_points;
  static RigidBody new_(const SpatialInertia& inertia) { return {inertia, {}}; }
  static RigidBody new_sphere(Float m, Float r, const std::string& frame) {
    const Float i = 2.0 / 5.0 * m * r * r;
    return {{frame, Matrix3::from_diagonal({i, i, i}), {}, m}, {}};
  }
  static RigidBody new_cube(Float m, Float l, const std::string& frame) {
    const Float i = m * l * l / 6.0;
    return {{frame, Matrix3::from_diagonal({i, i, i}), {}, m}, {}};
  }
  static RigidBody new_cuboid(Float m, Float w, Float d, Float h, const std::string& frame) {
    return {{frame, Matrix3::from_diagonal({m * (d * d + h * h) / 12.0, m * (w * w + h * h) / 12.0, m * (w * w + d * d) / 12.0}), {}, m}, {}};
  }
  static RigidBody new_cuboid_at(const Vector3& com, Float m, Float w, Float d, Float h, const std::string& frame) {
    const Matrix3 mc = Matrix3::from_diagonal({m * (d * d + h * h) / 12.0, m * (w * w + h * h) / 12.0, m * (w * w + d * d) / 12.0});
    return {{frame, parallel_axis(mc, m, com), com * m, m}, {}};
  }
  void add_contact_point(const ContactPoint& p) { contact_points.push_back(p); }
};

struct JointSpring {  // joint/prismatic.rs:13-16
  Float k = 0, l = 0;
};

struct Joint {  // enum Joint, joint/mod.rs:22-27
  int type = GP_JOINT_FIXED;
  Transform3D transform_;
  Vector3 axis{0, 0, 1};
  std::optional<JointSpring> spring;
  const Transform3D& transform() const { return transform_; }
};
struct RevoluteJoint {
  static Joint new_(const Transform3D& t, const Vector3& axis) { return {GP_JOINT_REVOLUTE, t, axis, {}}; }
  // joint whose current transform corresponds to position q (revolute.rs:62-73)
  static Joint new_with_q(Float q, Transform3D t, const Vector3& axis) {
    t.iso = t.iso * Isometry3{UnitQuaternion::from_axis_angle(axis, q).inverse(), {}};
    return {GP_JOINT_REVOLUTE, t, axis, {}};
  }
};
struct PrismaticJoint {
  static Joint new_(const Transform3D& t, const Vector3& axis) { return {GP_JOINT_PRISMATIC, t, axis, {}}; }
  static Joint new_with_spring(const Transform3D& t, const Vector3& axis, const JointSpring& s) {
    return {GP_JOINT_PRISMATIC, t, axis, s};
  }
};
struct FloatingJoint {
  static Joint new_(const Transform3D& t) { return {GP_JOINT_FLOATING, t, {0, 0, 1}, {}}; }
};
struct FixedJoint {
  static Joint new_(const Transform3D& t) { return {GP_JOINT_FIXED, t, {0, 0, 1}, {}}; }
};

struct Pose {  // spatial/pose.rs:6-9
  UnitQuaternion rotation;
  Vector3 translation;
  static Pose identity() { return {}; }
};
struct SpatialVector {
  Vector3 angular, linear;
  static SpatialVector zero() { return {}; }
};

// tagged unions of joint/mod.rs:186-390
struct JointPosition {
  enum Kind { None, Float_, Pose_ } kind = None;
  double value = 0;
  gorilla::Pose pose_;
  static JointPosition Float(double v) { JointPosition j; j.kind = Float_; j.value = v; return j; }
  static JointPosition Pose(const gorilla::Pose& p) { JointPosition j; j.kind = Pose_; j.pose_ = p; return j; }
  static JointPosition none() { return {}; }
  double float_() const { if (kind != Float_) throw Error(GP_ERR_INVALID, "JointPosition is not a Float"); return value; }
  const gorilla::Pose& pose() const { if (kind != Pose_) throw Error(GP_ERR_INVALID, "JointPosition is not a Pose"); return pose_; }
};
template <int Tag>
struct JointSpatialOrFloat {
  enum Kind { None, Float_, Spatial_ } kind = None;
  double value = 0;
  SpatialVector spatial_;
  static JointSpatialOrFloat Float(double v) { JointSpatialOrFloat j; j.kind = Float_; j.value = v; return j; }
  static JointSpatialOrFloat Spatial(const SpatialVector& s) { JointSpatialOrFloat j; j.kind = Spatial_; j.spatial_ = s; return j; }
  static JointSpatialOrFloat none() { return {}; }
  double float_() const { if (kind != Float_) throw Error(GP_ERR_INVALID, "joint value is not a Float"); return value; }
  const SpatialVector& spatial() const { if (kind != Spatial_) throw Error(GP_ERR_INVALID, "joint value is not Spatial"); return spatial_; }
};
using JointVelocity = JointSpatialOrFloat<0>;
using JointTorque = JointSpatialOrFloat<1>;
using JointAcceleration = JointSpatialOrFloat<2>;

enum class Integrator {  // integrators.rs:17-23
  SemiImplicitEuler = GP_SEMI_IMPLICIT_EULER,
  RungeKutta2 = GP_RUNGE_KUTTA_2,
  RungeKutta4 = GP_RUNGE_KUTTA_4,
  VelocityStepping = GP_VELOCITY_STEPPING,
  CCDVelocityStepping = GP_CCD_VELOCITY_STEPPING
};

// flat packing of joint/mod.rs:208-303
inline std::vector<Float> to_float_vec(const std::vector<JointPosition>& q) {
  std::vector<Float> out;
  for (const auto& j : q) {
    if (j.kind == JointPosition::Float_) out.push_back(j.value);
    else if (j.kind == JointPosition::Pose_) {
      const auto& p = j.pose_;
      out.insert(out.end(), {p.rotation.x, p.rotation.y, p.rotation.z, p.rotation.w, p.translation.x, p.translation.y, p.translation.z});
    }
  }
  return out;
}
template <int Tag>
std::vector<Float> to_float_vec(const std::vector<JointSpatialOrFloat<Tag>>& v) {
  std::vector<Float> out;
  for (const auto& j : v) {
    if (j.kind == JointSpatialOrFloat<Tag>::Float_) out.push_back(j.value);
    else if (j.kind == JointSpatialOrFloat<Tag>::Spatial_) {
      const auto& s = j.spatial_;
      out.insert(out.end(), {s.angular.x, s.angular.y, s.angular.z, s.linear.x, s.linear.y, s.linear.z});
    }
  }
  return out;
}
inline std::vector<JointPosition> to_joint_pos_vec(const std::vector<Float>& q) {
  std::vector<JointPosition> out;
  for (Float x : q) out.push_back(JointPosition::Float(x));
  return out;
}
inline std::vector<JointVelocity> to_joint_vel_vec(const std::vector<Float>& v) {
  std::vector<JointVelocity> out;
  for (Float x : v) out.push_back(JointVelocity::Float(x));
  return out;
}
inline std::vector<JointTorque> to_joint_torque_vec(const std::vector<Float>& v) {
  std::vector<JointTorque> out;
  for (Float x : v) out.push_back(JointTorque::Float(x));
  return out;
}

// ---------------------------------------------------------------- MechanismState (mechanism.rs:41-58)
class MechanismState {
 public:
  std::vector<Joint> treejoints;
  std::vector<RigidBody> bodies;
  std::vector<size_t> parents;  // parents[i-1] = parent body id of joint i, 0 = world
  std::vector<JointPosition> q;
  std::vector<JointVelocity> v;
  std::vector<HalfSpace> halfspaces;

  // MechanismState::new, mechanism.rs:62-148: joint i's child body must be body i; parents are
  // resolved by frame name; zero initial condition (identity pose for floating joints)
  static MechanismState new_(std::vector<Joint> joints, std::vector<RigidBody> bodies, int device = 0) {
    MechanismState s;
    s.device_ = device;
    const size_t n = joints.size();
    if (bodies.size() != n) throw Error(GP_ERR_INVALID, "number of joints and bodies differ");
    for (size_t i = 0; i < n; ++i) {
      const Joint& j = joints[i];
      if (j.transform().from != bodies[i].inertia.frame)
        throw Error(GP_ERR_INVALID, "joint " + std::to_string(i + 1) + "'s child body is not body " + std::to_string(i + 1));
      if (j.transform().to == WORLD_FRAME) {
        s.parents.push_back(0);
      } else {
        size_t parent = 0;
        for (size_t b = 0; b < n; ++b)
          if (bodies[b].inertia.frame == j.transform().to) { parent = b + 1; break; }
        if (parent == 0 || parent > i)
          throw Error(GP_ERR_INVALID, "joint " + std::to_string(i + 1) + " has no parent body of frame " + j.transform().to);
        s.parents.push_back(parent);
      }
      switch (j.type) {
        case GP_JOINT_REVOLUTE: case GP_JOINT_PRISMATIC:
          s.q.push_back(JointPosition::Float(0.0)); s.v.push_back(JointVelocity::Float(0.0)); break;
        case GP_JOINT_FLOATING:
          s.q.push_back(JointPosition::Pose(Pose::identity())); s.v.push_back(JointVelocity::Spatial(SpatialVector::zero())); break;
        default:
          s.q.push_back(JointPosition::none()); s.v.push_back(JointVelocity::none()); break;
      }
    }
    s.treejoints = std::move(joints);
    s.bodies = std::move(bodies);
    return s;
  }
  // wrap a mechanism made by gp_model_create (the build_* helpers below)
  static MechanismState from_model(const std::string& name, const std::vector<Float>& params = {}, int device = 0) {
    MechanismState s;
    s.device_ = device;
    gp_mechanism* m = nullptr;
    check(gp_model_create(name.c_str(), params.empty() ? nullptr : params.data(), (int)params.size(), &m));
    s.mech_.reset(m, gp_mechanism_destroy);
    gp_mechanism_desc d;
    check(gp_mechanism_get_desc(m, &d));
    for (int i = 0; i < d.n_bodies; ++i) {
      s.parents.push_back((size_t)d.parent[i]);
      s.joint_types_.push_back(d.joint_type[i]);
      switch (d.joint_type[i]) {
        case GP_JOINT_REVOLUTE: case GP_JOINT_PRISMATIC:
          s.q.push_back(JointPosition::Float(0.0)); s.v.push_back(JointVelocity::Float(0.0)); break;
        case GP_JOINT_FLOATING:
          s.q.push_back(JointPosition::Pose(Pose::identity())); s.v.push_back(JointVelocity::Spatial(SpatialVector::zero())); break;
        default:
          s.q.push_back(JointPosition::none()); s.v.push_back(JointVelocity::none()); break;
      }
    }
    s.from_model_ = true;
    return s;
  }

  void update(const std::vector<JointPosition>& q_, const std::vector<JointVelocity>& v_) {  // mechanism.rs:209
    if (q_.size() != q.size() || v_.size() != v.size()) throw Error(GP_ERR_INVALID, "update: wrong number of joints");
    q = q_;
    v = v_;
    state_dirty_ = true;
  }
  void set_joint_q(size_t jointid, const JointPosition& qi) {  // mechanism.rs:291 (1-based)
    if (joint_type(jointid - 1) == GP_JOINT_FIXED) throw Error(GP_ERR_INVALID, "Should not try to set a fixed joint's position");
    q.at(jointid - 1) = qi;
    state_dirty_ = true;
  }
  void set_joint_v(size_t jointid, const JointVelocity& vi) {  // mechanism.rs:328
    v.at(jointid - 1) = vi;
    state_dirty_ = true;
  }
  void add_halfspace(const HalfSpace& h) {  // mechanism.rs:379
    halfspaces.push_back(h);
    if (mech_) {
      const double p[3] = {h.point.x, h.point.y, h.point.z}, n[3] = {h.normal.x, h.normal.y, h.normal.z};
      check(gp_mechanism_add_halfspace(mech_.get(), p, n, h.alpha, h.mu));
      batch_.reset();
    }
  }
  void add_contact_point(const ContactPoint& p) {  // mechanism.rs:384: silently ignored if no body has that frame
    for (size_t b = 0; b < bodies.size(); ++b)
      if (bodies[b].inertia.frame == p.frame) {
        bodies[b].add_contact_point(p);
        if (mech_) {
          const double loc[3] = {p.location.x, p.location.y, p.location.z};
          check(gp_mechanism_add_contact_point(mech_.get(), (int32_t)b + 1, loc, p.k));
          batch_.reset();
        }
        return;
      }
  }
  // add a contact point by body id (for states made with from_model, which carry no frame names)
  void add_contact_point_on_body(size_t bodyid, const Vector3& location, Float k = 50e3) {
    ensure_mechanism();
    const double loc[3] = {location.x, location.y, location.z};
    check(gp_mechanism_add_contact_point(mech_.get(), (int32_t)bodyid, loc, k));
    batch_.reset();
  }

  Float kinetic_energy() { return energies()[0]; }        // mechanism.rs:334
  Float gravitational_energy() { return energies()[1]; }  // mechanism.rs:352
  Float spring_energy() { return energies()[2]; }         // mechanism.rs:365
  std::vector<Pose> poses() {                             // mechanism.rs:403
    sync_to_device();
    std::vector<double> buf(7 * q.size());
    check(gp_batch_poses(batch_.get(), buf.data()));
    std::vector<Pose> out;
    for (size_t i = 0; i < q.size(); ++i) {
      const double* p = &buf[7 * i];
      out.push_back({{p[0], p[1], p[2], p[3]}, {p[4], p[5], p[6]}});
    }
    return out;
  }
  // simulate() (simulate.rs:87-112) when the control law needs no host closure: zero torques, or one of the reference's
  // controllers that the library evaluates in-kernel (GP_CTRL_*, e.g. swingup_acrobot with {m, l}). Every state of the
  // rollout comes back - row 0 the initial state, then the state after each step, like the reference's vectors - but
  // the steps run fused on the device (gp_batch_simulate) instead of one launch and two copies per step.
  std::pair<std::vector<std::vector<JointPosition>>, std::vector<std::vector<JointVelocity>>> simulate_fused(
      Float final_time, Float dt, int integrator, int controller = GP_CTRL_NONE, const std::vector<Float>& ctrl_params = {}) {
    sync_to_device();
    const size_t nq = (size_t)n_q(), nv = (size_t)n_v();
    const int64_t n_counted = gp_simulate_step_count(final_time, dt);
    if (n_counted < 0) throw Error(GP_ERR_INVALID, "simulate: final_time / dt gives no countable number of steps");
    const size_t n = (size_t)n_counted;
    std::vector<double> qf = to_float_vec(q), vf = to_float_vec(v);
    std::vector<double> hq((n + 1) * nq + 1), hv((n + 1) * nv + 1), dummy(1, 0.0);
    int64_t steps = 0;
    check(gp_batch_simulate(batch_.get(), nq ? qf.data() : dummy.data(), nv ? vf.data() : dummy.data(), nullptr, final_time, dt,
                            integrator, controller, ctrl_params.empty() ? nullptr : ctrl_params.data(), (int)ctrl_params.size(),
                            &steps, hq.data(), hv.data()));
    std::vector<std::vector<JointPosition>> qs;
    std::vector<std::vector<JointVelocity>> vs;
    qs.reserve(n + 1);
    vs.reserve(n + 1);
    for (size_t s = 0; s <= (size_t)steps; ++s) {
      unpack(std::vector<double>(hq.begin() + s * nq, hq.begin() + (s + 1) * nq),
             std::vector<double>(hv.begin() + s * nv, hv.begin() + (s + 1) * nv));
      qs.push_back(q);
      vs.push_back(v);
    }
    state_dirty_ = false;  // q, v and the device hold the final state
    return {qs, vs};
  }
  bool has_spring_contacts() const { return false; }  // (this facade builds none; the C ABI has them: gp_mechanism_add_spring_contact)

  // ---- used by step / simulate / dynamics_continuous below
  gp_batch* batch() { sync_to_device(); return batch_.get(); }
  void pull_from_device() {
    std::vector<double> qf(n_q()), vf(n_v());
    check(gp_batch_get_state(batch_.get(), qf.data(), vf.data()));
    unpack(qf, vf);
    state_dirty_ = false;
  }
  int n_q() { ensure_mechanism(); return gp_mechanism_n_q(mech_.get()); }
  int n_v() { ensure_mechanism(); return gp_mechanism_n_v(mech_.get()); }
  int joint_type(size_t i) const { return from_model_ ? joint_types_.at(i) : treejoints.at(i).type; }
  std::string kernel_variant() { ensure_mechanism(); return gp_mechanism_kernel_variant(mech_.get()); }

 private:
  std::shared_ptr<gp_mechanism> mech_;
  std::shared_ptr<gp_batch> batch_;
  std::vector<int> joint_types_;
  bool from_model_ = false;
  bool state_dirty_ = true;
  int device_ = 0;

  void ensure_mechanism() {
    if (mech_) return;
    const int nb = (int)treejoints.size();
    std::vector<int32_t> parent(nb), jt(nb), has_spring(nb), cp_body;
    std::vector<double> axis, iso, moment, cross, mass, sk(nb, 0.0), sl(nb, 0.0), cp_loc, cp_k, hp, hn, ha, hm;
    for (int i = 0; i < nb; ++i) {
      const Joint& j = treejoints[i];
      parent[i] = (int32_t)parents[i];
      jt[i] = j.type;
      axis.insert(axis.end(), {j.axis.x, j.axis.y, j.axis.z});
      const Isometry3& t = j.transform().iso;
      iso.insert(iso.end(), {t.rotation.x, t.rotation.y, t.rotation.z, t.rotation.w, t.translation_.x, t.translation_.y, t.translation_.z});
      const SpatialInertia& I = bodies[i].inertia;
      moment.insert(moment.end(), I.moment.m.begin(), I.moment.m.end());
      cross.insert(cross.end(), {I.cross_part.x, I.cross_part.y, I.cross_part.z});
      mass.push_back(I.mass);
      has_spring[i] = j.spring ? 1 : 0;
      if (j.spring) { sk[i] = j.spring->k; sl[i] = j.spring->l; }
      for (const ContactPoint& c : bodies[i].contact_points) {
        cp_body.push_back(i + 1);
        cp_loc.insert(cp_loc.end(), {c.location.x, c.location.y, c.location.z});
        cp_k.push_back(c.k);
      }
    }
    for (const HalfSpace& h : halfspaces) {
      hp.insert(hp.end(), {h.point.x, h.point.y, h.point.z});
      hn.insert(hn.end(), {h.normal.x, h.normal.y, h.normal.z});
      ha.push_back(h.alpha);
      hm.push_back(h.mu);
    }
    gp_mechanism_desc d{};
    d.n_bodies = nb;
    d.parent = parent.data(); d.joint_type = jt.data(); d.axis = axis.data(); d.init_iso = iso.data();
    d.moment = moment.data(); d.cross_part = cross.data(); d.mass = mass.data();
    d.has_spring = has_spring.data(); d.spring_k = sk.data(); d.spring_l = sl.data();
    d.n_contact_points = (int)cp_body.size(); d.cp_body = cp_body.data(); d.cp_location = cp_loc.data(); d.cp_k = cp_k.data();
    d.n_halfspaces = (int)ha.size(); d.hs_point = hp.data(); d.hs_normal = hn.data(); d.hs_alpha = ha.data(); d.hs_mu = hm.data();
    gp_mechanism* m = nullptr;
    check(gp_mechanism_create(&d, &m));
    mech_.reset(m, gp_mechanism_destroy);
  }
  void sync_to_device() {
    ensure_mechanism();
    if (!batch_) {
      gp_batch* b = nullptr;
      check(gp_batch_create(mech_.get(), 1, device_, &b));
      batch_.reset(b, gp_batch_destroy);
      state_dirty_ = true;
    }
    if (state_dirty_) {
      const std::vector<double> qf = to_float_vec(q), vf = to_float_vec(v);
      check(gp_batch_set_state(batch_.get(), qf.empty() ? nullptr : qf.data(), vf.empty() ? nullptr : vf.data()));
      state_dirty_ = false;
    }
  }
  std::array<Float, 3> energies() {
    sync_to_device();
    std::array<Float, 3> e{};
    check(gp_batch_energy(batch_.get(), &e[0], &e[1], &e[2]));
    return e;
  }
  void unpack(const std::vector<double>& qf, const std::vector<double>& vf) {
    size_t qi = 0, vi = 0;
    for (size_t i = 0; i < q.size(); ++i) {
      switch (joint_type(i)) {
        case GP_JOINT_REVOLUTE: case GP_JOINT_PRISMATIC:
          q[i] = JointPosition::Float(qf[qi++]);
          v[i] = JointVelocity::Float(vf[vi++]);
          break;
        case GP_JOINT_FLOATING:
          q[i] = JointPosition::Pose({{qf[qi], qf[qi + 1], qf[qi + 2], qf[qi + 3]}, {qf[qi + 4], qf[qi + 5], qf[qi + 6]}});
          v[i] = JointVelocity::Spatial({{vf[vi], vf[vi + 1], vf[vi + 2]}, {vf[vi + 3], vf[vi + 4], vf[vi + 5]}});
          qi += 7; vi += 6;
          break;
        default: break;
      }
    }
  }
};

// ---------------------------------------------------------------- dynamics / step / simulate
inline void set_tau(MechanismState& state, const std::vector<JointTorque>& tau) {
  if (tau.empty()) {  // empty tau -> zero torques of the matching variant (simulate.rs:27-48)
    check(gp_batch_set_tau(state.batch(), nullptr));
    return;
  }
  if (tau.size() != state.v.size())  // simulate.rs:39-45 assert_eq!
    throw Error(GP_ERR_INVALID, "joint torques vector length " + std::to_string(tau.size()) +
                                    " and Joint velocity vector v length " + std::to_string(state.v.size()) + " differ!");
  const std::vector<double> tf = to_float_vec(tau);
  check(gp_batch_set_tau(state.batch(), tf.empty() ? nullptr : tf.data()));
}

// dynamics_continuous(state, tau), dynamics.rs:322-364
inline std::vector<JointAcceleration> dynamics_continuous(MechanismState& state, const std::vector<JointTorque>& tau) {
  set_tau(state, tau);
  std::vector<double> vdot(state.n_v() > 0 ? state.n_v() : 1);
  check(gp_batch_dynamics(state.batch(), vdot.data(), nullptr));
  std::vector<JointAcceleration> out;
  size_t i = 0;
  for (size_t j = 0; j < state.v.size(); ++j) {
    switch (state.joint_type(j)) {
      case GP_JOINT_REVOLUTE: case GP_JOINT_PRISMATIC: out.push_back(JointAcceleration::Float(vdot[i++])); break;
      case GP_JOINT_FLOATING:
        out.push_back(JointAcceleration::Spatial({{vdot[i], vdot[i + 1], vdot[i + 2]}, {vdot[i + 3], vdot[i + 4], vdot[i + 5]}}));
        i += 6;
        break;
      default: out.push_back(JointAcceleration::none()); break;
    }
  }
  return out;
}

// step(state, dt, tau, integrator), simulate.rs:20-83
inline std::pair<std::vector<JointPosition>, std::vector<JointVelocity>> step(MechanismState& state, Float dt,
                                                                            const std::vector<JointTorque>& tau,
                                                                            Integrator integrator) {
  set_tau(state, tau);
  check(gp_batch_step(state.batch(), dt, (int)integrator, 1, GP_CTRL_NONE, nullptr, 0));
  state.pull_from_device();
  return {state.q, state.v};
}

// simulate(state, final_time, dt, control_fn, integrator), simulate.rs:87-112
template <class ControlFn>
std::pair<std::vector<std::vector<JointPosition>>, std::vector<std::vector<JointVelocity>>> simulate(
    MechanismState& state, Float final_time, Float dt, ControlFn control_fn, Integrator integrator) {
  Float t = 0.0;
  std::vector<std::vector<JointPosition>> qs{state.q};
  std::vector<std::vector<JointVelocity>> vs{state.v};
  while (t < final_time) {
    const std::vector<JointTorque> tau = control_fn(state);
    auto qv = step(state, dt, tau, integrator);
    qs.push_back(std::move(qv.first));
    vs.push_back(std::move(qv.second));
    t += dt;
  }
  return {qs, vs};
}

// simulate(state, final_time, dt, |_| vec![], integrator): the closure that returns no torques needs no host round trip
inline std::pair<std::vector<std::vector<JointPosition>>, std::vector<std::vector<JointVelocity>>> simulate(
    MechanismState& state, Float final_time, Float dt, Integrator integrator) {
  return state.simulate_fused(final_time, dt, (int)integrator);
}

// ---------------------------------------------------------------- builders (helpers.rs, builders/*.rs)
inline std::vector<Float> pack(std::initializer_list<std::vector<Float>> parts) {
  std::vector<Float> out;
  for (const auto& p : parts) out.insert(out.end(), p.begin(), p.end());
  return out;
}
inline std::vector<Float> flat(const Matrix3& m) { return {m.m.begin(), m.m.end()}; }
inline std::vector<Float> flat(const Vector3& v) { return {v.x, v.y, v.z}; }
inline std::vector<Float> flat(const Isometry3& t) {
  return {t.rotation.x, t.rotation.y, t.rotation.z, t.rotation.w, t.translation_.x, t.translation_.y, t.translation_.z};
}
inline MechanismState build_pendulum(Float mass, const Matrix3& moment, const Vector3& cross_part, const Isometry3& rod_to_world,
                                     const Vector3& axis) {  // helpers.rs:24
  return MechanismState::from_model("pendulum", pack({{mass}, flat(moment), flat(cross_part), flat(rod_to_world), flat(axis)}));
}
inline MechanismState build_double_pendulum(Float mass, const Matrix3& moment, const Vector3& cross_part, const Isometry3& rod1_to_world,
                                            const Isometry3& rod2_to_rod1, const Vector3& axis) {  // helpers.rs:49
  return MechanismState::from_model("double_pendulum", pack({{mass}, flat(moment), flat(cross_part), flat(rod1_to_world),
                                                             flat(rod2_to_rod1), flat(axis)}));
}
inline MechanismState build_cart(Float mass, const Matrix3& moment, const Vector3& cross_part, const Vector3& axis) {  // helpers.rs:86
  return MechanismState::from_model("cart", pack({{mass}, flat(moment), flat(cross_part), flat(axis)}));
}
inline MechanismState build_cart_pole(Float m_cart, Float m_pole, const Matrix3& moment_cart, const Matrix3& moment_pole,
                                      const Vector3& cross_cart, const Vector3& cross_pole, const Vector3& axis_pole) {  // helpers.rs:111
  return MechanismState::from_model("cart_pole", pack({{m_cart, m_pole}, flat(moment_cart), flat(moment_pole), flat(cross_cart),
                                                       flat(cross_pole), flat(axis_pole)}));
}
inline MechanismState build_cube(Float mass, Float length) { return MechanismState::from_model("cube", {mass, length}); }  // helpers.rs:151
inline MechanismState build_rimless_wheel(Float m_body, Float r_body, Float l, size_t n_foot) {  // helpers.rs:168
  return MechanismState::from_model("rimless_wheel", {m_body, r_body, l, (Float)n_foot});
}
inline MechanismState build_hopper(Float m_foot, Float r_foot, Float m_hip, Float r_hip, Float m_body, Float r_body, Float l_foot_to_hip) {
  return MechanismState::from_model("hopper", {m_foot, r_foot, m_hip, r_hip, m_body, r_body, l_foot_to_hip});  // helpers.rs:345
}
inline MechanismState build_quadruped() { return MechanismState::from_model("quadruped"); }  // helpers.rs:423
inline MechanismState build_so101() { return MechanismState::from_model("so101"); }          // builders/mod.rs:252 (meshes are visual only)
inline MechanismState build_navbot() { return MechanismState::from_model("navbot"); }        // builders/navbot_builder.rs:682
inline MechanismState build_biped() { return MechanismState::from_model("biped"); }          // builders/biped_builder.rs:12
inline MechanismState build_leg() { return MechanismState::from_model("leg"); }              // builders/leg_builder.rs:8
inline MechanismState build_leg_from_foot() { return MechanismState::from_model("leg_from_foot"); }  // builders/leg_builder.rs:106

}  // namespace gorilla
