// gp_models.cpp — the reference's model builders as flat mechanism descriptions.
//
// Restates the numbers of src/helpers.rs, src/builders/mod.rs (SO-101) and
// src/builders/navbot_builder.rs (navbot) of the reference: masses, centres of mass,
// COM-frame inertia tensors moved to the frame origin with the parallel-axis formula the
// reference uses, joint origins (xyz + rpy) and joint lists. Visual meshes and colliders
// are not part of the in-scope path. Host only.
// tests/golden/model_literals.json (extracted from the reference sources by
// tools/extract_reference_literals.py) pins every literal below.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "gp_host.h"

namespace {

constexpr double PI = 3.14159265358979323846;

struct Vec3 {
  double x, y, z;
};

struct Iso {  // quaternion x,y,z,w + translation
  double v[7];
};
Iso iso_identity() { return Iso{{0, 0, 0, 1, 0, 0, 0}}; }
Iso iso_translation(double x, double y, double z) { return Iso{{0, 0, 0, 1, x, y, z}}; }
// Transform3D::new_xyz_rpy (reference spatial/transform.rs:77-94): UnitQuaternion::from_euler_angles
Iso iso_xyz_rpy(double x, double y, double z, double roll, double pitch, double yaw) {
  const double sr = std::sin(roll * 0.5), cr = std::cos(roll * 0.5);
  const double sp = std::sin(pitch * 0.5), cp = std::cos(pitch * 0.5);
  const double sy = std::sin(yaw * 0.5), cy = std::cos(yaw * 0.5);
  return Iso{{sr * cp * cy - cr * sp * sy, cr * sp * cy + sr * cp * sy, cr * cp * sy - sr * sp * cy,
              cr * cp * cy + sr * sp * sy, x, y, z}};
}

struct Inertia {
  double moment[9];
  double cross[3];
  double mass;
};
Inertia diag_inertia(double ix, double iy, double iz, Vec3 cross, double m) {
  return Inertia{{ix, 0, 0, 0, iy, 0, 0, 0, iz}, {cross.x, cross.y, cross.z}, m};
}
// moment_com + m (|c|^2 1 - c c^T), cross_part = m c   (reference builders/mod.rs:37-40)
Inertia com_inertia(double m, Vec3 c, double ixx, double ixy, double ixz, double iyy, double iyz, double izz) {
  const double n2 = c.x * c.x + c.y * c.y + c.z * c.z;
  const double cc[9] = {c.x * c.x, c.x * c.y, c.x * c.z, c.y * c.x, c.y * c.y, c.y * c.z, c.z * c.x, c.z * c.y, c.z * c.z};
  const double mc[9] = {ixx, ixy, ixz, ixy, iyy, iyz, ixz, iyz, izz};
  Inertia I;
  for (int r = 0; r < 3; ++r)
    for (int col = 0; col < 3; ++col) {
      const double id = (r == col) ? 1.0 : 0.0;
      I.moment[3 * r + col] = mc[3 * r + col] + m * (n2 * id - cc[3 * r + col]);
    }
  I.cross[0] = m * c.x;
  I.cross[1] = m * c.y;
  I.cross[2] = m * c.z;
  I.mass = m;
  return I;
}
Inertia sphere_inertia(double m, double r) {  // RigidBody::new_sphere, rigid_body.rs:111-125
  const double i = 2.0 / 5.0 * m * r * r;
  return diag_inertia(i, i, i, {0, 0, 0}, m);
}
Inertia cube_inertia(double m, double l) {  // RigidBody::new_cube, rigid_body.rs:145-156
  const double i = m * l * l / 6.0;
  return diag_inertia(i, i, i, {0, 0, 0}, m);
}

// RigidBody::new_cuboid (rigid_body.rs:160-172): uniform cuboid about its centre, w x d x h along x, y, z
Inertia cuboid_inertia(double m, double w, double d, double h) {
  return diag_inertia(m * (d * d + h * h) / 12.0, m * (w * w + h * h) / 12.0, m * (w * w + d * d) / 12.0, {0, 0, 0}, m);
}
// RigidBody::new_cuboid_at (rigid_body.rs:175-195): the same cuboid with its centre at `com` of the body frame
Inertia cuboid_inertia_at(Vec3 com, double m, double w, double d, double h) {
  return com_inertia(m, com, m * (d * d + h * h) / 12.0, 0, 0, m * (w * w + h * h) / 12.0, 0, m * (w * w + d * d) / 12.0);
}

struct Builder {
  std::vector<int32_t> parent, jtype, has_spring, cp_body;
  std::vector<double> axis, iso, moment, cross, mass, sk, sl, cp_loc, cp_k;

  int add(int parent_id, int jt, Vec3 ax, const Iso& t, const Inertia& I, bool spring = false, double k = 0,
          double l = 0) {
    parent.push_back(parent_id);
    jtype.push_back(jt);
    axis.insert(axis.end(), {ax.x, ax.y, ax.z});
    iso.insert(iso.end(), t.v, t.v + 7);
    moment.insert(moment.end(), I.moment, I.moment + 9);
    cross.insert(cross.end(), I.cross, I.cross + 3);
    mass.push_back(I.mass);
    has_spring.push_back(spring ? 1 : 0);
    sk.push_back(k);
    sl.push_back(l);
    return (int)parent.size();
  }
  void contact(int body, Vec3 loc, double k = 50e3) {  // ContactPoint::new default k, contact.rs:24-30
    cp_body.push_back(body);
    cp_loc.insert(cp_loc.end(), {loc.x, loc.y, loc.z});
    cp_k.push_back(k);
  }
  int create(gp_mechanism** out) const {
    gp_mechanism_desc d{};
    d.n_bodies = (int)parent.size();
    d.parent = parent.data();
    d.joint_type = jtype.data();
    d.axis = axis.data();
    d.init_iso = iso.data();
    d.moment = moment.data();
    d.cross_part = cross.data();
    d.mass = mass.data();
    d.has_spring = has_spring.data();
    d.spring_k = sk.data();
    d.spring_l = sl.data();
    d.n_contact_points = (int)cp_body.size();
    d.cp_body = cp_body.data();
    d.cp_location = cp_loc.data();
    d.cp_k = cp_k.data();
    d.n_halfspaces = 0;
    return gp_mechanism_create(&d, out);
  }
};

const Vec3 X{1, 0, 0}, Y{0, 1, 0}, Z{0, 0, 1}, NEG_Y{0, -1, 0}, NEG_Z{0, 0, -1};
enum { FIXED = GP_JOINT_FIXED, REV = GP_JOINT_REVOLUTE, PRIS = GP_JOINT_PRISMATIC, FLOAT = GP_JOINT_FLOATING };

Inertia inertia_from(const double* moment9, const double* cross3, double m) {
  Inertia I;
  std::memcpy(I.moment, moment9, sizeof(I.moment));
  std::memcpy(I.cross, cross3, sizeof(I.cross));
  I.mass = m;
  return I;
}
Iso iso_from(const double* p) {
  Iso t;
  std::memcpy(t.v, p, sizeof(t.v));
  return t;
}

// helpers.rs:388-421 add_cube_contacts: bottom face then top face
void add_cube_contacts(Builder& b, int body, double l) {
  const double h = l / 2.0;
  b.contact(body, {h, h, -h});
  b.contact(body, {h, -h, -h});
  b.contact(body, {-h, h, -h});
  b.contact(body, {-h, -h, -h});
  b.contact(body, {h, h, h});
  b.contact(body, {h, -h, h});
  b.contact(body, {-h, h, h});
  b.contact(body, {-h, -h, h});
}

// RigidBody::add_cuboid_contacts (rigid_body.rs:216-249): top face then bottom face, x fastest
void add_cuboid_contacts(Builder& b, int body, double w, double d, double h) {
  for (double sz : {1.0, -1.0})
    for (double sy : {1.0, -1.0})
      for (double sx : {-1.0, 1.0}) b.contact(body, {sx * w / 2.0, sy * d / 2.0, sz * h / 2.0});
}
// RigidBody::add_cuboid_contacts_with (rigid_body.rs:251-264): corners about `com`, z fastest
void add_cuboid_contacts_with(Builder& b, int body, Vec3 com, double w, double d, double h) {
  for (double i : {-1.0, 1.0})
    for (double j : {-1.0, 1.0})
      for (double k : {-1.0, 1.0}) b.contact(body, {com.x + i * w / 2.0, com.y + j * d / 2.0, com.z + k * h / 2.0});
}

int bad_params(const char* name, int got, int want) {
  gp::set_error("model '%s' takes %d parameters (or 0 for the reference's own values), got %d", name, want, got);
  return GP_ERR_INVALID;
}

}  // namespace

extern "C" int gp_model_create(const char* name_c, const double* p, int np, gp_mechanism** out) {
  if (!name_c || !out) {
    gp::set_error("gp_model_create: null argument");
    return GP_ERR_INVALID;
  }
  const std::string name(name_c);
  Builder b;

  if (name == "pendulum") {  // helpers.rs:24-46; defaults: dynamics.rs:883-906 (m=5, l=7, axis y)
    double d[23];
    if (np == 0) {
      const double m = 5.0, l = 7.0;
      const double dd[23] = {m, 0, 0, 0, 0, 1.0 / 3.0 * m * l * l, 0, 0, 0, 1.0 / 3.0 * m * l * l, m * l / 2.0, 0, 0,
                             0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
      std::memcpy(d, dd, sizeof(d));
    } else if (np == 23) {
      std::memcpy(d, p, sizeof(d));
    } else {
      return bad_params(name_c, np, 23);
    }
    b.add(0, REV, {d[20], d[21], d[22]}, iso_from(d + 13), inertia_from(d + 1, d + 10, d[0]));
    return b.create(out);
  }
  if (name == "double_pendulum") {  // helpers.rs:49-83; defaults: examples/acrobot.rs:14-34
    double d[30];
    if (np == 0) {
      const double m = 1.0, l = 7.0;
      const double dd[30] = {m, 0, 0, 0, 0, m * l * l, 0, 0, 0, m * l * l, m * l, 0, 0,
                             0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 1, l, 0, 0, 0, -1, 0};
      std::memcpy(d, dd, sizeof(d));
    } else if (np == 30) {
      std::memcpy(d, p, sizeof(d));
    } else {
      return bad_params(name_c, np, 30);
    }
    const Inertia I = inertia_from(d + 1, d + 10, d[0]);
    const Vec3 ax{d[27], d[28], d[29]};
    b.add(0, REV, ax, iso_from(d + 13), I);
    b.add(1, REV, ax, iso_from(d + 20), I);
    return b.create(out);
  }
  if (name == "cart") {  // helpers.rs:86-108; defaults: simulate.rs:177-190 (m=3, l=5, axis x)
    double d[16];
    if (np == 0) {
      const double m = 3.0, l = 5.0;
      const double dd[16] = {m, 0, 0, 0, 0, m * l * l / 12.0, 0, 0, 0, m * l * l / 12.0, 0, 0, 0, 1, 0, 0};
      std::memcpy(d, dd, sizeof(d));
    } else if (np == 16) {
      std::memcpy(d, p, sizeof(d));
    } else {
      return bad_params(name_c, np, 16);
    }
    b.add(0, PRIS, {d[13], d[14], d[15]}, iso_identity(), inertia_from(d + 1, d + 10, d[0]));
    return b.create(out);
  }
  if (name == "cart_pole") {  // helpers.rs:111-149; defaults: examples/cart_pole.rs:27-48
    double d[29];
    if (np == 0) {
      const double m_cart = 1.0, l_cart = 1.0, m_pole = 2.0, l_pole = 1.0;
      const double dd[29] = {m_cart, m_pole,
                             0, 0, 0, 0, m_cart * l_cart * l_cart / 12.0, 0, 0, 0, m_cart * l_cart * l_cart / 12.0,
                             m_pole * l_pole * l_pole, 0, 0, 0, m_pole * l_pole * l_pole, 0, 0, 0, 0,
                             0, 0, 0,
                             0, 0, -l_pole * m_pole,
                             0, -1, 0};
      std::memcpy(d, dd, sizeof(d));
    } else if (np == 29) {
      std::memcpy(d, p, sizeof(d));
    } else {
      return bad_params(name_c, np, 29);
    }
    b.add(0, PRIS, X, iso_identity(), inertia_from(d + 2, d + 20, d[0]));
    b.add(1, REV, {d[26], d[27], d[28]}, iso_identity(), inertia_from(d + 11, d + 23, d[1]));
    return b.create(out);
  }
  if (name == "cube") {  // helpers.rs:151-166; defaults: contact.rs:408-412 (m=3, l=1)
    if (np != 0 && np != 2) return bad_params(name_c, np, 2);
    const double m = np ? p[0] : 3.0, l = np ? p[1] : 1.0;
    b.add(0, FLOAT, Z, iso_identity(), cube_inertia(m, l));
    add_cube_contacts(b, 1, l);
    return b.create(out);
  }
  if (name == "ball") {  // joint/floating.rs:82-104 (m=5, r=1; no collider on this path)
    if (np != 0 && np != 2) return bad_params(name_c, np, 2);
    const double m = np ? p[0] : 5.0, r = np ? p[1] : 1.0;
    b.add(0, FLOAT, Z, iso_identity(), sphere_inertia(m, r));
    return b.create(out);
  }
  if (name == "slip") {  // helpers.rs:308-337 build_SLIP; defaults: contact.rs:839-846 SLIP_hopping
    if (np != 0 && np != 5) return bad_params(name_c, np, 5);
    const double m = np ? p[0] : 0.54, r = np ? p[1] : 0.1, l_rest = np ? p[2] : 0.2,
                 angle = np ? p[3] : 45.0 * PI / 180.0, k_spring = np ? p[4] : 500.0;
    b.add(0, FLOAT, Z, iso_identity(), sphere_inertia(m, r));
    int rc = b.create(out);
    if (rc != GP_OK) return rc;
    double dir[3] = {std::sin(angle), 0.0, -std::cos(angle)};
    const double n = std::sqrt(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
    for (double& x : dir) x /= n;  // UnitVector3::new_normalize
    return gp_mechanism_add_spring_contact(*out, 1, l_rest, dir, k_spring);
  }
  if (name == "rimless_wheel") {  // helpers.rs:168-201; defaults: examples/rimless_wheel.rs:14-21
    if (np != 0 && np != 4) return bad_params(name_c, np, 4);
    const double m_body = np ? p[0] : 10.0, r_body = np ? p[1] : 5.0, l = np ? p[2] : 10.0;
    const int n_foot = np ? (int)p[3] : 8;
    if (n_foot < 1 || n_foot > GP_MAX_CONTACT_POINTS) return bad_params(name_c, np, 4);
    b.add(0, FLOAT, Z, iso_identity(), sphere_inertia(m_body, r_body));
    const double alpha = 2.0 * PI / (double)n_foot / 2.0;
    for (int i = 0; i < n_foot; ++i) {
      // Rotation3::from_axis_angle(y, i * 2 alpha) * (0, 0, -l)
      const double ang = (double)i * 2.0 * alpha;
      const double s = std::sin(ang), c = std::cos(ang);
      b.contact(1, {s * (-l), 0.0, c * (-l)});
    }
    return b.create(out);
  }
  if (name == "hopper") {  // helpers.rs:345-386; defaults: control/hopper_control.rs:144-152
    if (np != 0 && np != 7) return bad_params(name_c, np, 7);
    const double m_foot = np ? p[0] : 1.0, r_foot = np ? p[1] : 1.0, m_hip = np ? p[2] : 0.5,
                 r_hip = np ? p[3] : 1.0, m_body = np ? p[4] : 9.5, r_body = np ? p[5] : 4.0,
                 l_foot_to_hip = np ? p[6] : 1.0;
    b.add(0, FLOAT, Z, iso_identity(), sphere_inertia(m_foot, r_foot));
    b.add(1, PRIS, Z, iso_translation(0, 0, l_foot_to_hip), sphere_inertia(m_hip, r_hip), true, 1e3, 0.0);
    b.add(2, REV, Y, iso_identity(), sphere_inertia(m_body, r_body));
    b.contact(1, {0, 0, 0}, 75e3);
    return b.create(out);
  }
  if (name == "hopper_1d") {  // examples/1D_hopper.rs:14-98
    if (np != 0) return bad_params(name_c, np, 0);
    const double w_body = 5.0, h_body = 0.1, r_leg = 0.5, r_foot = 0.5, body_leg_length = 2.0, leg_foot_length = 10.0;
    const double m_body = 10.0, m_leg = 1.0, m_foot = 1.0;
    const double mx = (w_body * w_body + h_body * h_body) * m_body / 12.0;
    const double mz = (w_body * w_body + w_body * w_body) * m_body / 12.0;
    b.add(0, FLOAT, Z, iso_identity(), diag_inertia(mx, mx, mz, {0, 0, 0}, m_body));
    b.add(1, PRIS, NEG_Z, iso_translation(0, 0, -body_leg_length), sphere_inertia(m_leg, r_leg));
    b.add(2, PRIS, NEG_Z, iso_translation(0, 0, -leg_foot_length), sphere_inertia(m_foot, r_foot));
    b.contact(1, {0, 0, 0});
    b.contact(2, {0, 0, 0});
    b.contact(3, {0, 0, 0});
    return b.create(out);
  }
  if (name == "hopper_2d") {  // helpers.rs:203-306
    if (np != 12) return bad_params(name_c, np, 12);
    const double m_body = p[0], w_body = p[1], h_body = p[2], m_hip = p[3], r_hip = p[4], body_hip_length = p[5],
                 m_piston = p[6], r_piston = p[7], hip_piston_length = p[8], m_leg = p[9], l_leg = p[10],
                 piston_leg_length = p[11];
    const double mx = (w_body * w_body + h_body * h_body) * m_body / 12.0;
    const double mz = (w_body * w_body + w_body * w_body) * m_body / 12.0;
    b.add(0, FLOAT, Z, iso_identity(), diag_inertia(mx, mx, mz, {0, 0, 0}, m_body));
    b.add(1, REV, Y, iso_translation(0, 0, -body_hip_length), sphere_inertia(m_hip, r_hip));
    b.add(2, PRIS, NEG_Z, iso_translation(0, 0, -hip_piston_length), sphere_inertia(m_piston, r_piston));
    const double ml = 1.0 / 3.0 * m_leg * l_leg * l_leg;
    b.add(3, PRIS, NEG_Z, iso_translation(0, 0, -piston_leg_length),
          diag_inertia(ml, ml, 0.0, {0, 0, m_leg * l_leg / 2.0}, m_leg));
    b.contact(4, {0, 0, 0});
    return b.create(out);
  }
  if (name == "quadruped") {  // helpers.rs:423-557
    if (np != 0) return bad_params(name_c, np, 0);
    const double m_body = 5.0, w_body = 1.5, d_body = 0.5, h_body = 0.5, l_leg = 1.0;
    // (sic) the reference adds h_body + h_body instead of h_body^2, helpers.rs:429-430
    const double mx = (d_body * d_body + h_body + h_body) * m_body / 12.0;
    const double my = (w_body * w_body + h_body + h_body) * m_body / 12.0;
    const double mz = (w_body * w_body + d_body * d_body) * m_body / 12.0;
    const double m_hip = 0.5, l_hip = 0.2, m_knee = 0.5, l_knee = 0.2;
    b.add(0, FLOAT, Z, iso_identity(), diag_inertia(mx, my, mz, {0, 0, 0}, m_body));
    const double sx[4] = {w_body / 2.0, w_body / 2.0, -w_body / 2.0, -w_body / 2.0};
    const double sy[4] = {-d_body / 2.0, d_body / 2.0, -d_body / 2.0, d_body / 2.0};  // fr, fl, br, bl
    int hip[4], knee[4];
    for (int k = 0; k < 4; ++k) {
      hip[k] = b.add(1, REV, NEG_Y, iso_translation(sx[k], sy[k], 0.0), cube_inertia(m_hip, l_hip));
      knee[k] = b.add(hip[k], REV, NEG_Y, iso_translation(0, 0, -l_leg / 2.0), cube_inertia(m_knee, l_knee));
    }
    for (int k = 0; k < 4; ++k) b.contact(hip[k], {0, 0, 0});
    for (int k = 0; k < 4; ++k) b.contact(knee[k], {0, 0, 0});
    for (int k = 0; k < 4; ++k) b.contact(knee[k], {0, 0, -l_leg / 2.0}, 10e3);
    return b.create(out);
  }
  if (name == "so101") {  // builders/mod.rs:19-341
    if (np != 0) return bad_params(name_c, np, 0);
    b.add(0, FIXED, Z, iso_identity(),
          com_inertia(0.147, {0.0137179, -5.19711e-05, 0.0334843}, 0.000114686, -4.59787e-07, 4.97151e-06,
                      0.000136117, 9.75275e-08, 0.000130364));
    b.add(1, REV, Z, iso_xyz_rpy(0.0388353, -8.97657e-09, 0.0624, 3.14159, 4.18253e-17, -3.14159),
          com_inertia(0.100006, {-0.0307604, -1.66727e-05, -0.0252713}, 8.3759e-05, 7.55525e-08, -1.16342e-06,
                      8.10403e-05, 1.54663e-07, 2.39783e-05));
    b.add(2, REV, Z, iso_xyz_rpy(-0.0303992, -0.0182778, -0.054, -1.5708, -1.5708, 0.),
          com_inertia(0.103, {-0.0898471, -0.00838224, 0.0184089}, 4.08002e-05, -1.97819e-05, -4.03016e-08,
                      0.000147318, 8.97326e-09, 0.000142487));
    b.add(3, REV, Z, iso_xyz_rpy(-0.11257, -0.028, 1.73763e-16, -3.63608e-16, 8.74301e-16, 1.5708),
          com_inertia(0.104, {-0.0980701, 0.00324376, 0.0182831}, 2.87438e-05, 7.41152e-06, 1.26409e-06,
                      0.000159844, -4.90188e-08, 0.00014529));
    b.add(4, REV, Z, iso_xyz_rpy(-0.1349, 0.0052, 3.62355e-17, 4.02456e-15, 8.67362e-16, -1.5708),
          com_inertia(0.079, {-0.000103312, -0.0386143, 0.0281156}, 3.68263e-05, 1.7893e-08, -5.28128e-08,
                      2.5391e-05, 3.6412e-06, 2.1e-05));
    b.add(5, REV, Z, iso_xyz_rpy(5.55112e-17, -0.0611, 0.0181, 1.5708, 0.0486795, 3.14159),
          com_inertia(0.087, {0.000213627, 0.000245138, -0.025187}, 2.75087e-05, -3.35241e-07, -5.7352e-06,
                      4.33657e-05, -5.17847e-08, 3.45059e-05));
    b.add(6, REV, Z, iso_xyz_rpy(0.0202, 0.0188, -0.0234, 1.5708, -5.24284e-08, -1.41553e-15),
          com_inertia(0.012, {-0.00157495, -0.0300244, 0.0192755}, 6.61427e-06, -3.19807e-07, -5.90717e-09,
                      1.89032e-06, -1.09945e-07, 5.28738e-06));
    return b.create(out);
  }
  if (name == "navbot") {  // builders/navbot_builder.rs:154-790 (loop constraints :792-822 are not on this path)
    if (np != 0) return bad_params(name_c, np, 0);
    const int base = b.add(0, FLOAT, Z, iso_identity(),
                           com_inertia(0.139444, {0.000112099, 0.0274141, -0.0131977}, 4.69114e-05, 4.31285e-12,
                                       5.70403e-10, 5.91667e-05, 1.50774e-06, 8.63546e-05));
    const int leg_left = b.add(base, REV, Z, iso_xyz_rpy(-0.0299877, 0.0274141, -0.0126354, -1.5708, -0.307769, -1.5708),
                               com_inertia(0.00775586, {-0.0134676, -0.000849763, -0.0089392}, 4.45178e-07,
                                           -2.74989e-08, 2.09677e-08, 2.67121e-06, 4.57695e-09, 2.81454e-06));
    const int foot_left = b.add(leg_left, REV, Z, iso_xyz_rpy(-0.052, 0.003, -0.0065, 4.51632e-25, 1.59286e-24, 3.23144e-17),
                                com_inertia(0.0379925, {0.03509, 0.027779, -0.000900737}, 4.76781e-06, -3.17164e-06,
                                            -1.40975e-07, 6.22098e-06, -1.19188e-07, 9.12834e-06));
    b.add(base, REV, Z, iso_xyz_rpy(-0.0406877, 0.0451087, 0.00357876, 1.5708, -0.943592, 1.5708),
          com_inertia(0.00242268, {0.00126762, -0.02515, -0.00146631}, 1.21113e-06, 7.99494e-09, -1.43206e-09,
                      2.07936e-08, -1.00885e-09, 1.22447e-06));  // link_left
    b.add(foot_left, REV, Z, iso_xyz_rpy(0.0391772, 0.0310668, -0.00035, 8.25667e-17, -3.43754e-16, 0.886077),
          com_inertia(0.0155748, {4.83102e-08, -1.61747e-09, -0.00780743}, 1.75465e-06, -3.92314e-13, 3.65986e-12,
                      1.75464e-06, -1.18704e-13, 2.81671e-06));  // wheel_left
    const int leg_right = b.add(base, REV, Z, iso_xyz_rpy(0.0302123, 0.0274141, -0.0126354, 1.5708, -0.307784, -1.5708),
                                com_inertia(0.00775582, {-0.0134676, 0.000849731, -0.0089392}, 4.45176e-07,
                                            2.74979e-08, 2.09678e-08, 2.67121e-06, -4.57706e-09, 2.81453e-06));
    const int foot_right = b.add(leg_right, REV, Z, iso_xyz_rpy(-0.052, -0.003, -0.0065, -3.24841e-15, -1.09622e-15, 0.700637),
                                 com_inertia(0.0379925, {0.00891353, -0.043858, -0.000900737}, 8.49931e-06,
                                             1.24815e-06, -3.10263e-08, 2.48948e-06, 1.81978e-07, 9.12834e-06));
    b.add(base, REV, Z, iso_xyz_rpy(0.0409123, 0.0451087, 0.00357876, -1.5708, -0.943581, 1.5708),
          com_inertia(0.00242268, {0.00126762, 0.02515, -0.00146631}, 1.21113e-06, 7.99494e-09, -1.43206e-09,
                      2.07936e-08, -1.00885e-09, 1.22447e-06));  // link_right
    b.add(foot_right, REV, Z, iso_xyz_rpy(0.00991772, -0.0490065, -0.00035, -4.01485e-15, 1.79841e-15, -1.84321),
          com_inertia(0.0155748, {-1.61747e-09, -4.83102e-08, -0.00780743}, 1.75464e-06, 3.92314e-13, -1.18704e-13,
                      1.75465e-06, -3.65986e-12, 2.81671e-06));  // wheel_right
    return b.create(out);
  }
  if (name == "biped") {  // builders/biped_builder.rs:12-187: floating base + two 6-joint legs, 13 bodies, 18 dof
    if (np != 0) return bad_params(name_c, np, 0);
    const double l1 = 0.05, l2 = 0.2, m = 0.1, w_foot = 0.2;
    const int base = b.add(0, FLOAT, Z, iso_identity(), cuboid_inertia(m, l1, l1, l2));
    for (double side : {1.0, -1.0}) {  // left (+x), then right
      const int pelvis = b.add(base, REV, Z, iso_translation(side * (l1 + l1) / 2.0, 0, 0),
                               cuboid_inertia_at({0, 0, -l2 / 2.0}, m, l1, l1, l2));
      const int hip = b.add(pelvis, REV, NEG_Y, iso_translation(0, 0, -l2),
                            cuboid_inertia_at({side * l2 / 2.0, 0, 0}, m, l2, l1, l1));
      const int thigh = b.add(hip, REV, X, iso_translation(side * l2, 0, 0),
                              cuboid_inertia_at({0, 0, -l2 / 2.0}, m, l1, l1, l2));
      const int calf = b.add(thigh, REV, X, iso_translation(0, 0, -l2), cuboid_inertia_at({0, 0, -l2 / 2.0}, m, l1, l1, l2));
      const int ankle = b.add(calf, REV, X, iso_translation(0, 0, -l2), cuboid_inertia(m, l1, l2, l1));
      const int foot = b.add(ankle, REV, NEG_Y, iso_translation(0, 0, -(l1 + l1) / 2.0), cuboid_inertia(m, w_foot, l2, l1));
      add_cuboid_contacts(b, foot, w_foot, l2, l1);
    }
    return b.create(out);
  }
  if (name == "leg") {  // builders/leg_builder.rs:8-104: floating base + one 5-joint leg, contacts on thigh, calf, foot
    if (np != 0) return bad_params(name_c, np, 0);
    const double l1 = 0.05, l2 = 0.2, m = 0.1, w_foot = 0.2;
    const int base = b.add(0, FLOAT, Z, iso_identity(), cuboid_inertia(m, l1, l1, l2));
    const int pelvis = b.add(base, REV, Z, iso_translation((l1 + l1) / 2.0, 0, 0), cuboid_inertia_at({0, 0, -l2 / 2.0}, m, l1, l1, l2));
    const int hip = b.add(pelvis, REV, NEG_Y, iso_translation(0, 0, -l2), cuboid_inertia_at({l2 / 2.0, 0, 0}, m, l2, l1, l1));
    const Vec3 down{0, 0, -l2 / 2.0};
    const int thigh = b.add(hip, REV, X, iso_translation(l2, 0, 0), cuboid_inertia_at(down, m, l1, l1, l2));
    add_cuboid_contacts_with(b, thigh, down, l1, l1, l2);
    const int calf = b.add(thigh, REV, X, iso_translation(0, 0, -l2), cuboid_inertia_at(down, m, l1, l1, l2));
    add_cuboid_contacts_with(b, calf, down, l1, l1, l2);
    const int foot = b.add(calf, REV, X, iso_translation(0, 0, -l2), cuboid_inertia(m, w_foot, l2, l1));
    add_cuboid_contacts(b, foot, w_foot, l2, l1);
    return b.create(out);
  }
  if (name == "leg_from_foot") {  // builders/leg_builder.rs:106-211: the same leg rooted at its (floating) foot
    if (np != 0) return bad_params(name_c, np, 0);
    const double l1 = 0.05, l2 = 0.2, m = 0.1, w_foot = 0.2;
    const int foot = b.add(0, FLOAT, Z, iso_identity(), cuboid_inertia(m, w_foot, l2, l1));
    add_cuboid_contacts(b, foot, w_foot, l2, l1);
    const Vec3 up{0, 0, l2 / 2.0};
    const int calf = b.add(foot, REV, X, iso_identity(), cuboid_inertia_at(up, m, l1, l1, l2));
    add_cuboid_contacts_with(b, calf, up, l1, l1, l2);
    const int thigh = b.add(calf, REV, X, iso_translation(0, 0, l2), cuboid_inertia_at(up, m, l1, l1, l2));
    add_cuboid_contacts_with(b, thigh, up, l1, l1, l2);
    const int hip = b.add(thigh, REV, X, iso_translation(0, 0, l2), cuboid_inertia_at({-l2 / 2.0, 0, 0}, m, l2, l1, l1));
    const int pelvis = b.add(hip, REV, NEG_Y, iso_translation(-l2, 0, 0), cuboid_inertia_at(up, m, l1, l1, l2));
    b.add(pelvis, REV, Z, iso_translation(0, 0, l2), cuboid_inertia_at({-(l1 + l1) / 2.0, 0, 0}, m, l1, l1, l2));
    return b.create(out);
  }
  gp::set_error("unknown model '%s'", name_c);
  return GP_ERR_INVALID;
}
