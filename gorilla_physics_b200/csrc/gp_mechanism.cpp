// gp_mechanism.cpp — mechanism handles of the C ABI: validation, flattening to device
// constants, kernel-variant selection. Host only; needs no GPU.
//
// Replaces the bookkeeping half of MechanismState::new (reference src/mechanism.rs:62-148):
// parents and supports become integer tables, joint transforms become the constant matrices
// the kernels combine with sin/cos of the joint angle.
#include <cmath>
#include <cstring>

#include "gp_host.h"
#include "gp_jit.h"
#include "gp_topology.cuh"

namespace gp {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}
const std::string& last_error() { return g_last_error; }

void quat_to_mat_host(const double q[4], double R[9]) {
  const double i = q[0], j = q[1], k = q[2], w = q[3];
  const double ww = w * w, ii = i * i, jj = j * j, kk = k * k;
  const double ij = i * j * 2.0, wk = w * k * 2.0, wj = w * j * 2.0;
  const double ik = i * k * 2.0, jk = j * k * 2.0, wi = w * i * 2.0;
  R[0] = ww + ii - jj - kk; R[1] = ij - wk;           R[2] = wj + ik;
  R[3] = wk + ij;           R[4] = ww - ii + jj - kk; R[5] = jk - wi;
  R[6] = ik - wj;           R[7] = wi + jk;           R[8] = ww - ii - jj + kk;
}

static void mat_mul(const double A[9], const double B[9], double C[9]) {
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}

int finalize_mechanism(gp_mechanism* m) {
  MechParams& P = m->params;
  std::memset(&P, 0, sizeof(P));
  const int nb = m->nb;

  TopoData td{};
  td.nb = nb;
  for (int i = 0; i < nb; ++i) {
    td.parent[i] = m->parent[i] - 1;
    td.jtype[i] = m->joint_type[i];
    const double* a = &m->axis[3 * i];
    const bool scalar = (td.jtype[i] == JRevolute || td.jtype[i] == JPrismatic);
    td.axis[i] = (scalar && a[0] == 0.0 && a[1] == 0.0 && a[2] == 1.0) ? AxZ : AxAny;
  }
  const TopoTables t = make_tables(td);
  if (t.nv > kMaxNV || t.nq > kMaxNQ) {
    set_error("mechanism has n_v=%d n_q=%d, limits are %d / %d", t.nv, t.nq, kMaxNV, kMaxNQ);
    return GP_ERR_LIMIT;
  }
  m->n_q = t.nq;
  m->n_v = t.nv;
  P.nb = nb;
  P.n_q = t.nq;
  P.n_v = t.nv;
  P.n_cp = m->n_cp();
  P.n_hs = m->n_hs();
  for (int i = 0; i < kMaxBodies; ++i) {
    P.parent[i] = t.parent[i];
    P.jtype[i] = t.jtype[i];
    P.qoff[i] = t.qoff[i];
    P.voff[i] = t.voff[i];
    P.depth[i] = t.depth[i];
    P.anc_mask[i] = t.anc_mask[i];
    P.has_children[i] = t.has_children[i];
    P.anchored[i] = t.anchored[i];
    for (int k = 0; k < kMaxBodies; ++k) P.anc_at[i][k] = t.anc_at[i][k];
  }
  for (int k = 0; k < kMaxNV; ++k) P.dof_body[k] = t.dof_body[k];

  for (int i = 0; i < nb; ++i) {
    const double* iso = &m->init_iso[7 * i];
    double E0[9];
    quat_to_mat_host(iso, E0);
    const double* a = &m->axis[3 * i];
    for (int k = 0; k < 4; ++k) P.iq[i][k] = iso[k];
    for (int k = 0; k < 3; ++k) {
      P.r0[i][k] = iso[4 + k];
      P.axis[i][k] = a[k];
      P.Ea[i][k] = E0[3 * k] * a[0] + E0[3 * k + 1] * a[1] + E0[3 * k + 2] * a[2];
    }
    if (td.jtype[i] == JRevolute) {
      // E0 * Rot(a, q) = Cm + cos(q) A + sin(q) B,  Cm = E0 a a^T, A = E0 - Cm, B = E0 [a]x
      const double aaT[9] = {a[0] * a[0], a[0] * a[1], a[0] * a[2], a[1] * a[0], a[1] * a[1],
                             a[1] * a[2], a[2] * a[0], a[2] * a[1], a[2] * a[2]};
      const double K[9] = {0.0, -a[2], a[1], a[2], 0.0, -a[0], -a[1], a[0], 0.0};
      double Cm[9], B[9];
      mat_mul(E0, aaT, Cm);
      mat_mul(E0, K, B);
      for (int k = 0; k < 9; ++k) {
        P.Cm[i][k] = Cm[k];
        P.A[i][k] = E0[k] - Cm[k];
        P.B[i][k] = B[k];
      }
    } else {
      for (int k = 0; k < 9; ++k) P.Cm[i][k] = E0[k];
    }
    const double* J = &m->moment[9 * i];
    P.J[i][0] = J[0]; P.J[i][1] = J[1]; P.J[i][2] = J[2];
    P.J[i][3] = J[4]; P.J[i][4] = J[5]; P.J[i][5] = J[8];
    for (int k = 0; k < 3; ++k) P.mc[i][k] = m->cross_part[3 * i + k];
    P.mass[i] = m->mass[i];
    P.has_spring[i] = m->has_spring[i];
    P.spring_k[i] = m->spring_k[i];
    P.spring_l[i] = m->spring_l[i];
    P.armature[i] = m->armature[i];
  }
  // constants of the composite-inertia pass (see gp_params.h): children come after their parents,
  // so one leaf-to-root sweep sees every subtree mass before it is needed
  for (int i = 0; i < nb; ++i) {
    P.msub[i] = P.mass[i];
    for (int k = 0; k < 6; ++k) P.Jacc0[i][k] = P.J[i][k];
    for (int k = 0; k < 3; ++k) {
      P.cacc0[i][k] = P.mc[i][k];
      P.r0x2[i][k] = 2.0 * P.r0[i][k];
    }
    const double* a = P.axis[i];
    const double* J = P.J[i];
    const double Ja[3] = {J[0] * a[0] + J[1] * a[1] + J[2] * a[2], J[1] * a[0] + J[3] * a[1] + J[4] * a[2],
                          J[2] * a[0] + J[4] * a[1] + J[5] * a[2]};
    const double ac = a[0] * P.mc[i][0] + a[1] * P.mc[i][1] + a[2] * P.mc[i][2];
    P.ne_a[i][0] = a[1] * Ja[2] - a[2] * Ja[1];
    P.ne_a[i][1] = a[2] * Ja[0] - a[0] * Ja[2];
    P.ne_a[i][2] = a[0] * Ja[1] - a[1] * Ja[0];
    const double aa = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
    for (int k = 0; k < 3; ++k) P.ne_l[i][k] = a[k] * ac - P.mc[i][k] * aa;
  }
  for (int i = nb - 1; i >= 0; --i) {
    const int p = P.parent[i];
    if (p < 0) continue;
    P.msub[p] += P.msub[i];
    if (P.jtype[i] == JRevolute || P.jtype[i] == JFixed) {
      const double m = P.msub[i];
      const double* r = P.r0[i];
      const double rr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
      P.Jacc0[p][0] += m * (rr - r[0] * r[0]);
      P.Jacc0[p][1] += -m * r[0] * r[1];
      P.Jacc0[p][2] += -m * r[0] * r[2];
      P.Jacc0[p][3] += m * (rr - r[1] * r[1]);
      P.Jacc0[p][4] += -m * r[1] * r[2];
      P.Jacc0[p][5] += m * (rr - r[2] * r[2]);
      for (int k = 0; k < 3; ++k) P.cacc0[p][k] += m * r[k];
    }
  }

  // inverse of the constant mass matrix of a single floating body (gp_params.h root_inv)
  P.root_inv_ok = 0;
  for (int k = 0; k < 21; ++k) P.root_inv[k] = 0.0;
  if (nb == 1 && P.jtype[0] == JFloating) {
    const double* J = P.J[0];
    const double* c = P.mc[0];
    const double ms = P.mass[0];
    long double H[6][6] = {{J[0], J[1], J[2], 0, -c[2], c[1]}, {J[1], J[3], J[4], c[2], 0, -c[0]},
                           {J[2], J[4], J[5], -c[1], c[0], 0}, {0, c[2], -c[1], ms, 0, 0},
                           {-c[2], 0, c[0], 0, ms, 0},         {c[1], -c[0], 0, 0, 0, ms}};
    long double L[6][6] = {};
    bool spd = true;
    for (int j = 0; j < 6 && spd; ++j) {  // Cholesky H = L L^T
      long double d = H[j][j];
      for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
      if (!(d > 0.0L)) { spd = false; break; }
      L[j][j] = sqrtl(d);
      for (int i = j + 1; i < 6; ++i) {
        long double t = H[i][j];
        for (int k = 0; k < j; ++k) t -= L[i][k] * L[j][k];
        L[i][j] = t / L[j][j];
      }
    }
    if (spd) {
      long double Li[6][6] = {};  // L^-1 (lower triangular), then H^-1 = L^-T L^-1
      for (int j = 0; j < 6; ++j) {
        Li[j][j] = 1.0L / L[j][j];
        for (int i = j + 1; i < 6; ++i) {
          long double t = 0.0L;
          for (int k = j; k < i; ++k) t -= L[i][k] * Li[k][j];
          Li[i][j] = t / L[i][i];
        }
      }
      for (int r = 0; r < 6; ++r)
        for (int cc = 0; cc <= r; ++cc) {
          long double t = 0.0L;
          for (int k = r; k < 6; ++k) t += Li[k][r] * Li[k][cc];
          P.root_inv[r * (r + 1) / 2 + cc] = (double)t;
        }
      P.root_inv_ok = 1;
    }
  }

  // contact points are stored body-major already
  int c = 0;
  for (int b = 0; b < nb; ++b) {
    P.cp_begin[b] = c;
    while (c < m->n_cp() && m->cp_body[c] == b + 1) ++c;
  }
  for (int b = nb; b <= kMaxBodies; ++b) P.cp_begin[b] = c;
  for (int k = 0; k < m->n_cp(); ++k) {
    for (int d = 0; d < 3; ++d) P.cp_loc[k][d] = m->cp_location[3 * k + d];
    const double k_A = m->cp_k[k], k_B = 50e3;  // reference contact.rs:328, :274
    P.cp_k[k] = k_A * k_B / (k_A + k_B);
  }
  for (int h = 0; h < m->n_hs(); ++h) {
    for (int d = 0; d < 3; ++d) {
      P.hs_point[h][d] = m->hs_point[3 * h + d];
      P.hs_normal[h][d] = m->hs_normal[3 * h + d];
    }
    P.hs_off[h] = P.hs_point[h][0] * P.hs_normal[h][0] + P.hs_point[h][1] * P.hs_normal[h][1] +
                  P.hs_point[h][2] * P.hs_normal[h][2];
    P.hs_alpha[h] = m->hs_alpha[h];
    P.hs_mu[h] = m->hs_mu[h];
  }

  P.n_sc = m->n_sc();
  for (int s = 0; s < m->n_sc(); ++s) {
    P.sc_body[s] = m->sc_body[s] - 1;
    P.sc_l_rest[s] = m->sc_l_rest[s];
    P.sc_k[s] = m->sc_k[s];
    for (int d = 0; d < 3; ++d) P.sc_dir[s][d] = m->sc_direction[3 * s + d];
  }

  // kernel variant (gp_kernel_mode): a shipped specialisation whose signature matches; else one compiled
  // at run time for this very tree (gp_jit.cpp); else the run-time-topology kernel.
  // (spring contacts: only kernels that implement them)
  int nvar = 0;
  const KernelTable* const* vars = all_variants(&nvar);
  const KernelTable* generic = vars[nvar - 1];
  m->table = nullptr;
  if (m->kernel_mode == GP_KERNEL_AUTO || m->kernel_mode == GP_KERNEL_SHIPPED) {
    for (int k = 0; k < nvar - 1; ++k)
      if (topo_matches(vars[k]->topo, td) && (m->n_sc() == 0 || vars[k]->springs)) {
        m->table = vars[k];
        break;
      }
  }
  if (!m->table && (m->kernel_mode == GP_KERNEL_AUTO || m->kernel_mode == GP_KERNEL_JIT)) {
    std::string why;
    if (jit_available(&why)) {
      m->table = jit_table(td, jit_policy_for(m, td));
    } else if (m->kernel_mode == GP_KERNEL_JIT) {
      set_error("GP_KERNEL_JIT: run-time specialisation is not available: %s", why.c_str());
      m->table = generic;
      m->revision++;
      return GP_ERR_JIT;
    }
  }
  if (!m->table) m->table = generic;
  m->revision++;
  return GP_OK;
}

}  // namespace gp

using namespace gp;

extern "C" {

int gp_abi_version(void) { return GP_ABI_VERSION; }

size_t gp_last_error(char* buf, size_t len) {
  const std::string& e = last_error();
  if (buf && len > 0) {
    const size_t n = e.size() < len - 1 ? e.size() : len - 1;
    std::memcpy(buf, e.data(), n);
    buf[n] = '\0';
  }
  return e.size();
}

// Every constant of a description ends up as an operand of the step kernels: a NaN or an infinity among them would
// not fail anywhere, it would turn every environment's state into NaNs at the first step. (The reference does not
// check either - f64 fields - but a C caller's uninitialised array is a likelier accident than a Rust one.)
static bool all_finite(const double* x, int n) {
  for (int k = 0; k < n; ++k)
    if (!std::isfinite(x[k])) return false;
  return true;
}

int gp_mechanism_create(const gp_mechanism_desc* d, gp_mechanism** out) {
  if (!d || !out) {
    set_error("gp_mechanism_create: null argument");
    return GP_ERR_INVALID;
  }
  *out = nullptr;
  const int nb = d->n_bodies;
  if (nb < 1 || nb > kMaxBodies) {
    set_error("n_bodies=%d outside [1, %d]", nb, kMaxBodies);
    return nb > kMaxBodies ? GP_ERR_LIMIT : GP_ERR_INVALID;
  }
  if (d->n_contact_points < 0 || d->n_contact_points > kMaxCP || d->n_halfspaces < 0 ||
      d->n_halfspaces > kMaxHS) {
    set_error("n_contact_points=%d (max %d) / n_halfspaces=%d (max %d)", d->n_contact_points, kMaxCP,
              d->n_halfspaces, kMaxHS);
    return GP_ERR_LIMIT;
  }
  if (!d->parent || !d->joint_type || !d->axis || !d->init_iso || !d->moment || !d->cross_part || !d->mass) {
    set_error("gp_mechanism_create: null array in description");
    return GP_ERR_INVALID;
  }
  if ((d->n_halfspaces > 0 && (!d->hs_point || !d->hs_normal || !d->hs_alpha || !d->hs_mu)) ||
      (d->n_contact_points > 0 && (!d->cp_body || !d->cp_location || !d->cp_k)) ||
      (d->n_spring_contacts > 0 && (!d->sc_body || !d->sc_l_rest || !d->sc_direction || !d->sc_k))) {
    set_error("gp_mechanism_create: a count is > 0 but its arrays are null");
    return GP_ERR_INVALID;
  }
  if (d->n_spring_contacts < 0 || d->n_spring_contacts > kMaxSC) {
    set_error("n_spring_contacts=%d (max %d)", d->n_spring_contacts, kMaxSC);
    return GP_ERR_LIMIT;
  }
  gp_mechanism* m = new gp_mechanism();
  m->nb = nb;
  for (int i = 0; i < nb; ++i) {
    const int p = d->parent[i];
    // reference mechanism.rs:98-125: the parent frame must be the world or an earlier body
    if (p < 0 || p > i) {
      set_error("joint %d has no parent body %d (parents must precede children)", i + 1, p);
      delete m;
      return GP_ERR_INVALID;
    }
    const int jt = d->joint_type[i];
    if (jt < GP_JOINT_FIXED || jt > GP_JOINT_FLOATING) {
      set_error("joint %d: unknown joint type %d", i + 1, jt);
      delete m;
      return GP_ERR_INVALID;
    }
    const double* a = d->axis + 3 * i;
    if (jt == GP_JOINT_REVOLUTE || jt == GP_JOINT_PRISMATIC) {
      const double n2 = a[0] * a[0] + a[1] * a[1] + a[2] * a[2];
      if (!(std::fabs(n2 - 1.0) < 1e-9)) {
        set_error("joint %d: axis is not a unit vector (|a|^2 = %.17g)", i + 1, n2);
        delete m;
        return GP_ERR_INVALID;
      }
    }
    const double* q = d->init_iso + 7 * i;
    const double qn = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    if (!(std::fabs(qn - 1.0) < 1e-9)) {
      set_error("joint %d: init_iso rotation is not a unit quaternion (|q|^2 = %.17g)", i + 1, qn);
      delete m;
      return GP_ERR_INVALID;
    }
    const double* J = d->moment + 9 * i;
    const bool sp_i = d->has_spring && d->has_spring[i] != 0;
    if (!all_finite(q, 7) || !all_finite(J, 9) || !all_finite(d->cross_part + 3 * i, 3) || !std::isfinite(d->mass[i]) ||
        d->mass[i] < 0.0 || (sp_i && d->spring_k && !std::isfinite(d->spring_k[i])) ||
        (sp_i && d->spring_l && !std::isfinite(d->spring_l[i]))) {
      set_error("body %d: joint origin, inertia, mass (>= 0) and joint spring must be finite numbers", i + 1);
      delete m;
      return GP_ERR_INVALID;
    }
    double jmax = 0.0;
    for (int k = 0; k < 9; ++k) jmax = std::fmax(jmax, std::fabs(J[k]));
    if (std::fabs(J[1] - J[3]) > 1e-12 * jmax || std::fabs(J[2] - J[6]) > 1e-12 * jmax ||
        std::fabs(J[5] - J[7]) > 1e-12 * jmax) {
      set_error("body %d: moment of inertia is not symmetric", i + 1);
      delete m;
      return GP_ERR_INVALID;
    }
    m->parent.push_back(p);
    m->joint_type.push_back(jt);
    m->axis.insert(m->axis.end(), a, a + 3);
    m->init_iso.insert(m->init_iso.end(), q, q + 7);
    m->moment.insert(m->moment.end(), J, J + 9);
    m->cross_part.insert(m->cross_part.end(), d->cross_part + 3 * i, d->cross_part + 3 * i + 3);
    m->mass.push_back(d->mass[i]);
    const bool sp = d->has_spring && d->has_spring[i] != 0;
    m->has_spring.push_back(sp ? 1 : 0);
    m->spring_k.push_back(sp && d->spring_k ? d->spring_k[i] : 0.0);
    m->spring_l.push_back(sp && d->spring_l ? d->spring_l[i] : 0.0);
    const double arm = d->armature ? d->armature[i] : 0.0;
    if (!(arm >= 0.0) || (arm != 0.0 && jt != GP_JOINT_REVOLUTE)) {
      // reference joint/mod.rs:105-108: "armature is only supported on revolute joints"
      set_error("joint %d: armature must be >= 0 and is only supported on revolute joints", i + 1);
      delete m;
      return GP_ERR_INVALID;
    }
    m->armature.push_back(arm);
  }
  for (int h = 0; h < d->n_halfspaces; ++h) {
    const double* nn = d->hs_normal + 3 * h;
    const double n2 = nn[0] * nn[0] + nn[1] * nn[1] + nn[2] * nn[2];
    if (!(std::fabs(n2 - 1.0) < 1e-9)) {  // the reference takes a UnitVector3 (halfspace.rs:6-11)
      set_error("halfspace %d: normal is not a unit vector (|n|^2 = %.17g)", h, n2);
      delete m;
      return GP_ERR_INVALID;
    }
    if (!all_finite(d->hs_point + 3 * h, 3) || !std::isfinite(d->hs_alpha[h]) || !std::isfinite(d->hs_mu[h])) {
      set_error("halfspace %d: point, alpha and mu must be finite numbers", h);
      delete m;
      return GP_ERR_INVALID;
    }
    m->hs_point.insert(m->hs_point.end(), d->hs_point + 3 * h, d->hs_point + 3 * h + 3);
    m->hs_normal.insert(m->hs_normal.end(), d->hs_normal + 3 * h, d->hs_normal + 3 * h + 3);
    m->hs_alpha.push_back(d->hs_alpha[h]);
    m->hs_mu.push_back(d->hs_mu[h]);
  }
  int rc = finalize_mechanism(m);
  for (int s = 0; rc == GP_OK && s < d->n_spring_contacts; ++s)
    rc = gp_mechanism_add_spring_contact(m, d->sc_body[s], d->sc_l_rest[s], d->sc_direction + 3 * s, d->sc_k[s]);
  for (int c = 0; rc == GP_OK && c < d->n_contact_points; ++c)
    rc = gp_mechanism_add_contact_point(m, d->cp_body[c], d->cp_location + 3 * c, d->cp_k[c]);
  if (rc != GP_OK) {
    delete m;
    return rc;
  }
  *out = m;
  return GP_OK;
}

void gp_mechanism_destroy(gp_mechanism* m) { delete m; }
int gp_mechanism_n_bodies(const gp_mechanism* m) { return m ? m->nb : 0; }
int gp_mechanism_n_q(const gp_mechanism* m) { return m ? m->n_q : 0; }
int gp_mechanism_n_v(const gp_mechanism* m) { return m ? m->n_v : 0; }
int gp_mechanism_n_contact_points(const gp_mechanism* m) { return m ? m->n_cp() : 0; }
int gp_mechanism_n_halfspaces(const gp_mechanism* m) { return m ? m->n_hs() : 0; }

int gp_mechanism_get_desc(const gp_mechanism* m, gp_mechanism_desc* o) {
  if (!m || !o) {
    set_error("gp_mechanism_get_desc: null argument");
    return GP_ERR_INVALID;
  }
  o->n_bodies = m->nb;
  o->parent = m->parent.data();
  o->joint_type = m->joint_type.data();
  o->axis = m->axis.data();
  o->init_iso = m->init_iso.data();
  o->moment = m->moment.data();
  o->cross_part = m->cross_part.data();
  o->mass = m->mass.data();
  o->has_spring = m->has_spring.data();
  o->spring_k = m->spring_k.data();
  o->spring_l = m->spring_l.data();
  o->n_contact_points = m->n_cp();
  o->cp_body = m->cp_body.data();
  o->cp_location = m->cp_location.data();
  o->cp_k = m->cp_k.data();
  o->n_halfspaces = m->n_hs();
  o->hs_point = m->hs_point.data();
  o->hs_normal = m->hs_normal.data();
  o->hs_alpha = m->hs_alpha.data();
  o->hs_mu = m->hs_mu.data();
  o->armature = m->armature.data();
  o->n_spring_contacts = m->n_sc();
  o->sc_body = m->sc_body.data();
  o->sc_l_rest = m->sc_l_rest.data();
  o->sc_direction = m->sc_direction.data();
  o->sc_k = m->sc_k.data();
  return GP_OK;
}

int gp_mechanism_add_halfspace(gp_mechanism* m, const double point[3], const double normal[3], double alpha,
                               double mu) {
  if (!m || !point || !normal) {
    set_error("gp_mechanism_add_halfspace: null argument");
    return GP_ERR_INVALID;
  }
  const double hn2 = normal[0] * normal[0] + normal[1] * normal[1] + normal[2] * normal[2];
  if (!(std::fabs(hn2 - 1.0) < 1e-9)) {  // the reference takes a UnitVector3 (halfspace.rs:6-11)
    set_error("halfspace normal is not a unit vector (|n|^2 = %.17g)", hn2);
    return GP_ERR_INVALID;
  }
  if (!all_finite(point, 3) || !std::isfinite(alpha) || !std::isfinite(mu)) {
    set_error("halfspace point, alpha and mu must be finite numbers");
    return GP_ERR_INVALID;
  }
  if (m->n_hs() >= kMaxHS) {
    set_error("more than %d halfspaces", kMaxHS);
    return GP_ERR_LIMIT;
  }
  m->hs_point.insert(m->hs_point.end(), point, point + 3);
  m->hs_normal.insert(m->hs_normal.end(), normal, normal + 3);
  m->hs_alpha.push_back(alpha);
  m->hs_mu.push_back(mu);
  return finalize_mechanism(m);
}

int gp_mechanism_add_contact_point(gp_mechanism* m, int32_t body, const double location[3], double k) {
  if (!m || !location) {
    set_error("gp_mechanism_add_contact_point: null argument");
    return GP_ERR_INVALID;
  }
  if (body < 1 || body > m->nb) {  // reference mechanism.rs:384-391 silently ignores unknown frames
    set_error("contact point on unknown body %d", body);
    return GP_ERR_INVALID;
  }
  if (!all_finite(location, 3) || !std::isfinite(k)) {
    set_error("contact point location and stiffness must be finite numbers");
    return GP_ERR_INVALID;
  }
  if (m->n_cp() >= kMaxCP) {
    set_error("more than %d contact points", kMaxCP);
    return GP_ERR_LIMIT;
  }
  // keep body-major order, appended after the body's existing points
  size_t pos = 0;
  while (pos < m->cp_body.size() && m->cp_body[pos] <= body) ++pos;
  m->cp_body.insert(m->cp_body.begin() + pos, body);
  m->cp_location.insert(m->cp_location.begin() + 3 * pos, location, location + 3);
  m->cp_k.insert(m->cp_k.begin() + pos, k);
  return finalize_mechanism(m);
}

int gp_mechanism_add_spring_contact(gp_mechanism* m, int32_t body, double l_rest, const double direction[3], double k) {
  if (!m || !direction) {
    set_error("gp_mechanism_add_spring_contact: null argument");
    return GP_ERR_INVALID;
  }
  if (body < 1 || body > m->nb) {
    set_error("spring contact on unknown body %d", body);
    return GP_ERR_INVALID;
  }
  const double n2 = direction[0] * direction[0] + direction[1] * direction[1] + direction[2] * direction[2];
  if (!(std::fabs(n2 - 1.0) < 1e-9)) {
    set_error("spring contact direction is not a unit vector");
    return GP_ERR_INVALID;
  }
  if (!std::isfinite(l_rest) || !std::isfinite(k)) {
    set_error("spring contact rest length and stiffness must be finite numbers");
    return GP_ERR_INVALID;
  }
  if (m->n_sc() >= kMaxSC) {
    set_error("more than %d spring contacts", kMaxSC);
    return GP_ERR_LIMIT;
  }
  m->sc_body.push_back(body);
  m->sc_l_rest.push_back(l_rest);
  m->sc_direction.insert(m->sc_direction.end(), direction, direction + 3);
  m->sc_k.push_back(k);
  return finalize_mechanism(m);
}

int gp_mechanism_n_spring_contacts(const gp_mechanism* m) { return m ? m->n_sc() : 0; }

int gp_mechanism_supports(const gp_mechanism* m, int32_t* out) {
  if (!m || !out) {
    set_error("gp_mechanism_supports: null argument");
    return GP_ERR_INVALID;
  }
  for (int j = 0; j < m->nb; ++j)
    for (int i = 0; i < m->nb; ++i) out[j * m->nb + i] = (m->params.anc_mask[i] >> j) & 1u;
  return GP_OK;
}

const char* gp_mechanism_kernel_variant(const gp_mechanism* m) {
  return (m && m->table) ? m->table->name : "";
}

int gp_mechanism_set_kernel_mode(gp_mechanism* m, int mode) {
  if (!m || mode < GP_KERNEL_AUTO || mode > GP_KERNEL_SHIPPED) {
    set_error("gp_mechanism_set_kernel_mode: bad argument");
    return GP_ERR_INVALID;
  }
  const int before = m->kernel_mode;
  m->kernel_mode = mode;
  const int rc = finalize_mechanism(m);
  if (rc != GP_OK) {
    m->kernel_mode = before;
    const std::string msg = last_error();
    finalize_mechanism(m);
    set_error("%s", msg.c_str());
  }
  return rc;
}

int gp_jit_available(void) { return jit_available(nullptr) ? 1 : 0; }

size_t gp_jit_cache_dir(char* buf, size_t len) {
  const std::string d = jit_cache_dir();
  if (buf && len > 0) {
    const size_t n = d.size() < len - 1 ? d.size() : len - 1;
    std::memcpy(buf, d.data(), n);
    buf[n] = '\0';
  }
  return d.size();
}

int gp_mechanism_precompile(const gp_mechanism* m, unsigned kinds, int* n_compiled) {
  if (!m || !m->table) {
    set_error("gp_mechanism_precompile: null mechanism");
    return GP_ERR_INVALID;
  }
  // (the contact mode a step launch of this mechanism uses: gp_batch.cu contact_mode)
  const int contact = m->n_sc() > 0 ? 2 : ((m->n_cp() == 0 || m->n_hs() == 0) ? 0 : (m->n_hs() == 1 ? 1 : 2));
  return jit_precompile(m->table, contact, kinds, n_compiled);
}

}  // extern "C"
