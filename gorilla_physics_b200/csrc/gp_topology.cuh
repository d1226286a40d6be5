// gp_topology.cuh — compile-time and run-time views of a mechanism's tree.
//
// The dynamics core is written once against a `Topo` policy. A StaticTopo<Spec> answers
// every structural question (parent, joint type, offsets, ancestor sets) from tables the
// front end evaluates at compile time, so that after full unrolling every body loop index,
// every branch on joint type and every mass-matrix sparsity test is a literal and all
// per-body arrays live in registers. DynTopo answers the same questions from MechParams at
// run time (arrays land in local memory) and serves any tree the specialisations don't cover.
//
// Tables replace the reference's string-matched `parents` and HashSet `supports`
// (src/mechanism.rs:98-125).
#pragma once
#include "gp_params.h"

#if defined(__CUDACC__)
#define GP_HD __host__ __device__ __forceinline__
#if defined(GP_HOST_DEBUG)
// tools/host_debug.cu only: lets the developer single-step the device functions on the CPU of a
// machine without a GPU. Never defined when building libgorilla_b200.so.
#define GP_D __host__ __device__ __forceinline__
#else
#define GP_D __device__ __forceinline__
#endif
#else
#define GP_HD inline
#define GP_D inline
#endif

namespace gp {

struct TopoTables {
  int nb, nq, nv;
  int parent[kMaxBodies];
  int jtype[kMaxBodies];
  int axis[kMaxBodies];
  int qoff[kMaxBodies];
  int voff[kMaxBodies];
  int depth[kMaxBodies];
  int anc_at[kMaxBodies][kMaxBodies];
  unsigned anc_mask[kMaxBodies];
  int has_children[kMaxBodies];
  int anchored[kMaxBodies];  // rigidly attached to the world through fixed joints only: zero twist
  int dof_body[kMaxNV];
  int q_body[kMaxNQ];  // body that owns entry k of the flat q vector
};

constexpr int joint_nq(int t) { return t == JFloating ? 7 : (t == JFixed ? 0 : 1); }
constexpr int joint_nv(int t) { return t == JFloating ? 6 : (t == JFixed ? 0 : 1); }

// derive every table from the signature (host + compile time)
constexpr TopoTables make_tables(const TopoData& d) {
  TopoTables t{};
  t.nb = d.nb;
  int qo = 0, vo = 0;
  for (int i = 0; i < kMaxBodies; ++i) {
    t.parent[i] = -1;
    t.jtype[i] = JFixed;
    t.axis[i] = AxAny;
    t.depth[i] = 0;
    t.anc_mask[i] = 0u;
    t.has_children[i] = 0;
    t.anchored[i] = 0;
    t.qoff[i] = 0;
    t.voff[i] = 0;
    for (int k = 0; k < kMaxBodies; ++k) t.anc_at[i][k] = -1;
  }
  for (int k = 0; k < kMaxNV; ++k) t.dof_body[k] = 0;
  for (int k = 0; k < kMaxNQ; ++k) t.q_body[k] = 0;
  for (int i = 0; i < d.nb; ++i) {
    t.parent[i] = d.parent[i];
    t.jtype[i] = d.jtype[i];
    t.axis[i] = d.axis[i];
    t.qoff[i] = qo;
    t.voff[i] = vo;
    for (int k = 0; k < joint_nv(d.jtype[i]); ++k)
      if (vo + k < kMaxNV) t.dof_body[vo + k] = i;
    for (int k = 0; k < joint_nq(d.jtype[i]); ++k)
      if (qo + k < kMaxNQ) t.q_body[qo + k] = i;
    qo += joint_nq(d.jtype[i]);
    vo += joint_nv(d.jtype[i]);
    int c = i, k = 0;
    while (c >= 0) {
      t.anc_at[i][k++] = c;
      t.anc_mask[i] |= (1u << c);
      c = d.parent[c];
    }
    t.depth[i] = k;
    if (d.parent[i] >= 0) t.has_children[d.parent[i]] = 1;
    t.anchored[i] = (d.jtype[i] == JFixed && (d.parent[i] < 0 || t.anchored[d.parent[i]])) ? 1 : 0;
  }
  t.nq = qo;
  t.nv = vo;
  return t;
}

// Compile-time lookups. A run-time index i is resolved through a chain of selects over
// template-constant K whose values are constant expressions, packed into integers where a
// second index is needed. After full unrolling i is a literal and the chain folds away; there
// is never a table object in device code (a local constexpr table per call site makes the
// optimiser chew through thousands of allocas).
template <class F, int K, int N>
GP_HD constexpr unsigned long long select_chain(int i) {
  if constexpr (K + 1 >= N) {
    return F::template at<K>();
  } else {
    return (i == K) ? F::template at<K>() : select_chain<F, K + 1, N>(i);
  }
}

// compile-time index + guaranteed unrolling: f(IC<0>{}), f(IC<1>{}), ... Each call instantiates
// the (generic) lambda with a literal index, so topology lookups fold in the front end and the
// per-body arrays are only ever indexed by constants. `#pragma unroll` alone is not reliable for
// the large body loops (the 9-body trees were left rolled, with their arrays in local memory).
template <int I>
struct IC {
  static constexpr int value = I;
  GP_HD constexpr operator int() const { return I; }
};
// the literal behind a loop index: IC<I> -> I; a run-time int (run-time-topology loops) -> -1
template <class T>
struct ic_of {
  static constexpr int value = -1;
};
template <int I>
struct ic_of<IC<I>> {
  static constexpr int value = I;
};
template <int B, int E, class F>
GP_HD void static_for(F&& f) {
  if constexpr (B < E) {
    f(IC<B>{});
    static_for<B + 1, E>(f);
  }
}
template <int B, int E, class F>
GP_HD void static_rfor(F&& f) {  // E-1, E-2, ..., B
  if constexpr (B < E) {
    f(IC<E - 1>{});
    static_rfor<B, E - 1>(f);
  }
}
// body loops: unrolled at compile time for static topologies, run-time loops otherwise. A sided topology
// (StaticTopo<Spec, SIDE>, warp-pair mapping: see below) visits only the bodies / dofs of its own half of
// the tree plus the root, which both halves carry.
template <class Topo, class F>
GP_HD void for_bodies(const MechParams& P, F&& f) {
  if constexpr (Topo::kStatic) {
    static_for<0, Topo::NB>([&](auto ii) {
      if constexpr (Topo::mine_body(ic_of<decltype(ii)>::value)) f(ii);
    });
  } else {
    for (int i = 0; i < P.nb; ++i) f(i);
  }
}
template <class Topo, class F>
GP_HD void for_bodies_reverse(const MechParams& P, F&& f) {
  if constexpr (Topo::kStatic) {
    static_rfor<0, Topo::NB>([&](auto ii) {
      if constexpr (Topo::mine_body(ic_of<decltype(ii)>::value)) f(ii);
    });
  } else {
    for (int i = P.nb - 1; i >= 0; --i) f(i);
  }
}
template <class Topo, class F>
GP_HD void for_dofs(const MechParams& P, F&& f) {
  if constexpr (Topo::kStatic) {
    static_for<0, Topo::kNVreal>([&](auto kk) {
      if constexpr (Topo::mine_dof(ic_of<decltype(kk)>::value)) f(kk);
    });
  } else {
    for (int k = 0; k < P.n_v; ++k) f(k);
  }
}
template <class Topo, class F>
GP_HD void for_dofs_reverse(const MechParams& P, F&& f) {
  if constexpr (Topo::kStatic) {
    static_rfor<0, Topo::kNVreal>([&](auto kk) {
      if constexpr (Topo::mine_dof(ic_of<decltype(kk)>::value)) f(kk);
    });
  } else {
    for (int k = P.n_v - 1; k >= 0; --k) f(k);
  }
}
// entries of the flat q / v vectors (loads, stores, integration): every entry for a whole-tree thread; a sided
// thread sees the entries of its own bodies and of the root
template <class Topo, class F>
GP_HD void for_q_entries(const MechParams& P, F&& f) {
  if constexpr (Topo::kStatic) {
    static_for<0, Topo::kNQreal>([&](auto kk) {
      if constexpr (Topo::mine_body(Topo::tables().q_body[ic_of<decltype(kk)>::value])) f(kk);
    });
  } else {
    for (int k = 0; k < P.n_q; ++k) f(k);
  }
}
template <class Topo, class F>
GP_HD void for_v_entries(const MechParams& P, F&& f) {
  for_dofs<Topo>(P, f);
}

// Warp-pair mapping (Spec::side_mask() != 0): the tree is cut at its root body (body 0) into two halves, each a
// set of whole child subtrees of the root. Two warps advance the same 32 environments, one half each
// (StaticTopo<Spec, 0> and StaticTopo<Spec, 1>), both carrying the root: half the bodies, columns of H and
// live state per thread. Per time step they exchange, through shared memory behind a 64-thread named
// barrier, what meets at the root: the halves' contributions to the root's composite inertia and force
// (leaf-to-root pass) and to the root block of H and of the right-hand side (factorisation); everything
// else - kinematics, contact, the columns of their own dofs, the back-substitution - is local.
// side_mask: bit i set = body i belongs to half 1 (bit 0, the root, stays clear).
template <class Spec, int SIDE = -1>
struct StaticTopo {
  static constexpr bool kStatic = true;
  static constexpr TopoTables tables() { return make_tables(Spec::data()); }
  static constexpr int NB = tables().nb;
  static constexpr int NQ = tables().nq > 0 ? tables().nq : 1;
  static constexpr int NV = tables().nv > 0 ? tables().nv : 1;
  static constexpr int kNVreal = tables().nv;
  static constexpr int kNQreal = tables().nq;
  // warp-pair mapping: SIDE -1 = the thread owns the whole tree; 0 / 1 = one half (plus the root)
  static constexpr int kSide = SIDE;
  static constexpr bool kSided = SIDE >= 0;
  static constexpr unsigned kSideMask = Spec::side_mask();
  static constexpr bool kHasSides = kSideMask != 0u;  // the step kernel of this topology runs warp pairs
  static constexpr int kRootNV = joint_nv(tables().jtype[0]);
  // floating base + four (hip, knee) revolute chains, the tree of build_quadruped (helpers.rs:423): the only one the
  // in-kernel QuadrupedTrottingController is compiled into
  static constexpr bool quadruped_like() {
    const TopoTables t = tables();
    if (t.nb != 9 || t.jtype[0] != JFloating || t.parent[0] != -1) return false;
    for (int i = 1; i < 9; ++i)
      if (t.jtype[i] != JRevolute || t.parent[i] != ((i & 1) ? 0 : i - 1)) return false;
    return true;
  }
  static constexpr bool kQuadrupedLike = quadruped_like();
  using Whole = StaticTopo<Spec, -1>;
  template <int S> using Half = StaticTopo<Spec, S>;
  // body i is advanced by this thread / is stored by this thread (the root is computed by both halves and
  // stored by half 0)
  GP_HD static constexpr bool mine_body(int i) { return SIDE < 0 || i == 0 || (int)((kSideMask >> i) & 1u) == SIDE; }
  GP_HD static constexpr bool owns_body(int i) { return SIDE < 0 || (i == 0 ? SIDE == 0 : (int)((kSideMask >> i) & 1u) == SIDE); }
  GP_HD static constexpr bool mine_dof(int k) { return mine_body(tables().dof_body[k]); }
  GP_HD static constexpr bool owns_dof(int k) { return owns_body(tables().dof_body[k]); }
  GP_HD static constexpr bool owns_q(int k) { return owns_body(tables().q_body[k]); }
  static constexpr int kUnroll = 64;
  static const char* name() { return Spec::name(); }
  static constexpr int min_blocks(int contact) { return Spec::min_blocks(contact); }
  static constexpr int kBlockSize = Spec::block_size();
  static constexpr bool kBatchedSinCos = Spec::batched_sincos();
  // stateful SpringContact legs (reference contact.rs:74-94, :133-186) compiled into the general-contact
  // kernels of this topology: the single floating body (SLIP, helpers.rs:308-337)
  static constexpr bool kSprings = Spec::springs();
  // per-lane list of the points in contact (gp_dynamics.cuh): kContactList = some body of the topology
  // uses it (the step kernel then stages the contact points in shared memory), contact_list = this body does
  static constexpr bool kContactList = Spec::contact_list(-1);
  static constexpr bool kTickets = Spec::tickets();  // step kernel compiled with ticket mode (gp_kernels.cuh)
  GP_HD static constexpr bool contact_list(const MechParams&, int body, int /*n_points*/) { return Spec::contact_list(body); }
  // factorise H column by column inside the leaf-to-root pass (gp_dynamics.cuh): pays where the kernel
  // has registers to spare, i.e. everywhere but the 14-dof trees
  static constexpr bool kColumnsInPass2 = tables().nv < 12 && !kSided;

  struct FParent { template <int K> static constexpr unsigned long long at() { return (unsigned long long)(tables().parent[K] + 1); } };
  struct FJtype { template <int K> static constexpr unsigned long long at() { return (unsigned long long)tables().jtype[K]; } };
  struct FAxis { template <int K> static constexpr unsigned long long at() { return (unsigned long long)tables().axis[K]; } };
  struct FQoff { template <int K> static constexpr unsigned long long at() { return (unsigned long long)tables().qoff[K]; } };
  struct FVoff { template <int K> static constexpr unsigned long long at() { return (unsigned long long)tables().voff[K]; } };
  struct FDepth { template <int K> static constexpr unsigned long long at() { return (unsigned long long)tables().depth[K]; } };
  struct FChildren { template <int K> static constexpr unsigned long long at() { return (unsigned long long)tables().has_children[K]; } };
  struct FAnchored { template <int K> static constexpr unsigned long long at() { return (unsigned long long)tables().anchored[K]; } };
  // ancestors of body K packed 4 bits each (value + 1), k-th nibble = k-th ancestor
  struct FAncRow {
    template <int K> static constexpr unsigned long long at() {
      unsigned long long w = 0;
      for (int k = 0; k < kMaxBodies; ++k) w |= (unsigned long long)(tables().anc_at[K][k] + 1) << (4 * k);
      return w;
    }
  };
  struct FDofBody { template <int K> static constexpr unsigned long long at() { return (unsigned long long)tables().dof_body[K]; } };
  // bit a set: dof a's body is an ancestor-or-self of dof K's body
  struct FDofAnc {
    template <int K> static constexpr unsigned long long at() {
      unsigned long long w = 0;
      for (int a = 0; a < tables().nv; ++a)
        if ((tables().anc_mask[tables().dof_body[K]] >> tables().dof_body[a]) & 1u) w |= 1ull << a;
      return w;
    }
  };

  // bodies that are ancestor-or-self of dof K's body, as a bit mask
  struct FDofBodies { template <int K> static constexpr unsigned long long at() { return (unsigned long long)tables().anc_mask[tables().dof_body[K]]; } };

  // loop bound helper: static topologies always loop to the compile-time maximum (with a
  // guard inside) so that every trip count is a literal when the unroller first sees the loop
  GP_HD static constexpr int lim(int /*runtime_bound*/, int static_bound) { return static_bound; }
  GP_HD static constexpr int nb(const MechParams&) { return NB; }
  GP_HD static constexpr int nq(const MechParams&) { return tables().nq; }
  GP_HD static constexpr int nv(const MechParams&) { return tables().nv; }
  GP_HD static constexpr int parent(const MechParams&, int i) { return (int)select_chain<FParent, 0, NB>(i) - 1; }
  GP_HD static constexpr int jtype(const MechParams&, int i) { return (int)select_chain<FJtype, 0, NB>(i); }
  GP_HD static constexpr int axis_kind(const MechParams&, int i) { return (int)select_chain<FAxis, 0, NB>(i); }
  GP_HD static constexpr int qoff(const MechParams&, int i) { return (int)select_chain<FQoff, 0, NB>(i); }
  GP_HD static constexpr int voff(const MechParams&, int i) { return (int)select_chain<FVoff, 0, NB>(i); }
  GP_HD static constexpr int depth(const MechParams&, int i) { return (int)select_chain<FDepth, 0, NB>(i); }
  GP_HD static constexpr int anc_at(const MechParams&, int i, int k) {
    return (int)((select_chain<FAncRow, 0, NB>(i) >> (4 * k)) & 15ull) - 1;
  }
  GP_HD static constexpr bool has_children(const MechParams&, int i) { return select_chain<FChildren, 0, NB>(i) != 0ull; }
  GP_HD static constexpr bool anchored(const MechParams&, int i) { return select_chain<FAnchored, 0, NB>(i) != 0ull; }
  GP_HD static constexpr int dof_body(const MechParams&, int k) { return (int)select_chain<FDofBody, 0, NV>(k); }
  // dof a's body is an ancestor-or-self of dof b's body (mass-matrix entry (b,a) is structurally non-zero)
  GP_HD static constexpr bool dof_anc(const MechParams&, int a, int b) {
    return ((select_chain<FDofAnc, 0, NV>(b) >> a) & 1ull) != 0ull;
  }
  // dof d belongs to the subtree rooted at body i (i itself included)
  GP_HD static constexpr bool dof_under(const MechParams&, int d, int i) {
    return ((select_chain<FDofBodies, 0, NV>(d) >> i) & 1ull) != 0ull;
  }
};

struct DynTopo {
  static constexpr bool kStatic = false;
  static constexpr int kSide = -1;
  static constexpr bool kSided = false;
  static constexpr bool kHasSides = false;
  static constexpr int kRootNV = 0;
  static constexpr bool kQuadrupedLike = false;
  using Whole = DynTopo;
  GP_HD static constexpr bool mine_body(int) { return true; }
  GP_HD static constexpr bool owns_body(int) { return true; }
  GP_HD static constexpr bool mine_dof(int) { return true; }
  GP_HD static constexpr bool owns_dof(int) { return true; }
  GP_HD static constexpr bool owns_q(int) { return true; }
  static constexpr int NB = kMaxBodies;
  static constexpr int NQ = kMaxNQ;
  static constexpr int NV = kMaxNV;
  static constexpr int kNVreal = kMaxNV;
  static constexpr int kBlockSize = 128;
  static constexpr bool kBatchedSinCos = false;
  static constexpr bool kSprings = true;
  static constexpr bool kContactList = true;
  static constexpr bool kTickets = true;
  // run-time topology: bodies with many points take the list (both loops exist once in the rolled body)
  GP_HD static bool contact_list(const MechParams&, int /*body*/, int n_points) { return n_points >= 4; }
  static constexpr bool kColumnsInPass2 = true;
  static constexpr int kUnroll = 1;
  static const char* name() { return "generic"; }

  GP_HD static int lim(int runtime_bound, int /*static_bound*/) { return runtime_bound; }
  GP_HD static int nb(const MechParams& P) { return P.nb; }
  GP_HD static int nq(const MechParams& P) { return P.n_q; }
  GP_HD static int nv(const MechParams& P) { return P.n_v; }
  GP_HD static int parent(const MechParams& P, int i) { return P.parent[i]; }
  GP_HD static int jtype(const MechParams& P, int i) { return P.jtype[i]; }
  GP_HD static int axis_kind(const MechParams&, int) { return AxAny; }
  GP_HD static int qoff(const MechParams& P, int i) { return P.qoff[i]; }
  GP_HD static int voff(const MechParams& P, int i) { return P.voff[i]; }
  GP_HD static int depth(const MechParams& P, int i) { return P.depth[i]; }
  GP_HD static int anc_at(const MechParams& P, int i, int k) { return P.anc_at[i][k]; }
  GP_HD static bool has_children(const MechParams& P, int i) { return P.has_children[i] != 0; }
  GP_HD static bool anchored(const MechParams& P, int i) { return P.anchored[i] != 0; }
  GP_HD static int dof_body(const MechParams& P, int k) { return P.dof_body[k]; }
  GP_HD static bool dof_anc(const MechParams& P, int a, int b) {
    return ((P.anc_mask[P.dof_body[b]] >> P.dof_body[a]) & 1u) != 0u;
  }
  GP_HD static bool dof_under(const MechParams& P, int d, int i) { return ((P.anc_mask[P.dof_body[d]] >> i) & 1u) != 0u; }
};

// ---------------------------------------------------------------- shipped specialisations
// One per tree the reference's configs use (BASELINE.json configs).
#define GP_R JRevolute
#define GP_P JPrismatic
#define GP_F JFloating
#define GP_X JFixed

struct SpecPendulum {  // helpers.rs:24 build_pendulum
  static constexpr TopoData data() { return {1, {-1}, {GP_R}, {AxAny}}; }
  static const char* name() { return "pendulum_R"; }
  static constexpr int min_blocks(int) { return 1; }
  static constexpr int block_size() { return 128; }
  static constexpr bool batched_sincos() { return true; }
  static constexpr bool springs() { return false; }
  static constexpr bool tickets() { return false; }
  static constexpr bool contact_list(int) { return false; }
  static constexpr unsigned side_mask() { return 0u; }
};
struct SpecDoublePendulum {  // helpers.rs:49 build_double_pendulum (acrobot, configs 1-2)
  static constexpr TopoData data() { return {2, {-1, 0}, {GP_R, GP_R}, {AxAny, AxAny}}; }
  static const char* name() { return "double_pendulum_RR"; }
  static constexpr int min_blocks(int) { return 6; }  // 80 registers, 24 warps per SM: +3.6 % over 150 registers / 12 warps
  static constexpr int block_size() { return 128; }
  static constexpr bool batched_sincos() { return true; }
  static constexpr bool springs() { return false; }
  static constexpr bool tickets() { return false; }
  static constexpr bool contact_list(int) { return false; }
  static constexpr unsigned side_mask() { return 0u; }
};
struct SpecCartPole {  // helpers.rs:111 build_cart_pole (config 2)
  static constexpr TopoData data() { return {2, {-1, 0}, {GP_P, GP_R}, {AxAny, AxAny}}; }
  static const char* name() { return "cart_pole_PR"; }
  static constexpr int min_blocks(int) { return 1; }
  static constexpr int block_size() { return 128; }
  static constexpr bool batched_sincos() { return true; }
  static constexpr bool springs() { return false; }
  static constexpr bool tickets() { return false; }
  static constexpr bool contact_list(int) { return false; }
  static constexpr unsigned side_mask() { return 0u; }
};
struct SpecSO101 {  // builders/mod.rs:252 build_so101: fixed base + 6 revolute(+z) chain (config 3)
  static constexpr TopoData data() {
    return {7, {-1, 0, 1, 2, 3, 4, 5}, {GP_X, GP_R, GP_R, GP_R, GP_R, GP_R, GP_R},
            {AxAny, AxZ, AxZ, AxZ, AxZ, AxZ, AxZ}};
  }
  static const char* name() { return "so101_X6Rz"; }
  // one 256-thread block per SM: its 8 warps walk the unrolled step body in lockstep (per-step
  // barrier) and share instruction-cache lines; 246 registers, no spills (profiles/r1_tuning.md)
  static constexpr int min_blocks(int) { return 1; }
  static constexpr int block_size() { return 256; }
  static constexpr bool batched_sincos() { return true; }
  static constexpr bool springs() { return false; }
  static constexpr bool tickets() { return false; }
  static constexpr bool contact_list(int) { return false; }
  static constexpr unsigned side_mask() { return 0u; }
};
struct SpecFloating {  // helpers.rs:151 build_cube, :168 build_rimless_wheel, ball (config 4a)
  static constexpr TopoData data() { return {1, {-1}, {GP_F}, {AxAny}}; }
  static const char* name() { return "floating_F"; }
  static constexpr int min_blocks(int) { return 4; }  // 128 registers: four 128-thread blocks per SM
  static constexpr int block_size() { return 128; }
  static constexpr bool batched_sincos() { return true; }
  static constexpr bool springs() { return true; }
  static constexpr bool tickets() { return true; }  // rimless wheel, 256 K: +3 %
  static constexpr bool contact_list(int) { return true; }  // cube corners, rimless-wheel spokes
  static constexpr unsigned side_mask() { return 0u; }
};
struct SpecHopper1D {  // examples/1D_hopper.rs: floating + 2 prismatic chain (config 4b)
  static constexpr TopoData data() { return {3, {-1, 0, 1}, {GP_F, GP_P, GP_P}, {AxAny, AxAny, AxAny}}; }
  static const char* name() { return "hopper1d_FPP"; }
  static constexpr int min_blocks(int) { return 4; }  // 128 registers (332 B of spills), 16 warps per SM: +2.6 % over 208 / 8
  static constexpr int block_size() { return 128; }
  static constexpr bool batched_sincos() { return true; }
  static constexpr bool springs() { return false; }
  static constexpr bool tickets() { return false; }
  static constexpr bool contact_list(int) { return false; }
  static constexpr unsigned side_mask() { return 0u; }
};
struct SpecHopper {  // helpers.rs:345 build_hopper: floating foot + prismatic(spring) + revolute
  static constexpr TopoData data() { return {3, {-1, 0, 1}, {GP_F, GP_P, GP_R}, {AxAny, AxAny, AxAny}}; }
  static const char* name() { return "hopper_FPR"; }
  static constexpr int min_blocks(int) { return 1; }
  static constexpr int block_size() { return 128; }
  static constexpr bool batched_sincos() { return true; }
  static constexpr bool springs() { return false; }
  static constexpr bool tickets() { return false; }
  static constexpr bool contact_list(int) { return false; }
  static constexpr unsigned side_mask() { return 0u; }
};
struct SpecQuadruped {  // helpers.rs:423 build_quadruped: floating + 4 x (hip, knee) revolute(-y)
  static constexpr TopoData data() {
    return {9, {-1, 0, 1, 0, 3, 0, 5, 0, 7}, {GP_F, GP_R, GP_R, GP_R, GP_R, GP_R, GP_R, GP_R, GP_R},
            {AxAny, AxAny, AxAny, AxAny, AxAny, AxAny, AxAny, AxAny, AxAny}};
  }
  static const char* name() { return "quadruped_F8R"; }
  static constexpr int min_blocks(int) { return 1; }
  static constexpr int block_size() { return 256; }
  // 8 angles in flight at the start of the step cost this kernel more in spills than the shared
  // literals save (profiles/r1_tuning.md)
  static constexpr bool batched_sincos() { return false; }
  static constexpr bool springs() { return false; }
  static constexpr bool tickets() { return true; }  // 64 K environments = 1.73 waves: +12 %
  static constexpr bool contact_list(int) { return false; }  // one or two points per body: the list only costs registers (-22 %)
#ifndef GP_NO_SIDES
  static constexpr unsigned side_mask() { return 0x1e0u; }  // half 0: legs 1-2, 3-4; half 1: legs 5-6, 7-8
#else
  static constexpr unsigned side_mask() { return 0u; }
#endif
};
struct SpecNavbot {  // navbot_builder.rs:682 build_navbot: floating + 8 revolute(+z) (config 5)
  static constexpr TopoData data() {
    return {9, {-1, 0, 1, 0, 2, 0, 5, 0, 6}, {GP_F, GP_R, GP_R, GP_R, GP_R, GP_R, GP_R, GP_R, GP_R},
            {AxAny, AxZ, AxZ, AxZ, AxZ, AxZ, AxZ, AxZ, AxZ}};
  }
  static const char* name() { return "navbot_F8Rz"; }
  static constexpr int min_blocks(int) { return 1; }
  static constexpr int block_size() { return 256; }
  static constexpr bool batched_sincos() { return true; }
  static constexpr bool springs() { return false; }
  static constexpr bool tickets() { return true; }  // 64 K environments = 1.73 waves: +12 %
  static constexpr bool contact_list(int body) { return body < 0 || body == 4 || body == 8; }  // the wheels (8 points on each rim)
#ifndef GP_NO_SIDES
  // half 0: leg 1 - foot 2 - wheel 4 and link 3; half 1: leg 5 - foot 6 - wheel 8 and link 7
  static constexpr unsigned side_mask() { return (1u << 5) | (1u << 6) | (1u << 7) | (1u << 8); }
#else
  static constexpr unsigned side_mask() { return 0u; }
#endif
};

// A specialisation for ONE more tree. Two users:
//  * gp_jit.cpp compiles it at RUN time (NVRTC) for any mechanism no shipped spec matches: the macros below are
//    the first lines of the translation unit it generates, the policies chosen from the mechanism itself
//    (which bodies carry many contact points, spring contacts, size);
//  * a library build with make variables, for deployments without NVRTC:
//      make CUSTOM_NB=4 CUSTOM_PARENTS=-1,0,1,2 CUSTOM_JOINTS=3,1,2,2 CUSTOM_AXES=0,0,0,0 CUSTOM_NAME=hopper2d_FRPP
// (0-based parents, -1 = world; joint types 0 fixed / 1 revolute / 2 prismatic / 3 floating; axes 1 = exactly +z,
// 0 = any; `python tools/custom_topo.py <model>` prints the line for a mechanism). Policy defaults are what the
// shipped specs converged to: one 256-thread block per SM beyond three bodies, ticket mode there.
#ifdef GP_CUSTOM_TOPO_NB
#ifndef GP_CUSTOM_BLOCK
#define GP_CUSTOM_BLOCK (GP_CUSTOM_TOPO_NB <= 3 ? 128 : 256)
#endif
#ifndef GP_CUSTOM_MIN_BLOCKS
#define GP_CUSTOM_MIN_BLOCKS 1
#endif
#ifndef GP_CUSTOM_SINCOS
#define GP_CUSTOM_SINCOS true
#endif
#ifndef GP_CUSTOM_SPRINGS
#define GP_CUSTOM_SPRINGS false
#endif
#ifndef GP_CUSTOM_TICKETS
#define GP_CUSTOM_TICKETS (GP_CUSTOM_TOPO_NB > 3)
#endif
#ifndef GP_CUSTOM_CONTACT_LIST_MASK
#define GP_CUSTOM_CONTACT_LIST_MASK 0u  // bit b: body b runs the per-lane list of points in contact
#endif
#ifndef GP_CUSTOM_SIDE_MASK
#define GP_CUSTOM_SIDE_MASK 0u  // bit b: body b belongs to half 1 of the warp-pair mapping (0: thread per environment)
#endif
struct SpecCustom {
  static constexpr TopoData data() {
    return {GP_CUSTOM_TOPO_NB, {GP_CUSTOM_TOPO_PARENTS}, {GP_CUSTOM_TOPO_JOINTS}, {GP_CUSTOM_TOPO_AXES}};
  }
  static const char* name() { return GP_CUSTOM_TOPO_NAME; }
  static constexpr int min_blocks(int) { return GP_CUSTOM_MIN_BLOCKS; }
  static constexpr int block_size() { return GP_CUSTOM_BLOCK; }
  static constexpr bool batched_sincos() { return GP_CUSTOM_SINCOS; }
  static constexpr bool springs() { return GP_CUSTOM_SPRINGS; }
  static constexpr bool tickets() { return GP_CUSTOM_TICKETS; }
  static constexpr bool contact_list(int body) {
    return body < 0 ? (GP_CUSTOM_CONTACT_LIST_MASK) != 0u : (((GP_CUSTOM_CONTACT_LIST_MASK) >> body) & 1u) != 0u;
  }
  static constexpr unsigned side_mask() { return GP_CUSTOM_SIDE_MASK; }
};
#endif

#undef GP_R
#undef GP_P
#undef GP_F
#undef GP_X

// does a mechanism's signature match a specialisation? (AxAny in the spec matches any axis)
constexpr bool topo_matches(const TopoData& spec, const TopoData& mech) {
  if (spec.nb != mech.nb) return false;
  for (int i = 0; i < spec.nb; ++i) {
    if (spec.parent[i] != mech.parent[i] || spec.jtype[i] != mech.jtype[i]) return false;
    if (spec.axis[i] != AxAny && spec.axis[i] != mech.axis[i]) return false;
  }
  return true;
}

}  // namespace gp
