// gp_jit.h — run-time specialisation: a StaticTopo kernel for ANY mechanism.
//
// The reference builds any joint tree at run time (MechanismState::new, src/mechanism.rs:62-148). The
// kernels here get their speed from a compile-time topology (gp_topology.cuh StaticTopo: every body
// index a literal, all per-body arrays in registers); the shipped specs cover the trees of the
// reference's configs, this unit covers every other tree: it writes the mechanism's signature and
// policies as the macros of SpecCustom, compiles the SAME kernel sources (embedded in the library) with
// NVRTC for sm_100a, caches the cubin on disk keyed by signature + source hash, and loads it with
// cudaLibraryLoadData. One kernel per compilation, compiled when it is first launched.
//
// NVRTC is loaded with dlopen at first use; the library has no link-time dependency on it. Without it
// (or with GP_JIT=0) such mechanisms run the run-time-topology kernel (variant_generic), as before.
#pragma once
#include <string>

#include "gp_launch.h"

struct gp_mechanism;

namespace gp {

struct JitPolicy {
  int block_size = 256;      // threads per block of the step kernels (Spec::block_size)
  int min_blocks = 1;        // __launch_bounds__ minimum resident blocks (Spec::min_blocks)
  bool batched_sincos = true;
  bool springs = false;      // SpringContact legs compiled into the general-contact kernels
  bool tickets = true;       // step kernels compiled with ticket mode
  unsigned contact_list_mask = 0u;  // bit b: body b runs the per-lane list of points in contact
  unsigned side_mask = 0u;  // warp-pair mapping: bit b = body b belongs to half 1 (0: thread per environment)
};

// which kernels of a table (bit mask; gp_mechanism_precompile)
enum JitKind : unsigned { JitStepSIE = 1u, JitStepRK = 2u, JitDynamics = 4u, JitEnergy = 8u, JitStepTauSeq = 16u };

// is NVRTC loadable in this process? (`why` receives the reason when not)
bool jit_available(std::string* why = nullptr);
// policies from the mechanism itself: size, contact points per body, spring contacts
JitPolicy jit_policy_for(const gp_mechanism* m, const TopoData& td);
// The table of (signature, policy), interned for the life of the process. Creating it compiles nothing.
const KernelTable* jit_table(const TopoData& td, const JitPolicy& pol);
bool is_jit_table(const KernelTable* t);
// Compile kernels of a JIT table into the on-disk cache without loading them (needs no GPU).
// Returns a gp_status_code; n_compiled counts the compilations that were not already cached.
int jit_precompile(const KernelTable* t, int contact, unsigned kinds, int* n_compiled);
// directory cubins are written to (first writable of $GP_JIT_CACHE, <library dir>/jit_cache,
// ~/.cache/gorilla_b200, /tmp/gorilla_b200_jit)
std::string jit_cache_dir();

}  // namespace gp
