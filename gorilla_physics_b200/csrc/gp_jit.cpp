// gp_jit.cpp — run-time specialisation of the kernels for a mechanism's own topology (see gp_jit.h).
//
// Host only. NVRTC (dlopen) turns "SpecCustom macros + #include gp_kernels.cuh" into a cubin for sm_100a;
// the cubin is cached on disk and loaded with cudaLibraryLoadData / cudaLibraryGetKernel (context
// independent: one load serves every device of the process). Compiling needs no GPU, loading does.
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/file.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "gp_host.h"
#include "gp_jit.h"
#include "gp_topology.cuh"

namespace gp {
namespace {

// the kernel sources, embedded when the library is built (csrc/Makefile -> tools/embed_sources.py):
//   kJitSourceNames[i] / kJitSourceTexts[i], kJitSourceCount
#include "gp_jit_sources.inc"

// standard headers the sources name but device code does not need (NVRTC has the CUDA math library,
// size_t and the vector types built in)
const char* const kStubNames[] = {"cstdint", "stdint.h", "stddef.h", "cmath", "cstdlib", "cuda_runtime.h"};
const char* const kStdintStub =
    "typedef signed char int8_t; typedef unsigned char uint8_t; typedef short int16_t; typedef unsigned short uint16_t;\n"
    "typedef int int32_t; typedef unsigned int uint32_t; typedef long long int64_t; typedef unsigned long long uint64_t;\n";
const char* const kStubTexts[] = {"#include <stdint.h>\n", kStdintStub, "\n", "\n", "\n", "\n"};

const char* const kArchOption = "--gpu-architecture=sm_100a";

// ---- NVRTC through dlopen ----------------------------------------------------------------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* handle = nullptr;
  std::string path, why;
  int major = 0, minor = 0;
  int (*Version)(int*, int*) = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*AddNameExpression)(nvrtcProgram, const char*) = nullptr;
  int (*GetLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};

const Nvrtc& nvrtc() {
  static const Nvrtc lib = [] {
    Nvrtc n;
    std::vector<std::string> candidates;
    if (const char* e = std::getenv("GP_NVRTC_LIB")) candidates.push_back(e);
    // The toolkit's own copy first, by full path: a bare soname resolves to whatever copy the process already holds
    // (a Python process that imported torch holds torch's bundled NVRTC, an older release), and the compiler
    // version is part of the cache key - kernels compiled ahead of time (gp_mechanism_precompile) would never be
    // found again.
    for (const char* var : {"CUDA_HOME", "CUDA_PATH"})
      if (const char* e = std::getenv(var)) {
        candidates.push_back(std::string(e) + "/lib64/libnvrtc.so.12");
        candidates.push_back(std::string(e) + "/lib64/libnvrtc.so");
      }
    candidates.push_back("/usr/local/cuda/lib64/libnvrtc.so.12");
    candidates.push_back("/usr/local/cuda/lib64/libnvrtc.so");
    candidates.push_back("libnvrtc.so.12");
    candidates.push_back("libnvrtc.so");
    for (const std::string& c : candidates) {
      n.handle = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL);
      if (n.handle) {
        n.path = c;
        break;
      }
    }
    if (!n.handle) {
      n.why = "libnvrtc.so not found (set GP_NVRTC_LIB to its path)";
      return n;
    }
    bool ok = true;
    auto sym = [&](const char* name) {
      void* p = dlsym(n.handle, name);
      if (!p) ok = false;
      return p;
    };
    n.Version = (decltype(n.Version))sym("nvrtcVersion");
    n.CreateProgram = (decltype(n.CreateProgram))sym("nvrtcCreateProgram");
    n.DestroyProgram = (decltype(n.DestroyProgram))sym("nvrtcDestroyProgram");
    n.CompileProgram = (decltype(n.CompileProgram))sym("nvrtcCompileProgram");
    n.GetProgramLogSize = (decltype(n.GetProgramLogSize))sym("nvrtcGetProgramLogSize");
    n.GetProgramLog = (decltype(n.GetProgramLog))sym("nvrtcGetProgramLog");
    n.GetCUBINSize = (decltype(n.GetCUBINSize))sym("nvrtcGetCUBINSize");
    n.GetCUBIN = (decltype(n.GetCUBIN))sym("nvrtcGetCUBIN");
    n.AddNameExpression = (decltype(n.AddNameExpression))sym("nvrtcAddNameExpression");
    n.GetLoweredName = (decltype(n.GetLoweredName))sym("nvrtcGetLoweredName");
    n.GetErrorString = (decltype(n.GetErrorString))sym("nvrtcGetErrorString");
    if (!ok || n.Version(&n.major, &n.minor) != 0) {
      n.why = "NVRTC at " + n.path + " lacks a required entry point";
      dlclose(n.handle);
      n.handle = nullptr;
      return n;
    }
    if (n.major < 12 || (n.major == 12 && n.minor < 8)) {  // sm_100a needs 12.8
      char buf[160];
      snprintf(buf, sizeof(buf), "NVRTC %d.%d at %s cannot target sm_100a (needs 12.8 or later)", n.major, n.minor, n.path.c_str());
      n.why = buf;
      dlclose(n.handle);
      n.handle = nullptr;
    }
    return n;
  }();
  return lib;
}

// ---- cache --------------------------------------------------------------------------------------
uint64_t fnv1a(const std::string& s, uint64_t h) {
  for (unsigned char c : s) {
    h ^= c;
    h *= 0x100000001b3ull;
  }
  return h;
}

const std::string& sources_digest() {
  static const std::string d = [] {
    uint64_t a = 0xcbf29ce484222325ull, b = 0x84222325cbf29ce4ull;
    for (int i = 0; i < kJitSourceCount; ++i) {
      a = fnv1a(kJitSourceNames[i], a);
      a = fnv1a(kJitSourceTexts[i], a);
      b = fnv1a(kJitSourceTexts[i], b);
    }
    char buf[40];
    snprintf(buf, sizeof(buf), "%016llx%016llx", (unsigned long long)a, (unsigned long long)b);
    return std::string(buf);
  }();
  return d;
}

std::string library_dir() {
  Dl_info info;
  if (dladdr((const void*)&sources_digest, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    const size_t k = p.rfind('/');
    if (k != std::string::npos) return p.substr(0, k);
  }
  return ".";
}

bool dir_usable(const std::string& d, bool create) {
  struct stat st;
  if (stat(d.c_str(), &st) != 0) {
    if (!create) return false;
    // (parents first)
    for (size_t k = 1; k <= d.size(); ++k)
      if (k == d.size() || d[k] == '/') mkdir(d.substr(0, k).c_str(), 0777);
    if (stat(d.c_str(), &st) != 0) return false;
  }
  return S_ISDIR(st.st_mode) && (!create || access(d.c_str(), W_OK | X_OK) == 0);
}

std::vector<std::string> cache_dirs() {
  std::vector<std::string> v;
  if (const char* e = std::getenv("GP_JIT_CACHE")) v.push_back(e);
  v.push_back(library_dir() + "/jit_cache");
  if (const char* e = std::getenv("XDG_CACHE_HOME")) v.push_back(std::string(e) + "/gorilla_b200");
  if (const char* e = std::getenv("HOME")) v.push_back(std::string(e) + "/.cache/gorilla_b200");
  v.push_back("/tmp/gorilla_b200_jit");
  return v;
}

// cache entry: "GPJIT1 <lowered name>\n" followed by the cubin
bool read_entry(const std::string& path, std::string* lowered, std::vector<char>* cubin) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<char> all;
  char buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) all.insert(all.end(), buf, buf + n);
  fclose(f);
  const char magic[] = "GPJIT1 ";
  if (all.size() < sizeof(magic) || std::memcmp(all.data(), magic, sizeof(magic) - 1) != 0) return false;
  size_t nl = sizeof(magic) - 1;
  while (nl < all.size() && all[nl] != '\n') ++nl;
  if (nl >= all.size()) return false;
  lowered->assign(all.data() + sizeof(magic) - 1, all.data() + nl);
  cubin->assign(all.begin() + nl + 1, all.end());
  return !cubin->empty();
}

bool write_entry(const std::string& dir, const std::string& file, const std::string& lowered, const std::vector<char>& cubin) {
  char tmp[64];
  snprintf(tmp, sizeof(tmp), "/.tmp.%d.%p", (int)getpid(), (const void*)&cubin);
  const std::string t = dir + tmp;
  FILE* f = fopen(t.c_str(), "wb");
  if (!f) return false;
  const std::string head = "GPJIT1 " + lowered + "\n";
  bool ok = fwrite(head.data(), 1, head.size(), f) == head.size() && fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
  ok = (fclose(f) == 0) && ok;
  if (ok) ok = rename(t.c_str(), (dir + "/" + file).c_str()) == 0;  // atomic: readers never see a partial entry
  if (!ok) unlink(t.c_str());
  return ok;
}

// ---- one kernel ---------------------------------------------------------------------------------
struct Slot {
  std::atomic<cudaKernel_t> kernel{nullptr};
  cudaLibrary_t library = nullptr;
  bool failed = false;
  std::string error;
  std::atomic<bool> smem_set[64] = {};  // per device: dynamic shared memory limit raised (warp-pair kernels)
};

struct JitTable : KernelTable {
  std::string defines;  // the SpecCustom macros: first lines of every translation unit of this table
  std::string label;    // storage behind KernelTable::name
  std::mutex mu;
  Slot step_slot[3][2][3], dynamics_slot, energy_slot;  // [contact][integrator class][plain / warp pairs / torque sequence]
};

// flavour of a semi-implicit-Euler step kernel: 0 plain, 1 warp pairs, 2 reads a torque sequence (the Runge-Kutta
// kernels have one flavour, which reads it)
std::string step_expr(int contact, int integ, int flavour) {
  char buf[160];
  snprintf(buf, sizeof(buf), "&gp::step_kernel<gp::StaticTopo<gp::SpecCustom>, %d, %d, %s, %s>", contact, integ,
           flavour == 1 ? "true" : "false", (integ != 0 || flavour == 2) ? "true" : "false");
  return buf;
}
const char* const kDynamicsExpr = "&gp::dynamics_kernel<gp::StaticTopo<gp::SpecCustom>, 2>";
const char* const kEnergyExpr = "&gp::energy_kernel<gp::StaticTopo<gp::SpecCustom>>";

std::string entry_file(const JitTable* t, const std::string& expr) {
  const Nvrtc& n = nvrtc();
  char ver[32];
  snprintf(ver, sizeof(ver), "nvrtc%d.%d", n.major, n.minor);
  const std::string key = t->defines + "\n" + expr + "\n" + kArchOption + "\n" + ver + "\n" + sources_digest();
  char buf[48];
  snprintf(buf, sizeof(buf), "%016llx%016llx.cubin", (unsigned long long)fnv1a(key, 0xcbf29ce484222325ull),
           (unsigned long long)fnv1a(key, 0x9ae16a3b2f90404full));
  return buf;
}

// compile `expr` of table t with NVRTC; on success fills lowered + cubin
int compile(const JitTable* t, const std::string& expr, std::string* lowered, std::vector<char>* cubin) {
  const Nvrtc& n = nvrtc();
  if (!n.handle) {
    set_error("run-time specialisation needs NVRTC: %s", n.why.c_str());
    return GP_ERR_JIT;
  }
  const std::string src = t->defines + "#include \"gp_kernels.cuh\"\n";
  std::vector<const char*> names, texts;
  for (int i = 0; i < kJitSourceCount; ++i) {
    names.push_back(kJitSourceNames[i]);
    texts.push_back(kJitSourceTexts[i]);
  }
  for (size_t i = 0; i < sizeof(kStubNames) / sizeof(kStubNames[0]); ++i) {
    names.push_back(kStubNames[i]);
    texts.push_back(kStubTexts[i]);
  }
  nvrtcProgram prog = nullptr;
  int rc = n.CreateProgram(&prog, src.c_str(), "gp_jit_unit.cu", (int)names.size(), texts.data(), names.data());
  if (rc != 0) {
    set_error("nvrtcCreateProgram: %s", n.GetErrorString(rc));
    return GP_ERR_JIT;
  }
  auto fail = [&](const char* what, int code) {
    size_t ls = 0;
    std::string log;
    if (n.GetProgramLogSize(prog, &ls) == 0 && ls > 1) {
      log.resize(ls);
      n.GetProgramLog(prog, &log[0]);
      if (log.size() > 700) log.resize(700);
    }
    set_error("%s: %s\n%s", what, n.GetErrorString(code), log.c_str());
    n.DestroyProgram(&prog);
    return GP_ERR_JIT;
  };
  if ((rc = n.AddNameExpression(prog, expr.c_str())) != 0) return fail("nvrtcAddNameExpression", rc);
  // -default-device: the unannotated constexpr helpers (make_tables, select_chain's functors) and the C ABI's
  // declarations are host functions to nvcc's relaxed-constexpr mode; here they are device functions
  const char* opts[] = {kArchOption, "-std=c++17", "-default-device", "-lineinfo"};
  if ((rc = n.CompileProgram(prog, (int)(sizeof(opts) / sizeof(opts[0])), opts)) != 0) return fail("nvrtcCompileProgram", rc);
  const char* low = nullptr;
  if ((rc = n.GetLoweredName(prog, expr.c_str(), &low)) != 0 || !low) return fail("nvrtcGetLoweredName", rc);
  *lowered = low;
  size_t cs = 0;
  if ((rc = n.GetCUBINSize(prog, &cs)) != 0 || cs == 0) return fail("nvrtcGetCUBINSize", rc);
  cubin->resize(cs);
  if ((rc = n.GetCUBIN(prog, cubin->data())) != 0) return fail("nvrtcGetCUBIN", rc);
  n.DestroyProgram(&prog);
  return GP_OK;
}

// cubin of `expr`: from the cache, else compiled (and cached). compiled_now reports which.
int obtain(const JitTable* t, const std::string& expr, std::string* lowered, std::vector<char>* cubin, bool* compiled_now) {
  const std::string file = entry_file(t, expr);
  const std::vector<std::string> dirs = cache_dirs();
  if (compiled_now) *compiled_now = false;
  static const bool no_cache = std::getenv("GP_JIT_NO_CACHE") != nullptr;  // tests
  auto lookup = [&] {
    if (no_cache) return false;
    for (const std::string& d : dirs)
      if (read_entry(d + "/" + file, lowered, cubin)) return true;
    return false;
  };
  if (lookup()) return GP_OK;
  // one compiler per entry across processes (the 8 ranks of a box want the same kernel at the same time)
  std::string wdir;
  for (const std::string& d : dirs)
    if (dir_usable(d, true)) {
      wdir = d;
      break;
    }
  int lock_fd = -1;
  if (!wdir.empty()) {
    lock_fd = open((wdir + "/" + file + ".lock").c_str(), O_CREAT | O_RDWR, 0666);
    if (lock_fd >= 0) flock(lock_fd, LOCK_EX);
  }
  int rc = GP_OK;
  if (!lookup()) {
    rc = compile(t, expr, lowered, cubin);
    if (rc == GP_OK) {
      if (compiled_now) *compiled_now = true;
      if (!wdir.empty() && !no_cache) write_entry(wdir, file, *lowered, *cubin);
    }
  }
  if (lock_fd >= 0) {
    flock(lock_fd, LOCK_UN);
    close(lock_fd);
    unlink((wdir + "/" + file + ".lock").c_str());
  }
  return rc;
}

// kernel handle of a slot, loading (and compiling) on first use
int get_kernel(JitTable* t, Slot& slot, const std::string& expr, cudaKernel_t* out) {
  cudaKernel_t k = slot.kernel.load(std::memory_order_acquire);
  if (k) {
    *out = k;
    return GP_OK;
  }
  std::lock_guard<std::mutex> lock(t->mu);
  k = slot.kernel.load(std::memory_order_acquire);
  if (k) {
    *out = k;
    return GP_OK;
  }
  if (slot.failed) {  // do not recompile a broken unit on every launch
    set_error("%s", slot.error.c_str());
    return GP_ERR_JIT;
  }
  std::string lowered;
  std::vector<char> cubin;
  int rc = obtain(t, expr, &lowered, &cubin, nullptr);
  if (rc == GP_OK) {
    cudaError_t e = cudaLibraryLoadData(&slot.library, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&k, slot.library, lowered.c_str());
    if (e != cudaSuccess) {
      set_error("loading the run-time-compiled kernel %s failed: %s", expr.c_str(), cudaGetErrorString(e));
      cudaGetLastError();
      rc = GP_ERR_JIT;
    }
  }
  if (rc != GP_OK) {
    slot.failed = true;
    slot.error = last_error();
    return rc;
  }
  slot.kernel.store(k, std::memory_order_release);
  *out = k;
  return GP_OK;
}

// the C ABI reports JIT failures through gp_last_error; the launch interface speaks cudaError_t
constexpr cudaError_t kJitFailed = cudaErrorJitCompilationDisabled;

cudaError_t jit_step(const KernelTable* self, int contact, int integ_class, cudaStream_t s, const MechParams& P, const StepArgs& A0) {
  JitTable* t = const_cast<JitTable*>(static_cast<const JitTable*>(self));
  const int c = contact < 0 ? 0 : (contact > 2 ? 2 : contact), ic = integ_class == IntegSIE ? 0 : 1;
  cudaKernel_t k = nullptr;
  const bool tauseq = ic == 0 && A0.tau_seq != nullptr;
  const bool pairs = ic == 0 && !tauseq && t->lanes_sie == 2 && use_pairs(A0.n, t->block_size);
  const int flavour = ic != 0 ? 0 : (tauseq ? 2 : (pairs ? 1 : 0));
  if (get_kernel(t, t->step_slot[c][ic][flavour], step_expr(c, ic, flavour), &k) != GP_OK) return kJitFailed;
  StepArgs A = A0;
  const int lanes = pairs ? 2 : 1;
  if (lanes == 2) {
    // exchange buffers of the warp pairs: more than the 48 KB a kernel gets without asking (per kernel and device)
    int dev = 0;
    cudaGetDevice(&dev);
    Slot& slot = t->step_slot[c][ic][1];
    if (dev >= 0 && dev < 64 && !slot.smem_set[dev].load(std::memory_order_acquire)) {
      cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_dynamic_smem(t->block_size, 2));
      slot.smem_set[dev].store(true, std::memory_order_release);
    }
  }
  const StepLaunchPlan plan = plan_step_launch((const void*)k, t->block_size, t->tickets, s, A, lanes);
  void* args[] = {(void*)&P, (void*)&A};
  return cudaLaunchKernel((const void*)k, dim3(plan.grid), dim3((unsigned)plan.block), args, plan.smem, s);
}

cudaError_t jit_dynamics(const KernelTable* self, int /*contact*/, cudaStream_t s, const MechParams& P, const DynArgs& A) {
  JitTable* t = const_cast<JitTable*>(static_cast<const JitTable*>(self));
  cudaKernel_t k = nullptr;
  if (get_kernel(t, t->dynamics_slot, kDynamicsExpr, &k) != GP_OK) return kJitFailed;
  void* args[] = {(void*)&P, (void*)&A};
  return cudaLaunchKernel((const void*)k, dim3(grid_for(A.n)), dim3(kBlock), args, 0, s);
}

cudaError_t jit_energy(const KernelTable* self, cudaStream_t s, const MechParams& P, const EnergyArgs& A) {
  JitTable* t = const_cast<JitTable*>(static_cast<const JitTable*>(self));
  cudaKernel_t k = nullptr;
  if (get_kernel(t, t->energy_slot, kEnergyExpr, &k) != GP_OK) return kJitFailed;
  void* args[] = {(void*)&P, (void*)&A};
  return cudaLaunchKernel((const void*)k, dim3(grid_for(A.n)), dim3(kBlock), args, 0, s);
}

std::string defines_of(const TopoData& td, const JitPolicy& pol) {
  std::string parents, joints, axes;
  char num[16];
  for (int i = 0; i < td.nb; ++i) {
    const char* sep = i ? "," : "";
    snprintf(num, sizeof(num), "%s%d", sep, td.parent[i]);
    parents += num;
    snprintf(num, sizeof(num), "%s%d", sep, td.jtype[i]);
    joints += num;
    snprintf(num, sizeof(num), "%s%d", sep, td.axis[i]);
    axes += num;
  }
  char buf[1024];
  snprintf(buf, sizeof(buf),
           "#define GP_CUSTOM_TOPO_NB %d\n#define GP_CUSTOM_TOPO_PARENTS %s\n#define GP_CUSTOM_TOPO_JOINTS %s\n"
           "#define GP_CUSTOM_TOPO_AXES %s\n#define GP_CUSTOM_TOPO_NAME \"jit\"\n#define GP_CUSTOM_BLOCK %d\n"
           "#define GP_CUSTOM_MIN_BLOCKS %d\n#define GP_CUSTOM_SINCOS %s\n#define GP_CUSTOM_SPRINGS %s\n"
           "#define GP_CUSTOM_TICKETS %s\n#define GP_CUSTOM_CONTACT_LIST_MASK 0x%xu\n#define GP_CUSTOM_SIDE_MASK 0x%xu\n",
           td.nb, parents.c_str(), joints.c_str(), axes.c_str(), pol.block_size, pol.min_blocks,
           pol.batched_sincos ? "true" : "false", pol.springs ? "true" : "false", pol.tickets ? "true" : "false",
           pol.contact_list_mask, pol.side_mask);
  return buf;
}

}  // namespace

bool jit_available(std::string* why) {
  static const bool disabled = [] {
    const char* e = std::getenv("GP_JIT");
    return e && (e[0] == '0' || e[0] == 'n' || e[0] == 'N' || e[0] == 'f' || e[0] == 'F');
  }();
  if (disabled) {
    if (why) *why = "disabled by GP_JIT=0";
    return false;
  }
  const Nvrtc& n = nvrtc();
  if (!n.handle && why) *why = n.why;
  return n.handle != nullptr;
}

JitPolicy jit_policy_for(const gp_mechanism* m, const TopoData& td) {
  // what the shipped specs converged to (gp_topology.cuh, profiles/r1_tuning.md)
  JitPolicy p;
  p.block_size = td.nb <= 3 ? 128 : 256;
  // a double pendulum (with or without a fixed body) needs few registers: asking for six
  // 128-thread blocks per SM is what the shipped double-pendulum spec runs with (gp_topology.cuh: +3.6 %, and 1.4x
  // against the one-block default on its run-time-compiled twin, profiles/r2_static_generic_jit.jsonl)
  int dofs = 0, revolute = 0;
  for (int i = 0; i < td.nb; ++i) {
    dofs += joint_nv(td.jtype[i]);
    revolute += td.jtype[i] == JRevolute;
  }
  p.min_blocks = (td.nb <= 3 && dofs == 2 && revolute == 2 && m->n_sc() == 0) ? 6 : 1;
  p.tickets = td.nb > 3;
  p.springs = m->n_sc() > 0;
  // the per-lane hit list pays on bodies with many points, of which a few touch at a time (wheels, spokes, corners)
  std::vector<int> per_body(td.nb, 0);
  for (int b : m->cp_body)
    if (b >= 1 && b <= td.nb) per_body[b - 1]++;
  for (int i = 0; i < td.nb; ++i)
    if (per_body[i] >= 4) p.contact_list_mask |= 1u << i;
  // eight general-axis angles in flight at the start of the step cost a big tree more in spills than the
  // shared literals of the lockstep sin/cos save (quadruped); +z joints need fewer live values (navbot)
  int general_revolute = 0;
  for (int i = 0; i < td.nb; ++i)
    if (td.jtype[i] == JRevolute && td.axis[i] == AxAny) general_revolute++;
  p.batched_sincos = general_revolute < 8;
  // Warp pairs (gp_topology.cuh): big trees whose root (body 0, the only child of the world) carries at least two
  // child subtrees are cut at the root into two halves of about equal size, one warp each: half the live state
  // per thread where a thread per environment spills (the 9-body trees). Whole subtrees go to a half, largest
  // first onto the lighter half.
  static const bool no_sides = std::getenv("GP_JIT_NO_SIDES") != nullptr;  // tuning only
  int nv_total = 0, roots = 0;
  for (int i = 0; i < td.nb; ++i) {
    nv_total += joint_nv(td.jtype[i]);
    if (td.parent[i] < 0) roots++;
  }
  if (!no_sides && !p.springs && roots == 1 && td.parent[0] < 0 && td.nb >= 7 && nv_total >= 12) {
    std::vector<int> top(td.nb, -1), weight(td.nb, 0);  // child subtree of the root a body belongs to, dofs per subtree
    for (int i = 1; i < td.nb; ++i) {
      top[i] = td.parent[i] == 0 ? i : top[td.parent[i]];
      weight[top[i]] += 1 + joint_nv(td.jtype[i]);
    }
    std::vector<int> subtrees;
    for (int i = 1; i < td.nb; ++i)
      if (top[i] == i) subtrees.push_back(i);
    if (subtrees.size() >= 2) {
      std::sort(subtrees.begin(), subtrees.end(), [&](int a, int b) { return weight[a] != weight[b] ? weight[a] > weight[b] : a < b; });
      int load[2] = {0, 0};
      unsigned mask = 0u;
      for (int sroot : subtrees) {
        const int half = load[1] < load[0] ? 1 : 0;
        load[half] += weight[sroot];
        if (half == 1)
          for (int i = 1; i < td.nb; ++i)
            if (top[i] == sroot) mask |= 1u << i;
      }
      // worth it only when the halves are balanced: the heavier one sets the pace
      if (mask != 0u && 3 * std::min(load[0], load[1]) >= 2 * std::max(load[0], load[1])) p.side_mask = mask;
    }
  }
  return p;
}

const KernelTable* jit_table(const TopoData& td, const JitPolicy& pol) {
  static std::mutex mu;
  static std::map<std::string, std::unique_ptr<JitTable>> tables;
  const std::string defs = defines_of(td, pol);
  std::lock_guard<std::mutex> lock(mu);
  auto it = tables.find(defs);
  if (it != tables.end()) return it->second.get();
  std::unique_ptr<JitTable> t(new JitTable());
  t->defines = defs;
  // "jit:<joint letters>" e.g. jit:XRRRRRRX for the SO-101 with a fixed leaf
  t->label = "jit:";
  for (int i = 0; i < td.nb; ++i) t->label += "XRPF"[td.jtype[i] & 3];
  t->name = t->label.c_str();
  t->topo = td;
  t->is_static = true;
  t->block_size = pol.block_size;
  t->springs = pol.springs;
  t->tickets = pol.tickets;
  t->lanes_sie = pol.side_mask != 0u ? 2 : 1;
  t->step = &jit_step;
  t->dynamics = &jit_dynamics;
  t->energy = &jit_energy;
  const KernelTable* out = t.get();
  tables.emplace(defs, std::move(t));
  return out;
}

bool is_jit_table(const KernelTable* t) { return t && t->step == &jit_step; }

int jit_precompile(const KernelTable* table, int contact, unsigned kinds, int* n_compiled) {
  if (n_compiled) *n_compiled = 0;
  if (!is_jit_table(table)) return GP_OK;  // build-time variants have nothing to compile
  const JitTable* t = static_cast<const JitTable*>(table);
  const int c = contact < 0 ? 0 : (contact > 2 ? 2 : contact);
  std::vector<std::string> exprs;
  if (kinds & JitStepSIE) {
    exprs.push_back(step_expr(c, 0, 0));
    if (t->lanes_sie == 2) exprs.push_back(step_expr(c, 0, 1));  // small batches run warp pairs
  }
  if (kinds & JitStepRK) exprs.push_back(step_expr(c, 1, 0));
  if (kinds & JitStepTauSeq) exprs.push_back(step_expr(c, 0, 2));
  if (kinds & JitDynamics) exprs.push_back(kDynamicsExpr);
  if (kinds & JitEnergy) exprs.push_back(kEnergyExpr);
  for (const std::string& e : exprs) {
    std::string lowered;
    std::vector<char> cubin;
    bool now = false;
    const int rc = obtain(t, e, &lowered, &cubin, &now);
    if (rc != GP_OK) return rc;
    if (now && n_compiled) ++*n_compiled;
  }
  return GP_OK;
}

std::string jit_cache_dir() {
  for (const std::string& d : cache_dirs())
    if (dir_usable(d, true)) return d;
  return "";
}

}  // namespace gp
