// gp_sharded.cpp — the environments of one MechanismState batch spread over several GPUs of a box from ONE
// host process (the reference is a single process), and the optional end-of-rollout diagnostic reduction
// across processes (one rank per GPU) with NCCL loaded at run time.
//
// Environments are independent: shard g owns the contiguous range
// [g*N/G, (g+1)*N/G), keeps its state resident on its own device and has its own stream; nothing is
// exchanged on the step path. Calls that only enqueue (step) return at once, so the devices run
// concurrently; calls that move host data run one worker thread per device. Built purely on the
// single-device C ABI (gp_batch_*), host code only.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gorilla_b200.h"
#include "gp_host.h"

using gp::set_error;

struct gp_sharded {
  const gp_mechanism* mech = nullptr;
  int64_t n = 0;
  int n_q = 0, n_v = 0;
  std::vector<gp_batch*> shard;
  std::vector<int64_t> lo, hi;
};

namespace {

// run fn(g) for every shard on its own thread; first failure wins (its message is re-raised on the caller's thread)
template <class F>
int each_shard(gp_sharded* s, F fn) {
  const int G = (int)s->shard.size();
  std::vector<int> rc(G, GP_OK);
  std::vector<std::string> msg(G);
  auto body = [&](int g) {
    rc[g] = fn(g);
    if (rc[g] != GP_OK) msg[g] = gp::last_error();  // (the error string is thread-local)
  };
  if (G == 1) {
    body(0);
  } else {
    std::vector<std::thread> th;
    th.reserve(G);
    for (int g = 0; g < G; ++g) th.emplace_back(body, g);
    for (auto& t : th) t.join();
  }
  for (int g = 0; g < G; ++g)
    if (rc[g] != GP_OK) {
      set_error("shard %d (device %d): %s", g, gp_batch_device(s->shard[g]), msg[g].c_str());
      return rc[g];
    }
  return GP_OK;
}

int check(const gp_sharded* s, const char* what) {
  if (!s || s->shard.empty()) {
    set_error("%s: null sharded batch", what);
    return GP_ERR_INVALID;
  }
  return GP_OK;
}

// ---- NCCL through dlopen ----------------------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
struct NcclId {
  char internal[128];
};
struct Nccl {
  void* handle = nullptr;
  std::string why;
  int (*GetVersion)(int*) = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;      // ncclSum

const Nccl& nccl() {
  static const Nccl lib = [] {
    Nccl n;
    std::vector<std::string> candidates;
    if (const char* e = std::getenv("GP_NCCL_LIB")) candidates.push_back(e);
    // a process that already holds NCCL (e.g. through a framework) gets that copy back by soname
    candidates.push_back("libnccl.so.2");
    candidates.push_back("libnccl.so");
    for (const std::string& c : candidates) {
      n.handle = dlopen(c.c_str(), RTLD_NOW | RTLD_GLOBAL);
      if (n.handle) break;
    }
    if (!n.handle) {
      n.why = "libnccl.so.2 not found (set GP_NCCL_LIB to its path)";
      return n;
    }
    bool ok = true;
    auto sym = [&](const char* name) {
      void* p = dlsym(n.handle, name);
      if (!p) ok = false;
      return p;
    };
    n.GetVersion = (decltype(n.GetVersion))sym("ncclGetVersion");
    n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
    n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
    n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
    n.AllReduce = (decltype(n.AllReduce))sym("ncclAllReduce");
    n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
    if (!ok) {
      n.why = "the NCCL library lacks a required entry point";
      n.handle = nullptr;
    }
    return n;
  }();
  return lib;
}

}  // namespace

struct gp_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  double* buf = nullptr;  // 4 doubles on the device
};

extern "C" {

// ---- sharded batch ----------------------------------------------------------------------------------
int gp_sharded_create(const gp_mechanism* mech, int64_t n_envs, const int* device_ids, int n_devices, gp_sharded** out) {
  if (!mech || !out || !device_ids || n_devices <= 0 || n_envs < n_devices) {
    set_error("gp_sharded_create: bad argument (need a mechanism, at least one device and one environment per device)");
    return GP_ERR_INVALID;
  }
  *out = nullptr;
  gp_sharded* s = new gp_sharded();
  s->mech = mech;
  s->n = n_envs;
  s->n_q = gp_mechanism_n_q(mech);
  s->n_v = gp_mechanism_n_v(mech);
  for (int g = 0; g < n_devices; ++g) {
    const int64_t lo = n_envs * g / n_devices, hi = n_envs * (g + 1) / n_devices;
    gp_batch* b = nullptr;
    const int rc = gp_batch_create(mech, hi - lo, device_ids[g], &b);
    if (rc != GP_OK) {
      for (gp_batch* x : s->shard) gp_batch_destroy(x);
      delete s;
      return rc;
    }
    s->shard.push_back(b);
    s->lo.push_back(lo);
    s->hi.push_back(hi);
  }
  *out = s;
  return GP_OK;
}

void gp_sharded_destroy(gp_sharded* s) {
  if (!s) return;
  for (gp_batch* b : s->shard) gp_batch_destroy(b);
  delete s;
}

int gp_sharded_n_shards(const gp_sharded* s) { return s ? (int)s->shard.size() : 0; }
int64_t gp_sharded_n_envs(const gp_sharded* s) { return s ? s->n : 0; }

gp_batch* gp_sharded_shard(gp_sharded* s, int g, int64_t* env_lo, int64_t* env_hi) {
  if (!s || g < 0 || g >= (int)s->shard.size()) return nullptr;
  if (env_lo) *env_lo = s->lo[g];
  if (env_hi) *env_hi = s->hi[g];
  return s->shard[g];
}

int gp_sharded_set_state(gp_sharded* s, const double* q_host, const double* v_host) {
  int rc = check(s, "gp_sharded_set_state");
  if (rc) return rc;
  return each_shard(s, [&](int g) {
    return gp_batch_set_state(s->shard[g], q_host ? q_host + s->lo[g] * s->n_q : nullptr,
                              v_host ? v_host + s->lo[g] * s->n_v : nullptr);
  });
}

int gp_sharded_get_state(gp_sharded* s, double* q_host, double* v_host) {
  int rc = check(s, "gp_sharded_get_state");
  if (rc) return rc;
  if (!q_host || !v_host) {
    set_error("gp_sharded_get_state: null output");
    return GP_ERR_INVALID;
  }
  return each_shard(s, [&](int g) {
    return gp_batch_get_state(s->shard[g], q_host + s->lo[g] * s->n_q, v_host + s->lo[g] * s->n_v);
  });
}

int gp_sharded_set_tau(gp_sharded* s, const double* tau_host) {
  int rc = check(s, "gp_sharded_set_tau");
  if (rc) return rc;
  return each_shard(s, [&](int g) { return gp_batch_set_tau(s->shard[g], tau_host ? tau_host + s->lo[g] * s->n_v : nullptr); });
}

int gp_sharded_step(gp_sharded* s, double dt, int integrator, int n_steps, int controller, const double* ctrl_params,
                    int n_ctrl_params) {
  int rc = check(s, "gp_sharded_step");
  if (rc) return rc;
  // enqueue only: one after the other from this thread, the devices then run side by side
  for (size_t g = 0; g < s->shard.size(); ++g)
    if ((rc = gp_batch_step(s->shard[g], dt, integrator, n_steps, controller, ctrl_params, n_ctrl_params))) return rc;
  return GP_OK;
}

int gp_sharded_sync(gp_sharded* s) {
  int rc = check(s, "gp_sharded_sync");
  if (rc) return rc;
  for (gp_batch* b : s->shard)
    if ((rc = gp_batch_sync(b))) return rc;
  return GP_OK;
}

int gp_sharded_simulate(gp_sharded* s, double* q_host, double* v_host, const double* tau_host, double final_time, double dt,
                        int integrator, int controller, const double* ctrl_params, int n_ctrl_params, int64_t* n_steps_out) {
  int rc = check(s, "gp_sharded_simulate");
  if (rc) return rc;
  if (!q_host || !v_host) {
    set_error("gp_sharded_simulate: null state buffers");
    return GP_ERR_INVALID;
  }
  std::vector<int64_t> steps(s->shard.size(), 0);
  rc = each_shard(s, [&](int g) {
    return gp_batch_simulate(s->shard[g], q_host + s->lo[g] * s->n_q, v_host + s->lo[g] * s->n_v,
                             tau_host ? tau_host + s->lo[g] * s->n_v : nullptr, final_time, dt, integrator, controller,
                             ctrl_params, n_ctrl_params, &steps[g], nullptr, nullptr);
  });
  if (rc == GP_OK && n_steps_out) *n_steps_out = steps[0];
  return rc;
}

int gp_sharded_status(gp_sharded* s, uint32_t* status_host) {
  int rc = check(s, "gp_sharded_status");
  if (rc) return rc;
  if (!status_host) {
    set_error("gp_sharded_status: null output");
    return GP_ERR_INVALID;
  }
  return each_shard(s, [&](int g) { return gp_batch_status(s->shard[g], status_host + s->lo[g]); });
}

int gp_sharded_energy_sums(gp_sharded* s, double out[4]) {
  int rc = check(s, "gp_sharded_energy_sums");
  if (rc) return rc;
  if (!out) {
    set_error("gp_sharded_energy_sums: null output");
    return GP_ERR_INVALID;
  }
  const int G = (int)s->shard.size();
  std::vector<double> part((size_t)4 * G, 0.0);
  rc = each_shard(s, [&](int g) { return gp_batch_reduce_diagnostics(s->shard[g], nullptr, &part[(size_t)4 * g]); });
  if (rc) return rc;
  for (int k = 0; k < 4; ++k) {
    out[k] = 0.0;
    for (int g = 0; g < G; ++g) out[k] += part[(size_t)4 * g + k];  // fixed order: reproducible
  }
  return GP_OK;
}

// ---- diagnostic reduction across processes ---------------------------------------------------------------
int gp_nccl_available(void) { return nccl().handle != nullptr; }

int gp_comm_unique_id(char id_out[GP_COMM_ID_BYTES]) {
  const Nccl& n = nccl();
  if (!n.handle) {
    set_error("gp_comm_unique_id: %s", n.why.c_str());
    return GP_ERR_UNSUPPORTED;
  }
  if (!id_out) {
    set_error("gp_comm_unique_id: null output");
    return GP_ERR_INVALID;
  }
  NcclId id;
  const int rc = n.GetUniqueId(&id);
  if (rc != 0) {
    set_error("ncclGetUniqueId: %s", n.GetErrorString(rc));
    return GP_ERR_CUDA;
  }
  static_assert(sizeof(NcclId) == GP_COMM_ID_BYTES, "NCCL unique id is 128 bytes");
  std::memcpy(id_out, id.internal, sizeof(id.internal));
  return GP_OK;
}

int gp_comm_create(int rank, int world, const char id[GP_COMM_ID_BYTES], int device, gp_comm** out) {
  if (!out || !id || world < 1 || rank < 0 || rank >= world) {
    set_error("gp_comm_create: bad argument");
    return GP_ERR_INVALID;
  }
  *out = nullptr;
  if (gp_device_count() <= 0) {
    set_error("no CUDA device available");
    return GP_ERR_NO_DEVICE;
  }
  const Nccl& n = nccl();
  if (!n.handle) {
    set_error("gp_comm_create: %s", n.why.c_str());
    return GP_ERR_UNSUPPORTED;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    set_error("gp_comm_create: cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(cudaGetLastError()));
    return GP_ERR_CUDA;
  }
  gp_comm* c = new gp_comm();
  c->rank = rank;
  c->world = world;
  c->device = device;
  NcclId nid;
  std::memcpy(nid.internal, id, sizeof(nid.internal));
  int rc = n.CommInitRank(&c->comm, world, nid, rank);
  if (rc != 0) {
    set_error("ncclCommInitRank: %s", n.GetErrorString(rc));
    delete c;
    return GP_ERR_CUDA;
  }
  if (cudaMalloc((void**)&c->buf, 4 * sizeof(double)) != cudaSuccess) {
    set_error("gp_comm_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    n.CommDestroy(c->comm);
    delete c;
    return GP_ERR_CUDA;
  }
  *out = c;
  return GP_OK;
}

void gp_comm_destroy(gp_comm* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->buf) cudaFree(c->buf);
  if (c->comm) nccl().CommDestroy(c->comm);
  delete c;
}

int gp_batch_reduce_diagnostics(gp_batch* b, gp_comm* comm, double out_host[4]) {
  if (!b || !out_host) {
    set_error("gp_batch_reduce_diagnostics: null argument");
    return GP_ERR_INVALID;
  }
  if (comm && comm->device != gp_batch_device(b)) {
    set_error("gp_batch_reduce_diagnostics: communicator is on device %d, the batch on device %d", comm->device,
              gp_batch_device(b));
    return GP_ERR_INVALID;
  }
  if (cudaSetDevice(gp_batch_device(b)) != cudaSuccess) {
    set_error("gp_batch_reduce_diagnostics: cudaSetDevice failed: %s", cudaGetErrorString(cudaGetLastError()));
    return GP_ERR_CUDA;
  }
  double* buf = comm ? comm->buf : nullptr;
  bool own = false;
  if (!buf) {
    if (cudaMalloc((void**)&buf, 4 * sizeof(double)) != cudaSuccess) {
      set_error("gp_batch_reduce_diagnostics: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
      return GP_ERR_CUDA;
    }
    own = true;
  }
  cudaStream_t stream = (cudaStream_t)gp_batch_stream(b);
  int rc = gp_batch_energy_sums_device(b, buf);  // written on the batch's stream
  if (rc == GP_OK && comm && comm->world > 1) {
    // in place, on the same stream: ordered after the sums, no host round trip in between
    const int nrc = nccl().AllReduce(buf, buf, 4, kNcclFloat64, kNcclSum, comm->comm, stream);
    if (nrc != 0) {
      set_error("ncclAllReduce: %s", nccl().GetErrorString(nrc));
      rc = GP_ERR_CUDA;
    }
  }
  if (rc == GP_OK) {
    if (cudaMemcpyAsync(out_host, buf, 4 * sizeof(double), cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess) {
      set_error("gp_batch_reduce_diagnostics: copy failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = GP_ERR_CUDA;
    }
  }
  if (own) cudaFree(buf);
  return rc;
}

}  // extern "C"
