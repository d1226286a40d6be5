/* Plain C against include/gorilla_b200.h: the several-GPUs-from-one-process contract (gp_sharded_*) and the
 * per-step torque sequence, as a C host (or a Rust `extern "C"` block) would drive them.
 *   usage: test_sharded [device ...]       (default: devices 0 0 0 - three shards on one GPU: the logic)
 * Built by __graft_entry__.build() with gcc (no C++), run on the GPU by tests/test_cpp_facade_gpu.py. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gorilla_b200.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc__ = (call);                                                       \
    if (rc__ != GP_OK) {                                                     \
      char msg[512];                                                         \
      gp_last_error(msg, sizeof msg);                                        \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc__, msg);                   \
      return 1;                                                              \
    }                                                                        \
  } while (0)

static double frand(unsigned long long* s) {
  *s = *s * 6364136223846793005ULL + 1442695040888963407ULL;
  return (double)(*s >> 11) / 9007199254740992.0 * 2.0 - 1.0;
}

int main(int argc, char** argv) {
  int devices[8] = {0, 0, 0}, n_dev = 3;
  if (argc > 1) {
    n_dev = argc - 1 > 8 ? 8 : argc - 1;
    for (int i = 0; i < n_dev; ++i) devices[i] = atoi(argv[i + 1]);
  }
  gp_mechanism* mech = NULL;
  CHECK(gp_model_create("so101", NULL, 0, &mech));
  const double origin[3] = {0, 0, 0}, up[3] = {0, 0, 1};
  for (int body = 3; body <= 7; ++body) CHECK(gp_mechanism_add_contact_point(mech, body, origin, 50e3));
  CHECK(gp_mechanism_add_halfspace(mech, origin, up, 0.9, 0.5));
  const int nq = gp_mechanism_n_q(mech), nv = gp_mechanism_n_v(mech);
  const long long n = 1001;
  const int steps = 40;
  double *q = malloc(sizeof(double) * n * nq), *v = malloc(sizeof(double) * n * nv), *tau = malloc(sizeof(double) * n * nv);
  double *q1 = malloc(sizeof(double) * n * nq), *v1 = malloc(sizeof(double) * n * nv);
  double *q2 = malloc(sizeof(double) * n * nq), *v2 = malloc(sizeof(double) * n * nv);
  unsigned long long seed = 12345;
  for (long long i = 0; i < n * nq; ++i) q[i] = frand(&seed);
  for (long long i = 0; i < n * nv; ++i) v[i] = frand(&seed), tau[i] = 0.2 * frand(&seed);

  /* one batch on one device */
  gp_batch* one = NULL;
  CHECK(gp_batch_create(mech, n, devices[0], &one));
  CHECK(gp_batch_set_state(one, q, v));
  CHECK(gp_batch_set_tau(one, tau));
  CHECK(gp_batch_step(one, 1.0 / 6000.0, GP_SEMI_IMPLICIT_EULER, steps, GP_CTRL_NONE, NULL, 0));
  CHECK(gp_batch_get_state(one, q1, v1));
  double sums1[4];
  CHECK(gp_batch_reduce_diagnostics(one, NULL, sums1));

  /* the same environments over n_dev shards */
  gp_sharded* sh = NULL;
  CHECK(gp_sharded_create(mech, n, devices, n_dev, &sh));
  if (gp_sharded_n_shards(sh) != n_dev || gp_sharded_n_envs(sh) != n) return 2;
  long long covered = 0;
  for (int g = 0; g < n_dev; ++g) {
    int64_t lo, hi;
    gp_batch* b = gp_sharded_shard(sh, g, &lo, &hi);
    if (!b || lo != covered || gp_batch_n_envs(b) != hi - lo || gp_batch_device(b) != devices[g]) return 3;
    covered = hi;
  }
  if (covered != n) return 4;
  CHECK(gp_sharded_set_state(sh, q, v));
  CHECK(gp_sharded_set_tau(sh, tau));
  CHECK(gp_sharded_step(sh, 1.0 / 6000.0, GP_SEMI_IMPLICIT_EULER, steps, GP_CTRL_NONE, NULL, 0));
  CHECK(gp_sharded_sync(sh));
  CHECK(gp_sharded_get_state(sh, q2, v2));
  if (memcmp(q1, q2, sizeof(double) * n * nq) || memcmp(v1, v2, sizeof(double) * n * nv)) {
    fprintf(stderr, "sharded step differs from the single batch\n");
    return 5;
  }
  double sums2[4];
  CHECK(gp_sharded_energy_sums(sh, sums2));
  for (int k = 0; k < 4; ++k)
    if (fabs(sums1[k] - sums2[k]) > 1e-9 * fmax(1.0, fabs(sums1[k]))) {
      fprintf(stderr, "diagnostic sum %d: %.17g vs %.17g\n", k, sums1[k], sums2[k]);
      return 6;
    }
  /* simulate() through host buffers on every shard at once */
  memcpy(q2, q, sizeof(double) * n * nq);
  memcpy(v2, v, sizeof(double) * n * nv);
  int64_t done = 0;
  CHECK(gp_sharded_simulate(sh, q2, v2, tau, (steps - 0.5) / 6000.0, 1.0 / 6000.0, GP_SEMI_IMPLICIT_EULER, GP_CTRL_NONE, NULL, 0, &done));
  if (done != steps || memcmp(q1, q2, sizeof(double) * n * nq) || memcmp(v1, v2, sizeof(double) * n * nv)) {
    fprintf(stderr, "sharded simulate differs (%lld steps)\n", (long long)done);
    return 7;
  }
  uint32_t* status = calloc(n, sizeof(uint32_t));
  CHECK(gp_sharded_status(sh, status));
  for (long long e = 0; e < n; ++e)
    if (status[e]) return 8;

  /* a torque per time step (the reference's control closure): sequence in one call == set_tau + step per time step */
  const int K = 12;
  double* seq = malloc(sizeof(double) * K * n * nv);
  for (long long i = 0; i < (long long)K * n * nv; ++i) seq[i] = 0.3 * frand(&seed);
  CHECK(gp_batch_set_state(one, q, v));
  for (int s = 0; s < K; ++s) {
    CHECK(gp_batch_set_tau(one, seq + (size_t)s * n * nv));
    CHECK(gp_batch_step(one, 1.0 / 6000.0, GP_SEMI_IMPLICIT_EULER, 1, GP_CTRL_NONE, NULL, 0));
  }
  CHECK(gp_batch_get_state(one, q1, v1));
  CHECK(gp_batch_set_state(one, q, v));
  CHECK(gp_batch_step_tau_sequence(one, 1.0 / 6000.0, GP_SEMI_IMPLICIT_EULER, K, seq));
  CHECK(gp_batch_get_state(one, q2, v2));
  if (memcmp(q1, q2, sizeof(double) * n * nq) || memcmp(v1, v2, sizeof(double) * n * nv)) {
    fprintf(stderr, "torque sequence differs from set_tau + step per time step\n");
    return 9;
  }
  gp_sharded_destroy(sh);
  gp_batch_destroy(one);
  gp_mechanism_destroy(mech);
  printf("sharded-ok %d shards, %lld environments, sum KE %.12g\n", n_dev, n, sums2[0]);
  return 0;
}
