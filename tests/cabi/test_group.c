/* Drives the C-ABI of include/tnco_b200.h from plain C -- no Python anywhere: a device group of two engines
 * (two GPUs when the box has them, else two engines on GPU 0; or the emulation build on a CPU box) anneals 64 chains
 * of a 3-regular 48-tensor network under a wall-clock budget, and the result must equal, chain by chain and bit for
 * bit, what ONE engine produces for the same seeds (results do not depend on the number of devices), the winner's
 * tree must be a valid contraction tree whose cost re-evaluates to the reported minimum.
 *
 *   gcc -O1 -I include tests/cabi/test_group.c -o /tmp/test_group -L tnco_b200 -ltnco_b200 -Wl,-rpath,$PWD/tnco_b200
 *   /tmp/test_group [device0 device1]
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "tnco_b200.h"

#define N 48
#define NI (3 * N / 2)
#define W ((NI + 31) / 32)
#define CHAINS 64
#define SWEEPS 300

#define CHECK(cond, ...)                         \
  do {                                           \
    if (!(cond)) {                               \
      fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
      fprintf(stderr, __VA_ARGS__);              \
      fprintf(stderr, "\n");                     \
      return 1;                                  \
    }                                            \
  } while (0)

int main(int argc, char** argv) {
  int devices[2] = {0, 0};
  if (argc >= 3) { devices[0] = atoi(argv[1]); devices[1] = atoi(argv[2]); }
  /* circulant 3-regular graph: vertex v -- v+1 (N edges) and v -- v+N/2 (N/2 edges); one index per edge */
  static uint32_t leaf[N][W];
  memset(leaf, 0, sizeof leaf);
  int k = 0;
  for (int v = 0; v < N; ++v, ++k) {
    leaf[v][k >> 5] |= 1u << (k & 31);
    leaf[(v + 1) % N][k >> 5] |= 1u << (k & 31);
  }
  for (int v = 0; v < N / 2; ++v, ++k) {
    leaf[v][k >> 5] |= 1u << (k & 31);
    leaf[v + N / 2][k >> 5] |= 1u << (k & 31);
  }
  CHECK(k == NI, "index count");
  uint64_t seeds[CHAINS];
  for (int i = 0; i < CHAINS; ++i) seeds[i] = 1000u + 7u * (uint64_t)i;
  double betas[SWEEPS];
  for (int s = 0; s < SWEEPS; ++s) betas[s] = 100.0 * s / SWEEPS;

  /* ---- the group */
  tnb_group* g = NULL;
  CHECK(tnb_group_create(&g, devices, 2) == 0, "tnb_group_create: %s", tnb_last_error(NULL));
  CHECK(tnb_group_size(g) == 2, "group size");
  CHECK(tnb_group_set_network(g, N, NI, &leaf[0][0], 2, NULL, NULL) == 0, "%s", tnb_group_last_error(g));
  CHECK(tnb_group_set_mode(g, -1.0, 10, 0, TNB_PROB_MH, TNB_RNG_PHILOX, TNB_LAYOUT_AUTO) == 0, "%s", tnb_group_last_error(g));
  CHECK(tnb_group_set_betas(g, betas, SWEEPS) == 0, "%s", tnb_group_last_error(g));
  CHECK(tnb_group_generate_chains(g, CHAINS, seeds, TNB_TREES_GREEDY) == 0, "%s", tnb_group_last_error(g));
  static double t0[CHAINS], t1[CHAINS], m1[CHAINS];
  CHECK(tnb_group_get_costs(g, t0, NULL) == 0, "%s", tnb_group_last_error(g));
  int64_t reached = 0;
  CHECK(tnb_group_run(g, SWEEPS, 60.0, &reached) == 0, "%s", tnb_group_last_error(g));
  CHECK(reached == SWEEPS, "reached %lld", (long long)reached);
  CHECK(tnb_group_get_costs(g, t1, m1) == 0, "%s", tnb_group_last_error(g));
  uint64_t props = 0, acc = 0, sw = 0;
  CHECK(tnb_group_get_counters(g, &props, &acc, &sw) == 0, "%s", tnb_group_last_error(g));
  CHECK(sw == (uint64_t)CHAINS * SWEEPS && props > sw && acc > 0 && acc <= props, "counters");
  double best = 0;
  int64_t chain = -1;
  static int32_t par[2 * N - 1], c0[2 * N - 1], c1[2 * N - 1];
  CHECK(tnb_group_get_best(g, &best, &chain, par, c0, c1, NULL) == 0, "%s", tnb_group_last_error(g));
  double mean0 = 0, mean1 = 0;
  for (int i = 0; i < CHAINS; ++i) {
    CHECK(m1[i] > 0 && m1[i] <= t1[i] && m1[i] <= t0[i], "chain %d: min %g total %g initial %g", i, m1[i], t1[i], t0[i]);
    CHECK(best <= m1[i], "best is not the minimum");
    mean0 += log2(t0[i]) / CHAINS;
    mean1 += log2(m1[i]) / CHAINS;
  }
  CHECK(mean1 < mean0, "annealing did not improve the mean cost (%g -> %g)", mean0, mean1);
  CHECK(chain >= 0 && chain < CHAINS && best == m1[chain], "winner");
  /* the winner's tree: a valid binary tree over the N leaves ... */
  int seen[2 * N - 1];
  memset(seen, 0, sizeof seen);
  for (int z = 0; z < 2 * N - 1; ++z) {
    if (z < N) { CHECK(c0[z] == -1 && c1[z] == -1, "leaf with children"); continue; }
    CHECK(c0[z] >= 0 && c0[z] < 2 * N - 1 && c1[z] >= 0 && c1[z] < 2 * N - 1 && c0[z] != c1[z], "children of %d", z);
    CHECK(par[c0[z]] == z && par[c1[z]] == z, "parent of the children of %d", z);
    seen[c0[z]]++; seen[c1[z]]++;
  }
  for (int z = 0; z < 2 * N - 2; ++z) CHECK(seen[z] == 1, "node %d is a child %d times", z, seen[z]);
  CHECK(par[2 * N - 2] == -1, "root");
  /* ... whose cost re-evaluates to the reported minimum (full-tree evaluation on device 0's engine) */
  double seq = 0, pc = 0, mw = 0;
  CHECK(tnb_eval_cost(tnb_group_engine(g, 0), 1, par, c0, c1, NULL, &seq, &pc, &mw) == 0, "%s", tnb_last_error(tnb_group_engine(g, 0)));
  CHECK(fabs(log2(pc) - log2(best)) < 1e-9, "re-evaluated cost %g vs reported %g", pc, best);

  /* ---- one engine, same seeds: identical chains (global chain ids enter the Philox counters) */
  tnb_engine* e = NULL;
  CHECK(tnb_create(&e, devices[0]) == 0, "tnb_create: %s", tnb_last_error(NULL));
  CHECK(tnb_set_network(e, N, NI, &leaf[0][0], 2, NULL) == 0, "%s", tnb_last_error(e));
  CHECK(tnb_set_mode(e, -1.0, 10, 0, TNB_PROB_MH, TNB_RNG_PHILOX, TNB_LAYOUT_AUTO) == 0, "%s", tnb_last_error(e));
  CHECK(tnb_set_betas(e, betas, SWEEPS) == 0, "%s", tnb_last_error(e));
  CHECK(tnb_generate_chains(e, CHAINS, seeds, 0, TNB_TREES_GREEDY) == 0, "%s", tnb_last_error(e));
  CHECK(tnb_run(e, SWEEPS) == 0, "%s", tnb_last_error(e));
  static double t2[CHAINS], m2[CHAINS];
  CHECK(tnb_get_costs(e, t2, m2) == 0, "%s", tnb_last_error(e));
  for (int i = 0; i < CHAINS; ++i)
    CHECK(t1[i] == t2[i] && m1[i] == m2[i], "chain %d differs between 2 devices and 1: %.17g vs %.17g", i, m1[i], m2[i]);
  /* ---- a wall-clock budget far too small for the request: stops early, chains stay valid */
  int64_t r2 = 0;
  CHECK(tnb_generate_chains(e, CHAINS, seeds, 0, TNB_TREES_GREEDY) == 0, "%s", tnb_last_error(e));
  static double many[1 << 20];
  for (int s = 0; s < (1 << 20); ++s) many[s] = 100.0 * s / (1 << 20);
  CHECK(tnb_set_betas(e, many, 1 << 20) == 0, "%s", tnb_last_error(e));
  CHECK(tnb_run_timed(e, 1 << 20, 0.2, &r2) == 0, "%s", tnb_last_error(e));
  CHECK(r2 >= 16 && r2 < (1 << 20), "timed run reached %lld", (long long)r2);
  CHECK(tnb_get_costs(e, t2, m2) == 0 && m2[0] > 0 && m2[0] <= t2[0], "costs after the timed run");
  tnb_destroy(e);
  tnb_group_destroy(g);
  printf("ok: 2 engines on devices %d,%d: %d chains x %d sweeps, %llu proposals, mean log2 cost %.3f -> %.3f, best %.3f (chain %lld); "
         "identical to one engine; timed run stopped at sweep %lld\n",
         devices[0], devices[1], CHAINS, SWEEPS, (unsigned long long)props, mean0, mean1, log2(best), (long long)chain, (long long)r2);
  return 0;
}
