/*
 * TEST INFRASTRUCTURE ONLY -- see sa_oracle.h for the scope statement and the list of reference
 * files (file:line) every function below follows.  Not part of the product path.
 */
#include "sa_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ std::mt19937 (libstdc++) */
typedef struct {
  uint32_t x[624];
  int p;
  uint32_t* rec;
  uint64_t rec_cap, rec_n, drawn;
  /* replay: draws come from a caller-supplied word stream instead of the generator (ora_set_replay) */
  const uint32_t* rp;
  uint64_t rp_n, rp_k;
  int rp_over;
} mt_t;

static void mt_seed(mt_t* m, uint32_t s) { /* optimize/optimizer.hpp:75 prng.seed(size_t) -> s mod 2^32 */
  m->x[0] = s;
  for (int i = 1; i < 624; ++i) m->x[i] = 1812433253u * (m->x[i - 1] ^ (m->x[i - 1] >> 30)) + (uint32_t)i;
  m->p = 624;
  m->rec = NULL;
  m->rec_cap = m->rec_n = m->drawn = 0;
  m->rp = NULL;
  m->rp_n = m->rp_k = 0;
  m->rp_over = 0;
}

static void mt_twist(mt_t* m) {
  uint32_t* x = m->x;
  for (int i = 0; i < 624; ++i) {
    uint32_t y = (x[i] & 0x80000000u) | (x[(i + 1) % 624] & 0x7fffffffu);
    x[i] = x[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  }
  m->p = 0;
}

static uint32_t mt_next(mt_t* m) {
  if (m->rp) {
    m->drawn++;
    if (m->rp_k >= m->rp_n) { m->rp_over = 1; return 0u; }
    return m->rp[m->rp_k++];
  }
  if (m->p >= 624) mt_twist(m);
  uint32_t y = m->x[m->p++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  m->drawn++;
  if (m->rec && m->rec_n < m->rec_cap) m->rec[m->rec_n] = y;
  if (m->rec) m->rec_n++;
  return y;
}

/* std::uniform_real_distribution<double>{}(prng) == generate_canonical<double,53>: two 32-bit draws,
 * (lo + hi*2^32) / 2^64, clamped below 1 (infinite_memory/optimizer.hpp:100,162). */
static double mt_uniform(mt_t* m) {
  const double lo = (double)mt_next(m);
  const double hi = (double)mt_next(m);
  double u = (lo + hi * 4294967296.0) / 18446744073709551616.0;
  if (u >= 1.0) u = nextafter(1.0, 0.0);
  return u;
}

/* libstdc++ uniform_int_distribution<size_t> over a 32-bit URNG, range = b-a+1 <= 2^32 (Lemire). */
static uint32_t mt_nd(mt_t* m, uint32_t range) {
  uint64_t prod = (uint64_t)mt_next(m) * (uint64_t)range;
  uint32_t low = (uint32_t)prod;
  if (low < range) {
    const uint32_t thr = (uint32_t)(0u - range) % range;
    while (low < thr) {
      prod = (uint64_t)mt_next(m) * (uint64_t)range;
      low = (uint32_t)prod;
    }
  }
  return (uint32_t)(prod >> 32);
}

/* std::shuffle, GCC 13 bits/stl_algo.h:3742-3805 ("two swaps per draw" branch; n*n <= 2^32 holds here) */
static void mt_shuffle(mt_t* m, int32_t* v, int n) {
  if (n == 0) return;
  int i = 1;
#define SWAP(a, b) do { int32_t t_ = v[a]; v[a] = v[b]; v[b] = t_; } while (0)
  if ((n % 2) == 0) {
    const uint32_t j = mt_nd(m, 2);
    SWAP(1, j);
    i = 2;
  }
  while (i != n) {
    const uint32_t r = (uint32_t)i + 1; /* swap range of first element = i+1, second = i+2 */
    const uint32_t x = mt_nd(m, r * (r + 1));
    const uint32_t a = x / (r + 1), b = x % (r + 1);
    SWAP(i, a);
    ++i;
    SWAP(i, b);
    ++i;
  }
#undef SWAP
}

/* ------------------------------------------------------------------ chain */
struct ora_chain {
  int n, N, n_inds, W;
  int finite, dsi;
  uint64_t dim;       /* uniform dim when dims == NULL */
  uint64_t* dims;     /* per-index dims or NULL */
  float max_width;    /* width_type = float32 (tnco/app/app.py:757) */
  uint32_t* skip;     /* [W] skip_slices: indices the slicer must not take, or NULL */
  uint32_t* sparse;   /* [W] sparse indices (SimpleCostModelSparseInds) or NULL */
  uint64_t n_projs;
  int32_t *par, *c0, *c1;
  uint32_t *bits, *hyper; /* [N][W] */
  double *cc, *pc;        /* contraction_cost, partial_cost */
  float* width;           /* WidthCache (finite) */
  uint32_t* slices;       /* [W] */
  /* min */
  int32_t *mpar, *mc0, *mc1;
  uint32_t *mbits, *mslices;
  double min_total;
  mt_t mt;
  uint64_t proposals, accepts, sweeps, width_rejects;
  int max_new_slices;         /* max_number_new_slices (finite_width/greedy/optimizer.hpp:77,226-321) */
  uint64_t new_slice_moves, new_slice_accepts;
  /* test hooks */
  ora_trace_rec* tr;          /* per-proposal trace (ora_trace) */
  uint64_t tr_cap, tr_n;
  const uint32_t* fs;         /* forced re-slices (ora_set_forced_slices): candidates [fs_n][W], keep flags */
  const uint8_t* fs_keep;
  uint64_t fs_n, fs_k;
  double* rl;                 /* re-slice log: (cost under candidate, cost under current) pairs */
  uint64_t rl_cap, rl_n;
};

static inline int popc_w(const uint32_t* a, int W) {
  int k = 0;
  for (int i = 0; i < W; ++i) k += __builtin_popcount(a[i]);
  return k;
}

/* infinite_memory/cost_model/simple.hpp:38-54 get_cost(inds, dims) */
static double cost_plain(const ora_chain* c, const uint32_t* u) {
  if (!c->dims) return pow((double)c->dim, (double)popc_w(u, c->W));
  double r = 1.0;
  for (int w = 0; w < c->W; ++w) {
    uint32_t x = u[w];
    while (x) {
      const int b = __builtin_ctz(x);
      r *= (double)c->dims[w * 32 + b];
      x &= x - 1;
    }
  }
  return r;
}

/* infinite_memory/cost_model/simple_sparse_inds.hpp:38-49:
 *   get_cost(inds - sparse) * min(get_cost(inds & sparse), n_projs),  min(x, y) = x < y ? x : y in cost_type */
static double cost_of(const ora_chain* c, const uint32_t* u) {
  if (!c->sparse) return cost_plain(c, u);
  uint32_t d[c->W], s[c->W];
  for (int i = 0; i < c->W; ++i) { d[i] = u[i] & ~c->sparse[i]; s[i] = u[i] & c->sparse[i]; }
  const double x = cost_plain(c, s), y = (double)c->n_projs;
  return cost_plain(c, d) * (x < y ? x : y);
}

/* contraction_cost(in1,in2,out,dims[,slices]) = get_cost(in1|in2[|slices]) (simple.hpp:66-83; fw :114-136;
 * sparse: simple_sparse_inds.hpp:69-86, fw :136-157) */
static double ccost_of(const ora_chain* c, const uint32_t* a, const uint32_t* b, const uint32_t* s) {
  uint32_t u[c->W];
  for (int i = 0; i < c->W; ++i) u[i] = a[i] | b[i] | (s ? s[i] : 0u);
  return cost_of(c, u);
}

/* finite_width/cost_model/simple.hpp:39-58 get_width<float> */
static float width_plain(const ora_chain* c, const uint32_t* u) {
  if (!c->dims) return (float)(log2((double)c->dim) * (double)popc_w(u, c->W));
  float wd = 0.0f;
  for (int w = 0; w < c->W; ++w) {
    uint32_t x = u[w];
    while (x) {
      const int b = __builtin_ctz(x);
      wd = (float)((double)wd + log2((double)c->dims[w * 32 + b]));
      x &= x - 1;
    }
  }
  return wd;
}

/* min(x, y) of finite_width/cost_model/simple_sparse_inds.hpp:43-45: x float, y = log2(n_projs) double */
static float min_w(float x, double y) { return (double)x < y ? x : (float)y; }

/* finite_width/cost_model/simple_sparse_inds.hpp:38-49 get_width<float> */
static float width_of(const ora_chain* c, const uint32_t* u) {
  if (!c->sparse) return width_plain(c, u);
  uint32_t d[c->W], s[c->W];
  for (int i = 0; i < c->W; ++i) { d[i] = u[i] & ~c->sparse[i]; s[i] = u[i] & c->sparse[i]; }
  return width_plain(c, d) + min_w(width_plain(c, s), log2((double)c->n_projs));
}

/* utils.hpp:35-52 traverse: post-order, children[0] subtree first; writes node ids to `order` */
static int post_order(int N, const int32_t* c0, const int32_t* c1, int32_t* order) {
  int32_t* stack = (int32_t*)malloc(sizeof(int32_t) * (size_t)(2 * N + 2));
  uint8_t* vis = (uint8_t*)calloc((size_t)N, 1);
  int sp = 0, k = 0;
  stack[sp++] = N - 1;
  while (sp) {
    const int32_t pos = stack[sp - 1];
    if (vis[pos] || c0[pos] < 0) {
      --sp;
      order[k++] = pos;
    } else {
      vis[pos] = 1;
      stack[sp++] = c1[pos];
      stack[sp++] = c0[pos];
    }
  }
  free(stack);
  free(vis);
  return k;
}

/* infinite_memory/utils.hpp:32-56 CostCache ctor (fw: with slices) */
static void build_cost_cache(const ora_chain* c, const uint32_t* slices, double* cc, double* pc) {
  int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)c->N);
  const int k = post_order(c->N, c->c0, c->c1, order);
  for (int i = 0; i < k; ++i) {
    const int pos = order[i];
    if (c->c0[pos] < 0) {
      cc[pos] = 0;
      pc[pos] = 0;
    } else {
      const int a = c->c0[pos], b = c->c1[pos];
      const double cost_A = ccost_of(c, c->bits + (size_t)a * c->W, c->bits + (size_t)b * c->W, slices);
      cc[pos] = cost_A;
      pc[pos] = cost_A + pc[a] + pc[b];
    }
  }
  free(order);
}

/* infinite_memory/utils.hpp:102-116 get_cost: sequential sum in traverse order */
static double seq_cost(const ora_chain* c, const uint32_t* slices) {
  int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)c->N);
  const int k = post_order(c->N, c->c0, c->c1, order);
  double t = 0;
  for (int i = 0; i < k; ++i) {
    const int pos = order[i];
    if (c->c0[pos] >= 0)
      t += ccost_of(c, c->bits + (size_t)c->c0[pos] * c->W, c->bits + (size_t)c->c1[pos] * c->W, slices);
  }
  free(order);
  return t;
}

/* finite_width/greedy/utils.hpp:24-125 get_slices_impl (skip_slices = nullopt, as the app drives it) */
static void get_slices(ora_chain* c, uint32_t* out) {
  const int W = c->W, N = c->N;
  memset(out, 0, sizeof(uint32_t) * (size_t)W);
  uint32_t* nbig = (uint32_t*)calloc((size_t)W * 32, sizeof(uint32_t));
  for (int t = 0; t < N; ++t)
    if (c->width[t] > c->max_width) {
      const uint32_t* x = c->bits + (size_t)t * W;
      for (int w = 0; w < W; ++w) {
        uint32_t v = x[w];
        while (v) {
          nbig[w * 32 + __builtin_ctz(v)]++;
          v &= v - 1;
        }
      }
    }
  int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
  int32_t* pos = (int32_t*)malloc(sizeof(int32_t) * (size_t)W * 32);
  uint32_t xs[W];
  const int k = post_order(N, c->c0, c->c1, order);
  for (int i = 0; i < k; ++i) {
    const int t = order[i];
    if (!(c->width[t] > c->max_width)) continue;
    for (int w = 0; w < W; ++w) xs[w] = c->bits[(size_t)t * W + w] & ~out[w];
    float sw = width_of(c, xs);
    if (!(sw > c->max_width)) continue;
    int np = 0;
    for (int w = 0; w < W; ++w) {
      uint32_t v = xs[w] & ~(c->skip ? c->skip[w] : 0u); /* sliced_xs - skip_slices (:76-79) */
      while (v) {
        pos[np++] = w * 32 + __builtin_ctz(v);
        v &= v - 1;
      }
    }
    mt_shuffle(&c->mt, pos, np);
    /* std::stable_sort(greater): insertion sort is stable and yields the same permutation */
    for (int a = 1; a < np; ++a) {
      const int32_t key = pos[a];
      int b = a - 1;
      while (b >= 0) {
        const int32_t y = pos[b];
        int key_before_y; /* greater(key, y) */
        if (!c->dims) key_before_y = nbig[key] > nbig[y];
        else {
          const float lk = (float)log2((double)c->dims[key]), ly = (float)log2((double)c->dims[y]);
          key_before_y = nbig[key] == nbig[y] ? (lk > ly) : (nbig[key] > nbig[y]);
        }
        if (!key_before_y) break;
        pos[b + 1] = y;
        --b;
      }
      pos[b + 1] = key;
    }
    for (int a = 0; a < np; ++a) {
      const int p = pos[a];
      out[p >> 5] |= 1u << (p & 31);
      /* get_delta_width: (1 - 2*test(pos)) * log2(dim) as float, added in float (simple.hpp:60-76) */
      const double l2 = log2((double)(c->dims ? c->dims[p] : c->dim));
      const int tst = (xs[p >> 5] >> (p & 31)) & 1;
      if (c->sparse && ((c->sparse[p >> 5] >> (p & 31)) & 1)) {
        /* sparse index (simple_sparse_inds.hpp:51-77): min(width(new & sparse), L) - min(width(old & sparse), L) */
        uint32_t so[W], sn[W];
        for (int w = 0; w < W; ++w) so[w] = sn[w] = xs[w] & c->sparse[w];
        sn[p >> 5] ^= 1u << (p & 31);
        const double L = log2((double)c->n_projs);
        sw += min_w(width_plain(c, sn), L) - min_w(width_plain(c, so), L);
      } else {
        sw += (float)((double)(1 - 2 * tst) * l2);
      }
      xs[p >> 5] &= ~(1u << (p & 31));
      if (sw <= c->max_width) break;
    }
  }
  free(order);
  free(pos);
  free(nbig);
}

static void copy_min(ora_chain* c) {
  memcpy(c->mpar, c->par, sizeof(int32_t) * (size_t)c->N);
  memcpy(c->mc0, c->c0, sizeof(int32_t) * (size_t)c->N);
  memcpy(c->mc1, c->c1, sizeof(int32_t) * (size_t)c->N);
  memcpy(c->mbits, c->bits, sizeof(uint32_t) * (size_t)c->N * c->W);
  if (c->finite) memcpy(c->mslices, c->slices, sizeof(uint32_t) * (size_t)c->W);
}

ora_chain* ora_create(int n, int n_inds, const int32_t* parent, const int32_t* child0, const int32_t* child1,
                      const uint32_t* node_bits, uint64_t dim, const uint64_t* dims, int finite,
                      float max_width, uint32_t seed, int dsi, int* err) {
  return ora_create_sparse(n, n_inds, parent, child0, child1, node_bits, dim, dims, finite, max_width, seed, dsi,
                           NULL, 0, err);
}

ora_chain* ora_create_sparse(int n, int n_inds, const int32_t* parent, const int32_t* child0,
                             const int32_t* child1, const uint32_t* node_bits, uint64_t dim, const uint64_t* dims,
                             int finite, float max_width, uint32_t seed, int dsi, const uint32_t* sparse_bits,
                             uint64_t n_projs, int* err) {
  return ora_create_ex(n, n_inds, parent, child0, child1, node_bits, dim, dims, finite, max_width, seed, dsi,
                       sparse_bits, n_projs, NULL, err);
}

ora_chain* ora_create_ex(int n, int n_inds, const int32_t* parent, const int32_t* child0, const int32_t* child1,
                         const uint32_t* node_bits, uint64_t dim, const uint64_t* dims, int finite, float max_width,
                         uint32_t seed, int dsi, const uint32_t* sparse_bits, uint64_t n_projs,
                         const uint32_t* skip_bits, int* err) {
  return ora_create_ex2(n, n_inds, parent, child0, child1, node_bits, dim, dims, finite, max_width, seed, dsi,
                        sparse_bits, n_projs, skip_bits, NULL, err);
}

ora_chain* ora_create_ex2(int n, int n_inds, const int32_t* parent, const int32_t* child0, const int32_t* child1,
                          const uint32_t* node_bits, uint64_t dim, const uint64_t* dims, int finite, float max_width,
                          uint32_t seed, int dsi, const uint32_t* sparse_bits, uint64_t n_projs,
                          const uint32_t* skip_bits, const uint32_t* init_slices, int* err) {
  if (err) *err = 0;
  if (sparse_bits && n_projs == 0) { /* "'n_projs' must be a positive number." (simple_sparse_inds.hpp:64-67) */
    if (err) *err = 3;
    return NULL;
  }
  ora_chain* c = (ora_chain*)calloc(1, sizeof(ora_chain));
  const int N = 2 * n - 1, W = (n_inds + 31) / 32;
  c->n = n; c->N = N; c->n_inds = n_inds; c->W = W;
  c->finite = finite; c->dsi = dsi; c->dim = dim; c->max_width = max_width;
  if (skip_bits) {
    c->skip = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)W);
    memcpy(c->skip, skip_bits, sizeof(uint32_t) * (size_t)W);
  }
  if (sparse_bits) {
    c->sparse = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)W);
    memcpy(c->sparse, sparse_bits, sizeof(uint32_t) * (size_t)W);
    c->n_projs = n_projs;
  }
  if (dims) {
    c->dims = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)W * 32);
    memset(c->dims, 0, sizeof(uint64_t) * (size_t)W * 32);
    memcpy(c->dims, dims, sizeof(uint64_t) * (size_t)n_inds);
  }
#define AL(T, cnt) (T*)calloc((size_t)(cnt), sizeof(T))
  c->par = AL(int32_t, N); c->c0 = AL(int32_t, N); c->c1 = AL(int32_t, N);
  c->mpar = AL(int32_t, N); c->mc0 = AL(int32_t, N); c->mc1 = AL(int32_t, N);
  c->bits = AL(uint32_t, (size_t)N * W); c->hyper = AL(uint32_t, (size_t)N * W);
  c->mbits = AL(uint32_t, (size_t)N * W);
  c->cc = AL(double, N); c->pc = AL(double, N); c->width = AL(float, N);
  c->slices = AL(uint32_t, W); c->mslices = AL(uint32_t, W);
#undef AL
  memcpy(c->par, parent, sizeof(int32_t) * (size_t)N);
  memcpy(c->c0, child0, sizeof(int32_t) * (size_t)N);
  memcpy(c->c1, child1, sizeof(int32_t) * (size_t)N);
  memcpy(c->bits, node_bits, sizeof(uint32_t) * (size_t)N * W);
  mt_seed(&c->mt, seed);
  /* HyperCache (infinite_memory/utils.hpp:68-100) */
  for (int z = 0; z < N; ++z)
    if (c->c0[z] >= 0)
      for (int w = 0; w < W; ++w)
        c->hyper[(size_t)z * W + w] = c->bits[(size_t)z * W + w] & c->bits[(size_t)c->c0[z] * W + w] &
                                      c->bits[(size_t)c->c1[z] * W + w];
  if (finite) {
    /* finite_width/greedy/optimizer.hpp:80-101: WidthCache -> slices (consumes prng) -> CostCache */
    for (int t = 0; t < N; ++t) c->width[t] = width_of(c, c->bits + (size_t)t * W);
    if (init_slices) memcpy(c->slices, init_slices, sizeof(uint32_t) * (size_t)W); /* `slices=` given (:80-88) */
    else get_slices(c, c->slices);
  }
  build_cost_cache(c, finite ? c->slices : NULL, c->cc, c->pc);
  copy_min(c);
  c->min_total = seq_cost(c, finite ? c->slices : NULL); /* get_cost(min_ctree[, min_slices]) */
  const double l1 = log2(c->pc[N - 1]), l2 = log2(c->min_total);
  if (isinf(l1) || isnan(l1) || isinf(l2) || isnan(l2)) {
    /* "Precision is too low." (infinite_memory/optimizer.hpp:77-87) */
    if (err) *err = 2;
    ora_destroy(c);
    return NULL;
  }
  return c;
}

void ora_destroy(ora_chain* c) {
  if (!c) return;
  free(c->skip); free(c->sparse); free(c->dims); free(c->par); free(c->c0); free(c->c1); free(c->mpar); free(c->mc0); free(c->mc1);
  free(c->bits); free(c->hyper); free(c->mbits); free(c->cc); free(c->pc); free(c->width);
  free(c->slices); free(c->mslices);
  free(c);
}

/* prob/mh.hpp:45-59, greedy.hpp:38-42, base.hpp:43-47 */
static double prob_of(int kind, double beta, double delta, double old) {
  if (kind == ORA_PROB_ALWAYS) return 1.0;
  if (kind == ORA_PROB_GREEDY) return delta <= 0 ? 1.0 : 0.0;
  if (delta <= 0) return 1.0;
  if (old == 0) return 0.0;
  return pow(1 + (delta / old), -beta);
}

void ora_update(ora_chain* c, int kind, double beta, int update_slices) {
  const int W = c->W, N = c->N;
  mt_t* mt = &c->mt;
#define BITS(t) (c->bits + (size_t)(t) * W)
#define HYP(t) (c->hyper + (size_t)(t) * W)
  c->sweeps++;
  int B = (int)(mt_next(mt) % (uint32_t)c->n); /* optimizer.hpp:103 */
  if ((B = c->par[B]) < 0) return;             /* :107 */
  double total = c->pc[N - 1];                 /* :112 */
  const uint32_t* S = c->finite ? c->slices : NULL;
  uint32_t newB[W], tmp[W];
  while (1) {
    /* get_ctree_nn (optimize/optimizer.hpp:86-172) */
    const int A = c->par[B];
    if (A < 0) break;
    int C = (c->c0[A] == B) ? c->c1[A] : c->c0[A];
    const int p0 = c->c0[B], p1 = c->c1[B];
    int i0 = 0, i1 = 0;
    for (int w = 0; w < W; ++w) {
      i0 |= (BITS(p0)[w] & BITS(C)[w]) != 0;
      i1 |= (BITS(p1)[w] & BITS(C)[w]) != 0;
    }
    int D, E;
    if (c->dsi || (i0 && i1)) {
      if (mt_next(mt) % 2) { D = p0; E = p1; } else { D = p1; E = p0; }
    } else if (i0) { D = p0; E = p1; } else { D = p1; E = p0; }
    for (int w = 0; w < W; ++w) newB[w] = (BITS(D)[w] ^ BITS(C)[w]) | HYP(A)[w] | HYP(B)[w]; /* :147 */
    c->proposals++;
    int gate = 1;
    float new_width_B = 0;
    if (c->finite) { /* finite_width/greedy/optimizer.hpp:176-188 */
      new_width_B = width_of(c, newB);
      for (int w = 0; w < W; ++w) tmp[w] = newB[w] & ~S[w];
      gate = width_of(c, tmp) <= c->max_width;
      if (!gate) c->width_rejects++;
    }
    ora_trace_rec* tr = NULL;
    if (c->tr) {
      if (c->tr_n < c->tr_cap) {
        tr = c->tr + c->tr_n;
        memset(tr, 0, sizeof *tr);
        tr->B = B; tr->A = A; tr->pick0 = (uint8_t)(D == p0); tr->gate = (uint8_t)gate;
        tr->coin = (uint8_t)(c->dsi || (i0 && i1)); tr->total = total;
      }
      c->tr_n++;
    }
    if (gate) {
      const double nA = ccost_of(c, newB, BITS(E), S);    /* :152-155 */
      const double nB = ccost_of(c, BITS(D), BITS(C), S);
      const double delta = (nB - c->cc[B]) + (nA - c->cc[A]); /* :158 */
      const double u = mt_uniform(mt);
      const double pr = prob_of(kind, beta, delta, total);
      if (tr) { tr->delta = delta; tr->u = u; tr->p = pr; tr->acc = (uint8_t)(u <= pr); }
      if (u <= pr) {        /* :162 */
        /* tree.hpp:141-192 swap_with_nn(E): E <-> C, child slots preserved */
        if (c->c0[A] == C) c->c0[A] = E; else c->c1[A] = E;
        if (c->c0[B] == E) c->c0[B] = C; else c->c1[B] = C;
        c->par[C] = B;
        c->par[E] = A;
        { const int t = C; C = E; E = t; }
        memcpy(BITS(B), newB, sizeof(uint32_t) * (size_t)W);
        for (int w = 0; w < W; ++w) {
          HYP(A)[w] = BITS(A)[w] & BITS(B)[w] & BITS(C)[w]; /* inds_E after the name swap == new C */
          HYP(B)[w] = BITS(B)[w] & BITS(D)[w] & BITS(E)[w];
        }
        c->cc[B] = nB;
        c->cc[A] = nA;
        total += delta;
        if (c->finite) c->width[B] = new_width_B;
        c->accepts++;
      }
    }
    int skip_prop = 0;
    if (!gate && c->max_new_slices > 0) { /* finite_width/greedy/optimizer.hpp:226-321: random new slices */
      uint32_t ns[W];
      memcpy(ns, S, sizeof(uint32_t) * (size_t)W);
      int32_t* pos = (int32_t*)malloc(sizeof(int32_t) * (size_t)W * 32);
      int n_pos = 0, n_new = 0;
      for (int w = 0; w < W; ++w) { /* (new_inds_B - slices [- skip_slices]).positions(), ascending */
        uint32_t v = newB[w] & ~S[w] & ~(c->skip ? c->skip[w] : 0u);
        while (v) { pos[n_pos++] = w * 32 + __builtin_ctz(v); v &= v - 1; }
      }
      float nsw = width_of(c, tmp); /* new_sliced_width_B (tmp = newB & ~S) */
      while (n_new < c->max_new_slices && nsw > c->max_width && n_pos > 0) {
        const uint32_t r = mt_next(mt) % (uint32_t)n_pos; /* :246 */
        const int32_t t_ = pos[r]; pos[r] = pos[n_pos - 1]; pos[n_pos - 1] = t_;
        const int p = pos[n_pos - 1];
        ns[p >> 5] |= 1u << (p & 31);
        /* new_sliced_width_B -= log2_dims[...] in width_type, DimsCache<width_type> (:256-266) */
        nsw -= (float)log2((double)(c->dims ? c->dims[p] : c->dim));
        --n_pos; ++n_new;
      }
      free(pos);
      if (nsw <= c->max_width) { /* :283-318 */
        c->new_slice_moves++;
        uint32_t oldB[W];
        memcpy(oldB, BITS(B), sizeof(uint32_t) * (size_t)W);
        memcpy(BITS(B), newB, sizeof(uint32_t) * (size_t)W);
        /* swap_with_nn(E): E <-> C */
        if (c->c0[A] == C) c->c0[A] = E; else c->c1[A] = E;
        if (c->c0[B] == E) c->c0[B] = C; else c->c1[B] = C;
        c->par[C] = B;
        c->par[E] = A;
        double* cc2 = (double*)malloc(sizeof(double) * (size_t)N);
        double* pc2 = (double*)malloc(sizeof(double) * (size_t)N);
        build_cost_cache(c, ns, cc2, pc2);
        const double delta = pc2[N - 1] - total;
        const double u = mt_uniform(mt);
        if (u <= prob_of(kind, beta, delta, total)) {
          memcpy(c->cc, cc2, sizeof(double) * (size_t)N);
          memcpy(c->pc, pc2, sizeof(double) * (size_t)N);
          for (int w = 0; w < W; ++w) { /* pos_C / pos_E are NOT renamed in this branch (:300-302) */
            HYP(A)[w] = BITS(A)[w] & BITS(B)[w] & BITS(E)[w];
            HYP(B)[w] = BITS(B)[w] & BITS(D)[w] & BITS(C)[w];
          }
          c->width[B] = new_width_B;
          total = c->pc[N - 1];
          memcpy(c->slices, ns, sizeof(uint32_t) * (size_t)W);
          skip_prop = 1;
          c->accepts++;
          c->new_slice_accepts++;
        } else { /* swap back: swap_with_nn(pos_C) */
          if (c->c0[A] == E) c->c0[A] = C; else c->c1[A] = C;
          if (c->c0[B] == C) c->c0[B] = E; else c->c1[B] = E;
          c->par[C] = A;
          c->par[E] = B;
          memcpy(BITS(B), oldB, sizeof(uint32_t) * (size_t)W);
        }
        free(cc2);
        free(pc2);
      }
    }
    if (!skip_prop) {
      c->pc[B] = c->pc[D] + c->pc[E] + c->cc[B]; /* :185-188 */
      c->pc[A] = c->pc[B] + c->pc[C] + c->cc[A];
    }
    B = A;
  }
  if (c->finite && update_slices) { /* finite_width/greedy/optimizer.hpp:360-376 */
    int any = 0;
    for (int w = 0; w < W; ++w) any |= c->slices[w] != 0;
    if (any) {
      uint32_t ns[W];
      int forced = -1; /* test hook: candidate slices and the keep decision come from the caller */
      if (c->fs) {
        if (c->fs_k < c->fs_n) {
          memcpy(ns, c->fs + (size_t)c->fs_k * W, sizeof(uint32_t) * (size_t)W);
          forced = c->fs_keep[c->fs_k];
        } else {
          memcpy(ns, c->slices, sizeof(uint32_t) * (size_t)W);
          forced = 0;
          c->mt.rp_over = 1; /* ran out of forced re-slices: flag it like a stream overrun */
        }
        c->fs_k++;
      } else {
        get_slices(c, ns);
      }
      double* cc2 = (double*)malloc(sizeof(double) * (size_t)N);
      double* pc2 = (double*)malloc(sizeof(double) * (size_t)N);
      build_cost_cache(c, ns, cc2, pc2);
      if (c->rl) {
        if (c->rl_n < c->rl_cap) { c->rl[2 * c->rl_n] = pc2[N - 1]; c->rl[2 * c->rl_n + 1] = c->pc[N - 1]; }
        c->rl_n++;
      }
      if (forced >= 0 ? forced : (pc2[N - 1] < c->pc[N - 1])) {
        memcpy(c->slices, ns, sizeof(uint32_t) * (size_t)W);
        memcpy(c->cc, cc2, sizeof(double) * (size_t)N);
        memcpy(c->pc, pc2, sizeof(double) * (size_t)N);
      }
      free(cc2);
      free(pc2);
    }
  }
  if (c->pc[N - 1] < c->min_total) { /* :197-201 / fw :384-389 */
    c->min_total = c->pc[N - 1];
    copy_min(c);
  }
#undef BITS
#undef HYP
}

void ora_run(ora_chain* c, int kind, const double* betas, int64_t n, int every, int64_t off) {
  for (int64_t i = 0; i < n; ++i) ora_update(c, kind, betas[i], every > 0 ? (((off + i) % every) == 0) : 0);
}

void ora_get_tree(const ora_chain* c, int m, int32_t* p, int32_t* a, int32_t* b) {
  memcpy(p, m ? c->mpar : c->par, sizeof(int32_t) * (size_t)c->N);
  memcpy(a, m ? c->mc0 : c->c0, sizeof(int32_t) * (size_t)c->N);
  memcpy(b, m ? c->mc1 : c->c1, sizeof(int32_t) * (size_t)c->N);
}
void ora_get_bits(const ora_chain* c, int m, uint32_t* o) {
  memcpy(o, m ? c->mbits : c->bits, sizeof(uint32_t) * (size_t)c->N * c->W);
}
void ora_get_slices(const ora_chain* c, int m, uint32_t* o) {
  memcpy(o, m ? c->mslices : c->slices, sizeof(uint32_t) * (size_t)c->W);
}
double ora_total_cost(const ora_chain* c) { return c->pc[c->N - 1]; }
double ora_min_total_cost(const ora_chain* c) { return c->min_total; }
double ora_log2_total_cost(const ora_chain* c) { return log2(c->pc[c->N - 1]); }
double ora_log2_min_total_cost(const ora_chain* c) { return log2(c->min_total); }
void ora_get_costs(const ora_chain* c, double* cc, double* pc) {
  memcpy(cc, c->cc, sizeof(double) * (size_t)c->N);
  memcpy(pc, c->pc, sizeof(double) * (size_t)c->N);
}
void ora_prng_state(const ora_chain* c, uint32_t* s, int* pos) {
  memcpy(s, c->mt.x, sizeof(uint32_t) * 624);
  *pos = c->mt.p;
}
void ora_counters(const ora_chain* c, uint64_t* p, uint64_t* a, uint64_t* s, uint64_t* w, uint64_t* wr) {
  if (p) *p = c->proposals;
  if (a) *a = c->accepts;
  if (s) *s = c->sweeps;
  if (w) *w = c->mt.drawn;
  if (wr) *wr = c->width_rejects;
}
void ora_set_max_new_slices(ora_chain* c, int max_number_new_slices) { c->max_new_slices = max_number_new_slices; }
void ora_new_slice_counters(const ora_chain* c, uint64_t* moves, uint64_t* accepts) {
  if (moves) *moves = c->new_slice_moves;
  if (accepts) *accepts = c->new_slice_accepts;
}
void ora_set_replay(ora_chain* c, const uint32_t* words, uint64_t n) {
  c->mt.rp = words;
  c->mt.rp_n = n;
  c->mt.rp_k = 0;
  c->mt.rp_over = 0;
}
void ora_replay_state(const ora_chain* c, uint64_t* consumed, int* overrun) {
  if (consumed) *consumed = c->mt.rp_k;
  if (overrun) *overrun = c->mt.rp_over;
}
void ora_trace(ora_chain* c, ora_trace_rec* buf, uint64_t cap) {
  c->tr = buf;
  c->tr_cap = cap;
  c->tr_n = 0;
}
uint64_t ora_traced(const ora_chain* c) { return c->tr_n; }
void ora_set_forced_slices(ora_chain* c, const uint32_t* cand, const uint8_t* keep, uint64_t n, double* log2x,
                           uint64_t log_cap) {
  c->fs = cand;
  c->fs_keep = keep;
  c->fs_n = n;
  c->fs_k = 0;
  c->rl = log2x;
  c->rl_cap = log_cap;
  c->rl_n = 0;
}
uint64_t ora_forced_used(const ora_chain* c) { return c->fs_k; }
void ora_record(ora_chain* c, uint32_t* buf, uint64_t cap) {
  c->mt.rec = buf;
  c->mt.rec_cap = cap;
  c->mt.rec_n = 0;
}
uint64_t ora_recorded(const ora_chain* c) { return c->mt.rec_n; }

void ora_mt_stream(uint32_t seed, uint64_t n, uint32_t* out) {
  mt_t m;
  mt_seed(&m, seed);
  for (uint64_t i = 0; i < n; ++i) out[i] = mt_next(&m);
}

double ora_tree_cost(int n, int n_inds, const int32_t* c0, const int32_t* c1, const uint32_t* node_bits,
                     uint64_t dim, const uint64_t* dims, const uint32_t* slices, double* mw, double* pcr) {
  ora_chain c;
  memset(&c, 0, sizeof c);
  c.n = n; c.N = 2 * n - 1; c.n_inds = n_inds; c.W = (n_inds + 31) / 32; c.dim = dim;
  c.dims = (uint64_t*)dims; c.c0 = (int32_t*)c0; c.c1 = (int32_t*)c1; c.bits = (uint32_t*)node_bits;
  const double t = seq_cost(&c, slices);
  if (mw) {
    double m = 0;
    uint32_t tmp[c.W];
    for (int z = 0; z < c.N; ++z) {
      for (int w = 0; w < c.W; ++w) tmp[w] = node_bits[(size_t)z * c.W + w] & ~(slices ? slices[w] : 0u);
      double wd = 0;
      if (!dims) wd = log2((double)dim) * popc_w(tmp, c.W);
      else
        for (int i = 0; i < n_inds; ++i)
          if ((tmp[i >> 5] >> (i & 31)) & 1) wd += log2((double)dims[i]);
      if (wd > m) m = wd;
    }
    *mw = m;
  }
  if (pcr) {
    double* cc = (double*)malloc(sizeof(double) * (size_t)c.N);
    double* pc = (double*)malloc(sizeof(double) * (size_t)c.N);
    build_cost_cache(&c, slices, cc, pc);
    *pcr = pc[c.N - 1];
    free(cc);
    free(pc);
  }
  return t;
}

int ora_get_contraction(int n, const int32_t* c0, const int32_t* c1, int32_t* tr) {
  const int N = 2 * n - 1;
  int32_t* order = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
  const int k = post_order(N, c0, c1, order);
  int m = 0;
  for (int i = 0; i < k; ++i) {
    const int pos = order[i];
    if (c0[pos] >= 0) {
      tr[3 * m] = c0[pos];
      tr[3 * m + 1] = c1[pos];
      tr[3 * m + 2] = pos;
      ++m;
    }
  }
  free(order);
  return m;
}
