// Host-side (no GPU) entry points of the C-ABI: initial trees, tree <-> linear path, mt19937 streams.
// See include/tnco_b200.h for the reference interfaces each one replaces.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "tnb_internal.h"

namespace tnb {

static thread_local std::string g_err;
void set_global_error(const std::string& s) { g_err = s; }
const char* global_error() { return g_err.c_str(); }

// ----------------------------------------------------------------------------------------- mt19937
// Bit-compatible with std::mt19937 seeded by prng.seed(s) (include/tnco/optimize/optimizer.hpp:75).
void Mt19937::seed(uint32_t s) {
  x[0] = s;
  for (int i = 1; i < 624; ++i) x[i] = 1812433253u * (x[i - 1] ^ (x[i - 1] >> 30)) + uint32_t(i);
  p = 624;
}
void Mt19937::refill() {
  auto mix = [](uint32_t a, uint32_t b) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
  };
  int i = 0;
  for (; i < 624 - 397; ++i) x[i] = x[i + 397] ^ mix(x[i], x[i + 1]);
  for (; i < 623; ++i) x[i] = x[i + 397 - 624] ^ mix(x[i], x[i + 1]);
  x[623] = x[396] ^ mix(x[623], x[0]);
  p = 0;
}
void Mt19937::fill(uint32_t* out, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) out[i] = next();
}

// ----------------------------------------------------------------------------------------- small rng
struct SplitMix {
  uint64_t s;
  explicit SplitMix(uint64_t seed) : s(seed) {}
  uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  uint32_t below(uint32_t n) { return uint32_t((uint64_t(uint32_t(next() >> 32)) * n) >> 32); }
};

// ----------------------------------------------------------------------------------------- trees
static inline int popc_row(const uint32_t* a, int W) {
  int k = 0;
  for (int i = 0; i < W; ++i) k += __builtin_popcount(a[i]);
  return k;
}

struct Net {
  int n, n_inds, W;
  const uint32_t* leaf_bits;
  std::vector<int32_t> own0, own1;            // the first two leaves holding each index
  std::vector<std::vector<int32_t>> holders;  // all leaves holding each index
  std::vector<int> hcount0;                   // hyper count: holders - 1 (+1 for output indices), ctree.py:138-156
  bool hyper = false;                         // some index has hyper count >= 2
};

static bool build_net(Net& net, const uint32_t* output_bits, std::string& err) {
  net.own0.assign(net.n_inds, -1);
  net.own1.assign(net.n_inds, -1);
  net.holders.assign(net.n_inds, {});
  for (int t = 0; t < net.n; ++t)
    for (int w = 0; w < net.W; ++w) {
      uint32_t v = net.leaf_bits[size_t(t) * net.W + w];
      while (v) {
        const int i = w * 32 + __builtin_ctz(v);
        v &= v - 1;
        if (i >= net.n_inds) { err = "leaf_bits has a bit beyond n_inds"; return false; }
        if (net.own0[i] < 0) net.own0[i] = t;
        else if (net.own1[i] < 0) net.own1[i] = t;
        net.holders[i].push_back(t);
      }
    }
  net.hcount0.assign(net.n_inds, 0);
  for (int i = 0; i < net.n_inds; ++i) {
    const int out = output_bits ? int((output_bits[i >> 5] >> (i & 31)) & 1u) : 0;
    net.hcount0[i] = std::max(0, int(net.holders[i].size()) - 1 + out);
    net.hyper |= net.hcount0[i] >= 2;
  }
  return true;
}

// One tree for a network WITH hyper-indices (general but slower than one_tree): clusters holding each index are
// tracked explicitly and intermediate index sets follow the hyper-count rule of tnco/ctree.py:169-189.
static bool one_tree_hyper(const Net& net, uint64_t seed, int method, int32_t* par, int32_t* c0, int32_t* c1) {
  const int n = net.n, N = 2 * n - 1, W = net.W;
  SplitMix rng(seed * 0x9E3779B97F4A7C15ull + 0x7654321ull + uint64_t(method));
  std::fill(par, par + N, -1);
  std::fill(c0, c0 + N, -1);
  std::fill(c1, c1 + N, -1);
  if (n == 1) return true;
  std::vector<uint32_t> bits(size_t(N) * W, 0u);
  std::memcpy(bits.data(), net.leaf_bits, sizeof(uint32_t) * size_t(n) * W);
  std::vector<int> cnt(net.hcount0);
  std::vector<std::vector<int32_t>> cur(net.holders);  // clusters currently holding each index
  std::vector<uint8_t> alive(N, 0);
  std::fill(alive.begin(), alive.begin() + n, 1);
  int nxt = n;
  auto out_size = [&](int a, int b) {  // popcount of the index set the contraction of a and b would get
    const uint32_t *ba = &bits[size_t(a) * W], *bb = &bits[size_t(b) * W];
    int k = 0;
    for (int w = 0; w < W; ++w) {
      k += __builtin_popcount(ba[w] ^ bb[w]);
      uint32_t v = ba[w] & bb[w];
      while (v) {
        const int i = w * 32 + __builtin_ctz(v);
        v &= v - 1;
        k += cnt[i] - 1 > 0;
      }
    }
    return k;
  };
  auto merge = [&](int a, int b) {
    const int z = nxt++;
    uint32_t* bz = &bits[size_t(z) * W];
    const uint32_t *ba = &bits[size_t(a) * W], *bb = &bits[size_t(b) * W];
    for (int w = 0; w < W; ++w) {
      uint32_t keep = 0u, v = ba[w] & bb[w];
      while (v) {
        const int bit = __builtin_ctz(v);
        v &= v - 1;
        if (--cnt[w * 32 + bit] > 0) keep |= 1u << bit;
      }
      bz[w] = (ba[w] ^ bb[w]) | keep;
      uint32_t u = ba[w] | bb[w];
      while (u) {
        const int bit = __builtin_ctz(u);
        u &= u - 1;
        auto& h = cur[size_t(w) * 32 + bit];
        h.erase(std::remove_if(h.begin(), h.end(), [&](int32_t x) { return x == a || x == b; }), h.end());
        if ((bz[w] >> bit) & 1u) h.push_back(z);
      }
    }
    c0[z] = a; c1[z] = b; par[a] = z; par[b] = z;
    alive[a] = alive[b] = 0; alive[z] = 1;
    return z;
  };
  if (method == TNB_TREES_RANDOM) {
    std::vector<int32_t> order(net.n_inds);
    for (int i = 0; i < net.n_inds; ++i) order[i] = i;
    while (nxt < N) {
      for (size_t i = order.size(); i > 1; --i) std::swap(order[i - 1], order[rng.below(uint32_t(i))]);
      bool progressed = false;
      for (int32_t i : order) {
        auto& h = cur[size_t(i)];
        if (h.size() < 2) continue;
        const uint32_t x = rng.below(uint32_t(h.size()));
        uint32_t y = rng.below(uint32_t(h.size() - 1));
        if (y >= x) ++y;
        merge(h[x], h[y]);
        progressed = true;
        if (nxt == N) break;
      }
      if (!progressed) return false;
    }
    return true;
  }
  struct Cand {
    double score;
    uint32_t tie;
    int32_t a, b;
    bool operator<(const Cand& o) const { return score != o.score ? score > o.score : tie > o.tie; }
  };
  std::priority_queue<Cand> pq;
  std::vector<int> kcache(N, 0);
  for (int t = 0; t < n; ++t) kcache[t] = popc_row(&bits[size_t(t) * W], W);
  auto sz = [](int k) { return std::ldexp(1.0, std::min(k, 1000)); };
  auto push = [&](int a, int b) {
    pq.push(Cand{sz(out_size(a, b)) - sz(kcache[a]) - sz(kcache[b]), uint32_t(rng.next() >> 32), a, b});
  };
  std::vector<int32_t> seen(N, -1);
  for (int t = 0; t < n; ++t) {  // every pair of leaves sharing an index, once
    for (int w = 0; w < W; ++w) {
      uint32_t v = bits[size_t(t) * W + w];
      while (v) {
        const int i = w * 32 + __builtin_ctz(v);
        v &= v - 1;
        for (int32_t o : cur[size_t(i)])
          if (o > t && seen[o] != t) { seen[o] = t; push(t, o); }
      }
    }
  }
  std::fill(seen.begin(), seen.end(), -1);
  while (nxt < N && !pq.empty()) {
    const Cand c = pq.top();
    pq.pop();
    if (!alive[c.a] || !alive[c.b]) continue;
    const int z = (c.tie & 1) ? merge(c.a, c.b) : merge(c.b, c.a);
    const uint32_t* bz = &bits[size_t(z) * W];
    kcache[z] = popc_row(bz, W);
    for (int w = 0; w < W; ++w) {
      uint32_t v = bz[w];
      while (v) {
        const int i = w * 32 + __builtin_ctz(v);
        v &= v - 1;
        for (int32_t o : cur[size_t(i)])
          if (o != z && seen[o] != z) { seen[o] = z; push(z, o); }
      }
    }
  }
  return nxt == N;
}

// One tree.  method 0: greedy on size(out)-size(a)-size(b) with random tie-breaks; 1: random edge order.
static bool one_tree(const Net& net, uint64_t seed, int method, int32_t* par, int32_t* c0, int32_t* c1) {
  const int n = net.n, N = 2 * n - 1, W = net.W;
  SplitMix rng(seed * 0x9E3779B97F4A7C15ull + 0x1234567ull + uint64_t(method));
  std::fill(par, par + N, -1);
  std::fill(c0, c0 + N, -1);
  std::fill(c1, c1 + N, -1);
  if (n == 1) return true;
  std::vector<uint32_t> bits(size_t(N) * W, 0u);
  std::memcpy(bits.data(), net.leaf_bits, sizeof(uint32_t) * size_t(n) * W);
  std::vector<int32_t> o0(net.own0), o1(net.own1);  // current cluster holding each index
  std::vector<uint8_t> alive(N, 0);
  std::fill(alive.begin(), alive.begin() + n, 1);
  int nxt = n;
  auto merge = [&](int a, int b) {
    const int z = nxt++;
    uint32_t* bz = &bits[size_t(z) * W];
    const uint32_t *ba = &bits[size_t(a) * W], *bb = &bits[size_t(b) * W];
    for (int w = 0; w < W; ++w) bz[w] = ba[w] ^ bb[w];
    for (int w = 0; w < W; ++w) {
      uint32_t v = bz[w];
      while (v) {
        const int i = w * 32 + __builtin_ctz(v);
        v &= v - 1;
        if (o0[i] == a || o0[i] == b) o0[i] = z;
        if (o1[i] == a || o1[i] == b) o1[i] = z;
      }
    }
    c0[z] = a; c1[z] = b; par[a] = z; par[b] = z;
    alive[a] = alive[b] = 0; alive[z] = 1;
    return z;
  };
  if (method == TNB_TREES_RANDOM) {
    std::vector<int32_t> edges;
    for (int i = 0; i < net.n_inds; ++i)
      if (net.own1[i] >= 0) edges.push_back(i);
    for (size_t i = edges.size(); i > 1; --i) std::swap(edges[i - 1], edges[rng.below(uint32_t(i))]);
    for (int32_t i : edges) {
      if (nxt == N) break;
      const int a = o0[i], b = o1[i];
      if (a == b || a < 0 || b < 0 || !alive[a] || !alive[b]) continue;  // contracted index: stale owners
      if (rng.next() & 1) merge(a, b); else merge(b, a);
    }
    return nxt == N;
  }
  struct Cand {
    double score;
    uint32_t tie;
    int32_t a, b;
    bool operator<(const Cand& o) const { return score != o.score ? score > o.score : tie > o.tie; }  // min-heap
  };
  std::priority_queue<Cand> pq;
  std::vector<int> kcache(N, 0);
  for (int t = 0; t < n; ++t) kcache[t] = popc_row(&bits[size_t(t) * W], W);
  auto sz = [](int k) { return std::ldexp(1.0, std::min(k, 1000)); };
  auto push = [&](int a, int b) {
    const uint32_t *ba = &bits[size_t(a) * W], *bb = &bits[size_t(b) * W];
    int ko = 0;
    for (int w = 0; w < W; ++w) ko += __builtin_popcount(ba[w] ^ bb[w]);
    pq.push(Cand{sz(ko) - sz(kcache[a]) - sz(kcache[b]), uint32_t(rng.next() >> 32), a, b});
  };
  for (int i = 0; i < net.n_inds; ++i)
    if (net.own1[i] >= 0 && net.own0[i] != net.own1[i]) push(net.own0[i], net.own1[i]);
  std::vector<int32_t> seen(N, -1);
  while (nxt < N && !pq.empty()) {
    const Cand c = pq.top();
    pq.pop();
    if (!alive[c.a] || !alive[c.b]) continue;
    const int z = (c.tie & 1) ? merge(c.a, c.b) : merge(c.b, c.a);
    const uint32_t* bz = &bits[size_t(z) * W];
    kcache[z] = popc_row(bz, W);
    for (int w = 0; w < W; ++w) {
      uint32_t v = bz[w];
      while (v) {
        const int i = w * 32 + __builtin_ctz(v);
        v &= v - 1;
        const int other = (o0[i] == z) ? o1[i] : o0[i];
        if (other >= 0 && other != z && seen[other] != z) {
          seen[other] = z;
          push(z, other);
        }
      }
    }
  }
  return nxt == N;
}

// Fenwick tree over list slots with "live" flags (prefix counts + k-th live slot).
struct Fenwick {
  int n;
  std::vector<int32_t> t;
  explicit Fenwick(int n_) : n(n_), t(n_ + 1, 0) {}
  void add(int i, int v) { for (++i; i <= n; i += i & -i) t[i] += v; }
  int prefix(int i) const { int s = 0; for (; i > 0; i -= i & -i) s += t[i]; return s; }  // live in [0,i)
  int kth(int k) const {  // slot of the k-th (0-based) live element
    int pos = 0, lg = 1;
    while ((lg << 1) <= n) lg <<= 1;
    for (int pw = lg; pw; pw >>= 1)
      if (pos + pw <= n && t[pos + pw] <= k) { pos += pw; k -= t[pos]; }
    return pos;
  }
};

}  // namespace tnb

using namespace tnb;

extern "C" {

int tnb_version(void) { return 100; }

void tnb_mt19937_stream(uint32_t seed, uint64_t n, uint32_t* out) {
  Mt19937 m;
  m.seed(seed);
  m.fill(out, n);
}

void tnb_mt19937_state(uint32_t seed, uint64_t n_draws, uint32_t* state624, int32_t* pos) {
  Mt19937 m;
  m.seed(seed);
  for (uint64_t i = 0; i < n_draws; ++i) (void)m.next();
  std::memcpy(state624, m.x, sizeof(m.x));
  *pos = m.p;
}

void tnb_mt19937_advance(uint32_t* state624, int32_t* pos, uint64_t n_draws) {
  Mt19937 m;
  std::memcpy(m.x, state624, sizeof(m.x));
  m.p = *pos;
  for (uint64_t i = 0; i < n_draws; ++i) (void)m.next();
  std::memcpy(state624, m.x, sizeof(m.x));
  *pos = m.p;
}

int tnb_random_trees(int n_leaves, int n_inds, const uint32_t* leaf_bits, int n_trees, const uint64_t* seeds,
                     int method, int n_threads, int32_t* parent, int32_t* child0, int32_t* child1) {
  return tnb_random_trees_out(n_leaves, n_inds, leaf_bits, nullptr, n_trees, seeds, method, n_threads, parent, child0,
                              child1);
}

int tnb_random_trees_out(int n_leaves, int n_inds, const uint32_t* leaf_bits, const uint32_t* output_bits, int n_trees,
                         const uint64_t* seeds, int method, int n_threads, int32_t* parent, int32_t* child0,
                         int32_t* child1) {
  if (n_leaves < 1 || n_inds < 0 || n_trees < 0 || !leaf_bits || !seeds) {
    set_global_error("tnb_random_trees: invalid arguments");
    return -1;
  }
  Net net;
  net.n = n_leaves; net.n_inds = n_inds; net.W = (n_inds + 31) / 32; net.leaf_bits = leaf_bits;
  std::string err;
  if (!build_net(net, output_bits, err)) { set_global_error("tnb_random_trees: " + err); return -2; }
  const int N = 2 * n_leaves - 1;
  if (n_threads <= 0) n_threads = int(std::max(1u, std::thread::hardware_concurrency()));
  n_threads = std::max(1, std::min(n_threads, n_trees));
  std::vector<int> ok(size_t(n_threads), 1);
  auto work = [&](int tid) {
    for (int t = tid; t < n_trees; t += n_threads)
      if (!(net.hyper ? one_tree_hyper : one_tree)(net, seeds[t], method, parent + size_t(t) * N,
                                                   child0 + size_t(t) * N, child1 + size_t(t) * N))
        ok[size_t(tid)] = 0;
  };
  if (n_threads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int i = 0; i < n_threads; ++i) th.emplace_back(work, i);
    for (auto& t : th) t.join();
  }
  for (int v : ok)
    if (!v) { set_global_error("tnb_random_trees: the network is not connected"); return -3; }
  return 0;
}

int tnb_tree_to_path(int n_leaves, int n_trees, const int32_t* child0, const int32_t* child1, int n_tensors,
                     const int32_t* tensors_pos, int32_t* path) {
  const int n = n_leaves, N = 2 * n - 1;
  if (!tensors_pos) n_tensors = n;
  if (n_tensors < n) { set_global_error("tnb_tree_to_path: n_tensors < n_leaves"); return -1; }
  std::vector<int32_t> stack, slot(N);
  std::vector<uint8_t> vis(N);
  for (int t = 0; t < n_trees; ++t) {
    const int32_t *a = child0 + size_t(t) * N, *b = child1 + size_t(t) * N;
    int32_t* out = path + size_t(t) * (n - 1) * 2;
    Fenwick fw(n_tensors + n - 1);
    for (int i = 0; i < n_tensors; ++i) fw.add(i, 1);
    for (int i = 0; i < n; ++i) {
      slot[i] = tensors_pos ? tensors_pos[i] : i;
      if (slot[i] < 0 || slot[i] >= n_tensors) { set_global_error("tnb_tree_to_path: bad tensors_pos"); return -1; }
    }
    std::fill(vis.begin(), vis.end(), 0);
    stack.assign(1, N - 1);
    int step = 0;
    while (!stack.empty()) {  // post-order, children[0] first (include/tnco/utils.hpp:35-52)
      const int pos = stack.back();
      if (pos < 0 || pos >= N) { set_global_error("tnb_tree_to_path: bad node id"); return -1; }
      if (vis[pos] || a[pos] < 0) {
        stack.pop_back();
        if (a[pos] >= 0) {
          if (step >= n - 1) { set_global_error("tnb_tree_to_path: not a tree"); return -1; }
          const int sx = slot[a[pos]], sy = slot[b[pos]];
          out[2 * step] = fw.prefix(sx);
          out[2 * step + 1] = fw.prefix(sy);
          fw.add(sx, -1);
          fw.add(sy, -1);
          slot[pos] = n_tensors + step;
          fw.add(slot[pos], 1);
          ++step;
        }
      } else {
        vis[pos] = 1;
        stack.push_back(b[pos]);
        stack.push_back(a[pos]);
      }
    }
    if (step != n - 1) { set_global_error("tnb_tree_to_path: not a full binary tree"); return -1; }
  }
  return 0;
}

int tnb_merge_paths(int n_tensors, int n_runs, int n_paths, const int32_t* lens, const int32_t* paths,
                    int32_t* merged) {
  if (n_tensors < 1 || n_runs < 0 || n_paths < 0 || (n_paths > 0 && (!lens || !paths)) || !merged) {
    set_global_error("tnb_merge_paths: invalid arguments");
    return -1;
  }
  int total = 0;
  for (int i = 0; i < n_paths; ++i) total += lens[i];
  if (total > n_tensors - 1) { set_global_error("tnb_merge_paths: too many contractions"); return -1; }
  std::vector<int32_t> l2m(size_t(n_tensors) + size_t(total));
  for (int r = 0; r < n_runs; ++r) {
    const int32_t* in = paths + size_t(r) * total * 2;
    int32_t* out = merged + size_t(r) * (n_tensors - 1) * 2;
    Fenwick mf(n_tensors + total);
    for (int i = 0; i < n_tensors; ++i) mf.add(i, 1);
    std::vector<uint8_t> mlive(size_t(n_tensors) + size_t(total), 0);
    std::fill(mlive.begin(), mlive.begin() + n_tensors, 1);
    int gstep = 0;
    for (int i = 0; i < n_paths; ++i) {
      Fenwick lf(n_tensors + lens[i]);
      for (int k = 0; k < n_tensors; ++k) { lf.add(k, 1); l2m[size_t(k)] = k; }
      int live = n_tensors;
      for (int k = 0; k < lens[i]; ++k, ++gstep) {
        int x = in[2 * gstep], y = in[2 * gstep + 1];
        if (x > y) std::swap(x, y);
        if (x < 0 || y >= live || x == y) { set_global_error("tnb_merge_paths: invalid path entry"); return -2; }
        const int sx = lf.kth(x), sy = lf.kth(y);
        const int mx = l2m[size_t(sx)], my = l2m[size_t(sy)];
        if (!mlive[size_t(mx)] || !mlive[size_t(my)]) {
          set_global_error("'paths' are not valid or not disconnected.");
          return -2;
        }
        int px = mf.prefix(mx), py = mf.prefix(my);
        if (px > py) std::swap(px, py);
        out[2 * gstep] = px;
        out[2 * gstep + 1] = py;
        mf.add(mx, -1); mf.add(my, -1);
        mlive[size_t(mx)] = mlive[size_t(my)] = 0;
        lf.add(sx, -1); lf.add(sy, -1);
        const int ls = n_tensors + k, ms = n_tensors + gstep;
        lf.add(ls, 1);
        mf.add(ms, 1);
        mlive[size_t(ms)] = 1;
        l2m[size_t(ls)] = ms;
        live -= 1;
      }
    }
    for (int k = gstep; k < n_tensors - 1; ++k) { out[2 * k] = 0; out[2 * k + 1] = 1; }  // autocomplete
  }
  return 0;
}

int tnb_path_to_tree(int n_leaves, const int32_t* path, int32_t* parent, int32_t* child0, int32_t* child1) {
  const int n = n_leaves, N = 2 * n - 1;
  std::fill(parent, parent + N, -1);
  std::fill(child0, child0 + N, -1);
  std::fill(child1, child1 + N, -1);
  Fenwick fw(N);
  std::vector<int32_t> node_of_slot(N);
  for (int i = 0; i < n; ++i) { fw.add(i, 1); node_of_slot[i] = i; }
  int live = n;
  for (int i = 0; i < n - 1; ++i) {
    int x = path[2 * i], y = path[2 * i + 1];
    if (x > y) std::swap(x, y);
    if (x < 0 || y >= live || x == y) { set_global_error("tnb_path_to_tree: invalid path entry"); return -1; }
    const int sy = fw.kth(y), sx = fw.kth(x);
    const int py = node_of_slot[sy], px = node_of_slot[sx];
    fw.add(sy, -1);
    fw.add(sx, -1);
    const int z = n + i;
    node_of_slot[z] = z;
    fw.add(z, 1);
    live -= 1;
    child0[z] = px; child1[z] = py; parent[px] = z; parent[py] = z;
  }
  return 0;
}

}  // extern "C"
