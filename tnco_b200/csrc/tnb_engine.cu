// The engine behind the C-ABI of include/tnco_b200.h: device memory, kernel launches, host-side MT19937
// stream feeding, result read-back.  Compiled by nvcc for sm_100a into libtnco_b200.so.
//
// (With -DTNB_EMU and a plain C++ compiler the same file builds tests/emu's logic-emulation harness: the
// "device" is host memory and a launch is a loop over chains with one lane per chain.  That build is test
// infrastructure; the tnco_b200 package never loads it.)
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "tnb_internal.h"
#include "tnb_launch.h"

namespace tnb {

#if !defined(TNB_EMU)
// one translation unit per tile shape (tnb_inst.cu)
#define TNB_EXTERN_SHAPE(T, W)                                                            \
  extern template bool launch_tw<T, W>(Rt&, const Params&, bool, bool, bool);             \
  extern template bool launch_treegen_t<T, W>(Rt&, const Params&);
TNB_EXTERN_SHAPE(4, 1)
TNB_EXTERN_SHAPE(8, 1)
TNB_EXTERN_SHAPE(16, 1)
TNB_EXTERN_SHAPE(32, 1)
TNB_EXTERN_SHAPE(32, 2)
TNB_EXTERN_SHAPE(32, 3)
TNB_EXTERN_SHAPE(32, 4)
#undef TNB_EXTERN_SHAPE
#endif

#if !defined(TNB_EMU)
// children words between the compact [n_chains][n_int] array (upload / read-back form) and the node records
static __global__ void __launch_bounds__(256) ch_scatter_kernel(const uint32_t* src, char* rec, int stride, size_t count) {
  const size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
  if (i < count) *reinterpret_cast<uint32_t*>(rec + i * size_t(stride)) = src[i];
}
static __global__ void __launch_bounds__(256) ch_gather_kernel(uint32_t* dst, const char* rec, int stride, size_t count) {
  const size_t i = size_t(blockIdx.x) * 256 + threadIdx.x;
  if (i < count) dst[i] = *reinterpret_cast<const uint32_t*>(rec + i * size_t(stride));
}

#endif

static bool launch(Rt& rt, const Params& P, int tile, int wpl, bool init, bool finite, bool stream_rng) {
#if defined(TNB_EMU)
  (void)tile;
  (void)wpl;
  if (P.W <= 4) return launch_tw<1, 4>(rt, P, init, finite, stream_rng);
  if (P.W <= 16) return launch_tw<1, 16>(rt, P, init, finite, stream_rng);
  if (P.W <= 48) return launch_tw<1, 48>(rt, P, init, finite, stream_rng);
  return launch_tw<1, 128>(rt, P, init, finite, stream_rng);
#else
  switch (tile * 16 + wpl) {
    case 4 * 16 + 1: return launch_tw<4, 1>(rt, P, init, finite, stream_rng);
    case 8 * 16 + 1: return launch_tw<8, 1>(rt, P, init, finite, stream_rng);
    case 16 * 16 + 1: return launch_tw<16, 1>(rt, P, init, finite, stream_rng);
    case 32 * 16 + 1: return launch_tw<32, 1>(rt, P, init, finite, stream_rng);
    case 32 * 16 + 2: return launch_tw<32, 2>(rt, P, init, finite, stream_rng);
    case 32 * 16 + 3: return launch_tw<32, 3>(rt, P, init, finite, stream_rng);
    case 32 * 16 + 4: return launch_tw<32, 4>(rt, P, init, finite, stream_rng);
  }
  rt.err = "unsupported tile shape";
  return false;
#endif
}

static bool ch_copy(Rt& rt, uint32_t* compact, char* rec, int stride, size_t count, bool to_records) {
#if defined(TNB_EMU)
  (void)rt;
  for (size_t i = 0; i < count; ++i) {
    uint32_t* r = reinterpret_cast<uint32_t*>(rec + i * size_t(stride));
    if (to_records) *r = compact[i]; else compact[i] = *r;
  }
  return true;
#else
  if (count == 0) return true;
  const unsigned grid = unsigned((count + 255) / 256);
  if (to_records) ch_scatter_kernel<<<grid, 256, 0, rt.stream>>>(compact, rec, stride, count);
  else ch_gather_kernel<<<grid, 256, 0, rt.stream>>>(compact, rec, stride, count);
  return rt.ok(cudaGetLastError(), "ch_copy launch");
#endif
}

static bool launch_treegen(Rt& rt, const Params& P, int tile, int wpl) {
#if defined(TNB_EMU)
  (void)tile;
  (void)wpl;
  if (P.W <= 4) return launch_treegen_t<1, 4>(rt, P);
  if (P.W <= 16) return launch_treegen_t<1, 16>(rt, P);
  if (P.W <= 48) return launch_treegen_t<1, 48>(rt, P);
  return launch_treegen_t<1, 128>(rt, P);
#else
  switch (tile * 16 + wpl) {
    case 4 * 16 + 1: return launch_treegen_t<4, 1>(rt, P);
    case 8 * 16 + 1: return launch_treegen_t<8, 1>(rt, P);
    case 16 * 16 + 1: return launch_treegen_t<16, 1>(rt, P);
    case 32 * 16 + 1: return launch_treegen_t<32, 1>(rt, P);
    case 32 * 16 + 2: return launch_treegen_t<32, 2>(rt, P);
    case 32 * 16 + 3: return launch_treegen_t<32, 3>(rt, P);
    case 32 * 16 + 4: return launch_treegen_t<32, 4>(rt, P);
  }
  rt.err = "unsupported tile shape";
  return false;
#endif
}

// ------------------------------------------------------------------------------------------ chain storage
struct ChainSet {
  int n_chains = 0;
  int16_t *par = nullptr, *bpar = nullptr;
  uint32_t *bch = nullptr, *slices = nullptr, *bslices = nullptr;
  char *rec = nullptr, *bitsb = nullptr;  // node headers / index sets, see Params::hdr (bitsb == rec + 16 when interleaved)
  char* bits_alloc = nullptr;             // SPLIT layout only: the separate index-set array
  // SPLIT layout only: par, the node headers and kwsz are carved out of ONE block -- everything the walk's
  // address-dependent loads touch (parent -> header -> sibling id), a tenth of the state -- so that one L2
  // access-policy window can keep it resident while the index sets stream through (Rt::l2_window)
  char* hot = nullptr;
  size_t hot_bytes = 0;
  int hstride = 0, bstride = 0;
  double* pc = nullptr;
  dbl2* cp2 = nullptr;
  double *total = nullptr, *min_total = nullptr, *out_seq = nullptr, *out_maxw = nullptr;
  unsigned long long *seeds = nullptr, *rng_ctr = nullptr, *n_prop = nullptr, *n_acc = nullptr, *n_wrej = nullptr,
                     *cursor = nullptr;
  long long* sweep_idx = nullptr;
  int* overrun = nullptr;
  uint16_t* nbig = nullptr;
  int16_t *posbuf = nullptr, *kpop = nullptr, *word = nullptr;
  uint32_t *wkey = nullptr, *kwsz = nullptr;
  int16_t* ksp = nullptr;
  int* tree_fail = nullptr;
  double* escore = nullptr;
  uint32_t* stream = nullptr;
  unsigned long long stream_len = 0;
  // decision trace (tnb_set_trace)
  TraceRec* trace = nullptr;
  unsigned long long *trace_n = nullptr, trace_cap = 0;
  uint32_t *trace_S = nullptr, *trace_sn = nullptr, trace_scap = 0;
  int trace_chains = 0;

  void release(Rt& rt) {
    if (hot) par = nullptr, rec = nullptr, kwsz = nullptr;  // (carved out of `hot`)
    void* ps[] = {hot, ksp, par, bpar, rec, bits_alloc, bch, pc, slices, bslices, cp2, total, min_total, out_seq, out_maxw, seeds,
                  rng_ctr, n_prop, n_acc, n_wrej, cursor, sweep_idx, overrun, nbig, posbuf, kpop, tree_fail, escore, stream, kwsz, word, wkey,
                  trace, trace_n, trace_S, trace_sn};
    for (void* p : ps) rt.free_(p);
    *this = ChainSet();
  }
};

}  // namespace tnb

using namespace tnb;

struct tnb_engine {
  Rt rt;
  std::string err;
  // network
  int n = 0, N = 0, n_int = 0, n_inds = 0, W = 0, Ws = 0, Npad = 0, stride = 0;
  int tile = 32, wpl = 1;
  uint64_t dim = 2;
  double log2d = 1.0;
  uint32_t* d_leaf_bits = nullptr;
  std::vector<double> h_betas;  // the schedule as given (tnb_set_betas)
  int inv_kind = 0;             // acceptance rule d_inv_betas was computed for
  double* d_pow_tab = nullptr;
  int16_t* d_net_own = nullptr;  // [2][n_inds] the leaves holding each index (device tree construction)
  uint16_t* d_hcount0 = nullptr;  // [Ws*32] initial hyper counts
  std::vector<uint16_t> h_holders;  // [n_inds] tensors holding each index
  std::vector<uint32_t> h_output;   // [W] output (open) indices
  bool hyper = false;
  // Per-index dimensions that are powers of two: index i of dimension 2^w is carried as w adjacent binary
  // ("virtual") indices [voff[i], voff[i] + vw[i]) that are always set together, so that cost = 2^popcount and
  // width = popcount keep holding and the sweep kernels are untouched; only the slicers know about the groups.
  // n_inds / W / Ws count virtual indices; n_inds_u / Wu are the caller's.  Without per-index dims they coincide.
  int n_inds_u = 0, Wu = 0;
  bool grouped = false;
  std::vector<int> voff, vw;
  uint32_t* d_leader = nullptr;  // [Ws] first virtual bit of every index
  // sparse-index cost model (tnb_set_sparse_inds): sparse indices in the virtual index space, or none
  uint32_t* d_sparse = nullptr;  // [Ws + tail]
  uint64_t n_projs = 0;
  uint32_t* d_skip = nullptr;    // [Ws + tail] skip_slices (tnb_set_skip_slices)
  // general per-index dimensions (not all equal, not all powers of two): the reference's sequential cost / width loops
  bool generic = false;
  double* d_gdims = nullptr;     // [Ws*32]
  double* d_glog2 = nullptr;     // [Ws*32]
  // skip_slices under the production generator: only the reference's slicer (get_slices_dev) knows them, so such a
  // batch takes the table-cost kernels, which re-slice with it.  (Teaching the production re-slicer to skip cost the
  // 2^popcount kernel 1.7 % on C4 -- ptxas allocates registers across the call, so a change behind it moves the level
  // loop -- for an option the app never sets.)
  bool skip_verbatim() const { return d_skip != nullptr && finite && rng_kind == TNB_RNG_PHILOX; }
  // costs are 2^popcount (uniform dimension 2 or power-of-two groups, simple cost model): DIM2 kernels, fast re-slicer
  bool pow2_costs() const { return dim == 2 && !d_sparse && !generic && !skip_verbatim(); }
  // The production re-slicer applies where a width is made of popcounts: any uniform dimension or power-of-two groups,
  // and the sparse-index model (two popcounts per node) unless it comes with groups; not general dimensions.
  bool popcount_widths() const { return !generic && !(d_sparse && grouped) && !skip_verbatim(); }
  uint8_t* d_gw = nullptr;       // [Ws*32] log2(dim) at the leader positions

  // caller's index space <-> virtual index space (rows of Wu / W words)
  void expand_row(const uint32_t* u, uint32_t* v) const {
    for (int w = 0; w < W; ++w) v[w] = 0u;
    for (int i = 0; i < n_inds_u; ++i)
      if ((u[i >> 5] >> (i & 31)) & 1u)
        for (int k = voff[size_t(i)]; k < voff[size_t(i)] + vw[size_t(i)]; ++k) v[k >> 5] |= 1u << (k & 31);
  }
  void contract_row(const uint32_t* v, uint32_t* u) const {
    for (int w = 0; w < Wu; ++w) u[w] = 0u;
    for (int i = 0; i < n_inds_u; ++i) {
      const int k = voff[size_t(i)];
      if ((v[k >> 5] >> (k & 31)) & 1u) u[i >> 5] |= 1u << (i & 31);
    }
  }
  std::vector<uint32_t> h_leaf_bits;  // [n][W]
  // mode
  bool finite = false;
  float max_width = 0.f;
  int every = 0, dsi = 0, prob_kind = TNB_PROB_MH, rng_kind = TNB_RNG_PHILOX, layout = TNB_LAYOUT_AUTO;
  // L2 carve-out for the hot block of a SPLIT batch in MiB (Rt::l2_window): 0 = off, < 0 = as much as the device
  // allows.  TNB_L2_PERSIST_MB overrides it (measurement switch).
  int l2_persist_mb = [] {
    const char* f = std::getenv("TNB_L2_PERSIST_MB");
    return f ? std::atoi(f) : 0;
  }();
  int max_new = 0;  // max_number_new_slices (tnb_set_new_slices)
  // chains
  ChainSet cs;
  bool initialized = false;
  uint64_t chain_id0 = 0;
  std::vector<uint64_t> h_seeds;
  // schedule
  double* d_betas = nullptr;
  float* d_inv_betas = nullptr;
  int64_t n_betas = 0;
  // MT19937 feeding
  std::vector<Mt19937> mts;
  std::vector<uint32_t> h_stream;
  std::vector<uint64_t> words_base;  // draws consumed before the current stream window, per chain
  // resume (tnb_set_resume): saved generator states, slices and best trees that replace what the constructors derive
  std::vector<uint32_t> rs_mt;                  // [n_chains][625]
  std::vector<uint32_t> rs_slices, rs_bslices;  // [n_chains][Wu], caller's index space
  std::vector<int32_t> rs_bp, rs_ba, rs_bb;     // [n_chains][N]
  void clear_resume() { rs_mt.clear(); rs_slices.clear(); rs_bslices.clear(); rs_bp.clear(); rs_ba.clear(); rs_bb.clear(); }
  void* d_flush = nullptr;
  // timing
  double kernel_ms = 0.0;
  int64_t launches = 0;

  bool fail(const std::string& m) { err = m; return false; }
  bool rtfail() { err = rt.err; return false; }
};

namespace tnb {

// Lanes per chain.  The memory layout does not depend on it (word w of an index set belongs to lane w), so it is
// chosen per batch: sharing a warp between chains saves instructions but costs divergence and parallelism, and the
// measurements reduce to one rule -- take the widest tile whose batch still fits one wave of
// resident warps (28 per SM); if even the narrowest tile does not fit, take the narrowest.
static int pick_tile(int W, int n_inds, int& wpl, bool hyper, long long n_chains, int n_sms) {
  wpl = 1;
  int tile = 32;
  // (one-word-per-lane kernels pack popcounts into 10-bit fields: 1024 indices exactly take two words per lane)
  if (n_inds >= 1024 && W <= 32) { wpl = 2; return 32; }
  if (hyper || W > 32) { wpl = (W + 31) / 32; return 32; }
  const int narrowest = W <= 4 ? 4 : W <= 8 ? 8 : W <= 16 ? 16 : 32;
  const long long capacity = 28ll * (n_sms > 0 ? n_sms : 148);
  tile = narrowest;
  for (int t = 32; t >= narrowest; t >>= 1)
    if ((n_chains * t + 31) / 32 <= capacity) { tile = t; break; }
  if (const char* f = std::getenv("TNB_TILE")) {
    const int t = std::atoi(f);
    if ((t == 4 || t == 8 || t == 16 || t == 32) && t * 1 >= (W <= 32 ? W : 32)) tile = t;
  }
  return tile;
}

static unsigned long long sweep_reserve(const tnb_engine* e) {
  // worst case draws of one sweep: leaf + (coin + 2) per level, depth <= n-1; plus slicer head-room
  unsigned long long r = 1ull + 3ull * (unsigned long long)(e->n > 1 ? e->n - 1 : 1);
  if (e->finite) r += 8ull * (unsigned long long)e->n_inds + 64ull;
  if (e->finite && e->max_new > 0) r += (unsigned long long)e->max_new * (unsigned long long)(e->n > 1 ? e->n - 1 : 1);
  return r;
}

static void fill_params(const tnb_engine* e, const ChainSet& cs, Params& P) {
  std::memset(&P, 0, sizeof(P));
  P.n = e->n; P.N = e->N; P.n_int = e->n_int; P.n_inds = e->n_inds; P.W = e->W; P.Ws = e->Ws;
  P.leaf_bits = e->d_leaf_bits; P.pow_tab = e->d_pow_tab; P.log2d = e->log2d;
  P.skip = e->d_skip;
  P.sparse = e->d_sparse;  // costs of a network with sparse indices come from the table kernels
  P.dim2 = e->pow2_costs();
  // 2^popcount production kernels keep the high word of every contraction cost at header +4 (Params::cost_hi)
  P.cost_hi = e->rng_kind == TNB_RNG_PHILOX && e->pow2_costs();
  P.gdims = e->generic ? e->d_gdims : nullptr;
  P.glog2 = e->generic ? e->d_glog2 : nullptr;
  P.n_projs = double(e->n_projs);
  P.log2_n_projs = e->n_projs ? std::log2(double(e->n_projs)) : 0.0;
  P.max_new = e->max_new;
  P.finite = e->finite; P.every = e->every; P.dsi = e->dsi; P.prob_kind = e->prob_kind; P.max_width = e->max_width;
  P.n_chains = cs.n_chains; P.Npad = e->Npad;
  P.par = cs.par; P.hdr = cs.rec; P.bitsb = cs.bitsb; P.hstride = cs.hstride; P.bstride = cs.bstride; P.pc = cs.pc; P.bpar = cs.bpar; P.bch = cs.bch;
  P.slices = cs.slices; P.bslices = cs.bslices; P.total = cs.total; P.min_total = cs.min_total;
  P.seeds = cs.seeds; P.rng_ctr = cs.rng_ctr; P.chain_id0 = e->chain_id0; P.sweep_idx = cs.sweep_idx;
  P.n_prop = cs.n_prop; P.n_acc = cs.n_acc; P.n_wrej = cs.n_wrej;
  P.stream = cs.stream; P.cursor = cs.cursor; P.stream_len = cs.stream_len; P.reserve = sweep_reserve(e);
  P.overrun = cs.overrun;
  P.betas = e->d_betas; P.inv_betas = e->d_inv_betas; P.n_betas = e->n_betas; P.until = 0;
  P.nbig = cs.nbig; P.posbuf = cs.posbuf; P.cp2 = cs.cp2;
  P.slices_given = 0; P.out_seq = cs.out_seq; P.out_maxw = cs.out_maxw;
  P.kwsz = cs.kwsz; P.ksp = cs.ksp; P.word = cs.word; P.wkey = cs.wkey;
  P.kthr = 0;
  if (e->finite)
    for (int k = 0; k <= e->n_inds; ++k)
      if (float(e->log2d * double(k)) <= e->max_width) P.kthr = k;
  P.grouped = e->grouped; P.leader = e->d_leader; P.gw = e->d_gw;
  P.hyper = e->hyper; P.hyp_off = 4 * e->Ws; P.hcount0 = e->d_hcount0;
  // shared-memory-resident chains (TNB_LAYOUT_SMEM): unconstrained production kernels whose state fits the SM
  P.smem_chain_bytes = 0;
  if (e->layout == TNB_LAYOUT_SMEM && !e->finite && e->rng_kind == TNB_RNG_PHILOX && e->pow2_costs() && !e->hyper &&
      e->wpl == 1 && !cs.bits_alloc) {
    const size_t b = smem_chain_bytes(e->Npad, e->n_int, e->stride);
    if (b * size_t(32 / e->tile) <= size_t(200) << 10) P.smem_chain_bytes = int(b);
  }
  P.trace = cs.trace; P.trace_n = cs.trace_n; P.trace_cap = cs.trace_cap; P.trace_S = cs.trace_S; P.trace_sn = cs.trace_sn;
  P.trace_scap = cs.trace_scap; P.trace_chains = cs.trace_chains;
  P.net_own = e->d_net_own; P.kpop = cs.kpop; P.escore = cs.escore; P.tree_fail = cs.tree_fail; P.tree_method = TNB_TREES_GREEDY;
}

template <class T>
static bool alloc_to(Rt& rt, T*& p, size_t count) {
  p = static_cast<T*>(rt.alloc(count * sizeof(T)));
  return p != nullptr;
}

static bool alloc_chains(tnb_engine* e, ChainSet& cs, int n_chains, bool with_best, bool with_slicer) {
  Rt& rt = e->rt;
  if (&cs == &e->cs) e->tile = pick_tile(e->W, e->n_inds, e->wpl, e->hyper, n_chains, rt.n_sms);
  cs.n_chains = n_chains;
  const size_t nc = size_t(n_chains), ni = size_t(std::max(e->n_int, 1));
  // layout: interleaved node records while the whole batch fits L2, split beyond
  const size_t state = nc * ni * size_t(e->stride);
  bool split = state > (size_t(96) << 20);
  if (e->layout == TNB_LAYOUT_INTERLEAVED || e->layout == TNB_LAYOUT_SMEM) split = false;
  if (e->layout == TNB_LAYOUT_SPLIT) split = true;
  cs.hstride = split ? 16 : e->stride;
  // (split rows start on 32-byte sector boundaries: a 188-byte row then costs 6 sectors, not 6 or 7, and one
  //  prefetch per 128-byte line covers it)
  cs.bstride = split ? (4 * e->Ws * (e->hyper ? 2 : 1) + 31) / 32 * 32 : e->stride;
  // (+ tail: load_bits reads whole tiles of words and masks the ones beyond the row)
  // per-node popcounts / leaf counts for the production re-slicer: wherever a width is a popcount times a constant
  // (any uniform dimension, power-of-two groups; no sparse indices, no general dimensions) -- otherwise the
  // table-cost kernels re-slice with the reference's slicer verbatim
  // (TNB_VERBATIM_RESLICER: measurement switch, keeps the table-cost kernels on the reference's slicer)
  const bool want_kwsz = with_slicer && e->finite && e->popcount_widths() &&
                         (e->pow2_costs() || !std::getenv("TNB_VERBATIM_RESLICER"));
  bool ok = true;
  if (split) {
    auto up = [](size_t b) { return (b + 255) / 256 * 256; };
    const size_t par_b = up(nc * e->Npad * sizeof(int16_t)), rec_b = up(nc * ni * 16 + 16 * kTailWords),
                 kw_b = want_kwsz ? up(nc * e->Npad * sizeof(uint32_t)) : 0;
    cs.hot_bytes = par_b + rec_b + kw_b;
    ok = alloc_to(rt, cs.hot, cs.hot_bytes);
    if (ok) {
      cs.rec = cs.hot;  // (headers first: 8-byte accesses on a 256-byte aligned base)
      cs.par = reinterpret_cast<int16_t*>(cs.hot + rec_b);
      if (want_kwsz) cs.kwsz = reinterpret_cast<uint32_t*>(cs.hot + rec_b + par_b);
    }
  } else {
    ok = alloc_to(rt, cs.par, nc * e->Npad) && alloc_to(rt, cs.rec, nc * ni * size_t(cs.hstride) + 4 * kTailWords);
  }
  ok = ok && (!split || alloc_to(rt, cs.bits_alloc, nc * ni * size_t(cs.bstride) + 4 * kTailWords)) &&
            alloc_to(rt, cs.pc, nc * ni) && alloc_to(rt, cs.bch, nc * ni) &&
            alloc_to(rt, cs.slices, nc * e->Ws) && alloc_to(rt, cs.total, nc) && alloc_to(rt, cs.min_total, nc) &&
            alloc_to(rt, cs.out_seq, nc) && alloc_to(rt, cs.out_maxw, nc) && alloc_to(rt, cs.seeds, nc) &&
            alloc_to(rt, cs.rng_ctr, nc) && alloc_to(rt, cs.n_prop, nc) && alloc_to(rt, cs.n_acc, nc) &&
            alloc_to(rt, cs.n_wrej, nc) && alloc_to(rt, cs.cursor, nc) && alloc_to(rt, cs.sweep_idx, nc) &&
            alloc_to(rt, cs.overrun, nc);
  cs.bitsb = split ? cs.bits_alloc : cs.rec + 16;
  if (ok && with_best) ok = alloc_to(rt, cs.bpar, nc * e->Npad) && alloc_to(rt, cs.bslices, nc * e->Ws);
  if (ok && with_slicer)
    ok = alloc_to(rt, cs.nbig, nc * e->Ws * 32) && alloc_to(rt, cs.posbuf, nc * e->Ws * 32) &&
         alloc_to(rt, cs.cp2, nc * ni);
  else if (ok && e->hyper)
    ok = alloc_to(rt, cs.nbig, nc * e->Ws * 32);  // hyper counters of build_sets
  if (ok && want_kwsz)
    ok = (cs.kwsz || alloc_to(rt, cs.kwsz, nc * e->Npad)) && (!e->d_sparse || alloc_to(rt, cs.ksp, nc * e->Npad)) &&
         alloc_to(rt, cs.word, nc * e->Npad) && alloc_to(rt, cs.wkey, nc * e->Npad);
  if (!ok) return e->rtfail();
  return true;
}

// allocate a chain set and upload the packed topology
// validate trees in the reference's node order and pack them: parents [n_chains][Npad] int16, children words
// child0 | child1 << 16 per internal node [n_chains][max(n_int, 1)]
static bool pack_trees(tnb_engine* e, int n_chains, const int32_t* parent, const int32_t* c0, const int32_t* c1,
                       std::vector<int16_t>& hp, std::vector<uint32_t>& hc) {
  const size_t nc = size_t(n_chains), ni = size_t(std::max(e->n_int, 1));
  const int N = e->N, n = e->n;
  hp.assign(nc * e->Npad, int16_t(-1));
  hc.assign(nc * ni, 0u);
  std::vector<int> seen(size_t(N), 0);
  for (int c = 0; c < n_chains; ++c) {
    const int32_t *p = parent + size_t(c) * N, *a = c0 + size_t(c) * N, *b = c1 + size_t(c) * N;
    std::fill(seen.begin(), seen.end(), 0);
    for (int z = 0; z < N; ++z) {
      const bool leaf = z < n;
      if (leaf ? (a[z] != -1 || b[z] != -1) : (a[z] < 0 || a[z] >= N || b[z] < 0 || b[z] >= N || a[z] == b[z]))
        return e->fail("invalid tree: leaves must come first and internal nodes need two children (chain " +
                       std::to_string(c) + ", node " + std::to_string(z) + ")");
      if (z == N - 1 ? p[z] != -1 : (p[z] < n || p[z] >= N))
        return e->fail("invalid tree: bad parent / root must be the last node (chain " + std::to_string(c) + ")");
      if (!leaf) {
        if (p[a[z]] != z || p[b[z]] != z) return e->fail("invalid tree: parent/children mismatch");
        seen[a[z]]++; seen[b[z]]++;
        hc[size_t(c) * ni + (z - n)] = uint32_t(a[z]) | (uint32_t(b[z]) << 16);
      }
      hp[size_t(c) * e->Npad + z] = int16_t(p[z]);
    }
    for (int z = 0; z < N - 1; ++z)
      if (seen[z] != 1) return e->fail("invalid tree: every non-root node must be a child exactly once");
  }
  return true;
}

static bool make_chains(tnb_engine* e, ChainSet& cs, int n_chains, const int32_t* parent, const int32_t* c0,
                        const int32_t* c1, bool with_best, bool with_slicer) {
  if (!alloc_chains(e, cs, n_chains, with_best, with_slicer)) return false;
  Rt& rt = e->rt;
  const size_t nc = size_t(n_chains);
  std::vector<int16_t> hp;
  std::vector<uint32_t> hc;
  if (!pack_trees(e, n_chains, parent, c0, c1, hp, hc)) return false;
  // the compact children words travel through bch (the init kernel overwrites it with the first snapshot)
  if (!rt.h2d(cs.par, hp.data(), hp.size() * sizeof(int16_t)) || !rt.h2d(cs.bch, hc.data(), hc.size() * sizeof(uint32_t)) ||
      !ch_copy(rt, cs.bch, cs.rec, cs.hstride, nc * size_t(e->n_int), true))
    return e->rtfail();
  if (!rt.sync()) return e->rtfail();  // hp/hc go out of scope
  return true;
}

static bool check_shared(tnb_engine* e, const int32_t* c0, const int32_t* c1, int n_chains) {
  // check_shared_inds (include/tnco/ctree.hpp:101-152): children of every contraction must share an index.
  // inds(z) by the hyper-count rule of tnco/ctree.py:169-189, computed here on the host for validation only.
  if (e->dsi) return true;
  const int N = e->N, n = e->n, W = e->W;
  std::vector<uint32_t> bits(size_t(N) * W);
  std::vector<int> cnt(size_t(W) * 32, 0);
  std::vector<int32_t> order;
  std::vector<int32_t> stack;
  std::vector<uint8_t> vis(N);
  for (int c = 0; c < n_chains; ++c) {
    const int32_t *a = c0 + size_t(c) * N, *b = c1 + size_t(c) * N;
    std::memcpy(bits.data(), e->h_leaf_bits.data(), sizeof(uint32_t) * size_t(n) * W);
    std::fill(vis.begin(), vis.end(), 0);
    for (int i = 0; i < e->n_inds; ++i)
      cnt[size_t(i)] = int(e->h_holders[size_t(i)]) - 1 + int((e->h_output[size_t(i) >> 5] >> (i & 31)) & 1u);
    stack.assign(1, N - 1);
    while (!stack.empty()) {
      const int z = stack.back();
      if (z < n || vis[z]) {
        stack.pop_back();
        if (z >= n) {
          bool inter = false;
          for (int w = 0; w < W; ++w) {
            const uint32_t x = bits[size_t(a[z]) * W + w], y = bits[size_t(b[z]) * W + w];
            inter |= (x & y) != 0;
            uint32_t keep = 0u, v = x & y;
            while (v) {
              const int bit = __builtin_ctz(v);
              v &= v - 1;
              if (--cnt[size_t(w) * 32 + bit] > 0) keep |= 1u << bit;
            }
            bits[size_t(z) * W + w] = (x ^ y) | keep;
          }
          if (!inter)
            return e->fail("invalid tree: contracted tensors share no index (check_shared_inds), chain " +
                           std::to_string(c) + " node " + std::to_string(z));
        }
      } else {
        vis[z] = 1;
        stack.push_back(b[z]);
        stack.push_back(a[z]);
      }
    }
  }
  return true;
}

static bool stream_mode(const tnb_engine* e) { return e->rng_kind != TNB_RNG_PHILOX; }

// MT19937: (re)fill every chain's window of the draw stream from the host generators
static bool mt_refill(tnb_engine* e, bool first) {
  ChainSet& cs = e->cs;
  const size_t L = cs.stream_len, nc = size_t(cs.n_chains);
  std::vector<unsigned long long> cur(nc, 0);
  if (!first && !e->rt.d2h(cur.data(), cs.cursor, nc * sizeof(unsigned long long))) return e->rtfail();
  for (size_t c = 0; c < nc; ++c) {
    uint32_t* w = e->h_stream.data() + c * L;
    const size_t used = first ? L : size_t(cur[c]);
    if (!first) e->words_base[c] += used;
    if (!first && used < L) std::memmove(w, w + used, (L - used) * sizeof(uint32_t));
    e->mts[c].fill(w + (L - used), used);
  }
  if (!e->rt.h2d(cs.stream, e->h_stream.data(), nc * L * sizeof(uint32_t))) return e->rtfail();
  if (!e->rt.zero(cs.cursor, nc * sizeof(unsigned long long))) return e->rtfail();
  return e->rt.sync() || e->rtfail();
}

// Philox kernels are compiled for the app's mode (Metropolis-Hastings, shared-index moves)
static bool mode_ok(tnb_engine* e) {
  // (greedy / always acceptance are limits of the production kernel's threshold test -- 1/beta = 0 / +inf, see
  //  upload_inv_betas -- so only disable_shared_inds is out of its reach)
  if (e->rng_kind == TNB_RNG_PHILOX && e->dsi)
    return e->fail("TNB_RNG_PHILOX runs Metropolis-Hastings / greedy / always acceptance with shared-index moves only: "
                   "disable_shared_inds needs TNB_RNG_MT19937 or TNB_RNG_REPLAY (invalid mode)");
  if (e->rng_kind == TNB_RNG_PHILOX && e->finite && e->max_new > 0)
    return e->fail("max_number_new_slices > 0 is a core-object option: use TNB_RNG_MT19937 or TNB_RNG_REPLAY "
                   "(invalid mode)");
  return true;
}

static bool ensure_init(tnb_engine* e) {
  if (e->initialized) return true;
  if (e->cs.n_chains == 0) return e->fail("no chains: call tnb_set_chains first");
  if (e->rng_kind == TNB_RNG_REPLAY && !e->cs.stream) return e->fail("TNB_RNG_REPLAY needs tnb_set_stream");
  if (!mode_ok(e)) return false;
  if (e->rng_kind == TNB_RNG_MT19937) {
    ChainSet& cs = e->cs;
    const unsigned long long res = sweep_reserve(e);
    unsigned long long L = (256ull << 20) / 4ull / (unsigned long long)cs.n_chains;
    L = std::max(L, 4ull * res);
    L = std::min(L, std::max(1ull << 22, 4ull * res));
    cs.stream_len = L;
    e->rt.free_(cs.stream);
    if (!alloc_to(e->rt, cs.stream, size_t(cs.n_chains) * L)) return e->rtfail();
    e->h_stream.assign(size_t(cs.n_chains) * L, 0u);
    e->mts.resize(size_t(cs.n_chains));
    e->words_base.assign(size_t(cs.n_chains), 0);
    for (int c = 0; c < cs.n_chains; ++c) {
      Mt19937& m = e->mts[size_t(c)];
      if (e->rs_mt.empty()) {
        m.seed(uint32_t(e->h_seeds[size_t(c)]));
      } else {  // resume: `iss >> prng` (optimize/optimizer.hpp:68-72)
        std::memcpy(m.x, &e->rs_mt[size_t(c) * 625], sizeof(m.x));
        m.p = int(e->rs_mt[size_t(c) * 625 + 624]);
      }
    }
    if (!mt_refill(e, true)) return false;
  }
  const size_t nc_ = size_t(e->cs.n_chains);
  Params P;
  fill_params(e, e->cs, P);
  if (!e->rs_slices.empty() && e->finite) {  // `slices=` given: the constructor does not run the slicer (:80-88)
    std::vector<uint32_t> hs(nc_ * e->Ws, 0u);
    for (size_t c = 0; c < nc_; ++c) e->expand_row(&e->rs_slices[c * e->Wu], &hs[c * e->Ws]);
    if (!e->rt.h2d(e->cs.slices, hs.data(), hs.size() * sizeof(uint32_t)) || !e->rt.sync()) return e->rtfail();
    P.slices_given = 1;
  }
  if (!launch(e->rt, P, e->tile, e->wpl, true, e->finite, stream_mode(e))) return e->rtfail();
  if (!e->rt.sync()) return e->rtfail();
  if (!e->rs_bp.empty()) {
    // saved min_ctree [+ min_slices]: min_total_cost = get_cost(min_ctree[, min_slices]) as the reference
    // constructors compute it (infinite_memory/optimizer.hpp:72-75, finite_width/greedy/optimizer.hpp:99-101)
    std::vector<int16_t> hp;
    std::vector<uint32_t> hc;
    if (!pack_trees(e, e->cs.n_chains, e->rs_bp.data(), e->rs_ba.data(), e->rs_bb.data(), hp, hc)) return false;
    std::vector<double> seq(nc_);
    const uint32_t* bs = !e->finite ? nullptr : !e->rs_bslices.empty() ? e->rs_bslices.data()
                         : !e->rs_slices.empty() ? e->rs_slices.data() : nullptr;
    std::vector<uint32_t> cur;
    if (e->finite && !bs) {  // min_slices defaults to the constructor's slices (:89-90)
      std::vector<uint32_t> v(nc_ * e->Ws);
      if (!e->rt.d2h(v.data(), e->cs.slices, v.size() * sizeof(uint32_t))) return e->rtfail();
      cur.assign(nc_ * e->Wu, 0u);
      for (size_t c = 0; c < nc_; ++c) e->contract_row(&v[c * e->Ws], &cur[c * e->Wu]);
      bs = cur.data();
    }
    if (tnb_eval_cost(e, e->cs.n_chains, e->rs_bp.data(), e->rs_ba.data(), e->rs_bb.data(), bs, seq.data(), nullptr,
                      nullptr) != 0)
      return false;
    if (!e->rt.h2d(e->cs.bpar, hp.data(), hp.size() * sizeof(int16_t)) ||
        !e->rt.h2d(e->cs.bch, hc.data(), hc.size() * sizeof(uint32_t)) ||
        !e->rt.h2d(e->cs.min_total, seq.data(), seq.size() * sizeof(double)))
      return e->rtfail();
    if (e->finite) {
      std::vector<uint32_t> hb(nc_ * e->Ws, 0u);
      for (size_t c = 0; c < nc_; ++c) e->expand_row(bs + c * e->Wu, &hb[c * e->Ws]);
      if (!e->rt.h2d(e->cs.bslices, hb.data(), hb.size() * sizeof(uint32_t))) return e->rtfail();
    }
    if (!e->rt.sync()) return e->rtfail();
  }
  // "Precision is too low." (infinite_memory/optimizer.hpp:77-87)
  std::vector<double> tot(size_t(e->cs.n_chains));
  if (!e->rt.d2h(tot.data(), e->cs.total, tot.size() * sizeof(double))) return e->rtfail();
  if (e->n_int > 0)
    for (double t : tot)
      if (!(t > 0.0) || std::isinf(t) || std::isnan(t)) return e->fail("Precision is too low.");
  e->initialized = true;
  return true;
}

// hyper counts (holders - 1, +1 for output indices), the hyper flag and everything that depends on it
static bool refresh_hyper(tnb_engine* e) {
  std::vector<uint16_t> hc(size_t(e->Ws) * 32, 0);
  e->hyper = false;
  for (int i = 0; i < e->n_inds; ++i) {
    const int c = int(e->h_holders[size_t(i)]) - 1 + int((e->h_output[size_t(i) >> 5] >> (i & 31)) & 1u);
    hc[size_t(i)] = uint16_t(c < 0 ? 0 : c);
    e->hyper |= c >= 2;
  }
  e->tile = pick_tile(e->W, e->n_inds, e->wpl, e->hyper, 0, e->rt.n_sms);
  e->stride = 16 + 4 * e->Ws * (e->hyper ? 2 : 1);
  if (!e->rt.h2d(e->d_hcount0, hc.data(), hc.size() * sizeof(uint16_t)) || !e->rt.sync()) return e->rtfail();
  return true;
}

}  // namespace tnb

// ============================================================================================ C-ABI
extern "C" {

const char* tnb_last_error(const tnb_engine* e) { return e ? e->err.c_str() : global_error(); }

int tnb_create(tnb_engine** out, int device) {
  if (!out) { set_global_error("tnb_create: null output pointer"); return -1; }
  *out = nullptr;
  tnb_engine* e = new tnb_engine();
  if (!e->rt.init(device)) {
    set_global_error("tnb_create: " + e->rt.err);
    delete e;
    return -2;
  }
  *out = e;
  return 0;
}

void tnb_destroy(tnb_engine* e) {
  if (!e) return;
  e->cs.release(e->rt);
  e->rt.free_(e->d_leaf_bits);
  e->rt.free_(e->d_pow_tab);
  e->rt.free_(e->d_net_own);
  e->rt.free_(e->d_hcount0);
  e->rt.free_(e->d_leader);
  e->rt.free_(e->d_gw);
  e->rt.free_(e->d_sparse);
  e->rt.free_(e->d_skip);
  e->rt.free_(e->d_gdims);
  e->rt.free_(e->d_glog2);
  e->rt.free_(e->d_betas);
  e->rt.free_(e->d_inv_betas);
  e->rt.free_(e->d_flush);
  e->rt.destroy();
  delete e;
}

int tnb_set_network(tnb_engine* e, int n_leaves, int n_inds_u, const uint32_t* leaf_bits_u, uint64_t dim,
                    const uint64_t* dims) {
  if (!e) return -1;
  if (n_leaves < 1 || n_inds_u < 1 || !leaf_bits_u) return e->fail("tnb_set_network: invalid arguments"), -1;
  const int Wu = (n_inds_u + 31) / 32;
  // per-index dims: uniform -> scalar dim (include/tnco/ctree.hpp:80-89); powers of two -> virtual binary indices
  std::vector<int> vw(size_t(n_inds_u), 1), voff(size_t(n_inds_u), 0);
  bool grouped = false, generic = false;
  if (dims) {
    bool uniform = true, pow2 = true;
    for (int i = 0; i < n_inds_u; ++i) {
      const uint64_t d = dims[i];
      if (d < 1) return e->fail("tnb_set_network: every dimension must be positive (invalid dims)"), -1;
      uniform &= d == dims[0];
      pow2 &= d >= 2 && (d & (d - 1)) == 0;
    }
    if (uniform) {
      dim = dims[0];
    } else if (pow2) {
      for (int i = 0; i < n_inds_u; ++i) vw[size_t(i)] = __builtin_ctzll(dims[i]);
      grouped = true;
      dim = 2;
    } else {
      // anything else: one bit per index, costs / widths by the reference's sequential loops (table-cost kernels)
      generic = true;
      dim = 2;  // (only feeds the unused power table)
    }
  }
  if (dim < 1) return e->fail("tnb_set_network: dim must be positive"), -1;
  if (2 * n_leaves - 1 > 32767) return e->fail("tnb_set_network: at most 16384 tensors"), -2;
  int n_inds = 0;
  for (int i = 0; i < n_inds_u; ++i) {
    voff[size_t(i)] = n_inds;
    n_inds += vw[size_t(i)];
  }
  const int W = (n_inds + 31) / 32;
  if (W > 128) return e->fail("tnb_set_network: at most 4096 (binary) indices"), -2;
  for (int t = 0; t < n_leaves; ++t)
    for (int i = n_inds_u; i < Wu * 32; ++i)
      if ((leaf_bits_u[size_t(t) * Wu + (i >> 5)] >> (i & 31)) & 1u)
        return e->fail("tnb_set_network: leaf_bits has a bit beyond n_inds"), -1;
  e->cs.release(e->rt);
  e->initialized = false;
  e->n = n_leaves; e->N = 2 * n_leaves - 1; e->n_int = n_leaves - 1; e->n_inds = n_inds; e->W = W;
  e->n_inds_u = n_inds_u; e->Wu = Wu; e->grouped = grouped; e->vw = vw; e->voff = voff;
  e->generic = generic;
  e->Ws = (W + 3) / 4 * 4;
  e->Npad = (e->N + 7) / 8 * 8;
  // everything below lives in the virtual index space
  std::vector<uint32_t> leaf_bits(size_t(n_leaves) * W);
  for (int t = 0; t < n_leaves; ++t) e->expand_row(leaf_bits_u + size_t(t) * Wu, &leaf_bits[size_t(t) * W]);
  // holders of every index (hyper-index = on 3+ tensors, or on 2 and open)
  std::vector<uint16_t> cnt(size_t(W) * 32, 0);
  std::vector<int16_t> own(size_t(2) * n_inds, int16_t(-1));
  for (int t = 0; t < n_leaves; ++t)
    for (int w = 0; w < W; ++w) {
      uint32_t v = leaf_bits[size_t(t) * W + w];
      while (v) {
        const int i = w * 32 + __builtin_ctz(v);
        v &= v - 1;
        if (++cnt[size_t(i)] <= 2) own[size_t(cnt[size_t(i)] - 1) * n_inds + i] = int16_t(t);
      }
    }
  e->h_holders.assign(cnt.begin(), cnt.begin() + n_inds);
  e->h_output.assign(size_t(W), 0u);
  e->dim = dim;
  e->log2d = std::log2(double(dim));
  e->h_leaf_bits = leaf_bits;
  std::vector<uint32_t> padded(size_t(n_leaves) * e->Ws, 0u);
  for (int t = 0; t < n_leaves; ++t)
    std::memcpy(&padded[size_t(t) * e->Ws], &leaf_bits[size_t(t) * W], sizeof(uint32_t) * size_t(W));
  std::vector<uint32_t> leader(size_t(e->Ws), 0u);
  std::vector<uint8_t> gw(size_t(e->Ws) * 32, 0);
  for (int i = 0; i < n_inds_u; ++i) {
    const int k = voff[size_t(i)];
    leader[size_t(k) >> 5] |= 1u << (k & 31);
    gw[size_t(k)] = uint8_t(vw[size_t(i)]);
  }
  void* old[] = {e->d_leaf_bits, e->d_pow_tab, e->d_net_own, e->d_hcount0, e->d_leader, e->d_gw, e->d_sparse,
                 e->d_skip, e->d_gdims, e->d_glog2};
  for (void* q : old) e->rt.free_(q);
  e->d_gdims = nullptr; e->d_glog2 = nullptr;
  e->d_sparse = nullptr; e->n_projs = 0;  // a new network starts with the simple cost model
  e->d_skip = nullptr;
  e->d_leaf_bits = nullptr; e->d_pow_tab = nullptr; e->d_net_own = nullptr; e->d_hcount0 = nullptr;
  e->d_leader = nullptr; e->d_gw = nullptr;
  if (!alloc_to(e->rt, e->d_net_own, own.size()) || !e->rt.h2d(e->d_net_own, own.data(), own.size() * sizeof(int16_t)) ||
      !alloc_to(e->rt, e->d_hcount0, size_t(e->Ws) * 32) || !alloc_to(e->rt, e->d_leader, leader.size()) ||
      !alloc_to(e->rt, e->d_gw, gw.size()) || !e->rt.h2d(e->d_leader, leader.data(), leader.size() * sizeof(uint32_t)) ||
      !e->rt.h2d(e->d_gw, gw.data(), gw.size()))
    return e->rtfail(), -3;
  if (!refresh_hyper(e)) return -3;
  if (generic) {
    std::vector<double> gd(size_t(e->Ws) * 32, 1.0), gl(size_t(e->Ws) * 32, 0.0);
    for (int i = 0; i < n_inds_u; ++i) {
      gd[size_t(i)] = double(dims[i]);
      gl[size_t(i)] = std::log2(double(dims[i]));
    }
    if (!alloc_to(e->rt, e->d_gdims, gd.size()) || !alloc_to(e->rt, e->d_glog2, gl.size()) ||
        !e->rt.h2d(e->d_gdims, gd.data(), gd.size() * sizeof(double)) ||
        !e->rt.h2d(e->d_glog2, gl.data(), gl.size() * sizeof(double)) || !e->rt.sync())
      return e->rtfail(), -3;
  }
  if (!alloc_to(e->rt, e->d_leaf_bits, padded.size() + kTailWords)) return e->rtfail(), -3;
  std::vector<double> tab(size_t(n_inds) + 1);
  for (int k = 0; k <= n_inds; ++k) tab[size_t(k)] = std::pow(double(dim), double(k));
  if (!alloc_to(e->rt, e->d_pow_tab, tab.size())) return e->rtfail(), -3;
  if (!e->rt.h2d(e->d_leaf_bits, padded.data(), padded.size() * sizeof(uint32_t)) ||
      !e->rt.h2d(e->d_pow_tab, tab.data(), tab.size() * sizeof(double)) || !e->rt.sync())
    return e->rtfail(), -3;
  return 0;
}

int tnb_set_output_inds(tnb_engine* e, const uint32_t* output_bits) {
  if (!e) return -1;
  if (e->n == 0) return e->fail("tnb_set_output_inds: call tnb_set_network first"), -1;
  e->cs.release(e->rt);
  e->initialized = false;
  e->h_output.assign(size_t(e->W), 0u);
  if (output_bits) {
    std::vector<uint32_t> u(output_bits, output_bits + e->Wu);
    if (e->n_inds_u & 31) u[size_t(e->Wu) - 1] &= (1u << (e->n_inds_u & 31)) - 1u;
    e->expand_row(u.data(), e->h_output.data());
  }
  return refresh_hyper(e) ? 0 : -3;
}

int tnb_is_hyper(tnb_engine* e) { return e && e->hyper ? 1 : 0; }

int tnb_set_skip_slices(tnb_engine* e, const uint32_t* skip_bits) {
  if (!e) return -1;
  if (e->n == 0) return e->fail("tnb_set_skip_slices: call tnb_set_network first"), -1;
  e->cs.release(e->rt);
  e->initialized = false;
  e->rt.free_(e->d_skip);
  e->d_skip = nullptr;
  if (!skip_bits) return 0;
  std::vector<uint32_t> u(skip_bits, skip_bits + e->Wu), v(size_t(e->Ws) + kTailWords, 0u);
  if (e->n_inds_u & 31) u[size_t(e->Wu) - 1] &= (1u << (e->n_inds_u & 31)) - 1u;
  e->expand_row(u.data(), v.data());
  if (!alloc_to(e->rt, e->d_skip, v.size()) || !e->rt.h2d(e->d_skip, v.data(), v.size() * sizeof(uint32_t)) ||
      !e->rt.sync())
    return e->rtfail(), -3;
  return 0;
}

int tnb_set_sparse_inds(tnb_engine* e, const uint32_t* sparse_bits, uint64_t n_projs) {
  if (!e) return -1;
  if (e->n == 0) return e->fail("tnb_set_sparse_inds: call tnb_set_network first"), -1;
  if (sparse_bits && n_projs == 0) return e->fail("'n_projs' must be a positive number."), -1;
  e->cs.release(e->rt);
  e->initialized = false;
  e->rt.free_(e->d_sparse);
  e->d_sparse = nullptr;
  e->n_projs = 0;
  if (!sparse_bits) return 0;
  std::vector<uint32_t> u(sparse_bits, sparse_bits + e->Wu), v(size_t(e->Ws) + kTailWords, 0u);
  if (e->n_inds_u & 31) u[size_t(e->Wu) - 1] &= (1u << (e->n_inds_u & 31)) - 1u;
  e->expand_row(u.data(), v.data());
  if (!alloc_to(e->rt, e->d_sparse, v.size()) || !e->rt.h2d(e->d_sparse, v.data(), v.size() * sizeof(uint32_t)) ||
      !e->rt.sync())
    return e->rtfail(), -3;
  e->n_projs = n_projs;
  return 0;
}

int tnb_set_mode(tnb_engine* e, double max_width, int update_slices_every, int disable_shared_inds, int prob_kind,
                 int rng_kind, int layout) {
  if (!e) return -1;
  if (prob_kind < 0 || prob_kind > 2 || rng_kind < 0 || rng_kind > 2 || layout < 0 || layout > 3)
    return e->fail("tnb_set_mode: invalid arguments"), -1;
  e->finite = !(max_width < 0.0) && !std::isinf(max_width) && !std::isnan(max_width);
  e->max_width = e->finite ? float(max_width) : 0.f;
  e->every = update_slices_every > 0 ? update_slices_every : 0;
  e->dsi = disable_shared_inds != 0;
  e->prob_kind = prob_kind;
  e->rng_kind = rng_kind;
  e->layout = layout;
  e->cs.release(e->rt);  // chains are (re)built under the new mode: call tnb_set_chains afterwards
  e->initialized = false;
  return 0;
}

int tnb_set_prob(tnb_engine* e, int prob_kind) {
  if (!e) return -1;
  if (prob_kind < 0 || prob_kind > 2) return e->fail("tnb_set_prob: invalid arguments"), -1;
  e->prob_kind = prob_kind;
  return 0;
}

int tnb_set_new_slices(tnb_engine* e, int max_number_new_slices) {
  if (!e) return -1;
  if (max_number_new_slices < 0) return e->fail("tnb_set_new_slices: invalid arguments"), -1;
  e->max_new = max_number_new_slices;
  return 0;
}

int tnb_set_update_slices(tnb_engine* e, int update_slices_every) {
  if (!e) return -1;
  e->every = update_slices_every > 0 ? update_slices_every : 0;
  return 0;
}

int tnb_set_chains(tnb_engine* e, int n_chains, const int32_t* parent, const int32_t* child0, const int32_t* child1,
                   const uint64_t* seeds, uint64_t chain_id0) {
  if (!e) return -1;
  if (e->n == 0) return e->fail("tnb_set_chains: call tnb_set_network first"), -1;
  if (n_chains < 1 || !parent || !child0 || !child1 || !seeds) return e->fail("tnb_set_chains: invalid arguments"), -1;
  e->cs.release(e->rt);
  e->initialized = false;
  e->clear_resume();
  e->chain_id0 = chain_id0;
  if (!make_chains(e, e->cs, n_chains, parent, child0, child1, true, true)) { e->cs.release(e->rt); return -2; }
  if (!check_shared(e, child0, child1, n_chains)) { e->cs.release(e->rt); return -2; }
  e->h_seeds.assign(seeds, seeds + n_chains);
  if (!e->rt.h2d(e->cs.seeds, seeds, size_t(n_chains) * sizeof(uint64_t)) || !e->rt.sync()) return e->rtfail(), -3;
  return 0;
}

int tnb_generate_chains(tnb_engine* e, int n_chains, const uint64_t* seeds, uint64_t chain_id0, int method) {
  if (!e) return -1;
  if (e->n == 0) return e->fail("tnb_generate_chains: call tnb_set_network first"), -1;
  if (n_chains < 1 || !seeds || (method != TNB_TREES_GREEDY && method != TNB_TREES_RANDOM))
    return e->fail("tnb_generate_chains: invalid arguments"), -1;
  if (e->hyper && e->n > 4 * e->Ws * 32)
    return e->fail("tnb_generate_chains: device tree construction is not supported for hyper-index networks with more "
                   "than 128 tensors per 32 indices (use tnb_random_trees_out + tnb_set_chains)"), -2;
  ChainSet& cs = e->cs;
  cs.release(e->rt);
  e->initialized = false;
  e->clear_resume();
  e->chain_id0 = chain_id0;
  if (!alloc_chains(e, cs, n_chains, true, true) || !alloc_to(e->rt, cs.kpop, size_t(n_chains) * e->Npad) ||
      !alloc_to(e->rt, cs.tree_fail, size_t(n_chains)) || !alloc_to(e->rt, cs.escore, size_t(n_chains) * e->Ws * 32)) {
    e->rtfail();
    cs.release(e->rt);
    return -3;
  }
  e->h_seeds.assign(seeds, seeds + n_chains);
  if (!e->rt.h2d(cs.seeds, seeds, size_t(n_chains) * sizeof(uint64_t)) ||
      !e->rt.fill_ff(cs.par, size_t(n_chains) * e->Npad * sizeof(int16_t)))
    return e->rtfail(), -3;
  if (e->n_int > 0) {
    Params P;
    fill_params(e, cs, P);
    P.tree_method = method;
    if (!launch_treegen(e->rt, P, e->tile, e->wpl)) return e->rtfail(), -3;
    std::vector<int> failed(size_t(n_chains), 0);
    if (!e->rt.d2h(failed.data(), cs.tree_fail, failed.size() * sizeof(int))) return e->rtfail(), -3;
    for (int f : failed)
      if (f) {
        cs.release(e->rt);
        return e->fail("tnb_generate_chains: the network is not connected"), -2;
      }
  } else if (!e->rt.sync()) {
    return e->rtfail(), -3;
  }
  return 0;
}

int tnb_set_resume(tnb_engine* e, const uint32_t* mt_state, const uint32_t* slices, const int32_t* best_parent,
                   const int32_t* best_child0, const int32_t* best_child1, const uint32_t* best_slices) {
  if (!e) return -1;
  if (e->cs.n_chains == 0) return e->fail("tnb_set_resume: call tnb_set_chains first"), -1;
  if (e->initialized) return e->fail("tnb_set_resume: the chains are already constructed; call it right after "
                                     "tnb_set_chains"), -1;
  if (mt_state && e->rng_kind != TNB_RNG_MT19937)
    return e->fail("tnb_set_resume: generator states belong to TNB_RNG_MT19937 mode"), -1;
  if ((best_parent || best_child0 || best_child1) && !(best_parent && best_child0 && best_child1))
    return e->fail("tnb_set_resume: invalid arguments"), -1;
  if ((slices || best_slices) && !e->finite) return e->fail("tnb_set_resume: slices need a max_width"), -1;
  const size_t nc = size_t(e->cs.n_chains);
  e->clear_resume();
  if (mt_state) {
    for (size_t c = 0; c < nc; ++c)
      if (mt_state[c * 625 + 624] > 624u) return e->fail("tnb_set_resume: invalid generator state"), -1;
    e->rs_mt.assign(mt_state, mt_state + nc * 625);
  }
  if (slices) e->rs_slices.assign(slices, slices + nc * e->Wu);
  if (best_slices) e->rs_bslices.assign(best_slices, best_slices + nc * e->Wu);
  if (best_parent) {
    e->rs_bp.assign(best_parent, best_parent + nc * e->N);
    e->rs_ba.assign(best_child0, best_child0 + nc * e->N);
    e->rs_bb.assign(best_child1, best_child1 + nc * e->N);
  }
  return 0;
}

int tnb_set_trace(tnb_engine* e, int n_chains, uint64_t cap_records, uint32_t cap_reslices) {
  if (!e) return -1;
  ChainSet& cs = e->cs;
  if (cs.n_chains == 0) return e->fail("tnb_set_trace: call tnb_set_chains / tnb_generate_chains first"), -1;
  if (n_chains < 0 || n_chains > cs.n_chains) return e->fail("tnb_set_trace: invalid arguments (chain count)"), -1;
  if (n_chains > 0 && (e->rng_kind != TNB_RNG_PHILOX || !e->pow2_costs() || e->hyper))
    return e->fail("tnb_set_trace: the decision trace is not supported outside the production kernels "
                   "(TNB_RNG_PHILOX, 2^popcount costs, no hyper-indices)"), -1;
  for (void* p : {(void*)cs.trace, (void*)cs.trace_n, (void*)cs.trace_S, (void*)cs.trace_sn}) e->rt.free_(p);
  cs.trace = nullptr; cs.trace_n = nullptr; cs.trace_S = nullptr; cs.trace_sn = nullptr;
  cs.trace_chains = 0; cs.trace_cap = 0; cs.trace_scap = 0;
  if (n_chains == 0) return 0;
  const size_t nc = size_t(n_chains);
  if (!alloc_to(e->rt, cs.trace, nc * std::max<uint64_t>(cap_records, 1)) || !alloc_to(e->rt, cs.trace_n, nc) ||
      !alloc_to(e->rt, cs.trace_S, nc * std::max<uint32_t>(cap_reslices, 1) * size_t(e->Ws)) ||
      !alloc_to(e->rt, cs.trace_sn, nc) || !e->rt.sync())
    return e->rtfail(), -3;
  cs.trace_chains = n_chains; cs.trace_cap = cap_records; cs.trace_scap = cap_reslices;
  return 0;
}

int tnb_get_trace(tnb_engine* e, int chain, uint64_t* n_records, void* records, uint32_t* n_reslices, uint32_t* slices) {
  if (!e) return -1;
  ChainSet& cs = e->cs;
  if (chain < 0 || chain >= cs.trace_chains) return e->fail("tnb_get_trace: chain range"), -1;
  unsigned long long n = 0;
  uint32_t ns = 0;
  if (!e->rt.d2h(&n, cs.trace_n + chain, sizeof n) || !e->rt.d2h(&ns, cs.trace_sn + chain, sizeof ns)) return e->rtfail(), -3;
  if (n_records) *n_records = n;
  if (n_reslices) *n_reslices = ns;
  const size_t have = size_t(std::min<unsigned long long>(n, cs.trace_cap));
  if (records && !e->rt.d2h(records, cs.trace + size_t(chain) * cs.trace_cap, have * sizeof(TraceRec))) return e->rtfail(), -3;
  if (slices) {
    const size_t hs = std::min<size_t>(ns, cs.trace_scap);
    std::vector<uint32_t> v(hs * size_t(e->Ws));
    if (!e->rt.d2h(v.data(), cs.trace_S + size_t(chain) * cs.trace_scap * size_t(e->Ws), v.size() * sizeof(uint32_t)))
      return e->rtfail(), -3;
    for (size_t k = 0; k < hs; ++k) e->contract_row(&v[k * size_t(e->Ws)], slices + k * size_t(e->Wu));
  }
  return 0;
}

int tnb_get_node_costs(tnb_engine* e, int chain, double* ccost) {
  if (!e) return -1;
  if (!ensure_init(e)) return -2;
  if (chain < 0 || chain >= e->cs.n_chains || !ccost) return e->fail("tnb_get_node_costs: arguments"), -1;
  for (int z = 0; z < e->n; ++z) ccost[z] = 0.0;
  if (e->n_int == 0) return 0;
  const size_t ni = size_t(e->n_int), hs = size_t(e->cs.hstride);
  std::vector<char> hb(ni * hs);
  if (!e->rt.d2h(hb.data(), e->cs.rec + size_t(chain) * ni * hs, (ni - 1) * hs + 16)) return e->rtfail(), -3;
  const bool hi_only = e->rng_kind == TNB_RNG_PHILOX && e->pow2_costs();  // Params::cost_hi
  for (size_t z = 0; z < ni; ++z) {
    if (hi_only) {
      uint32_t hi;
      std::memcpy(&hi, &hb[z * hs + 4], 4);
      const unsigned long long b = (unsigned long long)hi << 32;
      std::memcpy(&ccost[size_t(e->n) + z], &b, sizeof(double));
    } else {
      std::memcpy(&ccost[size_t(e->n) + z], &hb[z * hs + 8], sizeof(double));
    }
  }
  return 0;
}

int tnb_set_stream(tnb_engine* e, const uint32_t* words, uint64_t len) {
  if (!e) return -1;
  if (e->cs.n_chains == 0) return e->fail("tnb_set_stream: call tnb_set_chains first"), -1;
  if (e->rng_kind != TNB_RNG_REPLAY) return e->fail("tnb_set_stream: engine is not in TNB_RNG_REPLAY mode"), -1;
  if (!words || len == 0) return e->fail("tnb_set_stream: invalid arguments"), -1;
  ChainSet& cs = e->cs;
  e->rt.free_(cs.stream);
  cs.stream = nullptr;
  cs.stream_len = len;
  if (!alloc_to(e->rt, cs.stream, size_t(cs.n_chains) * len)) return e->rtfail(), -3;
  if (!e->rt.h2d(cs.stream, words, size_t(cs.n_chains) * len * sizeof(uint32_t)) ||
      !e->rt.zero(cs.cursor, size_t(cs.n_chains) * sizeof(unsigned long long)) ||
      !e->rt.zero(cs.overrun, size_t(cs.n_chains) * sizeof(int)) || !e->rt.sync())
    return e->rtfail(), -3;
  return 0;
}

// 1/beta per sweep for the production kernel's acceptance test  delta <= (2^(-log2(u)/beta) - 1) * total.  The other
// two rules of the reference are its limits: greedy (prob/greedy.hpp:38-42: accept iff delta <= 0) is beta = +inf,
// i.e. 1/beta = 0 and a threshold of exactly 0; always (prob/base.hpp:43-47) is beta = 0, a threshold of +inf.
static bool upload_inv_betas(tnb_engine* e) {
  const size_t n = e->h_betas.size();
  std::vector<float> inv(n);
  for (size_t i = 0; i < n; ++i) {
    if (e->prob_kind == TNB_PROB_GREEDY) inv[i] = 0.f;
    else if (e->prob_kind == TNB_PROB_ALWAYS) inv[i] = 3.0e38f;
    else inv[i] = e->h_betas[i] > 0.0 ? 1.f / float(e->h_betas[i]) : 3.0e38f;
  }
  if (!e->rt.h2d(e->d_inv_betas, inv.data(), n * sizeof(float)) || !e->rt.sync()) return e->rtfail();
  e->inv_kind = e->prob_kind;
  return true;
}

int tnb_set_betas(tnb_engine* e, const double* betas, int64_t n) {
  if (!e) return -1;
  if (!betas || n < 1) return e->fail("tnb_set_betas: invalid arguments"), -1;
  e->rt.free_(e->d_betas);
  e->rt.free_(e->d_inv_betas);
  e->d_betas = nullptr; e->d_inv_betas = nullptr;
  e->h_betas.assign(betas, betas + n);
  if (!alloc_to(e->rt, e->d_betas, size_t(n)) || !alloc_to(e->rt, e->d_inv_betas, size_t(n))) return e->rtfail(), -3;
  if (!e->rt.h2d(e->d_betas, betas, size_t(n) * sizeof(double)) || !upload_inv_betas(e)) return e->rtfail(), -3;
  e->n_betas = n;
  return 0;
}

int tnb_run(tnb_engine* e, int64_t until_sweep) {
  if (!e) return -1;
  if (!e->d_betas) return e->fail("tnb_run: call tnb_set_betas first"), -1;
  if (!ensure_init(e)) return -2;
  if (!mode_ok(e)) return -1;  // tnb_set_prob may have changed the rule since the chains were built
  if (until_sweep >= (int64_t(1) << 31)) return e->fail("tnb_run: until_sweep must be below 2^31"), -1;
  if (e->inv_kind != e->prob_kind && !upload_inv_betas(e)) return -3;  // tnb_set_prob since the schedule was uploaded
  Params P;
  // HBM-resident batches of the production kernels: the hot block stays in L2 for the launches of this call
  struct L2Guard {
    Rt& rt;
    ~L2Guard() { rt.l2_window_off(); }
  } l2guard{e->rt};
  if (e->cs.hot && e->rng_kind == TNB_RNG_PHILOX && e->l2_persist_mb != 0)
    e->rt.l2_window(e->cs.hot, e->cs.hot_bytes, e->l2_persist_mb > 0 ? size_t(e->l2_persist_mb) << 20 : 0);
  for (int guard = 0;; ++guard) {
    fill_params(e, e->cs, P);
    P.until = until_sweep;
#if !defined(TNB_EMU)
    cudaEventRecord(e->rt.ev0, e->rt.stream);
#else
    const auto t0 = std::chrono::steady_clock::now();
#endif
    if (!launch(e->rt, P, e->tile, e->wpl, false, e->finite, stream_mode(e))) return e->rtfail(), -3;
#if !defined(TNB_EMU)
    cudaEventRecord(e->rt.ev1, e->rt.stream);
    if (!e->rt.sync()) return e->rtfail(), -3;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e->rt.ev0, e->rt.ev1);
    e->kernel_ms += double(ms);
#else
    e->kernel_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
#endif
    e->launches += 1;
    if (e->rng_kind == TNB_RNG_PHILOX) break;
    // stream modes: did everybody get there?
    const size_t nc = size_t(e->cs.n_chains);
    std::vector<long long> sw(nc);
    std::vector<int> ov(nc);
    if (!e->rt.d2h(sw.data(), e->cs.sweep_idx, nc * sizeof(long long)) ||
        !e->rt.d2h(ov.data(), e->cs.overrun, nc * sizeof(int)))
      return e->rtfail(), -3;
    for (int o : ov)
      if (o) return e->fail("draw stream exhausted in the middle of a sweep (chain state is no longer valid)"), -4;
    bool done = true;
    for (long long s : sw) done &= s >= until_sweep;
    if (done) break;
    if (e->rng_kind == TNB_RNG_REPLAY) break;  // caller inspects tnb_get_progress and supplies more words
    if (!mt_refill(e, false)) return -3;
    if (guard > (1 << 20)) return e->fail("tnb_run: no progress"), -4;
  }
  return 0;
}

int tnb_run_timed(tnb_engine* e, int64_t until_sweep, double timeout_s, int64_t* reached) {
  if (!e) return -1;
  if (!e->d_betas) return e->fail("tnb_run_timed: call tnb_set_betas first"), -1;
  if (!ensure_init(e)) return -2;
  const size_t nc = size_t(e->cs.n_chains);
  std::vector<long long> sw(nc);
  auto min_sweep = [&](long long& out) {
    if (!e->rt.d2h(sw.data(), e->cs.sweep_idx, nc * sizeof(long long))) return false;
    out = *std::min_element(sw.begin(), sw.end());
    return true;
  };
  long long done = 0;
  if (!min_sweep(done)) return e->rtfail(), -3;
  const bool limited = !(timeout_s < 0.0) && !std::isinf(timeout_s) && !std::isnan(timeout_s);
  if (!limited) {
    const int rc = tnb_run(e, until_sweep);
    if (rc == 0 && !min_sweep(done)) return e->rtfail(), -3;
    if (reached) *reached = done;
    return rc;
  }
  const auto t0 = std::chrono::steady_clock::now();
  auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
  long long chunk = 16;
  while (done < until_sweep && elapsed() < timeout_s) {
    const double t1 = elapsed();
    const long long target = std::min<long long>(until_sweep, done + chunk);
    const int rc = tnb_run(e, target);
    if (rc != 0) return rc;
    if (!min_sweep(done)) return e->rtfail(), -3;
    if (done < target) break;  // REPLAY mode: a stream ran dry
    if (elapsed() - t1 < 0.05) chunk *= 2;
  }
  if (reached) *reached = done;
  return 0;
}

int tnb_get_timing(tnb_engine* e, double* kernel_ms, int64_t* launches) {
  if (!e) return -1;
  if (kernel_ms) *kernel_ms = e->kernel_ms;
  if (launches) *launches = e->launches;
  e->kernel_ms = 0.0;
  e->launches = 0;
  return 0;
}

int tnb_get_costs(tnb_engine* e, double* total, double* min_total) {
  if (!e) return -1;
  if (!ensure_init(e)) return -2;
  const size_t b = size_t(e->cs.n_chains) * sizeof(double);
  if ((total && !e->rt.d2h(total, e->cs.total, b)) || (min_total && !e->rt.d2h(min_total, e->cs.min_total, b)))
    return e->rtfail(), -3;
  return 0;
}

int tnb_get_trees(tnb_engine* e, int best, int chain0, int n, int32_t* parent, int32_t* child0, int32_t* child1) {
  if (!e) return -1;
  if (!ensure_init(e)) return -2;
  if (chain0 < 0 || n < 0 || chain0 + n > e->cs.n_chains) return e->fail("tnb_get_trees: chain range"), -1;
  const size_t ni = size_t(std::max(e->n_int, 1));
  std::vector<int16_t> hp(size_t(n) * e->Npad);
  std::vector<uint32_t> hc(size_t(n) * ni);
  uint32_t* d_ch = e->cs.bch + size_t(chain0) * ni;
  uint32_t* tmp = nullptr;
  if (!best && e->n_int > 0) {  // current trees: gather the children words out of the node records
    if (!alloc_to(e->rt, tmp, hc.size())) return e->rtfail(), -3;
    if (!ch_copy(e->rt, tmp, e->cs.rec + size_t(chain0) * ni * size_t(e->cs.hstride), e->cs.hstride, hc.size(), false)) {
      e->rt.free_(tmp);
      return e->rtfail(), -3;
    }
    d_ch = tmp;
  }
  const bool got = e->rt.d2h(hp.data(), (best ? e->cs.bpar : e->cs.par) + size_t(chain0) * e->Npad,
                             hp.size() * sizeof(int16_t)) &&
                   e->rt.d2h(hc.data(), d_ch, hc.size() * sizeof(uint32_t));
  e->rt.free_(tmp);
  if (!got) return e->rtfail(), -3;
  const int N = e->N, nl = e->n;
  for (int c = 0; c < n; ++c)
    for (int z = 0; z < N; ++z) {
      parent[size_t(c) * N + z] = hp[size_t(c) * e->Npad + z];
      if (z < nl) {
        child0[size_t(c) * N + z] = -1;
        child1[size_t(c) * N + z] = -1;
      } else {
        const uint32_t w = hc[size_t(c) * ni + (z - nl)];
        child0[size_t(c) * N + z] = int32_t(w & 0xffffu);
        child1[size_t(c) * N + z] = int32_t(w >> 16);
      }
    }
  return 0;
}

int tnb_get_trees_packed(tnb_engine* e, int best, int chain0, int n, uint32_t* children) {
  if (!e) return -1;
  if (!ensure_init(e)) return -2;
  if (chain0 < 0 || n < 0 || chain0 + n > e->cs.n_chains || !children) return e->fail("tnb_get_trees_packed: arguments"), -1;
  if (e->n_int == 0 || n == 0) return 0;
  const size_t ni = size_t(e->n_int), count = size_t(n) * ni;
  if (best) return e->rt.d2h(children, e->cs.bch + size_t(chain0) * ni, count * sizeof(uint32_t)) ? 0 : (e->rtfail(), -3);
  uint32_t* tmp = nullptr;
  if (!alloc_to(e->rt, tmp, count)) return e->rtfail(), -3;
  const bool got = ch_copy(e->rt, tmp, e->cs.rec + size_t(chain0) * ni * size_t(e->cs.hstride), e->cs.hstride, count, false) &&
                   e->rt.d2h(children, tmp, count * sizeof(uint32_t));
  e->rt.free_(tmp);
  return got ? 0 : (e->rtfail(), -3);
}

int tnb_get_bits(tnb_engine* e, int chain, uint32_t* node_bits) {
  if (!e) return -1;
  if (!ensure_init(e)) return -2;
  if (chain < 0 || chain >= e->cs.n_chains) return e->fail("tnb_get_bits: chain range"), -1;
  const int W = e->W, Wu = e->Wu;
  for (int t = 0; t < e->n; ++t) e->contract_row(&e->h_leaf_bits[size_t(t) * W], node_bits + size_t(t) * Wu);
  const size_t ni = size_t(std::max(e->n_int, 1));
  const size_t bs = size_t(e->cs.bstride);
  std::vector<char> hb(ni * bs);
  if (!e->rt.d2h(hb.data(), e->cs.bitsb + size_t(chain) * ni * bs, hb.size() - (e->cs.bits_alloc ? 0 : 16)))
    return e->rtfail(), -3;
  for (int z = 0; z < e->n_int; ++z)
    e->contract_row(reinterpret_cast<const uint32_t*>(&hb[size_t(z) * bs]), node_bits + size_t(e->n + z) * Wu);
  return 0;
}

int tnb_get_slices(tnb_engine* e, int best, int chain0, int n, uint32_t* slices) {
  if (!e) return -1;
  if (!ensure_init(e)) return -2;
  if (chain0 < 0 || n < 0 || chain0 + n > e->cs.n_chains) return e->fail("tnb_get_slices: chain range"), -1;
  std::vector<uint32_t> hs(size_t(n) * e->Ws);
  if (!e->rt.d2h(hs.data(), (best ? e->cs.bslices : e->cs.slices) + size_t(chain0) * e->Ws, hs.size() * sizeof(uint32_t)))
    return e->rtfail(), -3;
  for (int c = 0; c < n; ++c) e->contract_row(&hs[size_t(c) * e->Ws], slices + size_t(c) * e->Wu);
  return 0;
}

int tnb_get_progress(tnb_engine* e, int64_t* sweeps, uint64_t* proposals, uint64_t* accepts, uint64_t* width_rejects,
                     uint64_t* words) {
  if (!e) return -1;
  if (!ensure_init(e)) return -2;
  const size_t nc = size_t(e->cs.n_chains);
  if ((sweeps && !e->rt.d2h(sweeps, e->cs.sweep_idx, nc * 8)) || (proposals && !e->rt.d2h(proposals, e->cs.n_prop, nc * 8)) ||
      (accepts && !e->rt.d2h(accepts, e->cs.n_acc, nc * 8)) ||
      (width_rejects && !e->rt.d2h(width_rejects, e->cs.n_wrej, nc * 8)) ||
      (words && !e->rt.d2h(words, e->cs.cursor, nc * 8)))
    return e->rtfail(), -3;
  if (words && e->rng_kind == TNB_RNG_MT19937 && e->words_base.size() == nc)
    for (size_t c = 0; c < nc; ++c) words[c] += e->words_base[c];
  return 0;
}

int tnb_get_counters(tnb_engine* e, uint64_t* proposals, uint64_t* accepts, uint64_t* sweeps) {
  if (!e) return -1;
  if (!ensure_init(e)) return -2;
  const size_t nc = size_t(e->cs.n_chains);
  std::vector<uint64_t> a(nc), b(nc);
  std::vector<int64_t> s(nc);
  if (tnb_get_progress(e, s.data(), a.data(), b.data(), nullptr, nullptr) != 0) return -3;
  uint64_t sa = 0, sb = 0, ss = 0;
  for (size_t i = 0; i < nc; ++i) { sa += a[i]; sb += b[i]; ss += uint64_t(s[i]); }
  if (proposals) *proposals = sa;
  if (accepts) *accepts = sb;
  if (sweeps) *sweeps = ss;
  return 0;
}

int tnb_eval_cost(tnb_engine* e, int n_trees, const int32_t* parent, const int32_t* child0, const int32_t* child1,
                  const uint32_t* slices, double* total_seq, double* total_pc, double* max_width) {
  if (!e) return -1;
  if (e->n == 0) return e->fail("tnb_eval_cost: call tnb_set_network first"), -1;
  if (n_trees < 1 || !parent || !child0 || !child1) return e->fail("tnb_eval_cost: invalid arguments"), -1;
  ChainSet tmp;
  int rc = 0;
  if (!make_chains(e, tmp, n_trees, parent, child0, child1, false, false)) rc = -2;
  if (rc == 0 && slices) {
    std::vector<uint32_t> hs(size_t(n_trees) * e->Ws, 0u);
    for (int c = 0; c < n_trees; ++c)
      e->expand_row(slices + size_t(c) * e->Wu, &hs[size_t(c) * e->Ws]);
    if (!e->rt.h2d(tmp.slices, hs.data(), hs.size() * sizeof(uint32_t)) || !e->rt.sync()) { e->rtfail(); rc = -3; }
  }
  if (rc == 0) {
    Params P;
    fill_params(e, tmp, P);
    P.slices_given = 1;
    P.finite = slices != nullptr;
    if (!launch(e->rt, P, e->tile, e->wpl, true, slices != nullptr, false) || !e->rt.sync()) { e->rtfail(); rc = -3; }
  }
  const size_t b = size_t(n_trees) * sizeof(double);
  if (rc == 0 && ((total_seq && !e->rt.d2h(total_seq, tmp.out_seq, b)) || (total_pc && !e->rt.d2h(total_pc, tmp.total, b)) ||
                  (max_width && !e->rt.d2h(max_width, tmp.out_maxw, b)))) {
    e->rtfail();
    rc = -3;
  }
  tmp.release(e->rt);
  return rc;
}

int tnb_flush_l2(tnb_engine* e) {
  if (!e) return -1;
  const size_t bytes = size_t(256) << 20;  // 2 x the 126 MB L2
  if (!e->d_flush && !(e->d_flush = e->rt.alloc(bytes))) return e->rtfail(), -3;
  if (!e->rt.fill_ff(e->d_flush, bytes) || !e->rt.sync()) return e->rtfail(), -3;
  return 0;
}

int tnb_get_config(tnb_engine* e, int* tile, int* words_per_lane, int* layout, int* state_bytes_per_chain) {
  if (!e) return -1;
  if (tile) *tile = e->tile;
  if (words_per_lane) *words_per_lane = e->wpl;
  if (layout) {
    *layout = e->cs.n_chains ? (e->cs.bits_alloc ? TNB_LAYOUT_SPLIT : TNB_LAYOUT_INTERLEAVED) : e->layout;
    if (e->cs.n_chains) {
      Params P;
      fill_params(e, e->cs, P);
      if (P.smem_chain_bytes > 0) *layout = TNB_LAYOUT_SMEM;
    }
  }
  if (state_bytes_per_chain)
    *state_bytes_per_chain = int(size_t(e->n_int) * size_t(e->stride) + size_t(e->Npad) * 2);
  return 0;
}

// ============================================================================================ device groups
// Several GPUs of one box behind one handle, for hosts that are not Python (the Python layer runs one process per
// GPU instead, tnco_b200/dist.py).  Chains are sharded contiguously (chain i -> device floor(i*G/n)),
// seeds and Philox counters use GLOBAL chain ids, so results do not depend on G.  Every fan-out call runs one host
// thread per device.  The exchange step (minimum over devices + the winner's tree) happens on the host: inside one
// process the per-device minima are G doubles that tnb_get_costs has already brought back, so there is nothing for
// NCCL to do here; across processes it is NCCL (dist.py).
}  // extern "C"

#include <thread>

struct tnb_group {
  std::vector<tnb_engine*> eng;
  std::vector<int> lo, hi;  // chain range of every device
  int n_chains = 0;
  std::string err;
};

namespace tnb {
template <class F>
static int group_fan(tnb_group* g, F f) {
  const size_t G = g->eng.size();
  std::vector<int> rc(G, 0);
  std::vector<std::thread> th;
  for (size_t k = 0; k < G; ++k)
    th.emplace_back([&, k] {
#if !defined(TNB_EMU)
      cudaSetDevice(g->eng[k]->rt.device);
#endif
      rc[k] = f(int(k), g->eng[k]);
    });
  for (auto& t : th) t.join();
  for (size_t k = 0; k < G; ++k)
    if (rc[k] != 0) {
      g->err = "device " + std::to_string(k) + ": " + g->eng[k]->err;
      return rc[k];
    }
  return 0;
}
}  // namespace tnb

extern "C" {

int tnb_group_create(tnb_group** out, const int* devices, int n_dev) {
  if (!out || !devices || n_dev < 1) { set_global_error("tnb_group_create: invalid arguments"); return -1; }
  *out = nullptr;
  tnb_group* g = new tnb_group();
  for (int k = 0; k < n_dev; ++k) {
    tnb_engine* e = nullptr;
    const int rc = tnb_create(&e, devices[k]);
    if (rc != 0) {
      for (tnb_engine* x : g->eng) tnb_destroy(x);
      delete g;
      return rc;  // message in tnb_last_error(NULL)
    }
    g->eng.push_back(e);
  }
  *out = g;
  return 0;
}

void tnb_group_destroy(tnb_group* g) {
  if (!g) return;
  for (tnb_engine* e : g->eng) {
#if !defined(TNB_EMU)
    cudaSetDevice(e->rt.device);
#endif
    tnb_destroy(e);
  }
  delete g;
}

int tnb_group_size(const tnb_group* g) { return g ? int(g->eng.size()) : 0; }
tnb_engine* tnb_group_engine(tnb_group* g, int k) { return g && k >= 0 && k < int(g->eng.size()) ? g->eng[size_t(k)] : nullptr; }
const char* tnb_group_last_error(const tnb_group* g) { return g ? g->err.c_str() : global_error(); }

int tnb_group_set_network(tnb_group* g, int n_leaves, int n_inds, const uint32_t* leaf_bits, uint64_t dim,
                          const uint64_t* dims, const uint32_t* output_bits) {
  if (!g) return -1;
  return group_fan(g, [&](int, tnb_engine* e) {
    const int rc = tnb_set_network(e, n_leaves, n_inds, leaf_bits, dim, dims);
    return rc != 0 || !output_bits ? rc : tnb_set_output_inds(e, output_bits);
  });
}

int tnb_group_set_mode(tnb_group* g, double max_width, int update_slices_every, int disable_shared_inds, int prob_kind,
                       int rng_kind, int layout) {
  if (!g) return -1;
  return group_fan(g, [&](int, tnb_engine* e) {
    return tnb_set_mode(e, max_width, update_slices_every, disable_shared_inds, prob_kind, rng_kind, layout);
  });
}

int tnb_group_set_betas(tnb_group* g, const double* betas, int64_t n) {
  if (!g) return -1;
  return group_fan(g, [&](int, tnb_engine* e) { return tnb_set_betas(e, betas, n); });
}

int tnb_group_generate_chains(tnb_group* g, int n_chains, const uint64_t* seeds, int method) {
  if (!g) return -1;
  const int G = int(g->eng.size());
  if (n_chains < G || !seeds) { g->err = "tnb_group_generate_chains: need at least one chain per device"; return -1; }
  g->lo.assign(size_t(G), 0);
  g->hi.assign(size_t(G), 0);
  for (int k = 0; k < G; ++k) {
    g->lo[size_t(k)] = int((long long)n_chains * k / G);
    g->hi[size_t(k)] = int((long long)n_chains * (k + 1) / G);
  }
  g->n_chains = n_chains;
  return group_fan(g, [&](int k, tnb_engine* e) {
    const int lo = g->lo[size_t(k)], hi = g->hi[size_t(k)];
    return tnb_generate_chains(e, hi - lo, seeds + lo, uint64_t(lo), method);
  });
}

int tnb_group_run(tnb_group* g, int64_t until_sweep, double timeout_s, int64_t* reached) {
  if (!g) return -1;
  std::vector<int64_t> r(g->eng.size(), 0);
  const int rc = group_fan(g, [&](int k, tnb_engine* e) { return tnb_run_timed(e, until_sweep, timeout_s, &r[size_t(k)]); });
  if (reached) *reached = *std::min_element(r.begin(), r.end());
  return rc;
}

int tnb_group_get_costs(tnb_group* g, double* total, double* min_total) {
  if (!g) return -1;
  return group_fan(g, [&](int k, tnb_engine* e) {
    const int lo = g->lo[size_t(k)];
    return tnb_get_costs(e, total ? total + lo : nullptr, min_total ? min_total + lo : nullptr);
  });
}

int tnb_group_get_counters(tnb_group* g, uint64_t* proposals, uint64_t* accepts, uint64_t* sweeps) {
  if (!g) return -1;
  const size_t G = g->eng.size();
  std::vector<uint64_t> p(G, 0), a(G, 0), s(G, 0);
  const int rc = group_fan(g, [&](int k, tnb_engine* e) { return tnb_get_counters(e, &p[size_t(k)], &a[size_t(k)], &s[size_t(k)]); });
  uint64_t sp = 0, sa = 0, ss = 0;
  for (size_t k = 0; k < G; ++k) { sp += p[k]; sa += a[k]; ss += s[k]; }
  if (proposals) *proposals = sp;
  if (accepts) *accepts = sa;
  if (sweeps) *sweeps = ss;
  return rc;
}

int tnb_group_get_best(tnb_group* g, double* cost, int64_t* chain, int32_t* parent, int32_t* child0, int32_t* child1,
                       uint32_t* slices) {
  if (!g || g->n_chains == 0) return -1;
  std::vector<double> m(size_t(g->n_chains));
  int rc = tnb_group_get_costs(g, nullptr, m.data());
  if (rc != 0) return rc;
  // strict minimum, ties to the smaller global chain id: the packed-key order of the multi-process exchange
  int best = 0;
  for (int i = 1; i < g->n_chains; ++i)
    if (m[size_t(i)] < m[size_t(best)]) best = i;
  int k = 0;
  while (!(g->lo[size_t(k)] <= best && best < g->hi[size_t(k)])) ++k;
  tnb_engine* e = g->eng[size_t(k)];
#if !defined(TNB_EMU)
  cudaSetDevice(e->rt.device);
#endif
  const int local = best - g->lo[size_t(k)];
  if (parent && child0 && child1 && (rc = tnb_get_trees(e, 1, local, 1, parent, child0, child1)) != 0) { g->err = e->err; return rc; }
  if (slices && e->finite && (rc = tnb_get_slices(e, 1, local, 1, slices)) != 0) { g->err = e->err; return rc; }
  if (cost) *cost = m[size_t(best)];
  if (chain) *chain = best;
  return 0;
}

}  // extern "C"
