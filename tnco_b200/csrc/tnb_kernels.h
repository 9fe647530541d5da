// Chain kernels of the SA engine: one TILE of lanes owns one chain (one reference Optimizer object).
//
//   chain_init    builds a chain's caches from its tree, like the reference constructors
//                 (infinite_memory/optimizer.hpp:61-88, finite_width/greedy/optimizer.hpp:72-115)
//   chain_sweeps  runs leaf->root sweeps == Optimizer::update() (infinite_memory/optimizer.hpp:90-201,
//                 finite_width/greedy/optimizer.hpp:117-390, incl. the random-new-slice move :226-321)
//
// Data mapping: lane `tl` of the tile owns words tl, tl+TILE, ... of every index bitset, so all bitset
// traffic is lane-private (no cross-lane memory hand-off); tile-uniform scalars (topology, costs) are
// stored redundantly by all lanes with the same value (one coalesced transaction), so a lane only ever
// reads back its own stores.  Cross-lane reads exist only in the slicer scratch, the best-tree snapshots and the
// generator's shared-memory batch, and are fenced by Tile::sync().  Without hyper-indices inds(z) = inds(c0) ^
// inds(c1); the HYPER kernels carry the reference's hyper cache next to every index set.
#pragma once
#include <type_traits>

#include "tnb_simt.h"

namespace tnb {

constexpr int kProbMH = 0, kProbGreedy = 1, kProbAlways = 2;

// One record of the decision trace (32 bytes).  w0: bits 0-1 kind (0 sweep start, 1 proposal, 2 re-slice),
// proposal flags bit 2 pick0 (D = child slot 0), bit 3 width gate passed, bit 4 accepted, bit 5 coin drawn (both
// children of B intersect C); bits 16-31 node B.  w1: the event's random word (sweep start: leaf word; proposal:
// float bits of -log2(u) with the coin in the last bit; re-slice: its number).  w2: float bits of 1/beta.
// w3: sweep start: leaf; proposal: node A; re-slice: 1 if the new slices were kept.  d0 / d1: proposal: delta and
// the running total before the move; sweep start: total and min_total; re-slice: cost under the candidate slices
// and under the current ones.
struct TraceRec {
  uint32_t w0, w1, w2, w3;
  double d0, d1;
};

struct Params {
  // network
  int n, N, n_int, n_inds, W, Ws;  // leaves, nodes, internal nodes, indices, words per bitset, row stride
  const uint32_t* leaf_bits;        // [n][Ws]
  const double* pow_tab;            // [n_inds+1] dim^k (host std::pow), unused when dim2
  int dim2;
  double log2d;
  // sparse-index cost model (infinite_memory/cost_model/simple_sparse_inds.hpp:38-49, finite_width/...:38-77):
  //   cost = get_cost(inds - sparse) * min(get_cost(inds & sparse), n_projs), width likewise with log2(n_projs).
  // Served by the table-cost kernels only (the host clears dim2 for such networks); nullptr = simple cost model.
  const uint32_t* sparse;  // [Ws]
  double n_projs;          // (double)n_projs
  // General per-index dimensions (not all equal, not all powers of two): costs and widths are the reference's own
  // sequential loops over the set bits in ascending index order (cost_model/simple.hpp:46-54, finite_width/
  // cost_model/simple.hpp:47-58) -- fp64 products / float32 sums round per step, so the order is part of the
  // result.  Table-cost kernels only; nullptr = off.
  const double* gdims;     // [Ws*32] dimension of every index
  const double* glog2;     // [Ws*32] log2 of it (host std::log2)
  // skip_slices (finite_width/greedy/utils.hpp:76-79): indices the greedy slicer never takes; stream kernels only
  const uint32_t* skip;    // [Ws] or nullptr
  double log2_n_projs;     // log2(n_projs)
  // mode
  int finite, every, dsi, prob_kind;
  float max_width;
  int max_new;  // max_number_new_slices (finite_width/greedy/optimizer.hpp:226-321); stream kernels only, 0 = off
  // chains
  int n_chains, Npad;
  // hyper-indices (an index on 3+ tensors, or on 2 and open): HYPER kernels keep, next to every internal node's
  // index set, hyper[z] = inds[z] & inds[c0] & inds[c1] (HyperCache, infinite_memory/utils.hpp:68-100)
  // per-index dims (powers of two) as groups of adjacent binary indices: only the slicers look at these
  int grouped;
  const uint32_t* leader;   // [Ws] first bit of every group
  const uint8_t* gw;        // [Ws*32] group size (log2 of the index dimension) at the leader positions
  int hyper;                // network has hyper-indices
  int hyp_off;              // byte offset of the hyper row from the node's index-set row (= 4*Ws)
  const uint16_t* hcount0;  // [Ws*32] initial hyper count of every index: holders - 1 (+1 if output), ctree.py:138-156
  int16_t* par;    // [n_chains][Npad]
  // Per internal node a 16-byte header {+0 u32 child0 | child1 << 16, +4 see cost_hi, +8 f64 contraction cost} and
  // an index set u32[Ws].  INTERLEAVED layout (state fits L2): one record {header, index set} of
  // stride = 16 + 4*Ws bytes per node, so everything the walk needs of a node sits in 1-2 adjacent sectors.
  // SPLIT layout (HBM-resident state): headers [n_int] x 16 B and index sets [n_int] x 4*Ws B apart, so the
  // small, hot headers of all chains stay L2-resident while the index sets stream from HBM.
  char* hdr;       // header of node i of chain c at hdr + (c*n_int + i) * hstride
  char* bitsb;     // index set                 at bitsb + (c*n_int + i) * bstride
  int hstride, bstride;
  // 2^popcount kernels (DIM2; production RNG only): a contraction cost is an exact power of two (or +inf), so only the
  // HIGH word of the double is kept -- at header +4, right behind the children word, and +8 is unused: the walk
  // reads and writes a node as ONE 8-byte access and carries one register per cost instead of two.
  int cost_hi;
  double* pc;      // [n_chains][n_int]   partial costs (kept by the parity modes only)
  int16_t* bpar;   // best tree (reference min_ctree)
  uint32_t* bch;
  uint32_t* slices;   // [n_chains][Ws]
  uint32_t* bslices;  // [n_chains][Ws]
  double* total;      // [n_chains] partial_cost[root]
  double* min_total;  // [n_chains]
  const unsigned long long* seeds;
  unsigned long long* rng_ctr;
  unsigned long long chain_id0;
  long long* sweep_idx;
  unsigned long long *n_prop, *n_acc, *n_wrej;
  // draw stream (MT19937 / REPLAY)
  const uint32_t* stream;       // [n_chains][stream_len]
  unsigned long long* cursor;   // [n_chains]
  unsigned long long stream_len, reserve;
  int* overrun;                 // [n_chains]
  // schedule
  const double* betas;
  const float* inv_betas;  // 1/beta, or 3e38 for beta <= 0 (accept everything: (1+x)^-beta >= 1)
  long long n_betas, until;
  // shared-memory-resident chains (TNB_LAYOUT_SMEM): bytes per chain, 0 = state is walked in global memory
  int smem_chain_bytes;
  // scratch for the slicer
  uint16_t* nbig;   // [n_chains][Ws*32]
  int16_t* posbuf;  // [n_chains][Ws*32]
  dbl2* cp2;        // [n_chains][n_int]
  // production re-slicer (finite width, Philox): per-node popcount / leaf count kept in step with the tree
  uint32_t* kwsz;   // [n_chains][Npad] per node: popcount of its index set (unsliced width / log2 d) in the low half,
                    //                   leaves below it in the high half -- one word, one load / store for both
  int16_t* ksp;     // [n_chains][Npad] popcount of the SPARSE part of every node's index set (sparse-index model under
                    //                   the production re-slicer; nullptr otherwise)
  uint32_t* wkey;   // [n_chains][Npad] scratch: (post-order rank << 16 | node) of the wide nodes
  int16_t* word;    // [n_chains][Npad] scratch: wide nodes, then wide nodes in post-order
  int kthr;         // largest popcount whose width still fits max_width
  // initial-tree construction on the device
  const int16_t* net_own;  // [2][n_inds] the (<= 2) leaves holding each index, -1 if none
  int16_t* kpop;           // [n_chains][Npad] popcount of every cluster's index set
  double* escore;          // [n_chains][Ws*32] cached greedy score of every live edge
  int tree_method;         // TNB_TREES_GREEDY / TNB_TREES_RANDOM
  int* tree_fail;          // [n_chains] set when the network turned out to be disconnected
  // decision trace of the production kernel (TRACE instantiations; tests replay it through a CPU restatement of the reference)
  TraceRec* trace;                // [trace_chains][trace_cap]
  unsigned long long* trace_n;    // [trace_chains] records produced (beyond trace_cap: counted, not stored)
  unsigned long long trace_cap;
  uint32_t* trace_S;              // [trace_chains][trace_scap][Ws] candidate slices of every re-slice
  uint32_t* trace_sn;             // [trace_chains] re-slices produced
  uint32_t trace_scap;
  int trace_chains;               // chains [0, trace_chains) are traced
  // init / eval
  int slices_given;
  double* out_seq;   // [n_chains] cost summed in traversal order (get_cost)
  double* out_maxw;  // [n_chains] max log2 width after slicing
};

// ------------------------------------------------------------------------------------------ Philox
TNB_D TNB_INLINE void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                   uint32_t& o0, uint32_t& o1, uint32_t& o2, uint32_t& o3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  o0 = c0; o1 = c1; o2 = c2; o3 = c3;
}

TNB_D TNB_INLINE double uniform_from(uint32_t lo, uint32_t hi) {
  // libstdc++ generate_canonical<double,53> over a 32-bit engine: (lo + hi*2^32) / 2^64, kept below 1
  double u = (double(lo) + double(hi) * 4294967296.0) * 5.42101086242752217003726400434970855712890625e-20;
  if (u >= 1.0) u = 0.99999999999999988897769753748434595763683319091796875;
  return u;
}

template <int TILE>
TNB_D TNB_INLINE unsigned lane_in_tile_here(int tl) {
#if defined(TNB_EMU)
  return unsigned(tl);
#else
  (void)tl;
  unsigned l;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
  return l & unsigned(TILE - 1);
#endif
}

// Counter-based production RNG (Philox4x32-10; key = run seed, counter = (event index / 2, global chain id)).  An
// "event" of a chain is a sweep start (one word -> leaf), a level (one word -> D/E coin, one -> uniform) or a slicer
// draw; a Philox vector serves two events, and a tile generates a whole batch of them at once into shared memory.
template <int TILE>
struct RngPhilox {
  static constexpr bool kFast = true;  // log-domain fp32 acceptance test (see chain_sweeps)
  // Event e of a chain (one per sweep start, one per level, one per slicer draw) takes half (e & 1) of the Philox
  // vector with counter e >> 1: words (0,1) or (2,3) -- word A for a leaf draw / the D/E coin, word B for the uniform.
  // That mapping does not depend on the tile shape, so neither do results.
  //
  // Registers held across the sweep loop: ONE cursor and the current event's word.  The key (run seed), the global
  // chain id and the 64-bit event index of the held batch are only needed when a batch is generated, so they stay in
  // memory / are re-derived there; the batch itself -- kEB events per tile: their words A, and their level words
  // (-log2(u) as a float with the coin in its last bit) -- sits in shared memory, where an event is ONE broadcast load
  // at the cursor: against a shuffle out of two loop-carried registers (which the finite-width kernel, at its register
  // cap, spilled: it read the level word back from local memory at every level) plus the convergence check the
  // compiler brackets every shuffle with.  A full-warp tile holds 64 events (one Philox call per lane); sub-warp
  // tiles hold 32 each (32 / (2 TILE) calls per lane), so that the fixed cost of a refill -- key and event index from
  // memory, two warp syncs -- is paid once per 32 events whatever the tile shape (with TILE events per batch C1 at
  // TILE = 4 paid it every fourth iteration and lost 8 %).
  // Slot k of the batch belongs to event base + k, base = P.rng_ctr[chain] (even) while a generator is live.  `at`
  // is the shared-memory byte address of the next event's level word; the tile's kEB slots are aligned, so "batch
  // used up" is (at & (4 kEB - 1)) == 0 right after an event was consumed, and the next batch is generated then.
  // Outside the sweep loop (constructors) a generator is not `live`: no batch, rng_ctr is the next event itself.
  static constexpr int kEB = TILE == 32 ? 64 : (TILE == 1 ? 2 : 32);  // events per batch (TILE == 1: emulation build)
  static constexpr int kCalls = kEB / (2 * TILE);                      // Philox calls per lane and batch
  static constexpr int kSlotsW = kEB * (32 / (TILE == 1 ? 32 : TILE));  // slots of all tiles of a warp
  const Params* Pp;
  int chain;
  uint32_t at;
  uint32_t e0;   // current event: word A (leaf draw) / the level word (coin | -log2(u))
  uint32_t n_local;  // draws taken by local_next() since the last sync_from0() (zero outside the slicers)
  uint32_t cd;   // sub-warp tiles: iterations until all tiles of the warp start a fresh batch together (tick)
  bool live;
  bool ool = false;  // generate batches out of line (set by the kernels whose loop is instruction-cache bound)
#if defined(TNB_EMU)
  uint32_t batch_[2][kEB];
  TNB_D uint32_t slot0() { return 0u; }
  TNB_D uint32_t& word(uint32_t a, int which) { return batch_[which][(a >> 2) & uint32_t(kEB - 1)]; }
  TNB_D uint32_t get(uint32_t a, int which) { return word(a, which); }
  TNB_D void put2(uint32_t a, int which, uint32_t x, uint32_t y) { word(a, which) = x; word(a + 4u, which) = y; }
  TNB_D void ring_put(int) {}
  TNB_D void sweep_mark() {}
  TNB_D void note_refill() {}
  TNB_D uint32_t sweep_events(uint32_t& s0) const { s0 = 0; return 0; }
  TNB_D uint32_t ring_get(uint32_t) const { return 0; }
  TNB_D void set_flag(uint32_t) {}
  TNB_D uint32_t flag() const { return 0; }
#else
  // Shared memory of the warp (sweep kernels run one warp per block -- kSweepBlock): [0] words A of all slots,
  // [1] their level words, and for full-warp tiles [2] the ring of the walk (the node B of the level that consumed
  // each event, see ring_put) and [3] bookkeeping of the current sweep.  A tile's slots start at its first lane's.
  TNB_D TNB_INLINE uint32_t slot0() {
    __shared__ __align__(256) uint32_t sm[TILE == 32 ? 4 : 2][kSlotsW];
    return uint32_t(__cvta_generic_to_shared(&sm[1][0])) + 4u * uint32_t(kEB) * ((threadIdx.x & 31u) / uint32_t(TILE));
  }
  TNB_D TNB_INLINE uint32_t get(uint32_t a, int which) {
    uint32_t v;
    if (which) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    else asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(-4 * kSlotsW));
    return v;
  }
  TNB_D TNB_INLINE void put2(uint32_t a, int which, uint32_t x, uint32_t y) {
    if (which) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
    else asm volatile("st.shared.v2.u32 [%0+%3], {%1, %2};" ::"r"(a), "r"(x), "r"(y), "n"(-4 * kSlotsW) : "memory");
  }
  // ---- full-warp tiles: which nodes did the current sweep walk?  Every level leaves its node B next to the event
  // it consumed (ONE shared-memory store per level, addressed off the cursor), so a sweep of up to kEB - 1 levels can
  // be listed afterwards -- what the incremental best-tree snapshot needs (snapshot_walk).
  TNB_D TNB_INLINE void ring_put(int B) {
    asm volatile("st.shared.u32 [%0+%2], %1;" ::"r"(at), "r"(B), "n"(4 * kSlotsW) : "memory");
  }
  TNB_D TNB_INLINE uint32_t meta_addr() const { return (at & ~uint32_t(4 * kEB - 1)) + 8u * uint32_t(kSlotsW); }
  TNB_D TNB_INLINE void sweep_mark() {  // sweep start, before its leaf event: remember the cursor, no refill yet
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(meta_addr()), "r"(at), "r"(0u) : "memory");
  }
  TNB_D TNB_INLINE void note_refill() {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(r) : "r"(meta_addr()));
    asm volatile("st.shared.u32 [%0+4], %1;" ::"r"(meta_addr()), "r"(r + 1u) : "memory");
  }
  // events consumed since sweep_mark() (the leaf event + one per level); slot_start = ring slot of the leaf event
  TNB_D TNB_INLINE uint32_t sweep_events(uint32_t& slot_start) const {
    uint32_t a0, r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a0), "=r"(r) : "r"(meta_addr()));
    slot_start = (a0 >> 2) & uint32_t(kEB - 1);
    return r * uint32_t(kEB) + ((at >> 2) & uint32_t(kEB - 1)) - slot_start;
  }
  TNB_D TNB_INLINE uint32_t ring_get(uint32_t slot) const {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];"
                 : "=r"(v)
                 : "r"((at & ~uint32_t(4 * kEB - 1)) + 4u * uint32_t(kSlotsW) + 4u * (slot & uint32_t(kEB - 1))));
    return v;
  }
  TNB_D TNB_INLINE void set_flag(uint32_t f) { asm volatile("st.shared.u32 [%0+8], %1;" ::"r"(meta_addr()), "r"(f) : "memory"); }
  TNB_D TNB_INLINE uint32_t flag() const {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(v) : "r"(meta_addr()));
    return v;
  }
#endif
  TNB_D TNB_INLINE uint32_t pos() const { return live ? (at >> 2) & uint32_t(kEB - 1) : 0u; }
  TNB_D TNB_INLINE unsigned long long base() const { return Pp->rng_ctr[chain]; }
  TNB_D TNB_INLINE void set_base(unsigned long long b) const { Pp->rng_ctr[chain] = b; }
  TNB_D unsigned long long counter() const { return base() + pos(); }
  TNB_D void load(const Params& P, int chain_) {
    Pp = &P;
    chain = chain_;
    e0 = 0;
    at = 0;
    n_local = 0;
    cd = uint32_t(kEB - 1);
    live = false;
  }
  // a fresh batch whose next event is b (batches start at even events: an odd b skips the first slot)
  TNB_D TNB_INLINE void restart(const Tile<TILE>& t, unsigned long long b) {
    at -= 4u * pos();
    set_base(b & ~1ull);
    generate(t, b & ~1ull);
    at += 4u * uint32_t(b & 1ull);
  }
  TNB_D void begin_stream(const Tile<TILE>& t) {  // sweep kernels: the batch that holds the next event
    const unsigned long long b = base();
    live = true;
    at = slot0();
    restart(t, b);
  }
  TNB_D void store(const Params&, int) {
    set_base(counter());
    live = false;
  }
  TNB_D bool can_start(const Params&) const { return true; }
  TNB_D TNB_INLINE void philox(unsigned long long ctr, uint32_t& o0, uint32_t& o1, uint32_t& o2, uint32_t& o3) const {
    const unsigned long long s = Pp->seeds[chain], g = Pp->chain_id0 + (unsigned long long)chain;
    philox4x32_10(uint32_t(ctr), uint32_t(ctr >> 32), uint32_t(g), uint32_t(g >> 32), uint32_t(s), uint32_t(s >> 32),
                  o0, o1, o2, o3);
  }
  // word A of one event (leaf draw, slicer draw), whatever batch is held
  TNB_D TNB_INLINE uint32_t word_a(unsigned long long e) const {
    uint32_t r0, r1, r2, r3;
    philox(e >> 1, r0, r1, r2, r3);
    return (e & 1ull) ? r2 : r0;
  }
  // u = (x + 0.5) / 2^32 in (0,1);  -log2(u) = 32 - log2(x + 0.5), computed once per event.  The D/E coin of the
  // level (last bit of word A) rides in the last mantissa bit of that float -- one load per level instead of two;
  // the bit moves the acceptance threshold by one fp32 ulp, far inside the 3e-6 of ex2.approx.
  static TNB_D TNB_INLINE uint32_t level_word(uint32_t a, uint32_t x) {
#if defined(TNB_EMU)
    const float rf = 32.f - log2f(float(x) + 0.5f);
    uint32_t fb;
    std::memcpy(&fb, &rf, 4);
#else
    const uint32_t fb = __float_as_uint(32.f - __log2f(float(x) + 0.5f));
#endif
    return (fb & ~1u) | (a & 1u);
  }
  // The batch of events bb .. bb + kEB - 1 (bb even) into the tile's slots (at0 = address of slot 0).  Out of line
  // and by value: ONE copy of the ~100 instructions instead of one at each of the five places a batch can start (the
  // sweep loop of the finite-width kernel is close to the instruction cache's size: with the copies inlined 22 % of
  // its stall samples were instruction fetches), and no generator state is forced into local memory by the call.
#if !defined(TNB_EMU)
  static TNB_D TNB_NOINLINE void generate_batch(const Params* P, int chain, uint32_t at0, unsigned long long bb) {
    Tile<TILE> t;
    RngPhilox g;
    g.Pp = P;
    g.chain = chain;
    const unsigned tl = lane_in_tile_here<TILE>(t.tl);
    t.sync();  // (everybody is done with the previous batch)
#pragma unroll 1
    for (int j = 0; j < kCalls; ++j) {
      const unsigned q = tl + unsigned(j * TILE);  // events bb + 2q and bb + 2q + 1
      uint32_t r0, r1, r2, r3;
      g.philox((bb >> 1) + (unsigned long long)q, r0, r1, r2, r3);
      g.put2(at0 + 8u * q, 0, r0, r2);
      g.put2(at0 + 8u * q, 1, level_word(r0, r1), level_word(r2, r3));
    }
    t.sync();
  }
#endif
#if defined(TNB_EMU)
  TNB_D void generate(const Tile<TILE>&, unsigned long long bb) {  // (the emulation build keeps the batch in the object)
    uint32_t r0, r1, r2, r3;
    philox(bb >> 1, r0, r1, r2, r3);
    put2(at, 0, r0, r2);
    put2(at, 1, level_word(r0, r1), level_word(r2, r3));
  }
#else
  // (the small unconstrained kernels inline it -- 2 % faster there; see `ool`)
  TNB_D TNB_INLINE void generate(const Tile<TILE>& t, unsigned long long bb) {
    if (ool) {
      generate_batch(Pp, chain, at, bb);
    } else {
      const unsigned tl = lane_in_tile_here<TILE>(t.tl);
      t.sync();
#pragma unroll 1
      for (int j = 0; j < kCalls; ++j) {
        const unsigned q = tl + unsigned(j * TILE);
        uint32_t r0, r1, r2, r3;
        philox((bb >> 1) + (unsigned long long)q, r0, r1, r2, r3);
        put2(at + 8u * q, 0, r0, r2);
        put2(at + 8u * q, 1, level_word(r0, r1), level_word(r2, r3));
      }
      t.sync();
    }
  }
#endif
  // Sub-warp tiles: every kEB - 1 iterations of the sweep loop ALL tiles of the warp start a fresh batch together
  // (an iteration consumes at most one event and a batch holds at least kEB - 1, so nobody runs dry in between).
  // Refilling tile by tile made the instructions of a refill run with 1/8 of the lanes active several times per
  // iteration (C1: 13.5 of 32 lanes active on average).
  TNB_D TNB_INLINE void tick(const Tile<TILE>& t, uint32_t) {
    if (TILE < 32) {
      if (--cd == 0u) {
        cd = uint32_t(kEB - 1);
        if (pos() != 0u) restart(t, base() + pos());  // (pos == 0: fresh batch, unused)
      }
    }
  }
  TNB_D TNB_INLINE void consumed(const Tile<TILE>& t) {  // one event taken: step, and start the next batch if that was the last
    at += 4u;
    if ((at & uint32_t(4 * kEB - 1)) == 0u) {
      const unsigned long long bb = base() + uint32_t(kEB);
      set_base(bb);
      at -= 4u * uint32_t(kEB);
      if (TILE == 32) note_refill();
      generate(t, bb);
    }
  }
  TNB_D TNB_INLINE uint32_t leaf_word(const Tile<TILE>& t) {  // sweep start: the whole word A
    e0 = get(at, 0);
    consumed(t);
    return e0;
  }
  TNB_D TNB_INLINE void begin_level(const Tile<TILE>& t, int B) {  // one level: coin and -log2(u) in one word
    e0 = get(at, 1);
    if (TILE == 32) ring_put(B);
    consumed(t);
  }
  TNB_D TNB_INLINE uint32_t coin_word(const Tile<TILE>&) { return e0; }
  static TNB_D TNB_INLINE uint32_t float_bits(float f) {
#if defined(TNB_EMU)
    uint32_t b;
    std::memcpy(&b, &f, 4);
    return b;
#else
    return __float_as_uint(f);
#endif
  }
  TNB_D TNB_INLINE float neg_log2_u() const {
#if defined(TNB_EMU)
    float f;
    std::memcpy(&f, &e0, 4);
    return f;
#else
    return __uint_as_float(e0);
#endif
  }
  TNB_D TNB_INLINE double uniform(const Tile<TILE>&) {  // exact path (not used by the production kernels)
    const unsigned long long e = counter() - 1ull;       // the event begin_level() just consumed
    uint32_t r0, r1, r2, r3;
    philox(e >> 1, r0, r1, r2, r3);
    return (e & 1ull) ? uniform_from(r3, r2) : uniform_from(r1, r0);
  }
  // Draws outside the level stream (the slicers): word A of the events after the last one consumed.  A slicer may
  // draw on lane 0 only; sync_from0() afterwards makes every lane skip the events lane 0 used.
  TNB_D uint32_t local_next() {
    const uint32_t a = word_a(counter() + n_local);
    ++n_local;
    return a;
  }
  TNB_D void sync_from0(const Tile<TILE>& t) {
    const uint32_t n = t.bcast(n_local, 0);
    n_local = 0;
    if (n == 0u) return;
    if (!live) {
      set_base(base() + n);
    } else if (pos() + n < uint32_t(kEB)) {
      at += 4u * n;
    } else {  // past the held batch: a fresh one holding the next event
      restart(t, base() + pos() + n);
    }  // (slicer draws come after the walk of their sweep was counted: sweep_events() is taken before)
  }
  TNB_D unsigned long long words() const { return 0; }
  TNB_D int overrun() const { return 0; }
};

// Raw 32-bit draw stream consumed in the reference's own order (std::mt19937 words, or a recorded stream).
template <int TILE>
struct RngStream {
  static constexpr bool kFast = false;  // parity modes keep the reference's pow() acceptance in fp64
  TNB_D TNB_INLINE float neg_log2_u() const { return 0.f; }
  const uint32_t* w;
  unsigned long long cur, len;
  int over;
  TNB_D void load(const Params& P, int chain) {
    w = P.stream + size_t(chain) * P.stream_len;
    cur = P.cursor[chain];
    len = P.stream_len;
    over = P.overrun[chain];
  }
  TNB_D void store(const Params& P, int chain) const {
    P.cursor[chain] = cur;
    P.overrun[chain] = over;
  }
  TNB_D bool can_start(const Params& P) const { return !over && cur + P.reserve <= len; }
  TNB_D TNB_INLINE uint32_t next() {
    if (cur >= len) { over = 1; return 0u; }
    return w[cur++];
  }
  TNB_D TNB_INLINE uint32_t leaf_word(const Tile<TILE>&) { return next(); }
  TNB_D TNB_INLINE void tick(const Tile<TILE>&, uint32_t) {}
  TNB_D TNB_INLINE void begin_level(const Tile<TILE>&, int) {}
  TNB_D TNB_INLINE uint32_t coin_word(const Tile<TILE>&) { return next(); }
  TNB_D TNB_INLINE double uniform(const Tile<TILE>&) {
    const uint32_t lo = next();
    const uint32_t hi = next();
    return uniform_from(lo, hi);
  }
  TNB_D uint32_t local_next() { return next(); }
  TNB_D void sync_from0(const Tile<TILE>& t) {
    cur = t.bcast_u64(cur, 0);
    over = int(t.bcast(uint32_t(over), 0));
  }
  TNB_D unsigned long long words() const { return cur; }
  TNB_D int overrun() const { return over; }
};

// A node as the walk carries it: children word + contraction cost.  The 2^popcount kernels keep the pair in ONE 64-bit
// value -- children in the low word, the high word of the cost (an exact power of two: its low word is zero) in the
// high word -- exactly as it sits in memory: one 8-byte load brings it, a move updates its halves in place and one
// 8-byte store writes it back.  (Loaded into two separate 32-bit variables instead, the compiler moved the cost out of
// the destination pair right behind the load and the warp waited there for the load to land.)
struct NodePacked {
  unsigned long long v = 0ull;
  // (mov.b64 packs / unpacks a register pair for free; 64-bit masks and ORs became real 64-bit additions)
  static TNB_D TNB_INLINE unsigned long long pack(uint32_t lo, uint32_t hi) {
#if defined(TNB_EMU)
    return (unsigned long long)lo | ((unsigned long long)hi << 32);
#else
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
#endif
  }
  TNB_D TNB_INLINE uint32_t w() const {
#if defined(TNB_EMU)
    return uint32_t(v);
#else
    uint32_t lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
    return lo;
#endif
  }
  TNB_D TNB_INLINE uint32_t cost() const {
#if defined(TNB_EMU)
    return uint32_t(v >> 32);
#else
    uint32_t lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
    return hi;
#endif
  }
  TNB_D TNB_INLINE void set_w(uint32_t w) { v = pack(w, cost()); }
  TNB_D TNB_INLINE void set_cost(uint32_t c) { v = pack(w(), c); }
  TNB_D TNB_INLINE double cost_f64() const { return bits_to_f64(pack(0u, cost())); }
};
struct NodeWide {  // table-cost kernels: the cost is a general double
  uint32_t w_ = 0u;
  double c_ = 0.0;
  TNB_D TNB_INLINE uint32_t w() const { return w_; }
  TNB_D TNB_INLINE void set_w(uint32_t w) { w_ = w; }
  TNB_D TNB_INLINE double cost() const { return c_; }
  TNB_D TNB_INLINE void set_cost(double c) { c_ = c; }
  TNB_D TNB_INLINE double cost_f64() const { return c_; }
};

// ------------------------------------------------------------------------------------------ chain view
// node id -> 64-bit factor of an address product: node * stride as a widening 32x32->64 multiply-add onto the base
// pointer is ONE instruction (IMAD.WIDE.U32); a 32-bit product has wrap-around semantics and costs three
// (IMAD, IADD3, IMAD.X) -- with about six such addresses per level.
TNB_D TNB_INLINE unsigned long long w64(int x) { return (unsigned long long)unsigned(x); }

template <int TILE, int WPL>
struct ChainView {
  const Params& P;
  Tile<TILE> t;
  int chain;
  int n;        // leaves: node < n is a leaf
  unsigned Ws;  // words between consecutive bitset rows
  int16_t* par;
  // bases resolved for this lane and pre-offset by -n so that every array is indexed by the node id itself
  // (signed element offsets; the virtual bases are only dereferenced at indices >= n)
  const uint32_t* leaf_lane;  // leaf_bits + tl
  unsigned hstride, bstride;  // bytes between the headers / index sets of consecutive internal nodes
  char* rec;                  // header of internal node z at rec + z*hstride
  char* rec_lane;             // this lane's first word of z's index set at rec_lane + z*bstride (+ 4*k*TILE)
  double* pcv;                // pcv[z] = partial cost of internal node z (parity modes)
  bool lane_ok[WPL];
  bool smem = false;          // rec / rec_lane / par point into shared memory (chain_sweeps<..., SMEM>)

  TNB_D ChainView(const Params& P_, int chain_) : P(P_), chain(chain_) {
    n = P.n;
    Ws = unsigned(P.Ws);
    par = P.par + size_t(chain) * P.Npad;
    const long long row0 = (long long)chain * P.n_int - n;
    hstride = unsigned(P.hstride);
    bstride = unsigned(P.bstride);
    rec = P.hdr + row0 * (long long)P.hstride;
    rec_lane = P.bitsb + row0 * (long long)P.bstride + 4 * t.tl;
    leaf_lane = P.leaf_bits + t.tl;
    pcv = P.pc + row0;
#pragma unroll
    for (int k = 0; k < WPL; ++k) lane_ok[k] = (t.tl + k * TILE) < P.W;
  }
  TNB_D TNB_INLINE void load_bits(int node, uint32_t (&o)[WPL]) const {
    if constexpr (TILE == 32 && WPL == 1) {
      // full-warp tile: "leaf or internal" is warp-uniform, so two predicated loads behind a uniform branch (the
      // leaf one through the read-only path) beat a selected pointer: C2 @ 4096 chains 4.22e9 vs 4.06e9
      if (node < n) {
        const uint32_t* src = leaf_lane + w64(node) * Ws;
        o[0] = lane_ok[0] ? ldg(src) : 0u;
      } else {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(rec_lane + w64(node) * bstride);
        o[0] = lane_ok[0] ? *src : 0u;
      }
    } else {
      // sub-warp tiles (the branch would diverge between the chains of a warp) and multi-word lanes: one address
      // select and predicated loads (C1 +7 %, C5 +10 % against the branching form).  An unconditional load of the
      // whole tile of words plus a mask is shorter still but drags extra sectors through L1 (C2 -5 %).
      const bool leaf = node < n;
      if (smem) {  // leaves in global memory (read-only path), internal nodes in shared memory: two predicated loads
        const uint32_t* gl = leaf_lane + w64(node) * Ws;
        const uint32_t* sh = reinterpret_cast<const uint32_t*>(rec_lane + w64(node) * bstride);
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
          uint32_t v = 0u;
          if (lane_ok[k] && leaf) v = ldg(gl + k * TILE);
          if (lane_ok[k] && !leaf) v = sh[k * TILE];
          o[k] = v;
        }
        return;
      }
      const char* base = leaf ? reinterpret_cast<const char*>(leaf_lane) : const_cast<const char*>(rec_lane);
      const unsigned st = leaf ? 4u * Ws : bstride;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(base + w64(node) * st);
#if !defined(TNB_EMU)
      __builtin_assume(__isGlobal(src));
#endif
#pragma unroll
      for (int k = 0; k < WPL; ++k) o[k] = lane_ok[k] ? src[k * TILE] : 0u;
    }
  }
  // L2 prefetch of a node's whole index set (first and last byte: a row may straddle two lines)
  TNB_D TNB_INLINE void prefetch_row(int node) const {
#if !defined(TNB_EMU)
    const char* p = node < n ? reinterpret_cast<const char*>(leaf_lane - t.tl) + w64(node) * (4u * Ws)
                             : rec_lane - 4 * t.tl + w64(node) * bstride;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 4 * (P.W - 1)));
#else
    (void)node;
#endif
  }
  TNB_D TNB_INLINE void store_bits(int node, const uint32_t (&v)[WPL]) const {
    uint32_t* dst = reinterpret_cast<uint32_t*>(rec_lane + w64(node) * bstride);
#pragma unroll
    for (int k = 0; k < WPL; ++k)
      if (lane_ok[k]) dst[k * TILE] = v[k];
  }
  TNB_D TNB_INLINE void load_hyp(int node, uint32_t (&o)[WPL]) const {  // internal nodes only
    const uint32_t* src = reinterpret_cast<const uint32_t*>(rec_lane + w64(node) * bstride + unsigned(P.hyp_off));
#pragma unroll
    for (int k = 0; k < WPL; ++k) o[k] = lane_ok[k] ? src[k * TILE] : 0u;
  }
  TNB_D TNB_INLINE void store_hyp(int node, const uint32_t (&v)[WPL]) const {
    uint32_t* dst = reinterpret_cast<uint32_t*>(rec_lane + w64(node) * bstride + unsigned(P.hyp_off));
#pragma unroll
    for (int k = 0; k < WPL; ++k)
      if (lane_ok[k]) dst[k * TILE] = v[k];
  }
  // header fields of internal node z
  TNB_D TNB_INLINE uint32_t& ch(int z) const { return *reinterpret_cast<uint32_t*>(rec + w64(z) * hstride); }
  // contraction cost of z, whichever form the batch keeps (P.cost_hi is kernel-uniform)
  static TNB_D TNB_INLINE double hi_to_f64(uint32_t hi) { return bits_to_f64((unsigned long long)hi << 32); }
  static TNB_D TNB_INLINE uint32_t f64_hi(double v) {
#if defined(TNB_EMU)
    unsigned long long u;
    std::memcpy(&u, &v, 8);
    return uint32_t(u >> 32);
#else
    return uint32_t(__double2hiint(v));
#endif
  }
  TNB_D TNB_INLINE uint32_t& cc_hi(int z) const { return *reinterpret_cast<uint32_t*>(rec + w64(z) * hstride + 4); }
  TNB_D TNB_INLINE double& cc_f64(int z) const { return *reinterpret_cast<double*>(rec + w64(z) * hstride + 8); }
  TNB_D TNB_INLINE double cc_get(int z) const { return P.cost_hi ? hi_to_f64(cc_hi(z)) : cc_f64(z); }
  TNB_D TNB_INLINE void cc_set(int z, double v) const {
    if (P.cost_hi) cc_hi(z) = f64_hi(v);
    else cc_f64(z) = v;
  }
  // one node of the walk: children word + cost.  DIM2 kernels: one 8-byte access, the cost as its high word.
  TNB_D TNB_INLINE void load_node(int z, uint32_t& children, uint32_t& cost_hi_word) const {
#if defined(TNB_EMU)
    children = ch(z);
    cost_hi_word = cc_hi(z);
#else
    // (one 64-bit scalar: as a uint2 the compiler splits the access into two 32-bit ones)
    const unsigned long long v = *reinterpret_cast<const unsigned long long*>(rec + w64(z) * hstride);
    children = uint32_t(v);
    cost_hi_word = uint32_t(v >> 32);
#endif
  }
  TNB_D TNB_INLINE void load_node(int z, NodePacked& h) const {
#if defined(TNB_EMU)
    h.v = (unsigned long long)ch(z) | ((unsigned long long)cc_hi(z) << 32);
#else
    h.v = *reinterpret_cast<const unsigned long long*>(rec + w64(z) * hstride);
#endif
  }
  TNB_D TNB_INLINE void store_node(int z, const NodePacked& h) const {
#if defined(TNB_EMU)
    ch(z) = h.w();
    cc_hi(z) = h.cost();
#else
    *reinterpret_cast<unsigned long long*>(rec + w64(z) * hstride) = h.v;
#endif
  }
  TNB_D TNB_INLINE void load_node(int z, NodeWide& h) const {
    h.w_ = ch(z);
    h.c_ = cc_f64(z);
  }
  TNB_D TNB_INLINE void store_node(int z, const NodeWide& h) const { store_node(z, h.w_, h.c_); }
  TNB_D TNB_INLINE void load_node(int z, uint32_t& children, double& cost) const {
    children = ch(z);
    cost = cc_f64(z);
  }
  TNB_D TNB_INLINE void store_node(int z, uint32_t children, double cost) const {  // table-cost production kernels
#if defined(TNB_EMU)
    ch(z) = children;
    cc_f64(z) = cost;
#else
    const unsigned long long cb = (unsigned long long)__double_as_longlong(cost);
    *reinterpret_cast<uint4*>(rec + w64(z) * hstride) = make_uint4(children, 0u, uint32_t(cb), uint32_t(cb >> 32));
#endif
  }
  TNB_D TNB_INLINE void store_node(int z, uint32_t children, uint32_t cost_hi_word) const {
#if defined(TNB_EMU)
    ch(z) = children;
    cc_hi(z) = cost_hi_word;
#else
    *reinterpret_cast<unsigned long long*>(rec + w64(z) * hstride) =
        (unsigned long long)children | ((unsigned long long)cost_hi_word << 32);
#endif
  }
  TNB_D TNB_INLINE double pc_of(int node) const { return node < n ? 0.0 : pcv[node]; }
  TNB_D TNB_INLINE double cost_of(int k) const {
    // pow(dim, k) (infinite_memory/cost_model/simple.hpp:45) from the host-computed table (std::pow)
    return ldg(P.pow_tab + k);
  }
  TNB_D TNB_INLINE float width_of(int k) const { return float(P.log2d * double(k)); }  // fw simple.hpp:47
  // sparse-index model: k = popcount of the whole set, ks = popcount of its sparse part
  TNB_D TNB_INLINE double cost_sp(int k, int ks) const {
    const double x = cost_of(ks), y = P.n_projs;
    return cost_of(k - ks) * (x < y ? x : y);
  }
  TNB_D TNB_INLINE float width_sp(int k, int ks) const {  // float + min(float, double) -> float, as the reference
    const float ws = width_of(ks);
    return width_of(k - ks) + (double(ws) < P.log2_n_projs ? ws : float(P.log2_n_projs));
  }
  // ---- general per-index dimensions: every lane walks the whole set word by word (words fetched by shuffle from
  // their owners), so the result is tile-uniform and the rounding order is the reference's.  CONV: called from the
  // sweep loop (converged-lane shuffles).
  // (out of line and not unrolled: a rare path that would otherwise be inlined into every table-cost kernel at a
  //  dozen call sites)
  template <bool CONV>
  TNB_D TNB_NOINLINE double gcost(const uint32_t (&u)[WPL]) const {
    double r = 1.0;
#pragma unroll 1
    for (int k = 0; k < WPL; ++k)
#pragma unroll 1
      for (int o = 0; o < TILE; ++o) {
        const int w = k * TILE + o;
        if (w >= P.W) break;
        uint32_t v = CONV ? t.bcast_c(u[k], o) : t.bcast(u[k], o);
        while (v) {
          r *= ldg(P.gdims + w * 32 + ctz32(v));
          v &= v - 1;
        }
      }
    return r;
  }
  template <bool CONV>
  TNB_D TNB_NOINLINE float gwidth(const uint32_t (&u)[WPL]) const {
    float wd = 0.f;
#pragma unroll 1
    for (int k = 0; k < WPL; ++k)
#pragma unroll 1
      for (int o = 0; o < TILE; ++o) {
        const int w = k * TILE + o;
        if (w >= P.W) break;
        uint32_t v = CONV ? t.bcast_c(u[k], o) : t.bcast(u[k], o);
        while (v) {
          wd = float(double(wd) + ldg(P.glog2 + w * 32 + ctz32(v)));  // width_ += log2(dims[pos]), width_ float
          v &= v - 1;
        }
      }
    return wd;
  }
  // the same under the active cost model (sparse indices or not); SP = this lane's words of the sparse mask or 0
  template <bool CONV>
  TNB_D TNB_NOINLINE double gcost_model(const uint32_t (&u)[WPL], const uint32_t (&SP)[WPL]) const {
    if (P.sparse == nullptr) return gcost<CONV>(u);
    uint32_t d[WPL], sp[WPL];
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      d[k] = u[k] & ~SP[k];
      sp[k] = u[k] & SP[k];
    }
    const double x = gcost<CONV>(sp), y = P.n_projs;
    return gcost<CONV>(d) * (x < y ? x : y);
  }
  TNB_D TNB_INLINE float cap_w(float ws) const { return double(ws) < P.log2_n_projs ? ws : float(P.log2_n_projs); }
  template <bool CONV>
  TNB_D TNB_NOINLINE float gwidth_model(const uint32_t (&u)[WPL], const uint32_t (&SP)[WPL]) const {
    if (P.sparse == nullptr) return gwidth<CONV>(u);
    uint32_t d[WPL], sp[WPL];
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      d[k] = u[k] & ~SP[k];
      sp[k] = u[k] & SP[k];
    }
    return gwidth<CONV>(d) + cap_w(gwidth<CONV>(sp));
  }
  TNB_D TNB_INLINE void load_sparse(uint32_t (&o)[WPL]) const {
#pragma unroll
    for (int k = 0; k < WPL; ++k) o[k] = lane_ok[k] ? ldg(P.sparse + t.tl + k * TILE) : 0u;
  }
  // stackless post-order over the CURRENT topology, children[0] subtree first (utils.hpp:35-52)
  TNB_D int po_first() const {
    int x = P.N - 1;
    while (x >= n) x = int(ch(x) & 0xffffu);
    return x;
  }
  TNB_D int po_next(int x) const {
    if (x == P.N - 1) return -1;
    const int p = par[x];
    const uint32_t c = ch(p);
    if (int(c & 0xffffu) == x) {
      x = int(c >> 16);
      while (x >= n) x = int(ch(x) & 0xffffu);
      return x;
    }
    return p;
  }
};

template <int WPL>
TNB_D TNB_INLINE uint32_t popc_or3(const uint32_t (&a)[WPL], const uint32_t (&b)[WPL], const uint32_t (&c)[WPL]) {
  uint32_t k = 0;
#pragma unroll
  for (int i = 0; i < WPL; ++i) k += popc32(a[i] | b[i] | c[i]);
  return k;
}

// Post-order cost pass (CostCache ctor, infinite_memory/utils.hpp:32-56; with slices finite_width/utils.hpp:36-45).
// Optionally (re)builds the index sets of internal nodes and tracks get_cost's sequential sum and the widest node.
// where cost_pass puts its results: the chain's own caches, or a scratch array of {contraction, partial} pairs
template <int TILE, int WPL>
struct CacheSink {
  const ChainView<TILE, WPL>& c;
  TNB_D TNB_INLINE double pc(int node) const { return c.pc_of(node); }
  TNB_D TNB_INLINE void put(int z, double cost, double pcost) const {
    c.cc_set(z, cost);
    c.pcv[z] = pcost;
  }
};
struct ScratchSink {
  dbl2* dst;
  int n;
  TNB_D TNB_INLINE double pc(int node) const { return node < n ? 0.0 : dst[node - n].y; }
  TNB_D TNB_INLINE void put(int z, double cost, double pcost) const { dst[z - n] = make_dbl2(cost, pcost); }
};

template <int TILE, int WPL, bool WIDTHS, class Sink>
TNB_D void cost_pass(const ChainView<TILE, WPL>& c, const uint32_t (&S)[WPL], const Sink& dst, double& seq,
                     double& maxw) {
  const Params& P = c.P;
  seq = 0.0;
  uint32_t maxk = 0;
  float maxw_sp = 0.f;
  const bool sparse = P.sparse != nullptr;
  const bool gen = P.gdims != nullptr;
  uint32_t SP[WPL];
#pragma unroll
  for (int i = 0; i < WPL; ++i) SP[i] = 0u;
  if (sparse) c.load_sparse(SP);
  for (int z = c.po_first(); z >= 0; z = c.po_next(z)) {
    uint32_t kw_ = 0;
    if (gen) {  // general per-index dimensions: the reference's sequential loops
      if (WIDTHS) {
        uint32_t x[WPL];
        c.load_bits(z, x);
#pragma unroll
        for (int i = 0; i < WPL; ++i) x[i] &= ~S[i];
        const float w = c.template gwidth_model<false>(x, SP);
        maxw_sp = w > maxw_sp ? w : maxw_sp;
      }
      if (z < P.n) continue;
      const uint32_t cw = c.ch(z);
      const int a = int(cw & 0xffffu), b = int(cw >> 16);
      uint32_t xa[WPL], xb[WPL];
      c.load_bits(a, xa);
      c.load_bits(b, xb);
#pragma unroll
      for (int i = 0; i < WPL; ++i) xa[i] |= xb[i] | S[i];
      const double cost = c.template gcost_model<false>(xa, SP);
      dst.put(z, cost, cost + dst.pc(a) + dst.pc(b));
      seq += cost;
      continue;
    }
    if (WIDTHS) {  // sliced popcount of the node's own index set (widest node)
      uint32_t x[WPL];
      c.load_bits(z, x);
#pragma unroll
      for (int i = 0; i < WPL; ++i) kw_ += uint32_t(popc32(x[i] & ~S[i])) | (uint32_t(popc32(x[i] & ~S[i] & SP[i])) << 16);
      if (sparse) {
        kw_ = c.t.sum(kw_);
        const float w = c.width_sp(int(kw_ & 0xffffu), int(kw_ >> 16));
        maxw_sp = w > maxw_sp ? w : maxw_sp;
        kw_ = 0;
      }
    }
    if (z < P.n) {
      if (WIDTHS && !sparse) {
        kw_ = c.t.sum(kw_);
        maxk = kw_ > maxk ? kw_ : maxk;
      }
      continue;
    }
    const uint32_t cc_ = c.ch(z);
    const int a = int(cc_ & 0xffffu), b = int(cc_ >> 16);
    uint32_t xa[WPL], xb[WPL];
    c.load_bits(a, xa);
    c.load_bits(b, xb);
    uint32_t kk = c.t.sum(popc_or3<WPL>(xa, xb, S) | (kw_ << 16));
    if (WIDTHS && !sparse) {
      const uint32_t k = kk >> 16;
      maxk = k > maxk ? k : maxk;
      kk &= 0xffffu;
    }
    double cost;
    if (sparse) {
      uint32_t ks = 0;
#pragma unroll
      for (int i = 0; i < WPL; ++i) ks += uint32_t(popc32((xa[i] | xb[i] | S[i]) & SP[i]));
      cost = c.cost_sp(int(kk), int(c.t.sum(ks)));
    } else {
      cost = c.cost_of(int(kk));
    }
    const double pa = dst.pc(a);
    const double pb = dst.pc(b);
    dst.put(z, cost, cost + pa + pb);
    seq += cost;
  }
  maxw = (sparse || gen) ? double(maxw_sp) : P.log2d * double(maxk);
}

// Index sets (and, with HYPER, hyper rows) of all internal nodes from the topology, in post-order -- what
// ContractionTree.__init__ derives from a path (tnco/ctree.py:169-189): inds(z) = inds(a) ^ inds(b), plus every
// shared index whose hyper count is still positive after this contraction.  The counters are lane-private
// (a lane counts the indices of its own words), initialised from the network's hyper counts.
template <int TILE, int WPL, bool HYPER>
TNB_D void build_sets(const ChainView<TILE, WPL>& c) {
  const Params& P = c.P;
  uint16_t* cnt = HYPER ? P.nbig + size_t(c.chain) * P.Ws * 32 : nullptr;
  if (HYPER) {
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      const int w = c.t.tl + k * TILE;
      if (w < P.W)
        for (int b = 0; b < 32; ++b) cnt[w * 32 + b] = P.hcount0[w * 32 + b];
    }
  }
  for (int z = c.po_first(); z >= 0; z = c.po_next(z)) {
    if (z < P.n) continue;
    const uint32_t cc_ = c.ch(z);
    uint32_t xa[WPL], xb[WPL], keep[WPL];
    c.load_bits(int(cc_ & 0xffffu), xa);
    c.load_bits(int(cc_ >> 16), xb);
#pragma unroll
    for (int i = 0; i < WPL; ++i) {
      keep[i] = 0u;
      if (HYPER) {
        const int w = c.t.tl + i * TILE;
        uint32_t v = xa[i] & xb[i];
        while (v) {
          const int b = ctz32(v);
          v &= v - 1;
          if (--cnt[w * 32 + b] > 0) keep[i] |= 1u << b;
        }
      }
      xa[i] = (xa[i] ^ xb[i]) | keep[i];
    }
    c.store_bits(z, xa);
    if (HYPER) c.store_hyp(z, keep);  // == inds(z) & inds(a) & inds(b)
  }
}

TNB_D TNB_INLINE uint32_t mix32(uint32_t x) {  // murmur3 finaliser
  x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
  return x;
}

// bits of word `w` covered by the bit range [idx, idx + g)
TNB_D TNB_INLINE uint32_t span_mask(int idx, int g, int w) {
  const int lo = idx > 32 * w ? idx : 32 * w, hi = idx + g < 32 * w + 32 ? idx + g : 32 * w + 32;
  if (lo >= hi) return 0u;
  const uint32_t ones = hi - lo >= 32 ? 0xffffffffu : ((1u << (hi - lo)) - 1u);
  return ones << (lo - 32 * w);
}

// Greedy slicer (finite_width/greedy/utils.hpp:24-125, skip_slices = nullopt, uniform dims), including
// libstdc++'s std::shuffle / uniform_int_distribution draw pattern so that stream modes stay bit-exact.
template <int TILE, int WPL, class Rng>
TNB_D void get_slices_dev(const ChainView<TILE, WPL>& c, Rng& rng, uint32_t (&S2)[WPL]) {
  const Params& P = c.P;
  const Tile<TILE>& t = c.t;
  uint16_t* nbig = P.nbig + size_t(c.chain) * P.Ws * 32;
  int16_t* pos = P.posbuf + size_t(c.chain) * P.Ws * 32;
#pragma unroll
  for (int k = 0; k < WPL; ++k) {
    S2[k] = 0u;
    const int w = t.tl + k * TILE;
    if (w < P.W)
      for (int b = 0; b < 32; ++b) nbig[w * 32 + b] = 0;
  }
  const bool sparse = P.sparse != nullptr;
  const bool gen = P.gdims != nullptr;
  uint32_t SP[WPL];
#pragma unroll
  for (int i = 0; i < WPL; ++i) SP[i] = 0u;
  if (sparse) c.load_sparse(SP);
  // n_big_tensors (:41-47): every node, leaves included
  for (int z = 0; z < P.N; ++z) {
    uint32_t x[WPL];
    c.load_bits(z, x);
    uint32_t k = 0;
#pragma unroll
    for (int i = 0; i < WPL; ++i) k += uint32_t(popc32(x[i])) | (uint32_t(popc32(x[i] & SP[i])) << 16);
    k = t.sum(k);
    const float wz = gen ? c.template gwidth_model<false>(x, SP)
                         : sparse ? c.width_sp(int(k & 0xffffu), int(k >> 16)) : c.width_of(int(k));
    if (wz > P.max_width) {
#pragma unroll
      for (int i = 0; i < WPL; ++i) {
        const int w = t.tl + i * TILE;
        uint32_t v = x[i];
        while (v) {
          nbig[w * 32 + ctz32(v)]++;
          v &= v - 1;
        }
      }
    }
  }
  t.sync();
  for (int z = c.po_first(); z >= 0; z = c.po_next(z)) {
    uint32_t x[WPL];
    c.load_bits(z, x);
    uint32_t k = 0, ks = 0, kp = 0;  // kp: sparse parts of the whole and of the sliced set
    float wfull = 0.f;
    if (gen) wfull = c.template gwidth_model<false>(x, SP);
#pragma unroll
    for (int i = 0; i < WPL; ++i) {
      k += popc32(x[i]);
      kp += uint32_t(popc32(x[i] & SP[i]));
      x[i] &= ~S2[i];
      ks += popc32(x[i]);
      kp += uint32_t(popc32(x[i] & SP[i])) << 16;
    }
    k = t.sum(k | (ks << 16));
    ks = k >> 16;
    k &= 0xffffu;
    if (sparse) kp = t.sum(kp);
    if (!gen) wfull = sparse ? c.width_sp(int(k), int(kp & 0xffffu)) : c.width_of(int(k));
    if (!(wfull > P.max_width)) continue;
    float sw = gen ? c.template gwidth_model<false>(x, SP)
                   : sparse ? c.width_sp(int(ks), int(kp >> 16)) : c.width_of(int(ks));
    if (!(sw > P.max_width)) continue;
    int kss = int(kp >> 16);  // sparse indices still unsliced on this node
    uint32_t xsp[WPL];        // (general dims) the sparse part of the still unsliced set, kept in step with the picks
#pragma unroll
    for (int i = 0; i < WPL; ++i) xsp[i] = x[i] & SP[i];
    // ascending positions of the still unsliced indices of this node (group leaders when dims differ per index)
    uint32_t np = 0;
#pragma unroll
    for (int i = 0; i < WPL; ++i) {
      const int w = t.tl + i * TILE;
      uint32_t v = x[i];
      if (P.skip) v &= w < P.W ? ~P.skip[w] : 0u;  // sliced_xs - skip_slices
      if (P.grouped) v &= w < P.W ? P.leader[w] : 0u;
      uint32_t tot;
      uint32_t off = np + t.excl_scan_sum(uint32_t(popc32(v)), tot);
      while (v) {
        pos[off++] = int16_t(w * 32 + ctz32(v));
        v &= v - 1;
      }
      np += tot;
    }
    t.sync();
    if (t.tl == 0) {
      auto nd = [&](uint32_t range) -> uint32_t {  // uniform_int_distribution, 32-bit URNG (Lemire)
        unsigned long long prod = (unsigned long long)rng.local_next() * range;
        uint32_t low = uint32_t(prod);
        if (low < range) {
          const uint32_t thr = (0u - range) % range;
          while (low < thr) {
            prod = (unsigned long long)rng.local_next() * range;
            low = uint32_t(prod);
            if (rng.overrun()) break;
          }
        }
        return uint32_t(prod >> 32);
      };
      auto swp = [&](uint32_t a, uint32_t b) {
        const int16_t tmp = pos[a];
        pos[a] = pos[b];
        pos[b] = tmp;
      };
      // std::shuffle (GCC 13 bits/stl_algo.h:3768-3799), two swaps per draw
      uint32_t i = 1;
      if ((np % 2u) == 0u) {
        swp(1, nd(2));
        i = 2;
      }
      while (i != np) {
        const uint32_t r = i + 1;
        const uint32_t xx = nd(r * (r + 1));
        swp(i, xx / (r + 1));
        ++i;
        swp(i, xx % (r + 1));
        ++i;
      }
      // std::stable_sort by n_big_tensors, descending, then (dims vector) by log2 dim, descending (:52-62,:85);
      // insertion sort is stable
      for (uint32_t a = 1; a < np; ++a) {
        const int16_t key = pos[a];
        const uint16_t kb = nbig[key];
        const int gk = P.grouped ? int(P.gw[key]) : 0;
        int b = int(a) - 1;
        // (general dims: second key log2 dim as float, DimsCache<width_type>, :52-62)
        const float lk = gen ? float(P.glog2[key]) : 0.f;
        while (b >= 0 && (kb > nbig[pos[b]] || (P.grouped && kb == nbig[pos[b]] && gk > int(P.gw[pos[b]])) ||
                          (gen && kb == nbig[pos[b]] && lk > float(P.glog2[pos[b]])))) {
          pos[b + 1] = pos[b];
          --b;
        }
        pos[b + 1] = key;
      }
    }
    rng.sync_from0(t);
    t.sync();
    uint32_t m = 0;
    const float dw = float(-P.log2d);  // get_delta_width for a present index (fw simple.hpp:60-76)
    while (m < np) {
      const int idx = pos[m];
      if (gen) {
        if (sparse && ((P.sparse[idx >> 5] >> (idx & 31)) & 1u)) {
          const float wo = c.cap_w(c.template gwidth<false>(xsp));
#pragma unroll
          for (int i = 0; i < WPL; ++i) xsp[i] &= ~span_mask(idx, 1, t.tl + i * TILE);
          sw += c.cap_w(c.template gwidth<false>(xsp)) - wo;
        } else {
          sw += float(-P.glog2[idx]);  // (1 - 2*test(pos)) * log2(dims[pos]) as float, :60-76
        }
      } else if (sparse && ((P.sparse[idx >> 5] >> (idx & 31)) & 1u)) {
        // get_delta_width of a sparse index (fw simple_sparse_inds.hpp:51-77): difference of the capped widths
        const int g = P.grouped ? int(P.gw[idx]) : 1;
        const float wo = c.width_of(kss), wn = c.width_of(kss - g);
        const float L = float(P.log2_n_projs);
        sw += (double(wn) < P.log2_n_projs ? wn : L) - (double(wo) < P.log2_n_projs ? wo : L);
        kss -= g;
      } else {
        sw += P.grouped ? -float(int(P.gw[idx])) : dw;
      }
      ++m;
      if (sw <= P.max_width) break;
    }
    for (uint32_t j = 0; j < m; ++j) {
      const int idx = pos[j];
      const int g = P.grouped ? int(P.gw[idx]) : 1;
#pragma unroll
      for (int i = 0; i < WPL; ++i) S2[i] |= span_mask(idx, g, t.tl + i * TILE);
    }
    t.sync();
  }
}

// Production re-slicer: the same greedy rule as get_slices_dev / finite_width/greedy/utils.hpp:24-125 (walk the
// wide nodes in post-order; while a node is still too wide, slice the index that sits on the most wide nodes,
// ties broken uniformly at random -- which is what std::shuffle followed by a stable sort yields), restructured
// so that nothing walks the whole tree:
//   * wide nodes come from a scan of the per-node popcounts kw[], which the sweep keeps up to date;
//   * their post-order ranks come from the subtree leaf counts sz[] (also kept by the sweep): the nodes visited
//     before z are z's own subtree plus the left-sibling subtree of every right turn on the way to the root;
//   * the per-index counters are NB bit planes in registers (lane-private: a lane counts the 32 indices of its
//     own word), added with a ripple carry and saturating, so "largest count among the candidates" is NB votes.
// draw0: the re-slice's tie-break word, drawn by the caller (passing the generator itself by reference to this
// out-of-line function forced all of its state through local memory -- the sweep loop then read its level word back
// from the stack at every level).
// SPARSE: the sparse-index width model (finite_width/cost_model/simple_sparse_inds.hpp:39-77): width = w(dense part) +
// min(w(sparse part), log2 n_projs), so a node is described by two popcounts, slicing a sparse index above the cap
// does not narrow the node (the reference slices it all the same: candidates are taken in count order until the node
// fits), and candidates go one at a time.  A separate instantiation: the 2^popcount kernels keep the code they had.
template <int TILE, int WPL, bool SPARSE>
TNB_D TNB_NOINLINE void get_slices_fast(const ChainView<TILE, WPL>& c, const uint32_t draw0, uint32_t (&S2)[WPL]) {
  const Params& P = c.P;
  const Tile<TILE>& t = c.t;
  constexpr int NB = 8;
  const uint32_t* kwsz = P.kwsz + size_t(c.chain) * P.Npad;
  const int16_t* ksp = SPARSE ? P.ksp + size_t(c.chain) * P.Npad : nullptr;
  uint32_t SP[WPL];
#pragma unroll
  for (int k = 0; k < WPL; ++k) SP[k] = 0u;
  if (SPARSE) c.load_sparse(SP);
  uint32_t* wkey = P.wkey + size_t(c.chain) * P.Npad;
  int16_t* word = P.word + size_t(c.chain) * P.Npad;
  uint32_t cnt[WPL][NB];
#pragma unroll
  for (int k = 0; k < WPL; ++k) {
    S2[k] = 0u;
#pragma unroll
    for (int b = 0; b < NB; ++b) cnt[k][b] = 0u;
  }
  // (1) wide nodes (leaves included, :41-47): list them and count, per index, how many contain it
  int nw = 0;
  for (int base = 0; base < P.N; base += 4 * TILE) {  // four independent loads in flight per round trip
    int kv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int z = base + u * TILE + t.tl;
      kv[u] = z < P.N ? int(kwsz[z] & 0xffffu) : -1;
      if (SPARSE && z < P.N) kv[u] |= int(ksp[z]) << 16;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool wide = SPARSE ? (kv[u] >= 0 && c.width_sp(kv[u] & 0xffff, kv[u] >> 16) > P.max_width) : kv[u] > P.kthr;
      const uint32_t m = t.ballot(wide);
      if (wide) word[nw + popc32(m & ((1u << t.tl) - 1u))] = int16_t(base + u * TILE + t.tl);
      nw += popc32(m);
    }
  }
  if (nw == 0) return;
  t.sync();
  // The rows of the wide nodes are read twice below, one after the other (counting, then selection in post-order):
  // ask for all of them now, one row per lane, so that the sequential passes find them in L2 instead of paying one
  // HBM round trip per row.
  for (int j = t.tl; j < nw; j += TILE) c.prefetch_row(word[j]);
  {
    uint32_t xn[WPL];
    c.load_bits(word[0], xn);
    for (int j = 0; j < nw; ++j) {
      uint32_t x[WPL];
#pragma unroll
      for (int k = 0; k < WPL; ++k) x[k] = xn[k];
      if (j + 1 < nw) c.load_bits(word[j + 1], xn);  // the next row is in flight while this one is counted
#pragma unroll
      for (int k = 0; k < WPL; ++k) {
        uint32_t carry = x[k];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const uint32_t nc = cnt[k][b] & carry;
          cnt[k][b] ^= carry;
          carry = nc;
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) cnt[k][b] |= carry;  // saturate at 2^NB - 1
      }
    }
  }
  // (2) post-order rank of every wide node, one node per lane
  for (int j = t.tl; j < nw; j += TILE) {
    const int z = word[j];
    uint32_t post = 2u * (kwsz[z] >> 16) - 2u;
    int y = z;
    while (true) {
      const int p = c.par[y];
      if (p < 0) break;
      const int l = int(c.ch(p) & 0xffffu);
      if (l != y) post += 2u * (kwsz[l] >> 16) - 1u;
      y = p;
    }
    wkey[j] = (post << 16) | uint32_t(z);
  }
  t.sync();
  // rank sort (keys are distinct): word[] <- wide nodes in post-order
  for (int j = t.tl; j < nw; j += TILE) {
    const uint32_t key = wkey[j];
    int rank = 0;
    for (int k = 0; k < nw; ++k) rank += wkey[k] < key ? 1 : 0;
    word[rank] = int16_t(key & 0xffffu);
  }
  t.sync();
  // (3) greedy selection (:60-104)
  uint32_t n_draw = 0;
  uint32_t xnext[WPL];
  c.load_bits(word[0], xnext);
  for (int j = 0; j < nw; ++j) {
    uint32_t x[WPL];
#pragma unroll
    for (int k = 0; k < WPL; ++k) x[k] = xnext[k];
    if (j + 1 < nw) c.load_bits(word[j + 1], xnext);
    uint32_t ks = 0;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      x[k] &= ~S2[k];
      ks += uint32_t(popc32(x[k]));
      if (SPARSE) ks += uint32_t(popc32(x[k] & SP[k])) << 16;
    }
    ks = t.sum(ks);
    int k_all = int(ks & 0xffffu), k_sp = int(ks >> 16);  // SPARSE: all / sparse indices left on this node
    // m = binary indices still to be removed from this node (SPARSE: 1 while it is too wide, one candidate at a time)
    for (int m = SPARSE ? (c.width_sp(k_all, k_sp) > P.max_width ? 1 : 0) : int(ks) - P.kthr; m > 0;) {
      uint32_t cand[WPL];
#pragma unroll
      for (int k = 0; k < WPL; ++k) {
        cand[k] = x[k];
        if (P.grouped) {  // one candidate per index: its leader bit
          const int w = t.tl + k * TILE;
          cand[k] &= w < P.W ? P.leader[w] : 0u;
        }
      }
#pragma unroll
      for (int b = NB - 1; b >= 0; --b) {
        bool hit = false;
#pragma unroll
        for (int k = 0; k < WPL; ++k) hit |= (cand[k] & cnt[k][b]) != 0u;
        if (t.any(hit)) {
#pragma unroll
          for (int k = 0; k < WPL; ++k) cand[k] &= cnt[k][b];
        }
      }
      if (P.grouped) {  // second sort key: the largest dimension among the most frequent (:57-60)
        uint32_t gmax = 0;
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
          const int w = t.tl + k * TILE;
          for (uint32_t v = cand[k]; v; v &= v - 1) {
            const uint32_t g = P.gw[w * 32 + ctz32(v)];
            gmax = g > gmax ? g : gmax;
          }
        }
        gmax = t.max_u32(gmax);
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
          const int w = t.tl + k * TILE;
          for (uint32_t v = cand[k]; v; v &= v - 1)
            if (P.gw[w * 32 + ctz32(v)] != gmax) cand[k] &= ~(1u << ctz32(v));
        }
      }
      uint32_t mine = 0;
#pragma unroll
      for (int k = 0; k < WPL; ++k) mine += uint32_t(popc32(cand[k]));
      if (!P.grouped) {
        // Every candidate of the most frequent class goes anyway when the class is not larger than what is still
        // to be removed (the order inside a class is only the random tie-break): take the class in one step.
        const uint32_t all = t.sum(mine);
        if (!SPARSE && all != 0u && int(all) <= m) {
#pragma unroll
          for (int k = 0; k < WPL; ++k) {
            S2[k] |= cand[k];
            x[k] &= ~cand[k];
          }
          m -= int(all);
          continue;
        }
        if (SPARSE && all == 0u) break;  // (nothing left to slice on this node)
      }
      uint32_t tot;
      const uint32_t off = t.excl_scan_sum(mine, tot);
      // tie-break draw: one Philox word per re-slice, hashed with the pick number (every lane computes the same)
      uint32_t r = mulhi32(mix32(draw0 + 0x9e3779b9u * ++n_draw), tot);
      bool done = !(r >= off && r < off + mine);
      r -= off;
      uint32_t picked = 0;  // 1 + position of the chosen bit (on the lane that owns it)
#pragma unroll
      for (int k = 0; k < WPL; ++k) {
        const uint32_t pk = uint32_t(popc32(cand[k]));
        if (!done && r < pk) {
          const int bpos = nth_set_bit(cand[k], r);
          if (P.grouped) {
            picked = uint32_t((t.tl + k * TILE) * 32 + bpos) + 1u;
          } else {
            S2[k] |= 1u << bpos;
            x[k] &= ~(1u << bpos);
            if (SPARSE) picked = 1u + ((SP[k] >> bpos) & 1u);  // 2: the sliced index is a sparse one
          }
          done = true;
        }
        r -= pk;
      }
      if (P.grouped) {  // the whole index goes: every lane clears its share of the group
        picked = t.max_u32(picked);
        if (picked == 0u) break;  // (no candidate left; cannot happen while m > 0)
        const int idx = int(picked - 1u), g = int(P.gw[idx]);
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
          const uint32_t msk = span_mask(idx, g, t.tl + k * TILE);
          S2[k] |= msk;
          x[k] &= ~msk;
        }
        m -= g;
      } else if (SPARSE) {
        --k_all;
        if (t.max_u32(picked) == 2u) --k_sp;
        m = c.width_sp(k_all, k_sp) > P.max_width ? 1 : 0;
      } else {
        --m;
      }
    }
  }
}

// Contraction costs of every internal node under the slices S2, into dst[].x; returns their sum.  No traversal
// order is needed when only the total matters (production mode keeps no partial-cost cache).
template <int TILE, int WPL, bool DIM2>
TNB_D TNB_NOINLINE double recost_all(const ChainView<TILE, WPL>& c, const uint32_t (&S2)[WPL], dbl2* dst) {
  const Params& P = c.P;
  const Tile<TILE>& t = c.t;
  double acc = 0.0;
  const bool sparse = !DIM2 && P.sparse != nullptr;
  uint32_t SP[WPL];
#pragma unroll
  for (int i = 0; i < WPL; ++i) SP[i] = 0u;
  if (sparse) c.load_sparse(SP);
  for (int base = P.n; base < P.N; base += TILE) {
    const int zz = base + t.tl;
    const uint32_t mych = zz < P.N ? c.ch(zz) : 0u;
    const int cnt = P.N - base < TILE ? P.N - base : TILE;
    for (int q = 0; q < cnt; ++q) {
      const uint32_t w = t.bcast(mych, q);
      uint32_t xa[WPL], xb[WPL];
      c.load_bits(int(w & 0xffffu), xa);
      c.load_bits(int(w >> 16), xb);
      uint32_t k = popc_or3<WPL>(xa, xb, S2);
      if (sparse) {  // (kernel-uniform) the sparse part of the same union rides in the high half
#pragma unroll
        for (int i = 0; i < WPL; ++i) k += uint32_t(popc32((xa[i] | xb[i] | S2[i]) & SP[i])) << 16;
      }
      k = t.sum(k);
      const double cost = DIM2 ? bits_to_f64((unsigned long long)(1023u + (k > 1024u ? 1024u : k)) << 52)
                          : sparse ? c.cost_sp(int(k & 0xffffu), int(k >> 16)) : c.cost_of(int(k));
      dst[base - P.n + q].x = cost;
      acc += cost;
    }
  }
  return acc;
}

// Re-cost under new slices WITHOUT touching the index sets (dim == 2, where a contraction cost is the exact power
// 2^popc(U_z | S), U_z = union of the children's index sets).  Going from S to S2 multiplies the cost of node z by
//   2^(|S2 \ S| - |S \ S2| + #{i in S \ S2 : i in U_z} - #{i in S2 \ S : i in U_z}),
// and i lies in U_z exactly for the nodes on the two paths from the leaves holding i up to the node where i is
// contracted.  mark_slice_diff walks those paths for every index of the symmetric difference (one index per lane)
// and leaves the per-node exponent correction in dz[]; returns the common shift |S2 \ S| - |S \ S2|.
template <int TILE, int WPL>
TNB_D TNB_NOINLINE int mark_slice_diff(const ChainView<TILE, WPL>& c, const uint32_t (&S)[WPL],
                                       const uint32_t (&S2)[WPL], int* dz, int16_t* list) {
  const Params& P = c.P;
  const Tile<TILE>& t = c.t;
  for (int i = t.tl; i < P.n_int; i += TILE) dz[i] = 0;
  uint32_t mine = 0, n_add = 0;
#pragma unroll
  for (int k = 0; k < WPL; ++k) {
    mine += uint32_t(popc32(S[k] ^ S2[k]));
    n_add += uint32_t(popc32(S2[k] & ~S[k]));
  }
  uint32_t nd;
  uint32_t off = t.excl_scan_sum(mine, nd);
  n_add = t.sum(n_add);
#pragma unroll
  for (int k = 0; k < WPL; ++k) {
    uint32_t v = S[k] ^ S2[k];
    const int w = t.tl + k * TILE;
    while (v) {
      const int b = ctz32(v);
      v &= v - 1;
      // bit 15: the index leaves the slice set (its nodes get +1), otherwise it joins (-1)
      list[off++] = int16_t((w * 32 + b) | (((S[k] >> b) & 1u) ? 0x8000 : 0));
    }
  }
  t.sync();
  // The two leaves holding idx climb towards each other, the one with the smaller subtree first (an ancestor always
  // has the larger leaf count), and meet in the node that contracts idx; every node entered on the way has idx in
  // one of its children.  Only par[] and sz[] are read: small arrays that stay in L2 even when the index sets
  // live in HBM (testing the bit of idx in every node's index set cost one HBM round trip per step).
  const uint32_t* kwsz = P.kwsz + size_t(c.chain) * P.Npad;
  for (int j = t.tl; j < int(nd); j += TILE) {
    const int e = uint16_t(list[j]);
    const int idx = e & 0x7fff, sgn = (e & 0x8000) ? 1 : -1;
    int a = P.net_own[idx], b = P.net_own[P.n_inds + idx];
    auto mark = [&](int z) {
#if defined(TNB_EMU)
      dz[z - c.n] += sgn;
#else
      atomicAdd(dz + (z - c.n), sgn);
#endif
    };
    if (b < 0) {  // open index: it stays in every node up to the root
      while (true) {
        a = c.par[a];
        if (a < 0) break;
        mark(a);
      }
      continue;
    }
    int sa = 1, sb = 1;
    while (a != b) {
      if (sa <= sb) {
        a = c.par[a];
        sa = int(kwsz[a] >> 16);
        if (a != b) mark(a);
      } else {
        b = c.par[b];
        sb = int(kwsz[b] >> 16);
        if (a != b) mark(b);
      }
    }
    // (the contracting node was marked by whichever walker reached it first, exactly once)
  }
  t.sync();
  return 2 * int(n_add) - int(nd);  // |S2 \ S| - |S \ S2|
}

// cost * 2^e for a cost that is an exact power of two
TNB_D TNB_INLINE double scale_pow2(double cost, int e) {
#if defined(TNB_EMU)
  return std::ldexp(cost, e);
#else
  return __longlong_as_double(__double_as_longlong(cost) + ((long long)e << 52));
#endif
}

// Sum of the shifted costs; with STORE the shifted costs also replace the cached ones.
template <int TILE, int WPL, bool STORE>
TNB_D double sum_shifted(const ChainView<TILE, WPL>& c, const int* dz, int shift0) {
  double acc = 0.0;
  for (int z = c.n + c.t.tl; z < c.P.N; z += TILE) {
    const double v = scale_pow2(c.cc_get(z), shift0 + dz[z - c.n]);
    if (STORE) c.cc_set(z, v);
    acc += v;
  }
#if !defined(TNB_EMU)
#pragma unroll
  for (int d = TILE / 2; d > 0; d >>= 1)
    acc += TILE == 32 ? __shfl_xor_sync(0xffffffffu, acc, d, 32) : __shfl_xor_sync(c.t.mask, acc, d, TILE);
#endif
  return acc;
}

// kw[] / sz[] of every node from scratch (construction time).
template <int TILE, int WPL>
TNB_D void build_kw_sz(const ChainView<TILE, WPL>& c) {
  const Params& P = c.P;
  uint32_t* kwsz = P.kwsz + size_t(c.chain) * P.Npad;
  uint32_t SP[WPL];
#pragma unroll
  for (int i = 0; i < WPL; ++i) SP[i] = 0u;
  if (P.ksp) c.load_sparse(SP);
  for (int z = c.po_first(); z >= 0; z = c.po_next(z)) {
    uint32_t x[WPL];
    c.load_bits(z, x);
    uint32_t k = 0;
#pragma unroll
    for (int i = 0; i < WPL; ++i) k += uint32_t(popc32(x[i]));
    if (P.ksp) {
#pragma unroll
      for (int i = 0; i < WPL; ++i) k += uint32_t(popc32(x[i] & SP[i])) << 16;
    }
    k = c.t.sum(k);
    if (P.ksp) {
      P.ksp[size_t(c.chain) * P.Npad + z] = int16_t(k >> 16);
      k &= 0xffffu;
    }
    if (z < P.n) {
      kwsz[z] = k | (1u << 16);
    } else {
      const uint32_t w = c.ch(z);
      kwsz[z] = k | (((kwsz[w & 0xffffu] >> 16) + (kwsz[w >> 16] >> 16)) << 16);
    }
  }
}

template <int TILE, int WPL>
TNB_D void snapshot_best(const ChainView<TILE, WPL>& c, const uint32_t (&S)[WPL], bool finite) {
  const Params& P = c.P;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(c.par);
  uint32_t* dst = reinterpret_cast<uint32_t*>(P.bpar + size_t(c.chain) * P.Npad);
  c.t.sync();  // (see snapshot_walk)
  // (unrolled: the copies are independent, several loads in flight instead of one load-store round trip at a time)
#pragma unroll 8
  for (int i = c.t.tl; i < P.Npad / 2; i += TILE) dst[i] = src[i];
  uint32_t* dch = P.bch + size_t(c.chain) * P.n_int;
#pragma unroll 8
  for (int i = c.t.tl; i < P.n_int; i += TILE) dch[i] = c.ch(P.n + i);
  if (finite) {
    uint32_t* ds = P.bslices + size_t(c.chain) * P.Ws;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      const int w = c.t.tl + k * TILE;
      if (w < P.W) ds[w] = S[k];
    }
  }
}

// Incremental form of snapshot_best for a sweep that STARTED from the best tree (the previous sweep ended with a
// snapshot): the sweep changed the children words of the nodes it walked (the B of every level, and the root) and
// re-parented children of those nodes only -- a node's final parent is the last walked node that adopted it.  So the
// snapshot is brought up to date by one pass over the walked nodes, one per lane: its children word, and itself as the
// parent of both children.  (C5: 2 x 1999 words copied per snapshot, at almost every sweep of the descent phase --
// half of all stall samples of a short anneal -- against ~22 nodes here.)
template <int TILE, int WPL, class Rng>
TNB_D void snapshot_walk(const ChainView<TILE, WPL>& c, const Rng& rng, uint32_t slot_start, int levels,
                         const uint32_t (&S)[WPL], bool finite) {
  const Params& P = c.P;
  int16_t* bpar = P.bpar + size_t(c.chain) * P.Npad;
  uint32_t* dch = P.bch + size_t(c.chain) * P.n_int;
  c.t.sync();  // (the two snapshot forms write the same words from different lanes: keep them ordered)
  for (int i = c.t.tl; i <= levels; i += TILE) {
    const int X = i < levels ? int(rng.ring_get(slot_start + 1u + uint32_t(i))) : P.N - 1;
    const uint32_t w = c.ch(X);
    dch[X - P.n] = w;
    bpar[w & 0xffffu] = int16_t(X);
    bpar[w >> 16] = int16_t(X);
  }
  if (finite) {
    uint32_t* ds = P.bslices + size_t(c.chain) * P.Ws;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      const int w = c.t.tl + k * TILE;
      if (w < P.W) ds[w] = S[k];
    }
  }
}

template <int TILE, int WPL>
TNB_D void load_slices(const ChainView<TILE, WPL>& c, uint32_t (&S)[WPL]) {
  const uint32_t* s = c.P.slices + size_t(c.chain) * c.P.Ws;
#pragma unroll
  for (int k = 0; k < WPL; ++k) {
    const int w = c.t.tl + k * TILE;
    S[k] = w < c.P.W ? s[w] : 0u;
  }
}
template <int TILE, int WPL>
TNB_D void store_slices(const ChainView<TILE, WPL>& c, const uint32_t (&S)[WPL]) {
  uint32_t* s = c.P.slices + size_t(c.chain) * c.P.Ws;
#pragma unroll
  for (int k = 0; k < WPL; ++k) {
    const int w = c.t.tl + k * TILE;
    if (w < c.P.W) s[w] = S[k];
  }
}

// ------------------------------------------------------------------------------------------ initial trees


// One initial contraction tree per chain, built by the chain's own tile (replaces the host-side
// get_random_contraction_path + ContractionTree construction, tnco/utils/tn.py:109-273, tnco/ctree.py:108-226, for
// hyper-free connected networks).  n-1 merge steps; at every step each lane scores a strided share of the live
// edges (an index whose two owning clusters differ), the tile takes the arg-min, merges the two clusters
// (index set = XOR), and re-points the owners of the surviving indices.  GREEDY scores an edge like opt_einsum's
// greedy, size(out) - size(a) - size(b), ties broken by a per-(seed, step, edge) hash; RANDOM uses the hash only.
// Every merge is along a shared index, so check_shared_inds holds by construction.
// The same for networks WITH hyper-indices (an index on 3+ tensors, or on 2 and open): there is no "edge = index with
// two owners" any more, so every step scores all pairs of live clusters that share an index (strided over the lanes;
// O(n^2 W / TILE) per step -- these networks are small), with the index set of a contraction given by the hyper-count
// rule of tnco/ctree.py:169-189: a shared index survives while other tensors (or the output) still hold it.
template <int TILE, int WPL>
TNB_D void chain_treegen_hyper(const Params& P, int chain) {
  ChainView<TILE, WPL> c(P, chain);
  const Tile<TILE>& t = c.t;
  const int n = P.n, W = P.W;
  uint16_t* cnt = P.nbig + size_t(chain) * P.Ws * 32;  // remaining hyper count of every index
  int16_t* kpop = P.kpop + size_t(chain) * P.Npad;
  int16_t* live = reinterpret_cast<int16_t*>(P.escore + size_t(chain) * P.Ws * 32);  // live clusters (4*Ws*32 slots)
  const unsigned long long seed = P.seeds[chain];
  const uint32_t s0 = mix32(uint32_t(seed) ^ 0x9e3779b9u), s1 = mix32(uint32_t(seed >> 32) + 0x7f4a7c15u);
#pragma unroll
  for (int k = 0; k < WPL; ++k) {
    const int w = t.tl + k * TILE;
    if (w < W)
      for (int b = 0; b < 32; ++b) cnt[w * 32 + b] = P.hcount0[w * 32 + b];
  }
  for (int x = t.tl; x < P.N; x += TILE) {
    int k = 0;
    if (x < n)
      for (int w = 0; w < W; ++w) k += popc32(P.leaf_bits[size_t(x) * P.Ws + w]);
    kpop[x] = int16_t(k);
    c.par[x] = int16_t(-1);
    if (x < n) live[x] = int16_t(x);
  }
  t.sync();
  auto row = [&](int x) -> const uint32_t* {
    return x < n ? P.leaf_bits + size_t(x) * P.Ws
                 : reinterpret_cast<const uint32_t*>(c.rec_lane - 4 * t.tl + unsigned(x) * c.bstride);
  };
  auto pow2 = [](int k) { return bits_to_f64((unsigned long long)(1023 + (k > 1000 ? 1000 : k)) << 52); };
  int nl = n;
  for (int step = 0; step < n - 1; ++step) {
    const int z = n + step;
    double best = 1.0e308;
    uint32_t best_tie = 0xffffffffu;
    int best_i = -1, best_j = -1;
    const uint32_t hs = mix32(s0 + uint32_t(step) * 0x632be5abu);
    for (int i = 0; i + 1 < nl; ++i) {
      const int a = live[i];
      const uint32_t* ra = row(a);
      for (int j = i + 1 + t.tl; j < nl; j += TILE) {
        const int b = live[j];
        const uint32_t* rb = row(b);
        int ko = 0;
        bool shared = false;
        for (int w = 0; w < W; ++w) {
          const uint32_t xa = ra[w], xb = rb[w];
          ko += popc32(xa ^ xb);
          uint32_t v = xa & xb;
          shared |= v != 0u;
          while (v) {  // a shared index stays on the result while somebody else still holds it
            ko += int(cnt[w * 32 + ctz32(v)]) - 1 > 0 ? 1 : 0;
            v &= v - 1;
          }
        }
        if (!shared) continue;
        const double sc = P.tree_method != 0 ? 0.0 : pow2(ko) - pow2(kpop[a]) - pow2(kpop[b]);
        const uint32_t tie = mix32(hs ^ (uint32_t(a) * 0x9e3779b1u) ^ (uint32_t(b) * 0x85ebca77u) ^ s1);
        if (sc < best || (sc == best && tie < best_tie)) {
          best = sc;
          best_tie = tie;
          best_i = i;
          best_j = j;
        }
      }
    }
#if !defined(TNB_EMU)
#pragma unroll
    for (int d = TILE / 2; d > 0; d >>= 1) {
      const unsigned m = TILE == 32 ? 0xffffffffu : t.mask;
      const double os = __shfl_xor_sync(m, best, d, TILE);
      const uint32_t ot = __shfl_xor_sync(m, best_tie, d, TILE);
      const int oi = __shfl_xor_sync(m, best_i, d, TILE), oj = __shfl_xor_sync(m, best_j, d, TILE);
      const bool take = oi >= 0 && (best_i < 0 || os < best || (os == best && (ot < best_tie ||
                        (ot == best_tie && (oi < best_i || (oi == best_i && oj < best_j))))));
      if (take) { best = os; best_tie = ot; best_i = oi; best_j = oj; }
    }
#endif
    if (best_i < 0) {  // no two live clusters share an index: the network is disconnected
      P.tree_fail[chain] = 1;
      return;
    }
    int a = live[best_i], b = live[best_j];
    if (best_tie & 1u) { const int tmp = a; a = b; b = tmp; }
    uint32_t xa[WPL], xb[WPL], xz[WPL];
    c.load_bits(a, xa);
    c.load_bits(b, xb);
    uint32_t k = 0;
#pragma unroll
    for (int q = 0; q < WPL; ++q) {
      const int w = t.tl + q * TILE;
      uint32_t keep = 0u, v = xa[q] & xb[q];
      while (v) {
        const int bit = ctz32(v);
        v &= v - 1;
        if (--cnt[w * 32 + bit] > 0) keep |= 1u << bit;
      }
      xz[q] = (xa[q] ^ xb[q]) | keep;
      k += uint32_t(popc32(xz[q]));
    }
    c.store_bits(z, xz);
    k = t.sum(k);
    t.sync();  // everybody has read live[best_i], live[best_j]
    kpop[z] = int16_t(k);
    c.par[a] = int16_t(z);
    c.par[b] = int16_t(z);
    c.ch(z) = uint32_t(a) | (uint32_t(b) << 16);
    live[best_i] = int16_t(z);  // (best_i < best_j <= nl - 1)
    live[best_j] = live[nl - 1];
    --nl;
    t.sync();
  }
}

template <int TILE, int WPL>
TNB_D void chain_treegen(const Params& P, int chain) {
  if (P.hyper) {
    chain_treegen_hyper<TILE, WPL>(P, chain);
    return;
  }
  ChainView<TILE, WPL> c(P, chain);
  const Tile<TILE>& t = c.t;
  const int n = P.n, N = P.N, W = P.W;
  int16_t* own0 = reinterpret_cast<int16_t*>(P.nbig + size_t(chain) * P.Ws * 32);
  int16_t* own1 = P.posbuf + size_t(chain) * P.Ws * 32;
  int16_t* kpop = P.kpop + size_t(chain) * P.Npad;
  double* escore = P.escore + size_t(chain) * P.Ws * 32;  // cached score of every live edge, +inf otherwise
  const double kDead = 1.0e308;
  const unsigned long long seed = P.seeds[chain];
  const uint32_t s0 = mix32(uint32_t(seed) ^ 0x9e3779b9u), s1 = mix32(uint32_t(seed >> 32) + 0x7f4a7c15u);
  for (int i = t.tl; i < P.n_inds; i += TILE) {
    own0[i] = P.net_own[i];
    own1[i] = P.net_own[P.n_inds + i];
  }
  for (int x = t.tl; x < N; x += TILE) {
    int k = 0;
    if (x < n)
      for (int w = 0; w < W; ++w) k += popc32(P.leaf_bits[size_t(x) * P.Ws + w]);
    kpop[x] = int16_t(k);
    c.par[x] = int16_t(-1);
  }
  t.sync();
  auto row = [&](int x) -> const uint32_t* {
    return x < n ? P.leaf_bits + size_t(x) * P.Ws
                 : reinterpret_cast<const uint32_t*>(c.rec_lane - 4 * t.tl + unsigned(x) * c.bstride);
  };
  auto pow2 = [](int k) { return bits_to_f64((unsigned long long)(1023 + (k > 1000 ? 1000 : k)) << 52); };
  // score of contracting clusters a and b along a shared index (opt_einsum greedy: size(out) - size(a) - size(b))
  auto score = [&](int a, int b) -> double {
    if (P.tree_method != 0) return 0.0;
    const uint32_t *ra = row(a), *rb = row(b);
    int ko = 0;
    for (int w = 0; w < W; ++w) ko += popc32(ra[w] ^ rb[w]);
    return pow2(ko) - pow2(kpop[a]) - pow2(kpop[b]);
  };
  // every edge is scored once here; afterwards only the edges of the freshly merged cluster are re-scored
  for (int i = t.tl; i < P.n_inds; i += TILE) {
    const int a = own0[i], b = own1[i];
    escore[i] = (a < 0 || b < 0 || a == b) ? kDead : score(a, b);
  }
  t.sync();
  for (int step = 0; step < n - 1; ++step) {
    const int z = n + step;
    double best = kDead;
    uint32_t best_tie = 0xffffffffu;
    int best_i = -1;
    const uint32_t hs = mix32(s0 + uint32_t(step) * 0x632be5abu);
    for (int i = t.tl; i < P.n_inds; i += TILE) {
      const double sc = escore[i];
      if (!(sc < kDead)) continue;
      const uint32_t tie = mix32(hs ^ (uint32_t(i) * 0x9e3779b1u) ^ s1);
      if (sc < best || (sc == best && tie < best_tie)) {
        best = sc;
        best_tie = tie;
        best_i = i;
      }
    }
    // tile arg-min over (score, tie); lanes without a candidate carry best_i = -1
#if !defined(TNB_EMU)
#pragma unroll
    for (int d = TILE / 2; d > 0; d >>= 1) {
      const unsigned m = TILE == 32 ? 0xffffffffu : t.mask;
      const double os = __shfl_xor_sync(m, best, d, TILE);
      const uint32_t ot = __shfl_xor_sync(m, best_tie, d, TILE);
      const int oi = __shfl_xor_sync(m, best_i, d, TILE);
      const bool take = oi >= 0 && (best_i < 0 || os < best || (os == best && (ot < best_tie || (ot == best_tie && oi < best_i))));
      if (take) { best = os; best_tie = ot; best_i = oi; }
    }
#endif
    if (best_i < 0) {  // no live edge left: the network is disconnected
      P.tree_fail[chain] = 1;
      return;
    }
    int a = own0[best_i], b = own1[best_i];
    if (best_tie & 1u) { const int tmp = a; a = b; b = tmp; }
    // merge: index set of z, popcount, topology
    uint32_t xa[WPL], xb[WPL], xz[WPL];
    c.load_bits(a, xa);
    c.load_bits(b, xb);
    uint32_t k = 0;
#pragma unroll
    for (int q = 0; q < WPL; ++q) {
      xz[q] = xa[q] ^ xb[q];
      k += uint32_t(popc32(xz[q]));
      // owners: surviving indices now belong to z, contracted ones (in both) die
      const int w = t.tl + q * TILE;
      uint32_t v = xz[q];
      while (v) {
        const int idx = w * 32 + ctz32(v);
        v &= v - 1;
        if (own0[idx] == a || own0[idx] == b) own0[idx] = int16_t(z);
        if (own1[idx] == a || own1[idx] == b) own1[idx] = int16_t(z);
      }
      v = xa[q] & xb[q];
      while (v) {
        const int idx = w * 32 + ctz32(v);
        v &= v - 1;
        own0[idx] = int16_t(-1);
        own1[idx] = int16_t(-1);
        escore[idx] = kDead;
      }
    }
    c.store_bits(z, xz);
    k = t.sum(k);
    kpop[z] = int16_t(k);
    c.par[a] = int16_t(z);
    c.par[b] = int16_t(z);
    c.ch(z) = uint32_t(a) | (uint32_t(b) << 16);
    t.sync();  // the row of z (written word by word by its owners) is read whole below
    // re-score the edges of z: each lane takes the surviving indices of its own words
#pragma unroll
    for (int q = 0; q < WPL; ++q) {
      const int w = t.tl + q * TILE;
      uint32_t v = xz[q];
      while (v) {
        const int idx = w * 32 + ctz32(v);
        v &= v - 1;
        const int o0 = own0[idx], o1 = own1[idx];
        const int other = o0 == z ? o1 : o0;
        escore[idx] = (other < 0 || other == z) ? kDead : score(z, other);
      }
    }
    t.sync();  // owners / scores written by other lanes are read by everybody in the next step
  }
}

// ------------------------------------------------------------------------------------------ construction
template <int TILE, int WPL, bool FINITE, class Rng>
TNB_D void chain_init(const Params& P, int chain) {
  ChainView<TILE, WPL> c(P, chain);
  uint32_t S[WPL];
#pragma unroll
  for (int k = 0; k < WPL; ++k) S[k] = 0u;
  if (P.n_int == 0) {  // single tensor: nothing to contract
    P.total[chain] = 0.0;
    P.min_total[chain] = 0.0;
    if (P.out_seq) P.out_seq[chain] = 0.0;
    if (P.out_maxw) P.out_maxw[chain] = 0.0;
    return;
  }
  if (P.hyper) build_sets<TILE, WPL, true>(c);
  else build_sets<TILE, WPL, false>(c);
  if (FINITE) {
    if (P.slices_given) {
      load_slices(c, S);
    } else {
      // reference ctor order: seed PRNG -> WidthCache -> slices (consumes the PRNG) -> CostCache
      Rng rng;
      rng.load(P, chain);
      if constexpr (Rng::kFast) {
        if (P.kwsz) {  // production: the same greedy rule through the fast slicer
          build_kw_sz(c);
          const uint32_t draw0 = rng.local_next();
          rng.sync_from0(c.t);
          if (P.ksp) get_slices_fast<TILE, WPL, true>(c, draw0, S);
          else get_slices_fast<TILE, WPL, false>(c, draw0, S);
        } else {
          get_slices_dev(c, rng, S);
        }
      } else {
        get_slices_dev(c, rng, S);
      }
      rng.store(P, chain);
      store_slices(c, S);
    }
  }
  double seq;
  double maxw;
  cost_pass<TILE, WPL, true>(c, S, CacheSink<TILE, WPL>{c}, seq, maxw);
  const double rootpc = c.pcv[P.N - 1];
  P.total[chain] = rootpc;
  P.min_total[chain] = seq;  // get_cost(min_ctree) sums in traversal order (infinite_memory/utils.hpp:102-116)
  if (P.out_seq) P.out_seq[chain] = seq;
  if (P.out_maxw) P.out_maxw[chain] = maxw;
  if (FINITE && P.kwsz && (!Rng::kFast || P.slices_given)) build_kw_sz(c);
  if (P.bpar) snapshot_best(c, S, FINITE);
}

// ------------------------------------------------------------------------------------------ sweeps
TNB_D TNB_INLINE float exp2_fast(float x) { return exp2f(x); }  // ex2.approx under -ftz on the device
TNB_D TNB_INLINE int exp_of(double x) {  // biased binary exponent
#if defined(TNB_EMU)
  unsigned long long u;
  std::memcpy(&u, &x, 8);
  return int((u >> 52) & 0x7ffu);
#else
  return (__double2hiint(x) >> 20) & 0x7ff;
#endif
}

// Sum of all contraction costs of a chain (production mode keeps no partial-cost cache: the running total is
// re-based on this exact-as-possible sum every few sweeps, like the reference re-reads partial_cost.back()
// at the start of every update()).  Same association order on every lane -> tile-uniform result.
template <int TILE, int WPL>
TNB_D double sum_ccost(const ChainView<TILE, WPL>& c) {
  double acc = 0.0;
  for (int z = c.n + c.t.tl; z < c.P.N; z += TILE) acc += c.cc_get(z);
#if !defined(TNB_EMU)
#pragma unroll
  for (int d = TILE / 2; d > 0; d >>= 1)
    acc += TILE == 32 ? __shfl_xor_sync(0xffffffffu, acc, d, 32) : __shfl_xor_sync(c.t.mask, acc, d, TILE);
#endif
  return acc;
}

// Children words (child slot 0 in the low half, slot 1 in the high half): the other child, and one slot replaced.
TNB_D TNB_INLINE int other_child(uint32_t w, int x) { return int(w & 0xffffu) == x ? int(w >> 16) : int(w & 0xffffu); }
TNB_D TNB_INLINE uint32_t put_child(uint32_t w, int v, bool slot1) {
#if defined(TNB_EMU)
  return slot1 ? (w & 0xffffu) | (uint32_t(v) << 16) : (w & 0xffff0000u) | uint32_t(v);
#else
  return __byte_perm(w, uint32_t(v), slot1 ? 0x5410u : 0x3254u);  // one SEL + one PRMT
#endif
}

// One flat loop per tile: every iteration is either a sweep boundary (finish sweep s, start sweep s+1) or one
// level of the leaf->root walk, so the tiles sharing a warp stay converged instead of waiting for each other's
// walks.  The inputs of level k+1 (parent A', its children word and contraction cost, the sibling's index set
// and partial cost) do not depend on the move at level k, so they are loaded during level k in three stages and
// are in registers when level k+1 starts.
// Shared-memory-resident chain state (small networks, north_star item 4): bytes one chain occupies -- its parent
// array and its node records {header, index set} in the interleaved layout.
TNB_HD inline size_t smem_chain_bytes(int Npad, int n_int, int stride) {
  return (size_t(Npad) * 2 + 15) / 16 * 16 + size_t(n_int) * size_t(stride);
}

template <int TILE, int WPL, bool FINITE, class Rng, bool DIM2, bool HYPER = false, bool TRACE = false, bool SMEM = false>
TNB_D void chain_sweeps(const Params& P, int chain, char* smem_tile = nullptr) {
  ChainView<TILE, WPL> c(P, chain);
  const Tile<TILE>& t = c.t;
  const int n = P.n, root = P.N - 1;
  if (P.n_int == 0) {
    P.sweep_idx[chain] = P.until;
    return;
  }
  // SMEM: the chain's topology and node records live in shared memory for the whole launch (copied in here, copied
  // back at the end); the walk then never leaves the SM except for leaf index sets (read-only path) and snapshots.
  int16_t* g_par = c.par;
  char* g_rec = c.rec + w64(n) * c.hstride;  // record of internal node n, i.e. the chain's first
  const size_t par_bytes = (size_t(P.Npad) * 2 + 15) / 16 * 16;
  if constexpr (SMEM) {
#if !defined(TNB_EMU)
    uint32_t* sp = reinterpret_cast<uint32_t*>(smem_tile);
    const uint32_t* gp = reinterpret_cast<const uint32_t*>(g_par);
    for (int i = t.tl; i < P.Npad / 2; i += TILE) sp[i] = gp[i];
    uint32_t* sr = reinterpret_cast<uint32_t*>(smem_tile + par_bytes);
    const uint32_t* gr = reinterpret_cast<const uint32_t*>(g_rec);
    const int words = P.n_int * (P.hstride / 4);
    for (int i = t.tl; i < words; i += TILE) sr[i] = gr[i];
    t.sync();
    c.par = reinterpret_cast<int16_t*>(smem_tile);
    c.rec = smem_tile + par_bytes - w64(n) * c.hstride;
    c.rec_lane = c.rec + 16 + 4 * t.tl;
    c.smem = true;
#endif
  }
  // the per-chain bases stay in registers (otherwise they are re-derived from the kernel parameters at every use)
  keep_in_register<SMEM>(c.rec);
  keep_in_register<SMEM>(c.rec_lane);
  keep_in_register<SMEM>(c.par);
  if constexpr (WPL == 1) {  // "this lane owns a word": one live flag instead of S2R + AND + constant load + compare
    uint32_t okf = c.lane_ok[0] ? 1u : 0u;
    keep_in_register(okf);
    c.lane_ok[0] = okf != 0u;
  }
  // PC: keep the reference's partial-cost cache (parity modes).  The production kernel drops it: the walk then
  // needs no partial costs of D/E/C, no two 16-byte stores per level and half the fp64 adds; its total is a
  // running sum re-based on sum_ccost() every 64 sweeps.
  constexpr bool PC = !Rng::kFast;
  // INC: incremental best-tree snapshots (snapshot_walk) -- production generator, full-warp tiles
  constexpr bool INC = Rng::kFast && TILE == 32;
  // FS: production re-slicer (get_slices_fast); the walk then also maintains kw[] and sz[].  Only where a cost is
  // 2^popcount (DIM2): the table-cost kernels -- other dimensions, sparse indices -- re-slice with the reference's
  // slicer verbatim (get_slices_dev + a full cost pass), which knows every width model.
  // FSC: known at compile time (the 2^popcount kernels).  The table-cost production kernels take the same re-slicer
  // when widths are still plain popcounts -- a uniform dimension other than 2, no sparse indices, no general
  // dimensions (the host allocates kwsz[] exactly then) -- and re-cost with recost_all; otherwise the reference's
  // slicer verbatim.
  constexpr bool FSC = FINITE && Rng::kFast && DIM2;
  const bool FS = FSC || (FINITE && Rng::kFast && !DIM2 && P.kwsz != nullptr);
  // popcount | leaf count << 16 of every node; the per-chain base is pinned like the others (re-derived from the
  // kernel parameters it cost ten instructions per 2-byte access)
  uint32_t* kws = FS ? P.kwsz + size_t(chain) * P.Npad : nullptr;
  if constexpr (FSC) keep_in_register(kws);
  // FS: leaves below the children of B (slot order) and below C -- carried like the index sets, the sibling's count
  // loaded a level ahead (fetching them when a move is accepted stalled the warp on that load at every accept)
  int sz0 = 0, sz1 = 0, szC = 0;
  Rng rng;
  rng.load(P, chain);
  if constexpr (Rng::kFast) {
    rng.ool = FINITE || TILE < 32;  // (measured per kernel: C4 +15 %, C1 at TILE 4 +4 %, C2 -2 % when out of line)
    rng.begin_stream(t);
  }
  if constexpr (INC) rng.set_flag(0u);
  // Production kernels: every level consumes exactly one event of the generator and every sweep start one, so the
  // proposals of a launch are (events consumed) - (sweeps done) - (draws of the re-slicer): no counter in the loop.
  // P.n_prop[chain] is off by the starting values while the launch runs and is completed at its end.
  // Likewise width-gate rejections = proposals - proposals that passed the gate (counted where the gate is passed).
  if constexpr (Rng::kFast && FINITE) P.n_wrej[chain] -= P.n_prop[chain];
  if constexpr (Rng::kFast) P.n_prop[chain] += (unsigned long long)P.sweep_idx[chain] - rng.counter();
  // (sweep indices fit 32 bits: tnb_run refuses until_sweep >= 2^31; min_total stays in memory -- it is looked at
  //  once per sweep)
  int s = int(P.sweep_idx[chain]);
  const int until = int(P.until);
  double* const min_total_p = P.min_total + chain;
  bool rebase = true;
  // 16-bit counters in one register (proposals low, accepted moves high), folded into the 64-bit ones in memory
  // before either half can overflow (a sweep has fewer than 2^15 levels: node ids are int16)
  uint32_t q_pa = 0, q_wrej = 0;  // (production finite-width kernels: q_wrej counts the proposals that PASSED the width gate)
  uint32_t S[WPL];
#pragma unroll
  for (int k = 0; k < WPL; ++k) S[k] = 0u;
  if (FINITE) load_slices(c, S);
#if !defined(TNB_EMU) && !defined(TNB_NO_SFIX)
  // The slices are loaded here and first USED inside the loop.  ptxas then guards that first use -- the top of every
  // level, right behind the loads the level issues for the next one -- with a wait on the scoreboard of this load,
  // which is the scoreboard nearly all global loads of the kernel share: every level waited there for its own
  // prefetches to land (13.5 % of C4's stall samples on that one LOP3; `cuobjdump -sass` control words decoded by
  // scripts/sass_ctrl.py).  One ALU instruction on S before the loop consumes the load here, and the level's wait
  // moves to where the prefetched values are really taken over, at its end.
  if (FINITE) {
    const uint32_t z0 = uint32_t(P.n_chains) >> 31;  // zero, but not to the compiler
#pragma unroll
    for (int k = 0; k < WPL; ++k) S[k] ^= z0;
  }
#endif

  // The production (Philox) kernels run what the app runs -- Metropolis-Hastings with shared-index moves -- with the
  // two mode flags folded at compile time (per level: two constant loads, two compares and the generic acceptance
  // branch gone); other acceptance rules / disable_shared_inds belong to the core objects, i.e. the stream kernels
  // (the host refuses them in Philox mode).
  const int f_dsi = Rng::kFast ? 0 : P.dsi, f_prob = Rng::kFast ? kProbMH : P.prob_kind;
  bool in_sweep = false;
  int B = 0, A = -1, C = 0;
  // one level ahead (loop-carried): An = parent(A) and its header, loaded during the previous level
  int An = -1, Ann = -1;  // Ann = parent(An): two levels ahead, so that An's successor header can be requested early
  // contraction costs as the walk carries them: the high word of the double in the 2^popcount kernels (the low word
  // of a power of two is zero), the double itself otherwise
  using CR = typename std::conditional<DIM2, uint32_t, double>::type;
  auto f64 = [](CR v) -> double {
    if constexpr (DIM2) return ChainView<TILE, WPL>::hi_to_f64(v);
    else return v;
  };
  // B, A and An = parent(A) as the walk carries them (children word in slot order + contraction cost; a move replaces
  // one slot of B's and of A's word)
  using Node = typename std::conditional<DIM2, NodePacked, NodeWide>::type;
  Node nodeB, nodeA, nodeAn;
  uint32_t b0[WPL], b1[WPL], bC[WPL];
  // HYPER: inds[A], hyper[A] (loaded per level) and hyper[B] (carried: next level's B is this level's A)
  uint32_t bA[WPL], hA[WPL], hB[WPL];
#pragma unroll
  for (int k = 0; k < WPL; ++k) b0[k] = b1[k] = bC[k] = bA[k] = hA[k] = hB[k] = 0u;
  double pc0 = 0.0, pc1 = 0.0, pcC = 0.0, total = 0.0, root_pc = 0.0, beta = 0.0;
  float inv_beta_f = 0.f;
  // high word of the largest value the running total had since it was last re-summed (for positive doubles the
  // high words order like the values: one integer max per accepted move)
  int kmax_hi = 0;
  auto hi_of = [](double x) -> int { return int(ChainView<TILE, WPL>::f64_hi(x)); };

  // Two copies of the loop body for the unconstrained kernels: the rotation of the loop-carried registers (a fifth
  // of a level's instructions were MOVs) then happens by renaming.  (The finite-width kernel, with the re-slicer in
  // its boundary branch, loses a third of its speed to the doubled body.)
  constexpr int kUnroll = FINITE ? 1 : 2;
  if (!PC && s < until) {  // the running total starts from the exact sum of the contraction costs
    total = sum_ccost(c);
    kmax_hi = hi_of(total);
  }
  // decision trace (tests): lane 0 of a traced chain appends 32-byte records
  const bool tracing = TRACE && chain < P.trace_chains;
  unsigned long long tr_n = tracing ? P.trace_n[chain] : 0ull;
  uint32_t tr_sn = tracing ? P.trace_sn[chain] : 0u;
  auto trace_put = [&](uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, double d0, double d1) {
    if (t.tl == 0 && tr_n < P.trace_cap) P.trace[size_t(chain) * P.trace_cap + tr_n] = TraceRec{w0, w1, w2, w3, d0, d1};
    ++tr_n;
  };
  auto trace_slices = [&](const uint32_t (&S2)[WPL], bool kept, double r2, double r1) {
    if (tr_sn < P.trace_scap) {
      uint32_t* dst = P.trace_S + (size_t(chain) * P.trace_scap + tr_sn) * P.Ws;
#pragma unroll
      for (int k = 0; k < WPL; ++k) {
        const int w = t.tl + k * TILE;
        if (w < P.W) dst[w] = S2[k];
      }
    }
    trace_put(2u, tr_sn, 0u, kept ? 1u : 0u, r2, r1);
    ++tr_sn;
  };
  uint32_t iteration = 0;
#pragma unroll kUnroll
  while (true) {
    rng.tick(t, iteration++);
    if (A < 0) {
      // ------------------------------------------------------------------ sweep boundary
      if (in_sweep) {
        // INC: what did this sweep walk (taken before the re-slicer moves the generator's cursor)
        uint32_t n_ev = 0, slot_start = 0;
        if constexpr (INC) n_ev = rng.sweep_events(slot_start);
        if (FINITE && P.every > 0 && (s % P.every) == 0) {  // finite_width/greedy/optimizer.hpp:360-376
          bool anyS = false;
#pragma unroll
          for (int k = 0; k < WPL; ++k) anyS |= S[k] != 0u;
          if (t.any(anyS)) {
            uint32_t S2[WPL];
            dbl2* cp2 = P.cp2 + size_t(chain) * P.n_int;
            if constexpr (Rng::kFast) P.n_prop[chain] += rng.counter();  // (the re-slicer's draws are not proposals)
            bool fast_done = false;
            if constexpr (FINITE && Rng::kFast) if (FS) {
              fast_done = true;
              const uint32_t draw0 = rng.local_next();
              rng.sync_from0(t);
              if constexpr (DIM2) {
                get_slices_fast<TILE, WPL, false>(c, draw0, S2);
              } else {
                if (P.ksp) get_slices_fast<TILE, WPL, true>(c, draw0, S2);
                else get_slices_fast<TILE, WPL, false>(c, draw0, S2);
              }
              P.n_prop[chain] -= rng.counter();
              bool diff = false;
#pragma unroll
              for (int k = 0; k < WPL; ++k) diff |= S2[k] != S[k];
              if constexpr (TRACE) if (tracing && !t.any(diff)) trace_slices(S2, false, total, total);
              if (t.any(diff)) {  // same slices -> same costs: nothing to decide
                total = sum_ccost(c);  // the decision compares exact sums
                kmax_hi = hi_of(total);
                if (DIM2 && !HYPER && P.n_inds <= 1000) {
                  int* dz = reinterpret_cast<int*>(P.wkey + size_t(chain) * P.Npad);
                  const int shift0 = mark_slice_diff(c, S, S2, dz, P.word + size_t(chain) * P.Npad);
                  const double r2 = sum_shifted<TILE, WPL, false>(c, dz, shift0);
                  if constexpr (TRACE) if (tracing) trace_slices(S2, r2 < total, r2, total);
                  if (r2 < total) {
                    total = sum_shifted<TILE, WPL, true>(c, dz, shift0);
                    t.sync();
#pragma unroll
                    for (int k = 0; k < WPL; ++k) S[k] = S2[k];
                    store_slices(c, S);
                  }
                } else {
                  const double r2 = recost_all<TILE, WPL, DIM2>(c, S2, cp2);
                  if constexpr (TRACE) if (tracing) trace_slices(S2, r2 < total, r2, total);
                  if (r2 < total) {
                    t.sync();
                    for (int i = t.tl; i < P.n_int; i += TILE) c.cc_set(n + i, cp2[i].x);
                    t.sync();
#pragma unroll
                    for (int k = 0; k < WPL; ++k) S[k] = S2[k];
                    store_slices(c, S);
                    total = r2;
                  }
                }
              }
            }
            if (!fast_done) {
              get_slices_dev(c, rng, S2);
              if constexpr (Rng::kFast) P.n_prop[chain] -= rng.counter();
              double seq;
              double maxw;
              cost_pass<TILE, WPL, false>(c, S2, ScratchSink{cp2, n}, seq, maxw);
              const double r2 = cp2[P.n_int - 1].y;
              if (r2 < (PC ? root_pc : total)) {
                t.sync();
                for (int i = t.tl; i < P.n_int; i += TILE) {
                  c.cc_set(n + i, cp2[i].x);
                  c.pcv[n + i] = cp2[i].y;
                }
                t.sync();
#pragma unroll
                for (int k = 0; k < WPL; ++k) S[k] = S2[k];
                store_slices(c, S);
                root_pc = r2;
                total = r2;
              }
            }
          }
        }
        if (!PC) {
          // The running total is accurate only while nothing much larger than it has passed through it: a chain
          // that wandered up to 2^110 at small beta and came back carries the rounding of the large terms (absolute
          // error ~ 2^(kmax-52) per addition), up to a total of the wrong sign -- which would then be recorded as a
          // minimum that nothing can beat.  kmax is the largest exponent the total had since it was last re-summed;
          // re-sum whenever the total fell more than 2^10 below it, when it is not positive, and every 64 sweeps
          // regardless: the relative error of the total, and of min_total, stays below ~1e-10.  (The reference sums
          // positive partial costs and has no cancellation at all.  This is the only re-summation inside the loop:
          // more inlined copies of sum_ccost in the unrolled body cost more than they save.)
          const int et = exp_of(total);
          if (!(total > 0.0) || (kmax_hi >> 20) - et > 10 || ((s + 1) & 63) == 0) {
            total = sum_ccost(c);
            kmax_hi = hi_of(total);
          }
          root_pc = total;
        }
        if (root_pc < *min_total_p) {  // infinite_memory/optimizer.hpp:197-201
          *min_total_p = root_pc;
          bool done = false;
          if constexpr (INC) {
            if (rng.flag() != 0u && n_ev <= uint32_t(Rng::kEB)) {  // started from the best tree, and the ring still holds the walk
              snapshot_walk(c, rng, slot_start, int(n_ev) - 1, S, FINITE);
              done = true;
            }
            rng.set_flag(1u);
          }
          if (!done) snapshot_best(c, S, FINITE);
        } else {
          if constexpr (INC) rng.set_flag(0u);
        }
        ++s;
        in_sweep = false;
        if ((q_pa & 0x80008000u) != 0u || (FINITE && (s & 255) == 0)) {
          P.n_prop[chain] += q_pa & 0xffffu;
          P.n_acc[chain] += q_pa >> 16;
          if (FINITE) {
            if (Rng::kFast) P.n_wrej[chain] -= q_wrej;
            else P.n_wrej[chain] += q_wrej;
          }
          q_pa = q_wrej = 0;
        }
        if (rng.overrun()) break;
      }
      if (s >= until || !rng.can_start(P)) break;
      {
        const long long sb = (long long)s < P.n_betas ? (long long)s : P.n_betas - 1;
        if (Rng::kFast) inv_beta_f = P.inv_betas[sb];  // 1/beta (host-computed) for the threshold acceptance test
        else beta = P.betas[sb];
      }
      if constexpr (INC) rng.sweep_mark();
      // leaf = prng() % n_leaves (optimizer.hpp:103); the production RNG maps its word with a multiply-high instead
      const uint32_t lw = rng.leaf_word(t);
      const int leaf = Rng::kFast ? int(mulhi32(lw, uint32_t(n))) : int(lw % uint32_t(n));
      if constexpr (TRACE) if (tracing) trace_put(0u, lw, 0u, uint32_t(leaf), total, *min_total_p);
      B = c.par[leaf];
      if (PC) total = c.pcv[root];                       // :112
      rebase = false;
      root_pc = total;
      c.load_node(B, nodeB);
      // The sweep start is a chain of dependent loads -- leaf -> B -> A -> An, each with its header behind it -- and
      // the first level can only begin when its last link has landed.
      if constexpr (!FINITE) {
        // Chain first (C2 +2.5 %, C3 +1.8 %): every parent is asked for together with the header of its child, and
        // the next link is issued before the index sets that merely hang off the previous one.
        A = c.par[B];
        in_sweep = true;
        Ann = -1;
        if (A >= 0) {
          c.load_node(A, nodeA);
          An = c.par[A];
        }
        const int p0 = int(nodeB.w() & 0xffffu), p1 = int(nodeB.w() >> 16);
        c.load_bits(p0, b0);
        c.load_bits(p1, b1);
        if (PC) {
          pc0 = c.pc_of(p0);
          pc1 = c.pc_of(p1);
        }
        if (HYPER) c.load_hyp(B, hB);
        if (A >= 0) {
          if (An >= 0) {
            c.load_node(An, nodeAn);
            Ann = c.par[An];
          }
          C = other_child(nodeA.w(), B);
          c.load_bits(C, bC);
          if (HYPER) {
            c.load_bits(A, bA);
            c.load_hyp(A, hA);
          }
          if (PC) pcC = c.pc_of(C);
        }
      } else {
        // The finite-width kernel keeps the plain order: with header(A) issued early ptxas puts it on the scoreboard
        // of the index-set loads and the chain waits for them (C4 -0.6 %; with only the parents moved up -1.4 %).
        const int p0 = int(nodeB.w() & 0xffffu), p1 = int(nodeB.w() >> 16);
        c.load_bits(p0, b0);
        c.load_bits(p1, b1);
        if (PC) {
          pc0 = c.pc_of(p0);
          pc1 = c.pc_of(p1);
        }
        if (FS) {
          sz0 = int(kws[p0] >> 16);
          sz1 = int(kws[p1] >> 16);
        }
        if (HYPER) c.load_hyp(B, hB);
        A = c.par[B];
        in_sweep = true;
        if (A >= 0) {
          c.load_node(A, nodeA);
          C = other_child(nodeA.w(), B);
          c.load_bits(C, bC);
          if (HYPER) {
            c.load_bits(A, bA);
            c.load_hyp(A, hA);
          }
          if (PC) pcC = c.pc_of(C);
          if (FS) szC = int(kws[C] >> 16);
          An = c.par[A];
          Ann = -1;
          if (An >= 0) {
            c.load_node(An, nodeAn);
            Ann = c.par[An];
          }
        }
      }
    } else {
    // -------------------------------------------------------------------- one level (A >= 0)
    // (if/else, not `continue`: both kinds of iteration join again here so the tiles of a warp re-converge)
    // get_ctree_nn (optimize/optimizer.hpp:86-172): A = parent(B), C = sibling(B), D/E = children of B
    // Inputs of the NEXT level, issued before this level's own work so that every load has (at least) the whole
    // level to land.  They depend only on the chain of ancestors, which no move below them modifies:
    //   the sibling of A under An (from An's header, in registers) and its index set; the header of Ann = parent(An)
    //   (Ann itself was requested a level ago: a load whose ADDRESS is still in flight stalls the warp at issue);
    //   and Annn = parent(Ann) for the level after.
    rng.begin_level(t, B);  // (first: the event's word is needed right after the two votes below)
    const int Annn = Ann >= 0 ? int(c.par[Ann]) : -1;
    Node nodeAnn;
    if (Ann >= 0) c.load_node(Ann, nodeAnn);
    int Cn = 0;
    uint32_t bCn[WPL], bAn[WPL], hAn[WPL];
#pragma unroll
    for (int k = 0; k < WPL; ++k) bCn[k] = bAn[k] = hAn[k] = 0u;
    double pcCn = 0.0;
    int szCn = 0;
    if (An >= 0) {
      Cn = other_child(nodeAn.w(), A);
      c.load_bits(Cn, bCn);
      if (HYPER) {
        c.load_bits(An, bAn);
        c.load_hyp(An, hAn);
      }
      if (PC) pcCn = c.pc_of(Cn);
      if (FS) szCn = int(kws[Cn] >> 16);
    }
    const bool bslot0 = int(nodeA.w() & 0xffffu) == B;
    bool l0 = false, l1 = false;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      l0 |= (b0[k] & bC[k]) != 0u;
      l1 |= (b1[k] & bC[k]) != 0u;
    }
    const bool i0 = t.any_c(l0), i1 = t.any_c(l1);
    bool pick0;
    if constexpr (Rng::kFast) {  // (no disable_shared_inds here; plain predicate logic: coin when both intersect, else i0)
      const bool coin = (rng.coin_word(t) & 1u) != 0u;
      pick0 = i0 && (coin || !i1);
    } else {
      if (f_dsi || (i0 && i1)) pick0 = (rng.coin_word(t) & 1u) != 0u;
      else pick0 = i0;
    }
    int E = pick0 ? int(nodeB.w() >> 16) : int(nodeB.w() & 0xffffu);  // D = the other child
    uint32_t bD[WPL], bE[WPL], nb[WPL];
    // K10: no popcount of this kernel reaches 1024 (the host gives networks of 1024 and more indices two words per lane)
    constexpr bool K10 = DIM2 && TILE * WPL <= 32;
    uint32_t kpack = 0, ks = 0, ku = 0;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      bD[k] = pick0 ? b0[k] : b1[k];
      bE[k] = pick0 ? b1[k] : b0[k];
      nb[k] = bD[k] ^ bC[k];  // new inds of B (infinite_memory/optimizer.hpp:147)
      if (HYPER) nb[k] |= hA[k] | hB[k];
      const uint32_t ka_l = uint32_t(popc32(nb[k] | bE[k] | S[k])), kb_l = uint32_t(popc32(bD[k] | bC[k] | S[k]));
      if (FSC && K10) {
        // one reduction on the dependent path: sliced width of the new B and both cost exponents as 10-bit fields;
        // the unsliced popcount of the new B (only stored, on accept) goes through a second one that nothing waits for
        kpack += uint32_t(popc32(nb[k] & ~S[k])) | (ka_l << 10) | (kb_l << 20);
        ku += uint32_t(popc32(nb[k]));
      } else {
        kpack += ka_l | (kb_l << 16);
        if (FINITE) ks += uint32_t(popc32(nb[k] & ~S[k]));
        if (FS) ks += uint32_t(popc32(nb[k])) << 16;  // unsliced popcount of the new B rides in the same reduction
      }
    }
    // sparse-index cost model (table-cost kernels only): the sparse parts of the same three sets
    const bool sparse = !DIM2 && P.sparse != nullptr;
    const bool gen = !DIM2 && P.gdims != nullptr;  // general per-index dimensions: sequential products / sums
    uint32_t kspack = 0, kss = 0, kus = 0;
    uint32_t SP[WPL];
#pragma unroll
    for (int k = 0; k < WPL; ++k) SP[k] = 0u;
    if (sparse) {
      c.load_sparse(SP);
#pragma unroll
      for (int k = 0; k < WPL; ++k) {
        kspack += uint32_t(popc32((nb[k] | bE[k] | S[k]) & SP[k])) |
                  (uint32_t(popc32((bD[k] | bC[k] | S[k]) & SP[k])) << 16);
        if (FINITE) kss += uint32_t(popc32(nb[k] & ~S[k] & SP[k]));
        if (FS) kss += uint32_t(popc32(nb[k] & SP[k])) << 16;  // sparse popcount of the new B, for the re-slicer
      }
      kspack = t.sum_c(kspack);
      if (FINITE) kss = t.sum_c(kss);
      if (FS) {
        kus = kss >> 16;
        kss &= 0xffffu;
      }
    }
    const double pcD = pick0 ? pc0 : pc1;
    double pcE = pick0 ? pc1 : pc0;
    const int szD = pick0 ? sz0 : sz1;
    int szE = pick0 ? sz1 : sz0;
    if (!Rng::kFast) ++q_pa;
    bool gate = true;
    float swB = 0.f;  // new_sliced_width_B
    if (FINITE) {  // finite_width/greedy/optimizer.hpp:176-188
      if (FSC && K10) {
        kpack = t.sum_c(kpack);
        ku = t.sum_c(ku);
        ks = kpack & 0x3ffu;
      } else {
        ks = t.sum_c(ks);
        if (FS) {
          ku = ks >> 16;
          ks &= 0xffffu;
        }
      }
      if (gen) {
        uint32_t xs[WPL];
#pragma unroll
        for (int k = 0; k < WPL; ++k) xs[k] = nb[k] & ~S[k];
        swB = c.template gwidth_model<true>(xs, SP);
      } else if (!FS || sparse) {
        swB = sparse ? c.width_sp(int(ks), int(kss)) : c.width_of(int(ks));
      }
      // (2^popcount kernels: width = log2(d) * popcount is monotone in the popcount, and kthr is the largest popcount
      //  whose float32 width still fits -- the same decision without the fp64 product on the dependent path)
      gate = (FS && !sparse) ? int(ks) <= P.kthr : swB <= P.max_width;
      if (!Rng::kFast && !gate) ++q_wrej;
    }
    bool acc = false;
    CR nA{}, nB{};
    double delta = 0.0;
    if (gate) {
      if (Rng::kFast && FINITE) ++q_wrej;
      if (!(FSC && K10)) kpack = t.sum_c(kpack);
      if constexpr (DIM2) {  // dim == 2: the exact power of two built from its exponent (+inf past 2^1023, like pow)
        if constexpr (FSC && K10) {  // the high words straight from the fields
          nA = ((kpack << 10) & 0x3ff00000u) + 0x3ff00000u;
          nB = (kpack & 0xfff00000u) + 0x3ff00000u;
        } else if constexpr (K10) {
          nA = (kpack << 20) + 0x3ff00000u;
          nB = ((kpack >> 16) << 20) + 0x3ff00000u;
        } else {
          const uint32_t ka = kpack & 0xffffu, kb = kpack >> 16;
          nA = (1023u + (ka > 1024u ? 1024u : ka)) << 20;
          nB = (1023u + (kb > 1024u ? 1024u : kb)) << 20;
        }
      } else if (gen) {
        uint32_t uA[WPL], uB[WPL];
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
          uA[k] = nb[k] | bE[k] | S[k];
          uB[k] = bD[k] | bC[k] | S[k];
        }
        nA = c.template gcost_model<true>(uA, SP);
        nB = c.template gcost_model<true>(uB, SP);
      } else if (sparse) {
        nA = c.cost_sp(int(kpack & 0xffffu), int(kspack & 0xffffu));
        nB = c.cost_sp(int(kpack >> 16), int(kspack >> 16));
      } else {
        nA = c.cost_of(int(kpack & 0xffffu));  // cost(new_B | E [| slices])
        nB = c.cost_of(int(kpack >> 16));      // cost(D | C [| slices])
      }
      delta = (f64(nB) - nodeB.cost_f64()) + (f64(nA) - nodeA.cost_f64());  // :158, this association order

      if (Rng::kFast && f_prob == kProbMH) {
        // same rule as prob/mh.hpp:45-59, solved for the move: u <= (1 + delta/total)^-beta
        //   <=>  delta <= (2^(-log2(u)/beta) - 1) * total.
        // The threshold depends only on the event's uniform and beta, so it is ready before delta is; what is
        // left on the dependent path is one fp64 multiply and one compare (fp32 exp2: ~3e-6 relative error of
        // the threshold; an overflowing exponent gives +inf = accept, which is the limit of the rule).
        const float thr = exp2_fast(rng.neg_log2_u() * inv_beta_f) - 1.f;
        acc = delta <= double(thr) * total;
      } else {
        const double u = rng.uniform(t);  // always drawn (:162)
        double p;
        if (f_prob == kProbMH) {          // prob/mh.hpp:45-59
          if (delta <= 0.0) p = 1.0;
          else if (total == 0.0) p = 0.0;
          else p = pow(1.0 + delta / total, -beta);
        } else if (f_prob == kProbGreedy) {
          p = delta <= 0.0 ? 1.0 : 0.0;
        } else {
          p = 1.0;
        }
        acc = u <= p;
      }
    }
    if constexpr (TRACE) if (tracing) {
      uint32_t ib = 0u;
      if constexpr (Rng::kFast) ib = Rng::float_bits(inv_beta_f);
      trace_put(1u | (pick0 ? 4u : 0u) | (gate ? 8u : 0u) | (acc ? 16u : 0u) | ((i0 && i1) ? 32u : 0u) | (uint32_t(B) << 16),
                rng.coin_word(t), ib, uint32_t(A), delta, total);
    }
    // ---- random new slices (finite_width/greedy/optimizer.hpp:226-321; core-object option, stream kernels only):
    // the new tensor is too wide -> slice up to max_new random indices of it; if it fits then, swap, rebuild the whole
    // cost cache under the new slices and put the move to the acceptance rule with delta = new total - total.
    bool ns_taken = false;
    if constexpr (FINITE && !Rng::kFast) {
      if (!gate && P.max_new > 0) {
        int16_t* pos = P.posbuf + size_t(chain) * P.Ws * 32;
        uint32_t np = 0;
#pragma unroll
        for (int k = 0; k < WPL; ++k) {  // (new_inds_B - slices [- skip_slices]).positions(), ascending (:233-238)
          const int w = t.tl + k * TILE;
          uint32_t v = nb[k] & ~S[k];
          if (P.skip) v &= w < P.W ? ~P.skip[w] : 0u;
          if (P.grouped) v &= w < P.W ? P.leader[w] : 0u;
          uint32_t tot;
          uint32_t off = np + t.excl_scan_sum(uint32_t(popc32(v)), tot);
          while (v) {
            pos[off++] = int16_t(w * 32 + ctz32(v));
            v &= v - 1;
          }
          np += tot;
        }
        t.sync();
        uint32_t n_new = 0;
        float nsw = swB;
        if (t.tl == 0) {
          uint32_t n_pos = np;
          while (n_new < uint32_t(P.max_new) && nsw > P.max_width && n_pos > 0u) {
            const uint32_t r = rng.local_next() % n_pos;  // :246
            const int16_t tmp = pos[r];
            pos[r] = pos[n_pos - 1];
            pos[n_pos - 1] = tmp;
            const int idx = pos[n_pos - 1];
            // new_sliced_width_B -= log2_dims[...] in width_type (DimsCache<width_type>, :256-266)
            nsw -= gen ? float(P.glog2[idx]) : P.grouped ? float(int(P.gw[idx])) : float(P.log2d);
            --n_pos;
            ++n_new;
          }
        }
        rng.sync_from0(t);
        n_new = t.bcast(n_new, 0);
        const bool fits = t.bcast(nsw <= P.max_width ? 1u : 0u, 0) != 0u;
        t.sync();
        if (fits) {
          uint32_t S2[WPL];
#pragma unroll
          for (int k = 0; k < WPL; ++k) S2[k] = S[k];
          for (uint32_t j = np - n_new; j < np; ++j) {
            const int idx = pos[j];
            const int g = P.grouped ? int(P.gw[idx]) : 1;
#pragma unroll
            for (int k = 0; k < WPL; ++k) S2[k] |= span_mask(idx, g, t.tl + k * TILE);
          }
          // tentative swap_with_nn(E) in memory: the cost pass walks the tree as stored
          c.ch(A) = put_child(nodeA.w(), E, bslot0);
          c.ch(B) = put_child(nodeB.w(), C, pick0);
          c.par[C] = int16_t(B);
          c.par[E] = int16_t(A);
          c.store_bits(B, nb);
          t.sync();
          dbl2* cp2 = P.cp2 + size_t(chain) * P.n_int;
          double seq2, maxw2;
          cost_pass<TILE, WPL, false>(c, S2, ScratchSink{cp2, n}, seq2, maxw2);
          t.sync();
          const double r2 = cp2[P.n_int - 1].y;
          const double d2 = r2 - total;
          const double u2 = rng.uniform(t);
          double p2;
          if (f_prob == kProbMH) p2 = d2 <= 0.0 ? 1.0 : (total == 0.0 ? 0.0 : pow(1.0 + d2 / total, -beta));
          else if (f_prob == kProbGreedy) p2 = d2 <= 0.0 ? 1.0 : 0.0;
          else p2 = 1.0;
          if (u2 <= p2) {  // :290-312 (pos_C / pos_E keep their names in this branch)
            t.sync();
            for (int i = t.tl; i < P.n_int; i += TILE) {
              c.cc_set(n + i, cp2[i].x);
              c.pcv[n + i] = cp2[i].y;
            }
            t.sync();
            if (HYPER) {
#pragma unroll
              for (int k = 0; k < WPL; ++k) {
                hA[k] = bA[k] & nb[k] & bE[k];
                hB[k] = nb[k] & bD[k] & bC[k];
              }
              c.store_hyp(A, hA);
              c.store_hyp(B, hB);
            }
#pragma unroll
            for (int k = 0; k < WPL; ++k) S[k] = S2[k];
            store_slices(c, S);
            total = r2;
            root_pc = r2;
            q_pa += 0x10000u;
            ns_taken = true;
          } else {  // swap back (swap_with_nn(pos_C), :314-317)
            c.ch(A) = nodeA.w();
            c.ch(B) = nodeB.w();
            c.par[C] = int16_t(A);
            c.par[E] = int16_t(B);
            uint32_t ob[WPL];
#pragma unroll
            for (int k = 0; k < WPL; ++k) ob[k] = HYPER ? ((bD[k] ^ bE[k]) | hB[k]) : (bD[k] ^ bE[k]);
            c.store_bits(B, ob);
            t.sync();
          }
        }
      }
    }
    if (ns_taken) {
      // The whole cost cache was replaced (skip_cost_propagation, :311): nothing carried in registers for the levels
      // above is valid any more.  Re-enter the walk at B <- A from memory, like a sweep start does.
      if constexpr (FINITE && !Rng::kFast) {
        B = A;
        c.load_node(B, nodeB);
        const int p0 = int(nodeB.w() & 0xffffu), p1 = int(nodeB.w() >> 16);
        c.load_bits(p0, b0);
        c.load_bits(p1, b1);
        pc0 = c.pc_of(p0);
        pc1 = c.pc_of(p1);
        if (HYPER) c.load_hyp(B, hB);
        A = c.par[B];
        if (A >= 0) {
          c.load_node(A, nodeA);
          C = other_child(nodeA.w(), B);
          c.load_bits(C, bC);
          if (HYPER) {
            c.load_bits(A, bA);
            c.load_hyp(A, hA);
          }
          pcC = c.pc_of(C);
          An = c.par[A];
          Ann = -1;
          if (An >= 0) {
            c.load_node(An, nodeAn);
            Ann = c.par[An];
          }
        }
      }
    } else {
    uint32_t bB[WPL];
    if (acc) {
      // Tree::swap_with_nn(E): E <-> C, child slots preserved (tree.hpp:141-192)
      nodeA.set_w(put_child(nodeA.w(), E, bslot0));
      nodeB.set_w(put_child(nodeB.w(), C, pick0));
      nodeA.set_cost(nA);
      nodeB.set_cost(nB);
      if (PC) {
        c.ch(A) = nodeA.w();
        c.ch(B) = nodeB.w();
      } else {  // production: children word and new contraction cost of a node leave as ONE header store
        c.store_node(A, nodeA);
        c.store_node(B, nodeB);
      }
      c.par[C] = int16_t(B);
      c.par[E] = int16_t(A);
      c.store_bits(B, nb);
      if (HYPER) {  // :170-172 with E_old now under A and C_old now under B
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
          hA[k] = bA[k] & nb[k] & bE[k];
          hB[k] = nb[k] & bD[k] & bC[k];
        }
        c.store_hyp(A, hA);
        c.store_hyp(B, hB);
      }
      total += delta;
      if (!PC) {  // largest value of the running total since it was last re-summed (see the sweep boundary)
        const int h_now = hi_of(total);
        kmax_hi = h_now > kmax_hi ? h_now : kmax_hi;
      }
      q_pa += 0x10000u;
      if (FS) kws[B] = ku | (uint32_t(szD + szC) << 16);  // popcount and leaf count of the new B for the re-slicer
      if constexpr (!DIM2) {
        if (FS && sparse) P.ksp[size_t(chain) * P.Npad + B] = int16_t(kus);
      }
      {
        const int ti = C; C = E; E = ti;
        const double td = pcC; pcC = pcE; pcE = td;
        const int ts = szC; szC = szE; szE = ts;
      }
#pragma unroll
      for (int k = 0; k < WPL; ++k) {
        bB[k] = nb[k];
        bC[k] = bE[k];  // the node now called C is the old E
      }
    } else {
#pragma unroll
      for (int k = 0; k < WPL; ++k) bB[k] = HYPER ? ((bD[k] ^ bE[k]) | hB[k]) : (bD[k] ^ bE[k]);
    }
    // propagate partial costs (:185-188), post-swap names
    double pcB = 0.0;
    if (PC) {
      pcB = pcD + pcE + nodeB.cost_f64();
      const double pcA = pcB + pcC + nodeA.cost_f64();
      if constexpr (!DIM2) {  // (PC kernels are table-cost kernels)
        c.cc_f64(B) = nodeB.cost_f64();
        c.cc_f64(A) = nodeA.cost_f64();
      }
      c.pcv[B] = pcB;
      c.pcv[A] = pcA;
      root_pc = pcA;
    }
    // next level: B <- A, whose children are {B, C} in slot order
    nodeB = nodeA;
#pragma unroll
    for (int k = 0; k < WPL; ++k) {
      b0[k] = bslot0 ? bB[k] : bC[k];
      b1[k] = bslot0 ? bC[k] : bB[k];
    }
    if (PC) {
      pc0 = bslot0 ? pcB : pcC;
      pc1 = bslot0 ? pcC : pcB;
    }
    if (FS) {
      const int szB = szD + szE;  // post-swap names
      sz0 = bslot0 ? szB : szC;
      sz1 = bslot0 ? szC : szB;
    }
    B = A;
    A = An;
    {  // rotate the pipeline: what was loaded for the next level becomes current.  (Unconditional: at the root,
       // An < 0, the next iteration is a sweep boundary that reloads every one of these -- two branches less.)
      nodeA = nodeAn;
      C = Cn;
#pragma unroll
      for (int k = 0; k < WPL; ++k) bC[k] = bCn[k];
      if (HYPER) {
#pragma unroll
        for (int k = 0; k < WPL; ++k) {
          hB[k] = hA[k];
          bA[k] = bAn[k];
          hA[k] = hAn[k];
        }
      }
      if (PC) pcC = pcCn;
      if (FS) szC = szCn;
    }
    An = Ann;
    Ann = Annn;
    nodeAn = nodeAnn;
    }  // (no new-slice move taken)
    }  // level
  }
  if constexpr (TRACE) if (tracing) {
    P.trace_n[chain] = tr_n;
    P.trace_sn[chain] = tr_sn;
  }
  if constexpr (SMEM) {  // state back to its home in global memory
#if !defined(TNB_EMU)
    t.sync();
    const uint32_t* sp = reinterpret_cast<const uint32_t*>(smem_tile);
    uint32_t* gp = reinterpret_cast<uint32_t*>(g_par);
    for (int i = t.tl; i < P.Npad / 2; i += TILE) gp[i] = sp[i];
    const uint32_t* sr = reinterpret_cast<const uint32_t*>(smem_tile + par_bytes);
    uint32_t* gr = reinterpret_cast<uint32_t*>(g_rec);
    const int words = P.n_int * (P.hstride / 4);
    for (int i = t.tl; i < words; i += TILE) gr[i] = sr[i];
#endif
  }
  if constexpr (Rng::kFast) P.n_prop[chain] += rng.counter() - (unsigned long long)s;
  rng.store(P, chain);
  P.sweep_idx[chain] = s;
  P.total[chain] = PC ? c.pcv[root] : (rebase ? P.total[chain] : total);
  if constexpr (!Rng::kFast) P.n_prop[chain] += q_pa & 0xffffu;
  P.n_acc[chain] += q_pa >> 16;
  if constexpr (Rng::kFast && FINITE) P.n_wrej[chain] += P.n_prop[chain] - q_wrej;  // (n_prop is complete here)
  else P.n_wrej[chain] += q_wrej;
}

}  // namespace tnb
