/*
 * tnco_b200 -- C-ABI of the B200-native simulated-annealing engine for contraction trees.
 *
 * This is the drop-in boundary for ONE path of google-research/tnco (v0.4.0): what
 * `tnco.app.Optimizer(method='sa').optimize(tn, betas, n_steps, n_runs)` does below the Python driver,
 * i.e. the work the reference performs through its pybind module `tnco_core`
 *   tnco_core.optimize.infinite_memory.Optimizer_float64        (include/tnco/optimize/infinite_memory/main.hpp:33-51)
 *   tnco_core.optimize.finite_width.greedy.Optimizer_float64_float32 (…/finite_width/greedy/main.hpp:32-46)
 * one process and one object per run.  Here one engine owns one GPU and steps thousands of independent
 * chains (runs) per kernel launch.  All entry points are plain C: pointers + sizes, caller-owned host
 * buffers (C-contiguous), the engine owns all device memory.  Calls are blocking and not re-entrant per
 * engine.  Return value: 0 on success, negative on error with the text available from tnb_last_error().
 * There is NO CPU fallback: tnb_create() fails if no sm_100 device is usable.
 *
 * Node numbering is the reference's (include/tnco/tree.hpp:73-99): leaves are [0, n_leaves), internal
 * nodes follow, the root is node 2*n_leaves-2; -1 is "null" (include/tnco/node.hpp:33-43).
 * Index sets are bitsets of 32-bit words, bit i of the set = bit (i%32) of word (i/32)
 * (the reference's Bitset position i, include/tnco/bitset.hpp:54-80).  W32 = ceil(n_inds/32).
 */
#ifndef TNCO_B200_H
#define TNCO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tnb_engine tnb_engine;

/* acceptance rule (include/tnco/optimize/prob/{mh,greedy,base}.hpp) */
#define TNB_PROB_MH 0
#define TNB_PROB_GREEDY 1
#define TNB_PROB_ALWAYS 2

/* random-number source */
#define TNB_RNG_PHILOX 0  /* production: per-chain counter-based Philox4x32-10, generated in-kernel */
#define TNB_RNG_MT19937 1 /* parity: the reference's std::mt19937 stream per chain (seed = run seed), same
                             draw order as include/tnco/optimize/infinite_memory/optimizer.hpp:100-162 */
#define TNB_RNG_REPLAY 2  /* parity: caller-recorded raw 32-bit draw stream per chain (tnb_set_stream) */

/* chain-state layout in HBM (both are operated on in place) */
#define TNB_LAYOUT_AUTO 0        /* INTERLEAVED while the whole batch fits L2 (<= 96 MiB), SPLIT beyond */
#define TNB_LAYOUT_INTERLEAVED 1 /* one record {children, cost, index set} per tree node */
#define TNB_LAYOUT_SPLIT 2       /* node headers and index sets in separate arrays (headers stay L2-resident) */
#define TNB_LAYOUT_SMEM 3        /* chain state (parents + node records) resident in SHARED MEMORY for the whole launch:
                                    small networks, unconstrained production kernels; falls back to INTERLEAVED when
                                    a warp's chains do not fit 200 KB or the mode has no such kernel */

#define TNB_TREES_GREEDY 0 /* random tie-broken greedy merges (stand-in for opt_einsum 'greedy', tnco/utils/tn.py:225) */
#define TNB_TREES_RANDOM 1 /* uniformly random merges of index-sharing pairs */

int tnb_version(void);
/* last error of `e`, or of the failed tnb_create / host helper when e == NULL */
const char* tnb_last_error(const tnb_engine* e);

/* ---------------------------------------------------------------- host-side helpers (no GPU needed) */

/* Initial contraction trees, one per seed (replaces tnco/utils/tn.py:109-273 get_random_contraction_path
 * + tnco/ctree.py:108-226 for a CONNECTED network without hyper-indices).  Every contracted pair shares
 * an index (check_shared_inds).  Outputs are [n_trees][2*n_leaves-1] each. */
int tnb_random_trees(int n_leaves, int n_inds, const uint32_t* leaf_bits, int n_trees, const uint64_t* seeds,
                     int method, int n_threads, int32_t* parent, int32_t* child0, int32_t* child1);

/* Same for any network: index sets of intermediate tensors follow the hyper-count rule of
 * tnco/ctree.py:138-189 (an index shared by the two contracted tensors survives while other tensors, or the
 * output, still hold it).  output_bits [W32] marks the open (output) indices; NULL = none. */
int tnb_random_trees_out(int n_leaves, int n_inds, const uint32_t* leaf_bits, const uint32_t* output_bits,
                         int n_trees, const uint64_t* seeds, int method, int n_threads, int32_t* parent,
                         int32_t* child0, int32_t* child1);

/* Tree -> linear (einsum) contraction path; replaces include/tnco/utils.hpp:54-71 get_contraction +
 * tnco/ctree.py:350-388 ContractionTree.path().  The tree's leaf k is tensor tensors_pos[k] of a network of
 * n_tensors tensors (tensors_pos == NULL: identity, n_tensors = n_leaves); positions in the path count all
 * n_tensors tensors, as the reference's per-component paths do.  path is [n_trees][n_leaves-1][2]. */
int tnb_tree_to_path(int n_leaves, int n_trees, const int32_t* child0, const int32_t* child1, int n_tensors,
                     const int32_t* tensors_pos, int32_t* path);

/* Merge per-component linear paths (each expressed over all n_tensors tensors) into one path and connect the
 * components with trailing (0,1) steps; replaces tnco/utils/tn.py:334-401 merge_contraction_paths
 * (autocomplete=True), batched over runs.  lens[n_paths] = contractions per component path (same for every
 * run); paths is [n_runs][sum(lens)][2]; merged is [n_runs][n_tensors-1][2]. */
int tnb_merge_paths(int n_tensors, int n_runs, int n_paths, const int32_t* lens, const int32_t* paths,
                    int32_t* merged);

/* Linear path -> tree in reference numbering (tnco/ctree.py:108-131,208-218).  path is [n_leaves-1][2]. */
int tnb_path_to_tree(int n_leaves, const int32_t* path, int32_t* parent, int32_t* child0, int32_t* child1);

/* First n outputs of std::mt19937(seed) (include/tnco/optimize/optimizer.hpp:48,75). */
void tnb_mt19937_stream(uint32_t seed, uint64_t n, uint32_t* out);
/* Internal state of std::mt19937(seed) after n_draws outputs: the 624 words and the position that
 * libstdc++'s operator<< prints (reference Optimizer.prng_state, optimize/optimizer.hpp:191-195). */
void tnb_mt19937_state(uint32_t seed, uint64_t n_draws, uint32_t* state624, int32_t* pos);
/* advance a std::mt19937 state (624 words + position) by n_draws outputs, in place */
void tnb_mt19937_advance(uint32_t* state624, int32_t* pos, uint64_t n_draws);

/* ---------------------------------------------------------------- engine */

int tnb_create(tnb_engine** e, int device);
void tnb_destroy(tnb_engine* e);

/* Network = what ContractionTree carries besides the tree (include/tnco/ctree.hpp:32-40): per-leaf index
 * sets and dims.  leaf_bits [n_leaves][W32].  dims == NULL: every index has dimension `dim`
 * (ctree.hpp:80-89).  dims [n_inds] otherwise: all equal = the same as `dim`; all powers of two = carried as groups
 * of binary indices (costs stay 2^popcount); anything else = the reference's sequential product / width loops
 * (include/tnco/optimize/infinite_memory/cost_model/simple.hpp:46-54), bit-identical and much slower.
 * The network must be connected.  Output indices: tnb_set_output_inds. */
int tnb_set_network(tnb_engine* e, int n_leaves, int n_inds, const uint32_t* leaf_bits, uint64_t dim,
                    const uint64_t* dims);

/* Output (open) indices of the network, [W32] bitset or NULL for none (tnco/ctree.py output_inds).  Together with
 * the holders of every index this fixes the hyper counts; a network with hyper-indices (an index on 3+ tensors,
 * or on 2 tensors and open) runs the HYPER kernels, which keep the reference's HyperCache
 * (include/tnco/optimize/infinite_memory/utils.hpp:68-100).  Call after tnb_set_network; drops the chains. */
int tnb_set_output_inds(tnb_engine* e, const uint32_t* output_bits);
/* 1 if the current network + output indices have hyper-indices */
int tnb_is_hyper(tnb_engine* e);

/* Sparse-index cost model (SimpleCostModelSparseInds: include/tnco/optimize/infinite_memory/cost_model/
 * simple_sparse_inds.hpp:38-86 and finite_width/cost_model/simple_sparse_inds.hpp:38-157; what
 * tnco.app's optimize(..., n_projs=) selects when the network has sparse indices):
 *   cost  = get_cost(inds - sparse) * min(get_cost(inds & sparse), n_projs)
 *   width = get_width(inds - sparse) + min(get_width(inds & sparse), log2(n_projs))      (float32)
 * sparse_bits [W32] or NULL to return to the simple model; n_projs > 0.  Call after tnb_set_network; drops the
 * chains.  Served by the table-cost kernels; with max_width they re-slice with the reference's slicer verbatim. */
int tnb_set_sparse_inds(tnb_engine* e, const uint32_t* sparse_bits, uint64_t n_projs);

/* skip_slices of the finite-width core object (tnco/optimize/finite_width/optimizer.py:60,96-107;
 * include/tnco/optimize/finite_width/greedy/utils.hpp:76-79): indices the greedy slicer never takes.  [W32] or NULL
 * for none.  Honoured by the reference's slicer (stream kernels, table-cost kernels); under TNB_RNG_PHILOX a batch
 * with skip_slices is served by the table-cost kernels, which re-slice with that slicer (the production re-slicer
 * of the 2^popcount kernels does not know the option).  Call after tnb_set_network; drops the chains. */
int tnb_set_skip_slices(tnb_engine* e, const uint32_t* skip_bits);

/* max_width < 0 or +inf: unconstrained (infinite_memory optimizer).  Otherwise the memory-constrained
 * optimizer with float32 width arithmetic (tnco/app/app.py:757) and the greedy slicer, re-slicing on
 * sweeps s with s % update_slices_every == 0 (tnco/app/finite_width/sa.py:228). */
int tnb_set_mode(tnb_engine* e, double max_width, int update_slices_every, int disable_shared_inds,
                 int prob_kind, int rng_kind, int layout);
/* TNB_RNG_PHILOX (production) runs what tnco.app runs: shared-index moves, Metropolis-Hastings acceptance -- and
 * greedy / always acceptance as the two limits of the same threshold test (1/beta = 0 / +inf).  disable_shared_inds
 * -- a core-object option -- needs TNB_RNG_MT19937 or TNB_RNG_REPLAY; tnb_run and the chain constructors fail
 * otherwise. */

/* Change only the acceptance rule (the reference passes a prob object to every update()); chains are kept. */
int tnb_set_prob(tnb_engine* e, int prob_kind);

/* max_number_new_slices of the finite-width core object (tnco/optimize/finite_width/optimizer.py:59,
 * include/tnco/optimize/finite_width/greedy/optimizer.hpp:226-321): a move whose new tensor exceeds max_width may slice
 * up to that many random indices of it; if it fits then, the whole cost cache is rebuilt under the new slices and the
 * move is put to the acceptance rule.  0 (default -- what tnco.app uses) disables it.  Stream modes only
 * (TNB_RNG_MT19937 / TNB_RNG_REPLAY); chains are kept. */
int tnb_set_new_slices(tnb_engine* e, int max_number_new_slices);

/* Change only the re-slicing period (update(prob, update_slices) of the finite-width core object); chains kept. */
int tnb_set_update_slices(tnb_engine* e, int update_slices_every);

/* One chain per tree ([n_chains][2*n_leaves-1] each) and per seed.  Builds every chain's caches on the
 * device (index sets of internal nodes, contraction / partial costs, initial slices) exactly as the
 * reference constructors do (infinite_memory/optimizer.hpp:61-88, finite_width/greedy/optimizer.hpp:72-115).
 * chain_id0 = global id of chain 0 (multi-GPU sharding; enters the Philox counter). */
int tnb_set_chains(tnb_engine* e, int n_chains, const int32_t* parent, const int32_t* child0,
                   const int32_t* child1, const uint64_t* seeds, uint64_t chain_id0);

/* Same as tnb_set_chains, but the initial tree of every chain is built ON THE DEVICE by the chain's own lanes
 * (replaces tnco/utils/tn.py:109-273 get_random_contraction_path + tnco/ctree.py:108-226 for a connected,
 * hyper-index-free network; seed -> tree is deterministic).  method: TNB_TREES_GREEDY / TNB_TREES_RANDOM.
 * Fails with "not connected" if some chain runs out of index-sharing pairs. */
int tnb_generate_chains(tnb_engine* e, int n_chains, const uint64_t* seeds, uint64_t chain_id0, int method);

/* Resume chains from a saved state instead of constructing them afresh: what pickling a reference core object
 * restores through its constructor (tnco/optimize/infinite_memory/optimizer.py:234-247 and
 * finite_width/optimizer.py:330-346: seed = prng_state string, _min_ctree, _slices, _min_slices; C++ side
 * include/tnco/optimize/optimizer.hpp:57-72, finite_width/greedy/optimizer.hpp:72-101).  Call right after
 * tnb_set_chains (which gives the current trees); every argument may be NULL:
 *   mt_state     [n_chains][625]  std::mt19937 state, 624 words + position (TNB_RNG_MT19937 mode)
 *   slices       [n_chains][W32]  current slices: the constructor's slicer is not run
 *   best_*       [n_chains][2*n_leaves-1] saved min_ctree;  best_slices [n_chains][W32] saved min_slices
 * min_total_cost becomes get_cost(min_ctree[, min_slices]), as in the reference constructors. */
int tnb_set_resume(tnb_engine* e, const uint32_t* mt_state, const uint32_t* slices, const int32_t* best_parent,
                   const int32_t* best_child0, const int32_t* best_child1, const uint32_t* best_slices);

/* TNB_RNG_REPLAY: raw draw stream per chain, words [n_chains][len]; cursors reset to 0. */
int tnb_set_stream(tnb_engine* e, const uint32_t* words, uint64_t len);

/* Decision trace of the PRODUCTION kernels (TNB_RNG_PHILOX, 2^popcount costs, no hyper-indices), for parity tests:
 * chains [0, n_chains) append one 32-byte record per sweep start, per proposal and per re-slice
 *   struct { uint32_t w0, w1, w2, w3; double d0, d1; }
 *   w0: bits 0-1 kind (0 sweep start, 1 proposal, 2 re-slice); proposal flags: bit 2 D is child slot 0 of B,
 *       bit 3 width gate passed, bit 4 accepted, bit 5 a coin was drawn (both children of B intersect C);
 *       bits 16-31 node B
 *   w1: sweep start: the leaf's random word; proposal: float bits of -log2(u), coin in the last bit; re-slice: its number
 *   w2: proposal: float bits of 1/beta          w3: sweep start: leaf; proposal: node A; re-slice: 1 = new slices kept
 *   d0, d1: proposal: delta, running total before the move; sweep start: total, min_total; re-slice: cost under
 *       the candidate slices, cost under the current ones
 * so that a CPU restatement of the reference (include/tnco/optimize/infinite_memory/optimizer.hpp:90-201,
 * finite_width/greedy/optimizer.hpp:117-390) can be driven through the very same decisions and must arrive at the
 * same trees, index sets and contraction costs.  Call after tnb_set_chains / tnb_generate_chains; n_chains = 0
 * switches the trace off.  cap_records / cap_reslices: capacity per chain (further events are counted, not stored). */
int tnb_set_trace(tnb_engine* e, int n_chains, uint64_t cap_records, uint32_t cap_reslices);
/* records [min(*n_records, cap_records)] x 32 bytes and the candidate slices of every re-slice
 * [min(*n_reslices, cap_reslices)][W32] of one traced chain; any pointer may be NULL */
int tnb_get_trace(tnb_engine* e, int chain, uint64_t* n_records, void* records, uint32_t* n_reslices,
                  uint32_t* slices);
/* contraction cost of every node of the CURRENT tree of one chain, [2*n_leaves-1] (0 for leaves):
 * the reference's CostCache contraction_cost (include/tnco/optimize/infinite_memory/utils.hpp:32-56) */
int tnb_get_node_costs(tnb_engine* e, int chain, double* ccost);

/* Inverse temperatures, one per sweep (tnco/app/infinite_memory/sa.py:147-156,199-205). */
int tnb_set_betas(tnb_engine* e, const double* betas, int64_t n);

/* Advance every chain to sweep index `until_sweep` (one sweep == one reference Optimizer::update()).
 * In REPLAY mode a chain stops early when its stream cannot cover another sweep; see tnb_get_progress. */
int tnb_run(tnb_engine* e, int64_t until_sweep);

/* The same under a wall-clock budget: the reference's `timeout` (tnco/parallel.py:243-248 flips a stop flag that
 * every run polls once per sweep, tnco/app/infinite_memory/sa.py:201, and each run returns its best so far).  Here the
 * clock is checked between internal launches of K sweeps (K doubles from 16 while a launch takes under 50 ms), so the
 * call returns within one launch of the deadline.  timeout_s < 0 or +inf: no limit.  *reached = sweep index every
 * chain has reached (<= until_sweep); the chains stay valid and can be advanced further. */
int tnb_run_timed(tnb_engine* e, int64_t until_sweep, double timeout_s, int64_t* reached);

/* elapsed device time (ms, CUDA events on the engine's stream) and launch count of the sweep kernel
 * accumulated since the last call */
int tnb_get_timing(tnb_engine* e, double* kernel_ms, int64_t* launches);

/* per chain: partial_cost[root] of the current tree and min_total_cost (linear domain, fp64) */
int tnb_get_costs(tnb_engine* e, double* total, double* min_total);
/* per chain [n][2*n_leaves-1]; best != 0: the reference's min_ctree */
int tnb_get_trees(tnb_engine* e, int best, int chain0, int n, int32_t* parent, int32_t* child0, int32_t* child1);
/* the same trees in the engine's compact form: one word child0 | child1 << 16 per INTERNAL node,
 * [n][n_leaves-1] (node n_leaves + i at column i); a tenth of the bytes of tnb_get_trees, for bulk read-back */
int tnb_get_trees_packed(tnb_engine* e, int best, int chain0, int n, uint32_t* children);
/* index sets of every node of the CURRENT tree of one chain, [2*n_leaves-1][W32] */
int tnb_get_bits(tnb_engine* e, int chain, uint32_t* node_bits);
/* per chain [n][W32]; best != 0: min_slices */
int tnb_get_slices(tnb_engine* e, int best, int chain0, int n, uint32_t* slices);
/* per chain: sweeps done, proposals (level-loop iterations), accepted moves, width-gate rejections,
 * 32-bit draws consumed (stream modes).  Any pointer may be NULL. */
int tnb_get_progress(tnb_engine* e, int64_t* sweeps, uint64_t* proposals, uint64_t* accepts,
                     uint64_t* width_rejects, uint64_t* words);
/* sums over all chains */
int tnb_get_counters(tnb_engine* e, uint64_t* proposals, uint64_t* accepts, uint64_t* sweeps);

/* Full-tree evaluation of arbitrary trees (does not touch the chains): total cost summed in traversal
 * order (infinite_memory/utils.hpp:102-116 get_cost), partial_cost[root] (CostCache order, :32-56) and the
 * maximum log2 width over all nodes after removing `slices` ([n_trees][W32] or NULL). */
int tnb_eval_cost(tnb_engine* e, int n_trees, const int32_t* parent, const int32_t* child0,
                  const int32_t* child1, const uint32_t* slices, double* total_seq, double* total_pc,
                  double* max_width);

/* Evict the L2 cache (writes a buffer larger than L2); benchmarking aid. */
int tnb_flush_l2(tnb_engine* e);

/* the layout / tile shape the engine picked: lanes per chain, words per lane, TNB_LAYOUT_* , state bytes per chain */
int tnb_get_config(tnb_engine* e, int* tile, int* words_per_lane, int* layout, int* state_bytes_per_chain);

/* ---------------------------------------------------------------- several GPUs of one box behind one handle
 * For hosts that are not Python (the Python layer runs one process per GPU and exchanges over NCCL,
 * tnco_b200/dist.py).  What the reference does with one loky process per run (tnco/parallel.py:330-341) is done here
 * with one engine + one host thread per device: chains are sharded contiguously (chain i -> device floor(i*G/n)),
 * seeds and Philox counters use GLOBAL chain ids, so results do not depend on the number of devices.  The one
 * exchange step -- minimum over all chains + the winner's tree -- is tnb_group_get_best; inside one
 * process it is a host-side minimum over values tnb_get_costs has already brought back (nothing for NCCL to reduce).
 * `devices` may name the same device more than once (two engines on one GPU).  Per-device engines are reachable
 * through tnb_group_engine for everything not fanned out below (sparse indices, read-back of single chains, ...). */
typedef struct tnb_group tnb_group;
int tnb_group_create(tnb_group** g, const int* devices, int n_dev);
void tnb_group_destroy(tnb_group* g);
int tnb_group_size(const tnb_group* g);
tnb_engine* tnb_group_engine(tnb_group* g, int k);
const char* tnb_group_last_error(const tnb_group* g);
/* tnb_set_network (+ tnb_set_output_inds when output_bits != NULL), tnb_set_mode, tnb_set_betas on every device */
int tnb_group_set_network(tnb_group* g, int n_leaves, int n_inds, const uint32_t* leaf_bits, uint64_t dim,
                          const uint64_t* dims, const uint32_t* output_bits);
int tnb_group_set_mode(tnb_group* g, double max_width, int update_slices_every, int disable_shared_inds,
                       int prob_kind, int rng_kind, int layout);
int tnb_group_set_betas(tnb_group* g, const double* betas, int64_t n);
/* n_chains runs over all devices (>= one per device), seeds [n_chains]; trees are built on the devices */
int tnb_group_generate_chains(tnb_group* g, int n_chains, const uint64_t* seeds, int method);
/* all devices concurrently, each under the same wall-clock budget (tnb_run_timed); *reached = min over devices */
int tnb_group_run(tnb_group* g, int64_t until_sweep, double timeout_s, int64_t* reached);
/* [n_chains] in global chain order */
int tnb_group_get_costs(tnb_group* g, double* total, double* min_total);
int tnb_group_get_counters(tnb_group* g, uint64_t* proposals, uint64_t* accepts, uint64_t* sweeps);
/* the exchange step: best min_total_cost over all chains of all devices, its global chain id, its tree
 * ([2*n_leaves-1] each) and slices ([W32], finite width only); any pointer may be NULL */
int tnb_group_get_best(tnb_group* g, double* cost, int64_t* chain, int32_t* parent, int32_t* child0,
                       int32_t* child1, uint32_t* slices);

#ifdef __cplusplus
}
#endif
#endif
