// Kernels and their launchers, one set per tile shape (TILE lanes per chain, WPL bitset words per lane).
// The product build compiles every shape in its own translation unit (tnb_inst.cu, -DTNB_INST_TILE/-DTNB_INST_WPL)
// so that the shapes build in parallel; tnb_engine.cu only sees `extern template` declarations.  The TNB_EMU build
// (tests/emu) includes this header once and instantiates what it needs implicitly.
#pragma once
#include "tnb_kernels.h"
#include "tnb_rt.h"

namespace tnb {

// ------------------------------------------------------------------------------------------ kernels
constexpr int kBlock = 128;
constexpr size_t kTailWords = 256;  // padding words behind every array that load_bits reads rows from

#if !defined(TNB_EMU)
template <int TILE, int WPL, bool FINITE, class Rng>
__global__ void __launch_bounds__(kBlock) sa_init_kernel(const __grid_constant__ Params P) {
  const int chain = (blockIdx.x * kBlock + threadIdx.x) / TILE;
  if (chain >= P.n_chains) return;
  chain_init<TILE, WPL, FINITE, Rng>(P, chain);
}
// One warp per block: the hardware block scheduler then balances chains over the 148 SMs at warp granularity
// (4096 chains of 16 lanes = 2048 blocks = 13.8 per SM, all resident in a single wave at <= 128 registers).
constexpr int kSweepBlock = 32;
// Production (Philox) kernels are held to 72 registers (no spills) so that 28 single-warp blocks fit on an SM:
// 148 x 28 = 4144 resident chains at TILE = 32.  Parity kernels (fp64 pow, stream bookkeeping) keep 128.
// MINB = single-warp blocks resident per SM the register allocation must allow (28 -> 72 registers).
template <int TILE, int WPL, bool FINITE, class Rng, bool DIM2, int MINB, bool HYPER, bool TRACE>
__global__ void __launch_bounds__(kSweepBlock, MINB) sa_sweep_kernel(const __grid_constant__ Params P) {
  const int chain = (blockIdx.x * kSweepBlock + threadIdx.x) / TILE;
  if (chain >= P.n_chains) return;
  chain_sweeps<TILE, WPL, FINITE, Rng, DIM2, HYPER, TRACE>(P, chain);
}
// Shared-memory-resident variant: every tile of the block's warp owns P.smem_chain_bytes of dynamic shared memory.
template <int TILE, class Rng, int MINB>
__global__ void __launch_bounds__(kSweepBlock, MINB) sa_sweep_smem_kernel(const __grid_constant__ Params P) {
  extern __shared__ __align__(16) char smem_all[];
  const int chain = (blockIdx.x * kSweepBlock + threadIdx.x) / TILE;
  if (chain >= P.n_chains) return;
  chain_sweeps<TILE, 1, false, Rng, true, false, false, true>(P, chain, smem_all + size_t(threadIdx.x / TILE) * P.smem_chain_bytes);
}

template <int TILE, int WPL>
__global__ void __launch_bounds__(kBlock) sa_treegen_kernel(const __grid_constant__ Params P) {
  const int chain = (blockIdx.x * kBlock + threadIdx.x) / TILE;
  if (chain >= P.n_chains) return;
  chain_treegen<TILE, WPL>(P, chain);
}
#endif

template <int TILE, int WPL>
bool launch_treegen_t(Rt& rt, const Params& P) {
#if defined(TNB_EMU)
  (void)rt;
  for (int c = 0; c < P.n_chains; ++c) chain_treegen<TILE, WPL>(P, c);
  return true;
#else
  const long long threads = (long long)P.n_chains * TILE;
  const int grid = int((threads + kBlock - 1) / kBlock);
  if (grid == 0) return true;
  sa_treegen_kernel<TILE, WPL><<<grid, kBlock, 0, rt.stream>>>(P);
  return rt.ok(cudaGetLastError(), "sa_treegen_kernel launch");
#endif
}

template <int TILE, int WPL, bool FINITE, class Rng, bool DIM2, bool HYPER, bool TRACE = false>
static bool launch_h(Rt& rt, const Params& P, bool init) {
#if defined(TNB_EMU)
  (void)rt;
  for (int c = 0; c < P.n_chains; ++c) {
    if (init) chain_init<TILE, WPL, FINITE, Rng>(P, c);
    else chain_sweeps<TILE, WPL, FINITE, Rng, DIM2, HYPER, TRACE>(P, c);
  }
  return true;
#else
  const long long threads = (long long)P.n_chains * TILE;
  const int blk = init ? kBlock : kSweepBlock;
  const int grid = int((threads + blk - 1) / blk);
  if (grid == 0) return true;
  // occupancy class: production kernels 28 single-warp blocks per SM (<= 72 registers), parity kernels 16
  constexpr int MINB = Rng::kFast ? 28 : 16;
  if (init) sa_init_kernel<TILE, WPL, FINITE, Rng><<<grid, kBlock, 0, rt.stream>>>(P);
  else sa_sweep_kernel<TILE, WPL, FINITE, Rng, DIM2, MINB, HYPER, TRACE><<<grid, kSweepBlock, 0, rt.stream>>>(P);
  return rt.ok(cudaGetLastError(), init ? "sa_init_kernel launch" : "sa_sweep_kernel launch");
#endif
}

// HYPER kernels exist for full-warp tiles only (pick_tile gives hyper-index networks TILE = 32)
template <int TILE, int WPL, bool FINITE, class Rng, bool DIM2>
static bool launch_t(Rt& rt, const Params& P, bool init) {
  if constexpr (TILE == 32 || TILE == 1) {
    if (P.hyper && !init) return launch_h<TILE, WPL, FINITE, Rng, DIM2, true>(rt, P, init);
  } else {
    if (P.hyper) { rt.err = "hyper-index networks need TILE = 32"; return false; }
  }
#if !defined(TNB_EMU)
  // shared-memory-resident chains: unconstrained production kernel, one word per lane
  if constexpr (Rng::kFast && DIM2 && !FINITE && WPL == 1) {
    if (P.smem_chain_bytes > 0 && !init && !P.trace) {
      const long long threads = (long long)P.n_chains * TILE;
      const int grid = int((threads + kSweepBlock - 1) / kSweepBlock);
      if (grid == 0) return true;
      const size_t bytes = size_t(kSweepBlock / TILE) * size_t(P.smem_chain_bytes);
      auto kern = sa_sweep_smem_kernel<TILE, Rng, 28>;
      if (!rt.ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(bytes)), "smem size"))
        return false;
      kern<<<grid, kSweepBlock, bytes, rt.stream>>>(P);
      return rt.ok(cudaGetLastError(), "sa_sweep_smem_kernel launch");
    }
  }
#endif
  // decision trace (tests): the production kernel proper -- Philox, 2^popcount costs, no hyper-indices
  if constexpr (Rng::kFast && DIM2) {
    if (P.trace && !init) return launch_h<TILE, WPL, FINITE, Rng, DIM2, false, true>(rt, P, init);
  }
  return launch_h<TILE, WPL, FINITE, Rng, DIM2, false>(rt, P, init);
}

template <int TILE, int WPL>
bool launch_tw(Rt& rt, const Params& P, bool init, bool finite, bool stream_rng) {
  // parity (stream) modes always take costs from the std::pow table; the production kernel builds 2^k directly
  const bool d2 = P.dim2 != 0;
  if (finite) {
    if (stream_rng) return launch_t<TILE, WPL, true, RngStream<TILE>, false>(rt, P, init);
    return d2 ? launch_t<TILE, WPL, true, RngPhilox<TILE>, true>(rt, P, init)
              : launch_t<TILE, WPL, true, RngPhilox<TILE>, false>(rt, P, init);
  }
  if (stream_rng) return launch_t<TILE, WPL, false, RngStream<TILE>, false>(rt, P, init);
  return d2 ? launch_t<TILE, WPL, false, RngPhilox<TILE>, true>(rt, P, init)
            : launch_t<TILE, WPL, false, RngPhilox<TILE>, false>(rt, P, init);
}

}  // namespace tnb
