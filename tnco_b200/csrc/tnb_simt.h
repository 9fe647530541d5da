// SIMT portability layer for the chain kernels.
//
// Product build (nvcc, sm_100a): a chain is owned by a TILE of 4/8/16/32 lanes of one warp; tile
// collectives map to VOTE / REDUX / SHFL with the tile's member mask.
// TNB_EMU build (plain g++, used ONLY by tests/emu to exercise the kernel logic on machines without a GPU;
// it is never loaded by the tnco_b200 package): TILE == 1, the collectives are identities.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(TNB_EMU)
#define TNB_HD
#define TNB_D
#define TNB_INLINE inline
struct tnb_double2 {
  double x, y;
};
typedef tnb_double2 dbl2;
static inline dbl2 make_dbl2(double x, double y) { return dbl2{x, y}; }
#else
#include <cuda_runtime.h>
#define TNB_HD __host__ __device__
#define TNB_D __device__
#define TNB_INLINE __forceinline__
typedef double2 dbl2;
static __device__ __forceinline__ dbl2 make_dbl2(double x, double y) { return make_double2(x, y); }
#endif

namespace tnb {

// Make a value opaque to the optimiser (it then stays in its register instead of being re-derived from kernel
// parameters at every use; the per-chain base pointers of the sweep loop are kept this way).
template <bool SHARED = false, class T>
TNB_D TNB_INLINE void keep_in_register(T*& p) {
#if !defined(TNB_EMU)
  asm volatile("" : "+l"(p));
  // the asm hides the address space: keep LDG/STG (or LDS/STS) instead of generic LD/ST
  if (SHARED) __builtin_assume(__isShared(p));
  else __builtin_assume(__isGlobal(p));
#else
  (void)p;
#endif
}
// Same for a 32-bit value: the compiler otherwise re-derives cheap-looking values (lane id -> "does this lane own a
// word of the bitset") from special registers and kernel parameters at every use -- four instructions and an S2R
// latency each time instead of one compare against a live register.
TNB_D TNB_INLINE void keep_in_register(uint32_t& v) {
#if !defined(TNB_EMU)
  asm volatile("" : "+r"(v));
#else
  (void)v;
#endif
}
#if defined(TNB_EMU)
#define TNB_NOINLINE
#else
#define TNB_NOINLINE __noinline__
#endif

template <int TILE>
struct Tile {
#if defined(TNB_EMU)
  static_assert(TILE == 1, "the emulation build runs one lane per chain");
  int tl = 0;
  TNB_D bool any(bool p) const { return p; }
  TNB_D bool any_c(bool p) const { return p; }
  TNB_D uint32_t sum_c(uint32_t v) const { return v; }
  TNB_D uint32_t bcast_c(uint32_t v, int) const { return v; }
  TNB_D uint32_t ballot(bool p) const { return p ? 1u : 0u; }
  TNB_D uint32_t max_u32(uint32_t v) const { return v; }
  TNB_D uint32_t sum(uint32_t v) const { return v; }
  TNB_D uint32_t bcast(uint32_t v, int) const { return v; }
  TNB_D void sync() const {}
#else
  unsigned mask;
  int tl;  // lane within the tile
  TNB_D Tile() {
    const int lane = threadIdx.x & 31;
    tl = lane & (TILE - 1);
    mask = TILE == 32 ? 0xffffffffu : (((1u << TILE) - 1u) << (lane & ~(TILE - 1)));
  }
  // (TILE == 32 uses the literal full mask: with a run-time mask the compiler brackets every vote / shuffle
  // with a divergence check and a WARPSYNC)
  TNB_D TNB_INLINE bool any(bool p) const {
    if (TILE == 32) return __any_sync(0xffffffffu, p) != 0;
    return (__ballot_sync(mask, p) & mask) != 0u;
  }
  // tile-relative lane mask of the lanes whose predicate holds (bit k = lane k of the tile)
  TNB_D TNB_INLINE uint32_t ballot(bool p) const {
    if (TILE == 32) return __ballot_sync(0xffffffffu, p);
    return (__ballot_sync(mask, p) & mask) >> ((threadIdx.x & 31) & ~(TILE - 1));
  }
  TNB_D TNB_INLINE uint32_t max_u32(uint32_t v) const {
    if (TILE == 32) return __reduce_max_sync(0xffffffffu, v);
#pragma unroll
    for (int d = TILE / 2; d > 0; d >>= 1) {
      const uint32_t o = __shfl_xor_sync(mask, v, d, TILE);
      v = o > v ? o : v;
    }
    return v;
  }
  // Full warp: one REDUX.  Sub-warp tiles: a shuffle butterfly -- REDUX with a partial member mask makes the
  // compiler run every tile of the warp exclusively (WARPSYNC.EXCLUSIVE), which serialises the tiles.
  TNB_D TNB_INLINE uint32_t sum(uint32_t v) const {
    if (TILE == 32) return __reduce_add_sync(0xffffffffu, v);
#pragma unroll
    for (int d = TILE / 2; d > 0; d >>= 1) v += __shfl_xor_sync(mask, v, d, TILE);
    return v;
  }
  TNB_D TNB_INLINE uint32_t bcast(uint32_t v, int src) const {
    if (TILE == 32) return __shfl_sync(0xffffffffu, v, src, 32);
    return __shfl_sync(mask, v, src, TILE);
  }
  TNB_D TNB_INLINE void sync() const {
    if (TILE == 32) __syncwarp(0xffffffffu);
    else __syncwarp(mask);
  }
  // "_c" variants for the sweep loop, where all lanes of a tile are converged by construction (control flow there
  // is tile-uniform): the member mask is the set of lanes converged right now instead of the tile's own mask.
  // With per-tile masks the hardware executes a vote / shuffle once per distinct mask, i.e. once per tile (ncu on
  // C1: the ballots ran with 7.6 of 32 lanes active and were 10 % of all instructions); with the common mask the
  // tiles that sit in the same branch share one instruction.  The results are still taken per tile.
  TNB_D TNB_INLINE bool any_c(bool p) const {
    if (TILE == 32) return __any_sync(0xffffffffu, p) != 0;
    return (__ballot_sync(__activemask(), p) & mask) != 0u;
  }
  TNB_D TNB_INLINE uint32_t sum_c(uint32_t v) const {
    if (TILE == 32) return __reduce_add_sync(0xffffffffu, v);
    const unsigned m = __activemask();
#pragma unroll
    for (int d = TILE / 2; d > 0; d >>= 1) v += __shfl_xor_sync(m, v, d, TILE);
    return v;
  }
  TNB_D TNB_INLINE uint32_t bcast_c(uint32_t v, int src) const {
    if (TILE == 32) return __shfl_sync(0xffffffffu, v, src, 32);
    return __shfl_sync(__activemask(), v, src, TILE);
  }
#endif
  // exclusive prefix sum over the tile's lanes (lane order); returns this lane's offset
  TNB_D TNB_INLINE uint32_t excl_scan_sum(uint32_t v, uint32_t& total) const {
#if defined(TNB_EMU)
    total = v;
    return 0;
#else
    uint32_t x = v;
#pragma unroll
    for (int d = 1; d < TILE; d <<= 1) {
      const uint32_t y = TILE == 32 ? __shfl_up_sync(0xffffffffu, x, d, 32) : __shfl_up_sync(mask, x, d, TILE);
      if (tl >= d) x += y;
    }
    total = bcast(x, TILE - 1);
    return x - v;
#endif
  }
  TNB_D TNB_INLINE double bcast_f64(double v, int src) const {
#if defined(TNB_EMU)
    (void)src;
    return v;
#else
    return TILE == 32 ? __shfl_sync(0xffffffffu, v, src, 32) : __shfl_sync(mask, v, src, TILE);
#endif
  }
  TNB_D TNB_INLINE unsigned long long bcast_u64(unsigned long long v, int src) const {
#if defined(TNB_EMU)
    (void)src;
    return v;
#else
    return __shfl_sync(mask, v, src, TILE);
#endif
  }
};

TNB_D TNB_INLINE int popc32(uint32_t x) {
#if defined(TNB_EMU)
  return __builtin_popcount(x);
#else
  return __popc(x);
#endif
}
TNB_D TNB_INLINE int ctz32(uint32_t x) {
#if defined(TNB_EMU)
  return __builtin_ctz(x);
#else
  return __ffs(int(x)) - 1;
#endif
}
// position of the r-th (0-based) set bit of x; x must have more than r bits set
TNB_D TNB_INLINE int nth_set_bit(uint32_t x, uint32_t r) {
#if defined(TNB_EMU)
  for (; r > 0; --r) x &= x - 1;
  return __builtin_ctz(x);
#else
  return int(__fns(x, 0u, int(r) + 1));
#endif
}
template <class T>
TNB_D TNB_INLINE T ldg(const T* p) {
#if defined(TNB_EMU)
  return *p;
#else
  return __ldg(p);
#endif
}
TNB_D TNB_INLINE double bits_to_f64(unsigned long long v) {
#if defined(TNB_EMU)
  double d;
  std::memcpy(&d, &v, 8);
  return d;
#else
  return __longlong_as_double((long long)v);
#endif
}
TNB_D TNB_INLINE uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(TNB_EMU)
  return uint32_t((uint64_t(a) * uint64_t(b)) >> 32);
#else
  return __umulhi(a, b);
#endif
}

}  // namespace tnb
