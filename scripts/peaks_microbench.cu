// Measured per-SM integer / collective / on-chip-memory peaks of the B200 this repository runs on: the denominators
// of bench.py's roofline ("the per-SM POPC rate is to be measured by a microbenchmark on the
// box").  Stand-alone:
//     nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/peaks scripts/peaks_microbench.cu && /tmp/peaks
// prints one JSON object (committed as profiles/r02_int_peaks.json).
//
// Issue-rate kernels: every thread runs ILP independent dependency chains of one instruction (inline PTX, volatile,
// so nothing is folded), 1024 threads per block, 2 blocks per SM; rate = lane-operations / clock / SM from the
// kernel's own clock64() span and from CUDA-event time x the SM clock (both reported).
// Memory kernels: coalesced 128-byte warp loads over a working set sized for the level under test.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define CK(x)                                                                     \
  do {                                                                            \
    cudaError_t e_ = (x);                                                         \
    if (e_ != cudaSuccess) {                                                      \
      std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));               \
      std::exit(1);                                                               \
    }                                                                             \
  } while (0)

constexpr int ILP = 8;
constexpr int ITERS = 4096;

enum Op { POPC, LOP3, IADD, IMAD, REDUX, VOTE, SHFL, DADD, DMUL, EX2, FMUL };

template <int OP>
__global__ void __launch_bounds__(1024, 2) issue_kernel(unsigned* out, long long* span, unsigned seed) {
  unsigned v[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) v[k] = seed + threadIdx.x * 2654435761u + k * 40503u;
  double d[ILP];
  float f[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k) {
    d[k] = 1.0 + 1e-9 * double(v[k] & 1023u);
    f[k] = 0.5f + 1e-6f * float(v[k] & 1023u);
  }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      if (OP == POPC) asm volatile("popc.b32 %0, %0;" : "+r"(v[k]));
      if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[k]) : "r"(seed), "r"(it));
      if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(v[k]) : "r"(seed));
      if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[k]) : "r"(seed), "r"(it));
      if (OP == REDUX) asm volatile("redux.sync.add.u32 %0, %0, 0xffffffff;" : "+r"(v[k]));
      if (OP == VOTE)
        asm volatile("{ .reg .pred p; setp.ne.u32 p, %0, 0; vote.sync.ballot.b32 %0, p, 0xffffffff; }" : "+r"(v[k]));
      if (OP == SHFL) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(v[k]) : "r"(it & 31));
      if (OP == DADD) asm volatile("add.f64 %0, %0, %1;" : "+d"(d[k]) : "d"(1e-9));
      if (OP == DMUL) asm volatile("mul.f64 %0, %0, %1;" : "+d"(d[k]) : "d"(1.0000000001));
      if (OP == EX2) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[k]));
      if (OP == FMUL) asm volatile("mul.f32 %0, %0, %1;" : "+f"(f[k]) : "f"(1.0000001f));
    }
  }
  const long long t1 = clock64();
  unsigned acc = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) acc += v[k] + unsigned(d[k]) + unsigned(f[k]);
  if (acc == 0x12345u) out[0] = acc;
  if (threadIdx.x == 0) span[blockIdx.x] = t1 - t0;
}

// ---- memory: every warp reads `lines` consecutive 128-byte lines of its block's window, `reps` times
template <int SPACE>  // 0 = global through L1 (ld.global.ca), 1 = global bypassing L1 (ld.global.cg), 2 = shared
__global__ void __launch_bounds__(1024, 2) mem_kernel(const uint4* buf, size_t window_u4, int reps, unsigned* out,
                                                      long long* span) {
  extern __shared__ uint4 sm[];
  const uint4* base = buf + size_t(blockIdx.x) * window_u4;
  if (SPACE == 2) {
    for (size_t i = threadIdx.x; i < window_u4; i += blockDim.x) sm[i] = base[i];
    __syncthreads();
  }
  unsigned acc = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
#pragma unroll 4
    for (size_t i = threadIdx.x; i < window_u4; i += blockDim.x) {
      uint4 v;
      if (SPACE == 0) asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(base + i));
      if (SPACE == 1) asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(base + i));
      if (SPACE == 2) {
        const unsigned a = unsigned(__cvta_generic_to_shared(sm + i));
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
      }
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345u) out[0] = acc;
  if (threadIdx.x == 0) span[blockIdx.x] = t1 - t0;
}

// 4-byte-per-lane variant (what a lane-owns-a-word bitset layout issues): one 128-byte line per warp load
template <int SPACE>
__global__ void __launch_bounds__(1024, 2) mem4_kernel(const unsigned* buf, size_t window_u32, int reps, unsigned* out,
                                                       long long* span) {
  extern __shared__ uint4 sm[];
  unsigned* s32 = reinterpret_cast<unsigned*>(sm);
  const unsigned* base = buf + size_t(blockIdx.x) * window_u32;
  if (SPACE == 2) {
    for (size_t i = threadIdx.x; i < window_u32; i += blockDim.x) s32[i] = base[i];
    __syncthreads();
  }
  unsigned acc = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int r = 0; r < reps; ++r) {
#pragma unroll 8
    for (size_t i = threadIdx.x; i < window_u32; i += blockDim.x) {
      unsigned v;
      if (SPACE == 0) asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(base + i));
      if (SPACE == 1) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(base + i));
      if (SPACE == 2) {
        const unsigned a = unsigned(__cvta_generic_to_shared(s32 + i));
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
      }
      acc += v;
    }
  }
  const long long t1 = clock64();
  if (acc == 0x12345u) out[0] = acc;
  if (threadIdx.x == 0) span[blockIdx.x] = t1 - t0;
}

// dependent-load latency: one thread chases a random cycle through `n` 128-byte-strided slots
__global__ void chase_kernel(const unsigned* next, int steps, unsigned* out, long long* span, int cg) {
  unsigned p = 0;
  const long long t0 = clock64();
  for (int i = 0; i < steps; ++i) {
    if (cg) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(p) : "l"(next + size_t(p) * 32));
    else asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(p) : "l"(next + size_t(p) * 32));
  }
  const long long t1 = clock64();
  out[0] = p;
  span[0] = t1 - t0;
}

__global__ void stream_kernel(const uint4* buf, size_t n_u4, unsigned* out) {
  unsigned acc = 0;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_u4; i += size_t(gridDim.x) * blockDim.x) {
    const uint4 v = buf[i];
    acc += v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345u) out[0] = acc;
}

struct Timer {
  cudaEvent_t a, b;
  Timer() { CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); }
  void start() { CK(cudaEventRecord(a)); }
  float stop() {
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
  }
};

static long long max_span(long long* d_span, int n) {
  std::vector<long long> h(n);
  CK(cudaMemcpy(h.data(), d_span, n * sizeof(long long), cudaMemcpyDeviceToHost));
  long long m = 0;
  for (long long v : h) m = v > m ? v : m;
  return m;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  unsigned* d_out;
  long long* d_span;
  CK(cudaMalloc(&d_out, 64));
  CK(cudaMalloc(&d_span, sizeof(long long) * sms * 2));
  Timer tm;
  std::string js = "{";
  char line[512];
  std::snprintf(line, sizeof line, "\"gpu\": \"%s\", \"sms\": %d, \"sm_clock_khz_max\": %d, \"ilp\": %d, ", prop.name, sms,
                clock_khz, ILP);
  js += line;

  auto issue = [&](const char* name, auto kern) {
    const int blocks = sms * 2;
    kern<<<blocks, 1024>>>(d_out, d_span, 12345u);  // warm-up
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    long long best_span = 1ll << 62;
    for (int rep = 0; rep < 5; ++rep) {
      tm.start();
      kern<<<blocks, 1024>>>(d_out, d_span, 12345u + rep);
      const float ms = tm.stop();
      best_ms = ms < best_ms ? ms : best_ms;
      const long long sp = max_span(d_span, blocks);
      best_span = sp < best_span ? sp : best_span;
    }
    const double lane_ops_per_sm = 2.0 * 1024.0 * ILP * ITERS;
    const double per_clk = lane_ops_per_sm / double(best_span);
    const double g_per_s = lane_ops_per_sm * sms / (best_ms * 1e-3) / 1e9;
    std::snprintf(line, sizeof line,
                  "\"%s\": {\"lane_ops_per_clk_per_sm\": %.2f, \"warp_instr_per_clk_per_sm\": %.3f, \"glane_ops_per_s\": %.1f}, ",
                  name, per_clk, per_clk / 32.0, g_per_s);
    js += line;
  };
  issue("popc", issue_kernel<POPC>);
  issue("lop3", issue_kernel<LOP3>);
  issue("iadd", issue_kernel<IADD>);
  issue("imad", issue_kernel<IMAD>);
  issue("redux_add", issue_kernel<REDUX>);
  issue("vote_ballot", issue_kernel<VOTE>);
  issue("shfl_idx", issue_kernel<SHFL>);
  issue("dadd", issue_kernel<DADD>);
  issue("dmul", issue_kernel<DMUL>);
  issue("ex2_f32", issue_kernel<EX2>);
  issue("fmul", issue_kernel<FMUL>);

  // ---- on-chip and L2 bandwidth
  const size_t l2_total = size_t(48) << 20;  // 48 MiB: L2-resident on one B200 (126 MB L2)
  uint4* d_buf;
  const size_t hbm_bytes = size_t(4) << 30;
  CK(cudaMalloc(&d_buf, hbm_bytes));
  CK(cudaMemset(d_buf, 1, hbm_bytes));
  auto mem = [&](const char* name, auto kern, size_t window_bytes, int reps, size_t smem, size_t elt) {
    const int blocks = sms * 2;
    if (smem) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<blocks, 1024, smem>>>(reinterpret_cast<decltype(d_buf)>(d_buf), window_bytes / elt, 2, d_out, d_span);
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    long long best_span = 1ll << 62;
    for (int rep = 0; rep < 3; ++rep) {
      tm.start();
      kern<<<blocks, 1024, smem>>>(reinterpret_cast<decltype(d_buf)>(d_buf), window_bytes / elt, reps, d_out, d_span);
      const float ms = tm.stop();
      best_ms = ms < best_ms ? ms : best_ms;
      const long long sp = max_span(d_span, blocks);
      best_span = sp < best_span ? sp : best_span;
    }
    const double bytes_per_sm = 2.0 * double(window_bytes) * reps;
    std::snprintf(line, sizeof line, "\"%s\": {\"bytes_per_clk_per_sm\": %.1f, \"gb_per_s\": %.0f, \"window_bytes_per_block\": %zu}, ",
                  name, bytes_per_sm / double(best_span), bytes_per_sm * sms / (best_ms * 1e-3) / 1e9, window_bytes);
    js += line;
  };
  auto memv = [&](const char* name, auto kern, size_t window_bytes, int reps, size_t smem) {
    mem(name, kern, window_bytes, reps, smem, 16);
  };
  memv("l1_hit_ld128", mem_kernel<0>, 32 << 10, 400, 0);
  memv("smem_ld128", mem_kernel<2>, 32 << 10, 400, 32 << 10);
  memv("l2_hit_ld128", mem_kernel<1>, l2_total / (sms * 2) / 1024 * 1024, 40, 0);
  {
    auto k0 = mem4_kernel<0>;
    auto k1 = mem4_kernel<1>;
    auto k2 = mem4_kernel<2>;
    const int blocks = sms * 2;
    auto mem4 = [&](const char* name, void (*kern)(const unsigned*, size_t, int, unsigned*, long long*), size_t window_bytes,
                    int reps, size_t smem) {
      if (smem) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      kern<<<blocks, 1024, smem>>>(reinterpret_cast<const unsigned*>(d_buf), window_bytes / 4, 2, d_out, d_span);
      CK(cudaDeviceSynchronize());
      float best_ms = 1e30f;
      long long best_span = 1ll << 62;
      for (int rep = 0; rep < 3; ++rep) {
        tm.start();
        kern<<<blocks, 1024, smem>>>(reinterpret_cast<const unsigned*>(d_buf), window_bytes / 4, reps, d_out, d_span);
        const float ms = tm.stop();
        best_ms = ms < best_ms ? ms : best_ms;
        const long long sp = max_span(d_span, blocks);
        best_span = sp < best_span ? sp : best_span;
      }
      const double bytes_per_sm = 2.0 * double(window_bytes) * reps;
      std::snprintf(line, sizeof line, "\"%s\": {\"bytes_per_clk_per_sm\": %.1f, \"gb_per_s\": %.0f, \"window_bytes_per_block\": %zu}, ",
                    name, bytes_per_sm / double(best_span), bytes_per_sm * sms / (best_ms * 1e-3) / 1e9, window_bytes);
      js += line;
    };
    mem4("l1_hit_ld32", k0, 32 << 10, 400, 0);
    mem4("smem_ld32", k2, 32 << 10, 400, 32 << 10);
    mem4("l2_hit_ld32", k1, l2_total / (sms * 2) / 1024 * 1024, 40, 0);
  }
  {  // HBM stream
    stream_kernel<<<sms * 8, 1024>>>(d_buf, hbm_bytes / 16, d_out);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      tm.start();
      stream_kernel<<<sms * 8, 1024>>>(d_buf, hbm_bytes / 16, d_out);
      const float ms = tm.stop();
      best = ms < best ? ms : best;
    }
    std::snprintf(line, sizeof line, "\"hbm_read\": {\"gb_per_s\": %.0f, \"bytes\": %zu}, ", hbm_bytes / (best * 1e-3) / 1e9, hbm_bytes);
    js += line;
  }
  // ---- dependent-load latency
  auto chase = [&](const char* name, size_t slots, int cg) {
    std::vector<unsigned> h(slots * 32, 0u);
    std::vector<unsigned> perm(slots);
    for (size_t i = 0; i < slots; ++i) perm[i] = unsigned(i);
    unsigned s = 12345u;
    for (size_t i = slots - 1; i > 0; --i) {
      s = s * 1664525u + 1013904223u;
      const size_t j = s % (i + 1);
      std::swap(perm[i], perm[j]);
    }
    for (size_t i = 0; i < slots; ++i) h[size_t(perm[i]) * 32] = perm[(i + 1) % slots];
    CK(cudaMemcpy(d_buf, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    const int steps = 20000;
    chase_kernel<<<1, 1>>>(reinterpret_cast<const unsigned*>(d_buf), steps, d_out, d_span, cg);  // warm the level
    CK(cudaDeviceSynchronize());
    chase_kernel<<<1, 1>>>(reinterpret_cast<const unsigned*>(d_buf), steps, d_out, d_span, cg);
    CK(cudaDeviceSynchronize());
    const long long sp = max_span(d_span, 1);
    std::snprintf(line, sizeof line, "\"%s\": {\"cycles\": %.1f, \"working_set_bytes\": %zu}, ", name, double(sp) / steps, slots * 128);
    js += line;
  };
  chase("latency_l1_hit", 64, 0);                    // 8 KB
  chase("latency_l2_hit", (size_t(16) << 20) / 128, 1);  // 16 MB, L1 bypassed
  chase("latency_hbm", (size_t(2) << 30) / 128, 1);  // 2 GB > L2
  js += "\"note\": \"issue rates: lane-operations per clock per SM from the kernel's own clock64 span (2 blocks x 1024 threads per SM, 8 independent chains per thread); bandwidths: coalesced warp loads, 128-bit (ld128) and 32-bit per lane (ld32)\"}";
  std::printf("%s\n", js.c_str());
  return 0;
}
