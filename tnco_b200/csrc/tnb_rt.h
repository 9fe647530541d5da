// Runtime layer of the engine: device, stream, events, pooled device memory (or host memory in the TNB_EMU build).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>

#if !defined(TNB_EMU)
#include <cuda_runtime.h>
#endif

namespace tnb {

// ------------------------------------------------------------------------------------------ runtime layer
#if defined(TNB_EMU)
struct Rt {
  std::string err;
  int n_sms = 148;
  bool init(int) { return true; }
  void* alloc(size_t b) { return std::calloc(std::max<size_t>(b, 1), 1); }
  void free_(void* p) { std::free(p); }
  bool h2d(void* d, const void* h, size_t b) { std::memcpy(d, h, b); return true; }
  bool d2h(void* h, const void* d, size_t b) { std::memcpy(h, d, b); return true; }
  bool zero(void* d, size_t b) { std::memset(d, 0, b); return true; }
  bool fill_ff(void* d, size_t b) { std::memset(d, 0xff, b); return true; }
  bool sync() { return true; }
  void l2_window(void*, size_t, size_t) {}
  void l2_window_off() {}
  void destroy() {}
};
#else
struct Rt {
  std::string err;
  int device = 0, n_sms = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return true;
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
  }
  bool init(int dev) {
    int count = 0;
    if (!ok(cudaGetDeviceCount(&count), "cudaGetDeviceCount")) return false;
    if (dev < 0 || dev >= count) { err = "no such CUDA device"; return false; }
    cudaDeviceProp prop;
    if (!ok(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties")) return false;
    if (prop.major != 10) {
      err = "tnco_b200 needs an sm_100 (Blackwell B200) device; found sm_" + std::to_string(prop.major) +
            std::to_string(prop.minor) + " (there is no CPU or other-architecture fallback)";
      return false;
    }
    device = dev;
    n_sms = prop.multiProcessorCount;
    l2_persist_max = size_t(prop.persistingL2CacheMaxSize);
    l2_window_max = size_t(prop.accessPolicyMaxWindowSize);
    if (!ok(cudaSetDevice(dev), "cudaSetDevice")) return false;
    if (const char* f = std::getenv("TNB_L2_FETCH")) {  // measurement switch: DRAM -> L2 fetch granularity (32 / 64 / 128)
      if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, size_t(std::atoi(f))) != cudaSuccess) cudaGetLastError();
    }
    if (!ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
    if (!ok(cudaEventCreate(&ev0), "cudaEventCreate") || !ok(cudaEventCreate(&ev1), "cudaEventCreate")) return false;
    return true;
  }
  // L2 residency of the walk's hot block (ChainSet::hot: parents, node headers, kwsz -- the address-dependent loads
  // of the ancestor pipeline) while an HBM-resident batch streams its index sets through L2: the kernels launched
  // into the stream between l2_window() and l2_window_off() see [p, p + bytes) as persisting lines.  With a window
  // larger than the carve-out the hardware keeps a `hitRatio` share of its lines.
  size_t l2_persist_max = 0, l2_window_max = 0;
  bool l2_on = false;
  void l2_window(void* p, size_t bytes, size_t cap_bytes) {
    if (!p || !bytes || !l2_persist_max || !l2_window_max) return;
    const size_t carve = std::min(l2_persist_max, cap_bytes ? cap_bytes : l2_persist_max);
    const size_t win = std::min(bytes, l2_window_max);
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) != cudaSuccess) { cudaGetLastError(); return; }
    cudaStreamAttrValue a;
    std::memset(&a, 0, sizeof a);
    a.accessPolicyWindow.base_ptr = p;
    a.accessPolicyWindow.num_bytes = win;
    a.accessPolicyWindow.hitRatio = float(std::min(1.0, double(carve) / double(win)));
    a.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    a.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    if (cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &a) != cudaSuccess) { cudaGetLastError(); return; }
    l2_on = true;
  }
  void l2_window_off() {
    if (!l2_on) return;
    cudaStreamAttrValue a;
    std::memset(&a, 0, sizeof a);
    a.accessPolicyWindow.num_bytes = 0;
    cudaStreamSetAttribute(stream, cudaStreamAttributeAccessPolicyWindow, &a);
    cudaCtxResetPersistingL2Cache();
    l2_on = false;
  }
  // Device allocations are recycled through an exact-size free list: cudaMalloc / cudaFree synchronise the device
  // and were the most variable part (10-350 ms) of a back-to-back optimize() call that needs the same arrays again.
  std::multimap<size_t, void*> pool;
  std::unordered_map<void*, size_t> sizes;
  size_t pooled = 0;
  static constexpr size_t kPoolCap = size_t(16) << 30;
  void* alloc(size_t b) {
    b = std::max<size_t>(b, 16);
    void* p = nullptr;
    cudaSetDevice(device);
    auto it = pool.find(b);
    if (it != pool.end()) {
      p = it->second;
      pool.erase(it);
      pooled -= b;
    } else {
      if (cudaMalloc(&p, b) != cudaSuccess) {  // out of memory: give the cached blocks back and retry once
        cudaGetLastError();
        trim();
        if (!ok(cudaMalloc(&p, b), "cudaMalloc")) return nullptr;
      }
      sizes[p] = b;
    }
    cudaMemsetAsync(p, 0, b, stream);
    return p;
  }
  void free_(void* p) {
    if (!p) return;
    auto it = sizes.find(p);
    if (it == sizes.end()) { cudaFree(p); return; }
    if (pooled + it->second > kPoolCap) {
      cudaStreamSynchronize(stream);
      cudaFree(p);
      sizes.erase(it);
      return;
    }
    pool.emplace(it->second, p);  // later kernels of this stream are ordered after the ones still using it
    pooled += it->second;
  }
  void trim() {
    cudaStreamSynchronize(stream);
    for (auto& kv : pool) {
      cudaFree(kv.second);
      sizes.erase(kv.second);
    }
    pool.clear();
    pooled = 0;
  }
  bool h2d(void* d, const void* h, size_t b) {
    return b == 0 || ok(cudaMemcpyAsync(d, h, b, cudaMemcpyHostToDevice, stream), "cudaMemcpy H2D");
  }
  bool d2h(void* h, const void* d, size_t b) {
    if (b == 0) return true;
    return ok(cudaMemcpyAsync(h, d, b, cudaMemcpyDeviceToHost, stream), "cudaMemcpy D2H") && sync();
  }
  bool zero(void* d, size_t b) { return b == 0 || ok(cudaMemsetAsync(d, 0, b, stream), "cudaMemset"); }
  bool fill_ff(void* d, size_t b) { return b == 0 || ok(cudaMemsetAsync(d, 0xff, b, stream), "cudaMemset"); }
  bool sync() { return ok(cudaStreamSynchronize(stream), "cudaStreamSynchronize"); }
  void destroy() {
    trim();
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
  }
};
#endif

}  // namespace tnb
