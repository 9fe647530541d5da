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
    if (!ok(cudaSetDevice(dev), "cudaSetDevice")) return false;
    if (!ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "cudaStreamCreate")) return false;
    if (!ok(cudaEventCreate(&ev0), "cudaEventCreate") || !ok(cudaEventCreate(&ev1), "cudaEventCreate")) return false;
    return true;
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
