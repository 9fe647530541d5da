// Internal declarations shared by the host helpers (tnb_host.cpp) and the CUDA engine (tnb_engine.cu).
#pragma once
#include <cstdint>
#include <string>

#include "../../include/tnco_b200.h"

namespace tnb {

void set_global_error(const std::string& s);
const char* global_error();

// std::mt19937-compatible generator (host side of TNB_RNG_MT19937).
struct Mt19937 {
  uint32_t x[624];
  int p = 624;
  void seed(uint32_t s);
  void refill();
  inline uint32_t next() {
    if (p >= 624) refill();
    uint32_t y = x[p++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  void fill(uint32_t* out, uint64_t n);
};

}  // namespace tnb
