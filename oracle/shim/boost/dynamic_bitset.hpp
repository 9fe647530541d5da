// TEST INFRASTRUCTURE ONLY. Minimal stand-in for boost::dynamic_bitset (absent from this image),
// providing exactly the API subset the reference headers use, so that the UNMODIFIED reference C++
// core can be compiled as the parity oracle (oracle/_ref). Pure bit container: cannot change results.
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
namespace boost {
template <typename Block = unsigned long, typename Alloc = std::allocator<Block>>
class dynamic_bitset {
 public:
  using size_type = std::size_t;
  using block_type = Block;
  static constexpr size_type bits_per_block = 8 * sizeof(Block);
  static constexpr size_type npos = static_cast<size_type>(-1);
  class reference {
    Block& b_; Block m_;
   public:
    reference(Block& b, size_type p) : b_{b}, m_{Block{1} << p} {}
    operator bool() const { return (b_ & m_) != 0; }
    reference& operator=(bool x) { if (x) b_ |= m_; else b_ &= ~m_; return *this; }
    reference& operator=(const reference& r) { return *this = bool(r); }
  };
  dynamic_bitset() = default;
  explicit dynamic_bitset(size_type n, unsigned long value = 0)
      : n_{n}, w_((n + bits_per_block - 1) / bits_per_block, 0) {
    if (!w_.empty()) w_[0] = static_cast<Block>(value);
    sanitize();
  }
  // From string: rightmost char is bit 0 (boost semantics)
  template <typename CharT, typename Traits, typename A>
  explicit dynamic_bitset(const std::basic_string<CharT, Traits, A>& s) : dynamic_bitset(s.size()) {
    for (size_type i = 0; i < s.size(); ++i) {
      const char c = s[s.size() - 1 - i];
      if (c == '1') set(i); else if (c != '0') throw std::invalid_argument("bad bit");
    }
  }
  size_type size() const { return n_; }
  size_type num_blocks() const { return w_.size(); }
  bool test(size_type p) const { return (w_[p / bits_per_block] >> (p % bits_per_block)) & 1; }
  bool operator[](size_type p) const { return test(p); }
  reference operator[](size_type p) { return reference(w_[p / bits_per_block], p % bits_per_block); }
  dynamic_bitset& set(size_type p, bool v = true) {
    if (v) w_[p / bits_per_block] |= Block{1} << (p % bits_per_block); else reset(p);
    return *this;
  }
  dynamic_bitset& reset(size_type p) { w_[p / bits_per_block] &= ~(Block{1} << (p % bits_per_block)); return *this; }
  size_type count() const { size_type c = 0; for (auto x : w_) c += __builtin_popcountl(x); return c; }
  bool any() const { for (auto x : w_) if (x) return true; return false; }
  bool none() const { return !any(); }
  size_type find_first() const { return find_from(0); }
  size_type find_next(size_type p) const { return p + 1 >= n_ ? npos : find_from(p + 1); }
  dynamic_bitset& operator&=(const dynamic_bitset& o) { for (size_type i = 0; i < w_.size(); ++i) w_[i] &= o.w_[i]; return *this; }
  dynamic_bitset& operator|=(const dynamic_bitset& o) { for (size_type i = 0; i < w_.size(); ++i) w_[i] |= o.w_[i]; return *this; }
  dynamic_bitset& operator^=(const dynamic_bitset& o) { for (size_type i = 0; i < w_.size(); ++i) w_[i] ^= o.w_[i]; return *this; }
  dynamic_bitset& operator-=(const dynamic_bitset& o) { for (size_type i = 0; i < w_.size(); ++i) w_[i] &= ~o.w_[i]; return *this; }
  dynamic_bitset operator~() const { dynamic_bitset r{*this}; for (auto& x : r.w_) x = ~x; r.sanitize(); return r; }
  bool intersects(const dynamic_bitset& o) const { for (size_type i = 0; i < std::min(w_.size(), o.w_.size()); ++i) if (w_[i] & o.w_[i]) return true; return false; }
  bool is_subset_of(const dynamic_bitset& o) const { for (size_type i = 0; i < w_.size(); ++i) if (w_[i] & ~o.w_[i]) return false; return true; }
  bool is_proper_subset_of(const dynamic_bitset& o) const { return is_subset_of(o) && w_ != o.w_; }
  friend bool operator==(const dynamic_bitset& a, const dynamic_bitset& b) { return a.n_ == b.n_ && a.w_ == b.w_; }
  friend bool operator!=(const dynamic_bitset& a, const dynamic_bitset& b) { return !(a == b); }
 private:
  size_type find_from(size_type p) const {
    for (size_type i = p / bits_per_block; i < w_.size(); ++i) {
      Block x = w_[i];
      if (i == p / bits_per_block) x &= ~Block{0} << (p % bits_per_block);
      if (x) return i * bits_per_block + __builtin_ctzl(x);
    }
    return npos;
  }
  void sanitize() { if (n_ % bits_per_block && !w_.empty()) w_.back() &= (Block{1} << (n_ % bits_per_block)) - 1; }
  size_type n_{0};
  std::vector<Block> w_;
};
}  // namespace boost
