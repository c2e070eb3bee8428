// fpm_shim.h — the subset of fpm::fixed the scheduling path uses.
// fpm (MikeLankamp/fpm @ b46537fe, dependencies/cmake/fpm/CMakeLists.txt:6-8) is
// a fetched dependency that is not under /root/reference; its semantics are
// restated from its published header (fpm/fixed.hpp): value = raw / 2^F held in
// BaseType; from floating point rounds half away from zero; to integer divides
// the raw value by 2^F (C++ truncation); to floating point divides by 2^F;
// `*=`/`/=` with an integer act on the raw value; fixed*fixed rounds to nearest.
// Call sites on the path: SURVEY.md §8c. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cstdint>
#include <type_traits>

namespace fpm {
template <typename BaseType, typename IntermediateType, unsigned int FractionBits>
class fixed {
  static constexpr IntermediateType kMult = IntermediateType(1) << FractionBits;
  struct raw_tag {};
  constexpr fixed(BaseType v, raw_tag) : m_value(v) {}

 public:
  constexpr fixed() noexcept = default;
  template <typename T, std::enable_if_t<std::is_integral_v<T>>* = nullptr>
  constexpr explicit fixed(T val) noexcept : m_value(static_cast<BaseType>(val * kMult)) {}
  template <typename T, std::enable_if_t<std::is_floating_point_v<T>>* = nullptr>
  constexpr explicit fixed(T val) noexcept
      : m_value(static_cast<BaseType>((val >= 0.0) ? (val * static_cast<T>(kMult) + T{0.5})
                                                   : (val * static_cast<T>(kMult) - T{0.5}))) {}
  template <typename T, std::enable_if_t<std::is_floating_point_v<T>>* = nullptr>
  constexpr explicit operator T() const noexcept { return static_cast<T>(m_value) / static_cast<T>(kMult); }
  template <typename T, std::enable_if_t<std::is_integral_v<T>>* = nullptr>
  constexpr explicit operator T() const noexcept { return static_cast<T>(m_value / kMult); }
  constexpr BaseType raw_value() const noexcept { return m_value; }
  static constexpr fixed from_raw_value(BaseType v) noexcept { return fixed(v, raw_tag{}); }

  constexpr fixed operator-() const noexcept { return from_raw_value(-m_value); }
  constexpr fixed& operator+=(const fixed& y) noexcept { m_value += y.m_value; return *this; }
  constexpr fixed& operator-=(const fixed& y) noexcept { m_value -= y.m_value; return *this; }
  template <typename I, std::enable_if_t<std::is_integral_v<I>>* = nullptr>
  constexpr fixed& operator+=(I y) noexcept { m_value += y * kMult; return *this; }
  template <typename I, std::enable_if_t<std::is_integral_v<I>>* = nullptr>
  constexpr fixed& operator-=(I y) noexcept { m_value -= y * kMult; return *this; }
  constexpr fixed& operator*=(const fixed& y) noexcept {
    auto value = (static_cast<IntermediateType>(m_value) * y.m_value) / (kMult / 2);
    m_value = static_cast<BaseType>((value / 2) + (value % 2));
    return *this;
  }
  constexpr fixed& operator/=(const fixed& y) noexcept {
    auto value = (static_cast<IntermediateType>(m_value) * kMult * 2) / y.m_value;
    m_value = static_cast<BaseType>((value / 2) + (value % 2));
    return *this;
  }
  template <typename I, std::enable_if_t<std::is_integral_v<I>>* = nullptr>
  constexpr fixed& operator*=(I y) noexcept { m_value *= y; return *this; }
  template <typename I, std::enable_if_t<std::is_integral_v<I>>* = nullptr>
  constexpr fixed& operator/=(I y) noexcept { m_value /= y; return *this; }

  friend constexpr fixed operator+(fixed a, const fixed& b) noexcept { return a += b; }
  friend constexpr fixed operator-(fixed a, const fixed& b) noexcept { return a -= b; }
  friend constexpr fixed operator*(fixed a, const fixed& b) noexcept { return a *= b; }
  friend constexpr fixed operator/(fixed a, const fixed& b) noexcept { return a /= b; }
  template <typename I, std::enable_if_t<std::is_integral_v<I>>* = nullptr>
  friend constexpr fixed operator*(fixed a, I b) noexcept { return a *= b; }
  friend constexpr bool operator==(const fixed& a, const fixed& b) noexcept { return a.m_value == b.m_value; }
  friend constexpr bool operator!=(const fixed& a, const fixed& b) noexcept { return a.m_value != b.m_value; }
  friend constexpr bool operator<(const fixed& a, const fixed& b) noexcept { return a.m_value < b.m_value; }
  friend constexpr bool operator>(const fixed& a, const fixed& b) noexcept { return a.m_value > b.m_value; }
  friend constexpr bool operator<=(const fixed& a, const fixed& b) noexcept { return a.m_value <= b.m_value; }
  friend constexpr bool operator>=(const fixed& a, const fixed& b) noexcept { return a.m_value >= b.m_value; }

 private:
  BaseType m_value{0};
};
}  // namespace fpm
