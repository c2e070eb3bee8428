// absl_shim.h — the slice of Abseil the reference's scheduling hot path uses
// (abseil 20250814.1 is a fetched dependency, not under /root/reference;
// SURVEY.md §8c). TEST INFRASTRUCTURE ONLY: lets oracle/_ref compile the
// reference's own JobScheduler.{h,cpp} text with g++ 13.
//
//  * absl::Time / absl::Duration: int64 seconds with Abseil's saturation rules
//    for InfiniteFuture()/InfiniteDuration() (the path only ever holds whole
//    seconds: `now` is truncated at JobScheduler.cpp:1071, limits are seconds).
//  * absl::flat_hash_map / flat_hash_set: ORDERED std::map / std::set on a
//    monotonic bump arena. Two documented deviations hang on that
//    (SURVEY.md §8c D1/D3): iteration is key order instead of hash order, and a
//    NodeState inserted later has a higher address, so with nodes inserted in
//    index order the (cost, NodeState*) set of JobScheduler.h:588 breaks cost
//    ties by node index — the rule the oracle and the CUDA path implement.
#pragma once
#include <compare>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <limits>
#include <map>
#include <new>
#include <set>
#include <vector>

#include <sys/mman.h>

namespace absl {

class Duration {
 public:
  constexpr Duration() = default;
  static constexpr Duration FromSeconds(int64_t s) { Duration d; d.s_ = s; return d; }
  constexpr int64_t seconds() const { return s_; }
  constexpr bool is_inf() const { return s_ == std::numeric_limits<int64_t>::max(); }
  friend constexpr auto operator<=>(Duration a, Duration b) = default;
 private:
  int64_t s_{0};
};
constexpr Duration Seconds(int64_t n) { return Duration::FromSeconds(n); }
constexpr Duration Minutes(int64_t n) { return Duration::FromSeconds(n * 60); }
constexpr Duration Hours(int64_t n) { return Duration::FromSeconds(n * 3600); }
constexpr Duration ZeroDuration() { return Duration(); }
constexpr Duration InfiniteDuration() { return Duration::FromSeconds(std::numeric_limits<int64_t>::max()); }
constexpr int64_t ToInt64Seconds(Duration d) { return d.seconds(); }
constexpr Duration operator+(Duration a, Duration b) {
  return (a.is_inf() || b.is_inf()) ? InfiniteDuration() : Duration::FromSeconds(a.seconds() + b.seconds());
}
constexpr Duration operator-(Duration a, Duration b) {
  return a.is_inf() ? a : Duration::FromSeconds(a.seconds() - b.seconds());
}
constexpr Duration& operator+=(Duration& a, Duration b) { a = a + b; return a; }
constexpr Duration& operator-=(Duration& a, Duration b) { a = a - b; return a; }
constexpr Duration operator/(Duration a, int64_t k) { return a.is_inf() ? a : Duration::FromSeconds(a.seconds() / k); }
constexpr Duration operator*(Duration a, int64_t k) { return a.is_inf() ? a : Duration::FromSeconds(a.seconds() * k); }

class Time {
 public:
  constexpr Time() = default;  // UnixEpoch, like absl::Time{}
  static constexpr Time FromUnix(int64_t s) { Time t; t.s_ = s; return t; }
  constexpr int64_t unix_seconds() const { return s_; }
  constexpr bool is_inf() const { return s_ == std::numeric_limits<int64_t>::max(); }
  friend constexpr auto operator<=>(Time a, Time b) = default;
 private:
  int64_t s_{0};
};
constexpr Time InfiniteFuture() { return Time::FromUnix(std::numeric_limits<int64_t>::max()); }
constexpr Time FromUnixSeconds(int64_t s) { return Time::FromUnix(s); }
constexpr int64_t ToUnixSeconds(Time t) { return t.unix_seconds(); }
constexpr Time operator+(Time t, Duration d) {
  return (t.is_inf() || d.is_inf()) ? InfiniteFuture() : Time::FromUnix(t.unix_seconds() + d.seconds());
}
constexpr Time operator-(Time t, Duration d) { return t.is_inf() ? t : Time::FromUnix(t.unix_seconds() - d.seconds()); }
constexpr Duration operator-(Time a, Time b) {
  return a.is_inf() ? InfiniteDuration() : Duration::FromSeconds(a.unix_seconds() - b.unix_seconds());
}

// ---- monotonic arena -----------------------------------------------------------
namespace shim_internal {
// One lazily-touched virtual reservation per thread: addresses only grow, so
// "allocated later" == "higher address" holds for the whole call.
struct Arena {
  char* base = nullptr;
  size_t used = 0;
  static constexpr size_t kReserve = size_t(16) << 30;
  void* alloc(size_t bytes, size_t align) {
    if (!base) {
      void* p = mmap(nullptr, kReserve, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (p == MAP_FAILED) throw std::bad_alloc();
      base = static_cast<char*>(p);
    }
    size_t off = (used + align - 1) & ~(align - 1);
    if (off + bytes > kReserve) throw std::bad_alloc();
    used = off + bytes;
    return base + off;
  }
  void reset() {
    if (base && used) madvise(base, used, MADV_DONTNEED);
    used = 0;
  }
};
inline Arena& arena() { static thread_local Arena a; return a; }

template <class T>
struct ArenaAlloc {
  using value_type = T;
  ArenaAlloc() = default;
  template <class U> ArenaAlloc(const ArenaAlloc<U>&) {}
  T* allocate(size_t n) { return static_cast<T*>(arena().alloc(n * sizeof(T), alignof(T) < 16 ? 16 : alignof(T))); }
  void deallocate(T*, size_t) {}
  template <class U> bool operator==(const ArenaAlloc<U>&) const { return true; }
};
}  // namespace shim_internal

template <class K, class V, class Hash = void, class Eq = void>
class flat_hash_map : public std::map<K, V, std::less<K>, shim_internal::ArenaAlloc<std::pair<const K, V>>> {
 public:
  using Base = std::map<K, V, std::less<K>, shim_internal::ArenaAlloc<std::pair<const K, V>>>;
  using Base::Base;
  void reserve(size_t) {}
};
template <class K, class Hash = void, class Eq = void>
class flat_hash_set : public std::set<K, std::less<K>, shim_internal::ArenaAlloc<K>> {
 public:
  using Base = std::set<K, std::less<K>, shim_internal::ArenaAlloc<K>>;
  using Base::Base;
  void reserve(size_t) {}
};

class Mutex {};
class MutexLock { public: explicit MutexLock(Mutex*) {} };

}  // namespace absl
