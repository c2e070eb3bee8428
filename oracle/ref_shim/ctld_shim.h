// ctld_shim.h — what the sliced reference text expects to find around it:
// the typedefs of PublicHeader.h:35-47, logging/assert macros (Logger.h), the
// crane::grpc names it mentions, JobInCtld's accessors used by the
// *JobInScheduler constructors (JobScheduler.h:75-163), and the daemon
// singletons NodeSelect reads (JobScheduler.cpp:5566-5567, 5604-5664, 5770-5773,
// 5825; CtldPublicDefs.h:151-231; Node/NodeDefs.h:57-121) as plain structs that
// the harness fills from the C-ABI tables. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <expected>
#include <functional>
#include <limits>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <ranges>
#include <set>
#include <string>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <utility>
#include <variant>
#include <vector>

#include "absl_shim.h"
#include "fpm_shim.h"

// ---- crane::grpc / protobuf names that only appear in declarations -----------------
namespace crane::grpc {
class DeviceTypeSlotsMap;
class DedicatedResourceInNode;
class GresCount;
class GresMap;
class ResourceInNodeV3;
class ResourceV3;
class ResourceView;
enum PreemptType { PREEMPT_NONE = 0, PREEMPT_QOS = 1 };
enum PreemptMode { PREEMPT_MODE_OFF = 0, PREEMPT_MODE_CANCEL = 1 };
struct JobToCtld {
  struct License {};
  int licenses_count() const { return 0; }
  bool is_licenses_or() const { return false; }
};
}  // namespace crane::grpc
namespace google::protobuf {
template <class T>
struct RepeatedPtrField {
  RepeatedPtrField() = default;
  template <class U> RepeatedPtrField(U&&) {}
};
}  // namespace google::protobuf

// ---- PublicHeader.h:35-47 -----------------------------------------------------------------
using job_id_t = uint32_t;
using task_id_t = uint32_t;
using step_id_t = uint32_t;
using PartitionId = std::string;
using CranedId = std::string;
using ResvId = std::string;
using LicenseId = std::string;
using cpu_t = fpm::fixed<int64_t, __int128, 8>;

// ---- Logger.h: logging is dropped, assertions abort like the Debug build ------------------
#define CRANE_TRACE(...) ((void)0)
#define CRANE_DEBUG(...) ((void)0)
#define CRANE_INFO(...) ((void)0)
#define CRANE_WARN(...) ((void)0)
#define CRANE_ERROR(...) ((void)0)
#define CRANE_ASSERT(cond) do { if (!(cond)) { fprintf(stderr, "CRANE_ASSERT failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); abort(); } } while (0)
#define CRANE_ASSERT_MSG(cond, msg) CRANE_ASSERT(cond)
#define ABSL_ASSERT(cond) CRANE_ASSERT(cond)
template <class... A>
inline std::string fmt(A&&...) { return {}; }
inline const char* const kResourceTypeGpu = "gpu";  // PublicHeader.h:133
