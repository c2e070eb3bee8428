// algebra.cuh — the reference's per-node resource algebra on bit-mask rows.
//
// Device-side restatement of src/Utilities/PublicHeader/PublicHeader.cpp (R8 in
// SURVEY.md §8a): std::set<uint32_t> core ids and std::set<SlotId> slot sets
// become bit masks, so set inclusion / intersection / difference are
// AND/ANDN and "take the n lowest ids" is a select-lowest-n-bits.
// Functions are __host__ __device__ so the C-ABI layer can validate inputs
// with the same code; the scheduling itself only ever runs them on the GPU.
#pragma once

#include <stdint.h>

#include "../../include/crane_sched.h"

#ifndef CRANE_HD
#define CRANE_HD __host__ __device__ __forceinline__
#endif

namespace crane {

// Device-side views of the ABI rows with the SAME memory layout; the eight
// uint16 gres slot masks / counts are handled as two 64-bit words so set
// algebra on all gres entries is two AND/ANDN instructions.
struct __align__(8) Row {   // crane_res_in_node_t / ResourceInNodeV3 (PublicHeader.h:562)
  int64_t cpu_raw;
  uint64_t mem;
  uint64_t mem_sw;
  uint64_t core[CRANE_CORE_WORDS];
  uint64_t g[2];            // gres[0..3], gres[4..7]
};
struct __align__(8) View {  // crane_res_view_t / ResourceView (PublicHeader.h:671)
  int64_t cpu_raw;
  uint64_t mem;
  uint64_t mem_sw;
  uint64_t gtot[2];         // gres_total[0..7] as 16-bit fields
  uint64_t gspec[2];        // gres_spec[0..7]
};
static_assert(sizeof(Row) == sizeof(crane_res_in_node_t), "Row layout");
static_assert(sizeof(View) == sizeof(crane_res_view_t), "View layout");

struct GresDict {
  uint32_t n_entries;
  uint32_t pad;
  uint8_t entry_name[CRANE_GRES_ENTRIES];
  uint8_t name_first[CRANE_GRES_NAMES];   // first entry of each name (entries of a name are contiguous)
  uint8_t name_count[CRANE_GRES_NAMES];
  uint64_t name_mask8[CRANE_GRES_NAMES];  // 0xFF in byte e for every entry e of the name
};

CRANE_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
CRANE_HD int popc64(uint64_t v) {
#if defined(__CUDA_ARCH__)
  return __popcll(v);
#else
  return __builtin_popcountll(v);
#endif
}
// 16-bit field e (0..7) of a two-word array; selects instead of indexing so
// the words stay in registers
CRANE_HD uint32_t field16(const uint64_t* w, uint32_t e) {
  const uint64_t x = (e & 4u) ? w[1] : w[0];
  return (uint32_t)(x >> ((e & 3u) * 16u)) & 0xffffu;
}
CRANE_HD void set_field16(uint64_t* w, uint32_t e, uint32_t v) {
  const uint32_t sh = (e & 3u) * 16u;
  const uint64_t clr = ~(0xffffull << sh), val = (uint64_t)(v & 0xffffu) << sh;
  if (e & 4u) w[1] = (w[1] & clr) | val; else w[0] = (w[0] & clr) | val;
}

// the n lowest set bits of m: "it = begin(); n times ++it" over an ordered
// std::set (PublicHeader.cpp:535-537, 572-574).
CRANE_HD uint64_t lowest_bits64(uint64_t m, int n) {
  if (n >= popc64(m)) return m;
  if (n <= 0) return 0;
  // branch-free select of the n-th set bit: halve the search window six times
  uint32_t pos = 0;
  int rem = n;
#pragma unroll
  for (uint32_t w = 32; w >= 1; w >>= 1) {
    const int c = popc64((m >> pos) & ((1ull << w) - 1ull));
    if (c < rem) { rem -= c; pos += w; }
  }
  return m & ((2ull << pos) - 1ull);  // pos <= 62 here because n < popcount(m)
}

CRANE_HD void row_zero(Row& r) {
  r.cpu_raw = 0; r.mem = 0; r.mem_sw = 0;
  for (int w = 0; w < CRANE_CORE_WORDS; ++w) r.core[w] = 0;
  r.g[0] = r.g[1] = 0;
}
CRANE_HD bool core_empty(const Row& r) { return (r.core[0] | r.core[1] | r.core[2] | r.core[3]) == 0; }
CRANE_HD int core_count(const Row& r) {
  return popc64(r.core[0]) + popc64(r.core[1]) + popc64(r.core[2]) + popc64(r.core[3]);
}

// ResourceInNodeV3::operator-= (PublicHeader.cpp:789-796; CpuSet 758-766;
// DedicatedResourceInNode 204-217; TypeSlotsMap 316-328): tolerant core erase,
// counts subtract, slot sets difference (empty entries vanish = mask 0).
CRANE_HD void row_sub(Row& a, const Row& b) {
  a.cpu_raw -= b.cpu_raw;
  a.mem -= b.mem;
  a.mem_sw -= b.mem_sw;
  for (int w = 0; w < CRANE_CORE_WORDS; ++w) a.core[w] &= ~b.core[w];
  a.g[0] &= ~b.g[0];
  a.g[1] &= ~b.g[1];
}
// ResourceInNodeV3::operator+= (PublicHeader.cpp:781-787)
CRANE_HD void row_add(Row& a, const Row& b) {
  a.cpu_raw += b.cpu_raw;
  a.mem += b.mem;
  a.mem_sw += b.mem_sw;
  for (int w = 0; w < CRANE_CORE_WORDS; ++w) a.core[w] |= b.core[w];
  a.g[0] |= b.g[0];
  a.g[1] |= b.g[1];
}
// operator<=(ResourceInNodeV3, ResourceInNodeV3) (PublicHeader.cpp:886-890,
// 159-169, 334-343): cpu, mem, slot-set inclusion. core ids and mem_sw are NOT
// compared.
CRANE_HD bool row_le(const Row& a, const Row& b) {
  return a.cpu_raw <= b.cpu_raw && a.mem <= b.mem && ((a.g[0] & ~b.g[0]) | (a.g[1] & ~b.g[1])) == 0;
}

// ---- the window minimum ------------------------------------------------------
// JobScheduler.cpp:5314-5319 folds  row = res_avail; for seg in window:
// row.Ckmin(seg)  (PublicHeader.cpp:815-827). Its core rule ("intersect only
// when both sides are non-empty") is not associative as written, but the fold
// equals  core = res_avail.core & AND{ seg.core : seg.core != {} }  (once the
// running set is empty it stays empty either way); cpu/mem are minima and the
// gres sets plain intersections. The kernels therefore AND-reduce the masks
// with all-ones standing in for an empty core set, and test cpu/mem per
// segment.

// ---- requests ---------------------------------------------------------------
// req_node_res_view + req_task_res_view * t  (JobScheduler.cpp:5190-5192,
// 5350; PublicHeader.cpp:448-456, 473-481, 23-29, 45-51). gres counts clamp at
// 0xFFFF (no node can hold that many slots, so the verdict is unchanged).
CRANE_HD void view_node_plus_tasks(View& out, const View& node, const View& task, uint32_t t) {
  out.cpu_raw = node.cpu_raw + task.cpu_raw * (int64_t)t;
  out.mem = node.mem + task.mem * (uint64_t)t;
  out.mem_sw = node.mem_sw + task.mem_sw * (uint64_t)t;
  out.gtot[0] = out.gtot[1] = out.gspec[0] = out.gspec[1] = 0;
  for (uint32_t i = 0; i < 8; ++i) {
    uint64_t v = (uint64_t)field16(node.gtot, i) + (uint64_t)field16(task.gtot, i) * t;
    set_field16(out.gtot, i, (uint32_t)(v > 0xFFFF ? 0xFFFF : v));
    uint64_t s = (uint64_t)field16(node.gspec, i) + (uint64_t)field16(task.gspec, i) * t;
    set_field16(out.gspec, i, (uint32_t)(s > 0xFFFF ? 0xFFFF : s));
  }
}
CRANE_HD bool view_has_gres(const View& v) { return (v.gtot[0] | v.gtot[1] | v.gspec[0] | v.gspec[1]) != 0; }

// gres part of ResourceView::GetFeasibleResourceInNode (PublicHeader.cpp:545-595)
// on the two packed mask words `ag` of the available row. kAlloc=false: verdict
// only (a name is feasible iff every typed count fits its type and
// max(total, sum typed) fits the name's slots); kAlloc=true also picks the
// slots in the reference's order: typed first (dictionary order, deviation
// D4), leftovers of a typed entry serve the untyped part, then the other
// types in order.
template <bool kAlloc>
CRANE_HD bool feasible_gres(const View& req, const uint64_t* ag, const GresDict& d, uint64_t* out) {
#pragma unroll 1
  for (uint32_t g = 0; g < CRANE_GRES_NAMES; ++g) {
    // a requested name without dictionary entries has name_count 0: absent from every node
    const uint32_t e0 = d.name_first[g], e1 = e0 + d.name_count[g];
    const uint32_t want_total = field16(req.gtot, g);
    uint32_t typed_sum = 0, have_total = 0;
    bool wanted = want_total != 0;
#pragma unroll 1
    for (uint32_t e = e0; e < e1; ++e) {
      const uint32_t sp = field16(req.gspec, e);
      typed_sum += sp;
      wanted |= sp != 0;
      have_total += popc32(field16(ag, e));
    }
    if (!wanted) continue;
    if (have_total == 0) return false;  // name absent from avail (PH.cpp:551)
    uint32_t untyped = want_total > typed_sum ? want_total - typed_sum : 0;
#pragma unroll 1
    for (uint32_t e = e0; e < e1; ++e) {  // typed first (PH.cpp:563-579)
      const uint32_t sp = field16(req.gspec, e);
      if (sp == 0) continue;
      const uint32_t m = field16(ag, e);
      const uint32_t c = popc32(m);
      if (c < sp) return false;  // covers "type absent" (m == 0)
      uint32_t extra = c - sp;
      if (extra > untyped) extra = untyped;
      untyped -= extra;
      if (kAlloc) set_field16(out, e, (uint32_t)lowest_bits64(m, (int)(sp + extra)));
    }
    if (untyped > 0) {  // the other types (PH.cpp:581-592)
#pragma unroll 1
      for (uint32_t e = e0; e < e1 && untyped > 0; ++e) {
        if (field16(req.gspec, e) != 0) continue;
        const uint32_t m = field16(ag, e);
        const uint32_t c = popc32(m);
        const uint32_t take = c < untyped ? c : untyped;
        untyped -= take;
        if (kAlloc) set_field16(out, e, (uint32_t)lowest_bits64(m, (int)take));
      }
    }
    if (untyped != 0) return false;
  }
  return true;
}

// ResourceView::GetFeasibleResourceInNode (PublicHeader.cpp:519-599).
template <bool kAlloc>
CRANE_HD bool feasible(const View& req, const Row& avail, const GresDict& d, Row* alloc) {
  if (req.cpu_raw > avail.cpu_raw) return false;
  if (req.mem > avail.mem) return false;
  const int64_t whole = req.cpu_raw / 256;  // static_cast<int64_t>(cpu_t): truncates
  const bool integer_req = (whole * 256 == req.cpu_raw) && !core_empty(avail);
  if (integer_req && (int64_t)core_count(avail) < whole) return false;
  if (kAlloc) {
    row_zero(*alloc);
    alloc->cpu_raw = req.cpu_raw;
    alloc->mem = req.mem;
    alloc->mem_sw = req.mem_sw;
    if (integer_req) {
      int left = (int)whole;
#pragma unroll
      for (int w = 0; w < CRANE_CORE_WORDS; ++w) {
        if (left <= 0) break;
        const int c = popc64(avail.core[w]);
        const int take = c < left ? c : left;
        alloc->core[w] = lowest_bits64(avail.core[w], take);
        left -= take;
      }
    }
  }
  if (!view_has_gres(req)) return true;
  return feasible_gres<kAlloc>(req, avail.g, d, kAlloc ? alloc->g : nullptr);
}

// MinCpuTimeRatioFirst::UpdateCost delta (JobScheduler.h:46-48):
//   seconds * (double(res.cpu) / double(total.cpu)), double(cpu_t) = raw/256.0.
// Written with explicit rounding intrinsics so nvcc never contracts into FMA.
CRANE_HD double cost_delta(int64_t seconds, int64_t res_cpu_raw, int64_t total_cpu_raw) {
  // x / 256.0 is an exact power-of-two scaling, so it is written as a multiply
  // (bit-identical, and one fp64 division instead of three)
#if defined(__CUDA_ARCH__) || defined(CRANE_EMU)
  double a = __dmul_rn(__ll2double_rn(res_cpu_raw), 0.00390625);
  double b = __dmul_rn(__ll2double_rn(total_cpu_raw), 0.00390625);
  return __dmul_rn(__ll2double_rn(seconds), __ddiv_rn(a, b));
#else
  double a = (double)res_cpu_raw / 256.0;
  double b = (double)total_cpu_raw / 256.0;
  return (double)seconds * (a / b);
#endif
}

// IUpdateNodeCostPolicy::UpdateCost as a cost step: policy 0 MinCpuTimeRatioFirst
// (JobScheduler.h:40-54), policy 1 BestFit (ours, BASELINE config 4: the node's
// cost is its free cpu count over all allocations; integers as doubles, exact).
CRANE_HD double cost_step(uint32_t policy, int64_t seconds, int64_t res_cpu_raw, int64_t total_cpu_raw) {
#if defined(__CUDA_ARCH__) || defined(CRANE_EMU)
  if (policy == 1) return -__ll2double_rn(res_cpu_raw);
#else
  if (policy == 1) return -(double)res_cpu_raw;
#endif
  return cost_delta(seconds, res_cpu_raw, total_cpu_raw);
}

}  // namespace crane
