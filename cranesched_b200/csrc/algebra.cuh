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

typedef crane_res_in_node_t Row;  // ResourceInNodeV3 (PublicHeader.h:562)
typedef crane_res_view_t View;    // ResourceView     (PublicHeader.h:671)

struct GresDict {
  uint32_t n_entries;
  uint8_t entry_name[CRANE_GRES_ENTRIES];
};

CRANE_HD int popc16(uint32_t v) {
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

// the n lowest set bits of m (n <= popcount(m)); "it = begin(); n times ++it"
// over an ordered std::set (PublicHeader.cpp:535-537, 572-574).
CRANE_HD uint32_t lowest_bits32(uint32_t m, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) {
    uint32_t low = m & (0u - m);
    r |= low;
    m ^= low;
  }
  return r;
}
CRANE_HD uint64_t lowest_bits64(uint64_t m, int n) {
  if (n >= 64) return m;
  uint64_t r = 0;
  for (int i = 0; i < n; ++i) {
    uint64_t low = m & (0ull - m);
    r |= low;
    m ^= low;
  }
  return r;
}

CRANE_HD void row_zero(Row& r) {
  r.cpu_raw = 0; r.mem = 0; r.mem_sw = 0;
  for (int w = 0; w < CRANE_CORE_WORDS; ++w) r.core[w] = 0;
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) r.gres[e] = 0;
}
CRANE_HD bool core_empty(const Row& r) {
  uint64_t any = 0;
  for (int w = 0; w < CRANE_CORE_WORDS; ++w) any |= r.core[w];
  return any == 0;
}
CRANE_HD int core_count(const Row& r) {
  int n = 0;
  for (int w = 0; w < CRANE_CORE_WORDS; ++w) n += popc64(r.core[w]);
  return n;
}

// ResourceInNodeV3::operator-= (PublicHeader.cpp:789-796; CpuSet 758-766;
// DedicatedResourceInNode 204-217; TypeSlotsMap 316-328): tolerant core erase,
// counts subtract, slot sets difference (empty entries vanish = mask 0).
CRANE_HD void row_sub(Row& a, const Row& b) {
  a.cpu_raw -= b.cpu_raw;
  a.mem -= b.mem;
  a.mem_sw -= b.mem_sw;
  for (int w = 0; w < CRANE_CORE_WORDS; ++w) a.core[w] &= ~b.core[w];
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) a.gres[e] &= (uint16_t)~b.gres[e];
}
// ResourceInNodeV3::operator+= (PublicHeader.cpp:781-787)
CRANE_HD void row_add(Row& a, const Row& b) {
  a.cpu_raw += b.cpu_raw;
  a.mem += b.mem;
  a.mem_sw += b.mem_sw;
  for (int w = 0; w < CRANE_CORE_WORDS; ++w) a.core[w] |= b.core[w];
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) a.gres[e] |= b.gres[e];
}
// operator<=(ResourceInNodeV3, ResourceInNodeV3) (PublicHeader.cpp:886-890,
// 159-169, 334-343): cpu, mem, slot-set inclusion. core ids and mem_sw are NOT
// compared.
CRANE_HD bool row_le(const Row& a, const Row& b) {
  if (a.cpu_raw > b.cpu_raw) return false;
  if (a.mem > b.mem) return false;
  uint32_t bad = 0;
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) bad |= (uint32_t)(a.gres[e] & (uint16_t)~b.gres[e]);
  return bad == 0;
}
// ResourceInNodeV3::Ckmin (PublicHeader.cpp:815-827), literal form.
CRANE_HD void row_ckmin(Row& a, const Row& b) {
  if (b.cpu_raw < a.cpu_raw) a.cpu_raw = b.cpu_raw;
  if (!core_empty(a) && !core_empty(b))
    for (int w = 0; w < CRANE_CORE_WORDS; ++w) a.core[w] &= b.core[w];
  if (b.mem < a.mem) a.mem = b.mem;
  if (b.mem_sw < a.mem_sw) a.mem_sw = b.mem_sw;
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) a.gres[e] &= b.gres[e];
}

// ---- prefix form of Ckmin -------------------------------------------------
// The window minimum of JobScheduler.cpp:5314-5319 is the left fold
//   row = res_avail; for seg in window: row.Ckmin(seg)
// Its core rule ("intersect only when both sides are non-empty") is not
// associative as written, but the fold equals
//   core = res_avail.core & AND{ seg.core : seg.core != {} }
// (once the running set is empty it stays empty either way). So the per-node
// prefix array `pm` stores, for the core field, the AND over the non-empty
// segment masks with all-ones as the identity; every other field is a plain
// min / AND. pm_identity() is the fold's neutral element.
CRANE_HD void pm_identity(Row& r) {
  r.cpu_raw = INT64_MAX; r.mem = UINT64_MAX; r.mem_sw = UINT64_MAX;
  for (int w = 0; w < CRANE_CORE_WORDS; ++w) r.core[w] = ~0ull;
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) r.gres[e] = 0xFFFF;
}
// acc = acc (+) seg, seg a raw timeline segment
CRANE_HD void pm_absorb(Row& acc, const Row& seg) {
  if (seg.cpu_raw < acc.cpu_raw) acc.cpu_raw = seg.cpu_raw;
  if (seg.mem < acc.mem) acc.mem = seg.mem;
  if (seg.mem_sw < acc.mem_sw) acc.mem_sw = seg.mem_sw;
  if (!core_empty(seg))
    for (int w = 0; w < CRANE_CORE_WORDS; ++w) acc.core[w] &= seg.core[w];
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) acc.gres[e] &= seg.gres[e];
}
// acc = acc (+) other, both already in prefix form (associative combine)
CRANE_HD void pm_combine(Row& acc, const Row& o) {
  if (o.cpu_raw < acc.cpu_raw) acc.cpu_raw = o.cpu_raw;
  if (o.mem < acc.mem) acc.mem = o.mem;
  if (o.mem_sw < acc.mem_sw) acc.mem_sw = o.mem_sw;
  for (int w = 0; w < CRANE_CORE_WORDS; ++w) acc.core[w] &= o.core[w];
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) acc.gres[e] &= o.gres[e];
}
// the window-min row: fold started from the (stale) tick-start res_avail
CRANE_HD void window_row(Row& out, const Row& avail0, const Row& pm) {
  out = avail0;
  pm_combine(out, pm);
}

// ---- requests ---------------------------------------------------------------
// req_node_res_view + req_task_res_view * t  (JobScheduler.cpp:5190-5192,
// 5350; PublicHeader.cpp:448-456, 473-481, 23-29, 45-51). gres counts clamp at
// 0xFFFF (no node can hold that many slots, so the verdict is unchanged).
CRANE_HD void view_node_plus_tasks(View& out, const View& node, const View& task, uint32_t t) {
  out.cpu_raw = node.cpu_raw + task.cpu_raw * (int64_t)t;
  out.mem = node.mem + task.mem * (uint64_t)t;
  out.mem_sw = node.mem_sw + task.mem_sw * (uint64_t)t;
  for (int g = 0; g < CRANE_GRES_NAMES; ++g) {
    uint64_t v = (uint64_t)node.gres_total[g] + (uint64_t)task.gres_total[g] * t;
    out.gres_total[g] = (uint16_t)(v > 0xFFFF ? 0xFFFF : v);
  }
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) {
    uint64_t v = (uint64_t)node.gres_spec[e] + (uint64_t)task.gres_spec[e] * t;
    out.gres_spec[e] = (uint16_t)(v > 0xFFFF ? 0xFFFF : v);
  }
}

// ResourceView::GetFeasibleResourceInNode (PublicHeader.cpp:519-599).
// kAlloc=false evaluates only the verdict (counts suffice: a name is feasible
// iff every typed count fits its type and max(total, sum typed) fits the
// name's slots); kAlloc=true also picks the concrete cores/slots in the
// reference's order: typed first (dictionary order, deviation D4), leftovers of
// a typed entry serve the untyped part, then the other types in order.
template <bool kAlloc>
CRANE_HD bool feasible(const View& req, const Row& avail, const GresDict& d, Row* alloc) {
  if (req.cpu_raw > avail.cpu_raw) return false;
  if (req.mem > avail.mem) return false;

  int64_t whole = req.cpu_raw / 256;  // static_cast<int64_t>(cpu_t): truncates
  bool integer_req = (whole * 256 == req.cpu_raw) && !core_empty(avail);
  if (integer_req) {
    if ((int64_t)core_count(avail) < whole) return false;
  }
  if (kAlloc) {
    row_zero(*alloc);
    alloc->cpu_raw = req.cpu_raw;
    alloc->mem = req.mem;
    alloc->mem_sw = req.mem_sw;
    if (integer_req) {
      int left = (int)whole;
      for (int w = 0; w < CRANE_CORE_WORDS && left > 0; ++w) {
        int c = popc64(avail.core[w]);
        int take = c < left ? c : left;
        alloc->core[w] = lowest_bits64(avail.core[w], take);
        left -= take;
      }
    }
  }
  for (int g = 0; g < CRANE_GRES_NAMES; ++g) {
    uint32_t typed_sum = 0, have_total = 0;
    bool wanted = req.gres_total[g] != 0;
    for (uint32_t e = 0; e < d.n_entries; ++e) {
      if (d.entry_name[e] != g) continue;
      typed_sum += req.gres_spec[e];
      if (req.gres_spec[e]) wanted = true;
      have_total += popc16(avail.gres[e]);
    }
    if (!wanted) continue;
    if (have_total == 0) return false;  // name absent from avail (PH.cpp:551)
    uint32_t untyped = req.gres_total[g] > typed_sum ? req.gres_total[g] - typed_sum : 0;
    for (uint32_t e = 0; e < d.n_entries; ++e) {  // typed first (PH.cpp:563-579)
      if (d.entry_name[e] != g || req.gres_spec[e] == 0) continue;
      uint32_t m = avail.gres[e];
      uint32_t c = popc16(m);
      if (c < req.gres_spec[e]) return false;  // covers "type absent" (m == 0)
      uint32_t extra = c - req.gres_spec[e];
      if (extra > untyped) extra = untyped;
      untyped -= extra;
      if (kAlloc) alloc->gres[e] = (uint16_t)lowest_bits32(m, (int)(req.gres_spec[e] + extra));
    }
    if (untyped > 0) {  // the other types (PH.cpp:581-592)
      for (uint32_t e = 0; e < d.n_entries && untyped > 0; ++e) {
        if (d.entry_name[e] != g || req.gres_spec[e] != 0) continue;
        uint32_t m = avail.gres[e];
        uint32_t c = popc16(m);
        uint32_t take = c < untyped ? c : untyped;
        untyped -= take;
        if (kAlloc) alloc->gres[e] = (uint16_t)lowest_bits32(m, (int)take);
      }
    }
    if (untyped != 0) return false;
  }
  return true;
}

// MinCpuTimeRatioFirst::UpdateCost delta (JobScheduler.h:46-48):
//   seconds * (double(res.cpu) / double(total.cpu)), double(cpu_t) = raw/256.0.
// Written with explicit rounding intrinsics so nvcc never contracts into FMA.
CRANE_HD double cost_delta(int64_t seconds, int64_t res_cpu_raw, int64_t total_cpu_raw) {
#if defined(__CUDA_ARCH__) || defined(CRANE_EMU)
  double a = __ddiv_rn(__ll2double_rn(res_cpu_raw), 256.0);
  double b = __ddiv_rn(__ll2double_rn(total_cpu_raw), 256.0);
  return __dmul_rn(__ll2double_rn(seconds), __ddiv_rn(a, b));
#else
  double a = (double)res_cpu_raw / 256.0;
  double b = (double)total_cpu_raw / 256.0;
  return (double)seconds * (a / b);
#endif
}

}  // namespace crane
