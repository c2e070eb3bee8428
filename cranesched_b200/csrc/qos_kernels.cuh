// qos_kernels.cuh — the QoS post-filter of the commit loop.
//
// Reference: JobScheduler.cpp:1262 calls, for every job NodeSelect starts now
// and in job-id order, AccountMetaContainer::CheckAndMallocQosResource
// (Accounting/AccountMetaContainer.cpp:164-191): CheckQosResource_ (382-491)
// tests the job's allocation + current usage against the qos limits at three
// levels — (user,qos), every (account,qos) of the account chain, qos — and
// DoMallocResource_ (546-587) adds the job to the usage at all of them.
//
// Every usage entry is keyed by the qos, so jobs of different qos never touch
// the same entry: the pass is sequential only within one qos. One warp owns
// one qos id and walks the job table in job-id order (ballot-compacted, 32
// jobs per coalesced probe); for one job the levels are checked lane-parallel
// (lane 0 = user, lanes 1..C = account chain in order, lane C+1 = qos) and the
// lowest failing lane gives the reason, which is the reference's first-failure
// order.
#pragma once

#include "algebra.cuh"

namespace crane {

struct QosDev {
  uint32_t n_qos, n_users, n_accounts, n_jobs;
  const uint8_t* valid;
  const uint32_t* max_jobs_per_user;
  const uint32_t* max_jobs_per_account;
  const uint32_t* max_jobs;
  const int64_t* max_cpus_per_user_raw;
  const int64_t* max_wall;
  const crane_tres_limit_t* tres_user;
  const crane_tres_limit_t* tres_account;
  const crane_tres_limit_t* tres_qos;
  const uint32_t* chain_off;
  const uint32_t* chain_acct;
  crane_meta_resource_t* user_usage;
  crane_meta_resource_t* account_usage;
  crane_meta_resource_t* qos_usage;
  // job columns (device-resident inputs / outputs of the last run)
  const uint32_t* qos;
  const uint32_t* user;
  const int64_t* time_limit;
  const uint32_t* n_alloc;
  const uint32_t* alloc_off;
  const Row* alloc_res;
  uint8_t* reason;
};

// the job's allocation as a ResourceView (ResourceV3::View,
// PublicHeader.cpp:946-952, 399-427) plus what MetaResource carries
struct QosUse {
  int64_t cpu_raw;
  uint64_t mem, mem_sw;
  uint32_t tot[CRANE_GRES_NAMES];
  uint32_t spec[CRANE_GRES_ENTRIES];
};

// AccountMetaContainer::CheckTres_ / CheckGres_ (AccountMetaContainer.cpp:493-531).
// A name / type is "in the request map" iff its count is non-zero; names and
// types are walked in dictionary order (deviation D8).
__device__ __forceinline__ uint8_t qos_check_tres(const GresDict& d, const QosUse& u, const crane_tres_limit_t& L) {
  if (u.cpu_raw > L.view.cpu_raw) return CRANE_REASON_QOS_CPU;
  if (u.mem > L.view.mem) return CRANE_REASON_QOS_MEM;
  for (uint32_t g = 0; g < CRANE_GRES_NAMES; ++g) {
    const uint32_t first = d.name_first[g], cnt = d.name_count[g];
    if (cnt == 0) continue;
    bool in_req = u.tot[g] != 0;
    for (uint32_t e = first; e < first + cnt; ++e) in_req |= u.spec[e] != 0;
    if (!in_req) continue;
    if (!((L.gres_name_present >> g) & 1u)) return CRANE_REASON_NONE;
    if (u.tot[g] > L.view.gres_total[g]) return CRANE_REASON_QOS_GRES;
    for (uint32_t e = first; e < first + cnt; ++e) {
      if (u.spec[e] == 0) continue;
      if (!((L.gres_spec_present >> e) & 1u)) return CRANE_REASON_NONE;
      if (u.spec[e] > L.view.gres_spec[e]) return CRANE_REASON_QOS_GRES;
    }
  }
  return CRANE_REASON_NONE;
}

template <class T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

__global__ void __launch_bounds__(32) k_qos_filter(QosDev q, GresDict dict) {
  const uint32_t lane = threadIdx.x;
  const uint32_t w = blockIdx.x;
  const uint32_t N = q.n_jobs;
  if (w == q.n_qos) {
    // "InvalidQOS" (AccountMetaContainer.cpp:168-169): no usage is touched, so
    // the jobs of all unknown / deleted qos are independent
    for (uint32_t i = lane; i < N; i += 32) {
      const uint32_t jq = q.qos[i];
      if (q.reason[i] == CRANE_REASON_NONE && q.n_alloc[i] != 0 && (jq >= q.n_qos || !q.valid[jq]))
        q.reason[i] = CRANE_REASON_QOS_INVALID;
    }
    return;
  }
  if (!q.valid[w]) return;
  const int64_t max_wall = q.max_wall[w];
  for (uint32_t base = 0; base < N; base += 32) {
    const uint32_t i = base + lane;
    const bool mine = i < N && q.qos[i] == w && q.reason[i] == CRANE_REASON_NONE && q.n_alloc[i] != 0;
    uint32_t todo = __ballot_sync(0xffffffffu, mine);
    while (todo) {
      const uint32_t j = base + (uint32_t)__ffs((int)todo) - 1u;
      todo &= todo - 1u;
      // ---- allocated_res.View() -------------------------------------------
      const uint32_t n = q.n_alloc[j];
      const Row* rows = q.alloc_res + q.alloc_off[j];
      QosUse a;
      a.cpu_raw = 0; a.mem = 0; a.mem_sw = 0;
#pragma unroll
      for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) a.spec[e] = 0;
      if (n == 1) {
        const Row r = rows[0];
        a.cpu_raw = r.cpu_raw; a.mem = r.mem; a.mem_sw = r.mem_sw;
#pragma unroll
        for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) a.spec[e] = (uint32_t)popc32(field16(r.g, e));
      } else {
        for (uint32_t k = lane; k < n; k += 32) {
          const Row r = rows[k];
          a.cpu_raw += r.cpu_raw; a.mem += r.mem; a.mem_sw += r.mem_sw;
#pragma unroll
          for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) a.spec[e] += (uint32_t)popc32(field16(r.g, e));
        }
        a.cpu_raw = warp_sum(a.cpu_raw); a.mem = warp_sum(a.mem); a.mem_sw = warp_sum(a.mem_sw);
#pragma unroll
        for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) a.spec[e] = warp_sum(a.spec[e]);
      }
#pragma unroll
      for (int g = 0; g < CRANE_GRES_NAMES; ++g) a.tot[g] = 0;
      for (uint32_t e = 0; e < dict.n_entries; ++e) a.tot[dict.entry_name[e]] += a.spec[e];
      // ---- CheckQosResource_: one level per lane ---------------------------
      const uint32_t c0 = q.chain_off[j], C = q.chain_off[j + 1] - c0;
      const int64_t tl = q.time_limit[j];
      crane_meta_resource_t* val = nullptr;
      const crane_tres_limit_t* lim = nullptr;
      uint32_t max_jobs = 0;
      if (lane == 0) {
        val = q.user_usage + (size_t)q.user[j] * q.n_qos + w;
        lim = q.tres_user + w;
        max_jobs = q.max_jobs_per_user[w];
      } else if (lane <= C) {
        val = q.account_usage + (size_t)q.chain_acct[c0 + lane - 1] * q.n_qos + w;
        lim = q.tres_account + w;
        max_jobs = q.max_jobs_per_account[w];
      } else if (lane == C + 1) {
        val = q.qos_usage + w;
        lim = q.tres_qos + w;
        max_jobs = q.max_jobs[w];
      }
      uint8_t result = CRANE_REASON_NONE;
      if (val) {
        QosUse u = a;  // resource_use = allocated view + val.resource
        u.cpu_raw += val->cpu_raw; u.mem += val->mem; u.mem_sw += val->mem_sw;
#pragma unroll
        for (int g = 0; g < CRANE_GRES_NAMES; ++g) u.tot[g] += val->gres_total[g];
#pragma unroll
        for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) u.spec[e] += val->gres_spec[e];
        if (lane == 0 && u.cpu_raw > q.max_cpus_per_user_raw[w]) result = CRANE_REASON_QOS_CPU;
        else if ((uint64_t)val->jobs_count + 1ull > (uint64_t)max_jobs) result = CRANE_REASON_QOS_JOBS;
        else if (max_wall > 0 && val->wall_time + tl > max_wall) result = CRANE_REASON_QOS_WALL;
        else result = qos_check_tres(dict, u, *lim);
      }
      const uint32_t failed = __ballot_sync(0xffffffffu, result != CRANE_REASON_NONE);
      if (failed) {
        const uint8_t first = (uint8_t)__shfl_sync(0xffffffffu, (uint32_t)result, __ffs((int)failed) - 1);
        if (lane == 0) q.reason[j] = first;
      } else if (val) {
        // DoMallocResource_ / MetaResource::operator+= (AccountMetaContainer.cpp:33-39, 546-587)
        val->cpu_raw += a.cpu_raw; val->mem += a.mem; val->mem_sw += a.mem_sw;
#pragma unroll
        for (int g = 0; g < CRANE_GRES_NAMES; ++g) val->gres_total[g] += a.tot[g];
#pragma unroll
        for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) val->gres_spec[e] += a.spec[e];
        val->jobs_count += 1;
        val->wall_time += tl;
      }
      __syncwarp();  // usage written by one lane is read by another for the next job
    }
  }
}

}  // namespace crane
