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
// the same entry: the pass is sequential only within one qos.
//   k_qos_keys    key = qos for the jobs the pass has to walk (started now, known
//                 qos), sentinel otherwise; "InvalidQOS" is written here (it
//                 touches no usage). A stable radix sort of (key, job) then
//                 gives every qos its jobs in job-id order.
//   k_qos_offsets the segment of every qos in the sorted list.
//   k_qos_chain   one CTA per qos. The usage entries of the qos — its column of
//                 the (user,qos) and (account,qos) tables and the qos entry —
//                 are staged in shared memory when they fit (else they stay in
//                 global memory: same code, generic pointers). Warps 1-3 sum
//                 the allocations of the next 96 jobs (ResourceV3::View) into a
//                 shared buffer while warp 0 walks the current 96 in order: the
//                 levels of one job are checked lane-parallel (lane 0 = user,
//                 lanes 1..C = account chain in order, lane C+1 = qos) and the
//                 lowest failing lane gives the reason, which is the
//                 reference's first-failure order.
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
__device__ __forceinline__ uint8_t qos_check_tres(const GresDict& d, uint32_t n_names, const QosUse& u, const crane_tres_limit_t& L) {
  if (u.cpu_raw > L.view.cpu_raw) return CRANE_REASON_QOS_CPU;
  if (u.mem > L.view.mem) return CRANE_REASON_QOS_MEM;
  // (all loops have constant bounds and are unrolled: u stays in registers; the uniform
  // early exits keep the work proportional to the cluster's dictionary)
  const uint32_t n_entries = d.n_entries;
#pragma unroll
  for (int g = 0; g < CRANE_GRES_NAMES; ++g) {
    if ((uint32_t)g >= n_names) continue;
    const uint32_t first = d.name_first[g], cnt = d.name_count[g];
    if (cnt == 0) continue;
    bool in_req = u.tot[g] != 0;
#pragma unroll
    for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) {
      if ((uint32_t)e >= n_entries) continue;
      if ((uint32_t)e >= first && (uint32_t)e < first + cnt) in_req |= u.spec[e] != 0;
    }
    if (!in_req) continue;
    if (!((L.gres_name_present >> g) & 1u)) return CRANE_REASON_NONE;
    if (u.tot[g] > L.view.gres_total[g]) return CRANE_REASON_QOS_GRES;
#pragma unroll
    for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) {
      if ((uint32_t)e >= n_entries) continue;
      if ((uint32_t)e < first || (uint32_t)e >= first + cnt || u.spec[e] == 0) continue;
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

constexpr int kQosBatch = 96;     // jobs prepared by warps 1-3 per round
constexpr int kQosThreads = 128;

__global__ void k_qos_keys(QosDev q, uint64_t* keys, uint32_t* vals, uint64_t sentinel) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= q.n_jobs) return;
  uint64_t key = sentinel;
  if (q.reason[i] == CRANE_REASON_NONE && q.n_alloc[i] != 0) {
    const uint32_t jq = q.qos[i];
    // "InvalidQOS" (AccountMetaContainer.cpp:168-169): no usage is touched
    if (jq >= q.n_qos || !q.valid[jq]) q.reason[i] = CRANE_REASON_QOS_INVALID;
    else key = jq;
  }
  keys[i] = key;
  vals[i] = i;
}

// off[w] = first position of key >= w in the sorted keys, w = 0..n_qos
__global__ void k_qos_offsets(const uint64_t* keys, uint32_t n, uint32_t n_qos, uint32_t* off) {
  const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w > n_qos) return;
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) / 2;
    if (keys[mid] < (uint64_t)w) lo = mid + 1; else hi = mid;
  }
  off[w] = lo;
}

struct QosPrep {   // what the walk needs of one job, prepared off the chain
  QosUse a;
  int64_t tl;
  uint32_t job, user, C, pad;
  uint16_t chain[30];  // account chain, in order
  uint16_t pad2[2];
};

__global__ void __launch_bounds__(kQosThreads) k_qos_chain(QosDev q, GresDict dict, const uint32_t* list, const uint32_t* off,
                                                           uint32_t tables_in_smem) {
  unsigned char* const qos_dyn = CRANE_DYN_BASE();  // dynamic shared memory: the usage column of this qos
  __shared__ QosPrep s_prep[2][kQosBatch];
  __shared__ crane_tres_limit_t s_lim[3];
  __shared__ crane_meta_resource_t s_qos_entry;
  const uint32_t tid = threadIdx.x, lane = tid & 31u, wid = tid >> 5;
  const uint32_t w = blockIdx.x;
  const uint32_t begin = off[w], end = off[w + 1];
  if (begin == end) return;  // (an invalid qos has no jobs in the list)
  const uint32_t Q = q.n_qos, U = q.n_users, A = q.n_accounts;
  crane_meta_resource_t* const s_user = reinterpret_cast<crane_meta_resource_t*>(qos_dyn);
  crane_meta_resource_t* const s_acct = s_user + U;
  constexpr uint32_t kWords = sizeof(crane_meta_resource_t) / 4;
  if (tables_in_smem) {
    for (uint32_t i = tid; i < (U + A) * kWords; i += kQosThreads) {
      const uint32_t e = i / kWords, k = i % kWords;
      const crane_meta_resource_t* src = e < U ? q.user_usage + (size_t)e * Q + w : q.account_usage + (size_t)(e - U) * Q + w;
      reinterpret_cast<uint32_t*>(s_user)[i] = reinterpret_cast<const uint32_t*>(src)[k];
    }
  }
  if (tid < kWords) reinterpret_cast<uint32_t*>(&s_qos_entry)[tid] = reinterpret_cast<const uint32_t*>(q.qos_usage + w)[tid];
  if (tid == 0) { s_lim[0] = q.tres_user[w]; s_lim[1] = q.tres_account[w]; s_lim[2] = q.tres_qos[w]; }
  const int64_t max_wall = q.max_wall[w];
  const uint32_t mj_user = q.max_jobs_per_user[w], mj_acct = q.max_jobs_per_account[w], mj_qos = q.max_jobs[w];
  const int64_t max_cpus_user = q.max_cpus_per_user_raw[w];
  uint32_t n_names = 0;  // names beyond the last entry's have no entries: nothing to check or to count
#pragma unroll
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e)
    if ((uint32_t)e < dict.n_entries && n_names < dict.entry_name[e] + 1u) n_names = dict.entry_name[e] + 1u;
  const uint32_t n_entries = dict.n_entries;

  // allocated_res.View() of the jobs [b0, b0 + kQosBatch) of the list, one job per thread of warps 1-3
  auto prepare = [&](uint32_t b0, uint32_t buf) {
    const uint32_t t = tid - 32u;
    if (b0 + t >= end) return;
    const uint32_t j = list[b0 + t];
    QosPrep& P = s_prep[buf][t];
    const uint32_t n = q.n_alloc[j];
    const Row* rows = q.alloc_res + q.alloc_off[j];
    QosUse a;
    a.cpu_raw = 0; a.mem = 0; a.mem_sw = 0;
#pragma unroll
    for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) a.spec[e] = 0;
    for (uint32_t k = 0; k < n; ++k) {
      const Row r = rows[k];
      a.cpu_raw += r.cpu_raw; a.mem += r.mem; a.mem_sw += r.mem_sw;
#pragma unroll
      for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) a.spec[e] += (uint32_t)popc32(field16(r.g, e));
    }
#pragma unroll
    for (int g = 0; g < CRANE_GRES_NAMES; ++g) a.tot[g] = 0;
#pragma unroll
    for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) {
      if ((uint32_t)e >= dict.n_entries) continue;
      const uint32_t name = dict.entry_name[e];
#pragma unroll
      for (int g = 0; g < CRANE_GRES_NAMES; ++g)
        if (name == (uint32_t)g) a.tot[g] += a.spec[e];
    }
    P.a = a;
    P.tl = q.time_limit[j];
    P.job = j;
    P.user = q.user[j];
    const uint32_t c0 = q.chain_off[j];
    P.C = q.chain_off[j + 1] - c0;
    for (uint32_t c = 0; c < P.C; ++c) P.chain[c] = (uint16_t)q.chain_acct[c0 + c];
  };

  if (wid) prepare(begin, 0);
  __syncthreads();
  uint32_t buf = 0;
  for (uint32_t b0 = begin; b0 < end; b0 += kQosBatch, buf ^= 1u) {
    if (wid) {
      prepare(b0 + kQosBatch, buf ^ 1u);
    } else {
      const uint32_t cnt = end - b0 < (uint32_t)kQosBatch ? end - b0 : (uint32_t)kQosBatch;
      for (uint32_t k = 0; k < cnt; ++k) {
        const QosPrep& P = s_prep[buf][k];
        // (the gres counts of the allocation are read only where they are needed: most
        // refusals are decided by the cpu / job-count / wall-time limits)
        const int64_t a_cpu = P.a.cpu_raw;
        const uint64_t a_mem = P.a.mem, a_mem_sw = P.a.mem_sw;
        const int64_t tl = P.tl;
        const uint32_t C = P.C;
        // ---- CheckQosResource_: one level per lane ---------------------------
        crane_meta_resource_t* val = nullptr;
        const crane_tres_limit_t* lim = nullptr;
        uint32_t max_jobs = 0;
        if (lane == 0) {
          val = tables_in_smem ? s_user + P.user : q.user_usage + (size_t)P.user * Q + w;
          lim = &s_lim[0];
          max_jobs = mj_user;
        } else if (lane <= C) {
          const uint32_t acct = P.chain[lane - 1];
          val = tables_in_smem ? s_acct + acct : q.account_usage + (size_t)acct * Q + w;
          lim = &s_lim[1];
          max_jobs = mj_acct;
        } else if (lane == C + 1) {
          val = &s_qos_entry;
          lim = &s_lim[2];
          max_jobs = mj_qos;
        }
        uint8_t result = CRANE_REASON_NONE;
        int64_t u_cpu = a_cpu;  // resource_use = allocated view + val.resource
        uint64_t u_mem = a_mem, u_mem_sw = a_mem_sw;
        if (val) {
          u_cpu += val->cpu_raw; u_mem += val->mem; u_mem_sw += val->mem_sw;
          if (lane == 0 && u_cpu > max_cpus_user) result = CRANE_REASON_QOS_CPU;
          else if ((uint64_t)val->jobs_count + 1ull > (uint64_t)max_jobs) result = CRANE_REASON_QOS_JOBS;
          else if (max_wall > 0 && val->wall_time + tl > max_wall) result = CRANE_REASON_QOS_WALL;
        }
        // the reason is the lowest failing level's: levels behind one that already failed
        // the cheap tests need no tres check
        const uint32_t early = __ballot_sync(0xffffffffu, result != CRANE_REASON_NONE);
        const uint32_t lowest = early ? (uint32_t)__ffs((int)early) - 1u : 32u;
        if (val && result == CRANE_REASON_NONE && lane < lowest) {
          QosUse u;
          u.cpu_raw = u_cpu; u.mem = u_mem; u.mem_sw = u_mem_sw;
#pragma unroll
          for (int g = 0; g < CRANE_GRES_NAMES; ++g) u.tot[g] = (uint32_t)g < n_names ? P.a.tot[g] + val->gres_total[g] : 0u;
#pragma unroll
          for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) u.spec[e] = (uint32_t)e < n_entries ? P.a.spec[e] + val->gres_spec[e] : 0u;
          result = qos_check_tres(dict, n_names, u, *lim);
        }
        const uint32_t failed = __ballot_sync(0xffffffffu, result != CRANE_REASON_NONE);
        if (failed) {
          const uint8_t first = (uint8_t)__shfl_sync(0xffffffffu, (uint32_t)result, __ffs((int)failed) - 1);
          if (lane == 0) q.reason[P.job] = first;
        } else if (val) {
          // DoMallocResource_ / MetaResource::operator+= (AccountMetaContainer.cpp:33-39, 546-587)
          val->cpu_raw += a_cpu; val->mem += a_mem; val->mem_sw += a_mem_sw;
#pragma unroll
          for (int g = 0; g < CRANE_GRES_NAMES; ++g)
            if ((uint32_t)g < n_names) val->gres_total[g] += P.a.tot[g];
#pragma unroll
          for (int e = 0; e < CRANE_GRES_ENTRIES; ++e)
            if ((uint32_t)e < n_entries) val->gres_spec[e] += P.a.spec[e];
          val->jobs_count += 1;
          val->wall_time += tl;
        }
        __syncwarp();  // usage written by one lane is read by another for the next job
      }
    }
    __syncthreads();
  }
  // ---- usage back to the tables --------------------------------------------------
  if (tables_in_smem) {
    for (uint32_t i = tid; i < (U + A) * kWords; i += kQosThreads) {
      const uint32_t e = i / kWords, k = i % kWords;
      crane_meta_resource_t* dst = e < U ? q.user_usage + (size_t)e * Q + w : q.account_usage + (size_t)(e - U) * Q + w;
      reinterpret_cast<uint32_t*>(dst)[k] = reinterpret_cast<const uint32_t*>(s_user)[i];
    }
  }
  if (tid < kWords) reinterpret_cast<uint32_t*>(q.qos_usage + w)[tid] = reinterpret_cast<const uint32_t*>(&s_qos_entry)[tid];
}

}  // namespace crane
