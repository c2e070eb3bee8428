// sched_kernels.cuh — sm_100a kernels of the CraneCtld scheduling hot path.
//
// Reference functions restated here:
//   R3  NodeState::InitTimeAvailResMap           JobScheduler.h:295-332
//   R4  MinCpuTimeRatioFirst / NodeSelector cost JobScheduler.h:40-54,492-532
//   R5  BasicPriority                            JobScheduler.h:177-195
//   R6  MultiFactorPriority                      JobScheduler.cpp:6526-6739
//   R7  LocalScheduler::GetNodesAndTrySchedule_  JobScheduler.cpp:5185-5405
//   R8  ResourceView::GetFeasibleResourceInNode  PublicHeader.cpp:519-599 (algebra.cuh)
//   R9  EarliestStartSubsetSelector              JobScheduler.h:786-859
//   R10 NodeState::UpdateResourceInNode          JobScheduler.h:334-453
//   R11 SchedulerAlgo::NodeSelect job loop       JobScheduler.cpp:5777-5867
//
// Data layout in HBM (all per tick):
//   node slot g = compact index of an alive && !drain node inside its
//   partition's contiguous range [part_base[p], part_base[p+1]).
//   tl_time[g][CAP] int64, tl_seg[g][CAP] Row, tl_pm[g][CAP] Row (prefix
//   Ckmin), tl_n[g]; CAP = max_jobs_per_node + 1.
#pragma once

#include "algebra.cuh"

#ifndef CRANE_EMU
#define CRANE_DYN_SMEM(T, name) extern __shared__ __align__(16) unsigned char name##_raw[]; T* name = reinterpret_cast<T*>(name##_raw)
#define CRANE_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

namespace crane {

constexpr int64_t kInf = INT64_MAX;
constexpr uint32_t kFullMask = 0xffffffffu;

__constant__ GresDict c_dict;

// ------------------------------------------------------------------------
// device-side tables
// ------------------------------------------------------------------------
struct PendingDev {  // SoA mirror of crane_pending_t
  uint32_t n;
  const uint32_t* partition;
  const int64_t* time_limit;
  const int64_t* submit_time;
  const uint32_t* node_num;
  const uint32_t* ntasks_per_node_min;
  const uint8_t* exclusive;
  const uint32_t* partition_priority;
  const uint32_t* qos_priority;
  const uint32_t* account;
  const double* mandated_priority;
  const View* req_node;
  const View* req_task;
  const View* req_total;
  const uint32_t* incl_off;
  const uint32_t* incl_nodes;
  const uint32_t* excl_off;
  const uint32_t* excl_nodes;
  const uint32_t* alloc_off;  // exclusive prefix sum of node_num
};

struct RunningDev {
  uint32_t n;
  const int64_t* start_time;
  const int64_t* end_time;
  const uint32_t* node_num;
  const uint32_t* partition_priority;
  const uint32_t* qos_priority;
  const uint32_t* account;
  const int64_t* view_cpu_raw;
  const uint64_t* view_mem;
  // running allocations regrouped by node slot (input order kept)
  const uint32_t* slot_off;   // [n_slots+1]
  const int64_t* slot_end;    // end_time of the owning job (unclamped)
  const Row* slot_res;
  // running jobs regrouped by account (input order kept)
  uint32_t n_accounts;
  const uint32_t* acc_off;    // [n_accounts+1]
  const uint32_t* acc_job;
  const uint8_t* acc_present; // account appears in pending or running
};

struct ClusterDev {
  uint32_t n_slots;           // usable nodes
  uint32_t n_parts;
  uint32_t max_part_slots;
  const uint32_t* part_base;  // [n_parts+1] slot ranges
  const uint32_t* slot_node;  // slot -> global node index
  const uint32_t* node_slot;  // global node -> slot or 0xffffffff
  const Row* slot_total;      // res_total per slot
};

struct TimelineDev {
  uint32_t cap;               // entries per slot
  uint32_t* n;                // [n_slots]
  int64_t* time;              // [n_slots][cap]
  Row* seg;                   // [n_slots][cap]
  Row* pm;                    // [n_slots][cap]
  Row* avail0;                // [n_slots] tick-start res_avail (NodeState::res_avail)
  double* cost0;              // [n_slots] initial cost (NodeRater)
  uint8_t* skip;              // [n_slots] timeline size >= max_jobs_per_node
};

// per-job record in final queue order (partition-major, priority order inside)
struct __align__(16) JobQ {
  View req;            // req_node + req_task * ntasks_per_node (min_res_view)
  int64_t time_limit;
  uint32_t job;        // index into the pending table
  uint32_t node_num;
  uint32_t alloc_off;
  uint32_t ntasks_per_node;
  uint32_t flags;      // bit0 exclusive, bit1 has gres request
  uint32_t pad[3];
};
static_assert(sizeof(JobQ) == 96, "JobQ layout");

struct PlaceDev {
  uint8_t* reason;
  double* priority;
  int64_t* start_time;
  int64_t* end_time;
  uint32_t* n_alloc;
  uint32_t* alloc_node;
  uint32_t* alloc_ntasks;
  Row* alloc_res;
};

struct Bounds {  // MultiFactorPriority::FactorBound, JobScheduler.h:206-216
  unsigned long long age_max, age_min;
  unsigned long long qos_max, qos_min;
  unsigned long long part_max, part_min;
  unsigned long long nodes_max, nodes_min;
  unsigned long long mem_max, mem_min;
  unsigned long long cpus_max, cpus_min;  // raw cpu_t (monotone in the double)
  unsigned long long svc_max_bits, svc_min_bits;  // non-negative doubles as bits
};

// ------------------------------------------------------------------------
// warp helpers
// ------------------------------------------------------------------------
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  return (uint64_t)__shfl_sync(kFullMask, (unsigned long long)v, src);
}
__device__ __forceinline__ int64_t shfl_i64(int64_t v, int src) {
  return (int64_t)__shfl_sync(kFullMask, (long long)v, src);
}
__device__ __forceinline__ void row_shfl(Row& out, const Row& in, int src) {
  out.cpu_raw = shfl_i64(in.cpu_raw, src);
  out.mem = shfl_u64(in.mem, src);
  out.mem_sw = shfl_u64(in.mem_sw, src);
#pragma unroll
  for (int w = 0; w < CRANE_CORE_WORDS; ++w) out.core[w] = shfl_u64(in.core[w], src);
  uint64_t g0, g1;
  g0 = (uint64_t)in.gres[0] | (uint64_t)in.gres[1] << 16 | (uint64_t)in.gres[2] << 32 | (uint64_t)in.gres[3] << 48;
  g1 = (uint64_t)in.gres[4] | (uint64_t)in.gres[5] << 16 | (uint64_t)in.gres[6] << 32 | (uint64_t)in.gres[7] << 48;
  g0 = shfl_u64(g0, src);
  g1 = shfl_u64(g1, src);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    out.gres[e] = (uint16_t)(g0 >> (16 * e));
    out.gres[4 + e] = (uint16_t)(g1 >> (16 * e));
  }
}

// packed per-entry slot counts of a row (8 x u8) for the cheap pre-filter
__device__ __forceinline__ uint64_t pack_gres_counts(const Row& r) {
  uint64_t p = 0;
#pragma unroll
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) p |= (uint64_t)popc16(r.gres[e]) << (8 * e);
  return p;
}
// conservative count-only gres check against packed counts (same verdict as
// feasible<false> restricted to gres)
__device__ __forceinline__ bool gres_counts_ok(const View& req, uint64_t packed) {
  for (int g = 0; g < CRANE_GRES_NAMES; ++g) {
    uint32_t typed = 0, have = 0;
    bool wanted = req.gres_total[g] != 0;
    bool ok = true;
    for (uint32_t e = 0; e < c_dict.n_entries; ++e) {
      if (c_dict.entry_name[e] != g) continue;
      uint32_t c = (uint32_t)(packed >> (8 * e)) & 0xff;
      typed += req.gres_spec[e];
      if (req.gres_spec[e]) wanted = true;
      if (c < req.gres_spec[e]) ok = false;
      have += c;
    }
    if (!wanted) continue;
    uint32_t need = req.gres_total[g] > typed ? req.gres_total[g] : typed;
    if (!ok || have == 0 || have < need) return false;
  }
  return true;
}

// ------------------------------------------------------------------------
// K-prio part 1: factor bounds (JobScheduler.cpp:6553-6633)
// ------------------------------------------------------------------------
__global__ void k_bounds_init(Bounds* b) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    b->age_max = 0; b->age_min = ~0ull;
    b->qos_max = 0; b->qos_min = 0xffffffffull;
    b->part_max = 0; b->part_min = 0xffffffffull;
    b->nodes_max = 0; b->nodes_min = 0xffffffffull;
    b->mem_max = 0; b->mem_min = ~0ull;
    b->cpus_max = 0; b->cpus_min = ~0ull;  // "double max" sentinel handled at use
    b->svc_max_bits = 0;                               // 0.0
    b->svc_min_bits = 0x41EFFFFFFFE00000ull;           // 4294967295.0
  }
}

__device__ __forceinline__ void warp_minmax_commit(unsigned long long mn, unsigned long long mx,
                                                   unsigned long long* gmin, unsigned long long* gmax) {
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long a = __shfl_xor_sync(kFullMask, mn, o);
    unsigned long long c = __shfl_xor_sync(kFullMask, mx, o);
    mn = a < mn ? a : mn;
    mx = c > mx ? c : mx;
  }
  if (lane_id() == 0) {
    atomicMin(gmin, mn);
    atomicMax(gmax, mx);
  }
}

__global__ void k_bounds(PendingDev pd, RunningDev rn, int64_t now, uint64_t max_age, Bounds* b) {
  const uint32_t total = pd.n + rn.n;
  unsigned long long age_mn = ~0ull, age_mx = 0, qos_mn = ~0ull, qos_mx = 0, part_mn = ~0ull, part_mx = 0,
                     nod_mn = ~0ull, nod_mx = 0, mem_mn = ~0ull, mem_mx = 0, cpu_mn = ~0ull, cpu_mx = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    unsigned long long qos, part, nodes, mem, cpu;
    if (i < pd.n) {
      unsigned long long age = (unsigned long long)(now - pd.submit_time[i]);
      if (age > max_age) age = max_age;
      age_mn = age < age_mn ? age : age_mn;
      age_mx = age > age_mx ? age : age_mx;
      qos = pd.qos_priority[i];
      part = pd.partition_priority[i];
      nodes = pd.node_num[i];
      mem = pd.req_total[i].mem;
      cpu = (unsigned long long)pd.req_total[i].cpu_raw;
    } else {
      uint32_t k = i - pd.n;
      qos = rn.qos_priority[k];
      part = rn.partition_priority[k];
      nodes = rn.node_num[k];
      mem = rn.view_mem[k];
      cpu = (unsigned long long)rn.view_cpu_raw[k];
    }
    qos_mn = qos < qos_mn ? qos : qos_mn; qos_mx = qos > qos_mx ? qos : qos_mx;
    part_mn = part < part_mn ? part : part_mn; part_mx = part > part_mx ? part : part_mx;
    nod_mn = nodes < nod_mn ? nodes : nod_mn; nod_mx = nodes > nod_mx ? nodes : nod_mx;
    mem_mn = mem < mem_mn ? mem : mem_mn; mem_mx = mem > mem_mx ? mem : mem_mx;
    cpu_mn = cpu < cpu_mn ? cpu : cpu_mn; cpu_mx = cpu > cpu_mx ? cpu : cpu_mx;
  }
  warp_minmax_commit(age_mn, age_mx, &b->age_min, &b->age_max);
  warp_minmax_commit(qos_mn, qos_mx, &b->qos_min, &b->qos_max);
  warp_minmax_commit(part_mn, part_mx, &b->part_min, &b->part_max);
  warp_minmax_commit(nod_mn, nod_mx, &b->nodes_min, &b->nodes_max);
  warp_minmax_commit(mem_mn, mem_mx, &b->mem_min, &b->mem_max);
  warp_minmax_commit(cpu_mn, cpu_mx, &b->cpus_min, &b->cpus_max);
}

__device__ __forceinline__ double cpu_raw_to_double(unsigned long long raw) {
  return __ddiv_rn(__ll2double_rn((long long)raw), 256.0);
}
// cpus_alloc_min starts at numeric_limits<double>::max(), cpus_alloc_max at 0
// (JobScheduler.cpp:6575-6576); with no job at all the raw sentinel is kept.
__device__ __forceinline__ double cpus_min_double(const Bounds& b) {
  return b.cpus_min == ~0ull ? 1.7976931348623157e308 : cpu_raw_to_double(b.cpus_min);
}

// per-account service value (JobScheduler.cpp:6635-6671); one thread per
// account walks that account's running jobs in input order (deviation D6).
__global__ void k_service(RunningDev rn, int64_t now, Bounds* b, double* acc_service) {
  uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  bool present = a < rn.n_accounts && rn.acc_present[a];
  double sum = 0.0;
  if (present) {
    const double cmin = cpus_min_double(*b), cmax = cpu_raw_to_double(b->cpus_max);
    for (uint32_t k = rn.acc_off[a]; k < rn.acc_off[a + 1]; ++k) {
      uint32_t j = rn.acc_job[k];
      double sv = 0.0;
      if (cmax > cmin)
        sv = __dadd_rn(sv, __ddiv_rn(__dmul_rn(1.0, __dsub_rn(cpu_raw_to_double((unsigned long long)rn.view_cpu_raw[j]), cmin)),
                                     __dsub_rn(cmax, cmin)));
      else
        sv = __dadd_rn(sv, 1.0);
      if (b->nodes_max > b->nodes_min)
        sv = __dadd_rn(sv, __ddiv_rn(__dmul_rn(1.0, __uint2double_rn(rn.node_num[j] - (uint32_t)b->nodes_min)),
                                     __uint2double_rn((uint32_t)b->nodes_max - (uint32_t)b->nodes_min)));
      else
        sv = __dadd_rn(sv, 1.0);
      if (b->mem_max > b->mem_min)
        sv = __dadd_rn(sv, __ddiv_rn(__dmul_rn(1.0, __ull2double_rn(rn.view_mem[j] - b->mem_min)),
                                     __ull2double_rn(b->mem_max - b->mem_min)));
      else
        sv = __dadd_rn(sv, 1.0);
      unsigned long long run_time = (unsigned long long)(now - rn.start_time[j]);
      sum = __dadd_rn(sum, __dmul_rn(sv, __ull2double_rn(run_time)));
    }
    acc_service[a] = sum;
    unsigned long long bits = (unsigned long long)__double_as_longlong(sum);  // sum >= 0
    atomicMin(&b->svc_min_bits, bits);
    atomicMax(&b->svc_max_bits, bits);
  }
}

// ------------------------------------------------------------------------
// K-prio part 2: priority value and sort key (JobScheduler.cpp:6674-6739)
// ------------------------------------------------------------------------
struct PrioCfg {
  uint32_t type, favor_small, w_age, w_fs, w_size, w_part, w_qos;
  uint64_t max_age;
};

__global__ void k_priority(PendingDev pd, PrioCfg cfg, int64_t now, const Bounds* bp,
                           const double* acc_service, double* prio_out, uint64_t* key_out,
                           uint32_t* idx_out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pd.n) return;
  double p = pd.mandated_priority ? pd.mandated_priority[i] : 0.0;
  if (cfg.type != 0 && p == 0.0) {
    const Bounds b = *bp;
    unsigned long long age = (unsigned long long)(now - pd.submit_time[i]);
    if (age > cfg.max_age) age = cfg.max_age;
    double f_age = 0, f_qos = 0, f_part = 0, f_size = 0, f_fs = 0;
    if (b.age_max > b.age_min)
      f_age = __ddiv_rn(__dmul_rn(1.0, __ull2double_rn(age - b.age_min)), __ull2double_rn(b.age_max - b.age_min));
    if (b.qos_max > b.qos_min)
      f_qos = __ddiv_rn(__dmul_rn(1.0, __uint2double_rn(pd.qos_priority[i] - (uint32_t)b.qos_min)),
                        __uint2double_rn((uint32_t)b.qos_max - (uint32_t)b.qos_min));
    if (b.part_max > b.part_min)
      f_part = __ddiv_rn(__dmul_rn(1.0, __uint2double_rn(pd.partition_priority[i] - (uint32_t)b.part_min)),
                         __uint2double_rn((uint32_t)b.part_max - (uint32_t)b.part_min));
    const double cmin = cpus_min_double(b), cmax = cpu_raw_to_double(b.cpus_max);
    if (cmax > cmin)
      f_size = __dadd_rn(f_size, __ddiv_rn(__dmul_rn(1.0, __dsub_rn(cpu_raw_to_double((unsigned long long)pd.req_total[i].cpu_raw), cmin)),
                                           __dsub_rn(cmax, cmin)));
    if (b.nodes_max > b.nodes_min)
      f_size = __dadd_rn(f_size, __ddiv_rn(__dmul_rn(1.0, __uint2double_rn(pd.node_num[i] - (uint32_t)b.nodes_min)),
                                           __uint2double_rn((uint32_t)b.nodes_max - (uint32_t)b.nodes_min)));
    if (b.mem_max > b.mem_min)
      f_size = __dadd_rn(f_size, __ddiv_rn(__dmul_rn(1.0, __ull2double_rn(pd.req_total[i].mem - b.mem_min)),
                                           __ull2double_rn(b.mem_max - b.mem_min)));
    if (cfg.favor_small)
      f_size = __dsub_rn(1.0, __ddiv_rn(f_size, 3.0));
    else
      f_size = __ddiv_rn(f_size, 3.0);
    double smin = __longlong_as_double((long long)b.svc_min_bits);
    double smax = __longlong_as_double((long long)b.svc_max_bits);
    if (smax > smin)
      f_fs = __dsub_rn(1.0, __ddiv_rn(__dsub_rn(acc_service[pd.account[i]], smin), __dsub_rn(smax, smin)));
    p = __dmul_rn(__uint2double_rn(cfg.w_age), f_age);
    p = __dadd_rn(p, __dmul_rn(__uint2double_rn(cfg.w_part), f_part));
    p = __dadd_rn(p, __dmul_rn(__uint2double_rn(cfg.w_size), f_size));
    p = __dadd_rn(p, __dmul_rn(__uint2double_rn(cfg.w_fs), f_fs));
    p = __dadd_rn(p, __dmul_rn(__uint2double_rn(cfg.w_qos), f_qos));
  }
  prio_out[i] = p;
  // ascending radix order of the key == descending priority; ties keep input
  // order because the LSD sort is stable (deviation D2). BasicPriority: key 0.
  uint64_t bits = (uint64_t)__double_as_longlong(p);
  uint64_t orderable = (bits >> 63) ? ~bits : (bits | 0x8000000000000000ull);
  key_out[i] = cfg.type == 0 ? 0ull : ~orderable;
  idx_out[i] = i;
}

// ------------------------------------------------------------------------
// stable LSD radix sort, 8-bit digits, (u64 key, u32 value)
// one warp per block walks its tile row by row, so equal digits keep order.
// ------------------------------------------------------------------------
constexpr int kSortTile = 2048;

__device__ __forceinline__ unsigned match_any_u32(unsigned v) {
#ifdef CRANE_EMU
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) {
    unsigned o = __shfl_sync(kFullMask, v, l);
    if (o == v) m |= 1u << l;
  }
  return m;
#else
  return __match_any_sync(kFullMask, v);
#endif
}

__global__ void k_sort_hist(const uint64_t* keys, uint32_t n, int shift, uint32_t* hist, uint32_t nblocks) {
  __shared__ uint32_t s_cnt[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_cnt[i] = 0;
  __syncthreads();
  uint32_t base = blockIdx.x * kSortTile;
  for (uint32_t off = threadIdx.x; off < (uint32_t)kSortTile; off += blockDim.x) {
    uint32_t i = base + off;
    if (i < n) atomicAdd(&s_cnt[(keys[i] >> shift) & 0xff], 1u);
  }
  __syncthreads();
  for (int d = threadIdx.x; d < 256; d += blockDim.x) hist[(size_t)d * nblocks + blockIdx.x] = s_cnt[d];
}

// exclusive scan of hist[256*nblocks] (digit-major) by one block
__global__ void k_sort_scan(uint32_t* hist, uint32_t total) {
  __shared__ uint32_t s_part[1024];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < total; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < total ? hist[i] : 0;
    // block inclusive scan (Hillis-Steele in shared memory)
    s_part[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t o = 1; o < blockDim.x; o <<= 1) {
      uint32_t add = threadIdx.x >= o ? s_part[threadIdx.x - o] : 0;
      __syncthreads();
      s_part[threadIdx.x] += add;
      __syncthreads();
    }
    uint32_t incl = s_part[threadIdx.x];
    uint32_t carry = s_carry;
    if (i < total) hist[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = carry + incl;
    __syncthreads();
  }
}

__global__ void k_sort_scatter(const uint64_t* keys_in, const uint32_t* vals_in, uint64_t* keys_out,
                               uint32_t* vals_out, uint32_t n, int shift, const uint32_t* hist,
                               uint32_t nblocks) {
  // blockDim.x == 32
  __shared__ uint32_t s_off[256];
  for (int d = threadIdx.x; d < 256; d += 32) s_off[d] = hist[(size_t)d * nblocks + blockIdx.x];
  __syncwarp();
  uint32_t base = blockIdx.x * kSortTile;
  const int lane = threadIdx.x;
  for (int row = 0; row < kSortTile / 32; ++row) {
    uint32_t i = base + row * 32 + lane;
    bool valid = i < n;
    uint64_t k = valid ? keys_in[i] : 0;
    unsigned d = valid ? (unsigned)((k >> shift) & 0xff) : 0x100u;  // invalid lanes share a fake digit
    unsigned peers = match_any_u32(d);
    int rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t dst = 0;
    if (valid) dst = s_off[d] + rank;
    __syncwarp();
    if (valid && rank == 0) s_off[d] += __popc(peers);
    __syncwarp();
    if (valid) {
      keys_out[dst] = k;
      vals_out[dst] = vals_in[i];
    }
  }
}

// ------------------------------------------------------------------------
// queue build: after the priority sort, order[r] is the job at rank r.
// Ranks >= limit get "Priority" (JobScheduler.cpp:6545-6550 / JS.h:191-193);
// unknown partitions get "Partition Not Found" (JobScheduler.cpp:5783-5786).
// The remaining ranks are stably re-sorted by partition id (key2).
// ------------------------------------------------------------------------
__global__ void k_queue_keys(PendingDev pd, const uint32_t* order, const double* prio, uint32_t limit,
                             uint32_t n_parts, uint64_t* key2, PlaceDev out, uint32_t* part_count) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= pd.n) return;
  uint32_t j = order[r];
  out.priority[j] = prio[j];
  out.start_time[j] = 0;
  out.end_time[j] = 0;
  out.n_alloc[j] = 0;
  uint32_t p = pd.partition[j];
  uint8_t reason = CRANE_REASON_NONE;
  uint64_t k = 0;
  if (r >= limit) {
    reason = CRANE_REASON_PRIORITY;
    k = (uint64_t)n_parts + 1;
  } else if (p >= n_parts) {
    reason = CRANE_REASON_PART_NOT_FOUND;
    k = (uint64_t)n_parts;
  } else {
    k = p;
    atomicAdd(&part_count[p], 1u);
  }
  out.reason[j] = reason;
  key2[r] = k;
}

__global__ void k_part_offsets(const uint32_t* part_count, uint32_t n_parts, uint32_t* part_job_off) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    uint32_t acc = 0;
    for (uint32_t p = 0; p < n_parts; ++p) {
      part_job_off[p] = acc;
      acc += part_count[p];
    }
    part_job_off[n_parts] = acc;
  }
}

// JobQ records in final queue order (coalesced 96-byte records for the commit
// kernel); min_res_view of JobScheduler.cpp:5190-5192.
__global__ void k_build_jobq(PendingDev pd, const uint32_t* queue, const uint32_t* n_queued_ptr, JobQ* jobq) {
  uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= *n_queued_ptr) return;
  uint32_t j = queue[r];
  JobQ q;
  uint32_t t = pd.ntasks_per_node_min[j];
  view_node_plus_tasks(q.req, pd.req_node[j], pd.req_task[j], t);
  q.time_limit = pd.time_limit[j];
  q.job = j;
  q.node_num = pd.node_num[j];
  q.alloc_off = pd.alloc_off[j];
  q.ntasks_per_node = t;
  bool gres = false;
  for (int g = 0; g < CRANE_GRES_NAMES; ++g) gres |= q.req.gres_total[g] != 0;
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) gres |= q.req.gres_spec[e] != 0;
  q.flags = (pd.exclusive[j] ? 1u : 0u) | (gres ? 2u : 0u);
  for (int i = 0; i < 3; ++i) q.pad[i] = 0;
  jobq[r] = q;
}

// ------------------------------------------------------------------------
// K-init: per node slot, NodeState + timeline + initial cost
// (JobScheduler.cpp:5715-5753, JobScheduler.h:295-332, 492-505)
// ------------------------------------------------------------------------
__global__ void k_node_init(ClusterDev cl, RunningDev rn, TimelineDev tl, int64_t now, uint32_t max_jobs) {
  uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= cl.n_slots) return;
  const Row total = cl.slot_total[g];
  Row avail = total;
  double cost = 0.0;
  uint32_t lo = rn.n ? rn.slot_off[g] : 0, hi = rn.n ? rn.slot_off[g + 1] : 0;
  for (uint32_t k = lo; k < hi; ++k) {  // allocated_res in input order
    int64_t end = rn.slot_end[k];
    if (end < now + 1) end = now + 1;   // JobScheduler.cpp:5547-5548
    const Row res = rn.slot_res[k];
    row_sub(avail, res);
    cost = __dadd_rn(cost, cost_delta(end - now, res.cpu_raw, total.cpu_raw));
  }
  tl.avail0[g] = avail;
  tl.cost0[g] = cost;
  int64_t* T = tl.time + (size_t)g * tl.cap;
  Row* S = tl.seg + (size_t)g * tl.cap;
  Row* P = tl.pm + (size_t)g * tl.cap;
  uint32_t n = 1;
  T[0] = now;
  S[0] = avail;
  bool overflow = false;
  // value of the segment at time t = avail + sum of releases with end <= t
  for (uint32_t k = lo; k < hi && !overflow; ++k) {
    int64_t end = rn.slot_end[k];
    if (end < now + 1) end = now + 1;
    const Row res = rn.slot_res[k];
    uint32_t idx = 1;
    while (idx < n && T[idx] < end) ++idx;
    if (idx == n || T[idx] != end) {
      if (n + 2 > tl.cap) { overflow = true; break; }  // + sentinel would not fit
      for (uint32_t m = n; m > idx; --m) { T[m] = T[m - 1]; S[m] = S[m - 1]; }
      T[idx] = end;
      S[idx] = S[idx - 1];
      ++n;
    }
    for (uint32_t m = idx; m < n; ++m) row_add(S[m], res);
  }
  T[n] = kInf;  // time_avail_res_map[end].SetToZero(), JobScheduler.h:331
  row_zero(S[n]);
  ++n;
  Row acc;
  pm_identity(acc);
  for (uint32_t m = 0; m < n; ++m) {
    pm_absorb(acc, S[m]);
    P[m] = acc;
  }
  tl.n[g] = n;
  tl.skip[g] = (overflow || n >= max_jobs) ? 1 : 0;  // JobScheduler.cpp:5230
}

// ------------------------------------------------------------------------
// K-feas: jobs x nodes capability bitmap.
// bit(r, q) = node q of job r's partition passes the node-list filters
// (JobScheduler.cpp:5238-5256) and get_max_tasks(res_total) > 0 (:5258).
// One warp per queue rank; lane l evaluates node 32*w + l; __ballot_sync
// packs the word.
// ------------------------------------------------------------------------
__global__ void k_feas_bitmap(ClusterDev cl, PendingDev pd, const JobQ* jobq, const uint32_t* n_queued_ptr,
                              uint32_t words_per_row, uint32_t* bitmap) {
  const int lane = lane_id();
  const uint32_t n_queued = *n_queued_ptr;
  uint32_t warps_per_block = blockDim.x >> 5;
  for (uint32_t r = blockIdx.x * warps_per_block + warp_id(); r < n_queued; r += gridDim.x * warps_per_block) {
    JobQ jq = jobq[r];
    uint32_t p = pd.partition[jq.job];
    uint32_t base = cl.part_base[p], mp = cl.part_base[p + 1] - base;
    uint32_t il = 0, ih = 0, el = 0, eh = 0;
    if (pd.incl_off) { il = pd.incl_off[jq.job]; ih = pd.incl_off[jq.job + 1]; }
    if (pd.excl_off) { el = pd.excl_off[jq.job]; eh = pd.excl_off[jq.job + 1]; }
    for (uint32_t w = 0; w < words_per_row; ++w) {
      uint32_t q = w * 32 + lane;
      bool ok = false;
      if (q < mp) {
        uint32_t node = cl.slot_node[base + q];
        ok = true;
        if (ih > il) {  // included_nodes non-empty: node must be listed
          ok = false;
          for (uint32_t k = il; k < ih; ++k) ok |= pd.incl_nodes[k] == node;
        }
        for (uint32_t k = el; k < eh && ok; ++k) ok = pd.excl_nodes[k] != node;
        if (ok) ok = feasible<false>(jq.req, cl.slot_total[base + q], c_dict, nullptr);
      }
      unsigned word = __ballot_sync(kFullMask, ok);
      if (lane == 0) bitmap[(size_t)r * words_per_row + w] = word;
    }
  }
}

// ------------------------------------------------------------------------
// K-commit: the sequential job loop (JobScheduler.cpp:5777-5867), one
// persistent CTA per partition, node-parallel inside each job.
// ------------------------------------------------------------------------
struct CommitArgs {
  ClusterDev cl;
  TimelineDev tl;
  const JobQ* jobq;
  const uint32_t* part_job_off;  // [n_parts+1] ranges of jobq
  const uint32_t* bitmap;
  uint32_t words_per_row;
  PlaceDev out;
  Row* scratch_alloc;            // [n_slots] per-partition scratch (slot-range indexed)
  int64_t now;
  int64_t max_window;
  uint32_t max_jobs;
};

// shared-memory carve-up for a partition of mp nodes
struct CommitSmem {
  double* cost;        // [mp]  NodeRater::cost
  long long* cpu0;     // [mp]  cpu of the first timeline segment (pre-filter)
  unsigned long long* gcnt;  // [mp] packed gres slot counts of the first segment
  uint16_t* order;     // [mp]  position -> local node, ascending (cost, node)
  uint16_t* pos;       // [mp]  local node -> position
  uint16_t* cand;      // [mp]  scratch: candidates / selection
  uint8_t* skip;       // [mp]
  uint32_t* bits;      // [words] capability bitmap row of the current job
};
__host__ __device__ inline size_t commit_smem_bytes(uint32_t mp, uint32_t words) {
  size_t b = 0;
  b += (size_t)mp * 8 * 3;
  b += (size_t)mp * 2 * 3;
  b += ((size_t)mp + 7) / 8 * 8;
  b += (size_t)words * 4 + 16;
  return b + 64;
}

// number of timeline entries with time < bound (kStrict) or <= bound
template <bool kStrict>
__device__ __forceinline__ uint32_t tl_count_before(const int64_t* T, uint32_t n, int64_t bound) {
  const int lane = lane_id();
  uint32_t cnt = 0;
  for (uint32_t base = 0; base < n; base += 32) {
    uint32_t i = base + lane;
    int64_t t = i < n ? T[i] : kInf;
    bool in = i < n && (kStrict ? t < bound : t <= bound);
    unsigned m = __ballot_sync(kFullMask, in);
    cnt += __popc(m);
    if (m != kFullMask) break;
  }
  return cnt;
}

// the exact per-node test of JobScheduler.cpp:5285-5334 for one candidate:
// window minimum over the segments that start before now+time_limit, then
// get_max_tasks(min) > 0. All lanes of the warp return the same verdict.
__device__ __forceinline__ bool window_check(const TimelineDev& tl, const ClusterDev& cl, uint32_t g,
                                             const JobQ& jq, int64_t w_end, Row* win_row) {
  const int64_t* T = tl.time + (size_t)g * tl.cap;
  uint32_t n = tl.n[g];
  uint32_t cnt = tl_count_before<true>(T, n, w_end);
  const Row pm = tl.pm[(size_t)g * tl.cap + (cnt - 1)];
  if (jq.flags & 1u) {  // exclusive: every segment in the window must hold res_total
    const Row total = cl.slot_total[g];
    *win_row = total;
    return row_le(total, pm);
  }
  Row a0 = tl.avail0[g];
  // stale pre-filter on res_avail (JobScheduler.cpp:5310) is implied: the
  // window row is <= res_avail in every compared field.
  window_row(*win_row, a0, pm);
  return feasible<false>(jq.req, *win_row, c_dict, nullptr);
}

// earliest t >= T0 such that `alloc` <= every segment overlapping
// [t, t+limit) on node g; kInf if none. One warp, lanes = segments.
// (per-node half of EarliestStartSubsetSelector, JobScheduler.h:806-849)
__device__ __forceinline__ int64_t earliest_on_node(const TimelineDev& tl, uint32_t g, const Row& alloc,
                                                    int64_t T0, int64_t limit) {
  const int lane = lane_id();
  const int64_t* T = tl.time + (size_t)g * tl.cap;
  const Row* S = tl.seg + (size_t)g * tl.cap;
  const uint32_t n = tl.n[g];
  int64_t carry = -1;  // start time of the satisfied run that reaches this chunk, or -1
  for (uint32_t base = 0; base < n; base += 32) {
    uint32_t i = base + lane;
    bool valid = i < n;
    int64_t t = valid ? T[i] : kInf;
    int64_t tend = (i + 1 < n) ? T[i + 1] : kInf;
    bool sat = false;
    if (valid && tend > T0) {
      const Row s = S[i];
      sat = row_le(alloc, s);
    }
    bool breaker = !sat;  // unsatisfied, or entirely before T0, or past the end
    unsigned bm = __ballot_sync(kFullMask, breaker);
    bool prev_break = lane == 0 ? (carry < 0) : ((bm >> (lane - 1)) & 1u);
    int64_t startv = (!breaker && prev_break) ? (t > T0 ? t : T0) : -1;
    if (lane == 0 && !breaker && !prev_break) startv = carry;
    // inclusive max-scan: run starts are non-decreasing along the timeline
    for (int o = 1; o < 32; o <<= 1) {
      int64_t up = shfl_i64(startv, lane - o >= 0 ? lane - o : lane);
      if (lane >= o && up > startv) startv = up;
    }
    bool ok = !breaker && (tend == kInf || tend - startv >= limit);
    unsigned okm = __ballot_sync(kFullMask, ok);
    if (okm) {
      int first = __ffs((int)okm) - 1;
      return shfl_i64(startv, first);
    }
    bool last_break = (bm >> 31) & 1u;
    int64_t last_start = shfl_i64(startv, 31);
    carry = last_break ? -1 : last_start;
  }
  return kInf;
}

// NodeState::UpdateResourceInNode (JobScheduler.h:334-453, allocation
// direction) on the array timeline of slot g, by one warp, followed by the
// prefix-min refresh. Returns the new entry count.
__device__ __forceinline__ uint32_t timeline_update(const TimelineDev& tl, uint32_t g, int64_t start,
                                                    int64_t end, const Row& alloc) {
  const int lane = lane_id();
  int64_t* T = tl.time + (size_t)g * tl.cap;
  Row* S = tl.seg + (size_t)g * tl.cap;
  Row* P = tl.pm + (size_t)g * tl.cap;
  const uint32_t n = tl.n[g];
  const uint32_t i_s = tl_count_before<false>(T, n, start) - 1;  // last key <= start
  const uint32_t i_e = tl_count_before<false>(T, n, end) - 1;    // last key <= end
  const bool ins_s = T[i_s] != start;
  const bool ins_e = T[i_e] != end;
  const Row seg_s = S[i_s];  // values before any modification
  const Row seg_e = S[i_e];
  const uint32_t add = (ins_s ? 1u : 0u) + (ins_e ? 1u : 0u);
  __syncwarp();
  // move old entries (i_s, n) upward, top chunk first; subtract inside [start,end)
  if (n > i_s + 1) {
    int64_t hi = (int64_t)n - 1;
    const int64_t lo = (int64_t)i_s + 1;
    while (hi >= lo) {
      int64_t j = hi - lane;
      bool act = j >= lo;
      int64_t t = 0;
      Row r;
      if (act) { t = T[j]; r = S[j]; }
      __syncwarp();
      if (act) {
        if (t >= start && t < end) row_sub(r, alloc);
        uint32_t nj = (uint32_t)j + (ins_s ? 1u : 0u) + (((uint32_t)j > i_e && ins_e) ? 1u : 0u);
        T[nj] = t;
        S[nj] = r;
      }
      __syncwarp();
      hi -= 32;
    }
  }
  if (lane == 0) {
    if (ins_s) {  // case #3: copy of the covering segment, minus the job
      Row r = seg_s;
      row_sub(r, alloc);
      T[i_s + 1] = start;
      S[i_s + 1] = r;
    } else {      // case #4: key == start already exists
      Row r = seg_s;
      row_sub(r, alloc);
      S[i_s] = r;
    }
    if (ins_e) {  // new breakpoint at `end` keeps the un-subtracted value
      uint32_t ne = i_e + (ins_s ? 1u : 0u) + 1u;
      T[ne] = end;
      S[ne] = seg_e;
    }
  }
  __syncwarp();
  const uint32_t nn = n + add;
  // refresh prefix minima from the first changed entry
  uint32_t m0 = ins_s ? i_s + 1 : i_s;
  Row carry;
  if (m0 > 0) carry = P[m0 - 1]; else pm_identity(carry);
  for (uint32_t base = m0; base < nn; base += 32) {
    uint32_t i = base + lane;
    Row v;
    pm_identity(v);
    if (i < nn) { const Row s = S[i]; pm_absorb(v, s); }
    for (int o = 1; o < 32; o <<= 1) {
      Row up;
      row_shfl(up, v, lane - o >= 0 ? lane - o : lane);
      if (lane >= o) pm_combine(v, up);
    }
    pm_combine(v, carry);
    if (i < nn) P[i] = v;
    Row last;
    row_shfl(last, v, 31);
    carry = last;
  }
  __syncwarp();
  if (lane == 0) tl.n[g] = nn;
  return nn;
}

// move local node u (whose cost just grew to new_cost) toward the back of the
// (cost, node) order; block-wide. NodeSelector::UpdateCost's erase+emplace in
// std::set<pair<double,NodeState*>> (JobScheduler.h:520-532), tie = node index.
__device__ __forceinline__ void reorder_node(CommitSmem& sm, uint32_t mp, uint32_t u, double new_cost) {
  __shared__ uint32_t s_stop;
  const uint32_t p = sm.pos[u];
  if (threadIdx.x == 0) s_stop = 0xffffffffu;
  __syncthreads();
  for (uint32_t base = p + 1; base < mp; base += blockDim.x) {
    uint32_t i = base + threadIdx.x;
    uint32_t o = 0;
    bool lt = false;
    if (i < mp) {
      o = sm.order[i];
      double c = sm.cost[o];
      lt = (c < new_cost) || (c == new_cost && o < u);
      if (!lt) atomicMin(&s_stop, i);
    }
    __syncthreads();
    if (i < mp && lt) {  // sorted => the lt positions form a prefix of (p, mp)
      sm.order[i - 1] = (uint16_t)o;
      sm.pos[o] = (uint16_t)(i - 1);
    }
    __syncthreads();
    if (s_stop != 0xffffffffu) break;
  }
  if (threadIdx.x == 0) {
    uint32_t stop = s_stop == 0xffffffffu ? mp : s_stop;  // first position not less than u
    sm.order[stop - 1] = (uint16_t)u;
    sm.pos[u] = (uint16_t)(stop - 1);
    sm.cost[u] = new_cost;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512, 1) k_commit(CommitArgs a) {
  CRANE_DYN_SMEM(unsigned char, smem_raw);
  const uint32_t part = blockIdx.x;
  const uint32_t base = a.cl.part_base[part];
  const uint32_t mp = a.cl.part_base[part + 1] - base;
  const uint32_t words = a.words_per_row;
  const int lane = lane_id();
  const int wid = warp_id();
  const int nwarps = blockDim.x >> 5;

  CommitSmem sm;
  {
    unsigned char* ptr = smem_raw;
    // widest element type first so every array is naturally aligned
    sm.cost = reinterpret_cast<double*>(ptr); ptr += (size_t)mp * 8;
    sm.cpu0 = reinterpret_cast<long long*>(ptr); ptr += (size_t)mp * 8;
    sm.gcnt = reinterpret_cast<unsigned long long*>(ptr); ptr += (size_t)mp * 8;
    sm.bits = reinterpret_cast<uint32_t*>(ptr); ptr += (size_t)words * 4;
    sm.order = reinterpret_cast<uint16_t*>(ptr); ptr += (size_t)mp * 2;
    sm.pos = reinterpret_cast<uint16_t*>(ptr); ptr += (size_t)mp * 2;
    sm.cand = reinterpret_cast<uint16_t*>(ptr); ptr += (size_t)mp * 2;
    sm.skip = ptr;
  }
  __shared__ JobQ s_job;
  __shared__ uint32_t s_warp_cnt[32];
  __shared__ uint32_t s_nsel, s_flag[32];
  __shared__ int64_t s_tmax;
  __shared__ int64_t s_tnode[32];
  __shared__ uint32_t s_resource_label;

  // ---- load node state; initial order = ascending (cost, node) -----------
  for (uint32_t q = threadIdx.x; q < mp; q += blockDim.x) {
    const uint32_t g = base + q;
    sm.cost[q] = a.tl.cost0[g];
    const Row s0 = a.tl.seg[(size_t)g * a.tl.cap];
    sm.cpu0[q] = s0.cpu_raw;
    sm.gcnt[q] = pack_gres_counts(s0);
    sm.skip[q] = a.tl.skip[g];
  }
  __syncthreads();
  // rank sort: position = number of nodes with a smaller (cost, node) key
  for (uint32_t q = threadIdx.x; q < mp; q += blockDim.x) {
    const double c = sm.cost[q];
    uint32_t rank = 0;
    for (uint32_t o = 0; o < mp; ++o) {
      double co = sm.cost[o];
      rank += (co < c || (co == c && o < q)) ? 1u : 0u;
    }
    sm.order[rank] = (uint16_t)q;
    sm.pos[q] = (uint16_t)rank;
  }
  __syncthreads();

  const uint32_t r_begin = a.part_job_off[part], r_end = a.part_job_off[part + 1];
  for (uint32_t r = r_begin; r < r_end; ++r) {
    // ---- job record + capability row ------------------------------------
    if (threadIdx.x < sizeof(JobQ) / 4)
      reinterpret_cast<uint32_t*>(&s_job)[threadIdx.x] = reinterpret_cast<const uint32_t*>(&a.jobq[r])[threadIdx.x];
    for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) sm.bits[w] = a.bitmap[(size_t)r * words + w];
    if (threadIdx.x == 0) { s_nsel = 0; s_resource_label = 0; }
    __syncthreads();
    const JobQ jq = s_job;
    const uint32_t K = jq.node_num;
    const bool exclusive = jq.flags & 1u;
    const int64_t w_end = a.now + jq.time_limit;
    bool start_now = false;

    // ---- phase 1: nodes that can run the job now, in cost order ----------
    // (JobScheduler.cpp:5224-5336); sm.cand[0..K) collects the selection
    if (K <= mp) {
      for (uint32_t cbase = 0; cbase < mp && !start_now; cbase += blockDim.x) {
        const uint32_t i = cbase + threadIdx.x;
        bool cand = false;
        uint32_t q = 0;
        if (i < mp) {
          q = sm.order[i];
          cand = ((sm.bits[q >> 5] >> (q & 31)) & 1u) && !sm.skip[q];
          if (cand && !exclusive)
            cand = sm.cpu0[q] >= jq.req.cpu_raw && (!(jq.flags & 2u) || gres_counts_ok(jq.req, sm.gcnt[q]));
        }
        // ordered compaction of the candidates of this chunk
        unsigned bm = __ballot_sync(kFullMask, cand);
        if (lane == 0) s_warp_cnt[wid] = __popc(bm);
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < nwarps; ++w) {
          uint32_t c = s_warp_cnt[w];
          if (w < wid) before += c;
          total += c;
        }
        // candidates are staged behind the (< K) nodes already selected
        const uint32_t nsel0 = s_nsel;
        if (cand) sm.cand[nsel0 + before + __popc(bm & ((1u << lane) - 1u))] = (uint16_t)q;
        __syncthreads();
        // exact window test, one warp per candidate, batches in order
        for (uint32_t b0 = 0; b0 < total && !start_now; b0 += nwarps) {
          const uint32_t ci = b0 + wid;
          bool ok = false;
          if (ci < total) {
            Row wr;
            ok = window_check(a.tl, a.cl, base + sm.cand[nsel0 + ci], jq, w_end, &wr);
          }
          if (lane == 0) s_flag[wid] = ok ? 1u : 0u;
          __syncthreads();
          if (threadIdx.x == 0) {
            uint32_t ns = s_nsel;
            for (int w = 0; w < nwarps && ns < K; ++w) {
              if (b0 + w < total && s_flag[w]) {
                sm.cand[ns] = sm.cand[nsel0 + b0 + w];  // ns <= nsel0 + b0 + w: in-place compaction
                ++ns;
              }
            }
            s_nsel = ns;
          }
          __syncthreads();
          if (s_nsel >= K) start_now = true;
        }
        // keep the passing candidates compacted at the front for the next chunk:
        // (already done in place: sm.cand[0..s_nsel) holds the selection)
        __syncthreads();
      }
    }

    int64_t start_time = 0;
    bool placed = false;
    if (start_now) {
      start_time = a.now;
      placed = true;
      // allocation against the window minimum (JobScheduler.cpp:5338-5362)
      for (uint32_t k = wid; k < K; k += nwarps) {
        const uint32_t g = base + sm.cand[k];
        Row wr, alloc;
        window_check(a.tl, a.cl, g, jq, w_end, &wr);
        if (exclusive) alloc = wr; else feasible<true>(jq.req, wr, c_dict, &alloc);
        if (lane == 0) a.scratch_alloc[base + k] = alloc;
      }
    } else {
      // ---- phase 3: first K capable nodes, then backfill -----------------
      // (JobScheduler.cpp:5269-5278, 5371-5404, 5407-5412)
      __syncthreads();
      if (threadIdx.x == 0) s_nsel = 0;
      __syncthreads();
      for (uint32_t cbase = 0; cbase < mp; cbase += blockDim.x) {
        const uint32_t i = cbase + threadIdx.x;
        bool cap = false;
        uint32_t q = 0;
        if (i < mp) {
          q = sm.order[i];
          cap = ((sm.bits[q >> 5] >> (q & 31)) & 1u) && !sm.skip[q];
        }
        unsigned bm = __ballot_sync(kFullMask, cap);
        if (lane == 0) s_warp_cnt[wid] = __popc(bm);
        __syncthreads();
        uint32_t before = s_nsel, total = 0;
        for (int w = 0; w < nwarps; ++w) {
          uint32_t c = s_warp_cnt[w];
          if (w < wid) before += c;
          total += c;
        }
        uint32_t slot = before + __popc(bm & ((1u << lane) - 1u));
        if (cap && slot < K) sm.cand[slot] = (uint16_t)q;
        __syncthreads();
        if (threadIdx.x == 0) s_nsel = s_nsel + total;
        __syncthreads();
        if (s_nsel >= K) break;
      }
      if (K <= mp && s_nsel >= K) {
        // allocation against res_total (JobScheduler.cpp:5381-5403)
        for (uint32_t k = wid; k < K; k += nwarps) {
          const uint32_t g = base + sm.cand[k];
          const Row total = a.cl.slot_total[g];
          Row alloc;
          if (exclusive) alloc = total; else feasible<true>(jq.req, total, c_dict, &alloc);
          if (lane == 0) a.scratch_alloc[base + k] = alloc;
        }
        __syncthreads();
        // earliest common start: fixed point of the per-node earliest fits
        int64_t Tcur = a.now;
        bool found = false, failed = false;
        while (!found && !failed) {
          if (threadIdx.x == 0) s_tmax = Tcur;
          __syncthreads();
          for (uint32_t k0 = 0; k0 < K; k0 += nwarps) {
            const uint32_t k = k0 + wid;
            int64_t t = Tcur;
            if (k < K) {
              const Row alloc = a.scratch_alloc[base + k];
              t = earliest_on_node(a.tl, base + sm.cand[k], alloc, Tcur, jq.time_limit);
            }
            if (lane == 0) s_tnode[wid] = t;
            __syncthreads();
            if (threadIdx.x == 0) {
              int64_t m = s_tmax;
              for (int w = 0; w < nwarps; ++w) m = s_tnode[w] > m ? s_tnode[w] : m;
              s_tmax = m;
            }
            __syncthreads();
          }
          const int64_t Tn = s_tmax;
          __syncthreads();
          if (Tn == kInf) failed = true;
          else if (Tn == Tcur) found = true;
          else Tcur = Tn;
        }
        // `current_time - now > kAlgoMaxTimeWindow` (JobScheduler.h:809)
        if (found && Tcur - a.now <= a.max_window) {
          placed = true;
          start_time = Tcur;
        }
      }
    }

    __syncthreads();
    // ---- commit: timeline update, cost, order, outputs -------------------
    if (placed) {
      const int64_t end_time = start_time + jq.time_limit;
      for (uint32_t k = wid; k < K; k += nwarps) {
        const uint32_t q = sm.cand[k];
        const uint32_t g = base + q;
        const Row alloc = a.scratch_alloc[base + k];
        uint32_t nn = timeline_update(a.tl, g, start_time, end_time, alloc);
        if (lane == 0) {
          if (nn >= a.max_jobs) sm.skip[q] = 1;
          if (start_time == a.now) {
            const Row s0 = a.tl.seg[(size_t)g * a.tl.cap];
            sm.cpu0[q] = s0.cpu_raw;
            sm.gcnt[q] = pack_gres_counts(s0);
          }
          // pending-reason label for future starts (JobScheduler.cpp:5842-5848)
          if (start_time != a.now && !row_le(alloc, a.tl.avail0[g])) atomicOr(&s_resource_label, 1u);
        }
      }
      __syncthreads();
      // outputs, node-index ascending (deviation D3)
      for (uint32_t k = threadIdx.x; k < K; k += blockDim.x) {
        const uint32_t q = sm.cand[k];
        uint32_t rank = 0;
        for (uint32_t m = 0; m < K; ++m) rank += sm.cand[m] < q ? 1u : 0u;
        const uint32_t dst = jq.alloc_off + rank;
        a.out.alloc_node[dst] = a.cl.slot_node[base + q];
        a.out.alloc_ntasks[dst] = jq.ntasks_per_node;
        a.out.alloc_res[dst] = a.scratch_alloc[base + k];
      }
      if (threadIdx.x == 0) {
        a.out.start_time[jq.job] = start_time;
        a.out.end_time[jq.job] = end_time;
        a.out.n_alloc[jq.job] = K;
        uint8_t reason = CRANE_REASON_NONE;
        if (start_time != a.now) reason = s_resource_label ? CRANE_REASON_RESOURCE : CRANE_REASON_PRIORITY;
        a.out.reason[jq.job] = reason;
      }
      // cost += (end-start) * cpu ratio, then re-key (JobScheduler.h:46-52,520-532)
      for (uint32_t k = 0; k < K; ++k) {
        const uint32_t q = sm.cand[k];
        const uint32_t g = base + q;
        const double delta = cost_delta(jq.time_limit, a.scratch_alloc[base + k].cpu_raw, a.cl.slot_total[g].cpu_raw);
        const double nc = __dadd_rn(sm.cost[q], delta);
        reorder_node(sm, mp, q, nc);
      }
    } else {
      if (threadIdx.x == 0) a.out.reason[jq.job] = CRANE_REASON_RESOURCE;  // JobScheduler.cpp:5802
    }
    __syncthreads();
  }
}

}  // namespace crane
