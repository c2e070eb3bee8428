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
//   tl_ent[g][CAP] = {int64 t; Row seg} (80 B), tl_n[g]; CAP = max_jobs_per_node + 1.
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

constexpr int kMaxClasses = 16;  // distinct res_total rows cached per partition

struct ClusterDev {
  uint32_t n_slots;           // usable nodes
  uint32_t n_parts;
  uint32_t max_part_slots;
  const uint32_t* part_base;  // [n_parts+1] slot ranges
  const uint32_t* slot_node;  // slot -> global node index
  const uint32_t* node_slot;  // global node -> slot or 0xffffffff
  const Row* slot_total;      // res_total per slot
  const uint8_t* slot_class;  // per slot: index into its partition's class rows, 0xff = none
  const Row* class_rows;      // [n_parts][kMaxClasses]
};

// one timeline breakpoint: from time t on, `seg` is available on the node
// (an entry of NodeState::time_avail_res_map, JobScheduler.h:239,285)
struct __align__(16) TlEntry {
  int64_t t;
  Row seg;
};
static_assert(sizeof(TlEntry) == 80, "TlEntry layout");

struct TimelineDev {
  uint32_t cap;               // entries per slot
  uint32_t* n;                // [n_slots]
  TlEntry* ent;               // [n_slots][cap]
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
  uint32_t flags;      // bit0 exclusive, bit1 has gres request, bits 8-15 requested gres names
  uint32_t pad0;
  uint64_t spec8;      // per-entry typed counts, one byte each (clamped to 127)
  uint8_t name_need[CRANE_GRES_NAMES];  // per name max(total, sum typed), clamped to 255
  uint64_t pad1;
};
static_assert(sizeof(JobQ) == 112, "JobQ layout");

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
// packed per-entry slot counts of a row (8 x u8) for the cheap pre-filter
__device__ __forceinline__ uint64_t pack_gres_counts(const Row& r) {
  uint64_t p = 0;
#pragma unroll
  for (uint32_t e = 0; e < CRANE_GRES_ENTRIES; ++e) p |= (uint64_t)popc32(field16(r.g, e)) << (8 * e);
  return p;
}
// count-only gres verdict against packed per-entry slot counts (all < 128):
// every typed count fits its entry (byte-wise >= without borrows) and each
// requested name has max(total, sum typed) slots over its entries (byte sum by
// multiply). Same verdict as feasible_gres<false> on a row with these counts.
__device__ __forceinline__ bool gres_counts_ok(uint64_t packed, uint64_t spec8, uint32_t names, const uint8_t* name_need) {
  const uint64_t H = 0x8080808080808080ull;
  if ((((packed | H) - spec8) & H) != H) return false;
  while (names) {
    const uint32_t g = (uint32_t)__ffs((int)names) - 1u;
    names &= names - 1u;
    const uint32_t have = (uint32_t)(((packed & c_dict.name_mask8[g]) * 0x0101010101010101ull) >> 56);
    if (have < name_need[g]) return false;
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
  q.flags = (pd.exclusive[j] ? 1u : 0u) | (view_has_gres(q.req) ? 2u : 0u);
  q.pad0 = 0;
  q.pad1 = 0;
  q.spec8 = 0;
  for (uint32_t g = 0; g < CRANE_GRES_NAMES; ++g) {
    uint32_t typed = 0;
    for (uint32_t e = c_dict.name_first[g]; e < (uint32_t)c_dict.name_first[g] + c_dict.name_count[g]; ++e) {
      const uint32_t sp = field16(q.req.gspec, e);
      typed += sp;
      q.spec8 |= (uint64_t)(sp > 127 ? 127 : sp) << (8 * e);
    }
    const uint32_t tot = field16(q.req.gtot, g);
    const uint32_t need = tot > typed ? tot : typed;
    q.name_need[g] = (uint8_t)(need > 255 ? 255 : need);
    if (need) q.flags |= 1u << (8 + g);
  }
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
  TlEntry* E = tl.ent + (size_t)g * tl.cap;
  uint32_t n = 1;
  E[0].t = now;
  E[0].seg = avail;
  bool overflow = false;
  // value of the segment at time t = avail + sum of releases with end <= t
  for (uint32_t k = lo; k < hi && !overflow; ++k) {
    int64_t end = rn.slot_end[k];
    if (end < now + 1) end = now + 1;
    const Row res = rn.slot_res[k];
    uint32_t idx = 1;
    while (idx < n && E[idx].t < end) ++idx;
    if (idx == n || E[idx].t != end) {
      if (n + 2 > tl.cap) { overflow = true; break; }  // + sentinel would not fit
      for (uint32_t m = n; m > idx; --m) E[m] = E[m - 1];
      E[idx].t = end;
      E[idx].seg = E[idx - 1].seg;
      ++n;
    }
    for (uint32_t m = idx; m < n; ++m) row_add(E[m].seg, res);
  }
  E[n].t = kInf;  // time_avail_res_map[end].SetToZero(), JobScheduler.h:331
  row_zero(E[n].seg);
  ++n;
  tl.n[g] = n;
  tl.skip[g] = (overflow || n >= max_jobs) ? 1 : 0;  // JobScheduler.cpp:5230
}

// ------------------------------------------------------------------------
// K-feas: jobs x nodes capability bitmap.
// bit(r, q) = node q of job r's partition passes the node-list filters
// (JobScheduler.cpp:5238-5256) and get_max_tasks(res_total) > 0 (:5258).
// One warp per queue rank; lane l evaluates node 32*w + l; __ballot_sync
// packs the word. Rows are `words_per_row` (a multiple of 4) words apart so
// the commit kernel can fetch a row with one 16-byte-aligned bulk copy.
// ------------------------------------------------------------------------
__global__ void k_feas_bitmap(ClusterDev cl, PendingDev pd, const JobQ* jobq, const uint32_t* n_queued_ptr,
                              uint32_t words_per_row, uint32_t* bitmap, const uint32_t* part_owner, uint32_t rank) {
  const int lane = lane_id();
  const uint32_t n_queued = *n_queued_ptr;
  uint32_t warps_per_block = blockDim.x >> 5;
  for (uint32_t r = blockIdx.x * warps_per_block + warp_id(); r < n_queued; r += gridDim.x * warps_per_block) {
    JobQ jq = jobq[r];
    uint32_t p = pd.partition[jq.job];
    if (part_owner && part_owner[p] != rank) continue;  // another GPU commits this partition
    uint32_t base = cl.part_base[p], mp = cl.part_base[p + 1] - base;
    const uint32_t row_words = (mp + 31) / 32;          // words beyond the job's own partition are never read
    uint32_t il = 0, ih = 0, el = 0, eh = 0;
    if (pd.incl_off) { il = pd.incl_off[jq.job]; ih = pd.incl_off[jq.job + 1]; }
    if (pd.excl_off) { el = pd.excl_off[jq.job]; eh = pd.excl_off[jq.job + 1]; }
    for (uint32_t w = 0; w < row_words; ++w) {
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

// One queue over several GPUs: the placement columns of jobs another rank owns
// go to zero, so the union over ranks is a sum (crane_sched_set_shard).
__global__ void k_shard_mask(PendingDev pd, PlaceDev out, const uint32_t* part_owner, uint32_t n_parts, uint32_t rank) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= pd.n) return;
  const uint32_t p = pd.partition[j];
  const uint32_t owner = p < n_parts ? part_owner[p] : 0u;
  if (owner != rank) {
    out.reason[j] = 0;
    out.start_time[j] = 0;
    out.end_time[j] = 0;
    out.n_alloc[j] = 0;
  }
}

// ------------------------------------------------------------------------
// K-commit: the sequential job loop (JobScheduler.cpp:5777-5867), one
// persistent CTA per partition (the reference's LocalScheduler).
//
// The loop is a dependency chain (job j+1 sees job j's allocation and cost
// update), so the design minimises the latency of one job, not throughput:
//  * a DRIVER warp (warp 0) runs a one-node job entirely by itself — scan,
//    exact test, allocation, timeline update, re-keying — with no block
//    barrier; the other warps are HELPERS parked on a barrier that the driver
//    only touches for multi-node jobs (one node per helper);
//  * the (cost, node) order of NodeSelector (JobScheduler.h:588) is a bucketed
//    sorted list in shared memory (<= 64 nodes per bucket), so re-keying a node
//    shifts tens of entries, not half the partition; every bucket carries an
//    upper bound of its nodes' first-segment cpu / gres counts so the scan for
//    an immediate start jumps over buckets that cannot match (availability
//    only shrinks inside a tick, so the bounds stay valid and are tightened
//    lazily);
//  * timelines live in HBM/L2 as 80-byte entries; a warp opens a node by
//    loading up to 64 entries into registers once and runs the window test
//    (ballot + REDUX.AND), the allocation, the earliest-start search
//    (ballot/clz run detection) and the update (per-lane stores) on them;
//  * job records and capability-bitmap rows arrive through a 4-deep
//    shared-memory ring filled by TMA bulk copies (cp.async.bulk + mbarrier)
//    three jobs ahead.
// ------------------------------------------------------------------------
// optional phase profiling (-DCRANE_PROFILE builds only; never in the product .so)
#ifdef CRANE_PROFILE
#define PROF_DECL long long prof_last = clock64(); unsigned long long prof_acc[16] = {0}
#define PROF(i) do { if (threadIdx.x == 0) { long long t__ = clock64(); prof_acc[i] += (unsigned long long)(t__ - prof_last); prof_last = t__; } } while (0)
#define PROF_CNT(i, v) do { if (threadIdx.x == 0) prof_acc[i] += (v); } while (0)
#define PROF_FLUSH(dst) do { if (threadIdx.x == 0 && (dst)) for (int i__ = 0; i__ < 16; ++i__) (dst)[blockIdx.x * 16 + i__] = prof_acc[i__]; } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_CNT(i, v)
#define PROF_FLUSH(dst)
#endif

struct CommitArgs {
  ClusterDev cl;
  TimelineDev tl;
  const JobQ* jobq;
  const uint32_t* part_job_off;  // [n_parts+1] ranges of jobq
  const uint32_t* bitmap;
  uint32_t words_per_row;        // multiple of 4
  PlaceDev out;
  int64_t now;
  int64_t max_window;
  uint32_t max_jobs;
  unsigned long long* prof;      // [n_parts][16] cycle counters (profiling builds)
};

constexpr int kRing = 16;            // prefetch ring depth (jobs)
// Registers are handed out per group of 4 warps, so 9 warps cost as much as 12
// (168 registers per thread, spills); 8 warps leave 255.
#ifndef CRANE_COMMIT_THREADS
#define CRANE_COMMIT_THREADS 256
#endif
constexpr int kCommitThreads = CRANE_COMMIT_THREADS;  // CTA size of k_commit: driver warp + helpers
constexpr int kBatch = kCommitThreads / 32 - 1;       // nodes of the jobs dispatched together (one helper warp each)
constexpr int kBatchJobs = kBatch < 8 ? kBatch : 8;   // jobs per batch: the resolve step lays 8 jobs x 4 lanes over one warp
constexpr int kEnt = (kBatch + 3) / 4;                // list entries per lane in the resolve step
static_assert(kBatch >= 1 && kBatch <= 12, "the resolve step handles lists of up to 12 entries and 16 picks");
constexpr int kBucket = 64;          // bucket capacity of the cost order
constexpr int kBucketFill = 32;      // entries per bucket after a (re)build

struct CommitSmem {
  uint32_t* bits_ring;         // [kRing][words]
  double* cost;                // [mp]  NodeRater::cost
  long long* cpu0;             // [mp]  cpu of the first timeline segment
  unsigned long long* gcnt;    // [mp]  packed gres slot counts of the first segment
  long long* bmax_cpu;         // [nb]  >= cpu0 of every node in the bucket
  long long* bmax_cpug;        // [nb]  >= cpu0 of every node in the bucket that still has a free gres slot
  unsigned long long* bmax_g;  // [nb]  >= gcnt (per byte) of every node in the bucket
  uint16_t* bk;                // [nb][kBucket] node ids, ascending (cost, node); buckets ascending
  uint16_t* bcnt;              // [nb]
  uint16_t* blast;             // [nb]  last (largest-key) node of the bucket, 0xffff = empty
  uint16_t* bkt;               // [mp]  bucket of a node
  uint16_t* list;              // [mp]  nodes handed to the workers / selected nodes of the job
  uint16_t* tmp;               // [mp]  scratch of (re)builds
  uint16_t* nseg;              // [mp]  timeline entry counts
  uint8_t* skip;               // [mp]
  uint8_t* cls;                // [mp]
  uint8_t* bexact;             // [nb]  bounds are the exact maxima (nothing inserted since the last tightening)
  uint8_t* pend;               // [mp]  picked by the batch in flight: out of the order until re-inserted
  uint32_t nb;
};
__host__ __device__ inline uint32_t commit_nbuckets(uint32_t mp) { return (mp + kBucketFill - 1) / kBucketFill + 1; }
__host__ __device__ inline size_t commit_smem_bytes(uint32_t mp, uint32_t words) {
  const size_t nb = commit_nbuckets(mp);
  size_t b = (size_t)kRing * words * 4;
  b += (size_t)mp * 8 * 3 + nb * 8 * 3 + nb;
  b += nb * kBucket * 2 + nb * 2 * 2;
  b += (size_t)mp * 2 * 4;
  b += (size_t)mp * 3;
  return b + 128;
}

// ---- TMA 1-D bulk copy + mbarrier (sm_90+/sm_100a) --------------------------
#ifdef CRANE_EMU
__device__ __forceinline__ void mbar_init(uint64_t*, uint32_t) {}
__device__ __forceinline__ void mbar_expect_tx(uint64_t*, uint32_t) {}
__device__ __forceinline__ void mbar_wait(uint64_t*, uint32_t) {}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t*) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void fence_mbar_init() {}
#else
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  const uint32_t addr = smem_u32(bar);
  while (!done) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
#endif

// ResourceView::GetFeasibleResourceInNode with the concrete pick, one shared
// out-of-line instance (keeps the per-job instruction footprint small)
// body of a polling loop on a shared-memory word
__device__ __forceinline__ void spin_pause() {
#ifdef CRANE_EMU
  sched_yield();  // the emulation runs more host threads than cores
#endif
}

// named barrier among `nthreads` threads (whole warps) of the CTA; id 1..15 (0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
#ifdef CRANE_EMU
  emu_named_bar((int)id, (int)nthreads, true);
#else
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}

__device__ __noinline__ bool feasible_alloc(const View& req, const Row& avail, Row& alloc) {
  return feasible<true>(req, avail, c_dict, &alloc);
}

__device__ __forceinline__ void prefetch_l1(const void* p) {
#ifndef CRANE_EMU
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

__device__ __forceinline__ uint64_t warp_and64(uint64_t v) {
  const uint32_t lo = __reduce_and_sync(kFullMask, (uint32_t)v);
  const uint32_t hi = __reduce_and_sync(kFullMask, (uint32_t)(v >> 32));
  return (uint64_t)hi << 32 | lo;
}

// ---- a node's timeline held by one warp -------------------------------------
// lane l holds entries l and l+32 (timelines of up to 64 entries). Longer
// timelines take the *_big paths below, which walk the entries in memory.
struct NodeRegs {
  int64_t t0, t1;
  Row s0, s1;
  uint32_t n;
};

__device__ __forceinline__ void node_open(const TimelineDev& tl, uint32_t g, uint32_t n, NodeRegs& nr) {
  const TlEntry* E = tl.ent + (size_t)g * tl.cap;
  const uint32_t lane = lane_id();
  nr.n = n;
  nr.t0 = kInf;
  nr.t1 = kInf;
  row_zero(nr.s0);
  row_zero(nr.s1);
  if (n <= 64) {
    if (lane < n) { const TlEntry e = E[lane]; nr.t0 = e.t; nr.s0 = e.seg; }
    if (lane + 32 < n) { const TlEntry e = E[lane + 32]; nr.t1 = e.t; nr.s1 = e.seg; }
  }
}

// The exact per-node test of JobScheduler.cpp:5285-5334 and the allocation of
// :5340-5361 in one go: window minimum over the entries that start before
// now+time_limit, get_max_tasks(min) > 0, and the concrete cores/slots taken
// from that minimum. cpu/mem minima are tested per entry (ballot); the core
// and gres masks are AND-reduced (see "the window minimum" in algebra.cuh).
__device__ __forceinline__ bool window_finish(const View& req, const Row& a0, uint64_t c0, uint64_t c1, uint64_t c2,
                                              uint64_t c3, uint64_t g0, uint64_t g1, Row& alloc) {
  if (a0.cpu_raw < req.cpu_raw || a0.mem < req.mem) return false;  // res_avail itself (JS.cpp:5310)
  Row wr;
  wr.cpu_raw = req.cpu_raw;  // the minima were checked entry by entry
  wr.mem = req.mem;
  wr.mem_sw = 0;
  wr.core[0] = a0.core[0] & warp_and64(c0);
  wr.core[1] = a0.core[1] & warp_and64(c1);
  wr.core[2] = a0.core[2] & warp_and64(c2);
  wr.core[3] = a0.core[3] & warp_and64(c3);
  wr.g[0] = a0.g[0] & warp_and64(g0);
  wr.g[1] = a0.g[1] & warp_and64(g1);
  return feasible_alloc(req, wr, alloc);
}

__device__ __forceinline__ bool node_test_now(const NodeRegs& nr, const View& req, bool exclusive, const Row& tot,
                                              const Row& a0, int64_t w_end, Row& alloc) {
  const bool in0 = nr.t0 < w_end, in1 = nr.t1 < w_end;
  if (exclusive) {  // every entry in the window must still hold res_total (JS.cpp:5285-5293)
    const bool ok = (!in0 || row_le(tot, nr.s0)) && (!in1 || row_le(tot, nr.s1));
    alloc = tot;
    return __all_sync(kFullMask, ok);
  }
  const bool ok = (!in0 || (nr.s0.cpu_raw >= req.cpu_raw && nr.s0.mem >= req.mem)) &&
                  (!in1 || (nr.s1.cpu_raw >= req.cpu_raw && nr.s1.mem >= req.mem));
  if (!__all_sync(kFullMask, ok)) return false;
  const bool e0 = in0 && !core_empty(nr.s0), e1 = in1 && !core_empty(nr.s1);
  const uint64_t ones = ~0ull;
  return window_finish(req, a0, (e0 ? nr.s0.core[0] : ones) & (e1 ? nr.s1.core[0] : ones),
                       (e0 ? nr.s0.core[1] : ones) & (e1 ? nr.s1.core[1] : ones),
                       (e0 ? nr.s0.core[2] : ones) & (e1 ? nr.s1.core[2] : ones),
                       (e0 ? nr.s0.core[3] : ones) & (e1 ? nr.s1.core[3] : ones),
                       (in0 ? nr.s0.g[0] : ones) & (in1 ? nr.s1.g[0] : ones),
                       (in0 ? nr.s0.g[1] : ones) & (in1 ? nr.s1.g[1] : ones), alloc);
}

__device__ __noinline__ bool node_test_now_big(const TimelineDev& tl, uint32_t g, uint32_t n, const View& req,
                                               bool exclusive, const Row& tot, const Row& a0, int64_t w_end,
                                               Row& alloc) {
  const TlEntry* E = tl.ent + (size_t)g * tl.cap;
  const uint32_t lane = lane_id();
  const uint64_t ones = ~0ull;
  uint64_t c0 = ones, c1 = ones, c2 = ones, c3 = ones, g0 = ones, g1 = ones;
  bool ok = true;
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t i = base + lane;
    bool in = false;
    if (i < n) {
      const TlEntry e = E[i];
      in = e.t < w_end;
      if (in) {
        if (exclusive) ok = ok && row_le(tot, e.seg);
        else {
          ok = ok && e.seg.cpu_raw >= req.cpu_raw && e.seg.mem >= req.mem;
          if (!core_empty(e.seg)) { c0 &= e.seg.core[0]; c1 &= e.seg.core[1]; c2 &= e.seg.core[2]; c3 &= e.seg.core[3]; }
          g0 &= e.seg.g[0];
          g1 &= e.seg.g[1];
        }
      }
    }
    if (!__all_sync(kFullMask, in)) break;  // entries are time-sorted
  }
  if (!__all_sync(kFullMask, ok)) return false;
  if (exclusive) { alloc = tot; return true; }
  return window_finish(req, a0, c0, c1, c2, c3, g0, g1, alloc);
}

// one 32-entry chunk of the earliest-fit scan: sat = this lane's entry holds the
// allocation and ends after T0; carry = start of the satisfied run that reaches
// the chunk from the left, or -1. Run starts come from the ballot of breakers.
__device__ __forceinline__ bool earliest_chunk(int64_t t, int64_t tend, bool sat, int64_t T0, int64_t limit,
                                               int64_t& carry, int64_t& result) {
  const uint32_t lane = lane_id();
  const unsigned bm = __ballot_sync(kFullMask, !sat);
  const unsigned below = bm & ((1u << lane) - 1u);
  const uint32_t r = below ? 32u - (uint32_t)__clz((int)below) : 0u;  // first lane of my run in this chunk
  int64_t rs = shfl_i64(t, (int)r);
  rs = rs > T0 ? rs : T0;
  if (!below && carry >= 0) rs = carry;  // the run started in an earlier chunk
  const bool ok = sat && (tend == kInf || tend - rs >= limit);
  const unsigned okm = __ballot_sync(kFullMask, ok);
  if (okm) {
    result = shfl_i64(rs, __ffs((int)okm) - 1);
    return true;
  }
  const int64_t last = shfl_i64(rs, 31);
  carry = ((bm >> 31) & 1u) ? -1 : last;
  return false;
}

// earliest t >= T0 such that `alloc` <= every entry overlapping [t, t+limit)
// on this node, kInf if none (per-node half of EarliestStartSubsetSelector,
// JobScheduler.h:731-784, 806-849).
__device__ __forceinline__ int64_t node_earliest(const NodeRegs& nr, const Row& alloc, int64_t T0, int64_t limit) {
  const uint32_t lane = lane_id();
  int64_t carry = -1, result = kInf;
  {
    int64_t tend = shfl_i64(nr.t0, lane + 1 < 32 ? (int)lane + 1 : (int)lane);
    const int64_t t32 = shfl_i64(nr.t1, 0);
    if (lane == 31) tend = t32;
    const bool sat = lane < nr.n && tend > T0 && row_le(alloc, nr.s0);
    if (earliest_chunk(nr.t0, tend, sat, T0, limit, carry, result)) return result;
  }
  if (nr.n > 32) {
    int64_t tend = shfl_i64(nr.t1, lane + 1 < 32 ? (int)lane + 1 : (int)lane);
    if (lane == 31) tend = kInf;
    const bool sat = lane + 32 < nr.n && tend > T0 && row_le(alloc, nr.s1);
    if (earliest_chunk(nr.t1, tend, sat, T0, limit, carry, result)) return result;
  }
  return kInf;
}

__device__ __noinline__ int64_t node_earliest_big(const TimelineDev& tl, uint32_t g, uint32_t n, const Row& alloc,
                                                  int64_t T0, int64_t limit) {
  const uint32_t lane = lane_id();
  const TlEntry* E = tl.ent + (size_t)g * tl.cap;
  int64_t carry = -1, result = kInf;
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t i = base + lane;
    int64_t t = kInf, tend = kInf;
    bool sat = false;
    if (i < n) {
      const TlEntry e = E[i];
      t = e.t;
      tend = (i + 1 < n) ? E[i + 1].t : kInf;
      sat = tend > T0 && row_le(alloc, e.seg);
    }
    if (earliest_chunk(t, tend, sat, T0, limit, carry, result)) return result;
  }
  return kInf;
}

// NodeState::UpdateResourceInNode (JobScheduler.h:334-453, allocation
// direction): breakpoints at start/end, subtract inside [start, end). Every
// lane writes its own entries to their new places; the lanes holding the
// covering segments also write the two inserted breakpoints. Lane 0 receives
// the (new) first segment. Returns the new entry count.
__device__ __forceinline__ uint32_t node_update(const TimelineDev& tl, uint32_t g, const NodeRegs& nr, int64_t start,
                                                int64_t end, const Row& alloc, Row& seg0) {
  const uint32_t lane = lane_id();
  TlEntry* E = tl.ent + (size_t)g * tl.cap;
  const uint32_t n = nr.n;
  const uint32_t i_s = __popc(__ballot_sync(kFullMask, nr.t0 <= start)) + __popc(__ballot_sync(kFullMask, nr.t1 <= start)) - 1;
  const uint32_t i_e = __popc(__ballot_sync(kFullMask, nr.t0 <= end)) + __popc(__ballot_sync(kFullMask, nr.t1 <= end)) - 1;
  const uint32_t ins_s = __any_sync(kFullMask, nr.t0 == start || nr.t1 == start) ? 0u : 1u;
  const uint32_t ins_e = __any_sync(kFullMask, nr.t0 == end || nr.t1 == end) ? 0u : 1u;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t j = lane + 32u * h;
    if (j >= n) continue;
    TlEntry e;
    e.t = h ? nr.t1 : nr.t0;
    e.seg = h ? nr.s1 : nr.s0;
    if (j == i_e && ins_e) {  // new breakpoint at `end` keeps the un-subtracted value
      TlEntry f;
      f.t = end;
      f.seg = e.seg;
      E[i_e + ins_s + 1] = f;
    }
    if (j > i_s) {
      const uint32_t nj = j + ins_s + (j > i_e ? ins_e : 0u);
      const bool sub = e.t < end;  // e.t > start here
      if (sub) row_sub(e.seg, alloc);
      if (sub || nj != j) E[nj] = e;
    } else if (j == i_s) {
      row_sub(e.seg, alloc);
      if (ins_s) {  // case #3: copy of the covering segment at `start`, minus the job
        e.t = start;
        E[i_s + 1] = e;
      } else {      // case #4: the key `start` exists
        E[i_s] = e;
      }
    }
  }
  if (lane == 0) {
    seg0 = nr.s0;
    if (i_s == 0 && !ins_s) row_sub(seg0, alloc);
    tl.n[g] = n + ins_s + ins_e;
  }
  return n + ins_s + ins_e;
}

__device__ __noinline__ uint32_t node_update_big(const TimelineDev& tl, uint32_t g, uint32_t n, int64_t start,
                                                 int64_t end, const Row& alloc, Row& seg0) {
  const uint32_t lane = lane_id();
  TlEntry* E = tl.ent + (size_t)g * tl.cap;
  uint32_t cnt_s = 0, cnt_e = 0;
  bool has_s = false, has_e = false;
  for (uint32_t base = 0; base < n; base += 32) {
    const uint32_t i = base + lane;
    const int64_t t = i < n ? E[i].t : kInf;
    const unsigned ms = __ballot_sync(kFullMask, t <= start), me = __ballot_sync(kFullMask, t <= end);
    cnt_s += __popc(ms);
    cnt_e += __popc(me);
    has_s = has_s || __any_sync(kFullMask, t == start);
    has_e = has_e || __any_sync(kFullMask, t == end);
    if (me != kFullMask) break;
  }
  const uint32_t i_s = cnt_s - 1, i_e = cnt_e - 1;
  const uint32_t ins_s = has_s ? 0u : 1u, ins_e = has_e ? 0u : 1u;
  const TlEntry es = E[i_s], ee = E[i_e];  // values before any modification
  __syncwarp();
  int64_t hi = (int64_t)n - 1;
  const int64_t lo = (int64_t)i_s + 1;
  while (hi >= lo) {  // move entries (i_s, n) upward, top chunk first
    const int64_t j = hi - lane;
    const bool act = j >= lo;
    TlEntry e;
    e.t = 0;
    if (act) e = E[j];
    __syncwarp();
    if (act) {
      if (e.t < end) row_sub(e.seg, alloc);
      E[(uint32_t)j + ins_s + ((uint32_t)j > i_e ? ins_e : 0u)] = e;
    }
    __syncwarp();
    hi -= 32;
  }
  if (lane == 0) {
    TlEntry e = es;
    row_sub(e.seg, alloc);
    if (ins_s) { e.t = start; E[i_s + 1] = e; } else { E[i_s] = e; }
    if (ins_e) { TlEntry f; f.t = end; f.seg = ee.seg; E[i_e + ins_s + 1] = f; }
    tl.n[g] = n + ins_s + ins_e;
  }
  __syncwarp();
  seg0 = E[0].seg;
  return n + ins_s + ins_e;
}

__device__ __forceinline__ Row node_total(const ClusterDev& cl, const Row* class_rows, const CommitSmem& sm,
                                          uint32_t base, uint32_t q) {
  const uint8_t c = sm.cls[q];
  if (c != 0xff) return class_rows[c];
  return cl.slot_total[base + q];
}

__device__ __forceinline__ bool key_lt(double c, uint32_t o, double kc, uint32_t ko) {
  return (c < kc) || (c == kc && o < ko);
}
__device__ __forceinline__ unsigned long long vmax8(unsigned long long a, unsigned long long b) {
  return (unsigned long long)__vmaxu4((unsigned)a, (unsigned)b) |
         (unsigned long long)__vmaxu4((unsigned)(a >> 32), (unsigned)(b >> 32)) << 32;
}

// ---- bucketed (cost, node) order: driver-warp operations -------------------
// NodeSelector::UpdateCost erases and re-inserts the node in a
// std::set<pair<double,NodeState*>> (JobScheduler.h:520-532). Here: remove u from
// its bucket, find the first bucket whose largest key is not below the new key,
// insert in place. Costs only grow inside a tick, so the search starts at u's
// old bucket. Returns false when the target bucket is full (caller rebuilds).
__device__ __forceinline__ void bucket_remove(CommitSmem& sm, uint32_t u) {
  const uint32_t lane = lane_id();
  const uint32_t b = sm.bkt[u];
  uint16_t* B = sm.bk + (size_t)b * kBucket;
  const uint32_t n = sm.bcnt[b];
  const uint16_t e0 = lane < n ? B[lane] : (uint16_t)0xffff;
  const uint16_t e1 = lane + 32 < n ? B[lane + 32] : (uint16_t)0xffff;
  const unsigned m0 = __ballot_sync(kFullMask, e0 == u), m1 = __ballot_sync(kFullMask, e1 == u);
  const uint32_t idx = m0 ? (uint32_t)__ffs((int)m0) - 1u : 32u + (uint32_t)__ffs((int)m1) - 1u;
  // entries after idx move one slot down (values are already in registers)
  if (lane > idx && lane < n) B[lane - 1] = e0;
  if (lane + 32 > idx && lane + 32 < n) B[lane + 31] = e1;
  __syncwarp();
  if (lane == 0) {
    sm.bcnt[b] = (uint16_t)(n - 1);
    sm.blast[b] = n > 1 ? B[n - 2] : (uint16_t)0xffff;
  }
  __syncwarp();
}

// Removes every node flagged in sm.pend — the picks list[0..np) of the batch in
// flight — from its bucket, one pass per distinct bucket. Driver warp only.
__device__ __noinline__ void bucket_remove_pending(CommitSmem& sm, uint32_t np) {
  const uint32_t lane = lane_id();
  const uint32_t myb = lane < np ? (uint32_t)sm.bkt[sm.list[lane]] : 0xffffffffu;
  unsigned todo = __ballot_sync(kFullMask, lane < np);
  while (todo) {
    const uint32_t b = __shfl_sync(kFullMask, myb, __ffs((int)todo) - 1);
    todo &= ~__ballot_sync(kFullMask, myb == b);
    uint16_t* B = sm.bk + (size_t)b * kBucket;
    const uint32_t n = sm.bcnt[b];
    const uint16_t e0 = lane < n ? B[lane] : (uint16_t)0xffff;
    const uint16_t e1 = lane + 32 < n ? B[lane + 32] : (uint16_t)0xffff;
    const bool r0 = lane < n && sm.pend[e0], r1 = lane + 32 < n && sm.pend[e1];
    const unsigned m0 = __ballot_sync(kFullMask, r0), m1 = __ballot_sync(kFullMask, r1);
    const unsigned below = (1u << lane) - 1u;
    if (lane < n && !r0) B[lane - (uint32_t)__popc(m0 & below)] = e0;
    if (lane + 32 < n && !r1) B[lane + 32 - (uint32_t)__popc(m0) - (uint32_t)__popc(m1 & below)] = e1;
    __syncwarp();
    if (lane == 0) {
      const uint32_t nn = n - (uint32_t)__popc(m0) - (uint32_t)__popc(m1);
      sm.bcnt[b] = (uint16_t)nn;
      sm.blast[b] = nn ? B[nn - 1] : (uint16_t)0xffff;
    }
    __syncwarp();
  }
}

// The bucket a node with key (new_cost, u) belongs into: the first non-empty
// bucket >= from_bucket whose last key is not below the key; if there is none,
// the last non-empty bucket. Read-only (blast[] and cost[]), so several warps
// may search at once, each for its own node.
__device__ __forceinline__ uint32_t bucket_find(const CommitSmem& sm, uint32_t u, double new_cost, uint32_t from_bucket) {
  const uint32_t lane = lane_id();
  uint32_t tb = 0xffffffffu, last_nonempty = 0xffffffffu;
  for (uint32_t start = from_bucket;; start = 0) {
    for (uint32_t b0 = start; b0 < sm.nb && tb == 0xffffffffu; b0 += 32) {
      const uint32_t b = b0 + lane;
      bool nonempty = false, ge = false;
      if (b < sm.nb) {
        const uint32_t o = sm.blast[b];
        if (o != 0xffffu) {
          nonempty = true;
          ge = !key_lt(sm.cost[o], o, new_cost, u);
        }
      }
      const unsigned mg = __ballot_sync(kFullMask, ge), mn = __ballot_sync(kFullMask, nonempty);
      if (mg) tb = b0 + (uint32_t)__ffs((int)mg) - 1u;
      if (mn) last_nonempty = b0 + 31u - (uint32_t)__clz((int)mn);
    }
    if (tb != 0xffffffffu) break;
    if (last_nonempty != 0xffffffffu) { tb = last_nonempty; break; }
    // nothing at or after u's old bucket (its tail was removed with it): the
    // nodes before it are the whole order now
    if (start == 0) { tb = from_bucket; break; }
  }
  return tb;
}
// Inserts u with key (new_cost, u) into bucket tb at its sorted place; false if
// the bucket is full (the caller re-deals the order).
__device__ __forceinline__ bool bucket_place(CommitSmem& sm, uint32_t u, double new_cost, uint32_t tb) {
  const uint32_t lane = lane_id();
  uint16_t* B = sm.bk + (size_t)tb * kBucket;
  const uint32_t n = sm.bcnt[tb];
#ifdef CRANE_EMU_DEBUG
  if (lane == 0) fprintf(stderr, "  insert u=%u key=%.6f -> tb=%u n=%u\n", u, new_cost, tb, n);
#endif
  if (n >= (uint32_t)kBucket) return false;
  const uint16_t e0 = lane < n ? B[lane] : (uint16_t)0xffff;
  const uint16_t e1 = lane + 32 < n ? B[lane + 32] : (uint16_t)0xffff;
  const bool l0 = lane < n && key_lt(sm.cost[e0], e0, new_cost, u);
  const bool l1 = lane + 32 < n && key_lt(sm.cost[e1], e1, new_cost, u);
  const uint32_t pos = (uint32_t)__popc(__ballot_sync(kFullMask, l0)) + (uint32_t)__popc(__ballot_sync(kFullMask, l1));
  if (lane >= pos && lane < n) B[lane + 1] = e0;
  if (lane + 32 >= pos && lane + 32 < n) B[lane + 33] = e1;
  __syncwarp();
  if (lane == 0) {
    sm.cost[u] = new_cost;
    B[pos] = (uint16_t)u;
    sm.bcnt[tb] = (uint16_t)(n + 1);
    if (pos == n) sm.blast[tb] = (uint16_t)u;
    sm.bkt[u] = (uint16_t)tb;
    const long long c = sm.cpu0[u];
    const unsigned long long gc = sm.gcnt[u];
    if (c > sm.bmax_cpu[tb]) sm.bmax_cpu[tb] = c;
    if (gc && c > sm.bmax_cpug[tb]) sm.bmax_cpug[tb] = c;
    sm.bmax_g[tb] = vmax8(sm.bmax_g[tb], gc);
    sm.bexact[tb] = 0;
  }
  __syncwarp();
  return true;
}
__device__ __forceinline__ bool bucket_insert(CommitSmem& sm, uint32_t u, double new_cost, uint32_t from_bucket) {
  return bucket_place(sm, u, new_cost, bucket_find(sm, u, new_cost, from_bucket));
}
// Deal sm.tmp[0..total) (already in (cost, node) order) out to the buckets,
// kBucketFill per bucket, and refresh bkt[] and the per-bucket bounds.
__device__ __noinline__ void bucket_deal(CommitSmem& sm, uint32_t total) {
  const uint32_t lane = lane_id();
  for (uint32_t b = 0; b < sm.nb; ++b) {
    const uint32_t lo = b * kBucketFill;
    const uint32_t n = lo < total ? (total - lo < (uint32_t)kBucketFill ? total - lo : (uint32_t)kBucketFill) : 0u;
    long long mc = INT64_MIN, mcg = INT64_MIN;
    unsigned long long mg = 0;
    if (lane < n) {
      const uint32_t q = sm.tmp[lo + lane];
      sm.bk[(size_t)b * kBucket + lane] = (uint16_t)q;
      sm.bkt[q] = (uint16_t)b;
      mc = sm.cpu0[q];
      mg = sm.gcnt[q];
      if (mg) mcg = mc;
    }
    for (int o = 16; o > 0; o >>= 1) {
      const long long oc = __shfl_xor_sync(kFullMask, mc, o), ocg = __shfl_xor_sync(kFullMask, mcg, o);
      mc = oc > mc ? oc : mc;
      mcg = ocg > mcg ? ocg : mcg;
      mg = vmax8(mg, __shfl_xor_sync(kFullMask, mg, o));
    }
    if (lane == 0) {
      sm.bcnt[b] = (uint16_t)n; sm.bmax_cpu[b] = mc; sm.bmax_cpug[b] = mcg; sm.bmax_g[b] = mg; sm.bexact[b] = 1;
      sm.blast[b] = n ? sm.tmp[lo + n - 1] : (uint16_t)0xffff;
    }
  }
  __syncwarp();
}
// Re-spread all bucketed nodes evenly, keeping the order. Driver warp only;
// runs when a bucket overflows (rare: a bucket must gain 32 nodes net).
__device__ __noinline__ void bucket_rebuild(CommitSmem& sm) {
  const uint32_t lane = lane_id();
  uint32_t rank = 0;
  for (uint32_t b = 0; b < sm.nb; ++b) {
    const uint32_t n = sm.bcnt[b];
    for (uint32_t i = lane; i < n; i += 32) sm.tmp[rank + i] = sm.bk[(size_t)b * kBucket + i];
    rank += n;
  }
  __syncwarp();
  bucket_deal(sm, rank);
}

#ifdef CRANE_EMU_DEBUG
// emulation-only invariant check of the bucketed order (driver lane 0)
inline void bucket_check(const CommitSmem& sm, uint32_t mp, const char* where, uint32_t job) {
  if (lane_id() != 0) return;
  std::vector<int> seen(mp, 0);
  double pc = -1.0; uint32_t pq = 0; bool have = false;
  for (uint32_t b = 0; b < sm.nb; ++b) {
    const uint32_t n = sm.bcnt[b];
    if ((n == 0) != (sm.blast[b] == 0xffff) || (n && sm.blast[b] != sm.bk[(size_t)b * kBucket + n - 1])) {
      fprintf(stderr, "[%s job %u] blast mismatch bucket %u n=%u blast=%u\n", where, job, b, n, sm.blast[b]); abort();
    }
    for (uint32_t i = 0; i < n; ++i) {
      const uint32_t q = sm.bk[(size_t)b * kBucket + i];
      if (q >= mp || seen[q]++) { fprintf(stderr, "[%s job %u] node %u twice/out of range in bucket %u\n", where, job, q, b); abort(); }
      if (sm.bkt[q] != b) { fprintf(stderr, "[%s job %u] bkt[%u]=%u but in bucket %u\n", where, job, q, sm.bkt[q], b); abort(); }
      if (have && !key_lt(pc, pq, sm.cost[q], q)) {
        fprintf(stderr, "[%s job %u] order violated at bucket %u idx %u: (%g,%u) then (%g,%u)\n", where, job, b, i, pc, pq, sm.cost[q], q); abort();
      }
      pc = sm.cost[q]; pq = q; have = true;
    }
  }
  for (uint32_t q = 0; q < mp; ++q)
    if (!seen[q]) { fprintf(stderr, "[%s job %u] node %u missing (pend=%u)\n", where, job, q, sm.pend[q]); abort(); }
}
#define BUCKET_CHECK(where, job) bucket_check(sm, mp, where, job)
#else
#define BUCKET_CHECK(where, job)
#endif

// worker commands (driver -> helpers, through shared memory + the CTA barrier)
enum : uint32_t {
  OP_NOW_K1 = 0,     // one-node job: test the node now; on success update it
  OP_BF_K1 = 1,      // one-node job: earliest start on the node; on success update it
  OP_TEST = 2,       // worker w tests list[w] for an immediate start
  OP_EARLY = 3,      // workers: earliest fit >= t0 over their share of list[0..n)
  OP_UPDATE_NOW = 4, // workers: allocate against the window minimum and update their share
  OP_UPDATE_BF = 5,  // workers: allocate against res_total and update their share
  OP_NOW_MULTI = 9,  // K <= warps: worker w tests list[w]; if all K pass, each updates its node
  OP_BF_MULTI = 10,  // K <= warps: worker w iterates the common earliest start with the others, then updates
  OP_BATCH_P = 6,    // batch: helper w+1 evaluates task w, and after the verdict commits it if its job is placed
  OP_SELECT = 12,    // helper t < n lists the candidates of batch job t
  OP_EXIT = 8,
};
struct BatchTask {   // one (job, node) pair of the batch in flight; its node is sm.list[w]
  uint32_t slot;     // ring slot of the job
  uint32_t mode;     // 0 = immediate start, 1 = backfill
  uint32_t tfirst;   // first task of the same job (its nodes are list[tfirst .. tfirst + node_num))
  uint32_t job;      // index of the job in the batch
};
struct BatchJob {    // one job of the batch being formed
  uint32_t slot;     // ring slot
  uint32_t K;        // node_num
  uint32_t need;     // nodes of this job and of the jobs before it in the batch: candidates worth listing
  uint32_t n0, n1;   // candidates found: pre-filter (immediate start) / capable (backfill)
  uint32_t pad;
};
struct BatchSel {    // the first `need` candidates of one job in cost order, with the cost each would get
  double nc0[kBatch], nc1[kBatch];
  uint16_t c0[kBatch], c1[kBatch];
};
struct CommitCmd {
  uint32_t kind, n, slot, first;  // OP_TEST: worker w handles list[first + w]; others: list[w], list[w+nw], ...
  int64_t t0;                     // OP_EARLY: T0; OP_UPDATE_*: start time
};
struct WorkerCtx {  // lives in shared memory; read-only after set-up
  ClusterDev cl;
  TimelineDev tl;
  PlaceDev out;
  CommitSmem sm;
  const JobQ* jobs;
  const Row* classrow;
  uint32_t* label;
  uint32_t* ok;      // [kBatch] verdicts of the batch being evaluated
  long long* tbuf;   // [2][32] per-iteration earliest fits of a multi-node job
  int64_t now, max_window;
  uint32_t base, max_jobs;
  uint32_t words;    // capability bitmap words per job row
  BatchJob* bj;      // [kBatch]
  BatchSel* sel;     // [kBatch]
  const BatchTask* task;         // [kBatch]
  uint32_t* joblabel;            // [kBatch] "some node is short of resources now" of a multi-node backfill in the batch
  const double* newcost;         // [kBatch] cost of a task's node once its job is placed
  uint32_t* tbk;                 // [kBatch] bucket the task's node goes back into (found by its helper)
  uint32_t* found;               // number of helpers that published tbk[] for the batch in flight
  const uint32_t* first_bucket;  // buckets before it are empty
};

// barrier protocol of the fused multi-node steps for a warp that holds no node
__device__ __noinline__ void multi_idle(const WorkerCtx* cxp, uint32_t kind, uint32_t n) {
  const WorkerCtx& cx = *cxp;
  if (kind == 9u) {  // OP_NOW_MULTI
    __syncthreads();
    return;
  }
  int64_t T0 = cx.now;  // OP_BF_MULTI
  for (uint32_t it = 0;; ++it) {
    const long long* buf = cx.tbuf + (it & 1u) * 32;
    __syncthreads();
    int64_t tmax = T0;
    for (uint32_t i = 0; i < n; ++i) tmax = buf[i] > tmax ? buf[i] : tmax;
    if (tmax == kInf || tmax == T0) break;
    T0 = tmax;
  }
}

// Candidate list of batch job t (one helper warp per job, all jobs of the batch
// at once, the order is not modified meanwhile): the first `need` nodes in cost
// order that pass capability + pre-filter (JobScheduler.cpp:5224-5266) and — if
// there are fewer — the first `need` capable nodes, which is where a backfill
// would go (JobScheduler.cpp:5269-5278). `need` covers the nodes the jobs before
// it in the batch may take away. Each listed node comes with the cost it gets
// when the job is placed on it (JobScheduler.h:46-52).
__device__ __noinline__ void select_step(const WorkerCtx* cxp, uint32_t t) {
  const WorkerCtx& cx = *cxp;
  const CommitSmem& sm = cx.sm;
  const uint32_t lane = lane_id();
  const BatchJob bj = cx.bj[t];
  const JobQ& jq = cx.jobs[bj.slot];
  const uint32_t* bits = sm.bits_ring + (size_t)bj.slot * cx.words;
  BatchSel& out = cx.sel[t];
  const uint32_t need = bj.need;
  const uint32_t jflags = jq.flags;
  const bool exclusive = jflags & 1u;
  const int64_t req_cpu = jq.req.cpu_raw;
  const uint64_t spec8 = jq.spec8;
  const uint32_t gnames = (jflags >> 8) & 0xffu;
  const uint32_t fb = *cx.first_bucket;
  uint32_t c = 0;
  for (uint32_t b = fb; c < need;) {
    // next bucket whose bounds admit a candidate (32 buckets per probe)
    uint32_t nbk = 0xffffffffu;
    for (uint32_t b0 = b; b0 < sm.nb && nbk == 0xffffffffu; b0 += 32) {
      const uint32_t bb = b0 + lane;
      bool prom = false;
      if (bb < sm.nb && sm.bcnt[bb])
        prom = exclusive || ((jflags & 2u) ? (sm.bmax_cpug[bb] >= req_cpu && gres_counts_ok(sm.bmax_g[bb], spec8, gnames, jq.name_need))
                                           : sm.bmax_cpu[bb] >= req_cpu);
      const unsigned pm = __ballot_sync(kFullMask, prom);
      if (pm) nbk = b0 + (uint32_t)__ffs((int)pm) - 1u;
    }
    if (nbk == 0xffffffffu) break;
    b = nbk;
    const uint32_t n = sm.bcnt[b];
    const uint16_t* B = sm.bk + (size_t)b * kBucket;
    bool any_cand = false;
    for (uint32_t h = 0; h < 2 && c < need; ++h) {
      const uint32_t idx = lane + 32 * h;
      bool cand = false;
      uint32_t q = 0;
      if (idx < n) {
        q = B[idx];
        cand = ((bits[q >> 5] >> (q & 31)) & 1u) && !sm.skip[q];
        if (cand && !exclusive)
          cand = sm.cpu0[q] >= req_cpu && (!(jflags & 2u) || gres_counts_ok(sm.gcnt[q], spec8, gnames, jq.name_need));
      }
      const unsigned cm = __ballot_sync(kFullMask, cand);
      any_cand = any_cand || cm != 0;
      const uint32_t rank = c + (uint32_t)__popc(cm & ((1u << lane) - 1u));
      if (cand && rank < need) out.c0[rank] = (uint16_t)q;
      c += (uint32_t)__popc(cm);
    }
    if (!any_cand && !sm.bexact[b]) {
      // nothing in this bucket passes the pre-filter: tighten its bounds to the
      // exact maxima (several helpers may do this at once; they write the same values)
      long long mc = INT64_MIN, mcg = INT64_MIN;
      unsigned long long mg = 0;
      for (uint32_t h = 0; h < 2; ++h) {
        const uint32_t idx = lane + 32 * h;
        if (idx < n) {
          const uint32_t q = B[idx];
          const long long c0 = sm.cpu0[q];
          const unsigned long long gc = sm.gcnt[q];
          mc = c0 > mc ? c0 : mc;
          if (gc && c0 > mcg) mcg = c0;
          mg = vmax8(mg, gc);
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        const long long oc = __shfl_xor_sync(kFullMask, mc, o), ocg = __shfl_xor_sync(kFullMask, mcg, o);
        mc = oc > mc ? oc : mc;
        mcg = ocg > mcg ? ocg : mcg;
        mg = vmax8(mg, __shfl_xor_sync(kFullMask, mg, o));
      }
      if (lane == 0) { sm.bmax_cpu[b] = mc; sm.bmax_cpug[b] = mcg; sm.bmax_g[b] = mg; sm.bexact[b] = 1; }
      __syncwarp();
    }
    ++b;
  }
  const uint32_t n0 = c < need ? c : need;
  uint32_t n1 = 0;
  if (n0 < need) {
    uint32_t cum = 0;
    for (uint32_t b = fb; b < sm.nb && cum < need; ++b) {
      const uint16_t* B = sm.bk + (size_t)b * kBucket;
      const uint32_t n = sm.bcnt[b];
      for (uint32_t h = 0; h < 2 && cum < need; ++h) {
        const uint32_t idx = lane + 32 * h;
        bool cap = false;
        uint32_t q = 0;
        if (idx < n) {
          q = B[idx];
          cap = ((bits[q >> 5] >> (q & 31)) & 1u) && !sm.skip[q];
        }
        const unsigned m = __ballot_sync(kFullMask, cap);
        const uint32_t rank = cum + (uint32_t)__popc(m & ((1u << lane) - 1u));
        if (cap && rank < need) out.c1[rank] = (uint16_t)q;
        cum += (uint32_t)__popc(m);
      }
    }
    n1 = cum < need ? cum : need;
  }
  __syncwarp();
  // the cost each listed node would get: lanes 0..7 the immediate-start list, 8..15 the backfill list
  {
    const bool second = lane >= (uint32_t)kBatch;
    const uint32_t i = second ? lane - (uint32_t)kBatch : lane;
    if (lane < 2u * (uint32_t)kBatch && i < (second ? n1 : n0)) {
      const uint32_t q = second ? out.c1[i] : out.c0[i];
      const int64_t tot_cpu = sm.cls[q] != 0xff ? cx.classrow[sm.cls[q]].cpu_raw : cx.cl.slot_total[cx.base + q].cpu_raw;
      const double nc = __dadd_rn(sm.cost[q], cost_delta(jq.time_limit, exclusive ? tot_cpu : req_cpu, tot_cpu));
      if (second) out.nc1[i] = nc; else out.nc0[i] = nc;
    }
  }
  if (lane == 0) { cx.bj[t].n0 = n0; cx.bj[t].n1 = n1; }
  __syncwarp();
}

// The per-node work of one command on this warp's share of sm.list[first, n)
// with the given stride: open the node's timeline once, then — depending on the
// command — the immediate-start test and allocation against the window minimum
// (JobScheduler.cpp:5285-5361), or the allocation against res_total and the
// earliest fit (JobScheduler.cpp:5381-5403, JobScheduler.h:806-849), and the
// timeline update with its outputs (JobScheduler.h:334-453, JobScheduler.cpp:5827).
// One out-of-line instance shared by the driver and the helpers.
// Returns: OP_NOW_K1/OP_TEST pass flag; OP_BF_K1 start time or kInf; OP_EARLY
// the max earliest fit over the share.
__device__ __noinline__ long long worker_step(const WorkerCtx* cxp, uint32_t kind, uint32_t n, uint32_t slot,
                                              int64_t t0, uint32_t first, uint32_t stride, uint32_t batch) {
  const WorkerCtx& cx = *cxp;
  const CommitSmem& sm = cx.sm;
  const uint32_t lane = lane_id();
  const JobQ& jq = cx.jobs[slot];
  const View req = jq.req;
  const bool exclusive = jq.flags & 1u;
  const int64_t limit = jq.time_limit;
  const int64_t now = cx.now;
  const int64_t w_end = now + limit;
  const uint32_t K = jq.node_num;
  long long result = (kind == OP_EARLY) ? (long long)t0 : 0;
#pragma unroll 1
  for (uint32_t k = first; k < n; k += ((batch == 2 || kind == OP_NOW_MULTI || kind == OP_BF_MULTI) ? 0x7fffffffu : stride)) {
    const uint32_t q = sm.list[k];
#ifdef CRANE_EMU_DEBUG
    if (q > 60000) { fprintf(stderr, "worker_step: kind=%u n=%u first=%u stride=%u k=%u q=%u tid=%u\n", kind, n, first, stride, k, q, threadIdx.x); abort(); }
#endif
    const uint32_t g = cx.base + q;
    const uint32_t ns = sm.nseg[q];
    const bool from_window = kind == OP_NOW_K1 || kind == OP_TEST || kind == OP_UPDATE_NOW || kind == OP_NOW_MULTI;
    const Row a0 = cx.tl.avail0[g];
    Row tot;
    row_zero(tot);
    if (exclusive || !from_window) tot = node_total(cx.cl, cx.classrow, sm, cx.base, q);
    NodeRegs nr;
    node_open(cx.tl, g, ns, nr);
    Row alloc;
    bool ok = true;
    if (from_window) {
      ok = ns <= 64 ? node_test_now(nr, req, exclusive, tot, a0, w_end, alloc)
                    : node_test_now_big(cx.tl, g, ns, req, exclusive, tot, a0, w_end, alloc);
    } else if (exclusive) {
      alloc = tot;
    } else {
      feasible_alloc(req, tot, alloc);
    }
    int64_t start = t0;
    if (kind == OP_NOW_K1 || kind == OP_TEST) result = ok ? 1 : 0;
    if (kind == OP_NOW_K1) start = now;
    if (kind == OP_NOW_MULTI) {
      // all n nodes must pass (they are the first n candidates in cost order);
      // the verdict barrier is shared with the idle warps (multi_idle)
      if (lane == 0) cx.ok[first] = ok ? 1u : 0u;
      __syncthreads();
      bool all = true;
      for (uint32_t i = 0; i < n; ++i) all = all && cx.ok[i] != 0;
      ok = all;
      result = all ? 1 : 0;
      start = now;
    }
    if (kind == OP_BF_MULTI) {
      // common earliest start of the chosen nodes: every warp keeps its node's
      // timeline in registers and iterates T <- max over nodes of the earliest
      // fit >= T to the fixed point (JobScheduler.h:806-849), one barrier a round.
      // On the one-job path the nodes are list[0..n) and the barrier is the
      // CTA's (idle warps follow in multi_idle); inside a batch they are this
      // job's tasks and the barrier is a named one among just their warps.
      const uint32_t l0 = batch == 2 ? cx.task[first].tfirst : 0u;
      const uint32_t cnt = batch == 2 ? K : n;
      const uint32_t bar_id = batch == 2 ? 1u + cx.task[first].job : 0u;
      int64_t T0 = now;
      ok = false;
      for (uint32_t it = 0;; ++it) {
        const int64_t t = ns <= 64 ? node_earliest(nr, alloc, T0, limit) : node_earliest_big(cx.tl, g, ns, alloc, T0, limit);
        long long* buf = cx.tbuf + (it & 1u) * 32;
        if (lane == 0) buf[first] = t;
        if (batch == 2) named_bar_sync(bar_id, cnt * 32u); else __syncthreads();
        int64_t tmax = T0;
        for (uint32_t i = 0; i < cnt; ++i) tmax = buf[l0 + i] > tmax ? buf[l0 + i] : tmax;
        if (tmax == kInf) break;
        if (tmax == T0) { ok = T0 - now <= cx.max_window; break; }  // JobScheduler.h:809
        T0 = tmax;
      }
      start = T0;
      result = ok ? T0 : kInf;
    }
    if (kind == OP_BF_K1 || kind == OP_EARLY) {
      const int64_t t = ns <= 64 ? node_earliest(nr, alloc, t0, limit) : node_earliest_big(cx.tl, g, ns, alloc, t0, limit);
      if (kind == OP_EARLY) {
        result = t > result ? t : result;
      } else {
        ok = t != kInf && t - now <= cx.max_window;  // `current_time - now > kAlgoMaxTimeWindow` (JobScheduler.h:809)
        start = t;
        result = ok ? t : kInf;
      }
    }
    if (batch == 2) {
      // task `first` of a batch of `stride` one-node jobs: publish the verdict, wait
      // for the others, and commit only if every task before this one succeeded
      if (lane == 0) cx.ok[first] = ok ? 1u : 0u;
      __syncthreads();
      uint32_t f = stride;
      for (uint32_t i = 0; i < stride; ++i)
        if (!cx.ok[i]) { f = i; break; }
      // a job is placed only if all its nodes pass: cut at the first task of the failing job
      if (f < stride) f = cx.task[f].tfirst;
      ok = first < f;
      result = (long long)f;
      // where the node goes back into the order (the driver has taken all picks
      // out by now): at its new cost if the job is placed, at the old one otherwise.
      // Read-only search, published to the driver, which does the inserts while
      // the timelines are updated below.
      const uint32_t tbk = bucket_find(sm, q, ok ? cx.newcost[first] : sm.cost[q], *cx.first_bucket);
      if (lane == 0) {
        __threadfence_block();
        *(volatile uint32_t*)&cx.tbk[first] = tbk;  // the driver polls this word
      }
    }
    const bool do_update = ((kind == OP_NOW_K1 || kind == OP_BF_K1 || kind == OP_NOW_MULTI || kind == OP_BF_MULTI) && ok) ||
                           kind == OP_UPDATE_NOW || kind == OP_UPDATE_BF;
    if (do_update) {
      const int64_t end = start + limit;
      Row seg0;
      const uint32_t nn = ns <= 64 ? node_update(cx.tl, g, nr, start, end, alloc, seg0)
                                   : node_update_big(cx.tl, g, ns, start, end, alloc, seg0);
      uint32_t rank = 0;  // node-index ascending output slot (deviation D3)
      if ((n > 1 && !batch) || (batch == 2 && K > 1)) {
        const uint32_t l0 = batch == 2 ? cx.task[first].tfirst : 0u;  // the job's nodes are list[l0 .. l0+K)
        for (uint32_t m = lane; m < K; m += 32) rank += sm.list[l0 + m] < q ? 1u : 0u;
        for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(kFullMask, rank, o);
      }
      if (lane == 0) {
        sm.nseg[q] = (uint16_t)nn;
        if (nn >= cx.max_jobs) sm.skip[q] = 1;
        sm.cpu0[q] = seg0.cpu_raw;
        sm.gcnt[q] = (seg0.g[0] | seg0.g[1]) ? pack_gres_counts(seg0) : 0ull;
        const uint32_t dst = jq.alloc_off + rank;
        cx.out.alloc_node[dst] = cx.cl.slot_node[g];
        cx.out.alloc_ntasks[dst] = jq.ntasks_per_node;
        cx.out.alloc_res[dst] = alloc;
        // pending-reason label for future starts (JobScheduler.cpp:5842-5848)
        const bool short_now = start != now && !row_le(alloc, a0);
        if (batch) {  // job of a batch: its job-level outputs are written here (by each of its tasks, same values)
          cx.out.start_time[jq.job] = start;
          cx.out.end_time[jq.job] = end;
          cx.out.n_alloc[jq.job] = K;
          if (K == 1 || start == now)
            cx.out.reason[jq.job] = start == now ? CRANE_REASON_NONE : (short_now ? CRANE_REASON_RESOURCE : CRANE_REASON_PRIORITY);
          else if (short_now)
            atomicOr(&cx.joblabel[cx.task[first].job], 1u);
        } else if (short_now) {
          atomicOr(cx.label, 1u);
        }
      }
      __syncwarp();  // lane 0's shared-memory writes are visible to the whole warp
      if (batch == 2 && K > 1 && start != now) {
        // multi-node backfill inside a batch: "Resource" if any of its nodes is
        // short now, else "Priority" (JobScheduler.cpp:5842-5848)
        named_bar_sync(1u + cx.task[first].job, K * 32u);
        if (lane == 0 && first == cx.task[first].tfirst)
          cx.out.reason[jq.job] = cx.joblabel[cx.task[first].job] ? CRANE_REASON_RESOURCE : CRANE_REASON_PRIORITY;
      }
    }
  }
  return result;
}

__global__ void __launch_bounds__(kCommitThreads, 1) k_commit(CommitArgs a) {
  CRANE_DYN_SMEM(unsigned char, smem_raw);
  const uint32_t part = blockIdx.x;
  const uint32_t base = a.cl.part_base[part];
  const uint32_t mp = a.cl.part_base[part + 1] - base;
  const uint32_t words = a.words_per_row;
  const uint32_t lane = lane_id();
  const uint32_t wid = warp_id();
  const uint32_t nw = blockDim.x >> 5;

  CommitSmem sm;
  sm.nb = commit_nbuckets(mp);
  {
    unsigned char* ptr = smem_raw;  // 16-byte aligned; widest element types first
    sm.bits_ring = reinterpret_cast<uint32_t*>(ptr); ptr += (size_t)kRing * words * 4;
    sm.cost = reinterpret_cast<double*>(ptr); ptr += (size_t)mp * 8;
    sm.cpu0 = reinterpret_cast<long long*>(ptr); ptr += (size_t)mp * 8;
    sm.gcnt = reinterpret_cast<unsigned long long*>(ptr); ptr += (size_t)mp * 8;
    sm.bmax_cpu = reinterpret_cast<long long*>(ptr); ptr += (size_t)sm.nb * 8;
    sm.bmax_cpug = reinterpret_cast<long long*>(ptr); ptr += (size_t)sm.nb * 8;
    sm.bmax_g = reinterpret_cast<unsigned long long*>(ptr); ptr += (size_t)sm.nb * 8;
    sm.bk = reinterpret_cast<uint16_t*>(ptr); ptr += (size_t)sm.nb * kBucket * 2;
    sm.bcnt = reinterpret_cast<uint16_t*>(ptr); ptr += (size_t)sm.nb * 2;
    sm.blast = reinterpret_cast<uint16_t*>(ptr); ptr += (size_t)sm.nb * 2;
    sm.bkt = reinterpret_cast<uint16_t*>(ptr); ptr += (size_t)mp * 2;
    sm.list = reinterpret_cast<uint16_t*>(ptr); ptr += (size_t)mp * 2;
    sm.tmp = reinterpret_cast<uint16_t*>(ptr); ptr += (size_t)mp * 2;
    sm.nseg = reinterpret_cast<uint16_t*>(ptr); ptr += (size_t)mp * 2;
    sm.skip = ptr; ptr += mp;
    sm.cls = ptr; ptr += mp;
    sm.bexact = ptr; ptr += sm.nb;
    sm.pend = ptr;
  }
  __shared__ JobQ s_jobs[kRing];
  __shared__ __align__(8) uint64_t s_bar[kRing];
  __shared__ Row s_classrow[kMaxClasses];
  __shared__ CommitCmd s_cmd;
  __shared__ BatchTask s_task[kBatch];
  __shared__ BatchJob s_bj[kBatch];
  __shared__ BatchSel s_sel[kBatch];
  __shared__ uint32_t s_first_bucket;
  __shared__ uint32_t s_joblabel[kBatch];
  __shared__ uint32_t s_tbk[kBatch];
  __shared__ uint32_t s_found;
  __shared__ __align__(16) uint16_t s_pick[2][16];
  __shared__ uint32_t s_ok[32];
  __shared__ long long s_tbuf[2][32];
  __shared__ double s_newcost[kBatch];
  __shared__ WorkerCtx s_cx;
  __shared__ long long s_res[32];   // per-worker result of a multi-warp step
  __shared__ uint32_t s_label;

  // ---- load node state (all warps) --------------------------------------
  for (uint32_t q = threadIdx.x; q < mp; q += blockDim.x) {
    const uint32_t g = base + q;
    sm.cost[q] = a.tl.cost0[g];
    const Row s0 = a.tl.ent[(size_t)g * a.tl.cap].seg;
    sm.cpu0[q] = s0.cpu_raw;
    sm.gcnt[q] = pack_gres_counts(s0);
    sm.skip[q] = a.tl.skip[g];
    sm.nseg[q] = (uint16_t)a.tl.n[g];
    sm.cls[q] = a.cl.slot_class[g];
    sm.pend[q] = 0;
  }
  if (threadIdx.x < kMaxClasses) s_classrow[threadIdx.x] = a.cl.class_rows[(size_t)part * kMaxClasses + threadIdx.x];
  for (uint32_t b = threadIdx.x; b < sm.nb; b += blockDim.x) sm.bcnt[b] = 0;
  if (threadIdx.x == 0) {
    s_label = 0;
    s_cx.cl = a.cl; s_cx.tl = a.tl; s_cx.out = a.out; s_cx.sm = sm; s_cx.jobs = s_jobs; s_cx.classrow = s_classrow;
    s_cx.label = &s_label; s_cx.ok = s_ok; s_cx.tbuf = &s_tbuf[0][0]; s_cx.now = a.now; s_cx.max_window = a.max_window; s_cx.base = base; s_cx.max_jobs = a.max_jobs;
    s_cx.words = words; s_cx.bj = s_bj; s_cx.sel = s_sel; s_cx.task = s_task; s_cx.first_bucket = &s_first_bucket; s_cx.joblabel = s_joblabel; s_cx.newcost = s_newcost; s_cx.tbk = s_tbk; s_cx.found = &s_found;
    for (int s = 0; s < kRing; ++s) mbar_init(&s_bar[s], 1);
    fence_mbar_init();
  }
  __syncthreads();
  // initial order = ascending (cost, node): rank sort into sm.list, dealt out by the driver
  for (uint32_t q = threadIdx.x; q < mp; q += blockDim.x) {
    const double c = sm.cost[q];
    uint32_t rank = 0;
    for (uint32_t o = 0; o < mp; ++o) {
      const double co = sm.cost[o];
      rank += (co < c || (co == c && o < q)) ? 1u : 0u;
    }
    sm.tmp[rank] = (uint16_t)q;
  }
  __syncthreads();
  if (wid == 0) bucket_deal(sm, mp);
  __syncthreads();

  const uint32_t r_begin = a.part_job_off[part], r_end = a.part_job_off[part + 1];
  const uint32_t njobs = r_end - r_begin;
  const uint32_t row_bytes = words * 4;
  auto issue = [&](uint32_t i) {  // one driver lane per record
    const uint32_t slot = i % kRing;
    mbar_expect_tx(&s_bar[slot], (uint32_t)sizeof(JobQ) + row_bytes);
    tma_load_1d(&s_jobs[slot], &a.jobq[r_begin + i], (uint32_t)sizeof(JobQ), &s_bar[slot]);
    tma_load_1d(sm.bits_ring + (size_t)slot * words, a.bitmap + (size_t)(r_begin + i) * words, row_bytes, &s_bar[slot]);
  };

  if (wid != 0) {
    // ======================= helpers: parked on the barrier ================
    for (;;) {
      __syncthreads();  // a command is ready
      const CommitCmd c = s_cmd;
      if (c.kind == OP_EXIT) break;
      long long r = 0;
      if (c.kind == OP_BATCH_P) {
        const uint32_t t = wid - 1;  // task t of the batch: node sm.list[t], job in ring slot s_task[t].slot
        if (t < c.n) {
          const BatchTask tk = s_task[t];
          const uint32_t kind = !tk.mode ? OP_NOW_K1 : (s_jobs[tk.slot].node_num > 1 ? OP_BF_MULTI : OP_BF_K1);
          r = worker_step(&s_cx, kind, t + 1, tk.slot, s_cx.now, t, c.n, 2);
        } else {
          __syncthreads();  // the verdict barrier inside the batch step
        }
      } else if (c.kind == OP_SELECT) {
        if (wid - 1 < c.n) {
          select_step(&s_cx, wid - 1);
          if (lane == 0) {
            __threadfence_block();
            atomicAdd(&s_found, 1u);  // the driver polls this count instead of a CTA barrier
          }
        }
        continue;
      } else if (c.kind == OP_NOW_MULTI || c.kind == OP_BF_MULTI) {
        if (wid < c.n) r = worker_step(&s_cx, c.kind, c.n, c.slot, c.t0, wid, 1, 0);
        else multi_idle(&s_cx, c.kind, c.n);
      } else if (c.kind == OP_TEST) {
        r = worker_step(&s_cx, c.kind, c.n, c.slot, c.t0, c.first + wid, 0x7fffffffu, 0);
      } else {
        r = worker_step(&s_cx, c.kind, c.n, c.slot, c.t0, wid, nw, 0);
      }
      if (lane == 0) s_res[wid] = r;
      __syncthreads();  // results are in
    }
    return;
  }

  // ========================= driver warp ====================================
  // runs one multi-warp step: publish the command, join the helpers, reduce
  auto multi_step = [&](uint32_t kind, uint32_t n, uint32_t slot, int64_t t0, uint32_t first) -> long long {
    if (lane == 0) { s_cmd.kind = kind; s_cmd.n = n; s_cmd.slot = slot; s_cmd.first = first; s_cmd.t0 = t0; }
    __syncthreads();
    const long long r0 = kind == OP_TEST ? worker_step(&s_cx, kind, n, slot, t0, first, 0x7fffffffu, 0)
                                         : worker_step(&s_cx, kind, n, slot, t0, 0, nw, 0);
    if (lane == 0) s_res[0] = r0;
    __syncthreads();
    long long acc = r0;
    if (kind == OP_EARLY)
      for (uint32_t w = 1; w < nw; ++w) acc = s_res[w] > acc ? s_res[w] : acc;
    return acc;
  };

  uint32_t first_bucket = 0;  // buckets before it are empty
  uint32_t issued = 0;        // jobs whose records were requested from the ring
  PROF_DECL;

  // ring slot of job i is free once job i-kRing is finished
  auto ensure_issued = [&](uint32_t finished) {
    __syncwarp();  // no lane is still reading the ring slots of finished jobs
    const uint32_t hi = njobs < finished + (uint32_t)kRing ? njobs : finished + (uint32_t)kRing;
    if (issued < hi) {  // at most kRing <= 32 records: one lane each
      if (issued + lane < hi) issue(issued + lane);
      issued = hi;
    }
    __syncwarp();
  };
  // re-key node q to cost nc in the bucketed order (JobScheduler.h:520-532)
  auto rekey = [&](uint32_t q, double nc, uint32_t from) {
    bucket_remove(sm, q);
    if (!bucket_insert(sm, q, nc, from)) {
      // the target bucket is full: spread the other nodes evenly again (q is in
      // no bucket right now), then insert into a bucket with room
      bucket_rebuild(sm);
      first_bucket = 0;
      bucket_insert(sm, q, nc, 0);
    }
  };

  // a node the helpers could not put back (tiny partition, or its target bucket
  // was full): serial insert; once the order was re-dealt the old bucket index
  // of the remaining nodes means nothing
  auto leftover_insert = [&](uint32_t q, double nc, bool& rebuilt) {
    if (!bucket_insert(sm, q, nc, rebuilt ? 0u : (uint32_t)sm.bkt[q])) {
      bucket_rebuild(sm);
      PROF_CNT(12, 1);
      first_bucket = 0;
      rebuilt = true;
      bucket_insert(sm, q, nc, 0);
    }
    if (lane == 0) sm.pend[q] = 0;
    __syncwarp();
  };

  // ---- one job, start to finish (any node_num) ------------------------------
  auto process_single = [&](uint32_t ji) {
    PROF(15);
    const uint32_t slot = ji % kRing;
    mbar_wait(&s_bar[slot], (ji / kRing) & 1u);
    const JobQ& jq = s_jobs[slot];
    const uint32_t* bits = sm.bits_ring + (size_t)slot * words;
    const uint32_t K = jq.node_num;
    const uint32_t jflags = jq.flags;
    const bool exclusive = jflags & 1u;
    const int64_t limit = jq.time_limit;
    const int64_t req_cpu = jq.req.cpu_raw;
    const uint64_t spec8 = jq.spec8;
    const uint32_t gnames = (jflags >> 8) & 0xffu;
    while (first_bucket + 1 < sm.nb && sm.bcnt[first_bucket] == 0) ++first_bucket;
    if (lane == 0) s_label = 0;
    __syncwarp();
    PROF(0);

    int64_t start_time = 0;
    bool placed = false;
    uint32_t nsel = 0;  // nodes selected for an immediate start (in sm.list[0..nsel))

    // ---- multi-node job that fits the CTA (K <= warps): fused steps ---------
    // Immediate start: the first K pre-filter candidates in cost order are
    // tested in parallel; if all pass they are the reference's pick and are
    // updated in the same step. With fewer than K candidates the job can only
    // be backfilled: the first K capable nodes iterate their common earliest
    // start with their timelines resident in registers.
    bool handled = false;
    auto fused = [&](uint32_t kind) -> long long {
      if (lane == 0) { s_cmd.kind = kind; s_cmd.n = K; s_cmd.slot = slot; s_cmd.first = 0; s_cmd.t0 = a.now; }
      __syncthreads();
      const long long r = worker_step(&s_cx, kind, K, slot, a.now, 0, 1, 0);
      __syncthreads();
      return r;
    };
    if (K > 1 && K <= nw && K <= mp) {
      uint32_t c = 0;  // pre-filter candidates found, in cost order, in sm.list[0..c)
      for (uint32_t b = first_bucket; c < K;) {
        // next bucket whose bounds admit a candidate (32 buckets per probe)
        uint32_t nbk = 0xffffffffu;
        for (uint32_t b0 = b; b0 < sm.nb && nbk == 0xffffffffu; b0 += 32) {
          const uint32_t bb = b0 + lane;
          bool prom = false;
          if (bb < sm.nb && sm.bcnt[bb])
            prom = exclusive || ((jflags & 2u) ? (sm.bmax_cpug[bb] >= req_cpu && gres_counts_ok(sm.bmax_g[bb], spec8, gnames, jq.name_need))
                                               : sm.bmax_cpu[bb] >= req_cpu);
          const unsigned pm = __ballot_sync(kFullMask, prom);
          if (pm) nbk = b0 + (uint32_t)__ffs((int)pm) - 1u;
        }
        if (nbk == 0xffffffffu) break;
        b = nbk;
        const uint32_t n = sm.bcnt[b];
        const uint16_t* B = sm.bk + (size_t)b * kBucket;
        for (uint32_t h = 0; h < 2 && c < K; ++h) {
          const uint32_t idx = lane + 32 * h;
          bool cand = false;
          uint32_t q = 0;
          if (idx < n) {
            q = B[idx];
            cand = ((bits[q >> 5] >> (q & 31)) & 1u) && !sm.skip[q];
            if (cand && !exclusive)
              cand = sm.cpu0[q] >= req_cpu && (!(jflags & 2u) || gres_counts_ok(sm.gcnt[q], spec8, gnames, jq.name_need));
          }
          const unsigned cm = __ballot_sync(kFullMask, cand);
          const uint32_t rank = c + (uint32_t)__popc(cm & ((1u << lane) - 1u));
          if (cand && rank < K) sm.list[rank] = (uint16_t)q;
          c += (uint32_t)__popc(cm);
        }
        ++b;
      }
      __syncwarp();
      if (c >= K) {
        if (fused(OP_NOW_MULTI)) { placed = true; start_time = a.now; handled = true; nsel = K; }
        // else: some candidate failed the exact test -> the general path below
      } else {
        // fewer than K candidates: no immediate start is possible
        handled = true;
        uint32_t cum = 0;
        for (uint32_t b = first_bucket; b < sm.nb && cum < K; ++b) {
          const uint16_t* B = sm.bk + (size_t)b * kBucket;
          const uint32_t n = sm.bcnt[b];
          for (uint32_t h = 0; h < 2 && cum < K; ++h) {
            const uint32_t idx = lane + 32 * h;
            bool cap = false;
            uint32_t q = 0;
            if (idx < n) {
              q = B[idx];
              cap = ((bits[q >> 5] >> (q & 31)) & 1u) && !sm.skip[q];
            }
            const unsigned m = __ballot_sync(kFullMask, cap);
            const uint32_t rank = cum + (uint32_t)__popc(m & ((1u << lane) - 1u));
            if (cap && rank < K) sm.list[rank] = (uint16_t)q;
            cum += (uint32_t)__popc(m);
          }
        }
        __syncwarp();
        if (cum >= K) {
          const long long t = fused(OP_BF_MULTI);
          if (t != kInf) { placed = true; start_time = t; }
        }
      }
    }

    // ---- immediate start: walk the buckets in cost order -------------------
    // (JobScheduler.cpp:5224-5336). A bucket whose bounds cannot satisfy the
    // pre-filter holds no candidate and is skipped.
    if (K <= mp && !handled) {
      uint32_t b = first_bucket;
      while (nsel < K) {
        // next bucket that may hold a candidate
        uint32_t nbk = 0xffffffffu;
        for (uint32_t b0 = b; b0 < sm.nb && nbk == 0xffffffffu; b0 += 32) {
          const uint32_t bb = b0 + lane;
          bool prom = false;
          if (bb < sm.nb && sm.bcnt[bb])
            prom = exclusive || ((jflags & 2u) ? (sm.bmax_cpug[bb] >= req_cpu && gres_counts_ok(sm.bmax_g[bb], spec8, gnames, jq.name_need))
                                               : sm.bmax_cpu[bb] >= req_cpu);
          const unsigned pm = __ballot_sync(kFullMask, prom);
          if (pm) nbk = b0 + (uint32_t)__ffs((int)pm) - 1u;
        }
        if (nbk == 0xffffffffu) break;
        b = nbk;
        const uint16_t* B = sm.bk + (size_t)b * kBucket;
        const uint32_t n = sm.bcnt[b];
        bool any_cand = false;
        for (uint32_t h = 0; h < 2 && nsel < K; ++h) {
          const uint32_t idx = lane + 32 * h;
          bool cand = false;
          uint32_t q = 0;
          if (idx < n) {
            q = B[idx];
            cand = ((bits[q >> 5] >> (q & 31)) & 1u) && !sm.skip[q];
            if (cand && !exclusive)
              cand = sm.cpu0[q] >= req_cpu && (!(jflags & 2u) || gres_counts_ok(sm.gcnt[q], spec8, gnames, jq.name_need));
          }
          unsigned cm = __ballot_sync(kFullMask, cand);
          any_cand = any_cand || cm != 0;
          PROF_CNT(8, __popc(cm));
          while (cm && nsel < K) {
            {
              // hand the next candidates to the workers, one each, in order
              // (as many as there are workers: tests have no side effects, and the
              // walk usually has to go past a few candidates that fail the exact test)
              const uint32_t W = nw;
              uint32_t cnt = 0;
              unsigned taken = 0;
              while (cm && cnt < W) {
                const uint32_t l = (uint32_t)__ffs((int)cm) - 1u;
                cm &= cm - 1u;
                taken |= 1u << l;
                ++cnt;
              }
              // stage them after the already selected nodes: list[nsel .. nsel+cnt)
              if ((taken >> lane) & 1u) sm.list[nsel + __popc(taken & ((1u << lane) - 1u))] = (uint16_t)q;
              __syncwarp();
              PROF_CNT(9, 1);
              // worker w tests list[nsel + w]
              multi_step(OP_TEST, nsel + cnt, slot, a.now, nsel);
              // keep the passing ones, in order, compacted at list[nsel..)
              uint32_t keep = nsel;
              for (uint32_t w = 0; w < cnt; ++w) {
                if (s_res[w] && keep < K) {
                  const uint16_t v = sm.list[nsel + w];
                  __syncwarp();
                  if (lane == 0) sm.list[keep] = v;
                  ++keep;
                }
              }
              __syncwarp();
              nsel = keep;
            }
          }
        }
        if (!any_cand && !sm.bexact[b]) {
          // nothing in this bucket passes the pre-filter: tighten its bounds to
          // the exact maxima (once; an insert makes them inexact again)
          long long mc = INT64_MIN, mcg = INT64_MIN;
          unsigned long long mg = 0;
          for (uint32_t h = 0; h < 2; ++h) {
            const uint32_t idx = lane + 32 * h;
            if (idx < n) {
              const uint32_t q = B[idx];
              const long long c0 = sm.cpu0[q];
              const unsigned long long gc = sm.gcnt[q];
              mc = c0 > mc ? c0 : mc;
              if (gc && c0 > mcg) mcg = c0;
              mg = vmax8(mg, gc);
            }
          }
          for (int o = 16; o > 0; o >>= 1) {
            const long long oc = __shfl_xor_sync(kFullMask, mc, o), ocg = __shfl_xor_sync(kFullMask, mcg, o);
            mc = oc > mc ? oc : mc;
            mcg = ocg > mcg ? ocg : mcg;
            mg = vmax8(mg, __shfl_xor_sync(kFullMask, mg, o));
          }
          if (lane == 0) { sm.bmax_cpu[b] = mc; sm.bmax_cpug[b] = mcg; sm.bmax_g[b] = mg; sm.bexact[b] = 1; }
          __syncwarp();
        }
        ++b;
      }
    }
    PROF(2);

    if (handled) {
      // the fused step did everything
    } else if (nsel >= K && K <= mp) {
      placed = true;
      start_time = a.now;
      PROF_CNT(10, 1);
      multi_step(OP_UPDATE_NOW, K, slot, a.now, 0);
    } else {
      // ---- backfill: the first K capable nodes in cost order, allocation
      // against res_total, earliest common start (JobScheduler.cpp:5269-5278,
      // 5371-5404; JobScheduler.h:806-849)
      uint32_t cum = 0;
      for (uint32_t b = first_bucket; b < sm.nb && cum < K; ++b) {
        const uint16_t* B = sm.bk + (size_t)b * kBucket;
        const uint32_t n = sm.bcnt[b];
        for (uint32_t h = 0; h < 2 && cum < K; ++h) {
          const uint32_t idx = lane + 32 * h;
          bool cap = false;
          uint32_t q = 0;
          if (idx < n) {
            q = B[idx];
            cap = ((bits[q >> 5] >> (q & 31)) & 1u) && !sm.skip[q];
          }
          const unsigned m = __ballot_sync(kFullMask, cap);
          const uint32_t rank = cum + (uint32_t)__popc(m & ((1u << lane) - 1u));
          if (cap && rank < K) sm.list[rank] = (uint16_t)q;
          cum += (uint32_t)__popc(m);
        }
      }
      __syncwarp();
      PROF(4);
      if (K <= mp && cum >= K) {
        if (K == 1) {
          const long long t = worker_step(&s_cx, OP_BF_K1, 1, slot, a.now, 0, 1, 0);
          PROF_CNT(11, 1);
          if (t != kInf) { placed = true; start_time = t; }
        } else {
          int64_t Tcur = a.now;
          bool found = false, failed = false;
          while (!found && !failed) {
            PROF_CNT(11, 1);
            const int64_t tmax = multi_step(OP_EARLY, K, slot, Tcur, 0);
            if (tmax == kInf) failed = true;
            else if (tmax == Tcur) found = true;
            else Tcur = tmax;
          }
          if (found && Tcur - a.now <= a.max_window) {
            placed = true;
            start_time = Tcur;
            multi_step(OP_UPDATE_BF, K, slot, Tcur, 0);
          }
        }
      }
      PROF(5);
    }

    // ---- job-level outputs and re-keying of the chosen nodes ---------------
    if (placed) {
      PROF(6);
      if (lane == 0) {
        a.out.start_time[jq.job] = start_time;
        a.out.end_time[jq.job] = start_time + limit;
        a.out.n_alloc[jq.job] = K;
        uint8_t reason = CRANE_REASON_NONE;
        if (start_time != a.now) reason = s_label ? CRANE_REASON_RESOURCE : CRANE_REASON_PRIORITY;
        a.out.reason[jq.job] = reason;
      }
      // cost += (end-start) * cpu ratio (JobScheduler.h:46-52); the allocation's
      // cpu is the job's per-node request, or the node total for exclusive jobs
#pragma unroll 1
      for (uint32_t k = 0; k < K; ++k) {
        const uint32_t q = sm.list[k];
        const int64_t tot_cpu = sm.cls[q] != 0xff ? s_classrow[sm.cls[q]].cpu_raw : a.cl.slot_total[base + q].cpu_raw;
        const double delta = cost_delta(limit, exclusive ? tot_cpu : req_cpu, tot_cpu);
        const double oc = sm.cost[q];
        const double nc = __dadd_rn(oc, delta);
        if (nc > oc) rekey(q, nc, sm.bkt[q]);
      }
      PROF(7);
    } else {
      if (lane == 0) a.out.reason[jq.job] = CRANE_REASON_RESOURCE;  // JobScheduler.cpp:5802
    }
    };

  // ---- dispatcher: consecutive jobs go out as batches ---------------------------
  // A batch is a run of jobs with at most kBatch nodes in total (one helper warp
  // per node). Four steps, each parallel over the helpers:
  //  select   every job lists its first candidates in cost order, enough of them
  //           to survive whatever the jobs before it in the batch take away;
  //  resolve  (driver, a few shared-memory look-ups per job) in job order each job
  //           takes its first free candidates, exactly the reference's pick if
  //           the jobs before it get placed; a taken node is flagged (sm.pend);
  //           a job that could use a taken node at that node's NEW place in the
  //           order ends the batch — it needs that node's updated timeline;
  //  evaluate every (job, node) pair is tested exactly, without touching state,
  //           while the driver takes the picked nodes out of the order; the jobs
  //           before the first failing one are committed;
  //  re-key   every helper puts its node back: at the new cost if its job was
  //           committed, where it was otherwise.
  // The failing job, multi-node backfills and jobs larger than a batch take the
  // one-job path.
  uint32_t ji = 0;
  uint32_t single_job = 0xffffffffu;  // a job already known to need the one-job path
  while (ji < njobs) {
    ensure_issued(ji);
    PROF(1);
    // ---- form the batch: lanes 0..kBatch-1 look at one job each ---------------
    uint32_t nj = 0;
    {
      uint32_t myK = 0, myslot = 0;
      bool okj = false;
      if (lane < (uint32_t)kBatchJobs && lane + 1 < nw && ji + lane < njobs && ji + lane != single_job) {
        const uint32_t j = ji + lane;
        myslot = j % kRing;
        mbar_wait(&s_bar[myslot], (j / kRing) & 1u);
        myK = s_jobs[myslot].node_num;
        okj = myK >= 1 && myK <= mp && myK <= (uint32_t)kBatch;
      }
      uint32_t cum = myK;  // inclusive prefix sum of node_num over the lanes
      for (int o = 1; o < kBatchJobs; o <<= 1) {
        const uint32_t up = __shfl_up_sync(kFullMask, cum, o);
        if ((int)lane >= o) cum += up;
      }
      const unsigned good = __ballot_sync(kFullMask, okj && cum <= (uint32_t)kBatch && cum + 1 <= nw);
      nj = (uint32_t)__ffs((int)~good) - 1u;  // leading run of jobs that fit
      if (lane < nj) { s_bj[lane].slot = myslot; s_bj[lane].K = myK; s_bj[lane].need = cum; }
      __syncwarp();
    }
    bool single = nj == 0;
    if (nj) {
      while (first_bucket + 1 < sm.nb && sm.bcnt[first_bucket] == 0) ++first_bucket;
      if (lane == 0) { s_first_bucket = first_bucket; s_cmd.kind = OP_SELECT; s_cmd.n = nj; s_found = 0; }
      __syncthreads();  // helpers list the candidates
      while (*(volatile uint32_t*)&s_found < nj) spin_pause();
      __syncwarp();     // lists are in
      PROF(3);
      // ---- resolve: 4 lanes per job, 2 list entries per lane -----------------------
      // In job order each job takes its first K free candidates; "free" depends on
      // what the jobs before it took. Solved as a fixed point over all jobs at once:
      // every job recomputes its picks against the others' previous picks; job t is
      // final after round t+1. The first guess — job t skips as many entries as the
      // jobs before it need — is already the answer when the lists coincide.
      // A job with fewer than K pre-filter candidates in all can only be backfilled
      // (taken nodes lose resources, they do not gain candidates), so its list is the
      // capable one from the start.
      const uint32_t rt = lane >> 2, rk = lane & 3u, rg = lane & ~3u;
      const bool ract = rt < nj;
      uint32_t rK = 0, rfirst = 0, rslot = 0, rneed = 0, rmode = 0, rnl = 0;
      uint32_t ce[kEnt];               // my entries of the job's list: entries kEnt*rk + e
#pragma unroll
      for (int e = 0; e < kEnt; ++e) ce[e] = 0xffffu;
      if (ract) {
        const BatchJob bj = s_bj[rt];
        rK = bj.K; rneed = bj.need; rfirst = bj.need - bj.K; rslot = bj.slot;
        rmode = bj.n0 < bj.K ? 1u : 0u;
        rnl = rmode ? bj.n1 : bj.n0;
        const uint16_t* L = rmode ? s_sel[rt].c1 : s_sel[rt].c0;
#pragma unroll
        for (int e = 0; e < kEnt; ++e)
          if (kEnt * rk + e < rnl) ce[e] = L[kEnt * rk + e];
      }
      if (lane < 16) s_pick[0][lane] = 0xffffu;
      __syncwarp();
      if (ract && rnl >= rneed) {  // first guess: entries [rfirst, rfirst + K)
#pragma unroll
        for (int e = 0; e < kEnt; ++e) {
          const uint32_t i = kEnt * rk + e;
          if (i >= rfirst && i < rneed) s_pick[0][i] = (uint16_t)ce[e];
        }
      }
      PROF(4);
      uint32_t cur = 0, rstop = 0;
      bool che[kEnt];                  // my entries are chosen ...
      uint32_t ranke[kEnt];            // ... as the job's ranke-th node
      bool te[kEnt];                   // my entries are taken by an earlier job ...
      uint32_t se[kEnt];               // ... as its task se
      uint32_t chosenm = 0;
#pragma unroll
      for (int e = 0; e < kEnt; ++e) { che[e] = false; ranke[e] = 0; te[e] = false; se[e] = 0; }
      for (uint32_t round = 0; round < (uint32_t)kBatchJobs + 2; ++round) {
        __syncwarp();
        const uint4 pk0 = *reinterpret_cast<const uint4*>(&s_pick[cur][0]);
        const uint4 pk1 = *reinterpret_cast<const uint4*>(&s_pick[cur][8]);
        const uint32_t mine_old = lane < 16 ? (uint32_t)s_pick[cur][lane] : 0u;
        if (lane < 16) s_pick[cur ^ 1u][lane] = 0xffffu;
        __syncwarp();
#pragma unroll
        for (int e = 0; e < kEnt; ++e) { te[e] = false; se[e] = 0; }
#pragma unroll
        for (uint32_t w = 0; w < (uint32_t)kBatch; ++w) {
          const uint32_t wi = w >> 1;
          const uint32_t word = wi == 0 ? pk0.x : wi == 1 ? pk0.y : wi == 2 ? pk0.z : wi == 3 ? pk0.w
                              : wi == 4 ? pk1.x : wi == 5 ? pk1.y : wi == 6 ? pk1.z : pk1.w;
          const uint32_t v = (w & 1u) ? word >> 16 : word & 0xffffu;
          const bool earlier = w < rfirst && v != 0xffffu;  // tasks before mine belong to the jobs before mine
#pragma unroll
          for (int e = 0; e < kEnt; ++e)
            if (earlier && v == ce[e]) { te[e] = true; se[e] = w; }
        }
        // bit i of freem = entry i of the list is free: entry kEnt*k+e sits in lane rg+k, ballot e
        uint32_t freem = 0;
        bool mine_taken = false;
#pragma unroll
        for (int e = 0; e < kEnt; ++e) {
          const uint32_t x = (__ballot_sync(kFullMask, ce[e] != 0xffffu && !te[e]) >> rg) & 0xFu;
          freem |= ((x & 1u) | ((x & 2u) << (kEnt - 1)) | ((x & 4u) << (2 * kEnt - 2)) | ((x & 8u) << (3 * kEnt - 3))) << e;
          mine_taken = mine_taken || te[e];
        }
        const bool any_taken = ((__ballot_sync(kFullMask, mine_taken) >> rg) & 0xFu) != 0;
        rstop = 0;
        if (!ract) rstop = 1;
        else if ((uint32_t)__popc(freem) < rK) rstop = (rmode && !any_taken) ? 2u : 1u;  // too few capable nodes at all : wait for the taken ones
        chosenm = 0;
        if (!rstop) { chosenm = freem; while ((uint32_t)__popc(chosenm) > rK) chosenm &= ~(1u << (31 - __clz((int)chosenm))); }
#pragma unroll
        for (int e = 0; e < kEnt; ++e) {
          const uint32_t i = kEnt * rk + e;
          ranke[e] = (uint32_t)__popc(freem & ((1u << i) - 1u));
          che[e] = (chosenm >> i) & 1u;
          if (che[e]) s_pick[cur ^ 1u][rfirst + ranke[e]] = (uint16_t)ce[e];
        }
        __syncwarp();
        const bool changed = lane < 16 && (uint32_t)s_pick[cur ^ 1u][lane] != mine_old;
        cur ^= 1u;
        PROF_CNT(8, 1);
        if (!__any_sync(kFullMask, changed)) break;
      }
      PROF(5);
      // tasks: node, new cost, job
#pragma unroll
      for (int e = 0; e < kEnt; ++e) {
        if (che[e]) {
          const uint32_t w = rfirst + ranke[e];
          sm.list[w] = (uint16_t)ce[e];
          s_newcost[w] = rmode ? s_sel[rt].nc1[kEnt * rk + e] : s_sel[rt].nc0[kEnt * rk + e];
          s_task[w].slot = rslot; s_task[w].mode = rmode; s_task[w].tfirst = rfirst; s_task[w].job = rt;
        }
      }
      if (lane < (uint32_t)kBatch) s_joblabel[lane] = 0;
      __syncwarp();
      // a taken node listed before my job's last pick that sorts before it at its
      // new cost would be among the first K of the updated order: the job has to
      // wait for that node's update
      bool clash = false;
      {
        const bool live = !rstop && chosenm != 0;
        const uint32_t lastbit = live ? 31u - (uint32_t)__clz((int)chosenm) : 0u;
        uint32_t mine_sel = ce[0];
#pragma unroll
        for (int e = 1; e < kEnt; ++e)
          if (lastbit % (uint32_t)kEnt == (uint32_t)e) mine_sel = ce[e];
        const uint32_t q_last = __shfl_sync(kFullMask, mine_sel, (int)(rg + lastbit / (uint32_t)kEnt));
        if (live) {
          const double c_last = sm.cost[q_last];
#pragma unroll
          for (int e = 0; e < kEnt; ++e)
            if (te[e] && kEnt * rk + e < lastbit) clash = clash || key_lt(s_newcost[se[e]], ce[e], c_last, q_last);
        }
      }
      const unsigned badm = __ballot_sync(kFullMask, rstop != 0 || clash);
      // jobs before the first one that has to wait (lanes of jobs >= nj are "bad")
      const uint32_t njr = badm ? ((uint32_t)__ffs((int)badm) - 1u) >> 2 : nj;
      const uint32_t stop_cut = __shfl_sync(kFullMask, rstop, (int)((njr < 8u ? njr : 0u) * 4u));
      bool need_single = false;
      if (njr < nj && stop_cut == 2u) {
        single_job = ji + njr;  // placing the jobs before it cannot change that
        need_single = njr == 0;
      }
      const uint32_t NT = njr ? __shfl_sync(kFullMask, rneed, (int)((njr - 1u) * 4u)) : 0u;
#pragma unroll
      for (int e = 0; e < kEnt; ++e)
        if (che[e] && rt < njr) sm.pend[ce[e]] = 1;
      __syncwarp();
      PROF(6);
      if (NT) {
        if (lane == 0) { s_cmd.kind = OP_BATCH_P; s_cmd.n = NT; }
        if (lane < NT) s_tbk[lane] = 0xffffffffu;
        __syncthreads();                  // the helpers start evaluating
        bucket_remove_pending(sm, NT);    // meanwhile the picks leave the order
        __syncthreads();                  // verdicts are in
        uint32_t f = NT;
        for (uint32_t i = 0; i < NT; ++i)
          if (!s_ok[i]) { f = i; break; }
        if (f < NT) f = s_task[f].tfirst;
        PROF(9);
        // while the helpers commit, every picked node goes back into the order: at
        // its new cost if its job is placed, where it was otherwise. (The bucket
        // bounds read cpu0/gcnt while a commit may be lowering them: either value
        // is a valid upper bound.)
        // while the helpers commit, every picked node goes back into the order: at
        // its new cost if its job is placed, where it was otherwise. Each helper has
        // looked up its node's bucket; the inserts themselves are done here, one
        // after the other. (The bucket bounds read cpu0/gcnt while a commit may be
        // lowering them: either value is a valid upper bound.)
        bool rebuilt = false;
        for (uint32_t t = 0; t < NT; ++t) {
          const uint32_t q = sm.list[t];
          const double nc = t < f ? s_newcost[t] : sm.cost[q];
          uint32_t tbt;
          while ((tbt = *(volatile uint32_t*)&s_tbk[t]) == 0xffffffffu) spin_pause();  // its helper is still searching
          __syncwarp();
          if (rebuilt || !bucket_place(sm, q, nc, tbt)) leftover_insert(q, nc, rebuilt);
          else if (lane == 0) sm.pend[q] = 0;
        }
        __syncwarp();
        PROF(10);
        __syncthreads();                  // commits are done
        // jobs placed = those whose tasks all lie before the cut
        uint32_t done = 0;
        for (uint32_t t = 0; t < njr; ++t) done += s_bj[t].need <= f ? 1u : 0u;
        // A failed backfill is final: the reference takes exactly these nodes (the
        // first K capable ones) and gives up when they have no common start inside
        // the window (JobScheduler.cpp:5371-5404, 5802). Only a failed immediate
        // start has to continue its walk on the one-job path.
        const bool final_fail = f < NT && s_task[f].mode == 1u;
        if (final_fail) {
          if (lane == 0) a.out.reason[s_jobs[s_task[f].slot].job] = CRANE_REASON_RESOURCE;
          ++done;
        }
        PROF_CNT(13, done);
        PROF_CNT(14, 1);
        PROF(11);
        BUCKET_CHECK("batch", ji);
        ji += done;
        single = f < NT && !final_fail;  // the failing job is next
      } else {
        single = need_single || njr == 0;
      }
    }
    if (single) {
      ensure_issued(ji);
      process_single(ji);
      BUCKET_CHECK("single", ji);
      ++ji;
    }
  }
  // release the helpers
  if (lane == 0) s_cmd.kind = OP_EXIT;
  __syncthreads();
  PROF_FLUSH(a.prof);
}

}  // namespace crane
