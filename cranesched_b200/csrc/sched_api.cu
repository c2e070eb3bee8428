// sched_api.cu — C-ABI of the B200 scheduling hot path (include/crane_sched.h).
//
// Host side of the drop-in boundary behind SchedulerAlgo::NodeSelect
// (reference: src/CraneCtld/JobScheduler.cpp:1141, 5543-5868). This file only
// validates, lays tables out in HBM and launches kernels; every scheduling
// decision is taken on the device (sched_kernels.cuh). There is no CPU path:
// without a CUDA device crane_sched_create() fails with CRANE_ENODEV.
//
// Built two ways:
//   nvcc -gencode arch=compute_100a,code=sm_100a  -> libcrane_sched.so (product)
//   g++ -DCRANE_EMU (tests/cuda_emu)               -> tests/_emu/libcrane_sched_emu.so
//     a kernel-emulation harness used only by the CPU-side unit tests.
#ifdef CRANE_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif

#include <algorithm>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/crane_sched.h"
#include "sched_kernels.cuh"
#include "commit_v2.cuh"
#include "qos_kernels.cuh"

using namespace crane;

namespace {

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 8 + 16;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  // like ensure(), but the first `keep` elements survive a reallocation
  cudaError_t grow(size_t n, size_t keep, cudaStream_t st) {
    if (n <= cap) return cudaSuccess;
    T* old = p;
    size_t want = n + n / 2 + 16;
    T* q = nullptr;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&q), want * sizeof(T));
    if (e != cudaSuccess) return e;
    if (old && keep) {
      e = cudaMemcpyAsync(q, old, keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    }
    if (old) cudaFree(old);
    p = q;
    cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};


}  // namespace

struct crane_sched {
  crane_sched_config_t cfg{};
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[10]{};  // 0-1 upload, 2-7 run / fetch, 8-9 qos filter
  std::string err;
  crane_sched_timing_t timing{};
  size_t v2_budget = 0;       // dynamic shared memory k_commit2 may use
  bool h2d_timed = false;     // ev[0], ev[1] were recorded (crane_sched_upload)
  uint32_t v2_ring = 0;

  // cluster (host copies)
  bool have_cluster = false;
  uint32_t n_nodes = 0, n_parts = 0, n_slots = 0, max_part_slots = 0, words_per_row = 0;
  uint32_t n_vparts = 0, n_resv = 0;  // schedulers = groups of overlapping partitions + reservations
  uint32_t n_comp = 0, max_comp_parts = 1;
  std::vector<uint32_t> h_part_comp, h_part_cidx, h_sched_nparts;
  // host copies of the cluster and reservation tables (the slot layout is rebuilt when either changes)
  std::vector<Row> c_res_total, r_res;
  std::vector<uint8_t> c_alive, c_drain;
  std::vector<uint32_t> c_part_off, c_part_nodes, r_node_off, r_node;
  std::vector<int64_t> r_start, r_end;
  std::vector<uint32_t> h_vslot_off, h_vslot_node;  // per reservation: its nodes (ascending) -> slots h_part_base[n_parts + r] + k
  uint32_t n_gres_entries = 0;
  std::vector<uint32_t> h_part_base, h_slot_node, h_node_slot;
  GresDict dict{};
  uint32_t tl_cap = 0;

  // cluster (device)
  DBuf<uint32_t> d_part_base, d_slot_node, d_node_slot;
  DBuf<Row> d_slot_total;
  // timelines
  DBuf<uint32_t> d_tl_n;
  DBuf<TlEntry> d_tl_ent;
  DBuf<Row> d_avail0, d_class_rows;
  DBuf<uint8_t> d_slot_class;
  DBuf<double> d_cost0;
  DBuf<uint8_t> d_skip;
  DBuf<int64_t> d_first_resv, d_resv_start, d_resv_end;
  DBuf<uint32_t> d_q_off;
  DBuf<uint32_t> d_slot_resv, d_rsv_off, d_rsv_id, d_vpart, d_pd_resv, d_sched_nparts, d_part_comp, d_part_cidx, d_sched_owner;
  DBuf<uint8_t> d_slot_memb;
  DBuf<Row> d_rsv_res;
  bool have_pd_resv = false;

  // pending (device)
  uint32_t n_pending = 0, n_running = 0, n_accounts = 0;
  uint64_t total_alloc = 0;
  bool have_lists_incl = false, have_lists_excl = false, have_mandated = false;
  std::vector<uint32_t> h_alloc_off;
  // device-resident pending table: host mirrors of the small columns, tombstones
  std::vector<uint32_t> h_pd_account, h_incl_off, h_excl_off;
  std::vector<uint8_t> h_dead;
  uint32_t n_dead = 0;
  DBuf<uint8_t> d_dead;
  DBuf<uint32_t> d_erase_rows;
  DBuf<uint32_t> d_ntpn_max, d_ntasks;
  DBuf<uint32_t> d_partition, d_node_num, d_ntpn, d_part_prio, d_qos_prio, d_account, d_alloc_off;
  DBuf<int64_t> d_time_limit, d_submit;
  DBuf<uint8_t> d_exclusive;
  DBuf<double> d_mandated;
  DBuf<View> d_req_node, d_req_task, d_req_total;
  DBuf<uint32_t> d_incl_off, d_incl_nodes, d_excl_off, d_excl_nodes;
  // running (device)
  DBuf<int64_t> d_rn_start, d_rn_end, d_rn_cpu, d_rn_slot_end;
  DBuf<uint32_t> d_rn_node_num, d_rn_part_prio, d_rn_qos_prio, d_rn_account, d_rn_slot_off, d_rn_acc_off, d_rn_acc_job;
  DBuf<uint64_t> d_rn_mem;
  DBuf<Row> d_rn_slot_res;
  DBuf<uint8_t> d_acc_present;
  // work buffers
  DBuf<Bounds> d_bounds;
  DBuf<double> d_acc_service, d_prio;
  DBuf<uint64_t> d_keys_a, d_keys_b;
  DBuf<uint32_t> d_vals_a, d_vals_b, d_hist, d_part_count, d_part_job_off, d_bitmap;
  DBuf<JobQ> d_jobq;
  DBuf<unsigned long long> d_prof;
  uint32_t* d_queue = nullptr;  // points into vals_a / vals_b after the sorts
  // outputs (device)
  DBuf<uint8_t> d_reason;
  DBuf<double> d_out_prio;
  DBuf<int64_t> d_out_start, d_out_end;
  DBuf<uint32_t> d_out_nalloc, d_out_node, d_out_ntasks;
  DBuf<Row> d_out_res;
  bool uploaded = false, ran = false;
  int dict_slot = -1;  // this handle's entry of c_dicts[] on its device
  // one queue over several GPUs
  uint32_t shard_rank = 0, shard_n = 1;
  std::vector<uint32_t> h_part_owner, h_part_list;
  DBuf<uint32_t> d_part_owner, d_part_list;
  // QoS post-filter (R12)
  bool have_qos_cols = false;
  DBuf<uint32_t> d_qos, d_user, d_q_u32, d_q_chain_off, d_q_chain;
  DBuf<int64_t> d_q_i64;
  DBuf<uint8_t> d_q_valid;
  DBuf<crane_tres_limit_t> d_q_tres;
  DBuf<crane_meta_resource_t> d_q_user_usage, d_q_acct_usage, d_q_qos_usage;
};

namespace {

// c_dicts[] slots in use, per device
std::mutex g_slot_mu;
uint32_t g_slots_used[64] = {0};

int fail(crane_sched* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail(h, e__ == 2 ? CRANE_ENOMEM : CRANE_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

template <class T>
int h2d(crane_sched* h, DBuf<T>& dst, const T* src, size_t n) {
  CU(dst.ensure(n ? n : 1));
  if (n) CU(cudaMemcpyAsync(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  return CRANE_OK;
}
#define H2D(buf, src, n)                      \
  do {                                        \
    int rc__ = h2d(h, buf, src, (size_t)(n)); \
    if (rc__ != CRANE_OK) return rc__;        \
  } while (0)

// stable LSD radix sort over bits [0, nbits) of d_keys_a with payload
// d_vals_a; result pointers returned in keys/vals.
int radix_sort(crane_sched* h, uint32_t n, int nbits, uint64_t** keys, uint32_t** vals) {
  uint64_t* ka = h->d_keys_a.p;
  uint64_t* kb = h->d_keys_b.p;
  uint32_t* va = h->d_vals_a.p;
  uint32_t* vb = h->d_vals_b.p;
  uint32_t nblocks = (n + kSortTile - 1) / kSortTile;
  CU(h->d_hist.ensure((size_t)256 * nblocks));
  for (int shift = 0; shift < nbits; shift += 8) {
    CRANE_LAUNCH(k_sort_hist, nblocks, 256, 0, h->stream, ka, n, shift, h->d_hist.p, nblocks);
    CRANE_LAUNCH(k_sort_scan, 1, 1024, 0, h->stream, h->d_hist.p, 256u * nblocks);
    CRANE_LAUNCH(k_sort_scatter, nblocks, 32, 0, h->stream, ka, va, kb, vb, n, shift, h->d_hist.p, nblocks);
    h->timing.kernel_launches += 3;
    std::swap(ka, kb);
    std::swap(va, vb);
  }
  *keys = ka;
  *vals = va;
  return CRANE_OK;
}

}  // namespace

extern "C" {

const char* crane_sched_last_error(const crane_sched_t* h) { return h ? h->err.c_str() : "null handle"; }

int crane_sched_create(const crane_sched_config_t* cfg, int device, crane_sched_t** out) {
  if (!cfg || !out) return CRANE_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return CRANE_ENODEV;
  if (cfg->max_jobs_per_node < 2 || cfg->max_jobs_per_node > 65000) return CRANE_EINVAL;
  if (cfg->cost_policy > 1) return CRANE_ENOSYS;  // 0 MinCpuTimeRatioFirst (JobScheduler.h:40), 1 BestFit (ours, config 4)
  if (device >= 64) return CRANE_EINVAL;
  int slot = -1;
  {
    std::lock_guard<std::mutex> lk(g_slot_mu);
    for (int k = 0; k < kDictSlots; ++k)
      if (!(g_slots_used[device] >> k & 1u)) { slot = k; g_slots_used[device] |= 1u << k; break; }
  }
  if (slot < 0) return CRANE_ENOSYS;  // more than kDictSlots live handles on one device
  crane_sched* h = new crane_sched();
  h->cfg = *cfg;
  h->device = device;
  h->dict_slot = slot;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) {
    { std::lock_guard<std::mutex> lk(g_slot_mu); g_slots_used[device] &= ~(1u << slot); }
    delete h;
    return CRANE_ENODEV;
  }
  for (auto& e : h->ev) cudaEventCreate(&e);
  {
    // k_commit2: everything the SM has beyond the kernel's static tables
    size_t stat = 20 * 1024;
#ifndef CRANE_EMU
    cudaFuncAttributes fa;
    if (cudaFuncGetAttributes(&fa, k_commit2) == cudaSuccess) stat = fa.sharedSizeBytes;
#endif
    const size_t sm_total = 227 * 1024;
    h->v2_budget = sm_total > stat + 1024 ? sm_total - stat - 1024 : 0;
    cudaFuncSetAttribute(k_commit2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->v2_budget);
  }
  h->tl_cap = cfg->max_jobs_per_node + 1;
  *out = h;
  return CRANE_OK;
}

void crane_sched_destroy(crane_sched_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaStreamSynchronize(h->stream);
#define REL(b) h->b.release()
  REL(d_part_base); REL(d_slot_node); REL(d_node_slot); REL(d_slot_total); REL(d_tl_n); REL(d_tl_ent);
  REL(d_avail0); REL(d_class_rows); REL(d_slot_class); REL(d_cost0); REL(d_skip); REL(d_partition);
  REL(d_node_num); REL(d_ntpn); REL(d_part_prio); REL(d_qos_prio); REL(d_account); REL(d_alloc_off);
  REL(d_time_limit); REL(d_submit); REL(d_exclusive); REL(d_mandated); REL(d_req_node); REL(d_req_task);
  REL(d_req_total); REL(d_incl_off); REL(d_incl_nodes); REL(d_excl_off); REL(d_excl_nodes); REL(d_rn_start);
  REL(d_rn_end); REL(d_rn_cpu); REL(d_rn_slot_end); REL(d_rn_node_num); REL(d_rn_part_prio); REL(d_rn_qos_prio);
  REL(d_rn_account); REL(d_rn_slot_off); REL(d_rn_acc_off); REL(d_rn_acc_job); REL(d_rn_mem); REL(d_rn_slot_res);
  REL(d_acc_present); REL(d_bounds); REL(d_acc_service); REL(d_prio); REL(d_keys_a); REL(d_keys_b); REL(d_vals_a);
  REL(d_vals_b); REL(d_hist); REL(d_part_count); REL(d_part_job_off); REL(d_bitmap); REL(d_jobq); REL(d_reason);
  REL(d_out_prio); REL(d_out_start); REL(d_out_end); REL(d_out_nalloc); REL(d_out_node); REL(d_out_ntasks);
  REL(d_out_res); REL(d_prof); REL(d_qos); REL(d_user); REL(d_q_u32); REL(d_q_chain_off); REL(d_q_chain);
  REL(d_q_i64); REL(d_q_valid); REL(d_part_owner); REL(d_part_list); REL(d_ntpn_max); REL(d_ntasks); REL(d_first_resv); REL(d_resv_start); REL(d_resv_end); REL(d_slot_resv);
  REL(d_rsv_off); REL(d_rsv_id); REL(d_dead); REL(d_erase_rows); REL(d_q_off); REL(d_sched_nparts); REL(d_part_comp); REL(d_part_cidx); REL(d_sched_owner); REL(d_slot_memb); REL(d_vpart); REL(d_pd_resv); REL(d_rsv_res); REL(d_q_tres); REL(d_q_user_usage); REL(d_q_acct_usage); REL(d_q_qos_usage);
#undef REL
  for (auto& e : h->ev) cudaEventDestroy(e);
  cudaStreamDestroy(h->stream);
  { std::lock_guard<std::mutex> lk(g_slot_mu); g_slots_used[h->device] &= ~(1u << h->dict_slot); }
  delete h;
}

// Node state a running job's allocation on `node` counts against: the node's own, or —
// for a job inside a reservation — the reservation's state of that node
// (JobScheduler.cpp:5715-5741)
static uint32_t running_slot(const crane_sched* h, const crane_running_t* rn, uint32_t j, uint32_t node) {
  const uint32_t rv = rn->reservation ? rn->reservation[j] : 0xffffffffu;
  if (rv == 0xffffffffu) return h->h_node_slot[node];
  if (rv >= h->n_resv) return 0xffffffffu;  // the reference logs an error and skips the job
  const uint32_t lo = h->r_node_off[rv], hi = h->r_node_off[rv + 1];
  const auto it = std::lower_bound(h->r_node.begin() + lo, h->r_node.begin() + hi, node);
  if (it == h->r_node.begin() + hi || *it != node) return 0xffffffffu;
  return h->h_part_base[h->n_comp + rv] + (uint32_t)(it - (h->r_node.begin() + lo));
}

// Slot layout of the node states: the usable nodes of every partition (contiguous per
// partition), then the nodes of every reservation; device copies of the cluster tables.
static int build_layout(crane_sched* h) {
  h->h_node_slot.assign(h->n_nodes, 0xffffffffu);
  h->h_slot_node.clear();
  std::vector<Row> slot_total;
  uint32_t max_mp = 0;
  // Partitions that share a node share its NodeState (JobScheduler.cpp:5622): their job
  // loops are coupled through its timeline, so one scheduler (one CTA) takes every
  // connected component of overlapping partitions; each partition keeps its own order.
  std::vector<uint32_t> uf(h->n_parts);
  for (uint32_t p = 0; p < h->n_parts; ++p) uf[p] = p;
  auto find = [&](uint32_t x) { while (uf[x] != x) { uf[x] = uf[uf[x]]; x = uf[x]; } return x; };
  {
    std::vector<uint32_t> first(h->n_nodes, 0xffffffffu);
    for (uint32_t p = 0; p < h->n_parts; ++p) {
      uint32_t prev = 0;
      for (uint32_t k = h->c_part_off[p]; k < h->c_part_off[p + 1]; ++k) {
        const uint32_t n = h->c_part_nodes[k];
        if (n >= h->n_nodes) return fail(h, CRANE_EINVAL, "cluster: node index %u out of range", n);
        if (k > h->c_part_off[p] && n <= prev) return fail(h, CRANE_EINVAL, "cluster: partition %u node list must be ascending", p);
        prev = n;
        if (first[n] == 0xffffffffu) first[n] = p;
        else { const uint32_t a = find(first[n]), c = find(p); if (a != c) uf[std::max(a, c)] = std::min(a, c); }
      }
    }
  }
  h->h_part_comp.assign(h->n_parts, 0);
  h->h_part_cidx.assign(h->n_parts, 0);
  h->h_sched_nparts.clear();
  std::vector<std::vector<uint32_t>> comp_parts;
  {
    std::vector<uint32_t> id(h->n_parts, 0xffffffffu);
    for (uint32_t p = 0; p < h->n_parts; ++p) {
      const uint32_t r = find(p);
      if (id[r] == 0xffffffffu) { id[r] = (uint32_t)comp_parts.size(); comp_parts.emplace_back(); }
      h->h_part_comp[p] = id[r];
      h->h_part_cidx[p] = (uint32_t)comp_parts[id[r]].size();
      comp_parts[id[r]].push_back(p);
    }
  }
  h->n_comp = (uint32_t)comp_parts.size();
  h->h_part_base.assign(h->n_comp + 1, 0);
  h->h_part_base.reserve(h->n_comp + h->n_resv + 1);
  std::vector<uint8_t> slot_memb;
  uint32_t max_cp = 1;
  for (uint32_t c = 0; c < h->n_comp; ++c) {
    h->h_part_base[c] = (uint32_t)h->h_slot_node.size();
    if (comp_parts[c].size() > (size_t)kMaxCompParts)
      return fail(h, CRANE_ENOSYS, "cluster: %zu partitions overlap in one connected group (at most %d)", comp_parts[c].size(), kMaxCompParts);
    max_cp = std::max<uint32_t>(max_cp, (uint32_t)comp_parts[c].size());
    h->h_sched_nparts.push_back((uint32_t)comp_parts[c].size());
    std::vector<std::pair<uint32_t, uint8_t>> nodes;  // (node, membership bits), ascending node index
    for (uint32_t i = 0; i < comp_parts[c].size(); ++i) {
      const uint32_t p = comp_parts[c][i];
      for (uint32_t k = h->c_part_off[p]; k < h->c_part_off[p + 1]; ++k) nodes.push_back({h->c_part_nodes[k], (uint8_t)(1u << i)});
    }
    std::sort(nodes.begin(), nodes.end());
    for (size_t i = 0; i < nodes.size();) {
      const uint32_t n = nodes[i].first;
      uint8_t bits = 0;
      for (; i < nodes.size() && nodes[i].first == n; ++i) bits |= nodes[i].second;
      if (!h->c_alive[n] || h->c_drain[n]) continue;  // JobScheduler.cpp:5629
      const Row& t = h->c_res_total[n];
      if (t.cpu_raw < 0 || t.cpu_raw > (int64_t)1 << 40) return fail(h, CRANE_EINVAL, "cluster: node %u cpu out of range", n);
      for (uint32_t e = h->n_gres_entries; e < CRANE_GRES_ENTRIES; ++e)
        if (field16(t.g, e)) return fail(h, CRANE_EINVAL, "cluster: node %u has slots for an undeclared gres entry", n);
      h->h_node_slot[n] = (uint32_t)h->h_slot_node.size();
      h->h_slot_node.push_back(n);
      slot_total.push_back(t);
      slot_memb.push_back(bits);
    }
    max_mp = std::max(max_mp, (uint32_t)h->h_slot_node.size() - h->h_part_base[c]);
  }
  h->h_part_base[h->n_comp] = (uint32_t)h->h_slot_node.size();
  h->max_comp_parts = max_cp;
  const uint32_t n_phys = (uint32_t)h->h_slot_node.size();
  // one more scheduler per reservation: node states holding exactly the reserved
  // resources (JobScheduler.cpp:5689-5703), whatever the node's own state
  h->n_vparts = h->n_comp + h->n_resv;
  std::vector<uint32_t> slot_resv(n_phys, 0xffffffffu);
  for (uint32_t r = 0; r < h->n_resv; ++r) {
    for (uint32_t k = h->r_node_off[r]; k < h->r_node_off[r + 1]; ++k) {
      h->h_slot_node.push_back(h->r_node[k]);
      slot_total.push_back(h->r_res[k]);
      slot_resv.push_back(r);
      slot_memb.push_back(1);
    }
    h->h_sched_nparts.push_back(1);
    h->h_part_base.push_back((uint32_t)h->h_slot_node.size());
    max_mp = std::max(max_mp, h->r_node_off[r + 1] - h->r_node_off[r]);
  }
  // reservations holding resources of a partition's node, ascending reservation id
  std::vector<uint32_t> rsv_off((size_t)h->h_slot_node.size() + 1, 0), rsv_id;
  std::vector<Row> rsv_res;
  {
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> per(n_phys);
    for (uint32_t r = 0; r < h->n_resv; ++r)
      for (uint32_t k = h->r_node_off[r]; k < h->r_node_off[r + 1]; ++k) {
        const uint32_t g = h->h_node_slot[h->r_node[k]];
        if (g != 0xffffffffu) per[g].push_back({r, k});  // (nodes outside node_state_map are skipped, :5680)
      }
    for (uint32_t g = 0; g < n_phys; ++g) {
      for (auto& e : per[g]) { rsv_id.push_back(e.first); rsv_res.push_back(h->r_res[e.second]); }
      rsv_off[g + 1] = (uint32_t)rsv_id.size();
    }
    for (size_t g = n_phys; g < h->h_slot_node.size(); ++g) rsv_off[g + 1] = (uint32_t)rsv_id.size();
  }
  h->n_slots = (uint32_t)h->h_slot_node.size();
  h->max_part_slots = max_mp;
  h->words_per_row = std::max<uint32_t>(4, ((max_mp + 31) / 32 + 3) / 4 * 4);  // 16-byte rows for the bulk copies
  h->v2_ring = commit2_ring_slots(max_mp, h->words_per_row, h->n_gres_entries > 0, h->v2_budget, h->max_comp_parts);
  if (max_mp > 65000 || h->v2_ring == 0)
    return fail(h, CRANE_ENOSYS, "cluster: partition with %u usable nodes exceeds the per-SM state budget", max_mp);
  // res_total classes per partition (distinct rows), cached in shared memory by the commit kernel
  std::vector<uint8_t> slot_class(std::max<size_t>(slot_total.size(), 1), 0xff);
  std::vector<Row> class_rows((size_t)std::max<uint32_t>(h->n_vparts, 1) * kMaxClasses);
  memset(class_rows.data(), 0, class_rows.size() * sizeof(Row));
  for (uint32_t p = 0; p < h->n_vparts; ++p) {
    uint32_t ncls = 0;
    for (uint32_t g = h->h_part_base[p]; g < h->h_part_base[p + 1]; ++g) {
      uint32_t k = 0;
      for (; k < ncls; ++k)
        if (memcmp(&class_rows[(size_t)p * kMaxClasses + k], &slot_total[g], sizeof(Row)) == 0) break;
      if (k == ncls) {
        if (ncls == (uint32_t)kMaxClasses) continue;  // no class: the kernel reads slot_total
        class_rows[(size_t)p * kMaxClasses + ncls++] = slot_total[g];
      }
      slot_class[g] = (uint8_t)k;
    }
  }

  H2D(h->d_part_base, h->h_part_base.data(), h->h_part_base.size());
  H2D(h->d_slot_node, h->h_slot_node.data(), h->h_slot_node.size());
  H2D(h->d_node_slot, h->h_node_slot.data(), h->h_node_slot.size());
  H2D(h->d_slot_total, slot_total.data(), slot_total.size());
  H2D(h->d_slot_class, slot_class.data(), slot_class.size());
  H2D(h->d_class_rows, class_rows.data(), class_rows.size());
  H2D(h->d_slot_resv, slot_resv.data(), slot_resv.size());
  H2D(h->d_slot_memb, slot_memb.data(), slot_memb.size());
  H2D(h->d_sched_nparts, h->h_sched_nparts.data(), h->h_sched_nparts.size());
  H2D(h->d_part_comp, h->h_part_comp.data(), h->h_part_comp.size());
  H2D(h->d_part_cidx, h->h_part_cidx.data(), h->h_part_cidx.size());
  H2D(h->d_rsv_off, rsv_off.data(), rsv_off.size());
  H2D(h->d_rsv_id, rsv_id.data(), rsv_id.size());
  H2D(h->d_rsv_res, rsv_res.data(), rsv_res.size());
  H2D(h->d_resv_start, h->r_start.data(), h->r_start.size());
  H2D(h->d_resv_end, h->r_end.data(), h->r_end.size());
  CU(cudaMemcpyToSymbolAsync(c_dicts, &h->dict, sizeof(GresDict), (size_t)h->dict_slot * sizeof(GresDict), cudaMemcpyHostToDevice, h->stream));
  size_t ns = std::max<uint32_t>(h->n_slots, 1);
  CU(h->d_tl_n.ensure(ns));
  CU(h->d_tl_ent.ensure(ns * h->tl_cap));
  CU(h->d_avail0.ensure(ns));
  CU(h->d_cost0.ensure(ns));
  CU(h->d_skip.ensure(ns));
  CU(h->d_first_resv.ensure(ns));
  CU(h->d_part_count.ensure(h->n_vparts + 2));
  CU(h->d_part_job_off.ensure(h->n_vparts + 2));
  CU(cudaStreamSynchronize(h->stream));
  h->have_cluster = true;
  h->uploaded = false;
  h->shard_rank = 0;
  h->shard_n = 1;  // a new cluster: the partition -> rank table has to be set again
  return CRANE_OK;
}


int crane_sched_set_cluster(crane_sched_t* h, const crane_cluster_t* c) {
  if (!h || !c) return CRANE_EINVAL;
  if (!c->res_total || !c->alive || !c->drain || !c->part_off || (!c->part_nodes && c->n_nodes))
    return fail(h, CRANE_EINVAL, "cluster: null table");
  if (c->n_gres_entries > CRANE_GRES_ENTRIES) return fail(h, CRANE_EINVAL, "cluster: too many gres entries");
  for (uint32_t e = 0; e < c->n_gres_entries; ++e) {
    if (c->gres_entry_name[e] >= CRANE_GRES_NAMES) return fail(h, CRANE_EINVAL, "cluster: gres name id out of range");
    if (e && c->gres_entry_name[e] < c->gres_entry_name[e - 1])
      return fail(h, CRANE_EINVAL, "cluster: gres entries must be grouped by ascending name id");
  }
  CU(cudaSetDevice(h->device));
  h->have_cluster = false;  // state of the previous cluster / tick is void from here on
  h->uploaded = false;
  h->ran = false;
  if (c->n_partitions && c->part_off[0] != 0) return fail(h, CRANE_EINVAL, "cluster: part_off[0] != 0");
  for (uint32_t p = 0; p < c->n_partitions; ++p)
    if (c->part_off[p + 1] < c->part_off[p]) return fail(h, CRANE_EINVAL, "cluster: part_off is not monotonic");
  h->n_nodes = c->n_nodes;
  h->n_parts = c->n_partitions;
  memset(&h->dict, 0, sizeof h->dict);
  h->dict.n_entries = c->n_gres_entries;
  for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) h->dict.entry_name[e] = c->gres_entry_name[e];
  for (uint32_t e = 0; e < c->n_gres_entries; ++e) {
    const uint8_t g = c->gres_entry_name[e];
    if (h->dict.name_count[g] == 0) h->dict.name_first[g] = (uint8_t)e;
    h->dict.name_count[g]++;
    h->dict.name_mask8[g] |= 0xFFull << (8 * e);
  }
  h->n_gres_entries = c->n_gres_entries;
  h->c_res_total.assign(reinterpret_cast<const Row*>(c->res_total), reinterpret_cast<const Row*>(c->res_total) + c->n_nodes);
  h->c_alive.assign(c->alive, c->alive + c->n_nodes);
  h->c_drain.assign(c->drain, c->drain + c->n_nodes);
  h->c_part_off.assign(c->part_off, c->part_off + c->n_partitions + 1);
  h->c_part_nodes.assign(c->part_nodes, c->part_nodes + (c->n_partitions ? c->part_off[c->n_partitions] : 0));
  h->r_start.clear(); h->r_end.clear(); h->r_node_off.assign(1, 0); h->r_node.clear(); h->r_res.clear();  // a new cluster has no reservations yet
  h->n_resv = 0;
  return build_layout(h);
}

int crane_sched_set_reservations(crane_sched_t* h, const crane_reservations_t* rv) {
  if (!h) return CRANE_EINVAL;
  if (!h->have_cluster) return fail(h, CRANE_EINVAL, "set_reservations: set_cluster first");
  CU(cudaSetDevice(h->device));
  const uint32_t n = rv ? rv->n : 0;
  if (n && (!rv->start_time || !rv->end_time || !rv->node_off)) return fail(h, CRANE_EINVAL, "reservations: null table");
  if (n && rv->node_off[0] != 0) return fail(h, CRANE_EINVAL, "reservations: node_off[0] != 0");
  for (uint32_t r = 0; r < n; ++r) {
    if (rv->node_off[r + 1] < rv->node_off[r]) return fail(h, CRANE_EINVAL, "reservations: node_off is not monotonic");
    if (rv->end_time[r] < rv->start_time[r]) return fail(h, CRANE_EINVAL, "reservations[%u]: end before start", r);
    for (uint32_t k = rv->node_off[r]; k < rv->node_off[r + 1]; ++k) {
      if (!rv->node || !rv->res) return fail(h, CRANE_EINVAL, "reservations: null node table");
      if (rv->node[k] >= h->n_nodes) return fail(h, CRANE_EINVAL, "reservations[%u]: node out of range", r);
      for (uint32_t k2 = rv->node_off[r]; k2 < k; ++k2)
        if (rv->node[k2] == rv->node[k]) return fail(h, CRANE_EINVAL, "reservations[%u]: node %u listed twice", r, rv->node[k]);
    }
  }
  h->have_cluster = false;
  h->uploaded = false;
  h->ran = false;
  h->n_resv = n;
  h->r_start.assign(rv ? rv->start_time : nullptr, rv ? rv->start_time + n : nullptr);
  h->r_end.assign(rv ? rv->end_time : nullptr, rv ? rv->end_time + n : nullptr);
  h->r_node_off.assign(1, 0);
  h->r_node.clear();
  h->r_res.clear();
  for (uint32_t r = 0; r < n; ++r) {
    // node states in node-index order: equal-cost nodes are ordered by index (deviation D1)
    std::vector<uint32_t> idx(rv->node_off[r + 1] - rv->node_off[r]);
    for (uint32_t k = 0; k < idx.size(); ++k) idx[k] = rv->node_off[r] + k;
    std::sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return rv->node[x] < rv->node[y]; });
    for (uint32_t k : idx) {
      h->r_node.push_back(rv->node[k]);
      h->r_res.push_back(reinterpret_cast<const Row&>(rv->res[k]));
    }
    h->r_node_off.push_back((uint32_t)h->r_node.size());
  }
  return build_layout(h);
}

int crane_sched_set_shard(crane_sched_t* h, uint32_t rank, uint32_t n_ranks, const uint32_t* part_owner) {
  if (!h) return CRANE_EINVAL;
  if (!h->have_cluster) return fail(h, CRANE_EINVAL, "set_shard: set_cluster first");
  if (n_ranks == 0 || rank >= n_ranks) return fail(h, CRANE_EINVAL, "set_shard: rank %u of %u", rank, n_ranks);
  CU(cudaSetDevice(h->device));
  h->shard_rank = rank;
  h->shard_n = n_ranks;
  h->h_part_owner.clear();
  h->h_part_list.clear();
  if (n_ranks > 1) {
    if (!part_owner && h->n_parts) return fail(h, CRANE_EINVAL, "set_shard: null owner table");
    std::vector<uint32_t> sched_owner(h->n_vparts, 0);  // reservations: rank 0
    for (uint32_t p = 0; p < h->n_parts; ++p) {
      if (part_owner[p] >= n_ranks) return fail(h, CRANE_EINVAL, "set_shard: partition %u owned by rank %u of %u", p, part_owner[p], n_ranks);
      h->h_part_owner.push_back(part_owner[p]);
      const uint32_t c = h->h_part_comp[p];
      if (h->h_part_cidx[p] == 0) sched_owner[c] = part_owner[p];
      else if (sched_owner[c] != part_owner[p])
        return fail(h, CRANE_EINVAL, "set_shard: partitions that share nodes must have one owner (partition %u)", p);
    }
    for (uint32_t c = 0; c < h->n_vparts; ++c)
      if (sched_owner[c] == rank) h->h_part_list.push_back(c);
    H2D(h->d_sched_owner, sched_owner.data(), sched_owner.size());
    H2D(h->d_part_owner, h->h_part_owner.data(), h->h_part_owner.size());
    H2D(h->d_part_list, h->h_part_list.data(), h->h_part_list.size());
    CU(cudaStreamSynchronize(h->stream));
  }
  return CRANE_OK;
}

int crane_sched_device_placements(crane_sched_t* h, crane_device_placements_t* out) {
  if (!h || !out) return CRANE_EINVAL;
  if (!h->ran) return fail(h, CRANE_EINVAL, "device_placements: run first");
  out->reason = h->d_reason.p;
  out->start_time = h->d_out_start.p;
  out->end_time = h->d_out_end.p;
  out->n_alloc = h->d_out_nalloc.p;
  out->alloc_node = h->d_out_node.p;
  out->alloc_ntasks = h->d_out_ntasks.p;
  out->alloc_res = h->d_out_res.p;
  out->n_jobs = h->n_pending;
  out->n_rows = h->total_alloc;
  return CRANE_OK;
}

}  // extern "C" (the split upload lives in an unnamed namespace)

namespace {

template <class T>
int h2d_at(crane_sched* h, DBuf<T>& dst, const T* src, size_t base, size_t n) {
  CU(dst.grow(std::max<size_t>(base + n, 1), base, h->stream));
  if (n) CU(cudaMemcpyAsync(dst.p + base, src, n * sizeof(T), cudaMemcpyHostToDevice, h->stream));
  return CRANE_OK;
}
#define H2D_AT(buf, src, base, n)                                  \
  do {                                                             \
    int rc__ = h2d_at(h, buf, src, (size_t)(base), (size_t)(n));   \
    if (rc__ != CRANE_OK) return rc__;                             \
  } while (0)

// Appends the rows of `pd` to the device-resident pending table (rows keep job-id
// order: the reference's pending map is a btree over job ids, JobScheduler.cpp:1092).
int pending_append(crane_sched* h, const crane_pending_t* pd) {
  const uint32_t base = h->n_pending, N = pd->n;
  if (N && (!pd->partition || !pd->time_limit || !pd->submit_time || !pd->node_num || !pd->ntasks ||
            !pd->ntasks_per_node_min || !pd->ntasks_per_node_max || !pd->exclusive || !pd->partition_priority ||
            !pd->qos_priority || !pd->account || !pd->req_node || !pd->req_task || !pd->req_total))
    return fail(h, CRANE_EINVAL, "pending: null column");
  if ((uint64_t)base + N > 0xfffffff0ull) return fail(h, CRANE_EINVAL, "pending: too many rows");
  // optional columns are all-or-nothing over the life of the table
  if (base) {
    if ((pd->qos != nullptr && pd->user != nullptr) != h->have_qos_cols || (pd->mandated_priority != nullptr) != h->have_mandated ||
        (pd->reservation != nullptr) != h->have_pd_resv || (pd->incl_off != nullptr) != h->have_lists_incl ||
        (pd->excl_off != nullptr) != h->have_lists_excl)
      return fail(h, CRANE_EINVAL, "pending_append: optional columns (qos/user, mandated_priority, reservation, node lists) must match the table's");
  } else {
    h->have_qos_cols = pd->qos != nullptr && pd->user != nullptr;
    h->have_mandated = pd->mandated_priority != nullptr;
    h->have_pd_resv = pd->reservation != nullptr;
    h->have_lists_incl = pd->incl_off != nullptr;
    h->have_lists_excl = pd->excl_off != nullptr;
    h->h_alloc_off.assign(1, 0);
    h->h_incl_off.assign(1, 0);
    h->h_excl_off.assign(1, 0);
    h->h_pd_account.clear();
    h->h_dead.clear();
    h->n_dead = 0;
  }
  // ---- validation + alloc_off (prefix sum of node_num over all rows) -----------
  uint64_t acc = h->h_alloc_off.back();
  for (uint32_t i = 0; i < N; ++i) {
    if (pd->node_num[i] == 0) return fail(h, CRANE_EINVAL, "pending[%u]: node_num == 0", i);
    if (pd->time_limit[i] < 1) return fail(h, CRANE_EINVAL, "pending[%u]: time_limit < 1", i);
    const uint32_t t = pd->ntasks_per_node_min[i], tmax = pd->ntasks_per_node_max[i];
    if (t == 0 || tmax < t) return fail(h, CRANE_EINVAL, "pending[%u]: ntasks_per_node_min/max", i);
    // general task distribution (the top-K heaps of JobScheduler.cpp:5193-5205): heaps of up to 127 nodes
    if ((tmax != t || (uint64_t)t * pd->node_num[i] != pd->ntasks[i]) && pd->node_num[i] >= (uint32_t)kHeapMax)
      return fail(h, CRANE_ENOSYS, "pending[%u]: a job with a ntasks_per_node range wider than %d nodes", i, kHeapMax - 1);
    if ((uint64_t)tmax * pd->node_num[i] < pd->ntasks[i] || (uint64_t)t * pd->node_num[i] > pd->ntasks[i])
      return fail(h, CRANE_EINVAL, "pending[%u]: ntasks outside [node_num*ntpn_min, node_num*ntpn_max]", i);
    if (pd->req_node[i].cpu_raw < 0 || pd->req_task[i].cpu_raw < 0)
      return fail(h, CRANE_EINVAL, "pending[%u]: negative cpu request", i);
    if (pd->account[i] > (1u << 24)) return fail(h, CRANE_EINVAL, "account ids must be dense (< 2^24)");
    acc += pd->node_num[i];
    if (acc > 0xfffffff0ull) return fail(h, CRANE_EINVAL, "pending: sum(node_num) overflows");
  }
  if ((pd->incl_off && !pd->incl_nodes && pd->incl_off[N]) || (pd->excl_off && !pd->excl_nodes && pd->excl_off[N]))
    return fail(h, CRANE_EINVAL, "pending: node list CSR without node array");
  for (uint32_t i = 0; i < N; ++i)
    if ((pd->incl_off && pd->incl_off[i + 1] < pd->incl_off[i]) || (pd->excl_off && pd->excl_off[i + 1] < pd->excl_off[i]))
      return fail(h, CRANE_EINVAL, "pending[%u]: node list offsets are not monotonic", i);
  if (h->have_lists_incl)
    for (uint32_t k = 0; k < pd->incl_off[N]; ++k)
      if (pd->incl_nodes[k] >= h->n_nodes && pd->incl_nodes[k] != 0xFFFFFFFFu)  // 0xFFFFFFFF = a host the cluster does not know
        return fail(h, CRANE_EINVAL, "pending: included node out of range");
  // ---- columns ----------------------------------------------------------------
  H2D_AT(h->d_partition, pd->partition, base, N);
  H2D_AT(h->d_time_limit, pd->time_limit, base, N);
  H2D_AT(h->d_submit, pd->submit_time, base, N);
  H2D_AT(h->d_node_num, pd->node_num, base, N);
  H2D_AT(h->d_ntpn, pd->ntasks_per_node_min, base, N);
  H2D_AT(h->d_ntpn_max, pd->ntasks_per_node_max, base, N);
  H2D_AT(h->d_ntasks, pd->ntasks, base, N);
  H2D_AT(h->d_exclusive, pd->exclusive, base, N);
  H2D_AT(h->d_part_prio, pd->partition_priority, base, N);
  H2D_AT(h->d_qos_prio, pd->qos_priority, base, N);
  H2D_AT(h->d_account, pd->account, base, N);
  if (h->have_qos_cols) {
    H2D_AT(h->d_qos, pd->qos, base, N);
    H2D_AT(h->d_user, pd->user, base, N);
  }
  if (h->have_mandated) H2D_AT(h->d_mandated, pd->mandated_priority, base, N);
  H2D_AT(h->d_req_node, reinterpret_cast<const View*>(pd->req_node), base, N);
  H2D_AT(h->d_req_task, reinterpret_cast<const View*>(pd->req_task), base, N);
  H2D_AT(h->d_req_total, reinterpret_cast<const View*>(pd->req_total), base, N);
  if (h->have_pd_resv) H2D_AT(h->d_pd_resv, pd->reservation, base, N);
  {
    std::vector<uint8_t> zeros(N, 0);
    H2D_AT(h->d_dead, zeros.data(), base, N);
    CU(cudaStreamSynchronize(h->stream));
  }
  // host mirrors
  for (uint32_t i = 0; i < N; ++i) {
    h->h_alloc_off.push_back(h->h_alloc_off.back() + pd->node_num[i]);
    h->h_pd_account.push_back(pd->account[i]);
    h->h_dead.push_back(0);
  }
  H2D_AT(h->d_alloc_off, h->h_alloc_off.data() + base, base, (size_t)N + 1);
  if (h->have_lists_incl) {
    const uint32_t nb = h->h_incl_off.back();
    for (uint32_t i = 0; i < N; ++i) h->h_incl_off.push_back(nb + pd->incl_off[i + 1]);
    H2D_AT(h->d_incl_off, h->h_incl_off.data() + base, base, (size_t)N + 1);
    H2D_AT(h->d_incl_nodes, pd->incl_nodes, nb, pd->incl_off[N]);
  }
  if (h->have_lists_excl) {
    const uint32_t nb = h->h_excl_off.back();
    for (uint32_t i = 0; i < N; ++i) h->h_excl_off.push_back(nb + pd->excl_off[i + 1]);
    H2D_AT(h->d_excl_off, h->h_excl_off.data() + base, base, (size_t)N + 1);
    H2D_AT(h->d_excl_nodes, pd->excl_nodes, nb, pd->excl_off[N]);
  }
  CU(cudaStreamSynchronize(h->stream));
  h->n_pending = base + N;
  h->total_alloc = h->h_alloc_off.back();
  return CRANE_OK;
}

// The running table of this tick + the work and output buffers for the resident pending table.
int set_running(crane_sched* h, const crane_running_t* rn) {
  const uint32_t N = h->n_pending;
  const uint32_t R = rn ? rn->n : 0;
  if (R && (!rn->start_time || !rn->end_time || !rn->node_num || !rn->partition_priority || !rn->qos_priority ||
            !rn->account || !rn->view_cpu_raw || !rn->view_mem || !rn->alloc_off))
    return fail(h, CRANE_EINVAL, "running: null column");
  for (uint32_t j = 0; j < R; ++j)
    if (rn->alloc_off[j + 1] < rn->alloc_off[j]) return fail(h, CRANE_EINVAL, "running: alloc_off is not monotonic");
  uint32_t max_account = 0;
  for (uint32_t i = 0; i < N; ++i) max_account = std::max(max_account, h->h_pd_account[i]);
  for (uint32_t k = 0; k < R; ++k) max_account = std::max(max_account, rn->account[k]);
  if (max_account > (1u << 24)) return fail(h, CRANE_EINVAL, "account ids must be dense (< 2^24)");
  const uint32_t A = (N + R) ? max_account + 1 : 0;

  // ---- running jobs: columns + regrouping by node slot and by account ------
  // (flattening of RnJobInScheduler::allocated_res; order inside a node /
  // account is input order, deviation D6)
  std::vector<uint32_t> slot_off((size_t)h->n_slots + 1, 0), acc_off((size_t)A + 1, 0), acc_job(R);
  std::vector<uint8_t> acc_present(std::max<uint32_t>(A, 1), 0);
  std::vector<int64_t> slot_end;
  std::vector<Row> slot_res;
  for (uint32_t i = 0; i < N; ++i)
    if (!h->h_dead[i]) acc_present[h->h_pd_account[i]] = 1;
  if (R) {
    const uint32_t E = rn->alloc_off[R];
    if (E && (!rn->alloc_node || !rn->alloc_res)) return fail(h, CRANE_EINVAL, "running: null allocation table");
    for (uint32_t j = 0; j < R; ++j) {
      acc_present[rn->account[j]] = 1;
      acc_off[rn->account[j] + 1]++;
      for (uint32_t k = rn->alloc_off[j]; k < rn->alloc_off[j + 1]; ++k) {
        uint32_t n = rn->alloc_node[k];
        if (n >= h->n_nodes) return fail(h, CRANE_EINVAL, "running[%u]: node out of range", j);
        uint32_t g = running_slot(h, rn, j, n);
        if (g != 0xffffffffu) slot_off[g + 1]++;  // nodes outside node_state_map are ignored (JS.cpp:5719)
      }
    }
    for (uint32_t g = 0; g < h->n_slots; ++g) slot_off[g + 1] += slot_off[g];
    for (uint32_t a = 0; a < A; ++a) acc_off[a + 1] += acc_off[a];
    slot_end.resize(slot_off[h->n_slots]);
    slot_res.resize(slot_off[h->n_slots]);
    std::vector<uint32_t> fill_s(slot_off.begin(), slot_off.end() - 1), fill_a(acc_off.begin(), acc_off.end() - 1);
    for (uint32_t j = 0; j < R; ++j) {
      acc_job[fill_a[rn->account[j]]++] = j;
      for (uint32_t k = rn->alloc_off[j]; k < rn->alloc_off[j + 1]; ++k) {
        uint32_t g = running_slot(h, rn, j, rn->alloc_node[k]);
        if (g == 0xffffffffu) continue;
        uint32_t dst = fill_s[g]++;
        slot_end[dst] = rn->end_time[j];
        slot_res[dst] = reinterpret_cast<const Row&>(rn->alloc_res[k]);
      }
    }
    H2D(h->d_rn_start, rn->start_time, R);
    H2D(h->d_rn_end, rn->end_time, R);
    H2D(h->d_rn_node_num, rn->node_num, R);
    H2D(h->d_rn_part_prio, rn->partition_priority, R);
    H2D(h->d_rn_qos_prio, rn->qos_priority, R);
    H2D(h->d_rn_account, rn->account, R);
    H2D(h->d_rn_cpu, rn->view_cpu_raw, R);
    H2D(h->d_rn_mem, rn->view_mem, R);
    H2D(h->d_rn_slot_end, slot_end.data(), slot_end.size());
    H2D(h->d_rn_slot_res, slot_res.data(), slot_res.size());
  }
  // always present: k_service / k_node_init index these even with no running job
  H2D(h->d_rn_acc_off, acc_off.data(), acc_off.size());
  H2D(h->d_rn_acc_job, acc_job.data(), acc_job.size());
  H2D(h->d_rn_slot_off, slot_off.data(), slot_off.size());
  H2D(h->d_acc_present, acc_present.data(), acc_present.size());
  // staging vectors above are pageable: the copies have completed on return,
  // but make it explicit before they go out of scope
  CU(cudaStreamSynchronize(h->stream));
  h->n_running = R;
  h->n_accounts = A;

  // ---- work + output buffers ----------------------------------------------
  size_t n1 = std::max<uint32_t>(N, 1);
  CU(h->d_bounds.ensure(1));
  CU(h->d_acc_service.ensure(std::max<uint32_t>(A, 1)));
  CU(h->d_prio.ensure(n1));
  CU(h->d_keys_a.ensure(n1));
  CU(h->d_keys_b.ensure(n1));
  CU(h->d_vals_a.ensure(n1));
  CU(h->d_vals_b.ensure(n1));
  uint32_t nq = std::min<uint32_t>(N, h->cfg.scheduled_batch_size);
  CU(h->d_jobq.ensure(std::max<uint32_t>(nq, 1)));
  CU(h->d_bitmap.ensure((size_t)std::max<uint32_t>(nq, 1) * h->words_per_row));
  CU(h->d_reason.ensure(n1));
  CU(h->d_out_prio.ensure(n1));
  CU(h->d_out_start.ensure(n1));
  CU(h->d_out_end.ensure(n1));
  CU(h->d_out_nalloc.ensure(n1));
  size_t ta = std::max<uint64_t>(h->total_alloc, 1);
  CU(h->d_out_node.ensure(ta));
  CU(h->d_out_ntasks.ensure(ta));
  CU(h->d_out_res.ensure(ta));
  return CRANE_OK;
}

}  // namespace

extern "C" {

uint32_t crane_sched_pending_rows(const crane_sched_t* h) { return h ? h->n_pending : 0; }

int crane_sched_pending_reset(crane_sched_t* h) {
  if (!h) return CRANE_EINVAL;
  h->n_pending = 0;
  h->total_alloc = 0;
  h->n_dead = 0;
  h->h_alloc_off.assign(1, 0);
  h->h_pd_account.clear();
  h->h_dead.clear();
  h->uploaded = false;
  h->ran = false;
  return CRANE_OK;
}

int crane_sched_pending_append(crane_sched_t* h, const crane_pending_t* rows, uint32_t* first_row) {
  if (!h || !rows) return CRANE_EINVAL;
  if (!h->have_cluster) return fail(h, CRANE_EINVAL, "pending_append: set_cluster first");
  CU(cudaSetDevice(h->device));
  h->uploaded = false;  // the running table and the work buffers follow with crane_sched_set_running
  h->ran = false;
  const uint32_t base = h->n_pending;
  int rc = pending_append(h, rows);
  if (rc != CRANE_OK) return rc;
  if (first_row) *first_row = base;
  return CRANE_OK;
}

int crane_sched_pending_erase(crane_sched_t* h, const uint32_t* rows, uint32_t n) {
  if (!h || (n && !rows)) return CRANE_EINVAL;
  CU(cudaSetDevice(h->device));
  for (uint32_t k = 0; k < n; ++k)
    if (rows[k] >= h->n_pending) return fail(h, CRANE_EINVAL, "pending_erase: row %u out of range", rows[k]);
  h->uploaded = false;
  h->ran = false;
  if (!n) return CRANE_OK;
  for (uint32_t k = 0; k < n; ++k)
    if (!h->h_dead[rows[k]]) { h->h_dead[rows[k]] = 1; h->n_dead++; }
  H2D(h->d_erase_rows, rows, n);
  CRANE_LAUNCH(k_mark_dead, (n + 255) / 256, 256, 0, h->stream, h->d_erase_rows.p, n, h->d_dead.p);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(h->stream));
  return CRANE_OK;
}

int crane_sched_set_running(crane_sched_t* h, const crane_running_t* rn) {
  if (!h) return CRANE_EINVAL;
  if (!h->have_cluster) return fail(h, CRANE_EINVAL, "set_running: set_cluster first");
  CU(cudaSetDevice(h->device));
  h->uploaded = false;
  h->ran = false;
  int rc = set_running(h, rn);
  if (rc != CRANE_OK) return rc;
  h->uploaded = true;
  return CRANE_OK;
}

int crane_sched_upload(crane_sched_t* h, const crane_running_t* rn, const crane_pending_t* pd) {
  if (!h || !pd) return CRANE_EINVAL;
  if (!h->have_cluster) return fail(h, CRANE_EINVAL, "upload: set_cluster first");
  CU(cudaSetDevice(h->device));
  h->uploaded = false;  // a failed upload leaves nothing runnable behind
  h->ran = false;
  CU(cudaEventRecord(h->ev[0], h->stream));
  crane_sched_pending_reset(h);
  int rc = pending_append(h, pd);
  if (rc != CRANE_OK) { crane_sched_pending_reset(h); return rc; }
  rc = set_running(h, rn);
  if (rc != CRANE_OK) return rc;
  CU(cudaEventRecord(h->ev[1], h->stream));
  h->h2d_timed = true;
  h->uploaded = true;
  h->ran = false;
  return CRANE_OK;
}

int crane_sched_run(crane_sched_t* h, int64_t now) {
  if (!h) return CRANE_EINVAL;
  if (!h->uploaded) return fail(h, CRANE_EINVAL, "run: upload first");
  CU(cudaSetDevice(h->device));
  (void)cudaGetLastError();  // a stale error of an earlier, unrelated call in this process is not this run's
  cudaStream_t st = h->stream;
  const uint32_t N = h->n_pending, R = h->n_running;
  h->timing.kernel_launches = 0;

  PendingDev pd{};
  pd.n = N;
  pd.partition = h->d_partition.p;
  pd.time_limit = h->d_time_limit.p;
  pd.submit_time = h->d_submit.p;
  pd.node_num = h->d_node_num.p;
  pd.ntasks_per_node_min = h->d_ntpn.p;
  pd.ntasks_per_node_max = h->d_ntpn_max.p;
  pd.ntasks = h->d_ntasks.p;
  pd.exclusive = h->d_exclusive.p;
  pd.partition_priority = h->d_part_prio.p;
  pd.qos_priority = h->d_qos_prio.p;
  pd.account = h->d_account.p;
  pd.mandated_priority = h->have_mandated ? h->d_mandated.p : nullptr;
  pd.req_node = h->d_req_node.p;
  pd.req_task = h->d_req_task.p;
  pd.req_total = h->d_req_total.p;
  pd.incl_off = h->have_lists_incl ? h->d_incl_off.p : nullptr;
  pd.incl_nodes = h->d_incl_nodes.p;
  pd.excl_off = h->have_lists_excl ? h->d_excl_off.p : nullptr;
  pd.excl_nodes = h->d_excl_nodes.p;
  pd.alloc_off = h->d_alloc_off.p;
  pd.reservation = h->have_pd_resv ? h->d_pd_resv.p : nullptr;
  pd.dead = h->n_dead ? h->d_dead.p : nullptr;

  RunningDev rn{};
  rn.n = R;
  rn.start_time = h->d_rn_start.p;
  rn.end_time = h->d_rn_end.p;
  rn.node_num = h->d_rn_node_num.p;
  rn.partition_priority = h->d_rn_part_prio.p;
  rn.qos_priority = h->d_rn_qos_prio.p;
  rn.account = h->d_rn_account.p;
  rn.view_cpu_raw = h->d_rn_cpu.p;
  rn.view_mem = h->d_rn_mem.p;
  rn.slot_off = h->d_rn_slot_off.p;
  rn.slot_end = h->d_rn_slot_end.p;
  rn.slot_res = h->d_rn_slot_res.p;
  rn.n_accounts = h->n_accounts;
  rn.acc_off = h->d_rn_acc_off.p;
  rn.acc_job = h->d_rn_acc_job.p;
  rn.acc_present = h->d_acc_present.p;

  ClusterDev cl{};
  cl.n_slots = h->n_slots;
  cl.n_parts = h->n_parts;
  cl.n_vparts = h->n_vparts;
  cl.n_comp = h->n_comp;
  cl.part_comp = h->d_part_comp.p;
  cl.part_cidx = h->d_part_cidx.p;
  cl.slot_memb = h->max_comp_parts > 1 ? h->d_slot_memb.p : nullptr;
  cl.n_resv = h->n_resv;
  cl.resv_start = h->d_resv_start.p;
  cl.resv_end = h->d_resv_end.p;
  cl.slot_resv = h->d_slot_resv.p;
  cl.rsv_off = h->d_rsv_off.p;
  cl.rsv_id = h->d_rsv_id.p;
  cl.rsv_res = h->d_rsv_res.p;
  cl.max_part_slots = h->max_part_slots;
  cl.part_base = h->d_part_base.p;
  cl.slot_node = h->d_slot_node.p;
  cl.node_slot = h->d_node_slot.p;
  cl.slot_total = h->d_slot_total.p;
  cl.slot_class = h->d_slot_class.p;
  cl.class_rows = h->d_class_rows.p;

  TimelineDev tl{};
  tl.cap = h->tl_cap;
  tl.n = h->d_tl_n.p;
  tl.ent = h->d_tl_ent.p;
  tl.avail0 = h->d_avail0.p;
  tl.cost0 = h->d_cost0.p;
  tl.skip = h->d_skip.p;
  tl.first_resv = h->d_first_resv.p;

  PlaceDev out{};
  out.reason = h->d_reason.p;
  out.priority = h->d_out_prio.p;
  out.start_time = h->d_out_start.p;
  out.end_time = h->d_out_end.p;
  out.n_alloc = h->d_out_nalloc.p;
  out.alloc_node = h->d_out_node.p;
  out.alloc_ntasks = h->d_out_ntasks.p;
  out.alloc_res = h->d_out_res.p;

  // ---- node state / timelines (R2,R3,R4) -----------------------------------
  CU(cudaEventRecord(h->ev[2], st));
  if (h->total_alloc) {  // allocation slots of jobs without a placement read as zero
    CU(cudaMemsetAsync(h->d_out_node.p, 0, sizeof(uint32_t) * h->total_alloc, st));
    CU(cudaMemsetAsync(h->d_out_ntasks.p, 0, sizeof(uint32_t) * h->total_alloc, st));
    CU(cudaMemsetAsync(h->d_out_res.p, 0, sizeof(Row) * h->total_alloc, st));
  }
  if (h->n_slots) {
    CRANE_LAUNCH(k_node_init, (h->n_slots + 127) / 128, 128, 0, st, cl, rn, tl, now, h->cfg.max_jobs_per_node, h->cfg.cost_policy);
    h->timing.kernel_launches++;
  }
  CU(cudaEventRecord(h->ev[3], st));

  // ---- priority + queue (R5,R6) --------------------------------------------
  uint32_t nq_cap = std::min<uint32_t>(N, h->cfg.scheduled_batch_size);
  CU(cudaMemsetAsync(h->d_part_count.p, 0, sizeof(uint32_t) * (h->n_vparts + 2), st));
  CU(h->d_vpart.ensure(std::max<uint32_t>(N, 1)));
  if (N) {
    PrioCfg pc{};
    pc.type = h->cfg.priority_type;
    pc.favor_small = h->cfg.favor_small;
    pc.w_age = h->cfg.weight_age;
    pc.w_fs = h->cfg.weight_fair_share;
    pc.w_size = h->cfg.weight_job_size;
    pc.w_part = h->cfg.weight_partition;
    pc.w_qos = h->cfg.weight_qos;
    pc.max_age = h->cfg.max_age_s;
    uint64_t* keys = h->d_keys_a.p;
    uint32_t* vals = h->d_vals_a.p;
    if (pc.type != 0) {
      CRANE_LAUNCH(k_bounds_init, 1, 32, 0, st, h->d_bounds.p);
      uint32_t nb = std::min<uint32_t>((N + R + 255) / 256, 592);
      CRANE_LAUNCH(k_bounds, nb, 256, 0, st, pd, rn, now, (uint64_t)h->cfg.max_age_s, h->d_bounds.p);
      h->timing.kernel_launches += 2;
      if (h->n_accounts) {
        CU(cudaMemsetAsync(h->d_acc_service.p, 0, sizeof(double) * h->n_accounts, st));
        CRANE_LAUNCH(k_service, (h->n_accounts + 127) / 128, 128, 0, st, rn, now, h->d_bounds.p, h->d_acc_service.p);
        h->timing.kernel_launches++;
      }
    }
    CRANE_LAUNCH(k_priority, (N + 255) / 256, 256, 0, st, pd, pc, now, h->d_bounds.p, h->d_acc_service.p,
                 h->d_prio.p, h->d_keys_a.p, h->d_vals_a.p);
    h->timing.kernel_launches++;
    if (pc.type != 0 || h->n_dead) {  // (FIFO keeps input order; erased rows only have to go behind the live ones)
      int rc = radix_sort(h, N, pc.type != 0 ? 64 : 8, &keys, &vals);
      if (rc != CRANE_OK) return rc;
    }
    // keys2 go to the buffer not holding `vals`
    uint64_t* key2 = h->d_keys_a.p;
    uint32_t* order = vals;
    if (order != h->d_vals_a.p) {
      // radix_sort works on the (keys_a, vals_a) pair: bring the order back
      CU(cudaMemcpyAsync(h->d_vals_a.p, order, sizeof(uint32_t) * N, cudaMemcpyDeviceToDevice, st));
      order = h->d_vals_a.p;
    }
    CRANE_LAUNCH(k_queue_keys, (N + 255) / 256, 256, 0, st, pd, order, h->d_prio.p, h->cfg.scheduled_batch_size,
                 cl, now, key2, out, h->d_part_count.p, h->d_vpart.p);
    CRANE_LAUNCH(k_part_offsets, 1, 32, 0, st, h->d_part_count.p, h->n_vparts, h->d_part_job_off.p);
    h->timing.kernel_launches += 2;
    int bits = 8;
    while ((1ull << bits) < (uint64_t)h->n_vparts + 2) bits += 8;
    uint64_t* k2s;
    uint32_t* queue;
    int rc = radix_sort(h, N, bits, &k2s, &queue);
    if (rc != CRANE_OK) return rc;
    h->d_queue = queue;
    if (nq_cap) {
      // n_queued <= nq_cap lives on the device (part_job_off[n_parts]); the
      // kernels below bound themselves with it.
      CRANE_LAUNCH(k_build_jobq, (nq_cap + 127) / 128, 128, 0, st, pd, queue, h->d_part_job_off.p + h->n_vparts, h->d_jobq.p, (uint32_t)h->dict_slot, h->d_vpart.p, h->n_comp, h->d_part_cidx.p);
      h->timing.kernel_launches++;
    }
  } else {
    CRANE_LAUNCH(k_part_offsets, 1, 32, 0, st, h->d_part_count.p, h->n_vparts, h->d_part_job_off.p);
    h->timing.kernel_launches++;
  }
  CU(cudaEventRecord(h->ev[4], st));

  // ---- capability bitmap (R7 predicate + R8) -------------------------------
  if (nq_cap && h->n_vparts) {
    uint32_t wpb = 8;
    uint32_t nb = std::min<uint32_t>((nq_cap + wpb - 1) / wpb, 148 * 8);
    CRANE_LAUNCH(k_feas_bitmap, nb, wpb * 32, 0, st, cl, pd, h->d_jobq.p, h->d_part_job_off.p + h->n_vparts, h->words_per_row, h->d_bitmap.p,
                 h->shard_n > 1 ? h->d_sched_owner.p : nullptr, h->shard_rank, (uint32_t)h->dict_slot);
    h->timing.kernel_launches++;
  }
  CU(cudaEventRecord(h->ev[5], st));

  // ---- sequential commit (R7,R9,R10,R11) -----------------------------------
  if (nq_cap && h->n_vparts) {
    Commit2Args c2{};
    c2.cl = cl; c2.tl = tl; c2.jobq = h->d_jobq.p; c2.part_job_off = h->d_part_job_off.p; c2.bitmap = h->d_bitmap.p;
    c2.words_per_row = h->words_per_row; c2.ring = h->v2_ring; c2.out = out; c2.now = now;
    c2.max_window = h->cfg.max_time_window_s; c2.max_jobs = h->cfg.max_jobs_per_node; c2.cost_policy = h->cfg.cost_policy;
    CU(h->d_prof.ensure((size_t)h->n_vparts * 16));
    c2.prof = h->d_prof.p;
    c2.gres = h->dict.n_entries > 0 ? 1u : 0u;
    c2.sched_nparts = h->max_comp_parts > 1 ? h->d_sched_nparts.p : nullptr;
    c2.slot_memb = h->d_slot_memb.p;
    c2.dslot = (uint32_t)h->dict_slot;
    c2.req_node = h->d_req_node.p;
    c2.req_task = h->d_req_task.p;
    size_t smem = commit2_smem_bytes(h->max_part_slots, h->words_per_row, c2.gres != 0, h->v2_ring, h->max_comp_parts);
    uint32_t grid = h->n_vparts;
    if (h->shard_n > 1) {
      c2.part_list = h->d_part_list.p;
      grid = (uint32_t)h->h_part_list.size();
    }
    if (grid) CRANE_LAUNCH(k_commit2, grid, kT2, smem, st, c2);
    h->timing.kernel_launches++;
    if (h->shard_n > 1 && N) {
      CRANE_LAUNCH(k_shard_mask, (N + 255) / 256, 256, 0, st, pd, out, h->d_part_owner.p, h->n_parts, h->shard_rank);
      h->timing.kernel_launches++;
    }
  }
  CU(cudaEventRecord(h->ev[6], st));
  CU(cudaGetLastError());
  h->ran = true;
  return CRANE_OK;
}

int crane_sched_fetch(crane_sched_t* h, crane_placements_t* out) {
  if (!h || !out) return CRANE_EINVAL;
  if (!h->ran) return fail(h, CRANE_EINVAL, "fetch: run first");
  CU(cudaSetDevice(h->device));
  cudaStream_t st = h->stream;
  const uint32_t N = h->n_pending;
  if (N && (!out->reason || !out->priority || !out->start_time || !out->end_time || !out->n_alloc || !out->alloc_off))
    return fail(h, CRANE_EINVAL, "placements: null column");
  if (h->total_alloc && (!out->alloc_node || !out->alloc_ntasks || !out->alloc_res))
    return fail(h, CRANE_EINVAL, "placements: null allocation table");
  // slots of jobs without a placement stay zero
  if (N) {
    CU(cudaMemcpyAsync(out->reason, h->d_reason.p, N, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out->priority, h->d_out_prio.p, sizeof(double) * N, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out->start_time, h->d_out_start.p, sizeof(int64_t) * N, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out->end_time, h->d_out_end.p, sizeof(int64_t) * N, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out->n_alloc, h->d_out_nalloc.p, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, st));
  }
  if (h->total_alloc) {
    CU(cudaMemcpyAsync(out->alloc_node, h->d_out_node.p, sizeof(uint32_t) * h->total_alloc, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out->alloc_ntasks, h->d_out_ntasks.p, sizeof(uint32_t) * h->total_alloc, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out->alloc_res, h->d_out_res.p, sizeof(Row) * h->total_alloc, cudaMemcpyDeviceToHost, st));
  }
  memcpy(out->alloc_off, h->h_alloc_off.data(), sizeof(uint32_t) * ((size_t)N + 1));
  CU(cudaEventRecord(h->ev[7], st));
  CU(cudaStreamSynchronize(st));
  CU(cudaGetLastError());
  float ms = 0;
  // (the upload events exist only after crane_sched_upload: a resident table is filled piecewise)
  if (h->h2d_timed) { cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]); h->timing.h2d_ms = ms; } else h->timing.h2d_ms = 0.f;
  cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]); h->timing.init_ms = ms;
  cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]); h->timing.priority_ms = ms;
  cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]); h->timing.feas_ms = ms;
  cudaEventElapsedTime(&ms, h->ev[5], h->ev[6]); h->timing.commit_ms = ms;
  cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]); h->timing.d2h_ms = ms;
  cudaEventElapsedTime(&ms, h->ev[2], h->ev[6]); h->timing.total_ms = ms;
  return CRANE_OK;
}

int crane_sched_node_select(crane_sched_t* h, int64_t now, const crane_running_t* running,
                            const crane_pending_t* pending, crane_placements_t* out) {
  int rc = crane_sched_upload(h, running, pending);
  if (rc != CRANE_OK) return rc;
  rc = crane_sched_run(h, now);
  if (rc != CRANE_OK) return rc;
  return crane_sched_fetch(h, out);
}

int crane_sched_sync(crane_sched_t* h, float* run_ms) {
  if (!h) return CRANE_EINVAL;
  CU(cudaSetDevice(h->device));
  CU(cudaStreamSynchronize(h->stream));
  CU(cudaGetLastError());
  if (h->ran) {
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[3]); h->timing.init_ms = ms;
    cudaEventElapsedTime(&ms, h->ev[3], h->ev[4]); h->timing.priority_ms = ms;
    cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]); h->timing.feas_ms = ms;
    cudaEventElapsedTime(&ms, h->ev[5], h->ev[6]); h->timing.commit_ms = ms;
    cudaEventElapsedTime(&ms, h->ev[2], h->ev[6]); h->timing.total_ms = ms;
    if (run_ms) *run_ms = ms;
  } else if (run_ms) {
    *run_ms = 0.f;
  }
  return CRANE_OK;
}

int crane_sched_debug_profile(crane_sched_t* h, unsigned long long* dst, size_t cap) {
  if (!h || !h->ran || !dst) return CRANE_EINVAL;
  CU(cudaSetDevice(h->device));
  size_t n = std::min(cap, (size_t)h->n_vparts * 16);
  CU(cudaStreamSynchronize(h->stream));
  if (n) CU(cudaMemcpy(dst, h->d_prof.p, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return CRANE_OK;
}

int crane_sched_qos_filter(crane_sched_t* h, const crane_qos_table_t* qt, uint8_t* reason) {
  if (!h || !qt || !reason) return CRANE_EINVAL;
  if (!h->ran) return fail(h, CRANE_EINVAL, "qos_filter: run first");
  if (!h->have_qos_cols) return fail(h, CRANE_EINVAL, "qos_filter: pending.qos / pending.user were not uploaded");
  CU(cudaSetDevice(h->device));
  const uint32_t N = h->n_pending, Q = qt->n_qos, U = qt->n_users, A = qt->n_accounts;
  if (Q == 0 || Q >= 65535) return fail(h, CRANE_EINVAL, "qos_filter: n_qos out of range");
  if (A > 65535) return fail(h, CRANE_ENOSYS, "qos_filter: more than 65535 accounts");
  if (!qt->valid || !qt->max_jobs_per_user || !qt->max_jobs_per_account || !qt->max_jobs ||
      !qt->max_cpus_per_user_raw || !qt->max_wall || !qt->max_tres_per_user || !qt->max_tres_per_account ||
      !qt->max_tres || !qt->chain_off || !qt->user_usage || !qt->qos_usage || (A && !qt->account_usage))
    return fail(h, CRANE_EINVAL, "qos_filter: null column");
  // the levels of one job are checked lane-parallel: user + chain + qos <= 32
  // lanes, and one usage entry per lane (a chain never repeats an account)
  for (uint32_t i = 0; i < N; ++i) {
    const uint32_t c0 = qt->chain_off[i], c1 = qt->chain_off[i + 1];
    if (c1 < c0 || c1 - c0 > 30) return fail(h, CRANE_ENOSYS, "qos_filter: pending[%u]: account chain longer than 30", i);
    for (uint32_t a = c0; a < c1; ++a) {
      if (qt->chain_acct[a] >= A) return fail(h, CRANE_EINVAL, "qos_filter: pending[%u]: account out of range", i);
      for (uint32_t b = c0; b < a; ++b)
        if (qt->chain_acct[a] == qt->chain_acct[b])
          return fail(h, CRANE_EINVAL, "qos_filter: pending[%u]: account repeated in chain", i);
    }
  }
  // user ids are checked on the device side of the table: against n_users here
  // (pending.user was validated only for presence at upload)
  {
    std::vector<uint32_t> users(N);
    if (N) CU(cudaMemcpyAsync(users.data(), h->d_user.p, (size_t)N * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (uint32_t i = 0; i < N; ++i)
      if (users[i] >= U) return fail(h, CRANE_EINVAL, "qos_filter: pending[%u]: user out of range", i);
  }
  std::vector<uint32_t> u32(3 * (size_t)Q);
  std::vector<int64_t> i64(2 * (size_t)Q);
  std::vector<crane_tres_limit_t> tres(3 * (size_t)Q);
  for (uint32_t k = 0; k < Q; ++k) {
    u32[k] = qt->max_jobs_per_user[k]; u32[Q + k] = qt->max_jobs_per_account[k]; u32[2 * Q + k] = qt->max_jobs[k];
    i64[k] = qt->max_cpus_per_user_raw[k]; i64[Q + k] = qt->max_wall[k];
    tres[k] = qt->max_tres_per_user[k]; tres[Q + k] = qt->max_tres_per_account[k]; tres[2 * Q + k] = qt->max_tres[k];
  }
  H2D(h->d_q_u32, u32.data(), u32.size());
  H2D(h->d_q_i64, i64.data(), i64.size());
  H2D(h->d_q_tres, tres.data(), tres.size());
  H2D(h->d_q_valid, qt->valid, Q);
  H2D(h->d_q_chain_off, qt->chain_off, N + 1);
  H2D(h->d_q_chain, qt->chain_acct, qt->chain_off[N]);
  H2D(h->d_q_user_usage, qt->user_usage, (size_t)U * Q);
  H2D(h->d_q_acct_usage, qt->account_usage, (size_t)A * Q);
  H2D(h->d_q_qos_usage, qt->qos_usage, Q);
  QosDev q{};
  q.n_qos = Q; q.n_users = U; q.n_accounts = A; q.n_jobs = N;
  q.valid = h->d_q_valid.p;
  q.max_jobs_per_user = h->d_q_u32.p; q.max_jobs_per_account = h->d_q_u32.p + Q; q.max_jobs = h->d_q_u32.p + 2 * Q;
  q.max_cpus_per_user_raw = h->d_q_i64.p; q.max_wall = h->d_q_i64.p + Q;
  q.tres_user = h->d_q_tres.p; q.tres_account = h->d_q_tres.p + Q; q.tres_qos = h->d_q_tres.p + 2 * Q;
  q.chain_off = h->d_q_chain_off.p; q.chain_acct = h->d_q_chain.p;
  q.user_usage = h->d_q_user_usage.p; q.account_usage = h->d_q_acct_usage.p; q.qos_usage = h->d_q_qos_usage.p;
  q.qos = h->d_qos.p; q.user = h->d_user.p; q.time_limit = h->d_time_limit.p;
  q.n_alloc = h->d_out_nalloc.p; q.alloc_off = h->d_alloc_off.p; q.alloc_res = h->d_out_res.p;
  q.reason = h->d_reason.p;
  if (N) {
    // the jobs the pass walks, grouped by qos in job-id order: key + stable radix sort
    const int bits = Q < 255 ? 8 : 16;
    const uint64_t sentinel = (1ull << bits) - 1;
    CU(cudaEventRecord(h->ev[8], h->stream));
    CU(h->d_keys_a.ensure(N)); CU(h->d_keys_b.ensure(N)); CU(h->d_vals_a.ensure(N)); CU(h->d_vals_b.ensure(N));
    CRANE_LAUNCH(k_qos_keys, (N + 255) / 256, 256, 0, h->stream, q, h->d_keys_a.p, h->d_vals_a.p, sentinel);
    uint64_t* keys;
    uint32_t* list;
    int rc = radix_sort(h, N, bits, &keys, &list);
    if (rc != CRANE_OK) return rc;
    CU(h->d_q_off.ensure((size_t)Q + 1));
    CRANE_LAUNCH(k_qos_offsets, (Q + 1 + 127) / 128, 128, 0, h->stream, keys, N, Q, h->d_q_off.p);
    // the usage column of one qos in shared memory when it fits beside the job buffers
    const size_t table_bytes = ((size_t)U + A) * sizeof(crane_meta_resource_t);
    const bool force_global = getenv("CRANE_QOS_TABLES_GLOBAL") != nullptr;  // tests: the path of tables that do not fit
    const uint32_t in_smem = !force_global && table_bytes + 40 * 1024 <= h->v2_budget ? 1u : 0u;
    const size_t dyn = in_smem ? table_bytes : 0;
    CU(cudaFuncSetAttribute(k_qos_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    CRANE_LAUNCH(k_qos_chain, Q, kQosThreads, dyn, h->stream, q, h->dict, list, h->d_q_off.p, in_smem);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev[9], h->stream));
    h->timing.kernel_launches += 3;
  }
  if (N) CU(cudaMemcpyAsync(reason, h->d_reason.p, N, cudaMemcpyDeviceToHost, h->stream));
  if ((size_t)U * Q) CU(cudaMemcpyAsync(qt->user_usage, h->d_q_user_usage.p, (size_t)U * Q * sizeof(crane_meta_resource_t), cudaMemcpyDeviceToHost, h->stream));
  if ((size_t)A * Q) CU(cudaMemcpyAsync(qt->account_usage, h->d_q_acct_usage.p, (size_t)A * Q * sizeof(crane_meta_resource_t), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaMemcpyAsync(qt->qos_usage, h->d_q_qos_usage.p, (size_t)Q * sizeof(crane_meta_resource_t), cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  h->timing.qos_ms = 0.f;
  if (N) cudaEventElapsedTime(&h->timing.qos_ms, h->ev[8], h->ev[9]);
  return CRANE_OK;
}

int crane_sched_get_timing(const crane_sched_t* h, crane_sched_timing_t* t) {
  if (!h || !t) return CRANE_EINVAL;
  *t = h->timing;
  return CRANE_OK;
}

int crane_sched_debug_bitmap(crane_sched_t* h, uint32_t* dst, size_t cap_words, uint32_t* rows,
                             uint32_t* words_per_row) {
  if (!h || !h->ran) return CRANE_EINVAL;
  CU(cudaSetDevice(h->device));
  uint32_t nq = 0;
  CU(cudaMemcpy(&nq, h->d_part_job_off.p + h->n_vparts, sizeof(uint32_t), cudaMemcpyDeviceToHost));
  if (rows) *rows = nq;
  if (words_per_row) *words_per_row = h->words_per_row;
  size_t n = std::min(cap_words, (size_t)nq * h->words_per_row);
  if (dst && n) CU(cudaMemcpy(dst, h->d_bitmap.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return CRANE_OK;
}

}  // extern "C"
