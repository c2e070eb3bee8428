// ref_harness.cpp — drives the REFERENCE's own SchedulerAlgo::NodeSelect.
//
// TEST INFRASTRUCTURE ONLY (oracle/ rules). This translation unit is the only
// code of ours in oracle/_ref/libcrane_ref.so; everything between the
// `#include "*.inc"` lines is the reference's unmodified text, sliced out of
// /root/reference at build time by oracle/ref_build.py (never committed):
//   ph_h.inc    PublicHeader.h   TypeSlotsMap .. ResourceView             (:410-748)
//   ph_cpp.inc  PublicHeader.cpp the non-protobuf member functions        (:23-966)
//   js_h.inc    JobScheduler.h   IUpdateNodeCostPolicy .. SchedulerAlgo   (:30-979)
//   js_cpp.inc  JobScheduler.cpp LocalScheduler::*, SchedulerAlgo::NodeSelect,
//                                MultiFactorPriority::*                   (:5165-5868, 6526-6739)
// The harness converts the C-ABI tables (include/crane_sched.h) into the
// reference's objects with the same naming scheme as oracle/crane_oracle.cpp
// (zero-padded names, so lexicographic order == index order), calls
// NodeSelect(now, running, pending) and writes the PdJobInScheduler outputs
// back in the C-ABI layout.
//
// Where run-to-run nondeterminism of the reference is pinned by the shims, see
// ref_shim/absl_shim.h; what it canNOT pin (documented in tests/test_ref_pin.py):
//   * std::unordered_map iteration inside GetFeasibleResourceInNode
//     (PublicHeader.cpp:549,564,583) stays libstdc++ hash order -> cases where
//     that order is observable (a node with several types of one gres name AND a
//     request with an untyped remainder) are outside the pin;
//   * std::ranges::sort is unstable (JobScheduler.cpp:6541) -> cases with equal
//     computed priorities are outside the pin.
#include "ctld_shim.h"

// ---------------- reference text: resource algebra ----------------------------------------
#include "ph_h.inc"
#include "ph_cpp.inc"

#include "ctld_shim2.h"

// ---------------- reference text: scheduler --------------------------------------------------
// The slices are included with `private` opened up so that the harness can also
// drive single pieces (NodeState timelines, EarliestStartSubsetSelector) for the
// known-answer entry points below; the reference text itself is untouched.
#define private public
#define protected public
namespace Ctld {
#include "js_h.inc"
#include "js_cpp.inc"
// ---------------- reference text: the QoS check + malloc of the commit loop -------------------
//   amc_h.inc    Accounting/AccountMetaContainer.h    struct MetaResource                    (:30-47)
//   amc_cpp.inc  Accounting/AccountMetaContainer.cpp  MetaResource operators, CheckAndMallocQosResource,
//                CheckQosResource_, CheckTres_, CheckGres_, LockAccountStripes_, DoMallocResource_
#include "amc_h.inc"
#include "ctld_shim3.h"
#include "amc_cpp.inc"
}  // namespace Ctld
#undef private
#undef protected

#include <chrono>
#include <cstring>

#include "include/crane_sched.h"
#include "oracle/crane_ref.h"

namespace {

std::string NodeName(uint32_t i) { char b[24]; snprintf(b, sizeof b, "cn%08u", i); return b; }
std::string PartName(uint32_t i) { char b[24]; snprintf(b, sizeof b, "p%08u", i); return b; }
std::string ResvName(uint32_t i) { char b[24]; snprintf(b, sizeof b, "resv%06u", i); return b; }

struct Dict {
  uint32_t n_entries{0};
  uint8_t entry_name[CRANE_GRES_ENTRIES]{};
  std::string name_str[CRANE_GRES_NAMES], type_str[CRANE_GRES_ENTRIES], slot_str[CRANE_MAX_SLOTS];
  explicit Dict(const crane_cluster_t* c) {
    n_entries = std::min<uint32_t>(c->n_gres_entries, CRANE_GRES_ENTRIES);
    for (uint32_t e = 0; e < n_entries; ++e) entry_name[e] = c->gres_entry_name[e];
    char b[32];
    for (int g = 0; g < CRANE_GRES_NAMES; ++g) { snprintf(b, sizeof b, "gres%02d", g); name_str[g] = b; }
    for (int e = 0; e < CRANE_GRES_ENTRIES; ++e) { snprintf(b, sizeof b, "type%02d", e); type_str[e] = b; }
    for (int s = 0; s < CRANE_MAX_SLOTS; ++s) { snprintf(b, sizeof b, "/dev/slot%02d", s); slot_str[s] = b; }
  }
};

ResourceInNodeV3 FromAbi(const Dict& d, const crane_res_in_node_t& r) {
  ResourceInNodeV3 x;
  x.GetCpuSet().cpu_count = cpu_t::from_raw_value(r.cpu_raw);
  for (int w = 0; w < CRANE_CORE_WORDS; ++w)
    for (int b = 0; b < 64; ++b)
      if (r.core[w] >> b & 1) x.GetCpuSet().core_ids.insert(w * 64 + b);
  x.SetMemoryBytes(r.mem);
  x.SetMemorySwBytes(r.mem_sw);
  for (uint32_t e = 0; e < d.n_entries; ++e) {
    if (!r.gres[e]) continue;
    std::set<SlotId>& s = x.GetGres()[d.name_str[d.entry_name[e]]][d.type_str[e]];
    for (int b = 0; b < CRANE_MAX_SLOTS; ++b)
      if (r.gres[e] >> b & 1) s.insert(d.slot_str[b]);
  }
  return x;
}

void ToAbi(const Dict& d, const ResourceInNodeV3& x, crane_res_in_node_t* r) {
  memset(r, 0, sizeof *r);
  r->cpu_raw = x.GetCpuSet().cpu_count.raw_value();
  r->mem = x.GetMemoryBytes();
  r->mem_sw = x.GetMemorySwBytes();
  for (uint32_t id : x.GetCpuSet().core_ids)
    if (id < 64 * CRANE_CORE_WORDS) r->core[id / 64] |= 1ull << (id % 64);
  for (uint32_t e = 0; e < d.n_entries; ++e) {
    const auto& names = x.GetGres().name_type_slots_map;
    auto nit = names.find(d.name_str[d.entry_name[e]]);
    if (nit == names.end()) continue;
    auto tit = nit->second.type_slots_map.find(d.type_str[e]);
    if (tit == nit->second.type_slots_map.end()) continue;
    for (int b = 0; b < CRANE_MAX_SLOTS; ++b)
      if (tit->second.count(d.slot_str[b])) r->gres[e] |= uint16_t(1u << b);
  }
}

ResourceView ViewFromAbi(const Dict& d, const crane_res_view_t& v) {
  ResourceView x;
  x.SetCpuCount(cpu_t::from_raw_value(v.cpu_raw));
  x.SetMemoryBytes(v.mem);
  x.SetMemorySwBytes(v.mem_sw);
  for (int g = 0; g < CRANE_GRES_NAMES; ++g) {
    bool any = v.gres_total[g] != 0;
    for (uint32_t e = 0; e < d.n_entries; ++e)
      if (d.entry_name[e] == g && v.gres_spec[e]) any = true;
    if (!any) continue;
    GresCount& gc = x.GetGresMap()[d.name_str[g]];
    gc.total = v.gres_total[g];
    for (uint32_t e = 0; e < d.n_entries; ++e)
      if (d.entry_name[e] == g && v.gres_spec[e]) gc.specified[d.type_str[e]] = v.gres_spec[e];
  }
  return x;
}

int ReasonCode(const std::string& s) {
  if (s.empty()) return CRANE_REASON_NONE;
  if (s == "Priority") return CRANE_REASON_PRIORITY;
  if (s == "Resource") return CRANE_REASON_RESOURCE;
  if (s == "Resource Reserved") return CRANE_REASON_RESERVED;
  if (s == "Partition Not Found") return CRANE_REASON_PART_NOT_FOUND;
  if (s == "Reservation Not Found") return CRANE_REASON_RESV_NOT_FOUND;
  if (s == "Preempted") return CRANE_REASON_PREEMPTED;
  return 255;
}

}  // namespace

extern "C" const char* crane_ref_describe(void) {
  return "reference text of JobScheduler.{h,cpp} + PublicHeader.{h,cpp} (hot-path slices), "
         "compiled against oracle/ref_shim";
}

extern "C" int crane_ref_node_select(const crane_sched_config_t* cfg, const crane_cluster_t* cl, int64_t now_s,
                                     const crane_running_t* running, const crane_pending_t* pending,
                                     const crane_ref_extra_t* extra, crane_placements_t* out,
                                     double* elapsed_ms) {
  if (!cfg || !cl || !pending || !out) return CRANE_EINVAL;
  if (cfg->cost_policy != 0) return CRANE_ENOSYS;  // the reference only has MinCpuTimeRatioFirst
  using namespace Ctld;
  Dict dict(cl);
  const uint32_t M = cl->n_nodes, N = pending->n, R = running ? running->n : 0;
  const absl::Time now = absl::FromUnixSeconds(now_s);

  absl::shim_internal::arena().reset();
  g_config = Config{};
  g_config.PriorityConfig.Type = cfg->priority_type ? Config::Priority::MultiFactor : Config::Priority::Basic;
  g_config.PriorityConfig.FavorSmall = cfg->favor_small != 0;
  g_config.PriorityConfig.MaxAge = cfg->max_age_s;
  g_config.PriorityConfig.WeightAge = cfg->weight_age;
  g_config.PriorityConfig.WeightFairShare = cfg->weight_fair_share;
  g_config.PriorityConfig.WeightJobSize = cfg->weight_job_size;
  g_config.PriorityConfig.WeightPartition = cfg->weight_partition;
  g_config.PriorityConfig.WeightQoS = cfg->weight_qos;
  g_config.ScheduledBatchSize = cfg->scheduled_batch_size;
  // kAlgoMaxJobNumPerNode / kAlgoMaxTimeWindow are compile-time constants of the reference
  if (cfg->max_jobs_per_node != 1000 || cfg->max_time_window_s != 7 * 24 * 3600) return CRANE_ENOSYS;

  g_meta_container = std::make_unique<CranedMetaContainer>();
  g_account_manager = std::make_unique<AccountManager>();
  g_license_manager = std::make_unique<LicensesManager>();
  g_job_scheduler = std::make_unique<JobSchedulerStub>();

  std::vector<std::string> node_names(M);
  for (uint32_t i = 0; i < M; ++i) {
    node_names[i] = NodeName(i);
    CranedMeta& m = g_meta_container->craneds[node_names[i]].v;
    m.alive = cl->alive[i] != 0;
    m.drain = cl->drain[i] != 0;
    m.res_total = FromAbi(dict, cl->res_total[i]);
  }
  for (uint32_t p = 0; p < cl->n_partitions; ++p) {
    PartitionMeta& pm = g_meta_container->partitions[PartName(p)].v;
    for (uint32_t k = cl->part_off[p]; k < cl->part_off[p + 1]; ++k) pm.craned_ids.insert(node_names[cl->part_nodes[k]]);
  }
  const uint32_t n_resv = extra ? extra->n_resv : 0;
  for (uint32_t r = 0; r < n_resv; ++r) {
    ResvMeta& rm = g_meta_container->resvs[ResvName(r)].v;
    rm.start_time = absl::FromUnixSeconds(extra->resv_start[r]);
    rm.end_time = absl::FromUnixSeconds(extra->resv_end[r]);
    for (uint32_t k = extra->resv_off[r]; k < extra->resv_off[r + 1]; ++k) {
      rm.craned_ids.insert(node_names[extra->resv_node[k]]);
      rm.res_total.AddResourceInNode(node_names[extra->resv_node[k]], FromAbi(dict, extra->resv_res[k]));
    }
  }

  // --- the snapshot JobScheduler.cpp:1090-1127 takes before the timed call -----------------
  std::vector<JobInCtld> pd_src(N), rn_src(R);
  std::vector<std::unique_ptr<PdJobInScheduler>> pd;
  std::vector<std::unique_ptr<RnJobInScheduler>> rn;
  pd.reserve(N);
  rn.reserve(R);
  for (uint32_t i = 0; i < N; ++i) {
    JobInCtld& j = pd_src[i];
    j.job_id = i + 1;
    j.time_limit = absl::Seconds(pending->time_limit[i]);
    j.partition_id = PartName(pending->partition[i]);
    if (extra && extra->pd_resv && extra->pd_resv[i] != 0xffffffffu) j.reservation = ResvName(extra->pd_resv[i]);
    j.req_node_res_view = ViewFromAbi(dict, pending->req_node[i]);
    j.req_task_res_view = ViewFromAbi(dict, pending->req_task[i]);
    j.req_total_res_view = ViewFromAbi(dict, pending->req_total[i]);
    j.node_num = pending->node_num[i];
    j.ntasks_per_node_min = pending->ntasks_per_node_min[i];
    j.ntasks_per_node_max = pending->ntasks_per_node_max[i];
    j.ntasks = pending->ntasks[i];
    j.exclusive = pending->exclusive[i] != 0;
    if (pending->incl_off)
      for (uint32_t k = pending->incl_off[i]; k < pending->incl_off[i + 1]; ++k)
        j.included_nodes.insert(NodeName(pending->incl_nodes[k]));
    if (pending->excl_off)
      for (uint32_t k = pending->excl_off[i]; k < pending->excl_off[i + 1]; ++k)
        j.excluded_nodes.insert(NodeName(pending->excl_nodes[k]));
    j.submit_time = absl::FromUnixSeconds(pending->submit_time[i]);
    j.partition_priority = pending->partition_priority[i];
    j.qos_priority = pending->qos_priority[i];
    j.account = "acct" + std::to_string(pending->account[i]);
    j.qos = "qos" + std::to_string(pending->qos ? pending->qos[i] : 0);
    j.username = "user" + std::to_string(pending->user ? pending->user[i] : 0);
    j.mandated_priority = pending->mandated_priority ? pending->mandated_priority[i] : 0.0;
    pd.push_back(std::make_unique<PdJobInScheduler>(&j));
  }
  for (uint32_t i = 0; i < R; ++i) {
    JobInCtld& j = rn_src[i];
    j.job_id = 0x40000000u + i;
    j.start_time = absl::FromUnixSeconds(running->start_time[i]);
    j.end_time = absl::FromUnixSeconds(running->end_time[i]);
    j.time_limit = j.end_time - j.start_time;
    j.partition_priority = running->partition_priority[i];
    j.qos_priority = running->qos_priority[i];
    j.account = "acct" + std::to_string(running->account[i]);
    j.qos = "qos" + std::to_string(extra && extra->rn_qos ? extra->rn_qos[i] : 0);
    if (extra && extra->rn_resv && extra->rn_resv[i] != 0xffffffffu) j.reservation = ResvName(extra->rn_resv[i]);
    j.allocated_res_view.SetCpuCount(cpu_t::from_raw_value(running->view_cpu_raw[i]));
    j.allocated_res_view.SetMemoryBytes(running->view_mem[i]);
    for (uint32_t k = running->alloc_off[i]; k < running->alloc_off[i + 1]; ++k)
      j.allocated_res.AddResourceInNode(node_names[running->alloc_node[k]], FromAbi(dict, running->alloc_res[k]));
    rn.push_back(std::make_unique<RnJobInScheduler>(&j));
    rn.back()->node_num = running->node_num[i];  // the reference leaves it uninitialised (deviation D7)
  }
  if (extra && extra->preempt_qos) {
    g_config.Preempt.PreemptType = crane::grpc::PreemptType::PREEMPT_QOS;
    for (uint32_t q = 0; q < extra->n_qos; ++q) {
      auto qp = std::make_unique<Qos>();
      for (uint32_t v = 0; v < extra->n_qos; ++v)
        if (extra->preempt_qos[(size_t)q * extra->n_qos + v]) qp->preempt.insert("qos" + std::to_string(v));
      g_account_manager->qos_map["qos" + std::to_string(q)] = std::move(qp);
    }
  }

  BasicPriority basic;
  MultiFactorPriority multi;
  IPrioritySorter* sorter = cfg->priority_type ? static_cast<IPrioritySorter*>(&multi) : static_cast<IPrioritySorter*>(&basic);
  auto t0 = std::chrono::steady_clock::now();
  {
    SchedulerAlgo algo(sorter);
    algo.NodeSelect(now, rn, pd);  // JobScheduler.cpp:1141
  }
  auto t1 = std::chrono::steady_clock::now();
  if (elapsed_ms) *elapsed_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();

  // --- write-back -----------------------------------------------------------------------------
  std::unordered_map<std::string, uint32_t> node_index;
  node_index.reserve(M);
  for (uint32_t i = 0; i < M; ++i) node_index.emplace(node_names[i], i);
  uint32_t off = 0;
  for (uint32_t i = 0; i < N; ++i) {
    const PdJobInScheduler* j = pd[i].get();
    out->alloc_off[i] = off;
    out->reason[i] = (uint8_t)ReasonCode(j->reason);
    out->priority[i] = j->priority;
    const int64_t st = absl::ToUnixSeconds(j->start_time), en = absl::ToUnixSeconds(j->end_time);
    const bool placed = st != 0 && en != 0;
    out->start_time[i] = placed ? st : 0;
    out->end_time[i] = placed ? en : 0;
    out->n_alloc[i] = placed ? j->node_num : 0;
    for (uint32_t k = 0; k < j->node_num; ++k) {
      out->alloc_node[off + k] = 0;
      out->alloc_ntasks[off + k] = 0;
      memset(&out->alloc_res[off + k], 0, sizeof(crane_res_in_node_t));
    }
    if (placed) {
      std::map<uint32_t, const ResourceInNodeV3*> by_node;  // node-index ascending (deviation D3)
      for (const auto& [id, res] : j->allocated_res.EachNodeResMap()) by_node[node_index.at(id)] = &res;
      uint32_t k = 0;
      for (const auto& [idx, res] : by_node) {
        if (k >= j->node_num) break;
        out->alloc_node[off + k] = idx;
        out->alloc_ntasks[off + k] = j->craned_id_to_task_num.at(node_names[idx]);
        ToAbi(dict, *res, &out->alloc_res[off + k]);
        ++k;
      }
    }
    off += j->node_num;
  }
  out->alloc_off[N] = off;
  if (extra && extra->preempted_running && R)
    for (uint32_t i = 0; i < R; ++i) extra->preempted_running[i] = 0;
  if (extra && extra->preempted_running)
    for (job_id_t id : g_job_scheduler->preempt_cancelled)
      if (id >= 0x40000000u && id - 0x40000000u < R) extra->preempted_running[id - 0x40000000u] = 1;
  pd.clear();
  rn.clear();
  g_meta_container.reset();
  g_account_manager.reset();
  absl::shim_internal::arena().reset();
  return CRANE_OK;
}

// ResourceView::GetFeasibleResourceInNode (PublicHeader.cpp:519-599) of the reference itself
extern "C" int crane_ref_feasible(const crane_cluster_t* cl, const crane_res_view_t* req,
                                  const crane_res_in_node_t* avail, crane_res_in_node_t* alloc) {
  Dict d(cl);
  ResourceView v = ViewFromAbi(d, *req);
  ResourceInNodeV3 a = FromAbi(d, *avail), got;
  if (!v.GetFeasibleResourceInNode(a, &got)) return 0;
  if (alloc) ToAbi(d, got, alloc);
  return 1;
}

// ResourceInNodeV3::Ckmin (PublicHeader.cpp:815-827) of the reference itself
extern "C" void crane_ref_ckmin(const crane_cluster_t* cl, crane_res_in_node_t* a, const crane_res_in_node_t* b) {
  Dict d(cl);
  ResourceInNodeV3 x = FromAbi(d, *a);
  x.Ckmin(FromAbi(d, *b));
  ToAbi(d, x, a);
}

// operator<=(ResourceInNodeV3, ResourceInNodeV3) (PublicHeader.cpp:886-890) of the reference itself
extern "C" int crane_ref_res_le(const crane_cluster_t* cl, const crane_res_in_node_t* a, const crane_res_in_node_t* b) {
  Dict d(cl);
  return FromAbi(d, *a) <= FromAbi(d, *b) ? 1 : 0;
}

// EarliestStartSubsetSelector::CalcEarliestStartTime (JobScheduler.h:786-859) of the
// reference itself on K explicit timelines: node k has n_seg[k] entries
// (times[off[k]+i], rows[off[k]+i]) and the job holds alloc[k] on it. Returns 1 and
// *start when a start time is found, 0 otherwise.
extern "C" int crane_ref_earliest_start(const crane_cluster_t* cl, uint32_t n_nodes, uint32_t node_num,
                                        const uint32_t* off, const int64_t* times, const crane_res_in_node_t* rows,
                                        const crane_res_in_node_t* alloc, int64_t now_s, int64_t time_limit,
                                        int64_t* start) {
  using namespace Ctld;
  Dict d(cl);
  std::vector<std::unique_ptr<SchedulerAlgo::NodeState>> states;
  std::vector<SchedulerAlgo::NodeState*> ptrs;
  JobInCtld src;
  src.node_num = node_num;
  src.time_limit = absl::Seconds(time_limit);
  PdJobInScheduler job(&src);
  for (uint32_t k = 0; k < n_nodes; ++k) {
    states.push_back(std::make_unique<SchedulerAlgo::NodeState>(NodeName(k), ResourceInNodeV3{}));
    for (uint32_t i = off[k]; i < off[k + 1]; ++i)
      states.back()->time_avail_res_map.emplace(absl::FromUnixSeconds(times[i]), FromAbi(d, rows[i]));
    job.allocated_res.AddResourceInNode(NodeName(k), FromAbi(d, alloc[k]));
    ptrs.push_back(states.back().get());
  }
  SchedulerAlgo::EarliestStartSubsetSelector sel(&job, ptrs);
  if (!sel.CalcEarliestStartTime(absl::FromUnixSeconds(now_s), &job)) return 0;
  *start = absl::ToUnixSeconds(job.start_time);
  return 1;
}

// ---- AccountMetaContainer::CheckAndMallocQosResource over the jobs a tick starts ------------------
// Same contract as crane_oracle_qos_filter: job-id order, jobs with reason NONE and an
// allocation; placements->reason and the usage tables of `qt` are updated in place. Names are
// zero-padded (map order == index order); a usage entry lists a gres name / type only when its
// count is non-zero (the oracle's reading of "is in the map", deviation D8).
namespace {
std::string QosName(uint32_t i) { char b[24]; snprintf(b, sizeof b, "qos%06u", i); return b; }
std::string UserName(uint32_t i) { char b[24]; snprintf(b, sizeof b, "user%07u", i); return b; }
std::string AcctName(uint32_t i) { char b[24]; snprintf(b, sizeof b, "acct%07u", i); return b; }

ResourceView LimitFromAbi(const Dict& d, const crane_tres_limit_t& L) {
  ResourceView x;
  x.SetCpuCount(cpu_t::from_raw_value(L.view.cpu_raw));
  x.SetMemoryBytes(L.view.mem);
  x.SetMemorySwBytes(L.view.mem_sw);
  for (int g = 0; g < CRANE_GRES_NAMES; ++g) {
    if (!((L.gres_name_present >> g) & 1u)) continue;
    GresCount& gc = x.GetGresMap()[d.name_str[g]];
    gc.total = L.view.gres_total[g];
    for (uint32_t e = 0; e < d.n_entries; ++e)
      if (d.entry_name[e] == g && ((L.gres_spec_present >> e) & 1u)) gc.specified[d.type_str[e]] = L.view.gres_spec[e];
  }
  return x;
}
Ctld::MetaResource MetaFromAbi(const Dict& d, const crane_meta_resource_t& m) {
  Ctld::MetaResource r;
  r.resource.SetCpuCount(cpu_t::from_raw_value(m.cpu_raw));
  r.resource.SetMemoryBytes(m.mem);
  r.resource.SetMemorySwBytes(m.mem_sw);
  for (int g = 0; g < CRANE_GRES_NAMES; ++g) {
    bool any = m.gres_total[g] != 0;
    for (uint32_t e = 0; e < d.n_entries; ++e)
      if (d.entry_name[e] == g && m.gres_spec[e]) any = true;
    if (!any) continue;
    GresCount& gc = r.resource.GetGresMap()[d.name_str[g]];
    gc.total = m.gres_total[g];
    for (uint32_t e = 0; e < d.n_entries; ++e)
      if (d.entry_name[e] == g && m.gres_spec[e]) gc.specified[d.type_str[e]] = m.gres_spec[e];
  }
  r.jobs_count = m.jobs_count;
  r.wall_time = absl::Seconds(m.wall_time);
  return r;
}
void MetaToAbi(const Dict& d, const Ctld::MetaResource& r, crane_meta_resource_t* m) {
  memset(m, 0, sizeof *m);
  m->cpu_raw = r.resource.GetCpuCount().raw_value();
  m->mem = r.resource.GetMemoryBytes();
  m->mem_sw = r.resource.GetMemorySwBytes();
  for (int g = 0; g < CRANE_GRES_NAMES; ++g) {
    auto it = r.resource.GetGresMap().find(d.name_str[g]);
    if (it == r.resource.GetGresMap().end()) continue;
    m->gres_total[g] = (uint32_t)it->second.total;
    for (uint32_t e = 0; e < d.n_entries; ++e) {
      if (d.entry_name[e] != g) continue;
      auto tit = it->second.specified.find(d.type_str[e]);
      if (tit != it->second.specified.end()) m->gres_spec[e] = (uint32_t)tit->second;
    }
  }
  m->jobs_count = r.jobs_count;
  m->wall_time = absl::ToInt64Seconds(r.wall_time);
}
int QosReasonCode(const std::string& s) {
  if (s == "QosCpuResourceLimit") return CRANE_REASON_QOS_CPU;
  if (s == "QosJobsResourceLimit") return CRANE_REASON_QOS_JOBS;
  if (s == "QosWallTimeLimit") return CRANE_REASON_QOS_WALL;
  if (s == "QosMemResourceLimit") return CRANE_REASON_QOS_MEM;
  if (s == "QosGresResourceLimit") return CRANE_REASON_QOS_GRES;
  if (s == "InvalidQOS") return CRANE_REASON_QOS_INVALID;
  return 255;  // "QosResourceLimit": an entry missing from a map — cannot happen with dense tables
}
}  // namespace

extern "C" int crane_ref_qos_filter(const crane_cluster_t* cl, const crane_pending_t* pending,
                                    crane_placements_t* pl, const crane_qos_table_t* qt) {
  if (!cl || !pending || !pl || !qt || !pending->qos || !pending->user) return CRANE_EINVAL;
  using namespace Ctld;
  Dict dict(cl);
  const uint32_t N = pending->n, Q = qt->n_qos, U = qt->n_users, A = qt->n_accounts;
  absl::shim_internal::arena().reset();
  g_account_manager = std::make_unique<AccountManager>();
  for (uint32_t q = 0; q < Q; ++q) {
    auto qp = std::make_unique<Qos>();
    qp->deleted = !qt->valid[q];
    qp->max_jobs_per_user = qt->max_jobs_per_user[q];
    qp->max_jobs_per_account = qt->max_jobs_per_account[q];
    qp->max_jobs = qt->max_jobs[q];
    qp->max_cpus_per_user = cpu_t::from_raw_value(qt->max_cpus_per_user_raw[q]);
    qp->max_wall = absl::Seconds(qt->max_wall[q]);
    qp->max_tres = LimitFromAbi(dict, qt->max_tres[q]);
    qp->max_tres_per_user = LimitFromAbi(dict, qt->max_tres_per_user[q]);
    qp->max_tres_per_account = LimitFromAbi(dict, qt->max_tres_per_account[q]);
    g_account_manager->qos_map[QosName(q)] = std::move(qp);
  }
  AccountMetaContainer amc;
  for (uint32_t u = 0; u < U; ++u)
    for (uint32_t q = 0; q < Q; ++q) amc.m_user_meta_map_[UserName(u)][QosName(q)] = MetaFromAbi(dict, qt->user_usage[(size_t)u * Q + q]);
  for (uint32_t a = 0; a < A; ++a)
    for (uint32_t q = 0; q < Q; ++q) amc.m_account_meta_map_[AcctName(a)][QosName(q)] = MetaFromAbi(dict, qt->account_usage[(size_t)a * Q + q]);
  for (uint32_t q = 0; q < Q; ++q) amc.m_qos_meta_map_[QosName(q)] = MetaFromAbi(dict, qt->qos_usage[q]);

  for (uint32_t i = 0; i < N; ++i) {
    if (pl->reason[i] != CRANE_REASON_NONE || pl->n_alloc[i] == 0) continue;
    JobInCtld j;
    j.job_id = i + 1;
    j.time_limit = absl::Seconds(pending->time_limit[i]);
    j.qos = pending->qos[i] < Q ? QosName(pending->qos[i]) : std::string("no-such-qos");
    j.username = UserName(pending->user[i]);
    for (uint32_t c = qt->chain_off[i]; c < qt->chain_off[i + 1]; ++c) j.account_chain.push_back(AcctName(qt->chain_acct[c]));
    PdJobInScheduler job(&j);
    for (uint32_t k = pl->alloc_off[i]; k < pl->alloc_off[i] + pl->n_alloc[i]; ++k)
      job.allocated_res.AddResourceInNode(NodeName(pl->alloc_node[k]), FromAbi(dict, pl->alloc_res[k]));
    auto r = amc.CheckAndMallocQosResource(job);
    if (!r) pl->reason[i] = (uint8_t)QosReasonCode(r.error());
  }
  for (uint32_t u = 0; u < U; ++u)
    for (uint32_t q = 0; q < Q; ++q) MetaToAbi(dict, amc.m_user_meta_map_[UserName(u)][QosName(q)], &qt->user_usage[(size_t)u * Q + q]);
  for (uint32_t a = 0; a < A; ++a)
    for (uint32_t q = 0; q < Q; ++q) MetaToAbi(dict, amc.m_account_meta_map_[AcctName(a)][QosName(q)], &qt->account_usage[(size_t)a * Q + q]);
  for (uint32_t q = 0; q < Q; ++q) MetaToAbi(dict, amc.m_qos_meta_map_[QosName(q)], &qt->qos_usage[q]);
  return CRANE_OK;
}
