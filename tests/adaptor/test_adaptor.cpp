// test_adaptor.cpp — drives the NodeSelect-shaped adaptor with reference-style
// objects (hostnames, gres names/types, device paths) and checks the result
// against the oracle fed with the same scenario built directly as C-ABI tables.
// Linked against the product library (GPU) or the kernel-emulation library
// (CPU tests) plus the oracle.
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <climits>
#include <cstdint>
#include <set>
#include <string>

#include "../../cranesched_b200/adaptor/crane_adaptor.h"
#include "../../oracle/crane_oracle.h"

using namespace crane_b200;

static std::string NodeName(int i) { char b[32]; snprintf(b, sizeof b, "cn%04d", i); return b; }
static std::string Slot(int i) { char b[32]; snprintf(b, sizeof b, "/dev/nvidia%d", i); return b; }

int main() {
  const int64_t now = 1700000000;
  std::mt19937 rng(12345);
  auto U = [&](int lo, int hi) { return (int)(rng() % (uint32_t)(hi - lo + 1)) + lo; };
  const int M = 24, N = 160;
  // ---- cluster: two partitions; "gpu" nodes carry a100 or h100 devices -------
  std::map<CranedId, CranedMeta> craneds;
  std::map<PartitionId, std::set<CranedId>> parts;
  std::vector<crane_res_in_node_t> totals(M);
  std::vector<uint8_t> alive(M, 1), drain(M, 0);
  // dictionary order: (gpu,a100)=0, (gpu,h100)=1
  for (int i = 0; i < M; ++i) {
    CranedMeta meta;
    const int cores = i < 12 ? 16 : 32;
    for (int c = 0; c < cores; ++c) meta.res_total.core_ids.insert(c);
    meta.res_total.cpu_count_raw = cores * 256;
    meta.res_total.memory_bytes = meta.res_total.memory_sw_bytes = (uint64_t)cores * 4 << 30;
    crane_res_in_node_t& t = totals[i];
    memset(&t, 0, sizeof t);
    t.cpu_raw = cores * 256;
    t.mem = t.mem_sw = (uint64_t)cores * 4 << 30;
    t.core[0] = cores == 32 ? 0xFFFFFFFFull : 0xFFFFull;
    if (i >= 12) {
      const char* type = (i % 2) ? "h100" : "a100";
      for (int s = 0; s < 4; ++s) meta.res_total.gres["gpu"][type].insert(Slot(s));
      t.gres[(i % 2) ? 1 : 0] = 0xF;
    }
    if (i == 5) { meta.drain = true; drain[i] = 1; }
    craneds[NodeName(i)] = meta;
    parts[i < 12 ? "cpu" : "gpu"].insert(NodeName(i));
  }
  PriorityConfig prio;
  prio.WeightAge = 500; prio.WeightPartition = 1000; prio.WeightQoS = 1000000; prio.WeightJobSize = 300;
  SchedulerAlgo algo(prio, 100000, 0);
  algo.SetCluster(craneds, parts);

  // ---- jobs, in both representations --------------------------------------------
  std::vector<std::unique_ptr<PdJobInScheduler>> pending;
  std::vector<std::unique_ptr<RnJobInScheduler>> running;
  std::vector<uint32_t> partition(N), node_num(N), ntasks(N), ntpn(N), pp(N), qp(N), acc(N), zero(N, 0), incl_off(N + 1, 0), excl_off(N + 1, 0), incl, excl;
  std::vector<int64_t> tl(N), sub(N);
  std::vector<uint8_t> excl_flag(N, 0);
  std::vector<double> mand(N, 0.0);
  std::vector<crane_res_view_t> rn_(N), rt(N), rtot(N);
  std::map<std::string, uint32_t> acc_id, user_id;
  std::vector<uint32_t> user(N);
  for (int i = 0; i < N; ++i) {
    auto j = std::make_unique<PdJobInScheduler>();
    const bool gpu = U(0, 2) == 0;
    j->job_id = i + 1;
    j->partition_id = i == 17 ? "nosuch" : (gpu ? "gpu" : "cpu");
    j->time_limit = U(60, 20000);
    j->submit_time = now - U(0, 500000);
    j->node_num = U(0, 9) == 0 ? 2 : 1;
    j->ntasks = j->node_num;
    j->partition_priority = gpu ? 2000 : 1000;
    j->qos_priority = U(0, 1) ? 1000 : 5000;
    j->account = "acct" + std::to_string(U(0, 3));
    j->qos = "normal";
    j->username = "user" + std::to_string(U(0, 5));
    const int cpus = 1 << U(0, 3);
    j->req_task_res_view.cpu_count_raw = cpus * 256;
    j->req_task_res_view.memory_bytes = j->req_task_res_view.memory_sw_bytes = (uint64_t)cpus << 30;
    memset(&rn_[i], 0, sizeof rn_[i]);
    memset(&rt[i], 0, sizeof rt[i]);
    rt[i].cpu_raw = cpus * 256;
    rt[i].mem = rt[i].mem_sw = (uint64_t)cpus << 30;
    if (gpu) {
      const int g = U(1, 3);
      GresCount gc;
      gc.total = g;
      rn_[i].gres_total[0] = g;
      if (U(0, 1)) { const bool h = U(0, 1); gc.specified[h ? "h100" : "a100"] = 1; rn_[i].gres_spec[h ? 1 : 0] = 1; }
      j->req_node_res_view.gres_map["gpu"] = gc;
    }
    if (U(0, 11) == 0 && !gpu) { j->excluded_nodes.insert(NodeName(U(0, 11))); }
    // total view = node*node_num + task*ntasks (JobScheduler.cpp:6109)
    j->req_total_res_view.cpu_count_raw = j->req_task_res_view.cpu_count_raw * j->ntasks;
    j->req_total_res_view.memory_bytes = j->req_task_res_view.memory_bytes * j->ntasks;
    rtot[i] = rt[i];
    rtot[i].cpu_raw *= j->ntasks; rtot[i].mem *= j->ntasks; rtot[i].mem_sw *= j->ntasks;
    for (int g = 0; g < 8; ++g) { rtot[i].gres_total[g] = rn_[i].gres_total[g] * j->node_num; rtot[i].gres_spec[g] = rn_[i].gres_spec[g] * j->node_num; }
    partition[i] = j->partition_id == "cpu" ? 0 : j->partition_id == "gpu" ? 1 : 2;
    node_num[i] = j->node_num; ntasks[i] = j->ntasks; ntpn[i] = 1; pp[i] = j->partition_priority; qp[i] = j->qos_priority;
    if (!acc_id.count(j->account)) { uint32_t id = (uint32_t)acc_id.size(); acc_id[j->account] = id; }
    acc[i] = acc_id[j->account];
    if (!user_id.count(j->username)) { uint32_t id = (uint32_t)user_id.size(); user_id[j->username] = id; }
    user[i] = user_id[j->username];
    // account chain: the job's account, its department, the root (PdJobInScheduler::account_chain)
    j->account_chain = {j->account, std::string("dept") + (j->account.back() < '2' ? "A" : "B"), "root"};
    tl[i] = j->time_limit; sub[i] = j->submit_time;
    for (const auto& id : j->excluded_nodes) excl.push_back((uint32_t)atoi(id.c_str() + 2));
    excl_off[i + 1] = (uint32_t)excl.size();
    pending.push_back(std::move(j));
  }
  algo.NodeSelect(now, running, pending);

  // ---- oracle on the directly-built tables ----------------------------------------
  crane_sched_config_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.priority_type = 1; cfg.favor_small = 1; cfg.max_age_s = 7 * 24 * 3600; cfg.weight_age = 500; cfg.weight_partition = 1000;
  cfg.weight_qos = 1000000; cfg.weight_job_size = 300; cfg.scheduled_batch_size = 100000; cfg.max_jobs_per_node = 1000; cfg.max_time_window_s = 7 * 24 * 3600;
  crane_cluster_t cl;
  memset(&cl, 0, sizeof cl);
  std::vector<uint32_t> part_off{0, 12, 24}, part_nodes(M);
  for (int i = 0; i < M; ++i) part_nodes[i] = i;
  cl.n_nodes = M; cl.res_total = totals.data(); cl.alive = alive.data(); cl.drain = drain.data(); cl.n_partitions = 2;
  cl.part_off = part_off.data(); cl.part_nodes = part_nodes.data(); cl.n_gres_entries = 2;
  crane_pending_t pd;
  memset(&pd, 0, sizeof pd);
  pd.n = N; pd.partition = partition.data(); pd.time_limit = tl.data(); pd.submit_time = sub.data(); pd.node_num = node_num.data();
  pd.ntasks = ntasks.data(); pd.ntasks_per_node_min = ntpn.data(); pd.ntasks_per_node_max = ntpn.data(); pd.exclusive = excl_flag.data();
  pd.partition_priority = pp.data(); pd.qos_priority = qp.data(); pd.account = acc.data(); pd.qos = zero.data(); pd.user = user.data();
  pd.mandated_priority = mand.data(); pd.req_node = rn_.data(); pd.req_task = rt.data(); pd.req_total = rtot.data();
  pd.incl_off = incl_off.data(); pd.incl_nodes = incl.data(); pd.excl_off = excl_off.data(); pd.excl_nodes = excl.data();
  crane_running_t rnj;
  memset(&rnj, 0, sizeof rnj);
  uint32_t roff0 = 0;
  rnj.alloc_off = &roff0;
  uint64_t total = 0;
  for (int i = 0; i < N; ++i) total += node_num[i];
  std::vector<uint8_t> o_reason(N); std::vector<double> o_prio(N); std::vector<int64_t> o_s(N), o_e(N);
  std::vector<uint32_t> o_n(N), o_off(N + 1), o_node(total), o_nt(total); std::vector<crane_res_in_node_t> o_res(total);
  crane_placements_t out{o_reason.data(), o_prio.data(), o_s.data(), o_e.data(), o_n.data(), o_off.data(), o_node.data(), o_nt.data(), o_res.data()};
  double ms; uint32_t done;
  if (crane_oracle_node_select(&cfg, &cl, now, &rnj, &pd, &out, &ms, 0, &done) != 0) { printf("oracle failed\n"); return 2; }

  // ---- compare -----------------------------------------------------------------------
  const char* reason_str[] = {"", "Priority", "Resource", "Resource Reserved", "Partition Not Found"};
  int bad = 0, started = 0, reserved = 0;
  for (int i = 0; i < N; ++i) {
    const PdJobInScheduler& j = *pending[i];
    bool ok = j.reason == reason_str[o_reason[i]] && memcmp(&j.priority, &o_prio[i], 8) == 0 && j.start_time == o_s[i] && j.end_time == o_e[i] &&
              j.craned_ids.size() == o_n[i];
    for (uint32_t k = 0; ok && k < o_n[i]; ++k) {
      const std::string id = NodeName((int)o_node[o_off[i] + k]);
      ok = j.craned_ids[k] == id && j.craned_id_to_task_num.at(id) == o_nt[o_off[i] + k];
      const ResourceInNodeV3& r = j.allocated_res.at(id);
      const crane_res_in_node_t& e = o_res[o_off[i] + k];
      ok = ok && r.cpu_count_raw == e.cpu_raw && r.memory_bytes == e.mem;
      uint64_t cm = 0;
      for (uint32_t c : r.core_ids) cm |= 1ull << c;
      ok = ok && cm == e.core[0];
      for (int g = 0; g < 2 && ok; ++g) {
        uint16_t mask = 0;
        auto nit = r.gres.find("gpu");
        if (nit != r.gres.end()) {
          auto tit = nit->second.find(g ? "h100" : "a100");
          if (tit != nit->second.end())
            for (const auto& s : tit->second) mask |= (uint16_t)(1u << atoi(s.c_str() + 11));
        }
        ok = mask == e.gres[g];
      }
    }
    if (!ok) { if (bad < 5) printf("MISMATCH job %d: reason '%s' vs '%s' start %ld vs %ld\n", i, j.reason.c_str(), reason_str[o_reason[i]], (long)j.start_time, (long)o_s[i]); ++bad; }
    started += j.is_scheduled();
    reserved += !j.is_scheduled() && !j.craned_ids.empty();
  }
  printf("adaptor: %d jobs, %d start now, %d reserved, %d mismatches vs oracle\n", N, started, reserved, bad);
  if (bad) return 1;

  // ---- the QoS stage behind NodeSelect (JobScheduler.cpp:1262) ---------------------------
  // adaptor: reference-style objects; oracle: the same limits and usage as C-ABI tables
  std::map<std::string, Qos> qos_table;
  Qos normal;
  normal.max_jobs_per_user = 9;
  normal.max_jobs_per_account = 14;
  normal.max_cpus_per_user_raw = 40 * 256;
  normal.max_wall = 400000;
  normal.max_tres_per_account.cpu_count_raw = 70 * 256;
  normal.max_tres.memory_bytes = 300ull << 30;
  normal.max_tres_per_user.gres_map["gpu"].total = 6;
  normal.max_tres_per_user.gres_map["gpu"].specified["h100"] = 2;
  qos_table["normal"] = normal;
  QosUsage usage;
  usage.user["user1"]["normal"].jobs_count = 7;            // close to its job limit before the pass
  usage.user["user1"]["normal"].resource.cpu_count_raw = 8 * 256;
  usage.account["root"]["normal"].wall_time = 1000;
  algo.CheckAndMallocQosResource(qos_table, usage, pending);

  // account ids: job accounts in first-appearance order (as for the fair-share column), then chain parents
  for (int i = 0; i < N; ++i)
    for (const auto& a : pending[i]->account_chain)
      if (!acc_id.count(a)) { uint32_t id = (uint32_t)acc_id.size(); acc_id[a] = id; }
  const uint32_t Q = 1, NU = (uint32_t)user_id.size(), NA = (uint32_t)acc_id.size();
  std::vector<uint32_t> chain_off(N + 1, 0), chain;
  for (int i = 0; i < N; ++i) {
    for (const auto& a : pending[i]->account_chain) chain.push_back(acc_id[a]);
    chain_off[i + 1] = (uint32_t)chain.size();
  }
  uint8_t valid = 1;
  crane_tres_limit_t tu, ta, tq;
  memset(&tu, 0, sizeof tu); memset(&ta, 0, sizeof ta); memset(&tq, 0, sizeof tq);
  tu.view.cpu_raw = ta.view.cpu_raw = tq.view.cpu_raw = INT64_MAX;
  tu.view.mem = ta.view.mem = tq.view.mem = UINT64_MAX;
  tu.view.mem_sw = ta.view.mem_sw = tq.view.mem_sw = UINT64_MAX;
  ta.view.cpu_raw = 70 * 256;
  tq.view.mem = 300ull << 30;
  tu.gres_name_present = 1; tu.view.gres_total[0] = 6;
  tu.gres_spec_present = 2; tu.view.gres_spec[1] = 2;
  std::vector<crane_meta_resource_t> uu(NU), au(NA), qu(Q);
  memset(uu.data(), 0, sizeof(crane_meta_resource_t) * NU);
  memset(au.data(), 0, sizeof(crane_meta_resource_t) * NA);
  memset(qu.data(), 0, sizeof(crane_meta_resource_t) * Q);
  uu[user_id["user1"]].jobs_count = 7;
  uu[user_id["user1"]].cpu_raw = 8 * 256;
  au[acc_id["root"]].wall_time = 1000;
  crane_qos_table_t qt;
  memset(&qt, 0, sizeof qt);
  qt.n_qos = Q; qt.n_users = NU; qt.n_accounts = NA; qt.valid = &valid;
  qt.max_jobs_per_user = &normal.max_jobs_per_user; qt.max_jobs_per_account = &normal.max_jobs_per_account; qt.max_jobs = &normal.max_jobs;
  qt.max_cpus_per_user_raw = &normal.max_cpus_per_user_raw; qt.max_wall = &normal.max_wall;
  qt.max_tres_per_user = &tu; qt.max_tres_per_account = &ta; qt.max_tres = &tq;
  qt.chain_off = chain_off.data(); qt.chain_acct = chain.data();
  qt.user_usage = uu.data(); qt.account_usage = au.data(); qt.qos_usage = qu.data();
  if (crane_oracle_qos_filter(&cl, &pd, &out, &qt) != 0) { printf("oracle qos filter failed\n"); return 2; }
  auto qos_str = [&](uint8_t c) -> std::string {
    switch (c) {
      case 16: return "QosCpuResourceLimit"; case 17: return "QosJobsResourceLimit"; case 18: return "QosWallTimeLimit";
      case 19: return "QosMemResourceLimit"; case 20: return "QosGresResourceLimit"; case 21: return "InvalidQOS";
      default: return reason_str[c];
    }
  };
  int qbad = 0, refused = 0;
  std::set<std::string> kinds;
  for (int i = 0; i < N; ++i) {
    if (pending[i]->reason != qos_str(o_reason[i])) { if (qbad < 5) printf("QOS MISMATCH job %d: '%s' vs '%s'\n", i, pending[i]->reason.c_str(), qos_str(o_reason[i]).c_str()); ++qbad; }
    if (o_reason[i] >= 16) { ++refused; kinds.insert(qos_str(o_reason[i])); }
  }
  auto same = [&](const MetaResource& a, const crane_meta_resource_t& b) {
    uint64_t gt = 0, a100 = 0, h100 = 0;
    auto it = a.resource.gres_map.find("gpu");
    if (it != a.resource.gres_map.end()) {
      gt = it->second.total;
      if (it->second.specified.count("a100")) a100 = it->second.specified.at("a100");
      if (it->second.specified.count("h100")) h100 = it->second.specified.at("h100");
    }
    return a.resource.cpu_count_raw == b.cpu_raw && a.resource.memory_bytes == b.mem && a.jobs_count == b.jobs_count && a.wall_time == b.wall_time &&
           gt == b.gres_total[0] && a100 == b.gres_spec[0] && h100 == b.gres_spec[1];
  };
  for (const auto& [name, id] : user_id) if (!same(usage.user[name]["normal"], uu[id])) { printf("QOS usage of user %s differs\n", name.c_str()); ++qbad; }
  for (const auto& [name, id] : acc_id) if (!same(usage.account[name]["normal"], au[id])) { printf("QOS usage of account %s differs\n", name.c_str()); ++qbad; }
  if (!same(usage.qos["normal"], qu[0])) { printf("QOS usage of the qos differs\n"); ++qbad; }
  printf("adaptor qos: %d refused (%zu kinds), %d mismatches vs oracle\n", refused, kinds.size(), qbad);
  return qbad || refused == 0 ? 1 : 0;
}
