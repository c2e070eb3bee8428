// ctld_shim2.h — second half of the shims: needs the reference's resource types
// (ph_h.inc) to be declared already. TEST INFRASTRUCTURE ONLY.
#pragma once

namespace Ctld {

// JobInCtld (CtldPublicDefs.h:737): the accessors/fields the constructors at
// JobScheduler.h:75-88 and :140-163 read.
struct JobInCtld {
  job_id_t job_id{0};
  absl::Duration time_limit;
  PartitionId partition_id;
  std::string reservation;
  ResourceView req_node_res_view, req_task_res_view, req_total_res_view;
  uint32_t node_num{0}, ntasks_per_node_min{0}, ntasks_per_node_max{0}, ntasks{0};
  bool exclusive{false};
  std::unordered_set<std::string> included_nodes, excluded_nodes;
  absl::Time submit_time, start_time, end_time;
  uint32_t partition_priority{0}, qos_priority{0};
  std::string account, qos, username;
  double mandated_priority{0.0};
  std::list<std::string> account_chain;
  ResourceV3 allocated_res;
  ResourceView allocated_res_view;
  crane::grpc::JobToCtld job_to_ctld;

  job_id_t JobId() const { return job_id; }
  absl::Time SubmitTime() const { return submit_time; }
  absl::Time StartTime() const { return start_time; }
  absl::Time EndTime() const { return end_time; }
  const ResourceV3& AllocatedRes() const { return allocated_res; }
  const crane::grpc::JobToCtld& JobToCtld() const { return job_to_ctld; }
  const std::string& Username() const { return username; }
};

struct PdJobInScheduler;

// util::Synchronized-like wrappers (crane/Lock.h, crane/Pointer.h): no locking here
template <class T>
struct ExclPtr {
  T* p;
  T* operator->() const { return p; }
  T& operator*() const { return *p; }
  explicit operator bool() const { return p != nullptr; }
};
template <class T>
struct Guarded {
  mutable T v;
  ExclPtr<T> GetExclusivePtr() const { return {&v}; }
};

// Node/NodeDefs.h:57-121, reduced to what NodeSelect reads. craned_ids are ORDERED
// sets here (the reference uses unordered sets): insertion of NodeStates then
// follows node-name == node-index order (deviation D1).
struct CranedMeta {
  bool alive{false};
  bool drain{false};
  ResourceInNodeV3 res_total;
};
struct PartitionMeta {
  std::set<CranedId> craned_ids;
};
struct ResvMeta {
  absl::Time start_time, end_time;
  std::set<CranedId> craned_ids;
  ResourceV3 res_total;
};

struct CranedMetaContainer {
  std::map<PartitionId, Guarded<PartitionMeta>> partitions;
  std::map<CranedId, Guarded<CranedMeta>> craneds;
  std::map<ResvId, Guarded<ResvMeta>> resvs;
  const std::map<PartitionId, Guarded<PartitionMeta>>* GetAllPartitionsMetaMapConstPtr() const { return &partitions; }
  const std::map<CranedId, Guarded<CranedMeta>>* GetCranedMetaMapConstPtr() const { return &craneds; }
  const std::map<ResvId, Guarded<ResvMeta>>* GetResvMetaMapPtr() const { return &resvs; }
};

// Account/AccountDefs.h:27-50: the fields NodeSelect (deleted, preempt) and
// AccountMetaContainer::CheckQosResource_ (the limits) read
struct Qos {
  bool deleted = false;
  std::set<std::string> preempt;
  uint32_t max_jobs_per_user{0};
  uint32_t max_jobs_per_account{0};
  cpu_t max_cpus_per_user{};
  uint32_t max_jobs{0};
  absl::Duration max_wall;
  ResourceView max_tres;
  ResourceView max_tres_per_user;
  ResourceView max_tres_per_account;
};
struct AccountManager {
  std::map<std::string, std::unique_ptr<Qos>> qos_map;
  const std::map<std::string, std::unique_ptr<Qos>>* GetAllQosInfo() const { return &qos_map; }
  // AccountManager::GetExistedQosInfo: the qos if it exists and is not deleted
  const Qos* GetExistedQosInfo(const std::string& name) const {
    auto it = qos_map.find(name);
    return it != qos_map.end() && !it->second->deleted ? it->second.get() : nullptr;
  }
};
struct LicensesManager {
  void CheckLicenseCountSufficient(std::vector<PdJobInScheduler*>*) {}  // no licenses in scope
};
struct JobSchedulerStub {
  std::vector<job_id_t> preempt_cancelled;
  void EnqueuePreemptCancel(std::vector<job_id_t> ids) {
    preempt_cancelled.insert(preempt_cancelled.end(), ids.begin(), ids.end());
  }
};

// CtldPublicDefs.h:151-163, 176-181, 231
struct Config {
  struct Priority {
    enum TypeEnum { Basic, MultiFactor };
    TypeEnum Type{Basic};
    bool FavorSmall{true};
    uint64_t MaxAge{0};
    uint32_t WeightAge{0}, WeightFairShare{0}, WeightJobSize{0}, WeightPartition{0}, WeightQoS{0};
  };
  struct PreemptConfig {
    crane::grpc::PreemptType PreemptType{crane::grpc::PreemptType::PREEMPT_NONE};
  };
  Priority PriorityConfig;
  PreemptConfig Preempt;
  uint32_t ScheduledBatchSize{0};
};

}  // namespace Ctld

inline Ctld::Config g_config;
inline std::unique_ptr<Ctld::CranedMetaContainer> g_meta_container;
inline std::unique_ptr<Ctld::AccountManager> g_account_manager;
inline std::unique_ptr<Ctld::LicensesManager> g_license_manager;
inline std::unique_ptr<Ctld::JobSchedulerStub> g_job_scheduler;
