// ctld_shim3.h — what the sliced AccountMetaContainer text (amc_h.inc / amc_cpp.inc) expects
// around it: the class declaration reduced to the members those definitions touch
// (Accounting/AccountMetaContainer.h:49-170) and phmap::parallel_flat_hash_map's three
// calls as an ordered map. Included inside namespace Ctld, after the scheduler slices
// (PdJobInScheduler). TEST INFRASTRUCTURE ONLY.
#pragma once
// (<array>, <expected>, <mutex> come from ctld_shim.h: this file is included inside namespace Ctld)

template <class K, class V>
struct PhMapShim : std::map<K, V> {
  bool contains(const K& k) const { return this->find(k) != this->end(); }
  template <class F>
  bool if_contains(const K& k, F&& f) {
    auto it = this->find(k);
    if (it == this->end()) return false;
    f(*it);
    return true;
  }
  template <class F, class... A>
  bool try_emplace_l(const K& k, F&& f, A&&... a) {  // phmap: f on the existing entry, else construct
    auto it = this->find(k);
    if (it != this->end()) { f(*it); return false; }
    this->emplace(k, V(std::forward<A>(a)...));
    return true;
  }
};

class AccountMetaContainer final {
 public:
  using QosToResourceMap = std::unordered_map<std::string, MetaResource>;
  using ResourceMetaMap = PhMapShim<std::string, QosToResourceMap>;
  using QosResourceMap = PhMapShim<std::string, MetaResource>;

  std::expected<void, std::string> CheckAndMallocQosResource(const PdJobInScheduler& job);

  const static int kNumStripes = 128;
  static int StripeForKey_(const std::string& key) { return std::hash<std::string>{}(key) % kNumStripes; }
  std::expected<void, std::string> CheckQosResource_(const Qos& qos, const PdJobInScheduler& job);
  static std::expected<void, std::string> CheckTres_(const ResourceView& resource_req, const ResourceView& resource_total);
  static bool CheckGres_(const GresMap& device_req, const GresMap& device_total);
  std::vector<std::unique_lock<std::mutex>> LockAccountStripes_(const std::list<std::string>& account_chain);
  void DoMallocResource_(job_id_t job_id, const std::string& username, const std::list<std::string>& account_chain,
                         const std::string& qos, const MetaResource& meta_resource);

  ResourceMetaMap m_user_meta_map_;
  ResourceMetaMap m_account_meta_map_;
  QosResourceMap m_qos_meta_map_;
  std::array<std::mutex, kNumStripes> m_user_stripes_;
  std::array<std::mutex, kNumStripes> m_account_stripes_;
  std::array<std::mutex, kNumStripes> m_qos_stripes_;
  void UserAddJob(const std::string&) {}
};
