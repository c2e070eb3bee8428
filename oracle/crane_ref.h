/*
 * crane_ref.h — C interface of oracle/_ref/libcrane_ref.so: the REFERENCE's own
 * SchedulerAlgo::NodeSelect, compiled from its unmodified source text
 * (oracle/ref_build.py, oracle/ref_shim/). TEST INFRASTRUCTURE ONLY: only
 * tests/ and bench.py's reference / cpu_baseline legs may load it.
 */
#ifndef CRANE_REF_H_
#define CRANE_REF_H_

#include "../include/crane_sched.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Inputs of NodeSelect that crane_pending_t / crane_running_t do not carry.
 * Every pointer may be NULL (= feature unused). */
typedef struct crane_ref_extra {
  /* reservations (ResvMeta, Node/NodeDefs.h:81-97; JobScheduler.cpp:5655-5713) */
  uint32_t n_resv;
  const int64_t* resv_start;              /* [n_resv]                               */
  const int64_t* resv_end;                /* [n_resv]                               */
  const uint32_t* resv_off;               /* [n_resv+1] CSR over (node, res) pairs  */
  const uint32_t* resv_node;
  const crane_res_in_node_t* resv_res;    /* ResvMeta::res_total per node           */
  const uint32_t* pd_resv;                /* [n_pending] reservation id or 0xFFFFFFFF */
  const uint32_t* rn_resv;                /* [n_running] reservation id or 0xFFFFFFFF */
  /* QoS preemption (JobScheduler.cpp:5566-5577, 5414-5541) */
  uint32_t n_qos;
  const uint8_t* preempt_qos;             /* [n_qos][n_qos] row q: qos ids q may preempt */
  const uint32_t* rn_qos;                 /* [n_running]                            */
  uint8_t* preempted_running;             /* out [n_running]: EnqueuePreemptCancel'd */
} crane_ref_extra_t;

/* Same contract as crane_oracle_node_select (oracle/crane_oracle.h); `extra`
 * may be NULL. CRANE_ENOSYS when cfg asks for something the reference fixes at
 * compile time (max_jobs_per_node != 1000, max_time_window_s != 7 d) or does
 * not have (cost_policy != 0). */
int crane_ref_node_select(const crane_sched_config_t* cfg, const crane_cluster_t* cluster, int64_t now,
                          const crane_running_t* running, const crane_pending_t* pending,
                          const crane_ref_extra_t* extra, crane_placements_t* out, double* elapsed_ms);

int crane_ref_feasible(const crane_cluster_t* dict, const crane_res_view_t* req,
                       const crane_res_in_node_t* avail, crane_res_in_node_t* alloc);
void crane_ref_ckmin(const crane_cluster_t* dict, crane_res_in_node_t* a, const crane_res_in_node_t* b);
int crane_ref_res_le(const crane_cluster_t* dict, const crane_res_in_node_t* a, const crane_res_in_node_t* b);
/* AccountMetaContainer::CheckAndMallocQosResource (Accounting/AccountMetaContainer.cpp:164-191, 382-531,
 * 546-587; the reference's own text) over the jobs a tick starts; same contract as crane_oracle_qos_filter. */
int crane_ref_qos_filter(const crane_cluster_t* dict, const crane_pending_t* pending,
                         crane_placements_t* placements, const crane_qos_table_t* qos);
const char* crane_ref_describe(void);

#ifdef __cplusplus
}
#endif
#endif
