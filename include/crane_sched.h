/*
 * crane_sched.h — C-ABI of the B200-native CraneCtld scheduling hot path.
 *
 * This is the drop-in boundary for ONE call in the reference daemon:
 *
 *     m_node_selection_algo_->NodeSelect(now, running_jobs, pending_jobs);
 *         -- /root/reference/src/CraneCtld/JobScheduler.cpp:1141
 *         -- declared at src/CraneCtld/JobScheduler.h:254-257
 *
 * The reference has no FFI for this path (it is an in-process C++ virtual
 * interface); the entry points below are what a cgo/FFI-free
 * C++ caller (the NodeSelect-shaped adaptor in cranesched_b200/adaptor/) binds.
 * Everything is plain-old-data: pointers + sizes, no C++/torch types.
 *
 * Conventions
 *   - times are int64 unix seconds (the reference truncates `now` to 1 s,
 *     JobScheduler.cpp:1071); "unset" start/end time is 0.
 *   - cpu amounts are the raw value of the reference's
 *     cpu_t = fpm::fixed<int64_t,__int128,8> (PublicHeader.h:44): cpus * 256.
 *   - nodes, partitions, accounts, qos, users are dense indices. Node index
 *     order is the documented tie-break for equal-cost nodes (the reference
 *     uses heap-pointer order, JobScheduler.h:588).
 *   - sets of core ids / gres slot ids (std::set in the reference,
 *     PublicHeader.h:412-479,540-558) are bit masks: core bit i = core id i;
 *     slot bit i of gres entry e = the i-th slot (lexicographic slot-path
 *     order) of that (name,type) on that node.
 *   - all functions return 0 on success or a negative CRANE_E* code; they
 *     never throw and never fall back to a CPU implementation.
 */
#ifndef CRANE_SCHED_H_
#define CRANE_SCHED_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRANE_CORE_WORDS 4   /* <= 256 cores per node                       */
#define CRANE_GRES_ENTRIES 8 /* (name,type) dictionary entries, cluster-wide */
#define CRANE_GRES_NAMES 8   /* distinct gres names                          */
#define CRANE_MAX_SLOTS 16   /* slots per (name,type) per node               */

/* error codes */
#define CRANE_OK 0
#define CRANE_EINVAL (-22)   /* malformed input (see crane_sched_last_error)  */
#define CRANE_ENOMEM (-12)
#define CRANE_ENODEV (-19)   /* no CUDA device / kernel image not loadable    */
#define CRANE_ECUDA (-5)     /* CUDA runtime error during the call            */
#define CRANE_ENOSYS (-38)   /* feature outside the built scope (see DESIGN)  */

/* Mirrors ResourceInNodeV3 (PublicHeader.h:562-615): concrete per-node
 * resources. 72 bytes. */
typedef struct crane_res_in_node {
  int64_t cpu_raw;                    /* CpuSet::cpu_count raw               */
  uint64_t mem;                       /* bytes                               */
  uint64_t mem_sw;                    /* bytes                               */
  uint64_t core[CRANE_CORE_WORDS];    /* CpuSet::core_ids                    */
  uint16_t gres[CRANE_GRES_ENTRIES];  /* slot masks per dictionary entry     */
} crane_res_in_node_t;

/* Mirrors ResourceView (PublicHeader.h:671-737): counts only. 56 bytes.
 * gres_total[g] = GresCount::total of name g; gres_spec[e] =
 * GresCount::specified[type of entry e]. A name with total==0 and no
 * specified count is "not requested". */
typedef struct crane_res_view {
  int64_t cpu_raw;
  uint64_t mem;
  uint64_t mem_sw;
  uint16_t gres_total[CRANE_GRES_NAMES];
  uint16_t gres_spec[CRANE_GRES_ENTRIES];
} crane_res_view_t;

/* Scheduler knobs: Config::Priority (CtldPublicDefs.h:151-163),
 * ScheduledBatchSize (:231), and the constants at JobScheduler.h:262-264. */
typedef struct crane_sched_config {
  uint32_t priority_type; /* 0 = BasicPriority (JS.h:177), 1 = MultiFactor   */
  uint32_t favor_small;
  uint64_t max_age_s;
  uint32_t weight_age;
  uint32_t weight_fair_share;
  uint32_t weight_job_size;
  uint32_t weight_partition;
  uint32_t weight_qos;
  uint32_t scheduled_batch_size; /* `limit`, JS.cpp:5770                    */
  uint32_t max_jobs_per_node;    /* kAlgoMaxJobNumPerNode (1000)            */
  uint32_t cost_policy;          /* IUpdateNodeCostPolicy (JS.h:30-54): 0 = MinCpuTimeRatioFirst
                                    (the reference's only policy), 1 = BestFit: cost = the node's free
                                    cpu count over all allocations on it, fullest node first — not in
                                    the reference; BASELINE config 4 ("best-fit selection")   */
  int64_t max_time_window_s;     /* kAlgoMaxTimeWindow (7 d)                */
} crane_sched_config_t;

/* Node table + partition membership: what NodeSelect reads from
 * g_meta_container (JobScheduler.cpp:5597-5651; CranedMeta NodeDefs.h:57-79;
 * PartitionMeta::craned_ids NodeDefs.h:118-121). */
typedef struct crane_cluster {
  uint32_t n_nodes;
  const crane_res_in_node_t* res_total; /* [n_nodes]                          */
  const uint8_t* alive;                 /* [n_nodes]                          */
  const uint8_t* drain;                 /* [n_nodes]                          */
  uint32_t n_partitions;
  const uint32_t* part_off;   /* [n_partitions+1] CSR                         */
  const uint32_t* part_nodes; /* node indices, ascending within a partition   */
  uint32_t n_gres_entries;    /* <= CRANE_GRES_ENTRIES                        */
  uint8_t gres_entry_name[CRANE_GRES_ENTRIES]; /* name id of each entry; the
        entries of one name must be contiguous and names ascending            */
} crane_cluster_t;

/* Mirrors RnJobInScheduler (JobScheduler.h:56-89). */
typedef struct crane_running {
  uint32_t n;
  const int64_t* start_time; /* [n]                                          */
  const int64_t* end_time;   /* [n]                                          */
  const uint32_t* node_num;  /* [n] (the reference reads it uninitialised,
                                JS.cpp:6614; here it is an explicit input)   */
  const uint32_t* partition_priority;
  const uint32_t* qos_priority;
  const uint32_t* account;
  const int64_t* view_cpu_raw; /* allocated_res_view.CpuCount raw            */
  const uint64_t* view_mem;    /* allocated_res_view.GetMemoryBytes          */
  const uint32_t* alloc_off;   /* [n+1] CSR over (node, res) pairs           */
  const uint32_t* alloc_node;
  const crane_res_in_node_t* alloc_res;
  const uint32_t* reservation; /* [n] RnJobInScheduler::reservation as an index into
                                  crane_reservations_t, 0xFFFFFFFF = none; NULL = no job has one */
} crane_running_t;

/* Mirrors the inputs of PdJobInScheduler (JobScheduler.h:91-164). Input order
 * is job-id order (the pending map is a btree, JS.cpp:1092). */
typedef struct crane_pending {
  uint32_t n;
  const uint32_t* partition;     /* partition index; >= n_partitions = none  */
  const int64_t* time_limit;     /* seconds, >= 1                            */
  const int64_t* submit_time;
  const uint32_t* node_num;      /* >= 1                                     */
  const uint32_t* ntasks;
  const uint32_t* ntasks_per_node_min;
  const uint32_t* ntasks_per_node_max;
  const uint8_t* exclusive;
  const uint32_t* partition_priority;
  const uint32_t* qos_priority;
  const uint32_t* account;
  const uint32_t* qos;
  const uint32_t* user;
  const double* mandated_priority;   /* 0.0 = compute (JS.h:158, JS.cpp:6535)*/
  const crane_res_view_t* req_node;  /* req_node_res_view                    */
  const crane_res_view_t* req_task;  /* req_task_res_view                    */
  const crane_res_view_t* req_total; /* req_total_res_view                   */
  const uint32_t* incl_off;  /* [n+1] CSR of included_nodes, or NULL; the index
                                0xFFFFFFFF stands for a host unknown to the cluster */
  const uint32_t* incl_nodes;
  const uint32_t* excl_off;  /* [n+1] CSR of excluded_nodes, or NULL          */
  const uint32_t* excl_nodes;
  const uint32_t* reservation; /* [n] PdJobInScheduler::reservation as an index into
                                  crane_reservations_t, 0xFFFFFFFF = none, an index >= n
                                  = a name the daemon does not know ("Reservation Not
                                  Found"); NULL = no job has one */
} crane_pending_t;

/* Reservations: what NodeSelect reads from g_meta_container->GetResvMetaMapPtr()
 * (JobScheduler.cpp:5655-5713; ResvMeta, Node/NodeDefs.h:81-97). A reservation
 * that has started is an allocation on its nodes until its end and — for the
 * pending jobs submitted into it — a scheduler of its own over node states
 * holding exactly the reserved resources, with the timeline ending at the
 * reservation's end; one that starts later is taken out of its nodes' timelines
 * over [start, end); an expired one (now >= end) is ignored. A job backfilled
 * onto a node whose first reservation starts inside the job's window is
 * labelled "Resource Reserved" (JobScheduler.cpp:5829-5841). */
typedef struct crane_reservations {
  uint32_t n;
  const int64_t* start_time;        /* [n]                                    */
  const int64_t* end_time;          /* [n]                                    */
  const uint32_t* node_off;         /* [n+1] CSR over (node, res) pairs:
                                       ResvMeta::res_total.EachNodeResMap()   */
  const uint32_t* node;             /* node index (a node appears once per reservation) */
  const crane_res_in_node_t* res;   /* resources reserved on that node        */
} crane_reservations_t;

/* pending reasons: the strings NodeSelect writes (JS.cpp:192,5784-5864,6547;
 * docs/en/reference/pending_reason.md:42-54) as codes. */
enum {
  CRANE_REASON_NONE = 0,           /* "" : start now                          */
  CRANE_REASON_PRIORITY = 1,       /* "Priority"                              */
  CRANE_REASON_RESOURCE = 2,       /* "Resource"                              */
  CRANE_REASON_RESERVED = 3,       /* "Resource Reserved"                     */
  CRANE_REASON_PART_NOT_FOUND = 4, /* "Partition Not Found"                   */
  CRANE_REASON_RESV_NOT_FOUND = 5, /* "Reservation Not Found" (JS.cpp:5793)   */
  CRANE_REASON_PREEMPTED = 6,      /* "Preempted" (JS.cpp:5815)               */
  /* written by crane_sched_qos_filter (Accounting/AccountMetaContainer.cpp:382-531) */
  CRANE_REASON_QOS_CPU = 16,       /* "QosCpuResourceLimit"                   */
  CRANE_REASON_QOS_JOBS = 17,      /* "QosJobsResourceLimit"                  */
  CRANE_REASON_QOS_WALL = 18,      /* "QosWallTimeLimit"                      */
  CRANE_REASON_QOS_MEM = 19,       /* "QosMemResourceLimit"                   */
  CRANE_REASON_QOS_GRES = 20,      /* "QosGresResourceLimit"                  */
  CRANE_REASON_QOS_INVALID = 21,   /* "InvalidQOS"                            */
  CRANE_REASON_ERASED = 255,       /* a row removed with crane_sched_pending_erase */
};

/* The fields NodeSelect writes into PdJobInScheduler (JobScheduler.h:116-132).
 * Caller-owned, pre-sized. Job i owns allocation slots
 * [alloc_off[i], alloc_off[i] + node_num[i]) with alloc_off = exclusive prefix
 * sum of pending.node_num (written by the callee). n_alloc[i] is node_num[i]
 * when the job holds a placement (start now, or a backfill reservation) and 0
 * otherwise; slots are ordered by node index ascending. */
typedef struct crane_placements {
  uint8_t* reason;        /* [n]                                             */
  double* priority;       /* [n]                                             */
  int64_t* start_time;    /* [n] 0 = unset                                   */
  int64_t* end_time;      /* [n] 0 = unset                                   */
  uint32_t* n_alloc;      /* [n]                                             */
  uint32_t* alloc_off;    /* [n+1]                                           */
  uint32_t* alloc_node;   /* [sum node_num]                                  */
  uint32_t* alloc_ntasks; /* [sum node_num] craned_id_to_task_num            */
  crane_res_in_node_t* alloc_res; /* [sum node_num] allocated_res            */
} crane_placements_t;

typedef struct crane_sched crane_sched_t; /* opaque handle, one per GPU      */

/* Per-call device timings (CUDA events on the handle's stream), ms. */
typedef struct crane_sched_timing {
  float h2d_ms;
  float init_ms;     /* node state + timeline build (R2,R3,R4)               */
  float priority_ms; /* R5/R6                                                */
  float feas_ms;     /* capability bitmap (R7 predicate ⊕ R8)                */
  float commit_ms;   /* sequential select/backfill/update (R7,R9,R10)        */
  float d2h_ms;
  float total_ms;
  uint32_t kernel_launches;
  float qos_ms;      /* device time of the last crane_sched_qos_filter (R12)     */
} crane_sched_timing_t;

/* ---- lifecycle ---------------------------------------------------------- */
/* Replaces: construction of SchedulerAlgo + IPrioritySorter
 * (JobScheduler.cpp:111-121). device = CUDA ordinal. */
int crane_sched_create(const crane_sched_config_t* cfg, int device,
                       crane_sched_t** out);
void crane_sched_destroy(crane_sched_t* h);
const char* crane_sched_last_error(const crane_sched_t* h);

/* Replaces: the node/partition snapshot NodeSelect takes from
 * g_meta_container every tick (JobScheduler.cpp:5603-5651). Call when the node
 * set, alive/drain flags or partition membership change.
 * A node may be listed by several partitions: it then has ONE state shared by
 * their LocalSchedulers (JobScheduler.cpp:5622), and every connected group of
 * such partitions is scheduled as one unit, job by job in priority order, each
 * job in its own partition's node order (at most 8 partitions per group, else
 * CRANE_ENOSYS). Partitions that share nothing keep their independent loops. */
int crane_sched_set_cluster(crane_sched_t* h, const crane_cluster_t* cluster);

/* Replaces: the reservation snapshot of NodeSelect (JobScheduler.cpp:5655-5713).
 * Call after crane_sched_set_cluster (which clears the reservations) whenever
 * the reservation set changes; resv == NULL or n == 0 = none. */
int crane_sched_set_reservations(crane_sched_t* h, const crane_reservations_t* resv);

/* ---- the hot path -------------------------------------------------------- */
/* Replaces: SchedulerAlgo::NodeSelect (JobScheduler.cpp:5543-5868), host
 * buffers in, host buffers out; blocking; single caller thread. */
int crane_sched_node_select(crane_sched_t* h, int64_t now,
                            const crane_running_t* running,
                            const crane_pending_t* pending,
                            crane_placements_t* out);

/* The same call split at the PCIe boundary, for callers that keep the pending
 * table resident and for measurement:
 *   upload  = H2D of running+pending tables (async on the handle's stream)
 *   run     = every kernel of the tick, device-resident in and out
 *   fetch   = D2H of the placements
 * node_select == upload; run; fetch. */
int crane_sched_upload(crane_sched_t* h, const crane_running_t* running,
                       const crane_pending_t* pending);
int crane_sched_run(crane_sched_t* h, int64_t now);
int crane_sched_fetch(crane_sched_t* h, crane_placements_t* out);

/* The pending table kept resident on the device across ticks:
 * the reference rebuilds its PdJobInScheduler vector from the pending map every
 * tick (JobScheduler.cpp:1090-1113); here a submit appends its row(s)
 * (JobScheduler.cpp:4254), a start or cancel erases its row, and a tick is
 *     crane_sched_set_running(h, running); crane_sched_run(h, now); crane_sched_fetch(h, out)
 * without any H2D of pending jobs. Rows keep their indices (outputs are indexed
 * by row; `out` must be sized for crane_sched_pending_rows(h) rows and the sum of
 * node_num over ALL rows); an erased row reports CRANE_REASON_ERASED and takes
 * part in nothing (priority bounds, batch limit, queue). Rows must arrive in
 * job-id order (equal priorities keep row order). crane_sched_upload ==
 * pending_reset; pending_append; set_running. Re-upload to compact. */
int crane_sched_pending_reset(crane_sched_t* h);
int crane_sched_pending_append(crane_sched_t* h, const crane_pending_t* rows, uint32_t* first_row);
int crane_sched_pending_erase(crane_sched_t* h, const uint32_t* rows, uint32_t n);
int crane_sched_set_running(crane_sched_t* h, const crane_running_t* running);
uint32_t crane_sched_pending_rows(const crane_sched_t* h);

/* Blocks until the handle's stream is idle; *run_ms (optional) receives the
 * device time of the last crane_sched_run (CUDA events on that stream). */
int crane_sched_sync(crane_sched_t* h, float* run_ms);

int crane_sched_get_timing(const crane_sched_t* h, crane_sched_timing_t* t);

/* ---- one queue over several GPUs ----------------------------- */
/* The job loop is independent per partition (one LocalScheduler each,
 * JobScheduler.cpp:5757-5766; the reference's "TODO: do it in parallel", :5756):
 * with part_owner[p] == rank this handle commits partition p and leaves every
 * other partition to the handle (GPU) that owns it. Priority, batch limit and
 * queue order are computed from the WHOLE pending table on every rank (upload
 * the same tables everywhere), so they are identical on all ranks. After
 * crane_sched_run the device-side placement columns hold zeros for the jobs this
 * rank does not own (a job without a valid partition belongs to rank 0), which
 * makes the union over ranks a plain sum: one all-reduce(sum) per column of
 * crane_sched_device_placements(), then crane_sched_fetch on any rank returns the
 * whole tick. n_ranks == 1 (the default) switches sharding off. Partitions
 * that share nodes are one unit and must have one owner (CRANE_EINVAL). */
int crane_sched_set_shard(crane_sched_t* h, uint32_t rank, uint32_t n_ranks,
                          const uint32_t* part_owner /* [n_partitions] */);

/* Device pointers of the placement columns of the last crane_sched_run, for the
 * all-reduce above (NCCL through the caller's communicator). `priority` is
 * identical on all ranks and is not part of the exchange. */
typedef struct crane_device_placements {
  void* reason;       /* uint8  [n_jobs]                                      */
  void* start_time;   /* int64  [n_jobs]                                      */
  void* end_time;     /* int64  [n_jobs]                                      */
  void* n_alloc;      /* uint32 [n_jobs]                                      */
  void* alloc_node;   /* uint32 [n_rows]                                      */
  void* alloc_ntasks; /* uint32 [n_rows]                                      */
  void* alloc_res;    /* crane_res_in_node_t [n_rows] = 9 x uint64 per row    */
  uint64_t n_jobs;
  uint64_t n_rows;
} crane_device_placements_t;
int crane_sched_device_placements(crane_sched_t* h, crane_device_placements_t* out);

/* ---- QoS post-filter (R12) ------------------------------------------------ */
/* A ResourceView limit of struct Qos (Account/AccountDefs.h:27-50): max_tres,
 * max_tres_per_user, max_tres_per_account. A gres name / type that is absent
 * from the limit map is "unlimited"; presence is carried by the two masks
 * (bit g = name g present, bit e = dictionary entry e present in `specified`). */
typedef struct crane_tres_limit {
  crane_res_view_t view;
  uint8_t gres_name_present;
  uint8_t gres_spec_present;
  uint8_t pad[6];
} crane_tres_limit_t;

/* MetaResource (Accounting/AccountMetaContainer.h:30-47): running usage of one
 * (user,qos), (account,qos) or qos. */
typedef struct crane_meta_resource {
  int64_t cpu_raw;
  uint64_t mem;
  uint64_t mem_sw;
  uint32_t gres_total[CRANE_GRES_NAMES];
  uint32_t gres_spec[CRANE_GRES_ENTRIES];
  uint32_t jobs_count;
  uint32_t pad;
  int64_t wall_time;
} crane_meta_resource_t;

typedef struct crane_qos_table {
  uint32_t n_qos, n_users, n_accounts;
  /* limits, indexed by qos id (struct Qos, AccountDefs.h:27-50) */
  const uint8_t* valid;                 /* 0 -> "InvalidQOS" (deleted / unknown) */
  const uint32_t* max_jobs_per_user;
  const uint32_t* max_jobs_per_account;
  const uint32_t* max_jobs;
  const int64_t* max_cpus_per_user_raw;
  const int64_t* max_wall;              /* seconds; 0 = unlimited               */
  const crane_tres_limit_t* max_tres_per_user;
  const crane_tres_limit_t* max_tres_per_account;
  const crane_tres_limit_t* max_tres;
  /* account chain of every pending job (PdJobInScheduler::account_chain),
   * CSR over account ids, in chain order */
  const uint32_t* chain_off;            /* [n_pending + 1]                      */
  const uint32_t* chain_acct;
  /* usage, read and updated in place (host memory) */
  crane_meta_resource_t* user_usage;    /* [n_users][n_qos]                     */
  crane_meta_resource_t* account_usage; /* [n_accounts][n_qos]                  */
  crane_meta_resource_t* qos_usage;     /* [n_qos]                              */
} crane_qos_table_t;

/* Replaces: the CheckAndMallocQosResource calls of the commit loop
 * (JobScheduler.cpp:1262 -> Accounting/AccountMetaContainer.cpp:164-191, 382-531, 546-587)
 * for the placements of the last crane_sched_run: in job-id (input) order, every
 * job that NodeSelect starts now is checked against its (user,qos), account
 * chain and qos limits and, if it passes, added to the usage; a job that fails
 * gets the QoS reason code (it keeps the timeline reservation it made, as in
 * the reference). `reason` ([n_pending], host) receives the updated reasons.
 * Needs pending.qos and pending.user at upload. CRANE_ENOSYS for an account
 * chain longer than 30, CRANE_EINVAL for an account repeated inside a chain or
 * ids outside the tables. Usage maps are dense here, so the reference's
 * "QosResourceLimit" (entry missing from a map) cannot occur. */
int crane_sched_qos_filter(crane_sched_t* h, const crane_qos_table_t* qos, uint8_t* reason);

/* Capability bitmap of the last run (the jobs x nodes feasibility bitmap):
 * bit (rank r, local node q) set iff the r-th job in priority order can run
 * on the q-th node of its partition by total resources and node lists
 * (JobScheduler.cpp:5238-5266). words_per_row = ceil(max_part_size/32).
 * Copies min(cap, rows*words_per_row) words to `dst` (host). For tests. */
int crane_sched_debug_bitmap(crane_sched_t* h, uint32_t* dst, size_t cap_words,
                             uint32_t* rows, uint32_t* words_per_row);

/* Per-partition phase cycle counters of the commit kernel, [n_partitions][16];
 * all zero unless the library was built with -DCRANE_PROFILE (a profiling
 * build, never the shipped one). For tools/. */
int crane_sched_debug_profile(crane_sched_t* h, unsigned long long* dst, size_t cap);

#ifdef __cplusplus
}
#endif
#endif /* CRANE_SCHED_H_ */
