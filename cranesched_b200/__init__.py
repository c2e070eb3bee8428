"""cranesched_b200 — B200-native CraneCtld scheduling hot path.

A drop-in for ONE call of the reference daemon, SchedulerAlgo::NodeSelect
(src/CraneCtld/JobScheduler.cpp:1141): pending-queue -> node match, priority
sort and backfill, as hand-written sm_100a CUDA kernels behind a C-ABI
(include/crane_sched.h). See DESIGN.md.
"""
from .abi import Cluster, Config, Pending, Placements, Running  # noqa: F401
from .scheduler import CraneSchedError, GpuScheduler  # noqa: F401
