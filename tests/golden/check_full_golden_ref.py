"""Oracle vs the REFERENCE's own NodeSelect (oracle/_ref) on the BASELINE configs at sizes far above the
unit-test pins (tests/test_ref_pin.py). The reference sorts the queue with an UNSTABLE sort
(JobScheduler.cpp:6541), so equal computed priorities — which appear from a few thousand jobs on — make
its order of those jobs arbitrary; to compare whole queues the case is first given unique priorities:
every job gets `mandated_priority` = N - (its rank in the oracle's own order), which keeps exactly the
order the oracle (and the CUDA path) use. Both are then run on that case and every output column is
compared. Results: tests/golden/ref_check.json. CPU only, single thread, minutes per case.

    python -m tests.golden.check_full_golden_ref config2:20000:2000:3002 [...]   (config:jobs:nodes:seed)"""
import dataclasses
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cranesched_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "ref_check.json")


def main(specs):
    from oracle import pyoracle, pyref
    pyoracle.build()
    if not pyref.available():
        raise SystemExit("oracle/_ref/libcrane_ref.so is not built")
    for spec in specs:
        cid, nj, nn, seed = spec.split(":")
        gen = synth.CONFIGS[int(cid[-1])]
        cfg, cl, rn, pd, now = gen(n_jobs=int(nj), n_nodes=int(nn), seed_id=int(seed))
        # the order both use: priorities of the plain case (a run that schedules one job is enough for them)
        first, _, _ = pyoracle.node_select(dataclasses.replace(cfg, scheduled_batch_size=1), cl, rn, pd, now)
        order = np.argsort(-first.priority, kind="stable")
        mand = np.empty(pd.n, np.float64)
        mand[order] = pd.n - np.arange(pd.n)
        pd2 = dataclasses.replace(pd, mandated_priority=mand)
        t0 = time.time()
        a, oms, _ = pyoracle.node_select(cfg, cl, rn, pd2, now)
        b, rms = pyref.node_select(cfg, cl, rn, pd2, now)
        d = a.diff(b)
        res = {"columns_equal": not d, "first_differences": d[:3], "oracle_ms": round(oms, 1), "reference_ms": round(rms, 1),
               "n_started": int((a.reason == 0).sum()), "n_reserved": int(((a.reason != 0) & (a.n_alloc > 0)).sum()),
               "wall_s": round(time.time() - t0, 1)}
        allr = json.load(open(OUT)) if os.path.exists(OUT) else {}
        allr[spec] = res
        with open(OUT, "w") as f:
            json.dump(allr, f, indent=1, sort_keys=True)
        print(spec, res, flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
