"""The REFERENCE's own NodeSelect (oracle/_ref) on the whole queue of a full-size golden case, compared with
the digests of tests/golden/full_digests.json (the ones the CUDA path is held to). Unique priorities as in
check_full_golden_ref.py (mandated_priority = the oracle's own order, taken from a run that schedules one
job), so every column but `priority` must give the golden digest. One reference run: ~1.5 CPU-hours for
config2_full. Result: tests/golden/ref_check.json.

    python -m tests.golden.check_golden_ref_full config2_full"""
import dataclasses
import json
import os
import sys
import time

import numpy as np

from tests.golden.make_full_golden import CASES, COLUMNS, HERE, digest_of, load

OUT = os.path.join(HERE, "ref_check.json")


def main(names):
    from oracle import pyoracle, pyref
    pyoracle.build()
    want = load()
    for name in names:
        cfg, cl, rn, pd, now = CASES[name][0]()
        first, _, _ = pyoracle.node_select(dataclasses.replace(cfg, scheduled_batch_size=1), cl, rn, pd, now)
        order = np.argsort(-first.priority, kind="stable")
        mand = np.empty(pd.n, np.float64)
        mand[order] = pd.n - np.arange(pd.n)
        t0 = time.time()
        out, ms = pyref.node_select(cfg, cl, rn, dataclasses.replace(pd, mandated_priority=mand), now)
        got = digest_of(out)
        cols = [c for c in COLUMNS if c != "priority"]
        diff = [c for c in cols if got[c] != want[name][c]]
        res = {"reference_equals_golden_digests": not diff, "columns_compared": cols, "differing_columns": diff,
               "reference_ms": round(ms, 1), "wall_s": round(time.time() - t0, 1), "n_jobs": int(pd.n), "n_nodes": int(cl.n_nodes),
               "n_started": got["n_started"], "n_reserved": got["n_reserved"]}
        allr = json.load(open(OUT)) if os.path.exists(OUT) else {}
        allr["golden:" + name] = res
        with open(OUT, "w") as f:
            json.dump(allr, f, indent=1, sort_keys=True)
        print(name, res, flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
