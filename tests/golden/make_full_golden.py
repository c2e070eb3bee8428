"""Full-queue oracle results of the BASELINE.json configs, as SHA-256 digests.

    python -m tests.golden.make_full_golden <name> [...]      (from the repo root)

For every case in CASES the CPU oracle (oracle/crane_oracle.cpp) schedules the
WHOLE queue and one digest per output column is stored in
tests/golden/full_digests.json; results small enough to commit are also kept
as <name>.npz. The `-m gpu` tests in tests/test_gpu_full_golden.py run the
same synthetic case through the C-ABI on the B200 and compare every column.
Hours of CPU for the largest cases (single thread per case, as the reference's
NodeSelect is single-threaded) — which is why they are precomputed here and
not inside the test.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cranesched_b200 import abi, synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
DIGESTS = os.path.join(HERE, "full_digests.json")
COLUMNS = ("reason", "priority", "start_time", "end_time", "n_alloc", "alloc_off",
           "alloc_node", "alloc_ntasks", "alloc_res")

# name -> (builder, keep_npz)
CASES = {
    "config2_full": (lambda: synth.config2(), False),                       # 100k x 10k
    "config2_seed3002": (lambda: synth.config2(n_jobs=20_000, n_nodes=2_000, seed_id=3002), True),
    "config3_10pct": (lambda: synth.config3(n_jobs=100_000, n_nodes=5_000), False),
    "config4_10pct": (lambda: synth.config4(n_jobs=50_000, n_nodes=2_000), False),
    "config5_full": (lambda: synth.config5(), False),                       # 200k x 5k
    "config5_25pct": (lambda: synth.config5(n_jobs=50_000, n_nodes=1_250), False),
}
for _s in range(8):  # config 2 at medium size, eight draws (VERDICT r1 #1c)
    CASES["config2_med_s%d" % _s] = ((lambda s=_s: synth.config2(n_jobs=12_000, n_nodes=1_200, seed_id=3000 + s)), False)


def digest_of(out: abi.Placements) -> dict:
    d = {}
    for f in COLUMNS:
        a = np.ascontiguousarray(getattr(out, f))
        d[f] = hashlib.sha256(a.tobytes()).hexdigest()
    d["n_started"] = int((out.reason == abi.REASON_NONE).sum())
    d["n_reserved"] = int(((out.reason != abi.REASON_NONE) & (out.n_alloc > 0)).sum())
    d["reason_hist"] = np.bincount(out.reason, minlength=5).tolist()
    return d


def load() -> dict:
    if os.path.exists(DIGESTS):
        with open(DIGESTS) as f:
            return json.load(f)
    return {}


def main(names):
    from oracle import pyoracle
    pyoracle.build()
    for name in names:
        mk, keep = CASES[name]
        case = mk()
        t0 = time.time()
        out, ms, done = pyoracle.node_select(*case[:4], case[4])
        d = digest_of(out)
        d["oracle_ms"] = round(ms, 1)
        d["n_jobs"] = int(case[3].n)
        d["n_nodes"] = int(case[1].n_nodes)
        if keep:
            np.savez_compressed(os.path.join(HERE, name + ".npz"),
                                **{f: getattr(out, f) for f in abi.Placements.__dataclass_fields__})
        # several generator processes may run at once: re-read before writing
        all_d = load()
        all_d[name] = d
        tmp = DIGESTS + ".%d.tmp" % os.getpid()
        with open(tmp, "w") as f:
            json.dump(all_d, f, indent=1, sort_keys=True)
        os.replace(tmp, DIGESTS)
        print(name, "oracle %.1f s (wall %.1f s)" % (ms / 1e3, time.time() - t0), d["reason_hist"], flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
