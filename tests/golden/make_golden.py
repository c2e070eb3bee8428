"""Regenerates tests/golden/*.npz from the oracle (run from the repo root:
`python -m tests.golden.make_golden`). The reference cannot be built or
imported in this image, so these fixtures pin OUR restatement (parity is
"unpinned" w.r.t. upstream, see oracle/crane_oracle.cpp)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from cranesched_b200 import abi, synth  # noqa: E402

CASES = {
    "random_7": lambda: synth.random_case(7, n_jobs=200, n_nodes=32, n_running=25),
    "random_8": lambda: synth.random_case(8, n_jobs=200, n_nodes=40, n_parts=4, n_running=30, fifo=True),
    "config2_small": lambda: synth.config2(n_jobs=1500, n_nodes=200),
}

if __name__ == "__main__":
    from oracle import pyoracle
    here = os.path.dirname(os.path.abspath(__file__))
    for name, mk in CASES.items():
        case = mk()
        out, ms, _ = pyoracle.node_select(*case[:4], case[4])
        np.savez_compressed(os.path.join(here, name + ".npz"),
                            **{f: getattr(out, f) for f in abi.Placements.__dataclass_fields__})
        print(name, "%.1f ms" % ms, np.bincount(out.reason, minlength=5).tolist())
