"""Builds oracle/_ref/libcrane_ref.so: the REFERENCE's own scheduling code.

TEST INFRASTRUCTURE ONLY (same rules as the rest of oracle/).

The reference daemon cannot be built in this image (GCC >= 14, ~20 fetched
dependencies, protoc), but the scheduling hot path itself is a
few hundred lines of standard C++ on top of four third-party surfaces: absl
time + hash containers, fpm::fixed, a handful of crane::grpc enums and the
daemon's singletons. This recipe

  1. slices the hot path's own text out of /root/reference — by ANCHOR lines
     (it fails loudly if an anchor moved), never by hand —
        src/CraneCtld/JobScheduler.h      class IUpdateNodeCostPolicy .. end of class SchedulerAlgo
        src/CraneCtld/JobScheduler.cpp    LocalScheduler::CalculateRunningNodesAndStartTime_ .. end of
                                          SchedulerAlgo::NodeSelect, and MultiFactorPriority::*
        src/Utilities/PublicHeader/include/crane/PublicHeader.h   SlotId .. ResourceView operators
        src/Utilities/PublicHeader/PublicHeader.cpp               the non-protobuf member functions
        src/CraneCtld/Accounting/AccountMetaContainer.{h,cpp}     MetaResource, CheckAndMallocQosResource,
                                          CheckQosResource_, CheckTres_, CheckGres_, DoMallocResource_
     into oracle/_ref/gen/*.inc (git-ignored: reference text is never committed);
  2. compiles oracle/ref_shim/ref_harness.cpp, which includes those slices
     UNMODIFIED between small shim headers (oracle/ref_shim/*.h: absl::Time as
     saturating int64 seconds, flat_hash_map as an ordered map on a bump arena,
     the fpm::fixed subset, stub singletons), with g++ -std=c++23;
  3. exposes crane_ref_node_select() and crane_ref_qos_filter() with the oracle's C signatures.

What the shims decide (documented deviations): hash-container
iteration order becomes key order and NodeState addresses follow insertion
order (D1: equal-cost nodes by index), see ref_shim/absl_shim.h.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(OUT, "gen")
SO = os.path.join(OUT, "libcrane_ref.so")


def _lines(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read().split("\n")


def _find(lines, pattern, start=0):
    rx = re.compile(pattern)
    for i in range(start, len(lines)):
        if rx.search(lines[i]):
            return i
    raise SystemExit("ref_build: anchor %r not found — the reference changed, fix oracle/ref_build.py" % pattern)


def _slice(lines, first_pat, end_pat, include_end=False, start=0):
    a = _find(lines, first_pat, start)
    b = _find(lines, end_pat, a + 1)
    return lines[a:b + (1 if include_end else 0)], b


def _func_end(lines, a):
    """Index of the line holding the closing brace of the top-level definition starting at line a."""
    depth, seen = 0, False
    for i in range(a, len(lines)):
        for ch in lines[i]:
            if ch == "{":
                depth += 1
                seen = True
            elif ch == "}":
                depth -= 1
        if seen and depth == 0:
            return i
    raise SystemExit("ref_build: unbalanced braces after line %d" % a)


def extract():
    os.makedirs(GEN, exist_ok=True)
    # ---- JobScheduler.h: cost policy .. SchedulerAlgo -------------------------
    h = _lines("src/CraneCtld/JobScheduler.h")
    body, _ = _slice(h, r"^class IUpdateNodeCostPolicy \{", r"^class JobScheduler \{")
    _write("js_h.inc", "src/CraneCtld/JobScheduler.h", body)
    # ---- JobScheduler.cpp ------------------------------------------------------
    c = _lines("src/CraneCtld/JobScheduler.cpp")
    sel, _ = _slice(c, r"^bool SchedulerAlgo::LocalScheduler::CalculateRunningNodesAndStartTime_\(",
                    r"^void JobScheduler::ProcessFinalSteps_\(")
    a = _find(c, r"^void MultiFactorPriority::GetOrderedJobPtrVec\(")
    b = _find(c, r"^\}  // namespace Ctld", a)
    _write("js_cpp.inc", "src/CraneCtld/JobScheduler.cpp", sel + c[a:b])
    # ---- PublicHeader.h ----------------------------------------------------------
    ph = _lines("src/Utilities/PublicHeader/include/crane/PublicHeader.h")
    body, _ = _slice(ph, r"^using SlotId = std::string;", r"^template <class\.\.\. Ts>")
    _write("ph_h.inc", "src/Utilities/PublicHeader/include/crane/PublicHeader.h", body)
    # ---- PublicHeader.cpp: every top-level definition that does not touch protobuf types ----
    pc = _lines("src/Utilities/PublicHeader/PublicHeader.cpp")
    start = _find(pc, r"^GresCount& GresCount::operator\+=")
    out, i, skipped = [], start, []
    while i < len(pc):
        ln = pc[i]
        if ln and not ln.startswith((" ", "/", "}", "#")):
            # a definition starts here; it may span several header lines
            e = _func_end(pc, i)
            text = "\n".join(pc[i:e + 1])
            head = text.split("{", 1)[0]
            if "crane::grpc" in head or "Grpc" in head:
                skipped.append(pc[i])
            else:
                out += pc[i:e + 1] + [""]
            i = e + 1
        else:
            i += 1
    _write("ph_cpp.inc", "src/Utilities/PublicHeader/PublicHeader.cpp", out)
    # ---- AccountMetaContainer: the QoS check + malloc of the commit loop (JobScheduler.cpp:1262) ----
    ah = _lines("src/CraneCtld/Accounting/AccountMetaContainer.h")
    a = _find(ah, r"^struct MetaResource \{")
    _write("amc_h.inc", "src/CraneCtld/Accounting/AccountMetaContainer.h", ah[a:_func_end(ah, a) + 1])
    ac = _lines("src/CraneCtld/Accounting/AccountMetaContainer.cpp")
    body = []
    for pat in (r"^bool MetaResource::operator<=\(", r"^MetaResource& MetaResource::operator\+=\(",
                r"^MetaResource& MetaResource::operator-=\(", r"^void MetaResource::SetToZero\(",
                r"^AccountMetaContainer::CheckAndMallocQosResource\(", r"^std::expected<void, std::string> AccountMetaContainer::CheckQosResource_\(",
                r"^std::expected<void, std::string> AccountMetaContainer::CheckTres_\(", r"^bool AccountMetaContainer::CheckGres_\(",
                r"^AccountMetaContainer::LockAccountStripes_\(", r"^void AccountMetaContainer::DoMallocResource_\("):
        a = _find(ac, pat)
        first = a - 1 if not ac[a].startswith(("bool", "void", "MetaResource&", "std::expected")) else a  # return type on the line above
        body += ac[first:_func_end(ac, a) + 1] + [""]
    _write("amc_cpp.inc", "src/CraneCtld/Accounting/AccountMetaContainer.cpp", body)
    return skipped


def _write(name, origin, body):
    with open(os.path.join(GEN, name), "w") as f:
        f.write("// GENERATED by oracle/ref_build.py from /root/reference/%s — verbatim slice, not committed.\n" % origin)
        f.write("\n".join(body))
        f.write("\n")


def build(force: bool = False) -> str | None:
    """Returns the .so path, or None when /root/reference is absent and nothing
    prebuilt exists (GPU box: the prebuilt .so travels with the snapshot)."""
    have_ref = os.path.isdir(os.path.join(REF, "src", "CraneCtld"))
    shim = os.path.join(HERE, "ref_shim")
    srcs = [os.path.join(shim, f) for f in sorted(os.listdir(shim))] + [os.path.abspath(__file__)]
    if os.path.exists(SO) and not force:
        if not have_ref or all(os.path.getmtime(s) <= os.path.getmtime(SO) for s in srcs):
            return SO
    if not have_ref:
        return SO if os.path.exists(SO) else None
    extract()
    cmd = ["g++", "-std=c++23", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-fPIC", "-shared", "-w",
           "-I", shim, "-I", GEN, "-I", os.path.join(HERE, ".."),
           os.path.join(shim, "ref_harness.cpp"), "-o", SO]
    subprocess.check_call(cmd)
    return SO


if __name__ == "__main__":
    p = build(force="-f" in sys.argv)
    print(p or "no /root/reference and no prebuilt oracle/_ref/libcrane_ref.so")
