#!/usr/bin/env python
"""SASS opcode histogram of the hot kernels in libising_b200.so (cuobjdump -sass), written to
profiles/sass_opcodes_<kernel>.txt by __graft_entry__.build().  This is the committed evidence that the shipped
kernels are Blackwell-native: UTMALDG = cp.async.bulk.tensor (TMA load), SYNCS = mbarrier, UTCIMMA = tcgen05.mma
kind::i8, LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ising_gpu_b200", "libising_b200.so")
# kernel tag -> regex on the mangled name
KERNELS = {
    "update_color_kernel2_fast": r"update_color_kernel2ILb1E",
    "update_color_kernel_fast": r"update_color_kernelILb1ELb0ELb0E",
    "update_color_kernel_fast_push": r"update_color_kernelILb1ELb0ELb1E",
    "update_color_persistent_fast": r"update_color_persistentILb1ELb0ELb0E",
    "tc_update_pipe_kernel": r"tc_update_pipe_kernel",
    "init_kernel": r"init_kernel",
    "count_kernel": r"count_kernel",
}
MUST = {"update_color_kernel2_fast": ["UTMALDG", "SYNCS", "IMAD.WIDE.U32"],
        "update_color_kernel_fast": ["UTMALDG", "SYNCS", "IMAD.WIDE.U32"],
        "update_color_persistent_fast": ["UTMALDG", "SYNCS", "IMAD.WIDE.U32"],
        "tc_update_pipe_kernel": ["UTCIMMA", "LDTM", "UTMALDG", "SYNCS"]}


def histograms(lib=LIB):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    out, cur = {}, None
    for ln in txt.split("\n"):
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = None
            for tag, pat in KERNELS.items():
                if re.search(pat, m.group(1)) and tag not in out:
                    cur = tag
                    out[tag] = {"name": m.group(1), "ops": collections.Counter()}
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)", ln)
        if m:
            op = m.group(1)
            base = op if op.startswith("IMAD.WIDE") else op.split(".")[0]
            out[cur]["ops"][base] += 1
    return out


def write(outdir=os.path.join(ROOT, "profiles")):
    hs = histograms()
    for tag, h in hs.items():
        with open(os.path.join(outdir, "sass_opcodes_%s.txt" % tag), "w") as f:
            f.write("# cuobjdump -sass ising_gpu_b200/libising_b200.so, static opcode counts (tools/sass_histogram.py)\n")
            f.write("# function: %s\n# total instructions: %d\n" % (h["name"], sum(h["ops"].values())))
            for op, n in sorted(h["ops"].items(), key=lambda kv: (-kv[1], kv[0])):
                f.write("%6d  %s\n" % (n, op))
    return hs


if __name__ == "__main__":
    hs = write()
    bad = 0
    for tag, need in MUST.items():
        for op in need:
            if not hs.get(tag) or not any(k.startswith(op) for k in hs[tag]["ops"]):
                print("MISSING %s in %s" % (op, tag)); bad = 1
    for tag, h in hs.items():
        print("%-34s %5d instr  %s" % (tag, sum(h["ops"].values()),
                                      " ".join("%s:%d" % kv for kv in h["ops"].most_common(6))))
    sys.exit(bad)
