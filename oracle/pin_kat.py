#!/usr/bin/env python
"""Pin the CPU oracle against the known-answer spin counts in the reference's
optimized/README.md (default seed 463463564571, T=1.5).  Slow (minutes, ~9 GB RAM):
run once per oracle change; the fast subset lives in tests/test_oracle_kat.py.

    python oracle/pin_kat.py [--sweeps 16] [--skip-big]

Writes oracle/pin_kat.log (committed) with PASS/FAIL per vector.
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ising_oracle as o  # noqa: E402

# optimized/README.md:128-137 (KAT-1), :187-196 (KAT-4), :240 (KAT-2 it 0), :307 (KAT-3 it 0)
KAT1 = {0: (2147484090, 2147483206), 16: (2147575418, 2147391878), 32: (2147641872, 2147325424),
        48: (2147605659, 2147361637), 64: (2147701147, 2147266149)}
KAT4 = {0: (2147484090, 2147483206), 16: (2147594634, 2147372662)}
KAT2_INIT = (131072, 65536, (4294989182, 4294945410))
KAT3_INIT = (524288, 65536, (17179689306, 17180049062))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sweeps", type=int, default=16)
    ap.add_argument("--skip-big", action="store_true")
    a = ap.parse_args()
    out = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pin_kat.log"), "w")

    def say(msg):
        print(msg, flush=True)
        out.write(msg + "\n")
        out.flush()

    ok = True
    say("oracle threads: %d" % o.num_threads())
    for name, kat, xsl, ysl in (("KAT-1 65536x65536 T=1.5", KAT1, 0, 0),
                                ("KAT-4 65536x65536 T=1.5 --xsl 2048 --ysl 2048", KAT4, 2048, 2048)):
        t0 = time.time()
        L = o.OracleLattice(65536, 65536, xsl=xsl, ysl=ysl)
        L.init()
        for it in range(0, a.sweeps + 1, 16):
            if it:
                L.sweeps(16, 1.5)
            got = L.count()
            if it in kat:
                good = got == kat[it]
                ok &= good
                say("%s it %3d: up/down %d/%d  expect %d/%d  %s  (%.0f s)" %
                    (name, it, got[0], got[1], kat[it][0], kat[it][1], "PASS" if good else "FAIL",
                     time.time() - t0))
        L.close()
    if not a.skip_big:
        for rows, X, exp in (KAT2_INIT, KAT3_INIT):
            t0 = time.time()
            got = o.init_count(X, rows)
            good = got == exp
            ok &= good
            say("init %dx%d: up/down %d/%d expect %d/%d %s (%.0f s)" %
                (rows, X, got[0], got[1], exp[0], exp[1], "PASS" if good else "FAIL", time.time() - t0))
    say("ALL PASS" if ok else "SOME FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
