"""Helpers around tests/golden/reference_b200.json (outputs of the unmodified reference binary on
a B200, see tools/gen_golden_ref.sh / tools/pack_golden.py)."""
import json
import os
import re
import tarfile

HERE = os.path.dirname(os.path.abspath(__file__))
_INDEX = None


def cases():
    global _INDEX
    if _INDEX is None:
        _INDEX = json.load(open(os.path.join(HERE, "golden", "reference_b200.json")))["cases"]
    return _INDEX


def extract_files(case, dest):
    """Extract the golden dump/corr files of one case into dest; returns {basename: path}."""
    out = {}
    with tarfile.open(os.path.join(HERE, "golden", "reference_b200_files.tgz")) as t:
        for m in t.getmembers():
            if m.name.startswith(case + "/"):
                m.name = os.path.basename(m.name)
                t.extract(m, dest, filter="data")
                out[m.name] = os.path.join(dest, m.name)
    return out


_MAG = re.compile(r"(Initial|Final)?\s*magnetization:\s+([0-9.]+)(?: \(\^2:\s+[0-9.]+\))?, up_s:\s+(\d+), dw_s:\s+(\d+)"
                  r"(?: \(iter:\s+(\d+)\))?")


def parse_counts(lines):
    """[(kind, iter or None, up, down)] for every magnetization line, in order."""
    out = []
    for ln in lines:
        m = _MAG.search(ln)
        if m:
            kind = m.group(1) or "step"
            out.append((kind, int(m.group(5)) if m.group(5) else None, int(m.group(3)), int(m.group(4))))
    return out


def parse_args(args):
    """The subset of cuIsing options used by the golden cases -> dict (main.cu:1285-1393)."""
    o = {"x": 0, "y": 0, "n": 1, "p": 0, "t": None, "a": None, "s": 463463564571, "xsl": 0, "ysl": 0, "o": False,
         "u": None, "J": None, "c": False, "e": False, "m": None, "d": 1}
    i = 0
    while i < len(args):
        a = args[i]
        if a in ("-o", "-c", "-e"):
            o[a[1]] = True
            i += 1
            continue
        v = args[i + 1]
        if a == "-u":
            s, f = v.split(",")
            o["u"] = (float(s), int(f))
        elif a in ("-t", "-a", "-J", "-m"):
            o[a[1]] = float(v)
        elif a in ("--xsl", "--ysl"):
            o[a[2:]] = int(v)
        else:
            o[a.lstrip("-")] = int(v)
        i += 2
    return o
