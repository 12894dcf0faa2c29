#!/usr/bin/env python3
"""Python-3 replacement of the reference's optimized/plotLattice.py (Python 2 + matplotlib, :1-18):
reads a lattice dump written with -o (one text line of '0'/'1' per row, optimized/main.cu:1160-1175)
and writes <file>.png.  Several files (the per-device dumps of a -d N run) are stacked top to bottom.

    plot_lattice.py lattice_..._0.txt [lattice_..._1.txt ...] [--scale K] [-o OUT.png]

Uses matplotlib when it is installed (same picture as the reference script), Pillow otherwise, and a
built-in PNG writer when neither is there.  --scale K averages K x K spins per pixel (large lattices)."""
import argparse
import struct
import sys
import zlib

import numpy as np


def read_dump(path):
    rows = []
    with open(path, "rb") as f:
        for line in f:
            line = line.strip(b" \n\r")
            if line:
                rows.append(np.frombuffer(line, dtype=np.uint8) - ord("0"))
    if not rows or any(len(r) != len(rows[0]) for r in rows):
        raise ValueError("%s: not a lattice dump (ragged or empty)" % path)
    a = np.vstack(rows)
    if a.max() > 1:
        raise ValueError("%s: characters other than 0/1" % path)
    return a


def downscale(a, k):
    if k <= 1:
        return a.astype(np.float32)
    y, x = (a.shape[0] // k) * k, (a.shape[1] // k) * k
    return a[:y, :x].reshape(y // k, k, x // k, k).mean(axis=(1, 3), dtype=np.float32)


def write_png_gray(path, img8):
    """Minimal 8-bit grayscale PNG encoder (zlib + CRC from the standard library)."""
    h, w = img8.shape
    raw = b"".join(b"\x00" + img8[r].tobytes() for r in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 0, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("files", nargs="+")
    ap.add_argument("--scale", type=int, default=1)
    ap.add_argument("-o", "--out", default=None)
    ap.add_argument("--backend", choices=["auto", "matplotlib", "pillow", "builtin"], default="auto")
    a = ap.parse_args(argv)
    data = np.vstack([read_dump(f) for f in a.files])
    print(data.shape[0], "x", data.shape[1])          # the reference script prints the same line
    img = downscale(data, a.scale)
    out = a.out or (a.files[0] + ".png")
    backend = a.backend
    if backend in ("auto", "matplotlib"):
        try:
            import matplotlib
            matplotlib.use("Agg")
            from matplotlib import pyplot as plt
            plt.imshow(img, interpolation="nearest")
            plt.savefig(out)
            return out
        except ImportError:
            if backend == "matplotlib":
                raise
    img8 = np.rint(img * 255.0).astype(np.uint8)
    if backend in ("auto", "pillow"):
        try:
            from PIL import Image
            Image.fromarray(img8, mode="L").save(out)
            return out
        except ImportError:
            if backend == "pillow":
                raise
    write_png_gray(out, img8)
    return out


if __name__ == "__main__":
    main()
