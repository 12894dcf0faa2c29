"""tools/plot_lattice.py (Python-3 replacement of optimized/plotLattice.py) on an oracle dump."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import plot_lattice  # noqa: E402


def _dump(oracle, tmp_path, name, rows, row0=0):
    L = oracle.OracleLattice(2048, 32)
    L.init(); L.sweeps(2, 1.5)
    p = str(tmp_path / name)
    L.dump(p, row0, rows)
    lat = L.lattice()
    L.close()
    return p, lat


def test_reads_dump_and_writes_png(oracle, tmp_path, capsys):
    p, lat = _dump(oracle, tmp_path, "lattice_0.txt", 32)
    a = plot_lattice.read_dump(p)
    assert a.shape == (32, 2048) and np.array_equal(a, lat)
    for backend in ("builtin", "pillow"):
        out = plot_lattice.main([p, "--backend", backend, "-o", str(tmp_path / (backend + ".png"))])
        assert "32 x 2048" in capsys.readouterr().out
        from PIL import Image
        img = np.asarray(Image.open(out))
        assert img.shape == (32, 2048) and np.array_equal(img > 127, lat.astype(bool))


def test_stacks_device_files_and_downscales(oracle, tmp_path):
    p0, lat = _dump(oracle, tmp_path, "lattice_0.txt", 16, 0)
    p1, _ = _dump(oracle, tmp_path, "lattice_1.txt", 16, 16)
    out = plot_lattice.main([p0, p1, "--scale", "4", "--backend", "builtin", "-o", str(tmp_path / "s.png")])
    from PIL import Image
    img = np.asarray(Image.open(out))
    assert img.shape == (8, 512)
    want = np.rint(lat.reshape(8, 4, 512, 4).mean(axis=(1, 3), dtype=np.float32) * 255.0).astype(np.uint8)
    assert np.array_equal(img, want)
