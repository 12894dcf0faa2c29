"""ctypes mirror of the tensor-core variant (ising_tc_* in include/ising_b200.h): drop-in for the
hot path of the reference's tensorcore/main.cu, driven the way its main() drives it
(:492-497 init, :582-595 warm-up + trial iterations numbered consecutively from 1)."""
import ctypes as C
import os

import numpy as np

from .engine import IsingError, load_library

NN_MMA, NN_STENCIL, NN_MMA_TILE = 0, 1, 2   # persistent tcgen05 pipeline (product), integer stencil, one CTA per tile


class TensorcoreEngine:
    def __init__(self, nbx, nby, seed=1234, alpha=0.1, device=0, nn_path=None, ngpus=1):
        self._L = load_library()
        self.nbx, self.nby = int(nbx), int(nby)
        self.nx, self.ny = 256 * self.nbx, 256 * self.nby
        h = C.c_void_p()
        self._h = None
        self._ck(self._L.ising_tc_create_multi(self.nbx, self.nby, int(seed), C.c_float(float(np.float32(alpha))),
                                               int(device), int(ngpus), C.byref(h)))
        self._h = h
        self.iter = 0
        if nn_path is not None:
            self.set_nn_path(nn_path)

    def _ck(self, rc):
        if rc != 0:
            raise IsingError(rc, self._L.ising_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.ising_tc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_nn_path(self, path):
        self._ck(self._L.ising_tc_set_nn_path(self._h, int(path)))

    def init(self):
        self._ck(self._L.ising_tc_init(self._h))
        self.iter = 0

    def update(self, n=1, sync=True):
        self._ck(self._L.ising_tc_update(self._h, self.iter + 1, int(n)))
        self.iter += int(n)
        if sync:
            self.sync()

    def update_timed(self, n):
        ms = C.c_float()
        self._ck(self._L.ising_tc_update_timed(self._h, self.iter + 1, int(n), C.byref(ms)))
        self.iter += int(n)
        return float(ms.value)

    def sync(self):
        self._ck(self._L.ising_tc_sync(self._h))

    def sum(self):
        s = C.c_int64()
        self._ck(self._L.ising_tc_sum(self._h, C.byref(s)))
        return int(s.value)

    def storage(self):
        out = np.empty(self.nx * self.ny, dtype=np.int8)
        self._ck(self._L.ising_tc_export_storage(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def import_storage(self, arr):
        a = np.ascontiguousarray(arr, dtype=np.int8)
        assert a.size == self.nx * self.ny
        self._ck(self._L.ising_tc_import_storage(self._h, C.c_void_p(a.ctypes.data)))

    def true_lattice(self):
        out = np.empty((self.ny, self.nx), dtype=np.int8)
        self._ck(self._L.ising_tc_export_true(self._h, C.c_void_p(out.ctypes.data)))
        return out

    def write_lattice(self, fname):
        self._ck(self._L.ising_tc_write_lattice(self._h, os.fsencode(fname)))

    def accept_table(self):
        t = (C.c_float * 5)()
        self._ck(self._L.ising_tc_accept_table(self._h, t))
        return np.array(list(t), dtype=np.float32)

    def neighbour_sums(self, is_black):
        out = np.empty(self.nx * self.ny, dtype=np.int8)
        self._ck(self._L.ising_tc_neighbour_sums(self._h, int(bool(is_black)), C.c_void_p(out.ctypes.data)))
        return out

    @property
    def launches(self):
        return int(self._L.ising_tc_launch_count(self._h))
