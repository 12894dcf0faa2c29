"""ctypes binding of oracle/libising_oracle.so (TEST INFRASTRUCTURE -- see ising_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(quiet=True):
    """Compile libising_oracle.so (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", _HERE] + (["-s"] if quiet else []), check=True)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_HERE, "libising_oracle.so")
    if not os.path.exists(so):
        build()
    L = C.CDLL(so)
    u32p = C.POINTER(C.c_uint32)
    u64p = C.POINTER(C.c_ulonglong)
    L.ora_philox4x32_10.argtypes = [u32p, u32p, u32p]
    L.ora_uniform.argtypes = [C.c_uint32]
    L.ora_uniform.restype = C.c_float
    L.ora_exp_table.argtypes = [C.c_float, C.POINTER(C.c_float)]
    L.ora_create.argtypes = [C.c_int, C.c_longlong, C.c_int, C.c_longlong]
    L.ora_create.restype = C.c_void_p
    L.ora_destroy.argtypes = [C.c_void_p]
    L.ora_plane.argtypes = [C.c_void_p, C.c_int]
    L.ora_plane.restype = C.POINTER(C.c_uint8)
    L.ora_init_color.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong]
    L.ora_update_color.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong, C.c_int, C.POINTER(C.c_float)]
    L.ora_sweep.argtypes = [C.c_void_p, C.c_ulonglong, C.c_int, C.c_float]
    L.ora_count.argtypes = [C.c_void_p, u64p, u64p]
    L.ora_energy.argtypes = [C.c_void_p]
    L.ora_energy.restype = C.c_double
    L.ora_export_ref.argtypes = [C.c_void_p, C.c_int, u64p]
    L.ora_import_ref.argtypes = [C.c_void_p, C.c_int, u64p]
    L.ora_export_bits.argtypes = [C.c_void_p, C.c_int, u32p]
    L.ora_import_bits.argtypes = [C.c_void_p, C.c_int, u32p]
    L.ora_dump.argtypes = [C.c_void_p, C.c_longlong, C.c_longlong, C.c_char_p]
    L.ora_num_threads.restype = C.c_int
    L.ora_init_bonds.argtypes = [C.c_void_p, C.c_ulonglong, C.c_float]
    L.ora_bonds.argtypes = [C.c_void_p, C.c_int]
    L.ora_bonds.restype = C.POINTER(C.c_uint8)
    L.ora_correlation.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.ora_make_slab.argtypes = [C.c_void_p, C.c_longlong]
    L.ora_ghost.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ora_ghost.restype = C.POINTER(C.c_uint8)
    L.ora_init_count.argtypes = [C.c_int, C.c_longlong, C.c_ulonglong]
    L.ora_init_count.restype = C.c_ulonglong
    i8p = C.POINTER(C.c_int8)
    L.tc_create.argtypes = [C.c_int, C.c_int, C.c_ulonglong, C.c_float]
    L.tc_create.restype = C.c_void_p
    L.tc_destroy.argtypes = [C.c_void_p]
    L.tc_set_accept.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.tc_get_accept.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
    L.tc_storage.argtypes = [C.c_void_p]
    L.tc_storage.restype = i8p
    L.tc_init.argtypes = [C.c_void_p]
    L.tc_update_color.argtypes = [C.c_void_p, C.c_int, C.c_ulonglong]
    L.tc_update.argtypes = [C.c_void_p, C.c_ulonglong]
    L.tc_sum.argtypes = [C.c_void_p]
    L.tc_sum.restype = C.c_longlong
    L.tc_export_true.argtypes = [C.c_void_p, i8p]
    L.tc_neighbour_sums.argtypes = [C.c_void_p, C.c_int, i8p]
    L.tc_write_lattice.argtypes = [C.c_void_p, C.c_char_p]
    _LIB = L
    return L


DEFAULT_SEED = 463463564571          # optimized/main.cu:63
CRIT_TEMP = np.float32(2.26918531421)  # optimized/main.cu:42


def philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    lib().ora_philox4x32_10(c, k, o)
    return tuple(int(v) for v in o)


def uniform(x):
    return np.float32(lib().ora_uniform(int(x)))


def exp_table(temp):
    t = (C.c_float * 10)()
    lib().ora_exp_table(C.c_float(float(np.float32(temp))), t)
    return np.array(list(t), dtype=np.float32).reshape(2, 5)


class OracleLattice:
    """Byte-per-spin CPU lattice driven exactly like optimized/main.cu drives the GPU one."""

    def __init__(self, X, rows, seed=DEFAULT_SEED, xsl=0, ysl=0):
        self.X, self.rows, self.seed = int(X), int(rows), int(seed)
        self._h = lib().ora_create(self.X, self.rows, int(xsl), int(ysl))
        if not self._h:
            raise ValueError("bad lattice dims: X multiple of 2048, rows multiple of 16")
        self.it = 0

    def close(self):
        if self._h:
            lib().ora_destroy(self._h)
            self._h = None

    __del__ = close

    def init(self):
        lib().ora_init_color(self._h, 0, self.seed)
        lib().ora_init_color(self._h, 1, self.seed)
        self.it = 0

    def sweeps(self, n, temp):
        for _ in range(int(n)):
            lib().ora_sweep(self._h, self.seed, self.it, C.c_float(float(np.float32(temp))))
            self.it += 1

    def update_color(self, color, it, temp):
        tab = exp_table(temp)
        buf = (C.c_float * 10)(*[float(v) for v in tab.reshape(-1)])
        lib().ora_update_color(self._h, int(color), self.seed, int(it), buf)

    def count(self):
        up, dn = C.c_ulonglong(), C.c_ulonglong()
        lib().ora_count(self._h, C.byref(up), C.byref(dn))
        return int(up.value), int(dn.value)

    def energy(self):
        return float(lib().ora_energy(self._h))

    def plane(self, color):
        n = self.rows * (self.X // 2)
        p = lib().ora_plane(self._h, int(color))
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(self.rows, self.X // 2)

    def export_ref(self, color):
        out = np.empty((self.rows, self.X // 32), dtype=np.uint64)
        lib().ora_export_ref(self._h, int(color), out.ctypes.data_as(C.POINTER(C.c_ulonglong)))
        return out

    def import_ref(self, color, words):
        w = np.ascontiguousarray(words, dtype=np.uint64)
        lib().ora_import_ref(self._h, int(color), w.ctypes.data_as(C.POINTER(C.c_ulonglong)))

    def export_bits(self, color):
        out = np.empty((self.rows, self.X // 64), dtype=np.uint32)
        lib().ora_export_bits(self._h, int(color), out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out

    def import_bits(self, color, words):
        w = np.ascontiguousarray(words, dtype=np.uint32)
        lib().ora_import_bits(self._h, int(color), w.ctypes.data_as(C.POINTER(C.c_uint32)))

    def dump(self, fname, row0=0, nrows=None):
        nrows = self.rows - row0 if nrows is None else nrows
        if lib().ora_dump(self._h, int(row0), int(nrows), os.fsencode(fname)) != 0:
            raise OSError("cannot write " + fname)

    def init_bonds(self, prob):
        """-J prob: random +-J bonds (seed+1 inside, as main.cu:1734)."""
        if lib().ora_init_bonds(self._h, self.seed, C.c_float(float(np.float32(prob)))) != 0:
            raise MemoryError

    def bonds(self, color):
        n = self.rows * (self.X // 2)
        return np.ctypeslib.as_array(lib().ora_bonds(self._h, int(color)), shape=(n,)).reshape(self.rows, self.X // 2)

    def correlation(self, ncorr=128):
        out = (C.c_double * ncorr)()
        lib().ora_correlation(self._h, int(ncorr), out)
        return np.array(list(out))

    def make_slab(self, grow0):
        """Hold rows [grow0, grow0+rows) of a taller lattice; edge neighbours come from ghost rows."""
        if lib().ora_make_slab(self._h, int(grow0)) != 0:
            raise MemoryError
        self.is_slab = True

    def ghost(self, color, bottom):
        p = lib().ora_ghost(self._h, int(color), int(bool(bottom)))
        return np.ctypeslib.as_array(p, shape=(self.X // 2,))

    def lattice(self):
        """Full rows x X array of 0/1 in true lattice order (dumpLattice interleave)."""
        b, w = self.plane(0), self.plane(1)
        out = np.empty((self.rows, self.X), dtype=np.uint8)
        even = (np.arange(self.rows) % 2 == 0)
        out[even, 0::2] = b[even]; out[even, 1::2] = w[even]
        out[~even, 0::2] = w[~even]; out[~even, 1::2] = b[~even]
        return out


def num_threads():
    return int(lib().ora_num_threads())


def init_count(X, rows, seed=DEFAULT_SEED):
    """(up, down) right after the reference's lattice init, without storing the lattice."""
    up = int(lib().ora_init_count(int(X), int(rows), int(seed)))
    return up, int(X) * int(rows) - up


class TensorcoreOracle:
    """CPU restatement of tensorcore/main.cu (blocked +-1 lattice, its own RNG map)."""

    def __init__(self, nbx, nby, seed=1234, alpha=0.1):
        self.nbx, self.nby = int(nbx), int(nby)
        self.nx, self.ny = 256 * self.nbx, 256 * self.nby
        self._h = lib().tc_create(self.nbx, self.nby, int(seed), C.c_float(float(np.float32(alpha))))
        self.iter = 0          # iterations done so far (warm-up included), next one uses iter+1

    def close(self):
        if self._h:
            lib().tc_destroy(self._h)
            self._h = None

    __del__ = close

    def set_accept(self, table):
        t = (C.c_float * 5)(*[float(v) for v in table])
        lib().tc_set_accept(self._h, t)

    def accept(self):
        t = (C.c_float * 5)()
        lib().tc_get_accept(self._h, t)
        return np.array(list(t), dtype=np.float32)

    def init(self):
        lib().tc_init(self._h)
        self.iter = 0

    def update(self, n=1):
        for _ in range(int(n)):
            self.iter += 1
            lib().tc_update(self._h, self.iter)

    def update_color(self, is_black, it):
        lib().tc_update_color(self._h, int(bool(is_black)), int(it))

    def storage(self):
        p = lib().tc_storage(self._h)
        return np.ctypeslib.as_array(p, shape=(self.nx * self.ny,))

    def sum(self):
        return int(lib().tc_sum(self._h))

    def true_lattice(self):
        out = np.empty((self.ny, self.nx), dtype=np.int8)
        lib().tc_export_true(self._h, out.ctypes.data_as(C.POINTER(C.c_int8)))
        return out

    def neighbour_sums(self, is_black):
        out = np.empty(self.nx * self.ny, dtype=np.int8)
        lib().tc_neighbour_sums(self._h, int(bool(is_black)), out.ctypes.data_as(C.POINTER(C.c_int8)))
        return out

    def write_lattice(self, fname):
        if lib().tc_write_lattice(self._h, os.fsencode(fname)) != 0:
            raise OSError(fname)
