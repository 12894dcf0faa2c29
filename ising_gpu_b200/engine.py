"""ctypes mirror of include/ising_b200.h.

``IsingEngine`` follows the step order of the reference's main() (optimized/main.cu:1596-1925):
create -> init_lattice -> [sweeps | count_spins | dump]* -> destroy, with the same argument
meaning as the reference's options (-x, -y per device, -d, -s, -t/-a, --xsl/--ysl).
"""
import ctypes as C
import os

import numpy as np

DEFAULT_SEED = 463463564571            # SEED_DEF, optimized/main.cu:63
CRIT_TEMP = np.float32(2.26918531421)  # CRIT_TEMP, optimized/main.cu:42
ALPHA_DEF = np.float32(0.1)            # ALPHA_DEF, optimized/main.cu:43

BLACK, WHITE = 0, 1
PART_ALL, PART_EDGES, PART_INTERIOR = 0, 1, 2
F_NO_PEER_COPY = 1
F_BONDS = 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class IsingError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ising_b200 error %d: %s" % (code, msg))
        self.code = code


class _Cfg(C.Structure):
    _fields_ = [("X", C.c_int32), ("Y", C.c_int32), ("ndev", C.c_int32), ("first_device", C.c_int32),
                ("seed", C.c_uint64), ("temp", C.c_float), ("xsl", C.c_int32), ("ysl", C.c_int32),
                ("slab_first", C.c_int32), ("slab_count", C.c_int32), ("flags", C.c_int32),
                ("j_prob", C.c_float), ("reserved", C.c_int32 * 4)]


ABI_VERSION = 2      # ISING_B200_ABI_VERSION of include/ising_b200.h


def lib_path():
    # ISING_B200_LIB: developer override to A/B kernel variants; the product is the in-tree build
    return os.environ.get("ISING_B200_LIB") or os.path.join(_HERE, "libising_b200.so")


def load_library():
    """Load libising_b200.so (built in-tree by `make` / __graft_entry__.build()).  No fallback."""
    global _LIB
    if _LIB is not None:
        return _LIB
    p = lib_path()
    if not os.path.exists(p):
        raise IsingError(-100, "%s not built: run `make` (or __graft_entry__.build()) first" % p)
    L = C.CDLL(p)
    vp, u64p, u32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    sig = {
        "ising_create": ([C.POINTER(_Cfg), C.POINTER(vp)], C.c_int),
        "ising_destroy": ([vp], C.c_int),
        "ising_last_error": ([], C.c_char_p),
        "ising_abi_version": ([], C.c_int),
        "ising_set_stream": ([vp, C.c_int, vp], C.c_int),
        "ising_init_lattice": ([vp], C.c_int),
        "ising_update_color": ([vp, C.c_int, C.c_int], C.c_int),
        "ising_sweeps": ([vp, C.c_int, C.c_int], C.c_int),
        "ising_sync": ([vp], C.c_int),
        "ising_sweeps_timed": ([vp, C.c_int, C.c_int, C.POINTER(C.c_float)], C.c_int),
        "ising_set_temperature": ([vp, C.c_float], C.c_int),
        "ising_get_temperature": ([vp, C.POINTER(C.c_float)], C.c_int),
        "ising_count_spins": ([vp, u64p, u64p], C.c_int),
        "ising_bond_sum": ([vp, C.POINTER(C.c_int64)], C.c_int),
        "ising_correlation": ([vp, C.c_int, C.POINTER(C.c_double)], C.c_int),
        "ising_export_bonds_ref_layout": ([vp, C.c_int, vp, vp], C.c_int),
        "ising_export_ref_layout": ([vp, C.c_int, vp, vp], C.c_int),
        "ising_import_ref_layout": ([vp, C.c_int, vp, vp], C.c_int),
        "ising_export_bits": ([vp, C.c_int, vp, vp], C.c_int),
        "ising_import_bits": ([vp, C.c_int, vp, vp], C.c_int),
        "ising_dump_lattice": ([vp, C.c_char_p], C.c_int),
        "ising_save_checkpoint": ([vp, C.c_char_p, C.c_uint64], C.c_int),
        "ising_load_checkpoint": ([vp, C.c_char_p, u64p], C.c_int),
        "ising_run_from_host": ([vp, vp, vp, C.c_int, C.c_int, vp, vp, u64p, u64p], C.c_int),
        "ising_run_batch_from_host": ([vp, vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.c_int, C.c_int,
                                       C.POINTER(vp), C.POINTER(vp), u64p, u64p], C.c_int),
        "ising_run_slab_batch_from_host": ([vp, vp, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.c_int, C.c_int,
                                            C.POINTER(vp), C.POINTER(vp), u64p, u64p], C.c_int),
        "ising_row_ptr": ([vp, C.c_int, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t)], C.c_int),
        "ising_update_color_part": ([vp, C.c_int, C.c_int, C.c_int, C.c_int, vp], C.c_int),
        "ising_host_alloc": ([C.c_size_t, C.c_int, C.POINTER(vp)], C.c_int),
        "ising_run_batch_from_host_n": ([C.POINTER(vp), C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                         C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp), u64p, u64p], C.c_int),
        "ising_host_free": ([vp], C.c_int),
        "ising_device_count": ([C.POINTER(C.c_int)], C.c_int),
        "ising_device_info": ([C.c_int, C.c_char_p, C.c_size_t] + [C.POINTER(C.c_int)] * 5, C.c_int),
        "ising_can_access_peer": ([C.c_int, C.c_int, C.POINTER(C.c_int)], C.c_int),
        "ising_ipc_export": ([vp, vp, C.c_size_t], C.c_int),
        "ising_ipc_attach": ([vp, vp, vp], C.c_int),
        "ising_push_edges": ([vp, C.c_int], C.c_int),
        "ising_accept_count": ([C.c_float, C.c_int], C.c_uint64),
        "ising_accept_table": ([C.c_float, C.POINTER(C.c_float)], None),
        "ising_site_random": ([C.c_uint64, C.c_int32, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_int, u32p], None),
        "ising_launch_count": ([vp], C.c_uint64),
        "ising_tc_create": ([C.c_int, C.c_int, C.c_uint64, C.c_float, C.c_int, C.POINTER(vp)], C.c_int),
        "ising_tc_create_multi": ([C.c_int, C.c_int, C.c_uint64, C.c_float, C.c_int, C.c_int, C.POINTER(vp)], C.c_int),
        "ising_tc_destroy": ([vp], C.c_int),
        "ising_tc_set_nn_path": ([vp, C.c_int], C.c_int),
        "ising_tc_init": ([vp], C.c_int),
        "ising_tc_update": ([vp, C.c_uint64, C.c_int], C.c_int),
        "ising_tc_update_timed": ([vp, C.c_uint64, C.c_int, C.POINTER(C.c_float)], C.c_int),
        "ising_tc_sync": ([vp], C.c_int),
        "ising_tc_sum": ([vp, C.POINTER(C.c_int64)], C.c_int),
        "ising_tc_export_storage": ([vp, vp], C.c_int),
        "ising_tc_import_storage": ([vp, vp], C.c_int),
        "ising_tc_export_true": ([vp, vp], C.c_int),
        "ising_tc_write_lattice": ([vp, C.c_char_p], C.c_int),
        "ising_tc_accept_table": ([vp, C.POINTER(C.c_float)], C.c_int),
        "ising_tc_neighbour_sums": ([vp, C.c_int, vp], C.c_int),
        "ising_tc_launch_count": ([vp], C.c_uint64),
        "ising_kernel_name": ([vp], C.c_char_p),
    }
    for name, (args, res) in sig.items():
        fn = getattr(L, name)      # AttributeError here = header and library out of sync
        fn.argtypes, fn.restype = args, res
    if L.ising_abi_version() != ABI_VERSION:      # a stale build next to newer host code: fail loudly, never guess
        raise IsingError(-101, "%s has ABI version %d, this host code needs %d: rebuild with `make`" %
                         (p, L.ising_abi_version(), ABI_VERSION))
    _LIB = L
    return L


EXPORTED_SYMBOLS = ("ising_create ising_destroy ising_last_error ising_abi_version ising_set_stream "
                    "ising_init_lattice ising_update_color ising_sweeps ising_sync ising_sweeps_timed "
                    "ising_set_temperature ising_get_temperature ising_count_spins ising_bond_sum ising_correlation "
                    "ising_export_bonds_ref_layout ising_export_ref_layout "
                    "ising_import_ref_layout ising_export_bits ising_import_bits ising_dump_lattice "
                    "ising_save_checkpoint ising_load_checkpoint "
                    "ising_run_from_host ising_run_batch_from_host ising_run_slab_batch_from_host ising_row_ptr ising_update_color_part ising_host_alloc ising_host_free ising_run_batch_from_host_n ising_accept_count "
                    "ising_device_count ising_device_info ising_can_access_peer "
                    "ising_ipc_export ising_ipc_attach ising_push_edges "
                    "ising_accept_table ising_site_random ising_launch_count ising_kernel_name "
                    "ising_tc_create ising_tc_create_multi ising_tc_destroy ising_tc_set_nn_path ising_tc_init ising_tc_update "
                    "ising_tc_update_timed ising_tc_sync ising_tc_sum ising_tc_export_storage "
                    "ising_tc_import_storage ising_tc_export_true ising_tc_write_lattice ising_tc_accept_table "
                    "ising_tc_neighbour_sums ising_tc_launch_count").split()


class HostBuffer:
    """Page-locked staging buffer from ising_host_alloc as a numpy uint32 array (`.array`); write_combined buffers are
    for inputs only (the CPU should never read them)."""

    def __init__(self, shape, write_combined=False):
        self._L = load_library()
        n = int(np.prod(shape)) * 4
        p = C.c_void_p()
        rc = self._L.ising_host_alloc(n, int(bool(write_combined)), C.byref(p))
        if rc != 0:
            raise IsingError(rc, self._L.ising_last_error().decode())
        self._p = p
        self.array = np.ctypeslib.as_array((C.c_uint32 * (n // 4)).from_address(p.value)).reshape(shape)

    def close(self):
        if getattr(self, "_p", None):
            self.array = None
            self._L.ising_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def accept_count(p, strict=False):
    return int(load_library().ising_accept_count(C.c_float(float(np.float32(p))), int(bool(strict))))


def accept_table(temp):
    out = (C.c_float * 5)()
    load_library().ising_accept_table(C.c_float(float(np.float32(temp))), out)
    return np.array(list(out), dtype=np.float32)


def site_random(seed, X, it, color, grow, word, q):
    out = (C.c_uint32 * 4)()
    load_library().ising_site_random(int(seed), int(X), int(it), int(color), int(grow), int(word), int(q), out)
    return tuple(int(v) for v in out)


def resolve_temperature(temp=None, alpha=None):
    """-t wins over -a; default alpha 0.1 (optimized/main.cu:1465-1471), all in float32."""
    if temp is not None and np.float32(temp) != np.float32(-1.0):
        return np.float32(temp)
    if alpha is not None and np.float32(alpha) != np.float32(-1.0):
        return np.float32(alpha) * CRIT_TEMP
    return ALPHA_DEF * CRIT_TEMP


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


class IsingEngine:
    """One context = `ndev` slabs of Y rows x X columns driven by this process."""

    def __init__(self, X, Y, ndev=1, seed=DEFAULT_SEED, temp=None, alpha=None, xsl=0, ysl=0,
                 first_device=0, slab_first=0, slab_count=0, flags=0, j_prob=None):
        self._L = load_library()
        self.X, self.Y, self.ndev, self.seed = int(X), int(Y), int(ndev), int(seed)
        self.temp = resolve_temperature(temp, alpha)
        if j_prob is not None:          # -J PROB, clamped as main.cu:1377
            flags |= F_BONDS
            j_prob = min(max(0.0, float(np.float32(j_prob))), 1.0)
        cfg = _Cfg(X=self.X, Y=self.Y, ndev=self.ndev, first_device=int(first_device), seed=self.seed,
                   temp=float(self.temp), xsl=int(xsl), ysl=int(ysl), slab_first=int(slab_first),
                   slab_count=int(slab_count), flags=int(flags), j_prob=float(j_prob or 0.0))
        h = C.c_void_p()
        self._h = None
        self._ck(self._L.ising_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.it = 0                      # sweeps done so far; next sweep uses it+1 (main.cu:1771)
        self.pitch = self.X // 64        # uint32 words per colour row
        self.slab_count = int(slab_count) if slab_count else self.ndev

    # -- plumbing --------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise IsingError(rc, self._L.ising_last_error().decode())

    def close(self):
        if getattr(self, "_h", None):
            self._L.ising_destroy(self._h)
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

    # -- hot path ----------------------------------------------------------------------------
    def init_lattice(self):
        self._ck(self._L.ising_init_lattice(self._h))
        self.it = 0

    def update_color(self, color, it):
        self._ck(self._L.ising_update_color(self._h, int(color), int(it)))

    def update_color_part(self, dev, color, it, part, stream=None):
        self._ck(self._L.ising_update_color_part(self._h, int(dev), int(color), int(it), int(part),
                                                 C.c_void_p(stream) if stream else None))

    def sweeps(self, n, sync=True):
        self._ck(self._L.ising_sweeps(self._h, self.it + 1, int(n)))
        self.it += int(n)
        if sync:
            self.sync()

    def sweeps_timed(self, n):
        ms = C.c_float()
        self._ck(self._L.ising_sweeps_timed(self._h, self.it + 1, int(n), C.byref(ms)))
        self.it += int(n)
        return float(ms.value)

    def sync(self):
        self._ck(self._L.ising_sync(self._h))

    def set_temperature(self, temp):
        self.temp = np.float32(temp)
        self._ck(self._L.ising_set_temperature(self._h, C.c_float(float(self.temp))))

    def set_stream(self, dev, stream):
        self._ck(self._L.ising_set_stream(self._h, int(dev), C.c_void_p(stream)))

    # -- observables ---------------------------------------------------------------------------
    def count_spins(self):
        up, dn = C.c_uint64(), C.c_uint64()
        self._ck(self._L.ising_count_spins(self._h, C.byref(up), C.byref(dn)))
        return int(up.value), int(dn.value)

    def magnetization(self):
        up, dn = self.count_spins()
        return abs(up - dn) / float(up + dn)      # main.cu:1748

    def bond_sum(self):
        s = C.c_int64()
        self._ck(self._L.ising_bond_sum(self._h, C.byref(s)))
        return int(s.value)

    def correlation(self, ncorr=128):
        out = (C.c_double * int(ncorr))()
        self._ck(self._L.ising_correlation(self._h, int(ncorr), out))
        return np.array(list(out))

    def export_bonds_ref_layout(self, dev=0):
        r = np.empty((self.Y, self.X // 32), dtype=np.uint64)
        m = np.empty((self.Y, self.X // 32), dtype=np.uint64)
        self._ck(self._L.ising_export_bonds_ref_layout(self._h, int(dev), _ptr(r), _ptr(m)))
        return r, m

    def energy_per_site(self):
        return -self.bond_sum() / float(self.ndev * self.Y * self.X)

    # -- transfers -----------------------------------------------------------------------------
    def export_bits(self, dev=0):
        b = np.empty((self.Y, self.pitch), dtype=np.uint32)
        w = np.empty((self.Y, self.pitch), dtype=np.uint32)
        self._ck(self._L.ising_export_bits(self._h, int(dev), _ptr(b), _ptr(w)))
        return b, w

    def import_bits(self, black, white, dev=0):
        b = np.ascontiguousarray(black, dtype=np.uint32)
        w = np.ascontiguousarray(white, dtype=np.uint32)
        assert b.shape == (self.Y, self.pitch) and w.shape == (self.Y, self.pitch)
        self._ck(self._L.ising_import_bits(self._h, int(dev), _ptr(b), _ptr(w)))

    def export_ref_layout(self, dev=0):
        b = np.empty((self.Y, self.X // 32), dtype=np.uint64)
        w = np.empty((self.Y, self.X // 32), dtype=np.uint64)
        self._ck(self._L.ising_export_ref_layout(self._h, int(dev), _ptr(b), _ptr(w)))
        return b, w

    def import_ref_layout(self, black, white, dev=0):
        b = np.ascontiguousarray(black, dtype=np.uint64)
        w = np.ascontiguousarray(white, dtype=np.uint64)
        self._ck(self._L.ising_import_ref_layout(self._h, int(dev), _ptr(b), _ptr(w)))

    def dump_lattice(self, prefix):
        self._ck(self._L.ising_dump_lattice(self._h, os.fsencode(prefix)))

    def save_checkpoint(self, path):
        self._ck(self._L.ising_save_checkpoint(self._h, os.fsencode(path), self.it))

    def load_checkpoint(self, path):
        it = C.c_uint64()
        self._ck(self._L.ising_load_checkpoint(self._h, os.fsencode(path), C.byref(it)))
        self.it = int(it.value)
        return self.it

    def run_from_host(self, black_in, white_in, n, black_out, white_out, counts=True):
        up, dn = C.c_uint64(), C.c_uint64()
        self._ck(self._L.ising_run_from_host(self._h, _ptr(black_in), _ptr(white_in), self.it + 1, int(n),
                                             _ptr(black_out), _ptr(white_out),
                                             C.byref(up) if counts else None, C.byref(dn) if counts else None))
        self.it += int(n)
        return int(up.value), int(dn.value)

    def run_batch_from_host(self, other, black_in, white_in, n, black_out=None, white_out=None, first_it=1,
                            ghosts_in=None):
        """Replica batch (ising_run_batch_from_host): lattice k = (black_in[k], white_in[k]) is uploaded,
        evolved by n sweeps from iteration first_it, counted and downloaded; jobs alternate between this
        engine and `other` (same shape and device, or None) so copies overlap kernels.  Returns the
        per-lattice (up, down) counts.  With ghosts_in (slab contexts, fused halo) every rank passes its slab
        of each replica plus its two white ghost rows (ising_run_slab_batch_from_host)."""
        k = len(black_in)
        assert len(white_in) == k

        def ptrs(arrs):
            if arrs is None:
                return None
            return (C.c_void_p * k)(*[(a.ctypes.data if a is not None else None) for a in arrs])
        up, dn = (C.c_uint64 * k)(), (C.c_uint64 * k)()
        if isinstance(other, (list, tuple)):      # more than two contexts: job j runs on [self] + other, round robin
            hs = [self._h] + [o._h for o in other]
            arr = (C.c_void_p * len(hs))(*[h.value if hasattr(h, "value") else h for h in hs])
            if ghosts_in is not None:
                assert len(ghosts_in) == k and all(g.shape == (2, self.pitch) and g.dtype == np.uint32 for g in ghosts_in)
            self._ck(self._L.ising_run_batch_from_host_n(arr, len(hs), k, ptrs(black_in), ptrs(white_in), ptrs(ghosts_in),
                                                         int(first_it), int(n), ptrs(black_out), ptrs(white_out), up, dn))
            return [(int(up[i]), int(dn[i])) for i in range(k)]
        oh = other._h if other is not None else None
        if ghosts_in is None:
            self._ck(self._L.ising_run_batch_from_host(self._h, oh, k, ptrs(black_in), ptrs(white_in), int(first_it),
                                                       int(n), ptrs(black_out), ptrs(white_out), up, dn))
        else:   # one slab of a multi-process lattice: ghosts_in[k] = (2, pitch) white rows above / below the slab
            assert len(ghosts_in) == k and all(g.shape == (2, self.pitch) and g.dtype == np.uint32 for g in ghosts_in)
            self._ck(self._L.ising_run_slab_batch_from_host(self._h, oh, k, ptrs(black_in), ptrs(white_in),
                                                            ptrs(ghosts_in), int(first_it), int(n), ptrs(black_out),
                                                            ptrs(white_out), up, dn))
        return [(int(up[i]), int(dn[i])) for i in range(k)]

    def ipc_export(self):
        buf = C.create_string_buffer(192)
        self._ck(self._L.ising_ipc_export(self._h, buf, 192))
        return buf.raw

    def ipc_attach(self, up_handles, down_handles):
        self._ck(self._L.ising_ipc_attach(self._h, C.c_char_p(up_handles), C.c_char_p(down_handles)))

    def push_edges(self, color):
        self._ck(self._L.ising_push_edges(self._h, int(color)))

    def row_ptr(self, dev, color, row):
        p, n = C.c_void_p(), C.c_size_t()
        self._ck(self._L.ising_row_ptr(self._h, int(dev), int(color), int(row), C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    @property
    def launches(self):
        return int(self._L.ising_launch_count(self._h))

    @property
    def kernel_name(self):
        return self._L.ising_kernel_name(self._h).decode()
