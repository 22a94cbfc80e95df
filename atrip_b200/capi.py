"""ctypes binding of the C-ABI in include/atrip_b200.h (one Engine == one atrip_b200_ctx).

Mirrors the header one to one; every failure of the library raises EngineError carrying
atrip_b200_last_error().  Nothing here computes: if libatrip_b200.so is missing, loading raises.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")

# symbols declared in include/atrip_b200.h (tests check the library exports every one)
SYMBOLS = [
    "atrip_b200_create", "atrip_b200_destroy", "atrip_b200_last_error", "atrip_b200_version",
    "atrip_b200_set_epsilon", "atrip_b200_set_Tai", "atrip_b200_load_Tabij", "atrip_b200_load_Vabij",
    "atrip_b200_load_Vijka", "atrip_b200_load_Vabci", "atrip_b200_load_Jijka", "atrip_b200_load_Jabci",
    "atrip_b200_fill_synthetic", "atrip_b200_build_tuples", "atrip_b200_set_tuples",
    "atrip_b200_num_tuples", "atrip_b200_get_tuples", "atrip_b200_run", "atrip_b200_tuple_debug",
    "atrip_b200_read_slice", "atrip_b200_last_timing", "atrip_b200_kp", "atrip_b200_flops_per_tuple",
    "atrip_b200_host_tuples", "atrip_b200_host_slice_owner", "atrip_b200_measure_dmma_peak",
    "atrip_b200_synth_to_host", "atrip_b200_batch_tuples", "atrip_b200_host_plan", "atrip_b200_device_count",
    "atrip_b200_comm_unique_id", "atrip_b200_comm_init", "atrip_b200_allreduce", "atrip_b200_last_exchange",
    "atrip_b200_host_slice_slot", "atrip_b200_host_shard_sizes", "atrip_b200_host_plan_batch",
    "atrip_b200_host_cache_need", "atrip_b200_host_local_slot", "atrip_b200_host_store_source",
    "atrip_b200_host_energy_z", "atrip_b200_debug_cubes_checksum",
    "atrip_b200_upload_slices", "atrip_b200_upload_slice", "atrip_b200_read_slices",
    "atrip_b200_host_owned_slices", "atrip_b200_last_phases", "atrip_b200_host_check_schedule",
]

NAIVE, GROUP_AND_SORT = 0, 1
FIELD_REAL, FIELD_COMPLEX = 0, 1
TRANSPORT_DEFAULT, TRANSPORT_NCCL, TRANSPORT_P2P = 0, 1, 2
TA, VIJKA, VABCI, TABIJ, VABIJ = 100, 101, 200, 201, 202
JIJKA, JABCI = 111, 210  # (cT) tensors in the slice kinds of the upload / read-back entry points
VABCI_T = 203  # host-side name of the transposed-hole twin (x,x)' of a diagonal pair slice


class EngineError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("with_J", C.c_int32),
                ("No", C.c_int64), ("Nv", C.c_int64), ("batch_tuples", C.c_int64),
                ("resident", C.c_int32), ("transport", C.c_int32), ("field", C.c_int32)]


def lib_path():
    # ATRIP_B200_LIB: developer A/B builds of the same library (tools/ab_build.sh)
    return os.environ.get("ATRIP_B200_LIB") or os.path.join(CSRC, "libatrip_b200.so")


def build_library(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a (cross-compiles without a GPU)"""
    cmd = ["make", "-C", CSRC] + (["-B"] if force else []) + ([] if verbose else ["-s"])
    subprocess.check_call(cmd)
    return lib_path()


_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint64)
_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise EngineError(f"{path} is missing: run `make -C atrip_b200/csrc` (or __graft_entry__.build()); "
                          "there is no CPU fallback")
    L = C.CDLL(path)
    ctx = C.c_void_p
    L.atrip_b200_create.argtypes = [C.POINTER(ctx), C.POINTER(Config)]
    L.atrip_b200_destroy.argtypes = [ctx]
    L.atrip_b200_last_error.restype = C.c_char_p
    L.atrip_b200_version.restype = C.c_char_p
    L.atrip_b200_set_epsilon.argtypes = [ctx, _dp, _dp]
    for n in ("set_Tai", "load_Tabij", "load_Vabij", "load_Vijka", "load_Vabci", "load_Jijka", "load_Jabci"):
        getattr(L, "atrip_b200_" + n).argtypes = [ctx, _dp]
    L.atrip_b200_fill_synthetic.argtypes = [ctx, C.c_uint64, C.c_double]
    L.atrip_b200_build_tuples.argtypes = [ctx, C.c_int32]
    L.atrip_b200_set_tuples.argtypes = [ctx, _up, C.c_int64]
    L.atrip_b200_num_tuples.argtypes = [ctx]
    L.atrip_b200_num_tuples.restype = C.c_int64
    L.atrip_b200_get_tuples.argtypes = [ctx, _up, C.c_int64]
    L.atrip_b200_run.argtypes = [ctx, C.c_int64, C.c_int64, _dp, _dp]
    L.atrip_b200_tuple_debug.argtypes = [ctx, C.c_int64, C.c_int64, C.c_int64, _dp, _dp, _dp]
    L.atrip_b200_read_slice.argtypes = [ctx, C.c_int32, C.c_int64, C.c_int64, _dp]
    L.atrip_b200_last_timing.argtypes = [ctx, _dp]
    L.atrip_b200_kp.argtypes = [ctx]
    L.atrip_b200_kp.restype = C.c_int64
    L.atrip_b200_batch_tuples.argtypes = [ctx]
    L.atrip_b200_batch_tuples.restype = C.c_int64
    L.atrip_b200_flops_per_tuple.argtypes = [ctx]
    L.atrip_b200_flops_per_tuple.restype = C.c_double
    L.atrip_b200_host_tuples.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _up, C.c_int64]
    L.atrip_b200_host_tuples.restype = C.c_int64
    L.atrip_b200_host_slice_owner.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32]
    L.atrip_b200_host_slice_owner.restype = C.c_int32
    _ip = C.POINTER(C.c_int64)
    L.atrip_b200_host_slice_slot.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, _ip]
    L.atrip_b200_host_slice_slot.restype = C.c_int32
    L.atrip_b200_host_local_slot.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.c_int64, C.c_int32, C.c_int32]
    L.atrip_b200_host_local_slot.restype = C.c_int64
    L.atrip_b200_host_shard_sizes.argtypes = [C.c_int64, C.c_int32, C.c_int32, _ip]
    L.atrip_b200_host_plan_batch.argtypes = [C.c_int64, C.c_int32, C.c_int32, _up, C.c_int64, _ip,
                                             C.POINTER(C.c_int32), _ip, C.c_int64]
    L.atrip_b200_host_plan_batch.restype = C.c_int64
    L.atrip_b200_host_cache_need.argtypes = [C.c_int64, C.c_int32, C.c_int32, _up, C.c_int64, C.c_int64, _ip]
    L.atrip_b200_host_check_schedule.argtypes = [C.c_int64, C.c_int32, C.c_int32, _up, C.c_int64, C.c_int64, C.c_int32,
                                                 _ip, _dp]
    L.atrip_b200_upload_slices.argtypes = [ctx, C.c_int32, C.c_int64, _ip, _dp]
    L.atrip_b200_upload_slice.argtypes = [ctx, C.c_int32, C.c_int64, C.c_int64, _dp]
    L.atrip_b200_read_slices.argtypes = [ctx, C.c_int32, C.c_int64, _ip, _dp]
    L.atrip_b200_host_owned_slices.argtypes = [C.c_int32, C.c_int64, C.c_int32, C.c_int32, _ip, C.c_int64]
    L.atrip_b200_host_owned_slices.restype = C.c_int64
    L.atrip_b200_last_phases.argtypes = [ctx, _dp]
    L.atrip_b200_comm_unique_id.argtypes = [C.c_void_p]
    L.atrip_b200_comm_init.argtypes = [ctx, C.c_void_p]
    L.atrip_b200_allreduce.argtypes = [ctx, _dp, C.c_int32]
    L.atrip_b200_last_exchange.argtypes = [ctx, _dp]
    L.atrip_b200_host_store_source.argtypes = [C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int64, C.c_int64,
                                               C.c_int64, C.c_int64, _dp]
    L.atrip_b200_host_energy_z.argtypes = [C.c_int64, C.c_double, _dp, _dp, _dp, C.c_int32]
    L.atrip_b200_host_energy_z.restype = C.c_double
    L.atrip_b200_measure_dmma_peak.argtypes = [C.c_int32, _dp]
    L.atrip_b200_synth_to_host.argtypes = [C.c_int32, C.c_uint64, C.c_int32, C.c_double, C.c_uint64, C.c_uint64, _dp]
    _lib = L
    return L


def measure_dmma_peak(device=0):
    """FP64 tensor-core ceiling of the device in TFLOP/s (register-resident DMMA loop)"""
    L = load_library()
    v = C.c_double(0)
    if L.atrip_b200_measure_dmma_peak(device, C.byref(v)) != 0:
        raise EngineError(L.atrip_b200_last_error().decode())
    return v.value


def synth_to_host(device, seed, tensor_id, scale, first, count, host):
    """fill host[0:count] (numpy array or int address) with the synthetic tensor values"""
    L = load_library()
    if L.atrip_b200_synth_to_host(device, seed, tensor_id, scale, first, count, _ptr(host)) != 0:
        raise EngineError(L.atrip_b200_last_error().decode())


def host_plan(No, smem_limit=0):
    """the contraction kernel's tile plan for No (host-only)"""
    L = load_library()
    L.atrip_b200_host_plan.argtypes = [C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
    out = (C.c_int64 * 11)()
    if L.atrip_b200_host_plan(No, smem_limit, out) != 0:
        raise EngineError(L.atrip_b200_last_error().decode())
    k = ["MI", "NI", "warps", "tu", "tv", "row_tiles", "col_tiles", "stages", "smem", "arows", "useful"]
    d = dict(zip(k, list(out)))
    d["useful"] /= 1e6
    return d


def host_tuples(distribution, Nv, rank=0, nranks=1, pad=True):
    """tuple list of one rank, computed on the host by the library (no GPU needed)"""
    L = load_library()
    n = L.atrip_b200_host_tuples(distribution, Nv, rank, nranks, int(pad), None, 0)
    if n < 0:
        raise EngineError(L.atrip_b200_last_error().decode())
    out = np.empty((n, 3), dtype=np.uint64)
    L.atrip_b200_host_tuples(distribution, Nv, rank, nranks, int(pad), out.ctypes.data_as(_up), n)
    return out


def slice_owner(kind, x, y, Nv, nranks):
    return load_library().atrip_b200_host_slice_owner(kind, x, y, Nv, nranks)


def slice_slot(kind, x, y, Nv, nranks):
    """(owner rank, slot in the owner's store) of a slice"""
    s = C.c_int64(-1)
    o = load_library().atrip_b200_host_slice_slot(kind, x, y, Nv, nranks, C.byref(s))
    return o, s.value


def local_slot(kind, x, y, Nv, rank, nranks):
    """slot of a slice in the store of `rank`, or -1"""
    return load_library().atrip_b200_host_local_slot(kind, x, y, Nv, rank, nranks)


def shard_sizes(Nv, rank, nranks):
    """owned slots (AX, BY, VIJ) of one rank"""
    L = load_library()
    out = (C.c_int64 * 3)()
    if L.atrip_b200_host_shard_sizes(Nv, rank, nranks, out) != 0:
        raise EngineError(L.atrip_b200_last_error().decode())
    return tuple(out)


def plan_batch(Nv, rank, nranks, abc, cache_base):
    """host fetch schedule of one batch: (recs [n,16] int32, ranges [m,5] int64 =
    peer, store, first slot at owner, count, first cache slot in the region)"""
    L = load_library()
    abc = np.ascontiguousarray(abc, dtype=np.uint64).reshape(-1, 3)
    n = len(abc)
    base = (C.c_int64 * 3)(*[int(b) for b in cache_base])
    recs = np.zeros((n, 16), dtype=np.int32)
    cap = 12 * max(n, 1)
    ranges = np.zeros((cap, 5), dtype=np.int64)
    m = L.atrip_b200_host_plan_batch(Nv, rank, nranks, abc.ctypes.data_as(_up), n, base,
                                     recs.ctypes.data_as(C.POINTER(C.c_int32)),
                                     ranges.ctypes.data_as(C.POINTER(C.c_int64)), cap)
    if m < 0:
        raise EngineError(L.atrip_b200_last_error().decode())
    return recs, ranges[:m]


def cache_need(Nv, rank, nranks, abc, batch):
    L = load_library()
    abc = np.ascontiguousarray(abc, dtype=np.uint64).reshape(-1, 3)
    out = (C.c_int64 * 3)()
    if L.atrip_b200_host_cache_need(Nv, rank, nranks, abc.ctypes.data_as(_up), len(abc), batch, out) != 0:
        raise EngineError(L.atrip_b200_last_error().decode())
    return tuple(out)


def check_schedule(Nv, rank, nranks, abc, batch, caps, calls=1):
    """walk a tuple list through the persistent fetch cache as atrip_b200_run does and check it against an
    independent model (host-only); returns the traffic statistics, raises EngineError on a violation"""
    L = load_library()
    abc = np.ascontiguousarray(abc, dtype=np.uint64).reshape(-1, 3)
    cap = (C.c_int64 * 3)(*[int(x) for x in caps])
    out = (C.c_double * 8)()
    if L.atrip_b200_host_check_schedule(Nv, rank, nranks, abc.ctypes.data_as(_up), len(abc), batch, calls, cap, out) != 0:
        raise EngineError(L.atrip_b200_last_error().decode())
    return dict(fetched=[int(out[i]) for i in range(3)], hits=[int(out[3 + i]) for i in range(3)],
                ranges=int(out[6]), batches=int(out[7]))


def owned_slices(kind, Nv, rank, nranks):
    """(x, y) list [n,2] int64 of the slices of `kind` a rank has to be given (host-only)"""
    L = load_library()
    n = L.atrip_b200_host_owned_slices(kind, Nv, rank, nranks, None, 0)
    if n < 0:
        raise EngineError(L.atrip_b200_last_error().decode())
    out = np.zeros((n, 2), dtype=np.int64)
    L.atrip_b200_host_owned_slices(kind, Nv, rank, nranks, out.ctypes.data_as(C.POINTER(C.c_int64)), n)
    return out


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 creates it, the host ships it to the other ranks)"""
    L = load_library()
    buf = C.create_string_buffer(128)
    if L.atrip_b200_comm_unique_id(buf) != 0:
        raise EngineError(L.atrip_b200_last_error().decode())
    return buf.raw


def _ptr(a):
    """double* of a host buffer: numpy float64 (or complex128 = interleaved doubles) array, or an int
    address (e.g. pinned torch storage)"""
    if a is None:
        return None
    if isinstance(a, int):
        return C.cast(a, _dp)
    assert a.dtype in (np.float64, np.complex128) and (a.flags["C_CONTIGUOUS"] or a.flags["F_CONTIGUOUS"])
    return a.ctypes.data_as(_dp)


def store_source(store, No, Nv, a, x, y, row, kappa):
    """complex layout: (tensor, part, sign, element) held by a store element (host-only)"""
    L = load_library()
    out = (C.c_double * 4)()
    if L.atrip_b200_host_store_source(store, No, Nv, a, x, y, row, kappa, out) != 0:
        raise EngineError(L.atrip_b200_last_error().decode())
    return int(out[0]), int(out[1]), out[2], int(out[3])


def host_energy_z(No, epsabc, eps_i, Tijk, Zijk, same):
    """complex tuple energy through the device kernel's per-point function, on the host"""
    L = load_library()
    return L.atrip_b200_host_energy_z(No, epsabc, _ptr(eps_i), _ptr(Tijk), _ptr(Zijk), int(same))


class Engine:
    def __init__(self, No, Nv, device=0, rank=0, nranks=1, with_J=False, batch_tuples=0, resident=True,
                 transport=0, field=FIELD_REAL):
        self.L = load_library()
        self.No, self.Nv = int(No), int(Nv)
        self.field = int(field)
        self.dtype = np.complex128 if self.field == FIELD_COMPLEX else np.float64
        cfg = Config(device, rank, nranks, int(with_J), No, Nv, batch_tuples, int(resident), int(transport),
                     self.field)
        self.ctx = C.c_void_p()
        self._ck(self.L.atrip_b200_create(C.byref(self.ctx), C.byref(cfg)))

    def _ck(self, rc):
        if rc != 0:
            raise EngineError(self.L.atrip_b200_last_error().decode())

    def close(self):
        if getattr(self, "ctx", None) and self.ctx.value:
            self.L.atrip_b200_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- data
    def set_epsilon(self, eps_i, eps_a):
        self._ck(self.L.atrip_b200_set_epsilon(self.ctx, _ptr(eps_i), _ptr(eps_a)))

    def set_Tai(self, Tai):
        self._ck(self.L.atrip_b200_set_Tai(self.ctx, _ptr(Tai)))

    def load(self, name, host):
        self._ck(getattr(self.L, "atrip_b200_load_" + name)(self.ctx, _ptr(host)))

    def load_all(self, eps_i, eps_a, Tai, Tabij, Vabij, Vijka, Vabci, Jijka=None, Jabci=None):
        self.set_epsilon(eps_i, eps_a)
        self.set_Tai(Tai)
        self.load("Tabij", Tabij)
        self.load("Vabij", Vabij)
        self.load("Vijka", Vijka)
        self.load("Vabci", Vabci)
        if Jijka is not None:
            self.load("Jijka", Jijka)
        if Jabci is not None:
            self.load("Jabci", Jabci)

    def fill_synthetic(self, seed=12345, scale=0.1):
        self._ck(self.L.atrip_b200_fill_synthetic(self.ctx, seed, scale))

    # ---- tuples
    def build_tuples(self, distribution=GROUP_AND_SORT):
        self._ck(self.L.atrip_b200_build_tuples(self.ctx, distribution))
        return self.num_tuples()

    def set_tuples(self, abc):
        abc = np.ascontiguousarray(abc, dtype=np.uint64).reshape(-1, 3)
        self._ck(self.L.atrip_b200_set_tuples(self.ctx, abc.ctypes.data_as(_up), len(abc)))

    def num_tuples(self):
        return self.L.atrip_b200_num_tuples(self.ctx)

    def get_tuples(self):
        n = self.num_tuples()
        out = np.empty((n, 3), dtype=np.uint64)
        self._ck(self.L.atrip_b200_get_tuples(self.ctx, out.ctypes.data_as(_up), n))
        return out

    # ---- execute
    def run(self, first=0, count=None):
        if count is None:
            count = self.num_tuples() - first
        e, ct = C.c_double(0), C.c_double(0)
        self._ck(self.L.atrip_b200_run(self.ctx, first, count, C.byref(e), C.byref(ct)))
        return e.value, ct.value

    def tuple_debug(self, a, b, c, cubes=True):
        n = self.No ** 3
        T = np.empty(n, dtype=self.dtype) if cubes else None
        Z = np.empty(n, dtype=self.dtype) if cubes else None
        e = C.c_double(0)
        self._ck(self.L.atrip_b200_tuple_debug(self.ctx, a, b, c, _ptr(T), _ptr(Z), C.byref(e)))
        return e.value, T, Z

    def read_slice(self, kind, x, y=0):
        No, Nv = self.No, self.Nv
        n = {TA: Nv * No * No, VIJKA: No ** 3, VABCI: Nv * No, TABIJ: No * No, VABIJ: No * No}[kind]
        out = np.empty(n, dtype=self.dtype)
        self._ck(self.L.atrip_b200_read_slice(self.ctx, kind, x, y, _ptr(out)))
        return out

    def slice_elems(self, kind):
        No, Nv = self.No, self.Nv
        return {TA: Nv * No * No, VIJKA: No ** 3, JIJKA: No ** 3, VABCI: Nv * No, JABCI: Nv * No, TABIJ: No * No,
                VABIJ: No * No}[kind]

    def upload_slices(self, kind, xy, host):
        """n slices of one kind in the reference's slice layout, back to back (numpy array or address)"""
        xy = np.ascontiguousarray(xy, dtype=np.int64).reshape(-1, 2)
        self._ck(self.L.atrip_b200_upload_slices(self.ctx, kind, len(xy), xy.ctypes.data_as(C.POINTER(C.c_int64)),
                                                 _ptr(host)))

    def upload_slice(self, kind, x, y, host):
        self._ck(self.L.atrip_b200_upload_slice(self.ctx, kind, x, y, _ptr(host)))

    def read_slices(self, kind, xy, out=None):
        xy = np.ascontiguousarray(xy, dtype=np.int64).reshape(-1, 2)
        if out is None:
            out = np.empty(len(xy) * self.slice_elems(kind), dtype=self.dtype)
        self._ck(self.L.atrip_b200_read_slices(self.ctx, kind, len(xy), xy.ctypes.data_as(C.POINTER(C.c_int64)),
                                               _ptr(out)))
        return out

    def last_phases(self):
        out = (C.c_double * 6)()
        self.L.atrip_b200_last_phases(self.ctx, out)
        return dict(gap_ms=out[0], plan_ms=out[1], cache_hits=int(out[2]), fetched=int(out[3]), batches=int(out[4]),
                    startup_gap_ms=out[5])

    # ---- multi-GPU
    def comm_init(self, unique_id):
        assert len(unique_id) == 128
        self._ck(self.L.atrip_b200_comm_init(self.ctx, C.create_string_buffer(bytes(unique_id), 128)))

    def allreduce(self, vals):
        a = np.ascontiguousarray(vals, dtype=np.float64)
        self._ck(self.L.atrip_b200_allreduce(self.ctx, _ptr(a), len(a)))
        return a

    def last_exchange(self):
        out = (C.c_double * 2)()
        self.L.atrip_b200_last_exchange(self.ctx, out)
        return dict(bytes=out[0], messages=int(out[1]))

    def cubes_checksum(self):
        """debug: integer checksum of the class cubes of the last batch"""
        h = C.c_uint64(0)
        self.L.atrip_b200_debug_cubes_checksum.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        self._ck(self.L.atrip_b200_debug_cubes_checksum(self.ctx, C.byref(h)))
        return h.value

    def last_timing(self):
        out = (C.c_double * 6)()
        self.L.atrip_b200_last_timing(self.ctx, out)
        return dict(total_ms=out[0], contract_ms=out[1], reduce_ms=out[2], contract_launches=int(out[3]),
                    reduce_launches=int(out[4]), tuples=int(out[5]))

    @property
    def kp(self):
        return self.L.atrip_b200_kp(self.ctx)

    @property
    def batch_tuples(self):
        return self.L.atrip_b200_batch_tuples(self.ctx)

    @property
    def flops_per_tuple(self):
        return self.L.atrip_b200_flops_per_tuple(self.ctx)
