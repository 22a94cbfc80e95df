"""TEST INFRASTRUCTURE ONLY: ctypes bindings of the oracle.

``Oracle``   -> oracle/liboracle.so, the plain-C restatement (atrip_oracle.c)
``Reference`` -> oracle/_ref/libatrip_ref{,_loops}.so, the reference's own unmodified
                 sources compiled by oracle/Makefile (present only after
                 ``make -C oracle`` ran in a container that has /root/reference;
                 the prebuilt .so travels to the GPU box)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (atrip_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
EPS_I, EPS_A, TAI, TABIJ, VABIJ, VIJKA, VABCI, JIJKA, JABCI = range(9)

_dp = C.POINTER(C.c_double)
_up = C.POINTER(C.c_uint64)


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def build(force=False):
    """compile liboracle.so (and oracle/_ref when /root/reference is present)"""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def tensor_sizes(No, Nv):
    return {
        EPS_I: No, EPS_A: Nv, TAI: Nv * No, TABIJ: Nv * Nv * No * No,
        VABIJ: Nv * Nv * No * No, VIJKA: No * No * No * Nv, VABCI: Nv * Nv * Nv * No,
        JIJKA: No * No * No * Nv, JABCI: Nv * Nv * Nv * No,
    }


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = self.L = C.CDLL(path)
        L.oracle_synth.restype = C.c_double
        L.oracle_synth.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_double]
        L.oracle_fill.argtypes = [C.c_uint64, C.c_int, C.c_double, C.c_uint64, C.c_uint64, _dp]
        for f in (L.oracle_slice_TA, L.oracle_slice_HHHA):
            f.argtypes = [C.c_long, C.c_long, _dp, C.c_long, _dp]
        for f in (L.oracle_slice_ABPH, L.oracle_slice_ABHH):
            f.argtypes = [C.c_long, C.c_long, _dp, C.c_long, C.c_long, _dp]
        L.oracle_doubles.argtypes = [C.c_long, C.c_long] + [_dp] * 16
        L.oracle_singles.argtypes = [C.c_long] * 5 + [_dp] * 5
        for f in (L.oracle_energy_distinct, L.oracle_energy_same):
            f.restype = C.c_double
            f.argtypes = [C.c_double, C.c_long, _dp, _dp, _dp]
        L.oracle_tuple_energy.restype = C.c_double
        L.oracle_tuple_energy.argtypes = [C.c_long, C.c_long] + [_dp] * 9 + [C.c_long] * 3 + [_dp] * 3
        L.oracle_n_tuples.restype = C.c_long
        L.oracle_n_tuples.argtypes = [C.c_long]
        L.oracle_all_tuples.restype = C.c_long
        L.oracle_all_tuples.argtypes = [C.c_long, _up, C.c_long]
        L.oracle_group_and_sort.restype = C.c_long
        L.oracle_group_and_sort.argtypes = [C.c_long, C.c_long, C.c_long, _up, C.c_long]
        L.oracle_owner_single.restype = C.c_long
        L.oracle_owner_single.argtypes = [C.c_long, C.c_long]
        L.oracle_owner_pair.restype = C.c_long
        L.oracle_owner_pair.argtypes = [C.c_long] * 4
        L.oracle_run.argtypes = [C.c_long, C.c_long] + [_dp] * 9 + [_up, C.c_long, _dp, _dp]
        # F = Complex (atrip_oracle_z.c): complex128 arrays passed as interleaved doubles
        L.oracle_fill_z.argtypes = [C.c_uint64, C.c_int, C.c_double, C.c_uint64, C.c_uint64, _dp]
        L.oracle_doubles_z.argtypes = [C.c_long, C.c_long] + [_dp] * 16
        L.oracle_singles_z.argtypes = [C.c_long] * 5 + [_dp] * 5
        for f in (L.oracle_energy_distinct_z, L.oracle_energy_same_z):
            f.restype = C.c_double
            f.argtypes = [C.c_double, C.c_long, _dp, _dp, _dp]
        L.oracle_tuple_energy_z.restype = C.c_double
        L.oracle_tuple_energy_z.argtypes = [C.c_long, C.c_long] + [_dp] * 9 + [C.c_long] * 3 + [_dp] * 3
        L.oracle_run_z.argtypes = [C.c_long, C.c_long] + [_dp] * 9 + [_up, C.c_long, _dp, _dp]

    # ---- inputs
    def fill(self, seed, tensor_id, scale, count, first=0):
        out = np.empty(count, dtype=np.float64)
        self.L.oracle_fill(seed, tensor_id, scale, first, count, _d(out))
        return out

    def inputs(self, No, Nv, seed=12345, scale=0.1, with_J=False):
        """dict of the seven (nine with J) full tensors, flat column-major"""
        ids = [EPS_I, EPS_A, TAI, TABIJ, VABIJ, VIJKA, VABCI] + ([JIJKA, JABCI] if with_J else [])
        sz = tensor_sizes(No, Nv)
        return {t: self.fill(seed, t, scale, sz[t]) for t in ids}

    # ---- F = Complex: inputs are complex128 arrays (interleaved re, im in memory)
    def fill_z(self, seed, tensor_id, scale, count, first=0):
        out = np.empty(count, dtype=np.complex128)
        self.L.oracle_fill_z(seed, tensor_id, scale, first, count, _d(out))
        return out

    def inputs_z(self, No, Nv, seed=12345, scale=0.1, with_J=False):
        ids = [EPS_I, EPS_A, TAI, TABIJ, VABIJ, VIJKA, VABCI] + ([JIJKA, JABCI] if with_J else [])
        sz = tensor_sizes(No, Nv)
        return {t: self.fill_z(seed, t, scale, sz[t]) for t in ids}

    @staticmethod
    def tuple_slices_z(No, Nv, t, abc, J=False):
        """the 18 complex slices of one tuple (numpy views of the column-major tensors)"""
        a, b, c = abc
        T = t[TABIJ].reshape((Nv, Nv, No, No), order="F")
        Vij = t[VABIJ].reshape((Nv, Nv, No, No), order="F")
        Vp = t[JABCI if J else VABCI].reshape((Nv, Nv, Nv, No), order="F")
        Vh = t[JIJKA if J else VIJKA].reshape((No, No, No, Nv), order="F")
        flat = lambda x: np.ascontiguousarray(x.reshape(-1, order="F"))
        S = {}
        for nm, (x, y) in dict(VAB=(a, b), VAC=(a, c), VBC=(b, c), VBA=(b, a), VCA=(c, a), VCB=(c, b)).items():
            S[nm] = flat(Vp[x, y])
        for nm, x in dict(HA=a, HB=b, HC=c).items():
            S[nm] = flat(Vh[:, :, :, x])
        for nm, x in dict(TA=a, TB=b, TC=c).items():
            S[nm] = flat(T[x])
        for nm, (x, y) in dict(TAB=(a, b), TAC=(a, c), TBC=(b, c)).items():
            S[nm] = flat(T[x, y])
        for nm, (x, y) in dict(VABij=(a, b), VACij=(a, c), VBCij=(b, c)).items():
            S[nm] = flat(Vij[x, y])
        return S

    def doubles_z(self, No, Nv, S):
        out = np.empty(No ** 3, dtype=np.complex128)
        self.L.oracle_doubles_z(No, Nv, *[_d(S[k]) for k in self.DOUBLES_ORDER], _d(out))
        return out

    def singles_z(self, No, Nv, abc, Tai, S, Tijk):
        Z = Tijk.copy()
        self.L.oracle_singles_z(No, Nv, abc[0], abc[1], abc[2], _d(Tai), _d(S["VABij"]),
                                _d(S["VACij"]), _d(S["VBCij"]), _d(Z))
        return Z

    def energy_distinct_z(self, epsabc, No, epsi, Tijk, Zijk):
        return self.L.oracle_energy_distinct_z(epsabc, No, _d(epsi), _d(Tijk), _d(Zijk))

    def energy_same_z(self, epsabc, No, epsi, Tijk, Zijk):
        return self.L.oracle_energy_same_z(epsabc, No, _d(epsi), _d(Tijk), _d(Zijk))

    def tuple_energy_z(self, No, Nv, t, abc, want_cubes=False):
        T = np.empty(No ** 3, dtype=np.complex128) if want_cubes else None
        Z = np.empty(No ** 3, dtype=np.complex128) if want_cubes else None
        ct = C.c_double(0)
        e = self.L.oracle_tuple_energy_z(
            No, Nv, _d(t[EPS_I]), _d(t[EPS_A]), _d(t[TAI]), _d(t[TABIJ]), _d(t[VABIJ]),
            _d(t[VIJKA]), _d(t[VABCI]), _d(t.get(JIJKA)), _d(t.get(JABCI)),
            abc[0], abc[1], abc[2], _d(T), _d(Z), C.cast(C.byref(ct), _dp))
        return (e, ct.value, T, Z) if want_cubes else (e, ct.value)

    def run_z(self, No, Nv, t, tuples=None):
        e, ct = C.c_double(0), C.c_double(0)
        tp, n = (None, 0)
        if tuples is not None:
            tuples = np.ascontiguousarray(tuples, dtype=np.uint64)
            tp, n = tuples.ctypes.data_as(_up), len(tuples)
        self.L.oracle_run_z(No, Nv, _d(t[EPS_I]), _d(t[EPS_A]), _d(t[TAI]), _d(t[TABIJ]),
                            _d(t[VABIJ]), _d(t[VIJKA]), _d(t[VABCI]), _d(t.get(JIJKA)),
                            _d(t.get(JABCI)), tp, n, C.cast(C.byref(e), _dp), C.cast(C.byref(ct), _dp))
        return e.value, ct.value

    # ---- slices
    def slice_TA(self, No, Nv, Tabij, x):
        out = np.empty(Nv * No * No)
        self.L.oracle_slice_TA(No, Nv, _d(Tabij), x, _d(out))
        return out

    def slice_HHHA(self, No, Nv, Vijka, x):
        out = np.empty(No ** 3)
        self.L.oracle_slice_HHHA(No, Nv, _d(Vijka), x, _d(out))
        return out

    def slice_ABPH(self, No, Nv, Vabci, x, y):
        out = np.empty(Nv * No)
        self.L.oracle_slice_ABPH(No, Nv, _d(Vabci), x, y, _d(out))
        return out

    def slice_ABHH(self, No, Nv, Vabij, x, y):
        out = np.empty(No * No)
        self.L.oracle_slice_ABHH(No, Nv, _d(Vabij), x, y, _d(out))
        return out

    def tuple_slices(self, No, Nv, t, abc, J=False):
        """the 18 slices of one tuple in reference argument order"""
        a, b, c = abc
        vp = t[JABCI] if J else t[VABCI]
        vh = t[JIJKA] if J else t[VIJKA]
        S = {}
        for nm, (x, y) in dict(VAB=(a, b), VAC=(a, c), VBC=(b, c), VBA=(b, a), VCA=(c, a), VCB=(c, b)).items():
            S[nm] = self.slice_ABPH(No, Nv, vp, x, y)
        for nm, x in dict(HA=a, HB=b, HC=c).items():
            S[nm] = self.slice_HHHA(No, Nv, vh, x)
        for nm, x in dict(TA=a, TB=b, TC=c).items():
            S[nm] = self.slice_TA(No, Nv, t[TABIJ], x)
        for nm, (x, y) in dict(TAB=(a, b), TAC=(a, c), TBC=(b, c)).items():
            S[nm] = self.slice_ABHH(No, Nv, t[TABIJ], x, y)
        for nm, (x, y) in dict(VABij=(a, b), VACij=(a, c), VBCij=(b, c)).items():
            S[nm] = self.slice_ABHH(No, Nv, t[VABIJ], x, y)
        return S

    SLICE_ORDER = ["VAB", "VAC", "VBC", "VBA", "VCA", "VCB", "HA", "HB", "HC", "TA", "TB", "TC",
                   "TAB", "TAC", "TBC", "VABij", "VACij", "VBCij"]

    def synth_tuple_slices(self, No, Nv, abc, seed=12345, scale=0.1):
        """the 18 slices of one tuple straight from the generator (no full tensors needed)"""
        sizes = [Nv * No] * 6 + [No ** 3] * 3 + [Nv * No * No] * 3 + [No * No] * 6
        S = {k: np.empty(n) for k, n in zip(self.SLICE_ORDER, sizes)}
        arr = (_dp * 18)(*[_d(S[k]) for k in self.SLICE_ORDER])
        self.L.oracle_synth_tuple_slices(C.c_uint64(seed), C.c_double(scale), C.c_long(No), C.c_long(Nv),
                                         C.c_long(abc[0]), C.c_long(abc[1]), C.c_long(abc[2]), arr)
        return S

    # ---- equations
    DOUBLES_ORDER = ["VAB", "VAC", "VBC", "VBA", "VCA", "VCB", "HA", "HB", "HC",
                     "TA", "TB", "TC", "TAB", "TAC", "TBC"]

    def doubles(self, No, Nv, S):
        out = np.empty(No ** 3)
        self.L.oracle_doubles(No, Nv, *[_d(S[k]) for k in self.DOUBLES_ORDER], _d(out))
        return out

    def singles(self, No, Nv, abc, Tai, S, Tijk):
        Z = Tijk.copy()
        self.L.oracle_singles(No, Nv, abc[0], abc[1], abc[2], _d(Tai), _d(S["VABij"]),
                              _d(S["VACij"]), _d(S["VBCij"]), _d(Z))
        return Z

    def energy_distinct(self, epsabc, No, epsi, Tijk, Zijk):
        return self.L.oracle_energy_distinct(epsabc, No, _d(epsi), _d(Tijk), _d(Zijk))

    def energy_same(self, epsabc, No, epsi, Tijk, Zijk):
        return self.L.oracle_energy_same(epsabc, No, _d(epsi), _d(Tijk), _d(Zijk))

    def tuple_energy(self, No, Nv, t, abc, want_cubes=False):
        T = np.empty(No ** 3) if want_cubes else None
        Z = np.empty(No ** 3) if want_cubes else None
        ct = C.c_double(0)
        e = self.L.oracle_tuple_energy(
            No, Nv, _d(t[EPS_I]), _d(t[EPS_A]), _d(t[TAI]), _d(t[TABIJ]), _d(t[VABIJ]),
            _d(t[VIJKA]), _d(t[VABCI]), _d(t.get(JIJKA)), _d(t.get(JABCI)),
            abc[0], abc[1], abc[2], _d(T), _d(Z), C.cast(C.byref(ct), _dp))
        return (e, ct.value, T, Z) if want_cubes else (e, ct.value)

    # ---- tuples
    def n_tuples(self, Nv):
        return self.L.oracle_n_tuples(Nv)

    def all_tuples(self, Nv):
        n = self.n_tuples(Nv)
        out = np.empty((n, 3), dtype=np.uint64)
        self.L.oracle_all_tuples(Nv, out.ctypes.data_as(_up), n)
        return out

    def group_and_sort(self, n_nodes, node_id, Nv):
        cap = self.n_tuples(Nv)
        out = np.empty((cap, 3), dtype=np.uint64)
        n = self.L.oracle_group_and_sort(n_nodes, node_id, Nv, out.ctypes.data_as(_up), cap)
        return out[:n].copy()

    def run(self, No, Nv, t, tuples=None):
        e, ct = C.c_double(0), C.c_double(0)
        tp, n = (None, 0)
        if tuples is not None:
            tuples = np.ascontiguousarray(tuples, dtype=np.uint64)
            tp, n = tuples.ctypes.data_as(_up), len(tuples)
        self.L.oracle_run(No, Nv, _d(t[EPS_I]), _d(t[EPS_A]), _d(t[TAI]), _d(t[TABIJ]),
                          _d(t[VABIJ]), _d(t[VIJKA]), _d(t[VABCI]), _d(t.get(JIJKA)),
                          _d(t.get(JABCI)), tp, n, C.cast(C.byref(e), _dp), C.cast(C.byref(ct), _dp))
        return e.value, ct.value


class Reference:
    """the reference's own implementation (oracle/_ref), if it was built"""

    @staticmethod
    def path(loops=False):
        return os.path.join(HERE, "_ref", "libatrip_ref_loops.so" if loops else "libatrip_ref.so")

    @classmethod
    def available(cls, loops=False):
        return os.path.exists(cls.path(loops))

    def __init__(self, loops=False):
        L = self.L = C.CDLL(self.path(loops))
        L.ref_run.restype = C.c_int
        L.ref_run.argtypes = [C.c_int, C.c_int] + [_dp] * 9 + [C.c_long, _dp, _dp, C.c_char_p, C.c_int]
        L.ref_chrono.restype = C.c_double
        L.ref_chrono.argtypes = [C.c_char_p]
        L.ref_doubles.argtypes = [C.c_long, C.c_long] + [_dp] * 18
        L.ref_singles.argtypes = [C.c_long] * 5 + [_dp] * 5
        for f in (L.ref_energy_distinct, L.ref_energy_same):
            f.restype = C.c_double
            f.argtypes = [C.c_double, C.c_long, _dp, _dp, _dp]
        self.has_complex = hasattr(L, "ref_run_z")
        if self.has_complex:
            L.ref_run_z.restype = C.c_int
            L.ref_run_z.argtypes = L.ref_run.argtypes
            L.ref_doubles_z.argtypes = [C.c_long, C.c_long] + [_dp] * 18
            L.ref_singles_z.argtypes = [C.c_long] * 5 + [_dp] * 5
            for f in (L.ref_energy_distinct_z, L.ref_energy_same_z):
                f.restype = C.c_double
                f.argtypes = [C.c_double, C.c_long, _dp, _dp, _dp]
        L.ref_group_and_sort.restype = C.c_long
        L.ref_group_and_sort.argtypes = [C.c_long, C.c_long, C.c_long, _up, C.c_long]
        L.ref_all_tuples.restype = C.c_long
        L.ref_all_tuples.argtypes = [C.c_long, _up, C.c_long]

    def run(self, No, Nv, t, max_iterations=0):
        e, ct = C.c_double(0), C.c_double(0)
        err = C.create_string_buffer(512)
        rc = self.L.ref_run(No, Nv, _d(t[EPS_I]), _d(t[EPS_A]), _d(t[TAI]), _d(t[TABIJ]),
                            _d(t[VABIJ]), _d(t[VIJKA]), _d(t[VABCI]), _d(t.get(JIJKA)),
                            _d(t.get(JABCI)), max_iterations, C.cast(C.byref(e), _dp),
                            C.cast(C.byref(ct), _dp), err, 512)
        if rc:
            raise RuntimeError("reference threw: " + err.value.decode())
        return e.value, ct.value

    def set_ijkabc(self, on):
        """Input::ijkabc of the following run / run_z calls (needs an oracle/_ref built from this tree)"""
        self.L.ref_set_ijkabc(int(bool(on)))

    def chrono(self, name):
        return self.L.ref_chrono(name.encode())

    # ---- F = Complex (complex128 arrays)
    def run_z(self, No, Nv, t, max_iterations=0):
        e, ct = C.c_double(0), C.c_double(0)
        err = C.create_string_buffer(512)
        rc = self.L.ref_run_z(No, Nv, _d(t[EPS_I]), _d(t[EPS_A]), _d(t[TAI]), _d(t[TABIJ]),
                              _d(t[VABIJ]), _d(t[VIJKA]), _d(t[VABCI]), _d(t.get(JIJKA)),
                              _d(t.get(JABCI)), max_iterations, C.cast(C.byref(e), _dp),
                              C.cast(C.byref(ct), _dp), err, 512)
        if rc:
            raise RuntimeError("reference threw: " + err.value.decode())
        return e.value, ct.value

    def doubles_z(self, No, Nv, S):
        out = np.empty(No ** 3, dtype=np.complex128)
        tb, vh = np.empty(No ** 3, dtype=np.complex128), np.empty(No ** 3, dtype=np.complex128)
        self.L.ref_doubles_z(No, Nv, *[_d(S[k]) for k in Oracle.DOUBLES_ORDER], _d(out), _d(tb), _d(vh))
        return out

    def singles_z(self, No, Nv, abc, Tai, S, Tijk):
        Z = Tijk.copy()
        self.L.ref_singles_z(No, Nv, abc[0], abc[1], abc[2], _d(Tai), _d(S["VABij"]),
                             _d(S["VACij"]), _d(S["VBCij"]), _d(Z))
        return Z

    def energy_distinct_z(self, epsabc, No, epsi, Tijk, Zijk):
        return self.L.ref_energy_distinct_z(epsabc, No, _d(epsi), _d(Tijk), _d(Zijk))

    def energy_same_z(self, epsabc, No, epsi, Tijk, Zijk):
        return self.L.ref_energy_same_z(epsabc, No, _d(epsi), _d(Tijk), _d(Zijk))

    def doubles(self, No, Nv, S, scratch=None):
        out = np.empty(No ** 3)
        tb, vh = scratch if scratch is not None else (np.empty(No ** 3), np.empty(No ** 3))
        self.L.ref_doubles(No, Nv, *[_d(S[k]) for k in Oracle.DOUBLES_ORDER], _d(out), _d(tb), _d(vh))
        return out

    def singles(self, No, Nv, abc, Tai, S, Tijk):
        Z = Tijk.copy()
        self.L.ref_singles(No, Nv, abc[0], abc[1], abc[2], _d(Tai), _d(S["VABij"]),
                           _d(S["VACij"]), _d(S["VBCij"]), _d(Z))
        return Z

    def energy_distinct(self, epsabc, No, epsi, Tijk, Zijk):
        return self.L.ref_energy_distinct(epsabc, No, _d(epsi), _d(Tijk), _d(Zijk))

    def energy_same(self, epsabc, No, epsi, Tijk, Zijk):
        return self.L.ref_energy_same(epsabc, No, _d(epsi), _d(Tijk), _d(Zijk))

    def group_and_sort(self, n_nodes, node_id, Nv):
        cap = Nv * (Nv + 1) * (Nv + 2) // 6 - Nv
        out = np.empty((cap, 3), dtype=np.uint64)
        n = self.L.ref_group_and_sort(n_nodes, node_id, Nv, out.ctypes.data_as(_up), cap)
        return out[:n].copy()

    def all_tuples(self, Nv):
        cap = Nv * (Nv + 1) * (Nv + 2) // 6 - Nv
        out = np.empty((cap, 3), dtype=np.uint64)
        self.L.ref_all_tuples(Nv, out.ctypes.data_as(_up), cap)
        return out
