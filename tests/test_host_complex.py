"""CPU: the complex-field pieces of the product that are shared between host and device
(`__host__ __device__` code in atrip_b200/csrc/stores.cuh and reduction_z.cuh), checked against
the oracle without a GPU:
  * the store layout map (which tensor element, part and sign every AX / BY element holds) must
    make the REAL contraction the kernel runs, on K-doubled operands, equal the reference's complex
    Tijk (doubles_contribution<Complex>);
  * the per-point energy function of the complex reduction kernel must reproduce
    get_energy_distinct/same<Complex>.
No compute entry point is called (those need the GPU and fail without it)."""
import numpy as np
import pytest

from atrip_b200 import capi
from oracle.oracle import EPS_A, EPS_I, TABIJ, TAI, VABCI, VIJKA


def _stores(lib, No, Nv, t, abc):
    """AX variants 0/1 of a, b, c and the six BY slices of the tuple, built from the library's map"""
    Kh = No + Nv
    Kp = (2 * Kh + 15) // 16 * 16
    src = {1: t[TABIJ], 2: t[VIJKA], 3: t[VABCI]}

    def value(ref):
        tensor, part, sign, lin = ref
        if tensor == 0:
            return 0.0
        z = src[tensor][lin]
        return sign * (z.imag if part else z.real)

    AX = {}
    for x in set(abc):
        for var in (0, 1):
            m = np.zeros((No * No, Kp))
            for row in range(No * No):
                for kap in range(Kp):
                    m[row, kap] = value(capi.store_source(0, No, Nv, var, x, 0, row, kap))
            AX[(x, var)] = m
    a, b, c = abc
    BY = []
    for (y, z, tf) in [(b, c, 0), (a, c, 0), (c, b, 1), (a, b, 0), (c, a, 1), (b, a, 1)]:
        m = np.zeros((No, Kp))
        for r in range(No):
            for kap in range(Kp):
                m[r, kap] = value(capi.store_source(1, No, Nv, tf, y, z, r, kap))
        BY.append(m)
    return AX, BY


def _contract(No, AX, BY, abc, var):
    """the three class GEMMs of contraction.cuh (real arithmetic), assembled into Tijk[i,j,k]"""
    a, b, c = abc

    def T(m):  # rows (p,q) -> (q,p): the transposed tensor map
        return m.reshape(No, No, -1).transpose(1, 0, 2).reshape(No * No, -1)

    A = {x: AX[(x, var)] for x in set(abc)}
    Ck = A[a] @ BY[0].T + T(A[b]) @ BY[1].T          # [i + j No, k]
    Cj = A[a] @ BY[2].T + T(A[c]) @ BY[3].T          # [i + k No, j]
    Ci = A[b] @ BY[4].T + T(A[c]) @ BY[5].T          # [j + k No, i]
    # rows are u + v No with u fastest: reshape(No(v), No(u), n)
    Ck = Ck.reshape(No, No, No)  # [j, i, k]
    Cj = Cj.reshape(No, No, No)  # [k, i, j]
    Ci = Ci.reshape(No, No, No)  # [k, j, i]
    W = np.zeros((No, No, No))
    W += Ck.transpose(1, 0, 2)   # [i, j, k]
    W += Cj.transpose(1, 2, 0)   # [i, j, k] from [k, i, j]
    W += Ci.transpose(2, 1, 0)   # [i, j, k] from [k, j, i]
    return W


@pytest.mark.parametrize("abc", [(0, 2, 4), (1, 1, 3), (2, 4, 4)])
def test_complex_store_layout_reproduces_reference_Tijk(lib, oracle, abc):
    No, Nv = 3, 5
    t = oracle.inputs_z(No, Nv, seed=11, scale=0.3)
    AX, BY = _stores(lib, No, Nv, t, abc)
    W = _contract(No, AX, BY, abc, 0) + 1j * _contract(No, AX, BY, abc, 1)
    _, _, T, _ = oracle.tuple_energy_z(No, Nv, t, abc, want_cubes=True)
    T = T.reshape((No, No, No), order="F")
    assert np.abs(W - T).max() <= 1e-13 * np.abs(T).max()


def test_complex_point_energy_matches_oracle(lib, oracle):
    No, Nv = 9, 12
    t = oracle.inputs_z(No, Nv, seed=5, scale=0.2)
    t[EPS_I] = t[EPS_I] + 0.01j * oracle.fill(5, 4, 1.0, No)  # exercise a complex denominator too
    for abc in [(0, 1, 2), (3, 3, 7), (2, 11, 11), (9, 10, 11)]:
        _, _, T, Z = oracle.tuple_energy_z(No, Nv, t, abc, want_cubes=True)
        epsabc = float((t[EPS_A][abc[0]] + t[EPS_A][abc[1]] + t[EPS_A][abc[2]]).real)
        same = (abc[0] == abc[1]) != (abc[1] == abc[2])
        want = (oracle.energy_same_z if same else oracle.energy_distinct_z)(epsabc, No, t[EPS_I], T, Z)
        got = capi.host_energy_z(No, epsabc, t[EPS_I], T, Z, same)
        assert abs(got - want) <= 1e-12 * abs(want), (abc, got, want)
