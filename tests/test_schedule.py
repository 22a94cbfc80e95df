"""CPU: slice ownership and the per-batch fetch schedule (atrip_b200/csrc/schedule.hpp) through
the host-only C-ABI.  The single-process tests simulate every rank; tests/test_multirank_gloo.py
runs the same protocol over a real 2-process gloo group."""
import numpy as np
import pytest

from atrip_b200 import capi

KA, KB, KV = 0, 1, 2


def stores_of(Nv, n):
    """content[rank][store][slot] = global id of the slice held there (built from local_slot)"""
    out = []
    for r in range(n):
        sizes = capi.shard_sizes(Nv, r, n)
        content = [np.full(sizes[k], -1, dtype=np.int64) for k in range(3)]
        for x in range(Nv):
            s = capi.local_slot(capi.TA, x, 0, Nv, r, n)
            assert s == capi.local_slot(capi.VIJKA, x, 0, Nv, r, n)
            if s >= 0:
                assert content[KA][s] == -1
                content[KA][s] = x
            s = capi.local_slot(capi.VABCI_T, x, x, Nv, r, n)
            if s >= 0:
                assert content[KB][s] == -1
                content[KB][s] = Nv * Nv + x
            for y in range(Nv):
                s = capi.local_slot(capi.VABCI, x, y, Nv, r, n)
                if s >= 0:
                    assert content[KB][s] == -1
                    content[KB][s] = x + y * Nv
                if x <= y:
                    s = capi.local_slot(capi.VABIJ, x, y, Nv, r, n)
                    if s >= 0:
                        assert content[KV][s] == -1
                        content[KV][s] = x + y * Nv
        out.append(content)
    return out


def wanted_ids(abc, Nv):
    """global slice ids a tuple needs, in TupleRec order (ax[3], by[6], vij[3])"""
    a, b, c = (int(v) for v in abc)

    def bid(y, z, t):
        return Nv * Nv + y if (y == z and t) else y + z * Nv
    return ([a, b, c],
            [bid(b, c, 0), bid(a, c, 0), bid(c, b, 1), bid(a, b, 0), bid(c, a, 1), bid(b, a, 1)],
            [b + c * Nv, a + c * Nv, a + b * Nv])


@pytest.mark.parametrize("Nv,n", [(16, 2), (13, 3), (24, 4), (17, 8), (9, 1)])
def test_every_slice_has_one_home_and_slots_are_dense(lib, Nv, n):
    stores = stores_of(Nv, n)
    for k in range(3):
        for r in range(n):
            assert np.all(stores[r][k] >= 0), "holes in the slot numbering"
    # A and B slices: exactly one holder, the owner, at the slot the owner formula gives
    for x in range(Nv):
        o, s = capi.slice_slot(capi.TA, x, 0, Nv, n)
        assert o == x % n and stores[o][KA][s] == x
        assert sum(int(np.any(stores[r][KA] == x)) for r in range(n)) == 1
        for y in range(Nv):
            o, s = capi.slice_slot(capi.VABCI, x, y, Nv, n)
            assert o == x % n and stores[o][KB][s] == x + y * Nv
            if x <= y:  # V: owner(x) always, owner(y) as a replica
                o, s = capi.slice_slot(capi.VABIJ, x, y, Nv, n)
                assert o == x % n and stores[o][KV][s] == x + y * Nv
                holders = {r for r in range(n) if np.any(stores[r][KV] == x + y * Nv)}
                assert holders == {x % n, y % n}


@pytest.mark.parametrize("Nv,n,batch", [(16, 2, 7), (13, 3, 5), (24, 4, 16), (17, 8, 3), (12, 1, 10)])
def test_fetch_schedule_delivers_every_slice(lib, Nv, n, batch):
    stores = stores_of(Nv, n)
    total_ranges = total_slices = 0
    for r in range(n):
        tl = capi.host_tuples(capi.GROUP_AND_SORT, Nv, r, n, pad=True)
        need = capi.cache_need(Nv, r, n, tl, batch)
        owned = capi.shard_sizes(Nv, r, n)
        for k0 in range(0, len(tl), batch):
            region = (k0 // batch) % 2
            base = [owned[k] + region * need[k] for k in range(3)]
            recs, ranges = capi.plan_batch(Nv, r, n, tl[k0:k0 + batch], base)
            # the "exchange": copy what each owner holds into this rank's cache region
            cache = [dict() for _ in range(3)]
            for peer, kind, src, cnt, dst in ranges.tolist():
                assert peer != r and cnt > 0 and src + cnt <= len(stores[peer][kind])
                for i in range(cnt):
                    assert dst + i < need[kind], "cache_need underestimates"
                    assert dst + i not in cache[kind], "two ranges write the same cache slot"
                    cache[kind][dst + i] = stores[peer][kind][src + i]
                total_ranges += 1
                total_slices += cnt
            for t, rec in zip(tl[k0:k0 + batch], recs):
                if not t.any():
                    assert rec[3] == 1
                    continue
                assert rec[3] == 0 and tuple(rec[:3]) == tuple(int(v) for v in t)
                got = (rec[4:7], rec[7:13], rec[13:16])
                for kind, want in enumerate(wanted_ids(t, Nv)):
                    for slot, wid in zip(got[kind], want):
                        if slot < owned[kind]:
                            assert stores[r][kind][slot] == wid
                        else:
                            assert cache[kind][slot - base[kind]] == wid
    if n == 1:
        assert total_ranges == 0
    else:
        assert total_ranges > 0 and total_slices >= total_ranges


def test_ranges_merge_along_runs(lib):
    """group-and-sort runs (p0, p1, home z stepping by n) ask for consecutive slots of the owner:
    the B requests of a batch collapse to far fewer messages than slices"""
    Nv, n, r, batch = 64, 4, 1, 32
    tl = capi.host_tuples(capi.GROUP_AND_SORT, Nv, r, n, pad=False)
    owned = capi.shard_sizes(Nv, r, n)
    k0 = len(tl) // 2
    recs, ranges = capi.plan_batch(Nv, r, n, tl[k0:k0 + batch], owned)
    b = ranges[ranges[:, 1] == KB]
    assert b[:, 3].sum() >= 3 * len(b), (len(b), b[:, 3].sum())


def test_remote_fetches_per_tuple_match_survey(lib):
    """with GPU-as-node group-and-sort about 2 ABPH pair slices and ~20/Nv
    single-index slices per tuple are remote (8 owners)"""
    Nv, n, r = 64, 8, 3
    tl = capi.host_tuples(capi.GROUP_AND_SORT, Nv, r, n, pad=False)
    owned = capi.shard_sizes(Nv, r, n)
    tot = np.zeros(3)
    batch = 64
    for k0 in range(0, len(tl), batch):
        _, ranges = capi.plan_batch(Nv, r, n, tl[k0:k0 + batch], owned)
        for kind in range(3):
            tot[kind] += ranges[ranges[:, 1] == kind][:, 3].sum()
    per_tuple = tot / len(tl)
    assert per_tuple[KB] < 2.6 and per_tuple[KA] < 0.5 and per_tuple[KV] < 0.3, per_tuple


def _caps(need):
    return [max(3 * need[KA], 3), max(3 * need[KB], 6), max(3 * need[KV], 3)]


@pytest.mark.parametrize("Nv,n,batch,calls", [(16, 2, 7, 1), (13, 3, 5, 4), (24, 4, 16, 3), (17, 8, 3, 5), (40, 8, 24, 2),
                                              (12, 1, 10, 2)])
def test_persistent_cache_schedule_is_consistent(lib, Nv, n, batch, calls):
    """the engine's real schedule (atrip_b200_run: one fetch cache shared by all batches and calls, slots
    re-assigned in ring order two batches after their last use, prefetch for the next call) checked
    against an independent model of the cache contents, for every rank"""
    for r in range(n):
        tl = capi.host_tuples(capi.GROUP_AND_SORT, Nv, r, n, pad=True)
        need = capi.cache_need(Nv, r, n, tl, batch)
        st = capi.check_schedule(Nv, r, n, tl, batch, _caps(need), calls=calls)
        assert st["batches"] == sum(-(-len(c) // batch) for c in np.array_split(tl, range(-(-len(tl) // calls), len(tl), -(-len(tl) // calls))))
        if n == 1:
            assert st["fetched"] == [0, 0, 0] and st["hits"] == [0, 0, 0]


def test_cache_reuse_cuts_the_single_index_traffic(lib):
    """the two slowly varying indices of a group-and-sort run keep their A slices across batches
    (the reference's Recycled / exact-match reuse, SliceUnion.cxx:66-137): with the persistent cache
    an A slice is fetched far less often than once per batch"""
    Nv, n, r, batch = 64, 8, 3, 8
    tl = capi.host_tuples(capi.GROUP_AND_SORT, Nv, r, n, pad=False)
    owned = capi.shard_sizes(Nv, r, n)
    per_batch = 0
    for k0 in range(0, len(tl), batch):
        _, ranges = capi.plan_batch(Nv, r, n, tl[k0:k0 + batch], owned)
        per_batch += int(ranges[ranges[:, 1] == KA][:, 3].sum())
    need = capi.cache_need(Nv, r, n, tl, batch)
    st = capi.check_schedule(Nv, r, n, tl, batch, _caps(need))
    assert st["fetched"][KA] < 0.6 * per_batch, (st, per_batch)
    assert st["hits"][KA] > 0


def test_too_small_a_cache_is_reported(lib):
    Nv, n, r, batch = 24, 4, 1, 16
    tl = capi.host_tuples(capi.GROUP_AND_SORT, Nv, r, n, pad=True)
    with pytest.raises(capi.EngineError, match="overflow"):
        capi.check_schedule(Nv, r, n, tl, batch, [1, 2, 1])


def test_owned_slice_lists_cover_every_slice_once_per_holder(lib):
    """atrip_b200_host_owned_slices (what a rank has to upload, SliceUnion.cxx:305-332) against the
    ownership map: single-index and ordered-pair slices have one holder, x <= y pair slices are held by
    the owners of both indices"""
    Nv, n = 13, 3
    seen = {k: {} for k in (capi.TA, capi.VIJKA, capi.VABCI, capi.TABIJ, capi.VABIJ)}
    for r in range(n):
        for kind in seen:
            for x, y in capi.owned_slices(kind, Nv, r, n).tolist():
                seen[kind].setdefault((x, y), set()).add(r)
    for x in range(Nv):
        assert seen[capi.TA][(x, 0)] == {x % n} and seen[capi.VIJKA][(x, 0)] == {x % n}
        for y in range(Nv):
            assert seen[capi.VABCI][(x, y)] == {x % n}
            if x <= y:
                assert seen[capi.TABIJ][(x, y)] == {x % n, y % n}
                assert seen[capi.VABIJ][(x, y)] == {x % n, y % n}
    assert len(seen[capi.VABCI]) == Nv * Nv and len(seen[capi.TABIJ]) == Nv * (Nv + 1) // 2


def test_tuples_distribution_dry_run_tool(lib):
    """tools/tuples_distribution.py (the host-only analog of the reference's bench/tuples-distribution.cxx): whole
    lists of every rank through the persistent-cache replay checker; the fetch counts are consistent"""
    import os
    import subprocess
    import sys
    from conftest import ROOT
    out = subprocess.check_output([sys.executable, os.path.join(ROOT, "tools", "tuples_distribution.py"), "--no", "8",
                                   "--nv", "24", "--ranks", "3", "--batch", "16", "--calls", "3"], text=True)
    rows = [l.split() for l in out.splitlines() if l and l[0] == " " and l.split()[0].isdigit()]
    assert len(rows) == 3, out
    total = 0
    for r in rows:
        tuples, fakes = int(r[1]), int(r[2])
        total += tuples - fakes
    assert total == 24 * 25 * 26 // 6 - 24, out
