"""GPU parity tests of the payload-carrying join: tables that keep the payload column the reference
generates and drops (chaining_ht.cpp:21-23,34), probed through cc_probe_batch_payload, against the CPU oracle
(oracle/cc_oracle.c orc_join_payload).  Integer work => bit-exact; result rows are compared as sorted multisets."""
import numpy as np
import pytest
import torch

import oracle_lib as O

pytestmark = pytest.mark.gpu

M64 = (1 << 64) - 1


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def u64sum(a) -> int:
    return int(np.asarray(a, dtype=np.int64).view(np.uint64).sum(dtype=np.uint64))


def rows_of(r) -> np.ndarray:
    m = r["n_matches"]
    cols = [r["out_key"], r["out_build_key"]] + list(r["out_cols"])
    return np.stack([c[:m].cpu().numpy() for c in cols], axis=1) if m else np.empty((0, len(cols)), dtype=np.int64)


def make_payload(n, ncols, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    cols = [O.ref_payload(n)]  # column 0: the reference's own payload (row i -> i + 10000000)
    for _ in range(ncols - 1):
        cols.append(rng.integers(-(1 << 62), 1 << 62, size=n, dtype=np.int64))
    return cols


@pytest.fixture(params=["direct", "partitioned", "partitioned2"])
def strategy(request, ccb):
    if request.param == "direct":
        ccb.set_probe_strategy(1)
    else:
        ccb.set_probe_strategy(2 if request.param == "partitioned" else 3, 16 << 10)
    yield request.param
    ccb.set_probe_strategy(0, 32 << 20)


# ------------------------------------------------------------------ table layout
@pytest.mark.parametrize("n,cf,ncols", [(1000, 1, 1), (4096, 4, 2), (20000, 5, 4), (1, 1, 1), (300, 20, 3)])
def test_lp_payload_sits_at_the_slot_of_its_key(ccb, n, cf, ncols):
    bk = O.build_keys(n, cf)
    cols = make_payload(n, ncols, n + cf)
    t = ccb.LPHashTable(keys=bk, payload=cols)
    assert t.payload_cols() == ncols
    slots = t.export()
    pay = t.export_payload()
    occ = slots != -1
    assert int(occ.sum()) == n
    assert np.array_equal(slots, O.OracleLP(bk).slots())  # attaching payloads never moves a key
    got = np.stack([slots[occ]] + [p[occ] for p in pay], axis=1)
    want = np.stack([bk] + cols, axis=1)
    assert np.array_equal(O.sort_rows(got), O.sort_rows(want))  # every (key, payload...) build row sits in exactly one slot
    for p in pay:
        assert not p[~occ].any()  # rows of empty slots are zero


@pytest.mark.parametrize("n,cf,ncols", [(1000, 1, 1), (4096, 4, 2), (20000, 5, 4), (1, 1, 1), (300, 20, 3)])
def test_chain_payload_follows_chain_order(ccb, n, cf, ncols):
    bk = O.build_keys(n, cf)
    cols = make_payload(n, ncols, n + cf)
    t = ccb.HashTable(keys=bk, payload=cols)
    begin, count, ckeys = t.export()
    pay = t.export_payload()
    # chain entries are in insertion (FIFO) order, so column 0 (row id + 10000000) names the build row of every entry
    rows = pay[0] - 10000000
    assert np.array_equal(np.sort(rows), np.arange(n))
    assert np.array_equal(bk[rows], ckeys)
    for c in range(ncols):
        assert np.array_equal(pay[c], cols[c][rows])
    nz = count > 0
    first = begin[nz].astype(np.int64)
    for b0, c0 in list(zip(first, count[nz]))[:2000]:
        assert np.all(np.diff(rows[b0:b0 + c0]) > 0)  # FIFO within a bucket (chaining_ht.cpp:34 push_back)


def test_reference_builder_keeps_the_dropped_payload(ccb):
    for T, kind in ((ccb.LPHashTable, 0), (ccb.HashTable, 1)):
        n, cf = 5000, 4
        t = T(n, cf, keep_payload=True)
        assert t.payload_cols() == 1
        bk = O.build_keys(n, cf)
        keys = O.gen_keys_counter(30000, 2, 8191)
        r = t.probe_batch_payload(dev(keys), capacity=keys.size * cf)
        want = O.join_payload(kind, bk, [O.ref_payload(n)], keys)
        assert r["n_matches"] == want.shape[0] and r["overflow"] == 0
        assert np.array_equal(O.sort_rows(rows_of(r)), O.sort_rows(want))


# ------------------------------------------------------------------ probe parity
@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("n,cf,hit,nprobe,ncols", [(1024, 1, 1, 100000, 1), (1024, 1, 2, 100000, 2), (2000, 4, 1, 50000, 3),
                                                   (5000, 8, 4, 77777, 4), (200000, 5, 1, 300000, 2), (128, 1, 1, 1, 1),
                                                   (64, 2, 1, 0, 1), (3000, 20, 1, 20000, 1)])
def test_probe_batch_payload_matches_oracle(ccb, strategy, kind, n, cf, hit, nprobe, ncols):
    rng = np.random.Generator(np.random.PCG64(n + cf + hit + nprobe))
    keys = rng.integers(0, max(1, n * hit), size=nprobe, dtype=np.int64)
    bk = O.build_keys(n, cf)
    cols = make_payload(n, ncols, 7 * n + cf)
    t = (ccb.LPHashTable if kind == 0 else ccb.HashTable)(keys=bk, payload=cols)
    assert t.info().has_duplicates == int(cf > 1 and n > 1)
    want = O.join_payload(kind, bk, cols, keys)
    r = t.probe_batch_payload(dev(keys), capacity=max(1, nprobe * cf))
    assert (r["n_matches"], r["overflow"]) == (want.shape[0], 0)
    assert r["key_sum"] == u64sum(want[:, 0]) and r["payload_sum"] == u64sum(want[:, 1])
    assert r["col_sum"] == [u64sum(want[:, 2 + c]) for c in range(ncols)]
    assert np.array_equal(O.sort_rows(rows_of(r)), O.sort_rows(want))


@pytest.mark.parametrize("kind", [0, 1])
def test_probe_batch_payload_arbitrary_keys_and_independent_join(ccb, strategy, kind):
    """random 64-bit keys with duplicates on both sides, checked against the oracle AND a numpy sort-merge join"""
    rng = np.random.Generator(np.random.PCG64(99 + kind))
    bk = rng.integers(-(1 << 62), 1 << 62, size=40000, dtype=np.int64)
    bk[bk == -1] = 5
    bk[1000:1500] = bk[:500]      # duplicate build keys
    bk[2000:2032] = bk[2000]      # one key 32 times
    pay = [rng.integers(-(1 << 62), 1 << 62, size=bk.size, dtype=np.int64), np.arange(bk.size, dtype=np.int64)]
    keys = np.concatenate([rng.choice(bk, size=60000), rng.integers(-(1 << 62), 1 << 62, size=20000, dtype=np.int64)])
    rng.shuffle(keys)
    t = (ccb.LPHashTable if kind == 0 else ccb.HashTable)(keys=bk, payload=pay)
    r = t.probe_batch_payload(dev(keys), capacity=4 * keys.size)
    got = O.sort_rows(rows_of(r))
    assert np.array_equal(got, O.sort_rows(O.join_payload(kind, bk, pay, keys)))
    # independent: sort-merge join in numpy
    order = np.argsort(bk, kind="stable")
    sk = bk[order]
    lo, hi = np.searchsorted(sk, keys, "left"), np.searchsorted(sk, keys, "right")
    reps = hi - lo
    probe_idx = np.repeat(np.arange(keys.size), reps)
    build_idx = order[np.concatenate([np.arange(a, b) for a, b in zip(lo[reps > 0], hi[reps > 0])])] if reps.sum() else np.empty(0, dtype=np.int64)
    indep = np.stack([keys[probe_idx], bk[build_idx], pay[0][build_idx], pay[1][build_idx]], axis=1)
    assert np.array_equal(got, O.sort_rows(indep))


@pytest.mark.parametrize("kind", [0, 1])
def test_payload_probe_options(ccb, kind):
    """row ids, a subset of the payload columns, count-only, and a too-small output"""
    n, cf = 3000, 3
    bk = O.build_keys(n, cf)
    cols = make_payload(n, 3, 1)
    keys = O.gen_keys_counter(40000, 2, 4095)
    t = (ccb.LPHashTable if kind == 0 else ccb.HashTable)(keys=bk, payload=cols)
    want = O.join_payload(kind, bk, cols, keys)
    sums = [u64sum(want[:, 2 + c]) for c in range(3)]
    r = t.probe_batch_payload(dev(keys), capacity=keys.size * cf, rowid=True, n_out_cols=2)
    m = r["n_matches"]
    assert m == want.shape[0] and len(r["out_cols"]) == 2 and r["col_sum"] == sums  # every column is summed, two are written
    rid = r["out_rowid"][:m].cpu().numpy()
    assert np.array_equal(keys[rid], r["out_key"][:m].cpu().numpy())
    assert np.array_equal(O.sort_rows(rows_of(r)), O.sort_rows(want[:, :4]))
    r = t.probe_batch_payload(dev(keys), materialize=False)
    assert (r["n_matches"], r["overflow"], r["col_sum"]) == (m, 0, sums)
    cap = m // 3
    r = t.probe_batch_payload(dev(keys), capacity=cap)
    assert r["n_matches"] == m and r["overflow"] == 1 and r["col_sum"] == sums
    part = np.stack([c[:cap].cpu().numpy() for c in [r["out_key"], r["out_build_key"]] + r["out_cols"]], axis=1)
    full = {tuple(x) for x in want.tolist()}
    assert all(tuple(x) in full for x in part.tolist())  # the rows that fit are genuine result rows
    # the key-only entry point is untouched by an attached payload
    r0 = t.probe_batch(dev(keys), capacity=keys.size * cf)
    assert r0["n_matches"] == m and r0["key_sum"] == u64sum(want[:, 0])


def test_payload_errors(ccb):
    bk = O.build_keys(1000, 1)
    t = ccb.LPHashTable(keys=bk)
    with pytest.raises(ccb.CCError):
        t.probe_batch_payload(dev(bk))  # no payload attached
    with pytest.raises(ccb.CCError):
        t.attach_payload(bk + 1, [O.ref_payload(1000)])  # not the keys the table was built from
    assert t.payload_cols() == 0
    with pytest.raises(ccb.CCError):
        t.attach_payload(bk, [O.ref_payload(1000)] * 5)  # more than CC_MAX_PAYLOAD_COLS
    t.attach_payload(bk, [O.ref_payload(1000)])
    t.attach_payload(bk, [O.ref_payload(1000) * 2, O.ref_payload(1000)])  # re-attach replaces
    assert t.payload_cols() == 2
    r = t.probe_batch_payload(dev(bk[:10]))
    assert sorted(r["out_cols"][0][:10].cpu().tolist()) == [2 * (10000000 + i) for i in range(10)]


def test_payload_probe_large_properties(ccb):
    """DRAM-resident sizes (LP: 2^26 slots of keys + as many of payload = 1 GiB, beyond L2 => the partitioned strategy
    picks itself): size-independent properties.  payload = f(key), so every result row must satisfy payload == f(probe key),
    and the column sums follow from the probe keys alone."""
    n = 1 << 24
    bk = torch.arange(n, dtype=torch.int64, device="cuda")
    pay = bk * 3 + 1
    for T in (ccb.LPHashTable, ccb.HashTable):
        tab = T(keys=bk, payload=[pay])
        for hit in (1, 2):
            keys = ccb.gen_keys_counter(1 << 26, 2, n * hit - 1)
            r = tab.probe_batch_payload(keys, capacity=keys.numel())
            hits = keys[keys < n]
            m = r["n_matches"]
            assert m == hits.numel() and r["overflow"] == 0
            s = int(hits.sum().item()) & M64
            assert r["key_sum"] == s and r["payload_sum"] == s
            assert r["col_sum"] == [(3 * int(hits.sum().item()) + m) & M64]
            assert torch.equal(r["out_cols"][0][:m], r["out_key"][:m] * 3 + 1)
            assert torch.equal(r["out_build_key"][:m], r["out_key"][:m])
            assert torch.equal(torch.sort(r["out_key"][:m]).values, torch.sort(hits).values)
        del tab
