"""CPU tests: the oracle (oracle/cc_oracle.c) against the reference's known answers.

Pins: (1) SURVEY 8c known-answer vectors measured from the compiled reference,
(2) golden fixtures produced by the real reference classes (tests/golden/),
(3) an independent multiplicity oracle, (4) when oracle/_ref/ref_driver is present and the
host has AVX-512, the real reference run live.
"""
import numpy as np
import pytest

import golden_util as G
import oracle_lib as O

# (J, cf, lhs, rhs) -> (n_tuples, digest, colsums or None); SURVEY 8c, kBlockSize = 256
KNOWN = [
    ((1, 1, 100000, 1000), (99905, 10349564124114550294, [49892540, 0, 49892540])),
    ((2, 1, 200000, 20000), (199981, 6558397397398525477, [1999961664, 1993900611, 0, 1999961664, 0, 1993900611])),
    ((2, 2, 10000, 1000), (9952, 17583432778385065612, [4908616, 4845072, 0, 4908616, 0, 4845072])),
    ((3, 5, 300000, 20000), (285000, 2819346475311658104,
                             [2832133750, 2761510000, 2888024375, 0, 2832133750, 0, 2761510000, 0, 2888024375])),
    ((4, 8, 300000, 20000), (270336, 10954991527034855424, None)),
    ((4, 1, 300000, 20000), (299934, 3961181216388177270, None)),
    ((4, 5, 1000000, 100000), (1003125, 18084939429413948304,
                               [49208715625, 49143915625, 50306137500, 49458043750, 0, 49208715625, 0, 49143915625, 0,
                                50306137500, 0, 49458043750])),
]


def test_hash_known_values():
    # hash_functions.h:8-16 evaluated by hand-written python big-int arithmetic
    def ref(x):
        m = (1 << 64) - 1
        x ^= x >> 32
        x = (x * 0xD6E8FEB86659FD93) & m
        x ^= x >> 32
        x = (x * 0xD6E8FEB86659FD93) & m
        x ^= x >> 32
        return x

    xs = np.array([0, 1, 2, 0xFFFFFFFF, 0x100000000, (1 << 63), (1 << 64) - 1, 0x9E3779B97F4A7C15, 12345678901234567], dtype=np.uint64)
    got = O.murmurhash64(xs)
    for x, g in zip(xs, got):
        assert int(g) == ref(int(x))


def test_build_keys_generator():
    # chaining_ht.cpp:15-26: integer division step (cf=3, n=20 -> num_unique 7, step 2)
    k = O.build_keys(20, 3)
    assert k.tolist() == [0, 0, 0, 2, 2, 2, 4, 4, 4, 6, 6, 6, 8, 8, 8, 10, 10, 10, 12, 12]
    assert O.build_keys(8, 1).tolist() == list(range(8))
    assert O.build_keys(0, 1).size == 0


@pytest.mark.parametrize("cfg,expect", KNOWN, ids=[str(k[0]) for k in KNOWN])
def test_known_answer_vectors(cfg, expect):
    J, cf, lhs_n, rhs = cfg
    n, digest, colsum = expect
    lhs = O.gen_lhs_main(lhs_n, J, rhs)
    bk = O.build_keys(rhs, cf)
    tabs = [O.OracleChain(bk) for _ in range(J)]
    d = O.pipeline(tabs, lhs, 256)
    assert d["n_tuples"] == n and d["digest"] == digest
    if colsum:
        assert d["colsum"] == colsum
    # compaction must be result-transparent: full and threshold compaction
    for kw in (dict(compaction=1), dict(compaction=2, threshold=64), dict(use_inone=True)):
        d2 = O.pipeline(tabs, lhs, 256, **kw)
        assert (d2["n_tuples"], d2["digest"], d2["colsum"]) == (d["n_tuples"], d["digest"], d["colsum"])
    # LP tables give the same multiset; chunk size does not matter
    d3 = O.pipeline([O.OracleLP(bk) for _ in range(J)], lhs, 2048)
    assert (d3["n_tuples"], d3["digest"]) == (n, digest)
    # independent multiplicity oracle
    m = O.multiplicity_oracle([bk] * J, lhs)
    assert (m["n_tuples"], m["digest"], m["colsum"], m["probe_tuples"]) == (n, digest, d["colsum"], d["probe_tuples"])


def test_golden_pipeline_main():
    for g in G.load_index()["pipeline_main"]:
        lhs = O.gen_lhs_main(g["lhs"], g["J"], g["rhs"])
        bk = O.build_keys(g["rhs"], g["cf"])
        tabs = [O.OracleChain(bk) for _ in range(g["J"])]
        d = O.pipeline(tabs, lhs, g["block"], compaction=g["compact"])
        for key in ("n_tuples", "digest", "colsum", "probe_tuples", "next_calls", "level_in", "level_chunks"):
            assert d[key] == g[key], (g, key)


def test_golden_pipeline_explicit_tuples():
    for g in G.load_index()["pipeline_explicit"]:
        J = g["J"]
        lhs = G.load_i64(f"{g['name']}_lhs.bin").reshape(-1, J)
        bk = O.build_keys(g["rhs"], g["cf"])
        tabs = [(O.OracleLP if g["kind"] == 0 else O.OracleChain)(bk) for _ in range(J)]
        want = G.load_i64(f"{g['name']}_k{g['kind']}_tuples.bin").reshape(-1, 3 * J)
        for kw in (dict(), dict(compaction=1), dict(compaction=2, threshold=100)):
            d = O.pipeline(tabs, lhs, g["block"], collect=True, **kw)
            assert d["n_tuples"] == g["n_tuples"] and d["digest"] == g["digest"] and d["colsum"] == g["colsum"]
            assert np.array_equal(G.sort_rows(d["tuples"]), want)
        d = O.pipeline(tabs, lhs, g["block"])
        assert d["next_calls"] == g["next_calls"] and d["level_chunks"] == g["level_chunks"]


def test_golden_next_protocol():
    """Per-Next parity with the real ScanStructure / LPScanStructure."""
    for g in G.load_index()["nextdump"]:
        keys = G.load_i64(f"{g['name']}_keys.bin")
        want = G.load_nextdump(f"{g['name']}_k{g['kind']}_i{g['inone']}_next.bin")
        bk = O.build_keys(g["n"], g["cf"])
        table = (O.OracleLP if g["kind"] == 0 else O.OracleChain)(bk)
        B = g["block"]
        sel = np.arange(B, dtype=np.uint32)
        n_tuples = 0
        for ci, k0 in enumerate(range(0, keys.size, B)):
            blk = np.zeros(B, dtype=np.int64)
            fill = min(B, keys.size - k0)
            blk[:fill] = keys[k0:k0 + fill]
            got = O.scan_chunk(table, blk, sel, fill, B, inone=bool(g["inone"]))
            assert len(got) == len(want[ci]), (g, ci)
            for (gp, gv), (wp, wv) in zip(got, want[ci]):
                assert np.array_equal(gp, wp) and np.array_equal(gv, wv)
                n_tuples += gp.size
        assert n_tuples == g["n_tuples"]


@pytest.mark.parametrize("scale,hit,cf,expect", [(3, 2, 1, 67114250), (3, 1, 4, 134218336), (0, 2, 1, 67098719)])
def test_microbench_tuple_counts(scale, hit, cf, expect):
    """#tuples of simd_micro_bench, all variants: 2^27 glibc rand() keys."""
    block, n_rhs = 256 << scale, 128 << scale
    keys = O.gen_keys_rand(1 << 27, n_rhs * hit - 1)
    bk = O.build_keys(n_rhs, cf)
    n_lp, _ = O.microbench(O.OracleLP(bk), keys, block)
    assert n_lp == expect
    n_ch, _ = O.microbench(O.OracleChain(bk), keys[: 1 << 24], block, inone=True)
    n_lp2, _ = O.microbench(O.OracleLP(bk), keys[: 1 << 24], block, inone=True)
    assert n_ch == n_lp2


def test_bandit_matches_reference_golden():
    g = G.load_index()["bandit"]
    arms = [0, 32, 64, 128, 256, 384, 512, 768, 1024]
    b = O.OracleBandit(len(arms))
    lcg = 88172645463325252
    got = []
    for i in range(g["steps"]):
        a = b.select()
        thr = arms[a]
        lcg = (lcg * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
        noise = float((lcg >> 33) % 1000) / 1000.0
        best = 256.0 if i < g["steps"] // 2 else 768.0
        scale = 1.0 if i < g["steps"] // 2 else 4.0
        reward = scale * (2.0 - abs(float(thr) - best) / 1024.0) + 0.05 * noise
        b.update(a, reward)
        got.append(thr)
    assert got == g["arms"]


def test_edge_cases():
    # empty build side: one-slot table, nothing matches (chaining_ht.cpp:5-6 starts at 1 bucket)
    empty = np.empty(0, dtype=np.int64)
    lhs = np.arange(10, dtype=np.int64).reshape(-1, 1)
    for T in (O.OracleLP, O.OracleChain):
        d = O.pipeline([T(empty)], lhs, 256)
        assert d["n_tuples"] == 0 and d["next_calls"] == 0
    # ragged last chunk + all-duplicate build side
    bk = np.full(7, 5, dtype=np.int64)
    lhs = np.array([[5], [4], [5]], dtype=np.int64)
    for T in (O.OracleLP, O.OracleChain):
        d = O.pipeline([T(bk)], lhs, 2, collect=True)
        assert d["n_tuples"] == 14 and (d["tuples"] == 5).sum() == 28


@pytest.mark.skipif(not O.have_ref_driver(), reason="oracle/_ref/ref_driver not built or host lacks AVX-512")
def test_live_reference_agrees():
    r = O.ref_driver("main", 3, 4, 40000, 3000, 256, 0)
    lhs = O.gen_lhs_main(40000, 3, 3000)
    bk = O.build_keys(3000, 4)
    d = O.pipeline([O.OracleChain(bk) for _ in range(3)], lhs, 256)
    for key in ("n_tuples", "digest", "colsum", "next_calls", "level_chunks"):
        assert d[key] == r[key]
    r2 = O.ref_driver("main", 3, 4, 40000, 3000, 256, 1)
    assert (r2["n_tuples"], r2["digest"]) == (r["n_tuples"], r["digest"])


# ------------------------------------------------------------------ payload-keeping join
def test_payload_join_oracle_against_independent_join():
    """orc_join_payload (both table kinds) == a numpy sort-merge join; the reference payload is row + 10000000."""
    assert np.array_equal(O.ref_payload(5), np.arange(5) + 10000000)
    rng = np.random.Generator(np.random.PCG64(3))
    for n, cf in [(1000, 1), (4096, 4), (777, 20), (1, 1), (0, 1)]:
        bk = O.build_keys(n, cf)
        pay = [O.ref_payload(n), rng.integers(-(1 << 62), 1 << 62, size=n, dtype=np.int64)]
        keys = rng.integers(0, max(1, 2 * n), size=20000, dtype=np.int64)
        order = np.argsort(bk, kind="stable")
        sk = bk[order]
        lo, hi = np.searchsorted(sk, keys, "left"), np.searchsorted(sk, keys, "right")
        reps = hi - lo
        pi = np.repeat(np.arange(keys.size), reps)
        bi = order[np.concatenate([np.arange(a, b) for a, b in zip(lo[reps > 0], hi[reps > 0])])] if reps.sum() else np.empty(0, dtype=np.int64)
        want = O.sort_rows(np.stack([keys[pi], bk[bi], pay[0][bi], pay[1][bi]], axis=1))
        for kind in (0, 1):
            got = O.join_payload(kind, bk, pay, keys)
            assert np.array_equal(O.sort_rows(got), want)
            # key columns agree with the pinned key-only pipeline: same count, same key sums
            if n:
                tab = (O.OracleLP if kind == 0 else O.OracleChain)(bk)
                ref = O.pipeline([tab], keys.reshape(-1, 1), 2048)
                assert ref["n_tuples"] == got.shape[0]
                assert ref["colsum"][0] == int(got[:, 0].view(np.uint64).sum(dtype=np.uint64))
                assert ref["colsum"][2] == int(got[:, 1].view(np.uint64).sum(dtype=np.uint64))
