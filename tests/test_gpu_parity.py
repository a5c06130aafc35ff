"""GPU parity tests (run with -m gpu on a B200): every CUDA path of the product, called
through the C ABI, against the CPU oracle and the committed golden fixtures.
Integer work => bit-exact comparisons everywhere."""
import os

import numpy as np
import pytest
import torch

import golden_util as G
import oracle_lib as O

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def u64sum(a) -> int:
    return int(np.asarray(a, dtype=np.int64).view(np.uint64).sum(dtype=np.uint64))


# ------------------------------------------------------------------ hash / generators
def test_hash_bit_exact(ccb):
    rng = np.random.Generator(np.random.PCG64(1))
    xs = np.concatenate([rng.integers(0, 1 << 63, size=100000, dtype=np.uint64) * 2 + rng.integers(0, 2, size=100000, dtype=np.uint64),
                         np.array([0, 1, 0xFFFFFFFF, 1 << 32, (1 << 64) - 1, 1 << 63], dtype=np.uint64)])
    got = ccb.murmurhash64(dev(xs.view(np.int64))).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, O.murmurhash64(xs))


@pytest.mark.parametrize("n,cf", [(20, 3), (1000, 1), (4096, 8), (2000000, 5), (7, 20), (0, 1)])
def test_build_key_generator(ccb, n, cf):
    assert np.array_equal(ccb.gen_build_keys(n, cf).cpu().numpy(), O.build_keys(n, cf))


def test_counter_generator(ccb):
    got = ccb.gen_keys_counter(100000, 2, (1 << 20) - 1, first=12345).cpu().numpy()
    assert np.array_equal(got, O.gen_keys_counter(100000, 2, (1 << 20) - 1, first=12345))


# ------------------------------------------------------------------ table builds
@pytest.mark.parametrize("n,cf", [(128, 1), (1024, 4), (1000, 8), (20000, 5), (300, 3), (1, 1), (0, 1), (200000, 20)])
def test_lp_build_equals_reference_layout(ccb, n, cf):
    """Ordered GPU build == serial reference insertion (linear_probing_ht.cpp:28-36), slot for slot."""
    t = ccb.LPHashTable(n, cf)
    want = O.OracleLP(O.build_keys(n, cf))
    assert t.info().n_slots == want.n_slots
    assert np.array_equal(t.export(), want.slots())
    assert t.info().has_duplicates == int(cf > 1 and n > 1)


@pytest.mark.parametrize("cf,scrambled", [(1, False), (4, False), (1, True)])
def test_lp_streaming_build_equals_reference_layout(ccb, cf, scrambled):
    """Tables beyond L2 are built from keys grouped by table slice first (the streaming build, tables.cu): the table must still
    equal the serial reference insertion slot for slot -- for the reference's own key column (with and without duplicates) and
    for keys in arbitrary order (== serial insertion in ascending unsigned key order)."""
    n = (1 << 22) + 12345  # 2^25 slots = 256 MiB: beyond the 96 MiB threshold of the streaming build
    keys = O.build_keys(n, cf)
    if scrambled:
        keys = (O.murmurhash64(keys.view(np.uint64)) >> np.uint64(2)).astype(np.int64)
    t = ccb.LPHashTable(keys=keys)
    order = np.argsort(keys.view(np.uint64), kind="stable")
    want = O.OracleLP(keys[order])
    assert t.info().n_slots == want.n_slots == 1 << 25
    assert np.array_equal(t.export(), want.slots())
    assert t.info().has_duplicates == int(cf > 1)


def test_lp_build_arbitrary_keys(ccb):
    rng = np.random.Generator(np.random.PCG64(5))
    keys = rng.integers(-(1 << 62), 1 << 62, size=50000, dtype=np.int64)
    keys[keys == -1] = 7
    keys[100:200] = keys[0]  # duplicates
    # ordered build == serial insertion in ascending unsigned key order
    order = np.argsort(keys.view(np.uint64), kind="stable")
    want = O.OracleLP(keys[order]).slots()
    t = ccb.LPHashTable(keys=keys)
    assert np.array_equal(t.export(), want)
    # unordered build: same occupied slots (insertion-order independent) and same key multiset
    t2 = ccb.LPHashTable(keys=keys, flags=ccb.CC_BUILD_UNORDERED)
    got = t2.export()
    assert np.array_equal(got == -1, want == -1)
    assert np.array_equal(np.sort(got), np.sort(want))
    # -1 is the empty sentinel: rejected (linear_probing_ht.cpp:7)
    with pytest.raises(ccb.CCError):
        ccb.LPHashTable(keys=np.array([3, -1, 4], dtype=np.int64))


@pytest.mark.parametrize("n,cf", [(128, 1), (1024, 4), (20000, 5), (1000, 8), (1, 1), (0, 1), (50000, 20)])
def test_chain_build_equals_reference_chains(ccb, n, cf):
    """Every bucket holds the same keys in the same (FIFO) order as the reference's std::list."""
    t = ccb.HashTable(n, cf)
    begin, count, ckeys = t.export()
    oc = O.OracleChain(O.build_keys(n, cf))
    key, nxt, head = oc.arrays()
    assert begin.size == oc.n_buckets
    assert int(count.sum()) == n
    for b in np.nonzero((count > 0) | (head != O.NIL))[0]:
        chain = []
        i = head[b]
        while i != O.NIL:
            chain.append(key[i])
            i = nxt[i]
        assert ckeys[begin[b]:begin[b] + count[b]].tolist() == chain
    assert t.info().max_chain == (int(count.max()) if n else 0)


def test_chain_build_arbitrary_keys_fifo(ccb):
    rng = np.random.Generator(np.random.PCG64(6))
    keys = rng.integers(-50, 3000, size=20000, dtype=np.int64)
    t = ccb.HashTable(keys=keys)
    begin, count, ckeys = t.export()
    oc = O.OracleChain(keys)
    key, nxt, head = oc.arrays()
    for b in np.nonzero(count > 0)[0]:
        chain = []
        i = head[b]
        while i != O.NIL:
            chain.append(key[i])
            i = nxt[i]
        assert ckeys[begin[b]:begin[b] + count[b]].tolist() == chain


# ------------------------------------------------------------------ chunk-granular protocol
def run_gpu_chunk(ccb, table, blk, sel, count, B, inone):
    join_key = ccb.Vector(data=dev(blk))
    inp = ccb.DataChunk(1, B)
    inp.data_[0] = join_key
    inp.selection_vector_ = dev(sel.view(np.int32))
    inp.count_ = count
    out = ccb.DataChunk(3, B)
    ss = table.Probe(join_key, count, inp.selection_vector_, B)
    calls = []
    while ss.HasNext():
        rc = (ss.InOneNext if inone else ss.Next)(join_key, inp, out)
        rsel = ccb.to_u32_numpy(out.selection_vector_)
        col2 = out.data_[2].data_.cpu().numpy()
        assert np.array_equal(rsel[rc:], np.arange(rc, B, dtype=np.uint32))  # Reset identity tail (base.h:96-99)
        assert out.data_[0].data_.data_ptr() == join_key.data_.data_ptr()  # Slice shares LHS storage (base.cpp:40)
        calls.append((rsel[:rc].copy(), col2[rsel[:rc]].copy()))
    return calls


def test_chunk_protocol_matches_reference_golden(ccb):
    """Probe / Next / InOneNext call by call against records dumped from the real reference classes."""
    for g in G.load_index()["nextdump"]:
        keys = G.load_i64(f"{g['name']}_keys.bin")
        want = G.load_nextdump(f"{g['name']}_k{g['kind']}_i{g['inone']}_next.bin")
        table = (ccb.LPHashTable if g["kind"] == 0 else ccb.HashTable)(g["n"], g["cf"])
        B = g["block"]
        sel = np.arange(B, dtype=np.uint32)
        total = 0
        for ci, k0 in enumerate(range(0, keys.size, B)):
            blk = np.zeros(B, dtype=np.int64)
            fill = min(B, keys.size - k0)
            blk[:fill] = keys[k0:k0 + fill]
            got = run_gpu_chunk(ccb, table, blk, sel, fill, B, bool(g["inone"]))
            assert len(got) == len(want[ci]), (g, ci)
            for (gp, gv), (wp, wv) in zip(got, want[ci]):
                assert np.array_equal(gp, wp) and np.array_equal(gv, wv)
                total += gp.size
        assert total == g["n_tuples"]


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("B,count", [(256, 256), (2048, 1500), (5000, 4999), (256, 0), (300, 1)])
def test_chunk_protocol_permuted_selection(ccb, kind, B, count):
    """Non-identity selection vectors, partial chunks, block sizes above one tile."""
    rng = np.random.Generator(np.random.PCG64(B * 7 + count + kind))
    n, cf = 700, 3
    bk = O.build_keys(n, cf)
    otab = (O.OracleLP if kind == 0 else O.OracleChain)(bk)
    gtab = (ccb.LPHashTable if kind == 0 else ccb.HashTable)(n, cf)
    blk = rng.integers(0, n + 100, size=B, dtype=np.int64)
    sel = rng.permutation(B).astype(np.uint32)
    for inone in (False, True):
        want = O.scan_chunk(otab, blk, sel, count, B, inone=inone)
        got = run_gpu_chunk(ccb, gtab, blk, sel, count, B, inone)
        assert len(got) == len(want)
        for (gp, gv), (wp, wv) in zip(got, want):
            assert np.array_equal(gp, wp) and np.array_equal(gv, wv)


def test_datachunk_primitives(ccb):
    """Append / Slice / Reset / FetchChunk / AppendChunk (base.cpp, data_collection.cpp)."""
    rng = np.random.Generator(np.random.PCG64(11))
    B = 512
    rows = rng.integers(0, 1000, size=(1300, 3), dtype=np.int64)
    table = ccb.DataCollection(3)
    table.AppendRows(rows)
    assert table.NumTuples() == 1300
    c = table.FetchChunk(512, 1024, B)
    assert c.count_ == 512 and np.array_equal(c.rows(), rows[512:1024])
    last = table.FetchChunk(1024, 1300, B)
    assert last.count_ == 276 and np.array_equal(last.rows(), rows[1024:1300])
    sv = rng.permutation(512)[:100].astype(np.uint32)
    s = ccb.DataChunk(5, B)
    s.Slice(c, dev(sv.view(np.int32)), 100)
    assert np.array_equal(s.rows()[:, :3], rows[512:1024][sv])
    d = ccb.DataChunk(5, B)
    d.Append(s, 60)
    d.Append(s, 40, 60)
    assert d.count_ == 100 and np.array_equal(d.rows()[:, :3], rows[512:1024][sv])
    sink = ccb.DataCollection(5)
    sink.AppendChunk(d)
    sink.AppendChunk(s)
    assert sink.NumTuples() == 200 and np.array_equal(sink.numpy()[:100, :3], rows[512:1024][sv])
    s.Reset()
    assert s.count_ == 0 and np.array_equal(ccb.to_u32_numpy(s.selection_vector_), np.arange(B, dtype=np.uint32))


# ------------------------------------------------------------------ facade pipeline (main.cpp protocol)
def facade_pipeline(ccb, tables, lhs, B, compaction=None, threshold=None):
    J = lhs.shape[1]
    src = ccb.DataCollection(J)
    src.AppendRows(lhs)
    sink = ccb.DataCollection(3 * J)
    inter = [ccb.DataChunk(J + 2 * (i + 1), B) for i in range(J)]
    comps = [ccb.Compactor(J + 2 * (i + 1), B, threshold) for i in range(J)] if compaction else None

    def execute(inp, level):  # main.cpp:119-170
        if level == J:
            sink.AppendChunk(inp)
            return
        join_key = inp.data_[level]
        ss = tables[level].Probe(join_key, inp.count_, inp.selection_vector_, B)
        while ss.HasNext():
            ss.Next(join_key, inp, inter[level])
            result = inter[level]
            if comps:
                result = comps[level].Compact(result)
                if result.count_ == 0:
                    continue
            execute(result, level + 1)

    for start in range(0, lhs.shape[0], B):
        execute(src.FetchChunk(start, min(start + B, lhs.shape[0]), B), 0)
    if comps:  # FlushPipelineCache, main.cpp:172-191
        for level in range(J):
            execute(comps[level].Flush(), level + 1)
    return sink.numpy()


@pytest.mark.parametrize("name", ["e1", "e3"])
def test_facade_pipeline_matches_reference_tuples(ccb, name):
    for g in [x for x in G.load_index()["pipeline_explicit"] if x["name"] == name]:
        J = g["J"]
        lhs = G.load_i64(f"{name}_lhs.bin").reshape(-1, J)
        want = G.load_i64(f"{name}_k{g['kind']}_tuples.bin").reshape(-1, 3 * J)
        T = ccb.LPHashTable if g["kind"] == 0 else ccb.HashTable
        tables = [T(g["rhs"], g["cf"]) for _ in range(J)]
        for kw in (dict(), dict(compaction=True), dict(compaction=True, threshold=64)):
            got = facade_pipeline(ccb, tables, lhs, g["block"], **kw)
            assert np.array_equal(G.sort_rows(got), want), (g, kw)


# ------------------------------------------------------------------ batch probe (fast path)
@pytest.fixture(params=["direct", "partitioned", "partitioned_two_pass"])
def strategy(request, ccb):
    """All probe strategies: direct, and key partitioning by table slice (forced, with tiny 16 KiB slices so that
    even the small test tables are split into many partitions) with the single-pass and the two-pass partition."""
    if request.param == "direct":
        ccb.set_probe_strategy(1)
    else:
        ccb.set_probe_strategy(2 if request.param == "partitioned" else 3, 16 << 10)
    yield request.param
    ccb.set_probe_strategy(0, 32 << 20)


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("n,cf,hit,nprobe", [(1024, 1, 1, 100000), (1024, 1, 2, 100000), (2000, 4, 1, 50000), (5000, 8, 4, 77777),
                                             (200000, 5, 1, 300000), (128, 1, 1, 1), (64, 2, 1, 0), (0, 1, 1, 1000)])
def test_probe_batch_matches_oracle(ccb, strategy, kind, n, cf, hit, nprobe):
    rng = np.random.Generator(np.random.PCG64(n + cf + hit + nprobe))
    keys = rng.integers(0, max(1, n * hit), size=nprobe, dtype=np.int64)
    bk = O.build_keys(n, cf)
    otab = (O.OracleLP if kind == 0 else O.OracleChain)(bk)
    gtab = (ccb.LPHashTable if kind == 0 else ccb.HashTable)(n, cf)
    want = O.pipeline([otab], keys.reshape(-1, 1), 2048, collect=True) if nprobe else dict(n_tuples=0, tuples=np.empty((0, 3), dtype=np.int64))
    cap = max(1, nprobe * max(cf, 1) * 2)
    r = gtab.probe_batch(dev(keys), capacity=cap, rowid=True)
    assert r["n_matches"] == want["n_tuples"] and r["overflow"] == 0
    m = r["n_matches"]
    got = np.stack([r["out_key"][:m].cpu().numpy(), np.zeros(m, dtype=np.int64), r["out_payload"][:m].cpu().numpy()], axis=1)
    assert np.array_equal(G.sort_rows(got), G.sort_rows(want["tuples"]))
    assert r["key_sum"] == u64sum(want["tuples"][:, 0]) and r["payload_sum"] == u64sum(want["tuples"][:, 2])
    rid = r["out_rowid"][:m].cpu().numpy()
    assert np.array_equal(keys[rid], got[:, 0])  # row ids point at the probe rows that produced the match
    # without row ids the partitioned strategy is eligible: same multiset
    r1 = gtab.probe_batch(dev(keys), capacity=cap)
    got1 = np.stack([r1["out_key"][:m].cpu().numpy(), np.zeros(m, dtype=np.int64), r1["out_payload"][:m].cpu().numpy()], axis=1)
    assert r1["n_matches"] == m and np.array_equal(G.sort_rows(got1), G.sort_rows(want["tuples"]))
    # count-only mode and the overflow report
    r2 = gtab.probe_batch(dev(keys), materialize=False)
    assert (r2["n_matches"], r2["key_sum"], r2["payload_sum"]) == (r["n_matches"], r["key_sum"], r["payload_sum"])
    if m > 1:
        r3 = gtab.probe_batch(dev(keys), capacity=m // 2)
        assert r3["overflow"] == 1 and r3["n_matches"] == m


def test_probe_batch_negative_and_extreme_keys(ccb, strategy):
    rng = np.random.Generator(np.random.PCG64(99))
    bk = rng.integers(-(1 << 62), 1 << 62, size=30000, dtype=np.int64)
    bk[bk == -1] = 5
    probe = np.concatenate([bk[::3], rng.integers(-(1 << 62), 1 << 62, size=10000, dtype=np.int64)])
    for kind in (0, 1):
        otab = (O.OracleLP if kind == 0 else O.OracleChain)(bk)
        gtab = (ccb.LPHashTable if kind == 0 else ccb.HashTable)(keys=bk)
        want = O.pipeline([otab], probe.reshape(-1, 1), 2048, collect=True)
        r = gtab.probe_batch(dev(probe), capacity=probe.size * 2)
        m = r["n_matches"]
        assert m == want["n_tuples"]
        got = np.stack([r["out_key"][:m].cpu().numpy(), np.zeros(m, dtype=np.int64), r["out_payload"][:m].cpu().numpy()], axis=1)
        assert np.array_equal(G.sort_rows(got), G.sort_rows(want["tuples"]))


def test_probe_batch_clustered_keys(ccb, strategy):
    """Adversarial LP table for the deferred tail walk of probe_unique_lp_kernel: 2000 unique build keys whose home slots all
    fall into 64 of the 8192 slots, i.e. ONE cluster of ~2000 consecutive occupied slots.  Nearly every probe key finds another
    key in its home slot (all 128 keys of a warp are pending at once: the per-warp ring overflows and the excess is walked on
    the spot), probe sequences run over hundreds of sectors (entries circulate through the ring for hundreds of iterations), and
    the key column ends long before the rings have drained (tile-less drain iterations).  Misses walk to the end of the cluster."""
    n = 2000
    n_slots = 8192
    cand = np.arange(1, 4_000_000, dtype=np.int64)
    home = O.murmurhash64(cand.view(np.uint64)) & np.uint64(n_slots - 1)
    cand = cand[home < 64]
    assert cand.size > 3 * n
    bk = cand[:n].copy()
    absent = cand[n:3 * n]
    rng = np.random.Generator(np.random.PCG64(5))
    for nprobe in (100, 5000, 200000):
        probe = np.concatenate([rng.choice(bk, size=nprobe), rng.choice(absent, size=nprobe // 2), rng.integers(0, 1 << 40, size=nprobe // 4, dtype=np.int64)])
        rng.shuffle(probe)
        gtab = ccb.LPHashTable(keys=bk)
        assert gtab.info().n_slots == n_slots and gtab.info().has_duplicates == 0
        want = O.pipeline([O.OracleLP(bk)], probe.reshape(-1, 1), 2048, collect=True)
        r = gtab.probe_batch(dev(probe), capacity=probe.size)
        m = r["n_matches"]
        assert m == want["n_tuples"] and r["overflow"] == 0
        assert r["key_sum"] == u64sum(want["tuples"][:, 0]) and r["payload_sum"] == u64sum(want["tuples"][:, 2])
        got = np.stack([r["out_key"][:m].cpu().numpy(), np.zeros(m, dtype=np.int64), r["out_payload"][:m].cpu().numpy()], axis=1)
        assert np.array_equal(G.sort_rows(got), G.sort_rows(want["tuples"]))
        r2 = gtab.probe_batch(dev(probe), materialize=False)
        assert (r2["n_matches"], r2["key_sum"]) == (m, r["key_sum"])


def test_deferred_tail_kernel_parity(ccb):
    """probe_unique_lp_kernel (the deferred tail walk, CCB_LEAN_DEFERRED_TAIL=1 -- measured slower and therefore not the default,
    profiles/r2_probe_deferred_tail_ab.txt) must stay correct: the switch is read once per process, so the LP probe tests
    that stress it (clustered keys, extreme keys, DRAM-resident properties, all strategies) run again in a child process."""
    import subprocess
    import sys

    if os.environ.get("CCB_LEAN_DEFERRED_TAIL") == "1":
        pytest.skip("already inside the child run")
    env = dict(os.environ, CCB_LEAN_DEFERRED_TAIL="1")
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-q", "-m", "gpu", "-x", "-k",
                        "clustered_keys or negative_and_extreme or probe_batch_large_properties or (probe_batch_matches_oracle and 1024)"],
                       capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-2000:]
    assert " passed" in p.stdout and "failed" not in p.stdout


def test_probe_batch_microbench_known_answer(ccb):
    """simd_micro_bench --scale 3 --hit-frequency 2: #tuples == 67114250 over 2^27 glibc rand() keys."""
    keys = O.gen_keys_rand(1 << 27, 1024 * 2 - 1)
    for T in (ccb.LPHashTable, ccb.HashTable):
        r = T(1024, 1).probe_batch(dev(keys), materialize=False)
        assert r["n_matches"] == 67114250


def test_probe_batch_large_properties(ccb, strategy):
    """Size-independent properties at a DRAM-resident size: hit=1 => every probe matches exactly once,
    key checksum == payload checksum == sum of inputs; hit=2 => matches are exactly the keys < n."""
    n = 1 << 24
    if strategy != "direct":
        ccb.set_probe_strategy(2 if strategy == "partitioned" else 3, 8 << 20)
    for T in (ccb.LPHashTable, ccb.HashTable):
        tab = T(n, 1)
        assert tab.info().has_duplicates == 0
        for hit in (1, 2):
            keys = ccb.gen_keys_counter(1 << 26, 2, n * hit - 1)
            r = tab.probe_batch(keys, capacity=keys.numel())
            hits = keys[keys < n]
            assert r["n_matches"] == hits.numel()
            s = int(hits.sum().item()) & ((1 << 64) - 1)
            assert r["key_sum"] == s and r["payload_sum"] == s
            m = r["n_matches"]
            assert torch.equal(r["out_key"][:m], r["out_payload"][:m])
            assert torch.equal(torch.sort(r["out_key"][:m]).values, torch.sort(hits).values)
        del tab


def test_probe_batch_host_end_to_end(ccb):
    rng = np.random.Generator(np.random.PCG64(3))
    n = 100000
    keys = rng.integers(0, 2 * n, size=3_000_000, dtype=np.int64)
    tab = ccb.LPHashTable(n, 2)
    want = O.pipeline([O.OracleLP(O.build_keys(n, 2))], keys.reshape(-1, 1), 2048, collect=True)
    ok = np.empty(keys.size * 2, dtype=np.int64)
    op = np.empty(keys.size * 2, dtype=np.int64)
    r = tab.probe_batch_host(keys, ok, op)
    m = r["n_matches"]
    assert m == want["n_tuples"] and r["overflow"] == 0
    got = np.stack([ok[:m], np.zeros(m, dtype=np.int64), op[:m]], axis=1)
    assert np.array_equal(G.sort_rows(got), G.sort_rows(want["tuples"]))


# ------------------------------------------------------------------ compactor
def test_compactor_protocol(ccb):
    """NaiveCompactor::Compact/Flush (compactor.cpp:5-41): only full chunks leave, rows preserved in order."""
    rng = np.random.Generator(np.random.PCG64(21))
    B, ncol = 256, 3
    comp = ccb.NaiveCompactor(ncol, B)
    emitted, fed = [], []
    for it in range(40):
        cnt = int(rng.integers(0, B + 1)) if it % 7 else B
        ch = ccb.DataChunk(ncol, B)
        data = rng.integers(0, 1 << 40, size=(B, ncol), dtype=np.int64)
        for c in range(ncol):
            ch.data_[c].data_.copy_(dev(data[:, c]))
        sel = rng.permutation(B).astype(np.uint32)
        ch.selection_vector_ = dev(sel.view(np.int32))
        ch.count_ = cnt
        fed.append(data[sel[:cnt]])
        out = comp.Compact(ch)
        if out.count_:
            assert out.count_ == B
            emitted.append(out.rows())
    tail = comp.Flush()
    emitted.append(tail.rows())
    fed_rows = np.concatenate(fed)
    got = np.concatenate(emitted)
    assert got.shape == fed_rows.shape
    assert np.array_equal(G.sort_rows(got), G.sort_rows(fed_rows))
    # threshold semantics (Binary/Dynamic)
    comp2 = ccb.Compactor(ncol, B, threshold=64)
    assert comp2.GetThreshold() == 64
    comp2.SetThreshold(100)
    ch = ccb.DataChunk(ncol, B)
    ch.count_ = 120
    assert comp2.Compact(ch) is ch and ch.count_ == 120  # >= threshold: untouched
    ch.count_ = 50
    assert comp2.Compact(ch).count_ == 0  # buffered


# ------------------------------------------------------------------ fused join chain
THRESHOLD_SETS = [None, "zero", "mixed"]


def thresholds_for(kind, J):
    if kind is None:
        return None
    if kind == "zero":
        return [0] * J
    return [32, 512, 128, 64, 256, 0, 384, 768][:J]


@pytest.mark.parametrize("thr", THRESHOLD_SETS)
def test_chain_execute_matches_reference_golden(ccb, thr):
    """main.cpp pipelines (mt19937(2) LHS) through the fused kernel: count, digest, column sums, rows per level."""
    seen = set()
    for g in G.load_index()["pipeline_main"]:
        key = (g["J"], g["cf"], g["lhs"], g["rhs"])
        if key in seen:
            continue
        seen.add(key)
        J = g["J"]
        lhs = O.gen_lhs_main(g["lhs"], J, g["rhs"])
        cols = [dev(lhs[:, j].copy()) for j in range(J)]
        for T in (ccb.HashTable, ccb.LPHashTable):
            tables = [T(g["rhs"], g["cf"]) for _ in range(J)]
            r = ccb.chain_execute(tables, cols, thresholds=thresholds_for(thr, J))
            assert r["n_tuples"] == g["n_tuples"] and r["digest"] == g["digest"], (g, T)
            assert r["colsum"] == g["colsum"] and r["level_in"] == g["level_in"]
            assert r["probe_tuples"] == g["probe_tuples"]


def test_chain_execute_materialized_tuples(ccb):
    for g in G.load_index()["pipeline_explicit"]:
        J = g["J"]
        lhs = G.load_i64(f"{g['name']}_lhs.bin").reshape(-1, J)
        want = G.load_i64(f"{g['name']}_k{g['kind']}_tuples.bin").reshape(-1, 3 * J)
        T = ccb.LPHashTable if g["kind"] == 0 else ccb.HashTable
        tables = [T(g["rhs"], g["cf"]) for _ in range(J)]
        cols = [dev(lhs[:, j].copy()) for j in range(J)]
        for thr in (None, [0] * J, [100] * J):
            r = ccb.chain_execute(tables, cols, thresholds=thr, materialize=True, capacity=want.shape[0] + 10)
            assert r["n_tuples"] == want.shape[0] and r["overflow"] == 0
            got = torch.stack([c[: r["n_tuples"]] for c in r["out_cols"]], dim=1).cpu().numpy()
            assert np.array_equal(G.sort_rows(got), want)
        r = ccb.chain_execute(tables, cols, materialize=True, capacity=max(1, want.shape[0] // 2))
        assert r["overflow"] == (1 if want.shape[0] > 1 else 0) and r["n_tuples"] == want.shape[0]


def test_chain_execute_compaction_densifies(ccb):
    """Full compaction must need clearly fewer rounds, with clearly more live lanes each, than no compaction on a sparse chain
    (SURVEY 10).  The golden LHS is repeated 30 times so that every pipeline instance gets dozens of chunks: with less than one
    chunk per instance (300 000 rows over ~3500 instances) the run is one long flush of half-filled caches under any threshold."""
    J, cf, rhs, rows, reps = 4, 8, 20000, 300000, 30
    lhs = np.tile(O.gen_lhs_main(rows, J, rhs), (reps, 1))
    cols = [dev(lhs[:, j].copy()) for j in range(J)]
    tables = [ccb.HashTable(rhs, cf) for _ in range(J)]
    full = ccb.chain_execute(tables, cols)
    none = ccb.chain_execute(tables, cols, thresholds=[0] * J)
    want = (270336 * reps, (10954991527034855424 * reps) % (1 << 64))  # count and digest are sums over the result tuples
    assert (full["n_tuples"], full["digest"]) == (none["n_tuples"], none["digest"]) == want
    # A round of the fused kernel inspects a whole sector or two of a lane's chain and emits ALL of its matches at once (the GPU
    # form of InOneNext), so even without compaction a level hands down denser chunks than the reference's one-match-per-Next
    # protocol; what the threshold controls is how many lanes a round runs with.
    assert sum(full["level_steps"]) * 1.5 < sum(none["level_steps"]), (full["level_steps"], none["level_steps"])
    dens_full = sum(full["level_lanes"]) / max(1, sum(full["level_steps"]))
    dens_none = sum(none["level_lanes"]) / max(1, sum(none["level_steps"]))
    assert dens_full > 1.5 * dens_none, (dens_full, dens_none)


def test_chain_telemetry_histograms(ccb, tmp_path):
    """cc_chain_execute_ex: the chunk-density histograms (ZebraProfiler analogue, profiler.h:168-260).  Their totals must equal the
    step counters of cc_chain_result, full compaction must run clearly more of its Next rounds in the top density bin than no
    compaction does, histograms of several calls add up, and the CSV dump has one row per histogram, level and bin.
    (The golden LHS is repeated 30 times so that every pipeline instance gets dozens of chunks: with a few chunks per instance
    the final flush of the half-filled caches would dominate every histogram.)"""
    J, cf, rhs, rows, reps = 4, 8, 20000, 300000, 30
    lhs = np.tile(O.gen_lhs_main(rows, J, rhs), (reps, 1))
    cols = [dev(lhs[:, j].copy()) for j in range(J)]
    tables = [ccb.HashTable(rhs, cf) for _ in range(J)]
    tel_full, tel_none = ccb.new_chain_telemetry(), ccb.new_chain_telemetry()
    full = ccb.chain_execute(tables, cols, telemetry=tel_full)
    none = ccb.chain_execute(tables, cols, thresholds=[0] * J, telemetry=tel_none)
    want = (270336 * reps, (10954991527034855424 * reps) % (1 << 64))  # count and digest are sums over the result tuples
    assert (full["n_tuples"], full["digest"]) == (none["n_tuples"], none["digest"]) == want
    hf, hn = ccb.parse_chain_telemetry(tel_full, J), ccb.parse_chain_telemetry(tel_none, J)
    for r, h in ((full, hf), (none, hn)):
        assert [sum(h["round_lanes_hist"][l]) for l in range(J)] == r["level_steps"]
        assert all(sum(h["probe_rows_hist"][l]) > 0 for l in range(J))
    top = lambda h: sum(h["round_lanes_hist"][l][-1] for l in range(J)) / max(1, sum(sum(h["round_lanes_hist"][l]) for l in range(J)))
    assert top(hf) > 0.5 and top(hf) > 1.5 * top(hn), (top(hf), top(hn), hf["round_lanes_hist"], hn["round_lanes_hist"])
    again = ccb.chain_execute(tables, cols, telemetry=tel_full)  # accumulates (the number of rounds depends on the scheduling of the run)
    h2 = ccb.parse_chain_telemetry(tel_full, J)
    assert all(sum(h2["round_lanes_hist"][l]) == sum(hf["round_lanes_hist"][l]) + again["level_steps"][l] for l in range(J))
    path = str(tmp_path / "density.csv")
    ccb.chain_telemetry_csv(tel_none, J, path)
    lines = open(path).read().strip().splitlines()
    assert lines[0] == "histogram,level,density_from,density_to,chunks" and len(lines) == 1 + 2 * J * 8
    assert sum(int(l.split(",")[-1]) for l in lines[1:] if l.startswith("round_lanes")) == sum(none["level_steps"])


def test_chain_execute_edge_cases(ccb):
    t = ccb.HashTable(100, 1)
    empty = torch.empty(0, dtype=torch.int64, device="cuda")
    r = ccb.chain_execute([t, t], [empty, empty])
    assert r["n_tuples"] == 0 and r["level_in"] == [0, 0]
    one = dev(np.array([5], dtype=np.int64))
    r = ccb.chain_execute([t], [one], materialize=True, capacity=4)
    assert r["n_tuples"] == 1 and [int(c[0]) for c in r["out_cols"]] == [5, 0, 5]
    miss = dev(np.array([1000, 2000, 3000], dtype=np.int64))
    r = ccb.chain_execute([t, t], [miss, miss])
    assert r["n_tuples"] == 0 and r["level_in"] == [3, 0]


# ------------------------------------------------------------------ partitioning
@pytest.mark.parametrize("log2p,n,sequential", [(0, 333333, False), (1, 333333, False), (3, 333333, False), (8, 333333, False),
                                                 (0, 3 << 22, True), (4, (3 << 22) + 4099, True), (9, 1 << 24, False)])
def test_partition_kernels(ccb, log2p, n, sequential):
    """Many tiles per CTA (the TMA refill path of the scatter kernel) and sequential keys included: an early version lost
    whole warp slices of a tile when the async refill raced with the reads of the staging buffer."""
    rng = np.random.Generator(np.random.PCG64(log2p))
    keys = np.arange(n, dtype=np.int64) if sequential else rng.integers(-(1 << 62), 1 << 62, size=n, dtype=np.int64)
    out, counts, offsets = ccb.partition_keys(dev(keys), log2p)
    out = out.cpu().numpy()
    P = 1 << log2p
    pid = (O.murmurhash64(keys.view(np.uint64)) >> np.uint64(64 - log2p)).astype(np.int64) if log2p else np.zeros(keys.size, dtype=np.int64)
    assert np.array_equal(counts, np.bincount(pid, minlength=P))
    for p in range(P):
        seg = out[offsets[p]:offsets[p] + counts[p]]
        assert np.array_equal(np.sort(seg), np.sort(keys[pid == p]))


def test_chain_execute_tuned_negative_feedback(ccb):
    """Dynamic compaction: bandit-chosen thresholds never change the result, and the policy learns that
    compacting pays off on a sparse chain (most pulls go to the large thresholds)."""
    J, cf, rhs, rows = 4, 8, 20000, 300000
    lhs = O.gen_lhs_main(rows, J, rhs)
    cols = [dev(lhs[:, j].copy()) for j in range(J)]
    tables = [ccb.HashTable(rhs, cf) for _ in range(J)]
    tuner = ccb.CompactTuner()
    for l in range(J):
        tuner.Initialize(0x1000 + l)
    total_pulls = 0
    for rep in range(12):  # 12 passes x 30 batches = 360 pulls per bandit (warm-up is 36)
        r = ccb.chain_execute_tuned(tables, cols, tuner, batch_rows=10000)
        assert (r["n_tuples"], r["digest"]) == (270336, 10954991527034855424)
        assert r["level_in"][0] == rows
        total_pulls += 30
    arms = np.array(ccb.DEFAULT_ARMS)
    for l in range(1, J):  # the compactors in front of joins 1..J-1 see sparse chunks
        rewards, selects = ccb.CompactTuner.state(tuner, l - 1)
        assert int(selects.sum()) == total_pulls
        assert selects[arms >= 256].sum() > selects[arms < 64].sum(), (l, selects)
    # materialised output across batches
    r = ccb.chain_execute_tuned(tables, cols, tuner, batch_rows=77777, materialize=True, capacity=270336)
    got = torch.stack([c[: r["n_tuples"]] for c in r["out_cols"]], dim=1).cpu().numpy()
    want = O.pipeline([O.OracleChain(O.build_keys(rhs, cf)) for _ in range(J)], lhs, 256, collect=True)
    assert r["overflow"] == 0 and np.array_equal(G.sort_rows(got), G.sort_rows(want["tuples"]))


@pytest.mark.parametrize("log2p,n", [(0, 3 << 22), (1, 3 << 22), (3, (3 << 20) + 12345), (2, 333333)])
def test_partition_scatter_peers(ccb, log2p, n):
    """cc_partition_scatter_peers: partition p lands in its own destination buffer (peer memory in production),
    at the base offset handed in, complete and uncorrupted -- many tiles per CTA (TMA refill path) and ragged tails."""
    import ctypes as C

    rng = np.random.Generator(np.random.PCG64(n + log2p))
    keys = rng.integers(-(1 << 62), 1 << 62, size=n, dtype=np.int64)
    P = 1 << log2p
    pid = (O.murmurhash64(keys.view(np.uint64)) >> np.uint64(64 - log2p)).astype(np.int64) if log2p else np.zeros(n, dtype=np.int64)
    counts = np.bincount(pid, minlength=P)
    base = np.arange(P, dtype=np.int64) * 7 + 3  # arbitrary start row of this sender inside every destination buffer
    bufs = [torch.full((int(counts[p] + base[p]) + 16,), -7, dtype=torch.int64, device="cuda") for p in range(P)]
    dkeys = dev(keys)
    dbase = dev(base)
    cursors = torch.zeros(P, dtype=torch.int64, device="cuda")
    ptrs = (C.c_void_p * P)(*[b.data_ptr() for b in bufs])
    for blocks in (0, 8):
        ccb._lib.check(ccb.lib().cc_partition_set_peer_blocks(blocks))
        for b in bufs:
            b.fill_(-7)
        ccb._lib.check(ccb.lib().cc_partition_scatter_peers(dkeys.data_ptr(), n, log2p, dbase.data_ptr(), cursors.data_ptr(), ptrs,
                                                            torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        for p in range(P):
            got = bufs[p].cpu().numpy()
            assert np.all(got[: base[p]] == -7) and np.all(got[base[p] + counts[p]:] == -7)
            assert np.array_equal(np.sort(got[base[p]: base[p] + counts[p]]), np.sort(keys[pid == p])), (p, blocks)
    ccb._lib.check(ccb.lib().cc_partition_set_peer_blocks(0))


@pytest.mark.parametrize("kind", [0, 1])
def test_probe_batch_skewed_keys_take_the_fallback(ccb, kind):
    """Single-pass partition regions have 12.5 % slack; heavily skewed probe keys overrun a region, which must
    switch (on the device) to the two-pass partition with identical results."""
    n = 200000
    bk = O.build_keys(n, 2)
    T, OT = (ccb.LPHashTable, O.OracleLP) if kind == 0 else (ccb.HashTable, O.OracleChain)
    tab, otab = T(n, 2), OT(bk)
    rng = np.random.Generator(np.random.PCG64(77))
    for name, keys in {"all_equal_hit": np.full(300000, 1234, dtype=np.int64), "all_equal_miss": np.full(300000, 1235, dtype=np.int64),
                       "half_one_key": np.where(rng.random(400000) < 0.5, 4242, rng.integers(0, 2 * n, size=400000)).astype(np.int64),
                       "two_hot_keys": rng.choice(np.array([10, 199998], dtype=np.int64), size=250001)}.items():
        want = O.pipeline([otab], keys.reshape(-1, 1), 2048)
        try:
            ccb.set_probe_strategy(2, 64 << 10)
            r = tab.probe_batch(dev(keys), capacity=2 * keys.size + 8)
        finally:
            ccb.set_probe_strategy(0, 32 << 20)
        assert (r["n_matches"], r["key_sum"], r["payload_sum"], r["overflow"]) == (want["n_tuples"], want["colsum"][0], want["colsum"][2], 0), name
        m = r["n_matches"]
        got = r["out_key"][:m].cpu().numpy()
        assert np.array_equal(np.sort(got), np.sort(np.repeat(keys[np.isin(keys, bk)], 2)))
