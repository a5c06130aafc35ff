"""ctypes binding of oracle/libcc_oracle.so -- the CPU parity checker.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never by the product
package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libcc_oracle.so")
REF_DRIVER = os.path.join(ORACLE_DIR, "_ref", "ref_driver")

NIL = 0xFFFFFFFF


class ResultStats(C.Structure):
    _fields_ = [
        ("n_tuples", C.c_uint64),
        ("digest", C.c_uint64),
        ("colsum", C.c_uint64 * 64),
        ("probe_tuples", C.c_uint64),
        ("level_in", C.c_uint64 * 16),
        ("level_chunks", C.c_uint64 * 16),
        ("next_calls", C.c_uint64),
    ]


class PipelineCfg(C.Structure):
    _fields_ = [
        ("n_joins", C.c_size_t),
        ("block", C.c_size_t),
        ("table_kind", C.c_int),
        ("use_inone", C.c_int),
        ("compaction", C.c_int),
        ("threshold", C.c_size_t),
        ("collect", C.c_int),
    ]


class LPTable(C.Structure):
    _fields_ = [("n_slots", C.c_size_t), ("slots", C.POINTER(C.c_int64))]


class ChainTable(C.Structure):
    _fields_ = [
        ("n_buckets", C.c_size_t),
        ("n", C.c_size_t),
        ("key", C.POINTER(C.c_int64)),
        ("next", C.POINTER(C.c_uint32)),
        ("head", C.POINTER(C.c_uint32)),
        ("tail", C.POINTER(C.c_uint32)),
    ]


class Chunk(C.Structure):
    _fields_ = [
        ("block", C.c_size_t),
        ("ncol", C.c_size_t),
        ("count", C.c_size_t),
        ("col", C.POINTER(C.POINTER(C.c_int64))),
        ("own", C.POINTER(C.POINTER(C.c_int64))),
        ("sel", C.POINTER(C.c_uint32)),
    ]


_lib = None


def build(force: bool = False) -> None:
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB_PATH)
    vp, sz, u64, i64p, u32p = C.c_void_p, C.c_size_t, C.c_uint64, C.POINTER(C.c_int64), C.POINTER(C.c_uint32)
    L.orc_murmurhash64.restype = u64
    L.orc_murmurhash64.argtypes = [u64]
    L.orc_murmurhash64_batch.argtypes = [vp, vp, sz]
    L.orc_build_keys.argtypes = [sz, sz, vp]
    L.orc_gen_lhs_main.argtypes = [sz, sz, sz, vp]
    L.orc_gen_keys_rand.argtypes = [sz, u64, vp]
    L.orc_gen_keys_counter.argtypes = [sz, u64, u64, u64, vp]
    L.orc_lp_build.restype = C.POINTER(LPTable)
    L.orc_lp_build.argtypes = [vp, sz]
    L.orc_lp_build_reference.restype = C.POINTER(LPTable)
    L.orc_lp_build_reference.argtypes = [sz, sz]
    L.orc_lp_free.argtypes = [C.POINTER(LPTable)]
    L.orc_chain_build.restype = C.POINTER(ChainTable)
    L.orc_chain_build.argtypes = [vp, sz]
    L.orc_chain_build_reference.restype = C.POINTER(ChainTable)
    L.orc_chain_build_reference.argtypes = [sz, sz]
    L.orc_chain_free.argtypes = [C.POINTER(ChainTable)]
    L.orc_chunk_new.restype = C.POINTER(Chunk)
    L.orc_chunk_new.argtypes = [sz, sz]
    L.orc_chunk_free.argtypes = [C.POINTER(Chunk)]
    L.orc_chunk_reset.argtypes = [C.POINTER(Chunk)]
    L.orc_chunk_slice.argtypes = [C.POINTER(Chunk), C.POINTER(Chunk), vp, sz]
    L.orc_chunk_append.argtypes = [C.POINTER(Chunk), C.POINTER(Chunk), sz, sz]
    L.orc_lp_probe.restype = vp
    L.orc_lp_probe.argtypes = [C.POINTER(LPTable), vp, sz, vp, sz]
    L.orc_chain_probe.restype = vp
    L.orc_chain_probe.argtypes = [C.POINTER(ChainTable), vp, sz, vp, sz]
    L.orc_scan_has_next.argtypes = [vp]
    L.orc_scan_next.restype = sz
    L.orc_scan_next.argtypes = [vp, vp, C.POINTER(Chunk), C.POINTER(Chunk)]
    L.orc_scan_inone_next.restype = sz
    L.orc_scan_inone_next.argtypes = [vp, vp, C.POINTER(Chunk), C.POINTER(Chunk)]
    L.orc_scan_free.argtypes = [vp]
    L.orc_compactor_new.restype = vp
    L.orc_compactor_new.argtypes = [sz, sz, sz]
    L.orc_compactor_free.argtypes = [vp]
    L.orc_compactor_compact.argtypes = [vp, C.POINTER(C.POINTER(Chunk))]
    L.orc_compactor_flush.argtypes = [vp, C.POINTER(C.POINTER(Chunk))]
    L.orc_pipeline.restype = C.c_int
    L.orc_pipeline.argtypes = [C.POINTER(PipelineCfg), C.POINTER(vp), vp, sz, C.POINTER(ResultStats), C.POINTER(i64p)]
    L.orc_multiplicity_oracle.restype = C.c_int
    L.orc_multiplicity_oracle.argtypes = [sz, C.POINTER(vp), C.POINTER(sz), vp, sz, C.POINTER(ResultStats)]
    L.orc_digest_tuples.restype = u64
    L.orc_digest_tuples.argtypes = [vp, sz, sz, vp]
    L.orc_microbench_lp.restype = u64
    L.orc_microbench_lp.argtypes = [C.POINTER(LPTable), vp, sz, sz, C.c_int, C.POINTER(u64)]
    L.orc_microbench_chain.restype = u64
    L.orc_microbench_chain.argtypes = [C.POINTER(ChainTable), vp, sz, sz, C.c_int, C.POINTER(u64)]
    L.orc_bandit_new.restype = vp
    L.orc_bandit_new.argtypes = [sz]
    L.orc_bandit_free.argtypes = [vp]
    L.orc_bandit_select.restype = sz
    L.orc_bandit_select.argtypes = [vp]
    L.orc_bandit_update.argtypes = [vp, sz, C.c_double]
    L.orc_bandit_state.argtypes = [vp, vp, vp]
    L.orc_ref_payload.argtypes = [sz, vp]
    L.orc_join_payload.restype = C.c_int
    L.orc_join_payload.argtypes = [C.c_int, vp, sz, C.POINTER(vp), sz, vp, sz, C.POINTER(i64p), C.POINTER(sz)]
    L.orc_free.argtypes = [vp]
    _lib = L
    return L


def _p(a: np.ndarray) -> int:
    return a.ctypes.data


# ---------------------------------------------------------------- generators
def murmurhash64(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.uint64)
    out = np.empty_like(x)
    lib().orc_murmurhash64_batch(_p(x), _p(out), x.size)
    return out


def build_keys(n: int, cf: int) -> np.ndarray:
    out = np.empty(n, dtype=np.int64)
    lib().orc_build_keys(n, cf, _p(out))
    return out


def gen_lhs_main(rows: int, n_joins: int, rhs_size: int) -> np.ndarray:
    out = np.empty((rows, n_joins), dtype=np.int64)
    lib().orc_gen_lhs_main(rows, n_joins, rhs_size, _p(out))
    return out


def gen_keys_rand(n: int, mask: int) -> np.ndarray:
    out = np.empty(n, dtype=np.int64)
    lib().orc_gen_keys_rand(n, mask, _p(out))
    return out


def gen_keys_counter(n: int, seed: int, mask: int, first: int = 0) -> np.ndarray:
    out = np.empty(n, dtype=np.int64)
    lib().orc_gen_keys_counter(n, seed, first, mask, _p(out))
    return out


# -------------------------------------------------------------------- tables
class OracleLP:
    def __init__(self, keys: np.ndarray):
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        self.ptr = lib().orc_lp_build(_p(keys), keys.size)
        self.kind = 0

    @property
    def n_slots(self) -> int:
        return self.ptr.contents.n_slots

    def slots(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.ptr.contents.slots, shape=(self.n_slots,)).copy()

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().orc_lp_free(self.ptr)
            self.ptr = None


class OracleChain:
    def __init__(self, keys: np.ndarray):
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        self.ptr = lib().orc_chain_build(_p(keys), keys.size)
        self.kind = 1

    @property
    def n_buckets(self) -> int:
        return self.ptr.contents.n_buckets

    @property
    def n(self) -> int:
        return self.ptr.contents.n

    def arrays(self):
        c = self.ptr.contents
        n = max(c.n, 1)
        return (
            np.ctypeslib.as_array(c.key, shape=(n,))[: c.n].copy(),
            np.ctypeslib.as_array(c.next, shape=(n,))[: c.n].copy(),
            np.ctypeslib.as_array(c.head, shape=(c.n_buckets,)).copy(),
        )

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().orc_chain_free(self.ptr)
            self.ptr = None


def stats_dict(st: ResultStats, n_joins: int) -> dict:
    return {
        "n_tuples": int(st.n_tuples),
        "digest": int(st.digest),
        "colsum": [int(st.colsum[i]) for i in range(3 * n_joins)],
        "probe_tuples": int(st.probe_tuples),
        "level_in": [int(st.level_in[i]) for i in range(n_joins)],
        "level_chunks": [int(st.level_chunks[i]) for i in range(n_joins)],
        "next_calls": int(st.next_calls),
    }


def pipeline(tables, lhs: np.ndarray, block: int, *, use_inone=False, compaction=0, threshold=0, collect=False):
    """Run the reference pipeline protocol (main.cpp:79-191) over explicit tables/LHS."""
    lhs = np.ascontiguousarray(lhs, dtype=np.int64)
    rows, J = lhs.shape
    assert len(tables) == J
    cfg = PipelineCfg(J, block, tables[0].kind, int(use_inone), compaction, threshold, int(collect))
    tp = (C.c_void_p * J)(*[C.cast(t.ptr, C.c_void_p) for t in tables])
    st = ResultStats()
    out = C.POINTER(C.c_int64)()
    rc = lib().orc_pipeline(C.byref(cfg), tp, _p(lhs), rows, C.byref(st), C.byref(out) if collect else None)
    assert rc == 0
    d = stats_dict(st, J)
    if collect:
        n = d["n_tuples"]
        if n:
            d["tuples"] = np.ctypeslib.as_array(out, shape=(n, 3 * J)).copy()
        else:
            d["tuples"] = np.empty((0, 3 * J), dtype=np.int64)
        lib().orc_free(out)
    return d


def multiplicity_oracle(build_key_arrays, lhs: np.ndarray) -> dict:
    lhs = np.ascontiguousarray(lhs, dtype=np.int64)
    rows, J = lhs.shape
    arrs = [np.ascontiguousarray(a, dtype=np.int64) for a in build_key_arrays]
    kp = (C.c_void_p * J)(*[_p(a) for a in arrs])
    ns = (C.c_size_t * J)(*[a.size for a in arrs])
    st = ResultStats()
    assert lib().orc_multiplicity_oracle(J, kp, ns, _p(lhs), rows, C.byref(st)) == 0
    return stats_dict(st, J)


def ref_payload(n: int) -> np.ndarray:
    """payload the reference generates and drops: row i -> i + 10000000 (chaining_ht.cpp:21)"""
    out = np.empty(n, dtype=np.int64)
    lib().orc_ref_payload(n, _p(out))
    return out


def join_payload(kind: int, build_keys: np.ndarray, payload_cols, probe_keys: np.ndarray) -> np.ndarray:
    """orc_join_payload: rows (probe key, build key, payload...) of the single join that keeps the build row."""
    bk = np.ascontiguousarray(build_keys, dtype=np.int64)
    pk = np.ascontiguousarray(probe_keys, dtype=np.int64)
    cols = [np.ascontiguousarray(c, dtype=np.int64) for c in payload_cols]
    assert all(c.size == bk.size for c in cols)
    cp = (C.c_void_p * max(len(cols), 1))(*[_p(c) for c in cols])
    out = C.POINTER(C.c_int64)()
    n = C.c_size_t(0)
    assert lib().orc_join_payload(kind, _p(bk), bk.size, cp, len(cols), _p(pk), pk.size, C.byref(out), C.byref(n)) == 0
    w = 2 + len(cols)
    rows = np.ctypeslib.as_array(out, shape=(n.value, w)).copy() if n.value else np.empty((0, w), dtype=np.int64)
    lib().orc_free(out)
    return rows


def sort_rows(rows: np.ndarray) -> np.ndarray:
    """rows in lexicographic order (multiset comparison)"""
    rows = np.ascontiguousarray(rows)
    if rows.shape[0] == 0:
        return rows
    return rows[np.lexsort(rows.T[::-1])]


def digest_tuples(tuples: np.ndarray):
    tuples = np.ascontiguousarray(tuples, dtype=np.int64)
    n, nc = tuples.shape
    cs = np.zeros(nc, dtype=np.uint64)
    h = lib().orc_digest_tuples(_p(tuples), n, nc, _p(cs))
    return int(h), [int(x) for x in cs]


def microbench(table, keys: np.ndarray, block: int, inone: bool = False):
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    cs = C.c_uint64(0)
    f = lib().orc_microbench_lp if table.kind == 0 else lib().orc_microbench_chain
    n = f(table.ptr, _p(keys), keys.size, block, int(inone), C.byref(cs))
    return int(n), int(cs.value)


def scan_chunk(table, keys: np.ndarray, sel: np.ndarray, count: int, block: int, inone: bool = False):
    """Chunk-granular protocol for ONE input chunk: returns a list of
    (positions[u32], payloads[i64]) per Next call (physical LHS positions)."""
    L = lib()
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    sel = np.ascontiguousarray(sel, dtype=np.uint32)
    assert keys.size >= block and sel.size >= block
    inp = L.orc_chunk_new(1, block)
    res = L.orc_chunk_new(3, block)
    C.memmove(inp.contents.col[0], _p(keys), block * 8)
    C.memmove(inp.contents.sel, _p(sel), block * 4)
    inp.contents.count = count
    kcol = C.cast(inp.contents.col[0], C.c_void_p)
    probe = L.orc_lp_probe if table.kind == 0 else L.orc_chain_probe
    ss = probe(table.ptr, kcol, count, C.cast(inp.contents.sel, C.c_void_p), block)
    out = []
    while L.orc_scan_has_next(ss):
        rc = (L.orc_scan_inone_next if inone else L.orc_scan_next)(ss, kcol, inp, res)
        rsel = np.ctypeslib.as_array(res.contents.sel, shape=(block,))[:rc].copy()
        col2 = np.ctypeslib.as_array(res.contents.col[2], shape=(block,))
        out.append((rsel, col2[rsel].copy()))
    L.orc_scan_free(ss)
    L.orc_chunk_free(inp)
    L.orc_chunk_free(res)
    return out


class OracleBandit:
    def __init__(self, n_arms: int):
        self.n_arms = n_arms
        self.ptr = lib().orc_bandit_new(n_arms)

    def select(self) -> int:
        return lib().orc_bandit_select(self.ptr)

    def update(self, arm: int, reward: float) -> None:
        lib().orc_bandit_update(self.ptr, arm, reward)

    def state(self):
        r = np.zeros(self.n_arms, dtype=np.float64)
        s = np.zeros(self.n_arms, dtype=np.uint64)
        lib().orc_bandit_state(self.ptr, _p(r), _p(s))
        return r, s

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().orc_bandit_free(self.ptr)
            self.ptr = None


# ------------------------------------------------------- the real reference
def have_ref_driver() -> bool:
    if not os.path.exists(REF_DRIVER):
        return False
    try:
        with open("/proc/cpuinfo") as f:
            flags = f.read()
        return all(x in flags for x in ("avx512f", "avx512dq", "avx512vl", "avx512bw"))
    except OSError:
        return False


def ref_driver(*args, timeout=600) -> dict:
    import json

    out = subprocess.check_output([REF_DRIVER] + [str(a) for a in args], timeout=timeout)
    return json.loads(out.decode().strip().splitlines()[-1])
