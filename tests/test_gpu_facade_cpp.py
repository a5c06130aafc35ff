"""GPU test of the C++ facade (host/simd_compaction.hpp): the ported main.cpp driver
(host/pipeline_main.cpp) must reproduce the reference's known answers, both through the
chunk-granular Probe/Next/Compact protocol and through the fused chain kernel."""
import json
import os
import subprocess

import pytest

import golden_util as G
from conftest import PKG_NAME, ROOT

pytestmark = pytest.mark.gpu
DRIVER = os.path.join(ROOT, PKG_NAME, "host", "pipeline_main")


def run(*args):
    out = subprocess.check_output([DRIVER] + [str(a) for a in args], timeout=900)
    return json.loads(out.decode().strip().splitlines()[-1])


@pytest.mark.parametrize("J,cf,lhs,rhs", [(2, 2, 10000, 1000), (3, 5, 50000, 5000)])
def test_cpp_driver_matches_reference(ccb, J, cf, lhs, rhs):
    assert os.path.exists(DRIVER), "build it with make -C <pkg>/csrc driver"
    gold = {(g["J"], g["cf"], g["lhs"], g["rhs"]): g for g in G.load_index()["pipeline_main"]}
    g = gold[(J, cf, lhs, rhs)]
    base = ["--join-num", J, "--chunk-factor", cf, "--lhs-size", lhs, "--rhs-size", rhs]
    variants = [
        ["--block", g["block"], "--compact", "none"],
        ["--block", g["block"], "--compact", "full"],
        ["--block", g["block"], "--compact", 64],
        ["--block", g["block"], "--table", "lp"],
        ["--mode", "fused", "--compact", "full"],
        ["--mode", "fused", "--compact", "none"],
        ["--mode", "fused", "--compact", 128, "--table", "lp"],
    ]
    for v in variants:
        r = run(*base, *v)
        assert (r["n_tuples"], r["digest"], r["colsum"]) == (g["n_tuples"], g["digest"], g["colsum"]), (v, r)


@pytest.mark.parametrize("table", ["chain", "lp"])
def test_cpp_driver_payload_mode(ccb, table):
    """facade HashTable(n, cf, keep_payload) + ProbeBatchPayload vs the payload oracle"""
    import numpy as np

    import oracle_lib as O

    J, cf, lhs, rhs = 2, 4, 30000, 3000
    r = run("--join-num", J, "--chunk-factor", cf, "--lhs-size", lhs, "--rhs-size", rhs, "--mode", "payload", "--table", table)
    keys = O.gen_lhs_main(lhs, J, rhs)[:, 0]
    want = O.join_payload(0 if table == "lp" else 1, O.build_keys(rhs, cf), [O.ref_payload(rhs)], keys)
    digest, colsum = O.digest_tuples(want)
    assert (r["n_tuples"], r["colsum"]) == (want.shape[0], colsum)
    assert r["digest"] == digest  # the digest is a sum over rows, so row order does not matter


MICRO = os.path.join(ROOT, PKG_NAME, "host", "micro_bench_main")


@pytest.mark.parametrize("scale,hit,cf", [(3, 2, 1), (3, 1, 4), (0, 4, 8)])
def test_cpp_micro_bench_driver(ccb, scale, hit, cf):
    """the ported simd_micro_bench driver: every variant (chunk protocol Next / InOneNext and the fused batch probe, both
    table kinds) must print the same #tuples as the oracle's scalar Probe + Next over the same glibc rand() keys"""
    import oracle_lib as O

    assert os.path.exists(MICRO), "build it with make -C <pkg>/csrc driver"
    n_keys = 1 << (20 if scale else 17)  # scale 0 = 256-row blocks: one C-ABI call per Probe / Next, 110 s for 2^20 keys
    out = subprocess.check_output([MICRO, "--scale", str(scale), "--hit-frequency", str(hit), "--chunk-factor", str(cf),
                                   "--lhs-tuples", str(n_keys)], timeout=900, stderr=subprocess.DEVNULL)
    r = json.loads(out.decode().strip().splitlines()[-1])
    n_rhs, block = 128 << scale, 256 << scale
    keys = O.gen_keys_rand(n_keys, n_rhs * hit - 1)
    want, _ = O.microbench(O.OracleLP(O.build_keys(n_rhs, cf)), keys, block)
    want_chain, _ = O.microbench(O.OracleChain(O.build_keys(n_rhs, cf)), keys, block)
    assert want == want_chain
    assert len(r["variants"]) == 6
    for v in r["variants"]:
        assert v["n_tuples"] == want, v
