"""CPU test of the partitioned join's addressing logic: tests/native/partfn_host_check.cu is compiled with nvcc and run on the host
(no GPU needed) -- PartFn's owner / slice / fused ids against an independent numpy computation on the oracle's hash, and the
walk-order -> arena-region mapping of SegIn (slice-major walk over a [piece][sender][slice] arena)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_lib as O
from conftest import PKG_NAME, ROOT


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_partition_function_and_region_mapping(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "partfn_host_check")
    subprocess.check_call([nvcc, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, PKG_NAME, "csrc"),
                           os.path.join(ROOT, "tests", "native", "partfn_host_check.cu"), "-o", exe])
    out = subprocess.check_output([exe]).decode().strip().splitlines()
    rows = np.array([[int(v) for v in ln.split()] for ln in out[:-1]], dtype=np.uint64)
    assert rows.shape == (2000, 5)
    h = O.murmurhash64(rows[:, 0].copy())
    owner = h >> np.uint64(61)                                               # high log2(8) hash bits
    slice_ = ((h & np.uint64((1 << 20) - 1)) >> np.uint64(14)) & np.uint64(63)  # high 6 bits of the home slot in a 2^20-slot table
    assert np.array_equal(rows[:, 2], owner)
    assert np.array_equal(rows[:, 3], slice_)
    assert np.array_equal(rows[:, 1], owner * np.uint64(64) + slice_)  # fused id = owner * S + slice: owner o's regions are contiguous
    assert np.array_equal(rows[:, 4], slice_)                          # one rank: the slice alone
    assert len(set(rows[:, 1].tolist())) > 400                         # the ids spread over the 512 regions
    regions = [int(v) for v in out[-1].split()[1:]]
    # walk index p = slice * (pieces * senders) + (piece * senders + sender)  ->  region (piece * senders + sender) * slices_allocated + slice
    want = [(p % 12) * 5 + p // 12 for p in range(60)]
    assert regions == want
    assert sorted(regions) == list(range(60))  # a permutation: every region is probed exactly once


@pytest.mark.parametrize("world", [2, 4, 8])
def test_local_comm_control_plane(tmp_path, world):
    """LocalComm (host/simd_compaction.hpp), the fork + shared-memory control plane of host/pjoin_main.cpp, driven through the
    cc_comm callbacks the library calls: all-gathers and barriers across forked ranks, and a failing rank releases the others
    instead of leaving them in a barrier forever."""
    pkg = os.path.join(ROOT, PKG_NAME)
    exe = str(tmp_path / "localcomm_check")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "host"),
                           os.path.join(ROOT, "tests", "native", "localcomm_check.cpp"), "-o", exe, "-L", pkg, "-lccb200", f"-Wl,-rpath,{pkg}"])
    assert subprocess.run([exe, str(world)], timeout=120).returncode == 0
    assert subprocess.run([exe, str(world), "1"], timeout=120).returncode == 0
