"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/cc_api.h declares, fails loudly without a GPU, and its host-only policy code
(CompactTuner) is bit-identical to the reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import golden_util as G
from conftest import ROOT, load_pkg

HEADER = os.path.join(ROOT, "include", "cc_api.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    pkg = load_pkg()
    lib = pkg.lib()
    names = declared_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), f"{n} declared in cc_api.h but not exported by libccb200.so"
    # and the Python binding covers exactly the header
    assert sorted(pkg._lib.SIGNATURES) == names
    assert lib.cc_api_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_device():
    pkg = load_pkg()
    lib = pkg.lib()
    assert lib.cc_device_init(0) == pkg._lib.CC_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.cc_last_error()
    h = C.c_void_p()
    assert lib.cc_ht_build_reference(C.byref(h), 0, 16, 1, None) == pkg._lib.CC_ERR_NO_DEVICE
    r = pkg.ProbeResult()
    keys = np.zeros(4, dtype=np.int64)
    assert lib.cc_probe_batch_host(None, keys.ctypes.data, 4, None, None, 0, C.byref(r), None) != 0
    with pytest.raises(pkg.CCError):
        pkg.init(0)


def test_product_does_not_import_oracle():
    pkg_dir = os.path.join(ROOT, "chunk-compaction-in-vectorized-execution-simd_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "cc_oracle" not in txt and "oracle_lib" not in txt and "orc_" not in txt, f


def _drive(select, update, steps):
    arms = [0, 32, 64, 128, 256, 384, 512, 768, 1024]
    lcg = 88172645463325252
    out = []
    for i in range(steps):
        thr = select()
        lcg = (lcg * 6364136223846793005 + 1442695040888963407) & ((1 << 64) - 1)
        noise = float((lcg >> 33) % 1000) / 1000.0
        best = 256.0 if i < steps // 2 else 768.0
        scale = 1.0 if i < steps // 2 else 4.0
        update(thr, scale * (2.0 - abs(float(thr) - best) / 1024.0) + 0.05 * noise)
        out.append(thr)
    return out


def test_tuner_matches_reference_and_oracle():
    """CompactTuner (negative_feedback.hpp:165-260): same arm sequence as the real reference class
    (golden) and bit-identical FP64 state vs the oracle restatement."""
    import oracle_lib as O

    pkg = load_pkg()
    g = G.load_index()["bandit"]
    t = pkg.CompactTuner()
    t.Initialize(0xABC)
    assert t.GetId(0xABC) == 0 and t.GetId(7) == -1 and t.GetBanditSize() == 1
    got = _drive(lambda: t.SelectArm(0), lambda thr, r: t.UpdateArm(0, thr, r), g["steps"])
    assert got == g["arms"]
    arms = list(pkg.DEFAULT_ARMS)
    b = O.OracleBandit(len(arms))
    _drive(lambda: arms[b.select()], lambda thr, r: b.update(arms.index(thr), r), g["steps"])
    r1, s1 = t.state(0)
    r2, s2 = b.state()
    assert np.array_equal(r1.view(np.uint64), r2.view(np.uint64)) and np.array_equal(s1, s2)
    # unknown arm values are ignored (negative_feedback.hpp:193); double registration is an error
    t.UpdateArm(0, 12345, 1.0)
    with pytest.raises(pkg.CCError):
        t.Initialize(0xABC)


def test_tuner_log_csv(tmp_path):
    pkg = load_pkg()
    t = pkg.CompactTuner()
    t.Initialize(1, [0, 8, 16])
    for i in range(1200):
        a = t.SelectArm(0)
        t.UpdateArm(0, a, 1.0 + (a == 8))
    t.Reset(True, str(tmp_path / "bandit_log"))
    files = list((tmp_path / "bandit_log").iterdir())
    assert len(files) == 1 and files[0].read_text().count("\n") >= 3
    assert t.GetBanditSize() == 0


def test_header_is_plain_c(tmp_path):
    """include/cc_api.h is the drop-in boundary: it must compile as C99 (no C++, no CUDA or torch types) and as C++17."""
    import subprocess

    src_c = tmp_path / "use_api.c"
    src_c.write_text('#include "cc_api.h"\nint main(void) { cc_probe_result r; cc_probe_payload_result p; (void) r; (void) p; return cc_api_version() == CC_API_VERSION ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src_c)])
    src_cpp = tmp_path / "use_api.cpp"
    src_cpp.write_text(src_c.read_text())
    subprocess.check_call(["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src_cpp)])
    code = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S).lower()  # comments may name what a handle wraps
    code = code.replace("cc_err_cuda", "")  # the status code, not a type
    assert "cuda" not in code and "torch" not in code and "#include <std" in code


def test_library_holds_sm_100a_code_only():
    """no multi-arch fat binary, no PTX for a JIT to retarget: the product is written for B200 (sm_100a) alone, and the
    two hot kernels use what the design says they use (TMA bulk copies in the scatter, 256-bit sector loads in the probe)"""
    import shutil
    import subprocess

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    lib = os.path.join(ROOT, "chunk-compaction-in-vectorized-execution-simd_b200", "libccb200.so")
    elfs = subprocess.check_output([cuobjdump, "--list-elf", lib], text=True).split()
    cubins = [e for e in elfs if e.endswith(".cubin")]
    assert cubins and all(".sm_100a." in c for c in cubins), cubins
    sass = subprocess.check_output([cuobjdump, "-sass", lib], text=True)
    assert "UBLKCP" in sass          # cp.async.bulk (TMA) key ring of partition_scatter_kernel
    assert "LDG.E.ENL2.256" in sass  # one-request sector loads of the probe's tail walk
    assert "SYNCS" in sass           # mbarrier
