"""ctypes binding of libccb200.so (the C ABI declared in include/cc_api.h).

The library is built in-tree by `__graft_entry__.build()` / `make -C csrc`.
There is no fallback: if the shared object is missing or a call fails, an
exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CCB_LIB_PATH") or os.path.join(PKG_DIR, "libccb200.so")  # override: A/B builds of the same ABI

CC_HT_LP = 0
CC_HT_CHAIN = 1
CC_BUILD_ORDERED = 0
CC_BUILD_UNORDERED = 1
CC_CHAIN_WIDTH = 512
CC_MAX_JOINS = 8

CC_OK = 0
CC_ERR_NO_DEVICE = -2


class CCError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libccb200 error {code}: {msg}")
        self.code = code


class DeviceInfo(C.Structure):
    _fields_ = [
        ("device", C.c_int),
        ("sm_major", C.c_int),
        ("sm_minor", C.c_int),
        ("sm_count", C.c_int),
        ("l2_bytes", C.c_size_t),
        ("total_mem", C.c_size_t),
        ("free_mem", C.c_size_t),
        ("name", C.c_char * 128),
    ]


class HtInfo(C.Structure):
    _fields_ = [
        ("kind", C.c_int),
        ("n_keys", C.c_size_t),
        ("n_slots", C.c_size_t),
        ("bytes", C.c_size_t),
        ("has_duplicates", C.c_int),
        ("max_chain", C.c_size_t),
    ]


class ProbeResult(C.Structure):
    _fields_ = [
        ("n_matches", C.c_uint64),
        ("key_sum", C.c_uint64),
        ("payload_sum", C.c_uint64),
        ("overflow", C.c_uint64),
    ]


CC_MAX_PAYLOAD_COLS = 4


class ProbePayloadResult(C.Structure):
    _fields_ = [
        ("n_matches", C.c_uint64),
        ("key_sum", C.c_uint64),
        ("payload_sum", C.c_uint64),
        ("overflow", C.c_uint64),
        ("col_sum", C.c_uint64 * CC_MAX_PAYLOAD_COLS),
    ]


class ChainResult(C.Structure):
    _fields_ = [
        ("n_tuples", C.c_uint64),
        ("digest", C.c_uint64),
        ("colsum", C.c_uint64 * (3 * CC_MAX_JOINS)),
        ("level_in", C.c_uint64 * CC_MAX_JOINS),
        ("level_steps", C.c_uint64 * CC_MAX_JOINS),
        ("level_lanes", C.c_uint64 * CC_MAX_JOINS),
        ("overflow", C.c_uint64),
        ("device_ns", C.c_uint64),
        ("reserved", C.c_uint64 * 3),
    ]


CC_DENSITY_BINS = 8


class ChainTelemetry(C.Structure):
    _fields_ = [("probe_rows_hist", (C.c_uint64 * CC_DENSITY_BINS) * CC_MAX_JOINS), ("round_lanes_hist", (C.c_uint64 * CC_DENSITY_BINS) * CC_MAX_JOINS)]


# every symbol include/cc_api.h declares: (restype, argtypes)
_vp, _sz, _u64, _int = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
_pvp = C.POINTER(C.c_void_p)
SIGNATURES = {
    "cc_api_version": (_int, []),
    "cc_last_error": (C.c_char_p, []),
    "cc_device_init": (_int, [_int]),
    "cc_device_get_info": (_int, [C.POINTER(DeviceInfo)]),
    "cc_malloc": (_int, [_pvp, _sz]),
    "cc_free": (_int, [_vp]),
    "cc_host_alloc": (_int, [_pvp, _sz]),
    "cc_host_free": (_int, [_vp]),
    "cc_memcpy_h2d": (_int, [_vp, _vp, _sz, _vp]),
    "cc_memcpy_d2h": (_int, [_vp, _vp, _sz, _vp]),
    "cc_memcpy_d2d": (_int, [_vp, _vp, _sz, _vp]),
    "cc_memset": (_int, [_vp, _int, _sz, _vp]),
    "cc_stream_create": (_int, [_pvp]),
    "cc_stream_destroy": (_int, [_vp]),
    "cc_stream_sync": (_int, [_vp]),
    "cc_scratch_set_retention": (_int, [_u64]),
    "cc_scratch_release": (_int, []),
    "cc_launch_count": (_u64, []),
    "cc_hash_u64": (_int, [_vp, _vp, _sz, _vp]),
    "cc_gen_build_keys": (_int, [_vp, _sz, _sz, _vp]),
    "cc_gen_build_keys_range": (_int, [_vp, _sz, _sz, _sz, _sz, _vp]),
    "cc_gen_keys_counter": (_int, [_vp, _sz, _u64, _u64, _u64, _vp]),
    "cc_ht_build": (_int, [_pvp, _int, _vp, _sz, _int, _vp]),
    "cc_ht_build_sized": (_int, [_pvp, _int, _vp, _sz, _sz, _int, _vp]),
    "cc_ht_build_reference": (_int, [_pvp, _int, _sz, _sz, _vp]),
    "cc_ht_import_lp": (_int, [_pvp, _vp, _sz, _sz, _vp]),
    "cc_ht_attach_payload": (_int, [_vp, _vp, _pvp, _sz, _vp]),
    "cc_ht_build_reference_payload": (_int, [_pvp, _int, _sz, _sz, _vp]),
    "cc_ht_payload_cols": (_sz, [_vp]),
    "cc_ht_export_payload": (_int, [_vp, _pvp]),
    "cc_ht_get_info": (_int, [_vp, C.POINTER(HtInfo)]),
    "cc_ht_export_lp": (_int, [_vp, _vp]),
    "cc_ht_export_chain": (_int, [_vp, _vp, _vp, _vp]),
    "cc_ht_destroy": (_int, [_vp]),
    "cc_probe_chunk": (_int, [_vp, _vp, _sz, _vp, _sz, _pvp, _vp]),
    "cc_scan_has_next": (_int, [_vp]),
    "cc_scan_active": (_sz, [_vp]),
    "cc_scan_next": (_int, [_vp, _int, _vp, _vp, _vp, _vp, C.POINTER(_sz), _vp]),
    "cc_scan_destroy": (_int, [_vp]),
    "cc_chunk_append": (_int, [_pvp, _sz, _pvp, _vp, _sz, _sz, _sz, _vp]),
    "cc_sel_compose": (_int, [_vp, _vp, _vp, _sz, _vp]),
    "cc_sel_identity": (_int, [_vp, _sz, _vp]),
    "cc_rows_to_columns": (_int, [_vp, _sz, _sz, _pvp, _vp]),
    "cc_columns_to_rows": (_int, [_pvp, _vp, _sz, _sz, _vp, _vp]),
    "cc_probe_batch": (_int, [_vp, _vp, _sz, _vp, _vp, _vp, _sz, _vp, _vp]),
    "cc_probe_batch_payload": (_int, [_vp, _vp, _sz, _vp, _vp, _pvp, _sz, _vp, _sz, _vp, _vp]),
    "cc_probe_batch_segmented": (_int, [_vp, _vp, _int, _sz, _vp, _vp, _vp, _sz, _vp, _vp]),
    "cc_probe_stream_begin": (_int, [_pvp, _vp, _sz, _vp, _vp, _sz, _vp, _vp]),
    "cc_probe_stream_add": (_int, [_vp, _vp, _sz, _int, _sz, _vp, _vp]),
    "cc_probe_stream_finish": (_int, [_vp, _vp]),
    "cc_probe_set_strategy": (_int, [_int, _sz]),
    "cc_probe_set_cache_mode": (_int, [_int, _int]),
    "cc_probe_set_profiling": (_int, [_int]),
    "cc_probe_last_phase_ms": (_int, [_vp]),
    "cc_probe_batch_host": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, C.POINTER(ProbeResult), _vp]),
    "cc_probe_host_release": (_int, []),
    "cc_compactor_create": (_int, [_pvp, _sz, _sz, _sz]),
    "cc_compactor_set_threshold": (_int, [_vp, _sz]),
    "cc_compactor_get_threshold": (_sz, [_vp]),
    "cc_compactor_compact": (_int, [_vp, _pvp, _vp, C.POINTER(_sz), _pvp, _pvp, _vp]),
    "cc_compactor_flush": (_int, [_vp, _pvp, _pvp, C.POINTER(_sz), _vp]),
    "cc_compactor_destroy": (_int, [_vp]),
    "cc_chain_execute": (_int, [_pvp, _sz, _pvp, _sz, _vp, _pvp, _sz, _vp, _vp]),
    "cc_chain_execute_ex": (_int, [_pvp, _sz, _pvp, _sz, _vp, _pvp, _sz, _vp, _vp, _vp]),
    "cc_chain_telemetry_csv": (_int, [_vp, _sz, C.c_char_p]),
    "cc_chain_execute_tuned": (_int, [_pvp, _sz, _pvp, _sz, _sz, _vp, _sz, _pvp, _sz, C.POINTER(ChainResult), _vp]),
    "cc_tuner_create": (_int, [_pvp]),
    "cc_tuner_initialize": (_int, [_vp, _sz, _vp, _sz]),
    "cc_tuner_select_arm": (_int, [_vp, _sz, C.POINTER(_sz)]),
    "cc_tuner_update_arm": (_int, [_vp, _sz, _sz, C.c_double]),
    "cc_tuner_get_id": (C.c_int64, [_vp, _sz]),
    "cc_tuner_bandit_size": (_sz, [_vp]),
    "cc_tuner_state": (_int, [_vp, _sz, _vp, _vp, _sz]),
    "cc_tuner_reset": (_int, [_vp, _int, C.c_char_p]),
    "cc_tuner_destroy": (_int, [_vp]),
    "cc_partition_count": (_int, [_vp, _sz, _int, _vp, _vp]),
    "cc_partition_scatter": (_int, [_vp, _sz, _int, _vp, _vp, _vp, _vp]),
    "cc_partition_scatter_peers": (_int, [_vp, _sz, _int, _vp, _vp, _pvp, _vp]),
    "cc_partition_single": (_int, [_vp, _sz, _int, _sz, _vp, _vp, _vp, _int, _vp, _vp]),
    "cc_partition_set_peer_blocks": (_int, [_int]),
    "cc_ipc_export": (_int, [_vp, _vp]),
    "cc_ipc_open": (_int, [_vp, _pvp]),
    "cc_ipc_close": (_int, [_vp]),
    "cc_peer_copy_sm": (_int, [_vp, _vp, _sz, _int, _vp]),
}

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t)
BARRIER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p)


class Comm(C.Structure):
    """cc_comm: the caller's control plane for cc_pjoin_create / cc_pjoin_destroy"""
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("allgather", ALLGATHER_FN), ("barrier", BARRIER_FN), ("user", C.c_void_p)]


SIGNATURES.update({
    "cc_pjoin_create": (_int, [_pvp, C.POINTER(Comm), _int, _vp, _sz, _sz, _int, _vp]),
    "cc_pjoin_probe": (_int, [_vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp]),
    "cc_pjoin_probe_begin": (_int, [_vp, _vp, _sz, _vp]),
    "cc_pjoin_probe_end": (_int, [_vp, _vp, _vp, _sz, _vp, _vp]),
    "cc_pjoin_table": (_int, [_vp, _pvp]),
    "cc_pjoin_destroy": (_int, [_vp]),
})

_lib = None


def load() -> C.CDLL:
    """Load libccb200.so; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C chunk-compaction-in-vectorized-execution-simd_b200/csrc`). There is no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != CC_OK:
        raise CCError(rc, load().cc_last_error().decode(errors="replace"))
