"""B200-native hash-join probe + chunk-compaction engine -- Python host layer.

A thin mirror of the reference's operator surface (namespace simd_compaction:
HashTable / LPHashTable / ScanStructure / DataChunk / DataCollection /
NaiveCompactor / CompactTuner) on top of the C ABI in
include/cc_api.h.  PyTorch is used for device memory and streams only; all
compute is in libccb200.so (hand-written sm_100a kernels).  No CPU fallback.

The package directory name contains '-', so import it with
    importlib.import_module("chunk-compaction-in-vectorized-execution-simd_b200")
(tests/conftest.py and __graft_entry__.py also alias it as `ccb200`).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L
from ._lib import (CC_BUILD_ORDERED, CC_BUILD_UNORDERED, CC_CHAIN_WIDTH, CC_HT_CHAIN, CC_HT_LP, CC_MAX_JOINS, CCError,
                   ChainResult, DeviceInfo, HtInfo, ProbeResult)

# ---- reference globals (base.h:37-51) ------------------------------------------
kBlockSize = 2048
DEFAULT_ARMS = (0, 32, 64, 128, 256, 384, 512, 768, 1024)  # negative_feedback.hpp:172

_initialised = False


def lib():
    return L.load()


def init(device: int = 0) -> DeviceInfo:
    """cudaSetDevice + capability check; raises CCError when no sm_100 device exists."""
    global _initialised
    L.check(lib().cc_device_init(device))
    torch.cuda.set_device(device)
    _initialised = True
    info = DeviceInfo()
    L.check(lib().cc_device_get_info(C.byref(info)))
    return info


def _ensure():
    if not _initialised:
        init(torch.cuda.current_device() if torch.cuda.is_available() else 0)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous()
    return t.data_ptr()


def launch_count() -> int:
    return int(lib().cc_launch_count())


def _i64(x, device="cuda") -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.int64)).to(device)


def _u32(x, device="cuda") -> torch.Tensor:
    """uint32 payload carried in an int32 tensor (torch has no full uint32 support)."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=torch.int32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.uint32).view(np.int32)).to(device)


def to_u32_numpy(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().view(np.uint32)


# ---- utility kernels ------------------------------------------------------------
def murmurhash64(x) -> torch.Tensor:
    """hash_functions.h:8-16 on the device (uint64 carried as int64)."""
    _ensure()
    x = _i64(x)
    out = torch.empty_like(x)
    L.check(lib().cc_hash_u64(_ptr(x), _ptr(out), x.numel(), _stream()))
    return out


def gen_build_keys(n: int, chunk_factor: int) -> torch.Tensor:
    _ensure()
    out = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")[:n]
    L.check(lib().cc_gen_build_keys(_ptr(out), n, chunk_factor, _stream()))
    return out


def gen_keys_counter(n: int, seed: int, mask: int, first: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _ensure()
    if out is None:
        out = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")[:n]
    L.check(lib().cc_gen_keys_counter(_ptr(out), n, seed, first, mask, _stream()))
    return out


def set_probe_strategy(strategy: int = 0, slice_bytes: int = 0) -> None:
    """0 auto, 1 direct, 2 partitioned (cc_probe_set_strategy)."""
    L.check(lib().cc_probe_set_strategy(strategy, slice_bytes))


def set_probe_profiling(enable: bool) -> None:
    L.check(lib().cc_probe_set_profiling(int(enable)))


def probe_last_phase_ms():
    """(partition histogram, partition scatter, probe kernel) of the last cc_probe_batch, in ms."""
    ms = (C.c_float * 3)()
    L.check(lib().cc_probe_last_phase_ms(ms))
    return [float(x) for x in ms]


def set_probe_cache_mode(mode_direct: int = 0, mode_partitioned: int = 2) -> None:
    L.check(lib().cc_probe_set_cache_mode(mode_direct, mode_partitioned))


# ---- data model (base.h:54-100) ---------------------------------------------------
class Vector:
    """base.h:59-76: one column of kBlockSize int64 values (device resident, shared by reference)."""

    def __init__(self, block: Optional[int] = None, data: Optional[torch.Tensor] = None):
        _ensure()
        self.data_ = data if data is not None else torch.zeros(block or kBlockSize, dtype=torch.int64, device="cuda")

    def Reference(self, other: "Vector") -> None:  # base.cpp:5-8
        self.data_ = other.data_

    def Data(self) -> torch.Tensor:
        return self.data_


class DataChunk:
    """base.h:79-100: count_ + columns + selection vector, all on the device."""

    def __init__(self, n_cols: int, block: Optional[int] = None):
        _ensure()
        self.block = block or kBlockSize
        self.count_ = 0
        self.data_: List[Vector] = [Vector(self.block) for _ in range(n_cols)]
        self.selection_vector_ = torch.arange(self.block, dtype=torch.int32, device="cuda")

    def Reset(self) -> None:  # base.h:96-99
        self.count_ = 0
        L.check(lib().cc_sel_identity(_ptr(self.selection_vector_), self.block, _stream()))

    def Slice(self, other: "DataChunk", selection_vector: torch.Tensor, count: int) -> None:  # base.cpp:37-47
        assert len(other.data_) <= len(self.data_)
        self.count_ = count
        for c in range(len(other.data_)):
            self.data_[c].Reference(other.data_[c])
        L.check(lib().cc_sel_compose(_ptr(self.selection_vector_), _ptr(other.selection_vector_), _ptr(selection_vector), count, _stream()))

    SIMDSlice = Slice  # base.cpp:49-68 computes the same thing

    def Append(self, chunk: "DataChunk", num: int, offset: int = 0) -> None:  # base.cpp:15-27
        assert len(self.data_) == len(chunk.data_) and self.count_ + num <= self.block
        n = len(self.data_)
        dst = (C.c_void_p * n)(*[v.data_.data_ptr() for v in self.data_])
        src = (C.c_void_p * n)(*[v.data_.data_ptr() for v in chunk.data_])
        L.check(lib().cc_chunk_append(dst, self.count_, src, _ptr(chunk.selection_vector_), num, offset, n, _stream()))
        self.count_ += num

    def AppendTuple(self, tup: Sequence[int]) -> None:  # base.cpp:29-35 (host convenience, slow)
        for i, v in enumerate(tup):
            self.data_[i].data_[self.count_] = int(v)
        self.count_ += 1

    def rows(self) -> np.ndarray:
        """Materialise the logical rows (through the selection vector) on the host."""
        sel = self.selection_vector_[: self.count_].long()
        cols = [v.data_[sel] for v in self.data_]
        return torch.stack(cols, dim=1).cpu().numpy() if cols else np.empty((self.count_, 0), dtype=np.int64)


class DataCollection:
    """data_collection.h:15-33.  The reference keeps row-major host rows; here the table is
    device-columnar and the row<->column transposes run on the GPU (FetchChunk / AppendChunk)."""

    def __init__(self, n_cols: int):
        _ensure()
        self.n_cols = n_cols
        self._rows: List[torch.Tensor] = []  # list of row-major device blocks
        self.n_tuples_ = 0

    def AppendTuple(self, tup: Sequence[int]) -> None:
        self.AppendRows(np.asarray([tup], dtype=np.int64))

    def AppendRows(self, rows) -> None:
        rows = _i64(rows).reshape(-1, self.n_cols)
        self._rows.append(rows)
        self.n_tuples_ += rows.shape[0]

    def _flat(self) -> torch.Tensor:
        if len(self._rows) > 1:
            self._rows = [torch.cat(self._rows, dim=0)]
        return self._rows[0] if self._rows else torch.empty((0, self.n_cols), dtype=torch.int64, device="cuda")

    def AppendChunk(self, chunk: DataChunk) -> None:  # data_collection.cpp:10-21
        assert len(chunk.data_) == self.n_cols
        out = torch.empty((max(chunk.count_, 1), self.n_cols), dtype=torch.int64, device="cuda")[: chunk.count_]
        cols = (C.c_void_p * self.n_cols)(*[v.data_.data_ptr() for v in chunk.data_])
        L.check(lib().cc_columns_to_rows(cols, _ptr(chunk.selection_vector_), chunk.count_, self.n_cols, _ptr(out), _stream()))
        self._rows.append(out)
        self.n_tuples_ += chunk.count_

    def FetchChunk(self, start: int, end: int, block: Optional[int] = None) -> DataChunk:  # data_collection.cpp:23-27
        chunk = DataChunk(self.n_cols, block)
        rows = self._flat()[start:end].contiguous()
        cols = (C.c_void_p * self.n_cols)(*[v.data_.data_ptr() for v in chunk.data_])
        L.check(lib().cc_rows_to_columns(_ptr(rows), rows.shape[0], self.n_cols, cols, _stream()))
        chunk.count_ = rows.shape[0]
        return chunk

    def NumTuples(self) -> int:
        return self.n_tuples_

    def numpy(self) -> np.ndarray:
        return self._flat().cpu().numpy()

    def Print(self, n_tuple: int) -> None:  # data_collection.cpp:29-45
        for row in self.numpy()[: min(n_tuple, self.n_tuples_)]:
            print("".join(f"{int(v)}, " for v in row))


# ---- hash tables ------------------------------------------------------------------
class ScanStructure:
    """chaining_ht.h:29-84 / linear_probing_ht.h:24-55 on the device."""

    def __init__(self, handle: int, table: "_TableBase"):
        self._h = handle
        self._table = table

    def HasNext(self) -> bool:
        return bool(lib().cc_scan_has_next(self._h))

    def active(self) -> int:
        return int(lib().cc_scan_active(self._h))

    def _next(self, in_one: int, join_key: Vector, inp: DataChunk, result: DataChunk) -> int:
        n_in = len(inp.data_)
        assert len(result.data_) >= n_in + 2
        # Slice shares the LHS columns (base.cpp:40); the selection vector is composed in-kernel
        for c in range(n_in):
            result.data_[c].Reference(inp.data_[c])
        cnt = C.c_size_t(0)
        L.check(lib().cc_scan_next(self._h, in_one, _ptr(join_key.data_), _ptr(inp.selection_vector_), _ptr(result.selection_vector_),
                                   _ptr(result.data_[n_in + 1].data_), C.byref(cnt), _stream()))
        result.count_ = cnt.value
        return cnt.value

    def Next(self, join_key: Vector, inp: DataChunk, result: DataChunk) -> int:  # chaining_ht.cpp:60-80
        return self._next(0, join_key, inp, result)

    def InOneNext(self, join_key: Vector, inp: DataChunk, result: DataChunk) -> int:  # chaining_ht.cpp:138-173
        return self._next(1, join_key, inp, result)

    SIMDNext = Next  # the AVX-512 twins compute the same results
    SIMDInOneNext = InOneNext

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cc_scan_destroy(self._h)
            self._h = None


class _TableBase:
    kind = -1

    def __init__(self, n_rhs_tuples: Optional[int] = None, chunk_factor: int = 1, *, keys=None, flags: int = CC_BUILD_ORDERED,
                 payload=None, keep_payload: bool = False, n_slots: int = 0):
        """HashTable(n_rhs_tuples, chunk_factor) like the reference (chaining_ht.h:88), or keys=... for explicit
        build keys (the reference has no external build-input API).
        Payload columns: payload=[col, ...] with keys=..., or keep_payload=True to keep the column the
        reference generates and drops (payload of build row i = i + 10000000, chaining_ht.cpp:21)."""
        _ensure()
        h = C.c_void_p()
        if keys is not None:
            k = _i64(keys)
            L.check(lib().cc_ht_build_sized(C.byref(h), self.kind, _ptr(k) if k.numel() else None, k.numel(), int(n_slots), flags, _stream()))
            self._h = h.value
            if payload is not None:
                self.attach_payload(k, payload)
        elif keep_payload:
            L.check(lib().cc_ht_build_reference_payload(C.byref(h), self.kind, int(n_rhs_tuples), int(chunk_factor), _stream()))
            self._h = h.value
        else:
            L.check(lib().cc_ht_build_reference(C.byref(h), self.kind, int(n_rhs_tuples), int(chunk_factor), _stream()))
            self._h = h.value

    def attach_payload(self, build_keys, payload_cols) -> None:
        """cc_ht_attach_payload: keep int64 payload columns (one value per build row) next to the keys."""
        k = _i64(build_keys)
        cols = [_i64(c) for c in payload_cols]
        assert all(c.numel() == k.numel() for c in cols)
        arr = (C.c_void_p * len(cols))(*[_ptr(c) if c.numel() else None for c in cols])
        L.check(lib().cc_ht_attach_payload(self._h, _ptr(k) if k.numel() else None, arr, len(cols), _stream()))

    def payload_cols(self) -> int:
        return int(lib().cc_ht_payload_cols(self._h))

    def export_payload(self) -> List[np.ndarray]:
        """payload columns in table order (LP: one row per slot; chain: chain order)"""
        i = self.info()
        rows = i.n_slots if self.kind == CC_HT_LP else i.n_keys
        cols = [np.zeros(max(rows, 1), dtype=np.int64) for _ in range(self.payload_cols())]
        arr = (C.c_void_p * max(len(cols), 1))(*[c.ctypes.data for c in cols])
        L.check(lib().cc_ht_export_payload(self._h, arr))
        return [c[:rows] for c in cols]

    def probe_batch_payload(self, keys: torch.Tensor, *, capacity: Optional[int] = None, n_out_cols: Optional[int] = None,
                            materialize: bool = True, rowid: bool = False, sync: bool = True) -> dict:
        """cc_probe_batch_payload: result rows (probe key, matched build key, payload columns of the matched build row)."""
        n = keys.numel()
        cap = capacity if capacity is not None else n
        npay = self.payload_cols()
        nout = npay if n_out_cols is None else n_out_cols
        new = lambda: torch.empty(max(cap, 1), dtype=torch.int64, device="cuda")
        out_key = new() if materialize else None
        out_build = new() if materialize else None
        out_cols = [new() for _ in range(nout)] if materialize else []
        out_rowid = new() if rowid else None
        result = torch.zeros(4 + L.CC_MAX_PAYLOAD_COLS, dtype=torch.int64, device="cuda")
        arr = (C.c_void_p * max(len(out_cols), 1))(*[_ptr(c) for c in out_cols])
        L.check(lib().cc_probe_batch_payload(self._h, _ptr(keys) if n else None, n, _ptr(out_key), _ptr(out_build), arr, len(out_cols),
                                             _ptr(out_rowid), cap if (materialize or rowid) else 0, _ptr(result), _stream()))
        out = {"result_tensor": result, "out_key": out_key, "out_build_key": out_build, "out_cols": out_cols, "out_rowid": out_rowid}
        if sync:
            r = result.cpu().numpy().view(np.uint64)
            out.update(n_matches=int(r[0]), key_sum=int(r[1]), payload_sum=int(r[2]), overflow=int(r[3]),
                       col_sum=[int(x) for x in r[4:4 + npay]])
        return out

    @classmethod
    def import_slots(cls, slots: np.ndarray, n_keys: int) -> "_TableBase":
        _ensure()
        assert cls.kind == CC_HT_LP
        self = cls.__new__(cls)
        h = C.c_void_p()
        s = np.ascontiguousarray(slots, dtype=np.int64)
        L.check(lib().cc_ht_import_lp(C.byref(h), s.ctypes.data, s.size, n_keys, _stream()))
        self._h = h.value
        return self

    def info(self) -> HtInfo:
        i = HtInfo()
        L.check(lib().cc_ht_get_info(self._h, C.byref(i)))
        return i

    def Probe(self, join_key: Vector, count: int, sel_vec: torch.Tensor, block: Optional[int] = None) -> ScanStructure:
        h = C.c_void_p()
        block = block or join_key.data_.numel()
        L.check(lib().cc_probe_chunk(self._h, _ptr(join_key.data_), count, _ptr(sel_vec), block, C.byref(h), _stream()))
        return ScanStructure(h.value, self)

    SIMDProbe = Probe

    def probe_batch(self, keys: torch.Tensor, *, materialize: bool = True, rowid: bool = False, capacity: Optional[int] = None,
                    out_key: Optional[torch.Tensor] = None, out_payload: Optional[torch.Tensor] = None,
                    result: Optional[torch.Tensor] = None, sync: bool = True) -> dict:
        """One-launch probe of a whole key column with dense (compacted) output (cc_probe_batch)."""
        n = keys.numel()
        cap = capacity if capacity is not None else n
        if materialize:
            if out_key is None:
                out_key = torch.empty(max(cap, 1), dtype=torch.int64, device="cuda")
            if out_payload is None:
                out_payload = torch.empty(max(cap, 1), dtype=torch.int64, device="cuda")
        out_rowid = torch.empty(max(cap, 1), dtype=torch.int64, device="cuda") if rowid else None
        if result is None:
            result = torch.zeros(4, dtype=torch.int64, device="cuda")
        L.check(lib().cc_probe_batch(self._h, _ptr(keys) if n else None, n, _ptr(out_key), _ptr(out_payload), _ptr(out_rowid),
                                     cap if (materialize or rowid) else 0, _ptr(result), _stream()))
        out = {"result_tensor": result, "out_key": out_key, "out_payload": out_payload, "out_rowid": out_rowid}
        if sync:
            r = result.cpu().numpy().view(np.uint64)
            out.update(n_matches=int(r[0]), key_sum=int(r[1]), payload_sum=int(r[2]), overflow=int(r[3]))
        return out

    def probe_batch_segmented(self, keys: torch.Tensor, n_segments: int, segment_capacity: int, segment_counts: torch.Tensor, *,
                              capacity: int, out_key: Optional[torch.Tensor] = None, out_payload: Optional[torch.Tensor] = None,
                              result: Optional[torch.Tensor] = None, sync: bool = True) -> dict:
        """cc_probe_batch_segmented: probe a segmented key column (segment s = keys[s * cap : s * cap + counts[s]], the counts
        stay on the device).  `keys` must span n_segments * segment_capacity rows."""
        assert keys.numel() >= n_segments * segment_capacity and segment_counts.dtype == torch.int64
        if result is None:
            result = torch.zeros(4, dtype=torch.int64, device="cuda")
        L.check(lib().cc_probe_batch_segmented(self._h, _ptr(keys), n_segments, segment_capacity, _ptr(segment_counts), _ptr(out_key),
                                               _ptr(out_payload), capacity if (out_key is not None or out_payload is not None) else 0,
                                               _ptr(result), _stream()))
        out = {"result_tensor": result, "out_key": out_key, "out_payload": out_payload}
        if sync:
            r = result.cpu().numpy().view(np.uint64)
            out.update(n_matches=int(r[0]), key_sum=int(r[1]), payload_sum=int(r[2]), overflow=int(r[3]))
        return out

    def probe_stream(self, n_expected: int, *, capacity: int, out_key: Optional[torch.Tensor] = None,
                     out_payload: Optional[torch.Tensor] = None, result: Optional[torch.Tensor] = None) -> "ProbeStream":
        """cc_probe_stream_begin: an incremental probe of this table (add pieces, then finish)."""
        return ProbeStream(self, n_expected, capacity, out_key, out_payload, result)

    def probe_batch_host(self, h_keys: np.ndarray, h_out_key: Optional[np.ndarray], h_out_payload: Optional[np.ndarray]) -> dict:
        """End-to-end probe with HOST buffers (cc_probe_batch_host): H2D + probe + D2H inside."""
        r = ProbeResult()
        cap = h_out_key.size if h_out_key is not None else (h_out_payload.size if h_out_payload is not None else 0)
        L.check(lib().cc_probe_batch_host(self._h, h_keys.ctypes.data, h_keys.size, h_out_key.ctypes.data if h_out_key is not None else None,
                                          h_out_payload.ctypes.data if h_out_payload is not None else None, cap, C.byref(r), None))
        return dict(n_matches=int(r.n_matches), key_sum=int(r.key_sum), payload_sum=int(r.payload_sum), overflow=int(r.overflow))

    def export(self):
        i = self.info()
        if self.kind == CC_HT_LP:
            s = np.empty(i.n_slots, dtype=np.int64)
            L.check(lib().cc_ht_export_lp(self._h, s.ctypes.data))
            return s
        b = np.empty(i.n_slots, dtype=np.uint32)
        c = np.empty(i.n_slots, dtype=np.uint32)
        k = np.empty(max(i.n_keys, 1), dtype=np.int64)
        L.check(lib().cc_ht_export_chain(self._h, b.ctypes.data, c.ctypes.data, k.ctypes.data))
        return b, c, k[: i.n_keys]

    def destroy(self) -> None:
        if getattr(self, "_h", None) and lib is not None:
            lib().cc_ht_destroy(self._h)
            self._h = None

    def __del__(self):
        self.destroy()


class ProbeStream:
    """Incremental probe (cc_probe_stream_*): pieces of the key column are added as they arrive, one probe result."""

    def __init__(self, table: "_TableBase", n_expected: int, capacity: int, out_key, out_payload, result):
        self.result = result if result is not None else torch.zeros(4, dtype=torch.int64, device="cuda")
        self.out_key, self.out_payload = out_key, out_payload
        h = C.c_void_p()
        L.check(lib().cc_probe_stream_begin(C.byref(h), table._h, n_expected, _ptr(out_key), _ptr(out_payload),
                                            capacity if (out_key is not None or out_payload is not None) else 0, _ptr(self.result), _stream()))
        self._h = h

    def add(self, keys: torch.Tensor) -> None:
        L.check(lib().cc_probe_stream_add(self._h, _ptr(keys) if keys.numel() else None, keys.numel(), 0, 0, None, _stream()))

    def add_segmented(self, keys: torch.Tensor, n_segments: int, segment_capacity: int, segment_counts: torch.Tensor) -> None:
        assert keys.numel() >= n_segments * segment_capacity and segment_counts.dtype == torch.int64
        L.check(lib().cc_probe_stream_add(self._h, _ptr(keys), 0, n_segments, segment_capacity, _ptr(segment_counts), _stream()))

    def finish(self, sync: bool = True) -> dict:
        h, self._h = self._h, None
        L.check(lib().cc_probe_stream_finish(h, _stream()))
        out = {"result_tensor": self.result, "out_key": self.out_key, "out_payload": self.out_payload}
        if sync:
            r = self.result.cpu().numpy().view(np.uint64)
            out.update(n_matches=int(r[0]), key_sum=int(r[1]), payload_sum=int(r[2]), overflow=int(r[3]))
        return out

    def __del__(self):
        if getattr(self, "_h", None) and lib is not None:
            try:
                lib().cc_probe_stream_finish(self._h, None)
            except Exception:
                pass
            self._h = None


class HashTable(_TableBase):
    """chaining_ht.h:86-101 (separate chaining)."""

    kind = CC_HT_CHAIN


class LPHashTable(_TableBase):
    """linear_probing_ht.h:56-71 (linear probing, -1 == empty)."""

    kind = CC_HT_LP


# ---- compactor ----------------------------------------------------------------------
class NaiveCompactor:
    """compactor.h:14-29.  `Compact(chunk)` returns the chunk to push downstream (count_ == 0 when the
    rows were buffered) -- the Python spelling of the reference's unique_ptr<DataChunk>& in/out swap."""

    def __init__(self, n_cols: int, block: Optional[int] = None, threshold: Optional[int] = None):
        _ensure()
        self.block = block or kBlockSize
        self.n_cols = n_cols
        h = C.c_void_p()
        L.check(lib().cc_compactor_create(C.byref(h), n_cols, self.block, self.block if threshold is None else threshold))
        self._h = h.value

    def SetThreshold(self, threshold: int) -> None:  # main.cpp:141
        L.check(lib().cc_compactor_set_threshold(self._h, threshold))

    def GetThreshold(self) -> int:  # main.cpp:166
        return int(lib().cc_compactor_get_threshold(self._h))

    def _wrap(self, cols, sel_ptr, count, src: Optional[DataChunk]) -> DataChunk:
        if src is not None and sel_ptr == src.selection_vector_.data_ptr():
            src.count_ = count
            return src
        out = DataChunk.__new__(DataChunk)
        out.block = self.block
        out.count_ = count
        out.data_ = [Vector(data=_wrap_ptr(cols[j], self.block, torch.int64)) for j in range(self.n_cols)]
        out.selection_vector_ = _wrap_ptr(sel_ptr, self.block, torch.int32)
        return out

    def Compact(self, chunk: DataChunk) -> DataChunk:  # compactor.cpp:5-41
        n = self.n_cols
        cols = (C.c_void_p * n)(*[v.data_.data_ptr() for v in chunk.data_])
        out_cols = (C.c_void_p * n)()
        out_sel = C.c_void_p()
        cnt = C.c_size_t(chunk.count_)
        L.check(lib().cc_compactor_compact(self._h, cols, _ptr(chunk.selection_vector_), C.byref(cnt), out_cols, C.byref(out_sel), _stream()))
        return self._wrap(out_cols, out_sel.value, cnt.value, chunk)

    def Flush(self) -> DataChunk:  # compactor.h:23
        n = self.n_cols
        out_cols = (C.c_void_p * n)()
        out_sel = C.c_void_p()
        cnt = C.c_size_t(0)
        L.check(lib().cc_compactor_flush(self._h, out_cols, C.byref(out_sel), C.byref(cnt), _stream()))
        return self._wrap(out_cols, out_sel.value, cnt.value, None)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cc_compactor_destroy(self._h)
            self._h = None


Compactor = NaiveCompactor  # setting.h:17-29 picks the alias at compile time; here the threshold decides
BinaryCompactor = NaiveCompactor
DynamicCompactor = NaiveCompactor


class _ForeignBuffer:
    """Exposes library-owned device memory through __cuda_array_interface__ (zero copy)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _wrap_ptr(ptr: int, n: int, dtype: torch.dtype) -> torch.Tensor:
    typestr = "<i8" if dtype == torch.int64 else "<i4"
    return torch.as_tensor(_ForeignBuffer(ptr, n, typestr), device="cuda")


# ---- compaction policy ------------------------------------------------------------------
class CompactTuner:
    """negative_feedback.hpp:165-260 (the reference's is a singleton; Get() returns a process-wide one)."""

    _instance: Optional["CompactTuner"] = None

    def __init__(self):
        h = C.c_void_p()
        L.check(lib().cc_tuner_create(C.byref(h)))
        self._h = h.value

    @classmethod
    def Get(cls) -> "CompactTuner":
        if cls._instance is None:
            cls._instance = cls()
        return cls._instance

    def Initialize(self, address: int, arms: Optional[Sequence[int]] = None) -> None:
        if arms is None:
            L.check(lib().cc_tuner_initialize(self._h, address, None, 0))
        else:
            a = (C.c_size_t * len(arms))(*arms)
            L.check(lib().cc_tuner_initialize(self._h, address, a, len(arms)))

    def SelectArm(self, idx: int) -> int:
        v = C.c_size_t(0)
        L.check(lib().cc_tuner_select_arm(self._h, idx, C.byref(v)))
        return v.value

    def UpdateArm(self, idx: int, arm_value: int, reward: float) -> None:
        L.check(lib().cc_tuner_update_arm(self._h, idx, arm_value, reward))

    def GetId(self, address: int) -> int:
        return int(lib().cc_tuner_get_id(self._h, address))

    def GetBanditSize(self) -> int:
        return int(lib().cc_tuner_bandit_size(self._h))

    def state(self, idx: int, n_arms: int = len(DEFAULT_ARMS)):
        r = np.zeros(n_arms, dtype=np.float64)
        s = np.zeros(n_arms, dtype=np.uint64)
        L.check(lib().cc_tuner_state(self._h, idx, r.ctypes.data, s.ctypes.data, n_arms))
        return r, s

    def Reset(self, enable_log: bool = False, log_dir: Optional[str] = None) -> None:
        L.check(lib().cc_tuner_reset(self._h, int(enable_log), log_dir.encode() if log_dir else None))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cc_tuner_destroy(self._h)
            self._h = None


# ---- fused join chain ---------------------------------------------------------------------
def chain_execute(tables: Sequence[_TableBase], lhs_cols: Sequence[torch.Tensor], thresholds: Optional[Sequence[int]] = None,
                  materialize: bool = False, capacity: int = 0, result: Optional[torch.Tensor] = None, sync: bool = True,
                  telemetry: Optional[torch.Tensor] = None) -> dict:
    """ExecutePipeline + FlushPipelineCache (main.cpp:119-191) for a whole LHS table in one kernel.
    telemetry: an int64 device tensor from new_chain_telemetry() -- the chunk-density histograms of this call are ADDED to it
    (cc_chain_execute_ex; parse with parse_chain_telemetry)."""
    _ensure()
    J = len(tables)
    assert len(lhs_cols) == J
    n_rows = lhs_cols[0].numel()
    tp = (C.c_void_p * J)(*[t._h for t in tables])
    cp = (C.c_void_p * J)(*[c.data_ptr() for c in lhs_cols])
    thr = None
    if thresholds is not None:
        thr_arr = (C.c_uint32 * J)(*[int(x) for x in thresholds])
        thr = C.cast(thr_arr, C.c_void_p)
    outs = None
    op = None
    if materialize:
        outs = [torch.empty(max(capacity, 1), dtype=torch.int64, device="cuda") for _ in range(3 * J)]
        op = (C.c_void_p * (3 * J))(*[o.data_ptr() for o in outs])
    if result is None:
        result = torch.zeros(C.sizeof(ChainResult) // 8, dtype=torch.int64, device="cuda")
    L.check(lib().cc_chain_execute_ex(tp, J, cp, n_rows, thr, op, capacity, _ptr(result), _ptr(telemetry), _stream()))
    out = {"result_tensor": result, "out_cols": outs}
    if sync:
        out.update(parse_chain_result(result, J))
    return out


def chain_execute_tuned(tables: Sequence[_TableBase], lhs_cols: Sequence[torch.Tensor], tuner: "CompactTuner", batch_rows: int,
                        first_bandit_id: int = 0, materialize: bool = False, capacity: int = 0) -> dict:
    """Dynamic compaction (main.cpp:137-167): thresholds chosen per batch by the tuner's bandits (cc_chain_execute_tuned)."""
    _ensure()
    J = len(tables)
    n_rows = lhs_cols[0].numel()
    tp = (C.c_void_p * J)(*[t._h for t in tables])
    cp = (C.c_void_p * J)(*[c.data_ptr() for c in lhs_cols])
    outs, op = None, None
    if materialize:
        outs = [torch.empty(max(capacity, 1), dtype=torch.int64, device="cuda") for _ in range(3 * J)]
        op = (C.c_void_p * (3 * J))(*[o.data_ptr() for o in outs])
    r = ChainResult()
    L.check(lib().cc_chain_execute_tuned(tp, J, cp, n_rows, batch_rows, tuner._h, first_bandit_id, op, capacity, C.byref(r), _stream()))
    out = _chain_result_dict(r, J)
    out["out_cols"] = outs
    return out


def new_chain_telemetry() -> torch.Tensor:
    """zeroed cc_chain_telemetry on the device (chunk-density histograms, the ZebraProfiler analogue of profiler.h:168-260)"""
    _ensure()
    return torch.zeros(C.sizeof(L.ChainTelemetry) // 8, dtype=torch.int64, device="cuda")


def parse_chain_telemetry(telemetry: torch.Tensor, J: int) -> dict:
    t = L.ChainTelemetry.from_buffer_copy(telemetry.cpu().numpy().tobytes())
    return {"probe_rows_hist": [[int(t.probe_rows_hist[l][q]) for q in range(L.CC_DENSITY_BINS)] for l in range(J)],
            "round_lanes_hist": [[int(t.round_lanes_hist[l][q]) for q in range(L.CC_DENSITY_BINS)] for l in range(J)], "_struct": t}


def chain_telemetry_csv(telemetry: torch.Tensor, J: int, path: str) -> None:
    """cc_chain_telemetry_csv: histogram, level, density_from, density_to, chunks"""
    t = L.ChainTelemetry.from_buffer_copy(telemetry.cpu().numpy().tobytes())
    L.check(lib().cc_chain_telemetry_csv(C.byref(t), J, path.encode()))


def parse_chain_result(result: torch.Tensor, J: int) -> dict:
    raw = result.cpu().numpy().tobytes()
    return _chain_result_dict(ChainResult.from_buffer_copy(raw), J)


def _chain_result_dict(r: ChainResult, J: int) -> dict:
    return dict(n_tuples=int(r.n_tuples), digest=int(r.digest), colsum=[int(r.colsum[i]) for i in range(3 * J)],
                level_in=[int(r.level_in[i]) for i in range(J)], level_steps=[int(r.level_steps[i]) for i in range(J)],
                level_lanes=[int(r.level_lanes[i]) for i in range(J)], overflow=int(r.overflow), device_ns=int(r.device_ns),
                probe_tuples=sum(int(r.level_in[i]) for i in range(J)))


# ---- multi-GPU partitioning -----------------------------------------------------------------
def partition_single(keys: torch.Tensor, log2_parts: int, region_capacity: int, out: Optional[torch.Tensor] = None,
                     counts: Optional[torch.Tensor] = None, overflow: Optional[torch.Tensor] = None, self_part: int = -1,
                     self_out_ptr: Optional[int] = None):
    """cc_partition_single: single-pass hash partition into fixed regions of `region_capacity` rows.
    Returns (out[P * region_capacity], counts[P] int64 device, overflow int32[1] device); nothing is synchronised.
    self_part / self_out_ptr: that partition is written to the buffer at self_out_ptr (same region offset) instead."""
    _ensure()
    P = 1 << log2_parts
    n = keys.numel()
    if out is None:
        out = torch.empty(P * region_capacity, dtype=torch.int64, device="cuda")
    if counts is None:
        counts = torch.zeros(P, dtype=torch.int64, device="cuda")
    if overflow is None:
        overflow = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.check(lib().cc_partition_single(_ptr(keys) if n else None, n, log2_parts, region_capacity, _ptr(counts), _ptr(overflow), _ptr(out),
                                      self_part if self_out_ptr else -1, self_out_ptr, _stream()))
    return out, counts, overflow


def partition_keys(keys: torch.Tensor, log2_parts: int):
    """Hash-partition a key column into 2^log2_parts contiguous segments.
    Returns (partitioned keys, counts[P] as a host numpy array, offsets[P])."""
    _ensure()
    P = 1 << log2_parts
    n = keys.numel()
    counts = torch.zeros(P, dtype=torch.int64, device="cuda")
    L.check(lib().cc_partition_count(_ptr(keys) if n else None, n, log2_parts, _ptr(counts), _stream()))
    offsets = torch.cumsum(counts, 0) - counts
    cursors = torch.zeros(P, dtype=torch.int64, device="cuda")
    out = torch.empty(max(n, 1), dtype=torch.int64, device="cuda")[:n]
    L.check(lib().cc_partition_scatter(_ptr(keys) if n else None, n, log2_parts, _ptr(offsets), _ptr(cursors), _ptr(out) if n else None, _stream()))
    return out, counts.cpu().numpy(), offsets.cpu().numpy()
