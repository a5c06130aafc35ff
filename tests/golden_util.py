"""Readers for the committed golden fixtures (tests/golden/, produced by make_golden.py)."""
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_index() -> dict:
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        return json.load(f)


def load_i64(name: str) -> np.ndarray:
    return np.fromfile(os.path.join(GOLDEN_DIR, name), dtype=np.int64)


def load_nextdump(name: str):
    """-> list (per input chunk) of list (per Next call) of (positions u32[], payloads i64[])."""
    raw = np.fromfile(os.path.join(GOLDEN_DIR, name), dtype=np.uint8).tobytes()
    off = 0
    chunks, cur = [], []
    while off < len(raw):
        cnt = int(np.frombuffer(raw, dtype=np.uint32, count=1, offset=off)[0])
        off += 4
        if cnt == 0xFFFFFFFF:
            chunks.append(cur)
            cur = []
            continue
        rec = np.frombuffer(raw, dtype=np.dtype([("pos", "<u4"), ("payload", "<i8")]), count=cnt, offset=off)
        off += cnt * 12
        cur.append((rec["pos"].copy(), rec["payload"].copy()))
    return chunks


def sort_rows(t: np.ndarray) -> np.ndarray:
    if t.shape[0] == 0:
        return t
    return t[np.lexsort(t.T[::-1])]
