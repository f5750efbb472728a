"""ctypes binding of oracle/liboracle.so (ORACLE: test infrastructure only).

Mirrors the two reference classes at the level the parity tests need:
``Oracle.find(haystack_bytes, overlapping)`` is the drain of the reference's
iterator (src/lib.rs:42-68, 238-248, 433) and ``Oracle.find_str`` adds the
byte->code-point mapping of src/lib.rs:73-88,240-246.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

MATCHKIND_IDS = {"Standard": 0, "LeftmostFirst": 1, "LeftmostLongest": 2}


def build_oracle(force: bool = False) -> str:
    """Compile liboracle.so with gcc if missing or stale."""
    src = os.path.join(_HERE, "ac_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "liboracle.so"])
    return _LIB_PATH


_lib = None


def _load():
    global _lib
    if _lib is None:
        build_oracle()
        lib = ctypes.CDLL(_LIB_PATH)
        c = ctypes
        lib.orc_build.restype = c.c_void_p
        lib.orc_build.argtypes = [c.c_void_p, c.c_void_p, c.c_uint64, c.c_int, c.c_char_p, c.c_size_t]
        lib.orc_free.argtypes = [c.c_void_p]
        lib.orc_num_states.restype = c.c_uint32
        lib.orc_num_states.argtypes = [c.c_void_p]
        lib.orc_max_pattern_len.restype = c.c_uint32
        lib.orc_max_pattern_len.argtypes = [c.c_void_p]
        lib.orc_find_iter.restype = c.c_int64
        lib.orc_find_iter.argtypes = [c.c_void_p, c.c_void_p, c.c_uint64, c.c_int, c.c_int,
                                      c.c_void_p, c.c_void_p, c.c_void_p, c.c_uint64]
        lib.orc_byte_to_code_point.argtypes = [c.c_void_p, c.c_uint64, c.c_void_p]
        lib.orc_scan_batch.restype = c.c_uint64
        lib.orc_scan_batch.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_int,
                                       c.c_int, c.c_void_p, c.c_void_p, c.c_uint64]
        lib.orc_scan_batch_reps.restype = c.c_uint64
        lib.orc_scan_batch_reps.argtypes = lib.orc_scan_batch.argtypes + [c.c_int]
        _lib = lib
    return _lib


def _pack(patterns):
    pats = [bytes(p) for p in patterns]
    offs = np.zeros(len(pats) + 1, dtype=np.uint64)
    np.cumsum([len(p) for p in pats], out=offs[1:])
    blob = np.frombuffer(b"".join(pats) or b"\0", dtype=np.uint8).copy()
    return blob, offs


class Oracle:
    """CPU oracle automaton over byte patterns. kind: 0/1/2 or a MatchKind name."""

    def __init__(self, patterns, kind=0):
        lib = _load()
        if isinstance(kind, str):
            kind = MATCHKIND_IDS[kind]
        self.kind = int(kind)
        blob, offs = _pack(patterns)
        err = ctypes.create_string_buffer(256)
        self._h = lib.orc_build(blob.ctypes.data, offs.ctypes.data, len(offs) - 1, self.kind, err, 256)
        if not self._h:
            raise ValueError(err.value.decode())
        self.num_states = lib.orc_num_states(self._h)
        self.max_pattern_len = lib.orc_max_pattern_len(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_free(self._h)
            self._h = None

    def find(self, haystack: bytes, overlapping: bool = False, use_dfa: bool = True):
        """-> list[(pid, start, end)] in byte offsets (src/lib.rs:422-434)."""
        lib = _load()
        hay = np.frombuffer(bytes(haystack) or b"\0", dtype=np.uint8)
        n = len(haystack)
        cap = 1024
        while True:
            pid = np.empty(cap, dtype=np.uint32)
            st = np.empty(cap, dtype=np.uint64)
            en = np.empty(cap, dtype=np.uint64)
            got = lib.orc_find_iter(self._h, hay.ctypes.data, n, int(overlapping), int(use_dfa),
                                    pid.ctypes.data, st.ctypes.data, en.ctypes.data, cap)
            if got < 0:
                name = {1: "LeftmostFirst", 2: "LeftmostLongest"}[self.kind]
                raise ValueError(f"match kind {name} does not support overlapping searches")
            if got <= cap:
                return [(int(pid[i]), int(st[i]), int(en[i])) for i in range(got)]
            cap = int(got)

    def find_str(self, haystack: str, overlapping: bool = False, use_dfa: bool = True):
        """-> list[(pid, start, end)] in code points (src/lib.rs:229-249)."""
        lib = _load()
        raw = haystack.encode("utf-8")
        hay = np.frombuffer(raw or b"\0", dtype=np.uint8)
        b2c = np.empty(len(raw) + 1, dtype=np.uint64)
        lib.orc_byte_to_code_point(hay.ctypes.data, len(raw), b2c.ctypes.data)
        return [(p, int(b2c[s]), int(b2c[e])) for (p, s, e) in self.find(raw, overlapping, use_dfa)]

    def time_batch(self, data: np.ndarray, offsets: np.ndarray, overlapping=False, codepoints=False, nthreads=1, reps=1):
        """Timing helper: every thread scans its contiguous shard `reps` times. -> total matches over all reps."""
        lib = _load()
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        return int(lib.orc_scan_batch_reps(self._h, data.ctypes.data, offsets.ctypes.data, n, int(overlapping),
                                           int(codepoints), int(nthreads), None, None, 0, int(reps)))

    def scan_batch(self, data: np.ndarray, offsets: np.ndarray, overlapping=False, codepoints=False,
                   nthreads=1, want_records=True, rec_cap=None):
        """Batch drain. data: uint8 array, offsets: int64 (n+1).
        -> (total, counts[uint32 n], records[uint32 (k,4)] = hay,pid,start,end)."""
        lib = _load()
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        counts = np.zeros(n, dtype=np.uint32)
        if not want_records:
            total = lib.orc_scan_batch(self._h, data.ctypes.data, offsets.ctypes.data, n, int(overlapping),
                                       int(codepoints), int(nthreads), counts.ctypes.data, None, 0)
            return int(total), counts, None
        cap = rec_cap or max(1024, n)
        while True:
            rec = np.zeros((cap, 4), dtype=np.uint32)
            total = lib.orc_scan_batch(self._h, data.ctypes.data, offsets.ctypes.data, n, int(overlapping),
                                       int(codepoints), 1, counts.ctypes.data, rec.ctypes.data, cap)
            if total <= cap:
                return int(total), counts, rec[:total]
            cap = int(total)
