"""Test infrastructure: a pure-Python interpreter of the SIEVE IMAGE
(ahocorasick_rs_b200/csrc/sieve.h) that follows the CUDA code path of
scan_sieve.cuh position by position -- first filter probe, remaining probes and
the on-chip walk through the deeper filter levels, hash table + reverse-trie
walk, emission along the terminal links -- and the epilogue's per-haystack
selection of the non-overlapping matches (capi.cu: select_non_overlapping).
It lets the CPU-only test run check the host builder's tables and the scan's
logic against the oracle without a GPU.  It is NOT a product path (the product
has no CPU fallback) and is far too slow to be one."""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from ahocorasick_rs_b200 import _capi

U32 = 0xFFFFFFFF
MIX_HI, MUL_A, MUL_B, MUL_C = 0x9E3779B1, 0x85EBCA6B, 0xC2B2AE35, 0x27D4EB2F
SALT_TERM, MUL_SLOT = 0x5BD1E995, 0x7FEB352D
NO_NODE = 0xFFFFFFFF
TERMINAL = 1 << 17
HDR_FMT = "<16I6Q"


def step(x, b):
    x = ((x + b + 1) * MIX_HI) & U32
    return x ^ (x >> 15)


class SieveImage:
    def __init__(self, patterns, kind=0, bloom_bytes_max=200 * 1024, w_max=0):
        L = _capi.lib()
        pats = [bytes(p) for p in patterns]
        offs = np.zeros(len(pats) + 1, dtype=np.uint64)
        if pats:
            np.cumsum([len(p) for p in pats], out=offs[1:])
        blob = np.frombuffer(b"".join(pats) or b"\0", dtype=np.uint8)
        h = C.c_void_p()
        rc = L.acb_build(blob.ctypes.data, offs.ctypes.data, len(pats), kind, -1, C.byref(h))
        if rc != 0:
            raise ValueError(_capi.last_error())
        self._h, self._L = h, L
        n = L.acb_sieve_build(h, bloom_bytes_max, w_max)
        assert n, _capi.last_error()
        buf = np.zeros(n, dtype=np.uint8)
        assert L.acb_sieve_write(h, buf.ctypes.data, n) == 0
        self.raw = buf
        f = struct.unpack_from(HDR_FMT, buf.tobytes()[: struct.calcsize(HDR_FMT)])
        (magic, self.W, self.last_level, self.n_probes, self.bloom_words, self.ht_mask, self.n_nodes, self.n_pids,
         self.max_pat_len, self.min_pat_len, self.n_keys, self.n_entries, self.prim_words, self.term_levels, _p1, _p2, o_bloom, o_ht, o_na, o_nb, o_pids, total) = f
        assert magic == 0x32424341 and total == n
        self.bloom = buf[o_bloom:o_bloom + 4 * self.bloom_words].view(np.uint32)
        self.ht = buf[o_ht:o_ht + 16 * (self.ht_mask + 1)].view(np.uint32).reshape(-1, 4)
        nn = max(self.n_nodes, 1)
        self.na = buf[o_na:o_na + 8 * nn].view(np.uint32).reshape(-1, 2)
        self.nb = buf[o_nb:o_nb + 32 * nn].view(np.uint32).reshape(-1, 8)
        self.pids = buf[o_pids:o_pids + 4 * self.n_pids].view(np.uint32)
        self.kind = kind
        self.stats = {"pos": 0, "probe1": 0, "stage1": 0, "stage2": 0, "confirmed": 0}

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._L.acb_free(h)
            self._h = None

    # ---- the filter -----------------------------------------------------------------------
    def prim_bit(self, p):
        word = int(self.bloom[(p * self.prim_words) >> 32])
        return (word >> (p & 31)) & 1

    def has(self, x):
        """the secondary filter"""
        sec_words = self.bloom_words - self.prim_words
        for m in (MUL_B, MUL_C)[: self.n_probes]:
            p = (x * m) & U32
            word = int(self.bloom[self.prim_words + ((p * sec_words) >> 32)])
            if not (word >> (p & 31)) & 1:
                return False
        return True

    def key(self, text, e):
        """(lo', hi') of the W bytes before e (bytes before the stream read as zero, like the kernel's loads)."""
        W = self.W
        b = [text[e - W + i] if e - W + i >= 0 else 0 for i in range(W)]
        if W <= 4:
            lo = sum(b[i] << (8 * i) for i in range(W))
            return lo, 0
        lo = sum(b[W - 4 + i] << (8 * i) for i in range(4))
        hi = sum(b[i] << (8 * i) for i in range(W - 4))
        return lo, hi

    # ---- one position: every pattern ending at e inside [hs, ...) --------------------------------
    def matches_ending_at(self, text, e, hs):
        """-> [(pid, start)] longest first, like the kernel emits them; [] when the filter or the trie says no."""
        W = self.W
        st = self.stats
        st["pos"] += 1
        if e - W < 0:
            return []
        lo, hi = self.key(text, e)
        x = (lo + hi * MIX_HI) & U32
        if not self.prim_bit((x * MUL_A) & U32):
            return []
        st["probe1"] += 1
        if not self.has(x):
            return []
        st["stage1"] += 1
        d, go, xx = W, False, x
        while True:
            if d >= self.last_level and self.max_pat_len > self.last_level:
                go = True
                break
            if (self.term_levels >> d) & 1 and self.has(xx ^ SALT_TERM):
                go = True
                break
            if d >= self.last_level:
                break
            b = text[e - d - 1] if e - d - 1 >= 0 else 0
            xx = step(xx, b)
            d += 1
            if not self.has(xx):
                break
        if not go:
            return []
        st["stage2"] += 1
        if e - W < hs:
            return []
        size = self.ht_mask + 1
        s = (((x * MUL_SLOT) & U32) * size) >> 32
        v = NO_NODE
        while True:
            ent = self.ht[s]
            if int(ent[2]) == NO_NODE:
                break
            if int(ent[0]) == lo and int(ent[1]) == hi:
                v = int(ent[2])
                break
            s = (s + 1) & self.ht_mask
        best, d = NO_NODE, W
        while v != NO_NODE:
            first, meta = int(self.na[v][0]), int(self.na[v][1])
            if meta & TERMINAL:
                best = v
            nk = (meta >> 8) & 0x1FF
            if nk == 0 or e - 1 - d < hs:
                break
            b = text[e - 1 - d]
            c = NO_NODE
            for t in range(nk):
                cb = int(self.na[first + t][1]) & 0xFF
                if cb >= b:
                    if cb == b:
                        c = first + t
                    break
            v = c
            d += 1
        if best == NO_NODE:
            return []
        st["confirmed"] += 1
        out = []
        u = best
        while u != NO_NODE:
            own_off, own_cnt, link, depth, chain = (int(z) for z in self.nb[u][:5])
            if u == best:
                expect = chain
            for t in range(own_cnt):
                out.append((int(self.pids[own_off + t]), e - depth))
            u = link
        assert len(out) == expect, "chain_cnt disagrees with the chain"
        return out

    def overlapping(self, data, offs):
        """The overlapping list of a batch: [(haystack, pid, start, end)] byte offsets, in the reference's order."""
        text = bytes(data)
        out = []
        for h in range(len(offs) - 1):
            hs, he = int(offs[h]), int(offs[h + 1])
            for e in range(hs + 1, he + 1):
                for pid, start in self.matches_ending_at(text, e, hs):
                    out.append((h, pid, start - hs, e - hs))
        return out


def select(rows, mode, max_len, longest):
    """capi.cu select_non_overlapping on one haystack's rows [(h, pid, start, end)] sorted by (end, start, pid)."""
    out, s = [], 0
    if mode == 0:
        for r in rows:
            if r[2] >= s:
                out.append(r)
                s = r[3]
        return out
    i, n = 0, len(rows)
    while i < n:
        best = None
        for j in range(i, n):
            m = rows[j]
            if best is not None and m[3] > best[2] + max_len:
                break
            if m[2] < s:
                continue
            if best is None or m[2] < best[2]:
                best = m
            elif m[2] == best[2]:
                if longest and (m[3] > best[3] or (m[3] == best[3] and m[1] < best[1])):
                    best = m
                elif not longest and m[1] < best[1]:
                    best = m
        if best is None:
            break
        out.append(best)
        s = best[3]
        while i < n and rows[i][3] <= s:
            i += 1
    return out


def scan(image: SieveImage, data, offs, overlapping=False):
    """What acb_scan_batch returns through the sieve path, byte offsets."""
    rows = image.overlapping(data, offs)
    if overlapping:
        return rows
    out = []
    by_h = {}
    for r in rows:
        by_h.setdefault(r[0], []).append(r)
    for h in sorted(by_h):
        out += select(by_h[h], 0 if image.kind == 0 else 1, image.max_pat_len, image.kind == 2)
    return out
