"""Test infrastructure: a pure-Python interpreter of the DEVICE IMAGE
(ahocorasick_rs_b200/csrc/automaton.h) that follows the CUDA code paths line
by line -- exact_scan (scan_core.cuh) and the per-lane logic of the staged
kernel (scan_staged.cuh: head / 16-byte fast groups through the hot table with
the trap row / tail).  It lets the CPU-only test run check the host builder's
tables and the kernel's control flow against the oracle without a GPU.  It is
NOT a product path (the product has no CPU fallback) and is far too slow to be
one."""
from __future__ import annotations

import ctypes as C
import struct

import numpy as np

from ahocorasick_rs_b200 import _capi

K_DEAD, K_ROOT = 0, 1
FLAG = 0x80000000
MASK = 0x7FFFFFFF
HDR_FMT = "<12I7Q"


class Image:
    def __init__(self, patterns, kind=0, implementation=-1):
        L = _capi.lib()
        pats = [bytes(p) for p in patterns]
        offs = np.zeros(len(pats) + 1, dtype=np.uint64)
        if pats:
            np.cumsum([len(p) for p in pats], out=offs[1:])
        blob = np.frombuffer(b"".join(pats) or b"\0", dtype=np.uint8)
        h = C.c_void_p()
        rc = L.acb_build(blob.ctypes.data, offs.ctypes.data, len(pats), kind, implementation, C.byref(h))
        if rc != 0:
            raise ValueError(_capi.last_error())
        n = L.acb_image_bytes(h)
        buf = np.zeros(n, dtype=np.uint8)
        assert L.acb_image_write(h, buf.ctypes.data, n) == 0
        self._h = h
        self.raw = buf
        f = struct.unpack_from(HDR_FMT, buf.tobytes()[: struct.calcsize(HDR_FMT)])
        (self.magic, self.version, self.kind, self.col_mode, self.n_states, self.n_cols, self.col_lo, self.n_patterns,
         self.max_pat_len, self.min_pat_len, self.n_hot, _res, o_colmap, o_trans, o_moff, o_mpid, o_plen, o_pcp,
         self.total_bytes) = f
        assert self.magic == 0x30424341 and self.total_bytes == n
        self.colmap = buf[o_colmap:o_colmap + 256]
        self.trans = buf[o_trans:o_trans + 4 * self.n_states * self.n_cols].view(np.uint32).reshape(self.n_states, self.n_cols)
        self.match_off = buf[o_moff:o_moff + 4 * (self.n_states + 1)].view(np.uint32)
        self.match_pid = buf[o_mpid:o_mpid + 4 * int(self.match_off[-1])].view(np.uint32)
        self.pat_len = buf[o_plen:o_plen + 4 * self.n_patterns].view(np.uint32)
        self.pat_cplen = buf[o_pcp:o_pcp + 4 * self.n_patterns].view(np.uint32)

    def col(self, b):
        if self.col_mode == 0:
            return min((b - self.col_lo) & 0xFFFFFFFF, self.n_cols - 1)
        return int(self.colmap[b])

    def hot_image(self, visits=None, max_rows=4096):
        """acb_hot_build -> (table[(rows+1), n_cols], hot2full, full2hot, rows)."""
        L = _capi.lib()
        n = L.acb_hot_bytes(self._h, max_rows)
        buf = np.zeros(n, dtype=np.uint8)
        vp = None
        if visits is not None:
            visits = np.ascontiguousarray(visits, dtype=np.uint32)
            vp = visits.ctypes.data
        assert L.acb_hot_build(self._h, vp, max_rows, buf.ctypes.data, n) == 0
        rows = L.acb_hot_rows(buf.ctypes.data)
        magic, n_rows, n_cols, n_states, o_table, o_h2f, o_f2h, total = struct.unpack_from("<4I4Q", buf.tobytes()[:48])
        assert magic == 0x31424341 and n_rows == rows and n_cols == self.n_cols and total == n
        table = buf[o_table:o_table + 2 * (rows + 1) * n_cols].view(np.uint16).reshape(rows + 1, n_cols)
        h2f = buf[o_h2f:o_h2f + 4 * (rows + 1)].view(np.uint32)
        f2h = buf[o_f2h:o_f2h + 2 * n_states].view(np.uint16)
        return table, h2f, f2h, rows

    def hot_table(self, H, visits=None):
        """scan_staged_kernel prologue: the first H rows of the hot image, clamped to H."""
        table, h2f, f2h, rows = self.hot_image(visits)
        H = min(H, rows)
        row_bytes = 2 * self.n_cols
        assert np.all(table % row_bytes == 0) and table.max() == rows * row_bytes  # byte offsets of rows; trap = last row
        hot = np.minimum(table[: H + 1].astype(np.uint32) // row_bytes, H)
        hot[H, :] = H
        return hot, h2f, f2h, H


class Ctx:
    def __init__(self, hay, at, end, emit_from, cp):
        self.hay = hay
        self.at, self.end, self.emit_from = at, end, emit_from
        self.state = K_ROOT
        self.have = False
        self.last_pid = self.last_end = 0
        self.cp = cp
        self.cp_pos = at          # init_unit: counting starts at the first byte the unit reads
        self.cp_count = 0
        self.out = []


def report(c, im, pid, end):
    if end <= c.emit_from:
        return
    start = end - int(im.pat_len[pid])
    if c.cp:
        while c.cp_pos < end:
            c.cp_count += (c.hay[c.cp_pos] & 0xC0) != 0x80
            c.cp_pos += 1
        end = c.cp_count
        start = end - int(im.pat_cplen[pid])
    c.out.append((int(pid), start, end))


def exact_scan(c, im, mode, stop_hot=False, min_at=0, phase=0, hot_limit=0, f2h=None):
    s, at, end = c.state, c.at, c.end
    while True:
        if mode == 1:
            if at == end or s == K_DEAD:
                if c.have:
                    report(c, im, c.last_pid, c.last_end)
                    at = c.last_end
                    c.have = False
                    s = K_ROOT
                    continue
                if at == end:
                    break
                s = K_ROOT
        elif at == end:
            break
        if (stop_hot and at >= min_at and ((at - phase) & 15) == 0 and (mode != 1 or not c.have)
                and int(f2h[s]) < hot_limit):
            break
        e = int(im.trans[s, im.col(c.hay[at])])
        s = e & MASK
        at += 1
        if e & FLAG:
            m0 = int(im.match_off[s])
            if mode == 0:
                report(c, im, im.match_pid[m0], at)
                s = K_ROOT
            elif mode == 1:
                c.have, c.last_pid, c.last_end = True, im.match_pid[m0], at
            else:
                for k in range(m0, int(im.match_off[s + 1])):
                    report(c, im, im.match_pid[k], at)
    c.state, c.at = s, at


def cp_catch_up(c, to):
    while c.cp_pos < to:
        c.cp_count += (c.hay[c.cp_pos] & 0xC0) != 0x80
        c.cp_pos += 1


def staged_lane(im, hay: bytes, mode, H, base_addr=0, at=0, end=None, emit_from=0, cp=False, hot=None, stats=None,
                visits=None):
    """One lane of scan_staged_kernel over hay[at:end]; base_addr = absolute
    address of hay[0] (only its low bits matter: 16-byte group alignment)."""
    end = len(hay) if end is None else end
    c = Ctx(hay, at, end, emit_from, cp)
    hot, h2f, f2h, H = im.hot_table(H, visits) if hot is None else hot
    phase = (-base_addr) & 15
    p0 = base_addr + at
    a0 = p0 & ~63
    pe = base_addr + end
    nchunks = (pe - a0 + 63) // 64 if pe > a0 else 0
    rel0 = at - (p0 - a0)
    exact_scan(c, im, mode, True, c.at, phase, H, f2h)
    pos, s = c.at, int(f2h[c.state])
    cpd = 0
    if cp:
        cp_catch_up(c, pos)
        cpd = pos - c.cp_count
    for k in range(nchunks):
        relk = rel0 + 64 * k
        for j in range(4):
            g = relk + 16 * j
            if g == pos and g + 16 <= end:
                t = s
                for i in range(16):
                    t = int(hot[t, im.col(hay[pos + i])])
                if stats is not None:
                    stats["groups"] = stats.get("groups", 0) + 1
                if t != H:
                    s = t
                    if cp:
                        cpd += sum((b & 0xC0) == 0x80 for b in hay[pos:pos + 16])
                    pos += 16
                else:
                    if stats is not None:
                        stats["traps"] = stats.get("traps", 0) + 1
                    c.state, c.at = int(h2f[s]), pos
                    if cp:
                        c.cp_pos, c.cp_count = pos, pos - cpd
                    exact_scan(c, im, mode, True, pos + 16, phase, H, f2h)
                    s, pos = int(f2h[c.state]), c.at
                    if cp:
                        cp_catch_up(c, pos)
                        cpd = pos - c.cp_count
    if pos < c.end:
        c.state, c.at = int(h2f[s]), pos
        if cp:
            c.cp_pos, c.cp_count = pos, pos - cpd
        exact_scan(c, im, mode)
    return c.out


def find(im, hay: bytes, overlapping=False, cp=False):
    """Whole-haystack exact scan (the plain kernel)."""
    mode = 2 if overlapping else (0 if im.kind == 0 else 1)
    c = Ctx(hay, 0, len(hay), 0, cp)
    exact_scan(c, im, mode)
    return c.out


def find_chunked(im, hay: bytes, chunk, H=None, cp=False, base_addr=0):
    """acb_scan_chunked: overlapping, units = chunks with a halo of max_pat_len-1."""
    halo = max(im.max_pat_len - 1, 0)
    out = []
    n = (len(hay) + chunk - 1) // chunk
    cps = 0
    for u in range(n):
        lo, hi = u * chunk, min((u + 1) * chunk, len(hay))
        at = lo - halo if lo > halo else 0
        cps_at = cps - sum((b & 0xC0) != 0x80 for b in hay[at:lo])
        if H is None:
            c = Ctx(hay, at, hi, lo, cp)
            c.cp_count = cps_at
            exact_scan(c, im, 2)
            out += c.out
        else:
            # staged lane with pre-seeded code point count
            res = _staged_chunk(im, hay, H, base_addr, at, hi, lo, cp, cps_at)
            out += res
        cps += sum((b & 0xC0) != 0x80 for b in hay[lo:hi])
    return out


def _staged_chunk(im, hay, H, base_addr, at, end, emit_from, cp, cps):
    # same as staged_lane but with c.cp_count preset (Units.chunk_cp)
    orig = Ctx.__init__

    def patched(self, hay_, at_, end_, emit_from_, cp_):
        orig(self, hay_, at_, end_, emit_from_, cp_)
        self.cp_count = cps

    Ctx.__init__ = patched
    try:
        return staged_lane(im, hay, 2, H, base_addr, at, end, emit_from, cp)
    finally:
        Ctx.__init__ = orig
