"""Test infrastructure: a pure-Python interpreter of the DEVICE IMAGE
(ahocorasick_rs_b200/csrc/automaton.h) that follows the CUDA code paths line
by line -- exact_scan / scan_byte / leftmost_flush (scan_core.cuh), the
per-lane logic of the segment kernel (scan_staged.cuh: warm-up, pieces, 16-byte
fast groups through the hot table with the trap row, hand-over to the exact
scanner), the validate/repair pass (repair.cuh) and the ordering + code point
fix-up (capi.cu).  It lets the CPU-only test run check the host builder's
tables and the kernels' control flow against the oracle without a GPU.  It is
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
NO_STATE, SPEC_SKIPPED = 0xFFFFFFFF, 0xFFFFFFFE
WARM, HEAD, NORMAL = 0, 1, 2
U32 = 0xFFFFFFFF


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
        self._L = L
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

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            self._L.acb_free(h)
            self._h = None

    def col(self, b):
        if self.col_mode == 0:
            return min((b - self.col_lo) & U32, self.n_cols - 1)
        return int(self.colmap[b])

    def hot_image(self, visits=None, max_rows=4096):
        """acb_hot_build -> (table[(rows+1), n_cols], hot2full, full2hot, rows)."""
        L = self._L
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
        """scan_staged_kernel prologue: the first H rows of the hot image, clamped to H.
        Returned as hot INDICES (the device keeps byte offsets = index * row bytes)."""
        table, h2f, f2h, rows = self.hot_image(visits)
        H = min(H, rows)
        row_bytes = 2 * self.n_cols
        assert np.all(table % row_bytes == 0) and table.max() == rows * row_bytes  # byte offsets of rows; trap = last row
        hot = np.minimum(table[: H + 1].astype(np.uint32) // row_bytes, H)
        hot[H, :] = H
        return hot, h2f, f2h, H


# ------------------------------------------------------------------ scan_core.cuh
class Piece:
    """PieceCtx"""

    def __init__(self):
        self.base = 0           # stream position of relative position 0
        self.at = self.stop = self.limit = 0
        self.emit_from = 0
        self.state = K_ROOT
        self.have = 0
        self.last_pid = self.last_end = 0
        self.hay = 0
        self.hay_delta = 0
        self.unit = 0
        self.nemit = 0
        self.cp_pos = self.cp_cont = 0


class Sink:
    def __init__(self, n_units):
        self.raw = []           # (hay, pid, start, end, seq, unit, aux)
        self.unit_counts = [0] * n_units


def byte_at(data, c, p):
    return int(data[c.base + p])


def cp_catch_up(data, c, to):
    while c.cp_pos < to:
        c.cp_cont += (byte_at(data, c, c.cp_pos) & 0xC0) == 0x80
        c.cp_pos += 1


def report(data, c, im, out, cp, pid, end):
    if end <= c.emit_from:
        return
    aux = 0
    if cp:
        cp_catch_up(data, c, end)
        aux = c.cp_cont
    hend = (end + c.hay_delta) & U32
    out.raw.append((c.hay, int(pid), hend - int(im.pat_len[pid]), hend, c.nemit, c.unit, aux))
    c.nemit += 1


def scan_byte(data, c, im, mode, emit):
    e = int(im.trans[c.state, im.col(byte_at(data, c, c.at))])
    c.state = e & MASK
    c.at += 1
    if e & FLAG:
        m0 = int(im.match_off[c.state])
        if mode == 0:
            emit(im.match_pid[m0], c.at)
            c.state = K_ROOT
        elif mode == 1:
            c.have, c.last_pid, c.last_end = 1, im.match_pid[m0], c.at
        else:
            at = c.at
            for k in range(m0, int(im.match_off[c.state + 1])):
                emit(im.match_pid[k], at)


def leftmost_flush(c, mode, emit):
    if mode != 1:
        return False
    if c.state == K_DEAD or c.at == c.limit:
        if c.have:
            emit(c.last_pid, c.last_end)
            c.at = c.last_end
            c.have = 0
            c.state = K_ROOT
            return True
        if c.state == K_DEAD:
            c.state = K_ROOT
    return False


def exact_scan(data, c, im, out, mode, cp, stop_hot=False, min_at=0, hot_limit=0, f2h=None):
    def emit(pid, end):
        report(data, c, im, out, cp, pid, end)

    while True:
        if leftmost_flush(c, mode, emit):
            continue
        if c.at >= c.stop and (mode != 1 or not c.have):
            break
        if stop_hot and c.at >= min_at and (mode != 1 or not c.have) and int(f2h[c.state]) < hot_limit:
            break
        scan_byte(data, c, im, mode, emit)


def find_haystack(offsets, p):
    lo, hi = 0, len(offsets) - 1
    while lo < hi:
        mid = (lo + hi) >> 1
        if offsets[mid + 1] <= p:
            lo = mid + 1
        else:
            hi = mid
    return lo


# ------------------------------------------------------------------ capi.cu: acb_plan_scan
def plan_scan(im, base_addr, total_bytes, n_haystacks, segment_bytes=1024):
    warm = max(16, (im.max_pat_len + 15) & ~15)
    seg = max(segment_bytes, 8 * warm)
    seg = (seg + 63) & ~63
    mis = base_addr & 63
    n_seg = (total_bytes + mis + seg - 1) // seg
    stride = 1
    if n_haystacks > 1:
        avg = total_bytes // n_haystacks
        stride = min(max((avg + seg // 2) // seg, 1), 65536)
    return dict(seg=seg, warm=warm, n_seg=n_seg, stride=stride, origin=-mis)


# ------------------------------------------------------------------ scan_staged.cuh: one lane = one segment
class LaneSeg:
    pass


def advance_piece(data, offsets, c, L, out, seg_info, cp):
    n_hay = len(offsets) - 1
    if L.kind == WARM:
        piece_end = min(L.hi_rel, c.limit)
        if c.at == L.lo_rel:
            L.spec_state = c.state
            L.kind = HEAD
            c.stop = piece_end
            c.emit_from = 0
            c.cp_pos, c.cp_cont = c.at, 0
            return
        L.spec_state = SPEC_SKIPPED
        L.kind = HEAD
        c.at = c.stop = piece_end
        c.emit_from = 0
        c.state = K_ROOT
        c.have = 0
        c.cp_pos, c.cp_cont = L.lo_rel, 0
    if L.kind == HEAD:
        L.head_count = c.nemit
    if c.stop == c.limit:
        h = L.h + 1
        while h < n_hay and offsets[h + 1] == offsets[h]:
            h += 1
        hi_pos = L.org + L.hi_rel
        if h < n_hay and offsets[h] < hi_pos:
            hs, he = int(offsets[h]), int(offsets[h + 1])
            L.h = h
            L.kind = NORMAL
            c.at = hs - L.org
            c.limit = he - L.org
            c.stop = min(L.hi_rel, c.limit)
            c.emit_from = 0
            c.state = K_ROOT
            c.have = 0
            c.hay = h
            c.hay_delta = (L.org - hs) & U32
            c.cp_pos, c.cp_cont = c.at, 0
            return
    # finish the segment
    cont_tail = 0
    if cp:
        if c.cp_pos <= L.hi_rel:
            cp_catch_up(data, c, min(L.hi_rel, c.limit))
            cont_tail = c.cp_cont
        else:
            cont_tail = c.cp_cont - sum((byte_at(data, c, p) & 0xC0) == 0x80 for p in range(L.hi_rel, c.cp_pos))
    seg_info[L.seg] = dict(spec_state=L.spec_state, end_state=c.state, end_over=c.at - L.hi_rel, head_count=L.head_count,
                           drop=0, cont_tail=cont_tail)
    out.unit_counts[2 * L.seg] = 0
    out.unit_counts[2 * L.seg + 1] = c.nemit
    L.done = True


def settle(data, offsets, c, L, im, out, seg_info, mode, cp, H, f2h, min_at):
    while True:
        exact_scan(data, c, im, out, mode, cp, True, min_at, H, f2h)
        if c.at >= c.stop and (mode != 1 or not c.have):
            advance_piece(data, offsets, c, L, out, seg_info, cp)
            if L.done:
                return
            min_at = c.at
            continue
        return


def staged_segment(im, data, offsets, plan, seg, mode, cp, hotinfo, out, seg_info, base_addr, stats=None):
    """scan_staged_kernel, one lane.  Stream positions are indices into `data`;
    the address of data[0] is base_addr (only its low 6 bits matter)."""
    hot, h2f, f2h, H = hotinfo
    S, warm, origin = plan["seg"], plan["warm"], plan["origin"]
    stream_lo, stream_hi = int(offsets[0]), int(offsets[-1])
    glo = origin + seg * S
    lo, hi = max(glo, stream_lo), min(glo + S, stream_hi)
    if lo >= hi:
        seg_info[seg] = dict(spec_state=NO_STATE, end_state=K_ROOT, end_over=0, head_count=0, drop=0, cont_tail=0)
        out.unit_counts[2 * seg] = out.unit_counts[2 * seg + 1] = 0
        return
    h = find_haystack(offsets, lo)
    hs, he = int(offsets[h]), int(offsets[h + 1])
    cont = hs < lo
    w = max(hs, lo - warm) if cont else lo
    pw = base_addr + w
    a0 = pw & ~63
    L = LaneSeg()
    L.org = w - (pw - a0)
    L.seg = seg
    L.lo_rel, L.hi_rel = lo - L.org, hi - L.org
    L.h = h
    L.kind = WARM if cont else NORMAL
    L.done = False
    L.spec_state, L.head_count = NO_STATE, 0
    nchunks = (L.hi_rel + 63) // 64
    c = Piece()
    c.base = L.org
    c.at = w - L.org
    c.limit = he - L.org
    c.stop = L.lo_rel if cont else min(L.hi_rel, c.limit)
    c.emit_from = U32 if cont else 0
    c.hay = h
    c.hay_delta = (L.org - hs) & U32
    c.unit = 2 * seg + 1
    c.cp_pos, c.cp_cont = c.at, 0
    assert c.at < c.stop
    st = {"pos": c.at, "stop": c.stop, "s": 0, "cpd": 0, "warm": cont}   # root = hot row 0

    def leave_fast(min_at):
        c.state = int(h2f[st["s"]])
        c.at = st["pos"]
        if cp:
            c.cp_pos, c.cp_cont = st["pos"], st["cpd"]
        settle(data, offsets, c, L, im, out, seg_info, mode, cp, H, f2h, min_at)
        st["pos"], st["stop"] = c.at, c.stop
        st["warm"] = (not L.done) and L.kind == WARM
        if not L.done:
            st["s"] = int(f2h[c.state])
            assert st["s"] < H, "exact_scan must hand back a hot state"
            if cp:
                cp_catch_up(data, c, st["pos"])
                st["cpd"] = c.cp_cont

    def piece_end_fast():
        if st["warm"]:
            L.spec_state = int(h2f[st["s"]])
            L.kind = HEAD
            st["stop"] = min(L.hi_rel, c.limit)
            c.stop = st["stop"]
            c.emit_from = 0
            st["cpd"] = 0
            st["warm"] = False
            return True
        if st["stop"] == L.hi_rel:
            seg_info[L.seg] = dict(spec_state=L.spec_state, end_state=int(h2f[st["s"]]), end_over=0,
                                   head_count=c.nemit if L.kind == HEAD else L.head_count, drop=0,
                                   cont_tail=st["cpd"] if cp else 0)
            out.unit_counts[2 * L.seg] = 0
            out.unit_counts[2 * L.seg + 1] = c.nemit
            L.done = True
            return True
        return False

    def fast_bytes(pos, n):
        """n bytes from pos through the hot table; -> (state, number of bytes before the trap or n)"""
        t = st["s"]
        grp = data[c.base + pos: c.base + pos + n]
        for b in grp:
            t = int(hot[t, im.col(int(b))])
        return t, grp

    for k in range(nchunks):
        relk = 64 * k
        if not L.done and st["warm"] and st["pos"] == st["stop"] and st["pos"] == relk:
            piece_end_fast()
        if not L.done and st["pos"] == relk and relk + 64 <= st["stop"]:
            t, grp = fast_bytes(relk, 64)
            if t != H:
                st["s"] = t
                st["pos"] += 64
                if cp:
                    st["cpd"] += int(np.count_nonzero((grp & 0xC0) == 0x80))
                if stats is not None:
                    stats["groups"] = stats.get("groups", 0) + 4
                continue
        for j in range(4):
            g = relk + 16 * j
            if L.done or st["pos"] < g or st["pos"] >= g + 16:
                continue
            if st["pos"] == g and g + 16 <= st["stop"]:
                t, grp = fast_bytes(g, 16)
                if t != H:
                    st["s"] = t
                    st["pos"] += 16
                    if cp:
                        st["cpd"] += int(np.count_nonzero((grp & 0xC0) == 0x80))
                    if stats is not None:
                        stats["groups"] = stats.get("groups", 0) + 1
                    continue
            if stats is not None:
                stats["groups"] = stats.get("groups", 0) + 1
            trapped = False
            while not L.done and g <= st["pos"] < g + 16:
                if st["pos"] >= st["stop"]:
                    if not (st["pos"] == st["stop"] and piece_end_fast()):
                        leave_fast(st["stop"])
                    continue
                b = int(data[c.base + st["pos"]])
                t = int(hot[st["s"], im.col(b)])
                if t != H:
                    st["s"] = t
                    st["pos"] += 1
                    if cp:
                        st["cpd"] += (b & 0xC0) == 0x80
                else:
                    trapped = True
                    leave_fast(st["pos"] + 1)
            if trapped and stats is not None:
                stats["traps"] = stats.get("traps", 0) + 1
    while not L.done:
        if not (st["pos"] == st["stop"] and piece_end_fast()):
            leave_fast(st["stop"])


# ------------------------------------------------------------------ repair.cuh
def piece_finished(m, stop, mode):
    return m.at >= stop and (mode != 1 or not m.have)


def machine_step(data, m, im, mode, emit):
    if not leftmost_flush(m, mode, emit):
        if m.at < m.limit:
            scan_byte(data, m, im, mode, emit)


def repair(im, data, offsets, plan, mode, cp, out, seg_info, stats=None):
    S, origin = plan["seg"], plan["origin"]
    stream_hi = int(offsets[-1])
    for h in range(len(offsets) - 1):
        hs, he = int(offsets[h]), int(offsets[h + 1])
        if he <= hs:
            continue
        js, je = (hs - origin) // S, (he - 1 - origin) // S
        if js == je:
            continue
        limit = he - hs
        T = Piece()
        T.base = hs
        T.limit = limit
        T.hay = h
        slot = {"seg": -1, "seq": 0}

        def close_slot():
            if slot["seg"] >= 0:
                out.unit_counts[2 * slot["seg"]] = slot["seq"]

        def emit_true(pid, end):
            m = (hs + end - 1 - origin) // S
            if m != slot["seg"]:
                close_slot()
                slot["seg"], slot["seq"] = m, 0
            aux = 0
            if cp:
                lo_m = max(origin + m * S, hs)
                aux = sum((int(data[hs + p]) & 0xC0) == 0x80 for p in range(lo_m - hs, end))
            out.raw.append((h, int(pid), end - int(im.pat_len[pid]), end, slot["seq"], 2 * m, aux))
            slot["seq"] += 1

        k = js + 1
        while k <= je:
            prev, spec = seg_info[k - 1], seg_info[k]["spec_state"]
            if prev["end_over"] == 0 and spec == prev["end_state"]:
                k += 1
                continue
            if stats is not None:
                stats["repairs"] = stats.get("repairs", 0) + 1
            T.at = (origin + k * S - hs) + prev["end_over"]
            T.state = prev["end_state"]
            T.have = 0
            cur = k
            while True:
                info = seg_info[cur]
                lo_c = origin + cur * S
                hi_c = min(lo_c + S, stream_hi)
                pstop = min(hi_c, he) - hs
                p_alive = info["spec_state"] != SPEC_SKIPPED
                Pm = Piece()
                Pm.base = hs
                Pm.limit = limit
                Pm.at = lo_c - hs
                Pm.state = info["spec_state"]
                d = [0]

                def count_spec(pid, end):
                    d[0] += 1

                converged = False
                while p_alive and not piece_finished(Pm, pstop, mode):
                    if T.at == Pm.at and T.state == Pm.state and not T.have and not Pm.have:
                        converged = True
                        break
                    t_done = T.at >= limit and not T.have
                    if not t_done and T.at <= Pm.at:
                        machine_step(data, T, im, mode, emit_true)
                    else:
                        machine_step(data, Pm, im, mode, count_spec)
                if converged:
                    info["drop"] = d[0]
                    out.unit_counts[2 * cur + 1] -= d[0]
                    k = cur + 1
                    break
                info["drop"] = info["head_count"]
                out.unit_counts[2 * cur + 1] -= info["head_count"]
                while not piece_finished(T, pstop, mode):
                    machine_step(data, T, im, mode, emit_true)
                if pstop == limit:
                    k = je + 1
                    break
                cur += 1
                if T.at == origin + cur * S - hs and not T.have and seg_info[cur]["spec_state"] == T.state:
                    k = cur + 1
                    break
            close_slot()
            slot["seg"] = -1


# ------------------------------------------------------------------ capi.cu: ordering + code point fix-up
def order_matches(im, offsets, plan, cp, out, seg_info, segments=True):
    n_units = len(out.unit_counts)
    unit_off = np.zeros(n_units + 1, dtype=np.int64)
    np.cumsum(out.unit_counts, out=unit_off[1:])
    total = int(unit_off[-1])
    res = [None] * total
    cont_cum = None
    if segments and cp:
        tails = [seg_info[j]["cont_tail"] for j in range(plan["n_seg"])]
        cont_cum = np.zeros(plan["n_seg"] + 1, dtype=np.int64)
        np.cumsum(tails, out=cont_cum[1:])
    for (hay, pid, start, end, seq, unit, aux) in out.raw:
        if segments and (unit & 1):
            drop = seg_info[unit >> 1]["drop"]
            if seq < drop:
                continue
            seq -= drop
        dst = int(unit_off[unit]) + seq
        if cp:
            cont = aux
            if segments:
                j = unit >> 1
                hs = int(offsets[hay])
                if hs < plan["origin"] + j * plan["seg"]:
                    j0 = (hs - plan["origin"]) // plan["seg"]
                    cont += int(cont_cum[j] - cont_cum[j0])
            end = end - cont
            start = end - int(im.pat_cplen[pid])
        assert res[dst] is None, "two matches for one output slot"
        res[dst] = (hay, pid, start, end)
    assert all(r is not None for r in res), "hole in the ordered output"
    return res


def emulate_scan(im, data, offsets, overlapping=False, cp=False, segment_bytes=1024, H=4096, visits=None, base_addr=0,
                 stats=None):
    """acb_scan_batch (staged path) end to end: [(haystack, pattern, start, end)] in output order."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    offsets = [int(x) for x in offsets]
    mode = 2 if overlapping else (0 if im.kind == 0 else 1)
    plan = plan_scan(im, base_addr, len(data), len(offsets) - 1, segment_bytes)
    if len(offsets) < 2 or offsets[-1] == offsets[0]:
        return []
    hotinfo = im.hot_table(H, visits)
    out = Sink(2 * plan["n_seg"])
    seg_info = {}
    for seg in range(plan["n_seg"]):
        staged_segment(im, data, offsets, plan, seg, mode, cp, hotinfo, out, seg_info, base_addr, stats)
    if mode != 2:
        repair(im, data, offsets, plan, mode, cp, out, seg_info, stats)
    return order_matches(im, offsets, plan, cp, out, seg_info)


def emulate_plain(im, data, offsets, overlapping=False, cp=False):
    """scan_plain_kernel + ordering: one exact scanner per haystack."""
    data = np.ascontiguousarray(data, dtype=np.uint8)
    mode = 2 if overlapping else (0 if im.kind == 0 else 1)
    n = len(offsets) - 1
    out = Sink(n)
    for h in range(n):
        c = Piece()
        c.base = int(offsets[h])
        c.stop = c.limit = int(offsets[h + 1]) - int(offsets[h])
        c.hay = h
        c.unit = h
        exact_scan(data, c, im, out, mode, cp)
        out.unit_counts[h] = c.nemit
    return order_matches(im, offsets, None, cp, out, None, segments=False)


def find(im, hay: bytes, overlapping=False, cp=False):
    """One haystack through the plain path -> [(pid, start, end)]."""
    data = np.frombuffer(hay or b"\0", dtype=np.uint8)
    return [(p, s, e) for (_, p, s, e) in emulate_plain(im, data, [0, len(hay)], overlapping, cp)]


def find_staged(im, hay: bytes, overlapping=False, cp=False, **kw):
    """One haystack through the segment path -> [(pid, start, end)]."""
    data = np.frombuffer(hay or b"\0", dtype=np.uint8)
    return [(p, s, e) for (_, p, s, e) in emulate_scan(im, data, [0, len(hay)], overlapping, cp, **kw)]
