// scan_core.cuh -- device-side pieces shared by the scan kernels.
//
// The exact (byte-at-a-time) scanner below is the definition of the result:
// it is the reference's iterator drain (src/lib.rs:42-68, 238-248, 433; the
// crate's find / find_overlapping loops) run by ONE thread, reading the dense
// table in global memory.  The staged kernel (scan_staged.cuh) only
// accelerates the stretches where nothing can happen and drops back here for
// everything else, and the segment-parallel scheme (DESIGN.md "Segments")
// validates every speculative start against the true state and repairs the
// rare misses with this same scanner, so all kernels produce the same matches
// in the same order by construction.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/acb200.h"
#include "automaton.h"

namespace acb {

enum ScanMode : int {
    kModeStandard = 0,  // non-overlapping, report at the first match state, restart at the root
    kModeLeftmost = 1,  // non-overlapping, keep the last match until the dead state, restart after it
    kModeOverlap = 2,   // every pattern on every match state's list, never restart
};

constexpr uint32_t kNoState = 0xffffffffu;    // SegInfo.spec_state: the segment starts at a haystack start
constexpr uint32_t kSpecSkipped = 0xfffffffeu;  // SegInfo.spec_state: no usable speculation, head piece left to the repair pass

// Views into the device image (all global memory).
struct DevImage {
    const uint8_t *colmap;
    const uint32_t *trans;
    const uint32_t *match_off;
    const uint32_t *match_pid;
    const uint32_t *pat_len;
    const uint32_t *pat_cplen;
    uint32_t n_cols, col_lo, n_states, col_mode;
};

// The input: haystack h = bytes[offsets[h], offsets[h+1]).
struct Batch {
    const uint8_t *bytes;
    const int64_t *offsets;
    int64_t n_haystacks;
};

// Segments: the byte stream [offsets[0], offsets[n]) cut on a fixed grid of
// `seg_bytes` (a multiple of 64, anchored at a 64-byte aligned ADDRESS), so
// every lane gets the same amount of work whatever the haystack lengths are.
// A segment that starts in the middle of a haystack starts from a SPECULATED
// state (obtained by scanning `warm` bytes before it from the root); the
// validate/repair pass checks it against the true state.
struct SegPlan {
    int64_t origin;       // stream position of segment 0's grid start: the 64-byte aligned address at or before `bytes`
    uint32_t seg_bytes;   // (the stream itself is [offsets[0], offsets[n]), read on the device)
    uint32_t warm;        // multiple of 16, >= max_pattern_len - 1
    int64_t n_segments;
    uint32_t lane_stride;  // lane l of warp-task (a, j) scans segment (32 a + l) * lane_stride + j
    uint64_t avg_len;      // hint: total bytes / haystacks (exact for equal-length batches): O(1) haystack lookup
};

// Per-segment summary written by the scan kernel, read by validate/repair and the ordering pass.
struct SegInfo {
    uint32_t spec_state;  // state assumed at the segment start (kNoState / kSpecSkipped: see above)
    uint32_t end_state;   // state when the scan stopped (meaningful when the segment ends inside a haystack)
    uint32_t end_over;    // how far past the segment end the scan stopped (leftmost kinds may overrun)
    uint32_t head_count;  // matches reported by the speculative head piece (the first head_count of slot 1)
    uint32_t drop;        // leading slot-1 matches superseded by the repair pass
    uint32_t cont_tail;   // code points: continuation bytes of the last piece (the one that runs into the next segment)
    uint32_t reserved0, reserved1;
};

struct Sink {
    acb_match *raw;
    uint32_t *raw_seq;
    uint32_t *raw_unit;
    uint32_t *raw_aux;  // code points: continuation bytes between the piece's counting origin and the match end
    unsigned long long cap;
    uint32_t *unit_counts;
    unsigned long long *raw_total;
};

// Scanner state for one piece (the part of one haystack inside one scan unit).
// Positions are relative to `base`; position + hay_delta = byte offset in the haystack.
struct PieceCtx {
    const uint8_t *base;
    uint32_t at;         // next byte to read
    uint32_t stop;       // end of this piece (the scan runs past it only while a leftmost match is pending)
    uint32_t limit;      // end of the haystack = end of input
    uint32_t emit_from;  // report only matches that end after this position
    uint32_t state;
    uint32_t have, last_pid, last_end;  // leftmost: the match to report once the automaton dies / input ends
    uint32_t hay, hay_delta;
    uint32_t unit, nemit;      // output slot and rank of the next match in it
    uint32_t cp_pos, cp_cont;  // code points: continuation bytes counted in [counting origin, cp_pos)
};

__device__ __forceinline__ uint32_t ld_u8(const uint8_t *p) { return __ldg(p); }

// advance the continuation-byte counter to position `to` (no-op when already there or past it)
__device__ __forceinline__ void cp_catch_up(PieceCtx &c, uint32_t to) {
    uint32_t p = c.cp_pos, n = c.cp_cont;
    while (p < to) {
        n += (ld_u8(c.base + p) & 0xC0u) == 0x80u;
        p++;
    }
    c.cp_pos = p;
    c.cp_cont = n;
}

template <bool CP>
__device__ __forceinline__ void report(PieceCtx &c, const DevImage &im, const Sink &out, uint32_t pid, uint32_t end) {
    if (end <= c.emit_from) return;
    uint32_t aux = 0;
    if (CP) {
        // ends are reported in non-decreasing order, so one forward-only counter is enough
        cp_catch_up(c, end);
        aux = c.cp_cont;
    }
    const uint32_t hend = end + c.hay_delta;
    const unsigned long long i = atomicAdd(out.raw_total, 1ULL);
    if (i < out.cap) {
        acb_match m;
        m.haystack = c.hay;
        m.pattern = pid;
        m.start = hend - __ldg(im.pat_len + pid);
        m.end = hend;
        *reinterpret_cast<uint4 *>(out.raw + i) = *reinterpret_cast<uint4 *>(&m);
        out.raw_seq[i] = c.nemit;
        out.raw_unit[i] = c.unit;
        if (CP) out.raw_aux[i] = aux;
    }
    c.nemit++;
}

// Where the staged fast path may resume: full2hot[state] < hot_limit.
struct HotMap {
    const uint16_t *full2hot;
    uint32_t hot_limit;
};

// One transition and its consequences.  Shared by exact_scan and the repair
// pass (which steps two scanners side by side), so both follow the same rules.
template <int MODE, typename Emit>
__device__ __forceinline__ void scan_byte(PieceCtx &c, const DevImage &im, uint32_t &s, uint32_t &at, Emit &&emit) {
    const uint32_t b = ld_u8(c.base + at);
    // kColRange: the column is arithmetic (one dependent load less per byte)
    const uint32_t col = im.col_mode == kColRange ? min(b - im.col_lo, im.n_cols - 1) : (uint32_t)__ldg(im.colmap + b);
    const uint32_t e = __ldg(im.trans + (size_t)s * im.n_cols + col);
    s = e & kStateMask;
    at++;
    if (e & kMatchFlag) {
        const uint32_t m0 = __ldg(im.match_off + s);
        if (MODE == kModeStandard) {
            emit(__ldg(im.match_pid + m0), at);
            s = kRoot;
        } else if (MODE == kModeLeftmost) {
            c.have = 1;
            c.last_pid = __ldg(im.match_pid + m0);
            c.last_end = at;
        } else {
            const uint32_t m1 = __ldg(im.match_off + s + 1);
            for (uint32_t k = m0; k < m1; k++) emit(__ldg(im.match_pid + k), at);
        }
    }
}

// Leftmost kinds: what happens BEFORE reading the next byte.  Returns true when
// a pending match was reported and the search restarted right behind it.
template <int MODE, typename Emit>
__device__ __forceinline__ bool leftmost_flush(PieceCtx &c, uint32_t &s, uint32_t &at, Emit &&emit) {
    if (MODE != kModeLeftmost) return false;
    if (s == kDead || at == c.limit) {
        if (c.have) {
            // the crate's iterator: report, then search again from the match end
            emit(c.last_pid, c.last_end);
            at = c.last_end;
            c.have = 0;
            s = kRoot;
            return true;
        }
        if (s == kDead) s = kRoot;  // unreachable: the dead state is only entered below a match state
    }
    return false;
}

// Runs the exact scanner from c.at until the piece is finished: at >= c.stop
// with nothing pending (leftmost kinds run past c.stop while a match is
// pending, never past c.limit) -- or, if stop_hot is set, as soon as
//   at >= min_at, the state is hot and nothing is pending,
// i.e. at a point where the staged fast path may take over again.
template <int MODE, bool CP>
__device__ __forceinline__ void exact_scan(PieceCtx &c, const DevImage &im_in, const Sink &out, bool stop_hot,
                                        uint32_t min_at, HotMap hm) {
    const DevImage im = im_in;  // private copy: the loop keeps the table pointers in registers
    uint32_t s = c.state, at = c.at;
    const uint32_t stop = c.stop;
    auto emit = [&](uint32_t pid, uint32_t end) { report<CP>(c, im, out, pid, end); };
    for (;;) {
        if (leftmost_flush<MODE>(c, s, at, emit)) continue;
        if (at >= stop && (MODE != kModeLeftmost || !c.have)) break;
        if (stop_hot && at >= min_at && (MODE != kModeLeftmost || !c.have) && (uint32_t)__ldg(hm.full2hot + s) < hm.hot_limit)
            break;
        scan_byte<MODE>(c, im, s, at, emit);
    }
    c.state = s;
    c.at = at;
}

// index of the haystack containing stream position p (offsets[h] <= p < offsets[h+1]); n when p is at/after the end
__device__ __forceinline__ int64_t find_haystack(const Batch &B, int64_t p) {
    int64_t lo = 0, hi = B.n_haystacks;  // answer in [lo, hi]
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(B.offsets + mid + 1) <= p)
            lo = mid + 1;
        else
            hi = mid;
    }
    return lo;
}

}  // namespace acb
