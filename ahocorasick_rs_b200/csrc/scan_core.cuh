// scan_core.cuh -- device-side pieces shared by the scan kernels.
//
// The exact (byte-at-a-time) scanner below is the definition of the result:
// it is the reference's iterator drain (src/lib.rs:42-68, 238-248, 433; the
// crate's find / find_overlapping loops) run by ONE thread over ONE scan unit,
// reading the dense table in global memory.  The staged kernel (scan_staged.cuh)
// only accelerates the stretches where nothing can happen and drops back here
// for everything else, so both kernels produce the same matches in the same
// order by construction.
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

// What to scan.  Batch: unit u = haystack u = bytes[offsets[u], offsets[u+1]).
// Chunked: one haystack of `len` bytes, unit u emits matches that END in
// (u*chunk, (u+1)*chunk] and starts reading `halo` bytes earlier.
struct Units {
    const uint8_t *bytes;
    const int64_t *offsets;  // batch only
    int64_t n_units;
    uint64_t len;            // chunked only
    uint32_t chunk, halo;    // chunk == 0 => batch
    const uint64_t *chunk_cp;  // chunked + codepoints: code points before each chunk
};

struct Sink {
    acb_match *raw;
    uint32_t *raw_seq;
    uint32_t *raw_unit;
    unsigned long long cap;
    uint32_t *unit_counts;
    unsigned long long *total;
};

// Per-unit scanner state.  Positions are relative to `base` (the haystack's
// first byte), so they fit 32 bits for haystacks below 4 GiB.
struct UnitCtx {
    const uint8_t *base;
    uint32_t at;         // next byte to read
    uint32_t end;        // one past the last byte of the unit
    uint32_t emit_from;  // report only matches with end > emit_from (halo suppression)
    uint32_t state;
    uint32_t nemit;      // matches reported so far = rank of the next one inside the unit
    uint32_t unit, hay;
    // leftmost: the match that will be reported once the automaton dies / input ends
    uint32_t have, last_pid, last_end;
    // code points: number of non-continuation bytes in [base, cp_pos)
    uint32_t cp_pos, cp_count;
};

__device__ __forceinline__ uint32_t ld_u8(const uint8_t *p) { return __ldg(p); }

template <bool CP>
__device__ __forceinline__ void report(UnitCtx &c, const DevImage &im, const Sink &out, uint32_t pid, uint32_t end) {
    if (end <= c.emit_from) return;
    uint32_t start = end - __ldg(im.pat_len + pid);
    if (CP) {
        // ends are reported in non-decreasing order, so one forward-only counter is enough
        uint32_t pos = c.cp_pos, cnt = c.cp_count;
        while (pos < end) {
            cnt += (ld_u8(c.base + pos) & 0xC0u) != 0x80u;
            pos++;
        }
        c.cp_pos = pos;
        c.cp_count = cnt;
        end = cnt;
        start = cnt - __ldg(im.pat_cplen + pid);
    }
    unsigned long long i = atomicAdd(out.total, 1ULL);
    if (i < out.cap) {
        acb_match m;
        m.haystack = c.hay;
        m.pattern = pid;
        m.start = start;
        m.end = end;
        *reinterpret_cast<uint4 *>(out.raw + i) = *reinterpret_cast<uint4 *>(&m);
        out.raw_seq[i] = c.nemit;
        out.raw_unit[i] = c.unit;
    }
    c.nemit++;
}

// Where the staged fast path may resume: full2hot[state] < hot_limit.
struct HotMap {
    const uint16_t *full2hot;
    uint32_t hot_limit;
};

// Runs the exact scanner from c.at.  It returns when the unit is finished
// (c.at == c.end with nothing pending), or -- if stop_hot is set -- as soon as
//   at >= min_at, (at - phase) % 16 == 0, the state is hot and nothing is pending,
// i.e. at a point where the staged fast path may take over again.
template <int MODE, bool CP>
__device__ __noinline__ void exact_scan(UnitCtx &c, const DevImage &im, const Sink &out, bool stop_hot,
                                        uint32_t min_at, uint32_t phase, HotMap hm) {
    uint32_t s = c.state, at = c.at;
    const uint32_t end = c.end;
    for (;;) {
        if (MODE == kModeLeftmost) {
            if (at == end || s == kDead) {
                if (c.have) {
                    // the crate's iterator: report, then search again from the match end
                    report<CP>(c, im, out, c.last_pid, c.last_end);
                    at = c.last_end;
                    c.have = 0;
                    s = kRoot;
                    continue;
                }
                if (at == end) break;
                s = kRoot;  // unreachable: the dead state is only entered below a match state
            }
        } else {
            if (at == end) break;
        }
        if (stop_hot && at >= min_at && ((at - phase) & 15u) == 0 && (MODE != kModeLeftmost || !c.have) &&
            (uint32_t)__ldg(hm.full2hot + s) < hm.hot_limit)
            break;
        const uint32_t col = __ldg(im.colmap + ld_u8(c.base + at));
        const uint32_t e = __ldg(im.trans + (size_t)s * im.n_cols + col);
        s = e & kStateMask;
        at++;
        if (e & kMatchFlag) {
            const uint32_t m0 = __ldg(im.match_off + s);
            if (MODE == kModeStandard) {
                report<CP>(c, im, out, __ldg(im.match_pid + m0), at);
                s = kRoot;
            } else if (MODE == kModeLeftmost) {
                c.have = 1;
                c.last_pid = __ldg(im.match_pid + m0);
                c.last_end = at;
            } else {
                const uint32_t m1 = __ldg(im.match_off + s + 1);
                for (uint32_t k = m0; k < m1; k++) report<CP>(c, im, out, __ldg(im.match_pid + k), at);
            }
        }
    }
    c.state = s;
    c.at = at;
}

// Fills in a unit's context.  Returns false when u is out of range.
template <bool CP>
__device__ __forceinline__ bool init_unit(UnitCtx &c, const Units &U, int64_t u) {
    if (u >= U.n_units) return false;
    c.unit = (uint32_t)u;
    c.state = kRoot;
    c.nemit = 0;
    c.have = 0;
    c.last_pid = 0;
    c.last_end = 0;
    if (U.chunk == 0) {
        const int64_t b = U.offsets[u], e = U.offsets[u + 1];
        c.base = U.bytes + b;
        c.at = 0;
        c.end = (uint32_t)(e - b);
        c.emit_from = 0;
        c.hay = (uint32_t)u;
        c.cp_pos = 0;
        c.cp_count = 0;
    } else {
        const uint64_t lo = (uint64_t)u * U.chunk;
        uint64_t hi = lo + U.chunk;
        if (hi > U.len) hi = U.len;
        c.base = U.bytes;
        c.at = (uint32_t)(lo > U.halo ? lo - U.halo : 0);
        c.end = (uint32_t)hi;
        c.emit_from = (uint32_t)lo;
        c.hay = 0;
        // code points before the first byte this unit READS (the halo starts before the chunk)
        c.cp_pos = c.at;
        c.cp_count = 0;
        if (CP) {
            uint32_t n = (uint32_t)U.chunk_cp[u];
            for (uint32_t p = c.at; p < (uint32_t)lo; p++) n -= (ld_u8(c.base + p) & 0xC0u) != 0x80u;
            c.cp_count = n;
        }
    }
    return true;
}

}  // namespace acb
