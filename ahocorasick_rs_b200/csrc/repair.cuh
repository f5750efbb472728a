// repair.cuh -- validation of the speculative segment starts, and repair.
//
// The scan kernel starts every segment that begins inside a haystack from a
// GUESSED state (scan_staged.cuh).  For non-overlapping searches that guess can
// be wrong when a match ends just before the segment (the true scanner had
// restarted from the root there).  One thread per haystack walks the haystack's
// segment boundaries in order:
//
//   * boundary clean (the previous segment stopped exactly at the boundary in
//     the state this segment assumed): nothing to do, the segment is valid and
//     so is its recorded end state;
//   * otherwise it runs TWO exact scanners side by side from the boundary: T
//     from the true state (reports real matches into the segment's slot 0) and P
//     from the guessed state (only counts, replaying what the scan kernel
//     reported into slot 1).  As soon as both are at the same position in the
//     same state with nothing pending, everything P did from there on is valid:
//     the first `drop` matches of slot 1 are discarded and replaced by T's.  If
//     P reaches the end of its segment first, all of its matches are discarded
//     and T carries on into the next segment.
//
// Overlapping searches never need this (the state is a pure function of the
// last max_pattern_len bytes, which the warm-up covers).
#pragma once
#include "scan_core.cuh"

namespace acb {

template <int MODE>
__device__ __forceinline__ bool piece_finished(const PieceCtx &m, uint32_t stop) {
    return m.at >= stop && (MODE != kModeLeftmost || !m.have);
}

// one step of an exact scanner: either the leftmost flush/restart or one byte
template <int MODE, typename Emit>
__device__ __forceinline__ void machine_step(PieceCtx &m, const DevImage &im, Emit &&emit) {
    uint32_t s = m.state, at = m.at;
    if (!leftmost_flush<MODE>(m, s, at, emit)) {
        if (at < m.limit) scan_byte<MODE>(m, im, s, at, emit);
    }
    m.state = s;
    m.at = at;
}

template <int MODE, bool CP>
__device__ __forceinline__ void repair_body(const DevImage &im, const Batch &B, const SegPlan &P, const Sink &out,
                                            SegInfo *seg_info, unsigned long long *stats) {
    const int64_t stream_hi = __ldg(B.offsets + B.n_haystacks);
    for (int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; h < B.n_haystacks; h += (int64_t)gridDim.x * blockDim.x) {
        const int64_t hs = __ldg(B.offsets + h), he = __ldg(B.offsets + h + 1);
        if (he <= hs) continue;
        const int64_t S = P.seg_bytes;
        const int64_t js = (hs - P.origin) / S, je = (he - 1 - P.origin) / S;
        if (js == je) continue;
        const uint32_t limit = (uint32_t)(he - hs);

        // T: the true scanner, positions relative to the haystack start
        PieceCtx T;
        T.base = B.bytes + hs;
        T.limit = limit;
        T.emit_from = 0;
        T.hay = (uint32_t)h;
        T.hay_delta = 0;
        T.have = 0;
        T.last_pid = T.last_end = 0;
        int64_t slot_seg = -1;  // segment whose slot 0 T is currently filling
        uint32_t slot_seq = 0;
        auto close_slot = [&]() {
            if (slot_seg >= 0) out.unit_counts[2 * slot_seg] = slot_seq;
        };
        auto emit_true = [&](uint32_t pid, uint32_t end) {
            const int64_t m = (hs + (int64_t)end - 1 - P.origin) / S;  // the segment the match ends in
            if (m != slot_seg) {
                close_slot();
                slot_seg = m;
                slot_seq = 0;
            }
            uint32_t aux = 0;
            if (CP) {
                // continuation bytes between the segment start and the match end
                const int64_t lo_m = max(P.origin + m * S, hs);
                for (uint32_t p = (uint32_t)(lo_m - hs); p < end; p++) aux += (ld_u8(T.base + p) & 0xC0u) == 0x80u;
            }
            const unsigned long long i = atomicAdd(out.raw_total, 1ULL);
            if (i < out.cap) {
                acb_match mm;
                mm.haystack = (uint32_t)h;
                mm.pattern = pid;
                mm.start = end - __ldg(im.pat_len + pid);
                mm.end = end;
                *reinterpret_cast<uint4 *>(out.raw + i) = *reinterpret_cast<uint4 *>(&mm);
                out.raw_seq[i] = slot_seq;
                out.raw_unit[i] = (uint32_t)(2 * m);
                if (CP) out.raw_aux[i] = aux;
            }
            slot_seq++;
        };

        int64_t k = js + 1;
        while (k <= je) {
            const uint4 prev = *reinterpret_cast<const uint4 *>(seg_info + k - 1);  // spec, end_state, end_over, head_count
            const uint32_t spec = seg_info[k].spec_state;
            if (prev.z == 0 && spec == prev.y) {
                k++;  // clean: segment k started from the true state
                continue;
            }
            // ---- dirty boundary: repair from the true condition at the end of segment k-1 ----
            atomicAdd(stats, 1ULL);
            T.at = (uint32_t)(P.origin + k * S - hs) + prev.z;
            T.state = prev.y;
            T.have = 0;
            int64_t cur = k;
            for (;;) {
                const uint4 info = *reinterpret_cast<const uint4 *>(seg_info + cur);
                const int64_t lo_c = P.origin + cur * S;
                const int64_t hi_c = min(lo_c + S, stream_hi);
                const uint32_t pstop = (uint32_t)(min(hi_c, he) - hs);
                const bool p_alive = info.x != kSpecSkipped;
                PieceCtx Pm;
                Pm.base = T.base;
                Pm.limit = limit;
                Pm.at = (uint32_t)(lo_c - hs);
                Pm.state = info.x;
                Pm.have = 0;
                Pm.last_pid = Pm.last_end = 0;
                uint32_t d = 0;
                auto count_spec = [&](uint32_t, uint32_t) { d++; };
                bool converged = false;
                while (p_alive && !piece_finished<MODE>(Pm, pstop)) {
                    if (T.at == Pm.at && T.state == Pm.state && !T.have && !Pm.have) {
                        converged = true;
                        break;
                    }
                    const bool t_done = T.at >= limit && !T.have;
                    if (!t_done && T.at <= Pm.at)
                        machine_step<MODE>(T, im, emit_true);
                    else
                        machine_step<MODE>(Pm, im, count_spec);
                }
                if (converged) {
                    seg_info[cur].drop = d;
                    out.unit_counts[2 * cur + 1] -= d;
                    k = cur + 1;  // from here on segment cur is what the scan kernel recorded
                    break;
                }
                // the guess never met the truth inside this segment: all of its head-piece matches go
                seg_info[cur].drop = info.w;
                out.unit_counts[2 * cur + 1] -= info.w;
                while (!piece_finished<MODE>(T, pstop)) machine_step<MODE>(T, im, emit_true);
                if (pstop == limit) {
                    k = je + 1;  // T has finished the haystack
                    break;
                }
                cur++;
                // does the next segment's guess agree with the truth?  then it is valid as recorded
                if (T.at == (uint32_t)(P.origin + cur * S - hs) && !T.have && seg_info[cur].spec_state == T.state) {
                    k = cur + 1;
                    break;
                }
            }
            close_slot();
            slot_seg = -1;
        }
    }
}

}  // namespace acb
