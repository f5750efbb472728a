// scan_staged.cuh -- the hot kernel: one lane per SEGMENT of the byte stream,
// hot table rows in shared memory, bytes staged through shared memory with
// cp.async.
//
// Layout per CTA (dynamic shared memory):
//   [ hot table : (H + 1) rows x n_cols u16 ]  rows 0..H-1 are the H hottest
//        states of the hot image (automaton.h: HotHeader; ranked by sampled
//        visit counts).  An entry is the BYTE OFFSET of the next state's row
//        inside this table when that state is hot and is neither a match state
//        nor the dead state, else the offset of row H, the TRAP row, which maps
//        everything to itself.  One add therefore forms the next shared-memory
//        address, a lane that left the hot set stays trapped, and ONE compare
//        per 16 bytes detects it; the 16 bytes are then redone by exact_scan
//        from the saved state.  Brought in with one TMA bulk copy
//        (cp.async.bulk + mbarrier) when the whole hot image fits.
//   [ column map : 256 B ]  (kColClass only)
//   [ mbarrier ]
//   [ copy metadata : per warp 32 x (first 16-byte unit, number of chunks) ]
//   [ staging : per warp, 2 buffers x 32 lanes x 80 B ]  lane l's 64-byte
//        chunk at a pitch of 80 bytes: consecutive lanes start 5 sixteen-byte
//        units apart, so the per-lane LDS.128 reads of a quarter warp hit 8
//        different bank groups (conflict free) with plain immediate offsets.
//
// Work decomposition (scan_core.cuh: SegPlan).  The stream is cut into
// fixed-size segments; a lane scans one segment, walking through whatever
// haystack boundaries fall inside it (each haystack start resets the scanner to
// the root, exactly).  A segment that begins inside a haystack begins with a
// speculated state: the lane first scans `warm` bytes before the segment from
// the root, silently, and takes the state it arrives with.  It records that
// state and its end state in SegInfo; repair.cuh verifies the chain and redoes
// the few places where the guess was wrong.  Lanes of a warp take segments
// `lane_stride` apart, which for a batch of equal-length haystacks puts all 32
// lanes at the same offset of 32 different haystacks (same text -> same table
// row -> shared-memory broadcasts instead of bank conflicts).
//
// Each lane reads its bytes in 64-byte chunks of the ABSOLUTE address grid,
// so every cp.async is 16-byte aligned and a warp-wide copy instruction
// touches 8 x 64 contiguous bytes.  Chunk k+1 is in flight while chunk k is
// scanned (the scan of a chunk takes longer than an HBM round trip).
#pragma once
#include <type_traits>

#include "scan_core.cuh"

namespace acb {

#ifndef ACB_MAX_WARPS
#define ACB_MAX_WARPS 32
#endif
constexpr int kMaxWarps = ACB_MAX_WARPS;  // per CTA (one CTA per SM): 32 -> 64 registers per thread, 28 -> 72, 24 -> 80
constexpr int kChunk = 64;                // bytes per lane per stage
constexpr int kRow = kChunk + 16;         // a lane's row in the staging buffer: 80-byte pitch = conflict-free LDS.128 without a swizzle
constexpr int kStageBytes = 32 * kRow;    // per warp per buffer
constexpr int kStageOffset = 256 + 128;   // column map + mbarrier slot, after the hot table
constexpr int kMetaBytes = 32 * 8;        // per warp: (first 16-byte unit, chunk count) of each lane, read by the copy issue

struct FastTab {
    const uint8_t *hot;   // shared: the table, addressed in bytes
    const uint8_t *cmap;  // shared: byte -> column (kColClass)
    uint32_t lo, maxc;
};

constexpr int kColAscii = 2;  // hot table indexed by the raw byte (128 entries per row); bytes >= 128 clamp to 127

template <int COLMODE, bool CLAMP = true>
__device__ __forceinline__ uint32_t fstep(uint32_t s, uint32_t b, const FastTab &f) {
    uint32_t col;
    if (COLMODE == kColRange)
        col = min(b - f.lo, f.maxc);
    else if (COLMODE == kColClass)
        col = (uint32_t)f.cmap[b];
    else
        col = CLAMP ? min(b, 127u) : b;
    return *reinterpret_cast<const uint16_t *>(f.hot + s + (col << 1));
}

template <int COLMODE, bool CLAMP = true>
__device__ __forceinline__ uint32_t fstep4(uint32_t s, uint32_t w, const FastTab &f) {
    s = fstep<COLMODE, CLAMP>(s, __byte_perm(w, 0, 0x4440), f);
    s = fstep<COLMODE, CLAMP>(s, __byte_perm(w, 0, 0x4441), f);
    s = fstep<COLMODE, CLAMP>(s, __byte_perm(w, 0, 0x4442), f);
    s = fstep<COLMODE, CLAMP>(s, __byte_perm(w, 0, 0x4443), f);
    return s;
}

__device__ __forceinline__ uint4 lds128_volatile(const uint8_t *p) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}

// continuation bytes (10xxxxxx) in a word
__device__ __forceinline__ uint32_t cont_bytes(uint32_t w) { return __popc(w & ~(w << 1) & 0x80808080u); }

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// ---- TMA bulk copy (global -> shared) completing on an mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}

// Views into the hot image (global memory).
struct DevHot {
    const uint16_t *table;
    const uint32_t *hot2full;
    const uint16_t *full2hot;
    uint32_t n_rows;
    const uint16_t *table128;  // kColAscii
    uint32_t n_rows128;
};

// What a lane knows about its segment besides the scanner state (kept out of
// the hot loop's registers: only touched at piece boundaries).
enum PieceKind : uint32_t { kPieceWarm = 0, kPieceHead = 1, kPieceNormal = 2 };
struct LaneSeg {
    int64_t org;         // stream position of c.base
    int64_t seg;         // segment index
    uint32_t lo_rel, hi_rel;  // the segment, relative to org
    uint32_t h;          // current haystack
    uint32_t kind;
    uint32_t spec_state, head_count;
    uint32_t done;
};

// The current piece is finished (c.at >= c.stop, nothing pending): move on.  Either starts the next
// piece (c.at = its first byte, c.state = its start state) or sets L.done and writes the segment summary.
template <int MODE, bool CP>
__device__ __noinline__ void advance_piece(PieceCtx &c, LaneSeg &L, const Batch &B, const Sink &out, SegInfo *seg_info) {
    bool finish_segment = false;
    if (L.kind == kPieceWarm) {
        const uint32_t he_rel = c.limit;
        const uint32_t piece_end = min(L.hi_rel, he_rel);
        if (c.at == L.lo_rel) {
            // arrived at the segment start with a guess for the state there: scan the head piece from it
            L.spec_state = c.state;
            L.kind = kPieceHead;
            c.stop = piece_end;
            c.emit_from = 0;
            c.cp_pos = c.at;
            c.cp_cont = 0;
            return;
        }
        // a leftmost match kept the scanner busy past the segment start: no usable guess.
        // Leave the whole head piece to the repair pass.
        L.spec_state = kSpecSkipped;
        L.kind = kPieceHead;
        c.at = piece_end;
        c.stop = piece_end;
        c.emit_from = 0;
        c.state = kRoot;
        c.have = 0;
        c.cp_pos = L.lo_rel;  // its continuation bytes still have to be counted for the segments after it
        c.cp_cont = 0;
        // fall through: the (skipped) head piece is finished
    }
    if (L.kind == kPieceHead) L.head_count = c.nemit;
    if (c.stop == c.limit) {
        // the piece ended with its haystack: continue with the next non-empty haystack, if it starts inside the segment
        int64_t h = (int64_t)L.h + 1;
        while (h < B.n_haystacks && __ldg(B.offsets + h + 1) == __ldg(B.offsets + h)) h++;
        const int64_t hi_pos = L.org + L.hi_rel;
        if (h < B.n_haystacks && __ldg(B.offsets + h) < hi_pos) {
            const int64_t hs = __ldg(B.offsets + h), he = __ldg(B.offsets + h + 1);
            L.h = (uint32_t)h;
            L.kind = kPieceNormal;
            c.at = (uint32_t)(hs - L.org);
            c.limit = (uint32_t)(he - L.org);
            c.stop = min(L.hi_rel, c.limit);
            c.emit_from = 0;
            c.state = kRoot;
            c.have = 0;
            c.hay = (uint32_t)h;
            c.hay_delta = (uint32_t)(L.org - hs);
            c.cp_pos = c.at;
            c.cp_cont = 0;
            return;
        }
        finish_segment = true;
    } else {
        finish_segment = true;  // the segment ends inside this haystack
    }
    if (finish_segment) {
        SegInfo si;
        si.spec_state = L.spec_state;
        si.end_state = c.state;
        si.end_over = c.at - L.hi_rel;
        si.head_count = L.head_count;
        si.drop = 0;
        si.cont_tail = 0;
        si.reserved0 = si.reserved1 = 0;
        if (CP) {
            // continuation bytes of the last piece inside the segment
            if (c.cp_pos <= L.hi_rel) {
                cp_catch_up(c, min(L.hi_rel, c.limit));
                si.cont_tail = c.cp_cont;
            } else {
                uint32_t n = c.cp_cont;
                for (uint32_t p = L.hi_rel; p < c.cp_pos; p++) n -= (ld_u8(c.base + p) & 0xC0u) == 0x80u;
                si.cont_tail = n;
            }
        }
        uint4 *dst = reinterpret_cast<uint4 *>(seg_info + L.seg);
        dst[0] = make_uint4(si.spec_state, si.end_state, si.end_over, si.head_count);
        dst[1] = make_uint4(si.drop, si.cont_tail, 0u, 0u);
        out.unit_counts[2 * L.seg] = 0;
        out.unit_counts[2 * L.seg + 1] = c.nemit;
        L.done = 1;
    }
}

// Runs the exact scanner until the lane is at a point where the fast path can take over
// (hot state, nothing pending, inside a piece) or the segment is finished.
template <int MODE, bool CP>
__device__ __noinline__ void settle(PieceCtx &c, LaneSeg &L, const DevImage &im, const Batch &B, const Sink &out,
                                    SegInfo *seg_info, HotMap hm, uint32_t min_at) {
    for (;;) {
        exact_scan<MODE, CP>(c, im, out, true, min_at, hm);
        if (c.at >= c.stop && (MODE != kModeLeftmost || !c.have)) {
            advance_piece<MODE, CP>(c, L, B, out, seg_info);
            if (L.done) return;
            min_at = c.at;
            continue;
        }
        return;
    }
}

template <int MODE, bool CP, int COLMODE>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
scan_staged_kernel(DevImage im, DevHot hot_img, Batch B, SegPlan P, Sink out, SegInfo *seg_info, uint32_t H,
                   uint32_t hot_bytes, unsigned int *task_counter, unsigned long long *trap_stats) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *hot = smem;
    uint8_t *cmap = smem + hot_bytes;                     // 256 B
    uint8_t *stage_all = smem + hot_bytes + kStageOffset;  // 128-aligned by construction
    const uint32_t row_entries = COLMODE == kColAscii ? kAsciiCols : im.n_cols;
    const uint32_t row_bytes = row_entries * 2;
    const uint32_t trap = H * row_bytes;
    const uint16_t *src_table = COLMODE == kColAscii ? hot_img.table128 : hot_img.table;

    // ---- prologue: the hot table (L2 resident) ------------------------------------------
    if (H == (COLMODE == kColAscii ? hot_img.n_rows128 : hot_img.n_rows)) {
        // the whole image fits: one TMA bulk copy, completion on an mbarrier
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem + hot_bytes + 256);
        const uint32_t bytes = ((H + 1) * row_bytes + 15u) & ~15u;
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_expect_tx(bar, bytes);
            tma_bulk_g2s((uint32_t)__cvta_generic_to_shared(hot), src_table, bytes, bar);
        }
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) cmap[i] = __ldg(im.colmap + i);
        if (COLMODE == kColAscii) {
            // guard row behind the trap row: the speculative (unclamped) pass may read up to 254 bytes past it
            uint16_t *h16 = reinterpret_cast<uint16_t *>(hot);
            for (uint32_t i = threadIdx.x; i < kAsciiCols; i += blockDim.x) h16[(H + 1) * kAsciiCols + i] = (uint16_t)trap;
        }
        __syncthreads();  // the barrier is initialised before anyone polls it
        mbar_wait(bar, 0);
    } else {
        // a prefix of the image (rows are hottest-first): entries beyond it become the trap
        uint16_t *h16 = reinterpret_cast<uint16_t *>(hot);
        const uint32_t n = H * row_entries;
        const uint32_t *src32 = reinterpret_cast<const uint32_t *>(src_table);  // 16-byte aligned in the image
        uint32_t *dst32 = reinterpret_cast<uint32_t *>(hot);
        for (uint32_t i = threadIdx.x; i < n / 2; i += blockDim.x) {
            const uint32_t v = __ldg(src32 + i);
            dst32[i] = min(v & 0xffffu, trap) | (min(v >> 16, trap) << 16);
        }
        if ((n & 1u) && threadIdx.x == 0) h16[n - 1] = (uint16_t)min((uint32_t)__ldg(src_table + n - 1), trap);
        for (uint32_t i = threadIdx.x; i < row_entries * (COLMODE == kColAscii ? 2u : 1u); i += blockDim.x) h16[n + i] = (uint16_t)trap;
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) cmap[i] = __ldg(im.colmap + i);
    }
    __syncthreads();

    HotMap hm;
    hm.full2hot = hot_img.full2hot;
    hm.hot_limit = H;
    FastTab ft;
    ft.hot = hot;
    ft.cmap = cmap;
    ft.lo = im.col_lo;
    ft.maxc = im.n_cols - 1;
    uint32_t n_groups = 0, n_traps = 0;

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint2 *meta = reinterpret_cast<uint2 *>(stage_all + (size_t)warp * kMetaBytes);
    uint8_t *stage = stage_all + (size_t)(blockDim.x >> 5) * kMetaBytes + (size_t)warp * 2 * kStageBytes;
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const uintptr_t gbase = reinterpret_cast<uintptr_t>(B.bytes + P.origin);  // 64-byte aligned by construction of the plan
    // copy instruction i of a stage moves 16-byte unit (lane & 3) of lane (i * 8 + lane / 4)
    const uint32_t cp_dst = stage_s + (lane >> 2) * kRow + (lane & 3) * 16;
    const uint64_t q = P.lane_stride;
    const uint64_t n_tasks = ((uint64_t)P.n_segments + 32 * q - 1) / (32 * q) * q;
    const int64_t stream_lo = __ldg(B.offsets), stream_hi = __ldg(B.offsets + B.n_haystacks);

    for (;;) {
        // ---- claim the next warp-task: 32 segments, lane_stride apart -----------------
        unsigned int task = 0;
        if (lane == 0) task = atomicAdd(task_counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= n_tasks) break;

        PieceCtx c;
        LaneSeg L;
        L.seg = (int64_t)(((uint64_t)(task / q) * 32 + lane) * q + task % q);
        L.done = 1;
        L.spec_state = kNoState;
        L.head_count = 0;
        uint32_t off16 = 0, nchunks = 0;
        uint32_t pos = 0, s = 0, stop = 0, cpd = 0;
        bool warm = false;  // the current piece is the silent warm-up before the segment
        const int64_t glo = P.origin + L.seg * (int64_t)P.seg_bytes;
        const int64_t lo = max(glo, stream_lo), hi = min(glo + (int64_t)P.seg_bytes, stream_hi);
        if (L.seg < P.n_segments && lo >= hi) {
            // a segment outside the stream (the plan is sized from the buffer length): nothing to scan
            uint4 *dst = reinterpret_cast<uint4 *>(seg_info + L.seg);
            dst[0] = make_uint4(kNoState, kRoot, 0u, 0u);
            dst[1] = make_uint4(0u, 0u, 0u, 0u);
            out.unit_counts[2 * L.seg] = 0;
            out.unit_counts[2 * L.seg + 1] = 0;
        } else if (L.seg < P.n_segments) {
            // the haystack containing lo: try the position an equal-length batch would put it at, else search
            int64_t h = P.avg_len ? (lo - stream_lo) / (int64_t)P.avg_len : 0;
            if (h >= B.n_haystacks) h = B.n_haystacks - 1;
            int64_t hs = __ldg(B.offsets + h), he = __ldg(B.offsets + h + 1);
            if (!(hs <= lo && lo < he)) {
                h = find_haystack(B, lo);
                hs = __ldg(B.offsets + h);
                he = __ldg(B.offsets + h + 1);
            }
            const bool cont = hs < lo;
            const int64_t w = cont ? max(hs, lo - (int64_t)P.warm) : lo;
            const uintptr_t pw = reinterpret_cast<uintptr_t>(B.bytes + w);
            const uintptr_t a0 = pw & ~uintptr_t(kChunk - 1);
            L.org = w - (int64_t)(pw - a0);
            L.lo_rel = (uint32_t)(lo - L.org);
            L.hi_rel = (uint32_t)(hi - L.org);
            L.h = (uint32_t)h;
            L.kind = cont ? kPieceWarm : kPieceNormal;
            L.done = 0;
            off16 = (uint32_t)((a0 - gbase) >> 4);
            nchunks = (L.hi_rel + kChunk - 1) / kChunk;
            c.base = B.bytes + L.org;
            c.at = (uint32_t)(w - L.org);
            c.limit = (uint32_t)(he - L.org);
            c.stop = cont ? L.lo_rel : min(L.hi_rel, c.limit);
            c.emit_from = cont ? 0xffffffffu : 0u;  // the warm-up reports nothing
            c.state = kRoot;
            c.have = 0;
            c.last_pid = c.last_end = 0;
            c.hay = (uint32_t)h;
            c.hay_delta = (uint32_t)(L.org - hs);
            c.unit = (uint32_t)(2 * L.seg + 1);
            c.nemit = 0;
            c.cp_pos = c.at;
            c.cp_cont = 0;
            warm = cont;
            if (c.at < c.stop) {
                // a piece always starts in the root state, which is hot row 0: straight into the fast path
                pos = c.at;
                stop = c.stop;
                s = 0;
            } else {
                settle<MODE, CP>(c, L, im, B, out, seg_info, hm, c.at);  // (cannot happen: segments are never empty here)
                pos = c.at;
                stop = c.stop;
                warm = !L.done && L.kind == kPieceWarm;
                if (!L.done) {
                    s = (uint32_t)__ldg(hot_img.full2hot + c.state) * row_bytes;
                    if (CP) {
                        cp_catch_up(c, pos);
                        cpd = c.cp_cont;
                    }
                }
            }
        }
        bool done = L.done != 0;
        if (done) nchunks = 0;
        uint32_t kmax = nchunks;
#pragma unroll
        for (int d = 16; d; d >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));
        // where each lane's bytes are: read back by whichever lane copies them (no shuffles in the loop:
        // the compiler cannot prove the warp converged there and would emit a slow collective path)
        __syncwarp();  // the previous task no longer reads meta / the staging buffers
        meta[lane] = make_uint2(off16, nchunks);
        __syncwarp();
        if (!done) n_groups += (L.hi_rel - c.at) >> 4;

        // stage chunk k into buffer BUF (compile-time: every shared-memory address below is base + immediate)
        auto issue = [&](auto buf_tag, uint32_t k) {
            constexpr uint32_t BUF = decltype(buf_tag)::value;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint2 m = meta[i * 8 + (lane >> 2)];
                const uint8_t *src = reinterpret_cast<const uint8_t *>(gbase) + ((size_t)(m.x + k * 4 + (lane & 3)) << 4);
                cp_async16(cp_dst + BUF * kStageBytes + i * 8 * kRow, src, k < m.y ? 16u : 0u);
            }
            cp_async_commit();
        };
        // hand the lane over to the exact scanner at position `pos`, come back at the next fast-resume point
        auto leave_fast = [&](uint32_t min_at) {
            c.state = __ldg(hot_img.hot2full + s / row_bytes);
            c.at = pos;
            if (CP) {
                c.cp_pos = pos;
                c.cp_cont = cpd;
            }
            settle<MODE, CP>(c, L, im, B, out, seg_info, hm, min_at);
            done = L.done != 0;
            pos = c.at;
            stop = c.stop;
            warm = !done && L.kind == kPieceWarm;
            if (!done) {
                s = (uint32_t)__ldg(hot_img.full2hot + c.state) * row_bytes;
                if (CP) {
                    cp_catch_up(c, pos);
                    cpd = c.cp_cont;
                }
            }
        };
        // the piece ended exactly where the fast path stands: the cheap, common transitions
        // (warm-up -> head piece; end of the segment) without going through the exact scanner
        auto piece_end_fast = [&]() -> bool {
            if (warm) {
                // arrived at the segment start in state s: that is the guess; scan the head piece from it
                L.spec_state = __ldg(hot_img.hot2full + s / row_bytes);
                L.kind = kPieceHead;
                stop = min(L.hi_rel, c.limit);
                c.stop = stop;
                c.emit_from = 0;
                cpd = 0;
                warm = false;
                return true;
            }
            if (stop == L.hi_rel) {
                // end of the segment: write the summary
                uint4 *dst = reinterpret_cast<uint4 *>(seg_info + L.seg);
                const uint32_t nem = c.nemit;
                dst[0] = make_uint4(L.spec_state, __ldg(hot_img.hot2full + s / row_bytes), 0u,
                                    L.kind == kPieceHead ? nem : L.head_count);
                dst[1] = make_uint4(0u, CP ? cpd : 0u, 0u, 0u);
                out.unit_counts[2 * L.seg] = 0;
                out.unit_counts[2 * L.seg + 1] = nem;
                L.done = 1;
                done = true;
                return true;
            }
            return false;
        };

        // everything that is not a clean whole chunk: 16-byte groups, then single bytes (one instance of this code)
        auto generic = [&](uint32_t relk, const uint8_t *row) {
#pragma unroll 1
            for (int j = 0; j < 4; j++) {
                const uint32_t g = relk + j * 16;
                if (done || pos < g || pos >= g + 16) continue;  // this lane is not inside this group
                if (pos == g && g + 16 <= stop) {
                    // a whole 16-byte group in the fast path
                    const uint4 w = *reinterpret_cast<const uint4 *>(row + j * 16);
                    uint32_t t = fstep4<COLMODE>(s, w.x, ft);
                    t = fstep4<COLMODE>(t, w.y, ft);
                    t = fstep4<COLMODE>(t, w.z, ft);
                    t = fstep4<COLMODE>(t, w.w, ft);
                    if (t != trap) {
                        s = t;
                        pos += 16;
                        if (CP) {
                            if ((w.x | w.y | w.z | w.w) & 0x80808080u)
                                cpd += cont_bytes(w.x) + cont_bytes(w.y) + cont_bytes(w.z) + cont_bytes(w.w);
                        }
                        continue;
                    }
                }
                // byte by byte through the hot table (bytes from the staged row): piece boundaries,
                // unaligned positions, and the group something happens in -- the exact scanner only
                // gets the byte that left the hot set
                bool trapped = false;
                while (!done && pos >= g && pos < g + 16) {
                    if (pos >= stop) {
                        if (!(pos == stop && piece_end_fast())) leave_fast(stop);
                        continue;
                    }
                    const uint32_t b = row[pos - relk];
                    const uint32_t t = fstep<COLMODE>(s, b, ft);
                    if (t != trap) {
                        s = t;
                        pos++;
                        if (CP) cpd += (b & 0xC0u) == 0x80u;
                    } else {
                        trapped = true;
                        leave_fast(pos + 1);
                    }
                }
                n_traps += trapped;
            }
        };
        using Buf0 = std::integral_constant<uint32_t, 0>;
        using Buf1 = std::integral_constant<uint32_t, 1>;
        const uint8_t *row0 = stage + lane * kRow;
        auto body = [&](auto buf_tag, uint32_t k) {
            constexpr uint32_t BUF = decltype(buf_tag)::value;
            cp_async_wait_all();
            __syncwarp();
            if (k + 1 < kmax) issue(std::integral_constant<uint32_t, 1 - BUF>{}, k + 1);
            const uint8_t *row = row0 + BUF * kStageBytes;
            const uint32_t relk = k * kChunk;
            if (!done && warm && pos == stop && pos == relk) piece_end_fast();  // the warm-up ended right at this chunk
            if (!done && pos == relk && relk + kChunk <= stop) {
                // ---- the whole 64-byte chunk in the fast path: one trap check for all of it ----
                uint32_t t = s, hb = 0;
                if (COLMODE == kColAscii) {
                    // raw-byte indexing, no clamp: speculative.  A byte >= 128 would index past its row
                    // (into the next rows / the guard row, never outside the table) and the result is thrown
                    // away: the OR of all words tells afterwards whether that happened.
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint4 w = *reinterpret_cast<const uint4 *>(row + j * 16);
                        t = fstep4<COLMODE, false>(t, w.x, ft);
                        t = fstep4<COLMODE, false>(t, w.y, ft);
                        t = fstep4<COLMODE, false>(t, w.z, ft);
                        t = fstep4<COLMODE, false>(t, w.w, ft);
                        hb |= w.x | w.y | w.z | w.w;
                    }
                    if (!(hb & 0x80808080u) && t != trap) {
                        s = t;
                        pos += kChunk;
                        return;
                    }
                    // high bytes or an event: group by group below
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const uint4 w = *reinterpret_cast<const uint4 *>(row + j * 16);
                        t = fstep4<COLMODE>(t, w.x, ft);
                        t = fstep4<COLMODE>(t, w.y, ft);
                        t = fstep4<COLMODE>(t, w.z, ft);
                        t = fstep4<COLMODE>(t, w.w, ft);
                        if (CP) hb |= w.x | w.y | w.z | w.w;
                    }
                    if (t != trap) {
                        s = t;
                        pos += kChunk;
                        if (CP && (hb & 0x80808080u)) {
                            // multi-byte characters in this chunk (rare in mostly-ASCII text): count their
                            // continuation bytes from the staged row again (volatile: do not keep 16 words live for this)
#pragma unroll
                            for (int j = 0; j < 16; j++)
                                cpd += cont_bytes(*reinterpret_cast<const volatile uint32_t *>(row + j * 4));
                        }
                        return;
                    }
                    // something happened in these 64 bytes: go through them group by group below
                    // (s and pos are untouched); only the group it happened in is redone exactly
                }
            }
            generic(relk, row);
        };
        if (kmax) issue(Buf0{}, 0);
        for (uint32_t k = 0; k < kmax; k += 2) {
            body(Buf0{}, k);
            if (k + 1 < kmax) body(Buf1{}, k + 1);
        }
        // the end of the segment (normally reached in the fast path, right at its last byte)
        while (!done) {
            if (!(pos == stop && piece_end_fast())) leave_fast(stop);
        }
    }
    // how well the hot set fits the data: the host re-profiles when traps are frequent
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        n_groups += __shfl_xor_sync(0xffffffffu, n_groups, d);
        n_traps += __shfl_xor_sync(0xffffffffu, n_traps, d);
    }
    if (lane == 0) {
        atomicAdd(trap_stats, (unsigned long long)n_groups);
        atomicAdd(trap_stats + 1, (unsigned long long)n_traps);
    }
}

}  // namespace acb
