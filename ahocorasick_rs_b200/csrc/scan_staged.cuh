// scan_staged.cuh -- the hot kernel: one lane per scan unit, hot table rows in
// shared memory, haystack bytes staged through shared memory with cp.async.
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
//   [ staging : per warp, 2 buffers x 32 lanes x 64 B ]  lane l's 64-byte
//        chunk, 16-byte units XOR-swizzled with (l >> 1) & 3 so the per-lane
//        LDS.128 reads are bank-conflict free.
//
// Each lane walks its own unit in 64-byte chunks of the ABSOLUTE address grid,
// so every cp.async is 16-byte aligned and a warp-wide copy instruction
// touches 8 x 64 contiguous bytes.  Chunk k+1 is in flight while chunk k is
// scanned (the scan of a chunk takes longer than an HBM round trip).
#pragma once
#include "scan_core.cuh"

namespace acb {

constexpr int kChunk = 64;                // bytes per lane per stage
constexpr int kStageBytes = 32 * kChunk;  // per warp per buffer
constexpr int kStageOffset = 256 + 128;   // column map + mbarrier slot, after the hot table

struct FastTab {
    const uint8_t *hot;   // shared: the table, addressed in bytes
    const uint8_t *cmap;  // shared: byte -> column (kColClass)
    uint32_t lo, maxc;
};

template <int COLMODE>
__device__ __forceinline__ uint32_t fstep(uint32_t s, uint32_t b, const FastTab &f) {
    const uint32_t col = (COLMODE == kColRange) ? min(b - f.lo, f.maxc) : (uint32_t)f.cmap[b];
    return *reinterpret_cast<const uint16_t *>(f.hot + s + (col << 1));
}

template <int COLMODE>
__device__ __forceinline__ uint32_t fstep4(uint32_t s, uint32_t w, const FastTab &f) {
    s = fstep<COLMODE>(s, __byte_perm(w, 0, 0x4440), f);
    s = fstep<COLMODE>(s, __byte_perm(w, 0, 0x4441), f);
    s = fstep<COLMODE>(s, __byte_perm(w, 0, 0x4442), f);
    s = fstep<COLMODE>(s, __byte_perm(w, 0, 0x4443), f);
    return s;
}

// continuation bytes (10xxxxxx) in a word
__device__ __forceinline__ uint32_t cont_bytes(uint32_t w) { return __popc(w & ~(w << 1) & 0x80808080u); }

// advance the code point counter to position `to` (no-op when already there or past it)
__device__ __forceinline__ void cp_catch_up(UnitCtx &c, uint32_t to) {
    uint32_t p = c.cp_pos, n = c.cp_count;
    while (p < to) {
        n += (ld_u8(c.base + p) & 0xC0u) != 0x80u;
        p++;
    }
    c.cp_pos = p;
    c.cp_count = n;
}

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
};

template <int MODE, bool CP, int COLMODE>
__global__ void __launch_bounds__(1024, 1)
scan_staged_kernel(DevImage im, DevHot hot_img, Units U, Sink out, uint32_t H, uint32_t hot_bytes,
                   unsigned int *task_counter, unsigned long long *trap_stats) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *hot = smem;
    uint8_t *cmap = smem + hot_bytes;                     // 256 B
    uint8_t *stage_all = smem + hot_bytes + kStageOffset;  // 128-aligned by construction
    const uint32_t row_bytes = im.n_cols * 2;
    const uint32_t trap = H * row_bytes;

    // ---- prologue: the hot table (L2 resident) ------------------------------------------
    if (H == hot_img.n_rows) {
        // the whole image fits: one TMA bulk copy, completion on an mbarrier
        const uint32_t bar = (uint32_t)__cvta_generic_to_shared(smem + hot_bytes + 256);
        const uint32_t bytes = ((H + 1) * row_bytes + 15u) & ~15u;
        if (threadIdx.x == 0) {
            mbar_init(bar, 1);
            mbar_expect_tx(bar, bytes);
            tma_bulk_g2s((uint32_t)__cvta_generic_to_shared(hot), hot_img.table, bytes, bar);
        }
        for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) cmap[i] = __ldg(im.colmap + i);
        __syncthreads();  // the barrier is initialised before anyone polls it
        mbar_wait(bar, 0);
    } else {
        // a prefix of the image (rows are hottest-first): entries beyond it become the trap
        uint16_t *h16 = reinterpret_cast<uint16_t *>(hot);
        const uint32_t n = H * im.n_cols;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) h16[i] = (uint16_t)min((uint32_t)__ldg(hot_img.table + i), trap);
        for (uint32_t i = threadIdx.x; i < im.n_cols; i += blockDim.x) h16[n + i] = (uint16_t)trap;
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
    uint8_t *stage = stage_all + (size_t)warp * 2 * kStageBytes;
    const uint32_t stage_s = (uint32_t)__cvta_generic_to_shared(stage);
    const uintptr_t gbase = reinterpret_cast<uintptr_t>(U.bytes) & ~uintptr_t(kChunk - 1);
    const uint32_t my_swz = (lane >> 1) & 3;

    for (;;) {
        // ---- claim the next 32 units -------------------------------------------
        unsigned int task = 0;
        if (lane == 0) task = atomicAdd(task_counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if ((int64_t)task * 32 >= U.n_units) break;

        UnitCtx c;
        const bool valid = init_unit<CP>(c, U, (int64_t)task * 32 + lane);
        uint32_t phase = 0, off16 = 0, nchunks = 0;
        uint32_t rel0 = 0;      // position (relative to c.base, mod 2^32) of the first byte of chunk 0
        uint32_t pos = 0, s = 0;
        uint32_t fast_last = 0;  // last position a 16-byte fast group may start at
        bool fast_ok = false;
        uint32_t cpd = 0;        // code points: pos - (code points before pos)
        uint32_t end = 0;
        if (valid) {
            const uintptr_t p0 = reinterpret_cast<uintptr_t>(c.base + c.at);
            const uintptr_t pe = reinterpret_cast<uintptr_t>(c.base + c.end);
            const uintptr_t a0 = p0 & ~uintptr_t(kChunk - 1);
            phase = (uint32_t)(0 - reinterpret_cast<uintptr_t>(c.base)) & 15u;
            off16 = (uint32_t)((a0 - gbase) >> 4);
            nchunks = (pe > a0) ? (uint32_t)((pe - a0 + kChunk - 1) / kChunk) : 0;
            rel0 = c.at - (uint32_t)(p0 - a0);
            // head: bytes before the first 16-byte boundary
            exact_scan<MODE, CP>(c, im, out, true, c.at, phase, hm);
            pos = c.at;
            end = c.end;
            s = (uint32_t)__ldg(hot_img.full2hot + c.state) * row_bytes;  // a hot row unless the unit is already finished
            if (CP) {
                cp_catch_up(c, pos);
                cpd = pos - c.cp_count;
            }
            fast_ok = end >= 16 && pos < end;
            fast_last = end - 16;
        }
        uint32_t kmax = nchunks;
#pragma unroll
        for (int d = 16; d; d >>= 1) kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, d));

        // who loads what: copy instruction i moves 16-byte unit (i*32+lane)&3 of lane (i*32+lane)>>2
        uint32_t src_off16[4], src_nch[4], dst_off[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t idx = i * 32 + lane, ch = idx >> 2, un = idx & 3;
            src_off16[i] = __shfl_sync(0xffffffffu, off16, ch) + un;
            src_nch[i] = __shfl_sync(0xffffffffu, nchunks, ch);
            dst_off[i] = ch * kChunk + ((un ^ ((ch >> 1) & 3)) << 4);
        }
        auto issue = [&](uint32_t k) {
            const uint32_t buf = stage_s + (k & 1) * kStageBytes;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint8_t *src = reinterpret_cast<const uint8_t *>(gbase) + (((size_t)src_off16[i] + (size_t)k * 4) << 4);
                cp_async16(buf + dst_off[i], src, k < src_nch[i] ? 16u : 0u);
            }
            cp_async_commit();
        };

        __syncwarp();  // previous task's readers are done with both buffers
        if (kmax) issue(0);
        for (uint32_t k = 0; k < kmax; k++) {
            cp_async_wait_all();
            __syncwarp();
            if (k + 1 < kmax) issue(k + 1);
            const uint8_t *buf = stage + (k & 1) * kStageBytes + lane * kChunk;
            const uint32_t relk = rel0 + k * kChunk;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint32_t g = relk + j * 16;  // wraps for positions before the unit: never equal to pos then
                if (fast_ok && g == pos && g <= fast_last) {
                    const uint4 w = *reinterpret_cast<const uint4 *>(buf + ((j ^ my_swz) << 4));
                    uint32_t t = fstep4<COLMODE>(s, w.x, ft);
                    t = fstep4<COLMODE>(t, w.y, ft);
                    t = fstep4<COLMODE>(t, w.z, ft);
                    t = fstep4<COLMODE>(t, w.w, ft);
                    n_groups++;
                    if (t != trap) {
                        s = t;
                        pos += 16;
                        if (CP) {
                            if ((w.x | w.y | w.z | w.w) & 0x80808080u)
                                cpd += cont_bytes(w.x) + cont_bytes(w.y) + cont_bytes(w.z) + cont_bytes(w.w);
                        }
                    } else {
                        // something happened in these 16 bytes: redo them exactly
                        n_traps++;
                        c.state = __ldg(hot_img.hot2full + s / row_bytes);
                        c.at = pos;
                        if (CP) {
                            c.cp_pos = pos;
                            c.cp_count = pos - cpd;
                        }
                        exact_scan<MODE, CP>(c, im, out, true, pos + 16, phase, hm);
                        s = (uint32_t)__ldg(hot_img.full2hot + c.state) * row_bytes;
                        pos = c.at;
                        if (CP) {
                            cp_catch_up(c, pos);
                            cpd = pos - c.cp_count;
                        }
                        fast_ok = pos < end;
                    }
                }
            }
        }
        if (valid) {
            // tail: whatever is left after the last full 16-byte group
            if (pos < end) {
                c.state = __ldg(hot_img.hot2full + s / row_bytes);
                c.at = pos;
                if (CP) {
                    c.cp_pos = pos;
                    c.cp_count = pos - cpd;
                }
                exact_scan<MODE, CP>(c, im, out, false, 0, 0, hm);
            }
            out.unit_counts[c.unit] = c.nemit;
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
