// capi.cu -- the C ABI (include/acb200.h): kernel dispatch + ordering passes.
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "scan_staged.cuh"

namespace acb {

// ---------------------------------------------------------------------------
// plain kernel: the exact scanner over whole units, table in global memory
// ---------------------------------------------------------------------------
template <int MODE, bool CP>
__global__ void __launch_bounds__(128) scan_plain_kernel(DevImage im, Units U, Sink out) {
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < U.n_units; u += (int64_t)gridDim.x * blockDim.x) {
        UnitCtx c;
        init_unit<CP>(c, U, u);
        exact_scan<MODE, CP>(c, im, out, false, 0, 0, HotMap{nullptr, 0});
        out.unit_counts[u] = c.nemit;
    }
}

// ---------------------------------------------------------------------------
// profile kernel: walks a sample of the input through the dense table and
// counts state visits; the host ranks states by these counts to choose the rows
// the staged kernel keeps in shared memory (automaton.cpp: build_hot_image)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
profile_kernel(DevImage im, Units U, uint32_t *visits, int64_t n_samples, uint32_t max_bytes, int restart_on_match) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_samples) return;
    const uint8_t *p;
    uint64_t len;
    if (U.chunk == 0) {
        const int64_t u = (U.n_units * i) / n_samples;
        p = U.bytes + U.offsets[u];
        len = (uint64_t)(U.offsets[u + 1] - U.offsets[u]);
    } else {
        const uint64_t lo = (U.len / (uint64_t)n_samples) * (uint64_t)i;
        p = U.bytes + lo;
        len = U.len - lo;
    }
    if (len > max_bytes) len = max_bytes;
    uint32_t s = kRoot;
    for (uint64_t k = 0; k < len; k++) {
        const uint32_t e = __ldg(im.trans + (size_t)s * im.n_cols + __ldg(im.colmap + __ldg(p + k)));
        s = e & kStateMask;
        if (s == kDead || (restart_on_match && (e & kMatchFlag))) s = kRoot;
        atomicAdd(visits + s, 1u);
    }
}

// ---------------------------------------------------------------------------
// exclusive prefix sum u32[n] -> u64[n+1] (three small kernels)
// ---------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long *total) {
    __shared__ unsigned long long warp_sums[kScanThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < kScanThreads / 32; d <<= 1) {
            unsigned long long y = __shfl_up_sync(0xffffffffu, w, d);
            if (lane >= d) w += y;
        }
        if (lane < kScanThreads / 32) warp_sums[lane] = w;
    }
    __syncthreads();
    const unsigned long long before = warp ? warp_sums[warp - 1] : 0;
    *total = warp_sums[kScanThreads / 32 - 1];
    __syncthreads();
    return before + x - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums(const uint32_t *in, uint64_t n, unsigned long long *tile_sums) {
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    unsigned long long v = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++)
        if (base + i < n) v += in[base + i];
    unsigned long long total;
    block_exclusive_scan(v, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_offsets(unsigned long long *tile_sums, uint64_t n_tiles) {
    unsigned long long carry = 0;
    for (uint64_t base = 0; base < n_tiles; base += kScanThreads) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long v = i < n_tiles ? tile_sums[i] : 0;
        unsigned long long total;
        const unsigned long long ex = block_exclusive_scan(v, &total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(kScanThreads)
scan_apply(const uint32_t *in, uint64_t n, const unsigned long long *tile_offs, unsigned long long *out) {
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    unsigned long long vals[kScanItems], v = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        vals[i] = base + i < n ? in[base + i] : 0;
        v += vals[i];
    }
    unsigned long long total;
    unsigned long long run = tile_offs[blockIdx.x] + block_exclusive_scan(v, &total);
#pragma unroll
    for (int i = 0; i < kScanItems; i++) {
        if (base + i < n) out[base + i] = run;
        run += vals[i];
        if (base + i + 1 == n) out[n] = run;
    }
}

// final placement: match i of unit u with rank r goes to unit_offsets[u] + r
__global__ void __launch_bounds__(256)
order_matches_kernel(const acb_match *raw, const uint32_t *raw_seq, const uint32_t *raw_unit, unsigned long long raw_cap,
                     const unsigned long long *total, const unsigned long long *unit_offsets, acb_match *out,
                     unsigned long long out_cap) {
    unsigned long long n = *total;
    if (n > raw_cap) n = raw_cap;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long dst = unit_offsets[raw_unit[i]] + raw_seq[i];
        if (dst < out_cap) reinterpret_cast<uint4 *>(out)[dst] = reinterpret_cast<const uint4 *>(raw)[i];
    }
}

__global__ void finish_total_kernel(unsigned long long *total, unsigned long long raw_cap, unsigned long long out_cap) {
    const unsigned long long n = total[0];
    total[1] = (n <= raw_cap && n <= out_cap) ? n : 0;  // [1] = matches delivered in dev_out (0 = incomplete, retry)
}

// code points (non-continuation bytes) per chunk, for the chunked scan with codepoints
__global__ void __launch_bounds__(256) chunk_cp_count_kernel(const uint8_t *bytes, uint64_t len, uint32_t chunk, uint32_t *counts, uint64_t n_chunks) {
    // one warp per chunk
    const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (warp >= n_chunks) return;
    const uint64_t lo = warp * chunk;
    uint64_t hi = lo + chunk;
    if (hi > len) hi = len;
    uint32_t n = 0;
    for (uint64_t p = lo + lane; p < hi; p += 32) n += (__ldg(bytes + p) & 0xC0u) != 0x80u;
#pragma unroll
    for (int d = 16; d; d >>= 1) n += __shfl_xor_sync(0xffffffffu, n, d);
    if (lane == 0) counts[warp] = n;
}

}  // namespace acb

// ===========================================================================
// C ABI
// ===========================================================================
using namespace acb;

struct acb_automaton {
    Automaton *impl;
};

static thread_local std::string g_err;
static unsigned long long g_launches = 0;
static acb_tuning g_tuning = {0, 0, 0};

// optional device timing of the dominant (scan) kernel, for bench.py's roofline
static bool g_timing = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_timing_events;

static int fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}

#define CUDA_OK(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess) return fail(ACB_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

extern "C" {

const char *acb_last_error(void) { return g_err.c_str(); }
const char *acb_version(void) { return "acb200 0.1 (sm_100a)"; }
uint64_t acb_launch_count(void) { return g_launches; }

int acb_timing_enable(int on) {
    g_timing = on != 0;
    return ACB_OK;
}

int acb_timing_read(double *total_ms, uint64_t *n_scans) {
    double tot = 0;
    uint64_t n = 0;
    for (auto &p : g_timing_events) {
        float ms = 0;
        CUDA_OK(cudaEventSynchronize(p.second));
        CUDA_OK(cudaEventElapsedTime(&ms, p.first, p.second));
        tot += ms;
        n++;
        cudaEventDestroy(p.first);
        cudaEventDestroy(p.second);
    }
    g_timing_events.clear();
    if (total_ms) *total_ms = tot;
    if (n_scans) *n_scans = n;
    return ACB_OK;
}

int acb_set_tuning(const acb_tuning *t) {
    if (!t) return fail(ACB_EINVAL, "null tuning");
    g_tuning = *t;
    return ACB_OK;
}

int acb_build(const uint8_t *blob, const uint64_t *offsets, uint64_t n, int match_kind, int implementation,
              acb_automaton **out) {
    if (!out || !offsets || (!blob && n && offsets[n] != 0)) return fail(ACB_EINVAL, "null argument");
    if (implementation < -1 || implementation > 2) return fail(ACB_EINVAL, "unknown implementation");
    try {
        Automaton *impl = build_automaton(blob, offsets, n, match_kind, implementation);
        *out = new acb_automaton{impl};
        return ACB_OK;
    } catch (const std::exception &e) {
        return fail(ACB_EBUILD, e.what());
    }
}

void acb_free(acb_automaton *a) {
    if (!a) return;
    delete a->impl;
    delete a;
}

uint64_t acb_num_patterns(const acb_automaton *a) { return a->impl->hdr.n_patterns; }
uint64_t acb_num_states(const acb_automaton *a) { return a->impl->hdr.n_states; }
uint32_t acb_num_columns(const acb_automaton *a) { return a->impl->hdr.n_cols; }
uint32_t acb_max_pattern_len(const acb_automaton *a) { return a->impl->hdr.max_pat_len; }
uint32_t acb_min_pattern_len(const acb_automaton *a) { return a->impl->hdr.min_pat_len; }
int acb_match_kind(const acb_automaton *a) { return (int)a->impl->hdr.match_kind; }
uint64_t acb_image_bytes(const acb_automaton *a) { return a->impl->hdr.total_bytes; }

int acb_image_write(const acb_automaton *a, void *host_dst, uint64_t dst_bytes) {
    if (!a || !host_dst) return fail(ACB_EINVAL, "null argument");
    if (dst_bytes < a->impl->hdr.total_bytes) return fail(ACB_ECAPACITY, "image buffer too small");
    std::memcpy(host_dst, a->impl->image.data(), a->impl->hdr.total_bytes);
    return ACB_OK;
}

uint64_t acb_hot_bytes(const acb_automaton *a, uint32_t max_rows) { return hot_image_bytes(*a->impl, max_rows); }

int acb_hot_build(const acb_automaton *a, const uint32_t *host_visits, uint32_t max_rows, void *host_dst, uint64_t dst_bytes) {
    if (!a || !host_dst) return fail(ACB_EINVAL, "null argument");
    if (dst_bytes < hot_image_bytes(*a->impl, max_rows)) return fail(ACB_ECAPACITY, "hot image buffer too small");
    build_hot_image(*a->impl, host_visits, max_rows, static_cast<uint8_t *>(host_dst));
    return ACB_OK;
}

uint32_t acb_hot_rows(const void *host_hot) {
    const HotHeader *h = static_cast<const HotHeader *>(host_hot);
    return (h && h->magic == kHotMagic) ? h->n_rows : 0;
}

uint64_t acb_scratch_words(uint64_t n_units) {
    const uint64_t tiles = (n_units + kScanTile - 1) / kScanTile;
    // [0] task counter | tile sums | chunk code point counts (u32, n_units) | chunk code point offsets (n_units + 1)
    return 2 + tiles + 1 + (n_units + 1) / 2 + 1 + n_units + 1;
}

uint64_t acb_chunk_count(uint64_t len, uint32_t chunk_bytes) {
    if (!chunk_bytes) return 0;
    return (len + chunk_bytes - 1) / chunk_bytes;
}

}  // extern "C"

namespace {

struct DeviceInfo {
    int device = -1;
    int sms = 0;
    int max_smem_optin = 0;
};

int device_info(DeviceInfo &d) {
    static thread_local DeviceInfo cache;
    int dev;
    CUDA_OK(cudaGetDevice(&dev));
    if (cache.device != dev) {
        cache.device = dev;
        CUDA_OK(cudaDeviceGetAttribute(&cache.sms, cudaDevAttrMultiProcessorCount, dev));
        CUDA_OK(cudaDeviceGetAttribute(&cache.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    }
    d = cache;
    return ACB_OK;
}

DevImage make_view(const ImageHeader &h, const void *dev_image) {
    const uint8_t *b = static_cast<const uint8_t *>(dev_image);
    DevImage im;
    im.colmap = b + h.off_colmap;
    im.trans = reinterpret_cast<const uint32_t *>(b + h.off_trans);
    im.match_off = reinterpret_cast<const uint32_t *>(b + h.off_match_off);
    im.match_pid = reinterpret_cast<const uint32_t *>(b + h.off_match_pid);
    im.pat_len = reinterpret_cast<const uint32_t *>(b + h.off_pat_len);
    im.pat_cplen = reinterpret_cast<const uint32_t *>(b + h.off_pat_cplen);
    im.n_cols = h.n_cols;
    im.col_lo = h.col_lo;
    im.n_states = h.n_states;
    im.col_mode = h.col_mode;
    return im;
}

template <int MODE, bool CP>
int launch_plain(const DevImage &im, const Units &U, const Sink &out, const DeviceInfo &d, cudaStream_t st) {
    const int threads = 128;
    int64_t blocks = (U.n_units + threads - 1) / threads;
    const int64_t cap = (int64_t)d.sms * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    scan_plain_kernel<MODE, CP><<<(unsigned)blocks, threads, 0, st>>>(im, U, out);
    g_launches++;
    return ACB_OK;
}

template <int MODE, bool CP, int COLMODE>
int launch_staged(const DevImage &im, const DevHot &hot, const Units &U, const Sink &out, const DeviceInfo &d,
                  unsigned int *task_counter, cudaStream_t st) {
    auto kern = scan_staged_kernel<MODE, CP, COLMODE>;
    const int64_t tasks = (U.n_units + 31) / 32;
    const int ctas = d.sms;
    int warps = (int)((tasks + ctas - 1) / ctas);
    if (warps < 4) warps = 4;
    if (warps > 32) warps = 32;
    const uint32_t row_bytes = im.n_cols * 2;
    const uint32_t stage_bytes = (uint32_t)warps * 2 * kStageBytes;
    const uint32_t budget = (uint32_t)d.max_smem_optin;
    if (budget < stage_bytes + kStageOffset + 3 * row_bytes + 128) return fail(ACB_ECUDA, "not enough shared memory for the staged kernel");
    uint32_t rows = (budget - stage_bytes - kStageOffset - 128) / row_bytes;  // includes the trap row
    uint32_t H = rows - 1;
    if (H > hot.n_rows) H = hot.n_rows;
    if (g_tuning.hot_rows > 0 && (uint32_t)g_tuning.hot_rows < H) H = (uint32_t)g_tuning.hot_rows;
    if (H < 1) return fail(ACB_ECUDA, "rows too wide for the staged kernel");
    const uint32_t hot_bytes = (((H + 1) * row_bytes) + 127u) & ~127u;
    const uint32_t smem = hot_bytes + kStageOffset + stage_bytes;
    static thread_local int configured_for = -1;
    (void)configured_for;
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    CUDA_OK(cudaMemsetAsync(task_counter, 0, sizeof(unsigned int), st));
    kern<<<ctas, warps * 32, smem, st>>>(im, hot, U, out, H, hot_bytes, task_counter, out.total + 2);
    g_launches++;
    return ACB_OK;
}

template <int MODE, bool CP>
int launch_scan(const ImageHeader &h, const DevImage &im, const DevHot *hot, const Units &U, const Sink &out,
                const DeviceInfo &d, unsigned int *task_counter, cudaStream_t st) {
    int kernel = g_tuning.kernel;
    if (kernel == 0) kernel = 2;
    if (!hot) kernel = 1;  // no hot image: the plain kernel (table in global memory / L2)
    if (kernel == 1) return launch_plain<MODE, CP>(im, U, out, d, st);
    if (h.col_mode == kColRange) return launch_staged<MODE, CP, kColRange>(im, *hot, U, out, d, task_counter, st);
    return launch_staged<MODE, CP, kColClass>(im, *hot, U, out, d, task_counter, st);
}

// dev_hot points at a device copy of a hot image; hot_rows is its row count (the
// host knows it: acb_hot_rows on the host copy), because the header lives on the device
int make_hot_view(const acb_automaton *a, const void *dev_hot, uint32_t hot_rows, DevHot &v) {
    const ImageHeader &ih = a->impl->hdr;
    if (hot_rows < 1 || hot_rows > 65534) return fail(ACB_EINVAL, "bad hot image row count");
    auto align16 = [](uint64_t x) { return (x + 15) & ~uint64_t(15); };
    const uint8_t *b = static_cast<const uint8_t *>(dev_hot);
    uint64_t off = align16(sizeof(HotHeader));
    v.table = reinterpret_cast<const uint16_t *>(b + off);
    off = align16(off + uint64_t(hot_rows + 1) * ih.n_cols * 2);
    v.hot2full = reinterpret_cast<const uint32_t *>(b + off);
    off = align16(off + uint64_t(hot_rows + 1) * 4);
    v.full2hot = reinterpret_cast<const uint16_t *>(b + off);
    v.n_rows = hot_rows;
    return ACB_OK;
}

int run_scan(const acb_automaton *a, const void *dev_image, const DevHot *hot, const Units &U, int mode, int codepoints,
             const acb_workspace *ws, cudaStream_t st) {
    DeviceInfo d;
    int rc = device_info(d);
    if (rc) return rc;
    const ImageHeader &h = a->impl->hdr;
    const DevImage im = make_view(h, dev_image);
    Sink out;
    out.raw = ws->dev_raw;
    out.raw_seq = ws->dev_raw_seq;
    out.raw_unit = ws->dev_raw_unit;
    out.cap = ws->raw_capacity;
    out.unit_counts = ws->dev_unit_counts;
    out.total = reinterpret_cast<unsigned long long *>(ws->dev_total);
    unsigned int *task_counter = reinterpret_cast<unsigned int *>(ws->dev_scratch);
    unsigned long long *tile_sums = reinterpret_cast<unsigned long long *>(ws->dev_scratch) + 2;

    CUDA_OK(cudaMemsetAsync(ws->dev_total, 0, 4 * sizeof(uint64_t), st));
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (g_timing && U.n_units > 0) {
        CUDA_OK(cudaEventCreate(&ev0));
        CUDA_OK(cudaEventCreate(&ev1));
    }
    if (U.n_units > 0) {
        const bool cp = codepoints != 0;
        if (ev0) CUDA_OK(cudaEventRecord(ev0, st));
        if (mode == kModeStandard)
            rc = cp ? launch_scan<kModeStandard, true>(h, im, hot, U, out, d, task_counter, st)
                    : launch_scan<kModeStandard, false>(h, im, hot, U, out, d, task_counter, st);
        else if (mode == kModeLeftmost)
            rc = cp ? launch_scan<kModeLeftmost, true>(h, im, hot, U, out, d, task_counter, st)
                    : launch_scan<kModeLeftmost, false>(h, im, hot, U, out, d, task_counter, st);
        else
            rc = cp ? launch_scan<kModeOverlap, true>(h, im, hot, U, out, d, task_counter, st)
                    : launch_scan<kModeOverlap, false>(h, im, hot, U, out, d, task_counter, st);
        if (rc) return rc;
        CUDA_OK(cudaGetLastError());
        if (ev1) {
            CUDA_OK(cudaEventRecord(ev1, st));
            g_timing_events.emplace_back(ev0, ev1);
        }
    }
    // counts -> offsets -> ordered output
    const uint64_t n = (uint64_t)U.n_units;
    const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
    if (n == 0) {
        CUDA_OK(cudaMemsetAsync(ws->dev_unit_offsets, 0, sizeof(uint64_t), st));
    } else {
        scan_tile_sums<<<(unsigned)tiles, kScanThreads, 0, st>>>(ws->dev_unit_counts, n, tile_sums);
        scan_tile_offsets<<<1, kScanThreads, 0, st>>>(tile_sums, tiles);
        scan_apply<<<(unsigned)tiles, kScanThreads, 0, st>>>(ws->dev_unit_counts, n, tile_sums,
                                                             reinterpret_cast<unsigned long long *>(ws->dev_unit_offsets));
        order_matches_kernel<<<d.sms * 4, 256, 0, st>>>(ws->dev_raw, ws->dev_raw_seq, ws->dev_raw_unit, ws->raw_capacity,
                                                       out.total, reinterpret_cast<unsigned long long *>(ws->dev_unit_offsets),
                                                       ws->dev_out, ws->out_capacity);
        g_launches += 4;
    }
    finish_total_kernel<<<1, 1, 0, st>>>(out.total, ws->raw_capacity, ws->out_capacity);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return ACB_OK;
}

int check_ws(const acb_workspace *ws) {
    if (!ws || !ws->dev_raw || !ws->dev_raw_seq || !ws->dev_raw_unit || !ws->dev_unit_counts || !ws->dev_unit_offsets ||
        !ws->dev_scratch || !ws->dev_total || !ws->dev_out)
        return fail(ACB_EINVAL, "workspace has a null buffer");
    return ACB_OK;
}

}  // namespace

extern "C" {

int acb_profile(const acb_automaton *a, const void *dev_image, const uint8_t *dev_bytes, const int64_t *dev_offsets,
                int64_t n_haystacks, uint64_t len, int overlapping, uint32_t *dev_visits, void *stream) {
    if (!a || !dev_image || !dev_visits) return fail(ACB_EINVAL, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const ImageHeader &h = a->impl->hdr;
    CUDA_OK(cudaMemsetAsync(dev_visits, 0, uint64_t(h.n_states) * 4, st));
    Units U{};
    U.bytes = dev_bytes;
    U.offsets = dev_offsets;
    U.n_units = n_haystacks;
    U.len = len;
    U.chunk = dev_offsets ? 0 : 1;
    int64_t n_samples = 256;
    if (dev_offsets) {
        if (n_haystacks < 1) return ACB_OK;
        if (n_samples > n_haystacks) n_samples = n_haystacks;
    } else {
        if (len == 0) return ACB_OK;
        if ((uint64_t)n_samples > len / 1024 + 1) n_samples = (int64_t)(len / 1024 + 1);
    }
    const DevImage im = make_view(h, dev_image);
    const int restart = (!overlapping && h.match_kind == ACB_STANDARD) ? 1 : 0;
    profile_kernel<<<(unsigned)((n_samples + 127) / 128), 128, 0, st>>>(im, U, dev_visits, n_samples, 1024, restart);
    g_launches++;
    CUDA_OK(cudaGetLastError());
    return ACB_OK;
}

int acb_scan_batch(const acb_automaton *a, const void *dev_image, const void *dev_hot, uint32_t hot_rows,
                   const uint8_t *dev_bytes, const int64_t *dev_offsets, int64_t n_haystacks, int overlapping,
                   int codepoints, const acb_workspace *ws, void *stream) {
    if (!a || !dev_image || !dev_offsets || n_haystacks < 0) return fail(ACB_EINVAL, "bad argument");
    if (n_haystacks > 0xffffffffll) return fail(ACB_EINVAL, "too many haystacks in one batch");
    const int kind = (int)a->impl->hdr.match_kind;
    if (overlapping && kind != ACB_STANDARD)
        return fail(ACB_EUNSUPPORTED, std::string("match kind ") + (kind == ACB_LEFTMOST_FIRST ? "LeftmostFirst" : "LeftmostLongest") +
                                          " does not support overlapping searches");
    int rc = check_ws(ws);
    if (rc) return rc;
    Units U{};
    U.bytes = dev_bytes;
    U.offsets = dev_offsets;
    U.n_units = n_haystacks;
    U.chunk = 0;
    const int mode = overlapping ? kModeOverlap : (kind == ACB_STANDARD ? kModeStandard : kModeLeftmost);
    DevHot hot;
    if (dev_hot && (rc = make_hot_view(a, dev_hot, hot_rows, hot))) return rc;
    return run_scan(a, dev_image, dev_hot ? &hot : nullptr, U, mode, codepoints, ws, static_cast<cudaStream_t>(stream));
}

int acb_scan_chunked(const acb_automaton *a, const void *dev_image, const void *dev_hot, uint32_t hot_rows,
                     const uint8_t *dev_bytes, uint64_t len, uint32_t chunk_bytes, int codepoints,
                     const acb_workspace *ws, void *stream) {
    if (!a || !dev_image || (!dev_bytes && len)) return fail(ACB_EINVAL, "bad argument");
    if (chunk_bytes < 64) return fail(ACB_EINVAL, "chunk_bytes must be at least 64");
    if (len >= 0xffffffffull) return fail(ACB_EINVAL, "haystacks of 4 GiB and more are not supported yet");
    const int kind = (int)a->impl->hdr.match_kind;
    if (kind != ACB_STANDARD)
        return fail(ACB_EUNSUPPORTED, std::string("match kind ") + (kind == ACB_LEFTMOST_FIRST ? "LeftmostFirst" : "LeftmostLongest") +
                                          " does not support overlapping searches");
    int rc = check_ws(ws);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Units U{};
    U.bytes = dev_bytes;
    U.offsets = nullptr;
    U.n_units = (int64_t)acb_chunk_count(len, chunk_bytes);
    U.len = len;
    U.chunk = chunk_bytes;
    const uint32_t L = a->impl->hdr.max_pat_len;
    U.halo = L ? L - 1 : 0;
    U.chunk_cp = nullptr;
    if (codepoints && U.n_units) {
        // code points before each chunk: count per chunk, then prefix-sum
        const uint64_t n = (uint64_t)U.n_units;
        const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
        unsigned long long *tile_sums = reinterpret_cast<unsigned long long *>(ws->dev_scratch) + 2;
        uint32_t *cnt = reinterpret_cast<uint32_t *>(tile_sums + tiles + 1);
        unsigned long long *offs = tile_sums + tiles + 1 + (n + 1) / 2 + 1;
        const unsigned blocks = (unsigned)((n * 32 + 255) / 256);
        chunk_cp_count_kernel<<<blocks, 256, 0, st>>>(dev_bytes, len, chunk_bytes, cnt, n);
        scan_tile_sums<<<(unsigned)tiles, kScanThreads, 0, st>>>(cnt, n, tile_sums);
        scan_tile_offsets<<<1, kScanThreads, 0, st>>>(tile_sums, tiles);
        scan_apply<<<(unsigned)tiles, kScanThreads, 0, st>>>(cnt, n, tile_sums, offs);
        g_launches += 4;
        CUDA_OK(cudaGetLastError());
        U.chunk_cp = reinterpret_cast<const uint64_t *>(offs);
    }
    DevHot hot;
    if (dev_hot && (rc = make_hot_view(a, dev_hot, hot_rows, hot))) return rc;
    return run_scan(a, dev_image, dev_hot ? &hot : nullptr, U, kModeOverlap, codepoints, ws, st);
}

}  // extern "C"
