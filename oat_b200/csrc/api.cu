// api.cu -- the C ABI of liboatgpu.so (include/oatgpu.h).  Host-side orchestration only: pointer
// classification + staging, per-frame constants, kernel launches on the context's stream.
// There is NO CPU fallback anywhere in this file: without a CUDA device every compute entry
// point fails with OAT_ERR_CUDA.
#include <float.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <atomic>
#include <new>
#include <map>
#include <mutex>
#include <string>
#include <deque>
#include <vector>

#include "common.cuh"
#include "mog_fused.cuh"
#include "mog_pipe.cuh"
#include "pixel_ops.cuh"
#include "tail.cuh"
#include "tail_fast.cuh"
#include "posfilt.cuh"

using namespace oat;

// ---- errors --------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess)                                                                     \
            return fail(_e == cudaErrorMemoryAllocation ? OAT_ERR_NOMEM : OAT_ERR_CUDA,            \
                        std::string(#call) + ": " + cudaGetErrorString(_e));                       \
    } while (0)
#define CKRET(expr)          \
    do {                     \
        int _r = (expr);     \
        if (_r != OAT_OK) return _r; \
    } while (0)
#define REQUIRE(cond, msg) \
    do {                   \
        if (!(cond)) return fail(OAT_ERR_INVALID, msg); \
    } while (0)

// ---- context -------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return OAT_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        CK(cudaMalloc(&p, bytes));
        cap = bytes;
        return OAT_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

namespace {
struct ProfRec {  // one bracketed launch of the resident fused kernel (oat_ctx_profile_resident)
    cudaEvent_t e0, e1;
    uint64_t frames;
};
}  // namespace

// staging of one chunk of the resident clip engine (see clip_run)
struct ClipHalf {
    FrameDesc *h_desc = nullptr, *d_desc = nullptr;
    TailFrame *h_tf = nullptr, *d_tf = nullptr;
    unsigned int *d_ctr = nullptr;  // band counters of the tail server, 2 per frame, + its exit ticket
    size_t cap = 0;
    cudaEvent_t done = nullptr;
    int ensure(size_t n)
    {
        if (!done) CK(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        if (n <= cap) return OAT_OK;
        release_bufs();
        CK(cudaHostAlloc(&h_desc, n * sizeof(FrameDesc), cudaHostAllocDefault));
        CK(cudaHostAlloc(&h_tf, n * sizeof(TailFrame), cudaHostAllocDefault));
        CK(cudaMalloc(&d_desc, n * sizeof(FrameDesc)));
        CK(cudaMalloc(&d_tf, n * sizeof(TailFrame)));
        // (zeroed once: the tail server's last CTA to leave re-arms the counters of its launch and the ticket behind them)
        CK(cudaMalloc(&d_ctr, (2 * n + 1) * sizeof(unsigned int)));
        CK(cudaMemset(d_ctr, 0, (2 * n + 1) * sizeof(unsigned int)));
        cap = n;
        return OAT_OK;
    }
    void release_bufs()
    {
        if (h_desc) cudaFreeHost(h_desc);
        if (h_tf) cudaFreeHost(h_tf);
        if (d_desc) cudaFree(d_desc);
        if (d_tf) cudaFree(d_tf);
        if (d_ctr) cudaFree(d_ctr);
        h_desc = nullptr;
        h_tf = nullptr;
        d_desc = nullptr;
        d_tf = nullptr;
        d_ctr = nullptr;
        cap = 0;
    }
    void release()
    {
        release_bufs();
        if (done) cudaEventDestroy(done);
        done = nullptr;
    }
};

struct oat_ctx {
    int refs = 1;  // the caller's handle + one per live object created from it: the context is released by whoever leaves last
    int device = 0;
    cudaStream_t stream = nullptr;   // compute
    cudaStream_t h2d = nullptr;      // ingest copies (overlap with compute of the previous frame)
    static const int NTAIL = 8;
    unsigned tail_rr = 0;
    cudaStream_t tail[NTAIL] = {};  // detect tails of in-flight frames (higher priority than compute)
    int *hsv_lut = nullptr;          // sdiv[256] | hdiv[256]
    uint64_t launches = 0;
    int num_sms = 148;
    bool pipe_attr_set = false;
    // development switches, read once at creation: OAT_B200_NO_PIPE / _NO_FAST_TAIL / _NO_OVERLAP force the generic
    // fused kernel / the multi-launch tail / single-stream execution (A-B measurements; results are identical)
    bool no_pipe = false, no_fast_tail = false, no_overlap = false, pdl = true;
    bool no_prelabel = false;  // OAT_B200_NO_PRELABEL (A/B switch): the labelling CTA extracts every run table itself
    bool force_prelabel = false;  // OAT_B200_FORCE_PRELABEL (tests): the bands' tables are used whatever the size of the mask
    double heavy_tail_at = 10000.0;  // run-table entries per frame from which the tail server gets twice its share (OAT_B200_HEAVY_TAIL_AT)
    cudaStream_t post = nullptr;  // position epilogues (Kalman/mean), strictly in frame order
    // Work scheduler of the resident fused kernel (mog_pipe.cuh): every launch draws its (frame, tile) items from
    // one counter of this ring (slot = launch number % NSLOTS; counter and exit ticket are re-armed by the
    // launch's last CTA, which then writes the launch number into done_host[slot], pinned host memory).  A slot
    // is handed to a new launch only once done_host shows that its previous user has left -- so two live
    // launches never share a slot, whatever the number of launches in flight.
    static const unsigned NSLOTS = 64;
    unsigned int *work_counter = nullptr;   // [NSLOTS] counters, then [NSLOTS] exit tickets
    volatile unsigned int *done_host = nullptr;  // [NSLOTS] pinned + mapped
    unsigned int *done_dev = nullptr;            // device view of done_host
    unsigned int slot_user[NSLOTS] = {};         // launch number (+1) of the slot's last user, 0 = never used
    // model whose full-grid resident fused kernel was the LAST kernel enqueued on `stream` (0 = none): the
    // next launch may order itself behind it tile by tile instead of waiting for the whole grid
    unsigned long long chain_uid = 0;
    bool no_chain = false, no_mirror = false, no_track = false, no_clip = false, relaxed_publish = false, fence_always = false;
    uint64_t pipe_launches = 0;
    cudaStream_t aux = nullptr;  // small host-synchronous uploads (frame descriptors of a clip)
    cudaStream_t lane[2] = {nullptr, nullptr};      // oat_memcpy_async: ingest / egress copies of a host component
    cudaEvent_t lane_done[2] = {nullptr, nullptr};  // ... the lane's latest copy
    cudaEvent_t lane_after = nullptr;               // ... what the compute stream had been given when a copy was enqueued
    ClipHalf clip[2];            // two chunks of the resident clip engine in flight
    const void *clip_owner = nullptr;  // the tracker whose stream (oat_tracker_stream_*) holds chunks in flight, if any
    DevBuf tail_scratch[2];      // per chunk in flight: one global-memory labelling area per CTA of the tail server
    // device timing of the resident fused kernel: a CUDA-event pair on the compute stream around every launch
    // (consecutive launches overlap tile by tile, so the brackets partition the timeline: their sum is the time
    // from the first launch's start to the last one's end)
    int prof_resident = 0;
    std::vector<ProfRec> prof_recs;
    // host side of the resident clip engine: time inside clip_run() spent working (descriptors, launches, reading
    // results) and spent waiting for a chunk's completion event, and the frames served
    double clip_busy_ns = 0.0, clip_wait_ns = 0.0;
    uint64_t clip_frames_total = 0;
    unsigned int *slow_count = nullptr;    // census: 4-pixel groups that left the fused kernel's fast path (cumulative)
    DevBuf flush;
    DevBuf scratch_in, scratch_out, scratch_roi;  // staging for the stateless entry points
};

extern "C" int oat_abi_version(void) { return OATGPU_ABI_VERSION; }
extern "C" const char *oat_last_error(void) { return g_err.c_str(); }
extern "C" int oat_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int bind(oat_ctx *c)
{
    if (!c) return fail(OAT_ERR_INVALID, "null context");
    CK(cudaSetDevice(c->device));
    return OAT_OK;
}

extern "C" int oat_ctx_create(int device_index, oat_ctx **out)
{
    REQUIRE(out, "oat_ctx_create: out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(OAT_ERR_CUDA, std::string("no CUDA device available: ") +
                                      (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    // reference: "Selected GPU index is invalid." (src/framefilter/BackgroundSubtractorMOG.cpp:94-100)
    if (device_index < 0 || device_index >= n) return fail(OAT_ERR_INVALID, "Selected GPU index is invalid.");
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device_index));
    if (prop.major != 10)
        return fail(OAT_ERR_UNSUPPORTED, std::string("liboatgpu is built for sm_100a only; device is ") + prop.name);
    oat_ctx *c = new (std::nothrow) oat_ctx();
    if (!c) return fail(OAT_ERR_NOMEM, "out of host memory");
    c->device = device_index;
    c->num_sms = prop.multiProcessorCount;
    c->no_pipe = getenv("OAT_B200_NO_PIPE") != nullptr;
    c->no_fast_tail = getenv("OAT_B200_NO_FAST_TAIL") != nullptr;
    c->no_prelabel = getenv("OAT_B200_NO_PRELABEL") != nullptr;
    c->force_prelabel = getenv("OAT_B200_FORCE_PRELABEL") != nullptr;
    if (const char *e = getenv("OAT_B200_HEAVY_TAIL_AT")) c->heavy_tail_at = atof(e);
    c->no_overlap = getenv("OAT_B200_NO_OVERLAP") != nullptr;
    c->pdl = getenv("OAT_B200_NO_PDL") == nullptr;
    c->no_chain = getenv("OAT_B200_NO_CHAIN") != nullptr;
    c->no_mirror = getenv("OAT_B200_NO_MIRROR") != nullptr;
    c->no_track = getenv("OAT_B200_NO_TRACK") != nullptr;
    c->no_clip = getenv("OAT_B200_NO_CLIP") != nullptr;
    c->relaxed_publish = getenv("OAT_B200_RELAXED_PUBLISH") != nullptr;  // measurement only: prices the release fences
    c->fence_always = getenv("OAT_B200_FENCE_ALWAYS") != nullptr;        // measurement only: a fence for every tile
    CK(cudaSetDevice(device_index));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->h2d, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->post, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->aux, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;  // numerically lower = higher priority
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (int i = 0; i < oat_ctx::NTAIL; ++i) CK(cudaStreamCreateWithPriority(&c->tail[i], cudaStreamNonBlocking, hi));
    }
    int lut[512];
    lut[0] = lut[256] = 0;
    for (int i = 1; i < 256; ++i) {
        lut[i] = (int)nearbyint((255 << 12) / (1.0 * i));
        lut[256 + i] = (int)nearbyint((180 << 12) / (6.0 * i));
    }
    CK(cudaMalloc(&c->slow_count, sizeof(unsigned int)));
    CK(cudaMemset(c->slow_count, 0, sizeof(unsigned int)));
    CK(cudaMalloc(&c->work_counter, 2 * oat_ctx::NSLOTS * sizeof(unsigned int)));
    CK(cudaMemset(c->work_counter, 0, 2 * oat_ctx::NSLOTS * sizeof(unsigned int)));
    {
        void *hp = nullptr, *dp = nullptr;
        CK(cudaHostAlloc(&hp, oat_ctx::NSLOTS * sizeof(unsigned int), cudaHostAllocMapped));
        memset(hp, 0, oat_ctx::NSLOTS * sizeof(unsigned int));
        CK(cudaHostGetDevicePointer(&dp, hp, 0));
        c->done_host = (volatile unsigned int *)hp;
        c->done_dev = (unsigned int *)dp;
    }
    CK(cudaMalloc(&c->hsv_lut, sizeof(lut)));
    CK(cudaMemcpy(c->hsv_lut, lut, sizeof(lut), cudaMemcpyHostToDevice));
    *out = c;
    return OAT_OK;
}

static void ctx_ref(oat_ctx *c) { ++c->refs; }
static void ctx_free(oat_ctx *c);
static void ctx_unref(oat_ctx *c)
{
    if (--c->refs == 0) ctx_free(c);
}
// Objects created from a context keep it alive: destroying the context first (a garbage collector's order, an
// exception path) defers its release to the last object's destroy instead of leaving them with a dangling pointer.
extern "C" int oat_ctx_destroy(oat_ctx *c)
{
    if (!c) return OAT_OK;
    ctx_unref(c);
    return OAT_OK;
}
static void ctx_free(oat_ctx *c)
{
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->h2d);
    if (c->post) {
        cudaStreamSynchronize(c->post);
        cudaStreamDestroy(c->post);
    }
    for (int i = 0; i < oat_ctx::NTAIL; ++i)
        if (c->tail[i]) {
            cudaStreamSynchronize(c->tail[i]);
            cudaStreamDestroy(c->tail[i]);
        }
    c->flush.release();
    c->scratch_in.release();
    c->scratch_out.release();
    c->scratch_roi.release();
    if (c->hsv_lut) cudaFree(c->hsv_lut);
    c->clip[0].release();
    c->clip[1].release();
    c->tail_scratch[0].release();
    c->tail_scratch[1].release();
    if (c->work_counter) cudaFree(c->work_counter);
    if (c->done_host) cudaFreeHost((void *)c->done_host);
    if (c->aux) {
        cudaStreamSynchronize(c->aux);
        cudaStreamDestroy(c->aux);
        for (int l = 0; l < 2; ++l) {
            if (c->lane[l]) cudaStreamDestroy(c->lane[l]);
            if (c->lane_done[l]) cudaEventDestroy(c->lane_done[l]);
        }
        if (c->lane_after) cudaEventDestroy(c->lane_after);
    }
    if (c->slow_count) cudaFree(c->slow_count);
    cudaStreamDestroy(c->stream);
    cudaStreamDestroy(c->h2d);
    delete c;
}

extern "C" int oat_ctx_idle(oat_ctx *c, int *idle)
{
    CKRET(bind(c));
    REQUIRE(idle, "oat_ctx_idle: null argument");
    const cudaError_t e = cudaStreamQuery(c->stream);
    if (e != cudaSuccess && e != cudaErrorNotReady) CK(e);
    *idle = (e == cudaSuccess) ? 1 : 0;
    return OAT_OK;
}
extern "C" int oat_ctx_sync(oat_ctx *c)
{
    CKRET(bind(c));
    CK(cudaStreamSynchronize(c->h2d));
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < oat_ctx::NTAIL; ++i) CK(cudaStreamSynchronize(c->tail[i]));
    return OAT_OK;
}
extern "C" void *oat_ctx_stream(oat_ctx *c) { return c ? (void *)c->stream : nullptr; }
extern "C" uint64_t oat_ctx_kernel_launches(const oat_ctx *c) { return c ? c->launches : 0; }
extern "C" int oat_ctx_clip_host_stats(oat_ctx *c, double *busy_us, double *wait_us, uint64_t *frames)
{
    REQUIRE(c, "null context");
    if (busy_us) *busy_us = c->clip_busy_ns * 1e-3;
    if (wait_us) *wait_us = c->clip_wait_ns * 1e-3;
    if (frames) *frames = c->clip_frames_total;
    c->clip_busy_ns = c->clip_wait_ns = 0.0;
    c->clip_frames_total = 0;
    return OAT_OK;
}
extern "C" int oat_ctx_profile_resident(oat_ctx *c, int enable)
{
    CKRET(bind(c));
    c->prof_resident = enable ? 1 : 0;
    return OAT_OK;
}
extern "C" int oat_ctx_profile_resident_read(oat_ctx *c, double *total_ms, uint64_t *launches, uint64_t *frames)
{
    CKRET(bind(c));
    CK(cudaStreamSynchronize(c->stream));
    double ms = 0.0;
    uint64_t nf = 0;
    for (auto &r : c->prof_recs) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms += t;
        nf += r.frames;
        cudaEventDestroy(r.e0);
        cudaEventDestroy(r.e1);
    }
    if (total_ms) *total_ms = ms;
    if (launches) *launches = c->prof_recs.size();
    if (frames) *frames = nf;
    c->prof_recs.clear();
    return OAT_OK;
}

// ---- pointer classification + staging ------------------------------------------------------
enum MemKind { MEM_PAGEABLE, MEM_PINNED, MEM_DEVICE };
// Device allocations handed out by oat_alloc_device (frames a host component or a benchmark keeps in HBM): a pointer
// into one of them is device memory without asking the driver -- cudaPointerGetAttributes is a driver call per frame
// on the enqueue path, and driver calls are what gets slow when several processes drive their GPUs at once.
static std::mutex g_dev_mu;
static std::map<uintptr_t, size_t> g_dev_ranges;
static bool known_device_ptr(const void *p)
{
    const uintptr_t a = (uintptr_t)p;
    std::lock_guard<std::mutex> lk(g_dev_mu);
    auto it = g_dev_ranges.upper_bound(a);
    if (it == g_dev_ranges.begin()) return false;
    --it;
    return a - it->first < it->second;
}
static MemKind mem_kind(const void *p)
{
    if (known_device_ptr(p)) return MEM_DEVICE;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return MEM_PAGEABLE;
    }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return MEM_DEVICE;
    if (at.type == cudaMemoryTypeHost) return MEM_PINNED;
    return MEM_PAGEABLE;
}

// A frame into `stage` at a 16-byte aligned pitch (`tight`).  A contiguous frame whose rows are not a multiple of 16
// bytes (a 1000-column BGR frame: 3000 B) would need a strided 2-D DMA -- measured at about half the PCIe rate of a
// linear one -- so, given a second buffer, a HOST frame goes up linearly into `raw` and is re-pitched on the device.
static int repitch(cudaStream_t s, void *dst, size_t dpitch, const void *src, size_t spitch, size_t rowbytes, int rows)
{
    if ((((uintptr_t)dst | (uintptr_t)src | dpitch | spitch | rowbytes) & 3u) != 0) {
        CK(cudaMemcpy2DAsync(dst, dpitch, src, spitch, rowbytes, rows, cudaMemcpyDeviceToDevice, s));
        return OAT_OK;
    }
    const bool v16 = (((uintptr_t)dst | (uintptr_t)src | dpitch | spitch | rowbytes) & 15u) == 0;
    const unsigned units = (unsigned)(rowbytes / (v16 ? 16 : 4));
    const unsigned long long total = (unsigned long long)units * rows;
    const unsigned grid = (unsigned)std::min<unsigned long long>((total + 255) / 256, 148ull * 4);
    if (v16)
        repitch_kernel<uint4><<<grid, 256, 0, s>>>((uint8_t *)dst, dpitch, (const uint8_t *)src, spitch, units, (unsigned)rows);
    else
        repitch_kernel<uint32_t><<<grid, 256, 0, s>>>((uint8_t *)dst, dpitch, (const uint8_t *)src, spitch, units, (unsigned)rows);
    CK(cudaGetLastError());
    return OAT_OK;
}
static int copy_frame_in(cudaStream_t s, DevBuf &stage, DevBuf *raw, const void *src, size_t pitch, int rows, size_t rowbytes,
                         size_t tight, bool src_is_device)
{
    CKRET(stage.ensure(tight * rows));
    if (src_is_device) {
        if (pitch == rowbytes && tight == rowbytes)
            CKRET(repitch(s, stage.p, rowbytes * rows, src, rowbytes * rows, rowbytes * rows, 1));  // one long row
        else
            CKRET(repitch(s, stage.p, tight, src, pitch, rowbytes, rows));
    } else if (pitch == tight) {  // same pitch on both sides (tight frames of 16-byte-multiple rows included): one linear DMA
        CK(cudaMemcpyAsync(stage.p, src, pitch * (size_t)(rows - 1) + rowbytes, cudaMemcpyHostToDevice, s));
    } else if (raw && pitch == rowbytes) {
        CKRET(raw->ensure(rowbytes * rows));
        CK(cudaMemcpyAsync(raw->p, src, rowbytes * rows, cudaMemcpyHostToDevice, s));
        CKRET(repitch(s, stage.p, tight, raw->p, rowbytes, rowbytes, rows));
    } else {
        CK(cudaMemcpy2DAsync(stage.p, tight, src, pitch, rowbytes, rows, cudaMemcpyHostToDevice, s));
    }
    return OAT_OK;
}
// Input image: returns a device view (ptr,pitch); host images are copied into `stage`.
static int stage_in(oat_ctx *c, cudaStream_t s, DevBuf &stage, const void *src, size_t pitch, int rows,
                    size_t rowbytes, const uint8_t **dptr, size_t *dpitch, DevBuf *raw = nullptr)
{
    if (mem_kind(src) == MEM_DEVICE) {
        *dptr = (const uint8_t *)src;
        *dpitch = pitch;
        return OAT_OK;
    }
    const size_t tight = (rowbytes + 15) & ~(size_t)15;
    CKRET(copy_frame_in(s, stage, raw, src, pitch, rows, rowbytes, tight, false));
    *dptr = (const uint8_t *)stage.p;
    *dpitch = tight;
    return OAT_OK;
}
// Output image: device view to write into; if the user pointer is host memory the view is
// `stage` and finish_out() copies it back.
struct OutView {
    uint8_t *d = nullptr;
    size_t dpitch = 0;
    void *host = nullptr;
    size_t hpitch = 0;
    size_t rowbytes = 0;
    int rows = 0;
};
static int stage_out(DevBuf &stage, void *dst, size_t pitch, int rows, size_t rowbytes, OutView *v)
{
    *v = OutView();
    if (!dst) return OAT_OK;
    v->rows = rows;
    v->rowbytes = rowbytes;
    if (mem_kind(dst) == MEM_DEVICE) {
        v->d = (uint8_t *)dst;
        v->dpitch = pitch;
        return OAT_OK;
    }
    const size_t tight = (rowbytes + 15) & ~(size_t)15;
    CKRET(stage.ensure(tight * rows));
    v->d = (uint8_t *)stage.p;
    v->dpitch = tight;
    v->host = dst;
    v->hpitch = pitch;
    return OAT_OK;
}
static int finish_out(cudaStream_t s, const OutView &v)
{
    if (v.host)
    {
        if (v.hpitch == v.rowbytes && v.dpitch == v.rowbytes)
            CK(cudaMemcpyAsync(v.host, v.d, v.rowbytes * v.rows, cudaMemcpyDeviceToHost, s));
        else
            CK(cudaMemcpy2DAsync(v.host, v.hpitch, v.d, v.dpitch, v.rowbytes, v.rows, cudaMemcpyDeviceToHost, s));
    }
    return OAT_OK;
}

static inline unsigned nblocks(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }
#define LAUNCH_CHECK(c)                 \
    do {                                \
        ++(c)->launches;                \
        (c)->chain_uid = 0;             \
        CK(cudaGetLastError());         \
    } while (0)

// ---- GMM model -----------------------------------------------------------------------------
extern "C" void oat_mog_default_params(oat_mog_params *p)
{
    if (!p) return;
    p->history = 500;
    p->nmixtures = 5;
    p->var_threshold = 16.0f;
    p->var_threshold_gen = 9.0f;
    p->background_ratio = 0.9f;
    p->var_init = 15.0f;
    p->var_min = 4.0f;
    p->var_max = 75.0f;
    p->complexity_reduction_threshold = 0.05f;
    p->detect_shadows = 1;
    p->shadow_value = 127;
    p->shadow_threshold = 0.5f;
}

struct MogModel {
    BitGeom g{};
    int K = 5;
    oat_mog_params p{};
    float *state = nullptr;
    size_t plane = 0;
    uint8_t *nmodes = nullptr;
    int nframes = 0;
    unsigned long long *d_sum = nullptr;
    unsigned int *tile_seq = nullptr;  // per tile of the pipelined kernel: sequence number of the last launch that finished it
    unsigned int seq = 0;              // sequence number the next pipelined launch expects (and publishes + 1)
    bool flags_current = false;        // the last kernel that touched the state was the pipelined one (tile_seq is valid)
    unsigned long long uid = 0;

    int create(int rows, int cols, const oat_mog_params *params)
    {
        if (params)
            p = *params;
        else
            oat_mog_default_params(&p);
        REQUIRE(rows > 0 && cols > 0 && rows <= 32768 && cols <= 32768, "MOG: bad frame geometry");
        REQUIRE(p.nmixtures >= 1 && p.nmixtures <= 5, "MOG: nmixtures must be in 1..5");
        REQUIRE(p.history >= 1, "MOG: history must be >= 1");
        REQUIRE(p.shadow_value >= 0 && p.shadow_value <= 255, "MOG: shadow_value must be in 0..255");
        g.rows = rows;
        g.cols = cols;
        g.wpr = div_up(cols, 32);
        K = p.nmixtures;
        plane = (size_t)rows * g.pitch_px();
        CK(cudaMalloc(&state, plane * 5 * K * sizeof(float)));
        CK(cudaMalloc(&nmodes, plane));
        CK(cudaMalloc(&d_sum, sizeof(unsigned long long)));
        CK(cudaMemset(nmodes, 0, plane));
        const size_t ntiles = (plane + PIPE_TILE - 1) / PIPE_TILE;
        CK(cudaMalloc(&tile_seq, ntiles * sizeof(unsigned int)));
        CK(cudaMemset(tile_seq, 0, ntiles * sizeof(unsigned int)));
        CK(cudaStreamSynchronize(cudaStreamLegacy));  // the memsets ran on the legacy stream; the context's streams do not wait for it
        static std::atomic<unsigned long long> next_uid{1};
        uid = next_uid++;
        seq = 0;
        nframes = 0;
        return OAT_OK;
    }
    void destroy()
    {
        if (state) cudaFree(state);
        if (nmodes) cudaFree(nmodes);
        if (d_sum) cudaFree(d_sum);
        if (tile_seq) cudaFree(tile_seq);
        tile_seq = nullptr;
        state = nullptr;
        nmodes = nullptr;
        d_sum = nullptr;
    }
    // cv::BackgroundSubtractorMOG2Impl::apply: (re)initialise, count the frame, pick the rate.
    void frame_consts(double learning_rate, MogConsts *c, int *reset)
    {
        const bool init = (nframes == 0) || (learning_rate >= 1);
        if (init) nframes = 0;
        ++nframes;
        int d = 2 * nframes;
        if (d > p.history) d = p.history;
        const double lr = (learning_rate >= 0 && nframes > 1) ? learning_rate : 1.0 / d;
        c->aT = (float)lr;
        c->a1 = 1.0f - c->aT;
        c->prune = (float)(-lr * (double)p.complexity_reduction_threshold);
        c->Tb = p.var_threshold;
        c->TB = p.background_ratio;
        c->Tg = p.var_threshold_gen;
        c->varInit = p.var_init;
        c->varMin = p.var_min;
        c->varMax = p.var_max;
        c->tau = p.shadow_threshold;
        c->detect_shadows = p.detect_shadows;
        c->shadow_value = p.shadow_value;
        *reset = init ? 1 : 0;
    }
};

template <int PX, bool TRACK>
static void launch_fused_px(cudaStream_t s, int K, const FusedArgs &a)
{
    const long long threads = (long long)a.rows * a.wpr * (32 / PX);
    const unsigned grid = nblocks(threads, 128);
    switch (K) {
    case 1: mog_fused_kernel<1, PX, TRACK><<<grid, 128, 0, s>>>(a); break;
    case 2: mog_fused_kernel<2, PX, TRACK><<<grid, 128, 0, s>>>(a); break;
    case 3: mog_fused_kernel<3, PX, TRACK><<<grid, 128, 0, s>>>(a); break;
    case 4: mog_fused_kernel<4, PX, TRACK><<<grid, 128, 0, s>>>(a); break;
    default: mog_fused_kernel<5, PX, TRACK><<<grid, 128, 0, s>>>(a); break;
    }
}

static bool aligned4(const void *p, size_t pitch) { return (((uintptr_t)p | pitch) & 3u) == 0; }

// Scheduler slot for the next launch of the resident fused kernel: the slot's previous user (NSLOTS launches
// ago) must have left -- its last CTA says so in pinned host memory.  In practice that happened long ago; the
// wait is bounded so that a dead context surfaces as an error.
static int acquire_slot(oat_ctx *c, unsigned *slot_out)
{
    const unsigned slot = (unsigned)(c->pipe_launches % oat_ctx::NSLOTS);
    const unsigned want = c->slot_user[slot];
    if (want != 0u) {
        unsigned long long spins = 0;
        while (c->done_host[slot] != want) {
            if ((++spins & 0xfffull) == 0) {
                cudaError_t e = cudaStreamQuery(c->stream);
                if (e != cudaSuccess && e != cudaErrorNotReady)
                    return fail(OAT_ERR_CUDA, std::string("resident fused kernel: ") + cudaGetErrorString(e));
                if (spins > (1ull << 34)) return fail(OAT_ERR_CUDA, "resident fused kernel: scheduler slot never released");
            }
        }
    }
    *slot_out = slot;
    return OAT_OK;
}

// Geometry/aliasing class of a frame for the resident fused kernel: 2 = LINEAR (tight pitches, cols % 32 == 0:
// one bulk copy per tile), 1 = every row starts 16-byte aligned (one bulk copy per row segment), 0 = neither
// (generic per-thread kernel).  Egress images, when present, only need 4-byte alignment unless LINEAR.
static int pipe_class(const MogModel &m, const FusedArgs &a)
{
    const size_t tight3 = (size_t)3 * m.g.cols;
    const bool linear = (m.g.cols % 32 == 0) && a.in_pitch == tight3 && (((uintptr_t)a.bgr & 15u) == 0) &&
                        (!a.bgr_out || a.bgr_out_pitch == tight3) && (!a.hsv_out || a.hsv_pitch == tight3) &&
                        (!a.fg_out || a.fg_pitch == (size_t)m.g.cols);
    if (linear) return 2;
    if ((((uintptr_t)a.bgr | a.in_pitch) & 15u) == 0 && m.g.cols >= 64) return 1;
    return 0;
}

// One launch of the resident fused kernel over `pa.nframes` frames (descriptors in pa.descs, or pa.one).
// Fills in the scheduler slot; commits the host bookkeeping only once the launch call has succeeded.
// reserved: CTA slots of the persistent grid left to a detect tail running beside it (0 = none)
static int pipe_grid_full(const oat_ctx *c, int reserved)
{
    const int g = PIPE_CTAS_PER_SM * c->num_sms - reserved;
    return g > 0 ? g : 1;
}

static int launch_stream(oat_ctx *c, StreamArgs &pa, bool frozen, bool linear, int reserved)
{
    if (!c->pipe_attr_set) {
        CK(cudaFuncSetAttribute(mog_stream_kernel<5, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES));
        CK(cudaFuncSetAttribute(mog_stream_kernel<5, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES));
        CK(cudaFuncSetAttribute(mog_stream_kernel<5, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES));
        CK(cudaFuncSetAttribute(mog_stream_kernel<5, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BYTES));
        c->pipe_attr_set = true;
    }
    const long long total = (long long)pa.ntiles * pa.nframes;
    REQUIRE(total > 0 && total < (1ll << 30), "resident fused kernel: bad work size");
    const int full = pipe_grid_full(c, reserved);
    const int grid = total < full ? (int)total : full;
    unsigned slot = 0;
    CKRET(acquire_slot(c, &slot));
    pa.work_counter = c->work_counter + slot;
    pa.exit_ticket = c->work_counter + oat_ctx::NSLOTS + slot;
    pa.done_flag = c->done_dev + slot;
    pa.launch_id = (unsigned)(c->pipe_launches + 1);
    pa.relaxed_publish = c->relaxed_publish ? 1 : 0;
    pa.fence_always = c->fence_always ? 1 : 0;
    if (pa.launch_id == 0u) pa.launch_id = 1u;  // 0 means "never used"
    // programmatic dependent launch: the next launch's CTAs become resident (and run their prologue: mbarrier
    // + queue initialisation, first draw) while this launch's last CTAs drain
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3((unsigned)grid);
    lc.blockDim = dim3(PIPE_THREADS);
    lc.dynamicSmemBytes = PIPE_SMEM_BYTES;
    lc.stream = c->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = c->pdl ? 1u : 0u;
    if (!c->pdl) pa.wait_grid = 0;  // plain stream order: nothing to wait for inside the kernel
    ProfRec pr{nullptr, nullptr, (uint64_t)pa.nframes};
    if (c->prof_resident) {
        CK(cudaEventCreate(&pr.e0));
        CK(cudaEventCreate(&pr.e1));
        CK(cudaEventRecord(pr.e0, c->stream));
    }
    if (frozen && linear)
        CK(cudaLaunchKernelEx(&lc, mog_stream_kernel<5, true, true>, pa));
    else if (frozen)
        CK(cudaLaunchKernelEx(&lc, mog_stream_kernel<5, true, false>, pa));
    else if (linear)
        CK(cudaLaunchKernelEx(&lc, mog_stream_kernel<5, false, true>, pa));
    else
        CK(cudaLaunchKernelEx(&lc, mog_stream_kernel<5, false, false>, pa));
    c->slot_user[slot] = pa.launch_id;
    ++c->pipe_launches;
    ++c->launches;
    if (c->prof_resident) {
        CK(cudaEventRecord(pr.e1, c->stream));
        c->prof_recs.push_back(pr);
    }
    return OAT_OK;
}

static void stream_args_common(oat_ctx *c, const MogModel &m, const FusedArgs &a, StreamArgs &pa)
{
    pa.f = a;
    pa.ntiles = (int)((m.plane + PIPE_TILE - 1) / PIPE_TILE);
    pa.nframes = 1;
    pa.descs = nullptr;
    pa.div_magic = 0xffffffffffffffffull / (unsigned long long)m.g.pitch_px() + 1ull;
    pa.zero_in = a.do_hsv && a.lo[0] <= 0 && a.hi[0] >= 0 && a.lo[1] <= 0 && a.hi[1] >= 0 && a.lo[2] <= 0 && a.hi[2] >= 0;
    pa.wait_grid = 1;
}

static int launch_fused(oat_ctx *c, MogModel &m, FusedArgs &a, bool allow_pipe = true, bool allow_chain = false,
                        bool beside_tail = false)
{
    a.rows = m.g.rows;
    a.cols = m.g.cols;
    a.wpr = m.g.wpr;
    a.state = m.state;
    a.plane = m.plane;
    a.nmodes = m.nmodes;
    a.hsv_lut = c->hsv_lut;
    if (!a.slow_count) a.slow_count = c->slow_count;
    const bool vec = (m.g.cols % 4 == 0) && aligned4(a.bgr, a.in_pitch) &&
                     (!a.bgr_out || aligned4(a.bgr_out, a.bgr_out_pitch)) &&
                     (!a.fg_out || aligned4(a.fg_out, a.fg_pitch)) && (!a.hsv_out || aligned4(a.hsv_out, a.hsv_pitch));
    // a frozen model (learning rate 0) rewrites nothing: that variant tracks changes and skips
    // the state write-back; with a live rate every live mode changes every frame anyway.
    const bool frozen = (a.c.aT == 0.0f) && !a.reset && !c->no_track;
    const int cls = pipe_class(m, a);
    if (vec && cls > 0 && !a.reset && m.K == 5 && allow_pipe && !c->no_pipe) {
        // steady state: the resident bulk-async staged kernel (mog_pipe.cuh), here over a queue of one frame
        StreamArgs pa;
        stream_args_common(c, m, a, pa);
        const bool full = (long long)pa.ntiles >= (long long)pipe_grid_full(c, beside_tail ? PIPE_RESERVED_CTAS_CFG : 0);
        // order the frame behind the previous launch tile by tile (not grid by grid) when that launch was the
        // resident kernel too, it is the last kernel on the stream, both grids fill the machine, and this model's
        // tile flags are current (its own last launch was the resident kernel; the predecessor on the stream may
        // belong to another model -- independent streams share nothing)
        const bool chain = allow_chain && c->pdl && !c->no_chain && c->chain_uid != 0 && m.flags_current && full;
        FrameDesc &one = pa.inl[0];
        one.bgr = a.bgr;
        one.state = m.state;
        one.nmodes = m.nmodes;
        one.thr_bits = a.thr_bits;
        one.tile_seq = m.tile_seq;
        one.slow_count = a.slow_count;
        one.done_count = nullptr;
        one.seq_expect = m.seq;
        one.flags = chain ? FD_CHAIN : 0u;
        pa.wait_grid = chain ? 0 : 1;
        const int r = launch_stream(c, pa, frozen, cls == 2, beside_tail ? PIPE_RESERVED_CTAS_CFG : 0);
        if (r != OAT_OK) {  // nothing was enqueued: the model's sequence numbers and flags stand as they were
            c->chain_uid = 0;
            return r;
        }
        ++m.seq;
        c->chain_uid = full ? m.uid : 0;
        m.flags_current = true;
        return OAT_OK;
    }
    if (vec && frozen)
        launch_fused_px<4, true>(c->stream, m.K, a);
    else if (vec)
        launch_fused_px<4, false>(c->stream, m.K, a);
    else
        launch_fused_px<1, true>(c->stream, m.K, a);
    LAUNCH_CHECK(c);
    m.flags_current = false;  // the generic kernels do not publish tile flags
    return OAT_OK;
}

static int live_modes(oat_ctx *c, MogModel &m, uint64_t *sum)
{
    REQUIRE(sum, "live_modes: null output");
    CK(cudaMemsetAsync(m.d_sum, 0, sizeof(unsigned long long), c->stream));
    live_modes_kernel<<<nblocks((long long)m.plane, 256), 256, 0, c->stream>>>(m.nmodes, m.g, m.d_sum);
    LAUNCH_CHECK(c);
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, m.d_sum, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *sum = h;
    return OAT_OK;
}

// ---- framefilt mog -------------------------------------------------------------------------
struct oat_mog {
    oat_ctx *ctx;
    MogModel m;
    DevBuf in, out_bgr, out_mask;
};

extern "C" int oat_mog_create(oat_ctx *c, int rows, int cols, const oat_mog_params *params, oat_mog **out)
{
    REQUIRE(out, "oat_mog_create: out is null");
    *out = nullptr;
    CKRET(bind(c));
    oat_mog *h = new (std::nothrow) oat_mog();
    if (!h) return fail(OAT_ERR_NOMEM, "out of host memory");
    h->ctx = c;
    ctx_ref(c);
    int r = h->m.create(rows, cols, params);
    if (r != OAT_OK) {
        h->m.destroy();
        {
            oat_ctx *owner_ = h->ctx;
            delete h;
            ctx_unref(owner_);
        }
        return r;
    }
    *out = h;
    return OAT_OK;
}
extern "C" int oat_mog_destroy(oat_mog *h)
{
    if (!h) return OAT_OK;
    cudaSetDevice(h->ctx->device);
    cudaStreamSynchronize(h->ctx->stream);
    h->m.destroy();
    h->in.release();
    h->out_bgr.release();
    h->out_mask.release();
    {
        oat_ctx *owner_ = h->ctx;
        delete h;
        ctx_unref(owner_);
    }
    return OAT_OK;
}
extern "C" int oat_mog_reset(oat_mog *h)
{
    REQUIRE(h, "null handle");
    h->m.nframes = 0;
    return OAT_OK;
}

static int mog_apply_common(oat_mog *h, const uint8_t *bgr_in, size_t in_pitch, uint8_t *bgr_out, size_t out_pitch,
                            uint8_t *mask_out, size_t mask_pitch, double learning_rate, bool async)
{
    REQUIRE(h && bgr_in, "oat_mog_apply: null handle or input");
    oat_ctx *c = h->ctx;
    CKRET(bind(c));
    const int rows = h->m.g.rows, cols = h->m.g.cols;
    REQUIRE(in_pitch >= (size_t)3 * cols, "oat_mog_apply: input pitch too small");
    REQUIRE(!bgr_out || out_pitch >= (size_t)3 * cols, "oat_mog_apply: output pitch too small");
    REQUIRE(!mask_out || mask_pitch >= (size_t)cols, "oat_mog_apply: mask pitch too small");
    if (async)
        REQUIRE(mem_kind(bgr_in) == MEM_DEVICE && (!bgr_out || mem_kind(bgr_out) == MEM_DEVICE) &&
                    (!mask_out || mem_kind(mask_out) == MEM_DEVICE),
                "oat_mog_apply_async: device-resident images only");
    FusedArgs a{};
    CKRET(stage_in(c, c->stream, h->in, bgr_in, in_pitch, rows, (size_t)3 * cols, &a.bgr, &a.in_pitch));
    OutView ob, om;
    CKRET(stage_out(h->out_bgr, bgr_out, out_pitch, rows, (size_t)3 * cols, &ob));
    CKRET(stage_out(h->out_mask, mask_out, mask_pitch, rows, (size_t)cols, &om));
    h->m.frame_consts(learning_rate, &a.c, &a.reset);
    a.do_hsv = 0;
    a.thr_bits = nullptr;
    a.bgr_out = ob.d;
    a.bgr_out_pitch = ob.dpitch;
    a.fg_out = om.d;
    a.fg_pitch = om.dpitch;
    // device-resident frames in, device-resident frames out: nothing but this stream's own kernels orders the
    // frame, so consecutive launches may overlap tile by tile
    CKRET(launch_fused(c, h->m, a, true, async));
    if (async) return OAT_OK;
    CKRET(finish_out(c->stream, ob));
    CKRET(finish_out(c->stream, om));
    CK(cudaStreamSynchronize(c->stream));
    return OAT_OK;
}

extern "C" int oat_mog_apply(oat_mog *h, const uint8_t *bgr_in, size_t in_pitch, uint8_t *bgr_out, size_t out_pitch,
                             uint8_t *mask_out, size_t mask_pitch, double learning_rate)
{
    return mog_apply_common(h, bgr_in, in_pitch, bgr_out, out_pitch, mask_out, mask_pitch, learning_rate, false);
}

extern "C" int oat_mog_apply_async(oat_mog *h, const uint8_t *bgr_in, size_t in_pitch, uint8_t *bgr_out, size_t out_pitch,
                                   uint8_t *mask_out, size_t mask_pitch, double learning_rate)
{
    return mog_apply_common(h, bgr_in, in_pitch, bgr_out, out_pitch, mask_out, mask_pitch, learning_rate, true);
}

extern "C" int oat_mog_live_modes(oat_mog *h, uint64_t *sum)
{
    REQUIRE(h, "null handle");
    CKRET(bind(h->ctx));
    return live_modes(h->ctx, h->m, sum);
}

static int get_state(oat_ctx *c, MogModel &m, uint8_t *modes_used, float *weight, float *variance, float *mean)
{
    const size_t n = (size_t)m.g.rows * m.g.cols;
    const int K = m.K;
    uint8_t *dm = nullptr;
    float *dw = nullptr, *dv = nullptr, *dmean = nullptr;
    if (modes_used) CK(cudaMalloc(&dm, n));
    if (weight) CK(cudaMalloc(&dw, n * K * sizeof(float)));
    if (variance) CK(cudaMalloc(&dv, n * K * sizeof(float)));
    if (mean) CK(cudaMalloc(&dmean, n * K * 3 * sizeof(float)));
    state_export_kernel<<<nblocks((long long)n, 256), 256, 0, c->stream>>>(m.state, m.plane, m.nmodes, m.g, K, dm, dw,
                                                                           dv, dmean);
    LAUNCH_CHECK(c);
    CK(cudaStreamSynchronize(c->stream));
    if (dm) CK(cudaMemcpy(modes_used, dm, n, cudaMemcpyDeviceToHost));
    if (dw) CK(cudaMemcpy(weight, dw, n * K * sizeof(float), cudaMemcpyDeviceToHost));
    if (dv) CK(cudaMemcpy(variance, dv, n * K * sizeof(float), cudaMemcpyDeviceToHost));
    if (dmean) CK(cudaMemcpy(mean, dmean, n * K * 3 * sizeof(float), cudaMemcpyDeviceToHost));
    cudaFree(dm);
    cudaFree(dw);
    cudaFree(dv);
    cudaFree(dmean);
    return OAT_OK;
}

extern "C" int oat_mog_get_state(oat_mog *h, uint8_t *modes_used, float *weight, float *variance, float *mean)
{
    REQUIRE(h, "null handle");
    CKRET(bind(h->ctx));
    return get_state(h->ctx, h->m, modes_used, weight, variance, mean);
}

// ---- framefilt col -C HSV ------------------------------------------------------------------
extern "C" int oat_bgr2hsv(oat_ctx *c, const uint8_t *bgr, size_t in_pitch, uint8_t *hsv, size_t out_pitch, int rows,
                           int cols)
{
    CKRET(bind(c));
    REQUIRE(bgr && hsv && rows > 0 && cols > 0, "oat_bgr2hsv: bad arguments");
    REQUIRE(in_pitch >= (size_t)3 * cols && out_pitch >= (size_t)3 * cols, "oat_bgr2hsv: pitch too small");
    const uint8_t *d;
    size_t dp;
    CKRET(stage_in(c, c->stream, c->scratch_in, bgr, in_pitch, rows, (size_t)3 * cols, &d, &dp));
    OutView ov;
    CKRET(stage_out(c->scratch_out, hsv, out_pitch, rows, (size_t)3 * cols, &ov));
    bgr2hsv_kernel<<<nblocks((long long)rows * cols, 256), 256, 0, c->stream>>>(d, dp, ov.d, ov.dpitch, rows, cols,
                                                                              c->hsv_lut);
    LAUNCH_CHECK(c);
    CKRET(finish_out(c->stream, ov));
    CK(cudaStreamSynchronize(c->stream));
    return OAT_OK;
}

// ---- framefilt bsub ------------------------------------------------------------------------
struct oat_bsub {
    oat_ctx *ctx;
    int rows, cols, ch;
    double alpha;
    bool set;
    uint8_t *bg;
    float *bgf;
    DevBuf in, out;
};

extern "C" int oat_bsub_create(oat_ctx *c, int rows, int cols, int channels, double alpha, oat_bsub **out)
{
    REQUIRE(out, "oat_bsub_create: out is null");
    *out = nullptr;
    CKRET(bind(c));
    REQUIRE(rows > 0 && cols > 0 && (channels == 1 || channels == 3), "oat_bsub_create: bad geometry");
    // reference option range check: adaptation-coeff in [0,1] (BackgroundSubtractor.cpp:77)
    REQUIRE(alpha >= 0.0 && alpha <= 1.0, "oat_bsub_create: alpha must be in [0,1]");
    oat_bsub *b = new (std::nothrow) oat_bsub();
    if (!b) return fail(OAT_ERR_NOMEM, "out of host memory");
    b->ctx = c;
    ctx_ref(c);
    b->rows = rows;
    b->cols = cols;
    b->ch = channels;
    b->alpha = alpha;
    b->set = false;
    b->bg = nullptr;
    b->bgf = nullptr;
    const size_t n = (size_t)rows * cols * channels;
    if (cudaMalloc(&b->bg, n) != cudaSuccess || cudaMalloc(&b->bgf, n * sizeof(float)) != cudaSuccess) {
        if (b->bg) cudaFree(b->bg);
        {
            oat_ctx *owner_ = b->ctx;
            delete b;
            ctx_unref(owner_);
        }
        cudaGetLastError();
        return fail(OAT_ERR_NOMEM, "oat_bsub_create: device allocation failed");
    }
    *out = b;
    return OAT_OK;
}
extern "C" int oat_bsub_destroy(oat_bsub *b)
{
    if (!b) return OAT_OK;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    cudaFree(b->bg);
    cudaFree(b->bgf);
    b->in.release();
    b->out.release();
    {
        oat_ctx *owner_ = b->ctx;
        delete b;
        ctx_unref(owner_);
    }
    return OAT_OK;
}
extern "C" int oat_bsub_set_background(oat_bsub *b, const uint8_t *img, size_t pitch)
{
    REQUIRE(b && img, "oat_bsub_set_background: null argument");
    oat_ctx *c = b->ctx;
    CKRET(bind(c));
    const int rowbytes = b->cols * b->ch;
    REQUIRE(pitch >= (size_t)rowbytes, "oat_bsub_set_background: pitch too small");
    const uint8_t *d;
    size_t dp;
    CKRET(stage_in(c, c->stream, b->in, img, pitch, b->rows, rowbytes, &d, &dp));
    bsub_set_bg_kernel<<<nblocks((long long)b->rows * rowbytes, 256), 256, 0, c->stream>>>(d, dp, b->bg, b->bgf,
                                                                                         b->rows, rowbytes);
    LAUNCH_CHECK(c);
    CK(cudaStreamSynchronize(c->stream));
    b->set = true;
    return OAT_OK;
}
extern "C" int oat_bsub_apply(oat_bsub *b, const uint8_t *in, size_t in_pitch, uint8_t *out, size_t out_pitch)
{
    REQUIRE(b && in && out, "oat_bsub_apply: null argument");
    oat_ctx *c = b->ctx;
    CKRET(bind(c));
    const int rowbytes = b->cols * b->ch;
    REQUIRE(in_pitch >= (size_t)rowbytes && out_pitch >= (size_t)rowbytes, "oat_bsub_apply: pitch too small");
    const uint8_t *d;
    size_t dp;
    CKRET(stage_in(c, c->stream, b->in, in, in_pitch, b->rows, rowbytes, &d, &dp));
    OutView ov;
    CKRET(stage_out(b->out, out, out_pitch, b->rows, rowbytes, &ov));
    const float a = (float)b->alpha, bb = 1.0f - a;
    bsub_kernel<<<nblocks((long long)b->rows * rowbytes, 256), 256, 0, c->stream>>>(
        d, dp, ov.d, ov.dpitch, b->bg, b->bgf, b->rows, rowbytes, b->set ? 0 : 1, a, bb, b->alpha > 0.0 ? 1 : 0);
    LAUNCH_CHECK(c);
    b->set = true;
    CKRET(finish_out(c->stream, ov));
    CK(cudaStreamSynchronize(c->stream));
    return OAT_OK;
}

// ---- detect tail ---------------------------------------------------------------------------
extern "C" void oat_hsv_default_params(oat_hsv_params *p)
{
    if (!p) return;
    p->h_min = p->s_min = p->v_min = 0;
    p->h_max = p->s_max = p->v_max = 256;
    p->erode_px = 0;
    p->dilate_px = 10;
    p->min_area = 0.0;
    p->max_area = DBL_MAX;
}

static int check_hsv_params(const oat_hsv_params *p)
{
    REQUIRE(p, "null hsv params");
    // reference: thresholds are ints in 0..256 with min < max is NOT enforced for H/S/V
    // (HSVDetector.cpp:87-110); area needs min < max (HSVDetector.cpp:130-137).
    const int v[6] = {p->h_min, p->h_max, p->s_min, p->s_max, p->v_min, p->v_max};
    for (int i = 0; i < 6; ++i) REQUIRE(v[i] >= 0 && v[i] <= 256, "HSV thresholds must be in 0..256");
    REQUIRE(p->min_area < p->max_area, "area: min must be < max");
    return OAT_OK;
}

// Per-frame buffers of the one-launch tail (small: a 1080p set is ~270 KB), so the tails of
// consecutive frames can run on different streams while the next fused kernel is in flight.
struct FastBufs {
    uint32_t *di = nullptr;      // post-morphology bits
    int2 *rowext = nullptr;
    int2 *rowcnt = nullptr;      // per-row run counts
    int *bbox = nullptr;         // ymin, ymax
    unsigned int *ticket = nullptr;
    // band pre-labelling pool of the resident tail server (tail_fast.cuh: band_prelabel), allocated on first use
    uint2 *pool_runs = nullptr;
    uint4 *pool_sums = nullptr;
    uint4 *pool_agg = nullptr;
    uint4 *band_hdr = nullptr;
    unsigned int *pool_alloc = nullptr;
    int pool_cap = 0;
    // (the engine carves the pools of all ring slots of a tracker out of ONE allocation: oat_tracker::pool_block)
    // 32 runs per row on average before a frame falls back to the in-place extraction (a 1080p frame of 200 blobs
    // has ~14 k): 40 bytes per run, 1.4 MB per ring slot at 1080p -- allocated for streams whose masks are busy only
    static int pool_cap_for(int rows) { return std::min(std::max(rows * 32, 8192), 1 << 20); }
    static size_t pool_bytes(int rows, int nbands)
    {
        const size_t cap = (size_t)pool_cap_for(rows);
        return cap * (sizeof(uint2) + 2 * sizeof(uint4)) + (((size_t)nbands * sizeof(uint4) + 255) & ~(size_t)255) + 256;
    }
    void set_pool(uint8_t *block, int rows, int nbands)  // block: pool_bytes(rows, nbands) of zeroed device memory, 256-byte aligned
    {
        const size_t cap = (size_t)pool_cap_for(rows);
        pool_sums = reinterpret_cast<uint4 *>(block);
        pool_agg = pool_sums + cap;
        pool_runs = reinterpret_cast<uint2 *>(pool_agg + cap);
        band_hdr = reinterpret_cast<uint4 *>(block + cap * (sizeof(uint2) + 2 * sizeof(uint4)));
        pool_alloc = reinterpret_cast<unsigned int *>(reinterpret_cast<uint8_t *>(band_hdr) + (((size_t)nbands * sizeof(uint4) + 255) & ~(size_t)255));
        pool_cap = (int)cap;
    }
    int create(size_t nwords, int rows)
    {
        CK(cudaMalloc(&di, nwords * 4));
        CK(cudaMalloc(&rowext, (size_t)rows * sizeof(int2)));
        CK(cudaMalloc(&rowcnt, (size_t)rows * sizeof(int2)));
        CK(cudaMalloc(&bbox, 2 * sizeof(int)));
        CK(cudaMalloc(&ticket, sizeof(unsigned int)));
        const int init[2] = {INT_MAX, -1};
        CK(cudaMemcpy(bbox, init, sizeof(init), cudaMemcpyHostToDevice));
        CK(cudaMemset(ticket, 0, sizeof(unsigned int)));
        return OAT_OK;
    }
    void destroy()
    {
        cudaFree(di);
        cudaFree(rowext);
        cudaFree(rowcnt);
        cudaFree(bbox);
        cudaFree(ticket);
        *this = FastBufs();  // (the pool belongs to the tracker: oat_tracker::pool_block)
    }
};

struct Tail {
    TailBuffers tb{};
    uint32_t *bits0 = nullptr;  // threshold mask (input of the tail)
    uint32_t *tmp = nullptr, *er = nullptr, *di = nullptr;
    size_t nwords = 0;
    size_t prep_smem_set = 48 * 1024;
    FastBufs fb;                      // one-launch tail buffers for the synchronous entry points
    size_t fast_smem = 0;             // dynamic shared memory of tail_fast_kernel (0: fast path unusable)
    size_t fast_smem_set = 48 * 1024;
    int fast_comps = 128;

    // smem_kb: shared-memory budget of the one-launch tail (stand-alone detectors can afford a whole SM's worth;
    // the tracker keeps it small enough to co-reside with the fused kernel)
    int create(int rows, int cols, size_t smem_kb = 48)
    {
        REQUIRE(rows > 0 && cols > 0 && rows <= 32768 && cols <= 32768, "detector: bad frame geometry");
        tb.g.rows = rows;
        tb.g.cols = cols;
        tb.g.wpr = div_up(cols, 32);
        nwords = (size_t)rows * tb.g.wpr;
        tb.nnodes = nwords * 32 + 1;
        CK(cudaMalloc(&bits0, nwords * 4));
        CK(cudaMalloc(&tmp, nwords * 4));
        CK(cudaMalloc(&er, nwords * 4));
        CK(cudaMalloc(&di, nwords * 4));
        CK(cudaMalloc(&tb.G, nwords * 4));
        CK(cudaMalloc(&tb.parent, tb.nnodes * sizeof(int)));
        CK(cudaMalloc(&tb.acc, tb.nnodes * 3 * sizeof(unsigned long long)));
        CK(cudaMalloc(&tb.best, sizeof(unsigned long long)));
        CK(cudaMalloc(&tb.count, sizeof(unsigned int)));
        CK(cudaMalloc(&tb.ticket, sizeof(unsigned int)));
        CK(cudaMalloc(&tb.rowext, (size_t)rows * sizeof(int2)));
        CKRET(fb.create(nwords, rows));
        // label-phase budget (mask region + run table + accumulators in shared memory).  48 KB keeps a
        // tail CTA co-resident with two CTAs of the fused kernel (2 x 74.1 KB) on one SM, which is what
        // lets the tail of frame t overlap the fused kernel of frame t+1.  OAT_B200_TAIL_SMEM_KB overrides.
        size_t kb = smem_kb;
        if (const char *e = getenv("OAT_B200_TAIL_SMEM_KB")) kb = (size_t)atoi(e);
        if (kb > 200) kb = 200;
        fast_smem = kb * 1024;
        return OAT_OK;
    }
    void destroy()
    {
        cudaFree(bits0);
        cudaFree(tmp);
        cudaFree(er);
        cudaFree(di);
        cudaFree(tb.G);
        cudaFree(tb.parent);
        cudaFree(tb.acc);
        cudaFree(tb.best);
        cudaFree(tb.count);
        cudaFree(tb.ticket);
        cudaFree(tb.rowext);
        fb.destroy();
        *this = Tail();
    }
    // Fast path: ONE launch (tail_fast.cuh).  Returns false (nothing enqueued) if this geometry /
    // kernel size cannot use it; res->status == TAIL_OVERFLOW after completion means "replay with run()".
    bool run_fast(oat_ctx *c, cudaStream_t stream, const FastBufs &b, const uint32_t *src, const oat_hsv_params &p,
                  TailResult *d_res, uint8_t *thresh_dev, size_t thresh_pitch, int *err, unsigned int *slow_in = nullptr,
                  TailResult *h_mirror = nullptr)
    {
        *err = OAT_OK;
        const BitGeom g = tb.g;
        if (fast_smem == 0 || g.rows < 2 || c->no_fast_tail) return false;
        const int ke = p.erode_px > 0 ? p.erode_px : 0, kd = p.dilate_px > 0 ? p.dilate_px : 0;
        const int R = 8;
        const size_t nin = (size_t)R + (ke > 0 ? ke - 1 : 0) + (kd > 0 ? kd - 1 : 0);
        const size_t stage = 2 * nin * (size_t)g.wpr * sizeof(uint32_t);
        if (stage > 200 * 1024) return false;
        const size_t smem = stage > fast_smem ? stage : fast_smem;
        if (smem + 2048 > fast_smem_set) {
            cudaError_t e = cudaFuncSetAttribute(tail_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
            if (e != cudaSuccess) {
                *err = fail(OAT_ERR_CUDA, std::string("cudaFuncSetAttribute(tail_fast_kernel): ") + cudaGetErrorString(e));
                return true;
            }
            fast_smem_set = 200 * 1024;
        }
        FastArgs fa{};  // (no band pool: the per-frame launch labels from the mask)
        fa.in = src;
        fa.out = b.di;
        fa.ke = ke;
        fa.kd = kd;
        fa.R = R;
        fa.g = g;
        fa.rowext = b.rowext;
        fa.rowcnt = b.rowcnt;
        fa.bbox = b.bbox;
        fa.ticket = b.ticket;
        fa.min_area = p.min_area;
        fa.max_area = p.max_area;
        fa.res = d_res;
        fa.res_host = h_mirror;
        fa.smem_bytes = (int)smem;
        fa.max_comps = fast_comps;
        fa.slow_in = slow_in;
        fa.in_place_ok = 0;  // thresh egress reads b.di after the kernel
        tail_fast_kernel<<<div_up(g.rows, R), 256, smem, stream>>>(fa);
        ++c->launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            *err = fail(OAT_ERR_CUDA, std::string("tail_fast_kernel launch: ") + cudaGetErrorString(e));
            return true;
        }
        if (thresh_dev) {
            bits_to_mask_kernel<<<nblocks((long long)g.rows * g.pitch_px(), 256), 256, 0, stream>>>(b.di, g, thresh_dev,
                                                                                                  thresh_pitch);
            ++c->launches;
        }
        return true;
    }

    // [erode] -> [dilate] -> labelling -> moments -> select; result to d_out (device memory).
    // 5 launches: prep (morphology + row extents + union-find init), merge, fill, moments, select.
    // Unbounded (every structure is in global memory): the fallback of run_fast and the path
    // that can publish per-pixel labels.
    int run(oat_ctx *c, const uint32_t *src_bits, const oat_hsv_params &p, oat_detection *d_out, uint8_t *thresh_dev,
            size_t thresh_pitch, int32_t *labels_dev)
    {
        cudaStream_t s = c->stream;
        const BitGeom g = tb.g;
        const unsigned gw = nblocks((long long)nwords, 256);
        int ke = p.erode_px > 0 ? p.erode_px : 0, kd = p.dilate_px > 0 ? p.dilate_px : 0;
        const uint32_t *src = src_bits;
        const size_t smem_limit = 200 * 1024;
        auto smem_for = [&](int R, int e, int d) {
            const size_t nin = (size_t)R + (e > 0 ? e - 1 : 0) + (d > 0 ? d - 1 : 0);
            return 2 * nin * (size_t)g.wpr * sizeof(uint32_t);
        };
        int R = 8;
        if (smem_for(R, ke, kd) > smem_limit) R = 1;
        if (smem_for(R, ke, kd) > smem_limit) {
            // very wide frame x very large kernel: separable passes through global memory instead
            if (ke > 0) {
                morph_h_kernel<false><<<gw, 256, 0, s>>>(src, tmp, g, ke);
                LAUNCH_CHECK(c);
                morph_v_kernel<false><<<gw, 256, 0, s>>>(tmp, er, g, ke);
                LAUNCH_CHECK(c);
                src = er;
            }
            if (kd > 0) {
                morph_h_kernel<true><<<gw, 256, 0, s>>>(src, tmp, g, kd);
                LAUNCH_CHECK(c);
                uint32_t *dst = (er == src) ? tb.G : er;  // G is rewritten by the fill kernel later
                morph_v_kernel<true><<<gw, 256, 0, s>>>(tmp, dst, g, kd);
                LAUNCH_CHECK(c);
                src = dst;
            }
            ke = kd = 0;
            R = 8;
            if (smem_for(R, 0, 0) > smem_limit) R = 1;
        }
        PrepArgs pa;
        pa.in = src;
        pa.out = di;
        pa.ke = ke;
        pa.kd = kd;
        pa.R = R;
        pa.tb = tb;
        const size_t smem = smem_for(R, ke, kd);
        if (smem > prep_smem_set) {
            CK(cudaFuncSetAttribute(tail_prep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_limit));
            prep_smem_set = smem_limit;
        }
        tail_prep_kernel<<<div_up(g.rows, R), 256, smem, s>>>(pa);
        LAUNCH_CHECK(c);
        const uint32_t *cur = di;
        if (thresh_dev) {
            bits_to_mask_kernel<<<nblocks((long long)g.rows * g.pitch_px(), 256), 256, 0, s>>>(cur, g, thresh_dev,
                                                                                             thresh_pitch);
            LAUNCH_CHECK(c);
        }
        ccl_merge_kernel<<<gw, 256, 0, s>>>(cur, tb);
        LAUNCH_CHECK(c);
        if (labels_dev) {
            ccl_labels_kernel<<<gw, 256, 0, s>>>(cur, tb, labels_dev);
            LAUNCH_CHECK(c);
        }
        ccl_fill_kernel<<<gw, 256, 0, s>>>(cur, tb);
        LAUNCH_CHECK(c);
        if (g.rows > 1) {
            ccl_moments_kernel<<<nblocks((long long)(g.rows - 1) * g.wpr, 256), 256, 0, s>>>(tb);
            LAUNCH_CHECK(c);
        }
        ccl_select_kernel<<<gw, 256, 0, s>>>(cur, tb, p.min_area, p.max_area, d_out);
        LAUNCH_CHECK(c);
        return OAT_OK;
    }
};

struct oat_hsvdet {
    oat_ctx *ctx;
    Tail tail;
    TailResult *d_res;
    DevBuf in, out_thr, out_lab;
};

extern "C" int oat_hsvdet_create(oat_ctx *c, int rows, int cols, oat_hsvdet **out)
{
    REQUIRE(out, "oat_hsvdet_create: out is null");
    *out = nullptr;
    CKRET(bind(c));
    oat_hsvdet *h = new (std::nothrow) oat_hsvdet();
    if (!h) return fail(OAT_ERR_NOMEM, "out of host memory");
    h->ctx = c;
    ctx_ref(c);
    h->d_res = nullptr;
    int r = h->tail.create(rows, cols, 160);
    if (r == OAT_OK && cudaMalloc(&h->d_res, sizeof(TailResult)) != cudaSuccess)
        r = fail(OAT_ERR_NOMEM, "device allocation failed");
    if (r != OAT_OK) {
        h->tail.destroy();
        {
            oat_ctx *owner_ = h->ctx;
            delete h;
            ctx_unref(owner_);
        }
        return r;
    }
    *out = h;
    return OAT_OK;
}
extern "C" int oat_hsvdet_destroy(oat_hsvdet *h)
{
    if (!h) return OAT_OK;
    cudaSetDevice(h->ctx->device);
    cudaStreamSynchronize(h->ctx->stream);
    h->tail.destroy();
    cudaFree(h->d_res);
    h->in.release();
    h->out_thr.release();
    h->out_lab.release();
    {
        oat_ctx *owner_ = h->ctx;
        delete h;
        ctx_unref(owner_);
    }
    return OAT_OK;
}

// channels 3: HSV frame + bands of p; 1 with t_min >= 0: grey frame + [t_min, t_max]; 1 with t_min < 0: binary mask
static int detect_common(oat_hsvdet *h, const uint8_t *img, size_t pitch, int channels, const oat_hsv_params *p,
                         oat_detection *out, uint8_t *thresh_out, size_t thresh_pitch, int32_t *labels_out, int t_min = -1,
                         int t_max = -1)
{
    REQUIRE(h && img && out, "detect: null argument");
    CKRET(check_hsv_params(p));
    oat_ctx *c = h->ctx;
    CKRET(bind(c));
    const BitGeom g = h->tail.tb.g;
    REQUIRE(pitch >= (size_t)channels * g.cols, "detect: input pitch too small");
    REQUIRE(!thresh_out || thresh_pitch >= (size_t)g.cols, "detect: thresh pitch too small");
    const uint8_t *d;
    size_t dp;
    CKRET(stage_in(c, c->stream, h->in, img, pitch, g.rows, (size_t)channels * g.cols, &d, &dp));
    OutView ot;
    CKRET(stage_out(h->out_thr, thresh_out, thresh_pitch, g.rows, (size_t)g.cols, &ot));
    int32_t *lab_dev = nullptr;
    const size_t lab_bytes = (size_t)g.rows * g.cols * sizeof(int32_t);
    if (labels_out) {
        if (mem_kind(labels_out) == MEM_DEVICE)
            lab_dev = labels_out;
        else {
            CKRET(h->out_lab.ensure(lab_bytes));
            lab_dev = (int32_t *)h->out_lab.p;
        }
    }
    const unsigned gp = nblocks((long long)g.rows * g.pitch_px(), 256);
    if (channels == 3)
        inrange_bits_kernel<<<gp, 256, 0, c->stream>>>(d, dp, g, p->h_min, p->s_min, p->v_min, p->h_max, p->s_max,
                                                       p->v_max, h->tail.bits0);
    else if (t_min >= 0)
        inrange1_bits_kernel<<<gp, 256, 0, c->stream>>>(d, dp, g, t_min, t_max, h->tail.bits0);
    else
        mask_to_bits_kernel<<<gp, 256, 0, c->stream>>>(d, dp, g, h->tail.bits0);
    LAUNCH_CHECK(c);
    // one-launch fast path unless per-pixel labels are wanted; replay through the unbounded
    // path if the mask overflowed the shared-memory run table
    bool need_generic = true;
    if (!labels_out) {
        int err = OAT_OK;
        if (h->tail.run_fast(c, c->stream, h->tail.fb, h->tail.bits0, *p, h->d_res, ot.d, ot.dpitch, &err)) {
            CKRET(err);
            TailResult tr;
            CK(cudaMemcpyAsync(&tr, h->d_res, sizeof(tr), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            if (tr.status == TAIL_OK) {
                *out = tr.det;
                need_generic = false;
            }
        }
    }
    if (need_generic) {
        CKRET(h->tail.run(c, h->tail.bits0, *p, &h->d_res->det, ot.d, ot.dpitch, lab_dev));
        CK(cudaMemcpyAsync(out, &h->d_res->det, sizeof(oat_detection), cudaMemcpyDeviceToHost, c->stream));
    }
    CKRET(finish_out(c->stream, ot));
    if (labels_out && lab_dev != labels_out)
        CK(cudaMemcpyAsync(labels_out, lab_dev, lab_bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OAT_OK;
}

extern "C" int oat_hsvdet_detect(oat_hsvdet *h, const uint8_t *hsv, size_t pitch, const oat_hsv_params *p,
                                 oat_detection *out, uint8_t *thresh_out, size_t thresh_pitch, int32_t *labels_out)
{
    return detect_common(h, hsv, pitch, 3, p, out, thresh_out, thresh_pitch, labels_out);
}
extern "C" int oat_sift_contours(oat_hsvdet *h, const uint8_t *mask, size_t pitch, const oat_hsv_params *p,
                                 oat_detection *out, uint8_t *thresh_out, size_t thresh_pitch, int32_t *labels_out)
{
    return detect_common(h, mask, pitch, 1, p, out, thresh_out, thresh_pitch, labels_out);
}

extern "C" int oat_thresh_detect(oat_hsvdet *h, const uint8_t *grey, size_t pitch, int t_min, int t_max,
                                 const oat_hsv_params *p, oat_detection *out, uint8_t *thresh_out, size_t thresh_pitch,
                                 int32_t *labels_out)
{
    // SimpleThreshold.cpp:76-84: "Values of thresh should be between 0 and 256."
    REQUIRE(t_min >= 0 && t_min <= 256 && t_max >= 0 && t_max <= 256, "Values of thresh should be between 0 and 256.");
    return detect_common(h, grey, pitch, 1, p, out, thresh_out, thresh_pitch, labels_out, t_min, t_max);
}

// ---- posidet diff ------------------------------------------------------------------------------
struct oat_diffdet {
    oat_ctx *ctx;
    Tail tail;
    TailResult *d_res;
    uint8_t *last;  // previous frame (tight rows x cols)
    bool have_last;
    DevBuf in, out_thr;
};
extern "C" int oat_diffdet_create(oat_ctx *c, int rows, int cols, oat_diffdet **out)
{
    REQUIRE(out, "oat_diffdet_create: out is null");
    *out = nullptr;
    CKRET(bind(c));
    oat_diffdet *h = new (std::nothrow) oat_diffdet();
    if (!h) return fail(OAT_ERR_NOMEM, "out of host memory");
    h->ctx = c;
    ctx_ref(c);
    h->d_res = nullptr;
    h->last = nullptr;
    h->have_last = false;
    int r = h->tail.create(rows, cols, 160);
    if (r == OAT_OK && (cudaMalloc(&h->d_res, sizeof(TailResult)) != cudaSuccess || cudaMalloc(&h->last, (size_t)rows * cols) != cudaSuccess))
        r = fail(OAT_ERR_NOMEM, "device allocation failed");
    if (r != OAT_OK) {
        h->tail.destroy();
        cudaFree(h->d_res);
        cudaFree(h->last);
        {
            oat_ctx *owner_ = h->ctx;
            delete h;
            ctx_unref(owner_);
        }
        return r;
    }
    *out = h;
    return OAT_OK;
}
extern "C" int oat_diffdet_destroy(oat_diffdet *h)
{
    if (!h) return OAT_OK;
    cudaSetDevice(h->ctx->device);
    cudaStreamSynchronize(h->ctx->stream);
    h->tail.destroy();
    cudaFree(h->d_res);
    cudaFree(h->last);
    h->in.release();
    h->out_thr.release();
    {
        oat_ctx *owner_ = h->ctx;
        delete h;
        ctx_unref(owner_);
    }
    return OAT_OK;
}
extern "C" int oat_diffdet_reset(oat_diffdet *h)
{
    REQUIRE(h, "null handle");
    h->have_last = false;
    return OAT_OK;
}
extern "C" int oat_diffdet_detect(oat_diffdet *h, const uint8_t *grey, size_t pitch, int diff_threshold, int blur_px, double min_area,
                                  double max_area, oat_detection *out, uint8_t *thresh_out, size_t thresh_pitch)
{
    REQUIRE(h && grey && out, "oat_diffdet_detect: null argument");
    REQUIRE(diff_threshold >= 0 && blur_px >= 0, "diff-threshold and blur must be >= 0");
    REQUIRE(min_area < max_area, "area: min must be < max");
    oat_ctx *c = h->ctx;
    CKRET(bind(c));
    const BitGeom g = h->tail.tb.g;
    // cv::blur of a 0/255 image is non-zero wherever ONE set pixel falls in the box only while 255/k^2 rounds up
    if (blur_px > 22 || blur_px > g.rows || blur_px > g.cols)
        return fail(OAT_ERR_UNSUPPORTED, "posidet diff: blur sizes above 22 (or above the frame size) are not implemented");
    REQUIRE(pitch >= (size_t)g.cols, "detect: input pitch too small");
    REQUIRE(!thresh_out || thresh_pitch >= (size_t)g.cols, "detect: thresh pitch too small");
    const uint8_t *d;
    size_t dp;
    CKRET(stage_in(c, c->stream, h->in, grey, pitch, g.rows, (size_t)g.cols, &d, &dp));
    OutView ot;
    CKRET(stage_out(h->out_thr, thresh_out, thresh_pitch, g.rows, (size_t)g.cols, &ot));
    cudaStream_t s = c->stream;
    const unsigned gp = nblocks((long long)g.rows * g.pitch_px(), 256), gw = nblocks((long long)h->tail.nwords, 256);
    absdiff_bits_kernel<<<gp, 256, 0, s>>>(d, dp, h->last, g, diff_threshold, h->have_last ? 0 : 1, h->tail.bits0);
    LAUNCH_CHECK(c);
    const uint32_t *cur = h->tail.bits0;
    if (h->have_last && blur_px > 0) {  // (blurred != 0) == box dilation with cv::blur's anchor and reflected border
        const int a = blur_px / 2;
        morph_h_kernel<true><<<gw, 256, 0, s>>>(h->tail.bits0, h->tail.tmp, g, blur_px);
        LAUNCH_CHECK(c);
        if (blur_px % 2 == 0) {
            reflect_even_fix_kernel<<<nblocks(g.rows, 256), 256, 0, s>>>(h->tail.bits0, h->tail.tmp, g, a, 0);
            LAUNCH_CHECK(c);
        }
        morph_v_kernel<true><<<gw, 256, 0, s>>>(h->tail.tmp, h->tail.er, g, blur_px);
        LAUNCH_CHECK(c);
        if (blur_px % 2 == 0) {
            reflect_even_fix_kernel<<<nblocks(g.wpr, 256), 256, 0, s>>>(h->tail.tmp, h->tail.er, g, a, 1);
            LAUNCH_CHECK(c);
        }
        cur = h->tail.er;
    }
    h->have_last = true;
    oat_hsv_params p;
    oat_hsv_default_params(&p);
    p.erode_px = p.dilate_px = 0;
    p.min_area = min_area;
    p.max_area = max_area;
    bool need_generic = true;
    int err = OAT_OK;
    if (h->tail.run_fast(c, s, h->tail.fb, cur, p, h->d_res, ot.d, ot.dpitch, &err)) {
        CKRET(err);
        TailResult tr;
        CK(cudaMemcpyAsync(&tr, h->d_res, sizeof(tr), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        if (tr.status == TAIL_OK) {
            *out = tr.det;
            need_generic = false;
        }
    }
    if (need_generic) {
        CKRET(h->tail.run(c, cur, p, &h->d_res->det, ot.d, ot.dpitch, nullptr));
        CK(cudaMemcpyAsync(out, &h->d_res->det, sizeof(oat_detection), cudaMemcpyDeviceToHost, s));
    }
    CKRET(finish_out(s, ot));
    CK(cudaStreamSynchronize(s));
    return OAT_OK;
}

// framefilt thresh / framefilt mask: zero the pixels outside an intensity band / a region-of-interest mask
extern "C" int oat_keep_where(oat_ctx *c, const uint8_t *in, size_t in_pitch, uint8_t *out, size_t out_pitch, int rows, int cols,
                              int channels, const uint8_t *roi, size_t roi_pitch, int i_min, int i_max)
{
    CKRET(bind(c));
    REQUIRE(in && out && rows > 0 && cols > 0 && (channels == 1 || channels == 3), "oat_keep_where: bad arguments");
    REQUIRE(in_pitch >= (size_t)channels * cols && out_pitch >= (size_t)channels * cols, "oat_keep_where: pitch too small");
    REQUIRE(roi || (i_min >= 0 && i_min <= 256 && i_max >= 0 && i_max <= 256), "Values of intensity should be between 0 and 256.");
    REQUIRE(!roi || roi_pitch >= (size_t)cols, "oat_keep_where: mask pitch too small");
    const uint8_t *d;
    size_t dp;
    CKRET(stage_in(c, c->stream, c->scratch_in, in, in_pitch, rows, (size_t)channels * cols, &d, &dp));
    OutView ov;
    CKRET(stage_out(c->scratch_out, out, out_pitch, rows, (size_t)channels * cols, &ov));
    const uint8_t *droi = nullptr;
    size_t droi_pitch = 0;
    if (roi) CKRET(stage_in(c, c->stream, c->scratch_roi, roi, roi_pitch, rows, (size_t)cols, &droi, &droi_pitch));
    keep_where_kernel<<<nblocks((long long)rows * cols, 256), 256, 0, c->stream>>>(d, dp, ov.d, ov.dpitch, rows, cols, channels,
                                                                                 droi, droi_pitch, i_min, i_max);
    LAUNCH_CHECK(c);
    CKRET(finish_out(c->stream, ov));
    CK(cudaStreamSynchronize(c->stream));
    return OAT_OK;
}

// ---- fused tracker -------------------------------------------------------------------------
struct Slot {
    DevBuf in;                      // staged input frame (host-fed streams)
    DevBuf in_raw;                  // ... its linear landing area when the rows have to be re-pitched on the device
    TailResult *h_res = nullptr;    // pinned
    TailResult *d_res = nullptr;
    uint32_t *bits = nullptr;       // this frame's threshold mask (kept until collect: overflow replay)
    FastBufs fb;                    // this frame's one-launch tail buffers
    unsigned int *d_slow = nullptr; // fused kernel's slow-path census of this frame (re-armed by the tail)
    unsigned int *d_done = nullptr; // resident path: tiles of this slot's frames published so far (monotonic)
    unsigned int done_total = 0;    //                host mirror: value of *d_done once every frame enqueued so far is complete
    cudaEvent_t fused_done = nullptr;
    oat_hsv_params hp{};
    bool fast = false;              // the one-launch tail was used (status must be checked at collect)
    int ingest = 0;                 // how the frame's input is consumed: 1 = H2D copy (event `copied`), 2 = read in place by the fused kernel (event `fused_done`)
    cudaEvent_t copied = nullptr, done = nullptr;
    oat_position *d_pos = nullptr, *h_pos = nullptr;  // attached position filter: this frame's filtered position
    cudaEvent_t pos_done = nullptr;
    bool has_pos = false;
};

// ---- posifilt kalman + posicom mean (posfilt.cuh) ----------------------------------------------
struct oat_posfilt {
    oat_ctx *ctx;
    int n = 1, combine = 0, heading_anchor = -1;
    KalmanConsts kc{};
    PosfiltDev *dev = nullptr;
    oat_position *d_raw = nullptr, *d_out = nullptr;  // staging of oat_posfilt_apply
    oat_position *h_io = nullptr;                     // pinned: [0..n) raw in, [n..2n) out
    int attached = 0;
};

extern "C" void oat_kalman_default_params(oat_kalman_params *p)
{
    if (!p) return;
    p->dt = 0.02;  // KalmanFilter2D.h:66-70
    p->timeout = 0.0;
    p->sigma_accel = 5.0;
    p->sigma_noise = 0.0;
}

extern "C" int oat_posfilt_create(oat_ctx *c, int n_sources, const oat_kalman_params *kp, int combine_mean,
                                  int heading_anchor, oat_posfilt **out)
{
    REQUIRE(c && out, "oat_posfilt_create: null argument");
    REQUIRE(n_sources >= 1 && n_sources <= 8, "oat_posfilt_create: 1..8 sources");
    REQUIRE(heading_anchor < n_sources, "heading-anchor must be a SOURCE index");  // MeanPosition.cpp:52-54
    REQUIRE(!kp || kp->dt > 0.0, "kalman: dt must be positive");
    CKRET(bind(c));
    oat_posfilt *f = new (std::nothrow) oat_posfilt();
    REQUIRE(f, "out of memory");
    f->ctx = c;
    ctx_ref(c);
    f->n = n_sources;
    f->combine = combine_mean ? 1 : 0;
    f->heading_anchor = heading_anchor < 0 ? -1 : heading_anchor;
    if (kp) {  // initializeStaticMatracies, KalmanFilter2D.cpp:162-200
        const double dt = kp->dt, sa = kp->sigma_accel;
        f->kc.enabled = 1;
        f->kc.dt = dt;
        f->kc.q00 = sa * sa * (dt * dt * dt * dt) / 4.0;
        f->kc.q01 = sa * sa * (dt * dt * dt) / 2.0;
        f->kc.q11 = sa * sa * (dt * dt);
        f->kc.r = kp->sigma_noise * kp->sigma_noise;
        f->kc.not_found_thr = (int)(kp->timeout / kp->dt);  // KalmanFilter2D.cpp:74-76
    }
    CK(cudaMalloc(&f->dev, sizeof(PosfiltDev)));
    CK(cudaMalloc(&f->d_raw, 8 * sizeof(oat_position)));
    CK(cudaMalloc(&f->d_out, 8 * sizeof(oat_position)));
    CK(cudaHostAlloc(&f->h_io, 16 * sizeof(oat_position), cudaHostAllocDefault));
    posfilt_reset_kernel<<<1, 32, 0, c->post>>>(f->dev);
    const unsigned long long keep = c->chain_uid;
    LAUNCH_CHECK(c);
    c->chain_uid = keep;
    *out = f;
    return OAT_OK;
}

extern "C" int oat_posfilt_destroy(oat_posfilt *f)
{
    if (!f) return OAT_OK;
    REQUIRE(!f->attached, "oat_posfilt_destroy: still attached to a tracker");
    bind(f->ctx);
    cudaStreamSynchronize(f->ctx->post);
    if (f->dev) cudaFree(f->dev);
    if (f->d_raw) cudaFree(f->d_raw);
    if (f->d_out) cudaFree(f->d_out);
    if (f->h_io) cudaFreeHost(f->h_io);
    {
        oat_ctx *owner_ = f->ctx;
        delete f;
        ctx_unref(owner_);
    }
    return OAT_OK;
}

extern "C" int oat_posfilt_reset(oat_posfilt *f)
{
    REQUIRE(f, "null handle");
    oat_ctx *c = f->ctx;
    CKRET(bind(c));
    posfilt_reset_kernel<<<1, 32, 0, c->post>>>(f->dev);
    const unsigned long long keep = c->chain_uid;
    LAUNCH_CHECK(c);
    c->chain_uid = keep;
    return OAT_OK;
}

static int posfilt_launch(oat_posfilt *f, const oat_position *raw, const oat_detection *det, const int32_t *status,
                          int force, oat_position *out)
{
    oat_ctx *c = f->ctx;
    PosfiltArgs a{};
    a.dev = f->dev;
    a.kc = f->kc;
    a.n = f->n;
    a.combine = f->combine;
    a.heading_anchor = f->heading_anchor;
    a.raw = raw;
    a.det = det;
    a.det_status = status;
    a.force = force;
    a.out = out;
    posfilt_kernel<<<1, 32, 0, c->post>>>(a);
    const unsigned long long keep = c->chain_uid;  // not on the compute stream: the fused-kernel chain is intact
    LAUNCH_CHECK(c);
    c->chain_uid = keep;
    return OAT_OK;
}

extern "C" int oat_posfilt_apply(oat_posfilt *f, const oat_position *sources, oat_position *out)
{
    REQUIRE(f && sources && out, "oat_posfilt_apply: null argument");
    REQUIRE(!f->attached, "oat_posfilt_apply: the filter is fed by a tracker");
    oat_ctx *c = f->ctx;
    CKRET(bind(c));
    const int nout = f->combine ? 1 : f->n;
    memcpy(f->h_io, sources, f->n * sizeof(oat_position));
    CK(cudaMemcpyAsync(f->d_raw, f->h_io, f->n * sizeof(oat_position), cudaMemcpyHostToDevice, c->post));
    CKRET(posfilt_launch(f, f->d_raw, nullptr, nullptr, 1, f->d_out));
    CK(cudaMemcpyAsync(f->h_io + 8, f->d_out, nout * sizeof(oat_position), cudaMemcpyDeviceToHost, c->post));
    CK(cudaStreamSynchronize(c->post));
    memcpy(out, f->h_io + 8, nout * sizeof(oat_position));
    return OAT_OK;
}

struct oat_tracker {
    oat_ctx *ctx;
    oat_posfilt *pf = nullptr;
    MogModel m;
    Tail tail;
    std::vector<Slot> ring;
    uint64_t head = 0, tailpos = 0;  // submitted / collected
    uint64_t replays = 0;            // frames whose tail had to be replayed through the unbounded path
    // kernel choice: the pipelined fused kernel assumes most pixels take its fast path (<= 2 live modes, sample
    // fits the heaviest); a stream where that fails (busy multi-modal scenes) is faster on the generic kernel
    double slow_frac = 0.0;          // moving estimate of the fraction of 4-pixel groups leaving the fast path
    bool use_generic = false;
    uint64_t generic_frames = 0;
    DevBuf out_bgr, out_fg, out_hsv, out_thr;
    // profiling of the fused kernel
    int prof = 0;
    cudaEvent_t pe0 = nullptr, pe1 = nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_pending;
    double prof_ms = 0.0;
    uint64_t prof_n = 0;
    size_t last_slot = 0;            // ring slot of the most recently collected frame (oat_tracker_tail_stats)
    uint64_t clip_frames = 0;        // frames that went through the resident clip engine
    double tail_load = 0.0;          // moving estimate of the run-table entries a frame's mask needs (sizes the tail server's share)
    void *pool_block = nullptr;      // band pre-labelling pools of every ring slot (allocated when the stream's masks get busy)
    struct StreamState *stream = nullptr;  // oat_tracker_stream_*: the resident engine kept alive between calls
};
static bool stream_busy(const oat_tracker *t);
static void stream_destroy(oat_tracker *t);
#define REQUIRE_NOT_STREAMING(t, who) REQUIRE(!stream_busy(t), who ": frames are gathered or in flight in the tracker's stream (poll them first)")

extern "C" int oat_tracker_create(oat_ctx *c, int rows, int cols, const oat_mog_params *mp, int ring_depth,
                                  oat_tracker **out)
{
    REQUIRE(out, "oat_tracker_create: out is null");
    *out = nullptr;
    CKRET(bind(c));
    REQUIRE(ring_depth >= 0 && ring_depth <= 64, "oat_tracker_create: ring_depth must be in 0..64");
    if (ring_depth == 0) ring_depth = 4;
    oat_tracker *t = new (std::nothrow) oat_tracker();
    if (!t) return fail(OAT_ERR_NOMEM, "out of host memory");
    t->ctx = c;
    ctx_ref(c);
    int r = t->m.create(rows, cols, mp);
    if (r == OAT_OK) r = t->tail.create(rows, cols);
    if (r == OAT_OK) {
        t->ring.resize(ring_depth);
        for (auto &s : t->ring) {
            s.done_total = 0;
            if (cudaHostAlloc(&s.h_res, sizeof(TailResult), cudaHostAllocMapped) != cudaSuccess ||
                cudaMalloc(&s.d_res, sizeof(TailResult)) != cudaSuccess ||
                cudaMalloc(&s.bits, t->tail.nwords * 4) != cudaSuccess ||
                s.fb.create(t->tail.nwords, rows) != OAT_OK || cudaMalloc(&s.d_slow, 2 * sizeof(unsigned int)) != cudaSuccess ||
                cudaMemset(s.d_slow, 0, 2 * sizeof(unsigned int)) != cudaSuccess ||
                cudaEventCreateWithFlags(&s.fused_done, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming) != cudaSuccess) {
                r = fail(OAT_ERR_NOMEM, std::string("tracker ring allocation failed: ") +
                                            cudaGetErrorString(cudaGetLastError()));
                break;
            }
        }
    }
    if (r == OAT_OK) {
        for (auto &s : t->ring) s.d_done = s.d_slow + 1;
        if (cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess)  // the memsets ran on the legacy stream
            r = fail(OAT_ERR_CUDA, "tracker ring initialisation failed");
    }
    if (r != OAT_OK) {
        oat_tracker_destroy(t);
        return r;
    }
    *out = t;
    return OAT_OK;
}

extern "C" int oat_tracker_destroy(oat_tracker *t)
{
    if (!t) return OAT_OK;
    cudaSetDevice(t->ctx->device);
    stream_destroy(t);
    cudaStreamSynchronize(t->ctx->h2d);
    cudaStreamSynchronize(t->ctx->stream);
    for (int i = 0; i < oat_ctx::NTAIL; ++i) cudaStreamSynchronize(t->ctx->tail[i]);
    cudaStreamSynchronize(t->ctx->post);
    t->m.destroy();
    t->tail.destroy();
    for (auto &s : t->ring) {
        s.in.release();
        s.in_raw.release();
        if (s.h_res) cudaFreeHost(s.h_res);
        if (s.d_res) cudaFree(s.d_res);
        if (s.bits) cudaFree(s.bits);
        s.fb.destroy();
        if (s.d_slow) cudaFree(s.d_slow);
        if (s.fused_done) cudaEventDestroy(s.fused_done);
        if (s.copied) cudaEventDestroy(s.copied);
        if (s.done) cudaEventDestroy(s.done);
        if (s.pos_done) cudaEventDestroy(s.pos_done);
        if (s.d_pos) cudaFree(s.d_pos);
        if (s.h_pos) cudaFreeHost(s.h_pos);
    }
    if (t->pool_block) cudaFree(t->pool_block);
    if (t->pf) t->pf->attached = 0;
    for (auto &pr : t->prof_pending) {
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    t->out_bgr.release();
    t->out_fg.release();
    t->out_hsv.release();
    t->out_thr.release();
    {
        oat_ctx *owner_ = t->ctx;
        delete t;
        ctx_unref(owner_);
    }
    return OAT_OK;
}

extern "C" int oat_tracker_reset(oat_tracker *t)
{
    REQUIRE(t, "null handle");
    REQUIRE(t->head == t->tailpos, "oat_tracker_reset: frames are still outstanding");
    REQUIRE_NOT_STREAMING(t, "oat_tracker_reset");
    t->m.nframes = 0;
    return OAT_OK;
}

// overlap: run this frame's detect tail on one of the context's tail streams so that it overlaps the
// fused kernel of the next frame (submit/collect); otherwise everything stays on the compute stream.
static int tracker_enqueue(oat_tracker *t, const uint8_t *bgr_in, size_t in_pitch, double learning_rate,
                           const oat_hsv_params *p, OutView &ob, OutView &ofg, OutView &ohsv, OutView &othr, bool overlap)
{
    oat_ctx *c = t->ctx;
    const int rows = t->m.g.rows, cols = t->m.g.cols;
    Slot &s = t->ring[t->head % t->ring.size()];
    FusedArgs a{};
    if (mem_kind(bgr_in) == MEM_DEVICE) {
        a.bgr = bgr_in;
        a.in_pitch = in_pitch;
        s.ingest = 2;
    } else {
        // ingest on the copy stream so the DMA of frame t+1 overlaps the kernels of frame t
        CKRET(stage_in(c, c->h2d, s.in, bgr_in, in_pitch, rows, (size_t)3 * cols, &a.bgr, &a.in_pitch, &s.in_raw));
        CK(cudaEventRecord(s.copied, c->h2d));
        CK(cudaStreamWaitEvent(c->stream, s.copied, 0));
        s.ingest = 1;
    }
    t->m.frame_consts(learning_rate, &a.c, &a.reset);
    a.do_hsv = 1;
    a.lo[0] = p->h_min;
    a.lo[1] = p->s_min;
    a.lo[2] = p->v_min;
    a.hi[0] = p->h_max;
    a.hi[1] = p->s_max;
    a.hi[2] = p->v_max;
    a.thr_bits = s.bits;
    s.hp = *p;
    a.bgr_out = ob.d;
    a.bgr_out_pitch = ob.dpitch;
    a.fg_out = ofg.d;
    a.fg_pitch = ofg.dpitch;
    a.hsv_out = ohsv.d;
    a.hsv_pitch = ohsv.dpitch;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (t->prof) {
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, c->stream));
    }
    a.slow_count = s.d_slow;
    if (t->use_generic) ++t->generic_frames;
    // tile-granular chaining to the previous frame's launch: only when nothing but this stream's own
    // kernels order the frame (device-resident input, no egress buffers that a copy still reads)
    const bool chainable = mem_kind(bgr_in) == MEM_DEVICE && !ob.d && !ofg.d && !ohsv.d && !othr.d && !t->prof;
    CKRET(launch_fused(c, t->m, a, !t->use_generic, chainable, true));
    const unsigned long long chain_after_fused = c->chain_uid;
    if (t->prof) {
        CK(cudaEventRecord(e1, c->stream));
        t->prof_pending.emplace_back(e0, e1);
    }
    CKRET(finish_out(c->stream, ob));
    CKRET(finish_out(c->stream, ofg));
    CKRET(finish_out(c->stream, ohsv));
    cudaStream_t ts = c->stream;
    CK(cudaEventRecord(s.fused_done, c->stream));  // (also what oat_tracker_wait_ingest waits for when the frame is read in place)
    if (overlap && !c->no_overlap) {
        ts = c->tail[c->tail_rr++ % oat_ctx::NTAIL];  // context-wide round robin: trackers sharing a context do not pile onto one stream
        CK(cudaStreamWaitEvent(ts, s.fused_done, 0));
    }
    {
        int err = OAT_OK;
        s.fast = t->tail.run_fast(c, ts, s.fb, s.bits, *p, s.d_res, othr.d, othr.dpitch, &err, s.d_slow, c->no_mirror ? nullptr : s.h_res);
        CKRET(err);
        if (!s.fast) {
            ts = c->stream;  // the unbounded path owns shared buffers: compute stream only
            CKRET(t->tail.run(c, s.bits, *p, &s.d_res->det, othr.d, othr.dpitch, nullptr));
        }
    }
    // a tail that went to another stream left the fused kernel the last kernel on the compute stream
    if (ts != c->stream) c->chain_uid = chain_after_fused;
    CKRET(finish_out(ts, othr));
    // the one-launch tail wrote its result into the pinned mirror itself
    if (!s.fast || c->no_mirror) CK(cudaMemcpyAsync(s.h_res, s.d_res, sizeof(TailResult), cudaMemcpyDeviceToHost, ts));
    CK(cudaEventRecord(s.done, ts));
    s.has_pos = false;
    if (t->pf) {
        // the position epilogue: one warp on the in-order `post` stream behind this frame's tail.  A frame
        // whose one-launch tail overflowed has no final detection yet: the kernel holds the filter (and
        // every later frame's update) until collect has replayed the frame and re-runs them in order.
        CK(cudaStreamWaitEvent(c->post, s.done, 0));
        CKRET(posfilt_launch(t->pf, nullptr, &s.d_res->det, s.fast ? &s.d_res->status : nullptr, 0, s.d_pos));
        CK(cudaMemcpyAsync(s.h_pos, s.d_pos, sizeof(oat_position), cudaMemcpyDeviceToHost, c->post));
        CK(cudaEventRecord(s.pos_done, c->post));
        s.has_pos = true;
    }
    ++t->head;
    return OAT_OK;
}

static int tracker_check(oat_tracker *t, const uint8_t *bgr_in, size_t in_pitch, const oat_hsv_params *p)
{
    REQUIRE(t && bgr_in, "tracker: null handle or input");
    CKRET(check_hsv_params(p));
    REQUIRE(in_pitch >= (size_t)3 * t->m.g.cols, "tracker: input pitch too small");
    return OAT_OK;
}

// Diagnostic: enqueue ONLY the fused MOG+HSV+threshold kernel of the next frame (the GMM state advances
// exactly as in oat_tracker_submit; the detect tail is not run and there is nothing to collect).
// Lets a harness time back-to-back launches of the dominant kernel in isolation.
extern "C" int oat_tracker_submit_fused_only(oat_tracker *t, const uint8_t *bgr_in, size_t in_pitch,
                                             double learning_rate, const oat_hsv_params *p)
{
    CKRET(tracker_check(t, bgr_in, in_pitch, p));
    CKRET(bind(t->ctx));
    REQUIRE(t->head == t->tailpos, "oat_tracker_submit_fused_only: frames are still outstanding (collect first)");
    REQUIRE_NOT_STREAMING(t, "oat_tracker_submit_fused_only");
    REQUIRE(mem_kind(bgr_in) == MEM_DEVICE, "oat_tracker_submit_fused_only: device-resident frames only");
    oat_ctx *c = t->ctx;
    Slot &s = t->ring[t->head % t->ring.size()];
    FusedArgs a{};
    a.bgr = bgr_in;
    a.in_pitch = in_pitch;
    t->m.frame_consts(learning_rate, &a.c, &a.reset);
    a.do_hsv = 1;
    a.lo[0] = p->h_min;
    a.lo[1] = p->s_min;
    a.lo[2] = p->v_min;
    a.hi[0] = p->h_max;
    a.hi[1] = p->s_max;
    a.hi[2] = p->v_max;
    a.thr_bits = s.bits;
    return launch_fused(c, t->m, a, true, true);
}

static int tracker_collect(oat_tracker *t, oat_detection *out, oat_position *pos);
extern "C" int oat_tracker_collect(oat_tracker *t, oat_detection *out)
{
    REQUIRE(t && out, "oat_tracker_collect: null argument");
    return tracker_collect(t, out, nullptr);
}

extern "C" int oat_tracker_collect_position(oat_tracker *t, oat_detection *det, oat_position *pos)
{
    REQUIRE(t && det && pos, "oat_tracker_collect_position: null argument");
    REQUIRE(t->pf, "oat_tracker_collect_position: no position filter attached");
    return tracker_collect(t, det, pos);
}

extern "C" int oat_tracker_attach_posfilt(oat_tracker *t, oat_posfilt *f)
{
    REQUIRE(t, "null handle");
    REQUIRE(t->head == t->tailpos, "oat_tracker_attach_posfilt: frames are still outstanding");
    REQUIRE_NOT_STREAMING(t, "oat_tracker_attach_posfilt");
    REQUIRE(!f || (f->ctx == t->ctx && f->n == 1 && !f->attached), "oat_tracker_attach_posfilt: needs an unattached single-source filter of the same context");
    CKRET(bind(t->ctx));
    if (t->pf) t->pf->attached = 0;
    t->pf = f;
    if (f) {
        f->attached = 1;
        for (auto &s : t->ring) {
            if (!s.d_pos) CK(cudaMalloc(&s.d_pos, sizeof(oat_position)));
            if (!s.h_pos) CK(cudaHostAlloc(&s.h_pos, sizeof(oat_position), cudaHostAllocDefault));
            if (!s.pos_done) CK(cudaEventCreateWithFlags(&s.pos_done, cudaEventDisableTiming));
        }
    }
    return OAT_OK;
}

static int tracker_collect(oat_tracker *t, oat_detection *out, oat_position *pos)
{
    if (t->tailpos == t->head) return fail(OAT_ERR_STATE, "oat_tracker_collect: nothing outstanding");
    CKRET(bind(t->ctx));
    Slot &s = t->ring[t->tailpos % t->ring.size()];
    CK(cudaEventSynchronize(s.done));
    if (s.fast && s.h_res->status != TAIL_OK) {
        // the mask overflowed the one-launch tail's run table: replay this frame's mask through
        // the unbounded path (its bits are still in the slot)
        oat_ctx *c = t->ctx;
        CKRET(t->tail.run(c, s.bits, s.hp, &s.d_res->det, nullptr, 0, nullptr));
        CK(cudaMemcpyAsync(&s.h_res->det, &s.d_res->det, sizeof(oat_detection), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        ++t->replays;
    }
    if (s.fast) {  // the census rides on the one-launch tail's result
        const double groups = (double)t->m.g.rows * t->m.g.cols / 4.0;
        t->slow_frac = 0.75 * t->slow_frac + 0.25 * ((double)s.h_res->slow_groups / groups);
        if (!t->use_generic && t->slow_frac > 0.30) t->use_generic = true;
        else if (t->use_generic && t->slow_frac < 0.15) t->use_generic = false;
    }
    *out = s.h_res->det;
    if (s.has_pos) {
        oat_ctx *c = t->ctx;
        CK(cudaEventSynchronize(s.pos_done));
        if (s.h_pos->reserved == -1) {
            // the filter was held at this frame (its detection was not final when the epilogue ran): the
            // detection is final now -- run this frame, then re-run the frames submitted after it, in order
            CKRET(posfilt_launch(t->pf, nullptr, &s.d_res->det, nullptr, 1, s.d_pos));
            CK(cudaMemcpyAsync(s.h_pos, s.d_pos, sizeof(oat_position), cudaMemcpyDeviceToHost, c->post));
            for (uint64_t u = t->tailpos + 1; u < t->head; ++u) {
                Slot &n = t->ring[u % t->ring.size()];
                if (!n.has_pos) continue;
                CK(cudaStreamWaitEvent(c->post, n.done, 0));
                CKRET(posfilt_launch(t->pf, nullptr, &n.d_res->det, n.fast ? &n.d_res->status : nullptr, 0, n.d_pos));
                CK(cudaMemcpyAsync(n.h_pos, n.d_pos, sizeof(oat_position), cudaMemcpyDeviceToHost, c->post));
                CK(cudaEventRecord(n.pos_done, c->post));
            }
            CK(cudaStreamSynchronize(c->post));
        }
        if (pos) *pos = *s.h_pos;
    } else if (pos) {
        return fail(OAT_ERR_STATE, "oat_tracker_collect_position: the frame was submitted before the filter was attached");
    }
    t->last_slot = (size_t)(t->tailpos % t->ring.size());
    ++t->tailpos;
    return OAT_OK;
}

extern "C" int oat_tracker_submit(oat_tracker *t, const uint8_t *bgr_in, size_t in_pitch, double learning_rate,
                                  const oat_hsv_params *p, uint8_t *bgr_out, size_t bgr_out_pitch)
{
    CKRET(tracker_check(t, bgr_in, in_pitch, p));
    CKRET(bind(t->ctx));
    REQUIRE_NOT_STREAMING(t, "oat_tracker_submit");
    if (t->head - t->tailpos >= t->ring.size())
        return fail(OAT_ERR_STATE, "oat_tracker_submit: ring full (collect first)");
    const int rows = t->m.g.rows, cols = t->m.g.cols;
    REQUIRE(!bgr_out || bgr_out_pitch >= (size_t)3 * cols, "tracker: output pitch too small");
    // async egress needs a stable device view: device pointers are written directly, host
    // pointers go through a per-tracker staging buffer (one frame in flight for egress).
    REQUIRE(!bgr_out || mem_kind(bgr_out) == MEM_DEVICE || t->head == t->tailpos,
            "oat_tracker_submit: host bgr_out needs the previous frame collected first");
    OutView ob, none1, none2, none3;
    CKRET(stage_out(t->out_bgr, bgr_out, bgr_out_pitch, rows, (size_t)3 * cols, &ob));
    return tracker_enqueue(t, bgr_in, in_pitch, learning_rate, p, ob, none1, none2, none3, true);
}

// ---- resident clip engine ----------------------------------------------------------------------
// A clip of device-resident frames does not need the host per frame: the frames of a chunk (half of every
// tracker's ring) become a queue of FrameDesc / TailFrame descriptors, ONE launch of the resident fused kernel
// (compute stream) and ONE launch of the resident tail server (a tail stream) work through it, and the two
// kernels hand frames over on the device (Slot::d_done).  Two chunks are kept in flight: the fused kernel of
// chunk q+1 follows chunk q's tile by tile (FD_CHAIN), so neither a frame boundary nor a chunk boundary drains
// the machine.  Frames of several trackers (independent streams on one GPU) can be interleaved in one queue.

// Can this tracker's next frames go through the resident engine?  (Steady state only: the first frame of a
// model, a learning rate that changes per frame or re-initialises, a stream the census moved to the generic
// kernel, and geometries the one-launch tail cannot take stay on the per-frame path.)
static bool clip_eligible(const oat_tracker *t, double learning_rate, const oat_hsv_params *p, bool need_tail)
{
    const oat_ctx *c = t->ctx;
    if (c->no_pipe || c->no_clip || t->prof || t->use_generic) return false;
    if (t->m.K != 5 || t->m.nframes < 1 || !(learning_rate >= 0.0 && learning_rate < 1.0)) return false;
    if (t->m.g.cols % 4 != 0 || t->ring.size() < 2) return false;
    if (need_tail) {
        if (c->no_fast_tail || t->tail.fast_smem == 0 || t->m.g.rows < 2) return false;
        const int ke = p->erode_px > 0 ? p->erode_px : 0, kd = p->dilate_px > 0 ? p->dilate_px : 0;
        const size_t nin = (size_t)8 + (ke > 0 ? ke - 1 : 0) + (kd > 0 ? kd - 1 : 0);
        if (2 * nin * (size_t)t->m.g.wpr * sizeof(uint32_t) > 200 * 1024) return false;
    }
    return true;
}
// known_device: the caller has already established that the frame is device memory (cudaPointerGetAttributes is ~1 us)
static bool clip_frame_ok(const oat_tracker *t, const uint8_t *frame, size_t in_pitch, int *cls_out, bool known_device = false)
{
    if (!frame || (!known_device && mem_kind(frame) != MEM_DEVICE) || !aligned4(frame, in_pitch)) return false;
    FusedArgs a{};
    a.bgr = frame;
    a.in_pitch = in_pitch;
    const int cls = pipe_class(t->m, a);
    if (cls_out) *cls_out = cls;
    return cls > 0;
}

// Host side of the resident engine for a fixed set of trackers: geometry of a chunk and the two chunks in flight
// (ctx->clip[0..1]).  clip_run() drives it over a whole clip; the streaming entry points (oat_tracker_stream_*) keep
// one alive between calls.
struct ClipEngine {
    oat_ctx *c = nullptr;
    oat_tracker *trk[64] = {};
    int S = 0;
    bool fused_only = false;
    size_t chunkF = 0;  // frames (per tracker) per chunk: half of the smallest ring
    BitGeom g{};
    int ntiles = 0;
    size_t scratch_bytes = 0;
    static const int scratch_comps = 4096;
    struct Flight {
        size_t count = 0;
        bool live = false;
        uint64_t order = 0;
    } fl[2];
    uint64_t nchunk = 0;  // chunks launched; the next one uses half nchunk & 1
    int class_all = 2;
    double wait_ns = 0.0;
    bool frames_known_device = false;  // the stream checks every frame when it is pushed
    bool heavy_tail = false;           // the masks of these streams are busy: the tail server gets twice its usual share
    TailQueue tail_queue;              // kernel-parameter image of a short queue of tail descriptors

    void init(oat_ctx *ctx, oat_tracker *const *trackers, int n_trackers, bool fused)
    {
        c = ctx;
        S = n_trackers;
        fused_only = fused;
        chunkF = trackers[0]->ring.size() / 2;
        for (int s = 0; s < S; ++s) {
            trk[s] = trackers[s];
            chunkF = std::min(chunkF, trackers[s]->ring.size() / 2);
        }
        g = trk[0]->tail.tb.g;
        ntiles = (int)((trk[0]->m.plane + PIPE_TILE - 1) / PIPE_TILE);
        // second labelling attempt in global memory for masks whose tables do not fit the CTA's shared memory: room for
        // ~200 k runs and 4096 contours per CTA, plus the mask itself
        scratch_bytes = (((size_t)g.rows * g.wpr * 4 + (size_t)g.rows * 16 + (size_t)3 * scratch_comps * 8 + ((size_t)200 << 10) * 12 + 4096) + 255) & ~(size_t)255;
    }
    int next_half() const { return (int)(nchunk & 1); }
    bool half_free() const { return !fl[nchunk & 1].live; }
    int oldest_live() const
    {
        if (fl[0].live && fl[1].live) return fl[0].order < fl[1].order ? 0 : 1;
        return fl[0].live ? 0 : fl[1].live ? 1 : -1;
    }
    size_t frames_in_flight() const { return (fl[0].live ? fl[0].count : 0) + (fl[1].live ? fl[1].count : 0); }
    bool finished(int h) const { return fl[h].live && cudaEventQuery(c->clip[h].done) == cudaSuccess; }

    // One chunk: up to chunkF frames of `frames` ([n][S] frame-major) into the free half next_half().  *taken = frames
    // (per tracker) that went in -- fewer than min(n, chunkF) when a frame or a tracker is not eligible for the engine
    // (the caller takes it from there on the per-frame path).
    int launch(const uint8_t *const *frames, size_t n, size_t in_pitch, double learning_rate, const oat_hsv_params *p, size_t *taken)
    {
        *taken = 0;
        const int h = next_half();
        if (fl[h].live) return fail(OAT_ERR_STATE, "resident engine: both chunks are in flight");
        ClipHalf *half = c->clip;
        oat_tracker *t0 = trk[0];
        for (int s = 0; s < S; ++s)
            if (!clip_eligible(trk[s], learning_rate, p, !fused_only)) return OAT_OK;
        const int ke = p->erode_px > 0 ? p->erode_px : 0, kd = p->dilate_px > 0 ? p->dilate_px : 0;
        // bands of 32 rows (a band costs ~3 dependent L2 round trips whatever its size; most bands of a tracking mask are
        // empty) unless the morphology staging of such a band would not fit
        auto stage_for = [&](int rr) {
            const size_t nin = (size_t)rr + (ke > 0 ? ke - 1 : 0) + (kd > 0 ? kd - 1 : 0);
            return 2 * nin * (size_t)g.wpr * sizeof(uint32_t);
        };
        const int R = stage_for(32) <= (size_t)64 * 1024 ? 32 : 8;
        const size_t tail_smem = std::max(std::max(stage_for(R), t0->tail.fast_smem), (size_t)PIPE_TAIL_SMEM_KB_CFG * 1024);
        static size_t tail_stream_smem_set = 48 * 1024;
        if (!fused_only && tail_smem + 2048 > tail_stream_smem_set) {
            CK(cudaFuncSetAttribute(tail_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
            tail_stream_smem_set = 200 * 1024;
        }
        // how many frames can go into this chunk
        size_t cnt = 0;
        for (; cnt < chunkF && cnt < n; ++cnt) {
            bool ok = true;
            for (int s = 0; s < S && ok; ++s) {
                int cls = 0;
                ok = clip_frame_ok(trk[s], frames[cnt * S + s], in_pitch, &cls, frames_known_device);
                if (ok) class_all = std::min(class_all, cls);
            }
            if (!ok) break;  // this chunk ends before the frame
        }
        if (cnt == 0) return OAT_OK;
        CKRET(half[h].ensure(chunkF * (size_t)S));
        // The tail server's share of the machine follows the masks: a tracking stream (a blob or a few: ~100 run-table
        // entries per frame) keeps 24 of the 296 CTA slots busy at most, and so does a scene of dozens of blobs now that
        // the bands pre-label their rows (60 blobs per 1080p frame, 6.5 k entries: 36.7 k frames/s on 24 slots, 31.0 k on
        // 48 -- the fused kernel misses the slots more than the tail needs them).  From ~10 k entries per frame (200 blobs:
        // 14 k) the tail, not the HBM-bound fused kernel, sets the frame rate and gets twice the slots (27.3 k against
        // 25.8 k frames/s).
        for (int s = 0; s < S; ++s) {
            if (trk[s]->tail_load > c->heavy_tail_at) heavy_tail = true;
        }
        if (heavy_tail) {
            bool all_light = true;
            for (int s = 0; s < S; ++s) all_light = all_light && trk[s]->tail_load < 0.5 * c->heavy_tail_at;
            if (all_light) heavy_tail = false;
        }
        const int reserved = fused_only ? 0 : (heavy_tail ? 2 : 1) * (int)PIPE_RESERVED_CTAS_CFG;
        const bool full = (long long)ntiles * (long long)(cnt * S) >= (long long)pipe_grid_full(c, reserved);
        bool chain_launch = c->pdl && !c->no_chain && c->chain_uid != 0 && full;
        for (int s = 0; s < S; ++s) chain_launch = chain_launch && trk[s]->m.flags_current;
        // the band pre-labelling pools of the slots this chunk uses (first use of a slot only; before any bookkeeping moves)
        // (a tracking mask -- a blob or a few, ~100 runs -- is labelled from a staged copy in a few us: the bands'
        // pre-labelling would only lengthen every band by as much)
        if (!fused_only && !c->no_prelabel && R == 32)
            for (int s = 0; s < S; ++s)
                if ((trk[s]->tail_load > 300.0 || c->force_prelabel) && !trk[s]->pool_block) {
                    // one allocation for the pools of every ring slot of the stream, the first time its masks are busy
                    const size_t per = (FastBufs::pool_bytes(g.rows, div_up(g.rows, R)) + 255) & ~(size_t)255;
                    CK(cudaMalloc(&trk[s]->pool_block, per * trk[s]->ring.size()));
                    CK(cudaMemset(trk[s]->pool_block, 0, per * trk[s]->ring.size()));
                    for (size_t k = 0; k < trk[s]->ring.size(); ++k)
                        trk[s]->ring[k].fb.set_pool((uint8_t *)trk[s]->pool_block + k * per, g.rows, div_up(g.rows, R));
                }
        FusedArgs a{};
        bool seen[64] = {};  // (n_trackers <= 64)
        for (size_t i = 0; i < cnt; ++i)
            for (int s = 0; s < S; ++s) {
                oat_tracker *t = trk[s];
                Slot &sl = t->ring[(size_t)h * chunkF + i];
                MogConsts mc;
                int reset = 0;
                t->m.frame_consts(learning_rate, &mc, &reset);
                if (i == 0 && s == 0) a.c = mc;
                FrameDesc &d = half[h].h_desc[i * S + s];
                d.bgr = frames[i * S + s];
                d.state = t->m.state;
                d.nmodes = t->m.nmodes;
                d.thr_bits = sl.bits;
                d.tile_seq = t->m.tile_seq;
                d.slow_count = sl.d_slow;
                d.done_count = fused_only ? nullptr : sl.d_done;
                d.seq_expect = t->m.seq++;
                d.flags = (seen[s] || chain_launch) ? FD_CHAIN : 0u;  // later frames of a model follow its earlier ones in this launch
                seen[s] = true;
                sl.hp = *p;
                sl.fast = true;
                sl.has_pos = false;
                if (!fused_only) {
                    sl.done_total += (unsigned)ntiles;
                    TailFrame &tf = half[h].h_tf[i * S + s];
                    FastArgs &fa = tf.a;
                    fa.in = sl.bits;
                    fa.out = sl.fb.di;
                    fa.ke = ke;
                    fa.kd = kd;
                    fa.R = R;
                    fa.g = g;
                    fa.rowext = sl.fb.rowext;
                    fa.rowcnt = sl.fb.rowcnt;
                    fa.bbox = sl.fb.bbox;
                    fa.ticket = sl.fb.ticket;
                    fa.min_area = p->min_area;
                    fa.max_area = p->max_area;
                    fa.res = sl.d_res;
                    fa.res_host = sl.h_res;
                    fa.smem_bytes = (int)tail_smem;
                    fa.max_comps = t->tail.fast_comps;
                    fa.slow_in = sl.d_slow;
                    fa.in_place_ok = 1;  // (the engine has no thresh egress: nobody reads the slot's mask after the labelling)
                    fa.pool_runs = nullptr;  // (h_tf is reused from chunk to chunk: every field is set every time)
                    fa.pool_sums = nullptr;
                    fa.pool_agg = nullptr;
                    fa.band_hdr = nullptr;
                    fa.pool_alloc = nullptr;
                    fa.pool_cap = 0;
                    fa.force_pre = 0;
                    if (sl.fb.pool_runs && !c->no_prelabel && R == 32 && (t->tail_load > 300.0 || c->force_prelabel)) {
                        fa.pool_runs = sl.fb.pool_runs;
                        fa.pool_sums = sl.fb.pool_sums;
                        fa.pool_agg = sl.fb.pool_agg;
                        fa.band_hdr = sl.fb.band_hdr;
                        fa.pool_alloc = sl.fb.pool_alloc;
                        fa.pool_cap = sl.fb.pool_cap;
                        fa.force_pre = c->force_prelabel ? 1 : 0;
                    }
                    tf.done_count = sl.d_done;
                    tf.done_target = sl.done_total;
                    tf.pad = 0;
                }
            }
        const size_t nitems = cnt * (size_t)S;
        // descriptors: a short queue travels in the kernel parameters; a long one is uploaded on a side stream and
        // waited for by the HOST, so that nothing but the previous fused kernel precedes this launch on the compute
        // stream (the launches overlap tile by tile)
        const bool inline_descs = nitems <= (size_t)PIPE_INLINE_DESCS;
        if (!inline_descs) {
            CK(cudaMemcpyAsync(half[h].d_desc, half[h].h_desc, nitems * sizeof(FrameDesc), cudaMemcpyHostToDevice, c->aux));
            CK(cudaStreamSynchronize(c->aux));
        }
        a.in_pitch = in_pitch;
        a.rows = t0->m.g.rows;
        a.cols = t0->m.g.cols;
        a.wpr = t0->m.g.wpr;
        a.plane = t0->m.plane;
        a.hsv_lut = c->hsv_lut;
        a.do_hsv = 1;
        a.lo[0] = p->h_min;
        a.lo[1] = p->s_min;
        a.lo[2] = p->v_min;
        a.hi[0] = p->h_max;
        a.hi[1] = p->s_max;
        a.hi[2] = p->v_max;
        a.thr_bits = half[h].h_desc[0].thr_bits;  // (non-NULL: the kernel reads the per-frame pointer)
        a.bgr = half[h].h_desc[0].bgr;
        StreamArgs pa;
        stream_args_common(c, t0->m, a, pa);
        pa.nframes = (int)nitems;
        pa.descs = inline_descs ? nullptr : half[h].d_desc;
        if (inline_descs) memcpy(pa.inl, half[h].h_desc, nitems * sizeof(FrameDesc));
        pa.wait_grid = chain_launch ? 0 : 1;
        const bool frozen = (a.c.aT == 0.0f) && !c->no_track;
        CKRET(launch_stream(c, pa, frozen, class_all == 2, reserved));
        for (int s = 0; s < S; ++s) trk[s]->m.flags_current = true;
        c->chain_uid = full ? t0->m.uid : 0;
        if (!fused_only) {
            cudaStream_t ts = c->tail[c->tail_rr++ % oat_ctx::NTAIL];
            // a short queue of tail descriptors travels in the kernel parameters, like the fused kernel's (one driver call
            // less per chunk; the band counters need none either: the kernel leaves them re-armed)
            const bool inline_tf = nitems <= (size_t)TAIL_INLINE_DESCS;
            if (!inline_tf) CK(cudaMemcpyAsync(half[h].d_tf, half[h].h_tf, nitems * sizeof(TailFrame), cudaMemcpyHostToDevice, ts));
            const int nbands = div_up(g.rows, R);
            // (more CTAs than one frame has bands is useful: they work on different frames of the queue at the same time)
            const int gridT = std::max(1, std::min(nbands * (int)nitems, (heavy_tail ? 2 : 1) * (int)PIPE_TAIL_GRID_CFG));
            uint8_t *scratch = nullptr;
            if (scratch_bytes < ((size_t)1 << 31) && c->tail_scratch[h].ensure(scratch_bytes * gridT) == OAT_OK)
                scratch = (uint8_t *)c->tail_scratch[h].p;
            else
                cudaGetLastError();  // no scratch: overflowing masks are replayed by the host
            if (inline_tf) memcpy(tail_queue.inl, half[h].h_tf, nitems * sizeof(TailFrame));
            tail_stream_kernel<<<gridT, 256, tail_smem, ts>>>(inline_tf ? nullptr : half[h].d_tf, tail_queue, (int)nitems, half[h].d_ctr,
                                                              half[h].d_ctr + nitems, half[h].d_ctr + 2 * half[h].cap, scratch,
                                                              (int)scratch_bytes, scratch_comps);
            ++c->launches;
            CK(cudaGetLastError());
            CK(cudaEventRecord(half[h].done, ts));
        } else {
            CK(cudaEventRecord(half[h].done, c->stream));
        }
        fl[h].count = cnt;
        fl[h].live = true;
        fl[h].order = nchunk;
        ++nchunk;
        *taken = cnt;
        return OAT_OK;
    }

    // Waits for chunk h and hands out its detections ([count][S]; pos: [count], S == 1 with a position filter).
    int retire(int h, oat_detection *out, oat_position *pos, size_t *count)
    {
        Flight &f = fl[h];
        ClipHalf *half = c->clip;
        {
            const auto w0 = std::chrono::steady_clock::now();
            CK(cudaEventSynchronize(half[h].done));
            wait_ns += std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - w0).count();
        }
        if (!fused_only) {
            bool replayed = false;
            for (size_t i = 0; i < f.count; ++i)
                for (int s = 0; s < S; ++s) {
                    oat_tracker *t = trk[s];
                    Slot &sl = t->ring[(size_t)h * chunkF + i];
                    if (sl.h_res->status != TAIL_OK) {
                        // the mask outgrew even the tail server's global-memory tables: replay this frame's mask through
                        // the unbounded path (its bits are still in the slot)
                        CKRET(t->tail.run(c, sl.bits, sl.hp, &sl.d_res->det, nullptr, 0, nullptr));
                        CK(cudaMemcpyAsync(&sl.h_res->det, &sl.d_res->det, sizeof(oat_detection), cudaMemcpyDeviceToHost, c->stream));
                        ++t->replays;
                        replayed = true;
                    }
                }
            if (replayed) CK(cudaStreamSynchronize(c->stream));
            for (size_t i = 0; i < f.count; ++i)
                for (int s = 0; s < S; ++s) {
                    oat_tracker *t = trk[s];
                    Slot &sl = t->ring[(size_t)h * chunkF + i];
                    const double groups = (double)t->m.g.rows * t->m.g.cols / 4.0;
                    t->slow_frac = 0.75 * t->slow_frac + 0.25 * ((double)sl.h_res->slow_groups / groups);
                    if (!t->use_generic && t->slow_frac > 0.30) t->use_generic = true;
                    t->tail_load = 0.75 * t->tail_load + 0.25 * (double)sl.h_res->nodes;
                    if (out) out[i * S + s] = sl.h_res->det;
                    t->last_slot = (size_t)h * chunkF + i;
                }
            if (pos) {  // S == 1: the position epilogue over the chunk's final detections, in frame order
                oat_tracker *t = trk[0];
                for (size_t i = 0; i < f.count; ++i) {
                    Slot &sl = t->ring[(size_t)h * chunkF + i];
                    CKRET(posfilt_launch(t->pf, nullptr, &sl.d_res->det, nullptr, 1, sl.d_pos));
                    CK(cudaMemcpyAsync(sl.h_pos, sl.d_pos, sizeof(oat_position), cudaMemcpyDeviceToHost, c->post));
                }
                CK(cudaStreamSynchronize(c->post));
                for (size_t i = 0; i < f.count; ++i) pos[i] = *t->ring[(size_t)h * chunkF + i].h_pos;
            }
        }
        for (int s = 0; s < S; ++s) trk[s]->clip_frames += f.count;
        *count = f.count;
        f.live = false;
        return OAT_OK;
    }
};

// frames: [n][S] frame-major; out / pos likewise (pos only with S == 1).  *used = frames (per tracker) consumed;
// fewer than n if a frame or a tracker stopped being eligible -- the caller continues on the per-frame path.
static int clip_run(oat_ctx *c, oat_tracker *const *trk, int S, const uint8_t *const *frames, size_t n, size_t in_pitch,
                    double learning_rate, const oat_hsv_params *p, bool fused_only, oat_detection *out, oat_position *pos,
                    size_t *used)
{
    *used = 0;
    ClipEngine e;
    e.init(c, trk, S, fused_only);
    if (e.chunkF == 0 || n == 0) return OAT_OK;
    const auto t_enter = std::chrono::steady_clock::now();
    size_t next = 0, first[2] = {0, 0};
    bool stop = false;  // no new chunks: something stopped being eligible
    for (;;) {
        const int h = e.next_half();
        if (next < n && !stop && e.half_free()) {
            size_t taken = 0;
            CKRET(e.launch(frames + next * S, n - next, in_pitch, learning_rate, p, &taken));
            if (taken < std::min(e.chunkF, n - next)) stop = true;
            if (taken) {
                first[h] = next;
                next += taken;
                continue;
            }
        }
        const int r = e.oldest_live();
        if (r < 0) break;
        size_t cnt = 0;
        CKRET(e.retire(r, out ? out + first[r] * S : nullptr, pos ? pos + first[r] : nullptr, &cnt));
        *used = first[r] + cnt;
    }
    c->clip_wait_ns += e.wait_ns;
    c->clip_busy_ns += std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - t_enter).count() - e.wait_ns;
    c->clip_frames_total += *used * (uint64_t)S;
    return OAT_OK;
}

// the per-frame submit/collect loop, `depth` frames in flight (host-fed clips, first frames, generic-kernel streams)
static int run_clip_per_frame(oat_tracker *t, const uint8_t *const *frames, size_t n, size_t in_pitch, double learning_rate,
                              const oat_hsv_params *p, int depth, oat_detection *out, oat_position *pos)
{
    size_t d = depth < 1 ? 1 : (size_t)depth;
    if (d > t->ring.size()) d = t->ring.size();
    size_t done = 0;
    for (size_t i = 0; i < n; ++i) {
        if (i - done == d) {
            CKRET(tracker_collect(t, &out[done], pos ? &pos[done] : nullptr));
            ++done;
        }
        CKRET(oat_tracker_submit(t, frames[i], in_pitch, learning_rate, p, nullptr, 0));
    }
    for (; done < n; ++done) CKRET(tracker_collect(t, &out[done], pos ? &pos[done] : nullptr));
    return OAT_OK;
}

extern "C" int oat_tracker_wait_ingest(oat_tracker *t)
{
    REQUIRE(t, "null handle");
    REQUIRE(t->head != t->tailpos, "oat_tracker_wait_ingest: nothing outstanding");
    CKRET(bind(t->ctx));
    Slot &s = t->ring[(t->head - 1) % t->ring.size()];
    if (s.ingest == 1)
        CK(cudaEventSynchronize(s.copied));
    else
        CK(cudaEventSynchronize(s.fused_done));
    return OAT_OK;
}

extern "C" int oat_tracker_run_clip(oat_tracker *t, const uint8_t *const *frames, size_t n, size_t in_pitch,
                                    double learning_rate, const oat_hsv_params *p, int depth, oat_detection *out,
                                    oat_position *pos)
{
    REQUIRE(t && frames && out, "oat_tracker_run_clip: null argument");
    REQUIRE(t->head == t->tailpos, "oat_tracker_run_clip: frames are still outstanding (collect first)");
    REQUIRE_NOT_STREAMING(t, "oat_tracker_run_clip");
    REQUIRE(!t->ctx->clip_owner, "oat_tracker_run_clip: another tracker of this context is streaming");
    REQUIRE(!pos || t->pf, "oat_tracker_run_clip: positions requested but no position filter is attached");
    if (n == 0) return OAT_OK;
    CKRET(tracker_check(t, frames[0], in_pitch, p));
    CKRET(bind(t->ctx));
    size_t done = 0;
    while (done < n) {
        // device-resident frames in the steady state: the resident engine (one launch per chunk)
        if (clip_eligible(t, learning_rate, p, true) && clip_frame_ok(t, frames[done], in_pitch, nullptr)) {
            size_t used = 0;
            CKRET(clip_run(t->ctx, &t, 1, frames + done, n - done, in_pitch, learning_rate, p, false, out + done,
                           pos ? pos + done : nullptr, &used));
            done += used;
            if (used) continue;
        }
        // otherwise frame by frame: the first frame of a model on its own (the engine takes over from the second),
        // everything else (host frames, generic-kernel streams, ...) pipelined `depth` deep
        const bool first_only = t->m.nframes == 0 && learning_rate >= 0.0 && learning_rate < 1.0;
        const size_t cnt = first_only ? 1 : n - done;
        CKRET(run_clip_per_frame(t, frames + done, cnt, in_pitch, learning_rate, p, depth, out + done, pos ? pos + done : nullptr));
        done += cnt;
    }
    return OAT_OK;
}

// Frames of several independent streams on one GPU, interleaved in ONE queue of the resident engine:
// frames[i * n_trackers + s] is frame i of tracker s (all device-resident, same geometry, same parameters);
// out likewise.  flags & 1: fused kernel only (no detect tail, nothing returned; diagnostics / roofline).
extern "C" int oat_tracker_run_clips(oat_tracker *const *trackers, int n_trackers, const uint8_t *const *frames, size_t n_frames,
                                     size_t in_pitch, double learning_rate, const oat_hsv_params *p, int flags, oat_detection *out)
{
    REQUIRE(trackers && n_trackers >= 1 && n_trackers <= 64 && frames, "oat_tracker_run_clips: bad arguments");
    const bool fused_only = (flags & 1) != 0;
    REQUIRE(fused_only || out, "oat_tracker_run_clips: null output");
    oat_tracker *t0 = trackers[0];
    REQUIRE(t0, "oat_tracker_run_clips: null tracker");
    CKRET(check_hsv_params(p));
    for (int s = 0; s < n_trackers; ++s) {
        oat_tracker *t = trackers[s];
        REQUIRE(t && t->ctx == t0->ctx && t->m.g.rows == t0->m.g.rows && t->m.g.cols == t0->m.g.cols &&
                    memcmp(&t->m.p, &t0->m.p, sizeof(oat_mog_params)) == 0,
                "oat_tracker_run_clips: the trackers must share context, geometry and MOG parameters");
        REQUIRE(t->head == t->tailpos, "oat_tracker_run_clips: frames are still outstanding (collect first)");
        REQUIRE_NOT_STREAMING(t, "oat_tracker_run_clips");
        REQUIRE(!t->ctx->clip_owner, "oat_tracker_run_clips: another tracker of this context is streaming");
        REQUIRE(t->m.nframes == t0->m.nframes, "oat_tracker_run_clips: the trackers must have seen the same number of frames");
        for (int u = 0; u < s; ++u) REQUIRE(trackers[u] != t, "oat_tracker_run_clips: a tracker appears twice");
    }
    REQUIRE(in_pitch >= (size_t)3 * t0->m.g.cols, "tracker: input pitch too small");
    CKRET(bind(t0->ctx));
    size_t done = 0;
    while (done < n_frames) {
        bool ok = true;
        for (int s = 0; s < n_trackers; ++s)
            ok = ok && clip_eligible(trackers[s], learning_rate, p, !fused_only) &&
                 clip_frame_ok(trackers[s], frames[done * n_trackers + s], in_pitch, nullptr);
        if (ok) {
            size_t used = 0;
            CKRET(clip_run(t0->ctx, trackers, n_trackers, frames + done * n_trackers, n_frames - done, in_pitch, learning_rate, p,
                           fused_only, out ? out + done * n_trackers : nullptr, nullptr, &used));
            done += used;
            if (used) continue;
        }
        // one frame of every stream on the per-frame path
        for (int s = 0; s < n_trackers; ++s) {
            if (fused_only)
                CKRET(oat_tracker_submit_fused_only(trackers[s], frames[done * n_trackers + s], in_pitch, learning_rate, p));
            else
                CKRET(oat_tracker_submit(trackers[s], frames[done * n_trackers + s], in_pitch, learning_rate, p, nullptr, 0));
        }
        if (!fused_only)
            for (int s = 0; s < n_trackers; ++s) CKRET(tracker_collect(trackers[s], &out[done * n_trackers + s], nullptr));
        ++done;
    }
    return OAT_OK;
}

// ---- streaming use of the resident engine (oat_tracker_stream_*) -------------------------------------------
struct StreamState {
    ClipEngine eng;
    std::vector<const uint8_t *> gathered;  // device views of the frames of the chunk being gathered
    size_t pitch = 0;                       // ... their pitch, learning rate and detector parameters
    double lr = 0.0;
    oat_hsv_params p{};
    bool staged_any = false;                // a gathered frame sits in a staging buffer: its copy precedes the launch
    cudaEvent_t copied = nullptr;           // last staging copy on the ingest stream
    bool copy_pending = false;
    std::deque<std::pair<oat_detection, oat_position>> ready;
    std::vector<oat_detection> tmp_det;
    std::vector<oat_position> tmp_pos;
};
static bool stream_busy(const oat_tracker *t) { return t->stream && (!t->stream->gathered.empty() || t->stream->eng.frames_in_flight() > 0); }
static void stream_release_owner(oat_tracker *t)
{
    if (t->ctx->clip_owner == t && !stream_busy(t)) t->ctx->clip_owner = nullptr;
}
// the oldest chunk in flight -> ready
static int stream_retire_oldest(oat_tracker *t)
{
    StreamState *st = t->stream;
    const int h = st->eng.oldest_live();
    if (h < 0) return OAT_OK;
    st->tmp_det.resize(st->eng.chunkF);
    st->tmp_pos.resize(st->eng.chunkF);
    size_t cnt = 0;
    CKRET(st->eng.retire(h, st->tmp_det.data(), t->pf ? st->tmp_pos.data() : nullptr, &cnt));
    for (size_t i = 0; i < cnt; ++i) st->ready.emplace_back(st->tmp_det[i], t->pf ? st->tmp_pos[i] : oat_position{});
    t->ctx->clip_frames_total += cnt;
    return OAT_OK;
}
// one frame through the per-frame path, synchronously and in order (everything before it is retired first)
static int stream_fallback(oat_tracker *t, const uint8_t *frame, size_t pitch, double lr, const oat_hsv_params *p);
static int stream_launch_gathered(oat_tracker *t, bool block)
{
    StreamState *st = t->stream;
    if (st->gathered.empty()) return OAT_OK;
    if (!st->eng.half_free()) {
        if (!block) return OAT_OK;
        CKRET(stream_retire_oldest(t));
    }
    if (st->staged_any && st->copy_pending) {
        if (cudaEventQuery(st->copied) == cudaSuccess) {
            st->copy_pending = false;  // every staged frame has arrived: nothing but the previous fused kernel precedes this launch
        } else {
            // the last staged frame is still on its way (a host-fed stream: the launch is issued in the copy's shadow):
            // the compute stream waits for it, and the launch orders itself behind the whole previous grid, not tile by tile
            cudaGetLastError();
            CK(cudaStreamWaitEvent(t->ctx->stream, st->copied, 0));
            t->ctx->chain_uid = 0;
        }
    }
    size_t taken = 0;
    CKRET(st->eng.launch(st->gathered.data(), st->gathered.size(), st->pitch, st->lr, &st->p, &taken));
    std::vector<const uint8_t *> rest(st->gathered.begin() + taken, st->gathered.end());
    st->gathered.clear();
    st->staged_any = false;
    // (the tracker stopped being eligible between push and launch, e.g. the census moved it to the generic kernel)
    for (const uint8_t *f : rest) CKRET(stream_fallback(t, f, st->pitch, st->lr, &st->p));
    return OAT_OK;
}
static int stream_fallback(oat_tracker *t, const uint8_t *frame, size_t pitch, double lr, const oat_hsv_params *p)
{
    StreamState *st = t->stream;
    CKRET(stream_launch_gathered(t, true));
    while (st->eng.oldest_live() >= 0) CKRET(stream_retire_oldest(t));
    OutView n0, n1, n2, n3;
    CKRET(tracker_enqueue(t, frame, pitch, lr, p, n0, n1, n2, n3, true));
    oat_detection d;
    oat_position pos{};
    CKRET(tracker_collect(t, &d, t->pf ? &pos : nullptr));
    st->ready.emplace_back(d, pos);
    return OAT_OK;
}
static void stream_destroy(oat_tracker *t)
{
    StreamState *st = t->stream;
    if (!st) return;
    while (st->eng.oldest_live() >= 0)
        if (stream_retire_oldest(t) != OAT_OK) break;
    if (st->copied) {
        cudaEventSynchronize(st->copied);
        cudaEventDestroy(st->copied);
    }
    st->gathered.clear();
    if (t->ctx->clip_owner == t) t->ctx->clip_owner = nullptr;
    delete st;
    t->stream = nullptr;
}

extern "C" int oat_tracker_stream_push(oat_tracker *t, const uint8_t *bgr_in, size_t in_pitch, double learning_rate,
                                       const oat_hsv_params *p, unsigned flags)
{
    CKRET(tracker_check(t, bgr_in, in_pitch, p));
    REQUIRE(t->head == t->tailpos, "oat_tracker_stream_push: frames are still outstanding on the per-frame path (collect first)");
    REQUIRE(t->ring.size() >= 2, "oat_tracker_stream_push: the tracker needs a ring of at least 2");
    oat_ctx *c = t->ctx;
    REQUIRE(!c->clip_owner || c->clip_owner == t, "oat_tracker_stream_push: another tracker of this context is streaming");
    CKRET(bind(c));
    if (!t->stream) {
        t->stream = new (std::nothrow) StreamState();
        if (!t->stream) return fail(OAT_ERR_NOMEM, "out of host memory");
        t->stream->eng.init(c, &t, 1, false);
        t->stream->eng.frames_known_device = true;
        CK(cudaEventCreateWithFlags(&t->stream->copied, cudaEventDisableTiming));
    }
    StreamState *st = t->stream;
    c->clip_owner = t;
    const bool is_dev = mem_kind(bgr_in) == MEM_DEVICE;
    // staged: on request, host memory, or a device frame the engine cannot read where it is (rows not 16-byte aligned:
    // a tight 1000-column frame) -- a D2D copy into an aligned pitch costs microseconds, the per-frame path a launch pair
    const bool need_stage = (flags & OAT_STREAM_COPY) || !is_dev || !clip_frame_ok(t, bgr_in, in_pitch, nullptr, true);
    int r = OAT_OK;
    if (!clip_eligible(t, learning_rate, p, true)) {
        r = stream_fallback(t, bgr_in, in_pitch, learning_rate, p);  // (consumed when it returns)
        stream_release_owner(t);
        return r;
    }
    const int rows = t->m.g.rows;
    const size_t rowbytes = (size_t)3 * t->m.g.cols, tight = (rowbytes + 15) & ~(size_t)15;
    const size_t view_pitch = need_stage ? tight : in_pitch;
    if (!st->gathered.empty() && (st->pitch != view_pitch || st->lr != learning_rate || memcmp(&st->p, p, sizeof(*p)) != 0))
        CKRET(stream_launch_gathered(t, true));  // a chunk shares pitch, learning rate and detector parameters
    const uint8_t *view = bgr_in;
    if (need_stage) {
        // the staging buffer of the slot this frame will occupy: the half must have been retired
        if (!st->eng.half_free()) CKRET(stream_retire_oldest(t));
        Slot &sl = t->ring[(size_t)st->eng.next_half() * st->eng.chunkF + st->gathered.size()];
        CKRET(copy_frame_in(c->h2d, sl.in, &sl.in_raw, bgr_in, in_pitch, rows, rowbytes, tight, is_dev));
        CK(cudaEventRecord(st->copied, c->h2d));
        st->copy_pending = true;
        st->staged_any = true;
        view = (const uint8_t *)sl.in.p;
    }
    if (need_stage && !clip_frame_ok(t, view, view_pitch, nullptr, true)) {
        r = stream_fallback(t, view, view_pitch, learning_rate, p);
        stream_release_owner(t);
        return r;
    }
    if (st->gathered.empty()) {
        st->pitch = view_pitch;
        st->lr = learning_rate;
        st->p = *p;
    }
    st->gathered.push_back(view);
    if (st->gathered.size() >= st->eng.chunkF) CKRET(stream_launch_gathered(t, true));
    return OAT_OK;
}

extern "C" int oat_tracker_stream_wait_ingest(oat_tracker *t)
{
    REQUIRE(t, "null handle");
    if (!t->stream || !t->stream->copy_pending) return OAT_OK;
    CKRET(bind(t->ctx));
    CK(cudaEventSynchronize(t->stream->copied));
    t->stream->copy_pending = false;
    return OAT_OK;
}

extern "C" int oat_tracker_stream_flush(oat_tracker *t, int block)
{
    REQUIRE(t, "null handle");
    if (!t->stream) return OAT_OK;
    CKRET(bind(t->ctx));
    return stream_launch_gathered(t, block != 0);
}

extern "C" int oat_tracker_stream_poll(oat_tracker *t, oat_detection *out, oat_position *pos, size_t cap, int block, size_t *got)
{
    REQUIRE(t && got && (out || cap == 0), "oat_tracker_stream_poll: null argument");
    REQUIRE(!pos || t->pf, "oat_tracker_stream_poll: positions requested but no position filter is attached");
    *got = 0;
    StreamState *st = t->stream;
    if (!st) return OAT_OK;
    CKRET(bind(t->ctx));
    for (int h = st->eng.oldest_live(); h >= 0 && st->eng.finished(h); h = st->eng.oldest_live()) CKRET(stream_retire_oldest(t));
    if (block && st->ready.empty()) {
        if (st->eng.oldest_live() < 0) CKRET(stream_launch_gathered(t, true));
        CKRET(stream_retire_oldest(t));
    }
    while (*got < cap && !st->ready.empty()) {
        out[*got] = st->ready.front().first;
        if (pos) pos[*got] = st->ready.front().second;
        st->ready.pop_front();
        ++*got;
    }
    stream_release_owner(t);
    return OAT_OK;
}

extern "C" int oat_tracker_stream_pending(oat_tracker *t, size_t *gathered, size_t *in_flight, size_t *ready)
{
    REQUIRE(t, "null handle");
    const StreamState *st = t->stream;
    if (gathered) *gathered = st ? st->gathered.size() : 0;
    if (in_flight) *in_flight = st ? st->eng.frames_in_flight() : 0;
    if (ready) *ready = st ? st->ready.size() : 0;
    return OAT_OK;
}

extern "C" int oat_tracker_track(oat_tracker *t, const uint8_t *bgr_in, size_t in_pitch, double learning_rate,
                                 const oat_hsv_params *p, oat_detection *out, uint8_t *bgr_out, size_t bgr_out_pitch,
                                 uint8_t *fgmask_out, size_t fgmask_pitch, uint8_t *hsv_out, size_t hsv_pitch,
                                 uint8_t *thresh_out, size_t thresh_pitch)
{
    CKRET(tracker_check(t, bgr_in, in_pitch, p));
    REQUIRE(out, "oat_tracker_track: null output");
    REQUIRE(t->head == t->tailpos, "oat_tracker_track: frames are still outstanding (collect first)");
    REQUIRE_NOT_STREAMING(t, "oat_tracker_track");
    CKRET(bind(t->ctx));
    const int rows = t->m.g.rows, cols = t->m.g.cols;
    REQUIRE(!bgr_out || bgr_out_pitch >= (size_t)3 * cols, "tracker: bgr_out pitch too small");
    REQUIRE(!hsv_out || hsv_pitch >= (size_t)3 * cols, "tracker: hsv_out pitch too small");
    REQUIRE(!fgmask_out || fgmask_pitch >= (size_t)cols, "tracker: fgmask pitch too small");
    REQUIRE(!thresh_out || thresh_pitch >= (size_t)cols, "tracker: thresh pitch too small");
    OutView ob, ofg, ohsv, othr;
    CKRET(stage_out(t->out_bgr, bgr_out, bgr_out_pitch, rows, (size_t)3 * cols, &ob));
    CKRET(stage_out(t->out_fg, fgmask_out, fgmask_pitch, rows, (size_t)cols, &ofg));
    CKRET(stage_out(t->out_hsv, hsv_out, hsv_pitch, rows, (size_t)3 * cols, &ohsv));
    CKRET(stage_out(t->out_thr, thresh_out, thresh_pitch, rows, (size_t)cols, &othr));
    CKRET(tracker_enqueue(t, bgr_in, in_pitch, learning_rate, p, ob, ofg, ohsv, othr, false));
    return oat_tracker_collect(t, out);
}

extern "C" int oat_tracker_live_modes(oat_tracker *t, uint64_t *sum)
{
    REQUIRE(t, "null handle");
    CKRET(bind(t->ctx));
    return live_modes(t->ctx, t->m, sum);
}

extern "C" int oat_tracker_get_state(oat_tracker *t, uint8_t *modes_used, float *weight, float *variance,
                                     float *mean)
{
    REQUIRE(t, "null handle");
    CKRET(bind(t->ctx));
    return get_state(t->ctx, t->m, modes_used, weight, variance, mean);
}

// Diagnostic: what the one-launch tail needed for the most recently collected frame.
extern "C" int oat_tracker_tail_stats(oat_tracker *t, uint32_t *out /* [15]: status, nodes, replays, fast, cyc[8], generic frames, slow groups, clip frames */)
{
    REQUIRE(t && out, "null argument");
    const Slot &s = t->ring[t->last_slot];
    out[0] = (uint32_t)s.h_res->status;
    out[1] = s.h_res->nodes;
    out[2] = (uint32_t)t->replays;
    out[3] = (s.fast ? 1u : 0u) | (s.h_res->pad ? 2u : 0u);  // bit 1: the run table came pre-labelled from the bands
    for (int i = 0; i < 8; ++i) out[4 + i] = s.h_res->cyc[i];
    out[14] = (uint32_t)t->clip_frames;      // frames served by the resident clip engine (one launch per chunk)
    out[12] = (uint32_t)t->generic_frames;  // frames that ran the generic fused kernel (adaptive choice)
    out[13] = s.h_res->slow_groups;          // census of that frame
    return OAT_OK;
}

extern "C" int oat_tracker_profile(oat_tracker *t, int enable)
{
    REQUIRE(t, "null handle");
    t->prof = enable ? 1 : 0;
    return OAT_OK;
}
extern "C" int oat_tracker_profile_read(oat_tracker *t, double *mean_ms, uint64_t *launches)
{
    REQUIRE(t, "null handle");
    CKRET(bind(t->ctx));
    CK(cudaStreamSynchronize(t->ctx->stream));
    for (auto &pr : t->prof_pending) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, pr.first, pr.second));
        t->prof_ms += ms;
        ++t->prof_n;
        cudaEventDestroy(pr.first);
        cudaEventDestroy(pr.second);
    }
    t->prof_pending.clear();
    if (mean_ms) *mean_ms = t->prof_n ? t->prof_ms / (double)t->prof_n : 0.0;
    if (launches) *launches = t->prof_n;
    t->prof_ms = 0.0;
    t->prof_n = 0;
    return OAT_OK;
}

// ---- synthetic source ----------------------------------------------------------------------
extern "C" int oat_synth_frame(oat_ctx *c, uint8_t *dst, size_t pitch, int rows, int cols, uint32_t seed, uint32_t t)
{
    CKRET(bind(c));
    REQUIRE(dst && rows >= 20 && cols >= 4 && rows / 3 > 0 && cols / 2 > 0, "oat_synth_frame: bad arguments");
    REQUIRE(pitch >= (size_t)3 * cols, "oat_synth_frame: pitch too small");
    OutView ov;
    CKRET(stage_out(c->scratch_out, dst, pitch, rows, (size_t)3 * cols, &ov));
    const uint32_t kbg = fmix32_hd(seed ^ 0x9e3779b9u);
    const uint32_t knz = fmix32_hd(seed + 0x7f4a7c15u * (t + 1u));
    const int r = rows / 20;
    const int cx = cols / 4 + (int)((7u * t) % (uint32_t)(cols / 2));
    const int cy = rows / 3 + (int)((4u * t) % (uint32_t)(rows / 3));
    synth_kernel<<<nblocks((long long)rows * cols, 256), 256, 0, c->stream>>>(ov.d, ov.dpitch, rows, cols, kbg, knz,
                                                                            cx, cy, r, t != 0 ? 1 : 0);
    LAUNCH_CHECK(c);
    CKRET(finish_out(c->stream, ov));
    CK(cudaStreamSynchronize(c->stream));
    return OAT_OK;
}

// ---- memory helpers ------------------------------------------------------------------------
extern "C" int oat_alloc_device(oat_ctx *c, size_t bytes, void **out)
{
    CKRET(bind(c));
    REQUIRE(out && bytes > 0, "oat_alloc_device: bad arguments");
    CK(cudaMalloc(out, bytes));
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        g_dev_ranges[(uintptr_t)*out] = bytes;
    }
    return OAT_OK;
}
extern "C" int oat_free_device(oat_ctx *c, void *p)
{
    CKRET(bind(c));
    if (p) {
        {
            std::lock_guard<std::mutex> lk(g_dev_mu);
            g_dev_ranges.erase((uintptr_t)p);
        }
        CK(cudaFree(p));
    }
    return OAT_OK;
}
extern "C" int oat_alloc_pinned(size_t bytes, void **out)
{
    REQUIRE(out && bytes > 0, "oat_alloc_pinned: bad arguments");
    CK(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return OAT_OK;
}
extern "C" int oat_free_pinned(void *p)
{
    if (p) CK(cudaFreeHost(p));
    return OAT_OK;
}
extern "C" int oat_register_host(void *p, size_t bytes)
{
    REQUIRE(p && bytes > 0, "oat_register_host: bad arguments");
    CK(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return OAT_OK;
}
extern "C" int oat_unregister_host(void *p)
{
    if (p) CK(cudaHostUnregister(p));
    return OAT_OK;
}
extern "C" int oat_ipc_export(oat_ctx *c, const void *dev_ptr, unsigned char handle[64])
{
    CKRET(bind(c));
    REQUIRE(dev_ptr && handle, "oat_ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    memcpy(handle, &h, sizeof(h));
    return OAT_OK;
}
extern "C" int oat_ipc_open(oat_ctx *c, const unsigned char handle[64], void **dev_ptr)
{
    CKRET(bind(c));
    REQUIRE(dev_ptr && handle, "oat_ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return OAT_OK;
}
extern "C" int oat_ipc_close(oat_ctx *c, void *dev_ptr)
{
    CKRET(bind(c));
    if (dev_ptr) CK(cudaIpcCloseMemHandle(dev_ptr));
    return OAT_OK;
}
extern "C" int oat_memcpy(oat_ctx *c, void *dst, const void *src, size_t bytes)
{
    CKRET(bind(c));
    REQUIRE(dst && src, "oat_memcpy: null pointer");
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return OAT_OK;
}
extern "C" int oat_memcpy_async(oat_ctx *c, int lane, void *dst, const void *src, size_t bytes)
{
    CKRET(bind(c));
    REQUIRE(dst && src, "oat_memcpy_async: null pointer");
    REQUIRE(lane == 0 || lane == 1, "oat_memcpy_async: lane must be 0 (ingest) or 1 (egress)");
    if (!c->lane[lane]) {
        CK(cudaStreamCreateWithFlags(&c->lane[lane], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c->lane_done[lane], cudaEventDisableTiming));
    }
    if (lane == 1) {  // egress: the copy reads what the compute stream produces
        if (!c->lane_after) CK(cudaEventCreateWithFlags(&c->lane_after, cudaEventDisableTiming));
        CK(cudaEventRecord(c->lane_after, c->stream));
        CK(cudaStreamWaitEvent(c->lane[lane], c->lane_after, 0));
    }
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, c->lane[lane]));
    CK(cudaEventRecord(c->lane_done[lane], c->lane[lane]));
    return OAT_OK;
}
extern "C" int oat_memcpy_wait(oat_ctx *c, int lane)
{
    CKRET(bind(c));
    REQUIRE(lane == 0 || lane == 1, "oat_memcpy_wait: lane must be 0 or 1");
    if (c->lane[lane]) CK(cudaEventSynchronize(c->lane_done[lane]));
    return OAT_OK;
}
extern "C" int oat_memcpy_done(oat_ctx *c, int lane, int *done)
{
    CKRET(bind(c));
    REQUIRE(done && (lane == 0 || lane == 1), "oat_memcpy_done: bad argument");
    *done = 1;
    if (c->lane[lane]) {
        const cudaError_t e = cudaEventQuery(c->lane_done[lane]);
        if (e == cudaErrorNotReady) *done = 0;
        else CK(e);
    }
    return OAT_OK;
}
extern "C" int oat_flush_l2(oat_ctx *c)
{
    CKRET(bind(c));
    const size_t bytes = (size_t)256 << 20;
    CKRET(c->flush.ensure(bytes));
    static uint32_t v = 0;
    flush_kernel<<<148 * 8, 256, 0, c->stream>>>((uint4 *)c->flush.p, bytes / 16, ++v);
    CK(cudaGetLastError());  // benchmark utility: not counted in oat_ctx_kernel_launches
    return OAT_OK;
}
