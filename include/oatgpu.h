/*
 * oatgpu.h -- C ABI of liboatgpu.so, the B200 (sm_100a) implementation of Oat's per-frame
 * tracking hot path:  framefilt mog -> framefilt col -C HSV -> posidet hsv  (+ framefilt bsub).
 *
 * This header is the drop-in boundary.  Every entry point names the reference interface it
 * replaces (paths relative to the Oat source tree).  A maintainer of the reference binds these
 * from the `filter()` / `detectPosition()` overrides of `oat-framefilt` / `oat-posidet`; the
 * stub to add is shown in INTEGRATION.md.
 *
 * Conventions
 *   - plain C: opaque handles, plain pointers, sizes and pitches in BYTES, no C++/torch types;
 *   - every call returns OAT_OK (0) or a negative oat_status; the message of the most recent
 *     failure on the calling thread is oat_last_error() (the C++ host layer rethrows it as
 *     std::runtime_error so that main() keeps the reference's "name: message, exit -1" contract,
 *     src/framefilter/main.cpp:278-295);
 *   - image pointers may be device memory, pinned host memory or pageable host memory; the
 *     library inspects them (cudaPointerGetAttributes) and stages host buffers itself;
 *   - images are 8-bit, 1 or 3 interleaved channels, row-major (lib/datatypes/Color.h:29-38);
 *   - handles are NOT thread-safe; use one handle per thread/stream of frames, exactly like
 *     the single-threaded reference components (lib/base/Component.cpp:50-76);
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with
 *     OAT_ERR_CUDA.
 */
#ifndef OATGPU_H
#define OATGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OATGPU_ABI_VERSION 1

/* Every entry point below is exported with default visibility; everything else in the
 * library (including the statically linked CUDA runtime) stays hidden. */
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef enum oat_status {
    OAT_OK = 0,
    OAT_ERR_INVALID = -1,  /* bad argument (null handle, bad geometry, bad parameter range) */
    OAT_ERR_CUDA = -2,     /* CUDA runtime error / no device */
    OAT_ERR_NOMEM = -3,    /* allocation failed */
    OAT_ERR_STATE = -4,    /* call sequence error (collect without submit, ring full, ...) */
    OAT_ERR_UNSUPPORTED = -5
} oat_status;

/* ---- library ------------------------------------------------------------------------- */

int oat_abi_version(void);
/* Message of the last failure on this thread ("" if none). Never NULL. */
const char *oat_last_error(void);
/* Number of CUDA devices visible (0 if none / no driver). Never fails. */
int oat_device_count(void);

/* ---- context: one per (thread, device) ------------------------------------------------
 * Replaces BackgroundSubtractorMOG::configureGPU(index)
 * (src/framefilter/BackgroundSubtractorMOG.cpp:92-111): validates the index, selects the
 * device, owns the CUDA streams the handles below run on. */
typedef struct oat_ctx oat_ctx;
int oat_ctx_create(int device_index, oat_ctx **out);
/* Objects created from a context keep it alive: if any still exist, the context's resources are released when the
 * last of them is destroyed (so the order in which a garbage collector or an error path destroys handles does not
 * matter); the handle itself must not be used after this call. */
int oat_ctx_destroy(oat_ctx *ctx);
/* Block until everything queued through this context has finished. */
int oat_ctx_sync(oat_ctx *ctx);
/* Polls instead: *idle = 1 when everything given to the context's compute stream so far has finished. */
int oat_ctx_idle(oat_ctx *ctx, int *idle);
/* The cudaStream_t (as void*) the synchronous entry points launch on; lets a harness time
 * with CUDA events on the launching stream. */
void *oat_ctx_stream(oat_ctx *ctx);
/* Number of kernels of THIS library launched through ctx since creation. */
uint64_t oat_ctx_kernel_launches(const oat_ctx *ctx);
/* Device timing of the resident fused kernel (the dominant, HBM-bound kernel): while enabled, every launch is
 * bracketed by a CUDA-event pair on the compute stream; read returns the summed duration, the number of launches
 * and the number of frames they processed since the last read (and waits for the stream). */
int oat_ctx_profile_resident(oat_ctx *ctx, int enable);
/* Host side of the resident clip engine since the last call: microseconds the calling thread spent WORKING inside
 * oat_tracker_run_clip(s) (descriptors, launches, reading results), microseconds it spent waiting for chunks to
 * complete, and the frames served.  (The reference's component loop is one thread, lib/base/Component.cpp:56-76:
 * this is what that thread has left.) */
int oat_ctx_clip_host_stats(oat_ctx *ctx, double *busy_us, double *wait_us, uint64_t *frames);
int oat_ctx_profile_resident_read(oat_ctx *ctx, double *total_ms, uint64_t *launches, uint64_t *frames);

/* ---- framefilt mog --------------------------------------------------------------------
 * Parameters of cv::BackgroundSubtractorMOG2; oat_mog_default_params() gives what
 * cv::createBackgroundSubtractorMOG2() with no arguments gives, which is what the reference
 * constructs (src/framefilter/BackgroundSubtractorMOG.cpp:83). */
typedef struct oat_mog_params {
    int history;              /* 500 */
    int nmixtures;            /* 5 (1..5 supported) */
    float var_threshold;      /* Tb 16 */
    float var_threshold_gen;  /* Tg 9 */
    float background_ratio;   /* TB 0.9 */
    float var_init;           /* 15 */
    float var_min;            /* 4 */
    float var_max;            /* 75 */
    float complexity_reduction_threshold; /* CT 0.05 */
    int detect_shadows;       /* 1 */
    int shadow_value;         /* 127 */
    float shadow_threshold;   /* tau 0.5 */
} oat_mog_params;
void oat_mog_default_params(oat_mog_params *p);

typedef struct oat_mog oat_mog;
/* Replaces BackgroundSubtractorMOG::applyConfiguration (…MOG.cpp:69-89). params==NULL means
 * defaults. The GMM state (fp32, mode-major planes) lives in HBM for the life of the handle. */
int oat_mog_create(oat_ctx *ctx, int rows, int cols, const oat_mog_params *params, oat_mog **out);
int oat_mog_destroy(oat_mog *mog);
/* Forget the model (the next apply is "frame 1" again). */
int oat_mog_reset(oat_mog *mog);
/* Replaces BackgroundSubtractorMOG::filter, CPU branch (…MOG.cpp:124-125):
 *     mog2->apply(frame, mask, learning_rate);  frame.setTo(0, mask == 0);
 * bgr_in  : rows x cols x 3 BGR.     bgr_out : same geometry, may alias bgr_in, may be NULL.
 * mask_out: rows x cols x 1 in {0 background, shadow_value, 255 foreground}, may be NULL.
 * learning_rate is Oat's --adaptation-coeff (…MOG.cpp:56-59, :87-88); negative = automatic,
 * >= 1 re-initialises the model, exactly as cv::BackgroundSubtractorMOG2::apply does.
 * Synchronous: outputs are valid on return. */
int oat_mog_apply(oat_mog *mog, const uint8_t *bgr_in, size_t in_pitch, uint8_t *bgr_out,
                  size_t out_pitch, uint8_t *mask_out, size_t mask_pitch, double learning_rate);
/* Stream variant for device-resident frames (SharedFrameHeader memory kind DEVICE on both sides): enqueues
 * the frame on the context's stream and returns; every image pointer must be device memory; consecutive
 * frames overlap on the device.  oat_ctx_sync() (or any synchronous entry point) waits for them. */
int oat_mog_apply_async(oat_mog *h, const uint8_t *bgr_in, size_t in_pitch, uint8_t *bgr_out,
                        size_t out_pitch, uint8_t *mask_out, size_t mask_pitch, double learning_rate);
/* Test/diagnostic egress of the GMM state in OpenCV's layout (host pointers, any may be NULL):
 * modes_used u8[rows*cols]; weight,variance f32[rows*cols*K]; mean f32[rows*cols*K*3]. */
int oat_mog_get_state(oat_mog *mog, uint8_t *modes_used, float *weight, float *variance,
                      float *mean);
/* Sum over pixels of live GMM modes after the last apply (m-bar = sum / (rows*cols)); this is
 * the figure the algorithmic-bytes model 8 + 40*m-bar needs. */
int oat_mog_live_modes(oat_mog *mog, uint64_t *sum_modes);

/* ---- framefilt col -C HSV -------------------------------------------------------------
 * Replaces ColorConvert::filter for BGR->HSV (src/framefilter/ColorConvert.cpp:101-107,
 * lib/datatypes/Color.h:46-51): 8-bit cv::COLOR_BGR2HSV, H in [0,180). */
int oat_bgr2hsv(oat_ctx *ctx, const uint8_t *bgr, size_t in_pitch, uint8_t *hsv, size_t out_pitch,
                int rows, int cols);

/* ---- framefilt bsub -------------------------------------------------------------------
 * Replaces BackgroundSubtractor::filter (src/framefilter/BackgroundSubtractor.cpp:87-100):
 * first frame becomes the background; alpha>0 runs accumulateWeighted + convertTo(8U); output
 * is the saturating per-channel difference frame - background. channels is 1 or 3. */
typedef struct oat_bsub oat_bsub;
int oat_bsub_create(oat_ctx *ctx, int rows, int cols, int channels, double alpha, oat_bsub **out);
int oat_bsub_destroy(oat_bsub *b);
/* Optional explicit background image (BackgroundSubtractor.cpp:80-85). */
int oat_bsub_set_background(oat_bsub *b, const uint8_t *img, size_t pitch);
int oat_bsub_apply(oat_bsub *b, const uint8_t *in, size_t in_pitch, uint8_t *out, size_t out_pitch);

/* ---- posidet hsv ----------------------------------------------------------------------
 * Parameters of HSVDetector (src/positiondetector/HSVDetector.cpp:54-69, :87-140;
 * HSVDetector.h:86-94). oat_hsv_default_params: H,S,V = [0,256], erode 0 (off), dilate 10,
 * area = [0, DBL_MAX). */
typedef struct oat_hsv_params {
    int h_min, h_max; /* 0..256, inclusive band, 256 == 255 */
    int s_min, s_max;
    int v_min, v_max;
    int erode_px;     /* <=0 : off */
    int dilate_px;    /* <=0 : off */
    double min_area;  /* keep contours with min_area <= m00 < max_area */
    double max_area;
} oat_hsv_params;
void oat_hsv_default_params(oat_hsv_params *p);

/* What detectPosition() + siftContours() produce (HSVDetector.cpp:142-173,
 * DetectorFunc.cpp:31-66): Position2D::position / position_valid and the object area. */
typedef struct oat_detection {
    int32_t position_valid;
    int32_t n_components; /* external contours found (diagnostic; not in the reference) */
    double x;             /* m10 / m00 of the selected contour */
    double y;             /* m01 / m00 */
    double area;          /* m00 (0 if none selected) */
} oat_detection;

typedef struct oat_hsvdet oat_hsvdet;
int oat_hsvdet_create(oat_ctx *ctx, int rows, int cols, oat_hsvdet **out);
int oat_hsvdet_destroy(oat_hsvdet *det);
/* Replaces HSVDetector::detectPosition (HSVDetector.cpp:142-173): inRange -> erode -> dilate ->
 * findContours(RETR_EXTERNAL) -> moments -> largest area in [min,max).
 * hsv: rows x cols x 3 HSV frame. Optional egress (each may be NULL; device or host):
 *   thresh_out : rows x cols u8 {0,255}, the mask after inRange+erode+dilate;
 *   labels_out : rows x cols int32, 8-connected component label of every foreground pixel of
 *                that mask = linear index (y*cols+x) of the component's raster-first pixel,
 *                -1 for background. */
int oat_hsvdet_detect(oat_hsvdet *det, const uint8_t *hsv, size_t pitch, const oat_hsv_params *p,
                      oat_detection *out, uint8_t *thresh_out, size_t thresh_pitch,
                      int32_t *labels_out);
/* Same back-end on an already-thresholded binary mask (non-zero = foreground): the
 * siftContours() entry (src/positiondetector/DetectorFunc.cpp:31-66) that `posidet thresh`
 * and `posidet diff` also use. Morphology is applied first if erode_px/dilate_px > 0. */
int oat_sift_contours(oat_hsvdet *det, const uint8_t *mask, size_t pitch, const oat_hsv_params *p,
                      oat_detection *out, uint8_t *thresh_out, size_t thresh_pitch,
                      int32_t *labels_out);

/* posidet thresh: replaces SimpleThreshold::detectPosition (src/positiondetector/SimpleThreshold.cpp:
 * 130-143, :169-182): cv::inRange(grey, t_min, t_max) -> erode -> dilate -> siftContours on a 1-channel
 * GREY frame; thresholds in 0..256 (256 == 255), erode/dilate/area from p. Egress as oat_hsvdet_detect. */
int oat_thresh_detect(oat_hsvdet *det, const uint8_t *grey, size_t pitch, int t_min, int t_max,
                      const oat_hsv_params *p, oat_detection *out, uint8_t *thresh_out,
                      size_t thresh_pitch, int32_t *labels_out);

/* posidet diff: replaces DifferenceDetector::detectPosition (src/positiondetector/DifferenceDetector.cpp:118-173):
 * cv::absdiff(frame, last) -> cv::threshold(> diff_threshold, THRESH_BINARY) -> cv::blur(blur_px x blur_px) ->
 * siftContours (a blurred pixel counts as foreground when it is non-zero), last <- frame. GREY frames. The first
 * call has no previous image and, like the reference, sifts the raw frame. blur_px 0 = off; 1..22 supported
 * (OAT_ERR_UNSUPPORTED above: one set pixel then no longer survives the box average). thresh_out (optional): 255
 * where the sifted image is non-zero. */
typedef struct oat_diffdet oat_diffdet;
int oat_diffdet_create(oat_ctx *ctx, int rows, int cols, oat_diffdet **out);
int oat_diffdet_destroy(oat_diffdet *det);
int oat_diffdet_reset(oat_diffdet *det);
int oat_diffdet_detect(oat_diffdet *det, const uint8_t *grey, size_t pitch, int diff_threshold, int blur_px,
                       double min_area, double max_area, oat_detection *out, uint8_t *thresh_out,
                       size_t thresh_pitch);

/* framefilt thresh / framefilt mask: out = in where kept, 0 elsewhere (may alias in).
 *   roi == NULL: replaces Threshold::filter (src/framefilter/Threshold.cpp:67-81): keep pixels whose
 *                grey value (the frame itself if 1 channel, 8-bit cv::COLOR_BGR2GRAY if 3) lies in
 *                [i_min, i_max], 0..256;
 *   roi != NULL: replaces FrameMasker::filter (src/framefilter/FrameMasker.cpp:71-75):
 *                frame.setTo(0, roi == 0), roi = rows x cols u8. */
int oat_keep_where(oat_ctx *ctx, const uint8_t *in, size_t in_pitch, uint8_t *out, size_t out_pitch,
                   int rows, int cols, int channels, const uint8_t *roi, size_t roi_pitch, int i_min,
                   int i_max);

/* ---- fused tracker: mog -> col HSV -> hsv in one pass over HBM ------------------------
 * One handle = one video stream (its own GMM state). Equivalent to the three reference
 * components chained with no shared-memory hops in between. */
typedef struct oat_tracker oat_tracker;
int oat_tracker_create(oat_ctx *ctx, int rows, int cols, const oat_mog_params *mog_params,
                       int ring_depth /* frames in flight for submit/collect, 0 = default */,
                       oat_tracker **out);
int oat_tracker_destroy(oat_tracker *t);
int oat_tracker_reset(oat_tracker *t);
/* Synchronous: one frame in, one detection out. Optional egress, each may be NULL:
 * bgr_out (filtered frame = what framefilt mog publishes), fgmask_out (MOG2 mask),
 * hsv_out (what framefilt col publishes), thresh_out (post-morphology mask). */
int oat_tracker_track(oat_tracker *t, const uint8_t *bgr_in, size_t in_pitch, double learning_rate,
                      const oat_hsv_params *p, oat_detection *out, uint8_t *bgr_out,
                      size_t bgr_out_pitch, uint8_t *fgmask_out, size_t fgmask_pitch,
                      uint8_t *hsv_out, size_t hsv_pitch, uint8_t *thresh_out,
                      size_t thresh_pitch);
/* Asynchronous pair: submit queues a frame (returns once the work is enqueued; host input
 * buffers must stay valid and unchanged until the matching collect), collect returns the
 * detections in submission order. bgr_out as in oat_tracker_track (may be NULL).
 * OAT_ERR_STATE if more than ring_depth frames are outstanding / nothing is outstanding. */
int oat_tracker_submit(oat_tracker *t, const uint8_t *bgr_in, size_t in_pitch, double learning_rate,
                       const oat_hsv_params *p, uint8_t *bgr_out, size_t bgr_out_pitch);
int oat_tracker_collect(oat_tracker *t, oat_detection *out);
/* Blocks until the most recently submitted frame's INPUT has been consumed (host frame: its H2D copy is
 * complete; device frame: the fused kernel that reads it in place has finished) -- the moment a component may
 * hand the frame back to its SOURCE (Source::post(), lib/shmemdf/Source.h:217-232) while the frame's detection is
 * still in flight.  What lets a component keep several frames in flight behind a lock-step SOURCE. */
int oat_tracker_wait_ingest(oat_tracker *t);
int oat_tracker_live_modes(oat_tracker *t, uint64_t *sum_modes);

/* ---- posifilt kalman + posicom mean: the O(1) epilogue behind the detectors ----
 * The Position2D fields those components touch (lib/datatypes/Position2D.h:112-155). */
typedef struct oat_position {
    int32_t position_valid, velocity_valid, heading_valid;
    int32_t reserved;
    double x, y;   /* Position2D::position */
    double vx, vy; /* Position2D::velocity */
    double hx, hy; /* Position2D::heading (unit vector) */
} oat_position;
/* KalmanFilter2D's options (src/positionfilter/KalmanFilter2D.cpp:40-92; defaults
 * KalmanFilter2D.h:66-83: dt 0.02 s, timeout 0 s, sigma-accel 5, sigma-noise 0).  NOTE the reference's
 * semantics: not_found_threshold = int(timeout / dt), and a position is valid only while
 * not_found_count < threshold -- with the default timeout 0 no output is ever valid. */
typedef struct oat_kalman_params {
    double dt, timeout, sigma_accel, sigma_noise;
} oat_kalman_params;
void oat_kalman_default_params(oat_kalman_params *p);
typedef struct oat_posfilt oat_posfilt;
/* n_sources position streams (1..8); kalman != NULL: each goes through its own KalmanFilter2D
 * (replaces KalmanFilter2D::filter, KalmanFilter2D.cpp:95-145); combine_mean != 0: the (filtered)
 * sources are merged by MeanPosition::combine (src/positioncombiner/MeanPosition.cpp:60-118) with
 * heading_anchor = --heading-anchor (negative: headings are averaged instead of generated).
 * The filter state lives on the device; one apply = one kernel launch. */
int oat_posfilt_create(oat_ctx *ctx, int n_sources, const oat_kalman_params *kalman, int combine_mean,
                       int heading_anchor, oat_posfilt **out);
int oat_posfilt_destroy(oat_posfilt *f);
int oat_posfilt_reset(oat_posfilt *f);
/* sources: n_sources raw positions (host memory); out: 1 position if combine_mean else n_sources. */
int oat_posfilt_apply(oat_posfilt *f, const oat_position *sources, oat_position *out);
/* Fuse a single-source filter behind a tracker: every frame's detection feeds it on the device, in
 * frame order (also when detect tails of consecutive frames overlap), and
 * oat_tracker_collect_position returns the filtered position next to the raw detection.
 * Pass NULL to detach. */
int oat_tracker_attach_posfilt(oat_tracker *t, oat_posfilt *f);
int oat_tracker_collect_position(oat_tracker *t, oat_detection *det, oat_position *pos);
/* A whole clip -- what a file-fed graph (oat frameserve file, src/frameserver/FileReader.cpp:103-131)
 * amounts to.  out: n detections in frame order; pos: NULL, or n filtered positions when a position
 * filter is attached.
 * Device-resident frames in the steady state (every frame after the model's first, a fixed learning rate in
 * [0,1), rows 16-byte aligned) go through the RESIDENT engine: the clip is cut into chunks of half the
 * tracker's ring, and per chunk ONE launch of the fused kernel and ONE launch of the tail server work
 * through a queue of frame descriptors on the device -- no host work per frame (`depth` is not used there;
 * create the tracker with a ring of 64 for 32-frame chunks).  Anything else (host frames, the first frame,
 * learning rate < 0 or >= 1) runs frames[0..n) through submit/collect with up to `depth` frames in flight.
 * Results are identical either way. */
int oat_tracker_run_clip(oat_tracker *t, const uint8_t *const *frames, size_t n, size_t in_pitch,
                         double learning_rate, const oat_hsv_params *p, int depth, oat_detection *out,
                         oat_position *pos);
/* Several independent streams on one GPU (BASELINE configs 3-4: --gpu-index shards streams over GPUs,
 * src/framefilter/BackgroundSubtractorMOG.cpp:92-111; this batches the streams of ONE GPU): frame i of
 * tracker s is frames[i * n_trackers + s], its detection out[i * n_trackers + s].  The trackers must share
 * context, geometry and MOG parameters and have seen the same number of frames.  Device-resident frames are
 * interleaved in one queue of the resident engine.  flags & OAT_CLIPS_FUSED_ONLY: run only the fused
 * MOG+HSV+threshold kernel (no detect tail, out may be NULL) -- diagnostics / roofline timing. */
#define OAT_CLIPS_FUSED_ONLY 1
int oat_tracker_run_clips(oat_tracker *const *trackers, int n_trackers, const uint8_t *const *frames,
                          size_t n_frames, size_t in_pitch, double learning_rate, const oat_hsv_params *p,
                          int flags, oat_detection *out);
/* STREAMING use of the resident engine -- what lets a component in a shared-memory graph (a lock-step
 * SOURCE in front, a lock-step SINK behind: lib/shmemdf/Source.h:187-232, Sink.h:93-138) run at the engine's
 * rate instead of one launch pair per frame; the GPU-aware counterpart of putting `oat buffer`
 * (src/buffer/FrameBuffer.cpp:56-116) in front of a slow component.
 *   push        hands ONE frame to the tracker and returns without waiting for its detection.  Frames are
 *               gathered into a chunk (half the tracker's ring); a full chunk is launched at once (one fused
 *               launch + one tail-server launch), a partial one by flush.  Two chunks are in flight; a push
 *               that needs a third waits for the oldest and keeps its detections for poll.
 *               flags & OAT_STREAM_COPY: the caller takes the frame's memory back after wait_ingest, so the frame
 *               is staged into the tracker's own HBM first (H2D from host memory -- always staged --, D2D from
 *               device memory).  Without it a device frame is read IN PLACE and must stay valid and unchanged
 *               until its detection has been polled.
 *               A frame the engine cannot take (the model's first frame, learning rate < 0 or >= 1, ...) goes
 *               through the per-frame path in order, synchronously.
 *   wait_ingest blocks until the pixels of every pushed OAT_STREAM_COPY frame have left the caller's memory.
 *   flush       launches the gathered frames as a (short) chunk if a chunk slot is free; block != 0: waits for one.
 *   poll        detections (and filtered positions, pos may be NULL) of finished frames, in push order; block != 0:
 *               waits until at least one is there (flushing if need be) unless nothing is outstanding.
 *   pending     frames gathered / in flight on the GPU / finished and waiting for poll.
 * Results are identical to submit/collect and run_clip.  One tracker per context streams at a time; the other
 * tracker entry points return OAT_ERR_STATE while frames are gathered or in flight. */
#define OAT_STREAM_COPY 1u
int oat_tracker_stream_push(oat_tracker *t, const uint8_t *bgr_in, size_t in_pitch, double learning_rate,
                            const oat_hsv_params *p, unsigned flags);
int oat_tracker_stream_wait_ingest(oat_tracker *t);
int oat_tracker_stream_flush(oat_tracker *t, int block);
int oat_tracker_stream_poll(oat_tracker *t, oat_detection *out, oat_position *pos, size_t cap, int block,
                            size_t *got);
int oat_tracker_stream_pending(oat_tracker *t, size_t *gathered, size_t *in_flight, size_t *ready);
/* GMM state egress, as oat_mog_get_state. */
int oat_tracker_get_state(oat_tracker *t, uint8_t *modes_used, float *weight, float *variance,
                          float *mean);
/* Per-kernel device timing (CUDA events on the launching stream around the fused
 * MOG+HSV+threshold kernel): enable, run frames, read the mean. Adds two event records per
 * frame; off by default. */
int oat_tracker_profile(oat_tracker *t, int enable);
int oat_tracker_profile_read(oat_tracker *t, double *mean_mog_kernel_ms, uint64_t *launches);

/* Diagnostic: enqueue only the fused MOG+HSV+threshold kernel for the next frame (device-resident
 * frame; the model advances exactly as in oat_tracker_submit, no detection is produced, nothing
 * to collect). For timing back-to-back launches of the dominant kernel in isolation. */
int oat_tracker_submit_fused_only(oat_tracker *t, const uint8_t *bgr_in, size_t in_pitch,
                                  double learning_rate, const oat_hsv_params *p);
/* Diagnostic: the detect tail of the most recently collected frame. out[15] = { status (0 = the
 * one-launch tail sufficed, 1 = replayed through the unbounded path), run-table entries needed,
 * replays so far, one-launch tail used (bit 0; bit 1: run table pre-labelled by the bands), 8 SM-clock stamps of its labelling CTA, frames run on the generic fused
 * kernel so far (adaptive kernel choice), 4-pixel groups that left the fused kernel's fast path in that frame,
 * frames served by the resident clip engine so far }. */
int oat_tracker_tail_stats(oat_tracker *t, uint32_t *out);

/* ---- synthetic frame source (measurement + parity) --------------------
 * Deterministic counter-hash stream: static background 40..120, +-3 noise, a filled disc of
 * colour BGR (40,220,60), radius rows/20, centre (cols/4 + 7t mod cols/2, rows/3 + 4t mod
 * rows/3), absent at t=0. Bit-identical to oracle/synth.py and oracle/oat_oracle.c. Writes
 * rows x cols x 3 BGR to dst (device or host). */
int oat_synth_frame(oat_ctx *ctx, uint8_t *dst, size_t pitch, int rows, int cols, uint32_t seed,
                    uint32_t t);

/* ---- memory helpers for the pinned-host / device-resident Frame variants --------------- */
int oat_alloc_device(oat_ctx *ctx, size_t bytes, void **out);
int oat_free_device(oat_ctx *ctx, void *p);
int oat_alloc_pinned(size_t bytes, void **out);
int oat_free_pinned(void *p);
/* Page-lock an existing mapping (e.g. a shmemdf segment) so copies from/to it are async DMA. */
int oat_register_host(void *p, size_t bytes);
int oat_unregister_host(void *p);
/* Device-resident Frame variant across processes (SharedFrameHeader memory kind DEVICE): export a device
 * allocation made with oat_alloc_device as a 64-byte handle (cudaIpcMemHandle_t) that travels in the
 * shared frame header; another process on the same GPU opens it and passes the pointer to any entry
 * point above. The exporter must outlive the importers' use. */
int oat_ipc_export(oat_ctx *ctx, const void *dev_ptr, unsigned char handle[64]);
int oat_ipc_open(oat_ctx *ctx, const unsigned char handle[64], void **dev_ptr);
int oat_ipc_close(oat_ctx *ctx, void *dev_ptr);
/* Copy between any two of {device, pinned, pageable}; synchronous. */
int oat_memcpy(oat_ctx *ctx, void *dst, const void *src, size_t bytes);
/* The same, asynchronous, on one of the context's two copy lanes (0: ingest, 1: egress; each a stream of its own, so
 * a frame on its way back to the host overlaps the next one on its way up and the kernels between them).  An egress
 * copy is ordered behind everything the context's compute stream was given before the call (it reads what the kernels
 * produce); an ingest copy is ordered behind nothing but the lane's earlier copies -- its destination must not be in
 * use by work still in flight (double-buffer it).  oat_memcpy_wait returns once the lane's copies have landed;
 * oat_memcpy_done polls (*done = 1: nothing in flight on the lane). */
int oat_memcpy_async(oat_ctx *ctx, int lane, void *dst, const void *src, size_t bytes);
int oat_memcpy_wait(oat_ctx *ctx, int lane);
int oat_memcpy_done(oat_ctx *ctx, int lane, int *done);
/* L2 flush for benchmarking: overwrites an internal buffer larger than L2. */
int oat_flush_l2(oat_ctx *ctx);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif

#ifdef __cplusplus
}
#endif
#endif /* OATGPU_H */
