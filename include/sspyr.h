/* include/sspyr.h -- the drop-in boundary: a C ABI (plain pointers and sizes, no C++/torch types) for
 * the B200-native SIFT scale-space builder.  One shared library, libsspyr.so, exports exactly the
 * symbols declared here; every one of them replaces a piece of the reference's serial class
 * `GaussPyramid` (ZhangShuui/SIFT-parallel-optimization, GuassDePyramid.h) as cited per function.
 *
 * The reference has no FFI: a variant is selected at compile time by `#include "<variant>.h"` plus a
 * class name (main.cpp:2-13,61).  The binding a maintainer adds is therefore a header-only C++ class
 * over this ABI -- include/GaussDePyramid-CUDA.h (`GaussPyramid_cuda`, same public members) -- see
 * INTEGRATION.md.
 *
 * Conventions
 *  - every function returns an int status: 0 = SSPYR_OK, negative = error (sspyr_last_error() has text);
 *    nothing throws across this boundary.  (The reference has no error handling at all.)
 *  - the handle owns all device memory; the caller owns every host pointer it passes; no pointer
 *    returned by the library outlives sspyr_destroy().
 *  - one handle = one GPU = one host thread at a time.  All work is enqueued on the handle's stream
 *    (sspyr_set_stream); calls that touch host memory synchronise only when they say so.
 *  - there is NO CPU fallback: without a CUDA device sspyr_create() fails with SSPYR_ERR_CUDA.
 *
 * Device layout of one frame slot, per octave o (H_o = height>>o, W_o = width>>o, pitch_o = W_o rounded
 * up to 32 floats so that every row starts on a 128-byte line), planes of H_o x pitch_o floats:
 *
 *      [ G_0 .. G_{S+1} | DoG_0 .. DoG_{S+1} | G_{S+2} ]
 *
 * so that the reference's own in-place result (slots 0..S+1 = DoG, slot S+2 = top Gaussian,
 * GuassDePyramid.h:136-149) is the contiguous tail [DoG_0 .. G_{S+2}] of each octave.
 */
#ifndef SSPYR_H
#define SSPYR_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSPYR_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define SSPYR_API __attribute__((visibility("default")))
#else
#define SSPYR_API
#endif

/* status codes */
#define SSPYR_OK 0
#define SSPYR_ERR_ARG (-1)      /* bad argument / configuration */
#define SSPYR_ERR_CUDA (-2)     /* CUDA runtime failure (no device, launch or copy error) */
#define SSPYR_ERR_NOMEM (-3)    /* host or device allocation failed */
#define SSPYR_ERR_STATE (-4)    /* call made in the wrong state (e.g. download before build) */
#define SSPYR_ERR_UNSUPPORTED (-5)

/* mode: which "Gaussian" the levels hold */
#define SSPYR_MODE_REF 0  /* the reference's semantics: pointwise centred window, GuassDePyramid.h:106-134 */
#define SSPYR_MODE_CONV 1 /* true separable Gaussian blur + decimation chain (north_star); no upstream parity */

/* outputs mask */
#define SSPYR_OUT_GAUSS 1     /* all S+3 Gaussian levels */
#define SSPYR_OUT_DOG 2       /* S+2 DoG levels */
#define SSPYR_OUT_GAUSS_TOP 4 /* only G_{S+2}: with DOG this is exactly the reference's in-place result */
#define SSPYR_OUT_EXTREMA 8   /* 26-neighbour DoG extremum flags (uint8, S planes per octave); beyond the reference */
#define SSPYR_OUT_KEYPOINTS 16 /* the same extrema as a compacted list of sspyr_keypoint records per frame slot */
#define SSPYR_OUT_INPLACE (SSPYR_OUT_DOG | SSPYR_OUT_GAUSS_TOP)
#define SSPYR_OUT_ALL (SSPYR_OUT_GAUSS | SSPYR_OUT_DOG)

/* pixel type of the input frame */
#define SSPYR_PIXEL_I32 0 /* the reference's `int**` image (GuassDePyramid.h:13,36) */
#define SSPYR_PIXEL_F32 1 /* already-float pixels (e.g. [0,1]-normalised) */
#define SSPYR_PIXEL_U8 2  /* 8-bit grayscale: 1 byte/pixel ingest */

/* plane kinds for download / device_ptr */
#define SSPYR_KIND_GAUSS 0
#define SSPYR_KIND_DOG 1
#define SSPYR_KIND_INPLACE 2 /* slot s of the reference's in-place layout: s<=S+1 -> DoG_s, s==S+2 -> G_{S+2} */
#define SSPYR_KIND_EXTREMA 3
#define SSPYR_KIND_KEYPOINTS 4 /* sspyr_device_ptr only: the slot's keypoint buffer, [uint32 count, uint32 capacity, 0, 0][records] */

/* stages of the reference's pipeline a build can stop after (sspyr_build_stage) */
#define SSPYR_STAGE_INIT 0   /* K0 only: every level = decimated original, GuassDePyramid.h:76-86 */
#define SSPYR_STAGE_FILTER 1 /* K0 + GaussFilter on every octave: Gaussian levels, no DoG (:106-134) */
#define SSPYR_STAGE_DOG 2    /* the full GenerateDoG result (:136-149); what sspyr_build does */

#define SSPYR_MAX_OCTAVES 16
#define SSPYR_MAX_LEVELS 16 /* S + 3 <= 16 */

typedef struct sspyr_ctx* sspyr_handle;

/* One DoG extremum (SSPYR_OUT_KEYPOINTS): a pixel of DoG level `level` (1..S) of octave `octave` that is strictly greater
 * or strictly smaller than its 26 neighbours in (level-1, level, level+1) and exceeds extrema_thresh in magnitude.
 * x = column, y = row in that octave's coordinates (of the full image; row bands are not supported).  16 bytes. */
typedef struct sspyr_keypoint {
    int32_t x, y;
    int32_t octave_level;  /* octave << 16 | level */
    float value;           /* the DoG value */
} sspyr_keypoint;

/* Replaces the reference's compile-time constants and constructor arguments:
 *   sigma (GuassDePyramid.h:7), ctor (int** img, int len, int S) (:36), layer = floor(log2 len)+1 (:48-53). */
typedef struct sspyr_config {
    int32_t height;        /* rows held by THIS handle: a whole frame, or one row band of it */
    int32_t width;         /* columns */
    int32_t octaves;       /* 0 = all = floor(log2(min(full_height,width)))+1, the reference's rule */
    int32_t S;             /* scales per octave; S+3 levels, S+2 DoGs (GuassDePyramid.h:64,140) */
    float sigma0;          /* <= 0: mode default (REF 2.0 as GuassDePyramid.h:7; CONV 1.6) */
    int32_t mode;          /* SSPYR_MODE_* */
    int32_t outputs;       /* SSPYR_OUT_* mask; 0 = SSPYR_OUT_ALL */
    int32_t pixel_type;    /* SSPYR_PIXEL_* */
    int32_t frames;        /* resident frame slots (>= 1); build_batch cycles through them */
    int32_t device;        /* CUDA ordinal; -1 = current device */
    int32_t band_row0;     /* first row of this band in the full image (multiple of 2^(octaves-1)); 0 for a whole frame */
    int32_t full_height;   /* rows of the full image; 0 = height */
    float sigma_in;        /* CONV: blur assumed present in the input (default 0.5) */
    float radius_sigmas;   /* CONV: tap radius = ceil(radius_sigmas * sigma_inc) (default 3.0) */
    float extrema_thresh;  /* SSPYR_OUT_EXTREMA / _KEYPOINTS: |DoG| must exceed this */
    int32_t max_keypoints; /* SSPYR_OUT_KEYPOINTS: records kept per frame slot (0 = 1 << 20); extrema beyond it are counted, not stored */
    int32_t reserved[7];   /* must be zero */
} sspyr_config;

/* ---- life cycle ---------------------------------------------------------------------------------- */
SSPYR_API int sspyr_version(void);
/* Fill *cfg with defaults: S=3, mode REF, outputs ALL, i32 pixels, 1 frame, current device. */
SSPYR_API int sspyr_default_config(sspyr_config* cfg);
/* Allocates device buffers and the window/tap tables.  Replaces GaussPyramid::GaussPyramid(img,len,S)
 * minus the image copy (GuassDePyramid.h:36-58) and the allocation half of GaussPyInit (:62-72). */
SSPYR_API int sspyr_create(const sspyr_config* cfg, sspyr_handle* out);
/* Replaces GaussPyramid::~GaussPyramid (GuassDePyramid.h:151-170). */
SSPYR_API int sspyr_destroy(sspyr_handle h);
/* Text of the last error on this handle (or of the last failed sspyr_create when h == NULL). */
SSPYR_API const char* sspyr_last_error(sspyr_handle h);

/* ---- geometry: the integer quantities that must be bit-exact ---------------------------------------- */
SSPYR_API int sspyr_num_octaves(sspyr_handle h);  /* GuassDePyramid.h:48-53 */
SSPYR_API int sspyr_num_levels(sspyr_handle h);   /* S+3, GuassDePyramid.h:64 */
SSPYR_API int sspyr_num_dogs(sspyr_handle h);     /* S+2, GuassDePyramid.h:140 */
/* H_o, W_o (GuassDePyramid.h:66 `length/step`) and the device row pitch in floats. */
SSPYR_API int sspyr_level_dims(sspyr_handle h, int octave, int* rows, int* cols, size_t* pitch_floats);
/* Compulsory traffic of one frame in bytes for the configured outputs: input read once + every
 * requested output plane written once (B_full / B_ref). */
SSPYR_API int sspyr_algorithmic_bytes(sspyr_handle h, uint64_t* bytes);

/* ---- streams ------------------------------------------------------------------------------------ */
/* Use the caller's cudaStream_t (passed as void*; NULL = the legacy default stream). */
SSPYR_API int sspyr_set_stream(sspyr_handle h, void* cuda_stream);

/* ---- input: replaces the deep copy in the constructor (GuassDePyramid.h:38-46) ------------------- */
/* Copy one host frame (pixel_type elements, `pitch_bytes` between rows, 0 = tight) into frame slot
 * `frame`.  Asynchronous on the handle's stream when `host` is pinned memory. */
SSPYR_API int sspyr_upload(sspyr_handle h, int frame, const void* host, size_t pitch_bytes);
/* Point frame slot `frame` at a device-resident image instead (no copy; caller keeps it alive).
 * pitch_bytes must be a multiple of 16. */
SSPYR_API int sspyr_set_input_device(sspyr_handle h, int frame, const void* dev, size_t pitch_bytes);

/* ---- the hot path: GaussPyInit + GaussFilter(all octaves) + GenerateDoG in one fused pass ---------- */
/* Replaces GaussPyInit's decimating copy (GuassDePyramid.h:76-86), GaussFilter (:106-134) and
 * GenerateDoG (:136-149) for frame slot `frame`.  Asynchronous. */
SSPYR_API int sspyr_build(sspyr_handle h, int frame);
/* Stop after an earlier stage of the reference's pipeline (SSPYR_STAGE_*): INIT and FILTER write the
 * S+3 Gaussian-level planes only (needs SSPYR_OUT_GAUSS) -- the states the reference object is in after
 * GaussPyInit() and after GaussFilter() on every octave.  REF mode. */
SSPYR_API int sspyr_build_stage(sspyr_handle h, int frame, int stage);
/* Same for `count` consecutive slots starting at `first` (wraps modulo cfg.frames). */
SSPYR_API int sspyr_build_batch(sspyr_handle h, int first, int count);
SSPYR_API int sspyr_sync(sspyr_handle h);
/* Device time of the most recent sspyr_build / sspyr_build_batch (CUDA events on the handle's stream);
 * synchronises.  Replaces the wall-clock bracket of main.cpp:67-69.  Needs sspyr_set_tuning(h,"timing",1)
 * first: event records between two builds keep consecutive frames from overlapping, so they are opt-in. */
SSPYR_API int sspyr_elapsed_ms(sspyr_handle h, float* ms);
/* Number of kernels the most recent build enqueued. */
SSPYR_API int sspyr_last_launches(sspyr_handle h);

/* ---- results: replaces reading the public `float**** GaussPy` (GuassDePyramid.h:16) ---------------- */
/* Copy one plane to host (`pitch_bytes` between rows, 0 = tight).  Synchronises the stream. */
SSPYR_API int sspyr_download(sspyr_handle h, int frame, int octave, int level, int kind, void* dst, size_t pitch_bytes);
/* Copy the reference's whole in-place result of one frame, dense: for each octave, S+3 planes of
 * H_o x W_o floats back to back (DoG_0..DoG_{S+1}, G_{S+2}).  `dst` must hold
 * (S+3) * sum_o H_o*W_o floats.  Asynchronous when `dst` is pinned; call sspyr_sync() before reading. */
SSPYR_API int sspyr_download_inplace(sspyr_handle h, int frame, float* dst);
/* Same for all Gaussian levels: S+3 planes per octave. */
SSPYR_API int sspyr_download_gauss(sspyr_handle h, int frame, float* dst);
/* The keypoint list of a frame slot (SSPYR_OUT_KEYPOINTS; whole frames only -- a row band would treat its seams as image
 * borders, so SSPYR_OUT_EXTREMA / _KEYPOINTS are rejected on band handles).  Enqueues two copies on the handle's stream:
 * the number of extrema found -> *count, and the first min(capacity, max_keypoints) records -> dst, in no particular order.
 * Both are asynchronous when the destinations are pinned (sspyr_host_alloc); call sspyr_sync() before reading.  *count may
 * exceed what was stored: the list then holds the first max_keypoints extrema that reached the cursor. */
SSPYR_API int sspyr_download_keypoints(sspyr_handle h, int frame, sspyr_keypoint* dst, int capacity, int* count);
/* Device pointer of a plane (valid until sspyr_destroy).  From the first call on, CONV builds are ordered after
 * everything already enqueued on the handle's stream (the caller's kernels may now read the planes). */
SSPYR_API int sspyr_device_ptr(sspyr_handle h, int frame, int octave, int level, int kind, void** ptr);

/* ---- pinned host memory: lets a C/C++ caller get async copies without including CUDA headers ------- */
SSPYR_API int sspyr_host_alloc(size_t bytes, void** ptr);
SSPYR_API int sspyr_host_free(void* ptr);

/* ---- introspection used by the tests ------------------------------------------------------------ */
/* Copy the window table of (octave, level) out: axis 0 = row window (over this band's rows),
 * axis 1 = column window.  REF mode only (K1, GuassDePyramid.h:118-121). */
SSPYR_API int sspyr_window_table(sspyr_handle h, int octave, int level, int axis, float* dst, int capacity);
/* CONV mode: taps of level s (2R+1 floats, centre at R); returns R through *radius. */
SSPYR_API int sspyr_conv_taps(sspyr_handle h, int level, float* dst, int capacity, int* radius);
/* Kernel tuning knobs (bench/sweeps): key in {"rows_per_thread","block","bx","pdl","timing","occ","prefetch_next","conv_streams",
 * "conv_march","conv_graph","conv_tma","conv_waves","conv_seg_min","conv_fused_sync","conv_chain","conv_l2hint","conv_lanes",
 * "conv_cascade","conv_casc_seg","conv_band_chain","conv_band_lanes"}.  The ones that change behaviour a caller can observe:
 *   conv_lanes      (default 8) CONV builds of different frame slots that may be in flight at once; 1 = strictly one
 *                   after the other.  Work enqueued on the handle's stream after a build always sees it complete.
 *   conv_band_lanes (default 6) the same for row bands that read their neighbours' planes over peer memory.
 *   conv_chain      (default 1) consecutive levels of an octave overlap through per-segment counters; 0 = a level
 *                   starts when the previous grid has completed; 2 = chained even for grids of less than a wave.
 *   conv_cascade    (default 0) 2 = build a whole pyramid with ONE launch (all levels pipelined through L2; same bits;
 *                   measured slower than one launch per level); 1 = only for frames of >= 4 Mpixel.
 *   conv_band_chain (default 0) 1 = chain levels across a band seam through the neighbours' segment counters. */
SSPYR_API int sspyr_set_tuning(sspyr_handle h, const char* key, int value);

/* ---- row-band halo exchange (CONV mode, multi-GPU): see DESIGN.md "Row bands" ----------------------- */
/* Rows of halo this band needs from each neighbour before level `level` of `octave` can be blurred. */
SSPYR_API int sspyr_halo_rows(sspyr_handle h, int octave, int level, int* rows);
/* Device pointers of the halo staging areas for (octave, level): the rows this band must SEND up/down
 * (inside its own planes) and the buffers it RECEIVES into; each is rows x pitch_o floats. */
SSPYR_API int sspyr_halo_ptrs(sspyr_handle h, int frame, int octave, int level, void** send_up, void** send_down,
                    void** recv_up, void** recv_down, size_t* bytes);
/* CONV build split at level granularity so the host can exchange halos between steps:
 * runs the blur that PRODUCES (octave, level). */
SSPYR_API int sspyr_conv_step(sspyr_handle h, int frame, int octave, int level);

/* ---- row-band CONV over NVLink peer memory: halo rows are read inside the blur kernel ------------------- */
/* Instead of exchanging halo rows into staging buffers (sspyr_halo_ptrs), a band can ATTACH its neighbours:
 * the level kernels then load the rows they need straight from the neighbour's planes over NVLink, and tiny
 * signal / wait kernels on a progress counter in peer memory keep the bands in step.  One process per GPU:
 * export a blob here, ship it to the neighbours (any host transport), attach it there (CUDA IPC).  Bands of one
 * process attach each other's handles directly.  Every band must issue the same sspyr_conv_step sequence;
 * inputs must be uploaded (and the ranks synchronised) before the first step of a frame. */
#define SSPYR_IPC_BLOB_BYTES 1024
#define SSPYR_SIDE_ABOVE 0
#define SSPYR_SIDE_BELOW 1
SSPYR_API int sspyr_ipc_export(sspyr_handle h, void* blob, size_t capacity, size_t* bytes);
SSPYR_API int sspyr_ipc_attach(sspyr_handle h, int side, const void* blob, size_t bytes);
SSPYR_API int sspyr_peer_attach_local(sspyr_handle h, int side, sspyr_handle neighbour);

#ifdef __cplusplus
}
#endif
#endif /* SSPYR_H */
