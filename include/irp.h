/*
 * irp.h — C ABI of libirp_b200.so: the B200-native degradation-analysis +
 * preprocess hot path of RazonIn4K/image-restoration-platform.
 *
 * This is the drop-in boundary.  Every entry point below is
 * what a Node N-API addon (addon/irp_addon.cc) or the Python ctypes mirror
 * (image-restoration-platform_b200/_ffi.py) binds; signatures use plain
 * pointers and sizes only.  Each group cites the reference call it replaces
 * (paths relative to the reference repo root).
 *
 * Pixel layout everywhere: u8, interleaved HWC, `pitch` bytes between rows.
 * There is NO CPU fallback: without a usable CUDA device irp_create() fails
 * with IRP_ERR_NO_DEVICE and nothing else can be called.
 */
#ifndef IRP_H_
#define IRP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IRP_ABI_VERSION 1

/* status codes (0 = ok, <0 = error; message via irp_last_error) */
enum {
  IRP_OK = 0,
  IRP_ERR_BAD_ARG = -1,     /* null pointer, bad dims, bad channel count        */
  IRP_ERR_UNSUPPORTED = -2, /* C == 2, arithmetic-coded / lossless / CMYK JPEG, ...       */
  IRP_ERR_CUDA = -3,        /* a CUDA runtime call failed                       */
  IRP_ERR_NOMEM = -4,       /* host or device allocation failed                 */
  IRP_ERR_NO_DEVICE = -5,   /* no CUDA device / wrong architecture              */
  IRP_ERR_CAPACITY = -6     /* caller-provided output buffer too small          */
};

/* score order == key order of the object ClassifierService.analyze returns
 * (server-node/src/services/classifier.js:62-70). */
enum {
  IRP_SCORE_BLUR = 0,
  IRP_SCORE_NOISE = 1,
  IRP_SCORE_LOWLIGHT = 2,
  IRP_SCORE_COMPRESSION = 3,
  IRP_SCORE_SCRATCH = 4,
  IRP_SCORE_FADE = 5,
  IRP_SCORE_COLORSHIFT = 6,
  IRP_NUM_SCORES = 7
};

/* Switches for the libvips details SURVEY.md §8a tags MED/LOW confidence. */
enum { IRP_LUMA_VIPS_USUAL = 0 /* 0.2/0.7/0.1 */, IRP_LUMA_CIE = 1 /* 0.2126/0.7152/0.0722 */ };
enum { IRP_COEF_FIXED_POINT_SUM = 0 /* vips_vector_to_fixed_point */, IRP_COEF_TRUNCATE = 1 };
/* gaussblur(1) of _detectBlockiness (classifier.js:297): libvips' C path of vips_convi, (sum + 22) / 44 exactly,
 * or its SIMD path as recalled: 8-bit mantissas {35, 58, 35}, (sum + 64) >> 7 */
enum { IRP_BLUR_EXACT = 0, IRP_BLUR_VECTOR = 1 };
/* reducev / reduceh: 12-bit integer coefficients (the C and highway kernels), or 6 fractional bits ((sum + 32) >> 6,
 * the precision of libvips' older orc vector path) */
enum { IRP_REDUCE_INT12 = 0, IRP_REDUCE_VECTOR_2_6 = 1 };

#define IRP_MAX_DIMENSION 2048 /* imagePreprocess.js:4  MAX_DIMENSION */
#define IRP_FUSION_CANVAS 2048 /* SURVEY.md §8a row P5                */
#define IRP_FUSION_MAX_IMAGES 3

typedef struct irp_ctx irp_ctx;

typedef struct irp_opts {
  uint32_t struct_size; /* sizeof(irp_opts), for forward compatibility */
  int32_t luma_mode;    /* IRP_LUMA_*                                   */
  int32_t coef_mode;    /* IRP_COEF_*                                   */
  int32_t reserved0;
  uint64_t staging_bytes; /* pixels per pipeline chunk of a host-resident batch; 0 = 96 MiB */
  int32_t blur_mode;    /* IRP_BLUR_*   (read when struct_size covers it)       */
  int32_t reduce_mode;  /* IRP_REDUCE_*                                         */
} irp_opts;

typedef struct irp_image_desc {
  const uint8_t *pixels;    /* host or device pointer (see on_device)                 */
  size_t pitch;             /* bytes per row, >= width*channels                       */
  int32_t width, height;    /* STORED dims (pre-EXIF), as sharp metadata() reports    */
  int32_t channels;         /* 1, 3 or 4                                              */
  int32_t is_jpeg;          /* metadata.format === 'jpeg' (classifier.js:180)         */
  int32_t exif_orientation; /* 1..8; anything else is treated as 1                    */
  int32_t on_device;        /* 1: `pixels` already lives in this context's device HBM */
} irp_image_desc;

/* Everything the classifier derives from pixels.  Integer fields are exact and
 * order-independent; score[] is computed from them in IEEE double. */
typedef struct irp_result {
  double score[IRP_NUM_SCORES];
  uint64_t sum[4], sumsq[4];     /* per channel Σx, Σx²      (sharp.stats, classifier.js:52)   */
  uint64_t e_sum[2], e_sumsq[2]; /* clipped Lap8 / Sharp9 responses (classifier.js:107-118,135-145) */
  uint64_t b_sum, b_sumsq;       /* pooled Σ, Σ² of gaussblur(σ=1) bytes (classifier.js:297)  */
  uint32_t scratch_v, scratch_h; /* _detectLinearFeatures counts (classifier.js:310-337)      */
  uint32_t block_edges[2];       /* additive diagnostic: strong grey steps across 8-px column / row boundaries */
  uint32_t luma_hist[256];       /* additive diagnostic: histogram of the libvips B_W grey    */
  int32_t status;                /* per-image status (IRP_OK or an error code)                */
  /* PromptEnhancerService._identifyTopIssues (promptEnhancer.js:121-145): the scores above 0.3, highest first (ties
   * in key order, as a stable sort leaves them), at most three.  issues[k] = IRP_ISSUE(severity, score index) or
   * IRP_NO_ISSUE; issues[3] = how many.  Severity: >= 0.7 high, >= 0.5 medium, else low. */
  uint8_t issues[4];
} irp_result;
enum { IRP_SEVERITY_LOW = 1, IRP_SEVERITY_MEDIUM = 2, IRP_SEVERITY_HIGH = 3 };
#define IRP_ISSUE(severity, index) ((uint8_t)(((severity) << 4) | (index)))
#define IRP_ISSUE_INDEX(v) ((v) & 15)
#define IRP_ISSUE_SEVERITY(v) ((v) >> 4)
#define IRP_NO_ISSUE 0xFF
#define IRP_ISSUE_THRESHOLD 0.3   /* promptEnhancer.js:122 (and the logging filter, classifier.js:74) */

typedef struct irp_out_desc {
  uint8_t *pixels;  /* caller-owned destination (host or device)                        */
  size_t pitch;     /* in: bytes per row of the destination; 0 = tight (width*channels) */
  size_t capacity;  /* in: bytes available at `pixels`                                  */
  int32_t width, height, channels; /* out: dims of what was written                    */
  int32_t on_device;
} irp_out_desc;

/* per-call device timings of the last completed call on this context (ms, CUDA events
 * on the launch stream).  For a pipelined (chunks > 1) call classify_ms / preprocess_ms are
 * sums over chunks and h2d_ms / d2h_ms are 0: the copies overlap the kernels. */
typedef struct irp_timing {
  float h2d_ms, classify_ms, preprocess_ms, d2h_ms, total_ms;
  uint32_t kernel_launches; /* kernels of this library launched by the call */
  uint32_t chunks;          /* pipeline chunks the batch was cut into (1 = not pipelined) */
} irp_timing;

/* ---- library / context ------------------------------------------------- */
int irp_abi_version(void);
int irp_device_count(void);
/* Replaces `new ClassifierService({logger})` + sharp's global libvips init
 * (server-node/src/context/services.js:48-50).  One context per GPU. */
irp_ctx *irp_create(int device, const irp_opts *opts);
void irp_destroy(irp_ctx *ctx);
/* ctx may be NULL to read the error of a failed irp_create on this thread. */
const char *irp_last_error(const irp_ctx *ctx);
/* Launch on a caller-owned CUDA stream (cudaStream_t as void*); NULL = the
 * context's own stream. */
int irp_set_stream(irp_ctx *ctx, void *cuda_stream);
int irp_get_timing(const irp_ctx *ctx, irp_timing *out);

/* ---- pure host helpers (no GPU work) ----------------------------------- */
/* calculateResizeDimensions + sharp fit:'inside' after .rotate()
 * (server-node/src/middleware/imagePreprocess.js:7-22,42-53). */
int irp_preprocess_dims(int width, int height, int exif_orientation, int *out_w, int *out_h);
/* fusion canvas placement: resized dims + top-left offset. */
int irp_fusion_dims(int width, int height, int exif_orientation, int *out_w, int *out_h, int *off_x,
                    int *off_y);
/* the 7 JS formulas (classifier.js:119-121,146,159-167,180-186,223-228,240-253)
 * applied to the integer moments already in `r`; fills r->score. */
int irp_scores_from_moments(irp_result *r, int width, int height, int channels, int is_jpeg);
/* r->issues from r->score (called by irp_scores_from_moments; promptEnhancer.js:121-145) */
int irp_top_issues(irp_result *r);
/* copies of the baked grey tables (for the exhaustive 2^24 verification test) */
int irp_grey_tables(int luma_mode, uint32_t lut_r[256], uint32_t lut_g[256], uint32_t lut_b[256],
                    uint32_t inv[4096]);

/* ---- the hot path ------------------------------------------------------ */
/* ClassifierService.analyze for n images (classifier.js:40-99): stats,
 * grey, 3 stencils, blur delta, scratch grid — one pass over each image. */
int irp_classify_batch(irp_ctx *ctx, const irp_image_desc *imgs, int n, irp_result *results);
/* preprocessImage pixel stages (imagePreprocess.js:42-53): auto-orient,
 * lanczos3 fit-inside <= 2048 (with vips_resize's integer box pre-shrink when an
 * axis shrinks 4x or more), normalise to u8 RGB (grey stays 1 channel). */
int irp_preprocess_batch(irp_ctx *ctx, const irp_image_desc *imgs, int n, irp_out_desc *outs);
/* classify + preprocess of the same sources in one submission (BASELINE.json
 * configs[1]).  Two kernels, each reading the sources from HBM once: a batch is
 * far larger than the 126 MB L2, so the second pass is NOT served from cache
 * (profiles/: DRAM traffic 1.76x the fused figure; the kernels are issue-bound). */
int irp_analyze_batch(irp_ctx *ctx, const irp_image_desc *imgs, int n, irp_result *results,
                      irp_out_desc *outs);
/* up to 3 images -> aligned 2048x2048x3 canvases, centred, black pad
 * (SURVEY.md §8a row P5).  `canvases` has n_groups*3 entries; unused slots of
 * a group (pixels == NULL in `imgs`) are skipped. */
int irp_fusion_prepare_batch(irp_ctx *ctx, const irp_image_desc *imgs, int n_groups,
                             irp_out_desc *canvases);

/* ---- compressed input ------------------------------------------------- */
/* What the reference's callers actually hold is the FILE: analyze(imageBuffer)
 * and preprocessImage(req.file.buffer) hand sharp compressed bytes
 * (classifier.js:40,51-52; imagePreprocess.js:36-42), and sharp decodes them
 * with libjpeg-turbo (JDCT_ISLOW, fancy upsampling).  These entry points take
 * baseline JPEG bytes, decode them ON THE DEVICE pixel-exact with that decoder
 * (parallel Huffman decode, integer IDCT, triangle chroma upsampling, YCbCr ->
 * RGB) and feed the pixels to the same kernels — ~4 MB instead of 36 MB per
 * 12 MP photo crosses PCIe and the host decodes nothing.  8-bit Huffman JPEG:
 * baseline (SOF0/SOF1; one interleaved scan — the parallel self-synchronising
 * decoder — or one scan per component) and progressive (SOF2, what
 * imagePreprocess.js:57-61 writes: one warp per scan, dependent scans
 * overlapped block by block), 1 or 3 components, 4:4:4 / 4:2:2 / 4:2:0 / 4:4:0, with or without
 * restart markers; anything else returns IRP_ERR_UNSUPPORTED (no CPU fallback). */
typedef struct irp_jpeg_desc {
  const uint8_t *data;      /* host pointer to the JPEG file bytes          */
  size_t size;
  int32_t exif_orientation; /* 1..8 (from the caller's header parse), else 1 */
  int32_t scale_denom;      /* irp_decode_jpeg_batch only: decode at libjpeg scale 1 / 2, 4 or 8 (0, 1 = full size) */
} irp_jpeg_desc;
/* header only: stored dims and channels (1 grey, 3 colour) — sharp(buf).metadata(), classifier.js:51 */
int irp_jpeg_info(const uint8_t *data, size_t size, int *width, int *height, int *channels);
/* decoded pixels (u8 RGB or grey), to host or device buffers */
int irp_decode_jpeg_batch(irp_ctx *ctx, const irp_jpeg_desc *jpegs, int n, irp_out_desc *outs);
/* decode + classify (is_jpeg = 1) + preprocess; `results` or `outs` may be NULL */
int irp_analyze_jpeg_batch(irp_ctx *ctx, const irp_jpeg_desc *jpegs, int n, irp_result *results,
                           irp_out_desc *outs);

/* ---- compressed output ------------------------------------------------ */
/* preprocessImage ends in `.jpeg({quality: 85, chromaSubsampling: '4:4:4',
 * mozjpeg: true})` and returns the FILE (imagePreprocess.js:50-58).  These entry
 * points encode on the device, bit-identical to libjpeg-turbo's baseline encoder
 * (RGB->YCbCr, accurate integer DCT, reciprocal quantisation, Annex K Huffman
 * tables, 4:4:4, JFIF header) — a 2048x1536 result leaves the GPU as ~1 MB of
 * file bytes instead of 9.4 MB of pixels.  mozjpeg's trellis quantisation and
 * progressive scan search are NOT reproduced (same picture, ~10 % larger file). */
/* OR-ed into `quality`: Huffman tables optimised per image (libjpeg's optimize_coding, part of sharp's
 * `mozjpeg: true`): one more pass over the coefficients, files ~5-10 % smaller, byte-identical to
 * libjpeg-turbo's optimised sequential file */
#define IRP_JPEG_OPTIMIZE 0x100
typedef struct irp_jpeg_out {
  uint8_t *data;    /* caller-owned HOST buffer for the file                     */
  size_t capacity;  /* in: bytes available at `data`                             */
  size_t size;      /* out: bytes written (on IRP_ERR_CAPACITY: bytes required)  */
  int32_t width, height, channels; /* out: dims of the encoded image             */
  int32_t reserved;
} irp_jpeg_out;
/* `.withMetadata({icc: 'srgb'})` (imagePreprocess.js:57-67): an ICC profile attached to the files of an encode
 * call as APP2 "ICC_PROFILE" segments right behind the JFIF header (jpeg_write_icc_profile's layout, 65519 bytes
 * per segment).  The profile is chosen PER CALL: IRP_JPEG_ICC(id) OR-ed into `quality` names an immutable profile
 * of the context's registry — IRP_ICC_SRGB (always there: an sRGB IEC 61966-2.1 v4 matrix/TRC profile generated
 * by the library; libvips' own built-in file is not redistributed) or an id irp_register_icc returned.  Concurrent
 * calls with different profiles do not interfere.  Output buffers must hold the profile as well.
 * id 0 (no IRP_JPEG_ICC bits) = the context default set by irp_set_output_icc (none until then; NULL / 0 clears it):
 * kept for callers that attach one profile to everything — do not toggle it around calls. */
#define IRP_ICC_SRGB 1
#define IRP_JPEG_ICC(id) (((id) & 0xFF) << 16)
int irp_set_output_icc(irp_ctx *ctx, const uint8_t *profile, size_t size);
/* returns the id (>= 2) of the profile, or a negative status; the same bytes always get the same id */
int irp_register_icc(irp_ctx *ctx, const uint8_t *profile, size_t size);
/* the bytes of profile `id` (out may be NULL to ask for the size); ctx may be NULL for IRP_ICC_SRGB */
int irp_get_icc(irp_ctx *ctx, int id, uint8_t *out, size_t capacity, size_t *size);
/* pixels (host or device, 1 or 3 channels) -> baseline JPEG files */
int irp_encode_jpeg_batch(irp_ctx *ctx, const irp_image_desc *imgs, int n, int quality,
                          irp_jpeg_out *outs);
/* classify (results may be NULL) + preprocess + encode: raw pixels in, scores and the
 * preprocessed FILE out — analyze() and preprocessImage() of one upload */
int irp_analyze_encode_batch(irp_ctx *ctx, const irp_image_desc *imgs, int n, irp_result *results,
                             int quality, irp_jpeg_out *outs);
/* the same from JPEG FILE bytes: decode, classify, preprocess and re-encode without a
 * pixel crossing PCIe */
int irp_transcode_jpeg_batch(irp_ctx *ctx, const irp_jpeg_desc *jpegs, int n, irp_result *results,
                             int quality, irp_jpeg_out *outs);

/* ---- concurrent single-image requests --------------------------------- */
/* The reference's callers issue ONE image per call from several in-flight
 * promises (ClassifierService.analyze is async, classifier.js:40; restoreBatch
 * fans out with pLimit(3), restorator.js:196-211).  irp_submit queues one
 * image and returns at once; a dispatcher thread owned by the context gathers
 * whatever is queued (up to 64 requests, waiting at most ~100 us for company)
 * into ONE batched submission, so concurrent callers share launches and the
 * H2D / kernel / D2H pipeline instead of serialising on the context.  `result`
 * and / or `out` select classify, preprocess or both; the pointers (and the
 * pixels) must stay valid until irp_wait returns.  irp_wait blocks until that
 * request is done, returns its status, copies its error text (if any) and
 * releases the ticket.  A failing request does not fail its batch-mates. */
typedef struct irp_request *irp_ticket;
int irp_submit(irp_ctx *ctx, const irp_image_desc *img, irp_result *result, irp_out_desc *out,
               irp_ticket *ticket);
/* the same for a JPEG FILE (the bytes must stay valid until irp_wait returns): what analyze(imageBuffer)
 * hands over, decoded on the device with the rest of its batch */
int irp_submit_jpeg(irp_ctx *ctx, const irp_jpeg_desc *jpeg, irp_result *result, irp_out_desc *out,
                    irp_ticket *ticket);
/* the same with the preprocessed FILE as output (irp_transcode_jpeg_batch for one upload): analyze(buf) and
 * preprocessImage(buf) of one request; `result` may be NULL; on IRP_ERR_CAPACITY out->size is the size needed */
int irp_submit_transcode(irp_ctx *ctx, const irp_jpeg_desc *jpeg, irp_result *result, int quality,
                         irp_jpeg_out *out, irp_ticket *ticket);
int irp_wait(irp_ctx *ctx, irp_ticket ticket, char *err, size_t err_capacity);

/* ---- memory helpers (so non-CUDA hosts can stage device-resident data) -- */
void *irp_dev_alloc(irp_ctx *ctx, size_t bytes);
int irp_dev_free(irp_ctx *ctx, void *p);
void *irp_host_alloc_pinned(irp_ctx *ctx, size_t bytes);
int irp_host_free_pinned(irp_ctx *ctx, void *p);
int irp_memcpy_h2d(irp_ctx *ctx, void *dst, const void *src, size_t bytes);
int irp_memcpy_d2h(irp_ctx *ctx, void *dst, const void *src, size_t bytes);
int irp_synchronize(irp_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* IRP_H_ */
