/*
 * irp_oracle.c — CPU restatement of the reference hot path.  TEST
 * INFRASTRUCTURE ONLY: nothing in the product (image-restoration-platform_b200/,
 * include/) links, imports or calls this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may.
 *
 * PARITY UNPINNED at exact-value level: the arithmetic of this path lives in
 * sharp 0.33.5 (server-node/package.json:37) -> @img/sharp-libvips-linux-x64 1.0.4
 * (libvips 8.15.x, server-node/package-lock.json:5215), neither vendored under the
 * reference tree nor installable here (no node, no libvips, no network).  The
 * reference's own tests pin only thresholds (server-node/tests/classifierService.test.js
 * :23,32,39,46,53-56).  What pins this file: those thresholds + the hand-derived
 * known answers of SURVEY.md §8c (tests/test_oracle.py), PIL exif_transpose for P2
 * and scipy.ndimage.correlate for K1.  What WOULD pin it: tools/sharp_golden.mjs runs the
 * reference's own ClassifierService / preprocessImage and their literal sharp calls on the
 * seeded fixtures of tests/golden/sharp_cases.py wherever Node exists; tests/test_sharp_golden.py
 * consumes that dump and votes on every switch below (luma_mode, coef_mode, blur_mode,
 * reduce_mode — the _m entry points take them all).
 *
 * Every function cites the reference lines it follows.  JS-side formulas are
 * literal (two-pass, sequential double accumulation as Array.prototype.reduce
 * does); libvips-side steps follow the published algorithms as recalled in
 * SURVEY.md §8a.
 */
#include "irp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../include/irp_spec.h"

/* ------------------------------------------------------------------ */
/* G1: libvips colourspace(B_W) for sRGB u8                           */
/* libvips colour/LabQ2sRGB.c calcul_tables(), colour/sRGB2scRGB.c,    */
/* colour/scRGB2BW.c (vips_col_scRGB2BW_8) — SURVEY.md §8a row G1      */
/* ------------------------------------------------------------------ */
/* libvips' VIPS_RINT: round half away from zero through a double add + int cast (NOT C rint) */
#define VIPS_RINT(R) ((int)((R) > 0 ? ((R) + 0.5) : ((R)-0.5)))

static float v2Y_8[256]; /* sRGB byte -> linear, float  */
static int Y2v_8[257];   /* linear*255 -> sRGB byte, int, [256] = [255] */
static int tables_ready = 0;

static void make_tables(void) {
  if (tables_ready) return;
  for (int i = 0; i < 256; i++) {
    float f = (float)i / 255;
    float v;
    if (f <= 0.0031308)
      v = 12.92 * f;
    else
      v = (1.0 + 0.055) * pow(f, 1.0 / 2.4) - 0.055;
    Y2v_8[i] = VIPS_RINT(255 * v);
  }
  Y2v_8[256] = Y2v_8[255];
  for (int i = 0; i < 256; i++) {
    float f = (float)i / 255;
    if (f <= 0.04045)
      v2Y_8[i] = f / 12.92;
    else
      v2Y_8[i] = pow((f + 0.055) / (1 + 0.055), 2.4);
  }
  tables_ready = 1;
}

void orc_tables(float v2y[256], int y2v[257]) {
  make_tables();
  memcpy(v2y, v2Y_8, sizeof v2Y_8);
  memcpy(y2v, Y2v_8, sizeof Y2v_8);
}

uint8_t orc_grey_rgb(int r, int g, int b, int luma_mode) {
  make_tables();
  float R = v2Y_8[r], G = v2Y_8[g], B = v2Y_8[b];
  float Y;
  /* double-promoted products, stored to float (C semantics of the libvips line) */
  if (luma_mode == 1)
    Y = 0.2126 * R + 0.7152 * G + 0.0722 * B;
  else
    Y = 0.2 * R + 0.7 * G + 0.1 * B; /* "the usual ratio" */
  float Yf = Y * 255;
  if (Yf < 0) Yf = 0;
  if (Yf > 255) Yf = 255;
  int Yi = (int)Yf;
  float v = Y2v_8[Yi] + (Y2v_8[Yi + 1] - Y2v_8[Yi]) * (Yf - Yi);
  return (uint8_t)VIPS_RINT(v);
}

int orc_grey(const uint8_t *px, int w, int h, int c, size_t pitch, int luma_mode, uint8_t *out) {
  if (!px || !out || w <= 0 || h <= 0) return -1;
  if (c != 1 && c != 3 && c != 4) return -2;
  for (int y = 0; y < h; y++) {
    const uint8_t *row = px + (size_t)y * pitch;
    for (int x = 0; x < w; x++)
      out[(size_t)y * w + x] =
          c == 1 ? row[x] : orc_grey_rgb(row[x * c], row[x * c + 1], row[x * c + 2], luma_mode);
  }
  return 0;
}

/* ------------------------------------------------------------------ */
/* K1 + K2: .convolve(3x3) on the grey image then .raw() u8 cast      */
/* classifier.js:107-115,135-143,199-207; sharp defaults scale = Σk    */
/* clamped to >= 1, offset 0; vips_conv replicate edges; cast clips.   */
/* ------------------------------------------------------------------ */
static const int K_LAP8[9] = {-1, -1, -1, -1, 8, -1, -1, -1, -1};   /* classifier.js:112 */
static const int K_SHARP9[9] = {-1, -1, -1, -1, 9, -1, -1, -1, -1}; /* classifier.js:140 */
static const int K_LAP4[9] = {0, -1, 0, -1, 4, -1, 0, -1, 0};       /* classifier.js:204 */

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

int orc_stencil(const uint8_t *grey, int w, int h, int which, uint8_t *out) {
  const int *k = which == 0 ? K_LAP8 : which == 1 ? K_SHARP9 : K_LAP4;
  int scale = 0;
  for (int i = 0; i < 9; i++) scale += k[i];
  if (scale < 1) scale = 1; /* sharp lib/operation.js convolve(): scale default = sum, min 1 */
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int acc = 0;
      for (int j = 0; j < 3; j++)
        for (int i = 0; i < 3; i++) {
          int yy = clampi(y + j - 1, 0, h - 1), xx = clampi(x + i - 1, 0, w - 1);
          acc += k[j * 3 + i] * grey[(size_t)yy * w + xx];
        }
      float r = (float)acc / (float)scale; /* float conv; exact integer here */
      out[(size_t)y * w + x] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
    }
  return 0;
}

/* ------------------------------------------------------------------ */
/* A4 inner: .blur(1) -> vips_gaussblur(sigma 1, min_ampl 0.2, integer) */
/* classifier.js:297; separable [12,20,12]/44, horizontal then         */
/* vertical, u8 between passes, replicate edges                        */
/* ------------------------------------------------------------------ */
/* blur_mode 0: libvips' C path of vips_convi — exact (sum + scale / 2) / scale.
 * blur_mode 1 (IRP_BLUR_VECTOR): the SIMD path of vips_convi as recalled (convi.c vips_convi_intize + the
 * highway kernel): the mask is baked into 8-bit mantissas with a shared exponent, here
 * shift = ceil(log2(20/44) + 1) = 0, exp = 7, mant = rint(128 * {12,20,12}/44) = {35, 58, 35} (sum 128), and a
 * pixel is (sum mant*p + 64) >> 7.  Which of the two the sharp prebuilt takes is not checkable offline (SURVEY.md
 * section 8a row A4 tags it LOW); tests/test_sharp_golden.py decides once a sharp dump exists. */
static inline int blur_tap(int l, int c, int r, int blur_mode) {
  if (blur_mode == 1) return (35 * (l + r) + 58 * c + 64) >> 7;
  return (IRP_GAUSS_EDGE * (l + r) + IRP_GAUSS_CENTRE * c + IRP_GAUSS_SCALE / 2) / IRP_GAUSS_SCALE;
}

int orc_blur1_m(const uint8_t *px, int w, int h, int c, size_t pitch, int blur_mode, uint8_t *out) {
  size_t n = (size_t)w * h * c;
  uint8_t *tmp = (uint8_t *)malloc(n);
  if (!tmp) return -4;
  for (int y = 0; y < h; y++) {
    const uint8_t *row = px + (size_t)y * pitch;
    for (int x = 0; x < w; x++) {
      int xl = x > 0 ? x - 1 : 0, xr = x < w - 1 ? x + 1 : w - 1;
      for (int ch = 0; ch < c; ch++)
        tmp[((size_t)y * w + x) * c + ch] = (uint8_t)blur_tap(row[xl * c + ch], row[x * c + ch], row[xr * c + ch], blur_mode);
    }
  }
  for (int y = 0; y < h; y++) {
    int yu = y > 0 ? y - 1 : 0, yd = y < h - 1 ? y + 1 : h - 1;
    for (size_t i = 0; i < (size_t)w * c; i++)
      out[(size_t)y * w * c + i] =
          (uint8_t)blur_tap(tmp[(size_t)yu * w * c + i], tmp[(size_t)y * w * c + i], tmp[(size_t)yd * w * c + i], blur_mode);
  }
  free(tmp);
  return 0;
}

int orc_blur1(const uint8_t *px, int w, int h, int c, size_t pitch, uint8_t *out) {
  return orc_blur1_m(px, w, h, c, pitch, 0, out);
}

/* ------------------------------------------------------------------ */
/* JS helpers, literal: classifier.js:262-270                          */
/* ------------------------------------------------------------------ */
static double js_variance(const uint8_t *buf, size_t n) {
  double sum = 0; /* buffer.reduce((sum, val) => sum + val, 0) */
  for (size_t i = 0; i < n; i++) sum = sum + buf[i];
  double mean = sum / (double)n;
  double acc = 0; /* reduce((sum, val) => sum + Math.pow(val - mean, 2), 0) */
  for (size_t i = 0; i < n; i++) {
    double d = buf[i] - mean;
    acc = acc + d * d;
  }
  return acc / (double)n;
}

static void moments(const uint8_t *buf, size_t n, uint64_t *s, uint64_t *s2) {
  uint64_t a = 0, b = 0;
  for (size_t i = 0; i < n; i++) {
    a += buf[i];
    b += (uint64_t)buf[i] * buf[i];
  }
  *s = a;
  *s2 = b;
}

/* _detectLinearFeatures, classifier.js:310-337 */
static void linear_features(const uint8_t *e, int width, int height, uint32_t *vc, uint32_t *hc) {
  uint32_t verticalCount = 0, horizontalCount = 0;
  const int threshold = IRP_SCRATCH_THRESHOLD;
  for (int y = 0; y < height; y += IRP_SCRATCH_STRIDE)
    for (int x = 0; x < width; x += IRP_SCRATCH_STRIDE) {
      size_t idx = (size_t)y * width + x;
      if (e[idx] > threshold) {
        if (x + 1 < width) verticalCount += e[idx + 1] > threshold ? 1 : 0;
        if (y + 1 < height) horizontalCount += e[idx + width] > threshold ? 1 : 0;
      }
    }
  *vc = verticalCount;
  *hc = horizontalCount;
}

static double dmin(double a, double b) { return (a != a || b != b) ? NAN : (a < b ? a : b); } /* Math.min */
static double dmax(double a, double b) { return (a != a || b != b) ? NAN : (a > b ? a : b); } /* Math.max */

/* ------------------------------------------------------------------ */
/* A0: ClassifierService.analyze, classifier.js:40-99                  */
/* ------------------------------------------------------------------ */
int orc_classify(const uint8_t *px, int w, int h, int c, size_t pitch, int is_jpeg, int luma_mode,
                 irp_result *out) {
  return orc_classify_m(px, w, h, c, pitch, is_jpeg, luma_mode, 0, out);
}

int orc_classify_m(const uint8_t *px, int w, int h, int c, size_t pitch, int is_jpeg, int luma_mode, int blur_mode,
                   irp_result *out) {
  if (!px || !out || w <= 0 || h <= 0 || pitch < (size_t)w * c) return -1;
  if (c != 1 && c != 3 && c != 4) return -2;
  memset(out, 0, sizeof *out);
  const size_t N = (size_t)w * h;

  /* S0: sharp(buf).stats() -> vips_stats: per-channel sum, sum of squares;
   * mean = s/N; deviation = sqrt(fabs(s2 - s*s/N) / (N - 1))  (classifier.js:52) */
  double mean[4], stdev[4];
  for (int ch = 0; ch < c; ch++) {
    uint64_t s = 0, s2 = 0;
    for (int y = 0; y < h; y++) {
      const uint8_t *row = px + (size_t)y * pitch;
      for (int x = 0; x < w; x++) {
        unsigned v = row[x * c + ch];
        s += v;
        s2 += v * v;
      }
    }
    out->sum[ch] = s;
    out->sumsq[ch] = s2;
    double ds = (double)s, ds2 = (double)s2, vals = (double)N;
    mean[ch] = ds / vals;
    stdev[ch] = sqrt(fabs(ds2 - (ds * ds / vals)) / (vals - 1));
  }

  uint8_t *grey = (uint8_t *)malloc(N), *e = (uint8_t *)malloc(N);
  uint8_t *o = (uint8_t *)malloc(N * c), *b = (uint8_t *)malloc(N * c);
  if (!grey || !e || !o || !b) {
    free(grey); free(e); free(o); free(b);
    return -4;
  }
  orc_grey(px, w, h, c, pitch, luma_mode, grey);
  for (int i = 0; i < 256; i++) out->luma_hist[i] = 0;
  for (size_t i = 0; i < N; i++) out->luma_hist[grey[i]]++;
  /* additive diagnostic, defined in include/irp.h */
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      int g = grey[(size_t)y * w + x];
      if ((x & 7) == 7 && x + 1 < w && abs(grey[(size_t)y * w + x + 1] - g) > IRP_BLOCK_EDGE_THRESHOLD)
        out->block_edges[0]++;
      if ((y & 7) == 7 && y + 1 < h && abs(grey[(size_t)(y + 1) * w + x] - g) > IRP_BLOCK_EDGE_THRESHOLD)
        out->block_edges[1]++;
    }

  /* A1 _analyzeBlur, classifier.js:104-127 */
  orc_stencil(grey, w, h, 0, e);
  moments(e, N, &out->e_sum[0], &out->e_sumsq[0]);
  {
    double edgeVariance = js_variance(e, N);
    double normalizedVariance = dmin(edgeVariance / IRP_BLUR_VAR_DIVISOR, 1.0);
    out->score[IRP_SCORE_BLUR] = dmax(0, 1.0 - normalizedVariance);
  }
  /* A2 _analyzeNoise, classifier.js:132-151 */
  orc_stencil(grey, w, h, 1, e);
  moments(e, N, &out->e_sum[1], &out->e_sumsq[1]);
  {
    double noiseLevel = sqrt(js_variance(e, N));
    out->score[IRP_SCORE_NOISE] = dmin(noiseLevel / IRP_NOISE_STD_DIVISOR, 1.0);
  }
  /* A3 _analyzeLowLight, classifier.js:156-172 */
  {
    double sum = 0;
    for (int ch = 0; ch < c; ch++) sum = sum + mean[ch];
    double normalizedBrightness = (sum / c) / 255;
    out->score[IRP_SCORE_LOWLIGHT] =
        normalizedBrightness < IRP_LOWLIGHT_KNEE ? dmin((IRP_LOWLIGHT_KNEE - normalizedBrightness) * 2, 1.0) : 0.0;
  }
  /* A4 _analyzeCompression + _detectBlockiness, classifier.js:177-191,288-308 */
  for (int y = 0; y < h; y++) memcpy(o + (size_t)y * w * c, px + (size_t)y * pitch, (size_t)w * c);
  orc_blur1_m(px, w, h, c, pitch, blur_mode, b);
  moments(b, N * c, &out->b_sum, &out->b_sumsq);
  if (is_jpeg) {
    double originalVariance = js_variance(o, N * c);
    double blurredVariance = js_variance(b, N * c);
    double varianceDelta = dmax(0, originalVariance - blurredVariance);
    out->score[IRP_SCORE_COMPRESSION] = dmin(dmin(varianceDelta / IRP_COMPRESSION_DIVISOR, 1.0), 1.0);
  } else {
    out->score[IRP_SCORE_COMPRESSION] = 0.0;
  }
  /* A5 _analyzeScratch, classifier.js:196-215 */
  orc_stencil(grey, w, h, 2, e);
  linear_features(e, w, h, &out->scratch_v, &out->scratch_h);
  {
    double total = (double)out->scratch_v + (double)out->scratch_h;
    out->score[IRP_SCORE_SCRATCH] = dmin(dmin(total / IRP_SCRATCH_DIVISOR, 1.0), 1.0);
  }
  /* A6 _analyzeFade, classifier.js:220-233,272-286 */
  {
    double colorfulness;
    if (c < 3)
      colorfulness = 0.5;
    else
      colorfulness = dmin(sqrt(pow(stdev[0], 2) + pow(stdev[1], 2) + pow(stdev[2], 2)) / 255, 1.0);
    double sum = 0;
    for (int ch = 0; ch < c; ch++) sum = sum + stdev[ch];
    double contrast = dmin((sum / c) / IRP_CONTRAST_DIVISOR, 1.0);
    double fadeScore = (1.0 - colorfulness) * 0.6 + (1.0 - contrast) * 0.4;
    out->score[IRP_SCORE_FADE] = dmin(fadeScore, 1.0);
  }
  /* A7 _analyzeColorShift, classifier.js:238-258 */
  if (c < 3) {
    out->score[IRP_SCORE_COLORSHIFT] = 0.0;
  } else {
    double avgMean = (mean[0] + mean[1] + mean[2]) / 3;
    double rd = avgMean > 0 ? fabs(mean[0] - avgMean) / avgMean : 0;
    double gd = avgMean > 0 ? fabs(mean[1] - avgMean) / avgMean : 0;
    double bd = avgMean > 0 ? fabs(mean[2] - avgMean) / avgMean : 0;
    double maxDeviation = dmax(dmax(rd, gd), bd);
    out->score[IRP_SCORE_COLORSHIFT] = dmin(maxDeviation * 2, 1.0);
  }
  free(grey); free(e); free(o); free(b);
  out->status = 0;
  orc_top_issues(out->score, out->issues);
  return 0;
}

/* PromptEnhancerService._identifyTopIssues + _determineSeverity, promptEnhancer.js:121-145: entries above 0.3,
 * Array.prototype.sort by descending confidence (stable: ties keep the key order of classifier.js:62-70), first 3 */
void orc_top_issues(const double score[IRP_NUM_SCORES], uint8_t issues[4]) {
  int idx[IRP_NUM_SCORES], n = 0;
  for (int k = 0; k < IRP_NUM_SCORES; k++)
    if (score[k] > 0.3) idx[n++] = k;
  for (int i = 1; i < n; i++) { /* insertion sort: stable */
    int v = idx[i], j = i - 1;
    while (j >= 0 && score[idx[j]] < score[v]) {
      idx[j + 1] = idx[j];
      j--;
    }
    idx[j + 1] = v;
  }
  if (n > 3) n = 3;
  for (int k = 0; k < 3; k++) {
    if (k >= n) {
      issues[k] = IRP_NO_ISSUE;
      continue;
    }
    double c = score[idx[k]];
    int sev = c >= 0.7 ? IRP_SEVERITY_HIGH : (c >= 0.5 ? IRP_SEVERITY_MEDIUM : IRP_SEVERITY_LOW);
    issues[k] = IRP_ISSUE(sev, idx[k]);
  }
  issues[3] = (uint8_t)n;
}

/* ------------------------------------------------------------------ */
/* P2: sharp .rotate() auto-orient, imagePreprocess.js:42              */
/* sharp src/common.cc EXIF map — SURVEY.md §8a row P2                 */
/* ------------------------------------------------------------------ */
void orc_orient_dims(int w, int h, int orientation, int *ow, int *oh) {
  if (orientation >= 5 && orientation <= 8) {
    *ow = h;
    *oh = w;
  } else {
    *ow = w;
    *oh = h;
  }
}

/* source coordinate of oriented pixel (x, y) */
static inline void orient_src(int orientation, int w, int h, int x, int y, int *sx, int *sy) {
  switch (orientation) {
    default: *sx = x; *sy = y; break;
    case 2: *sx = w - 1 - x; *sy = y; break;         /* mirror horizontal */
    case 3: *sx = w - 1 - x; *sy = h - 1 - y; break; /* rotate 180        */
    case 4: *sx = x; *sy = h - 1 - y; break;         /* mirror vertical   */
    case 5: *sx = y; *sy = x; break;                 /* transpose         */
    case 6: *sx = y; *sy = h - 1 - x; break;         /* rotate 90 CW      */
    case 7: *sx = w - 1 - y; *sy = h - 1 - x; break; /* transverse        */
    case 8: *sx = w - 1 - y; *sy = x; break;         /* rotate 270 CW     */
  }
}

int orc_orient(const uint8_t *px, int w, int h, int c, size_t pitch, int orientation, uint8_t *out) {
  int ow, oh;
  orc_orient_dims(w, h, orientation, &ow, &oh);
  for (int y = 0; y < oh; y++)
    for (int x = 0; x < ow; x++) {
      int sx, sy;
      orient_src(orientation, w, h, x, y, &sx, &sy);
      memcpy(out + ((size_t)y * ow + x) * c, px + (size_t)sy * pitch + (size_t)sx * c, c);
    }
  return 0;
}

/* ------------------------------------------------------------------ */
/* P3: target dims. imagePreprocess.js:7-22 (needsResize,              */
/* calculateResizeDimensions) then sharp ResolveShrink for             */
/* fit:'inside' + withoutEnlargement on the ROTATED image, output      */
/* size = round(in / shrink) as vips_reduce{h,v} compute it.           */
/* ------------------------------------------------------------------ */
static double js_round(double v) { return floor(v + 0.5); } /* Math.round */

static void fit_inside(int wo, int ho, int tw, int th, int *ow, int *oh, double *shrink) {
  double hs = (double)wo / tw, vs = (double)ho / th;
  double f = hs > vs ? hs : vs;        /* Canvas::MAX */
  if (f < 1.0) f = 1.0;                /* withoutEnlargement */
  if (f > wo) f = wo;
  if (f > ho) f = ho;
  *shrink = f;
  *ow = (int)round((double)wo / f);
  *oh = (int)round((double)ho / f);
  if (*ow < 1) *ow = 1;
  if (*oh < 1) *oh = 1;
}

int orc_preprocess_dims(int w, int h, int orientation, int *ow, int *oh, double *shrink) {
  int wo, ho;
  orc_orient_dims(w, h, orientation, &wo, &ho);
  if (w > IRP_MAX_DIMENSION || h > IRP_MAX_DIMENSION) {
    double scale = (double)IRP_MAX_DIMENSION / (w > h ? w : h);
    int tw = (int)js_round(w * scale), th = (int)js_round(h * scale);
    if (tw < 1) tw = 1; /* sharp rejects 0; keep >= 1 */
    if (th < 1) th = 1;
    fit_inside(wo, ho, tw, th, ow, oh, shrink);
  } else {
    *ow = wo;
    *oh = ho;
    *shrink = 1.0;
  }
  return 0;
}

int orc_fusion_dims(int w, int h, int orientation, int *ow, int *oh, int *offx, int *offy, double *shrink) {
  int wo, ho;
  orc_orient_dims(w, h, orientation, &wo, &ho);
  fit_inside(wo, ho, IRP_FUSION_CANVAS, IRP_FUSION_CANVAS, ow, oh, shrink);
  *offx = (IRP_FUSION_CANVAS - *ow) / 2;
  *offy = (IRP_FUSION_CANVAS - *oh) / 2;
  return 0;
}

/* ------------------------------------------------------------------ */
/* P3: vips_resize lanczos3 = reducev then reduceh                     */
/* libvips resample/reducev.cpp, reduceh.cpp, templates.h              */
/* (calculate_coefficients_lanczos, vips_vector_to_fixed_point)        */
/* ------------------------------------------------------------------ */
static int reduce_points(double shrink) { return 2 * (int)rint(IRP_LANCZOS_A * shrink) + 1; }

static void lanczos_mask(double *c, int n, double shrink, double x) {
  const double a = IRP_LANCZOS_A;
  const double half = x + n / 2 - 1; /* integer n/2: centre between taps n/2-1 and n/2 */
  double sum = 0;
  for (int i = 0; i < n; i++) {
    double xp = (i - half) / shrink;
    double l;
    if (xp == 0.0)
      l = 1.0;
    else if (xp < -a || xp > a)
      l = 0.0;
    else
      l = a * sin(M_PI * xp) * sin(M_PI * xp / a) / (M_PI * M_PI * xp * xp);
    c[i] = l;
    sum += l;
  }
  for (int i = 0; i < n; i++) c[i] /= sum;
}

/* vips_vector_to_fixed_point: rint each, bisect a scale so the ints sum to rint(Σ·scale) */
static void to_fixed_point(const double *in, int16_t *out, int n, int scale) {
  double fsum = 0;
  for (int i = 0; i < n; i++) fsum += in[i];
  int target = (int)rint(fsum * scale);
  double high = scale + (n + 1) / 2, low = scale - (n + 1) / 2, guess;
  int sum;
  do {
    guess = (high + low) / 2.0;
    for (int i = 0; i < n; i++) out[i] = (int16_t)rint(in[i] * guess);
    sum = 0;
    for (int i = 0; i < n; i++) sum += out[i];
    if (sum == target) break;
    if (sum < target) low = guess;
    if (sum > target) high = guess;
  } while (high - low > 0.01);
  if (sum != target) {
    int each_error = (target - sum) / n;
    int extra_error = (target - sum) % n;
    int direction = extra_error > 0 ? 1 : -1;
    int n_elements = abs(extra_error);
    for (int i = 0; i < n; i++) out[i] += each_error;
    for (int i = 0; i < n_elements; i++) out[i] += direction;
  }
}

int orc_reduce_plan(int in_size, int out_size, double shrink, int coef_mode, int *n_taps, int32_t *start,
                    int32_t *phase, int16_t *coefs /* [65][IRP_MAX_TAPS] */) {
  return orc_reduce_plan_m(in_size, out_size, shrink, coef_mode, 0, n_taps, start, phase, coefs);
}

/* reduce_mode 0: the 12-bit integer path of reducev / reduceh (C and highway kernels agree on it).
 * reduce_mode 1 (IRP_REDUCE_VECTOR_2_6): coefficients with 6 fractional bits, (sum + 32) >> 6 — the precision of
 * libvips' older orc vector path; stated here as 12-bit coefficients that are multiples of 64, which is the same
 * arithmetic ((sum 64 c6 p + 2048) >> 12 == (sum c6 p + 32) >> 6) and leaves every kernel untouched. */
int orc_reduce_plan_m(int in_size, int out_size, double shrink, int coef_mode, int reduce_mode, int *n_taps,
                      int32_t *start, int32_t *phase, int16_t *coefs /* [65][IRP_MAX_TAPS] */) {
  int n = reduce_points(shrink);
  if (n > IRP_MAX_TAPS) return -2;
  *n_taps = n;
  double mask[IRP_MAX_TAPS];
  for (int t = 0; t <= IRP_PHASES; t++) {
    lanczos_mask(mask, n, shrink, (double)t / IRP_PHASES);
    int16_t *ci = coefs + (size_t)t * IRP_MAX_TAPS;
    memset(ci, 0, sizeof(int16_t) * IRP_MAX_TAPS);
    if (coef_mode == 1 && reduce_mode != 1)
      for (int i = 0; i < n; i++) ci[i] = (int16_t)(mask[i] * (1 << IRP_INTERP_SHIFT)); /* C cast: truncate */
    else if (reduce_mode == 1) {
      to_fixed_point(mask, ci, n, 64);
      for (int i = 0; i < n; i++) ci[i] = (int16_t)(ci[i] * 64);
    } else
      to_fixed_point(mask, ci, n, 1 << IRP_INTERP_SHIFT);
  }
  /* keep the image centred when out*shrink != in (libvips "extra_pixels") */
  double extra = out_size * shrink - in_size;
  for (int o = 0; o < out_size; o++) {
    double c = (o + 0.5) * shrink - 0.5 - extra / 2.0;
    int p = (int)floor(c);
    int sy = (int)floor(c * IRP_PHASES * 2);
    int siy = sy & (IRP_PHASES * 2 - 1);
    phase[o] = (siy + 1) >> 1;
    start[o] = p - (n / 2 - 1);
  }
  (void)in_size;
  return 0;
}

static void reduce_axis(const uint8_t *in, int in_w, int in_h, int c, int out_size, int vertical,
                        const int32_t *start, const int32_t *phase, const int16_t *coefs, int n, uint8_t *out) {
  int ow = vertical ? in_w : out_size, oh = vertical ? out_size : in_h;
  int lim = vertical ? in_h : in_w;
  for (int y = 0; y < oh; y++)
    for (int x = 0; x < ow; x++) {
      int o = vertical ? y : x;
      const int16_t *ci = coefs + (size_t)phase[o] * IRP_MAX_TAPS;
      for (int ch = 0; ch < c; ch++) {
        int sum = 0;
        for (int i = 0; i < n; i++) {
          int s = clampi(start[o] + i, 0, lim - 1);
          size_t idx = vertical ? ((size_t)s * in_w + x) * c + ch : ((size_t)y * in_w + s) * c + ch;
          sum += ci[i] * in[idx];
        }
        sum = (sum + (1 << (IRP_INTERP_SHIFT - 1))) >> IRP_INTERP_SHIFT; /* unsigned_fixed_round */
        out[((size_t)y * ow + x) * c + ch] = (uint8_t)clampi(sum, 0, 255);
      }
    }
}

/* vips_resize's integer pre-shrink (resample/resize.c, gap = 2.0, the default sharp leaves alone): when an axis
 * shrinks by 4 or more, int_shrink = floor(in / target / gap) rows / columns are box-averaged first — vips_shrinkv
 * then vips_shrinkh with "ceil" (the image is edge-extended to a multiple of the factor), each
 * (sum + factor / 2) / factor in integers, u8 between the passes — and lanczos3 only does the remaining factor in
 * [2, 4).   */
int orc_box_factor(int in_size, int out_size) {
  int k = (int)floor((double)in_size / out_size / 2.0);
  return k < 1 ? 1 : k;
}

int orc_box_shrink(const uint8_t *in, int w, int h, int c, int kh, int kv, uint8_t *out /* ceil(w/kh) x ceil(h/kv) */) {
  const int ow = (w + kh - 1) / kh, oh = (h + kv - 1) / kv;
  uint8_t *tmp = (uint8_t *)malloc((size_t)w * oh * c);
  if (!tmp) return -4;
  for (int y = 0; y < oh; y++)
    for (size_t i = 0; i < (size_t)w * c; i++) {
      int sum = 0;
      for (int j = 0; j < kv; j++) sum += in[(size_t)clampi(y * kv + j, 0, h - 1) * w * c + i];
      tmp[(size_t)y * w * c + i] = (uint8_t)((sum + kv / 2) / kv);
    }
  for (int y = 0; y < oh; y++)
    for (int x = 0; x < ow; x++)
      for (int ch = 0; ch < c; ch++) {
        int sum = 0;
        for (int j = 0; j < kh; j++) sum += tmp[((size_t)y * w + clampi(x * kh + j, 0, w - 1)) * c + ch];
        out[((size_t)y * ow + x) * c + ch] = (uint8_t)((sum + kh / 2) / kh);
      }
  free(tmp);
  return 0;
}

/* oriented u8 image (tight) -> resized u8 image (tight); box pre-shrink if due, reducev, then reduceh */
static int resize_tight(const uint8_t *img, int wo, int ho, int c, int ow, int oh, double shrink, int coef_mode,
                        int reduce_mode, uint8_t *out) {
  if (ow == wo && oh == ho) {
    memcpy(out, img, (size_t)wo * ho * c);
    return 0;
  }
  int rc = 0;
  uint8_t *boxed = NULL;
  double vshrink = shrink, hshrink = shrink;
  const int kh = orc_box_factor(wo, ow), kv = orc_box_factor(ho, oh);
  if (kh > 1 || kv > 1) {
    const int bw = (wo + kh - 1) / kh, bh = (ho + kv - 1) / kv;
    boxed = (uint8_t *)malloc((size_t)bw * bh * c);
    if (!boxed) return -4;
    rc = orc_box_shrink(img, wo, ho, c, kh, kv, boxed);
    if (rc) { free(boxed); return rc; }
    img = boxed;
    wo = bw;
    ho = bh;
    hshrink = shrink / kh;
    vshrink = shrink / kv;
  }
  int16_t *coefs = (int16_t *)malloc(sizeof(int16_t) * (IRP_PHASES + 1) * IRP_MAX_TAPS);
  int32_t *start = (int32_t *)malloc(sizeof(int32_t) * (ow > oh ? ow : oh));
  int32_t *phase = (int32_t *)malloc(sizeof(int32_t) * (ow > oh ? ow : oh));
  uint8_t *mid = (uint8_t *)malloc((size_t)wo * oh * c);
  int n;
  const uint8_t *hsrc = img;
  if (oh != ho) {
    rc = orc_reduce_plan_m(ho, oh, vshrink, coef_mode, reduce_mode, &n, start, phase, coefs);
    if (rc) goto done;
    reduce_axis(img, wo, ho, c, oh, 1, start, phase, coefs, n, mid);
    hsrc = mid;
  }
  if (ow != wo) {
    rc = orc_reduce_plan_m(wo, ow, hshrink, coef_mode, reduce_mode, &n, start, phase, coefs);
    if (rc) goto done;
    reduce_axis(hsrc, wo, oh, c, ow, 0, start, phase, coefs, n, out);
  } else {
    memcpy(out, hsrc, (size_t)wo * oh * c);
  }
done:
  free(coefs); free(start); free(phase); free(mid); free(boxed);
  return rc;
}

/* P4 "normalisation" to the raw pixels the JPEG encoder would see: RGBA is
 * flattened on black (libvips flatten: p*a/255, integer), grey stays grey. */
static void normalise(const uint8_t *in, int w, int h, int c, uint8_t *out) {
  size_t n = (size_t)w * h;
  if (c == 4) {
    for (size_t i = 0; i < n; i++)
      for (int ch = 0; ch < 3; ch++) out[i * 3 + ch] = (uint8_t)((in[i * 4 + ch] * in[i * 4 + 3]) / 255);
  } else {
    memcpy(out, in, n * c);
  }
}

int orc_preprocess(const uint8_t *px, int w, int h, int c, size_t pitch, int orientation, int coef_mode,
                   uint8_t *out, int *ow, int *oh, int *oc) {
  return orc_preprocess_m(px, w, h, c, pitch, orientation, coef_mode, 0, out, ow, oh, oc);
}

int orc_preprocess_m(const uint8_t *px, int w, int h, int c, size_t pitch, int orientation, int coef_mode,
                     int reduce_mode, uint8_t *out, int *ow, int *oh, int *oc) {
  if (!px || !out || w <= 0 || h <= 0 || pitch < (size_t)w * c) return -1;
  if (c != 1 && c != 3 && c != 4) return -2;
  double shrink;
  int wo, ho;
  orc_orient_dims(w, h, orientation, &wo, &ho);
  orc_preprocess_dims(w, h, orientation, ow, oh, &shrink);
  *oc = c == 4 ? 3 : c;
  uint8_t *oriented = (uint8_t *)malloc((size_t)wo * ho * c);
  uint8_t *resized = (uint8_t *)malloc((size_t)*ow * *oh * c);
  if (!oriented || !resized) { free(oriented); free(resized); return -4; }
  orc_orient(px, w, h, c, pitch, orientation, oriented);
  int rc = resize_tight(oriented, wo, ho, c, *ow, *oh, shrink, coef_mode, reduce_mode, resized);
  if (!rc) normalise(resized, *ow, *oh, c, out);
  free(oriented); free(resized);
  return rc;
}

/* P5 (defined by this build): orient -> lanczos3 fit inside
 * 2048x2048 (no enlargement) -> centred on a black 2048x2048x3 canvas. */
int orc_fusion_canvas(const uint8_t *px, int w, int h, int c, size_t pitch, int orientation, int coef_mode,
                      uint8_t *canvas) {
  return orc_fusion_canvas_m(px, w, h, c, pitch, orientation, coef_mode, 0, canvas);
}

int orc_fusion_canvas_m(const uint8_t *px, int w, int h, int c, size_t pitch, int orientation, int coef_mode,
                        int reduce_mode, uint8_t *canvas) {
  if (!px || !canvas || w <= 0 || h <= 0 || pitch < (size_t)w * c) return -1;
  if (c != 1 && c != 3 && c != 4) return -2;
  int ow, oh, ox, oy, wo, ho;
  double shrink;
  orc_orient_dims(w, h, orientation, &wo, &ho);
  orc_fusion_dims(w, h, orientation, &ow, &oh, &ox, &oy, &shrink);
  uint8_t *oriented = (uint8_t *)malloc((size_t)wo * ho * c);
  uint8_t *resized = (uint8_t *)malloc((size_t)ow * oh * c);
  uint8_t *norm = (uint8_t *)malloc((size_t)ow * oh * 3);
  if (!oriented || !resized || !norm) { free(oriented); free(resized); free(norm); return -4; }
  orc_orient(px, w, h, c, pitch, orientation, oriented);
  int rc = resize_tight(oriented, wo, ho, c, ow, oh, shrink, coef_mode, reduce_mode, resized);
  if (!rc) {
    normalise(resized, ow, oh, c, norm);
    const int S = IRP_FUSION_CANVAS;
    memset(canvas, 0, (size_t)S * S * 3);
    for (int y = 0; y < oh; y++)
      for (int x = 0; x < ow; x++)
        for (int ch = 0; ch < 3; ch++)
          canvas[((size_t)(y + oy) * S + x + ox) * 3 + ch] =
              c == 1 ? norm[(size_t)y * ow + x] : norm[((size_t)y * ow + x) * 3 + ch];
  }
  free(oriented); free(resized); free(norm);
  return rc;
}

/* ------------------------------------------------------------------ */
/* batch driver for the CPU-baseline timing (bench.py): images are     */
/* independent, so worker threads pull image indices from a shared     */
/* counter (pthreads; no OpenMP runtime needed on the box)             */
/* ------------------------------------------------------------------ */
#include <pthread.h>

typedef struct {
  const irp_image_desc *imgs;
  int n, luma_mode, coef_mode;
  irp_result *results;
  uint8_t **outs;
  int next, rc;
  pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg) {
  batch_job *j = (batch_job *)arg;
  for (;;) {
    pthread_mutex_lock(&j->mu);
    int i = j->next++;
    pthread_mutex_unlock(&j->mu);
    if (i >= j->n) break;
    const irp_image_desc *d = &j->imgs[i];
    int rc = orc_classify(d->pixels, d->width, d->height, d->channels, d->pitch, d->is_jpeg, j->luma_mode,
                          &j->results[i]);
    if (!rc && j->outs && j->outs[i]) {
      int ow, oh, oc;
      rc = orc_preprocess(d->pixels, d->width, d->height, d->channels, d->pitch, d->exif_orientation,
                          j->coef_mode, j->outs[i], &ow, &oh, &oc);
    }
    if (rc) {
      pthread_mutex_lock(&j->mu);
      j->rc = rc;
      pthread_mutex_unlock(&j->mu);
    }
  }
  return NULL;
}

int orc_analyze_batch(const irp_image_desc *imgs, int n, int luma_mode, int coef_mode, irp_result *results,
                      uint8_t **outs, int threads) {
  batch_job j = {imgs, n, luma_mode, coef_mode, results, outs, 0, 0, PTHREAD_MUTEX_INITIALIZER};
  if (threads < 1) threads = 1;
  if (threads > n) threads = n;
  if (threads > 256) threads = 256;
  pthread_t tid[256];
  for (int t = 1; t < threads; t++) pthread_create(&tid[t], NULL, batch_worker, &j);
  batch_worker(&j);
  for (int t = 1; t < threads; t++) pthread_join(tid[t], NULL);
  return j.rc;
}
