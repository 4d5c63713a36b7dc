/*
 * jpeg_oracle.c — TEST INFRASTRUCTURE ONLY (like irp_oracle.c): a CPU restatement of baseline JPEG
 * decoding as libjpeg-turbo performs it with its default settings (JDCT_ISLOW, fancy upsampling), the
 * decoder inside libvips / sharp that runs in front of every pipeline of the hot path
 * (reference: server-node/src/services/classifier.js:51-52,107,135,199,296 `sharp(imageBuffer)`,
 *  server-node/src/middleware/imagePreprocess.js:40-42; accepted uploads uploadValidation.js:7).
 * SURVEY.md §8f rank 1 ("JPEG bitstream decode on device") is the next row of the hot-path table; this
 * file is its oracle.  Unlike the rest of oracle/, it IS pinned: tests/test_jpeg_oracle.py compares
 * it bit for bit with Pillow's decoder, which is libjpeg-turbo itself (the same library, the same
 * defaults as libvips).
 *
 * Follows, by algorithm (nothing is copied; libjpeg-turbo is not in /root/reference):
 *   jdmarker.c  marker parsing (SOF0, DQT, DHT, DRI, SOS, RSTn)
 *   jdhuff.c    Huffman entropy decoding, HUFF_EXTEND, restart handling
 *   jidctint.c  jpeg_idct_islow (CONST_BITS 13, PASS1_BITS 2)
 *   jdsample.c  h2v1 / h2v2 / h1v2 fancy (triangle) upsampling, replicated edges
 *   jdcolor.c   YCbCr -> RGB with the 16-bit fixed-point tables
 *   jdphuff.c   progressive scans: DC / AC first passes and refinements, EOB runs, correction bits
 * Scope: 8-bit Huffman JPEG, baseline sequential (SOF0 / SOF1, one interleaved scan or one scan per component) and
 * progressive (SOF2: spectral selection + successive approximation, what mozjpeg / `progressive: true` writes and
 * imagePreprocess.js:57-61 therefore hands on), 1 or 3 components, sampling factors 1 or 2.  A complete progressive
 * file needs no block smoothing (jdcoefct.c smooths only while AC precision is still missing), so its pixels are the
 * IDCT of the final coefficients.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { JO_OK = 0, JO_ERR_FORMAT = -1, JO_ERR_UNSUPPORTED = -2, JO_ERR_NOMEM = -4 };

static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

typedef struct {
  int maxcode[18], valptr[17], mincode[17];
  uint8_t bits[17], vals[256];
  int present;
} HuffTab;

typedef struct {
  int id, h, v, tq, td, ta;
  int bw, bh;        /* blocks per row / block rows, MCU padded */
  int dw, dh;        /* downsampled (real) width / height in samples (at the decode scale) */
  int S, hx, vx;     /* DCT_scaled_size (8 = full size) and the upsampler's expansion factors */
  int fw, fh;        /* downsampled size at FULL scale: what a scan of this component alone walks */
  int16_t* coef;     /* [bh][bw][64] natural order, DC already integrated */
  uint8_t* plane;    /* [bh*8][bw*8] */
} Comp;

typedef struct {
  int w, h, ncomp, hmax, vmax, mcux, mcuy, restart;
  uint16_t q[4][64]; /* natural order */
  int qpresent[4];
  HuffTab dc[4], ac[4];
  Comp comp[3];
  const uint8_t* scan;
  size_t scan_len;
  int adobe_transform;  /* -1 none */
  int progressive;      /* SOF2 */
  int out_w, out_h, fancy; /* output size and upsampler kind at the decode scale (w, h, 1 at full size) */
  int multi;            /* more than the one interleaved full-band scan: decode_multi() walks the scans */
  size_t sos_pos;       /* index of the first SOS marker's 0xFF */
} Jpeg;

static int build_huff(HuffTab* t) {
  int code = 0, k = 0;
  for (int l = 1; l <= 16; l++) {
    t->valptr[l] = k;
    t->mincode[l] = code;
    code += t->bits[l];
    k += t->bits[l];
    t->maxcode[l] = t->bits[l] ? code - 1 : -1;
    code <<= 1;
  }
  t->maxcode[17] = 0xFFFFF;
  t->present = 1;
  return k <= 256 ? JO_OK : JO_ERR_FORMAT;
}

static int parse_dqt(Jpeg* J, const uint8_t* s, size_t sl) {
  size_t o = 0;
  while (o < sl) {
    const int pq = s[o] >> 4, tq = s[o] & 15;
    o++;
    if (tq > 3) return JO_ERR_FORMAT;
    for (int i = 0; i < 64; i++) {
      int v;
      if (pq) {
        if (o + 2 > sl) return JO_ERR_FORMAT;
        v = (s[o] << 8) | s[o + 1];
        o += 2;
      } else {
        if (o + 1 > sl) return JO_ERR_FORMAT;
        v = s[o++];
      }
      J->q[tq][kZigzag[i]] = (uint16_t)v;
    }
    J->qpresent[tq] = 1;
  }
  return JO_OK;
}
static int parse_dht(Jpeg* J, const uint8_t* s, size_t sl) {
  size_t o = 0;
  while (o < sl) {
    if (o + 17 > sl) return JO_ERR_FORMAT;
    const int tc = s[o] >> 4, th = s[o] & 15;
    if (tc > 1 || th > 3) return JO_ERR_FORMAT;
    HuffTab* t = tc ? &J->ac[th] : &J->dc[th];
    int cnt = 0;
    t->bits[0] = 0;
    for (int l = 1; l <= 16; l++) {
      t->bits[l] = s[o + l];
      cnt += t->bits[l];
    }
    o += 17;
    if (cnt > 256 || o + cnt > sl) return JO_ERR_FORMAT;
    memcpy(t->vals, s + o, cnt);
    o += cnt;
    if (build_huff(t)) return JO_ERR_FORMAT;
  }
  return JO_OK;
}

static int parse(const uint8_t* d, size_t n, Jpeg* J) {
  memset(J, 0, sizeof *J);
  J->adobe_transform = -1;
  if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) return JO_ERR_FORMAT;
  size_t p = 2;
  for (;;) {
    if (p + 4 > n) return JO_ERR_FORMAT;
    if (d[p] != 0xFF) return JO_ERR_FORMAT;
    while (p < n && d[p] == 0xFF) p++;
    if (p >= n) return JO_ERR_FORMAT;
    const int m = d[p++];
    if (m == 0xD9) return JO_ERR_FORMAT; /* EOI before SOS */
    if (p + 2 > n) return JO_ERR_FORMAT;
    const size_t len = ((size_t)d[p] << 8) | d[p + 1];
    if (len < 2 || p + len > n) return JO_ERR_FORMAT;
    const uint8_t* s = d + p + 2;
    const size_t sl = len - 2;
    if (m == 0xDB) { /* DQT */
      const int rc = parse_dqt(J, s, sl);
      if (rc) return rc;
    } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) { /* SOF0 / SOF1 / SOF2 */
      J->progressive = m == 0xC2;
      if (sl < 6 || s[0] != 8) return JO_ERR_UNSUPPORTED;
      J->h = (s[1] << 8) | s[2];
      J->w = (s[3] << 8) | s[4];
      J->ncomp = s[5];
      if ((J->ncomp != 1 && J->ncomp != 3) || sl < (size_t)(6 + 3 * J->ncomp) || J->w == 0 || J->h == 0) return JO_ERR_UNSUPPORTED;
      for (int i = 0; i < J->ncomp; i++) {
        J->comp[i].id = s[6 + 3 * i];
        J->comp[i].h = s[7 + 3 * i] >> 4;
        J->comp[i].v = s[7 + 3 * i] & 15;
        J->comp[i].tq = s[8 + 3 * i];
        if (J->comp[i].h < 1 || J->comp[i].h > 2 || J->comp[i].v < 1 || J->comp[i].v > 2 || J->comp[i].tq > 3) return JO_ERR_UNSUPPORTED;
      }
    } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
      return JO_ERR_UNSUPPORTED; /* lossless, arithmetic, hierarchical */
    } else if (m == 0xC4) { /* DHT */
      const int rc = parse_dht(J, s, sl);
      if (rc) return rc;
    } else if (m == 0xDD) { /* DRI */
      if (sl < 2) return JO_ERR_FORMAT;
      J->restart = (s[0] << 8) | s[1];
    } else if (m == 0xEE && sl >= 12 && !memcmp(s, "Adobe", 5)) {
      J->adobe_transform = s[11];
    } else if (m == 0xDA) { /* SOS */
      if (!J->w || sl < 1) return JO_ERR_FORMAT;
      const int ns = s[0];
      if (ns < 1 || ns > J->ncomp || sl < (size_t)(1 + 2 * ns + 3)) return JO_ERR_FORMAT;
      J->sos_pos = p - 2;
      J->multi = J->progressive || ns != J->ncomp || s[1 + 2 * ns] != 0 || s[2 + 2 * ns] != 63 || s[3 + 2 * ns] != 0;
      if (!J->multi) {
        for (int i = 0; i < ns; i++) {
          int ci = -1;
          for (int k = 0; k < J->ncomp; k++)
            if (J->comp[k].id == s[1 + 2 * i]) ci = k;
          if (ci != i) return JO_ERR_UNSUPPORTED;
          J->comp[ci].td = s[2 + 2 * i] >> 4;
          J->comp[ci].ta = s[2 + 2 * i] & 15;
        }
        J->scan = d + p + len;
        J->scan_len = n - (p + len);
      }
      break;
    }
    p += len;
  }
  J->hmax = J->vmax = 1;
  for (int i = 0; i < J->ncomp; i++) {
    if (J->comp[i].h > J->hmax) J->hmax = J->comp[i].h;
    if (J->comp[i].v > J->vmax) J->vmax = J->comp[i].v;
  }
  if (J->ncomp == 1) J->comp[0].h = J->comp[0].v = J->hmax = J->vmax = 1; /* single-component scans are never interleaved */
  J->mcux = (J->w + 8 * J->hmax - 1) / (8 * J->hmax);
  J->mcuy = (J->h + 8 * J->vmax - 1) / (8 * J->vmax);
  for (int i = 0; i < J->ncomp; i++) {
    Comp* c = &J->comp[i];
    c->bw = J->mcux * c->h;
    c->bh = J->mcuy * c->v;
    c->dw = (J->w * c->h + J->hmax - 1) / J->hmax;
    c->dh = (J->h * c->v + J->vmax - 1) / J->vmax;
    c->fw = c->dw;
    c->fh = c->dh;
    c->S = 8;
    c->hx = J->hmax / c->h;
    c->vx = J->vmax / c->v;
    if (!J->qpresent[c->tq]) return JO_ERR_FORMAT;
    if (!J->multi && (!J->dc[c->td].present || !J->ac[c->ta].present)) return JO_ERR_FORMAT;
  }
  J->out_w = J->w;
  J->out_h = J->h;
  J->fancy = 1;
  return JO_OK;
}

/* jdmaster.c, jpeg_core_output_dimensions + jdsample.c jinit_upsampler for scale 1/denom (denom 2, 4, 8 — libvips'
 * shrink-on-load): the output is ceil(size / denom); every component gets the LARGEST IDCT size that still needs no
 * more than the frame's upsampling (a 2x subsampled chroma component is scaled up by its IDCT instead of the
 * upsampler), and fancy upsampling is only used while the smallest IDCT is larger than 1x1. */
static void set_scale(Jpeg* J, int denom) {
  const int minS = 8 / denom;
  J->out_w = (J->w + denom - 1) / denom;
  J->out_h = (J->h + denom - 1) / denom;
  J->fancy = minS > 1;
  for (int i = 0; i < J->ncomp; i++) {
    Comp* c = &J->comp[i];
    int ssize = minS;
    while (ssize < 8 && (J->hmax * minS) % (c->h * ssize * 2) == 0 && (J->vmax * minS) % (c->v * ssize * 2) == 0) ssize *= 2;
    c->S = ssize;
    c->dw = (J->w * c->h * ssize + J->hmax * 8 - 1) / (J->hmax * 8);
    c->dh = (J->h * c->v * ssize + J->vmax * 8 - 1) / (J->vmax * 8);
    c->hx = J->hmax / ((c->h * ssize) / minS);
    c->vx = J->vmax / ((c->v * ssize) / minS);
  }
}

/* ---- entropy decoding (jdhuff.c) ---- */
typedef struct {
  const uint8_t* d;
  size_t n, p;
  uint32_t acc;
  int cnt;
  int marker; /* pending marker byte, 0 none */
} Bits;

static void fill(Bits* b) {
  while (b->cnt <= 24) {
    int byte = 0;
    if (!b->marker && b->p < b->n) {
      byte = b->d[b->p++];
      if (byte == 0xFF) {
        int nx = b->p < b->n ? b->d[b->p] : 0xD9;
        while (nx == 0xFF && b->p + 1 < b->n) { /* fill bytes */
          b->p++;
          nx = b->d[b->p];
        }
        if (nx == 0) {
          b->p++;
        } else {
          b->marker = nx;
          b->p++;
          byte = 0;
        }
      }
    }
    b->acc |= (uint32_t)byte << (24 - b->cnt);
    b->cnt += 8;
  }
}
static int getbits(Bits* b, int n) {
  if (!n) return 0;
  if (b->cnt < n) fill(b);
  const int v = (int)(b->acc >> (32 - n));
  b->acc <<= n;
  b->cnt -= n;
  return v;
}
static int decode_sym(Bits* b, const HuffTab* t) {
  int code = getbits(b, 1), l = 1;
  while (l <= 16 && code > t->maxcode[l]) {
    code = (code << 1) | getbits(b, 1);
    l++;
  }
  if (l > 16) return 0;
  return t->vals[t->valptr[l] + code - t->mincode[l]];
}
static int extend(int x, int s) { return x < (1 << (s - 1)) ? x + (int)((~0u) << s) + 1 : x; }

static int decode_scan(Jpeg* J) {
  Bits b = {J->scan, J->scan_len, 0, 0, 0, 0};
  int pred[3] = {0, 0, 0};
  int togo = J->restart;
  for (int my = 0; my < J->mcuy; my++)
    for (int mx = 0; mx < J->mcux; mx++) {
      if (J->restart && togo == 0) {
        /* byte-align, expect RSTn */
        b.acc = 0;
        b.cnt = 0;
        if (!b.marker) { /* the marker has not been reached by the bit reader yet: scan for it */
          while (b.p + 1 < b.n && !(b.d[b.p] == 0xFF && b.d[b.p + 1] >= 0xD0 && b.d[b.p + 1] <= 0xD7)) b.p++;
          if (b.p + 1 >= b.n) return JO_ERR_FORMAT;
          b.p += 2;
        }
        b.marker = 0;
        pred[0] = pred[1] = pred[2] = 0;
        togo = J->restart;
      }
      for (int ci = 0; ci < J->ncomp; ci++) {
        Comp* c = &J->comp[ci];
        for (int by = 0; by < c->v; by++)
          for (int bx = 0; bx < c->h; bx++) {
            int16_t* blk = c->coef + ((size_t)(my * c->v + by) * c->bw + (mx * c->h + bx)) * 64;
            int s = decode_sym(&b, &J->dc[c->td]);
            if (s) pred[ci] += extend(getbits(&b, s), s);
            blk[0] = (int16_t)pred[ci];
            for (int k = 1; k < 64; k++) {
              const int rs = decode_sym(&b, &J->ac[c->ta]);
              const int r = rs >> 4;
              s = rs & 15;
              if (s) {
                k += r;
                if (k > 63) break;
                blk[kZigzag[k]] = (int16_t)extend(getbits(&b, s), s);
              } else {
                if (r != 15) break;
                k += 15;
              }
            }
          }
      }
      togo--;
    }
  return JO_OK;
}

/* ---- every scan of a multi-scan file (jdphuff.c; sequential scans of single components go the same way) ---- */
typedef struct {
  int ns, ci[3], td[3], ta[3], ss, se, ah, al;
} ScanHdr;

static int restart_sync(Bits* b) { /* byte-align, step over the RSTn marker */
  b->acc = 0;
  b->cnt = 0;
  if (!b->marker) {
    while (b->p + 1 < b->n && !(b->d[b->p] == 0xFF && b->d[b->p + 1] >= 0xD0 && b->d[b->p + 1] <= 0xD7)) b->p++;
    if (b->p + 1 >= b->n) return JO_ERR_FORMAT;
    b->p += 2;
  }
  b->marker = 0;
  return JO_OK;
}

/* one block of an AC refinement scan (decode_mcu_AC_refine) */
static void ac_refine_block(Bits* b, const HuffTab* t, int16_t* blk, const ScanHdr* h, int* eobrun) {
  const int p1 = 1 << h->al, m1 = -(1 << h->al);
  int k = h->ss;
  if (*eobrun == 0) {
    for (; k <= h->se; k++) {
      const int rs = decode_sym(b, t);
      int r = rs >> 4, s = rs & 15;
      if (s) {
        s = getbits(b, 1) ? p1 : m1; /* a newly nonzero coefficient is +-1 at this bit position */
      } else if (r != 15) {
        *eobrun = 1 << r;
        if (r) *eobrun += getbits(b, r);
        break; /* the rest of the band only takes correction bits */
      }
      /* step over r still-zero coefficients; every nonzero one met on the way takes a correction bit */
      do {
        int16_t* c = blk + kZigzag[k];
        if (*c) {
          if (getbits(b, 1) && !(*c & p1)) *c = (int16_t)(*c >= 0 ? *c + p1 : *c + m1);
        } else if (--r < 0) {
          break;
        }
        k++;
      } while (k <= h->se);
      if (s && k <= h->se) blk[kZigzag[k]] = (int16_t)s;
    }
  }
  if (*eobrun > 0) {
    for (; k <= h->se; k++) {
      int16_t* c = blk + kZigzag[k];
      if (*c && getbits(b, 1) && !(*c & p1)) *c = (int16_t)(*c >= 0 ? *c + p1 : *c + m1);
    }
    (*eobrun)--;
  }
}

static int decode_one_scan(Jpeg* J, const ScanHdr* h, const uint8_t* data, size_t len) {
  Bits b = {data, len, 0, 0, 0, 0};
  int pred[3] = {0, 0, 0}, eobrun = 0, togo = J->restart;
  /* a scan of one component walks that component's own blocks (ceil(samples / 8)), not the MCU-padded grid */
  const int single = h->ns == 1;
  const Comp* c0 = &J->comp[h->ci[0]];
  const int mx_n = single ? (c0->fw + 7) / 8 : J->mcux, my_n = single ? (c0->fh + 7) / 8 : J->mcuy;
  for (int my = 0; my < my_n; my++)
    for (int mx = 0; mx < mx_n; mx++) {
      if (J->restart && togo == 0) {
        if (restart_sync(&b)) return JO_ERR_FORMAT;
        pred[0] = pred[1] = pred[2] = 0;
        eobrun = 0;
        togo = J->restart;
      }
      for (int i = 0; i < h->ns; i++) {
        Comp* c = &J->comp[h->ci[i]];
        const int nh = single ? 1 : c->h, nv = single ? 1 : c->v;
        for (int by = 0; by < nv; by++)
          for (int bx = 0; bx < nh; bx++) {
            int16_t* blk = c->coef + ((size_t)(my * nv + by) * c->bw + (mx * nh + bx)) * 64;
            if (h->ss == 0) {
              if (h->ah == 0) { /* DC first pass (also the DC of a sequential scan) */
                const int s = decode_sym(&b, &J->dc[h->td[i]]);
                if (s) pred[i] += extend(getbits(&b, s), s);
                blk[0] = (int16_t)(pred[i] * (1 << h->al));
              } else if (getbits(&b, 1)) { /* DC refinement: one more bit */
                blk[0] |= (int16_t)(1 << h->al);
              }
              if (!J->progressive) { /* sequential: the AC coefficients follow in the same scan */
                for (int k = 1; k < 64; k++) {
                  const int rs = decode_sym(&b, &J->ac[h->ta[i]]);
                  const int r = rs >> 4, s = rs & 15;
                  if (s) {
                    k += r;
                    if (k > 63) break;
                    blk[kZigzag[k]] = (int16_t)extend(getbits(&b, s), s);
                  } else {
                    if (r != 15) break;
                    k += 15;
                  }
                }
              }
            } else if (h->ah == 0) { /* AC first pass */
              if (eobrun > 0) {
                eobrun--;
              } else {
                for (int k = h->ss; k <= h->se; k++) {
                  const int rs = decode_sym(&b, &J->ac[h->ta[i]]);
                  const int r = rs >> 4, s = rs & 15;
                  if (s) {
                    k += r;
                    if (k > 63) break;
                    blk[kZigzag[k]] = (int16_t)(extend(getbits(&b, s), s) * (1 << h->al));
                  } else if (r == 15) {
                    k += 15;
                  } else {
                    eobrun = 1 << r;
                    if (r) eobrun += getbits(&b, r);
                    eobrun--; /* this block is the first of the run */
                    break;
                  }
                }
              }
            } else {
              ac_refine_block(&b, &J->ac[h->ta[i]], blk, h, &eobrun);
            }
          }
      }
      togo--;
    }
  return JO_OK;
}

static int decode_multi(Jpeg* J, const uint8_t* d, size_t n) {
  size_t p = J->sos_pos;
  for (;;) {
    if (p + 2 > n) return JO_OK; /* no EOI: what was decoded stands */
    if (d[p] != 0xFF) return JO_ERR_FORMAT;
    while (p < n && d[p] == 0xFF) p++;
    if (p >= n) return JO_OK;
    const int m = d[p++];
    if (m == 0xD9) return JO_OK;
    if (p + 2 > n) return JO_ERR_FORMAT;
    const size_t len = ((size_t)d[p] << 8) | d[p + 1];
    if (len < 2 || p + len > n) return JO_ERR_FORMAT;
    const uint8_t* s = d + p + 2;
    const size_t sl = len - 2;
    int rc = JO_OK;
    if (m == 0xC4) rc = parse_dht(J, s, sl);
    else if (m == 0xDB) rc = parse_dqt(J, s, sl);
    else if (m == 0xDD) {
      if (sl < 2) return JO_ERR_FORMAT;
      J->restart = (s[0] << 8) | s[1];
    } else if (m == 0xDA) {
      ScanHdr h;
      if (sl < 1) return JO_ERR_FORMAT;
      h.ns = s[0];
      if (h.ns < 1 || h.ns > J->ncomp || sl < (size_t)(1 + 2 * h.ns + 3)) return JO_ERR_FORMAT;
      for (int i = 0; i < h.ns; i++) {
        h.ci[i] = -1;
        for (int k = 0; k < J->ncomp; k++)
          if (J->comp[k].id == s[1 + 2 * i]) h.ci[i] = k;
        if (h.ci[i] < 0) return JO_ERR_FORMAT;
        h.td[i] = s[2 + 2 * i] >> 4;
        h.ta[i] = s[2 + 2 * i] & 15;
        if (h.td[i] > 3 || h.ta[i] > 3) return JO_ERR_FORMAT;
      }
      h.ss = s[1 + 2 * h.ns];
      h.se = s[2 + 2 * h.ns];
      h.ah = s[3 + 2 * h.ns] >> 4;
      h.al = s[3 + 2 * h.ns] & 15;
      if (J->progressive) {
        if (h.ss > h.se || h.se > 63 || (h.ss == 0 && h.se != 0) || (h.ss > 0 && h.ns != 1) || h.al > 13) return JO_ERR_FORMAT;
      } else if (h.ss != 0 || h.se != 63 || h.ah || h.al) {
        return JO_ERR_FORMAT;
      }
      for (int i = 0; i < h.ns; i++) {
        if (h.ss == 0 && h.ah == 0 && !J->dc[h.td[i]].present) return JO_ERR_FORMAT;
        if ((h.ss > 0 || !J->progressive) && !J->ac[h.ta[i]].present) return JO_ERR_FORMAT;
      }
      /* the entropy-coded segment runs to the next marker that is neither a stuffed zero nor RSTn */
      size_t e = p + len;
      const size_t d0 = e;
      while (e + 1 < n && !(d[e] == 0xFF && d[e + 1] != 0x00 && d[e + 1] != 0xFF && !(d[e + 1] >= 0xD0 && d[e + 1] <= 0xD7))) e++;
      if (e + 1 >= n) e = n;
      if ((rc = decode_one_scan(J, &h, d + d0, e - d0))) return rc;
      p = e;
      continue;
    }
    if (rc) return rc;
    p += len;
  }
}

/* ---- jidctint.c ---- */
#define CONST_BITS 13
#define PASS1_BITS 2
#define DESCALE(x, n) (((x) + ((int32_t)1 << ((n)-1))) >> (n))
static uint8_t range_limit(int32_t x) {
  int v = x & 1023;
  if (v >= 512) v -= 1024;
  v += 128;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}
static void idct_islow(const int16_t* in, const uint16_t* q, uint8_t* out, int stride) {
  int32_t ws[64];
  for (int c = 0; c < 8; c++) {
    const int16_t* ip = in + c;
    const uint16_t* qp = q + c;
    int32_t* wp = ws + c;
    int32_t z2 = ip[16] * qp[16], z3 = ip[48] * qp[48];
    int32_t z1 = (z2 + z3) * 4433;
    int32_t tmp2 = z1 + z3 * (-15137), tmp3 = z1 + z2 * 6270;
    z2 = ip[0] * qp[0];
    z3 = ip[32] * qp[32];
    int32_t tmp0 = (z2 + z3) * (1 << CONST_BITS), tmp1 = (z2 - z3) * (1 << CONST_BITS);
    const int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = ip[56] * qp[56];
    tmp1 = ip[40] * qp[40];
    tmp2 = ip[24] * qp[24];
    tmp3 = ip[8] * qp[8];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    int32_t z4 = tmp1 + tmp3;
    const int32_t z5 = (z3 + z4) * 9633;
    tmp0 *= 2446;
    tmp1 *= 16819;
    tmp2 *= 25172;
    tmp3 *= 12299;
    z1 *= -7373;
    z2 *= -20995;
    z3 *= -16069;
    z4 *= -3196;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    wp[0] = DESCALE(tmp10 + tmp3, CONST_BITS - PASS1_BITS);
    wp[56] = DESCALE(tmp10 - tmp3, CONST_BITS - PASS1_BITS);
    wp[8] = DESCALE(tmp11 + tmp2, CONST_BITS - PASS1_BITS);
    wp[48] = DESCALE(tmp11 - tmp2, CONST_BITS - PASS1_BITS);
    wp[16] = DESCALE(tmp12 + tmp1, CONST_BITS - PASS1_BITS);
    wp[40] = DESCALE(tmp12 - tmp1, CONST_BITS - PASS1_BITS);
    wp[24] = DESCALE(tmp13 + tmp0, CONST_BITS - PASS1_BITS);
    wp[32] = DESCALE(tmp13 - tmp0, CONST_BITS - PASS1_BITS);
  }
  for (int r = 0; r < 8; r++) {
    const int32_t* wp = ws + 8 * r;
    uint8_t* op = out + (size_t)r * stride;
    int32_t z2 = wp[2], z3 = wp[6];
    int32_t z1 = (z2 + z3) * 4433;
    int32_t tmp2 = z1 + z3 * (-15137), tmp3 = z1 + z2 * 6270;
    int32_t tmp0 = (wp[0] + wp[4]) * (1 << CONST_BITS), tmp1 = (wp[0] - wp[4]) * (1 << CONST_BITS);
    const int32_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = wp[7];
    tmp1 = wp[5];
    tmp2 = wp[3];
    tmp3 = wp[1];
    z1 = tmp0 + tmp3;
    z2 = tmp1 + tmp2;
    z3 = tmp0 + tmp2;
    int32_t z4 = tmp1 + tmp3;
    const int32_t z5 = (z3 + z4) * 9633;
    tmp0 *= 2446;
    tmp1 *= 16819;
    tmp2 *= 25172;
    tmp3 *= 12299;
    z1 *= -7373;
    z2 *= -20995;
    z3 *= -16069;
    z4 *= -3196;
    z3 += z5;
    z4 += z5;
    tmp0 += z1 + z3;
    tmp1 += z2 + z4;
    tmp2 += z2 + z3;
    tmp3 += z1 + z4;
    const int sh = CONST_BITS + PASS1_BITS + 3;
    op[0] = range_limit(DESCALE(tmp10 + tmp3, sh));
    op[7] = range_limit(DESCALE(tmp10 - tmp3, sh));
    op[1] = range_limit(DESCALE(tmp11 + tmp2, sh));
    op[6] = range_limit(DESCALE(tmp11 - tmp2, sh));
    op[2] = range_limit(DESCALE(tmp12 + tmp1, sh));
    op[5] = range_limit(DESCALE(tmp12 - tmp1, sh));
    op[3] = range_limit(DESCALE(tmp13 + tmp0, sh));
    op[4] = range_limit(DESCALE(tmp13 - tmp0, sh));
  }
}

/* ---- jidctred.c: reduced-size inverse DCTs (4x4, 2x2, 1x1 output from one 8x8 coefficient block) ---- */
static void idct_4x4(const int16_t* in, const uint16_t* q, uint8_t* out, int stride) {
  int32_t ws[8 * 4];
  for (int c = 0; c < 8; c++) {
    if (c == 4) continue; /* the second pass never reads column 4 */
    const int16_t* ip = in + c;
    const uint16_t* qp = q + c;
    int32_t tmp0 = (int32_t)(ip[0] * qp[0]) * (1 << (CONST_BITS + 1));
    int32_t z2 = ip[16] * qp[16], z3 = ip[48] * qp[48];
    int32_t tmp2 = z2 * 15137 + z3 * (-6270);
    const int32_t tmp10 = tmp0 + tmp2, tmp12 = tmp0 - tmp2;
    const int32_t z1 = ip[56] * qp[56];
    z2 = ip[40] * qp[40];
    z3 = ip[24] * qp[24];
    const int32_t z4 = ip[8] * qp[8];
    tmp0 = z1 * (-1730) + z2 * 11893 + z3 * (-17799) + z4 * 8697;
    tmp2 = z1 * (-4176) + z2 * (-4926) + z3 * 7373 + z4 * 20995;
    ws[c + 0] = DESCALE(tmp10 + tmp2, CONST_BITS - PASS1_BITS + 1);
    ws[c + 24] = DESCALE(tmp10 - tmp2, CONST_BITS - PASS1_BITS + 1);
    ws[c + 8] = DESCALE(tmp12 + tmp0, CONST_BITS - PASS1_BITS + 1);
    ws[c + 16] = DESCALE(tmp12 - tmp0, CONST_BITS - PASS1_BITS + 1);
  }
  for (int r = 0; r < 4; r++) {
    const int32_t* w = ws + 8 * r;
    int32_t tmp0 = w[0] * (1 << (CONST_BITS + 1));
    int32_t tmp2 = w[2] * 15137 + w[6] * (-6270);
    const int32_t tmp10 = tmp0 + tmp2, tmp12 = tmp0 - tmp2;
    const int32_t z1 = w[7], z2 = w[5], z3 = w[3], z4 = w[1];
    tmp0 = z1 * (-1730) + z2 * 11893 + z3 * (-17799) + z4 * 8697;
    tmp2 = z1 * (-4176) + z2 * (-4926) + z3 * 7373 + z4 * 20995;
    uint8_t* o = out + (size_t)r * stride;
    o[0] = range_limit(DESCALE(tmp10 + tmp2, CONST_BITS + PASS1_BITS + 3 + 1));
    o[3] = range_limit(DESCALE(tmp10 - tmp2, CONST_BITS + PASS1_BITS + 3 + 1));
    o[1] = range_limit(DESCALE(tmp12 + tmp0, CONST_BITS + PASS1_BITS + 3 + 1));
    o[2] = range_limit(DESCALE(tmp12 - tmp0, CONST_BITS + PASS1_BITS + 3 + 1));
  }
}
static void idct_2x2(const int16_t* in, const uint16_t* q, uint8_t* out, int stride) {
  int32_t ws[8 * 2];
  for (int c = 0; c < 8; c++) {
    if (c == 2 || c == 4 || c == 6) continue; /* never read by the second pass */
    const int16_t* ip = in + c;
    const uint16_t* qp = q + c;
    const int32_t tmp10 = (int32_t)(ip[0] * qp[0]) * (1 << (CONST_BITS + 2));
    const int32_t tmp0 = (ip[56] * qp[56]) * (-5906) + (ip[40] * qp[40]) * 6967 + (ip[24] * qp[24]) * (-10426) + (ip[8] * qp[8]) * 29692;
    ws[c] = DESCALE(tmp10 + tmp0, CONST_BITS - PASS1_BITS + 2);
    ws[c + 8] = DESCALE(tmp10 - tmp0, CONST_BITS - PASS1_BITS + 2);
  }
  for (int r = 0; r < 2; r++) {
    const int32_t* w = ws + 8 * r;
    const int32_t tmp10 = w[0] * (1 << (CONST_BITS + 2));
    const int32_t tmp0 = w[7] * (-5906) + w[5] * 6967 + w[3] * (-10426) + w[1] * 29692;
    out[(size_t)r * stride] = range_limit(DESCALE(tmp10 + tmp0, CONST_BITS + PASS1_BITS + 3 + 2));
    out[(size_t)r * stride + 1] = range_limit(DESCALE(tmp10 - tmp0, CONST_BITS + PASS1_BITS + 3 + 2));
  }
}
static void idct_1x1(const int16_t* in, const uint16_t* q, uint8_t* out) { out[0] = range_limit(DESCALE((int32_t)(in[0] * q[0]), 3)); }

/* ---- jdsample.c: fancy upsampling of one component to full resolution (at least w x h) ---- */
static uint8_t* upsample(const Jpeg* J, const Comp* c, int* out_stride) {
  const int hx = c->hx, vx = c->vx;
  const int pw = c->bw * c->S; /* padded plane width */
  const int ow = c->dw * hx, oh = c->dh * vx;
  uint8_t* o = (uint8_t*)malloc((size_t)(ow + 2) * (oh + 2));
  if (!o) return NULL;
  *out_stride = ow;
#define ROW(r) (c->plane + (size_t)((r) < 0 ? 0 : ((r) >= c->dh ? c->dh - 1 : (r))) * pw)
  if (hx == 1 && vx == 1) {
    for (int y = 0; y < oh; y++) memcpy(o + (size_t)y * ow, ROW(y), ow);
  } else if ((hx == 2 && (c->dw <= 2 || !J->fancy)) || (hx == 1 && vx == 2 && !J->fancy)) {
    /* jinit_upsampler: the fancy h2v1 / h2v2 routines only when downsampled_width > 2; else h2v1_upsample /
     * h2v2_upsample, plain replication (vertically too) */
    for (int y = 0; y < oh; y++) {
      const uint8_t* in = ROW(vx == 2 ? y >> 1 : y);
      for (int x = 0; x < ow; x++) o[(size_t)y * ow + x] = in[hx == 2 ? x >> 1 : x];
    }
  } else if (hx == 2 && vx == 1) {
    for (int y = 0; y < oh; y++) {
      const uint8_t* in = ROW(y);
      uint8_t* op = o + (size_t)y * ow;
      const int n = c->dw;
      if (n == 1) {
        op[0] = in[0];
        op[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2); /* libjpeg reads the padded column here */
        continue;
      }
      op[0] = in[0];
      op[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
      for (int x = 1; x < n - 1; x++) {
        op[2 * x] = (uint8_t)((in[x] * 3 + in[x - 1] + 1) >> 2);
        op[2 * x + 1] = (uint8_t)((in[x] * 3 + in[x + 1] + 2) >> 2);
      }
      op[2 * (n - 1)] = (uint8_t)((in[n - 1] * 3 + in[n - 2] + 1) >> 2);
      op[2 * (n - 1) + 1] = in[n - 1];
    }
  } else if (hx == 2 && vx == 2) {
    for (int r = 0; r < c->dh; r++)
      for (int v = 0; v < 2; v++) {
        const uint8_t* in0 = ROW(r);
        const uint8_t* in1 = v == 0 ? ROW(r - 1) : ROW(r + 1);
        uint8_t* op = o + (size_t)(2 * r + v) * ow;
        const int n = c->dw;
        int thiscol = in0[0] * 3 + in1[0], nextcol = in0[1] * 3 + in1[1], lastcol;
        op[0] = (uint8_t)((thiscol * 4 + 8) >> 4);
        op[1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
        if (n == 1) continue;
        lastcol = thiscol;
        thiscol = nextcol;
        for (int x = 1; x < n - 1; x++) {
          nextcol = in0[x + 1] * 3 + in1[x + 1];
          op[2 * x] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
          op[2 * x + 1] = (uint8_t)((thiscol * 3 + nextcol + 7) >> 4);
          lastcol = thiscol;
          thiscol = nextcol;
        }
        op[2 * (n - 1)] = (uint8_t)((thiscol * 3 + lastcol + 8) >> 4);
        op[2 * (n - 1) + 1] = (uint8_t)((thiscol * 4 + 7) >> 4);
      }
  } else if (hx == 1 && vx == 2) {
    for (int r = 0; r < c->dh; r++)
      for (int v = 0; v < 2; v++) {
        const uint8_t* in0 = ROW(r);
        const uint8_t* in1 = v == 0 ? ROW(r - 1) : ROW(r + 1);
        uint8_t* op = o + (size_t)(2 * r + v) * ow;
        const int bias = v == 0 ? 1 : 2;
        for (int x = 0; x < ow; x++) op[x] = (uint8_t)((in0[x] * 3 + in1[x] + bias) >> 2);
      }
  } else {
    free(o);
    return NULL;
  }
#undef ROW
  return o;
}

static void free_all(Jpeg* J) {
  for (int i = 0; i < 3; i++) {
    free(J->comp[i].coef);
    free(J->comp[i].plane);
  }
}

int irp_jpeg_info(const uint8_t* data, size_t len, int* w, int* h, int* ncomp, int* restart, int sampling[6]) {
  Jpeg J;
  const int rc = parse(data, len, &J);
  if (rc) return rc;
  *w = J.w;
  *h = J.h;
  *ncomp = J.ncomp;
  *restart = J.restart;
  for (int i = 0; i < 3; i++) {
    sampling[2 * i] = i < J.ncomp ? J.comp[i].h : 0;
    sampling[2 * i + 1] = i < J.ncomp ? J.comp[i].v : 0;
  }
  return JO_OK;
}

/* out: w*h*3 (RGB) for 3 components, w*h for 1.  coef_out (optional): the quantised coefficients of
 * component `coef_comp`, [block rows][blocks per row][64] natural order (for the device stages' tests). */
static int decode_at_scale(const uint8_t* data, size_t len, int denom, uint8_t* out, int16_t* coef_out, int coef_comp);
int irp_jpeg_decode(const uint8_t* data, size_t len, uint8_t* out, int16_t* coef_out, int coef_comp) {
  return decode_at_scale(data, len, 1, out, coef_out, coef_comp);
}
/* libjpeg's scale_num / scale_denom = 1 / denom (2, 4 or 8): out holds ceil(h / denom) x ceil(w / denom) pixels */
int irp_jpeg_decode_scaled(const uint8_t* data, size_t len, int denom, uint8_t* out) {
  if (denom != 1 && denom != 2 && denom != 4 && denom != 8) return JO_ERR_UNSUPPORTED;
  return decode_at_scale(data, len, denom, out, NULL, -1);
}
static int decode_at_scale(const uint8_t* data, size_t len, int denom, uint8_t* out, int16_t* coef_out, int coef_comp) {
  Jpeg J;
  int rc = parse(data, len, &J);
  if (rc) return rc;
  if (denom > 1) set_scale(&J, denom);
  for (int i = 0; i < J.ncomp; i++) {
    Comp* c = &J.comp[i];
    c->coef = (int16_t*)calloc((size_t)c->bw * c->bh * 64, sizeof(int16_t));
    c->plane = (uint8_t*)malloc((size_t)c->bw * c->bh * 64);
    if (!c->coef || !c->plane) {
      free_all(&J);
      return JO_ERR_NOMEM;
    }
  }
  if ((rc = J.multi ? decode_multi(&J, data, len) : decode_scan(&J))) {
    free_all(&J);
    return rc;
  }
  if (coef_out && coef_comp >= 0 && coef_comp < J.ncomp)
    memcpy(coef_out, J.comp[coef_comp].coef, (size_t)J.comp[coef_comp].bw * J.comp[coef_comp].bh * 64 * sizeof(int16_t));
  for (int i = 0; i < J.ncomp; i++) {
    Comp* c = &J.comp[i];
    const int S = c->S, pw = c->bw * S;
    for (int by = 0; by < c->bh; by++)
      for (int bx = 0; bx < c->bw; bx++) {
        const int16_t* blk = c->coef + ((size_t)by * c->bw + bx) * 64;
        uint8_t* o = c->plane + ((size_t)by * S * c->bw + bx) * S;
        if (S == 8) idct_islow(blk, J.q[c->tq], o, pw);
        else if (S == 4) idct_4x4(blk, J.q[c->tq], o, pw);
        else if (S == 2) idct_2x2(blk, J.q[c->tq], o, pw);
        else idct_1x1(blk, J.q[c->tq], o);
      }
  }
  if (!out) {
    free_all(&J);
    return JO_OK;
  }
  if (J.ncomp == 1) {
    for (int y = 0; y < J.out_h; y++) memcpy(out + (size_t)y * J.out_w, J.comp[0].plane + (size_t)y * J.comp[0].bw * J.comp[0].S, J.out_w);
    free_all(&J);
    return JO_OK;
  }
  uint8_t* up[3] = {0, 0, 0};
  int st[3];
  for (int i = 0; i < 3; i++) {
    up[i] = upsample(&J, &J.comp[i], &st[i]);
    if (!up[i]) {
      for (int k = 0; k < 3; k++) free(up[k]);
      free_all(&J);
      return JO_ERR_UNSUPPORTED;
    }
  }
  const int rgb_passthrough = J.adobe_transform == 0;
  for (int y = 0; y < J.out_h; y++)
    for (int x = 0; x < J.out_w; x++) {
      const int Y = up[0][(size_t)y * st[0] + x], cb = up[1][(size_t)y * st[1] + x], cr = up[2][(size_t)y * st[2] + x];
      uint8_t* o = out + ((size_t)y * J.out_w + x) * 3;
      if (rgb_passthrough) {
        o[0] = (uint8_t)Y;
        o[1] = (uint8_t)cb;
        o[2] = (uint8_t)cr;
        continue;
      }
      /* jdcolor.c build_ycc_rgb_table, SCALEBITS 16 */
      const int xb = cb - 128, xr = cr - 128;
      const int r = Y + (int)((91881 * (int64_t)xr + 32768) >> 16);
      const int g = Y + (int)(((-22554) * (int64_t)xb + 32768 + (-46802) * (int64_t)xr) >> 16);
      const int bl = Y + (int)((116130 * (int64_t)xb + 32768) >> 16);
      o[0] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
      o[1] = (uint8_t)(g < 0 ? 0 : (g > 255 ? 255 : g));
      o[2] = (uint8_t)(bl < 0 ? 0 : (bl > 255 ? 255 : bl));
    }
  for (int k = 0; k < 3; k++) free(up[k]);
  free_all(&J);
  return JO_OK;
}
