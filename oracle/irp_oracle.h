/*
 * irp_oracle.h — CPU oracle (TEST INFRASTRUCTURE; parity unpinned, see irp_oracle.c).
 * Shares struct layouts with include/irp.h so tests compare field by field.
 */
#ifndef IRP_ORACLE_H_
#define IRP_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#include "../include/irp.h"

#ifdef __cplusplus
extern "C" {
#endif

void orc_tables(float v2y[256], int y2v[257]);
uint8_t orc_grey_rgb(int r, int g, int b, int luma_mode);
int orc_grey(const uint8_t *px, int w, int h, int c, size_t pitch, int luma_mode, uint8_t *out);
/* which: 0 = Lap8 (blur), 1 = Sharp9 (noise), 2 = Lap4 (scratch) */
int orc_stencil(const uint8_t *grey, int w, int h, int which, uint8_t *out);
int orc_blur1(const uint8_t *px, int w, int h, int c, size_t pitch, uint8_t *out);
/* the _m variants take the remaining switches of include/irp.h (IRP_BLUR_*, IRP_REDUCE_*); the plain ones use 0 */
int orc_blur1_m(const uint8_t *px, int w, int h, int c, size_t pitch, int blur_mode, uint8_t *out);
int orc_classify(const uint8_t *px, int w, int h, int c, size_t pitch, int is_jpeg, int luma_mode,
                 irp_result *out);
int orc_classify_m(const uint8_t *px, int w, int h, int c, size_t pitch, int is_jpeg, int luma_mode, int blur_mode,
                   irp_result *out);

void orc_top_issues(const double score[IRP_NUM_SCORES], uint8_t issues[4]);

void orc_orient_dims(int w, int h, int orientation, int *ow, int *oh);
int orc_orient(const uint8_t *px, int w, int h, int c, size_t pitch, int orientation, uint8_t *out);
int orc_preprocess_dims(int w, int h, int orientation, int *ow, int *oh, double *shrink);
int orc_fusion_dims(int w, int h, int orientation, int *ow, int *oh, int *offx, int *offy, double *shrink);
int orc_reduce_plan(int in_size, int out_size, double shrink, int coef_mode, int *n_taps, int32_t *start,
                    int32_t *phase, int16_t *coefs);
int orc_reduce_plan_m(int in_size, int out_size, double shrink, int coef_mode, int reduce_mode, int *n_taps,
                      int32_t *start, int32_t *phase, int16_t *coefs);
int orc_box_factor(int in_size, int out_size);
int orc_box_shrink(const uint8_t *in, int w, int h, int c, int kh, int kv, uint8_t *out);
int orc_preprocess_m(const uint8_t *px, int w, int h, int c, size_t pitch, int orientation, int coef_mode,
                     int reduce_mode, uint8_t *out, int *ow, int *oh, int *oc);
int orc_fusion_canvas_m(const uint8_t *px, int w, int h, int c, size_t pitch, int orientation, int coef_mode,
                        int reduce_mode, uint8_t *canvas);
int orc_preprocess(const uint8_t *px, int w, int h, int c, size_t pitch, int orientation, int coef_mode,
                   uint8_t *out, int *ow, int *oh, int *oc);
int orc_fusion_canvas(const uint8_t *px, int w, int h, int c, size_t pitch, int orientation, int coef_mode,
                      uint8_t *canvas);
int orc_analyze_batch(const irp_image_desc *imgs, int n, int luma_mode, int coef_mode, irp_result *results,
                      uint8_t **outs, int threads);

#ifdef __cplusplus
}
#endif
#endif
