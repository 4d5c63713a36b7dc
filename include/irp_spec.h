/*
 * irp_spec.h — the hard-coded constants of the reference hot path, in one
 * place, shared by the CUDA kernels, the host library and the CPU oracle
 * (SURVEY.md §5 "Config/flags").  Paths cite the reference repo.
 */
#ifndef IRP_SPEC_H_
#define IRP_SPEC_H_

/* server-node/src/services/classifier.js */
#define IRP_BLUR_VAR_DIVISOR 1000.0   /* :119  edgeVariance / 1000            */
#define IRP_NOISE_STD_DIVISOR 50.0    /* :146  noiseLevel / 50                */
#define IRP_LOWLIGHT_KNEE 0.3         /* :163  normalizedBrightness < 0.3     */
#define IRP_COMPRESSION_DIVISOR 500.0 /* :303  varianceDelta / 500            */
#define IRP_SCRATCH_THRESHOLD 200     /* :318  threshold = 200                */
#define IRP_SCRATCH_STRIDE 4          /* :320-321  y += 4, x += 4             */
#define IRP_SCRATCH_DIVISOR 1000.0    /* :336  total / 1000                   */
#define IRP_CONTRAST_DIVISOR 64.0     /* :285  avgStdev / 64                  */

/* libvips gaussblur(sigma=1, min_ampl=0.2, precision=integer): gaussmat gives
 * taps [12,20,12], scale 44; convi rounds with (sum + scale/2) / scale.
 * (12*(l+r) + 20*c + 22) / 44 == (3*(l+r) + 5*c + 5) / 11 for all u8 inputs
 * (tests/test_oracle.py::test_blur_div11_identity checks every value). */
#define IRP_GAUSS_EDGE 12
#define IRP_GAUSS_CENTRE 20
#define IRP_GAUSS_SCALE 44

/* libvips reducev/reduceh fixed point */
#define IRP_INTERP_SHIFT 12  /* VIPS_INTERPOLATE_SHIFT */
#define IRP_PHASES 64        /* VIPS_TRANSFORM_SCALE   */
#define IRP_LANCZOS_A 3
#define IRP_MAX_TAPS 25      /* 2*rint(3*shrink)+1 with shrink < 4 */

/* additive diagnostic (north_star "8x8 JPEG blockiness"): a grey step counts
 * when |g(a) - g(b)| > this, for neighbours a|b straddling an 8-px boundary */
#define IRP_BLOCK_EDGE_THRESHOLD 12

#endif /* IRP_SPEC_H_ */
