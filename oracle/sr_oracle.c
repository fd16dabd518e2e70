/*
 * sr_oracle.c — CPU restatement of the rusty_sr super-resolution forward pass.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (rusty_sr_b200/, include/) never links or calls anything in oracle/.
 *
 * What it restates (reference = millardjn/rusty_sr @ /root/reference):
 *   - graph topology, kernel sizes, padding, parameter order ... src/network.rs:16-76,104-109
 *   - which blob / factor ........................................ src/main.rs:26-31
 *   - u8 <-> f32 image mapping .................................. src/main.rs:168-175
 * The arithmetic itself lives in the un-vendored crate alumina ^0.1.1 (Cargo.toml:11; no
 * Cargo.lock, no Rust toolchain here), so the op semantics (kernel memory order
 * [Cout][Kh][Kw][Cin], cross-correlation without flip, zero "Same" padding, BeLU formula,
 * Expand channel order, half-pixel bilinear) follow SURVEY.md Appendix A, which was pinned
 * empirically against the reference's committed output images.
 *
 * Parity pinning: the reference has no tests.  This oracle is pinned by the two image pairs
 * the reference commits as its own output (docs/cartoon_lr.png -> docs/cartoon_rsa.png with
 * anime.rsr, bit-level; docs/butterfly_lr.png -> docs/butterfly_rs.png with imagenet.rsr,
 * 55.7 dB) and the fp64 known-answer values of SURVEY.md App. B.4 — see
 * tests/test_oracle_golden.py.
 *
 * Two precisions: sr_oracle_forward_f32 (what the reference computes in: fp32) and
 * sr_oracle_forward_f64 (ground truth for tolerances).  OpenMP over image rows.
 */
#define _POSIX_C_SOURCE 200112L
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SR_MAXC 64 /* max channels of any node (3*f*f for f<=4 is 48) */

/* ---- parameter layout: order in which ops are added to the graph, network.rs:27-76 ---- */
typedef struct {
    int factor;
    size_t conv0, f_bias, f_activ, expand_bias;
    size_t l1_bias, l2_bias, l3_bias, l1_activ, l2_activ, l3_activ;
    size_t conv1, conv2, conv3, conv5, conv6, conv7, conv8, conv9, conv10;
    size_t total;
} sr_layout;

static void sr_make_layout(int factor, sr_layout *L) {
    size_t o = 0;
    const size_t ce = (size_t)3 * factor * factor; /* network.rs:37 CHANNELS*factor*factor */
    L->factor = factor;
    L->conv0 = o;       o += 32 * 5 * 5 * 3;   /* network.rs:33 */
    L->f_bias = o;      o += 32;               /* :34 */
    L->f_activ = o;     o += 32;               /* :35 */
    L->expand_bias = o; o += ce;               /* :38 */
    L->l1_bias = o;     o += 32;               /* :50 */
    L->l2_bias = o;     o += 32;               /* :51 */
    L->l3_bias = o;     o += 32;               /* :52 */
    L->l1_activ = o;    o += 32;               /* :54 */
    L->l2_activ = o;    o += 32;               /* :55 */
    L->l3_activ = o;    o += 32;               /* :56 */
    L->conv1 = o;       o += 32 * 25 * 32;     /* :60 */
    L->conv2 = o;       o += 32 * 25 * 32;     /* :61 */
    L->conv3 = o;       o += 32 * 25 * 32;     /* :62 */
    L->conv5 = o;       o += 32 * 9 * 32;      /* :65 */
    L->conv6 = o;       o += 32 * 9 * 32;      /* :66 */
    L->conv7 = o;       o += ce * 9 * 32;      /* :67 */
    L->conv8 = o;       o += 32 * 9 * 32;      /* :69 */
    L->conv9 = o;       o += ce * 9 * 32;      /* :70 */
    L->conv10 = o;      o += ce * 9 * 32;      /* :72 */
    L->total = o;
}

/* main.rs:162 — graph.num_params(); 130459 for factor 3 */
size_t sr_oracle_num_params(int factor) {
    sr_layout L;
    sr_make_layout(factor, &L);
    return L.total;
}

int sr_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void sr_oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- .rsr decode: bytevec `<Vec<f32>>::decode::<u32>` (main.rs:138,146,149,152) --------
 * u32 LE N; N x u32 LE element byte sizes (all 4); N x f32 LE.  Returns element count, or
 * -1 on malformed input.  out may be NULL to query the count. */
long sr_oracle_decode_rsr(const uint8_t *bytes, size_t len, float *out, size_t out_cap) {
    if (len < 4) return -1;
    uint32_t n = (uint32_t)bytes[0] | ((uint32_t)bytes[1] << 8) | ((uint32_t)bytes[2] << 16) |
                 ((uint32_t)bytes[3] << 24);
    if ((size_t)4 + (size_t)8 * n != len) return -1;
    for (uint32_t i = 0; i < n; i++) {
        const uint8_t *p = bytes + 4 + 4 * (size_t)i;
        uint32_t sz = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) |
                      ((uint32_t)p[3] << 24);
        if (sz != 4) return -1;
    }
    if (out) {
        if (out_cap < n) return -1;
        memcpy(out, bytes + 4 + 4 * (size_t)n, 4 * (size_t)n); /* little-endian host */
    }
    return (long)n;
}

/* The generic body is instantiated twice (float / double) through the preprocessor. */
#define REAL float
#define SR_LANES 16
typedef float sr_v16f __attribute__((vector_size(64), aligned(64)));
#define SR_VEC sr_v16f
#define SUF(x) x##_f32
#define SQRT sqrtf
#define FLOOR floorf
#include "sr_oracle_body.inc"
#undef REAL
#undef SUF
#undef SQRT
#undef FLOOR
#undef SR_LANES
#undef SR_VEC

#define REAL double
#define SR_LANES 8
typedef double sr_v8d __attribute__((vector_size(64), aligned(64)));
#define SR_VEC sr_v8d
#define SUF(x) x##_f64
#define SQRT sqrt
#define FLOOR floor
#include "sr_oracle_body.inc"
#undef REAL
#undef SUF
#undef SQRT
#undef FLOOR

/* ---- image <-> tensor (alumina img_to_data / data_to_img; call sites main.rs:170,175) --- */

/* u8 RGB or RGBA -> f32 [H][W][3], v = u8/255, alpha discarded. */
void sr_oracle_img_to_data(const uint8_t *img, int w, int h, int channels, float *out) {
    for (long i = 0; i < (long)w * h; i++)
        for (int c = 0; c < 3; c++) out[i * 3 + c] = (float)img[i * channels + c] / 255.0f;
}

/* f32 [H][W][3] -> u8 RGBA: round(clamp(v,0,1)*255), alpha = 255 (`.to_rgba()`). */
void sr_oracle_data_to_img(const float *data, int w, int h, uint8_t *rgba) {
    for (long i = 0; i < (long)w * h; i++) {
        for (int c = 0; c < 3; c++) {
            float v = data[i * 3 + c];
            v = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
            rgba[i * 4 + c] = (uint8_t)(v * 255.0f + 0.5f);
        }
        rgba[i * 4 + 3] = 255;
    }
}

/* ---- the two parameter-free graphs of the reference CLI (test oracle for rsr_bilinear_* / rsr_downsample_*)
 * bilinear_net  (network.rs:111-123): SrgbToLinear -> LinearInterp(factor) -> LinearToSrgb
 * downsample_net(network.rs:125-138): SrgbToLinear -> Pooling(factor, average) -> LinearToSrgb
 * Transfer functions: IEC 61966-2-1.  alumina's op sources are unavailable and the reference commits no
 * output of these two modes, so this part of the oracle is NOT pinned ("parity unpinned"); it documents
 * the semantics the GPU path implements.  Edge windows of the pooling average the pixels they cover. */
static double sr_srgb_to_linear(double v) { return v <= 0.04045 ? v / 12.92 : pow((v + 0.055) / 1.055, 2.4); }
static double sr_linear_to_srgb(double v) { return v <= 0.0031308 ? v * 12.92 : 1.055 * pow(v, 1.0 / 2.4) - 0.055; }

void sr_oracle_bilinear_f64(const float *in, int W, int H, int F, double *out) {
    const int OW = W * F, OH = H * F;
#pragma omp parallel for schedule(static)
    for (int Y = 0; Y < OH; Y++) {
        const double sy = (Y + 0.5) / F - 0.5, fy = floor(sy), ty = sy - fy;
        int y0 = (int)fy, y1 = y0 + 1;
        y0 = y0 < 0 ? 0 : (y0 > H - 1 ? H - 1 : y0);
        y1 = y1 < 0 ? 0 : (y1 > H - 1 ? H - 1 : y1);
        for (int X = 0; X < OW; X++) {
            const double sx = (X + 0.5) / F - 0.5, fx = floor(sx), tx = sx - fx;
            int x0 = (int)fx, x1 = x0 + 1;
            x0 = x0 < 0 ? 0 : (x0 > W - 1 ? W - 1 : x0);
            x1 = x1 < 0 ? 0 : (x1 > W - 1 ? W - 1 : x1);
            for (int c = 0; c < 3; c++) {
                const double v00 = sr_srgb_to_linear(in[((size_t)y0 * W + x0) * 3 + c]);
                const double v01 = sr_srgb_to_linear(in[((size_t)y0 * W + x1) * 3 + c]);
                const double v10 = sr_srgb_to_linear(in[((size_t)y1 * W + x0) * 3 + c]);
                const double v11 = sr_srgb_to_linear(in[((size_t)y1 * W + x1) * 3 + c]);
                const double top = v00 + (v01 - v00) * tx, bot = v10 + (v11 - v10) * tx;
                out[((size_t)Y * OW + X) * 3 + c] = sr_linear_to_srgb(top + (bot - top) * ty);
            }
        }
    }
}

void sr_oracle_downsample_f64(const float *in, int W, int H, int F, double *out) {
    const int OW = (W + F - 1) / F, OH = (H + F - 1) / F;
#pragma omp parallel for schedule(static)
    for (int Y = 0; Y < OH; Y++)
        for (int X = 0; X < OW; X++)
            for (int c = 0; c < 3; c++) {
                double acc = 0;
                int cnt = 0;
                for (int dy = 0; dy < F; dy++)
                    for (int dx = 0; dx < F; dx++) {
                        const int y = Y * F + dy, x = X * F + dx;
                        if (y < H && x < W) { acc += sr_srgb_to_linear(in[((size_t)y * W + x) * 3 + c]); cnt++; }
                    }
                out[((size_t)Y * OW + X) * 3 + c] = sr_linear_to_srgb(acc / cnt);
            }
}

/* Generic zero-padded KxK convolution (the op behind every `Convolution::new(.., Padding::Same, ..)`,
 * network.rs:33,60-72) exposed for the per-layer sweep tests of rsr_conv2d_3x3_f32: in f32 [H][W][Ci],
 * weights f32 [Co][K][K][Ci], out f64 [H][W][Co].  Returns 0 or -1. */
int sr_oracle_conv2d_f64(const float *in, int W, int H, int Ci, const float *w_okki, int Co, int K, double *out) {
    const size_t n = (size_t)W * H;
    double *x = (double *)malloc(n * Ci * sizeof(double));
    if (!x) return -1;
    for (size_t i = 0; i < n * Ci; i++) x[i] = in[i];
    memset(out, 0, n * Co * sizeof(double));
    int rc = 0;
    /* conv_acc handles up to 96 output channels per call (register tile); do wider layers in slices */
    for (int co0 = 0; co0 < Co && !rc; co0 += 32) {
        const int nco = Co - co0 < 32 ? Co - co0 : 32;
        double *tmp = (double *)calloc(n * nco, sizeof(double));
        if (!tmp) { rc = -1; break; }
        rc = conv_acc_f64(x, H, W, Ci, w_okki + (size_t)co0 * K * K * Ci, nco, K, tmp);
        for (size_t i = 0; i < n && !rc; i++)
            for (int o = 0; o < nco; o++) out[i * Co + co0 + o] = tmp[i * nco + o];
        free(tmp);
    }
    free(x);
    return rc;
}
