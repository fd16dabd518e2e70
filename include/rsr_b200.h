/*
 * rsr_b200.h — C ABI of the B200-native rusty_sr forward path (librsr_b200.so).
 *
 * Drop-in boundary: the single call `graph.forward(1, vec![input], &params).remove(0)` at
 * reference src/main.rs:171, with the graph built by `sr_net(FACTOR, None)`
 * (src/network.rs:16-109, src/main.rs:138,146,149,152).  Everything the reference does on
 * either side of that call (clap CLI, .rsr decode, image decode/encode) stays host-side.
 *
 * Plain pointers and sizes only.  All functions return 0 on success and a negative RSR_E_*
 * code on failure (the reference panics instead: main.rs:134-138,162,164,175); the message is
 * available from rsr_last_error().  A context is bound to one GPU and is not thread-safe
 * (the reference's Graph is `&mut self`, main.rs:133,171).  There is no CPU fallback.
 */
#ifndef RSR_B200_H
#define RSR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rsr_ctx rsr_ctx;

enum {
    RSR_OK = 0,
    RSR_E_PARAM_COUNT = -1, /* params.len() != graph.num_params()            (main.rs:162) */
    RSR_E_BAD_ARG = -2,     /* null pointer, non-positive dims, bad channels/factor        */
    RSR_E_CUDA = -3,        /* CUDA runtime / driver error, no device, kernel fault        */
    RSR_E_OOM = -4,         /* device or host allocation failed                            */
    RSR_E_UNSUPPORTED = -5  /* valid request this build does not implement                 */
};

/* Flags for rsr_create. */
enum {
    RSR_FLAG_NONE = 0,
    /* Launch plan of the two middle levels of the DAG (network.rs:65-70).  Results agree to fp32 rounding.
     * FUSED_MID: conv5|6|7 and conv8|9 run as ONE launch that keeps l2 in shared memory and the running
     *            partial sums in tensor memory (5 launches per call, 768 B/pixel less HBM traffic);
     * SPLIT_MID: two launches with l2 and the partial sums stored to HBM in between (6 launches per call).
     * Neither flag: chosen per call from the image width (the fused launch works on 124-pixel strips, the split launches
     *            on 128-pixel strips, so e.g. 256-pixel-wide patches run the split plan).  Factor 4 always runs SPLIT_MID. */
    RSR_FLAG_FUSED_MID = 1,
    RSR_FLAG_SPLIT_MID = 2,
    /* Run every conv launch on single CTAs (conv_strip.cuh) even where the library default uses CTA pairs
     * (tcgen05 cta_group::2, conv_pair.cuh). */
    RSR_FLAG_NO_PAIRS = 4,
    /* Operand precision of the 32-channel convolutions.  Default: fp16 x fp16 plus two fp8 (e4m3) correction products per
     * term, 2 tensor-core MMAs per K chunk: max-abs error vs the fp64 reference <= 3e-4 in linear light on uniform-noise
     * input, <= 5e-5 on natural images (contract: 1e-3).  BF16X3: three bf16 products per term (3 MMAs per K chunk):
     * <= 8e-5 everywhere, ~1.3x slower. */
    RSR_FLAG_BF16X3 = 8
};

/* graph.num_params() of sr_net(factor, None): 130459 for factor 3 (main.rs:162; layout in
 * network.rs:27-76 op-registration order; the three `expand` convolutions have 3 * factor^2 output channels,
 * network.rs:37).  Factors 1..4 are supported (the reference is compiled with FACTOR = 3, main.rs:31); 0 otherwise. */
size_t rsr_num_params(int factor);

/* Builds the device-side network from the flat parameter vector the reference passes as
 * `&params` (decoded .rsr, main.rs:138-152): copies, splits them into the kernels' operand format and repacks the
 * weights once.  The caller keeps ownership of `params`.  `device` is a CUDA ordinal.
 * Returns NULL on failure (see rsr_last_error / rsr_last_error_code). */
rsr_ctx* rsr_create(const float* params, size_t n_params, int factor, int device, int flags);
void rsr_destroy(rsr_ctx* ctx);

/* upscale, f32 in / f32 out — the exact tensor contract of graph.forward (main.rs:168-171):
 * in  : host [h][w][3] f32, sRGB-encoded 0..1 (what img_to_data produces, main.rs:170)
 * out : host [factor*h][factor*w][3] f32, NOT clamped (clamping is data_to_img's job).
 * The host calls are synchronous.  On one large frame the result's way back over PCIe takes as long as the forward pass or
 * longer, so the library runs such a frame as up to four horizontal bands (7 halo rows recomputed, results identical bit for
 * bit) and sends each band's rows to the host while the next band computes; pinned `out` memory makes that overlap real. */
int rsr_upscale_f32(rsr_ctx* ctx, const float* in_hwc3, int w, int h, float* out_hwc3);

/* upscale with the reference's image<->tensor mapping fused on the device
 * (img_to_data main.rs:170: v = u8/255, alpha dropped; data_to_img(..).to_rgba() main.rs:175:
 * round(clamp(v,0,1)*255), alpha = 255).
 * in  : host [h][w][in_channels] u8, in_channels 3 (RGB) or 4 (RGBA)
 * out : host [factor*h][factor*w][4] u8 RGBA. */
int rsr_upscale_u8(rsr_ctx* ctx, const uint8_t* in, int w, int h, int in_channels, uint8_t* out_rgba);
/* The same without the constant alpha channel: out is [factor*h][factor*w][3] u8 RGB = data_to_img(..) (main.rs:175);
 * `.to_rgba()` is a host-side step in the reference, so a host that adds alpha itself saves a quarter of the
 * device->host bytes. */
int rsr_upscale_u8_rgb(rsr_ctx* ctx, const uint8_t* in, int w, int h, int in_channels, uint8_t* out_rgb);

/* Batches of n equally sized images, contiguous [n][h][w][c] (image-batch sharding unit). */
int rsr_upscale_batch_f32(rsr_ctx* ctx, const float* in, int n, int w, int h, float* out);
int rsr_upscale_batch_u8(rsr_ctx* ctx, const uint8_t* in, int n, int w, int h, int in_channels,
                         uint8_t* out_rgba);

/* Device-resident variants: `d_in` / `d_out` are device pointers on the context's GPU; the work
 * is enqueued on the context's stream and NOT synchronised (use rsr_synchronize).  Exactly one
 * of d_out_f32 / d_out_rgba may be non-null per call; in_channels is 3 or 4 for u8 input. */
int rsr_upscale_device_f32(rsr_ctx* ctx, const float* d_in, int n, int w, int h, float* d_out_f32,
                           uint8_t* d_out_rgba);
int rsr_upscale_device_u8(rsr_ctx* ctx, const uint8_t* d_in, int n, int w, int h, int in_channels,
                          float* d_out_f32, uint8_t* d_out_rgba);
int rsr_upscale_device_u8_rgb(rsr_ctx* ctx, const uint8_t* d_in, int n, int w, int h, int in_channels, uint8_t* d_out_rgb);
/* Device-resident band of a larger frame (see rsr_upscale_band_f32): d_in_rows holds rows_in input rows including
 * top_halo / bottom_halo halo rows; only the factor * (rows_in - top_halo - bottom_halo) output rows of the band are
 * computed by the last launch and written to d_out_rows (out_channels 3 = RGB, 4 = RGBA).  Enqueued, not synchronised. */
int rsr_upscale_band_device_f32(rsr_ctx* ctx, const float* d_in_rows, int w, int rows_in, int top_halo, int bottom_halo,
                                float* d_out_rows);
int rsr_upscale_band_device_u8(rsr_ctx* ctx, const uint8_t* d_in_rows, int w, int rows_in, int top_halo, int bottom_halo,
                               int in_channels, uint8_t* d_out_rows, int out_channels);
int rsr_synchronize(rsr_ctx* ctx);

/* Pipelined host entry points: submit returns at once with a ticket (>= 0, or a negative error);
 * the frame's host->device copy, forward pass and device->host copy are enqueued on three streams
 * so that the copies of neighbouring frames overlap the compute of the current one (3 frames in
 * flight).  `in` and `out_rgba` must stay valid — and should be pinned host memory for the copies to
 * be asynchronous — until rsr_wait(ticket) returns.  Frames complete in submission order. */
int rsr_submit_u8(rsr_ctx* ctx, const uint8_t* in, int w, int h, int in_channels, uint8_t* out_rgba);
int rsr_submit_u8_rgb(rsr_ctx* ctx, const uint8_t* in, int w, int h, int in_channels, uint8_t* out_rgb);
int rsr_submit_f32(rsr_ctx* ctx, const float* in_hwc3, int w, int h, float* out_hwc3);
int rsr_wait(rsr_ctx* ctx, int ticket);
int rsr_wait_all(rsr_ctx* ctx);

/* Spatial-band entry point for sharding one large frame across GPUs (scheme
 * "overlap-recompute"): `in` holds rows [band_y0 - top_halo, band_y1 + bottom_halo) of a frame,
 * where the halos are min(7, rows available) — 7 = receptive field of the net.  Writes the
 * factor*(band_y1-band_y0) output rows of the band to `out`.  Rows of `in` beyond the halo
 * are not needed; image edges are zero padded exactly as in the unsharded frame. */
int rsr_upscale_band_f32(rsr_ctx* ctx, const float* in_rows, int w, int rows_in, int top_halo,
                         int bottom_halo, float* out_rows);

/* Same for the u8 image form (RGB / RGBA rows in, RGBA rows out). */
int rsr_upscale_band_u8(rsr_ctx* ctx, const uint8_t* in_rows, int w, int rows_in, int top_halo, int bottom_halo,
                        int in_channels, uint8_t* out_rows_rgba);

/* Several GPUs from ONE host process — what a single-process host such as the reference CLI (main.rs:33-178) binds
 * to use a whole box (SURVEY.md 8(b) "multi-GPU variants", 8(e)).  One context per listed device (a device may be
 * listed more than once), one host thread per context inside each call, no collective:
 *   rsr_multi_upscale_batch_*: image i of the batch runs on context i mod G (independent units);
 *   rsr_multi_upscale_frame_*: ONE frame split into G horizontal bands with 7-row input halos (overlap-recompute);
 *                              the result is bit-identical to rsr_upscale_* of the whole frame on one GPU.
 * devices == NULL: all visible devices (at most n_devices of them if n_devices > 0).  Errors: as rsr_create /
 * rsr_upscale_*; the message names the failing device.  Not thread-safe per multi context. */
typedef struct rsr_multi rsr_multi;
rsr_multi* rsr_create_multi(const float* params, size_t n_params, int factor, const int* devices, int n_devices, int flags);
void rsr_destroy_multi(rsr_multi* m);
int rsr_multi_size(rsr_multi* m);               /* number of contexts G */
rsr_ctx* rsr_multi_ctx(rsr_multi* m, int i);    /* borrowed: context i (timing / plan queries) */
int rsr_multi_upscale_batch_f32(rsr_multi* m, const float* in, int n, int w, int h, float* out);
int rsr_multi_upscale_batch_u8(rsr_multi* m, const uint8_t* in, int n, int w, int h, int in_channels, uint8_t* out_rgba);
int rsr_multi_upscale_frame_f32(rsr_multi* m, const float* in_hwc3, int w, int h, float* out_hwc3);
int rsr_multi_upscale_frame_u8(rsr_multi* m, const uint8_t* in, int w, int h, int in_channels, uint8_t* out_rgba);

/* The two parameter-free modes of the reference CLI (no conv net, no rsr_ctx):
 *   rsr_bilinear_*   = bilinear_net(factor)   (network.rs:111-123; `-p bilinear`, main.rs:153-155):
 *                      sRGB -> linear, LinearInterp x factor (half-pixel centres, edge clamp), linear -> sRGB
 *   rsr_downsample_* = downsample_net(factor) (network.rs:125-138; `-d`, main.rs:139-141):
 *                      sRGB -> linear, factor x factor average pooling, linear -> sRGB;
 *                      output is ceil(h/factor) x ceil(w/factor), edge windows average the pixels they cover.
 * The transfer functions are the IEC 61966-2-1 sRGB curves.  alumina's own op sources are not available, so
 * these two modes are NOT pinned by any reference artefact.  f32 variants follow the tensor
 * contract of graph.forward ([h][w][3] in, unclamped out); u8 variants apply img_to_data / data_to_img. */
int rsr_bilinear_f32(int device, const float* in_hwc3, int w, int h, int factor, float* out_hwc3);
int rsr_bilinear_u8(int device, const uint8_t* in, int w, int h, int in_channels, int factor, uint8_t* out_rgba);
int rsr_downsample_f32(int device, const float* in_hwc3, int w, int h, int factor, float* out_hwc3);
int rsr_downsample_u8(int device, const uint8_t* in, int w, int h, int in_channels, int factor, uint8_t* out_rgba);

/* Generic "Same"-padded 3x3 convolution with the same kernel template as the sr_net launches (BASELINE config 5:
 * per-layer conv sweep).  x: host f32 [h][w][cin], weights: host f32 [cout][3][3][cin] (alumina Convolution
 * layout), out: host f32 [h][w][cout]; cin and cout must be multiples of 32.  The kernel part
 * (input already packed, device resident) is repeated `iters` times and its mean duration in ms is written to
 * *kernel_ms (may be NULL).  No context needed. */
int rsr_conv2d_3x3_f32(int device, const float* x, const float* weights, int w, int h, int cin, int cout,
                       float* out, int iters, float* kernel_ms);

/* The parameter-file codec of the reference: bytevec `<Vec<f32>>::decode::<u32>` (main.rs:138,146-152,194) and
 * `.encode::<u32>()` (main.rs:213): u32 LE count | count x u32 LE element sizes (all 4) | count x f32 LE.  Host only.
 * encode returns the bytes written (rsr_rsr_encoded_size(n)), decode the number of parameters in the blob (at most `cap`
 * of them are written; params may be NULL to query); negative RSR_E_BAD_ARG on a malformed blob / short buffer. */
size_t rsr_rsr_encoded_size(size_t n_params);
long long rsr_rsr_encode(const float* params, size_t n_params, uint8_t* out, size_t cap);
long long rsr_rsr_decode(const uint8_t* data, size_t len, float* params, size_t cap);

/* Use an externally created CUDA stream (cudaStream_t passed as void*) instead of the
 * context's own; NULL restores the internal stream. */
int rsr_set_stream(rsr_ctx* ctx, void* cuda_stream);

/* Per-kernel device times (ms, CUDA events on the launching stream) of the most recent call,
 * in launch order: conv0 (reads the image itself: img_to_data and the x-im2col happen on the way into shared memory),
 * conv1|2|3, conv5|6|7, conv8|9, conv10+expand.  Synchronises.
 * Returns the number of entries written (<= cap) or a negative error; 0 unless rsr_set_kernel_timing(ctx, 1)
 * was called before the forward pass.  Timing is off by default: the events between the launches would serialise them,
 * and the conv launches are chained with programmatic dependent launch (each sets itself up under its predecessor's tail). */
int rsr_set_kernel_timing(rsr_ctx* ctx, int enable);
int rsr_last_kernel_times(rsr_ctx* ctx, float* ms, int cap);
/* Names matching rsr_last_kernel_times entries. */
const char* rsr_kernel_name(int index);
/* Number of kernel launches issued by the most recent call. */
int rsr_last_launch_count(rsr_ctx* ctx);
/* Human-readable launch plan of this context (split / fused middle levels, which launches run on CTA pairs). */
const char* rsr_plan(rsr_ctx* ctx);

/* Test hooks (used by tests/ to localise a parity failure to one launch; not needed by a host
 * integration).  rsr_debug_create = rsr_create with an explicit CTA-pair launch mask (bit 2: conv1|2|3, bit 3: the fused
 * middle launch; negative = library default).  rsr_debug_stop_after(k): subsequent calls stop after launch index k
 * (k = 1 + the index of rsr_kernel_name: 1 = conv0 .. 5 = conv10; negative = run everything; outputs are then not written).  rsr_debug_read_node
 * copies one workspace buffer of the most recent call to the host as f32 [n*h*w][32]:
 * which = 0: activation buffer A (f, later l2), 1: activation buffer B (l1, later l3),
 *         2: partial buffer A (conv2(f), later expand), 3: partial buffer B (l3_conv partial),
 *         4: partial buffer C (expand partial of the fused middle launch; there l3 lives in buffer A),
 *         5, 6: debug dumps of the fused middle launch (rsr_debug_fused). */
rsr_ctx* rsr_debug_create(const float* params, size_t n_params, int factor, int device, int flags, int pair_mask);
int rsr_debug_stop_after(rsr_ctx* ctx, int launch_index);
int rsr_debug_read_node(rsr_ctx* ctx, int which, float* out, size_t n_pixels);
/* Per-CTA wait-cycle counters of the conv launches (which mbarrier each warp role waited on, in SM
 * cycles): enable, run a call, then read [grid][16] rows for launch index 1..5; returns the grid. */
/* Fused middle launch: flags bit 0 = skip the second-level MMAs, bit 1 = dump l2 (read_node which = 5) and the raw
 * accumulator columns conv6(l1)+conv8(l2) (which = 6) as f32 [n*h*w][32]; bits 0-2 select the single-CTA debug
 * kernel.  Bit 2 = wait-cycle counters of every warp role, read as [grid][32] rows (slot names: the FP_* enum of
 * csrc/conv_fused.cuh); bit 3 = the same counters from the plan's own kernel (single CTA or CTA pair), which
 * carries them only in a library built with -DRSR_FUSED_PROF=1 (all zero otherwise). */
int rsr_debug_fused(rsr_ctx* ctx, int flags);
int rsr_debug_read_fused_profile(rsr_ctx* ctx, long long* out, int cap_rows);
int rsr_debug_profile(rsr_ctx* ctx, int enable);
int rsr_debug_read_profile(rsr_ctx* ctx, int launch_index, long long* out, int cap_rows);

/* Thread-local description of the last failure in this thread ("" if none). */
const char* rsr_last_error(void);
int rsr_last_error_code(void);

/* Page-locked host memory for image buffers (cudaHostAlloc / cudaFreeHost, so that a host program does not have to link the
 * CUDA runtime itself).  Copies from / into such buffers run at PCIe speed and overlap the computation (rsr_upscale_* on a large
 * frame, rsr_submit_*): 1080p u8 -> RGBA takes 1.8 ms per synchronous call with pinned buffers, 4.5 ms with pageable ones.
 * rsr_host_alloc returns NULL on failure (see rsr_last_error). */
void* rsr_host_alloc(size_t bytes);
void rsr_host_free(void* p);

/* Build identification, e.g. "rsr_b200 0.2 sm_100a ...". */
const char* rsr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RSR_B200_H */
