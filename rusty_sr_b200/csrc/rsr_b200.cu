// rsr_b200.cu — C-ABI library (include/rsr_b200.h) over the sm_100a kernels.
//
// Host side of the drop-in boundary `graph.forward(1, vec![input], &params)` (reference
// src/main.rs:171): parameter-vector layout of sr_net (src/network.rs:27-76), weight split +
// repack, workspace management, TMA tensor maps, the launch sequence and error mapping.
// No CPU fallback: every entry point fails with RSR_E_CUDA if no sm_100 device is usable.
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <new>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/rsr_b200.h"
#include "conv_strip.cuh"
#include "conv_fused.cuh"
#include "conv_pair.cuh"

using namespace rsr;

// ------------------------------------------------------------------------------------ errors
static thread_local std::string g_err;
static thread_local int g_err_code = 0;

static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    g_err_code = code;
    return code;
}
#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(e__ == cudaErrorMemoryAllocation ? RSR_E_OOM : RSR_E_CUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                          \
    } while (0)

// --------------------------------------------------------------- parameter layout (network.rs)
// Order = order in which ops are added to the graph (network.rs:27-76): LinearInterp (no
// params) then the `ops` vec.  Conv kernels are [Cout][Kh][Kw][Cin].
struct Layout {
    size_t conv0, f_bias, f_activ, expand_bias, l1_bias, l2_bias, l3_bias, l1_activ, l2_activ, l3_activ;
    size_t conv1, conv2, conv3, conv5, conv6, conv7, conv8, conv9, conv10, total;
    int ce;
};
static Layout make_layout(int factor) {
    Layout L{};
    size_t o = 0;
    L.ce = 3 * factor * factor;                       // network.rs:37
    L.conv0 = o;       o += 32 * 5 * 5 * 3;           // :33
    L.f_bias = o;      o += 32;                       // :34
    L.f_activ = o;     o += 32;                       // :35
    L.expand_bias = o; o += L.ce;                     // :38
    L.l1_bias = o;     o += 32;                       // :50
    L.l2_bias = o;     o += 32;                       // :51
    L.l3_bias = o;     o += 32;                       // :52
    L.l1_activ = o;    o += 32;                       // :54
    L.l2_activ = o;    o += 32;                       // :55
    L.l3_activ = o;    o += 32;                       // :56
    L.conv1 = o;       o += 32 * 25 * 32;             // :60
    L.conv2 = o;       o += 32 * 25 * 32;             // :61
    L.conv3 = o;       o += 32 * 25 * 32;             // :62
    L.conv5 = o;       o += 32 * 9 * 32;              // :65
    L.conv6 = o;       o += 32 * 9 * 32;              // :66
    L.conv7 = o;       o += (size_t)L.ce * 9 * 32;    // :67
    L.conv8 = o;       o += 32 * 9 * 32;              // :69
    L.conv9 = o;       o += (size_t)L.ce * 9 * 32;    // :70
    L.conv10 = o;      o += (size_t)L.ce * 9 * 32;    // :72
    L.total = o;
    return L;
}

// --------------------------------------------------------------------------- kernel configs
// Ring geometry of the launches, overridable at build time for same-box A/B runs (tools/ab_session.py):
// rows per pass R, input-row slots NR, taps per weight stage TPS, weight stages NBS, partial-sum tile slots NPIN.
#ifndef RSR_CFG0_R
#define RSR_CFG0_R 8   // 8 rows per pass: conv0's MMAs are tiny, so the per-pass hand-offs dominate; 4 -> 8: -20 % (same-box A/B, r01j)
#endif
#ifndef RSR_CFG0_NR
#define RSR_CFG0_NR 20   // window of 12 rows + the 8 new rows of the next pass (the raw-image producers work a pass ahead)
#endif
#ifndef RSR_CFG1_NR
#define RSR_CFG1_NR 7
#endif
#ifndef RSR_CFG1_TPS
#define RSR_CFG1_TPS 2
#endif
#ifndef RSR_CFG1_NBS
#define RSR_CFG1_NBS 3
#endif
#ifndef RSR_L2PROMO
#define RSR_L2PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_256B
#endif
#ifndef RSR_CFG2_R
#define RSR_CFG2_R 2
#endif
#ifndef RSR_CFG2_NR
#define RSR_CFG2_NR 5
#endif
#ifndef RSR_CFG2_TPS
#define RSR_CFG2_TPS 2
#endif
#ifndef RSR_CFG2_NBS
#define RSR_CFG2_NBS 3
#endif
#ifndef RSR_CFG2_NPIN
#define RSR_CFG2_NPIN 2
#endif
#define RSR_CFG2_GEOM RSR_CFG2_R, RSR_CFG2_NR, RSR_CFG2_TPS, RSR_CFG2_NBS, RSR_CFG2_NPIN
#ifndef RSR_CFG3_R
#define RSR_CFG3_R 3   // 3 rows per pass, 2 weight stages: -2 % on conv8|9, same-box A/B r01j
#endif
#ifndef RSR_CFG3_NR
#define RSR_CFG3_NR 6
#endif
#ifndef RSR_CFG3_TPS
#define RSR_CFG3_TPS 2
#endif
#ifndef RSR_CFG3_NBS
#define RSR_CFG3_NBS 2
#endif
#ifndef RSR_CFG3_NPIN
#define RSR_CFG3_NPIN 3
#endif
#define RSR_CFG3_GEOM RSR_CFG3_R, RSR_CFG3_NR, RSR_CFG3_TPS, RSR_CFG3_NBS, RSR_CFG3_NPIN
#ifndef RSR_CFG4_R
#define RSR_CFG4_R 2
#endif
#ifndef RSR_CFG4_NR
#define RSR_CFG4_NR 6
#endif
#ifndef RSR_CFG4_NBS
#define RSR_CFG4_NBS 3
#endif
#ifndef RSR_CFG4_NPIN
#define RSR_CFG4_NPIN 3
#endif
#ifndef RSR_CFG3_NEPI
#define RSR_CFG3_NEPI 4
#endif
#ifndef RSR_PCFG1_NR
#define RSR_PCFG1_NR 7
#endif
#ifndef RSR_PCFG1_TPS
#define RSR_PCFG1_TPS 2
#endif
#ifndef RSR_PCFG1_NBS
#define RSR_PCFG1_NBS 3
#endif
#ifndef RSR_PCFG1_NEPI
#define RSR_PCFG1_NEPI 4
#endif
// Launches that run on CTA pairs (bit = launch index): conv1|2|3 (index 2: -12..18 % on pairs) and the fused middle launch
// (index 3: 455 vs 605 us).  The separate 3x3 launches, conv0 and conv10 were slower on pairs (r01i/r01j, profiles/README.md)
// and have no pair instantiation any more.
#ifndef RSR_DEFAULT_PAIR_MASK
#define RSR_DEFAULT_PAIR_MASK 12
#endif
#ifndef RSR_DEFAULT_FUSED_MID
#define RSR_DEFAULT_FUSED_MID 1   // conv5|6|7 + conv8|9 as one launch on CTA pairs: 455 vs 583 us (same-box bench A/B, r01j)
#endif

// The launch configurations of sr_net(F, None) for operand format PREC.  NGE = 32-column accumulator groups of the `expand`
// node (3 F^2 channels, network.rs:37): 1 for F <= 3, 2 for F = 4.
template <int PREC, int F>
struct Net {
    static constexpr int NGE = (3 * F * F + 31) / 32;
    //                      KH KW  RB NG  R NR TPS NBS NPIN NOUT SHUF NEPI PAIR  FOLD  PREC OFMT
    // conv0 straight from the image: four producer warps build the x-im2col rows in shared memory (K = 5 taps x 16, always
    // BF16x3 operands), row-folded; writes f in format PREC
    using C0 = ConvCfg<5, 1,  64, 1, RSR_CFG0_R, RSR_CFG0_NR, 5, 2, 0, 2, 0, 8, false, true, PREC_BF16X3, PREC, true>;
    // f -> conv1 | conv2 | conv3
    using C1 = ConvCfg<5, 5, 128, 3, 2, RSR_CFG1_NR, RSR_CFG1_TPS, RSR_CFG1_NBS, 0, 2, 0, 4, false, false, PREC>;
    using P1 = ConvCfg<5, 5, 128, 3, 2, RSR_PCFG1_NR, RSR_PCFG1_TPS, RSR_PCFG1_NBS, 0, 2, 0, RSR_PCFG1_NEPI, true, false, PREC>;
    // l1 -> conv5 | conv6 | conv7 (+ P2, P3)   [split plan]
    using C2 = std::conditional_t<NGE == 1, ConvCfg<3, 3, 128, 3, RSR_CFG2_GEOM, 2, 0, 8, false, false, PREC>,
                                            ConvCfg<3, 3, 128, 4, 1, 4, 1, 4, 2, 2, 0, 8, false, false, PREC>>;
    // l2 -> conv8 | conv9 (+ P3, E)            [split plan]
    using C3 = std::conditional_t<NGE == 1, ConvCfg<3, 3, 128, 2, RSR_CFG3_GEOM, 2, 0, RSR_CFG3_NEPI, false, false, PREC>,
                                            ConvCfg<3, 3, 128, 3, 2, 5, 2, 2, 3, 2, 0, 4, false, false, PREC>>;
    // l3 -> conv10 (+ E) + expand + bilinear skip, row-folded
    using C4 = std::conditional_t<NGE == 1, ConvCfg<3, 3, 128, 1, RSR_CFG4_R, RSR_CFG4_NR, 3, RSR_CFG4_NBS, RSR_CFG4_NPIN, 2, F, 8, false, true, PREC>,
                                            ConvCfg<3, 3, 128, 2, 2, 6, 3, 2, 2, 2, F, 8, false, true, PREC>>;
};

// Launch indices (pair masks, rsr_debug_stop_after, profile rows).  Index 0 was the separate input-packing launch of earlier
// versions; conv0 now reads the image itself, so the public per-launch times / names start at index 1.
enum { K_PACK = 0, K_CONV0, K_F, K_L1, K_L2, K_L3, K_COUNT };
static const char* kKernelNames[K_COUNT] = {"", "conv0", "conv1|2|3", "conv5|6|7", "conv8|9", "conv10+expand+skip"};

// ---------------------------------------------------------------------------------- context
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct rsr_ctx {
    int device = 0, factor = 3, num_sms = 148;
    int prec = PREC_F16MX;      // operand / activation storage format of the 32-channel launches (RSR_FLAG_BF16X3 -> PREC_BF16X3)
    Layout L{};
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // packed weights (conv_strip.cuh storage formats), biases / slopes (32 floats each, zero padded; expand_bias: 64)
    uint8_t *w0 = nullptr, *w1 = nullptr, *w2 = nullptr, *w3 = nullptr, *w4 = nullptr;
    float* vecs = nullptr;  // 8 x 32 floats: f_bias f_activ l1_bias l1_activ l2_bias l2_activ l3_bias l3_activ, then 64: expand_bias
    std::vector<float> vecs_host;   // the same on the host (the fused launch takes its bias / slope vectors as kernel parameters)
    // workspace
    DevBuf xf4, actA, actB, partA, partB, partC, out_f32, out_u8, in_stage;
    int plan = 0;               // middle levels: 0 = by width (fused unless the 124-pixel strips waste too much), 1 = fused, 2 = split
    bool fused_last = false;    // plan the most recent call ran
    int pair_mask = 0;          // bit 2: conv1|2|3 on CTA pairs, bit 3: fused middle launch on CTA pairs (conv_pair.cuh)
    int fused_dbg = 0;          // test hook rsr_debug_fused: bit 0 skip S2 MMAs, bit 1 dump l2 / raw accumulators, bit 2 profile (debug kernel),
                                // bit 3 profile of the plan's own kernel (single CTA or pair; needs a -DRSR_FUSED_PROF=1 build)
    DevBuf dbg_l2, dbg_t3, dbg_prof;
    int* fault = nullptr;       // device alias of fault_host (mapped pinned memory: readable after a kernel trap)
    int* fault_host = nullptr;  // [0] = code, [4 + 4*warp ..] = which barrier each warp role was stuck on
    cudaEvent_t ev[K_COUNT + 1] = {};
    bool kernel_timing = false; // record CUDA events around every launch (rsr_set_kernel_timing): costs the launch overlap
    bool timed = false;
    int last_launches = 0;
    int stop_after = -1;  // test hook
    // pipelined host API: 3 frames in flight
    struct Slot {
        DevBuf in, out;
        cudaEvent_t ev_in = nullptr, ev_done = nullptr, ev_out = nullptr;
        bool used = false;
    } slot[3];
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_band[4] = {};   // synchronous host calls on large frames: one event per band (upscale_host)
    long long next_ticket = 0;
    long long* prof = nullptr;  // test hook: [K_COUNT][256][16] wait-cycle counters
    int prof_grid[K_COUNT] = {};
    PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
};

constexpr int kFaultInts = 4 + 4 * 32;   // fault word + one record per warp of a CTA (mbar_timeout, ptx.cuh)

static int ensure(DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(RSR_E_OOM, "cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    }
    b.cap = bytes;
    return 0;
}
static void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

static inline void split_bf16(float w, __nv_bfloat16& hi, __nv_bfloat16& mid) {
    hi = __float2bfloat16_rn(w);
    mid = __float2bfloat16_rn(w - __bfloat162float(hi));
}

// 32 input-channel weights of one (tap, output channel) -> the 128-byte weight row of operand format `prec` (conv_strip.cuh):
// two 64-byte K chunks of 16 channels, BF16X3 [16 x bf16 hi | 16 x bf16 mid]; F16MX [16 x fp16 w_h | 16 x e4m3(w) | 16 x e4m3(2^11 w_l)].
static void pack_weight_row(const float* w, uint8_t* row, int prec) {
    for (int ci = 0; ci < 32; ++ci) {
        uint8_t* chunk = row + (ci >> 4) * 64;
        const int k = ci & 15;
        if (prec == PREC_F16MX) {
            const __half wh = __float2half_rn(w[ci]);
            reinterpret_cast<__half*>(chunk)[k] = wh;
            const float wl = w[ci] - __half2float(wh);
            chunk[32 + k] = (uint8_t)__nv_cvt_float_to_fp8(w[ci], __NV_SATFINITE, __NV_E4M3);
            chunk[48 + k] = (uint8_t)__nv_cvt_float_to_fp8(wl * F16MX_SCALE, __NV_SATFINITE, __NV_E4M3);
        } else {
            __nv_bfloat16* q = reinterpret_cast<__nv_bfloat16*>(chunk);
            split_bf16(w[ci], q[k], q[16 + k]);
        }
    }
}

// One 32-column accumulator group of a launch: output channels [o0, o0 + n) of the convolution stored at params + off.
struct WGroup { size_t off; int o0, n; };
// Packs the convolutions that read one input node into the per-tap B tiles of the kernel: out[t][group * 32 + o][128 B];
// groups shorter than 32 output channels are padded with zero rows.  fold: store the taps of a tap column bottom tap first
// (block tb * K + (K-1 - ta)), the order ConvCfg::FOLD reads them in.
static void pack_group(const float* params, const std::vector<WGroup>& groups, int K, int prec, bool fold, std::vector<uint8_t>& out) {
    const int taps = K * K, ncol = 32 * (int)groups.size();
    out.assign((size_t)taps * ncol * 128, 0);
    for (size_t g = 0; g < groups.size(); ++g)
        for (int o = 0; o < groups[g].n; ++o)
            for (int t = 0; t < taps; ++t) {
                const int blk = fold ? (t % K) * K + (K - 1 - t / K) : t;
                pack_weight_row(params + groups[g].off + ((size_t)(groups[g].o0 + o) * taps + t) * 32,
                                &out[((size_t)blk * ncol + g * 32 + o) * 128], prec);
            }
}

extern "C" size_t rsr_num_params(int factor) {
    if (factor < 1 || factor > 4) return 0;
    return make_layout(factor).total;
}

extern "C" const char* rsr_last_error(void) { return g_err.c_str(); }
extern "C" void* rsr_host_alloc(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        fail(e == cudaErrorMemoryAllocation ? RSR_E_OOM : RSR_E_CUDA, "cudaHostAlloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
extern "C" void rsr_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
extern "C" int rsr_last_error_code(void) { return g_err_code; }
extern "C" const char* rsr_version(void) { return "rsr_b200 0.2 sm_100a (tcgen05 implicit-GEMM, fp16 + mxfp8 split operands)"; }
extern "C" const char* rsr_kernel_name(int i) { return (i >= 0 && i + 1 < K_COUNT) ? kKernelNames[i + 1] : ""; }

extern "C" void rsr_destroy(rsr_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    cudaFree(c->w0); cudaFree(c->w1); cudaFree(c->w2); cudaFree(c->w3); cudaFree(c->w4);
    cudaFree(c->vecs); cudaFree(c->prof);
    if (c->fault_host) cudaFreeHost(c->fault_host);
    release(c->xf4); release(c->actA); release(c->actB); release(c->partA);
    release(c->partB); release(c->partC); release(c->dbg_l2); release(c->dbg_t3); release(c->dbg_prof); release(c->out_f32); release(c->out_u8); release(c->in_stage);
    for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    for (auto& sl : c->slot) {
        release(sl.in); release(sl.out);
        if (sl.ev_in) cudaEventDestroy(sl.ev_in);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
        if (sl.ev_out) cudaEventDestroy(sl.ev_out);
    }
    if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
    if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
    for (auto& e : c->ev_band) if (e) cudaEventDestroy(e);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

template <class T>
static int upload(T** dst, const void* src, size_t bytes) {
    CUDA_TRY(cudaMalloc((void**)dst, bytes));
    CUDA_TRY(cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice));
    return 0;
}

template <class C>
static int set_smem(void (*k)(ConvParams)) {
    CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    return 0;
}
template <int PREC, int F>
static int set_attrs() {
    using N = Net<PREC, F>;
    if (int rc = set_smem<typename N::C0>(conv_strip_kernel<typename N::C0>)) return rc;
    if (int rc = set_smem<typename N::C1>(conv_strip_kernel<typename N::C1>)) return rc;
    if (int rc = set_smem<typename N::P1>(conv_pair_kernel<typename N::P1>)) return rc;
    if (int rc = set_smem<typename N::C2>(conv_strip_kernel<typename N::C2>)) return rc;
    if (int rc = set_smem<typename N::C3>(conv_strip_kernel<typename N::C3>)) return rc;
    if (int rc = set_smem<typename N::C4>(conv_strip_kernel<typename N::C4>)) return rc;
    if (N::NGE == 1) {
        CUDA_TRY(cudaFuncSetAttribute(conv_fused_mid_kernel<false, false, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfgT<false, PREC>::SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(conv_fused_mid_kernel<true, false, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfgT<false, PREC>::SMEM_BYTES));
        CUDA_TRY(cudaFuncSetAttribute(conv_fused_mid_kernel<false, true, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfgT<true, PREC>::SMEM_BYTES));
    }
    return 0;
}
// Calls fn.template operator()<PREC, F>() for the context's operand format and factor.
template <class Fn>
static int dispatch(int prec, int factor, Fn&& fn) {
    if (prec == PREC_F16MX) {
        switch (factor) {
            case 1: return fn.template operator()<PREC_F16MX, 1>();
            case 2: return fn.template operator()<PREC_F16MX, 2>();
            case 3: return fn.template operator()<PREC_F16MX, 3>();
            case 4: return fn.template operator()<PREC_F16MX, 4>();
        }
    } else {
        switch (factor) {
            case 1: return fn.template operator()<PREC_BF16X3, 1>();
            case 2: return fn.template operator()<PREC_BF16X3, 2>();
            case 3: return fn.template operator()<PREC_BF16X3, 3>();
            case 4: return fn.template operator()<PREC_BF16X3, 4>();
        }
    }
    return fail(RSR_E_UNSUPPORTED, "factor %d", factor);
}
struct SetAttrs { template <int PREC, int F> int operator()() const { return set_attrs<PREC, F>(); } };

static int create_impl(rsr_ctx* c, const float* params) {
    const Layout& L = c->L;
    CUDA_TRY(cudaSetDevice(c->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, c->device));
    if (prop.major != 10)
        return fail(RSR_E_CUDA, "device %d is sm_%d%d; this library contains sm_100a code only (no fallback)",
                    c->device, prop.major, prop.minor);
    c->num_sms = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) return fail(RSR_E_CUDA, "driver lacks cuTensorMapEncodeTiled");
    c->encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (auto& e : c->ev) CUDA_TRY(cudaEventCreate(&e));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking));
    for (auto& e : c->ev_band) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& sl : c->slot) {
        CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&sl.ev_out, cudaEventDisableTiming));
    }
    CUDA_TRY(cudaHostAlloc((void**)&c->fault_host, kFaultInts * sizeof(int), cudaHostAllocMapped));
    memset(c->fault_host, 0, kFaultInts * sizeof(int));
    CUDA_TRY(cudaHostGetDevicePointer((void**)&c->fault, c->fault_host, 0));

    // conv0: [32][5][5][3] -> w0[4 - a][n][k = b*3 + c | 16 + k]  (x-im2col form, 64 B rows of bf16 hi | mid; tap rows stored
    // bottom tap first: the row-folded order)
    {
        std::vector<__nv_bfloat16> w((size_t)5 * 32 * 32, __float2bfloat16_rn(0.f));
        for (int n = 0; n < 32; ++n)
            for (int a = 0; a < 5; ++a)
                for (int b = 0; b < 5; ++b)
                    for (int ch = 0; ch < 3; ++ch) {
                        __nv_bfloat16 h, m;
                        split_bf16(params[L.conv0 + (((size_t)n * 5 + a) * 5 + b) * 3 + ch], h, m);
                        const size_t row = ((size_t)(4 - a) * 32 + n) * 32;
                        w[row + b * 3 + ch] = h;
                        w[row + 16 + b * 3 + ch] = m;
                    }
        if (int rc = upload(&c->w0, w.data(), w.size() * 2)) return rc;
    }
    const int ce = L.ce, nge = (ce + 31) / 32;
    auto expand_groups = [&](std::vector<WGroup>& g, size_t off) {
        for (int k = 0; k < nge; ++k) g.push_back({off, k * 32, std::min(32, ce - k * 32)});
    };
    std::vector<uint8_t> w;
    {
        std::vector<WGroup> g = {{L.conv1, 0, 32}, {L.conv2, 0, 32}, {L.conv3, 0, 32}};
        pack_group(params, g, 5, c->prec, false, w);
        if (int rc = upload(&c->w1, w.data(), w.size())) return rc;
    }
    {
        std::vector<WGroup> g = {{L.conv5, 0, 32}, {L.conv6, 0, 32}};
        expand_groups(g, L.conv7);
        pack_group(params, g, 3, c->prec, false, w);
        if (int rc = upload(&c->w2, w.data(), w.size())) return rc;
    }
    {
        std::vector<WGroup> g = {{L.conv8, 0, 32}};
        expand_groups(g, L.conv9);
        pack_group(params, g, 3, c->prec, false, w);
        if (int rc = upload(&c->w3, w.data(), w.size())) return rc;
    }
    {
        std::vector<WGroup> g;
        expand_groups(g, L.conv10);
        pack_group(params, g, 3, c->prec, true, w);
        if (int rc = upload(&c->w4, w.data(), w.size())) return rc;
    }
    {
        std::vector<float> v(8 * 32 + 64, 0.f);
        const size_t src[8] = {L.f_bias, L.f_activ, L.l1_bias, L.l1_activ, L.l2_bias, L.l2_activ, L.l3_bias, L.l3_activ};
        for (int i = 0; i < 8; ++i) memcpy(&v[i * 32], params + src[i], sizeof(float) * 32);
        memcpy(&v[8 * 32], params + L.expand_bias, sizeof(float) * ce);
        if (int rc = upload(&c->vecs, v.data(), v.size() * 4)) return rc;
        c->vecs_host = v;
    }
    return dispatch(c->prec, c->factor, SetAttrs{});
}

// pair_mask < 0: the library default
static rsr_ctx* create_ctx(const float* params, size_t n_params, int factor, int device, int flags, int pair_mask) {
    g_err.clear();
    g_err_code = 0;
    if (!params) { fail(RSR_E_BAD_ARG, "params is null"); return nullptr; }
    if (factor < 1 || factor > 4) {
        fail(RSR_E_BAD_ARG, "factor %d out of range 1..4 (reference FACTOR is 3, main.rs:31)", factor);
        return nullptr;
    }
    if ((flags & RSR_FLAG_FUSED_MID) && factor == 4) {
        fail(RSR_E_UNSUPPORTED, "the fused middle launch holds 32 expand channels; factor 4 (48) runs the split plan");
        return nullptr;
    }
    const Layout L = make_layout(factor);
    if (n_params != L.total) {
        fail(RSR_E_PARAM_COUNT,
             "Parameters selected do not have the size required by the neural net (%zu given, %zu required)",
             n_params, L.total);
        return nullptr;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        fail(RSR_E_CUDA, "no CUDA device available (this library has no CPU fallback)");
        return nullptr;
    }
    if (device < 0 || device >= ndev) { fail(RSR_E_BAD_ARG, "device %d out of range (0..%d)", device, ndev - 1); return nullptr; }
    rsr_ctx* c = new (std::nothrow) rsr_ctx();
    if (!c) { fail(RSR_E_OOM, "host allocation failed"); return nullptr; }
    c->device = device;
    c->factor = factor;
    c->L = L;
    c->prec = (flags & RSR_FLAG_BF16X3) ? PREC_BF16X3 : PREC_F16MX;
    c->plan = (flags & RSR_FLAG_SPLIT_MID) ? 2 : (flags & RSR_FLAG_FUSED_MID) ? 1 : (RSR_DEFAULT_FUSED_MID ? 0 : 2);
    if (factor == 4) c->plan = 2;
    c->pair_mask = (flags & RSR_FLAG_NO_PAIRS) ? 0 : (pair_mask >= 0 ? pair_mask & 12 : RSR_DEFAULT_PAIR_MASK);
#ifdef RSR_LAB
    // lab builds only (tools/ab_session.py): environment overrides for same-box A/B runs
    if (const char* e = getenv("RSR_B200_FUSED")) c->plan = e[0] == '1' ? 1 : 2;
    if (const char* e = getenv("RSR_B200_PAIR")) c->pair_mask = atoi(e) & 12;
    if (const char* e = getenv("RSR_B200_PREC")) c->prec = e[0] == '0' ? PREC_BF16X3 : PREC_F16MX;
#endif
    if (create_impl(c, params) != 0) {
        const std::string keep = g_err;
        const int code = g_err_code;
        rsr_destroy(c);
        g_err = keep;
        g_err_code = code;
        return nullptr;
    }
    return c;
}
extern "C" rsr_ctx* rsr_create(const float* params, size_t n_params, int factor, int device, int flags) {
    return create_ctx(params, n_params, factor, device, flags, -1);
}
extern "C" rsr_ctx* rsr_debug_create(const float* params, size_t n_params, int factor, int device, int flags, int pair_mask) {
    return create_ctx(params, n_params, factor, device, flags, pair_mask);
}

extern "C" int rsr_set_stream(rsr_ctx* c, void* s) {
    if (!c) return fail(RSR_E_BAD_ARG, "ctx is null");
    c->stream = s ? reinterpret_cast<cudaStream_t>(s) : c->own_stream;
    return 0;
}

extern "C" int rsr_synchronize(rsr_ctx* c) {
    if (!c) return fail(RSR_E_BAD_ARG, "ctx is null");
    CUDA_TRY(cudaSetDevice(c->device));
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        std::string log;
        if (c->fault_host)
            for (int wq = 0; wq < 32; ++wq) {
                const int* r = c->fault_host + 4 + wq * 4;
                if (r[0]) { char b[96]; snprintf(b, sizeof b, " [warp %d: tag %d cta %d parity %d bar@0x%x]", wq, r[0], r[1], r[2], r[3]); log += b; }
            }
        return fail(RSR_E_CUDA, "kernel failed: %s (fault word 0x%x)%s", cudaGetErrorString(e), c->fault_host ? c->fault_host[0] : 0, log.c_str());
    }
    return 0;
}

// ------------------------------------------------------------------------------ tensor maps
// activation / xcol tensor: rows of `row_bytes` (128: swizzle 128B; 64: swizzle 64B) per pixel, described as 2-byte elements
// box_bytes: bytes of a pixel row one box covers (default: the whole row); 64-byte boxes use the 64-byte swizzle
static int make_map_act(rsr_ctx* c, CUtensorMap* m, const void* base, int row_bytes, int W, int H, int N, int boxw, int box_bytes = 0) {
    const cuuint64_t rb = (cuuint64_t)row_bytes;
    const int elems = row_bytes / 2;
    if (!box_bytes) box_bytes = row_bytes;
    cuuint64_t dims[4] = {(cuuint64_t)elems, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {rb, rb * W, rb * W * H};
    cuuint32_t box[4] = {(cuuint32_t)(box_bytes / 2), (cuuint32_t)boxw, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = c->encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           box_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                           RSR_L2PROMO, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RSR_E_CUDA, "cuTensorMapEncodeTiled(activation %dx%dx%d) failed: %d", W, H, N, (int)r);
    return 0;
}
static int make_map_part(rsr_ctx* c, CUtensorMap* m, const void* base, int W, int H, int N, int boxw, int box_ch = 32) {
    cuuint64_t dims[4] = {32, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {128, (cuuint64_t)128 * W, (cuuint64_t)128 * W * H};
    cuuint32_t box[4] = {(cuuint32_t)box_ch, (cuuint32_t)boxw, 1, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = c->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, box_ch == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RSR_E_CUDA, "cuTensorMapEncodeTiled(partial %dx%dx%d) failed: %d", W, H, N, (int)r);
    return 0;
}
static int make_map_w(rsr_ctx* c, CUtensorMap* m, const void* base, int row_bytes, int rows, int boxrows) {
    const int elems = row_bytes / 2;
    cuuint64_t dims[2] = {(cuuint64_t)elems, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
    cuuint32_t box[2] = {(cuuint32_t)elems, (cuuint32_t)boxrows};
    cuuint32_t es[2] = {1, 1};
    CUresult r = c->encode(m, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE,
                           row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                           RSR_L2PROMO, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RSR_E_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    return 0;
}

// One kernel launch: optional cluster of 2 CTAs, optional programmatic dependent launch (the kernel calls griddep_wait before it
// touches the previous launch's data, so its prologue overlaps that launch's tail).
template <class P>
static int launch_kernel(void (*kernel)(P), int grid, int block, int smem, cudaStream_t stream, bool pair, bool pdl, const P& params) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (pair) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
#ifdef RSR_NO_PDL   // lab builds: A/B of the launch overlap
    pdl = false;
#endif
    if (pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at;
    cfg.numAttrs = na;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kernel, params));
    return 0;
}

// CTA-pair launch (conv_pair.cuh): p.tmB must have been built with a box of NCOL / 2 rows (one CTA's half tap).
template <class C>
static int launch_conv_pair(rsr_ctx* c, ConvParams& p, int kindex) {
    p.n_strips = (p.W + 127) / 128;
    p.total_rowtiles = (long long)p.NIMG * p.n_strips * (p.row_end - p.row_begin);
    const int ncl = (int)std::min<long long>(p.total_rowtiles, c->num_sms / 2);
    p.prof = nullptr;
    c->prof_grid[kindex] = 0;
    if (int rc = launch_kernel(conv_pair_kernel<C>, 2 * ncl, C::NTHREADS, C::SMEM_BYTES, c->stream, true, !c->kernel_timing, p)) return rc;
    c->last_launches++;
    return 0;
}
template <class C>
static int launch_conv(rsr_ctx* c, ConvParams& p, int kindex) {
    p.n_strips = (p.W + 127) / 128;
    p.total_rowtiles = (long long)p.NIMG * p.n_strips * (p.row_end - p.row_begin);
    const int grid = (int)std::min<long long>(p.total_rowtiles, c->num_sms);
    p.prof = (c->prof && grid <= 256) ? c->prof + (size_t)kindex * 256 * 16 : nullptr;
    c->prof_grid[kindex] = grid;
    if (int rc = launch_kernel(conv_strip_kernel<C>, grid, C::NTHREADS, C::SMEM_BYTES, c->stream, false, !c->kernel_timing, p)) return rc;
    c->last_launches++;
    return 0;
}

// Output selection of one forward pass: exactly one of f32 / u8 (u8_ch = 4 RGBA or 3 RGB); top / bottom = halo rows of a band
// (rows of the input whose outputs are not produced; 0 for whole images).
struct OutSpec {
    float* f32 = nullptr;
    uint8_t* u8 = nullptr;
    int u8_ch = 4;
    int top = 0, bottom = 0;
};

// Middle levels on the fused launch?  It walks 124-pixel strips where the split launches walk 128-pixel strips, so narrow
// images can pay for a nearly empty extra strip; per strip the fused launch costs FUSED_STRIP_COST of the two split launches
// (same-box measurements, profiles/README.md).
#ifndef RSR_FUSED_STRIP_COST
#define RSR_FUSED_STRIP_COST 0.62   // (0.75 before the fused launch's epilogues were rebuilt; tools/plan_by_width.py, profiles/r03p_*)
#endif
static bool use_fused(const rsr_ctx* c, int w) {
    if (c->plan == 1) return true;
    if (c->plan == 2) return false;
    const int nf = (w + FusedCfg::STRIP - 1) / FusedCfg::STRIP, ns = (w + 127) / 128;
    return nf * RSR_FUSED_STRIP_COST <= ns;
}

// The whole forward pass on device-resident input.  in: f32 [n][h][w][3] or u8 [n][h][w][ch].
template <int PREC, int F>
static int forward_impl(rsr_ctx* c, const void* d_in, bool u8, int in_ch, int n, int w, int h, const OutSpec& o) {
    using N = Net<PREC, F>;
    constexpr int NGE = N::NGE;
    const size_t npx = (size_t)n * w * h;
    if (int rc = ensure(c->xf4, npx * 16)) return rc;
    if (int rc = ensure(c->actA, npx * 128)) return rc;
    if (int rc = ensure(c->actB, npx * 128)) return rc;
    if (int rc = ensure(c->partA, npx * 128)) return rc;
    if (int rc = ensure(c->partB, npx * 128)) return rc;
    const bool fused = NGE == 1 && use_fused(c, w);
    if (fused || NGE == 2) { if (int rc = ensure(c->partC, npx * 128)) return rc; }
    c->fused_last = fused;

    uint8_t* f_act = (uint8_t*)c->actA.p;   // f, later l2
    uint8_t* l1_act = (uint8_t*)c->actB.p;  // l1, later l3
    uint8_t* l2_act = f_act;
    uint8_t* l3_act = l1_act;               // (the fused middle launch puts l3 into f's buffer instead)
    float* P2 = (float*)c->partA.p;  // conv2(f): fp32 partial of l2_conv; later E (expand partial)
    float* P3 = (float*)c->partB.p;  // conv3(f) (+conv6(l1)): fp32 partial of l3_conv, updated in place
    float* E[2] = {P2, (float*)c->partC.p};   // conv7(l1) (+conv9(l2)) per 32-column group, updated in place (aliasing)
    const float* V = c->vecs;
    const float *f_bias = V, *f_slope = V + 32, *l1_bias = V + 64, *l1_slope = V + 96, *l2_bias = V + 128,
                *l2_slope = V + 160, *l3_bias = V + 192, *l3_slope = V + 224, *e_bias = V + 256;

    c->last_launches = 0;
    if (c->kernel_timing) CUDA_TRY(cudaEventRecord(c->ev[1], c->stream));
    if (c->stop_after == 0) return 0;

    ConvParams p{};
    p.W = w; p.H = h; p.NIMG = n; p.row_begin = 0; p.row_end = h; p.fault = c->fault;
    p.n_cb = 1;
    p.xf4 = (const float4*)c->xf4.p;
    const EpiGroup none{EPI_NONE, 0, nullptr, nullptr, 0, 0};
    auto set_groups = [&](std::initializer_list<EpiGroup> g) {
        int i = 0;
        for (const EpiGroup& e : g) p.grp[i++] = e;
        for (; i < 8; ++i) p.grp[i] = none;
    };
    auto in_map = [&](CUtensorMap* m, const void* base, int boxw) { return make_map_act(c, m, base, 128, w, h, n, boxw); };
    auto pair = [&](int kindex) { return (c->pair_mask >> kindex & 1) != 0; };
    auto part_map = [&](CUtensorMap* m, const void* base, int boxw) { return make_map_part(c, m, base, w, h, n, boxw); };

    // conv0: image -> f = BeLU(conv0(x) + f_bias), x as float4 for the skip    main.rs:170, network.rs:33-35
    p.raw_in = d_in; p.raw_u8 = u8 ? 1 : 0; p.raw_ch = u8 ? in_ch : 3;
    p.xf4_out = (float4*)c->xf4.p;
    if (int rc = make_map_w(c, &p.tmB, c->w0, 64, 5 * 32, N::C0::TPS * 32)) return rc;
    if (int rc = in_map(&p.tmOut[0], f_act, 32)) return rc;
    set_groups({EpiGroup{EPI_ACT, 0, f_bias, f_slope, 0, 0}});
    if (int rc = launch_conv<typename N::C0>(c, p, 1)) return rc;
    if (c->kernel_timing) CUDA_TRY(cudaEventRecord(c->ev[2], c->stream));
    if (c->stop_after == 1) return 0;

    // f -> l1 = BeLU(conv1(f) + l1_bias); P2 = conv2(f); P3 = conv3(f)  network.rs:60-62,50,54
    if (int rc = in_map(&p.tmA, f_act, N::C1::BOXW)) return rc;
    if (int rc = make_map_w(c, &p.tmB, c->w1, 128, 25 * 96, pair(2) ? 48 : N::C1::TPS * 96)) return rc;
    if (int rc = in_map(&p.tmOut[0], l1_act, 32)) return rc;
    if (int rc = part_map(&p.tmOut[1], P2, 32)) return rc;
    if (int rc = part_map(&p.tmOut[2], P3, 32)) return rc;
    set_groups({EpiGroup{EPI_ACT, 0, l1_bias, l1_slope, 0, 0}, EpiGroup{EPI_RAW, 0, nullptr, nullptr, 0, 0},
                EpiGroup{EPI_RAW, 0, nullptr, nullptr, 0, 0}});
    if (pair(2)) { if (int rc = launch_conv_pair<typename N::P1>(c, p, 2)) return rc; }
    else         { if (int rc = launch_conv<typename N::C1>(c, p, 2)) return rc; }
    if (c->kernel_timing) CUDA_TRY(cudaEventRecord(c->ev[3], c->stream));
    if (c->stop_after == 2) return 0;

    if constexpr (NGE == 1) {
    if (fused) {
        // l1, P2, P3 -> l3, E in one launch; l2 lives in shared memory only (conv_fused.cuh)   :65-70,51-52,55-56
        // l3 goes to the buffer of the (dead) f; E to a buffer of its own: the fused launch reads P2 rows of
        // neighbouring work items (l2 halo rows), so E may not overwrite P2 here.
        l3_act = f_act;
        E[0] = (float*)c->partC.p;
        FusedParams q{};
        q.W = w; q.H = h; q.NIMG = n; q.fault = c->fault;
        q.n_strips = (w + FusedCfg::STRIP - 1) / FusedCfg::STRIP;
        q.total_rowtiles = (long long)n * q.n_strips * h;
        if (int rc = in_map(&q.tmA, l1_act, 128)) return rc;
        // the fused launch runs on CTA pairs when launch index 3 is in the pair mask (half-tap weight boxes);
        // the debug hooks exist in the single-CTA debug instantiation only
        const bool fpair = pair(3) && (c->fused_dbg & 7) == 0;
        if (int rc = make_map_w(c, &q.tmB1, c->w2, 128, 9 * 96, fpair ? 48 : 96)) return rc;
        if (int rc = make_map_w(c, &q.tmB2, c->w3, 128, 9 * 64, fpair ? 32 : 64)) return rc;
        if (int rc = part_map(&q.tmP2, P2, 128)) return rc;
        if (int rc = part_map(&q.tmP3, P3, 128)) return rc;
        {   // half-row (one K chunk = 64 B) boxes of 32 / 31 / 29 pixels for the E2 warps' stores
            const int bw[3] = {32, 31, 29};
            for (int k = 0; k < 3; ++k) {
                if (int rc = make_map_act(c, &q.tmL3[k], l3_act, 128, w, h, n, bw[k], 64)) return rc;
                if (int rc = make_map_part(c, &q.tmE[k], E[0], w, h, n, bw[k], 16)) return rc;
            }
        }
        memcpy(q.l2_bias, &c->vecs_host[4 * 32], sizeof q.l2_bias);
        memcpy(q.l2_slope, &c->vecs_host[5 * 32], sizeof q.l2_slope);
        memcpy(q.l3_bias, &c->vecs_host[6 * 32], sizeof q.l3_bias);
        memcpy(q.l3_slope, &c->vecs_host[7 * 32], sizeof q.l3_slope);
        q.dbg_flags = c->fused_dbg & 1;
        if (c->fused_dbg & 2) {
            if (int rc = ensure(c->dbg_l2, npx * 128)) return rc;
            if (int rc = ensure(c->dbg_t3, npx * 128)) return rc;
            CUDA_TRY(cudaMemsetAsync(c->dbg_l2.p, 0, npx * 128, c->stream));
            CUDA_TRY(cudaMemsetAsync(c->dbg_t3.p, 0, npx * 128, c->stream));
            q.dbg_l2 = (float*)c->dbg_l2.p;
            q.dbg_t3 = (float*)c->dbg_t3.p;
        }
        if (c->fused_dbg & (4 | 8)) {
            if (int rc = ensure(c->dbg_prof, 256 * FPROF_N * sizeof(long long))) return rc;
            CUDA_TRY(cudaMemsetAsync(c->dbg_prof.p, 0, 256 * FPROF_N * sizeof(long long), c->stream));
            q.prof = (long long*)c->dbg_prof.p;
        }
        const bool pdl = !c->kernel_timing;
        if (fpair) {
            const int ncl = (int)std::min<long long>(q.total_rowtiles, c->num_sms / 2);
            if (int rc = launch_kernel(conv_fused_mid_kernel<false, true, PREC>, 2 * ncl, FusedCfg::NTHREADS, FusedCfgT<true, PREC>::SMEM_BYTES, c->stream, true, pdl, q)) return rc;
        } else {
            const int grid = (int)std::min<long long>(q.total_rowtiles, c->num_sms);
            if (c->fused_dbg & 7) { if (int rc = launch_kernel(conv_fused_mid_kernel<true, false, PREC>, grid, FusedCfg::NTHREADS, FusedCfgT<false, PREC>::SMEM_BYTES, c->stream, false, pdl, q)) return rc; }
            else { if (int rc = launch_kernel(conv_fused_mid_kernel<false, false, PREC>, grid, FusedCfg::NTHREADS, FusedCfgT<false, PREC>::SMEM_BYTES, c->stream, false, pdl, q)) return rc; }
        }
        CUDA_TRY(cudaGetLastError());
        c->last_launches++;
        // reported in the conv5|6|7 slot of rsr_last_kernel_times; the conv8|9 slot then reads ~0
        if (c->kernel_timing) CUDA_TRY(cudaEventRecord(c->ev[4], c->stream));
        if (c->kernel_timing) CUDA_TRY(cudaEventRecord(c->ev[5], c->stream));
        if (c->stop_after == 3 || c->stop_after == 4) return 0;
    }
    }
    if (!fused) {
        // l1 -> l2 = BeLU(P2 + conv5(l1) + l2_bias); P3 += conv6(l1); E = conv7(l1)   :65-67,51,55
        // (l2 overwrites f, E[0] overwrites P2: both are dead; a thread block only overwrites pixels whose
        //  partial tile it has already consumed — see DESIGN.md "aliasing")
        if (int rc = in_map(&p.tmA, l1_act, N::C2::BOXW)) return rc;
        if (int rc = make_map_w(c, &p.tmB, c->w2, 128, 9 * N::C2::NCOL, N::C2::TPS * N::C2::NCOL)) return rc;
        if (int rc = part_map(&p.tmPin[0], P2, 128)) return rc;
        if (int rc = part_map(&p.tmPin[1], P3, 128)) return rc;
        if (int rc = in_map(&p.tmOut[0], l2_act, 32)) return rc;
        if (int rc = part_map(&p.tmOut[1], P3, 32)) return rc;
        for (int k = 0; k < NGE; ++k)
            if (int rc = part_map(&p.tmOut[2 + k], E[k], 32)) return rc;
        set_groups({EpiGroup{EPI_ACT, 1, l2_bias, l2_slope, 0, 0}, EpiGroup{EPI_RAW, 1, nullptr, nullptr, 0, 0},
                    EpiGroup{EPI_RAW, 0, nullptr, nullptr, 0, 0}, NGE == 2 ? EpiGroup{EPI_RAW, 0, nullptr, nullptr, 0, 0} : none});
        if (int rc = launch_conv<typename N::C2>(c, p, 3)) return rc;
        if (c->kernel_timing) CUDA_TRY(cudaEventRecord(c->ev[4], c->stream));
        if (c->stop_after == 3) return 0;

        // l2 -> l3 = BeLU(P3 + conv8(l2) + l3_bias); E += conv9(l2)                  :69-70,52,56
        if (int rc = in_map(&p.tmA, l2_act, N::C3::BOXW)) return rc;
        if (int rc = make_map_w(c, &p.tmB, c->w3, 128, 9 * N::C3::NCOL, N::C3::TPS * N::C3::NCOL)) return rc;
        if (int rc = part_map(&p.tmPin[0], P3, 128)) return rc;
        if (int rc = in_map(&p.tmOut[0], l3_act, 32)) return rc;
        for (int k = 0; k < NGE; ++k) {
            if (int rc = part_map(&p.tmPin[1 + k], E[k], 128)) return rc;
            if (int rc = part_map(&p.tmOut[1 + k], E[k], 32)) return rc;
        }
        set_groups({EpiGroup{EPI_ACT, 1, l3_bias, l3_slope, 0, 0}, EpiGroup{EPI_RAW, 1, nullptr, nullptr, 0, 0},
                    NGE == 2 ? EpiGroup{EPI_RAW, 1, nullptr, nullptr, 0, 0} : none});
        if (int rc = launch_conv<typename N::C3>(c, p, 4)) return rc;
        if (c->kernel_timing) CUDA_TRY(cudaEventRecord(c->ev[5], c->stream));
        if (c->stop_after == 4) return 0;
    }

    // l3 -> output = Expand(E + conv10(l3) + expand_bias) + LinearInterp(x)          :72,38-39,27
    if (int rc = in_map(&p.tmA, l3_act, N::C4::BOXW)) return rc;
    if (int rc = make_map_w(c, &p.tmB, c->w4, 128, 9 * N::C4::NCOL, N::C4::TPS * N::C4::NCOL)) return rc;
    for (int k = 0; k < NGE; ++k)
        if (int rc = part_map(&p.tmPin[k], E[k], 128)) return rc;
    set_groups({EpiGroup{EPI_SHUFFLE, 1, e_bias, nullptr, 0, 0}, NGE == 2 ? EpiGroup{EPI_SHUFFLE, 1, e_bias + 32, nullptr, 0, 0} : none});
    p.row_begin = o.top; p.row_end = h - o.bottom;      // band mode: the halo rows' outputs are not produced
    p.out_row0 = o.top; p.out_rows = h - o.top - o.bottom;
    p.out_f32 = o.f32;
    p.out_u8 = o.u8;
    p.out_u8_ch = o.u8_ch;
    p.out_vec_ok = ((w * F) % 4 == 0) && ((reinterpret_cast<uintptr_t>(o.f32 ? (void*)o.f32 : (void*)o.u8) & 15) == 0);
    if (int rc = launch_conv<typename N::C4>(c, p, 5)) return rc;
    if (c->kernel_timing) CUDA_TRY(cudaEventRecord(c->ev[6], c->stream));
    c->timed = c->kernel_timing;
    return 0;
}

struct Forward {
    rsr_ctx* c; const void* d_in; bool u8; int in_ch, n, w, h; const OutSpec& o;
    template <int PREC, int F> int operator()() const { return forward_impl<PREC, F>(c, d_in, u8, in_ch, n, w, h, o); }
};
static int forward_device(rsr_ctx* c, const void* d_in, bool u8, int in_ch, int n, int w, int h, const OutSpec& o) {
    if (!c) return fail(RSR_E_BAD_ARG, "ctx is null");
    if (!d_in || (!o.f32 && !o.u8)) return fail(RSR_E_BAD_ARG, "null buffer");
    if (o.f32 && o.u8) return fail(RSR_E_BAD_ARG, "exactly one of the f32 / u8 outputs may be requested");
    if (n < 1 || w < 1 || h < 1) return fail(RSR_E_BAD_ARG, "bad dimensions n=%d w=%d h=%d", n, w, h);
    if (u8 && in_ch != 3 && in_ch != 4) return fail(RSR_E_BAD_ARG, "in_channels must be 3 or 4 (got %d)", in_ch);
    if (o.u8 && o.u8_ch != 3 && o.u8_ch != 4) return fail(RSR_E_BAD_ARG, "out_channels must be 3 or 4 (got %d)", o.u8_ch);
    if (o.top < 0 || o.bottom < 0 || h - o.top - o.bottom < 1 || ((o.top || o.bottom) && n != 1))
        return fail(RSR_E_BAD_ARG, "bad band: rows_in=%d top_halo=%d bottom_halo=%d", h, o.top, o.bottom);
    CUDA_TRY(cudaSetDevice(c->device));
    return dispatch(c->prec, c->factor, Forward{c, d_in, u8, in_ch, n, w, h, o});
}

static OutSpec out_spec(float* f32, uint8_t* u8, int u8_ch, int top = 0, int bottom = 0) {
    OutSpec o;
    o.f32 = f32; o.u8 = u8; o.u8_ch = u8_ch; o.top = top; o.bottom = bottom;
    return o;
}
extern "C" int rsr_upscale_device_f32(rsr_ctx* c, const float* d_in, int n, int w, int h, float* d_out_f32,
                                      uint8_t* d_out_rgba) {
    return forward_device(c, d_in, false, 3, n, w, h, out_spec(d_out_f32, d_out_rgba, 4));
}
extern "C" int rsr_upscale_device_u8(rsr_ctx* c, const uint8_t* d_in, int n, int w, int h, int in_channels,
                                     float* d_out_f32, uint8_t* d_out_rgba) {
    return forward_device(c, d_in, true, in_channels, n, w, h, out_spec(d_out_f32, d_out_rgba, 4));
}
extern "C" int rsr_upscale_device_u8_rgb(rsr_ctx* c, const uint8_t* d_in, int n, int w, int h, int in_channels, uint8_t* d_out_rgb) {
    return forward_device(c, d_in, true, in_channels, n, w, h, out_spec(nullptr, d_out_rgb, 3));
}
extern "C" int rsr_upscale_band_device_f32(rsr_ctx* c, const float* d_in_rows, int w, int rows_in, int top_halo, int bottom_halo,
                                           float* d_out_rows) {
    return forward_device(c, d_in_rows, false, 3, 1, w, rows_in, out_spec(d_out_rows, nullptr, 4, top_halo, bottom_halo));
}
extern "C" int rsr_upscale_band_device_u8(rsr_ctx* c, const uint8_t* d_in_rows, int w, int rows_in, int top_halo, int bottom_halo,
                                          int in_channels, uint8_t* d_out_rows, int out_channels) {
    return forward_device(c, d_in_rows, true, in_channels, 1, w, rows_in, out_spec(nullptr, d_out_rows, out_channels, top_halo, bottom_halo));
}

// Host-buffer entry points: H2D, forward, D2H on the context's stream, then synchronise.  out_ch: 0 = f32 RGB, 3 / 4 = u8.
// Receptive field of sr_net in rows: 2 + 2 + 1 + 1 + 1 (network.rs:33,60-72) — the halo a band needs on each side.
constexpr int kHaloRows = 7;
// The synchronous host calls on one large frame: the device->host copy of the result takes as long as the forward pass
// (u8) or several times longer (f32), so the frame runs as up to kSyncBands horizontal bands (halo rows recomputed; the
// results are bit-identical to the whole frame, like rsr_upscale_band_*) and each band's rows travel to the host while the
// next band computes.
constexpr int kSyncBands = 4;
constexpr int kSyncBandMinRows = 128;           // LR rows per band below which the halo recompute (14 rows) is not worth it
constexpr size_t kSyncBandMinBytes = 8u << 20;  // result size below which the copy is too short to matter

static int upscale_host(rsr_ctx* c, const void* in, bool u8, int in_ch, int n, int w, int h, void* out, int out_ch,
                        int top = 0, int bottom = 0) {
    if (!c) return fail(RSR_E_BAD_ARG, "ctx is null");
    if (!in || !out) return fail(RSR_E_BAD_ARG, "null buffer");
    if (n < 1 || w < 1 || h < 1) return fail(RSR_E_BAD_ARG, "bad dimensions n=%d w=%d h=%d", n, w, h);
    if (u8 && in_ch != 3 && in_ch != 4) return fail(RSR_E_BAD_ARG, "in_channels must be 3 or 4 (got %d)", in_ch);
    if (top < 0 || bottom < 0 || h - top - bottom < 1)
        return fail(RSR_E_BAD_ARG, "bad band: rows_in=%d top_halo=%d bottom_halo=%d", h, top, bottom);
    CUDA_TRY(cudaSetDevice(c->device));
    const int F = c->factor;
    const size_t npx = (size_t)n * w * h;
    const size_t in_px = u8 ? (size_t)in_ch : 3 * sizeof(float), out_px = out_ch ? (size_t)out_ch : 3 * sizeof(float);
    const size_t in_bytes = npx * in_px;
    if (int rc = ensure(c->in_stage, in_bytes)) return rc;
    CUDA_TRY(cudaMemcpyAsync(c->in_stage.p, in, in_bytes, cudaMemcpyHostToDevice, c->stream));
    const size_t opx = (size_t)n * w * (h - top - bottom) * F * F;
    const size_t out_bytes = opx * out_px;
    DevBuf& ob = out_ch ? c->out_u8 : c->out_f32;
    if (int rc = ensure(ob, out_bytes)) return rc;
    auto spec = [&](uint8_t* d_out, int t, int b) {
        return out_spec(out_ch ? nullptr : (float*)d_out, out_ch ? d_out : nullptr, out_ch ? out_ch : 4, t, b);
    };
    // (not under the test hooks / per-launch timing, which look at the buffers and events of ONE whole-frame pass)
    const bool whole = c->kernel_timing || c->stop_after >= 0 || c->fused_dbg || c->prof;
    const int r0 = top, r1 = h - bottom;             // LR rows whose outputs are produced (a band call's own halo rows are not)
    const int nb = (n == 1 && !whole && out_bytes >= kSyncBandMinBytes) ? std::min(kSyncBands, (r1 - r0) / kSyncBandMinRows) : 1;
    if (nb < 2) {
        if (int rc = forward_device(c, c->in_stage.p, u8, in_ch, n, w, h, spec((uint8_t*)ob.p, top, bottom))) return rc;
        if (int rc = rsr_synchronize(c)) return rc;   // surfaces a kernel fault (with its fault word) before the copy
        CUDA_TRY(cudaMemcpyAsync(out, ob.p, out_bytes, cudaMemcpyDeviceToHost, c->stream));
        return rsr_synchronize(c);
    }
    const size_t in_row = (size_t)w * in_px, out_row = (size_t)w * F * out_px;
    // band b owns LR rows [y0, y1); its halo rows are rows of the same input (none at the frame's own edges)
    auto band_rows = [&](int b, int& y0, int& y1) {
        y0 = r0 + (int)((long long)b * (r1 - r0) / nb);
        y1 = r0 + (int)((long long)(b + 1) * (r1 - r0) / nb);
    };
    // every band's forward pass is queued first and the copies after them: a copy into PAGEABLE host memory blocks the calling
    // thread until it is done, and the bands behind it should be computing meanwhile
    for (int b = 0; b < nb; ++b) {
        int y0, y1;
        band_rows(b, y0, y1);
        const int t = std::min(y0, kHaloRows), bt = std::min(h - y1, kHaloRows);
        uint8_t* d_band = (uint8_t*)ob.p + (size_t)(y0 - r0) * F * out_row;
        if (int rc = forward_device(c, (const uint8_t*)c->in_stage.p + (size_t)(y0 - t) * in_row, u8, in_ch, 1, w, y1 - y0 + t + bt,
                                    spec(d_band, t, bt)))
            return rc;
        CUDA_TRY(cudaEventRecord(c->ev_band[b], c->stream));
    }
    for (int b = 0; b < nb; ++b) {
        int y0, y1;
        band_rows(b, y0, y1);
        const size_t off = (size_t)(y0 - r0) * F * out_row;
        CUDA_TRY(cudaStreamWaitEvent(c->s_d2h, c->ev_band[b], 0));
        CUDA_TRY(cudaMemcpyAsync((uint8_t*)out + off, (const uint8_t*)ob.p + off, (size_t)(y1 - y0) * F * out_row, cudaMemcpyDeviceToHost, c->s_d2h));
    }
    if (int rc = rsr_synchronize(c)) return rc;       // kernel faults surface here, with the fault word
    cudaError_t e = cudaStreamSynchronize(c->s_d2h);
    if (e != cudaSuccess) return fail(RSR_E_CUDA, "device->host copy failed: %s", cudaGetErrorString(e));
    return 0;
}

// ---- pipelined host API ---------------------------------------------------------------------
static int submit_host(rsr_ctx* c, const void* in, bool u8, int in_ch, int w, int h, void* out, int out_ch) {
    if (!c) return fail(RSR_E_BAD_ARG, "ctx is null");
    if (!in || !out) return fail(RSR_E_BAD_ARG, "null buffer");
    if (w < 1 || h < 1) return fail(RSR_E_BAD_ARG, "bad dimensions w=%d h=%d", w, h);
    if (u8 && in_ch != 3 && in_ch != 4) return fail(RSR_E_BAD_ARG, "in_channels must be 3 or 4 (got %d)", in_ch);
    CUDA_TRY(cudaSetDevice(c->device));
    const long long ticket = c->next_ticket;
    rsr_ctx::Slot& sl = c->slot[ticket % 3];
    const int F = c->factor;
    const size_t npx = (size_t)w * h, opx = npx * F * F;
    const size_t in_bytes = u8 ? npx * in_ch : npx * 3 * sizeof(float);
    const size_t out_bytes = opx * (out_ch ? (size_t)out_ch : 3 * sizeof(float));
    if (in_bytes > sl.in.cap || out_bytes > sl.out.cap) {
        // growing a slot re-allocates: drain whatever still uses it first
        if (sl.used) CUDA_TRY(cudaEventSynchronize(sl.ev_out));
        if (int rc = ensure(sl.in, in_bytes)) return rc;
        if (int rc = ensure(sl.out, out_bytes)) return rc;
    }
    // H2D may overwrite the slot's input only after the forward pass that read it has finished
    if (sl.used) CUDA_TRY(cudaStreamWaitEvent(c->s_h2d, sl.ev_done, 0));
    CUDA_TRY(cudaMemcpyAsync(sl.in.p, in, in_bytes, cudaMemcpyHostToDevice, c->s_h2d));
    CUDA_TRY(cudaEventRecord(sl.ev_in, c->s_h2d));
    CUDA_TRY(cudaStreamWaitEvent(c->stream, sl.ev_in, 0));
    // the forward pass may overwrite the slot's output only after its previous D2H has finished
    if (sl.used) CUDA_TRY(cudaStreamWaitEvent(c->stream, sl.ev_out, 0));
    if (int rc = forward_device(c, sl.in.p, u8, in_ch, 1, w, h,
                                out_spec(out_ch ? nullptr : (float*)sl.out.p, out_ch ? (uint8_t*)sl.out.p : nullptr, out_ch ? out_ch : 4)))
        return rc;
    CUDA_TRY(cudaEventRecord(sl.ev_done, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->s_d2h, sl.ev_done, 0));
    CUDA_TRY(cudaMemcpyAsync(out, sl.out.p, out_bytes, cudaMemcpyDeviceToHost, c->s_d2h));
    CUDA_TRY(cudaEventRecord(sl.ev_out, c->s_d2h));
    sl.used = true;
    c->next_ticket++;
    return (int)(ticket & 0x3fffffff);
}
extern "C" int rsr_submit_u8(rsr_ctx* c, const uint8_t* in, int w, int h, int in_channels, uint8_t* out_rgba) {
    return submit_host(c, in, true, in_channels, w, h, out_rgba, 4);
}
extern "C" int rsr_submit_u8_rgb(rsr_ctx* c, const uint8_t* in, int w, int h, int in_channels, uint8_t* out_rgb) {
    return submit_host(c, in, true, in_channels, w, h, out_rgb, 3);
}
extern "C" int rsr_submit_f32(rsr_ctx* c, const float* in, int w, int h, float* out) {
    return submit_host(c, in, false, 3, w, h, out, 0);
}
extern "C" int rsr_wait(rsr_ctx* c, int ticket) {
    if (!c) return fail(RSR_E_BAD_ARG, "ctx is null");
    const long long newest = c->next_ticket - 1;
    // tickets are handed out modulo 2^30: the one meant is the most recent ticket number congruent to `ticket`
    const long long t = newest - ((newest - (long long)ticket) & 0x3fffffff);
    if (ticket < 0 || ticket > 0x3fffffff || t < 0) return fail(RSR_E_BAD_ARG, "unknown ticket %d (%lld issued so far)", ticket, newest + 1);
    CUDA_TRY(cudaSetDevice(c->device));
    // Frame t used slot t % 3.  If the slot has been re-used since, its event belongs to a LATER frame of the same slot;
    // device->host copies are ordered on one stream, so that event completing implies frame t's copy has completed too
    // (conservative, never early).
    cudaError_t e = cudaEventSynchronize(c->slot[t % 3].ev_out);
    if (e != cudaSuccess)
        return fail(RSR_E_CUDA, "frame failed: %s (fault word 0x%x)", cudaGetErrorString(e), c->fault_host ? c->fault_host[0] : 0);
    return 0;
}
extern "C" int rsr_wait_all(rsr_ctx* c) {
    if (!c) return fail(RSR_E_BAD_ARG, "ctx is null");
    CUDA_TRY(cudaSetDevice(c->device));
    for (auto& sl : c->slot)
        if (sl.used) CUDA_TRY(cudaEventSynchronize(sl.ev_out));
    return rsr_synchronize(c);
}

extern "C" int rsr_upscale_f32(rsr_ctx* c, const float* in, int w, int h, float* out) {
    return upscale_host(c, in, false, 3, 1, w, h, out, 0);
}
extern "C" int rsr_upscale_u8(rsr_ctx* c, const uint8_t* in, int w, int h, int in_channels, uint8_t* out_rgba) {
    return upscale_host(c, in, true, in_channels, 1, w, h, out_rgba, 4);
}
extern "C" int rsr_upscale_u8_rgb(rsr_ctx* c, const uint8_t* in, int w, int h, int in_channels, uint8_t* out_rgb) {
    return upscale_host(c, in, true, in_channels, 1, w, h, out_rgb, 3);
}
extern "C" int rsr_upscale_batch_f32(rsr_ctx* c, const float* in, int n, int w, int h, float* out) {
    return upscale_host(c, in, false, 3, n, w, h, out, 0);
}
extern "C" int rsr_upscale_batch_u8(rsr_ctx* c, const uint8_t* in, int n, int w, int h, int in_channels,
                                    uint8_t* out_rgba) {
    return upscale_host(c, in, true, in_channels, n, w, h, out_rgba, 4);
}
extern "C" int rsr_upscale_band_f32(rsr_ctx* c, const float* in_rows, int w, int rows_in, int top_halo,
                                    int bottom_halo, float* out_rows) {
    return upscale_host(c, in_rows, false, 3, 1, w, rows_in, out_rows, 0, top_halo, bottom_halo);
}
extern "C" int rsr_upscale_band_u8(rsr_ctx* c, const uint8_t* in_rows, int w, int rows_in, int top_halo, int bottom_halo,
                                   int in_channels, uint8_t* out_rows_rgba) {
    return upscale_host(c, in_rows, true, in_channels, 1, w, rows_in, out_rows_rgba, 4, top_halo, bottom_halo);
}

extern "C" int rsr_set_kernel_timing(rsr_ctx* c, int enable) {
    if (!c) return fail(RSR_E_BAD_ARG, "ctx is null");
    c->kernel_timing = enable != 0;
    if (!enable) c->timed = false;
    return 0;
}
extern "C" int rsr_last_kernel_times(rsr_ctx* c, float* ms, int cap) {
    if (!c || !ms) return fail(RSR_E_BAD_ARG, "null argument");
    if (!c->timed) return 0;
    if (int rc = rsr_synchronize(c)) return rc;
    int nout = 0;
    for (int i = 1; i < K_COUNT && nout < cap; ++i, ++nout) CUDA_TRY(cudaEventElapsedTime(&ms[nout], c->ev[i], c->ev[i + 1]));
    return nout;
}
extern "C" int rsr_last_launch_count(rsr_ctx* c) { return c ? c->last_launches : 0; }
extern "C" const char* rsr_plan(rsr_ctx* c) {
    static thread_local std::string s;
    if (!c) return "";
    const bool fused = c->plan == 1 || (c->plan == 0 && (c->last_launches == 0 || c->fused_last));
    s = fused ? "4 launches: conv0 (from the image), conv1|2|3, conv5|6|7+conv8|9 fused, conv10+expand"
              : "5 launches: conv0 (from the image), conv1|2|3, conv5|6|7, conv8|9, conv10+expand";
    if (c->plan == 0) s += " (middle levels by image width)";
    s += c->prec == PREC_F16MX ? "; operands fp16 + mxfp8 (2 MMAs per K chunk)" : "; operands bf16x3 (3 MMAs per K chunk)";
    s += "; CTA pairs (cta_group::2):";
    bool any = false;
    if (c->pair_mask & 4) { s += " conv1|2|3"; any = true; }
    if ((c->pair_mask & 8) && fused) { s += std::string(any ? "," : " ") + "conv5|6|7+conv8|9 fused"; any = true; }
    if (!any) s += " none";
    return s.c_str();
}

// ------------------------------------------------------------------------------- test hooks
extern "C" int rsr_debug_stop_after(rsr_ctx* c, int k) {
    if (!c) return fail(RSR_E_BAD_ARG, "ctx is null");
    c->stop_after = k;
    return 0;
}
// one stored activation row (128 B, conv_strip.cuh storage formats) -> 32 floats
static void unpack_act_row(const uint8_t* row, int prec, float* out) {
    for (int ch = 0; ch < 32; ++ch) {
        const uint8_t* chunk = row + (ch >> 4) * 64;
        const int k = ch & 15;
        if (prec == PREC_F16MX) {
            const __nv_fp8_storage_t lo = chunk[32 + k];
            out[ch] = __half2float(reinterpret_cast<const __half*>(chunk)[k]) + __half2float(__half(__nv_cvt_fp8_to_halfraw(lo, __NV_E4M3))) / F16MX_SCALE;
        } else {
            const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(chunk);
            out[ch] = __bfloat162float(q[k]) + __bfloat162float(q[16 + k]);
        }
    }
}
extern "C" int rsr_debug_read_node(rsr_ctx* c, int which, float* out, size_t npx) {
    if (!c || !out) return fail(RSR_E_BAD_ARG, "null argument");
    if (int rc = rsr_synchronize(c)) return rc;
    DevBuf* b = which == 0 ? &c->actA : which == 1 ? &c->actB : which == 2 ? &c->partA : which == 3 ? &c->partB
              : which == 4 ? &c->partC : which == 5 ? &c->dbg_l2 : which == 6 ? &c->dbg_t3 : nullptr;
    if (!b || !b->p || b->cap < npx * 128) return fail(RSR_E_BAD_ARG, "node %d not available for %zu pixels", which, npx);
    if (which >= 2) {
        CUDA_TRY(cudaMemcpy(out, b->p, npx * 128, cudaMemcpyDeviceToHost));
        return 0;
    }
    std::vector<uint8_t> tmp(npx * 128);
    CUDA_TRY(cudaMemcpy(tmp.data(), b->p, npx * 128, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < npx; ++i) unpack_act_row(&tmp[i * 128], c->prec, out + i * 32);
    return 0;
}

extern "C" int rsr_debug_fused(rsr_ctx* c, int flags) {
    if (!c) return fail(RSR_E_BAD_ARG, "ctx is null");
    c->fused_dbg = flags;
    return 0;
}
// [grid][32] wait-cycle counters of every warp role of the fused launch (slot names: FP_* in conv_fused.cuh)
extern "C" int rsr_debug_read_fused_profile(rsr_ctx* c, long long* out, int cap_rows) {
    if (!c || !out || !c->dbg_prof.p) return fail(RSR_E_BAD_ARG, "fused profile not enabled");
    if (int rc = rsr_synchronize(c)) return rc;
    const int rows = cap_rows < 256 ? cap_rows : 256;
    CUDA_TRY(cudaMemcpy(out, c->dbg_prof.p, sizeof(long long) * rows * FPROF_N, cudaMemcpyDeviceToHost));
    return rows;
}

extern "C" int rsr_debug_profile(rsr_ctx* c, int enable) {
    if (!c) return fail(RSR_E_BAD_ARG, "ctx is null");
    CUDA_TRY(cudaSetDevice(c->device));
    if (enable && !c->prof) {
        CUDA_TRY(cudaMalloc((void**)&c->prof, sizeof(long long) * K_COUNT * 256 * 16));
        CUDA_TRY(cudaMemset(c->prof, 0, sizeof(long long) * K_COUNT * 256 * 16));
    } else if (!enable && c->prof) {
        if (int rc = rsr_synchronize(c)) return rc;
        cudaFree(c->prof);
        c->prof = nullptr;
    }
    return 0;
}
// out: [grid][16] counters of launch `kindex` of the most recent call; returns the grid size.
extern "C" int rsr_debug_read_profile(rsr_ctx* c, int kindex, long long* out, int cap_rows) {
    if (!c || !out || kindex < 1 || kindex >= K_COUNT || !c->prof) return fail(RSR_E_BAD_ARG, "profile not enabled");
    if (int rc = rsr_synchronize(c)) return rc;
    const int rows = c->prof_grid[kindex] < cap_rows ? c->prof_grid[kindex] : cap_rows;
    CUDA_TRY(cudaMemcpy(out, c->prof + (size_t)kindex * 256 * 16, sizeof(long long) * rows * 16, cudaMemcpyDeviceToHost));
    return rows;
}

// ------------------------------------------------------------------------------- .rsr codec
// bytevec `<Vec<f32>>::encode::<u32>` / `decode::<u32>` (call sites main.rs:138-152,194,213):
// u32 LE N | N x u32 LE element sizes (all 4) | N x f32 LE.  Pure host code, no context.
extern "C" size_t rsr_rsr_encoded_size(size_t n_params) { return 4 + 8 * n_params; }
extern "C" long long rsr_rsr_encode(const float* params, size_t n_params, uint8_t* out, size_t cap) {
    if (!params || !out) return fail(RSR_E_BAD_ARG, "null argument");
    if (n_params > 0xffffffffull) return fail(RSR_E_BAD_ARG, "too many parameters for a u32 length prefix");
    const size_t need = rsr_rsr_encoded_size(n_params);
    if (cap < need) return fail(RSR_E_BAD_ARG, "output buffer too small (%zu < %zu)", cap, need);
    auto put32 = [](uint8_t* q, uint32_t v) { q[0] = (uint8_t)v; q[1] = (uint8_t)(v >> 8); q[2] = (uint8_t)(v >> 16); q[3] = (uint8_t)(v >> 24); };
    put32(out, (uint32_t)n_params);
    for (size_t i = 0; i < n_params; ++i) put32(out + 4 + 4 * i, 4);
    for (size_t i = 0; i < n_params; ++i) {
        uint32_t bits;
        memcpy(&bits, params + i, 4);
        put32(out + 4 + 4 * n_params + 4 * i, bits);
    }
    return (long long)need;
}
// Returns the number of parameters (writing at most `cap` of them to `params`, which may be NULL to query the count).
extern "C" long long rsr_rsr_decode(const uint8_t* data, size_t len, float* params, size_t cap) {
    if (!data) return fail(RSR_E_BAD_ARG, "null argument");
    auto get32 = [](const uint8_t* q) { return (uint32_t)q[0] | (uint32_t)q[1] << 8 | (uint32_t)q[2] << 16 | (uint32_t)q[3] << 24; };
    if (len < 4) return fail(RSR_E_BAD_ARG, "Parameter file is malformed: shorter than its length prefix");
    const size_t n = get32(data);
    if (len != 4 + 8 * n) return fail(RSR_E_BAD_ARG, "Parameter file is malformed: %zu bytes for %zu parameters", len, n);
    for (size_t i = 0; i < n; ++i)
        if (get32(data + 4 + 4 * i) != 4) return fail(RSR_E_BAD_ARG, "Parameter file is malformed: element %zu has size %u", i, get32(data + 4 + 4 * i));
    if (params)
        for (size_t i = 0; i < n && i < cap; ++i) {
            const uint32_t bits = get32(data + 4 + 4 * n + 4 * i);
            memcpy(params + i, &bits, 4);
        }
    return (long long)n;
}

// ------------------------------------------------------------------ parameter-free CLI modes
// bilinear_net / downsample_net (network.rs:111-138): elementwise + small-stencil, HBM bound.
__device__ __forceinline__ float srgb_to_linear(float v) {
    return v <= 0.04045f ? v * (1.0f / 12.92f) : powf((v + 0.055f) * (1.0f / 1.055f), 2.4f);
}
__device__ __forceinline__ float linear_to_srgb(float v) {
    return v <= 0.0031308f ? v * 12.92f : 1.055f * powf(v, 1.0f / 2.4f) - 0.055f;
}
template <bool U8>
__device__ __forceinline__ float load_px(const void* in, size_t idx, int ch, int c) {
    return U8 ? (float)reinterpret_cast<const uint8_t*>(in)[idx * ch + c] / 255.0f
              : reinterpret_cast<const float*>(in)[idx * 3 + c];
}
template <bool U8>
__device__ __forceinline__ void store_px(void* out, size_t idx, float r, float g, float b) {
    if (U8) {
        uchar4 q;
        q.x = (unsigned char)(fminf(fmaxf(r, 0.f), 1.f) * 255.0f + 0.5f);
        q.y = (unsigned char)(fminf(fmaxf(g, 0.f), 1.f) * 255.0f + 0.5f);
        q.z = (unsigned char)(fminf(fmaxf(b, 0.f), 1.f) * 255.0f + 0.5f);
        q.w = 255;
        reinterpret_cast<uchar4*>(out)[idx] = q;
    } else {
        float* o = reinterpret_cast<float*>(out) + idx * 3;
        o[0] = r; o[1] = g; o[2] = b;
    }
}
template <bool U8>
__global__ void __launch_bounds__(256) bilinear_kernel(const void* __restrict__ in, int ch, int W, int H, int F, void* __restrict__ out) {
    const size_t n = (size_t)W * F * H * F;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int X = (int)(i % ((size_t)W * F)), Y = (int)(i / ((size_t)W * F));
        const float sy = ((float)Y + 0.5f) / (float)F - 0.5f, sx = ((float)X + 0.5f) / (float)F - 0.5f;
        const float fy = floorf(sy), fx = floorf(sx);
        const float ty = sy - fy, tx = sx - fx;
        const int y0 = min(max((int)fy, 0), H - 1), y1 = min(max((int)fy + 1, 0), H - 1);
        const int x0 = min(max((int)fx, 0), W - 1), x1 = min(max((int)fx + 1, 0), W - 1);
        float o[3];
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v00 = srgb_to_linear(load_px<U8>(in, (size_t)y0 * W + x0, ch, c));
            const float v01 = srgb_to_linear(load_px<U8>(in, (size_t)y0 * W + x1, ch, c));
            const float v10 = srgb_to_linear(load_px<U8>(in, (size_t)y1 * W + x0, ch, c));
            const float v11 = srgb_to_linear(load_px<U8>(in, (size_t)y1 * W + x1, ch, c));
            const float top = v00 + (v01 - v00) * tx, bot = v10 + (v11 - v10) * tx;
            o[c] = linear_to_srgb(top + (bot - top) * ty);
        }
        store_px<U8>(out, i, o[0], o[1], o[2]);
    }
}
template <bool U8>
__global__ void __launch_bounds__(256) downsample_kernel(const void* __restrict__ in, int ch, int W, int H, int F, void* __restrict__ out) {
    const int OW = (W + F - 1) / F, OH = (H + F - 1) / F;
    const size_t n = (size_t)OW * OH;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int X = (int)(i % OW), Y = (int)(i / OW);
        float acc[3] = {0.f, 0.f, 0.f};
        int cnt = 0;
        for (int dy = 0; dy < F; ++dy)
            for (int dx = 0; dx < F; ++dx) {
                const int y = Y * F + dy, x = X * F + dx;
                if (y < H && x < W) {
                    ++cnt;
                    #pragma unroll
                    for (int c = 0; c < 3; ++c) acc[c] += srgb_to_linear(load_px<U8>(in, (size_t)y * W + x, ch, c));
                }
            }
        const float inv = 1.0f / (float)cnt;
        store_px<U8>(out, i, linear_to_srgb(acc[0] * inv), linear_to_srgb(acc[1] * inv), linear_to_srgb(acc[2] * inv));
    }
}

template <bool U8, bool DOWN>
static int aux_run(int device, const void* in, int w, int h, int in_ch, int factor, void* out) {
    if (!in || !out) return fail(RSR_E_BAD_ARG, "null buffer");
    if (w < 1 || h < 1 || factor < 1 || factor > 16) return fail(RSR_E_BAD_ARG, "bad dimensions w=%d h=%d factor=%d", w, h, factor);
    if (U8 && in_ch != 3 && in_ch != 4) return fail(RSR_E_BAD_ARG, "in_channels must be 3 or 4 (got %d)", in_ch);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(RSR_E_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(RSR_E_BAD_ARG, "device %d out of range", device);
    CUDA_TRY(cudaSetDevice(device));
    const size_t ipx = (size_t)w * h;
    const size_t opx = DOWN ? (size_t)((w + factor - 1) / factor) * ((h + factor - 1) / factor) : ipx * factor * factor;
    const size_t in_bytes = U8 ? ipx * in_ch : ipx * 12, out_bytes = U8 ? opx * 4 : opx * 12;
    void *d_in = nullptr, *d_out = nullptr;
    CUDA_TRY(cudaMalloc(&d_in, in_bytes));
    cudaError_t e = cudaMalloc(&d_out, out_bytes);
    if (e != cudaSuccess) { cudaFree(d_in); return fail(RSR_E_OOM, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    e = cudaMemcpy(d_in, in, in_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const int blocks = (int)std::min<size_t>((opx + 255) / 256, 148 * 16);
        if (DOWN) downsample_kernel<U8><<<blocks, 256>>>(d_in, in_ch, w, h, factor, d_out);
        else      bilinear_kernel<U8><<<blocks, 256>>>(d_in, in_ch, w, h, factor, d_out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, out_bytes, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(RSR_E_CUDA, "%s", cudaGetErrorString(e));
    return 0;
}
extern "C" int rsr_bilinear_f32(int device, const float* in, int w, int h, int factor, float* out) {
    return aux_run<false, false>(device, in, w, h, 3, factor, out);
}
extern "C" int rsr_bilinear_u8(int device, const uint8_t* in, int w, int h, int in_channels, int factor, uint8_t* out) {
    return aux_run<true, false>(device, in, w, h, in_channels, factor, out);
}
extern "C" int rsr_downsample_f32(int device, const float* in, int w, int h, int factor, float* out) {
    return aux_run<false, true>(device, in, w, h, 3, factor, out);
}
extern "C" int rsr_downsample_u8(int device, const uint8_t* in, int w, int h, int in_channels, int factor, uint8_t* out) {
    return aux_run<true, true>(device, in, w, h, in_channels, factor, out);
}

// --------------------------------------------------------------- generic 3x3 convolution (config 5 sweep)
// Same kernel template; the input's Cin/32 channel blocks are stored as planes in the activation storage format
// ([cb][H][W][128 B], conv_strip.cuh) and accumulated in TMEM (ConvParams::n_cb), the output is fp32 [H][W][Cout]
// written 32 channels at a time by the RAW epilogue through a tensor map with a channel coordinate.
// Operands are fp16 + mxfp8 (2 MMAs per 16-channel chunk) for every width.  Output blocks of <= 64 channels use row-folded
// MMAs (an N <= 64 MMA is bound by its A-tile read, so the three vertical taps become one MMA); wider outputs run as blocks of
// 128 channels, one row per pass (N = 128 MMAs are tensor-bound; 2 x 128 accumulator columns + the scale-factor columns fit
// tensor memory, two rows would not).  With several input-channel blocks every (pass, block) segment loads its own window of
// rows, so the row ring holds two windows: the next segment's rows land while the current one's MMAs run.
//                        KH KW  RB NG R NR TPS NBS NPIN NOUT SHUF NEPI PAIR FOLD PREC
using CvN32  = ConvCfg<3, 3, 128, 1, 2, 6, 3, 3, 0, 2, 0, 4, false, true, PREC_F16MX>;
using CvN64  = ConvCfg<3, 3, 128, 2, 2, 8, 3, 2, 0, 2, 0, 4, false, true, PREC_F16MX>;
using CvN128 = ConvCfg<3, 3, 128, 4, 1, 6, 1, 4, 0, 2, 0, 4, false, false, PREC_F16MX>;
// the same on CTA pairs (conv_pair.cuh): each CTA holds half of every tap, so an N = 128 MMA reads 4 + 2 KB of operands per
// 64 tensor cycles instead of 4 + 4 (the shared-memory port's 128 B/cycle) and becomes tensor-bound
using CvP128 = ConvCfg<3, 3, 128, 4, 1, 6, 1, 4, 0, 2, 0, 4, true, false, PREC_F16MX, PREC_F16MX, false, true>;

#ifndef RSR_CONV_N128
#define RSR_CONV_N128 CvP128   // blocks of 128 output channels on CTA pairs (CvN128: single CTAs)
#endif

template <int PREC>
__global__ void __launch_bounds__(256) pack_planes_kernel(const float* __restrict__ x, int W, int H, int Ci,
                                                           uint8_t* __restrict__ planes) {
    // one thread = one pixel x one channel block of 32 -> its 128-byte row in storage format PREC
    const size_t npx = (size_t)W * H;
    const int ncb = Ci / 32;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx * ncb; i += (size_t)gridDim.x * blockDim.x) {
        const size_t px = i % npx;
        const int cb = (int)(i / npx);
        const float* src = x + px * Ci + cb * 32;
        float v[32];
        #pragma unroll
        for (int k = 0; k < 32; ++k) v[k] = src[k];
        uint32_t w[32];
        pack_split_row<PREC>(v, w);
        uint4* dst = reinterpret_cast<uint4*>(planes + ((size_t)cb * npx + px) * 128);
        #pragma unroll
        for (int k = 0; k < 8; ++k) dst[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    }
}

struct AuxMaps {
    PFN_cuTensorMapEncodeTiled_v12000 encode;
};
static int encode_map(PFN_cuTensorMapEncodeTiled_v12000 enc, CUtensorMap* m, CUtensorMapDataType dt, int rank, const void* base,
                      const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box) {
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(m, dt, rank, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, RSR_L2PROMO, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RSR_E_CUDA, "cuTensorMapEncodeTiled failed: %d", (int)r);
    return 0;
}

template <class C>
static int conv2d_block(PFN_cuTensorMapEncodeTiled_v12000 enc, int num_sms, cudaStream_t st, const uint8_t* planes,
                        const uint8_t* wpack, float* d_out, int W, int H, int Ci, int Co, int co0, int* fault) {
    constexpr bool PAIR = C::PAIR;
    ConvParams p{};
    p.W = W; p.H = H; p.NIMG = 1; p.row_begin = 0; p.row_end = H; p.fault = fault;
    p.n_cb = Ci / 32;
    {
        cuuint64_t dims[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)p.n_cb};
        cuuint64_t strides[3] = {128, (cuuint64_t)128 * W, (cuuint64_t)128 * W * H};
        cuuint32_t box[4] = {64, (cuuint32_t)C::BOXW, 1, 1};
        if (int rc = encode_map(enc, &p.tmA, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, planes, dims, strides, box)) return rc;
    }
    {
        cuuint64_t dims[2] = {64, (cuuint64_t)p.n_cb * 9 * C::NCOL};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)(PAIR ? C::HB : C::TPS * C::NCOL)};   // (pairs load their half of a tap at a time)
        if (int rc = encode_map(enc, &p.tmB, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, wpack, dims, strides, box)) return rc;
    }
    for (int g = 0; g < C::NG; ++g) {
        cuuint64_t dims[4] = {(cuuint64_t)Co, (cuuint64_t)W, (cuuint64_t)H, 1};
        cuuint64_t strides[3] = {(cuuint64_t)Co * 4, (cuuint64_t)Co * 4 * W, (cuuint64_t)Co * 4 * W * H};
        cuuint32_t box[4] = {32, 32, 1, 1};
        if (int rc = encode_map(enc, &p.tmOut[g], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d_out, dims, strides, box)) return rc;
        p.grp[g] = EpiGroup{EPI_RAW, 0, nullptr, nullptr, co0 + g * 32, 0};
    }
    p.n_strips = (W + 127) / 128;
    p.total_rowtiles = (long long)p.n_strips * H;
    if constexpr (PAIR) {
        const int ncl = (int)std::min<long long>(p.total_rowtiles, num_sms / 2);
        CUDA_TRY(cudaFuncSetAttribute(conv_pair_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        if (int rc = launch_kernel(conv_pair_kernel<C>, 2 * ncl, C::NTHREADS, C::SMEM_BYTES, st, true, false, p)) return rc;
    } else {
        const int grid = (int)std::min<long long>(p.total_rowtiles, num_sms);
        CUDA_TRY(cudaFuncSetAttribute(conv_strip_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        conv_strip_kernel<C><<<grid, C::NTHREADS, C::SMEM_BYTES, st>>>(p);
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

extern "C" int rsr_conv2d_3x3_f32(int device, const float* x, const float* weights, int w, int h, int cin, int cout,
                                  float* out, int iters, float* kernel_ms) {
    if (!x || !weights || !out) return fail(RSR_E_BAD_ARG, "null buffer");
    if (w < 1 || h < 1 || cin < 32 || cout < 32 || cin % 32 || cout % 32)
        return fail(RSR_E_UNSUPPORTED, "conv2d: cin and cout must be positive multiples of 32 (got %d, %d)", cin, cout);
    if (iters < 1) iters = 1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(RSR_E_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device < 0 || device >= ndev) return fail(RSR_E_BAD_ARG, "device %d out of range", device);
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(RSR_E_CUDA, "device is sm_%d%d; sm_100a code only", prop.major, prop.minor);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) return fail(RSR_E_CUDA, "driver lacks cuTensorMapEncodeTiled");
    auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);

    const size_t npx = (size_t)w * h;
    const int ncb = cin / 32;
    // output-channel blocks: the widest kernel (N = 128, 64, 32) that fits what is left
    struct Blk { int co0, n; };
    std::vector<Blk> blocks;
    for (int co0 = 0; co0 < cout;) {
        const int left = cout - co0;
        const int n = left >= 128 ? 128 : left >= 64 ? 64 : 32;
        blocks.push_back({co0, n});
        co0 += n;
    }
    const int prec = PREC_F16MX;
    // weights [cout][3][3][cin] -> per block [cb][tap block][n][128 B weight row]
    std::vector<std::vector<uint8_t>> packs(blocks.size());
    for (size_t b = 0; b < blocks.size(); ++b) {
        const int n = blocks[b].n, co0 = blocks[b].co0;
        auto& pk = packs[b];
        pk.assign(((size_t)ncb * 9 * n + (size_t)n * 4) * 128, 0);  // tail pad: last stage may over-read
        for (int o = 0; o < n; ++o)
            for (int t = 0; t < 9; ++t)
                for (int cb = 0; cb < ncb; ++cb) {
                    // blocks of <= 64 channels run row-folded MMAs: tap blocks stored bottom tap first, tb * 3 + (2 - ta)
                    const int blk = n <= 64 ? (t % 3) * 3 + (2 - t / 3) : t;
                    pack_weight_row(&weights[(((size_t)(co0 + o)) * 9 + t) * cin + cb * 32], &pk[(((size_t)cb * 9 + blk) * n + o) * 128], prec);
                }
    }
    float* d_x = nullptr; float* d_out = nullptr; uint8_t* d_planes = nullptr; int* d_fault = nullptr;
    std::vector<uint8_t*> d_w(blocks.size(), nullptr);
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(d_x); cudaFree(d_out); cudaFree(d_planes); cudaFree(d_fault);
        for (auto p : d_w) cudaFree(p);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (st) cudaStreamDestroy(st);
    };
#define CV_TRY(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { rc = fail(e__ == cudaErrorMemoryAllocation ? RSR_E_OOM : RSR_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); cleanup(); return rc; } } while (0)
    CV_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    CV_TRY(cudaEventCreate(&e0));
    CV_TRY(cudaEventCreate(&e1));
    CV_TRY(cudaMalloc((void**)&d_x, npx * cin * sizeof(float)));
    CV_TRY(cudaMalloc((void**)&d_out, npx * cout * sizeof(float)));
    CV_TRY(cudaMalloc((void**)&d_planes, npx * ncb * 128));
    CV_TRY(cudaMalloc((void**)&d_fault, 64 * sizeof(int)));   // fault code + per-warp records (ptx.cuh mbar_wait)
    CV_TRY(cudaMemsetAsync(d_fault, 0, 64 * sizeof(int), st));
    for (size_t b = 0; b < blocks.size(); ++b) {
        CV_TRY(cudaMalloc((void**)&d_w[b], packs[b].size()));
        CV_TRY(cudaMemcpyAsync(d_w[b], packs[b].data(), packs[b].size(), cudaMemcpyHostToDevice, st));
    }
    CV_TRY(cudaMemcpyAsync(d_x, x, npx * cin * sizeof(float), cudaMemcpyHostToDevice, st));
    {
        const int nblk = (int)std::min<size_t>((npx * ncb + 255) / 256, (size_t)prop.multiProcessorCount * 16);
        pack_planes_kernel<PREC_F16MX><<<nblk, 256, 0, st>>>(d_x, w, h, cin, d_planes);
        CV_TRY(cudaGetLastError());
    }
    for (int it = 0; it < iters + 1; ++it) {   // first round is a warm-up
        if (it == 1) CV_TRY(cudaEventRecord(e0, st));
        for (size_t b = 0; b < blocks.size(); ++b) {
            const int n = blocks[b].n, co0 = blocks[b].co0;
            if (n == 128)     rc = conv2d_block<RSR_CONV_N128>(enc, prop.multiProcessorCount, st, d_planes, d_w[b], d_out, w, h, cin, cout, co0, d_fault);
            else if (n == 64) rc = conv2d_block<CvN64>(enc, prop.multiProcessorCount, st, d_planes, d_w[b], d_out, w, h, cin, cout, co0, d_fault);
            else              rc = conv2d_block<CvN32>(enc, prop.multiProcessorCount, st, d_planes, d_w[b], d_out, w, h, cin, cout, co0, d_fault);
            if (rc) { cleanup(); return rc; }
        }
    }
    CV_TRY(cudaEventRecord(e1, st));
    CV_TRY(cudaMemcpyAsync(out, d_out, npx * cout * sizeof(float), cudaMemcpyDeviceToHost, st));
    {
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            int f = 0;
            cudaMemcpy(&f, d_fault, sizeof f, cudaMemcpyDeviceToHost);
            rc = fail(RSR_E_CUDA, "conv2d kernel failed: %s (fault word 0x%x)", cudaGetErrorString(e), f);
            cleanup();
            return rc;
        }
    }
    if (kernel_ms) {
        float ms = 0.f;
        CV_TRY(cudaEventElapsedTime(&ms, e0, e1));
        *kernel_ms = ms / iters;
    }
#undef CV_TRY
    cleanup();
    return 0;
}

// ------------------------------------------------------------------ several GPUs from ONE host process
// The reference CLI is a single process (main.rs:33-178); these entry points let such a host use every GPU of the box
// (SURVEY.md 8(b) "multi-GPU variants", 8(e)): one context per device, one host thread per device, no collective —
// image batches shard as independent units (image i -> device i mod G), one large frame shards as horizontal bands with
// 7-row input halos (overlap-recompute; bit-identical to the unsharded frame).  The one-process-per-GPU form of the same
// plans (torchrun, NCCL) lives in rusty_sr_b200/sharding.py and bench.py.
struct rsr_multi {
    std::vector<rsr_ctx*> ctx;
};

extern "C" rsr_multi* rsr_create_multi(const float* params, size_t n_params, int factor, const int* devices, int n_devices, int flags) {
    g_err.clear();
    g_err_code = 0;
    std::vector<int> dev;
    if (devices) {
        if (n_devices < 1) { fail(RSR_E_BAD_ARG, "n_devices must be >= 1 (got %d)", n_devices); return nullptr; }
        dev.assign(devices, devices + n_devices);
    } else {   // all visible devices (at most n_devices of them when n_devices > 0)
        int cnt = 0;
        cudaError_t e = cudaGetDeviceCount(&cnt);
        if (e != cudaSuccess || cnt < 1) {
            fail(RSR_E_CUDA, "no CUDA device: %s", e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
            return nullptr;
        }
        if (n_devices > 0 && n_devices < cnt) cnt = n_devices;
        for (int i = 0; i < cnt; ++i) dev.push_back(i);
    }
    rsr_multi* m = new (std::nothrow) rsr_multi;
    if (!m) { fail(RSR_E_OOM, "out of host memory"); return nullptr; }
    for (int d : dev) {
        rsr_ctx* c = rsr_create(params, n_params, factor, d, flags);
        if (!c) {   // keep rsr_create's error for the caller
            const std::string keep = g_err;
            const int code = g_err_code;
            for (rsr_ctx* x : m->ctx) rsr_destroy(x);
            delete m;
            g_err = keep;
            g_err_code = code;
            return nullptr;
        }
        m->ctx.push_back(c);
    }
    return m;
}

extern "C" void rsr_destroy_multi(rsr_multi* m) {
    if (!m) return;
    for (rsr_ctx* c : m->ctx) rsr_destroy(c);
    delete m;
}

extern "C" int rsr_multi_size(rsr_multi* m) { return m ? (int)m->ctx.size() : 0; }
extern "C" rsr_ctx* rsr_multi_ctx(rsr_multi* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

// Runs job(g) on one host thread per context; the first failure (lowest device slot) becomes the caller's error.
template <class Job>
static int multi_run(rsr_multi* m, Job job) {
    const int G = (int)m->ctx.size();
    std::vector<int> rc(G, 0);
    std::vector<std::string> msg(G);
    std::vector<std::thread> th;
    th.reserve(G);
    for (int g = 0; g < G; ++g)
        th.emplace_back([&, g]() {
            rc[g] = job(g);
            if (rc[g] != 0) msg[g] = g_err;   // thread-local in the worker: hand it to the caller
        });
    for (auto& t : th) t.join();
    for (int g = 0; g < G; ++g)
        if (rc[g] != 0) return fail(rc[g] < 0 ? rc[g] : RSR_E_CUDA, "device slot %d (cuda:%d): %s", g, m->ctx[g]->device, msg[g].c_str());
    return 0;
}

static int multi_batch(rsr_multi* m, const void* in, bool u8, int in_ch, int n, int w, int h, void* out) {
    if (!m || m->ctx.empty()) return fail(RSR_E_BAD_ARG, "multi context is null");
    if (!in || !out) return fail(RSR_E_BAD_ARG, "null buffer");
    if (n < 1 || w < 1 || h < 1) return fail(RSR_E_BAD_ARG, "bad dimensions n=%d w=%d h=%d", n, w, h);
    if (u8 && in_ch != 3 && in_ch != 4) return fail(RSR_E_BAD_ARG, "in_channels must be 3 or 4 (got %d)", in_ch);
    const int G = (int)m->ctx.size(), F = m->ctx[0]->factor;
    const size_t in_img = (size_t)w * h * (u8 ? (size_t)in_ch : 3 * sizeof(float));
    const size_t out_img = (size_t)w * h * F * F * (u8 ? 4 : 3 * sizeof(float));
    return multi_run(m, [&](int g) -> int {
        rsr_ctx* c = m->ctx[g];
        for (int i = g; i < n; i += G) {   // image i -> device slot i mod G; up to three frames in flight per device
            const uint8_t* src = (const uint8_t*)in + (size_t)i * in_img;
            uint8_t* dst = (uint8_t*)out + (size_t)i * out_img;
            const int t = u8 ? rsr_submit_u8(c, src, w, h, in_ch, dst) : rsr_submit_f32(c, (const float*)src, w, h, (float*)dst);
            if (t < 0) { rsr_wait_all(c); return t; }
        }
        return rsr_wait_all(c);
    });
}
extern "C" int rsr_multi_upscale_batch_u8(rsr_multi* m, const uint8_t* in, int n, int w, int h, int in_channels, uint8_t* out_rgba) {
    return multi_batch(m, in, true, in_channels, n, w, h, out_rgba);
}
extern "C" int rsr_multi_upscale_batch_f32(rsr_multi* m, const float* in, int n, int w, int h, float* out) {
    return multi_batch(m, in, false, 3, n, w, h, out);
}

static int multi_frame(rsr_multi* m, const void* in, bool u8, int in_ch, int w, int h, void* out) {
    if (!m || m->ctx.empty()) return fail(RSR_E_BAD_ARG, "multi context is null");
    if (!in || !out) return fail(RSR_E_BAD_ARG, "null buffer");
    if (w < 1 || h < 1) return fail(RSR_E_BAD_ARG, "bad dimensions w=%d h=%d", w, h);
    if (u8 && in_ch != 3 && in_ch != 4) return fail(RSR_E_BAD_ARG, "in_channels must be 3 or 4 (got %d)", in_ch);
    const int G = (int)m->ctx.size(), F = m->ctx[0]->factor;
    const int HALO = 7;   // receptive field of sr_net in rows: 2 + 2 + 1 + 1 + 1 (network.rs:33,60-72)
    const size_t in_row = (size_t)w * (u8 ? (size_t)in_ch : 3 * sizeof(float));
    const size_t out_row = (size_t)w * F * (u8 ? 4 : 3 * sizeof(float));
    return multi_run(m, [&](int g) -> int {
        // band g owns LR rows [y0, y1) (sizes differ by at most one row; slots beyond the row count stay idle)
        const int y0 = (int)((long long)g * h / G), y1 = (int)((long long)(g + 1) * h / G);
        if (y1 <= y0) return 0;
        const int top = y0 < HALO ? y0 : HALO, bot = h - y1 < HALO ? h - y1 : HALO;
        const uint8_t* src = (const uint8_t*)in + (size_t)(y0 - top) * in_row;
        uint8_t* dst = (uint8_t*)out + (size_t)y0 * F * out_row;
        const int rows_in = y1 - y0 + top + bot;
        return u8 ? rsr_upscale_band_u8(m->ctx[g], src, w, rows_in, top, bot, in_ch, dst)
                  : rsr_upscale_band_f32(m->ctx[g], (const float*)src, w, rows_in, top, bot, (float*)dst);
    });
}
extern "C" int rsr_multi_upscale_frame_f32(rsr_multi* m, const float* in, int w, int h, float* out) {
    return multi_frame(m, in, false, 3, w, h, out);
}
extern "C" int rsr_multi_upscale_frame_u8(rsr_multi* m, const uint8_t* in, int w, int h, int in_channels, uint8_t* out_rgba) {
    return multi_frame(m, in, true, in_channels, w, h, out_rgba);
}
