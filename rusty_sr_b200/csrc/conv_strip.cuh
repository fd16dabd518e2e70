// conv_strip.cuh — the implicit-GEMM convolution kernel of the rusty_sr forward pass on sm_100a.
//
// Replaces alumina's Convolution (+ Bias + BeLU + Expand + LinearInterp) ops that
// `graph.forward` executes for sr_net (reference src/network.rs:27-72, src/main.rs:171).
//
// Formulation ("by input"): one launch takes ONE input node (32 channels, 128 B per pixel in a split
// storage format, below) and computes EVERY convolution that reads it, as one GEMM with the output
// channels of those convolutions concatenated along N:
//     D[pixel, n] = sum_{tap, ci} A[pixel + tap, ci] * W[tap][n][ci]
//   M = 128 consecutive pixels of one image row, N = 32 * NG, K = taps * 32.
//
// Operand precision (template parameter PREC; SURVEY.md F8 — one bf16 / tf32 product misses the 1e-3 contract):
//   PREC_F16MX  (default of the library)  a*w ~= a_h*w_h + 2^-11 * ( e4m3(2^11 a_l) * e4m3(w) + e4m3(a) * e4m3(2^11 w_l) )
//               with x_h = fp16(x), x_l = x - x_h: per 16-channel K chunk ONE kind::f16 MMA (K = 16) plus ONE
//               kind::mxf8f6f4 block-scaled MMA (K = 32: both cross terms concatenated along K) whose UE8M0 scale
//               factor 2^-11 is applied by the tensor core, so everything lands in one fp32 accumulator.  Executed
//               cost 2 MMAs per chunk; error <= 2.5e-4 linear light on uniform noise, <= 5e-5 on natural images
//               (tools/precision_probe.py; tests/cuda/umma_mx_probe.cu pins the instruction semantics).
//               Pixel row (128 B) = two 64-byte K chunks of 16 channels: [16 x fp16 a_h | 16 x e4m3 2^11 a_l | 16 x e4m3 a];
//               weight row: per chunk [16 x fp16 w_h | 16 x e4m3 w | 16 x e4m3 2^11 w_l].
//   PREC_BF16X3 (RSR_FLAG_BF16X3, conv0's x-im2col input, the generic conv) a_hi*w_hi + a_mid*w_hi + a_hi*w_mid, all
//               bf16: three kind::f16 MMAs per chunk, error <= 8e-5.  Per 64-byte chunk: [16 x bf16 hi | 16 x bf16 mid].
//   Either way a 16-channel chunk is 64 contiguous bytes (first plane 32 B, second plane 32 B), so an epilogue can split a
//   pixel's channels over two warps that each own four consecutive 16-byte pieces of the row.
//
// Data movement: a CTA walks down a 128-pixel-wide column strip.  Input rows are TMA-loaded ONCE
// each into a ring of row slots in shared memory (box = 128 + KW - 1 pixels, SWIZZLE_128B,
// out-of-bounds = zero fill = `Padding::Same`); the KH*KW taps are NOT re-loaded: the tap
// (a, b) operand is the same shared-memory row slot (row + a) read through a UMMA descriptor whose
// start address is shifted by b pixels.  Weights stream per tap through a small TMA ring.
//
// Partial pre-activation sums that an earlier launch produced for the same node (alumina nodes
// sum all their writers: network.rs:61+65, :62+66+69, :67+70+72) travel as fp32 [pixel][32] tiles:
// a producer warp TMA-loads each 128-pixel tile into shared memory ahead of time and the epilogue
// adds it from there, so the epilogue never waits on global memory:
// TMEM -> registers (+ partial from smem) -> bias / BeLU / split -> swizzled staging -> TMA store.  R output rows are
// accumulated per pass in TMEM, double buffered, so the epilogue of pass j overlaps the MMAs of pass j + 1.
//
// Row folding (ConvCfg::FOLD; conv0 and conv10): at N <= 64 an MMA is bound by the shared-memory read of its 4 KB A tile
// (tests/cuda/umma_bench.cu: max(N/2, 32 + N/4) cycles).  Input row i of a pass window contributes to the output rows
// r = i - ta, ta = 0..KH-1; with the accumulators of the R rows of a pass adjacent in TMEM and the KH weight taps of one tap
// column stored bottom tap first, those contributions are ONE MMA with N = (rows hit) x NCOL: same tensor work, KH times
// fewer A reads (measured, profiles/r02a_fold_session.log: conv10 207 -> 172 us, conv0 83 -> 79 us).
//
// conv0 (ConvCfg::RAWIN) has no TMA input at all: four extra warps read the image (f32 or u8) and write the x-im2col'd,
// split rows straight into the row ring (raw_row_producer_role) — the 80 B/px tensor a separate packing launch used to
// write and conv0 to read back is gone (same-box A/B: pack 50-54 + conv0 79-84 us -> 96 us).
//
// Warp roles: warp 0 = input-row producer, warp 3 = weight producer, warp 2 = TMEM allocator +
// partial-sum producer, warp 1 = MMA issuer (one lane), warps 4.. = epilogue (4 or 8 warps per config;
// with 8, two warps per TMEM lane quarter take alternate (row, group) tiles of a pass, which hides the
// instruction latency of the ALU-heavy epilogues: conv0 -11 %, conv5|6|7 -6 %, conv10 -10 %, measured).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "ptx.cuh"

// Timing-ablation switches (tools/ablate.sh builds variants; results are then WRONG on purpose).
#ifndef RSR_ABL
#define RSR_ABL 0
#endif
#define ABL_NO_A_LOAD (RSR_ABL & 1)     // input rows: barrier protocol only, no TMA
#define ABL_NO_B_LOAD (RSR_ABL & 2)     // weights: barrier protocol only, no TMA
#define ABL_NO_PIN (RSR_ABL & 4)        // partial-sum tiles: no TMA (MMAs still issued)
#define ABL_NO_MMA (RSR_ABL & 8)        // no tcgen05.mma at all
#define ABL_NO_EPI_OUT (RSR_ABL & 16)   // epilogue: TMEM load only, no math / staging / store
#define ABL_NO_E2_STORE (RSR_ABL & 128)  // fused kernel, second-level epilogue: math kept, global stores dropped
#define ABL_NO_E1_MATH (RSR_ABL & 64)   // fused kernel, first-level epilogue: no bias / BeLU / split (raw bits into the l2 ring)

namespace rsr {

enum : int { PREC_BF16X3 = 0, PREC_F16MX = 1 };
constexpr float F16MX_SCALE = 2048.0f;          // 2^11: the fp8 "low" operands are stored pre-multiplied by this ...
constexpr uint32_t F16MX_SFA_BYTES = 0x74747474u;   // ... and the block-scaled MMA multiplies A rows by UE8M0 116 = 2^-11
constexpr uint32_t F16MX_SFB_BYTES = 0x7f7f7f7fu;   // B columns by UE8M0 127 = 2^0

// What happens to one 32-column group of the accumulator.
enum EpiKind : int {
    EPI_NONE = 0,
    EPI_ACT = 1,      // y = BeLU(acc + bias; slope); store split activation
    EPI_RAW = 2,      // store acc (+ partial) as fp32 [pixel][32]: partial sum for a later launch
    EPI_SHUFFLE = 3,  // acc + bias; depth-to-space + bilinear skip; store image
};

struct EpiGroup {
    int kind;
    int has_pin;          // a partial sum from an earlier launch is added to this group
    const float* bias;    // 32 floats (zero padded) or null
    const float* slope;   // 32 floats: BeLU p (EPI_ACT)
    int c0;               // channel offset of this group in the output tensor (TMA coordinate 0)
    int pad_;
};

struct ConvParams {
    CUtensorMap tmA;       // input node, rank 4: (RB/2 elems, W, H, NIMG), box (RB/2, BOXW, 1, 1)
    CUtensorMap tmB;       // packed weights, rank 2: (RB/2 elems, taps * NCOL), box (RB/2, NCOL)
    CUtensorMap tmPin[3];  // per group: fp32 partial-sum input, rank 4 (32, W, H, NIMG), box (32, 128, 1, 1)
    CUtensorMap tmOut[8];  // per group: output, rank 4, box (.., 32 px, 1, 1): split activation (64 x 2 B) or fp32 (32)
    int W, H, NIMG;
    int row_begin, row_end;  // output rows computed; [0, H) normally
    int n_strips;
    long long total_rowtiles;  // NIMG * n_strips * (row_end - row_begin)
    EpiGroup grp[8];
    int n_cb;           // input channel blocks of 32 accumulated per output (1 for the sr_net launches; the
                        // generic conv uses Cin / 32, planes stored as images img * n_cb + cb)
    const float4* xf4;  // EPI_SHUFFLE: input image as float4 [N][H][W] (bilinear skip)
    // ConvCfg::RAWIN (conv0): the image itself, f32 [N][H][W][3] or u8 [N][H][W][raw_ch]; the producer warps also leave
    // every pixel as a float4 in xf4_out for the last launch's bilinear skip
    const void* raw_in;
    int raw_u8, raw_ch;
    float4* xf4_out;
    float* out_f32;     // EPI_SHUFFLE: [N][fH][fW][3] (or null)
    uint8_t* out_u8;    // EPI_SHUFFLE: [N][fH][fW][out_u8_ch] (or null)
    int out_u8_ch;      // 4 = RGBA (alpha 255, data_to_img(..).to_rgba()), 3 = RGB (data_to_img alone)
    int out_row0;       // EPI_SHUFFLE: LR row that maps to output row 0 (band mode: the rows above are halo)
    int out_rows;       // EPI_SHUFFLE: LR rows per image in the output tensor
    int out_vec_ok;     // output rows are 16-byte aligned (W * factor % 4 == 0 and aligned base)
    int* fault;
    long long* prof;    // test hook: per-CTA wait-cycle counters [gridDim.x][16] (null = off)
};

// Wait on an mbarrier and, when profiling, charge the cycles to counter `slot`.
#ifndef RSR_PROFILE
#define RSR_PROFILE 0   // tools/ablate.sh builds the profiling variant with -DRSR_PROFILE=1
#endif
#define RSR_PROF_ON (RSR_PROFILE && p.prof)
#define RSR_WAIT(bar, par, tag, slot)                                   \
    do {                                                                 \
        if (RSR_PROF_ON) {                                               \
            const long long t0__ = clock64();                            \
            mbar_wait(bar, par, p.fault, tag);                           \
            prof_acc[slot] += clock64() - t0__;                          \
        } else {                                                         \
            mbar_wait(bar, par, p.fault, tag);                           \
        }                                                                \
    } while (0)

// SHUF_: 0 = ordinary launch, f > 0 = last launch of the net with upscaling factor f (Expand + LinearInterp epilogue).
// PREC_: operand format of the INPUT node and the weights; OFMT_: storage format of the activations this launch writes.
// RAWIN_: the input node is the IMAGE (conv0): four extra warps read the raw pixels and build the x-im2col'd, split rows
// straight into the shared-memory row ring (no TMA, no intermediate tensor in HBM).
template <int KH_, int KW_, int RB_, int NG_, int R_, int NR_, int TPS_, int NBS_, int NPIN_, int NOUT_, int SHUF_, int NEPI_ = 4,
          bool PAIR_ = false, bool FOLD_ = false, int PREC_ = PREC_BF16X3, int OFMT_ = PREC_, bool RAWIN_ = false, bool MCB_ = false>
struct ConvCfg {
    static constexpr bool RAWIN = RAWIN_;
    static constexpr bool MCB = MCB_;                                  // conv_pair_kernel: several input-channel blocks per pass (ConvParams::n_cb)
    static constexpr int NPROD = RAWIN ? 4 : 0;                        // raw-image producer warps (each builds every 4th row of the ring)
    static constexpr bool PAIR = PAIR_;                                // CTA-pair kernel (conv_pair.cuh): half a weight tap per CTA
    // Row-folded MMAs (see the MMA role): the KH vertical taps of one tap column become ONE MMA per input row whose N spans
    // the up to KH output rows that row contributes to.
    static constexpr bool FOLD = FOLD_;
    static constexpr int PREC = PREC_, OFMT = OFMT_;
    static constexpr int KH = KH_, KW = KW_, RB = RB_, NG = NG_, R = R_;
    static constexpr int NEPI = NEPI_;                                 // epilogue warps (4 or 8)
    static_assert(NEPI == 4 || NEPI == 8, "4 or 8 epilogue warps");
    static constexpr int NTHREADS = 128 + 32 * NEPI + 32 * NPROD;
    static constexpr int NBS = NBS_, NPIN = NPIN_, NOUT = NEPI == 8 ? 1 : NOUT_;   // staging tiles per epilogue warp
    static constexpr bool SHUF = SHUF_ > 0;
    static constexpr int FACTOR = SHUF_;
    static constexpr int NCOL = 32 * NG;
    static constexpr int PH = KH / 2, PW = KW / 2;
    static constexpr int TAPS = KH * KW;
    static constexpr int BOXW = 128 + KW - 1;                          // pixels per loaded row
    static constexpr int ROW_BYTES = BOXW * RB;                        // TMA bytes per row
    static constexpr int SLOT = (ROW_BYTES + 1023) / 1024 * 1024;      // ring slot pitch
    static constexpr int NR = NR_;                                     // input-row ring slots (>= R + 2*PH)
    static constexpr int TPS = TPS_;                                   // taps per weight stage
    static constexpr int NSTG = (TAPS + TPS - 1) / TPS;                // weight stages per pass
    static constexpr int HB = PAIR ? NCOL / 2 : NCOL;                  // weight rows of one tap held by this CTA
    static constexpr int TAP_BYTES = HB * RB;                          // weight bytes of one tap (in this CTA)
    static constexpr int BST = TPS * TAP_BYTES;                        // bytes per weight stage
    static constexpr int KC = RB / 64;                                 // 16-channel chunks per plane
    static constexpr int PIN_BYTES = 128 * 128;                        // one partial-sum tile
    static constexpr int ACC_COLS = 2 * R * NCOL;                      // double-buffered accumulators
    static constexpr int SF_COL = ACC_COLS;                            // PREC_F16MX: 16 scale-factor columns for A, 16 for B
    static constexpr int USED_COLS = ACC_COLS + (PREC == PREC_F16MX ? 32 : 0);
    static constexpr int TMEM_COLS = USED_COLS <= 32 ? 32 : USED_COLS <= 64 ? 64 : USED_COLS <= 128 ? 128
                                   : USED_COLS <= 256 ? 256 : 512;
    static constexpr int OFF_B = NR * SLOT;
    static constexpr int OFF_P = OFF_B + NBS * BST;
    static constexpr int OFF_O = OFF_P + NPIN * PIN_BYTES;
    static constexpr int RAW_SCRATCH = 2048;                           // RAWIN: scratch per producer warp (132 x 3 floats)
    static constexpr int OFF_RAW = OFF_O + NEPI * NOUT * 4096;
    static constexpr int OFF_BAR = OFF_RAW + NPROD * RAW_SCRATCH;
    static constexpr int NBAR = 2 * NR + 2 * NBS + 2 * (NPIN > 0 ? NPIN : 1) + 4;
    static constexpr int SMEM_BYTES = 1024 + OFF_BAR + NBAR * 8 + 16;
    static constexpr uint32_t SWZ = RB == 128 ? UMMA_SW_128B : UMMA_SW_64B;
    static_assert(USED_COLS <= 512, "accumulators (+ scale factors) exceed TMEM");
    static_assert(RB == 128 || RB == 64, "pixel row must be 128 or 64 bytes");
    static_assert(PREC == PREC_BF16X3 || RB == 128, "the F16MX pixel row is 128 bytes");
    static_assert(TAP_BYTES % 1024 == 0, "weight tap tile must keep 1024-byte alignment");
    static_assert(PAIR || TPS * NCOL <= 256, "weight stage exceeds the TMA box limit");   // (pairs load tap by tap)
    static_assert(HB % 8 == 0, "a CTA's share of a tap must be whole 8-row core-matrix groups");
    static_assert(NR >= R + 2 * PH, "row ring smaller than one pass window");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
    static_assert(!FOLD || (!PAIR && TPS == KH && KH * NCOL <= 256), "row folding: one weight stage = the KH taps of one tap column");
    static_assert(!SHUF || 3 * FACTOR * FACTOR <= NCOL, "expand channels must fit the accumulator groups");
    static_assert(!RAWIN || NR % 4 == 0, "a producer warp keeps to its own ring slots");
    static_assert(!RAWIN || (KW == 1 && KH == 5 && RB == 64 && PREC == PREC_BF16X3 && !PAIR), "raw input: the 5x5x3 conv0 on 64-byte x-im2col rows");
};

struct WorkItem {
    int img, x0, y_b, y_e, npass;
};

// Static range partition: the NIMG * n_strips * rows row-tiles are linearised strip-major and cut
// into gridDim.x equal ranges; a range that crosses a strip boundary becomes two work items.
struct ItemIter {
    long long cur, end;
    int rows;
};
__device__ __forceinline__ ItemIter item_begin(const ConvParams& p) {
    ItemIter it;
    it.rows = p.row_end - p.row_begin;
    it.cur = p.total_rowtiles * (long long)blockIdx.x / gridDim.x;
    it.end = p.total_rowtiles * (long long)(blockIdx.x + 1) / gridDim.x;
    return it;
}
template <class C>
__device__ __forceinline__ bool item_next(const ConvParams& p, ItemIter& it, WorkItem& w) {
    if (it.cur >= it.end) return false;
    const int col = (int)(it.cur / it.rows);
    const int y0 = (int)(it.cur % it.rows);
    const int n = (int)min((long long)(it.rows - y0), it.end - it.cur);
    w.img = col / p.n_strips;
    w.x0 = (col % p.n_strips) * 128;
    w.y_b = p.row_begin + y0;
    w.y_e = w.y_b + n;
    w.npass = (n + C::R - 1) / C::R;
    it.cur += n;
    return true;
}

// Packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2, one instruction for two lanes of IEEE fp32 math).  The epilogues are bound
// by instruction issue, not by the FP32 pipe, so the bias / BeLU / split arithmetic runs on pairs: same results bit for bit
// as the scalar forms above, ~1/3 fewer instructions per activation value.
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t r, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));   // one MUFU op, rel. error ~2^-23
    return r;
}
__device__ __forceinline__ float belu(float x, float p) {
    // alumina BeLU (network.rs:35,54-56): y = p*x + sqrt(x^2 + 1) - 1   
    return fmaf(p, x, sqrt_approx(fmaf(x, x, 1.0f)) - 1.0f);
}

// (x0, x1) <- BeLU(x0 + b0; p0), BeLU(x1 + b1; p1): the same operations in the same order as belu(x + b, p)
__device__ __forceinline__ void bias_belu2(float& x0, float& x1, float b0, float b1, float p0, float p1) {
    const uint64_t x = add2(pack2(x0, x1), pack2(b0, b1));
    float t0, t1;
    unpack2(fma2(x, x, pack2(1.0f, 1.0f)), t0, t1);
    const uint64_t s = add2(pack2(sqrt_approx(t0), sqrt_approx(t1)), pack2(-1.0f, -1.0f));
    unpack2(fma2(pack2(p0, p1), x, s), x0, x1);
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
// (lo, hi) floats -> packed bf16x2 with round-to-nearest-even, one instruction
__device__ __forceinline__ uint32_t cvt_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// split two floats into bf16 hi parts (packed) and bf16 mid parts (packed): v = hi + mid + O(2^-17 v)
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& mid) {
    hi = cvt_bf16x2(v0, v1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    float m0, m1;
    unpack2(sub2(pack2(v0, v1), pack2(h0, h1)), m0, m1);
    mid = cvt_bf16x2(m0, m1);
}
// (lo, hi) floats -> two e4m3 bytes (lo in bits 0-7), round to nearest even, saturating
__device__ __forceinline__ uint32_t cvt_e4m3x2(float lo, float hi) {
    uint16_t r;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
    return r;
}
// F16MX split of two floats: fp16 parts (packed), e4m3(2^11 * residual) (2 bytes), e4m3(value) (2 bytes)
__device__ __forceinline__ void split_pair_f16mx(float v0, float v1, uint32_t& h, uint32_t& lo8, uint32_t& a8) {
    const __half2 h2 = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h2);
    h = *reinterpret_cast<const uint32_t*>(&h2);
    float l0, l1;
    unpack2(mul2(sub2(pack2(v0, v1), pack2(hf.x, hf.y)), pack2(F16MX_SCALE, F16MX_SCALE)), l0, l1);
    lo8 = cvt_e4m3x2(l0, l1);
    a8 = cvt_e4m3x2(v0, v1);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// explicit shared-space accesses for the staging tiles (pointers derived from the aligned dynamic-smem base are generic to the compiler)
__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_shared_b32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// A pixel's 128-byte row as 8 chunks of 16 bytes, written at `row` (shared-memory address of the row inside a tile laid out
// with the 128-byte TMA swizzle: chunk j of row r sits at chunk j ^ (r & 7)); sw = r & 7.
__device__ __forceinline__ void store_row_swizzled(const uint32_t (&w)[32], uint32_t row, uint32_t sw) {
    #pragma unroll
    for (int j = 0; j < 8; ++j) st_shared_v4(row + ((j ^ sw) << 4), w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}
// 16 activation values (one K chunk) -> the 16 words of its 64 bytes in storage format OFMT:
// F16MX [8 words fp16 a_h | 4 words e4m3(2^11 a_l) | 4 words e4m3(a)], BF16X3 [8 words bf16 hi | 8 words bf16 mid].
template <int OFMT>
__device__ __forceinline__ void pack_split_chunk(const float* v, uint32_t* w) {
    if (OFMT == PREC_F16MX) {
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t la, lb, aa, ab;
            split_pair_f16mx(v[4 * i], v[4 * i + 1], w[2 * i], la, aa);
            split_pair_f16mx(v[4 * i + 2], v[4 * i + 3], w[2 * i + 1], lb, ab);
            w[8 + i] = la | (lb << 16);
            w[12 + i] = aa | (ab << 16);
        }
    } else {
        #pragma unroll
        for (int i = 0; i < 8; ++i) split_pair(v[2 * i], v[2 * i + 1], w[i], w[8 + i]);
    }
}
// One pixel's 32 activation values -> the 32 words of its row (zero: a padding pixel, all zeros).
template <int OFMT>
__device__ __forceinline__ void pack_split_row(const float (&v)[32], uint32_t (&w)[32], bool zero = false) {
    pack_split_chunk<OFMT>(&v[0], &w[0]);
    pack_split_chunk<OFMT>(&v[16], &w[16]);
    if (zero) {
        #pragma unroll
        for (int i = 0; i < 32; ++i) w[i] = 0u;
    }
}

// Half rows (16 channels = one K chunk = the row's 16-byte pieces 4h .. 4h+3), for epilogues that split a pixel's 32
// channels over two warps.
template <int OFMT>
__device__ __forceinline__ void pack_split_half(const float (&v)[16], uint32_t (&w)[16], bool zero = false) {
    pack_split_chunk<OFMT>(&v[0], &w[0]);
    if (zero) {
        #pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = 0u;
    }
}
// into a tile of 128-byte rows with the 128-byte swizzle (sw = row & 7)
__device__ __forceinline__ void store_half_row_swizzled(const uint32_t (&w)[16], uint32_t row, uint32_t sw, int h) {
    #pragma unroll
    for (int k = 0; k < 4; ++k)
        st_shared_v4(row + (((uint32_t)(4 * h + k) ^ sw) << 4), w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
}
// into a tile of 64-byte rows with the 64-byte TMA swizzle (16-byte piece j of row r sits at piece j ^ ((r >> 1) & 3))
__device__ __forceinline__ void store_row64_swizzled(const uint32_t (&w)[16], uint32_t tile, int r) {
    const uint32_t row = tile + r * 64, sw = (r >> 1) & 3;
    #pragma unroll
    for (int k = 0; k < 4; ++k) st_shared_v4(row + (((uint32_t)k ^ sw) << 4), w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
}
// v += channels 16h .. 16h+15 of the pixel's fp32 partial sums in a TMA-loaded (128 B swizzled) tile of 128 pixels
__device__ __forceinline__ void epi_add_partial_half(float (&v)[16], uint32_t tile_addr, int row, int h) {
    const uint32_t base = tile_addr + row * 128;
    const uint32_t sw = row & 7;
    #pragma unroll
    for (int j = 0; j < 4; ++j) {
        float4 t;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(base + (((uint32_t)(4 * h + j) ^ sw) << 4)));
        v[4 * j + 0] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
    }
}

// One thread = one pixel; v[] = its 32 accumulator columns.  Applies bias / BeLU, splits every value into the storage
// format and writes the pixel's 128-byte row into the warp's staging tile.
template <int OFMT>
__device__ __forceinline__ void epi_split_to_stage(const EpiGroup& g, float (&v)[32], uint32_t stage_addr, int lane) {
    // (every EPI_ACT group has a bias and a slope vector: network.rs:35,50-56)
    const float4* b4 = reinterpret_cast<const float4*>(g.bias);
    const float4* s4 = reinterpret_cast<const float4*>(g.slope);
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 b = __ldg(b4 + i), t = __ldg(s4 + i);
        bias_belu2(v[4 * i + 0], v[4 * i + 1], b.x, b.y, t.x, t.y);
        bias_belu2(v[4 * i + 2], v[4 * i + 3], b.z, b.w, t.z, t.w);
    }
    uint32_t w[32];
    pack_split_row<OFMT>(v, w);
    store_row_swizzled(w, stage_addr + lane * 128, lane & 7);
}

// 3x3 neighbourhood of the input pixel, edge clamped, for the LinearInterp skip (network.rs:27).
struct SkipNb {
    float v[3][3][3];
};
// What one lane fetches of it: its own pixel column (rows y-1, y, y+1) — the left / right columns come from the neighbouring
// lanes by shuffle — plus, in lanes 0 and 31 only, the column just outside the warp's 32 pixels.  6 load instructions per
// tile instead of 9, few enough registers (18) to prefetch a whole pass ahead.
struct SkipCol {
    float v[3][3], e[3][3];
};
__device__ __forceinline__ void load_skip_col(const ConvParams& p, int img, int y, int x, int lane, SkipCol& c) {
    const int W = p.W, H = p.H;
    const int ys[3] = {max(y - 1, 0), y, min(y + 1, H - 1)};
    const int xo = min(x, W - 1);                                        // lanes past the image edge clamp (they serve as
    const int xe = lane == 0 ? min(max(x - 1, 0), W - 1) : min(x + 1, W - 1);   // right neighbour of the last pixel; a warp whose
                                                                          // whole slice lies past the edge still loads, in bounds)
    const float4* xi = p.xf4 + (size_t)img * H * W;
    #pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float4 t = __ldg(xi + (size_t)ys[a] * W + xo);
        c.v[a][0] = t.x; c.v[a][1] = t.y; c.v[a][2] = t.z;
    }
    if (lane == 0 || lane == 31) {
        #pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float4 t = __ldg(xi + (size_t)ys[a] * W + xe);
            c.e[a][0] = t.x; c.e[a][1] = t.y; c.e[a][2] = t.z;
        }
    }
}
// whole warp
__device__ __forceinline__ void expand_skip(const SkipCol& c, int lane, SkipNb& nb) {
    #pragma unroll
    for (int a = 0; a < 3; ++a)
        #pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float m = c.v[a][ch];
            float l = __shfl_up_sync(0xffffffffu, m, 1), r = __shfl_down_sync(0xffffffffu, m, 1);
            if (lane == 0) l = c.e[a][ch];
            if (lane == 31) r = c.e[a][ch];
            nb.v[a][0][ch] = l; nb.v[a][1][ch] = m; nb.v[a][2][ch] = r;
        }
}

// LinearInterp (network.rs:27): half-pixel centres.  Sub-pixel d of F sits at s = (d + 0.5) / F - 0.5 relative to its LR pixel:
// s < 0 blends neighbours (-1, 0) with weight s + 1 on the pixel itself, else (0, +1) with weight s on the next one.
template <int F>
struct Lerp {
    __host__ __device__ static constexpr int first(int d) { return 2 * d + 1 < F ? 0 : 1; }   // index into the 3-wide neighbourhood
    __host__ __device__ static constexpr float t(int d) {
        return (float)(2 * d + 1 < F ? ((double)d + 0.5) / F + 0.5 : ((double)d + 0.5) / F - 0.5);
    }
};

// Expand (depth-to-space, network.rs:39: out[F*h+dy, F*w+dx, c] = e[h, w, (dy*F+dx)*3 + c]) plus the bilinear skip, for one
// output row dy of one LR pixel: o[dx*3 + c].
template <int F, int NV>
__device__ __forceinline__ void shuffle_row(const float (&v)[NV], const SkipNb& nb, int dy, float (&o)[3 * F]) {
    const int ya = Lerp<F>::first(dy);
    const float ty = Lerp<F>::t(dy);
    #pragma unroll
    for (int dx = 0; dx < F; ++dx) {
        const int xa = Lerp<F>::first(dx);
        const float tx = Lerp<F>::t(dx);
        #pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v00 = nb.v[ya][xa][c], v01 = nb.v[ya][xa + 1][c];
            const float v10 = nb.v[ya + 1][xa][c], v11 = nb.v[ya + 1][xa + 1][c];
            const float top = v00 + (v01 - v00) * tx;
            const float bot = v10 + (v11 - v10) * tx;
            o[dx * 3 + c] = v[(dy * F + dx) * 3 + c] + top + (bot - top) * ty;
        }
    }
}

__device__ __forceinline__ uint32_t quant_u8(float v) {   // data_to_img (main.rs:175): round(clamp(v, 0, 1) * 255)
    return (uint32_t)(fminf(fmaxf(v, 0.f), 1.f) * 255.0f + 0.5f);
}

// The image store of the last launch for one warp = 32 LR pixels of LR row y: F output rows of 32 * F pixels each, staged
// in the warp's 4 KB tile `stg` and written with coalesced (vector where aligned) stores as f32 RGB, u8 RGBA or u8 RGB.
template <int F, int NV>
__device__ __forceinline__ void epi_shuffle_store(const ConvParams& p, const float (&v)[NV], const SkipNb& nb, uint32_t stg, int img,
                                                  int y, int xw, int lane) {
    const int npx_valid = min(32, p.W - xw);
    const size_t OW = (size_t)p.W * F;                                              // output pixels per row
    const size_t opx0 = ((size_t)img * p.out_rows * F + (size_t)(y - p.out_row0) * F) * OW + (size_t)xw * F;
    if (p.out_f32) {
        constexpr int DYC = F * F * 384 <= 4096 ? F : 4096 / (F * 384);              // output rows staged at once
        #pragma unroll
        for (int dy0 = 0; dy0 < F; dy0 += DYC) {
            #pragma unroll
            for (int dd = 0; dd < DYC; ++dd) {
                if (dy0 + dd >= F) continue;
                float o[3 * F];
                shuffle_row<F, NV>(v, nb, dy0 + dd, o);
                #pragma unroll
                for (int k = 0; k < 3 * F; ++k) st_shared_b32(stg + 4 * (dd * (96 * F) + lane * (3 * F) + k), __float_as_uint(o[k]));
            }
            __syncwarp();
            #pragma unroll
            for (int dd = 0; dd < DYC; ++dd) {
                if (dy0 + dd >= F) continue;
                float* dst = p.out_f32 + (opx0 + (size_t)(dy0 + dd) * OW) * 3;
                if (npx_valid == 32 && p.out_vec_ok) {
                    for (int i = lane; i < 24 * F; i += 32) reinterpret_cast<uint4*>(dst)[i] = ld_shared_v4(stg + 4 * (dd * (96 * F)) + 16 * i);
                } else {
                    for (int i = lane; i < npx_valid * 3 * F; i += 32) dst[i] = __uint_as_float(ld_shared_b32(stg + 4 * (dd * (96 * F) + i)));
                }
            }
            __syncwarp();
        }
    } else {
        // u8: stage RGBA words (alpha 255: data_to_img(..).to_rgba(), main.rs:175), [dy][32 * F pixels]
        #pragma unroll
        for (int dy = 0; dy < F; ++dy) {
            float o[3 * F];
            shuffle_row<F, NV>(v, nb, dy, o);
            #pragma unroll
            for (int dx = 0; dx < F; ++dx)
                st_shared_b32(stg + 4 * (dy * (32 * F) + lane * F + dx),
                              quant_u8(o[dx * 3]) | (quant_u8(o[dx * 3 + 1]) << 8) | (quant_u8(o[dx * 3 + 2]) << 16) | 0xFF000000u);
        }
        __syncwarp();
        const int nop = npx_valid * F;   // valid output pixels per row
        if (p.out_u8_ch == 4) {
            #pragma unroll
            for (int dy = 0; dy < F; ++dy) {
                uint32_t* dst = reinterpret_cast<uint32_t*>(p.out_u8) + opx0 + (size_t)dy * OW;
                for (int i = lane; i < nop; i += 32) dst[i] = ld_shared_b32(stg + 4 * (dy * (32 * F) + i));
            }
        } else {
            // packed RGB: 4 pixels = 3 words; word j takes its bytes from staged pixels 4(j/3) + j%3 and the next one
            #pragma unroll
            for (int dy = 0; dy < F; ++dy) {
                uint8_t* dst = p.out_u8 + (opx0 + (size_t)dy * OW) * 3;
                const uint32_t src = stg + 4 * (dy * (32 * F));
                if (npx_valid == 32 && p.out_vec_ok) {
                    for (int j = lane; j < 24 * F; j += 32) {
                        const int q = j / 3, m = j - 3 * q;
                        const uint32_t w0 = ld_shared_b32(src + 4 * (4 * q + m)), w1 = ld_shared_b32(src + 4 * (4 * q + m + 1));
                        reinterpret_cast<uint32_t*>(dst)[j] = __byte_perm(w0, w1, m == 0 ? 0x4210 : m == 1 ? 0x5421 : 0x6542);
                    }
                } else {
                    for (int i = lane; i < nop * 3; i += 32) dst[i] = (uint8_t)(ld_shared_b32(src + 4 * (i / 3)) >> (8 * (i % 3)));
                }
            }
        }
        __syncwarp();
    }
}

// fp32 [pixel][32] row (128 B) into the warp's staging tile, 128-byte swizzle.
__device__ __forceinline__ void epi_raw_to_stage(const float (&v)[32], uint32_t stage_addr, int lane) {
    uint32_t w[32];
    #pragma unroll
    for (int i = 0; i < 32; ++i) w[i] = __float_as_uint(v[i]);
    store_row_swizzled(w, stage_addr + lane * 128, lane & 7);
}

// v += the pixel's 32 fp32 partial sums from a TMA-loaded (128 B swizzled) tile of 128 pixels.
__device__ __forceinline__ void epi_add_partial_tile(float* v, uint32_t tile_addr, int row) {
    const uint32_t base = tile_addr + row * 128;
    const uint32_t sw = row & 7;
    #pragma unroll
    for (int j = 0; j < 8; ++j) {
        float4 t;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(base + ((j ^ sw) << 4)));
        v[4 * j + 0] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
    }
}

// ---- MMA issue: the split products of one 16-channel K chunk ---------------------------------------------------
// a_e / b_e: encoded (>> 4) shared-memory addresses of the pixel row / weight row starts (with the LBO bits);
// chunk c lives at +64c bytes (encoded +4c), its second plane (bf16 mid / fp8 pair) 32 bytes further (encoded +2).
template <int PREC, int CG, int RB>
struct SplitMma {
    static constexpr uint32_t HI = (uint32_t)make_sdesc_hi(8 * RB, RB == 128 ? UMMA_SW_128B : UMMA_SW_64B);
    static constexpr uint32_t PLANE2 = 2, CHUNK = 4;
    static __device__ __forceinline__ uint64_t desc(uint32_t lo) { return ((uint64_t)HI << 32) | lo; }
    static __host__ __device__ constexpr uint32_t idesc(int M, int N) {
        return PREC == PREC_F16MX ? make_idesc_f16(M, N) : make_idesc_bf16(M, N);
    }
    // the leading product (hi x hi): the only one that may overwrite the accumulator
    static __device__ __forceinline__ void head(uint32_t d, uint32_t a_e, uint32_t b_e, int c, int M, int N, uint32_t acc) {
        umma_f16_cg<CG>(d, desc(a_e + CHUNK * c), desc(b_e + CHUNK * c), idesc(M, N), acc);
    }
    // the correction products, always accumulated; sf = TMEM address of the scale-factor columns (F16MX)
    static __device__ __forceinline__ void tail(uint32_t d, uint32_t a_e, uint32_t b_e, int c, int M, int N, uint32_t sf) {
        if (PREC == PREC_F16MX) {
            umma_mxf8<CG>(d, desc(a_e + CHUNK * c + PLANE2), desc(b_e + CHUNK * c + PLANE2), make_idesc_mxf8(M, N), sf, sf + 16, 1);
        } else {
            umma_f16_cg<CG>(d, desc(a_e + CHUNK * c + PLANE2), desc(b_e + CHUNK * c), idesc(M, N), 1);
            umma_f16_cg<CG>(d, desc(a_e + CHUNK * c), desc(b_e + CHUNK * c + PLANE2), idesc(M, N), 1);
        }
    }
    static __device__ __forceinline__ void chunk(uint32_t d, uint32_t a_e, uint32_t b_e, int c, int M, int N, uint32_t sf, uint32_t acc) {
        head(d, a_e, b_e, c, M, N, acc);
        tail(d, a_e, b_e, c, M, N, sf);
    }
};

// PREC_F16MX: the four epilogue warps 4..7 (one per TMEM lane quarter) fill the scale-factor columns; call from ALL threads
// right after the TMEM base address is known (contains a CTA barrier).
template <class C>
__device__ __forceinline__ void init_scale_factors(uint32_t tmem_base, int warp) {
    if (C::PREC == PREC_F16MX) {
        if (warp >= 4 && warp < 8) {
            const uint32_t t = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + C::SF_COL;
            tmem_fill_x16(t, F16MX_SFA_BYTES);
            tmem_fill_x16(t + 16, F16MX_SFB_BYTES);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
}

// ---- the epilogue role (shared by conv_strip_kernel and conv_pair_kernel) ------------------------------------------
// next_item(w): the CTA's next work item; arrive_t_empty(buf): "accumulator buffer drained" (local or cluster arrive).
template <class C, class NextItem, class ArriveTEmpty>
__device__ __forceinline__ void conv_epilogue_role(const ConvParams& p, uint8_t* smem, uint64_t* p_full, uint64_t* p_empty,
                                                   uint64_t* t_full, uint32_t tmem_base, int warp, int lane, long long* prof_acc,
                                                   NextItem next_item, ArriveTEmpty arrive_t_empty) {
    constexpr int NPB = C::NPIN > 0 ? C::NPIN : 1;
    uint8_t* sO = smem + C::OFF_O;
    const int e = warp - 4;   // epilogue warp index
    const int ew = e & 3;     // == warp % 4: TMEM lane quarter this warp may read
    const int eh = e >> 2;    // with 8 epilogue warps: which half of the pass's tiles this warp takes
    const uint32_t stage0 = smem_u32(sO) + e * (C::NOUT * 4096);
    uint32_t buf = 0, tpar = 0, oc = 0;
    uint32_t pst = 0, ppar = 0;  // partial-sum tile ring (consumed in producer order: row, then group)
    const uint32_t sP_u = smem_u32(smem + C::OFF_P);
    // observe (and, when `use`, add) the next partial-sum tile: every epilogue warp sees every tile (wait full, arrive empty);
    // a warp that skipped the wait could fall two phases behind the barrier and mistake parity for completion
    auto take_pin = [&](float* v, bool use) {
        RSR_WAIT(&p_full[pst], ppar, 9, 3);
        if (use) epi_add_partial_tile(v, sP_u + pst * C::PIN_BYTES, ew * 32 + lane);
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_empty[pst]);
        if (++pst == NPB) { pst = 0; ppar ^= 1; }
    };
    auto wait_stage = [&]() {   // the staging tile used NOUT stores ago must have been read out by the TMA engine
        const long long t0s = RSR_PROF_ON ? clock64() : 0;
        if (elect_one()) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(C::NOUT - 1) : "memory");
        __syncwarp();
        if (RSR_PROF_ON) prof_acc[1] += clock64() - t0s;
    };
    WorkItem w;
    SkipCol next_col;
    bool have_next = false;
    while (next_item(w)) {
        const int xw = w.x0 + ew * 32;  // first pixel of this warp's 32-pixel slice
        const int x = xw + lane;
        have_next = false;
        for (int j = 0; j < w.npass; ++j, buf ^= 1, tpar ^= (buf == 0)) {
            SkipCol col;
            if (C::SHUF) {
                // the skip input does not depend on the MMAs: this warp's first row of the pass was fetched a pass ago (or is
                // fetched now, at the start of an item), and the next pass's row is requested before this pass's work
                // (SHUF launches: one tile per row; with 8 warps this warp owns rows r % 2 == eh)
                const int y = w.y_b + j * C::R + eh;
                if (have_next) col = next_col;
                else if (y < w.y_e) load_skip_col(p, w.img, y, x, lane, col);
                have_next = j + 1 < w.npass && y + C::R < w.y_e;
                if (have_next) load_skip_col(p, w.img, y + C::R, x, lane, next_col);
            }
            RSR_WAIT(&t_full[buf], tpar, 6, 0);
            tc_fence_after();
            const long long te0 = RSR_PROF_ON ? clock64() : 0;
            const uint32_t t_pass = tmem_base + ((uint32_t)(ew * 32) << 16) + buf * (C::R * C::NCOL);
            #pragma unroll 1
            for (int r = 0; r < C::R; ++r) {
                const int y = w.y_b + j * C::R + r;
                const bool row_live = y < w.y_e && xw < p.W;   // warp-uniform
                if constexpr (C::SHUF) {
                    // bias, depth-to-space, bilinear skip -> staging -> coalesced stores; all groups of the row together
                    constexpr int NV = 32 * C::NG;
                    const bool mine = C::NEPI == 4 || ((r & 1) == eh);
                    if (!mine || !row_live) {
                        #pragma unroll
                        for (int g = 0; g < C::NG; ++g)
                            if (C::NPIN > 0 && p.grp[g].has_pin) take_pin(nullptr, false);
                        continue;
                    }
                    float v[NV];
                    #pragma unroll
                    for (int g = 0; g < C::NG; ++g) {
                        uint32_t raw[32];
                        tmem_ld_32x32(t_pass + r * C::NCOL + g * 32, raw);
                        tmem_ld_wait();
                        #pragma unroll
                        for (int k = 0; k < 32; ++k) v[g * 32 + k] = __uint_as_float(raw[k]);
                        if (C::NPIN > 0 && p.grp[g].has_pin) take_pin(&v[g * 32], true);
                        const float4* b4 = reinterpret_cast<const float4*>(p.grp[g].bias);
                        #pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float4 t = __ldg(b4 + i);
                            v[g * 32 + 4 * i + 0] += t.x; v[g * 32 + 4 * i + 1] += t.y; v[g * 32 + 4 * i + 2] += t.z; v[g * 32 + 4 * i + 3] += t.w;
                        }
                    }
                    if (ABL_NO_EPI_OUT) { if (v[0] == 123.456f && p.out_f32) p.out_f32[0] = v[1]; ++oc; continue; }
                    const long long to0 = RSR_PROF_ON ? clock64() : 0;
                    if (r > eh) load_skip_col(p, w.img, y, x, lane, col);   // (4 epilogue warps only: further rows of the pass)
                    SkipNb nb;
                    expand_skip(col, lane, nb);
                    epi_shuffle_store<C::FACTOR, NV>(p, v, nb, stage0 + (oc % C::NOUT) * 4096, w.img, y, xw, lane);
                    if (RSR_PROF_ON) prof_acc[6] += clock64() - to0;
                    ++oc;
                } else {
                    #pragma unroll 1
                    for (int g = 0; g < C::NG; ++g) {
                        const EpiGroup& eg = p.grp[g];
                        if (eg.kind == EPI_NONE) continue;
                        const bool pin = C::NPIN > 0 && eg.has_pin;
                        const bool mine = C::NEPI == 4 || (((r * C::NG + g) & 1) == eh);   // warp-uniform tile ownership
                        if (!mine || !row_live) {
                            // the other warp of this lane quarter handles the tile (or nothing is stored): the partial
                            // tile the producer fetched must still be observed and released
                            if (pin) take_pin(nullptr, false);
                            continue;
                        }
                        uint32_t raw[32];
                        const long long tl0 = RSR_PROF_ON ? clock64() : 0;
                        tmem_ld_32x32(t_pass + r * C::NCOL + g * 32, raw);
                        tmem_ld_wait();
                        if (RSR_PROF_ON) prof_acc[5] += clock64() - tl0;
                        float v[32];
                        #pragma unroll
                        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(raw[k]);
                        if (pin) take_pin(v, true);
                        if (ABL_NO_EPI_OUT) { if (v[0] == 123.456f && p.out_f32) p.out_f32[0] = v[1]; ++oc; continue; }
                        const long long to0 = RSR_PROF_ON ? clock64() : 0;
                        const uint32_t stage = stage0 + (oc % C::NOUT) * 4096;
                        wait_stage();
                        if (eg.kind == EPI_RAW) epi_raw_to_stage(v, stage, lane);
                        else epi_split_to_stage<C::OFMT>(eg, v, stage, lane);
                        fence_proxy_async_smem();
                        __syncwarp();
                        if (elect_one()) {
                            asm volatile(
                                "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                                ::"l"(reinterpret_cast<uint64_t>(&p.tmOut[g])), "r"(stage), "r"(eg.c0), "r"(xw), "r"(y),
                                  "r"(w.img)
                                : "memory");
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        }
                        if (RSR_PROF_ON) prof_acc[6] += clock64() - to0;
                        ++oc;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (RSR_PROF_ON) prof_acc[2] += clock64() - te0;
            if (lane == 0) arrive_t_empty(buf);
        }
    }
    // all TMA stores of this warp must have completed before the CTA exits
    if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- raw-image row producer (ConvCfg::RAWIN: conv0 straight from the image) ------------------------------------------
// img_to_data (main.rs:170) and the layout change conv0 needs, done on the way into shared memory: for every pixel the 5
// horizontal neighbours x-2..x+2 (zero outside the image = Padding::Same) x 3 channels become one 16-element K chunk
// (k = dx*3 + c, k = 15 zero), split bf16 hi | mid = 64 B per pixel, written with the 64-byte swizzle exactly where a TMA
// load of an x-im2col tensor would have put it.  Producer warp pw of NPROD builds every NPROD-th row of the ring, all 128
// pixels of it (four per lane: the role is a chain of dependent latencies — global load, scratch round trip, split, proxy
// fence, barrier — so rows are spread over warps and every step of the chain carries four pixels): its lanes fetch the
// 132 x 3 values of the row segment with coalesced loads, issued two of its rows (= 2 NPROD ring rows) ahead, pass them
// through 1.6 KB of scratch so that a lane can pick up the 15 consecutive values of a pixel (stride 3 words: conflict-free),
// and arrive on the row's full barrier after a proxy fence.  Rows the CTA owns also leave each pixel as a float4 in xf4_out
// for the bilinear skip of the last launch.
template <class C>
__device__ __forceinline__ void raw_row_producer_role(const ConvParams& p, uint8_t* smem, uint64_t* a_full, uint64_t* a_empty,
                                                      int pw, int lane, long long* prof_acc) {
    constexpr int NE = 13;                       // 32-lane load rounds covering the 396 values of a row segment
    constexpr int PF = 2;                        // rows of this warp in flight
    const uint32_t scr = smem_u32(smem + C::OFF_RAW) + pw * C::RAW_SCRATCH;
    const uint32_t sA_u = smem_u32(smem);
    const int W = p.W, H = p.H;
    const float* in_f = reinterpret_cast<const float*>(p.raw_in);
    const uint8_t* in_b = reinterpret_cast<const uint8_t*>(p.raw_in);
    const bool u8 = p.raw_u8 != 0;
    const int rch = p.raw_ch;
    // values lane + 32 k (k < NE) of the segment that starts at pixel xs (may be negative) of row y; value e sits at position
    // 3 xs + e of the row's packed RGB values.  u8 / 255 is the IEEE quotient (as img_to_data's division) without the
    // division's slow path: one Newton step on x * (1/255) is exact for all 256 inputs (checked exhaustively).
    auto load_row = [&](int img, int y, int xs, float (&v)[NE]) {
        const bool rowok = y >= 0 && y < H;
        const size_t rowbase = ((size_t)img * H + (rowok ? y : 0)) * W;
        const int idx0 = xs * 3 + lane;
        const unsigned lim = rowok ? 3u * (unsigned)W : 0u;
        if (!u8) {
            const float* q = in_f + rowbase * 3;
            #pragma unroll
            for (int k = 0; k < NE; ++k) {
                const int idx = idx0 + 32 * k;
                const bool ok = (unsigned)idx < lim && (k < NE - 1 || lane < 12);
                v[k] = ok ? __ldg(q + idx) : 0.f;
            }
        } else {
            uint32_t b[NE];
            if (rch == 3) {
                const uint8_t* q = in_b + rowbase * 3;
                #pragma unroll
                for (int k = 0; k < NE; ++k) {
                    const int idx = idx0 + 32 * k;
                    const bool ok = (unsigned)idx < lim && (k < NE - 1 || lane < 12);
                    b[k] = ok ? (uint32_t)__ldg(q + idx) : 0u;
                }
            } else {
                const uint8_t* q = in_b + rowbase * 4;
                #pragma unroll
                for (int k = 0; k < NE; ++k) {
                    const int idx = idx0 + 32 * k;
                    const bool ok = (unsigned)idx < lim && (k < NE - 1 || lane < 12);
                    b[k] = ok ? (uint32_t)__ldg(q + idx + (unsigned)idx / 3u) : 0u;   // 4 px + c = (3 px + c) + px
                }
            }
            const float rcp = 1.0f / 255.0f;
            #pragma unroll
            for (int k = 0; k < NE; ++k) {
                const float x = (float)b[k], q0 = x * rcp;
                v[k] = fmaf(fmaf(-q0, 255.0f, x), rcp, q0);
            }
        }
    };
    uint32_t g_item = 0;                         // ring rows produced (by all producer warps) before the current item
    ItemIter it = item_begin(p);
    WorkItem w;
    while (item_next<C>(p, it, w)) {
        const int nrows = w.npass * C::R + 2 * C::PH;
        const int y0 = w.y_b - C::PH;
        const int xs = w.x0 - 2;
        const int i_first = (int)((pw - g_item) & (C::NPROD - 1));   // this warp's first row of the item
        float buf[PF][NE];
        #pragma unroll
        for (int j = 0; j < PF; ++j)
            if (i_first + j * C::NPROD < nrows) load_row(w.img, y0 + i_first + j * C::NPROD, xs, buf[j]);
        for (int i0 = i_first; i0 < nrows; i0 += PF * C::NPROD) {
            #pragma unroll
            for (int j = 0; j < PF; ++j) {
                const int i = i0 + j * C::NPROD, y = y0 + i;
                if (i >= nrows) break;           // (warp-uniform)
                #pragma unroll
                for (int k = 0; k < NE; ++k)
                    if (k < NE - 1 || lane < 12) st_shared_b32(scr + 4 * (lane + 32 * k), __float_as_uint(buf[j][k]));
                if (i + PF * C::NPROD < nrows) load_row(w.img, y + PF * C::NPROD, xs, buf[j]);
                __syncwarp();
                const uint32_t g = g_item + i, slot = g % C::NR, par = ((g / C::NR) & 1) ^ 1;
                RSR_WAIT(&a_empty[slot], par, 1, 0);
                const bool own = y >= w.y_b && y < w.y_e;        // rows of this item (not its halo): keep x for the skip
                float4* xrow = p.xf4_out + ((size_t)w.img * H + (own ? y : 0)) * W + w.x0;
                #pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = 32 * q + lane;                 // pixel of the strip = row of the slot
                    float v[16];
                    #pragma unroll
                    for (int k = 0; k < 15; ++k) v[k] = __uint_as_float(ld_shared_b32(scr + 4 * (3 * r + k)));
                    v[15] = 0.f;
                    uint32_t hi[8], mid[8];
                    #pragma unroll
                    for (int k = 0; k < 8; ++k) split_pair(v[2 * k], v[2 * k + 1], hi[k], mid[k]);
                    const uint32_t row = sA_u + slot * C::SLOT + r * 64, sw = (r >> 1) & 3;
                    st_shared_v4(row + ((0u ^ sw) << 4), hi[0], hi[1], hi[2], hi[3]);
                    st_shared_v4(row + ((1u ^ sw) << 4), hi[4], hi[5], hi[6], hi[7]);
                    st_shared_v4(row + ((2u ^ sw) << 4), mid[0], mid[1], mid[2], mid[3]);
                    st_shared_v4(row + ((3u ^ sw) << 4), mid[4], mid[5], mid[6], mid[7]);
                    if (own && w.x0 + r < W) xrow[r] = make_float4(v[6], v[7], v[8], 0.f);
                }
                fence_proxy_async_smem();        // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();                    // (also: the scratch is rewritten by this warp's next row)
                if (lane == 0) mbar_arrive(&a_full[slot]);
            }
        }
        g_item += nrows;
    }
}

// ---- the kernel ------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(C::NTHREADS, 1) conv_strip_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + C::OFF_B;
    uint8_t* sP = smem + C::OFF_P;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* a_empty = a_full + C::NR;
    uint64_t* b_full = a_empty + C::NR;
    uint64_t* b_empty = b_full + C::NBS;
    constexpr int NPB = C::NPIN > 0 ? C::NPIN : 1;
    uint64_t* p_full = b_empty + C::NBS;
    uint64_t* p_empty = p_full + NPB;
    uint64_t* t_full = p_empty + NPB;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long prof_t0 = RSR_PROF_ON ? clock64() : 0;

    // which groups fold a partial sum (uniform per launch)
    int pin_g[3], npg = 0;
    #pragma unroll
    for (int g = 0; g < C::NG; ++g)
        if (C::NPIN > 0 && p.grp[g].has_pin && npg < 3) pin_g[npg++] = g;
    const int n_pin = npg * C::R;  // partial tiles per pass

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NR; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < C::NBS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < NPB; ++i) { mbar_init(&p_full[i], 1); mbar_init(&p_empty[i], C::NEPI); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], C::NEPI); }
        fence_mbar_init();
        if (!C::RAWIN) tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
    }
    if (warp == 2) { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // warp-uniform for the compiler
    init_scale_factors<C>(tmem_base, warp);
    // programmatic dependent launch: everything above (barriers, tensor memory, scale factors, descriptor prefetch) may run
    // under the previous launch's tail; its results are needed from here on.  The next launch may be set up at once.
    griddep_launch_dependents();
    griddep_wait();

    // All role loops below are executed by the whole warp with warp-uniform control flow; only the
    // TMA / MMA / commit instructions themselves are issued by one lane, chosen with elect.sync —
    // the pattern ptxas recognises.  (Issuing from inside a divergent `if (lane == 0)` region makes
    // it wrap every tcgen05.mma in an ELECT / R2UR "waterfall" loop: ~100 issue cycles per MMA.)
    if (C::RAWIN && warp >= 4 + C::NEPI) {
        raw_row_producer_role<C>(p, smem, a_full, a_empty, warp - 4 - C::NEPI, lane, prof_acc);
    } else if (warp == 0) {
        // ===== input-row producer (TMA; idle when the raw-image producer warps fill the ring) =====
        uint32_t slot = 0, par = 1;   // waiting on parity 1 of a fresh barrier returns immediately
        ItemIter it = item_begin(p);
        WorkItem w;
        while (!C::RAWIN && item_next<C>(p, it, w)) {
            // n_cb == 1: one sliding window over the item's rows.  n_cb > 1: every (pass, channel block)
            // loads its own window of R + 2*PH rows from plane img * n_cb + cb.
            const int nseg = p.n_cb == 1 ? 1 : w.npass * p.n_cb;
            const int nrows = p.n_cb == 1 ? w.npass * C::R + 2 * C::PH : C::R + 2 * C::PH;
            for (int sg = 0; sg < nseg; ++sg) {
                const int cb = p.n_cb == 1 ? 0 : sg % p.n_cb;
                const int y0 = w.y_b - C::PH + (p.n_cb == 1 ? 0 : (sg / p.n_cb) * C::R);
                for (int i = 0; i < nrows; ++i) {
                    RSR_WAIT(&a_empty[slot], par, 1, 0);
                    if (elect_one()) {
                        if (ABL_NO_A_LOAD) { mbar_arrive(&a_full[slot]); }
                        else {
                            mbar_arrive_expect_tx(&a_full[slot], C::ROW_BYTES);
                            tma_load_4d(sA + slot * C::SLOT, &p.tmA, &a_full[slot], 0, w.x0 - C::PW, y0 + i, w.img * p.n_cb + cb);
                        }
                    }
                    __syncwarp();
                    if (++slot == C::NR) { slot = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ===== weight producer: one stage = TPS consecutive taps =====
        uint32_t st = 0, par = 1;
        ItemIter it = item_begin(p);
        WorkItem w;
        while (item_next<C>(p, it, w)) {
            for (int j = 0; j < w.npass; ++j)
                for (int cb = 0; cb < p.n_cb; ++cb)
                    for (int sp = 0; sp < C::NSTG; ++sp) {
                        RSR_WAIT(&b_empty[st], par, 2, 0);
                        if (elect_one()) {
                            if (ABL_NO_B_LOAD) { mbar_arrive(&b_full[st]); }
                            else {
                                mbar_arrive_expect_tx(&b_full[st], C::BST);   // rows past the last tap: next block / zero fill
                                tma_load_2d(sB + st * C::BST, &p.tmB, &b_full[st], 0, (cb * C::TAPS + sp * C::TPS) * C::NCOL);
                            }
                        }
                        __syncwarp();
                        if (++st == C::NBS) { st = 0; par ^= 1; }
                    }
        }
    } else if (warp == 2) {
        // ===== partial-sum tile producer =====
        // Tile order: per pass, row then group (the order in which the epilogue consumes them).
        if (C::NPIN > 0 && n_pin > 0) {
            uint32_t st = 0, par = 1;
            ItemIter it = item_begin(p);
            WorkItem w;
            while (item_next<C>(p, it, w)) {
                for (int j = 0; j < w.npass; ++j)
                    for (int r = 0; r < C::R; ++r)
                        for (int k = 0; k < npg; ++k) {
                            RSR_WAIT(&p_empty[st], par, 7, 0);
                            if (elect_one()) {
                                if (ABL_NO_PIN) { mbar_arrive(&p_full[st]); }
                                else {
                                    mbar_arrive_expect_tx(&p_full[st], C::PIN_BYTES);
                                    tma_load_4d(sP + st * C::PIN_BYTES, &p.tmPin[pin_g[k]], &p_full[st], 0, w.x0,
                                                w.y_b + j * C::R + r, w.img);
                                }
                            }
                            __syncwarp();
                            if (++st == NPB) { st = 0; par ^= 1; }
                        }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // Descriptors: the upper 32 bits are constant, the lower 32 are (address >> 4) | LBO << 16, so
        // a descriptor for a neighbouring chunk / plane / pixel shift is one 32-bit add.  All ring
        // positions advance incrementally (no division); the tap loop is fully unrolled so that tap
        // offsets are immediates.  The issue thread must stay ahead of an 8-deep MMA queue
        // (tests/cuda/umma_queue.cu), so every instruction here counts.
        using MM = SplitMma<C::PREC, 1, C::RB>;
        constexpr uint32_t LBO = 1u << 16;
        constexpr int WIN = C::R + 2 * C::PH;        // input rows one pass reads
        constexpr int NEARLY = C::FOLD ? 0 : (C::R < C::KH ? C::R : C::KH);   // rows retired during the pass (folded: every
                                                                               // stage reads the whole window)
        const uint32_t sA_e = (smem_u32(sA) >> 4) | LBO, sB_e = (smem_u32(sB) >> 4) | LBO;
        const uint32_t sf = tmem_base + C::SF_COL;
        uint32_t slot0 = 0;              // ring slot of the first window row of the current pass
        uint32_t wslot = 0, wpar = 0;    // next row slot to wait for, and its parity
        uint32_t bst = 0, bpar = 0;      // weight stage ring
        uint32_t buf = 0, tpar = 1;      // accumulator buffer, parity of its t_empty barrier
        ItemIter it = item_begin(p);
        WorkItem w;
        while (item_next<C>(p, it, w)) {
            for (int j = 0; j < w.npass; ++j) {
                RSR_WAIT(&t_empty[buf], tpar, 3, 0);
                for (int cb = 0; cb < p.n_cb; ++cb) {
                // rows this segment adds to the window: all of it for the first pass of an item (and for every
                // channel block of the generic conv), else R
                const bool fresh = j == 0 || p.n_cb > 1;
                const int nnew = fresh ? WIN : C::R;
                for (int i = 0; i < nnew; ++i) {
                    RSR_WAIT(&a_full[wslot], wpar, 4, 1);
                    if (++wslot == C::NR) { wslot = 0; wpar ^= 1; }
                }
                tc_fence_after();
                // encoded smem addresses of the window rows
                uint32_t arow[WIN];
                {
                    uint32_t sl = slot0;
                    #pragma unroll
                    for (int i = 0; i < WIN; ++i) {
                        arow[i] = sA_e + sl * (C::SLOT >> 4);
                        if (++sl == C::NR) sl = 0;
                    }
                }
                const uint32_t d_pass = tmem_base + buf * (C::R * C::NCOL);
                if constexpr (C::FOLD) {
                    // stage sp = tap column tb: weight rows [j][n], j = KH-1-ta (bottom tap first), so the taps that input row i
                    // feeds into output rows r_lo..r_hi (ta = i-r_lo down to i-r_hi) are one contiguous block of rows
                    #pragma unroll
                    for (int sp = 0; sp < C::NSTG; ++sp) {
                        RSR_WAIT(&b_full[bst], bpar, 5, 2);
                        tc_fence_after();
                        const uint32_t b_st = sB_e + bst * (C::BST >> 4);
                        const long long ti0 = RSR_PROF_ON ? clock64() : 0;
                        if (elect_one()) {
                            #pragma unroll
                            for (int i = 0; i < WIN; ++i) {
                                const int r_lo = i - (C::KH - 1) > 0 ? i - (C::KH - 1) : 0;
                                const int r_hi = i < C::R - 1 ? i : C::R - 1;
                                const int n = r_hi - r_lo + 1;                       // output rows this input row feeds
                                const uint32_t a_e = arow[i] + sp * (C::RB >> 4);
                                const uint32_t b_e = b_st + (C::KH - 1 - (i - r_lo)) * (C::TAP_BYTES >> 4);
                                const uint32_t d = d_pass + r_lo * C::NCOL;
                                #pragma unroll
                                for (int c = 0; c < C::KC; ++c) {
                                    if (ABL_NO_MMA) continue;
                                    if (sp == 0 && c == 0 && i < C::R && cb == 0) {
                                        // the first product that reaches output row i (= r_hi) overwrites its accumulator; the
                                        // rows above it in the same block already hold sums: two MMAs instead of one
                                        // (further channel blocks of the generic conv accumulate onto the first one's sums)
                                        if (n > 1) MM::head(d, a_e, b_e, c, 128, (n - 1) * C::NCOL, 1);
                                        MM::head(d + (n - 1) * C::NCOL, a_e, b_e + (n - 1) * (C::TAP_BYTES >> 4), c, 128, C::NCOL, 0);
                                    } else {
                                        MM::head(d, a_e, b_e, c, 128, n * C::NCOL, 1);
                                    }
                                    MM::tail(d, a_e, b_e, c, 128, n * C::NCOL, sf);
                                }
                            }
                            umma_commit(&b_empty[bst]);
                        }
                        __syncwarp();
                        if (RSR_PROF_ON) prof_acc[4] += clock64() - ti0;
                        if (++bst == C::NBS) { bst = 0; bpar ^= 1; }
                    }
                } else {
                #pragma unroll
                for (int sp = 0; sp < C::NSTG; ++sp) {
                    RSR_WAIT(&b_full[bst], bpar, 5, 2);
                    tc_fence_after();
                    const uint32_t b_st = sB_e + bst * (C::BST >> 4);
                    const long long ti0 = RSR_PROF_ON ? clock64() : 0;
                    if (elect_one()) {
                        #pragma unroll
                        for (int tt = 0; tt < C::TPS; ++tt) {
                            const int t = sp * C::TPS + tt;
                            if (t < C::TAPS) {
                                const int ta = t / C::KW, tb = t % C::KW;
                                const uint32_t b_e = b_st + tt * (C::TAP_BYTES >> 4);
                                #pragma unroll
                                for (int r = 0; r < C::R; ++r) {
                                    const uint32_t a_e = arow[r + ta] + tb * (C::RB >> 4);
                                    const uint32_t d = d_pass + r * C::NCOL;
                                    #pragma unroll
                                    for (int c = 0; c < C::KC; ++c) {
                                        if (ABL_NO_MMA) continue;
                                        MM::chunk(d, a_e, b_e, c, 128, C::NCOL, sf, (t | c) != 0 ? 1u : (uint32_t)(cb != 0));
                                    }
                                }
                                // the top window rows are dead once their tap row is done: free them early
                                if (tb == C::KW - 1 && ta < NEARLY) {
                                    uint32_t sl = slot0 + ta;
                                    if (sl >= C::NR) sl -= C::NR;
                                    umma_commit(&a_empty[sl]);
                                }
                            }
                        }
                        umma_commit(&b_empty[bst]);
                    }
                    __syncwarp();
                    if (RSR_PROF_ON) prof_acc[4] += clock64() - ti0;
                    if (++bst == C::NBS) { bst = 0; bpar ^= 1; }
                }
                }  // !FOLD
                const bool last = j == w.npass - 1 || p.n_cb > 1;
                const int n_retire = last ? WIN : C::R;
                if (elect_one()) {
                    if (cb == p.n_cb - 1) umma_commit(&t_full[buf]);
                    uint32_t sl = slot0 + NEARLY;
                    if (sl >= C::NR) sl -= C::NR;
                    for (int i = NEARLY; i < n_retire; ++i) {
                        umma_commit(&a_empty[sl]);
                        if (++sl == C::NR) sl = 0;
                    }
                }
                __syncwarp();
                slot0 += n_retire;
                if (slot0 >= C::NR) slot0 -= C::NR;
                if (slot0 >= C::NR) slot0 -= C::NR;
                }  // channel blocks
                buf ^= 1;
                if (buf == 0) tpar ^= 1;
            }
        }
    } else if (warp >= 4 && warp < 4 + C::NEPI) {
        // ===== epilogue =====
        ItemIter it = item_begin(p);
        conv_epilogue_role<C>(p, smem, p_full, p_empty, t_full, tmem_base, warp, lane, prof_acc,
                              [&](WorkItem& w) { return item_next<C>(p, it, w); },
                              [&](uint32_t buf) { mbar_arrive(&t_empty[buf]); });
    }
    if (RSR_PROF_ON && lane == 0 && (warp < 5)) {
        // row per CTA: [warp0: a_empty] [warp1: t_empty a_full b_full p_full] [warp2: p_empty] [warp3: b_empty]
        //              [warp4: t_full store_read] [9: MMA issue blocks] [10: epilogue busy] [11: warp4 p_full] [12: warp4 tcgen05.ld + wait]
        //              [13: warp4 output path: math, staging, store issue (store_read wait included)] [15: total cycles]
        long long* dst = p.prof + (size_t)blockIdx.x * 16;
        const int base = warp == 0 ? 0 : warp == 1 ? 1 : warp == 2 ? 5 : warp == 3 ? 6 : 7;
        const int n = warp == 1 ? 4 : warp == 4 ? 2 : 1;
        for (int i = 0; i < n; ++i) dst[base + i] = prof_acc[i];
        if (warp == 1) { dst[9] = prof_acc[4]; dst[15] = clock64() - prof_t0; }
        if (warp == 4) { dst[10] = prof_acc[2]; dst[11] = prof_acc[3]; dst[12] = prof_acc[5]; dst[13] = prof_acc[6]; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace rsr
