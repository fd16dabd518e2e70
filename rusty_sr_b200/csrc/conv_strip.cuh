// conv_strip.cuh — the implicit-GEMM convolution kernel of the rusty_sr forward pass on sm_100a.
//
// Replaces alumina's Convolution (+ Bias + BeLU + Expand + LinearInterp) ops that
// `graph.forward` executes for sr_net (reference src/network.rs:27-72, src/main.rs:171).
//
// Formulation ("by input"): one launch takes ONE input node (32 channels, stored as a split-bf16
// pair hi|mid = 128 B per pixel) and computes EVERY convolution that reads it, as one GEMM with
// the output channels of those convolutions concatenated along N:
//     D[pixel, n] = sum_{tap, ci} A[pixel + tap, ci] * W[tap][n][ci]
//   M = 128 consecutive pixels of one image row, N = 32 * NG, K = taps * 32.
// fp32-grade accuracy on the bf16 tensor pipe comes from the 3-term split product
//     a*w ~= a_hi*w_hi + a_mid*w_hi + a_hi*w_mid          ("BF16x3")
// i.e. three tcgen05.mma (kind::f16, fp32 accumulate in TMEM) per 16-channel K chunk.
//
// Data movement: a CTA walks down a 128-pixel-wide column strip.  Input rows are TMA-loaded ONCE
// each into a ring of row slots in shared memory (box = 128 + KW - 1 pixels, SWIZZLE_128B,
// out-of-bounds = zero fill = `Padding::Same`); the KH*KW taps are NOT re-loaded: the tap
// (a, b) operand is the same shared-memory row slot (row + a) read through a UMMA descriptor whose
// start address is shifted by b pixels.  Weights stream per tap through a small ring.  R output
// rows are accumulated per pass in TMEM (double buffered) so that the epilogue of pass j overlaps
// the MMAs of pass j + 1.
//
// Warp roles (256 threads): warp 0 = input-row TMA producer, warp 3 = weight TMA producer,
// warp 1 = MMA issuer (one lane), warp 2 = TMEM allocator, warps 4..7 = epilogue (TMEM -> regs ->
// bias / BeLU / split / partial sums / depth-to-space + bilinear skip -> global).
#pragma once
#include <cuda_bf16.h>
#include "ptx.cuh"

namespace rsr {

// What happens to one 32-column group of the accumulator.
enum EpiKind : int {
    EPI_NONE = 0,
    EPI_ACT = 1,      // v = acc + bias (+ partial_in); y = BeLU(v; slope); store split-bf16 activation
    EPI_PARTIAL = 2,  // v = acc (+ partial_in); store fp32 partial pre-activation sum
    EPI_SHUFFLE = 3,  // v = acc + bias (+ partial_in); depth-to-space + bilinear skip; store image
};

struct EpiGroup {
    int kind;
    int pad_;
    const float* partial_in;   // [N][H][W][32] fp32 or null
    float* partial_out;        // EPI_PARTIAL
    __nv_bfloat16* act_out;    // EPI_ACT: [N][H][W][64] (hi 32 | mid 32)
    const float* bias;         // 32 floats (zero padded)
    const float* slope;        // 32 floats: BeLU p
};

struct ConvParams {
    CUtensorMap tmA;  // input node, rank 4: (RB/2 elems, W, H, NIMG)
    CUtensorMap tmB;  // packed weights, rank 2: (RB/2 elems, taps * NCOL)
    int W, H, NIMG;
    int row_begin, row_end;  // output rows computed (band mode); [0, H) normally
    int n_strips, n_segs, seg_rows, total_items;
    int factor;
    EpiGroup grp[3];
    const float4* xf4;   // EPI_SHUFFLE: input image as float4 [N][H][W] (bilinear skip)
    float* out_f32;      // EPI_SHUFFLE: [N][fH][fW][3] (or null)
    uint8_t* out_u8;     // EPI_SHUFFLE: [N][fH][fW][4] RGBA (or null)
    int* fault;
};

#ifndef RSR_UMMA_BASE_OFFSET_POLICY
#define RSR_UMMA_BASE_OFFSET_POLICY 0   // set from the hardware probe (tests/cuda/umma_probe.cu)
#endif

template <int KH_, int KW_, int RB_, int NG_, int R_>
struct ConvCfg {
    static constexpr int KH = KH_, KW = KW_, RB = RB_, NG = NG_, R = R_;
    static constexpr int NCOL = 32 * NG;
    static constexpr int PH = KH / 2, PW = KW / 2;
    static constexpr int TAPS = KH * KW;
    static constexpr int BOXW = 128 + KW - 1;                          // pixels per loaded row
    static constexpr int ROW_BYTES = BOXW * RB;                        // TMA bytes per row
    static constexpr int SLOT = (ROW_BYTES + 1023) / 1024 * 1024;      // ring slot pitch
    static constexpr int NR = 2 * R + 2 * PH;                          // ring slots (one pass lookahead)
    static constexpr int BST = NCOL * RB;                              // bytes per weight stage (one tap)
    static constexpr int NBS = 4;                                      // weight stages
    static constexpr int KC = RB / 64;                                 // 16-channel chunks per plane
    static constexpr int ACC_COLS = 2 * R * NCOL;                      // double-buffered accumulators
    static constexpr int TMEM_COLS = ACC_COLS <= 32 ? 32 : ACC_COLS <= 64 ? 64 : ACC_COLS <= 128 ? 128
                                   : ACC_COLS <= 256 ? 256 : 512;
    static constexpr int NBAR = 2 * NR + 2 * NBS + 4;
    static constexpr int SMEM_BYTES = 1024 + NR * SLOT + NBS * BST + NBAR * 8 + 16;
    static constexpr uint32_t SWZ = RB == 128 ? UMMA_SW_128B : UMMA_SW_64B;
    static_assert(ACC_COLS <= 512, "accumulators exceed TMEM");
    static_assert(RB == 128 || RB == 64, "pixel row must be 128 or 64 bytes");
    static_assert(BST % 1024 == 0, "weight stage must keep 1024-byte alignment");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

struct WorkItem {
    int img, x0, y_b, y_e, npass;
};

template <class C>
__device__ __forceinline__ WorkItem decode_item(const ConvParams& p, int item) {
    WorkItem w;
    const int seg = item % p.n_segs;
    const int t = item / p.n_segs;
    const int strip = t % p.n_strips;
    w.img = t / p.n_strips;
    w.x0 = strip * 128;
    w.y_b = p.row_begin + seg * p.seg_rows;
    w.y_e = min(w.y_b + p.seg_rows, p.row_end);
    w.npass = (w.y_e - w.y_b + C::R - 1) / C::R;
    return w;
}

__device__ __forceinline__ float belu(float x, float p) {
    // alumina BeLU (network.rs:35,54-56): y = p*x + sqrt(x^2 + 1) - 1   
    return fmaf(p, x, sqrtf(fmaf(x, x, 1.0f)) - 1.0f);
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
    return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

// ---- epilogue pieces (one thread = one pixel, v[] = its 32 accumulator columns) -------------
__device__ __forceinline__ void epi_add_vec(float (&v)[32], const float* __restrict__ src) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = __ldg(s4 + i);
        v[4 * i + 0] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
}
// Partial sums may be updated in place (and the buffers are re-used between nodes), so they are
// read with ordinary coherent loads, never through the read-only path.
__device__ __forceinline__ void epi_add_partial(float (&v)[32], const float* src) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = s4[i];
        v[4 * i + 0] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
}

__device__ __forceinline__ void epi_act(const EpiGroup& g, float (&v)[32], size_t px) {
    epi_add_vec(v, g.bias);
    if (g.partial_in) epi_add_partial(v, g.partial_in + px * 32);
    uint32_t hi[16], mid[16];
    #pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float y0 = belu(v[2 * i], __ldg(g.slope + 2 * i));
        const float y1 = belu(v[2 * i + 1], __ldg(g.slope + 2 * i + 1));
        const __nv_bfloat16 h0 = __float2bfloat16_rn(y0), h1 = __float2bfloat16_rn(y1);
        const __nv_bfloat16 m0 = __float2bfloat16_rn(y0 - __bfloat162float(h0));
        const __nv_bfloat16 m1 = __float2bfloat16_rn(y1 - __bfloat162float(h1));
        hi[i] = pack_bf16x2(h0, h1);
        mid[i] = pack_bf16x2(m0, m1);
    }
    uint4* dst = reinterpret_cast<uint4*>(g.act_out + px * 64);
    #pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
    #pragma unroll
    for (int i = 0; i < 4; ++i) dst[4 + i] = make_uint4(mid[4 * i], mid[4 * i + 1], mid[4 * i + 2], mid[4 * i + 3]);
}

__device__ __forceinline__ void epi_partial(const EpiGroup& g, float (&v)[32], size_t px) {
    if (g.partial_in) epi_add_partial(v, g.partial_in + px * 32);
    float4* dst = reinterpret_cast<float4*>(g.partial_out + px * 32);
    #pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// Expand (depth-to-space, network.rs:39) + LinearInterp skip (network.rs:27) + image store.
__device__ __forceinline__ void epi_shuffle(const ConvParams& p, const EpiGroup& g, float (&v)[32], int img,
                                            int y, int x) {
    const int F = p.factor, W = p.W, H = p.H;
    const size_t px = ((size_t)img * H + y) * W + x;
    epi_add_vec(v, g.bias);
    if (g.partial_in) epi_add_partial(v, g.partial_in + px * 32);
    // 3x3 neighbourhood of the input, edge clamped (half-pixel-centre bilinear)
    const int ym = max(y - 1, 0), yp = min(y + 1, H - 1), xm = max(x - 1, 0), xp = min(x + 1, W - 1);
    const float4* xi = p.xf4 + (size_t)img * H * W;
    float nb[3][3][3];
    {
        const int ys[3] = {ym, y, yp}, xs[3] = {xm, x, xp};
        #pragma unroll
        for (int a = 0; a < 3; ++a)
            #pragma unroll
            for (int b = 0; b < 3; ++b) {
                const float4 t = __ldg(xi + (size_t)ys[a] * W + xs[b]);
                nb[a][b][0] = t.x; nb[a][b][1] = t.y; nb[a][b][2] = t.z;
            }
    }
    const float invF = 1.0f / (float)F;
    for (int dy = 0; dy < F; ++dy) {
        const float sy = ((float)dy + 0.5f) * invF - 0.5f;     // in (-0.5, 0.5)
        const int ya = sy < 0.f ? 0 : 1;                       // rows (ya, ya+1) of the 3x3 block
        const float ty = sy < 0.f ? sy + 1.0f : sy;
        const size_t orow = ((size_t)img * H * F + (size_t)y * F + dy) * ((size_t)W * F) + (size_t)x * F;
        for (int dx = 0; dx < F; ++dx) {
            const float sx = ((float)dx + 0.5f) * invF - 0.5f;
            const int xa = sx < 0.f ? 0 : 1;
            const float tx = sx < 0.f ? sx + 1.0f : sx;
            float o[3];
            #pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v00, v01, v10, v11;
                // select without dynamic register indexing
                v00 = ya == 0 ? (xa == 0 ? nb[0][0][c] : nb[0][1][c]) : (xa == 0 ? nb[1][0][c] : nb[1][1][c]);
                v01 = ya == 0 ? (xa == 0 ? nb[0][1][c] : nb[0][2][c]) : (xa == 0 ? nb[1][1][c] : nb[1][2][c]);
                v10 = ya == 0 ? (xa == 0 ? nb[1][0][c] : nb[1][1][c]) : (xa == 0 ? nb[2][0][c] : nb[2][1][c]);
                v11 = ya == 0 ? (xa == 0 ? nb[1][1][c] : nb[1][2][c]) : (xa == 0 ? nb[2][1][c] : nb[2][2][c]);
                const float top = v00 + (v01 - v00) * tx;
                const float bot = v10 + (v11 - v10) * tx;
                // Expand channel order: e[(dy*F + dx)*3 + c]
                float e = 0.f;
                const int ch = (dy * F + dx) * 3 + c;
                #pragma unroll
                for (int k = 0; k < 32; ++k) e = (k == ch) ? v[k] : e;
                o[c] = e + top + (bot - top) * ty;
            }
            if (p.out_f32) {
                float* d = p.out_f32 + (orow + dx) * 3;
                d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
            }
            if (p.out_u8) {
                // alumina data_to_img + to_rgba (main.rs:175): round(clamp(v,0,1)*255), alpha 255
                uchar4 q;
                q.x = (unsigned char)(fminf(fmaxf(o[0], 0.f), 1.f) * 255.0f + 0.5f);
                q.y = (unsigned char)(fminf(fmaxf(o[1], 0.f), 1.f) * 255.0f + 0.5f);
                q.z = (unsigned char)(fminf(fmaxf(o[2], 0.f), 1.f) * 255.0f + 0.5f);
                q.w = 255;
                reinterpret_cast<uchar4*>(p.out_u8)[orow + dx] = q;
            }
        }
    }
}

// ---- the kernel ------------------------------------------------------------------------------
template <class C>
__global__ void __launch_bounds__(256, 1) conv_strip_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = sA + C::NR * C::SLOT;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + C::NBS * C::BST);
    uint64_t* a_empty = a_full + C::NR;
    uint64_t* b_full = a_empty + C::NR;
    uint64_t* b_empty = b_full + C::NBS;
    uint64_t* t_full = b_empty + C::NBS;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NR; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < C::NBS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        fence_mbar_init();
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
    }
    if (warp == 2) { tmem_alloc(tmem_slot, C::TMEM_COLS); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== input-row producer =====
        if (lane == 0) {
            uint32_t s = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const WorkItem w = decode_item<C>(p, item);
                const int nrows = w.npass * C::R + 2 * C::PH;
                for (int i = 0; i < nrows; ++i, ++s) {
                    const uint32_t slot = s % C::NR, par = (s / C::NR) & 1;
                    mbar_wait(&a_empty[slot], par ^ 1, p.fault, 1);
                    mbar_arrive_expect_tx(&a_full[slot], C::ROW_BYTES);
                    tma_load_4d(sA + slot * C::SLOT, &p.tmA, &a_full[slot], 0, w.x0 - C::PW, w.y_b - C::PH + i, w.img);
                }
            }
        }
    } else if (warp == 3) {
        // ===== weight producer =====
        if (lane == 0) {
            uint32_t q = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const WorkItem w = decode_item<C>(p, item);
                for (int j = 0; j < w.npass; ++j)
                    for (int t = 0; t < C::TAPS; ++t, ++q) {
                        const uint32_t st = q % C::NBS, par = (q / C::NBS) & 1;
                        mbar_wait(&b_empty[st], par ^ 1, p.fault, 2);
                        mbar_arrive_expect_tx(&b_full[st], C::BST);
                        tma_load_2d(sB + st * C::BST, &p.tmB, &b_full[st], 0, t * C::NCOL);
                    }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t HI = (uint32_t)make_sdesc_hi(8 * C::RB, C::SWZ);
            constexpr uint32_t IDESC = make_idesc_bf16(128, C::NCOL);
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            uint32_t s_base = 0, waited = 0, q = 0, pq = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
                const WorkItem w = decode_item<C>(p, item);
                for (int j = 0; j < w.npass; ++j, ++pq) {
                    const uint32_t buf = pq & 1, tpar = (pq >> 1) & 1;
                    mbar_wait(&t_empty[buf], tpar ^ 1, p.fault, 3);
                    const uint32_t need = s_base + j * C::R + C::R + 2 * C::PH;
                    while (waited < need) {
                        mbar_wait(&a_full[waited % C::NR], (waited / C::NR) & 1, p.fault, 4);
                        ++waited;
                    }
                    tc_fence_after();
                    for (int t = 0; t < C::TAPS; ++t, ++q) {
                        const int ta = t / C::KW, tb = t % C::KW;
                        const uint32_t st = q % C::NBS;
                        mbar_wait(&b_full[st], (q / C::NBS) & 1, p.fault, 5);
                        tc_fence_after();
                        const uint32_t b_addr = sB_u + st * C::BST;
                        #pragma unroll
                        for (int r = 0; r < C::R; ++r) {
                            const uint32_t slot = (s_base + j * C::R + r + ta) % C::NR;
                            const uint32_t a_addr = sA_u + slot * C::SLOT + tb * C::RB;
                            const uint32_t d = tmem_base + buf * (C::R * C::NCOL) + r * C::NCOL;
                            #pragma unroll
                            for (int c = 0; c < C::KC; ++c) {
                                const uint32_t a_hi = a_addr + c * 32, a_mid = a_addr + C::RB / 2 + c * 32;
                                const uint32_t b_hi = b_addr + c * 32, b_mid = b_addr + C::RB / 2 + c * 32;
                                const uint32_t bo_a = RSR_UMMA_BASE_OFFSET_POLICY ? ((a_addr >> 7) & 7) : 0;
                                const uint64_t dA_hi = make_sdesc(a_hi, 16, HI, bo_a);
                                const uint64_t dA_mid = make_sdesc(a_mid, 16, HI, bo_a);
                                const uint64_t dB_hi = make_sdesc(b_hi, 16, HI, 0);
                                const uint64_t dB_mid = make_sdesc(b_mid, 16, HI, 0);
                                umma_f16(d, dA_hi, dB_hi, IDESC, (t | c) != 0);
                                umma_f16(d, dA_mid, dB_hi, IDESC, 1);
                                umma_f16(d, dA_hi, dB_mid, IDESC, 1);
                            }
                        }
                        umma_commit(&b_empty[st]);
                    }
                    umma_commit(&t_full[buf]);
                    const int n_retire = (j == w.npass - 1) ? C::R + 2 * C::PH : C::R;
                    for (int i = 0; i < n_retire; ++i) umma_commit(&a_empty[(s_base + j * C::R + i) % C::NR]);
                }
                s_base += w.npass * C::R + 2 * C::PH;
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue =====
        const int ew = warp - 4;  // == warp % 4: TMEM lane quarter this warp may read
        uint32_t pq = 0;
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            const WorkItem w = decode_item<C>(p, item);
            const int x = w.x0 + ew * 32 + lane;
            for (int j = 0; j < w.npass; ++j, ++pq) {
                const uint32_t buf = pq & 1, tpar = (pq >> 1) & 1;
                mbar_wait(&t_full[buf], tpar, p.fault, 6);
                tc_fence_after();
                #pragma unroll 1
                for (int r = 0; r < C::R; ++r) {
                    const int y = w.y_b + j * C::R + r;
                    if (y >= w.y_e) break;  // warp-uniform
                    const bool valid = x < p.W;
                    const size_t px = ((size_t)w.img * p.H + y) * p.W + (valid ? x : 0);
                    #pragma unroll 1
                    for (int g = 0; g < C::NG; ++g) {
                        const EpiGroup& eg = p.grp[g];
                        if (eg.kind == EPI_NONE) continue;
                        uint32_t raw[32];
                        tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + buf * (C::R * C::NCOL) + r * C::NCOL + g * 32, raw);
                        tmem_ld_wait();
                        if (valid) {
                            float v[32];
                            #pragma unroll
                            for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(raw[k]);
                            if (eg.kind == EPI_ACT) epi_act(eg, v, px);
                            else if (eg.kind == EPI_PARTIAL) epi_partial(eg, v, px);
                            else epi_shuffle(p, eg, v, w.img, y, x);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[buf]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, C::TMEM_COLS);
}

}  // namespace rsr
