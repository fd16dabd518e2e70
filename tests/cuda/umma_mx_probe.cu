// umma_mx_probe.cu — hardware probe + micro-benchmark (run under gpurun) for the "F16 + MXF8" operand split:
//     a*w ~= a_h*w_h  (kind::f16, fp16 operands, K = 16)
//          + 2^-11 * ( e4m3(2^11 a_l) * e4m3(w)  +  e4m3(a) * e4m3(2^11 w_l) )   (ONE kind::mxf8f6f4 MMA, K = 32)
// where the 2^-11 is applied by the tensor core itself: a block-scaled MMA multiplies every row of A by a UE8M0 scale
// factor read from tensor memory, and this probe fills the scale-factor columns with ONE constant byte (2^-11 for A,
// 2^0 for B), so the result does not depend on the scale-factor layout.  Checks, against an exact CPU expectation
// (small integers / powers of two: every product and sum is exact in fp32):
//   1. kind::f16 with fp16 operands from a SWIZZLE_128B K-major tile (the conv kernels' pixel-row layout);
//   2. kind::mxf8f6f4.block_scale (e4m3 x e4m3, K = 32) from the same 128-byte rows, +32 B chunk advance, scale 2^-11;
//   3. both accumulating into the same fp32 TMEM columns;
//   4. the same with cta_group::2 (M = 256, half of the B rows per CTA), as conv_pair.cuh / conv_fused.cuh issue it.
// Then times: cycles per MMA of (f16, mxf8) pairs vs BF16x3 triples at N = 32 / 64 / 96.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include "../../rusty_sr_b200/csrc/conv_pair.cuh"
using namespace rsr;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// kind::f16 instruction descriptor with fp16 operands (a_format = b_format = 0)
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
    return (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// block-scaled descriptor (cute::UMMA::InstrDescriptorBlockScaled): a/b format E4M3 = 0, scale format UE8M0 (bit 23)
__host__ __device__ constexpr uint32_t idesc_mxf8(int M, int N, int sfa_id = 0, int sfb_id = 0) {
    return (uint32_t(sfb_id) << 4) | (uint32_t(N >> 3) << 17) | (1u << 23) | (uint32_t(M >> 4) << 24) | (uint32_t(sfa_id) << 29);
}
template <int CG>
__device__ __forceinline__ void umma_mxf8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t sfa, uint32_t sfb, uint32_t acc) {
    if (CG == 1)
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %6, 0;\n"
                     "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale.scale_vec::1X [%0], %1, %2, %3, [%4], [%5], p;\n}\n"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(sfa), "r"(sfb), "r"(acc) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %6, 0;\n"
                     "tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale.scale_vec::1X [%0], %1, %2, %3, [%4], [%5], p;\n}\n"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(sfa), "r"(sfb), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, uint32_t v) {   // this warp's 32 lanes x 16 columns := v
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};"
                 ::"r"(taddr), "r"(v) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

constexpr int NB = 96;          // B rows (N)
constexpr int SF_A = 448, SF_B = 464;   // scale-factor columns (16 each), beyond the accumulators

// A: [128 rows][128 B], B: [NB rows][128 B] given in LOGICAL row-major form; the kernel stores them with the 128-byte swizzle.
// test: 0 = f16 only (2 chunks), 1 = mxf8 only (2 chunks), 2 = both, accumulated.  PAIR: two CTAs, M = 256, each holds NB/2 B rows.
template <bool PAIR>
__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ B, int test,
                                                         uint32_t sfa_byte, int sf_id, float* __restrict__ D, int* fault) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // 16 KB
    uint8_t* sB = smem + 16 * 1024;     // up to 12 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0;
    const int nb_local = PAIR ? NB / 2 : NB;
    // operands into shared memory, 128 B swizzle: 16-byte chunk j of row r at chunk j ^ (r & 7)
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
        const int r = i >> 3, j = i & 7;
        reinterpret_cast<uint4*>(sA)[r * 8 + (j ^ (r & 7))] = reinterpret_cast<const uint4*>(A)[((size_t)rank * 128 + r) * 8 + j];
    }
    for (int i = threadIdx.x; i < nb_local * 8; i += blockDim.x) {
        const int r = i >> 3, j = i & 7;
        reinterpret_cast<uint4*>(sB)[r * 8 + (j ^ (r & 7))] = reinterpret_cast<const uint4*>(B)[((size_t)rank * nb_local + r) * 8 + j];
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_s)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else { tmem_alloc(&tmem_s, 512); tmem_relinquish(); }
    }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_s;
    // scale factors: every byte of 16 columns x 128 lanes the same value
    tmem_st_x16(tmem + ((uint32_t)(warp * 32) << 16) + SF_A, sfa_byte * 0x01010101u);
    tmem_st_x16(tmem + ((uint32_t)(warp * 32) << 16) + SF_B, 127u * 0x01010101u);
    tc_fence_before(); __syncthreads();
    if (PAIR) cluster_sync_all();
    tc_fence_after();
    constexpr uint32_t HI = (uint32_t)make_sdesc_hi(1024, UMMA_SW_128B);
    const uint32_t a_e = (smem_u32(sA) >> 4) | (1u << 16), b_e = (smem_u32(sB) >> 4) | (1u << 16);
    auto desc = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
    constexpr int M = PAIR ? 256 : 128;
    if (warp == 0 && rank == 0) {
        if (elect_one()) {
            uint32_t acc = 0;
            if (test == 0 || test == 2)
                for (int c = 0; c < 2; ++c) {
                    if (PAIR) umma2_f16(tmem, desc(HI, a_e + 2 * c), desc(HI, b_e + 2 * c), idesc_f16(M, NB), acc);
                    else umma_f16(tmem, desc(HI, a_e + 2 * c), desc(HI, b_e + 2 * c), idesc_f16(M, NB), acc);
                    acc = 1;
                }
            if (test == 1 || test == 2)
                for (int c = 0; c < 2; ++c) {
                    umma_mxf8<PAIR ? 2 : 1>(tmem, desc(HI, a_e + 4 + 2 * c), desc(HI, b_e + 4 + 2 * c), idesc_mxf8(M, NB, sf_id, sf_id),
                                            tmem + SF_A, tmem + SF_B, acc);
                    acc = 1;
                }
            if (PAIR) umma2_commit(&bar); else umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0, fault, 1);
    tc_fence_after();
    uint32_t v[32];
    for (int g = 0; g < NB / 32; ++g) {
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + g * 32, v);
        tmem_ld_wait();
        for (int k = 0; k < 32; ++k) D[((size_t)rank * 128 + warp * 32 + lane) * NB + g * 32 + k] = __uint_as_float(v[k]);
    }
    tc_fence_before(); __syncthreads();
    if (PAIR) {
        cluster_sync_all();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    } else if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- timing: per 16-channel chunk either a BF16x3 triple (pattern 0) or an (f16, mxf8) pair (pattern 1); two accumulators
// alternate and the A start walks as in the conv kernels.  Pattern 2: mxf8 only; pattern 3: f16 only.
struct Result { long long cycles; int n, pattern, count; };
__global__ void __launch_bounds__(128, 1) bench_kernel(Result* out, int nres, const int* ns, const int* pats, int count, int* fault) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_s, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tmem_s;
    tmem_st_x16(tmem + ((uint32_t)(warp * 32) << 16) + SF_A, 116u * 0x01010101u);
    tmem_st_x16(tmem + ((uint32_t)(warp * 32) << 16) + SF_B, 127u * 0x01010101u);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    constexpr uint32_t HI = (uint32_t)make_sdesc_hi(1024, UMMA_SW_128B);
    const uint32_t a_e = (smem_u32(smem) >> 4) | (1u << 16);
    const uint32_t b_e = (smem_u32(smem + 96 * 1024) >> 4) | (1u << 16);
    auto desc = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
    if (warp == 0) {
        uint32_t phase = 0;
        for (int r = 0; r < nres; ++r) {
            const int N = ns[r], pat = pats[r];
            const uint32_t i_bf = make_idesc_bf16(128, N), i_f16 = idesc_f16(128, N), i_mx = idesc_mxf8(128, N);
            __syncwarp();
            const long long t0 = clock64();
            if (elect_one()) {
                if (pat == 0) {
                    for (int i = 0; i < count / 6; ++i) {
                        const uint32_t a = a_e + (i & 3) * 8 + (i & 1) * 2;
                        #pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const uint32_t aa = a + rr * (17408 >> 4), d = tmem + rr * 192;
                            umma_f16(d, desc(HI, aa), desc(HI, b_e), i_bf, 1);
                            umma_f16(d, desc(HI, aa + 4), desc(HI, b_e), i_bf, 1);
                            umma_f16(d, desc(HI, aa), desc(HI, b_e + 4), i_bf, 1);
                        }
                    }
                } else if (pat == 1) {
                    for (int i = 0; i < count / 4; ++i) {
                        const uint32_t a = a_e + (i & 3) * 8 + (i & 1) * 2;
                        #pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const uint32_t aa = a + rr * (17408 >> 4), d = tmem + rr * 192;
                            umma_f16(d, desc(HI, aa), desc(HI, b_e), i_f16, 1);
                            umma_mxf8<1>(d, desc(HI, aa + 4), desc(HI, b_e + 4), i_mx, tmem + SF_A, tmem + SF_B, 1);
                        }
                    }
                } else if (pat == 2) {
                    for (int i = 0; i < count; ++i) {
                        const uint32_t a = a_e + (i & 3) * 8 + (i & 1) * 2;
                        umma_mxf8<1>(tmem + (i & 1) * 192, desc(HI, a + 4), desc(HI, b_e + 4), i_mx, tmem + SF_A, tmem + SF_B, 1);
                    }
                } else {
                    for (int i = 0; i < count; ++i) {
                        const uint32_t a = a_e + (i & 3) * 8 + (i & 1) * 2;
                        umma_f16(tmem + (i & 1) * 192, desc(HI, a), desc(HI, b_e), i_f16, 1);
                    }
                }
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, phase, fault, 1);
            phase ^= 1;
            const long long t1 = clock64();
            if (threadIdx.x == 0) out[r] = Result{t1 - t0, N, pat, count};
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- host ----------------------------------------------------------------------------------------------------
static uint8_t e4m3_of(float v) {   // exact for the values used here (0, +-2^e * {1, 1.25, 1.5, 1.75}, e in -6..8)
    if (v == 0.f) return 0;
    uint8_t s = v < 0 ? 0x80 : 0;
    float a = v < 0 ? -v : v;
    int e = 0;
    while (a >= 2.f) { a /= 2.f; ++e; }
    while (a < 1.f) { a *= 2.f; --e; }
    const int m = (int)((a - 1.f) * 8.f + 0.5f);
    return s | (uint8_t)((e + 7) << 3) | (uint8_t)m;
}
static float e4m3_val(int k) {      // a small table of exactly representable values
    static const float t[] = {0.f, 1.f, -1.f, 2.f, 0.5f, -1.5f, 3.f, -0.25f, 4.f, 1.25f, -2.f, 0.75f, 6.f, -3.f, 1.75f, -0.5f};
    return t[k & 15];
}

int main() {
    // logical operands: per row 32 fp16 (64 B) | chunk0: 16 "x" e4m3 + 16 "y" e4m3 | chunk1: same
    const int MR = 256;
    std::vector<uint8_t> A((size_t)MR * 128), B((size_t)NB * 128);
    std::vector<float> Ah((size_t)MR * 32), A8((size_t)MR * 64), Bh((size_t)NB * 32), B8((size_t)NB * 64);
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 16; };
    for (int r = 0; r < MR; ++r) {
        for (int k = 0; k < 32; ++k) { const float v = (float)((int)(rnd() % 17) - 8); Ah[r * 32 + k] = v; reinterpret_cast<__half*>(&A[r * 128])[k] = __float2half(v); }
        for (int k = 0; k < 64; ++k) { const float v = e4m3_val(rnd()); A8[r * 64 + k] = v; A[r * 128 + 64 + k] = e4m3_of(v); }
    }
    for (int r = 0; r < NB; ++r) {
        for (int k = 0; k < 32; ++k) { const float v = (float)((int)(rnd() % 9) - 4) * 0.5f; Bh[r * 32 + k] = v; reinterpret_cast<__half*>(&B[r * 128])[k] = __float2half(v); }
        for (int k = 0; k < 64; ++k) { const float v = e4m3_val(rnd()); B8[r * 64 + k] = v; B[r * 128 + 64 + k] = e4m3_of(v); }
    }
    uint8_t *dA, *dB; float* dD; int* df;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, (size_t)MR * NB * 4)); CK(cudaMalloc(&df, 256));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(df, 0, 256));
    const int smem = 64 * 1024;
    CK(cudaFuncSetAttribute(probe_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(probe_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    std::vector<float> D((size_t)MR * NB);
    int bad_total = 0;
    for (int pair = 0; pair < 2; ++pair)
        for (int test = 0; test < 3; ++test)
            for (int sfcase = 0; sfcase < (test == 0 ? 1 : 3); ++sfcase) {
                const uint32_t sfa_byte = sfcase == 2 ? 127u : 116u;   // 2^-11, or 2^0
                const int sf_id = sfcase == 1 ? 2 : 0;
                CK(cudaMemset(dD, 0xff, (size_t)MR * NB * 4));
                if (pair) {
                    cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
                    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    CK(cudaLaunchKernelEx(&cfg, probe_kernel<true>, (const uint8_t*)dA, (const uint8_t*)dB, test, sfa_byte, sf_id, dD, df));
                } else {
                    probe_kernel<false><<<1, 128, smem>>>(dA, dB, test, sfa_byte, sf_id, dD, df);
                }
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("probe pair=%d test=%d: kernel error %s\n", pair, test, cudaGetErrorString(e)); return 1; }
                const int rows = pair ? 256 : 128;
                CK(cudaMemcpy(D.data(), dD, (size_t)rows * NB * 4, cudaMemcpyDeviceToHost));
                const double sc = sfa_byte == 116u ? 1.0 / 2048.0 : 1.0;
                int bad = 0; double worst = 0;
                for (int r = 0; r < rows; ++r)
                    for (int n = 0; n < NB; ++n) {
                        double ref = 0;
                        if (test == 0 || test == 2) for (int k = 0; k < 32; ++k) ref += (double)Ah[r * 32 + k] * Bh[n * 32 + k];
                        if (test == 1 || test == 2) for (int k = 0; k < 64; ++k) ref += sc * (double)A8[r * 64 + k] * B8[n * 64 + k];
                        const double d = fabs((double)D[(size_t)r * NB + n] - ref);
                        if (d > worst) worst = d;
                        if (d != 0.0) { if (bad < 4) printf("   mismatch r=%d n=%d got %.9g want %.9g\n", r, n, D[(size_t)r * NB + n], ref); ++bad; }
                    }
                printf("probe cta_group::%d test %d (%s) sfa=2^%d sf_id=%d : %d / %d mismatches, max |diff| %.3g\n", pair ? 2 : 1, test,
                       test == 0 ? "f16 only" : test == 1 ? "mxf8 only" : "f16 + mxf8", (int)sfa_byte - 127, sf_id, bad, rows * NB, worst);
                bad_total += bad;
            }
    printf(bad_total == 0 ? "PROBE OK\n" : "PROBE FAILED\n");

    // timing
    std::vector<int> ns, pats;
    for (int p = 0; p < 4; ++p) for (int n : {32, 64, 96}) { ns.push_back(n); pats.push_back(p); }
    const int nres = (int)ns.size(), count = 1200;
    int *dn, *dp; Result* dr;
    CK(cudaMalloc(&dn, nres * 4)); CK(cudaMalloc(&dp, nres * 4)); CK(cudaMalloc(&dr, nres * sizeof(Result)));
    CK(cudaMemcpy(dn, ns.data(), nres * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dp, pats.data(), nres * 4, cudaMemcpyHostToDevice));
    const int bsm = 200 * 1024;
    CK(cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bsm));
    for (int rep = 0; rep < 2; ++rep) {
        bench_kernel<<<1, 128, bsm>>>(dr, nres, dn, dp, count, df);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("bench kernel error %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<Result> res(nres);
    CK(cudaMemcpy(res.data(), dr, nres * sizeof(Result), cudaMemcpyDeviceToHost));
    const char* names[] = {"bf16x3 triple (2 acc)", "f16 + mxf8 pair (2 acc)", "mxf8 only", "f16 only"};
    for (auto& r : res) {
        const double per = (double)r.cycles / r.count;
        const double per_chunk = per * (r.pattern == 0 ? 3 : r.pattern == 1 ? 2 : 1);
        printf("bench %-26s N=%3d : %7.1f cycles/MMA, %7.1f cycles per 16-channel chunk\n", names[r.pattern], r.n, per, per_chunk);
    }
    return bad_total == 0 ? 0 : 1;
}
