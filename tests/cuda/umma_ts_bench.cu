// umma_ts_bench.cu — probe + micro-benchmark (run under gpurun) for feeding the A operand of tcgen05.mma from
// TENSOR MEMORY instead of shared memory.
//
// Why: profiles/README.md (r01i/r01j) — at N <= 96 the conv launches are bound by the 128 B/cycle shared-memory
// operand port: every tcgen05.mma (M=128, K=16) re-reads its 4 KB A tile (32 cycles) plus N x 32 B of weights, and
// the BF16x3 split issues three MMAs per K chunk.  In the conv kernels one A tile (input row y, pixel shift b,
// chunk c, plane hi|mid) is used by KH output rows x (2 | 1) products, so copying it ONCE into tensor memory
// (tcgen05.cp.128x256b) and issuing the MMAs with A = [tmem] would cut the A side of the port traffic by
// 3-6x: an N = 32 MMA could cost 16 cycles instead of 40.  This program measures whether the hardware delivers that.
//
//   part 1: correctness — D = A * B^T computed with A from shared memory (SS) and with A copied to TMEM by
//           tcgen05.cp.128x256b and used from there (TS); both must match bit for bit.
//   part 2: cycles per tcgen05.mma, SS vs TS, N in {32, 64, 96, 128}; and cycles per tcgen05.cp.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../../rusty_sr_b200/csrc/ptx.cuh"
using namespace rsr;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 128 rows x 256 bits (= one K = 16 bf16 chunk of a 128-row operand) from shared memory to tensor memory
__device__ __forceinline__ void tmem_cp_128x256b(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}

struct Timing { long long cycles; int n, mode, count; };

// shared memory: A region = 128 rows x 128 B (SW128, K-major: one swizzle row per M row), B region likewise (N rows)
__global__ void __launch_bounds__(128, 1) ts_kernel(const __nv_bfloat16* a_g, const __nv_bfloat16* b_g, float* d_ss, float* d_ts,
                                                    Timing* tim, int ntim, const int* ns, const int* modes, int count, int* fault) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                 // 16 KB
    uint8_t* sB = smem + 16 * 1024;     // 16 KB (up to 128 rows)
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // fill A and B: row r, 16-byte chunk j of the 128-byte row goes to chunk j ^ (r & 7)  (128 B swizzle)
    for (int i = threadIdx.x; i < 128 * 8; i += blockDim.x) {
        const int r = i >> 3, j = i & 7;
        const uint4 va = reinterpret_cast<const uint4*>(a_g)[i];
        const uint4 vb = reinterpret_cast<const uint4*>(b_g)[i];
        *reinterpret_cast<uint4*>(sA + r * 128 + ((j ^ (r & 7)) << 4)) = va;
        *reinterpret_cast<uint4*>(sB + r * 128 + ((j ^ (r & 7)) << 4)) = vb;
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_s, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_s, 0);
    constexpr uint32_t HI = (uint32_t)make_sdesc_hi(1024, UMMA_SW_128B);
    const uint32_t a_e = (smem_u32(sA) >> 4) | (1u << 16);
    const uint32_t b_e = (smem_u32(sB) >> 4) | (1u << 16);
    auto desc = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
    // TMEM plan: D_ss columns 0..127, D_ts columns 128..255, A chunks at columns 256 + 8 * c (c = 0..3: K = 64)
    const uint32_t d_ss_t = tmem, d_ts_t = tmem + 128, a_t = tmem + 256;
    uint32_t phase = 0;
    if (warp == 0) {
        // ---- part 1: K = 64 (4 chunks), N = 128
        const uint32_t idesc = make_idesc_bf16(128, 128);
        if (elect_one()) {
            for (int c = 0; c < 4; ++c) umma_f16(d_ss_t, desc(HI, a_e + 2 * c), desc(HI, b_e + 2 * c), idesc, c != 0);
            for (int c = 0; c < 4; ++c) tmem_cp_128x256b(a_t + 8 * c, desc(HI, a_e + 2 * c));
            for (int c = 0; c < 4; ++c) umma_f16_ts(d_ts_t, a_t + 8 * c, desc(HI, b_e + 2 * c), idesc, c != 0);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, phase, fault, 1);
    phase ^= 1;
    tc_fence_after();
    {
        // every warp reads its 32 lanes of both accumulators (columns in 4 groups of 32)
        for (int g = 0; g < 4; ++g) {
            uint32_t raw[32];
            tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + g * 32, raw);
            tmem_ld_wait();
            for (int k = 0; k < 32; ++k) d_ss[(warp * 32 + lane) * 128 + g * 32 + k] = __uint_as_float(raw[k]);
            tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + 128 + g * 32, raw);
            tmem_ld_wait();
            for (int k = 0; k < 32; ++k) d_ts[(warp * 32 + lane) * 128 + g * 32 + k] = __uint_as_float(raw[k]);
        }
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    // ---- part 2: timing
    if (warp == 0) {
        for (int r = 0; r < ntim; ++r) {
            const int N = ns[r], mode = modes[r];
            const uint32_t idesc = make_idesc_bf16(128, N);
            __syncwarp();
            const long long t0 = clock64();
            if (elect_one()) {
                if (mode == 0) {          // SS, BF16x3-like operand walk (A tile changes every MMA)
                    for (int i = 0; i < count; ++i) umma_f16(d_ss_t, desc(HI, a_e + 2 * (i & 3)), desc(HI, b_e + 2 * (i & 3)), idesc, 1);
                } else if (mode == 1) {   // TS: A from tensor memory
                    for (int i = 0; i < count; ++i) umma_f16_ts(d_ss_t, a_t + 8 * (i & 3), desc(HI, b_e + 2 * (i & 3)), idesc, 1);
                } else if (mode == 2) {   // tcgen05.cp alone
                    for (int i = 0; i < count; ++i) tmem_cp_128x256b(a_t + 8 * (i & 3), desc(HI, a_e + 2 * (i & 3)));
                } else {                  // the conv pattern: one cp feeds 3 TS MMAs (three output rows share an input row tile)
                    for (int i = 0; i < count / 3; ++i) {
                        tmem_cp_128x256b(a_t + 8 * (i & 3), desc(HI, a_e + 2 * (i & 3)));
                        umma_f16_ts(d_ss_t, a_t + 8 * (i & 3), desc(HI, b_e + 2 * (i & 3)), idesc, 1);
                        umma_f16_ts(d_ss_t + 128, a_t + 8 * (i & 3), desc(HI, b_e + 2 * ((i + 1) & 3)), idesc, 1);
                        umma_f16_ts(d_ss_t, a_t + 8 * (i & 3), desc(HI, b_e + 2 * ((i + 2) & 3)), idesc, 1);
                    }
                }
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, phase, fault, 2);
            phase ^= 1;
            const long long t1 = clock64();
            if (lane == 0) tim[r] = Timing{t1 - t0, N, mode, count};
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    // A: 128 x 64 bf16, B: 128 x 64 bf16 (row-major, K contiguous), small integers so that fp32 sums are exact
    std::vector<__nv_bfloat16> a(128 * 64), b(128 * 64);
    for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 64; ++k) {
            a[r * 64 + k] = __float2bfloat16((float)(((r * 7 + k * 3) % 13) - 6));
            b[r * 64 + k] = __float2bfloat16((float)(((r * 5 + k * 11) % 9) - 4));
        }
    std::vector<int> ns, modes;
    for (int m = 0; m < 4; ++m)
        for (int n : {32, 64, 96, 128}) { if (m == 2 && n != 32) continue; ns.push_back(n); modes.push_back(m); }
    const int ntim = (int)ns.size(), count = 1200;
    __nv_bfloat16 *da, *db; float *dss, *dts; Timing* dt; int *dn, *dm, *df;
    CK(cudaMalloc(&da, a.size() * 2)); CK(cudaMalloc(&db, b.size() * 2));
    CK(cudaMalloc(&dss, 128 * 128 * 4)); CK(cudaMalloc(&dts, 128 * 128 * 4));
    CK(cudaMalloc(&dt, ntim * sizeof(Timing))); CK(cudaMalloc(&dn, ntim * 4)); CK(cudaMalloc(&dm, ntim * 4)); CK(cudaMalloc(&df, 64 * 4));
    CK(cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dn, ns.data(), ntim * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dm, modes.data(), ntim * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(df, 0, 64 * 4)); CK(cudaMemset(dss, 0, 128 * 128 * 4)); CK(cudaMemset(dts, 0xff, 128 * 128 * 4));
    const int smem = 40 * 1024;
    CK(cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 2; ++rep) {
        ts_kernel<<<1, 128, smem>>>(da, db, dss, dts, dt, ntim, dn, dm, count, df);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<float> ss(128 * 128), ts(128 * 128);
    CK(cudaMemcpy(ss.data(), dss, ss.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ts.data(), dts, ts.size() * 4, cudaMemcpyDeviceToHost));
    int bad_ss = 0, bad_ts = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
            float ref = 0.f;
            for (int k = 0; k < 64; ++k) ref += __bfloat162float(a[m * 64 + k]) * __bfloat162float(b[n * 64 + k]);
            bad_ss += ss[m * 128 + n] != ref;
            bad_ts += ts[m * 128 + n] != ref;
        }
    printf("part 1: A from shared memory: %d / 16384 wrong;  A from tensor memory (tcgen05.cp.128x256b): %d / 16384 wrong\n", bad_ss, bad_ts);
    std::vector<Timing> tim(ntim);
    CK(cudaMemcpy(tim.data(), dt, ntim * sizeof(Timing), cudaMemcpyDeviceToHost));
    const char* names[] = {"SS  (A, B from shared memory)", "TS  (A from tensor memory)", "tcgen05.cp.128x256b alone", "1 cp + 3 TS MMAs (conv pattern)"};
    for (auto& t : tim)
        printf("part 2: %-34s N=%3d : %7.1f cycles per instruction (tensor-ideal per MMA %d)\n", names[t.mode], t.n,
               (double)t.cycles / t.count, t.n / 2);
    return (bad_ss || bad_ts) ? 3 : 0;
}
