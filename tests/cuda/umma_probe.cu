// umma_probe.cu — hardware probe (run under gpurun) that pins the tcgen05 shared-memory
// descriptor semantics the conv kernels rely on, against an exact CPU expectation:
//   * TMA SWIZZLE_128B / SWIZZLE_64B tile -> K-major UMMA descriptor with the same swizzle
//   * K-advance inside the swizzle row (start address + k*2 bytes)
//   * pixel-shifted start addresses (start + dx * row_bytes): the implicit-GEMM tap shift,
//     with base_offset = 0 (policy 0) or (addr >> 7) & 7 (policy 1)
//   * TMA out-of-bounds zero fill at negative / past-the-end coordinates ("Same" padding)
//   * TMEM accumulator layout as read by tcgen05.ld.32x32b
// All operands are small integers (exact in bf16, sums exact in fp32) so the check is ==.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_bf16.h>
#include <cudaTypedefs.h>
#include "../../rusty_sr_b200/csrc/ptx.cuh"

using namespace rsr;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Case { int a_off; int b_off; int policy; int pad; };

constexpr int PXBOX = 136;   // pixels per A box row
constexpr int NB = 32;       // B rows (N)
constexpr int MAXCASES = 96;

template <int RB>  // row bytes: 128 (SW128) or 64 (SW64)
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const Case* __restrict__ cases, int ncases, int x0, int y0, float* __restrict__ D, int* fault) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                       // PXBOX * RB
    uint8_t* sB = smem + 24 * 1024;           // NB * RB
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar_tma, 1);
        mbar_init(&bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc(&tmem_base_s, 64); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar_tma, PXBOX * RB + NB * RB);
        tma_load_3d(sA, &tmA, &bar_tma, 0, x0, y0);
        tma_load_2d(sB, &tmB, &bar_tma, 0, 0);
    }
    mbar_wait(&bar_tma, 0, fault, 1);
    constexpr uint32_t LAYOUT = RB == 128 ? UMMA_SW_128B : UMMA_SW_64B;
    constexpr uint32_t HI = (uint32_t)make_sdesc_hi(8 * RB, LAYOUT);
    const uint32_t idesc = make_idesc_bf16(128, NB);
    for (int c = 0; c < ncases; ++c) {
        if (threadIdx.x == 0) {
            const uint32_t a_addr = smem_u32(sA) + cases[c].a_off;
            const uint32_t b_addr = smem_u32(sB) + cases[c].b_off;
            const uint32_t boA = cases[c].policy ? ((a_addr >> 7) & 7) : 0;
            const uint32_t boB = cases[c].policy ? ((b_addr >> 7) & 7) : 0;
            tc_fence_after();
            umma_f16(tmem_base, make_sdesc(a_addr, 16, HI, boA), make_sdesc(b_addr, 16, HI, boB), idesc, 0);
            umma_commit(&bar_mma);
        }
        mbar_wait(&bar_mma, c & 1, fault, 2);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
        tmem_ld_wait();
        float* out = D + ((size_t)c * 128 + warp * 32 + lane_id()) * NB;
        #pragma unroll
        for (int j = 0; j < 32; ++j) out[j] = __uint_as_float(v[j]);
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
}

template <int RB>
static int run(const char* label) {
    const int E = RB / 2;           // bf16 elements per pixel row
    const int W = 150, H = 3;       // global image: narrower than x0+PXBOX so the right edge is OOB too
    const int x0 = -2, y0 = 1;
    std::vector<__nv_bfloat16> hA((size_t)H * W * E), hB((size_t)NB * E);
    std::vector<float> fA(hA.size()), fB(hB.size());
    srand(1234 + RB);
    for (size_t i = 0; i < hA.size(); ++i) { int v = rand() % 17 - 8; fA[i] = (float)v; hA[i] = __float2bfloat16((float)v); }
    for (size_t i = 0; i < hB.size(); ++i) { int v = rand() % 9 - 4; fB[i] = (float)v; hB[i] = __float2bfloat16((float)v); }
    __nv_bfloat16 *dA, *dB; float* dD; Case* dC; int* dF;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    std::vector<Case> cases;
    for (int policy = 0; policy < 2; ++policy)
        for (int dx = 0; dx <= 8; ++dx)
            for (int k = 0; k < E; k += 16)
                if ((int)cases.size() < MAXCASES) cases.push_back({dx * RB + k * 2, ((k + 16) % E) * 2, policy, 0});
    const int nc = (int)cases.size();
    CK(cudaMalloc(&dC, nc * sizeof(Case))); CK(cudaMemcpy(dC, cases.data(), nc * sizeof(Case), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&dD, (size_t)nc * 128 * NB * 4)); CK(cudaMemset(dD, 0xFF, (size_t)nc * 128 * NB * 4));
    CK(cudaMalloc(&dF, 4)); CK(cudaMemset(dF, 0, 4));

    auto enc = get_encode();
    CUtensorMap tmA, tmB;
    const CUtensorMapSwizzle sw = RB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    {
        cuuint64_t dims[3] = {(cuuint64_t)E, (cuuint64_t)W, (cuuint64_t)H};
        cuuint64_t strides[2] = {(cuuint64_t)RB, (cuuint64_t)RB * W};
        cuuint32_t box[3] = {(cuuint32_t)E, PXBOX, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dA, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)E, (cuuint64_t)NB};
        cuuint64_t strides[1] = {(cuuint64_t)RB};
        cuuint32_t box[2] = {(cuuint32_t)E, NB};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
    }
    const int smem_bytes = 24 * 1024 + NB * RB + 2048;
    CK(cudaFuncSetAttribute(probe_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    probe_kernel<RB><<<1, 128, smem_bytes>>>(tmA, tmB, dC, nc, x0, y0, dD, dF);
    cudaError_t e = cudaDeviceSynchronize();
    int fault = 0; cudaMemcpy(&fault, dF, 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("[%s] kernel error: %s fault=0x%x\n", label, cudaGetErrorString(e), fault); return 1; }
    std::vector<float> hD((size_t)nc * 128 * NB);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    int ok_by_policy[2] = {0, 0}, n_by_policy[2] = {0, 0};
    for (int c = 0; c < nc; ++c) {
        const int dx = cases[c].a_off / RB, kA = (cases[c].a_off % RB) / 2, kB = cases[c].b_off / 2;
        double maxerr = 0; int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < NB; ++n) {
                float ref = 0;
                const int gx = x0 + m + dx;
                if (gx >= 0 && gx < W && m + dx < PXBOX)
                    for (int k = 0; k < 16; ++k) ref += fA[((size_t)y0 * W + gx) * E + kA + k] * fB[(size_t)n * E + kB + k];
                const float got = hD[((size_t)c * 128 + m) * NB + n];
                const double err = fabs((double)got - ref);
                if (!(err <= 0)) bad++;
                if (err > maxerr || err != err) maxerr = err;
            }
        printf("[%s] case %2d policy=%d dx=%d kA=%2d kB=%2d : %s (bad=%d maxerr=%g)\n", label, c, cases[c].policy, dx, kA, kB,
               bad ? "FAIL" : "ok", bad, maxerr);
        n_by_policy[cases[c].policy]++; ok_by_policy[cases[c].policy] += bad == 0;
    }
    printf("[%s] SUMMARY policy0(base_offset=0): %d/%d ok ; policy1(base_offset=(addr>>7)&7): %d/%d ok\n", label,
           ok_by_policy[0], n_by_policy[0], ok_by_policy[1], n_by_policy[1]);
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dD); cudaFree(dF);
    return 0;
}

int main() {
    int rc = run<128>("SW128");
    rc |= run<64>("SW64");
    return rc;
}
