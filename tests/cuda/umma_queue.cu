// umma_queue.cu — how deep is the tcgen05.mma issue queue?  Times the ISSUE of k back-to-back MMAs
// (clock after the k-th issue, before any wait) against their completion time.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../rusty_sr_b200/csrc/ptx.cuh"
using namespace rsr;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(128, 1) k(long long* out, int N, int* fault) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_s, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_s, 0);
    constexpr uint32_t HI = (uint32_t)make_sdesc_hi(1024, UMMA_SW_128B);
    const uint32_t a_e = (smem_u32(smem) >> 4) | (1u << 16);
    const uint32_t b_e = (smem_u32(smem + 64 * 1024) >> 4) | (1u << 16);
    auto desc = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
    const uint32_t idesc = make_idesc_bf16(128, N);
    if (warp == 0) {
        uint32_t phase = 0;
        for (int kk = 1; kk <= 24; ++kk) {
            __syncwarp();
            long long t0 = 0, t1 = 0;
            if (elect_one()) {
                t0 = clock64();
                for (int i = 0; i < kk; ++i) umma_f16(tmem, desc(HI, a_e + (i & 3) * 8), desc(HI, b_e), idesc, 1);
                t1 = clock64();
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, phase, fault, 1);
            phase ^= 1;
            const long long t2 = clock64();
            if (threadIdx.x == 0) { out[kk * 2] = t1 - t0; }
            t0 = __shfl_sync(0xffffffffu, t0, 0);
            if (threadIdx.x == 0) out[kk * 2 + 1] = t2 - t0;
        }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
    long long* d; int* f; CK(cudaMalloc(&d, 64 * 8)); CK(cudaMalloc(&f, 4)); CK(cudaMemset(f, 0, 4));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024));
    for (int N : {32, 96, 192}) {
        for (int rep = 0; rep < 2; ++rep) { k<<<1, 128, 140 * 1024>>>(d, N, f); CK(cudaDeviceSynchronize()); }
        std::vector<long long> h(64); CK(cudaMemcpy(h.data(), d, 64 * 8, cudaMemcpyDeviceToHost));
        printf("N=%d: k: issue_cycles / complete_cycles\n", N);
        for (int kk = 1; kk <= 24; ++kk) printf("  k=%2d  issue %5lld  complete %5lld\n", kk, h[kk * 2], h[kk * 2 + 1]);
    }
    return 0;
}
