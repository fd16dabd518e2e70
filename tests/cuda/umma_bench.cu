// umma_bench.cu — micro-benchmark (run under gpurun): cycles per tcgen05.mma (kind::f16, M=128, K=16)
// as a function of N and of the operand pattern, from one CTA.  Guides the tile-shape decisions in
// DESIGN.md (is a small-N MMA bound by the tensor pipe, by the shared-memory operand read, or by a
// fixed per-instruction cost?).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../rusty_sr_b200/csrc/ptx.cuh"
using namespace rsr;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Result { long long cycles; int n, pattern, count; };

// pattern 0: same A, same B, same D every time          (pure issue/tensor rate)
// pattern 1: BF16x3 triple: (A0,B0) (A1,B0) (A0,B1) into one D, A shifted by 128 B per triple (tap walk)
// pattern 2: like 1 but two accumulators alternate (two rows per tap, as the kernel does)
// pattern 3: N-concat variant: (A0, B[0..2N)) into D[0..2N), then (A1, B0) into D[0..N)
__global__ void __launch_bounds__(128, 1) bench_kernel(Result* out, int nres, const int* ns, const int* pats, int count, int* fault) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;  // 192 KB of zeros
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) { tmem_alloc(&tmem_s, 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_s, 0);
    constexpr uint32_t HI = (uint32_t)make_sdesc_hi(1024, UMMA_SW_128B);
    const uint32_t a_e = (smem_u32(smem) >> 4) | (1u << 16);                 // A region: 64 KB
    const uint32_t b_e = (smem_u32(smem + 96 * 1024) >> 4) | (1u << 16);     // B region: 64 KB
    auto desc = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
    if (warp == 0) {
        uint32_t phase = 0;
        for (int r = 0; r < nres; ++r) {
            const int N = ns[r], pat = pats[r];
            const uint32_t idesc = make_idesc_bf16(128, N);
            const uint32_t idesc2 = make_idesc_bf16(128, 2 * N);
            __syncwarp();
            const long long t0 = clock64();
            if (elect_one()) {
                if (pat == 0) {
                    for (int i = 0; i < count; ++i) umma_f16(tmem, desc(HI, a_e), desc(HI, b_e), idesc, 1);
                } else if (pat == 1) {
                    for (int i = 0; i < count / 3; ++i) {
                        const uint32_t a = a_e + (i & 3) * 8 + (i & 1) * 2;
                        umma_f16(tmem, desc(HI, a), desc(HI, b_e), idesc, 1);
                        umma_f16(tmem, desc(HI, a + 4), desc(HI, b_e), idesc, 1);
                        umma_f16(tmem, desc(HI, a), desc(HI, b_e + 4), idesc, 1);
                    }
                } else if (pat == 2) {
                    for (int i = 0; i < count / 6; ++i) {
                        const uint32_t a = a_e + (i & 3) * 8 + (i & 1) * 2;
                        #pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const uint32_t aa = a + rr * (17408 >> 4), d = tmem + rr * 256;
                            umma_f16(d, desc(HI, aa), desc(HI, b_e), idesc, 1);
                            umma_f16(d, desc(HI, aa + 4), desc(HI, b_e), idesc, 1);
                            umma_f16(d, desc(HI, aa), desc(HI, b_e + 4), idesc, 1);
                        }
                    }
                } else {
                    for (int i = 0; i < count / 2; ++i) {
                        const uint32_t a = a_e + (i & 3) * 8 + (i & 1) * 2;
                        umma_f16(tmem, desc(HI, a), desc(HI, b_e), idesc2, 1);
                        umma_f16(tmem, desc(HI, a + 4), desc(HI, b_e), idesc, 1);
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

int main() {
    std::vector<int> ns, pats;
    const int Ns[] = {32, 64, 96, 128, 192, 256};
    for (int p = 0; p < 3; ++p) for (int n : Ns) { ns.push_back(n); pats.push_back(p); }
    for (int n : {32, 64, 96, 128}) { ns.push_back(n); pats.push_back(3); }
    const int nres = (int)ns.size(), count = 1200;
    int *dn, *dp, *df; Result* dr;
    CK(cudaMalloc(&dn, nres * 4)); CK(cudaMalloc(&dp, nres * 4)); CK(cudaMalloc(&df, 4)); CK(cudaMalloc(&dr, nres * sizeof(Result)));
    CK(cudaMemcpy(dn, ns.data(), nres * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dp, pats.data(), nres * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(df, 0, 4));
    const int smem = 200 * 1024;
    CK(cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int rep = 0; rep < 2; ++rep) {
        bench_kernel<<<1, 128, smem>>>(dr, nres, dn, dp, count, df);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("kernel error %s\n", cudaGetErrorString(e)); return 1; }
    }
    std::vector<Result> res(nres);
    CK(cudaMemcpy(res.data(), dr, nres * sizeof(Result), cudaMemcpyDeviceToHost));
    const char* names[] = {"same-operands", "bf16x3 triple (1 acc)", "bf16x3 triple (2 acc, as kernel)", "n-concat pair"};
    for (auto& r : res) {
        double per = (double)r.cycles / r.count;
        double ideal = r.pattern == 3 ? (r.n * 1.5 / 2.0) / 2.0 : r.n / 2.0;  // tensor cycles per MMA (avg)
        printf("pattern %d %-34s N=%3d : %8.1f cycles/MMA   (tensor-ideal %.0f, ratio %.2f)\n", r.pattern, names[r.pattern], r.n, per, ideal, per / ideal);
    }
    return 0;
}
