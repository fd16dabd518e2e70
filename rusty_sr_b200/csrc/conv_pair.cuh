// conv_pair.cuh — the conv_strip kernel run by CTA PAIRS (tcgen05 cta_group::2, M = 256).
//
// Same formulation, data layout, epilogues and results as conv_strip.cuh (the implicit-GEMM convolution of the
// sr_net forward pass, reference src/network.rs:33-72); what changes is who feeds the tensor core.  Measured on
// B200 (profiles/README.md, r01i): with BF16x3 at N <= 96 the conv launches are bound by the 128 B/cycle
// shared-memory port — a tcgen05.mma re-reads its 4 KB A tile and its N x 32 B weight tile from shared memory for
// each of the three split products, and the weight stages are re-streamed into shared memory for every R rows.
// A CTA pair halves the weight side of that traffic: the two CTAs of a cluster (two SMs of one TPC) each keep
// only HALF of every weight tap (N/2 rows), one tcgen05.mma.cta_group::2 of the leader CTA multiplies both CTAs'
// pixel tiles (M = 2 x 128) with the weights gathered from both halves, and each CTA's accumulator lands in its
// own tensor memory.
//
// Work split: a cluster owns a (strip, row range) work item like a single CTA does in conv_strip.cuh; the leader
// takes the upper half of the rows and the peer the lower half, so both walk down the same strip in lock step —
// same number of passes, same ring slots at every step (the A and B descriptors of a cta_group::2 MMA are
// shared-memory OFFSETS applied in both CTAs).  Each CTA has its own row ring, partial-sum ring, epilogue and
// TMA stores, exactly as in conv_strip.cuh.
//
// Cross-CTA hand-offs:  * input rows / weight halves: each CTA's TMA (cta_group::2 form) signals complete_tx on
//                         the LEADER's full barrier, which expects the bytes of both CTAs;
//                       * slot release and "accumulator ready": tcgen05.commit.cta_group::2 multicast to the same
//                         barrier in both CTAs;
//                       * "accumulator drained": the epilogue warps of both CTAs arrive on the leader's t_empty
//                         (remote arrive with cluster scope from the peer).
#pragma once
#include "conv_strip.cuh"

namespace rsr {

// ---- cluster / cta_group::2 PTX ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, int* g_fault, int tag) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > (1u << 24)) mbar_timeout(bar, parity, g_fault, tag);
    }
}
// TMA loads whose completion is signalled on a barrier of the CTA pair (here: always the leader's)
__device__ __forceinline__ void tma2_load_2d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t smem_dst, const CUtensorMap* m, uint32_t bar_cluster, int c0, int c1,
                                             int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (count 1) on the barrier at this shared-memory offset in BOTH CTAs once all prior MMAs of this thread completed
__device__ __forceinline__ void umma2_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}

struct PairIter {
    long long cur, end;
    int rows;
};
template <class C>
__device__ __forceinline__ bool pair_item_next(const ConvParams& p, PairIter& it, uint32_t rank, WorkItem& w) {
    if (it.cur >= it.end) return false;
    const int col = (int)(it.cur / it.rows);
    const int y0 = (int)(it.cur % it.rows);
    const int n = (int)min((long long)(it.rows - y0), it.end - it.cur);
    const int l0 = (n + 1) / 2;            // leader: upper half, peer: the rest (possibly empty)
    w.img = col / p.n_strips;
    w.x0 = (col % p.n_strips) * 128;
    w.y_b = p.row_begin + y0 + (rank ? l0 : 0);
    w.y_e = p.row_begin + y0 + (rank ? n : l0);
    w.npass = (l0 + C::R - 1) / C::R;      // common to both CTAs: they run in lock step
    it.cur += n;
    return true;
}

template <class C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C::NTHREADS, 1) conv_pair_kernel(const __grid_constant__ ConvParams p) {
    static_assert(C::PAIR, "conv_pair_kernel needs a ConvCfg with PAIR = true (half weight taps)");
    using P = C;   // shared-memory plan: ConvCfg computes it with half a weight tap per CTA
    extern __shared__ uint8_t smem_raw[];
    // (the dynamic shared-memory window starts at the same offset in both CTAs, so the aligned base does too)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + P::OFF_B;
    uint8_t* sP = smem + P::OFF_P;
    uint8_t* sO = smem + P::OFF_O;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + P::OFF_BAR);
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
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    int pin_g[3], npg = 0;
    #pragma unroll
    for (int g = 0; g < C::NG; ++g)
        if (C::NPIN > 0 && p.grp[g].has_pin) pin_g[npg++] = g;
    const int n_pin = npg * C::R;

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NR; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < C::NBS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < NPB; ++i) { mbar_init(&p_full[i], 1); mbar_init(&p_empty[i], C::NEPI); }
        // t_empty (used in the leader only): the epilogue warps of BOTH CTAs arrive
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 2 * C::NEPI); }
        fence_mbar_init();
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB);
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();     // both CTAs' barriers are initialised before any cross-CTA signal
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

    PairIter it;
    {
        const long long ncl = gridDim.x >> 1, cid = blockIdx.x >> 1;
        it.rows = p.row_end - p.row_begin;
        it.cur = p.total_rowtiles * cid / ncl;
        it.end = p.total_rowtiles * (cid + 1) / ncl;
    }
    WorkItem w;

    if (warp == 0) {
        // ===== input-row producer (each CTA its own rows; completion signalled on the LEADER's a_full) =====
        const uint32_t a_full0 = map_to_cta(smem_u32(a_full), 0);
        uint32_t slot = 0, par = 1;
        while (pair_item_next<C>(p, it, rank, w)) {
            const int nrows = w.npass * C::R + 2 * C::PH;
            const int y0 = w.y_b - C::PH;
            for (int i = 0; i < nrows; ++i) {
                mbar_wait(&a_empty[slot], par, p.fault, 1);
                if (elect_one()) {
                    if (leader) mbar_arrive_expect_tx(&a_full[slot], 2 * C::ROW_BYTES);
                    tma2_load_4d(smem_u32(sA + slot * C::SLOT), &p.tmA, a_full0 + slot * 8, 0, w.x0 - C::PW, y0 + i, w.img);
                }
                __syncwarp();
                if (++slot == C::NR) { slot = 0; par ^= 1; }
            }
        }
    } else if (warp == 3) {
        // ===== weight producer: this CTA's half (rows rank*HB .. +HB) of every tap of the stage =====
        const uint32_t b_full0 = map_to_cta(smem_u32(b_full), 0);
        uint32_t st = 0, par = 1;
        while (pair_item_next<C>(p, it, rank, w)) {
            for (int j = 0; j < w.npass; ++j)
                for (int sp = 0; sp < C::NSTG; ++sp) {
                    mbar_wait(&b_empty[st], par, p.fault, 2);
                    if (elect_one()) {
                        const int t0 = sp * C::TPS;
                        const int nt = C::TAPS - t0 < C::TPS ? C::TAPS - t0 : C::TPS;   // taps really in this stage
                        if (leader) mbar_arrive_expect_tx(&b_full[st], 2 * nt * P::TAP_BYTES);
                        for (int tt = 0; tt < nt; ++tt)
                            tma2_load_2d(smem_u32(sB + st * P::BST + tt * P::TAP_BYTES), &p.tmB, b_full0 + st * 8, 0,
                                         (t0 + tt) * C::NCOL + (int)rank * P::HB);
                    }
                    __syncwarp();
                    if (++st == C::NBS) { st = 0; par ^= 1; }
                }
        }
    } else if (warp == 2) {
        // ===== partial-sum tile producer (local to each CTA) =====
        if (C::NPIN > 0 && n_pin > 0) {
            uint32_t st = 0, par = 1;
            while (pair_item_next<C>(p, it, rank, w)) {
                for (int j = 0; j < w.npass; ++j)
                    for (int r = 0; r < C::R; ++r)
                        for (int k = 0; k < npg; ++k) {
                            mbar_wait(&p_empty[st], par, p.fault, 7);
                            if (elect_one()) {
                                mbar_arrive_expect_tx(&p_full[st], C::PIN_BYTES);
                                tma_load_4d(sP + st * C::PIN_BYTES, &p.tmPin[pin_g[k]], &p_full[st], 0, w.x0,
                                            w.y_b + j * C::R + r, w.img);
                            }
                            __syncwarp();
                            if (++st == NPB) { st = 0; par ^= 1; }
                        }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: leader CTA only; every MMA multiplies both CTAs' pixel tiles (M = 256) =====
        if (leader) {
            constexpr uint32_t HI = (uint32_t)make_sdesc_hi(8 * C::RB, C::SWZ);
            constexpr uint32_t IDESC = make_idesc_bf16(256, C::NCOL);
            constexpr uint32_t LBO = 1u << 16;
            constexpr uint32_t MID = (C::RB / 2) >> 4;
            constexpr int WIN = C::R + 2 * C::PH;
            constexpr int NEARLY = C::R < C::KH ? C::R : C::KH;
            const uint32_t sA_e = (smem_u32(sA) >> 4) | LBO, sB_e = (smem_u32(sB) >> 4) | LBO;
            auto desc = [](uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 32) | lo; };
            uint32_t slot0 = 0;
            uint32_t wslot = 0, wpar = 0;
            uint32_t bst = 0, bpar = 0;
            uint32_t buf = 0, tpar = 1;
            while (pair_item_next<C>(p, it, rank, w)) {
                for (int j = 0; j < w.npass; ++j) {
                    mbar_wait_cluster(&t_empty[buf], tpar, p.fault, 3);
                    const int nnew = j == 0 ? WIN : C::R;
                    for (int i = 0; i < nnew; ++i) {
                        mbar_wait(&a_full[wslot], wpar, p.fault, 4);
                        if (++wslot == C::NR) { wslot = 0; wpar ^= 1; }
                    }
                    tc_fence_after();
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
                    #pragma unroll
                    for (int sp = 0; sp < C::NSTG; ++sp) {
                        mbar_wait(&b_full[bst], bpar, p.fault, 5);
                        tc_fence_after();
                        const uint32_t b_st = sB_e + bst * (P::BST >> 4);
                        if (elect_one()) {
                            #pragma unroll
                            for (int tt = 0; tt < C::TPS; ++tt) {
                                const int t = sp * C::TPS + tt;
                                if (t < C::TAPS) {
                                    const int ta = t / C::KW, tb = t % C::KW;
                                    const uint32_t b_e = b_st + tt * (P::TAP_BYTES >> 4);
                                    #pragma unroll
                                    for (int r = 0; r < C::R; ++r) {
                                        const uint32_t a_e = arow[r + ta] + tb * (C::RB >> 4);
                                        const uint32_t d = d_pass + r * C::NCOL;
                                        #pragma unroll
                                        for (int c = 0; c < C::KC; ++c) {
                                            umma2_f16(d, desc(HI, a_e + 2 * c), desc(HI, b_e + 2 * c), IDESC, (t | c) != 0 ? 1u : 0u);
                                            umma2_f16(d, desc(HI, a_e + MID + 2 * c), desc(HI, b_e + 2 * c), IDESC, 1);
                                            umma2_f16(d, desc(HI, a_e + 2 * c), desc(HI, b_e + MID + 2 * c), IDESC, 1);
                                        }
                                    }
                                    if (tb == C::KW - 1 && ta < NEARLY) {
                                        uint32_t sl = slot0 + ta;
                                        if (sl >= C::NR) sl -= C::NR;
                                        umma2_commit(&a_empty[sl]);
                                    }
                                }
                            }
                            umma2_commit(&b_empty[bst]);
                        }
                        __syncwarp();
                        if (++bst == C::NBS) { bst = 0; bpar ^= 1; }
                    }
                    const bool last = j == w.npass - 1;
                    const int n_retire = last ? WIN : C::R;
                    if (elect_one()) {
                        umma2_commit(&t_full[buf]);
                        uint32_t sl = slot0 + NEARLY;
                        if (sl >= C::NR) sl -= C::NR;
                        for (int i = NEARLY; i < n_retire; ++i) {
                            umma2_commit(&a_empty[sl]);
                            if (++sl == C::NR) sl = 0;
                        }
                    }
                    __syncwarp();
                    slot0 += n_retire;
                    if (slot0 >= C::NR) slot0 -= C::NR;
                    if (slot0 >= C::NR) slot0 -= C::NR;
                    buf ^= 1;
                    if (buf == 0) tpar ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (per CTA, as in conv_strip.cuh; "drained" is signalled to the leader's t_empty) =====
        const int e = warp - 4;
        const int ew = e & 3;
        const int eh = e >> 2;
        const uint32_t stage0 = smem_u32(sO) + e * (C::NOUT * 4096);
        const uint32_t t_empty0 = map_to_cta(smem_u32(t_empty), 0);
        uint32_t buf = 0, tpar = 0, oc = 0;
        uint32_t pst = 0, ppar = 0;
        const uint32_t sP_u = smem_u32(sP);
        while (pair_item_next<C>(p, it, rank, w)) {
            const int xw = w.x0 + ew * 32;
            const int x = xw + lane;
            for (int j = 0; j < w.npass; ++j, buf ^= 1, tpar ^= (buf == 0)) {
                SkipNb nb;
                if (C::SHUF) {
                    const int y = w.y_b + j * C::R + eh;
                    if (y < w.y_e && x < p.W) load_skip_nb(p, w.img, y, x, nb);
                }
                mbar_wait(&t_full[buf], tpar, p.fault, 6);
                tc_fence_after();
                #pragma unroll 1
                for (int r = 0; r < C::R; ++r) {
                    const int y = w.y_b + j * C::R + r;
                    const bool row_live = y < w.y_e && xw < p.W;
                    #pragma unroll 1
                    for (int g = 0; g < C::NG; ++g) {
                        const EpiGroup& eg = p.grp[g];
                        if (eg.kind == EPI_NONE) continue;
                        const bool pin = C::NPIN > 0 && eg.has_pin;
                        const bool mine = C::NEPI == 4 || (((r * C::NG + g) & 1) == eh);
                        if (!mine || !row_live) {
                            // every epilogue warp observes every partial-sum tile (see conv_strip.cuh)
                            if (pin) {
                                mbar_wait(&p_full[pst], ppar, p.fault, 9);
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&p_empty[pst]);
                                if (++pst == NPB) { pst = 0; ppar ^= 1; }
                            }
                            continue;
                        }
                        uint32_t raw[32];
                        tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + buf * (C::R * C::NCOL) + r * C::NCOL + g * 32, raw);
                        tmem_ld_wait();
                        float v[32];
                        #pragma unroll
                        for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(raw[k]);
                        if (pin) {
                            mbar_wait(&p_full[pst], ppar, p.fault, 9);
                            epi_add_partial_tile(v, sP_u + pst * C::PIN_BYTES, ew * 32 + lane);
                            __syncwarp();
                            if (lane == 0) mbar_arrive(&p_empty[pst]);
                            if (++pst == NPB) { pst = 0; ppar ^= 1; }
                        }
                        const uint32_t stage = stage0 + (oc % C::NOUT) * 4096;
                        if (elect_one())
                            asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(C::NOUT - 1) : "memory");
                        __syncwarp();
                        if (!C::SHUF) {
                            if (eg.kind == EPI_RAW) epi_raw_to_stage(v, stage, lane);
                            else epi_split_to_stage(eg, v, stage, lane);
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
                        } else {
                            if (r > eh && x < p.W) load_skip_nb(p, w.img, y, x, nb);
                            {
                                const float4* b4 = reinterpret_cast<const float4*>(eg.bias);
                                #pragma unroll
                                for (int i = 0; i < 8; ++i) {
                                    const float4 t = __ldg(b4 + i);
                                    v[4 * i + 0] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
                                }
                            }
                            float o[3][9];
                            if (x >= p.W) {
                                #pragma unroll
                                for (int a = 0; a < 3; ++a)
                                    #pragma unroll
                                    for (int b = 0; b < 3; ++b) { nb.v[a][b][0] = 0.f; nb.v[a][b][1] = 0.f; nb.v[a][b][2] = 0.f; }
                            }
                            shuffle_pixel(v, nb, o);
                            const int npx_valid = min(32, p.W - xw);
                            const size_t W3 = (size_t)p.W * 3;
                            const size_t orow0 = ((size_t)w.img * p.H * 3 + (size_t)y * 3) * W3 + (size_t)xw * 3;
                            if (p.out_f32) {
                                float* sf = reinterpret_cast<float*>(smem + P::OFF_O) + (e * C::NOUT + (oc % C::NOUT)) * 1024;
                                #pragma unroll
                                for (int dy = 0; dy < 3; ++dy)
                                    #pragma unroll
                                    for (int k = 0; k < 9; ++k) sf[dy * 288 + lane * 9 + k] = o[dy][k];
                                __syncwarp();
                                const int nfl = npx_valid * 9;
                                #pragma unroll
                                for (int dy = 0; dy < 3; ++dy) {
                                    float* dst = p.out_f32 + (orow0 + (size_t)dy * W3) * 3;
                                    if (npx_valid == 32 && p.out_vec_ok) {
                                        for (int i = lane; i < 72; i += 32)
                                            reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(sf + dy * 288)[i];
                                    } else {
                                        for (int i = lane; i < nfl; i += 32) dst[i] = sf[dy * 288 + i];
                                    }
                                }
                            }
                            if (p.out_u8) {
                                uint32_t* sq = reinterpret_cast<uint32_t*>(smem + P::OFF_O) + (e * C::NOUT + (oc % C::NOUT)) * 1024;
                                #pragma unroll
                                for (int dy = 0; dy < 3; ++dy)
                                    #pragma unroll
                                    for (int dx = 0; dx < 3; ++dx) {
                                        const uint32_t r8 = (uint32_t)(fminf(fmaxf(o[dy][dx * 3 + 0], 0.f), 1.f) * 255.0f + 0.5f);
                                        const uint32_t g8 = (uint32_t)(fminf(fmaxf(o[dy][dx * 3 + 1], 0.f), 1.f) * 255.0f + 0.5f);
                                        const uint32_t b8 = (uint32_t)(fminf(fmaxf(o[dy][dx * 3 + 2], 0.f), 1.f) * 255.0f + 0.5f);
                                        sq[dy * 96 + lane * 3 + dx] = r8 | (g8 << 8) | (b8 << 16) | 0xFF000000u;
                                    }
                                __syncwarp();
                                const int nu = npx_valid * 3;
                                #pragma unroll
                                for (int dy = 0; dy < 3; ++dy) {
                                    uint32_t* dst = reinterpret_cast<uint32_t*>(p.out_u8) + orow0 + (size_t)dy * W3;
                                    for (int i = lane; i < nu; i += 32) dst[i] = sq[dy * 96 + i];
                                }
                            }
                            __syncwarp();
                        }
                        ++oc;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(t_empty0 + buf * 8);
            }
        }
        if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    // neither CTA may leave (or free its tensor memory) while the pair's MMAs / remote arrivals can still touch it
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
}

}  // namespace rsr
