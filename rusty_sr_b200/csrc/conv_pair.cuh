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
// The generic 3x3 convolution (rsr_conv2d_3x3_f32, 128 output channels per launch) runs here too: several input-channel
// blocks (ConvParams::n_cb) are accumulated per pass, each from its own window of rows, as in conv_strip.cuh.
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
    const int n_cb = C::MCB ? p.n_cb : 1;   // (a compile-time 1 for the sr_net launch: its loops stay as they were)

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
    // a cta_group::2 allocation is performed for the pair: either CTA's tcgen05.alloc may write the (common) address into
    // both CTAs' tmem_slot, so read it only after both allocations have completed (compute-sanitizer racecheck, r02)
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    init_scale_factors<C>(tmem_base, warp);   // (each CTA its own tensor memory; before the second cluster barrier)
    tc_fence_before();
    cluster_sync_all();     // both CTAs' barriers are initialised (and scale factors written) before any cross-CTA signal
    tc_fence_after();
    griddep_launch_dependents();   // programmatic dependent launch (see conv_strip.cuh)
    griddep_wait();

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
            // n_cb == 1: one sliding window over the item's rows.  n_cb > 1 (generic conv): every (pass, channel block)
            // segment loads its own window of R + 2*PH rows from plane img * n_cb + cb, as in conv_strip.cuh.
            const int nseg = n_cb == 1 ? 1 : w.npass * n_cb;
            const int nrows = n_cb == 1 ? w.npass * C::R + 2 * C::PH : C::R + 2 * C::PH;
            for (int sg = 0; sg < nseg; ++sg) {
                const int cb = n_cb == 1 ? 0 : sg % n_cb;
                const int y0 = w.y_b - C::PH + (n_cb == 1 ? 0 : (sg / n_cb) * C::R);
                for (int i = 0; i < nrows; ++i) {
                    mbar_wait(&a_empty[slot], par, p.fault, 1);
                    if (elect_one()) {
                        if (leader) mbar_arrive_expect_tx(&a_full[slot], 2 * C::ROW_BYTES);
                        tma2_load_4d(smem_u32(sA + slot * C::SLOT), &p.tmA, a_full0 + slot * 8, 0, w.x0 - C::PW, y0 + i, w.img * n_cb + cb);
                    }
                    __syncwarp();
                    if (++slot == C::NR) { slot = 0; par ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ===== weight producer: this CTA's half (rows rank*HB .. +HB) of every tap of the stage =====
        const uint32_t b_full0 = map_to_cta(smem_u32(b_full), 0);
        uint32_t st = 0, par = 1;
        while (pair_item_next<C>(p, it, rank, w)) {
            for (int j = 0; j < w.npass; ++j)
                for (int cb = 0; cb < n_cb; ++cb)
                for (int sp = 0; sp < C::NSTG; ++sp) {
                    mbar_wait(&b_empty[st], par, p.fault, 2);
                    if (elect_one()) {
                        const int t0 = sp * C::TPS;
                        const int nt = C::TAPS - t0 < C::TPS ? C::TAPS - t0 : C::TPS;   // taps really in this stage
                        if (leader) mbar_arrive_expect_tx(&b_full[st], 2 * nt * P::TAP_BYTES);
                        for (int tt = 0; tt < nt; ++tt)
                            tma2_load_2d(smem_u32(sB + st * P::BST + tt * P::TAP_BYTES), &p.tmB, b_full0 + st * 8, 0,
                                         (cb * C::TAPS + t0 + tt) * C::NCOL + (int)rank * P::HB);
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
            using MM = SplitMma<C::PREC, 2, C::RB>;
            constexpr uint32_t LBO = 1u << 16;
            constexpr int WIN = C::R + 2 * C::PH;
            constexpr int NEARLY = C::R < C::KH ? C::R : C::KH;
            const uint32_t sA_e = (smem_u32(sA) >> 4) | LBO, sB_e = (smem_u32(sB) >> 4) | LBO;
            const uint32_t sf = tmem_base + C::SF_COL;
            uint32_t slot0 = 0;
            uint32_t wslot = 0, wpar = 0;
            uint32_t bst = 0, bpar = 0;
            uint32_t buf = 0, tpar = 1;
            while (pair_item_next<C>(p, it, rank, w)) {
                for (int j = 0; j < w.npass; ++j) {
                    mbar_wait_cluster(&t_empty[buf], tpar, p.fault, 3);
                    for (int cb = 0; cb < n_cb; ++cb) {
                    // rows this segment adds to the window: all of it for the first pass of an item and for every channel
                    // block of the generic conv, else R
                    const int nnew = (j == 0 || n_cb > 1) ? WIN : C::R;
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
                                            if (ABL_NO_MMA) continue;
                                            MM::chunk(d, a_e, b_e, c, 256, C::NCOL, sf, (t | c) != 0 ? 1u : (uint32_t)(cb != 0));
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
                    const bool last = j == w.npass - 1 || n_cb > 1;
                    const int n_retire = last ? WIN : C::R;
                    if (elect_one()) {
                        if (cb == n_cb - 1) umma2_commit(&t_full[buf]);
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
                    }  // channel blocks
                    buf ^= 1;
                    if (buf == 0) tpar ^= 1;
                }
            }
        }
    } else if (warp >= 4) {
        // ===== epilogue (per CTA, conv_strip.cuh; "drained" is signalled to the LEADER's t_empty) =====
        const uint32_t t_empty0 = map_to_cta(smem_u32(t_empty), 0);
        long long prof_acc[8];   // (the pair kernel carries no wait profile; dead)
        conv_epilogue_role<C>(p, smem, p_full, p_empty, t_full, tmem_base, warp, lane, prof_acc,
                              [&](WorkItem& wi) { return pair_item_next<C>(p, it, rank, wi); },
                              [&](uint32_t buf) { mbar_arrive_cluster(t_empty0 + buf * 8); });
    }
    // neither CTA may leave (or free its tensor memory) while the pair's MMAs / remote arrivals can still touch it
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
}

}  // namespace rsr
