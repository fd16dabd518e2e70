// conv_fused.cuh — two levels of the sr_net DAG in ONE launch, the middle node kept in shared memory.
//
// Replaces the two launches `conv5|6|7` and `conv8|9` of conv_strip.cuh (reference src/network.rs:65-70
// with Bias :51-52 and BeLU :55-56):
//     l2 = BeLU(conv2(f) + conv5(l1) + l2_bias)                        (network.rs:61,65,51,55)
//     l3 = BeLU(conv3(f) + conv6(l1) + conv8(l2) + l3_bias)            (network.rs:62,66,69,52,56)
//     E  = conv7(l1) + conv9(l2)            (partial of `expand`;      network.rs:67,70)
// Inputs from HBM: l1 (split storage format, conv_strip.cuh), P2 = conv2(f), P3 = conv3(f) (fp32, written by the conv1|2|3
// launch).  Outputs to HBM: l3 (split format) and E (fp32).  l2, the updated P3 and the first half of E never leave
// the SM: 640 B per pixel of HBM traffic instead of 1408 B for the two separate launches.
//
// How: a CTA walks down a 124-pixel-wide column strip, one image row per step.
//   S1(y): tcgen05.mma M=128 N=96 over the 3x3 taps of l1 rows y-1..y+1 -> TMEM slot T[y] = [l2pre | P3acc | Eacc].
//   E1(y): 8 warps: l2pre (+ P2 tile) -> bias, BeLU, split -> row y of a 3-slot l2 ring in shared memory, written
//          in exactly the swizzled layout a TMA load would have produced (zero outside the image = Padding::Same).
//   S2(y): tcgen05.mma N=64 over the taps of l2 rows y-1..y+1, ACCUMULATING onto T[y][32:96] (conv8 onto the
//          conv6 partial, conv9 onto the conv7 partial) — the partial sums never leave TMEM.
//   E2(y): 8 more warps: T[y][32:64] (+ P3 tile) -> bias, BeLU, split -> l3 (global);  T[y][64:96] -> E (global).
// 128 l1 pixels give 126 l2 pixels give 124 output pixels per strip, so a strip recomputes 2 of 126 l2 columns;
// a work item also recomputes one l2 row above and below its row range.  MMA row (= TMEM lane) m stands for
// pixel x0 - 1 + m in BOTH levels — S2 accumulates onto S1's columns, so the lanes must agree: S1 reads l1 ring
// pixels m..m+2, S2 reads l2 ring pixels m-1..m+1, and the strip's outputs are lanes 1..124 (lane 0 and lanes
// 125..127 see out-of-window pixels and are discarded).  The MMA warp issues S1 three rows ahead of S2 so that
// E1 runs under the MMAs.
//
// Warp roles (21 warps): 0 l1-row producer (TMA), 1 MMA issuer, 2 TMEM allocator + P2 producer, 3 weight
// producer, 4-11 E1, 12-19 E2, 20 P3 producer.  All hand-offs are mbarriers and every wait is bounded (ptx.cuh
// mbar_wait).  The epilogues are the critical path of this kernel (r02 ablations: without their math it runs 250 us
// instead of 450): each pixel's 32 channels are therefore split over TWO warps per TMEM lane quarter (16 channels =
// one K chunk = 64 contiguous bytes of the pixel row each); every E2 warp stages its half rows of l3 and of E in two
// 2 KB tiles of its own and stores them with TMA (64-byte-wide boxes), so the bulk-store wait before a tile is re-written
// is always for a store issued a whole row earlier.  (Storing straight from registers was tried: the 16-byte-per-lane
// strided stores cost 63 us of load/store-unit time, profiles/r02 A/B.)
//
// CTA-pair variant (template parameter PAIR; tcgen05 cta_group::2 as in conv_pair.cuh): the kernel is bound by the
// shared-memory port — with one row per step every weight tap is re-streamed into shared memory and re-read by
// the MMAs for every row — so a cluster of two CTAs shares the weights: each CTA holds half of every tap, the
// leader's MMAs (M = 256) process the row of both CTAs at once.  The two CTAs take the upper / lower half of the
// cluster's row range of one strip and run in lock step; l1 rows and weight halves signal the leader's barriers,
// commits are multicast, the peer's E1 / E2 warps arrive remotely on the leader's l2_full / t_empty.
#pragma once
#include "conv_pair.cuh"

namespace rsr {

struct FusedParams {
    CUtensorMap tmA;      // l1: rank 4 (64, W, H, NIMG) bf16, box (64, 128, 1, 1), SWIZZLE_128B
    CUtensorMap tmB1;     // conv5|6|7 weights: rank 2 (64, 9 * 96), box (64, 96)
    CUtensorMap tmB2;     // conv8|9 weights:   rank 2 (64, 9 * 64), box (64, 64)
    CUtensorMap tmP2;     // conv2(f): rank 4 (32, W, H, NIMG) fp32, box (32, 128, 1, 1)
    CUtensorMap tmP3;     // conv3(f): same
    CUtensorMap tmL3[3];  // l3 out: half rows (32 x 2 B = one K chunk), boxes of 32, 31 (first slice of a strip) and 29 pixels (last slice), SWIZZLE_64B
    CUtensorMap tmE[3];   // E out: half rows (16 fp32), same boxes
    int W, H, NIMG, n_strips;
    long long total_rowtiles;   // NIMG * n_strips * H
    // bias / BeLU slope of the two activations, in the kernel parameter (constant) bank: network.rs:51-52,55-56
    float l2_bias[32], l2_slope[32], l3_bias[32], l3_slope[32];
    int* fault;
    // test hooks (rsr_debug_fused): fp32 [NIMG][H][W][32] dumps of l2 (as written to the ring) and of the raw
    // accumulator columns 32..63 (conv6(l1) + conv8(l2)); dbg_flags bit 0 = skip the S2 MMAs
    float* dbg_l2;
    float* dbg_t3;
    int dbg_flags;
    long long* prof;            // test hook: per-CTA wait-cycle counters of every warp role [gridDim.x][FPROF_N] (slots: FP_*)
};

// Wait-cycle profile of the fused launch (rsr_debug_fused bit 2 / bit 3, tools/fused_debug.py --profile): every
// role charges the cycles it spends blocked on each mbarrier to its own slot of the CTA's row.  Compiled into the
// DBG instantiation, and into every instantiation (single CTA and CTA pair) when the library is built with
// -DRSR_FUSED_PROF=1 (tools/ablate.sh builds that variant; the shipped library carries none of it).
#ifndef RSR_FUSED_PROF
#define RSR_FUSED_PROF 0
#endif
enum : int {
    FP_A_EMPTY = 0,    // warp 0  l1 row producer: ring slot not yet released by the MMAs
    FP_B_EMPTY,        // warp 3  weight producer: stage not yet released
    FP_P2_EMPTY,       // warp 2  P2 producer: E1 still reading the tile
    FP_P3_EMPTY,       // warp 12 P3 producer: E2 still reading the tile
    FP_T_EMPTY,        // warp 1  MMA: accumulator slot not drained by E2 (or E1 for halo rows)
    FP_A_FULL,         //         l1 row not landed
    FP_B_FULL_S1,      //         weight stage not landed (S1)
    FP_L2_FULL,        //         l2 ring row not written by E1
    FP_B_FULL_S2,      //         weight stage not landed (S2)
    FP_ISSUE_S1,       //         inside the issue blocks of S1 (MMA queue back-pressure included)
    FP_ISSUE_S2,       //         same for S2
    FP_MMA_TOTAL,      //         cycles of the whole role loop
    FP_E1_T1_FULL,     // warp 4  E1: S1 accumulator not complete
    FP_E1_P2_FULL,     //         P2 tile not landed
    FP_E1_L2_EMPTY,    //         l2 ring slot still read by S2
    FP_E1_TOTAL,
    FP_E2_T2_FULL,     // warp 8  E2: S2 accumulation not complete
    FP_E2_P3_FULL,     //         P3 tile not landed
    FP_E2_STORE_READ,  //         staging tile still being read by the previous TMA store
    FP_E2_TOTAL,
    FP_ROWS,           // warp 1  S1 rows issued by this CTA
    FP_SMID,           //         SM the CTA ran on; FP_T_START / FP_T_END: %globaltimer (ns) at the start / end of the MMA role
    FP_T_START,        //         (tools/cta_spread.py: how long every cluster takes and where it sat)
    FP_T_END,
    FPROF_N = 32
};

#ifndef RSR_FUSED_PAIR_TPS
#define RSR_FUSED_PAIR_TPS 3
#endif
#ifndef RSR_FUSED_PAIR_NB
#define RSR_FUSED_PAIR_NB 3
#endif
#ifndef RSR_FUSED_PAIR_NR1
#define RSR_FUSED_PAIR_NR1 3   // the 3-row window only: the next row loads behind the first tap row of S1 (no loss vs 4, r02 A/B); frees 16 KB
#endif
template <bool PAIR_, int PREC_ = PREC_BF16X3>
struct FusedCfgT {
    static constexpr bool PAIR = PAIR_;
    static constexpr int PREC = PREC_;          // operand / activation storage format (conv_strip.cuh)
    static constexpr int STRIP = 124;          // output pixels per strip
    static constexpr int L2W = 126;            // valid l2 pixels per strip row
    static constexpr int SLOT = 128 * 128;     // one row slot: 128 pixels x 128 B
    static constexpr int NR1 = PAIR ? RSR_FUSED_PAIR_NR1 : 3;     // l1 ring (the 3-row window of S1)
    static constexpr int NR2 = 3;              // l2 ring
    static constexpr int TPS = PAIR ? RSR_FUSED_PAIR_TPS : 1;     // taps per weight stage (each tcgen05.commit that
                                               // releases a stage costs tensor-pipe time: fewer, larger stages)
    static constexpr int NSTG = 9 / TPS;       // stages per S1 / S2
    static constexpr int NB = PAIR ? RSR_FUSED_PAIR_NB : 4;       // weight stages in the ring
    static constexpr int H1 = PAIR ? 48 : 96;  // weight rows per conv5|6|7 tap in this CTA
    static constexpr int H2 = PAIR ? 32 : 64;  // weight rows per conv8|9 tap in this CTA
    static constexpr int TAP1 = H1 * 128, TAP2 = H2 * 128;        // bytes of one tap in this CTA
    static constexpr int BST = TPS * TAP1;     // stage pitch
    static_assert(9 % TPS == 0, "taps per stage must divide 9");
    static constexpr int NT = 5;               // TMEM row slots of 96 columns
    static constexpr int TCOLS = 96;
    static constexpr int SF_COL = NT * TCOLS;  // PREC_F16MX: scale-factor columns 480..511 of the 512 allocated
    static constexpr int LOOKAHEAD = 4;        // S1 rows issued before the first S2 of a work item
    static constexpr int NTHREADS = 21 * 32;
    static constexpr int NEW = 8;              // warps per epilogue level: 2 per TMEM lane quarter (16 channels each)
    static constexpr int NARR = PAIR ? 2 * NEW : NEW;  // epilogue-warp arrivals on l2_full / t_empty (both CTAs of a pair)
    static constexpr int OFF_L2 = NR1 * SLOT;
    static constexpr int OFF_B = OFF_L2 + NR2 * SLOT;
    static constexpr int OFF_P2 = OFF_B + NB * BST;
    static constexpr int OFF_P3 = OFF_P2 + SLOT;
    static constexpr int OFF_O = OFF_P3 + SLOT;            // E2 staging: per warp a 2 KB half tile of l3 and one of E
    static constexpr int OFF_BAR = OFF_O + NEW * 4096;
    // a_full/a_empty[NR1] b_full/b_empty[NB] p2_full p2_empty p3_full p3_empty t1_full[NT] l2_full/l2_empty[NR2]
    // t2_full[NT] t_empty[NT]
    static constexpr int NBAR = 2 * NR1 + 2 * NB + 4 + NT + 2 * NR2 + NT + NT;
    static constexpr int SMEM_BYTES = 1024 + OFF_BAR + NBAR * 8 + 16;
    static_assert(OFF_BAR % 1024 == 0 && BST % 1024 == 0, "1024-byte swizzle atoms");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};
using FusedCfg = FusedCfgT<false>;   // (geometry constants that do not depend on PAIR / PREC)

struct FItem {
    int img, x0, y_b, n, y_e;   // rows y_b .. y_b+n-1 are computed, rows >= y_e are not stored
};
// Cluster-level work item (strip, row range) -> this CTA's rows.  Single CTA: the whole range.  Pair: the leader
// takes the upper half, the peer the lower half, both run `n` = ceil(range / 2) rows in lock step (y_e clips the
// peer's stores when the range is odd or a single row).
template <bool PAIR>
__device__ __forceinline__ bool fitem_next(const FusedParams& p, long long& cur, long long end, uint32_t rank, FItem& w) {
    if (cur >= end) return false;
    const int col = (int)(cur / p.H);
    const int y0 = (int)(cur % p.H);
    const int n = (int)min((long long)(p.H - y0), end - cur);
    w.img = col / p.n_strips;
    w.x0 = (col % p.n_strips) * FusedCfg::STRIP;
    if (PAIR) {
        const int l0 = (n + 1) / 2;
        w.y_b = y0 + (rank ? l0 : 0);
        w.n = l0;
        w.y_e = y0 + n;
    } else {
        w.y_b = y0;
        w.n = n;
        w.y_e = y0 + n;
    }
    cur += n;
    return true;
}

#define FW(bar, par, tag) mbar_wait(bar, par, p.fault, tag)
#define FPROF_ON ((DBG || RSR_FUSED_PROF) && p.prof)
// waits charged to a profile slot of this CTA's row when the profiling hook is on (FWPC: cluster-scope acquire)
#define FWP(bar, par, tag, slot)                                   \
    do {                                                            \
        if (FPROF_ON) {                                             \
            const long long t0__ = clock64();                       \
            mbar_wait(bar, par, p.fault, tag);                      \
            prof_acc[slot] += clock64() - t0__;                     \
        } else {                                                    \
            mbar_wait(bar, par, p.fault, tag);                      \
        }                                                           \
    } while (0)
#define FWPC(bar, par, tag, slot)                                  \
    do {                                                            \
        if (FPROF_ON) {                                             \
            const long long t0__ = clock64();                       \
            mbar_wait_cluster(bar, par, p.fault, tag);              \
            prof_acc[slot] += clock64() - t0__;                     \
        } else {                                                    \
            mbar_wait_cluster(bar, par, p.fault, tag);              \
        }                                                           \
    } while (0)
// end of a role: lane 0 adds the role's counters [first, first + n) to the CTA's row
#define FPROF_FLUSH(first, n)                                                              \
    do {                                                                                    \
        if (FPROF_ON && lane == 0) {                                                        \
            long long* dst__ = p.prof + (size_t)blockIdx.x * FPROF_N;                       \
            for (int i__ = (first); i__ < (first) + (n); ++i__) dst__[i__] = prof_acc[i__]; \
        }                                                                                   \
    } while (0)

// DBG: compile the test hooks (dumps, S2 skip, wait profile) in — they cost ~25 % when merely present.
// PAIR: CTA-pair variant; must be launched with cluster dimension 2.
template <bool DBG, bool PAIR, int PREC>
__global__ void __launch_bounds__(FusedCfg::NTHREADS, 1) conv_fused_mid_kernel(const __grid_constant__ FusedParams p) {
    using C = FusedCfgT<PAIR, PREC>;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sL = smem + C::OFF_L2;
    uint8_t* sB = smem + C::OFF_B;
    uint8_t* sP2 = smem + C::OFF_P2;
    uint8_t* sP3 = smem + C::OFF_P3;
    uint8_t* sO = smem + C::OFF_O;
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* a_empty = a_full + C::NR1;
    uint64_t* b_full = a_empty + C::NR1;
    uint64_t* b_empty = b_full + C::NB;
    uint64_t* p2_full = b_empty + C::NB;
    uint64_t* p2_empty = p2_full + 1;
    uint64_t* p3_full = p2_empty + 1;
    uint64_t* p3_empty = p3_full + 1;
    uint64_t* t1_full = p3_empty + 1;
    uint64_t* l2_full = t1_full + C::NT;
    uint64_t* l2_empty = l2_full + C::NR2;
    uint64_t* t2_full = l2_empty + C::NR2;
    uint64_t* t_empty = t2_full + C::NT;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + C::NT);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    long long prof_acc[FPROF_N] = {};   // (dead code unless the profile is compiled in)

    if (threadIdx.x == 0) {
        for (int i = 0; i < C::NR1; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < C::NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        mbar_init(p2_full, 1); mbar_init(p2_empty, C::NEW);
        mbar_init(p3_full, 1); mbar_init(p3_empty, C::NEW);
        for (int i = 0; i < C::NT; ++i) { mbar_init(&t1_full[i], 1); mbar_init(&t2_full[i], 1); mbar_init(&t_empty[i], C::NARR); }
        for (int i = 0; i < C::NR2; ++i) { mbar_init(&l2_full[i], C::NARR); mbar_init(&l2_empty[i], 1); }
        fence_mbar_init();
        tma_prefetch_desc(&p.tmA);
        tma_prefetch_desc(&p.tmB1);
        tma_prefetch_desc(&p.tmB2);
        tma_prefetch_desc(&p.tmP2);
        tma_prefetch_desc(&p.tmP3);
    }
    if (warp == 2) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            tmem_alloc(tmem_slot, 512);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) cluster_sync_all();   // both CTAs' cta_group::2 allocations (each may write both tmem_slots) have completed
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    init_scale_factors<C>(tmem_base, warp);
    if (PAIR) {   // both CTAs' barriers exist (and scale factors are written) before any cross-CTA signal
        tc_fence_before();
        cluster_sync_all();
        tc_fence_after();
    }
    griddep_launch_dependents();   // programmatic dependent launch (see conv_strip.cuh)
    griddep_wait();

    // work ranges are per cluster (pair) or per CTA
    const long long nrange = PAIR ? gridDim.x >> 1 : gridDim.x, irange = PAIR ? blockIdx.x >> 1 : blockIdx.x;
    long long cur = p.total_rowtiles * irange / nrange;
    const long long end = p.total_rowtiles * (irange + 1) / nrange;
    FItem w;
    // barriers the MMA warp of the LEADER waits on, as seen from this CTA (pair: remote for the peer)
    const uint32_t a_full0 = PAIR ? map_to_cta(smem_u32(a_full), 0) : smem_u32(a_full);
    const uint32_t b_full0 = PAIR ? map_to_cta(smem_u32(b_full), 0) : smem_u32(b_full);
    const uint32_t l2_full0 = PAIR ? map_to_cta(smem_u32(l2_full), 0) : smem_u32(l2_full);
    const uint32_t t_empty0 = PAIR ? map_to_cta(smem_u32(t_empty), 0) : smem_u32(t_empty);

    // Counters shared by convention between the roles (each role keeps its own copy, advanced identically):
    //   c1 = number of S1 rows so far (all items): T slot c1 % NT, l2 slot c1 % NR2
    //   a  = number of l1 rows loaded so far:      l1 slot a % NR1
    // The "full" barrier of use number u of a slot completes phase u, so a consumer waits parity u & 1 and
    // a producer waits parity (u & 1) ^ 1 of the "empty" barrier (returns at once on a fresh barrier).
    if (warp == 0) {
        // ===== l1 row producer: rows y_b-2 .. y_e+1 of the strip, pixels x0-2 .. x0+125 =====
        uint32_t a = 0;
        while (fitem_next<PAIR>(p, cur, end, rank, w)) {
            for (int i = 0; i < w.n + 4; ++i, ++a) {
                const uint32_t sl = a % C::NR1, par = ((a / C::NR1) & 1) ^ 1;
                FWP(&a_empty[sl], par, 21, FP_A_EMPTY);
                if (elect_one()) {
                    if (PAIR) {
                        if (leader) mbar_arrive_expect_tx(&a_full[sl], 2 * C::SLOT);
                        tma2_load_4d(smem_u32(sA + sl * C::SLOT), &p.tmA, a_full0 + sl * 8, 0, w.x0 - 2, w.y_b - 2 + i, w.img);
                    } else {
                        mbar_arrive_expect_tx(&a_full[sl], C::SLOT);
                        tma_load_4d(sA + sl * C::SLOT, &p.tmA, &a_full[sl], 0, w.x0 - 2, w.y_b - 2 + i, w.img);
                    }
                }
                __syncwarp();
            }
        }
        FPROF_FLUSH(FP_A_EMPTY, 1);
    } else if (warp == 3) {
        // ===== weight producer: one tap per stage, in the MMA warp's S1 / S2 order =====
        uint32_t st = 0, par = 1;
        // ncol = output columns of the op (96 / 64); a pair CTA loads its half of the tap's rows
        auto taps = [&](const CUtensorMap* tm, int ncol) {
            const int tapb = (PAIR ? ncol / 2 : ncol) * 128;   // bytes of one tap in this CTA
            for (int sp = 0; sp < C::NSTG; ++sp) {
                FWP(&b_empty[st], par, 22, FP_B_EMPTY);
                if (elect_one()) {
                    if (PAIR) {
                        if (leader) mbar_arrive_expect_tx(&b_full[st], C::TPS * ncol * 128);   // both halves of every tap
                        for (int tt = 0; tt < C::TPS; ++tt)
                            tma2_load_2d(smem_u32(sB + st * C::BST + tt * tapb), tm, b_full0 + st * 8, 0,
                                         (sp * C::TPS + tt) * ncol + (int)rank * (ncol / 2));
                    } else {
                        mbar_arrive_expect_tx(&b_full[st], C::TPS * ncol * 128);
                        for (int tt = 0; tt < C::TPS; ++tt)
                            tma_load_2d(sB + st * C::BST + tt * tapb, tm, &b_full[st], 0, (sp * C::TPS + tt) * ncol);
                    }
                }
                __syncwarp();
                if (++st == C::NB) { st = 0; par ^= 1; }
            }
        };
        while (fitem_next<PAIR>(p, cur, end, rank, w)) {
            // op order of the MMA warp: S1 runs LOOKAHEAD rows ahead of S2
            const int ns1 = w.n + 2;
            for (int i1 = 0, r = 0; r < w.n;) {
                const bool is_s1 = i1 < ns1 && i1 < r + C::LOOKAHEAD;
                taps(is_s1 ? &p.tmB1 : &p.tmB2, is_s1 ? 96 : 64);
                if (is_s1) ++i1; else ++r;
            }
        }
        FPROF_FLUSH(FP_B_EMPTY, 1);
    } else if (warp == 2) {
        // ===== P2 tile producer: one tile per S1 row (pixels x0-1 .. x0+126 of row y_b-1+q) =====
        uint32_t u = 0;
        while (fitem_next<PAIR>(p, cur, end, rank, w)) {
            for (int q = 0; q < w.n + 2; ++q, ++u) {
                FWP(p2_empty, (u & 1) ^ 1, 23, FP_P2_EMPTY);
                if (elect_one()) {
                    mbar_arrive_expect_tx(p2_full, C::SLOT);
                    tma_load_4d(sP2, &p.tmP2, p2_full, 0, w.x0 - 1, w.y_b - 1 + q, w.img);
                }
                __syncwarp();
            }
        }
        FPROF_FLUSH(FP_P2_EMPTY, 1);
    } else if (warp == 20) {
        // ===== P3 tile producer: one tile per S2 row (pixels x0-1 .. x0+126 of row y_b+r: tile row = TMEM lane) =====
        uint32_t u = 0;
        while (fitem_next<PAIR>(p, cur, end, rank, w)) {
            for (int r = 0; r < w.n; ++r, ++u) {
                FWP(p3_empty, (u & 1) ^ 1, 24, FP_P3_EMPTY);
                if (elect_one()) {
                    mbar_arrive_expect_tx(p3_full, C::SLOT);
                    tma_load_4d(sP3, &p.tmP3, p3_full, 0, w.x0 - 1, w.y_b + r, w.img);
                }
                __syncwarp();
            }
        }
        FPROF_FLUSH(FP_P3_EMPTY, 1);
    } else if (warp == 1) {
        // ===== MMA issuer (pair: the leader's, for both CTAs) =====
        if (leader) {
        using MM = SplitMma<PREC, PAIR ? 2 : 1, 128>;
        constexpr int MM_M = PAIR ? 256 : 128;
        const uint32_t sf = tmem_base + C::SF_COL;
        auto commit = [](uint64_t* bar) { if (PAIR) umma2_commit(bar); else umma_commit(bar); };
        // (barriers that the epilogue warps of both CTAs arrive on need a cluster-scope acquire in the pair kernel: FWPC)
        constexpr uint32_t LBO = 1u << 16;
        const uint32_t sA_e = (smem_u32(sA) >> 4) | LBO, sL_e = (smem_u32(sL) >> 4) | LBO, sB_e = (smem_u32(sB) >> 4) | LBO;
        uint32_t c1_item = 0, a_item = 0;   // counters at the start of the current item
        uint32_t a_wait = 0, l2_wait = 0;   // rows already waited for
        uint32_t c2 = 0;                    // S2 rows so far
        uint32_t bst = 0, bpar = 0;
        int n_item = 0;
        const long long prof_t0 = FPROF_ON ? clock64() : 0;
        long long prof_g0 = 0;
        if (FPROF_ON) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(prof_g0));

        // S1(q): l1 rows q..q+2 of the item -> T slot of row q
        auto s1 = [&](int q) {
            const uint32_t c = c1_item + q;
            const uint32_t ts = c % C::NT;
            if (PAIR) FWPC(&t_empty[ts], ((c / C::NT) & 1) ^ 1, 25, FP_T_EMPTY);
            else FWP(&t_empty[ts], ((c / C::NT) & 1) ^ 1, 25, FP_T_EMPTY);
            while (a_wait <= a_item + q + 2) {
                FWP(&a_full[a_wait % C::NR1], (a_wait / C::NR1) & 1, 26, FP_A_FULL);
                ++a_wait;
            }
            if (FPROF_ON) ++prof_acc[FP_ROWS];
            tc_fence_after();
            uint32_t arow[3];
            #pragma unroll
            for (int ta = 0; ta < 3; ++ta) arow[ta] = sA_e + ((a_item + q + ta) % C::NR1) * (C::SLOT >> 4);
            const uint32_t d = tmem_base + ts * C::TCOLS;
            #pragma unroll
            for (int sp = 0; sp < C::NSTG; ++sp) {
                FWP(&b_full[bst], bpar, 27, FP_B_FULL_S1);
                tc_fence_after();
                const long long ti0 = FPROF_ON ? clock64() : 0;
                if (elect_one()) {
                    #pragma unroll
                    for (int tt = 0; tt < C::TPS; ++tt) {
                        const int t = sp * C::TPS + tt;
                        const int ta = t / 3, tb = t % 3;
                        const uint32_t b_e = sB_e + bst * (C::BST >> 4) + tt * (C::TAP1 >> 4);
                        const uint32_t a_e = arow[ta] + tb * (128 >> 4);
                        #pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            if (ABL_NO_MMA) continue;
                            MM::chunk(d, a_e, b_e, k, MM_M, 96, sf, (t | k) != 0 ? 1u : 0u);
                        }
                        if (t == 2) commit(&a_empty[(a_item + q) % C::NR1]);   // top window row is dead
                    }
                    commit(&b_empty[bst]);
                }
                __syncwarp();
                if (FPROF_ON) prof_acc[FP_ISSUE_S1] += clock64() - ti0;
                if (++bst == C::NB) { bst = 0; bpar ^= 1; }
            }
            if (elect_one()) {
                commit(&t1_full[ts]);
                if (q == n_item + 1) {   // last S1 of the item: the rest of its window is dead too
                    commit(&a_empty[(a_item + q + 1) % C::NR1]);
                    commit(&a_empty[(a_item + q + 2) % C::NR1]);
                }
            }
            __syncwarp();
        };
        // S2(r): l2 rows r..r+2 of the item, accumulated onto columns 32..95 of the T slot of l2 row r+1; lane m is
        // pixel x0-1+m as in S1 (lane 0 reads one pixel before the ring slot: garbage in, discarded)
        auto s2 = [&](int r) {
            const uint32_t ts = (c1_item + r + 1) % C::NT;
            while (l2_wait <= c1_item + r + 2) {
                if (PAIR) FWPC(&l2_full[l2_wait % C::NR2], (l2_wait / C::NR2) & 1, 28, FP_L2_FULL);
                else FWP(&l2_full[l2_wait % C::NR2], (l2_wait / C::NR2) & 1, 28, FP_L2_FULL);
                ++l2_wait;
            }
            tc_fence_after();
            uint32_t lrow[3];
            #pragma unroll
            for (int ta = 0; ta < 3; ++ta) lrow[ta] = sL_e + ((c1_item + r + ta) % C::NR2) * (C::SLOT >> 4);
            const uint32_t d = tmem_base + ts * C::TCOLS + 32;
            #pragma unroll
            for (int sp = 0; sp < C::NSTG; ++sp) {
                FWP(&b_full[bst], bpar, 29, FP_B_FULL_S2);
                tc_fence_after();
                const long long ti0 = FPROF_ON ? clock64() : 0;
                if (elect_one()) {
                    #pragma unroll
                    for (int tt = 0; tt < C::TPS; ++tt) {
                        const int t = sp * C::TPS + tt;
                        const int ta = t / 3, tb = t % 3;
                        const uint32_t b_e = sB_e + bst * (C::BST >> 4) + tt * (C::TAP2 >> 4);
                        const uint32_t a_e = lrow[ta] + (tb - 1) * (128 >> 4);   // lane m reads ring pixels m-1..m+1
                        #pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            if ((DBG && (p.dbg_flags & 1)) || ABL_NO_MMA) continue;
                            MM::chunk(d, a_e, b_e, k, MM_M, 64, sf, 1);
                        }
                        if (t == 2) commit(&l2_empty[(c1_item + r) % C::NR2]);
                    }
                    commit(&b_empty[bst]);
                }
                __syncwarp();
                if (FPROF_ON) prof_acc[FP_ISSUE_S2] += clock64() - ti0;
                if (++bst == C::NB) { bst = 0; bpar ^= 1; }
            }
            if (elect_one()) {
                commit(&t2_full[c2 % C::NT]);
                if (r == n_item - 1) {
                    commit(&l2_empty[(c1_item + r + 1) % C::NR2]);
                    commit(&l2_empty[(c1_item + r + 2) % C::NR2]);
                }
            }
            __syncwarp();
            ++c2;
        };
        while (fitem_next<PAIR>(p, cur, end, rank, w)) {
            n_item = w.n;
            // S1 runs LOOKAHEAD rows ahead of S2: S1(0..3) S2(0) S1(4) S2(1) ... ; all n+2 S1 rows precede S2(n-1)
            const int ns1 = w.n + 2;
            for (int i1 = 0, r = 0; r < w.n;) {
                if (i1 < ns1 && i1 < r + C::LOOKAHEAD) s1(i1++);
                else s2(r++);
            }
            c1_item += ns1;
            a_item += w.n + 4;
        }
        if (FPROF_ON) prof_acc[FP_MMA_TOTAL] = clock64() - prof_t0;
        FPROF_FLUSH(FP_T_EMPTY, FP_MMA_TOTAL - FP_T_EMPTY + 1);
        if (FPROF_ON) {
            uint32_t smid;
            long long t1;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            prof_acc[FP_SMID] = smid; prof_acc[FP_T_START] = prof_g0; prof_acc[FP_T_END] = t1;
        }
        FPROF_FLUSH(FP_ROWS, 4);
        }  // leader
    } else if (warp >= 4 && warp < 12) {
        // ===== E1: l2pre (+ P2) -> l2 row in the shared-memory ring; this warp: lane quarter ew, channels 16*half .. +15 =====
        const int ew = warp & 3, half = (warp - 4) >> 2;
        const int m = ew * 32 + lane;            // MMA row = l2 ring pixel index
        const uint32_t sL_u = smem_u32(sL), sP2_u = smem_u32(sP2);
        const float* bias = p.l2_bias + 16 * half;
        const float* slope = p.l2_slope + 16 * half;
        uint32_t c = 0;
        const long long prof_t0 = FPROF_ON ? clock64() : 0;
        while (fitem_next<PAIR>(p, cur, end, rank, w)) {
            for (int q = 0; q < w.n + 2; ++q, ++c) {
                const uint32_t ts = c % C::NT, ls = c % C::NR2;
                const int y = w.y_b - 1 + q, px = w.x0 - 1 + m;
                FWP(&t1_full[ts], (c / C::NT) & 1, 30, FP_E1_T1_FULL);
                tc_fence_after();
                uint32_t raw[16];
                tmem_ld_32x16(tmem_base + ((uint32_t)(ew * 32) << 16) + ts * C::TCOLS + 16 * half, raw);
                tmem_ld_wait();
                float v[16];
                #pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(raw[k]);
                FWP(p2_full, c & 1, 31, FP_E1_P2_FULL);
                epi_add_partial_half(v, sP2_u, m, half);
                __syncwarp();
                if (lane == 0) mbar_arrive(p2_empty);
                if (!ABL_NO_E1_MATH) {
                    #pragma unroll
                    for (int i = 0; i < 16; i += 2) bias_belu2(v[i], v[i + 1], bias[i], bias[i + 1], slope[i], slope[i + 1]);
                }
                // Padding::Same of conv8 / conv9: l2 is ZERO outside the image (not BeLU(bias)); lanes past the
                // strip's 126 valid l2 pixels hold garbage accumulators and are zeroed too
                const bool valid = y >= 0 && y < p.H && px >= 0 && px < p.W && m < C::L2W;
                if (DBG && p.dbg_l2 && valid) {
                    float4* dst = reinterpret_cast<float4*>(p.dbg_l2 + (((size_t)w.img * p.H + y) * p.W + px) * 32 + 16 * half);
                    #pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
                uint32_t words[16];
                if (ABL_NO_E1_MATH) {
                    #pragma unroll
                    for (int i = 0; i < 16; ++i) words[i] = __float_as_uint(v[i]);
                } else pack_split_half<PREC>(v, words, !valid);
                // the MMAs that read the previous occupant of this ring slot must have completed
                FWP(&l2_empty[ls], ((c / C::NR2) & 1) ^ 1, 32, FP_E1_L2_EMPTY);
                store_half_row_swizzled(words, sL_u + ls * C::SLOT + m * 128, m & 7, half);
                fence_proxy_async_smem();   // generic-proxy writes -> visible to the tensor core's async-proxy reads
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    // (pair: both CTAs' warps arrive on the leader's barriers)
                    if (PAIR) mbar_arrive_cluster(l2_full0 + ls * 8); else mbar_arrive(&l2_full[ls]);
                    // halo rows have no S2 / E2: their T slot is released here
                    if (q == 0 || q == w.n + 1) {
                        if (PAIR) mbar_arrive_cluster(t_empty0 + ts * 8); else mbar_arrive(&t_empty[ts]);
                    }
                }
            }
        }
        if (FPROF_ON) prof_acc[FP_E1_TOTAL] = clock64() - prof_t0;
        if (warp == 4) FPROF_FLUSH(FP_E1_T1_FULL, 4);
    } else if (warp >= 12 && warp < 20) {
        // ===== E2: T[32:64] (+ P3) -> l3;  T[64:96] -> E; this warp: lane quarter ew, channels 16*half .. +15 of both =====
        // Lane m of the strip is pixel x0-1+m; the strip's outputs are lanes 1..124, so quarter 0 drops its lane 0 (staging
        // row = lane-1, 31-pixel store at x0) and quarter 3 its lanes 29..31 (29-pixel store).
        const int ew = warp & 3, half = (warp - 12) >> 2;
        const int m = ew * 32 + lane;
        const uint32_t stage_l3 = smem_u32(sO) + (warp - 12) * 4096, stage_e = stage_l3 + 2048;
        const uint32_t sP3_u = smem_u32(sP3);
        const int bx = ew == 0 ? 1 : ew == 3 ? 2 : 0;     // tensor maps with 32 / 31 / 29 pixel boxes
        const int srow = ew == 0 ? lane - 1 : lane;       // staging row of this lane (-1: not stored)
        const float* bias = p.l3_bias + 16 * half;
        const float* slope = p.l3_slope + 16 * half;
        uint32_t c1_item = 0, c2 = 0;
        const long long prof_t0 = FPROF_ON ? clock64() : 0;
        // one TMA store of this warp's half tile; the previous store from the same tile must have finished READING it: at
        // most the one newer store (of the other tile) may still be pending
        auto store_tile = [&](const uint32_t (&words)[16], uint32_t stage, const CUtensorMap* tm, int c0, int xs, int y, int img) {
            {
                const long long t0s = FPROF_ON ? clock64() : 0;
                if (elect_one()) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                if (FPROF_ON) prof_acc[FP_E2_STORE_READ] += clock64() - t0s;
            }
            if (srow >= 0) store_row64_swizzled(words, stage, srow);
            fence_proxy_async_smem();
            __syncwarp();
            if (elect_one()) {
                asm volatile(
                    "cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                    ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(stage), "r"(c0), "r"(xs), "r"(y), "r"(img)
                    : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        };
        while (fitem_next<PAIR>(p, cur, end, rank, w)) {
            const int xl = w.x0 - 1 + ew * 32;
            const int xs = ew == 0 ? w.x0 : xl;            // first pixel this warp stores
            const bool live_x = xs < p.W;   // warp-uniform
            for (int r = 0; r < w.n; ++r, ++c2) {
                const uint32_t ts = (c1_item + r + 1) % C::NT;
                const int y = w.y_b + r;
                const bool live = live_x && y < w.y_e;   // (pair: the peer's padding rows are computed, not stored)
                FWP(&t2_full[c2 % C::NT], (c2 / C::NT) & 1, 33, FP_E2_T2_FULL);
                tc_fence_after();
                const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16) + ts * C::TCOLS + 16 * half;
                uint32_t raw[16], raw_e[16];
                tmem_ld_32x16(trow + 32, raw);
                tmem_ld_32x16(trow + 64, raw_e);
                tmem_ld_wait();
                float v[16];
                #pragma unroll
                for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(raw[k]);
                // --- E: raw fp32 partial of `expand` ---
                if (!ABL_NO_EPI_OUT && !ABL_NO_E2_STORE && live) store_tile(raw_e, stage_e, &p.tmE[bx], 16 * half, xs, y, w.img);
                if (DBG && p.dbg_t3 && live && m >= 1 && m <= C::STRIP && xl + lane < p.W) {
                    float4* dst = reinterpret_cast<float4*>(p.dbg_t3 + (((size_t)w.img * p.H + y) * p.W + xl + lane) * 32 + 16 * half);
                    #pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
                // --- l3 ---
                FWP(p3_full, c2 & 1, 34, FP_E2_P3_FULL);
                epi_add_partial_half(v, sP3_u, m, half);
                __syncwarp();
                if (lane == 0) mbar_arrive(p3_empty);
                if (ABL_NO_EPI_OUT) { if (v[0] == 123.456f) p.dbg_l2[0] = v[1] + __uint_as_float(raw_e[0]); }
                else if (live) {
                    #pragma unroll
                    for (int i = 0; i < 16; i += 2) bias_belu2(v[i], v[i + 1], bias[i], bias[i + 1], slope[i], slope[i + 1]);
                    uint32_t words[16];
                    pack_split_half<PREC>(v, words);
                    if (ABL_NO_E2_STORE) { if (words[0] == 0x12345678u) p.dbg_l2[0] = __uint_as_float(words[1] ^ words[5] ^ words[9] ^ words[13]); }
                    else store_tile(words, stage_l3, &p.tmL3[bx], 32 * half, xs, y, w.img);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_cluster(t_empty0 + ts * 8); else mbar_arrive(&t_empty[ts]);
                }
            }
            c1_item += w.n + 2;
        }
        if (FPROF_ON) prof_acc[FP_E2_TOTAL] = clock64() - prof_t0;
        if (warp == 12) FPROF_FLUSH(FP_E2_T2_FULL, 4);
        if (elect_one()) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (PAIR) {
        cluster_sync_all();   // the pair's MMAs and remote arrivals are over before either CTA leaves
        if (warp == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    } else if (warp == 2) {
        tmem_dealloc(tmem_base, 512);
    }
}

#undef FW
#undef FWP
#undef FWPC
#undef FPROF_FLUSH

}  // namespace rsr
