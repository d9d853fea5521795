// fused_async.cu -- the fast HOGWILD form of the legacy fused pass (stale train_sg: per window pair the o3 community
// gradient of x_j, the SGNS pair update and the combined write; utils/training_sdg_inner.c:1597-1905, 2520-2715,
// 2988-3740) at size 128 with pi in top-1 form: ONE persistent cooperative kernel, warp-specialised, no grid barrier.
//
// The o3 term is a 128x128 mat-vec per pair against the inverse covariance of x_j's community; a window holds ~4.5
// communities, so a warp alone can batch only ~4 rows per community (round 1 streamed 64 KB of inv_cov from L2 per such
// group).  Here the batching is done ACROSS the ~5900 walks in flight, asynchronously:
//
//   one CTA per SM; 20 WALKER warps + 8 SERVICE warps (a front and a back team of 4) per CTA; setmaxnreg gives the walkers
//   80 registers and the service warps 48.
//   walker warp (interleaves two walks; per walk one centre at a time, exactly the o2 kernel's order):
//     [stage]  lists the rows of its next window; every row whose o3 term can be taken from the row's value at the start
//              of the centre (all but a node that repeats inside the window) becomes a request {row, result slot} appended
//              to the queue of (its community, replica) -- a ring in global memory, tail reserved with one warp-aggregated
//              atomicAdd per community, entries published with single 8-byte stores;
//     [wait]   until its completion counter shows that all requests of the centre were served;
//     [sgns]   the centre's pairs (fused_sgns.cuh): SGNS on the o2 size-128 code path + the combined write with the o3
//              term from its result slots; repeated nodes get the term in-warp from the current value, so a walk sees the
//              reference's sequential semantics exactly.
//   service, front team: looks at all queues, takes the fullest (the resident community counts twice), claims up to 64
//   entries with one CAS (any CTA serves any community: the load balances itself), keeps inv_cov_c resident in shared
//   memory as the tcgen05 A operand (hi/lo TF32 images, fetched by the TMA engine with cp.async.bulk only on a community
//   switch), gathers the rows, forms x - mu_c and splits hi/lo into the swizzled B operand;
//   service, back team: one thread issues 48 tcgen05.mma (3xTF32: fp32-level accuracy) into one of two TMEM accumulators,
//   the team waits for them, reads the accumulator back with tcgen05.ld, writes Y to the result slots (L2 evict_first),
//   release-increments the requesters' counters and hands the buffer back -- while the front team is already claiming
//   and gathering the next tile.
//
// Nothing waits on a barrier that another CTA must reach: walkers wait only for results, service warps only for
// published entries, so the SGNS half runs at the o2 kernel's pace while the tensor half hides behind it.
#include <cstdio>
#include <cstdlib>

#include "comemb_common.cuh"
#include "fused_sgns.cuh"
#include "umma.cuh"

#ifndef COMEMB_ASYNC_STATS
#define COMEMB_ASYNC_STATS 0  // 1: per-phase cycle counters (printed with COMEMB_ROUND_STATS=1); they cost the 48-register
#endif                        //    service warps local-memory spills, so the product build leaves them out

namespace {
constexpr bool STATS = COMEMB_ASYNC_STATS != 0;
__device__ __forceinline__ long long tick() { return STATS ? clock64() : 0LL; }

using fused::INFO_INWARP;
constexpr int D = 128;
constexpr int TN = 64;                      // requests per GEMM tile (tcgen05 N)
#ifndef COMEMB_VMAX
#define COMEMB_VMAX 64
#endif
#ifndef COMEMB_WCTX
#define COMEMB_WCTX 2
#endif
constexpr int VMAX = COMEMB_VMAX;           // window rows per centre (2*window <= VMAX)
constexpr int A_IMG_BYTES = 2 * D * D * 4;  // hi + lo operand images of one community
constexpr int NSVC = 4;                     // warps per service team (one per TMEM lane quarter): front = warps 0..3, back = 4..7
constexpr int ASYNC_WARPS = 28;             // warps per CTA: 8 service + 20 walkers.  896 threads leave 72 registers per thread at
                                            // launch; the service warpgroups give registers back (setmaxnreg.dec -> 48) and the
                                            // walker warpgroups take them (setmaxnreg.inc -> 80), like the 24-warp build
constexpr int SVC_REGS = 48, WALK_REGS = 80;
// Registers are allocated per SM sub-partition (16384 each; warp w lives on sub-partition w % 4): with the 8 service warps
// first, a sub-partition holds 2 service warps and ceil((ASYNC_WARPS - 8) / 4) walkers.  A split that does not fit there --
// or fits exactly (56 / 80 at 28 warps, 48 / 88 at 26) -- leaves setmaxnreg.inc waiting forever.
static_assert((2 * SVC_REGS + ((ASYNC_WARPS - 8 + 3) / 4) * WALK_REGS) * 32 <= 16384 - 512, "register split per sub-partition");
constexpr int MAXQ = 64;                    // ints of front-team scratch
constexpr int WCTX = COMEMB_WCTX;           // walks interleaved per walker warp
constexpr unsigned long long EMPTY = ~0ULL;
constexpr long long WAIT_TIMEOUT = 8000000000LL;  // ~4 s of SM clocks: a protocol bug must end the kernel, not hang the GPU

struct AsyncParams {
    float *node, *ctx;
    const uint32_t *walks;
    const int64_t *walk_off;
    int64_t n_walks;
    const int32_t *rw;
    const uint64_t *seeds;
    uint64_t base_seed;
    const uint32_t *table;
    TableMod mod;
    const float *mu, *inv_cov;
    const char *a_img;
    const int32_t *comm;
    const float *weight;
    int K, window, is_node;
    float lr, lambda1, lambda2;
    const float *glut;
    // scratch
    float *ybuf;                // [total warps * vslots][128]
    unsigned long long *ring;   // [Q][cap] request entries: row | slot << 32, EMPTY when free
    unsigned *tail;             // [Q * 8]: reserved entries per queue (one 32-byte sector each)
    unsigned *claim;            // [Q * 8]: entries claimed by service CTAs (claim <= tail)
    unsigned *done;             // [total warps]: served requests per walker warp (monotonic)
    int *live;                  // walker warps that still have work
    unsigned long long *walk_cursor;
    int *err;
    long long *stats;
    int y_hint;                 // 1: L2 evict_first on the result-slot stores and loads
    int stick;                  // a CTA leaves its resident community only for a queue with more than `stick` times its backlog
    int64_t cap;                // ring capacity (power of two >= all requests that can be outstanding)
    int n_rep, Q, vslots;
    int64_t active_warps;
};

struct WalkCtx {  // one walk in progress (shared memory; lane 0 writes, the warp reads)
    const uint32_t *path;
    const int32_t *rwp;
    uint64_t rnd;       // LCG state the next pair's samples are drawn from
    int len, ci, V;     // walk length, current centre, rows of its window
    uint32_t wi;        // centre token
    unsigned expected;  // requests posted so far (the completion counter must reach it)
    int have;           // a staged centre is waiting for its sgns phase
};
static_assert(sizeof(WalkCtx) == 48, "layout");

template <int NW>
struct AsyncSmem {
    static constexpr int A_HI = 0, A_LO = D * D * 4, B_HI = 2 * D * D * 4, B_LO = B_HI + TN * D * 4;
    static constexpr int MU = B_LO + TN * D * 4;   // float[128]
    static constexpr int ROW = MU + D * 4;         // uint32[2][TN] rows of the tiles in flight
    static constexpr int SLOT = ROW + 2 * TN * 4;  // uint32[2][TN] their result slots
    static constexpr int LUT = SLOT + 2 * TN * 4;  // float[1000]
    static constexpr int HEAD = LUT + 4096;        // uint32[MAXQ] consumed entries per owned queue
    static constexpr int NWALK = NW - 8;           // walker warps (8 service warps)
    static constexpr int WTOK = HEAD + MAXQ * 4;   // per walker warp and walk context: uint32 tok[VMAX], int32 info[VMAX]
    static constexpr int WINF = WTOK + NWALK * WCTX * VMAX * 4;
    static constexpr int WCTXS = WINF + NWALK * WCTX * VMAX * 4;  // per walker warp: WalkCtx[WCTX]
    static constexpr int BAR = (WCTXS + NWALK * WCTX * 48 + 15) & ~15;
    static constexpr int TOTAL = BAR + 128;
};

template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void front_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NSVC * 32) : "memory"); }
__device__ __forceinline__ void back_barrier() { asm volatile("bar.sync 2, %0;" ::"n"(NSVC * 32) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ unsigned ld_vol(const unsigned *p) { return *reinterpret_cast<const volatile unsigned *>(p); }
// release / acquire at gpu scope: an entry (or a counter increment) published with release makes every write that
// happened before it -- the publisher's own and, cumulatively, those it observed through a CTA/warp barrier -- visible to
// the thread that acquires it
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Rows r = first, first + STRIDE, ... (COUNT of them, those < n) of the tile: gather from the node table, subtract mu_c,
// split hi/lo TF32 and store into the swizzled B operand images, GB rows in flight per warp.  The tile's 64 rows are dealt
// out in 16 classes r mod 16; front warp w takes classes w, w + 4, w + 8, w + 12.
template <int GB, int STRIDE, int COUNT>
__device__ __forceinline__ void gather_rows(const AsyncParams &P, char *b_hi_img, char *b_lo_img, const uint32_t *row_b,
                                            const float *mu_s, int first, int n, int lane) {
    static_assert(COUNT % GB == 0, "batches");
    const float4 m = *reinterpret_cast<const float4 *>(mu_s + 4 * lane);
#pragma unroll
    for (int i0 = 0; i0 < COUNT; i0 += GB) {
        float4 xv[GB];
#pragma unroll
        for (int qq = 0; qq < GB; qq++) {
            const int r = first + STRIDE * (i0 + qq);
            if (r < n) xv[qq] = __ldcg(reinterpret_cast<const float4 *>(P.node + (int64_t)row_b[r] * D + 4 * lane));
        }
#pragma unroll
        for (int qq = 0; qq < GB; qq++) {
            const int r = first + STRIDE * (i0 + qq);
            if (r < n) {
                const float4 x = xv[qq];
                const float4 df = make_float4(x.x - m.x, x.y - m.y, x.z - m.z, x.w - m.w);
                const float4 hi = make_float4(umma::tf32_round(df.x), umma::tf32_round(df.y), umma::tf32_round(df.z),
                                              umma::tf32_round(df.w));
                const float4 lo = make_float4(umma::tf32_round(df.x - hi.x), umma::tf32_round(df.y - hi.y),
                                              umma::tf32_round(df.z - hi.z), umma::tf32_round(df.w - hi.w));
                const uint32_t off = umma::sw128_offset(TN, r, 4 * lane);
                *reinterpret_cast<float4 *>(b_hi_img + off) = hi;
                *reinterpret_cast<float4 *>(b_lo_img + off) = lo;
            }
        }
    }
}

template <bool ATOMIC, int NEG, int NW>
__global__ void __launch_bounds__(NW * 32, 1) sg_async_kernel(const AsyncParams P) {
    using L = AsyncSmem<NW>;
    constexpr LcgJump<NEG> J{};
    constexpr uint32_t TMEM_COLS = 2 * TN;  // two accumulators: the back team drains one while the MMAs fill the other
    extern __shared__ char smem_raw[];
    char *smem = reinterpret_cast<char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float *lut = reinterpret_cast<float *>(smem + L::LUT);
    uint64_t *bar_a = reinterpret_cast<uint64_t *>(smem + L::BAR), *bar_mma = bar_a + 1, *bar_free = bar_a + 3;
    uint64_t *bar_full = bar_a + 5, *bar_ent = bar_a + 7;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar_a + 9);
    int *sel = reinterpret_cast<int *>(bar_a + 10);  // front scratch: 4 ints
    int *n_s = reinterpret_cast<int *>(bar_a + 12);  // [0..1] requests of the tile in buffer 0 / 1 (-1: no more tiles),
                                                     // [2..3] 1 if that tile waits for an operand fetch
    uint32_t *row_s = reinterpret_cast<uint32_t *>(smem + L::ROW);
    uint32_t *slot_s = reinterpret_cast<uint32_t *>(smem + L::SLOT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int e = threadIdx.x; e < EXP_TABLE_SIZE; e += blockDim.x) lut[e] = P.glut[e];
    if (warp == 0) umma::tmem_alloc(tmem_slot, TMEM_COLS);
    if (threadIdx.x == 0) {
        umma::mbar_init(bar_a, 1);
        umma::mbar_init(bar_mma, 1);
        umma::mbar_init(bar_mma + 1, 1);
        umma::mbar_init(bar_free, NSVC);
        umma::mbar_init(bar_free + 1, NSVC);
        umma::mbar_init(bar_full, 1);
        umma::mbar_init(bar_full + 1, 1);
        umma::mbar_init(bar_ent, 1);
        umma::mbar_init(bar_ent + 1, 1);
        umma::fence_mbar_init();
    }
    umma::tc_fence_before();
    __syncthreads();
    umma::tc_fence_after();
    const uint32_t taddr = *tmem_slot;
    const int K = P.K;

    if (warp < NSVC) {
        // =============================== SERVICE, FRONT TEAM (warps 0..3) ================================================
        // claim a tile -> pop its entries -> gather rows -> B operand -> hand it to the back team
        if (NW * 32 > 768) reg_dec<SVC_REGS>();
        float *mu_s = reinterpret_cast<float *>(smem + L::MU);
        const uint32_t a_hi = umma::smem_u32(smem + L::A_HI), a_lo = umma::smem_u32(smem + L::A_LO);
        const uint32_t b_hi = umma::smem_u32(smem + L::B_HI), b_lo = umma::smem_u32(smem + L::B_LO);
        const int tid = threadIdx.x;  // 0..127
        int *pick = sel;              // pick[0..2] = {queue, first ring index, count} of the claimed tile
        int *wbest = reinterpret_cast<int *>(smem + L::HEAD);  // per front warp: best key of the scan
        int cur_c = -1, empty_scans = 0;
        bool a_pending = false;
        long long tiles = 0, rows_served = 0, idle_polls = 0;
        long long ph[6] = {0, 0, 0, 0, 0, 0}, tp = tick();
        auto lap = [&](int k) {
            if (STATS) {
                const long long now = clock64();
                ph[k] += now - tp;
                tp = now;
            }
        };
        // Pick the next tile: every front thread looks at the queues t, t+128, ...; the queue with the largest backlog wins,
        // the resident community counts `stick` (2) times plus 1.5 tiles (a switch costs a 128 KB operand fetch from L2:
        // with plain largest-backlog-first 31 % of the tiles switched, +10 % L2 traffic; measured +5 %), and the thread that saw
        // the winner claims up to TN of its entries with one compare-and-swap (any CTA may serve any community: the load
        // balances itself).  Result in pick[]; count 0 = nothing claimed, -1 = all walkers finished and all queues empty.
        auto choose_and_claim = [&]() {
            const int live = tid == 0 ? *reinterpret_cast<volatile int *>(P.live) : 1;  // read BEFORE the queues
            int key = 0, my_backlog = 0;
            unsigned my_h = 0;
            for (int q = tid; q < P.Q; q += NSVC * 32) {
                const unsigned h = ld_vol(P.claim + q * 8);
                const int backlog = (int)(ld_vol(P.tail + q * 8) - h);
                if (backlog > 0) {
                    // + a CTA-specific tie-breaker so that the CTAs do not all rush to the same queue
                    const int score = min(backlog, 1 << 16) * (q == cur_c ? P.stick : 1) + (q == cur_c ? (3 * TN) / 2 : 0) +
                                      ((q * 7 + (int)blockIdx.x * 13) & 31);
                    const int k2 = (score << 11) | q;
                    if (k2 > key) {
                        key = k2;
                        my_h = h;
                        my_backlog = backlog;
                    }
                }
            }
            int best = key;
#pragma unroll
            for (int o = 16; o; o >>= 1) best = max(best, __shfl_xor_sync(FULL, best, o));
            if (lane == 0) wbest[warp] = best;
            front_barrier();
            const int k = max(max(wbest[0], wbest[1]), max(wbest[2], wbest[3]));
            if (k > 0 && key == k) {  // the one thread that saw the winning queue claims from what it saw (no re-read)
                const int q = k & 2047;
                int n = 0;
                unsigned h = my_h;
                int backlog = my_backlog;
                for (int attempt = 0; attempt < 3 && n == 0 && backlog > 0; attempt++) {
                    const int want = min(backlog, TN);
                    const unsigned old = atomicCAS(P.claim + q * 8, h, h + (unsigned)want);
                    if (old == h) {
                        n = want;
                    } else {  // another CTA claimed meanwhile: retry from what the CAS returned
                        h = old;
                        backlog = (int)(ld_vol(P.tail + q * 8) - h);
                    }
                }
                pick[0] = q;
                pick[1] = (int)h;
                pick[2] = n;
            }
            if (tid == 0) {
                if (k > 0) {
                    empty_scans = 0;
                } else {
                    int n = 0;
                    if (live == 0) {
                        if (++empty_scans >= 2) n = -1;
                    } else {
                        empty_scans = 0;
                    }
                    pick[0] = 0;
                    pick[1] = 0;
                    pick[2] = n;
                }
            }
            front_barrier();
        };
        choose_and_claim();
        uint32_t t = 0;  // tiles issued so far; tile t uses list / accumulator buffer t & 1
        while (true) {
            const int q = pick[0], n = pick[2];
            const uint32_t base = (uint32_t)pick[1];
            front_barrier();  // everyone has read pick[] before it is rewritten
            if (n == 0) {
                idle_polls++;
                __nanosleep(100);
                choose_and_claim();
                continue;
            }
            const uint32_t b = t & 1u;
            // the back team must have drained what tile t-2 left in buffer b (slot list, accumulator)
            if (t >= 2) umma::mbar_wait(bar_free + b, ((t >> 1) - 1u) & 1u);
            if (n < 0) {  // tell the back team to leave
                if (tid == 0) {
                    n_s[b] = -1;
                    mbar_arrive(bar_ent + b);
                }
                break;
            }
            lap(0);
            const int c = q;
            const int n16 = (n + 15) & ~15;
            uint32_t *row_b = row_s + b * TN, *slot_b = slot_s + b * TN;
            if (tid < n) {
                // Relaxed (volatile) loads: everything read on behalf of this entry is addressed THROUGH its value (row ->
                // ld.global.cg of the row from L2, where the requester's red.adds were performed before its release store
                // made the entry visible), and no mutable data is ever read through L1 here, so the L1 invalidation that an
                // acquire at gpu scope costs (CCTL.IVALL per thread per tile) buys nothing.  The claimed entries are
                // reserved, hence published within moments.
                volatile unsigned long long *p = P.ring + (int64_t)q * P.cap + ((base + tid) & (uint32_t)(P.cap - 1));
                unsigned long long e = *p;
                const long long t0 = clock64();
                while (e == EMPTY) {
                    if (clock64() - t0 > WAIT_TIMEOUT) {
                        *P.err = 2;
                        e = 0;
                        break;
                    }
                    e = *p;
                }
                *p = EMPTY;
                row_b[tid] = (uint32_t)e;
                slot_b[tid] = (uint32_t)(e >> 32);
            }
            if (tid == 0) {
                n_s[b] = n;
                n_s[2 + b] = (c != cur_c) ? 1 : 0;  // the tile switches the community: its operand images will be in flight
            }
            // the previous tile's MMAs have read the B operand (and A) completely before either is overwritten
            if (t >= 1) umma::mbar_wait(bar_mma + ((t - 1u) & 1u), ((t - 1u) >> 1) & 1u);
            a_pending = false;
            if (c != cur_c) {
                if (STATS) idle_polls += 1LL << 32;  // high half: community switches
                if (tid == 0) {
                    umma::mbar_expect_tx(bar_a, A_IMG_BYTES);
                    const char *src = P.a_img + (int64_t)c * A_IMG_BYTES;
#pragma unroll
                    for (int qq = 0; qq < 8; qq++) umma::bulk_g2s(smem + L::A_HI + qq * 16384, src + qq * 16384, 16384, bar_a);
                }
                if (warp == 1)
                    *reinterpret_cast<float4 *>(mu_s + 4 * lane) =
                        __ldg(reinterpret_cast<const float4 *>(P.mu + (int64_t)c * D + 4 * lane));
                a_pending = true;
                cur_c = c;
            }
            front_barrier();
            lap(1);
            if (tid == 0) mbar_arrive(bar_ent + b);  // the tile's slots and size are published to the back team
            // ---- B operand: warp w gathers rows w, w+4, ... (16 each), 4 in flight.  The back team's loop (MMAs -> drain the
            // accumulator -> result stores -> signals) is the longer one, so the whole gather stays here: splitting it between
            // the teams was measured at 7.15e8 (half / half), 7.5e8 (three quarters here) and 7.8e8 (all here) -----------------
#pragma unroll
            for (int cls = 0; cls < 4; cls++)
                gather_rows<4, 16, TN / 16>(P, smem + L::B_HI, smem + L::B_LO, row_b, mu_s, warp + 4 * cls, n, lane);
            umma::fence_proxy_async_smem();
            front_barrier();
            lap(2);
            if (tid == 0) mbar_arrive(bar_full + b);  // the B operand is in place; the back team issues the MMAs
                                                      // (issuing blocks the thread for about their duration: off this path)
            a_pending = false;
            t++;
            tiles++;
            rows_served += n;
            lap(3);
            choose_and_claim();  // the next tile is claimed while the tensor cores work on this one
            lap(4);
        }
        if (STATS && P.stats && tid == 0) {
            atomicAdd(reinterpret_cast<unsigned long long *>(P.stats + 0), (unsigned long long)tiles);
            atomicAdd(reinterpret_cast<unsigned long long *>(P.stats + 1), (unsigned long long)rows_served);
            atomicAdd(reinterpret_cast<unsigned long long *>(P.stats + 2), (unsigned long long)idle_polls);
            for (int k = 0; k < 6; k++) atomicAdd(reinterpret_cast<unsigned long long *>(P.stats + 8 + k), (unsigned long long)ph[k]);
        }
    } else if (warp < 2 * NSVC) {
        // =============================== SERVICE, BACK TEAM (warps 4..7) =================================================
        // wait for a tile's MMAs -> read the accumulator back (warp w reads TMEM lanes 32(w%4)..+31 = output coordinates) ->
        // write Y to the requests' result slots -> release-increment the requesters' counters -> hand the buffer back
        if (NW * 32 > 768) reg_dec<SVC_REGS>();
        const int bw = warp - NSVC, btid = threadIdx.x - NSVC * 32;
        const uint32_t a_hi = umma::smem_u32(smem + L::A_HI), a_lo = umma::smem_u32(smem + L::A_LO);
        const uint32_t b_hi = umma::smem_u32(smem + L::B_HI), b_lo = umma::smem_u32(smem + L::B_LO);
        uint32_t par_a = 0;
        long long t_wait = 0, t_epi = 0, tq = tick(), bph[4] = {0, 0, 0, 0};
        auto blap = [&](int k) {
            if (STATS) {
                const long long now = clock64();
                bph[k] += now - tq;
                t_epi += now - tq;
                tq = now;
            }
        };
        for (uint32_t u = 0;; u++) {
            const uint32_t b = u & 1u;
            umma::mbar_wait(bar_ent + b, (u >> 1) & 1u);  // the front team has popped the tile's entries (rows, slots, mu_c)
            const int n = n_s[b];
            if (n < 0) break;
            if (STATS) { const long long now = clock64(); t_wait += now - tq; tq = now; }
            const int n16 = (n + 15) & ~15;
            blap(0);
            umma::mbar_wait(bar_full + b, (u >> 1) & 1u);  // the front team has written the B operand
            if (bw == 0) {
                if (n_s[2 + b]) {  // this tile switched the community: its operand images are still in flight
                    umma::mbar_wait(bar_a, par_a);
                    par_a ^= 1;
                }
                umma::tc_fence_after();
                if (lane == 0) {
                    umma::issue_3xtf32(taddr + b * (uint32_t)TN, a_hi, a_lo, b_hi, b_lo, TN, n16);
                    umma::mma_commit(bar_mma + b);
                }
                __syncwarp();
            }
            umma::mbar_wait(bar_mma + b, (u >> 1) & 1u);
            umma::tc_fence_after();
            blap(1);
            const uint32_t *slot_b = slot_s + b * TN;
            const int a = 32 * bw + lane;
            const uint64_t y_pol = l2_policy_evict_first();
            for (int ch = 0; ch * 16 < n16; ch++) {
                float v[16];
                umma::tmem_ld16(taddr + ((uint32_t)(32 * bw) << 16) + b * (uint32_t)TN + (uint32_t)(ch * 16), v);
#pragma unroll
                for (int qq = 0; qq < 16; qq++) {
                    const int nn = ch * 16 + qq;
                    if (nn < n) {  // Y itself: the requester applies its responsibility
                        float *yp = P.ybuf + (int64_t)slot_b[nn] * D + a;
                        if (P.y_hint) st_f32_hint(yp, v[qq], y_pol);
                        else *yp = v[qq];
                    }
                }
            }
            umma::tc_fence_before();
            back_barrier();  // all four coordinate quarters of every result row are written ...
            blap(2);
            if (btid < n) red_release_add_u32(P.done + slot_b[btid] / (uint32_t)P.vslots, 1u);  // ... before the counter moves
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_free + b);  // 4 arrivals: list and accumulator b are free for tile u+2
            blap(3);
        }
        if (STATS && P.stats && btid == 0) {
            atomicAdd(reinterpret_cast<unsigned long long *>(P.stats + 14), (unsigned long long)t_wait);
            atomicAdd(reinterpret_cast<unsigned long long *>(P.stats + 15), (unsigned long long)t_epi);
            for (int k = 0; k < 4; k++) atomicAdd(reinterpret_cast<unsigned long long *>(P.stats + 16 + k), (unsigned long long)bph[k]);
        }
    } else {
        // =============================== WALKER WARPS ====================================================================
        // A warp interleaves WCTX walks: while the requests of one walk's centre are being served it runs the SGNS pairs of
        // the other walk's centre, so the result latency (claim + gather + MMAs + epilogue, ~25K cycles) is hidden and twice
        // as many requests are outstanding (fuller tiles).  Inside a walk everything stays sequential.
        if (NW * 32 > 768) reg_inc<WALK_REGS>();
        constexpr int NWALK = NW - 2 * NSVC;
        const int ww = warp - 2 * NSVC;
        uint32_t *tok_base = reinterpret_cast<uint32_t *>(smem + L::WTOK) + ww * WCTX * VMAX;
        int32_t *inf_base = reinterpret_cast<int32_t *>(smem + L::WINF) + ww * WCTX * VMAX;
        WalkCtx *cx = reinterpret_cast<WalkCtx *>(smem + L::WCTXS) + ww * WCTX;
        const int W = P.window;
        const int64_t gwarp = (int64_t)blockIdx.x * NWALK + ww;
        const bool walker = gwarp * WCTX < P.active_warps;  // active_warps counts walk CONTEXTS (the Hogwild concurrency cap)
        fused::SgnsArgs SA;
        SA.node = P.node; SA.ctx = P.ctx; SA.table = P.table; SA.mod = P.mod; SA.mu = P.mu; SA.inv_cov = P.inv_cov;
        SA.weight = P.weight; SA.pi = nullptr; SA.ybuf = P.ybuf; SA.K = K; SA.dense = false; SA.o3_on = true; SA.y_evict_first = P.y_hint != 0;
        SA.is_node = P.is_node != 0; SA.lr = P.lr; SA.lambda1 = P.lambda1; SA.nl2 = -P.lambda2;            // c:3132
        SA.clipv = __double2float_rn(__dmul_rn((double)P.lr, 0.1));                                         // c:2556
        uint64_t myA = 1, myC = 0;
#pragma unroll
        for (int k = 0; k < NEG; k++)
            if (lane == k) {
                myA = J.A[k];
                myC = J.C[k];
            }
        bool cursor_done = !walker;  // no walk left to take
        long long t_wait = 0, t_sgns = 0, t_stage = 0, centres = 0;
        if (lane < WCTX) {
            cx[lane].path = nullptr; cx[lane].rwp = nullptr; cx[lane].rnd = 0; cx[lane].len = 0; cx[lane].ci = -1;
            cx[lane].V = 0; cx[lane].wi = 0; cx[lane].expected = 0; cx[lane].have = 0;
        }
        __syncwarp();

        // ---- [stage]: advance context c to its next centre with a non-empty window and post that centre's requests -----------
        auto stage = [&](int c) {
            uint32_t *tokS = tok_base + c * VMAX;
            int32_t *infS = inf_base + c * VMAX;
            const int64_t slot0 = (gwarp * WCTX + c) * P.vslots;
            const uint32_t *path = cx[c].path;
            const int32_t *rwp = cx[c].rwp;
            uint64_t rnd = cx[c].rnd;
            int len = cx[c].len, ci = cx[c].ci, V = 0, n_req = 0;
            uint32_t wi = 0;
            bool have = false;
            while (true) {
                ci++;
                if (ci >= len) {  // next walk
                    unsigned long long w = ~0ULL;
                    if (gwarp * WCTX + c >= P.active_warps) {  // this context is beyond the concurrency cap
                        len = 0;
                        break;
                    }
                    if (!cursor_done) {
                        if (lane == 0) w = atomicAdd(P.walk_cursor, 1ULL);
                        w = __shfl_sync(FULL, w, 0);
                    }
                    if (cursor_done || (int64_t)w >= P.n_walks) {
                        cursor_done = true;
                        len = 0;
                        break;
                    }
                    const int64_t o0 = __ldg(P.walk_off + w), o1 = __ldg(P.walk_off + w + 1);
                    path = P.walks + o0;
                    rwp = P.rw ? P.rw + o0 : nullptr;
                    len = (int)min((int64_t)MAX_SENTENCE_LEN, o1 - o0);
                    // the LCG state the NEXT pair's samples are drawn from (pyx:133-134 order: lookup, then advance)
                    rnd = P.seeds ? P.seeds[w] : (splitmix64(P.base_seed ^ splitmix64((uint64_t)w)) & LCG_MASK);
                    ci = -1;
                    continue;
                }
                wi = __ldg(path + ci);
                if (wi == COMEMB_TOKEN_NONE) continue;
                const int r = rwp ? rwp[ci] : 0;
                const int ja = max(0, ci - W + r), jb = min(len, ci + W + 1 - r);
                int v = 0;
                __syncwarp();
                for (int base = ja; base < jb; base += 32) {  // window rows in position order
                    const int jl = base + lane;
                    uint32_t tk = COMEMB_TOKEN_NONE;
                    if (jl < jb && jl != ci) tk = __ldg(path + jl);
                    const bool valid = tk != COMEMB_TOKEN_NONE;
                    const unsigned vm = __ballot_sync(FULL, valid);
                    if (valid) tokS[v + __popc(vm & ((1u << lane) - 1u))] = tk;
                    v += __popc(vm);
                }
                V = v;
                if (V == 0) continue;
                __syncwarp();  // every lane's row updates of the previous centre happen before the release stores below
                for (int base = 0; base < V; base += 32) {
                    const int vv = base + lane;
                    const bool mine = vv < V;
                    const uint32_t tk = mine ? tokS[vv] : 0u;
                    bool dup = false;  // an earlier window position holds the same node: o3 from the current value, in-warp
                    for (int u = 0; u < vv && mine; u++) dup = dup || (tokS[u] == tk);
                    int cm = -1;
                    if (mine) {
                        cm = __ldg(P.comm + tk);
                        if (cm >= K || (cm >= 0 && __ldg(P.weight + tk) == 0.f)) cm = -1;
                    }
                    const int key = (mine && cm >= 0 && !dup) ? cm : -1 - lane;  // unique negative keys for non-requests
                    const unsigned peers = __match_any_sync(FULL, key);
                    if (key >= 0) {
                        const int leader = __ffs(peers) - 1;
                        unsigned basep = 0;
                        if (lane == leader) basep = atomicAdd(P.tail + cm * 8, (unsigned)__popc(peers));
                        basep = __shfl_sync(peers, basep, leader);
                        const unsigned at = (basep + (unsigned)__popc(peers & ((1u << lane) - 1u))) & (unsigned)(P.cap - 1);
                        st_release_u64(P.ring + (int64_t)cm * P.cap + at,
                                       (unsigned long long)tk | ((unsigned long long)(uint32_t)(slot0 + vv) << 32));
                    }
                    n_req += __popc(__ballot_sync(FULL, key >= 0));
                    if (mine) infS[vv] = cm < 0 ? -1 : (dup ? (cm | INFO_INWARP) : cm);
                }
                have = true;
                break;
            }
            __syncwarp();
            if (lane == 0) {
                cx[c].path = path; cx[c].rwp = rwp; cx[c].rnd = rnd; cx[c].len = len; cx[c].ci = ci; cx[c].V = V;
                cx[c].wi = wi; cx[c].expected += (unsigned)n_req; cx[c].have = have ? 1 : 0;
            }
            __syncwarp();
        };

        int remaining = 0;
        {
            const long long t0 = tick();
            for (int c = 0; c < WCTX; c++) {
                stage(c);
                remaining += cx[c].have;
            }
            t_stage += tick() - t0;
        }
        for (int c = 0; remaining > 0; c = (c + 1) % WCTX) {
            if (!cx[c].have) continue;
            const long long t1 = clock64();
            // ---- [wait] for the centre's results ----------------------------------------------------------------------------
            const unsigned expected = cx[c].expected;
            bool ok = true;
            if (lane == 0) {
                // Poll with a plain volatile load: an acquire load inside the loop makes ptxas emit CCTL.IVALL -- an
                // invalidation of the SM's whole L1 -- per iteration (ncu: 2.25e9 of them per step, 16 % of all stall samples,
                // and every spilled register of the resident warps refetched from L2).  One acquire fence after the counter
                // has been seen is what the protocol needs.
                while ((int)(ld_vol(P.done + gwarp * WCTX + c) - expected) < 0) {
                    __nanosleep(64);
                    if (clock64() - t1 > WAIT_TIMEOUT) {
                        *P.err = 3;
                        ok = false;
                        break;
                    }
                }
                __threadfence();
            }
            ok = __shfl_sync(FULL, ok, 0);
            __syncwarp();
            if (!ok) break;
            const long long t2 = tick();
            // ---- [sgns] -----------------------------------------------------------------------------------------------------
            {
                const uint64_t r0 = cx[c].rnd;
                const int V = cx[c].V;
                uint32_t tnext = (lane < NEG) ? __ldg(P.table + table_slot((myA * r0 + myC) & LCG_MASK, P.mod)) : 0u;
                uint64_t radv = (J.A[NEG] * r0 + J.C[NEG]) & LCG_MASK;
                fused::sgns_centre<ATOMIC, NEG>(SA, cx[c].wi, V, tok_base + c * VMAX, inf_base + c * VMAX, lut,
                                                (gwarp * WCTX + c) * P.vslots, radv, tnext, myA, myC, lane);
                __syncwarp();
                if (lane == 0) cx[c].rnd = lcg_skip(r0, (uint64_t)V * (uint64_t)NEG);  // V pairs consumed NEG draws each
                __syncwarp();
            }
            const long long t3 = tick();
            stage(c);
            if (!cx[c].have) remaining--;
            if (STATS) {
                const long long t4 = clock64();
                t_wait += t2 - t1;
                t_sgns += t3 - t2;
                t_stage += t4 - t3;
                centres++;
            }
        }
        __syncwarp();
        if (walker && lane == 0) {
            __threadfence();
            atomicSub(P.live, 1);
            if (STATS && P.stats && centres) {
                atomicAdd(reinterpret_cast<unsigned long long *>(P.stats + 3), (unsigned long long)t_stage);
                atomicAdd(reinterpret_cast<unsigned long long *>(P.stats + 4), (unsigned long long)t_wait);
                atomicAdd(reinterpret_cast<unsigned long long *>(P.stats + 5), (unsigned long long)t_sgns);
                atomicAdd(reinterpret_cast<unsigned long long *>(P.stats + 6), (unsigned long long)centres);
            }
        }
    }
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(taddr, TMEM_COLS);
}

template <int NEG>
cudaError_t launch_async_t(const AsyncParams &P, bool atomic, int grid, cudaStream_t st) {
    constexpr int NW = ASYNC_WARPS;
    const int smem = AsyncSmem<NW>::TOTAL + 1024;
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        void *args[] = {const_cast<AsyncParams *>(&P)};
        return cudaLaunchCooperativeKernel(reinterpret_cast<void *>(kernel), dim3(grid), dim3(NW * 32), args, (size_t)smem, st);
    };
    return atomic ? go(sg_async_kernel<true, NEG, NW>) : go(sg_async_kernel<false, NEG, NW>);
}

}  // namespace

int launch_umma_prep_a(const float *P, char *out, int K, cudaStream_t st);

// pi in top-1 form (comm / weight per table row), lambda2 != 0.  Returns COMEMB_E_UNSUPPORTED when the shape does not fit
// (the caller falls back to the round-synchronous or the generic kernel).
int launch_sg_fused_async(float *node, float *negemb, const uint32_t *walks, const int64_t *walk_off, int64_t n_walks,
                          const int32_t *reduced_windows, const uint64_t *seeds, uint64_t base_seed, const uint32_t *table,
                          uint64_t table_len, const float *mu, const float *inv_cov, int K, int window, int negative,
                          float lr, float lambda1, float lambda2, int is_node_embedding, bool atomic,
                          const int32_t *comm, const float *weight, cudaStream_t st) {
    constexpr int NW = ASYNC_WARPS, NWALK = NW - 2 * NSVC;
    if (negative < 1 || negative > 7 || window < 1 || 2 * window > VMAX || lambda2 == 0.f) return COMEMB_E_UNSUPPORTED;
    if (K < 1 || !mu || !inv_cov || !comm || !weight) return COMEMB_E_UNSUPPORTED;
    if (!is_node_embedding && negemb == node) return COMEMB_E_UNSUPPORTED;
    if (n_walks <= 0) return 0;
    int dev = 0, sms = 148, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop) return COMEMB_E_UNSUPPORTED;
    int64_t warps = (int64_t)sms * NWALK * WCTX;  // walks in flight (walk contexts)
    if (comemb_opts().max_warps > 0 && comemb_opts().max_warps < warps) warps = comemb_opts().max_warps;
    if (n_walks < warps) warps = n_walks;
    const int grid = (int)((warps + NWALK * WCTX - 1) / (NWALK * WCTX));
    const int n_rep = 1, Q = K;  // one queue per community, served by whichever CTAs find it the fullest
    if (Q > 2047) return COMEMB_E_UNSUPPORTED;
    const int64_t total_warps = (int64_t)grid * NWALK * WCTX;  // walk contexts: each has its result slots and counter
    const int vslots = 2 * window;
    // a walker has at most vslots requests in flight, so total_warps * vslots bounds the entries of one queue that are
    // reserved and not yet popped; twice that keeps a ring position's reuse far away from the pop that freed it
    int64_t cap = 64;
    while (cap < 2 * total_warps * vslots) cap <<= 1;
    size_t off = 0;
    auto carve = [&](size_t bytes) {
        const size_t at = off;
        off = (off + bytes + 1023) & ~(size_t)1023;
        return at;
    };
    const size_t sz_ctl = 256 + (size_t)Q * 64 + (size_t)total_warps * 4;
    const size_t o_ctl = carve(sz_ctl), o_y = carve((size_t)total_warps * vslots * D * 4);
    const size_t o_ring = carve((size_t)Q * cap * 8), o_img = carve((size_t)K * A_IMG_BYTES);
    if (off > ((size_t)24 << 30)) return COMEMB_E_UNSUPPORTED;
    char *scratch = nullptr;
    CUDA_TRY(cudaMallocAsync(&scratch, off, st));
    auto fail = [&](cudaError_t e) {
        cudaFreeAsync(scratch, st);
        return (int)e;
    };
    cudaError_t e = cudaMemsetAsync(scratch + o_ctl, 0, sz_ctl, st);
    if (e != cudaSuccess) return fail(e);
    if ((e = cudaMemsetAsync(scratch + o_ring, 0xFF, (size_t)Q * cap * 8, st)) != cudaSuccess) return fail(e);
    const int h_live = (int)((warps + WCTX - 1) / WCTX);  // walker warps with at least one active context
    if ((e = cudaMemcpyAsync(scratch + o_ctl + 48, &h_live, 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(e);
    const int r = launch_umma_prep_a(inv_cov, scratch + o_img, K, st);
    if (r) return fail((cudaError_t)r);
    AsyncParams P;
    P.node = node; P.ctx = negemb; P.walks = walks; P.walk_off = walk_off; P.n_walks = n_walks; P.rw = reduced_windows;
    P.seeds = seeds; P.base_seed = base_seed; P.table = table; P.mod = make_table_mod(table_len);
    P.mu = mu; P.inv_cov = inv_cov; P.a_img = scratch + o_img; P.comm = comm; P.weight = weight;
    P.K = K; P.window = window; P.is_node = is_node_embedding; P.lr = lr; P.lambda1 = lambda1; P.lambda2 = lambda2;
    P.glut = comemb_lut_device();
    P.ybuf = reinterpret_cast<float *>(scratch + o_y);
    P.ring = reinterpret_cast<unsigned long long *>(scratch + o_ring);
    P.walk_cursor = reinterpret_cast<unsigned long long *>(scratch + o_ctl + 16);
    P.err = reinterpret_cast<int *>(scratch + o_ctl + 32);
    P.live = reinterpret_cast<int *>(scratch + o_ctl + 48);
    P.tail = reinterpret_cast<unsigned *>(scratch + o_ctl + 256);
    P.claim = reinterpret_cast<unsigned *>(scratch + o_ctl + 256 + (size_t)Q * 32);
    P.done = reinterpret_cast<unsigned *>(scratch + o_ctl + 256 + (size_t)Q * 64);
    static const int stick_env = getenv("COMEMB_FUSED_STICK") ? atoi(getenv("COMEMB_FUSED_STICK")) : 2;
    P.stick = stick_env < 1 ? 1 : (stick_env > 16 ? 16 : stick_env);
    P.y_hint = 1;  // +4.5 % on the SBM workload: the 60 MB of result slots stop displacing table rows in L2
    P.cap = cap; P.n_rep = n_rep; P.Q = Q; P.vslots = vslots; P.active_warps = warps;
    static const bool want_stats = STATS && getenv("COMEMB_ROUND_STATS") != nullptr;  // needs -DCOMEMB_ASYNC_STATS=1
    long long *d_stats = nullptr;
    if (want_stats) {
        cudaMalloc(&d_stats, 24 * sizeof(long long));
        cudaMemset(d_stats, 0, 24 * sizeof(long long));
    }
    P.stats = d_stats;
    switch (negative) {
        case 1: e = launch_async_t<1>(P, atomic, grid, st); break;
        case 2: e = launch_async_t<2>(P, atomic, grid, st); break;
        case 3: e = launch_async_t<3>(P, atomic, grid, st); break;
        case 4: e = launch_async_t<4>(P, atomic, grid, st); break;
        case 5: e = launch_async_t<5>(P, atomic, grid, st); break;
        case 6: e = launch_async_t<6>(P, atomic, grid, st); break;
        default: e = launch_async_t<7>(P, atomic, grid, st); break;
    }
    if (want_stats) {
        long long h[24];
        int h_err = 0;
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_stats, sizeof(h), cudaMemcpyDeviceToHost);
        cudaMemcpy(&h_err, P.err, 4, cudaMemcpyDeviceToHost);
        cudaFree(d_stats);
        fprintf(stderr,
                "[async stats] grid %d n_rep %d Q %d err %d: tiles %lld rows %lld (%.1f rows/tile) idle polls %lld community switches %lld | per centre "
                "cycles: stage %lld wait %lld sgns %lld (centres %lld)\n",
                grid, n_rep, Q, h_err, h[0], h[1], h[0] ? (double)h[1] / h[0] : 0.0, h[2] & 0xFFFFFFFFLL, h[2] >> 32, h[3] / (h[6] + 1),
                h[4] / (h[6] + 1), h[5] / (h[6] + 1), h[6]);
        fprintf(stderr, "[async stats] front cycles per tile: wait-free %lld pop %lld gather %lld issue %lld claim-next %lld | back: "
                "wait %lld epilogue+signal %lld\n",
                h[8] / (h[0] + 1), h[9] / (h[0] + 1), h[10] / (h[0] + 1), h[11] / (h[0] + 1), h[12] / (h[0] + 1),
                h[14] / (h[0] + 1), h[15] / (h[0] + 1));
        fprintf(stderr, "[async stats] back cycles per tile: gather-half %lld wait-front+mma %lld epilogue %lld signal %lld\n",
                h[16] / (h[0] + 1), h[17] / (h[0] + 1), h[18] / (h[0] + 1), h[19] / (h[0] + 1));
    }
    cudaFreeAsync(scratch, st);
    return (int)e;
}
