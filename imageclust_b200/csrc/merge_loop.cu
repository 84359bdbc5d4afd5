// merge_loop.cu -- K3: the whole agglomeration loop as ONE persistent cooperative kernel.
//
// Replaces the body of the reference's merge loop (clustering.go:220-246):
//   FindClosestClusters (:119-133)  -> reduction over the per-row NN cache, u64 order
//                                      (dist bits, row key) == the reference's (d, i, j)
//   maxSize check (:228-234)        -> eager admissibility: an entry whose size sum
//                                      exceeds maxSize is stored as +inf when written,
//                                      so the device never "rejects"; it stops when the
//                                      global minimum is not selectable (:222-225)
//   MergeClusters (:29-47)          -> slot b keeps the merged cluster (key N+t, the
//                                      highest so far == appended last, :241), slot a
//                                      (the larger position) is retired
//   UpdateDistanceMatrix (:76-96)   -> Lance-Williams recurrence from rows a and b
//                                      (two coalesced row reads + one row write + the
//                                      mirrored column write), in double, stored fp32
//   RemoveRowsAndColumns (:100-116) -> nothing moves; retired slots are skipped
//
// Every block owns a contiguous slice of slots and keeps that slice's {key, size} and
// NN cache in SHARED MEMORY for the whole loop (only the owner ever writes them; the
// global copies are written through for read-back, resume and the rare whole-row path).
// One merge = two grid-wide barriers, each phase has at most one L2 and one HBM round trip:
//   phase A  every block folds its slice of the NN cache (shared memory) into one
//            candidate record and scans its slice of each row whose cached partner died
//   phase B  every block folds all records (identically), applies the bookkeeping
//            of the previous merge for the slots it owns, picks the merge, updates
//            its slice of row/column b and collects the rows that need a rescan
// HBM roofline: algorithmic bytes = 12*n per merge; in practice the loop
// is bound by the two barriers + two dependent memory round trips per merge (merges/s).
#include "common.cuh"
#include "kernels.h"

namespace ic {

namespace {

constexpr int kRC = kLoopRescanSlots;
constexpr uint32_t kSpinLimit = 1u << 24;

struct __align__(16) PartA {  // best cached candidate of one block's slice
    uint64_t m1;              // (dist bits << 32) | row key; kPackInf if none
    uint64_t m2;              // runner-up of the slice
    int32_t a, b;             // row slot, partner slot
    int32_t sa, sb;           // their sizes
    uint32_t pkey;            // partner key
    uint32_t pad[3];
};
struct __align__(16) PartB {  // best entry of the freshly written row within one block's slice
    uint64_t pack;            // (dist bits << 32) | key of k
    int32_t slot, size;       // k and its size
    uint32_t runner;          // bits of min over the slice of d(k,a), d(k,b)
    uint32_t pad[3];
};
struct __align__(16) PartR {  // best lower-key partner of a rescanned row within one block's slice
    uint64_t pack;
    int32_t slot, size;
};
static_assert(sizeof(PartA) == 48 && sizeof(PartB) == 32 && sizeof(PartR) == 16, "record layout");

IC_DEVINL uint32_t ld_relaxed_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
IC_DEVINL void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

// Release by one thread is cumulative over the block's earlier writes (ordered by bar.sync);
// polls are relaxed, one acquire fence after the last one.
IC_DEVINL void grid_barrier(uint32_t* bar, uint32_t& target, uint32_t nblocks) {
    __syncthreads();
    target += nblocks;
    if (threadIdx.x == 0) {
        red_release_add_u32(bar, 1u);
        uint32_t spins = 0;
        while (static_cast<int32_t>(ld_relaxed_u32(bar) - target) < 0) {
            if (++spins > kSpinLimit) __trap();  // a protocol bug must not hang the GPU box
        }
        fence_acq_rel_gpu();
    }
    __syncthreads();
}

template <typename T>
IC_DEVINL T ldcg_as(const void* p) {
    return __ldcg(reinterpret_cast<const T*>(p));
}

// Lance-Williams update of Ward's distance (clustering.go:141-144 is the closed form it
// equals): exact integer weights, double arithmetic, one rounding to fp32.  The CPU
// oracle's LW mode (oracle/ward_fast.c) performs the same operations in the same order.
IC_DEVINL float lance_williams(int sa, int sb, int sk, float dka, float dkb, float dab) {
    const double t1 = static_cast<double>(sa + sk) * static_cast<double>(dka);
    const double t2 = static_cast<double>(sb + sk) * static_cast<double>(dkb);
    const double t3 = static_cast<double>(sk) * static_cast<double>(dab);
    const double num = (t1 + t2) - t3;
    return canon_dist(static_cast<float>(num / static_cast<double>(sa + sb + sk)));
}

IC_DEVINL uint4 nn_none() { return make_uint4(kNoPartner, kNoPartner, kNoPartner, 0u); }

}  // namespace

size_t merge_loop_part_a_bytes() { return sizeof(PartA); }
size_t merge_loop_part_b_bytes() { return sizeof(PartB); }
size_t merge_loop_part_r_bytes() { return sizeof(PartR); }
size_t merge_loop_smem_bytes(int64_t n, int grid) {
    const int64_t chunk = (n + grid - 1) / grid;
    return static_cast<size_t>(chunk > 0 ? chunk : 1) * (sizeof(uint4) + sizeof(int2));
}

template <int kT>
__global__ void __launch_bounds__(kT, 1) merge_loop_kernel(LoopState st, LoopParams prm) {
    constexpr int kW = kT / 32;
    static_assert(kW >= kRC + 2, "warp 0 folds the slice records, warp 1 the new row, warps 2.. one rescanned row each");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int blk = blockIdx.x, G = gridDim.x;
    const int32_t n = st.n;
    const int64_t ld = st.ld;
    float* const dm = st.dm;
    PartA* const part_a = static_cast<PartA*>(st.part_a);
    PartB* const part_b = static_cast<PartB*>(st.part_b);
    PartR* const part_r = static_cast<PartR*>(st.part_r);
    int4* const rlist = reinterpret_cast<int4*>(st.rlist);  // [2][n] {slot, key, size, 0}

    // slot slice owned by this block; its state lives in shared memory
    const int32_t chunk = (n + G - 1) / G;
    const int32_t lo = min(n, blk * chunk), hi = min(n, lo + chunk);
    const int32_t cnt = hi - lo;
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    uint4* const s_nn = reinterpret_cast<uint4*>(dyn_smem);
    int2* const s_ks = reinterpret_cast<int2*>(dyn_smem + static_cast<size_t>(chunk > 0 ? chunk : 1) * sizeof(uint4));
    for (int32_t i = tid; i < cnt; i += kT) {
        s_ks[i] = __ldcg(st.ks + lo + i);
        s_nn[i] = __ldcg(st.nn + lo + i);
    }

    __shared__ uint64_t s_m1[kW], s_m2[kW];
    __shared__ uint64_t s_rr[kRC][kW];
    __shared__ uint64_t s_up[kW], s_ur[kW];
    __shared__ PartA s_win;        // winning record among the blocks' slice candidates (+ folded m1/m2)
    __shared__ PartB s_candb;      // best entry of the previous merge's new row (+ folded runner)
    __shared__ PartR s_rfold[kRC];
    __shared__ int32_t s_rl[kRC];  // rows being rescanned this iteration
    __shared__ int2 s_rks[kRC];    // their {key, size}
    __shared__ uint64_t s_big[kW];

    uint32_t bar_target = 0;
    int32_t n_live = st.ctl[CTL_N_LIVE];
    int32_t t = st.ctl[CTL_N_MERGES];  // merges done so far == index of the next merge
    int32_t launched = 0;
    int32_t exhausted = 0;
    // pending merge (bookkeeping applied in the next phase B)
    bool pending = false;
    int32_t pa = -1, pb = -1, p_snew = 0, p_keyhi = 0, p_keylo = 0;
    float p_dist = 0.0f;
    uint32_t p_second = kInfBits;  // runner-up among the cached candidates when the pending merge was picked
    int par = 0;                   // rescan list written by the pending merge
    __syncthreads();

    const bool timed = st.prof != nullptr && blk == 0 && tid == 0;
    long long c_a = 0, c_b1 = 0, c_fold = 0, c_upd = 0, c_b2 = 0;
    long long c_a1 = 0, c_a2 = 0, c_a3 = 0, c_f1 = 0, c_u1 = 0;
    for (;;) {
        // =========================== phase A ===========================
        const long long t0 = timed ? clock64() : 0;
        // the count and (speculatively) the first kRC entries of the rescan list: one round trip
        int32_t R = 0;
        int4 rent = make_int4(-1, 0, 0, 0);
        if (pending) {
            if (tid < kRC) rent = __ldcg(rlist + static_cast<int64_t>(par) * n + tid);
            R = __ldcg(st.rcount + par);
        }
        if (R > kRC) {
            // Rare: many rows lost their partner (e.g. duplicates of one point).  One
            // block per row, whole-row scans over the global state, then one extra barrier.
            const int4* rl = rlist + static_cast<int64_t>(par) * n;
            for (int32_t ri = blk; ri < R; ri += G) {
                const int4 e = __ldcg(rl + ri);
                const int32_t r = e.x, kr = e.y;
                const float* row = dm + static_cast<int64_t>(r) * ld;
                uint64_t best = kPackInf;
                int32_t bslot = -1, bsize = 0;
                for (int32_t u = tid; u < n; u += kT) {
                    if (u == pa || u == pb) continue;
                    const int2 ku = __ldcg(st.ks + u);
                    if (ku.x < 0 || ku.x >= kr) continue;
                    const uint64_t c = pack_cand(__ldcg(row + u), static_cast<uint32_t>(ku.x));
                    if (c < best) {
                        best = c;
                        bslot = u;
                        bsize = ku.y;
                    }
                }
                const uint64_t wmin = warp_min_u64(best);
                __syncthreads();  // s_big free
                if (lane == 0) s_big[warp] = wmin;
                __syncthreads();
                uint64_t bmin = s_big[0];
#pragma unroll
                for (int w = 1; w < kW; ++w) bmin = umin64(bmin, s_big[w]);
                if (!pack_selectable(bmin)) {
                    if (tid == 0) st.nn[r] = nn_none();
                } else if (best == bmin) {
                    st.nn[r] = make_uint4(pack_key(bmin), static_cast<uint32_t>(bmin >> 32),
                                          static_cast<uint32_t>(bslot), static_cast<uint32_t>(bsize));
                }
            }
            if (blk == 0 && tid == 0) {
                atomicAdd(st.ctl + CTL_RESCANS, R);
                atomicAdd(st.ctl + CTL_BIG_RESCANS, 1);
            }
            grid_barrier(st.barrier, bar_target, G);
            // refresh the shared copy of the rows this block owns
            for (int32_t ri = tid; ri < R; ri += kT) {
                const int32_t r = __ldcg(rl + ri).x;
                if (r >= lo && r < hi) s_nn[r - lo] = __ldcg(st.nn + r);
            }
            R = 0;
            __syncthreads();
        }
        if (tid < R) {
            s_rl[tid] = rent.x;
            s_rks[tid] = make_int2(rent.y, rent.z);
        }
        if (R > 0) __syncthreads();
        const long long ta1 = timed ? clock64() : 0;

        // A1: slice of the NN cache -> two smallest (dist, row key) candidates
        // A2: cooperative rescans: this block's slice of every row whose partner died
        Top2 top = {kPackInf, kPackInf};
        int32_t w_a = -1, w_b = -1, w_sa = 0, w_sb = 0;
        uint32_t w_pkey = 0;
        uint64_t rbest[kRC];
        int32_t rslot[kRC], rsize[kRC];
#pragma unroll
        for (int q = 0; q < kRC; ++q) {
            rbest[q] = kPackInf;
            rslot[q] = -1;
            rsize[q] = 0;
        }
        for (int32_t i = tid; i < cnt; i += kT) {
            const int32_t s = lo + i;
            if (pending && (s == pa || s == pb)) continue;  // bookkeeping not applied yet: a is retired, b gets the highest key
            const int2 k = s_ks[i];
            if (k.x < 0) continue;
            bool stale = false;
#pragma unroll
            for (int q = 0; q < kRC; ++q) {
                if (q < R) {
                    stale |= (s == s_rl[q]);  // its own cache is being rebuilt, folded in phase B
                    if (k.x < s_rks[q].x) {
                        const float v = __ldcg(dm + static_cast<int64_t>(s_rl[q]) * ld + s);
                        const uint64_t c = pack_cand(v, static_cast<uint32_t>(k.x));
                        if (c < rbest[q]) {
                            rbest[q] = c;
                            rslot[q] = s;
                            rsize[q] = k.y;
                        }
                    }
                }
            }
            if (stale) continue;
            const uint4 q = s_nn[i];
            if (q.y >= kMaxFloatBits) continue;  // nothing selectable in this row
            const uint64_t cand = (static_cast<uint64_t>(q.y) << 32) | static_cast<uint32_t>(k.x);
            if (cand < top.m1) {
                w_a = s;
                w_b = static_cast<int32_t>(q.z);
                w_sa = k.y;
                w_sb = static_cast<int32_t>(q.w);
                w_pkey = q.x;
            }
            top2_insert(top, cand);
        }
        // block reduce (one sync for everything)
        const long long ta2 = timed ? clock64() : 0;
        const Top2 wt = warp_top2(top);
        if (lane == 0) {
            s_m1[warp] = wt.m1;
            s_m2[warp] = wt.m2;
        }
#pragma unroll
        for (int q = 0; q < kRC; ++q) {
            if (q < R) {
                const uint64_t wm = warp_min_u64(rbest[q]);
                if (lane == 0) s_rr[q][warp] = wm;
            }
        }
        __syncthreads();
        const long long ta3 = timed ? clock64() : 0;
        {
            Top2 bt = {s_m1[0], s_m2[0]};
#pragma unroll
            for (int w = 1; w < kW; ++w) top2_merge(bt, s_m1[w], s_m2[w]);
            if (bt.m1 == kPackInf) {
                if (tid == 0) {
                    PartA rec = {kPackInf, kPackInf, -1, -1, 0, 0, 0u, {0u, 0u, 0u}};
                    part_a[blk] = rec;
                }
            } else if (top.m1 == bt.m1) {  // row keys are unique: exactly one thread
                PartA rec = {bt.m1, bt.m2, w_a, w_b, w_sa, w_sb, w_pkey, {0u, 0u, 0u}};
                part_a[blk] = rec;
            }
#pragma unroll
            for (int q = 0; q < kRC; ++q) {
                if (q < R) {
                    uint64_t bm = s_rr[q][0];
#pragma unroll
                    for (int w = 1; w < kW; ++w) bm = umin64(bm, s_rr[q][w]);
                    if (bm == kPackInf) {
                        if (tid == 0) {
                            PartR rec = {kPackInf, -1, 0};
                            part_r[static_cast<int64_t>(q) * G + blk] = rec;
                        }
                    } else if (rbest[q] == bm) {  // partner keys are unique: exactly one thread
                        PartR rec = {bm, rslot[q], rsize[q]};
                        part_r[static_cast<int64_t>(q) * G + blk] = rec;
                    }
                }
            }
        }
        const long long t1 = timed ? clock64() : 0;
        grid_barrier(st.barrier, bar_target, G);
        const long long t2 = timed ? clock64() : 0;
        c_a += t1 - t0;
        c_b1 += t2 - t1;
        c_a1 += ta1 - t0;
        c_a2 += ta2 - ta1;
        c_a3 += ta3 - ta2;

        // =========================== phase B ===========================
        // B-a: fold every block's records (all blocks compute the same result); one warp per fold
        if (warp == 0) {  // slice candidates: two smallest + the winner's payload
            Top2 ft = {kPackInf, kPackInf};
            uint64_t best = kPackInf;
            int bidx = -1;
            for (int g = lane; g < G; g += 32) {
                const uint4 r0 = ldcg_as<uint4>(part_a + g);
                const uint64_t m1 = (static_cast<uint64_t>(r0.y) << 32) | r0.x;
                const uint64_t m2 = (static_cast<uint64_t>(r0.w) << 32) | r0.z;
                if (m1 < best) {
                    best = m1;
                    bidx = g;
                }
                top2_merge(ft, m1, m2);
            }
            ft = warp_top2(ft);
            const unsigned who = __ballot_sync(0xffffffffu, best == ft.m1 && bidx >= 0);
            if (who == 0u) {
                if (lane == 0) {
                    s_win.m1 = kPackInf;
                    s_win.m2 = kPackInf;
                }
            } else if (lane == __ffs(who) - 1) {
                const uint4 r1 = ldcg_as<uint4>(reinterpret_cast<const uint4*>(part_a + bidx) + 1);
                const uint4 r2 = ldcg_as<uint4>(reinterpret_cast<const uint4*>(part_a + bidx) + 2);
                s_win.m1 = ft.m1;
                s_win.m2 = ft.m2;
                s_win.a = static_cast<int32_t>(r1.x);
                s_win.b = static_cast<int32_t>(r1.y);
                s_win.sa = static_cast<int32_t>(r1.z);
                s_win.sb = static_cast<int32_t>(r1.w);
                s_win.pkey = r2.x;
            }
        } else if (warp == 1) {  // the pending merge's new row: its best entry + the exact runner-up distance
            uint64_t best = kPackInf, run = kPackInf;
            int32_t bslot = -1, bsize = 0;
            if (pending) {
                for (int g = lane; g < G; g += 32) {
                    const uint4 b0 = ldcg_as<uint4>(part_b + g);
                    const uint4 b1 = ldcg_as<uint4>(reinterpret_cast<const uint4*>(part_b + g) + 1);
                    const uint64_t c = (static_cast<uint64_t>(b0.y) << 32) | b0.x;
                    if (c < best) {
                        best = c;
                        bslot = static_cast<int32_t>(b0.z);
                        bsize = static_cast<int32_t>(b0.w);
                    }
                    run = umin64(run, b1.x);
                }
            }
            const uint64_t wm = warp_min_u64(best);
            run = warp_min_u64(run);
            const unsigned who = __ballot_sync(0xffffffffu, best == wm);
            if (lane == __ffs(who) - 1) {
                s_candb.pack = wm;
                s_candb.slot = bslot;
                s_candb.size = bsize;
                s_candb.runner = static_cast<uint32_t>(run);
            }
        } else if (warp - 2 < R) {  // one warp folds one rescanned row
            const int q = warp - 2;
            uint64_t best = kPackInf;
            int32_t bslot = -1, bsize = 0;
            for (int g = lane; g < G; g += 32) {
                const uint4 r = ldcg_as<uint4>(part_r + static_cast<int64_t>(q) * G + g);
                const uint64_t c = (static_cast<uint64_t>(r.y) << 32) | r.x;
                if (c < best) {
                    best = c;
                    bslot = static_cast<int32_t>(r.z);
                    bsize = static_cast<int32_t>(r.w);
                }
            }
            const uint64_t wm = warp_min_u64(best);
            const unsigned who = __ballot_sync(0xffffffffu, best == wm);
            if (lane == __ffs(who) - 1) {
                PartR rec = {wm, bslot, bsize};
                if (!pack_selectable(wm)) {
                    rec.pack = kPackInf;
                    rec.slot = -1;
                    rec.size = 0;
                }
                s_rfold[q] = rec;
            }
        }
        __syncthreads();
        const long long tf1 = timed ? clock64() : 0;
        c_f1 += tf1 - t2;
        Top2 gt = {s_win.m1, s_win.m2};
        const uint64_t g_bp = s_candb.pack;

        // B-b: bookkeeping of the pending merge, by the blocks that own the slots
        int src = 0;  // 0: a slice candidate, 1: the pending merge's new row, 2+q: rescanned row q
        if (pending) {
            const bool b_sel = pack_selectable(g_bp);
            const int32_t new_key = n + t - 1;  // key of the cluster made by merge t-1 (appended last, :241)
            if (tid == 0) {
                if (pa >= lo && pa < hi) {
                    s_ks[pa - lo] = make_int2(-1, 0);
                    st.ks[pa] = make_int2(-1, 0);
                }
                if (pb >= lo && pb < hi) {
                    const uint4 nb = b_sel ? make_uint4(pack_key(g_bp), static_cast<uint32_t>(g_bp >> 32),
                                                        static_cast<uint32_t>(s_candb.slot),
                                                        static_cast<uint32_t>(s_candb.size))
                                           : nn_none();
                    s_ks[pb - lo] = make_int2(new_key, p_snew);
                    s_nn[pb - lo] = nb;
                    st.ks[pb] = make_int2(new_key, p_snew);
                    st.nn[pb] = nb;
                    // trace entry of merge t-1 with the exact runner-up distance
                    const uint32_t second = min(p_second, s_candb.runner);
                    const float sd = __uint_as_float(second);
                    const float gap = (sd - p_dist) / fmaxf(p_dist, 1e-30f);
                    st.tr_key_hi[t - 1] = p_keyhi;
                    st.tr_key_lo[t - 1] = p_keylo;
                    st.tr_dist[t - 1] = p_dist;
                    st.tr_size[t - 1] = p_snew;
                    st.tr_gap[t - 1] = gap;
                    if (gap < prm.near_tie_tol) atomicAdd(st.ctl + CTL_NEAR_TIES, 1);
                    if (R > 0) atomicAdd(st.ctl + CTL_RESCANS, R);
                }
                if (blk == 0) st.rcount[par] = 0;  // everyone read it before the barrier above
            }
            if (b_sel) {
                const uint64_t cand = (g_bp & 0xFFFFFFFF00000000ull) | static_cast<uint32_t>(new_key);
                if (cand < gt.m1) src = 1;
                top2_insert(gt, cand);
            }
            for (int q = 0; q < R; ++q) {
                const PartR rr = s_rfold[q];
                const int32_t r = s_rl[q];
                if (tid == 0 && r >= lo && r < hi) {
                    const uint4 nr = rr.pack == kPackInf
                                         ? nn_none()
                                         : make_uint4(pack_key(rr.pack), static_cast<uint32_t>(rr.pack >> 32),
                                                      static_cast<uint32_t>(rr.slot), static_cast<uint32_t>(rr.size));
                    s_nn[r - lo] = nr;
                    st.nn[r] = nr;
                }
                if (rr.pack != kPackInf) {
                    const uint64_t cand = (rr.pack & 0xFFFFFFFF00000000ull) | static_cast<uint32_t>(s_rks[q].x);
                    if (cand < gt.m1) src = 2 + q;
                    top2_insert(gt, cand);
                }
            }
        }
        pending = false;

        // B-c: termination (clustering.go:220 loop condition, :222-225 exhaustion)
        const bool selectable = pack_selectable(gt.m1);
        const bool stop = (n_live <= prm.n_target) || !selectable || (prm.max_merges >= 0 && launched >= prm.max_merges);
        if (stop) {
            if (n_live > prm.n_target && !selectable) exhausted = 1;
            if (blk == 0 && tid == 0) {  // what FindClosestClusters would return next
                int32_t khi = -1, klo = -1;
                if (selectable) {
                    khi = static_cast<int32_t>(pack_key(gt.m1));
                    klo = src == 0 ? static_cast<int32_t>(s_win.pkey)
                                   : (src == 1 ? static_cast<int32_t>(pack_key(g_bp))
                                               : static_cast<int32_t>(pack_key(s_rfold[src - 2].pack)));
                }
                st.ctl[CTL_NEXT_HI] = khi;
                st.ctl[CTL_NEXT_LO] = klo;
                st.ctl[CTL_NEXT_DIST] = selectable ? static_cast<int32_t>(gt.m1 >> 32) : static_cast<int32_t>(kInfBits);
            }
            break;
        }

        // B-d: the merge.  a = row slot (higher key), b = partner slot (lower key)
        int32_t a, b, sa, sb;
        uint32_t key_lo;
        if (src == 0) {
            a = s_win.a;
            b = s_win.b;
            sa = s_win.sa;
            sb = s_win.sb;
            key_lo = s_win.pkey;
        } else if (src == 1) {
            a = pb;
            b = s_candb.slot;
            sa = p_snew;
            sb = s_candb.size;
            key_lo = pack_key(g_bp);
        } else {
            const PartR rr = s_rfold[src - 2];
            a = s_rl[src - 2];
            b = rr.slot;
            sa = s_rks[src - 2].y;
            sb = rr.size;
            key_lo = pack_key(rr.pack);
        }
        const float dab = __uint_as_float(static_cast<uint32_t>(gt.m1 >> 32));
        const int32_t snew = sa + sb;
        const int npar = par ^ 1;
        __syncthreads();  // the owner's shared-memory stores above are visible to the update pass
        const long long t3 = timed ? clock64() : 0;

        uint64_t ubest = kPackInf;
        int32_t uslot = -1, usize = 0;
        uint32_t urun = kInfBits;
        {
            const float* row_a = dm + static_cast<int64_t>(a) * ld;
            float* row_b = dm + static_cast<int64_t>(b) * ld;
            int4* rl_out = rlist + static_cast<int64_t>(npar) * n;
            for (int32_t i = tid; i < cnt; i += kT) {
                const int32_t k = lo + i;
                const int2 kk = s_ks[i];
                if (k == a || k == b || kk.x < 0) continue;
                const float dka = __ldcg(row_a + k);
                const float dkb = __ldcg(row_b + k);
                const uint4 q = s_nn[i];
                float v;
                if (kk.y + snew > prm.max_size)
                    v = __uint_as_float(kInfBits);  // inadmissible for good: sizes only grow (:228)
                else
                    v = lance_williams(sa, sb, kk.y, dka, dkb, dab);
                __stcg(row_b + k, v);
                __stcg(dm + static_cast<int64_t>(k) * ld + b, v);
                const uint64_t c = pack_cand(v, static_cast<uint32_t>(kk.x));
                if (c < ubest) {
                    ubest = c;
                    uslot = k;
                    usize = kk.y;
                }
                urun = min(urun, min(__float_as_uint(dka), __float_as_uint(dkb)));
                if (q.y != kNoPartner && (static_cast<int32_t>(q.z) == a || static_cast<int32_t>(q.z) == b)) {
                    const int32_t idx = atomicAdd(st.rcount + npar, 1);
                    rl_out[idx] = make_int4(k, kk.x, kk.y, 0);
                }
            }
        }
        const long long tu1 = timed ? clock64() : 0;
        c_u1 += tu1 - t3;
        {
            const uint64_t wu = warp_min_u64(ubest);
            const uint64_t wr = warp_min_u64(static_cast<uint64_t>(urun));
            if (lane == 0) {
                s_up[warp] = wu;
                s_ur[warp] = wr;
            }
        }
        __syncthreads();
        {
            uint64_t bu = s_up[0], br = s_ur[0];
#pragma unroll
            for (int w = 1; w < kW; ++w) {
                bu = umin64(bu, s_up[w]);
                br = umin64(br, s_ur[w]);
            }
            if (bu == kPackInf) {
                if (tid == 0) {
                    PartB rec = {kPackInf, -1, 0, static_cast<uint32_t>(br), {0u, 0u, 0u}};
                    part_b[blk] = rec;
                }
            } else if (ubest == bu) {  // keys are unique: exactly one thread
                PartB rec = {bu, uslot, usize, static_cast<uint32_t>(br), {0u, 0u, 0u}};
                part_b[blk] = rec;
            }
        }
        // remember the merge; its bookkeeping is applied in the next phase B
        pending = true;
        pa = a;
        pb = b;
        p_snew = snew;
        p_keyhi = static_cast<int32_t>(pack_key(gt.m1));
        p_keylo = static_cast<int32_t>(key_lo);
        p_dist = dab;
        p_second = static_cast<uint32_t>(gt.m2 >> 32);
        par = npar;
        ++t;
        ++launched;
        --n_live;
        const long long t4 = timed ? clock64() : 0;
        grid_barrier(st.barrier, bar_target, G);
        if (timed) {
            c_fold += t3 - t2;
            c_upd += t4 - t3;
            c_b2 += clock64() - t4;
        }
    }
    if (timed) {
        st.prof[0] = c_a;
        st.prof[1] = c_b1;
        st.prof[2] = c_fold;
        st.prof[3] = c_upd;
        st.prof[4] = c_b2;
        st.prof[5] = launched;
        st.prof[8] = c_a1;
        st.prof[9] = c_a2;
        st.prof[10] = c_a3;
        st.prof[11] = c_f1;
        st.prof[12] = c_u1;
    }

    if (blk == 0 && tid == 0) {
        st.ctl[CTL_N_LIVE] = n_live;
        st.ctl[CTL_N_MERGES] = t;
        st.ctl[CTL_EXHAUSTED] = exhausted;
        st.ctl[CTL_DONE] = 1;
    }
}

int merge_loop_threads(int64_t n, int num_sms) {
    // a few slots per thread
    return n > static_cast<int64_t>(num_sms) * 512 ? 512 : 384;
}

namespace {
template <int kT>
cudaError_t prepare(size_t smem, int* per_sm) {
    cudaError_t e = cudaFuncSetAttribute(merge_loop_kernel<kT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, merge_loop_kernel<kT>, kT, smem);
}
}  // namespace

cudaError_t merge_loop_max_grid(int threads, int num_sms, int64_t n, int* grid) {
    int per_sm = 0;
    const size_t smem = merge_loop_smem_bytes(n, num_sms);
    cudaError_t e = threads == 512 ? prepare<512>(smem, &per_sm) : prepare<384>(smem, &per_sm);
    if (e != cudaSuccess) return e;
    *grid = per_sm > 0 ? num_sms : 0;  // one CTA per SM: the barrier cost grows with the grid
    return cudaSuccess;
}

cudaError_t launch_merge_loop(const LoopState& st, const LoopParams& p, int grid, int threads, cudaStream_t s) {
    if (grid <= 0) return cudaErrorInvalidConfiguration;
    LoopState st_copy = st;
    LoopParams p_copy = p;
    void* args[] = {&st_copy, &p_copy};
    const size_t smem = merge_loop_smem_bytes(st.n, grid);
    if (threads == 512)
        return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(merge_loop_kernel<512>), dim3(grid), dim3(512),
                                           args, smem, s);
    return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(merge_loop_kernel<384>), dim3(grid), dim3(384), args,
                                       smem, s);
}

}  // namespace ic
