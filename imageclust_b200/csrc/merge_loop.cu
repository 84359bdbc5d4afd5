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
// The loop is a chain of ~N dependent steps with a few hundred KB of traffic each, so it
// is bound by synchronisation latency, not bandwidth.  Design for that:
//   * every block owns a contiguous slice of slots and keeps that slice's {key, size}
//     and NN cache in SHARED MEMORY for the whole loop; only the owner writes them
//     (global copies are written through for read-back and resume);
//   * ONE all-to-all exchange per merge and no separate barrier: at the end of an
//     iteration each block publishes one 128-byte record (its slice's best cached
//     candidate + its part of the freshly written row's minimum) whose 16-byte chunks
//     carry the epoch as a tag; at the start of the next iteration two warps poll all
//     records directly and fold them with shuffles (release fence before the publish,
//     acquire fence after the poll);
//   * every block folds the same records, so all blocks take the same decision without
//     a broadcast, then update their slice of row/column b;
//   * every row caches its kNNK smallest partners (common.cuh); only when ALL of them have died
//     is the row rescanned, by its OWNER block alone (whole row, many 16-byte loads in flight)
//     right after the update -- the row only depends on entries its owner wrote itself or that
//     were published at least one exchange ago.
// HBM roofline: algorithmic bytes = 12*n per merge; reported as merges/s too.
#include "common.cuh"
#include "kernels.h"

namespace ic {

namespace {

constexpr uint32_t kSpinLimit = 1u << 24;
constexpr int kRecWords = 32;  // 128 bytes per block record: one cache line per writer

// record chunks (uint4 each, .w = epoch tag)
//   c0 {row key, dist bits, runner-up dist bits, tag}      slice's best cached candidate
//   c1 {row slot a, partner slot b, size a, tag}
//   c2 {size b, partner key, 0, tag}
//   c3 {key of k, dist bits, slot k, tag}                   best entry of the new row in this slice
//   c4 {size k, runner bits, 0, tag}
constexpr int kChunks = 5;

IC_DEVINL uint4 ld_volatile_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
IC_DEVINL void st_volatile_u4(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
IC_DEVINL void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

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

struct Decision {  // what every block derives from the exchange
    uint64_t m1, m2;   // best / second best (dist bits << 32 | row key)
    int32_t a, b, sa, sb;
    uint32_t pkey;
};
struct NewRow {  // fold of the B-parts: best entry of the previous merge's new row
    uint64_t pack;
    int32_t slot, size;
    uint32_t runner;
};

}  // namespace

size_t merge_loop_record_bytes() { return kRecWords * sizeof(uint32_t); }
// Keys of ALL slots replicated in every block's shared memory when they fit: a whole-row
// rescan then streams only the row itself.
constexpr int64_t kReplicaMaxSlots = 36 * 1024;
bool merge_loop_uses_replica(int64_t n) { return n <= kReplicaMaxSlots; }
size_t merge_loop_smem_bytes(int64_t n, int grid) {
    const int64_t chunk = (n + grid - 1) / grid;
    const int64_t c = chunk > 0 ? chunk : 1;
    size_t bytes = static_cast<size_t>(c) * (kNNK * sizeof(uint4) + sizeof(int2) + 2 * sizeof(int32_t));
    bytes = (bytes + 15) & ~size_t(15);
    if (merge_loop_uses_replica(n)) bytes += static_cast<size_t>((n + 3) / 4 * 4) * sizeof(int32_t);
    return bytes;
}

template <int kT, bool kReplica>
__global__ void __launch_bounds__(kT, 1) merge_loop_kernel(LoopState st, LoopParams prm) {
    constexpr int kW = kT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int blk = blockIdx.x, G = gridDim.x;
    const int32_t n = st.n;
    const int64_t ld = st.ld;
    float* const dm = st.dm;
    // mailboxes: every writer pushes its record into a private line of every reader, so a polled
    // line has exactly one reader and one writer (148 blocks polling shared lines was 3x slower)
    uint4* const records = static_cast<uint4*>(st.records);  // [reader G][2][writer G][8]

    // slot slice owned by this block; its state lives in shared memory
    const int32_t chunk = (n + G - 1) / G;
    const int32_t lo = min(n, blk * chunk), hi = min(n, lo + chunk);
    const int32_t cnt = hi - lo;
    extern __shared__ __align__(16) uint8_t dyn_smem[];
    const size_t c1 = static_cast<size_t>(chunk > 0 ? chunk : 1);
    uint4* const s_nn = reinterpret_cast<uint4*>(dyn_smem);                                  // [chunk][kNNK]
    int2* const s_ks = reinterpret_cast<int2*>(dyn_smem + c1 * kNNK * sizeof(uint4));        // [chunk]
    int32_t* const s_more = reinterpret_cast<int32_t*>(dyn_smem + c1 * (kNNK * sizeof(uint4) + sizeof(int2)));
    int32_t* const s_resc = s_more + c1;
    const int32_t n4 = (n + 3) & ~3;
    int32_t* const s_key = reinterpret_cast<int32_t*>(
        dyn_smem + ((c1 * (kNNK * sizeof(uint4) + sizeof(int2) + 2 * sizeof(int32_t)) + 15) & ~size_t(15)));  // [n4] if kReplica
    for (int32_t i = tid; i < cnt; i += kT) {
        s_ks[i] = __ldcg(st.ks + lo + i);
        s_more[i] = __ldcg(st.nn_more + lo + i);
    }
    for (int32_t i = tid; i < cnt * kNNK; i += kT) s_nn[i] = __ldcg(st.nn + static_cast<int64_t>(lo) * kNNK + i);
    if (kReplica)
        for (int32_t u = tid; u < n4; u += kT) s_key[u] = __ldcg(st.gkey + u);
    const int npw = (G + 31) / 32;  // warps that poll one part of the records (one record per lane)

    __shared__ uint64_t s_m1[kW], s_m2[kW], s_up[kW], s_ur[kW];
    __shared__ Decision s_dec;
    __shared__ NewRow s_new;
    __shared__ Decision s_pdec[kW];
    __shared__ NewRow s_pnew[kW];
    __shared__ uint4 s_pub[kChunks];
    __shared__ int32_t s_rcount;
    __shared__ int32_t s_bwin[2];
    __shared__ TopKScratch s_topk;

    int32_t n_live = st.ctl[CTL_N_LIVE];
    int32_t t = st.ctl[CTL_N_MERGES];  // merges done so far == index of the next merge
    int32_t launched = 0;
    int32_t exhausted = 0;
    int32_t my_rescans = 0;
    long long my_rescan_cycles = 0;
    // pending merge (bookkeeping applied after the next exchange)
    bool pending = false;
    int32_t pa = -1, pb = -1, p_snew = 0, p_keyhi = 0, p_keylo = 0;
    float p_dist = 0.0f;
    uint32_t p_second = kInfBits;  // runner-up among the cached candidates when the pending merge was picked
    // this block's part of the freshly written row (B-part of the record it publishes)
    uint64_t pub_bpack = kPackInf;
    int32_t pub_bslot = -1, pub_bsize = 0;
    uint32_t pub_brun = kInfBits;
    if (tid == 0) s_rcount = 0;
    __syncthreads();

    const bool timed = st.prof != nullptr && blk == 0 && tid == 0;
    long long c_poll = 0, c_fold = 0, c_upd = 0, c_resc = 0, c_pub = 0;
    uint32_t epoch = 0;  // iteration index; records read in iteration i carry tag i+1

    for (;;) {
        // ====== publish: slice argmin over the shared NN cache -> record of this epoch ======
        const long long t0 = timed ? clock64() : 0;
        {
            Top2 top = {kPackInf, kPackInf};
            int32_t w_a = -1, w_b = -1, w_sa = 0, w_sb = 0;
            uint32_t w_pkey = 0;
            for (int32_t i = tid; i < cnt; i += kT) {
                const int32_t s = lo + i;
                if (pending && (s == pa || s == pb)) continue;  // a is retired; b's candidate travels in the B-part
                const int2 k = s_ks[i];
                if (k.x < 0) continue;
                const uint4 q = s_nn[i * kNNK];      // head of the row's partner list
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
            const Top2 wt = warp_top2(top);
            if (lane == 0) {
                s_m1[warp] = wt.m1;
                s_m2[warp] = wt.m2;
            }
            __syncthreads();  // also: every global store of this iteration was issued before this point
            Top2 bt = {s_m1[0], s_m2[0]};
#pragma unroll
            for (int w = 1; w < kW; ++w) top2_merge(bt, s_m1[w], s_m2[w]);
            const uint32_t tag = epoch + 1u;
            const bool none = bt.m1 == kPackInf;
            if (none ? tid == 0 : top.m1 == bt.m1) {  // row keys are unique: exactly one thread
                s_pub[0] = make_uint4(static_cast<uint32_t>(bt.m1), static_cast<uint32_t>(bt.m1 >> 32),
                                      static_cast<uint32_t>(bt.m2 >> 32), tag);
                s_pub[1] = make_uint4(static_cast<uint32_t>(w_a), static_cast<uint32_t>(w_b), static_cast<uint32_t>(w_sa), tag);
                s_pub[2] = make_uint4(static_cast<uint32_t>(w_sb), w_pkey, 0u, tag);
                s_pub[3] = make_uint4(static_cast<uint32_t>(pub_bpack), static_cast<uint32_t>(pub_bpack >> 32),
                                      static_cast<uint32_t>(pub_bslot), tag);
                s_pub[4] = make_uint4(static_cast<uint32_t>(pub_bsize), pub_brun, 0u, tag);
            }
            __syncthreads();
            if (tid < G) {  // push to reader `tid`
                uint4* rec = records + ((static_cast<size_t>(tid) * 2 + (epoch & 1u)) * G + blk) * (kRecWords / 4);
                fence_acq_rel_gpu();  // release: the block's stores (ordered by the bar.sync above) before the record
#pragma unroll
                for (int c = 0; c < kChunks; ++c) st_volatile_u4(rec + c, s_pub[c]);
            }
        }
        const long long t1 = timed ? clock64() : 0;

        // ====== exchange: poll every block's record (one record per lane) and fold with shuffles ======
        {
            const uint32_t tag = epoch + 1u;
            const uint4* base = records + (static_cast<size_t>(blk) * 2 + (epoch & 1u)) * G * (kRecWords / 4);
            if (warp < npw) {  // A-parts: two smallest candidates + the winner's payload
                const int g = warp * 32 + lane;
                Top2 ft = {kPackInf, kPackInf};
                uint4 r1 = make_uint4(0, 0, 0, 0), r2 = make_uint4(0, 0, 0, 0);
                if (g < G) {
                    const uint4* rec = base + static_cast<size_t>(g) * (kRecWords / 4);
                    uint4 r0;
                    uint32_t spins = 0;
                    for (;;) {
                        r0 = ld_volatile_u4(rec + 0);
                        r1 = ld_volatile_u4(rec + 1);
                        r2 = ld_volatile_u4(rec + 2);
                        if (r0.w == tag && r1.w == tag && r2.w == tag) break;
                        if (++spins > kSpinLimit) __trap();  // a protocol bug must not hang the GPU box
                    }
                    ft.m1 = (static_cast<uint64_t>(r0.y) << 32) | r0.x;
                    ft.m2 = (static_cast<uint64_t>(r0.z) << 32) | 0xFFFFFFFFull;  // only its distance matters
                    fence_acq_rel_gpu();  // acquire: everything published before that record
                }
                const uint64_t mine = ft.m1;
                ft = warp_top2(ft);
                const unsigned who = __ballot_sync(0xffffffffu, mine == ft.m1 && mine != kPackInf);
                if (who == 0u) {
                    if (lane == 0) {
                        s_pdec[warp].m1 = kPackInf;
                        s_pdec[warp].m2 = kPackInf;
                    }
                } else if (lane == __ffs(who) - 1) {
                    Decision d;
                    d.m1 = ft.m1;
                    d.m2 = ft.m2;
                    d.a = static_cast<int32_t>(r1.x);
                    d.b = static_cast<int32_t>(r1.y);
                    d.sa = static_cast<int32_t>(r1.z);
                    d.sb = static_cast<int32_t>(r2.x);
                    d.pkey = r2.y;
                    s_pdec[warp] = d;
                }
            } else if (warp < 2 * npw) {  // B-parts: the pending merge's new row: best entry + exact runner-up distance
                const int g = (warp - npw) * 32 + lane;
                uint64_t best = kPackInf;
                uint32_t run = 0xFFFFFFFFu;
                int32_t bslot = -1, bsize = 0;
                if (g < G) {
                    const uint4* rec = base + static_cast<size_t>(g) * (kRecWords / 4);
                    uint4 r3, r4;
                    uint32_t spins = 0;
                    for (;;) {
                        r3 = ld_volatile_u4(rec + 3);
                        r4 = ld_volatile_u4(rec + 4);
                        if (r3.w == tag && r4.w == tag) break;
                        if (++spins > kSpinLimit) __trap();
                    }
                    best = (static_cast<uint64_t>(r3.y) << 32) | r3.x;
                    bslot = static_cast<int32_t>(r3.z);
                    bsize = static_cast<int32_t>(r4.x);
                    run = r4.y;
                    fence_acq_rel_gpu();
                }
                const uint64_t wm = warp_min_u64(best);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) run = min(run, __shfl_xor_sync(0xffffffffu, run, o));
                const unsigned who = __ballot_sync(0xffffffffu, best == wm);
                if (lane == __ffs(who) - 1) {
                    NewRow nr;
                    nr.pack = wm;
                    nr.slot = bslot;
                    nr.size = bsize;
                    nr.runner = run;
                    s_pnew[warp - npw] = nr;
                }
            }
            __syncthreads();
            if (tid == 0) {  // combine the per-warp folds
                Decision d = s_pdec[0];
                Top2 ft = {d.m1, d.m2};
                for (int w = 1; w < npw; ++w) {
                    const Decision o = s_pdec[w];
                    if (o.m1 < d.m1) d = o;
                    top2_merge(ft, o.m1, o.m2);
                }
                d.m1 = ft.m1;
                d.m2 = ft.m2;
                s_dec = d;
            } else if (tid == 32) {
                NewRow nr = s_pnew[0];
                for (int w = 1; w < npw; ++w) {
                    const NewRow o = s_pnew[w];
                    const uint32_t run = min(nr.runner, o.runner);
                    if (o.pack < nr.pack) nr = o;
                    nr.runner = run;
                }
                s_new = nr;
            }
        }
        __syncthreads();
        const long long t2 = timed ? clock64() : 0;

        // ====== decision (identical in every block) ======
        Top2 gt = {s_dec.m1, s_dec.m2};
        const uint64_t g_bp = pending ? s_new.pack : kPackInf;
        bool from_new = false;
        if (pending) {
            const bool b_sel = pack_selectable(g_bp);
            const int32_t new_key = n + t - 1;  // key of the cluster made by merge t-1 (appended last, :241)
            if (kReplica && tid == 1) {
                s_key[pa] = -1;
                s_key[pb] = new_key;
            }
            if (tid == 0) {  // bookkeeping of merge t-1 for the slots this block owns
                if (pa >= lo && pa < hi) {
                    s_ks[pa - lo] = make_int2(-1, 0);
                    st.ks[pa] = make_int2(-1, 0);
                    st.gkey[pa] = -1;
                }
                if (pb >= lo && pb < hi) {
                    const uint4 nb = b_sel ? make_uint4(pack_key(g_bp), static_cast<uint32_t>(g_bp >> 32),
                                                        static_cast<uint32_t>(s_new.slot), static_cast<uint32_t>(s_new.size))
                                           : nn_none();
                    s_ks[pb - lo] = make_int2(new_key, p_snew);
                    s_nn[(pb - lo) * kNNK] = nb;  // only the head of the new row is known: the rest is "more"
#pragma unroll
                    for (int j = 1; j < kNNK; ++j) s_nn[(pb - lo) * kNNK + j] = nn_none();
                    s_more[pb - lo] = b_sel ? 1 : 0;
                    st.ks[pb] = make_int2(new_key, p_snew);
                    st.gkey[pb] = new_key;
                    // trace entry of merge t-1 with the exact runner-up distance
                    const uint32_t second = min(p_second, s_new.runner);
                    const float sd = __uint_as_float(second);
                    const float gap = (sd - p_dist) / fmaxf(p_dist, 1e-30f);
                    st.tr_key_hi[t - 1] = p_keyhi;
                    st.tr_key_lo[t - 1] = p_keylo;
                    st.tr_dist[t - 1] = p_dist;
                    st.tr_size[t - 1] = p_snew;
                    st.tr_gap[t - 1] = gap;
                    if (gap < prm.near_tie_tol) atomicAdd(st.ctl + CTL_NEAR_TIES, 1);
                }
            }
            if (b_sel) {
                const uint64_t cand = (g_bp & 0xFFFFFFFF00000000ull) | static_cast<uint32_t>(new_key);
                from_new = cand < gt.m1;
                top2_insert(gt, cand);
            }
        }
        const int32_t prev_a = pending ? pa : -1, prev_b = pending ? pb : -1;
        pending = false;

        // termination (clustering.go:220 loop condition, :222-225 exhaustion)
        const bool selectable = pack_selectable(gt.m1);
        const bool stop = (n_live <= prm.n_target) || !selectable || (prm.max_merges >= 0 && launched >= prm.max_merges);
        if (stop) {
            if (n_live > prm.n_target && !selectable) exhausted = 1;
            if (blk == 0 && tid == 0) {  // what FindClosestClusters would return next
                st.ctl[CTL_NEXT_HI] = selectable ? static_cast<int32_t>(pack_key(gt.m1)) : -1;
                st.ctl[CTL_NEXT_LO] = selectable ? static_cast<int32_t>(from_new ? pack_key(g_bp) : s_dec.pkey) : -1;
                st.ctl[CTL_NEXT_DIST] = selectable ? static_cast<int32_t>(gt.m1 >> 32) : static_cast<int32_t>(kInfBits);
            }
            break;
        }

        // the merge.  a = row slot (higher key), b = partner slot (lower key)
        const int32_t a = from_new ? prev_b : s_dec.a;
        const int32_t b = from_new ? s_new.slot : s_dec.b;
        const int32_t sa = from_new ? p_snew : s_dec.sa;
        const int32_t sb = from_new ? s_new.size : s_dec.sb;
        const uint32_t key_lo = from_new ? pack_key(g_bp) : s_dec.pkey;
        const float dab = __uint_as_float(static_cast<uint32_t>(gt.m1 >> 32));
        const int32_t snew = sa + sb;
        __syncthreads();  // the owner's shared-memory stores above are visible; s_dec / s_new were read

        // ====== update pass over the own slice: Lance-Williams row / column b ======
        uint64_t ubest = kPackInf;
        int32_t uslot = -1, usize = 0;
        uint32_t urun = kInfBits;
        {
            // a pair is stored in the row of its higher-key cluster (the new cluster always has the
            // highest key, so its distances are one coalesced row write and nothing is mirrored)
            const float* row_a = dm + static_cast<int64_t>(a) * ld;
            float* row_b = dm + static_cast<int64_t>(b) * ld;
            const int32_t key_a = static_cast<int32_t>(pack_key(gt.m1)), key_b = static_cast<int32_t>(key_lo);
            for (int32_t i = tid; i < cnt; i += kT) {
                const int32_t k = lo + i;
                const int2 kk = s_ks[i];
                if (k == a || k == b || kk.x < 0) continue;
                const float* own = dm + static_cast<int64_t>(k) * ld;
                const float dka = __ldcg(kk.x < key_a ? row_a + k : own + a);
                const float dkb = __ldcg(kk.x < key_b ? row_b + k : own + b);
                float v;
                if (kk.y + snew > prm.max_size)
                    v = __uint_as_float(kInfBits);  // inadmissible for good: sizes only grow (:228)
                else
                    v = lance_williams(sa, sb, kk.y, dka, dkb, dab);
                __stcg(row_b + k, v);
                const uint64_t c = pack_cand(v, static_cast<uint32_t>(kk.x));
                if (c < ubest) {
                    ubest = c;
                    uslot = k;
                    usize = kk.y;
                }
                urun = min(urun, min(__float_as_uint(dka), __float_as_uint(dkb)));
                // drop a and b from the row's partner list; rescan only when the list runs dry
                {
                    uint4 e[kNNK];
                    int kept = 0;
                    bool changed = false;
#pragma unroll
                    for (int j = 0; j < kNNK; ++j) {
                        const uint4 q = s_nn[i * kNNK + j];
                        if (q.y == kNoPartner) continue;
                        if (static_cast<int32_t>(q.z) == a || static_cast<int32_t>(q.z) == b) {
                            changed = true;
                            continue;
                        }
                        e[kept++] = q;
                    }
                    if (changed) {
#pragma unroll
                        for (int j = 0; j < kNNK; ++j) s_nn[i * kNNK + j] = j < kept ? e[j] : nn_none();
                        if (kept == 0 && s_more[i]) s_resc[atomicAdd(&s_rcount, 1)] = i;
                    }
                }
            }
        }
        {
            const uint64_t wu = warp_min_u64(ubest);
            const uint64_t wr = warp_min_u64(static_cast<uint64_t>(urun));
            if (lane == 0) {
                s_up[warp] = wu;
                s_ur[warp] = wr;
            }
        }
        __syncthreads();
        const long long t3 = timed ? clock64() : 0;
        {
            uint64_t bu = s_up[0], br = s_ur[0];
#pragma unroll
            for (int w = 1; w < kW; ++w) {
                bu = umin64(bu, s_up[w]);
                br = umin64(br, s_ur[w]);
            }
            // every thread keeps the block's B-part; the publishing thread may be any of them
            pub_bpack = bu;
            pub_brun = static_cast<uint32_t>(br);
            // slot / size of the block's best entry: broadcast through shared memory
            if (bu != kPackInf && ubest == bu) {
                s_bwin[0] = uslot;
                s_bwin[1] = usize;
            }
        }
        const int32_t R = s_rcount;
        if (kReplica && tid == 1) {  // a is retired, b is about to carry the highest key: never partners again
            s_key[a] = -1;
            s_key[b] = -1;
        }
        __syncthreads();
        pub_bslot = pub_bpack != kPackInf ? s_bwin[0] : -1;
        pub_bsize = pub_bpack != kPackInf ? s_bwin[1] : 0;

        // ====== owner rescans: whole rows of the own slice whose cached partners have all died ======
        const long long tr0 = (st.prof != nullptr && tid == 0 && R > 0) ? clock64() : 0;
        for (int32_t ri = 0; ri < R; ++ri) {
            const int32_t i = s_resc[ri];
            const int32_t r = lo + i;
            const int32_t kr = s_ks[i].x;
            const float* row = dm + static_cast<int64_t>(r) * ld;
            Cand2 c;
            cand2_init(c);
            constexpr int kU = kReplica ? 12 : 6;  // 16-byte row loads in flight per thread
            const uint32_t ukr = static_cast<uint32_t>(kr);
            for (int32_t base = 0; base < n4; base += kT * 4 * kU) {
                float4 v[kU];
                int4 kq[kReplica ? 1 : kU];
#pragma unroll
                for (int j = 0; j < kU; ++j) {
                    const int32_t u0 = base + (j * kT + tid) * 4;
                    v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!kReplica) kq[j] = make_int4(-1, -1, -1, -1);
                    if (u0 < n4) {
                        v[j] = __ldcg(reinterpret_cast<const float4*>(row + u0));
                        if (!kReplica) kq[j] = __ldcg(reinterpret_cast<const int4*>(st.gkey + u0));
                    }
                }
#pragma unroll
                for (int j = 0; j < kU; ++j) {
                    if (base + j * kT * 4 >= n4) break;  // block-uniform: nothing of this slab is inside the row
                    const int32_t u0 = base + (j * kT + tid) * 4;
                    int4 kj = make_int4(-1, -1, -1, -1);
                    if (kReplica) {
                        if (u0 < n4) kj = *reinterpret_cast<const int4*>(s_key + u0);
                    } else {
                        kj = kq[j];
                    }
                    const uint32_t ks4[4] = {static_cast<uint32_t>(kj.x), static_cast<uint32_t>(kj.y),
                                             static_cast<uint32_t>(kj.z), static_cast<uint32_t>(kj.w)};
                    const uint32_t vs4[4] = {__float_as_uint(v[j].x), __float_as_uint(v[j].y), __float_as_uint(v[j].z),
                                             __float_as_uint(v[j].w)};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int32_t u = u0 + e;
                        const uint32_t ku = ks4[e];  // retired slots hold -1 == 0xFFFFFFFF: never below kr
                        bool ok = ku < ukr && vs4[e] < kMaxFloatBits;
                        // without the replica the four slots in flux are excluded by index: a and prev_a
                        // are retired, b and prev_b carry the two highest keys
                        if (!kReplica) ok = ok && u != a && u != b && u != prev_a && u != prev_b;
                        if (ok) cand2_insert(c, (static_cast<uint64_t>(vs4[e]) << 32) | ku, u);
                    }
                }
            }
            bool more = false;
            const int m = block_select_topk<kT>(c, s_topk, more);
            if (tid < kNNK) {  // sizes of the listed partners: one round trip
                uint4 nr = nn_none();
                if (tid < m) {
                    const uint64_t p = s_topk.pack[tid];
                    const int32_t ws = s_topk.slot[tid];
                    const int32_t wsz = __ldcg(st.ks + ws).y;
                    nr = make_uint4(pack_key(p), static_cast<uint32_t>(p >> 32), static_cast<uint32_t>(ws),
                                    static_cast<uint32_t>(wsz));
                }
                s_nn[i * kNNK + tid] = nr;
                if (tid == 0) s_more[i] = more ? 1 : 0;
            }
            __syncthreads();
        }
        if (st.prof != nullptr && tid == 0 && R > 0) my_rescan_cycles += clock64() - tr0;
        if (tid == 0) {
            my_rescans += R;
            s_rcount = 0;
        }
        const long long t4 = timed ? clock64() : 0;

        // remember the merge; its bookkeeping is applied after the next exchange
        pending = true;
        pa = a;
        pb = b;
        p_snew = snew;
        p_keyhi = static_cast<int32_t>(pack_key(gt.m1));
        p_keylo = static_cast<int32_t>(key_lo);
        p_dist = dab;
        p_second = static_cast<uint32_t>(gt.m2 >> 32);
        ++t;
        ++launched;
        --n_live;
        ++epoch;
        if (timed) {
            c_pub += t1 - t0;
            c_poll += t2 - t1;
            c_upd += t3 - t2;
            c_resc += t4 - t3;
            c_fold += clock64() - t4;
        }
    }
    if (timed) {
        st.prof[0] = c_pub;
        st.prof[1] = c_poll;
        st.prof[2] = c_upd;
        st.prof[3] = c_resc;
        st.prof[4] = c_fold;
        st.prof[5] = launched;
    }
    // the partner lists live in shared memory during the loop: write the slice back for resume / read-back
    __syncthreads();
    for (int32_t i = tid; i < cnt * kNNK; i += kT) st.nn[static_cast<int64_t>(lo) * kNNK + i] = s_nn[i];
    for (int32_t i = tid; i < cnt; i += kT) st.nn_more[lo + i] = s_more[i];
    if (tid == 0 && my_rescans > 0) atomicAdd(st.ctl + CTL_RESCANS, my_rescans);
    if (st.prof != nullptr && tid == 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(st.prof + 8), static_cast<unsigned long long>(my_rescan_cycles));
    if (blk == 0 && tid == 0) {
        st.ctl[CTL_N_LIVE] = n_live;
        st.ctl[CTL_N_MERGES] = t;
        st.ctl[CTL_EXHAUSTED] = exhausted;
        st.ctl[CTL_DONE] = 1;
    }
}

int merge_loop_threads(int64_t n, int num_sms) {
    (void)n;
    (void)num_sms;
    return 512;
}

namespace {
template <int kT, bool kReplica>
cudaError_t prepare(size_t smem, int* per_sm) {
    cudaError_t e = cudaFuncSetAttribute(merge_loop_kernel<kT, kReplica>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, merge_loop_kernel<kT, kReplica>, kT, smem);
}
template <int kT, bool kReplica>
cudaError_t launch(void** args, int grid, size_t smem, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(merge_loop_kernel<kT, kReplica>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    // cooperative launch: guarantees that all blocks are resident at once (they wait on one another)
    return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(merge_loop_kernel<kT, kReplica>), dim3(grid), dim3(kT),
                                       args, smem, s);
}
}  // namespace

cudaError_t merge_loop_max_grid(int threads, int num_sms, int64_t n, int* grid) {
    int per_sm = 0;
    const size_t smem = merge_loop_smem_bytes(n, num_sms);
    const bool rep = merge_loop_uses_replica(n);
    cudaError_t e = threads == 256 ? (rep ? prepare<256, true>(smem, &per_sm) : prepare<256, false>(smem, &per_sm))
                                   : (rep ? prepare<512, true>(smem, &per_sm) : prepare<512, false>(smem, &per_sm));
    if (e != cudaSuccess) return e;
    *grid = per_sm > 0 ? num_sms : 0;  // one CTA per SM
    if (2 * ((*grid + 31) / 32) > threads / 32 || *grid > threads) *grid = 0;  // one lane per record, one thread per reader
    return cudaSuccess;
}

cudaError_t launch_merge_loop(const LoopState& st, const LoopParams& p, int grid, int threads, cudaStream_t s) {
    if (grid <= 0) return cudaErrorInvalidConfiguration;
    LoopState st_copy = st;
    LoopParams p_copy = p;
    void* args[] = {&st_copy, &p_copy};
    const size_t smem = merge_loop_smem_bytes(st.n, grid);
    const bool rep = merge_loop_uses_replica(st.n);
    if (threads == 256) return rep ? launch<256, true>(args, grid, smem, s) : launch<256, false>(args, grid, smem, s);
    return rep ? launch<512, true>(args, grid, smem, s) : launch<512, false>(args, grid, smem, s);
}

}  // namespace ic
