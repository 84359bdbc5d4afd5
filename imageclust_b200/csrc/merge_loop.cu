// merge_loop.cu -- K3: the whole agglomeration loop as ONE persistent kernel per GPU.
//
// Replaces the body of the reference's merge loop (clustering.go:220-246):
//   FindClosestClusters (:119-133)  -> reduction over the per-row partner lists, u64 order
//                                      (dist bits, row key) == the reference's (d, i, j)
//   maxSize check (:228-234)        -> eager admissibility: an entry whose size sum
//                                      exceeds maxSize is stored as +inf when written,
//                                      so the device never "rejects"; it stops when the
//                                      global minimum is not selectable (:222-225)
//   MergeClusters (:29-47)          -> slot b keeps the merged cluster (key N+t, the
//                                      highest so far == appended last, :241), slot a
//                                      (the larger position) is retired
//   UpdateDistanceMatrix (:76-96)   -> Lance-Williams recurrence from rows a and b, in
//                                      double, stored fp32, one coalesced row write
//   RemoveRowsAndColumns (:100-116) -> nothing moves; retired slots are skipped
//
// The loop is a chain of ~N dependent steps with a few hundred KB of traffic each, so it
// is bound by synchronisation latency, not bandwidth.  Design for that:
//   * the matrix is row-block sharded over P ranks (kernels.h); inside a rank every block
//     owns a contiguous slice of slots and keeps that slice's {key, size} and partner
//     lists in SHARED MEMORY for the whole loop;
//   * ONE all-to-all exchange per iteration and no separate barrier: each block pushes a
//     128-byte record (best candidate of its slice + its part of the freshly written row's
//     minimum + its rescan requests) into a private mailbox line of every block of its
//     rank; 16-byte chunks carry (launch generation, epoch) as a tag, readers poll their own
//     lines (one record per lane) and fold with shuffles.  With P > 1 the rank's block 0
//     then pushes the rank's fold into every rank's mailbox (peer-mapped memory over
//     NVLink) and all blocks fold those P records: same decision everywhere, no broadcast;
//   * a row whose cached partners have all died is NOT rescanned by its owner alone (one SM
//     streams a 100k-column row in ~26 us): the row stays in the reduction with a LOWER
//     BOUND (the distance of its last listed partner; the list is sorted and static), its
//     owner publishes a request, and every block of the rank scans its own column window of
//     that row in the same iteration and mails its partial top-k to the owner.  If a bound
//     wins the global reduction nothing is merged in that iteration (a "bubble"): a merge
//     is only taken when the minimum is an exact candidate below every bound, so the merge
//     sequence is exactly the sequential one.
// HBM roofline: algorithmic bytes = 12*n per merge; reported as merges/s too.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "loop_common.cuh"

namespace ic {

namespace {

constexpr uint32_t kSpinLimit = 1u << 24;
constexpr int kRecU4 = 8;  // 128 bytes per record: one cache line per (reader, writer) pair
constexpr int kT = kLoopThreads;
constexpr int kW = kT / 32;
constexpr int kMaxBlocks = 160;  // blocks per rank: one record per lane of 5 polling warps, x3 parts
constexpr int kMaxReqTotal = kMaxBlocks * kReqPerBlock;
constexpr uint32_t kMoreBit = 1u, kDryBit = 2u, kReqBit = 4u;
constexpr int kMaxSplit = 4;  // parts a block's scan window may be split into
constexpr bool kSharedRecords = true;  // block records: one shared line per writer (false: one per reader and writer)
constexpr int kRankboxFlagBytes = 256;

// block record chunks (uint4 each, .w = tag)
//   c0 {row key, dist bits, runner-up dist bits, tag}      slice's best candidate (exact or bound)
//   c1 {row slot a, partner slot b, size a, tag}
//   c2 {size b, partner key, 1 if the candidate is only a lower bound, tag}
//   c3 {key of k, dist bits, slot k, tag}                   best entry of the new row in this slice
//   c4 {size k, runner bits, number of requests, tag}
//   c5 {request 0 row, request 0 key, request 1 row, tag}   rows of this slice to rescan (-1: none);
//   c6 {request 1 key, 0, 0, tag}                           c5, c6 are only written when there are requests
constexpr int kChunks = 7;

// partial record (one per request and scanning block): c0..c3 {partner key, dist bits, slot, tag},
// c4 {count, more, 0, tag}; the owner looks the partners' sizes up when it folds the lists



struct Decision {  // what every block derives from the exchange
    uint64_t m1, m2;   // best / second best (dist bits << 32 | row key)
    int32_t a, b, sa, sb;
    uint32_t pkey;
    uint32_t stale;    // the best candidate is a lower bound, not a pair
};
struct NewRow {  // fold of the B-parts: best entry of the previous merge's new row
    uint64_t pack;
    int32_t slot, size;
    uint32_t runner;
};

}  // namespace

size_t merge_loop_records_bytes(int G) { return static_cast<size_t>(G) * 2 * G * kRecU4 * sizeof(uint4); }
size_t merge_loop_partials_bytes(int G) {
    return static_cast<size_t>(G) * 2 * kReqPerBlock * G * kMaxSplit * kRecU4 * sizeof(uint4);
}
size_t merge_loop_rankbox_bytes() { return kRankboxFlagBytes + 2 * kMaxRanks * kRecU4 * sizeof(uint4); }

// Keys of ALL slots replicated in every block's shared memory when they fit: a row scan then
// streams only the row itself.
constexpr int64_t kReplicaMaxSlots = 36 * 1024;
bool merge_loop_replica_fits(int64_t n) { return n <= kReplicaMaxSlots; }
// rows per rank and slots per block are multiples of 4: every block's slice of a matrix row starts on a 16-byte
// boundary (bulk stores of the new row's slices)
int64_t merge_loop_rows_per_rank(int64_t n, int n_ranks) { return ((n + n_ranks - 1) / n_ranks + 3) / 4 * 4; }
static int64_t slice_slots(int64_t n, int n_ranks, int G) {
    const int64_t C = merge_loop_rows_per_rank(n, n_ranks);
    const int64_t c = ((C + G - 1) / G + 3) / 4 * 4;
    return c > 0 ? c : 4;
}
size_t merge_loop_smem_bytes(int64_t n, int n_ranks, int G, bool replica) {
    const int64_t c = slice_slots(n, n_ranks, G);
    // partner lists, {key,size}, state bits, dry queue (2 entries per slot)
    size_t bytes = static_cast<size_t>(c) * (kNNK * sizeof(uint4) + sizeof(int2) + 3 * sizeof(int32_t));
    bytes = (bytes + 15) & ~size_t(15);
    if (replica) bytes += static_cast<size_t>((n + 3) / 4 * 4) * sizeof(int32_t);
    if (n_ranks > 1) bytes += static_cast<size_t>(c) * sizeof(float);  // staging of the new row's slice
    return bytes;
}

template <bool kReplica, bool kMulti>
__global__ void __launch_bounds__(kT, 1)
merge_loop_kernel(const __grid_constant__ LoopState st, const __grid_constant__ LoopParams prm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int G = static_cast<int>(gridDim.x) / st.n_local;  // blocks per rank
    const int v = static_cast<int>(blockIdx.x) / G;           // local rank of this block
    const int blk = static_cast<int>(blockIdx.x) - v * G;
    const int P = st.n_ranks, rank = st.rank0 + v;
    const int32_t n = st.n, C = st.rows_per_rank;
    const int64_t ld = st.ld;
    const int32_t n4 = (n + 3) & ~3;

    // this rank's views
    SlotKS* const g_ks = st.ks + static_cast<size_t>(v) * n;
    int32_t* const g_key = st.gkey + static_cast<size_t>(v) * n4;
    int32_t* const ctl = st.ctl + v * kCtlWords;
    // mailboxes: every writer pushes its record into a private line of every reader, so a polled
    // line has exactly one reader and one writer (148 blocks polling shared lines was 3x slower)
    uint4* const records = reinterpret_cast<uint4*>(static_cast<uint8_t*>(st.records) + v * st.records_stride);
    uint4* const partials = reinterpret_cast<uint4*>(static_cast<uint8_t*>(st.partials) + v * st.partials_stride);
    float* const dm_own = st.dm_rank[rank];

    // slot slice owned by this block (rows of this rank); its state lives in shared memory
    const int32_t r_lo = min(n, rank * C), r_hi = min(n, r_lo + C);
    const int32_t chunk = (((C + G - 1) / G) + 3) & ~3;
    const int32_t lo = min(r_hi, r_lo + blk * chunk), hi = min(r_hi, lo + chunk);
    const int32_t cnt = hi - lo;
    // column window this block scans when a row of its rank is rescanned
    const int32_t W = (((n4 + G - 1) / G) + 3) & ~3;
    const int32_t w0 = min(n4, blk * W), w1 = min(n4, w0 + W);
    const int split = W > 1024 ? 4 : (W > 512 ? 2 : 1);  // warps per (request, block): ~one batch of 4 x 16-byte loads per lane

    extern __shared__ __align__(16) uint8_t dyn_smem[];
    const size_t c1 = static_cast<size_t>(chunk > 0 ? chunk : 4);
    uint4* const s_nn = reinterpret_cast<uint4*>(dyn_smem);                                  // [chunk][kNNK]
    int2* const s_ks = reinterpret_cast<int2*>(dyn_smem + c1 * kNNK * sizeof(uint4));        // [chunk]
    uint32_t* const s_more = reinterpret_cast<uint32_t*>(dyn_smem + c1 * (kNNK * sizeof(uint4) + sizeof(int2)));
    int32_t* const s_dryq = reinterpret_cast<int32_t*>(s_more + c1);                         // ring, [2 * chunk]
    const int32_t qcap = static_cast<int32_t>(2 * c1);
    const size_t off_key = (c1 * (kNNK * sizeof(uint4) + sizeof(int2) + 3 * sizeof(int32_t)) + 15) & ~size_t(15);
    int32_t* const s_key = reinterpret_cast<int32_t*>(dyn_smem + off_key);  // [n4] if kReplica
    // [chunk] if kMulti: this block's slice of the new row, shipped to the row's owner as ONE bulk store
    float* const s_newrow = reinterpret_cast<float*>(dyn_smem + off_key + (kReplica ? static_cast<size_t>(n4) * 4 : 0));
    const int32_t cnt4 = (cnt + 3) & ~3;  // bulk stores move multiples of 16 bytes; columns >= n of a row are padding

    __shared__ uint64_t s_m1[kW], s_m2[kW], s_up[kW], s_ur[kW];
    __shared__ Decision s_dec;
    __shared__ NewRow s_new;
    __shared__ Decision s_pdec[kW];
    __shared__ NewRow s_pnew[kW];
    __shared__ uint4 s_pub[kChunks];
    __shared__ uint4 s_wpay[kW];   // per warp: payload of its best candidate {a, b, size a << 1 | bound, size b}
    __shared__ uint32_t s_wkey[kW];  //           and the partner's key
    __shared__ int2 s_uwin[kW];    // per warp: {slot, size} of its best entry of the new row
    __shared__ int32_t s_qhead, s_qtail;        // dry-row queue of this slice
    __shared__ int32_t s_nmine[2], s_req_i[2][kReqPerBlock];  // requests this block published, by epoch parity
    __shared__ int32_t s_nreq;                  // requests of the whole rank in this iteration
    __shared__ int4 s_rlist[kMaxReqTotal];      // {row, row key, owner block, request index}
    __shared__ PartList s_wl[kW];
    __shared__ int32_t s_err;
    __shared__ int32_t s_uwork;  // next unprocessed slot of the update pass
    __shared__ __align__(16) uint4 s_rr[2][kMaxRanks][kRecU4];  // leader: staging of the rank records (by epoch parity)

    if (tid == 0) {
        s_qhead = 0;
        s_qtail = 0;
        s_nmine[0] = s_nmine[1] = 0;
        s_nreq = 0;
        s_err = 0;
    }
    for (int32_t i = tid; i < cnt; i += kT) {
        s_ks[i] = __ldcg(g_ks + lo + i);
        s_more[i] = static_cast<uint32_t>(__ldcg(st.nn_more + lo + i)) & (kMoreBit | kDryBit);
    }
    for (int32_t i = tid; i < cnt * kNNK; i += kT) s_nn[i] = __ldcg(st.nn + static_cast<int64_t>(lo) * kNNK + i);
    if (kMulti)
        for (int32_t i = tid; i < static_cast<int32_t>(c1); i += kT) s_newrow[i] = __uint_as_float(kInfBits);
    if (kReplica)
        for (int32_t u = tid; u < n4; u += kT) s_key[u] = __ldcg(g_key + u);
    __syncthreads();
    for (int32_t i = tid; i < cnt; i += kT)  // rows that were dry when the previous launch stopped
        if (s_more[i] & kDryBit) s_dryq[atomicAdd(&s_qtail, 1) % qcap] = i;
    const int npw = (G + 31) / 32;  // warps that poll one part of the records (one record per lane)

    int32_t n_live = ctl[CTL_N_LIVE];
    int32_t t = ctl[CTL_N_MERGES];  // merges done so far == index of the next merge
    int32_t launched = 0;
    int32_t stop_reason = 0;
    int32_t my_rescans = 0, n_bubbles = 0, bubbles_in_a_row = 0;
    // pending merge (bookkeeping applied after the next exchange)
    bool pending = false;
    int32_t pa = -1, pb = -1, p_snew = 0, p_keyhi = 0, p_keylo = 0;
    float p_dist = 0.0f;
    uint32_t p_second = kInfBits;  // runner-up among the candidates when the pending merge was picked
    if (tid < kW) {  // this block's part of the freshly written row (B-part of the record): none yet
        s_up[tid] = kPackInf;
        s_ur[tid] = kInfBits;
    }
    __syncthreads();

    const bool timed = st.prof != nullptr && blockIdx.x == 0 && tid == 0;
    long long c_pub = 0, c_exch = 0, c_scan = 0, c_upd = 0, c_fold = 0;
    long long blk_wait = 0;  // cycles this block waited for the slowest record of every exchange (profile_loop)
    long long c_sub[6] = {0, 0, 0, 0, 0, 0};  // publish: argmin, reduce, fence, stores; exchange: poll, fold
    uint32_t epoch = 0;  // iteration index; records of iteration i carry tag (gen, i+1)
    const uint32_t tagbase = st.gen << 20;
    const uint32_t scan_every = prm.scan_every > 0 ? static_cast<uint32_t>(prm.scan_every) : 1u;

    for (;;) {
        const uint32_t tag = tagbase | (epoch + 1u);
        const uint32_t par = epoch & 1u;
        // ====== publish: slice argmin over the partner lists -> record of this epoch ======
        const long long t0 = timed ? clock64() : 0;
        {
            Top2 top = {kPackInf, kPackInf};
            int32_t w_a = -1, w_b = -1, w_sa = 0, w_sb = 0;
            uint32_t w_pkey = 0, w_stale = 0;
            for (int32_t i = tid; i < cnt; i += kT) {
                const int32_t s = lo + i;
                if (pending && (s == pa || s == pb)) continue;  // a is retired; b's candidate travels in the B-part
                const int2 k = s_ks[i];
                if (k.x < 0) continue;
                const uint4 q = s_nn[i * kNNK];      // head of the row's partner list, or its lower bound
                if (q.y >= kMaxFloatBits) continue;  // nothing selectable in this row
                const uint64_t cand = (static_cast<uint64_t>(q.y) << 32) | static_cast<uint32_t>(k.x);
                if (cand < top.m1) {
                    w_a = s;
                    w_b = static_cast<int32_t>(q.z);
                    w_sa = k.y;
                    w_sb = static_cast<int32_t>(q.w);
                    w_pkey = q.x;
                    w_stale = q.z == kNoPartner ? 1u : 0u;
                }
                top2_insert(top, cand);
            }
            if (tid == 0) {  // rescan requests: oldest dry rows of the slice first
                int32_t nreq = 0;
                // requests go out every scan_every-th iteration only: whenever any row of the rank is rescanned every
                // block pays the scan phase (~7 k cycles), so the rescans are batched; a row that waits keeps its
                // lower bound in the reduction (bubbles stay rare: the bound of a dry row is rarely the minimum)
                while ((epoch % scan_every) == 0u && nreq < kReqPerBlock && s_qhead != s_qtail) {
                    const int32_t i = s_dryq[s_qhead % qcap];
                    ++s_qhead;
                    const int32_t s = lo + i;
                    if (pending && (s == pa || s == pb)) continue;  // the row is gone / is being rebuilt
                    if (s_ks[i].x < 0 || (s_more[i] & (kDryBit | kReqBit)) != kDryBit) continue;
                    s_more[i] |= kReqBit;
                    s_req_i[par][nreq++] = i;
                }
                s_nmine[par] = nreq;
                const int32_t i0 = nreq > 0 ? s_req_i[par][0] : -1, i1 = nreq > 1 ? s_req_i[par][1] : -1;
                s_pub[5] = make_uint4(static_cast<uint32_t>(i0 >= 0 ? lo + i0 : -1),
                                      static_cast<uint32_t>(i0 >= 0 ? s_ks[i0].x : -1),
                                      static_cast<uint32_t>(i1 >= 0 ? lo + i1 : -1), tag);
                s_pub[6] = make_uint4(static_cast<uint32_t>(i1 >= 0 ? s_ks[i1].x : -1), 0u, 0u, tag);
            }
            const long long ta = timed ? clock64() : 0;
            const Top2 wt = warp_top2(top);
            if (top.m1 == wt.m1 && wt.m1 != kPackInf)  // row keys are unique: exactly one lane of the warp
                s_wpay[warp] = make_uint4(static_cast<uint32_t>(w_a), static_cast<uint32_t>(w_b),
                                          (static_cast<uint32_t>(w_sa) << 1) | w_stale, static_cast<uint32_t>(w_sb));
            if (top.m1 == wt.m1 && wt.m1 != kPackInf) s_wkey[warp] = w_pkey;
            if (lane == 0) {
                s_m1[warp] = wt.m1;
                s_m2[warp] = wt.m2;
            }
            __syncthreads();  // also: every global store of this iteration was issued before this point
            if (warp == 0) {  // final fold of the kW warp results (A-part) and of the update pass's B-part
                Top2 bt = {lane < kW ? s_m1[lane] : kPackInf, lane < kW ? s_m2[lane] : kPackInf};
                const uint64_t mine = bt.m1;
                bt = warp_top2(bt);
                const int wwin = __ffs(__ballot_sync(0xffffffffu, mine == bt.m1 && mine != kPackInf)) - 1;
                const uint64_t ub = lane < kW ? s_up[lane] : kPackInf;
                const uint64_t bu = warp_min_u64(ub);
                const uint64_t br = warp_min_u64(lane < kW ? s_ur[lane] : static_cast<uint64_t>(kInfBits));
                const int uwin = __ffs(__ballot_sync(0xffffffffu, ub == bu && bu != kPackInf)) - 1;
                if (lane == 0) {
                    const uint4 pay = wwin >= 0 ? s_wpay[wwin] : make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
                    const uint32_t pkey = wwin >= 0 ? s_wkey[wwin] : 0u;
                    const int2 uw = uwin >= 0 ? s_uwin[uwin] : make_int2(-1, 0);
                    s_pub[0] = make_uint4(static_cast<uint32_t>(bt.m1), static_cast<uint32_t>(bt.m1 >> 32),
                                          static_cast<uint32_t>(bt.m2 >> 32), tag);
                    s_pub[1] = make_uint4(pay.x, pay.y, pay.z >> 1, tag);
                    s_pub[2] = make_uint4(pay.w, pkey, pay.z & 1u, tag);
                    s_pub[3] = make_uint4(static_cast<uint32_t>(bu), static_cast<uint32_t>(bu >> 32),
                                          static_cast<uint32_t>(uw.x), tag);
                    s_pub[4] = make_uint4(static_cast<uint32_t>(uw.y), static_cast<uint32_t>(br),
                                          static_cast<uint32_t>(s_nmine[par]), tag);
                }
            }
            // release: ONE thread fences (every store of the block was ordered before it by the bar.sync above,
            // the pushes below are ordered after it by the next bar.sync -- the grid-barrier idiom); a fence per
            // pushing warp serialises and cost 3x more
            // With P > 1 every store to a peer is a bulk asynchronous store: this thread issued the block's slice of
            // the new row an iteration ago and only has to see it completed (it long has); everything else the block
            // wrote lives in this GPU's own memory, so the release stays at gpu scope -- a system-scope fence here
            // waited ~2.7 us per merge for NVLink acknowledgements.
            if (tid == kT - 1) {
                if (kMulti) bulk_wait_all();
                fence_acq_rel<false>();
            }
            __syncthreads();
            const long long tb = timed ? clock64() : 0;
            long long tc = 0;
            if (kSharedRecords) {
                // one line per writer, polled by every block of the rank: 5 stores per block instead of 5 * G
                // (draining G private copies through one SM's store path took ~3 000 cycles at G = 111)
                const int nch = s_nmine[par] > 0 ? kChunks : kChunks - 2;  // the request chunks travel only when used
                tc = timed ? clock64() : 0;
                if (tid < nch) st_cg_u4(records + (static_cast<size_t>(par) * G + blk) * kRecU4 + tid, s_pub[tid]);
            } else if (tid < G) {  // push to reader `tid`: a private line per (reader, writer) pair
                uint4* rec = records + ((static_cast<size_t>(tid) * 2 + par) * G + blk) * kRecU4;
                tc = timed ? clock64() : 0;
                const int nch = s_nmine[par] > 0 ? kChunks : kChunks - 2;  // the request chunks travel only when used
                for (int c = 0; c < nch; ++c) st_cg_u4(rec + c, s_pub[c]);
            }
            if (timed) {
#ifndef IC_SCAN_PROF
                c_sub[0] += ta - t0;
                c_sub[1] += tb - ta;
                c_sub[2] += tc - tb;
#endif
                c_sub[3] += clock64() - tc;
            }
        }
        // ====== owner: fold the partial lists of the rows this block asked for in the PREVIOUS iteration ======
        // (their scans had a whole iteration to arrive; the rows stayed in the reduction with their bounds, and
        //  this work overlaps the flight time of the records just pushed)
        const long long tf0 = timed ? clock64() : 0;
        const uint32_t ptag = tagbase | epoch, ppar = par ^ 1u;  // tag / parity of iteration epoch-1
        const int32_t nprev = epoch > 0 ? s_nmine[ppar] : 0;
        for (int32_t q = 0; q < nprev; ++q) {
            const int32_t i = s_req_i[ppar][q];
            const int32_t s = lo + i;
            // merged away by the merge decided in that iteration (still pending here): the scanners skipped it
            // and the slot's state is rebuilt by the next decision
            if (pending && (s == pa || s == pb)) continue;  // block uniform
            const int nsend = G * split;            // partial lists per request
            const int npf = (nsend + 31) / 32;      // warps that poll them
            if (warp < npf) {
                const int g = warp * 32 + lane;
                PartList in;
                in.m = 0;
                in.more = 0;
#pragma unroll
                for (int j = 0; j < kNNK; ++j) {
                    in.pk[j] = kPackInf;
                    in.sl[j] = -1;
                    in.sz[j] = 0;
                }
                if (g < nsend) {
                    const uint4* prec = partials + (((static_cast<size_t>(blk) * 2 + ppar) * kReqPerBlock + q) * (G * kMaxSplit) + g) * kRecU4;
                    uint4 e0, e1, e2, e3, z0;
                    uint32_t spins = 0;
                    for (;;) {
                        e0 = ld_volatile_u4(prec + 0);
                        e1 = ld_volatile_u4(prec + 1);
                        e2 = ld_volatile_u4(prec + 2);
                        e3 = ld_volatile_u4(prec + 3);
                        z0 = ld_volatile_u4(prec + 4);
                        if (e0.w == ptag && e1.w == ptag && e2.w == ptag && e3.w == ptag && z0.w == ptag) break;
                        if (++spins > kSpinLimit) __trap();
                    }
                    in.m = static_cast<int32_t>(z0.x);
                    in.more = static_cast<int32_t>(z0.y);
                    in.pk[0] = (static_cast<uint64_t>(e0.y) << 32) | e0.x;
                    in.pk[1] = (static_cast<uint64_t>(e1.y) << 32) | e1.x;
                    in.pk[2] = (static_cast<uint64_t>(e2.y) << 32) | e2.x;
                    in.pk[3] = (static_cast<uint64_t>(e3.y) << 32) | e3.x;
                    in.sl[0] = static_cast<int32_t>(e0.z);
                    in.sl[1] = static_cast<int32_t>(e1.z);
                    in.sl[2] = static_cast<int32_t>(e2.z);
                    in.sl[3] = static_cast<int32_t>(e3.z);
                }
                PartList out;
                warp_merge_lists(in, out);
                if (lane == 0) s_wl[warp] = out;
            }
            __syncthreads();
            if (tid == 0) {  // merge the per-warp lists with the same cut rule
                int ptr[kW];
                for (int w = 0; w < npf; ++w) ptr[w] = 0;
                int m = 0;
                bool cut = false;
                for (int r = 0; r < kNNK && !cut; ++r) {
                    int bw = -1;
                    uint64_t bp = kPackInf;
                    for (int w = 0; w < npf; ++w)
                        if (ptr[w] < s_wl[w].m && s_wl[w].pk[ptr[w]] < bp) {
                            bp = s_wl[w].pk[ptr[w]];
                            bw = w;
                        }
                    if (bw < 0) break;
                    const int p = ptr[bw]++;
                    s_nn[i * kNNK + r] = make_uint4(pack_key(bp), static_cast<uint32_t>(bp >> 32),
                                                    static_cast<uint32_t>(s_wl[bw].sl[p]), static_cast<uint32_t>(s_wl[bw].sz[p]));
                    m = r + 1;
                    cut = ptr[bw] == s_wl[bw].m && s_wl[bw].more != 0;
                }
                bool more = false;
                for (int w = 0; w < npf; ++w) more = more || ptr[w] < s_wl[w].m || s_wl[w].more != 0;
                for (int r = m; r < kNNK; ++r) s_nn[i * kNNK + r] = nn_none();
                s_more[i] = more ? kMoreBit : 0u;  // fresh again (an empty list without `more`: no partner left)
                ++my_rescans;
            }
            __syncthreads();
            if (tid < kNNK) {  // sizes of the listed partners (none of them is in flux): one L2 round trip, off the critical path
                const uint4 q4 = s_nn[i * kNNK + tid];
                if (q4.z != kNoPartner) s_nn[i * kNNK + tid].w = static_cast<uint32_t>(__ldcg(g_ks + q4.z).y);
            }
        }
        const long long tfold = timed ? clock64() - tf0 : 0;
        const long long t1 = timed ? clock64() : 0;

        // ====== exchange: poll every block's record (one record per lane) and fold with shuffles ======
        const long long t1b = (st.prof != nullptr && tid == 0) ? clock64() : 0;
        {
            const uint4* base = kSharedRecords ? records + static_cast<size_t>(par) * G * kRecU4
                                               : records + (static_cast<size_t>(blk) * 2 + par) * G * kRecU4;
            if (warp < npw) {
                // One lane per record.  Only the LAST chunk of the record (c4) is polled: three groups of warps
                // polling five chunks of every record kept ~130 k 16-byte L2 requests in flight per polling round and
                // every block waited ~7 k cycles per exchange.  Every chunk carries its own tag, so the rest of the
                // record is read once and re-read in the (never observed) case that it lags behind c4.
                const int g = warp * 32 + lane;
                Top2 ft = {kPackInf, kPackInf};
                uint4 r1 = make_uint4(0, 0, 0, 0), r2 = r1, r3 = r1, r4 = r1;
                uint64_t best = kPackInf;
                uint32_t run = 0xFFFFFFFFu;
                if (g < G) {
                    const uint4* rec = base + static_cast<size_t>(g) * kRecU4;
                    uint4 r0;
                    uint32_t spins = 0;
                    for (;;) {
                        r4 = ld_volatile_u4(rec + 4);
                        if (r4.w == tag) break;
                        if (++spins > kSpinLimit) __trap();  // a protocol bug must not hang the GPU box
                    }
                    if (timed) c_sub[4] += clock64() - t1;
                    for (;;) {
                        r0 = ld_volatile_u4(rec + 0);
                        r1 = ld_volatile_u4(rec + 1);
                        r2 = ld_volatile_u4(rec + 2);
                        r3 = ld_volatile_u4(rec + 3);
                        if (r0.w == tag && r1.w == tag && r2.w == tag && r3.w == tag) break;
                        if (++spins > kSpinLimit) __trap();
                    }
                    ft.m1 = (static_cast<uint64_t>(r0.y) << 32) | r0.x;
                    ft.m2 = (static_cast<uint64_t>(r0.z) << 32) | 0xFFFFFFFFull;  // only its distance matters
                    best = (static_cast<uint64_t>(r3.y) << 32) | r3.x;
                    run = r4.y;
                    if (r4.z != 0u) {  // rescan requests of that block
                        uint4 r5, r6;
                        for (;;) {
                            r5 = ld_volatile_u4(rec + 5);
                            r6 = ld_volatile_u4(rec + 6);
                            if (r5.w == tag && r6.w == tag) break;
                            if (++spins > kSpinLimit) __trap();
                        }
                        if (static_cast<int32_t>(r5.x) >= 0)
                            s_rlist[atomicAdd(&s_nreq, 1)] = make_int4(static_cast<int32_t>(r5.x), static_cast<int32_t>(r5.y), g, 0);
                        if (static_cast<int32_t>(r5.z) >= 0)
                            s_rlist[atomicAdd(&s_nreq, 1)] = make_int4(static_cast<int32_t>(r5.z), static_cast<int32_t>(r6.x), g, 1);
                    }
                }
                // A-parts: two smallest candidates + the winner's payload
                const uint64_t mine = ft.m1;
                ft = warp_top2(ft);
                const unsigned who = __ballot_sync(0xffffffffu, mine == ft.m1 && mine != kPackInf);
                if (who == 0u) {
                    if (lane == 0) {
                        s_pdec[warp].m1 = kPackInf;
                        s_pdec[warp].m2 = kPackInf;
                        s_pdec[warp].stale = 0u;
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
                    d.stale = r2.z;
                    s_pdec[warp] = d;
                }
                // B-parts: the pending merge's new row: best entry + exact runner-up distance
                const uint64_t wm = warp_min_u64(best);
                run = __reduce_min_sync(0xffffffffu, run);
                const unsigned whob = __ballot_sync(0xffffffffu, best == wm);
                if (lane == __ffs(whob) - 1) {
                    NewRow nr;
                    nr.pack = wm;
                    nr.slot = static_cast<int32_t>(r3.z);
                    nr.size = static_cast<int32_t>(r4.x);
                    nr.runner = run;
                    s_pnew[warp] = nr;
                }
            }
            __syncthreads();
            if (st.prof != nullptr && tid == 0 && v == 0) blk_wait += clock64() - t1b;
            // acquire: one thread fences after ALL polls of the block completed (ordered by the bar.sync above); the
            // data reads of this iteration come after the bar.sync below
            // With P > 1 nobody fences here: every block of the rank completed a system-scope release fence before
            // it wrote its local record, so what the leader forwards is already performed system-wide, and
            // everybody acquires at system scope after the rank records below (a third fence.sys in the chain cost
            // ~3 us per merge).
            if (tid == kT - 1 && !kMulti) fence_acq_rel<false>();
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
            if (kMulti) {
                // ---- second level: the rank's fold goes to every rank, all blocks fold the P rank records ----
                const long long tr0 = timed ? clock64() : 0;
                __syncthreads();
                if (blk == 0 && tid < P) {
                    const Decision d = s_dec;
                    const NewRow nr = s_new;
                    uint4* rec = static_cast<uint4*>(static_cast<void*>(static_cast<uint8_t*>(st.rankbox[tid]) + kRankboxFlagBytes)) +
                                 (static_cast<size_t>(par) * kMaxRanks + rank) * kRecU4;
                    // what this rank's blocks published was visible in this GPU's L2 before their local records were
                    // written, i.e. before the leader could see them.  The record travels as one bulk asynchronous
                    // store per peer, so that the acquire fence below does not wait for NVLink acknowledgements.
                    bulk_wait_read_1();  // the staging line of two iterations ago has been read
                    uint4* stg = s_rr[par][tid];
                    stg[0] = make_uint4(static_cast<uint32_t>(d.m1), static_cast<uint32_t>(d.m1 >> 32),
                                        static_cast<uint32_t>(d.m2 >> 32), tag);
                    stg[1] = make_uint4(static_cast<uint32_t>(d.a), static_cast<uint32_t>(d.b), static_cast<uint32_t>(d.sa), tag);
                    stg[2] = make_uint4(static_cast<uint32_t>(d.sb), d.pkey, d.stale, tag);
                    stg[3] = make_uint4(static_cast<uint32_t>(nr.pack), static_cast<uint32_t>(nr.pack >> 32),
                                        static_cast<uint32_t>(nr.slot), tag);
                    stg[4] = make_uint4(static_cast<uint32_t>(nr.size), nr.runner, 0u, tag);
                    bulk_store(rec, stg, 5 * sizeof(uint4));
                }
                __syncthreads();  // s_dec / s_new were read by the pushing threads
                if (warp == 0) {
                    const uint4* rbase = static_cast<const uint4*>(static_cast<const void*>(
                                             static_cast<const uint8_t*>(st.rankbox[rank]) + kRankboxFlagBytes)) +
                                         static_cast<size_t>(par) * kMaxRanks * kRecU4;
                    Top2 ft = {kPackInf, kPackInf};
                    uint4 r1 = make_uint4(0, 0, 0, 0), r2 = r1, r3 = r1, r4 = r1;
                    uint64_t best = kPackInf;
                    uint32_t run = 0xFFFFFFFFu;
                    if (lane < P) {
                        const uint4* rec = rbase + static_cast<size_t>(lane) * kRecU4;
                        uint4 r0;
                        uint32_t spins = 0;
                        for (;;) {
                            r0 = ld_volatile_u4(rec + 0);
                            r1 = ld_volatile_u4(rec + 1);
                            r2 = ld_volatile_u4(rec + 2);
                            r3 = ld_volatile_u4(rec + 3);
                            r4 = ld_volatile_u4(rec + 4);
                            if (r0.w == tag && r1.w == tag && r2.w == tag && r3.w == tag && r4.w == tag) break;
                            if (++spins > kSpinLimit) __trap();
                        }
                        ft.m1 = (static_cast<uint64_t>(r0.y) << 32) | r0.x;
                        ft.m2 = (static_cast<uint64_t>(r0.z) << 32) | 0xFFFFFFFFull;
                        best = (static_cast<uint64_t>(r3.y) << 32) | r3.x;
                        run = r4.y;
                        // acquire at gpu scope: every data load of the loop is ld.global.cg / volatile, which a B200
                        // never serves from a copy on the requesting GPU when the address is a peer's (measured:
                        // experiments/nvlink_pingpong.cu, 3 244 cycles for every re-read), and the peers' bulk stores
                        // into this GPU's memory land in its own L2.  A system-scope fence here cost 3 400 cycles.
                        fence_acq_rel<false>();
                    }
                    const uint64_t mine = ft.m1;
                    ft = warp_top2(ft);
                    const unsigned who = __ballot_sync(0xffffffffu, mine == ft.m1 && mine != kPackInf);
                    if (who == 0u) {
                        if (lane == 0) {
                            s_dec.m1 = kPackInf;
                            s_dec.m2 = kPackInf;
                            s_dec.stale = 0u;
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
                        d.stale = r2.z;
                        s_dec = d;
                    }
                    const uint64_t wm = warp_min_u64(best);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) run = min(run, __shfl_xor_sync(0xffffffffu, run, o));
                    const unsigned whob = __ballot_sync(0xffffffffu, best == wm && lane < P);
                    if (lane == __ffs(whob) - 1) {
                        NewRow nr;
                        nr.pack = wm;
                        nr.slot = static_cast<int32_t>(r3.z);
                        nr.size = static_cast<int32_t>(r4.x);
                        nr.runner = run;
                        s_new = nr;
                    }
                }
                if (timed) c_sub[5] += clock64() - tr0;
            }
        }
        __syncthreads();
        const long long t2 = timed ? clock64() : 0;

        // ====== decision (identical in every block of every rank) ======
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
                    s_more[pa - lo] = 0u;
                }
                if (pb >= lo && pb < hi) {
                    const uint4 nb = b_sel ? make_uint4(pack_key(g_bp), static_cast<uint32_t>(g_bp >> 32),
                                                        static_cast<uint32_t>(s_new.slot), static_cast<uint32_t>(s_new.size))
                                           : nn_none();
                    s_ks[pb - lo] = make_int2(new_key, p_snew);
                    s_nn[(pb - lo) * kNNK] = nb;  // only the head of the new row is known: the rest is "more"
#pragma unroll
                    for (int j = 1; j < kNNK; ++j) s_nn[(pb - lo) * kNNK + j] = nn_none();
                    s_more[pb - lo] = b_sel ? kMoreBit : 0u;
                }
            }
            if (blk == 0 && tid == 64) {  // the rank's replica of all slots + the merge trace
                g_ks[pa] = make_int2(-1, 0);
                g_key[pa] = -1;
                g_ks[pb] = make_int2(new_key, p_snew);
                g_key[pb] = new_key;
                // trace entry of merge t-1 with the runner-up distance
                const uint32_t second = min(p_second, s_new.runner);
                const float sd = __uint_as_float(second);
                const float gap = (sd - p_dist) / fmaxf(p_dist, 1e-30f);
                const size_t to = static_cast<size_t>(v) * n + (t - 1);
                st.tr_key_hi[to] = p_keyhi;
                st.tr_key_lo[to] = p_keylo;
                st.tr_dist[to] = p_dist;
                st.tr_size[to] = p_snew;
                st.tr_gap[to] = gap;
                if (gap < prm.near_tie_tol) atomicAdd(ctl + CTL_NEAR_TIES, 1);
            }
            if (b_sel) {
                const uint64_t cand = (g_bp & 0xFFFFFFFF00000000ull) | static_cast<uint32_t>(new_key);
                from_new = cand < gt.m1;
                top2_insert(gt, cand);
            }
        }
        const int32_t prev_a = pending ? pa : -1, prev_b = pending ? pb : -1;
        pending = false;

        // termination (clustering.go:220 loop condition, :222-225 exhaustion); a lower bound at the top of
        // the reduction is resolved (bubble) before it may be reported or before max_merges stops the launch
        const bool selectable = pack_selectable(gt.m1);
        const bool bubble = selectable && !from_new && s_dec.stale != 0u;
        if (n_live <= prm.n_target)
            stop_reason = STOP_TARGET;
        else if (!selectable)
            stop_reason = STOP_EXHAUSTED;
        else if (!bubble && prm.max_merges >= 0 && launched >= prm.max_merges)
            stop_reason = STOP_MAX_MERGES;
        else if (epoch + 8u >= (1u << 20))
            stop_reason = STOP_EPOCHS;
        else if (bubbles_in_a_row > chunk + 64)
            stop_reason = STOP_ERROR;  // cannot happen: every bubble serves kReqPerBlock dry rows per block
        if (stop_reason != 0) {
            if (blk == 0 && tid == 0) {  // what FindClosestClusters would return next
                const bool exact = selectable && !bubble;
                ctl[CTL_NEXT_HI] = exact ? static_cast<int32_t>(pack_key(gt.m1)) : -1;
                ctl[CTL_NEXT_LO] = exact ? static_cast<int32_t>(from_new ? pack_key(g_bp) : s_dec.pkey) : -1;
                ctl[CTL_NEXT_DIST] = exact ? static_cast<int32_t>(gt.m1 >> 32) : static_cast<int32_t>(kInfBits);
            }
            break;
        }

        // the merge.  a = row slot (higher key), b = partner slot (lower key)
        const bool merged = !bubble;
        const int32_t a = !merged ? -1 : (from_new ? prev_b : s_dec.a);
        const int32_t b = !merged ? -1 : (from_new ? s_new.slot : s_dec.b);
        const int32_t sa = from_new ? p_snew : s_dec.sa;
        const int32_t sb = from_new ? s_new.size : s_dec.sb;
        const uint32_t key_lo = from_new ? pack_key(g_bp) : s_dec.pkey;
        const float dab = __uint_as_float(static_cast<uint32_t>(gt.m1 >> 32));
        const int32_t snew = sa + sb;
        const int32_t nreq_all = s_nreq;
        const int32_t qt0 = s_qtail;  // dry queue before this iteration's update pass
        if (tid == 0) s_uwork = 0;
        __syncthreads();  // the owner's shared-memory stores above are visible; s_dec / s_new / s_nreq were read

        // Lance-Williams inputs of this thread's first slot: issued now so that their DRAM round trip overlaps
        // the row scans below
        const int32_t qa = merged ? a / C : 0, qb = merged ? b / C : 0;
        const float* row_a = merged ? st.dm_rank[qa] + static_cast<int64_t>(a - qa * C) * ld : nullptr;
        float* row_b = merged ? st.dm_rank[qb] + static_cast<int64_t>(b - qb * C) * ld : nullptr;
        // warps [0, ns) scan the requested rows while the others update: the two passes are independent.  A wide
        // window is split over kSplit warps, each mailing its own partial list (the owner folds G * split lists one
        // iteration later, off the critical path), so that a scanning warp needs one batch of loads, not two.
        constexpr int kScanWarps = (3 * kW) / 4;  // the update pass of a block is short: most warps may scan
        const int nparts = nreq_all * split;                 // (request, part) pairs of this iteration
        const int ns = min(nparts, kScanWarps);

        // ====== cooperative row scans: this block's column window of every requested row of the rank ======
        // scan_part: one warp scans part `sub` of `parts` of the window for request j -> its sorted partial list
        auto scan_part = [&](int32_t j, int sub, int parts) {
#ifdef IC_SCAN_PROF
            const long long ts0 = timed ? clock64() : 0;
            if (timed) c_sub[0] += ts0 - t2;  // decision + barrier
#endif
            const int4 rq = s_rlist[j];
            const int32_t r = rq.x;
            PartList out;
            out.m = 0;
            out.more = 0;
#pragma unroll
            for (int q = 0; q < kNNK; ++q) {
                out.pk[q] = kPackInf;
                out.sl[q] = -1;
                out.sz[q] = 0;
            }
            if (r != a && r != b) {  // else: merged away in this very iteration, its owner drops the request
                const uint32_t ukr = static_cast<uint32_t>(rq.y);
                const float* rowp = dm_own + static_cast<int64_t>(r - r_lo) * ld;
                const int32_t part = ((((w1 - w0) >> 2) + parts - 1) / parts) << 2;
                const int32_t sw0 = min(w1, w0 + sub * part), sw1 = min(w1, sw0 + part);
                ScanCand c;
                scan_init(c);
                constexpr int kU = 4;  // 16-byte row loads in flight per lane
                for (int32_t base = sw0; base < sw1; base += 128 * kU) {
                    float4 vv[kU];
                    int4 kq[kU];
#pragma unroll
                    for (int q = 0; q < kU; ++q) {
                        const int32_t u0 = base + (q * 32 + lane) * 4;
                        vv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                        kq[q] = make_int4(-1, -1, -1, -1);
                        if (u0 < sw1) {
                            vv[q] = __ldcg(reinterpret_cast<const float4*>(rowp + u0));
                            kq[q] = kReplica ? *reinterpret_cast<const int4*>(s_key + u0)
                                             : __ldcg(reinterpret_cast<const int4*>(g_key + u0));
                        }
                    }
#pragma unroll
                    for (int q = 0; q < kU; ++q) {
                        const int32_t u0 = base + (q * 32 + lane) * 4;
                        const uint32_t ks4[4] = {static_cast<uint32_t>(kq[q].x), static_cast<uint32_t>(kq[q].y),
                                                 static_cast<uint32_t>(kq[q].z), static_cast<uint32_t>(kq[q].w)};
                        const uint32_t vs4[4] = {__float_as_uint(vv[q].x), __float_as_uint(vv[q].y),
                                                 __float_as_uint(vv[q].z), __float_as_uint(vv[q].w)};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            // one compare per element; survivors (a handful per lane) take the full test
                            if (vs4[e] <= static_cast<uint32_t>(c.c2 >> 32)) {
                                const int32_t u = u0 + e;
                                const uint32_t ku = ks4[e];  // retired slots hold -1 == 0xFFFFFFFF: never below the row's key
                                // the four slots in flux are excluded by index: a and prev_a are retired, b and
                                // prev_b carry the two highest keys (their replica entries may lag one exchange)
                                if (ku < ukr && vs4[e] < kMaxFloatBits && u != a && u != b && u != prev_a && u != prev_b)
                                    scan_insert(c, (static_cast<uint64_t>(vs4[e]) << 32) | ku, u);
                            } else {
                                c.extra = true;
                            }
                        }
                    }
                }
#ifdef IC_SCAN_PROF
                const long long ts1 = timed ? clock64() : 0;
                if (timed) c_sub[1] += ts1 - ts0;  // loads + filtering
#endif
                bool more = false;
                out.m = warp_select_scan(c, out.pk, out.sl, more);
#ifdef IC_SCAN_PROF
                if (timed) c_sub[2] += clock64() - ts1;  // selection
#endif
                out.more = more ? 1 : 0;
            }
            if (r != a && r != b) {  // mail the list of this part to the row's owner
                uint4* prec = partials + (((static_cast<size_t>(rq.z) * 2 + par) * kReqPerBlock + rq.w) * (G * kMaxSplit) +
                                          blk * parts + sub) * kRecU4;
                if (lane < kNNK) {
                    const uint64_t myp = sel4(out.pk, lane);
                    st_volatile_u4(prec + lane, lane < out.m ? make_uint4(pack_key(myp), static_cast<uint32_t>(myp >> 32),
                                                                          static_cast<uint32_t>(sel4(out.sl, lane)), tag)
                                                             : make_uint4(kNoPartner, kNoPartner, kNoPartner, tag));
                } else if (lane == kNNK) {
                    st_volatile_u4(prec + kNNK, make_uint4(static_cast<uint32_t>(out.m), static_cast<uint32_t>(out.more), 0u, tag));
                }
            }
        };
        for (int32_t jj = warp; warp < ns && jj < nparts; jj += ns) scan_part(jj / split, jj % split, split);
        const long long t3 = timed ? clock64() : 0;

        // ====== update pass over the own slice: Lance-Williams row b ======
        uint64_t ubest = kPackInf;
        int32_t uslot = -1, usize = 0;
        uint32_t urun = kInfBits;
        if (merged) {
            // The matrix is kept symmetric: d(k,a) and d(k,b) come from rows a and b (two coalesced reads; gathering
            // the entries of newer clusters from their own rows cost one 32-byte sector per value on the critical
            // path, ~6 MB per merge at N = 100k), the result goes to row b (coalesced) and to dm[k][b] (a scattered
            // store into this block's own row: fire and forget).  Rows a and b may live on another rank:
            // peer-mapped loads / stores.
            // Work units of 64 slots are handed out dynamically: the warps that scanned rows above join late.
            auto update_slot = [&](int32_t i, int2 kk, float dka, float dkb) {
                const int32_t k = lo + i;
                float val;
                if (kk.y + snew > prm.max_size)
                    val = __uint_as_float(kInfBits);  // inadmissible for good: sizes only grow (:228)
                else
                    val = lance_williams(sa, sb, kk.y, dka, dkb, dab);
                if (kMulti)
                    s_newrow[i] = val;       // the new cluster's row: staged, shipped to its owner as one bulk store
                else
                    __stcg(row_b + k, val);  // the new cluster's row (coalesced)
                __stcg(dm_own + static_cast<int64_t>(k - r_lo) * ld + b, val);       // mirrored entry in this block's own row k
                const uint64_t cd = pack_cand(val, static_cast<uint32_t>(kk.x));
                if (cd < ubest) {
                    ubest = cd;
                    uslot = k;
                    usize = kk.y;
                }
                urun = min(urun, min(__float_as_uint(dka), __float_as_uint(dkb)));
                // drop a and b from the row's partner list; when the list runs dry the row keeps the
                // distance of its last listed partner as a lower bound and queues for a rescan
                {
                    uint4 e[kNNK];
                    int kept = 0;
                    bool changed = false;
                    uint32_t last_removed = 0u;
#pragma unroll
                    for (int j = 0; j < kNNK; ++j) {
                        const uint4 q = s_nn[i * kNNK + j];
                        if (q.z == kNoPartner) continue;  // empty entry or a bound
                        if (static_cast<int32_t>(q.z) == a || static_cast<int32_t>(q.z) == b) {
                            changed = true;
                            last_removed = q.y;
                            continue;
                        }
                        e[kept++] = q;
                    }
                    if (changed) {
                        const bool dry = kept == 0 && (s_more[i] & kMoreBit) != 0u;
#pragma unroll
                        for (int j = 0; j < kNNK; ++j) s_nn[i * kNNK + j] = j < kept ? e[j] : nn_none();
                        if (dry) {
                            s_nn[i * kNNK] = nn_bound(last_removed);
                            s_more[i] = kMoreBit | kDryBit;
                            s_dryq[atomicAdd(&s_qtail, 1) % qcap] = i;
                        }
                    }
                }
            };
            for (;;) {
                int32_t base = 0;
                if (lane == 0) base = atomicAdd(&s_uwork, 64);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= cnt) break;
                const int32_t i0 = base + lane, i1 = base + 32 + lane;
                int2 kk0 = make_int2(-1, 0), kk1 = make_int2(-1, 0);
                float da0 = 0.f, db0 = 0.f, da1 = 0.f, db1 = 0.f;
                bool ok0 = false, ok1 = false;
                if (i0 < cnt) {
                    kk0 = s_ks[i0];
                    const int32_t k = lo + i0;
                    ok0 = k != a && k != b && kk0.x >= 0;
                    if (ok0) {
                        da0 = __ldcg(row_a + k);
                        db0 = __ldcg(row_b + k);
                    }
                }
                if (i1 < cnt) {
                    kk1 = s_ks[i1];
                    const int32_t k = lo + i1;
                    ok1 = k != a && k != b && kk1.x >= 0;
                    if (ok1) {
                        da1 = __ldcg(row_a + k);
                        db1 = __ldcg(row_b + k);
                    }
                }
                if (ok0)
                    update_slot(i0, kk0, da0, db0);
                else if (kMulti && i0 < cnt)
                    s_newrow[i0] = __uint_as_float(kInfBits);  // columns of retired slots / of a and b: never read
                if (ok1)
                    update_slot(i1, kk1, da1, db1);
                else if (kMulti && i1 < cnt)
                    s_newrow[i1] = __uint_as_float(kInfBits);
            }
        }
        {  // per-warp part of the new row's minimum; the block's fold happens in the next publish
            const uint64_t wu = warp_min_u64(ubest);
            const uint64_t wr = warp_min_u64(static_cast<uint64_t>(urun));
            if (ubest == wu && wu != kPackInf) s_uwin[warp] = make_int2(uslot, usize);
            if (lane == 0) {
                s_up[warp] = wu;
                s_ur[warp] = wr;
            }
        }
        if (kReplica && merged && tid == 1) {  // a is retired, b is about to carry the highest key: never partners again
            s_key[a] = -1;
            s_key[b] = -1;
        }
        if (tid == 0 && s_qtail - s_qhead > qcap) s_err = 1;  // dry queue overflow (cannot happen)
        __syncthreads();
        if (kMulti && merged && tid == kT - 1 && cnt4 > 0) bulk_store(row_b + lo, s_newrow, static_cast<uint32_t>(cnt4) * 4u);
        const long long t4 = timed ? clock64() : 0;

        // rows that ran dry in this update pass will be scanned one or two iterations from now: pull them into L2
        // (a cold 100k-column row costs every scanning warp two DRAM round trips otherwise)
        {
            const int32_t qt1 = s_qtail;
            for (int32_t qi = qt0; qi != qt1; ++qi) {
                const float* rowp = dm_own + static_cast<int64_t>(lo + s_dryq[qi % qcap] - r_lo) * ld;
                for (int32_t off = tid * 32; off < n; off += kT * 32)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(rowp + off));
            }
        }
        if (tid == 0) {
            s_nreq = 0;
        }

        // remember the merge; its bookkeeping is applied after the next exchange
        if (merged) {
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
            bubbles_in_a_row = 0;
        } else {
            ++n_bubbles;
            ++bubbles_in_a_row;
        }
        ++epoch;
        if (timed) {
            c_pub += t1 - t0 - tfold;
            c_fold += tfold;
            c_exch += t2 - t1;
            c_scan += t3 - t2;
            c_upd += t4 - t3;
        }
    }
    if (st.prof != nullptr && tid == 0 && v == 0 && blk < 240) st.prof[16 + blk] = blk_wait;
    if (timed) {
        st.prof[0] = c_pub;
        st.prof[1] = c_exch;
        st.prof[2] = c_upd;
        st.prof[3] = c_scan;
        st.prof[4] = c_fold;
        st.prof[5] = launched;
        st.prof[6] = epoch;
        for (int i = 0; i < 6; ++i) st.prof[10 + i] = c_sub[i];
    }
    if (kMulti && (tid == kT - 1 || (blk == 0 && tid < P))) bulk_wait_all();  // this thread's bulk stores have landed
    // the partner lists live in shared memory during the loop: write the slice back for resume / read-back
    __syncthreads();
    for (int32_t i = tid; i < cnt * kNNK; i += kT) st.nn[static_cast<int64_t>(lo) * kNNK + i] = s_nn[i];
    for (int32_t i = tid; i < cnt; i += kT) st.nn_more[lo + i] = static_cast<int32_t>(s_more[i] & (kMoreBit | kDryBit));
    if (tid == 0 && my_rescans > 0) atomicAdd(ctl + CTL_RESCANS, my_rescans);
    if (tid == 0 && s_err) atomicExch(ctl + CTL_ERROR, 1);
    if (blk == 0 && tid == 0) {
        ctl[CTL_N_LIVE] = n_live;
        ctl[CTL_N_MERGES] = t;
        ctl[CTL_EXHAUSTED] = stop_reason == STOP_EXHAUSTED ? 1 : 0;
        ctl[CTL_BUBBLES] = ctl[CTL_BUBBLES] + n_bubbles;
        ctl[CTL_STOP] = stop_reason;
        if (stop_reason == STOP_ERROR) ctl[CTL_ERROR] = 2;
        __threadfence();
        ctl[CTL_DONE] = 1;
    }
}

// ---- rank barrier: the P single-GPU processes line up before they start talking -------------------------
__global__ void rank_barrier_kernel(void* b0, void* b1, void* b2, void* b3, void* b4, void* b5, void* b6, void* b7,
                                    int n_ranks, int rank, unsigned long long seq) {
    void* boxes[kMaxRanks] = {b0, b1, b2, b3, b4, b5, b6, b7};
    const int q = threadIdx.x;
    if (q >= n_ranks) return;
    unsigned long long* theirs = static_cast<unsigned long long*>(boxes[q]) + rank;  // my flag in rank q's box
    const unsigned long long* mine = static_cast<const unsigned long long*>(boxes[rank]) + q;
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(theirs), "l"(seq) : "memory");
    unsigned long long seen = 0;
    for (uint32_t spins = 0;; ++spins) {
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
        if (seen >= seq) break;
        if (spins > (1u << 26)) __trap();  // a missing peer must not hang the GPU box
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
}

cudaError_t launch_rank_barrier(void* const* rankbox, int n_ranks, int rank, uint64_t seq, cudaStream_t s) {
    void* b[kMaxRanks];
    for (int i = 0; i < kMaxRanks; ++i) b[i] = i < n_ranks ? rankbox[i] : nullptr;
    rank_barrier_kernel<<<1, 32, 0, s>>>(b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7], n_ranks, rank,
                                         static_cast<unsigned long long>(seq));
    return cudaGetLastError();
}

namespace {
template <bool kReplica, bool kMulti>
cudaError_t prepare(size_t smem, int* per_sm) {
    cudaError_t e = cudaFuncSetAttribute(merge_loop_kernel<kReplica, kMulti>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, merge_loop_kernel<kReplica, kMulti>, kT, smem);
}
template <bool kReplica, bool kMulti>
cudaError_t launch(void** args, int grid, size_t smem, cudaStream_t s) {
    cudaError_t e = cudaFuncSetAttribute(merge_loop_kernel<kReplica, kMulti>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    // cooperative launch: guarantees that all blocks are resident at once (they wait on one another)
    return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(merge_loop_kernel<kReplica, kMulti>), dim3(grid), dim3(kT),
                                       args, smem, s);
}
}  // namespace

cudaError_t merge_loop_grid(int num_sms, int64_t n, int n_ranks, int n_local, int want_blocks, bool rep,
                            int* blocks_per_rank) {
    *blocks_per_rank = 0;
    if (n_ranks < 1 || n_ranks > kMaxRanks || n_local < 1 || n_local > n_ranks) return cudaErrorInvalidValue;
    const int64_t C = merge_loop_rows_per_rank(n, n_ranks);
    // the exchange costs grow with the number of blocks: small problems use fewer
    // (measured on B200: 48 blocks at n = 20k, 111 at n = 100k; more blocks make the all-to-all dearer than the
    // per-block work they save)
    int64_t G = want_blocks;
    if (G <= 0) {
        // ~400 slots of the update pass per block, and scan windows (n / G columns of a rescanned row, whatever the
        // number of ranks) of at most ~900 columns
        G = std::max<int64_t>((C + 399) / 400, (n + 899) / 900);
        if (G < 8) G = std::min<int64_t>(8, (C + 63) / 64);
        if (G > 111) G = 111;
    }
    const int64_t cap = num_sms / n_local;  // one CTA per SM, all resident
    if (G > cap) G = cap;
    if (G > kMaxBlocks) G = kMaxBlocks;
    if (G < 1) G = 1;
    for (;;) {  // the slice state must fit one SM's shared memory
        int per_sm = 0;
        const size_t smem = merge_loop_smem_bytes(n, n_ranks, static_cast<int>(G), rep);
        const bool multi = n_ranks > 1;
        cudaError_t e = rep ? (multi ? prepare<true, true>(smem, &per_sm) : prepare<true, false>(smem, &per_sm))
                            : (multi ? prepare<false, true>(smem, &per_sm) : prepare<false, false>(smem, &per_sm));
        if (e == cudaSuccess && per_sm > 0) break;
        if (e != cudaSuccess && e != cudaErrorInvalidValue) return e;
        (void)cudaGetLastError();
        if (G >= cap || G >= kMaxBlocks) return cudaSuccess;  // does not fit: *blocks_per_rank stays 0
        G = G * 2 < cap ? G * 2 : cap;
    }
    *blocks_per_rank = static_cast<int>(G);
    return cudaSuccess;
}

cudaError_t launch_merge_loop(const LoopState& st, const LoopParams& p, int blocks_per_rank, bool rep, cudaStream_t s) {
    if (blocks_per_rank <= 0 || blocks_per_rank > kMaxBlocks) return cudaErrorInvalidConfiguration;
    LoopState st_copy = st;
    LoopParams p_copy = p;
    void* args[] = {&st_copy, &p_copy};
    const size_t smem = merge_loop_smem_bytes(st.n, st.n_ranks, blocks_per_rank, rep);
    const bool multi = st.n_ranks > 1;
    const int grid = blocks_per_rank * st.n_local;
    if (rep) return multi ? launch<true, true>(args, grid, smem, s) : launch<true, false>(args, grid, smem, s);
    return multi ? launch<false, true>(args, grid, smem, s) : launch<false, false>(args, grid, smem, s);
}

}  // namespace ic
