// merge_batch.cu -- K3b: the agglomeration loop with BATCHES of provably consecutive merges (one GPU).
//
// Same reference semantics as merge_loop.cu (clustering.go:220-246; FindClosestClusters :119-133, maxSize branch
// :228-234, MergeClusters :29-47, UpdateDistanceMatrix :76-96), same state (slots, keys, static sorted partner
// lists, eager admissibility, Lance-Williams in double) and the SAME merge sequence, but one iteration of this kernel
// takes every merge that is certain to be the next one, the one after it, ... of the sequential algorithm:
//
//   Every live row r has a head h_r = (d, key_r) -- its smallest pair with a lower-key partner -- and the pairs
//   of the matrix in scan order are the k-way merge of the rows' sorted lists.  Ward's update is reducible:
//   d(k, a u b) >= min(d(k,a), d(k,b)), so a merge only ever creates distances that are >= every pair that
//   touched one of its two clusters.  Walk the pairs in scan order and stop at the first pair that touches a
//   cluster of an earlier pair (the "stopper" T).  Every pair before T is disjoint from the others, is a row head,
//   and every distance the earlier merges create is >= T (a tie goes to the older pair: new clusters carry the
//   highest keys).  Hence the sequential algorithm performs exactly those merges, in that order.
//   T = min( second list entry of any row,  any head that is not the first head at both of its slots ).
//
//   Measured batch sizes: 10.2 merges per iteration at config C (N = 100 000), 12.4 at B, 60 at E (maxSize 8): 9 511
//   iterations instead of 97 250.  The stopper is almost always a second list entry, so batches do not grow with N.
//
// One iteration = three grid-wide phases of a persistent cooperative kernel (one CTA per SM):
//   P1 rescans    rows whose cached partners all died (and the rows of the clusters just created) are scanned
//                 by 2048-column windows, one warp each; the warp that finishes a row's last window folds the
//                 partial lists (exact cut rule).  Short rows (n <= 32k): one block per row, fold in shared memory
//   P2 heads      per row: head + stopper; every block publishes its heads below its own stopper minimum
//                 (any value >= T is a valid filter) and that minimum
//   P3 batch      every block reads all published candidates (a few dozen), derives T -- stopper minimum and
//                 slot conflicts, pairwise in shared memory -- and ranks the pairs below T; then: trace / slot
//                 bookkeeping, validation of the partner lists, Lance-Williams rows (the new cluster's row is written
//                 in full; a pair lives in the row of its HIGHER-key cluster, entries of newer clusters are gathered
//                 with predicated loads), the m x m cross terms of the batch (two chained updates from the four old
//                 entries)
// Sharded runs (template parameter kMulti, one process per GPU): every rank runs the kernel on its row block -- own rows
// in P1 / P2 / validation, the columns of its own row block in the rows phase (rows a, b: coalesced peer loads; gathers:
// local; its slice of the new row: coalesced peer store) -- candidates and rank minima are pushed into every rank's
// exchange box, and two cross-rank barriers per iteration (after P2, after the rows) replace the grid barriers.
// HBM roofline: algorithmic bytes = 12*n per merge.
#include <algorithm>

#include "common.cuh"
#include "exact.cuh"
#include "kernels.h"
#include "loop_common.cuh"

namespace ic {

namespace {

constexpr int kBT = kBatchThreads;
constexpr int kBW = kBT / 32;
constexpr uint32_t kMoreBit = 1u, kDryBit = 2u;
constexpr uint32_t kBarSpin = 1u << 24;
#ifndef IC_UPD_COLS
#define IC_UPD_COLS 256
#endif
#ifndef IC_UPD_GROUP
#define IC_UPD_GROUP 1
#endif
constexpr int kUpdGroup = IC_UPD_GROUP;  // merges of one update unit (they share the unit's slot-table chunk)
constexpr int kUpdCols = IC_UPD_COLS;  // columns of one update unit (one warp: 4 x 32 lanes x 4, all loads in flight at once)
constexpr int kRcpTab = 1024;
// exact phase: 1 = register-staged chunks (one in flight; A/B against the cp.async ring of exact.cuh)
#ifndef IC_EXACT_SYNC
#define IC_EXACT_SYNC 0
#endif
// Exact phase: an iteration has about one group of four pairs per 2.5 warps (config C), and what bounds the phase is how
// many bytes of centroid rows are in flight -- so the first kExWarps warps of a block take the groups, each with a ring of
// kExStages chunks (2 640 bytes each) in flight, instead of all sixteen with a shallow one.
#ifndef IC_PROF_EXACT  // debug: the selection phase's profile slots carry block 0 / warp 0's exact-phase group timing instead
#define IC_PROF_EXACT 0
#endif
#ifndef IC_EX_STAGES
#define IC_EX_STAGES 6
#endif
#ifndef IC_EX_WARPS
#define IC_EX_WARPS 8
#endif
constexpr int kExStages = IC_EX_STAGES, kExWarps = IC_EX_WARPS;
constexpr int kExRing = kExStages * kExAsyncStage > kExGroup * kExStride ? kExStages * kExAsyncStage : kExGroup * kExStride;  // floats per ring
constexpr int kExSmemFloats = kExWarps * kExRing + kBW * kExChunk;  // + one register-staged chunk per warp (single pairs)
static_assert(kExRing % 4 == 0 && kExWarps >= 1 && kExWarps <= kBW, "exact phase staging memory");
constexpr int kVR = 1;         // rows per thread whose lists are loaded ahead of the update pass

// counters[slot][*]
enum { CN_DRY = 0, CN_CAND = 1, CN_XQ = 2, CN_NEAR = 3 };

// Every lane reserves `cnt` consecutive entries of a global queue with one atomic per warp; returns the lane's first index.
IC_DEVINL int32_t warp_reserve(int32_t* counter, int cnt, int lane) {
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    const int total = __shfl_sync(0xffffffffu, inc, 31);
    int32_t base = 0;
    if (lane == 31 && total > 0) base = atomicAdd(counter, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + inc - cnt;
}

// A barrier that is not completed after kBarSpin polls (seconds: a protocol bug, or a peer that died) must neither hang
// the GPU nor poison the CUDA context (__trap() does: every later call of the process fails).  The block that gives up
// raises an abort word next to the barrier counter and reports CTL_ERROR = 4; every poll loop looks at the word now and
// then; a block that sees it leaves the kernel (all its threads, right after the barrier's __syncthreads).  The host finds
// CTL_ERROR / no CTL_DONE and returns IC_ERR_INTERNAL; the context stays usable.
constexpr int kBarAbortWord = 40;  // (the barrier scratch is 64 words, zeroed by the host before every launch)
IC_DEVINL bool bar_give_up(uint32_t* bar, int32_t* ctl, uint32_t spins) {
    if ((spins & 0x3FFu) != 0u) return false;
    if (spins <= kBarSpin && ld_acquire_u32(bar + kBarAbortWord) == 0u) return false;
    atomicExch(bar + kBarAbortWord, 1u);
    atomicExch(ctl + CTL_ERROR, 4);
    return true;
}
#define IC_BAR_EXIT_IF(flag)                  \
    do {                                      \
        if (flag) asm volatile("exit;");      \
    } while (0)

// grid-wide barrier on one monotone counter: arrive with a release reduction, poll with acquire loads
IC_DEVINL void grid_sync(uint32_t* bar, int32_t* ctl, uint32_t& phase, uint32_t G, long long* wait_acc = nullptr) {
    __shared__ int s_abort;
    __syncthreads();
    ++phase;
    if (threadIdx.x == 0) {
        const long long t_arrive = wait_acc ? clock64() : 0;
        s_abort = 0;
        red_release_add_u32(bar, 1u);
        const uint32_t target = phase * G;
        uint32_t spins = 0;
        while (ld_acquire_u32(bar) < target)
            if (bar_give_up(bar, ctl, ++spins)) {
                s_abort = 1;
                break;
            }
        if (wait_acc) *wait_acc += clock64() - t_arrive;
    }
    __syncthreads();
    IC_BAR_EXIT_IF(s_abort);
}

// Barrier of all blocks of ALL ranks (sharded runs): every block fences its stores to peer memory at system scope and
// arrives on the local counter; block 0 waits for its rank, exchanges a sequence number with every peer through the
// peer-mapped flags, and releases the local blocks.
// `remote_stores`: the blocks wrote to peer memory in this phase (their stores must be performed at system scope before
// the rank reports; stores to the rank's own memory only need the gpu-scope release of the arrival -- block 0's system
// fence after it has acquired all arrivals is cumulative).
IC_DEVINL void grid_sync_ranks(const BatchState& st, uint32_t& phase, uint32_t& xcount, uint32_t G, uint32_t bid, bool remote_stores,
                               int publish_slot, long long* wait_acc = nullptr) {
    __shared__ int s_abort_r;
    __syncthreads();
    ++phase;
    ++xcount;
    if (threadIdx.x == 0) {
        const long long t_arrive = wait_acc ? clock64() : 0;
        s_abort_r = 0;
        if (remote_stores) asm volatile("fence.acq_rel.sys;" ::: "memory");
        red_release_add_u32(st.bar, 1u);
        uint32_t spins = 0;
        if (bid == 0) {
            const uint32_t target = phase * G;
            while (ld_acquire_u32(st.bar) < target)
                if (bar_give_up(st.bar, st.ctl, ++spins)) {
                    s_abort_r = 1;
                    break;
                }
            const unsigned long long seq = (static_cast<unsigned long long>(st.gen) << 32) | xcount;
            if (publish_slot >= 0) {  // this rank's minima and candidate count of the iteration go into every rank's box
                const uint8_t* acc = st.xbox[st.rank] + kBatchXAccum;
                const unsigned long long sstop = __ldcg(reinterpret_cast<const unsigned long long*>(acc) + publish_slot);
                const unsigned long long shead = __ldcg(reinterpret_cast<const unsigned long long*>(acc) + 3 + publish_slot);
                const int32_t scnt = __ldcg(reinterpret_cast<const int32_t*>(acc + 48) + publish_slot);
                if (scnt > kBatchXCand) {
                    // more candidates than this rank's region holds (which ones were dropped depends on the atomics'
                    // order): publish the rank's smallest head as well -- every rank then merges the global minimum only
                    uint64_t best = kPackInf;
                    uint32_t bb = 0;
                    for (uint32_t g = 0; g < G; ++g) {
                        const uint4 r0 = __ldcg(st.blockmin + 2 * g);
                        const uint64_t hp = (static_cast<uint64_t>(r0.y) << 32) | r0.x;
                        if (hp < best) {
                            best = hp;
                            bb = g;
                        }
                    }
                    const uint4 m0 = __ldcg(st.blockmin + 2 * bb), m1 = __ldcg(st.blockmin + 2 * bb + 1);
                    for (int q = 0; q < st.n_ranks; ++q) {
                        uint4* dst = reinterpret_cast<uint4*>(st.xbox[q] + kBatchXRankMin + (static_cast<size_t>(st.rank) * 3 + publish_slot) * 32);
                        __stcg(dst, m0);
                        __stcg(dst + 1, m1);
                    }
                }
                for (int q = 0; q < st.n_ranks; ++q) {
                    uint4* dst = reinterpret_cast<uint4*>(st.xbox[q] + kBatchXSummary + (static_cast<size_t>(st.rank) * 3 + publish_slot) * 32);
                    __stcg(dst, make_uint4(static_cast<uint32_t>(sstop), static_cast<uint32_t>(sstop >> 32),
                                           static_cast<uint32_t>(shead), static_cast<uint32_t>(shead >> 32)));
                    __stcg(dst + 1, make_uint4(static_cast<uint32_t>(scnt), static_cast<uint32_t>((__ldcg(st.ctl + CTL_XQ_OVERFLOW) != 0 ? 1 : 0) | (__ldcg(st.ctl + CTL_ORDER_VIOL) != 0 ? 2 : 0)), 0u, 0u));
                }
            }
            // block 0's own pushes (the summary) must be performed before the flags; the other blocks fenced theirs
            if (publish_slot >= 0 || remote_stores) asm volatile("fence.acq_rel.sys;" ::: "memory");
            for (int q = 0; q < st.n_ranks; ++q)
                if (q != st.rank)
                    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(reinterpret_cast<unsigned long long*>(st.xbox[q]) + st.rank), "l"(seq) : "memory");
            for (int q = 0; q < st.n_ranks; ++q) {
                if (q == st.rank) continue;
                unsigned long long seen = 0;
                for (spins = 1; s_abort_r == 0; ++spins) {
                    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(reinterpret_cast<unsigned long long*>(st.xbox[st.rank]) + q) : "memory");
                    if (seen >= seq) break;
                    if (bar_give_up(st.bar, st.ctl, spins)) s_abort_r = 1;  // a missing peer must not hang the GPU box
                }
            }
            // acquire at system scope (the flags were polled with relaxed loads), then release the rank's blocks at gpu scope:
            // cumulativity makes what the peers pushed before their flags visible to every block of this rank.  ~3 400
            // cycles per barrier (profiles/r01_nvlink_pingpong.txt); with dozens of merges per iteration that is noise
            asm volatile("fence.acq_rel.sys;" ::: "memory");
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(st.bar + 32), "r"(phase) : "memory");
        } else {
            while (ld_acquire_u32(st.bar + 32) < phase)
                if (bar_give_up(st.bar, st.ctl, ++spins)) {
                    s_abort_r = 1;
                    break;
                }
        }
        if (wait_acc) *wait_acc += clock64() - t_arrive;
    }
    __syncthreads();
    IC_BAR_EXIT_IF(s_abort_r);
}

IC_DEVINL uint64_t block_min_u64(uint64_t v, uint64_t* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t w = warp_min_u64(v);
    __syncthreads();
    if (lane == 0) s_red[warp] = w;
    __syncthreads();
    uint64_t r = s_red[0];
#pragma unroll
    for (int i = 1; i < kBW; ++i) r = umin64(r, s_red[i]);
    return r;
}
IC_DEVINL int block_sum_i32(int v, int* s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    int r = 0;
#pragma unroll
    for (int i = 0; i < kBW; ++i) r += s_red[i];
    return r;
}

// predicated 4-byte load (no branch: the sixteen elements of an update unit stay independent instruction streams)
IC_DEVINL float ldcg_if(const float* p, bool pred, float other) {
    float v = other;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q ld.global.cg.f32 %0, [%1];\n\t}" : "+f"(v) : "l"(p), "r"(static_cast<int>(pred)));
    return v;
}

// head and stopper of a row from its list.  Lists are compacted: valid entries first, then (if entries were
// removed and more partners may exist) one bound placeholder, then empty entries.  The distance of the last
// non-empty entry is a lower bound of every unlisted partner.
struct RowHead {
    uint64_t head;  // (dist bits << 32 | row key), kPackInf: none
    uint64_t stop;  // smallest pack a second pair of this row may have, +1 (see P2); kPackInf: none
    uint64_t bound; // the row has no listed partner but may have partners at or above this pack (beyond the horizon)
    uint32_t partner_slot, partner_key;
    int32_t partner_size;
};
IC_DEVINL RowHead row_head(uint4 e0, uint4 e1, uint32_t more_bits, uint32_t key_r) {
    RowHead h;
    h.head = h.stop = h.bound = kPackInf;
    h.partner_slot = h.partner_key = kNoPartner;
    h.partner_size = 0;
    if (e0.z == kNoPartner) {  // no listed partner; a bound here comes from a near list whose entries all died (near.cu):
        if (e0.y != kNoPartner) h.bound = (static_cast<uint64_t>(e0.y) << 32) | key_r;  // the row's pairs are beyond the horizon
        return h;
    }
    h.head = (static_cast<uint64_t>(e0.y) << 32) | key_r;
    h.partner_slot = e0.z;
    h.partner_key = e0.x;
    h.partner_size = static_cast<int32_t>(e0.w);
    if (e1.y != kNoPartner)  // a second partner, or the bound of the unlisted ones
        h.stop = ((static_cast<uint64_t>(e1.y) << 32) | key_r) + 1ull;
    else if ((more_bits & kMoreBit) != 0u)
        h.stop = h.head + 1ull;  // list cut after the head: the rest is >= the head's distance
    return h;
}

}  // namespace

size_t merge_batch_smem_bytes(int64_t n) {
    const size_t n4 = static_cast<size_t>((n + 3) / 4 * 4);
    const size_t bitmap = ((n4 + 31) / 32 + 3) / 4 * 4 * sizeof(uint32_t);
    return bitmap + sizeof(float) * kExSmemFloats;  // + the exact phase's rings and staging buffers
}
static int64_t batch_window_cols(int64_t n) {  // at most kBatchMaxWin windows per row: <= 4 partial lists per lane in the fold
    const int64_t n4 = (n + 3) / 4 * 4;
    return std::max<int64_t>(kBatchWinMin, ((n4 + kBatchMaxWin - 1) / kBatchMaxWin + 127) / 128 * 128);
}
int64_t merge_batch_windows(int64_t n) {
    const int64_t n4 = (n + 3) / 4 * 4, win = batch_window_cols(n);
    return std::max<int64_t>(1, (n4 + win - 1) / win);
}

// kVirt (test hook, option "virtual_ranks"): the ranks of a sharded run are emulated on ONE GPU by one cooperative launch --
// blocks [v * G, (v + 1) * G) are rank v and take their BatchState from vstates[v]; the peers' rows and exchange boxes are
// simply other addresses of the same device.  Same code path as one process per GPU (kMulti), so pytest -m gpu covers it.
template <bool kMulti, bool kVirt>
__global__ void __launch_bounds__(kBT, 1)
merge_batch_kernel(const __grid_constant__ BatchState st_param, const __grid_constant__ LoopParams prm,
                   const BatchState* __restrict__ vstates, int32_t blocks_per_rank) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ BatchState s_vst;
    if (kVirt) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(vstates + blockIdx.x / blocks_per_rank);
        uint32_t* dst = reinterpret_cast<uint32_t*>(&s_vst);
        for (int i = tid; i < static_cast<int>(sizeof(BatchState) / 4); i += kBT) dst[i] = src[i];
        __syncthreads();
    }
    const BatchState& st = kVirt ? s_vst : st_param;
    const uint32_t bid = kVirt ? blockIdx.x % static_cast<uint32_t>(blocks_per_rank) : blockIdx.x;
    const uint32_t G = kVirt ? static_cast<uint32_t>(blocks_per_rank) : gridDim.x;
    const int32_t n = st.n;
    const int32_t n4 = (n + 3) & ~3;
    const int32_t kbase = st.key_base;
    const int64_t ld = st.ld;
    const int32_t gtid = static_cast<int32_t>(bid) * kBT + tid, GT = static_cast<int32_t>(G) * kBT;
    // work units go to warps block-interleaved: consecutive units run on different SMs
    const int32_t gw = warp * static_cast<int32_t>(G) + static_cast<int32_t>(bid), GW = static_cast<int32_t>(G) * kBW;
    const int32_t win = st.win_cols, nwin = st.n_win;
    float* const dm = st.dm;  // first row of this rank's row block (row 0 on one GPU)
    int32_t* const ctl = st.ctl;
    // row-block shard of this rank (sharded runs): rows [r_lo, r_hi) live here, with all their columns
    const int32_t C = kMulti ? st.rows_per_rank : n4;
    const int32_t r_lo = kMulti ? min(n, st.rank * C) : 0, r_hi = kMulti ? min(n, r_lo + C) : n;
    auto row_of = [&](int32_t slot) -> float* {  // any cluster's row (peer mapped when it lives on another rank)
        if (!kMulti) return dm + static_cast<int64_t>(slot) * ld;
        const int32_t q = slot / C;
        return st.dm_rank[q] + static_cast<int64_t>(slot - q * C) * ld;
    };
    uint8_t* const xb = kMulti ? st.xbox[st.rank] : nullptr;  // this rank's exchange box
    unsigned long long* const xb_stop = reinterpret_cast<unsigned long long*>(xb + kBatchXAccum);
    unsigned long long* const xb_head = xb_stop + 3;
    int32_t* const xb_cnt = reinterpret_cast<int32_t*>(xb_head + 3);
    uint32_t xcount = 0;
    __shared__ int32_t s_xcnt[kMaxRanks + 1];  // candidate pairs per rank (prefix sums)
    __shared__ int32_t s_xover[kMaxRanks];     // rank dropped candidates (its region of the exchange box was full)
    __shared__ int32_t s_xcntm[kMaxBatch], s_upre[kMaxBatch + 1], s_wsum[kBW];  // exact phase: queued pairs / groups per merge

    extern __shared__ __align__(16) uint8_t dyn_smem[];
    // exact phase: squared differences of one chunk of up to kExGroup pairs per warp
    float (*const s_ex)[kExRing] = reinterpret_cast<float (*)[kExRing]>(dyn_smem);  // rings of the first kExWarps warps
    float (*const s_one)[kExChunk] = reinterpret_cast<float (*)[kExChunk]>(dyn_smem + sizeof(float) * kExWarps * kExRing);
    uint32_t* const s_bits = reinterpret_cast<uint32_t*>(dyn_smem + sizeof(float) * kExSmemFloats);  // merged-slot bitmap of the current batch
    const int32_t n_words = (n4 + 31) >> 5;

    __shared__ uint64_t s_red[kBW];
    __shared__ int s_redi[kBW];
    // candidate pairs as collected (unordered) ...
    __shared__ uint64_t s_hp[kMaxBatch];
    __shared__ int32_t s_ca[kMaxBatch], s_cb[kMaxBatch], s_csa[kMaxBatch], s_csb[kMaxBatch], s_ckb[kMaxBatch];
    // ... and the batch in scan order
    __shared__ int32_t s_a[kMaxBatch], s_b[kMaxBatch], s_sa[kMaxBatch], s_sb[kMaxBatch], s_ka[kMaxBatch], s_kb[kMaxBatch];
    __shared__ uint32_t s_d[kMaxBatch + 1];
    __shared__ int32_t s_m;
    __shared__ uint64_t s_ppk[kBW][kNNK];  // row-per-block rescans: the warps' partial lists
    __shared__ int32_t s_psl[kBW][kNNK];
    __shared__ int2 s_pm[kBW];
    __shared__ double s_rcp[kRcpTab];  // 1.0 / size sum, correctly rounded (what lance_williams() computes inline)
    const bool exact = prm.exact != 0;
    const bool use_xres = exact && !kMulti;
    // what a list says about partners that are not in a near list / not re-evaluated: they are ABOVE the horizon, i.e. at
    // least the next float (a bound AT the horizon could equal the safe bound when the horizon is 0: duplicates)
    const uint32_t hz_bound = prm.horizon >= 0.0 ? __float_as_uint(static_cast<float>(prm.horizon)) + 1u : 0u;
    const int d4 = static_cast<int>(st.ldc);

    uint32_t phase = 0;
    int32_t m_prev = 0;  // merges of the previous iteration
    int32_t n_live = __ldcg(ctl + CTL_N_LIVE);
    int32_t t = __ldcg(ctl + CTL_N_MERGES);
    int32_t launched = 0, stop_reason = 0, iters = 0;
    long long n_rescans = 0;
    const bool timed = st.prof != nullptr && bid == 0 && tid == 0;
    const bool small_sizes = prm.max_size < kRcpTab;  // every admissible size sum has its reciprocal in the table
    __shared__ long long c_ph[10];  // cycles of block 0 per phase (profile_loop)
    if (tid < 10) c_ph[tid] = 0;
    __shared__ long long c_sel[4];  // selection phase: candidates loaded + minima, theta, filter, (rest: conflicts + ranks); bookkeeping
    if (tid < 4) c_sel[tid] = 0;

    for (int i = tid; i < kRcpTab; i += kBT) s_rcp[i] = 1.0 / static_cast<double>(i > 0 ? i : 1);
    __syncthreads();
    // rows that were dry when the previous launch stopped (or rows the other loop left dry): queue slot 0
    // a row whose partner list has to be rebuilt: re-selected from its near list if it has one, else scanned
    auto queue_row = [&](int32_t r, int32_t key, int slot3) {
        if (st.near_meta != nullptr && __ldcg(&st.near_meta[r].y) >= 0)
            st.nearq[atomicAdd(st.counters + slot3 * 4 + CN_NEAR, 1)] = make_int2(r, key);
        else
            st.dryq[atomicAdd(st.counters + slot3 * 4 + CN_DRY, 1)] = make_int2(r, key);
    };
    for (int32_t r = r_lo + gtid; r < r_hi; r += GT)
        if ((static_cast<uint32_t>(__ldcg(st.nn_more + r)) & kDryBit) != 0u && __ldcg(st.ks + r).x >= 0)
            queue_row(r, __ldcg(st.ks + r).x, 0);
    // live size of every slot (0: retired / padding): what the update pass streams beside the two rows
    for (int32_t r = gtid; r < n4; r += GT) {
        const int2 k = __ldcg(st.ks + r);
        st.lsize[r] = (r < n && k.x >= 0) ? k.y : 0;
    }
    grid_sync(st.bar, st.ctl, phase, G);

    for (uint32_t it = 0;; ++it) {
        const int sl = static_cast<int>(it % 3u), sl1 = static_cast<int>((it + 1u) % 3u), sl2 = static_cast<int>((it + 2u) % 3u);
        const long long tp0 = timed ? clock64() : 0;

        // ---- partner lists of the clusters the previous iteration created, selected from their re-evaluated pairs ----
        // (one GPU, reference arithmetic: every pair of the new row at or below the horizon went through the exact phase;
        // the rest of the row is above the horizon, which therefore bounds the unlisted partners)
        if (exact && !use_xres)
            for (int32_t j = gtid; j < m_prev; j += GT) st.xhit[j] = 0;  // (read by every block in the exact phase before the barrier)
        if (use_xres) {
            for (int32_t j = gw; j < m_prev; j += GW) {
                const int32_t c = __ldcg(st.xhit + j);
                // the rest of the row is above the horizon (assumed to hold selectable values: if it does not -- everything else
                // masked by maxSize -- the spurious bound costs one more horizon raise at the very end, whose sweep clears it)
                const bool far = true;
                if (lane == 0) st.xhit[j] = 0;
                // too many pairs, or a queue overflowed (the host re-evaluates the rows): the row was queued for a scan
                if (c > kXResCap || __ldcg(ctl + CTL_XQ_OVERFLOW) != 0) continue;
                const uint4* src = st.xres + static_cast<int64_t>(j) * kXResCap;
                {   // the re-evaluated pairs ARE the row's near list (near.cu): everything else in the row is above the horizon
                    int32_t base = 0;
                    if (st.near_meta != nullptr) {
                        if (lane == 0 && c > 0) base = atomicAdd(st.near_cursor, c);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        const bool fits = base + c <= st.near_pool_cap;
                        for (int32_t i = lane; i < c && fits; i += 32) {
                            const uint4 e = __ldcg(src + i);
                            st.near_pool[static_cast<int64_t>(base) + i] = make_uint2(e.x, e.y);
                        }
                        if (lane == 0) st.near_meta[s_b[j]] = fits ? make_int2(base, c | (far ? kNearFarBit : 0)) : make_int2(0, -1);
                    }
                }
                constexpr int kPer = kXResCap / 32;
                uint64_t pk[kPer];
                uint32_t taken = 0u;
#pragma unroll
                for (int x = 0; x < kPer; ++x) {
                    const int32_t idx = x * 32 + lane;
                    pk[x] = kPackInf;
                    if (idx < c) {
                        const uint4 e = __ldcg(src + idx);
                        pk[x] = (static_cast<uint64_t>(e.x) << 32) | e.y;
                    }
                }
                const int32_t b = s_b[j];
                for (int r = 0; r < kNNK; ++r) {
                    uint64_t best = kPackInf;
                    int bx = -1;
#pragma unroll
                    for (int x = 0; x < kPer; ++x)
                        if (((taken >> x) & 1u) == 0u && pk[x] < best) {
                            best = pk[x];
                            bx = x;
                        }
                    const uint64_t wm = warp_min_u64(best);
                    uint4 out = nn_none();
                    if (wm != kPackInf) {
                        const bool win = best == wm;  // packs are unique (distinct partner keys)
                        const int src_lane = __ffs(__ballot_sync(0xffffffffu, win)) - 1;
                        if (win) taken |= 1u << bx;
                        uint4 e = make_uint4(0u, 0u, 0u, 0u);
                        if (win) e = __ldcg(src + bx * 32 + lane);
                        out.x = __shfl_sync(0xffffffffu, e.y, src_lane);  // partner key
                        out.y = __shfl_sync(0xffffffffu, e.x, src_lane);  // distance bits
                        out.z = __shfl_sync(0xffffffffu, e.z, src_lane);  // partner slot
                        out.w = __shfl_sync(0xffffffffu, e.w, src_lane);  // partner size
                    } else if (r == c && far) {
                        out = nn_bound(hz_bound);  // the unlisted partners are above the horizon
                    }
                    if (lane == 0) __stcg(st.nn + static_cast<int64_t>(b) * kNNK + r, out);
                }
                if (lane == 0) __stcg(st.nn_more + b, (c > kNNK || far) ? static_cast<int32_t>(kMoreBit) : 0);
            }
        }

        // One (row, window) unit of a rescan: a warp scans 2048 columns of a queued row; the warp that delivers a row's last
        // window folds the partial lists and writes the row's list.  `during_batch`: the unit runs inside the rows phase of
        // the batch that follows the one the row ran dry in -- clusters of the current batch are excluded through the
        // bitmap (their slots are being rewritten), and a queued row that was itself merged is skipped.
        auto scan_window_unit = [&](int32_t base, int32_t rows, int64_t u, bool during_batch) {
                // windows of one row go to warps of different blocks: q = u % rows
                const int32_t w = static_cast<int32_t>(u / rows), q = static_cast<int32_t>(u - static_cast<int64_t>(w) * rows);
                const int2 rq = __ldcg(st.dryq + base + q);
                const int32_t r = rq.x;
                if (during_batch && ((s_bits[r >> 5] >> (r & 31)) & 1u)) return;  // merged in this batch: no list to rebuild
                const uint32_t ukr = static_cast<uint32_t>(rq.y);
                const float* rowp = dm + static_cast<int64_t>(r - r_lo) * ld;
                // a cluster older than the last compaction sits in key order: its lower-key partners are the columns before it
                const int32_t lim = static_cast<int32_t>(ukr) < st.order_key ? min(n4, (r + 3) & ~3) : n4;
                const int32_t sw0 = min(lim, w * win), sw1 = min(lim, sw0 + win);
                ScanCand c;
                scan_init(c);
                constexpr int kU = 8;  // 16-byte loads of the row and of the keys in flight per lane
                for (int32_t b0 = sw0; b0 < sw1; b0 += 128 * kU) {
                    float4 vv[kU];
                    int4 kq[kU];
#pragma unroll
                    for (int x = 0; x < kU; ++x) {
                        const int32_t u0 = b0 + (x * 32 + lane) * 4;
                        vv[x] = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
                        kq[x] = make_int4(-1, -1, -1, -1);
                        if (u0 < sw1) {
                            vv[x] = __ldcg(reinterpret_cast<const float4*>(rowp + u0));
                            kq[x] = __ldcg(reinterpret_cast<const int4*>(st.gkey + u0));
                        }
                    }
#pragma unroll
                    for (int x = 0; x < kU; ++x) {
                        const int32_t u0 = b0 + (x * 32 + lane) * 4;
                        const uint32_t vs4[4] = {__float_as_uint(vv[x].x), __float_as_uint(vv[x].y), __float_as_uint(vv[x].z),
                                                 __float_as_uint(vv[x].w)};
                        const uint32_t ks4[4] = {static_cast<uint32_t>(kq[x].x), static_cast<uint32_t>(kq[x].y),
                                                 static_cast<uint32_t>(kq[x].z), static_cast<uint32_t>(kq[x].w)};
#pragma unroll
                        for (int e = 0; e < 4; ++e)  // retired slots and padding hold key -1 == 0xFFFFFFFF: never below the row's key
                            if (ks4[e] < ukr && vs4[e] < kMaxFloatBits &&
                                !(during_batch && ((s_bits[(u0 + e) >> 5] >> ((u0 + e) & 31)) & 1u)))
                                scan_insert(c, (static_cast<uint64_t>(vs4[e]) << 32) | ks4[e], u0 + e);
                    }
                }
                PartList out;
                bool more = false;
                out.m = warp_select_scan(c, out.pk, out.sl, more);
                out.more = more ? 1 : 0;
#pragma unroll
                for (int x = 0; x < kNNK; ++x) out.sz[x] = 0;
                bool folder = nwin == 1;
                if (nwin > 1) {
                    uint4* prec = st.partials + (static_cast<size_t>(q) * kBatchMaxWin + w) * 8;
                    if (lane < kNNK) {
                        const uint64_t myp = sel4(out.pk, lane);
                        __stcg(prec + lane, lane < out.m ? make_uint4(pack_key(myp), static_cast<uint32_t>(myp >> 32),
                                                                      static_cast<uint32_t>(sel4(out.sl, lane)), 0u)
                                                         : nn_none());
                    } else if (lane == kNNK) {
                        __stcg(prec + kNNK, make_uint4(static_cast<uint32_t>(out.m), static_cast<uint32_t>(out.more), 0u, 0u));
                    }
                    __syncwarp();
                    int old = 0;
                    if (lane == 0) {
                        __threadfence();
                        old = atomicAdd(st.part_cnt + q, 1);
                        __threadfence();
                    }
                    old = __shfl_sync(0xffffffffu, old, 0);
                    folder = old == nwin - 1;
                    if (folder) {  // last window of the row: fold the nwin partial lists (up to four per lane)
                        PartList acc;
                        acc.m = 0;
                        acc.more = 0;
#pragma unroll
                        for (int x = 0; x < kNNK; ++x) {
                            acc.pk[x] = kPackInf;
                            acc.sl[x] = -1;
                            acc.sz[x] = 0;
                        }
                        for (int32_t ww = lane; ww < nwin; ww += 32) {
                            const uint4* rec = st.partials + (static_cast<size_t>(q) * kBatchMaxWin + ww) * 8;
                            PartList in;
                            const uint4 hd = __ldcg(rec + kNNK);
                            in.m = static_cast<int32_t>(hd.x);
                            in.more = static_cast<int32_t>(hd.y);
#pragma unroll
                            for (int x = 0; x < kNNK; ++x) {
                                const uint4 e = __ldcg(rec + x);
                                in.pk[x] = x < in.m ? ((static_cast<uint64_t>(e.y) << 32) | e.x) : kPackInf;
                                in.sl[x] = x < in.m ? static_cast<int32_t>(e.z) : -1;
                                in.sz[x] = 0;
                            }
                            if (ww == lane)
                                acc = in;
                            else
                                lane_merge2(acc, in);
                        }
                        warp_merge_lists(acc, out);
                        if (lane == 0) st.part_cnt[q] = 0;
                    }
                }
                if (folder) {
                    if (lane < kNNK) {  // entry: {partner key, distance bits, partner slot, partner size}
                        const uint64_t myp = sel4(out.pk, lane);
                        const int32_t ps = sel4(out.sl, lane);
                        const int32_t psz = lane < out.m ? __ldcg(st.lsize + ps) : 0;
                        __stcg(st.nn + static_cast<int64_t>(r) * kNNK + lane,
                               lane < out.m ? make_uint4(pack_key(myp), static_cast<uint32_t>(myp >> 32),
                                                         static_cast<uint32_t>(ps), static_cast<uint32_t>(psz))
                                            : nn_none());
                    } else if (lane == kNNK) {
                        __stcg(st.nn_more + r, out.more ? static_cast<int32_t>(kMoreBit) : 0);
                    }
                }
        };
        // ================= P1: row rescans =================
        // rows with a near list: one warp re-selects the 4 smallest LIVE entries (liveness: slot_of_key); the unlisted
        // partners are the remaining near entries and, if the row has any, everything beyond the horizon
        const int32_t Qn = __ldcg(st.counters + sl * 4 + CN_NEAR);
        for (int32_t q = gw; q < Qn; q += GW) {
            const int32_t r = __ldcg(st.nearq + q).x;
            const int2 meta = __ldcg(st.near_meta + r);
            const int32_t cnt = meta.y & kNearCntMask;
            const bool far = (meta.y & kNearFarBit) != 0;
            const uint2* seg = st.near_pool + meta.x;
            uint64_t t0 = kPackInf, t1 = kPackInf, t2 = kPackInf, t3 = kPackInf;  // this lane's smallest live entries, ascending
            int32_t alive = 0;
            for (int32_t i = lane; i < cnt; i += 32) {
                const uint2 e = __ldcg(seg + i);
                if (__ldcg(st.slot_of_key + e.y) < 0) continue;
                ++alive;
                uint64_t p = (static_cast<uint64_t>(e.x) << 32) | e.y;
                if (p < t0) { const uint64_t x = t0; t0 = p; p = x; }
                if (p < t1) { const uint64_t x = t1; t1 = p; p = x; }
                if (p < t2) { const uint64_t x = t2; t2 = p; p = x; }
                if (p < t3) t3 = p;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) alive += __shfl_xor_sync(0xffffffffu, alive, o);
            int n_out = 0;
            for (int k = 0; k < kNNK; ++k) {
                const uint64_t wm = warp_min_u64(t0);
                uint4 out = nn_none();
                if (wm != kPackInf) {
                    if (t0 == wm) {  // packs are unique (distinct partner keys): one lane
                        t0 = t1;
                        t1 = t2;
                        t2 = t3;
                        t3 = kPackInf;
                    }
                    const uint32_t pkey = pack_key(wm);
                    const int32_t ps = __ldcg(st.slot_of_key + pkey);
                    out = make_uint4(pkey, static_cast<uint32_t>(wm >> 32), static_cast<uint32_t>(ps), static_cast<uint32_t>(__ldcg(st.lsize + ps)));
                    n_out = k + 1;
                } else if (k == n_out && far) {
                    out = nn_bound(hz_bound);
                }
                if (lane == 0) __stcg(st.nn + static_cast<int64_t>(r) * kNNK + k, out);
            }
            if (lane == 0) __stcg(st.nn_more + r, (alive > kNNK || far) ? static_cast<int32_t>(kMoreBit) : 0);
        }
        const int32_t Q = __ldcg(st.counters + sl * 4 + CN_DRY);
        n_rescans += Q + Qn;
        // Short rows, a few dozen of them: ONE BLOCK PER ROW.  Its sixteen warps scan a sixteenth of the row each and the
        // partial lists are folded in shared memory -- no partial records in global memory, no fence / atomic / re-read
        // chain.  Long rows must be spread over many SMs: one SM streams a 400 KB row (+ its keys) in ~50 k cycles
        // (measured at config C; the window mode below takes 32 k per iteration for ~28 rows).
        const bool row_per_block = Q <= 2 * static_cast<int32_t>(G) && n4 <= 32768;
        for (int32_t q = static_cast<int32_t>(bid); row_per_block && q < Q; q += static_cast<int32_t>(G)) {
            const int2 rq = __ldcg(st.dryq + q);
            const int32_t r = rq.x;
            const uint32_t ukr = static_cast<uint32_t>(rq.y);
            const float* rowp = dm + static_cast<int64_t>(r - r_lo) * ld;
            const int32_t lim = static_cast<int32_t>(ukr) < st.order_key ? min(n4, (r + 3) & ~3) : n4;
            const int32_t seg = (((lim + kBW - 1) / kBW) + 127) & ~127;
            const int32_t sw0 = min(lim, warp * seg), sw1 = min(lim, sw0 + seg);
            ScanCand c;
            scan_init(c);
            constexpr int kU = 8;
            for (int32_t b0 = sw0; b0 < sw1; b0 += 128 * kU) {
                float4 vv[kU];
                int4 kq[kU];
#pragma unroll
                for (int x = 0; x < kU; ++x) {
                    const int32_t u0 = b0 + (x * 32 + lane) * 4;
                    vv[x] = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
                    kq[x] = make_int4(-1, -1, -1, -1);
                    if (u0 < sw1) {
                        vv[x] = __ldcg(reinterpret_cast<const float4*>(rowp + u0));
                        kq[x] = __ldcg(reinterpret_cast<const int4*>(st.gkey + u0));
                    }
                }
#pragma unroll
                for (int x = 0; x < kU; ++x) {
                    const int32_t u0 = b0 + (x * 32 + lane) * 4;
                    const uint32_t vs4[4] = {__float_as_uint(vv[x].x), __float_as_uint(vv[x].y), __float_as_uint(vv[x].z),
                                             __float_as_uint(vv[x].w)};
                    const uint32_t ks4[4] = {static_cast<uint32_t>(kq[x].x), static_cast<uint32_t>(kq[x].y),
                                             static_cast<uint32_t>(kq[x].z), static_cast<uint32_t>(kq[x].w)};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if (ks4[e] < ukr && vs4[e] < kMaxFloatBits)
                            scan_insert(c, (static_cast<uint64_t>(vs4[e]) << 32) | ks4[e], u0 + e);
                }
            }
            PartList out;
            bool more = false;
            out.m = warp_select_scan(c, out.pk, out.sl, more);
            out.more = more ? 1 : 0;
            if (lane < kNNK) {
                s_ppk[warp][lane] = sel4(out.pk, lane);
                s_psl[warp][lane] = sel4(out.sl, lane);
            } else if (lane == kNNK) {
                s_pm[warp] = make_int2(out.m, out.more);
            }
            __syncthreads();
            if (warp == 0) {
                PartList in;
                in.m = 0;
                in.more = 0;
#pragma unroll
                for (int x = 0; x < kNNK; ++x) {
                    in.pk[x] = kPackInf;
                    in.sl[x] = -1;
                    in.sz[x] = 0;
                }
                if (lane < kBW) {
                    in.m = s_pm[lane].x;
                    in.more = s_pm[lane].y;
#pragma unroll
                    for (int x = 0; x < kNNK; ++x) {
                        in.pk[x] = x < in.m ? s_ppk[lane][x] : kPackInf;
                        in.sl[x] = x < in.m ? s_psl[lane][x] : -1;
                    }
                }
                warp_merge_lists(in, out);
                if (lane < kNNK) {  // entry: {partner key, distance bits, partner slot, partner size}
                    const uint64_t myp = sel4(out.pk, lane);
                    const int32_t ps = sel4(out.sl, lane);
                    const int32_t psz = lane < out.m ? __ldcg(st.lsize + ps) : 0;
                    __stcg(st.nn + static_cast<int64_t>(r) * kNNK + lane,
                           lane < out.m ? make_uint4(pack_key(myp), static_cast<uint32_t>(myp >> 32), static_cast<uint32_t>(ps),
                                                     static_cast<uint32_t>(psz))
                                        : nn_none());
                } else if (lane == kNNK) {
                    __stcg(st.nn_more + r, out.more ? static_cast<int32_t>(kMoreBit) : 0);
                }
            }
            __syncthreads();
        }
        // Many rows at once (thousands of duplicates, ...): cooperative, one warp per (row, window)
        for (int32_t base = 0; !row_per_block && base < Q; base += kBatchMaxDry) {
            const int32_t rows = min(Q - base, kBatchMaxDry);
            const int64_t units = static_cast<int64_t>(rows) * nwin;
            for (int64_t u = gw; u < units; u += GW) scan_window_unit(base, rows, u, false);
            if (base + kBatchMaxDry < Q) grid_sync(st.bar, st.ctl, phase, G);  // the partial buffers are reused
        }
        grid_sync(st.bar, st.ctl, phase, G, timed ? &c_ph[5] : nullptr);
        const long long tp1 = timed ? clock64() : 0;

        // (test hook, option loop_debug = 77: block 1 leaves in its third iteration -- the others must give the barrier up
        // and report it, not hang: tests/test_gpu_parity.py::test_barrier_timeout_is_reported_and_the_context_survives)
        if (prm.debug == 77 && it == 2u && bid == 1u && G > 1u) asm volatile("exit;");
        // ================= P2: heads and stoppers; candidates go to one global list =================
        int p2_remote = 0;  // this block pushed candidates into the peers' boxes
        {
            uint64_t bstop = kPackInf, bhead = kPackInf, dropped = kPackInf;
            uint64_t my_head = kPackInf;  // sharded: this thread's smallest head as a candidate record
            uint4 my_w0 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u), my_w1 = make_uint4(0u, 0u, 0u, 0u);
            bool pushed = false;
            int32_t* const cnt_cand = kMulti ? xb_cnt + sl : st.counters + sl * 4 + CN_CAND;
            uint4* const cand_out = st.cand;
            constexpr int kP2R = 2;  // rows per thread and tile: one round trip and one block reduction for both
            for (int32_t r0 = r_lo; r0 < r_hi; r0 += kP2R * GT) {
                RowHead h[kP2R];
                int32_t sr[kP2R], rr[kP2R];
#pragma unroll
                for (int x = 0; x < kP2R; ++x) {
                    const int32_t r = r0 + x * GT + gtid;
                    rr[x] = r;
                    h[x].head = h[x].stop = h[x].bound = kPackInf;
                    h[x].partner_slot = h[x].partner_key = kNoPartner;
                    h[x].partner_size = 0;
                    sr[x] = 0;
                    if (r < r_hi) {  // one round trip: key, own size, list flags and the first two entries
                        const int32_t key_r = __ldcg(st.gkey + r);
                        sr[x] = __ldcg(st.lsize + r);
                        const uint32_t mb = static_cast<uint32_t>(__ldcg(st.nn_more + r));
                        const uint4 e0 = __ldcg(st.nn + static_cast<int64_t>(r) * kNNK);
                        const uint4 e1 = __ldcg(st.nn + static_cast<int64_t>(r) * kNNK + 1);
                        if (key_r >= 0) h[x] = row_head(e0, e1, mb, static_cast<uint32_t>(key_r));
                    }
                }
                uint64_t tstop_min = h[0].stop;
#pragma unroll
                for (int x = 1; x < kP2R; ++x) tstop_min = umin64(tstop_min, h[x].stop);
                bstop = umin64(bstop, block_min_u64(tstop_min, s_red));  // running minimum: any value >= T is a valid filter
#pragma unroll
                for (int x = 0; x < kP2R; ++x) {
                    bhead = umin64(bhead, umin64(h[x].head, h[x].bound));  // (a bound only ever stops the loop at the horizon)
                    if (kMulti && h[x].head < my_head) {
                        my_head = h[x].head;
                        my_w0 = make_uint4(static_cast<uint32_t>(h[x].head), static_cast<uint32_t>(h[x].head >> 32),
                                           static_cast<uint32_t>(rr[x]), h[x].partner_slot);
                        my_w1 = make_uint4(static_cast<uint32_t>(sr[x]), static_cast<uint32_t>(h[x].partner_size), h[x].partner_key, 0u);
                    }
                    if (h[x].head < bstop) {
                        const int32_t k = atomicAdd(cnt_cand, 1);
                        if (!kMulti || k < kBatchXCand) {
                            const uint4 w0 = make_uint4(static_cast<uint32_t>(h[x].head), static_cast<uint32_t>(h[x].head >> 32),
                                                        static_cast<uint32_t>(rr[x]), h[x].partner_slot);
                            const uint4 w1 = make_uint4(static_cast<uint32_t>(sr[x]), static_cast<uint32_t>(h[x].partner_size),
                                                        h[x].partner_key, 0u);
                            pushed = true;
                            if (kMulti) {  // into every rank's box (region of this rank)
                                for (int q = 0; q < st.n_ranks; ++q) {
                                    uint4* dst = reinterpret_cast<uint4*>(st.xbox[q] + kBatchXCandBase) +
                                                 2 * (static_cast<int64_t>(st.rank) * kBatchXCand + k);
                                    __stcg(dst, w0);
                                    __stcg(dst + 1, w1);
                                }
                            } else {
                                uint4* dst = cand_out + 2 * static_cast<int64_t>(k);
                                __stcg(dst, w0);
                                __stcg(dst + 1, w1);
                            }
                        } else {
                            dropped = umin64(dropped, h[x].head);  // does not fit the exchange box: nothing at or above it may be taken
                        }
                    }
                }
            }
            bhead = block_min_u64(bhead, s_red);
            if (kMulti) {  // the block's smallest REAL head (packs are unique: one owner), or "none" (bhead may be a bound)
                const uint64_t breal = block_min_u64(my_head, s_red);
                if (breal == kPackInf ? tid == 0 : my_head == breal) {
                    if (breal == kPackInf) my_w0 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
                    __stcg(st.blockmin + 2 * bid, my_w0);
                    __stcg(st.blockmin + 2 * bid + 1, my_w1);
                }
            }
            if (kMulti) p2_remote = __syncthreads_or(pushed ? 1 : 0);
            if (kMulti) {  // rank-wide minima in the exchange box
                dropped = block_min_u64(dropped, s_red);
                if (tid == 0) {
                    const uint64_t bs = umin64(bstop, dropped);
                    if (bs != kPackInf) atomicMin(xb_stop + sl, static_cast<unsigned long long>(bs));
                    if (bhead != kPackInf) atomicMin(xb_head + sl, static_cast<unsigned long long>(bhead));
                }
            } else if (tid == 0) {
                __stcg(st.hdr + bid, make_uint4(static_cast<uint32_t>(bstop), static_cast<uint32_t>(bstop >> 32),
                                                       static_cast<uint32_t>(bhead), static_cast<uint32_t>(bhead >> 32)));
            }
        }
        if (kMulti)
            grid_sync_ranks(st, phase, xcount, G, bid, p2_remote != 0, sl, timed ? &c_ph[6] : nullptr);
        else
            grid_sync(st.bar, st.ctl, phase, G, timed ? &c_ph[6] : nullptr);
        const long long tp2 = timed ? clock64() : 0;

        // ================= P3: the batch =================
        // lists of this thread's first rows: loaded now, validated after the update pass (one DRAM round trip less)
        uint4 ve[kVR][kNNK];
#pragma unroll
        for (int x = 0; x < kVR; ++x) {
            const int32_t r = r_lo + gtid + x * GT;
#pragma unroll
            for (int y = 0; y < kNNK; ++y)
                ve[x][y] = r < r_hi ? __ldcg(st.nn + static_cast<int64_t>(r) * kNNK + y) : nn_none();
        }
        // (one GPU: this thread's first candidate record, requested before the count is known -- the list holds a record per
        // row -- which takes a dependent round trip off the phase)
        uint4 spec0 = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u), spec1 = make_uint4(0u, 0u, 0u, 0u);
        if (!kMulti && tid < n) {
            spec0 = __ldcg(st.cand + 2 * static_cast<int64_t>(tid));
            spec1 = __ldcg(st.cand + 2 * static_cast<int64_t>(tid) + 1);
        }
        uint64_t tstop = kPackInf, H = kPackInf;
        int32_t n_pub;  // heads below their block's stopper minimum, all ranks
        int over_local = 0;  // some rank's exact-evaluation queue overflowed in the previous iteration
        if (kMulti) {   // every rank's minima and candidate count, read from its exchange box
            if (tid < st.n_ranks) {  // summary of rank `tid`, pushed into this rank's box
                const uint4* sm = reinterpret_cast<const uint4*>(xb + kBatchXSummary + (static_cast<size_t>(tid) * 3 + sl) * 32);
                const uint4 s0 = __ldcg(sm), s1 = __ldcg(sm + 1);
                tstop = (static_cast<uint64_t>(s0.y) << 32) | s0.x;
                H = (static_cast<uint64_t>(s0.w) << 32) | s0.z;
                s_xcnt[tid + 1] = min(static_cast<int32_t>(s1.x), kBatchXCand);
                s_xover[tid] = static_cast<int32_t>(s1.x) > kBatchXCand ? 1 : 0;
                over_local = static_cast<int>(s1.y) | (s_xover[tid] ? 4 : 0);
            }
            __syncthreads();
            if (tid == 0) {
                s_xcnt[0] = 0;
                for (int q = 0; q < st.n_ranks; ++q) s_xcnt[q + 1] += s_xcnt[q];
            }
            __syncthreads();
            n_pub = s_xcnt[st.n_ranks];
        } else {
            n_pub = __ldcg(st.counters + sl * 4 + CN_CAND);
            if (tid == 0) over_local = (__ldcg(ctl + CTL_XQ_OVERFLOW) != 0 ? 1 : 0) | (__ldcg(ctl + CTL_ORDER_VIOL) != 0 ? 2 : 0);
            if (tid < static_cast<int>(G)) {
                const uint4 h0 = __ldcg(st.hdr + tid);
                tstop = (static_cast<uint64_t>(h0.y) << 32) | h0.x;
                H = (static_cast<uint64_t>(h0.w) << 32) | h0.z;
            }
        }
        auto cand_ptr = [&](int32_t i) -> const uint4* {  // candidate pair i of the concatenated lists
            if (!kMulti) return st.cand + 2 * static_cast<int64_t>(i);
            int q = 0;
            while (i >= s_xcnt[q + 1]) ++q;
            return reinterpret_cast<const uint4*>(xb + kBatchXCandBase) + 2 * (static_cast<int64_t>(q) * kBatchXCand + (i - s_xcnt[q]));
        };
        // this thread's candidates tid, tid + kBT, ...: first halves {pack, slots} kept in registers for the count and the filter
        // pass (a few thousand heads are published for a batch of dozens)
        constexpr int kCandReg = 4;
        uint4 cq[kCandReg];
        uint4 c1 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int x = 0; x < kCandReg; ++x) {
            cq[x] = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u);
            const int32_t i = tid + x * kBT;
            if (i < n_pub) cq[x] = (!kMulti && x == 0) ? spec0 : __ldcg(cand_ptr(i));
        }
        if (tid < n_pub) c1 = kMulti ? __ldcg(cand_ptr(tid) + 1) : spec1;
        tstop = block_min_u64(tstop, s_red);
        H = block_min_u64(H, s_red);
        const bool xq_over = __syncthreads_or(over_local & 1) != 0;     // (every rank sees every rank's flags)
        const bool cand_over = kMulti && __syncthreads_or(over_local & 4) != 0;  // some rank's candidates did not fit its region
        const bool order_viol = __syncthreads_or(over_local & 2) != 0;
        const long long ts1 = timed ? clock64() : 0;
        // termination (clustering.go:220 loop condition, :222-225 exhaustion)
        if (n_live <= prm.n_target)
            stop_reason = STOP_TARGET;
        else if (!pack_selectable(H))
            stop_reason = STOP_EXHAUSTED;
        else if (prm.max_merges >= 0 && launched >= prm.max_merges)
            stop_reason = STOP_MAX_MERGES;
        else if (it + 8u >= (1u << 22))
            stop_reason = STOP_EPOCHS;
        else if (st.compact_at > 0 && n_live <= st.compact_at)
            stop_reason = STOP_COMPACT;  // half of the slots are retired: the host renumbers the live clusters (compact.cu)
        else if (exact && order_viol)
            stop_reason = STOP_ORDER;  // the previous batch was not the reference's sequence: the host starts over with delta_cut
        else if (exact && xq_over)
            stop_reason = STOP_XQ;  // the previous iteration could not queue all its pairs: the host re-evaluates its rows
        else if (exact && static_cast<double>(pack_dist(H)) > prm.safe)
            stop_reason = STOP_HORIZON;  // the minimum reached the horizon: the host raises it (refine.cu)
        if (stop_reason != 0) {
            if (bid == 0) {  // what FindClosestClusters would return now
                if (!pack_selectable(H)) {
                    if (tid == 0) {
                        ctl[CTL_NEXT_HI] = -1;
                        ctl[CTL_NEXT_LO] = -1;
                        ctl[CTL_NEXT_DIST] = static_cast<int32_t>(kInfBits);
                    }
                } else {
                    if (tid == 0) {  // distance and row key are the pack itself; the partner comes from the pair's record
                        ctl[CTL_NEXT_HI] = static_cast<int32_t>(pack_key(H));
                        ctl[CTL_NEXT_DIST] = static_cast<int32_t>(H >> 32);
                    }
                    for (int32_t i = tid; i < n_pub; i += kBT) {
                        const uint4 p = __ldcg(cand_ptr(i));
                        if (((static_cast<uint64_t>(p.y) << 32) | p.x) == H)
                            ctl[CTL_NEXT_LO] = static_cast<int32_t>(__ldcg(cand_ptr(i) + 1).z);
                    }
                }
            }
            break;
        }
        int32_t limit = n_live - prm.n_target;
        if (prm.max_merges >= 0) limit = min(limit, prm.max_merges - launched);
        limit = min(limit, kMaxBatch);

        // candidate pairs below theta are examined; theta = the stopper minimum unless more than kMaxBatch pairs are
        // below it (any prefix of a valid batch is a valid batch): bisection on the packed value, packs are unique
        uint64_t theta = tstop;
        if (cand_over) theta = H + 1ull;  // only the global minimum (always valid); its record comes from the rank minima
        if (!cand_over && n_pub > kMaxBatch) {
            auto count_lt = [&](uint64_t th) {
                int c = 0;
#pragma unroll
                for (int x = 0; x < kCandReg; ++x) c += ((static_cast<uint64_t>(cq[x].y) << 32) | cq[x].x) < th ? 1 : 0;  // (none: all ones)
                for (int32_t i = tid + kCandReg * kBT; i < n_pub; i += kBT) {
                    const uint4 p = __ldcg(cand_ptr(i));
                    c += ((static_cast<uint64_t>(p.y) << 32) | p.x) < th ? 1 : 0;
                }
                return block_sum_i32(c, s_redi);
            };
            if (count_lt(tstop) > kMaxBatch) {
                uint64_t lo_t = 0, hi_t = tstop;  // count(lo) <= kMaxBatch < count(hi)
                while (hi_t - lo_t > 1) {
                    const uint64_t mid = lo_t + (hi_t - lo_t) / 2;
                    if (count_lt(mid) <= kMaxBatch)
                        lo_t = mid;
                    else
                        hi_t = mid;
                }
                theta = lo_t;
            }
        }
        const long long ts2 = timed ? clock64() : 0;
        if (tid == 0) s_m = 0;
        for (int32_t w = tid; w < n_words; w += kBT) s_bits[w] = 0u;
        __syncthreads();
        auto take_cand = [&](int32_t i, const uint4& p, bool have_p1) {
            const uint64_t hp = (static_cast<uint64_t>(p.y) << 32) | p.x;
            if (hp < theta) {  // (theta <= all ones: an empty register slot never passes)
                const uint4 p1 = have_p1 ? c1 : __ldcg(cand_ptr(i) + 1);
                const int k = atomicAdd(&s_m, 1);
                s_hp[k] = hp;
                s_ca[k] = static_cast<int32_t>(p.z);
                s_cb[k] = static_cast<int32_t>(p.w);
                s_csa[k] = static_cast<int32_t>(p1.x);
                s_csb[k] = static_cast<int32_t>(p1.y);
                s_ckb[k] = static_cast<int32_t>(p1.z);
            }
        };
#pragma unroll
        for (int x = 0; x < kCandReg; ++x) take_cand(tid + x * kBT, cq[x], x == 0);
        for (int32_t i = tid + kCandReg * kBT; i < n_pub; i += kBT) take_cand(i, __ldcg(cand_ptr(i)), false);
        __syncthreads();
        if (cand_over && tid == 0) {  // the smallest head of every rank that dropped candidates (if it is not listed already)
            for (int q = 0; q < st.n_ranks; ++q) {
                if (!s_xover[q]) continue;
                const uint4* rec = reinterpret_cast<const uint4*>(xb + kBatchXRankMin + (static_cast<size_t>(q) * 3 + sl) * 32);
                const uint4 p = __ldcg(rec), p1 = __ldcg(rec + 1);
                const uint64_t hp = (static_cast<uint64_t>(p.y) << 32) | p.x;
                bool have = !(hp < theta);
                for (int k = 0; k < s_m && !have; ++k) have = s_hp[k] == hp;
                if (!have && s_m < kMaxBatch) {
                    const int k = s_m++;
                    s_hp[k] = hp;
                    s_ca[k] = static_cast<int32_t>(p.z);
                    s_cb[k] = static_cast<int32_t>(p.w);
                    s_csa[k] = static_cast<int32_t>(p1.x);
                    s_csb[k] = static_cast<int32_t>(p1.y);
                    s_ckb[k] = static_cast<int32_t>(p1.z);
                }
            }
        }
        if (cand_over) __syncthreads();
        const int32_t n_cand = s_m;  // <= kMaxBatch
        const int32_t n_cand_prof = n_cand + (n_pub << 10) * 0;
        const long long ts3 = timed ? clock64() : 0;
        // slot conflicts: a pair that shares a cluster with an earlier pair is a stopper
        uint64_t mine = kPackInf, conf = kPackInf;
        int n_less = 0;
        if (tid < n_cand) {
            mine = s_hp[tid];
            const int32_t a = s_ca[tid], b = s_cb[tid];
            // branch-free and unrolled: with a short-circuit || every trip waited for its own shared-memory loads (~80 cycles
            // per candidate, 13 k cycles per iteration at config C)
            int hit = 0;
#pragma unroll 8
            for (int j = 0; j < n_cand; ++j) {
                const int32_t aj = s_ca[j], bj = s_cb[j];
                const int less = s_hp[j] < mine ? 1 : 0;
                n_less += less;
                hit |= less & ((aj == a ? 1 : 0) | (aj == b ? 1 : 0) | (bj == a ? 1 : 0) | (bj == b ? 1 : 0));
            }
            if (hit != 0) conf = mine;
        }
        const uint64_t T = umin64(theta, block_min_u64(conf, s_red));
        bool take = tid < n_cand && mine < T;
        int was_cut = 0;
        if (exact && take) {
            // decisions only among values the horizon guarantees to be the reference's own; and because fp32 centroid
            // distances are reducible only up to rounding, a pair after the first must stay clear of the stopper: every
            // distance the earlier merges create is >= T (1 - rounding) (both predicates are thresholds on the distance, so
            // the accepted pairs remain a prefix of the scan order)
            const uint32_t tb = static_cast<uint32_t>(T >> 32);
            const double dmine = static_cast<double>(pack_dist(mine));
            const double dT = tb < kInfBits ? static_cast<double>(__uint_as_float(tb)) : static_cast<double>(INFINITY);
            if (dmine > prm.safe) {
                take = false;
            } else if (n_less > 0 && prm.delta_cut > 0.0 && !(dmine <= dT * (1.0 - prm.delta_cut))) {
                take = false;
                was_cut = 1;
            }
        }
        const int rank = take ? n_less : -1;  // everything below an accepted pair is accepted
        const int32_t m_all = block_sum_i32(rank >= 0 ? 1 : 0, s_redi);  // >= 1: the global minimum head is always among them
        if (m_all <= 0) {  // cannot happen (the global minimum is a candidate below every stopper): never spin on it
            stop_reason = STOP_ERROR;
            break;
        }
        const int32_t m = min(m_all, limit);                              // merges of this iteration
        if (exact && bid == 0 && __syncthreads_or(was_cut) != 0 && tid == 0) atomicAdd(ctl + CTL_N_CUT, 1);
        if (rank >= 0) {
            s_d[rank] = static_cast<uint32_t>(mine >> 32);
            if (rank < m) {
                const uint32_t a = static_cast<uint32_t>(s_ca[tid]), b = static_cast<uint32_t>(s_cb[tid]);
                s_a[rank] = static_cast<int32_t>(a);
                s_b[rank] = static_cast<int32_t>(b);
                s_sa[rank] = s_csa[tid];
                s_sb[rank] = s_csb[tid];
                s_ka[rank] = static_cast<int32_t>(pack_key(mine));
                s_kb[rank] = s_ckb[tid];
                atomicOr(&s_bits[a >> 5], 1u << (a & 31u));
                atomicOr(&s_bits[b >> 5], 1u << (b & 31u));
            }
        }
        if (tid == 0) {  // runner-up of the last pair: the stopper (a lower bound of what comes next)
            const uint32_t tb = static_cast<uint32_t>(T >> 32);
            s_d[m_all] = tb < kInfBits ? tb : kInfBits;
        }
        __syncthreads();
        const long long tp3 = timed ? clock64() : 0;

        // ---- bookkeeping (block 0): trace, slot table; counters of the next iterations ----
        if (bid == 0) {
            if (tid < m) {
                const int32_t a = s_a[tid], b = s_b[tid], snew = s_sa[tid] + s_sb[tid], new_key = kbase + t + tid;
                const float d = __uint_as_float(s_d[tid]), sd = __uint_as_float(s_d[tid + 1]);
                const float gap = (sd - d) / fmaxf(d, 1e-30f);
                st.tr_key_hi[t + tid] = s_ka[tid];
                st.tr_key_lo[t + tid] = s_kb[tid];
                st.tr_dist[t + tid] = d;
                st.tr_size[t + tid] = snew;
                st.tr_gap[t + tid] = gap;
                if (gap < prm.near_tie_tol) atomicAdd(ctl + CTL_NEAR_TIES, 1);
                __stcg(st.ks + a, make_int2(-1, 0));
                __stcg(st.gkey + a, -1);
                __stcg(st.ks + b, make_int2(new_key, snew));
                __stcg(st.gkey + b, new_key);
                __stcg(st.lsize + a, 0);
                __stcg(st.lsize + b, snew);
                // the new cluster's row: queued for a scan (all live clusters have a lower key)
                __stcg(st.nn + static_cast<int64_t>(b) * kNNK, nn_bound(0u));
#pragma unroll
                for (int x = 1; x < kNNK; ++x) __stcg(st.nn + static_cast<int64_t>(b) * kNNK + x, nn_none());
                __stcg(st.nn_more + b, static_cast<int32_t>(kMoreBit | kDryBit));
                __stcg(st.nn_more + a, 0);
                if (st.near_meta != nullptr) {
                    st.slot_of_key[s_ka[tid]] = -1;
                    st.slot_of_key[s_kb[tid]] = -1;
                    st.slot_of_key[new_key] = b;
                    st.near_meta[b] = make_int2(0, -1);  // (one GPU: replaced by the row's re-evaluated pairs after the exact phase)
                }
                // (one GPU with the horizon: the exact phase decides whether the row needs a scan at all)
                if (!use_xres && b >= r_lo && b < r_hi) st.dryq[atomicAdd(st.counters + sl1 * 4 + CN_DRY, 1)] = make_int2(b, new_key);
            }
            if (tid == kBT - 1) {
                st.counters[sl2 * 4 + CN_DRY] = 0;
                st.counters[sl2 * 4 + CN_NEAR] = 0;
                st.counters[sl1 * 4 + CN_CAND] = 0;
                st.counters[sl2 * 4 + CN_XQ] = 0;   // last read in the exact phase of the previous iteration
                ctl[CTL_XQ_FIRST_KEY] = kbase + t;  // the clusters this iteration creates carry the keys N + t ...
                if (kMulti) {  // this rank's exchange-box slot of the next iteration (last read two iterations ago)
                    xb_cnt[sl1] = 0;
                    xb_stop[sl1] = kPackInf;
                    xb_head[sl1] = kPackInf;
                }
            }
        }

        bool wrote_remote = false;
        // ---- Lance-Williams rows: unit = (merge, kUpdCols columns), one warp each, all loads of a unit in flight ----
        // A pair's distance lives in the row of its HIGHER-key cluster (the new cluster's row is written in full, nothing
        // is mirrored into the older rows: that cost one scattered sector per live cluster and merge).  d(c,a) of a
        // cluster c newer than a is therefore gathered from row c; early in the loop almost every column is older.
        {
            // sharded: a rank updates the columns of its own row block -- rows a and b are coalesced (possibly remote) reads,
            // the gathers from the rows of newer clusters c are local, its part of the new row is a coalesced remote store
            // (multiples of 4: a rank whose block starts beyond the last slot has r_lo == n, which need not be one)
            const int32_t c_lo = kMulti ? min(n4, st.rank * C) : 0, c_hi = kMulti ? min(n4, c_lo + C) : n4;
            const int32_t n_chunks = (c_hi - c_lo + kUpdCols - 1) / kUpdCols;
            // a unit = kUpdGroup consecutive merges of the batch x kUpdCols columns: the slot-table chunk ({key, size} of the
            // columns: a third of the phase's L2 sectors when every merge fetched its own) is loaded once per unit
            const int32_t mg = (m + kUpdGroup - 1) / kUpdGroup;
            const int32_t units = mg * n_chunks;
            for (int32_t u = gw; u < units; u += GW) {
              const int32_t ch = u / mg, i_first = (u - ch * mg) * kUpdGroup;
              constexpr int kI = kUpdCols / 128;
              int4 k01[kI], k23[kI];
#pragma unroll
              for (int x = 0; x < kI; ++x) {
                  const int32_t c0 = c_lo + ch * kUpdCols + x * 128 + lane * 4;
                  k01[x] = k23[x] = make_int4(-1, 0, -1, 0);
                  if (c0 < c_hi) {
                      k01[x] = __ldcg(reinterpret_cast<const int4*>(st.ks + c0));
                      k23[x] = __ldcg(reinterpret_cast<const int4*>(st.ks + c0 + 2));
                  }
              }
              for (int32_t i = i_first; i < min(m, i_first + kUpdGroup); ++i) {
                const int32_t a = s_a[i], b = s_b[i], sa = s_sa[i], sb = s_sb[i], snew = sa + sb;
                const int32_t ka = s_ka[i], kb = s_kb[i];
                if (kMulti && (b < r_lo || b >= r_hi)) wrote_remote = true;  // this unit's slice of the new row goes to a peer
                const float dab = __uint_as_float(s_d[i]);
                const double sad = static_cast<double>(sa), sbd = static_cast<double>(sb), dabd = static_cast<double>(dab);
                const float* row_a = row_of(a);
                float* row_b = row_of(b);
                float4 va[kI], vb[kI];
#pragma unroll
                for (int x = 0; x < kI; ++x) {
                    const int32_t c0 = c_lo + ch * kUpdCols + x * 128 + lane * 4;
                    va[x] = vb[x] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c0 < c_hi) {
                        va[x] = __ldcg(reinterpret_cast<const float4*>(row_a + c0));
                        vb[x] = __ldcg(reinterpret_cast<const float4*>(row_b + c0));
                    }
                }
                // Everything below is branch-free per element (predicated gathers, selects): a branch per element made the
                // sixteen Lance-Williams chains of a lane run one after the other (measured: ~2 000 cycles per element).
                float da[kI][4], db[kI][4];
                uint32_t livem[kI];
#pragma unroll
                for (int x = 0; x < kI; ++x) {  // second round trip, only for the columns of newer clusters
                    const int32_t c0 = c_lo + ch * kUpdCols + x * 128 + lane * 4;
                    const uint32_t bits = c0 < c_hi ? (s_bits[c0 >> 5] >> (c0 & 31)) & 0xFu : 0xFu;  // c0 % 4 == 0: one word
                    const int32_t keys[4] = {k01[x].x, k01[x].z, k23[x].x, k23[x].z};
                    const float ra[4] = {va[x].x, va[x].y, va[x].z, va[x].w};
                    const float rb[4] = {vb[x].x, vb[x].y, vb[x].z, vb[x].w};
                    livem[x] = 0u;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const bool live = keys[e] >= 0 && ((bits >> e) & 1u) == 0u;
                        livem[x] |= live ? (1u << e) : 0u;
                        const float* rc = dm + static_cast<int64_t>(c0 + e - r_lo) * ld;  // row of cluster c: local
                        // (pairs of two clusters older than the last compaction are stored in both rows)
                        db[x][e] = ldcg_if(rc + b, live && keys[e] > kb && keys[e] >= st.mirror_key, rb[e]);
                        da[x][e] = ldcg_if(rc + a, live && keys[e] > ka && keys[e] >= st.mirror_key, ra[e]);
                    }
                    livem[x] |= bits << 4;
                }
                float outv[kI][4];
                uint32_t hitm = 0u;  // elements written at or below the horizon: the exact phase owns them
#pragma unroll
                for (int x = 0; x < kI; ++x) {
                    const int32_t sizes[4] = {k01[x].y, k01[x].w, k23[x].y, k23[x].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        // lance_williams_rcp() with the integer -> double conversions hoisted (sums of small integers are exact);
                        // computed for every element, kept where the pair is live and admissible (else +inf for good, :228)
                        const int den = snew + sizes[e];
                        const double skd = static_cast<double>(sizes[e]);
                        const double num = ((sad + skd) * static_cast<double>(da[x][e]) + (sbd + skd) * static_cast<double>(db[x][e])) - skd * dabd;
                        const double rcp = small_sizes ? s_rcp[min(max(den, 0), kRcpTab - 1)] : 1.0 / static_cast<double>(max(den, 1));
                        const float lw = canon_dist(static_cast<float>(num * rcp));
                        const bool keep = ((livem[x] >> e) & 1u) != 0u && den <= prm.max_size;
                        outv[x][e] = keep ? lw : __uint_as_float(kInfBits);
                        if (exact && keep && static_cast<double>(lw) <= prm.horizon) hitm |= 1u << (x * 4 + e);
                    }
                }
                if (exact && __any_sync(0xffffffffu, hitm != 0u)) {  // the merge's queue: {column, Lance-Williams value}
                    int32_t pos = warp_reserve(st.xhit + i, __popc(hitm), lane);
#pragma unroll
                    for (int x = 0; x < kI; ++x)
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if ((hitm >> (x * 4 + e)) & 1u) {
                                const int32_t col = c_lo + ch * kUpdCols + x * 128 + lane * 4 + e;
                                const int32_t lwb = static_cast<int32_t>(__float_as_uint(outv[x][e]));
                                if (pos < kXResCap) {
                                    st.xqm[static_cast<int64_t>(i) * kXResCap + pos] = make_int2(col, lwb);
                                } else {  // more pairs than a merge's queue holds: the shared overflow queue
                                    const int32_t idx = atomicAdd(st.counters + sl * 4 + CN_XQ, 1);
                                    if (idx < st.xq_cap) {
                                        st.xq[idx] = make_int4(i, col, lwb, pos);
                                    } else {  // nowhere to queue it: the Lance-Williams value is stored, and the host's sweep
                                        ctl[CTL_XQ_OVERFLOW] = 1;  // of the band (STOP_XQ) finds and replaces it
                                        hitm &= ~(1u << (x * 4 + e));
                                    }
                                }
                                ++pos;
                            }
                }
#pragma unroll
                for (int x = 0; x < kI; ++x) {
                    const int32_t c0 = c_lo + ch * kUpdCols + x * 128 + lane * 4;
                    if (c0 >= c_hi) continue;
                    // columns of this batch's clusters belong to the cross-term pass, queued pairs to the exact phase
                    const uint32_t bits = (livem[x] >> 4) | ((hitm >> (x * 4)) & 0xFu);
                    if (bits == 0u) {
                        __stcg(reinterpret_cast<float4*>(row_b + c0), make_float4(outv[x][0], outv[x][1], outv[x][2], outv[x][3]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (((bits >> e) & 1u) == 0u) __stcg(row_b + c0 + e, outv[x][e]);
                    }
                }
              }
            }
        }
        // ---- centroids of the new clusters (MergeClusters, clustering.go:36-40): row N + t of the centroid store ----
        // (every rank keeps a replica; read by the exact phase after the next barrier)
        if (exact) {
            const int32_t nch = (d4 + 127) / 128;
            for (int32_t u = gw; u < m * nch; u += GW) {
                const int32_t j = u / nch, e = ((u - j * nch) * 32 + lane) * 4;
                if (e < d4) {
                    const float* pa = st.cen + static_cast<int64_t>(s_ka[j]) * st.ldc + e;  // a: the larger position (clustering.go:237)
                    const float* pb = st.cen + static_cast<int64_t>(s_kb[j]) * st.ldc + e;
                    const float4 ca = __ldcg(reinterpret_cast<const float4*>(pa)), cb = __ldcg(reinterpret_cast<const float4*>(pb));
                    const float fa = static_cast<float>(s_sa[j]), fb = static_cast<float>(s_sb[j]), fs = static_cast<float>(s_sa[j] + s_sb[j]);
                    __stcg(reinterpret_cast<float4*>(st.cen + static_cast<int64_t>(kbase + t + j) * st.ldc + e),
                           make_float4(ex_merge1(fa, ca.x, fb, cb.x, fs), ex_merge1(fa, ca.y, fb, cb.y, fs),
                                       ex_merge1(fa, ca.z, fb, cb.z, fs), ex_merge1(fa, ca.w, fb, cb.w, fs)));
                }
            }
        }
        const long long tq1 = timed ? clock64() : 0;
        // ---- cross terms: d(new_i, new_j), i < j, two chained updates from the four old entries ----
        for (int32_t j = gw; j < m; j += GW) {
            const int32_t aj = s_a[j], bj = s_b[j], saj = s_sa[j], sbj = s_sb[j], kaj = s_ka[j], kbj = s_kb[j];
            const float dj = __uint_as_float(s_d[j]);
            if (kMulti && (bj < r_lo || bj >= r_hi)) continue;  // the owner of new cluster j's row computes its cross terms
            for (int32_t i = lane; i < j; i += 32) {
                const int32_t ai = s_a[i], bi = s_b[i], sai = s_sa[i], sbi = s_sb[i], si = sai + sbi, kai = s_ka[i], kbi = s_kb[i];
                const float di = __uint_as_float(s_d[i]);
                auto pair = [&](int32_t p, int32_t kp, int32_t q, int32_t kq) {  // row of the higher key
                    return kp > kq ? __ldcg(row_of(p) + q) : __ldcg(row_of(q) + p);
                };
                const float x1 = pair(ai, kai, aj, kaj), x2 = pair(bi, kbi, aj, kaj), x3 = pair(ai, kai, bj, kbj),
                            x4 = pair(bi, kbi, bj, kbj);
                // merge i seen from k = a_j and k = b_j
                float t1 = __uint_as_float(kInfBits), t2 = __uint_as_float(kInfBits);
                if (saj + si <= prm.max_size) t1 = lance_williams(sai, sbi, saj, x1, x2, di);
                if (sbj + si <= prm.max_size) t2 = lance_williams(sai, sbi, sbj, x3, x4, di);
                // merge j seen from k = new_i (slot b_i, size si)
                float val = __uint_as_float(kInfBits);
                if (si + saj + sbj <= prm.max_size) val = lance_williams(saj, sbj, si, t1, t2, dj);
                if (exact && static_cast<double>(val) <= prm.horizon) {  // (lanes diverge here: rare)
                    const int32_t pos = atomicAdd(st.xhit + j, 1);
                    const int32_t colx = static_cast<int32_t>(0x80000000u | static_cast<uint32_t>(i));
                    const int32_t lwb = static_cast<int32_t>(__float_as_uint(val));
                    if (pos < kXResCap) {
                        st.xqm[static_cast<int64_t>(j) * kXResCap + pos] = make_int2(colx, lwb);
                    } else {
                        const int32_t idx = atomicAdd(st.counters + sl * 4 + CN_XQ, 1);
                        if (idx < st.xq_cap)
                            st.xq[idx] = make_int4(j, colx, lwb, pos);
                        else
                            ctl[CTL_XQ_OVERFLOW] = 1;
                    }
                } else {
                    __stcg(row_of(bj) + bi, val);  // new_j carries the higher key
                }
            }
        }
        const long long tq2 = timed ? clock64() : 0;
        // ---- list validation: partners merged in this batch are dead ----
        auto validate = [&](int32_t r, const uint4 (&e)[kNNK]) {
            if ((s_bits[r >> 5] >> (r & 31)) & 1u) return;  // merged rows: handled by block 0
            bool dead = false;
#pragma unroll
            for (int x = 0; x < kNNK; ++x)
                if (e[x].z != kNoPartner && ((s_bits[e[x].z >> 5] >> (e[x].z & 31u)) & 1u)) dead = true;
            if (!dead) return;
            const int32_t key_r = __ldcg(st.gkey + r);
            if (key_r < 0) return;
            uint4 keep[kNNK];
            int kept = 0;
            uint32_t last = 0u;
#pragma unroll
            for (int x = 0; x < kNNK; ++x) {
                if (e[x].y == kNoPartner) continue;  // empty
                last = e[x].y;                        // distance of the last listed entry (valid, dead or bound)
                if (e[x].z == kNoPartner) continue;  // an old bound
                if ((s_bits[e[x].z >> 5] >> (e[x].z & 31u)) & 1u) continue;
#pragma unroll
                for (int y = 0; y < kNNK; ++y)
                    if (y == kept) keep[y] = e[x];
                ++kept;
            }
            const bool more = (static_cast<uint32_t>(__ldcg(st.nn_more + r)) & kMoreBit) != 0u;
#pragma unroll
            for (int x = 0; x < kNNK; ++x) {
                uint4 v = nn_none();
                if (x < kept)
                    v = keep[x];
                else if (x == kept && more)
                    v = nn_bound(last);
                __stcg(st.nn + static_cast<int64_t>(r) * kNNK + x, v);
            }
            // a list that is down to one entry is refilled, too: its bound placeholder would be a stopper right above its
            // head and cut every batch there (measured at config C: 73 merges per iteration instead of 11, 2x the scans)
            if (kept < prm.refill_at && more) {
                __stcg(st.nn_more + r, static_cast<int32_t>(kMoreBit | kDryBit));
                queue_row(r, key_r, sl1);
            }
        };
#pragma unroll
        for (int x = 0; x < kVR; ++x)
            if (r_lo + gtid + x * GT < r_hi) validate(r_lo + gtid + x * GT, ve[x]);
        for (int32_t r = r_lo + gtid + kVR * GT; r < r_hi; r += GT) {
            uint4 e[kNNK];
#pragma unroll
            for (int y = 0; y < kNNK; ++y) e[y] = __ldcg(st.nn + static_cast<int64_t>(r) * kNNK + y);
            validate(r, e);
        }
        t += m;
        launched += m;
        n_live -= m;
        ++iters;
        const long long tq3 = timed ? clock64() : 0;
        // ---- exact phase: the pairs this iteration wrote at or below the horizon get the reference's own value ----
        // WardDistance(centroid k, centroid of the new cluster), clustering.go:83-86; the new centroid (clustering.go:39) is
        // formed on the fly from the two stored ones, which are only overwritten after the barrier that ends this phase.
        if (exact) {
            grid_sync(st.bar, st.ctl, phase, G);  // gpu scope: every queue is local to its rank
            const long long te0 = timed ? clock64() : 0;
            const int32_t nx = min(__ldcg(st.counters + sl * 4 + CN_XQ), st.xq_cap);
            const float d_last = __uint_as_float(s_d[m > 0 ? m - 1 : 0]);
            if (use_xres && bid == 0 && tid < m) {  // new rows without a pair at or below the horizon (or with too many): scan
                const int32_t c = __ldcg(st.xhit + tid);  // (a row without any pair at or below the horizon needs no scan: its list is a bound)
                if (c > kXResCap || __ldcg(st.counters + sl * 4 + CN_XQ) > st.xq_cap)
                    st.dryq[atomicAdd(st.counters + sl1 * 4 + CN_DRY, 1)] = make_int2(s_b[tid], kbase + t - m + tid);
            }
            int32_t my_exact = 0;
            float my_maxerr = 0.0f;
            // what one lane does with the reference's value of its pair: monitors, the matrix entry, the new row's list
            auto finish = [&](int32_t j, int32_t col, int32_t key_b, int32_t size_b, int32_t lwb, int32_t pos, float dsq) {
                const int32_t bj = s_b[j];
                const float w = ward_weight(s_sa[j] + s_sb[j], size_b, dsq);
                my_maxerr = fmaxf(my_maxerr, exact_monitor(ctl, __uint_as_float(static_cast<uint32_t>(lwb)), w, prm.eps_filter, prm.abs_slack));
                if (j + 1 < m && w < d_last) atomicAdd(ctl + CTL_ORDER_VIOL, 1);  // would have preceded a later pair of the batch
                __stcg(row_of(bj) + col, w);
                if (use_xres && pos < kXResCap)
                    __stcg(st.xres + static_cast<int64_t>(j) * kXResCap + pos,
                           make_uint4(__float_as_uint(w), static_cast<uint32_t>(key_b), static_cast<uint32_t>(col), static_cast<uint32_t>(size_b)));
                if (kMulti && (bj < r_lo || bj >= r_hi)) wrote_remote = true;
            };
            auto decode = [&](int32_t colx, int32_t& col, int32_t& key_b, int32_t& size_b) {
                if (colx < 0) {  // cross term: the other cluster was created by this batch, too
                    const int32_t i = colx & 0x7FFFFFFF;
                    size_b = s_sa[i] + s_sb[i];
                    col = s_b[i];
                    key_b = kbase + t - m + i;
                } else {
                    col = colx;
                    size_b = __ldcg(st.lsize + col);
                    key_b = __ldcg(st.gkey + col);
                }
            };
            // groups of up to kExGroup queued pairs of one merge: they share the new cluster's centroid and the chain latency
            {
                int32_t cj = 0;
                if (tid < m) cj = min(__ldcg(st.xhit + tid), kXResCap);
                if (tid < kMaxBatch) s_xcntm[tid] = cj;
                int32_t g = (cj + kExGroup - 1) / kExGroup, inc = g;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int32_t v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                if (lane == 31) s_wsum[warp] = inc;
                __syncthreads();
                int32_t woff = 0;
                for (int w2 = 0; w2 < warp; ++w2) woff += s_wsum[w2];
                if (tid < kMaxBatch) s_upre[tid + 1] = woff + inc;
                if (tid == 0) s_upre[0] = 0;
                __syncthreads();
            }
            const int32_t n_groups = s_upre[min(m, kMaxBatch)];
            for (int32_t u = gw; warp < kExWarps && u < n_groups; u += static_cast<int32_t>(G) * kExWarps) {
#if IC_PROF_EXACT
                const long long tx0 = timed ? clock64() : 0;
#endif
                int32_t lo_j = 0, hi_j = m;  // largest j with s_upre[j] <= u
                while (hi_j - lo_j > 1) {
                    const int32_t mid = (lo_j + hi_j) >> 1;
                    if (s_upre[mid] <= u)
                        lo_j = mid;
                    else
                        hi_j = mid;
                }
                const int32_t j = lo_j, p0 = (u - s_upre[j]) * kExGroup, np = min(kExGroup, s_xcntm[j] - p0);
                const float* pa = st.cen + static_cast<int64_t>(kbase + t - m + j) * st.ldc;  // written in the phase before the barrier
                int32_t col = 0, key_b = 0, size_b = 0, lwb = 0;
                const float* pb = nullptr;
                if (lane < np) {
                    const int2 ent = __ldcg(st.xqm + static_cast<int64_t>(j) * kXResCap + p0 + lane);
                    lwb = ent.y;
                    decode(ent.x, col, key_b, size_b);
                    pb = st.cen + static_cast<int64_t>(key_b) * st.ldc;
                }
#if IC_PROF_EXACT
                const long long tx1 = timed ? clock64() : 0;
#endif
#if IC_EXACT_SYNC  // (A/B: register-staged chunks, one in flight)
                const float dsq = warp_exact_dsq_group<kExGroup, kExChunk>(pa, pb, np, d4, s_ex[warp]);
#else
                const float dsq = warp_exact_dsq_group_async<kExStages>(pa, pb, np, d4, s_ex[warp]);
#endif
#if IC_PROF_EXACT
                const long long tx2 = timed ? clock64() : 0;
#endif
                if (lane < np) finish(j, col, key_b, size_b, lwb, p0 + lane, dsq);
                if (lane == 0) my_exact += np;
#if IC_PROF_EXACT
                if (timed) {
                    c_sel[1] += tx1 - tx0;
                    c_sel[2] += tx2 - tx1;
                    c_sel[3] += 1;
                }
#endif
            }
            // pairs beyond a merge's queue (rare): one at a time
            for (int32_t q = gw; q < nx; q += GW) {
                const int4 ent = __ldcg(st.xq + q);
                const int32_t j = ent.x;
                const float* pa = st.cen + static_cast<int64_t>(kbase + t - m + j) * st.ldc;
                int32_t size_b, col, key_b;
                decode(ent.y, col, key_b, size_b);
                const float dsq = warp_exact_dsq(pa, st.cen + static_cast<int64_t>(key_b) * st.ldc, d4, s_one[warp]);
                if (lane == 0) {
                    finish(j, col, key_b, size_b, ent.z, ent.w, dsq);
                    ++my_exact;
                }
            }
            if (lane == 0 && my_exact > 0) atomicAdd(ctl + CTL_N_EXACT, my_exact);
            exact_monitor_flush(ctl, my_maxerr);
            if (timed) {
                c_ph[8] += te0 - tq3;        // waiting for the slowest block's rows phase
                c_ph[9] += clock64() - te0;  // this block's share of the exact phase
            }
        }
        if (kMulti)
            grid_sync_ranks(st, phase, xcount, G, bid, __syncthreads_or(wrote_remote ? 1 : 0) != 0, -1);
        else
            grid_sync(st.bar, st.ctl, phase, G);
        m_prev = m;
        if (timed) {
            const long long tp4 = clock64();
            c_ph[0] += tp1 - tp0;
            c_ph[1] += tp2 - tp1;
            c_ph[2] += tp3 - tp2;
            c_ph[3] += tp4 - tp3;
            c_ph[4] += tq1 - tp3;
            c_ph[7] += tp4 - tq3;
            c_sel[0] += ts1 - tp2;
#if !IC_PROF_EXACT
            c_sel[1] += ts2 - ts1;
            c_sel[2] += ts3 - ts2;
            c_sel[3] += n_cand_prof;
#endif
        }
    }
    if (timed) {  // accumulated over the launches of one clustering (the host zeroes them when it starts)
        st.prof[0] += c_ph[0];
        st.prof[1] += c_ph[1];
        st.prof[2] += c_ph[2];
        st.prof[3] += c_ph[3];
        st.prof[5] += launched;
        st.prof[6] += iters;
        for (int i = 0; i < 6; ++i) st.prof[10 + i] += c_ph[4 + i];
        st.prof[4] += c_sel[0];
        st.prof[7] += c_sel[1];
        st.prof[8] += c_sel[2];
        st.prof[9] += c_sel[3];
    }
    if (bid == 0 && tid == 0) {
        ctl[CTL_N_LIVE] = n_live;
        ctl[CTL_N_MERGES] = t;
        ctl[CTL_EXHAUSTED] = stop_reason == STOP_EXHAUSTED ? 1 : 0;
        ctl[CTL_ITERS] = ctl[CTL_ITERS] + iters;
        ctl[CTL_RESCANS] = ctl[CTL_RESCANS] + static_cast<int32_t>(n_rescans);
        ctl[CTL_STOP] = stop_reason;
        if (stop_reason == STOP_ERROR) ctl[CTL_ERROR] = 3;  // no candidate below the stopper: a protocol bug, reported by the host
        __threadfence();
        ctl[CTL_DONE] = 1;
    }
}

cudaError_t merge_batch_grid(int num_sms, int64_t n, int* blocks) {
    *blocks = 0;
    const size_t smem = merge_batch_smem_bytes(n);
    cudaError_t e = cudaFuncSetAttribute(merge_batch_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(merge_batch_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(merge_batch_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return cudaSuccess;  // does not fit: *blocks stays 0
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, merge_batch_kernel<true, true>, kBT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaSuccess;
    // small problems: fewer blocks make the grid barriers cheaper
    int64_t want = std::max<int64_t>(8, (n + 127) / 128);
    *blocks = static_cast<int>(std::min<int64_t>(std::min(num_sms, kBatchMaxBlocks), want));
    return cudaSuccess;
}

cudaError_t launch_merge_batch(const BatchState& st, const LoopParams& p, int blocks, cudaStream_t s) {
    if (blocks <= 0) return cudaErrorInvalidConfiguration;
    BatchState st_copy = st;
    st_copy.win_cols = static_cast<int32_t>(batch_window_cols(st.n));
    st_copy.n_win = static_cast<int32_t>(merge_batch_windows(st.n));
    LoopParams p_copy = p;
    const BatchState* none = nullptr;
    int32_t bpr = blocks;
    void* args[] = {&st_copy, &p_copy, &none, &bpr};
    const size_t smem = merge_batch_smem_bytes(st.n);
    const bool multi = st.n_ranks > 1;
    void* fn = multi ? reinterpret_cast<void*>(merge_batch_kernel<true, false>) : reinterpret_cast<void*>(merge_batch_kernel<false, false>);
    cudaError_t e = multi ? cudaFuncSetAttribute(merge_batch_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
                          : cudaFuncSetAttribute(merge_batch_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    return cudaLaunchCooperativeKernel(fn, dim3(blocks), dim3(kBT), args, smem, s);
}

// P virtual ranks in one cooperative launch of P * blocks_per_rank blocks; d_states = the ranks' BatchStates in device memory
// (win_cols / n_win filled in by the caller through merge_batch_fill_windows)
void merge_batch_fill_windows(BatchState* st) {
    st->win_cols = static_cast<int32_t>(batch_window_cols(st->n));
    st->n_win = static_cast<int32_t>(merge_batch_windows(st->n));
}
cudaError_t launch_merge_batch_virtual(const BatchState* d_states, int n_ranks, int64_t n, const LoopParams& p, int blocks_per_rank,
                                       cudaStream_t s) {
    if (blocks_per_rank <= 0 || n_ranks < 2) return cudaErrorInvalidConfiguration;
    BatchState dummy{};
    LoopParams p_copy = p;
    int32_t bpr = blocks_per_rank;
    void* args[] = {&dummy, &p_copy, &d_states, &bpr};
    const size_t smem = merge_batch_smem_bytes(n);
    cudaError_t e = cudaFuncSetAttribute(merge_batch_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(merge_batch_kernel<true, true>), dim3(blocks_per_rank * n_ranks), dim3(kBT),
                                       args, smem, s);
}

}  // namespace ic
