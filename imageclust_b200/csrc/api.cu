// api.cu -- C ABI (include/imageclust_b200.h) and the host engine behind it.
//
// The engine owns one device, one stream and the resident problem:
//   X [n x d] fp32  ->  K0 prep (centre, TF32 hi/lo split, norms)
//                   ->  K1 initial Ward distances, full square [n x ld] fp32 in HBM
//                   ->  K2 first nearest-neighbour sweep
//                   ->  K3 persistent merge loop (one launch for all merges)
//                   ->  merge trace back to the host -> cluster lists
// mirroring PerformClusteringWithConstraints (clustering.go:198-284).  There is no CPU
// path: every entry point that computes needs the device, and fails without one.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/imageclust_b200.h"
#include "kernels.h"

using namespace ic;

struct ic_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int num_sms = 0;
    std::string err;
    // options
    double near_tie_tol = 1e-5;
    int center = 1;
    int gram_mode = IC_GRAM_TCGEN05_I8;
    int verbose = 0;
    int gram_terms = 23;      // debug: which products of the split K1 issues
    int loop_blocks = 0;      // merge-loop blocks per rank, 0 = auto
    int loop_blocks_alloc = -1;
    int profile_loop = 0;     // debug: per-phase cycle counters of the merge loop
    long long h_prof[256] = {0};  // [0,16): phases of block 0; [16, 16+G): exchange wait of every block
    // row-block sharding.  vranks > 1: P virtual shards emulated on this one GPU by one
    // cooperative launch (same kernel code path as the multi-GPU build; test hook).
    int vranks = 1;
    int vranks_alloc = -1;
    int scan_every = 4;  // merge loop: rescans are requested every scan_every-th iteration
    int loop_debug = 0;  // experiments only
    int gram_debug = 0;  // experiments only (gram_i8.cu)
    bool dm_lower_only = false;  // K1 stored dm[i][j] for j <= i only (batched loop: the upper triangle is never read)
    int loop_mode_alloc = -1, loop_mode_used = -1;  // layout the allocation was made for; loop that touched the matrix
    bool batch_layout = false;  // scratch of the batched loop is allocated
    int loop_mode = 1;   // 1: batched loop (merge_batch.cu) on an unsharded context; 0: one merge per iteration (merge_loop.cu)
    int batch_grid = 0;
    uint8_t* batch_scratch = nullptr;  // hdr | cand | counters | dryq | partials | part_cnt | bar
    size_t batch_scratch_bytes = 0, batch_off[12] = {0};
    int no_replica = 0, no_replica_alloc = -1;  // test hook: stream the keys from L2 even when the replica would fit
    bool loop_replica = false;
    uint32_t loop_gen = 0;    // generation of the last merge-loop launch (mailbox tags)
    int64_t loop_launches = 0;
    int32_t loop_launches_run = 0;  // launches of the loop kernel in the current clustering
    int64_t loop_n_target = 0, loop_max_size = 0, loop_max_merges = -1;
    int32_t loop_merges_at_start = 0;
    // real shards: one process / GPU / rank (ic_shard_*)
    int shard_rank = 0, shard_world = 1;
    int shard_rank_alloc = -1, shard_world_alloc = -1;
    bool peers_open = false;
    void* peer_dm[kMaxRanks] = {nullptr};   // every rank's row block (own + cudaIpc mapped)
    void* peer_dm_b[kMaxRanks] = {nullptr}; // every rank's second matrix buffer (compaction target), if allocated
    void* peer_box[kMaxRanks] = {nullptr};  // every rank's inter-rank mailbox
    uint64_t barrier_seq = 0;
    // resident problem
    int64_t n = 0, d = 0, n_pad = 0, d_pad = 0, ld = 0;
    float* x = nullptr;  // [n x d] dense
    float *hi = nullptr, *lo = nullptr;
    double *colsum = nullptr, *norms = nullptr;
    int2* tiles = nullptr;
    int n_tiles = 0;
    // int8 path (IC_GRAM_TCGEN05_I8)
    int8_t *i8h = nullptr, *i8m = nullptr, *i8l = nullptr;
    float* quanta = nullptr;
    int64_t d_pad8 = 0;
    int2* tiles8 = nullptr;
    int n_tiles8 = 0;
    bool prepped = false, prepped_i8 = false;
    float* dm = nullptr;
    SlotKS* ks = nullptr;
    int32_t* gkey = nullptr;
    SlotNN* nn = nullptr;
    int32_t* nn_more = nullptr;
    int32_t *tr_key_hi = nullptr, *tr_key_lo = nullptr, *tr_size = nullptr;
    float *tr_dist = nullptr, *tr_gap = nullptr;
    uint8_t *records = nullptr, *partials = nullptr, *rankbox = nullptr;  // mailboxes, zeroed once at allocation
    BatchState* vstates = nullptr;  // the virtual ranks' states of the batched loop (test hook)
    long long* prof = nullptr;
    int32_t* ctl = nullptr;
    int loop_grid = 0;  // blocks per rank
    bool loaded = false, have_dm = false, have_nn = false;
    int gram_mode_used = -1;
    // loop state mirrored on the host
    int64_t n_target = 0;
    int32_t n_live = 0, n_merges = 0, exhausted = 0;
    int32_t h_ctl[kCtlWords] = {0};
    // reference arithmetic: centroids, horizon, queues of refine.cu and of the loop's exact phase
    int exact_opt = 1;            // option "exact": 0 off (Lance-Williams values only), 1 auto, 2 on even for ic_set_matrix
    bool exact_on = false;        // this clustering runs with the horizon
    bool dm_is_reference = false; // the initial matrix already holds the reference's values (gram_mode 1)
    // delta_cut = 0: batches are taken optimistically and CHECKED (every pair a batch creates is compared with the batch's
    // later members, CTL_ORDER_VIOL); a failed check restarts the clustering with delta_cut_fallback
    double hz_factor_first = 0.0;  // option "horizon_factor_first": factor of the first horizon (0: the same as later ones)
    double hz_factor = 1.18, eps_filter = 3e-5, delta_cut = 0.0, delta_cut_fallback = 1e-5, abs_slack_opt = -1.0;
    double delta_cut_cur = 0.0;
    int32_t n_restarts = 0;
    double horizon = -1.0, abs_slack = 0.0, hz_factor_cur = 1.18;
    int32_t merges_at_raise = 0;
    float* cen = nullptr;
    int64_t ldc = 0;
    int4* xq = nullptr;
    int32_t xq_cap = 0;
    uint4* xres = nullptr;
    int32_t* xhit = nullptr;
    int2* xqm = nullptr;
    int32_t* xfar = nullptr;
    // near lists (near.cu)
    int2 *near_meta = nullptr, *near_meta_b = nullptr;
    uint2* near_pool = nullptr;
    int32_t near_pool_cap = 0;
    int32_t *near_cursor = nullptr, *slot_of_key = nullptr;
    int near_opt = 1;  // option "near_lists"
    double ms_near = 0.0;
    // K4 compaction (compact.cu): current epoch's geometry, second copies of the per-slot state and of the matrix
    int compact_opt = 1, compact_opt_alloc = -1;  // option "compact"
    int refill_at = 2;            // option "refill_at" (1 or 2)
    double compact_ratio = 0.7, compact_ratio_alloc = -1.0;  // option "compact_ratio": compact when live <= ratio * slots
    int mirror_init = 1;          // option "mirror_init": fill the upper triangle after a lower-triangle-only K1
    // option "fast_start": with the horizon and near lists on, the first sweep (K2) is not needed -- its only product that
    // survives the first horizon is the global minimum, which the mirror pass after K1 collects on the way
    int fast_start = 1;
    uint32_t* gmin_dev = nullptr;  // [1] bits of the smallest selectable initial distance (mirror pass)
    bool gmin_valid = false;
    int64_t compact_min = 4096;   // no compaction below this many slots
    int compact_tiles = 1;        // option "compact_tiles": one-pass tile kernel (one process); 0: compact_rows + mirror_lower
    int64_t n_cur = 0, ld_cur = 0;
    float* dm_cur = nullptr;      // dm (buffer A) or dm_b
    float* dm_b = nullptr;
    size_t dm_b_floats = 0;
    SlotKS* ks_b = nullptr;
    int32_t *gkey_b = nullptr, *nn_more_b = nullptr;
    SlotNN* nn_b = nullptr;
    int32_t *keymap = nullptr, *newslot = nullptr, *oldslot = nullptr, *nlive_dev = nullptr;
    int32_t order_key = 0, mirror_key = 0;
    int32_t n_compactions = 0;
    double ms_compact = 0.0;
    int2* rq = nullptr;
    int32_t rq_cap = 0;
    int32_t* rq_cnt = nullptr;
    double ms_refine = 0.0;
    int32_t n_raises = 0;
    std::vector<int32_t> h_key_hi, h_key_lo, h_size;
    std::vector<float> h_dist, h_gap;
    bool trace_on_host = false;
    ic_stats stats{};
    cudaEvent_t ev[12] = {nullptr};  // [10], [11]: around every launch of the loop kernel
    double ms_loop_kernel = 0.0;     // device time of the loop kernel's launches alone (no sweeps, no compactions)
    bool loop_launch_pending = false;
};

namespace {

// NVTX range of one phase of the path (visible in nsys / ncu --nvtx): upload, K0, K1, K2, loop, refine, compact, readback
struct Nvtx {
    explicit Nvtx(const char* name) { nvtxRangePushA(name); }
    ~Nvtx() { nvtxRangePop(); }
};

int fail(ic_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define IC_CUDA(call)                                                                                    \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(ctx, e__ == cudaErrorMemoryAllocation ? IC_ERR_OOM : IC_ERR_CUDA,                \
                        std::string(#call) + ": " + cudaGetErrorString(e__));                            \
    } while (0)

template <typename T>
void dev_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

void close_peers(ic_ctx* c) {
    for (int q = 0; q < kMaxRanks; ++q) {
        if (c->peers_open && q != c->shard_rank_alloc) {
            if (c->peer_dm[q]) cudaIpcCloseMemHandle(c->peer_dm[q]);
            if (c->peer_dm_b[q]) cudaIpcCloseMemHandle(c->peer_dm_b[q]);
            if (c->peer_box[q]) cudaIpcCloseMemHandle(c->peer_box[q]);
        }
        c->peer_dm[q] = c->peer_dm_b[q] = c->peer_box[q] = nullptr;
    }
    c->peers_open = false;
}

void release_problem(ic_ctx* c) {
    close_peers(c);
    dev_free(c->x);
    dev_free(c->hi);
    dev_free(c->lo);
    dev_free(c->colsum);
    dev_free(c->norms);
    dev_free(c->tiles);
    dev_free(c->i8h);
    dev_free(c->i8m);
    dev_free(c->i8l);
    dev_free(c->quanta);
    dev_free(c->tiles8);
    dev_free(c->dm);
    dev_free(c->ks);
    dev_free(c->gkey);
    dev_free(c->nn);
    dev_free(c->nn_more);
    dev_free(c->tr_key_hi);
    dev_free(c->tr_key_lo);
    dev_free(c->tr_size);
    dev_free(c->tr_dist);
    dev_free(c->tr_gap);
    dev_free(c->records);
    dev_free(c->partials);
    dev_free(c->rankbox);
    dev_free(c->vstates);
    dev_free(c->batch_scratch);
    dev_free(c->cen);
    dev_free(c->xq);
    dev_free(c->xres);
    dev_free(c->xhit);
    dev_free(c->xqm);
    dev_free(c->xfar);
    dev_free(c->near_meta);
    dev_free(c->near_meta_b);
    dev_free(c->near_pool);
    dev_free(c->near_cursor);
    dev_free(c->slot_of_key);
    dev_free(c->dm_b);
    dev_free(c->gmin_dev);
    dev_free(c->ks_b);
    dev_free(c->gkey_b);
    dev_free(c->nn_more_b);
    dev_free(c->nn_b);
    dev_free(c->keymap);
    dev_free(c->newslot);
    dev_free(c->oldslot);
    dev_free(c->nlive_dev);
    c->dm_cur = nullptr;
    dev_free(c->rq);
    dev_free(c->rq_cnt);
    dev_free(c->prof);
    dev_free(c->ctl);
    c->loaded = c->have_dm = c->have_nn = c->prepped = c->prepped_i8 = false;
    c->n = c->d = 0;
    c->trace_on_host = false;
}

int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is resolved at run time so that the library loads (and exports its symbols) on a
// machine without a driver.
EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int make_operand_map(ic_ctx* ctx, CUtensorMap* map, float* base, int64_t rows, int64_t cols) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(ctx, IC_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * sizeof(float)};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kGramBK), 128u};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, IC_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string(r));
    return IC_OK;
}

// Lower-triangular tile list in square super-tiles of 2048 x 2048 outputs: the 16 + 8
// operand panels of a super-tile (64 MB with hi+lo at D=2048) stay L2 resident while its
// 128 tiles are computed, so every operand panel is fetched from HBM once per super-tile.
std::vector<int2> build_tile_list(int64_t n, int64_t row_begin, int64_t row_end) {
    std::vector<int2> tiles;
    const int rbs = static_cast<int>((n + kGramBM - 1) / kGramBM);
    const int SR = 16, SC = 8;
    for (int sr = 0; sr * SR < rbs; ++sr)
        for (int sc = 0; sc <= sr; ++sc)
            for (int rb = sr * SR; rb < (sr + 1) * SR && rb < rbs; ++rb)
                for (int cb = sc * SC; cb < (sc + 1) * SC; ++cb) {
                    const int64_t col0 = static_cast<int64_t>(cb) * kGramBN;
                    const int64_t row_last = static_cast<int64_t>(rb) * kGramBM + kGramBM - 1;
                    if (col0 >= n || col0 > row_last) continue;  // outside, or wholly above the diagonal
                    // a rank keeps its rows at full width and symmetric: it needs the tiles whose rows (direct entries)
                    // or whose columns (mirrored entries) touch its row block
                    const bool rows_in = !(row_last < row_begin || static_cast<int64_t>(rb) * kGramBM >= row_end);
                    const bool cols_in = !(col0 + kGramBN <= row_begin || col0 >= row_end);
                    if (!rows_in && !cols_in) continue;
                    tiles.push_back(make_int2(rb, cb));
                }
    return tiles;
}

// sharding geometry of the resident problem
int n_ranks(const ic_ctx* c) { return c->shard_world > 1 ? c->shard_world : c->vranks; }
int n_local(const ic_ctx* c) { return c->shard_world > 1 ? 1 : c->vranks; }
int rank0(const ic_ctx* c) { return c->shard_world > 1 ? c->shard_rank : 0; }
int64_t rows_per_rank(const ic_ctx* c) { return merge_loop_rows_per_rank(c->n, n_ranks(c)); }
// rows of the distance matrix resident on this device: [row_begin, row_end)
int64_t row_begin(const ic_ctx* c) { return c->shard_world > 1 ? std::min(c->n, c->shard_rank * rows_per_rank(c)) : 0; }
int64_t row_end(const ic_ctx* c) {
    return c->shard_world > 1 ? std::min(c->n, row_begin(c) + rows_per_rank(c)) : c->n;
}

// geometry of the CURRENT epoch (after compactions the matrix has n_cur dense slots, re-blocked over the ranks)
int64_t rows_per_rank_cur(const ic_ctx* c) { return merge_loop_rows_per_rank(c->n_cur, n_ranks(c)); }
int64_t row_begin_cur(const ic_ctx* c) { return c->shard_world > 1 ? std::min(c->n_cur, c->shard_rank * rows_per_rank_cur(c)) : 0; }
int64_t row_end_cur(const ic_ctx* c) {
    return c->shard_world > 1 ? std::min(c->n_cur, row_begin_cur(c) + rows_per_rank_cur(c)) : c->n_cur;
}
// first row of rank q's block in the current matrix buffer (own memory, a peer mapping, or -- virtual ranks -- this device)
float* rank_block(const ic_ctx* c, int q) {
    if (c->shard_world > 1) return static_cast<float*>(c->dm_cur == c->dm ? c->peer_dm[q] : c->peer_dm_b[q]);
    return c->dm_cur + static_cast<int64_t>(q) * rows_per_rank_cur(c) * c->ld_cur;
}

// 128 x 128 tiles of the int8 kernel, lower triangle, in 2048 x 2048 super-tiles (operand panels stay in L2)
std::vector<int2> build_tile_list_i8(int64_t n, int64_t row_begin, int64_t row_end, bool lower_only) {
    std::vector<int2> tiles;
    const int nb = static_cast<int>((n + kI8Tile - 1) / kI8Tile);
    const int S = 16;
    for (int sr = 0; sr * S < nb; ++sr)
        for (int sc = 0; sc <= sr; ++sc)
            for (int rb = sr * S; rb < (sr + 1) * S && rb < nb; ++rb)
                for (int cb = sc * S; cb < (sc + 1) * S && cb <= rb; ++cb) {
                    const int64_t r0 = static_cast<int64_t>(rb) * kI8Tile, c0 = static_cast<int64_t>(cb) * kI8Tile;
                    const bool rows_in = !(r0 + kI8Tile <= row_begin || r0 >= row_end);
                    // (a tile whose columns only touch the row block serves the mirrored stores: not needed when K1 stores the
                    // lower triangle only -- the batched loop -- and the upper one is filled by the mirror pass)
                    const bool cols_in = !lower_only && !(c0 + kI8Tile <= row_begin || c0 >= row_end);
                    if (!rows_in && !cols_in) continue;
                    tiles.push_back(make_int2(rb, cb));
                }
    return tiles;
}

int make_operand_map_i8(ic_ctx* ctx, CUtensorMap* map, int8_t* base, int64_t rows, int64_t cols) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(ctx, IC_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols)};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kI8BK), static_cast<cuuint32_t>(kI8Tile)};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(ctx, IC_ERR_CUDA, "cuTensorMapEncodeTiled (int8) failed: " + std::to_string(r));
    return IC_OK;
}

int alloc_problem(ic_ctx* ctx, int64_t n, int64_t d) {
    if (ctx->x && ctx->dm && ctx->n == n && ctx->d == d && n > 0 && ctx->loop_blocks == ctx->loop_blocks_alloc &&
        ctx->vranks == ctx->vranks_alloc && ctx->no_replica == ctx->no_replica_alloc && (ctx->exact_opt != 0) == (ctx->cen != nullptr) && ctx->compact_opt == ctx->compact_opt_alloc && ctx->compact_ratio == ctx->compact_ratio_alloc && ctx->loop_mode == ctx->loop_mode_alloc && ctx->shard_world == ctx->shard_world_alloc &&
        ctx->shard_rank == ctx->shard_rank_alloc) {
        // same shape as the resident problem: keep the HBM allocations (40 GB at N=100k)
        ctx->loaded = ctx->have_dm = ctx->have_nn = ctx->prepped = ctx->prepped_i8 = false;
        ctx->trace_on_host = false;
        return IC_OK;
    }
    release_problem(ctx);
    ctx->n = n;
    ctx->d = d;
    ctx->n_pad = round_up(n, kGramBN);
    ctx->d_pad = round_up(d, kGramBK);
    ctx->d_pad8 = round_up(d, kI8BK);
    ctx->ld = round_up(n, 32);
    ctx->loop_mode_alloc = ctx->loop_mode;
    ctx->batch_layout = false;
    ctx->batch_grid = 0;
    ctx->vranks_alloc = ctx->vranks;
    ctx->no_replica_alloc = ctx->no_replica;
    ctx->loop_replica = merge_loop_replica_fits(n) && !ctx->no_replica;
    ctx->shard_world_alloc = ctx->shard_world;
    ctx->shard_rank_alloc = ctx->shard_rank;
    ctx->loop_blocks_alloc = ctx->loop_blocks;
    const int P = n_ranks(ctx), NL = n_local(ctx);
    if (P > kMaxRanks) return fail(ctx, IC_ERR_BAD_ARG, "at most 8 ranks");
    // a sharded context holds only its own row block (all columns)
    const int64_t rows = ctx->shard_world > 1 ? rows_per_rank(ctx) : n;
    size_t free_b = 0, total_b = 0;
    IC_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const double other = 4.0 * n * d + (ctx->gram_mode == IC_GRAM_TCGEN05_3XTF32 ? 8.0 * ctx->n_pad * ctx->d_pad : 0.0) +
                         (ctx->gram_mode == IC_GRAM_TCGEN05_I8 ? 3.0 * ctx->n_pad * ctx->d_pad8 : 0.0) + 260.0 * n + (160 << 20) +
                         (ctx->exact_opt ? 8.0 * 192 * n + 8.0 * n * round_up(d, 4) + 16.0 * std::max<int64_t>(1 << 20, 16 * n) + 8.0 * (16 << 20) : 0.0);
    if (ctx->loop_mode == 1 && n > 0) {  // batched loop: one GPU, real shards, or virtual ranks (test hook)
        int grid = 0;
        IC_CUDA(merge_batch_grid(ctx->num_sms, n, &grid));
        if (grid > 0) {
            ctx->batch_layout = true;
            ctx->batch_grid = ctx->loop_blocks > 0 ? std::min(ctx->loop_blocks, ctx->num_sms) : grid;
        }
    }
    const double need = other + 4.0 * static_cast<double>(rows) * ctx->ld;
    if (need > static_cast<double>(free_b))
        return fail(ctx, IC_ERR_OOM, "problem needs " + std::to_string(need / 1e9) + " GB, device has " +
                                         std::to_string(free_b / 1e9) + " GB free");
    const size_t nn1 = static_cast<size_t>(n > 0 ? n : 1);
    const size_t n4 = (nn1 + 3) / 4 * 4;
    IC_CUDA(cudaMalloc(&ctx->x, sizeof(float) * nn1 * static_cast<size_t>(d > 0 ? d : 1)));
    IC_CUDA(cudaMalloc(&ctx->dm, sizeof(float) * static_cast<size_t>(rows > 0 ? rows : 1) *
                                     static_cast<size_t>(ctx->ld > 0 ? ctx->ld : 1)));
    // (virtual) rank v keeps its replica at ks + v * n (one-merge-per-iteration loop) or ks + v * (n4 + 4) (batched loop, which
    // reads the table four slots at a time: padding key -1)
    IC_CUDA(cudaMalloc(&ctx->ks, sizeof(SlotKS) * (n4 + 4) * NL));
    IC_CUDA(cudaMemsetAsync(ctx->ks, 0xFF, sizeof(SlotKS) * (n4 + 4) * NL, ctx->stream));
    IC_CUDA(cudaMalloc(&ctx->gkey, sizeof(int32_t) * n4 * NL));
    IC_CUDA(cudaMalloc(&ctx->nn, sizeof(SlotNN) * nn1 * kNNK));
    IC_CUDA(cudaMalloc(&ctx->nn_more, sizeof(int32_t) * nn1));
    IC_CUDA(cudaMalloc(&ctx->tr_key_hi, sizeof(int32_t) * nn1 * NL));
    IC_CUDA(cudaMalloc(&ctx->tr_key_lo, sizeof(int32_t) * nn1 * NL));
    IC_CUDA(cudaMalloc(&ctx->tr_size, sizeof(int32_t) * nn1 * NL));
    IC_CUDA(cudaMalloc(&ctx->tr_dist, sizeof(float) * nn1 * NL));
    IC_CUDA(cudaMalloc(&ctx->tr_gap, sizeof(float) * nn1 * NL));
    IC_CUDA(cudaMalloc(&ctx->ctl, sizeof(int32_t) * kCtlWords * NL));
    ctx->ldc = round_up(d > 0 ? d : 1, 4);
    ctx->compact_opt_alloc = ctx->compact_opt;
    ctx->compact_ratio_alloc = ctx->compact_ratio;
    if (ctx->exact_opt) {
        ctx->xq_cap = static_cast<int32_t>(std::max<int64_t>(1 << 20, 16 * static_cast<int64_t>(nn1)));
        ctx->rq_cap = static_cast<int32_t>(std::min<int64_t>(16 << 20, std::max<int64_t>(1024, static_cast<int64_t>(nn1) * static_cast<int64_t>(nn1) / 2)));
        IC_CUDA(cudaMalloc(&ctx->cen, sizeof(float) * 2 * nn1 * static_cast<size_t>(ctx->ldc)));  // by key: N items + up to N merges
        IC_CUDA(cudaMalloc(&ctx->xq, sizeof(int4) * static_cast<size_t>(ctx->xq_cap) * NL));
        IC_CUDA(cudaMalloc(&ctx->xres, sizeof(uint4) * static_cast<size_t>(kMaxBatch) * kXResCap));
        IC_CUDA(cudaMalloc(&ctx->xhit, sizeof(int32_t) * kMaxBatch * NL));
        IC_CUDA(cudaMalloc(&ctx->xqm, sizeof(int2) * static_cast<size_t>(kMaxBatch) * kXResCap * NL));
        IC_CUDA(cudaMalloc(&ctx->xfar, sizeof(int32_t) * kMaxBatch * NL));
        // near lists: ~150 pairs per item are at or below the horizon on the benchmark mixtures (initial + created)
        ctx->near_pool_cap = static_cast<int32_t>(std::min<int64_t>(0x7FFFFFF0, std::max<int64_t>(1 << 20, 192 * static_cast<int64_t>(nn1))));
        IC_CUDA(cudaMalloc(&ctx->near_meta, sizeof(int2) * (nn1 + 4)));
        IC_CUDA(cudaMalloc(&ctx->near_meta_b, sizeof(int2) * (nn1 + 4)));
        IC_CUDA(cudaMalloc(&ctx->near_pool, sizeof(uint2) * static_cast<size_t>(ctx->near_pool_cap)));
        IC_CUDA(cudaMalloc(&ctx->near_cursor, sizeof(int32_t) * 4));
        IC_CUDA(cudaMalloc(&ctx->slot_of_key, sizeof(int32_t) * (2 * nn1 + 4)));
        IC_CUDA(cudaMalloc(&ctx->rq, sizeof(int2) * static_cast<size_t>(ctx->rq_cap)));
        IC_CUDA(cudaMalloc(&ctx->rq_cnt, sizeof(int32_t) * 4));
    }
    IC_CUDA(cudaMalloc(&ctx->prof, sizeof(long long) * 256));
    // merge-loop launch geometry and mailboxes (zeroed once: tags carry the launch generation)
    IC_CUDA(merge_loop_grid(ctx->num_sms, n, P, NL, ctx->loop_blocks, ctx->loop_replica, &ctx->loop_grid));
    if (ctx->loop_grid <= 0) return fail(ctx, IC_ERR_OOM, "merge loop slice does not fit an SM's shared memory");
    const size_t rb = merge_loop_records_bytes(ctx->loop_grid), pb = merge_loop_partials_bytes(ctx->loop_grid),
                 xb = merge_loop_rankbox_bytes();
    IC_CUDA(cudaMalloc(&ctx->records, rb * NL));
    IC_CUDA(cudaMalloc(&ctx->partials, pb * NL));
    // (+ the exchange box of the sharded batched loop, behind the rank mailboxes: one IPC handle covers both)
    IC_CUDA(cudaMalloc(&ctx->rankbox, xb * NL + kBatchXBoxBytes * NL));
    IC_CUDA(cudaMalloc(&ctx->vstates, sizeof(BatchState) * kMaxRanks));
    IC_CUDA(cudaMemsetAsync(ctx->records, 0, rb * NL, ctx->stream));
    IC_CUDA(cudaMemsetAsync(ctx->partials, 0, pb * NL, ctx->stream));
    IC_CUDA(cudaMemsetAsync(ctx->rankbox, 0, xb * NL + kBatchXBoxBytes * NL, ctx->stream));
    IC_CUDA(cudaMemsetAsync(ctx->prof, 0, sizeof(long long) * 256, ctx->stream));
    if (ctx->batch_layout) {  // batched loop: scratch
        const size_t sizes[11] = {static_cast<size_t>(kBatchMaxBlocks) * 32, 32 * nn1,
                                  3 * 4 * 4, 8 * nn1, static_cast<size_t>(kBatchMaxDry) * kBatchMaxWin * 128,
                                  static_cast<size_t>(kBatchMaxDry) * 4, 256, 4 * (nn1 + 4), static_cast<size_t>(kBatchMaxBlocks) * 32, 8 * nn1, 0};
        size_t off = 0;
        for (int i = 0; i < 11; ++i) {
            ctx->batch_off[i] = off;
            off += (sizes[i] + 255) / 256 * 256;
        }
        ctx->batch_scratch_bytes = off;
        IC_CUDA(cudaMalloc(&ctx->batch_scratch, off * NL));
        IC_CUDA(cudaMemsetAsync(ctx->batch_scratch, 0, off * NL, ctx->stream));
    }
    // K4: second matrix buffer (a quarter of the first) and second copies of the per-slot state; one GPU, batched loop
    if (ctx->compact_opt && ctx->batch_layout && n >= ctx->compact_min) {
        const size_t half = static_cast<size_t>(static_cast<double>(n) * ctx->compact_ratio + 32);
        // a sharded context holds its row block of the compacted matrix only
        const size_t half_rows = ctx->shard_world > 1 ? static_cast<size_t>(merge_loop_rows_per_rank(static_cast<int64_t>(half), P)) : half;
        const size_t want = half_rows * static_cast<size_t>(round_up(static_cast<int64_t>(half), 32));
        size_t fb = 0, tb = 0;
        IC_CUDA(cudaMemGetInfo(&fb, &tb));
        if (static_cast<double>(want) * 4.0 + 64.0 * nn1 + (256 << 20) < static_cast<double>(fb)) {
            ctx->dm_b_floats = want;
            IC_CUDA(cudaMalloc(&ctx->dm_b, sizeof(float) * want));
            IC_CUDA(cudaMalloc(&ctx->ks_b, sizeof(SlotKS) * (n4 + 4) * NL));
            IC_CUDA(cudaMemsetAsync(ctx->ks_b, 0xFF, sizeof(SlotKS) * (n4 + 4) * NL, ctx->stream));
            IC_CUDA(cudaMalloc(&ctx->gkey_b, sizeof(int32_t) * n4 * NL));
            IC_CUDA(cudaMalloc(&ctx->nn_b, sizeof(SlotNN) * nn1 * kNNK));
            IC_CUDA(cudaMalloc(&ctx->nn_more_b, sizeof(int32_t) * nn1));
            IC_CUDA(cudaMalloc(&ctx->keymap, sizeof(int32_t) * (2 * nn1 + 4)));
            IC_CUDA(cudaMalloc(&ctx->newslot, sizeof(int32_t) * (nn1 + 4)));
            IC_CUDA(cudaMalloc(&ctx->oldslot, sizeof(int32_t) * (nn1 + 4)));
            IC_CUDA(cudaMalloc(&ctx->nlive_dev, sizeof(int32_t) * 4));
        }
    }
    IC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->loop_gen = 0;
    ctx->loop_launches = 0;
    ctx->barrier_seq = 0;
    ctx->h_key_hi.clear();
    ctx->h_key_lo.clear();
    ctx->h_size.clear();
    ctx->h_dist.clear();
    ctx->h_gap.clear();
    return IC_OK;
}

int loop_state(ic_ctx* c, LoopState* out) {
    LoopState st{};
    const int P = n_ranks(c), NL = n_local(c);
    st.n = static_cast<int32_t>(c->n);
    st.n_ranks = P;
    st.rank0 = rank0(c);
    st.n_local = NL;
    st.rows_per_rank = static_cast<int32_t>(rows_per_rank(c));
    st.ld = c->ld;
    if (c->shard_world > 1) {
        if (!c->peers_open) return fail(c, IC_ERR_STATE, "sharded context: ic_shard_connect has not run");
        for (int q = 0; q < P; ++q) {
            st.dm_rank[q] = static_cast<float*>(c->peer_dm[q]);
            st.rankbox[q] = c->peer_box[q];
        }
    } else {
        for (int q = 0; q < P; ++q) {
            st.dm_rank[q] = c->dm + static_cast<int64_t>(q) * st.rows_per_rank * c->ld;
            st.rankbox[q] = c->rankbox + static_cast<size_t>(q) * merge_loop_rankbox_bytes();
        }
    }
    st.ks = c->ks;
    st.gkey = c->gkey;
    st.nn = c->nn;
    st.nn_more = c->nn_more;
    st.tr_key_hi = c->tr_key_hi;
    st.tr_key_lo = c->tr_key_lo;
    st.tr_dist = c->tr_dist;
    st.tr_size = c->tr_size;
    st.tr_gap = c->tr_gap;
    st.records = c->records;
    st.records_stride = static_cast<int64_t>(merge_loop_records_bytes(c->loop_grid));
    st.partials = c->partials;
    st.partials_stride = static_cast<int64_t>(merge_loop_partials_bytes(c->loop_grid));
    st.ctl = c->ctl;
    st.prof = c->profile_loop ? c->prof : nullptr;
    *out = st;
    return IC_OK;
}

bool use_batch(const ic_ctx* c);

int do_prep_i8(ic_ctx* ctx) {
    if (ctx->prepped_i8) return IC_OK;
    Nvtx range("ic K0 prep (centre, int8 digits, norms)");
    const size_t pad_elems = static_cast<size_t>(ctx->n_pad) * static_cast<size_t>(ctx->d_pad8);
    if (!ctx->i8h) IC_CUDA(cudaMalloc(&ctx->i8h, pad_elems ? pad_elems : 1));
    if (!ctx->i8m) IC_CUDA(cudaMalloc(&ctx->i8m, pad_elems ? pad_elems : 1));
    if (!ctx->i8l) IC_CUDA(cudaMalloc(&ctx->i8l, pad_elems ? pad_elems : 1));
    if (!ctx->quanta) IC_CUDA(cudaMalloc(&ctx->quanta, sizeof(float) * static_cast<size_t>(ctx->n_pad ? ctx->n_pad : 1)));
    if (!ctx->colsum) IC_CUDA(cudaMalloc(&ctx->colsum, sizeof(double) * static_cast<size_t>(std::max(ctx->d_pad, ctx->d_pad8))));
    if (!ctx->norms) IC_CUDA(cudaMalloc(&ctx->norms, sizeof(double) * static_cast<size_t>(ctx->n_pad ? ctx->n_pad : 1)));
    IC_CUDA(launch_colsum(ctx->x, ctx->n, ctx->d, ctx->d, ctx->colsum, ctx->stream));
    IC_CUDA(launch_split_i8(ctx->x, ctx->n, ctx->d, ctx->d, ctx->colsum, ctx->center, ctx->i8h, ctx->i8m, ctx->i8l,
                            ctx->quanta, ctx->norms, ctx->n_pad, ctx->d_pad8, ctx->stream));
    ctx->stats.kernel_launches += 2;
    if (!ctx->tiles8) {
        const std::vector<int2> tiles = build_tile_list_i8(ctx->n, row_begin(ctx), row_end(ctx), use_batch(ctx));
        ctx->n_tiles8 = static_cast<int>(tiles.size());
        IC_CUDA(cudaMalloc(&ctx->tiles8, sizeof(int2) * (tiles.size() ? tiles.size() : 1)));
        IC_CUDA(cudaMemcpyAsync(ctx->tiles8, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice, ctx->stream));
        IC_CUDA(cudaStreamSynchronize(ctx->stream));  // the host vector goes out of scope
    }
    ctx->prepped_i8 = true;
    ctx->prepped = false;  // norms now belong to the int8 representation
    return IC_OK;
}

int do_prep(ic_ctx* ctx) {
    if (ctx->prepped) return IC_OK;
    Nvtx range("ic K0 prep (centre, tf32 split, norms)");
    ctx->prepped_i8 = false;  // norms will belong to the tf32 representation
    const size_t pad_elems = static_cast<size_t>(ctx->n_pad) * static_cast<size_t>(ctx->d_pad);
    if (!ctx->hi) IC_CUDA(cudaMalloc(&ctx->hi, sizeof(float) * pad_elems));
    if (!ctx->lo) IC_CUDA(cudaMalloc(&ctx->lo, sizeof(float) * pad_elems));
    if (!ctx->colsum) IC_CUDA(cudaMalloc(&ctx->colsum, sizeof(double) * static_cast<size_t>(std::max(ctx->d_pad, ctx->d_pad8))));
    if (!ctx->norms) IC_CUDA(cudaMalloc(&ctx->norms, sizeof(double) * static_cast<size_t>(ctx->n_pad)));
    IC_CUDA(launch_colsum(ctx->x, ctx->n, ctx->d, ctx->d, ctx->colsum, ctx->stream));
    IC_CUDA(launch_split(ctx->x, ctx->n, ctx->d, ctx->d, ctx->colsum, ctx->center, ctx->hi, ctx->lo, ctx->norms,
                         ctx->n_pad, ctx->d_pad, ctx->stream));
    ctx->stats.kernel_launches += 2;
    if (!ctx->tiles) {
        const std::vector<int2> tiles = build_tile_list(ctx->n, row_begin(ctx), row_end(ctx));
        ctx->n_tiles = static_cast<int>(tiles.size());
        IC_CUDA(cudaMalloc(&ctx->tiles, sizeof(int2) * (tiles.size() ? tiles.size() : 1)));
        IC_CUDA(cudaMemcpyAsync(ctx->tiles, tiles.data(), sizeof(int2) * tiles.size(), cudaMemcpyHostToDevice,
                                ctx->stream));
        IC_CUDA(cudaStreamSynchronize(ctx->stream));  // the host vector goes out of scope
    }
    ctx->prepped = true;
    return IC_OK;
}

bool use_batch(const ic_ctx* c);

int do_gram(ic_ctx* ctx, int mode) {
    Nvtx range("ic K1 initial distances");
    ctx->dm_lower_only = false;
    if (mode == IC_GRAM_EXACT_FP32) {
        IC_CUDA(launch_gram_exact(ctx->x, ctx->n, ctx->d, ctx->d, ctx->dm, ctx->ld, row_begin(ctx), row_end(ctx), ctx->stream));
        ctx->stats.kernel_launches += 1;
        return IC_OK;
    }
    if (mode == IC_GRAM_TCGEN05_I8) {
        GramI8Plan p8{};
        int rc8 = make_operand_map_i8(ctx, &p8.map_h, ctx->i8h, ctx->n_pad, ctx->d_pad8);
        if (rc8 == IC_OK) rc8 = make_operand_map_i8(ctx, &p8.map_m, ctx->i8m, ctx->n_pad, ctx->d_pad8);
        if (rc8 == IC_OK) rc8 = make_operand_map_i8(ctx, &p8.map_l, ctx->i8l, ctx->n_pad, ctx->d_pad8);
        if (rc8 != IC_OK) return rc8;
        p8.tiles = ctx->tiles8;
        p8.n_tiles = ctx->n_tiles8;
        p8.k_blocks = static_cast<int>(ctx->d_pad8 / kI8BK);
        ctx->dm_lower_only = use_batch(ctx);
        IC_CUDA(launch_gram_i8(p8, ctx->norms, ctx->quanta, ctx->dm, ctx->n, ctx->ld, row_begin(ctx), row_end(ctx),
                               ctx->num_sms, ctx->stream, ctx->gram_debug | (ctx->dm_lower_only ? 8 : 0)));
        ctx->stats.kernel_launches += 1;
        return IC_OK;
    }
    GramPlan plan{};
    int rc = make_operand_map(ctx, &plan.map_hi, ctx->hi, ctx->n_pad, ctx->d_pad);
    if (rc != IC_OK) return rc;
    rc = make_operand_map(ctx, &plan.map_lo, ctx->lo, ctx->n_pad, ctx->d_pad);
    if (rc != IC_OK) return rc;
    plan.tiles = ctx->tiles;
    plan.n_tiles = ctx->n_tiles;
    plan.k_blocks = static_cast<int>(ctx->d_pad / kGramBK);
    IC_CUDA(launch_gram_tcgen05(plan, ctx->norms, ctx->dm, ctx->n, ctx->ld, row_begin(ctx), row_end(ctx), ctx->num_sms,
                                ctx->stream, ctx->gram_terms));
    ctx->stats.kernel_launches += 1;
    return IC_OK;
}

// a fresh matrix: one slot per item, in key order, in buffer A
void reset_epoch(ic_ctx* ctx) {
    ctx->n_cur = ctx->n;
    ctx->ld_cur = ctx->ld;
    ctx->dm_cur = ctx->dm;
    ctx->order_key = static_cast<int32_t>(ctx->n);  // singletons: key == slot
    // a symmetric initial matrix holds every pair of two items in both rows; the lower-triangle-only K1 does not
    ctx->mirror_key = ctx->dm_lower_only ? 0 : static_cast<int32_t>(ctx->n);
    ctx->n_compactions = 0;
    ctx->ms_compact = 0.0;
}

int init_loop_state(ic_ctx* ctx) {
    const int NL = n_local(ctx);
    const size_t n = static_cast<size_t>(ctx->n), n4 = (n + 3) / 4 * 4;
    IC_CUDA(launch_init_slots(ctx->ks, ctx->gkey, ctx->n, ctx->stream));
    ctx->stats.kernel_launches += 1;
    if (ctx->exact_on && ctx->cen) {  // singleton centroids (clustering.go:19-20): rows 0..N-1 of the store (by key)
        IC_CUDA(launch_init_centroids(ctx->x, ctx->n, ctx->d, ctx->d, ctx->cen, ctx->ldc, ctx->stream));
        ctx->stats.kernel_launches += 1;
    }
    reset_epoch(ctx);
    if (ctx->slot_of_key) {
        IC_CUDA(launch_slot_of_key_init(ctx->slot_of_key, static_cast<int32_t>(ctx->n), static_cast<int32_t>(2 * ctx->n), ctx->stream));
        IC_CUDA(cudaMemsetAsync(ctx->near_meta, 0xFF, sizeof(int2) * (n + 4), ctx->stream));  // no near lists yet
        IC_CUDA(cudaMemsetAsync(ctx->near_cursor, 0, sizeof(int32_t) * 4, ctx->stream));
        ctx->stats.kernel_launches += 1;
    }
    ctx->ms_near = 0.0;
    ctx->ms_loop_kernel = 0.0;
    ctx->loop_launches_run = 0;
    if (ctx->prof) IC_CUDA(cudaMemsetAsync(ctx->prof, 0, sizeof(long long) * 256, ctx->stream));
    ctx->horizon = -1.0;
    ctx->abs_slack = ctx->abs_slack_opt >= 0.0 ? ctx->abs_slack_opt : 0.0;
    ctx->hz_factor_cur = ctx->hz_factor;
    ctx->merges_at_raise = 0;
    ctx->n_raises = 0;
    ctx->ms_refine = 0.0;
    ctx->n_live = static_cast<int32_t>(ctx->n);
    ctx->loop_mode_used = -1;
    ctx->n_merges = 0;
    ctx->exhausted = 0;
    ctx->trace_on_host = false;
    std::memset(ctx->h_ctl, 0, sizeof(ctx->h_ctl));
    ctx->h_ctl[CTL_N_LIVE] = ctx->n_live;
    for (int v = 0; v < NL; ++v) {
        if (v > 0 && n > 0) {  // every (virtual) rank keeps its own replica of the slot table
            const size_t ks_stride = use_batch(ctx) ? n4 + 4 : n;
            IC_CUDA(cudaMemcpyAsync(ctx->ks + v * ks_stride, ctx->ks, sizeof(SlotKS) * n, cudaMemcpyDeviceToDevice, ctx->stream));
            IC_CUDA(cudaMemcpyAsync(ctx->gkey + v * n4, ctx->gkey, sizeof(int32_t) * n4, cudaMemcpyDeviceToDevice,
                                    ctx->stream));
        }
        IC_CUDA(cudaMemcpyAsync(ctx->ctl + kCtlWords * v, ctx->h_ctl, sizeof(ctx->h_ctl), cudaMemcpyHostToDevice, ctx->stream));
    }
    return IC_OK;
}

int raise_horizon(ic_ctx* ctx);
int nn_init(ic_ctx* ctx);

// the batched loop runs on an unsharded context whose slice state fits (it always does below ~1e6 items)
bool use_batch(const ic_ctx* c) {
    return c->loop_mode == 1 && c->batch_layout && c->batch_grid > 0 && c->n > 0;
}

// One launch of the persistent loop (enqueued; sync_loop_result waits and relaunches if the kernel ran out of
// mailbox epochs, which takes ~1e6 iterations).
int enqueue_loop(ic_ctx* ctx, int64_t n_target, int64_t max_size, int64_t max_merges) {
    LoopState st{};
    int rc = loop_state(ctx, &st);
    if (rc != IC_OK) return rc;
    LoopParams p{};
    p.n_target = static_cast<int32_t>(n_target);
    p.max_size = static_cast<int32_t>(max_size > 0x3FFFFFFF ? 0x3FFFFFFF : max_size);
    p.max_merges = static_cast<int32_t>(max_merges < 0 ? -1 : (max_merges > 0x7FFFFFFF ? 0x7FFFFFFF : max_merges));
    p.near_tie_tol = static_cast<float>(ctx->near_tie_tol);
    p.scan_every = ctx->scan_every;
    p.debug = ctx->loop_debug;
    p.refill_at = ctx->refill_at;
    p.exact = (ctx->exact_on && use_batch(ctx) && ctx->cen) ? 1 : 0;
    p.eps_filter = static_cast<float>(ctx->eps_filter);
    p.abs_slack = static_cast<float>(ctx->abs_slack);
    p.horizon = ctx->horizon;
    p.safe = ctx->horizon < 0.0 ? -1.0 : (ctx->horizon - ctx->abs_slack) / (1.0 + 2.0 * ctx->eps_filter);
    p.delta_cut = ctx->delta_cut_cur;
    {   // the batched loop stops mirroring distances into the older clusters' rows: one loop per clustering
        const int mode = use_batch(ctx) ? 1 : 0;
        if (ctx->loop_mode_used >= 0 && ctx->loop_mode_used != mode && ctx->n_merges > 0)
            return fail(ctx, IC_ERR_STATE, "loop_mode changed in the middle of a clustering");
        if (mode == 0 && ctx->dm_lower_only)
            return fail(ctx, IC_ERR_STATE, "loop_mode changed after ic_initial_distances (the matrix holds the lower triangle only)");
        ctx->loop_mode_used = mode;
    }
    if (use_batch(ctx)) {
        const int VR = ctx->shard_world > 1 ? 1 : ctx->vranks;  // virtual ranks of this launch (test hook), else 1
        const size_t n4s = (static_cast<size_t>(ctx->n) + 3) / 4 * 4;
        if (ctx->shard_world > 1 && !ctx->peers_open) return fail(ctx, IC_ERR_STATE, "sharded context: ic_shard_connect has not run");
        if (VR > 1 || ctx->shard_world > 1) ++ctx->barrier_seq;
        std::vector<BatchState> states(static_cast<size_t>(VR));
        for (int v = 0; v < VR; ++v) {
            BatchState& bs = states[static_cast<size_t>(v)];
            bs = BatchState{};
            uint8_t* sc = ctx->batch_scratch + static_cast<size_t>(v) * ctx->batch_scratch_bytes;
            bs.n = static_cast<int32_t>(ctx->n_cur);
            bs.key_base = static_cast<int32_t>(ctx->n);
            bs.order_key = ctx->order_key;
            bs.mirror_key = ctx->mirror_key;
            bs.compact_at = (ctx->dm_b && ctx->n_cur >= ctx->compact_min) ? static_cast<int32_t>(static_cast<double>(ctx->n_cur) * ctx->compact_ratio) : 0;
            bs.n_ranks = 1;
            bs.rank = 0;
            bs.rows_per_rank = static_cast<int32_t>(rows_per_rank_cur(ctx));
            bs.ld = ctx->ld_cur;
            bs.dm = ctx->dm_cur;
            if (ctx->shard_world > 1) {  // real shards: peer-mapped row blocks and exchange boxes
                bs.n_ranks = ctx->shard_world;
                bs.rank = ctx->shard_rank;
                for (int q = 0; q < ctx->shard_world; ++q) {
                    bs.dm_rank[q] = rank_block(ctx, q);
                    bs.xbox[q] = static_cast<uint8_t*>(ctx->peer_box[q]) + merge_loop_rankbox_bytes();
                }
                bs.gen = static_cast<uint32_t>(ctx->barrier_seq);
            } else if (VR > 1) {  // virtual ranks: the peers' row blocks and boxes are addresses of this device
                bs.n_ranks = VR;
                bs.rank = v;
                for (int q = 0; q < VR; ++q) {
                    bs.dm_rank[q] = rank_block(ctx, q);
                    bs.xbox[q] = ctx->rankbox + merge_loop_rankbox_bytes() * VR + static_cast<size_t>(q) * kBatchXBoxBytes;
                }
                bs.dm = bs.dm_rank[v];
                bs.gen = static_cast<uint32_t>(ctx->barrier_seq);
            }
            // per (virtual) rank: replica of the slot table and keys, its own trace, control words, scratch and queue
            bs.ks = ctx->ks + static_cast<size_t>(v) * (n4s + 4);
            bs.gkey = ctx->gkey + static_cast<size_t>(v) * n4s;
            bs.nn = ctx->nn;
            bs.nn_more = ctx->nn_more;
            const size_t tro = static_cast<size_t>(v) * static_cast<size_t>(ctx->n > 0 ? ctx->n : 1);
            bs.tr_key_hi = ctx->tr_key_hi + tro;
            bs.tr_key_lo = ctx->tr_key_lo + tro;
            bs.tr_dist = ctx->tr_dist + tro;
            bs.tr_size = ctx->tr_size + tro;
            bs.tr_gap = ctx->tr_gap + tro;
            bs.ctl = ctx->ctl + kCtlWords * v;
            bs.prof = (ctx->profile_loop && v == 0) ? ctx->prof : nullptr;
            bs.hdr = reinterpret_cast<uint4*>(sc + ctx->batch_off[0]);
            bs.cand = reinterpret_cast<uint4*>(sc + ctx->batch_off[1]);
            bs.counters = reinterpret_cast<int32_t*>(sc + ctx->batch_off[2]);
            bs.dryq = reinterpret_cast<int2*>(sc + ctx->batch_off[3]);
            bs.partials = reinterpret_cast<uint4*>(sc + ctx->batch_off[4]);
            bs.part_cnt = reinterpret_cast<int32_t*>(sc + ctx->batch_off[5]);
            bs.bar = reinterpret_cast<uint32_t*>(sc + ctx->batch_off[6]);
            bs.lsize = reinterpret_cast<int32_t*>(sc + ctx->batch_off[7]);
            bs.blockmin = reinterpret_cast<uint4*>(sc + ctx->batch_off[8]);
            bs.nearq = reinterpret_cast<int2*>(sc + ctx->batch_off[9]);
            const bool near_on = p.exact != 0 && ctx->near_opt != 0 && ctx->near_meta != nullptr && ctx->horizon >= 0.0;
            bs.near_meta = near_on ? ctx->near_meta : nullptr;
            bs.near_pool = ctx->near_pool;
            bs.near_pool_cap = ctx->near_pool_cap;
            bs.near_cursor = ctx->near_cursor;
            bs.slot_of_key = ctx->slot_of_key;
            bs.xfar = ctx->xfar ? ctx->xfar + static_cast<size_t>(v) * kMaxBatch : nullptr;
            bs.cen = ctx->cen;  // (virtual ranks share one centroid store: every rank writes the same values)
            bs.ldc = ctx->ldc;
            bs.xq = ctx->xq ? ctx->xq + static_cast<size_t>(v) * static_cast<size_t>(ctx->xq_cap) : nullptr;
            bs.xq_cap = ctx->xq_cap;
            bs.xres = ctx->xres;
            bs.xhit = ctx->xhit ? ctx->xhit + static_cast<size_t>(v) * kMaxBatch : nullptr;
            bs.xqm = ctx->xqm ? ctx->xqm + static_cast<size_t>(v) * kMaxBatch * kXResCap : nullptr;
            // scratch of a launch: counters and the barrier at zero
            IC_CUDA(cudaMemsetAsync(bs.counters, 0, 3 * 4 * 4, ctx->stream));
            IC_CUDA(cudaMemsetAsync(bs.part_cnt, 0, static_cast<size_t>(kBatchMaxDry) * 4, ctx->stream));
            IC_CUDA(cudaMemsetAsync(bs.bar, 0, 256, ctx->stream));
            IC_CUDA(cudaMemsetAsync(bs.ctl + CTL_DONE, 0, sizeof(int32_t), ctx->stream));
            if (bs.n_ranks > 1) {
                // this rank's exchange-box slots: no candidates, no minima (the flags keep counting up across launches)
                uint8_t* xb = bs.xbox[bs.rank];
                if (ctx->shard_world > 1) xb = ctx->rankbox + merge_loop_rankbox_bytes();  // (own box through the local mapping)
                IC_CUDA(cudaMemsetAsync(xb + kBatchXAccum, 0xFF, 48, ctx->stream));
                IC_CUDA(cudaMemsetAsync(xb + kBatchXAccum + 48, 0, 16, ctx->stream));
            }
            merge_batch_fill_windows(&bs);
        }
        if (ctx->xhit) IC_CUDA(cudaMemsetAsync(ctx->xhit, 0, sizeof(int32_t) * kMaxBatch * VR, ctx->stream));
        if (ctx->xfar) IC_CUDA(cudaMemsetAsync(ctx->xfar, 0, sizeof(int32_t) * kMaxBatch * VR, ctx->stream));
        if (ctx->shard_world > 1) {  // all ranks line up: nobody starts before every box is reset
            IC_CUDA(launch_rank_barrier(ctx->peer_box, ctx->shard_world, ctx->shard_rank, ctx->barrier_seq, ctx->stream));
            ctx->stats.kernel_launches += 1;
        }
        IC_CUDA(cudaEventRecord(ctx->ev[10], ctx->stream));
        if (VR > 1) {
            IC_CUDA(cudaMemcpyAsync(ctx->vstates, states.data(), sizeof(BatchState) * static_cast<size_t>(VR), cudaMemcpyHostToDevice, ctx->stream));
            IC_CUDA(cudaStreamSynchronize(ctx->stream));  // (the host vector goes out of scope; test hook)
            IC_CUDA(launch_merge_batch_virtual(ctx->vstates, VR, ctx->n_cur, p, std::max(1, std::min(ctx->batch_grid, ctx->num_sms / VR)), ctx->stream));
        } else {
            IC_CUDA(launch_merge_batch(states[0], p, ctx->batch_grid, ctx->stream));
        }
        IC_CUDA(cudaEventRecord(ctx->ev[11], ctx->stream));
        ctx->loop_launch_pending = true;
        ++ctx->loop_launches;
        ++ctx->loop_launches_run;
        ctx->stats.kernel_launches += 1;
        IC_CUDA(cudaMemcpyAsync(ctx->h_ctl, ctx->ctl, sizeof(ctx->h_ctl), cudaMemcpyDeviceToHost, ctx->stream));
        if (ctx->profile_loop)
            IC_CUDA(cudaMemcpyAsync(ctx->h_prof, ctx->prof, sizeof(ctx->h_prof), cudaMemcpyDeviceToHost, ctx->stream));
        ctx->trace_on_host = false;
        return IC_OK;
    }
    ctx->loop_gen = ctx->loop_gen % 4095u + 1u;
    st.gen = ctx->loop_gen;
    if (ctx->loop_gen == 1u && ctx->loop_launches > 0) {  // generation wrapped: forget every old tag
        const int NL = n_local(ctx);
        IC_CUDA(cudaMemsetAsync(ctx->records, 0, merge_loop_records_bytes(ctx->loop_grid) * NL, ctx->stream));
        IC_CUDA(cudaMemsetAsync(ctx->partials, 0, merge_loop_partials_bytes(ctx->loop_grid) * NL, ctx->stream));
        IC_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(ctx->rankbox) + 256, 0, merge_loop_rankbox_bytes() * NL - 256,
                                ctx->stream));
    }
    for (int v = 0; v < n_local(ctx); ++v)
        IC_CUDA(cudaMemsetAsync(ctx->ctl + kCtlWords * v + CTL_DONE, 0, sizeof(int32_t), ctx->stream));
    if (ctx->shard_world > 1) {
        // all ranks have left their previous launch (and cleared what they had to) before anyone talks
        ++ctx->barrier_seq;
        IC_CUDA(launch_rank_barrier(ctx->peer_box, ctx->shard_world, ctx->shard_rank, ctx->barrier_seq, ctx->stream));
        ctx->stats.kernel_launches += 1;
    }
    IC_CUDA(launch_merge_loop(st, p, ctx->loop_grid, ctx->loop_replica, ctx->stream));
    ++ctx->loop_launches;
    ctx->stats.kernel_launches += 1;
    IC_CUDA(cudaMemcpyAsync(ctx->h_ctl, ctx->ctl, sizeof(ctx->h_ctl), cudaMemcpyDeviceToHost, ctx->stream));
    if (st.prof)
        IC_CUDA(cudaMemcpyAsync(ctx->h_prof, st.prof, sizeof(ctx->h_prof), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->trace_on_host = false;
    return IC_OK;
}

// Pairs of the resident rows whose stored value lies in (lo, hi] get the reference's own value (refine.cu).  Rows are
// swept in one go; if the queue overflows, in row chunks sized from the count.
int refine_band(ic_ctx* ctx, double lo, double hi, int32_t min_row_key) {
    Nvtx range("ic refine (horizon sweep, reference arithmetic)");
    const double t0 = now_ms();
    RefineArgs a{};
    a.dm = ctx->dm_cur;
    a.ld = ctx->ld_cur;
    a.n_slots = static_cast<int32_t>(ctx->n_cur);
    a.mirror_key = ctx->mirror_key;
    a.rows_per_rank = ctx->shard_world > 1 ? static_cast<int32_t>(rows_per_rank_cur(ctx)) : 0x40000000;
    for (int q = 0; q < kMaxRanks; ++q)
        a.dm_rank[q] = (ctx->shard_world > 1 && q < ctx->shard_world) ? rank_block(ctx, q) : ctx->dm_cur;
    a.r_lo = static_cast<int32_t>(row_begin_cur(ctx));
    a.r_hi = static_cast<int32_t>(row_end_cur(ctx));  // (after a compaction: n_cur dense slots)
    a.ks = ctx->ks;
    a.gkey = ctx->gkey;
    a.cen = ctx->cen;
    a.ldc = ctx->ldc;
    a.lo = lo;
    a.hi = hi;
    a.min_row_key = min_row_key;
    a.lower_only = ctx->n_merges == 0 ? 1 : 0;
    a.q = ctx->rq;
    a.cap = ctx->rq_cap;
    a.cnt = ctx->rq_cnt;
    a.ctl = ctx->ctl;
    a.nn_more = ctx->nn_more;
    a.eps_filter = static_cast<float>(ctx->eps_filter);
    a.abs_slack = static_cast<float>(ctx->abs_slack);
    int64_t r = a.r_lo, step = std::max<int64_t>(1, a.r_hi - a.r_lo);
    while (r < a.r_hi) {
        a.row0 = static_cast<int32_t>(r);
        a.row1 = static_cast<int32_t>(std::min<int64_t>(a.r_hi, r + step));
        int32_t cnt = 0;
        IC_CUDA(cudaMemsetAsync(ctx->rq_cnt, 0, sizeof(int32_t) * 4, ctx->stream));
        IC_CUDA(launch_refine_collect(a, ctx->stream));
        IC_CUDA(cudaMemcpyAsync(&cnt, ctx->rq_cnt, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        IC_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->stats.kernel_launches += 1;
        if (cnt > a.cap) {
            if (step == 1) return fail(ctx, IC_ERR_INTERNAL, "refine: one row holds more pairs than the queue");
            step = std::max<int64_t>(1, step * a.cap / cnt / 2);
            continue;
        }
        if (cnt > 0) {
            IC_CUDA(launch_refine_eval(a, ctx->num_sms, ctx->stream));
            ctx->stats.kernel_launches += 1;
        }
        r = a.row1;
    }
    IC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->ms_refine += now_ms() - t0;
    return IC_OK;
}

// Near lists of the resident rows for the current horizon (near.cu): one sweep of the live matrix.
int build_near(ic_ctx* ctx, bool mark_dry) {
    if (!ctx->near_meta || !ctx->near_opt) {  // no near lists: stale bounds are replaced by full row scans
        if (mark_dry) {
            IC_CUDA(launch_mark_rows_dry(ctx->gkey, ctx->nn_more, static_cast<int32_t>(row_begin_cur(ctx)),
                                         static_cast<int32_t>(row_end_cur(ctx)), ctx->stream));
            ctx->stats.kernel_launches += 1;
        }
        return IC_OK;
    }
    Nvtx range("ic near lists (one sweep)");
    const double t0 = now_ms();
    NearArgs a{};
    a.dm = ctx->dm_cur;
    a.ld = ctx->ld_cur;
    a.n_slots = static_cast<int32_t>(ctx->n_cur);
    a.r_lo = static_cast<int32_t>(row_begin_cur(ctx));
    a.r_hi = static_cast<int32_t>(row_end_cur(ctx));
    a.gkey = ctx->gkey;
    a.order_key = ctx->order_key;
    a.horizon = ctx->horizon;
    a.meta = ctx->near_meta;
    a.pool = ctx->near_pool;
    a.pool_cap = ctx->near_pool_cap;
    a.cursor = ctx->near_cursor;
    IC_CUDA(launch_near_build(a, ctx->stream));
    ctx->stats.kernel_launches += 1;
    if (mark_dry) {  // bounds at the old horizon are stale: every live row re-selects its list (cheap: from its near list)
        IC_CUDA(launch_mark_rows_dry(ctx->gkey, ctx->nn_more, a.r_lo, a.r_hi, ctx->stream));
        ctx->stats.kernel_launches += 1;
    }
    IC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->ms_near += now_ms() - t0;
    return IC_OK;
}

// STOP_HORIZON: the smallest candidate H is above the range in which stored values are guaranteed to be the reference's.
// New horizon = factor * (H (1 + 2 eps) + 2 slack); the band between the old and the new one is re-evaluated.  The factor
// escalates while a horizon buys fewer than 64 merges (isolated tiny distances below the bulk).
int raise_horizon(ic_ctx* ctx) {
    float h = 0.0f;
    const uint32_t hb = static_cast<uint32_t>(ctx->h_ctl[CTL_NEXT_DIST]);
    std::memcpy(&h, &hb, 4);
    const bool first = ctx->horizon < 0.0;
    if (first && ctx->abs_slack_opt < 0.0 && !ctx->dm_is_reference && ctx->norms && ctx->gram_mode_used != IC_GRAM_EXACT_FP32 &&
        ctx->gram_mode_used >= 0) {
        // absolute error of a tensor-core Gram value: 2^-20 of the largest centred squared norm
        std::vector<double> hn(static_cast<size_t>(ctx->n));
        IC_CUDA(cudaMemcpyAsync(hn.data(), ctx->norms, sizeof(double) * hn.size(), cudaMemcpyDeviceToHost, ctx->stream));
        IC_CUDA(cudaStreamSynchronize(ctx->stream));
        double mx = 0.0;
        for (double v : hn) mx = std::max(mx, v);
        ctx->abs_slack = mx * (1.0 / 1048576.0);
    }
    if (!first && ctx->n_merges - ctx->merges_at_raise < 64)
        ctx->hz_factor_cur = std::min(ctx->hz_factor_cur * ctx->hz_factor_cur, 1e6);
    else
        ctx->hz_factor_cur = (first && ctx->hz_factor_first > 1.0) ? ctx->hz_factor_first : ctx->hz_factor;
    const double base = std::max(static_cast<double>(h), ctx->horizon);
    const double hi = ctx->hz_factor_cur * (base * (1.0 + 2.0 * ctx->eps_filter) + 2.0 * ctx->abs_slack);
    if (!(first && ctx->dm_is_reference)) {
        const int rc = refine_band(ctx, first ? -1.0 : ctx->horizon, hi, 0);
        if (rc != IC_OK) return rc;
        // the partner lists of the rows whose values changed are stale.  With near lists every live row simply re-selects its
        // list from them (below: build_near marks all rows); without, the first sweep kernel rebuilds every list from the
        // whole rows while keys are still the slot indices (later: the rows marked by refine.cu are scanned by the loop)
        if (ctx->n_merges == 0 && !(ctx->near_meta && ctx->near_opt)) {
            IC_CUDA(launch_nn_sweep(ctx->dm, row_begin(ctx), row_end(ctx), ctx->ld, ctx->nn, ctx->nn_more, ctx->stream));
            ctx->stats.kernel_launches += 1;
        }
    }
    ctx->horizon = hi;
    // (first horizon over an initial matrix that already holds the reference's values: the lists of the first sweep stand)
    const bool lists_stand = ctx->n_merges == 0 && (!(ctx->near_meta && ctx->near_opt) || (first && ctx->dm_is_reference));
    ctx->merges_at_raise = ctx->n_merges;
    ++ctx->n_raises;
    return build_near(ctx, !lists_stand);
}

// K4 (compact.cu): renumber the live clusters densely in key order, move the matrix into the other buffer (both
// triangles), permute the per-slot state.  Host-driven: the loop kernel stopped with STOP_COMPACT.
int do_compact(ic_ctx* ctx) {
    Nvtx range("ic K4 compaction");
    const double t0 = now_ms();
    const int P = n_ranks(ctx), NL = n_local(ctx);
    const bool real = ctx->shard_world > 1;
    const int32_t n_old = static_cast<int32_t>(ctx->n_cur), n_new = ctx->n_live, n_new4 = (n_new + 3) & ~3;
    const int64_t ld_new = round_up(n_new, 32);
    const int32_t c_old = static_cast<int32_t>(rows_per_rank_cur(ctx));
    const int32_t c_new = static_cast<int32_t>(merge_loop_rows_per_rank(n_new, P));
    const bool to_b = ctx->dm_cur == ctx->dm;
    float* target = to_b ? ctx->dm_b : ctx->dm;
    const size_t cap = to_b ? ctx->dm_b_floats
                            : static_cast<size_t>(real ? rows_per_rank(ctx) : ctx->n) * static_cast<size_t>(ctx->ld);
    const size_t rows_here = real ? static_cast<size_t>(c_new) : static_cast<size_t>(n_new4);
    if (rows_here * static_cast<size_t>(ld_new) > cap) return fail(ctx, IC_ERR_INTERNAL, "compaction target buffer too small");
    if (real && !ctx->peers_open) return fail(ctx, IC_ERR_STATE, "sharded context: ic_shard_connect has not run");
    IC_CUDA(launch_compact_map(ctx->ks, n_old, ctx->keymap, static_cast<int32_t>(ctx->n + ctx->n_merges), ctx->newslot,
                               ctx->oldslot, n_new4, ctx->nlive_dev, ctx->stream));
    const size_t n4s = (static_cast<size_t>(ctx->n) + 3) / 4 * 4;
    CompactArgs a{};
    a.n_old = n_old;
    a.n_new = n_new;
    a.n_new4 = n_new4;
    a.oldslot = ctx->oldslot;
    a.newslot = ctx->newslot;
    a.nn_old = ctx->nn;
    a.nn_more_old = ctx->nn_more;
    a.nn_new = ctx->nn_b;
    a.nn_more_new = ctx->nn_more_b;
    a.my_rank = real ? ctx->shard_rank : -1;  // real shards: the partner list of a row that changes owner is rebuilt
    a.order_key_old = ctx->order_key;
    a.near_meta_old = ctx->near_meta;
    a.near_meta_new = ctx->near_meta_b;
    a.slot_of_key = ctx->slot_of_key;
    for (int q = 0; q < kMaxRanks; ++q) {
        a.dm_old[q] = q < P ? rank_block(ctx, q) : nullptr;
        a.dm_new_rank[q] = q >= P ? nullptr
                           : real ? static_cast<float*>(to_b ? ctx->peer_dm_b[q] : ctx->peer_dm[q])
                                  : target + static_cast<int64_t>(q) * c_new * ld_new;
    }
    a.rows_per_rank_old = c_old;
    a.rows_per_rank_new = c_new;
    a.ld_old = ctx->ld_cur;
    a.dm_new = target;
    a.row_base_new = real ? std::min(n_new, ctx->shard_rank * c_new) : 0;
    a.row0 = a.row_base_new;
    a.row1 = real ? std::min(n_new, a.row_base_new + c_new) : n_new;
    a.ld_new = ld_new;
    for (int v = 0; v < NL; ++v) {  // every (virtual) rank's replica of the slot table; the partner lists once
        a.ks_old = ctx->ks + static_cast<size_t>(v) * (n4s + 4);
        a.ks_new = ctx->ks_b + static_cast<size_t>(v) * (n4s + 4);
        a.gkey_new = ctx->gkey_b + static_cast<size_t>(v) * n4s;
        CompactArgs av = a;
        if (v > 0) av.nn_new = nullptr;
        IC_CUDA(launch_compact_state(av, ctx->stream));
    }
    a.ks_old = ctx->ks;
    if (!real && ctx->compact_tiles) {  // one process: old and new matrix are contiguous -- both triangles in one pass
        IC_CUDA(launch_compact_tiles(a, ctx->stream));
    } else {
        IC_CUDA(launch_compact_rows(a, ctx->stream));
        if (real) {  // every rank's lower triangle is complete before anyone mirrors it
            ++ctx->barrier_seq;
            IC_CUDA(launch_rank_barrier(ctx->peer_box, ctx->shard_world, ctx->shard_rank, ctx->barrier_seq, ctx->stream));
        }
        IC_CUDA(launch_mirror_lower(a, ctx->stream));
    }
    // scatter + rank, one state kernel per (virtual) rank, then tiles | rows (+ rank barrier) + mirror
    ctx->stats.kernel_launches += 2 + NL + ((!real && ctx->compact_tiles) ? 1 : (real ? 3 : 2));
    int32_t found = 0;
    IC_CUDA(cudaMemcpyAsync(&found, ctx->nlive_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    IC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (found != n_new) return fail(ctx, IC_ERR_INTERNAL, "compaction found " + std::to_string(found) + " live clusters, expected " + std::to_string(n_new));
    std::swap(ctx->ks, ctx->ks_b);
    std::swap(ctx->gkey, ctx->gkey_b);
    std::swap(ctx->nn, ctx->nn_b);
    std::swap(ctx->nn_more, ctx->nn_more_b);
    std::swap(ctx->near_meta, ctx->near_meta_b);
    ctx->dm_cur = target;
    ctx->n_cur = n_new;
    ctx->ld_cur = ld_new;
    ctx->order_key = ctx->mirror_key = static_cast<int32_t>(ctx->n + ctx->n_merges);
    ++ctx->n_compactions;
    ctx->ms_compact += now_ms() - t0;
    return IC_OK;
}

constexpr int kRestart = 1;  // sync_loop_result: an optimistic batch failed its check (STOP_ORDER)

int run_loop(ic_ctx* ctx, int64_t n_target, int64_t max_size, int64_t max_merges) {
    ctx->loop_n_target = n_target;
    ctx->loop_max_size = max_size;
    ctx->loop_max_merges = max_merges;
    ctx->loop_merges_at_start = ctx->n_merges;
    return enqueue_loop(ctx, n_target, max_size, max_merges);
}

float ev_ms(cudaEvent_t a, cudaEvent_t b);

int sync_loop_result(ic_ctx* ctx) {
    Nvtx range("ic K3b merge loop (sync, horizon raises, compactions, relaunches)");
    for (;;) {
        IC_CUDA(cudaStreamSynchronize(ctx->stream));
        if (ctx->loop_launch_pending) {
            ctx->ms_loop_kernel += ev_ms(ctx->ev[10], ctx->ev[11]);
            ctx->loop_launch_pending = false;
        }
        if (ctx->h_ctl[CTL_ERROR] != 0)  // (4: a grid / cross-rank barrier was given up -- the kernel left without hanging)
            return fail(ctx, IC_ERR_INTERNAL, "merge loop protocol error " + std::to_string(ctx->h_ctl[CTL_ERROR]));
        if (ctx->h_ctl[CTL_DONE] != 1) return fail(ctx, IC_ERR_INTERNAL, "merge loop did not complete");
        ctx->n_live = ctx->h_ctl[CTL_N_LIVE];
        ctx->n_merges = ctx->h_ctl[CTL_N_MERGES];
        ctx->exhausted = ctx->h_ctl[CTL_EXHAUSTED];
        const int stop = ctx->h_ctl[CTL_STOP];
        if (stop == STOP_HORIZON) {  // the minimum reached the horizon: raise it, re-evaluate the band (refine.cu)
            const int rc = raise_horizon(ctx);
            if (rc != IC_OK) return rc;
        } else if (stop == STOP_XQ) {  // the exact-evaluation queue of the last iteration overflowed: redo its rows
            int rc = refine_band(ctx, -1.0, ctx->horizon, ctx->h_ctl[CTL_XQ_FIRST_KEY]);
            if (rc != IC_OK) return rc;
            for (int v = 0; v < n_local(ctx); ++v)
                IC_CUDA(cudaMemsetAsync(ctx->ctl + kCtlWords * v + CTL_XQ_OVERFLOW, 0, sizeof(int32_t), ctx->stream));
        } else if (stop == STOP_COMPACT) {
            const int rc = do_compact(ctx);
            if (rc != IC_OK) return rc;
        } else if (stop == STOP_ORDER) {
            return kRestart;
        } else if (stop != STOP_EPOCHS) {
            return IC_OK;
        }
        int64_t left = ctx->loop_max_merges;
        if (left >= 0) left -= ctx->n_merges - ctx->loop_merges_at_start;
        const int rc = enqueue_loop(ctx, ctx->loop_n_target, ctx->loop_max_size, left < 0 ? -1 : left);
        if (rc != IC_OK) return rc;
    }
}

int fetch_trace(ic_ctx* ctx) {
    if (ctx->trace_on_host) return IC_OK;
    Nvtx range("ic merge trace D2H");
    const size_t m = static_cast<size_t>(ctx->n_merges);
    ctx->h_key_hi.resize(m);
    ctx->h_key_lo.resize(m);
    ctx->h_size.resize(m);
    ctx->h_dist.resize(m);
    ctx->h_gap.resize(m);
    if (m) {
        IC_CUDA(cudaMemcpyAsync(ctx->h_key_hi.data(), ctx->tr_key_hi, 4 * m, cudaMemcpyDeviceToHost, ctx->stream));
        IC_CUDA(cudaMemcpyAsync(ctx->h_key_lo.data(), ctx->tr_key_lo, 4 * m, cudaMemcpyDeviceToHost, ctx->stream));
        IC_CUDA(cudaMemcpyAsync(ctx->h_size.data(), ctx->tr_size, 4 * m, cudaMemcpyDeviceToHost, ctx->stream));
        IC_CUDA(cudaMemcpyAsync(ctx->h_dist.data(), ctx->tr_dist, 4 * m, cudaMemcpyDeviceToHost, ctx->stream));
        IC_CUDA(cudaMemcpyAsync(ctx->h_gap.data(), ctx->tr_gap, 4 * m, cudaMemcpyDeviceToHost, ctx->stream));
        IC_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    ctx->stats.d2h_bytes += static_cast<int64_t>(20 * m + sizeof(ctx->h_ctl));
    ctx->trace_on_host = true;
    return IC_OK;
}

// Output assembly (clustering.go:265-280).  Slice order == key order; members of a merged
// cluster are members(hi) ++ members(lo) (clustering.go:31,237); clusters below minSize are
// skipped without consuming an id (:268-271).
int assemble(ic_ctx* ctx, int64_t min_size, int64_t max_size, int32_t* offsets, int32_t* members, int32_t* n_out) {
    Nvtx range("ic output assembly (host)");
    const int64_t n = ctx->n, m = ctx->n_merges;
    std::vector<uint8_t> consumed(static_cast<size_t>(n + m), 0);
    for (int64_t t = 0; t < m; ++t) {
        const int32_t khi = ctx->h_key_hi[t], klo = ctx->h_key_lo[t];
        if (khi < 0 || klo < 0 || khi >= n + t || klo >= n + t || khi <= klo || consumed[khi] || consumed[klo])
            return fail(ctx, IC_ERR_INTERNAL, "corrupt merge trace at merge " + std::to_string(t));
        consumed[khi] = consumed[klo] = 1;
    }
    std::vector<int32_t> stack;
    int32_t cid = 0;
    int64_t pos = 0;
    offsets[0] = 0;
    for (int64_t k = 0; k < n + m; ++k) {
        if (consumed[k]) continue;
        const int64_t size = k < n ? 1 : ctx->h_size[k - n];
        // the reference would try to split an oversized cluster (clustering.go:249-262); no merge
        // can create one (:228), so this is an invariant check, not a code path
        if (max_size > 0 && size > max_size) return fail(ctx, IC_ERR_INTERNAL, "cluster exceeds maxSize");
        if (size < min_size) continue;
        stack.clear();
        stack.push_back(static_cast<int32_t>(k));
        while (!stack.empty()) {
            const int32_t kk = stack.back();
            stack.pop_back();
            if (kk < n) {
                if (pos >= n) return fail(ctx, IC_ERR_INTERNAL, "member overflow");
                members[pos++] = kk;
            } else {
                stack.push_back(ctx->h_key_lo[kk - n]);
                stack.push_back(ctx->h_key_hi[kk - n]);
            }
        }
        offsets[++cid] = static_cast<int32_t>(pos);
    }
    *n_out = cid;
    ctx->stats.n_out = cid;
    return IC_OK;
}

int validate_sizes(ic_ctx* ctx, int64_t n, int64_t d, int64_t ldx) {
    if (n < 0 || d < 0 || (n > 0 && d > 0 && ldx < d)) return fail(ctx, IC_ERR_BAD_ARG, "bad n / d / ldx");
    if (n > 0x3FFFFFFF) return fail(ctx, IC_ERR_BAD_ARG, "too many items");
    return IC_OK;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int load_common(ic_ctx* ctx, const float* x, int64_t n, int64_t d, int64_t ldx, cudaMemcpyKind kind) {
    Nvtx range("ic upload X");
    int rc = validate_sizes(ctx, n, d, ldx);
    if (rc != IC_OK) return rc;
    if (n > 0 && d > 0 && !x) return fail(ctx, IC_ERR_BAD_ARG, "x is NULL");
    IC_CUDA(cudaSetDevice(ctx->device));
    rc = alloc_problem(ctx, n, d);
    if (rc != IC_OK) return rc;
    std::memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->stats.n_items = n;
    ctx->stats.dim = d;
    ctx->stats.near_tie_tol = static_cast<float>(ctx->near_tie_tol);
    ctx->stats.matrix_bytes = static_cast<int64_t>(sizeof(float)) * n * ctx->ld;
    IC_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (n > 0 && d > 0) {
        if (ldx == d)
            IC_CUDA(cudaMemcpyAsync(ctx->x, x, sizeof(float) * n * d, kind, ctx->stream));
        else
            IC_CUDA(cudaMemcpy2DAsync(ctx->x, sizeof(float) * d, x, sizeof(float) * ldx, sizeof(float) * d, n, kind,
                                      ctx->stream));
        if (kind == cudaMemcpyHostToDevice) ctx->stats.h2d_bytes += static_cast<int64_t>(sizeof(float)) * n * d;
    }
    IC_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    ctx->loaded = true;
    return IC_OK;
}

int initial_distances(ic_ctx* ctx, int mode, int64_t max_size) {
    if (!ctx->loaded) return fail(ctx, IC_ERR_STATE, "no matrix loaded");
    if (mode != IC_GRAM_TCGEN05_3XTF32 && mode != IC_GRAM_EXACT_FP32 && mode != IC_GRAM_TCGEN05_I8)
        return fail(ctx, IC_ERR_BAD_ARG, "bad gram mode");
    IC_CUDA(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (mode == IC_GRAM_TCGEN05_3XTF32) {
        const int rc = do_prep(ctx);
        if (rc != IC_OK) return rc;
    } else if (mode == IC_GRAM_TCGEN05_I8) {
        const int rc = do_prep_i8(ctx);
        if (rc != IC_OK) return rc;
    }
    IC_CUDA(cudaEventRecord(ctx->ev[3], ctx->stream));
    ctx->gmin_valid = false;
    if (max_size >= 2) {
        const int rc = do_gram(ctx, mode);
        if (rc != IC_OK) return rc;
    } else {
        ctx->dm_lower_only = false;
        // 1 + 1 > maxSize: every pair is inadmissible from the start (clustering.go:228)
        IC_CUDA(launch_fill(ctx->dm, (row_end(ctx) - row_begin(ctx)) * ctx->ld, INFINITY, ctx->stream));
        ctx->stats.kernel_launches += 1;
    }
    if (ctx->mirror_init && ctx->dm_lower_only && ctx->n > 1 && max_size >= 2 && (ctx->shard_world <= 1 || ctx->peers_open)) {
        // fill the upper triangle: every row then holds all its partners (fewer gathers in the loop's first epoch).  Sharded:
        // the transposed tiles are read from their owners, once every rank's K1 is complete
        const bool real = ctx->shard_world > 1;
        CompactArgs a{};
        a.n_new = static_cast<int32_t>(ctx->n);
        a.dm_new = ctx->dm;
        a.rows_per_rank_new = real ? static_cast<int32_t>(rows_per_rank(ctx)) : 0x40000000;
        for (int q = 0; q < kMaxRanks; ++q) a.dm_new_rank[q] = real ? static_cast<float*>(ctx->peer_dm[q]) : ctx->dm;
        a.row_base_new = static_cast<int32_t>(row_begin(ctx));
        a.row0 = a.row_base_new;
        a.row1 = static_cast<int32_t>(row_end(ctx));
        a.ld_new = ctx->ld;
        if (real) {
            ++ctx->barrier_seq;
            IC_CUDA(launch_rank_barrier(ctx->peer_box, ctx->shard_world, ctx->shard_rank, ctx->barrier_seq, ctx->stream));
        }
        if (!real) {
            if (!ctx->gmin_dev) IC_CUDA(cudaMalloc(&ctx->gmin_dev, sizeof(uint32_t) * 4));
            IC_CUDA(cudaMemsetAsync(ctx->gmin_dev, 0xFF, sizeof(uint32_t) * 4, ctx->stream));
            a.gmin = ctx->gmin_dev;
        }
        IC_CUDA(launch_mirror_lower(a, ctx->stream));
        ctx->stats.kernel_launches += 1;
        ctx->dm_lower_only = false;
        ctx->gmin_valid = !real;
    }
    IC_CUDA(cudaEventRecord(ctx->ev[4], ctx->stream));
    ctx->have_dm = true;
    ctx->have_nn = false;
    reset_epoch(ctx);
    ctx->gram_mode_used = mode;
    ctx->stats.gram_mode = mode;
    ctx->exact_on = ctx->exact_opt != 0 && ctx->cen != nullptr;
    ctx->dm_is_reference = mode == IC_GRAM_EXACT_FP32 || max_size < 2;
    return IC_OK;
}

int nn_init(ic_ctx* ctx) {
    Nvtx range("ic K2 first nearest-neighbour sweep");
    if (!ctx->have_dm) return fail(ctx, IC_ERR_STATE, "no distance matrix");
    int rc = init_loop_state(ctx);
    if (rc != IC_OK) return rc;
    IC_CUDA(launch_nn_sweep(ctx->dm, row_begin(ctx), row_end(ctx), ctx->ld, ctx->nn, ctx->nn_more, ctx->stream));
    ctx->stats.kernel_launches += 1;
    IC_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));
    ctx->have_nn = true;
    return IC_OK;
}

void fill_stats(ic_ctx* ctx) {
    ic_stats& s = ctx->stats;
    s.n_target = static_cast<int32_t>(ctx->n_target);
    s.n_merges = ctx->n_merges;
    s.n_final = ctx->n_live;
    s.exhausted = ctx->exhausted;
    s.n_near_ties = ctx->h_ctl[CTL_NEAR_TIES];
    s.n_rescans = ctx->h_ctl[CTL_RESCANS];
    s.loop_mode = ctx->loop_mode_used > 0 ? 1 : 0;
    s.exact = (ctx->exact_on && use_batch(ctx) && ctx->cen) ? 1 : 0;
    s.n_horizon_raises = ctx->n_raises;
    s.n_exact = ctx->h_ctl[CTL_N_EXACT];
    s.n_filter_viol = ctx->h_ctl[CTL_FILTER_VIOL];
    s.n_order_viol = ctx->h_ctl[CTL_ORDER_VIOL];
    s.n_cut = ctx->h_ctl[CTL_N_CUT];
    {
        const uint32_t eb = static_cast<uint32_t>(ctx->h_ctl[CTL_FILTER_MAXERR]);
        std::memcpy(&s.filter_max_err, &eb, 4);
    }
    s.horizon = ctx->horizon;
    s.ms_refine = static_cast<float>(ctx->ms_refine + ctx->ms_near);  // horizon sweeps: re-evaluation + near lists
    s.n_restarts = ctx->n_restarts;
    s.ms_loop_kernel = static_cast<float>(ctx->ms_loop_kernel);
    s.loop_launches = static_cast<int32_t>(ctx->loop_launches_run);
    s.n_compactions = ctx->n_compactions;
    s.ms_compact = static_cast<float>(ctx->ms_compact);
    s.n_iterations = s.loop_mode ? ctx->h_ctl[CTL_ITERS] : ctx->n_merges + ctx->h_ctl[CTL_BUBBLES];
}

// Start of a resident run.  With reference arithmetic and near lists on (one GPU, batched loop) the first launch of the loop
// only ever reports the global minimum -- the first horizon is set from it, the band below it re-evaluated, and every row's
// list re-selected from its near list.  So the first sweep (K2: 4 bytes per pair) is skipped: the mirror pass after K1 has
// collected the minimum, the lists start empty and "dry", and the horizon is raised right away.  Anything else: K2.
int fast_start(ic_ctx* ctx, int64_t n_target) {
    const bool applies = ctx->fast_start && ctx->gmin_valid && ctx->exact_on && ctx->cen != nullptr && !ctx->dm_is_reference &&
                         use_batch(ctx) && ctx->shard_world <= 1 && ctx->vranks <= 1 && ctx->near_meta != nullptr &&
                         ctx->near_opt != 0 && ctx->n > n_target;
    if (!applies) return nn_init(ctx);
    uint32_t gmin = 0xFFFFFFFFu;
    IC_CUDA(cudaMemcpyAsync(&gmin, ctx->gmin_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    IC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (gmin >= 0x7F7FFFFFu) return nn_init(ctx);  // (>= MaxFloat32) nothing selectable: the loop reports exhaustion the usual way
    Nvtx range("ic start without the first sweep");
    int rc = init_loop_state(ctx);
    if (rc != IC_OK) return rc;
    IC_CUDA(launch_init_lists_dry(ctx->nn, ctx->nn_more, static_cast<int32_t>(ctx->n), ctx->stream));
    ctx->stats.kernel_launches += 1;
    IC_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));
    ctx->have_nn = true;
    ctx->h_ctl[CTL_NEXT_DIST] = static_cast<int32_t>(gmin);
    return raise_horizon(ctx);
}

int run_resident(ic_ctx* ctx, int64_t min_size, int64_t max_size, int32_t* offsets, int32_t* members,
                 int32_t* n_clusters, ic_stats* stats, double t_wall0) {
    if (!ctx->loaded) return fail(ctx, IC_ERR_STATE, "no matrix loaded");
    if (!offsets || !members || !n_clusters) return fail(ctx, IC_ERR_BAD_ARG, "NULL output");
    int64_t n_target = 0;
    int rc = ic_optimal_clusters(ctx->n, min_size, max_size, &n_target);
    if (rc != IC_OK) return fail(ctx, rc, "cluster size constraints cannot be satisfied");
    ctx->n_target = n_target;
    ctx->delta_cut_cur = ctx->delta_cut;
    ctx->n_restarts = 0;
    for (int attempt = 0;; ++attempt) {
        ctx->prepped = ctx->prepped_i8 = false;  // K0 is part of the path: redo it on every run
        rc = initial_distances(ctx, ctx->gram_mode, max_size);
        if (rc != IC_OK) return rc;
        rc = fast_start(ctx, n_target);  // (K2 only if the start without it does not apply)
        if (rc != IC_OK) return rc;
        rc = run_loop(ctx, n_target, max_size, -1);
        if (rc != IC_OK) return rc;
        rc = sync_loop_result(ctx);  // (relaunches after a horizon raise: the loop's time includes the sweeps)
        if (rc == kRestart && attempt == 0 && ctx->delta_cut_cur < ctx->delta_cut_fallback) {
            // a pair created inside an optimistic batch came out below a later member of it (fp32 centroid distances
            // are reducible only up to rounding): start over, keeping every batch clear of its stopper
            ctx->delta_cut_cur = ctx->delta_cut_fallback;
            ++ctx->n_restarts;
            continue;
        }
        if (rc == kRestart) return fail(ctx, IC_ERR_INTERNAL, "batch order check failed even with delta_cut_fallback");
        if (rc != IC_OK) return rc;
        break;
    }
    IC_CUDA(cudaEventRecord(ctx->ev[6], ctx->stream));
    rc = fetch_trace(ctx);
    if (rc != IC_OK) return rc;
    IC_CUDA(cudaEventRecord(ctx->ev[7], ctx->stream));
    IC_CUDA(cudaEventSynchronize(ctx->ev[7]));
    const double th0 = now_ms();
    rc = assemble(ctx, min_size, max_size, offsets, members, n_clusters);
    if (rc != IC_OK) return rc;
    const double th1 = now_ms();
    fill_stats(ctx);
    ic_stats& s = ctx->stats;
    s.ms_h2d = ev_ms(ctx->ev[0], ctx->ev[1]);
    s.ms_prep = ev_ms(ctx->ev[2], ctx->ev[3]);
    s.ms_gram = ev_ms(ctx->ev[3], ctx->ev[4]);
    s.ms_nn_init = ev_ms(ctx->ev[4], ctx->ev[5]);
    s.ms_loop = ev_ms(ctx->ev[5], ctx->ev[6]);
    s.ms_d2h = ev_ms(ctx->ev[6], ctx->ev[7]);
    s.ms_host = static_cast<float>(th1 - th0);
    s.ms_total = static_cast<float>(th1 - t_wall0);
    if (stats) *stats = s;
    return IC_OK;
}

}  // namespace

extern "C" {

int ic_create(ic_ctx** out, int device) {
    if (!out) return IC_ERR_BAD_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return IC_ERR_CUDA;  // no CPU fallback
    if (device < 0 || device >= count) return IC_ERR_BAD_ARG;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return IC_ERR_CUDA;
    if (prop.major != 10) return IC_ERR_CUDA;  // sm_100a code only
    if (cudaSetDevice(device) != cudaSuccess) return IC_ERR_CUDA;
    ic_ctx* c = new ic_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return IC_ERR_CUDA;
    }
    for (auto& e : c->ev)
        if (cudaEventCreate(&e) != cudaSuccess) {
            ic_destroy(c);
            return IC_ERR_CUDA;
        }
    *out = c;
    return IC_OK;
}

void ic_destroy(ic_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    release_problem(ctx);
    for (auto& e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* ic_last_error(const ic_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context"; }

void* ic_pinned_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}

void ic_pinned_free(void* p) {
    if (p) cudaFreeHost(p);
}

int ic_set_option(ic_ctx* ctx, const char* name, double value) {
    if (!ctx || !name) return IC_ERR_BAD_ARG;
    const std::string k(name);
    if (k == "near_tie_tol")
        ctx->near_tie_tol = value;
    else if (k == "center")
        ctx->center = value != 0.0;
    else if (k == "gram_mode") {
        const int m = static_cast<int>(value);
        if (m != IC_GRAM_TCGEN05_3XTF32 && m != IC_GRAM_EXACT_FP32 && m != IC_GRAM_TCGEN05_I8)
            return fail(ctx, IC_ERR_BAD_ARG, "bad gram_mode");
        ctx->gram_mode = m;
    } else if (k == "loop_threads") {
        const int t = static_cast<int>(value);  // kept for compatibility: the loop kernel has 512 threads
        if (t != 0 && t != 512) return fail(ctx, IC_ERR_BAD_ARG, "loop_threads must be 0 or 512");
    } else if (k == "virtual_ranks") {
        const int r = static_cast<int>(value);
        if (r < 1 || r > kMaxRanks) return fail(ctx, IC_ERR_BAD_ARG, "virtual_ranks must be 1..8");
        if (ctx->shard_world > 1) return fail(ctx, IC_ERR_STATE, "virtual_ranks on a sharded context");
        ctx->vranks = r;
    } else if (k == "scan_every") {
        const int e = static_cast<int>(value);
        if (e < 1 || e > 64) return fail(ctx, IC_ERR_BAD_ARG, "scan_every must be 1..64");
        ctx->scan_every = e;
    } else if (k == "loop_mode") {
        const int m = static_cast<int>(value);
        if (m != 0 && m != 1) return fail(ctx, IC_ERR_BAD_ARG, "loop_mode must be 0 (sequential) or 1 (batched)");
        ctx->loop_mode = m;
    } else if (k == "gram_debug") {
        ctx->gram_debug = static_cast<int>(value);
    } else if (k == "loop_debug") {
        ctx->loop_debug = static_cast<int>(value);
    } else if (k == "no_replica") {
        ctx->no_replica = value != 0.0;
    } else if (k == "loop_blocks") {
        ctx->loop_blocks = static_cast<int>(value);
    } else if (k == "exact") {
        const int m = static_cast<int>(value);
        if (m < 0 || m > 2) return fail(ctx, IC_ERR_BAD_ARG, "exact must be 0, 1 or 2");
        ctx->exact_opt = m;
    } else if (k == "horizon_factor") {
        if (!(value > 1.0)) return fail(ctx, IC_ERR_BAD_ARG, "horizon_factor must be > 1");
        ctx->hz_factor = value;
    } else if (k == "horizon_factor_first") {
        if (!(value == 0.0 || value > 1.0)) return fail(ctx, IC_ERR_BAD_ARG, "horizon_factor_first must be 0 or > 1");
        ctx->hz_factor_first = value;
    } else if (k == "eps_filter") {
        if (!(value >= 0.0 && value < 0.1)) return fail(ctx, IC_ERR_BAD_ARG, "eps_filter must be in [0, 0.1)");
        ctx->eps_filter = value;
    } else if (k == "delta_cut") {
        if (!(value >= 0.0 && value < 0.1)) return fail(ctx, IC_ERR_BAD_ARG, "delta_cut must be in [0, 0.1)");
        ctx->delta_cut = value;
    } else if (k == "refill_at") {
        const int v = static_cast<int>(value);
        if (v < 1 || v > 2) return fail(ctx, IC_ERR_BAD_ARG, "refill_at must be 1 or 2");
        ctx->refill_at = v;
    } else if (k == "compact") {
        ctx->compact_opt = value != 0.0;
    } else if (k == "near_lists") {
        ctx->near_opt = value != 0.0;
    } else if (k == "compact_ratio") {
        if (!(value >= 0.25 && value <= 0.9)) return fail(ctx, IC_ERR_BAD_ARG, "compact_ratio must be in [0.25, 0.9]");
        ctx->compact_ratio = value;
    } else if (k == "fast_start") {
        ctx->fast_start = value != 0.0;
    } else if (k == "mirror_init") {
        ctx->mirror_init = value != 0.0;
    } else if (k == "compact_tiles") {
        ctx->compact_tiles = value != 0.0;
    } else if (k == "compact_min") {
        ctx->compact_min = std::max<int64_t>(64, static_cast<int64_t>(value));
    } else if (k == "delta_cut_fallback") {
        if (!(value >= 0.0 && value < 0.1)) return fail(ctx, IC_ERR_BAD_ARG, "delta_cut_fallback must be in [0, 0.1)");
        ctx->delta_cut_fallback = value;
    } else if (k == "abs_slack") {
        ctx->abs_slack_opt = value;
    } else if (k == "profile_loop") {
        ctx->profile_loop = value != 0.0;
    } else if (k == "gram_terms") {
        ctx->gram_terms = static_cast<int>(value);
    } else if (k == "verbose")
        ctx->verbose = static_cast<int>(value);
    else
        return fail(ctx, IC_ERR_BAD_ARG, "unknown option " + k);
    return IC_OK;
}

// CalculateOptimalClusters, clustering.go:168-186.
int ic_optimal_clusters(int64_t total_items, int64_t min_size, int64_t max_size, int64_t* out) {
    if (out) *out = 0;
    // Go divides by float64(0) here (+Inf -> implementation-defined int conversion); refuse instead
    if (min_size < 1 || max_size < 1 || total_items < 0 || !out) return IC_ERR_BAD_ARG;
    if (total_items < min_size) return IC_ERR_TOO_FEW;  // :169-171
    const int64_t lo = static_cast<int64_t>(std::ceil(static_cast<double>(total_items) / static_cast<double>(max_size)));
    const int64_t hi = static_cast<int64_t>(std::floor(static_cast<double>(total_items) / static_cast<double>(min_size)));
    if (lo > hi) return IC_ERR_UNSAT;  // :175-177
    *out = lo < hi ? (lo + hi) / 2 : lo;  // :180-183
    return IC_OK;
}

int ic_load(ic_ctx* ctx, const float* x_host, int64_t n, int64_t d, int64_t ldx) {
    if (!ctx) return IC_ERR_BAD_ARG;
    return load_common(ctx, x_host, n, d, ldx, cudaMemcpyHostToDevice);
}

int ic_load_device(ic_ctx* ctx, const float* x_dev, int64_t n, int64_t d, int64_t ldx) {
    if (!ctx) return IC_ERR_BAD_ARG;
    return load_common(ctx, x_dev, n, d, ldx, cudaMemcpyDeviceToDevice);
}

int ic_load_combined(ic_ctx* ctx, const float* img_host, int64_t n, int64_t d_img, int64_t ld_img,
                     const int32_t* label_offsets, const int32_t* label_ids, int64_t n_labels) {
    if (!ctx) return IC_ERR_BAD_ARG;
    if (n_labels < 0 || d_img < 0 || (n > 0 && !label_offsets)) return fail(ctx, IC_ERR_BAD_ARG, "bad label arguments");
    const int64_t d = d_img + n_labels;
    int rc = validate_sizes(ctx, n, d_img, ld_img);
    if (rc != IC_OK) return rc;
    if (n > 0 && d_img > 0 && !img_host) return fail(ctx, IC_ERR_BAD_ARG, "img is NULL");
    const int64_t nnz = n > 0 ? label_offsets[n] : 0;
    for (int64_t i = 0; i < n; ++i)
        if (label_offsets[i] < 0 || label_offsets[i] > label_offsets[i + 1])
            return fail(ctx, IC_ERR_BAD_ARG, "label_offsets must be non-decreasing");
    if (nnz > 0 && !label_ids) return fail(ctx, IC_ERR_BAD_ARG, "label_ids is NULL");
    for (int64_t q = 0; q < nnz; ++q)  // -1: a label that is not in the set (ignored, embeddings.go:169); anything else must index the set
        if (label_ids[q] < -1 || label_ids[q] >= n_labels)
            return fail(ctx, IC_ERR_BAD_ARG, "label id " + std::to_string(label_ids[q]) + " outside [-1, n_labels)");
    IC_CUDA(cudaSetDevice(ctx->device));
    rc = alloc_problem(ctx, n, d);
    if (rc != IC_OK) return rc;
    std::memset(&ctx->stats, 0, sizeof(ctx->stats));
    ctx->stats.n_items = n;
    ctx->stats.dim = d;
    ctx->stats.near_tie_tol = static_cast<float>(ctx->near_tie_tol);
    ctx->stats.matrix_bytes = static_cast<int64_t>(sizeof(float)) * (row_end(ctx) - row_begin(ctx)) * ctx->ld;
    struct Temp {  // freed on every path out of this function
        int32_t* p = nullptr;
        ~Temp() {
            if (p) cudaFree(p);
        }
    } t_off, t_ids;
    int32_t *&d_off = t_off.p, *&d_ids = t_ids.p;
    IC_CUDA(cudaEventRecord(ctx->ev[0], ctx->stream));
    if (n > 0) {
        if (d_img > 0)  // copy(combined, embedding): the image block of every row (embeddings.go:180)
            IC_CUDA(cudaMemcpy2DAsync(ctx->x, sizeof(float) * d, img_host, sizeof(float) * ld_img, sizeof(float) * d_img, n,
                                      cudaMemcpyHostToDevice, ctx->stream));
        IC_CUDA(cudaMalloc(&d_off, sizeof(int32_t) * (n + 1)));
        IC_CUDA(cudaMalloc(&d_ids, sizeof(int32_t) * (nnz > 0 ? nnz : 1)));
        IC_CUDA(cudaMemcpyAsync(d_off, label_offsets, sizeof(int32_t) * (n + 1), cudaMemcpyHostToDevice, ctx->stream));
        if (nnz > 0)
            IC_CUDA(cudaMemcpyAsync(d_ids, label_ids, sizeof(int32_t) * nnz, cudaMemcpyHostToDevice, ctx->stream));
        IC_CUDA(launch_label_block(ctx->x, n, d, d_img, d_off, d_ids, ctx->stream));
        ctx->stats.kernel_launches += 1;
        ctx->stats.h2d_bytes += static_cast<int64_t>(sizeof(float)) * n * d_img + 4 * (n + 1 + nnz);
    }
    IC_CUDA(cudaEventRecord(ctx->ev[1], ctx->stream));
    IC_CUDA(cudaStreamSynchronize(ctx->stream));  // the host arrays and the two temporaries are done with
    ctx->loaded = true;
    return IC_OK;
}

int ic_read_x(ic_ctx* ctx, float* out_host, int64_t ld) {
    if (!ctx || !out_host) return IC_ERR_BAD_ARG;
    if (!ctx->loaded) return fail(ctx, IC_ERR_STATE, "no problem loaded");
    if (ld < ctx->d) return fail(ctx, IC_ERR_BAD_ARG, "ld < d");
    IC_CUDA(cudaSetDevice(ctx->device));
    if (ctx->n > 0 && ctx->d > 0)
        IC_CUDA(cudaMemcpy2DAsync(out_host, sizeof(float) * ld, ctx->x, sizeof(float) * ctx->d, sizeof(float) * ctx->d,
                                  ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
    IC_CUDA(cudaStreamSynchronize(ctx->stream));
    return IC_OK;
}

int ic_initial_distances(ic_ctx* ctx, int mode, int64_t max_size) {
    if (!ctx) return IC_ERR_BAD_ARG;
    IC_CUDA(cudaSetDevice(ctx->device));
    const int rc = initial_distances(ctx, mode, max_size);
    if (rc != IC_OK) return rc;
    IC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.ms_prep = ev_ms(ctx->ev[2], ctx->ev[3]);
    ctx->stats.ms_gram = ev_ms(ctx->ev[3], ctx->ev[4]);
    return IC_OK;
}

int ic_set_matrix(ic_ctx* ctx, const float* m_host, int64_t ld) {
    if (!ctx || !m_host) return IC_ERR_BAD_ARG;
    if (!ctx->loaded) return fail(ctx, IC_ERR_STATE, "no problem loaded");
    if (ld < ctx->n) return fail(ctx, IC_ERR_BAD_ARG, "ld < n");
    IC_CUDA(cudaSetDevice(ctx->device));
    const int64_t r0 = row_begin(ctx), r1 = row_end(ctx);  // a sharded context keeps its own row block
    ctx->dm_lower_only = false;
    if (r1 > r0)
        IC_CUDA(cudaMemcpy2DAsync(ctx->dm, sizeof(float) * ctx->ld, m_host + r0 * ld, sizeof(float) * ld,
                                  sizeof(float) * ctx->n, r1 - r0, cudaMemcpyHostToDevice, ctx->stream));
    IC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->have_dm = true;
    ctx->have_nn = false;
    reset_epoch(ctx);
    // a supplied matrix need not be the distances of the resident X: Lance-Williams values only, unless "exact" = 2
    ctx->exact_on = ctx->exact_opt == 2 && ctx->cen != nullptr;
    ctx->dm_is_reference = false;
    ctx->gram_mode_used = -1;
    return IC_OK;
}

int ic_nn_init(ic_ctx* ctx) {
    if (!ctx) return IC_ERR_BAD_ARG;
    IC_CUDA(cudaSetDevice(ctx->device));
    IC_CUDA(cudaEventRecord(ctx->ev[4], ctx->stream));
    const int rc = nn_init(ctx);
    if (rc != IC_OK) return rc;
    IC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->stats.ms_nn_init = ev_ms(ctx->ev[4], ctx->ev[5]);
    return IC_OK;
}

int ic_find_closest(ic_ctx* ctx, int32_t* key_hi, int32_t* key_lo, float* dist) {
    if (!ctx || !key_hi || !key_lo || !dist) return IC_ERR_BAD_ARG;
    if (!ctx->have_nn) return fail(ctx, IC_ERR_STATE, "ic_nn_init has not run");
    IC_CUDA(cudaSetDevice(ctx->device));
    ctx->delta_cut_cur = ctx->delta_cut;
    int rc = IC_OK;
    for (int tries = 0; tries < 16; ++tries) {
        rc = run_loop(ctx, 0, 0x3FFFFFFF, 0);  // zero merges: fold the NN cache only
        if (rc != IC_OK) return rc;
        rc = sync_loop_result(ctx);
        if (rc != IC_OK) return rc < 0 ? rc : fail(ctx, IC_ERR_INTERNAL, "batch order check failed");
        // with the horizon, the minimum is the reference's own value only once it lies at or below the safe bound
        if (!(ctx->exact_on && use_batch(ctx) && ctx->cen) || ctx->h_ctl[CTL_NEXT_HI] < 0) break;
        float h = 0.0f;
        const uint32_t hb = static_cast<uint32_t>(ctx->h_ctl[CTL_NEXT_DIST]);
        std::memcpy(&h, &hb, 4);
        const double safe = ctx->horizon < 0.0 ? -1.0 : (ctx->horizon - ctx->abs_slack) / (1.0 + 2.0 * ctx->eps_filter);
        if (static_cast<double>(h) <= safe) break;
        rc = raise_horizon(ctx);
        if (rc != IC_OK) return rc;
    }
    *key_hi = ctx->h_ctl[CTL_NEXT_HI];
    *key_lo = ctx->h_ctl[CTL_NEXT_LO];
    const uint32_t bits = static_cast<uint32_t>(ctx->h_ctl[CTL_NEXT_DIST]);
    std::memcpy(dist, &bits, 4);
    return IC_OK;
}

int ic_merge_loop(ic_ctx* ctx, int64_t min_size, int64_t max_size, int64_t max_merges) {
    if (!ctx) return IC_ERR_BAD_ARG;
    if (!ctx->have_nn) return fail(ctx, IC_ERR_STATE, "ic_nn_init has not run");
    int64_t n_target = 0;
    int rc = ic_optimal_clusters(ctx->n, min_size, max_size, &n_target);
    if (rc != IC_OK) return fail(ctx, rc, "cluster size constraints cannot be satisfied");
    ctx->n_target = n_target;
    IC_CUDA(cudaSetDevice(ctx->device));
    IC_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));
    ctx->delta_cut_cur = ctx->delta_cut;
    rc = run_loop(ctx, n_target, max_size, max_merges);
    if (rc != IC_OK) return rc;
    rc = sync_loop_result(ctx);
    if (rc == kRestart)
        return fail(ctx, IC_ERR_INTERNAL, "batch order check failed: set option delta_cut (e.g. 1e-5) and start again from ic_initial_distances");
    if (rc != IC_OK) return rc;
    IC_CUDA(cudaEventRecord(ctx->ev[6], ctx->stream));
    IC_CUDA(cudaEventSynchronize(ctx->ev[6]));
    ctx->stats.ms_loop = ev_ms(ctx->ev[5], ctx->ev[6]);
    fill_stats(ctx);
    return IC_OK;
}

int ic_build_clusters(ic_ctx* ctx, int64_t min_size, int32_t* cluster_offsets, int32_t* members,
                      int32_t* n_clusters) {
    if (!ctx || !cluster_offsets || !members || !n_clusters) return IC_ERR_BAD_ARG;
    if (!ctx->have_nn) return fail(ctx, IC_ERR_STATE, "no merge state");
    IC_CUDA(cudaSetDevice(ctx->device));
    const int rc = fetch_trace(ctx);
    if (rc != IC_OK) return rc;
    return assemble(ctx, min_size, 0, cluster_offsets, members, n_clusters);
}

int ic_run_resident(ic_ctx* ctx, int64_t min_size, int64_t max_size, int32_t* cluster_offsets, int32_t* members,
                    int32_t* n_clusters, ic_stats* stats) {
    if (!ctx) return IC_ERR_BAD_ARG;
    IC_CUDA(cudaSetDevice(ctx->device));
    const double t0 = now_ms();
    ctx->stats.kernel_launches = 0;
    ctx->stats.d2h_bytes = 0;
    return run_resident(ctx, min_size, max_size, cluster_offsets, members, n_clusters, stats, t0);
}

int ic_cluster_with_constraints(ic_ctx* ctx, const float* x, int64_t n, int64_t d, int64_t ldx, int64_t min_size,
                                int64_t max_size, int32_t* cluster_offsets, int32_t* members, int32_t* n_clusters,
                                ic_stats* stats) {
    if (!ctx) return IC_ERR_BAD_ARG;
    if (n_clusters) *n_clusters = 0;
    const double t0 = now_ms();
    // the reference checks the constraints before touching the data (clustering.go:203-207)
    int64_t n_target = 0;
    int rc = ic_optimal_clusters(n, min_size, max_size, &n_target);
    if (rc != IC_OK) return fail(ctx, rc, "cluster size constraints cannot be satisfied");
    rc = load_common(ctx, x, n, d, ldx, cudaMemcpyHostToDevice);
    if (rc != IC_OK) return rc;
    return run_resident(ctx, min_size, max_size, cluster_offsets, members, n_clusters, stats, t0);
}


// ---- row-block sharding across GPUs: one process per GPU --------------------------------
namespace {
struct ShardHandle {  // what ic_shard_export writes (IC_SHARD_HANDLE_BYTES)
    uint32_t magic;
    int32_t rank, world;
    int32_t pad;
    int64_t n, ld, rows;
    cudaIpcMemHandle_t dm, box, dm_b;
    int32_t has_b, pad2;
};
static_assert(sizeof(ShardHandle) <= IC_SHARD_HANDLE_BYTES, "handle blob too small");
constexpr uint32_t kShardMagic = 0x49435348u;
}  // namespace

int ic_shard_init(ic_ctx* ctx, int rank, int world) {
    if (!ctx) return IC_ERR_BAD_ARG;
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail(ctx, IC_ERR_BAD_ARG, "bad rank / world");
    if (ctx->vranks > 1 && world > 1) return fail(ctx, IC_ERR_STATE, "virtual_ranks is set");
    ctx->shard_rank = rank;
    ctx->shard_world = world;
    return IC_OK;
}

int ic_shard_export(ic_ctx* ctx, void* handle) {
    if (!ctx || !handle) return IC_ERR_BAD_ARG;
    if (ctx->shard_world <= 1) return fail(ctx, IC_ERR_STATE, "not a sharded context");
    if (!ctx->loaded) return fail(ctx, IC_ERR_STATE, "no problem loaded");
    IC_CUDA(cudaSetDevice(ctx->device));
    ShardHandle h{};
    h.magic = kShardMagic;
    h.rank = ctx->shard_rank;
    h.world = ctx->shard_world;
    h.n = ctx->n;
    h.ld = ctx->ld;
    h.rows = rows_per_rank(ctx);
    IC_CUDA(cudaIpcGetMemHandle(&h.dm, ctx->dm));
    IC_CUDA(cudaIpcGetMemHandle(&h.box, ctx->rankbox));
    h.has_b = ctx->dm_b ? 1 : 0;
    if (ctx->dm_b) IC_CUDA(cudaIpcGetMemHandle(&h.dm_b, ctx->dm_b));
    std::memset(handle, 0, IC_SHARD_HANDLE_BYTES);
    std::memcpy(handle, &h, sizeof(h));
    return IC_OK;
}

int ic_shard_connect(ic_ctx* ctx, const void* handles) {
    if (!ctx || !handles) return IC_ERR_BAD_ARG;
    if (ctx->shard_world <= 1) return fail(ctx, IC_ERR_STATE, "not a sharded context");
    if (!ctx->loaded) return fail(ctx, IC_ERR_STATE, "no problem loaded");
    IC_CUDA(cudaSetDevice(ctx->device));
    if (ctx->peers_open) return IC_OK;  // same allocations as before (alloc_problem keeps them across loads)
    for (int q = 0; q < ctx->shard_world; ++q) {
        ShardHandle h{};
        std::memcpy(&h, static_cast<const uint8_t*>(handles) + static_cast<size_t>(q) * IC_SHARD_HANDLE_BYTES, sizeof(h));
        if (h.magic != kShardMagic || h.rank != q || h.world != ctx->shard_world || h.n != ctx->n || h.ld != ctx->ld ||
            h.rows != rows_per_rank(ctx))
            return fail(ctx, IC_ERR_BAD_ARG, "shard handle of rank " + std::to_string(q) + " does not match this problem");
        if (q == ctx->shard_rank) {
            ctx->peer_dm[q] = ctx->dm;
            ctx->peer_box[q] = ctx->rankbox;
            ctx->peer_dm_b[q] = ctx->dm_b;
        } else {
            IC_CUDA(cudaIpcOpenMemHandle(&ctx->peer_dm[q], h.dm, cudaIpcMemLazyEnablePeerAccess));
            IC_CUDA(cudaIpcOpenMemHandle(&ctx->peer_box[q], h.box, cudaIpcMemLazyEnablePeerAccess));
            if ((h.has_b != 0) != (ctx->dm_b != nullptr))
                return fail(ctx, IC_ERR_BAD_ARG, "rank " + std::to_string(q) + " and this rank disagree about the compaction buffer");
            if (h.has_b) IC_CUDA(cudaIpcOpenMemHandle(&ctx->peer_dm_b[q], h.dm_b, cudaIpcMemLazyEnablePeerAccess));
        }
    }
    ctx->peers_open = true;
    return IC_OK;
}

int ic_shard_rows(ic_ctx* ctx, int64_t* row_begin_out, int64_t* row_end_out) {
    if (!ctx || !row_begin_out || !row_end_out) return IC_ERR_BAD_ARG;
    *row_begin_out = row_begin(ctx);
    *row_end_out = row_end(ctx);
    return IC_OK;
}

int ic_read_matrix(ic_ctx* ctx, float* out_host, int64_t ld) {
    if (!ctx || !out_host) return IC_ERR_BAD_ARG;
    if (!ctx->have_dm) return fail(ctx, IC_ERR_STATE, "no distance matrix");
    if (ld < ctx->n) return fail(ctx, IC_ERR_BAD_ARG, "ld < n");
    if (ctx->n_compactions > 0) return fail(ctx, IC_ERR_STATE, "the matrix has been compacted (option \"compact\" = 0 keeps the slot layout)");
    IC_CUDA(cudaSetDevice(ctx->device));
    const int64_t r0 = row_begin(ctx), r1 = row_end(ctx);  // a sharded context fills its own rows only
    if (r1 > r0)
        IC_CUDA(cudaMemcpy2DAsync(out_host + r0 * ld, sizeof(float) * ld, ctx->dm, sizeof(float) * ctx->ld,
                                  sizeof(float) * ctx->n, r1 - r0, cudaMemcpyDeviceToHost, ctx->stream));
    IC_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->dm_lower_only && ctx->shard_world <= 1)  // K1 skipped the mirrored triangle: the caller still gets the symmetric matrix
        for (int64_t i = 0; i < ctx->n; ++i)
            for (int64_t j = i + 1; j < ctx->n; ++j) out_host[i * ld + j] = out_host[j * ld + i];
    return IC_OK;
}

int ic_read_slots(ic_ctx* ctx, int32_t* key, int32_t* size) {
    if (!ctx || !key || !size) return IC_ERR_BAD_ARG;
    if (!ctx->have_nn) return fail(ctx, IC_ERR_STATE, "no merge state");
    IC_CUDA(cudaSetDevice(ctx->device));
    std::vector<int2> h(static_cast<size_t>(ctx->n), make_int2(-1, 0));  // (after a compaction: n_cur dense slots, the rest retired)
    if (ctx->n_cur > 0)
        IC_CUDA(cudaMemcpyAsync(h.data(), ctx->ks, sizeof(int2) * ctx->n_cur, cudaMemcpyDeviceToHost, ctx->stream));
    IC_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int64_t i = 0; i < ctx->n; ++i) {
        key[i] = h[i].x;
        size[i] = h[i].y;
    }
    return IC_OK;
}

int ic_get_merge_trace(ic_ctx* ctx, int32_t* key_hi, int32_t* key_lo, float* dist, int32_t* size, float* gap,
                       int64_t capacity, int64_t* n_merges) {
    if (!ctx || !n_merges) return IC_ERR_BAD_ARG;
    if (!ctx->have_nn) return fail(ctx, IC_ERR_STATE, "no merge state");
    IC_CUDA(cudaSetDevice(ctx->device));
    const int rc = fetch_trace(ctx);
    if (rc != IC_OK) return rc;
    const int64_t m = ctx->n_merges;
    *n_merges = m;
    if (capacity < m) return fail(ctx, IC_ERR_BAD_ARG, "trace capacity too small");
    if (key_hi) std::memcpy(key_hi, ctx->h_key_hi.data(), 4 * m);
    if (key_lo) std::memcpy(key_lo, ctx->h_key_lo.data(), 4 * m);
    if (dist) std::memcpy(dist, ctx->h_dist.data(), 4 * m);
    if (size) std::memcpy(size, ctx->h_size.data(), 4 * m);
    if (gap) std::memcpy(gap, ctx->h_gap.data(), 4 * m);
    return IC_OK;
}

int ic_get_linkage(ic_ctx* ctx, double* z, int64_t capacity, int64_t* n_rows) {
    if (!ctx || !n_rows) return IC_ERR_BAD_ARG;
    if (!ctx->have_nn) return fail(ctx, IC_ERR_STATE, "no merge state");
    IC_CUDA(cudaSetDevice(ctx->device));
    const int rc = fetch_trace(ctx);
    if (rc != IC_OK) return rc;
    const int64_t m = ctx->n_merges;
    *n_rows = m;
    if (capacity < m || (m > 0 && !z)) return fail(ctx, IC_ERR_BAD_ARG, "linkage capacity too small");
    for (int64_t t = 0; t < m; ++t) {
        z[4 * t + 0] = static_cast<double>(ctx->h_key_lo[t]);
        z[4 * t + 1] = static_cast<double>(ctx->h_key_hi[t]);
        z[4 * t + 2] = std::sqrt(2.0 * static_cast<double>(ctx->h_dist[t]));  // Ward distance d = h^2 / 2
        z[4 * t + 3] = static_cast<double>(ctx->h_size[t]);
    }
    return IC_OK;
}

int ic_get_loop_block_waits(ic_ctx* ctx, int64_t* out, int64_t capacity, int64_t* n_blocks) {
    if (!ctx || !out || !n_blocks) return IC_ERR_BAD_ARG;
    *n_blocks = ctx->loop_grid;
    for (int64_t i = 0; i < capacity && i < ctx->loop_grid && i < 240; ++i) out[i] = ctx->h_prof[16 + i];
    return IC_OK;
}

int ic_get_loop_profile(ic_ctx* ctx, int64_t* out16) {
    if (!ctx || !out16) return IC_ERR_BAD_ARG;
    for (int i = 0; i < 16; ++i) out16[i] = ctx->h_prof[i];
    if (ctx->loop_mode_used != 1) {  // (the batched loop keeps cycle counters of its selection phase in these two)
        out16[9] = ctx->h_ctl[CTL_BUBBLES];
        out16[7] = ctx->h_ctl[CTL_RESCANS];
    }
    return IC_OK;
}

int ic_get_stats(ic_ctx* ctx, ic_stats* stats) {
    if (!ctx || !stats) return IC_ERR_BAD_ARG;
    fill_stats(ctx);
    *stats = ctx->stats;
    return IC_OK;
}

int ic_time_kernel(ic_ctx* ctx, const char* which, int repeats, float* ms_each) {
    if (!ctx || !which || !ms_each || repeats < 1) return IC_ERR_BAD_ARG;
    if (!ctx->loaded) return fail(ctx, IC_ERR_STATE, "no problem loaded");
    IC_CUDA(cudaSetDevice(ctx->device));
    const std::string k(which);
    if (k == "gram" || k == "split") {
        const int rc = do_prep(ctx);
        if (rc != IC_OK) return rc;
    }
    if (k == "gram_i8") {
        const int rc = do_prep_i8(ctx);
        if (rc != IC_OK) return rc;
    }
    if (k == "nn_sweep" && !ctx->have_dm) return fail(ctx, IC_ERR_STATE, "no distance matrix");
    IC_CUDA(cudaStreamSynchronize(ctx->stream));
    IC_CUDA(cudaEventRecord(ctx->ev[8], ctx->stream));
    for (int r = 0; r < repeats; ++r) {
        if (k == "gram") {
            const int rc = do_gram(ctx, IC_GRAM_TCGEN05_3XTF32);
            if (rc != IC_OK) return rc;
        } else if (k == "gram_i8") {
            const int rc = do_gram(ctx, IC_GRAM_TCGEN05_I8);
            if (rc != IC_OK) return rc;
        } else if (k == "gram_exact") {
            const int rc = do_gram(ctx, IC_GRAM_EXACT_FP32);
            if (rc != IC_OK) return rc;
        } else if (k == "split") {
            IC_CUDA(launch_split(ctx->x, ctx->n, ctx->d, ctx->d, ctx->colsum, ctx->center, ctx->hi, ctx->lo, ctx->norms,
                                 ctx->n_pad, ctx->d_pad, ctx->stream));
        } else if (k == "nn_sweep") {
            IC_CUDA(launch_nn_sweep(ctx->dm, row_begin(ctx), row_end(ctx), ctx->ld, ctx->nn, ctx->nn_more, ctx->stream));
        } else {
            return fail(ctx, IC_ERR_BAD_ARG, "unknown kernel " + k);
        }
    }
    IC_CUDA(cudaEventRecord(ctx->ev[9], ctx->stream));
    IC_CUDA(cudaEventSynchronize(ctx->ev[9]));
    *ms_each = ev_ms(ctx->ev[8], ctx->ev[9]) / static_cast<float>(repeats);
    if (k == "gram" || k == "gram_exact" || k == "gram_i8") ctx->have_dm = true;
    if (k != "split") ctx->have_nn = false;  // the loop state no longer matches the matrix
    return IC_OK;
}

}  // extern "C"
