// kernels.h -- host-callable launchers of the sm_100a kernels (internal to the library).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ic {

// ---- resident per-slot state -------------------------------------------------------------
// A "slot" is a row/column of the distance matrix.  Slot s holds one live cluster
// (the reference's clusters[] entry, clustering.go:11-15) or is retired.
//   ks[s]  = {key, size}: key is the monotone order id (item index for singletons,
//            N + t for the cluster made by merge t) == the reference's slice order;
//            key < 0 when the slot is retired.
//   nn[s*kNNK + j] = j-th smallest cached partner among clusters with a LOWER key:
//            {partner key, distance bits, partner slot, partner size};
//            y == 0xFFFFFFFF for an empty entry.  nn_more[s] != 0: selectable partners
//            may exist beyond the list (rescan the row when the list runs dry).
typedef int2 SlotKS;
typedef uint4 SlotNN;
constexpr int kNNK = 4;  // cached partners per row
constexpr uint32_t kNoPartner = 0xFFFFFFFFu;

// ---- K0 prep (prep.cu) -------------------------------------------------------------------
// column sums of X [n x d] (row stride ldx) in double
cudaError_t launch_colsum(const float* x, int64_t n, int64_t d, int64_t ldx, double* colsum, cudaStream_t s);
// xc = x - mean; hi = tf32(xc); lo = tf32(xc - hi); norm[i] = sum_k (hi+lo)^2 (double).
// hi/lo are [n_pad x d_pad] (zero padded), norms [n_pad].
cudaError_t launch_split(const float* x, int64_t n, int64_t d, int64_t ldx, const double* colsum, int center,
                         float* hi, float* lo, double* norms, int64_t n_pad, int64_t d_pad, cudaStream_t s);
cudaError_t launch_fill(float* dm, int64_t count, float value, cudaStream_t s);
// ks[s] = {s, 1}, gkey[s] = s (padding up to a multiple of 4: -1)
cudaError_t launch_init_slots(SlotKS* ks, int32_t* gkey, int64_t n, cudaStream_t s);

// ---- K1 Gram / initial Ward distances ----------------------------------------------------
constexpr int kGramBM = 128;  // tile rows   (UMMA M)
constexpr int kGramBN = 256;  // tile cols   (UMMA N)
constexpr int kGramBK = 32;   // fp32 elements per 128-byte swizzle row
struct GramPlan {
    CUtensorMap map_hi;  // box {32 floats, 128 rows} over hi [n_pad x d_pad]
    CUtensorMap map_lo;
    const int2* tiles;   // (row block, col block) list in L2-friendly order
    int n_tiles;
    int k_blocks;        // d_pad / 32
};
// dm[i][j] = dm[j][i] = max(0.5*(norm_i + norm_j) - <x_i, x_j>, 0), diag 0; full square, row stride ld
cudaError_t launch_gram_tcgen05(const GramPlan& plan, const double* norms, float* dm, int64_t n, int64_t ld,
                                int num_sms, cudaStream_t s, int terms = 23);
size_t gram_tcgen05_smem_bytes();
// audit kernel: the reference's own arithmetic (sequential fp32, clustering.go:136-157), bit exact
cudaError_t launch_gram_exact(const float* x, int64_t n, int64_t d, int64_t ldx, float* dm, int64_t ld,
                              cudaStream_t s);

// ---- K2 nearest-neighbour sweep ----------------------------------------------------------
// First sweep (keys are the slot indices): nn[s][*] = the kNNK smallest (dm[s][u], u) over columns u < s.
cudaError_t launch_nn_sweep(const float* dm, int64_t n, int64_t ld, SlotNN* nn, int32_t* nn_more, cudaStream_t s);

// ---- K3 persistent merge loop ------------------------------------------------------------
struct LoopState {
    float* dm;
    int64_t ld;
    int32_t n;       // slots (== items)
    SlotKS* ks;      // [n]
    int32_t* gkey;   // [round_up(n, 4)] keys only (what a whole-row rescan streams); padding = -1
    SlotNN* nn;      // [n][kNNK]
    int32_t* nn_more; // [n]
    // merge trace, capacity n
    int32_t* tr_key_hi;
    int32_t* tr_key_lo;
    float* tr_dist;
    int32_t* tr_size;
    float* tr_gap;
    // scratch (zeroed by the host before every launch)
    void* records;   // [2][grid] 128-byte exchange records
    int32_t* ctl;    // [16], see CTL_*
    long long* prof; // [16] or NULL: SM cycles block 0 spent per phase
};
struct LoopParams {
    int32_t n_target;    // CalculateOptimalClusters result (clustering.go:220)
    int32_t max_size;    // clustering.go:228
    int32_t max_merges;  // stop after this many merges in this launch (<0: unlimited)
    float near_tie_tol;
};
// ctl[] indices.  N_LIVE and N_MERGES are read at launch (resume) and written at exit.
enum { CTL_N_LIVE = 0, CTL_N_MERGES = 1, CTL_EXHAUSTED = 2, CTL_ERROR = 3, CTL_NEAR_TIES = 4, CTL_RESCANS = 5,
       CTL_DONE = 6, CTL_BIG_RESCANS = 7, CTL_NEXT_HI = 8, CTL_NEXT_LO = 9, CTL_NEXT_DIST = 10 };
int merge_loop_threads(int64_t n, int num_sms);
cudaError_t merge_loop_max_grid(int threads, int num_sms, int64_t n, int* grid);
size_t merge_loop_smem_bytes(int64_t n, int grid);
size_t merge_loop_record_bytes();
cudaError_t launch_merge_loop(const LoopState& st, const LoopParams& p, int grid, int threads, cudaStream_t s);

}  // namespace ic
