// kernels.h -- host-callable launchers of the sm_100a kernels (internal to the library).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ic {

// ---- K0 prep (prep.cu) -------------------------------------------------------------------
// column sums of X [n x d] (row stride ldx) in double
cudaError_t launch_colsum(const float* x, int64_t n, int64_t d, int64_t ldx, double* colsum, cudaStream_t s);
// xc = x - mean; hi = tf32(xc); lo = tf32(xc - hi); norm[i] = sum_k (hi+lo)^2 (double).
// hi/lo are [n_pad x d_pad] (zero padded), norms [n_pad].
cudaError_t launch_split(const float* x, int64_t n, int64_t d, int64_t ldx, const double* colsum, int center,
                         float* hi, float* lo, double* norms, int64_t n_pad, int64_t d_pad, cudaStream_t s);

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
                                int num_sms, cudaStream_t s);
size_t gram_tcgen05_smem_bytes();
// audit kernel: the reference's own arithmetic (sequential fp32, clustering.go:136-157), bit exact
cudaError_t launch_gram_exact(const float* x, int64_t n, int64_t d, int64_t ldx, float* dm, int64_t ld,
                              cudaStream_t s);
cudaError_t launch_fill(float* dm, int64_t count, float value, cudaStream_t s);

// ---- K2 nearest-neighbour sweep ----------------------------------------------------------
// nn_pack[s] = min over partners u with key[u] < key[s] (alive) of (dm[s][u], key[u]);
// nn_slot[s] = slot of that partner (-1 if none).  identity_keys: key[s] == s (first sweep).
cudaError_t launch_nn_sweep(const float* dm, int64_t n, int64_t ld, const int32_t* key, const int32_t* slot_of_key,
                            int identity_keys, unsigned long long* nn_pack, int32_t* nn_slot, cudaStream_t s);

// ---- K3 persistent merge loop ------------------------------------------------------------
struct LoopState {
    float* dm;
    int64_t ld;
    int32_t n;                    // slots
    int32_t* key;                 // [n] -1 = retired
    int32_t* size;                // [n]
    int32_t* slot_of_key;         // [2n]
    unsigned long long* nn_pack;  // [n]
    int32_t* nn_slot;             // [n]
    // trace, capacity n
    int32_t* tr_key_hi;
    int32_t* tr_key_lo;
    float* tr_dist;
    int32_t* tr_size;
    float* tr_gap;
    // scratch
    unsigned long long* partials;  // [2][grid][4]
    int32_t* rlist;                // [3][n]
    int32_t* rcount;               // [3]
    uint32_t* barrier;             // [1]
    int32_t* ctl;                  // [16]: see merge_loop.cu
};
struct LoopParams {
    int32_t n_target;
    int32_t max_size;
    int32_t max_merges;  // <0: unlimited
    float near_tie_tol;
};
int merge_loop_max_grid(int threads);
cudaError_t launch_merge_loop(const LoopState& st, const LoopParams& p, int grid, int threads, cudaStream_t s);

// ctl[] indices
enum { CTL_N_LIVE = 0, CTL_N_MERGES = 1, CTL_EXHAUSTED = 2, CTL_ERROR = 3, CTL_NEAR_TIES = 4, CTL_RESCANS = 5,
       CTL_DONE = 6 };

}  // namespace ic
