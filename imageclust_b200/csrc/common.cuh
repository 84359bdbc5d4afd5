// common.cuh -- shared device helpers (sm_100a inline PTX) for imageclust_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

#ifndef IC_DEVINL
#define IC_DEVINL __device__ __forceinline__
#endif

namespace ic {

constexpr uint64_t kPackInf = 0xFFFFFFFFFFFFFFFFull;  // "no candidate" for packed (dist,key) minima
constexpr uint32_t kMaxFloatBits = 0x7F7FFFFFu;       // bits of MaxFloat32 (clustering.go:120)
constexpr uint32_t kInfBits = 0x7F800000u;

// ---- packed candidates -------------------------------------------------------------------
// Non-negative floats order like their bit patterns, so (dist_bits << 32 | key) ordered as
// u64 is the lexicographic order (dist, key) that FindClosestClusters' strict '<' scan
// implies (clustering.go:123-131).  +inf and NaN are above MaxFloat32 and never win.
IC_DEVINL uint64_t pack_cand(float d, uint32_t key) {
    return (static_cast<uint64_t>(__float_as_uint(d)) << 32) | key;
}
IC_DEVINL float pack_dist(uint64_t p) { return __uint_as_float(static_cast<uint32_t>(p >> 32)); }
IC_DEVINL uint32_t pack_key(uint64_t p) { return static_cast<uint32_t>(p); }
IC_DEVINL bool pack_selectable(uint64_t p) { return static_cast<uint32_t>(p >> 32) < kMaxFloatBits; }

IC_DEVINL uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// warp-wide minimum of a u64 with two redux.sync.min.u32: high words first, then the low words of the
// lanes that hold the minimal high word (a shuffle butterfly costs 10 shuffles and 5 dependent steps)
IC_DEVINL uint64_t warp_min_u64(uint64_t v) {
    const uint32_t hi = static_cast<uint32_t>(v >> 32);
    const uint32_t mh = __reduce_min_sync(0xffffffffu, hi);
    const uint32_t lo = hi == mh ? static_cast<uint32_t>(v) : 0xFFFFFFFFu;
    const uint32_t ml = __reduce_min_sync(0xffffffffu, lo);
    return (static_cast<uint64_t>(mh) << 32) | ml;
}

// canonical non-negative distance: NaN -> +inf (never selectable, like the reference's
// strict '<'), negative / -0 -> +0
IC_DEVINL float canon_dist(float v) {
    if (!(v >= 0.0f)) v = (v != v) ? __uint_as_float(kInfBits) : 0.0f;
    return v + 0.0f;
}

// two smallest of a set of packed candidates (m1 <= m2); kPackInf when absent
struct Top2 {
    uint64_t m1, m2;
};
IC_DEVINL void top2_insert(Top2& t, uint64_t v) {
    if (v < t.m1) {
        t.m2 = t.m1;
        t.m1 = v;
    } else if (v < t.m2) {
        t.m2 = v;
    }
}
IC_DEVINL void top2_merge(Top2& t, uint64_t o1, uint64_t o2) {
    const uint64_t lo = umin64(t.m1, o1);
    const uint64_t hi = t.m1 < o1 ? o1 : t.m1;
    t.m2 = umin64(hi, umin64(t.m2, o2));
    t.m1 = lo;
}
// warp-wide two smallest of every lane's two smallest (m1 values are unique unless kPackInf): the runner-up is
// the winner lane's second or another lane's first
IC_DEVINL Top2 warp_top2(Top2 t) {
    Top2 r;
    r.m1 = warp_min_u64(t.m1);
    r.m2 = warp_min_u64(t.m1 == r.m1 ? t.m2 : t.m1);
    return r;
}

// ---- per-row partner lists ----------------------------------------------------------------
// d(r,u) between two unchanged clusters never changes and a new cluster is never a partner of an
// existing row (it carries the highest key), so a row's sorted partner list is static: partners
// only die.  Each row therefore caches its kNNK smallest partners; a whole-row rescan is needed
// only when all of them have died.
struct Cand2 {  // two smallest packed candidates seen by one thread, and how many it saw
    uint64_t c1, c2;
    int32_t s1, s2;
    int32_t cnt;
};
IC_DEVINL void cand2_init(Cand2& c) {
    c.c1 = c.c2 = kPackInf;
    c.s1 = c.s2 = -1;
    c.cnt = 0;
}
IC_DEVINL void cand2_insert(Cand2& c, uint64_t p, int32_t slot) {
    ++c.cnt;
    if (p < c.c2) {
        if (p < c.c1) {
            c.c2 = c.c1;
            c.s2 = c.s1;
            c.c1 = p;
            c.s1 = slot;
        } else {
            c.c2 = p;
            c.s2 = slot;
        }
    }
}

struct TopKScratch {
    uint64_t red[32];
    int32_t cnt[32];
    uint64_t pack[kNNK];
    int32_t slot[kNNK];
    int32_t stop;
};

// Block-wide selection of the (up to) kNNK smallest candidates from every thread's two smallest.
// EXACT: the list is cut right after an entry that was a thread's second smallest while that
// thread saw more than two candidates (its unseen third could be smaller than what follows).
// Returns the number of entries m (sc.pack / sc.slot[0..m) valid for all threads after return)
// and sets `more` when eligible candidates exist beyond the list.
template <int kT>
IC_DEVINL int block_select_topk(const Cand2& c, TopKScratch& sc, bool& more) {
    constexpr int kW = kT / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int taken = 0, m = 0, total = 0;
    for (int r = 0; r < kNNK; ++r) {
        const uint64_t cand = taken == 0 ? c.c1 : (taken == 1 ? c.c2 : kPackInf);
        const uint64_t wm = warp_min_u64(cand);
        int wc = c.cnt;
        if (r == 0) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wc += __shfl_xor_sync(0xffffffffu, wc, o);
        }
        if (lane == 0) {
            sc.red[warp] = wm;
            if (r == 0) sc.cnt[warp] = wc;
        }
        __syncthreads();
        uint64_t bm = sc.red[0];
#pragma unroll
        for (int w = 1; w < kW; ++w) bm = umin64(bm, sc.red[w]);
        if (r == 0) {
#pragma unroll
            for (int w = 0; w < kW; ++w) total += sc.cnt[w];
        }
        if (bm == kPackInf) break;  // uniform
        if (cand == bm) {           // packs are unique: exactly one thread
            sc.pack[r] = bm;
            sc.slot[r] = taken == 0 ? c.s1 : c.s2;
            ++taken;
            sc.stop = (taken == 2 && c.cnt > 2) ? 1 : 0;
        }
        __syncthreads();
        m = r + 1;
        if (sc.stop) break;  // uniform
    }
    more = total > m;
    __syncthreads();
    return m;
}

// ---- memory ------------------------------------------------------------------------------
IC_DEVINL float4 ld_stream_f4(const float4* p) {  // read-once data: bypass L1
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
IC_DEVINL uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
IC_DEVINL void red_release_add_u32(uint32_t* p, uint32_t v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

IC_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier ----------------------------------------------------------------------------
IC_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
IC_DEVINL void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
IC_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
IC_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
IC_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap, not hang the GPU box.
IC_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
        if (spin > (1u << 22)) __trap();
    }
}

// ---- TMA ---------------------------------------------------------------------------------
IC_DEVINL void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
IC_DEVINL void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t c0, int32_t c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

// ---- tcgen05 / TMEM ----------------------------------------------------------------------
IC_DEVINL void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
                 "r"(ncols)
                 : "memory");
}
IC_DEVINL void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
IC_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
IC_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
IC_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], kind::tf32, issued by ONE thread
IC_DEVINL void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::i8 (signed 8-bit operands, exact 32-bit integer accumulation), ONE thread
IC_DEVINL void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread completed
IC_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 consecutive columns -> 32 registers per thread (thread = lane/row)
IC_DEVINL void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
IC_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled shared-memory operand descriptor (UMMA SmemDescriptor):
//   [0,14) start address >> 4, [16,30) leading byte offset >> 4 (unused for swizzled K-major),
//   [32,46) stride byte offset >> 4 = 1024 B between 8-row groups, [46,48) version = 1,
//   [61,64) layout = 2 (SWIZZLE_128B).
IC_DEVINL uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// round-to-nearest TF32 (10-bit mantissa) kept in an fp32 container
IC_DEVINL float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

}  // namespace ic
