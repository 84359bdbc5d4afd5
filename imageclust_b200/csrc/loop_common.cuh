// loop_common.cuh -- device helpers shared by the two merge-loop kernels (merge_loop.cu, merge_batch.cu):
// Lance-Williams update, partner-list selection / merging with the exactness cut rule, mailbox loads / stores.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace ic {

IC_DEVINL uint4 ld_volatile_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// weak write-through store: the five chunks of a record are independent (each carries its own tag), and a thread's
// volatile stores are performed one after the other (measured: 3 000 cycles until the fifth one was visible)
IC_DEVINL void st_cg_u4(uint4* p, uint4 v) {
    asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
IC_DEVINL void st_volatile_u4(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// fences: gpu scope on one device; system scope as soon as peers' memory is involved
template <bool kSys>
IC_DEVINL void fence_acq_rel() {
    if (kSys)
        asm volatile("fence.acq_rel.sys;" ::: "memory");
    else
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
}

// bulk asynchronous stores (TMA, async proxy): shared -> global (possibly a peer's memory over NVLink).  Unlike
// generic stores they are not waited for by a later fence of the issuing SM; completion is tracked per thread.
IC_DEVINL void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the generic-proxy writes of the source first
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)),
                 "r"(smem_u32(ssrc)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
IC_DEVINL void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
IC_DEVINL void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// Lance-Williams update of Ward's distance (clustering.go:141-144 is the closed form it
// equals): exact integer weights, double arithmetic, the quotient as a product with the
// correctly rounded reciprocal of the (small integer) size sum, one rounding to fp32.  The CPU
// oracle's LW mode (oracle/ward_fast.c) performs the same operations in the same order.
IC_DEVINL float lance_williams_rcp(int sa, int sb, int sk, float dka, float dkb, float dab, double rcp_den) {
    const double t1 = static_cast<double>(sa + sk) * static_cast<double>(dka);
    const double t2 = static_cast<double>(sb + sk) * static_cast<double>(dkb);
    const double t3 = static_cast<double>(sk) * static_cast<double>(dab);
    const double num = (t1 + t2) - t3;
    return canon_dist(static_cast<float>(num * rcp_den));
}
IC_DEVINL float lance_williams(int sa, int sb, int sk, float dka, float dkb, float dab) {
    return lance_williams_rcp(sa, sb, sk, dka, dkb, dab, 1.0 / static_cast<double>(sa + sb + sk));
}

IC_DEVINL uint4 nn_none() { return make_uint4(kNoPartner, kNoPartner, kNoPartner, 0u); }
// a dry row's placeholder: no partner, only the lower bound of whatever a rescan will find
IC_DEVINL uint4 nn_bound(uint32_t dist_bits) { return make_uint4(kNoPartner, dist_bits, kNoPartner, 0u); }

struct PartList {  // a sorted partner list under construction
    uint64_t pk[kNNK];
    int32_t sl[kNNK], sz[kNNK];
    int32_t m, more;
};

template <typename T>
IC_DEVINL T sel4(const T (&v)[kNNK], int i) {
    return i == 0 ? v[0] : (i == 1 ? v[1] : (i == 2 ? v[2] : v[3]));
}

// Per-lane state of a row scan: the two smallest candidates and whether the lane may have passed over others.
// Elements are pre-filtered on their distance bits alone (one compare); only the rare survivor pays for the key /
// liveness tests, so `extra` is conservative: a passed-over element counts as a possible candidate.  That only
// ever shortens the exact list (see the cut rule below) -- it never admits a wrong entry.
struct ScanCand {
    uint64_t c1, c2;
    int32_t s1, s2;
    bool extra;
};
IC_DEVINL void scan_init(ScanCand& c) {
    c.c1 = c.c2 = kPackInf;
    c.s1 = c.s2 = -1;
    c.extra = false;
}
IC_DEVINL void scan_insert(ScanCand& c, uint64_t p, int32_t slot) {
    if (p < c.c2) {
        if (c.c2 != kPackInf) c.extra = true;  // the old second is passed over
        if (p < c.c1) {
            c.c2 = c.c1;
            c.s2 = c.s1;
            c.c1 = p;
            c.s1 = slot;
        } else {
            c.c2 = p;
            c.s2 = slot;
        }
    } else {
        c.extra = true;
    }
}
// Warp-wide selection of the (up to) kNNK smallest candidates from every lane's two smallest.  EXACT: the list
// is cut right after an entry that was a lane's second smallest while that lane may have passed over others.
IC_DEVINL int warp_select_scan(const ScanCand& c, uint64_t (&pk)[kNNK], int32_t (&sl)[kNNK], bool& more) {
    int taken = 0, m = 0;
#pragma unroll
    for (int r = 0; r < kNNK; ++r) {
        pk[r] = kPackInf;
        sl[r] = -1;
    }
#pragma unroll
    for (int r = 0; r < kNNK; ++r) {
        const uint64_t cand = taken == 0 ? c.c1 : (taken == 1 ? c.c2 : kPackInf);
        const uint64_t wm = warp_min_u64(cand);
        if (wm == kPackInf) break;  // warp uniform
        const bool win = cand == wm;  // packs are unique: exactly one lane
        const int src = __ffs(__ballot_sync(0xffffffffu, win)) - 1;
        const int32_t myslot = taken == 0 ? c.s1 : c.s2;
        pk[r] = wm;
        sl[r] = __shfl_sync(0xffffffffu, myslot, src);
        if (win) ++taken;
        m = r + 1;
        if (__ballot_sync(0xffffffffu, win && taken == 2 && c.extra)) break;
    }
    const int held = (c.c1 != kPackInf ? 1 : 0) + (c.c2 != kPackInf ? 1 : 0);
    more = __any_sync(0xffffffffu, taken < held || c.extra);
    return m;
}

// Warp-wide merge of one sorted list per lane (m entries, `more`: unlisted entries >= the last one
// exist).  Same exactness rule: cut right after a list's last entry if that list has more.
IC_DEVINL void warp_merge_lists(const PartList& in, PartList& out) {
    int ptr = 0;
    out.m = 0;
#pragma unroll
    for (int r = 0; r < kNNK; ++r) {
        out.pk[r] = kPackInf;
        out.sl[r] = -1;
        out.sz[r] = 0;
    }
#pragma unroll
    for (int r = 0; r < kNNK; ++r) {
        const uint64_t head = ptr < in.m ? sel4(in.pk, ptr) : kPackInf;
        const uint64_t wm = warp_min_u64(head);
        if (wm == kPackInf) break;
        const bool win = head == wm;
        const int src = __ffs(__ballot_sync(0xffffffffu, win)) - 1;
        out.pk[r] = wm;
        out.sl[r] = __shfl_sync(0xffffffffu, sel4(in.sl, ptr), src);
        out.sz[r] = __shfl_sync(0xffffffffu, sel4(in.sz, ptr), src);
        if (win) ++ptr;
        out.m = r + 1;
        if (__ballot_sync(0xffffffffu, win && ptr == in.m && in.more != 0)) break;
    }
    out.more = __any_sync(0xffffffffu, ptr < in.m || in.more != 0) ? 1 : 0;
}

// Lane-local merge of two sorted lists (same exactness rule as warp_merge_lists): acc <- merge(acc, b).
IC_DEVINL void lane_merge2(PartList& acc, const PartList& b) {
    PartList o;
    int pa = 0, pb = 0;
    o.m = 0;
#pragma unroll
    for (int r = 0; r < kNNK; ++r) {
        o.pk[r] = kPackInf;
        o.sl[r] = -1;
        o.sz[r] = 0;
    }
    bool cut = false;
#pragma unroll
    for (int r = 0; r < kNNK; ++r) {
        const uint64_t ha = pa < acc.m ? sel4(acc.pk, pa) : kPackInf;
        const uint64_t hb = pb < b.m ? sel4(b.pk, pb) : kPackInf;
        if (cut || (ha == kPackInf && hb == kPackInf)) continue;
        if (ha < hb) {
            o.pk[r] = ha;
            o.sl[r] = sel4(acc.sl, pa);
            ++pa;
            cut = pa == acc.m && acc.more != 0;
        } else {
            o.pk[r] = hb;
            o.sl[r] = sel4(b.sl, pb);
            ++pb;
            cut = pb == b.m && b.more != 0;
        }
        o.m = r + 1;
    }
    o.more = (pa < acc.m || pb < b.m || acc.more != 0 || b.more != 0) ? 1 : 0;
    acc = o;
}

}  // namespace ic
