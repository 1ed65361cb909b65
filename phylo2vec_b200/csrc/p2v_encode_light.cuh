// Warp-per-tree encoder with two bytes of shared memory per leaf (round 2).
//
// Same algorithm as encode_small_kernel (p2v_encode.cu; reference: build_cherries convert.rs:399-440,
// order_cherries :449-477, build_vector + Fenwick :286-337), laid out so that many more trees fit on an SM:
//
//  * rows are not staged: they are read from global memory two batches ahead of the wavefront;
//  * ONE table T by internal label: NONE until the row defining the label is reached, "pending in lane j"
//    while that row's batch is in flight, the label's smallest descendant leaf once the row has fired
//    (round 1 kept the defining row and the smallest descendant in two tables, plus both relabelled
//    children and the parent of every row: 10 bytes per leaf);
//  * build_vector needs, for the cherry of every row, idx = #{earlier rows with a smaller c_max}.  Round 1 walked
//    the VALUES upwards over a "rows seen" mask, which needs the rows sorted by c_max first (a table, or the
//    (row, c_min) words a first version of this kernel parked in the output row and read back).  The same count
//    taken in ROW order needs nothing of the kind: it is the number of c_max values already seen below this
//    one -- a prefix count on a "values seen" mask -- plus the lower lanes of the batch with a smaller c_max.
//    So a row's v entry is final when its cherry is: ONE pass over the rows, v[c_max - 1] stored from there
//    (scattered; a tree's output row stays in L2 until complete), and the mask doubles as the permutation check;
//  * the mask is k bits; every lane keeps the count of the mask words it owns, so the prefix count is a warp
//    scan, a walk over at most WPL words of the owner, and one partial word.
//
// from_edges scatters (c1, c2) by parent first (order_cherries) into a workspace slice and then runs the same
// sweep over that table.  Above T_SMEM_K_MAX the table T moves to the workspace as 32-bit entries.
#pragma once

#include <type_traits>

#include "p2v_common.cuh"
#include "p2v_internal.h"

namespace p2v {

constexpr int ENC_LIGHT_K_MAX = 65535;       // c_max values of two batches share a shuffle (16-bit fields) up to here
constexpr int ENC_LIGHT_WIDE_K_MAX = 262143; // above: a shuffle per batch (WIDE); 32 KB of mask per tree at the end
constexpr int ENC_LIGHT_T_SMEM_K_MAX = 32767;   // table in shared memory (uint16) up to here: 66 KB, three trees per SM

__host__ __device__ inline int enc_light_wpl(int K) {   // mask words per lane: a power of two >= 4
    const int need = ((K >> 5) + 31) >> 5;
    int w = 4;
    while (w < need) w <<= 1;
    return w;
}
// uint32 words of shared memory per warp
__host__ __device__ inline size_t enc_light_smem_words(int K, bool table_in_smem) {
    return (size_t)32 * enc_light_wpl(K) + 32 + (table_in_smem ? (size_t)(K + 64) / 2 : 0);
}
// uint32 words of workspace per resident warp: (c1, c2) by row for from_edges + the table when it is global
__host__ __device__ inline size_t enc_light_ws_words(int K, bool table_in_smem) {
    return (size_t)2 * (K + 32) + (table_in_smem ? 0 : (size_t)(K + 64));
}

template <bool GT> struct EncTable;
template <> struct EncTable<false> {
    uint16_t* t;
    static constexpr uint32_t NONE = 0xFFFFu;
    __device__ __forceinline__ uint32_t get(uint32_t i) const { return t[i]; }
    __device__ __forceinline__ void set(uint32_t i, uint32_t x) const { t[i] = (uint16_t)x; }
    __device__ __forceinline__ void clear(int n, int lane) const {
        uint32_t* w = reinterpret_cast<uint32_t*>(t);
        for (int i = lane; i < n / 2; i += 32) w[i] = 0xFFFFFFFFu;
    }
};
template <> struct EncTable<true> {
    uint32_t* t;
    static constexpr uint32_t NONE = 0xFFFFFFFFu;
    __device__ __forceinline__ uint32_t get(uint32_t i) const { return __ldcg(t + i); }
    __device__ __forceinline__ void set(uint32_t i, uint32_t x) const { __stcg(t + i, x); }
    __device__ __forceinline__ void clear(int n, int lane) const {
        uint4* w = reinterpret_cast<uint4*>(t);
        for (int i = lane; i < n / 4; i += 32) __stcg(w + i, make_uint4(NONE, NONE, NONE, NONE));
    }
};

__device__ __forceinline__ uint32_t popc4(const uint4 q) { return __popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w); }

// WIDE (k > 65535, table in the workspace): the in-batch counts take one shuffle per batch and step.
template <typename IdxT, int MODE, bool GT, bool WIDE = false>
__global__ void __launch_bounds__(128)
encode_light_kernel(const IdxT* __restrict__ in, int64_t B, int k, IdxT* __restrict__ vout,
                    int32_t* __restrict__ status, int K, uint32_t* __restrict__ ws, size_t ws_words) {
    extern __shared__ __align__(16) uint32_t enc_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int wpl = enc_light_wpl(K), wsh = 31 - __clz(wpl);
    const bool has_par = (MODE != FROM_PAIRS);
    const int in_cols = (MODE == FROM_PAIRS) ? 2 : (MODE == FROM_ANCESTRY ? 3 : 4);
    const uint32_t kint = (uint32_t)k + 1;
    const int nb = K >> 5;
    const int TN = K + 64;                                         // table entries (labels k+1 .. 2k+1 -> 0 .. k)

    uint32_t* mask = enc_sm + (size_t)warp * enc_light_smem_words(K, !GT);
    uint32_t* lcs = mask + 32 * wpl;                               // seen rows among the words lane L owns
    uint32_t* slice = ws + ((size_t)blockIdx.x * nwarps + warp) * ws_words;
    uint32_t* E = slice;                                           // from_edges: (c1, c2) by row
    EncTable<GT> T;
    if (GT) T.t = reinterpret_cast<decltype(T.t)>(slice + 2 * (size_t)(K + 32));
    else T.t = reinterpret_cast<decltype(T.t)>(lcs + 32);
    constexpr uint32_t NONE = EncTable<GT>::NONE;
    constexpr uint32_t PEND0 = NONE - 32;                          // NONE - 1 - j: the row of lane j has not fired

    for (int64_t tree = (int64_t)blockIdx.x * nwarps + warp; tree < B; tree += (int64_t)gridDim.x * nwarps) {
        const IdxT* it = in + tree * ((int64_t)k * in_cols);
        IdxT* vo = vout + tree * k;
        bool bad = false;
        for (int j = 0; j < wpl; ++j) mask[lane * wpl + j] = 0;
        lcs[lane] = 0;
        if (has_par) T.clear(TN, lane);
        __syncwarp();

        long long offset = 0;
        if (MODE == FROM_EDGES) {
            // order_cherries (convert.rs:449-477): row -> index parent - k - offset, offset = max parent - 2k + 1
            const long long top = 2LL * k + 1;
            uint32_t mx = 0;
            for (int b0 = 0; b0 < nb; b0 += 4) {
                long long pv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int s = (b0 + j) * 32 + lane;
                    pv[j] = (s < k) ? (long long)__ldg(it + 4 * (size_t)s + 1) : 0;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (pv[j] < 0 || pv[j] > top) bad = true;
                    else mx = max(mx, (uint32_t)pv[j]);
                }
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, d));
            offset = (long long)mx - 2LL * k + 1;
            for (int b0 = 0; b0 < nb; b0 += 2) {
                long long a[2], c[2], p[2];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int s = (b0 + j) * 32 + lane;
                    a[j] = c[j] = p[j] = 0;
                    if (s < k) {
                        a[j] = (long long)__ldg(it + 4 * (size_t)s);
                        p[j] = (long long)__ldg(it + 4 * (size_t)s + 1);
                        c[j] = (long long)__ldg(it + 4 * (size_t)s + 2);
                    }
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int s = (b0 + j) * 32 + lane;
                    if (s >= k) continue;
                    const long long idx = p[j] - k - offset;
                    if (a[j] < 0 || a[j] > top || c[j] < 0 || c[j] > top || p[j] < 0 || idx < 0 || idx >= k) { bad = true; continue; }
                    const uint32_t bit = 1u << (idx & 31);
                    P2V_CHECK_INDEX(idx >> 5, 32 * wpl);
                    if (atomicOr(&mask[idx >> 5], bit) & bit) bad = true;               // two rows, one parent
                    __stcg(reinterpret_cast<uint2*>(E) + idx, make_uint2((uint32_t)a[j], (uint32_t)c[j]));
                }
            }
            bad = __any_sync(FULL, bad);
            __syncwarp();
            for (int j = 0; j < wpl; ++j) mask[lane * wpl + j] = 0;
            __syncwarp();
        }

        // ---- ONE pass over the rows, in order, two batches at a time: relabel by smallest descendant
        //      (build_cherries), then build_vector's index of the cherry right away --
        //          idx = #{earlier rows with a smaller c_max}
        //      = the c_max values already seen that lie below this one (prefix count on the "values seen" mask,
        //      state before the batch) + the lower lanes of the batch with a smaller c_max; v[c_max - 1] is stored
        //      from here (a scattered store: a tree's output row stays in L2 until it is complete).  The same
        //      mask is the permutation check: a value seen twice is a malformed tree.
        if (!bad) {
            long long rcur[3], rnxt[3];
            auto issue = [&](int b, long long (&r)[3]) {
                const int s = b * 32 + lane;
                r[0] = r[1] = r[2] = 0;
                if (s < k) {
                    if (MODE == FROM_EDGES) {
                        const uint2 q = __ldcg(reinterpret_cast<const uint2*>(E) + s);
                        r[0] = q.x; r[1] = q.y; r[2] = (long long)s + k + offset;
                    } else {
                        r[0] = (long long)__ldg(it + (size_t)in_cols * s);
                        r[1] = (long long)__ldg(it + (size_t)in_cols * s + 1);
                        if (MODE == FROM_ANCESTRY) r[2] = (long long)__ldg(it + (size_t)in_cols * s + 2);
                    }
                }
            };
            issue(0, rcur);
            issue(1, rnxt);
            // one batch of rows through build_cherries: (c_max - 1, c_min) of the lane's row; false = malformed (uniform)
            auto cherries = [&](int b, uint32_t& vmax, uint32_t& vmin) -> bool {
                const int s = b * 32 + lane;
                const bool act = s < k;
                const long long a = rcur[0], c = rcur[1], pl = rcur[2];
                rcur[0] = rnxt[0]; rcur[1] = rnxt[1]; rcur[2] = rnxt[2];
                issue(b + 2, rnxt);
                const long long top = (MODE == FROM_PAIRS) ? (long long)k : 2LL * k + 1;
                bool ok = !(a < 0 || a > top || c < 0 || c > top || pl < 0 || pl > 2LL * k + 1);
                const uint32_t c1 = (uint32_t)a, c2 = (uint32_t)c, p = (uint32_t)pl;
                uint32_t m1 = c1, m2 = c2;
                if (has_par) {
                    if (ok && act && p < kint) ok = false;         // a leaf label as parent: not a tree in this format
                    const uint32_t pi = p - kint;
                    const bool live = act && ok;
                    if (live) {
                        P2V_CHECK_INDEX(pi, TN);
                        if (T.get(pi) != NONE) ok = false;         // the label already has a defining row
                        T.set(pi, NONE - 1 - lane);
                    }
                    __syncwarp();
                    if (live && T.get(pi) != NONE - 1 - (uint32_t)lane) ok = false;    // ... or another one in this batch
                    if (!__all_sync(FULL, ok || !act)) return false;
                    int j1 = -1, j2 = -1;                          // lanes this row waits for
                    if (act) {
                        if (c1 >= kint) {
                            P2V_CHECK_INDEX(c1 - kint, TN);
                            const uint32_t t = T.get(c1 - kint);
                            if (t >= PEND0) { if (t != NONE && (int)(NONE - 1 - t) < lane) j1 = (int)(NONE - 1 - t); }
                            else m1 = t;
                        }
                        if (c2 >= kint) {
                            P2V_CHECK_INDEX(c2 - kint, TN);
                            const uint32_t t = T.get(c2 - kint);
                            if (t >= PEND0) { if (t != NONE && (int)(NONE - 1 - t) < lane) j2 = (int)(NONE - 1 - t); }
                            else m2 = t;
                        }
                    }
                    uint32_t done = __ballot_sync(FULL, !act);
                    // A row fires when the rows defining its internal children have fired; a run of rows that
                    // each only wait for the row just before them is resolved in one step by a segmented
                    // min-scan (see encode_small_kernel).
                    while (done != FULL) {
                        const bool mine_pending = !((done >> lane) & 1u);
                        const bool w1 = j1 >= 0 && !((done >> j1) & 1u);
                        const bool w2 = j2 >= 0 && !((done >> j2) & 1u);
                        const bool prev1 = w1 && j1 + 1 == lane, prev2 = w2 && j2 + 1 == lane;
                        const bool other = (w1 && !prev1) || (w2 && !prev2) || (prev1 && prev2);
                        const bool avail = mine_pending && !other;
                        const uint32_t A = __ballot_sync(FULL, avail);
                        const uint32_t Lk = __ballot_sync(FULL, avail && (prev1 || prev2));
                        const uint32_t H = A & ~Lk;
                        const uint32_t Tl = (H << 1) & Lk;
                        const uint32_t Rm = H | (((Lk + Tl) ^ Lk) & Lk);
                        const bool mine = (Rm >> lane) & 1u;
                        uint32_t o1 = NONE32, o2 = NONE32;
                        if (mine) {
                            if (!prev1) o1 = (j1 >= 0) ? T.get(c1 - kint) : m1;
                            if (!prev2) o2 = (j2 >= 0) ? T.get(c2 - kint) : m2;
                        }
                        uint32_t val = min(o1, o2);
                        const uint32_t hb = H & lanes_le(lane);
                        const int dist = hb ? lane - (31 - __clz(hb)) : 0;
#pragma unroll
                        for (int d = 1; d < 32; d <<= 1) {
                            const uint32_t y = __shfl_up_sync(FULL, val, d);
                            if (mine && dist >= d) val = min(val, y);
                        }
                        const uint32_t pm = __shfl_up_sync(FULL, val, 1);
                        if (mine) {
                            m1 = prev1 ? pm : o1; m2 = prev2 ? pm : o2;
                            T.set(pi, val);
                        }
                        __syncwarp();
                        done |= Rm;
                    }
                } else if (!__all_sync(FULL, ok || !act)) return false;
                const uint32_t cm = max(m1, m2);
                vmin = min(m1, m2);
                vmax = cm - 1u;
                return __all_sync(FULL, !act || (cm >= 1u && cm <= (uint32_t)k));
            };
            uint32_t word1 = 0;                                        // K <= 1024: mask word `lane`
            for (int bp = 0; bp < nb; bp += 2) {
                uint32_t va = 0, la = 0, vb = 0, lb = 0;               // c_max - 1 and c_min of the lane's rows
                if (!cherries(bp, va, la)) { bad = true; break; }
                if (bp + 1 < nb && !cherries(bp + 1, vb, lb)) { bad = true; break; }
                const bool act_a = bp * 32 + lane < k, act_b = (bp + 1) * 32 + lane < k;
                uint32_t inb_a = 0, inb_b = 0;                           // lower lanes of the batch with a smaller c_max
                if (WIDE) {                                              // values of up to 18 bits: a shuffle per batch and step
                    const uint32_t ka = act_a ? va : 0x7FFFFFFFu, kb = act_b ? vb : 0x7FFFFFFFu;
#pragma unroll
                    for (int d = 1; d < 32; ++d) {
                        count_smaller_step(inb_a, ka, d);
                        count_smaller_step(inb_b, kb, d);
                    }
                } else if (K <= 32768) {
                    // 15-bit values: both batches compared by ONE subtraction per step.  Field f of (M - y) keeps bit 15
                    // iff mine_f - 1 >= y_f (no borrow crosses a field: 0x8000 + mine_f - 1 - y_f >= 0).
                    const uint32_t ya = act_a ? va : 0x7FFFu, yb = act_b ? vb : 0x7FFFu;
                    const uint32_t packed = ya | (yb << 16);
                    const uint32_t M = (ya + 0x7FFFu) | ((yb + 0x7FFFu) << 16);
                    uint32_t acc = 0;
#pragma unroll
                    for (int d = 1; d < 32; ++d) {
                        const uint32_t y = __shfl_up_sync(FULL, packed, d);    // below lane d: the lane's own value
                        acc += ((M - y) >> 15) & 0x00010001u;
                    }
                    inb_a = acc & 0xFFFFu; inb_b = acc >> 16;
                } else {
                    const uint32_t ka = act_a ? va : 0xFFFFu, kb = act_b ? vb : 0xFFFFu;
                    const uint32_t packed = ka | (kb << 16);
#pragma unroll
                    for (int d = 1; d < 32; ++d) {
                        const uint32_t y = __shfl_up_sync(FULL, packed, d);
                        inb_a += ((y & 0xFFFFu) < ka);
                        inb_b += ((y >> 16) < kb);
                    }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const bool act = h ? act_b : act_a;
                    const uint32_t val = act ? (h ? vb : va) : 0u;
                    const uint32_t lo = h ? lb : la;
                    const uint32_t g = val >> 5;
                    const uint32_t bit = 1u << (val & 31);
                    uint32_t idx = h ? inb_b : inb_a;
                    if (K <= 1024) {
                        // at most 32 mask words: lane l keeps word l in a register -- its popcount is the lane's count, the
                        // word a value falls into comes by one shuffle, and there is nothing to walk
                        const uint32_t cnt = __popc(word1);
                        const uint32_t excl = warp_incl_scan(cnt, lane) - cnt;
                        idx += __shfl_sync(FULL, excl, g) + __popc(__shfl_sync(FULL, word1, g) & (bit - 1u));
                        if (act) {
                            vo[val] = (IdxT)(idx == 0 ? lo : val + idx);
                            if (atomicOr(&mask[g], bit) & bit) bad = true;   // c_max must be a permutation of 1 .. k
                        }
                        __syncwarp();
                        word1 = mask[lane];
                        continue;
                    }
                    const uint32_t cnt = lcs[lane];
                    const uint32_t excl = warp_incl_scan(cnt, lane) - cnt;
                    const uint32_t src = g >> wsh;
                    P2V_CHECK_INDEX(g, 32 * wpl); P2V_CHECK_INDEX(src, 32);
                    idx += __shfl_sync(FULL, excl, src);
                    const uint4* q = reinterpret_cast<const uint4*>(mask + ((size_t)src << wsh));
                    const int nq = (int)(g - (src << wsh)) >> 2;
                    for (int j = 0; j < nq; ++j) idx += popc4(q[j]);
                    {
                        const uint4 t = q[nq];
                        const uint32_t jj = g & 3u, low = bit - 1u;
                        const uint32_t e0 = t.x, e1 = t.y, e2 = t.z, e3 = t.w;
                        idx += jj == 0 ? __popc(e0 & low)
                             : jj == 1 ? __popc(e0) + __popc(e1 & low)
                             : jj == 2 ? __popc(e0) + __popc(e1) + __popc(e2 & low)
                                       : __popc(e0) + __popc(e1) + __popc(e2) + __popc(e3 & low);
                    }
                    __syncwarp();                                        // every lane has read the mask
                    if (act) {
                        vo[val] = (IdxT)(idx == 0 ? lo : val + idx);
                        if (atomicOr(&mask[g], bit) & bit) bad = true;   // c_max must be a permutation of 1 .. k
                        atomicAdd(&lcs[src], 1u);
                    }
                    __syncwarp();
                }
            }
        }
        bad = __any_sync(FULL, bad);
        __syncwarp();
        if (lane == 0 && status) status[tree] = bad ? TREE_MALFORMED : 0;
        __syncwarp();
    }
}

}  // namespace p2v
