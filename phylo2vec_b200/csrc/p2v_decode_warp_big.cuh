// Warp-per-tree decode for BATCHES of big trees (16384 < k <= 262143) -- the same backward slot assignment and
// forward relabel sweep as p2v_decode_warp.cuh, laid out so that dozens of such trees fit an SM:
//   * shared memory holds the free-slot bit mask only (k / 8 bytes) plus one staged batch of rows;
//   * lane l owns the mask words [l * wpl, (l + 1) * wpl) and keeps THE NUMBER of free slots among them in a
//     register (not the words: wpl goes up to 64); a batch's 32 ranks are located with a warp scan of those
//     counts, then a walk over the owner lane's words, four at a time, straight from shared memory;
//   * v is read from global memory a pair of batches ahead of the rank scan (and validated on the way);
//   * the (time, value) of every slot is parked as a 64-bit word at the END of the tree's own output rows (row s is
//     written after word s was read and only reaches words below s) and read back, coalesced, by the sweep;
//   * "last row so far whose label is x" lives in global memory (one table per resident warp, L2; uint16 entries
//     while slots fit 16 bits, uint32 above).
// The merge kernel (decode_large_kernel) does O(k log^2 k) work per tree; this one O(32 k) like the small kernel.
#pragma once
#include "p2v_decode_warp.cuh"

namespace p2v {

// words of the mask a lane owns: a multiple of four (the walk reads 128 bits at a time)
__host__ __device__ inline int decode_big_wpl(int K) { return ((((K >> 5) + 31) >> 5) + 3) & ~3; }
// uint16 entries of shared memory per warp: mask words (2 u16 each), per-lane placement counters, a staged batch
__host__ __device__ inline size_t decode_big_scratch_u16(int K) { return (size_t)2 * 32 * decode_big_wpl(K) + 64 + 16 + 98 * 4; }

template <typename IdxT, typename OutT, int MODE, typename TblT>
__device__ __forceinline__ int decode_tree_warp_big(const IdxT* __restrict__ vt, int k, int K, OutT* ot, uint16_t* base,
                                                    TblT* __restrict__ gtbl /* K + 32 entries, global */, int lane) {
    constexpr uint32_t TNONE = (sizeof(TblT) == 2) ? 0xFFFFu : 0xFFFFFFFFu;
    const int KA = K + 32;
    const int W = K >> 5;                          // mask words in use
    const int wpl = decode_big_wpl(K);             // words per lane; words W .. 32 * wpl - 1 hold no free slot
    uint32_t* smask = reinterpret_cast<uint32_t*>(base);
    uint32_t* ldec = smask + 32 * wpl;             // 32 counters: placements that landed in lane l's words
    OutT* srow = reinterpret_cast<OutT*>(ldec + 32);   // 16-byte aligned: 128 * wpl + 128 bytes into the warp's scratch
    const int nb = K >> 5;
    constexpr int ROWE = (MODE == MODE_PAIRS || MODE == MODE_ANC2) ? 2 : (MODE == MODE_ANCESTRY ? 3 : 4);
    // (rounded down to 8 bytes: 12-byte int32 ancestry rows of an odd k end on a 4-byte boundary; the words still
    // lie inside the tree's rows and row s still only reaches words <= s)
    unsigned long long* const park = reinterpret_cast<unsigned long long*>(
        (reinterpret_cast<uintptr_t>(ot) + (size_t)ROWE * sizeof(OutT) * k - 8 * (size_t)k) & ~(uintptr_t)7);

    if (MODE != MODE_PAIRS) {                      // "no row yet" for every label
        uint4 ones; ones.x = ones.y = ones.z = ones.w = 0xFFFFFFFFu;
        constexpr int PER = 16 / (int)sizeof(TblT);
        for (int i = lane * PER; i < KA; i += 32 * PER) __stcg(reinterpret_cast<uint4*>(gtbl + i), ones);
    }
    for (int g = lane; g < 32 * wpl; g += 32) smask[g] = (g < W) ? 0xFFFFFFFFu : 0u;   // every slot below K is free (padding slots too)
    ldec[lane] = 0u;
    const int w_lo = min(W, lane * wpl), w_hi = min(W, w_lo + wpl);
    uint32_t cnt = 32u * (uint32_t)(w_hi - w_lo);
    __syncwarp();

    // ---- B: slot assignment, walking time backwards, two batches per rank scan -------------------------------
    uint32_t xv_hi = 0, xv_lo = 0;
    long long raw_hi = 0, raw_lo = 0;
    int badmin = 0x7FFFFFFF;
    auto issue_v = [&](int b) {
        const int t1 = b * 32 + lane, t0 = t1 - 32;
        raw_hi = (b >= 0 && t1 < k) ? (long long)__ldg(vt + t1) : 0;
        raw_lo = (t0 >= 0) ? (long long)__ldg(vt + t0) : 0;
    };
    auto take_v = [&](int b) {
        const int t1 = b * 32 + lane, t0 = t1 - 32;
        const bool bad1 = (unsigned long long)raw_hi > (unsigned long long)(2 * (long long)t1);
        const bool bad0 = t0 >= 0 && (unsigned long long)raw_lo > (unsigned long long)(2 * (long long)t0);
        if (bad1 && t1 < k) badmin = min(badmin, t1);
        if (bad0) badmin = min(badmin, t0);
        xv_hi = bad1 ? 0u : (uint32_t)raw_hi;
        xv_lo = bad0 ? 0u : (uint32_t)raw_lo;
    };
    auto place = [&](int b, uint32_t Rin, uint32_t xval) {
        const int t = b * 32 + lane;
        const uint32_t R = (t < k) ? Rin : (uint32_t)t;           // padding elements take the top slots
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += y;
        }
        uint32_t w = 0;
        { const uint32_t y = __shfl_sync(FULL, incl, 15); if (y <= R) w = 16; }
        { const uint32_t y = __shfl_sync(FULL, incl, w + 7); if (y <= R) w += 8; }
        { const uint32_t y = __shfl_sync(FULL, incl, w + 3); if (y <= R) w += 4; }
        { const uint32_t y = __shfl_sync(FULL, incl, w + 1); if (y <= R) w += 2; }
        { const uint32_t y = __shfl_sync(FULL, incl, w); if (y <= R) w += 1; }
        w &= 31;
        const uint32_t before = __shfl_sync(FULL, incl - cnt, w);
        uint32_t rr = R - before;                                  // rank inside lane w's words
        // walk lane w's words four at a time
        uint32_t g = w * (uint32_t)wpl;
        const uint32_t gend = g + (uint32_t)wpl;
        uint32_t wd = 0;
        bool found = false;
        for (; g < gend && !found; g += 4) {
            const uint4 q = *reinterpret_cast<const uint4*>(smask + g);
            const uint32_t c0 = __popc(q.x), c1 = __popc(q.y), c2 = __popc(q.z), c3 = __popc(q.w);
            if (rr < c0) { wd = q.x; found = true; break; }
            rr -= c0;
            if (rr < c1) { wd = q.y; g += 1; found = true; break; }
            rr -= c1;
            if (rr < c2) { wd = q.z; g += 2; found = true; break; }
            rr -= c2;
            if (rr < c3) { wd = q.w; g += 3; found = true; break; }
            rr -= c3;
        }
        const uint32_t pos = nth_set_bit(wd, rr);
        const int slot = (int)(g * 32 + pos);
        if (found) { P2V_CHECK_INDEX(g, 32 * wpl); P2V_CHECK_INDEX(pos, 32); }
        if (found && slot < k) __stcg(park + slot, (unsigned long long)(uint32_t)t | ((unsigned long long)xval << 32));
        __syncwarp();                                              // every lane has read the mask
        if (found) { atomicAnd(&smask[g], ~(1u << pos)); atomicAdd(&ldec[w], 1u); }
        __syncwarp();
        cnt -= ldec[lane];
        __syncwarp();
        ldec[lane] = 0u;
    };
    int bp = nb - 1;
    uint32_t Rn_hi, Rn_lo;
    issue_v(bp);
    if (nb & 1) {
        take_v(bp); issue_v(bp - 1);
        batch_rank_pair_vals(xv_hi, xv_lo, bp, k, lane, Rn_hi, Rn_lo);
        place(bp, Rn_hi, xv_hi);
        --bp;
    }
    if (bp >= 1) { take_v(bp); issue_v(bp - 2); batch_rank_pair_vals(xv_hi, xv_lo, bp, k, lane, Rn_hi, Rn_lo); }
    for (; bp >= 3; bp -= 2) {
        const uint32_t R1 = Rn_hi, R0 = Rn_lo, X1 = xv_hi, X0 = xv_lo;
        take_v(bp - 2); issue_v(bp - 4);
        batch_rank_pair_vals(xv_hi, xv_lo, bp - 2, k, lane, Rn_hi, Rn_lo);
        place(bp, R1, X1);
        place(bp - 1, R0, X0);
    }
    if (bp >= 1) {
        place(bp, Rn_hi, xv_hi);
        place(bp - 1, Rn_lo, xv_lo);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) badmin = min(badmin, __shfl_xor_sync(FULL, badmin, o));
    if (badmin != 0x7FFFFFFF) return badmin;
    __syncwarp();

    // ---- D: forward sweep over the slots, 32 rows at a time (see p2v_decode_warp.cuh) --------------------------
    uint32_t lab_carry = 0;
    uint32_t n_t = 0, n_lab = 0, n_peers = 0;
    bool n_root = false;
    auto label_stage = [&](int bb, uint32_t t, uint32_t x) {
        const int s = bb * 32 + lane;
        const bool act = s < k;
        n_root = act && x <= t;
        const uint32_t m = __ballot_sync(FULL, n_root) & lanes_le(lane);
        const int src = m ? 31 - __clz(m) : 0;
        uint32_t lab = __shfl_sync(FULL, x, src);
        if (!m) lab = lab_carry;
        lab_carry = __shfl_sync(FULL, lab, 31);
        n_t = t; n_lab = lab;
        if (MODE != MODE_PAIRS) n_peers = __match_any_sync(FULL, act ? lab : (0x40000000u + (uint32_t)lane));
    };
    auto parked = [&](int s) -> unsigned long long { return (s < k) ? __ldcg(park + s) : 0ull; };
    unsigned long long w0 = parked(lane), w1 = parked(32 + lane), w2 = parked(64 + lane), w3 = parked(96 + lane);
    label_stage(0, (uint32_t)w0, (uint32_t)(w0 >> 32));
    for (int b = 0; b < nb; ++b) {
        const int s = b * 32 + lane;
        const bool act = s < k;
        const uint32_t t = n_t, lab = n_lab, peers = n_peers;
        const bool root = n_root;
        const unsigned long long w4 = parked(s + 128);             // four batches ahead
        if (MODE == MODE_PAIRS) {
            if (act) store_pair<OutT>(ot + 2 * (size_t)s, (OutT)lab, (OutT)(t + 1));
            label_stage(b + 1, (uint32_t)w1, (uint32_t)(w1 >> 32));
        } else {
            const uint32_t lower = peers & lanes_lt(lane);
            uint32_t cand = TNONE;
            if (root && !lower) { P2V_CHECK_INDEX(lab, KA); cand = __ldcg(gtbl + lab); }
            label_stage(b + 1, (uint32_t)w1, (uint32_t)(w1 >> 32));          // overlaps the table lookup
            uint32_t c1p = (uint32_t)(k + s);
            if (root) {
                if (lower) cand = (uint32_t)(b * 32 + 31 - __clz(lower));
                c1p = (cand != TNONE) ? (uint32_t)(k + 1) + cand : lab;
            }
            __syncwarp();
            if (act && (peers & lanes_gt(lane)) == 0) { P2V_CHECK_INDEX(lab, KA); __stcg(gtbl + lab, (TblT)s); }
            __syncwarp();
            if (act) P2V_CHECK_INDEX(t + 1, KA);
            const uint32_t q = act ? (uint32_t)__ldcg(gtbl + t + 1) : TNONE;
            const uint32_t c2p = (q != TNONE) ? (uint32_t)(k + 1) + q : t + 1;
            if (MODE == MODE_ANCESTRY) {
                OutT* const dst = ot + 96 * (size_t)b;
                const int a = (int)((reinterpret_cast<uintptr_t>(dst) / sizeof(OutT)) & 1);
                OutT* const st = srow + a;
                st[3 * lane] = (OutT)c1p; st[3 * lane + 1] = (OutT)c2p; st[3 * lane + 2] = (OutT)(k + 1 + s);
                __syncwarp();
                const int lim = 3 * min(32, k - b * 32);
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
                    const int qv = lane + 32 * rep;
                    const int f = 2 * qv - a;
                    if (qv <= 48) {
                        if (f >= 0 && f + 1 < lim) store_pair<OutT>(dst + f, st[f], st[f + 1]);
                        else if (f < 0 && f + 1 < lim) dst[f + 1] = st[f + 1];
                        else if (f >= 0 && f < lim) dst[f] = st[f];
                    }
                }
                __syncwarp();
            } else if (act) {
                store_row<OutT, MODE>(ot, s, (OutT)c1p, (OutT)c2p, (OutT)(k + 1 + s));
            }
        }
        w1 = w2; w2 = w3; w3 = w4;
    }
    __syncwarp();
    return -1;
}

}  // namespace p2v
