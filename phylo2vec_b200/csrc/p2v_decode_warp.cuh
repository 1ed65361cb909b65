// Warp-per-tree decode of one tree with all state in shared memory (sm_100a); shared by the
// batched decode kernel (p2v_decode.cu) and the fused small-tree distance kernel
// (p2v_cophenetic.cu).  See p2v_decode.cu for the formulation.
#pragma once
#include <cuda_fp16.h>

#include "p2v_common.cuh"
#include "p2v_internal.h"

// 1: issue the rank scan of the next pair of batches between the dependent shuffles of the current
// placements.  Measured on B200 (profiles/decode_timing.py): no gain at n = 1000 (the kernel is
// bound by issue slots and the shared-memory pipe, not by that chain), -17 % at n = 4096; off.
#ifndef P2V_DECODE_INTERLEAVE
#define P2V_DECODE_INTERLEAVE 0
#endif
// 1: clear the free-mask bits of the batch's densest word with one REDUX.OR instead of ATOMS.AND
#ifndef P2V_DECODE_REDUX
#define P2V_DECODE_REDUX 1
#endif

namespace p2v {

constexpr uint32_t BIG = 0x7FFFFFFFu;

template <typename IdxT> struct Vec2;
template <> struct Vec2<long long> { using type = longlong2; };
template <> struct Vec2<int> { using type = int2; };
template <typename IdxT> struct Vec4;
template <> struct Vec4<int> { using type = int4; };

template <typename IdxT>
__device__ __forceinline__ void store_pair(IdxT* p, IdxT a, IdxT b) {
    typename Vec2<IdxT>::type v;
    v.x = a; v.y = b;
    *reinterpret_cast<typename Vec2<IdxT>::type*>(p) = v;
}

template <typename IdxT, int MODE>
__device__ __forceinline__ void store_row(IdxT* tree_out, int s, IdxT c1p, IdxT c2p, IdxT par) {
    if (MODE == MODE_ANCESTRY) {
        IdxT* o = tree_out + 3 * (size_t)s;
        o[0] = c1p; o[1] = c2p; o[2] = par;
    } else if (MODE == MODE_ANC2) {
        store_pair<IdxT>(tree_out + 2 * (size_t)s, c1p, c2p);
    } else {  // MODE_EDGES: (c1p,par),(c2p,par)
        IdxT* o = tree_out + 4 * (size_t)s;
        store_pair<IdxT>(o, c1p, par);
        store_pair<IdxT>(o + 2, c2p, par);
    }
}

// --------------------------------------------------------------------------
// k <= 1024: one warp per tree, all state in shared memory (uint16).
// --------------------------------------------------------------------------

// Rank of element t = 32b+lane among the elements [0, 32b+32) at the end of its batch.
// Lane i scans the later lanes i+1.. in time order: R = ins_i; R += (ins_j <= R).  SHFL.DOWN past
// lane 31 hands back the lane's own value, which always compares <= R, so the `lane` wrapped
// steps are taken off at the end (they all come after the real ones).
// Two batches (b and b-1) are scanned at once in the two halves of a half2: all values are
// integers <= 1024, exact in fp16, so one 32-bit shuffle, one HSET2.LE and one HADD2 advance
// both scans.  The 31 shuffles do not depend on R and are issued back to back.
__device__ __forceinline__ void batch_rank_pair(const uint16_t* __restrict__ sv, int b, int k, int lane,
                                                uint32_t& R_hi, uint32_t& R_lo) {
    const int t1 = b * 32 + lane, t0 = t1 - 32;            // batch b and batch b-1 (t0 < 0 if b == 0)
    const uint32_t x1 = sv[t1];
    const uint32_t x0 = (t0 >= 0) ? (uint32_t)sv[t0] : 0u;
    const __half inf = __ushort_as_half((unsigned short)0x7C00);
    const __half i1 = (t1 < k) ? __uint2half_rn(x1 <= (uint32_t)t1 ? 0u : x1 - (uint32_t)t1) : inf;
    const __half i0 = (t0 >= 0) ? __uint2half_rn(x0 <= (uint32_t)t0 ? 0u : x0 - (uint32_t)t0) : inf;
    const __half2 ins = __halves2half2(i0, i1);
    __half2 R = ins;
#pragma unroll
    for (int d = 1; d < 32; ++d) {
        const __half2 y = __shfl_down_sync(FULL, ins, d);
        R = __hadd2(R, __hle2(y, R));
    }
    R_lo = (uint32_t)__half2int_rz(__low2half(R)) - (uint32_t)lane;    // garbage for inactive lanes
    R_hi = (uint32_t)__half2int_rz(__high2half(R)) - (uint32_t)lane;
}

// The same scan with 32-bit integers (SHFL.DOWN + ISETP + predicated IADD per step and batch) for
// trees whose ranks do not fit the exact range of fp16.  rank_step uses SHFL's in-range predicate,
// so there are no wrapped steps to take off.
__device__ __forceinline__ void batch_rank_pair_int(const uint16_t* __restrict__ sv, int b, int k, int lane,
                                                    uint32_t& R_hi, uint32_t& R_lo) {
    const int t1 = b * 32 + lane, t0 = t1 - 32;
    const uint32_t x1 = sv[t1];
    const uint32_t x0 = (t0 >= 0) ? (uint32_t)sv[t0] : 0u;
    const uint32_t i1 = (t1 < k) ? (x1 <= (uint32_t)t1 ? 0u : x1 - (uint32_t)t1) : BIG;
    const uint32_t i0 = (t0 >= 0) ? (x0 <= (uint32_t)t0 ? 0u : x0 - (uint32_t)t0) : BIG;
    uint32_t R1 = i1, R0 = i0;
#pragma unroll
    for (int d = 1; d < 32; ++d) {
        rank_step(R1, i1, d);
        rank_step(R0, i0, d);
    }
    R_hi = R1; R_lo = R0;
}

// The integer scan with the two batches' values handed in (the light form reads v from global memory).
__device__ __forceinline__ void batch_rank_pair_vals(uint32_t x1, uint32_t x0, int b, int k, int lane,
                                                     uint32_t& R_hi, uint32_t& R_lo) {
    const int t1 = b * 32 + lane, t0 = t1 - 32;
    const uint32_t i1 = (t1 < k) ? (x1 <= (uint32_t)t1 ? 0u : x1 - (uint32_t)t1) : BIG;
    const uint32_t i0 = (t0 >= 0) ? (x0 <= (uint32_t)t0 ? 0u : x0 - (uint32_t)t0) : BIG;
    uint32_t R1 = i1, R0 = i0;
#pragma unroll
    for (int d = 1; d < 32; ++d) {
        rank_step(R1, i1, d);
        rank_step(R0, i0, d);
    }
    R_hi = R1; R_lo = R0;
}

// The same for k <= 16384 with ONE shuffle per step for both batches: the two 15-bit values share a word, and
// field f of ((R | 0x8000 in both fields) - y) keeps its bit 15 iff y_f <= R_f; no borrow crosses the fields
// (0x8000 + R_f - y_f >= 0) and no carry does either (R_f stays below 2^15 + 64).  SHFL.DOWN past lane 31
// hands back the lane's own value, which always counts: those `lane` steps come last and are taken off.
__device__ __forceinline__ void batch_rank_pair_vals15(uint32_t x1, uint32_t x0, int b, int k, int lane,
                                                       uint32_t& R_hi, uint32_t& R_lo) {
    const int t1 = b * 32 + lane, t0 = t1 - 32;
    const uint32_t i1 = (t1 < k) ? (x1 <= (uint32_t)t1 ? 0u : x1 - (uint32_t)t1) : 0x7FFFu;
    const uint32_t i0 = (t0 >= 0) ? (x0 <= (uint32_t)t0 ? 0u : x0 - (uint32_t)t0) : 0x7FFFu;
    const uint32_t ins = i0 | (i1 << 16);
    uint32_t R = ins;
#pragma unroll
    for (int d = 1; d < 32; ++d) {
        const uint32_t y = __shfl_down_sync(FULL, ins, d);
        R += (((R | 0x80008000u) - y) >> 15) & 0x00010001u;
    }
    R_lo = (R & 0xFFFFu) - (uint32_t)lane;                 // garbage for inactive lanes
    R_hi = (R >> 16) - (uint32_t)lane;
}

// The same shuffles, but the two fields are taken apart before the compare: two independent chains of
// (ISETP, predicated IADD) per step instead of one chain of five dependent operations.  One more instruction per
// step, half the dependent latency: for the trees that leave few warps on an SM (k > 8192) latency is what counts.
__device__ __forceinline__ void batch_rank_pair_vals15s(uint32_t x1, uint32_t x0, int b, int k, int lane,
                                                        uint32_t& R_hi, uint32_t& R_lo) {
    const int t1 = b * 32 + lane, t0 = t1 - 32;
    const uint32_t i1 = (t1 < k) ? (x1 <= (uint32_t)t1 ? 0u : x1 - (uint32_t)t1) : 0x7FFFu;
    const uint32_t i0 = (t0 >= 0) ? (x0 <= (uint32_t)t0 ? 0u : x0 - (uint32_t)t0) : 0x7FFFu;
    const uint32_t ins = i0 | (i1 << 16);
    uint32_t R0 = i0, R1 = i1;
#pragma unroll
    for (int d = 1; d < 32; ++d) {
        const uint32_t y = __shfl_down_sync(FULL, ins, d);
        const uint32_t y0 = y & 0xFFFFu, y1 = y >> 16;
        if (y0 <= R0) ++R0;
        if (y1 <= R1) ++R1;
    }
    R_lo = R0 - (uint32_t)lane;                            // the `lane` wrapped steps always count
    R_hi = R1 - (uint32_t)lane;
}

// uint16 entries of per-tree scratch the warp needs: sv, stime, (stbl), then the free-slot mask.
// Mid-size trees (wpl > 2, the "light" form): only stbl and the mask live in shared memory -- v is read
// from global memory batch by batch, and the (time, value) of every slot goes through the tree's own
// output rows as a 32-bit scratch word -- so that three times as many trees fit an SM.
__host__ __device__ inline size_t decode_warp_scratch_u16(int K, int mode, int wpl) {
    if (wpl > 2) return (size_t)((mode == MODE_PAIRS) ? 0 : 1) * (K + 32) + 64 * wpl + 16 + 98 * 4;   // + a staged batch of rows

    return (size_t)((mode == MODE_PAIRS) ? 2 : 3) * (K + 32) + 64 * wpl;
}

// Decodes v (k entries, K = k rounded up to 32) into `ot` (pairs / ancestry / edges rows of
// OutT; any address space).  Returns -1, or the first index i with v[i] outside [0, 2i]
// (check_v, vector/base.rs:50-67), in which case nothing is written.
// WPL = free-mask words per lane: 1 for k <= 1024, 2 for k <= 2016 (the half2 rank scan is exact
// for integers up to 2048 and overshoots the true rank by at most 31 before the correction);
// 4 / 8 / 16 for mid-size trees (k <= 4096 / 8192 / 16384) with the integer rank scan.
// VSTAGED: vt points into shared memory (a staged copy of the row), so plain loads instead of LDG.
template <typename IdxT, typename OutT, int MODE, int WPL, bool VSTAGED = false>
__device__ __forceinline__ int decode_tree_warp(const IdxT* __restrict__ vt, int k, int K, OutT* ot,
                                                uint16_t* base, int lane, const uint8_t* __restrict__ lut) {
    auto ldv = [&](int i) -> long long { return VSTAGED ? (long long)vt[i] : (long long)__ldg(vt + i); };
    const int KA = K + 32;
    constexpr bool LIGHT = WPL > 2;                 // k <= 16384: time < 2^14 and v[t] < 2^15 share a 32-bit word
    const int narr = LIGHT ? ((MODE == MODE_PAIRS) ? 0 : 1) : ((MODE == MODE_PAIRS) ? 2 : 3);
    uint16_t* sv = base;            // v[t]                                   (by time)            [!LIGHT]
    uint16_t* stime = sv + KA;      // element at slot s                      (by slot)            [!LIGHT]
    uint16_t* stbl = LIGHT ? base : stime + KA;    // last row so far whose label is x (by label)  [!PAIRS]
    uint32_t* smask = reinterpret_cast<uint32_t*>(base + narr * KA);  // 32*WPL words of free-slot bits
    const int nb = K >> 5;
    // light form: slot s parks (time | v[time] << 14) in the first four bytes of its own output row until the
    // forward sweep reads it back (three batches ahead of the rows it writes)
    // The words sit packed at the END of the tree's rows: row s is written after word s was read, and it only
    // reaches words below s (its bytes start at R s, word j at R k - 4 k + 4 j, R >= 8 bytes per row).
    constexpr int ROWE = (MODE == MODE_PAIRS || MODE == MODE_ANC2) ? 2 : (MODE == MODE_ANCESTRY ? 3 : 4);
    uint32_t* const park_base = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ot) + (size_t)ROWE * sizeof(OutT) * k - 4 * (size_t)k);
    auto park = [&](int s) -> uint32_t* { return park_base + s; };
    // light form, ancestry rows (24 / 12 bytes, three scalar stores at that stride each): a batch of rows is
    // staged in shared memory and leaves as whole two-element vectors -- every 32-byte sector is written once
    OutT* srow = reinterpret_cast<OutT*>(smask + 32 * WPL);       // 16-byte aligned: every piece before it is
    // ---- A: load (8 loads in flight per lane), validate (check_v), narrow ----------
    if (MODE != MODE_PAIRS) {                   // "no row yet" for every label, 8 entries per store
        uint4 ones; ones.x = ones.y = ones.z = ones.w = 0xFFFFFFFFu;
        for (int i = lane * 8; i < KA; i += 256) *reinterpret_cast<uint4*>(stbl + i) = ones;
    }
    bool anybad = false;
    for (int b0 = 0; b0 < (LIGHT ? 0 : nb); b0 += 8) {          // the light form validates in phase B
        long long xs[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = (b0 + j) * 32 + lane;
            xs[j] = (i < k) ? ldv(i) : 0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = (b0 + j) * 32 + lane;
            anybad |= (unsigned long long)xs[j] > (unsigned long long)(2 * i);     // also catches negatives
            if (!LIGHT && b0 + j < nb) sv[i] = (uint16_t)xs[j];
        }
    }
    if (__any_sync(FULL, anybad)) {             // rare: find the first offending index
        for (int b = 0; b < nb; ++b) {
            const int i = b * 32 + lane;
            const bool ok = (i >= k) || (unsigned long long)ldv(i) <= (unsigned long long)(2 * i);
            const unsigned bm = __ballot_sync(FULL, !ok);
            if (bm) return b * 32 + __ffs(bm) - 1;
        }
    }
    // free-slot mask: bit i of word g <-> slot 32g+i, 1 = free; lane l holds words l*WPL .. l*WPL+WPL-1.
    // The K - k padding elements (times k .. K-1, top batch only) are placed like real ones: with
    // rank t they take slots k .. K-1 before anything else is placed, so the placement code needs
    // no "is this lane a real element" predicate.
    uint32_t word[WPL];
#pragma unroll
    for (int j = 0; j < WPL; ++j) {
        const int g = lane * WPL + j;
        word[j] = (g * 32 < K) ? 0xFFFFFFFFu : 0u;
        smask[g] = word[j];
    }
    __syncwarp();

    // ---- B: slot assignment, walking time backwards, two batches per rank scan; the scan
    //         of the next pair is independent of the placement of the current one ------------
    // light form: v of the batches straight from global memory, one pair of batches ahead of the scan that uses
    // it (the loads have a whole pair of placements to land); check_v on the way -- an invalid entry counts as 0
    // here and the smallest offending index is reported at the end (the rows of such a tree are garbage).
    uint32_t xv_hi = 0, xv_lo = 0;                 // v[t] of the lane's element in the scanned batches
    long long raw_hi = 0, raw_lo = 0;              // the pair of batches loaded ahead
    int badmin = 0x7FFFFFFF;
    auto issue_v = [&](int b) {                    // batches b (hi) and b - 1 (lo); b may be < 1 past the end
        const int t1 = b * 32 + lane, t0 = t1 - 32;
        raw_hi = (b >= 0 && t1 < k) ? ldv(t1) : 0;
        raw_lo = (t0 >= 0) ? ldv(t0) : 0;
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
    auto rank_pair = [&](int b, uint32_t& hi, uint32_t& lo) {      // batches b (hi) and b - 1 (lo)
        if (WPL <= 2) batch_rank_pair(sv, b, k, lane, hi, lo);
        else {
            take_v(b); issue_v(b - 2);
            if (WPL >= 16) batch_rank_pair_vals15s(xv_hi, xv_lo, b, k, lane, hi, lo);
            else batch_rank_pair_vals15(xv_hi, xv_lo, b, k, lane, hi, lo);
        }
    };
    // The rank scan of the NEXT pair of batches is resumable (scan_begin / scan_step / scan_end) and
    // its 31 steps are issued between the dependent shuffles of the two placements of the CURRENT
    // pair (`fill(i)` at the 13 stall points of a placement), so that the two chains overlap.
    __half2 h_ins = __float2half2_rn(0.0f), h_R = h_ins;
    uint32_t i_ins1 = 0, i_ins0 = 0, i_R1 = 0, i_R0 = 0;
    auto scan_begin = [&](int b) {                 // batches b (hi half) and b - 1 (lo half)
        const int t1 = b * 32 + lane, t0 = t1 - 32;
        const uint32_t x1 = sv[t1];
        const uint32_t x0 = (t0 >= 0) ? (uint32_t)sv[t0] : 0u;
        const uint32_t a1 = x1 <= (uint32_t)t1 ? 0u : x1 - (uint32_t)t1;
        const uint32_t a0 = x0 <= (uint32_t)t0 ? 0u : x0 - (uint32_t)t0;
        if (WPL <= 2) {
            const __half inf = __ushort_as_half((unsigned short)0x7C00);
            h_ins = __halves2half2((t0 >= 0) ? __uint2half_rn(a0) : inf, (t1 < k) ? __uint2half_rn(a1) : inf);
            h_R = h_ins;
        } else {
            i_ins1 = (t1 < k) ? a1 : BIG; i_ins0 = (t0 >= 0) ? a0 : BIG;
            i_R1 = i_ins1; i_R0 = i_ins0;
        }
    };
    auto scan_step = [&](int d) {
        if (WPL <= 2) {
            const __half2 y = __shfl_down_sync(FULL, h_ins, d);
            h_R = __hadd2(h_R, __hle2(y, h_R));
        } else {
            rank_step(i_R1, i_ins1, d);
            rank_step(i_R0, i_ins0, d);
        }
    };
    auto scan_end = [&](uint32_t& hi, uint32_t& lo) {
        if (WPL <= 2) {      // the `lane` wrapped steps (SHFL.DOWN past lane 31 returns the own value)
            lo = (uint32_t)__half2int_rz(__low2half(h_R)) - (uint32_t)lane;
            hi = (uint32_t)__half2int_rz(__high2half(h_R)) - (uint32_t)lane;
        } else {
            hi = i_R1; lo = i_R0;
        }
    };
    auto fill_none = [&](int) {};
    auto fill_first = [&](int i) {                 // steps 1..16 over the 13 points of the first placement
        if (i < 3) { scan_step(2 * i + 1); scan_step(2 * i + 2); }
        else scan_step(i + 4);
    };
    auto fill_second = [&](int i) {                // steps 17..31 over the second placement
        if (i < 2) { scan_step(17 + 2 * i); scan_step(18 + 2 * i); }
        else scan_step(i + 19);
    };
    auto place = [&](int b, uint32_t Rin, uint32_t xval, auto&& fill) {
            const int t = b * 32 + lane;
            const uint32_t R = (t < k) ? Rin : (uint32_t)t;
            // rank-select R among the free slots
            uint32_t cnt = 0;
#pragma unroll
            for (int j = 0; j < WPL; ++j) cnt += __popc(word[j]);
            uint32_t incl = cnt;
            { const uint32_t y = __shfl_up_sync(FULL, incl, 1); fill(0); if (lane >= 1) incl += y; }
            { const uint32_t y = __shfl_up_sync(FULL, incl, 2); fill(1); if (lane >= 2) incl += y; }
            { const uint32_t y = __shfl_up_sync(FULL, incl, 4); fill(2); if (lane >= 4) incl += y; }
            { const uint32_t y = __shfl_up_sync(FULL, incl, 8); fill(3); if (lane >= 8) incl += y; }
            { const uint32_t y = __shfl_up_sync(FULL, incl, 16); fill(4); if (lane >= 16) incl += y; }
            uint32_t w = 0;
            { const uint32_t y = __shfl_sync(FULL, incl, 15); fill(5); if (y <= R) w = 16; }
            { const uint32_t y = __shfl_sync(FULL, incl, w + 7); fill(6); if (y <= R) w += 8; }
            { const uint32_t y = __shfl_sync(FULL, incl, w + 3); fill(7); if (y <= R) w += 4; }
            { const uint32_t y = __shfl_sync(FULL, incl, w + 1); fill(8); if (y <= R) w += 2; }
            { const uint32_t y = __shfl_sync(FULL, incl, w); fill(9); if (y <= R) w += 1; }
            w &= 31;
            const uint32_t before = __shfl_sync(FULL, incl - cnt, w);
            uint32_t wd0 = 0;
            if (WPL <= 2) wd0 = __shfl_sync(FULL, word[0], w);
            fill(10);
            uint32_t rr = R - before;                                // rank inside lane w's words
            uint32_t wd, g = w * WPL;
            if (WPL <= 2) {
                wd = wd0;
#pragma unroll
                for (int j = 1; j < WPL; ++j) {
                    const uint32_t nxt = __shfl_sync(FULL, word[j], w);
                    const uint32_t c = __popc(wd);
                    if (rr >= c) { rr -= c; wd = nxt; g = w * WPL + j; }
                }
            } else {
                // the words of lane w straight from shared memory (current: see the barrier below),
                // four at a time; first the group of four, then the word inside it
                const uint4* q4 = reinterpret_cast<const uint4*>(smask + w * WPL);
                uint4 q = q4[0];
                bool adv = true;
#pragma unroll
                for (int jj = 1; jj < WPL / 4; ++jj) {
                    const uint4 nq = q4[jj];
                    const uint32_t c4 = __popc(q.x) + __popc(q.y) + __popc(q.z) + __popc(q.w);
                    adv = adv && rr >= c4;
                    if (adv) { rr -= c4; q = nq; g += 4; }
                }
                wd = q.x;
                uint32_t c = __popc(q.x);
                adv = rr >= c;
                if (adv) { rr -= c; wd = q.y; ++g; }
                c = __popc(q.y);
                adv = adv && rr >= c;
                if (adv) { rr -= c; wd = q.z; ++g; }
                c = __popc(q.z);
                adv = adv && rr >= c;
                if (adv) { rr -= c; wd = q.w; ++g; }
            }
            const uint32_t pos = LIGHT ? nth_set_bit(wd, rr) : nth_set_bit_lut(wd, rr, lut);    // lut: nth_bit_lut_init, 2 KB per CTA
            P2V_CHECK_INDEX(g, 32 * WPL); P2V_CHECK_INDEX(pos, 32);
            if (LIGHT) { if ((int)(g * 32 + pos) < k) *park((int)(g * 32 + pos)) = (uint32_t)t | (xval << 14); }
            else { P2V_CHECK_INDEX(g * 32 + pos, KA); stime[g * 32 + pos] = (uint16_t)t; }
            if (WPL == 1 && P2V_DECODE_REDUX && K >= 512) {     // measured: +1.5 % at n = 512..1024, -3 % at n = 100
                // The low ranks of a batch (its roots) all land in the first word that still has free
                // slots, and same-word ATOMS serialise: those lanes are combined with one warp-wide
                // REDUX.OR and cleared by the word's owner; only the scattered rest uses atomics.
                const uint32_t wfirst = (uint32_t)__ffs(__ballot_sync(FULL, cnt != 0u)) - 1u;
                const bool infirst = g == wfirst;
                const uint32_t bits = __reduce_or_sync(FULL, infirst ? (1u << pos) : 0u);
                if (!infirst) atomicAnd(&smask[g], ~(1u << pos));
                if ((uint32_t)lane == wfirst) smask[lane] = word[0] & ~bits;
            } else {
                atomicAnd(&smask[g], ~(1u << pos));
            }
            fill(11);
            __syncwarp();
            if (WPL <= 2) {
#pragma unroll
                for (int j = 0; j < WPL; ++j) word[j] = smask[lane * WPL + j];
            } else {
#pragma unroll
                for (int jj = 0; jj < WPL / 4; ++jj) {
                    const uint4 q = reinterpret_cast<const uint4*>(smask + lane * WPL)[jj];
                    word[4 * jj] = q.x; word[4 * jj + 1] = q.y; word[4 * jj + 2] = q.z; word[4 * jj + 3] = q.w;
                }
            }
            fill(12);
    };
    int bp = nb - 1;
    uint32_t Rn_hi, Rn_lo;
    if (LIGHT) issue_v(bp);
    if (nb & 1) {                                  // odd number of batches: the top one goes alone
        if (LIGHT) {                               // (its "lower" batch bp - 1 is scanned again as the next pair's upper one)
            take_v(bp); issue_v(bp - 1);
            if (WPL >= 16) batch_rank_pair_vals15s(xv_hi, xv_lo, bp, k, lane, Rn_hi, Rn_lo);
            else batch_rank_pair_vals15(xv_hi, xv_lo, bp, k, lane, Rn_hi, Rn_lo);
        } else rank_pair(bp, Rn_hi, Rn_lo);
        place(bp, Rn_hi, xv_hi, fill_none);
        --bp;
    }
    if (bp >= 1) rank_pair(bp, Rn_hi, Rn_lo);
    for (; bp >= 3; bp -= 2) {                     // whole pairs (bp, bp - 1), the next pair's scan inside
        const uint32_t R1 = Rn_hi, R0 = Rn_lo;
        const uint32_t X1 = xv_hi, X0 = xv_lo;
#if P2V_DECODE_INTERLEAVE
        scan_begin(bp - 2);
        place(bp, R1, X1, fill_first);
        place(bp - 1, R0, X0, fill_second);
        scan_end(Rn_hi, Rn_lo);
#else
        rank_pair(bp - 2, Rn_hi, Rn_lo);
        place(bp, R1, X1, fill_none);
        place(bp - 1, R0, X0, fill_none);
#endif
    }
    if (bp >= 1) {                                 // last pair
        place(bp, Rn_hi, xv_hi, fill_none);
        place(bp - 1, Rn_lo, xv_lo, fill_none);
    }

    if (LIGHT) {
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) badmin = min(badmin, __shfl_xor_sync(FULL, badmin, o));
        if (badmin != 0x7FFFFFFF) return badmin;
    }

    // ---- D: forward sweep over the slots, 32 rows at a time -------------------------------
    // label of a row = v[t] of the nearest root (rank-0 insertion) at or left of it.
    // Ancestry / edges: the reference's parents[] pass (convert.rs:174-185) only ever looks
    // up the last earlier row with a given label; stbl keeps that row per label, same-batch
    // occurrences are resolved with MATCH.ANY.
    // Software pipeline: the loads of a batch are issued two iterations ahead (stime) and one
    // iteration ahead (sv), and its label resolution (ballot / shuffles / MATCH) runs while the
    // table lookup of the batch before it is in flight.
    uint32_t lab_carry = 0;
    uint32_t n_t, n_lab = 0, n_peers = 0;        // batch b + 1 after its label stage
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
        if (MODE != MODE_PAIRS) n_peers = __match_any_sync(FULL, act ? lab : (0x10000u + lane));
    };
    if (LIGHT) __syncwarp();                       // every lane's parked words are written
    // light form: the parked word of a slot; the L2 copy is the one that is sure to hold the other lanes' stores
    auto parked = [&](int s) -> uint32_t { return (s < k) ? __ldcg(park(s)) : 0u; };
    uint32_t t1, x1, t2, w2 = 0, w3 = 0, w4 = 0;   // parked words of the batches b + 2 .. b + 4 (they come from DRAM)
    if (LIGHT) {
        const uint32_t w0 = parked(lane);
        label_stage(0, w0 & 0x3FFFu, w0 >> 14);
        const uint32_t w1 = parked(32 + lane);
        t1 = w1 & 0x3FFFu; x1 = w1 >> 14;
        w2 = parked(64 + lane); w3 = parked(96 + lane); w4 = parked(128 + lane);
        t2 = 0;
    } else {
        t1 = (lane < k) ? (uint32_t)stime[lane] : 0u;              // batch 0
        x1 = sv[t1];
        label_stage(0, t1, x1);
        t1 = (32 + lane < k) ? (uint32_t)stime[32 + lane] : 0u;             // batch 1 (stime has K + 32 entries)
        x1 = sv[t1];
        t2 = (64 + lane < k) ? (uint32_t)stime[64 + lane] : 0u;    // batch 2
    }
    for (int b = 0; b < nb; ++b) {
        const int s = b * 32 + lane;
        const bool act = s < k;
        const uint32_t t = n_t, lab = n_lab, peers = n_peers;
        const bool root = n_root;
        // loads of the batches ahead
        const int s3 = s + 96;
        uint32_t x2, t3, w5 = 0;
        if (LIGHT) { t2 = w2 & 0x3FFFu; x2 = w2 >> 14; w5 = parked(s + 160); t3 = 0; }
        else { P2V_CHECK_INDEX(t2, KA); x2 = sv[t2]; t3 = (s3 < k) ? (uint32_t)stime[s3] : 0u; }
        if (MODE == MODE_PAIRS) {
            if (act) store_pair<OutT>(ot + 2 * (size_t)s, (OutT)lab, (OutT)(t + 1));
            label_stage(b + 1, t1, x1);
        } else {
            const uint32_t lower = peers & lanes_lt(lane);
            uint32_t cand = NONE16;
            if (root && !lower) { P2V_CHECK_INDEX(lab, KA); cand = stbl[lab]; }
            label_stage(b + 1, t1, x1);                 // independent of the table: overlaps the lookup
            uint32_t c1p = (uint32_t)(k + s);            // non-root: the previous row's parent
            if (root) {
                if (lower) cand = (uint32_t)(b * 32 + 31 - __clz(lower));
                c1p = (cand != NONE16) ? (uint32_t)(k + 1) + cand : lab;
            }
            __syncwarp();
            if (act && (peers & lanes_gt(lane)) == 0) { P2V_CHECK_INDEX(lab, KA); stbl[lab] = (uint16_t)s; }
            __syncwarp();
            if (LIGHT && MODE == MODE_ANCESTRY) {
                if (act) P2V_CHECK_INDEX(t + 1, KA);
                const uint32_t q = act ? (uint32_t)stbl[t + 1] : NONE16;
                const uint32_t c2p = (q != NONE16) ? (uint32_t)(k + 1) + q : t + 1;
                // element parity of the batch's first element against a two-element (16 / 8 byte) boundary
                OutT* const dst = ot + 96 * (size_t)b;
                const int a = (int)((reinterpret_cast<uintptr_t>(dst) / sizeof(OutT)) & 1);
                OutT* const st = srow + a;                 // the staged copy has the alignment of the destination
                P2V_CHECK_INDEX(a + 3 * lane + 2, 98);
                st[3 * lane] = (OutT)c1p; st[3 * lane + 1] = (OutT)c2p; st[3 * lane + 2] = (OutT)(k + 1 + s);
                __syncwarp();
                const int lim = 3 * min(32, k - b * 32);   // valid elements of this batch
                // vectors (f, f + 1) with f + a even: f = 2 q - a
#pragma unroll
                for (int rep = 0; rep < 2; ++rep) {
                    const int qv = lane + 32 * rep;
                    const int f = 2 * qv - a;
                    if (qv <= 48) {
                        if (f >= 0 && f + 1 < lim) store_pair<OutT>(dst + f, st[f], st[f + 1]);
                        else if (f + 1 >= 0 && f + 1 < lim && f < 0) dst[f + 1] = st[f + 1];     // the odd first element
                        else if (f >= 0 && f < lim) dst[f] = st[f];                                // the odd last element
                    }
                }
                __syncwarp();
            } else if (act) {
                P2V_CHECK_INDEX(t + 1, KA);
                const uint32_t q = stbl[t + 1];
                const uint32_t c2p = (q != NONE16) ? (uint32_t)(k + 1) + q : t + 1;
                store_row<OutT, MODE>(ot, s, (OutT)c1p, (OutT)c2p, (OutT)(k + 1 + s));
            }
        }
        t1 = t2; x1 = x2; t2 = t3; w2 = w3; w3 = w4; w4 = w5;
    }
    __syncwarp();
    return -1;
}

}  // namespace p2v
