// Batched encode pairs / ancestry / edges -> v for sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   vector::convert::from_pairs      phylo2vec/src/vector/convert.rs:23-30
//   vector::convert::from_ancestry   :129-136   + build_cherries :399-440
//   vector::convert::from_edges      :201-213   + order_cherries :449-477
//   vector::convert::build_vector    :320-337   + Fenwick :286-316
//
// Three kernels: encode_light_kernel (p2v_encode_light.cuh, round 2: one pass over the rows, 2 bytes of shared
// memory per leaf -- from_pairs / from_ancestry at every size up to k = 262143, from_edges from k = 800),
// encode_small_kernel below (round 1; today only from_edges with k < 800, which must not need a workspace) and
// encode_large_kernel (CTA per tree, k > 262143).
//
// Formulation of the two round-1 kernels (validated on the CPU by tests/parallel_model.py):
//  * build_vector: for the cherry whose c_max is c, idx = number of EARLIER rows
//    with a smaller c_max.  Walking the values c upwards in batches of 32, idx is
//    a prefix-popcount on the "rows already seen" bit mask (state at batch
//    start) plus an in-batch count of lower lanes with a smaller row.
//  * build_cherries (min-descendant relabel) keeps the reference's row-order
//    semantics but runs as a wavefront: 32 rows at a time, a row fires as soon
//    as the rows defining its internal children have fired.
#include <cuda_fp16.h>

#include <cstdlib>

#include "p2v_common.cuh"
#include "p2v_internal.h"
#include "p2v_encode_light.cuh"

namespace p2v {

constexpr int ENC_SMALL_WARPS = 4;
constexpr int ENC_WARP_K_MAX = 4096;      // round-1 warp kernel with staged rows: above k = 800 only reachable with
                                          // P2V_ENC_LIGHT_FROM set higher (tuning / comparison runs)
constexpr int ENC_LARGE_THREADS = 256;
constexpr int ENC_LARGE_WARPS = ENC_LARGE_THREADS / 32;

// status codes written per tree
constexpr int32_t TREE_OK = 0;

// ---- generic per-tree body, parameterised on the storage type of the scratch
//      arrays (uint16 in shared memory for k<=1024, uint32 in global above) ----
template <typename IdxT, int MODE>
__device__ __forceinline__ bool load_row(const IdxT* __restrict__ in, int s, int k, uint32_t& c1,
                                         uint32_t& c2, uint32_t& p) {
    long long a, b, c;
    if (MODE == FROM_PAIRS) {
        a = (long long)__ldg(in + 2 * (size_t)s);
        b = (long long)__ldg(in + 2 * (size_t)s + 1);
        c = 0;
        if (a < 0 || a > k || b < 0 || b > k) return false;
    } else if (MODE == FROM_ANCESTRY) {
        a = (long long)__ldg(in + 3 * (size_t)s);
        b = (long long)__ldg(in + 3 * (size_t)s + 1);
        c = (long long)__ldg(in + 3 * (size_t)s + 2);
        const long long top = 2LL * k + 1;
        if (a < 0 || a > top || b < 0 || b > top || c < 0 || c > top) return false;
    } else {  // FROM_EDGES: edges[2s] = (c1, p), edges[2s+1] = (c2, .)   convert.rs:206-208
        a = (long long)__ldg(in + 4 * (size_t)s);
        c = (long long)__ldg(in + 4 * (size_t)s + 1);
        b = (long long)__ldg(in + 4 * (size_t)s + 2);
        const long long top = 2LL * k + 1;
        if (a < 0 || a > top || b < 0 || b > top || c < 0 || c > top) return false;
    }
    c1 = (uint32_t)a; c2 = (uint32_t)b; p = (uint32_t)c;
    return true;
}

// --------------------------------------------------------------------------
// k <= 1024: one warp per tree, shared memory.
// --------------------------------------------------------------------------

// For two batches of 32 values at once (halves of a half2; rows are integers < 1024, exact in
// fp16): number of LOWER lanes of the same batch whose row is smaller.  SHFL.UP below lane 0
// hands back the lane's own value, which never compares smaller.
__device__ __forceinline__ void count_smaller_pair(uint32_t r_a, bool act_a, uint32_t r_b, bool act_b,
                                                   uint32_t& cnt_a, uint32_t& cnt_b) {
    const __half inf = __ushort_as_half((unsigned short)0x7C00);
    const __half2 mine = __halves2half2(act_a ? __uint2half_rn(r_a) : inf, act_b ? __uint2half_rn(r_b) : inf);
    __half2 cnt = __float2half2_rn(0.0f);
#pragma unroll
    for (int d = 1; d < 32; ++d) {
        const __half2 y = __shfl_up_sync(FULL, mine, d);
        cnt = __hadd2(cnt, __hlt2(y, mine));
    }
    cnt_a = (uint32_t)__half2int_rz(__low2half(cnt));
    cnt_b = (uint32_t)__half2int_rz(__high2half(cnt));
}

// The same count with 32-bit integers for rows >= 2048 (mid-size trees).
__device__ __forceinline__ void count_smaller_pair_int(uint32_t r_a, bool act_a, uint32_t r_b, bool act_b,
                                                       uint32_t& cnt_a, uint32_t& cnt_b) {
    const uint32_t ma = act_a ? r_a : 0x7FFFFFFFu, mb = act_b ? r_b : 0x7FFFFFFFu;
    cnt_a = 0; cnt_b = 0;
#pragma unroll
    for (int d = 1; d < 32; ++d) {
        count_smaller_step(cnt_a, ma, d);
        count_smaller_step(cnt_b, mb, d);
    }
}

// WPL = "rows seen" mask words per lane: 1 / 2 for k <= 1024 / 2047 (four warps per CTA, half2
// counts), 4 for batches of mid-size trees (k <= 4096; one warp per CTA, integer counts).
template <typename IdxT, int MODE, int WPL>
__global__ void __launch_bounds__(ENC_SMALL_WARPS * 32)
encode_small_kernel(const IdxT* __restrict__ in, int64_t B, int k, IdxT* __restrict__ vout,
                    int32_t* __restrict__ status, int K) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int KA = K + 32;
    const bool has_par = (MODE != FROM_PAIRS);
    const size_t per_warp = (size_t)(has_par ? 5 : 3) * KA + 64 * WPL;
    uint16_t* base = reinterpret_cast<uint16_t*>(smem_raw) + warp * per_warp;
    uint16_t* sc1 = base;
    uint16_t* sc2 = sc1 + KA;
    uint16_t* spr = sc2 + KA;       // parent label by row; later row_of[value]
    uint16_t* sdef = spr + KA;      // defining row of internal label k+1+i      [has_par]
    uint16_t* smd = sdef + KA;      // min descendant of internal label k+1+i    [has_par]
    uint32_t* smask = reinterpret_cast<uint32_t*>(base + per_warp - 64 * WPL);   // 32*WPL words
    const int nb = K >> 5;
    const int in_cols = (MODE == FROM_PAIRS) ? 2 : (MODE == FROM_ANCESTRY ? 3 : 4);
    const uint32_t kint = (uint32_t)k + 1;      // first internal label

    const int nwarps = blockDim.x >> 5;
    for (int64_t tree = (int64_t)blockIdx.x * nwarps + warp; tree < B;
         tree += (int64_t)gridDim.x * nwarps) {
        const IdxT* it = in + tree * ((int64_t)k * in_cols);
        bool bad = false;
        // ---- E1: load rows (4 batches of loads in flight) -------------------------------------
        if (has_par) {
            for (int i = lane; i < KA; i += 32) { sdef[i] = (uint16_t)NONE16; smd[i] = (uint16_t)NONE16; }
        }
#pragma unroll
        for (int j = 0; j < WPL; ++j) smask[lane * WPL + j] = 0;
        __syncwarp();
        long long offset = 0;
        if (MODE == FROM_EDGES) {
            // order_cherries (convert.rs:449-477): row -> index parent - k - offset,
            // offset = max parent - 2k + 1
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
                    if (pv[j] < 0 || pv[j] > 2LL * k + 1) bad = true;
                    else mx = max(mx, (uint32_t)pv[j]);
                }
            }
#pragma unroll
            for (int d = 16; d; d >>= 1) mx = max(mx, __shfl_xor_sync(FULL, mx, d));
            offset = (long long)mx - 2LL * k + 1;
        }
        for (int b0 = 0; b0 < nb; b0 += 4) {
            uint32_t c1[4], c2[4], pp[4];
            bool okr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int s = (b0 + j) * 32 + lane;
                c1[j] = c2[j] = pp[j] = 0; okr[j] = true;
                if (s < k) okr[j] = load_row<IdxT, MODE>(it, s, k, c1[j], c2[j], pp[j]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int s = (b0 + j) * 32 + lane;
                if (s < k) {
                    if (!okr[j]) { bad = true; continue; }
                    int dst = s;
                    if (MODE == FROM_EDGES) {
                        const long long idx = (long long)pp[j] - k - offset;
                        if (idx < 0 || idx >= k) { bad = true; continue; }
                        const uint32_t old = atomicOr(&smask[idx >> 5], 1u << (idx & 31));
                        if (old & (1u << (idx & 31))) bad = true;      // two rows, one parent
                        dst = (int)idx;
                    }
                    P2V_CHECK_INDEX(dst, KA);
                    sc1[dst] = (uint16_t)c1[j]; sc2[dst] = (uint16_t)c2[j]; spr[dst] = (uint16_t)pp[j];
                }
            }
        }
        bad = __any_sync(FULL, bad);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < WPL; ++j) smask[lane * WPL + j] = 0;
        // ---- E2: min-descendant relabel (build_cherries) as a wavefront ---------------
        if (has_par && !bad) {
            for (int b = 0; b < nb; ++b) {
                const int s = b * 32 + lane;
                if (s < k) {
                    const uint32_t p = spr[s];
                    if (p < kint) bad = true;                 // a leaf label as parent: not a tree in this format
                    else { P2V_CHECK_INDEX(p - kint, KA); sdef[p - kint] = (uint16_t)s; }
                }
            }
            bad = __any_sync(FULL, bad);
            __syncwarp();
        }
        if (has_par && !bad) {
            for (int b = 0; b < nb; ++b) {
                const int base_row = b * 32;
                const int s = base_row + lane;
                const bool act = s < k;
                uint32_t c1 = 0, c2 = 0, p = kint, d1 = NONE16, d2 = NONE16;
                if (act) {
                    c1 = sc1[s]; c2 = sc2[s]; p = spr[s];
                    if (c1 >= kint) { P2V_CHECK_INDEX(c1 - kint, KA); d1 = sdef[c1 - kint]; }
                    if (c2 >= kint) { P2V_CHECK_INDEX(c2 - kint, KA); d2 = sdef[c2 - kint]; }
                    if (d1 != NONE16 && d1 >= (uint32_t)s) d1 = NONE16;  // defined later: a leaf, as in the reference
                    if (d2 != NONE16 && d2 >= (uint32_t)s) d2 = NONE16;
                }
                uint32_t done = __ballot_sync(FULL, !act);
                // A row fires when the rows defining its internal children have fired.  Rows that only
                // wait for the row just before them (the usual case: a non-root row continues the
                // previous row's subtree) form chains; a whole chain hanging off a fired or ready row
                // is resolved in one step with a segmented min-scan, so the number of rounds is the
                // depth of the OTHER in-batch dependencies, not the chain length.
                while (done != FULL) {
                    const bool mine_pending = !((done >> lane) & 1u);
                    const bool w1 = d1 != NONE16 && d1 >= (uint32_t)base_row && !((done >> (d1 - base_row)) & 1u);
                    const bool w2 = d2 != NONE16 && d2 >= (uint32_t)base_row && !((done >> (d2 - base_row)) & 1u);
                    bool prev1 = w1 && d1 + 1 == (uint32_t)s, prev2 = w2 && d2 + 1 == (uint32_t)s;
                    const bool other = (w1 && !prev1) || (w2 && !prev2) || (prev1 && prev2);
                    const bool avail = mine_pending && !other;
                    const uint32_t A = __ballot_sync(FULL, avail);
                    const uint32_t Lk = __ballot_sync(FULL, avail && (prev1 || prev2));
                    const uint32_t H = A & ~Lk;                      // ready on their own
                    const uint32_t T = (H << 1) & Lk;                // chain links right above a ready row
                    const uint32_t Rm = H | (((Lk + T) ^ Lk) & Lk);  // ... and the whole runs of links above them
                    const bool mine = (Rm >> lane) & 1u;
                    uint32_t o1 = 0xFFFFu, o2 = 0xFFFFu;
                    if (mine) {
                        if (!prev1) o1 = (d1 != NONE16) ? (uint32_t)smd[c1 - kint] : c1;
                        if (!prev2) o2 = (d2 != NONE16) ? (uint32_t)smd[c2 - kint] : c2;
                    }
                    uint32_t val = min(o1, o2);
                    const uint32_t hb = H & lanes_le(lane);
                    const int dist = hb ? lane - (31 - __clz(hb)) : 0;  // distance to the head of my run
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t y = __shfl_up_sync(FULL, val, d);
                        if (mine && dist >= d) val = min(val, y);
                    }
                    const uint32_t pm = __shfl_up_sync(FULL, val, 1);   // min descendant of the row before
                    if (mine) {
                        const uint32_t m1 = prev1 ? pm : o1, m2 = prev2 ? pm : o2;
                        smd[p - kint] = (uint16_t)val;
                        sc1[s] = (uint16_t)m1; sc2[s] = (uint16_t)m2;
                    }
                    __syncwarp();
                    done |= Rm;
                }
            }
            __syncwarp();
        }
        // ---- E3: row_of[c_max], c_max must be a permutation of 1..k ---------------------
        __syncwarp();
        if (!bad) {
            for (int b = 0; b < nb; ++b) {
                const int s = b * 32 + lane;
                if (s < k) {
                    const uint32_t cm = max((uint32_t)sc1[s], (uint32_t)sc2[s]);
                    if (cm < 1 || cm > (uint32_t)k) bad = true;
                    else {
                        const uint32_t bit = 1u << ((cm - 1) & 31);
                        if (atomicOr(&smask[(cm - 1) >> 5], bit) & bit) bad = true;
                        P2V_CHECK_INDEX(cm, KA);
                        spr[cm] = (uint16_t)s;
                    }
                }
            }
            bad = __any_sync(FULL, bad);
        }
        __syncwarp();
        if (bad) {
            if (lane == 0 && status) status[tree] = TREE_MALFORMED;
            continue;
        }
        if (lane == 0 && status) status[tree] = TREE_OK;
#pragma unroll
        for (int j = 0; j < WPL; ++j) smask[lane * WPL + j] = 0;
        __syncwarp();
        // ---- E4: build_vector, values ascending, two batches per in-batch count ----------------
        IdxT* vo = vout + tree * k;
        uint32_t word[WPL];  // rows already seen (rows of smaller values): words lane*WPL .. +WPL-1
#pragma unroll
        for (int j = 0; j < WPL; ++j) word[j] = 0;
        for (int bp = 0; bp < nb; bp += 2) {
            const int ca = bp * 32 + lane + 1, cb = ca + 32;
            const bool act_a = ca <= k, act_b = cb <= k;
            const uint32_t ra = act_a ? (uint32_t)spr[ca] : 0u;
            const uint32_t rb = act_b ? (uint32_t)spr[cb] : 0u;
            uint32_t inb_a, inb_b;
            if (WPL <= 2) count_smaller_pair(ra, act_a, rb, act_b, inb_a, inb_b);
            else count_smaller_pair_int(ra, act_a, rb, act_b, inb_a, inb_b);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = h ? cb : ca;
                const bool act = h ? act_b : act_a;
                const uint32_t r = h ? rb : ra;
                uint32_t cnt = 0;
#pragma unroll
                for (int j = 0; j < WPL; ++j) cnt += __popc(word[j]);
                const uint32_t excl = warp_incl_scan(cnt, lane) - cnt;
                const uint32_t g = r >> 5;                       // word that holds row r
                const uint32_t src = g / WPL;
                uint32_t idx = __shfl_sync(FULL, excl, src) + (h ? inb_b : inb_a);
                const uint32_t low = (1u << (r & 31)) - 1u;
                if (WPL <= 2) {
#pragma unroll
                    for (int j = 0; j < WPL; ++j) {
                        const uint32_t wj = __shfl_sync(FULL, word[j], src);
                        const uint32_t jj = g - src * WPL;
                        if ((uint32_t)j < jj) idx += __popc(wj);
                        else if ((uint32_t)j == jj) idx += __popc(wj & low);
                    }
                } else {      // the words of lane `src` straight from shared memory (current: barrier below)
                    const uint32_t jj = g - src * WPL;
#pragma unroll
                    for (int j4 = 0; j4 < WPL / 4; ++j4) {
                        const uint4 q = reinterpret_cast<const uint4*>(smask + src * WPL)[j4];
                        const uint32_t qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint32_t jx = (uint32_t)(4 * j4 + j);
                            if (jx < jj) idx += __popc(qq[j]);
                            else if (jx == jj) idx += __popc(qq[j] & low);
                        }
                    }
                }
                if (act) {
                    P2V_CHECK_INDEX(r, KA); P2V_CHECK_INDEX(g, 32 * WPL);
                    const uint32_t lo = min((uint32_t)sc1[r], (uint32_t)sc2[r]);
                    vo[c - 1] = (IdxT)(idx == 0 ? lo : (uint32_t)(c - 1) + idx);
                    atomicOr(&smask[g], 1u << (r & 31));
                }
                __syncwarp();
#pragma unroll
                for (int j = 0; j < WPL; ++j) word[j] = smask[lane * WPL + j];
            }
        }
        __syncwarp();
    }
}

// --------------------------------------------------------------------------
// k > 1024: one CTA per tree, scratch in a global workspace, bit mask + Fenwick
// tree in shared memory.
// --------------------------------------------------------------------------
template <typename IdxT, int MODE>
__global__ void __launch_bounds__(ENC_LARGE_THREADS)
encode_large_kernel(const IdxT* __restrict__ in, int64_t B, int k, IdxT* __restrict__ vout,
                    int32_t* __restrict__ status, uint32_t* __restrict__ ws, size_t ws_stride) {
    extern __shared__ __align__(16) uint32_t sm32[];
    __shared__ int s_bad;
    __shared__ uint32_t s_max;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = (k + 31) >> 5;
    const int KA = ((k + 1 + 31) & ~31) + 32;
    const int KL = ((2 * k + 2 + 31) & ~31) + 32;
    uint32_t* bits = sm32;
    uint32_t* fen = bits + nw;      // fen[1..nw]
    uint32_t* wc1 = ws + (size_t)blockIdx.x * ws_stride;
    uint32_t* wc2 = wc1 + KA;
    uint32_t* wpr = wc2 + KA;       // parent by row; later row_of[value]
    uint32_t* wcnt = wpr + KA;      // in-batch counts (by value-1)
    uint32_t* wdef = wcnt + KA;
    uint32_t* wmd = wdef + KL;
    const bool has_par = (MODE != FROM_PAIRS);
    const int in_cols = (MODE == FROM_PAIRS) ? 2 : (MODE == FROM_ANCESTRY ? 3 : 4);

    for (int64_t tree = blockIdx.x; tree < B; tree += gridDim.x) {
        const IdxT* it = in + tree * ((int64_t)k * in_cols);
        if (tid == 0) { s_bad = 0; s_max = 0; }
        for (int w = tid; w < nw; w += ENC_LARGE_THREADS) bits[w] = 0;
        for (int i = tid; i <= nw; i += ENC_LARGE_THREADS) fen[i] = 0;
        if (has_par)
            for (int i = tid; i < KL; i += ENC_LARGE_THREADS) { wdef[i] = NONE32; wmd[i] = NONE32; }
        __syncthreads();
        // ---- E1 ---------------------------------------------------------------------
        if (MODE == FROM_EDGES) {
            uint32_t mx = 0;
            for (int s = tid; s < k; s += ENC_LARGE_THREADS) {
                uint32_t c1 = 0, c2 = 0, p = 0;
                if (!load_row<IdxT, MODE>(it, s, k, c1, c2, p)) s_bad = 1;
                mx = max(mx, p);
            }
            atomicMax(&s_max, mx);
            __syncthreads();
            const long long offset = (long long)s_max - 2LL * k + 1;
            if (!s_bad) {
                for (int s = tid; s < k; s += ENC_LARGE_THREADS) {
                    uint32_t c1, c2, p;
                    load_row<IdxT, MODE>(it, s, k, c1, c2, p);
                    const long long idx = (long long)p - k - offset;
                    if (idx < 0 || idx >= k) s_bad = 1;
                    else {
                        const uint32_t bit = 1u << (idx & 31);
                        if (atomicOr(&bits[idx >> 5], bit) & bit) s_bad = 1;
                        wc1[idx] = c1; wc2[idx] = c2; wpr[idx] = p;
                    }
                }
            }
            __syncthreads();
            for (int w = tid; w < nw; w += ENC_LARGE_THREADS) bits[w] = 0;
        } else {
            for (int s = tid; s < k; s += ENC_LARGE_THREADS) {
                uint32_t c1 = 0, c2 = 0, p = 0;
                if (!load_row<IdxT, MODE>(it, s, k, c1, c2, p)) s_bad = 1;
                wc1[s] = c1; wc2[s] = c2; wpr[s] = p;
            }
        }
        __syncthreads();
        // ---- E2: wavefront over 32-row batches, one warp (rows depend on earlier rows) ----
        if (has_par && !s_bad) {
            for (int s = tid; s < k; s += ENC_LARGE_THREADS) wdef[wpr[s]] = (uint32_t)s;
            __syncthreads();
            if (warp == 0) {
                for (int b = 0; b < nw; ++b) {
                    const int base_row = b * 32;
                    const int s = base_row + lane;
                    const bool act = s < k;
                    uint32_t c1 = 0, c2 = 0, p = 0, d1 = NONE32, d2 = NONE32;
                    if (act) {
                        c1 = wc1[s]; c2 = wc2[s]; p = wpr[s];
                        d1 = wdef[c1]; d2 = wdef[c2];
                        if (d1 != NONE32 && d1 >= (uint32_t)s) d1 = NONE32;
                        if (d2 != NONE32 && d2 >= (uint32_t)s) d2 = NONE32;
                    }
                    uint32_t done = __ballot_sync(FULL, !act);
                    while (done != FULL) {
                        bool ready = false;
                        if (!((done >> lane) & 1u)) {
                            const bool w1 = d1 != NONE32 && d1 >= (uint32_t)base_row && !((done >> (d1 - base_row)) & 1u);
                            const bool w2 = d2 != NONE32 && d2 >= (uint32_t)base_row && !((done >> (d2 - base_row)) & 1u);
                            ready = !w1 && !w2;
                            if (ready) {
                                const uint32_t m1 = (d1 != NONE32) ? wmd[c1] : c1;
                                const uint32_t m2 = (d2 != NONE32) ? wmd[c2] : c2;
                                wmd[p] = min(m1, m2);
                                wc1[s] = m1; wc2[s] = m2;
                            }
                        }
                        __syncwarp();
                        done |= __ballot_sync(FULL, ready);
                    }
                }
            }
            __syncthreads();
        }
        // ---- E3 ---------------------------------------------------------------------
        if (!s_bad) {
            for (int s = tid; s < k; s += ENC_LARGE_THREADS) {
                const uint32_t cm = max(wc1[s], wc2[s]);
                if (cm < 1 || cm > (uint32_t)k) s_bad = 1;
                else {
                    const uint32_t bit = 1u << ((cm - 1) & 31);
                    if (atomicOr(&bits[(cm - 1) >> 5], bit) & bit) s_bad = 1;
                    wpr[cm] = (uint32_t)s;
                }
            }
        }
        __syncthreads();
        if (s_bad) {
            if (tid == 0 && status) status[tree] = TREE_MALFORMED;
            __syncthreads();
            continue;
        }
        if (tid == 0 && status) status[tree] = TREE_OK;
        for (int w = tid; w < nw; w += ENC_LARGE_THREADS) bits[w] = 0;
        // ---- E4a: in-batch counts, all warps --------------------------------------------
        for (int b = warp; b < nw; b += ENC_LARGE_WARPS) {
            const int c = b * 32 + lane + 1;
            const bool act = c <= k;
            const uint32_t mine = act ? wpr[c] : 0xFFFFFFFFu;
            uint32_t cnt = 0;
#pragma unroll
            for (int d = 1; d < 32; ++d) count_smaller_step(cnt, mine, d);
            if (act) wcnt[c - 1] = cnt;
        }
        __syncthreads();
        // ---- E4b: prefix counts on the seen-rows mask, serial over value batches ------------
        if (warp == 0) {
            IdxT* vo = vout + tree * k;
            for (int b = 0; b < nw; ++b) {
                const int c = b * 32 + lane + 1;
                const bool act = c <= k;
                uint32_t r = 0;
                if (act) {
                    r = wpr[c];
                    const int w = (int)(r >> 5);
                    uint32_t idx = wcnt[c - 1] + __popc(bits[w] & ((1u << (r & 31)) - 1u));
                    for (int i = w; i > 0; i -= i & -i) idx += fen[i];
                    const uint32_t lo = min(wc1[r], wc2[r]);
                    vo[c - 1] = (IdxT)(idx == 0 ? lo : (uint32_t)(c - 1) + idx);
                }
                __syncwarp();
                if (act) {
                    const int w = (int)(r >> 5);
                    atomicOr(&bits[w], 1u << (r & 31));
                    for (int i = w + 1; i <= nw; i += i & -i) atomicAdd(&fen[i], 1u);
                }
                __syncwarp();
            }
        }
        __syncthreads();
    }
}

// --------------------------------------------------------------------------
static size_t enc_large_stride(int64_t k) {
    const size_t KA = (((size_t)k + 1 + 31) & ~(size_t)31) + 32;
    const size_t KL = ((2 * (size_t)k + 2 + 31) & ~(size_t)31) + 32;
    return 4 * KA + 2 * KL;
}
static int enc_large_grid(int64_t B) {
    const int64_t cap = (int64_t)sm_count() * 4;
    return (int)(B < cap ? B : cap);
}

// ---- the light warp kernel (p2v_encode_light.cuh): shape of a launch, shared by the workspace query ----
struct LightPlan { int warps, per_sm, K; bool gt; size_t smem, ws_words; };
static int enc_light_from() {               // smallest k that takes the light kernel: measured +7 % at k = 999, -3 % at 511
                                            // against encode_small_kernel (P2V_ENC_LIGHT_FROM to experiment)
    static const int v = [] { const char* e = getenv("P2V_ENC_LIGHT_FROM"); return e ? atoi(e) : 800; }();
    return v;
}
static int enc_env(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
static int enc_light_tsmem() { static const int v = enc_env("P2V_ENC_TSMEM", ENC_LIGHT_T_SMEM_K_MAX); return v; }
static int enc_light_gt_per_sm() { static const int v = enc_env("P2V_ENC_GT_PERSM", 16); return v; }
static bool enc_light(int64_t k) { return k >= enc_light_from() && k <= ENC_LIGHT_WIDE_K_MAX; }
static LightPlan enc_light_plan(int64_t k) {
    LightPlan p;
    p.K = (int)((k + 31) & ~31);
    p.gt = k > enc_light_tsmem();
    const size_t per_warp = enc_light_smem_words(p.K, !p.gt) * sizeof(uint32_t);
    p.warps = per_warp <= 6144 ? 4 : (per_warp <= 12288 ? 2 : 1);
    p.smem = per_warp * p.warps;
    int per_sm = (int)((size_t)(227 * 1024) / (p.smem + 1024));
    if (per_sm > 32) per_sm = 32;
    if (per_sm * p.warps > 48) per_sm = 48 / p.warps;
    if (p.gt && per_sm > enc_light_gt_per_sm()) per_sm = enc_light_gt_per_sm();       // tables in L2: more trees in flight only evict one another
    if (per_sm < 1) per_sm = 1;
    p.per_sm = per_sm;
    p.ws_words = enc_light_ws_words(p.K, !p.gt);
    return p;
}

size_t encode_workspace_bytes(int64_t B, int64_t k) {
    if (B <= 0) return 0;
    if (enc_light(k)) {
        const LightPlan p = enc_light_plan(k);
        const int64_t want = (B + p.warps - 1) / p.warps, cap = (int64_t)sm_count() * p.per_sm;
        return (size_t)(want < cap ? want : cap) * p.warps * p.ws_words * sizeof(uint32_t);
    }
    if (k <= ENC_SMALL_K_MAX) return 0;
    return (size_t)enc_large_grid(B) * enc_large_stride(k) * sizeof(uint32_t);
}

template <typename IdxT, int MODE, bool GT, bool WIDE = false>
static cudaError_t launch_encode_light(const void* in, int64_t B, int64_t k, void* v, int32_t* status, void* ws,
                                       size_t ws_bytes, cudaStream_t stream) {
    const LightPlan p = enc_light_plan(k);
    auto kern = encode_light_kernel<IdxT, MODE, GT, WIDE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return e;
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, p.warps * 32, p.smem);
    if (e != cudaSuccess) return e;
    const int per_sm = occ < 1 ? 1 : (occ < p.per_sm ? occ : p.per_sm);
    const int64_t want = (B + p.warps - 1) / p.warps, cap = (int64_t)sm_count() * per_sm;
    const int grid = (int)(want < cap ? want : cap);
    // the workspace holds from_edges' (c1, c2) table and the label table when that is not in shared memory
    const bool needs_ws = (MODE == FROM_EDGES) || GT;
    if (needs_ws && (!ws || ws_bytes < (size_t)grid * p.warps * p.ws_words * sizeof(uint32_t))) return cudaErrorInvalidValue;
    kern<<<grid, p.warps * 32, p.smem, stream>>>((const IdxT*)in, B, (int)k, (IdxT*)v, status, p.K, (uint32_t*)ws,
                                                 p.ws_words);
    count_launch();
    return cudaGetLastError();
}

template <typename IdxT, int MODE>
static cudaError_t launch_encode_t(const void* in, int64_t B, int64_t k, void* v, int32_t* status,
                                   void* ws, size_t ws_bytes, cudaStream_t stream) {
    // below enc_light_from() the light kernel is still the faster one (+2 ... 14 %), but from_edges would need a
    // workspace there, which the interface promises not to ask for: that mode stays on the round-1 kernel
    if (enc_light(k) || (MODE != FROM_EDGES && k < enc_light_from())) {
        if (k > ENC_LIGHT_K_MAX) return launch_encode_light<IdxT, MODE, true, true>(in, B, k, v, status, ws, ws_bytes, stream);
        return k > enc_light_tsmem() ? launch_encode_light<IdxT, MODE, true>(in, B, k, v, status, ws, ws_bytes, stream)
                                     : launch_encode_light<IdxT, MODE, false>(in, B, k, v, status, ws, ws_bytes, stream);
    }
    // batches of mid-size trees stay on the warp-per-tree kernel (a fraction of the CTA kernel's work)
    if (k <= ENC_SMALL_K_MAX || (k <= ENC_WARP_K_MAX && B * 4 > sm_count())) {
        const int K = (int)((k + 31) & ~31);
        const int wpl = (k <= 1024) ? 1 : (k <= ENC_SMALL_K_MAX ? 2 : 4);
        const int warps = (wpl <= 2) ? ENC_SMALL_WARPS : 1;
        const size_t per_warp = (size_t)(MODE != FROM_PAIRS ? 5 : 3) * (K + 32) + 64 * wpl;
        const size_t smem = warps * per_warp * sizeof(uint16_t);
        auto kern = encode_small_kernel<IdxT, MODE, 1>;
        if (wpl == 2) kern = encode_small_kernel<IdxT, MODE, 2>;
        else if (wpl == 4) kern = encode_small_kernel<IdxT, MODE, 4>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        int64_t want = (B + warps - 1) / warps;
        int64_t cap = (int64_t)sm_count() * per_sm;
        const int grid = (int)(want < cap ? want : cap);
        kern<<<grid, warps * 32, smem, stream>>>((const IdxT*)in, B, (int)k, (IdxT*)v, status, K);
    } else {
        const size_t stride = enc_large_stride(k);
        const int grid = enc_large_grid(B);
        if (!ws || ws_bytes < (size_t)grid * stride * sizeof(uint32_t)) return cudaErrorInvalidValue;
        const int nw = (int)((k + 31) >> 5);
        const size_t smem = ((size_t)2 * nw + 8) * sizeof(uint32_t);
        auto kern = encode_large_kernel<IdxT, MODE>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, ENC_LARGE_THREADS, smem, stream>>>((const IdxT*)in, B, (int)k, (IdxT*)v, status,
                                                       (uint32_t*)ws, stride);
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_encode(int mode, const void* in, int idx_bytes, int64_t B, int64_t k, void* v,
                          int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (B <= 0 || k <= 0) return cudaSuccess;
#define P2V_DISPATCH(T)                                                                             \
    switch (mode) {                                                                                 \
    case FROM_PAIRS: return launch_encode_t<T, FROM_PAIRS>(in, B, k, v, status, ws, ws_bytes, stream);       \
    case FROM_ANCESTRY: return launch_encode_t<T, FROM_ANCESTRY>(in, B, k, v, status, ws, ws_bytes, stream); \
    case FROM_EDGES: return launch_encode_t<T, FROM_EDGES>(in, B, k, v, status, ws, ws_bytes, stream);       \
    default: return cudaErrorInvalidValue;                                                          \
    }
    if (idx_bytes == 8) { P2V_DISPATCH(long long) }
    if (idx_bytes == 4) { P2V_DISPATCH(int) }
#undef P2V_DISPATCH
    return cudaErrorInvalidValue;
}

}  // namespace p2v
