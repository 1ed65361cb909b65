// Batched decode v -> pairs / ancestry / edges for sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   vector::avl::AVLTree::with_vector + inorder_traversal   phylo2vec/src/vector/avl.rs:35-52,193-211
//   vector::convert::to_pairs / to_ancestry / to_edges      phylo2vec/src/vector/convert.rs:103-109,167-188,227-248
//   vector::base::check_v                                    phylo2vec/src/vector/base.rs:50-67
//
// Formulation (validated on the CPU by tests/parallel_model.py):
//  * Element t (pair (.,t+1)) is inserted at rank ins_t = 0 if v[t] <= t else
//    v[t]-t.  Walking time backwards, element t takes the ins_t-th slot that
//    later elements left free.  Batches of 32 consecutive times are resolved
//    together: a 31-step SHFL.DOWN scan gives each lane its rank R among the
//    elements of [0, batch end), then all 32 lanes rank-select against the same
//    free-slot bit mask (one 32-bit word per lane + popc prefix for k <= 1024,
//    a shared-memory Fenwick tree over word counts above that).
//  * Rank-0 insertions ("roots") carry their own label v[t]; every other element
//    inherits the label of the nearest root on its left, so pair labels are a
//    segmented broadcast over slots.
//  * The sequential parents[] relabel pass of to_ancestry/to_edges is replaced by
//    closed forms over two tables filled in the same backward sweep
//    (first root per label value, next root with the same value) and the end
//    slot of every root's segment.
#include "p2v_common.cuh"
#include "p2v_internal.h"

namespace p2v {

constexpr int SMALL_WARPS = 4;             // warps (= trees in flight) per CTA, small kernel
constexpr int LARGE_THREADS = 256;
constexpr int LARGE_WARPS = LARGE_THREADS / 32;
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
// Lane i scans the later lanes i+1.. in time order; SHFL.DOWN past lane 31 hands back the
// lane's own value, which always compares <= R, so the `lane` wrapped steps are taken off
// at the end (they all come after the real ones).  The 31 shuffles do not depend on R and
// are issued back to back; only the compare/add chain is serial.
__device__ __forceinline__ uint32_t batch_rank(const uint16_t* __restrict__ sv, int b, int k, int lane) {
    const int t = b * 32 + lane;
    const uint32_t x = sv[t];
    const uint32_t ins = (t < k) ? (x <= (uint32_t)t ? 0u : x - (uint32_t)t) : BIG;
    uint32_t R = ins;
#pragma unroll
    for (int d = 1; d < 32; ++d) {
        const uint32_t y = __shfl_down_sync(FULL, ins, d);
        R += (y <= R) ? 1u : 0u;
    }
    return R - (uint32_t)lane;
}

template <typename IdxT, int MODE>
__global__ void __launch_bounds__(SMALL_WARPS * 32)
decode_small_kernel(const IdxT* __restrict__ v, int64_t B, int k, IdxT* __restrict__ out,
                    int64_t out_stride, int32_t* __restrict__ status, int K) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int KA = K + 32;
    const int narr = (MODE == MODE_PAIRS) ? 2 : 3;
    uint16_t* base = reinterpret_cast<uint16_t*>(smem_raw) + (size_t)warp * (narr * KA + 64);
    uint16_t* sv = base;            // v[t]                                   (by time)
    uint16_t* stime = sv + KA;      // element at slot s                      (by slot)
    uint16_t* stbl = stime + KA;    // last row so far whose label is x       (by label)  [!PAIRS]
    uint32_t* smask = reinterpret_cast<uint32_t*>(base + narr * KA);  // 32 words of free-slot bits
    const int nb = K >> 5;

    for (int64_t tree = (int64_t)blockIdx.x * SMALL_WARPS + warp; tree < B;
         tree += (int64_t)gridDim.x * SMALL_WARPS) {
        const IdxT* vt = v + tree * k;
        // ---- A: load (8 loads in flight per lane), validate (check_v), narrow ----------
        int bad = -1;
        for (int b0 = 0; b0 < nb; b0 += 8) {
            long long xs[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = (b0 + j) * 32 + lane;
                xs[j] = (i < k) ? (long long)__ldg(vt + i) : 0;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int i = (b0 + j) * 32 + lane;
                if (b0 + j < nb) {
                    const bool ok = (i >= k) || (xs[j] >= 0 && xs[j] <= 2LL * i);
                    const unsigned bm = __ballot_sync(FULL, !ok);
                    if (bm && bad < 0) bad = (b0 + j) * 32 + __ffs(bm) - 1;
                    sv[i] = (uint16_t)xs[j];
                    if (MODE != MODE_PAIRS) stbl[i] = (uint16_t)NONE16;
                }
            }
        }
        if (MODE != MODE_PAIRS) stbl[K + lane] = (uint16_t)NONE16;
        if (bad >= 0) {
            if (lane == 0 && status) status[tree] = bad + 1;
            __syncwarp();
            continue;
        }
        if (lane == 0 && status) status[tree] = 0;
        // free-slot mask: bit i of word w <-> slot 32w+i, 1 = free
        uint32_t word = (lane * 32 + 32 <= k) ? 0xFFFFFFFFu
                        : (lane * 32 >= k ? 0u : ((1u << (k - lane * 32)) - 1u));
        smask[lane] = word;
        __syncwarp();

        // ---- B: slot assignment, walking time backwards; the rank scan of the next batch
        //         is independent of the placement of the current one --------------------------
        uint32_t Rnext = batch_rank(sv, nb - 1, k, lane);
        for (int b = nb - 1; b >= 0; --b) {
            const int t = b * 32 + lane;
            const bool act = t < k;
            const uint32_t R = Rnext;
            if (b > 0) Rnext = batch_rank(sv, b - 1, k, lane);
            // rank-select R among the free slots
            const uint32_t cnt = __popc(word);
            const uint32_t incl = warp_incl_scan(cnt, lane);
            uint32_t w = 0;
#pragma unroll
            for (int step = 16; step >= 1; step >>= 1) {
                const uint32_t y = __shfl_sync(FULL, incl, (w + step - 1) & 31);
                if (y <= R) w += step;
            }
            w &= 31;
            const uint32_t wd = __shfl_sync(FULL, word, w);
            const uint32_t ex = __shfl_sync(FULL, incl - cnt, w);
            const uint32_t pos = nth_set_bit(wd, R - ex);
            if (act) {
                stime[w * 32 + pos] = (uint16_t)t;
                atomicAnd(&smask[w], ~(1u << pos));
            }
            __syncwarp();
            word = smask[lane];
        }

        // ---- D: forward sweep over the slots, 32 rows at a time -------------------------------
        // label of a row = v[t] of the nearest root (rank-0 insertion) at or left of it.
        // Ancestry / edges: the reference's parents[] pass (convert.rs:174-185) only ever looks
        // up the last earlier row with a given label; stbl keeps that row per label, same-batch
        // occurrences are resolved with MATCH.ANY.
        IdxT* ot = out + tree * out_stride;
        uint32_t lab_carry = 0;
        for (int b = 0; b < nb; ++b) {
            const int s = b * 32 + lane;
            const bool act = s < k;
            const uint32_t t = act ? stime[s] : 0u;
            const uint32_t x = sv[t];
            const bool root = act && x <= t;
            const uint32_t m = __ballot_sync(FULL, root) & lanes_le(lane);
            const int src = m ? 31 - __clz(m) : 0;
            uint32_t lab = __shfl_sync(FULL, x, src);
            if (!m) lab = lab_carry;
            lab_carry = __shfl_sync(FULL, lab, 31);
            if (MODE == MODE_PAIRS) {
                if (act) store_pair<IdxT>(ot + 2 * (size_t)s, (IdxT)lab, (IdxT)(t + 1));
            } else {
                const uint32_t peers = __match_any_sync(FULL, act ? lab : (0x10000u + lane));
                uint32_t c1p = (uint32_t)(k + s);            // non-root: the previous row's parent
                if (root) {
                    const uint32_t lower = peers & lanes_lt(lane);
                    const uint32_t cand = lower ? (uint32_t)(b * 32 + 31 - __clz(lower)) : (uint32_t)stbl[lab];
                    c1p = (cand != NONE16) ? (uint32_t)(k + 1) + cand : lab;
                }
                __syncwarp();
                if (act && (peers & lanes_gt(lane)) == 0) stbl[lab] = (uint16_t)s;
                __syncwarp();
                if (act) {
                    const uint32_t q = stbl[t + 1];
                    const uint32_t c2p = (q != NONE16) ? (uint32_t)(k + 1) + q : t + 1;
                    store_row<IdxT, MODE>(ot, s, (IdxT)c1p, (IdxT)c2p, (IdxT)(k + 1 + s));
                }
            }
        }
        __syncwarp();
    }
}

// --------------------------------------------------------------------------
// k > 1024: one CTA per tree.  Per-element arrays live in a global workspace
// (L2 resident), the free-slot mask and its Fenwick tree in shared memory.
// --------------------------------------------------------------------------
template <typename IdxT, int MODE>
__global__ void __launch_bounds__(LARGE_THREADS)
decode_large_kernel(const IdxT* __restrict__ v, int64_t B, int k, IdxT* __restrict__ out,
                    int64_t out_stride, int32_t* __restrict__ status, uint32_t* __restrict__ ws,
                    size_t ws_stride) {
    extern __shared__ __align__(16) uint32_t sm32[];
    __shared__ int s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nw = (k + 31) >> 5;
    const int KA = ((k + 1 + 31) & ~31) + 32;
    uint32_t* bits = sm32;            // nw words, 1 = free slot
    uint32_t* fen = bits + nw;        // fen[1..nw]: Fenwick tree over popc(bits[w])
    uint32_t* grp = fen + nw + 1;     // nw: per 32-slot group scan scratch
    uint32_t* wv = ws + (size_t)blockIdx.x * ws_stride;
    uint32_t* wR = wv + KA;
    uint32_t* wtime = wR + KA;
    uint32_t* wlast = wtime + KA;
    uint32_t* wnext = wlast + KA;
    uint32_t* wseg = wnext + KA;
    int hb = 1;
    while (hb * 2 <= nw) hb *= 2;

    for (int64_t tree = blockIdx.x; tree < B; tree += gridDim.x) {
        const IdxT* vt = v + tree * k;
        if (tid == 0) s_bad = 0x7FFFFFFF;
        __syncthreads();
        // ---- A ---------------------------------------------------------------------
        for (int i = tid; i < k; i += LARGE_THREADS) {
            long long x = (long long)__ldg(vt + i);
            if (x < 0 || x > 2LL * i) { atomicMin(&s_bad, i); x = 0; }
            wv[i] = (uint32_t)x;
            wlast[i] = NONE32;
        }
        if (tid == 0) wlast[k] = NONE32;
        for (int w = tid; w < nw; w += LARGE_THREADS)
            bits[w] = (w * 32 + 32 <= k) ? 0xFFFFFFFFu : ((1u << (k - w * 32)) - 1u);
        for (int i = tid + 1; i <= nw; i += LARGE_THREADS) {
            const int lo = i - (i & -i);
            fen[i] = (uint32_t)(min(32 * i, k) - min(32 * lo, k));
        }
        __syncthreads();
        if (s_bad != 0x7FFFFFFF) {
            if (tid == 0 && status) status[tree] = s_bad + 1;
            __syncthreads();
            continue;
        }
        if (tid == 0 && status) status[tree] = 0;
        // ---- B1: in-batch ranks, all warps ---------------------------------------------
        for (int b = warp; b < nw; b += LARGE_WARPS) {
            const int t = b * 32 + lane;
            const bool act = t < k;
            const uint32_t x = act ? wv[t] : 0u;
            const uint32_t ins = act ? (x <= (uint32_t)t ? 0u : x - (uint32_t)t) : BIG;
            uint32_t R = ins;
#pragma unroll
            for (int d = 1; d < 32; ++d) rank_step(R, ins, d);
            if (act) wR[t] = R;
        }
        __syncthreads();
        // ---- B2: warp 0 places the batches (serial over batches); warp 1 fills the
        //          root-value tables concurrently -------------------------------------------
        if (warp == 0) {
            for (int b = nw - 1; b >= 0; --b) {
                const int t = b * 32 + lane;
                const bool act = t < k;
                uint32_t rem = act ? wR[t] : 0u;
                int pos = 0;
                for (int step = hb; step; step >>= 1) {
                    const int np = pos + step;
                    if (np <= nw) {
                        const uint32_t f = fen[np];
                        if (f <= rem) { pos = np; rem -= f; }
                    }
                }
                uint32_t bit = 0;
                if (act) {
                    bit = nth_set_bit(bits[pos], rem);
                    wtime[pos * 32 + bit] = (uint32_t)t;
                }
                __syncwarp();
                if (act) {
                    atomicAnd(&bits[pos], ~(1u << bit));
                    for (int i = pos + 1; i <= nw; i += i & -i) atomicSub(&fen[i], 1u);
                }
                __syncwarp();
            }
        } else if (warp == 1 && MODE != MODE_PAIRS) {
            for (int b = nw - 1; b >= 0; --b) {
                const int t = b * 32 + lane;
                const bool act = t < k;
                const uint32_t x = act ? wv[t] : 0u;
                const bool root = act && x <= (uint32_t)t;
                // distinct keys for non-roots: values are < 2^20, lanes tag bit 30
                const uint32_t key = root ? x : (0x40000000u + lane);
                const uint32_t peers = __match_any_sync(FULL, key);
                const uint32_t higher = peers & lanes_gt(lane);
                uint32_t nxt = NONE32;
                if (root) nxt = higher ? (uint32_t)(b * 32 + __ffs(higher) - 1) : wlast[x];
                __syncwarp();
                if (root) {
                    wnext[t] = nxt;
                    if ((peers & lanes_lt(lane)) == 0) wlast[x] = (uint32_t)t;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        IdxT* ot = out + tree * out_stride;
        if (MODE == MODE_PAIRS) {
            // last root slot at or left of every slot: per-group last root, prefix-max
            for (int g = warp; g < nw; g += LARGE_WARPS) {
                const int s = g * 32 + lane;
                const bool act = s < k;
                const uint32_t t = act ? wtime[s] : 0u;
                const uint32_t x = act ? wv[t] : 1u;
                const uint32_t rm = __ballot_sync(FULL, act && x <= t);
                if (lane == 0) grp[g] = rm ? (uint32_t)(g * 32 + 31 - __clz(rm)) : 0u;
            }
            __syncthreads();
            if (warp == 0) {  // exclusive prefix max over groups (slot 0 is always a root)
                uint32_t carry = 0;
                for (int c = 0; c < nw; c += 32) {
                    const int g = c + lane;
                    const uint32_t val = g < nw ? grp[g] : 0u;
                    uint32_t m = val;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t y = __shfl_up_sync(FULL, m, d);
                        if (lane >= d) m = max(m, y);
                    }
                    uint32_t ex = __shfl_up_sync(FULL, m, 1);
                    if (lane == 0) ex = 0;
                    ex = max(ex, carry);
                    carry = max(carry, __shfl_sync(FULL, m, 31));
                    if (g < nw) grp[g] = ex;
                }
            }
            __syncthreads();
            for (int g = warp; g < nw; g += LARGE_WARPS) {
                const int s = g * 32 + lane;
                const bool act = s < k;
                const uint32_t t = act ? wtime[s] : 0u;
                const uint32_t x = act ? wv[t] : 1u;
                const uint32_t m = __ballot_sync(FULL, act && x <= t) & lanes_le(lane);
                if (act) {
                    const uint32_t rs = m ? (uint32_t)(g * 32 + 31 - __clz(m)) : grp[g];
                    const uint32_t lab = wv[wtime[rs]];
                    store_pair<IdxT>(ot + 2 * (size_t)s, (IdxT)lab, (IdxT)(t + 1));
                }
            }
        } else {
            // first root slot right of every slot: per-group first root, suffix-min
            for (int g = warp; g < nw; g += LARGE_WARPS) {
                const int s = g * 32 + lane;
                const bool act = s < k;
                const uint32_t t = act ? wtime[s] : 0u;
                const uint32_t x = act ? wv[t] : 1u;
                const uint32_t rm = __ballot_sync(FULL, act && x <= t);
                if (lane == 0) grp[g] = rm ? (uint32_t)(g * 32 + __ffs(rm) - 1) : NONE32;
            }
            __syncthreads();
            if (warp == 0) {  // exclusive suffix min over groups, "none" == k
                uint32_t carry = (uint32_t)k;
                for (int c = ((nw - 1) / 32) * 32; c >= 0; c -= 32) {
                    const int g = c + lane;
                    const uint32_t val = g < nw ? grp[g] : NONE32;
                    uint32_t m = val;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const uint32_t y = __shfl_down_sync(FULL, m, d);
                        if (lane + d < 32) m = min(m, y);
                    }
                    uint32_t ex = __shfl_down_sync(FULL, m, 1);
                    if (lane == 31) ex = NONE32;
                    ex = min(ex, carry);
                    carry = min(carry, __shfl_sync(FULL, m, 0));
                    if (g < nw) grp[g] = ex;
                }
            }
            __syncthreads();
            for (int g = warp; g < nw; g += LARGE_WARPS) {
                const int s = g * 32 + lane;
                const bool act = s < k;
                const uint32_t t = act ? wtime[s] : 0u;
                const uint32_t x = act ? wv[t] : 1u;
                const bool root = act && x <= t;
                const uint32_t rm = __ballot_sync(FULL, root);
                const uint32_t higher = rm & lanes_gt(lane);
                if (root) wseg[t] = (uint32_t)k + (higher ? (uint32_t)(g * 32 + __ffs(higher) - 1) : grp[g]);
            }
            __syncthreads();
            for (int s = tid; s < k; s += LARGE_THREADS) {
                const uint32_t t = wtime[s];
                const uint32_t x = wv[t];
                const uint32_t j2 = wlast[t + 1];
                const uint32_t c2p = (j2 != NONE32) ? wseg[j2] : t + 1;
                uint32_t c1p;
                if (x <= t) {
                    const uint32_t j1 = wnext[t];
                    c1p = (j1 != NONE32) ? wseg[j1] : x;
                } else {
                    c1p = (uint32_t)(k + s);
                }
                store_row<IdxT, MODE>(ot, s, (IdxT)c1p, (IdxT)c2p, (IdxT)(k + 1 + s));
            }
        }
        __syncthreads();
    }
}

// --------------------------------------------------------------------------
// check_v only
// --------------------------------------------------------------------------
template <typename IdxT>
__global__ void check_v_kernel(const IdxT* __restrict__ v, int64_t B, int k, int32_t* __restrict__ status) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t tree = wid; tree < B; tree += nwarps) {
        const IdxT* vt = v + tree * k;
        int bad = -1;
        for (int b = 0; b < k && bad < 0; b += 32) {
            const int i = b + lane;
            bool ok = true;
            if (i < k) { const long long x = (long long)__ldg(vt + i); ok = x >= 0 && x <= 2LL * i; }
            const unsigned bm = __ballot_sync(FULL, !ok);
            if (bm) bad = b + __ffs(bm) - 1;
        }
        if (lane == 0) status[tree] = bad + 1;
    }
}

// --------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------
static size_t large_ws_stride(int64_t k) {  // uint32 elements per CTA
    const size_t KA = (((size_t)k + 1 + 31) & ~(size_t)31) + 32;
    return 6 * KA;
}
static int large_grid(int64_t B) {
    const int64_t cap = (int64_t)sm_count() * 4;
    return (int)(B < cap ? B : cap);
}

size_t decode_workspace_bytes(int64_t B, int64_t k) {
    if (k <= SMALL_K_MAX || B <= 0) return 0;
    return (size_t)large_grid(B) * large_ws_stride(k) * sizeof(uint32_t);
}

template <typename IdxT, int MODE>
static cudaError_t launch_decode_t(const void* v, int64_t B, int64_t k, void* out, int32_t* status,
                                   void* ws, size_t ws_bytes, cudaStream_t stream, int64_t out_stride) {
    if (out_stride <= 0) out_stride = k * (MODE == MODE_PAIRS ? 2 : (MODE == MODE_ANCESTRY ? 3 : 4));
    if (k <= SMALL_K_MAX) {
        const int K = (int)((k + 31) & ~31);
        const int narr = (MODE == MODE_PAIRS) ? 2 : 3;
        const size_t smem = (size_t)SMALL_WARPS * ((size_t)narr * (K + 32) + 64) * sizeof(uint16_t);
        auto kern = decode_small_kernel<IdxT, MODE>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SMALL_WARPS * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        int64_t want = (B + SMALL_WARPS - 1) / SMALL_WARPS;
        int64_t cap = (int64_t)sm_count() * per_sm;
        int grid = (int)(want < cap ? want : cap);
        kern<<<grid, SMALL_WARPS * 32, smem, stream>>>((const IdxT*)v, B, (int)k, (IdxT*)out, out_stride, status, K);
    } else {
        const size_t stride = large_ws_stride(k);
        const int grid = large_grid(B);
        if (!ws || ws_bytes < (size_t)grid * stride * sizeof(uint32_t)) return cudaErrorInvalidValue;
        const int nw = (int)((k + 31) >> 5);
        const size_t smem = ((size_t)3 * nw + 8) * sizeof(uint32_t);
        auto kern = decode_large_kernel<IdxT, MODE>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kern<<<grid, LARGE_THREADS, smem, stream>>>((const IdxT*)v, B, (int)k, (IdxT*)out, out_stride,
                                                   status, (uint32_t*)ws, stride);
    }
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_decode(int mode, const void* v, int idx_bytes, int64_t B, int64_t k, void* out,
                          int32_t* status, void* ws, size_t ws_bytes, cudaStream_t stream,
                          int64_t out_stride) {
    if (B <= 0 || k <= 0) return cudaSuccess;
#define P2V_DISPATCH(T)                                                                            \
    switch (mode) {                                                                                \
    case MODE_PAIRS: return launch_decode_t<T, MODE_PAIRS>(v, B, k, out, status, ws, ws_bytes, stream, out_stride);     \
    case MODE_ANCESTRY: return launch_decode_t<T, MODE_ANCESTRY>(v, B, k, out, status, ws, ws_bytes, stream, out_stride); \
    case MODE_EDGES: return launch_decode_t<T, MODE_EDGES>(v, B, k, out, status, ws, ws_bytes, stream, out_stride);     \
    default: return cudaErrorInvalidValue;                                                         \
    }
    if (idx_bytes == 8) { P2V_DISPATCH(long long) }
    if (idx_bytes == 4) { P2V_DISPATCH(int) }
#undef P2V_DISPATCH
    return cudaErrorInvalidValue;
}

cudaError_t launch_check_v(const void* v, int idx_bytes, int64_t B, int64_t k, int32_t* status,
                           cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    const int threads = 256;
    int64_t want = (B * 32 + threads - 1) / threads;
    int64_t cap = (int64_t)sm_count() * 8;
    const int grid = (int)(want < cap ? want : cap);
    if (idx_bytes == 8)
        check_v_kernel<long long><<<grid, threads, 0, stream>>>((const long long*)v, B, (int)k, status);
    else if (idx_bytes == 4)
        check_v_kernel<int><<<grid, threads, 0, stream>>>((const int*)v, B, (int)k, status);
    else
        return cudaErrorInvalidValue;
    count_launch();
    return cudaGetLastError();
}

}  // namespace p2v
