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
//    free-slot bit mask (one 32-bit word per lane + popc prefix).  Trees with k > 1024
//    sort the batches by rank and merge them pairwise instead (see decode_large_kernel).
//  * Rank-0 insertions ("roots") carry their own label v[t]; every other element
//    inherits the label of the nearest root on its left, so pair labels are a
//    segmented broadcast over slots.
//  * The sequential parents[] relabel pass of to_ancestry/to_edges is replaced by
//    closed forms over two tables filled in the same backward sweep
//    (first root per label value, next root with the same value) and the end
//    slot of every root's segment.
#include <cstdlib>

#include "p2v_decode_warp.cuh"
#include "p2v_decode_warp_big.cuh"

namespace p2v {

cudaError_t debug_phase_ns_decode(unsigned long long* out) {
    return cudaMemcpyFromSymbol(out, g_phase_ns, sizeof(unsigned long long) * PHASE_SLOTS);
}

constexpr int SMALL_WARPS = 4;             // warps (= trees in flight) per CTA, small kernel

// uint32 entries of the merge kernel's per-team work arrays (K = k rounded up to 32)
__host__ __device__ inline size_t large_base_stride(size_t K) { return 7 * K + (K >> 5) + 128; }

// One warp per tree, all state in shared memory (p2v_decode_warp.cuh).  k <= 2016: four warps
// per CTA; mid-size trees (k <= 16384): one warp per CTA, so that the per-tree shared memory
// (6 bytes per leaf) still leaves several trees in flight per SM.
template <typename IdxT, int MODE, int WPL>
__global__ void __launch_bounds__(SMALL_WARPS * 32)
decode_small_kernel(const IdxT* __restrict__ v, int64_t B, int k, IdxT* __restrict__ out,
                    int64_t out_stride, int32_t* __restrict__ status, int K) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint8_t s_lut[WPL > 2 ? 16 : 2048];       // the light form selects bits with ALU instructions
    if (WPL <= 2) nth_bit_lut_init(s_lut, threadIdx.x, blockDim.x);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    uint16_t* base = reinterpret_cast<uint16_t*>(smem_raw) + (size_t)warp * decode_warp_scratch_u16(K, MODE, WPL);
    for (int64_t tree = (int64_t)blockIdx.x * nwarps + warp; tree < B;
         tree += (int64_t)gridDim.x * nwarps) {
        const int bad = decode_tree_warp<IdxT, IdxT, MODE, WPL>(v + tree * k, k, K, out + tree * out_stride, base, lane, s_lut);
        if (lane == 0 && status) status[tree] = bad + 1;
    }
}

// Batches of big trees (16384 < k <= 262143): one warp per tree, one tree per CTA (p2v_decode_warp_big.cuh);
// CTA c keeps its label table in slice c of the workspace.
template <typename IdxT, int MODE, typename TblT>
__global__ void __launch_bounds__(32)
decode_big_warp_kernel(const IdxT* __restrict__ v, int64_t B, int k, IdxT* __restrict__ out, int64_t out_stride,
                       int32_t* __restrict__ status, int K, TblT* __restrict__ tables, size_t table_stride) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    TblT* gtbl = tables + (size_t)blockIdx.x * table_stride;
    for (int64_t tree = blockIdx.x; tree < B; tree += gridDim.x) {
        const int bad = decode_tree_warp_big<IdxT, IdxT, MODE, TblT>(v + tree * k, k, K, out + tree * out_stride,
                                                                     reinterpret_cast<uint16_t*>(smem_raw), gtbl, lane);
        if (lane == 0 && status) status[tree] = bad + 1;
        __syncwarp();
    }
}

// --------------------------------------------------------------------------
// k > 1024: one team per tree -- a CTA for batches, a thread-block cluster of 16 (8) CTAs when
// there are only a few large trees.  Per-element arrays live in a global workspace (L2 resident).
//
// Slot assignment without any serial pass: level 0 ranks every element inside its 32-element
// batch (the same SHFL scan as the small kernel) and stores the batch sorted by rank; then blocks
// are merged pairwise, log2(k/32) levels.  For a pair X (earlier times) | Y (later times), the
// ranks of Y are already final for the merged block; an X element of rank r moves up by the
// number of Y elements that land before it, #{l : y_l - l <= r}, and a Y element (index l, rank
// y) ends up after #{x : r_x < y - l} X elements -- one binary search per element and level.
// After the last level the array of elements sorted by rank IS the slot order.
// Labels: segmented broadcast with a per-32 group scan.  Ancestry / edges: the forward relabel
// sweep of the small kernel, run by one warp with the last-row-per-label table in shared memory.
// --------------------------------------------------------------------------
// WT / SMEMARR: work arrays as uint16 in shared memory for mid-size trees (k <= MID_K_MAX),
// as uint32 in the global workspace above that.
template <typename IdxT, int MODE, int CSIZE, int THREADS, bool SMEMTBL, typename WT, bool SMEMARR>
__global__ void __launch_bounds__(THREADS)
decode_large_kernel(const IdxT* __restrict__ v, int64_t B, int k, IdxT* __restrict__ out,
                    int64_t out_stride, int32_t* __restrict__ status, uint32_t* __restrict__ ws,
                    size_t ws_stride) {
    extern __shared__ __align__(16) uint16_t s_tbl[];        // k+2 entries (rounded to 8) when SMEMTBL
    constexpr uint32_t ROOTBIT = 1u << (sizeof(WT) * 8 - 1);
    const int crank = team_rank<CSIZE>();
    const int tid = crank * THREADS + threadIdx.x;            // thread index inside the team
    const int TT = team_ctas<CSIZE>() * THREADS;
    const int NWT = TT / 32;
    const int lane = threadIdx.x & 31;
    const int gw = tid >> 5;                                   // warp index inside the team
    const int K = (k + 31) & ~31;
    const int nbat = K >> 5;
    const int64_t team = team_index<CSIZE>(), nteams = team_count<CSIZE>();
    const size_t tbl_entries = SMEMTBL ? (((size_t)k + 2 + 7) & ~(size_t)7) : 0;
    WT* wv = SMEMARR ? reinterpret_cast<WT*>(s_tbl + tbl_entries)
                     : reinterpret_cast<WT*>(ws + (size_t)team * ws_stride);
    WT* rk0 = wv + K;
    WT* rk1 = rk0 + K;
    WT* el0 = rk1 + K;
    WT* el1 = el0 + K;
    WT* wlab = el1 + K;
    WT* grp = wlab + K;                       // nbat + 32
    WT* gtbl = grp + nbat + 32;               // K + 32 (only without SMEMTBL)
    int* flags = reinterpret_cast<int*>(gtbl + (SMEMTBL ? 0 : K + 32) + ((nbat & 1) ? 1 : 0));   // 16, 4-byte aligned

    for (int64_t tree = team; tree < B; tree += nteams) {
        const IdxT* vt = v + tree * k;
        phase_stamp(15, CSIZE == 0 && tid == 0);
        if (tid < 16) flags[tid] = 0;
        // ---- A: load, validate, narrow ---------------------------------------------------
        int badlocal = 0x7FFFFFFF;
        for (int i = tid; i < K; i += TT) {
            long long x = (i < k) ? (long long)__ldg(vt + i) : 0;
            if (i < k && (x < 0 || x > 2LL * i)) { badlocal = min(badlocal, i); x = 0; }
            wv[i] = (WT)x;
        }
        team_sync<CSIZE>();
        phase_stamp(0, CSIZE == 0 && tid == 0);
        if (badlocal != 0x7FFFFFFF) atomicMax(&flags[1], K - badlocal);     // K - i >= 1
        team_sync<CSIZE>();
        const int badmark = *reinterpret_cast<volatile int*>(&flags[1]);
        if (badmark) {
            if (tid == 0 && status) status[tree] = (K - badmark) + 1;
            team_sync<CSIZE>();
            continue;
        }
        if (tid == 0 && status) status[tree] = 0;

        // ---- level 0: rank inside each batch of 32, stored sorted by rank ----------------------
        // Whole-grid teams: warp-sized work is dealt round-robin over the CTAs, and never to warp 0 of a CTA
        // (the master thread's warp runs shuffle-heavy code ~10x slower after a grid barrier, see
        // p2v_cophenetic.cu blocks_kernel); gwr = -1 marks the idle warp, NWR the stride.
        const int wq0 = (int)(threadIdx.x >> 5);
        const int NWR = (CSIZE == 0) ? (THREADS / 32 - 1) * team_ctas<CSIZE>() : NWT;
        const int gwr = (CSIZE == 0) ? (wq0 > 0 ? (wq0 - 1) * team_ctas<CSIZE>() + crank : NWR * 0 + 0x40000000) : gw;
        for (int b = gwr; b < nbat; b += NWR) {
            const int t = b * 32 + lane;
            const bool act = t < k;
            const uint32_t x = wv[t];
            const uint32_t ins = act ? (x <= (uint32_t)t ? 0u : x - (uint32_t)t) : BIG;
            uint32_t R = ins;
#pragma unroll
            for (int d = 1; d < 32; ++d) {
                const uint32_t y = __shfl_down_sync(FULL, ins, d);
                R += (y <= R) ? 1u : 0u;
            }
            R -= (uint32_t)lane;
            const uint32_t key = act ? R : BIG;
            uint32_t p = 0;
#pragma unroll
            for (int d = 1; d < 32; ++d) {
                const uint32_t y = __shfl_sync(FULL, key, (lane + d) & 31);
                p += (y < key) ? 1u : 0u;
            }
            if (act) { rk0[b * 32 + p] = (WT)R; el0[b * 32 + p] = (WT)t; }
        }
        team_sync<CSIZE>();
        phase_stamp(1, CSIZE == 0 && tid == 0);

        // ---- merge levels ----------------------------------------------------------------------
        int cur = 0;
        for (int S = 32; S < k; S <<= 1) {
            const WT* rin = cur ? rk1 : rk0;
            const WT* ein = cur ? el1 : el0;
            WT* rout = cur ? rk0 : rk1;
            WT* eout = cur ? el0 : el1;
            for (int i = ((CSIZE == 0) ? (wq0 * team_ctas<CSIZE>() + crank) * 32 + lane : tid); i < k; i += TT) {
                const int a = (i / (2 * S)) * (2 * S), m = a + S;
                const uint32_t r = rin[i], e = ein[i];
                if (m >= k) { rout[i] = (WT)r; eout[i] = (WT)e; continue; }      // lone block: copy through
                const int end = min(k, a + 2 * S);
                if (i < m) {                       // X: how many Y elements land before me
                    int lo = 0, hi = end - m;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if ((uint32_t)rin[m + mid] - (uint32_t)mid <= r) lo = mid + 1; else hi = mid;
                    }
                    rout[i + lo] = (WT)(r + (uint32_t)lo); eout[i + lo] = (WT)e;
                } else {                           // Y: how many X elements stay before me
                    const int l = i - m;
                    const uint32_t z = r - (uint32_t)l;
                    int lo = 0, hi = S;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if ((uint32_t)rin[a + mid] < z) lo = mid + 1; else hi = mid;
                    }
                    rout[a + l + lo] = (WT)r; eout[a + l + lo] = (WT)e;
                }
            }
            cur ^= 1;
            team_sync<CSIZE>();
        }
        phase_stamp(2, CSIZE == 0 && tid == 0);
        const WT* time_at = cur ? el1 : el0;                  // element at slot s

        // ---- labels: nearest root (rank-0 insertion) at or left of every slot ------------------
        for (int g = gwr; g < nbat; g += NWR) {
            const int s = g * 32 + lane;
            const bool act = s < k;
            const uint32_t t = act ? time_at[s] : 0u;
            const uint32_t x = act ? wv[t] : 1u;
            const uint32_t rm = __ballot_sync(FULL, act && x <= t);
            if (lane == 0) grp[g] = (WT)(rm ? (uint32_t)(g * 32 + 31 - __clz(rm)) : 0u);
        }
        team_sync<CSIZE>();
        if (crank == 0) {  // exclusive prefix max over groups (slot 0 is always a root): block scan in CTA 0
            __shared__ uint32_t s_scan[32];
            const int per = (nbat + THREADS - 1) / THREADS;
            const int g0 = min(nbat, (int)threadIdx.x * per), g1 = min(nbat, g0 + per);
            uint32_t loc = 0;
            for (int g = g0; g < g1; ++g) loc = max(loc, (uint32_t)grp[g]);
            uint32_t inc = loc;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(FULL, inc, d);
                if (lane >= d) inc = max(inc, y);
            }
            const int wq = threadIdx.x >> 5;
            if (lane == 31) s_scan[wq] = inc;
            __syncthreads();
            if (wq == 0) {
                uint32_t t = (lane < THREADS / 32) ? s_scan[lane] : 0u;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t y = __shfl_up_sync(FULL, t, d);
                    if (lane >= d) t = max(t, y);
                }
                uint32_t ex = __shfl_up_sync(FULL, t, 1);
                if (lane == 0) ex = 0;
                s_scan[lane] = ex;
            }
            __syncthreads();
            uint32_t run = __shfl_up_sync(FULL, inc, 1);
            if (lane == 0) run = 0;
            run = max(run, s_scan[wq]);
            for (int g = g0; g < g1; ++g) {
                const uint32_t cur = grp[g];
                grp[g] = (WT)run;
                run = max(run, cur);
            }
        }
        team_sync<CSIZE>();
        phase_stamp(3, CSIZE == 0 && tid == 0);
        IdxT* ot = out + tree * out_stride;
        for (int g = gwr; g < nbat; g += NWR) {
            const int s = g * 32 + lane;
            const bool act = s < k;
            const uint32_t t = act ? time_at[s] : 0u;
            const uint32_t x = act ? wv[t] : 1u;
            const bool root = act && x <= t;
            const uint32_t m = __ballot_sync(FULL, root) & lanes_le(lane);
            if (act) {
                const uint32_t rs = m ? (uint32_t)(g * 32 + 31 - __clz(m)) : grp[g];
                const uint32_t lab = root ? x : wv[time_at[rs]];
                if (MODE == MODE_PAIRS) store_pair<IdxT>(ot + 2 * (size_t)s, (IdxT)lab, (IdxT)(t + 1));
                else wlab[s] = (WT)(lab | (root ? ROOTBIT : 0u));
            }
        }
        phase_stamp(4, CSIZE == 0 && tid == 0);
        if (MODE == MODE_PAIRS) { team_sync<CSIZE>(); continue; }

        // ---- ancestry / edges, one huge tree on the whole grid, slots fit 16 bits: one worker per CTA with
        //      its "last row with label x" table in SHARED memory.  1. warp 1 of every CTA records the last
        //      row of each label inside its slot range; 2. the tables are published (uint16) and an exclusive
        //      prefix over the workers -- eight labels per thread and step -- turns table w into "last row with
        //      label x before range w"; 3. every CTA loads its prefix state back and warp 1 runs the ordinary
        //      sweep on its range.  All table traffic of the two sequential passes is shared-memory latency.
        if (CSIZE == 0 && SMEMTBL && !SMEMARR) {
            const int TBL = (int)tbl_entries;                                  // k + 2 rounded up to 8
            const size_t base16 = (large_base_stride((size_t)K) + 3) & ~(size_t)3;     // 16-byte aligned rows
            uint16_t* ptab16 = reinterpret_cast<uint16_t*>(ws + (size_t)team * ws_stride + base16);
            const int cap = (int)(((ws_stride - base16) * 2) / (size_t)TBL);
            const int W = min(team_ctas<CSIZE>(), cap);
            const int bpw = (nbat + W - 1) / W;
            const int w = crank;
            const int b0 = min(nbat, w * bpw), b1 = (w < W) ? min(nbat, b0 + bpw) : b0;
            const int wq = threadIdx.x >> 5;
            uint4 ones; ones.x = ones.y = ones.z = ones.w = 0xFFFFFFFFu;
            for (int i = threadIdx.x * 8; i < TBL; i += THREADS * 8) *reinterpret_cast<uint4*>(s_tbl + i) = ones;
            team_sync<CSIZE>();                          // the labels of every slot are written (and the table is filled)
            if (wq == 1) {            // not warp 0: see gwr above
                // (the labels come from L2: the loads of the next two batches are in flight while one is recorded)
                auto ldlab = [&](int b) -> uint32_t { const int s = b * 32 + lane; return (b < b1 && s < k) ? (uint32_t)wlab[s] : 0u; };
                uint32_t l1 = ldlab(b0), l2 = ldlab(b0 + 1);
                for (int b = b0; b < b1; ++b) {
                    const int s = b * 32 + lane;
                    const bool act = s < k;
                    const uint32_t lab = l1 & (ROOTBIT - 1u);
                    l1 = l2; l2 = ldlab(b + 2);
                    const uint32_t peers = __match_any_sync(FULL, act ? lab : (0x40000000u + lane));
                    if (act && (peers & lanes_gt(lane)) == 0) s_tbl[lab] = (uint16_t)s;
                    __syncwarp();
                }
            }
            __syncthreads();
            if (w < W)
                for (int i = threadIdx.x * 8; i < TBL; i += THREADS * 8)
                    *reinterpret_cast<uint4*>(ptab16 + (size_t)w * TBL + i) = *reinterpret_cast<const uint4*>(s_tbl + i);
            team_sync<CSIZE>();
            phase_stamp(6, tid == 0);
            // exclusive prefix over the workers: four labels per thread; the loads of 16 workers are issued
            // together (a store to the same row between them would serialise the L2 round trips)
            // (warps are dealt round-robin over the CTAs: 4 labels x 32 lanes per warp, every SM takes part)
            const int nct = team_ctas<CSIZE>();
            for (int x4 = ((wq * nct + crank) * 32 + lane) * 4; x4 < TBL; x4 += TT * 4) {
                uint2 run; run.x = run.y = 0xFFFFFFFFu;
                for (int w0 = 0; w0 < W; w0 += 16) {
                    uint2 cur[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (w0 + j < W) cur[j] = __ldcg(reinterpret_cast<const uint2*>(ptab16 + (size_t)(w0 + j) * TBL + x4));
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (w0 + j < W) {
                            __stcg(reinterpret_cast<uint2*>(ptab16 + (size_t)(w0 + j) * TBL + x4), run);
                            // per 16-bit lane: keep `run` where cur is "none" (0xFFFF), else take cur (rows grow with w)
                            const uint32_t mx = __vcmpeq2(cur[j].x, 0xFFFFFFFFu), my = __vcmpeq2(cur[j].y, 0xFFFFFFFFu);
                            run.x = (run.x & mx) | (cur[j].x & ~mx); run.y = (run.y & my) | (cur[j].y & ~my);
                        }
                    }
                }
            }
            team_sync<CSIZE>();
            phase_stamp(7, tid == 0);
            if (w < W)
                for (int i = threadIdx.x * 8; i < TBL; i += THREADS * 8)
                    *reinterpret_cast<uint4*>(s_tbl + i) = *reinterpret_cast<const uint4*>(ptab16 + (size_t)w * TBL + i);
            __syncthreads();
            if (wq == 1 && b0 < b1) {
                uint32_t nlab = (b0 * 32 + lane < k) ? (uint32_t)wlab[b0 * 32 + lane] : 0u;
                uint32_t nt = (b0 * 32 + lane < k) ? (uint32_t)time_at[b0 * 32 + lane] : 0u;
                for (int b = b0; b < b1; ++b) {
                    const int s = b * 32 + lane;
                    const bool act = s < k;
                    const uint32_t labr = nlab, t = nt;
                    if (b + 1 < b1) {                        // prefetch the next batch
                        const int s2 = s + 32;
                        nlab = (s2 < k) ? (uint32_t)wlab[s2] : 0u;
                        nt = (s2 < k) ? (uint32_t)time_at[s2] : 0u;
                    }
                    const uint32_t lab = labr & (ROOTBIT - 1u);
                    const bool root = act && (labr & ROOTBIT);
                    const uint32_t peers = __match_any_sync(FULL, act ? lab : (0x40000000u + lane));
                    uint32_t c1p = (uint32_t)(k + s);
                    if (root) {
                        const uint32_t lower = peers & lanes_lt(lane);
                        uint32_t cand;
                        if (lower) cand = (uint32_t)(b * 32 + 31 - __clz(lower));
                        else { const uint32_t q = s_tbl[lab]; cand = (q == NONE16) ? NONE32 : q; }
                        c1p = (cand != NONE32) ? (uint32_t)(k + 1) + cand : lab;
                    }
                    __syncwarp();
                    if (act && (peers & lanes_gt(lane)) == 0) s_tbl[lab] = (uint16_t)s;
                    __syncwarp();
                    if (act) {
                        const uint32_t q16 = s_tbl[t + 1];
                        const uint32_t c2p = (q16 != NONE16) ? (uint32_t)(k + 1) + q16 : t + 1;
                        store_row<IdxT, MODE>(ot, s, (IdxT)c1p, (IdxT)c2p, (IdxT)(k + 1 + s));
                    }
                }
            }
            team_sync<CSIZE>();
            phase_stamp(8, tid == 0);
            continue;
        }

        // ---- ancestry / edges, few large trees: the relabel sweep split over W warps ------------------
        // Worker w owns a contiguous range of slots and a private "last row with label x" table.
        // 1. every worker records the last row of each label inside its range; 2. an exclusive
        // prefix over the workers turns table w into "last row with label x before range w";
        // 3. every worker runs the sweep below on its range, starting from that state.
        const int K2 = K + 32;
        const int W = (CSIZE != 1) ? (int)((ws_stride - large_base_stride((size_t)K)) / (size_t)K2) : 0;
        if (CSIZE != 1 && W >= 2) {
            uint32_t* ptab = reinterpret_cast<uint32_t*>(ws + (size_t)team * ws_stride) + large_base_stride((size_t)K);
            for (size_t i = tid; i < (size_t)W * K2; i += TT) ptab[i] = NONE32;
            team_sync<CSIZE>();
        phase_stamp(5, CSIZE == 0 && tid == 0);
            const int Wd = (W > 0) ? W : 1;
            const int wstep = max(1, NWT / Wd);              // workers are spread over the CTAs of the team
            const bool worker = (gw % wstep) == 0 && (gw / wstep) < W;
            const int w = gw / wstep;
            const int bpw = (nbat + Wd - 1) / Wd;
            const int b0 = min(nbat, w * bpw), b1 = min(nbat, b0 + bpw);
            uint32_t* mytab = ptab + (size_t)w * K2;
            if (worker) {
                for (int b = b0; b < b1; ++b) {
                    const int s = b * 32 + lane;
                    const bool act = s < k;
                    const uint32_t lab = act ? ((uint32_t)wlab[s] & (ROOTBIT - 1u)) : 0u;
                    const uint32_t peers = __match_any_sync(FULL, act ? lab : (0x40000000u + lane));
                    if (act && (peers & lanes_gt(lane)) == 0) mytab[lab] = (uint32_t)s;
                    __syncwarp();                            // the next batch may write the same label
                }
            }
            team_sync<CSIZE>();
        phase_stamp(6, CSIZE == 0 && tid == 0);
            for (int x = tid; x < K2; x += TT) {             // exclusive prefix over the workers
                uint32_t run = NONE32;
                for (int ww = 0; ww < W; ++ww) {
                    const uint32_t cur = ptab[(size_t)ww * K2 + x];
                    ptab[(size_t)ww * K2 + x] = run;
                    if (cur != NONE32) run = cur;            // rows grow with the worker index
                }
            }
            team_sync<CSIZE>();
        phase_stamp(7, CSIZE == 0 && tid == 0);
            if (worker && b0 < b1) {
                uint32_t nlab = (b0 * 32 + lane < k) ? (uint32_t)wlab[b0 * 32 + lane] : 0u;
                uint32_t nt = (b0 * 32 + lane < k) ? (uint32_t)time_at[b0 * 32 + lane] : 0u;
                for (int b = b0; b < b1; ++b) {
                    const int s = b * 32 + lane;
                    const bool act = s < k;
                    const uint32_t labr = nlab, t = nt;
                    if (b + 1 < b1) {                        // prefetch the next batch
                        const int s2 = s + 32;
                        nlab = (s2 < k) ? (uint32_t)wlab[s2] : 0u;
                        nt = (s2 < k) ? (uint32_t)time_at[s2] : 0u;
                    }
                    const uint32_t lab = labr & (ROOTBIT - 1u);
                    const bool root = act && (labr & ROOTBIT);
                    const uint32_t peers = __match_any_sync(FULL, act ? lab : (0x40000000u + lane));
                    uint32_t c1p = (uint32_t)(k + s);
                    if (root) {
                        const uint32_t lower = peers & lanes_lt(lane);
                        const uint32_t cand = lower ? (uint32_t)(b * 32 + 31 - __clz(lower)) : mytab[lab];
                        c1p = (cand != NONE32) ? (uint32_t)(k + 1) + cand : lab;
                    }
                    __syncwarp();
                    if (act && (peers & lanes_gt(lane)) == 0) mytab[lab] = (uint32_t)s;
                    __syncwarp();
                    if (act) {
                        const uint32_t q = mytab[t + 1];
                        const uint32_t c2p = (q != NONE32) ? (uint32_t)(k + 1) + q : t + 1;
                        store_row<IdxT, MODE>(ot, s, (IdxT)c1p, (IdxT)c2p, (IdxT)(k + 1 + s));
                    }
                }
            }
            team_sync<CSIZE>();
        phase_stamp(8, CSIZE == 0 && tid == 0);
            continue;
        }

        // ---- ancestry / edges: forward relabel sweep, one warp, table of the last row per label --
        if (SMEMTBL) {
            if (crank == 0)
                for (int i = threadIdx.x; i <= k + 1; i += THREADS) s_tbl[i] = (uint16_t)NONE16;
        } else {
            for (int i = tid; i <= k + 1; i += TT) gtbl[i] = (WT)NONE32;
        }
        team_sync<CSIZE>();
        if (gw == 0) {
            uint32_t nlab = (lane < k) ? wlab[lane] : 0u;
            uint32_t nt = (lane < k) ? time_at[lane] : 0u;
            for (int b = 0; b < nbat; ++b) {
                const int s = b * 32 + lane;
                const bool act = s < k;
                const uint32_t labr = nlab, t = nt;
                if (b + 1 < nbat) {                       // prefetch the next batch
                    const int s2 = s + 32;
                    nlab = (s2 < k) ? wlab[s2] : 0u;
                    nt = (s2 < k) ? time_at[s2] : 0u;
                }
                const uint32_t lab = labr & (ROOTBIT - 1u);
                const bool root = act && (labr & ROOTBIT);
                const uint32_t peers = __match_any_sync(FULL, act ? lab : (0x40000000u + lane));
                uint32_t c1p = (uint32_t)(k + s);
                if (root) {
                    const uint32_t lower = peers & lanes_lt(lane);
                    uint32_t cand;
                    if (lower) cand = (uint32_t)(b * 32 + 31 - __clz(lower));
                    else if (SMEMTBL) { const uint32_t q = s_tbl[lab]; cand = (q == NONE16) ? NONE32 : q; }
                    else cand = gtbl[lab];
                    c1p = (cand != NONE32) ? (uint32_t)(k + 1) + cand : lab;
                }
                __syncwarp();
                if (act && (peers & lanes_gt(lane)) == 0) {
                    if (SMEMTBL) s_tbl[lab] = (uint16_t)s; else gtbl[lab] = (WT)s;
                }
                __syncwarp();
                if (act) {
                    uint32_t q;
                    if (SMEMTBL) { const uint32_t q16 = s_tbl[t + 1]; q = (q16 == NONE16) ? NONE32 : q16; }
                    else q = gtbl[t + 1];
                    const uint32_t c2p = (q != NONE32) ? (uint32_t)(k + 1) + q : t + 1;
                    store_row<IdxT, MODE>(ot, s, (IdxT)c1p, (IdxT)c2p, (IdxT)(k + 1 + s));
                }
            }
        }
        team_sync<CSIZE>();
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
constexpr int LARGE_THREADS_SINGLE = 512;
constexpr int LARGE_THREADS_CLUSTER = 1024;
constexpr int MID_THREADS = 256;
constexpr int MID_K_MAX = 12287;           // work arrays as uint16 in shared memory up to here
constexpr int WARP_K_MAX = 16384;          // warp-per-tree kernel for batches of trees up to here
static int warp_k_max() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("P2V_DECODE_WARP_KMAX");      // tuning knob (comparison runs)
        v = e ? atoi(e) : WARP_K_MAX;
        if (v > WARP_K_MAX) v = WARP_K_MAX;
        if (v < SMALL_K_MAX) v = SMALL_K_MAX;
    }
    return v;
}
static size_t mid_smem_bytes(int64_t k) {
    const size_t K = ((size_t)k + 31) & ~(size_t)31;
    const size_t tbl = ((size_t)k + 2 + 7) & ~(size_t)7;
    return (tbl + 6 * K + (K >> 5) + 32 + 2) * sizeof(uint16_t) + 16 * sizeof(int) + 16;
}

constexpr int SWEEP_WORKERS = 64;          // warps sharing the relabel sweep of one large tree (cluster mode)
static bool few_large_trees(int64_t B, int64_t k) { return (B * 8 <= sm_count()) && k >= 8192; }
static size_t large_ws_stride(int64_t B, int64_t k) {  // uint32 elements per team
    const size_t K = ((size_t)k + 31) & ~(size_t)31;
    size_t stride = large_base_stride(K);
    if (few_large_trees(B, k)) stride += (size_t)SWEEP_WORKERS * (K + 32);   // per-worker label tables
    return stride;
}
static int large_teams(int64_t B) {
    const int64_t cap = (int64_t)sm_count() * 4;
    return (int)(B < cap ? B : cap);
}

constexpr int BIG_WARP_K_MAX = 262143;      // 32 KB of free-slot mask per tree; label tables uint16 up to 65535 slots, uint32 above
// Resident trees per SM the workspace is sized for.  The kernel is latency-bound (two dependent label-table
// accesses per batch of rows), so more residents help until their tables fall out of the L2: measured 32 per SM
// at n = 32768 (+13 % over 16), 16 above (P2V_DECODE_BIG_PER_SM forces a value).
static int big_warp_ctas_per_sm(int64_t k) {
    static const int v = [] { const char* e = getenv("P2V_DECODE_BIG_PER_SM"); return e ? atoi(e) : 0; }();
    return v > 0 ? v : (k <= 32768 ? 32 : 16);
}
static bool big_warp_batch(int64_t B, int64_t k) { return k > warp_k_max() && k <= BIG_WARP_K_MAX && B * 4 > sm_count(); }
static size_t big_warp_table_stride(int64_t k) { return (((size_t)k + 31) & ~(size_t)31) + 32; }   // entries
static size_t big_warp_entry_bytes(int64_t k) { return k <= 65535 ? 2 : 4; }
static int64_t big_warp_ctas(int64_t B, int64_t k) {
    const int64_t cap = (int64_t)sm_count() * big_warp_ctas_per_sm(k);
    return B < cap ? B : cap;
}

size_t decode_workspace_bytes(int64_t B, int64_t k) {
    if (B > 0 && big_warp_batch(B, k)) return (size_t)big_warp_ctas(B, k) * big_warp_table_stride(k) * big_warp_entry_bytes(k);
    if (k <= MID_K_MAX || B <= 0) return 0;
    return (size_t)large_teams(B) * large_ws_stride(B, k) * sizeof(uint32_t);
}

template <typename IdxT, int MODE, int CSIZE, int THREADS, bool SMEMTBL, typename WT = uint32_t, bool SMEMARR = false>
static bool launch_large(const void* v, int64_t B, int64_t k, void* out, int64_t out_stride, int32_t* status,
                         void* ws, size_t stride, int teams, size_t smem, cudaStream_t stream) {
    auto kern = decode_large_kernel<IdxT, MODE, CSIZE, THREADS, SMEMTBL, WT, SMEMARR>;
    if (smem > 0 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (CSIZE > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (CSIZE == 0) {      // the whole cooperatively launched grid per tree
        const IdxT* vp0 = (const IdxT*)v;
        IdxT* op0 = (IdxT*)out;
        uint32_t* wp0 = (uint32_t*)ws;
        return launch_grid_team(kern, THREADS, smem, grid_team_ctas(2 * k, THREADS), stream, vp0, B, (int)k, op0, out_stride,
                                status, wp0, stride);
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(teams * (CSIZE ? CSIZE : 1)));
    cfg.blockDim = dim3(THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CSIZE ? CSIZE : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = (CSIZE > 1) ? 1 : 0;
    if (CSIZE > 1) {
        int max_clusters = 0;
        if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters < 1) {
            cudaGetLastError();
            return false;
        }
    }
    const IdxT* vp = (const IdxT*)v;
    IdxT* op = (IdxT*)out;
    uint32_t* wp = (uint32_t*)ws;
    const int kk = (int)k;
    if (cudaLaunchKernelEx(&cfg, kern, vp, B, kk, op, out_stride, status, wp, stride) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return true;
}

template <typename IdxT, int MODE>
static cudaError_t launch_decode_t(const void* v, int64_t B, int64_t k, void* out, int32_t* status,
                                   void* ws, size_t ws_bytes, cudaStream_t stream, int64_t out_stride) {
    if (out_stride <= 0) out_stride = k * ((MODE == MODE_PAIRS || MODE == MODE_ANC2) ? 2 : (MODE == MODE_ANCESTRY ? 3 : 4));
    // few large trees go to the merge kernel (all SMs on one tree); batches of mid-size trees keep
    // the warp-per-tree kernel, which does a fraction of the work per tree
    const bool warp_per_tree = k <= SMALL_K_MAX || (k <= warp_k_max() && B * 4 > sm_count());
    if (warp_per_tree) {
        const int K = (int)((k + 31) & ~31);
        static const int light_from = [] { const char* e = getenv("P2V_DECODE_LIGHT_FROM"); return e ? atoi(e) : SMALL_K_MAX + 1; }();
        static const int light_warps = [] { const char* e = getenv("P2V_DECODE_LIGHT_WARPS"); return e ? atoi(e) : 2; }();
        const int wpl = (k <= 1024) ? 1 : (k < light_from ? 2 : (k <= 4096 ? 4 : (k <= 8192 ? 8 : 16)));
        const int warps = (wpl <= 2) ? SMALL_WARPS : (k <= 4096 ? light_warps : 1);
        const size_t smem = (size_t)warps * decode_warp_scratch_u16(K, MODE, wpl) * sizeof(uint16_t);
        auto kern = decode_small_kernel<IdxT, MODE, 1>;
        if (wpl == 2) kern = decode_small_kernel<IdxT, MODE, 2>;
        else if (wpl == 4) kern = decode_small_kernel<IdxT, MODE, 4>;
        else if (wpl == 8) kern = decode_small_kernel<IdxT, MODE, 8>;
        else if (wpl == 16) kern = decode_small_kernel<IdxT, MODE, 16>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) per_sm = 1;
        int64_t want = (B + warps - 1) / warps;
        int64_t cap = (int64_t)sm_count() * per_sm;
        int grid = (int)(want < cap ? want : cap);
        kern<<<grid, warps * 32, smem, stream>>>((const IdxT*)v, B, (int)k, (IdxT*)out, out_stride, status, K);
    } else if (big_warp_batch(B, k)) {
        // batches of big trees: warp per tree, mask in shared memory, label table in the workspace
        const int K = (int)((k + 31) & ~31);
        const size_t smem = decode_big_scratch_u16(K) * sizeof(uint16_t);
        const size_t tstride = big_warp_table_stride(k);
        int64_t ctas = big_warp_ctas(B, k);
        if (!ws || ws_bytes < (size_t)ctas * tstride * big_warp_entry_bytes(k)) return cudaErrorInvalidValue;
        auto go = [&](auto kern, auto* tables) -> cudaError_t {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            int per_sm = 0;
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem);
            if (e != cudaSuccess) return e;
            if (per_sm < 1) per_sm = 1;
            if (per_sm > big_warp_ctas_per_sm(k)) per_sm = big_warp_ctas_per_sm(k);
            const int64_t cap = (int64_t)sm_count() * per_sm;
            if (ctas > cap) ctas = cap;
            kern<<<(int)ctas, 32, smem, stream>>>((const IdxT*)v, B, (int)k, (IdxT*)out, out_stride, status, K, tables, tstride);
            return cudaSuccess;
        };
        cudaError_t e = (k <= 65535) ? go(decode_big_warp_kernel<IdxT, MODE, uint16_t>, reinterpret_cast<uint16_t*>(ws))
                                     : go(decode_big_warp_kernel<IdxT, MODE, uint32_t>, reinterpret_cast<uint32_t*>(ws));
        if (e != cudaSuccess) return e;
    } else if (k <= MID_K_MAX) {
        // mid-size trees: the same merge kernel with uint16 work arrays in shared memory
        const size_t smem = mid_smem_bytes(k);
        const int64_t cap = (int64_t)sm_count() * 8;
        const int teams = (int)(B < cap ? B : cap);
        if (!launch_large<IdxT, MODE, 1, MID_THREADS, true, uint16_t, true>(v, B, k, out, out_stride, status, nullptr, 0,
                                                                            teams, smem, stream))
            return cudaErrorLaunchFailure;
    } else {
        if (k > LARGE_K_MAX) return cudaErrorNotSupported;
        const size_t stride = large_ws_stride(B, k);
        const int teams = large_teams(B);
        if (!ws || ws_bytes < (size_t)teams * stride * sizeof(uint32_t)) return cudaErrorInvalidValue;
        // last-row-per-label table in shared memory as uint16 while slots fit (k <= 65535)
        const bool smemtbl = (MODE != MODE_PAIRS) && (k <= 65535);
        const size_t smem = smemtbl ? ((size_t)k + 2 + 7) / 8 * 8 * sizeof(uint16_t) : 0;
        bool ok = false;
        const bool few_large = few_large_trees(B, k);
#define P2V_LARGE(CS_, TH_)                                                                            \
        (smemtbl ? launch_large<IdxT, MODE, CS_, TH_, (MODE != MODE_PAIRS)>(v, B, k, out, out_stride, status, ws, stride, teams, smem, stream) \
                 : launch_large<IdxT, MODE, CS_, TH_, false>(v, B, k, out, out_stride, status, ws, stride, teams, smem, stream))
        if (few_large && grid_team_trees(B, k)) ok = P2V_LARGE(0, LARGE_THREADS_CLUSTER);
        if (!ok && few_large) ok = P2V_LARGE(16, LARGE_THREADS_CLUSTER) || P2V_LARGE(8, LARGE_THREADS_CLUSTER);
        if (!ok) ok = P2V_LARGE(1, LARGE_THREADS_SINGLE);
#undef P2V_LARGE
        if (!ok) return cudaErrorLaunchFailure;
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
    case MODE_ANC2: return launch_decode_t<T, MODE_ANC2>(v, B, k, out, status, ws, ws_bytes, stream, out_stride);       \
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
