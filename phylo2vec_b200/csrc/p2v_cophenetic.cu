// Cophenetic (leaf-to-leaf path length) matrices for sm_100a.
//
// Replaces (reference, paths relative to /root/reference):
//   vector::graph::_cophenetic_distances   phylo2vec/src/vector/graph.rs:20-90
//   matrix::graph::cophenetic_distances    phylo2vec/src/matrix/graph.rs:23-26
//   matrix::base::parse_matrix             phylo2vec/src/matrix/base.rs:108-121
//
// The reference fills a (2k+1)^2 scratch matrix with ~1.5 k^2 x 4 scattered
// stores and slices the leaf block out of it.  Here the n x n leaf block is
// written once, row by row, with 128-bit stores, and nothing else of size n^2
// is touched:
//   D[i][j] = depth(i) + depth(j) - 2 depth(LCA(i,j)),   D[i][i] = 0,
// "unrooted" zeroes the root edge that leads to node 2k-1 (graph.rs:46-50 with
// the write order at :72-79), k = 0 and k = 1 are the special cases at :28-42.
//
// Pipeline per tree (k >= 2):
//   1. decode v -> ancestry (p2v_decode.cu)
//   2. tables_kernel: parent pointers; root-path sums (topological depth and
//      double-double branch-length depth) by pointer jumping; Euler tour of the
//      tree ranked by pointer jumping -> DFS position of every leaf and, for
//      every DFS position r, the internal node sep[r] that separates leaves
//      r-1 and r.  LCA(a,b) for DFS positions a<b is the shallowest sep in (a,b].
//      Per DFS block of R positions: prefix / suffix minima and the block minimum.
//   3. rows_kernel: one CTA = (DFS block of R rows) x (chunk of columns in
//      label order).  The range minimum splits into a per-row part (inside the
//      row's block), a per-column part (inside the column's block, combined once
//      per CTA with the minimum over the whole blocks in between), so the
//      streamed inner loop is a compare, a select and an add per element.
//      Columns whose leaf lies inside the row block are patched afterwards from
//      the running minima of that block.
// fp64 branch lengths: depths are kept as double-double and the two candidates
// (row part wins / column part wins) are pre-combined so that each element costs
// three DADD and is rounded once (DESIGN.md "fp64 parity").
#include "p2v_common.cuh"
#include "p2v_internal.h"

namespace p2v {

constexpr int TAB_THREADS = 512;
constexpr int ROW_THREADS = 256;
constexpr int ROW_R = 64;                  // DFS positions (matrix rows) per CTA
constexpr int ROW_CPT = 4;                 // columns per thread (two 16-byte stores per row)
constexpr int ROW_CHUNK = ROW_THREADS * ROW_CPT;   // 1024 columns per CTA
constexpr int KEY_INF = 0x3FFFFFFF;

struct CophLayout {
    // byte offsets inside one tree's table area
    size_t anc, jA, jDI, jDH, jDL, tn, tc, pos, leaf_at, sepk, sephi, seplo, prek, sufk, prehi,
        prelo, sufhi, suflo, blkk, blkhi, blklo, total;
    int n, N, A, nb;
};

__host__ __device__ inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

__host__ __device__ inline CophLayout coph_layout(int64_t k) {
    CophLayout L;
    L.n = (int)(k + 1); L.N = (int)(2 * k + 1); L.A = (int)(4 * k);
    L.nb = (L.n + ROW_R - 1) / ROW_R;
    size_t o = 0;
    L.anc = o; o = al16(o + 4 * 3 * (size_t)k);
    L.jA = o; o = al16(o + 4 * 2 * (size_t)L.N);
    L.jDI = o; o = al16(o + 4 * 2 * (size_t)L.N);
    L.jDH = o; o = al16(o + 8 * 2 * (size_t)L.N);
    L.jDL = o; o = al16(o + 8 * 2 * (size_t)L.N);
    L.tn = o; o = al16(o + 4 * 2 * (size_t)L.A);
    L.tc = o; o = al16(o + 4 * 2 * (size_t)L.A);
    L.pos = o; o = al16(o + 4 * (size_t)L.n);
    L.leaf_at = o; o = al16(o + 4 * (size_t)L.n);
    L.sepk = o; o = al16(o + 4 * (size_t)(L.n + 1));
    L.sephi = o; o = al16(o + 8 * (size_t)(L.n + 1));
    L.seplo = o; o = al16(o + 8 * (size_t)(L.n + 1));
    L.prek = o; o = al16(o + 4 * (size_t)L.n);
    L.sufk = o; o = al16(o + 4 * (size_t)L.n);
    L.prehi = o; o = al16(o + 8 * (size_t)L.n);
    L.prelo = o; o = al16(o + 8 * (size_t)L.n);
    L.sufhi = o; o = al16(o + 8 * (size_t)L.n);
    L.suflo = o; o = al16(o + 8 * (size_t)L.n);
    L.blkk = o; o = al16(o + 4 * (size_t)L.nb);
    L.blkhi = o; o = al16(o + 8 * (size_t)L.nb);
    L.blklo = o; o = al16(o + 8 * (size_t)L.nb);
    L.total = al16(o + 64);
    return L;
}

// ---- input adapters ----------------------------------------------------------------
__global__ void narrow_v_kernel(const long long* __restrict__ v, int64_t count, int* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const long long x = v[i];
        out[i] = (x < 0 || x > 0x7FFFFFFFLL) ? -1 : (int)x;  // -1 fails check_v downstream
    }
}

// parse_matrix (matrix/base.rs:108-121): v = m[:,0] as usize (truncating, saturating cast)
__global__ void matrix_v_kernel(const double* __restrict__ m, int64_t count, int* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double x = m[3 * i];
        int r;
        if (!(x > 0.0)) r = 0;                    // negatives and NaN saturate to 0
        else if (x >= 2147483647.0) r = 0x7FFFFFFF;
        else r = (int)x;
        out[i] = r;
    }
}

// ---- k < 2 special cases (graph.rs:28-42) ---------------------------------------------
__global__ void tiny_sum_kernel(const double* __restrict__ bl, int64_t bl_tree_stride, int64_t B,
                                double* __restrict__ sums) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (int64_t)gridDim.x * blockDim.x)
        sums[2 * t] = bl ? bl[t * bl_tree_stride] + bl[t * bl_tree_stride + 1] : 2.0;
}

__global__ void tiny_kernel(const double* __restrict__ sums, int64_t B, int k, int64_t row_begin,
                            int64_t row_end, double* __restrict__ out) {
    const int n = k + 1;
    const int64_t rows = row_end - row_begin;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < B; t += (int64_t)gridDim.x * blockDim.x) {
        const double s = (k == 1) ? sums[2 * t] : 0.0;
        for (int64_t r = row_begin; r < row_end; ++r)
            for (int c = 0; c < n; ++c) out[(t * rows + (r - row_begin)) * n + c] = (r == c) ? 0.0 : s;
    }
}

// ---- 2. per-tree tables ------------------------------------------------------------------
__global__ void __launch_bounds__(TAB_THREADS)
tables_kernel(unsigned char* __restrict__ tab, size_t tab_stride, const double* __restrict__ bl,
              int bl_stride, int64_t bl_tree_stride, int64_t B, int k, int unrooted,
              const int32_t* __restrict__ status) {
    const CophLayout L = coph_layout(k);
    const int tid = threadIdx.x;
    const int n = L.n, N = L.N, A = L.A, root = 2 * k;
    for (int64_t tree = blockIdx.x; tree < B; tree += gridDim.x) {
        if (status && status[tree] != 0) continue;   // uniform per CTA
        unsigned char* T = tab + (size_t)tree * tab_stride;
        const int* anc = reinterpret_cast<const int*>(T + L.anc);
        int* jA = reinterpret_cast<int*>(T + L.jA);
        int* jDI = reinterpret_cast<int*>(T + L.jDI);
        double* jDH = reinterpret_cast<double*>(T + L.jDH);
        double* jDL = reinterpret_cast<double*>(T + L.jDL);
        int* tn = reinterpret_cast<int*>(T + L.tn);
        int* tc = reinterpret_cast<int*>(T + L.tc);
        int* pos = reinterpret_cast<int*>(T + L.pos);
        int* leaf_at = reinterpret_cast<int*>(T + L.leaf_at);
        int* sepk = reinterpret_cast<int*>(T + L.sepk);
        double* sephi = reinterpret_cast<double*>(T + L.sephi);
        double* seplo = reinterpret_cast<double*>(T + L.seplo);
        const double* blt = bl ? bl + tree * bl_tree_stride : nullptr;

        // parents and edge weights (row r of bls: lengths above ancestry[r][0], [r][1];
        // graph.rs:57-58).  jA/jDI/jDH/jDL buffer 0 = initial state of the jumping.
        for (int r = tid; r < k; r += TAB_THREADS) {
            const int c1 = anc[3 * r], c2 = anc[3 * r + 1], p = k + 1 + r;
            double b1 = 1.0, b2 = 1.0;
            if (blt) { b1 = blt[(size_t)r * bl_stride]; b2 = blt[(size_t)r * bl_stride + 1]; }
            int w1 = 1, w2 = 1;
            if (unrooted) {
                if (c1 == 2 * k - 1) { w1 = 0; b1 = 0.0; }
                if (c2 == 2 * k - 1) { w2 = 0; b2 = 0.0; }
            }
            jA[c1] = p; jA[c2] = p;
            jDI[c1] = w1; jDI[c2] = w2;
            jDH[c1] = b1; jDH[c2] = b2;
            jDL[c1] = 0.0; jDL[c2] = 0.0;
        }
        if (tid == 0) { jA[root] = -1; jDI[root] = 0; jDH[root] = 0.0; jDL[root] = 0.0; }
        __syncthreads();
        // Euler tour arcs: down(c) = 2c, up(c) = 2c+1 for every non-root node c.
        for (int c = tid; c < 2 * k; c += TAB_THREADS) {
            const int p = jA[c];
            int nd, wd;
            if (c <= k) { nd = 2 * c + 1; wd = 1; }                     // leaf: turn around
            else { nd = 2 * anc[3 * (c - k - 1)]; wd = 0; }              // internal: into child 0
            int nu;
            const int* prow = anc + 3 * (p - k - 1);
            if (prow[0] == c) nu = 2 * prow[1];                          // then the sibling (child 1)
            else nu = (p == root) ? -1 : 2 * p + 1;                      // or further up
            tn[2 * c] = nd; tc[2 * c] = wd;
            tn[2 * c + 1] = nu; tc[2 * c + 1] = 0;
        }
        __syncthreads();
        // root-path sums by pointer jumping; an even number of rounds leaves the result in buffer 0
        {
            int cur = 0;
            for (int round = 0;; ++round) {
                const int* a_in = jA + cur * N; int* a_out = jA + (cur ^ 1) * N;
                const int* i_in = jDI + cur * N; int* i_out = jDI + (cur ^ 1) * N;
                const double* h_in = jDH + cur * N; double* h_out = jDH + (cur ^ 1) * N;
                const double* l_in = jDL + cur * N; double* l_out = jDL + (cur ^ 1) * N;
                int live = 0;
                for (int u = tid; u < N; u += TAB_THREADS) {
                    const int a = a_in[u];
                    int di = i_in[u];
                    dd d = {h_in[u], l_in[u]};
                    int na = -1;
                    if (a >= 0) {
                        di += i_in[a];
                        d = dd_add(d, dd{h_in[a], l_in[a]});
                        na = a_in[a];
                        live |= (na >= 0);
                    }
                    a_out[u] = na; i_out[u] = di; h_out[u] = d.hi; l_out[u] = d.lo;
                }
                cur ^= 1;
                const int any = __syncthreads_or(live);
                if (!any && cur == 0) break;
            }
        }
        // rank the tour: tc[a] becomes the number of leaf arcs from a to the end (inclusive)
        {
            int cur = 0;
            for (int round = 0;; ++round) {
                const int* n_in = tn + cur * A; int* n_out = tn + (cur ^ 1) * A;
                const int* c_in = tc + cur * A; int* c_out = tc + (cur ^ 1) * A;
                int live = 0;
                for (int a = tid; a < A; a += TAB_THREADS) {
                    const int nx = n_in[a];
                    int c = c_in[a];
                    int nn = -1;
                    if (nx >= 0) {
                        c += c_in[nx];
                        nn = n_in[nx];
                        live |= (nn >= 0);
                    }
                    n_out[a] = nn; c_out[a] = c;
                }
                cur ^= 1;
                const int any = __syncthreads_or(live);
                if (!any && cur == 0) break;
            }
        }
        // DFS positions of the leaves; separators by position
        for (int c = tid; c <= k; c += TAB_THREADS) {
            const int p = n - tc[2 * c];
            pos[c] = p;
            leaf_at[p] = c;
        }
        if (tid == 0) { sepk[0] = KEY_INF; sephi[0] = 0.0; seplo[0] = 0.0; sepk[n] = KEY_INF; sephi[n] = 0.0; seplo[n] = 0.0; }
        for (int u = k + 1 + tid; u <= 2 * k; u += TAB_THREADS) {
            const int c1 = anc[3 * (u - k - 1) + 1];      // child 1: its first leaf opens the right part
            const int r = n - tc[2 * c1];
            sepk[r] = jDI[u]; sephi[r] = jDH[u]; seplo[r] = jDL[u];
        }
        __syncthreads();
        // per-block prefix / suffix minima over the separators (by DFS position)
        {
            int* prek = reinterpret_cast<int*>(T + L.prek);
            int* sufk = reinterpret_cast<int*>(T + L.sufk);
            double* prehi = reinterpret_cast<double*>(T + L.prehi);
            double* prelo = reinterpret_cast<double*>(T + L.prelo);
            double* sufhi = reinterpret_cast<double*>(T + L.sufhi);
            double* suflo = reinterpret_cast<double*>(T + L.suflo);
            int* blkk = reinterpret_cast<int*>(T + L.blkk);
            double* blkhi = reinterpret_cast<double*>(T + L.blkhi);
            double* blklo = reinterpret_cast<double*>(T + L.blklo);
            for (int b = tid; b < L.nb; b += TAB_THREADS) {
                const int lo = b * ROW_R, hi = min(n, lo + ROW_R);
                int bk = KEY_INF; double bh = 0.0, bl_ = 0.0;
                for (int p = lo; p < hi; ++p) {          // pre[p] = min sep[lo..p]
                    const int kk = sepk[p];
                    if (kk < bk) { bk = kk; bh = sephi[p]; bl_ = seplo[p]; }
                    prek[p] = bk; prehi[p] = bh; prelo[p] = bl_;
                }
                blkk[b] = bk; blkhi[b] = bh; blklo[b] = bl_;
                bk = KEY_INF; bh = 0.0; bl_ = 0.0;
                for (int p = hi - 1; p >= lo; --p) {     // suf[p] = min sep[p+1..hi-1]
                    sufk[p] = bk; sufhi[p] = bh; suflo[p] = bl_;
                    const int kk = sepk[p];
                    if (kk < bk) { bk = kk; bh = sephi[p]; bl_ = seplo[p]; }
                }
            }
        }
        __syncthreads();
    }
}

// ---- 3. rows --------------------------------------------------------------------------------
__device__ __forceinline__ double int_to_f64(int v) {
    // exact for |v| < 2^31 : one DADD instead of an I2F.F64 on the slow pipe
    return __hiloint2double(0x43300000, (int)(0x80000000u ^ (unsigned)v)) - 4503601774854144.0;  // 2^52 + 2^31
}

template <bool VEC2>
__device__ __forceinline__ void store2(double* p, double a, double b) {
    if (VEC2) {
        double2 v; v.x = a; v.y = b;
        __stcs(reinterpret_cast<double2*>(p), v);
    } else {
        __stcs(p, a); __stcs(p + 1, b);
    }
}

struct RowShared {
    int label[ROW_R];
    int bk[2][ROW_R];        // [0]: prefix key (columns left of the tile)  [1]: suffix key (columns right)
    int di[ROW_R];           // topological depth of the row leaf
    int fi[2][ROW_R];        // di - 2 * row-part depth                       (topological)
    double dhi[ROW_R], dlo[ROW_R];
    double fhi[2][ROW_R], flo[2][ROW_R];
    int sk[ROW_R + 1];       // separators of the tile (for the in-tile patch)
    double shi[ROW_R + 1], slo[ROW_R + 1];
};

template <bool TOPO, bool VEC2>
__global__ void __launch_bounds__(ROW_THREADS)
rows_kernel(const unsigned char* __restrict__ tab, size_t tab_stride, int64_t B, int k,
            int64_t row_begin, int64_t row_end, double* __restrict__ out,
            const int32_t* __restrict__ status, int chunks_per_tree) {
    extern __shared__ __align__(16) unsigned char dyn[];
    __shared__ RowShared S;
    const CophLayout L = coph_layout(k);
    const int n = L.n, nb = L.nb;
    int* midk = reinterpret_cast<int*>(dyn);          // nb
    int* mida = midk + nb;                            // nb
    unsigned long long* scanR = reinterpret_cast<unsigned long long*>(dyn + (((size_t)8 * nb + 15) & ~(size_t)15));
    unsigned long long* scanL = scanR + ROW_THREADS;
    const int tid = threadIdx.x;
    const int64_t rows = row_end - row_begin;

    const int64_t tiles_per_tree = (int64_t)nb * chunks_per_tree;
    for (int64_t work = blockIdx.x; work < B * tiles_per_tree; work += gridDim.x) {
        const int64_t tree = work / tiles_per_tree;
        const int rem = (int)(work % tiles_per_tree);
        const int I = rem / chunks_per_tree;
        const int chunk = rem % chunks_per_tree;
        if (status && status[tree] != 0) continue;
        const unsigned char* T = tab + (size_t)tree * tab_stride;
        const int* jDI = reinterpret_cast<const int*>(T + L.jDI);
        const double* jDH = reinterpret_cast<const double*>(T + L.jDH);
        const double* jDL = reinterpret_cast<const double*>(T + L.jDL);
        const int* pos = reinterpret_cast<const int*>(T + L.pos);
        const int* leaf_at = reinterpret_cast<const int*>(T + L.leaf_at);
        const int* sepk = reinterpret_cast<const int*>(T + L.sepk);
        const double* sephi = reinterpret_cast<const double*>(T + L.sephi);
        const double* seplo = reinterpret_cast<const double*>(T + L.seplo);
        const int* prek = reinterpret_cast<const int*>(T + L.prek);
        const int* sufk = reinterpret_cast<const int*>(T + L.sufk);
        const double* prehi = reinterpret_cast<const double*>(T + L.prehi);
        const double* prelo = reinterpret_cast<const double*>(T + L.prelo);
        const double* sufhi = reinterpret_cast<const double*>(T + L.sufhi);
        const double* suflo = reinterpret_cast<const double*>(T + L.suflo);
        const int* blkk = reinterpret_cast<const int*>(T + L.blkk);
        const double* blkhi = reinterpret_cast<const double*>(T + L.blkhi);
        const double* blklo = reinterpret_cast<const double*>(T + L.blklo);
        const int lo = I * ROW_R, hi = min(n, lo + ROW_R), nrow = hi - lo;
        __syncthreads();   // previous iteration's readers of S / dyn are done

        // ---- row side of the tile -------------------------------------------------------
        bool any_row = false;
        if (tid < nrow) {
            const int q = lo + tid;
            const int i = leaf_at[q];
            S.label[tid] = i;
            const int pk = prek[q], sk = sufk[q];
            S.bk[0][tid] = pk; S.bk[1][tid] = sk;
            const int di = jDI[i];
            S.di[tid] = di;
            if (TOPO) {
                S.fi[0][tid] = di - 2 * pk;     // key == topological depth of the candidate
                S.fi[1][tid] = di - 2 * sk;
            } else {
                const dd d = {jDH[i], jDL[i]};
                S.dhi[tid] = d.hi; S.dlo[tid] = d.lo;
                const dd f0 = dd_add(d, dd_neg2(dd{prehi[q], prelo[q]}));
                const dd f1 = dd_add(d, dd_neg2(dd{sufhi[q], suflo[q]}));
                S.fhi[0][tid] = f0.hi; S.flo[0][tid] = f0.lo;
                S.fhi[1][tid] = f1.hi; S.flo[1][tid] = f1.lo;
            }
            any_row = (i >= row_begin && i < row_end);
        }
        if (tid <= nrow) {
            const int r = lo + tid;
            S.sk[tid] = (tid == 0) ? KEY_INF : sepk[min(r, n)];
            if (!TOPO) { S.shi[tid] = sephi[min(r, n)]; S.slo[tid] = seplo[min(r, n)]; }
        }
        if (!__syncthreads_or(any_row)) continue;    // none of this tile's rows is in [row_begin,row_end)

        // ---- minima over the whole blocks between the tile and every other block --------
        // midk[J] / mida[J] = min / argmin of blkk over the blocks strictly between I and J.
        // Each thread owns a contiguous segment of blocks; segment minima are combined with
        // a block-wide prefix (right of I) and suffix (left of I) min-scan on packed
        // (key << 32 | block) words, then every thread walks its own segment.
        {
            const int SEG = (nb + ROW_THREADS - 1) / ROW_THREADS;
            const int s0 = tid * SEG, s1 = min(nb, s0 + SEG);
            const unsigned long long INF64 = ((unsigned long long)KEY_INF << 32);
            unsigned long long pr = INF64, pl = INF64;      // this segment's minimum right / left of I
            for (int J = s0; J < s1; ++J) {
                const unsigned long long e = ((unsigned long long)(unsigned)blkk[J] << 32) | (unsigned)J;
                if (J > I) pr = min(pr, e);
                if (J < I) pl = min(pl, e);
            }
            scanR[tid] = pr; scanL[tid] = pl;
            __syncthreads();
            for (int d = 1; d < ROW_THREADS; d <<= 1) {     // inclusive prefix-min (R) / suffix-min (L)
                unsigned long long a = scanR[tid], b = scanL[tid];
                if (tid >= d) a = min(a, scanR[tid - d]);
                if (tid + d < ROW_THREADS) b = min(b, scanL[tid + d]);
                __syncthreads();
                scanR[tid] = a; scanL[tid] = b;
                __syncthreads();
            }
            unsigned long long run = (tid > 0) ? scanR[tid - 1] : INF64;      // blocks right of I, left of my segment
            for (int J = s0; J < s1; ++J) {
                if (J > I) {
                    midk[J] = (int)(run >> 32); mida[J] = (int)(run & 0xFFFFFFFFu);
                    run = min(run, ((unsigned long long)(unsigned)blkk[J] << 32) | (unsigned)J);
                }
            }
            run = (tid + 1 < ROW_THREADS) ? scanL[tid + 1] : INF64;            // blocks left of I, right of my segment
            for (int J = s1 - 1; J >= s0; --J) {
                if (J < I) {
                    midk[J] = (int)(run >> 32); mida[J] = (int)(run & 0xFFFFFFFFu);
                    run = min(run, ((unsigned long long)(unsigned)blkk[J] << 32) | (unsigned)J);
                }
                if (J == I) { midk[J] = KEY_INF; mida[J] = 0; }
            }
        }
        __syncthreads();

        // ---- column side: 4 columns per thread (two adjacent pairs) ----------------------
        const int cbase = chunk * ROW_CHUNK;
        int col[ROW_CPT];
        int ak[ROW_CPT], side[ROW_CPT];
        int ei[ROW_CPT], dji[ROW_CPT];
        double ehi[ROW_CPT], elo[ROW_CPT], djh[ROW_CPT], djl[ROW_CPT];
#pragma unroll
        for (int c = 0; c < ROW_CPT; ++c) {
            const int j = cbase + (c >> 1) * (ROW_THREADS * 2) + 2 * tid + (c & 1);
            col[c] = j;
            ak[c] = KEY_INF; side[c] = 0; ei[c] = 0; dji[c] = 0;
            ehi[c] = elo[c] = djh[c] = djl[c] = 0.0;
            if (j < n) {
                const int p = pos[j];
                const int J = p / ROW_R;
                const bool right = J > I;
                side[c] = right ? 1 : 0;
                int ck = right ? prek[p] : sufk[p];
                dd cv = {0.0, 0.0};
                if (!TOPO) cv = right ? dd{prehi[p], prelo[p]} : dd{sufhi[p], suflo[p]};
                const int mk = midk[J];
                if (mk < ck) {
                    ck = mk;
                    if (!TOPO) { const int a = mida[J]; cv = dd{blkhi[a], blklo[a]}; }
                }
                if (J == I) ck = KEY_INF;      // patched after the streamed part
                ak[c] = ck;
                dji[c] = jDI[j];
                if (TOPO) {
                    ei[c] = dji[c] - 2 * ck;
                } else {
                    const dd dj = {jDH[j], jDL[j]};
                    djh[c] = dj.hi; djl[c] = dj.lo;
                    const dd e = dd_add(dj, dd_neg2(cv));
                    ehi[c] = e.hi; elo[c] = e.lo;
                }
            }
        }

        // ---- streamed part: R rows x 4 columns per thread ----------------------------------
        double* obase = out + (size_t)tree * rows * n;
        for (int r = 0; r < nrow; ++r) {
            const int i = S.label[r];
            if (i < row_begin || i >= row_end) continue;
            double* orow = obase + (size_t)(i - row_begin) * n;
            double val[ROW_CPT];
            if (TOPO) {
                const int b0 = S.bk[0][r], b1 = S.bk[1][r], f0 = S.fi[0][r], f1 = S.fi[1][r], di = S.di[r];
#pragma unroll
                for (int c = 0; c < ROW_CPT; ++c) {
                    const int rk = side[c] ? b1 : b0;
                    const int f = side[c] ? f1 : f0;
                    const int x = (ak[c] < rk) ? (di + ei[c]) : (f + dji[c]);
                    val[c] = int_to_f64(x);
                }
            } else {
                const int b0 = S.bk[0][r], b1 = S.bk[1][r];
                const double dh = S.dhi[r], dl = S.dlo[r];
                const double f0h = S.fhi[0][r], f0l = S.flo[0][r], f1h = S.fhi[1][r], f1l = S.flo[1][r];
#pragma unroll
                for (int c = 0; c < ROW_CPT; ++c) {
                    const int rk = side[c] ? b1 : b0;
                    const double fh = side[c] ? f1h : f0h;
                    const double fl = side[c] ? f1l : f0l;
                    const bool colwin = ak[c] < rk;
                    const double xh = colwin ? dh : fh, xl = colwin ? dl : fl;
                    const double yh = colwin ? ehi[c] : djh[c], yl = colwin ? elo[c] : djl[c];
                    val[c] = __dadd_rn(__dadd_rn(xh, yh), __dadd_rn(xl, yl));
                }
            }
#pragma unroll
            for (int c = 0; c < ROW_CPT; c += 2) {
                if (col[c] + 1 < n) store2<VEC2>(orow + col[c], val[c], val[c + 1]);
                else if (col[c] < n) __stcs(orow + col[c], val[c]);
            }
        }
        __syncthreads();   // streamed stores of this CTA are ordered before the patch below

        // ---- patch: columns whose leaf sits inside the tile (incl. the zero diagonal) ------
        // thread a walks b = a+1.. with the running minimum of sep(a, b]
        if (tid < nrow) {
            const int a = tid;
            const int ia = S.label[a];
            const bool ia_row = ia >= row_begin && ia < row_end;
            const bool ia_col = ia >= cbase && ia < cbase + ROW_CHUNK;
            if (ia_row && ia_col) obase[(size_t)(ia - row_begin) * n + ia] = 0.0;
            int bk = KEY_INF; dd bv = {0.0, 0.0};
            for (int b = a + 1; b < nrow; ++b) {
                if (S.sk[b] < bk) { bk = S.sk[b]; if (!TOPO) bv = dd{S.shi[b], S.slo[b]}; }
                const int ib = S.label[b];
                const bool ib_row = ib >= row_begin && ib < row_end;
                const bool ib_col = ib >= cbase && ib < cbase + ROW_CHUNK;
                if (!((ia_row && ib_col) || (ib_row && ia_col))) continue;
                double d;
                if (TOPO) {
                    d = int_to_f64(S.di[a] + S.di[b] - 2 * bk);
                } else {
                    // (d_a - lca) + (d_b - lca), each difference in double-double
                    const dd m = dd{-bv.hi, -bv.lo};
                    const dd xa = dd_add(dd{S.dhi[a], S.dlo[a]}, m);
                    const dd xb = dd_add(dd{S.dhi[b], S.dlo[b]}, m);
                    d = __dadd_rn(__dadd_rn(xa.hi, xb.hi), __dadd_rn(xa.lo, xb.lo));
                }
                if (ia_row && ib_col) obase[(size_t)(ia - row_begin) * n + ib] = d;
                if (ib_row && ia_col) obase[(size_t)(ib - row_begin) * n + ia] = d;
            }
        }
    }
}

__global__ void fill_f64_kernel(double* __restrict__ out, int64_t count, double value) {
    const int64_t n2 = count >> 1;
    double2 v; v.x = value; v.y = value;
    double2* o2 = reinterpret_cast<double2*>(out);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
        __stcs(o2 + i, v);
    if ((count & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[count - 1] = value;
}

// --------------------------------------------------------------------------------------------
// workspace: [ status scratch (B int32) | int32 v (B*k) | decode workspace | tables (B * total) ]
struct CophWs { size_t status, v32, dec, tab, total; };
static CophWs coph_ws(int64_t B, int64_t k) {
    CophWs w;
    size_t o = 0;
    w.status = o; o = al16(o + 4 * (size_t)B);
    w.v32 = o; o = al16(o + 4 * (size_t)B * (size_t)(k > 0 ? k : 1));
    w.dec = o; o = al16(o + decode_workspace_bytes(B, k));
    w.tab = o; o = al16(o + (size_t)B * (k >= 2 ? coph_layout(k).total : 16));
    w.total = o;
    return w;
}

size_t cophenetic_workspace_bytes(int64_t B, int64_t k) {
    if (B <= 0) return 0;
    return coph_ws(B, k).total;
}

// Phase 1+2: decode and build the per-tree tables inside the workspace.
cudaError_t launch_cophenetic_prepare(const void* v_or_matrix, int idx_bytes, const double* bls, int64_t B,
                                      int64_t k, int unrooted, int32_t* status, void* ws, size_t ws_bytes,
                                      cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (k > LARGE_K_MAX) return cudaErrorNotSupported;
    const CophWs W = coph_ws(B, k);
    if (!ws || ws_bytes < W.total) return cudaErrorInvalidValue;
    unsigned char* wsb = static_cast<unsigned char*>(ws);
    int32_t* st = reinterpret_cast<int32_t*>(wsb + W.status);
    int* v32 = reinterpret_cast<int*>(wsb + W.v32);
    const int grid1d = sm_count() * 8;
    cudaError_t e;

    // branch lengths: bls (B,k,2) or columns 1..2 of the (B,k,3) matrix
    const double* bl = bls;
    int bl_stride = 2;
    int64_t bl_tree_stride = 2 * k;
    const void* vin = v_or_matrix;
    int vin_bytes = idx_bytes;
    if (idx_bytes == 0) {  // matrix input
        const double* m = static_cast<const double*>(v_or_matrix);
        bl = m + 1; bl_stride = 3; bl_tree_stride = 3 * k;
        if (k > 0) {
            matrix_v_kernel<<<grid1d, 256, 0, stream>>>(m, B * k, v32);
            count_launch();
        }
        vin = v32; vin_bytes = 4;
    } else if (idx_bytes == 8) {
        if (k > 0) {
            narrow_v_kernel<<<grid1d, 256, 0, stream>>>(static_cast<const long long*>(v_or_matrix), B * k, v32);
            count_launch();
        }
        vin = v32; vin_bytes = 4;
    } else if (idx_bytes != 4) {
        return cudaErrorInvalidValue;
    }
    if (k < 2) {
        if (k == 1) e = launch_check_v(vin, vin_bytes, B, k, st, stream);
        else e = cudaMemsetAsync(st, 0, sizeof(int32_t) * B, stream);
        if (e != cudaSuccess) return e;
        // k == 1: the rows phase needs bl0 + bl1 per tree; keep it in the table area (16 B per tree)
        if (k == 1) {
            double* sums = reinterpret_cast<double*>(wsb + W.tab);
            tiny_sum_kernel<<<(int)((B + 255) / 256 < 1024 ? (B + 255) / 256 : 1024), 256, 0, stream>>>(
                bl, bl_tree_stride, B, sums);
            count_launch();
        }
    } else {
        const CophLayout L = coph_layout(k);
        unsigned char* tab = wsb + W.tab;
        // decode straight into every tree's table area (per-tree output stride)
        e = launch_decode(MODE_ANCESTRY, vin, vin_bytes, B, k, tab + L.anc, st, wsb + W.dec, W.tab - W.dec,
                          stream, (int64_t)(L.total / sizeof(int)));
        if (e != cudaSuccess) return e;
        const int64_t cap = (int64_t)sm_count() * 4;
        const int grid = (int)(B < cap ? B : cap);
        tables_kernel<<<grid, TAB_THREADS, 0, stream>>>(tab, L.total, bl, bl_stride, bl_tree_stride, B,
                                                       (int)k, unrooted, st);
        count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    if (status) {
        e = cudaMemcpyAsync(status, st, sizeof(int32_t) * B, cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

// Phase 3: rows [row_begin,row_end) of every prepared tree.
cudaError_t launch_cophenetic_rows(const void* ws, size_t ws_bytes, int64_t B, int64_t k, int weighted,
                                   int64_t row_begin, int64_t row_end, double* out, cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    const int64_t n = k + 1;
    if (row_begin < 0 || row_end > n || row_begin > row_end) return cudaErrorInvalidValue;
    if (row_begin == row_end) return cudaSuccess;
    const CophWs W = coph_ws(B, k);
    if (!ws || ws_bytes < W.total) return cudaErrorInvalidValue;
    const unsigned char* wsb = static_cast<const unsigned char*>(ws);
    const int32_t* st = reinterpret_cast<const int32_t*>(wsb + W.status);
    cudaError_t e;
    if (k < 2) {
        tiny_kernel<<<(int)((B + 255) / 256 < 1024 ? (B + 255) / 256 : 1024), 256, 0, stream>>>(
            reinterpret_cast<const double*>(wsb + W.tab), B, (int)k, row_begin, row_end, out);
        count_launch();
        return cudaGetLastError();
    }
    const CophLayout L = coph_layout(k);
    const unsigned char* tab = wsb + W.tab;
    const int chunks = (int)((n + ROW_CHUNK - 1) / ROW_CHUNK);
    const int64_t work = B * (int64_t)L.nb * chunks;
    const size_t smem = (((size_t)8 * L.nb + 15) & ~(size_t)15) + 2 * ROW_THREADS * sizeof(unsigned long long);
    const bool topo = !weighted;
    const bool vec2 = (n % 2 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    const int64_t cap = (int64_t)sm_count() * 64;
    const int grid = (int)(work < cap ? work : cap);
#define P2V_ROWS(T_, V_)                                                                              \
    do {                                                                                              \
        auto kern = rows_kernel<T_, V_>;                                                              \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
        if (e != cudaSuccess) return e;                                                               \
        kern<<<grid, ROW_THREADS, smem, stream>>>(tab, L.total, B, (int)k, row_begin, row_end, out, st, chunks); \
    } while (0)
    if (topo) { if (vec2) P2V_ROWS(true, true); else P2V_ROWS(true, false); }
    else { if (vec2) P2V_ROWS(false, true); else P2V_ROWS(false, false); }
#undef P2V_ROWS
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_cophenetic(const void* v_or_matrix, int idx_bytes, const double* bls, int64_t B,
                              int64_t k, int unrooted, int64_t row_begin, int64_t row_end,
                              double* out, int32_t* status, void* ws, size_t ws_bytes,
                              cudaStream_t stream) {
    if (B <= 0) return cudaSuccess;
    if (row_begin < 0 || row_end > k + 1 || row_begin > row_end) return cudaErrorInvalidValue;
    cudaError_t e = launch_cophenetic_prepare(v_or_matrix, idx_bytes, bls, B, k, unrooted, status, ws, ws_bytes, stream);
    if (e != cudaSuccess) return e;
    const int weighted = (idx_bytes == 0) || (bls != nullptr);
    return launch_cophenetic_rows(ws, ws_bytes, B, k, weighted, row_begin, row_end, out, stream);
}

cudaError_t launch_fill_f64(double* out, int64_t count, double value, cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    fill_f64_kernel<<<sm_count() * 8, 256, 0, stream>>>(out, count, value);
    count_launch();
    return cudaGetLastError();
}

}  // namespace p2v
