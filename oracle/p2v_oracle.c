/*
 * p2v_oracle.c -- CPU restatement of the phylo2vec hot path (TEST INFRASTRUCTURE).
 *
 * This file is the parity oracle and the CPU baseline ("port") for the B200
 * kernels in phylo2vec_b200/csrc.  It is NOT part of the product path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it.  The product path never falls back to it.
 *
 * Parity status: PINNED.  Every function below follows the cited lines of the
 * reference (paths relative to /root/reference) and is checked against the
 * reference's own golden vectors in tests/golden/reference_goldens.json
 * (tests/test_oracle_goldens.py).  The Rust reference cannot be compiled in
 * this image (no cargo/rustc), so there is no oracle/_ref build.
 *
 * All index arrays are int64 (usize on the reference's 64-bit targets,
 * phylo2vec/src/types.rs:2-8).  Functions return 0 on success or a positive
 * code: (first offending index + 1) for an invalid vector, P2VO_EBADINPUT for
 * malformed encode input.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define P2VO_EBADINPUT (-2)
#define P2VO_ENOMEM (-3)
#define P2VO_EINVAL (-4)

typedef int64_t i64;

/* ------------------------------------------------------------------------ */
/* Validation: phylo2vec/src/vector/base.rs:50-67 (check_v / check_max) and  */
/* :92-101 (is_unordered).                                                   */
/* ------------------------------------------------------------------------ */

/* returns 0 if valid, else (index of first element with v[i] > 2i or < 0)+1 */
i64 p2vo_check_v(const i64 *v, i64 k) {
    for (i64 i = 0; i < k; ++i)
        if (v[i] < 0 || v[i] > 2 * i) return i + 1;
    return 0;
}

/* base.rs:92-101.  *bad receives idx+1 if check_max would panic before the
 * first unordered element is met (the reference returns early at the first
 * v[i] > i+1 and so never validates the tail). */
int p2vo_is_unordered(const i64 *v, i64 k, i64 *bad) {
    *bad = 0;
    for (i64 i = 0; i < k; ++i) {
        if (v[i] < 0 || v[i] > 2 * i) { *bad = i + 1; return 0; }
        if (v[i] > i + 1) return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Order-statistic AVL tree: phylo2vec/src/vector/avl.rs:5-218.              */
/* Same observable behaviour (insert at rank, lookup at rank, in-order dump, */
/* AVL rebalancing after every insert) on a node pool instead of Box<Node>.  */
/* ------------------------------------------------------------------------ */
typedef struct {
    i64 first, second;   /* the Pair payload */
    int32_t lc, rc;      /* children, -1 = none */
    int32_t height, size;
} avl_node;

typedef struct {
    avl_node *pool;
    int32_t used, root;
} avl_tree;

static inline int32_t avl_h(const avl_tree *t, int32_t x) { return x < 0 ? 0 : t->pool[x].height; }
static inline int32_t avl_s(const avl_tree *t, int32_t x) { return x < 0 ? 0 : t->pool[x].size; }

static inline void avl_pull(avl_tree *t, int32_t x) { /* avl.rs:68-71 */
    avl_node *n = &t->pool[x];
    int32_t hl = avl_h(t, n->lc), hr = avl_h(t, n->rc);
    n->height = 1 + (hl > hr ? hl : hr);
    n->size = 1 + avl_s(t, n->lc) + avl_s(t, n->rc);
}

static int32_t avl_rot_right(avl_tree *t, int32_t y) { /* avl.rs:73-95 */
    int32_t x = t->pool[y].lc;
    t->pool[y].lc = t->pool[x].rc;
    t->pool[x].rc = y;
    avl_pull(t, y);
    avl_pull(t, x);
    return x;
}

static int32_t avl_rot_left(avl_tree *t, int32_t x) { /* avl.rs:97-119 */
    int32_t y = t->pool[x].rc;
    t->pool[x].rc = t->pool[y].lc;
    t->pool[y].lc = x;
    avl_pull(t, x);
    avl_pull(t, y);
    return y;
}

static inline int32_t avl_bf(const avl_tree *t, int32_t x) { /* avl.rs:121-127 */
    return x < 0 ? 0 : avl_h(t, t->pool[x].lc) - avl_h(t, t->pool[x].rc);
}

static int32_t avl_fix(avl_tree *t, int32_t x) { /* avl.rs:129-151 */
    int32_t bf = avl_bf(t, x);
    if (bf > 1) {
        if (avl_bf(t, t->pool[x].lc) < 0) t->pool[x].lc = avl_rot_left(t, t->pool[x].lc);
        return avl_rot_right(t, x);
    }
    if (bf < -1) {
        if (avl_bf(t, t->pool[x].rc) > 0) t->pool[x].rc = avl_rot_right(t, t->pool[x].rc);
        return avl_rot_left(t, x);
    }
    return x;
}

/* avl.rs:157-173: descend by left-subtree size, then rebalance on the way up */
static int32_t avl_insert_at(avl_tree *t, int32_t x, i64 rank, int32_t fresh) {
    if (x < 0) return fresh;
    int32_t ls = avl_s(t, t->pool[x].lc);
    if (rank <= ls)
        t->pool[x].lc = avl_insert_at(t, t->pool[x].lc, rank, fresh);
    else
        t->pool[x].rc = avl_insert_at(t, t->pool[x].rc, rank - ls - 1, fresh);
    avl_pull(t, x);
    return avl_fix(t, x);
}

static void avl_insert(avl_tree *t, i64 rank, i64 a, i64 b) { /* avl.rs:153-155 */
    int32_t id = t->used++;
    avl_node *n = &t->pool[id];
    n->first = a; n->second = b; n->lc = n->rc = -1; n->height = 1; n->size = 1;
    t->root = avl_insert_at(t, t->root, rank, id);
}

/* avl.rs:175-191: out-of-range lookups yield (0,0) */
static void avl_lookup(const avl_tree *t, i64 rank, i64 *a, i64 *b) {
    int32_t x = t->root;
    while (x >= 0) {
        i64 ls = avl_s(t, t->pool[x].lc);
        if (rank < ls) x = t->pool[x].lc;
        else if (rank == ls) { *a = t->pool[x].first; *b = t->pool[x].second; return; }
        else { rank -= ls + 1; x = t->pool[x].rc; }
    }
    *a = 0; *b = 0;
}

/* avl.rs:193-211 */
static void avl_inorder(const avl_tree *t, i64 *pairs, int32_t *stack) {
    int32_t sp = 0, x = t->root;
    i64 w = 0;
    while (x >= 0 || sp > 0) {
        while (x >= 0) { stack[sp++] = x; x = t->pool[x].lc; }
        x = stack[--sp];
        pairs[2 * w] = t->pool[x].first;
        pairs[2 * w + 1] = t->pool[x].second;
        ++w;
        x = t->pool[x].rc;
    }
}

/* AVLTree::with_vector + inorder_traversal: avl.rs:35-52, convert.rs:83-87.
 * scratch: k avl_nodes + 128 int32 for the traversal stack. */
static int to_pairs_avl_ws(const i64 *v, i64 k, i64 *pairs, avl_node *pool, int32_t *stack) {
    if (k == 0) return 0;
    avl_tree t = { pool, 0, -1 };
    avl_insert(&t, 0, 0, 1);
    for (i64 i = 1; i < k; ++i) {
        i64 next_leaf = i + 1;
        if (v[i] <= i) {
            avl_insert(&t, 0, v[i], next_leaf);
        } else {
            i64 index = v[i] - next_leaf, a, b;
            avl_lookup(&t, index, &a, &b);
            avl_insert(&t, index + 1, a, next_leaf);
        }
    }
    avl_inorder(&t, pairs, stack);
    return 0;
}

int p2vo_to_pairs_avl(const i64 *v, i64 k, i64 *pairs) {
    if (k == 0) return 0;
    avl_node *pool = (avl_node *)malloc(sizeof(avl_node) * (size_t)k);
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * 128);
    if (!pool || !stack) { free(pool); free(stack); return P2VO_ENOMEM; }
    to_pairs_avl_ws(v, k, pairs, pool, stack);
    free(pool); free(stack);
    return 0;
}

/* convert.rs:39-76 (to_pairs_loop): Vec::insert == memmove */
int p2vo_to_pairs_loop(const i64 *v, i64 k, i64 *pairs) {
    i64 len = 0;
    for (i64 i = k - 1; i >= 0; --i) {
        if (v[i] <= i) { pairs[2 * len] = v[i]; pairs[2 * len + 1] = i + 1; ++len; }
    }
    for (i64 j = 1; j < k; ++j) {
        i64 next_leaf = j + 1;
        if (v[j] == 2 * j) {
            pairs[2 * len] = 0; pairs[2 * len + 1] = next_leaf; ++len;
        } else if (v[j] > j) {
            i64 index = len + v[j] - 2 * j;
            i64 first = pairs[2 * (index - 1)];
            memmove(&pairs[2 * (index + 1)], &pairs[2 * index], sizeof(i64) * 2 * (size_t)(len - index));
            pairs[2 * index] = first; pairs[2 * index + 1] = next_leaf; ++len;
        }
    }
    return 0;
}

/* Naive list-insert form of avl.rs:35-52 (O(k^2)); used by the tests as an
 * independent check of the AVL restatement. */
int p2vo_to_pairs_list(const i64 *v, i64 k, i64 *pairs) {
    if (k == 0) return 0;
    i64 len = 0;
    pairs[0] = 0; pairs[1] = 1; len = 1;
    for (i64 i = 1; i < k; ++i) {
        i64 rank, first;
        if (v[i] <= i) { rank = 0; first = v[i]; }
        else {
            i64 index = v[i] - (i + 1);
            first = (index < len) ? pairs[2 * index] : 0;
            rank = index + 1;
            if (rank > len) rank = len;
        }
        memmove(&pairs[2 * (rank + 1)], &pairs[2 * rank], sizeof(i64) * 2 * (size_t)(len - rank));
        pairs[2 * rank] = first; pairs[2 * rank + 1] = i + 1; ++len;
    }
    return 0;
}

/* convert.rs:103-109 */
static int to_pairs_ws(const i64 *v, i64 k, i64 *pairs, avl_node *pool, int32_t *stack) {
    i64 bad;
    int unordered = p2vo_is_unordered(v, k, &bad);
    if (bad) return (int)bad;           /* the reference panics here */
    if (unordered) {
        /* the reference does not validate the tail; we do, and report it */
        i64 b2 = p2vo_check_v(v, k);
        if (b2) return (int)b2;
        return to_pairs_avl_ws(v, k, pairs, pool, stack);
    }
    return p2vo_to_pairs_loop(v, k, pairs);
}

int p2vo_to_pairs(const i64 *v, i64 k, i64 *pairs) {
    if (k == 0) return 0;
    avl_node *pool = (avl_node *)malloc(sizeof(avl_node) * (size_t)k);
    int32_t *stack = (int32_t *)malloc(sizeof(int32_t) * 128);
    if (!pool || !stack) { free(pool); free(stack); return P2VO_ENOMEM; }
    int rc = to_pairs_ws(v, k, pairs, pool, stack);
    free(pool); free(stack);
    return rc;
}

/* convert.rs:167-188 (to_ancestry) given the pairs */
static void ancestry_from_pairs(const i64 *pairs, i64 k, i64 *anc, i64 *parents) {
    for (i64 i = 0; i <= 2 * k + 1; ++i) parents[i] = i;
    for (i64 i = 0; i < k; ++i) {
        i64 c1 = pairs[2 * i], c2 = pairs[2 * i + 1], p = k + 1 + i;
        anc[3 * i] = parents[c1]; anc[3 * i + 1] = parents[c2]; anc[3 * i + 2] = p;
        parents[c1] = p; parents[c2] = p;
    }
}

/* convert.rs:227-248 (to_edges) given the pairs */
static void edges_from_pairs(const i64 *pairs, i64 k, i64 *edges, i64 *parents) {
    for (i64 i = 0; i <= 2 * k + 1; ++i) parents[i] = i;
    for (i64 i = 0; i < k; ++i) {
        i64 c1 = pairs[2 * i], c2 = pairs[2 * i + 1], p = k + 1 + i;
        edges[4 * i] = parents[c1]; edges[4 * i + 1] = p;
        edges[4 * i + 2] = parents[c2]; edges[4 * i + 3] = p;
        parents[c1] = p; parents[c2] = p;
    }
}

typedef struct {
    avl_node *pool; int32_t *stack; i64 *pairs; i64 *parents; i64 cap;
} decode_ws;

static int ws_init(decode_ws *w, i64 k) {
    w->cap = k;
    w->pool = (avl_node *)malloc(sizeof(avl_node) * (size_t)(k + 1));
    w->stack = (int32_t *)malloc(sizeof(int32_t) * 128);
    w->pairs = (i64 *)malloc(sizeof(i64) * 2 * (size_t)(k + 1));
    w->parents = (i64 *)malloc(sizeof(i64) * (size_t)(2 * k + 2));
    return (w->pool && w->stack && w->pairs && w->parents) ? 0 : P2VO_ENOMEM;
}
static void ws_free(decode_ws *w) { free(w->pool); free(w->stack); free(w->pairs); free(w->parents); }

int p2vo_to_ancestry(const i64 *v, i64 k, i64 *anc) {
    if (k == 0) return 0;
    decode_ws w; if (ws_init(&w, k)) { ws_free(&w); return P2VO_ENOMEM; }
    int rc = to_pairs_ws(v, k, w.pairs, w.pool, w.stack);
    if (!rc) ancestry_from_pairs(w.pairs, k, anc, w.parents);
    ws_free(&w);
    return rc;
}

int p2vo_to_edges(const i64 *v, i64 k, i64 *edges) {
    if (k == 0) return 0;
    decode_ws w; if (ws_init(&w, k)) { ws_free(&w); return P2VO_ENOMEM; }
    int rc = to_pairs_ws(v, k, w.pairs, w.pool, w.stack);
    if (!rc) edges_from_pairs(w.pairs, k, edges, w.parents);
    ws_free(&w);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* Encode: Fenwick convert.rs:286-316, build_vector :320-337,                */
/* build_cherries :399-440, order_cherries :449-477, from_* :23-30,129-136,  */
/* 201-213.                                                                  */
/* ------------------------------------------------------------------------ */
static inline i64 fen_prefix(const i64 *d, i64 i) {
    i64 s = 0;
    while (i > 0) { s += d[i]; i -= i & (-i); }
    return s;
}
static inline void fen_update(i64 *d, i64 n, i64 i, i64 delta) {
    while (i <= n) { d[i] += delta; i += i & (-i); }
}

/* cherries: k rows [c1, c2, c_max]; fen: k+1 scratch */
static int build_vector_ws(const i64 *ch, i64 k, i64 *v, i64 *fen) {
    memset(fen, 0, sizeof(i64) * (size_t)(k + 1));
    memset(v, 0, sizeof(i64) * (size_t)k);
    for (i64 r = 0; r < k; ++r) {
        i64 c1 = ch[3 * r], c2 = ch[3 * r + 1], cm = ch[3 * r + 2];
        if (cm < 1 || cm > k) return P2VO_EBADINPUT; /* reference: index panic */
        i64 idx = fen_prefix(fen, cm - 1);
        v[cm - 1] = (idx == 0) ? (c1 < c2 ? c1 : c2) : cm - 1 + idx;
        fen_update(fen, k, cm, 1);
    }
    return 0;
}

int p2vo_build_vector(const i64 *cherries, i64 k, i64 *v) {
    if (k == 0) return 0;
    i64 *fen = (i64 *)malloc(sizeof(i64) * (size_t)(k + 1));
    if (!fen) return P2VO_ENOMEM;
    int rc = build_vector_ws(cherries, k, v, fen);
    free(fen);
    return rc;
}

int p2vo_from_pairs(const i64 *pairs, i64 k, i64 *v) { /* convert.rs:23-30 */
    if (k == 0) return 0;
    i64 *ch = (i64 *)malloc(sizeof(i64) * 3 * (size_t)k);
    if (!ch) return P2VO_ENOMEM;
    for (i64 r = 0; r < k; ++r) {
        i64 c1 = pairs[2 * r], c2 = pairs[2 * r + 1];
        ch[3 * r] = c1; ch[3 * r + 1] = c2; ch[3 * r + 2] = c1 > c2 ? c1 : c2;
    }
    int rc = p2vo_build_vector(ch, k, v);
    free(ch);
    return rc;
}

/* convert.rs:399-440.  In place; min_desc: 2k+2 scratch. */
static int build_cherries_ws(i64 *anc, i64 k, i64 *min_desc) {
    i64 nn = 2 * k + 2;
    for (i64 i = 0; i < nn; ++i) min_desc[i] = -1;
    for (i64 r = 0; r < k; ++r) {
        i64 c1 = anc[3 * r], c2 = anc[3 * r + 1], p = anc[3 * r + 2];
        if (c1 < 0 || c1 >= nn || c2 < 0 || c2 >= nn || p < 0 || p >= nn) return P2VO_EBADINPUT;
        i64 m1 = min_desc[c1] >= 0 ? min_desc[c1] : c1;
        i64 m2 = min_desc[c2] >= 0 ? min_desc[c2] : c2;
        i64 lo = m1 < m2 ? m1 : m2, hi = m1 < m2 ? m2 : m1;
        min_desc[p] = lo;
        anc[3 * r] = m1; anc[3 * r + 1] = m2; anc[3 * r + 2] = hi;
    }
    return 0;
}

int p2vo_from_ancestry(const i64 *anc, i64 k, i64 *v) { /* convert.rs:129-136 */
    if (k == 0) return 0;
    i64 *ch = (i64 *)malloc(sizeof(i64) * 3 * (size_t)k);
    i64 *md = (i64 *)malloc(sizeof(i64) * (size_t)(2 * k + 2));
    if (!ch || !md) { free(ch); free(md); return P2VO_ENOMEM; }
    memcpy(ch, anc, sizeof(i64) * 3 * (size_t)k);
    int rc = build_cherries_ws(ch, k, md);
    if (!rc) rc = p2vo_build_vector(ch, k, v);
    free(ch); free(md);
    return rc;
}

/* convert.rs:449-477 (order_cherries), in place */
int p2vo_order_cherries(i64 *anc, i64 k) {
    if (k == 0) return 0;
    i64 mx = anc[2];
    for (i64 r = 1; r < k; ++r) if (anc[3 * r + 2] > mx) mx = anc[3 * r + 2];
    i64 offset = mx - 2 * k + 1;
    i64 *tmp = (i64 *)malloc(sizeof(i64) * 3 * (size_t)k);
    i64 *md = (i64 *)malloc(sizeof(i64) * (size_t)(2 * k + 2));
    if (!tmp || !md) { free(tmp); free(md); return P2VO_ENOMEM; }
    memcpy(tmp, anc, sizeof(i64) * 3 * (size_t)k);
    int rc = 0;
    for (i64 r = 0; r < k; ++r) {
        i64 idx = anc[3 * r + 2] - k - offset;
        if (idx < 0 || idx >= k) { rc = P2VO_EBADINPUT; break; }
        tmp[3 * idx] = anc[3 * r]; tmp[3 * idx + 1] = anc[3 * r + 1]; tmp[3 * idx + 2] = anc[3 * r + 2];
    }
    if (!rc) { memcpy(anc, tmp, sizeof(i64) * 3 * (size_t)k); rc = build_cherries_ws(anc, k, md); }
    free(tmp); free(md);
    return rc;
}

/* convert.rs:201-213: edges is (2k, 2) (child, parent) */
int p2vo_from_edges(const i64 *edges, i64 n_edges, i64 *v) {
    if (n_edges % 2 != 0) return P2VO_EBADINPUT; /* assert at convert.rs:202 */
    i64 k = n_edges / 2;
    if (k == 0) return 0;
    i64 *anc = (i64 *)malloc(sizeof(i64) * 3 * (size_t)k);
    if (!anc) return P2VO_ENOMEM;
    for (i64 i = 0; i < k; ++i) {
        anc[3 * i] = edges[4 * i];         /* edges[2i].child   */
        anc[3 * i + 1] = edges[4 * i + 2]; /* edges[2i+1].child */
        anc[3 * i + 2] = edges[4 * i + 1]; /* edges[2i].parent  */
    }
    int rc = p2vo_order_cherries(anc, k);
    if (!rc) rc = p2vo_build_vector(anc, k, v);
    free(anc);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* Cophenetic distances.                                                     */
/* ------------------------------------------------------------------------ */

/* Literal restatement of vector/graph.rs:20-90 including the (2k+1)^2 scratch
 * matrix, the visiting order and the write order at :61-79.  bls == NULL means
 * unit lengths (graph.rs:10-12).  out is (k+1)x(k+1), row-major. */
int p2vo_cophenetic_literal(const i64 *v, const double *bls, i64 k, int unrooted, double *out) {
    if (k == 0) { out[0] = 0.0; return 0; }
    if (k == 1) { /* graph.rs:37-42 */
        double s = bls ? bls[0] + bls[1] : 2.0;
        out[0] = 0.0; out[1] = s; out[2] = s; out[3] = 0.0;
        return 0;
    }
    i64 *anc = (i64 *)malloc(sizeof(i64) * 3 * (size_t)k);
    if (!anc) return P2VO_ENOMEM;
    int rc = p2vo_to_ancestry(v, k, anc);
    if (rc) { free(anc); return rc; }
    if (unrooted) { /* graph.rs:46-50 */
        i64 *last = &anc[3 * (k - 1)];
        i64 mx = last[0] > last[1] ? last[0] : last[1];
        if (last[2] > mx) mx = last[2];
        last[2] = mx - 1;
    }
    i64 N = 2 * k + 1;
    double *dist = (double *)calloc((size_t)N * (size_t)N, sizeof(double));
    i64 *visited = (i64 *)malloc(sizeof(i64) * 3 * (size_t)k);
    if (!dist || !visited) { free(anc); free(dist); free(visited); return P2VO_ENOMEM; }
    i64 nv = 0;
    for (i64 i = 0; i < k; ++i) {
        i64 r = k - i - 1;
        i64 c1 = anc[3 * r], c2 = anc[3 * r + 1], p = anc[3 * r + 2];
        double bl1 = bls ? bls[2 * r] : 1.0, bl2 = bls ? bls[2 * r + 1] : 1.0;
        if (nv > 0) {
            for (i64 t = 0; t < nv - 1; ++t) { /* all_visited[0..len-1] */
                i64 x = visited[t];
                double d1 = dist[p * N + x] + bl1;
                double d2 = dist[p * N + x] + bl2;
                dist[c1 * N + x] = d1; dist[x * N + c1] = d1;
                dist[c2 * N + x] = d2; dist[x * N + c2] = d2;
            }
        }
        dist[c1 * N + c2] = bl1 + bl2; dist[c2 * N + c1] = bl1 + bl2;
        dist[c1 * N + p] = bl1; dist[p * N + c1] = bl1;
        dist[c2 * N + p] = bl2; dist[p * N + c2] = bl2;
        visited[nv++] = c1; visited[nv++] = c2; visited[nv++] = p;
    }
    i64 n = k + 1;
    for (i64 a = 0; a < n; ++a) memcpy(&out[a * n], &dist[a * N], sizeof(double) * (size_t)n);
    free(anc); free(dist); free(visited);
    return 0;
}

/* Closed-form rows [row_begin,row_end) of the same matrix for sizes where the
 * literal scratch does not fit (n = 65536 would need 137 GB).  Each entry is
 * the sum of edge lengths on the leaf-to-leaf path, accumulated in long double
 * (64-bit mantissa) and rounded once; "unrooted" zeroes the root edge that
 * leads to node 2k-1 (consequence of graph.rs:46-50 with the write order at
 * :72-79).  Pinned against p2vo_cophenetic_literal in tests/test_oracle.py.
 * out is (row_end-row_begin) x (k+1). */
int p2vo_cophenetic_rows(const i64 *v, const double *bls, i64 k, int unrooted,
                         i64 row_begin, i64 row_end, double *out) {
    i64 n = k + 1;
    if (k == 0) { if (row_end > row_begin) out[0] = 0.0; return 0; }
    if (k == 1) {
        double s = bls ? bls[0] + bls[1] : 2.0;
        for (i64 a = row_begin; a < row_end; ++a) {
            out[(a - row_begin) * 2 + a] = 0.0;
            out[(a - row_begin) * 2 + (1 - a)] = s;
        }
        return 0;
    }
    i64 *anc = (i64 *)malloc(sizeof(i64) * 3 * (size_t)k);
    i64 N = 2 * k + 1;
    i64 *par = (i64 *)malloc(sizeof(i64) * (size_t)N);
    long double *len = (long double *)malloc(sizeof(long double) * (size_t)N);
    long double *up = (long double *)malloc(sizeof(long double) * (size_t)N);
    i64 *stamp = (i64 *)malloc(sizeof(i64) * (size_t)N);
    if (!anc || !par || !len || !up || !stamp) { free(anc); free(par); free(len); free(up); free(stamp); return P2VO_ENOMEM; }
    int rc = p2vo_to_ancestry(v, k, anc);
    if (rc) { free(anc); free(par); free(len); free(up); free(stamp); return rc; }
    for (i64 r = 0; r < k; ++r) {
        i64 c1 = anc[3 * r], c2 = anc[3 * r + 1], p = anc[3 * r + 2];
        par[c1] = p; par[c2] = p;
        len[c1] = bls ? (long double)bls[2 * r] : 1.0L;
        len[c2] = bls ? (long double)bls[2 * r + 1] : 1.0L;
    }
    par[2 * k] = -1; len[2 * k] = 0.0L;
    if (unrooted) len[2 * k - 1] = 0.0L;
    for (i64 x = 0; x < N; ++x) stamp[x] = -1;
    for (i64 a = row_begin; a < row_end; ++a) {
        /* distance from leaf a up to each of its ancestors */
        long double acc = 0.0L;
        for (i64 x = a; x >= 0; x = par[x]) { stamp[x] = a; up[x] = acc; acc += len[x]; }
        double *row = &out[(a - row_begin) * n];
        for (i64 b = 0; b < n; ++b) {
            if (b == a) { row[b] = 0.0; continue; }
            long double s = 0.0L;
            i64 x = b;
            while (stamp[x] != a) { s += len[x]; x = par[x]; }
            row[b] = (double)(s + up[x]);
        }
    }
    free(anc); free(par); free(len); free(up); free(stamp);
    return 0;
}

/* matrix/base.rs:108-121 (parse_matrix) + matrix/graph.rs:23-26.  m is (k,3)
 * f64 row-major; v = (usize) m[r][0] (truncating cast), bls = m[r][1..3]. */
int p2vo_cophenetic_matrix_literal(const double *m, i64 k, int unrooted, double *out) {
    i64 *v = (i64 *)malloc(sizeof(i64) * (size_t)(k + 1));
    double *bls = (double *)malloc(sizeof(double) * 2 * (size_t)(k + 1));
    if (!v || !bls) { free(v); free(bls); return P2VO_ENOMEM; }
    for (i64 r = 0; r < k; ++r) {
        double x = m[3 * r];
        v[r] = (x != x || x <= 0.0) ? 0 : (i64)x;  /* Rust `as usize` saturates; NaN -> 0 */
        bls[2 * r] = m[3 * r + 1]; bls[2 * r + 1] = m[3 * r + 2];
    }
    int rc = p2vo_cophenetic_literal(v, bls, k, unrooted, out);
    free(v); free(bls);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* Variance-covariance matrix and node depths.             */
/* ------------------------------------------------------------------------ */

/* Literal restatement of vector/graph.rs:211-258 (_vcv) with
 * get_descendants_and_edges :171-204: walking the edges from the root down,
 * root_to_x[child] = root_to_x[parent] + bl (plain f64, that order), and the
 * two leaf sets under a cherry's children get out[l][r] = out[r][l] =
 * root_to_x[parent]; the diagonal is root_to_x[leaf].  The leaf sets are kept
 * as linked lists (a BitSet in the reference; only membership matters).
 * bls == NULL means unit lengths.  k == 0 is the reference's assert (:214-217)
 * and returns P2VO_EINVAL.  out is (k+1)x(k+1). */
int p2vo_vcv_literal(const i64 *v, const double *bls, i64 k, double *out) {
    if (k <= 0) return P2VO_EINVAL;
    i64 n = k + 1, N = 2 * k + 1;
    i64 *edges = (i64 *)malloc(sizeof(i64) * 4 * (size_t)k);
    i64 *head = (i64 *)malloc(sizeof(i64) * (size_t)N);
    i64 *tail = (i64 *)malloc(sizeof(i64) * (size_t)N);
    i64 *next = (i64 *)malloc(sizeof(i64) * (size_t)n);
    double *root_to_x = (double *)calloc((size_t)N, sizeof(double));
    if (!edges || !head || !tail || !next || !root_to_x) {
        free(edges); free(head); free(tail); free(next); free(root_to_x); return P2VO_ENOMEM;
    }
    int rc = p2vo_to_edges(v, k, edges);
    if (rc) { free(edges); free(head); free(tail); free(next); free(root_to_x); return rc; }
    /* desc[leaf] = {leaf}; desc[parent] = desc[c1] U desc[c2] in cherry order.  The lists are
     * concatenated destructively, so the sets are rebuilt top-down below from a second copy:
     * keep, for every node, the [first,last] leaves of its list and cut nothing -- a child's
     * list is a contiguous run of its parent's list. */
    for (i64 i = 0; i < n; ++i) { head[i] = i; tail[i] = i; next[i] = -1; }
    for (i64 r = 0; r < k; ++r) {
        i64 c1 = edges[4 * r], c2 = edges[4 * r + 2], p = edges[4 * r + 1];
        next[tail[c1]] = head[c2];
        head[p] = head[c1]; tail[p] = tail[c2];
    }
    for (size_t i = 0; i < (size_t)n * (size_t)n; ++i) out[i] = 0.0;
    for (i64 i = 2 * k - 1; i >= 0; --i) {          /* graph.rs:236-254 */
        i64 ci = edges[2 * i], p = edges[2 * i + 1];
        double var = root_to_x[p];
        root_to_x[ci] = var + (bls ? bls[i] : 1.0);  /* bls[i/2][i%2] == flat bls[i] */
        if (i % 2 == 1) {
            i64 cj = edges[2 * (i - 1)];
            for (i64 l = head[cj];; l = next[l]) {
                for (i64 r = head[ci];; r = next[r]) {
                    out[l * n + r] = var; out[r * n + l] = var;
                    if (r == tail[ci]) break;
                }
                if (l == tail[cj]) break;
            }
        }
    }
    for (i64 i = 0; i < n; ++i) out[i * n + i] = root_to_x[i];   /* :257 */
    free(edges); free(head); free(tail); free(next); free(root_to_x);
    return 0;
}

/* matrix/graph.rs:61-64 */
int p2vo_vcv_matrix_literal(const double *m, i64 k, double *out) {
    if (k <= 0) return P2VO_EINVAL;
    i64 *v = (i64 *)malloc(sizeof(i64) * (size_t)(k + 1));
    double *bls = (double *)malloc(sizeof(double) * 2 * (size_t)(k + 1));
    if (!v || !bls) { free(v); free(bls); return P2VO_ENOMEM; }
    for (i64 r = 0; r < k; ++r) {
        double x = m[3 * r];
        v[r] = (x != x || x <= 0.0) ? 0 : (i64)x;
        bls[2 * r] = m[3 * r + 1]; bls[2 * r + 1] = m[3 * r + 2];
    }
    int rc = p2vo_vcv_literal(v, bls, k, out);
    free(v); free(bls);
    return rc;
}

/* vector/ops.rs:201-233 (_get_node_depths): ancestry walked from the last row up,
 * depths[child] = depths[parent] + bl.  out has 2k+1 entries. */
int p2vo_node_depths(const i64 *v, const double *bls, i64 k, double *out) {
    if (k == 0) { out[0] = 0.0; return 0; }
    i64 *anc = (i64 *)malloc(sizeof(i64) * 3 * (size_t)k);
    if (!anc) return P2VO_ENOMEM;
    int rc = p2vo_to_ancestry(v, k, anc);
    if (rc) { free(anc); return rc; }
    for (i64 i = 0; i <= 2 * k; ++i) out[i] = 0.0;
    for (i64 r = k - 1; r >= 0; --r) {
        i64 c1 = anc[3 * r], c2 = anc[3 * r + 1], p = anc[3 * r + 2];
        out[c1] = out[p] + (bls ? bls[2 * r] : 1.0);
        out[c2] = out[p] + (bls ? bls[2 * r + 1] : 1.0);
    }
    free(anc);
    return 0;
}

/* Rows [row_begin,row_end) of the vcv matrix for sizes where the full matrix is too large
 * for the host: out[a][b] = node depth (as p2vo_node_depths) of the LCA of leaves a and b. */
int p2vo_vcv_rows(const i64 *v, const double *bls, i64 k, i64 row_begin, i64 row_end, double *out) {
    if (k <= 0) return P2VO_EINVAL;
    i64 n = k + 1, N = 2 * k + 1;
    i64 *anc = (i64 *)malloc(sizeof(i64) * 3 * (size_t)k);
    i64 *par = (i64 *)malloc(sizeof(i64) * (size_t)N);
    i64 *stamp = (i64 *)malloc(sizeof(i64) * (size_t)N);
    double *depth = (double *)malloc(sizeof(double) * (size_t)N);
    if (!anc || !par || !stamp || !depth) { free(anc); free(par); free(stamp); free(depth); return P2VO_ENOMEM; }
    int rc = p2vo_to_ancestry(v, k, anc);
    if (!rc) rc = p2vo_node_depths(v, bls, k, depth);
    if (rc) { free(anc); free(par); free(stamp); free(depth); return rc; }
    for (i64 r = 0; r < k; ++r) { par[anc[3 * r]] = anc[3 * r + 2]; par[anc[3 * r + 1]] = anc[3 * r + 2]; }
    par[2 * k] = -1;
    for (i64 x = 0; x < N; ++x) stamp[x] = -1;
    for (i64 a = row_begin; a < row_end; ++a) {
        for (i64 x = a; x >= 0; x = par[x]) stamp[x] = a;
        double *row = &out[(a - row_begin) * n];
        for (i64 b = 0; b < n; ++b) {
            i64 x = b;
            while (stamp[x] != a) x = par[x];
            row[b] = depth[x];
        }
    }
    free(anc); free(par); free(stamp); free(depth);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* to_newick (convert.rs:341-397).  Host-only; needed for BASELINE config C1  */
/* and because the reference's Newick goldens also pin to_pairs.             */
/* Returns the string length (excluding NUL), or <0.  buf may be NULL to size */
/* the output.                                                               */
/* ------------------------------------------------------------------------ */
static size_t dec_len(i64 x) { size_t l = 1; while (x >= 10) { x /= 10; ++l; } return l; }

i64 p2vo_to_newick(const i64 *v, i64 k, char *buf, i64 cap) {
    i64 n = k + 1;
    if (k == 0) { /* no pairs: the cache of the single leaf, convert.rs:361-381 */
        if (buf && cap > 2) { buf[0] = '0'; buf[1] = ';'; buf[2] = 0; }
        return 2;
    }
    i64 *pairs = (i64 *)malloc(sizeof(i64) * 2 * (size_t)k);
    if (!pairs) return P2VO_ENOMEM;
    int rc = p2vo_to_pairs(v, k, pairs);
    if (rc) { free(pairs); return -1000 - rc; }
    /* per-leaf growable strings, concatenated exactly like the reference's cache */
    char **cache = (char **)calloc((size_t)n, sizeof(char *));
    size_t *len = (size_t *)calloc((size_t)n, sizeof(size_t));
    size_t *capv = (size_t *)calloc((size_t)n, sizeof(size_t));
    size_t *nopen = (size_t *)calloc((size_t)n, sizeof(size_t));
    if (!cache || !len || !capv || !nopen) { free(pairs); free(cache); free(len); free(capv); free(nopen); return P2VO_ENOMEM; }
    for (i64 i = 0; i < k; ++i) nopen[pairs[2 * i]]++;          /* prepare_cache :341-358 */
    for (i64 c = 0; c < n; ++c) {
        size_t need = nopen[c] + dec_len(c) + 1;
        capv[c] = need < 32 ? 32 : need * 2;
        cache[c] = (char *)malloc(capv[c]);
        memset(cache[c], '(', nopen[c]);
        len[c] = nopen[c] + (size_t)sprintf(cache[c] + nopen[c], "%lld", (long long)c);
    }
    for (i64 i = 0; i < k; ++i) {                               /* build_newick :361-381 */
        i64 c1 = pairs[2 * i], c2 = pairs[2 * i + 1];
        char sp[24]; size_t spl = (size_t)sprintf(sp, "%lld", (long long)(n + i));
        size_t need = len[c1] + 1 + len[c2] + 1 + spl + 1;
        if (need > capv[c1]) { capv[c1] = need * 2; cache[c1] = (char *)realloc(cache[c1], capv[c1]); }
        cache[c1][len[c1]++] = ',';
        memcpy(cache[c1] + len[c1], cache[c2], len[c2]); len[c1] += len[c2];
        cache[c1][len[c1]++] = ')';
        memcpy(cache[c1] + len[c1], sp, spl); len[c1] += spl;
        free(cache[c2]); cache[c2] = NULL; len[c2] = 0;
    }
    i64 total = (i64)len[0] + 1;
    if (buf && cap > total) { memcpy(buf, cache[0], len[0]); buf[len[0]] = ';'; buf[len[0] + 1] = 0; }
    for (i64 c = 0; c < n; ++c) free(cache[c]);
    free(pairs); free(cache); free(len); free(capv); free(nopen);
    return total;
}

/* ------------------------------------------------------------------------ */
/* Balance indices (vector/balance.rs:14-58), literally: sums over the leaf    */
/* depths of get_node_depths in leaf order.  out[0] = sackin (the reference    */
/* casts the f64 sum to usize), out[1] = leaf_depth_variance, out[2] = b2.     */
/* ------------------------------------------------------------------------ */
int p2vo_balance(const i64 *v, i64 k, double *out) {
    i64 n = k + 1;
    double *depths = (double *)malloc(sizeof(double) * (size_t)(2 * k + 1));
    if (!depths) return P2VO_ENOMEM;
    int rc = p2vo_node_depths(v, NULL, k, depths);
    if (rc) { free(depths); return rc; }
    double sum = 0.0, sum_sq = 0.0, b2 = 0.0;
    for (i64 i = 0; i < n; ++i) sum += depths[i];                                  /* :17 */
    out[0] = (double)(unsigned long long)sum;
    sum = 0.0;
    for (i64 i = 0; i < n; ++i) { sum += depths[i]; sum_sq += depths[i] * depths[i]; }   /* :25-30 */
    double nn = (double)n, m = sum / nn;
    out[1] = sum_sq / nn - m * m;                                                  /* powi(2) */
    for (i64 i = 0; i < n; ++i) b2 += depths[i] * ldexp(1.0, -(int)depths[i]);     /* 2f64.powi(-d) :55 */
    out[2] = b2;
    free(depths);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* from_newick, vectors (convert.rs:268-278) on top of newick::parse          */
/* (newick/mod.rs:73-127): the literal stack walk; has_parents = some ")" is  */
/* followed by a digit (newick_patterns.rs `\)(\d+)`).                        */
/* anc: up to cap rows [c1, c2, p]; returns the number of rows or <0.         */
/* ------------------------------------------------------------------------ */
i64 p2vo_newick_parse(const char *s, i64 len, i64 *anc, i64 cap) {
    i64 *stack = (i64 *)malloc(sizeof(i64) * (size_t)(len + 2));
    if (!stack) return P2VO_ENOMEM;
    i64 sp = 0, rows = 0, i = 0, rc = 0;
    while (i < len) {
        char c = s[i];
        if (c == ')') {
            if (sp < 2) { rc = P2VO_EBADINPUT; break; }          /* StackUnderflow */
            i64 c2 = stack[--sp], c1 = stack[--sp];
            i64 e = i + 1;                                        /* node_substr: up to , ) ; */
            while (e < len && s[e] != ',' && s[e] != ')' && s[e] != ';') ++e;
            if (e == len) e = i + 1;                              /* no terminator: node_substr returns "" (mod.rs:23-38) */
            if (rows >= cap) { rc = P2VO_EBADINPUT; break; }
            if (e == i + 1) {                                     /* no parent label */
                i64 lo = c1 < c2 ? c1 : c2, hi = c1 < c2 ? c2 : c1;
                anc[3 * rows] = c1; anc[3 * rows + 1] = c2; anc[3 * rows + 2] = hi; ++rows;
                stack[sp++] = lo;
            } else {
                i64 p = 0;
                for (i64 j = i + 1; j < e; ++j) {
                    if (s[j] < '0' || s[j] > '9') { rc = P2VO_EBADINPUT; break; }   /* ParseIntError */
                    p = p * 10 + (s[j] - '0');
                }
                if (rc) break;
                anc[3 * rows] = c1; anc[3 * rows + 1] = c2; anc[3 * rows + 2] = p; ++rows;
                stack[sp++] = p;
                i = e - 1;
            }
        } else if (c >= '0' && c <= '9') {
            i64 e = i;
            while (e < len && s[e] != ',' && s[e] != ')' && s[e] != ';') ++e;
            if (e == len) { rc = P2VO_EBADINPUT; break; }         /* "" does not parse as an integer */
            i64 x = 0;
            for (i64 j = i; j < e; ++j) {
                if (s[j] < '0' || s[j] > '9') { rc = P2VO_EBADINPUT; break; }
                x = x * 10 + (s[j] - '0');
            }
            if (rc) break;
            stack[sp++] = x;
            i = e - 1;
        }
        ++i;
    }
    free(stack);
    return rc ? rc : rows;
}

/* convert.rs:534-573 (order_cherries_no_parents), in place */
static int order_cherries_no_parents(i64 *anc, i64 k) {
    i64 *sister = (i64 *)malloc(sizeof(i64) * (size_t)k);
    i64 *visited = (i64 *)malloc(sizeof(i64) * (size_t)(2 * k + 2));
    i64 *idx = (i64 *)malloc(sizeof(i64) * (size_t)k);
    i64 *tmp = (i64 *)malloc(sizeof(i64) * 3 * (size_t)k);
    if (!sister || !visited || !idx || !tmp) { free(sister); free(visited); free(idx); free(tmp); return P2VO_ENOMEM; }
    int rc = 0;
    for (i64 i = 0; i < 2 * k + 2; ++i) visited[i] = -1;
    for (i64 r = 0; r < k && !rc; ++r) {
        i64 c1 = anc[3 * r], c2 = anc[3 * r + 1], cm = anc[3 * r + 2];
        i64 lo = c1 < c2 ? c1 : c2;
        if (lo < 0 || lo > 2 * k + 1) { rc = P2VO_EBADINPUT; break; }
        i64 sis = (visited[lo] >= 0 && visited[lo] < cm) ? visited[lo] : cm;
        sister[r] = sis; visited[lo] = sis;
    }
    if (!rc) {      /* stable argsort, descending by sister: insertion sort keeps equal keys in row order */
        for (i64 r = 0; r < k; ++r) idx[r] = r;
        for (i64 a = 1; a < k; ++a) {
            i64 x = idx[a], b = a;
            while (b > 0 && sister[idx[b - 1]] < sister[x]) { idx[b] = idx[b - 1]; --b; }
            idx[b] = x;
        }
        for (i64 r = 0; r < k; ++r) memcpy(tmp + 3 * r, anc + 3 * idx[r], sizeof(i64) * 3);
        memcpy(anc, tmp, sizeof(i64) * 3 * (size_t)k);
    }
    free(sister); free(visited); free(idx); free(tmp);
    return rc;
}

/* v gets k = (number of cherries) entries; returns k or <0 */
i64 p2vo_from_newick(const char *s, i64 len, i64 *v, i64 cap) {
    if (len == 0) return 0;
    i64 *anc = (i64 *)malloc(sizeof(i64) * 3 * (size_t)(len + 1));
    if (!anc) return P2VO_ENOMEM;
    i64 k = p2vo_newick_parse(s, len, anc, len);
    if (k < 0 || k > cap) { free(anc); return k < 0 ? k : P2VO_EBADINPUT; }
    int parents = 0;
    for (i64 i = 0; i + 1 < len; ++i) if (s[i] == ')' && s[i + 1] >= '0' && s[i + 1] <= '9') { parents = 1; break; }
    int rc = 0;
    if (k > 0) {
        rc = parents ? p2vo_order_cherries(anc, k) : order_cherries_no_parents(anc, k);
        if (!rc) rc = p2vo_build_vector(anc, k, v);
    }
    free(anc);
    return rc ? rc : k;
}

/* strings at in + b*stride, lengths[b] bytes each; v (B,k); status per tree */
int p2vo_from_newick_batch(const char *in, i64 stride, const i64 *lengths, i64 B, i64 k, i64 *v, int32_t *status,
                           int nthreads) {
    int bad = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(| : bad)
#endif
    for (i64 b = 0; b < B; ++b) {
        i64 got = p2vo_from_newick(in + b * stride, lengths[b], v + b * k, k);
        int rc = (got == k) ? 0 : (got < 0 ? (int)got : P2VO_EBADINPUT);
        if (status) status[b] = rc;
        if (rc) bad |= 1;
    }
    return bad;
}

/* ------------------------------------------------------------------------ */
/* Batch drivers for the CPU baseline: the reference has no batch API and no  */
/* threads; one reference call == one tree on one core, so the    */
/* baseline is an OpenMP parallel-for over trees.                            */
/* op: 0 to_pairs (B,k,2)  1 to_ancestry (B,k,3)  2 to_edges (B,2k,2)        */
/*     3 from_pairs  4 from_ancestry  5 from_edges -> (B,k)                  */
/* status (may be NULL) gets the per-tree return code.                       */
/* ------------------------------------------------------------------------ */
int p2vo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int p2vo_batch(int op, const i64 *in, i64 B, i64 k, i64 *out, int32_t *status, int nthreads) {
    i64 in_stride, out_stride;
    switch (op) {
    case 0: in_stride = k; out_stride = 2 * k; break;
    case 1: in_stride = k; out_stride = 3 * k; break;
    case 2: in_stride = k; out_stride = 4 * k; break;
    case 3: in_stride = 2 * k; out_stride = k; break;
    case 4: in_stride = 3 * k; out_stride = k; break;
    case 5: in_stride = 4 * k; out_stride = k; break;
    default: return P2VO_EBADINPUT;
    }
    int any = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel num_threads(nthreads)
#endif
    {
        decode_ws w; int ok = (op <= 2) ? !ws_init(&w, k) : 1;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
        for (i64 b = 0; b < B; ++b) {
            int rc = 0;
            const i64 *x = in + b * in_stride;
            i64 *y = out + b * out_stride;
            if (!ok) rc = P2VO_ENOMEM;
            else if (op <= 2) {
                rc = to_pairs_ws(x, k, op == 0 ? y : w.pairs, w.pool, w.stack);
                if (!rc && op == 1) ancestry_from_pairs(w.pairs, k, y, w.parents);
                if (!rc && op == 2) edges_from_pairs(w.pairs, k, y, w.parents);
            } else if (op == 3) rc = p2vo_from_pairs(x, k, y);
            else if (op == 4) rc = p2vo_from_ancestry(x, k, y);
            else rc = p2vo_from_edges(x, 2 * k, y);
            if (status) status[b] = rc;
            if (rc) any = 1;
        }
        if (op <= 2) ws_free(&w);
    }
    return any;
}

/* batch to_newick: string b at out + b*stride (NUL-terminated), lengths[b] without the NUL */
int p2vo_to_newick_batch(const i64 *v, i64 B, i64 k, char *out, i64 stride, i64 *lengths, int nthreads) {
    int bad = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 16) num_threads(nthreads) reduction(| : bad)
#endif
    for (i64 b = 0; b < B; ++b) {
        i64 len = p2vo_to_newick(v + b * k, k, out + b * stride, stride);
        if (len < 0 || len >= stride) bad |= 1;
        if (lengths) lengths[b] = len;
    }
    return bad ? P2VO_EBADINPUT : 0;
}

/* batch cophenetic (literal algorithm), out (B, n, n); bls may be NULL */
int p2vo_cophenetic_batch(const i64 *v, const double *bls, i64 B, i64 k, int unrooted,
                          double *out, int nthreads) {
    i64 n = k + 1;
    int any = 0;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
    for (i64 b = 0; b < B; ++b) {
        int rc = p2vo_cophenetic_literal(v + b * k, bls ? bls + b * 2 * k : NULL, k, unrooted,
                                         out + b * n * n);
        if (rc) any = 1;
    }
    return any;
}
