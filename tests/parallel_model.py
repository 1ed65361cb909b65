"""Executable model of the data-parallel algorithms the CUDA kernels implement.

Pure Python/numpy, no CUDA: it exists so that the *derivations* behind the
kernels (DESIGN.md "Decode", "Encode", "Distance rows") are checked against the
oracle on the CPU, where there is no GPU.  tests/test_parallel_model.py runs it.
It mirrors the kernel phases one to one (same batch size, same formulas) but is
not used by the product.
"""

import numpy as np

W = 32  # batch width == warp width


# --------------------------------------------------------------------------
# Decode: v -> slot order -> pairs / ancestry / edges
# --------------------------------------------------------------------------
def slot_assignment(v):
    """slot_of[t] for every element t (element t <-> leaf t+1, pair (., t+1)).

    Insertion rank ins_t = 0 if v[t] <= t else v[t]-t (avl.rs:41-47: rank 0, or
    index+1 = v[t]-(t+1)+1).  Going backwards in time, element t ends up in the
    ins_t-th slot that later elements left free.  Batched: for a batch of W
    consecutive times [a, a+W) the rank R_t of element t among elements
    [0, a+W) follows from a forward scan inside the batch
    (R = ins_t; for j in t+1..a+W-1: R += ins_j <= R); all W selects then run
    against the same free mask."""
    k = len(v)
    ins = np.where(v <= np.arange(k), 0, v - np.arange(k))
    free = np.ones(k, dtype=bool)
    slot_of = np.full(k, -1, dtype=np.int64)
    nb = (k + W - 1) // W
    for b in range(nb - 1, -1, -1):
        a, e = b * W, min(k, b * W + W)
        R = ins[a:e].copy()
        for lane in range(e - a):           # lanes run this in lock step with shfl.down
            for d in range(1, e - a - lane):
                if ins[a + lane + d] <= R[lane]:
                    R[lane] += 1
        free_idx = np.flatnonzero(free)      # rank-select on the mask
        assert len(free_idx) == e
        s = free_idx[R]
        assert len(set(s.tolist())) == e - a
        slot_of[a:e] = s
        free[s] = False
    return slot_of, ins


def decode_model(v):
    v = np.asarray(v, dtype=np.int64)
    k = len(v)
    if k == 0:
        z = np.zeros((0, 2), dtype=np.int64)
        return z, np.zeros((0, 3), dtype=np.int64), z
    slot_of, ins = slot_assignment(v)
    time_at = np.empty(k, dtype=np.int64)
    time_at[slot_of] = np.arange(k)
    is_root_t = ins == 0                     # inserted at rank 0 with its own label v[t]
    # pairs: label of a slot = v[time] of the nearest root slot at or left of it
    root_slot = np.where(is_root_t[time_at], np.arange(k), -1)
    seg_root = np.maximum.accumulate(root_slot)       # segmented broadcast (max-scan)
    assert seg_root[0] == 0
    label = v[time_at[seg_root]]
    pairs = np.stack([label, time_at + 1], axis=1)

    # ancestry without the sequential parents[] pass (convert.rs:174-185):
    #  * roots appear in slot order by DEcreasing time, so the segment of root j
    #    ends right before the slot of the previous root in time.
    #  * c1 of a non-root row is the previous row's parent (same label, same segment).
    #  * c1 of a root row with label L: last earlier row carrying L == last row of the
    #    segment of the next root in time with the same v value, else leaf L.
    #  * c2 = leaf t+1: last earlier row carrying it == last row of the segment of
    #    the first root j (in time) with v[j] == t+1, else leaf t+1.
    roots_t = np.flatnonzero(is_root_t)
    prev_root = np.full(k, -1, dtype=np.int64)
    prev_root[roots_t[1:]] = roots_t[:-1]
    seg_end_node = np.full(k, -1, dtype=np.int64)     # by root time j: k+1+(last slot of its segment)
    for j in roots_t:
        end = k - 1 if prev_root[j] < 0 else slot_of[prev_root[j]] - 1
        seg_end_node[j] = k + 1 + end
    next_occ = np.full(k, -1, dtype=np.int64)         # next root time with the same v value
    first_root = np.full(k + 1, -1, dtype=np.int64)   # by value
    for t in range(k - 1, -1, -1):
        if is_root_t[t]:
            next_occ[t] = first_root[v[t]]
            first_root[v[t]] = t
    anc = np.empty((k, 3), dtype=np.int64)
    for s in range(k):
        t = time_at[s]
        if is_root_t[t]:
            j = next_occ[t]
            c1p = seg_end_node[j] if j >= 0 else v[t]
        else:
            c1p = k + s
        j = first_root[t + 1] if t + 1 <= k else -1
        c2p = seg_end_node[j] if j >= 0 else t + 1
        anc[s] = (c1p, c2p, k + 1 + s)
    edges = np.empty((2 * k, 2), dtype=np.int64)
    edges[0::2, 0] = anc[:, 0]
    edges[1::2, 0] = anc[:, 1]
    edges[0::2, 1] = anc[:, 2]
    edges[1::2, 1] = anc[:, 2]
    return pairs, anc, edges


# --------------------------------------------------------------------------
# Encode: cherries [c1,c2,cmax] (row order) -> v    (convert.rs:320-337)
# --------------------------------------------------------------------------
def build_vector_model(cherries):
    ch = np.asarray(cherries, dtype=np.int64).reshape(-1, 3)
    k = len(ch)
    row_of = np.full(k + 1, -1, dtype=np.int64)
    row_of[ch[:, 2]] = np.arange(k)
    assert (row_of[1:] >= 0).all(), "c_max must be a permutation of 1..k"
    seen = np.zeros(k, dtype=bool)          # rows whose value is < current batch
    v = np.zeros(k, dtype=np.int64)
    for a in range(1, k + 1, W):
        e = min(k + 1, a + W)
        rows = row_of[a:e]
        idx = np.array([seen[:r].sum() for r in rows])           # rank query on the mask
        for lane in range(e - a):                                 # in-batch pairs
            idx[lane] += sum(1 for l2 in range(lane) if rows[l2] < rows[lane])
        for lane, c in enumerate(range(a, e)):
            r = rows[lane]
            v[c - 1] = min(ch[r, 0], ch[r, 1]) if idx[lane] == 0 else c - 1 + idx[lane]
        seen[rows] = True
    return v


def build_vector_rows_model(cherries):
    """The round-2 form (p2v_encode_light.cuh): the same index taken in ROW order, 32 rows at a time --
    idx(row) = #{earlier rows with a smaller c_max} = the c_max values already seen below this one (prefix count
    on a "values seen" mask, state before the batch) + the lower lanes of the batch with a smaller c_max.  No
    table of rows by value; the mask doubles as the permutation check."""
    ch = np.asarray(cherries, dtype=np.int64).reshape(-1, 3)
    k = len(ch)
    seen = np.zeros(k + 1, dtype=bool)                           # by value
    v = np.zeros(k, dtype=np.int64)
    for a in range(0, k, W):
        e = min(k, a + W)
        cm = ch[a:e, 2]
        assert ((cm >= 1) & (cm <= k)).all() and not seen[cm].any() and len(set(cm.tolist())) == len(cm), "not a permutation"
        idx = np.array([seen[:c].sum() for c in cm])             # prefix count on the mask
        for lane in range(e - a):                                 # in-batch: lower lanes with a smaller value
            idx[lane] += sum(1 for l2 in range(lane) if cm[l2] < cm[lane])
        for lane in range(e - a):
            r = a + lane
            v[cm[lane] - 1] = min(ch[r, 0], ch[r, 1]) if idx[lane] == 0 else cm[lane] - 1 + idx[lane]
        seen[cm] = True
    return v


def min_desc_cherries_model(anc):
    """build_cherries (convert.rs:399-440): sequential semantics."""
    anc = np.asarray(anc, dtype=np.int64).reshape(-1, 3)
    k = len(anc)
    md = np.full(2 * k + 2, -1, dtype=np.int64)
    out = np.empty_like(anc)
    for r in range(k):
        c1, c2, p = anc[r]
        m1 = md[c1] if md[c1] >= 0 else c1
        m2 = md[c2] if md[c2] >= 0 else c2
        md[p] = min(m1, m2)
        out[r] = (m1, m2, max(m1, m2))
    return out


# --------------------------------------------------------------------------
# Distance rows: tables + blocked range-minimum scheme of the row kernel
# --------------------------------------------------------------------------
def tree_tables(v, bls, unrooted):
    """Everything the row kernel reads, for one tree with k >= 2.

    DFS order = the leaf order of the reference's Newick string (each row
    appends the absorbed leaf's whole sequence to the label's sequence,
    convert.rs:361-381).  sep[r] (r = 1..n-1) is the internal node that
    separates DFS leaves r-1 and r: it is the absorbing row of the leaf at DFS
    position r.  LCA(a,b) is then the shallowest sep in (a,b]."""
    v = np.asarray(v, dtype=np.int64)
    k = len(v)
    n = k + 1
    pairs, anc, _ = decode_model(v)
    N = 2 * k + 1
    par = np.full(N, -1, dtype=np.int64)
    wlen = np.zeros(N)
    for r in range(k):
        c1, c2, p = anc[r]
        par[c1] = par[c2] = p
        wlen[c1] = 1.0 if bls is None else bls[r][0]
        wlen[c2] = 1.0 if bls is None else bls[r][1]
    wint = np.ones(N, dtype=np.int64)
    wint[2 * k] = 0
    if unrooted:
        wlen[2 * k - 1] = 0.0
        wint[2 * k - 1] = 0
    # root-path sums by pointer jumping (log rounds)
    depth_i = wint.copy()
    depth_f = wlen.astype(np.longdouble)      # stands in for the kernel's double-double
    anc_ptr = par.copy()
    while (anc_ptr >= 0).any():
        has = anc_ptr >= 0
        ni, nf, na = depth_i.copy(), depth_f.copy(), anc_ptr.copy()
        ni[has] += depth_i[anc_ptr[has]]
        nf[has] += depth_f[anc_ptr[has]]
        na[has] = anc_ptr[anc_ptr[has]]
        depth_i, depth_f, anc_ptr = ni, nf, na
    # leaf order: linked list of leaves, then ranking
    nxt = np.full(n, -1, dtype=np.int64)
    tail = np.arange(n)                       # current tail of each label's sequence
    for s in range(k):
        c1, c2 = pairs[s]
        nxt[tail[c1]] = c2
        tail[c1] = tail[c2]
    order = []
    x = 0
    while x >= 0:
        order.append(x)
        x = nxt[x]
    leaf_at = np.array(order)
    assert len(leaf_at) == n
    pos = np.empty(n, dtype=np.int64)
    pos[leaf_at] = np.arange(n)
    # absorbing row of leaf L>=1 is the slot of element L-1 -> node k+1+slot
    slot_of, _ = slot_assignment(v)
    sep_node = np.full(n, -1, dtype=np.int64)          # indexed by DFS position r>=1
    sep_node[pos[1:]] = k + 1 + slot_of
    return dict(n=n, k=k, pos=pos, leaf_at=leaf_at, depth_i=depth_i, depth_f=depth_f,
                sep_node=sep_node, anc=anc)


def distance_rows_model(v, bls, unrooted, R=8, rows=None):
    """Blocked RMQ evaluation used by the row kernel.  Returns (n,n) float64
    (only `rows` filled if given).  key = topological depth of the separator;
    the LCA of DFS positions a<b is the separator with the smallest key in
    (a,b]; that minimum is assembled from
       row part   : suffix/prefix minimum inside the row's DFS block
       column part: min(prefix/suffix minimum inside the column's block,
                        minimum over the whole blocks in between)
    and an R x R table for columns inside the row's own block."""
    T = tree_tables(v, bls, unrooted)
    n, pos, leaf_at = T["n"], T["pos"], T["leaf_at"]
    BIG = 1 << 40
    topo = bls is None
    dep = T["depth_i"] if topo else T["depth_f"]
    key = np.full(n + 1, BIG, dtype=np.int64)           # key[r], r=1..n-1; key[0], key[n] = +inf
    key[1:n] = T["depth_i"][T["sep_node"][1:n]]
    kval = np.zeros(n + 1, dtype=dep.dtype)
    kval[1:n] = dep[T["sep_node"][1:n]]
    nb = (n + R - 1) // R
    # per-position within-block prefix / suffix minima (static tables)
    pre_k = np.full(n, BIG, dtype=np.int64); pre_v = np.zeros(n, dtype=dep.dtype)
    suf_k = np.full(n, BIG, dtype=np.int64); suf_v = np.zeros(n, dtype=dep.dtype)
    blk_k = np.full(nb, BIG, dtype=np.int64); blk_v = np.zeros(nb, dtype=dep.dtype)
    for b in range(nb):
        lo, hi = b * R, min(n, b * R + R)
        # prefix: min key[lo..p];   suffix: min key[p+1..hi-1]
        bk, bv = BIG, 0
        for p in range(lo, hi):
            if key[p] < bk:
                bk, bv = key[p], kval[p]
            pre_k[p], pre_v[p] = bk, bv
        blk_k[b], blk_v[b] = bk, bv
        bk, bv = BIG, 0
        for p in range(hi - 1, lo - 1, -1):
            suf_k[p], suf_v[p] = bk, bv
            if key[p] < bk:
                bk, bv = key[p], kval[p]
    out = np.zeros((n, n))
    for I in range(nb):
        lo, hi = I * R, min(n, I * R + R)
        # per tile: minima over whole blocks strictly between I and J
        mid_k = np.full(nb, BIG, dtype=np.int64); mid_v = np.zeros(nb, dtype=dep.dtype)
        bk, bv = BIG, 0
        for J in range(I + 1, nb):
            mid_k[J], mid_v[J] = bk, bv
            if blk_k[J] < bk:
                bk, bv = blk_k[J], blk_v[J]
        bk, bv = BIG, 0
        for J in range(I - 1, -1, -1):
            mid_k[J], mid_v[J] = bk, bv
            if blk_k[J] < bk:
                bk, bv = blk_k[J], blk_v[J]
        for j in range(n):                      # column setup, once per tile
            p = pos[j]
            J = p // R
            if J == I:
                continue
            if J > I:                           # column right of the tile: key[JR..p]
                ck, cv = pre_k[p], pre_v[p]
            else:                               # column left of the tile: key[p+1..(J+1)R-1]
                ck, cv = suf_k[p], suf_v[p]
            if mid_k[J] < ck:
                ck, cv = mid_k[J], mid_v[J]
            dj = dep[j]
            for q in range(lo, hi):             # the streamed part
                i = leaf_at[q]
                if rows is not None and i not in rows:
                    continue
                if J > I:
                    rk, rv = suf_k[q], suf_v[q]     # key[q+1..hi-1]
                else:
                    rk, rv = pre_k[q], pre_v[q]     # key[lo..q]
                lv = cv if ck < rk else rv
                out[i, j] = float((dep[i] - lv) + (dj - lv))
        # columns inside the tile's own block: direct minima
        for q in range(lo, hi):
            i = leaf_at[q]
            if rows is not None and i not in rows:
                continue
            for p in range(lo, hi):
                if p == q:
                    continue
                a, b = min(p, q), max(p, q)
                r = a + 1 + int(np.argmin(key[a + 1:b + 1]))
                j = leaf_at[p]
                out[i, j] = float((dep[i] - kval[r]) + (dep[j] - kval[r]))
    return out


# --------------------------------------------------------------------------
# Row kernel, current formulation: blocks sized by the REQUESTED rows, and
#   D = (d_i - d_rl) + (d_j - d_cl) + |d_rl - d_cl|
# with rl / cl the row's / column's own candidate (see p2v_cophenetic.cu).
# --------------------------------------------------------------------------
def distance_rows_model_v2(v, bls, unrooted, row_begin, row_end, R=4, PMAX=16):
    T = tree_tables(v, bls, unrooted)
    n, pos, leaf_at = T["n"], T["pos"], T["leaf_at"]
    INF = 1 << 40
    topo = bls is None
    dep = T["depth_i"] if topo else T["depth_f"]
    key = np.full(n + 1, INF, dtype=np.int64)
    key[1:n] = T["depth_i"][T["sep_node"][1:n]]
    val = np.zeros(n + 1, dtype=dep.dtype)
    val[1:n] = dep[T["sep_node"][1:n]]
    own = np.array([(row_begin <= leaf_at[p] < row_end) for p in range(n)], dtype=np.int64)
    c = np.cumsum(own)
    blk_of = np.arange(n) // PMAX + np.maximum(0, c - 1) // R
    nb = int(blk_of[-1]) + 1
    rowlist = np.flatnonzero(own)
    bstart = np.full(nb, -1); bend = np.full(nb, -1)
    for p in range(n):
        b = blk_of[p]
        if bstart[b] < 0:
            bstart[b] = p
        bend[b] = p + 1
    pre = [None] * n; suf = [None] * n; blk = [(INF, -1)] * nb
    for b in range(nb):
        if bstart[b] < 0:
            continue
        lo, hi = bstart[b], bend[b]
        assert hi - lo <= PMAX and own[lo:hi].sum() <= R
        best = (INF, -1)
        for p in range(lo, hi):
            best = min(best, (key[p], p))
            pre[p] = best
        blk[b] = best
        best = (INF, -1)
        for p in range(hi - 1, lo - 1, -1):
            suf[p] = best
            best = min(best, (key[p], p))
    out = np.full((row_end - row_begin, n), np.nan)
    for I in range(nb):
        if bstart[I] < 0:
            continue
        rows_q = [q for q in rowlist if blk_of[q] == I]
        if not rows_q:
            continue
        mid = [(INF, -1)] * nb
        best = (INF, -1)
        for J in range(I + 1, nb):
            mid[J] = best
            best = min(best, blk[J])
        best = (INF, -1)
        for J in range(I - 1, -1, -1):
            mid[J] = best
            best = min(best, blk[J])
        for q in rows_q:
            i = leaf_at[q]
            cand = {}
            for side, (ck, ca) in ((0, pre[q]), (1, suf[q])):
                if ck >= INF:
                    cand[side] = (dep[i], dep[i] - dep[i])      # candidate = the leaf itself
                else:
                    cand[side] = (val[ca], dep[i] - val[ca])
            for j in range(n):
                p = pos[j]
                J = blk_of[p]
                if J == I:                                       # in-block: direct range minimum
                    if p == q:
                        d = 0.0
                    else:
                        a, b = min(p, q), max(p, q)
                        r = a + 1 + int(np.argmin(key[a + 1:b + 1]))
                        d = float((dep[i] - val[r]) + (dep[j] - val[r]))
                else:
                    right = J > I
                    ck, ca = pre[p] if right else suf[p]
                    if mid[J][0] < ck:
                        ck, ca = mid[J]
                    if ck >= INF:
                        B, uj = dep[j], dep[j] - dep[j]
                    else:
                        B, uj = val[ca], dep[j] - val[ca]
                    A, ui = cand[1 if right else 0]
                    d = float(ui + uj + abs(A - B))
                out[i - row_begin, j] = d
    assert not np.isnan(out).any()
    return out


# --------------------------------------------------------------------------
# Decode, large trees: merge-based slot assignment (decode_large_kernel) and the
# forward relabel sweep used by both decode kernels.
# --------------------------------------------------------------------------
def slot_assignment_merge(v):
    """Level 0: rank inside each batch of W, stored sorted by rank.  Then pairwise merges:
    for X (earlier times) | Y (later times), Y's ranks are final for the merged block; an X
    element of rank r gains #{l : y_l - l <= r}; a Y element (index l, rank y) is preceded
    by #{x : r_x < y - l} X elements."""
    v = np.asarray(v, dtype=np.int64)
    k = len(v)
    ins = np.where(v <= np.arange(k), 0, v - np.arange(k))
    rk = np.zeros(k, dtype=np.int64)
    el = np.zeros(k, dtype=np.int64)
    for a in range(0, k, W):
        e = min(k, a + W)
        R = ins[a:e].copy()
        for lane in range(e - a):
            for d in range(1, e - a - lane):
                if ins[a + lane + d] <= R[lane]:
                    R[lane] += 1
        order = np.argsort(R, kind="stable")
        rk[a:e] = R[order]
        el[a:e] = np.arange(a, e)[order]
    S = W
    while S < k:
        nrk, nel = rk.copy(), el.copy()
        for a in range(0, k, 2 * S):
            m = a + S
            if m >= k:
                continue
            e = min(k, a + 2 * S)
            X, Y = rk[a:m], rk[m:e]
            z = Y - np.arange(len(Y))
            for ix in range(len(X)):
                cnt = int(np.searchsorted(z, X[ix], side="right"))
                nrk[a + ix + cnt] = X[ix] + cnt
                nel[a + ix + cnt] = el[a + ix]
            for l in range(len(Y)):
                cntx = int(np.searchsorted(X, z[l], side="left"))
                nrk[a + l + cntx] = Y[l]
                nel[a + l + cntx] = el[m + l]
        rk, el = nrk, nel
        S *= 2
    assert np.array_equal(rk, np.arange(k))
    slot_of = np.empty(k, dtype=np.int64)
    slot_of[el] = np.arange(k)
    return slot_of


def ancestry_forward_model(v):
    """The reference's parents[] pass (convert.rs:174-185) as a forward sweep over slots, W
    rows at a time, with one table: the last row so far whose label is x.  Non-root rows point
    at the previous row; a root row looks its label up (same-batch occurrences first); the
    absorbed leaf t+1 is looked up after the batch's updates (all rows labelled t+1 lie left)."""
    v = np.asarray(v, dtype=np.int64)
    k = len(v)
    slot_of, _ = slot_assignment(v)
    time_at = np.empty(k, dtype=np.int64)
    time_at[slot_of] = np.arange(k)
    tbl = np.full(k + 2, -1, dtype=np.int64)
    anc = np.empty((k, 3), dtype=np.int64)
    carry_label = 0
    for a in range(0, k, W):
        e = min(k, a + W)
        t = time_at[a:e]
        x = v[t]
        root = x <= t
        lab = np.empty(e - a, dtype=np.int64)
        for l in range(e - a):
            if root[l]:
                carry_label = x[l]
            lab[l] = carry_label
        pre = tbl[lab].copy()
        c1 = np.empty(e - a, dtype=np.int64)
        for l in range(e - a):
            if root[l]:
                lower = [l2 for l2 in range(l) if lab[l2] == lab[l]]
                cand = a + max(lower) if lower else pre[l]
                c1[l] = k + 1 + cand if cand >= 0 else lab[l]
            else:
                c1[l] = k + a + l
        for l in range(e - a):
            tbl[lab[l]] = max(tbl[lab[l]], a + l)
        for l in range(e - a):
            q = tbl[t[l] + 1]
            anc[a + l] = (c1[l], k + 1 + q if q >= 0 else t[l] + 1, k + 1 + a + l)
    return anc


# --------------------------------------------------------------------------
# Robinson-Foulds by Day's algorithm (p2v_treewise.cu; reference: vector/distance.rs:41-162)
# --------------------------------------------------------------------------
def robinson_foulds_day_model(v1, v2):
    """Leaves renumbered by their DFS rank in tree 1, rotated so that leaf 0 has rank 0; a split of tree 1 is the
    interval of ranks on the side WITHOUT leaf 0; tree 2 supplies (min, max, count) of the ranks below every node and of
    everything outside it; a contiguous side without leaf 0 is looked up.  RF = |S1| + |S2| - 2 shared."""
    from oracle import p2v_oracle as O
    n = len(v1) + 1
    if n < 4:
        return 0.0
    k = n - 1

    def children(v):
        anc = O.to_ancestry(np.asarray(v))
        ch = {}
        for c1, c2, p in anc.tolist():
            ch[p] = (c1, c2)
        return ch, 2 * k

    def below(ch, root):                                          # leaves below every node, iteratively
        leaves = {x: [x] for x in range(n)}
        for p in sorted(ch):                                      # parents are created after their children
            leaves[p] = leaves[ch[p][0]] + leaves[ch[p][1]]
        return leaves

    ch1, root1 = children(v1)
    order = below(ch1, root1)[root1]                              # DFS order of tree 1
    at0 = order.index(0)
    rank = {leaf: (i - at0) % n for i, leaf in enumerate(order)}

    def splits(ch, root, contiguous_only):
        out = set()
        for p, leaves in below(ch, root).items():
            if p == root:
                continue
            r = sorted(rank[x] for x in leaves)
            side = r if 0 not in r else sorted(set(range(n)) - set(r))      # the side without leaf 0
            if len(side) <= 1 or len(side) >= n - 1:
                continue                                          # trivial split
            if side[-1] - side[0] + 1 == len(side):
                out.add(("interval", side[0], side[-1]))
            elif not contiguous_only:
                out.add(("set",) + tuple(side))
        return out

    s1 = splits(ch1, root1, False)
    assert all(t[0] == "interval" for t in s1)                   # in tree 1's own order every split is an interval
    ch2, root2 = children(v2)
    s2_all = splits(ch2, root2, False)
    shared = len(s1 & s2_all)
    return float(len(s1) + len(s2_all) - 2 * shared)
