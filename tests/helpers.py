"""Shared test helpers: seeded input generation (the reference's
sample_vector/sample_matrix are unseeded, so the harness owns the seeds) and an
independent, definition-level tree-distance check."""

import numpy as np


def sample_vectors(B, n_leaves, seed, ordered=False):
    """v[b,i] ~ UniformInt{0..2i} (or {0..i}); phylo2vec/src/vector/base.rs:19-34."""
    rng = np.random.Generator(np.random.Philox(seed))
    k = n_leaves - 1
    hi = (np.arange(k, dtype=np.int64) * (1 if ordered else 2) + 1)
    return (rng.random((B, k)) * hi).astype(np.int64)


def sample_bls(B, n_leaves, seed):
    """bls ~ U[0,1); phylo2vec/src/matrix/base.rs:30-37."""
    rng = np.random.Generator(np.random.Philox(seed + 7919))
    return rng.random((B, n_leaves - 1, 2))


def to_matrix_input(v, bls):
    """(k,3) f64 = [v, bl1, bl2]; phylo2vec/src/matrix/base.rs:42-48."""
    return np.concatenate([v[..., None].astype(np.float64), bls], axis=-1)


def adversarial_vectors(n_leaves):
    k = n_leaves - 1
    i = np.arange(k, dtype=np.int64)
    out = {
        "ladder_zero": np.zeros(k, dtype=np.int64),
        "two_i": 2 * i,                      # phylo2vec/src/profile_main.rs:44
        "identity": i.copy(),
        "i_plus_1": np.minimum(i + 1, 2 * i),
        "alt": np.where(i % 2 == 0, 0, 2 * i),
    }
    return out


def tree_distances_by_definition(ancestry, bls, unrooted):
    """Leaf-to-leaf path lengths straight from the tree definition: build the
    parent map, walk both leaves to the root, sum the disjoint edges.  Mirrors
    what the reference's Python tests obtain from ete4
    (py-phylo2vec/tests/test_stats.py:28-68): 'unrooted' deletes the
    higher-numbered child of the root, dropping its branch length."""
    anc = np.asarray(ancestry)
    k = anc.shape[0]
    n = k + 1
    if bls is None:
        bls = np.ones((k, 2))
    par = {}
    ln = {}
    for r in range(k):
        c1, c2, p = (int(x) for x in anc[r])
        par[c1] = p
        par[c2] = p
        ln[c1] = float(bls[r][0])
        ln[c2] = float(bls[r][1])
    root = 2 * k
    if unrooted and k >= 2:
        a, b = int(anc[k - 1][0]), int(anc[k - 1][1])
        ln[max(a, b)] = 0.0
    d = np.zeros((n, n))
    paths = []
    for leaf in range(n):
        path = {}
        x, acc = leaf, 0.0
        while True:
            path[x] = acc
            if x == root:
                break
            acc += ln[x]
            x = par[x]
        paths.append(path)
    for i in range(n):
        for j in range(i):
            x, acc = j, 0.0
            while x not in paths[i]:
                acc += ln[x]
                x = par[x]
            d[i, j] = d[j, i] = acc + paths[i][x]
    return d
