#!/usr/bin/env python
"""Transcription of the reference's own known-answer tests for the hot path.

The Rust reference cannot be compiled or imported in this image (no cargo /
rustc / wheel), so these vectors are copied BY HAND from the reference's test
tables and doctests; each entry cites the file:line (relative to
/root/reference) it was read from.  Running this script rewrites
tests/golden/reference_goldens.json; it reads nothing from /root/reference.
"""

import json
import os

G = {
    # ---------------------------------------------------------------- decode
    "to_pairs": [
        {"v": [0, 0, 0, 1], "pairs": [[1, 4], [0, 3], [0, 2], [0, 1]],
         "src": "phylo2vec/src/vector/convert.rs:99-101"},
        {"v": [0, 2, 2, 5, 4, 1], "pairs": [[1, 6], [4, 5], [2, 3], [0, 1], [0, 4], [0, 2]],
         "src": "docs/demo.ipynb cell 30 (to_pairs(v_fixed))"},
    ],
    "to_ancestry": [
        {"v": [0, 2, 2], "ancestry": [[2, 3, 4], [0, 1, 5], [5, 4, 6]],
         "src": "phylo2vec/src/vector/convert.rs:159-165"},
        {"v": [0, 0, 0, 1, 3], "ancestry": [[3, 5, 6], [1, 4, 7], [0, 6, 8], [8, 2, 9], [9, 7, 10]],
         "src": "phylo2vec/src/tree_vec/mod.rs:162-166"},
        {"v": [0, 1, 2, 3], "ancestry": [[3, 4, 5], [2, 5, 6], [1, 6, 7], [0, 7, 8]],
         "src": "phylo2vec/src/tree_vec/mod.rs:167-170"},
        {"v": [0, 0, 1], "ancestry": [[1, 3, 4], [0, 2, 5], [5, 4, 6]],
         "src": "phylo2vec/src/tree_vec/mod.rs:171-173"},
        {"v": [0, 2, 2, 5, 4, 1],
         "ancestry": [[1, 6, 7], [4, 5, 8], [2, 3, 9], [0, 7, 10], [10, 8, 11], [11, 9, 12]],
         "src": "docs/demo.ipynb cell 26 (to_ancestry(v_fixed))"},
    ],
    "to_edges": [
        {"v": [0, 2, 2], "edges": [[2, 4], [3, 4], [0, 5], [1, 5], [5, 6], [4, 6]],
         "src": "phylo2vec/src/vector/convert.rs:223-225"},
        {"v": [0, 0, 0, 1, 3],
         "edges": [[3, 6], [5, 6], [1, 7], [4, 7], [0, 8], [6, 8], [8, 9], [2, 9], [9, 10], [7, 10]],
         "src": "phylo2vec/src/tree_vec/mod.rs:181-191"},
        {"v": [0, 1, 2, 3],
         "edges": [[3, 5], [4, 5], [2, 6], [5, 6], [1, 7], [6, 7], [0, 8], [7, 8]],
         "src": "phylo2vec/src/tree_vec/mod.rs:192-200"},
        {"v": [0, 0, 1], "edges": [[1, 4], [3, 4], [0, 5], [2, 5], [5, 6], [4, 6]],
         "src": "phylo2vec/src/tree_vec/mod.rs:201-207"},
        {"v": [0, 2, 2, 5, 4, 1],
         "edges": [[1, 7], [6, 7], [4, 8], [5, 8], [2, 9], [3, 9], [0, 10], [7, 10], [10, 11],
                   [8, 11], [11, 12], [9, 12]],
         "src": "docs/demo.ipynb cell 22 (to_edges(v_fixed))"},
    ],
    # Newick strings are a pure function of to_pairs (convert.rs:394-397), so
    # they pin the pair order and labels as well.
    "to_newick": [
        {"v": [0, 2, 2], "newick": "((0,1)5,(2,3)4)6;", "src": "phylo2vec/src/vector/convert.rs:389-393"},
        {"v": [0, 0, 0, 1, 3], "newick": "(((0,(3,5)6)8,2)9,(1,4)7)10;",
         "src": "phylo2vec/src/vector/convert.rs:662; tree_vec/mod.rs:149"},
        {"v": [0, 1, 2, 3, 4], "newick": "(0,(1,(2,(3,(4,5)6)7)8)9)10;",
         "src": "phylo2vec/src/vector/convert.rs:663; tree_vec/mod.rs:150"},
        {"v": [0, 0, 1], "newick": "((0,2)5,(1,3)4)6;",
         "src": "phylo2vec/src/vector/convert.rs:664; tree_vec/mod.rs:151"},
    ],
    # from_newick (vectors): with parent labels (has_parents arm) and without (order_cherries_no_parents)
    "from_newick": [
        {"newick": "((0,1)5,(2,3)4)6;", "v": [0, 2, 2], "src": "phylo2vec/src/vector/convert.rs:259-262"},
        {"newick": "((0,1),(2,3));", "v": [0, 2, 2], "src": "phylo2vec/src/vector/convert.rs:264-266"},
        {"newick": "(((0,(3,5)6)8,2)9,(1,4)7)10;", "v": [0, 0, 0, 1, 3], "src": "phylo2vec/src/vector/convert.rs:662"},
        {"newick": "(0,(1,(2,(3,(4,5)6)7)8)9)10;", "v": [0, 1, 2, 3, 4], "src": "phylo2vec/src/vector/convert.rs:663"},
        {"newick": "((0,2)5,(1,3)4)6;", "v": [0, 0, 1], "src": "phylo2vec/src/vector/convert.rs:664"},
        {"newick": "(((0,(3,5)),2),(1,4));", "v": [0, 0, 0, 1, 3], "src": "phylo2vec/src/vector/convert.rs:674"},
        {"newick": "(0,(1,(2,(3,(4,5)))));", "v": [0, 1, 2, 3, 4], "src": "phylo2vec/src/vector/convert.rs:675"},
        {"newick": "((0,2),(1,3));", "v": [0, 0, 1], "src": "phylo2vec/src/vector/convert.rs:676"},
    ],
    "newick_parse": [
        {"newick": "((0,2)5,(1,3)4)6;", "ancestry": [[0, 2, 5], [1, 3, 4], [5, 4, 6]],
         "src": "phylo2vec/src/newick/mod.rs:61-67"},
    ],
    # balance indices (vector/balance.rs tests :77-93 sackin, :103-123 variance, :139-156 b2; rtol 1e-5 for the floats)
    "balance": [
        {"v": [0], "sackin": 2, "variance": 0.0, "b2": 1.0, "src": "phylo2vec/src/vector/balance.rs:79,105,141"},
        {"v": [0, 2, 2, 3], "sackin": 12, "variance": 0.24, "b2": 2.25, "src": "balance.rs:80,106,142"},
        {"v": [0, 0], "sackin": 5, "variance": 0.2222222, "b2": 1.5, "src": "balance.rs:82,112,144"},
        {"v": [0, 0, 0], "sackin": 9, "variance": 0.6875, "b2": 1.75, "src": "balance.rs:83,113,145"},
        {"v": [0, 0, 0, 0], "sackin": 14, "variance": 1.36, "b2": 1.875, "src": "balance.rs:84,114,146"},
        {"v": [0, 0, 0, 0, 0], "b2": 1.9375, "src": "balance.rs:147"},
        {"v": [0] * 49, "sackin": 1274, "variance": 207.2896, "b2": 2.0, "src": "balance.rs:85,115,148"},
        {"v": [0] * 99, "sackin": 5049, "b2": 2.0, "src": "balance.rs:86,149"},
        {"v": [0, 2, 2], "sackin": 8, "variance": 0.0, "b2": 2.0, "src": "balance.rs:88,108,151"},
        {"v": [0, 2, 2, 6, 4, 6, 6], "sackin": 24, "variance": 0.0, "b2": 3.0, "src": "balance.rs:89,109,152"},
        {"v": [0, 2, 2, 6, 4, 6, 6, 14, 8, 10, 10, 14, 12, 14, 14], "sackin": 64, "variance": 0.0, "b2": 4.0,
         "src": "balance.rs:90,110,153"},
    ],
    # ---------------------------------------------------------------- encode
    "from_pairs": [
        {"pairs": [[1, 4], [0, 3], [0, 2], [0, 1]], "v": [0, 0, 0, 1],
         "src": "phylo2vec/src/vector/convert.rs:19-21"},
    ],
    "from_ancestry": [
        {"ancestry": [[2, 3, 4], [0, 1, 5], [5, 4, 6]], "v": [0, 2, 2],
         "src": "phylo2vec/src/vector/convert.rs:121-127"},
    ],
    "from_edges": [
        {"edges": [[2, 4], [3, 4], [0, 5], [1, 5], [4, 6], [5, 6]], "v": [0, 2, 2],
         "src": "phylo2vec/src/vector/convert.rs:197-199"},
    ],
    "order_cherries": [
        {"in": [[0, 3, 6], [0, 2, 5], [0, 1, 4]], "out": [[0, 1, 1], [0, 2, 2], [0, 3, 3]],
         "src": "phylo2vec/src/vector/convert.rs:700-709"},
        {"in": [[0, 1, 5], [0, 3, 4], [0, 2, 6]], "out": [[0, 3, 3], [0, 1, 1], [0, 2, 2]],
         "src": "phylo2vec/src/vector/convert.rs:710-719"},
    ],
    "build_vector": [
        {"cherries": [[0, 3, 3], [0, 2, 2], [0, 1, 1]], "v": [0, 0, 0],
         "src": "phylo2vec/src/vector/convert.rs:755-760"},
        {"cherries": [[2, 3, 3], [1, 2, 2], [0, 1, 1]], "v": [0, 1, 2],
         "src": "phylo2vec/src/vector/convert.rs:761-766"},
        {"cherries": [[6, 7, 7], [6, 9, 9], [0, 6, 6], [1, 5, 5], [1, 4, 4], [0, 3, 3], [0, 2, 2],
                      [0, 8, 8], [0, 1, 1]],
         "v": [0, 0, 0, 1, 1, 0, 6, 13, 9],
         "src": "phylo2vec/src/vector/convert.rs:767-779"},
    ],
    # ------------------------------------------------------------ validation
    # Robinson-Foulds: the reference's own assertions (phylo2vec/src/vector/distance.rs:168-239).  "expected"
    # is a number, or a property: "symmetric", "unit_interval" (normalised), "range_0_2".
    # (The docstring example in py-phylo2vec/phylo2vec/stats/treewise.py:46-51 prints 2.0 for
    # [0,1,2,3] vs [0,0,1,2]; the two 5-leaf trees share no bipartition, so the code -- and ete4, which the
    # reference's tests compare with, tests/test_stats.py:283-310 -- give 2 (n - 3) = 4: not used as a golden.)
    "robinson_foulds": [
        {"v1": [0, 1, 2, 3], "v2": [0, 1, 2, 3], "normalize": False, "expected": 0.0, "src": "distance.rs:169-172"},
        {"v1": [0, 1, 2, 3, 4, 5, 6], "v2": [0, 1, 2, 3, 4, 5, 6], "normalize": False, "expected": 0.0,
         "src": "distance.rs:175-178"},
        {"v1": [0, 1, 2, 3], "v2": [0, 0, 1, 2], "normalize": False, "expected": "symmetric", "src": "distance.rs:181-188"},
        {"v1": [0, 1, 2, 3], "v2": [0, 0, 1, 2], "normalize": True, "expected": "unit_interval", "src": "distance.rs:191-196"},
        {"v1": [0, 0], "v2": [0, 1], "normalize": False, "expected": 0.0, "src": "distance.rs:199-204"},
        {"v1": [0, 0, 0], "v2": [0, 1, 0], "normalize": False, "expected": "range_0_2", "src": "distance.rs:207-217"},
        {"v1": [0, 1, 2], "v2": [0, 1, 2], "normalize": False, "expected": 0.0, "src": "distance.rs:33-34 (doctest)"},
    ],
    # matrices <-> Newick with branch lengths: phylo2vec/src/matrix/convert.rs doctests and test tables
    "to_newick_matrix": [
        {"m": [[0.0, 0.5, 0.8], [1.0, 0.7, 0.6]], "newick": "(0:0.7,(1:0.5,2:0.8)3:0.6)4;", "src": "matrix/convert.rs:18-24"},
        {"m": [[0.0, 0.9, 0.4], [0.0, 0.8, 3.12], [3.0, 0.4, 0.5]], "newick": "(((0:0.9,2:0.4)4:0.8,3:3.12)5:0.4,1:0.5)6;",
         "src": "matrix/convert.rs:191-196"},
        {"m": [[0.0, 0.1, 0.2]], "newick": "(0:0.1,1:0.2)2;", "src": "matrix/convert.rs:197"},
        {"m": [[0.0, 1.0, 3.0], [0.0, 0.1, 0.2], [1.0, 0.5, 0.7]], "newick": "((0:0.1,2:0.2)5:0.5,(1:1.0,3:3.0)4:0.7)6;",
         "src": "matrix/convert.rs:198-202"},
    ],
    "from_newick_matrix": [
        {"newick": "(0:0.1,1:0.2)2;", "m": [[0.0, 0.1, 0.2]], "src": "matrix/convert.rs:73-78, 133"},
        {"newick": "(1:0.2,0:0.1)2;", "m": [[0.0, 0.1, 0.2]], "src": "matrix/convert.rs:135 (bl_rows_to_swap)"},
        {"newick": "(((0:0.9,2:0.4)4:0.8,3:3.0)5:0.4,1:0.5)6;", "m": [[0.0, 0.9, 0.4], [0.0, 0.8, 3.0], [3.0, 0.4, 0.5]],
         "src": "matrix/convert.rs:136-140"},
        {"newick": "(0:0.7,(1:0.5,2:0.8)3:0.6)4;", "m": [[0.0, 0.5, 0.8], [1.0, 0.7, 0.6]], "src": "matrix/convert.rs:141-144"},
        {"newick": "((((0:0.30118185,3:0.69665915)5:0.69915295,4:0.059750594)6:0.0017238181,1:0.34410468)7:0.2021168,2:0.7084421)8;",
         "m": [[0.0, 0.30118185, 0.69665915], [2.0, 0.69915295, 0.059750594], [0.0, 0.0017238181, 0.34410468],
               [4.0, 0.2021168, 0.7084421]], "src": "matrix/convert.rs:145-150"},
        {"newick": "(0:0.5,1:0.6);", "m": [[0.0, 0.5, 0.6]], "src": "matrix/convert.rs:160-162 (no parents)"},
        {"newick": "((0:0.1,2:0.2):0.3,(1:0.5,3:0.7):0.4);", "m": [[0.0, 0.5, 0.7], [0.0, 0.1, 0.2], [1.0, 0.3, 0.4]],
         "src": "matrix/convert.rs:163-167 (no parents)"},
    ],
    # test_check_m, phylo2vec/src/matrix/base.rs:131-170 (the empty-array case is a shape check)
    "check_m": {
        "ok": [
            [[0.0, 0.1, 0.4], [0.0, 0.2, 0.5], [1.0, 0.3, 0.6]],
            [[0.0, 0.33, 0.72], [0.0, 0.61, 0.14], [2.0, 0.90, 0.45], [1.0, 0.12, 0.34], [8.0, 0.78, 0.23]],
            [[0.0, 0.0, 0.0], [0.0, 0.1, 0.2], [1.0, 0.5, 0.7]],
        ],
        "panic": [
            {"m": [[0.0, 0.0, 9.0, 1.0], [0.1, 0.2, 0.3, 0.7], [0.4, 0.5, 0.6, 0.8]], "why": "columns"},
            {"m": [[0.0, 0.0, 0.1, 0.2], [0.0, 0.3, 0.4, 0.5], [9.0, 0.6, 0.7, 0.8], [1.0, 0.9, 1.0, 1.1]], "why": "v"},
            {"m": [[0.0, 0.33, 0.72], [0.0, 0.61, 0.14], [2.0, 0.90, 0.45], [1.0, 0.12, 0.34], [8.0, -0.78, 0.23]],
             "why": "negative"},
        ],
        "src": "phylo2vec/src/matrix/base.rs:131-170; doctest :70-75",
    },
    "check_v": {
        "ok": [[0, 0, 1], [0, 0, 2, 1, 8], [0, 2, 4, 6]],
        "panic": [[0, 0, 9, 1], [1, 0, 0, 0], [0, 3, 5, 7]],
        "src": "phylo2vec/src/vector/base.rs:120-132",
    },
    "is_unordered": [
        {"v": [0, 0, 0, 1, 3, 3, 1, 4, 4], "expected": False, "src": "phylo2vec/src/vector/base.rs:135"},
        {"v": [0, 0, 0, 3, 2, 9, 4, 1, 12], "expected": True, "src": "phylo2vec/src/vector/base.rs:136"},
        {"v": [0, 0, 1, 10, 2, 9, 4, 1, 12], "expected": "panic", "src": "phylo2vec/src/vector/base.rs:137-138"},
    ],
    # ------------------------------------------- cophenetic, topological (exact)
    "cophenetic_vector": [
        {"v": [], "unrooted": False, "d": [[0.0]], "src": "phylo2vec/src/vector/graph.rs:562"},
        {"v": [0], "unrooted": False, "d": [[0, 2], [2, 0]], "src": "phylo2vec/src/vector/graph.rs:563"},
        {"v": [0], "unrooted": True, "d": [[0, 2], [2, 0]], "src": "phylo2vec/src/vector/graph.rs:564"},
        {"v": [0, 0, 1], "unrooted": False,
         "d": [[0, 4, 2, 4], [4, 0, 4, 2], [2, 4, 0, 4], [4, 2, 4, 0]],
         "src": "phylo2vec/src/vector/graph.rs:565-570"},
        {"v": [0, 0, 1], "unrooted": True,
         "d": [[0, 3, 2, 3], [3, 0, 3, 2], [2, 3, 0, 3], [3, 2, 3, 0]],
         "src": "phylo2vec/src/vector/graph.rs:571-576"},
        {"v": [0, 1, 2], "unrooted": False,
         "d": [[0, 3, 4, 4], [3, 0, 3, 3], [4, 3, 0, 2], [4, 3, 2, 0]],
         "src": "phylo2vec/src/vector/graph.rs:577-582 (also doctest :102-105)"},
        {"v": [0, 1, 2], "unrooted": True,
         "d": [[0, 2, 3, 3], [2, 0, 3, 3], [3, 3, 0, 2], [3, 3, 2, 0]],
         "src": "phylo2vec/src/vector/graph.rs:583-588"},
        {"v": [0, 1, 0, 4, 4, 2, 6], "unrooted": False,
         "d": [[0, 5, 6, 2, 4, 4, 7, 7],
               [5, 0, 3, 5, 5, 5, 4, 4],
               [6, 3, 0, 6, 6, 6, 3, 3],
               [2, 5, 6, 0, 4, 4, 7, 7],
               [4, 5, 6, 4, 0, 2, 7, 7],
               [4, 5, 6, 4, 2, 0, 7, 7],
               [7, 4, 3, 7, 7, 7, 0, 2],
               [7, 4, 3, 7, 7, 7, 2, 0]],
         "src": "phylo2vec/src/vector/graph.rs:589-598"},
        {"v": [0, 2, 2, 5, 4, 1], "unrooted": False,
         "d": [[0, 3, 5, 5, 4, 4, 3],
               [3, 0, 6, 6, 5, 5, 2],
               [5, 6, 0, 2, 5, 5, 6],
               [5, 6, 2, 0, 5, 5, 6],
               [4, 5, 5, 5, 0, 2, 5],
               [4, 5, 5, 5, 2, 0, 5],
               [3, 2, 6, 6, 5, 5, 0]],
         "src": "docs/demo.ipynb cell 40 (pairwise_distances(v_fixed))"},
    ],
    # ------------------- cophenetic with branch lengths (reference: allclose,
    # rtol 1e-5 / atol 1e-8, phylo2vec/src/matrix/graph.rs:99-109)
    "cophenetic_matrix": [
        {"m": [[0.0, 1.0, 10.0]], "unrooted": False, "d": [[0.0, 11.0], [11.0, 0.0]],
         "src": "phylo2vec/src/matrix/graph.rs:112"},
        {"m": [[0.0, 1.0, 10.0]], "unrooted": True, "d": [[0.0, 11.0], [11.0, 0.0]],
         "src": "phylo2vec/src/matrix/graph.rs:113"},
        {"m": [[0.0, 1.0, 3.0], [0.0, 0.1, 2.0], [1.0, 4.0, 5.0]], "unrooted": False,
         "d": [[0.0, 10.1, 2.1, 12.1], [10.1, 0.0, 12.0, 4.0], [2.1, 12.0, 0.0, 14.0],
               [12.1, 4.0, 14.0, 0.0]],
         "src": "phylo2vec/src/matrix/graph.rs:114-123"},
        {"m": [[0.0, 1.0, 3.0], [0.0, 0.1, 2.0], [1.0, 5.0, 4.0]], "unrooted": True,
         "d": [[0.0, 5.1, 2.1, 7.1], [5.1, 0.0, 7.0, 4.0], [2.1, 7.0, 0.0, 9.0], [7.1, 4.0, 9.0, 0.0]],
         "src": "phylo2vec/src/matrix/graph.rs:124-133"},
        {"m": [[0.0, 0.4, 0.5], [2.0, 0.1, 0.2], [2.0, 0.3, 0.6]], "unrooted": False,
         "d": [[0.0, 0.3, 1.4, 1.5], [0.3, 0.0, 1.5, 1.6], [1.4, 1.5, 0.0, 0.9], [1.5, 1.6, 0.9, 0.0]],
         "src": "phylo2vec/src/matrix/graph.rs:134-143"},
        {"m": [[0.0, 0.6, 0.2], [0.0, 0.1, 0.4], [2.0, 0.2, 0.6], [3.0, 0.8, 0.9]], "unrooted": False,
         "d": [[0.0, 1.9, 0.9, 1.8, 1.4], [1.9, 0.0, 2.4, 3.3, 2.9], [0.9, 2.4, 0.0, 1.1, 0.7],
               [1.8, 3.3, 1.1, 0.0, 0.8], [1.4, 2.9, 0.7, 0.8, 0.0]],
         "src": "phylo2vec/src/matrix/graph.rs:144-155"},
        {"m": [[0.0, 1.0, 0.2], [2.0, 0.9, 0.7], [1.0, 0.1, 0.2], [6.0, 0.8, 0.3]], "unrooted": False,
         "d": [[0.0, 2.6, 1.2, 1.8, 2.1], [2.6, 0.0, 2.0, 1.2, 2.9], [1.2, 2.0, 0.0, 1.2, 1.3],
               [1.8, 1.2, 1.2, 0.0, 2.1], [2.1, 2.9, 1.3, 2.1, 0.0]],
         "src": "phylo2vec/src/matrix/graph.rs:156-167"},
        {"m": [[0.0, 0.9, 0.9], [1.0, 0.5, 0.6], [0.0, 0.7, 0.9], [4.0, 0.6, 0.7], [4.0, 0.1, 0.1],
               [2.0, 0.7, 0.6], [6.0, 0.7, 0.3]], "unrooted": False,
         "d": [[0.0, 2.4, 2.8, 1.3, 1.5, 1.7, 3.8, 3.8],
               [2.4, 0.0, 1.8, 2.5, 2.5, 2.7, 2.8, 2.8],
               [2.8, 1.8, 0.0, 2.9, 2.9, 3.1, 2.0, 2.0],
               [1.3, 2.5, 2.9, 0.0, 1.6, 1.8, 3.9, 3.9],
               [1.5, 2.5, 2.9, 1.6, 0.0, 1.6, 3.9, 3.9],
               [1.7, 2.7, 3.1, 1.8, 1.6, 0.0, 4.1, 4.1],
               [3.8, 2.8, 2.0, 3.9, 3.9, 4.1, 0.0, 1.8],
               [3.8, 2.8, 2.0, 3.9, 3.9, 4.1, 1.8, 0.0]],
         "src": "phylo2vec/src/matrix/graph.rs:168-188"},
        {"m": [[0.0, 0.05586577, 0.73334581], [0.0, 0.38201654, 0.61142737],
               [4.0, 0.85919964, 0.30591026], [0.0, 0.25750902, 0.02408398]],
         "unrooted": False, "round": 3,
         "d": [[0.0, 1.603, 1.049, 1.579, 0.789], [1.603, 0.0, 1.777, 0.588, 2.28],
               [1.049, 1.777, 0.0, 1.752, 1.727], [1.579, 0.588, 1.752, 0.0, 2.256],
               [0.789, 2.28, 1.727, 2.256, 0.0]],
         "src": "docs/demo.ipynb cell 41 (pairwise_distances(m5).round(3))"},
    ],
    # ------------------------------------------------- vcv and node depths
    "vcv_vector": [
        {"v": [0], "c": [[1.0, 0.0], [0.0, 1.0]], "src": "phylo2vec/src/vector/graph.rs:607"},
        {"v": [0, 0, 1],
         "c": [[2.0, 0.0, 1.0, 0.0], [0.0, 2.0, 0.0, 1.0], [1.0, 0.0, 2.0, 0.0], [0.0, 1.0, 0.0, 2.0]],
         "src": "phylo2vec/src/vector/graph.rs:608"},
        {"v": [0, 1, 2],
         "c": [[1.0, 0.0, 0.0, 0.0], [0.0, 2.0, 1.0, 1.0], [0.0, 1.0, 3.0, 2.0], [0.0, 1.0, 2.0, 3.0]],
         "src": "phylo2vec/src/vector/graph.rs:609-614 (also doctest :270-278)"},
        {"v": [0, 1, 1, 4],
         "c": [[1.0, 0.0, 0.0, 0.0, 0.0], [0.0, 4.0, 1.0, 3.0, 2.0], [0.0, 1.0, 2.0, 1.0, 1.0],
               [0.0, 3.0, 1.0, 4.0, 2.0], [0.0, 2.0, 1.0, 2.0, 3.0]],
         "src": "phylo2vec/src/vector/graph.rs:615-621"},
        {"v": [0, 1, 0, 4, 4, 2, 6],
         "c": [[3.0, 0.0, 0.0, 2.0, 1.0, 1.0, 0.0, 0.0],
               [0.0, 2.0, 1.0, 0.0, 0.0, 0.0, 1.0, 1.0],
               [0.0, 1.0, 3.0, 0.0, 0.0, 0.0, 2.0, 2.0],
               [2.0, 0.0, 0.0, 3.0, 1.0, 1.0, 0.0, 0.0],
               [1.0, 0.0, 0.0, 1.0, 3.0, 2.0, 0.0, 0.0],
               [1.0, 0.0, 0.0, 1.0, 2.0, 3.0, 0.0, 0.0],
               [0.0, 1.0, 2.0, 0.0, 0.0, 0.0, 4.0, 3.0],
               [0.0, 1.0, 2.0, 0.0, 0.0, 0.0, 3.0, 4.0]],
         "src": "phylo2vec/src/vector/graph.rs:622-630"},
    ],
    "vcv_matrix": [
        {"m": [[0.0, 1.0, 10.0]], "c": [[1.0, 0.0], [0.0, 10.0]], "src": "phylo2vec/src/matrix/graph.rs:200"},
        {"m": [[0.0, 0.4, 0.5], [2.0, 0.1, 0.2], [2.0, 0.3, 0.6]],
         "c": [[0.4, 0.3, 0.0, 0.0], [0.3, 0.5, 0.0, 0.0], [0.0, 0.0, 1.0, 0.6], [0.0, 0.0, 0.6, 1.1]],
         "src": "phylo2vec/src/matrix/graph.rs:201-210 (also doctest :48-59)"},
        {"m": [[0.0, 0.5, 0.5], [2.0, 1.5, 1.5], [2.0, 1.0, 2.0]],
         "c": [[2.5, 1.0, 0.0, 0.0], [1.0, 2.5, 0.0, 0.0], [0.0, 0.0, 2.5, 2.0], [0.0, 0.0, 2.0, 2.5]],
         "src": "phylo2vec/src/matrix/graph.rs:211-221"},
    ],
    "node_depths": [
        {"tree": [0], "depths": {"0": 1.0, "1": 1.0}, "src": "phylo2vec/src/vector/ops.rs:585-586"},
        {"tree": [0, 0, 0], "depths": {"6": 0.0, "5": 1.0, "4": 2.0, "3": 3.0, "2": 2.0, "0": 3.0, "1": 1.0},
         "src": "phylo2vec/src/vector/ops.rs:588-594"},
        {"tree": [0, 0, 1], "depths": {"6": 0.0, "4": 1.0, "5": 1.0, "0": 2.0, "1": 2.0, "2": 2.0, "3": 2.0},
         "src": "phylo2vec/src/vector/ops.rs:596-602"},
        {"tree": [0, 1, 2], "depths": {"6": 0.0, "0": 1.0},
         "src": "py-phylo2vec/phylo2vec/utils/vector.py:322-329 (doctest)"},
        {"tree": [[0.0, 0.5, 0.8], [1.0, 0.7, 0.6]],
         "depths": {"4": 0.0, "3": 0.6, "0": 0.7, "1": 1.1, "2": 1.4},
         "src": "phylo2vec/src/matrix/ops.rs:91-115"},
        {"tree": [[0.0, 0.9, 0.4], [0.0, 0.8, 3.0], [3.0, 0.4, 0.5]],
         "depths": {"6": 0.0, "5": 0.4, "4": 1.2, "1": 0.5, "3": 3.4, "0": 2.1, "2": 1.6},
         "src": "phylo2vec/src/matrix/ops.rs:117-157"},
    ],
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_goldens.json")
    with open(out, "w") as f:
        json.dump(G, f, indent=1)
    print("wrote", out)
