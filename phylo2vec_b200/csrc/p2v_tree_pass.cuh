// Warp-wide passes over the rows of an ancestry-like table in row order (children before parents), shared by
// the Robinson-Foulds kernel and the matrix Newick kernels.  ch[2r], ch[2r+1] = the two children of row r as
// node labels (leaf <= k, internal = k + 1 + row); par[r] = parent row.
//
// 32 rows at a time; a row whose child / parent row lies in the same batch waits for it, and the batch is
// relaxed until every lane has fired (a few rounds on random trees, 32 on a ladder).
#pragma once
#include "p2v_common.cuh"

namespace p2v {

// Bottom-up pass over rows [0,k): fire(r) runs once both children of row r are final.
// Top-down pass: fire(r) runs once the parent row of r is final (the root row k-1 fires first).
template <typename ST, typename Fire>
__device__ __forceinline__ void rf_bottom_up(const ST* ch, int k, int lane, Fire fire) {
    for (int base = 0; base < k; base += 32) {
        const int r = base + lane;
        const bool act = r < k;
        int c1 = 0, c2 = 0;
        if (act) { c1 = (int)ch[2 * r]; c2 = (int)ch[2 * r + 1]; }
        // a child is a leaf (label <= k), a row of an earlier batch, or a row of this batch
        const int w1 = (act && c1 > k && c1 - k - 1 >= base) ? c1 - k - 1 - base : -1;
        const int w2 = (act && c2 > k && c2 - k - 1 >= base) ? c2 - k - 1 - base : -1;
        unsigned done_mask = ~__ballot_sync(FULL, act);          // inactive lanes count as done
        bool fired = !act;
        while (done_mask != FULL) {
            const bool ready = !fired && (w1 < 0 || ((done_mask >> w1) & 1u)) && (w2 < 0 || ((done_mask >> w2) & 1u));
            if (ready) { fire(r, c1, c2); fired = true; }
            __syncwarp();
            done_mask |= __ballot_sync(FULL, ready);
        }
    }
}
template <typename ST, typename Fire>
__device__ __forceinline__ void rf_top_down(const ST* par, int k, int lane, Fire fire) {
    for (int base = ((k - 1) / 32) * 32; base >= 0; base -= 32) {
        const int r = base + lane;
        const bool act = r < k;
        const int p = (act && r != k - 1) ? (int)par[r] : -1;        // parent row; the root has none
        const int w = (p >= 0 && p < base + 32) ? p - base : -1;      // parent in this batch
        unsigned done_mask = ~__ballot_sync(FULL, act);
        bool fired = !act;
        while (done_mask != FULL) {
            const bool ready = !fired && (w < 0 || ((done_mask >> w) & 1u));
            if (ready) { fire(r, p); fired = true; }
            __syncwarp();
            done_mask |= __ballot_sync(FULL, ready);
        }
    }
}


}  // namespace p2v
