// gram_plan.h -- host-side work plan of the K2 Gram kernel (plain C++, no CUDA: also compiled into the CPU
// test emulation, tests/host_emu/fokl_emu.cpp).
//
// The block K2 has to form is   [X_old  X_new  y]' X_new     ((P_old + C + 1) x C).
// Plan vocabulary:
//   A-list   the (P_old + C + 1) operand columns in the order  new_0 .. new_{C-1} | pad to 8 | old_0 .. old_{P_old-1}, y |
//            pad to 8.  Putting the new columns first makes the symmetric X_new' X_new part start on an 8-column
//            boundary, so "this 8 x 8 fragment lies entirely below the diagonal" is the plain test  i > j.
//   fragment one 8 x 8 piece of the output (one mma.sync.m8n8k4.f64 accumulator): rows = 8 consecutive A-list entries,
//            columns = 8 consecutive new columns.
//   block    2 x 2 fragments (16 x 16 outputs): what one warp keeps in registers per block slot.
//   tile     up to warps * 4 blocks + the staged-column ("slot") list they touch: the work of one CTA per row split.
// Needed fragments: every (old / y row, new col) fragment and the fragments of the new x new part with i <= j
// (fokl_gram_scatter mirrors the upper triangle); a block carries a 4-bit mask of its needed fragments, the others are
// neither computed nor read back.  The needed blocks are walked in 8 x 8-block rectangles (so a tile's
// blocks share operand columns) and cut into tiles of equal size.
#pragma once
#include <stdint.h>
#include <algorithm>
#include <vector>

namespace fokl {

constexpr int kGramBlocksPerWarp = 4;   // 2 x 2-fragment blocks per warp (4 * 4 * 2 = 32 accumulator doubles per lane)
constexpr int kGramMaxWarps = 16;       // warps per CTA: 16 (one CTA per SM) or 8 (two CTAs per SM)
inline int gram_tile_blocks(int warps) { return warps * kGramBlocksPerWarp; }
constexpr int kGramRect = 8;            // rectangles of 8 x 8 blocks = 128 x 128 outputs

struct GramTileMeta {
    int32_t slot_off, n_slots;          // into slot_* arrays; n_slots is a multiple of 8
    int32_t blk_off, n_blk;             // into blocks
};

struct GramBlockMeta {
    uint16_t a_slot, b_slot;            // first of 16 consecutive tile-local slots on the row / column side
    uint32_t mask;                      // bit f = fragment (f >> 1, f & 1) of the block is needed
};

struct GramPlan {
    std::vector<GramTileMeta> tiles;
    std::vector<int32_t> slot_src;      // per slot: column of X (0 .. p-1), p = y, -1 = zero padding
    std::vector<int32_t> slot_arow;     // per slot: row of the output block, -1 = none
    std::vector<int32_t> slot_bcol;     // per slot: column of the output block (new-column index), -1 = none
    std::vector<GramBlockMeta> blocks;
    int max_slots = 0;
};

// max_slots_cap: most staged columns a tile may have (shared-memory budget), multiple of 16 and >= 32.
inline GramPlan gram_make_plan(int p_old, int c, int max_slots_cap, int warps = kGramMaxWarps)
{
    const int kTileBlocks = gram_tile_blocks(warps);
    GramPlan pl;
    const int p = p_old + c;
    const int fb = (c + 7) / 8;                      // fragment rows/cols of the new columns
    const int fo = (p_old + 1 + 7) / 8;              // fragment rows of old columns + y
    const int fa = fb + fo;                          // A-list length in fragments
    auto alist_src = [&](int e) -> int {             // A-list entry -> X column (p = y, -1 = pad)
        if (e < fb * 8) return e < c ? p_old + e : -1;
        int o = e - fb * 8;
        return o < p_old ? o : (o == p_old ? p : -1);
    };
    auto frag_needed = [&](int i, int j) { return i >= fb || i <= j; };
    const int ba = (fa + 1) / 2, bb = (fb + 1) / 2;  // block rows / cols (a block = fragment pair)
    auto block_needed = [&](int ib, int jb) {
        for (int di = 0; di < 2; ++di)
            for (int dj = 0; dj < 2; ++dj) {
                int i = 2 * ib + di, j = 2 * jb + dj;
                if (i < fa && j < fb && frag_needed(i, j)) return true;
            }
        return false;
    };
    // needed blocks in rectangle order
    std::vector<std::pair<int, int>> order;
    for (int rj = 0; rj < bb; rj += kGramRect)
        for (int ri = 0; ri < ba; ri += kGramRect)
            for (int ib = ri; ib < std::min(ri + kGramRect, ba); ++ib)
                for (int jb = rj; jb < std::min(rj + kGramRect, bb); ++jb)
                    if (block_needed(ib, jb)) order.push_back({ib, jb});
    const int total = (int)order.size();
    const int n_tiles_min = (total + kTileBlocks - 1) / kTileBlocks;
    const int target = n_tiles_min ? (total + n_tiles_min - 1) / n_tiles_min : 0;

    std::vector<int> frag_local(fa + 1, -1);         // A-list fragment row -> tile-local fragment index
    size_t pos = 0;
    while (pos < order.size()) {
        // grow a tile: up to `target` blocks while the slot union stays under the cap
        std::vector<int> used;                       // A-list fragment rows touched (as row or column operand)
        std::fill(frag_local.begin(), frag_local.end(), -1);
        auto touch_count = [&](int ib, int jb) {
            int add = 0;
            int fr[4] = {2 * ib, 2 * ib + 1, 2 * jb, 2 * jb + 1};
            for (int q = 0; q < 4; ++q) {
                bool dup = false;
                for (int r = 0; r < q; ++r) dup |= fr[r] == fr[q];
                if (!dup && frag_local[fr[q]] < 0) ++add;
            }
            return add;
        };
        size_t first = pos;
        int n_blk = 0;
        while (pos < order.size() && n_blk < target) {
            int ib = order[pos].first, jb = order[pos].second;
            int add = touch_count(ib, jb);
            if (n_blk > 0 && ((int)used.size() + add) * 8 > max_slots_cap) break;
            int fr[4] = {2 * ib, 2 * ib + 1, 2 * jb, 2 * jb + 1};
            for (int q = 0; q < 4; ++q)
                if (frag_local[fr[q]] < 0) { frag_local[fr[q]] = 1; used.push_back(fr[q]); }
            ++pos;
            ++n_blk;
        }
        std::sort(used.begin(), used.end());
        // both fragment rows of a block are always touched together, so they stay adjacent after sorting
        for (size_t u = 0; u < used.size(); ++u) frag_local[used[u]] = (int)u;
        GramTileMeta tm;
        tm.slot_off = (int32_t)pl.slot_src.size();
        tm.n_slots = (int32_t)used.size() * 8;
        tm.blk_off = (int32_t)pl.blocks.size();
        tm.n_blk = n_blk;
        for (int f : used)
            for (int r = 0; r < 8; ++r) {
                int e = f * 8 + r;
                int src = f < fa ? alist_src(e) : -1;
                pl.slot_src.push_back(src);
                pl.slot_arow.push_back(src < 0 ? -1 : src);             // output row index = X column index (y -> p)
                pl.slot_bcol.push_back((src >= p_old && src < p) ? src - p_old : -1);
            }
        std::vector<GramBlockMeta> full_blk, part_blk;
        for (size_t q = first; q < pos; ++q) {
            GramBlockMeta bm;
            bm.a_slot = (uint16_t)(frag_local[2 * order[q].first] * 8);
            bm.b_slot = (uint16_t)(frag_local[2 * order[q].second] * 8);
            bm.mask = 0;
            for (int f = 0; f < 4; ++f) {
                int i = 2 * order[q].first + (f >> 1), j = 2 * order[q].second + (f & 1);
                if (i < fa && j < fb && frag_needed(i, j)) bm.mask |= 1u << f;
            }
            (bm.mask == 15u ? full_blk : part_blk).push_back(bm);
        }
        // Warp w owns the blocks at positions w, w + warps, ... < n_blk.  Blocks with skipped fragments need the predicated
        // DMMA form, which costs issue slots: hand them to as few warps as possible (warp 0's positions first, then
        // warp 1's, ...), the fully needed blocks to the rest.
        {
            std::vector<GramBlockMeta> seq(part_blk);
            seq.insert(seq.end(), full_blk.begin(), full_blk.end());
            std::vector<GramBlockMeta> placed(n_blk);
            size_t next = 0;
            for (int w = 0; w < warps; ++w)
                for (int q = w; q < n_blk; q += warps) placed[q] = seq[next++];
            pl.blocks.insert(pl.blocks.end(), placed.begin(), placed.end());
        }
        pl.tiles.push_back(tm);
        pl.max_slots = std::max(pl.max_slots, (int)tm.n_slots);
    }
    return pl;
}

}  // namespace fokl
