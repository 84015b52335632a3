// gram_plan.h -- host-side work plan of the K2 Gram kernel (plain C++, no CUDA: also compiled into the CPU
// test emulation, tests/host_emu/fokl_emu.cpp).
//
// The block K2 has to form is   [X_old  X_new  y]' X_new     ((P_old + C + 1) x C).
// Plan vocabulary:
//   A-list   the (P_old + C + 1) operand columns in the order  new_0 .. new_{C-1} | pad to 8 | y, old_0 .. old_{P_old-1} |
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
// (Measured and dropped, profiles/r02_gram_round2.txt: a fifth block for the three consumer warps that share their SM
// sub-partition with the producer warp, dealt so that the four sub-partitions carry equal block counts -- the 40
// accumulators leave no registers for loads in flight under the 128-register cap of a 16-warp CTA, those warps become
// the slowest of every stage and the ring makes everybody wait for them: C = 168 launches 18.5 -> 19.8 ms.)
inline int gram_warp_cap(int, int) { return kGramBlocksPerWarp; }
inline int gram_tile_cap(int warps) { return warps * kGramBlocksPerWarp; }       // most blocks (work items) of a tile
inline int gram_tile_blocks(int warps) { return warps * kGramBlocksPerWarp; }    // positions per tile (workspace stride)
constexpr int kGramRect = 8;            // rectangles of 8 x 8 blocks = 128 x 128 outputs

struct GramTileMeta {
    int32_t slot_off, n_slots;          // into slot_* arrays; n_slots is a multiple of 8
    int32_t blk_off, n_blk;             // into blocks; n_blk = positions in use (work items and holes)
    int32_t ksplit, pad;                // every block of the tile is shared by `ksplit` work items (power of two)
    int32_t box_off, n_box;             // into boxes (TMA path)
};

// TMA path: the engine keeps y in the row in front of X ("[y | X]": column 0 = y, column 1 + j = X column j), so every
// 8-slot group of a tile is 8 consecutive columns of that buffer and consecutive groups merge into runs; a run is
// fetched as boxes of kGramBoxCols[kind] columns x 16 rows, one 2-D tensor-map copy each.  Columns past the end of the
// buffer's used part (padding slots of the last new group) are zero-filled by the copy engine.
constexpr int kGramBoxKinds = 3;
constexpr int kGramBoxCols[kGramBoxKinds] = {8, 32, 128};
struct GramBoxMeta {
    int32_t slot0;                      // first tile-local slot
    int32_t xcol0;                      // first column of [y | X]
    int32_t kind, ncols;                // box kind and its column count (kGramBoxCols[kind])
};

// One work item = one 2 x 2-fragment block restricted to the 16-row chunks  phase, phase + ksplit, ...  of every slab.
struct GramBlockMeta {
    uint16_t a_slot, b_slot;            // regular block: first of 16 consecutive tile-local slots on the row / column side
    uint8_t mask;                       // bit f = fragment f of the block is needed
    uint8_t phase;                      // 0 .. ksplit - 1; the phase-0 item is the head of its block's item chain
    int16_t next;                       // tile-local position of the block's next item, -1 = last
    // Fragment f of the block = rows: slots la[f] .. la[f] + 7, columns: slots lb[f] .. lb[f] + 7.  A regular block is
    // the 2 x 2 arrangement la[f] = a_slot + 8 (f >> 1), lb[f] = b_slot + 8 (f & 1) (four operand loads per four DMMAs).
    // A LOOSE block (loose = 1) packs four unrelated fragments -- the needed fragments of what would otherwise be
    // half-empty diagonal / edge blocks, whose skipped fragments would still cost their DMMA issue slots (eight operand
    // loads per four DMMAs, but no wasted slot).
    uint16_t la[4], lb[4];
    uint16_t loose, pad;
};

struct GramPlan {
    std::vector<GramTileMeta> tiles;
    std::vector<int32_t> slot_src;      // per slot: column of X (0 .. p-1), p = y, -1 = zero padding
    std::vector<int32_t> slot_arow;     // per slot: row of the output block, -1 = none
    std::vector<int32_t> slot_bcol;     // per slot: column of the output block (new-column index), -1 = none
    std::vector<GramBlockMeta> blocks;
    std::vector<GramBoxMeta> boxes;
    int max_slots = 0;
};

// k-split of a tile with few blocks.  A warp keeps at most 4 blocks' accumulators, and a block's DMMAs all issue from
// the one warp that owns it, so a tile with fewer blocks than the CTA has positions (warps * 4) leaves tensor-pipe issue
// slots idle (8 new columns against 50 old ones = 5 blocks: 5 of 16 warps busy, one block each).  Such a tile hands
// every block to `ksplit` work items that take the 16-row chunks  phase, phase + ksplit, ...  of each slab; the items of
// a block are summed (in phase order, then in row-split order) by gram_reduce_kernel.
//
// Placement.  Warp w owns the positions w, w + warps, ... (at most 4, filled from the first).  Per k-step an item costs
// its warp four DMMA issue slots of ~25 cycles (a lone warp cannot issue them faster, and a predicated-off DMMA keeps its
// slot: ncu source view, profiles/r01_gram_ksplit.txt) and the tensor pipe of the warp's SM sub-partition (w mod 4)
// 16 cycles per needed fragment; the split with the smallest per-slab makespan
//     max(100 * max_warp(items), 16 * max_subpartition(fragments)) / ksplit
// wins, ties to the smaller split.  Two deals:
//   mode 1 (used by K2)  sequential: the items in list order -- blocks with skipped fragments first, so that as few
//          warps as possible run the predicated DMMA form, then the full blocks in rectangle order, the items of a
//          block adjacent -- fill warp 0's positions, then warp 1's, ...
//   mode 0 balanced: items with skipped fragments are packed into as few warps as an even deal allows, full ones go
//          to the other warps, and within a kind an item goes to the sub-partition with the least fragments so far.
//          Measured 4 - 6 % slower than mode 1 on the wide three-way tiles and equal within 2 % elsewhere.
// Unused positions below n_blk are holes (mask 0).
struct GramPlacement {
    std::vector<int> pos;               // item (block g, phase f) = index g * ksplit + f -> tile-local position
    int n_pos = 0;
    double cost = 0.0;
};

inline int gram_popcount4(unsigned m) { return (int)((m & 1u) + (m >> 1 & 1u) + (m >> 2 & 1u) + (m >> 3 & 1u)); }

inline GramPlacement gram_place_items(const std::vector<int> &weight, int r, int warps, int mode = 0,
                                      const std::vector<int> *flex = nullptr)
{
    const int nb = (int)weight.size(), n_items = nb * r;
    const int per = (n_items + warps - 1) / warps;                 // items per warp when dealt evenly
    GramPlacement pc;
    pc.pos.assign(n_items, -1);
    if (mode == 2) {
        // even deal over the SM sub-partitions (warp w issues on sub-partition w & 3, whose tensor pipe is the shared
        // resource): an item goes to the sub-partition with the fewest items, inside it to the warp with the fewest.
        // Loose / masked items (flex) take the FIRST position of a warp, one per warp, so that every warp runs the
        // one-loose-block form at most; when the tile has more of them than busy warps the items are dealt in list
        // order (loose first) and the first warps take the all-loose form.
        std::vector<int> cnt(warps, 0);
        int sl[4] = {0, 0, 0, 0};
        for (int i = 0; i < n_items; ++i) {
            int best = -1;
            for (int w = 0; w < warps; ++w) {
                if (cnt[w] >= gram_warp_cap(warps, w)) continue;
                if (best < 0) { best = w; continue; }
                const int a = sl[w & 3], b = sl[best & 3];
                if (a != b ? a < b : cnt[w] < cnt[best]) best = w;
            }
            ++cnt[best];
            ++sl[best & 3];
        }
        int busy = 0, max_cnt = 0;
        for (int w = 0; w < warps; ++w) { busy += cnt[w] > 0; max_cnt = std::max(max_cnt, cnt[w]); }
        std::vector<int> flex_items, full_items;
        for (int it = 0; it < n_items; ++it) ((flex && (*flex)[it / r]) ? flex_items : full_items).push_back(it);
        std::vector<int> used(warps, 0);
        size_t nf = 0;
        double all_loose_penalty = 1.0;
        if ((int)flex_items.size() <= busy) {
            for (int w = 0; w < warps && nf < flex_items.size(); ++w)
                if (cnt[w] > 0) { pc.pos[flex_items[nf++]] = w; used[w] = 1; }
        } else {
            full_items.insert(full_items.begin(), flex_items.begin(), flex_items.end());
            // warps whose blocks all take the loose form are the slowest of every slab (twice the operand loads):
            // 14 old + 56 new columns, 1.83 ms unsplit against 2.37 ms split in four (profiles/r02_gram_round2.txt)
            all_loose_penalty = 1.3;
        }
        size_t nx = 0;
        for (int w = 0; w < warps; ++w)
            for (int b = used[w]; b < cnt[w]; ++b) {
                pc.pos[full_items[nx++]] = w + warps * b;
                pc.n_pos = std::max(pc.n_pos, w + warps * b + 1);
            }
        for (int w = 0; w < warps; ++w)
            if (used[w]) pc.n_pos = std::max(pc.n_pos, w + 1);
        pc.cost = all_loose_penalty * std::max(100.0 * max_cnt, 64.0 * std::max(std::max(sl[0], sl[1]), std::max(sl[2], sl[3]))) / r;
        return pc;
    }
    if (mode == 1) {
        // sequential deal: the items in list order (blocks with skipped fragments come first) fill warp 0's positions,
        // then warp 1's, ...  Cost: a skipped fragment keeps its issue slot, so a sub-partition's load is counted in
        // items (4 slots each), not in needed fragments.
        int next = 0, max_cnt = 0;
        int sl[4] = {0, 0, 0, 0};
        for (int w = 0; w < warps; ++w) {
            int cnt = 0;
            for (int q = w; q < n_items; q += warps) { pc.pos[next] = q; ++sl[w & 3]; ++next; ++cnt; }
            max_cnt = std::max(max_cnt, cnt);
        }
        pc.n_pos = n_items;
        pc.cost = std::max(100.0 * max_cnt, 64.0 * std::max(std::max(sl[0], sl[1]), std::max(sl[2], sl[3]))) / r;
        return pc;
    }
    std::vector<int> order(n_items);
    for (int i = 0; i < n_items; ++i) order[i] = i;
    // items with skipped fragments first (lightest last among them), then the full ones
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
        const int wx = weight[x / r], wy = weight[y / r];
        if ((wx == 4) != (wy == 4)) return wx != 4;
        return wx > wy;
    });
    std::vector<int> wload(warps, 0), wcnt(warps, 0), wmasked(warps, 0);
    int sload[4] = {0, 0, 0, 0};
    for (int it : order) {
        const bool part = weight[it / r] != 4;
        // preference classes (lower = better).  An item with skipped fragments joins a warp that already runs the
        // predicated form while that warp is under the even deal, else opens an untouched warp; a full item goes to
        // an untouched or unpredicated warp under the even deal (spread, not packed), then to any unpredicated one.
        int best = -1, best_cls = 9;
        for (int w = 0; w < warps; ++w) {
            if (wcnt[w] >= gram_warp_cap(warps, w)) continue;
            int cls;
            if (part) cls = (wmasked[w] && wcnt[w] < per) ? 0 : (wcnt[w] == 0 ? 1 : (wmasked[w] ? 2 : 3));
            else cls = (!wmasked[w] && wcnt[w] < per) ? 0 : (!wmasked[w] ? 1 : 2);
            bool better;
            if (best < 0 || cls < best_cls) better = true;
            else if (cls > best_cls) better = false;
            else if (sload[w & 3] != sload[best & 3]) better = sload[w & 3] < sload[best & 3];
            else if (wcnt[w] != wcnt[best]) better = wcnt[w] < wcnt[best];
            else better = wload[w] < wload[best];
            if (better) { best = w; best_cls = cls; }
        }
        pc.pos[it] = best + warps * wcnt[best];
        pc.n_pos = std::max(pc.n_pos, pc.pos[it] + 1);
        ++wcnt[best];
        wmasked[best] |= part ? 1 : 0;
        wload[best] += weight[it / r];
        sload[best & 3] += weight[it / r];
    }
    const int max_cnt = *std::max_element(wcnt.begin(), wcnt.end());
    const int max_frag = std::max(std::max(sload[0], sload[1]), std::max(sload[2], sload[3]));
    pc.cost = std::max(100.0 * max_cnt, 16.0 * max_frag) / r;
    return pc;
}

// Place the blocks of every tile of a fresh plan for slabs of `kchunks` 16-row chunks (KB / 16 = the largest
// admissible ksplit).  Call once per plan.
inline void gram_plan_place(GramPlan &pl, int warps, int kchunks, int mode = 0, int force_r = 0)
{
    const int cap = gram_tile_cap(warps);
    std::vector<GramBlockMeta> out;
    for (GramTileMeta &tm : pl.tiles) {
        std::vector<GramBlockMeta> base;
        std::vector<int> weight, flex;
        for (int q = 0; q < tm.n_blk; ++q) {
            const GramBlockMeta &bm = pl.blocks[tm.blk_off + q];
            if (bm.phase == 0 && bm.mask != 0) {
                base.push_back(bm);
                weight.push_back(gram_popcount4(bm.mask));
                flex.push_back(bm.loose != 0 || bm.mask != 15u);
            }
        }
        const int nb = (int)base.size();
        GramPlacement best;
        int best_r = 0;
        for (int r = 1; r <= kchunks && nb * r <= cap; r *= 2) {
            GramPlacement pc = gram_place_items(weight, r, warps, mode, &flex);
            if (best_r == 0 || (force_r ? r <= force_r : pc.cost < best.cost * 0.999)) { best = pc; best_r = r; }
        }
        GramBlockMeta hole;
        hole.a_slot = hole.b_slot = 0; hole.mask = 0; hole.phase = 0; hole.next = -1;
        for (int f = 0; f < 4; ++f) hole.la[f] = hole.lb[f] = 0;
        hole.loose = 0; hole.pad = 0;
        std::vector<GramBlockMeta> placed(best.n_pos, hole);
        for (int g = 0; g < nb; ++g)
            for (int f = 0; f < best_r; ++f) {
                GramBlockMeta bm = base[g];
                bm.phase = (uint8_t)f;
                bm.next = (int16_t)(f + 1 < best_r ? best.pos[g * best_r + f + 1] : -1);
                placed[best.pos[g * best_r + f]] = bm;
            }
        tm.blk_off = (int32_t)out.size();
        tm.n_blk = best.n_pos;
        tm.ksplit = best_r;
        out.insert(out.end(), placed.begin(), placed.end());
    }
    pl.blocks.swap(out);
}

// max_slots_cap: most staged columns a tile may have (shared-memory budget), multiple of 16 and >= 32.
// gap: the new columns are not stored right behind the old ones but `gap` columns further up (X columns
// p_old + gap .. p_old + gap + c - 1: a block that was built ahead of time, while the columns in between were still
// candidates of the previous substage); slot_src and the boxes then hold these PHYSICAL column indices and y is marked
// by p_old + c + gap, while slot_arow / slot_bcol keep the logical output indices.  cross_only: only the
// (old | y) x new fragments (the new x new part was formed by an earlier call).
inline GramPlan gram_make_plan(int p_old, int c, int max_slots_cap, int warps = kGramMaxWarps, int gap = 0,
                               bool cross_only = false, int pad_mode = -1, bool allow_loose = true)
{
    if (pad_mode < 0) {
        // pad_mode: an all-padding fragment row between an odd number of new fragment rows and the old ones (see
        // below) -- whichever of the two layouts needs fewer blocks
        if (((c + 7) / 8) % 2 == 0) return gram_make_plan(p_old, c, max_slots_cap, warps, gap, cross_only, 0, allow_loose);
        GramPlan a = gram_make_plan(p_old, c, max_slots_cap, warps, gap, cross_only, 0, allow_loose);
        GramPlan b = gram_make_plan(p_old, c, max_slots_cap, warps, gap, cross_only, 1, allow_loose);
        auto n_loose = [](const GramPlan &pl) {
            size_t k = 0;
            for (const GramBlockMeta &bm : pl.blocks) k += bm.loose;
            return k;
        };
        if (b.blocks.size() != a.blocks.size()) return b.blocks.size() < a.blocks.size() ? b : a;
        return n_loose(b) < n_loose(a) ? b : a;
    }
    const int kTileBlocks = gram_tile_cap(warps);
    GramPlan pl;
    const int p = p_old + c;
    auto phys = [&](int src) { return src < 0 ? -1 : (src >= p_old ? src + gap : src); };    // y (= p) -> p + gap
    const int fb = (c + 7) / 8;                      // fragment rows/cols of the new columns
    const int fo = (p_old + 1 + 7) / 8;              // fragment rows of old columns + y
    // pad_mode 1: the old rows start on an even fragment row (one all-padding fragment row after an odd number of new
    // ones), so that no 2 x 2 block straddles the new and the old part.  With an even number of old fragment rows the
    // straddling block row costs a whole extra row of half-empty blocks (C = 168 over 59 old rows: 120 blocks instead
    // of 110); with an odd number it pairs the last new with the first old fragment row and is the better layout.
    const int fbp = fb + (pad_mode ? (fb & 1) : 0);
    const int fa = fbp + fo;                         // A-list length in fragments
    auto alist_src = [&](int e) -> int {             // A-list entry -> X column (p = y, -1 = pad)
        if (e < fbp * 8) return e < c ? p_old + e : -1;
        int o = e - fbp * 8;                         // y first: [y | X_old] is one run of the engine's [y | X] buffer
        return o == 0 ? p : (o <= p_old ? o - 1 : -1);
    };
    auto frag_needed = [&](int i, int j) { return i >= fbp || (i < fb && !cross_only && i <= j); };
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
    // work items once the half-empty blocks are repacked into loose blocks (estimate for the whole block: every tile
    // rounds its own loose fragments up to a multiple of four, so the tile count is raised until every tile fits)
    int full_total = 0, part_frags = 0;
    std::vector<int> order_q(order.size(), 0);      // needed fragments of every block of the walk
    for (size_t o = 0; o < order.size(); ++o) {
        const auto &ob = order[o];
        int cntf = 0;
        for (int f = 0; f < 4; ++f) {
            int i = 2 * ob.first + (f >> 1), j = 2 * ob.second + (f & 1);
            if (i < fa && j < fb && frag_needed(i, j)) ++cntf;
        }
        order_q[o] = allow_loose ? cntf : 4;         // (predicated form: a half-empty block still costs a whole item)
        if (cntf == 4) ++full_total;
        else part_frags += cntf;
    }
    long total_q = 0;
    for (int q : order_q) total_q += q;
    const int items_est = allow_loose ? full_total + (part_frags + 3) / 4 : total;
    std::vector<int> frag_local(fa + 1, -1);         // A-list fragment row -> tile-local fragment index
  for (int n_tiles_try = std::max(1, (items_est + kTileBlocks - 1) / kTileBlocks);; ++n_tiles_try) {
    pl = GramPlan();
    // tiles of equal WORK: a tile ends where the needed fragments walked so far are nearest to its share (the blocks with
    // skipped fragments are repacked four fragments to an item, so items ~ needed fragments / 4)
    const double target_q = (double)total_q / n_tiles_try;
    long cum_q = 0;
    int tile_idx = 0;
    bool fits = true;
    size_t pos = 0;
    while (pos < order.size()) {
        // grow a tile up to its share of the work while the slot union stays under the cap
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
        while (pos < order.size()) {
            if (n_blk > 0 && tile_idx + 1 < n_tiles_try && cum_q + 0.5 * order_q[pos] > (tile_idx + 1) * target_q) break;
            int ib = order[pos].first, jb = order[pos].second;
            int add = touch_count(ib, jb);
            if (n_blk > 0 && ((int)used.size() + add) * 8 > max_slots_cap) break;
            int fr[4] = {2 * ib, 2 * ib + 1, 2 * jb, 2 * jb + 1};
            for (int q = 0; q < 4; ++q)
                if (frag_local[fr[q]] < 0) { frag_local[fr[q]] = 1; used.push_back(fr[q]); }
            cum_q += order_q[pos];
            ++pos;
            ++n_blk;
        }
        ++tile_idx;
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
                pl.slot_src.push_back(phys(src));
                pl.slot_arow.push_back(src < 0 ? -1 : src);             // output row index = logical column index (y -> p)
                pl.slot_bcol.push_back((src >= p_old && src < p) ? src - p_old : -1);
            }
        std::vector<GramBlockMeta> full_blk, part_blk;
        for (size_t q = first; q < pos; ++q) {
            GramBlockMeta bm;
            bm.a_slot = (uint16_t)(frag_local[2 * order[q].first] * 8);
            bm.b_slot = (uint16_t)(frag_local[2 * order[q].second] * 8);
            bm.mask = 0;
            bm.phase = 0;
            bm.next = -1;
            for (int f = 0; f < 4; ++f) {
                int i = 2 * order[q].first + (f >> 1), j = 2 * order[q].second + (f & 1);
                if (i < fa && j < fb && frag_needed(i, j)) bm.mask |= (uint8_t)(1u << f);
            }
            bm.loose = 0;
            bm.pad = 0;
            (bm.mask == 15u ? full_blk : part_blk).push_back(bm);
        }
        for (GramBlockMeta &bm : full_blk) {
            for (int f = 0; f < 4; ++f) {
                bm.la[f] = (uint16_t)(bm.a_slot + 8 * (f >> 1));
                bm.lb[f] = (uint16_t)(bm.b_slot + 8 * (f & 1));
            }
            bm.loose = 0;
            bm.pad = 0;
        }
        // the needed fragments of the half-empty blocks, repacked four to a loose block
        std::vector<GramBlockMeta> loose_blk;
        if (!allow_loose) {
            // predicated form (the cp.async kernel): the half-empty blocks stay 2 x 2 blocks with skipped fragments
            for (GramBlockMeta &bm : part_blk)
                for (int f = 0; f < 4; ++f) {
                    bm.la[f] = (uint16_t)(bm.a_slot + 8 * (f >> 1));
                    bm.lb[f] = (uint16_t)(bm.b_slot + 8 * (f & 1));
                }
            loose_blk = part_blk;
        } else {
            std::vector<std::pair<uint16_t, uint16_t>> lf;
            for (const GramBlockMeta &bm : part_blk)
                for (int f = 0; f < 4; ++f)
                    if (bm.mask >> f & 1u)
                        lf.push_back({(uint16_t)(bm.a_slot + 8 * (f >> 1)), (uint16_t)(bm.b_slot + 8 * (f & 1))});
            for (size_t g = 0; g < lf.size(); g += 4) {
                GramBlockMeta bm;
                bm.a_slot = lf[g].first;
                bm.b_slot = lf[g].second;
                bm.mask = 0;
                bm.phase = 0;
                bm.next = -1;
                bm.loose = 1;
                bm.pad = 0;
                for (int f = 0; f < 4; ++f) {
                    const bool have = g + f < lf.size();
                    bm.la[f] = have ? lf[g + f].first : lf[g].first;
                    bm.lb[f] = have ? lf[g + f].second : lf[g].second;
                    if (have) bm.mask |= (uint8_t)(1u << f);
                }
                loose_blk.push_back(bm);
            }
        }
        tm.n_blk = (int32_t)(loose_blk.size() + full_blk.size());
        // loose blocks first: as few warps as possible run the eight-load form (see gram_plan_place)
        pl.blocks.insert(pl.blocks.end(), loose_blk.begin(), loose_blk.end());
        pl.blocks.insert(pl.blocks.end(), full_blk.begin(), full_blk.end());
        tm.ksplit = 1;
        tm.pad = 0;
        // boxes: runs of 8-slot groups that are consecutive in [y | X], cut greedily into the largest box kinds
        tm.box_off = (int32_t)pl.boxes.size();
        {
            auto xcol = [&](int src) { return src == p + gap ? 0 : src + 1; };     // physical index -> column of [y | X]
            const int n_grp = tm.n_slots / 8;
            int g = 0;
            while (g < n_grp) {
                if (pl.slot_src[tm.slot_off + 8 * g] < 0) { ++g; continue; }       // all-padding fragment row: never read
                const int x0 = xcol(pl.slot_src[tm.slot_off + 8 * g]);
                int len = 1;
                while (g + len < n_grp && pl.slot_src[tm.slot_off + 8 * (g + len)] >= 0 &&
                       xcol(pl.slot_src[tm.slot_off + 8 * (g + len)]) == x0 + 8 * len)
                    ++len;
                int done = 0;
                while (done < len) {
                    int kind = kGramBoxKinds - 1;
                    while (kind > 0 && kGramBoxCols[kind] > 8 * (len - done)) --kind;
                    GramBoxMeta bx;
                    bx.slot0 = 8 * (g + done);
                    bx.xcol0 = x0 + 8 * done;
                    bx.kind = kind;
                    bx.ncols = kGramBoxCols[kind];
                    pl.boxes.push_back(bx);
                    done += kGramBoxCols[kind] / 8;
                }
                g += len;
            }
        }
        tm.n_box = (int32_t)pl.boxes.size() - tm.box_off;
        pl.tiles.push_back(tm);
        pl.max_slots = std::max(pl.max_slots, (int)tm.n_slots);
        fits = fits && tm.n_blk <= kTileBlocks;
    }
    if (fits || n_tiles_try >= total) break;
  }
    return pl;      // blocks still in list order, one item each: the caller runs gram_plan_place() once
}

}  // namespace fokl
