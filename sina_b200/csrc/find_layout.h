// Launch geometry of the k-mer search (find_tile_kernel / find_merge_kernel, search.cu): plain host arithmetic, kept
// apart from the CUDA sources so that the host unit checks (host/host_unit.cpp, no GPU) can walk every layout.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <algorithm>

namespace sg {

constexpr uint32_t SUB_DEFAULT = 4096;       // references per search sub-tile: one warp owns its u16 score counters (8 KB)
constexpr uint32_t SUB_MAX = 32768;          // a local id never equals 0xffff, the search kernel's "no posting"
constexpr uint32_t TILE_WARPS_MAX = 24;      // sub-tiles (= warps) per search CTA: 24 x 8 KB = 192 KB of counters
constexpr uint32_t FIND_MAX_SORT = 16384;    // candidates the top-k merge sorts in shared memory
constexpr int FIND_KC = 192;                 // k-mers whose offsets are staged at a time (at most)
constexpr int FIND_PRE = 8;                  // staged offsets a thread fetches per chunk (at most)
constexpr uint32_t SEL_BINS = 1024;          // widest score window histogrammed at once
constexpr uint32_t TIE_CAP = 1024;           // ties at the threshold ranked in shared memory (more: id-ordered walk)
constexpr size_t FIND_SMEM_SM = 228 * 1024, FIND_SMEM_CTA = 227 * 1024;   // shared memory of an SM / of one CTA (sm_100)
constexpr size_t FIND_SMEM_FIXED = 1280;     // per CTA: 1 KB reserved by the system + the kernel's static shared memory

// kernel variants by tile size: lists in flight per request (registers) / largest tile in warps / CTAs per SM aimed at
struct FindVariant { int g; uint32_t max_warps, ctas; };
constexpr FindVariant FIND_VARIANTS[3] = {{8, 12, 2}, {8, 14, 2}, {16, TILE_WARPS_MAX, 1}};

// sub-tiles per search CTA for an index of n_sub sub-tiles: up to 14 in one tile; beyond that tiles of at most 12,
// balanced, so that two CTAs share an SM (one CTA's selection phase and barriers hide behind the other's counting;
// 500 k references: 11 tiles of 12 instead of 6 of 24)
inline uint32_t find_auto_tile_warps(uint32_t n_sub) {
    if (n_sub <= FIND_VARIANTS[1].max_warps) return n_sub ? n_sub : 1;
    const uint32_t n_tiles = (n_sub + 11) / 12;
    return (n_sub + n_tiles - 1) / n_tiles;
}

// scratch behind the counters: while counting, the staged offsets off[tile_warps + 1][ks] (ks = kc + 2 G + 4: the
// columns past kc stay zero, requests past the chunk see empty lists; + 4 rotates the banks between rows); while
// selecting, hist2 + tie
struct FindLayout {
    int variant;
    uint32_t kc;                // k-mers staged at a time: a multiple of 2 G, kc * (tile_warps + 1) <= FIND_PRE * threads
    uint32_t ks;                // row stride of the staged offsets in words (a multiple of 4)
    uint32_t scratch_words;
    size_t smem;                // dynamic shared memory of the launch
};
inline FindLayout find_layout(uint32_t tile_warps, uint32_t sub_size) {
    const uint32_t tw = tile_warps, ow = tw + 1, nt = 32 * tw;
    FindLayout L;
    L.variant = tw <= FIND_VARIANTS[0].max_warps ? 0 : tw <= FIND_VARIANTS[1].max_warps ? 1 : 2;
    const FindVariant& V = FIND_VARIANTS[L.variant];
    const uint32_t g2 = 2u * (uint32_t)V.g;
    const size_t counters = (size_t)tw * sub_size * 2;
    // what a CTA may use if V.ctas of them are to share an SM
    const size_t per_cta = std::min<size_t>(FIND_SMEM_CTA, FIND_SMEM_SM / V.ctas) - FIND_SMEM_FIXED;
    const size_t room = std::max<size_t>(per_cta > counters ? per_cta - counters : 0, (SEL_BINS + TIE_CAP) * 4) / 4;   // words
    uint32_t kc = std::min<uint32_t>(FIND_KC, FIND_PRE * nt / ow) & ~(g2 - 1u);
    while (kc > g2 && (size_t)ow * (kc + g2 + 4) > room) kc -= g2;
    L.kc = kc;
    L.ks = kc + g2 + 4;
    L.scratch_words = std::max<uint32_t>(ow * L.ks, SEL_BINS + TIE_CAP);
    L.smem = counters + (size_t)L.scratch_words * 4;
    return L;
}

// Top-k merge plan for a window of `max` candidates per tile: the merge sorts up to FIND_MAX_SORT keys per CTA in
// shared memory, so it takes the tiles in groups of *group tiles (one level when a single group holds them all) and
// merges the groups' winners in a second level. false: the window does not fit two levels either.
inline bool find_merge_plan(uint64_t max, uint32_t n_tiles, uint32_t* group, uint32_t* n_groups) {
    if (max == 0 || max > FIND_MAX_SORT) return false;
    const uint32_t gs = (uint32_t)std::min<uint64_t>(n_tiles, FIND_MAX_SORT / max);
    const uint32_t ng = (n_tiles + gs - 1) / gs;
    if (ng > 1 && (gs < 2 || (uint64_t)ng * max > FIND_MAX_SORT)) return false;
    if (group) *group = gs;
    if (n_groups) *n_groups = ng;
    return true;
}

}  // namespace sg
