// Aligner pre-steps and family-graph construction on the device.
//   contains_kernel / partition_kernel   aligner::operator() pre-steps (reference src/align.cpp:320-389)
//   graph_kernel                         mseq::mseq column sweep + dag::link + sort + reduce_edges
//                                        (src/mseq.cpp:47-118, src/graph.h:332-357,451-488, src/align.cpp:399-402)
//                                        followed by the DP plan (groups, ring/spill classification, arenas)
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace sg {

constexpr uint32_t NONE = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------------
// One warp per (query, relative): does the relative's base string contain the query's, ignoring case?
// (boost::algorithm::icontains on getBases(), src/align.cpp:329-332). Stores 1 + first offset, 0 = no.
__global__ void __launch_bounds__(128) contains_kernel(const uint8_t* __restrict__ qmasks,
                                                       const uint64_t* __restrict__ qoff, uint32_t nq,
                                                       const uint8_t* __restrict__ masks,
                                                       const uint64_t* __restrict__ row_off,
                                                       const uint32_t* __restrict__ fam_ids,
                                                       const int32_t* __restrict__ fam_n, uint32_t fam_cap,
                                                       uint32_t* __restrict__ contains) {
    const uint32_t pair = blockIdx.x * (blockDim.x >> 5) + warp_id();
    const uint32_t q = pair / fam_cap, j = pair % fam_cap;
    if (q >= nq) return;
    const int32_t F = fam_n[q];
    if (F <= 0 || j >= (uint32_t)F) return;
    const uint32_t id = fam_ids[(uint64_t)q * fam_cap + j];
    const uint8_t* r = masks + row_off[id];
    const uint32_t rl = (uint32_t)(row_off[id + 1] - row_off[id]);
    const uint8_t* m = qmasks + qoff[q];
    const uint32_t ql = (uint32_t)(qoff[q + 1] - qoff[q]);
    uint32_t found = NONE;
    if (ql <= rl) {
        for (uint32_t base = 0; base + ql <= rl && found == NONE; base += 32) {
            uint32_t p = base + lane_id();
            bool ok = p + ql <= rl;
            for (uint32_t t = 0; ok && t < ql; t++) ok = ((r[p + t] ^ m[t]) & 15u) == 0;
            uint32_t b = __ballot_sync(0xffffffffu, ok);
            if (b) found = base + (uint32_t)__ffs((int)b) - 1;
        }
    }
    if (lane_id() == 0) contains[(uint64_t)q * fam_cap + j] = found == NONE ? 0u : found + 1u;
}

// One thread per query: std::partition(vc, !contains) exactly as libstdc++'s bidirectional __partition
// permutes (bits/stl_algo.h:1472-1495; the reference's call is src/align.cpp:333), then --realign erase
// (:337-348) or the alignment copy decision (:349-388).
__global__ void partition_kernel(uint32_t nq, const uint32_t* __restrict__ fam_ids, const int32_t* __restrict__ fam_n,
                                 uint32_t fam_cap, uint32_t* __restrict__ contains, const uint64_t* __restrict__ qoff,
                                 const uint64_t* __restrict__ row_off, int realign, uint32_t* __restrict__ afam,
                                 uint32_t* __restrict__ afam_n, uint32_t* __restrict__ copy_src,
                                 GraphHdr* __restrict__ hdr) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    GraphHdr h = {};
    h.qlen = (uint32_t)(qoff[q + 1] - qoff[q]);
    const int32_t Fi = fam_n[q];
    uint32_t* a = afam + (uint64_t)q * fam_cap;
    uint32_t* c = contains + (uint64_t)q * fam_cap;
    if (Fi < 0) { h.status = SG_Q_NOFAMILY; afam_n[q] = 0; hdr[q] = h; return; }
    uint32_t F = (uint32_t)Fi;
    for (uint32_t j = 0; j < F; j++) a[j] = fam_ids[(uint64_t)q * fam_cap + j];
    uint32_t first = 0, last = F;
    for (;;) {
        for (;;) { if (first == last) goto done; else if (c[first] == 0) ++first; else break; }
        --last;
        for (;;) { if (first == last) goto done; else if (c[last] != 0) --last; else break; }
        { uint32_t t = a[first]; a[first] = a[last]; a[last] = t; t = c[first]; c[first] = c[last]; c[last] = t; }
        ++first;
    }
done:
    if (first != F) {
        if (realign) {
            F = first;
            if (F == 0) h.status = SG_Q_SKIPPED;
        } else {
            uint32_t src = first;
            for (uint32_t j = first; j < F; j++) {  // iequals: a containing relative of the same length
                uint32_t id = a[j];
                if ((uint32_t)(row_off[id + 1] - row_off[id]) == h.qlen) { src = j; break; }
            }
            copy_src[2 * q] = a[src];
            copy_src[2 * q + 1] = c[src] - 1;
            h.status = SG_Q_COPIED;
        }
    }
    afam_n[q] = F;
    hdr[q] = h;
}

// ---------------------------------------------------------------------------------------------------
struct GraphArgs {
    const uint8_t* masks; const uint32_t* cols; const uint64_t* row_off; uint32_t W;
    const uint32_t* afam; const uint32_t* afam_n; uint32_t fam_cap;
    uint32_t icap, ncap, gcap, q0;  // q0: first query of the chunk (workspace arrays are chunk-local)
    uint32_t itemcap;               // stride of the per-item arrays (generic path)
    GraphHdr* hdr;
    uint8_t *tab, *tabli; uint32_t *colof, *colbase, *item_node, *slot;
    uint32_t* ncol; uint8_t* nmask; uint16_t* ncount; float* nweight; uint32_t* nsigma;
    uint32_t *slotbase, *cursor, *pred_off, *preds, *pdesc; int32_t* spillrow; uint8_t* nflags;
    uint32_t* lastnodes; GroupInfo* groups;
    uint32_t* nmaxins; int forbid;   // --insertion forbid
    uint32_t* order; rcol_t* rcol; uint16_t* nthr; uint32_t* pdesc2; GhostInfo* ghosts; uint32_t* writers; int force_generic;
    float pen_max, pen_min;   // larger / smaller of the gap open and gap extension penalties
    unsigned long long* cells; unsigned long long* cursors; uint64_t tb_words, spill_elems;
    float fs_weight;
    uint32_t stab_bytes;   // shared-memory budget of the column table (fast path of steps 2-5)
};

// running-carry exclusive scan of arr[0..n) (u32, global or shared), in place -> exclusive; returns total
template <typename T>
__device__ uint32_t scan_array_inplace(T* arr, uint32_t n, uint32_t* red) {
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n ? (uint32_t)arr[i] : 0, tot;
        uint32_t ex = block_exscan(v, red, &tot);
        if (i < n) arr[i] = (T)(carry + ex);
        carry += tot;
    }
    __syncthreads();
    return carry;
}


// ---- per-column helpers of the shared-memory path (one thread per column of the family's column table) ----
constexpr int GK = 6;   // distinct (local node, predecessor) pairs of a column kept in registers; more take the local-memory path
// predecessor node of row j's base in column c: the node of the row's previous base (NONE: c holds the row's first base)
// (the table holds one NIBBLE per (column, family row): Fh = (F + 1) / 2 bytes per column, row j in the low (even j) or
// high (odd j) half of byte j / 2)
__device__ __forceinline__ uint32_t prev_node(const uint8_t* stab, const uint32_t* scolbase, uint32_t Fh, uint32_t c, uint32_t j) {
    uint32_t r = c, pe = 0;
    const uint32_t jb = j >> 1, sh = (j & 1u) * 4u;
    while (r > 0) { r--; pe = (stab[r * Fh + jb] >> sh) & 15u; if (pe) break; }
    return pe ? scolbase[r] + pe - 1u : NONE;
}
// the column's distinct keys (local node << 24 | predecessor), ascending, into k[GK]; returns their number, or GK + 1
// if there are more than GK (k is then incomplete)
__device__ __forceinline__ uint32_t column_keys(const uint8_t* stab, const uint32_t* scolbase, uint32_t F, uint32_t c, uint32_t (&k)[GK]) {
#pragma unroll
    for (int i = 0; i < GK; i++) k[i] = NONE;
    uint32_t nu = 0, last = NONE;
    const uint32_t Fh = (F + 1u) >> 1;
    const uint8_t* t = stab + c * Fh;
    // the previous base of a row nearly always sits in the previous column: its byte is fetched together with the
    // column's own, the walk back through the table (prev_node) is left to the rows with a gap there
    const uint8_t* tp = c ? t - Fh : t;
    const uint32_t pbase = c ? scolbase[c - 1] : 0u;
    for (uint32_t jb = 0; jb < Fh; jb++) {
        const uint32_t cur = t[jb];
        if (!cur) continue;
        const uint32_t prv = c ? tp[jb] : 0u;
#pragma unroll
        for (uint32_t half = 0; half < 2; half++) {
            const uint32_t e = (cur >> (4u * half)) & 15u;
            if (!e) continue;
            const uint32_t pe = (prv >> (4u * half)) & 15u;
            const uint32_t p = pe ? pbase + pe - 1u : (c ? prev_node(stab, scolbase, Fh, c - 1u, 2u * jb + half) : NONE);
            if (p == NONE) continue;
            const uint32_t key = ((e - 1u) << 24) | p;
            if (key == last) continue;            // most rows of a family run through the same pair of nodes
            last = key;
            bool dup = false;
#pragma unroll
            for (int i = 0; i < GK; i++) dup |= k[i] == key;
            if (dup) continue;
            if (nu >= (uint32_t)GK) return GK + 1;
#pragma unroll
            for (int i = 0; i < GK; i++) if ((uint32_t)i == nu) k[i] = key;
            nu++;
        }
    }
    // sorting network for 6 keys (unused slots hold NONE = the largest value)
    static_assert(GK == 6, "sorting network below");
#define GK_CX(a, b) { const uint32_t lo_ = min(k[a], k[b]), hi_ = max(k[a], k[b]); k[a] = lo_; k[b] = hi_; }
    GK_CX(0, 5) GK_CX(1, 3) GK_CX(2, 4) GK_CX(1, 2) GK_CX(3, 4) GK_CX(0, 3) GK_CX(2, 5) GK_CX(0, 1) GK_CX(2, 3) GK_CX(4, 5) GK_CX(1, 2) GK_CX(3, 4)
#undef GK_CX
    return nu;
}
// the same for a column with many distinct pairs: sorted list of distinct keys in local memory; returns their number
__device__ __noinline__ uint32_t column_keys_many(const uint8_t* stab, const uint32_t* scolbase, uint32_t F, uint32_t c, uint32_t* u) {
    uint32_t nu = 0;
    const uint32_t Fh = (F + 1u) >> 1;
    const uint8_t* t = stab + c * Fh;
    for (uint32_t j = 0; j < F; j++) {
        const uint32_t e = (t[j >> 1] >> ((j & 1u) * 4u)) & 15u;
        if (!e) continue;
        const uint32_t p = prev_node(stab, scolbase, Fh, c, j);
        if (p == NONE) continue;
        const uint32_t key = ((e - 1u) << 24) | p;
        uint32_t y = nu;
        while (y > 0 && u[y - 1] > key) y--;
        if (y > 0 && u[y - 1] == key) continue;
        for (uint32_t z = nu; z > y; z--) u[z] = u[z - 1];
        u[y] = key;
        nu++;
    }
    return nu;
}

__global__ void __launch_bounds__(GRAPH_BLOCK) graph_kernel(GraphArgs A) {
    extern __shared__ uint32_t sm[];
    const uint32_t q = A.q0 + blockIdx.x;  // global query index (hdr, family); workspace uses the local index
    const uint32_t ql = blockIdx.x;
    GraphHdr* hdr = A.hdr + q;
    if (hdr->status == GS_ARENA_FULL) { if (threadIdx.x == 0) hdr->status = GS_OK; }  // retry pass, arenas were reset
    else if (hdr->status != GS_OK) return;
    __syncthreads();
    const uint32_t words = (A.W + 31) >> 5;
    uint32_t* bitmap = sm;            // [words]   columns used by any family row
    uint32_t* wrank = sm + words;     // [words]   exclusive popcount prefix
    uint32_t* famoff = wrank + words; // [fam_cap+1] item offsets of the family rows
    uint32_t* red = famoff + A.fam_cap + 1;  // [33]
    uint32_t* shv = red + 33;         // [8] misc broadcast
    const uint32_t F = A.afam_n[q];
    const uint32_t* fam = A.afam + (uint64_t)q * A.fam_cap;
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
    const uint32_t Lq = hdr->qlen;

    // per-query views
    uint8_t* tab = A.tab + (uint64_t)ql * A.ncap * A.fam_cap;
    uint8_t* tabli = A.tabli + (uint64_t)ql * A.ncap * A.fam_cap;
    uint32_t* colof = A.colof + (uint64_t)ql * A.ncap;
    uint32_t* colbase = A.colbase + (uint64_t)ql * (A.ncap + 1);
    const uint64_t io = (uint64_t)ql * A.icap;
    uint32_t* item_node = A.item_node + (uint64_t)ql * A.itemcap;
    uint32_t* slot = A.slot + (uint64_t)ql * A.itemcap;
    uint32_t* ncol = A.ncol + io; uint8_t* nmask = A.nmask + io; uint16_t* ncount = A.ncount + io;
    float* nweight = A.nweight + io; uint32_t* nsigma = A.nsigma + io;
    uint32_t* slotbase = A.slotbase + (uint64_t)ql * (A.icap + 1);
    uint32_t* cursor = A.cursor + io;
    uint32_t* pred_off = A.pred_off + (uint64_t)ql * (A.icap + 1);
    uint32_t* preds = A.preds + io; uint32_t* pdesc = A.pdesc + io;
    int32_t* spillrow = A.spillrow + io; uint8_t* nflags = A.nflags + io;
    uint32_t* lastnodes = A.lastnodes + io;
    GroupInfo* groups = A.groups + (uint64_t)ql * A.gcap;

    // ---- 0. item offsets of the family rows
    for (uint32_t i = tid; i < words; i += nt) bitmap[i] = 0;
    uint32_t* rb32 = shv + 8 + 2 * FARLIST_CAP + 2 * DP_G + DP_T;
    if (reinterpret_cast<uintptr_t>(rb32) & 7u) rb32++;
    uint64_t* rowbase = reinterpret_cast<uint64_t*>(rb32);  // [fam_cap] first item of family row j in the index
    for (uint32_t j = tid; j < F; j += nt) {
        const uint64_t a = A.row_off[fam[j]];
        rowbase[j] = a;
        famoff[j] = (uint32_t)(A.row_off[fam[j] + 1] - a);
    }
    __syncthreads();
    const uint32_t I = scan_array_inplace(famoff, F, red);
    if (tid == 0) famoff[F] = I;
    __syncthreads();
    if (I > A.itemcap || F == 0) { if (tid == 0) hdr->status = F == 0 ? (uint32_t)SG_Q_SKIPPED : GS_LIMIT; return; }
    // item passes: one warp per family row at a time, 4 items per lane and round so that the global loads of a
    // round are all in flight together. ROW_ITEMS(body) runs body(x, at, u) for item x (flat index over the family),
    // `at` = its position in the index arrays.
    const uint32_t wid = warp_id(), nwarp = nt >> 5, lane = lane_id();
    // ---- 1. used-column bitmap and column ranks (= the sweep's `min_next` column order, mseq.cpp:76-84)
    for (uint32_t j = wid; j < F; j += nwarp) {
        const uint64_t a = rowbase[j];
        const uint32_t len = famoff[j + 1] - famoff[j];
        for (uint32_t i0 = lane; i0 < len; i0 += 128) {
            uint32_t c[4];
#pragma unroll
            for (int u = 0; u < 4; u++) c[u] = i0 + 32 * u < len ? A.cols[a + i0 + 32 * u] : NONE;
#pragma unroll
            for (int u = 0; u < 4; u++) if (c[u] != NONE) atomicOr(&bitmap[c[u] >> 5], 1u << (c[u] & 31));
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < words; i += nt) wrank[i] = __popc(bitmap[i]);
    __syncthreads();
    const uint32_t n_cols = scan_array_inplace(wrank, words, red);
    if (n_cols > A.ncap) { if (tid == 0) hdr->status = GS_LIMIT; return; }
    for (uint32_t w = tid; w < words; w += nt) {
        uint32_t bits = bitmap[w], r = wrank[w];
        while (bits) { uint32_t b = __ffs((int)bits) - 1; colof[r++] = (w << 5) + b; bits &= bits - 1; }
    }
    auto colrank = [&](uint32_t c) -> uint32_t { return wrank[c >> 5] + __popc(bitmap[c >> 5] & ((1u << (c & 31)) - 1u)); };

    uint32_t my_max = 0;
    uint32_t E = 0, shv_V = 0;
    // Steps 2-5 (column table, nodes, edges). Fast path: the column table of the family (n_cols x F bytes, ~64 KB for
    // 40 full-length rows) lives in SHARED memory and the nodes, their per-row local indices and the de-duplicated,
    // sorted predecessor lists are all derived from it, one thread per column; global memory only sees the item reads
    // and the final node / edge arrays. (The generic path below keeps tab / tabli / item_node / slot in global
    // scratch: 3.65 MB of DRAM traffic per query against ~0.4 MB algorithmic, 22 long-scoreboard stalls per issue.)
    uint8_t* stab = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(rowbase + A.fam_cap) + 15u) & ~(uintptr_t)15u);   // [n_cols][Fh] nibbles: base masks, then local node index + 1
    const uint32_t Fh = (F + 1u) >> 1;
    bool fast = (uint64_t)n_cols * Fh <= A.stab_bytes && n_cols + 1 <= 2 * words && F <= FAM_CAP_MAX;
    if (fast) {
        // ---- 2. column table (one nibble per entry: a family with lower-case bases, whose nodes differ by case, takes
        //         the generic path). A warp fills the two rows sharing the table's bytes one after the other.
        {
            uint4* z = reinterpret_cast<uint4*>(stab);
            const uint32_t nz = (n_cols * Fh + 15) >> 4;
            for (uint32_t i = tid; i < nz; i += nt) z[i] = make_uint4(0, 0, 0, 0);
            if (tid == 0) shv[5] = 0;
        }
        __syncthreads();
        uint32_t lower = 0;
        for (uint32_t jp = wid; jp < Fh; jp += nwarp) {
            for (uint32_t half = 0; half < 2; half++) {
                const uint32_t j = 2 * jp + half;
                if (j >= F) break;
                const uint64_t a = rowbase[j];
                const uint32_t len = famoff[j + 1] - famoff[j];
                for (uint32_t i0 = lane; i0 < len; i0 += 128) {
                    uint32_t c[4];
                    uint8_t mk[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        c[u] = NONE;
                        if (i0 + 32 * u < len) { c[u] = A.cols[a + i0 + 32 * u]; mk[u] = A.masks[a + i0 + 32 * u]; }
                    }
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        if (c[u] == NONE) continue;
                        lower |= mk[u] & 16u;
                        uint8_t* e = stab + colrank(c[u]) * Fh + jp;
                        *e = (uint8_t)(*e | ((mk[u] & 15u) << (4 * half)));
                    }
                }
                __syncwarp();
            }
        }
        if (lower) shv[5] = 1;
        __syncthreads();   // bitmap / wrank are dead from here on
        if (shv[5]) fast = false;
    }
    if (fast) {
        uint32_t* scolbase = bitmap;   // [n_cols + 1], over the bitmap / rank arrays once the table is filled
        // ---- 3. nodes: one per (column, IUPAC char), local order = first family row bringing it (mseq.cpp:88-98).
        //         Pass A counts per column, pass B writes node arrays and turns the table entries into local node index + 1.
        auto count_column = [&](uint32_t c) -> uint32_t {
            uint32_t seen = 0, nn = 0;
            const uint8_t* t = stab + c * Fh;
            for (uint32_t jb = 0; jb < Fh; jb++) {
                const uint32_t byte = t[jb], lo = byte & 15u, hi = byte >> 4;
                if (lo && !((seen >> lo) & 1u)) { seen |= 1u << lo; nn++; }
                if (hi && !((seen >> hi) & 1u)) { seen |= 1u << hi; nn++; }
            }
            return nn;
        };
        uint32_t cnt_mine[8];   // nodes of the (up to 8) columns this thread owns: columns tid, tid + nt, ...
        {
            uint32_t q8 = 0;
            for (uint32_t c = tid; c < n_cols; c += nt, q8++) { const uint32_t nn = count_column(c); if (q8 < 8) cnt_mine[q8] = nn; }
        }
        __syncthreads();
        {
            uint32_t q8 = 0;
            for (uint32_t c = tid; c < n_cols; c += nt, q8++) scolbase[c] = q8 < 8 ? cnt_mine[q8] : count_column(c);
        }
        __syncthreads();
        const uint32_t V = scan_array_inplace(scolbase, n_cols, red);
        if (tid == 0) scolbase[n_cols] = V;
        if (V > A.icap) { if (tid == 0) hdr->status = GS_LIMIT; return; }
        __syncthreads();
        uint32_t my_masks = 0;   // IUPAC masks of the nodes this thread creates
        if (tid == 0) shv[4] = 0;
        for (uint32_t c = tid; c < n_cols; c += nt) {
            uint8_t* t = stab + c * Fh;
            const uint32_t base = scolbase[c], nn = scolbase[c + 1] - base, col = colof[c];
            colbase[c] = base;
            auto emit = [&](uint32_t k2, uint32_t mask, uint32_t count) {
                const uint32_t m = base + k2;
                my_masks |= 1u << (mask & 15u);
                ncol[m] = col; nmask[m] = (uint8_t)mask; ncount[m] = (uint16_t)count; nsigma[m] = c;
                // node->weight = 1.0/(weight+1) + weight * (node->weight/num_seqs)   (mseq.cpp:111-116)
                const float fr = __fdiv_rn((float)count, (float)F);
                const float b2 = __fmul_rn(A.fs_weight, fr);
                nweight[m] = (float)(1.0 / (double)__fadd_rn(A.fs_weight, 1.0f) + (double)b2);
                nflags[m] = 0; spillrow[m] = 0;
            };
            if (nn <= 4) {   // the usual column: its (up to four) characters and their counts stay in registers
                uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0, have = 0;
                auto place = [&](uint32_t bb) -> uint32_t {   // local node index + 1 of a base, 0 for none
                    if (!bb) return 0u;
                    if (have > 0 && bb == m0) { c0++; return 1u; }
                    if (have > 1 && bb == m1) { c1++; return 2u; }
                    if (have > 2 && bb == m2) { c2++; return 3u; }
                    if (have > 3 && bb == m3) { c3++; return 4u; }
                    if (have == 0) { m0 = bb; c0 = 1; } else if (have == 1) { m1 = bb; c1 = 1; }
                    else if (have == 2) { m2 = bb; c2 = 1; } else { m3 = bb; c3 = 1; }
                    return ++have;
                };
                for (uint32_t jb = 0; jb < Fh; jb++) {
                    const uint32_t byte = t[jb];
                    if (!byte) continue;
                    const uint32_t lo = place(byte & 15u), hi = place(byte >> 4);
                    t[jb] = (uint8_t)(lo | (hi << 4));
                }
                if (nn > 0) emit(0, m0, c0);
                if (nn > 1) emit(1, m1, c1);
                if (nn > 2) emit(2, m2, c2);
                if (nn > 3) emit(3, m3, c3);
            } else {
                uint32_t seen = 0, k3 = 0;
                uint8_t li_of[16];
                uint16_t cnt[16];
                uint8_t nm[16];
                auto place = [&](uint32_t bb) -> uint32_t {
                    if (!bb) return 0u;
                    if (!((seen >> bb) & 1u)) { seen |= 1u << bb; li_of[bb] = (uint8_t)k3; cnt[k3] = 1; nm[k3] = (uint8_t)bb; k3++; }
                    else cnt[li_of[bb]]++;
                    return li_of[bb] + 1u;
                };
                for (uint32_t jb = 0; jb < Fh; jb++) {
                    const uint32_t byte = t[jb];
                    if (!byte) continue;
                    const uint32_t lo = place(byte & 15u), hi = place(byte >> 4);
                    t[jb] = (uint8_t)(lo | (hi << 4));
                }
                for (uint32_t k2 = 0; k2 < nn; k2++) emit(k2, nm[k2], cnt[k2]);
            }
        }
        if (tid == 0) colbase[n_cols] = V;
        __syncthreads();
        if (my_masks) atomicOr(&shv[4], my_masks);
        // ---- 4 + 5. edges: the predecessor of (column c, row j) is the node of row j's previous base (dag::link,
        //         graph.h:332-340); per node the distinct predecessors in ascending id (reduce_edges, graph.h:466-488).
        //         A thread collects its column's (local node, predecessor) pairs as a sorted list of distinct keys
        //         (pass 0: in-degree per node; pass 1, after the scan: the predecessor ids)
        const bool keep_keys = (uint64_t)n_cols * GK <= A.itemcap;   // `slot` (global scratch of the generic path) is free here
        for (int pass = 0; pass < 2; pass++) {
            for (uint32_t c = tid; c < n_cols; c += nt) {
                const uint32_t base = scolbase[c], nn = scolbase[c + 1] - base, cm = colof[c];
                uint32_t k[GK];
                uint32_t nu;
                if (pass == 0 || !keep_keys) {
                    nu = column_keys(stab, scolbase, F, c, k);
                    if (pass == 0 && keep_keys) {   // pass 1 reads the column's keys back instead of deriving them again
#pragma unroll
                        for (int x = 0; x < GK; x++) slot[(uint64_t)c * GK + x] = nu <= (uint32_t)GK ? k[x] : 0xFFFFFFFEu;
                    }
                } else {
                    nu = 0;
#pragma unroll
                    for (int x = 0; x < GK; x++) { k[x] = slot[(uint64_t)c * GK + x]; nu += k[x] != NONE ? 1u : 0u; }
                    if (k[0] == 0xFFFFFFFEu) nu = GK + 1;
                }
                // one run of equal local node per node with predecessors: in-degree = run length (pass 0), the ids (pass 1)
                auto edge = [&](uint32_t key, uint32_t o) {   // pass 1: predecessor number o (global index) of its node
                    const uint32_t p = key & 0xffffffu;
                    preds[o] = p;
                    nflags[p] = 1;  // has a successor (benign same-value race)
                    if (A.forbid) atomicMin(&A.nmaxins[io + p], cm);
                };
                if (pass == 0) for (uint32_t k2 = 0; k2 < nn; k2++) pred_off[base + k2] = 0;
                if (nu <= (uint32_t)GK) {
                    uint32_t run = 0, o = 0;
#pragma unroll
                    for (int x = 0; x < GK; x++) {
                        if ((uint32_t)x < nu) {
                            const uint32_t li = k[x] >> 24;
                            const bool first = x == 0 || (k[x - (x > 0)] >> 24) != li;
                            if (pass == 0) {
                                run = first ? 1u : run + 1u;
                                const bool lastofrun = (uint32_t)x + 1 == nu || (k[x + (x + 1 < GK)] >> 24) != li;
                                if (lastofrun) { pred_off[base + li] = run; my_max = max(my_max, run); }
                            } else {
                                o = first ? pred_off[base + li] : o + 1u;
                                edge(k[x], o);
                            }
                        }
                    }
                } else {
                    uint32_t u[FAM_CAP_MAX + 1];
                    nu = column_keys_many(stab, scolbase, F, c, u);
                    uint32_t x = 0;
                    while (x < nu) {
                        const uint32_t li = u[x] >> 24;
                        uint32_t x2 = x;
                        while (x2 < nu && (u[x2] >> 24) == li) x2++;
                        if (pass == 0) { pred_off[base + li] = x2 - x; my_max = max(my_max, x2 - x); }
                        else { const uint32_t o = pred_off[base + li]; for (uint32_t y = x; y < x2; y++) edge(u[y], o + (y - x)); }
                        x = x2;
                    }
                }
            }
            __syncthreads();
            if (pass == 0) {
                E = scan_array_inplace(pred_off, V, red);
                if (E > A.icap) { if (tid == 0) hdr->status = GS_LIMIT; return; }
                if (tid == 0) pred_off[V] = E;
                if (A.forbid) for (uint32_t m = tid; m < V; m += nt) A.nmaxins[io + m] = 1000000u;
                __syncthreads();
            }
        }
        if (A.forbid) {
            // --insertion forbid: max_insert of a node = columns free before its nearest successor
            // (compute_node_simple::calc, src/mesh.h:480-484: min over next nodes of their position, 1000000 without one)
            uint32_t* nmaxins = A.nmaxins + io;
            for (uint32_t m = tid; m < V; m += nt) nmaxins[m] = nmaxins[m] - ncol[m] - 1u;
        }
        shv_V = V;
    } else {
        // ---- 2. column table: tab[c][j] = base mask of family row j in column rank c
        for (uint32_t i = tid; i < n_cols * A.fam_cap; i += nt) tab[i] = 0;
        __syncthreads();
        // (the column rank of every item is kept in item_node until the nodes exist)
        for (uint32_t j = wid; j < F; j += nwarp) {
            const uint64_t a = rowbase[j];
            const uint32_t fo = famoff[j], len = famoff[j + 1] - fo;
            for (uint32_t i0 = lane; i0 < len; i0 += 128) {
                uint32_t c[4];
                uint8_t mk[4];
    #pragma unroll
                for (int u = 0; u < 4; u++) {
                    c[u] = NONE;
                    if (i0 + 32 * u < len) { c[u] = A.cols[a + i0 + 32 * u]; mk[u] = A.masks[a + i0 + 32 * u]; }
                }
    #pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (c[u] == NONE) continue;
                    const uint32_t r = colrank(c[u]);
                    tab[(uint64_t)r * A.fam_cap + j] = mk[u] & 31u;
                    item_node[fo + i0 + 32 * u] = r;
                }
            }
        }
        __syncthreads();

        // ---- 3. nodes: one per (column, IUPAC char incl. case), local order = first family row bringing it
        //         (mseq.cpp:88-98). Pass A counts per column, pass B writes node arrays.
        for (uint32_t c = tid; c < n_cols; c += nt) {
            uint32_t seen = 0, nn = 0;
            const uint8_t* t = tab + (uint64_t)c * A.fam_cap;
            for (uint32_t j = 0; j < F; j++) { uint32_t b = t[j]; if (b && !((seen >> b) & 1u)) { seen |= 1u << b; nn++; } }
            colbase[c] = nn;
        }
        __syncthreads();
        const uint32_t V = scan_array_inplace(colbase, n_cols, red);
        shv_V = V;
        if (tid == 0) colbase[n_cols] = V;
        if (V > A.icap) { if (tid == 0) hdr->status = GS_LIMIT; return; }
        uint32_t my_masks = 0;   // IUPAC masks (case ignored) of the nodes this thread creates
        if (tid == 0) shv[4] = 0;
        for (uint32_t c = tid; c < n_cols; c += nt) {
            uint32_t seen = 0, nn = 0;
            uint8_t li_of[32];
            uint16_t cnt[32];
            uint8_t nm[32];
            const uint8_t* t = tab + (uint64_t)c * A.fam_cap;
            uint8_t* tl = tabli + (uint64_t)c * A.fam_cap;
            for (uint32_t j = 0; j < F; j++) {
                uint32_t b = t[j];
                if (!b) continue;
                if (!((seen >> b) & 1u)) { seen |= 1u << b; li_of[b] = (uint8_t)nn; cnt[nn] = 1; nm[nn] = (uint8_t)b; nn++; }
                else cnt[li_of[b]]++;
                tl[j] = li_of[b];
            }
            const uint32_t base = colbase[c], col = colof[c];
            for (uint32_t k2 = 0; k2 < nn; k2++) {
                const uint32_t m = base + k2;
                my_masks |= 1u << (nm[k2] & 15u);
                ncol[m] = col; nmask[m] = nm[k2]; ncount[m] = cnt[k2]; nsigma[m] = c;
                // node->weight = 1.0/(weight+1) + weight * (node->weight/num_seqs)   (mseq.cpp:111-116)
                float fr = __fdiv_rn((float)cnt[k2], (float)F);
                float b2 = __fmul_rn(A.fs_weight, fr);
                nweight[m] = (float)(1.0 / (double)__fadd_rn(A.fs_weight, 1.0f) + (double)b2);
                cursor[m] = 0; nflags[m] = 0; spillrow[m] = 0;
                slotbase[m] = cnt[k2];
            }
        }
        __syncthreads();
        if (my_masks) atomicOr(&shv[4], my_masks);

        // ---- 4. node of every item; predecessor candidates grouped per node (dag::link, graph.h:332-340)
        scan_array_inplace(slotbase, V, red);
        if (tid == 0) slotbase[V] = I;
        for (uint32_t j = wid; j < F; j += nwarp) {
            const uint32_t fo = famoff[j], len = famoff[j + 1] - fo;
            for (uint32_t i0 = lane; i0 < len; i0 += 128) {
                uint32_t r[4], cb[4], li[4];
    #pragma unroll
                for (int u = 0; u < 4; u++) r[u] = i0 + 32 * u < len ? item_node[fo + i0 + 32 * u] : NONE;
    #pragma unroll
                for (int u = 0; u < 4; u++) if (r[u] != NONE) { cb[u] = colbase[r[u]]; li[u] = tabli[(uint64_t)r[u] * A.fam_cap + j]; }
    #pragma unroll
                for (int u = 0; u < 4; u++) if (r[u] != NONE) item_node[fo + i0 + 32 * u] = cb[u] + li[u];
            }
        }
        __syncthreads();
        for (uint32_t j = wid; j < F; j += nwarp) {
            const uint32_t fo = famoff[j], len = famoff[j + 1] - fo;
            for (uint32_t i0 = lane; i0 < len; i0 += 128) {
                uint32_t node[4], from[4], pos[4], sb[4];
    #pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + 32 * u;
                    node[u] = NONE;
                    if (i < len) {
                        node[u] = item_node[fo + i];
                        from[u] = i ? item_node[fo + i - 1] : NONE;   // previous item of the same family row
                    }
                }
    #pragma unroll
                for (int u = 0; u < 4; u++) if (node[u] != NONE) { pos[u] = atomicAdd(&cursor[node[u]], 1u); sb[u] = slotbase[node[u]]; }
    #pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (node[u] == NONE) continue;
                    slot[sb[u] + pos[u]] = from[u];
                    if (from[u] != NONE) nflags[from[u]] = 1;  // has a successor (benign same-value race)
                }
            }
        }
        __syncthreads();

        // ---- 5. reduce_edges: sort predecessor ids, drop duplicates (graph.h:466-488); in-degree per node
        for (uint32_t m = tid; m < V; m += nt) {
            uint32_t* sl = slot + slotbase[m];
            const uint32_t n = slotbase[m + 1] - slotbase[m];
            // sorted list of the distinct predecessors, built in thread-local memory (n <= family size, and most of the
            // n candidates repeat one of a handful of nodes)
            uint32_t u[FAM_CAP_MAX + 1];
            uint32_t deg = 0;
            for (uint32_t x0 = 0; x0 < n; x0 += 4) {
                uint32_t key[4];
    #pragma unroll
                for (int k2 = 0; k2 < 4; k2++) key[k2] = x0 + k2 < n ? sl[x0 + k2] : NONE;
    #pragma unroll
                for (int k2 = 0; k2 < 4; k2++) {
                    if (key[k2] == NONE) continue;
                    uint32_t y = deg;
                    while (y > 0 && u[y - 1] > key[k2]) y--;
                    if (y > 0 && u[y - 1] == key[k2]) continue;
                    for (uint32_t z = deg; z > y; z--) u[z] = u[z - 1];
                    u[y] = key[k2];
                    deg++;
                }
            }
            for (uint32_t x = 0; x < deg; x++) sl[x] = u[x];
            pred_off[m] = deg;
            my_max = max(my_max, deg);
        }
        __syncthreads();
        E = scan_array_inplace(pred_off, V, red);
        if (E > A.icap) { if (tid == 0) hdr->status = GS_LIMIT; return; }
        if (tid == 0) pred_off[V] = E;
        __syncthreads();
        for (uint32_t m = tid; m < V; m += nt) {
            const uint32_t deg = pred_off[m + 1] - pred_off[m];
            for (uint32_t x = 0; x < deg; x++) preds[pred_off[m] + x] = slot[slotbase[m] + x];
        }
        if (A.forbid) {
            // --insertion forbid: max_insert of a node = columns free before its nearest successor
            // (compute_node_simple::calc, src/mesh.h:480-484: min over next nodes of their position, 1000000 without one)
            uint32_t* nmaxins = A.nmaxins + io;
            for (uint32_t m = tid; m < V; m += nt) nmaxins[m] = 1000000u;
            __syncthreads();
            for (uint32_t m = tid; m < V; m += nt) {
                const uint32_t deg = pred_off[m + 1] - pred_off[m], cm = ncol[m];
                for (uint32_t x = 0; x < deg; x++) atomicMin(&nmaxins[slot[slotbase[m] + x]], cm);
            }
            __syncthreads();
            for (uint32_t m = tid; m < V; m += nt) nmaxins[m] = nmaxins[m] - ncol[m] - 1u;
        }
    }
    const uint32_t V = shv_V;
    // block max of in-degree
    for (int o = 16; o > 0; o >>= 1) my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, o));
    __syncthreads();
    if (lane_id() == 0) red[warp_id()] = my_max;
    __syncthreads();
    if (tid == 0) { uint32_t mx = 0; for (uint32_t w = 0; w < (nt >> 5); w++) mx = max(mx, red[w]); shv[0] = mx; }
    __syncthreads();
    const uint32_t max_indeg = shv[0];
    const uint32_t wide = max_indeg > 8;

    // ---- 6. DP plan. Group g = nodes [g*T, (g+1)*T); a node at column rank sigma runs query position
    // s at step s + sigma - sigma_lo(g). An edge is "near" (served from the shared-memory ring) when both
    // ends are in the same group and at most DP_RING-2 column ranks apart; otherwise the predecessor row
    // is spilled to global memory ("far").
    const uint32_t T = DP_T;
    static_assert(DP_T % 8 == 0 && GRAPH_BLOCK >= DP_T, "the v2 plan below gives every row of a group its own thread");
    const uint32_t n_groups = (V + T - 1) / T;
    if (n_groups > A.gcap) { if (tid == 0) hdr->status = GS_LIMIT; return; }
    for (uint32_t m = tid; m < V; m += nt) {
        const uint32_t g = m / T, sg_ = nsigma[m];
        for (uint32_t e = pred_off[m]; e < pred_off[m + 1]; e++) {
            const uint32_t p = preds[e], d = sg_ - nsigma[p];
            if (p / T == g && d <= (uint32_t)DP_RING - 2) pdesc[e] = (d << 16) | (p - g * T);
            else pdesc[e] = FAR_BIT;
            if (p / T != g || d > (uint32_t)DP_MAXD) spillrow[p] = 1;   // the v2 kernel's ring reaches DP_MAXD ranks
        }
    }
    __syncthreads();
    const uint32_t n_spill = scan_array_inplace(spillrow, V, red);  // exclusive index for flagged rows
    // rows not flagged must read -1: flagged iff next prefix differs
    for (uint32_t m = tid; m < V; m += nt) {
        uint32_t nxt = (m + 1 < V) ? (uint32_t)spillrow[m + 1] : n_spill;
        cursor[m] = (nxt != (uint32_t)spillrow[m]) ? (uint32_t)spillrow[m] : NONE;
    }
    __syncthreads();
    for (uint32_t m = tid; m < V; m += nt) spillrow[m] = (int32_t)cursor[m];
    __syncthreads();
    for (uint32_t e = tid; e < E; e += nt) if (pdesc[e] == FAR_BIT) pdesc[e] = FAR_BIT | (uint32_t)spillrow[preds[e]];
    // last nodes (no successor), ascending id: sentinel's _previous list (graph.h:332-357)
    for (uint32_t m = tid; m < V; m += nt) cursor[m] = nflags[m] ? 0u : 1u;
    __syncthreads();
    const uint32_t n_last = scan_array_inplace(cursor, V, red);
    for (uint32_t m = tid; m < V; m += nt) if (!nflags[m]) lastnodes[cursor[m]] = m;

    // ---- 7. plan for the v2 kernel: inside a group rows are sorted by in-degree (warps become uniform),
    // every far predecessor is served by a "ghost" ring column that a loader lane streams from the spill
    // buffer, and rows that must be spilled (far successors) or min-tracked (last nodes) get a writer lane.
    uint32_t* order = A.order + (uint64_t)ql * A.gcap * T;
    uint16_t* nthr = A.nthr + io;
    uint32_t* pdesc2 = A.pdesc2 + io;
    GhostInfo* ghosts = A.ghosts + (uint64_t)ql * A.gcap * DP_G;
    uint32_t* writers = A.writers + (uint64_t)ql * A.gcap * DP_G;
    uint32_t* far_e = shv + 8;                   // [FARLIST_CAP] edge index
    uint32_t* far_gi = far_e + FARLIST_CAP;      // [FARLIST_CAP] ghost index of the edge
    uint32_t* gh_p = far_gi + FARLIST_CAP;       // [DP_G] ghost source node
    uint32_t* gh_b = gh_p + DP_G;                // [DP_G] ghost bucket
    uint32_t* rwant = gh_b + DP_G;               // [DP_T] ring column wish, then ring column, of the row at a thread
    if (tid == 0) shv[1] = 0;                    // overflow flag
    __syncthreads();
    for (uint32_t g = 0; g < n_groups; g++) {
        const uint32_t lo = g * T, n = min(T, V - lo);
        const uint32_t sigma_lo = nsigma[lo];
        const bool valid = tid < n;
        const uint32_t m = lo + tid;
        uint32_t np = 0;
        if (valid) np = pred_off[m + 1] - pred_off[m];
        const uint32_t key = !valid ? 0u : (np == 0 ? 1u : min(np, 5u));
        uint32_t base = 0;
        for (uint32_t b = 1; b <= 5; b++) {  // stable counting sort by in-degree class
            uint32_t tot;
            const uint32_t ex = block_exscan(key == b ? 1u : 0u, red, &tot);
            if (key == b) { nthr[m] = (uint16_t)(base + ex); order[(uint64_t)g * T + base + ex] = m; }
            base += tot;
        }
        if (tid >= n && tid < T) order[(uint64_t)g * T + tid] = NONE;
        if (tid < T) rwant[tid] = NONE;
        if (tid == 0) { shv[2] = 0; shv[3] = 0; }  // far edges, ghosts
        __syncthreads();
        // Ring column of every row. A row publishes its cells into one column of the DP's shared-memory ring and
        // its successors read them with LDS.128; the 8 lanes of a quarter-warp are served in one wavefront only if
        // their cells lie in 8 different 16-byte bank groups. Bank group of (column c, distance d) = (c - d) mod 8
        // (the ring's slot stride is one cell more than a multiple of eight, mesh.cu). So a row asks for the column
        // whose bank group equals the LANE (mod 8) of the successor that will read it, preferring the successor with
        // the fewest predecessors: if every lane of a quarter-warp gets its wish the load is conflict free. The
        // column stays inside the row's own block of 8 threads (the stores of a warp keep covering whole lines),
        // wishes that collide inside a block are granted in thread order, the others take what is left.
        if (valid) {
            const uint32_t sg_ = nsigma[m], me = nthr[m];
            for (uint32_t e = pred_off[m]; e < pred_off[m + 1]; e++) {
                const uint32_t p = preds[e], d = sg_ - nsigma[p];
                if (p >= lo && d <= (uint32_t)DP_MAXD) atomicMin(&rwant[nthr[p]], (min(np, 15u) << 3) | ((me + d) & 7u));
                else {
                    const uint32_t i = atomicAdd(&shv[2], 1u);
                    if (i < FARLIST_CAP) far_e[i] = e; else shv[1] = 1;
                    pdesc2[e] = m;  // remember the consumer until the ghost is known
                }
            }
        }
        __syncthreads();
        if (tid < T / 8) {
            uint32_t used = 0, w[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                w[i] = rwant[tid * 8 + i];
                const uint32_t b = w[i] & 7u;
                if (w[i] != NONE && !((used >> b) & 1u)) { used |= 1u << b; w[i] = b; } else w[i] = NONE;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (w[i] == NONE) { w[i] = (uint32_t)__ffs((int)~used) - 1u; used |= 1u << w[i]; }
                rwant[tid * 8 + i] = tid * 8 + w[i];
            }
        }
        __syncthreads();
        if (tid < T) A.rcol[((uint64_t)ql * A.gcap + g) * T + tid] = (rcol_t)rwant[tid];
        if (valid) {
            const uint32_t sg_ = nsigma[m];
            for (uint32_t e = pred_off[m]; e < pred_off[m + 1]; e++) {
                const uint32_t p = preds[e], d = sg_ - nsigma[p];
                if (p >= lo && d <= (uint32_t)DP_MAXD) pdesc2[e] = (d << 16) | rwant[nthr[p]];
            }
        }
        __syncthreads();
        const uint32_t nfar = min(shv[2], FARLIST_CAP);
        if (tid == 0) {  // ghosts = distinct (source row, DP_MAXD-rank bucket of the consumer)
            uint32_t ng = 0;
            for (uint32_t i = 0; i < nfar; i++) {
                const uint32_t e = far_e[i], p = preds[e], mm = pdesc2[e];
                const uint32_t bk = (nsigma[mm] - sigma_lo) / DP_MAXD;
                uint32_t j = 0;
                while (j < ng && !(gh_p[j] == p && gh_b[j] == bk)) j++;
                if (j == ng) {
                    if (ng == DP_G - 2) { shv[1] = 1; j = 0; }   // the last two loader columns are the DP's constant columns
                    else { gh_p[ng] = p; gh_b[ng] = bk; ng++; }
                }
                far_gi[i] = j;
            }
            shv[3] = ng;
        }
        __syncthreads();
        const uint32_t ng = shv[3];
        auto ghost_sigma = [&](uint32_t j) -> int {  // column rank the ghost pretends to sit at
            int sgm = (int)sigma_lo + (int)(gh_b[j] * DP_MAXD) - 1;
            if (gh_p[j] >= lo) sgm = max(sgm, (int)nsigma[gh_p[j]] + GHOST_LEAD);  // source still being computed
            return sgm;
        };
        for (uint32_t i = tid; i < nfar; i += nt) {
            const uint32_t e = far_e[i], mm = pdesc2[e], j = far_gi[i];
            const uint32_t d = (uint32_t)((int)nsigma[mm] - ghost_sigma(j));
            pdesc2[e] = (d << 16) | (T + j);
        }
        if (tid < ng) {
            GhostInfo gi;
            gi.spillrow = (uint32_t)spillrow[gh_p[tid]];
            gi.soff = ghost_sigma(tid) - (int)sigma_lo;
            ghosts[(uint64_t)g * DP_G + tid] = gi;
        }
        // writer lanes: rows with a far successor (spill) or without successor (end-cell search needs the row min)
        const uint32_t wflag = valid && (spillrow[m] >= 0 || nflags[m] == 0) ? 1u : 0u;
        uint32_t nwr;
        const uint32_t wex = block_exscan(wflag, red, &nwr);
        if (wflag && wex < DP_G) writers[(uint64_t)g * DP_G + wex] = m;
        if (tid == 0) {
            if (nwr > DP_G) shv[1] = 1;
            groups[g].n_ghost = ng;
            groups[g].n_writer = min(nwr, (uint32_t)DP_G);
        }
        __syncthreads();
    }
    // The v2 kernel never materialises the reference's initial cell value 1000000 (mesh.cu): it relies on every cell having
    // a candidate below it, value(m, s) <= 1 + (column rank + s) * max(gap, gapext) by a chain of gap steps. Penalties that
    // could break that bound (or negative ones) send the query through the generic kernel, which keeps the initial value.
    const bool bound_ok = A.pen_min >= 0.f && 1.f + (float)(n_cols + Lq) * A.pen_max < 900000.f;
    const uint32_t mode = (shv[1] || A.force_generic || wide || !bound_ok) ? 1u : 2u;   // v2 holds in-degree <= 8 and u8 cells
    if (tid == 0) {
        uint64_t words_total = 0;
        for (uint32_t g = 0; g < n_groups; g++) {
            const uint32_t lo = nsigma[g * T];
            const uint32_t hi = nsigma[min(V, (g + 1) * T) - 1];
            groups[g].sigma_lo = lo;
            groups[g].depth = hi - lo + 1;
            groups[g].tb_off = words_total;
            if (mode == 2) {   // two positions per step, one word per thread and pair of steps (mesh.cu), behind one pad
                               // row: the DP kernel stores a pair's word one row back while it computes the next pair
                const uint32_t steps8 = (((Lq + 1u) >> 1) + (hi - lo) + 7u) & ~7u;
                groups[g].tb_off = words_total + T;
                words_total += (uint64_t)((steps8 >> 1) + 1u) * T;
            } else {
                const uint32_t steps = (Lq + (hi - lo) + 3) & ~3u;
                const uint32_t per_word = wide ? 2 : 4;
                words_total += (uint64_t)(steps / per_word) * T;
            }
        }
        const uint64_t spill_need = (uint64_t)n_spill * ((Lq + 1u) & ~1u);   // rows of an even number of positions (16-byte cells of the v2 kernel)
        if (words_total > A.tb_words || spill_need > A.spill_elems) { hdr->status = GS_LIMIT; return; }   // not even alone
        const uint64_t tb_off = atomicAdd(&A.cursors[0], (unsigned long long)words_total);
        const uint64_t sp_off = atomicAdd(&A.cursors[1], (unsigned long long)spill_need);
        hdr->V = V; hdr->E = E; hdr->n_cols = n_cols; hdr->n_groups = n_groups;
        hdr->n_last = n_last; hdr->n_spill = n_spill; hdr->max_indeg = max_indeg; hdr->wide = wide;
        hdr->maskset = shv[4];
        hdr->mode = mode;
        hdr->tb_off = tb_off; hdr->spill_off = sp_off;
        if (tb_off + words_total > A.tb_words || sp_off + spill_need > A.spill_elems) hdr->status = GS_ARENA_FULL;
        else atomicAdd(A.cells, (unsigned long long)V * Lq);
    }
}


int launch_prealign(Session* s, const sg_align_params& ap, uint32_t q0, uint32_t n) {
    Index* ix = s->ix;
    if (n == 0) { q0 = 0; n = s->nq; }
    const uint64_t fo = (uint64_t)q0 * s->fam_cap;
    uint32_t pairs = n * s->fam_cap;
    contains_kernel<<<(pairs + 3) / 4, 128, 0, s->stream>>>(s->d_qmasks, s->d_qoff + q0, n, ix->d_masks, ix->d_row_off,
                                                           s->d_fam_ids + fo, s->d_fam_n + q0, s->fam_cap,
                                                           s->d_contains + fo);
    partition_kernel<<<(n + 127) / 128, 128, 0, s->stream>>>(n, s->d_fam_ids + fo, s->d_fam_n + q0, s->fam_cap,
                                                            s->d_contains + fo, s->d_qoff + q0, ix->d_row_off,
                                                            ap.realign, s->d_afam + fo, s->d_afam_n + q0, s->d_copy_src + 2 * (uint64_t)q0,
                                                            s->d_hdr + q0);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 2;
    return SG_OK;
}

static uint64_t env_kb(const char* name, uint64_t dflt) {
    const char* v = getenv(name);
    return (v && *v) ? strtoull(v, nullptr, 10) : dflt;
}

int launch_graph(Session* s, Workspace* w, const sg_align_params& ap, uint32_t q0, uint32_t n) {
    Index* ix = s->ix;
    GraphArgs A;
    A.masks = ix->d_masks; A.cols = ix->d_cols; A.row_off = ix->d_row_off; A.W = ix->W;
    A.afam = s->d_afam; A.afam_n = s->d_afam_n; A.fam_cap = s->fam_cap;
    A.icap = s->icap; A.ncap = s->ncap; A.gcap = s->gcap; A.q0 = q0; A.itemcap = s->itemcap;
    A.hdr = s->d_hdr; A.tab = w->d_tab; A.tabli = w->d_tabli; A.colof = w->d_colof; A.colbase = w->d_colbase;
    A.item_node = w->d_item_node; A.slot = w->d_slot; A.ncol = w->d_ncol; A.nmask = w->d_nmask;
    A.ncount = w->d_ncount; A.nweight = w->d_nweight; A.nsigma = w->d_nsigma; A.slotbase = w->d_slotbase;
    A.cursor = w->d_cursor; A.pred_off = w->d_pred_off; A.preds = w->d_preds; A.pdesc = w->d_pdesc;
    A.spillrow = w->d_spillrow; A.nflags = w->d_nflags; A.lastnodes = w->d_lastnodes; A.groups = w->d_groups;
    A.order = w->d_order; A.rcol = w->d_rcol; A.nthr = w->d_nthr; A.pdesc2 = w->d_pdesc2; A.ghosts = w->d_ghosts; A.writers = w->d_writers;
    A.pen_max = fmaxf(ap.gap_penalty, ap.gap_ext_penalty); A.pen_min = fminf(ap.gap_penalty, ap.gap_ext_penalty);
    A.force_generic = s->force_generic || ap.insertion == 1 || ix->d_colw != nullptr;   // the aspace-aware transition and the weighted scheme live in the generic kernel
    A.nmaxins = w->d_nmaxins; A.forbid = ap.insertion == 1;
    A.cells = s->d_counters + 1; A.cursors = w->d_cursors; A.tb_words = s->tb_words; A.spill_elems = s->spill_elems;
    A.fs_weight = ap.fs_weight;
    const uint32_t words = (ix->W + 31) >> 5;
    const size_t base_smem = (size_t)(2 * words + s->fam_cap + 1 + 33 + 8 + 2 * FARLIST_CAP + 2 * DP_G + DP_T + 1 + 2 * s->fam_cap) * 4;
    // column table in shared memory: what the batch's families can need (columns <= items of the longest rows), capped so
    // that two CTAs stay resident per SM; a family that needs more takes the global-scratch path
    const uint64_t want = std::min<uint64_t>(s->ncap, (uint64_t)ix->max_row_len * 2) * s->fam_cap;
    A.stab_bytes = s->graph_generic ? 0u : (uint32_t)((std::min<uint64_t>((want + 1) / 2, env_kb("SG_STAB_KB", 40) * 1024) + 15u) & ~15ull);   // a nibble per entry
    size_t smem = base_smem + 16 + A.stab_bytes;
    if (smem > 200 * 1024) { A.stab_bytes = 0; smem = base_smem + 16; }   // very wide alignments: the bitmap alone fills the SM
    SG_CUDA(cudaFuncSetAttribute(graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    graph_kernel<<<n, GRAPH_BLOCK, smem, w->stream>>>(A);
    s->stats.kernel_launches += 1;
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

}  // namespace sg
