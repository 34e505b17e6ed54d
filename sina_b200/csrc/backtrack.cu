// Backtrack + gap placement on the device, one warp per query.
// Replaces backtrack() (reference src/mesh.h:534-739) and cseq::fix_duplicate_positions
// (src/cseq.cpp:456-594), decoding the packed traceback written by mesh.cu instead of the reference's
// 28-byte cells:
//   value_(midx,sidx) of a cell  = NONE: (0,0) | MATCH via pred i: (pred_i, s-1) | INS: (m, gaps_idx(m,s))
//                                  | DEL via pred i: ob(pred_i, s) ? (pred_i, s) : (gapm_idx(pred_i, s), s)
//   gaps_idx(m,s) = ins-open(m,s) ? s-1 : gaps_idx(m,s-1), gaps_idx(m,0) = 0            (mesh.h:340-349)
//   gapm_idx(x,s) = no preds ? 0 : ob(lastpred(x), s) ? lastpred(x) : gapm_idx(lastpred(x), s)  (mesh.h:315-323)
//   ob(x,s) = "a deletion leaving (x,s) opens the gap", stored in (x,s)'s own traceback cell (common.cuh)
//
// The walk is a pointer chase through HBM (one traceback byte per step, 3-4 MB per query), so the kernel is
// organised around the length of the dependent-load chain:
//   * prologue: the warp packs everything the walk needs to know about a node into one 32-byte record
//     (traceback row base, step offset, predecessor-slot shift, in-degree, column, first four predecessors);
//   * walk: all lanes run the same chase. As soon as a node's record is there the records of its (<= 4 inline)
//     predecessors are requested, so that when the node's traceback cell arrives and names the predecessor, that
//     record is already in registers and the next cell can be requested at once: one memory latency per step.
//     The walk only records which node every query position landed on;
//   * epilogue (lane-parallel): columns of the overhangs and the aligned part in closed form, cseq::append's
//     running width as a prefix maximum, sum_weight added in the reference's order, reverse + setWidth, and the
//     serial gap placement only for queries that actually have bases sharing a column.
#include "common.cuh"

namespace sg {

struct __align__(16) NodeRec {
    uint32_t tbbase;    // generic kernel: index of the row's first cell pair inside the query's traceback block (u16 or u32
                        // units); v2 kernel: byte offset of the row's first word (4 cells)
    uint32_t meta;      // [15:0] sigma - sigma_lo(group)  [23:16] predecessor-slot shift  [31:24] in-degree
    uint32_t ncol;
    uint32_t pred_off;
    uint32_t pred[4];   // predecessors 0..3 (ascending id)
    uint32_t ptb[3];    // tbbase of predecessors 0..2: their traceback cells can be requested as soon as THIS record is known,
    uint32_t psoff;     // [9:0], [19:10], [29:20] their step offsets           without waiting for their own records
};
static_assert(sizeof(NodeRec) == 48, "NodeRec is read as three 16-byte loads");
static_assert(DP_T <= 1024, "step offsets inside a group are packed into 10 bits");

struct BtArgs {
    uint32_t nq, W, q0;  // nq queries of the chunk starting at q0
    const uint8_t* qmasks; const uint64_t* qoff;
    GraphHdr* hdr; const GroupInfo* groups; uint32_t gcap, icap; uint32_t* remaining;
    const uint32_t* ncol; const float* nweight; const uint32_t* nsigma; const uint32_t* pred_off;
    const uint32_t* preds; const uint32_t* lastnodes; const uint32_t* afam_n; const uint16_t* nthr; const uint8_t* nshift;
    const uint32_t* tb; const float* lastcol; const float* rowmin; const uint32_t* rowarg;
    const uint32_t* copy_src; const uint8_t* masks; const uint32_t* cols; const uint64_t* row_off;
    NodeRec* rec;
    uint32_t* out_cols; uint8_t* out_masks; sg_align_result* results;
    float ms; int overhang, lowercase;
    const float* colw; uint32_t ncolw;   // positional weights (scoring_scheme_weighted, generic DP kernel), null = none
};

constexpr int BT_WARPS = 2;   // queries per CTA
constexpr uint32_t FULL = 0xffffffffu;

// cseq_base::fix_duplicate_positions (src/cseq.cpp:456-594); returns 1 for its runtime_error
__device__ int fix_duplicate_positions(uint32_t* pos, uint8_t* masks, uint32_t n, uint32_t width, int lowercase) {
    uint32_t last = 0;
    for (uint32_t curr = 0; curr < n; ++curr) {
        if (pos[last] == pos[curr]) {
            if (curr + 1 != n) continue;
            ++curr;
        }
        uint32_t num_inserts = curr - last - 1;
        if (num_inserts == 0) { last = curr; continue; }
        uint32_t range_begin = pos[last] + 1;
        uint32_t range_end = (curr == n) ? width : pos[curr];
        ++last;
        --curr;
        if (range_end - range_begin < num_inserts) {
            while (range_end - range_begin < num_inserts) {
                int next_left_gap, next_right_gap;
                uint32_t left = last, right = curr;
                if (left == 0) {
                    next_left_gap = (range_begin > 0) ? (int)(range_begin - 1) : -1;
                } else if (pos[left - 1] + 1 < range_begin) {
                    next_left_gap = (int)(range_begin - 1);
                } else {
                    --left;
                    while (left != 0 && pos[left - 1] + 1 >= pos[left]) --left;
                    next_left_gap = (int)(pos[left] - 1);
                }
                if (right + 1 == n) {
                    next_right_gap = (range_end < width) ? (int)range_end : -1;
                } else if (pos[right + 1] > range_end) {
                    next_right_gap = (int)range_end;
                } else {
                    ++right;
                    while (right + 1 != n && pos[right] + 1 >= pos[right + 1]) ++right;
                    next_right_gap = (int)(pos[right] + 1);
                }
                if (next_right_gap == -1 ||
                    (next_left_gap != -1 &&
                     range_begin - (uint32_t)next_left_gap <= (uint32_t)next_right_gap - (range_end - 1))) {
                    if (next_left_gap == -1) return 1;
                    num_inserts += last - left;
                    range_begin = (uint32_t)next_left_gap;
                    last = left;
                } else {
                    num_inserts += right - curr;
                    range_end = (uint32_t)next_right_gap + 1;
                    curr = right;
                }
            }
        } else {
            range_begin = range_end - num_inserts;
        }
        ++curr;
        for (; last != curr; ++last) {
            pos[last] = range_begin++;
            if (lowercase) masks[last] |= 16;
        }
        last = curr;
    }
    return 0;
}

// strict-'<' argmin in ascending index order (the reference's scan): (value, first index reaching it) over the
// lanes' partial results
__device__ __forceinline__ void warp_argmin(float& v, uint32_t& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float v2 = __shfl_xor_sync(FULL, v, o);
        const uint32_t i2 = __shfl_xor_sync(FULL, i, o);
        if (v2 < v || (v2 == v && i2 < i)) { v = v2; i = i2; }
    }
}

__global__ void __launch_bounds__(32 * BT_WARPS) backtrack_kernel(BtArgs A) {
    const uint32_t ql = blockIdx.x * BT_WARPS + warp_id();
    if (ql >= A.nq) return;
    const uint32_t lane = lane_id();
    const uint32_t q = A.q0 + ql;
    const GraphHdr h = A.hdr[q];
    if (h.status == GS_DONE) return;  // finished in an earlier pass
    if (h.status == GS_ARENA_FULL) { if (lane == 0) atomicAdd(A.remaining, 1u); return; }  // host resets the arenas and re-runs
    sg_align_result r = {};
    const uint64_t qo = A.qoff[q];
    const uint32_t L = h.qlen;
    const uint8_t* qm = A.qmasks + qo;
    uint32_t* ocols = A.out_cols + qo;
    uint8_t* omasks = A.out_masks + qo;
    r.fam_used = A.afam_n[q];
    if (h.status == SG_Q_COPIED) {  // setAlignedBases from the containing relative (src/align.cpp:349-388)
        const uint32_t id = A.copy_src[2 * q], at = A.copy_src[2 * q + 1];
        const uint64_t ro = A.row_off[id] + at;
        for (uint32_t i = lane; i < L; i += 32) { ocols[i] = A.cols[ro + i]; omasks[i] = A.masks[ro + i]; }
        r.status = SG_Q_COPIED; r.score = 1.f; r.qual = 100; r.n_out = L;
        if (lane == 0) { A.results[q] = r; A.hdr[q].status = GS_DONE; }
        return;
    }
    if (h.status != GS_OK) {
        r.status = h.status == GS_LIMIT ? (int32_t)SG_Q_LIMIT : (int32_t)h.status;
        if (lane == 0) { A.results[q] = r; A.hdr[q].status = GS_DONE; }
        return;
    }

    const uint64_t io = (uint64_t)ql * A.icap;
    const uint32_t* pred_off = A.pred_off + (uint64_t)ql * (A.icap + 1);
    const uint32_t* preds = A.preds + io;
    const float* nweight = A.nweight + io;
    const uint32_t* tbq = A.tb + h.tb_off;
    const uint16_t* tbq16 = reinterpret_cast<const uint16_t*>(tbq);
    const uint8_t* tbq8 = reinterpret_cast<const uint8_t*>(tbq);
    const bool wide = h.wide != 0;
    const bool v2 = h.mode == 2;   // cells written by the v2 DP kernel: two query positions per step (common.cuh)
    const bool weighted = A.colw != nullptr;
    const uint32_t V = h.V, W = A.W;
    const uint32_t T = DP_T;
    NodeRec* rec = A.rec + io;

    // ---- prologue: node records
    {
        const uint32_t* nsigma = A.nsigma + io;
        const uint32_t* ncol = A.ncol + io;
        const GroupInfo* groups = A.groups + (uint64_t)ql * A.gcap;
        const uint16_t* nthr = A.nthr + io;
        const uint8_t* nshift = A.nshift + io;
        const bool sorted = h.mode >= 2;  // v2 kernel: rows sorted inside the group, predecessor slots right-aligned
        auto tb_of = [&](uint32_t x, uint32_t& soff) -> uint32_t {   // (tbbase, step offset) of node x
            const uint32_t g = x / T, tid = sorted ? (uint32_t)nthr[x] : x - g * T;
            const GroupInfo gi = groups[g];
            soff = nsigma[x] - gi.sigma_lo;
            return v2 ? (uint32_t)(4 * gi.tb_off) + 4u * tid : (uint32_t)(wide ? gi.tb_off : 2 * gi.tb_off) + tid;
        };
        for (uint32_t m = lane; m < V; m += 32) {
            const uint32_t po = pred_off[m], np = pred_off[m + 1] - po;
            uint4 a, b, c = make_uint4(0, 0, 0, 0);
            uint32_t soff;
            a.x = tb_of(m, soff);
            a.y = soff | ((uint32_t)nshift[m] << 16) | (min(np, 255u) << 24);
            a.z = ncol[m];
            a.w = po;
            b.x = np > 0 ? preds[po] : 0u;
            b.y = np > 1 ? preds[po + 1] : 0u;
            b.z = np > 2 ? preds[po + 2] : 0u;
            b.w = np > 3 ? preds[po + 3] : 0u;
            if (np > 0) { c.x = tb_of(b.x, soff); c.w |= soff; }
            if (np > 1) { c.y = tb_of(b.y, soff); c.w |= soff << 10; }
            if (np > 2) { c.z = tb_of(b.z, soff); c.w |= soff << 20; }
            uint4* dst = reinterpret_cast<uint4*>(rec + m);
            __stcg(dst, a);
            __stcg(dst + 1, b);
            __stcg(dst + 2, c);
        }
    }
    __syncwarp();

    auto ldrec = [&](uint32_t x, uint4& a, uint4& b) {
        const uint4* p = reinterpret_cast<const uint4*>(rec + x);
        a = __ldcg(p);
        b = __ldcg(p + 1);
    };
    auto ldrec3 = [&](uint32_t x, uint4& a, uint4& b, uint4& c) {
        const uint4* p = reinterpret_cast<const uint4*>(rec + x);
        a = __ldcg(p);
        b = __ldcg(p + 1);
        c = __ldcg(p + 2);
    };
    // traceback cell: the load (cell_raw) and its decoding (cell_dec) are separate so that a speculative load can stay
    // in flight; decoded = src | slot<<8 | ob<<2 (the wide layout; ob = a deletion leaving the cell opens, common.cuh)
    auto cell_raw = [&](const uint4& a, uint32_t s) -> uint32_t {
        if (v2) {
            const uint32_t t = (s >> 1) + (a.y & 0xffffu);
            return (uint32_t)__ldcg(&tbq8[a.x + (t >> 1) * (T * 4u) + 2u * (t & 1u) + (s & 1u)]);
        }
        const uint32_t t = s + (a.y & 0xffffu);
        const uint32_t idx = a.x + (t >> 1) * T;
        return wide ? __ldcg(&tbq[idx]) : (uint32_t)__ldcg(&tbq16[idx]);
    };
    auto cell_dec = [&](const uint4& a, uint32_t s, uint32_t raw) -> uint32_t {
        const uint32_t sh = (a.y >> 16) & 0xffu, np = a.y >> 24;
        if (v2) {
            const uint32_t c = raw & 0xffu;
            if (sh & TBR_FLAG) {
                // raw cell (common.cuh): the first flag in the order insertion, deletion slots, match slots names the
                // source; no flag = the last match slot; a row without predecessor has no deletion / match source
                const uint32_t mt = (c >> 4) & 3u, dl = c & 7u;
                uint32_t out = 0;
                if (c & TBR_INS) out = TB_SRC_INS;
                else if (np == 0) out = TB_SRC_NONE;
                else if (dl) out = TB_SRC_DEL | (((uint32_t)__ffs((int)dl) - 1u) << 8);
                else if (mt) out = TB_SRC_MATCH | (((uint32_t)__ffs((int)mt) - 1u) << 8);
                else out = TB_SRC_MATCH | (((sh & 0x7fu) + np - 1u) << 8);
                return out | ((c >> 7) << 2);
            }
            // index cell (common.cuh): 0 insertion, 1..slots deletion, slots+1.. match
            const uint32_t idx = c & 31u, slots = (sh & 0x7fu) + np;
            uint32_t out = 0;
            if (idx == 0) out = TB_SRC_INS;
            else if (np == 0) out = TB_SRC_NONE;
            else if (idx <= slots) out = TB_SRC_DEL | ((idx - 1u) << 8);
            else out = TB_SRC_MATCH | ((idx - 1u - slots) << 8);
            return out | ((c >> 7) << 2);
        }
        // generic kernel: bit 2 = ob (weighted scheme: the LAST predecessor's deletion opened), bit 3 (weighted scheme
        // only) = the chosen deletion opened
        const uint32_t t = s + (a.y & 0xffffu);
        if (wide) return (raw >> (16 * (t & 1))) & 0xffffu;
        const uint32_t c = (raw >> (8 * (t & 1))) & 0xffu;
        return (c & 3u) | (((c >> 2) & 7u) << 8) | (((c >> 5) & 1u) << 2) | (((c >> 6) & 1u) << 3);
    };
    auto cell = [&](const uint4& a, uint32_t s) -> uint32_t { return cell_dec(a, s, cell_raw(a, s)); };
    auto np_of = [](const uint4& a) -> uint32_t { return a.y >> 24; };
    auto pred_of = [&](const uint4& a, const uint4& b, uint32_t ord) -> uint32_t {
        if (ord == 0) return b.x;
        if (ord == 1) return b.y;
        if (ord == 2) return b.z;
        if (ord == 3) return b.w;
        return __ldcg(&preds[a.w + ord]);
    };
    // gaps_idx(m, s): first query position of the insertion run ending at (m, s). The insertion at (m, s') extends
    // iff gaps_val == value at (m, s'-1) (mesh.h:340-349), i.e. iff the source of (m, s'-1) is an insertion (the
    // insertion candidate wins ties against everything evaluated before it and loses only to a strictly smaller
    // match), so the run starts at the first cell left of s whose source is something else. Lanes look at 32 cells
    // of the row at a time.
    auto gaps_idx = [&](const uint4& a, uint32_t s) -> uint32_t {
        uint32_t top = s;                       // cells top-1, top-2, ...
        while (top > 0) {
            const bool in = lane < top;
            const uint32_t cc = in ? cell(a, top - 1 - lane) : TB_SRC_INS;
            const uint32_t hit = __ballot_sync(FULL, in && (cc & 3u) != TB_SRC_INS);
            if (hit) return top - 1u - ((uint32_t)__ffs((int)hit) - 1u);
            top = top > 32 ? top - 32 : 0;
        }
        return 0;
    };
    // (value_midx, value_sidx) of cell (m,s) = c
    auto follow = [&](const uint4& a, const uint4& b, uint32_t m, uint32_t s, uint32_t c, uint32_t& nm, uint32_t& ns) {
        const uint32_t src = c & 3u, sl = c >> 8, sh = (a.y >> 16) & (wide ? 0xffu : 0x7fu);
        const uint32_t ord = sl > sh ? sl - sh : 0u;  // v2 slots are right-aligned, leading slots repeat ordinal 0
        if (src == TB_SRC_NONE) { nm = 0; ns = 0; }
        else if (src == TB_SRC_MATCH) { nm = pred_of(a, b, ord); ns = s - 1; }
        else if (src == TB_SRC_INS) { nm = m; ns = gaps_idx(a, s); }
        else if (weighted) {
            // weighted scheme: whether a deletion opens depends on the TARGET node's column weight, so the facts are in
            // the target's cells: the chosen deletion via p opened (bit 3 of (m,s)) -> (p, s); else gapm_idx(p, s) with
            // gapm_idx(x, s) = "last predecessor's deletion opened" (bit 2 of (x,s)) ? lastpred(x) : gapm_idx(lastpred(x), s),
            // 0 for a row without predecessor (mesh.h:305-330)
            uint32_t x = pred_of(a, b, ord);
            if (!(c & 8u)) {
                for (;;) {
                    uint4 xa, xb;
                    ldrec(x, xa, xb);
                    const uint32_t np = np_of(xa);
                    if (np == 0) { x = 0; break; }
                    const bool opened = cell(xa, s) & 4u;
                    x = pred_of(xa, xb, np - 1);
                    if (opened) break;
                }
            }
            nm = x;
            ns = s;
        } else {
            // deletion via predecessor p: (p, s) if it opened there, else gapm_idx(p, s): down the chain of last
            // predecessors to the first cell whose deletion opens; a row without predecessor ends it at node 0
            // (init_edge, mesh.h:294-297)
            uint32_t x = pred_of(a, b, ord);
            for (;;) {
                uint4 xa, xb;
                ldrec(x, xa, xb);
                if (cell(xa, s) & 4u) break;
                const uint32_t np = np_of(xa);
                if (np == 0) { x = 0; break; }
                x = pred_of(xa, xb, np - 1);
            }
            nm = x;
            ns = s;
        }
    };

    // ---- starting point (mesh.h:567-592): strict-'<' scans in id order, starting from the first last-node
    const uint32_t send = L - 1;
    const float* lastcol = A.lastcol + io;
    const uint32_t* lastnodes = A.lastnodes + io;
    uint32_t m = lastnodes[0];
    float best = lastcol[m];
    {
        float bv = __int_as_float(0x7f800000);
        uint32_t bi = 0xffffffffu;
        for (uint32_t t = lane; t < V; t += 32) { const float v = lastcol[t]; if (v < bv) { bv = v; bi = t; } }
        warp_argmin(bv, bi);
        if (bv < best) { best = bv; m = bi; }
    }
    uint32_t s = send;
    {
        float bv = __int_as_float(0x7f800000);
        uint32_t bi = 0xffffffffu;
        for (uint32_t i = lane; i < h.n_last; i += 32) {
            const float v = A.rowmin[io + lastnodes[i]];
            if (v < bv) { bv = v; bi = i; }
        }
        warp_argmin(bv, bi);
        if (bv < best) { best = bv; m = lastnodes[bi]; s = A.rowarg[io + m]; }
    }
    r.end_m = m; r.end_s = s;
    const uint32_t m_end = m, s_end = s;

    // ---- walk back (mesh.h:642-685): ocols[s] = node the query position s is aligned to
    uint4 ra, rb, rc;
    ldrec3(m, ra, rb, rc);
    uint32_t c = cell(ra, s);
    if (lane == 0) ocols[s] = m;
    auto bcast4 = [&](const uint4& v, uint32_t from) -> uint4 {
        return make_uint4(__shfl_sync(FULL, v.x, from), __shfl_sync(FULL, v.y, from), __shfl_sync(FULL, v.z, from), __shfl_sync(FULL, v.w, from));
    };
    while (s != 0 && np_of(ra) != 0) {
        const uint32_t np = np_of(ra);
        // Speculation: the most common step is a match through one of the first predecessors. Lane k < 4 fetches
        // predecessor k: its record, and its cell (p_k, s-1) -- for k < 3 straight from THIS node's record, which carries
        // their traceback bases, so that record and cell travel together: one memory latency per step. The winner's
        // record and cell are then broadcast from its lane (a first version kept all four records in every lane and spent
        // ~450 dependent instructions per step selecting among them: the walk was bound by issue latency, not memory).
        uint4 qa = ra, qb = rb, qc = rc;
        uint32_t spec_raw = 0;
        if (lane < min(np, 4u)) {
            const uint32_t pk = lane == 0 ? rb.x : (lane == 1 ? rb.y : (lane == 2 ? rb.z : rb.w));
            if (lane < 3) {
                uint4 fake = ra;
                fake.x = lane == 0 ? rc.x : (lane == 1 ? rc.y : rc.z);
                fake.y = (rc.w >> (10 * lane)) & 0x3ffu;
                spec_raw = cell_raw(fake, s - 1);
            }
            ldrec3(pk, qa, qb, qc);
            if (lane == 3) spec_raw = cell_raw(qa, s - 1);
        }
        uint32_t ord = 0xffffffffu;   // predecessor ordinal of a match step
        if ((c & 3u) == TB_SRC_MATCH) {
            const uint32_t sl = c >> 8, sh = (ra.y >> 16) & (wide ? 0xffu : 0x7fu);
            ord = sl > sh ? sl - sh : 0u;
        }
        uint32_t snew, c2 = 0;
        if (ord < 4u) {
            m = ord == 0 ? rb.x : (ord == 1 ? rb.y : (ord == 2 ? rb.z : rb.w));
            snew = s - 1;
            ra = bcast4(qa, ord); rb = bcast4(qb, ord); rc = bcast4(qc, ord);
            const uint32_t raw = __shfl_sync(FULL, spec_raw, ord);
            if (snew != 0) c2 = cell_dec(ra, snew, raw);
        } else {
            uint32_t nm;
            follow(ra, rb, m, s, c, nm, snew);
            if (nm != m) ldrec3(nm, ra, rb, rc);
            m = nm;
            if (snew != 0) c2 = cell(ra, snew);
        }
        if (snew != 0 && (c2 & 3u) == TB_SRC_DEL) {   // landing on a cell reached by deletion (its value_sidx == snew): skip it (mesh.h:653-655)
            uint32_t m2, s2;
            follow(ra, rb, m, snew, c2, m2, s2);
            m = m2;
            ldrec3(m, ra, rb, rc);
            c2 = cell(ra, snew);
        }
        if (snew + 1 == s) { if (lane == 0) ocols[snew] = m; }   // one base per step, almost always
        else for (int ss = (int)s - 1 - (int)lane; ss >= (int)snew; ss -= 32) ocols[ss] = m;
        s = snew;
        c = c2;
    }
    const uint32_t m_fin = m, s_fin = s;
    __syncwarp();

    // ---- epilogue. Elements are appended for s descending: right overhang (mesh.h:594-615), aligned part,
    // left overhang (mesh.h:690-721); cseq::append (src/cseq.cpp:79-95) places element j at the running maximum of
    // the requested positions, i.e. a prefix maximum.
    const bool keep_case = A.lowercase == 1;
    const bool lc_unaligned = A.lowercase == 2;
    const int cutoff_tail = (int)(send - s_end);
    const bool right_oh = cutoff_tail != 0 && A.overhang != 1;
    const bool left_oh = s_fin != 0 && A.overhang != 1;
    const uint32_t top_s = right_oh ? send : s_end;
    const uint32_t bottom_s = left_oh ? 0u : s_fin;
    const uint32_t n = top_s - bottom_s + 1;
    const int rbase = (A.overhang == 0) ? (int)W - 1 - (int)A.ncol[io + m_end] - cutoff_tail : 0;
    const uint32_t lpos = W - 1 - A.ncol[io + m_fin];   // `pos` when the walk ended
    // sum_weight: the aligned elements' match scores added in append order (mesh.h:640,683)
    float sum_weight = 0.f;
    for (uint32_t j0 = 0; j0 <= s_end - s_fin; j0 += 32) {
        const uint32_t j = j0 + lane;
        const bool in = j <= s_end - s_fin;
        float pw = 0.f;
        if (in) {   // match score of a forced match (mesh.h:631-638,683); weighted scheme: (match * weights[col]) * node weight
            const uint32_t node = ocols[s_end - j];
            pw = weighted ? __fmul_rn(__fmul_rn(A.ms, A.colw[min(A.ncol[io + node], A.ncolw - 1u)]), nweight[node])
                          : __fmul_rn(A.ms, nweight[node]);
        }
        const uint32_t cnt = min(32u, s_end - s_fin + 1 - j0);
        for (uint32_t l = 0; l < cnt; l++) sum_weight = __fadd_rn(sum_weight, __shfl_sync(FULL, pw, l));
    }
    uint32_t width = 0;
    for (uint32_t j0 = 0; j0 < n; j0 += 32) {
        const uint32_t j = j0 + lane;
        const bool in = j < n;
        const uint32_t sx = top_s - (in ? j : 0u);
        uint32_t p = 0;
        bool unal = false;
        if (in) {
            if (sx > s_end) {            // right overhang, element i = send - sx: position max(rbase + i, 0)
                const int pp = rbase + (int)(send - sx);
                p = (uint32_t)(pp > 0 ? pp : 0);
                unal = true;
            } else if (sx >= s_fin) {
                p = W - 1 - A.ncol[io + ocols[sx]];
            } else if (A.overhang == 0) {  // attach: ++pos per base, clamped to the last column
                const uint32_t pp = lpos + (s_fin - sx);
                p = pp < W - 1 ? pp : W - 1;
                unal = true;
            } else {                       // edge
                p = W - sx - 1;
                unal = true;
            }
        }
        uint32_t x = in ? p : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(FULL, x, o);
            if (lane >= (uint32_t)o) x = max(x, y);
        }
        x = max(x, width);
        width = __shfl_sync(FULL, x, 31);
        if (in) {
            // setWidth + reverse (mesh.h:723-724; cseq.cpp:283-289): column W-1-position, ascending s
            ocols[sx] = W - 1 - x;
            uint8_t b = keep_case ? qm[sx] : (uint8_t)(qm[sx] & 15u);
            if (unal && lc_unaligned) b |= 16;
            omasks[sx] = b;
        }
    }
    __syncwarp();
    if (bottom_s != 0) {  // --overhang remove dropped the head: the output starts at its first base
        for (uint32_t i0 = 0; i0 < n; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint32_t cv = 0; uint8_t mv = 0;
            if (i < n) { cv = ocols[bottom_s + i]; mv = omasks[bottom_s + i]; }
            __syncwarp();
            if (i < n) { ocols[i] = cv; omasks[i] = mv; }
            __syncwarp();
        }
    }
    // bases sharing a column (insertions) need the serial gap placement
    uint32_t dup = 0;
    for (uint32_t i0 = 0; i0 + 1 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        dup |= (i + 1 < n && ocols[i] == ocols[i + 1]) ? 1u : 0u;
    }
    dup = __any_sync(FULL, dup);
    int nospace = 0;
    if (dup) {
        if (lane == 0) nospace = fix_duplicate_positions(ocols, omasks, n, W, lc_unaligned);
        nospace = __shfl_sync(FULL, nospace, 0);
    }
    if (s_fin != 0) r.head = (int)s_fin;
    r.tail = cutoff_tail;
    r.n_out = n;
    r.raw = best; r.sum_weight = sum_weight;
    r.score = __fdiv_rn(best, sum_weight);
    const float q100 = __fmul_rn(100.f, r.score);  // src/align.cpp:509
    r.qual = (int)(q100 < 0.f ? 0.f : (q100 > 100.f ? 100.f : q100));
    r.n_nodes = V;
    r.status = nospace ? SG_Q_NOSPACE : SG_Q_ALIGNED;
    if (lane == 0) { A.results[q] = r; A.hdr[q].status = GS_DONE; }
}

int launch_backtrack(Session* s, Workspace* w, const sg_align_params& ap, uint32_t q0, uint32_t n) {
    Index* ix = s->ix;
    BtArgs A;
    A.nq = n; A.q0 = q0; A.W = ix->W; A.qmasks = s->d_qmasks; A.qoff = s->d_qoff; A.hdr = s->d_hdr; A.groups = w->d_groups;
    A.gcap = s->gcap; A.icap = s->icap; A.remaining = w->d_remaining; A.ncol = w->d_ncol; A.nweight = w->d_nweight; A.nsigma = w->d_nsigma;
    A.pred_off = w->d_pred_off; A.preds = w->d_preds; A.lastnodes = w->d_lastnodes; A.afam_n = s->d_afam_n;
    A.nthr = w->d_nthr; A.nshift = w->d_nshift;
    A.tb = w->d_tb; A.lastcol = w->d_lastcol; A.rowmin = w->d_rowmin; A.rowarg = w->d_rowarg;
    A.copy_src = s->d_copy_src; A.masks = ix->d_masks; A.cols = ix->d_cols; A.row_off = ix->d_row_off;
    A.rec = reinterpret_cast<NodeRec*>(w->d_rec);
    A.out_cols = s->d_out_cols; A.out_masks = s->d_out_masks; A.results = s->d_results;
    A.ms = -ap.match_score; A.overhang = ap.overhang; A.lowercase = ap.lowercase;
    A.colw = ix->d_colw; A.ncolw = ix->W;
    backtrack_kernel<<<(n + BT_WARPS - 1) / BT_WARPS, 32 * BT_WARPS, 0, w->bt_stream ? w->bt_stream : w->stream>>>(A);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 1;
    return SG_OK;
}

}  // namespace sg
