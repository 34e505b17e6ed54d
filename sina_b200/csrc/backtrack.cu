// Backtrack + gap placement on the device, one thread per query.
// Replaces backtrack() (reference src/mesh.h:534-739) and cseq::fix_duplicate_positions
// (src/cseq.cpp:456-594), decoding the packed traceback written by mesh.cu instead of the reference's
// 28-byte cells:
//   value_(midx,sidx) of a cell  = NONE: (0,0) | MATCH via pred i: (pred_i, s-1) | INS: (m, gaps_idx(m,s))
//                                  | DEL via pred i: opened ? (pred_i, s) : (gapm_idx(pred_i, s), s)
//   gaps_idx(m,s) = ins-open(m,s) ? s-1 : gaps_idx(m,s-1), gaps_idx(m,0) = 0            (mesh.h:340-349)
//   gapm_idx(x,s) = no preds ? 0 : last-open(x,s) ? lastpred(x) : gapm_idx(lastpred(x), s)  (mesh.h:315-323)
#include "common.cuh"

namespace sg {

struct BtArgs {
    uint32_t nq, W, q0;  // nq queries of the chunk starting at q0
    const uint8_t* qmasks; const uint64_t* qoff;
    GraphHdr* hdr; const GroupInfo* groups; uint32_t gcap, icap; uint32_t* remaining;
    const uint32_t* ncol; const float* nweight; const uint32_t* nsigma; const uint32_t* pred_off;
    const uint32_t* preds; const uint32_t* lastnodes; const uint32_t* afam_n; const uint16_t* nthr; const uint8_t* nshift;
    const uint32_t* tb; const float* lastcol; const float* rowmin; const uint32_t* rowarg;
    const uint32_t* copy_src; const uint8_t* masks; const uint32_t* cols; const uint64_t* row_off;
    uint32_t* out_cols; uint8_t* out_masks; sg_align_result* results;
    float ms; int overhang, lowercase;
};

struct Out {  // cseq under construction: append() semantics of src/cseq.cpp:79-95
    uint32_t* pos; uint8_t* mask; uint32_t n, width;
    __device__ void append(uint32_t p, uint8_t b) {
        if (p >= width) { pos[n] = p; mask[n] = b; n++; width = p; }
        else { pos[n] = width; mask[n] = b; n++; }
    }
};

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

__global__ void __launch_bounds__(64) backtrack_kernel(BtArgs A) {
    const uint32_t ql = blockIdx.x * blockDim.x + threadIdx.x;
    if (ql >= A.nq) return;
    const uint32_t q = A.q0 + ql;
    const GraphHdr h = A.hdr[q];
    if (h.status == GS_DONE) return;  // finished in an earlier pass
    if (h.status == GS_ARENA_FULL) { atomicAdd(A.remaining, 1u); return; }  // host resets the arenas and re-runs
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
        for (uint32_t i = 0; i < L; i++) { ocols[i] = A.cols[ro + i]; omasks[i] = A.masks[ro + i]; }
        r.status = SG_Q_COPIED; r.score = 1.f; r.qual = 100; r.n_out = L;
        A.results[q] = r;
        A.hdr[q].status = GS_DONE;
        return;
    }
    if (h.status != GS_OK) { r.status = (int32_t)h.status; A.results[q] = r; A.hdr[q].status = GS_DONE; return; }

    const uint64_t io = (uint64_t)ql * A.icap;
    const uint32_t* pred_off = A.pred_off + (uint64_t)ql * (A.icap + 1);
    const uint32_t* preds = A.preds + io;
    const uint32_t* nsigma = A.nsigma + io;
    const uint32_t* ncol = A.ncol + io;
    const float* nweight = A.nweight + io;
    const GroupInfo* groups = A.groups + (uint64_t)ql * A.gcap;
    const uint32_t* tbq = A.tb + h.tb_off;
    const bool wide = h.wide != 0;
    const uint32_t V = h.V, W = A.W;
    const uint32_t T = DP_T;
    const uint16_t* nthr = A.nthr + io;
    const uint8_t* nshift = A.nshift + io;
    const bool sorted = h.mode >= 2;  // v2 kernel: rows sorted inside the group, predecessor slots right-aligned

    // decoded traceback cell: src | ord<<8 | chosen_open<<2 | last_open<<3 | ins_open<<4 (the wide layout)
    auto cell = [&](uint32_t m, uint32_t s) -> uint32_t {
        const uint32_t g = m / T, tid = sorted ? (uint32_t)nthr[m] : m - g * T;
        const GroupInfo gi = groups[g];
        const uint32_t t = s + nsigma[m] - gi.sigma_lo;
        if (wide) {
            const uint32_t w = tbq[gi.tb_off + (uint64_t)(t >> 1) * T + tid];
            return (w >> (16 * (t & 1))) & 0xffffu;
        }
        const uint16_t* tb16 = reinterpret_cast<const uint16_t*>(tbq + gi.tb_off);
        const uint32_t c = ((uint32_t)tb16[(uint64_t)(t >> 1) * T + tid] >> (8 * (t & 1))) & 0xffu;
        return (c & 3u) | (((c >> 2) & 7u) << 8) | (((c >> 5) & 1u) << 2) | (((c >> 6) & 1u) << 3) | (((c >> 7) & 1u) << 4);
    };
    auto gaps_idx = [&](uint32_t m, uint32_t s) -> uint32_t {
        for (uint32_t cur = s; cur > 0; cur--) if (cell(m, cur) & 16u) return cur - 1;
        return 0;
    };
    auto gapm_idx = [&](uint32_t x, uint32_t s) -> uint32_t {
        for (;;) {
            if (pred_off[x + 1] == pred_off[x]) return 0;
            const uint32_t lp = preds[pred_off[x + 1] - 1];
            if (cell(x, s) & 8u) return lp;
            x = lp;
        }
    };
    // (value_midx, value_sidx) of cell (m,s)
    auto follow = [&](uint32_t m, uint32_t s, uint32_t c, uint32_t& nm, uint32_t& ns) {
        const uint32_t src = c & 3u, sl = c >> 8, sh = nshift[m];
        const uint32_t ord = sl > sh ? sl - sh : 0u;  // v2 slots are right-aligned, leading slots repeat ordinal 0
        if (src == TB_SRC_NONE) { nm = 0; ns = 0; }
        else if (src == TB_SRC_MATCH) { nm = preds[pred_off[m] + ord]; ns = s - 1; }
        else if (src == TB_SRC_INS) { nm = m; ns = gaps_idx(m, s); }
        else {
            const uint32_t p = preds[pred_off[m] + ord];
            // deletion opened at p? For the last predecessor that is the cell's last-opened bit (the specialised
            // DP step does not set the chosen-opened bit for its last slot)
            const bool opened = (c & 4u) || (ord + 1 == pred_off[m + 1] - pred_off[m] && (c & 8u));
            nm = opened ? p : gapm_idx(p, s);
            ns = s;
        }
    };

    // ---- starting point (mesh.h:567-592)
    const uint32_t send = L - 1;
    const float* lastcol = A.lastcol + io;
    const uint32_t* lastnodes = A.lastnodes + io;
    uint32_t m = lastnodes[0];
    float best = lastcol[m];
    for (uint32_t t = 0; t < V; t++) { const float v = lastcol[t]; if (v < best) { best = v; m = t; } }
    uint32_t s = send;
    for (uint32_t i = 0; i < h.n_last; i++) {
        const uint32_t mt = lastnodes[i];
        const float v = A.rowmin[io + mt];
        if (v < best) { best = v; m = mt; s = A.rowarg[io + mt]; }
    }
    r.end_m = m; r.end_s = s;
    const bool keep_case = A.lowercase == 1;
    const bool lc_unaligned = A.lowercase == 2;
    auto qbase = [&](uint32_t i) -> uint8_t { return keep_case ? qm[i] : (uint8_t)(qm[i] & 15u); };

    Out o = { ocols, omasks, 0, 0 };
    // ---- right overhang (mesh.h:594-615)
    const int cutoff_tail = (int)(send - s);
    if (cutoff_tail && A.overhang != 1) {
        int pos = (A.overhang == 0) ? (int)W - 1 - (int)ncol[m] - cutoff_tail : 0;
        for (int i = 0; i < cutoff_tail; i++) {
            uint8_t b = qbase(L - 1 - i);
            if (lc_unaligned) b |= 16;
            const int pp = pos++;
            o.append((uint32_t)(pp > 0 ? pp : 0), b);
        }
    }
    const float rval = best;
    uint32_t pos = W - 1 - ncol[m];
    float sum_weight = 0.f;
    o.append(pos, qbase(s));
    sum_weight = __fadd_rn(sum_weight, __fmul_rn(A.ms, nweight[m]));
    // ---- walk back (mesh.h:642-685)
    while (s != 0 && pred_off[m + 1] != pred_off[m]) {
        uint32_t nm, snew;
        follow(m, s, cell(m, s), nm, snew);
        m = nm;
        if (snew != 0) {  // landing on a cell reached by deletion (its value_sidx == snew): skip it (mesh.h:653-655)
            const uint32_t c2 = cell(m, snew);
            if ((c2 & 3u) == TB_SRC_DEL) {
                uint32_t m2, s2;
                follow(m, snew, c2, m2, s2);
                m = m2;
            }
        }
        pos = W - 1 - ncol[m];
        while (s != snew) {
            --s;
            o.append(pos, qbase(s));
            sum_weight = __fadd_rn(sum_weight, __fmul_rn(A.ms, nweight[m]));
        }
    }
    // ---- left overhang (mesh.h:690-721)
    if (s != 0) {
        r.head = (int)s;
        if (A.overhang == 0) {
            while (s-- != 0) {
                uint8_t b = qbase(s);
                ++pos;
                if (lc_unaligned) b |= 16;
                o.append(pos < W - 1 ? pos : W - 1, b);
            }
        } else if (A.overhang == 2) {
            int n = (int)s;
            while (n--) {
                uint8_t b = qbase((uint32_t)n);
                if (lc_unaligned) b |= 16;
                o.append(W - (uint32_t)n - 1, b);
            }
        }
    }
    r.tail = cutoff_tail;
    // ---- setWidth + reverse (mesh.h:723-724; cseq.cpp:283-289)
    for (uint32_t i = 0; i < o.n / 2; i++) {
        const uint32_t tp = o.pos[i]; o.pos[i] = o.pos[o.n - 1 - i]; o.pos[o.n - 1 - i] = tp;
        const uint8_t tm = o.mask[i]; o.mask[i] = o.mask[o.n - 1 - i]; o.mask[o.n - 1 - i] = tm;
    }
    for (uint32_t i = 0; i < o.n; i++) o.pos[i] = W - 1 - o.pos[i];
    r.n_out = o.n;
    r.raw = rval; r.sum_weight = sum_weight;
    r.score = __fdiv_rn(rval, sum_weight);
    const float q100 = __fmul_rn(100.f, r.score);  // src/align.cpp:509
    r.qual = (int)(q100 < 0.f ? 0.f : (q100 > 100.f ? 100.f : q100));
    r.n_nodes = V;
    r.status = fix_duplicate_positions(o.pos, o.mask, o.n, W, lc_unaligned) ? SG_Q_NOSPACE : SG_Q_ALIGNED;
    A.results[q] = r;
    A.hdr[q].status = GS_DONE;
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
    A.out_cols = s->d_out_cols; A.out_masks = s->d_out_masks; A.results = s->d_results;
    A.ms = -ap.match_score; A.overhang = ap.overhang; A.lowercase = ap.lowercase;
    backtrack_kernel<<<(n + 63) / 64, 64, 0, w->stream>>>(A);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 1;
    return SG_OK;
}

}  // namespace sg
