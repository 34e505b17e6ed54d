// Mesh DP: query x family-graph affine-gap dynamic program with SINA's exact cell semantics.
// Replaces compute() + compute_node_simple::calc + transition_simple + scoring_scheme_simple
// (reference src/mesh.h:305-374,453-528, src/scoring_schemes.h:102-164). Scores are MINIMISED.
//
// Schedule (exact by construction, no scan): one CTA per query, one thread per graph node ("row").
// Nodes are processed in groups of DP_T consecutive ids (ids are topological: column-major).
// A row whose node sits at column rank sigma computes query position s at step t = s + sigma - sigma_lo,
// so every predecessor cell (p,s) and (p,s-1) was produced at an earlier step; the row-internal insertion
// chain (m,s-1) -> (m,s) lives in the thread's registers and is evaluated with the reference's scalar
// recurrence in the reference's order (deletion, insertion, match; tie rules <, <=, <).
// Each step a row publishes (value, gapm_val) into a shared-memory ring indexed by time.
//
// mesh_v2_kernel (normal path): rows of a group are sorted by in-degree so that warps are uniform and run a
// branch-free, fully unrolled step specialised on the warp's in-degree; far predecessors (other group, or
// more than DP_RING-2 column ranks away) are turned into near ones by "ghost" ring columns that loader lanes
// stream from the global spill buffer one step ahead; the same loader lanes spill rows that have far
// successors and track the row minimum of last nodes, so compute lanes touch global memory only for the
// packed traceback (coalesced, one word per 4 steps) and the last-column value.
// mesh_v1_kernel (fallback, hdr.mode == 1): generic per-lane loops, far predecessors read from global.
//
// Traceback: one byte (or halfword when some node has more than 8 predecessors) per cell, packed over time
// and stored coalesced as tb[group][t/4][thread].
#include "common.cuh"

namespace sg {

struct MeshArgs {
    const uint8_t* qmasks; const uint64_t* qoff;
    const GraphHdr* hdr; const GroupInfo* groups; uint32_t gcap, icap, q0;
    const uint8_t* nmask; const float* nweight; const uint32_t* nsigma;
    const uint32_t* pred_off; const uint32_t* pdesc; const int32_t* spillrow; const uint8_t* nflags;
    const uint32_t* pdesc2; const uint32_t* order; const uint16_t* nthr; uint8_t* nshift;
    const GhostInfo* ghosts; const uint32_t* writers;
    uint32_t* tb; float2* spill;
    float* lastcol; float* rowmin; uint32_t* rowarg;
    float ms, mms, gp, gpe;  // -match_score, -mismatch_score, gap_penalty, gap_ext_penalty (align.cpp:406-407)
};

constexpr int T = DP_T;        // rows per group
constexpr int S = DP_BLOCK;    // ring columns (= CTA threads)
constexpr int R = DP_RING;
constexpr int NPR = 4;         // predecessors held in registers (generic kernel)
constexpr int NPF = 8;         // largest in-degree the specialised v2 step is instantiated for
constexpr int QPAD = 512;      // padding either side of the query in shared memory (s runs out of range)
constexpr uint32_t RING_BYTES = sizeof(float2) * R * S;

// ====================================================================================================
// v2: branch-free specialised step
// ====================================================================================================
__device__ __forceinline__ float2 lds_f2(uint32_t byte_off, const unsigned char* smem) {
    return *reinterpret_cast<const float2*>(smem + byte_off);
}

// One group for the lanes of a warp whose rows all have <= NPW predecessors. Slots are right-aligned: a row
// with np < NPW repeats its first predecessor in the NPW-np leading slots (a repeated candidate ties with
// its first copy and strict '<' keeps the first, so nothing changes; backtrack maps slot -> ordinal with
// max(0, slot - shift)). Rows without predecessor get dgp = dgpe = msw = mmsw = +inf so that no deletion or
// match candidate can win, and has_real = false keeps gapm_val at its edge value.
template <int NPW, bool WIDE>
__device__ __forceinline__ void v2_fast_group(const MeshArgs& A, unsigned char* smem, const uint8_t* qm, uint32_t Lq,
                                              uint32_t steps4, const uint32_t* ck, float dgp, float dgpe, int soff,
                                              float initv, bool has_real, uint32_t mask,
                                              float msw, float mmsw, float* lastcol_ptr, uint32_t* tbg) {
    const float gp = A.gp, gpe = A.gpe;
    const float INF = __int_as_float(0x7f800000);
    float pvp[NPW];
#pragma unroll
    for (int k = 0; k < NPW; k++) pvp[k] = 0.f;
    float Ep = 1.0f, Hp = 1.0f;
    uint32_t x = 0;                       // t * S * 8: byte offset of time slot t before wrapping
    int s = -soff;
    const uint32_t wofs = threadIdx.x * 8u;
    const uint32_t RMASK = RING_BYTES - 1;
    for (uint32_t t0 = 0; t0 < steps4; t0 += 4) {
        uint32_t tbw = 0, tbw2 = 0;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t SH = WIDE ? 16u * (u & 1) : 8u * u;   // compile-time shift of this step's cell
            const bool s0 = (s == 0);
            float value = s0 ? 1.0f : initv;                     // init_edge / init (mesh.h:294-301,469-473)
            float gm = 1.0f;
            uint32_t code = 0;
            bool open = false;
            float cur[NPW];
            // ---- deletion over predecessor slots, ascending id (mesh.h:475-478 -> 305-330)
#pragma unroll
            for (int k = 0; k < NPW; k++) {
                const uint32_t a = (x + ck[k]) & RMASK;
                const float2 c = lds_f2(a, smem);
                cur[k] = c.x;
                const float v = __fadd_rn(c.x, dgp);
                const float gv = __fadd_rn(c.y, dgpe);
                open = v < gv;
                gm = open ? v : gv;                               // last predecessor wins
                const bool win = gm < value;
                value = win ? gm : value;
                const uint32_t cd = WIDE ? ((TB_SRC_DEL | (k << 8)) << SH) : ((TB_SRC_DEL | (k << 2)) << SH);
                const uint32_t co = WIDE ? (cd | (4u << SH)) : (cd | (32u << SH));
                code = win ? (open ? co : cd) : code;
            }
            const float gapm = has_real ? gm : 1.0f;
            // ---- insertion from (m, s-1) (mesh.h:486-490 -> 332-358)
            const bool ext = (Ep == Hp);
            float E = ext ? __fadd_rn(Ep, gpe) : __fadd_rn(Hp, gp);
            E = s0 ? 1.0f : E;
            const bool iwin = (E <= value) && !s0;
            value = iwin ? E : value;
            code = iwin ? (TB_SRC_INS << SH) : code;
            // ---- match from (p, s-1) (mesh.h:492-500 -> 360-374)
            float sc = (mask & qm[s]) ? msw : mmsw;
            sc = s0 ? INF : sc;
#pragma unroll
            for (int k = 0; k < NPW; k++) {
                const float v = __fadd_rn(pvp[k], sc);
                const bool win = v < value;
                value = win ? v : value;
                code = win ? (WIDE ? ((TB_SRC_MATCH | (k << 8)) << SH) : ((TB_SRC_MATCH | (k << 2)) << SH)) : code;
            }
            const uint32_t f_open = WIDE ? (8u << SH) : (64u << SH);
            const uint32_t f_ins = WIDE ? (16u << SH) : (128u << SH);
            code |= (open ? f_open : 0u) | ((!ext && !s0) ? f_ins : 0u);
            if (WIDE && u >= 2) tbw2 |= code; else tbw |= code;
#pragma unroll
            for (int k = 0; k < NPW; k++) pvp[k] = cur[k];
            Ep = E;
            Hp = value;
            *reinterpret_cast<float2*>(smem + ((x + wofs) & RMASK)) = make_float2(value, gapm);
            if (s == (int)Lq - 1) *lastcol_ptr = value;
            x += S * 8u;
            s++;
            __syncthreads();
        }
        if (WIDE) {
            tbg[(uint64_t)(t0 >> 1) * T] = tbw;
            tbg[(uint64_t)((t0 >> 1) + 1) * T] = tbw2;
        } else {
            tbg[(uint64_t)(t0 >> 2) * T] = tbw;
        }
    }
}

// Warps holding a row with more than NPR predecessors: slots are looped over.
template <bool WIDE>
__device__ __forceinline__ void v2_generic_group(const MeshArgs& A, unsigned char* smem, const uint8_t* qm, uint32_t Lq,
                                                 uint32_t steps4, uint32_t npw, uint32_t np, const uint32_t* pd,
                                                 int soff, float initv, uint32_t mask, float msw, float mmsw,
                                                 float* lastcol_ptr, uint32_t* tbg) {
    // slot k of this lane is real iff k >= npw - np; real slot k is predecessor ordinal k - (npw - np)
    const float gp = A.gp, gpe = A.gpe;
    const uint32_t shift = npw - np;
    float Ep = 1.0f, Hp = 1.0f;
    const float2* ring = reinterpret_cast<const float2*>(smem);
    float2* ringw = reinterpret_cast<float2*>(smem);
    for (uint32_t t0 = 0; t0 < steps4; t0 += 4) {
        uint32_t tbw = 0, tbw2 = 0;
        for (uint32_t u = 0; u < 4; u++) {
            const uint32_t t = t0 + u;
            const int s = (int)t - soff;
            uint32_t code = 0;
            if (s >= 0 && s < (int)Lq) {
                const bool s0 = s == 0;
                float value = s0 ? 1.0f : initv, gapm = value;
                uint32_t open_last = 0;
                for (uint32_t k = shift; k < npw; k++) {
                    const uint32_t d = __ldg(&pd[k - shift]);
                    const float2 c = ring[((t - (d >> 16)) & (R - 1)) * S + (d & 0xffffu)];
                    const float v = __fadd_rn(c.x, gp), gv = __fadd_rn(c.y, gpe);
                    const bool open = v < gv;
                    const float gm = open ? v : gv;
                    gapm = gm; open_last = open;
                    if (gm < value) { value = gm; code = WIDE ? (TB_SRC_DEL | ((uint32_t)open << 2) | (k << 8)) : (TB_SRC_DEL | (k << 2) | ((uint32_t)open << 5)); }
                }
                float E = 1.0f;
                uint32_t ins_open = 0;
                if (!s0) {
                    const bool ext = (Ep == Hp);
                    E = ext ? __fadd_rn(Ep, gpe) : __fadd_rn(Hp, gp);
                    ins_open = !ext;
                    if (E <= value) { value = E; code = TB_SRC_INS; }
                    const float sc = (mask & qm[s]) ? msw : mmsw;
                    for (uint32_t k = shift; k < npw; k++) {
                        const uint32_t d = __ldg(&pd[k - shift]);
                        const float v = __fadd_rn(ring[((t - 1 - (d >> 16)) & (R - 1)) * S + (d & 0xffffu)].x, sc);
                        if (v < value) { value = v; code = WIDE ? (TB_SRC_MATCH | (k << 8)) : (TB_SRC_MATCH | (k << 2)); }
                    }
                }
                code |= WIDE ? ((open_last << 3) | (ins_open << 4)) : ((open_last << 6) | (ins_open << 7));
                Ep = E; Hp = value;
                ringw[(t & (R - 1)) * S + threadIdx.x] = make_float2(value, gapm);
                if (s == (int)Lq - 1) *lastcol_ptr = value;
            }
            if (WIDE) { if (u >= 2) tbw2 |= code << (16 * (u & 1)); else tbw |= code << (16 * (u & 1)); }
            else tbw |= code << (8 * u);
            __syncthreads();
        }
        if (WIDE) { tbg[(uint64_t)(t0 >> 1) * T] = tbw; tbg[(uint64_t)((t0 >> 1) + 1) * T] = tbw2; }
        else tbg[(uint64_t)(t0 >> 2) * T] = tbw;
    }
}

// Loader lanes (threads DP_T .. DP_BLOCK-1): lane j feeds ghost column DP_T+j from the spill buffer one step
// ahead of its consumers, and drains one row (spill store and/or running row minimum for last nodes).
__device__ __forceinline__ void v2_loader_group(const MeshArgs& A, unsigned char* smem, uint32_t ql, const GraphHdr& h,
                                                const GroupInfo& gi, uint32_t g, uint32_t steps4) {
    const uint32_t j = threadIdx.x - T;
    const uint32_t Lq = h.qlen;
    const uint64_t io = (uint64_t)ql * A.icap;
    float2* ring = reinterpret_cast<float2*>(smem);
    float2* spill = A.spill + h.spill_off;
    // ghost
    const bool is_ghost = j < gi.n_ghost;
    const float2* gsrc = nullptr;
    int gsoff = 0;
    if (is_ghost) {
        const GhostInfo gh = A.ghosts[((uint64_t)ql * A.gcap + g) * DP_G + j];
        gsrc = spill + (uint64_t)gh.spillrow * Lq;
        gsoff = gh.soff;
    }
    // writer
    const bool is_writer = j < gi.n_writer;
    uint32_t wnode = 0, wcol = 0;
    int wsoff = 0, wsr = -1;
    bool wlast = false;
    if (is_writer) {
        wnode = A.writers[((uint64_t)ql * A.gcap + g) * DP_G + j];
        wcol = A.nthr[io + wnode];
        wsoff = (int)(A.nsigma[io + wnode] - gi.sigma_lo);
        wsr = A.spillrow[io + wnode];
        wlast = A.nflags[io + wnode] == 0;
    }
    float rmin = 0.f;
    uint32_t rarg = 0;
    auto gload = [&](int t) -> float2 {  // what the ghost publishes at step t: query position t - gsoff
        const int col = t - gsoff;
        if (is_ghost && col >= 0 && col < (int)Lq) return __ldcg(&gsrc[col]);
        return make_float2(0.f, 0.f);
    };
    // prologue = step -1 (a ghost with soff -1 must have position 0 in slot -1 before step 0); afterwards the
    // data of step t+4 is requested at step t, so the L2 latency never sits between two barriers
    // (GHOST_LEAD = 4 + 2 keeps that request behind the source row's spill store).
    float2 pf[4];
    {
        const float2 c = gload(-1);
        if (is_ghost) ring[((uint32_t)(-1) & (R - 1)) * S + threadIdx.x] = c;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) pf[k] = gload(k);
    __syncthreads();
    auto drain = [&](uint32_t t) {  // what the writer's row published at step t-1
        const int sw = (int)t - 1 - wsoff;
        if (is_writer && sw >= 0 && sw < (int)Lq) {
            const float2 c = ring[((t - 1) & (R - 1)) * S + wcol];
            if (wsr >= 0) __stcg(&spill[(uint64_t)wsr * Lq + sw], c);
            if (wlast && (sw == 0 || c.x < rmin)) { rmin = c.x; rarg = (uint32_t)sw; }
        }
    };
    for (uint32_t t0 = 0; t0 < steps4; t0 += 4) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t t = t0 + u;
            const float2 c = pf[u];
            pf[u] = gload((int)t + 4);
            if (is_ghost) ring[(t & (R - 1)) * S + threadIdx.x] = c;
            if (t >= 1) drain(t);
            __syncthreads();
        }
    }
    drain(steps4);
    if (is_writer && wlast) { A.rowmin[io + wnode] = rmin; A.rowarg[io + wnode] = rarg; }
}

template <bool WIDE>
__device__ __forceinline__ void v2_query(const MeshArgs& A, const GraphHdr& h, uint32_t ql, unsigned char* smem,
                                         const uint8_t* qm) {
    const uint32_t tid = threadIdx.x;
    const uint32_t Lq = h.qlen;
    const uint64_t io = (uint64_t)ql * A.icap;
    const uint32_t* pred_off = A.pred_off + (uint64_t)ql * (A.icap + 1);
    const uint32_t* pdesc2 = A.pdesc2 + io;
    uint32_t* tbq = A.tb + h.tb_off;
    const uint32_t RMASK = RING_BYTES - 1;

    for (uint32_t g = 0; g < h.n_groups; g++) {
        const GroupInfo gi = A.groups[(uint64_t)ql * A.gcap + g];
        const uint32_t steps4 = (Lq + gi.depth - 1 + 3) & ~3u;
        if (tid >= (uint32_t)T) {
            v2_loader_group(A, smem, ql, h, gi, g, steps4);
        } else {
            const uint32_t m = A.order[((uint64_t)ql * A.gcap + g) * T + tid];
            const bool valid = m != 0xFFFFFFFFu;
            uint32_t np = 0, pbase = 0, mask = 0;
            int soff = 0;
            float msw = 0.f, mmsw = 0.f;
            float* lastcol_ptr = A.lastcol + io;  // never stored through for invalid lanes (s stays out of range)
            if (valid) {
                pbase = pred_off[m];
                np = pred_off[m + 1] - pbase;
                mask = A.nmask[io + m] & 15u;
                const float w = A.nweight[io + m];
                msw = __fmul_rn(A.ms, w);    // (comp ? match : mismatch) * weight  (scoring_schemes.h:150-156)
                mmsw = __fmul_rn(A.mms, w);
                soff = (int)(A.nsigma[io + m] - gi.sigma_lo);
                lastcol_ptr = A.lastcol + io + m;
            } else {
                soff = (int)(steps4 + 8);    // idle lane: s stays negative, nothing it computes is ever read
            }
            const uint32_t npw = max(1u, __reduce_max_sync(0xffffffffu, np));
            if (valid) A.nshift[io + m] = (uint8_t)(npw - np);
            const float initv = np == 0 ? 1.0f : 1000000.0f;
            uint32_t* tbg = tbq + gi.tb_off + tid;
            __syncthreads();  // matches the loader's prologue barrier
            if (npw <= (uint32_t)NPF) {
                uint32_t ck[NPF];
                const uint32_t shift = npw - np;
#pragma unroll
                for (int k = 0; k < NPF; k++) {
                    ck[k] = tid * 8u;  // rows without predecessor: any valid cell, its candidates are +inf
                    if (np > 0 && k < (int)npw) {
                        const uint32_t ord = (uint32_t)k > shift ? (uint32_t)k - shift : 0u;
                        const uint32_t d = pdesc2[pbase + ord];
                        // slot (t - delta) & 15, column c  ->  byte ((t - delta)*S + c)*8, wrapped by RMASK
                        ck[k] = ((d & 0xffffu) * 8u - (d >> 16) * (S * 8u)) & RMASK;
                    }
                }
                const float INF = __int_as_float(0x7f800000);
                const bool hr = np > 0;
                const float dgp = hr ? A.gp : INF, dgpe = hr ? A.gpe : INF;
                if (!hr) { msw = INF; mmsw = INF; }
#define V2_CASE(N) case N: v2_fast_group<N, WIDE>(A, smem, qm, Lq, steps4, ck, dgp, dgpe, soff, initv, hr, mask, msw, mmsw, lastcol_ptr, tbg); break;
                switch (npw) {
                    V2_CASE(1) V2_CASE(2) V2_CASE(3) V2_CASE(4) V2_CASE(5) V2_CASE(6) V2_CASE(7)
                    default: v2_fast_group<8, WIDE>(A, smem, qm, Lq, steps4, ck, dgp, dgpe, soff, initv, hr, mask, msw, mmsw, lastcol_ptr, tbg); break;
                }
#undef V2_CASE
            } else {
                v2_generic_group<WIDE>(A, smem, qm, Lq, steps4, npw, np, pdesc2 + pbase, soff,
                                       initv, mask, msw, mmsw, lastcol_ptr, tbg);
            }
        }
        __syncthreads();  // ring and spill rows of this group are complete before the next group starts
    }
}

__global__ void __launch_bounds__(DP_BLOCK, 2) mesh_v2_kernel(MeshArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t q = A.q0 + blockIdx.x;
    const GraphHdr h = A.hdr[q];
    if (h.status != GS_OK || h.mode != 2) return;
    uint8_t* qm = smem + RING_BYTES + 16 + QPAD;   // valid for s in [-QPAD, Lq + QPAD)
    const uint8_t* src = A.qmasks + A.qoff[q];
    for (uint32_t i = threadIdx.x; i < h.qlen + 2 * QPAD; i += blockDim.x) {
        const int s = (int)i - QPAD;
        qm[s] = (s >= 0 && s < (int)h.qlen) ? (src[s] & 15u) : 0;
    }
    for (uint32_t i = threadIdx.x; i < RING_BYTES / 8; i += blockDim.x)  // no NaN bit patterns in unwritten cells
        reinterpret_cast<float2*>(smem)[i] = make_float2(0.f, 0.f);
    __syncthreads();
    if (h.wide) v2_query<true>(A, h, blockIdx.x, smem, qm);
    else v2_query<false>(A, h, blockIdx.x, smem, qm);
}

// ====================================================================================================
// v1: generic fallback (hdr.mode == 1). Rows keep id order inside a group; far predecessors are read
// straight from the spill buffer; every row tracks its own spill / row-minimum.
// ====================================================================================================
template <bool WIDE>
__device__ __forceinline__ void v1_query(const MeshArgs& A, const GraphHdr& h, uint32_t ql, float2* ring,
                                         const uint8_t* qm) {
    const uint32_t tid = threadIdx.x;
    const uint32_t Lq = h.qlen, V = h.V;
    const uint64_t io = (uint64_t)ql * A.icap;
    const uint32_t* pred_off = A.pred_off + (uint64_t)ql * (A.icap + 1);
    const uint32_t* pdesc = A.pdesc + io;
    const float2* spill = A.spill + h.spill_off;
    float2* spill_w = A.spill + h.spill_off;
    uint32_t* tbq = A.tb + h.tb_off;
    const float gp = A.gp, gpe = A.gpe;

    for (uint32_t g = 0; g < h.n_groups; g++) {
        const GroupInfo gi = A.groups[(uint64_t)ql * A.gcap + g];
        const uint32_t m = g * T + tid;
        const bool valid = tid < (uint32_t)T && m < V;
        uint32_t np = 0, pbase = 0, mask = 0;
        int soff = 0, sr = -1;
        float msw = 0.f, mmsw = 0.f;
        bool is_last = false;
        uint32_t pd[NPR];
        float pv_prev[NPR];
#pragma unroll
        for (int i = 0; i < NPR; i++) { pd[i] = 0; pv_prev[i] = 0.f; }
        if (valid) {
            pbase = pred_off[m];
            np = pred_off[m + 1] - pbase;
            mask = A.nmask[io + m];
            const float w = A.nweight[io + m];
            msw = __fmul_rn(A.ms, w);
            mmsw = __fmul_rn(A.mms, w);
            soff = (int)(A.nsigma[io + m] - gi.sigma_lo);
            sr = A.spillrow[io + m];
            is_last = A.nflags[io + m] == 0;
            A.nshift[io + m] = 0;
#pragma unroll
            for (int i = 0; i < NPR; i++) if ((uint32_t)i < np) pd[i] = pdesc[pbase + i];
        }
        float E_prev = 1.0f, H_prev = 1.0f;  // gaps_val / value of (m, s-1)
        float rmin = 0.f;
        uint32_t rarg = 0;
        const uint32_t steps = (Lq + gi.depth - 1 + 3) & ~3u;
        uint32_t tbw = 0;
        uint32_t* tbg = tbq + gi.tb_off;

        for (uint32_t t = 0; t < steps; t++) {
            const int s = (int)t - soff;
            uint32_t code = 0;
            if (valid && s >= 0 && s < (int)Lq) {
                const bool edge = (np == 0) || (s == 0);                 // init_edge / init (mesh.h:294-301,469-473)
                float value = edge ? 1.0f : 1000000.0f;
                float gapm = value;
                uint32_t open_last = 0;
                float pv_cur[NPR];
                auto del_step = [&](uint32_t i, float2 c) {              // mesh.h:305-330
                    const float v = __fadd_rn(c.x, gp);
                    const float gv = __fadd_rn(c.y, gpe);
                    const bool open = v < gv;
                    const float gm = open ? v : gv;
                    gapm = gm;                 // last predecessor wins
                    open_last = open;
                    if (gm < value) {
                        value = gm;
                        code = WIDE ? (TB_SRC_DEL | ((uint32_t)open << 2) | (i << 8))
                                    : (TB_SRC_DEL | (i << 2) | ((uint32_t)open << 5));
                    }
                };
                auto load_cell = [&](uint32_t d, int ss, uint32_t tt) -> float2 {
                    if (d & FAR_BIT) return __ldcg(&spill[(uint64_t)(d & ~FAR_BIT) * Lq + ss]);
                    return ring[((tt - (d >> 16)) & (R - 1)) * S + (d & 0xffffu)];
                };
#pragma unroll
                for (int i = 0; i < NPR; i++) {
                    if ((uint32_t)i < np) {
                        const float2 c = load_cell(pd[i], s, t);
                        pv_cur[i] = c.x;
                        del_step(i, c);
                    } else pv_cur[i] = 0.f;
                }
                for (uint32_t i = NPR; i < np; i++) {
                    const uint32_t d = __ldg(&pdesc[pbase + i]);
                    del_step(i, load_cell(d, s, t));
                }
                float E = 1.0f;
                uint32_t ins_open = 0;
                if (s > 0) {
                    const bool ext = (E_prev == H_prev);                 // mesh.h:332-358
                    E = ext ? __fadd_rn(E_prev, gpe) : __fadd_rn(H_prev, gp);
                    ins_open = !ext;
                    if (E <= value) { value = E; code = TB_SRC_INS; }
                    const float sc = (mask & qm[s] & 15u) ? msw : mmsw;  // mesh.h:360-374
#pragma unroll
                    for (int i = 0; i < NPR; i++) {
                        if ((uint32_t)i < np) {
                            const float v = __fadd_rn(pv_prev[i], sc);
                            if (v < value) { value = v; code = WIDE ? (TB_SRC_MATCH | ((uint32_t)i << 8)) : (TB_SRC_MATCH | ((uint32_t)i << 2)); }
                        }
                    }
                    for (uint32_t i = NPR; i < np; i++) {
                        const uint32_t d = __ldg(&pdesc[pbase + i]);
                        const float v = __fadd_rn(load_cell(d, s - 1, t - 1).x, sc);
                        if (v < value) { value = v; code = WIDE ? (TB_SRC_MATCH | (i << 8)) : (TB_SRC_MATCH | (i << 2)); }
                    }
                }
                code |= WIDE ? ((open_last << 3) | (ins_open << 4)) : ((open_last << 6) | (ins_open << 7));
#pragma unroll
                for (int i = 0; i < NPR; i++) pv_prev[i] = pv_cur[i];
                E_prev = E;
                H_prev = value;
                const float2 out = make_float2(value, gapm);
                ring[(t & (R - 1)) * S + tid] = out;
                if (sr >= 0) __stcg(&spill_w[(uint64_t)sr * Lq + s], out);
                if (s == (int)Lq - 1) A.lastcol[io + m] = value;
                if (is_last && (s == 0 || value < rmin)) { rmin = value; rarg = (uint32_t)s; }
            }
            if (tid < (uint32_t)T) {
                if (WIDE) {
                    tbw |= code << (16 * (t & 1));
                    if ((t & 1) == 1) { tbg[(uint64_t)(t >> 1) * T + tid] = tbw; tbw = 0; }
                } else {
                    tbw |= code << (8 * (t & 3));
                    if ((t & 3) == 3) { tbg[(uint64_t)(t >> 2) * T + tid] = tbw; tbw = 0; }
                }
            }
            __syncthreads();
        }
        if (valid && is_last) { A.rowmin[io + m] = rmin; A.rowarg[io + m] = rarg; }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(DP_BLOCK, 2) mesh_v1_kernel(MeshArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    float2* ring = reinterpret_cast<float2*>(smem);            // [R][S]
    uint8_t* qm = smem + RING_BYTES + 16 + QPAD;
    const uint32_t q = A.q0 + blockIdx.x;
    const GraphHdr h = A.hdr[q];
    if (h.status != GS_OK || h.mode != 1) return;
    const uint8_t* src = A.qmasks + A.qoff[q];
    for (uint32_t i = threadIdx.x; i < h.qlen; i += blockDim.x) qm[i] = src[i];
    __syncthreads();
    if (h.wide) v1_query<true>(A, h, blockIdx.x, ring, qm);
    else v1_query<false>(A, h, blockIdx.x, ring, qm);
}

int launch_mesh(Session* s, Workspace* w, const sg_align_params& ap, uint32_t q0, uint32_t n) {
    MeshArgs A;
    A.qmasks = s->d_qmasks; A.qoff = s->d_qoff; A.hdr = s->d_hdr; A.groups = w->d_groups;
    A.gcap = s->gcap; A.icap = s->icap; A.q0 = q0;
    A.nmask = w->d_nmask; A.nweight = w->d_nweight; A.nsigma = w->d_nsigma; A.pred_off = w->d_pred_off;
    A.pdesc = w->d_pdesc; A.spillrow = w->d_spillrow; A.nflags = w->d_nflags;
    A.pdesc2 = w->d_pdesc2; A.order = w->d_order; A.nthr = w->d_nthr; A.nshift = w->d_nshift;
    A.ghosts = w->d_ghosts; A.writers = w->d_writers;
    A.tb = w->d_tb; A.spill = w->d_spill; A.lastcol = w->d_lastcol; A.rowmin = w->d_rowmin; A.rowarg = w->d_rowarg;
    A.ms = -ap.match_score; A.mms = -ap.mismatch_score; A.gp = ap.gap_penalty; A.gpe = ap.gap_ext_penalty;
    uint32_t max_qlen = 0;
    for (uint32_t i = q0; i < q0 + n; i++) {
        uint32_t l = (uint32_t)(s->h_qoff[i + 1] - s->h_qoff[i]);
        if (l > max_qlen) max_qlen = l;
    }
    size_t smem = RING_BYTES + 16 + 2 * QPAD + ((max_qlen + 15) & ~15u);
    if (smem > 220 * 1024) SG_FAIL(SG_ERR_LIMIT, "query too long for the DP kernel's shared memory");
    SG_CUDA(cudaFuncSetAttribute(mesh_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SG_CUDA(cudaFuncSetAttribute(mesh_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mesh_v2_kernel<<<n, DP_BLOCK, smem, w->stream>>>(A);
    mesh_v1_kernel<<<n, DP_BLOCK, smem, w->stream>>>(A);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 2;
    return SG_OK;
}

}  // namespace sg
