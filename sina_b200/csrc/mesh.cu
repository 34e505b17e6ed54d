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
    const uint32_t* pdesc2; const uint32_t* order; const uint8_t* rcol; const uint16_t* nthr; uint8_t* nshift;
    const GhostInfo* ghosts; const uint32_t* writers;
    const uint32_t* nmaxins; int forbid;   // --insertion forbid (generic kernel only): free columns after each node
    uint32_t* tb; float2* spill;
    float* lastcol; float* rowmin; uint32_t* rowarg;
    float ms, mms, gp, gpe;  // -match_score, -mismatch_score, gap_penalty, gap_ext_penalty (align.cpp:406-407)
};

constexpr int T = DP_T;        // rows per group
constexpr int S = DP_BLOCK;    // ring columns (= CTA threads)
constexpr int R = DP_RING;
constexpr int NPR = 4;         // predecessors held in registers (generic kernel)
constexpr int NPF = 8;         // largest in-degree the specialised v2 step is instantiated for
constexpr int QPAD = 256;      // padding either side of the query in shared memory (s runs out of range by < T + 4)
static_assert(QPAD >= DP_T + 8, "query padding must cover the group skew");
constexpr int RS = S + 3;      // cells per time slot of the v2 ring: the odd stride rotates the banks by three cells from one
                               // time slot to the next, so that readers of the same column at different column-rank
                               // distances (a very common pattern: a row skipping a column) do not collide
constexpr uint32_t SLOT_BYTES = sizeof(float2) * RS;      // one time slot of the ring
constexpr uint32_t RB = SLOT_BYTES * R;                   // one copy of the ring
constexpr uint32_t RING_BYTES = 2 * RB;                   // the ring is stored twice back to back (see below)

// ====================================================================================================
// v2: branch-free specialised step
//
// Ring addressing. A row publishes the cell of step t into time slot t & (R-1), column = its thread, of BOTH
// copies of the ring. A consumer reads predecessor p (delta = difference of column ranks, 1..R-2) at byte
//     slot(t) + ck,   slot(t) = (t & (R-1)) * SLOT_BYTES  (uniform),  ck = col(p)*8 + ((R - delta) & (R-1)) * SLOT_BYTES
// which lands on time slot (t - delta) & (R-1) of copy A when it does not wrap and on the same slot of copy B
// when it does, so the per-lane address needs no add-and-mask: it is register + uniform register in the LDS itself.
// ====================================================================================================
// explicit shared-window accesses: address = per-lane register + uniform slot offset, which ptxas folds into
// the LDS/STS operand ([R + UR + imm]) so that no per-lane integer op is spent on addressing
__device__ __forceinline__ float2 lds_f2(uint32_t addr) {
    float2 r;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
    return r;
}
// compare into a float 0/1 (FSET.BF): the raw traceback bits are accumulated with FFMA, off the ALU pipe
__device__ __forceinline__ float fset_lt(float a, float b) {
    float r;
    asm("set.lt.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float fset_le(float a, float b) {
    float r;
    asm("set.le.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
template <uint32_t OFF>
__device__ __forceinline__ void sts_f2(uint32_t addr, float2 v) {
    asm volatile("st.shared.v2.f32 [%0+%3], {%1, %2};" ::"r"(addr), "f"(v.x), "f"(v.y), "n"(OFF) : "memory");
}

// Per-lane constants of the specialised step.
template <int NPW>
struct V2Lane {
    uint32_t pk[NPW];        // shared-window byte address of predecessor slot k at time slot 0
    uint32_t wadr;           // own ring column
    uint32_t qrow;           // query-table address of (position 0, plane of the node's mask) minus 8*t_first
    uint32_t mmsw_bits, dsc; // PLANES > 1: bits of mismatch*weight, and bits(match*weight) - bits(mismatch*weight)
    float msw, mmsw;         // PLANES == 1: (mis)match score * weight (+inf for rows without predecessor)
    uint32_t mask;           // PLANES == 1: the node's IUPAC mask
    float initv;
    uint32_t t_first, t_last;
    float* lastcol_ptr;
};

// Two steps t0, t0+1 of one group for the lanes of a warp whose rows all have <= NPW predecessors.
// Slots are right-aligned: the NPW-np leading slots of a row with np < NPW read a constant (+inf, +inf) ring column,
// a candidate that never wins (backtrack maps slot -> ordinal with slot - shift).
// A ring cell is (value, dm): dm = min(value + gap, gapm_val + gapext) is the deletion candidate this cell offers
// to every successor row (deletion(), mesh.h:305-330, evaluated once by the row it leaves from instead of once per
// edge); the row's own gapm_val is the dm of its last predecessor ("last predecessor wins").
// Rows without predecessor read the constant ring column (value 1, dm 1): the deletion candidate 1 is never below
// the initial 1 and leaves gapm_val = 1 (init_edge, mesh.h:294-301); their match score is +inf (mmsw_bits = +inf,
// dsc = 0), so no match candidate can win either.
// The match score is read from the query table as a 0/1 byte and turned into the float's bits with integer
// arithmetic (exact, and off the ALU pipe that bounds this kernel).
// En carries gaps_val of the next cell of the row: (m,s)'s insertion candidate is fixed when (m,s-1) is finished.
// EDGES: some lane may be at s == 0 or s == Lq-1 in these steps; the other (vast majority of) steps skip those tests.
template <int NPW, bool WIDE, bool EDGES, int PLANES>
__device__ __forceinline__ void v2_steps2(const V2Lane<NPW>& L, const float gp, const float gpe, const uint32_t t0,
                                          float (&pvp)[NPW], float& En, uint32_t& tbw) {
    const float INF = __int_as_float(0x7f800000);
    constexpr bool RAW = v2_raw_cells(NPW, WIDE);
    float acc = 0.f;   // RAW: the two cells' bits as a small integer held in a float
#pragma unroll
    for (int u = 0; u < 2; u++) {
        const uint32_t t = t0 + u;
        const uint32_t xs = (t & (R - 1)) * SLOT_BYTES;      // uniform
        const uint32_t SH = (WIDE ? 16u : 8u) * u;           // compile-time shift of this step's cell
        const bool s0 = EDGES && (t == L.t_first);
        // init (mesh.h:294-301,469-473): 1000000, or 1 for rows without predecessor. The s == 0 column is an
        // edge too (init 1): there the forced insertion candidate E = 1 below supplies that 1.
        float value = L.initv;
        float gmin = 1.0f;
        uint32_t code = 0;
        float cur[NPW];
        // ---- deletion over predecessor slots, ascending id (mesh.h:475-478 -> 305-330)
#pragma unroll
        for (int k = 0; k < NPW; k++) {
            const float2 c = lds_f2(L.pk[k] + xs);
            cur[k] = c.x;
            gmin = c.y;                                       // last predecessor wins
            // min() and the compare feed different consumers: the running value only depends on the FMNMX chain
            // (this step's critical path to the ring store), the predicates only feed the traceback code.
            // min(a, b) == (a < b ? a : b) here: no NaN, and a -0 cannot arise from these sums.
            if (RAW) {
                acc = fmaf(fset_lt(c.y, value), (float)((TBR_DEL << k) << (8 * u)), acc);
                value = fminf(value, c.y);
            } else {
                const bool win = c.y < value;
                value = fminf(value, c.y);
                code = win ? (WIDE ? ((TB_SRC_DEL | (k << 8)) << SH) : ((TB_SRC_DEL | (k << 2)) << SH)) : code;
            }
        }
        // ---- insertion from (m, s-1) (mesh.h:486-490 -> 332-358). At s == 0 the reference evaluates no
        // insertion, gaps_val stays 1 and value starts from 1: E is forced to 1, so value = min(1, deletions)
        // exactly as there (the traceback of an s == 0 cell is never followed).
        float E = En;
        if (EDGES) E = s0 ? 1.0f : E;
        if (RAW) {
            acc = fmaf(fset_le(E, value), (float)(TBR_INS << (8 * u)), acc);
            value = fminf(value, E);
        } else {
            const bool iwin = (E <= value);
            value = fminf(value, E);
            code = iwin ? (TB_SRC_INS << SH) : code;
        }
        // ---- match from (p, s-1) (mesh.h:492-500 -> 360-374)
        float sc;
        if (PLANES == 1) sc = (L.mask & lds_u8(L.qrow + t)) ? L.msw : L.mmsw;   // comp(): the IUPAC masks intersect
        else sc = __uint_as_float(lds_u8(L.qrow + (uint32_t)PLANES * t) * L.dsc + L.mmsw_bits);
        if (EDGES) sc = s0 ? INF : sc;
#pragma unroll
        for (int k = 0; k < NPW; k++) {
            const float v = __fadd_rn(pvp[k], sc);
            if (RAW) {
                acc = fmaf(fset_lt(v, value), (float)((TBR_MATCH << k) << (8 * u)), acc);
                value = fminf(value, v);
            } else {
                const bool win = v < value;
                value = fminf(value, v);
                code = win ? (WIDE ? ((TB_SRC_MATCH | (k << 8)) << SH) : ((TB_SRC_MATCH | (k << 2)) << SH)) : code;
            }
        }
        // ---- what this cell offers: the deletion candidate of its successors and the row's next insertion
        const float vgp = __fadd_rn(value, gp);
        const float ggpe = __fadd_rn(gmin, gpe);
        const float dm = fminf(vgp, ggpe);
        if (RAW) acc = fmaf(fset_lt(vgp, ggpe), (float)(TBR_OB << (8 * u)), acc);
        else tbw |= code | ((vgp < ggpe) ? (WIDE ? (4u << SH) : (32u << SH)) : 0u);
        En = (E == value) ? __fadd_rn(E, gpe) : vgp;          // extension iff gaps_val == value (mesh.h:340-349)
#pragma unroll
        for (int k = 0; k < NPW; k++) pvp[k] = cur[k];
        const float2 out = make_float2(value, dm);
        sts_f2<0>(L.wadr + xs, out);
        sts_f2<RB>(L.wadr + xs, out);
        if (EDGES && t == L.t_last) *L.lastcol_ptr = value;
        __syncthreads();
    }
    // the integer sits in the low mantissa bits of acc + 2^23 (acc < 2^16, exact)
    if (RAW) tbw = __float_as_uint(__fadd_rn(acc, 8388608.0f));
}

// [w_first, w_last] = steps at which some row of this warp is inside the query; outside of it the warp only keeps
// the barriers (nothing it would publish is read: a consumer's window starts after its predecessors' and ends after
// theirs). [b_first, b_last] = steps at which every row of the warp is strictly inside (0 < s < Lq-1).
template <int NPW, bool WIDE, int PLANES>
__device__ __forceinline__ void v2_fast_group(const V2Lane<NPW>& L, const float gp, const float gpe, uint32_t steps4,
                                              uint32_t w_first, uint32_t w_last, uint32_t b_first, uint32_t b_last,
                                              uint32_t* tbg) {
    float pvp[NPW];
#pragma unroll
    for (int k = 0; k < NPW; k++) pvp[k] = 0.f;
    float En = 1.0f;
    uint16_t* tbg16 = reinterpret_cast<uint16_t*>(tbg);   // u8 cells: this lane's halfword of step pair 0
    // Step pairs [0, e0) and [e1, steps4): the warp only keeps the barriers. [e0, c0) and [c1, e1): some row is at
    // an edge of the query (EDGES variant). [c0, c1): every row strictly inside; that loop carries no window tests.
    const uint32_t e0 = min(w_first & ~1u, steps4);
    const uint32_t e1 = min((w_last & ~1u) + 2u, steps4);
    uint32_t c0 = (b_first + 1u) & ~1u, c1 = (b_last + 1u) & ~1u;   // bulk pairs: t0 >= b_first && t0 + 1 <= b_last
    if (b_last == 0xFFFFFFFFu || b_last < b_first || c0 >= c1 || c0 < e0 || c1 > e1) c0 = c1 = e1;
    for (uint32_t t0 = 0; t0 < e0; t0 += 2) { __syncthreads(); __syncthreads(); }
#pragma unroll 1
    for (int ph = 0; ph < 2; ph++) {
        const uint32_t a = ph ? c1 : e0, b = ph ? e1 : c0;
        for (uint32_t t0 = a; t0 < b; t0 += 2) {
            uint32_t tbw = 0;
            v2_steps2<NPW, WIDE, true, PLANES>(L, gp, gpe, t0, pvp, En, tbw);
            if (WIDE) tbg[(uint64_t)(t0 >> 1) * T] = tbw;
            else tbg16[(uint64_t)(t0 >> 1) * T] = (uint16_t)tbw;
        }
        if (ph == 0) {
            for (uint32_t t0 = c0; t0 < c1; t0 += 2) {
                uint32_t tbw = 0;
                v2_steps2<NPW, WIDE, false, PLANES>(L, gp, gpe, t0, pvp, En, tbw);
                if (WIDE) tbg[(uint64_t)(t0 >> 1) * T] = tbw;
                else tbg16[(uint64_t)(t0 >> 1) * T] = (uint16_t)tbw;
            }
        }
    }
    for (uint32_t t0 = e1; t0 < steps4; t0 += 2) { __syncthreads(); __syncthreads(); }
}

template <int NPW, bool WIDE, int PLANES>
__device__ __forceinline__ void v2_fast_dispatch(const MeshArgs& A, uint32_t sring, uint32_t sq, const uint32_t* ck, uint32_t rcol,
                                                 bool valid, uint32_t np, int soff, uint32_t Lq, uint32_t plane, uint32_t mask, float w,
                                                 uint32_t steps4, float* lastcol_ptr, uint32_t* tbg) {
    V2Lane<NPW> L;
#pragma unroll
    for (int k = 0; k < NPW; k++) L.pk[k] = sring + ck[k];
    L.wadr = sring + rcol * 8u;
    const bool hr = np > 0;
    const float msw = __fmul_rn(A.ms, w);    // (comp ? match : mismatch) * weight  (scoring_schemes.h:150-156)
    const float mmsw = __fmul_rn(A.mms, w);
    L.initv = hr ? 1000000.0f : 1.0f;
    L.mmsw_bits = hr ? __float_as_uint(mmsw) : 0x7f800000u;
    L.dsc = hr ? __float_as_uint(msw) - __float_as_uint(mmsw) : 0u;
    L.msw = hr ? msw : __int_as_float(0x7f800000);
    L.mmsw = hr ? mmsw : __int_as_float(0x7f800000);
    L.mask = mask;
    L.t_first = valid ? (uint32_t)soff : 0xFFFFFFFFu;
    L.t_last = valid ? (uint32_t)soff + Lq - 1 : 0xFFFFFFFFu;
    L.lastcol_ptr = lastcol_ptr;
    L.qrow = sq + (valid ? (PLANES == 1 ? 0u : plane) - (uint32_t)PLANES * (uint32_t)soff : 0u);
    const uint32_t w_first = __reduce_min_sync(0xffffffffu, L.t_first);
    const uint32_t w_last = __reduce_max_sync(0xffffffffu, valid ? L.t_last : 0u);
    const uint32_t b_first = __reduce_max_sync(0xffffffffu, valid ? L.t_first : 0u) + 1u;
    const uint32_t b_last = __reduce_min_sync(0xffffffffu, L.t_last) - 1u;
    v2_fast_group<NPW, WIDE, PLANES>(L, A.gp, A.gpe, steps4, w_first, w_last, b_first, b_last, tbg);
}

// Warps holding a row with more than NPF predecessors: slots are looped over.
template <bool WIDE, int PLANES>
__device__ __forceinline__ void v2_generic_group(const MeshArgs& A, unsigned char* smem, const uint8_t* qt, uint32_t Lq,
                                                 uint32_t steps4, uint32_t npw, uint32_t np, const uint32_t* pd, uint32_t rcol,
                                                 int soff, float initv, uint32_t plane, uint32_t mask, float msw, float mmsw,
                                                 float* lastcol_ptr, uint32_t* tbg) {
    // slot k of this lane is real iff k >= npw - np; real slot k is predecessor ordinal k - (npw - np)
    const float gp = A.gp, gpe = A.gpe;
    const uint32_t shift = npw - np;
    float En = 1.0f;
    const float2* ring = reinterpret_cast<const float2*>(smem);
    float2* ringw = reinterpret_cast<float2*>(smem);
    uint16_t* tbg16 = reinterpret_cast<uint16_t*>(tbg);
    for (uint32_t t0 = 0; t0 < steps4; t0 += 2) {
        uint32_t tbw = 0;
        for (uint32_t u = 0; u < 2; u++) {
            const uint32_t t = t0 + u;
            const int s = (int)t - soff;
            uint32_t code = 0;
            if (s >= 0 && s < (int)Lq) {
                const bool s0 = s == 0;
                float value = s0 ? 1.0f : initv, gapm = value;
                for (uint32_t k = shift; k < npw; k++) {
                    const uint32_t d = __ldg(&pd[k - shift]);
                    const float gm = ring[((t - (d >> 16)) & (R - 1)) * RS + (d & 0xffffu)].y;   // the predecessor's dm
                    gapm = gm;
                    if (gm < value) { value = gm; code = WIDE ? (TB_SRC_DEL | (k << 8)) : (TB_SRC_DEL | (k << 2)); }
                }
                float E = 1.0f;
                if (!s0) {
                    E = En;
                    if (E <= value) { value = E; code = TB_SRC_INS; }
                    const float sc = (PLANES == 1 ? (mask & qt[s]) : qt[s * PLANES + (int)plane]) ? msw : mmsw;
                    for (uint32_t k = shift; k < npw; k++) {
                        const uint32_t d = __ldg(&pd[k - shift]);
                        const float v = __fadd_rn(ring[((t - 1 - (d >> 16)) & (R - 1)) * RS + (d & 0xffffu)].x, sc);
                        if (v < value) { value = v; code = WIDE ? (TB_SRC_MATCH | (k << 8)) : (TB_SRC_MATCH | (k << 2)); }
                    }
                }
                const float vgp = __fadd_rn(value, gp), ggpe = __fadd_rn(gapm, gpe);
                if (vgp < ggpe) code |= WIDE ? 4u : 32u;
                En = (E == value) ? __fadd_rn(E, gpe) : vgp;
                const float2 out = make_float2(value, fminf(vgp, ggpe));
                ringw[(t & (R - 1)) * RS + rcol] = out;
                ringw[(R + (t & (R - 1))) * RS + rcol] = out;   // second copy, read by the specialised warps
                if (s == (int)Lq - 1) *lastcol_ptr = value;
            }
            tbw |= code << ((WIDE ? 16 : 8) * u);
            __syncthreads();
        }
        if (WIDE) tbg[(uint64_t)(t0 >> 1) * T] = tbw;
        else tbg16[(uint64_t)(t0 >> 1) * T] = (uint16_t)tbw;
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
        wcol = A.rcol[((uint64_t)ql * A.gcap + g) * T + A.nthr[io + wnode]];   // ring column of the row's thread
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
    auto gstore = [&](uint32_t t, float2 c) {
        ring[(t & (R - 1)) * RS + threadIdx.x] = c;
        ring[(R + (t & (R - 1))) * RS + threadIdx.x] = c;
    };
    // the last loader lane owns the constant edge column (value 1, dm 1) read by rows without predecessor, the one
    // before it the constant column (+inf, +inf) read by the padding slots of rows with fewer predecessors than
    // their warp is specialised on: a candidate that can never win, and one shared address for all of them
    if (threadIdx.x >= S - 2) {
        const float cv = threadIdx.x == S - 1 ? 1.0f : __int_as_float(0x7f800000);
        for (uint32_t r = 0; r < 2 * R; r++) ring[r * RS + threadIdx.x] = make_float2(cv, cv);
    }
    // prologue = step -1 (a ghost with soff -1 must have position 0 in slot -1 before step 0); afterwards the
    // data of step t+4 is requested at step t, so the L2 latency never sits between two barriers
    // (GHOST_LEAD = 4 + 2 keeps that request behind the source row's spill store).
    float2 pf[4];
    {
        const float2 c = gload(-1);
        if (is_ghost) gstore((uint32_t)(-1), c);
    }
#pragma unroll
    for (int k = 0; k < 4; k++) pf[k] = gload(k);
    __syncthreads();
    auto drain = [&](uint32_t t) {  // what the writer's row published at step t-1
        const int sw = (int)t - 1 - wsoff;
        if (is_writer && sw >= 0 && sw < (int)Lq) {
            const float2 c = ring[((t - 1) & (R - 1)) * RS + wcol];
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
            if (is_ghost) gstore(t, c);
            if (t >= 1) drain(t);
            __syncthreads();
        }
    }
    drain(steps4);
    if (is_writer && wlast) { A.rowmin[io + wnode] = rmin; A.rowarg[io + wnode] = rarg; }
}

template <bool WIDE, int PLANES>
__device__ __forceinline__ void v2_query(const MeshArgs& A, const GraphHdr& h, uint32_t ql, unsigned char* smem,
                                         const uint8_t* qt) {
    const uint32_t tid = threadIdx.x;
    const uint32_t Lq = h.qlen;
    const uint64_t io = (uint64_t)ql * A.icap;
    const uint32_t* pred_off = A.pred_off + (uint64_t)ql * (A.icap + 1);
    const uint32_t* pdesc2 = A.pdesc2 + io;
    uint32_t* tbq = A.tb + h.tb_off;
    const uint32_t sring = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t sq = (uint32_t)__cvta_generic_to_shared(qt);

    for (uint32_t g = 0; g < h.n_groups; g++) {
        const GroupInfo gi = A.groups[(uint64_t)ql * A.gcap + g];
        const uint32_t steps4 = (Lq + gi.depth - 1 + 3) & ~3u;
        if (tid >= (uint32_t)T) {
            v2_loader_group(A, smem, ql, h, gi, g, steps4);
        } else {
            const uint32_t m = A.order[((uint64_t)ql * A.gcap + g) * T + tid];
            const uint32_t rc = A.rcol[((uint64_t)ql * A.gcap + g) * T + tid];   // ring column this thread publishes to
            const bool valid = m != 0xFFFFFFFFu;
            uint32_t np = 0, pbase = 0, plane = 0, mask = 0;
            int soff = 0;
            float w = 0.f;
            float* lastcol_ptr = A.lastcol + io;  // never stored through for lanes without a row
            if (valid) {
                pbase = pred_off[m];
                np = pred_off[m + 1] - pbase;
                mask = A.nmask[io + m] & 15u;
                plane = __popc(h.maskset & ((1u << mask) - 1u));   // rank of the node's mask among the graph's
                w = A.nweight[io + m];
                soff = (int)(A.nsigma[io + m] - gi.sigma_lo);
                lastcol_ptr = A.lastcol + io + m;
            }
            const uint32_t npw = max(1u, __reduce_max_sync(0xffffffffu, np));
            const bool warp_has_rows = __any_sync(0xffffffffu, valid);
            if (valid) A.nshift[io + m] = (uint8_t)((npw - np) | (v2_raw_cells((int)npw, WIDE) ? TBR_FLAG : 0u));
            uint32_t* tbg = tbq + gi.tb_off;
            __syncthreads();  // matches the loader's prologue barrier
            if (!warp_has_rows) {
                for (uint32_t t = 0; t < steps4; t++) __syncthreads();   // a warp without rows only keeps the barriers
            } else if (npw <= (uint32_t)NPF) {
                uint32_t ck[NPF];
                const uint32_t shift = npw - np;
#pragma unroll
                for (int k = 0; k < NPF; k++) {
                    ck[k] = (S - 1) * 8u;  // rows without predecessor: the constant edge column (1, 1)
                    if (np > 0 && k < (int)npw) {
                        ck[k] = (S - 2) * 8u;   // padding slot: the constant (+inf, +inf) column, never a winner
                        if ((uint32_t)k >= shift) {
                            const uint32_t d = pdesc2[pbase + (uint32_t)k - shift];
                            ck[k] = (d & 0xffffu) * 8u + (((uint32_t)R - (d >> 16)) & (R - 1)) * SLOT_BYTES;
                        }
                    }
                }
                // lane's cell of step pair 0: halfword (u8 cells) or word (u16 cells) number tid
                uint32_t* tbl = WIDE ? tbg + tid : reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(tbg) + tid);
#define V2_CASE(N) case N: v2_fast_dispatch<N, WIDE, PLANES>(A, sring, sq, ck, rc, valid, np, soff, Lq, plane, mask, w, steps4, lastcol_ptr, tbl); break;
                switch (npw) {
                    V2_CASE(1) V2_CASE(2) V2_CASE(3) V2_CASE(4) V2_CASE(5) V2_CASE(6) V2_CASE(7)
                    default: v2_fast_dispatch<8, WIDE, PLANES>(A, sring, sq, ck, rc, valid, np, soff, Lq, plane, mask, w, steps4, lastcol_ptr, tbl); break;
                }
#undef V2_CASE
            } else {
                if (!valid) soff = (int)(steps4 + 8);    // lane without a row: s stays negative
                const float initv = np == 0 ? 1.0f : 1000000.0f;
                uint32_t* tbl = WIDE ? tbg + tid : reinterpret_cast<uint32_t*>(reinterpret_cast<uint16_t*>(tbg) + tid);
                v2_generic_group<WIDE, PLANES>(A, smem, qt, Lq, steps4, npw, np, pdesc2 + pbase, rc, soff,
                                       initv, plane, mask, __fmul_rn(A.ms, w), __fmul_rn(A.mms, w), lastcol_ptr, tbl);
            }
        }
        __syncthreads();  // ring and spill rows of this group are complete before the next group starts
    }
}

// Query table in shared memory.
// PLANES == 8 (hdr.mode 2, graphs with at most 8 distinct node masks: the four bases and a few ambiguity codes):
//   8 bytes per query position, byte p = 1 iff the query base matches the p-th IUPAC mask occurring among the
//   graph's nodes (comp(): the masks intersect, src/aligned_base.h:153-156); a row reads its match/mismatch selector
//   as one byte and turns it into the score's float bits with one integer multiply-add.
// PLANES == 1 (hdr.mode 3, more distinct masks): one byte per position = the query base's mask; rows test
//   `mask & byte`.
// Positions outside the query read 0.
template <int PLANES>
__global__ void __launch_bounds__(DP_BLOCK, DP_CTAS_PER_SM) mesh_v2_kernel(MeshArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t q = A.q0 + blockIdx.x;
    const GraphHdr h = A.hdr[q];
    if (h.status != GS_OK || h.mode != (PLANES == 8 ? 2u : 3u)) return;
    uint8_t* qt = smem + RING_BYTES + 16 + PLANES * QPAD;   // valid for s in [-QPAD, Lq + QPAD)
    const uint8_t* src = A.qmasks + A.qoff[q];
    if (PLANES == 1) {
        for (uint32_t i = threadIdx.x; i < h.qlen + 2 * QPAD; i += blockDim.x) {
            const int s = (int)i - QPAD;
            qt[s] = (s >= 0 && s < (int)h.qlen) ? (src[s] & 15u) : 0;
        }
    } else {
        uint32_t pm[8];   // mask of plane p
        uint32_t ms = h.maskset;
#pragma unroll
        for (int p2 = 0; p2 < 8; p2++) { pm[p2] = ms ? (uint32_t)__ffs((int)ms) - 1u : 0u; ms &= ms - 1u; }
        for (uint32_t i = threadIdx.x; i < h.qlen + 2 * QPAD; i += blockDim.x) {
            const int s = (int)i - QPAD;
            const uint32_t b = (s >= 0 && s < (int)h.qlen) ? (src[s] & 15u) : 0u;
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int p2 = 0; p2 < 4; p2++) { lo |= ((pm[p2] & b) ? 1u : 0u) << (8 * p2); hi |= ((pm[p2 + 4] & b) ? 1u : 0u) << (8 * p2); }
            *reinterpret_cast<uint2*>(qt + (int64_t)s * 8) = make_uint2(lo, hi);
        }
    }
    for (uint32_t i = threadIdx.x; i < RING_BYTES / 8; i += blockDim.x)  // no NaN bit patterns in unwritten cells
        reinterpret_cast<float2*>(smem)[i] = make_float2(0.f, 0.f);
    __syncthreads();
    if (h.wide) v2_query<true, PLANES>(A, h, blockIdx.x, smem, qt);
    else v2_query<false, PLANES>(A, h, blockIdx.x, smem, qt);
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
        uint32_t gmax_prev = 0;              // --insertion forbid: insertions the run at (m, s-1) may still take
        const uint32_t maxins = (valid && A.forbid) ? A.nmaxins[io + m] : 0u;
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
                float pv_cur[NPR];
                auto del_step = [&](uint32_t i, float2 c) {              // mesh.h:305-330; c.y = the predecessor's dm
                    gapm = c.y;                // last predecessor wins
                    if (c.y < value) {
                        value = c.y;
                        code = WIDE ? (TB_SRC_DEL | (i << 8)) : (TB_SRC_DEL | (i << 2));
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
                float E = edge ? 1.0f : 1000000.0f;                      // gaps_val as initialised (mesh.h:294-301)
                uint32_t gmax = 0;
                if (s > 0) {
                    bool evaluated = true;
                    if (!A.forbid) {                                     // transition_simple::insertion, mesh.h:332-358
                        E = (E_prev != H_prev) ? __fadd_rn(H_prev, gp) : __fadd_rn(E_prev, gpe);
                    } else if (maxins < 1) {                             // transition_aspace_aware, mesh.h:403-438
                        evaluated = false;
                    } else if (E_prev != H_prev) {
                        E = __fadd_rn(H_prev, gp); gmax = maxins - 1;
                    } else if (gmax_prev > 0) {
                        E = __fadd_rn(E_prev, gpe); gmax = gmax_prev - 1;
                    } else {
                        evaluated = false;
                    }
                    if (evaluated && E <= value) { value = E; code = TB_SRC_INS; }
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
                const float vgp = __fadd_rn(value, gp), ggpe = __fadd_rn(gapm, gpe);
                if (vgp < ggpe) code |= WIDE ? 4u : 32u;                 // ob: a deletion leaving this cell opens
                E_prev = E; H_prev = value; gmax_prev = gmax;
#pragma unroll
                for (int i = 0; i < NPR; i++) pv_prev[i] = pv_cur[i];
                const float2 out = make_float2(value, fminf(vgp, ggpe));
                ring[(t & (R - 1)) * S + tid] = out;
                if (sr >= 0) __stcg(&spill_w[(uint64_t)sr * Lq + s], out);
                if (s == (int)Lq - 1) A.lastcol[io + m] = value;
                if (is_last && (s == 0 || value < rmin)) { rmin = value; rarg = (uint32_t)s; }
            }
            if (tid < (uint32_t)T) {   // cells of two consecutive steps share one store (see common.cuh)
                tbw |= code << ((WIDE ? 16 : 8) * (t & 1));
                if ((t & 1) == 1) {
                    if (WIDE) tbg[(uint64_t)(t >> 1) * T + tid] = tbw;
                    else reinterpret_cast<uint16_t*>(tbg)[(uint64_t)(t >> 1) * T + tid] = (uint16_t)tbw;
                    tbw = 0;
                }
            }
            __syncthreads();
        }
        if (valid && is_last) { A.rowmin[io + m] = rmin; A.rowarg[io + m] = rarg; }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(DP_BLOCK, DP_CTAS_PER_SM) mesh_v1_kernel(MeshArgs A) {
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
    A.pdesc2 = w->d_pdesc2; A.order = w->d_order; A.rcol = w->d_rcol; A.nthr = w->d_nthr; A.nshift = w->d_nshift;
    A.ghosts = w->d_ghosts; A.writers = w->d_writers;
    A.nmaxins = w->d_nmaxins; A.forbid = ap.insertion == 1;
    A.tb = w->d_tb; A.spill = w->d_spill; A.lastcol = w->d_lastcol; A.rowmin = w->d_rowmin; A.rowarg = w->d_rowarg;
    A.ms = -ap.match_score; A.mms = -ap.mismatch_score; A.gp = ap.gap_penalty; A.gpe = ap.gap_ext_penalty;
    uint32_t max_qlen = 0;
    for (uint32_t i = q0; i < q0 + n; i++) {
        uint32_t l = (uint32_t)(s->h_qoff[i + 1] - s->h_qoff[i]);
        if (l > max_qlen) max_qlen = l;
    }
    const size_t qpos = 2 * QPAD + ((max_qlen + 15) & ~15u);
    const size_t smem8 = RING_BYTES + 16 + 8 * qpos, smem1 = RING_BYTES + 16 + qpos;  // ring + query table
    if (smem8 > 220 * 1024) SG_FAIL(SG_ERR_LIMIT, "query too long for the DP kernel's shared memory");
    SG_CUDA(cudaFuncSetAttribute(mesh_v2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
    SG_CUDA(cudaFuncSetAttribute(mesh_v2_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    SG_CUDA(cudaFuncSetAttribute(mesh_v1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem8));
    mesh_v2_kernel<8><<<n, DP_BLOCK, smem8, w->stream>>>(A);
    mesh_v2_kernel<1><<<n, DP_BLOCK, smem1, w->stream>>>(A);
    mesh_v1_kernel<<<n, DP_BLOCK, smem8, w->stream>>>(A);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 3;
    return SG_OK;
}

}  // namespace sg
