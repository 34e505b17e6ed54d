// Mesh DP: query x family-graph affine-gap dynamic program with SINA's exact cell semantics.
// Replaces compute() + compute_node_simple::calc + transition_simple + scoring_scheme_simple
// (reference src/mesh.h:305-374,453-528, src/scoring_schemes.h:102-164). Scores are MINIMISED.
//
// Schedule (exact by construction, no scan): one CTA per query, one thread per graph node ("row").
// Nodes are processed in groups of DP_T consecutive ids (ids are topological: column-major).
// The row-internal insertion chain (m,s-1) -> (m,s) lives in the thread's registers and is evaluated with the
// reference's scalar recurrence; predecessor rows are read from a shared-memory ring indexed by time.
//
// v2 (normal path, hdr.mode == 2): a row whose node sits at column rank sigma computes the query positions
// 2j and 2j+1 at step t = j + sigma - sigma_lo: two cells per step and barrier, ring cells are float4
// (value, dm, value, dm) moved with one LDS.128 / STS.128. Rows of a group are sorted by in-degree so that warps
// are uniform and run a branch-free step specialised on the warp's in-degree, unrolled over the 8 phases of the ring
// (every shared-memory offset is an immediate). Far predecessors (other group, or more than DP_MAXD column ranks
// away) are turned into near ones by "ghost" ring columns that loader lanes stream from the global spill buffer;
// the same loader lanes spill rows that have far successors and track the row minimum of last nodes.
// v1 (fallback, hdr.mode == 1): generic per-lane loops, one query position per step, far predecessors read from
// global. Also runs --insertion forbid (transition_aspace_aware, src/mesh.h:377-438) and graphs with in-degree > 8.
//
// Traceback: one byte per cell (v1: halfword when some node has more than 8 predecessors), see common.cuh.
#include "common.cuh"

namespace sg {

struct MeshArgs {
    const uint8_t* qmasks; const uint64_t* qoff;
    const GraphHdr* hdr; const GroupInfo* groups; uint32_t gcap, icap, q0;
    const uint8_t* nmask; const float* nweight; const uint32_t* nsigma;
    const uint32_t* pred_off; const uint32_t* pdesc; const int32_t* spillrow; const uint8_t* nflags;
    const uint32_t* pdesc2; const uint32_t* order; const rcol_t* rcol; const uint16_t* nthr; uint8_t* nshift;
    const GhostInfo* ghosts; const uint32_t* writers;
    const uint32_t* nmaxins; int forbid;   // --insertion forbid (generic kernel only): free columns after each node
    uint32_t* tb; float2* spill;
    float* lastcol; float* rowmin; uint32_t* rowarg;
    float ms, mms, gp, gpe;  // -match_score, -mismatch_score, gap_penalty, gap_ext_penalty (align.cpp:406-407)
    uint32_t nw;             // v2: 32-bit words per plane of the query match-bit table
    const float* colw; uint32_t ncolw; const uint32_t* ncol;   // positional weights (scoring_scheme_weighted): generic kernel only
};

constexpr int T = DP_T;        // rows per group
constexpr int S = DP_BLOCK;    // CTA threads
constexpr int R = DP_RING;     // ring phases
constexpr int NPR = 4;         // predecessors held in registers (generic kernel)
constexpr int NPF = 8;         // largest in-degree the specialised v2 step is instantiated for
constexpr int QPAD = 256;      // v1: padding either side of the query bytes in shared memory

// ====================================================================================================
// v2 ring. A cell is float4 (value(2j), dm(2j), value(2j+1), dm(2j+1)); a time slot holds RS2 cells: columns
// 0..T-1 rows, T..T+29 ghosts, COL_PAD a constant (+inf) column read by the padding slots of rows with fewer
// predecessors than their warp is specialised on, COL_EDGE the constant column read by rows without predecessor.
// A row publishes the cell of step t (phase u = t & 7) and a consumer at column-rank distance d (1..DP_MAXD) reads
// it at step t + d. To keep every address "per-lane register + immediate" the ring is laid out over LINEAR slot
// indices: the cell of phase u lives at index u ("main", needed only when a reader wraps: u >= R - DP_MAXD) and at
// index u + R ("mirror", u <= R - 2); a reader at phase u reads index R - d + u. Indices below R - DP_MAXD are
// never read and are not allocated: NSLOT2 = R + DP_MAXD - 1 slots. The slot stride is 257 cells, i.e. one 16-byte
// bank group more than a multiple of eight, which rotates the bank groups between slots.
// ====================================================================================================
constexpr int RS2 = DP_RS2;
constexpr uint32_t SLOT2 = 16u * RS2;
constexpr int NSLOT2 = R + DP_MAXD - 1;
constexpr int LBASE = R - DP_MAXD;            // first linear slot index that exists
constexpr uint32_t RING2_BYTES = SLOT2 * NSLOT2;
constexpr uint32_t COL_PAD = DP_COL_PAD, COL_EDGE = DP_COL_EDGE;
constexpr int QB_PAD = 2 * DP_T + 64 <= 512 ? 512 : 1024;                   // zero bits either side of the query in the match-bit planes
constexpr int QB_BASE_PLANES = 15;            // planes 15..18: scratch (one plane per base bit A G C U)
constexpr int QB_PLANES = 19;
static_assert(DP_T + DP_G - 2 <= (int)COL_PAD, "ghost columns must end below the constant columns");
static_assert(2 * DP_T + 64 <= QB_PAD, "pre-start steps of a row must stay inside the padding");

template <int OFF>
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr), "n"(OFF) : "memory");
    return r;
}
template <int OFF>
__device__ __forceinline__ void sts_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared.v4.f32 [%0+%5], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "n"(OFF) : "memory");
}
// predicated store (no branch): executed iff flag != 0
template <int OFF>
__device__ __forceinline__ void sts_f4_if(uint32_t addr, float4 v, uint32_t flag) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t@p st.shared.v4.f32 [%0+%6], {%1, %2, %3, %4};\n\t}" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(flag), "n"(OFF) : "memory");
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
__device__ __forceinline__ float fset_eq(float a, float b) {
    float r;
    asm("set.eq.f32.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
// three-input minimum (FMNMX3, sm_100)
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// minimum of a[0..N) and e with FMNMX3s, e (the operand that arrives last) in the final one
template <int N>
__device__ __forceinline__ float min_with(const float (&a)[N], float e) {
    if constexpr (N == 1) return fminf(a[0], e);
    else if constexpr (N == 2) return fmin3(a[0], a[1], e);
    else {
        constexpr int M = (N + 2) / 3;
        float b[M];
#pragma unroll
        for (int j = 0; j < M; j++) {
            if (3 * j + 2 < N) b[j] = fmin3(a[3 * j], a[3 * j + 1], a[3 * j + 2]);
            else if (3 * j + 1 < N) b[j] = fminf(a[3 * j], a[3 * j + 1]);
            else b[j] = a[3 * j];
        }
        return min_with<M>(b, e);
    }
}

// Per-lane constants of the specialised step.
template <int NPW>
struct V2Lane {
    uint32_t pk[NPW];        // shared address of predecessor slot k's cell at phase 0
    uint32_t wadr;           // own ring column (linear slot LBASE)
    uint32_t qaddr;          // shared address of the row's match-bit plane
    float msw, mmsw;         // (mis)match score * weight  (scoring_schemes.h:150-156)
    uint32_t tf, tl;         // first / last step of the row (0xFFFFFFFF: lane without a row)
    int soff;                // column rank minus the group's first
    uint32_t odd;            // Lq & 1: the last query position is the first cell of the last step
    float* lastcol_ptr;
};

template <int NPW>
struct V2Addr {
    uint32_t ak[NPW];        // shared addresses of the predecessor cells read by the next pair of steps
    uint32_t aw;             // ... and of the row's own cell
};

// Two steps (four query positions) of a row from step t0 (even).
// A ring cell is (value, dm) per position: dm = min(value + gap, gapm_val + gapext) is the deletion candidate the
// cell offers to every successor row (deletion(), mesh.h:305-330, evaluated once by the row it leaves from instead
// of once per edge); the row's own gapm_val is the dm of its last predecessor ("last predecessor wins").
// Slots are right-aligned: the NPW-np leading slots of a row with np < NPW read COL_PAD (+inf: never a winner,
// never equal to a value). Rows without predecessor read COL_EDGE = (value +inf, dm 1): the deletion candidate 1
// plays the role of the initial value 1 and leaves gapm_val = 1 (init_edge, mesh.h:294-301); no match can win.
// The reference's initial value 1000000 of the other cells is never the minimum: every such cell has a
// predecessor, and value(m, s) <= 1 + s * max(gap, gapext) < 1000000 (launch_mesh sends a batch that could break
// this bound through the generic kernel). So value = min over all candidates, and the reference's winner (last
// update in its evaluation order: deletions '<', insertion '<=', matches '<') is the first candidate EQUAL to the
// minimum in the order insertion, deletions by slot, matches by slot: that is what the raw cells record
// (common.cuh), all flags against one final value so that the minimum itself is a chain of FMNMX3.
// E carries gaps_val of the row's next cell: (m,s)'s insertion candidate is fixed when (m,s-1) is finished;
// extension iff gaps_val == value (mesh.h:340-349).
// qb: the row's match bits, bit i = comp(node, query position of cell i of these two steps); four are consumed.
// EDGES: some lane of the warp may start (s == 0) or end (s == Lq-1) in these steps.
// The loop over steps is kept this short on purpose: a body unrolled over the 8 ring phases (every offset an
// immediate) measured 3x slower, the warps of a CTA run up to five differently specialised bodies and 5-13 KB
// each do not fit the instruction caches (ncu: 15 % of the stall samples "no instruction", the rest waiting at
// the barrier for the warp that had none).
template <int NPW, bool EDGES>
__device__ __forceinline__ uint32_t v2_steps2(const V2Lane<NPW>& L, const float gp, const float gpe, const uint32_t t0,
                                              const uint32_t qb, float (&pvp)[NPW], float& E, V2Addr<NPW>& AD,
                                              uint32_t* const tbp, const uint32_t prev_w) {
    const float INF = __int_as_float(0x7f800000);
    constexpr bool RAW = v2_raw_cells(NPW);
    uint32_t tbw = 0;
    // ring phase of the first step: 0, 2, 4 or 6; the second step is one slot further (an immediate, no wrap).
    // The addresses were computed during the previous pair of steps: the loads are the first instructions behind
    // the barrier, nothing else sits between the release and their issue.
    uint32_t ak[NPW];
#pragma unroll
    for (int k = 0; k < NPW; k++) ak[k] = AD.ak[k];
    const uint32_t aw = AD.aw;
    // main copy (index u, read by consumers whose phase wrapped): phases LBASE..R-1; mirror (index u + R): phases 0..R-2
    static_assert(LBASE == 4 && R == 8, "the store predicates below are written for 8 phases, main copy from phase 4 on");
    const uint32_t p_main = t0 & 4u, p_mir1 = (t0 & 6u) ^ 6u;
#pragma unroll
    for (int v = 0; v < 2; v++) {
        const uint32_t t = t0 + v;
        float4 c[NPW];
#pragma unroll
        for (int k = 0; k < NPW; k++) c[k] = v ? lds_f4<(int)SLOT2>(ak[k]) : lds_f4<0>(ak[k]);
        if (v == 0) {
            __stcs(tbp - T, prev_w);  // the previous pair's traceback word leaves in the shadow of the loads (streaming:
                                      // the traceback must not push the spill rows out of L2)
        } else {
            const uint32_t xn = ((t0 + 2u) & (R - 1)) * SLOT2;
#pragma unroll
            for (int k = 0; k < NPW; k++) AD.ak[k] = L.pk[k] + xn;
            AD.aw = L.wadr + xn;
        }
        if (EDGES) {
            // s == 0 (mesh.h:294-301,469-473): value starts from 1, no insertion, no match. E = 1 supplies that 1 and
            // makes the next insertion an extension exactly when value(m,0) == 1 (gaps_val == value)
            if (t == L.tf) {
                E = 1.0f;
#pragma unroll
                for (int k = 0; k < NPW; k++) pvp[k] = INF;
            }
        }
        const float sc0 = (qb & (1u << (2 * v))) ? L.msw : L.mmsw;       // comp(): the IUPAC masks intersect
        const float sc1 = (qb & (2u << (2 * v))) ? L.msw : L.mmsw;
        float out[4];
        // the two cells' traceback bytes as a small integer held in the low mantissa bits of a float: the sum starts at
        // 2^23 (every partial sum is an integer below 2^24, exact), so the first flag needs no separate addition
        float acc = 8388608.0f;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            float del[NPW], mt[NPW];
#pragma unroll
            for (int k = 0; k < NPW; k++) {
                del[k] = h ? c[k].w : c[k].y;                                  // deletion via slot k (mesh.h:305-330)
                mt[k] = __fadd_rn(h ? c[k].x : pvp[k], h ? sc1 : sc0);         // match via slot k   (mesh.h:360-374)
            }
            const float gmin = del[NPW - 1];                                    // last predecessor wins
            float value;
            if (RAW) {
                // minimum of all candidates, the insertion (latest operand to arrive) in the last FMNMX3
                if (NPW == 1) value = fmin3(del[0], mt[0], E);
                else if (NPW == 2) value = fmin3(fmin3(del[0], del[1], mt[0]), mt[1], E);
                else value = fmin3(fmin3(fmin3(del[0], del[1], del[2]), mt[0], mt[1]), mt[2], E);
                const float sh = h ? 256.f : 1.f;
#pragma unroll
                for (int k = 0; k < NPW; k++) acc = fmaf(fset_eq(del[k], value), (float)(TBR_DEL << k) * sh, acc);
#pragma unroll
                for (int k = 0; k + 1 < NPW; k++) acc = fmaf(fset_eq(mt[k], value), (float)(TBR_MATCH << k) * sh, acc);
            } else {
                // wider rows: the cell records the INDEX of the source, i.e. of the first candidate equal to the value
                // in the order insertion (0), deletion slots (1..NPW), match slots (NPW+1..2 NPW). key = index if equal,
                // 64 otherwise; the minimum of the keys is the index (shallow FMNMX3 trees instead of a chain of
                // compare / select pairs: these warps are the slowest of the CTA and every other warp waits for them)
                float cand[2 * NPW];
#pragma unroll
                for (int k = 0; k < NPW; k++) { cand[k] = del[k]; cand[NPW + k] = mt[k]; }
                value = min_with<2 * NPW>(cand, E);
                float key[2 * NPW];
#pragma unroll
                for (int k = 0; k < 2 * NPW; k++) key[k] = fmaf(fset_eq(cand[k], value), (float)(k + 1) - 64.f, 64.f);
                const float idx = min_with<2 * NPW>(key, fmaf(fset_eq(E, value), -64.f, 64.f));
                acc = fmaf(idx, h ? 256.f : 1.f, acc);
            }
            // ---- what this cell offers: the deletion candidate of its successors and the row's next insertion
            const float vgp = __fadd_rn(value, gp);
            const float ggpe = __fadd_rn(gmin, gpe);
            const float egpe = __fadd_rn(E, gpe);
            const bool ext = (E == value);                                      // gaps_val == value (mesh.h:340-349)
            acc = fmaf(fset_lt(vgp, ggpe), (float)TBR_OB * (h ? 256.f : 1.f), acc);
            if (RAW) acc = ext ? __fadd_rn(acc, (float)TBR_INS * (h ? 256.f : 1.f)) : acc;
            E = ext ? egpe : vgp;
            out[2 * h] = value;
            out[2 * h + 1] = fminf(vgp, ggpe);
        }
#pragma unroll
        for (int k = 0; k < NPW; k++) pvp[k] = c[k].z;
        const float4 o4 = make_float4(out[0], out[1], out[2], out[3]);
        if (v == 0) {
            sts_f4_if<-(int)(LBASE * SLOT2)>(aw, o4, p_main);
            sts_f4<(int)(DP_MAXD * SLOT2)>(aw, o4);
        } else {
            sts_f4_if<(int)SLOT2 - (int)(LBASE * SLOT2)>(aw, o4, p_main);
            sts_f4_if<(int)SLOT2 + (int)(DP_MAXD * SLOT2)>(aw, o4, p_mir1);
        }
        if (EDGES) {
            if (t == L.tl) *L.lastcol_ptr = L.odd ? out[0] : out[2];
        }
        __syncthreads();
        const uint32_t w16 = __float_as_uint(acc);
        tbw = v ? __byte_perm(tbw, w16, 0x5410) : w16;
    }
    return tbw;
}

// the row's match bits from step t (any parity) on: bit i = comp(node, query[2 * (t - soff) + i]), 32 positions
template <int NPW>
__device__ __forceinline__ uint32_t v2_qbits(const V2Lane<NPW>& L, uint32_t t) {
    const int bp = 2 * ((int)t - L.soff) + QB_PAD;
    const uint32_t wa = L.qaddr + ((uint32_t)bp >> 5) * 4u;
    return __funnelshift_r(lds_u32(wa), lds_u32(wa + 4u), (uint32_t)bp & 31u);
}

// All steps of one group for a warp whose rows have <= NPW predecessors.
// [e0, e1) = steps at which some row of the warp is inside the query; outside of them the warp only keeps the
// barriers (nothing it would publish is read: a consumer's window starts after its predecessors' and ends after
// theirs). [c0, c1) = steps at which every row of the warp is strictly inside (no start, no end): no edge tests.
template <int NPW>
__device__ __forceinline__ void v2_group(const V2Lane<NPW>& L, const float gp, const float gpe, const uint32_t steps8,
                                         const bool valid, uint32_t* tbl) {
    float pvp[NPW];
#pragma unroll
    for (int k = 0; k < NPW; k++) pvp[k] = 0.f;
    float E = 1.0f;
    const uint32_t w_first = __reduce_min_sync(0xffffffffu, L.tf);
    const uint32_t w_last = __reduce_max_sync(0xffffffffu, valid ? L.tl : 0u);
    const uint32_t i_first = __reduce_max_sync(0xffffffffu, valid ? L.tf : 0u) + 1u;   // every row has started
    const uint32_t i_last = __reduce_min_sync(0xffffffffu, L.tl);                       // first step at which a row ends
    const uint32_t e0 = min(w_first & ~1u, steps8);
    const uint32_t e1 = min((w_last & ~1u) + 2u, steps8);
    uint32_t c0 = (i_first + 1u) & ~1u, c1 = i_last & ~1u;      // pairs [t0, t0+2) with i_first <= t0 and t0 + 1 < i_last
    if (c0 >= c1 || c0 < e0 || c1 > e1) c0 = c1 = e1;
    for (uint32_t t0 = 0; t0 < e0; t0 += 2) { __syncthreads(); __syncthreads(); }
    // three phases: edge steps [e0, c0), inner steps [c0, c1), edge steps [c1, e1); the match bits are fetched for 16
    // steps at a time (32 query positions) and consumed from a register, four per pair of steps
    uint32_t* tbp = tbl + (uint64_t)(e0 >> 1) * T;
    // a pair's traceback word is stored during the next pair, one row of words back (tbp - T); the first store writes 0
    // to the word before the warp's first one: a cell before the row's first step, never read, or the pad row every
    // group's block starts with (graph.cu)
    uint32_t prev_w = 0;
    V2Addr<NPW> AD;
    {
        const uint32_t xn = (e0 & (R - 1)) * SLOT2;
#pragma unroll
        for (int k = 0; k < NPW; k++) AD.ak[k] = L.pk[k] + xn;
        AD.aw = L.wadr + xn;
    }
#pragma unroll 1
    for (int ph = 0; ph < 3; ph++) {
        uint32_t t0 = ph == 0 ? e0 : (ph == 1 ? c0 : c1);
        const uint32_t b = ph == 0 ? c0 : (ph == 1 ? c1 : e1);
        while (t0 < b) {
            const uint32_t stop = min(b, (t0 & ~15u) + 16u);
            uint32_t qb = v2_qbits<NPW>(L, t0);
            if (ph == 1) {
#pragma unroll 1
                for (; t0 < stop; t0 += 2) {
                    prev_w = v2_steps2<NPW, false>(L, gp, gpe, t0, qb, pvp, E, AD, tbp, prev_w);
                    tbp += T;
                    qb >>= 4;
                }
            } else {
#pragma unroll 1
                for (; t0 < stop; t0 += 2) {
                    prev_w = v2_steps2<NPW, true>(L, gp, gpe, t0, qb, pvp, E, AD, tbp, prev_w);
                    tbp += T;
                    qb >>= 4;
                }
            }
        }
    }
    __stcs(tbp - T, prev_w);
    for (uint32_t t0 = e1; t0 < steps8; t0 += 2) { __syncthreads(); __syncthreads(); }
}

template <int NPW>
__device__ __forceinline__ void v2_dispatch(const MeshArgs& A, uint32_t sring, uint32_t sqb, const uint32_t* ck, uint32_t rcol,
                                            bool valid, uint32_t np, int soff, uint32_t Lq, uint32_t plane, float w,
                                            uint32_t steps8, float* lastcol_ptr, uint32_t* tbl) {
    V2Lane<NPW> L;
#pragma unroll
    for (int k = 0; k < NPW; k++) L.pk[k] = sring + ck[k];
    L.wadr = sring + rcol * 16u;
    L.qaddr = sqb + plane * A.nw * 4u;
    L.msw = __fmul_rn(A.ms, w);    // (comp ? match : mismatch) * weight  (scoring_schemes.h:150-156)
    L.mmsw = __fmul_rn(A.mms, w);
    L.tf = valid ? (uint32_t)soff : 0xFFFFFFFFu;
    L.tl = valid ? (uint32_t)soff + ((Lq - 1u) >> 1) : 0xFFFFFFFFu;
    L.soff = soff;
    L.odd = Lq & 1u;
    L.lastcol_ptr = lastcol_ptr;
    v2_group<NPW>(L, A.gp, A.gpe, steps8, valid, tbl);
}

// Loader lanes (threads DP_T .. DP_BLOCK-1): lane j feeds ghost column DP_T+j from the spill buffer ahead of its
// consumers, and drains one row (spill store and/or running row minimum for last nodes).
__device__ __forceinline__ void v2_loader_group(const MeshArgs& A, unsigned char* smem, uint32_t ql, const GraphHdr& h,
                                                const GroupInfo& gi, uint32_t g, uint32_t steps8) {
    const uint32_t j = threadIdx.x - T;
    const uint32_t Lq = h.qlen, npairs = (Lq + 1u) >> 1;
    const uint64_t io = (uint64_t)ql * A.icap;
    float4* spill = reinterpret_cast<float4*>(A.spill + h.spill_off);   // rows of npairs cells
    // ghost
    const bool is_ghost = j < gi.n_ghost;
    const float4* gsrc = nullptr;
    int gsoff = 0;
    if (is_ghost) {
        const GhostInfo gh = A.ghosts[((uint64_t)ql * A.gcap + g) * DP_G + j];
        gsrc = spill + (uint64_t)gh.spillrow * npairs;
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
    const uint32_t sring = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t gaddr = sring + threadIdx.x * 16u;      // the ghost's ring column (= this thread's), linear slot LBASE
    const uint32_t waddr = sring + wcol * 16u;             // the drained row's ring column
    float4* const sdst = spill + (uint64_t)(wsr >= 0 ? wsr : 0) * npairs;
    const bool wspill = is_writer && wsr >= 0;
    const bool any_last = __any_sync(0xffffffffu, is_writer && wlast);   // uniform: groups without last nodes skip the minimum
    // This warp is alone in its role and every row warp waits for it at each step's barrier: its step is kept
    // straight-line and predicated (a first version with per-lane branches was the slowest warp of the CTA: ncu
    // showed it busy 83 % of the time against 42 % for the row warps).
    // A ghost publishes at step t the cell of the query positions 2(t - gsoff), +1 and asks for the cell of step
    // t + GHOST_PF at step t, so the L2 latency never sits between two barriers (GHOST_LEAD = GHOST_PF + 2 keeps that
    // request behind the source row's spill store). Cells outside the query are never read by a consumer.
    float4 pf[GHOST_PF];
#pragma unroll
    for (int k = 0; k < GHOST_PF; k++) pf[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    {   // prologue = step -1 (phase 7, main copy): a ghost with soff -1 must have its first cell in the ring before step 0
        const int jj = -1 - gsoff;
        if (is_ghost && jj >= 0 && jj < (int)npairs) sts_f4<(int)((R - 1 - LBASE) * SLOT2)>(gaddr, __ldcg(&gsrc[jj]));
    }
    int gj = -gsoff;                        // pair index of the ghost cell requested next
#pragma unroll
    for (int k = 0; k < GHOST_PF; k++, gj++)
        if (is_ghost && (uint32_t)gj < npairs) pf[k] = __ldcg(&gsrc[gj]);
    __syncthreads();
    int wj = -1 - wsoff;                    // pair index the drained row published at the previous step
    // what the writer's row published at step t-1: spill it (rows with far successors), track the row minimum (last nodes)
    auto drain = [&](uint32_t xd) {
        const bool on = is_writer && (uint32_t)wj < npairs;
        const float4 c = lds_f4<0>(waddr + xd);
        if (on && wspill) __stcg(&sdst[wj], c);
        if (any_last) {
            const uint32_t s0 = 2u * (uint32_t)wj;
            const bool l0 = on && wlast && (s0 == 0 || c.x < rmin);
            rmin = l0 ? c.x : rmin; rarg = l0 ? s0 : rarg;
            const bool l1 = on && wlast && s0 + 1 < Lq && c.z < rmin;
            rmin = l1 ? c.z : rmin; rarg = l1 ? s0 + 1 : rarg;
        }
        wj++;
    };
    static_assert(GHOST_PF == 2, "the prefetch queue is rotated by the two-step unrolling below");
#pragma unroll 1
    for (uint32_t t0 = 0; t0 < steps8; t0 += 2) {
        const uint32_t ph = t0 & (R - 1);                       // 0, 2, 4 or 6
        const uint32_t x0 = ph * SLOT2;
        const uint32_t gm = is_ghost ? 1u : 0u;
        const uint32_t p_main = gm & (ph >> 2), p_mir1 = gm & (ph != 6u ? 1u : 0u);
        // the cell of step t0 - 1: mirror copy of phase ph - 1, or the main copy of phase 7
        const uint32_t xd0 = ph ? x0 + (uint32_t)(DP_MAXD - 1) * SLOT2 : (uint32_t)(R - 1 - LBASE) * SLOT2;
        // ---- step t0
        sts_f4_if<-(int)(LBASE * SLOT2)>(gaddr + x0, pf[0], p_main);
        sts_f4_if<(int)(DP_MAXD * SLOT2)>(gaddr + x0, pf[0], gm);
        if (is_ghost && (uint32_t)gj < npairs) pf[0] = __ldcg(&gsrc[gj]);
        gj++;
        drain(xd0);
        __syncthreads();
        // ---- step t0 + 1
        sts_f4_if<(int)SLOT2 - (int)(LBASE * SLOT2)>(gaddr + x0, pf[1], p_main);
        sts_f4_if<(int)SLOT2 + (int)(DP_MAXD * SLOT2)>(gaddr + x0, pf[1], p_mir1);
        if (is_ghost && (uint32_t)gj < npairs) pf[1] = __ldcg(&gsrc[gj]);
        gj++;
        drain(x0 + (uint32_t)DP_MAXD * SLOT2);                  // the cell of step t0: mirror copy of phase ph
        __syncthreads();
    }
    drain((uint32_t)(R - 1 - LBASE) * SLOT2);                   // steps8 is a multiple of 8: the last step's phase is 7
    if (is_writer && wlast) { A.rowmin[io + wnode] = rmin; A.rowarg[io + wnode] = rarg; }
}

__device__ __forceinline__ void v2_query(const MeshArgs& A, const GraphHdr& h, uint32_t ql, unsigned char* smem,
                                         const uint32_t* qbt) {
    const uint32_t tid = threadIdx.x;
    const uint32_t Lq = h.qlen;
    const uint64_t io = (uint64_t)ql * A.icap;
    const uint32_t* pred_off = A.pred_off + (uint64_t)ql * (A.icap + 1);
    const uint32_t* pdesc2 = A.pdesc2 + io;
    uint32_t* tbq = A.tb + h.tb_off;
    const uint32_t sring = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t sqb = (uint32_t)__cvta_generic_to_shared(qbt);

    for (uint32_t g = 0; g < h.n_groups; g++) {
        const GroupInfo gi = A.groups[(uint64_t)ql * A.gcap + g];
        const uint32_t steps8 = (((Lq + 1u) >> 1) + gi.depth - 1u + 7u) & ~7u;
        if (tid >= (uint32_t)T) {
            v2_loader_group(A, smem, ql, h, gi, g, steps8);
        } else {
            const uint32_t m = A.order[((uint64_t)ql * A.gcap + g) * T + tid];
            const uint32_t rc = A.rcol[((uint64_t)ql * A.gcap + g) * T + tid];   // ring column this thread publishes to
            const bool valid = m != 0xFFFFFFFFu;
            uint32_t np = 0, pbase = 0, plane = 0;
            int soff = 0;
            float w = 0.f;
            float* lastcol_ptr = A.lastcol + io;  // never stored through for lanes without a row
            if (valid) {
                pbase = pred_off[m];
                np = pred_off[m + 1] - pbase;
                const uint32_t mask = A.nmask[io + m] & 15u;
                plane = __popc(h.maskset & ((1u << mask) - 1u));   // rank of the node's mask among the graph's
                w = A.nweight[io + m];
                soff = (int)(A.nsigma[io + m] - gi.sigma_lo);
                lastcol_ptr = A.lastcol + io + m;
            }
            const uint32_t npw = max(1u, __reduce_max_sync(0xffffffffu, np));
            const bool warp_has_rows = __any_sync(0xffffffffu, valid);
            if (valid) A.nshift[io + m] = (uint8_t)((npw - np) | (v2_raw_cells((int)npw) ? TBR_FLAG : 0u));
            uint32_t* tbl = tbq + gi.tb_off + tid;   // this lane's word of step pair 0
            __syncthreads();  // matches the loader's prologue barrier
            if (!warp_has_rows) {
                for (uint32_t t = 0; t < steps8; t++) __syncthreads();   // a warp without rows only keeps the barriers
            } else {
                // shared offset of predecessor slot k at phase 0: column * 16 + (DP_MAXD - distance) slots
                uint32_t ck[NPF];
                const uint32_t shift = npw - np;
#pragma unroll
                for (int k = 0; k < NPF; k++) {
                    ck[k] = COL_EDGE * 16u;     // rows without predecessor (and lanes without a row)
                    if (np > 0 && k < (int)npw) {
                        ck[k] = COL_PAD * 16u;  // padding slot
                        if ((uint32_t)k >= shift) {
                            const uint32_t d = pdesc2[pbase + (uint32_t)k - shift];
                            ck[k] = (d & 0xffffu) * 16u + ((uint32_t)DP_MAXD - (d >> 16)) * SLOT2;
                        }
                    }
                }
#define V2_CASE(N) case N: v2_dispatch<N>(A, sring, sqb, ck, rc, valid, np, soff, Lq, plane, w, steps8, lastcol_ptr, tbl); break;
                switch (npw) {
                    V2_CASE(1) V2_CASE(2) V2_CASE(3) V2_CASE(4) V2_CASE(5) V2_CASE(6) V2_CASE(7)
                    default: v2_dispatch<8>(A, sring, sqb, ck, rc, valid, np, soff, Lq, plane, w, steps8, lastcol_ptr, tbl); break;
                }
#undef V2_CASE
            }
        }
        __syncthreads();  // ring and spill rows of this group are complete before the next group starts
    }
}

// Query match-bit planes in shared memory (v2): plane p, bit QB_PAD + s = comp(p-th IUPAC mask occurring among
// the graph's nodes, query[s]) (the masks intersect, src/aligned_base.h:153-156); zero outside the query. A row
// fetches the 16 bits of a block of steps with two LDS.32 and a funnel shift and tests them from a register.
__device__ __forceinline__ void v2_build_qbits(const GraphHdr& h, const uint8_t* __restrict__ src, uint32_t* qbt, uint32_t nw) {
    const uint32_t lane = lane_id(), wid = warp_id(), nwarp = blockDim.x >> 5;
    for (uint32_t w = wid; w < nw; w += nwarp) {
        const int s = (int)(w * 32u + lane) - QB_PAD;
        const uint32_t b = (s >= 0 && s < (int)h.qlen) ? (src[s] & 15u) : 0u;
#pragma unroll
        for (int p = 0; p < 4; p++) {
            const uint32_t word = __ballot_sync(0xffffffffu, (b >> p) & 1u);
            if (lane == (uint32_t)p) qbt[(QB_BASE_PLANES + p) * nw + w] = word;
        }
    }
    __syncthreads();
    const uint32_t np = __popc(h.maskset);
    for (uint32_t i = threadIdx.x; i < np * nw; i += blockDim.x) {
        const uint32_t p = i / nw, w = i - p * nw;
        uint32_t ms = h.maskset;
        for (uint32_t r = 0; r < p; r++) ms &= ms - 1u;
        const uint32_t mask = (uint32_t)__ffs((int)ms) - 1u;     // the p-th mask of the graph
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; b++) if ((mask >> b) & 1u) word |= qbt[(QB_BASE_PLANES + b) * nw + w];
        qbt[p * nw + w] = word;
    }
}

// ====================================================================================================
// v1: generic fallback (hdr.mode == 1). Rows keep id order inside a group; far predecessors are read
// straight from the spill buffer; every row tracks its own spill / row-minimum.
// ====================================================================================================
// WEIGHTED: scoring_scheme_weighted (src/scoring_schemes.h:166-241, chosen when --filter supplies positional weights,
// src/align.cpp:409-415): gap costs and match scores are multiplied by the weight of the TARGET node's column, so a
// deletion candidate cannot be computed once by the row it leaves from: the ring / spill cells hold (value, gapm_val) and
// every edge evaluates deletion() as the reference does; the traceback cell records whether the chosen deletion opened and
// whether the last predecessor's did (the latter is what gapm_idx chains follow, see backtrack.cu).
template <bool WIDE, bool WEIGHTED>
__device__ __forceinline__ void v1_query(const MeshArgs& A, const GraphHdr& h, uint32_t ql, float2* ring,
                                         const uint8_t* qm) {
    const uint32_t tid = threadIdx.x;
    const uint32_t Lq = h.qlen, V = h.V, Lq2 = (Lq + 1u) & ~1u;   // spill rows hold an even number of positions
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
        float msw = 0.f, mmsw = 0.f, gpm = gp, gpem = gpe;
        uint32_t icol = 0, gidx_prev = 0;       // WEIGHTED: column after the node; gaps_idx of (m, s-1)
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
            if (WEIGHTED) {
                const uint32_t cm = A.ncol[io + m];
                const float wm = A.colw[min(cm, A.ncolw - 1u)];
                msw = __fmul_rn(__fmul_rn(A.ms, wm), w);      // (match * weights[col]) * node weight (scoring_schemes.h:224-232)
                mmsw = __fmul_rn(__fmul_rn(A.mms, wm), w);
                gpm = __fmul_rn(gp, wm);                      // deletions happen in the node's own column (:203-222)
                gpem = __fmul_rn(gpe, wm);
                icol = cm + 1u;                               // insertions go to the columns after it (:178-201)
            } else {
                msw = __fmul_rn(A.ms, w);
                mmsw = __fmul_rn(A.mms, w);
            }
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
                bool last_open = false;
                auto del_step = [&](uint32_t i, float2 c) {              // mesh.h:305-330; c.y = the predecessor's dm
                    if (WEIGHTED) {            // c.y = the predecessor's gapm_val: the edge is evaluated here
                        const float v = __fadd_rn(c.x, gpm), gv = __fadd_rn(c.y, gpem);
                        last_open = v < gv;
                        const float cand = last_open ? v : gv;
                        gapm = cand;           // last predecessor wins
                        if (cand < value) {
                            value = cand;
                            code = (WIDE ? (TB_SRC_DEL | (i << 8)) : (TB_SRC_DEL | (i << 2))) | (last_open ? (WIDE ? 8u : 64u) : 0u);
                        }
                        return;
                    }
                    gapm = c.y;                // last predecessor wins
                    if (c.y < value) {
                        value = c.y;
                        code = WIDE ? (TB_SRC_DEL | (i << 8)) : (TB_SRC_DEL | (i << 2));
                    }
                };
                auto load_cell = [&](uint32_t d, int ss, uint32_t tt) -> float2 {
                    if (d & FAR_BIT) return __ldcg(&spill[(uint64_t)(d & ~FAR_BIT) * Lq2 + ss]);
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
                uint32_t gmax = 0, gidx = 0;
                if (s > 0) {
                    bool evaluated = true;
                    float gpi = gp, gpei = gpe;
                    const bool opening = E_prev != H_prev;
                    if (WEIGHTED) {   // the column the inserted base will be placed in (scoring_schemes.h:178-201)
                        gpi = __fmul_rn(gp, A.colw[min(icol, A.ncolw - 1u)]);
                        gpei = __fmul_rn(gpe, A.colw[min(icol + ((uint32_t)s - 1u - gidx_prev), A.ncolw - 1u)]);
                        gidx = opening ? (uint32_t)s - 1u : gidx_prev;    // gaps_idx (mesh.h:343-348), kept as initialised (0) if not evaluated
                    }
                    if (!A.forbid) {                                     // transition_simple::insertion, mesh.h:332-358
                        E = opening ? __fadd_rn(H_prev, gpi) : __fadd_rn(E_prev, gpei);
                    } else if (maxins < 1) {                             // transition_aspace_aware, mesh.h:403-438
                        evaluated = false;
                    } else if (opening) {
                        E = __fadd_rn(H_prev, gpi); gmax = maxins - 1;
                    } else if (gmax_prev > 0) {
                        E = __fadd_rn(E_prev, gpei); gmax = gmax_prev - 1;
                    } else {
                        evaluated = false;
                    }
                    if (WEIGHTED && !evaluated) gidx = 0;
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
                if (WEIGHTED) { if (last_open) code |= WIDE ? 4u : 32u; }   // the last predecessor's deletion opened (gapm_idx chains)
                else if (vgp < ggpe) code |= WIDE ? 4u : 32u;            // ob: a deletion leaving this cell opens
                E_prev = E; H_prev = value; gmax_prev = gmax; gidx_prev = gidx;
#pragma unroll
                for (int i = 0; i < NPR; i++) pv_prev[i] = pv_cur[i];
                const float2 out = WEIGHTED ? make_float2(value, gapm) : make_float2(value, fminf(vgp, ggpe));
                ring[(t & (R - 1)) * S + tid] = out;
                if (sr >= 0) __stcg(&spill_w[(uint64_t)sr * Lq2 + s], out);
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


__global__ void __launch_bounds__(DP_BLOCK, DP_CTAS_PER_SM) mesh_kernel(MeshArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    const uint32_t q = A.q0 + blockIdx.x;
    const GraphHdr h = A.hdr[q];
    if (h.status != GS_OK) return;
    const uint8_t* src = A.qmasks + A.qoff[q];
    if (h.mode == 1) {
        float2* ring = reinterpret_cast<float2*>(smem);            // [R][S]
        uint8_t* qm = smem + RING2_BYTES + 16 + QPAD;
        for (uint32_t i = threadIdx.x; i < h.qlen; i += blockDim.x) qm[i] = src[i];
        __syncthreads();
        if (A.colw) {
            if (h.wide) v1_query<true, true>(A, h, blockIdx.x, ring, qm);
            else v1_query<false, true>(A, h, blockIdx.x, ring, qm);
        } else if (h.wide) v1_query<true, false>(A, h, blockIdx.x, ring, qm);
        else v1_query<false, false>(A, h, blockIdx.x, ring, qm);
        return;
    }
    uint32_t* qbt = reinterpret_cast<uint32_t*>(smem + RING2_BYTES + 16);
    float4* ring = reinterpret_cast<float4*>(smem);
    for (uint32_t i = threadIdx.x; i < RING2_BYTES / 16; i += blockDim.x)  // no NaN bit patterns in unwritten cells
        ring[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (threadIdx.x < (uint32_t)NSLOT2) {
        const float inf = __int_as_float(0x7f800000);
        ring[threadIdx.x * RS2 + COL_PAD] = make_float4(inf, inf, inf, inf);
        ring[threadIdx.x * RS2 + COL_EDGE] = make_float4(inf, 1.0f, inf, 1.0f);
    }
    v2_build_qbits(h, src, qbt, A.nw);
    __syncthreads();
    v2_query(A, h, blockIdx.x, smem, qbt);
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
    A.colw = s->ix->d_colw; A.ncolw = s->ix->W; A.ncol = w->d_ncol;
    uint32_t max_qlen = 0;
    for (uint32_t i = q0; i < q0 + n; i++) {
        uint32_t l = (uint32_t)(s->h_qoff[i + 1] - s->h_qoff[i]);
        if (l > max_qlen) max_qlen = l;
    }
    A.nw = (max_qlen + 2 * QB_PAD + 31) / 32 + 2;
    const size_t v2_bytes = (size_t)QB_PLANES * A.nw * 4, v1_bytes = 2 * QPAD + ((max_qlen + 15) & ~15u);
    const size_t smem = RING2_BYTES + 16 + (v2_bytes > v1_bytes ? v2_bytes : v1_bytes);  // ring + query table
    if (smem > 220 * 1024) SG_FAIL(SG_ERR_LIMIT, "query too long for the DP kernel's shared memory");
    SG_CUDA(cudaFuncSetAttribute(mesh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mesh_kernel<<<n, DP_BLOCK, smem, w->dp_stream ? w->dp_stream : w->stream>>>(A);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 1;
    return SG_OK;
}

}  // namespace sg
