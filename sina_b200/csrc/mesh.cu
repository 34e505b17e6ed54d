// Mesh DP: query x family-graph affine-gap dynamic program with SINA's exact cell semantics.
// Replaces compute() + compute_node_simple::calc + transition_simple + scoring_scheme_simple
// (reference src/mesh.h:305-374,453-528, src/scoring_schemes.h:102-164). Scores are MINIMISED.
//
// Schedule (exact by construction, no scan): one CTA per query, one thread per graph node ("row").
// Nodes are processed in groups of DP_THREADS consecutive ids (ids are topological: column-major).
// A row whose node sits at column rank sigma computes query position s at step t = s + sigma - sigma_lo,
// so every predecessor cell (p,s) and (p,s-1) was produced at an earlier step; the row-internal insertion
// chain (m,s-1) -> (m,s) lives in the thread's registers and is evaluated with the reference's scalar
// recurrence in the reference's order (deletion, insertion, match; tie rules <, <=, <).
// Each step a row publishes (value, gapm_val) into a shared-memory ring indexed by time; predecessors in
// the same group within DP_RING-2 column ranks are read from the ring, all others from the row's global
// spill buffer (written by every row that has such a far successor).
// Traceback: one byte (or halfword when some node has more than 8 predecessors) per cell, packed over time
// and stored coalesced as tb[group][t/4][thread].
#include "common.cuh"

namespace sg {

struct MeshArgs {
    const uint8_t* qmasks; const uint64_t* qoff;
    const GraphHdr* hdr; const GroupInfo* groups; uint32_t gcap, icap, q0;
    const uint8_t* nmask; const float* nweight; const uint32_t* nsigma;
    const uint32_t* pred_off; const uint32_t* pdesc; const int32_t* spillrow; const uint8_t* nflags;
    uint32_t* tb; float2* spill;
    float* lastcol; float* rowmin; uint32_t* rowarg;
    float ms, mms, gp, gpe;  // -match_score, -mismatch_score, gap_penalty, gap_ext_penalty (align.cpp:406-407)
};

constexpr int T = DP_THREADS;
constexpr int R = DP_RING;
constexpr int NPR = 4;  // predecessors cached in registers

template <bool WIDE>
__device__ __forceinline__ void mesh_query(const MeshArgs& A, const GraphHdr& h, uint32_t ql, float2* ring,
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
        const bool valid = m < V;
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
            msw = __fmul_rn(A.ms, w);    // (comp ? match : mismatch) * weight  (scoring_schemes.h:150-156)
            mmsw = __fmul_rn(A.mms, w);
            soff = (int)(A.nsigma[io + m] - gi.sigma_lo);
            sr = A.spillrow[io + m];
            is_last = A.nflags[io + m] == 0;
#pragma unroll
            for (int i = 0; i < NPR; i++) if ((uint32_t)i < np) pd[i] = pdesc[pbase + i];
        }
        float E_prev = 1.0f, H_prev = 1.0f;  // gaps_val / value of (m, s-1)
        float rmin = 0.f;
        uint32_t rarg = 0;
        const uint32_t steps = Lq + gi.depth - 1;
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
                // ---- deletion over predecessors, ascending id (mesh.h:475-478 -> 305-330)
                auto del_step = [&](uint32_t d, uint32_t i, float2 c) {
                    (void)d;
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
                    return ring[((tt - (d >> 16)) & (R - 1)) * T + (d & 0xffffu)];
                };
#pragma unroll
                for (int i = 0; i < NPR; i++) {
                    if ((uint32_t)i < np) {
                        const float2 c = load_cell(pd[i], s, t);
                        pv_cur[i] = c.x;
                        del_step(pd[i], i, c);
                    } else pv_cur[i] = 0.f;
                }
                for (uint32_t i = NPR; i < np; i++) {
                    const uint32_t d = __ldg(&pdesc[pbase + i]);
                    del_step(d, i, load_cell(d, s, t));
                }
                float E = 1.0f;
                uint32_t ins_open = 0;
                if (s > 0) {
                    // ---- insertion from (m, s-1) (mesh.h:486-490 -> 332-358)
                    const bool ext = (E_prev == H_prev);
                    E = ext ? __fadd_rn(E_prev, gpe) : __fadd_rn(H_prev, gp);
                    ins_open = !ext;
                    if (E <= value) { value = E; code = TB_SRC_INS; }
                    // ---- match from (p, s-1) (mesh.h:492-500 -> 360-374)
                    const float sc = (mask & qm[s] & 15u) ? msw : mmsw;
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
                ring[(t & (R - 1)) * T + tid] = out;
                if (sr >= 0) __stcg(&spill_w[(uint64_t)sr * Lq + s], out);
                if (s == (int)Lq - 1) A.lastcol[io + m] = value;
                if (is_last && (s == 0 || value < rmin)) { rmin = value; rarg = (uint32_t)s; }
            }
            if (WIDE) {
                tbw |= code << (16 * (t & 1));
                if ((t & 1) == 1 || t + 1 == steps) { tbg[(uint64_t)(t >> 1) * T + tid] = tbw; tbw = 0; }
            } else {
                tbw |= code << (8 * (t & 3));
                if ((t & 3) == 3 || t + 1 == steps) { tbg[(uint64_t)(t >> 2) * T + tid] = tbw; tbw = 0; }
            }
            __syncthreads();
        }
        if (valid && is_last) { A.rowmin[io + m] = rmin; A.rowarg[io + m] = rarg; }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(DP_THREADS, 2) mesh_kernel(MeshArgs A) {
    extern __shared__ __align__(16) unsigned char smem[];
    float2* ring = reinterpret_cast<float2*>(smem);            // [R][T]
    uint8_t* qm = smem + sizeof(float2) * R * T;                // [Lq]
    const uint32_t q = A.q0 + blockIdx.x;
    const GraphHdr h = A.hdr[q];
    if (h.status != GS_OK) return;
    const uint8_t* src = A.qmasks + A.qoff[q];
    for (uint32_t i = threadIdx.x; i < h.qlen; i += blockDim.x) qm[i] = src[i];
    __syncthreads();
    if (h.wide) mesh_query<true>(A, h, blockIdx.x, ring, qm);
    else mesh_query<false>(A, h, blockIdx.x, ring, qm);
}

int launch_mesh(Session* s, const sg_align_params& ap, uint32_t q0, uint32_t n) {
    MeshArgs A;
    A.qmasks = s->d_qmasks; A.qoff = s->d_qoff; A.hdr = s->d_hdr; A.groups = s->d_groups;
    A.gcap = s->gcap; A.icap = s->icap; A.q0 = q0;
    A.nmask = s->d_nmask; A.nweight = s->d_nweight; A.nsigma = s->d_nsigma; A.pred_off = s->d_pred_off;
    A.pdesc = s->d_pdesc; A.spillrow = s->d_spillrow; A.nflags = s->d_nflags;
    A.tb = s->d_tb; A.spill = s->d_spill; A.lastcol = s->d_lastcol; A.rowmin = s->d_rowmin; A.rowarg = s->d_rowarg;
    A.ms = -ap.match_score; A.mms = -ap.mismatch_score; A.gp = ap.gap_penalty; A.gpe = ap.gap_ext_penalty;
    uint32_t max_qlen = 0;
    for (uint32_t i = q0; i < q0 + n; i++) {
        uint32_t l = (uint32_t)(s->h_qoff[i + 1] - s->h_qoff[i]);
        if (l > max_qlen) max_qlen = l;
    }
    size_t smem = sizeof(float2) * R * T + ((max_qlen + 15) & ~15u);
    if (smem > 220 * 1024) SG_FAIL(SG_ERR_LIMIT, "query too long for the DP kernel's shared memory");
    SG_CUDA(cudaFuncSetAttribute(mesh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mesh_kernel<<<n, DP_THREADS, smem, s->stream>>>(A);
    SG_CUDA(cudaGetLastError());
    s->stats.kernel_launches += 1;
    return SG_OK;
}

}  // namespace sg
