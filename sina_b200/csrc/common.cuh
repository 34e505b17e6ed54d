// Shared declarations for the sina_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include <mutex>
#include <string>

#include "../../include/sina_b200.h"
#include "find_layout.h"

namespace sg {

// ------------------------------------------------------------------ errors
void set_error(const std::string& msg);
#define SG_CUDA(call)                                                                                   \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess) {                                                                       \
            sg::set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" + \
                          std::to_string(__LINE__) + ")");                                              \
            return SG_ERR_CUDA;                                                                         \
        }                                                                                               \
    } while (0)
#define SG_FAIL(code, msg)      \
    do {                        \
        sg::set_error(msg);     \
        return code;            \
    } while (0)

#define SG_TRY(x)                       \
    do {                                \
        int rc__ = (x);                 \
        if (rc__ != SG_OK) return rc__; \
    } while (0)

// ------------------------------------------------------------------ constants
constexpr int MAX_K = 16;
constexpr uint32_t FAM_CAP_MAX = 255;        // family members per query (predecessor ordinal fits 8 bits)
constexpr uint32_t W_MAX = 1u << 20;         // alignment columns (used-column bitmap lives in shared memory)
constexpr uint32_t QLEN_MAX = 1u << 16;      // bases per query
typedef uint16_t rcol_t;                     // ring column index
// DP CTA shape: DP_THREADS threads = DP_THREADS - 32 row lanes + one loader warp; DP_CTAS resident CTAs per SM (the
// register cap the kernels are compiled for). Measured on B200 (full-length 16S queries, GCUPS of the DP kernel):
// 128 threads x 7 CTAs 481, 192 x 5 560, 224 x 5 552, 256 x 4 604, 288 x 3 574, 320 x 3 539, 512 x 2 537.
#ifndef DP_THREADS
#define DP_THREADS 256
#endif
#ifndef DP_CTAS
#define DP_CTAS (DP_THREADS >= 512 ? 2 : DP_THREADS >= 256 ? 4 : DP_THREADS >= 192 ? 5 : 7)
#endif
constexpr int DP_BLOCK = DP_THREADS;         // threads per DP CTA = columns of the shared-memory ring
constexpr int DP_CTAS_PER_SM = DP_CTAS;
constexpr int DP_T = DP_BLOCK - 32;          // node rows per DP group (compute lanes, one row each)
#ifndef GRAPH_THREADS
#define GRAPH_THREADS 384
#endif
constexpr int GRAPH_BLOCK = DP_BLOCK > GRAPH_THREADS ? DP_BLOCK : GRAPH_THREADS;             // threads of the graph kernel's CTA
constexpr int DP_G = DP_BLOCK - DP_T;        // loader lanes: ghost columns (far predecessors) + spill writers
constexpr int DP_RING = 8;                   // ring depth (time slots) of the shared-memory row window
constexpr int DP_MAXD = 4;                   // v2 kernel: largest column-rank distance of a predecessor served by the ring
                                             // (the generic kernel serves DP_RING - 2); farther ones go through ghosts
constexpr int DP_RS2 = DP_BLOCK + 1;         // v2 kernel: 16-byte cells per time slot of the ring (one bank group more than a multiple of 8)
constexpr uint32_t DP_COL_PAD = DP_BLOCK - 2, DP_COL_EDGE = DP_BLOCK - 1;   // v2 kernel: constant ring columns (mesh.cu)
constexpr int GHOST_PF = 2;                  // steps a ghost requests its data ahead of publishing it
constexpr int GHOST_LEAD = GHOST_PF + 2;     // a ghost trails a source row of its own group by >= this many column ranks
static_assert(GHOST_LEAD <= DP_MAXD, "an in-group edge too long for the ring must be long enough for a ghost");
constexpr uint32_t FARLIST_CAP = 1024;       // far edges per group the v2 plan can hold
constexpr uint32_t FAR_BIT = 0x80000000u;    // predecessor descriptor: row lives in the global spill buffer

// traceback cell layout (mesh.cu writes, backtrack.cu decodes).
// generic kernel (hdr.mode 1): one query position per step, cells of two consecutive steps share one store:
//   tb16[group][t/2][thread] (byte t&1) for u8 cells, tb32[group][t/2][thread] (half t&1) for u16 cells.
// v2 kernel (hdr.mode 2): two query positions per step, u8 cells, 2 steps (4 cells) per 32-bit store:
//   tb32[group][t/2][thread], byte 2*(t&1) + (s&1), with s = 2*(t - (sigma - sigma_lo)) + (s&1).
// Deletion candidates are computed once, by the row they leave from: a row publishes (value, dm) with
//   dm(x,s) = min(value(x,s) + gap, gapm_val(x,s) + gapext)            (deletion(), src/mesh.h:305-330, seen from src)
// and records in ITS OWN cell the bit ob(x,s) = value(x,s) + gap < gapm_val(x,s) + gapext ("a deletion leaving this
// cell opens the gap"). The reference's per-edge facts follow from it: the deletion via predecessor p opened iff
// ob(p,s); gapm_idx(x,s) = ob(lastpred(x), s) ? lastpred(x) : gapm_idx(lastpred(x), s).
constexpr uint32_t TB_SRC_NONE = 0, TB_SRC_DEL = 1, TB_SRC_INS = 2, TB_SRC_MATCH = 3;
// u8 : [1:0] src  [4:2] pred slot  [5] ob  [7] insertion opened (generic kernel only)
// u16: [1:0] src  [2] ob  [4] ins-open  [15:8] pred slot
// raw u8 (rows of v2 warps specialised on <= 3 predecessor slots): which candidates EQUAL the cell's value. The
// reference's source (the last update in its evaluation order: deletions with '<', insertion with '<=', matches
// with '<') is the first set flag in the order insertion, deletion slots, match slots; the last match slot has no
// flag (it is the source when no flag is set). A row without predecessor has no deletion / match source whatever
// the flags say (its deletion candidate is the constant that stands for the initial value). The insertion-opened
// flag is not stored: it is "the source of (m, s-1) is not an insertion" (see backtrack.cu).
//   [2:0] deletion via slot k equals  [3] insertion equals  [5:4] match via slot k equals (k < slots-1)  [7] ob
// index u8 (rows of v2 warps specialised on 4..8 slots): [4:0] index of the source = of the first candidate equal to the
// value in the order insertion (0), deletion slots (1..slots), match slots (slots+1..2*slots)  [7] ob
constexpr uint32_t TBR_DEL = 1, TBR_INS = 8, TBR_MATCH = 16, TBR_OB = 128;
constexpr uint32_t TBR_FLAG = 0x80;          // in nshift[]: the row's cells are raw
__host__ __device__ constexpr bool v2_raw_cells(int npw) { return npw <= 3; }

// ------------------------------------------------------------------ index
struct Index {
    int device = 0;
    uint32_t N = 0, W = 0;
    int k = 10, nofast = 0;
    uint32_t max_row_len = 0;
    uint64_t total_bases = 0;
    uint8_t* d_masks = nullptr;    // [total_bases]
    uint32_t* d_cols = nullptr;    // [total_bases]
    uint64_t* d_row_off = nullptr; // [N+1]
    uint32_t n_tiles = 1, tile_size = 0;   // search CTA tile = tile_warps sub-tiles
    uint32_t sub_size = SUB_DEFAULT, n_sub = 1, tile_warps = 1;  // sub-tile j = references [j*sub_size, (j+1)*sub_size)
    uint64_t n_slots = 0;          // k-mer slots: 4^(k-1) in fast mode (first base A), else 4^k
    uint32_t* d_list_off = nullptr; // [n_slots*n_sub + 1], k-mer-major: list (v, j) = postings[off[v*n_sub+j] .. off[v*n_sub+j+1])
    uint16_t* d_postings = nullptr; // reference id minus the sub-tile's first id, unordered inside a list
    uint64_t n_postings = 0;
    float* d_colw = nullptr;        // [W] positional column weights (scoring_scheme_weighted), null = none
    uint32_t* d_name_rank = nullptr; // [N] rank of the reference's name in ascending order (ties of the --search stage), null = id
    void* cached = nullptr;  // Session reused by the host-buffer entry points
    std::mutex mu;           // serialises host-buffer calls on this index
};

// per-query graph header written by the graph kernel, read by DP / backtrack / host
struct GraphHdr {
    uint32_t V, E, n_cols, n_groups;
    uint32_t n_last, n_spill, max_indeg, wide;  // wide: traceback uses u16 cells
    uint32_t mode;       // DP kernel: 2 = sorted rows + ghost columns (v2: two positions per step), 1 = generic fallback (v1)
    uint32_t maskset;    // bit b set iff some node has IUPAC mask b (1..15); the v2 kernel keeps one query match-bit plane per mask
    uint64_t tb_off;     // offset (in 4-byte words) of this query's traceback in the arena
    uint64_t spill_off;  // offset (in float2) of this query's spill rows in the arena
    uint32_t status;     // 0 ok, else SG_Q_* / internal failure code (see GS_*)
    uint32_t qlen;
};
constexpr uint32_t GS_OK = 0, GS_DONE = 50, GS_ARENA_FULL = 100, GS_LIMIT = 101;

struct GroupInfo {
    uint32_t sigma_lo, depth;  // first column rank of the group, number of column ranks it spans
    uint64_t tb_off;           // word offset inside the query's traceback block
    uint32_t n_ghost, n_writer;
};
struct GhostInfo {   // a ring column fed from a spilled row: far predecessors become near ones
    uint32_t spillrow;
    int32_t soff;    // column rank of the ghost minus the group's sigma_lo (>= -1)
};

// Chunk-local device workspace of the aligner stage (graph arrays, DP plan, traceback/spill arenas). A
// session owns several and deals the batch's chunks to them round-robin; each workspace has its own stream,
// so the graph kernel of one chunk, the DP of another and the (latency-bound) backtrack of a third overlap.
// Per-query arrays use a uniform stride (icap items / ncap columns).
constexpr int MAX_WS = 8;
struct Workspace {
    cudaStream_t stream = nullptr;   // graph, backtrack, bookkeeping (high priority when dp_stream is used)
    cudaStream_t dp_stream = nullptr; // SG_PRIO=1: the DP kernel runs on its own low-priority stream, so that the latency-bound
                                     // graph / backtrack kernels of other chunks get the SM slots a retiring DP CTA frees
    cudaStream_t bt_stream = nullptr; // SG_PRIO=2: the backtrack kernel runs on its own HIGH-priority stream: its small CTAs take the
                                     // slots retiring DP CTAs free instead of queueing behind the DP grids of the other chunks
    cudaEvent_t ev[4] = {};          // stage boundaries: graph | dp | backtrack | end
    cudaEvent_t done = nullptr;
    bool busy = false;               // a chunk is in flight (retire() has not run yet)
    uint32_t q0 = 0, n = 0;          // the chunk in flight
    uint32_t last_q0 = 0, last_n = 0; // the chunk whose arrays the workspace holds (sg_session_dump_graph)
    uint32_t prev_remaining = 0xffffffffu;
    unsigned long long* d_cursors = nullptr;  // [0] traceback arena cursor, [1] spill arena cursor
    uint32_t* d_remaining = nullptr; // queries of the chunk that did not fit the arenas in this pass
    uint32_t* h_remaining = nullptr; // pinned host copy
    uint8_t* d_tab = nullptr;        // [n][ncap][fam_cap] base mask of family row j at column rank c
    uint8_t* d_tabli = nullptr;      // [n][ncap][fam_cap] local node index in that column
    uint32_t* d_colof = nullptr;     // [n][ncap] column of rank c
    uint32_t* d_colbase = nullptr;   // [n][ncap+1] first node id of column rank c
    uint32_t* d_item_node = nullptr; // [n][icap]
    uint32_t* d_slot = nullptr;      // [n][icap] predecessor candidates grouped by node
    uint32_t* d_ncol = nullptr;      // [n][icap] node column
    uint8_t* d_nmask = nullptr;      // [n][icap]
    uint16_t* d_ncount = nullptr;    // [n][icap] family rows through the node
    float* d_nweight = nullptr;      // [n][icap]
    uint32_t* d_nsigma = nullptr;    // [n][icap] column rank
    uint32_t* d_slotbase = nullptr;  // [n][icap+1]
    uint32_t* d_cursor = nullptr;    // [n][icap]
    uint32_t* d_pred_off = nullptr;  // [n][icap+1]
    uint32_t* d_preds = nullptr;     // [n][icap]
    uint32_t* d_pdesc = nullptr;     // [n][icap] near/far descriptor per edge (generic kernel)
    uint32_t* d_pdesc2 = nullptr;    // [n][icap] (delta<<16 | ring column) per edge (v2 kernel)
    uint32_t* d_order = nullptr;     // [n][gcap*DP_T] node handled by (group, thread) in the v2 kernel
    rcol_t* d_rcol = nullptr;       // [n][gcap*DP_T] ring column (group, thread) publishes to: a permutation inside every 16-thread block
    uint16_t* d_nthr = nullptr;      // [n][icap] thread (ring column) of a node inside its group
    uint8_t* d_nshift = nullptr;     // [n][icap] predecessor-slot shift of a node (v2: slot = ordinal + shift)
    uint32_t* d_nmaxins = nullptr;   // [n][icap] --insertion forbid: free columns between a node and its nearest successor
    GhostInfo* d_ghosts = nullptr;   // [n][gcap][DP_G]
    uint32_t* d_writers = nullptr;   // [n][gcap][DP_G] node whose row a loader lane spills / min-tracks
    int32_t* d_spillrow = nullptr;   // [n][icap] spill row of node or -1
    uint8_t* d_nflags = nullptr;     // [n][icap] bit0 has successor
    uint32_t* d_lastnodes = nullptr; // [n][icap]
    GroupInfo* d_groups = nullptr;   // [n][gcap]
    float* d_lastcol = nullptr;      // [n][icap] value(m, Lq-1)
    float* d_rowmin = nullptr;       // [n][icap] min over s of value(m, s) (last nodes only)
    uint32_t* d_rowarg = nullptr;    // [n][icap] first s reaching it
    void* d_rec = nullptr;           // [n][icap] 48-byte node records of the backtrack kernel
    uint32_t* d_tb = nullptr;        // traceback arena (words)
    float2* d_spill = nullptr;       // spill arena
};

// All per-batch device state.
struct Session {
    Index* ix = nullptr;
    cudaStream_t stream = nullptr;
    uint32_t max_q = 0, nq = 0;
    uint32_t chunk = 0;              // queries per align pass: one workspace handles `chunk` queries at a time
    uint64_t max_bases = 0;
    // queries
    uint8_t* d_qmasks = nullptr;
    uint64_t* d_qoff = nullptr;
    int64_t* d_excl = nullptr;
    uint64_t* h_qoff = nullptr;  // host copy
    // find
    uint32_t find_max = 0, find_cap = 0;
    uint64_t* d_cand = nullptr;    // [nq][n_tiles][find_max] keys (score<<32 | id), unordered
    uint32_t* d_cand_n = nullptr;  // [nq][n_tiles]
    uint64_t* d_cand2 = nullptr;   // two-level merge: [nq][groups][find_max] keys of the tile groups
    uint32_t* d_cand2_n = nullptr; // [nq][groups]
    uint64_t cand2_cap = 0;        // keys d_cand2 holds
    uint64_t* d_ranked = nullptr;  // [nq][find_max] keys in rank order
    uint32_t* d_nres = nullptr;    // [nq]
    uint32_t* d_kmers = nullptr;   // [max_bases] valid (fast: A-prefixed) k-mers of query q at qoff[q].., duplicates kept
    uint32_t* d_nk = nullptr;      // [nq] how many
    unsigned long long* d_counters = nullptr;  // [8]: 0 postings, 1 cells
    // full ranking (family walk over the whole index for the queries whose window the top-k merge cannot hold)
    uint32_t full_cap = 0;           // queries the buffers below hold
    uint16_t* d_full_scores = nullptr;   // [full_cap][N]
    uint64_t* d_full_tmp = nullptr;      // [full_cap][N]
    uint64_t* d_full_keys = nullptr;     // [full_cap][N] rank order
    // family
    uint32_t fam_cap = 0;
    uint32_t* d_fam_ids = nullptr;   // [nq][fam_cap] rank order
    float* d_fam_scores = nullptr;   // [nq][fam_cap]
    int32_t* d_fam_n = nullptr;      // [nq] (-1: too few, -2: window too small)
    uint32_t* d_retry = nullptr;     // [0] queries needing a larger candidate window
    // --fs-msc-max < 1: identity of the query (at its own input positions) with every candidate of the ranked window
    uint32_t* d_qcols = nullptr;     // [max_bases] positions of the query bases (sg_session_set_query_columns)
    bool have_qcols = false;
    float* d_fpair = nullptr;        // [queries of the pass][window]
    uint64_t fpair_cap = 0;
    // --search stage (lazily allocated)
    uint32_t* d_acols = nullptr;     // [max_bases] alignment columns of the aligned queries
    float* d_pair = nullptr;         // [max_q][pair_cap] identity of (query, candidate)
    uint32_t pair_cap = 0;
    uint32_t* d_sids = nullptr;      // [max_q][sres_cap] results
    float* d_sscores = nullptr;
    uint32_t* d_sn = nullptr;        // [max_q]
    uint32_t sres_cap = 0;
    // orientation check (--turn)
    int32_t* d_turn_scores = nullptr; // [4][nq] top k-mer score of the query as is / reversed / complemented / both
    int32_t* d_turn = nullptr;        // [nq] chosen orientation 0..3
    uint8_t* d_turn_ops = nullptr;    // [nq]
    // align (whole batch)
    uint32_t icap = 0, ncap = 0, gcap = 0;  // per-query capacities: nodes (= edges; the stride of all per-node arrays), column ranks, DP groups
    uint32_t itemcap = 0;                   // items (bases of the family rows): item_node / slot of the global-scratch graph path
    uint32_t* d_afam = nullptr;      // [nq][fam_cap] family after the contains-query partition
    uint32_t* d_afam_n = nullptr;    // [nq]
    uint32_t* d_contains = nullptr;  // [nq][fam_cap] 1 + first offset of the query inside the relative, 0 = none
    uint32_t* d_copy_src = nullptr;  // [nq][2] (ref id, offset) for SG_Q_COPIED
    GraphHdr* d_hdr = nullptr;       // [nq]
    // align (chunk workspaces)
    int n_ws = 0;
    Workspace ws[MAX_WS];
    uint64_t tb_words = 0, spill_elems = 0;  // arena sizes per workspace
    cudaEvent_t ev_pre = nullptr;    // prealign finished on `stream`
    // outputs
    uint32_t* d_out_cols = nullptr;  // [max_bases]
    uint8_t* d_out_masks = nullptr;  // [max_bases]
    sg_align_result* d_results = nullptr;  // [nq]
    // streaming download (host-buffer entry points): a finished chunk's columns/bases go to a pinned staging slot on
    // `cstream` and from there to the caller's buffers while the following chunks compute
    uint32_t* host_cols = nullptr;   // destination of query 0's bases for the align call in flight (null: no streaming)
    uint8_t* host_masks = nullptr;
    cudaStream_t cstream = nullptr;
    cudaEvent_t cev[2] = {};
    uint8_t* stage[2] = {};
    uint64_t stage_cap = 0;          // bases a staging slot holds (5 bytes each: u32 column + u8 base)
    struct { bool pending; uint64_t off, nb; } stage_info[2] = {};
    uint32_t stage_next = 0;
    // timing
    cudaEvent_t ev[8] = {};
    sg_stage_stats stats = {};
    bool have_family = false, have_find = false, have_align = false;
    int pair = 0;                    // SG_PAIR=1: see enqueue_chunk
    cudaEvent_t last_dp = nullptr;   // end of the DP kernel queued last
    int force_generic = 0;           // SG_DP_GENERIC=1: run every query through the generic DP kernel (testing)
    int graph_generic = 0;           // SG_GRAPH_GENERIC=1: family graph through the global-scratch path (testing)
};

// ------------------------------------------------------------------ kernel launchers (one per .cu)
int launch_index_build(Index* ix, cudaStream_t st);
// q0 / n: query range of the batch (n == 0: all of it)
int launch_find(Session* s, uint32_t max, uint32_t q0 = 0, uint32_t n = 0);
// ranked: rank-ordered keys of the range (stride `window`); null = the session's d_ranked
int launch_family(Session* s, const sg_fam_params& fp, uint32_t window, uint32_t q0 = 0, uint32_t n = 0, const uint64_t* ranked = nullptr,
                  const float* ident = nullptr);
int launch_identity(Session* s, const uint8_t* d_amasks, const uint32_t* d_acols, const uint64_t* d_aoff, uint32_t nq,
                    const uint64_t* ranked, const uint32_t* nres, uint32_t stride, const uint32_t* pair_ids,
                    const uint64_t* pair_off, int iupac, int cover, int filter_lc, int ignore_super, float* d_scores);
int launch_search_select(Session* s, const float* d_scores, const uint64_t* ranked, const uint32_t* nres, uint32_t stride,
                         uint32_t nq, uint32_t max_result, float min_sim, uint32_t* d_ids, float* d_out, uint32_t* d_n);
int launch_find_full(Session* s, uint32_t q0, uint32_t n);   // every reference in rank order for n <= full_cap queries
int launch_turn(Session* s, int all);
int launch_prealign(Session* s, const sg_align_params& ap, uint32_t q0 = 0, uint32_t n = 0);
int launch_graph(Session* s, Workspace* w, const sg_align_params& ap, uint32_t q0, uint32_t n);
int launch_mesh(Session* s, Workspace* w, const sg_align_params& ap, uint32_t q0, uint32_t n);
int launch_backtrack(Session* s, Workspace* w, const sg_align_params& ap, uint32_t q0, uint32_t n);

// ------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t warp_id() { return threadIdx.x >> 5; }

// Block-wide exclusive prefix sum of one value per thread (blockDim.x multiple of 32, <= 1024).
// `red` is shared scratch of 33 uint32. Returns the exclusive prefix; *total = block sum.
__device__ __forceinline__ uint32_t block_exscan(uint32_t v, uint32_t* red, uint32_t* total) {
    uint32_t lane = lane_id(), w = warp_id(), nw = (blockDim.x + 31) >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    __syncthreads();  // protect red[] from a previous call
    if (lane == 31) red[w] = x;
    __syncthreads();
    if (w == 0) {
        uint32_t s = lane < nw ? red[lane] : 0, t = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= (uint32_t)o) t += y;
        }
        if (lane < nw) red[lane] = t - s;
        if (lane == 31) red[32] = t;
    }
    __syncthreads();
    if (total) *total = red[32];
    return red[w] + x - v;
}

// 2-bit code of an unambiguous base mask (A0 G1 C2 U3), reference src/aligned_base.h:113-115
__device__ __forceinline__ uint32_t base_code(uint32_t m) { return (uint32_t)__ffs((int)(m & 15u)) - 1u; }
__device__ __forceinline__ bool base_ambig(uint32_t m) { return __popc(m & 15u) > 1; }

// K-mer ending at position i of a sequence (window [i-k+1, i]); returns false if the window holds an
// ambiguous base. Shared by index build and search.
__device__ __forceinline__ bool kmer_at(const uint8_t* __restrict__ m, uint32_t i, int k, uint32_t& val) {
    uint32_t v = 0;
    bool ok = true;
    for (int j = k - 1; j >= 0; j--) {
        uint32_t b = m[i - j];
        ok &= !base_ambig(b);
        v = (v << 2) | (base_code(b) & 3u);
    }
    val = v;
    return ok;
}

#endif

}  // namespace sg
