// C-ABI of libsina_b200.so (see include/sina_b200.h): index / session lifetime, stage drivers, host-buffer
// wrappers. No torch types, no exceptions across the boundary, no CPU fallback.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>
#include <vector>

#include <future>

#include "common.cuh"

namespace sg {
static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

template <typename T>
static int dmalloc(T** p, uint64_t n) {
    *p = nullptr;
    SG_CUDA(cudaMalloc((void**)p, (n ? n : 1) * sizeof(T)));
    return SG_OK;
}

static uint64_t env_mb(const char* name, uint64_t dflt_mb) {
    const char* v = getenv(name);
    if (!v || !*v) return dflt_mb;
    return strtoull(v, nullptr, 10);
}

static void free_workspace(Workspace* w) {
    void* ptrs[] = {w->d_cursors, w->d_remaining, w->d_tab, w->d_tabli, w->d_colof, w->d_colbase, w->d_item_node, w->d_slot,
                    w->d_ncol, w->d_nmask, w->d_ncount, w->d_nweight, w->d_nsigma, w->d_slotbase, w->d_cursor,
                    w->d_pred_off, w->d_preds, w->d_pdesc, w->d_pdesc2, w->d_order, w->d_rcol, w->d_nthr, w->d_nshift, w->d_nmaxins, w->d_ghosts,
                    w->d_writers, w->d_spillrow, w->d_nflags, w->d_lastnodes, w->d_groups, w->d_lastcol, w->d_rowmin,
                    w->d_rowarg, w->d_rec, w->d_tb, w->d_spill};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (w->h_remaining) cudaFreeHost(w->h_remaining);
    for (auto& e : w->ev) if (e) cudaEventDestroy(e);
    if (w->done) cudaEventDestroy(w->done);
    if (w->stream) cudaStreamDestroy(w->stream);
    if (w->dp_stream) cudaStreamDestroy(w->dp_stream);
    if (w->bt_stream) cudaStreamDestroy(w->bt_stream);
    *w = Workspace{};
}

static void free_align(Session* s) {
    for (int i = 0; i < MAX_WS; i++) free_workspace(&s->ws[i]);
    s->n_ws = 0;
    void* ptrs[] = {s->d_afam, s->d_afam_n, s->d_contains, s->d_copy_src, s->d_fam_ids, s->d_fam_scores};
    for (void* p : ptrs) if (p) cudaFree(p);
    s->d_afam = nullptr; s->d_afam_n = nullptr; s->d_contains = nullptr; s->d_copy_src = nullptr;
    s->d_fam_ids = nullptr; s->d_fam_scores = nullptr;
    s->fam_cap = 0; s->icap = 0;
}

static int alloc_workspace(Session* s, Workspace* w) {
    const uint64_t C = s->chunk, I = s->icap;
    if (env_mb("SG_PRIO", 0) == 2) {
        int lo = 0, hi = 0;
        SG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        SG_CUDA(cudaStreamCreateWithPriority(&w->stream, cudaStreamNonBlocking, lo));
        SG_CUDA(cudaStreamCreateWithPriority(&w->bt_stream, cudaStreamNonBlocking, hi));
    } else if (env_mb("SG_PRIO", 0) == 1) {
        int lo = 0, hi = 0;
        SG_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = least priority (largest number)
        SG_CUDA(cudaStreamCreateWithPriority(&w->stream, cudaStreamNonBlocking, hi));
        SG_CUDA(cudaStreamCreateWithPriority(&w->dp_stream, cudaStreamNonBlocking, lo));
    } else {
        SG_CUDA(cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking));
    }
    for (auto& e : w->ev) SG_CUDA(cudaEventCreate(&e));
    SG_CUDA(cudaEventCreateWithFlags(&w->done, cudaEventDisableTiming));
    SG_CUDA(cudaHostAlloc((void**)&w->h_remaining, sizeof(uint32_t), cudaHostAllocDefault));
    SG_TRY(dmalloc(&w->d_cursors, 2)); SG_TRY(dmalloc(&w->d_remaining, 1));
    SG_TRY(dmalloc(&w->d_tab, C * s->ncap * s->fam_cap)); SG_TRY(dmalloc(&w->d_tabli, C * s->ncap * s->fam_cap));
    SG_TRY(dmalloc(&w->d_colof, C * s->ncap)); SG_TRY(dmalloc(&w->d_colbase, C * (s->ncap + 1)));
    SG_TRY(dmalloc(&w->d_item_node, C * s->itemcap)); SG_TRY(dmalloc(&w->d_slot, C * s->itemcap));
    SG_TRY(dmalloc(&w->d_ncol, C * I)); SG_TRY(dmalloc(&w->d_nmask, C * I)); SG_TRY(dmalloc(&w->d_ncount, C * I));
    SG_TRY(dmalloc(&w->d_nweight, C * I)); SG_TRY(dmalloc(&w->d_nsigma, C * I));
    SG_TRY(dmalloc(&w->d_slotbase, C * (I + 1))); SG_TRY(dmalloc(&w->d_cursor, C * I));
    SG_TRY(dmalloc(&w->d_pred_off, C * (I + 1))); SG_TRY(dmalloc(&w->d_preds, C * I));
    SG_TRY(dmalloc(&w->d_pdesc, C * I)); SG_TRY(dmalloc(&w->d_spillrow, C * I)); SG_TRY(dmalloc(&w->d_nflags, C * I));
    SG_TRY(dmalloc(&w->d_lastnodes, C * I)); SG_TRY(dmalloc(&w->d_groups, C * s->gcap));
    SG_TRY(dmalloc(&w->d_lastcol, C * I)); SG_TRY(dmalloc(&w->d_rowmin, C * I)); SG_TRY(dmalloc(&w->d_rowarg, C * I));
    SG_TRY(dmalloc(&w->d_pdesc2, C * I)); SG_TRY(dmalloc(&w->d_order, C * s->gcap * DP_T)); SG_TRY(dmalloc(&w->d_rcol, C * s->gcap * DP_T)); SG_TRY(dmalloc(&w->d_nthr, C * I));
    SG_TRY(dmalloc(&w->d_nshift, C * I)); SG_TRY(dmalloc(&w->d_nmaxins, C * I)); SG_TRY(dmalloc(&w->d_ghosts, C * s->gcap * DP_G));
    SG_TRY(dmalloc(&w->d_writers, C * s->gcap * DP_G));
    { uint8_t* p = nullptr; SG_TRY(dmalloc(&p, C * I * 48)); w->d_rec = p; }
    SG_TRY(dmalloc(&w->d_tb, s->tb_words)); SG_TRY(dmalloc(&w->d_spill, s->spill_elems));
    return SG_OK;
}

// (re)allocate everything whose size depends on the family capacity
static int ensure_family_capacity(Session* s, uint32_t fam_cap) {
    if (fam_cap <= s->fam_cap) return SG_OK;
    if (fam_cap > FAM_CAP_MAX) SG_FAIL(SG_ERR_LIMIT, "family size above 255 is not supported");
    free_align(s);
    Index* ix = s->ix;
    const uint64_t Q = s->max_q;   // per-query results of the whole batch
    const uint64_t C = s->chunk;   // graph/DP workspace: one align pass handles `chunk` queries
    s->fam_cap = fam_cap;
    // capacities per query. Items (bases of the family rows) are bounded by fam_cap x longest row, but only the
    // global-scratch graph path stores anything per item. Nodes and edges of a family graph are far fewer (a node per
    // column and character: ~1.6 per column for 40 relatives): their arrays, with one stride for all of them, are sized for
    // 8 x the longest row (SG_NODE_CAP overrides); a query whose graph needs more gets SG_Q_LIMIT. (Sizing them by the
    // item bound, 28 x more than a full-length graph uses, left no room for chunks large enough to amortise the
    // latency-bound kernels.)
    s->itemcap = (uint32_t)std::min<uint64_t>((uint64_t)fam_cap * ix->max_row_len, 0xffffff);
    s->icap = (uint32_t)std::min<uint64_t>(s->itemcap, env_mb("SG_NODE_CAP", std::max<uint64_t>(4096, 8ull * ix->max_row_len)));
    s->ncap = ix->W < s->icap ? ix->W : s->icap;
    s->gcap = s->icap / DP_T + 1;
    SG_TRY(dmalloc(&s->d_fam_ids, Q * fam_cap)); SG_TRY(dmalloc(&s->d_fam_scores, Q * fam_cap));
    SG_TRY(dmalloc(&s->d_afam, Q * fam_cap)); SG_TRY(dmalloc(&s->d_afam_n, Q));
    SG_TRY(dmalloc(&s->d_contains, Q * fam_cap)); SG_TRY(dmalloc(&s->d_copy_src, Q * 2));
    // arenas per workspace: traceback (1-2 B per DP cell) and spill rows; SG_TB_ARENA_MB / SG_SPILL_ARENA_MB override
    const uint64_t max_qlen_guess = std::max<uint64_t>(ix->max_row_len, s->max_bases / std::max<uint64_t>(1, Q));
    uint64_t tb_mb = env_mb("SG_TB_ARENA_MB", std::min<uint64_t>(32768, std::max<uint64_t>(64, C * (2 * ix->max_row_len * (max_qlen_guess + 512) / 1000000 + 1))));   // V <= ~1.5 row lengths in practice; a chunk that does not fit is re-run (retire_chunk)
    uint64_t sp_mb = env_mb("SG_SPILL_ARENA_MB", std::min<uint64_t>(16384, std::max<uint64_t>(64, C * (128 * max_qlen_guess * 8 / 1000000 + 1))));   // ~100 spilled rows per query (group boundaries); a chunk that does not fit is re-run
    s->tb_words = tb_mb * 1024 * 1024 / 4;
    s->spill_elems = sp_mb * 1024 * 1024 / 8;
    // one workspace when the batch is a single chunk, else SG_STREAMS (default 3) so that chunks overlap: measured on
    // B200 (10k full-length queries, chunks of 2368): 2 workspaces 127.4k seq/s, 3 workspaces 130.4k
    const uint64_t n_chunks = (Q + C - 1) / C;
    int want = (int)std::min<uint64_t>(std::max<uint64_t>(1, env_mb("SG_STREAMS", 3)), MAX_WS);
    if ((uint64_t)want > n_chunks) want = (int)n_chunks;
    for (int i = 0; i < want; i++) {
        const int rc = alloc_workspace(s, &s->ws[i]);
        if (rc != SG_OK) {
            // out of device memory for another workspace (two sessions of 10 k full-length queries with three workspaces
            // each fill most of 180 GB): run with the workspaces there are; without any the call fails
            free_workspace(&s->ws[i]);
            s->ws[i] = Workspace();
            cudaGetLastError();
            if (i == 0) return rc;
            break;
        }
        s->n_ws = i + 1;
    }
    return SG_OK;
}

static int stage_begin(Session* s, int st) { SG_CUDA(cudaEventRecord(s->ev[0], s->stream)); (void)st; return SG_OK; }
static int stage_end(Session* s, float* acc) {
    SG_CUDA(cudaEventRecord(s->ev[1], s->stream));
    SG_CUDA(cudaEventSynchronize(s->ev[1]));
    float ms = 0.f;
    SG_CUDA(cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]));
    *acc += ms;
    return SG_OK;
}

static int validate_align_params(const sg_align_params* ap) {
    if (!ap) SG_FAIL(SG_ERR_ARG, "align params missing");
    if (ap->insertion < 0 || ap->insertion > 2 || ap->overhang < 0 || ap->overhang > 2 || ap->lowercase < 0 || ap->lowercase > 2)
        SG_FAIL(SG_ERR_ARG, "align params: enum out of range");
    return SG_OK;
}
static int validate_fam_params(const sg_fam_params* fp) {
    if (!fp) SG_FAIL(SG_ERR_ARG, "family params missing");
    if (fp->fs_max == 0) SG_FAIL(SG_ERR_ARG, "--fs-max must be > 0");
    return SG_OK;
}
}  // namespace sg

using namespace sg;

extern "C" {

const char* sg_last_error(void) { return g_err.c_str(); }

int sg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

void sg_default_search_params(sg_search_params* p) {
    if (!p) return;
    p->kmer_candidates = 1000; p->max_result = 10; p->min_sim = 0.7f; p->ignore_super = 0;
    p->iupac = 0; p->correction = 0; p->cover = 1; p->filter_lowercase = 0;
}

void sg_default_fam_params(sg_fam_params* p) {
    p->fs_min = 40; p->fs_max = 40; p->fs_msc = 0.7f; p->fs_msc_max = 2.0f; p->fs_min_len = 150; p->fs_req_full = 1;
    p->fs_full_len = 1400; p->fs_req_gaps = 10; p->fs_req = 1; p->leave_query_out = 0;
}
void sg_default_align_params(sg_align_params* p) {
    p->match_score = 2.f; p->mismatch_score = -1.f; p->gap_penalty = 5.f; p->gap_ext_penalty = 2.f; p->fs_weight = 1.f;
    p->overhang = 0; p->lowercase = 0; p->insertion = 0; p->realign = 0;
}

// ------------------------------------------------------------------------------------------------ index
int sg_index_create(const uint8_t* masks, const uint32_t* cols, const uint64_t* row_off, uint32_t N, uint32_t W,
                    int k, int nofast, int device, sg_index** out) {
    if (!masks || !cols || !row_off || !out) SG_FAIL(SG_ERR_ARG, "sg_index_create: null argument");
    if (N == 0) SG_FAIL(SG_ERR_ARG, "sg_index_create: empty reference");
    if (k < 1 || k > MAX_K) SG_FAIL(SG_ERR_ARG, "K must be in 1..16");  // src/kmer.h:57-62
    if (W == 0 || W > 786432) SG_FAIL(SG_ERR_LIMIT, "alignment width must be in 1..786432 columns");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        SG_FAIL(SG_ERR_CUDA, "no CUDA device: sina_b200 has no CPU path");
    if (device < 0 || device >= ndev) SG_FAIL(SG_ERR_ARG, "sg_index_create: bad device ordinal");
    const uint64_t total = row_off[N];
    uint32_t max_len = 0;
    for (uint32_t i = 0; i < N; i++) {
        if (row_off[i + 1] < row_off[i]) SG_FAIL(SG_ERR_ARG, "row_off must be non-decreasing");
        const uint64_t a = row_off[i], b = row_off[i + 1];
        if (b - a > max_len) max_len = (uint32_t)(b - a);
        for (uint64_t j = a; j < b; j++) {
            if ((masks[j] & 15) == 0 || masks[j] > 31) SG_FAIL(SG_ERR_ARG, "reference base is not an IUPAC mask");
            if (cols[j] >= W || (j > a && cols[j] <= cols[j - 1]))
                SG_FAIL(SG_ERR_ARG, "reference columns must be < W and strictly increasing inside a row");
        }
    }
    Index* ix = new Index;
    ix->device = device; ix->N = N; ix->W = W; ix->k = k; ix->nofast = nofast ? 1 : 0;
    ix->max_row_len = max_len ? max_len : 1; ix->total_bases = total;
    ix->n_slots = 1ull << (2 * (nofast ? k : k - 1));
    // sub-tile size: SG_SUBTILE (power of two, 32..32768: the u16 value 0xffff is the search kernel's "no posting";
    // tests use small ones to reach the multi-tile paths),
    // doubled until the (k-mer, sub-tile) offset table stays below 2^32 entries
    uint64_t sub = env_mb("SG_SUBTILE", SUB_DEFAULT);
    if (sub < 32 || sub > SUB_MAX || (sub & (sub - 1))) { delete ix; SG_FAIL(SG_ERR_ARG, "SG_SUBTILE must be a power of two in 32..32768"); }
    while (sub < SUB_MAX && ix->n_slots * (((uint64_t)N + sub - 1) / sub) >= (1ull << 32)) sub <<= 1;
    ix->sub_size = (uint32_t)sub;
    ix->n_sub = (uint32_t)(((uint64_t)N + sub - 1) / sub);
    if (ix->n_slots * ix->n_sub >= (1ull << 32)) {
        delete ix;
        SG_FAIL(SG_ERR_LIMIT, "k-mer table too large for this k and reference size");
    }
    // sub-tiles per search CTA (find_layout.h); SG_TILE_WARPS overrides (<= 24)
    const uint64_t tw_auto = find_auto_tile_warps(ix->n_sub);
    uint64_t tw = env_mb("SG_TILE_WARPS", tw_auto);
    tw = std::max<uint64_t>(1, std::min<uint64_t>(tw, std::min<uint64_t>(TILE_WARPS_MAX, (uint64_t)TILE_WARPS_MAX * SUB_DEFAULT / sub)));
    ix->tile_warps = (uint32_t)std::min<uint64_t>(tw, ix->n_sub);
    ix->tile_size = ix->tile_warps * ix->sub_size;
    ix->n_tiles = (ix->n_sub + ix->tile_warps - 1) / ix->tile_warps;
    auto fail = [&](int rc) { sg_index_destroy((sg_index*)ix); return rc; };
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice failed"); return fail(SG_ERR_CUDA); }
    int rc;
    if ((rc = dmalloc(&ix->d_masks, total + 16)) || (rc = dmalloc(&ix->d_cols, total + 4)) ||
        (rc = dmalloc(&ix->d_row_off, (uint64_t)N + 1)))
        return fail(rc);
    if (cudaMemcpy(ix->d_masks, masks, total, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(ix->d_cols, cols, total * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(ix->d_row_off, row_off, ((uint64_t)N + 1) * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("sg_index_create: upload failed");
        return fail(SG_ERR_CUDA);
    }
    cudaStream_t st;
    if (cudaStreamCreate(&st) != cudaSuccess) { set_error("cudaStreamCreate failed"); return fail(SG_ERR_CUDA); }
    rc = launch_index_build(ix, st);
    cudaStreamDestroy(st);
    if (rc) return fail(rc);
    *out = (sg_index*)ix;
    return SG_OK;
}

void sg_index_destroy(sg_index* h) {
    Index* ix = (Index*)h;
    if (!ix) return;
    cudaSetDevice(ix->device);
    if (ix->cached) { sg_session_destroy((sg_session*)ix->cached); ix->cached = nullptr; }
    cudaFree(ix->d_masks); cudaFree(ix->d_cols); cudaFree(ix->d_row_off); cudaFree(ix->d_list_off); cudaFree(ix->d_colw);
    cudaFree(ix->d_postings); cudaFree(ix->d_name_rank);
    delete ix;
}

int sg_index_set_column_weights(sg_index* h, const float* weights, uint32_t n) {
    Index* ix = (Index*)h;
    if (!ix) SG_FAIL(SG_ERR_ARG, "null index");
    if (n != 0 && (n != ix->W || !weights)) SG_FAIL(SG_ERR_ARG, "sg_index_set_column_weights: one weight per alignment column (or n = 0 for none)");
    std::lock_guard<std::mutex> lock(ix->mu);
    SG_CUDA(cudaSetDevice(ix->device));
    SG_CUDA(cudaDeviceSynchronize());
    if (ix->d_colw) { cudaFree(ix->d_colw); ix->d_colw = nullptr; }
    if (n == 0) return SG_OK;
    for (uint32_t i = 0; i < n; i++) if (!(weights[i] == weights[i])) SG_FAIL(SG_ERR_ARG, "sg_index_set_column_weights: NaN weight");
    SG_TRY(dmalloc(&ix->d_colw, (uint64_t)n));
    SG_CUDA(cudaMemcpy(ix->d_colw, weights, (size_t)n * 4, cudaMemcpyHostToDevice));
    return SG_OK;
}

int sg_index_set_name_ranks(sg_index* h, const uint32_t* rank, uint32_t n) {
    Index* ix = (Index*)h;
    if (!ix) SG_FAIL(SG_ERR_ARG, "null index");
    if (n != 0 && (n != ix->N || !rank)) SG_FAIL(SG_ERR_ARG, "sg_index_set_name_ranks: one rank per reference (or n = 0 for none)");
    std::lock_guard<std::mutex> lock(ix->mu);
    SG_CUDA(cudaSetDevice(ix->device));
    SG_CUDA(cudaDeviceSynchronize());
    if (ix->d_name_rank) { cudaFree(ix->d_name_rank); ix->d_name_rank = nullptr; }
    if (n == 0) return SG_OK;
    SG_TRY(dmalloc(&ix->d_name_rank, (uint64_t)n));
    SG_CUDA(cudaMemcpy(ix->d_name_rank, rank, (size_t)n * 4, cudaMemcpyHostToDevice));
    return SG_OK;
}

int sg_index_info(const sg_index* h, uint32_t* N, uint32_t* W, int* k, int* nofast, uint64_t* n_postings,
                  uint32_t* n_tiles, uint32_t* tile_size) {
    const Index* ix = (const Index*)h;
    if (!ix) SG_FAIL(SG_ERR_ARG, "null index");
    if (N) *N = ix->N;
    if (W) *W = ix->W;
    if (k) *k = ix->k;
    if (nofast) *nofast = ix->nofast;
    if (n_postings) *n_postings = ix->n_postings;
    if (n_tiles) *n_tiles = ix->n_tiles;
    if (tile_size) *tile_size = ix->tile_size;
    return SG_OK;
}

int sg_index_list(const sg_index* h, uint32_t kmer, uint32_t* ids, uint64_t cap, uint64_t* n) {
    const Index* ix = (const Index*)h;
    if (!ix || !n) SG_FAIL(SG_ERR_ARG, "null argument");
    *n = 0;
    if (kmer >= ix->n_slots) return SG_OK;  // fast mode: k-mers not starting with A have no list
    SG_CUDA(cudaSetDevice(ix->device));
    std::vector<uint32_t> offs(ix->n_sub + 1);
    SG_CUDA(cudaMemcpy(offs.data(), ix->d_list_off + (uint64_t)kmer * ix->n_sub, ((size_t)ix->n_sub + 1) * 4, cudaMemcpyDeviceToHost));
    std::vector<uint16_t> loc(offs[ix->n_sub] - offs[0]);
    if (!loc.empty())
        SG_CUDA(cudaMemcpy(loc.data(), ix->d_postings + offs[0], loc.size() * 2, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> all(loc.size());
    for (uint32_t j = 0; j < ix->n_sub; j++)
        for (uint32_t e = offs[j]; e < offs[j + 1]; e++) all[e - offs[0]] = j * ix->sub_size + loc[e - offs[0]];
    std::sort(all.begin(), all.end());
    *n = all.size();
    if (ids) memcpy(ids, all.data(), std::min<uint64_t>(cap, all.size()) * 4);
    return SG_OK;
}

int sg_index_export_lists(const sg_index* h, uint64_t* list_off, uint32_t* ids) {
    const Index* ix = (const Index*)h;
    if (!ix || !list_off) SG_FAIL(SG_ERR_ARG, "sg_index_export_lists: null argument");
    SG_CUDA(cudaSetDevice(ix->device));
    const uint64_t n_off = ix->n_slots * ix->n_sub + 1;
    std::vector<uint32_t> offs(n_off);
    SG_CUDA(cudaMemcpy(offs.data(), ix->d_list_off, n_off * 4, cudaMemcpyDeviceToHost));
    for (uint64_t v = 0; v <= ix->n_slots; v++) list_off[v] = offs[v * ix->n_sub];
    if (!ids) return SG_OK;
    std::vector<uint16_t> loc(ix->n_postings);
    if (ix->n_postings) SG_CUDA(cudaMemcpy(loc.data(), ix->d_postings, ix->n_postings * 2, cudaMemcpyDeviceToHost));
    // (k-mer, sub-tile) lists are stored k-mer-major with ids local to the sub-tile and unordered inside a list: global
    // ids, ascending inside every k-mer's list
    for (uint64_t v = 0; v < ix->n_slots; v++) {
        for (uint32_t j = 0; j < ix->n_sub; j++) {
            const uint32_t a = offs[v * ix->n_sub + j], b = offs[v * ix->n_sub + j + 1];
            for (uint32_t e = a; e < b; e++) ids[e] = j * ix->sub_size + loc[e];
            std::sort(ids + a, ids + b);
        }
    }
    return SG_OK;
}

int sg_index_list_sizes(const sg_index* h, const uint32_t* kmers, uint32_t n, uint64_t* sizes) {
    const Index* ix = (const Index*)h;
    if (!ix || !kmers || !sizes) SG_FAIL(SG_ERR_ARG, "null argument");
    SG_CUDA(cudaSetDevice(ix->device));
    for (uint32_t i = 0; i < n; i++) {
        sizes[i] = 0;
        if (kmers[i] >= ix->n_slots) continue;
        uint32_t a = 0, b = 0;
        SG_CUDA(cudaMemcpy(&a, ix->d_list_off + (uint64_t)kmers[i] * ix->n_sub, 4, cudaMemcpyDeviceToHost));
        SG_CUDA(cudaMemcpy(&b, ix->d_list_off + ((uint64_t)kmers[i] + 1) * ix->n_sub, 4, cudaMemcpyDeviceToHost));
        sizes[i] = b - a;
    }
    return SG_OK;
}

// ---------------------------------------------------------------------------------------------- session
int sg_session_create(sg_index* h, uint32_t max_queries, uint64_t max_bases, sg_session** out) {
    Index* ix = (Index*)h;
    if (!ix || !out || max_queries == 0) SG_FAIL(SG_ERR_ARG, "sg_session_create: bad argument");
    SG_CUDA(cudaSetDevice(ix->device));
    Session* s = new Session;
    s->ix = ix; s->max_q = max_queries; s->max_bases = max_bases ? max_bases : 1;
    s->chunk = std::min<uint32_t>(max_queries, (uint32_t)std::max<uint64_t>(1, env_mb("SG_BATCH", 2368)));
    s->force_generic = (int)env_mb("SG_DP_GENERIC", 0);
    s->pair = (int)env_mb("SG_PAIR", 0);
    s->graph_generic = (int)env_mb("SG_GRAPH_GENERIC", 0);
    *out = (sg_session*)s;
    SG_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    for (auto& e : s->ev) SG_CUDA(cudaEventCreate(&e));
    const uint64_t Q = max_queries;
    SG_TRY(dmalloc(&s->d_qmasks, s->max_bases + 16)); SG_TRY(dmalloc(&s->d_qoff, Q + 1)); SG_TRY(dmalloc(&s->d_excl, Q));
    SG_TRY(dmalloc(&s->d_kmers, s->max_bases + 32)); SG_TRY(dmalloc(&s->d_nk, Q));
    SG_TRY(dmalloc(&s->d_cand_n, Q * ix->n_tiles)); SG_TRY(dmalloc(&s->d_nres, Q)); SG_TRY(dmalloc(&s->d_counters, 8));
    SG_TRY(dmalloc(&s->d_fam_n, Q)); SG_TRY(dmalloc(&s->d_retry, 2)); SG_TRY(dmalloc(&s->d_hdr, Q));
    SG_TRY(dmalloc(&s->d_turn_scores, 4 * Q)); SG_TRY(dmalloc(&s->d_turn, Q)); SG_TRY(dmalloc(&s->d_turn_ops, Q));
    SG_TRY(dmalloc(&s->d_out_cols, s->max_bases + 4)); SG_TRY(dmalloc(&s->d_out_masks, s->max_bases + 16));
    SG_TRY(dmalloc(&s->d_results, Q));
    SG_CUDA(cudaMemset(s->d_counters, 0, 64));
    s->h_qoff = (uint64_t*)malloc((Q + 1) * 8);
    return SG_OK;
}

void sg_session_destroy(sg_session* h) {
    Session* s = (Session*)h;
    if (!s) return;
    cudaSetDevice(s->ix->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    free_align(s);
    void* ptrs[] = {s->d_full_scores, s->d_full_tmp, s->d_full_keys, s->d_qmasks, s->d_qoff, s->d_excl, s->d_kmers, s->d_nk, s->d_cand, s->d_cand_n, s->d_cand2, s->d_cand2_n, s->d_ranked, s->d_nres, s->d_counters,
                    s->d_fam_n, s->d_retry, s->d_hdr, s->d_out_cols, s->d_out_masks, s->d_results,
                    s->d_turn_scores, s->d_turn, s->d_turn_ops, s->d_qcols, s->d_fpair, s->d_acols, s->d_pair, s->d_sids, s->d_sscores, s->d_sn};
    for (void* p : ptrs) if (p) cudaFree(p);
    for (auto& e : s->ev) if (e) cudaEventDestroy(e);
    for (auto& e : s->cev) if (e) cudaEventDestroy(e);
    for (auto& p : s->stage) if (p) cudaFreeHost(p);
    if (s->cstream) cudaStreamDestroy(s->cstream);
    if (s->stream) cudaStreamDestroy(s->stream);
    free(s->h_qoff);
    delete s;
}

int sg_session_upload(sg_session* h, const uint8_t* qmasks, const uint64_t* qoff, uint32_t nq,
                      const int64_t* exclude_ids) {
    Session* s = (Session*)h;
    if (!s || !qmasks || !qoff) SG_FAIL(SG_ERR_ARG, "sg_session_upload: null argument");
    if (nq == 0 || nq > s->max_q) SG_FAIL(SG_ERR_ARG, "sg_session_upload: query count outside 1..max_queries");
    const uint64_t base = qoff[0], total = qoff[nq] - base;
    if (total > s->max_bases) SG_FAIL(SG_ERR_ARG, "sg_session_upload: more bases than the session holds");
    for (uint32_t i = 0; i < nq; i++) {
        if (qoff[i + 1] < qoff[i]) SG_FAIL(SG_ERR_ARG, "qoff must be non-decreasing");
        const uint64_t l = qoff[i + 1] - qoff[i];
        if (l < 2 || l > QLEN_MAX) SG_FAIL(SG_ERR_ARG, "query length must be in 2..65536 bases");
        s->h_qoff[i] = qoff[i] - base;
    }
    s->h_qoff[nq] = total;
    // the bases are validated on a second host thread while this one copies them to the device (a pageable source makes
    // cudaMemcpyAsync a blocking staged copy); nothing is launched on them before the verdict is in
    const uint8_t* qb = qmasks + base;
    auto scan = [qb, total]() -> unsigned {
        unsigned bad = 0;   // branch-free so that the compiler vectorises the scan
        for (uint64_t j = 0; j < total; j++) bad |= (unsigned)((qb[j] & 15) == 0) | (unsigned)(qb[j] > 31);
        return bad;
    };
    std::future<unsigned> verdict;
    bool threaded = total > (1u << 20);
    if (threaded) {
        try { verdict = std::async(std::launch::async, scan); } catch (...) { threaded = false; }   // no thread: scan here
    }
    if (!threaded && scan()) SG_FAIL(SG_ERR_ARG, "query base is not an IUPAC mask");
    SG_CUDA(cudaSetDevice(s->ix->device));
    s->nq = 0;
    SG_CUDA(cudaMemcpyAsync(s->d_qmasks, qmasks + base, total, cudaMemcpyHostToDevice, s->stream));
    if (threaded) {
        unsigned bad = 1;
        try { bad = verdict.get(); } catch (...) { bad = scan(); }
        if (bad) SG_FAIL(SG_ERR_ARG, "query base is not an IUPAC mask");
    }
    s->nq = nq;
    SG_CUDA(cudaMemcpyAsync(s->d_qoff, s->h_qoff, ((uint64_t)nq + 1) * 8, cudaMemcpyHostToDevice, s->stream));
    if (exclude_ids) SG_CUDA(cudaMemcpyAsync(s->d_excl, exclude_ids, (uint64_t)nq * 8, cudaMemcpyHostToDevice, s->stream));
    else SG_CUDA(cudaMemsetAsync(s->d_excl, 0xff, (uint64_t)nq * 8, s->stream));
    s->have_find = s->have_family = s->have_align = false;
    s->have_qcols = false;
    return SG_OK;
}

int sg_session_set_query_columns(sg_session* h, const uint32_t* qcols) {
    Session* s = (Session*)h;
    if (!s || s->nq == 0 || !qcols) SG_FAIL(SG_ERR_ARG, "sg_session_set_query_columns: no queries uploaded");
    SG_CUDA(cudaSetDevice(s->ix->device));
    const uint64_t total = s->h_qoff[s->nq];
    for (uint32_t q = 0; q < s->nq; q++)
        for (uint64_t j = s->h_qoff[q] + 1; j < s->h_qoff[q + 1]; j++)
            if (qcols[j] <= qcols[j - 1]) SG_FAIL(SG_ERR_ARG, "query positions must be strictly increasing");
    if (!s->d_qcols) SG_TRY(dmalloc(&s->d_qcols, s->max_bases));
    SG_CUDA(cudaMemcpyAsync(s->d_qcols, qcols, total * 4, cudaMemcpyHostToDevice, s->stream));
    SG_CUDA(cudaStreamSynchronize(s->stream));
    s->have_qcols = true;
    return SG_OK;
}

int sg_session_find(sg_session* h, uint32_t max) {
    Session* s = (Session*)h;
    if (!s || s->nq == 0) SG_FAIL(SG_ERR_ARG, "sg_session_find: no queries uploaded");
    SG_CUDA(cudaSetDevice(s->ix->device));
    SG_TRY(stage_begin(s, 0));
    SG_TRY(launch_find(s, max));
    SG_TRY(stage_end(s, &s->stats.ms_find));
    s->have_find = true;
    return SG_OK;
}

int sg_session_turn(sg_session* h, int mode, int32_t* turn) {
    Session* s = (Session*)h;
    if (!s || s->nq == 0) SG_FAIL(SG_ERR_ARG, "sg_session_turn: no queries uploaded");
    if (mode < 0 || mode > 2) SG_FAIL(SG_ERR_ARG, "sg_session_turn: mode must be 0 (none), 1 (revcomp) or 2 (all)");
    SG_CUDA(cudaSetDevice(s->ix->device));
    if (mode == 0) {
        if (turn) memset(turn, 0, (size_t)s->nq * 4);
        return SG_OK;
    }
    SG_TRY(stage_begin(s, 0));
    SG_TRY(launch_turn(s, mode == 2));
    SG_TRY(stage_end(s, &s->stats.ms_find));
    if (turn) {
        SG_CUDA(cudaMemcpyAsync(turn, s->d_turn, (uint64_t)s->nq * 4, cudaMemcpyDeviceToHost, s->stream));
        SG_CUDA(cudaStreamSynchronize(s->stream));
    }
    s->have_find = s->have_family = s->have_align = false;   // the query buffer changed
    return SG_OK;
}

// Family finding for queries [q0, q0 + n) of the batch (n == 0: all of it): k-mer search + selection with the
// reference's first window fs_max + 1; the queries whose quotas that window does not meet (famfinder.cpp:591-608 widens
// it x10 until it covers the index) are re-ranked alone, run by run: with the next windows while the shared-memory
// merge holds them, then over the whole index (rank_full_kernel). A wider window never changes the result of the
// walk, so the outcome equals the reference's loop.
static int family_range(Session* s, const sg_fam_params* fp, uint32_t q0, uint32_t n) {
    Index* ix = s->ix;
    if (n == 0) { q0 = 0; n = s->nq; }
    uint64_t window = (uint64_t)fp->fs_max + 1;  // famfinder.cpp:590
    auto fits_merge = [&](uint64_t w) { return find_merge_plan(w, ix->n_tiles, nullptr, nullptr); };
    // remove_similar (famfinder.cpp:553-556): cseq_comparator(optimistic, none, query cover, no filter) of the query at its
    // own input positions against the candidate; identities are <= 1, so the test only bites below 1
    const bool similar = fp->fs_msc_max < 1.0f;
    if (similar && !s->have_qcols) SG_FAIL(SG_ERR_ARG, "--fs-msc-max < 1 compares positions: give the queries' columns (sg_session_set_query_columns / sg_family_batch_aligned)");
    auto ident_buffer = [&](uint64_t need) -> int {
        if (need <= s->fpair_cap) return SG_OK;
        if (s->d_fpair) { cudaFree(s->d_fpair); s->d_fpair = nullptr; s->fpair_cap = 0; }
        SG_TRY(dmalloc(&s->d_fpair, need));
        s->fpair_cap = need;
        return SG_OK;
    };
    auto pass = [&](uint32_t w, uint32_t a, uint32_t cnt, uint32_t* retry) -> int {
        SG_TRY(stage_begin(s, 0));
        SG_TRY(launch_find(s, w, a, cnt));
        SG_TRY(stage_end(s, &s->stats.ms_find));
        SG_TRY(stage_begin(s, 1));
        const float* ident = nullptr;
        if (similar) {
            SG_TRY(ident_buffer((uint64_t)cnt * s->find_max));
            SG_TRY(launch_identity(s, s->d_qmasks, s->d_qcols, s->d_qoff + a, cnt, s->d_ranked + (uint64_t)a * s->find_max, s->d_nres + a,
                                   s->find_max, nullptr, nullptr, 0, 1, 0, 0, s->d_fpair));
            ident = s->d_fpair;
        }
        SG_TRY(launch_family(s, *fp, s->find_max, a, cnt, nullptr, ident));
        SG_TRY(stage_end(s, &s->stats.ms_family));
        SG_CUDA(cudaMemcpyAsync(retry, s->d_retry, 4, cudaMemcpyDeviceToHost, s->stream));
        SG_CUDA(cudaStreamSynchronize(s->stream));
        return SG_OK;
    };
    uint32_t retry = 0;
    const uint32_t w0 = (uint32_t)std::min<uint64_t>(window, ix->N);
    if (!fits_merge(w0)) SG_FAIL(SG_ERR_LIMIT, "--fs-max too large for this reference size (first window exceeds the top-k merge)");
    SG_TRY(pass(w0, q0, n, &retry));
    if (retry == 0 || w0 >= ix->N) return SG_OK;
    // runs of consecutive queries still asking for a wider window
    std::vector<int32_t> fam_n(n);
    auto flagged_runs = [&](std::vector<std::pair<uint32_t, uint32_t>>& runs) -> int {
        SG_CUDA(cudaMemcpyAsync(fam_n.data(), s->d_fam_n + q0, (size_t)n * 4, cudaMemcpyDeviceToHost, s->stream));
        SG_CUDA(cudaStreamSynchronize(s->stream));
        runs.clear();
        for (uint32_t i = 0; i < n;) {
            if (fam_n[i] != -2) { i++; continue; }
            uint32_t j = i;
            while (j < n && fam_n[j] == -2) j++;
            runs.emplace_back(q0 + i, j - i);
            i = j;
        }
        return SG_OK;
    };
    std::vector<std::pair<uint32_t, uint32_t>> runs;
    for (;;) {
        window *= 10;  // famfinder.cpp:607
        const uint32_t w = (uint32_t)std::min<uint64_t>(window, ix->N);
        if (!fits_merge(w)) break;
        SG_TRY(flagged_runs(runs));
        uint32_t left = 0;
        for (auto& r : runs) { uint32_t rr = 0; SG_TRY(pass(w, r.first, r.second, &rr)); left += rr; }
        if (left == 0 || w >= ix->N) return SG_OK;
    }
    // the rest walks the whole index
    SG_TRY(flagged_runs(runs));
    const uint64_t per_q = (uint64_t)ix->N * (2 + 8 + 8);
    const uint32_t cap = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(256, (512ull << 20) / per_q));
    if (cap > s->full_cap) {
        if (s->d_full_scores) cudaFree(s->d_full_scores);
        if (s->d_full_tmp) cudaFree(s->d_full_tmp);
        if (s->d_full_keys) cudaFree(s->d_full_keys);
        s->d_full_scores = nullptr; s->d_full_tmp = nullptr; s->d_full_keys = nullptr; s->full_cap = 0;
        SG_TRY(dmalloc(&s->d_full_scores, (uint64_t)cap * ix->N)); SG_TRY(dmalloc(&s->d_full_tmp, (uint64_t)cap * ix->N));
        SG_TRY(dmalloc(&s->d_full_keys, (uint64_t)cap * ix->N));
        s->full_cap = cap;
    }
    for (auto& r : runs) {
        for (uint32_t a = r.first; a < r.first + r.second; a += s->full_cap) {
            const uint32_t cnt = std::min(s->full_cap, r.first + r.second - a);
            SG_TRY(stage_begin(s, 0));
            SG_TRY(launch_find_full(s, a, cnt));
            SG_TRY(stage_end(s, &s->stats.ms_find));
            SG_TRY(stage_begin(s, 1));
            const float* ident = nullptr;
            if (similar) {
                SG_TRY(ident_buffer((uint64_t)cnt * ix->N));
                SG_TRY(launch_identity(s, s->d_qmasks, s->d_qcols, s->d_qoff + a, cnt, s->d_full_keys, s->d_nres + a, ix->N, nullptr, nullptr,
                                       0, 1, 0, 0, s->d_fpair));
                ident = s->d_fpair;
            }
            SG_TRY(launch_family(s, *fp, ix->N, a, cnt, s->d_full_keys, ident));
            SG_TRY(stage_end(s, &s->stats.ms_family));
        }
    }
    return SG_OK;
}

int sg_session_family(sg_session* h, const sg_fam_params* fp) {
    Session* s = (Session*)h;
    if (!s || s->nq == 0) SG_FAIL(SG_ERR_ARG, "sg_session_family: no queries uploaded");
    SG_TRY(validate_fam_params(fp));
    SG_CUDA(cudaSetDevice(s->ix->device));
    // the quota only starts removing items once fs_min are kept (famfinder.cpp:558-586): with --fs-min > --fs-max the family
    // grows to fs_min members
    SG_TRY(ensure_family_capacity(s, std::max(fp->fs_min, fp->fs_max) + fp->fs_req_full + 1));
    SG_TRY(family_range(s, fp, 0, 0));
    s->have_find = true;
    s->have_family = true;
    return SG_OK;
}

int sg_session_set_family(sg_session* h, const uint32_t* fam_ids, const uint64_t* fam_off) {
    Session* s = (Session*)h;
    if (!s || s->nq == 0 || !fam_ids || !fam_off) SG_FAIL(SG_ERR_ARG, "sg_session_set_family: bad argument");
    uint32_t cap = 1;
    for (uint32_t i = 0; i < s->nq; i++) {
        if (fam_off[i + 1] < fam_off[i]) SG_FAIL(SG_ERR_ARG, "fam_off must be non-decreasing");
        cap = std::max<uint32_t>(cap, (uint32_t)(fam_off[i + 1] - fam_off[i]));
    }
    for (uint64_t j = fam_off[0]; j < fam_off[s->nq]; j++)
        if (fam_ids[j] >= s->ix->N) SG_FAIL(SG_ERR_ARG, "family id outside the index");
    SG_CUDA(cudaSetDevice(s->ix->device));
    SG_TRY(ensure_family_capacity(s, cap));
    std::vector<uint32_t> ids((uint64_t)s->nq * s->fam_cap, 0);
    std::vector<int32_t> n(s->nq);
    for (uint32_t i = 0; i < s->nq; i++) {
        n[i] = (int32_t)(fam_off[i + 1] - fam_off[i]);
        memcpy(&ids[(uint64_t)i * s->fam_cap], fam_ids + fam_off[i], (size_t)n[i] * 4);
    }
    SG_CUDA(cudaMemcpyAsync(s->d_fam_ids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, s->stream));
    SG_CUDA(cudaMemcpyAsync(s->d_fam_n, n.data(), n.size() * 4, cudaMemcpyHostToDevice, s->stream));
    SG_CUDA(cudaMemsetAsync(s->d_fam_scores, 0, (uint64_t)s->nq * s->fam_cap * 4, s->stream));
    SG_CUDA(cudaStreamSynchronize(s->stream));
    s->have_family = true;
    return SG_OK;
}

// Streaming download: copy slot `slot` of the staging area into the caller's buffers once its D2H has landed.
static int flush_stage(Session* s, int slot) {
    auto& si = s->stage_info[slot];
    if (!si.pending) return SG_OK;
    SG_CUDA(cudaEventSynchronize(s->cev[slot]));
    if (s->host_cols) memcpy(s->host_cols + si.off, s->stage[slot], si.nb * 4);
    if (s->host_masks) memcpy(s->host_masks + si.off, s->stage[slot] + s->stage_cap * 4, si.nb);
    si.pending = false;
    return SG_OK;
}

// Queue the D2H of a finished chunk's output (queries q0 .. q0+n) into the next staging slot.
static int stage_chunk(Session* s, uint32_t q0, uint32_t n) {
    if (!s->host_cols && !s->host_masks) return SG_OK;
    const uint64_t off = s->h_qoff[q0], nb = s->h_qoff[q0 + n] - off;
    if (nb == 0) return SG_OK;
    if (!s->cstream) {
        SG_CUDA(cudaStreamCreateWithFlags(&s->cstream, cudaStreamNonBlocking));
        for (auto& e : s->cev) SG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (nb > s->stage_cap) {
        for (int i = 0; i < 2; i++) {
            SG_TRY(flush_stage(s, i));
            if (s->stage[i]) { cudaFreeHost(s->stage[i]); s->stage[i] = nullptr; }
        }
        s->stage_cap = nb + nb / 4 + 1024;
        for (int i = 0; i < 2; i++) SG_CUDA(cudaMallocHost((void**)&s->stage[i], s->stage_cap * 5));
    }
    const int slot = (int)(s->stage_next++ & 1u);
    SG_TRY(flush_stage(s, slot));
    if (s->host_cols) SG_CUDA(cudaMemcpyAsync(s->stage[slot], s->d_out_cols + off, nb * 4, cudaMemcpyDeviceToHost, s->cstream));
    if (s->host_masks) SG_CUDA(cudaMemcpyAsync(s->stage[slot] + s->stage_cap * 4, s->d_out_masks + off, nb, cudaMemcpyDeviceToHost, s->cstream));
    SG_CUDA(cudaEventRecord(s->cev[slot], s->cstream));
    s->stage_info[slot].pending = true; s->stage_info[slot].off = off; s->stage_info[slot].nb = nb;
    return SG_OK;
}

// Retire the chunk in flight on a workspace: wait for it, add its stage times, and if some of its queries
// did not fit the traceback/spill arenas re-run those (arenas reset) until none is left.
static int retire_chunk(Session* s, Workspace* w, const sg_align_params& ap);

static int enqueue_chunk(Session* s, Workspace* w, const sg_align_params& ap, uint32_t q0, uint32_t n) {
    w->q0 = q0; w->n = n; w->busy = true; w->last_q0 = q0; w->last_n = n;
    SG_CUDA(cudaMemsetAsync(w->d_cursors, 0, 16, w->stream));  // arena cursors
    SG_CUDA(cudaMemsetAsync(w->d_remaining, 0, 4, w->stream));
    // SG_PAIR=1: the graph kernel of this chunk starts when the DP kernel of the previous chunk (other workspace) ends,
    // i.e. together with that chunk's backtrack kernel: the two latency-bound kernels share the GPU between two DP kernels
    // instead of each trickling in beside one
    if (s->pair && s->last_dp) SG_CUDA(cudaStreamWaitEvent(w->stream, s->last_dp, 0));
    SG_CUDA(cudaEventRecord(w->ev[0], w->stream));
    SG_TRY(launch_graph(s, w, ap, q0, n));
    SG_CUDA(cudaEventRecord(w->ev[1], w->stream));
    if (w->dp_stream) SG_CUDA(cudaStreamWaitEvent(w->dp_stream, w->ev[1], 0));
    SG_TRY(launch_mesh(s, w, ap, q0, n));
    SG_CUDA(cudaEventRecord(w->ev[2], w->dp_stream ? w->dp_stream : w->stream));
    s->last_dp = w->ev[2];
    if (w->dp_stream) SG_CUDA(cudaStreamWaitEvent(w->stream, w->ev[2], 0));
    cudaStream_t bs = w->bt_stream ? w->bt_stream : w->stream;
    if (w->bt_stream) SG_CUDA(cudaStreamWaitEvent(bs, w->ev[2], 0));
    SG_TRY(launch_backtrack(s, w, ap, q0, n));
    SG_CUDA(cudaEventRecord(w->ev[3], bs));
    SG_CUDA(cudaMemcpyAsync(w->h_remaining, w->d_remaining, 4, cudaMemcpyDeviceToHost, bs));
    SG_CUDA(cudaEventRecord(w->done, bs));
    return SG_OK;
}

static int retire_chunk(Session* s, Workspace* w, const sg_align_params& ap) {
    while (w->busy) {
        SG_CUDA(cudaEventSynchronize(w->done));
        float ms = 0.f;
        SG_CUDA(cudaEventElapsedTime(&ms, w->ev[0], w->ev[1])); s->stats.ms_graph += ms;
        SG_CUDA(cudaEventElapsedTime(&ms, w->ev[1], w->ev[2])); s->stats.ms_dp += ms;
        SG_CUDA(cudaEventElapsedTime(&ms, w->ev[2], w->ev[3])); s->stats.ms_backtrack += ms;
        if (getenv("SG_TRACE")) {   // timeline of the chunk pipeline: stage boundaries of every chunk relative to the align call's start
            float t[4];
            for (int i = 0; i < 4; i++) cudaEventElapsedTime(&t[i], s->ev[0], w->ev[i]);
            fprintf(stderr, "chunk q0=%u n=%u ws=%d graph %.3f dp %.3f backtrack %.3f end %.3f ms\n", w->q0, w->n, (int)(w - s->ws), t[0], t[1], t[2], t[3]);
        }
        w->busy = false;
        const uint32_t remaining = *w->h_remaining;
        if (remaining == 0) { w->prev_remaining = 0xffffffffu; SG_TRY(stage_chunk(s, w->q0, w->n)); break; }
        if (remaining >= w->prev_remaining)
            SG_FAIL(SG_ERR_LIMIT, "align stage made no progress (internal error: every query left fits an empty arena)");
        w->prev_remaining = remaining;
        SG_TRY(enqueue_chunk(s, w, ap, w->q0, w->n));  // finished queries are skipped (GS_DONE), the rest redone
    }
    return SG_OK;
}

// Chunk pipeline of the aligner stage from query q_start on: chunk k goes to workspace k % n_ws (k_start = chunks
// already queued), then everything in flight is retired.
static int align_chunks(Session* s, const sg_align_params* ap, uint32_t q_start, int k) {
    // the first chunk is a third of the others: its graph kernel is all the GPU has to do until the first DP kernel can
    // start, so a short one shortens the pipeline's fill
    for (uint32_t q0 = q_start, n; q0 < s->nq; q0 += n, k++) {
        Workspace* w = &s->ws[k % s->n_ws];
        static const uint32_t first_div = (uint32_t)std::max<uint64_t>(1, env_mb("SG_FIRST_DIV", 3));
        n = std::min((k == 0 && s->n_ws > 1) ? std::max(1u, s->chunk / first_div) : s->chunk, s->nq - q0);
        SG_TRY(retire_chunk(s, w, *ap));
        SG_TRY(enqueue_chunk(s, w, *ap, q0, n));
        // the retired chunk's output goes to the caller's buffers while the GPU runs the chunks just queued
        for (int i = 0; i < 2; i++) SG_TRY(flush_stage(s, i));
    }
    for (int i = 0; i < s->n_ws; i++) {
        SG_TRY(retire_chunk(s, &s->ws[(k + i) % s->n_ws], *ap));   // oldest chunk first
        if (i + 1 < s->n_ws) for (int j = 0; j < 2; j++) SG_TRY(flush_stage(s, j));
    }
    for (int i = 0; i < 2; i++) SG_TRY(flush_stage(s, i));
    unsigned long long cnt[2];
    SG_CUDA(cudaMemcpyAsync(cnt, s->d_counters, 16, cudaMemcpyDeviceToHost, s->stream));
    SG_CUDA(cudaStreamSynchronize(s->stream));
    s->stats.postings = cnt[0];
    s->stats.cells = cnt[1];
    s->have_align = true;
    return SG_OK;
}

// Bring the chunk pipeline back to idle: after an error exit of the align stage some workspaces still carry a chunk
// "in flight" (stale q0 / n) and kernels of the failed call may still run; the next call must not retire those chunks
// into its own output buffers.
static void reset_pipeline(Session* s) {
    for (int i = 0; i < s->n_ws; i++) {
        Workspace* w = &s->ws[i];
        if (w->stream) cudaStreamSynchronize(w->stream);
        if (w->dp_stream) cudaStreamSynchronize(w->dp_stream);
        if (w->bt_stream) cudaStreamSynchronize(w->bt_stream);
        w->busy = false;
        w->prev_remaining = 0xffffffffu;
    }
    if (s->cstream) cudaStreamSynchronize(s->cstream);
    s->stage_info[0].pending = s->stage_info[1].pending = false;
    cudaGetLastError();
}

int sg_session_align(sg_session* h, const sg_align_params* ap) {
    Session* s = (Session*)h;
    if (!s || s->nq == 0) SG_FAIL(SG_ERR_ARG, "sg_session_align: no queries uploaded");
    if (!s->have_family) SG_FAIL(SG_ERR_ARG, "sg_session_align: run sg_session_family or sg_session_set_family first");
    SG_TRY(validate_align_params(ap));
    SG_CUDA(cudaSetDevice(s->ix->device));
    for (int i = 0; i < s->n_ws; i++) if (s->ws[i].busy) { reset_pipeline(s); break; }   // a previous call failed half way
    SG_TRY(stage_begin(s, 2));
    SG_TRY(launch_prealign(s, *ap));
    SG_TRY(stage_end(s, &s->stats.ms_graph));   // synchronises: the workspace streams may start
    const int rc = align_chunks(s, ap, 0, 0);
    if (rc != SG_OK) { const std::string msg = g_err; reset_pipeline(s); set_error(msg); }
    return rc;
}

// famfinder + aligner on the resident batch in one call (what sg_run_batch runs). The family finding of the whole batch
// runs first: overlapping it with the first chunks' graph / DP kernels was measured and loses (B200, 10 k queries: the
// k-mer search CTAs, 96 KB of shared memory each, starve behind the resident DP CTAs, 3.6 -> 9.5 .. 26.7 ms, and the
// later chunks start late: 99.2 k -> 97.4 k / 96.8 k sequences/s).
int sg_session_run(sg_session* h, const sg_fam_params* fp, const sg_align_params* ap) {
    Session* s = (Session*)h;
    if (!s || s->nq == 0) SG_FAIL(SG_ERR_ARG, "sg_session_run: no queries uploaded");
    SG_TRY(validate_fam_params(fp));
    SG_TRY(validate_align_params(ap));
    SG_TRY(sg_session_family(h, fp));
    return sg_session_align(h, ap);
}

int sg_session_sync(sg_session* h) {
    Session* s = (Session*)h;
    if (!s) SG_FAIL(SG_ERR_ARG, "null session");
    SG_CUDA(cudaStreamSynchronize(s->stream));
    return SG_OK;
}

int sg_session_download_find(sg_session* h, int16_t* scores, uint32_t* ids, uint32_t* nres) {
    Session* s = (Session*)h;
    if (!s || !s->have_find) SG_FAIL(SG_ERR_ARG, "sg_session_download_find: no find results");
    std::vector<uint64_t> keys((uint64_t)s->nq * s->find_max);
    SG_CUDA(cudaMemcpyAsync(keys.data(), s->d_ranked, keys.size() * 8, cudaMemcpyDeviceToHost, s->stream));
    std::vector<uint32_t> nr(s->nq);
    SG_CUDA(cudaMemcpyAsync(nr.data(), s->d_nres, (uint64_t)s->nq * 4, cudaMemcpyDeviceToHost, s->stream));
    SG_CUDA(cudaStreamSynchronize(s->stream));
    for (uint32_t q = 0; q < s->nq; q++) {
        if (nres) nres[q] = nr[q];
        for (uint32_t i = 0; i < nr[q]; i++) {
            const uint64_t kx = keys[(uint64_t)q * s->find_max + i];
            if (scores) scores[(uint64_t)q * s->find_max + i] = (int16_t)(uint16_t)(kx >> 32);
            if (ids) ids[(uint64_t)q * s->find_max + i] = (uint32_t)kx;
        }
    }
    return SG_OK;
}

int sg_session_download_family(sg_session* h, uint32_t fam_stride, uint32_t* fam_ids, float* fam_scores,
                               int32_t* fam_n) {
    Session* s = (Session*)h;
    if (!s || !s->have_family) SG_FAIL(SG_ERR_ARG, "sg_session_download_family: no family");
    std::vector<uint32_t> ids((uint64_t)s->nq * s->fam_cap);
    std::vector<float> sc((uint64_t)s->nq * s->fam_cap);
    std::vector<int32_t> n(s->nq);
    SG_CUDA(cudaMemcpyAsync(ids.data(), s->d_fam_ids, ids.size() * 4, cudaMemcpyDeviceToHost, s->stream));
    SG_CUDA(cudaMemcpyAsync(sc.data(), s->d_fam_scores, sc.size() * 4, cudaMemcpyDeviceToHost, s->stream));
    SG_CUDA(cudaMemcpyAsync(n.data(), s->d_fam_n, n.size() * 4, cudaMemcpyDeviceToHost, s->stream));
    SG_CUDA(cudaStreamSynchronize(s->stream));
    for (uint32_t q = 0; q < s->nq; q++) {
        if (fam_n) fam_n[q] = n[q];
        const uint32_t c = n[q] > 0 ? std::min<uint32_t>((uint32_t)n[q], fam_stride) : 0;
        if (n[q] > 0 && (uint32_t)n[q] > fam_stride) SG_FAIL(SG_ERR_ARG, "fam_stride smaller than a family");
        for (uint32_t i = 0; i < c; i++) {
            if (fam_ids) fam_ids[(uint64_t)q * fam_stride + i] = ids[(uint64_t)q * s->fam_cap + i];
            if (fam_scores) fam_scores[(uint64_t)q * fam_stride + i] = sc[(uint64_t)q * s->fam_cap + i];
        }
    }
    return SG_OK;
}

int sg_session_download_align(sg_session* h, uint32_t* out_cols, uint8_t* out_masks, sg_align_result* results) {
    Session* s = (Session*)h;
    if (!s || !s->have_align) SG_FAIL(SG_ERR_ARG, "sg_session_download_align: no alignment");
    const uint64_t total = s->h_qoff[s->nq];
    if (out_cols) SG_CUDA(cudaMemcpyAsync(out_cols, s->d_out_cols, total * 4, cudaMemcpyDeviceToHost, s->stream));
    if (out_masks) SG_CUDA(cudaMemcpyAsync(out_masks, s->d_out_masks, total, cudaMemcpyDeviceToHost, s->stream));
    std::vector<sg_align_result> tmp;
    sg_align_result* r = results;
    if (!r) { tmp.resize(s->nq); r = tmp.data(); }
    SG_CUDA(cudaMemcpyAsync(r, s->d_results, (uint64_t)s->nq * sizeof(sg_align_result), cudaMemcpyDeviceToHost, s->stream));
    SG_CUDA(cudaStreamSynchronize(s->stream));
    return SG_OK;   // per-query failures (SG_Q_LIMIT, SG_Q_NOSPACE, ...) are in results[].status
}

int sg_session_stats(sg_session* h, sg_stage_stats* st, int reset) {
    Session* s = (Session*)h;
    if (!s) SG_FAIL(SG_ERR_ARG, "null session");
    {   // the device-side counters (postings scanned, DP cells) as of now
        unsigned long long cnt[2];
        SG_CUDA(cudaSetDevice(s->ix->device));
        SG_CUDA(cudaMemcpyAsync(cnt, s->d_counters, 16, cudaMemcpyDeviceToHost, s->stream));
        SG_CUDA(cudaStreamSynchronize(s->stream));
        s->stats.postings = cnt[0];
        s->stats.cells = cnt[1];
    }
    if (st) *st = s->stats;
    if (reset) {
        s->stats = sg_stage_stats{};
        SG_CUDA(cudaMemsetAsync(s->d_counters, 0, 16, s->stream));
    }
    return SG_OK;
}

// Device clock around a run of stage calls: CUDA events on the session's stream. Every stage call returns with the
// session's streams drained, so the two events bracket all the kernels launched between them.
int sg_session_timer(sg_session* h, int stop, float* ms) {
    Session* s = (Session*)h;
    if (!s || (stop && !ms)) SG_FAIL(SG_ERR_ARG, "sg_session_timer: bad argument");
    SG_CUDA(cudaSetDevice(s->ix->device));
    SG_CUDA(cudaEventRecord(s->ev[stop ? 7 : 6], s->stream));
    if (stop) {
        SG_CUDA(cudaEventSynchronize(s->ev[7]));
        SG_CUDA(cudaEventElapsedTime(ms, s->ev[6], s->ev[7]));
    }
    return SG_OK;
}

int sg_session_dump_graph(sg_session* h, uint32_t q, uint32_t cap_nodes, uint32_t cap_edges, uint32_t* V, uint32_t* E,
                          uint32_t* col, uint8_t* mask, float* weight, uint32_t* pred_off, uint32_t* preds) {
    Session* s = (Session*)h;
    if (!s || !s->have_align || q >= s->nq) SG_FAIL(SG_ERR_ARG, "sg_session_dump_graph: bad argument");
    GraphHdr hd;
    SG_CUDA(cudaMemcpy(&hd, s->d_hdr + q, sizeof(hd), cudaMemcpyDeviceToHost));
    if (V) *V = hd.V;
    if (E) *E = hd.E;
    if (hd.V > cap_nodes || hd.E > cap_edges) SG_FAIL(SG_ERR_ARG, "sg_session_dump_graph: capacity too small");
    // the workspace that handled q's chunk last: its arrays are only still there if no later chunk reused it
    const Workspace* w = nullptr;
    for (int i = 0; i < s->n_ws; i++)
        if (q >= s->ws[i].last_q0 && q < s->ws[i].last_q0 + s->ws[i].last_n) w = &s->ws[i];
    if (!w) SG_FAIL(SG_ERR_ARG, "sg_session_dump_graph: the query's workspace has been reused");
    const uint64_t ql = q - w->last_q0, io = ql * s->icap;
    if (col) SG_CUDA(cudaMemcpy(col, w->d_ncol + io, (uint64_t)hd.V * 4, cudaMemcpyDeviceToHost));
    if (mask) SG_CUDA(cudaMemcpy(mask, w->d_nmask + io, hd.V, cudaMemcpyDeviceToHost));
    if (weight) SG_CUDA(cudaMemcpy(weight, w->d_nweight + io, (uint64_t)hd.V * 4, cudaMemcpyDeviceToHost));
    if (pred_off) SG_CUDA(cudaMemcpy(pred_off, w->d_pred_off + ql * (s->icap + 1), ((uint64_t)hd.V + 1) * 4, cudaMemcpyDeviceToHost));
    if (preds) SG_CUDA(cudaMemcpy(preds, w->d_preds + io, (uint64_t)hd.E * 4, cudaMemcpyDeviceToHost));
    return SG_OK;
}

// ------------------------------------------------------------------------- host-buffer entry points
// One call = upload + kernels + download. The session (device workspace) is cached in the index and reused
// by later calls, so steady-state calls do no cudaMalloc. Calls on one index are serialised by its mutex.
namespace {
constexpr uint32_t HOST_CALL_MAX_Q = 65536;  // queries per internal session; larger batches are looped

struct SessionLease {
    Index* ix; Session* s = nullptr; int rc = SG_OK;
    SessionLease(sg_index* h, const uint64_t* qoff, uint32_t nq) : ix((Index*)h) {
        ix->mu.lock();
        uint64_t maxlen = 2;
        for (uint32_t i = 0; i < nq; i++) maxlen = std::max<uint64_t>(maxlen, qoff[i + 1] - qoff[i]);
        const uint32_t want_q = std::min(nq, HOST_CALL_MAX_Q);
        uint64_t want_bases = 0;
        for (uint32_t a = 0; a < nq; a += want_q)
            want_bases = std::max<uint64_t>(want_bases, qoff[std::min(nq, a + want_q)] - qoff[a]);
        Session* c = (Session*)ix->cached;
        if (c && c->max_q >= want_q && c->max_bases >= want_bases) { s = c; return; }
        if (c) { sg_session_destroy((sg_session*)c); ix->cached = nullptr; }
        sg_session* ns = nullptr;
        rc = sg_session_create(h, want_q, want_bases, &ns);
        if (rc != SG_OK) { if (ns) sg_session_destroy(ns); return; }
        ix->cached = ns;
        s = (Session*)ns;
    }
    ~SessionLease() { ix->mu.unlock(); }
    uint32_t step() const { return s->max_q; }
};
// aligner stage (fp == null) or famfinder + aligner pipeline whose output columns/bases stream into the caller's
// buffers chunk by chunk (stage_chunk above)
static int align_streaming(Session* s, const sg_fam_params* fp, const sg_align_params* ap, uint32_t* cols, uint8_t* masks) {
    s->host_cols = cols; s->host_masks = masks; s->stage_next = 0;
    const int rc = fp ? sg_session_run((sg_session*)s, fp, ap) : sg_session_align((sg_session*)s, ap);
    s->host_cols = nullptr; s->host_masks = nullptr;
    s->stage_info[0].pending = s->stage_info[1].pending = false;
    return rc;
}
}  // namespace

int sg_find_batch(sg_index* ix, const uint8_t* qmasks, const uint64_t* qoff, uint32_t nq, uint32_t max,
                  int16_t* scores, uint32_t* ids, uint32_t* nres) {
    if (!ix || !qmasks || !qoff || nq == 0) SG_FAIL(SG_ERR_ARG, "sg_find_batch: bad argument");
    SessionLease L(ix, qoff, nq);
    if (L.rc) return L.rc;
    sg_session* s = (sg_session*)L.s;
    const uint32_t m = std::min(max, ((Index*)ix)->N);
    for (uint32_t a = 0; a < nq; a += L.step()) {
        const uint32_t n = std::min(L.step(), nq - a);
        SG_TRY(sg_session_upload(s, qmasks, qoff + a, n, nullptr));
        SG_TRY(sg_session_find(s, max));
        SG_TRY(sg_session_download_find(s, scores ? scores + (uint64_t)a * m : nullptr, ids ? ids + (uint64_t)a * m : nullptr,
                                        nres ? nres + a : nullptr));
    }
    return SG_OK;
}

int sg_turn_batch(sg_index* ix, const uint8_t* qmasks, const uint64_t* qoff, uint32_t nq, int mode, int32_t* turn) {
    if (!ix || !qmasks || !qoff || nq == 0 || !turn) SG_FAIL(SG_ERR_ARG, "sg_turn_batch: bad argument");
    SessionLease L(ix, qoff, nq);
    if (L.rc) return L.rc;
    sg_session* s = (sg_session*)L.s;
    for (uint32_t a = 0; a < nq; a += L.step()) {
        const uint32_t n = std::min(L.step(), nq - a);
        SG_TRY(sg_session_upload(s, qmasks, qoff + a, n, nullptr));
        SG_TRY(sg_session_turn(s, mode, turn + a));
    }
    return SG_OK;
}

int sg_family_batch(sg_index* ix, const uint8_t* qmasks, const uint64_t* qoff, uint32_t nq,
                    const int64_t* exclude_ids, const sg_fam_params* fp, uint32_t fam_stride, uint32_t* fam_ids,
                    float* fam_scores, int32_t* fam_n) {
    return sg_family_batch_aligned(ix, qmasks, nullptr, qoff, nq, exclude_ids, fp, fam_stride, fam_ids, fam_scores, fam_n);
}

int sg_family_batch_aligned(sg_index* ix, const uint8_t* qmasks, const uint32_t* qcols, const uint64_t* qoff, uint32_t nq,
                            const int64_t* exclude_ids, const sg_fam_params* fp, uint32_t fam_stride, uint32_t* fam_ids,
                            float* fam_scores, int32_t* fam_n) {
    if (!ix || !qmasks || !qoff || nq == 0) SG_FAIL(SG_ERR_ARG, "sg_family_batch: bad argument");
    SessionLease L(ix, qoff, nq);
    if (L.rc) return L.rc;
    sg_session* s = (sg_session*)L.s;
    for (uint32_t a = 0; a < nq; a += L.step()) {
        const uint32_t n = std::min(L.step(), nq - a);
        SG_TRY(sg_session_upload(s, qmasks, qoff + a, n, exclude_ids ? exclude_ids + a : nullptr));
        if (qcols) SG_TRY(sg_session_set_query_columns(s, qcols + qoff[a]));
        SG_TRY(sg_session_family(s, fp));
        SG_TRY(sg_session_download_family(s, fam_stride, fam_ids ? fam_ids + (uint64_t)a * fam_stride : nullptr,
                                          fam_scores ? fam_scores + (uint64_t)a * fam_stride : nullptr,
                                          fam_n ? fam_n + a : nullptr));
    }
    return SG_OK;
}

int sg_align_batch(sg_index* ix, const uint8_t* qmasks, const uint64_t* qoff, uint32_t nq, const uint32_t* fam_ids,
                   const uint64_t* fam_off, const sg_align_params* ap, uint32_t* out_cols, uint8_t* out_masks,
                   sg_align_result* results) {
    if (!ix || !qmasks || !qoff || nq == 0 || !fam_ids || !fam_off) SG_FAIL(SG_ERR_ARG, "sg_align_batch: bad argument");
    SessionLease L(ix, qoff, nq);
    if (L.rc) return L.rc;
    sg_session* s = (sg_session*)L.s;
    for (uint32_t a = 0; a < nq; a += L.step()) {
        const uint32_t n = std::min(L.step(), nq - a);
        SG_TRY(sg_session_upload(s, qmasks, qoff + a, n, nullptr));
        SG_TRY(sg_session_set_family(s, fam_ids, fam_off + a));
        SG_TRY(align_streaming(L.s, nullptr, ap, out_cols ? out_cols + qoff[a] : nullptr, out_masks ? out_masks + qoff[a] : nullptr));
        SG_TRY(sg_session_download_align(s, nullptr, nullptr, results ? results + a : nullptr));
    }
    return SG_OK;
}

int sg_run_batch(sg_index* ix, const uint8_t* qmasks, const uint64_t* qoff, uint32_t nq, const int64_t* exclude_ids,
                 const sg_fam_params* fp, const sg_align_params* ap, uint32_t* out_cols, uint8_t* out_masks,
                 sg_align_result* results) {
    if (!ix || !qmasks || !qoff || nq == 0) SG_FAIL(SG_ERR_ARG, "sg_run_batch: bad argument");
    SessionLease L(ix, qoff, nq);
    if (L.rc) return L.rc;
    sg_session* s = (sg_session*)L.s;
    for (uint32_t a = 0; a < nq; a += L.step()) {
        const uint32_t n = std::min(L.step(), nq - a);
        SG_TRY(sg_session_upload(s, qmasks, qoff + a, n, exclude_ids ? exclude_ids + a : nullptr));
        SG_TRY(align_streaming(L.s, fp, ap, out_cols ? out_cols + qoff[a] : nullptr, out_masks ? out_masks + qoff[a] : nullptr));
        SG_TRY(sg_session_download_align(s, nullptr, nullptr, results ? results + a : nullptr));
    }
    return SG_OK;
}

// ---- --search stage and the sequence comparator ------------------------------------------------------
namespace {
int validate_cmp(int iupac, int correction, int cover, const char* who) {
    if (iupac < 0 || iupac > 2 || correction < 0 || correction > 1 || cover < 0 || cover > 8) SG_FAIL(SG_ERR_ARG, std::string(who) + ": comparator rule out of range");
    if (cover == 0 && correction != 0) SG_FAIL(SG_ERR_ARG, std::string(who) + ": only fractional identity can be distance corrected");   // src/cseq_comparator.cpp:478-481
    return SG_OK;
}
// jukes_cantor() of the reference (src/cseq_comparator.cpp:42-44): host double log, as the reference computes it
float jukes_cantor(float in) { return (float)(-3.0 / 4 * std::log(1.0 - 4.0 / 3 * in)); }

int upload_aligned(Session* s, const uint8_t* amasks, const uint32_t* acols, const uint64_t* aoff, uint32_t n) {
    SG_TRY(sg_session_upload((sg_session*)s, amasks, aoff, n, nullptr));
    if (!s->d_acols) SG_TRY(dmalloc(&s->d_acols, s->max_bases));
    const uint64_t base = aoff[0], total = aoff[n] - base;
    for (uint32_t q = 0; q < n; q++)
        for (uint64_t j = aoff[q] + 1; j < aoff[q + 1]; j++)
            if (acols[j] <= acols[j - 1]) SG_FAIL(SG_ERR_ARG, "aligned sequence: columns must be strictly increasing");
    SG_CUDA(cudaMemcpyAsync(s->d_acols, acols + base, total * 4, cudaMemcpyHostToDevice, s->stream));
    return SG_OK;
}
}  // namespace

int sg_identity_batch(sg_index* ix, const uint8_t* amasks, const uint32_t* acols, const uint64_t* aoff, uint32_t nq,
                      const uint32_t* ref_ids, const uint64_t* ref_off, int iupac, int correction, int cover,
                      int filter_lowercase, float* out) {
    if (!ix || !amasks || !acols || !aoff || !ref_ids || !ref_off || !out || nq == 0) SG_FAIL(SG_ERR_ARG, "sg_identity_batch: bad argument");
    SG_TRY(validate_cmp(iupac, correction, cover, "sg_identity_batch"));
    Index* X = (Index*)ix;
    for (uint64_t i = ref_off[0]; i < ref_off[nq]; i++) if (ref_ids[i] >= X->N) SG_FAIL(SG_ERR_ARG, "sg_identity_batch: reference id out of range");
    SessionLease L(ix, aoff, nq);
    if (L.rc) return L.rc;
    Session* s = L.s;
    for (uint32_t a = 0; a < nq; a += L.step()) {
        const uint32_t n = std::min(L.step(), nq - a);
        SG_TRY(upload_aligned(s, amasks, acols, aoff + a, n));
        const uint64_t p0 = ref_off[a], np = ref_off[a + n] - p0;
        if (np == 0) continue;
        uint32_t* d_ids = nullptr; uint64_t* d_off = nullptr; float* d_sc = nullptr;
        std::vector<uint64_t> off(n + 1);
        for (uint32_t i = 0; i <= n; i++) off[i] = ref_off[a + i] - p0;
        int rc = dmalloc(&d_ids, np);
        if (rc == SG_OK) rc = dmalloc(&d_off, (uint64_t)n + 1);
        if (rc == SG_OK) rc = dmalloc(&d_sc, np);
        if (rc == SG_OK && cudaMemcpyAsync(d_ids, ref_ids + p0, np * 4, cudaMemcpyHostToDevice, s->stream) != cudaSuccess) rc = SG_ERR_CUDA;
        if (rc == SG_OK && cudaMemcpyAsync(d_off, off.data(), ((uint64_t)n + 1) * 8, cudaMemcpyHostToDevice, s->stream) != cudaSuccess) rc = SG_ERR_CUDA;
        if (rc == SG_OK) rc = launch_identity(s, s->d_qmasks, s->d_acols, s->d_qoff, n, nullptr, nullptr, 0, d_ids, d_off, iupac, cover, filter_lowercase, 0, d_sc);
        if (rc == SG_OK && cudaMemcpyAsync(out + p0, d_sc, np * 4, cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) rc = SG_ERR_CUDA;
        if (rc == SG_OK && cudaStreamSynchronize(s->stream) != cudaSuccess) rc = SG_ERR_CUDA;
        cudaFree(d_ids); cudaFree(d_off); cudaFree(d_sc);
        if (rc != SG_OK) { if (rc == SG_ERR_CUDA) set_error("sg_identity_batch: CUDA failure"); return rc; }
        if (correction == 1) for (uint64_t i = p0; i < p0 + np; i++) out[i] = jukes_cantor(out[i]);
    }
    return SG_OK;
}

int sg_search_batch(sg_index* ix, const uint8_t* amasks, const uint32_t* acols, const uint64_t* aoff, uint32_t nq,
                    const sg_search_params* sp, uint32_t* out_ids, float* out_scores, uint32_t* out_n) {
    if (!ix || !amasks || !acols || !aoff || !sp || !out_ids || !out_scores || !out_n || nq == 0) SG_FAIL(SG_ERR_ARG, "sg_search_batch: bad argument");
    SG_TRY(validate_cmp(sp->iupac, sp->correction, sp->cover, "sg_search_batch"));
    if (sp->kmer_candidates == 0 || sp->max_result == 0) SG_FAIL(SG_ERR_ARG, "sg_search_batch: kmer_candidates and max_result must be > 0");
    Index* X = (Index*)ix;
    SessionLease L(ix, aoff, nq);
    if (L.rc) return L.rc;
    Session* s = L.s;
    const uint32_t cand = std::min(sp->kmer_candidates, X->N), mr = sp->max_result;
    if (cand > s->pair_cap) {
        if (s->d_pair) { cudaFree(s->d_pair); s->d_pair = nullptr; s->pair_cap = 0; }
        SG_TRY(dmalloc(&s->d_pair, (uint64_t)s->max_q * cand));
        s->pair_cap = cand;
    }
    if (mr > s->sres_cap) {
        cudaFree(s->d_sids); cudaFree(s->d_sscores); s->d_sids = nullptr; s->d_sscores = nullptr; s->sres_cap = 0;
        SG_TRY(dmalloc(&s->d_sids, (uint64_t)s->max_q * mr));
        SG_TRY(dmalloc(&s->d_sscores, (uint64_t)s->max_q * mr));
        s->sres_cap = mr;
    }
    if (!s->d_sn) SG_TRY(dmalloc(&s->d_sn, (uint64_t)s->max_q));
    for (uint32_t a = 0; a < nq; a += L.step()) {
        const uint32_t n = std::min(L.step(), nq - a);
        SG_TRY(upload_aligned(s, amasks, acols, aoff + a, n));
        SG_TRY(sg_session_find((sg_session*)s, cand));   // index->find(*c, vc, kmer_candidates) (src/search_filter.cpp:303)
        SG_TRY(launch_identity(s, s->d_qmasks, s->d_acols, s->d_qoff, n, s->d_ranked, s->d_nres, s->find_max, nullptr, nullptr,
                               sp->iupac, sp->cover, sp->filter_lowercase, sp->ignore_super, s->d_pair));
        if (sp->correction == 0) {
            SG_TRY(launch_search_select(s, s->d_pair, s->d_ranked, s->d_nres, s->find_max, n, mr, sp->min_sim, s->d_sids, s->d_sscores, s->d_sn));
            SG_CUDA(cudaMemcpyAsync(out_ids + (uint64_t)a * mr, s->d_sids, (uint64_t)n * mr * 4, cudaMemcpyDeviceToHost, s->stream));
            SG_CUDA(cudaMemcpyAsync(out_scores + (uint64_t)a * mr, s->d_sscores, (uint64_t)n * mr * 4, cudaMemcpyDeviceToHost, s->stream));
            SG_CUDA(cudaMemcpyAsync(out_n + a, s->d_sn, (uint64_t)n * 4, cudaMemcpyDeviceToHost, s->stream));
            SG_CUDA(cudaStreamSynchronize(s->stream));
        } else {
            // --search-correction jc: the reference corrects with the host's double log before it sorts; the identities
            // come from the device, the logarithm and the (score, name) selection of max_result items run here
            const uint32_t st = s->find_max;
            std::vector<float> pair((uint64_t)n * st);
            std::vector<uint64_t> keys((uint64_t)n * st);
            std::vector<uint32_t> nr(n), rank;
            SG_CUDA(cudaMemcpyAsync(pair.data(), s->d_pair, pair.size() * 4, cudaMemcpyDeviceToHost, s->stream));
            SG_CUDA(cudaMemcpyAsync(keys.data(), s->d_ranked, keys.size() * 8, cudaMemcpyDeviceToHost, s->stream));
            SG_CUDA(cudaMemcpyAsync(nr.data(), s->d_nres, (uint64_t)n * 4, cudaMemcpyDeviceToHost, s->stream));
            if (X->d_name_rank) { rank.resize(X->N); SG_CUDA(cudaMemcpyAsync(rank.data(), X->d_name_rank, (uint64_t)X->N * 4, cudaMemcpyDeviceToHost, s->stream)); }
            SG_CUDA(cudaStreamSynchronize(s->stream));
            struct item { float score; uint32_t rk, id; };
            std::vector<item> v;
            for (uint32_t q = 0; q < n; q++) {
                v.clear();
                for (uint32_t i = 0; i < nr[q]; i++) {
                    uint32_t bits; const float raw = pair[(uint64_t)q * st + i];
                    memcpy(&bits, &raw, 4);
                    if (bits >= 0x7f800000u) continue;
                    const float jc = jukes_cantor(raw);
                    if (!(jc == jc)) continue;
                    const uint32_t id = (uint32_t)keys[(uint64_t)q * st + i];
                    v.push_back({jc, rank.empty() ? id : rank[id], id});
                }
                auto gt = [](const item& x, const item& y) { return x.score > y.score || (x.score == y.score && x.rk > y.rk); };
                const size_t m = std::min<size_t>(mr, v.size());
                std::partial_sort(v.begin(), v.begin() + m, v.end(), gt);
                uint32_t e = 0;
                while (e < m && v[e].score > sp->min_sim) { out_ids[(uint64_t)(a + q) * mr + e] = v[e].id; out_scores[(uint64_t)(a + q) * mr + e] = v[e].score; e++; }
                out_n[a + q] = e;
            }
        }
        // search_filter skips sequences shorter than 20 bases (src/search_filter.cpp:253-256)
        for (uint32_t q = 0; q < n; q++) if (aoff[a + q + 1] - aoff[a + q] < 20) out_n[a + q] = 0;
    }
    return SG_OK;
}

}  // extern "C"
