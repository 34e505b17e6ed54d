/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of SINA's per-query hot path.
 *
 * Plain C, flat arrays, one function per reference function on the path; every function cites the
 * reference file:line it follows (paths relative to the SINA source tree, commit b0763146).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this. The product (sina_b200/) never links or calls it.
 *
 * Parity status: PINNED. tests/test_oracle_vs_ref.py diffs every function below against the
 * reference's own sources compiled in place (oracle/_ref/libsina_ref.so, built by oracle/Makefile
 * from /root/reference/src) -- all seven mesh cell fields bitwise, graph arrays, output strings --
 * and tests/golden/ holds vectors generated from that build plus the reference's unit-test KATs.
 * Not pinnable: reference index order vs a real ARB database, --fs-kmer-mm != 0.
 */
#ifndef SINA_ORACLE_H
#define SINA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- base encoding (src/aligned_base.h:38-52, src/aligned_base.cpp:70-121) */
#define SO_BASEM_LC 0x10
int so_char_to_mask(int c);          /* -1: bad character; 0: '-' or '.' */
int so_mask_to_char(int mask, int dna);
/* cseq::append(const char*) (src/cseq.cpp:63-77): returns #bases, *width = #columns; -1-i on bad char at i */
int64_t so_encode_aligned(const char* str, uint8_t* masks, uint32_t* cols, uint64_t cap, uint32_t* width);
/* cseq::getAligned(nodots=true) (src/cseq.cpp:135-174). out must hold width+1 bytes */
void so_render_aligned(const uint8_t* masks, const uint32_t* cols, uint32_t n, uint32_t width, int dna, char* out);

/* ---- k-mers (src/kmer.h:46-203). mode: 0 all, 1 unique, 2 prefix(A), 3 unique_prefix(A) */
int64_t so_kmers(const uint8_t* masks, uint32_t n, int k, int mode, uint32_t* out, uint64_t cap);

/* ---- index build + find (src/kmer_search.cpp:152-276, 365-420) as CSR posting lists */
typedef struct so_index {
    uint32_t N;
    int k, nofast;
    uint64_t n_kmers;
    uint64_t* list_off; /* n_kmers+1 */
    uint32_t* postings; /* ascending ids per list */
} so_index;
so_index* so_index_build(uint32_t N, const uint8_t* masks, const uint64_t* off, int k, int nofast);
void so_index_free(so_index* ix);
uint32_t so_find(const so_index* ix, const uint8_t* q, uint32_t qlen, uint32_t max, int16_t* scores,
                 uint32_t* ids, uint64_t* postings);

/* ---- family selection (src/famfinder.cpp:497-612, 474-491) */
typedef struct so_fam_params {
    uint32_t fs_min, fs_max;
    float fs_msc, fs_msc_max;
    uint32_t fs_min_len, fs_req_full, fs_full_len, fs_req_gaps, fs_req;
    int leave_query_out;
} so_fam_params;
/* exclude_id: reference with the query's name (-1: none). returns family size, -1 if < fs_req */
/* famfinder::impl::turn_check (--turn): 0 none, 1 reversed, 2 complemented, 3 reversed + complemented */
int so_turn_check(const so_index* ix, const uint8_t* q, uint32_t qlen, int all, int32_t* scores4);
int so_family(const so_index* ix, const uint64_t* off, const uint32_t* cols, const uint8_t* q, uint32_t qlen,
              int64_t exclude_id, const so_fam_params* p, uint32_t* ids, float* scores, uint32_t cap,
              uint64_t* postings);
/* same filter applied to an already ranked candidate list (rank order), n_total = index size;
 * returns family size or -2 if the window was too small (caller must retry with more candidates) */
int so_family_from_ranked(const uint32_t* cand_ids, const int16_t* cand_scores, uint32_t n_cand, uint32_t n_total,
                          const uint64_t* off, const uint32_t* cols, int64_t exclude_id, const so_fam_params* p,
                          uint32_t* ids, float* scores, uint32_t cap);

/* ---- family graph (src/mseq.cpp:47-118, src/graph.h:332-357,451-488) */
typedef struct so_graph {
    uint32_t V, E, n_first, n_last, W;
    uint32_t* col;
    uint8_t* mask;
    float* weight;
    uint32_t* pred_off; /* V+1 */
    uint32_t* preds;    /* ascending node id per node */
    uint32_t* first;    /* nodes without predecessor, ascending */
    uint32_t* last;     /* nodes without successor, ascending */
} so_graph;
so_graph* so_graph_build(const uint32_t* fam, uint32_t F, const uint8_t* masks, const uint32_t* cols,
                         const uint64_t* off, uint32_t W, float fs_weight);
void so_graph_free(so_graph* g);

/* ---- mesh DP + backtrack + gap placement (src/mesh.h:282-374,453-528,534-739; src/cseq.cpp:456-594) */
typedef struct so_align_params {
    float match_score, mismatch_score, gap_penalty, gap_ext_penalty, fs_weight;
    int overhang;  /* 0 attach 1 remove 2 edge */
    int lowercase; /* 0 none 1 original 2 unaligned */
    int insertion; /* 0 shift, 1 forbid (transition_aspace_aware), 2 remove (= shift, src/cseq.cpp:462-464) */
    int realign;
} so_align_params;

typedef struct so_align_result {
    int status; /* 0 DP, 1 copied, 2 skipped, 3 no space (runtime_error), 4 no family */
    float score, raw, sum_weight;
    int head, tail, qual;
    uint32_t n_nodes, fam_used, n_out; /* n_out: bases in output (overhang remove drops some) */
    uint32_t end_m, end_s;
} so_align_result;

typedef struct so_mesh { /* full mesh in the reference's field set (src/mesh.h:282-290) */
    uint32_t V, L;
    uint32_t *value_midx, *value_sidx, *gapm_idx, *gaps_idx;
    float *value, *gapm_val, *gaps_val;
} so_mesh;
/* positional column weights for every later so_mesh_compute / so_backtrack / so_align / so_run_batch (n = 0: none).
 * The pointer is kept, not copied. src/scoring_schemes.h:166-241 */
void so_set_column_weights(const float* w, uint32_t n);
so_mesh* so_mesh_compute(const so_graph* g, const uint8_t* q, uint32_t qlen, const so_align_params* p);
void so_mesh_free(so_mesh* m);

/* backtrack on a computed mesh. out_cols/out_masks: qlen entries. returns status (0 or 3) */
int so_backtrack(const so_graph* g, const so_mesh* m, const uint8_t* q, uint32_t qlen, const so_align_params* p,
                 so_align_result* r, uint32_t* out_cols, uint8_t* out_masks);

/* cseq::fix_duplicate_positions (src/cseq.cpp:456-594). returns 0, or 1 for the runtime_error case */
int so_fix_duplicate_positions(uint32_t* pos, uint8_t* masks, uint32_t n, uint32_t width, int lowercase);

/* aligner::operator() (src/align.cpp:307-460): pre-steps + graph + DP + backtrack. fam is permuted in
 * place by the (unstable) partition exactly as libstdc++ does. */
int so_align(uint32_t* fam, uint32_t F, const uint8_t* masks, const uint32_t* cols, const uint64_t* off, uint32_t W,
             const uint8_t* q, uint32_t qlen, const so_align_params* p, so_align_result* r, uint32_t* out_cols,
             uint8_t* out_masks);

/* whole path for a batch, threaded over queries (cpu_baseline "port" leg). qoff: nq+1 offsets into qmasks.
 * out_cols/out_masks indexed like qmasks. returns threads used */
int so_run_batch(const so_index* ix, const uint8_t* masks, const uint32_t* cols, const uint64_t* off, uint32_t W,
                 uint32_t nq, const uint8_t* qmasks, const uint64_t* qoff, const int64_t* exclude_ids,
                 const so_fam_params* fp, const so_align_params* ap, int nthreads, so_align_result* results,
                 uint32_t* out_cols, uint8_t* out_masks, uint64_t* cells_total, uint64_t* postings_total);

#ifdef __cplusplus
}
#endif
#endif
