/* sina_b200 -- C-ABI of the B200-native replacement for SINA's per-query hot path
 * (k-mer family finding -> family graph -> mesh DP -> backtrack -> gap placement).
 *
 * Plain C types only; the library (sina_b200/libsina_b200.so) is self-contained CUDA (static cudart).
 * Every entry point returns 0 on success, non-zero on error; sg_last_error() gives the message
 * (thread-local). Nothing throws across this boundary. There is NO CPU fallback: without a CUDA
 * device every compute call fails with SG_ERR_CUDA.
 *
 * Reference interfaces replaced (paths in the SINA source tree, commit b0763146):
 *   sg_index_create      kmer_search::get_kmer_search + impl::build   src/kmer_search.cpp:118-134,152-276
 *                        (+ the reference rows query_arb::getCseq hands out, src/query_arb.cpp:742-770)
 *   sg_find_batch        search::find / kmer_search::impl::find       src/search.h:103, src/kmer_search.cpp:365-420
 *   sg_family_batch      famfinder::impl::match + stage body          src/famfinder.cpp:497-612,439-494
 *   sg_align_batch       aligner::operator() / do_align: mseq ctor, compute(), backtrack(),
 *                        cseq::fix_duplicate_positions                src/align.cpp:307-521, src/mseq.cpp:47-118,
 *                                                                     src/mesh.h:453-739, src/cseq.cpp:456-594
 *   sg_run_batch         the famfinder -> aligner node pair           src/sina.cpp:511,516
 *   sg_turn_batch        famfinder::impl::turn_check (--turn)          src/famfinder.cpp:344-378
 *   sg_index_set_column_weights   scoring_scheme_weighted + alignment_stats weights (--filter)
 *                                                                     src/scoring_schemes.h:166-241, src/align.cpp:409-415
 *
 * Data layout: bases are SINA's IUPAC bit masks, one byte each (A=1 G=2 C=4 T/U=8, +16 lowercase;
 * src/aligned_base.h:38-52). A set of sequences is (masks[], off[n+1]); aligned rows add cols[] (alignment
 * column of every base, strictly increasing inside a row).
 */
#ifndef SINA_B200_H
#define SINA_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SG_OK 0
#define SG_ERR_ARG 1    /* invalid argument / unsupported option */
#define SG_ERR_CUDA 2   /* CUDA runtime error or no device */
#define SG_ERR_LIMIT 3  /* a documented capacity limit was exceeded */

/* per-query status codes (sg_align_result.status) */
#define SG_Q_ALIGNED 0  /* aligned by the mesh DP */
#define SG_Q_COPIED 1   /* alignment copied from a relative containing the query (src/align.cpp:349-388) */
#define SG_Q_SKIPPED 2  /* all relatives contained the query and --realign removed them (src/align.cpp:337-348) */
#define SG_Q_NOSPACE 3  /* fix_duplicate_positions' runtime_error: more bases than columns (src/cseq.cpp:557-560) */
#define SG_Q_NOFAMILY 4 /* fewer than fs_req relatives (src/famfinder.cpp:486-491) */
#define SG_Q_LIMIT 5    /* this query's family graph exceeds a device capacity (nodes, columns, traceback arena): it is left
                         * unaligned, the rest of the batch is not affected (per-query failure, as src/famfinder.cpp:486-491
                         * and src/cseq.cpp:557-560 fail one sequence, not the run) */

typedef struct sg_index sg_index;     /* reference MSA + k-mer posting lists, resident in one GPU's HBM */
typedef struct sg_session sg_session; /* a batch of queries + all per-batch device workspace */

/* famfinder options, defaults = reference defaults (src/famfinder.cpp:155-195) */
typedef struct sg_fam_params {
    uint32_t fs_min;       /* --fs-min 40 */
    uint32_t fs_max;       /* --fs-max 40 */
    float fs_msc;          /* --fs-msc 0.7 */
    float fs_msc_max;      /* --fs-msc-max 2: candidates more identical to the query than this are dropped (remove_similar,
                            * src/famfinder.cpp:553-556). Identities are <= 1; values < 1 need the queries' own positions
                            * (sg_family_batch_aligned / sg_session_set_query_columns) */
    uint32_t fs_min_len;   /* --fs-min-len 150 */
    uint32_t fs_req_full;  /* --fs-req-full 1 */
    uint32_t fs_full_len;  /* --fs-full-len 1400 */
    uint32_t fs_req_gaps;  /* --fs-req-gaps 10 */
    uint32_t fs_req;       /* --fs-req 1 */
    int32_t leave_query_out; /* --fs-leave-query-out */
} sg_fam_params;

/* aligner options, defaults = reference defaults (src/align.cpp:232-259) */
typedef struct sg_align_params {
    float match_score;     /* --match-score 2 */
    float mismatch_score;  /* --mismatch-score -1 */
    float gap_penalty;     /* --pen-gap 5 */
    float gap_ext_penalty; /* --pen-gapext 2 */
    float fs_weight;       /* --fs-weight 1 */
    int32_t overhang;      /* --overhang: 0 attach, 1 remove, 2 edge */
    int32_t lowercase;     /* --lowercase: 0 none, 1 original, 2 unaligned */
    int32_t insertion;     /* --insertion: 0 shift, 1 forbid (transition_aspace_aware, src/mesh.h:377-438; runs in the
                            * generic DP kernel), 2 remove (= shift, as in the reference) */
    int32_t realign;       /* --realign */
} sg_align_params;

typedef struct sg_align_result {
    int32_t status;     /* SG_Q_* */
    float score;        /* raw / sum_weight (backtrack() return value, src/mesh.h:738) */
    float raw;          /* value of the end cell */
    float sum_weight;
    int32_t head, tail; /* cutoff_head / cutoff_tail */
    int32_t qual;       /* align_quality_slv (src/align.cpp:509) */
    uint32_t n_nodes;   /* family graph size */
    uint32_t fam_used;  /* relatives left after the contains-query partition */
    uint32_t n_out;     /* bases written to out_cols/out_masks (--overhang remove drops some) */
    uint32_t end_m, end_s;
} sg_align_result;

void sg_default_fam_params(sg_fam_params* p);
void sg_default_align_params(sg_align_params* p);
const char* sg_last_error(void);
int sg_device_count(void);

/* ---- index ------------------------------------------------------------------------------------- */
/* Uploads the reference alignment (N rows, W columns) to `device` and builds the k-mer posting lists
 * there. Reference index i = row i of the input. fast mode (nofast=0) indexes only k-mers starting
 * with A, as the reference does. */
int sg_index_create(const uint8_t* masks, const uint32_t* cols, const uint64_t* row_off, uint32_t N, uint32_t W,
                    int k, int nofast, int device, sg_index** out);
void sg_index_destroy(sg_index* ix);
/* Positional column weights of the alignment (alignment_stats::getWeights(), what --filter selects in the reference,
 * src/alignment_stats.cpp:54-112): from then on the aligner scores with scoring_scheme_weighted
 * (src/scoring_schemes.h:166-241, src/align.cpp:409-415) instead of scoring_scheme_simple. weights[W], one per column;
 * n = 0 clears them. The reference reads weights[column + 1 + insertion length] without a bounds check; an index past
 * W - 1 reads weights[W - 1] here. Not to be called while a batch is in flight on this index. */
int sg_index_set_column_weights(sg_index* ix, const float* weights, uint32_t n);
/* n_postings: total posting entries; n_tiles: reference-id tiles of the search histogram */
int sg_index_info(const sg_index* ix, uint32_t* N, uint32_t* W, int* k, int* nofast, uint64_t* n_postings,
                  uint32_t* n_tiles, uint32_t* tile_size);
/* All posting lists at once, as the reference keeps them (kmer_idx[kmer] = ascending reference ids,
 * src/kmer_search.cpp:152-211): list_off[n_slots + 1] with n_slots = 4^(k-1) in fast mode (k-mers starting with A; the
 * k-mer value IS the slot) and 4^k otherwise; ids[n_postings] (may be null to get the offsets only). This is what the
 * host writes into a .sidx index cache (src/kmer_search.cpp:278-303). */
int sg_index_export_lists(const sg_index* ix, uint64_t* list_off, uint32_t* ids);
/* test hook: list sizes (summed over tiles) for n k-mers */
int sg_index_list_sizes(const sg_index* ix, const uint32_t* kmers, uint32_t n, uint64_t* sizes);
/* test hook: copy the posting list of one k-mer (ids ascending) into ids[cap]; *n = list length */
int sg_index_list(const sg_index* ix, uint32_t kmer, uint32_t* ids, uint64_t cap, uint64_t* n);

/* ---- host-buffer entry points (what the reference-side binding calls) -------------------------- */
/* search::find for nq queries. scores/ids: nq rows of `max` entries, rank order (score desc, id desc);
 * nres[q] = min(max, N). */
int sg_find_batch(sg_index* ix, const uint8_t* qmasks, const uint64_t* qoff, uint32_t nq, uint32_t max,
                  int16_t* scores, uint32_t* ids, uint32_t* nres);
/* --turn orientation check (famfinder::turn_check, src/famfinder.cpp:344-378): find(max = 1) on the query, its
 * reverse, its complement and its reverse complement (mode 2 = "all"; mode 1 = "revcomp" skips the middle two).
 * turn[q] = 0 none, 1 reversed, 2 complemented, 3 reversed and complemented: the first orientation with the strictly
 * largest top score. The caller applies it to its sequence (cseq::reverse / complement) before family finding. */
int sg_turn_batch(sg_index* ix, const uint8_t* qmasks, const uint64_t* qoff, uint32_t nq, int mode, int32_t* turn);
/* famfinder stage: family per query in rank order. fam_ids/fam_scores: nq rows of fam_stride entries;
 * fam_n[q] = family size, or -1 when fewer than fs_req relatives remain. exclude_ids (optional): id of the
 * reference carrying the query's name (for --fs-leave-query-out), -1 for none. */
int sg_family_batch(sg_index* ix, const uint8_t* qmasks, const uint64_t* qoff, uint32_t nq,
                    const int64_t* exclude_ids, const sg_fam_params* fp, uint32_t fam_stride, uint32_t* fam_ids,
                    float* fam_scores, int32_t* fam_n);
/* aligner stage with the family given (ids into the index, rank order; fam_off: nq+1 offsets).
 * out_cols/out_masks are indexed like qmasks (query q's bases at qoff[q]..). */
int sg_align_batch(sg_index* ix, const uint8_t* qmasks, const uint64_t* qoff, uint32_t nq, const uint32_t* fam_ids,
                   const uint64_t* fam_off, const sg_align_params* ap, uint32_t* out_cols, uint8_t* out_masks,
                   sg_align_result* results);
/* sg_family_batch for queries that carry positions (a pre-aligned input, as the reference's evaluation runs with
 * --fs-msc-max < 1 use): qcols[j] = position of base j, strictly increasing inside a query; null = sg_family_batch. */
int sg_family_batch_aligned(sg_index* ix, const uint8_t* qmasks, const uint32_t* qcols, const uint64_t* qoff, uint32_t nq,
                            const int64_t* exclude_ids, const sg_fam_params* fp, uint32_t fam_stride, uint32_t* fam_ids,
                            float* fam_scores, int32_t* fam_n);
/* famfinder + aligner for a batch: host in, host out. */
int sg_run_batch(sg_index* ix, const uint8_t* qmasks, const uint64_t* qoff, uint32_t nq, const int64_t* exclude_ids,
                 const sg_fam_params* fp, const sg_align_params* ap, uint32_t* out_cols, uint8_t* out_masks,
                 sg_align_result* results);

/* ---- --search stage and the sequence comparator ------------------------------------------------------ */
/* search_filter options, defaults = reference defaults (src/search_filter.cpp:96-133, src/cseq_comparator.cpp:432-462) */
typedef struct sg_search_params {
    uint32_t kmer_candidates;  /* --search-kmer-candidates 1000 */
    uint32_t max_result;       /* --search-max-result 10 */
    float min_sim;             /* --search-min-sim 0.7 */
    int32_t ignore_super;      /* --search-ignore-super (the reference's partition + erase KEEPS the candidates that contain
                                * the query, src/search_filter.cpp:313-316: reproduced) */
    int32_t iupac;             /* --search-iupac: 0 optimistic, 1 pessimistic, 2 exact */
    int32_t correction;        /* --search-correction: 0 none, 1 jc */
    int32_t cover;             /* --search-cover: 0 abs, 1 query, 2 target, 3 overlap, 4 all, 5 average, 6 min, 7 max, 8 nogap */
    int32_t filter_lowercase;  /* --search-filter-lowercase */
} sg_search_params;
void sg_default_search_params(sg_search_params* p);
/* rank[i] = position of reference i's name in ascending lexicographic order: search::result_item orders equal scores by
 * name (src/search.h:56-68). n = 0 clears them (ties then go by reference id). */
int sg_index_set_name_ranks(sg_index* ix, const uint32_t* rank, uint32_t n);
/* cseq_comparator::operator() (src/cseq_comparator.cpp:209-293) of aligned sequence q (bases amasks, strictly increasing
 * alignment columns acols, rows aoff[q]..aoff[q+1]) against the reference rows ref_ids[ref_off[q]..ref_off[q+1]).
 * out[ref_off[nq]]: match count / cover-rule base as float; 0/0 (nothing to compare) is NaN as in the reference. */
int sg_identity_batch(sg_index* ix, const uint8_t* amasks, const uint32_t* acols, const uint64_t* aoff, uint32_t nq,
                      const uint32_t* ref_ids, const uint64_t* ref_off, int iupac, int correction, int cover,
                      int filter_lowercase, float* out);
/* search_filter::operator() (src/search_filter.cpp:244-330, the k-mer branch: --search-all is not offered) for a batch of
 * ALIGNED sequences: k-mer search for kmer_candidates references, identity of every candidate, the max_result best by
 * (score, name) descending with score > min_sim. out_ids / out_scores [nq * max_result], out_n [nq] (0 for sequences
 * shorter than 20 bases). kmer_candidates must fit the two-level top-k merge (at the default 1000: up to 256 index tiles of
 * 49152 references). With correction = jc the logarithm and
 * the final selection run on the host (the reference's double log), everything before it on the device. */
int sg_search_batch(sg_index* ix, const uint8_t* amasks, const uint32_t* acols, const uint64_t* aoff, uint32_t nq,
                    const sg_search_params* sp, uint32_t* out_ids, float* out_scores, uint32_t* out_n);

/* ---- session: the same stages with the batch resident in HBM (used by bench.py for the device-only
 * number and by the host-buffer calls above internally) ------------------------------------------- */
int sg_session_create(sg_index* ix, uint32_t max_queries, uint64_t max_bases, sg_session** out);
void sg_session_destroy(sg_session* s);
int sg_session_upload(sg_session* s, const uint8_t* qmasks, const uint64_t* qoff, uint32_t nq,
                      const int64_t* exclude_ids);
/* positions of the uploaded batch's bases (same layout as qmasks, starting at the batch's first base): needed by
 * --fs-msc-max < 1 only; forgotten at the next upload */
int sg_session_set_query_columns(sg_session* s, const uint32_t* qcols);
int sg_session_find(sg_session* s, uint32_t max);
/* orientation check on the resident batch; the queries are left in the chosen orientation (turn may be null) */
int sg_session_turn(sg_session* s, int mode, int32_t* turn);
int sg_session_family(sg_session* s, const sg_fam_params* fp);
int sg_session_set_family(sg_session* s, const uint32_t* fam_ids, const uint64_t* fam_off);
int sg_session_align(sg_session* s, const sg_align_params* ap);
/* sg_session_family + sg_session_align in one call (what sg_run_batch runs) */
int sg_session_run(sg_session* s, const sg_fam_params* fp, const sg_align_params* ap);
int sg_session_sync(sg_session* s);
int sg_session_download_find(sg_session* s, int16_t* scores, uint32_t* ids, uint32_t* nres);
int sg_session_download_family(sg_session* s, uint32_t fam_stride, uint32_t* fam_ids, float* fam_scores,
                               int32_t* fam_n);
int sg_session_download_align(sg_session* s, uint32_t* out_cols, uint8_t* out_masks, sg_align_result* results);

/* measurement hooks (CUDA events recorded on the session's own stream around every stage) */
typedef struct sg_stage_stats {
    float ms_find;      /* k-mer search + top-k kernels */
    float ms_family;    /* family selection kernel */
    float ms_graph;     /* family-graph construction + DP plan */
    float ms_dp;        /* mesh DP kernel */
    float ms_backtrack; /* backtrack + gap placement */
    uint64_t cells;     /* DP cells computed (sum of V*Lq) */
    uint64_t postings;  /* posting entries scanned by the search kernel */
    uint64_t kernel_launches;
} sg_stage_stats;
int sg_session_stats(sg_session* s, sg_stage_stats* st, int reset);
/* device clock over a run of stage calls: stop=0 records the start event on the session's stream, stop=1 records the
 * end event, waits for it and returns the elapsed milliseconds between the two (cudaEventElapsedTime) */
int sg_session_timer(sg_session* s, int stop, float* ms);

/* test hooks: device graph / traceback dumps for one query of the session after sg_session_align */
int sg_session_dump_graph(sg_session* s, uint32_t q, uint32_t cap_nodes, uint32_t cap_edges, uint32_t* V,
                          uint32_t* E, uint32_t* col, uint8_t* mask, float* weight, uint32_t* pred_off,
                          uint32_t* preds);

#ifdef __cplusplus
}
#endif
#endif
