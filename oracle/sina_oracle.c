/* TEST INFRASTRUCTURE ONLY -- see sina_oracle.h. CPU restatement of SINA's per-query hot path.
 * Reference citations are file:line in the SINA source tree (commit b0763146). */
#include "sina_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ------------------------------------------------------------------ base encoding */
/* src/aligned_base.cpp:70-107: A=1 G=2 C=4 T/U=8, IUPAC = ORs, lowercase adds 0x10 */
int so_char_to_mask(int c) {
    int lc = 0, m;
    if (c == '-' || c == '.') return 0;
    if (c >= 'a' && c <= 'z') { lc = SO_BASEM_LC; c -= 32; }
    switch (c) {
        case 'A': m = 1; break;
        case 'G': m = 2; break;
        case 'C': m = 4; break;
        case 'T': case 'U': m = 8; break;
        case 'R': m = 1 | 2; break;
        case 'Y': m = 4 | 8; break;
        case 'K': m = 2 | 8; break;
        case 'M': m = 1 | 4; break;
        case 'S': m = 2 | 4; break;
        case 'W': m = 1 | 8; break;
        case 'B': m = 2 | 8 | 4; break;
        case 'D': m = 2 | 1 | 8; break;
        case 'H': m = 1 | 4 | 8; break;
        case 'V': m = 2 | 4 | 1; break;
        case 'N': m = 15; break;
        default: return -1; /* bad_character_exception, src/aligned_base.h:76-82 */
    }
    return m | lc;
}

/* src/aligned_base.cpp:109-121 */
int so_mask_to_char(int mask, int dna) {
    static const char rna[] = ".AGRCMSVUWKDYHBN.agrcmsvuwkdyhbn";
    int c = rna[mask & 31];
    if (dna && c == 'U') c = 'T';
    if (dna && c == 'u') c = 't';
    return c;
}

/* cseq_base::append(const char*) src/cseq.cpp:63-77 */
int64_t so_encode_aligned(const char* str, uint8_t* masks, uint32_t* cols, uint64_t cap, uint32_t* width) {
    uint32_t w = 0;
    uint64_t n = 0;
    for (uint64_t i = 0; str[i]; i++) {
        char c = str[i];
        if (c == ' ' || c == '\t' || c == '\n' || c == '\r') continue;
        if (c != '-' && c != '.') {
            int m = so_char_to_mask((unsigned char)c);
            if (m <= 0) return -1 - (int64_t)i;
            if (n < cap) { masks[n] = (uint8_t)m; if (cols) cols[n] = w; }
            n++;
        }
        w++;
    }
    if (width) *width = w;
    return (int64_t)n;
}

/* cseq_base::getAligned(nodots=true, dna) src/cseq.cpp:135-174 */
void so_render_aligned(const uint8_t* masks, const uint32_t* cols, uint32_t n, uint32_t width, int dna, char* out) {
    uint32_t cursor = 0;
    uint64_t o = 0;
    for (uint32_t i = 0; i < n; i++) {
        uint32_t pos = cols[i];
        while (cursor < pos) { out[o++] = '-'; cursor++; }
        cursor = pos;
        out[o++] = (char)so_mask_to_char(masks[i], dna);
        cursor++;
    }
    while (cursor < width) { out[o++] = '-'; cursor++; }
    out[o] = 0;
}

/* ------------------------------------------------------------------ k-mers */
typedef struct { uint32_t* tab; uint32_t cap; } u32set; /* open addressing, key+1 stored, 0 = empty */
static void set_init(u32set* s, uint32_t n) {
    uint32_t c = 16;
    while (c < 2 * n + 2) c <<= 1;
    s->cap = c;
    s->tab = (uint32_t*)calloc(c, 4);
}
static void set_clear(u32set* s) { memset(s->tab, 0, (size_t)s->cap * 4); }
static int set_insert(u32set* s, uint32_t key) { /* 1 if newly inserted */
    uint32_t h = (key * 2654435761u) & (s->cap - 1);
    for (;;) {
        if (s->tab[h] == 0) { s->tab[h] = key + 1; return 1; }
        if (s->tab[h] == key + 1) return 0;
        h = (h + 1) & (s->cap - 1);
    }
}

/* src/kmer.h:54-83 (generator::push/good), :110-125 (prefix_filter), :128-151 (unique_filter),
 * :174-202 (iterable::iterator: the k-mer ending on the last base is never yielded). */
static int64_t kmers_impl(const uint8_t* masks, uint32_t n, int k, int mode, uint32_t* out, uint64_t cap, u32set* seen) {
    uint32_t kmask = (k >= 16) ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    uint32_t pmask = 3u << (2 * (k - 1));
    uint32_t val = (mode == 0) ? 1 : 0, good = 0;
    int64_t cnt = 0;
    if (seen) set_clear(seen);
    for (uint32_t i = 0; i < n; i++) {
        uint32_t b = masks[i] & 0xf;
        if (__builtin_popcount(b) > 1) {
            good = 0;
        } else {
            good++;
            val = ((val << 2) & kmask) + (uint32_t)__builtin_ctz(b);
        }
        int ok = good >= (uint32_t)k;
        if (ok && (mode & 2)) ok = (val & pmask) == 0; /* first base A (BASE_A=0) */
        if (ok && (mode & 1)) ok = set_insert(seen, val);
        if (ok && i != n - 1) {
            if ((uint64_t)cnt < cap && out) out[cnt] = val;
            cnt++;
        }
    }
    return cnt;
}

int64_t so_kmers(const uint8_t* masks, uint32_t n, int k, int mode, uint32_t* out, uint64_t cap) {
    u32set s;
    int64_t r;
    if (k < 1 || k > 16) return -1;
    if (mode & 1) { set_init(&s, n); r = kmers_impl(masks, n, k, mode, out, cap, &s); free(s.tab); return r; }
    return kmers_impl(masks, n, k, mode, out, cap, NULL);
}

/* ------------------------------------------------------------------ index build (src/kmer_search.cpp:152-276) */
so_index* so_index_build(uint32_t N, const uint8_t* masks, const uint64_t* off, int k, int nofast) {
    so_index* ix = (so_index*)calloc(1, sizeof(so_index));
    uint32_t maxlen = 0;
    int mode = nofast ? 1 : 3; /* unique_kmers : unique_prefix_kmers(A), :164-177 */
    ix->N = N; ix->k = k; ix->nofast = nofast;
    ix->n_kmers = 1ull << (2 * k);
    ix->list_off = (uint64_t*)calloc(ix->n_kmers + 1, 8);
    for (uint32_t i = 0; i < N; i++) if (off[i + 1] - off[i] > maxlen) maxlen = (uint32_t)(off[i + 1] - off[i]);
    uint32_t* buf = (uint32_t*)malloc(((size_t)maxlen + 1) * 4);
    u32set seen;
    set_init(&seen, maxlen);
    for (uint32_t i = 0; i < N; i++) {
        int64_t c = kmers_impl(masks + off[i], (uint32_t)(off[i + 1] - off[i]), k, mode, buf, maxlen, &seen);
        for (int64_t j = 0; j < c; j++) ix->list_off[buf[j] + 1]++;
    }
    for (uint64_t i = 0; i < ix->n_kmers; i++) ix->list_off[i + 1] += ix->list_off[i];
    ix->postings = (uint32_t*)malloc((ix->list_off[ix->n_kmers] + 1) * 4);
    uint64_t* cur = (uint64_t*)malloc(ix->n_kmers * 8);
    memcpy(cur, ix->list_off, ix->n_kmers * 8);
    for (uint32_t i = 0; i < N; i++) { /* ids ascend within each list (push_back(i), :169,176) */
        int64_t c = kmers_impl(masks + off[i], (uint32_t)(off[i + 1] - off[i]), k, mode, buf, maxlen, &seen);
        for (int64_t j = 0; j < c; j++) ix->postings[cur[buf[j]]++] = i;
    }
    /* lists longer than N/2 are stored inverted by the reference (:264-266); counting through the
     * complement and adding `offset` (src/idset.h:315-337, kmer_search.cpp:392-408) yields the same
     * totals as plain counting, so the CSR keeps plain lists. */
    free(cur); free(buf); free(seen.tab);
    return ix;
}

void so_index_free(so_index* ix) {
    if (!ix) return;
    free(ix->list_off); free(ix->postings); free(ix);
}

/* rank order of std::greater<pair<int16,int>> (src/kmer_search.cpp:412): score desc, then id desc */
typedef struct { int16_t score; uint32_t id; } rank_t;
static inline int rank_before(rank_t a, rank_t b) { return a.score > b.score || (a.score == b.score && a.id > b.id); }

static void heap_sift(rank_t* h, uint32_t n, uint32_t i) { /* "worst on top" heap */
    for (;;) {
        uint32_t l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && rank_before(h[w], h[l])) w = l;
        if (r < n && rank_before(h[w], h[r])) w = r;
        if (w == i) return;
        rank_t t = h[i]; h[i] = h[w]; h[w] = t;
        i = w;
    }
}

/* top-`max` of scores[0..N) in rank order */
static void top_ranks(const int16_t* scores, uint32_t N, uint32_t max, rank_t* out) {
    uint32_t n = 0;
    for (uint32_t i = 0; i < N; i++) {
        rank_t r = { scores[i], i };
        if (n < max) {
            out[n++] = r;
            if (n == max) for (int64_t j = (int64_t)n / 2 - 1; j >= 0; j--) heap_sift(out, n, (uint32_t)j);
        } else if (rank_before(r, out[0])) {
            out[0] = r;
            heap_sift(out, n, 0);
        }
    }
    if (n < max) for (int64_t j = (int64_t)n / 2 - 1; j >= 0; j--) heap_sift(out, n, (uint32_t)j);
    for (uint32_t m = n; m > 1; m--) { /* sort_heap: worst to the back */
        rank_t t = out[0]; out[0] = out[m - 1]; out[m - 1] = t;
        heap_sift(out, m - 1, 0);
    }
}

/* scores[i] = sum over query k-mers (duplicates counted) of [i in list] (src/kmer_search.cpp:389-409) */
static uint64_t count_scores(const so_index* ix, const uint8_t* q, uint32_t qlen, int16_t* scores, uint32_t* kbuf) {
    int mode = ix->nofast ? 0 : 2; /* all_kmers : prefix_kmers(A), :391,397 */
    uint64_t P = 0;
    int64_t c = kmers_impl(q, qlen, ix->k, mode, kbuf, qlen, NULL);
    memset(scores, 0, (size_t)ix->N * 2);
    for (int64_t j = 0; j < c; j++) {
        uint64_t a = ix->list_off[kbuf[j]], b = ix->list_off[kbuf[j] + 1];
        for (uint64_t e = a; e < b; e++) scores[ix->postings[e]]++; /* int16 wraps like idset::inc_t */
        P += b - a;
    }
    return P;
}

/* kmer_search::impl::find src/kmer_search.cpp:365-420 */
uint32_t so_find(const so_index* ix, const uint8_t* q, uint32_t qlen, uint32_t max, int16_t* scores, uint32_t* ids,
                 uint64_t* postings) {
    if (max > ix->N) max = ix->N;
    if (postings) *postings = 0;
    if (max == 0) return 0;
    int16_t* sc = (int16_t*)malloc((size_t)ix->N * 2);
    uint32_t* kbuf = (uint32_t*)malloc(((size_t)qlen + 1) * 4);
    rank_t* top = (rank_t*)malloc((size_t)max * sizeof(rank_t));
    uint64_t P = count_scores(ix, q, qlen, sc, kbuf);
    top_ranks(sc, ix->N, max, top);
    for (uint32_t i = 0; i < max; i++) { scores[i] = top[i].score; ids[i] = top[i].id; }
    if (postings) *postings = P;
    free(sc); free(kbuf); free(top);
    return max;
}

/* famfinder::impl::turn_check src/famfinder.cpp:344-378: top find() score of the query as is, reversed
 * (cseq_base::reverse, src/cseq.cpp:284-289), complemented (base_iupac::complement, src/aligned_base.h:117-124)
 * and reverse-complemented; the first orientation whose score is strictly above the running maximum (from 0) wins.
 * all == 0 ("revcomp") leaves scores 1 and 2 at 0. scores4 (optional) receives the four scores. */
static uint8_t mask_complement(uint8_t m) {
    return (uint8_t)(((m & 2u) << 1) | ((m & 4u) >> 1) | ((m & 1u) << 3) | ((m & 8u) >> 3) | (m & 16u));
}
int so_turn_check(const so_index* ix, const uint8_t* q, uint32_t qlen, int all, int32_t* scores4) {
    uint8_t* t = (uint8_t*)malloc(qlen ? qlen : 1);
    int32_t score[4] = {0, 0, 0, 0};
    int16_t sc;
    uint32_t id;
    if (so_find(ix, q, qlen, 1, &sc, &id, NULL)) score[0] = sc;
    for (uint32_t i = 0; i < qlen; i++) t[i] = q[qlen - 1 - i];                      /* turn.reverse() */
    if (all) {
        if (so_find(ix, t, qlen, 1, &sc, &id, NULL)) score[1] = sc;
        uint8_t* c = (uint8_t*)malloc(qlen ? qlen : 1);
        for (uint32_t i = 0; i < qlen; i++) c[i] = mask_complement(q[i]);            /* comp.complement() */
        if (so_find(ix, c, qlen, 1, &sc, &id, NULL)) score[2] = sc;
        free(c);
    }
    for (uint32_t i = 0; i < qlen; i++) t[i] = mask_complement(t[i]);                /* turn.complement() */
    if (so_find(ix, t, qlen, 1, &sc, &id, NULL)) score[3] = sc;
    free(t);
    int32_t max = 0;
    int best = 0;
    for (int i = 0; i < 4; i++)
        if (max < score[i]) { max = score[i]; best = i; }
    if (scores4) memcpy(scores4, score, sizeof(score));
    return best;
}

/* ------------------------------------------------------------------ family selection */
typedef struct {
    const so_fam_params* p;
    const uint64_t* off;
    int64_t exclude_id;
    uint32_t have, have_full;
} fam_state;

/* the `remove` lambda of famfinder::impl::match, src/famfinder.cpp:578-589 (with :508-576).
 * remove_superstring is hard-wired off (noid=false, :503); remove_similar needs cseq_comparator and
 * is a no-op at the default fs_msc_max=2 (identity <= 1); range-cover terms are off (fs_cover_gene=0). */
static int fam_remove(fam_state* st, uint32_t id, float score) {
    const so_fam_params* p = st->p;
    uint32_t len = (uint32_t)(st->off[id + 1] - st->off[id]);
    int is_full = len >= p->fs_full_len;
    if (len < p->fs_min_len) return 1;
    if (p->leave_query_out && st->exclude_id == (int64_t)id) return 1;
    if (st->have >= p->fs_min && (st->have >= p->fs_max || !(score < p->fs_msc)) &&
        !(p->fs_req_full && st->have_full < p->fs_req_full && is_full))
        return 1;
    st->have++;
    if (p->fs_req_full && is_full) st->have_full++;
    return 0;
}

/* too_few_gaps, src/famfinder.cpp:474-480 */
static int fam_too_few_gaps(const uint64_t* off, const uint32_t* cols, uint32_t id, uint32_t req) {
    uint32_t len = (uint32_t)(off[id + 1] - off[id]);
    if (len == 0) return 1;
    return cols[off[id + 1] - 1] - len + 1 < req; /* unsigned arithmetic as in the reference */
}

int so_family_from_ranked(const uint32_t* cand_ids, const int16_t* cand_scores, uint32_t n_cand, uint32_t n_total,
                          const uint64_t* off, const uint32_t* cols, int64_t exclude_id, const so_fam_params* p,
                          uint32_t* ids, float* scores, uint32_t cap) {
    fam_state st = { p, off, exclude_id, 0, 0 };
    uint64_t max_results = (uint64_t)p->fs_max + 1;
    uint32_t n = 0;
    if (!(st.have < p->fs_max || st.have_full < p->fs_req_full)) return 0;
    for (;;) { /* src/famfinder.cpp:591-608 */
        uint32_t window = (uint32_t)(max_results < n_total ? max_results : n_total);
        if (window > n_cand) return -2;
        if (window == 0) return p->fs_req > 0 ? -1 : 0;
        st.have = st.have_full = 0;
        n = 0;
        for (uint32_t i = 0; i < window; i++) {
            if (!fam_remove(&st, cand_ids[i], (float)cand_scores[i])) {
                if (n < cap) { ids[n] = cand_ids[i]; scores[n] = (float)cand_scores[i]; }
                n++;
            }
        }
        if (max_results >= n_total) break;
        if (!(st.have < p->fs_max || st.have_full < p->fs_req_full)) break;
        max_results *= 10;
    }
    if (n > cap) n = cap;
    if (p->fs_req_gaps != 0) {
        uint32_t m = 0;
        for (uint32_t i = 0; i < n; i++)
            if (!fam_too_few_gaps(off, cols, ids[i], p->fs_req_gaps)) { ids[m] = ids[i]; scores[m] = scores[i]; m++; }
        n = m;
    }
    if (n < p->fs_req) return -1; /* src/famfinder.cpp:486-491 */
    return (int)n;
}

int so_family(const so_index* ix, const uint64_t* off, const uint32_t* cols, const uint8_t* q, uint32_t qlen,
              int64_t exclude_id, const so_fam_params* p, uint32_t* ids, float* scores, uint32_t cap,
              uint64_t* postings) {
    int16_t* sc = (int16_t*)malloc((size_t)ix->N * 2 + 2);
    uint32_t* kbuf = (uint32_t*)malloc(((size_t)qlen + 1) * 4);
    uint64_t P = count_scores(ix, q, qlen, sc, kbuf);
    uint64_t window = (uint64_t)p->fs_max + 1;
    int r;
    if (postings) *postings = P;
    for (;;) {
        uint32_t max = (uint32_t)(window < ix->N ? window : ix->N);
        rank_t* top = (rank_t*)malloc(((size_t)max + 1) * sizeof(rank_t));
        uint32_t* cid = (uint32_t*)malloc(((size_t)max + 1) * 4);
        int16_t* csc = (int16_t*)malloc(((size_t)max + 1) * 2);
        top_ranks(sc, ix->N, max, top);
        for (uint32_t i = 0; i < max; i++) { cid[i] = top[i].id; csc[i] = top[i].score; }
        r = so_family_from_ranked(cid, csc, max, ix->N, off, cols, exclude_id, p, ids, scores, cap);
        free(top); free(cid); free(csc);
        if (r != -2) break;
        window *= 10;
    }
    free(sc); free(kbuf);
    return r;
}

/* ------------------------------------------------------------------ family graph */
typedef struct { uint32_t* v; uint32_t n, cap; } u32vec;
static void vec_push(u32vec* a, uint32_t x) {
    if (a->n == a->cap) { a->cap = a->cap ? a->cap * 2 : 256; a->v = (uint32_t*)realloc(a->v, (size_t)a->cap * 4); }
    a->v[a->n++] = x;
}
static int cmp_u64(const void* a, const void* b) {
    uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : x > y;
}

/* mseq::mseq src/mseq.cpp:47-118; dag::insert/link src/graph.h:332-357; sort + reduce_edges
 * src/align.cpp:401-402, src/graph.h:451-488. Node id = creation order (column-major, then first family
 * row to bring a new (column, IUPAC char)); preds ascend by id; first/last = sentinel lists. */
so_graph* so_graph_build(const uint32_t* fam, uint32_t F, const uint8_t* masks, const uint32_t* cols,
                         const uint64_t* off, uint32_t W, float fs_weight) {
    so_graph* g = (so_graph*)calloc(1, sizeof(so_graph));
    uint64_t* cur = (uint64_t*)malloc(((size_t)F + 1) * 8);
    uint64_t* end = (uint64_t*)malloc(((size_t)F + 1) * 8);
    int64_t* last = (int64_t*)malloc(((size_t)F + 1) * 8);
    u32vec ncol = {0}, nmask = {0}, ncount = {0}, efrom = {0}, eto = {0};
    g->W = W;
    for (uint32_t j = 0; j < F; j++) { cur[j] = off[fam[j]]; end[j] = off[fam[j] + 1]; last[j] = -1; }
    for (;;) {
        uint32_t i = 0xFFFFFFFFu; /* next column holding any base (min_next, :76-84,106-108) */
        int64_t slot[32];
        uint32_t first_new = ncol.n;
        for (uint32_t j = 0; j < F; j++) if (cur[j] < end[j] && cols[cur[j]] < i) i = cols[cur[j]];
        if (i == 0xFFFFFFFFu || i >= W) break;
        for (int c = 0; c < 32; c++) slot[c] = -1;
        for (uint32_t j = 0; j < F; j++) {
            if (cur[j] < end[j] && cols[cur[j]] == i) {
                uint8_t b = masks[cur[j]] & 31; /* node key = IUPAC char incl. case (:92) */
                if (slot[b] < 0) { /* :93-94 */
                    slot[b] = ncol.n;
                    vec_push(&ncol, i); vec_push(&nmask, b); vec_push(&ncount, 1);
                } else {
                    ncount.v[slot[b]]++; /* weight += 1 (:96-97) */
                }
                if (last[j] >= 0) { vec_push(&efrom, (uint32_t)last[j]); vec_push(&eto, (uint32_t)slot[b]); } /* :99-101 */
                last[j] = slot[b];
                cur[j]++;
            }
        }
        (void)first_new;
    }
    uint32_t V = ncol.n, E0 = efrom.n;
    g->V = V;
    g->col = (uint32_t*)malloc(((size_t)V + 1) * 4);
    g->mask = (uint8_t*)malloc((size_t)V + 1);
    g->weight = (float*)malloc(((size_t)V + 1) * 4);
    for (uint32_t v = 0; v < V; v++) {
        g->col[v] = ncol.v[v];
        g->mask[v] = (uint8_t)nmask.v[v];
        /* node->weight = 1.0/(weight+1) + weight * (node->weight/num_seqs)  (:111-116): double + float mix */
        float cnt = (float)ncount.v[v];
        float b = fs_weight * (cnt / (float)F);
        g->weight[v] = (float)(1.0 / (double)(fs_weight + 1.0f) + (double)b);
    }
    /* reduce_edges: sort + unique per node */
    uint64_t* ek = (uint64_t*)malloc(((size_t)E0 + 1) * 8);
    for (uint32_t e = 0; e < E0; e++) ek[e] = ((uint64_t)eto.v[e] << 32) | efrom.v[e];
    qsort(ek, E0, 8, cmp_u64);
    uint32_t E = 0;
    for (uint32_t e = 0; e < E0; e++) if (e == 0 || ek[e] != ek[e - 1]) ek[E++] = ek[e];
    g->E = E;
    g->pred_off = (uint32_t*)calloc((size_t)V + 2, 4);
    g->preds = (uint32_t*)malloc(((size_t)E + 1) * 4);
    uint8_t* has_succ = (uint8_t*)calloc((size_t)V + 1, 1);
    for (uint32_t e = 0; e < E; e++) {
        uint32_t to = (uint32_t)(ek[e] >> 32), from = (uint32_t)ek[e];
        g->pred_off[to + 1]++;
        g->preds[e] = from;
        has_succ[from] = 1;
    }
    for (uint32_t v = 0; v < V; v++) g->pred_off[v + 1] += g->pred_off[v];
    g->first = (uint32_t*)malloc(((size_t)V + 1) * 4);
    g->last = (uint32_t*)malloc(((size_t)V + 1) * 4);
    for (uint32_t v = 0; v < V; v++) {
        if (g->pred_off[v + 1] == g->pred_off[v]) g->first[g->n_first++] = v;
        if (!has_succ[v]) g->last[g->n_last++] = v;
    }
    free(cur); free(end); free(last); free(ncol.v); free(nmask.v); free(ncount.v); free(efrom.v); free(eto.v);
    free(ek); free(has_succ);
    return g;
}

void so_graph_free(so_graph* g) {
    if (!g) return;
    free(g->col); free(g->mask); free(g->weight); free(g->pred_off); free(g->preds); free(g->first); free(g->last);
    free(g);
}

/* ------------------------------------------------------------------ positional weights */
/* scoring_scheme_weighted (src/scoring_schemes.h:166-241, chosen when the alignment statistics have a width,
 * src/align.cpp:409-415): gap costs and match scores are multiplied by the weight of an alignment column. The
 * reference reads weights[position + 1 + insertion length] without a bounds check; here an index past the end reads
 * the last weight (the only place where this restatement has to DEFINE something the reference leaves undefined). */
static const float* g_colw = NULL;
static uint32_t g_ncolw = 0;
void so_set_column_weights(const float* w, uint32_t n) { g_colw = n ? w : NULL; g_ncolw = n; }
static inline float colw(uint32_t i) { return g_colw[i < g_ncolw ? i : g_ncolw - 1]; }

/* ------------------------------------------------------------------ mesh DP */
/* compute() src/mesh.h:509-528 over compute_node_simple::calc :453-502 with transition_simple
 * :305-374 and scoring_scheme_simple src/scoring_schemes.h:102-164. Scores are minimised. */
so_mesh* so_mesh_compute(const so_graph* g, const uint8_t* q, uint32_t L, const so_align_params* p) {
    so_mesh* M = (so_mesh*)calloc(1, sizeof(so_mesh));
    uint64_t n = (uint64_t)g->V * L;
    const float ms = -p->match_score, mms = -p->mismatch_score; /* src/align.cpp:406-407 */
    const float gp = p->gap_penalty, gpe = p->gap_ext_penalty;
    M->V = g->V; M->L = L;
    M->value_midx = (uint32_t*)malloc((n + 1) * 4); M->value_sidx = (uint32_t*)malloc((n + 1) * 4);
    M->gapm_idx = (uint32_t*)malloc((n + 1) * 4);   M->gaps_idx = (uint32_t*)malloc((n + 1) * 4);
    M->value = (float*)malloc((n + 1) * 4); M->gapm_val = (float*)malloc((n + 1) * 4);
    M->gaps_val = (float*)malloc((n + 1) * 4);
    /* --insertion forbid: transition_aspace_aware (src/mesh.h:377-438, chosen in src/align.cpp:466-468): a cell also
     * carries gaps_max, the number of insertions still accommodated by the free columns between the node and its
     * nearest successor (max_insert, compute_node_simple::calc src/mesh.h:480-484) */
    const int forbid = p->insertion == 1;
    uint32_t* gaps_max = forbid ? (uint32_t*)calloc(n + 1, 4) : NULL;
    uint32_t* min_mpos = forbid ? (uint32_t*)malloc(((size_t)g->V + 1) * 4) : NULL;
    if (forbid) {
        for (uint32_t m = 0; m < g->V; m++) min_mpos[m] = 1000000u;
        for (uint32_t m = 0; m < g->V; m++)
            for (uint32_t e = g->pred_off[m]; e < g->pred_off[m + 1]; e++)
                if (g->col[m] < min_mpos[g->preds[e]]) min_mpos[g->preds[e]] = g->col[m];
    }
    for (uint32_t m = 0; m < g->V; m++) {
        uint32_t pb = g->pred_off[m], pe = g->pred_off[m + 1];
        float w = g->weight[m];
        /* weighted scheme: deletions cost gap * weights[col(m)] (scoring_schemes.h:203-222), matches score
         * (match * weights[col(m)]) * weight(m) (:224-232), all with the TARGET node's column */
        const float wm = g_colw ? colw(g->col[m]) : 1.0f;
        const float gpm = g_colw ? gp * wm : gp, gpem = g_colw ? gpe * wm : gpe;
        const uint32_t smax = forbid ? (uint32_t)(int)(min_mpos[m] - g->col[m] - 1) : 0;
        for (uint32_t s = 0; s < L; s++) {
            uint64_t o = (uint64_t)m * L + s;
            float value, gapm_val, gaps_val;
            uint32_t value_midx = 0, value_sidx = 0, gapm_idx = 0, gaps_idx = 0;
            if (pb == pe || s == 0) value = gapm_val = gaps_val = 1.0f;  /* init_edge :294-297,469-470 */
            else value = gapm_val = gaps_val = 1000000.0f;               /* init :298-301 */
            for (uint32_t e = pb; e < pe; e++) { /* deletion :475-478 -> :305-330 */
                uint32_t mi = g->preds[e];
                uint64_t so = (uint64_t)mi * L + s;
                float v = M->value[so] + gpm;
                float gv = M->gapm_val[so] + gpem;
                uint32_t midx = mi;
                if (v < gv) { gapm_val = v; gapm_idx = mi; }
                else { gapm_val = gv; gapm_idx = M->gapm_idx[so]; v = gv; midx = M->gapm_idx[so]; }
                if (v < value) { value = v; value_midx = midx; value_sidx = s; }
            }
            if (s > 0) { /* insertion :486-490 -> :332-358 */
                uint64_t so = o - 1;
                int evaluated = 1;
                /* insertion costs: gap * weights[col(m) + 1] to open, gapext * weights[col(m) + 1 + run length] to
                 * extend (scoring_schemes.h:178-201: the column the inserted base will be placed in) */
                const float gpi = g_colw ? gp * colw(g->col[m] + 1) : gp;
                const float gpei = g_colw ? gpe * colw(g->col[m] + 1 + ((s - 1) - M->gaps_idx[so])) : gpe;
                if (!forbid) {
                    if (M->gaps_val[so] != M->value[so]) { gaps_val = M->value[so] + gpi; gaps_idx = s - 1; }
                    else { gaps_val = M->gaps_val[so] + gpei; gaps_idx = M->gaps_idx[so]; }
                } else if (smax < 1) {
                    evaluated = 0;                                        /* can't insert :412-414 */
                } else if (M->gaps_val[so] != M->value[so]) {              /* opening gap :416-420 */
                    gaps_val = M->value[so] + gpi; gaps_idx = s - 1; gaps_max[o] = smax - 1;
                } else if (gaps_max[so] > 0) {                             /* extending gap :421-426 */
                    gaps_val = M->gaps_val[so] + gpei; gaps_idx = M->gaps_idx[so]; gaps_max[o] = gaps_max[so] - 1;
                } else {
                    evaluated = 0;                                        /* :427-429 */
                }
                if (evaluated && gaps_val <= value) { value = gaps_val; value_sidx = gaps_idx; value_midx = m; }
                for (uint32_t e = pb; e < pe; e++) { /* match :492-500 -> :360-374 */
                    uint32_t mi = g->preds[e];
                    uint64_t po = (uint64_t)mi * L + (s - 1);
                    float sc = (g->mask[m] & q[s] & 0xf) ? ms : mms;
                    if (g_colw) sc = sc * wm;                              /* scoring_schemes.h:224-232 */
                    sc = sc * w;                                           /* scoring_schemes.h:150-156 */
                    float v = M->value[po] + sc;
                    if (v < value) { value = v; value_midx = mi; value_sidx = s - 1; }
                }
            }
            M->value[o] = value; M->gapm_val[o] = gapm_val; M->gaps_val[o] = gaps_val;
            M->value_midx[o] = value_midx; M->value_sidx[o] = value_sidx;
            M->gapm_idx[o] = gapm_idx; M->gaps_idx[o] = gaps_idx;
        }
    }
    free(gaps_max); free(min_mpos);
    return M;
}

void so_mesh_free(so_mesh* M) {
    if (!M) return;
    free(M->value_midx); free(M->value_sidx); free(M->gapm_idx); free(M->gaps_idx);
    free(M->value); free(M->gapm_val); free(M->gaps_val); free(M);
}

/* ------------------------------------------------------------------ gap placement */
/* cseq_base::fix_duplicate_positions src/cseq.cpp:456-594 (iterators restated as indices; idx_type is
 * unsigned int, next_left_gap/next_right_gap are int exactly as in the reference). */
int so_fix_duplicate_positions(uint32_t* pos, uint8_t* masks, uint32_t n, uint32_t width, int lowercase) {
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
                    if (next_left_gap == -1) return 1; /* runtime_error "no space to left and right" :557-560 */
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
            if (lowercase) masks[last] |= SO_BASEM_LC;
        }
        last = curr;
    }
    return 0;
}

/* ------------------------------------------------------------------ backtrack */
typedef struct { uint32_t* pos; uint8_t* mask; uint32_t n; uint32_t width; } outseq;
/* cseq_base::append(const aligned_base&) src/cseq.cpp:79-95 */
static void out_append(outseq* o, uint32_t pos, uint8_t mask) {
    if (pos >= o->width) { o->pos[o->n] = pos; o->mask[o->n] = mask; o->n++; o->width = pos; }
    else { o->pos[o->n] = o->width; o->mask[o->n] = mask; o->n++; }
}

/* backtrack() src/mesh.h:534-739 */
int so_backtrack(const so_graph* g, const so_mesh* M, const uint8_t* q, uint32_t L, const so_align_params* p,
                 so_align_result* r, uint32_t* out_cols, uint8_t* out_masks) {
    const uint32_t W = g->W;
    const float ms = -p->match_score;
    const uint32_t send = L - 1;
    outseq o = { out_cols, out_masks, 0, 0 };
    uint8_t* is_first = (uint8_t*)calloc((size_t)g->V + 1, 1);
    for (uint32_t i = 0; i < g->n_first; i++) is_first[g->first[i]] = 1;
#define CELL(mm, ss) ((uint64_t)(mm) * L + (ss))
    /* starting point :567-592 */
    uint32_t m = g->last[0];
    for (uint32_t t = 0; t < g->V; t++) if (M->value[CELL(t, send)] < M->value[CELL(m, send)]) m = t;
    uint32_t s = send;
    for (uint32_t i = 0; i < g->n_last; i++) {
        uint32_t mt = g->last[i];
        for (uint32_t st = 0; st < L; st++)
            if (M->value[CELL(mt, st)] < M->value[CELL(m, s)]) { m = mt; s = st; }
    }
    r->end_m = m; r->end_s = s;
    /* right overhang :594-615 */
    int cutoff_tail = (int)(send - s);
    if (cutoff_tail && p->overhang != 1) {
        int pos = (p->overhang == 0) ? (int)W - 1 - (int)g->col[m] - cutoff_tail : 0;
        for (int i = 0; i < cutoff_tail; i++) {
            uint8_t b = q[L - 1 - i];
            if (p->lowercase == 2) b |= SO_BASEM_LC;
            int pp = pos++;
            out_append(&o, (uint32_t)(pp > 0 ? pp : 0), b);
        }
    }
    float rval = M->value[CELL(m, s)];                 /* :618 */
    uint32_t pos = W - 1 - g->col[m];                  /* :620 */
    float sum_weight = 0;
    out_append(&o, pos, q[s]);                         /* :626-628 */
    sum_weight = sum_weight + (g_colw ? ms * colw(g->col[m]) : ms) * g->weight[m];   /* :631-638: forced match */
    while (s != 0 && !is_first[m]) {                   /* :642-685 */
        uint32_t snew = M->value_sidx[CELL(m, s)];
        m = M->value_midx[CELL(m, s)];
        if (snew == M->value_sidx[CELL(m, snew)] && snew != 0) m = M->value_midx[CELL(m, snew)]; /* :653-655 */
        pos = W - 1 - g->col[m];
        while (s != snew) {
            --s;
            out_append(&o, pos, q[s]);
            sum_weight = sum_weight + (g_colw ? ms * colw(g->col[m]) : ms) * g->weight[m];
        }
    }
#undef CELL
    if (s != 0) { /* left overhang :690-721 */
        r->head = (int)s;
        if (p->overhang == 0) {
            while (s-- != 0) {
                uint8_t b = q[s];
                ++pos;
                if (p->lowercase == 2) b |= SO_BASEM_LC;
                out_append(&o, pos < W - 1 ? pos : W - 1, b);
            }
        } else if (p->overhang == 2) {
            int n = (int)s;
            while (n--) {
                uint8_t b = q[n];
                if (p->lowercase == 2) b |= SO_BASEM_LC;
                out_append(&o, W - (uint32_t)n - 1, b);
            }
        }
    } else {
        r->head = 0;
    }
    r->tail = cutoff_tail;
    free(is_first);
    /* out.setWidth(W); out.reverse() :723-724, src/cseq.cpp:98-132,283-289 */
    for (uint32_t i = 0; i < o.n / 2; i++) {
        uint32_t tp = o.pos[i]; o.pos[i] = o.pos[o.n - 1 - i]; o.pos[o.n - 1 - i] = tp;
        uint8_t tm = o.mask[i]; o.mask[i] = o.mask[o.n - 1 - i]; o.mask[o.n - 1 - i] = tm;
    }
    for (uint32_t i = 0; i < o.n; i++) o.pos[i] = W - 1 - o.pos[i];
    r->n_out = o.n;
    r->raw = rval; r->sum_weight = sum_weight;
    r->score = rval / sum_weight;                      /* :738 */
    float q100 = 100.f * r->score;                     /* src/align.cpp:509 */
    r->qual = (int)(q100 < 0.f ? 0.f : (q100 > 100.f ? 100.f : q100));
    if (so_fix_duplicate_positions(o.pos, o.mask, o.n, W, p->lowercase == 2)) return 3; /* :725 */
    return 0;
}

/* ------------------------------------------------------------------ aligner stage */
static int contains_query(const uint8_t* ref, uint32_t rlen, const uint8_t* q, uint32_t qlen, uint32_t* at) {
    if (qlen > rlen) return 0;
    for (uint32_t i = 0; i + qlen <= rlen; i++) { /* boost::icontains on the base strings, src/align.cpp:329-332 */
        uint32_t j = 0;
        while (j < qlen && ((ref[i + j] ^ q[j]) & 0xf) == 0) j++;
        if (j == qlen) { if (at) *at = i; return 1; }
    }
    return 0;
}

/* aligner::operator() src/align.cpp:307-460 (graph + simple scheme path) and do_align :475-521 */
int so_align(uint32_t* fam, uint32_t F, const uint8_t* masks, const uint32_t* cols, const uint64_t* off, uint32_t W,
             const uint8_t* q_in, uint32_t L, const so_align_params* p, so_align_result* r, uint32_t* out_cols,
             uint8_t* out_masks) {
    memset(r, 0, sizeof(*r));
    uint8_t* q = (uint8_t*)malloc((size_t)L + 1);
    for (uint32_t i = 0; i < L; i++) q[i] = (p->lowercase != 1) ? (q_in[i] & 0xf) : q_in[i]; /* :324-326 */
#define NOT_CONTAINS(j) (!contains_query(masks + off[fam[j]], (uint32_t)(off[fam[j] + 1] - off[fam[j]]), q_in, L, NULL))
    /* std::partition (libstdc++ bidirectional __partition, bits/stl_algo.h:1472-1495), src/align.cpp:333 */
    uint32_t first = 0, lastp = F;
    for (;;) {
        for (;;) { if (first == lastp) goto part_done; else if (NOT_CONTAINS(first)) ++first; else break; }
        --lastp;
        for (;;) { if (first == lastp) goto part_done; else if (!NOT_CONTAINS(lastp)) --lastp; else break; }
        { uint32_t t = fam[first]; fam[first] = fam[lastp]; fam[lastp] = t; }
        ++first;
    }
part_done:;
#undef NOT_CONTAINS
    uint32_t begin_containing = first;
    if (begin_containing != F) {
        if (p->realign) { /* :337-348 */
            F = begin_containing;
            if (F == 0) { r->status = 2; free(q); return 2; }
        } else { /* :349-388 steal the alignment */
            uint32_t src = begin_containing, at = 0;
            for (uint32_t j = begin_containing; j < F; j++)
                if (off[fam[j] + 1] - off[fam[j]] == L) { src = j; break; } /* iequals (containing + same length) */
            contains_query(masks + off[fam[src]], (uint32_t)(off[fam[src] + 1] - off[fam[src]]), q_in, L, &at);
            for (uint32_t i = 0; i < L; i++) { /* setAlignedBases: takes the relative's bases AND columns */
                out_cols[i] = cols[off[fam[src]] + at + i];
                out_masks[i] = masks[off[fam[src]] + at + i];
            }
            r->status = 1; r->score = 1.f; r->qual = 100; r->n_out = L; r->fam_used = F;
            free(q);
            return 1;
        }
    }
    r->fam_used = F;
    so_graph* g = so_graph_build(fam, F, masks, cols, off, W, p->fs_weight);
    r->n_nodes = g->V;
    so_mesh* M = so_mesh_compute(g, q, L, p);
    int st = so_backtrack(g, M, q, L, p, r, out_cols, out_masks);
    r->status = st;
    so_mesh_free(M);
    so_graph_free(g);
    free(q);
    return st;
}

/* ------------------------------------------------------------------ threaded whole path */
typedef struct {
    const so_index* ix; const uint8_t* masks; const uint32_t* cols; const uint64_t* off; uint32_t W, nq;
    const uint8_t* qmasks; const uint64_t* qoff; const int64_t* exclude_ids;
    const so_fam_params* fp; const so_align_params* ap; so_align_result* results;
    uint32_t* out_cols; uint8_t* out_masks;
    uint32_t next; uint64_t cells, posts;
    pthread_mutex_t mu;
} batch_ctx;

static void* batch_worker(void* arg) {
    batch_ctx* c = (batch_ctx*)arg;
    uint32_t cap = c->fp->fs_max * 4 + 64;
    uint32_t* ids = (uint32_t*)malloc((size_t)cap * 4);
    float* sc = (float*)malloc((size_t)cap * 4);
    for (;;) {
        pthread_mutex_lock(&c->mu);
        uint32_t i = c->next++;
        pthread_mutex_unlock(&c->mu);
        if (i >= c->nq) break;
        const uint8_t* q = c->qmasks + c->qoff[i];
        uint32_t L = (uint32_t)(c->qoff[i + 1] - c->qoff[i]);
        uint64_t P = 0;
        int n = so_family(c->ix, c->off, c->cols, q, L, c->exclude_ids ? c->exclude_ids[i] : -1, c->fp, ids, sc, cap, &P);
        uint64_t cells = 0;
        if (n < 0) { memset(&c->results[i], 0, sizeof(so_align_result)); c->results[i].status = 4; }
        else {
            so_align(ids, (uint32_t)n, c->masks, c->cols, c->off, c->W, q, L, c->ap, &c->results[i],
                     c->out_cols + c->qoff[i], c->out_masks + c->qoff[i]);
            if (c->results[i].status == 0 || c->results[i].status == 3) cells = (uint64_t)c->results[i].n_nodes * L;
        }
        pthread_mutex_lock(&c->mu);
        c->cells += cells; c->posts += P;
        pthread_mutex_unlock(&c->mu);
    }
    free(ids); free(sc);
    return NULL;
}

int so_run_batch(const so_index* ix, const uint8_t* masks, const uint32_t* cols, const uint64_t* off, uint32_t W,
                 uint32_t nq, const uint8_t* qmasks, const uint64_t* qoff, const int64_t* exclude_ids,
                 const so_fam_params* fp, const so_align_params* ap, int nthreads, so_align_result* results,
                 uint32_t* out_cols, uint8_t* out_masks, uint64_t* cells_total, uint64_t* postings_total) {
    batch_ctx c = { ix, masks, cols, off, W, nq, qmasks, qoff, exclude_ids, fp, ap, results, out_cols, out_masks, 0, 0, 0,
                    PTHREAD_MUTEX_INITIALIZER };
    if (nthreads <= 0) { long nc = sysconf(_SC_NPROCESSORS_ONLN); nthreads = nc > 0 ? (int)nc : 1; }
    pthread_t* th = (pthread_t*)malloc((size_t)nthreads * sizeof(pthread_t));
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, batch_worker, &c);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
    if (cells_total) *cells_total = c.cells;
    if (postings_total) *postings_total = c.posts;
    return nthreads;
}
