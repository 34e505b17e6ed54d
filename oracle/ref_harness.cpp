// TEST INFRASTRUCTURE ONLY -- never linked or loaded by the product path.
//
// C-API harness around the *unmodified* reference sources, compiled in place from
// /root/reference/src (mesh.h, graph.h, scoring_schemes.h, mseq.cpp, cseq.cpp,
// aligned_base.cpp, kmer.h, idset.h) with the boost->std shims in oracle/shim.
// Output: oracle/_ref/libsina_ref.so (git-ignored, travels to the GPU box prebuilt).
//
// What is the reference's own code here:      mseq ctor, dag::sort/reduce_edges, compute(),
//   backtrack(), cseq::append/reverse/setWidth/fix_duplicate_positions/getAligned,
//   kmer generators (all_/prefix_/unique_ kmers), vlimap (push_back/invert/increment).
//   cseq_comparator::operator() (cseq_comparator.cpp, compiled whole; its option parsing only has to compile).
// What is restated (needs ARB/boost/TBB in the reference, so cannot be compiled):
//   search_filter::operator(), k-mer branch (src/search_filter.cpp:244-330)
//   kmer_search::impl::build/find  (src/kmer_search.cpp:152-276, 365-420)
//   famfinder::impl::match + gap filter + fs_req (src/famfinder.cpp:497-612, 474-491)
//   aligner::operator() pre-steps and do_align glue (src/align.cpp:320-460, 475-521)
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <fstream>
#include <sstream>
#include <thread>
#include <atomic>
#include <unordered_set>
#include <memory>

#include "log.h"
#include "spdlog/sinks/null_sink.h"

namespace sina {
std::shared_ptr<spdlog::logger> Log::create_logger(std::string name) {
    return std::make_shared<spdlog::logger>(name, std::make_shared<spdlog::sinks::null_sink_mt>());
}
}  // namespace sina

#include "mesh.h"
#include "mseq.h"
#include "cseq.h"
#include "kmer.h"
#include "idset.h"
#include "cseq_comparator.h"

using namespace sina;

extern "C" {

struct ref_align_params {
    float match_score;     // --match-score      (align.cpp:250) default 2
    float mismatch_score;  // --mismatch-score   (align.cpp:253) default -1
    float gap_penalty;     // --pen-gap          (align.cpp:256) default 5
    float gap_ext_penalty; // --pen-gapext       (align.cpp:259) default 2
    float fs_weight;       // --fs-weight        (align.cpp:247) default 1
    int overhang;          // 0 attach, 1 remove, 2 edge  (align.h:42-46)
    int lowercase;         // 0 none, 1 original, 2 unaligned (align.h:51-55)
    int insertion;         // 0 shift, 1 forbid, 2 remove (align.h:60-64)
    int realign;           // --realign (align.cpp:232)
};

struct ref_fam_params {
    uint32_t fs_min, fs_max;      // famfinder.cpp:163,165
    float fs_msc, fs_msc_max;     // :167, :193
    uint32_t fs_min_len;          // :175
    uint32_t fs_req_full;         // :169
    uint32_t fs_full_len;         // :171
    uint32_t fs_req_gaps;         // :173
    uint32_t fs_req;              // :160
    int leave_query_out;          // :195
};

struct ref_db {
    uint32_t W;
    std::vector<cseq> seqs;
};

struct ref_kidx {
    ref_db* db;
    int k;
    bool nofast;
    uint32_t n_kmers;
    std::vector<vlimap*> lists;
    uint32_t n_seqs = 0;   // index built from bare lists (ref_kidx_from_lists): number of sequences, else db->seqs.size()
    uint32_t size() const { return db ? (uint32_t)db->seqs.size() : n_seqs; }
    ~ref_kidx() { for (auto* l : lists) delete l; }
};

// ---------------------------------------------------------------- database
ref_db* ref_db_create(uint32_t N, uint32_t W, const char* const* names, const char* const* rows) {
    auto* db = new ref_db;
    db->W = W;
    db->seqs.reserve(N);
    for (uint32_t i = 0; i < N; i++) {
        db->seqs.emplace_back(names ? names[i] : "", rows[i]);
        db->seqs.back().setWidth(W);
    }
    return db;
}

// rows as packed (char, column) pairs
ref_db* ref_db_create_packed(uint32_t N, uint32_t W, const char* const* names, const uint8_t* chars,
                             const uint32_t* cols, const uint64_t* off) {
    auto* db = new ref_db;
    db->W = W;
    db->seqs.reserve(N);
    for (uint32_t i = 0; i < N; i++) {
        db->seqs.emplace_back(names ? names[i] : "", nullptr);
        cseq& c = db->seqs.back();
        for (uint64_t j = off[i]; j < off[i + 1]; j++) c.append(aligned_base(cols[j], chars[j]));
        c.setWidth(W);
    }
    return db;
}

void ref_db_free(ref_db* db) { delete db; }
uint32_t ref_db_size(ref_db* db) { return db->seqs.size(); }

// ---------------------------------------------------------------- base / cseq KAT helpers
int ref_char_to_mask(unsigned char c) {
    try { aligned_base b(0, c); return base_iupac::iupac_char_to_bmask[c]; } catch (...) { return -1; }
}

// cseq::getAligned of a row given as aligned string (round trip, KAT helper)
int ref_cseq_roundtrip(const char* row, int nodots, int dna, char* out, int cap) {
    cseq c("x", row);
    std::string s = c.getAligned(nodots, dna);
    if ((int)s.size() + 1 > cap) return -1;
    memcpy(out, s.c_str(), s.size() + 1);
    return (int)s.size();
}

// fix_duplicate_positions on arbitrary (monotone) positions. returns 0 ok, 1 = runtime_error thrown
int ref_fix_duplicate_positions(uint32_t n, const uint32_t* pos_in, const uint8_t* chars, uint32_t width,
                                int lowercase, uint32_t* pos_out, uint8_t* chars_out) {
    cseq c("x", nullptr);
    for (uint32_t i = 0; i < n; i++) c.append(aligned_base(pos_in[i], chars[i]));
    c.setWidth(width);
    std::stringstream log;
    try {
        c.fix_duplicate_positions(log, lowercase != 0, false);
    } catch (std::runtime_error&) {
        return 1;
    }
    const auto& ab = c.getAlignedBases();
    for (uint32_t i = 0; i < n; i++) {
        pos_out[i] = ab[i].getPosition();
        chars_out[i] = ab[i].getBase();
    }
    return 0;
}

// ---------------------------------------------------------------- k-mer generators (KATs)
// mode: 0 all_kmers, 1 unique_kmers, 2 prefix_kmers(A), 3 unique_prefix_kmers(A)
int ref_kmers(const char* seq, int k, int mode, uint32_t* out, int cap) {
    cseq c("q", seq);
    const std::vector<aligned_base>& bases = c.getAlignedBases();
    std::unordered_set<unsigned int> seen;
    int n = 0;
    auto put = [&](unsigned int v) { if (n < cap) out[n] = v; n++; };
    switch (mode) {
        case 0: for (unsigned int v : all_kmers(bases, k, 1)) put(v); break;
        case 1: for (unsigned int v : unique_kmers(bases, seen, k)) put(v); break;
        case 2: for (unsigned int v : prefix_kmers(bases, k, 1, BASE_A)) put(v); break;
        case 3: for (unsigned int v : unique_prefix_kmers(bases, seen, k, 1, BASE_A)) put(v); break;
        default: return -1;
    }
    return n;
}

// vlimap KAT: build from ascending ids, optionally invert, increment a zeroed int16 vector.
int ref_vlimap_increment(uint32_t maxsize, const uint32_t* ids, uint32_t n, int invert, int16_t* scores) {
    vlimap v(maxsize);
    for (uint32_t i = 0; i < n; i++) v.push_back(ids[i]);
    if (invert) v.invert();
    idset::inc_t t(maxsize, 0);
    int r = v.increment(t);
    for (uint32_t i = 0; i < maxsize; i++) scores[i] = t[i];
    return r;
}

// ---------------------------------------------------------------- k-mer index (restated build/find)
ref_kidx* ref_kidx_build(ref_db* db, int k, int nofast) {
    auto* ix = new ref_kidx;
    ix->db = db; ix->k = k; ix->nofast = nofast != 0;
    ix->n_kmers = 1u << (2 * k);
    ix->lists.assign(ix->n_kmers, nullptr);
    uint32_t N = db->seqs.size();
    std::unordered_set<unsigned int> seen;
    for (uint32_t i = 0; i < N; i++) {  // kmer_search.cpp:158-181 (serial; join order == index order)
        const auto& bases = db->seqs[i].getAlignedBases();
        if (ix->nofast) {
            for (const auto& kmer : unique_kmers(bases, seen, k)) {
                if (!ix->lists[kmer]) ix->lists[kmer] = new vlimap(N);
                ix->lists[kmer]->push_back(i);
            }
        } else {
            for (unsigned int kmer : unique_prefix_kmers(bases, seen, (int)k, 1, BASE_A)) {
                if (!ix->lists[kmer]) ix->lists[kmer] = new vlimap(N);
                ix->lists[kmer]->push_back(i);
            }
        }
    }
    for (uint32_t i = 0; i < ix->n_kmers; i++)  // kmer_search.cpp:264-266
        if (ix->lists[i] && ix->lists[i]->size() > N / 2) ix->lists[i]->invert();
    return ix;
}
void ref_kidx_free(ref_kidx* ix) { delete ix; }

// kmer_search::impl::store (kmer_search.cpp:278-303) over the reference's own vlimap::write (idset.h:386-398)
int ref_kidx_store(ref_kidx* ix, const char* const* names, const char* path) {
    std::ofstream out(path, std::ofstream::binary);
    if (!out) return -1;
    struct idx_header { uint64_t magic{0x5844494b414e4953}; uint16_t vers{0}; uint32_t n_sequences{0}; uint16_t flags{0}; };  // kmer_search.cpp:66-88
    idx_header header;
    memset((void*)&header, 0, sizeof(header));
    header.magic = 0x5844494b414e4953; header.vers = 0;
    header.n_sequences = ix->size();
    header.flags = (uint16_t)((ix->k & 0xff) | (ix->nofast ? 0x100 : 0));
    out.write((char*)&header, sizeof(idx_header));
    for (uint32_t i = 0; i < header.n_sequences; i++) out << names[i] << std::endl;
    vlimap emptymap(header.n_sequences);
    for (unsigned int i = 0; i < ix->n_kmers; i++)
        if (ix->lists[i] != nullptr && ix->lists[i]->size() > 0) emptymap.push_back(i);
    emptymap.write(out);
    size_t idxno = 0;
    for (auto inc : emptymap) { idxno += inc; ix->lists[idxno]->write(out); }
    return out ? 0 : -1;
}

// index from bare posting lists (ascending ids), lists longer than N / 2 inverted as build() does (kmer_search.cpp:264-266)
ref_kidx* ref_kidx_from_lists(uint32_t N, int k, int nofast, uint32_t n_lists, const uint32_t* kmers, const uint64_t* list_off,
                              const uint32_t* ids) {
    auto* ix = new ref_kidx;
    ix->db = nullptr; ix->n_seqs = N; ix->k = k; ix->nofast = nofast != 0;
    ix->n_kmers = 1u << (2 * k);
    ix->lists.assign(ix->n_kmers, nullptr);
    for (uint32_t i = 0; i < n_lists; i++) {
        auto* l = new vlimap(N);
        for (uint64_t e = list_off[i]; e < list_off[i + 1]; e++) l->push_back(ids[e]);
        if (l->size() > N / 2) l->invert();
        ix->lists[kmers[i]] = l;
    }
    return ix;
}
// scores[N] after incrementing with list `kmer` (+ the offset an inverted list asks for): what find() adds for one k-mer
int ref_kidx_list_scores(ref_kidx* ix, uint32_t kmer, int16_t* scores) {
    idset::inc_t sc(ix->size(), 0);
    if (!ix->lists[kmer]) return -1;
    const int off = ix->lists[kmer]->increment(sc);
    for (uint32_t i = 0; i < ix->size(); i++) scores[i] = sc[i] + off;
    return off;
}

// kmer_search::impl::try_load (kmer_search.cpp:305-351) over the reference's own vlimap::read (idset.h:400-409);
// returns null on a header mismatch
ref_kidx* ref_kidx_load(ref_db* db, const char* path, int k, int nofast) {
    std::ifstream in(path, std::ifstream::binary);
    if (!in) return nullptr;
    struct idx_header { uint64_t magic; uint16_t vers; uint32_t n_sequences; uint16_t flags; };
    idx_header header;
    in.read((char*)&header, sizeof(idx_header));
    if (header.magic != 0x5844494b414e4953 || header.vers != 0) return nullptr;
    if ((header.flags & 0xff) != k || ((header.flags >> 8) & 1) != (nofast != 0)) return nullptr;
    if (db && header.n_sequences != db->seqs.size()) return nullptr;
    for (unsigned int i = 0; i < header.n_sequences; i++) { std::string name; getline(in, name); }
    auto* ix = new ref_kidx;
    ix->db = db; ix->n_seqs = header.n_sequences; ix->k = k; ix->nofast = nofast != 0;
    ix->n_kmers = 1u << (2 * k);
    ix->lists.assign(ix->n_kmers, nullptr);
    vlimap emptymap(header.n_sequences);
    emptymap.read(in);
    size_t idxno = 0;
    for (auto inc : emptymap) {
        idxno += inc;
        auto* idx = new vlimap(header.n_sequences);
        idx->read(in);
        ix->lists[idxno] = idx;
    }
    return ix;
}

// number of postings in list `kmer` (un-inverted size), and total
uint64_t ref_kidx_list_size(ref_kidx* ix, uint32_t kmer) { return ix->lists[kmer] ? ix->lists[kmer]->size() : 0; }

using rank_pair = std::pair<idset::inc_t::value_type, int>;

static void kidx_rank(const ref_kidx* ix, const cseq& query, std::vector<rank_pair>& ranks, uint64_t* postings) {
    uint32_t N = ix->db->seqs.size();
    idset::inc_t scores(N, 0);
    const std::vector<aligned_base>& bases = query.getAlignedBases();
    int offset = 0;
    uint64_t P = 0;
    if (ix->nofast) {  // kmer_search.cpp:389-395
        for (unsigned int kmer : all_kmers(bases, ix->k, 1))
            if (ix->lists[kmer]) { offset += ix->lists[kmer]->increment(scores); P += ix->lists[kmer]->size(); }
    } else {           // :396-402
        for (unsigned int kmer : prefix_kmers(bases, ix->k, 1, BASE_A))
            if (ix->lists[kmer]) { offset += ix->lists[kmer]->increment(scores); P += ix->lists[kmer]->size(); }
    }
    ranks.clear();
    ranks.reserve(N);
    int n = 0;
    for (auto score : scores) ranks.emplace_back(score + offset, n++);  // :405-409
    if (postings) *postings = P;
}

// kmer_search::impl::find (kmer_search.cpp:365-420). returns number of results (min(max,N))
uint32_t ref_kidx_find(ref_kidx* ix, const char* query, uint32_t max, int16_t* scores, uint32_t* ids,
                       uint64_t* postings) {
    uint32_t N = ix->db->seqs.size();
    if (max > N) max = N;
    if (postings) *postings = 0;
    if (max == 0) return 0;
    cseq q("q", query);
    std::vector<rank_pair> ranks;
    kidx_rank(ix, q, ranks, postings);
    std::partial_sort(ranks.begin(), ranks.begin() + max, ranks.end(), std::greater<rank_pair>());
    for (uint32_t i = 0; i < max; i++) { scores[i] = ranks[i].first; ids[i] = ranks[i].second; }
    return max;
}

// famfinder::impl::turn_check (famfinder.cpp:344-378) on the reference's own cseq::reverse / complement
// (cseq.cpp:284-296, aligned_base.h:117-124) and the restated find above. scores4 (optional) = the four top scores.
int ref_turn_check(ref_kidx* ix, const char* query, int all, int32_t* scores4) {
    const uint32_t N = ix->db->seqs.size();
    auto top = [&](const cseq& c) -> double {
        if (N == 0) return 0;
        std::vector<rank_pair> ranks;
        kidx_rank(ix, c, ranks, nullptr);
        std::partial_sort(ranks.begin(), ranks.begin() + 1, ranks.end(), std::greater<rank_pair>());
        return (float)ranks[0].first;
    };
    cseq q("q", query);
    double score[4];
    score[0] = top(q);
    cseq turn(q);
    turn.reverse();
    if (all) {
        score[1] = top(turn);
        cseq comp(q);
        comp.complement();
        score[2] = top(comp);
    } else {
        score[1] = score[2] = 0;
    }
    turn.complement();
    score[3] = top(turn);
    double max = 0;
    int best = 0;
    for (int i = 0; i < 4; i++) {
        if (max < score[i]) {
            max = score[i], best = i;
        }
    }
    if (scores4) for (int i = 0; i < 4; i++) scores4[i] = (int32_t)score[i];
    return best;
}

// ---------------------------------------------------------------- family selection (restated)
struct fam_item { float score; uint32_t id; };

// famfinder::impl::match (famfinder.cpp:497-612) + gap filter (:474-480); remove_similar with the reference's comparator
// (identity filter needs cseq_comparator; off at default). returns family size.
static uint32_t select_family(const ref_kidx* ix, const cseq& query, const ref_fam_params& p,
                              std::vector<fam_item>& fam, uint64_t* postings) {
    const auto& seqs = ix->db->seqs;
    uint32_t N = seqs.size();
    std::vector<rank_pair> ranks;
    kidx_rank(ix, query, ranks, postings);
    size_t have = 0, have_full = 0;
    auto is_full = [&](const fam_item& r) { return seqs[r.id].size() >= p.fs_full_len; };
    cseq_comparator similar(CMP_IUPAC_OPTIMISTIC, CMP_DIST_NONE, CMP_COVER_QUERY, false);   // famfinder.cpp:554
    auto remove = [&](const fam_item& r) {
        bool rm = seqs[r.id].size() < p.fs_min_len ||
                  (p.leave_query_out && query.getName() == seqs[r.id].getName()) ||
                  (p.fs_msc_max <= 2 && similar(query, seqs[r.id]) > p.fs_msc_max) ||   // remove_similar (:553-556)
                  (have >= p.fs_min && (have >= p.fs_max || !(r.score < p.fs_msc)) &&
                   !(p.fs_req_full && have_full < p.fs_req_full && is_full(r)));
        if (rm) return true;
        ++have;
        if (p.fs_req_full && is_full(r)) ++have_full;
        return false;
    };
    size_t max_results = p.fs_max + 1;
    std::vector<fam_item> results;
    std::vector<fam_item>::iterator from;
    fam.clear();
    bool entered = false;
    while (have < p.fs_max || have_full < p.fs_req_full) {
        entered = true;
        results.clear();
        uint32_t max = std::min<size_t>(max_results, N);
        if (max == 0) return 0;
        std::partial_sort(ranks.begin(), ranks.begin() + max, ranks.end(), std::greater<rank_pair>());
        for (uint32_t i = 0; i < max; i++) results.push_back({(float)ranks[i].first, (uint32_t)ranks[i].second});
        have = 0; have_full = 0;
        from = std::remove_if(results.begin(), results.end(), remove);
        if (max_results >= N) break;
        max_results *= 10;
    }
    if (!entered) return 0;
    results.erase(from, results.end());
    // stage body: remove sequences having too few gaps (famfinder.cpp:474-480)
    if (p.fs_req_gaps != 0) {
        auto too_few_gaps = [&](const fam_item& i) {
            const cseq& s = seqs[i.id];
            return 0 == s.size() || s.rbegin()->getPosition() - s.size() + 1 < p.fs_req_gaps;
        };
        results.erase(std::remove_if(results.begin(), results.end(), too_few_gaps), results.end());
    }
    fam = results;
    return fam.size();
}

// returns family size; -1 if < fs_req ("unable to align: too few relatives", famfinder.cpp:486-491)
int ref_family(ref_kidx* ix, const char* qname, const char* query, const ref_fam_params* p, uint32_t* ids,
               float* scores, uint32_t cap) {
    cseq q(qname ? qname : "", query);
    std::vector<fam_item> fam;
    uint32_t n = select_family(ix, q, *p, fam, nullptr);
    for (uint32_t i = 0; i < n && i < cap; i++) { ids[i] = fam[i].id; scores[i] = fam[i].score; }
    if (n < p->fs_req) return -1;
    return (int)n;
}

// ---------------------------------------------------------------- identity / --search
static cseq packed_cseq(const char* name, uint32_t n, const uint8_t* chars, const uint32_t* cols) {
    cseq c(name, nullptr);
    for (uint32_t j = 0; j < n; j++) c.append(aligned_base(cols[j], chars[j]));
    return c;
}

// cseq_comparator::operator() (src/cseq_comparator.cpp:209-293), the reference's own code.
// iupac: 0 optimistic 1 pessimistic 2 exact; dist: 0 none 1 jc; cover: CMP_COVER_TYPE order (abs, query, target, overlap,
// all, average, min, max, nogap)
float ref_compare(uint32_t na, const uint8_t* achars, const uint32_t* acols, uint32_t nb, const uint8_t* bchars,
                  const uint32_t* bcols, int iupac, int dist, int cover, int filter_lc) {
    cseq a = packed_cseq("a", na, achars, acols), b = packed_cseq("b", nb, bchars, bcols);
    cseq_comparator cmp((CMP_IUPAC_TYPE)iupac, (CMP_DIST_TYPE)dist, (CMP_COVER_TYPE)cover, filter_lc != 0);
    return cmp(a, b);
}

// search_filter::operator() (src/search_filter.cpp:244-330), the branch without --search-all: find(kmer_candidates),
// --search-ignore-super's partition + erase (which KEEPS the candidates containing the query, :313-316), comparator
// score of every candidate (:318-320), partial_sort by greater<result_item> = (score, name) descending (:322-330),
// cut at the first score <= min_sim. Returns the number of results, 0 for queries shorter than 20 bases (:253-256).
int ref_search(ref_kidx* ix, uint32_t n, const uint8_t* chars, const uint32_t* cols, uint32_t kmer_candidates,
               uint32_t max_result, float min_sim, int ignore_super, int iupac, int dist, int cover, int filter_lc,
               uint32_t* ids, float* scores) {
    cseq c = packed_cseq("query", n, chars, cols);
    if (c.size() < 20) return 0;
    struct item {
        float score;
        const cseq* sequence;
        bool operator<(const item& o) const {   // search::result_item (src/search.h:56-68)
            if (score < o.score) return true;
            if (score > o.score) return false;
            return *sequence < *o.sequence;
        }
        bool operator>(const item& o) const { return !operator<(o); }
    };
    std::vector<item> vc;
    {   // index->find(*c, vc, kmer_candidates) (src/kmer_search.cpp:365-420)
        uint32_t max = std::min<uint32_t>(kmer_candidates, ix->db->seqs.size());
        std::vector<rank_pair> ranks;
        kidx_rank(ix, c, ranks, nullptr);
        std::partial_sort(ranks.begin(), ranks.begin() + max, ranks.end(), std::greater<rank_pair>());
        for (uint32_t i = 0; i < max; i++) vc.push_back({(float)ranks[i].first, &ix->db->seqs[ranks[i].second]});
    }
    auto iupac_compare = [](const aligned_base& a, const aligned_base& b) { return a.comp(b); };
    auto contains_query = [&](item& it) {
        const auto& hay = it.sequence->getAlignedBases();
        const auto& needle = c.getAlignedBases();
        return std::search(hay.begin(), hay.end(), needle.begin(), needle.end(), iupac_compare) != hay.end();   // boost::algorithm::contains
    };
    if (ignore_super) {
        auto it = std::partition(vc.begin(), vc.end(), contains_query);
        vc.erase(it, vc.end());
    }
    cseq_comparator cmp((CMP_IUPAC_TYPE)iupac, (CMP_DIST_TYPE)dist, (CMP_COVER_TYPE)cover, filter_lc != 0);
    for (auto& r : vc) r.score = cmp(c, *r.sequence);
    auto it = vc.begin();
    auto middle = vc.begin() + max_result;
    auto end = vc.end();
    if (middle > end) middle = end;
    std::partial_sort(it, middle, end, std::greater<item>());
    while (it != middle && it->score > min_sim) ++it;
    vc.erase(it, vc.end());
    for (size_t i = 0; i < vc.size(); i++) { ids[i] = (uint32_t)(vc[i].sequence - ix->db->seqs.data()); scores[i] = vc[i].score; }
    return (int)vc.size();
}

// ---------------------------------------------------------------- graph dump
using mesh_tr = transition_simple<scoring_scheme_simple, mseq, cseq>;
using mesh_cell = mesh_tr::data_type;

struct graph_dump {
    std::vector<uint32_t> col, pred_off, preds, first, last;
    std::vector<uint8_t> ch;
    std::vector<float> weight;
};

static void dump_graph(mseq& m, graph_dump& g) {
    for (auto it = m.begin(); it != m.end(); ++it) {
        g.col.push_back(it->getPosition());
        g.ch.push_back((unsigned char)it->getBase());
        g.weight.push_back(it->getWeight());
        g.pred_off.push_back(g.preds.size());
        for (auto p = it.prev_begin(); p != it.prev_end(); ++p) g.preds.push_back(get_node_id(m, p));
    }
    g.pred_off.push_back(g.preds.size());
    for (auto p = m.pn_first_begin(); p != m.pn_first_end(); ++p) g.first.push_back(get_node_id(m, p));
    for (auto p = m.pn_last_begin(); p != m.pn_last_end(); ++p) g.last.push_back(get_node_id(m, p));
}

// Build mseq for family rows `fam` and dump it. Any out pointer may be null. Returns #nodes, or
// -(needed) if a capacity is too small.
int ref_graph(ref_db* db, const uint32_t* fam, uint32_t F, float fs_weight, uint32_t cap_nodes, uint32_t cap_edges,
              uint32_t* col, uint8_t* ch, float* weight, uint32_t* pred_off, uint32_t* preds, uint32_t* n_edges,
              uint32_t* first, uint32_t* n_first, uint32_t* last, uint32_t* n_last) {
    std::vector<const cseq*> vcp;
    for (uint32_t i = 0; i < F; i++) vcp.push_back(&db->seqs[fam[i]]);
    mseq m(vcp.begin(), vcp.end(), fs_weight);
    m.sort();
    m.reduce_edges();
    graph_dump g;
    dump_graph(m, g);
    uint32_t V = g.col.size();
    if (V > cap_nodes || g.preds.size() > cap_edges) return -(int)std::max<size_t>(V, g.preds.size());
    if (col) memcpy(col, g.col.data(), V * 4);
    if (ch) memcpy(ch, g.ch.data(), V);
    if (weight) memcpy(weight, g.weight.data(), V * 4);
    if (pred_off) memcpy(pred_off, g.pred_off.data(), (V + 1) * 4);
    if (preds) memcpy(preds, g.preds.data(), g.preds.size() * 4);
    if (n_edges) *n_edges = g.preds.size();
    if (first) memcpy(first, g.first.data(), g.first.size() * 4);
    if (n_first) *n_first = g.first.size();
    if (last) memcpy(last, g.last.data(), g.last.size() * 4);
    if (n_last) *n_last = g.last.size();
    return (int)V;
}

// ---------------------------------------------------------------- alignment
struct ref_align_result {
    int status;        // 0 aligned by DP, 1 copied from containing relative, 2 skipped (all relatives contained
                       // query and --realign), 3 runtime_error from fix_duplicate_positions, 4 no family
    float score;       // backtrack() return value = raw / sum_weight
    int head, tail;    // cutoff_head / cutoff_tail
    int qual;          // align_quality_slv
    uint32_t n_nodes;  // graph size
    uint32_t fam_used; // relatives left after the contains-query partition
};

static bool icontains(const std::string& hay, const std::string& needle) {
    auto it = std::search(hay.begin(), hay.end(), needle.begin(), needle.end(),
                          [](char a, char b) { return std::toupper((unsigned char)a) == std::toupper((unsigned char)b); });
    return it != hay.end();
}
static bool iequals(const std::string& a, const std::string& b) {
    return a.size() == b.size() && icontains(a, b);
}

struct align_item { float score; const cseq* sequence; };

// per-column weights of the alignment (alignment_stats::getWeights(), what --filter selects; empty = no filter).
// The reference indexes them up to position + 1 + insertion length without a bounds check (scoring_schemes.h:186-201):
// the test harness pads the vector so that those reads are defined (the padding repeats the last weight, the rule
// the oracle and the CUDA path use).
static std::vector<float> g_col_weights;

// do_align (align.cpp:475-521) for one transition type: transition_simple, or transition_aspace_aware for
// --insertion forbid (choose_transition, align.cpp:462-473)
extern "C++" {
template <typename TR, typename SCHEME>
static bool run_dp(mseq& m, cseq& c, const SCHEME& s, const ref_align_params& P, ref_align_result& R,
                   std::stringstream& log, std::string* logstr, uint32_t* cells, uint64_t cells_cap_words) {
    using cell_t = typename TR::data_type;
    TR tr(s);
    compute_node_simple<TR> cns(tr);
    mesh<mseq, cseq, cell_t> A(m, c);
    compute(A, cns);
    if (cells) {
        uint64_t n = (uint64_t)m.size() * c.size();
        if (n * 7 <= cells_cap_words) {
            for (uint64_t i = 0; i < n; i++) {
                cell_t& d = A(i);
                uint32_t* o = cells + i * 7;
                o[0] = d.value_midx; o[1] = d.value_sidx; o[2] = d.gapm_idx; o[3] = d.gaps_idx;
                memcpy(o + 4, &d.value, 4); memcpy(o + 5, &d.gapm_val, 4); memcpy(o + 6, &d.gaps_val, 4);
            }
        }
    }
    c.clearSequence();
    int oh_head = 0, oh_tail = 0;
    try {
        float score = backtrack(A, c, tr, (OVERHANG_TYPE)P.overhang, (LOWERCASE_TYPE)P.lowercase,
                                (INSERTION_TYPE)P.insertion, oh_head, oh_tail, log);
        R.score = score;
        R.head = oh_head; R.tail = oh_tail;
        R.qual = (int)std::min(100.f, std::max(0.f, 100.f * score));  // align.cpp:509
        R.status = 0;
    } catch (std::runtime_error& e) {
        R.status = 3;
        if (logstr) *logstr = log.str() + e.what();
        return false;
    }
    return true;
}
}  // extern "C++"

// aligner::operator() (align.cpp:307-460) + do_align (:475-521) for the graph/simple-scheme path.
// `cells` (optional) receives the full mesh, 7 x u32/f32 per cell in data_type field order
// {value_midx, value_sidx, gapm_idx, gaps_idx, value, gapm_val, gaps_val}; must hold n_nodes*qlen*7 words.
static void align_one(ref_db* db, std::vector<align_item>& vc, const cseq& input, const ref_align_params& P,
                      ref_align_result& R, std::string& aligned, std::vector<uint32_t>* out_cols,
                      std::string* logstr, uint32_t* cells, uint64_t cells_cap_words) {
    R = ref_align_result{};
    cseq c(input);
    const std::string bases = c.getBases();
    if (P.lowercase != LOWERCASE_ORIGINAL) c.upperCaseAll();
    std::stringstream log;

    auto not_contains_query = [&](align_item& item) { return !icontains(item.sequence->getBases(), bases); };
    auto begin_containing = std::partition(vc.begin(), vc.end(), not_contains_query);  // align.cpp:333
    bool copied = false;
    if (begin_containing != vc.end()) {
        if (P.realign) {
            vc.erase(begin_containing, vc.end());
            if (vc.empty()) { R.status = 2; return; }
        } else {
            auto exact = std::find_if(begin_containing, vc.end(),
                                      [&](align_item& it) { return iequals(bases, it.sequence->getBases()); });
            if (exact != vc.end()) {
                c.setAlignedBases(exact->sequence->getAlignedBases());
            } else {
                const std::vector<aligned_base>& refal = begin_containing->sequence->getAlignedBases();
                std::string refseq = begin_containing->sequence->getBases();
                auto it = std::search(refseq.begin(), refseq.end(), bases.begin(), bases.end(), [](char a, char b) {
                    return std::toupper((unsigned char)a) == std::toupper((unsigned char)b);
                });
                size_t off = it - refseq.begin();
                std::vector<aligned_base> sub(refal.begin() + off, refal.begin() + off + bases.size());
                c.setAlignedBases(sub);
            }
            c.setWidth(begin_containing->sequence->getWidth());
            R.status = 1; R.score = 1.f; R.qual = 100; R.head = 0; R.tail = 0;
            copied = true;
        }
    }
    R.fam_used = vc.size();
    if (!copied) {
        std::vector<const cseq*> vcp;
        for (auto& r : vc) vcp.push_back(r.sequence);
        mseq m(vcp.begin(), vcp.end(), P.fs_weight);
        m.sort();
        m.reduce_edges();
        R.n_nodes = m.size();
        bool ok;
        if (g_col_weights.empty()) {   // astats width 0: scoring_scheme_simple (align.cpp:405-408)
            scoring_scheme_simple s(-P.match_score, -P.mismatch_score, P.gap_penalty, P.gap_ext_penalty);
            ok = P.insertion == INSERTION_FORBID
                ? run_dp<transition_aspace_aware<scoring_scheme_simple, mseq, cseq>>(m, c, s, P, R, log, logstr, cells, cells_cap_words)
                : run_dp<mesh_tr>(m, c, s, P, R, log, logstr, cells, cells_cap_words);
        } else {                       // positional weights (--filter): scoring_scheme_weighted (align.cpp:409-415)
            std::vector<float> weights = g_col_weights;
            scoring_scheme_weighted s(-P.match_score, -P.mismatch_score, P.gap_penalty, P.gap_ext_penalty, weights);
            ok = P.insertion == INSERTION_FORBID
                ? run_dp<transition_aspace_aware<scoring_scheme_weighted, mseq, cseq>>(m, c, s, P, R, log, logstr, cells, cells_cap_words)
                : run_dp<transition_simple<scoring_scheme_weighted, mseq, cseq>>(m, c, s, P, R, log, logstr, cells, cells_cap_words);
        }
        if (!ok) return;
    }
    aligned = c.getAligned(true, false);  // rw_fasta.cpp:520
    if (out_cols) {
        out_cols->clear();
        for (const auto& ab : c.getAlignedBases()) out_cols->push_back(ab.getPosition());
    }
    if (logstr) *logstr = log.str();
}

void ref_set_column_weights(const float* w, uint32_t n, uint32_t pad) {
    g_col_weights.assign(w, w + n);
    if (n) g_col_weights.resize((size_t)n + pad, w[n - 1]);
}

// Align `query` against the given family (ids into db, family order as given).
// out_aligned: W+1 bytes. out_cols: qlen entries (may be null). log: optional.
int ref_align(ref_db* db, const uint32_t* fam, uint32_t F, const char* qname, const char* query,
              const ref_align_params* P, ref_align_result* R, char* out_aligned, uint32_t* out_cols,
              uint32_t* n_cols, char* logbuf, uint32_t logcap, uint32_t* cells, uint64_t cells_cap_words) {
    std::vector<align_item> vc;
    for (uint32_t i = 0; i < F; i++) vc.push_back({0.f, &db->seqs[fam[i]]});
    cseq q(qname ? qname : "q", query);
    std::string aligned, logstr;
    std::vector<uint32_t> cols;
    align_one(db, vc, q, *P, *R, aligned, &cols, &logstr, cells, cells_cap_words);
    if (out_aligned) { memcpy(out_aligned, aligned.c_str(), aligned.size() + 1); }
    if (out_cols) memcpy(out_cols, cols.data(), cols.size() * 4);
    if (n_cols) *n_cols = cols.size();
    if (logbuf && logcap) { strncpy(logbuf, logstr.c_str(), logcap - 1); logbuf[logcap - 1] = 0; }
    return R->status;
}

// ---------------------------------------------------------------- whole path, threaded (CPU baseline)
// For each query: family selection + alignment. out_cols: concatenated per-query columns (qoff gives
// offsets in bases); status per query; returns total DP cells via *cells_total. nthreads<=0 -> hw threads.
int ref_run_batch(ref_kidx* ix, uint32_t nq, const char* const* qnames, const char* const* queries,
                  const ref_fam_params* fp, const ref_align_params* ap, int nthreads, const uint64_t* qoff,
                  uint32_t* out_cols, ref_align_result* results, uint64_t* cells_total, uint64_t* postings_total) {
    if (nthreads <= 0) nthreads = std::max(1u, std::thread::hardware_concurrency());
    std::atomic<uint32_t> next(0);
    std::atomic<uint64_t> cells(0), posts(0);
    auto work = [&]() {
        std::vector<fam_item> fam;
        std::vector<align_item> vc;
        std::vector<uint32_t> cols;
        std::string aligned;
        for (;;) {
            uint32_t i = next.fetch_add(1);
            if (i >= nq) break;
            cseq q(qnames ? qnames[i] : "q", queries[i]);
            uint64_t P = 0;
            uint32_t n = select_family(ix, q, *fp, fam, &P);
            posts += P;
            if (n < fp->fs_req) { results[i] = ref_align_result{}; results[i].status = 4; continue; }
            vc.clear();
            for (auto& f : fam) vc.push_back({f.score, &ix->db->seqs[f.id]});
            align_one(ix->db, vc, q, *ap, results[i], aligned, &cols, nullptr, nullptr, 0);
            if (results[i].status == 0) cells += (uint64_t)results[i].n_nodes * q.size();
            if (out_cols && qoff && (results[i].status == 0 || results[i].status == 1))
                memcpy(out_cols + qoff[i], cols.data(), cols.size() * 4);
        }
    };
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(work);
    for (auto& t : th) t.join();
    if (cells_total) *cells_total = cells;
    if (postings_total) *postings_total = posts;
    return nthreads;
}

}  // extern "C"
