#include "search_filter.h"

#include <cstdio>

#include "../../include/sina_b200.h"
#include "famfinder.h"
#include "kmer_search.h"

namespace sina {

search_filter::options* search_filter::opts = nullptr;
const char* const fn_nearest = "nearest_slv";

static std::function<void(const std::string&)> name_parser(int* target, std::vector<std::pair<std::string, int>> names, const std::string& what, bool prefix) {
    return [target, names, what, prefix](const std::string& v) {
        std::string s;
        for (char c : v) s.push_back((char)tolower((unsigned char)c));
        for (const auto& n : names)
            if (prefix ? (!s.empty() && n.first.compare(0, s.size(), s) == 0) : n.first == s) { *target = n.second; return; }
        throw std::logic_error(what);
    };
}

void search_filter::get_options_description(po::options_description& main, po::options_description& adv) {
    if (!opts) opts = new options();
    po::options_description mid("Search & Classify");
    mid.value<std::string>("search-db", &opts->search_db, "", "reference db if different from -r/--db");
    mid.unsupported("search-engine", true, "only the internal k-mer engine");
    mid.value<float>("search-min-sim", &opts->min_sim, 0.7f, "required sequence similarity (0.7)");
    mid.value<unsigned int>("search-max-result", &opts->max_result, 10u, "desired number of search results (10)");
    mid.unsupported("lca-fields", true, "taxonomy fields live in an ARB database");
    mid.unsupported("lca-quorum", true, "taxonomy fields live in an ARB database");
    main.add(mid);
    po::options_description od("Search & Classify");
    od.unsupported("search-port", true, "PT server");
    od.unsupported("search-all", false, "comparing against every reference sequence");
    od.flag("search-no-fast", &opts->search_no_fast, "don't use fast family search");
    od.value<unsigned int>("search-kmer-candidates", &opts->kmer_candidates, 1000u, "number of most similar sequences to acquire via kmer-step (1000)");
    od.value<unsigned int>("search-kmer-len", &opts->kmer_len, 10u, "length of k-mers (10)");
    od.unsupported("search-kmer-mm", true, "the internal engine ignores it");
    od.unsupported("search-kmer-norel", false, "the internal engine ignores it");
    od.flag("search-ignore-super", &opts->ignore_super, "ignore sequences containing query");
    od.unsupported("search-copy-fields", true, "fields live in an ARB database");
    // cseq_comparator::get_options_description("search-") (src/cseq_comparator.cpp:432-462)
    od.custom("search-iupac", "optimistic", "strategy for comparing ambiguous bases [pessimistic|*optimistic*|exact]",
              name_parser(&opts->iupac, {{"optimistic", 0}, {"pessimistic", 1}, {"exact", 2}},
                          "iupac matching must be either optimistic or pessimistic", true));
    od.custom("search-correction", "none", "apply distance correction. [*none*|jc]",
              name_parser(&opts->correction, {{"none", 0}, {"jc", 1}}, "distance correction must be either none or jc", false));
    od.custom("search-cover", "query", "compute comparative measure relative to [abs|*query*|target|min|max|average|overlap|all|nogap]",
              name_parser(&opts->cover, {{"abs", 0}, {"query", 1}, {"target", 2}, {"overlap", 3}, {"all", 4}, {"average", 5}, {"min", 6}, {"max", 7}, {"nogap", 8}},
                          "coverage type must be one of abs, query, target, overlap,average, nogap, min or max", false));
    od.flag("search-filter-lowercase", &opts->filter_lowercase, "ignore bases in lowercase when comparing sequences");
    adv.add(od);
}

void search_filter::validate_vm(po::variables_map& /*vm*/, po::options_description& /*desc*/) {
    // src/search_filter.cpp:140-176: the search database defaults to --db, its engine to --fs-engine
    if (opts->search_db.empty()) opts->search_db = famfinder::opts.database;
    if (opts->search_db.empty()) throw std::logic_error("Search module requires reference database (--db or --search-db)");
    if (opts->cover == 0 && opts->correction != 0) throw std::logic_error("only fractional identity can be distance corrected");   // src/cseq_comparator.cpp:478-481
}

search_filter::search_filter(int device)
    : index(kmer_search::get_kmer_search(opts->search_db, (int)opts->kmer_len, opts->search_no_fast, device)) {}
search_filter::search_filter(const search_filter& rhs)
    : index(kmer_search::get_kmer_search(opts->search_db, (int)opts->kmer_len, opts->search_no_fast, rhs.index->device())) {}
search_filter& search_filter::operator=(const search_filter& /*rhs*/) { return *this; }
search_filter::~search_filter() { delete index; }

void search_filter::run(std::vector<tray*>& trays) {
    const reference_db& db = index->db();
    std::vector<tray*> live;
    std::vector<uint8_t> masks;
    std::vector<uint32_t> cols;
    std::vector<uint64_t> off(1, 0);
    for (tray* t : trays) {
        cseq* c = t->aligned_sequence;
        if (c == nullptr) { t->log << "search: no sequence?!;"; continue; }            // src/search_filter.cpp:246-251
        if (c->size() < 20) { t->log << "search:sequence too short (<20 bases);"; continue; }   // :253-256
        live.push_back(t);
        for (const aligned_base& b : c->getAlignedBases()) { masks.push_back(b.getBase()); cols.push_back(b.getPosition()); }
        off.push_back(masks.size());
    }
    if (live.empty()) return;
    sg_search_params sp;
    sg_default_search_params(&sp);
    sp.kmer_candidates = opts->kmer_candidates; sp.max_result = opts->max_result; sp.min_sim = opts->min_sim;
    sp.ignore_super = opts->ignore_super; sp.iupac = opts->iupac; sp.correction = opts->correction; sp.cover = opts->cover;
    sp.filter_lowercase = opts->filter_lowercase;
    std::vector<uint32_t> ids((size_t)live.size() * sp.max_result), n(live.size());
    std::vector<float> scores((size_t)live.size() * sp.max_result);
    check_sg(sg_search_batch(index->handle(), masks.data(), cols.data(), off.data(), (uint32_t)live.size(), &sp, ids.data(),
                             scores.data(), n.data()),
             "search");
    for (size_t q = 0; q < live.size(); q++) {
        tray& t = *live[q];
        t.search_result = new search::result_vector();
        std::string nearest;
        for (uint32_t i = 0; i < n[q]; i++) {
            const cseq& r = db.getCseq(ids[q * sp.max_result + i]);
            t.search_result->emplace_back(scores[q * sp.max_result + i], &r);
            char buf[64];
            snprintf(buf, sizeof(buf), "~%.3f ", scores[q * sp.max_result + i]);   // "{acc}.{version}.{start}.{stop}~{:.3f} " (:363-369)
            nearest += r.getName() + buf;
        }
        t.aligned_sequence->set_attr<std::string>(fn_nearest, nearest);   // :377
    }
}

void search_filter::run(std::vector<tray>& trays) {
    std::vector<tray*> p;
    for (tray& t : trays) p.push_back(&t);
    run(p);
}

tray search_filter::operator()(tray t) {
    std::vector<tray*> p{&t};
    run(p);
    return t;
}

}  // namespace sina
