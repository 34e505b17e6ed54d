#include "kmer_search.h"

#include <algorithm>

#include <cstdlib>
#include <fstream>
#include <iostream>

#include "sidx.h"

#include <map>
#include <mutex>
#include <tuple>

#include "../../include/sina_b200.h"

namespace sina {

void check_sg(int rc, const char* what) {
    if (rc != SG_OK) throw std::runtime_error(std::string(what) + ": " + sg_last_error());
}

void pack_queries(const std::vector<const cseq*>& queries, std::vector<uint8_t>& masks, std::vector<uint64_t>& off) {
    off.assign(1, 0);
    masks.clear();
    for (const cseq* q : queries) {
        for (const auto& b : q->getAlignedBases()) masks.push_back(b.getBase());
        off.push_back(masks.size());
    }
    if (masks.empty()) masks.push_back(0);
}

class kmer_search::impl {
public:
    reference_db* rdb = nullptr;
    sg_index* ix = nullptr;
    int dev = 0;
    ~impl() { if (ix) sg_index_destroy(ix); }
};

namespace {
using key_t = std::tuple<std::string, int, bool, int>;
std::mutex g_mu;
std::map<key_t, std::shared_ptr<kmer_search::impl>> g_indices;
}  // namespace

kmer_search::kmer_search(std::shared_ptr<impl> p) : pimpl(std::move(p)) {}
kmer_search::~kmer_search() = default;

kmer_search* kmer_search::get_kmer_search(const std::string& database, int k, bool nofast, int device) {
    std::lock_guard<std::mutex> lock(g_mu);
    const key_t key(database, k, nofast, device);
    auto it = g_indices.find(key);
    if (it == g_indices.end()) {
        auto p = std::make_shared<impl>();
        p->rdb = reference_db::getDB(database);
        p->dev = device;
        check_sg(sg_index_create(p->rdb->masks().data(), p->rdb->cols().data(), p->rdb->offsets().data(),
                                 p->rdb->getSeqCount(), p->rdb->getAlignmentWidth(), k, nofast ? 1 : 0, device, &p->ix),
                 "building the k-mer index");
        {   // order of the names: search::result_item breaks score ties by name (src/search.h:56-68)
            const std::vector<std::string> names = p->rdb->getSequenceNames();
            std::vector<uint32_t> order(names.size()), rank(names.size());
            for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
            std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return names[a] != names[b] ? names[a] < names[b] : a < b; });
            for (uint32_t i = 0; i < order.size(); i++) rank[order[i]] = i;
            check_sg(sg_index_set_name_ranks(p->ix, rank.data(), (uint32_t)rank.size()), "setting the name order");
        }
        // index cache next to the database, as kmer_search::impl::impl keeps one (src/kmer_search.cpp:213-242): written when
        // there is none for this (k, fast) yet. The GPU rebuilds the index faster than the file is read, so the cache is never
        // loaded for its lists: it records the index order (reference_db::getDB) and serves SINA itself.
        if (!getenv("SINA_B200_NO_SIDX") && std::ifstream(database).good()) {   // only next to a database that is a file
            try {
                sidx::info have;
                bool ok = false;
                try { ok = sidx::read(database + ".sidx", have) && have.k == (unsigned)k && have.nofast == nofast &&
                           have.n_sequences == p->rdb->getSeqCount(); } catch (std::exception&) { ok = false; }
                if (!ok) {
                    uint64_t n_post = 0;
                    check_sg(sg_index_info(p->ix, nullptr, nullptr, nullptr, nullptr, &n_post, nullptr, nullptr), "index info");
                    const uint64_t n_slots = 1ull << (2 * (nofast ? k : k - 1));
                    std::vector<uint64_t> off(n_slots + 1);
                    std::vector<uint32_t> ids(n_post ? n_post : 1);
                    check_sg(sg_index_export_lists(p->ix, off.data(), ids.data()), "exporting the posting lists");
                    sidx::write(database + ".sidx", (unsigned)k, nofast, p->rdb->getSequenceNames(), off.data(), ids.data(), n_slots);
                }
            } catch (std::exception& e) {
                std::cerr << "warning: index cache not written: " << e.what() << std::endl;
            }
        }
        it = g_indices.emplace(key, std::move(p)).first;
    }
    return new kmer_search(it->second);
}

void kmer_search::set_column_weights(const std::vector<float>& w) {
    check_sg(sg_index_set_column_weights(pimpl->ix, w.empty() ? nullptr : w.data(), (uint32_t)w.size()), "setting the column weights");
}

void kmer_search::release_kmer_search(const std::string& database, int k, bool nofast, int device) {
    std::lock_guard<std::mutex> lock(g_mu);
    g_indices.erase(key_t(database, k, nofast, device));
}

unsigned int kmer_search::size() const { return pimpl->rdb->getSeqCount(); }
sg_index* kmer_search::handle() const { return pimpl->ix; }
const reference_db& kmer_search::db() const { return *pimpl->rdb; }
int kmer_search::device() const { return pimpl->dev; }

void kmer_search::find(const std::vector<const cseq*>& queries, std::vector<result_vector>& results, unsigned int max) {
    results.assign(queries.size(), result_vector());
    const unsigned int m = std::min<unsigned int>(max, size());
    if (m == 0 || queries.empty()) return;   // src/kmer_search.cpp:367-370
    // queries shorter than 2 bases have no k-mers: they score 0 everywhere; the device call wants >= 2 bases,
    // so they get a two-base stand-in made of ambiguity codes (no valid k-mer either)
    std::vector<cseq> standins;
    std::vector<const cseq*> qs(queries);
    standins.reserve(queries.size());
    for (auto& q : qs)
        if (q->size() < 2) { standins.emplace_back("", "NN"); q = &standins.back(); }
    std::vector<uint8_t> masks;
    std::vector<uint64_t> off;
    pack_queries(qs, masks, off);
    std::vector<int16_t> scores((size_t)queries.size() * m);
    std::vector<uint32_t> ids((size_t)queries.size() * m), nres(queries.size());
    check_sg(sg_find_batch(pimpl->ix, masks.data(), off.data(), (uint32_t)queries.size(), max, scores.data(), ids.data(),
                           nres.data()),
             "k-mer search");
    for (size_t q = 0; q < queries.size(); q++) {
        results[q].reserve(nres[q]);
        for (uint32_t i = 0; i < nres[q]; i++)
            results[q].emplace_back((float)scores[q * m + i], &pimpl->rdb->getCseq(ids[q * m + i]));
    }
}

void kmer_search::find(const cseq& query, result_vector& results, unsigned int max) {
    std::vector<result_vector> r;
    find(std::vector<const cseq*>{&query}, r, max);
    results = r.empty() ? result_vector() : std::move(r[0]);
}

}  // namespace sina
