// The internal k-mer search engine (--fs-engine internal) backed by the GPU index.
// Mirrors kmer_search of the reference (src/kmer_search.h:38-76): a factory returning a handle on a
// process-wide index keyed by (database, k, nofast) -- here also by device -- and search::find.
#ifndef SINA_B200_HOST_KMER_SEARCH_H
#define SINA_B200_HOST_KMER_SEARCH_H
#include <memory>
#include <string>

#include "reference_db.h"
#include "search.h"

struct sg_index;

namespace sina {

class kmer_search : public search {
public:
    // src/kmer_search.cpp:118-134. `nofast` is what famfinder passes as opts.fs_no_fast (src/famfinder.cpp:289-292).
    static kmer_search* get_kmer_search(const std::string& database, int k = 10, bool nofast = false, int device = 0);
    static void release_kmer_search(const std::string& database, int k = 10, bool nofast = false, int device = 0);
    ~kmer_search() override;

    // src/kmer_search.cpp:365-420: top-`max` references by shared k-mer count, (score desc, index desc)
    void find(const cseq& query, result_vector& results, unsigned int max) override;
    // the same for many queries in one device call
    void find(const std::vector<const cseq*>& queries, std::vector<result_vector>& results, unsigned int max);
    unsigned int size() const override;

    sg_index* handle() const;
    /// positional weights of the alignment's columns (alignment_stats::getWeights()): the aligner then scores with
    /// scoring_scheme_weighted (src/align.cpp:409-415); an empty vector switches back to scoring_scheme_simple
    void set_column_weights(const std::vector<float>& w);
    const reference_db& db() const;
    int device() const;

    class impl;
private:
    explicit kmer_search(std::shared_ptr<impl> p);
    std::shared_ptr<impl> pimpl;
};

// query bases as the C-ABI wants them: masks of all queries back to back + offsets
void pack_queries(const std::vector<const cseq*>& queries, std::vector<uint8_t>& masks, std::vector<uint64_t>& off);
// throws std::runtime_error carrying sg_last_error() when rc != 0
void check_sg(int rc, const char* what);

}  // namespace sina
#endif
