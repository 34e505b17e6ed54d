// Search stage: the reference's search_filter functor (src/search_filter.h, src/search_filter.cpp:244-330) over the
// GPU k-mer search + sequence comparator (sg_search_batch). What needs ARB fields of the reference database
// (acc/version/start/stop in `nearest_slv`, --lca-fields classification, --search-copy-fields) has no source in a
// FASTA database: `nearest_slv` carries the reference's name instead, the other two are rejected.
#ifndef SINA_B200_HOST_SEARCH_FILTER_H
#define SINA_B200_HOST_SEARCH_FILTER_H
#include <string>
#include <vector>

#include "options.h"
#include "tray.h"

namespace sina {

class kmer_search;
extern const char* const fn_nearest;   // "nearest_slv" (query_arb::fn_nearest)

class search_filter {
public:
    struct options {   // src/search_filter.cpp:66-133, src/cseq_comparator.cpp:432-462
        std::string search_db;
        unsigned int kmer_candidates = 1000, max_result = 10;
        float min_sim = 0.7f;
        bool ignore_super = false, search_no_fast = false, filter_lowercase = false;
        unsigned int kmer_len = 10;
        int iupac = 0, correction = 0, cover = 1;
    };
    static options* opts;

    explicit search_filter(int device = 0);
    search_filter(const search_filter& rhs);
    ~search_filter();
    search_filter& operator=(const search_filter& rhs);
    tray operator()(tray t);
    void run(std::vector<tray>& trays);

    static void get_options_description(po::options_description& all, po::options_description& adv);
    static void validate_vm(po::variables_map& vm, po::options_description& desc);

private:
    void run(std::vector<tray*>& trays);
    kmer_search* index;
};

}  // namespace sina
#endif
