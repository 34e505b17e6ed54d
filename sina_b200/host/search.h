// search engine interface of the reference, kept verbatim in shape (src/search.h:51-106).
#ifndef SINA_B200_HOST_SEARCH_H
#define SINA_B200_HOST_SEARCH_H
#include <vector>

#include "cseq.h"

namespace sina {

enum ENGINE_TYPE { ENGINE_ARB_PT = 0, ENGINE_SINA_KMER = 1 };

class search {
protected:
    search() = default;
public:
    search(const search&) = delete;
    search& operator=(const search&) = delete;
    virtual ~search() = default;
    struct result_item {
        result_item(float sc, const cseq* seq) : score(sc), sequence(seq) {}
        float score;
        const cseq* sequence;
        bool operator<(const result_item& o) const {
            if (score < o.score) return true;
            if (score > o.score) return false;
            return *sequence < *o.sequence;
        }
        bool operator>(const result_item& o) const { return !operator<(o); }
    };
    using result_vector = std::vector<result_item>;

    virtual void find(const cseq& query, result_vector& results, unsigned int max) = 0;
    virtual unsigned int size() const = 0;
};

}  // namespace sina
#endif
