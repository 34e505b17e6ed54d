// Aligner stage: the reference's aligner functor (src/align.h:70-84) over the GPU graph/DP/backtrack path.
#ifndef SINA_B200_HOST_ALIGN_H
#define SINA_B200_HOST_ALIGN_H
#include <vector>

#include "options.h"
#include "tray.h"

namespace sina {

enum OVERHANG_TYPE { OVERHANG_ATTACH, OVERHANG_REMOVE, OVERHANG_EDGE };
enum LOWERCASE_TYPE { LOWERCASE_NONE, LOWERCASE_ORIGINAL, LOWERCASE_UNALIGNED };
enum INSERTION_TYPE { INSERTION_SHIFT, INSERTION_FORBID, INSERTION_REMOVE };

class kmer_search;

class aligner {
public:
    struct options {  // src/align.cpp:225-277
        bool realign = false;
        OVERHANG_TYPE overhang = OVERHANG_ATTACH;
        LOWERCASE_TYPE lowercase = LOWERCASE_NONE;
        INSERTION_TYPE insertion = INSERTION_SHIFT;
        float fs_weight = 1.f, match_score = 2.f, mismatch_score = -1.f, gap_penalty = 5.f, gap_ext_penalty = 2.f;
        bool write_used_rels = false, calc_idty = false;
    };
    static options* opts;

    explicit aligner(int device = 0);
    aligner(const aligner& rhs);
    ~aligner();
    aligner& operator=(const aligner& rhs);
    tray operator()(tray t);
    void run(std::vector<tray>& trays);

    static void get_options_description(po::options_description& all, po::options_description& adv);
    static void validate_vm(po::variables_map& vm, po::options_description& desc);

private:
    void run(std::vector<tray*>& trays, bool rethrow);
    kmer_search* index;  // gives the device handle holding the reference rows
};

}  // namespace sina
#endif
