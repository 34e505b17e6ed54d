// Family finder stage: the reference's famfinder functor (src/famfinder.h:51-69) over the GPU k-mer search.
// operator()(tray) is the per-query drop-in; run(trays) handles a whole batch in one device call and is what
// the CLI uses (the reference's flow-graph node at unlimited concurrency, src/sina.cpp:511).
#ifndef SINA_B200_HOST_FAMFINDER_H
#define SINA_B200_HOST_FAMFINDER_H
#include <memory>
#include <vector>

#include "options.h"
#include "tray.h"

namespace sina {

enum TURN_TYPE { TURN_NONE = 0, TURN_REVCOMP = 1, TURN_ALL = 2 };

class famfinder {
    class impl;
    std::shared_ptr<impl> pimpl;
public:
    explicit famfinder(int device = 0);
    famfinder(const famfinder& o);
    famfinder& operator=(const famfinder& o);
    ~famfinder();
    tray operator()(const tray& t);
    void run(std::vector<tray>& trays);

    int turn_check(const cseq& query, bool all);  // src/famfinder.cpp:344-378: 0 none, 1 reversed, 2 complemented, 3 both

    static void get_options_description(po::options_description& main, po::options_description& adv);
    static void validate_vm(po::variables_map& vm, po::options_description& desc);
    static ENGINE_TYPE get_engine();

    // option values (reference defaults, src/famfinder.cpp:141-213); public so that embedders can set them
    struct options {
        TURN_TYPE turn_which = TURN_NONE;
        ENGINE_TYPE engine = ENGINE_SINA_KMER;
        unsigned int fs_min = 40, fs_max = 40;
        float fs_msc = 0.7f, fs_msc_max = 2.f;
        bool fs_leave_query_out = false;
        unsigned int fs_req = 1, fs_req_full = 1, fs_full_len = 1400, fs_req_gaps = 10, fs_min_len = 150;
        bool fs_no_fast = false;
        unsigned int fs_kmer_len = 10, fs_kmer_mm = 0;
        bool fs_kmer_norel = false;
        std::string database;
        std::string filter_weights;   // [sina_b200] file with one positional weight per alignment column (what --filter computes from an ARB SAI)
    };
    static options opts;
};

}  // namespace sina
#endif
