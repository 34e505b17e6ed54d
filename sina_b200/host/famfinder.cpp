#include "famfinder.h"

#include <algorithm>

#include <cstdio>

#include "../../include/sina_b200.h"
#include "kmer_search.h"

namespace sina {

famfinder::options famfinder::opts;

void famfinder::get_options_description(po::options_description& main, po::options_description& adv) {
    // names, defaults and help texts: src/famfinder.cpp:141-213
    main.value<std::string>("db,r", &opts.database, "", "reference database (aligned FASTA)");
    main.custom("turn,t", "none", "check other strand as well ('all' checks all four frames)", [](const std::string& v) {
        if (v == "none") opts.turn_which = TURN_NONE;
        else if (v == "revcomp") opts.turn_which = TURN_REVCOMP;
        else if (v == "all") opts.turn_which = TURN_ALL;
        else throw std::logic_error("Turn type must be one of 'none', 'revcomp' or 'all'");
    });
    po::options_description mid("Reference Selection");
    mid.custom("fs-engine", "internal", "search engine to use for reference selection [pt-server|*internal*]",
               [](const std::string& v) {
                   if (v == "internal") opts.engine = ENGINE_SINA_KMER;
                   else if (v == "pt-server")
                       throw std::logic_error("--fs-engine pt-server is not supported by sina_b200 (no ARB PT server); use 'internal'");
                   else throw std::logic_error("engine must be one of 'internal' or 'pt-server'");
               });
    mid.value<unsigned int>("fs-kmer-len", &opts.fs_kmer_len, 10u, "length of k-mers (10)");
    mid.value<unsigned int>("fs-req", &opts.fs_req, 1u, "required number of reference sequences (1)");
    mid.value<unsigned int>("fs-min", &opts.fs_min, 40u, "number of references used regardless of shared fraction (40)");
    mid.value<unsigned int>("fs-max", &opts.fs_max, 40u, "number of references used at most (40)");
    mid.value<float>("fs-msc", &opts.fs_msc, 0.7f, "required fractional identity of references (0.7)");
    mid.value<unsigned int>("fs-req-full", &opts.fs_req_full, 1u, "required number of full length references (1)");
    mid.value<unsigned int>("fs-full-len", &opts.fs_full_len, 1400u, "minimum length of full length reference (1400)");
    mid.value<unsigned int>("fs-req-gaps", &opts.fs_req_gaps, 10u, "ignore references with less internal gaps (10)");
    mid.value<unsigned int>("fs-min-len", &opts.fs_min_len, 150u, "minimal reference length (150)");
    main.add(mid);

    po::options_description od("Advanced Reference Selection");
    od.unsupported("ptdb", true, "PT server database");
    od.unsupported("ptport", true, "PT server port");
    od.flag("fs-kmer-no-fast", &opts.fs_no_fast, "don't use fast family search");
    // accepted and ignored, exactly as the reference's internal engine does (src/famfinder.cpp:279-292)
    od.value<unsigned int>("fs-kmer-mm", &opts.fs_kmer_mm, 0u, "allowed mismatches per k-mer (0) [ignored by the internal engine]");
    od.flag("fs-kmer-norel", &opts.fs_kmer_norel, "don't score k-mer distance relative to target length [ignored by the internal engine]");
    od.value<float>("fs-msc-max", &opts.fs_msc_max, 2.f, "max identity of used references (for evaluation)");
    od.flag("fs-leave-query-out", &opts.fs_leave_query_out, "ignore candidate if found in reference (for evaluation)");
    od.unsupported("gene-start", true, "gene range quotas need ARB field data");
    od.unsupported("gene-end", true, "gene range quotas need ARB field data");
    od.unsupported("fs-cover-gene", true, "gene range quotas need ARB field data");
    od.unsupported("filter", true, "positional variability filters are ARB SAI data; pass the column weights themselves with --filter-weights FILE");
    od.value<std::string>("filter-weights", &opts.filter_weights, "", "[sina_b200] file with one positional weight per alignment column "
                          "(whitespace separated; alignment_stats::getWeights(), src/alignment_stats.cpp:54-112): selects the weighted scoring scheme");
    od.unsupported("auto-filter-field", true, "positional variability filters are ARB SAI data");
    od.unsupported("auto-filter-threshold", true, "positional variability filters are ARB SAI data");
    od.unsupported("fs-oldmatch", false, "legacy PT-server family composition");
    adv.add(od);
}

void famfinder::validate_vm(po::variables_map& vm, po::options_description& /*desc*/) {
    if (vm.count("db") == 0) throw std::logic_error("Family Finder: Must have reference database (--db/-r)");  // famfinder.cpp:218-220
    if (opts.fs_kmer_len < 1 || opts.fs_kmer_len > 16) throw std::logic_error("Family Finder: K must be in 1..16");
    if (opts.fs_max == 0) throw std::logic_error("Family Finder: --fs-max must be > 0");
}

ENGINE_TYPE famfinder::get_engine() { return opts.engine; }

class famfinder::impl {
public:
    explicit impl(int device) : index(kmer_search::get_kmer_search(opts.database, (int)opts.fs_kmer_len, opts.fs_no_fast, device)) {}
    ~impl() { delete index; }   // src/famfinder.cpp:306-308
    kmer_search* index;
    void run(std::vector<tray*>& trays);
};

famfinder::famfinder(int device) : pimpl(new impl(device)) {}
famfinder::famfinder(const famfinder& o) = default;
famfinder& famfinder::operator=(const famfinder& o) = default;
famfinder::~famfinder() = default;

int famfinder::turn_check(const cseq& query, bool all) {
    if (query.size() < 2) return 0;
    std::vector<const cseq*> qs{&query};
    std::vector<uint8_t> masks;
    std::vector<uint64_t> off;
    pack_queries(qs, masks, off);
    int32_t turn = 0;
    check_sg(sg_turn_batch(pimpl->index->handle(), masks.data(), off.data(), 1, all ? 2 : 1, &turn), "orientation check");
    return turn;
}

// famfinder::impl::do_turn_check (src/famfinder.cpp:311-341) for a batch of trays
static void do_turn_check(sg_index* handle, std::vector<tray*>& live, const std::vector<uint8_t>& masks,
                          const std::vector<uint64_t>& off, TURN_TYPE which) {
    if (which == TURN_NONE) {
        for (tray* t : live) t->input_sequence->set_attr<std::string>(fn_turn, "turn-check disabled");
        return;
    }
    std::vector<int32_t> turn(live.size());
    check_sg(sg_turn_batch(handle, masks.data(), off.data(), (uint32_t)live.size(), which == TURN_ALL ? 2 : 1, turn.data()),
             "orientation check");
    static const char* const what[4] = {"none", "reversed", "complemented", "reversed and complemented"};
    for (size_t i = 0; i < live.size(); i++) {
        cseq& c = *live[i]->input_sequence;
        c.set_attr<std::string>(fn_turn, what[turn[i] & 3]);
        if (turn[i] & 1) c.reverse();
        if (turn[i] & 2) c.complement();
    }
}

void famfinder::impl::run(std::vector<tray*>& trays) {
    if (trays.empty()) return;
    const reference_db& db = index->db();
    std::vector<const cseq*> qs;
    std::vector<tray*> live;
    for (tray* t : trays) {
        delete t->alignment_reference;
        t->alignment_reference = nullptr;
        if (t->input_sequence == nullptr) continue;
        if (t->input_sequence->size() < 2) {  // no k-mers at all: no relatives (the device path wants >= 2 bases)
            t->log << "unable to align: too few relatives (0);";
            continue;
        }
        qs.push_back(t->input_sequence);
        live.push_back(t);
    }
    if (live.empty()) return;
    std::vector<uint8_t> masks;
    std::vector<uint64_t> off;
    pack_queries(qs, masks, off);
    do_turn_check(index->handle(), live, masks, off, opts.turn_which);
    if (opts.turn_which != TURN_NONE) pack_queries(qs, masks, off);   // the sequences may have been turned
    sg_fam_params fp;
    sg_default_fam_params(&fp);
    fp.fs_min = opts.fs_min; fp.fs_max = opts.fs_max; fp.fs_msc = opts.fs_msc; fp.fs_msc_max = opts.fs_msc_max;
    fp.fs_min_len = opts.fs_min_len; fp.fs_req_full = opts.fs_req_full; fp.fs_full_len = opts.fs_full_len;
    fp.fs_req_gaps = opts.fs_req_gaps; fp.fs_req = opts.fs_req; fp.leave_query_out = opts.fs_leave_query_out ? 1 : 0;
    std::vector<int64_t> excl;
    if (opts.fs_leave_query_out) {  // remove_query compares names (src/famfinder.cpp:542-544)
        excl.reserve(qs.size());
        for (const cseq* q : qs) excl.push_back(db.indexOf(q->getName()));
    }
    // remove_similar (src/famfinder.cpp:553-556) compares the query at the positions it came with: only a pre-aligned
    // input makes that meaningful, and only a threshold below 1 can remove anything
    std::vector<uint32_t> qcols;
    if (opts.fs_msc_max < 1.0f) {
        qcols.reserve(masks.size());
        for (const cseq* q : qs) for (const aligned_base& b : q->getAlignedBases()) qcols.push_back(b.getPosition());
    }
    const uint32_t stride = std::max(fp.fs_min, fp.fs_max) + fp.fs_req_full + 1;
    const uint32_t nq = (uint32_t)qs.size();
    std::vector<uint32_t> ids((size_t)nq * stride);
    std::vector<float> scores((size_t)nq * stride);
    std::vector<int32_t> fam_n(nq);
    check_sg(sg_family_batch_aligned(index->handle(), masks.data(), qcols.empty() ? nullptr : qcols.data(), off.data(), nq,
                                     excl.empty() ? nullptr : excl.data(), &fp, stride, ids.data(), scores.data(), fam_n.data()),
             "family selection");
    for (uint32_t q = 0; q < nq; q++) {
        tray& t = *live[q];
        if (fam_n[q] < 0) {  // src/famfinder.cpp:486-491
            t.log << "unable to align: too few relatives (<" << opts.fs_req << ");";
            continue;
        }
        t.alignment_reference = new search::result_vector();
        t.alignment_reference->reserve(fam_n[q]);
        std::string famstr;
        famstr.reserve((size_t)fam_n[q] * 24);
        char buf[64];
        for (int32_t i = 0; i < fam_n[q]; i++) {
            const uint32_t id = ids[(size_t)q * stride + i];
            const float sc = scores[(size_t)q * stride + i];
            t.alignment_reference->emplace_back(sc, &db.getCseq(id));
            // "{acc}.{start}:{score:.2f} " (src/famfinder.cpp:458-470); FASTA references have no acc/start fields.
            // The internal engine's scores are k-mer counts: integral, so ".00" needs no float formatting.
            famstr += db.familyLabel(id);
            if (sc >= 0.f && sc < 1e9f && sc == (float)(uint32_t)sc) {
                char* p = buf + sizeof(buf);
                *--p = 0; *--p = ' '; *--p = '0'; *--p = '0'; *--p = '.';
                uint32_t v = (uint32_t)sc;
                do { *--p = (char)('0' + v % 10); v /= 10; } while (v);
                *--p = ':';
                famstr += p;
            } else {
                snprintf(buf, sizeof(buf), ":%.2f ", sc);
                famstr += buf;
            }
        }
        t.input_sequence->set_attr<std::string>(fn_family, famstr);
    }
}

tray famfinder::operator()(const tray& in) {
    tray t(in);
    std::vector<tray*> v{&t};
    pimpl->run(v);
    return t;
}

void famfinder::run(std::vector<tray>& trays) {
    std::vector<tray*> v;
    v.reserve(trays.size());
    for (auto& t : trays) v.push_back(&t);
    pimpl->run(v);
}

}  // namespace sina
