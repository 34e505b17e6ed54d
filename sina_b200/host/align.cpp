#include "align.h"

#include <algorithm>
#include <cstdio>
#include <ctime>
#include <fstream>

#include "../../include/sina_b200.h"
#include "famfinder.h"
#include "kmer_search.h"

namespace sina {

aligner::options* aligner::opts = nullptr;

template <typename E>
static std::function<void(const std::string&)> enum_parser(E* target, std::vector<std::pair<std::string, E>> names, const std::string& what) {
    return [target, names, what](const std::string& v) {
        for (const auto& n : names) if (n.first == v) { *target = n.second; return; }
        throw std::logic_error(what);
    };
}

void aligner::get_options_description(po::options_description& /*main*/, po::options_description& adv) {
    if (!opts) opts = new options();
    po::options_description od("Aligner");
    od.flag("realign", &opts->realign, "do not copy alignment from reference");
    od.custom("overhang", "attach", "select type of overhang placement [*attach*|remove|edge]",
              enum_parser(&opts->overhang, {{"attach", OVERHANG_ATTACH}, {"remove", OVERHANG_REMOVE}, {"edge", OVERHANG_EDGE}},
                          "overhang type must be one of 'attach', 'remove' or 'edge'"));
    od.custom("lowercase", "none", "select which bases to put in lower case [*none*|original|unaligned]",
              enum_parser(&opts->lowercase, {{"none", LOWERCASE_NONE}, {"original", LOWERCASE_ORIGINAL}, {"unaligned", LOWERCASE_UNALIGNED}},
                          "legal lowercase settings are none, original and unaligned"));
    od.custom("insertion", "shift", "handling of insertions not accomodatable by reference alignment [*shift*|forbid|remove]",
              [](const std::string& v) {
                  if (v == "shift") opts->insertion = INSERTION_SHIFT;
                  else if (v == "remove") opts->insertion = INSERTION_REMOVE;  // "using shift" in the reference too (src/cseq.cpp:462-464)
                  else if (v == "forbid") opts->insertion = INSERTION_FORBID;  // transition_aspace_aware (src/align.cpp:466-468)
                  else throw std::logic_error("insertion type must be one of 'shift', 'forbid' or 'remove'");
              });
    od.unsupported("fs-no-graph", false, "profile-vector alignment (pseq) is a test-only path of the reference");
    od.value<float>("fs-weight", &opts->fs_weight, 1.f, "scales weight derived from fs base freq (1)");
    od.value<float>("match-score", &opts->match_score, 2.f, "score awarded for a match (2)");
    od.value<float>("mismatch-score", &opts->mismatch_score, -1.f, "score awarded for a mismatch (-1)");
    od.value<float>("pen-gap", &opts->gap_penalty, 5.f, "gap open penalty (5)");
    od.value<float>("pen-gapext", &opts->gap_ext_penalty, 2.f, "gap extend penalty (2)");
    od.unsupported("debug-graph", false, "graphviz dumps");
    od.unsupported("use-subst-matrix", false, "experimental scoring system of the reference");
    od.flag("write-used-rels", &opts->write_used_rels, "write used reference sequences to field 'used_rels'");
    od.flag("calc-idty", &opts->calc_idty, "calculate highest identity of aligned sequence with any reference");
    adv.add(od);
}

void aligner::validate_vm(po::variables_map& /*vm*/, po::options_description& /*desc*/) {}

// --filter-weights: the positional weights alignment_stats computes from an ARB SAI (src/alignment_stats.cpp:54-112),
// given directly, one per alignment column
static std::vector<float> load_weights(const std::string& path, unsigned int width) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("Unable to open column weights file '" + path + "'");
    std::vector<float> w;
    float x;
    while (in >> x) w.push_back(x);
    if (!in.eof()) throw std::runtime_error("column weights file '" + path + "': not a number after " + std::to_string(w.size()) + " values");
    if (w.size() != width)
        throw std::runtime_error("column weights file '" + path + "' holds " + std::to_string(w.size()) + " values, the alignment has " +
                                 std::to_string(width) + " columns");
    return w;
}

aligner::aligner(int device)
    : index(kmer_search::get_kmer_search(famfinder::opts.database, (int)famfinder::opts.fs_kmer_len, famfinder::opts.fs_no_fast, device)) {
    if (!opts) opts = new options();
    if (!famfinder::opts.filter_weights.empty())
        index->set_column_weights(load_weights(famfinder::opts.filter_weights, index->db().getAlignmentWidth()));
}
aligner::aligner(const aligner& rhs)
    : index(kmer_search::get_kmer_search(famfinder::opts.database, (int)famfinder::opts.fs_kmer_len, famfinder::opts.fs_no_fast, rhs.index->device())) {}
aligner& aligner::operator=(const aligner& /*rhs*/) { return *this; }
aligner::~aligner() { delete index; }

static std::string make_datetime() {  // src/align.cpp:286-297
    time_t t = time(nullptr);
    struct tm lt;
    localtime_r(&t, &lt);
    char buf[50];
    strftime(buf, sizeof(buf), "%F %T", &lt);
    return buf;
}

void aligner::run(std::vector<tray*>& trays, bool rethrow) {
    const reference_db& db = index->db();
    std::vector<tray*> live;
    std::vector<const cseq*> qs;
    std::vector<uint32_t> fam_ids;
    std::vector<uint64_t> fam_off(1, 0);
    for (tray* t : trays) {
        // skip if requirements missing (src/align.cpp:309-318; astats is always absent here)
        if (t->input_sequence == nullptr || t->alignment_reference == nullptr) continue;
        if (t->input_sequence->size() < 2) { t->log << "sequence too short to align;"; continue; }
        live.push_back(t);
        qs.push_back(t->input_sequence);
        for (const auto& r : *t->alignment_reference) fam_ids.push_back(db.indexOf(r.sequence));
        fam_off.push_back(fam_ids.size());
    }
    if (live.empty()) return;
    std::vector<uint8_t> masks;
    std::vector<uint64_t> off;
    pack_queries(qs, masks, off);
    sg_align_params ap;
    sg_default_align_params(&ap);
    ap.match_score = opts->match_score; ap.mismatch_score = opts->mismatch_score; ap.gap_penalty = opts->gap_penalty;
    ap.gap_ext_penalty = opts->gap_ext_penalty; ap.fs_weight = opts->fs_weight; ap.overhang = (int)opts->overhang;
    ap.lowercase = (int)opts->lowercase; ap.insertion = (int)opts->insertion; ap.realign = opts->realign ? 1 : 0;
    std::vector<uint32_t> out_cols(masks.size());
    std::vector<uint8_t> out_masks(masks.size());
    std::vector<sg_align_result> res(live.size());
    if (fam_ids.empty()) fam_ids.push_back(0);
    check_sg(sg_align_batch(index->handle(), masks.data(), off.data(), (uint32_t)live.size(), fam_ids.data(), fam_off.data(),
                            &ap, out_cols.data(), out_masks.data(), res.data()),
             "alignment");
    const std::string now = make_datetime();
    for (size_t q = 0; q < live.size(); q++) {
        tray& t = *live[q];
        const sg_align_result& r = res[q];
        if (r.status == SG_Q_SKIPPED) {  // src/align.cpp:337-348
            t.log << "sequences containing exact candidate removed from family;that's ALL of them. skipping sequence;";
            continue;
        }
        if (r.status == SG_Q_NOSPACE) {  // cseq::fix_duplicate_positions throws (src/cseq.cpp:557-560)
            if (rethrow) throw std::runtime_error("ERROR: no space to left and right?? sequence longer than alignment?!");
            t.log << "ERROR: no space to left and right?? sequence longer than alignment?!;";
            continue;
        }
        if (r.status == SG_Q_LIMIT) {  // this sequence only: its family graph exceeds a device capacity (include/sina_b200.h)
            t.log << "family graph exceeds the GPU aligner's capacity; sequence not aligned;";
            continue;
        }
        cseq* c = new cseq(cseq::withoutBases(*t.input_sequence));  // working copy: name and attributes
        std::vector<aligned_base> v;
        v.reserve(r.n_out);
        for (uint32_t i = 0; i < r.n_out; i++) v.emplace_back(out_cols[off[q] + i], out_masks[off[q] + i]);
        uint32_t W = db.getAlignmentWidth();
        if (!v.empty() && v.back().getPosition() + 1 > W) W = v.back().getPosition() + 1;  // the reference's right-edge quirk
        c->setAlignedBases(std::move(v));
        c->setWidth(W);
        if (r.status == SG_Q_COPIED) {  // src/align.cpp:349-388
            t.log << "copied alignment from template sequence; ";
            c->set_attr(fn_qual, 100);
            c->set_attr(fn_head, 0);
            c->set_attr(fn_tail, 0);
        } else {
            c->set_attr(fn_head, r.head);
            c->set_attr(fn_tail, r.tail);
            c->set_attr(fn_qual, r.qual);
            t.log << "scoring: raw=" << r.raw << ", weight=" << r.sum_weight << ", query-len=" << t.input_sequence->size()
                  << ", aligned-bases=" << r.n_out << ", score=" << r.score << "; ";  // src/mesh.h:733-736
            if (opts->write_used_rels) {
                std::string rels;
                for (const auto& s : *t.alignment_reference) rels += s.sequence->getName() + " ";
                c->set_attr<std::string>(fn_used_rels, rels);
            }
        }
        c->set_attr<std::string>(fn_date, now);
        c->set_attr<std::string>(fn_filter, famfinder::opts.filter_weights);   // astats->getName() in the reference (src/align.cpp:456)
        delete t.aligned_sequence;
        t.aligned_sequence = c;
    }
    if (opts->calc_idty) {
        // --calc-idty (src/align.cpp:443-453): highest identity (optimistic, no correction, relative to the overlap) of the
        // aligned sequence with any relative it was aligned against; 100 for a copied alignment (:380-382)
        std::vector<uint8_t> am;
        std::vector<uint32_t> ac, ids;
        std::vector<uint64_t> aoff(1, 0), roff(1, 0);
        std::vector<tray*> who;
        for (size_t q = 0; q < live.size(); q++) {
            tray& t = *live[q];
            if (res[q].status == SG_Q_COPIED && t.aligned_sequence) { t.aligned_sequence->set_attr<std::string>(fn_idty, "100"); continue; }
            if (res[q].status != SG_Q_ALIGNED || !t.aligned_sequence || t.aligned_sequence->size() < 2) continue;
            std::string qb;
            if (opts->realign) { qb = t.input_sequence->getBases(); for (char& ch : qb) ch = (char)toupper((unsigned char)ch); }
            for (const auto& r : *t.alignment_reference) {
                if (opts->realign) {   // the relatives containing the query were removed from the family (src/align.cpp:337-343)
                    std::string rb = r.sequence->getBases();
                    for (char& ch : rb) ch = (char)toupper((unsigned char)ch);
                    if (rb.find(qb) != std::string::npos) continue;
                }
                ids.push_back(db.indexOf(r.sequence));
            }
            roff.push_back(ids.size());
            for (const aligned_base& b : t.aligned_sequence->getAlignedBases()) { am.push_back(b.getBase()); ac.push_back(b.getPosition()); }
            aoff.push_back(am.size());
            who.push_back(&t);
        }
        if (!who.empty()) {
            std::vector<float> idty(ids.size() + 1);
            if (ids.empty()) ids.push_back(0);
            check_sg(sg_identity_batch(index->handle(), am.data(), ac.data(), aoff.data(), (uint32_t)who.size(), ids.data(), roff.data(),
                                       0, 0, 3, 0, idty.data()),
                     "identity");
            for (size_t i = 0; i < who.size(); i++) {
                float best = 0;
                for (uint64_t j = roff[i]; j < roff[i + 1]; j++) best = std::max(best, idty[j]);   // NaN never wins, as in std::max(idty, x)
                char buf[32];
                snprintf(buf, sizeof(buf), "%.9g", 100.f * best);
                who[i]->aligned_sequence->set_attr<std::string>(fn_idty, buf);
            }
        }
    }
}

tray aligner::operator()(tray t) {
    std::vector<tray*> v{&t};
    run(v, true);
    return t;
}

void aligner::run(std::vector<tray>& trays) {
    std::vector<tray*> v;
    v.reserve(trays.size());
    for (auto& t : trays) v.push_back(&t);
    run(v, false);
}

}  // namespace sina
