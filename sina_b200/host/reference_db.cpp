#include "reference_db.h"

#include <zlib.h>

#include "sidx.h"

#include <fstream>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>

namespace sina {

namespace {
std::mutex g_mu;
std::map<std::string, std::unique_ptr<reference_db>> g_dbs;
}  // namespace

void reference_db::pack() {
    row_off.assign(1, 0);
    uint64_t total = 0;
    for (const auto& s : seqs) total += s.size();
    packed_masks.reserve(total);
    packed_cols.reserve(total);
    by_name.clear();
    for (uint32_t i = 0; i < seqs.size(); i++) {
        if (seqs[i].getWidth() > width) width = seqs[i].getWidth();
        for (const auto& b : seqs[i].getAlignedBases()) {
            packed_masks.push_back(b.getBase());
            packed_cols.push_back(b.getPosition());
        }
        row_off.push_back(packed_masks.size());
        by_name.emplace(seqs[i].getName(), i);
    }
    // mseq requires all rows to have the alignment's width (src/mseq.cpp:56-65)
    for (auto& s : seqs) s.setWidth(width);
    labels.clear();
    labels.reserve(seqs.size());
    for (const auto& r : seqs) labels.push_back(r.get_attr_string(fn_acc, r.getName()) + "." + r.get_attr_string(fn_start, "0"));
}

reference_db* reference_db::fromSequences(const std::string& key, std::vector<cseq>&& v) {
    std::lock_guard<std::mutex> lock(g_mu);
    std::unique_ptr<reference_db> db(new reference_db());
    db->filename = key;
    db->seqs = std::move(v);
    db->pack();
    reference_db* p = db.get();
    g_dbs[key] = std::move(db);
    return p;
}

reference_db* reference_db::getDB(const std::string& path) {
    {
        std::lock_guard<std::mutex> lock(g_mu);
        auto it = g_dbs.find(path);
        if (it != g_dbs.end()) return it->second.get();
    }
    // plain or gzip-compressed aligned FASTA (SILVA ships its alignments as .fasta.gz)
    const bool gz = path.size() > 3 && path.compare(path.size() - 3, 3, ".gz") == 0;
    std::ifstream in;
    gzFile zin = nullptr;
    if (gz) {
        zin = gzopen(path.c_str(), "rb");
        if (zin) gzbuffer(zin, 1u << 20);
    } else {
        in.open(path);
    }
    if (gz ? zin == nullptr : !in) throw std::runtime_error("Unable to open reference database '" + path + "'");
    struct zcloser { gzFile f; ~zcloser() { if (f) gzclose(f); } } zguard{zin};
    std::vector<char> zbuf(gz ? 1u << 16 : 0);
    auto next_line = [&](std::string& line) -> bool {
        if (!gz) return (bool)std::getline(in, line);
        line.clear();
        for (;;) {   // gzgets stops at the buffer's end or behind a newline
            if (!gzgets(zin, zbuf.data(), (int)zbuf.size())) {
                int err = 0;
                gzerror(zin, &err);
                if (err != Z_OK && err != Z_STREAM_END) throw std::runtime_error("Error reading compressed reference database '" + path + "'");
                return !line.empty();
            }
            line += zbuf.data();
            if (!line.empty() && line.back() == '\n') { line.pop_back(); return true; }
        }
    };
    std::vector<cseq> v;
    std::string line;
    cseq* cur = nullptr;
    size_t lineno = 0;
    while (next_line(line)) {
        lineno++;
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty() || line[0] == ';') continue;
        if (line[0] == '>') {
            const auto blank = line.find_first_of(" \t");
            v.emplace_back(line.substr(1, blank == std::string::npos ? std::string::npos : blank - 1).c_str());
            cur = &v.back();
            if (blank != std::string::npos) cur->set_attr<std::string>(fn_fullname, line.substr(blank + 1));
        } else if (cur) {
            try {
                cur->append(line);
            } catch (base_iupac::bad_character_exception& e) {
                throw std::runtime_error("reference database '" + path + "' line " + std::to_string(lineno) +
                                         ": character '" + std::string(1, (char)e.character) + "' is not IUPAC");
            }
        }
    }
    if (v.empty()) throw std::runtime_error("reference database '" + path + "' holds no sequences");
    // Index order. In the reference the id of a sequence is its position in query_arb::getSequenceNames()
    // (src/kmer_search.cpp:248-249), which a .sidx index cache records (:289-291). When `<db>.sidx` exists and lists
    // exactly this database's sequences, its order is adopted, so that ids -- and with them the rank order of equal
    // k-mer scores -- are those of the SINA run that built the cache. Otherwise: order of the file.
    sidx::info cache;
    bool have_cache = false;
    try { have_cache = sidx::read(path + ".sidx", cache); } catch (std::exception&) { have_cache = false; }
    if (have_cache && cache.names.size() == v.size()) {
        std::unordered_map<std::string, size_t> at;
        for (size_t i = 0; i < v.size(); i++) at.emplace(v[i].getName(), i);
        std::vector<size_t> order;
        std::vector<char> used(v.size(), 0);
        for (const auto& n : cache.names) {
            auto it = at.find(n);
            if (it == at.end() || used[it->second]) break;
            used[it->second] = 1;
            order.push_back(it->second);
        }
        if (order.size() == v.size()) {
            std::vector<cseq> r;
            r.reserve(v.size());
            for (size_t i : order) r.push_back(std::move(v[i]));
            v.swap(r);
        }
    }
    return fromSequences(path, std::move(v));
}

std::vector<std::string> reference_db::getSequenceNames() const {
    std::vector<std::string> n;
    n.reserve(seqs.size());
    for (const auto& s : seqs) n.push_back(s.getName());
    return n;
}

const cseq& reference_db::getCseq(const std::string& name) const {
    auto it = by_name.find(name);
    if (it == by_name.end()) throw std::runtime_error("sequence '" + name + "' not in reference database");
    return seqs[it->second];
}

int64_t reference_db::indexOf(const std::string& name) const {
    auto it = by_name.find(name);
    return it == by_name.end() ? -1 : (int64_t)it->second;
}

}  // namespace sina
