#include "reference_db.h"

#include "rw_fasta.h"
#include "sidx.h"

#include <condition_variable>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
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
    // plain or gzip-compressed aligned FASTA (SILVA ships its alignments as .fasta.gz). One thread cuts the file into
    // records (rw_fasta's block reader), a few threads turn them into sequences: a 50 000-column row is 50 kB of text
    // and a 500 000-row database 25 GB of it.
    // a chunk = the text of up to per_chunk records back to back; chunks are recycled, so the reader copies into warm
    // pages instead of faulting in 25 GB of fresh ones
    struct chunk_t { size_t first = 0; std::string text; std::vector<size_t> end; std::vector<unsigned int> lineno; };
    std::vector<chunk_t> spare;
    std::mutex mu;
    std::condition_variable cv_work, cv_room;
    std::deque<chunk_t> todo;
    bool closed = false;
    std::string failure;
    std::vector<std::vector<cseq>> parts;   // by chunk number
    const size_t per_chunk = 64;
    auto parse = [&](const std::string& rec, size_t at, const size_t stop, unsigned int lineno, cseq& c) {   // one record: title, sequence lines
        bool title = true;
        while (at < stop) {
            const void* nl = memchr(rec.data() + at, '\n', stop - at);
            size_t e = nl ? (size_t)((const char*)nl - rec.data()) : stop;
            const size_t next = e + 1;
            if (e > at && rec[e - 1] == '\r') e--;
            if (title) {
                const std::string line(rec, at, e - at);
                const auto blank = line.find_first_of(" \t");
                c.setName(line.substr(1, blank == std::string::npos ? std::string::npos : blank - 1));
                if (blank != std::string::npos) c.set_attr<std::string>(fn_fullname, line.substr(blank + 1));
                title = false;
            } else if (e > at && rec[at] != ';') {
                try {
                    c.append(rec.data() + at, e - at);
                } catch (base_iupac::bad_character_exception& ex) {
                    throw std::runtime_error("reference database '" + path + "' line " + std::to_string(lineno) +
                                             ": character '" + std::string(1, (char)ex.character) + "' is not IUPAC");
                }
            }
            at = next;
            lineno++;
        }
    };
    auto worker = [&] {
        for (;;) {
            chunk_t ch;
            {
                std::unique_lock<std::mutex> l(mu);
                cv_work.wait(l, [&] { return !todo.empty() || closed; });
                if (todo.empty()) return;
                ch = std::move(todo.front());
                todo.pop_front();
                cv_room.notify_one();
            }
            std::vector<cseq> out(ch.end.size());
            try {
                for (size_t i = 0; i < ch.end.size(); i++) parse(ch.text, i ? ch.end[i - 1] : 0, ch.end[i], ch.lineno[i], out[i]);
            } catch (std::exception& e) {
                std::lock_guard<std::mutex> l(mu);
                if (failure.empty()) failure = e.what();
            }
            std::lock_guard<std::mutex> l(mu);
            const size_t no = ch.first / per_chunk;
            if (parts.size() <= no) parts.resize(no + 1);
            parts[no] = std::move(out);
            ch.text.clear(); ch.end.clear(); ch.lineno.clear();   // capacity kept
            spare.push_back(std::move(ch));
        }
    };
    std::vector<std::thread> pool;
    const unsigned int n_threads = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
    for (unsigned int i = 0; i < n_threads; i++) pool.emplace_back(worker);
    size_t n_records = 0;
    try {
        rw_fasta::reader rd(path, true);   // whole file, whatever --fasta-block says about the query file
        chunk_t ch;
        unsigned int seqno = 0, lineno = 0;
        auto push = [&] {
            std::unique_lock<std::mutex> l(mu);
            cv_room.wait(l, [&] { return todo.size() < 2 * n_threads; });
            todo.push_back(std::move(ch));
            cv_work.notify_one();
            if (spare.empty()) ch = chunk_t();
            else { ch = std::move(spare.back()); spare.pop_back(); }
        };
        for (;;) {
            const bool more = rd.next_record(ch.text, seqno, lineno, true);
            if (!more) break;
            if (ch.end.empty()) ch.first = n_records;
            ch.end.push_back(ch.text.size());
            ch.lineno.push_back(lineno);
            n_records++;
            if (ch.end.size() == per_chunk) push();
        }
        if (!ch.end.empty()) push();
    } catch (std::exception& e) {
        std::lock_guard<std::mutex> l(mu);
        if (failure.empty()) failure = std::string(e.what()).find("Unable to open") == 0 ? "Unable to open reference database '" + path + "'" : e.what();
    }
    { std::lock_guard<std::mutex> l(mu); closed = true; }
    cv_work.notify_all();
    for (auto& t : pool) t.join();
    if (!failure.empty()) throw std::runtime_error(failure);
    std::vector<cseq> v;
    v.reserve(n_records);
    for (auto& part : parts) for (auto& c : part) v.push_back(std::move(c));
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
