#include "sidx.h"

#include <cstring>
#include <fstream>
#include <stdexcept>

namespace sina {
namespace sidx {

namespace {
const uint64_t MAGIC = 0x5844494b414e4953ull;  // "SINAKIDX", src/kmer_search.cpp:66
const uint16_t VERSION = 0;
const size_t HEADER_BYTES = 24;                 // sizeof(idx_header): u64, u16, (2 pad), u32, u16, (6 pad)

struct vli {   // one vlimap on disk
    uint32_t inc = 1, last = 0, size = 0;
    std::vector<uint8_t> data;
    void push(uint32_t id) {   // vlimap::push_back: distance to the previous id, 7 bits per byte (src/idset.h:279-311)
        uint32_t n = id - last;
        while (n > 127) { data.push_back((uint8_t)(n | 0x80)); n >>= 7; }
        data.push_back((uint8_t)n);
        last = id;
    }
    void write(std::ostream& out) const {   // vlimap::write, src/idset.h:386-398
        const uint32_t head[4] = {inc, last, (uint32_t)data.size(), size};
        out.write((const char*)head, sizeof(head));
        out.write((const char*)data.data(), (std::streamsize)data.size());
    }
    bool read(std::istream& in) {           // vlimap::read, src/idset.h:400-409
        uint32_t head[4];
        if (!in.read((char*)head, sizeof(head))) return false;
        inc = head[0]; last = head[1]; size = head[3];
        data.resize(head[2]);
        return head[2] == 0 || (bool)in.read((char*)data.data(), head[2]);
    }
    std::vector<uint32_t> decode() const {   // the stored ids (distances summed up)
        std::vector<uint32_t> v;
        uint32_t cur = 0;
        for (size_t i = 0; i < data.size();) {
            uint32_t val = data[i] & 0x7f, shift = 7;
            while (data[i] & 0x80) {
                if (++i >= data.size()) throw std::runtime_error("truncated variable-length integer");
                val |= (uint32_t)(data[i] & 0x7f) << shift;
                shift += 7;
            }
            i++;
            cur += val;
            v.push_back(cur);
        }
        return v;
    }
};
}  // namespace

void write(const std::string& path, unsigned int k, bool nofast, const std::vector<std::string>& names,
           const uint64_t* list_off, const uint32_t* ids, uint64_t n_slots) {
    std::ofstream out(path, std::ofstream::binary);
    if (!out) throw std::runtime_error("Unable to write index cache '" + path + "'");
    const uint32_t N = (uint32_t)names.size();
    unsigned char hdr[HEADER_BYTES];
    memset(hdr, 0, sizeof(hdr));
    const uint16_t flags = (uint16_t)((k & 0xff) | (nofast ? 0x100 : 0));
    memcpy(hdr, &MAGIC, 8); memcpy(hdr + 8, &VERSION, 2); memcpy(hdr + 12, &N, 4); memcpy(hdr + 16, &flags, 2);
    out.write((const char*)hdr, sizeof(hdr));
    for (const auto& n : names) out << n << '\n';
    vli nonempty;
    for (uint64_t v = 0; v < n_slots; v++)
        if (list_off[v + 1] > list_off[v]) { nonempty.push((uint32_t)v); nonempty.size++; }
    nonempty.write(out);
    for (uint64_t v = 0; v < n_slots; v++) {
        const uint64_t a = list_off[v], b = list_off[v + 1];
        if (b == a) continue;
        vli l;
        l.size = (uint32_t)(b - a);
        if (b - a > N / 2) {   // inverted: the ids not in the list (vlimap::invert, src/idset.h:367-384)
            l.inc = 0xffffffffu;
            uint32_t next = 0;
            for (uint64_t e = a; e < b; e++) {
                while (next < ids[e]) l.push(next++);
                next = ids[e] + 1;
            }
            while (next < N) l.push(next++);
        } else {
            for (uint64_t e = a; e < b; e++) l.push(ids[e]);
        }
        l.write(out);
    }
    if (!out) throw std::runtime_error("Error writing index cache '" + path + "'");
}

bool read(const std::string& path, info& out, std::vector<uint32_t>* kmers, std::vector<std::vector<uint32_t>>* lists) {
    std::ifstream in(path, std::ifstream::binary);
    if (!in) return false;
    unsigned char hdr[HEADER_BYTES];
    if (!in.read((char*)hdr, sizeof(hdr))) throw std::runtime_error("Index file " + path + " is truncated");
    uint64_t magic; uint16_t vers, flags; uint32_t N;
    memcpy(&magic, hdr, 8); memcpy(&vers, hdr + 8, 2); memcpy(&N, hdr + 12, 4); memcpy(&flags, hdr + 16, 2);
    if (magic != MAGIC) throw std::runtime_error("Index file " + path + " has wrong magic");           // kmer_search.cpp:311-315
    if (vers != VERSION) throw std::runtime_error("Index file " + path + " created by different version");
    out.k = flags & 0xff; out.nofast = (flags >> 8) & 1; out.n_sequences = N;
    out.names.clear();
    for (uint32_t i = 0; i < N; i++) {
        std::string name;
        if (!std::getline(in, name)) throw std::runtime_error("Index file " + path + " is truncated (names)");
        out.names.push_back(name);
    }
    if (!kmers && !lists) return true;
    vli nonempty;
    if (!nonempty.read(in)) throw std::runtime_error("Index file " + path + " is truncated (k-mer map)");
    const std::vector<uint32_t> km = nonempty.decode();
    if (kmers) *kmers = km;
    if (lists) {
        lists->clear();
        for (size_t i = 0; i < km.size(); i++) {
            vli l;
            if (!l.read(in)) throw std::runtime_error("Index file " + path + " is truncated (posting lists)");
            std::vector<uint32_t> v = l.decode();
            if (l.inc == 0xffffffffu || l.inc == 0xffffu) {   // inverted list: complement
                std::vector<uint32_t> r;
                uint32_t next = 0;
                for (uint32_t x : v) { while (next < x) r.push_back(next++); next = x + 1; }
                while (next < N) r.push_back(next++);
                v.swap(r);
            }
            lists->push_back(std::move(v));
        }
    }
    return true;
}

}  // namespace sidx
}  // namespace sina

// C hooks for the tests (ctypes): the writer fed with lists from outside, the reader flattened
extern "C" int sina_sidx_write(const char* path, unsigned int k, int nofast, const char* const* names, uint32_t n,
                               const uint64_t* list_off, const uint32_t* ids, uint64_t n_slots) {
    try {
        sina::sidx::write(path, k, nofast != 0, std::vector<std::string>(names, names + n), list_off, ids, n_slots);
        return 0;
    } catch (std::exception&) { return -1; }
}
// returns the number of postings (or -1); with non-null outputs fills kmers[n_kmers], list_off[n_kmers + 1], ids[]
extern "C" int64_t sina_sidx_read(const char* path, uint32_t* k, uint32_t* nofast, uint32_t* n_sequences, uint32_t* n_kmers,
                                  uint32_t* kmers, uint64_t* list_off, uint32_t* ids, char* names, uint64_t names_cap) {
    try {
        sina::sidx::info inf;
        std::vector<uint32_t> km;
        std::vector<std::vector<uint32_t>> lists;
        if (!sina::sidx::read(path, inf, &km, &lists)) return -1;
        if (k) *k = inf.k;
        if (nofast) *nofast = inf.nofast;
        if (n_sequences) *n_sequences = inf.n_sequences;
        if (n_kmers) *n_kmers = (uint32_t)km.size();
        uint64_t total = 0, np = 0;
        for (size_t i = 0; i < lists.size(); i++) {
            if (kmers) kmers[i] = km[i];
            if (list_off) list_off[i] = total;
            if (ids) for (uint32_t x : lists[i]) ids[total++] = x; else total += lists[i].size();
        }
        if (list_off) list_off[lists.size()] = total;
        if (names) for (const auto& n : inf.names) { if (np + n.size() + 1 > names_cap) break; memcpy(names + np, n.c_str(), n.size()); np += n.size(); names[np++] = '\n'; }
        return (int64_t)total;
    } catch (std::exception&) { return -2; }
}
