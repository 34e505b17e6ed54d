// .sidx index cache of the internal k-mer search engine: reader / writer of the reference's file format
// (kmer_search::impl::store / try_load, src/kmer_search.cpp:66-88,278-351; posting lists as vlimap, src/idset.h:279-414).
//
//   idx_header   { u64 magic "SINAKIDX"; u16 version 0; u32 n_sequences; u16 flags = k | nofast << 8 }  (24 bytes with
//                the compiler's padding, as the reference writes the struct)
//   names        n_sequences lines: the reference sequence behind every index id (= the index ORDER)
//   vlimap       ids of the k-mers with a non-empty list
//   vlimap ...   one posting list per such k-mer
//   vlimap       { u32 inc, last, bytesize, size } + bytesize bytes: the ids as distances to the previous id, 7 bits per
//                byte, low bits first, 0x80 = more bytes follow; inc = -1: the list is INVERTED, i.e. holds the ids NOT in
//                it (lists longer than n_sequences / 2, src/kmer_search.cpp:264-266; `size` stays the un-inverted size)
#ifndef SINA_B200_SIDX_H
#define SINA_B200_SIDX_H
#include <cstdint>
#include <string>
#include <vector>

namespace sina {
namespace sidx {

struct info {
    unsigned int k = 0;
    bool nofast = false;
    uint32_t n_sequences = 0;
    std::vector<std::string> names;
};

/// Writes `path`. list_off / ids: every k-mer's ascending reference ids (sg_index_export_lists); n_slots lists.
void write(const std::string& path, unsigned int k, bool nofast, const std::vector<std::string>& names,
           const uint64_t* list_off, const uint32_t* ids, uint64_t n_slots);

/// Reads header and names; with `kmers` / `lists` also every posting list (un-inverted, ascending ids).
/// Returns false if the file does not exist; throws std::runtime_error on a wrong magic / version / truncated file.
bool read(const std::string& path, info& out, std::vector<uint32_t>* kmers = nullptr,
          std::vector<std::vector<uint32_t>>* lists = nullptr);

}  // namespace sidx
}  // namespace sina
#endif
