// Reference alignment access for the drop-in. The reference reads its --db through libARBDB
// (query_arb::getARBDB / getSequenceNames / getCseq, src/query_arb.cpp:412-476,732-770), which cannot be built
// here; this class offers the same calls over an aligned FASTA file. Reference index i = i-th record of the
// file (the reference's order is an unordered_map iteration, src/query_arb.cpp:732-739, so ties between equal
// k-mer scores cannot be compared with a real ARB run).
#ifndef SINA_B200_HOST_REFERENCE_DB_H
#define SINA_B200_HOST_REFERENCE_DB_H
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "cseq.h"

namespace sina {

class reference_db {
public:
    // process-wide instance per path, like query_arb::getARBDB (src/query_arb.cpp:412-420)
    static reference_db* getDB(const std::string& path);
    // build from sequences already in memory (tests, embedding)
    static reference_db* fromSequences(const std::string& key, std::vector<cseq>&& seqs);

    uint32_t getAlignmentWidth() const { return width; }
    uint32_t getSeqCount() const { return (uint32_t)seqs.size(); }
    std::vector<std::string> getSequenceNames() const;
    const cseq& getCseq(const std::string& name) const;     // throws std::runtime_error if unknown
    const cseq& getCseq(uint32_t index) const { return seqs[index]; }
    int64_t indexOf(const std::string& name) const;         // -1 if unknown
    uint32_t indexOf(const cseq* s) const { return (uint32_t)(s - seqs.data()); }
    const std::string& getFileName() const { return filename; }
    // "{acc}.{start}" of a reference as the family attribute prints it (src/famfinder.cpp:458-470), built once
    const std::string& familyLabel(uint32_t index) const { return labels[index]; }

    // packed form handed to sg_index_create
    const std::vector<uint8_t>& masks() const { return packed_masks; }
    const std::vector<uint32_t>& cols() const { return packed_cols; }
    const std::vector<uint64_t>& offsets() const { return row_off; }

private:
    reference_db() = default;
    void pack();
    std::string filename;
    uint32_t width = 0;
    std::vector<cseq> seqs;
    std::vector<std::string> labels;
    std::unordered_map<std::string, uint32_t> by_name;
    std::vector<uint8_t> packed_masks;
    std::vector<uint32_t> packed_cols;
    std::vector<uint64_t> row_off;
};

}  // namespace sina
#endif
