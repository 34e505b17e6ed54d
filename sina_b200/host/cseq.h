// Host-side sequence model of the sina_b200 drop-in: the subset of SINA's cseq / aligned_base the hot path
// touches (reference src/cseq.h:66-299, src/aligned_base.h:55-245). Bases are kept as the reference's IUPAC
// bit masks (A=1 G=2 C=4 T/U=8, +16 lowercase) so that a sequence can be handed to the C-ABI without recoding.
#ifndef SINA_B200_HOST_CSEQ_H
#define SINA_B200_HOST_CSEQ_H
#include <cstdint>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace sina {

class base_iupac {
public:
    // thrown for characters outside the IUPAC alphabet (src/aligned_base.h:62-82)
    class bad_character_exception : public std::exception {
    public:
        explicit bad_character_exception(unsigned char c) noexcept : character(c) {}
        const char* what() const noexcept override { return "Character not IUPAC encoded base or gap"; }
        unsigned char character;
    };
    static uint8_t from_char(unsigned char c);            // src/aligned_base.cpp:70-107; throws bad_character_exception
    static char iupac_rna(uint8_t mask);                  // src/aligned_base.cpp:109-114
    static char iupac_dna(uint8_t mask);                  // src/aligned_base.cpp:116-121
};

class aligned_base {
public:
    aligned_base(uint32_t pos = 0, uint8_t mask = 0) : position(pos), base(mask) {}
    uint32_t getPosition() const { return position; }
    void setPosition(uint32_t p) { position = p; }
    uint8_t getBase() const { return base; }              // IUPAC mask incl. the lowercase bit
    bool isLowerCase() const { return (base & 16) != 0; }
    void setLowerCase() { base |= 16; }
    void setUpperCase() { base &= 15; }
private:
    uint32_t position;
    uint8_t base;
};

class cseq {
public:
    cseq() = default;
    cseq(const char* name, const char* data = nullptr);
    const std::string& getName() const { return name; }
    void setName(const std::string& n) { name = n; }
    // src/cseq.cpp:63-77: blanks ignored, '-' and '.' advance the column, anything else must be IUPAC
    cseq& append(const char* str);
    cseq& append(const char* str, size_t n);              // n characters (the FASTA reader hands whole records)
    cseq& append(const std::string& s) { return append(s.c_str()); }
    // src/cseq.cpp:79-95: positions must not decrease; a base placed before the previous one is forced onto it
    cseq& append(const aligned_base& ab);
    void clearSequence() { bases.clear(); alignment_width = 0; }
    void setAlignedBases(const std::vector<aligned_base>& v) { bases = v; }
    void setAlignedBases(std::vector<aligned_base>&& v) { bases = std::move(v); }
    // name and attributes of `o` without its bases (the aligner's working copy, src/align.cpp:320-323, starts from
    // the input sequence and replaces the bases)
    static cseq withoutBases(const cseq& o) { cseq c; c.name = o.name; c.attributes = o.attributes; return c; }
    const std::vector<aligned_base>& getAlignedBases() const { return bases; }
    std::vector<aligned_base>& getAlignedBasesMutable() { return bases; }
    uint32_t size() const { return (uint32_t)bases.size(); }
    uint32_t getWidth() const { return alignment_width; }
    void setWidth(uint32_t w);                             // src/cseq.cpp:98-132 (growing only; shrinking moves bases)
    std::string getBases() const;                          // unaligned RNA string (src/cseq.cpp:176-188)
    std::string getAligned(bool nodots = false, bool dna = false) const;  // src/cseq.cpp:135-174
    void upperCaseAll();
    void reverse();                                        // src/cseq.cpp:284-289: order and positions mirrored
    void complement();                                     // src/cseq.cpp:292-296, src/aligned_base.h:117-124
    bool operator<(const cseq& o) const { return name < o.name; }

    // attributes (the reference keeps a boost::variant map, src/cseq.h:236-262; strings are enough here)
    template <typename T>
    void set_attr(const std::string& key, const T& val) {
        std::ostringstream o;
        o << val;
        attributes[key] = o.str();
    }
    void set_attr(const std::string& key, const std::string& val) { attributes[key] = val; }
    void set_attr(const std::string& key, int val) { attributes[key] = std::to_string(val); }
    template <typename T>
    T get_attr(const std::string& key, const T& dflt = T()) const {
        auto it = attributes.find(key);
        if (it == attributes.end()) return dflt;
        std::istringstream i(it->second);
        T v;
        if (!(i >> v)) return dflt;
        return v;
    }
    std::string get_attr_string(const std::string& key, const std::string& dflt = "") const {
        auto it = attributes.find(key);
        return it == attributes.end() ? dflt : it->second;
    }
    const std::map<std::string, std::string>& get_attrs() const { return attributes; }

private:
    std::string name;
    std::vector<aligned_base> bases;
    uint32_t alignment_width = 0;
    std::map<std::string, std::string> attributes;
};

// attribute names (src/query_arb.cpp:107-126)
extern const char* const fn_turn;
extern const char* const fn_acc;
extern const char* const fn_start;
extern const char* const fn_fullname;
extern const char* const fn_qual;
extern const char* const fn_idty;
extern const char* const fn_head;
extern const char* const fn_tail;
extern const char* const fn_date;
extern const char* const fn_family;
extern const char* const fn_filter;
extern const char* const fn_used_rels;
extern const char* const fn_align_log;

}  // namespace sina
#endif
