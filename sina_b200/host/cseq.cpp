#include "cseq.h"

#include <algorithm>

#include <cstring>

namespace sina {

const char* const fn_acc = "acc";
const char* const fn_start = "start";
const char* const fn_fullname = "full_name";
const char* const fn_qual = "align_quality_slv";
const char* const fn_idty = "align_ident_slv";
const char* const fn_head = "align_cutoff_head_slv";
const char* const fn_tail = "align_cutoff_tail_slv";
const char* const fn_date = "aligned_slv";
const char* const fn_turn = "turn";
const char* const fn_family = "align_family_slv";
const char* const fn_filter = "align_filter_slv";
const char* const fn_used_rels = "used_rels";
const char* const fn_align_log = "align_log_slv";

namespace {
struct tables {
    int8_t c2m[256];
    tables() {
        memset(c2m, -1, sizeof(c2m));
        const char* chars = "AGCTURYKMSWBDHVN";
        const int masks[] = {1, 2, 4, 8, 8, 1 | 2, 4 | 8, 2 | 8, 1 | 4, 2 | 4, 1 | 8, 2 | 8 | 4, 2 | 1 | 8, 1 | 4 | 8, 2 | 4 | 1, 15};
        for (int i = 0; chars[i]; i++) {
            c2m[(unsigned char)chars[i]] = (int8_t)masks[i];
            c2m[(unsigned char)(chars[i] + 32)] = (int8_t)(masks[i] | 16);
        }
    }
};
const tables T;
const char RNA[] = ".AGRCMSVUWKDYHBN.agrcmsvuwkdyhbn";
}  // namespace

uint8_t base_iupac::from_char(unsigned char c) {
    const int m = T.c2m[c];
    if (m <= 0) throw bad_character_exception(c);
    return (uint8_t)m;
}
char base_iupac::iupac_rna(uint8_t mask) { return RNA[mask & 31]; }
char base_iupac::iupac_dna(uint8_t mask) {
    const char c = RNA[mask & 31];
    return c == 'U' ? 'T' : (c == 'u' ? 't' : c);
}

cseq::cseq(const char* n, const char* data) : name(n ? n : "") {
    if (data) append(data);
}

cseq& cseq::append(const char* str) { return append(str, strlen(str)); }

// src/cseq.cpp:63-77: blanks are skipped, '-' and '.' advance the column, anything else is a base (bad characters throw).
// Runs of '-' are skipped eight at a time: an aligned row is 97 % gaps, and reference databases and --prealigned input
// come as such rows.
cseq& cseq::append(const char* str, size_t n) {
    bases.reserve(bases.size() + std::min<size_t>(n, 2048));
    size_t i = 0;
    while (i < n) {
        const char c = str[i];
        if (c == '-') {
            size_t j = i + 1;
            while (j + 8 <= n) {
                uint64_t w;
                memcpy(&w, str + j, 8);
                if (w != 0x2d2d2d2d2d2d2d2dULL) break;
                j += 8;
            }
            while (j < n && str[j] == '-') j++;
            alignment_width += (uint32_t)(j - i);
            i = j;
            continue;
        }
        i++;
        if (c == ' ' || c == '\t' || c == '\n' || c == '\r') continue;
        if (c != '.') bases.emplace_back(alignment_width, base_iupac::from_char((unsigned char)c));
        alignment_width++;
    }
    return *this;
}

cseq& cseq::append(const aligned_base& ab) {
    if (ab.getPosition() >= alignment_width) {
        bases.push_back(ab);
        alignment_width = ab.getPosition();
    } else {
        bases.emplace_back(alignment_width, ab.getBase());
    }
    return *this;
}

void cseq::setWidth(uint32_t w) {
    if (bases.empty() || w >= bases.back().getPosition() + 1) {
        alignment_width = w;
        return;
    }
    // shrinking (src/cseq.cpp:105-128): never below the base count; the last bases are moved left
    if (w < size()) throw std::runtime_error("Attempted to shrink alignment width below base count");
    uint32_t skip;
    for (skip = 0; skip < size(); skip++)
        if (bases[size() - skip - 1].getPosition() + skip < w) break;
    for (uint32_t i = skip; i > 0; --i) bases[size() - i].setPosition(w - i);
    alignment_width = w;
}

std::string cseq::getBases() const {
    std::string s;
    s.reserve(bases.size());
    for (const auto& b : bases) s.push_back(base_iupac::iupac_rna(b.getBase()));
    return s;
}

std::string cseq::getAligned(bool nodots, bool dna) const {
    std::string aligned;
    aligned.reserve(alignment_width);
    char dot = nodots ? '-' : '.';
    uint32_t cursor = 0;
    for (const auto& b : bases) {
        const uint32_t pos = b.getPosition();
        aligned.append(pos - cursor, dot);
        dot = '-';
        cursor = pos;
        aligned.push_back(dna ? base_iupac::iupac_dna(b.getBase()) : base_iupac::iupac_rna(b.getBase()));
        cursor++;
    }
    if (cursor < alignment_width) {
        if (!nodots) dot = '.';
        aligned.append(alignment_width - cursor, dot);
    }
    return aligned;
}

void cseq::reverse() {
    std::reverse(bases.begin(), bases.end());
    for (auto& b : bases) b.setPosition(alignment_width - 1 - b.getPosition());
}

void cseq::complement() {
    for (auto& b : bases) {
        const uint8_t m = b.getBase();   // A<->T/U, G<->C on the IUPAC bits, case kept
        b = aligned_base(b.getPosition(), (uint8_t)(((m & 2u) << 1) | ((m & 4u) >> 1) | ((m & 1u) << 3) | ((m & 8u) >> 3) | (m & 16u)));
    }
}

void cseq::upperCaseAll() {
    for (auto& b : bases) b.setUpperCase();
}

}  // namespace sina
