// Shim (ours): the two case-insensitive predicates cseq_comparator.cpp's option parsing uses.
#pragma once
#include <cctype>
#include <string>
namespace boost {
namespace algorithm {
inline bool iequals(const std::string& a, const std::string& b) {
    if (a.size() != b.size()) return false;
    for (size_t i = 0; i < a.size(); i++)
        if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
    return true;
}
inline bool istarts_with(const std::string& s, const std::string& prefix) {
    return s.size() >= prefix.size() && iequals(s.substr(0, prefix.size()), prefix);
}
}  // namespace algorithm
}  // namespace boost
