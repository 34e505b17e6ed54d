#pragma once
#include <sstream>
#include <stdexcept>
namespace boost {
struct bad_lexical_cast : std::runtime_error { bad_lexical_cast() : std::runtime_error("bad lexical cast") {} };
template <class T, class S> T lexical_cast(const S& s) {
    std::stringstream ss; T t;
    if (!(ss << s) || !(ss >> t)) throw bad_lexical_cast();
    return t;
}
}  // namespace boost
