// Shim (ours): lets the reference sources that mention boost::program_options compile without boost.
// The hot-path headers only need the class names; cseq_comparator.cpp also DEFINES its option parsing
// (validate(), get_options_description(), make_from_variables_map()), which the harness never calls: the
// classes below only have to make those definitions compile.
#pragma once
#include <any>
#include <stdexcept>
#include <string>
#include <vector>
namespace boost {
using any = std::any;
namespace program_options {
struct invalid_option_value : std::logic_error {
    using std::logic_error::logic_error;
};
namespace validators {
inline void check_first_occurrence(const boost::any&) {}
inline const std::string& get_single_string(const std::vector<std::string>& v) { return v.at(0); }
}  // namespace validators
template <class T>
struct typed_value {
    typed_value* default_value(const T&, const std::string& = "") { return this; }
};
template <class T>
typed_value<T>* value(T* = nullptr) {
    static typed_value<T> v;
    return &v;
}
inline typed_value<bool>* bool_switch(bool* = nullptr) {
    static typed_value<bool> v;
    return &v;
}
struct options_description_easy_init {
    template <class V>
    options_description_easy_init& operator()(const char*, V*, const char*) { return *this; }
    options_description_easy_init& operator()(const char*, const char*) { return *this; }
};
class options_description {
public:
    options_description() = default;
    explicit options_description(const std::string&) {}
    options_description_easy_init add_options() { return {}; }
    options_description& add(const options_description&) { return *this; }
};
struct variable_value {
    boost::any v;
    template <class T>
    const T& as() const { return *std::any_cast<T>(&v); }
};
class variables_map {
public:
    const variable_value& operator[](const std::string&) const {
        static variable_value x;
        return x;
    }
    size_t count(const std::string&) const { return 0; }
};
}  // namespace program_options
}  // namespace boost
