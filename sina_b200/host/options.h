// Minimal stand-in for the slice of boost::program_options the reference's stages use
// (options_description::add_options, variables_map::count, typed values with defaults, bool switches).
// Unknown options, missing values and bad values throw std::logic_error, which the reference's main()
// reports and turns into exit code 1 (src/sina.cpp:429-438, 595-607).
#ifndef SINA_B200_HOST_OPTIONS_H
#define SINA_B200_HOST_OPTIONS_H
#include <functional>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace sina {
namespace po {

struct option {
    std::string name;       // long name without dashes
    char short_name = 0;
    bool takes_value = true;
    std::string help, default_text;
    std::string unsupported;  // non-empty: recognised but rejected with this reason
    std::function<void(const std::string&)> assign;
};

class variables_map {
public:
    std::map<std::string, std::vector<std::string>> values;
    size_t count(const std::string& n) const { auto it = values.find(n); return it == values.end() ? 0 : it->second.size(); }
    const std::string& operator[](const std::string& n) const { return values.at(n).back(); }
};

template <typename T>
inline void parse_value(const std::string& name, const std::string& text, T* out) {
    std::istringstream in(text);
    T v;
    if (!(in >> v) || !in.eof()) throw std::logic_error("the argument ('" + text + "') for option '--" + name + "' is invalid");
    *out = v;
}
template <>
inline void parse_value<std::string>(const std::string&, const std::string& text, std::string* out) { *out = text; }

class options_description {
public:
    explicit options_description(std::string caption = "") : caption_(std::move(caption)) {}
    // ("name,s", &target, default, help)
    template <typename T>
    options_description& value(const std::string& spec, T* target, const T& dflt, const std::string& help) {
        option o = make(spec, help);
        *target = dflt;
        std::ostringstream d; d << dflt; o.default_text = d.str();
        const std::string n = o.name;
        o.assign = [n, target](const std::string& t) { parse_value<T>(n, t, target); };
        opts_.push_back(o);
        return *this;
    }
    // enum-like values parsed by a user function (throws std::logic_error itself)
    options_description& custom(const std::string& spec, const std::string& dflt, const std::string& help,
                                std::function<void(const std::string&)> assign) {
        option o = make(spec, help);
        o.default_text = dflt;
        o.assign = std::move(assign);
        opts_.push_back(o);
        return *this;
    }
    options_description& flag(const std::string& spec, bool* target, const std::string& help) {
        option o = make(spec, help);
        o.takes_value = false;
        *target = false;
        o.assign = [target](const std::string&) { *target = true; };
        opts_.push_back(o);
        return *this;
    }
    // option of the reference that this build recognises but cannot honour
    options_description& unsupported(const std::string& spec, bool takes_value, const std::string& why) {
        option o = make(spec, why);
        o.takes_value = takes_value;
        o.unsupported = why;
        opts_.push_back(o);
        return *this;
    }
    options_description& add(const options_description& other) {
        for (const auto& o : other.opts_) opts_.push_back(o);
        return *this;
    }
    const std::vector<option>& options() const { return opts_; }
    std::string usage() const;

private:
    static option make(const std::string& spec, const std::string& help) {
        option o;
        const auto comma = spec.find(',');
        o.name = spec.substr(0, comma);
        if (comma != std::string::npos && comma + 1 < spec.size()) o.short_name = spec[comma + 1];
        o.help = help;
        return o;
    }
    std::string caption_;
    std::vector<option> opts_;
};

// parse argv[1..]; fills vm and assigns targets. Throws std::logic_error.
void store(int argc, const char* const* argv, const options_description& desc, variables_map& vm);

}  // namespace po
}  // namespace sina
#endif
