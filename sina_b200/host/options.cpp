#include "options.h"

namespace sina {
namespace po {

std::string options_description::usage() const {
    std::ostringstream o;
    if (!caption_.empty()) o << caption_ << ":\n";
    for (const auto& op : opts_) {
        std::string left = "  ";
        if (op.short_name) left += std::string("-") + op.short_name + " [ --" + op.name + " ]";
        else left += "--" + op.name;
        if (op.takes_value) left += " arg";
        if (!op.default_text.empty()) left += " (=" + op.default_text + ")";
        o << left;
        if (left.size() < 38) o << std::string(38 - left.size(), ' ');
        else o << "\n" << std::string(38, ' ');
        o << (op.unsupported.empty() ? op.help : "[not supported] " + op.unsupported) << "\n";
    }
    return o.str();
}

void store(int argc, const char* const* argv, const options_description& desc, variables_map& vm) {
    auto find_long = [&](const std::string& n) -> const option* {
        for (const auto& o : desc.options()) if (o.name == n) return &o;
        return nullptr;
    };
    auto find_short = [&](char c) -> const option* {
        for (const auto& o : desc.options()) if (o.short_name == c) return &o;
        return nullptr;
    };
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        const option* o = nullptr;
        std::string val;
        bool have_val = false;
        if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
            std::string n = a.substr(2);
            const auto eq = n.find('=');
            if (eq != std::string::npos) { val = n.substr(eq + 1); n = n.substr(0, eq); have_val = true; }
            o = find_long(n);
            if (!o) throw std::logic_error("unrecognised option '--" + n + "'");
        } else if (a.size() >= 2 && a[0] == '-' && a != "-") {
            o = find_short(a[1]);
            if (!o) throw std::logic_error(std::string("unrecognised option '-") + a[1] + "'");
            if (a.size() > 2) { val = a.substr(2); have_val = true; }
        } else {
            throw std::logic_error("too many positional options have been specified on the command line: '" + a + "'");
        }
        if (o->takes_value && !have_val) {
            if (i + 1 >= argc) throw std::logic_error("the required argument for option '--" + o->name + "' is missing");
            val = argv[++i];
        }
        if (!o->takes_value && have_val) throw std::logic_error("option '--" + o->name + "' does not take any arguments");
        if (!o->unsupported.empty())
            throw std::logic_error("option '--" + o->name + "' is not supported by sina_b200: " + o->unsupported);
        vm.values[o->name].push_back(val);
        if (o->assign) o->assign(val);
    }
}

}  // namespace po
}  // namespace sina
