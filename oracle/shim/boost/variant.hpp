// Shim (ours): boost::variant -> std::variant (only used for cseq attributes, off the hot path)
#pragma once
#include <variant>
namespace boost {
template <class... T> using variant = std::variant<T...>;
template <class R = void> struct static_visitor { using result_type = R; };
template <class V, class Var> auto apply_visitor(const V& v, Var&& var) {
    return std::visit(v, std::forward<Var>(var));
}
}  // namespace boost
