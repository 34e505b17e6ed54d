// Shim (ours): lets the reference headers that only *mention* boost::program_options compile
// without boost. Only forward declarations are needed on the hot path.
#pragma once
#include <any>
#include <string>
#include <vector>
namespace boost {
using any = std::any;
namespace program_options {
class options_description;
class variables_map;
}  // namespace program_options
}  // namespace boost
