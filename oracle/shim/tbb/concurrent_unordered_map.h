#pragma once
#include <unordered_map>
namespace tbb { template <class K, class V, class... R> using concurrent_unordered_map = std::unordered_map<K, V, R...>; }
