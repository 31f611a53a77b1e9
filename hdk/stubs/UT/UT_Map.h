#pragma once
/* stub */
#include <map>
#include <set>
template <class K, class V> using UT_Map = std::map<K, V>;
template <class K> using UT_Set = std::set<K>;
