// TEST INFRASTRUCTURE ONLY -- stand-in for the absent Boost header so the reference's
// match4pcsBase.cc (which calls boost::hash_value on std::tuple<int,int,int> at
// /root/reference/src/3rdparty/super4pcs/src/super4pcs/algorithms/match4pcsBase.cc:65-74)
// compiles in place for oracle/_ref.  Only operMode 2 (not shipped) reaches it.
#pragma once
#include <cstddef>
#include <tuple>
namespace boost {
inline void pgp_hash_combine(std::size_t& seed, std::size_t v) {
  seed ^= v + 0x9e3779b97f4a7c15ULL + (seed << 6) + (seed >> 2);
}
template <typename A, typename B, typename C>
inline std::size_t hash_value(const std::tuple<A, B, C>& t) {
  std::size_t seed = 0;
  pgp_hash_combine(seed, static_cast<std::size_t>(std::get<0>(t)));
  pgp_hash_combine(seed, static_cast<std::size_t>(std::get<1>(t)));
  pgp_hash_combine(seed, static_cast<std::size_t>(std::get<2>(t)));
  return seed;
}
}  // namespace boost
