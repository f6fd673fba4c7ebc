// TEST INFRASTRUCTURE ONLY (oracle/): std-backed stand-in for the Boost.Random names the
// reference uses (src/distributions_boost.cpp:10-14,17,23-24,29-30,62-63,68-69).
// Boost is not installed in this image and the reference does not pin a version, so seed->variate
// parity with upstream binaries is out of reach ("parity unpinned" for that boundary);
// the oracle and the product host code share THIS mapping instead:
//   mt19937                  -> std::mt19937
//   uniform_real_distribution-> std::uniform_real_distribution<double>   (fresh object per call)
//   normal_distribution      -> std::normal_distribution<double>         (fresh object per call)
//   gamma_distribution       -> std::gamma_distribution<double>          (fresh object per call)
#ifndef BRR_ORACLE_SHIM_BOOST_RANDOM_HPP
#define BRR_ORACLE_SHIM_BOOST_RANDOM_HPP
#include <random>
#include <mutex>
#include <shared_mutex>
#include <ctime>
#include <string>
#include <sstream>

namespace boost {
using mt19937 = std::mt19937;

namespace random {
template <class T = double> using gamma_distribution = std::gamma_distribution<T>;
template <class T = double> using uniform_real_distribution = std::uniform_real_distribution<T>;
template <class T = double> using normal_distribution = std::normal_distribution<T>;

template <class Engine, class Dist>
class variate_generator {
public:
    variate_generator(Engine e, Dist d) : m_eng(e), m_dist(d) {}
    typename Dist::result_type operator()() { return m_dist(m_eng); }
private:
    Engine m_eng;   // Engine is a reference type (mt19937&) at every reference call site
    Dist m_dist;
};
} // namespace random

template <class T = double> using normal_distribution = std::normal_distribution<T>;
using random::variate_generator;

template <class Engine, class T = double>
class uniform_01 {
public:
    explicit uniform_01(Engine e) : m_eng(e) {}
    T operator()() { return std::uniform_real_distribution<T>(0, 1)(m_eng); }
private:
    Engine m_eng;
};
} // namespace boost
#endif
