// TEST INFRASTRUCTURE ONLY (oracle/): minimal boost::format stand-in. The hot path never
// formats; only headers (src/options.hpp:11, src/data.hpp:13) pull it in.
#ifndef BRR_ORACLE_SHIM_BOOST_FORMAT_HPP
#define BRR_ORACLE_SHIM_BOOST_FORMAT_HPP
#include <mutex>
#include <shared_mutex>
#include <ctime>
#include <sstream>
#include <string>
#include <ostream>
namespace boost {
class format {
public:
    explicit format(const std::string &f) : m_fmt(f) {}
    template <class T> format &operator%(const T &v) { m_args << v << ' '; return *this; }
    std::string str() const { return m_args.str(); }
private:
    std::string m_fmt;
    std::ostringstream m_args;
};
inline std::ostream &operator<<(std::ostream &os, const format &f) { return os << f.str(); }
} // namespace boost
#endif
