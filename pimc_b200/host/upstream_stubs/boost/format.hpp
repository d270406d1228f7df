// Stand-in for boost::format: the interface the reference's headers and the adaptors use (syntax check only).
#pragma once
#include <ostream>
#include <string>
namespace boost {
class format {
public:
    format() = default;
    format(const char*) {}
    format(const std::string&) {}
    template <class T> format& operator%(const T&) { return *this; }
    std::string str() const { return std::string(); }
};
inline std::string str(const format& f) { return f.str(); }
inline std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.str(); }
}  // namespace boost
