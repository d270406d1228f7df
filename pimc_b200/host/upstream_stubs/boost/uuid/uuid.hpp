#pragma once
#include <ostream>
namespace boost { namespace uuids {
struct uuid { unsigned char data[16]; };
struct random_generator { uuid operator()() const { return uuid{}; } };
inline std::ostream& operator<<(std::ostream& os, const uuid&) { return os; }
}}
