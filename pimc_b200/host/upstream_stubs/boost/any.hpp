#pragma once
#include <any>
namespace boost {
using any = std::any;
template <class T> T any_cast(const any& a) { return std::any_cast<T>(a); }
}
