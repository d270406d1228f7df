#pragma once
#include <istream>
#include <string>
#include "ptree.hpp"
namespace boost { namespace property_tree {
namespace xml_parser { enum { no_comments = 1, trim_whitespace = 2 }; }
inline void read_xml(const std::string&, ptree&, int = 0) {}
template <class Ch> void read_xml(std::basic_istream<Ch>&, ptree&, int = 0) {}
}}  // namespace boost::property_tree
