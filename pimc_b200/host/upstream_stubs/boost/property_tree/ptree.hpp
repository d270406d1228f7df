#pragma once
#include <cstddef>
#include <string>
#include <utility>
#include <list>
#include "../optional.hpp"
namespace boost { namespace property_tree {
class ptree {
public:
    using value_type = std::pair<const std::string, ptree>;
    using iterator = std::list<value_type>::iterator;
    using const_iterator = std::list<value_type>::const_iterator;
    ptree() = default;
    ptree(const ptree&) = default;
    ptree& operator=(const ptree& o) { if (this != &o) { kids_.clear(); for (const value_type& kv : o.kids_) kids_.emplace_back(kv); } return *this; }
    std::size_t count(const std::string&) const { return 0; }
    template <class T> T get(const std::string&) const { return T{}; }
    template <class T> T get(const std::string&, const T& d) const { return d; }
    ptree& get_child(const std::string&) { return *this; }
    const ptree& get_child(const std::string&) const { return *this; }
    boost::optional<const ptree&> get_child_optional(const std::string&) const { return boost::optional<const ptree&>(); }
    boost::optional<ptree&> get_child_optional(const std::string&) { return boost::optional<ptree&>(); }
    template <class T> T get_value() const { return T{}; }
    const std::string& data() const { static const std::string s; return s; }
    bool empty() const { return true; }
    iterator begin() { return kids_.begin(); }
    iterator end() { return kids_.end(); }
    const_iterator begin() const { return kids_.begin(); }
    const_iterator end() const { return kids_.end(); }
private:
    std::list<value_type> kids_;
};
}}  // namespace boost::property_tree
