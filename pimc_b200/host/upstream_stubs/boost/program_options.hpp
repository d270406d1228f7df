// Stand-in for the part of boost::program_options the reference's setup / constants code uses (syntax check only:
// every function has the right shape and does nothing).
#pragma once
#include <any>
#include <istream>
#include <map>
#include <ostream>
#include <string>
#include <vector>
namespace boost { namespace program_options {
class variable_value {
public:
    variable_value() = default;
    template <class T> variable_value(const T& v, bool defaulted) : v_(v), defaulted_(defaulted) {}
    bool empty() const { return !v_.has_value(); }
    bool defaulted() const { return defaulted_; }
    std::any& value() { return v_; }
    const std::any& value() const { return v_; }
    template <class T> const T& as() const { static const T t{}; return t; }
    template <class T> T& as() { static T t{}; return t; }
private:
    std::any v_;
    bool defaulted_ = false;
};
class variables_map : public std::map<std::string, variable_value> {
public:
    using base = std::map<std::string, variable_value>;
    const variable_value& operator[](const std::string& k) const { static const variable_value v; auto it = find(k); return it == end() ? v : it->second; }
    variable_value& operator[](const std::string& k) { return base::operator[](k); }
};
template <class T, class Ch = char>
class typed_value {
public:
    typed_value* default_value(const T&) { return this; }
    typed_value* default_value(const T&, const std::string&) { return this; }
    typed_value* implicit_value(const T&) { return this; }
    typed_value* multitoken() { return this; }
    typed_value* composing() { return this; }
    typed_value* required() { return this; }
    typed_value* zero_tokens() { return this; }
};
template <class T> typed_value<T>* value() { static typed_value<T> v; return &v; }
template <class T> typed_value<T>* value(T*) { static typed_value<T> v; return &v; }
inline typed_value<bool>* bool_switch() { static typed_value<bool> v; return &v; }
class options_description;
class options_description_easy_init {
public:
    options_description_easy_init& operator()(const char*, const char*) { return *this; }
    template <class T> options_description_easy_init& operator()(const char*, const typed_value<T>*, const char*) { return *this; }
    template <class T> options_description_easy_init& operator()(const char*, const typed_value<T>*) { return *this; }
};
class options_description {
public:
    options_description() = default;
    options_description(const std::string&) {}
    options_description(const std::string&, unsigned, unsigned = 0) {}
    options_description_easy_init add_options() { return options_description_easy_init(); }
    options_description& add(const options_description&) { return *this; }
};
inline std::ostream& operator<<(std::ostream& os, const options_description&) { return os; }
struct parsed_options {};
inline parsed_options parse_command_line(int, const char* const*, const options_description&) { return {}; }
inline parsed_options parse_command_line(int, char**, const options_description&) { return {}; }
template <class Ch> parsed_options parse_config_file(std::basic_istream<Ch>&, const options_description&, bool = false) { return {}; }
inline parsed_options parse_config_file(const char*, const options_description&, bool = false) { return {}; }
inline void store(const parsed_options&, variables_map&) {}
inline void notify(variables_map&) {}
}}  // namespace boost::program_options
