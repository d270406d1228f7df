#pragma once
#include <cstddef>
#include <map>
namespace boost {
template <class K, class T>
class ptr_map {
    using map_type = std::map<K, T*>;
public:
    using iterator = typename map_type::iterator;
    using const_iterator = typename map_type::const_iterator;
    ptr_map() = default;
    ptr_map(const ptr_map&) = delete;
    ~ptr_map() { for (auto& kv : m_) delete kv.second; }
    std::pair<iterator, bool> insert(K& k, T* p) { return m_.insert({k, p}); }
    std::pair<iterator, bool> insert(const K& k, T* p) { return m_.insert({k, p}); }
    T& at(const K& k) { return *m_.at(k); }
    const T& at(const K& k) const { return *m_.at(k); }
    T& operator[](const K& k) { return *m_[k]; }
    std::size_t count(const K& k) const { return m_.count(k); }
    iterator find(const K& k) { return m_.find(k); }
    iterator begin() { return m_.begin(); }
    iterator end() { return m_.end(); }
    const_iterator begin() const { return m_.begin(); }
    const_iterator end() const { return m_.end(); }
    std::size_t size() const { return m_.size(); }
    bool empty() const { return m_.empty(); }
    std::size_t erase(const K& k) { auto it = m_.find(k); if (it == m_.end()) return 0; delete it->second; m_.erase(it); return 1; }
private:
    map_type m_;
};
}  // namespace boost
