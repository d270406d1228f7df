#pragma once
#include <cstddef>
#include <iterator>
#include <vector>
namespace boost {
template <class T>
class ptr_vector {
public:
    class iterator {
    public:
        using iterator_category = std::random_access_iterator_tag;
        using value_type = T;
        using difference_type = std::ptrdiff_t;
        using pointer = T*;
        using reference = T&;
        iterator() = default;
        explicit iterator(typename std::vector<T*>::iterator i) : i_(i) {}
        T& operator*() const { return **i_; }
        T* operator->() const { return *i_; }
        iterator& operator++() { ++i_; return *this; }
        iterator operator++(int) { iterator t = *this; ++i_; return t; }
        iterator& operator--() { --i_; return *this; }
        iterator& operator+=(difference_type n) { i_ += n; return *this; }
        iterator operator+(difference_type n) const { return iterator(i_ + n); }
        iterator operator-(difference_type n) const { return iterator(i_ - n); }
        difference_type operator-(const iterator& o) const { return i_ - o.i_; }
        T& operator[](difference_type n) const { return *i_[n]; }
        bool operator!=(const iterator& o) const { return i_ != o.i_; }
        bool operator==(const iterator& o) const { return i_ == o.i_; }
        bool operator<(const iterator& o) const { return i_ < o.i_; }
        typename std::vector<T*>::iterator base() const { return i_; }
    private:
        typename std::vector<T*>::iterator i_;
    };
    using const_iterator = iterator;
    ptr_vector() = default;
    ptr_vector(const ptr_vector&) = delete;
    ~ptr_vector() { for (T* p : v_) delete p; }
    void push_back(T* p) { v_.push_back(p); }
    T& operator[](std::size_t k) { return *v_[k]; }
    const T& operator[](std::size_t k) const { return *v_[k]; }
    T& at(std::size_t k) { return *v_.at(k); }
    const T& at(std::size_t k) const { return *v_.at(k); }
    T& back() { return *v_.back(); }
    T& front() { return *v_.front(); }
    std::size_t size() const { return v_.size(); }
    bool empty() const { return v_.empty(); }
    void clear() { for (T* p : v_) delete p; v_.clear(); }
    iterator begin() { return iterator(v_.begin()); }
    iterator end() { return iterator(v_.end()); }
    iterator begin() const { return iterator(const_cast<std::vector<T*>&>(v_).begin()); }
    iterator end() const { return iterator(const_cast<std::vector<T*>&>(v_).end()); }
    std::reverse_iterator<iterator> rbegin() { return std::reverse_iterator<iterator>(end()); }
    std::reverse_iterator<iterator> rend() { return std::reverse_iterator<iterator>(begin()); }
    iterator erase(iterator it) { delete &*it; return iterator(v_.erase(it.base())); }
private:
    std::vector<T*> v_;
};
}  // namespace boost
