#pragma once
namespace boost {
template <class T> class optional;
template <class T>
class optional<T&> {
public:
    optional() = default;
    optional(T& t) : p_(&t) {}
    explicit operator bool() const { return p_ != nullptr; }
    bool operator!() const { return p_ == nullptr; }
    T& operator*() const { return *p_; }
    T* operator->() const { return p_; }
private:
    T* p_ = nullptr;
};
}  // namespace boost
