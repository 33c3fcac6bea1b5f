// TEST INFRASTRUCTURE ONLY (oracle/_ref build). TBB is not installed; the reference uses concurrent_vector only
// as a push_back-able container that is later iterated and moved from.
#pragma once
#include <mutex>
#include <vector>
namespace tbb {
template <class T> class concurrent_vector : public std::vector<T> {
    std::mutex m_;
public:
    concurrent_vector() = default;
    concurrent_vector(concurrent_vector &&o) noexcept : std::vector<T>(std::move(static_cast<std::vector<T> &>(o))) {}
    void push_back(T v) { std::scoped_lock l(m_); std::vector<T>::push_back(std::move(v)); }
};
}
