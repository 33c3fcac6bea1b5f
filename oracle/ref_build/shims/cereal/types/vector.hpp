// TEST INFRASTRUCTURE ONLY (oracle/_ref build). cereal is not installed; storage.h only names the two binary
// archive classes inside save()/load(), which the oracle never calls. Declaring them is enough to compile find().
#pragma once
#include <iosfwd>
namespace cereal {
struct BinaryOutputArchive {
    explicit BinaryOutputArchive(std::ostream &) {}
    template <class... T> void operator()(T &&...) {}
};
struct BinaryInputArchive {
    explicit BinaryInputArchive(std::istream &) {}
    template <class... T> void operator()(T &&...) {}
};
}
