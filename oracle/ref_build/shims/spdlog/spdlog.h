// TEST INFRASTRUCTURE ONLY (oracle/_ref build). spdlog is not installed; logging is a no-op in the oracle.
#pragma once
namespace spdlog {
template <class... A> inline void info(A &&...) {}
template <class... A> inline void error(A &&...) {}
}
