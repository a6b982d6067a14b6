// Minimal stand-in for tbb/cache_aligned_allocator.h -- ORACLE BUILD ONLY.
// Used as a std::vector allocator at source/scene.h:41, source/scene.cpp:12,
// source/main.cpp:315.  Alignment is a performance hint only.
#pragma once
#include <memory>
namespace tbb {
template <typename T>
using cache_aligned_allocator = std::allocator<T>;
}  // namespace tbb
