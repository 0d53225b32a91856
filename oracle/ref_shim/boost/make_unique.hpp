// Test-infrastructure shim (NOT product code): the few Boost symbols the reference's headers need,
// so the genuine reference sources under /root/reference compile in an image without Boost.
#pragma once
#include <memory>
namespace boost { template <class T, class... A> std::unique_ptr<T> make_unique(A&&... a){ return std::unique_ptr<T>(new T(std::forward<A>(a)...)); } }
