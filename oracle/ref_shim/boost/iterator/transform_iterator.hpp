#pragma once
#include <iterator>
namespace boost { template <class F, class It> class transform_iterator { It it; F f; public:
 using iterator_category=std::random_access_iterator_tag; using difference_type=std::ptrdiff_t; using value_type=decltype(std::declval<F>()(*std::declval<It>())); using pointer=void; using reference=value_type;
 transform_iterator(){} transform_iterator(It it, F f):it(it),f(f){}
 reference operator*() const {return f(*it);} transform_iterator& operator++(){++it;return *this;} transform_iterator& operator+=(difference_type n){it+=n;return *this;}
 transform_iterator operator+(difference_type n) const {return transform_iterator(it+n,f);} difference_type operator-(const transform_iterator& o) const {return it-o.it;}
 bool operator==(const transform_iterator& o) const {return it==o.it;} bool operator!=(const transform_iterator& o) const {return it!=o.it;} bool operator<(const transform_iterator& o) const {return it<o.it;} }; }
