#pragma once
#include <iterator>
#include <cstddef>
namespace boost { template <class T> class counting_iterator { T v; public:
 using iterator_category=std::random_access_iterator_tag; using value_type=T; using difference_type=std::ptrdiff_t; using pointer=const T*; using reference=const T&;
 counting_iterator():v(){} counting_iterator(T v):v(v){}
 reference operator*() const {return v;} counting_iterator& operator++(){++v;return *this;} counting_iterator operator++(int){auto c=*this;++v;return c;}
 counting_iterator& operator--(){--v;return *this;} counting_iterator& operator+=(difference_type n){v+=n;return *this;} counting_iterator& operator-=(difference_type n){v-=n;return *this;}
 counting_iterator operator+(difference_type n) const {return counting_iterator(v+n);} counting_iterator operator-(difference_type n) const {return counting_iterator(v-n);}
 difference_type operator-(const counting_iterator& o) const {return (difference_type)v-(difference_type)o.v;}
 T operator[](difference_type n) const {return v+n;}
 bool operator==(const counting_iterator& o) const {return v==o.v;} bool operator!=(const counting_iterator& o) const {return v!=o.v;}
 bool operator<(const counting_iterator& o) const {return v<o.v;} bool operator>(const counting_iterator& o) const {return v>o.v;}
 bool operator<=(const counting_iterator& o) const {return v<=o.v;} bool operator>=(const counting_iterator& o) const {return v>=o.v;} }; }
