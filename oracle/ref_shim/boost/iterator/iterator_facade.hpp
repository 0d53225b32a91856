#pragma once
#include <iterator>
#include <cstddef>
namespace boost { struct random_access_traversal_tag{}; class iterator_core_access { public:
 template <class I> static typename I::reference dereference(const I& i){return i.dereference();}
 template <class I> static void increment(I& i){i.increment();} template <class I> static void decrement(I& i){i.decrement();}
 template <class I> static void advance(I& i, std::ptrdiff_t n){i.advance(n);} template <class I> static bool equal(const I& a,const I& b){return a.equal(b);}
 template <class I> static std::ptrdiff_t distance_to(const I& a,const I& b){return a.distance_to(b);} };
template <class D, class V, class Cat, class Ref=V&, class Diff=std::ptrdiff_t> class iterator_facade { D& d(){return *static_cast<D*>(this);} const D& d() const {return *static_cast<const D*>(this);} public:
 using iterator_category=std::random_access_iterator_tag; using value_type=typename std::remove_const<V>::type; using reference=Ref; using difference_type=Diff; using pointer=void;
 reference operator*() const {return iterator_core_access::dereference(d());} D& operator++(){iterator_core_access::increment(d());return d();} D operator++(int){D c=d();++*this;return c;}
 D& operator--(){iterator_core_access::decrement(d());return d();} D operator--(int){D c=d();--*this;return c;} D& operator+=(Diff n){iterator_core_access::advance(d(),n);return d();} D& operator-=(Diff n){iterator_core_access::advance(d(),-n);return d();}
 D operator+(Diff n) const {D c=d();c+=n;return c;} D operator-(Diff n) const {D c=d();c-=n;return c;} Diff operator-(const D& o) const {return iterator_core_access::distance_to(o,d());}
 reference operator[](Diff n) const {return *(d()+n);}
 bool operator==(const D& o) const {return iterator_core_access::equal(d(),o);} bool operator!=(const D& o) const {return !(*this==o);}
 bool operator<(const D& o) const {return iterator_core_access::distance_to(d(),o)>0;} bool operator>(const D& o) const {return iterator_core_access::distance_to(d(),o)<0;}
 bool operator<=(const D& o) const {return !(*this>o);} bool operator>=(const D& o) const {return !(*this<o);} }; }
