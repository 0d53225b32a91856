#pragma once
#define BOOST_PP_CAT_I(a,b) a##b
#define BOOST_PP_CAT(a,b) BOOST_PP_CAT_I(a,b)
#define BOOST_PP_VARIADIC_SIZE_I(e0,e1,e2,e3,e4,e5,e6,e7,size,...) size
#define BOOST_PP_VARIADIC_SIZE(...) BOOST_PP_VARIADIC_SIZE_I(__VA_ARGS__,8,7,6,5,4,3,2,1,)
#define BOOST_PP_OVERLOAD(prefix,...) BOOST_PP_CAT(prefix,BOOST_PP_VARIADIC_SIZE(__VA_ARGS__))
