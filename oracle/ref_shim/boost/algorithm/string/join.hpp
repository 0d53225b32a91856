#pragma once
#include <string>
namespace boost { namespace algorithm { template <class C> std::string join(const C& c, const std::string& sep){ std::string r; bool f=true; for(auto& s:c){ if(!f) r+=sep; r+=s; f=false;} return r; } } }
