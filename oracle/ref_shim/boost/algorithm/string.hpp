#pragma once
#include <string>
#include "string/join.hpp"
namespace boost { namespace algorithm { inline void trim(std::string& s){ size_t b=s.find_first_not_of(" \t\r\n"); if(b==std::string::npos){s.clear();return;} size_t e=s.find_last_not_of(" \t\r\n"); s=s.substr(b,e-b+1);} } }
