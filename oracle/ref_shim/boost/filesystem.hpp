#pragma once
#include <filesystem>
#include <string>
#include <stdexcept>
#include <system_error>
namespace boost { namespace system { namespace errc { enum errc_t { io_error = 5 }; inline std::error_code make_error_code(errc_t e){ return std::error_code((int)e, std::generic_category()); } } }
namespace filesystem { class path : public std::filesystem::path { public: using std::filesystem::path::path; path(const std::filesystem::path& p):std::filesystem::path(p){} path()=default;
 path& remove_trailing_separator(){ std::string s=this->string(); while(s.size()>1 && s.back()=='/') s.pop_back(); *this=path(s); return *this; }
 path filename() const { return path(std::filesystem::path::filename()); } path& replace_extension(const std::string& e){ std::filesystem::path::replace_extension(e); return *this; } };
 inline path operator/(const path& a, const char* b){ std::filesystem::path r(a); r /= std::filesystem::path(b); return path(r); } inline path operator/(const path& a, const std::string& b){ std::filesystem::path r(a); r /= std::filesystem::path(b); return path(r); }
 inline bool exists(const path& p){ return std::filesystem::exists(p);} inline bool is_directory(const path& p){return std::filesystem::is_directory(p);} inline std::string extension(const path& p){ return p.extension().string(); }
 class filesystem_error : public std::runtime_error { public: filesystem_error(const std::string& m, std::error_code):std::runtime_error(m){} }; } }
