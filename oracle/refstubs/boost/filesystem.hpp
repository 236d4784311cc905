#pragma once
// Minimal stand-in for the parts of boost::filesystem the reference's headers and the compiled sources name.
#include <string>
#include <fstream>
#include <ostream>
#include <sys/stat.h>
#include <cstdio>
#include <stdexcept>
#include <unistd.h>
namespace boost { namespace system { class error_code { int v; public: error_code() : v(0) {} int value() const { return v; } void assign(int x) { v = x; } }; } }
namespace boost { namespace filesystem {
class path {
    std::string s;
public:
    path() {}
    path(const std::string &t) : s(t) {}
    path(const char *t) : s(t) {}
    const std::string &string() const { return s; }
    const char *c_str() const { return s.c_str(); }
    bool empty() const { return s.empty(); }
    path filename() const { size_t p = s.find_last_of('/'); return path(p == std::string::npos ? s : s.substr(p + 1)); }
    path extension() const { std::string f = filename().string(); size_t p = f.find_last_of('.'); return path(p == std::string::npos ? std::string() : f.substr(p)); }
    path parent_path() const { size_t p = s.find_last_of('/'); return path(p == std::string::npos ? std::string() : s.substr(0, p)); }
    path stem() const { std::string f = filename().string(); size_t p = f.find_last_of('.'); return path(p == std::string::npos ? f : f.substr(0, p)); }
    path &replace_extension(const path &e = path()) { std::string f = s; size_t p = f.find_last_of('.'); size_t q = f.find_last_of('/'); if (p != std::string::npos && (q == std::string::npos || p > q)) f = f.substr(0, p); s = f + e.string(); return *this; }
    path &remove_filename() { s = parent_path().string(); return *this; }
    path operator/(const path &o) const { return path(s.empty() ? o.s : s + "/" + o.s); }
    path &operator/=(const path &o) { s = s.empty() ? o.s : s + "/" + o.s; return *this; }
    bool operator==(const path &o) const { return s == o.s; }
    bool operator!=(const path &o) const { return s != o.s; }
    bool operator<(const path &o) const { return s < o.s; }
    bool is_absolute() const { return !s.empty() && s[0] == '/'; }
    bool is_complete() const { return is_absolute(); }
};
inline std::ostream &operator<<(std::ostream &os, const path &p) { return os << p.string(); }
inline bool exists(const path &p) { struct stat st; return ::stat(p.c_str(), &st) == 0; }
inline bool is_directory(const path &p) { struct stat st; return ::stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
inline unsigned long file_size(const path &p) { struct stat st; return ::stat(p.c_str(), &st) == 0 ? (unsigned long) st.st_size : 0; }
inline long last_write_time(const path &p) { struct stat st; return ::stat(p.c_str(), &st) == 0 ? (long) st.st_mtime : 0; }
inline bool remove(const path &p) { return ::remove(p.c_str()) == 0; }
inline bool create_directory(const path &p) { return ::mkdir(p.c_str(), 0777) == 0; }
inline long last_write_time(const path &p, boost::system::error_code &ec) { struct stat st; if (::stat(p.c_str(), &st) != 0) { ec.assign(1); return 0; } return (long) st.st_mtime; }
inline void resize_file(const path &p, unsigned long size) { if (::truncate(p.c_str(), (off_t) size) != 0) throw std::runtime_error("resize_file failed"); }
inline path absolute(const path &p) { return p; }
inline path complete(const path &p) { return p; }
inline path current_path() { return path("."); }
class ifstream : public std::ifstream { public: ifstream() {} explicit ifstream(const path &p, std::ios_base::openmode m = std::ios_base::in) : std::ifstream(p.c_str(), m) {} void open(const path &p, std::ios_base::openmode m = std::ios_base::in) { std::ifstream::open(p.c_str(), m); } };
class ofstream : public std::ofstream { public: ofstream() {} explicit ofstream(const path &p, std::ios_base::openmode m = std::ios_base::out) : std::ofstream(p.c_str(), m) {} void open(const path &p, std::ios_base::openmode m = std::ios_base::out) { std::ofstream::open(p.c_str(), m); } };
class fstream : public std::fstream { public: fstream() {} };
} }
