#pragma once
namespace boost { template <typename T> class scoped_array { T *p; public: explicit scoped_array(T *q = 0) : p(q) {} ~scoped_array() { delete[] p; } T *get() const { return p; } T &operator[](long i) const { return p[i]; } void reset(T *q = 0) { delete[] p; p = q; } private: scoped_array(const scoped_array &); scoped_array &operator=(const scoped_array &); }; }
