#pragma once
#include <memory>
namespace boost { template <typename T> class scoped_ptr { T *p; public: explicit scoped_ptr(T *q = 0) : p(q) {} ~scoped_ptr() { delete p; } T *get() const { return p; } T *operator->() const { return p; } T &operator*() const { return *p; } void reset(T *q = 0) { delete p; p = q; } private: scoped_ptr(const scoped_ptr &); scoped_ptr &operator=(const scoped_ptr &); }; }
