/* Test-infrastructure shim (NOT product code): MSVC-isms used by the
 * reference solver (Backend.hpp:57 __int64; Defs.cpp:53-56 and
 * Solver.cpp:606-609 _vscprintf / vsprintf_s). Force-included with -include. */
#pragma once
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#define __int64 long long
static inline int _vscprintf(const char *fmt, va_list ap) {
    va_list cp; va_copy(cp, ap); int n = vsnprintf(NULL, 0, fmt, cp); va_end(cp); return n;
}
#define vsprintf_s(buf, size, fmt, ap) vsnprintf((buf), (size), (fmt), (ap))
