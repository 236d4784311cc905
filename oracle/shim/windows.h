/* Test-infrastructure shim (NOT product code): lets the reference's
 * poisson_solver sources compile on Linux without edits.  Provides the three
 * Win32 timer symbols used at Backend.cpp:31,41-47,531-551. */
#pragma once
#include <stdint.h>
#include <time.h>
typedef union { int64_t QuadPart; } LARGE_INTEGER;
static inline int QueryPerformanceFrequency(LARGE_INTEGER *f) { f->QuadPart = 1000000000LL; return 1; }
static inline int QueryPerformanceCounter(LARGE_INTEGER *t) {
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    t->QuadPart = (int64_t)ts.tv_sec * 1000000000LL + ts.tv_nsec; return 1;
}
