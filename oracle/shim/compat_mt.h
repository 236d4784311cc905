/* As compat.h, plus: neutralise the per-loop omp_set_num_threads(
 * getPhysicalCoreCount()) calls (BackendOpenMP.cpp:166,...,448; the helper
 * returns 1 off-Windows, :76-80) so OMP_NUM_THREADS decides. No source edit. */
#pragma once
#include <omp.h>
#include "compat.h"
#define omp_set_num_threads(x) ((void)0)
