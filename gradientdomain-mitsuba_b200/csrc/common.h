// Shared host-side helpers for the gdb200 CUDA library (error plumbing).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include <string>
#include "../../include/gdb200.h"

namespace gdb200 {

// Thread-local message returned by gdb200_last_error().
std::string &last_error();
int set_error(int code, const char *fmt, ...);

#define GDB_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return gdb200::set_error(GDB200_ERR_CUDA, "%s failed: %s (%s:%d)", #call,    \
                                     cudaGetErrorString(e__), __FILE__, __LINE__);       \
    } while (0)

// Fails loudly when no usable device exists: the product has no CPU path.
int require_device();

struct DeviceInfo { int device; int sms; };
int device_info(DeviceInfo *out);

}  // namespace gdb200
