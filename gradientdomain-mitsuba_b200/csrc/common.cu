// gdb200 C ABI: misc entry points + error plumbing.
#include "common.h"

namespace gdb200 {

std::string &last_error()
{
    static thread_local std::string msg;
    return msg;
}

int set_error(int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    last_error() = buf;
    return code;
}

int require_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return set_error(GDB200_ERR_NO_DEVICE,
                         "no CUDA device available (%s); gdb200 has no CPU fallback",
                         e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    return GDB200_OK;
}

int device_info(DeviceInfo *out)
{
    if (int rc = require_device()) return rc;
    GDB_CUDA(cudaGetDevice(&out->device));
    GDB_CUDA(cudaDeviceGetAttribute(&out->sms, cudaDevAttrMultiProcessorCount, out->device));
    return GDB200_OK;
}

}  // namespace gdb200

extern "C" {

int gdb200_version(void) { return 101; }

int gdb200_abi_sizes(int *out, int capacity)
{
    const int sizes[] = {(int)sizeof(gdb200_stats), (int)sizeof(gdb200_poisson_config), (int)sizeof(gdb200_camera), (int)sizeof(gdb200_shape),
                         (int)sizeof(gdb200_material), (int)sizeof(gdb200_emitter), (int)sizeof(gdb200_envmap), (int)sizeof(gdb200_scene_desc),
                         (int)sizeof(gdb200_gpt_params), (int)sizeof(gdb200_buffers)};
    const int n = (int)(sizeof(sizes) / sizeof(sizes[0]));
    for (int i = 0; i < n && i < capacity; i++) out[i] = sizes[i];
    return n;
}

const char *gdb200_last_error(void) { return gdb200::last_error().c_str(); }

int gdb200_device_count(int *out_count)
{
    if (!out_count) return gdb200::set_error(GDB200_ERR_ARGUMENT, "out_count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *out_count = n;
    return GDB200_OK;
}

int gdb200_set_device(int device)
{
    if (int rc = gdb200::require_device()) return rc;
    GDB_CUDA(cudaSetDevice(device));
    return GDB200_OK;
}

int gdb200_host_alloc(void **out_ptr, size_t bytes)
{
    if (!out_ptr) return gdb200::set_error(GDB200_ERR_ARGUMENT, "out_ptr is NULL");
    if (int rc = gdb200::require_device()) return rc;
    GDB_CUDA(cudaMallocHost(out_ptr, bytes));
    return GDB200_OK;
}

int gdb200_host_free(void *ptr)
{
    if (!ptr) return GDB200_OK;
    GDB_CUDA(cudaFreeHost(ptr));
    return GDB200_OK;
}

int gdb200_device_alloc(void **out_ptr, size_t bytes)
{
    if (!out_ptr) return gdb200::set_error(GDB200_ERR_ARGUMENT, "out_ptr is NULL");
    if (int rc = gdb200::require_device()) return rc;
    GDB_CUDA(cudaMalloc(out_ptr, bytes));
    return GDB200_OK;
}

int gdb200_device_free(void *ptr)
{
    if (!ptr) return GDB200_OK;
    GDB_CUDA(cudaFree(ptr));
    return GDB200_OK;
}

int gdb200_device_download(void *host_dst, const void *device_src, size_t bytes)
{
    if (!host_dst || !device_src) return gdb200::set_error(GDB200_ERR_ARGUMENT, "NULL pointer");
    GDB_CUDA(cudaMemcpy(host_dst, device_src, bytes, cudaMemcpyDeviceToHost));
    return GDB200_OK;
}

}  // extern "C"
