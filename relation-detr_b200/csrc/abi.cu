// abi.cu -- error plumbing shared by the C-ABI entry points of librdetr_ops.so.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace rdetr {

static thread_local char g_last_error[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
    return code;
}

int check_cuda(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return RDETR_OK;
    return fail(RDETR_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

int enter_device_of(const void *ptr)
{
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
    if (e != cudaSuccess) return check_cuda(e, "cudaPointerGetAttributes");
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "buffer %p is not device memory (no CPU path exists in this library)", ptr);
    int cur = -1;
    e = cudaGetDevice(&cur);
    if (e != cudaSuccess) return check_cuda(e, "cudaGetDevice");
    if (cur != attr.device) return check_cuda(cudaSetDevice(attr.device), "cudaSetDevice");
    return RDETR_OK;
}

}  // namespace rdetr

extern "C" int rdetr_abi_version(void) { return RDETR_ABI_VERSION; }

extern "C" const char *rdetr_last_error(void) { return rdetr::g_last_error; }
