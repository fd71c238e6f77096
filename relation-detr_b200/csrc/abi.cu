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

DeviceGuard::DeviceGuard(const void *ptr)
{
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
    if (e != cudaSuccess) { rc_ = check_cuda(e, "cudaPointerGetAttributes"); return; }
    if (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged) {
        rc_ = fail(RDETR_ERR_INVALID_ARGUMENT, "buffer %p is not device memory (no CPU path exists in this library)", ptr);
        return;
    }
    e = cudaGetDevice(&prev_);
    if (e != cudaSuccess) { rc_ = check_cuda(e, "cudaGetDevice"); return; }
    if (prev_ != attr.device) {
        rc_ = check_cuda(cudaSetDevice(attr.device), "cudaSetDevice");
        switched_ = rc_ == RDETR_OK;
    }
}

DeviceGuard::~DeviceGuard()
{
    if (switched_) cudaSetDevice(prev_);
}

}  // namespace rdetr

extern "C" int rdetr_abi_version(void) { return RDETR_ABI_VERSION; }

extern "C" const char *rdetr_last_error(void) { return rdetr::g_last_error; }
