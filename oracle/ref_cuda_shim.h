// Pre-included (nvcc -include) when compiling the UNMODIFIED reference extension
// models/bricks/ops/cuda/ms_deform_attn_cuda.cu against torch >= 2.x.  TEST INFRASTRUCTURE ONLY.
//
// The reference calls AT_DISPATCH_FLOATING_TYPES(value.type(), ...) (ms_deform_attn_cuda.cu:56,126);
// current torch no longer ships the ::detail::scalar_type overload for DeprecatedTypeProperties, which
// is the only reason the extension fails to build.  This header restores that overload;
// no reference source is copied or edited.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
}  // namespace detail
