// TEST INFRASTRUCTURE ONLY (oracle/): force-included (-include) in front of the UNMODIFIED reference
// sources under /root/reference/lib/cuda when building oracle/_ref.  The reference calls
// AT_DISPATCH_FLOATING_TYPES(tensor.type(), ...) which no longer converts to c10::ScalarType on
// torch >= 2.x.  Re-defining the macro here lets the reference files
// compile where they lie, without copying or editing them.
#pragma once
#include <torch/extension.h>

namespace voxurf_ref_compat {
inline c10::ScalarType to_scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
inline c10::ScalarType to_scalar_type(c10::ScalarType t) { return t; }
}  // namespace voxurf_ref_compat

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(::voxurf_ref_compat::to_scalar_type(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
