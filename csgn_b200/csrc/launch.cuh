// launch.cuh -- kernel launch with Programmatic Dependent Launch (PDL).
//
// The hot kernels of this library are 25-30 us long at the headline size, and ~2.3 us
// of that is launch latency plus the ramp in which CTAs are being placed.  Every kernel
// therefore (a) starts with griddepcontrol.wait -- nothing of global memory is touched
// before the previous kernel in the stream has completed and flushed -- and (b) issues
// griddepcontrol.launch_dependents right away, so that the NEXT kernel of the stream
// (when launched through launch_kernel below) may have its CTAs resident and waiting
// while this one drains.  Kernels launched without the attribute, or after work that is
// not a PDL kernel, behave exactly as before.
#pragma once

#include <cuda_runtime.h>

#include "kernels.cuh"

namespace csgn {

__device__ __forceinline__ void pdl_enter() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline bool pdl_enabled() { return env_long("CSGN_PDL", 1) != 0; }

template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace csgn
