// common.cuh -- shared helpers for libpn12_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "pn12_b200.h"

#define PN_EXPORT extern "C" __attribute__((visibility("default")))

namespace pn {

// Thread-local description of the last failure; read through pn_last_error_string().
void set_error(const char* fmt, ...);

inline int finish_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();  // clear the sticky launch error
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return PN_OK;
}

#define PN_REQUIRE(cond, code, ...)      \
    do {                                 \
        if (!(cond)) {                   \
            pn::set_error(__VA_ARGS__);  \
            return (code);               \
        }                                \
    } while (0)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Layout of the per-cloud bucket workspace of pn_ball_grid_build_f32 (ball_grid.cu): bytes per cloud and the offset of the
// cell-sorted float4 (x, y, z, original index) records, for kernels of other files that walk the cloud in bucket order.
size_t ball_grid_cloud_bytes(int N);
size_t ball_grid_sorted_offset();

// Launch with the extended API (dynamic shared memory above 48 KB was enabled by the caller via cudaFuncSetAttribute).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Squared norm exactly as torch.sum(p ** 2, -1) evaluates it on three components:
// ((x*x + y*y) + z*z), each operation rounded to fp32, no FMA contraction.
__device__ __forceinline__ float sqnorm3(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

// square_distance of the reference (pointnet_util.py:37-39) in its CPU rounding sequence:
//   dot = fma(az,bz, fma(ay,by, ax*bx));  d = ((-2*dot) + |a|^2) + |b|^2
// (-2*dot is exact, so fma(-2, dot, sa) rounds exactly like (-2*dot) + sa.)
__device__ __forceinline__ float sqdist_expand(float ax, float ay, float az, float sa, float bx, float by,
                                               float bz, float sb) {
    const float dot = __fmaf_rn(az, bz, __fmaf_rn(ay, by, __fmul_rn(ax, bx)));
    return __fadd_rn(__fmaf_rn(-2.0f, dot, sa), sb);
}

// FPS distance of the reference (pointnet_util.py:80): sum((p - c) ** 2) = ((dx*dx + dy*dy) + dz*dz).
__device__ __forceinline__ float sqdist_diff(float px, float py, float pz, float cx, float cy, float cz) {
    const float dx = __fsub_rn(px, cx), dy = __fsub_rn(py, cy), dz = __fsub_rn(pz, cz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

}  // namespace pn
