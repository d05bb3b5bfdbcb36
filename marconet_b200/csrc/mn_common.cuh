// Shared helpers for libmarconet_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/marconet_b200.h"

void mn_set_error(const char* fmt, ...);

#define MN_REQUIRE(cond, ...)                         \
    do {                                              \
        if (!(cond)) {                                \
            mn_set_error(__VA_ARGS__);                \
            return MN_ERR_INVALID;                    \
        }                                             \
    } while (0)

#define MN_CUDA_CHECK(expr)                                                              \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            mn_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return MN_ERR_CUDA;                                                          \
        }                                                                                \
    } while (0)

#define MN_LAUNCH_CHECK() MN_CUDA_CHECK(cudaGetLastError())

static inline int mn_cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t mn_cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

__device__ __forceinline__ float mn_apply_act(float v, int act) {
    switch (act) {
        case MN_ACT_RELU: return fmaxf(v, 0.f);
        case MN_ACT_LRELU02: return v > 0.f ? v : 0.2f * v;
        case MN_ACT_TANH: return tanhf(v);
        case MN_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f));
        case MN_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
        case MN_ACT_RSQRT_EPS: return rsqrtf(v + 1e-8f);
        default: return v;
    }
}

__device__ __forceinline__ float mn_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double mn_warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float mn_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Programmatic dependent launch: every kernel of the library starts with griddepcontrol.wait (wait for the previous kernel in
// the stream to complete and flush -- normal stream semantics) followed by launch_dependents, and is launched with the
// programmatic-stream-serialization attribute, so launch latency / block scheduling of kernel i+1 overlaps the tail of kernel i.
__device__ __forceinline__ void mn_pdl_prologue() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
int mn_pdl_enabled();   // programmatic dependent launch on (default) / off (mn_set_pdl(0) or MN_PDL=0)
template <typename... KArgs, typename... Args>
static inline cudaError_t mn_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = mn_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel: raise it once per (kernel, device), so that one
// process may drive several GPUs.  `done_mask` is a function-local static of the caller (bit d = done on device d).
template <typename F>
static inline cudaError_t mn_ensure_dyn_smem(F* kernel, int bytes, unsigned long long* done_mask) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64 && ((*done_mask >> dev) & 1ull)) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64) *done_mask |= 1ull << dev;
    return e;
}

int mn_num_sms();
int mn_max_ctas();   // 0 = use every SM; >0 = cap for persistent kernels (leaves SMs to concurrent NCCL kernels)
