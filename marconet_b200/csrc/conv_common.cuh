// Geometry + fused epilogue shared by the convolution kernels (SIMT fp32 and tcgen05).
#pragma once
#include "mn_common.cuh"

struct ConvGeom {
    const float* x; const float* w;
    float* y; float* y2;
    const float* bias; const float* out_scale; const float* residual; const float* y2_scale;
    const int32_t* valid_w; float* ws; int64_t ws_bytes;
    const float2* gn_mr; const float* gn_gamma; const float* gn_beta; int gn_swish;   // fused GroupNorm(+swish) on the input (tc2 only)
    int N, H, W, Cin, x_cs;
    int KH, KW, sh, sw, ph, pw, Cout;
    int OH, OW, y_cs, y2_cs, res_cs, res_bcast, os_stride, y2s_stride;
    int act; float gain;
    int M, K;
    int ktiles, ktiles_per_split, splits;
    float x_scale; float* x_absmax; int32_t* range_flag; int32_t range_tag;   // fp16-range management (tensor-core precisions)
};

// Range bookkeeping of the operand-split stage: `amax` is the running max of (bits(|x * x_scale|)) a thread has seen.
__device__ __forceinline__ void conv_range_report(const ConvGeom& g, uint32_t amax, bool fp16_mode) {
    if (!g.x_absmax && !g.range_flag) return;
    amax = __reduce_max_sync(0xffffffffu, amax);
    if ((threadIdx.x & 31) == 0) {
        if (g.x_absmax) atomicMax(reinterpret_cast<unsigned int*>(g.x_absmax), amax);
        // 65504 = 0x477FE000; Inf / NaN >= 0x7F800000
        if (g.range_flag && amax >= (fp16_mode ? 0x477FE000u : 0x7F800000u)) *reinterpret_cast<volatile int32_t*>(g.range_flag) = g.range_tag;
    }
}

// Epilogue for 4 consecutive output channels [o, o+4) of GEMM row m (pixel index in [N,OH,OW]).
//   v = acc*out_scale[n][o] + bias[o] + residual ; v = act(v)*gain ; masked by valid_w ; y, y2 stores.
__device__ __forceinline__ void conv_epilogue4(const ConvGeom& g, int m, int o, float v[4]) {
    const int hw = g.OH * g.OW;
    const int n = m / hw;
    const int nvalid = min(4, g.Cout - o);
    bool masked = false;
    if (g.valid_w) {
        const int ox = m % g.OW;
        masked = ox >= g.valid_w[n];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (j < nvalid) {
            float t = v[j];
            if (g.out_scale) t *= g.out_scale[(size_t)n * g.os_stride + o + j];
            if (g.bias) t += g.bias[o + j];
            if (g.residual) {
                const size_t rm = g.res_bcast ? (size_t)(m - n * hw) : (size_t)m;
                t += g.residual[rm * g.res_cs + o + j];
            }
            t = mn_apply_act(t, g.act) * g.gain;
            v[j] = masked ? 0.f : t;
        }
    }
    const bool full = nvalid == 4;
    if (g.y) {
        float* dst = g.y + (size_t)m * g.y_cs + o;
        if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        else for (int j = 0; j < nvalid; ++j) dst[j] = v[j];
    }
    if (g.y2) {
        float u[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = (j < nvalid && g.y2_scale) ? v[j] * g.y2_scale[(size_t)n * g.y2s_stride + o + j] : v[j];
        float* dst = g.y2 + (size_t)m * g.y2_cs + o;
        if (full && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) *reinterpret_cast<float4*>(dst) = make_float4(u[0], u[1], u[2], u[3]);
        else for (int j = 0; j < nvalid; ++j) dst[j] = u[j];
    }
}

// Vectorised epilogue for the tensor-core kernels: 4 consecutive channels [o,o+4) (o % 4 == 0, Cout % 4 == 0) of pixel m
// belonging to sample n; float4 loads of the per-channel / per-sample vectors, no integer division.
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void conv_epilogue_vec4(const ConvGeom& g, int m, int n, bool masked, int o, float4 v, float4 bias4) {
    if (g.out_scale) {
        const float4 s = ldg4(g.out_scale + (size_t)n * g.os_stride + o);
        v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
    }
    v.x += bias4.x; v.y += bias4.y; v.z += bias4.z; v.w += bias4.w;
    if (g.residual) {
        const size_t rm = g.res_bcast ? (size_t)(m - n * g.OH * g.OW) : (size_t)m;
        const float4 r = ldg4(g.residual + rm * g.res_cs + o);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    if (g.act != MN_ACT_NONE) {
        v.x = mn_apply_act(v.x, g.act); v.y = mn_apply_act(v.y, g.act); v.z = mn_apply_act(v.z, g.act); v.w = mn_apply_act(v.w, g.act);
    }
    v.x *= g.gain; v.y *= g.gain; v.z *= g.gain; v.w *= g.gain;
    if (masked) v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (g.y) *reinterpret_cast<float4*>(g.y + (size_t)m * g.y_cs + o) = v;
    if (g.y2) {
        if (g.y2_scale) {
            const float4 s = ldg4(g.y2_scale + (size_t)n * g.y2s_stride + o);
            v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
        }
        *reinterpret_cast<float4*>(g.y2 + (size_t)m * g.y2_cs + o) = v;
    }
}

int mn_conv_simt_plan_splits(const ConvGeom& g, int64_t ws_bytes, int requested);
int mn_conv_simt_launch(ConvGeom g, const float* unused, cudaStream_t st);
int mn_conv_splitk_reduce_launch(const ConvGeom& g, cudaStream_t st);   // sums g.splits partial tiles in g.ws and runs the fused epilogue

// tcgen05 path (conv_tc.cu)
int mn_conv_tc_supported(const ConvGeom& g, const char** why);
int mn_conv_tc_launch(const ConvGeom& g, const void* w_hi, const void* w_lo, const float* w_scale, int prec, cudaStream_t st);
// tcgen05 path v2: halo tiles + weight multicast + persistent CTAs (conv_tc2.cu)
int mn_conv_tc2_supported(const ConvGeom& g, const char** why);
int mn_conv_tc2_launch(const ConvGeom& g, const void* w_hi, const void* w_lo, const float* w_scale, int prec, cudaStream_t st);
// direct 3x3 conv for Cout <= 4 (conv_small.cu)
bool mn_conv_small_supported(const ConvGeom& g);
int mn_conv_small_launch(const ConvGeom& g, cudaStream_t st);
