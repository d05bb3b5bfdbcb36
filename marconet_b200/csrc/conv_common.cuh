// Geometry + fused epilogue shared by the convolution kernels (SIMT fp32 and tcgen05).
#pragma once
#include "mn_common.cuh"

struct ConvGeom {
    const float* x; const float* w;
    float* y; float* y2;
    const float* bias; const float* out_scale; const float* residual; const float* y2_scale;
    const int32_t* valid_w; float* ws;
    int N, H, W, Cin, x_cs;
    int KH, KW, sh, sw, ph, pw, Cout;
    int OH, OW, y_cs, y2_cs, res_cs, res_bcast, os_stride, y2s_stride;
    int act; float gain;
    int M, K;
    int ktiles, ktiles_per_split, splits;
};

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

int mn_conv_simt_plan_splits(const ConvGeom& g, int64_t ws_bytes, int requested);
int mn_conv_simt_launch(ConvGeom g, const float* unused, cudaStream_t st);

// tcgen05 path (conv_tc.cu)
int mn_conv_tc_supported(const ConvGeom& g, const char** why);
int mn_conv_tc_launch(const ConvGeom& g, const void* w_hi, const void* w_lo, const float* w_scale, int prec, cudaStream_t st);
// tcgen05 path v2: halo tiles + weight multicast + persistent CTAs (conv_tc2.cu)
int mn_conv_tc2_supported(const ConvGeom& g, const char** why);
int mn_conv_tc2_launch(const ConvGeom& g, const void* w_hi, const void* w_lo, const float* w_scale, int prec, cudaStream_t st);
