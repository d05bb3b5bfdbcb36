// TSPGAN (StyleGAN-like structure-prior generator) helper operators.  HBM-bound elementwise /
// reduction kernels; see include/marconet_b200.h for the reference call sites they replace.
#include "mn_common.cuh"

namespace {

// ---------------------------------------------------------------- PixelNorm (networks.py:170-171)
__global__ void pixelnorm_kernel(const float* __restrict__ x, float* __restrict__ y, int N, int C) {
    mn_pdl_prologue();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    const float* xr = x + (size_t)row * C;
    float ss = 0.f;
    for (int c = lane; c < C; c += 32) ss = fmaf(xr[c], xr[c], ss);
    ss = mn_warp_sum(ss);
    const float r = rsqrtf(ss / (float)C + 1e-8f);
    for (int c = lane; c < C; c += 32) y[(size_t)row * C + c] = xr[c] * r;
}

// ---------------------------------------------------------------- SelectText (networks.py:205-215)
__global__ void select_text_kernel(const float* __restrict__ emb, const int64_t* __restrict__ labels,
                                   const float* __restrict__ s, int s_stride, float* __restrict__ out,
                                   int N, int L, int C) {
    mn_pdl_prologue();
    // out: [N, 4, 4*L, C]; one thread per 4 channels of one output pixel
    const int c4 = C >> 2;
    const int64_t total = (int64_t)N * 16 * L * c4;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;     // host guarantees total < 2^31: 32-bit div/mod only
    if (idx >= (uint32_t)total) return;
    const int c = (int)(idx % c4) * 4;
    uint32_t pix = idx / c4;
    const int xw = (int)(pix % (4 * L));
    pix /= (4 * L);
    const int n = (int)(pix / 4);
    const int l = xw >> 2;
    const int64_t lab = labels[(size_t)n * L + l];
    float4 e = *reinterpret_cast<const float4*>(emb + (size_t)lab * C + c);
    if (s) {
        const float* sn = s + (size_t)n * s_stride + c;
        e.x *= sn[0]; e.y *= sn[1]; e.z *= sn[2]; e.w *= sn[3];
    }
    *reinterpret_cast<float4*>(out + (size_t)idx * 4) = e;
}

// ---------------------------------------------------------------- demodulation (networks.py:284-287)
__global__ void demod_kernel(const float* __restrict__ s, int s_stride, const float* __restrict__ wsq,
                             float* __restrict__ demod, int N, int Cin, int Cout) {
    mn_pdl_prologue();
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (o >= Cout) return;
    const float* sn = s + (size_t)n * s_stride;
    float acc = 0.f;
    for (int c = 0; c < Cin; ++c) {
        const float sv = sn[c];
        acc = fmaf(sv * sv, wsq[(size_t)c * Cout + o], acc);
    }
    demod[(size_t)n * Cout + o] = rsqrtf(acc + 1e-8f);
}

// All demodulation tables of a generator pass in ONE launch: grid (cout tiles, N, layers), 64 columns x 4 k-slices.
__global__ void __launch_bounds__(256) demod_batched_kernel(const float* __restrict__ s_all, int s_stride, const mn_demod_desc* __restrict__ descs,
                                                            float* __restrict__ out_all, int out_stride) {
    mn_pdl_prologue();
    const mn_demod_desc d = descs[blockIdx.z];
    const int col = threadIdx.x & 63, ks = threadIdx.x >> 6;
    const int o = blockIdx.x * 64 + col;
    const int n = blockIdx.y;
    __shared__ float red[4][64];
    float acc = 0.f;
    if (o < d.cout) {
        const float* sn = s_all + (size_t)n * s_stride + d.s_off;
        const float* w = d.wsq + o;
#pragma unroll 4
        for (int c = ks; c < d.cin; c += 4) {
            const float sv = sn[c];
            acc = fmaf(sv * sv, w[(size_t)c * d.cout], acc);
        }
    }
    red[ks][col] = acc;
    __syncthreads();
    if (ks == 0 && o < d.cout)
        out_all[(size_t)n * out_stride + d.out_off + o] = rsqrtf((red[0][col] + red[1][col]) + (red[2][col] + red[3][col]) + 1e-8f);
}

// ---------------------------------------------------------------- bilinear x2 (+ per-sample channel scale)
// PyTorch upsample_bilinear2d, align_corners=False, scale 2: src = 0.5*(dst+0.5)-0.5 clamped at 0.
__device__ __forceinline__ void bilin_coords(int o, int size, int& i0, int& i1, float& l1) {
    float src = 0.5f * ((float)o + 0.5f) - 0.5f;
    src = src < 0.f ? 0.f : src;
    i0 = (int)src;
    i1 = i0 + (i0 < size - 1 ? 1 : 0);
    l1 = src - (float)i0;
}

__global__ void resample_modulate_kernel(const float* __restrict__ x, int x_cs, float* __restrict__ y, int y_cs,
                                         const float* __restrict__ s, int s_stride,
                                         int N, int H, int W, int C, int up) {
    mn_pdl_prologue();
    const int OH = up ? 2 * H : H, OW = up ? 2 * W : W;
    const int c4 = C >> 2;
    const int64_t total = (int64_t)N * OH * OW * c4;
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;     // host guarantees total < 2^31: 32-bit div/mod only
    if (idx >= (uint32_t)total) return;
    const int c = (int)(idx % c4) * 4;
    uint32_t pix = idx / c4;
    const int ox = (int)(pix % OW);
    pix /= OW;
    const int oy = (int)(pix % OH);
    const int n = (int)(pix / OH);
    const float* xn = x + (size_t)n * H * W * x_cs + c;
    float4 v;
    if (up) {
        int y0, y1, x0, x1; float ly, lx;
        bilin_coords(oy, H, y0, y1, ly);
        bilin_coords(ox, W, x0, x1, lx);
        const float hy = 1.f - ly, hx = 1.f - lx;
        const float4 a = *reinterpret_cast<const float4*>(xn + ((size_t)y0 * W + x0) * x_cs);
        const float4 b = *reinterpret_cast<const float4*>(xn + ((size_t)y0 * W + x1) * x_cs);
        const float4 cc = *reinterpret_cast<const float4*>(xn + ((size_t)y1 * W + x0) * x_cs);
        const float4 d = *reinterpret_cast<const float4*>(xn + ((size_t)y1 * W + x1) * x_cs);
        v.x = hy * (hx * a.x + lx * b.x) + ly * (hx * cc.x + lx * d.x);
        v.y = hy * (hx * a.y + lx * b.y) + ly * (hx * cc.y + lx * d.y);
        v.z = hy * (hx * a.z + lx * b.z) + ly * (hx * cc.z + lx * d.z);
        v.w = hy * (hx * a.w + lx * b.w) + ly * (hx * cc.w + lx * d.w);
    } else {
        v = *reinterpret_cast<const float4*>(xn + ((size_t)oy * W + ox) * x_cs);
    }
    if (s) {
        const float* sn = s + (size_t)n * s_stride + c;
        v.x *= sn[0]; v.y *= sn[1]; v.z *= sn[2]; v.w *= sn[3];
    }
    *reinterpret_cast<float4*>(y + (((size_t)n * OH + oy) * OW + ox) * y_cs + c) = v;
}

// ---------------------------------------------------------------- ToRGB (networks.py:313-321)
// One warp walks pixels of one sample; the modulated 1x1 weights (3 x C) live in registers.
template <int CPL /* channels per lane = C/32 */>
__global__ void torgb_kernel(const float* __restrict__ x, int x_cs, const float* __restrict__ s, int s_stride,
                             const float* __restrict__ w, const float* __restrict__ bias,
                             const float* __restrict__ skip, float* __restrict__ out,
                             int N, int H, int W, int C) {
    mn_pdl_prologue();
    const int n = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    // lane owns channels { q*128 + lane*4 + j }  (q < CPL/4)
    float wm[3][CPL];
#pragma unroll
    for (int q = 0; q < CPL / 4; ++q)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = q * 128 + lane * 4 + j;
            const float sv = s[(size_t)n * s_stride + c];
#pragma unroll
            for (int o = 0; o < 3; ++o) wm[o][q * 4 + j] = w[(size_t)o * C + c] * sv;
        }
    const int HW = H * W;
    const float* xn = x + (size_t)n * HW * x_cs;
    for (int p = warp; p < HW; p += nwarps) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
        for (int q = 0; q < CPL / 4; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(xn + (size_t)p * x_cs + q * 128 + lane * 4);
            a0 = fmaf(v.x, wm[0][q * 4], a0); a0 = fmaf(v.y, wm[0][q * 4 + 1], a0);
            a0 = fmaf(v.z, wm[0][q * 4 + 2], a0); a0 = fmaf(v.w, wm[0][q * 4 + 3], a0);
            a1 = fmaf(v.x, wm[1][q * 4], a1); a1 = fmaf(v.y, wm[1][q * 4 + 1], a1);
            a1 = fmaf(v.z, wm[1][q * 4 + 2], a1); a1 = fmaf(v.w, wm[1][q * 4 + 3], a1);
            a2 = fmaf(v.x, wm[2][q * 4], a2); a2 = fmaf(v.y, wm[2][q * 4 + 1], a2);
            a2 = fmaf(v.z, wm[2][q * 4 + 2], a2); a2 = fmaf(v.w, wm[2][q * 4 + 3], a2);
        }
        a0 = mn_warp_sum(a0); a1 = mn_warp_sum(a1); a2 = mn_warp_sum(a2);
        if (lane < 3) {
            float v = lane == 0 ? a0 : (lane == 1 ? a1 : a2);
            v += bias[lane];
            if (skip) {
                const int oy = p / W, ox = p - oy * W;
                const int h2 = H >> 1, w2 = W >> 1;
                int y0, y1, x0, x1; float ly, lx;
                bilin_coords(oy, h2, y0, y1, ly);
                bilin_coords(ox, w2, x0, x1, lx);
                const float hy = 1.f - ly, hx = 1.f - lx;
                const float* sk = skip + (size_t)n * h2 * w2 * 3 + lane;
                const float a = sk[((size_t)y0 * w2 + x0) * 3], b = sk[((size_t)y0 * w2 + x1) * 3];
                const float c = sk[((size_t)y1 * w2 + x0) * 3], d = sk[((size_t)y1 * w2 + x1) * 3];
                v += hy * (hx * a + lx * b) + ly * (hx * c + lx * d);
            }
            out[((size_t)n * HW + p) * 3 + lane] = tanhf(v);
        }
    }
}

}  // namespace

extern "C" int mn_pixelnorm(const float* x, float* y, int N, int C, void* stream) {
    MN_REQUIRE(x && y && N > 0 && C > 0, "mn_pixelnorm: bad args");
    MN_CUDA_CHECK((mn_launch(pixelnorm_kernel, dim3(mn_cdiv(N, 4)), dim3(128), 0, (cudaStream_t)stream, x, y, N, C)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_select_text(const float* emb, const int64_t* labels, const float* s, int s_stride,
                              float* out, int N, int L, int C, void* stream) {
    MN_REQUIRE(emb && labels && out && N > 0 && L > 0 && C > 0 && (C & 3) == 0, "mn_select_text: bad args");
    const int64_t total = (int64_t)N * 16 * L * (C >> 2);
    MN_REQUIRE(total < (1ll << 31), "tensor too large for 32-bit indexing");
    MN_CUDA_CHECK((mn_launch(select_text_kernel, dim3((unsigned)mn_cdiv64(total, 256)), dim3(256), 0, (cudaStream_t)stream, emb, labels, s, s_stride, out, N, L, C)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_demod(const float* s, int s_stride, const float* wsq, float* demod, int N, int Cin, int Cout, void* stream) {
    MN_REQUIRE(s && wsq && demod && N > 0 && Cin > 0 && Cout > 0, "mn_demod: bad args");
    dim3 grid(mn_cdiv(Cout, 128), N);
    MN_CUDA_CHECK((mn_launch(demod_kernel, dim3(grid), dim3(128), 0, (cudaStream_t)stream, s, s_stride, wsq, demod, N, Cin, Cout)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_resample_modulate(const float* x, int x_cs, float* y, int y_cs, const float* s, int s_stride,
                                    int N, int H, int W, int C, int up, void* stream) {
    MN_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0, "mn_resample_modulate: bad args");
    MN_REQUIRE((C & 3) == 0 && (x_cs & 3) == 0 && (y_cs & 3) == 0, "mn_resample_modulate: channels must be multiples of 4");
    MN_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "mn_resample_modulate: pointers must be 16B aligned");
    const int OH = up ? 2 * H : H, OW = up ? 2 * W : W;
    const int64_t total = (int64_t)N * OH * OW * (C >> 2);
    MN_REQUIRE(total < (1ll << 31), "tensor too large for 32-bit indexing");
    MN_CUDA_CHECK((mn_launch(resample_modulate_kernel, dim3((unsigned)mn_cdiv64(total, 256)), dim3(256), 0, (cudaStream_t)stream, x, x_cs, y, y_cs, s, s_stride, N, H, W, C, up)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_torgb(const float* x, int x_cs, const float* s, int s_stride, const float* w, const float* bias,
                        const float* skip, float* out, int N, int H, int W, int C, void* stream) {
    MN_REQUIRE(x && s && w && bias && out && N > 0 && H > 0 && W > 0, "mn_torgb: bad args");
    MN_REQUIRE(C % 128 == 0 && C <= 512 && (x_cs & 3) == 0 && ((uintptr_t)x & 15) == 0, "mn_torgb: C must be 128/256/384/512, 16B aligned");
    MN_REQUIRE(!skip || ((H & 1) == 0 && (W & 1) == 0), "mn_torgb: skip needs even H, W");
    const int HW = H * W;
    int blocks = mn_cdiv(HW, 8 * 4);           // 8 warps per block, >= 4 pixels per warp
    const int cap = mn_cdiv(mn_num_sms() * 8, N);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    dim3 grid(blocks, N);
    cudaStream_t st = (cudaStream_t)stream;
    switch (C / 32) {
        case 4: MN_CUDA_CHECK((mn_launch(torgb_kernel<4>, dim3(grid), dim3(256), 0, st, x, x_cs, s, s_stride, w, bias, skip, out, N, H, W, C))); break;
        case 8: MN_CUDA_CHECK((mn_launch(torgb_kernel<8>, dim3(grid), dim3(256), 0, st, x, x_cs, s, s_stride, w, bias, skip, out, N, H, W, C))); break;
        case 12: MN_CUDA_CHECK((mn_launch(torgb_kernel<12>, dim3(grid), dim3(256), 0, st, x, x_cs, s, s_stride, w, bias, skip, out, N, H, W, C))); break;
        case 16: MN_CUDA_CHECK((mn_launch(torgb_kernel<16>, dim3(grid), dim3(256), 0, st, x, x_cs, s, s_stride, w, bias, skip, out, N, H, W, C))); break;
        default: mn_set_error("mn_torgb: unsupported C=%d", C); return MN_ERR_UNSUPPORTED;
    }
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_demod_batched(const float* s_all, int s_stride, const mn_demod_desc* descs, int n_layers, int max_cout,
                                float* out_all, int out_stride, int N, void* stream) {
    MN_REQUIRE(s_all && descs && out_all && n_layers > 0 && max_cout > 0 && N > 0, "mn_demod_batched: bad args");
    MN_CUDA_CHECK((mn_launch(demod_batched_kernel, dim3(dim3(mn_cdiv(max_cout, 64), N, n_layers)), dim3(256), 0, (cudaStream_t)stream, s_all, s_stride, descs, out_all, out_stride)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}
