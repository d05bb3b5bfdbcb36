// Direct fp32 3x3 convolution for very few output channels (Cout <= 4): the RGB head of TSPSRNet
// (conv_final.6: 64 -> 3 at 128x2048, reference networks.py:374).  An implicit-GEMM tile would waste >95 % of its
// N dimension; here a block owns an 8 x 32 pixel patch, stages the (10 x 34) halo in shared memory 16 channels at a
// time, and each thread produces all Cout channels of one pixel.  Bound: HBM (one read of x, 12 bytes written per pixel).
#include "conv_common.cuh"
#include "mn_common.cuh"

namespace {

constexpr int PH = 8, PW = 32, CCH = 16;      // patch height / width, channels per chunk
constexpr int PIX_PITCH = CCH + 4;            // floats per halo pixel (padding keeps float4 reads conflict-free)

template <int COUT>
__global__ void __launch_bounds__(PH * PW) conv3x3_small_cout_kernel(const ConvGeom g) {
    mn_pdl_prologue();
    __shared__ __align__(16) float halo[(PH + 2) * (PW + 2) * PIX_PITCH];
    __shared__ __align__(16) float wsm[9 * CCH * 4];       // [tap][c][4] (COUT padded to 4)
    const int tx = threadIdx.x % PW, ty = threadIdx.x / PW;
    const int tiles_w = (g.W + PW - 1) / PW, tiles_h = (g.H + PH - 1) / PH;
    int b = blockIdx.x;
    const int twi = b % tiles_w; b /= tiles_w;
    const int thi = b % tiles_h;
    const int n = b / tiles_h;
    const int ox0 = twi * PW, oy0 = thi * PH;
    const float* xn = g.x + (size_t)n * g.H * g.W * g.x_cs;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = 0; c0 < g.Cin; c0 += CCH) {
        __syncthreads();
        for (int i = threadIdx.x; i < (PH + 2) * (PW + 2) * (CCH / 4); i += PH * PW) {
            const int q = i % (CCH / 4), pix = i / (CCH / 4);
            const int hx = pix % (PW + 2), hy = pix / (PW + 2);
            const int iy = oy0 + hy - 1, ix = ox0 + hx - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if ((unsigned)iy < (unsigned)g.H && (unsigned)ix < (unsigned)g.W)
                v = __ldg(reinterpret_cast<const float4*>(xn + ((size_t)iy * g.W + ix) * g.x_cs + c0 + q * 4));
            *reinterpret_cast<float4*>(&halo[pix * PIX_PITCH + q * 4]) = v;
        }
        for (int i = threadIdx.x; i < 9 * CCH * 4; i += PH * PW) {
            const int o = i & 3, c = (i >> 2) % CCH, tap = i / (4 * CCH);
            wsm[i] = o < COUT ? __ldg(g.w + ((size_t)tap * g.Cin + c0 + c) * g.Cout + o) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float* hp = &halo[((ty + ky) * (PW + 2) + tx + kx) * PIX_PITCH];
                const float* wp = &wsm[(ky * 3 + kx) * CCH * 4];
#pragma unroll
                for (int q = 0; q < CCH / 4; ++q) {
                    const float4 a = *reinterpret_cast<const float4*>(hp + q * 4);
                    const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 wv = *reinterpret_cast<const float4*>(wp + (q * 4 + j) * 4);
                        acc[0] = fmaf(av[j], wv.x, acc[0]);
                        if (COUT > 1) acc[1] = fmaf(av[j], wv.y, acc[1]);
                        if (COUT > 2) acc[2] = fmaf(av[j], wv.z, acc[2]);
                        if (COUT > 3) acc[3] = fmaf(av[j], wv.w, acc[3]);
                    }
                }
            }
    }
    const int oy = oy0 + ty, ox = ox0 + tx;
    if (oy < g.H && ox < g.W) {
        const int m = (n * g.H + oy) * g.W + ox;
        conv_epilogue4(g, m, 0, acc);
    }
}

}  // namespace

bool mn_conv_small_supported(const ConvGeom& g) {
    return g.KH == 3 && g.KW == 3 && g.sh == 1 && g.sw == 1 && g.ph == 1 && g.pw == 1 && g.Cout <= 4 && g.Cin % 16 == 0 &&
           g.x_cs % 4 == 0 && (reinterpret_cast<uintptr_t>(g.x) & 15) == 0;
}

int mn_conv_small_launch(const ConvGeom& g, cudaStream_t st) {
    const int blocks = g.N * ((g.H + PH - 1) / PH) * ((g.W + PW - 1) / PW);
    switch (g.Cout) {
        case 1: MN_CUDA_CHECK((mn_launch(conv3x3_small_cout_kernel<1>, dim3(blocks), dim3(PH * PW), 0, st, g))); break;
        case 2: MN_CUDA_CHECK((mn_launch(conv3x3_small_cout_kernel<2>, dim3(blocks), dim3(PH * PW), 0, st, g))); break;
        case 3: MN_CUDA_CHECK((mn_launch(conv3x3_small_cout_kernel<3>, dim3(blocks), dim3(PH * PW), 0, st, g))); break;
        default: MN_CUDA_CHECK((mn_launch(conv3x3_small_cout_kernel<4>, dim3(blocks), dim3(PH * PW), 0, st, g))); break;
    }
    MN_LAUNCH_CHECK();
    return MN_OK;
}
