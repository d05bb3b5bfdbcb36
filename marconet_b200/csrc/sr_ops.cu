// TSPSRNet helper operators: GroupNorm(+swish), AdaIN+concat over per-character windows, and the
// last-writer-wins window write-back.  HBM-bound; see include/marconet_b200.h for call sites.
#include "mn_common.cuh"

namespace {

// ---------------------------------------------------------------- GroupNorm statistics
// One warp = one group of 32 channels of one pixel per step; block walks a pixel chunk.
__global__ void gn_stats_kernel(const float* __restrict__ x, int x_cs, int H, int W, int C, int cpg,
                                const int32_t* __restrict__ valid_w, double* __restrict__ stats, int pix_per_block) {
    const int n = blockIdx.y;
    const int G = C / cpg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int wv = valid_w ? valid_w[n] : W;
    const int HW = H * W;
    const int p_begin = blockIdx.x * pix_per_block;
    const int p_end = min(HW, p_begin + pix_per_block);
    const float* xn = x + (size_t)n * HW * x_cs;
    // channel slots of this lane inside a group: lane, lane+32, ... (cpg is a multiple of 32 here: 32)
    for (int g = warp; g < G; g += nwarp) {
        float s = 0.f, ss = 0.f;
        double ds = 0.0, dss = 0.0;
        int cnt = 0;
        for (int p = p_begin; p < p_end; ++p) {
            const int px = p % W;
            if (px >= wv) continue;
            for (int c = lane; c < cpg; c += 32) {
                const float v = xn[(size_t)p * x_cs + g * cpg + c];
                s += v; ss = fmaf(v, v, ss);
            }
            if (++cnt == 64) { ds += (double)s; dss += (double)ss; s = 0.f; ss = 0.f; cnt = 0; }
        }
        ds += (double)s; dss += (double)ss;
        ds = mn_warp_sum_d(ds); dss = mn_warp_sum_d(dss);
        if (lane == 0) {
            atomicAdd(&stats[((size_t)n * G + g) * 2 + 0], ds);
            atomicAdd(&stats[((size_t)n * G + g) * 2 + 1], dss);
        }
    }
}

__global__ void gn_apply_kernel(const float* __restrict__ x, int x_cs, float* __restrict__ y, int y_cs,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                int N, int H, int W, int C, int cpg, float eps, int swish,
                                const int32_t* __restrict__ valid_w, const double* __restrict__ stats) {
    const int c4 = C >> 2;
    const int64_t total = (int64_t)N * H * W * c4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % c4) * 4;
    const int64_t pix = idx / c4;
    const int px = (int)(pix % W);
    const int n = (int)(pix / ((int64_t)H * W));
    const int wv = valid_w ? valid_w[n] : W;
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (px < wv) {
        const int G = C / cpg, g = c / cpg;
        const double cnt = (double)H * wv * cpg;
        const double mean_d = stats[((size_t)n * G + g) * 2] / cnt;
        double var_d = stats[((size_t)n * G + g) * 2 + 1] / cnt - mean_d * mean_d;
        if (var_d < 0.0) var_d = 0.0;
        const float mean = (float)mean_d;
        const float rstd = (float)(1.0 / sqrt(var_d + (double)eps));
        const float4 v = *reinterpret_cast<const float4*>(x + pix * x_cs + c);
        float t[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float u = (t[j] - mean) * rstd * gamma[c + j] + beta[c + j];
            if (swish) u = u * (1.f / (1.f + expf(-u)));
            t[j] = u;
        }
        o = make_float4(t[0], t[1], t[2], t[3]);
    }
    *reinterpret_cast<float4*>(y + pix * y_cs + c) = o;
}

// ---------------------------------------------------------------- AdaIN + concat
// grid (Nc, C/32), 256 threads: lane = channel inside the 32-chunk, 8 warps stride over window pixels.
__global__ void adain_concat_kernel(const float* __restrict__ prior, int prior_cs, const float* __restrict__ feat, int feat_cs,
                                    const mn_window* __restrict__ win, float* __restrict__ out,
                                    int H, int Wp, int W, int C) {
    const int i = blockIdx.x;
    const int c = blockIdx.y * 32 + (threadIdx.x & 31);
    const int warp = threadIdx.x >> 5;
    const mn_window wn = win[i];
    const int wv = wn.x2 - wn.x1;
    const int npix = H * wv;
    const float* pr = prior + (size_t)i * H * Wp * prior_cs + c;
    const float* ft = feat + (size_t)wn.line * H * W * feat_cs + c;
    float* op = out + (size_t)i * H * Wp * (2 * C);
    __shared__ float red[4][8][32];
    __shared__ float st[4][32];   // pm, ps, lm, ls

    if (wv > 0) {
        float sp = 0.f, sl = 0.f;
        for (int p = warp; p < npix; p += 8) {
            const int yy = p / wv, xx = p - yy * wv;
            sp += pr[((size_t)yy * Wp + wn.y1 + xx) * prior_cs];
            sl += ft[((size_t)yy * W + wn.x1 + xx) * feat_cs];
        }
        red[0][warp][threadIdx.x & 31] = sp;
        red[1][warp][threadIdx.x & 31] = sl;
        __syncthreads();
        if (warp == 0) {
            float a = 0.f, b = 0.f;
            for (int k = 0; k < 8; ++k) { a += red[0][k][threadIdx.x]; b += red[1][k][threadIdx.x]; }
            st[0][threadIdx.x] = a / (float)npix;
            st[2][threadIdx.x] = b / (float)npix;
        }
        __syncthreads();
        const float pm = st[0][threadIdx.x & 31], lm = st[2][threadIdx.x & 31];
        float vp = 0.f, vl = 0.f;
        for (int p = warp; p < npix; p += 8) {
            const int yy = p / wv, xx = p - yy * wv;
            const float a = pr[((size_t)yy * Wp + wn.y1 + xx) * prior_cs] - pm;
            const float b = ft[((size_t)yy * W + wn.x1 + xx) * feat_cs] - lm;
            vp = fmaf(a, a, vp); vl = fmaf(b, b, vl);
        }
        red[2][warp][threadIdx.x & 31] = vp;
        red[3][warp][threadIdx.x & 31] = vl;
        __syncthreads();
        if (warp == 0) {
            float a = 0.f, b = 0.f;
            for (int k = 0; k < 8; ++k) { a += red[2][k][threadIdx.x]; b += red[3][k][threadIdx.x]; }
            st[1][threadIdx.x] = sqrtf(a / (float)(npix - 1) + 1e-5f);
            st[3][threadIdx.x] = sqrtf(b / (float)(npix - 1) + 1e-5f);
        }
        __syncthreads();
    }
    const float pm = st[0][threadIdx.x & 31], ps = st[1][threadIdx.x & 31];
    const float lm = st[2][threadIdx.x & 31], ls = st[3][threadIdx.x & 31];
    for (int p = warp; p < H * Wp; p += 8) {
        const int yy = p / Wp, xx = p - yy * Wp;
        float a = 0.f, b = 0.f;
        if (xx < wv) {
            const float pv = pr[((size_t)yy * Wp + wn.y1 + xx) * prior_cs];
            b = ft[((size_t)yy * W + wn.x1 + xx) * feat_cs];
            a = __fadd_rn(__fmul_rn(__fdiv_rn(pv - pm, ps), ls), lm);
        }
        op[(size_t)p * (2 * C) + c] = a;
        op[(size_t)p * (2 * C) + C + c] = b;
    }
}

// ---------------------------------------------------------------- window write-back
__global__ void window_scatter_kernel(const float* __restrict__ feat, int feat_cs, const float* __restrict__ scale,
                                      const float* __restrict__ shift, const int32_t* __restrict__ owner,
                                      const mn_window* __restrict__ win, float* __restrict__ out, int out_cs,
                                      int B, int H, int W, int Wp, int C) {
    const int c4 = C >> 2;
    const int64_t total = (int64_t)B * H * W * c4;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c = (int)(idx % c4) * 4;
    const int64_t pix = idx / c4;
    const int x = (int)(pix % W);
    const int y = (int)((pix / W) % H);
    const int b = (int)(pix / ((int64_t)H * W));
    const float4 f = *reinterpret_cast<const float4*>(feat + pix * feat_cs + c);
    float4 o = f;
    const int i = owner[(size_t)b * W + x];
    if (i >= 0) {
        const int xx = x - win[i].x1;
        const size_t off = (((size_t)i * H + y) * Wp + xx) * C + c;
        const float4 sc = *reinterpret_cast<const float4*>(scale + off);
        const float4 sh = *reinterpret_cast<const float4*>(shift + off);
        o.x = __fadd_rn(f.x, __fadd_rn(__fmul_rn(f.x, sc.x), sh.x));
        o.y = __fadd_rn(f.y, __fadd_rn(__fmul_rn(f.y, sc.y), sh.y));
        o.z = __fadd_rn(f.z, __fadd_rn(__fmul_rn(f.z, sc.z), sh.z));
        o.w = __fadd_rn(f.w, __fadd_rn(__fmul_rn(f.w, sc.w), sh.w));
    }
    *reinterpret_cast<float4*>(out + pix * out_cs + c) = o;
}

}  // namespace

extern "C" int mn_groupnorm_swish(const float* x, int x_cs, float* y, int y_cs, const float* gamma, const float* beta,
                                  int N, int H, int W, int C, int cpg, float eps, int swish,
                                  const int32_t* valid_w, double* stats_ws, void* stream) {
    MN_REQUIRE(x && y && gamma && beta && stats_ws, "mn_groupnorm_swish: null pointer");
    MN_REQUIRE(N > 0 && H > 0 && W > 0 && C > 0 && cpg > 0 && C % cpg == 0 && cpg % 32 == 0, "mn_groupnorm_swish: bad dims (cpg must be a multiple of 32)");
    MN_REQUIRE((C & 3) == 0 && (x_cs & 3) == 0 && (y_cs & 3) == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0,
               "mn_groupnorm_swish: alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const int G = C / cpg;
    MN_CUDA_CHECK(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * N * G, st));
    const int HW = H * W;
    int blocks = mn_cdiv(mn_num_sms() * 4, N);
    if (blocks > mn_cdiv(HW, 16)) blocks = mn_cdiv(HW, 16);
    if (blocks < 1) blocks = 1;
    const int ppb = mn_cdiv(HW, blocks);
    blocks = mn_cdiv(HW, ppb);
    gn_stats_kernel<<<dim3(blocks, N), 256, 0, st>>>(x, x_cs, H, W, C, cpg, valid_w, stats_ws, ppb);
    MN_LAUNCH_CHECK();
    const int64_t total = (int64_t)N * H * W * (C >> 2);
    gn_apply_kernel<<<(unsigned)mn_cdiv64(total, 256), 256, 0, st>>>(x, x_cs, y, y_cs, gamma, beta, N, H, W, C, cpg, eps, swish, valid_w, stats_ws);
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_adain_concat(const float* prior, int prior_cs, const float* feat, int feat_cs, const mn_window* win,
                               float* out, int Nc, int H, int Wp, int W, int C, void* stream) {
    MN_REQUIRE(prior && feat && win && out && Nc > 0 && H > 0 && Wp > 0 && W > 0 && C > 0 && C % 32 == 0, "mn_adain_concat: bad args");
    adain_concat_kernel<<<dim3(Nc, C / 32), 256, 0, (cudaStream_t)stream>>>(prior, prior_cs, feat, feat_cs, win, out, H, Wp, W, C);
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_window_scatter(const float* feat, int feat_cs, const float* scale, const float* shift,
                                 const int32_t* owner, const mn_window* win, float* out, int out_cs,
                                 int B, int H, int W, int Wp, int C, void* stream) {
    MN_REQUIRE(feat && scale && shift && owner && win && out, "mn_window_scatter: null pointer");
    MN_REQUIRE(B > 0 && H > 0 && W > 0 && Wp > 0 && (C & 3) == 0 && (feat_cs & 3) == 0 && (out_cs & 3) == 0, "mn_window_scatter: bad dims");
    const int64_t total = (int64_t)B * H * W * (C >> 2);
    window_scatter_kernel<<<(unsigned)mn_cdiv64(total, 256), 256, 0, (cudaStream_t)stream>>>(feat, feat_cs, scale, shift, owner, win, out, out_cs, B, H, W, Wp, C);
    MN_LAUNCH_CHECK();
    return MN_OK;
}
