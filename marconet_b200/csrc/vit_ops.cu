// TextViT helper operators (LayerNorm, token-axis LayerNorm+Linear, fused MHA) and NCHW<->NHWC
// boundary conversion.  All fp32; these are latency/HBM-bound (S<=64 tokens, 512 features).
#include "mn_common.cuh"

namespace {

__global__ void layernorm_kernel(const float* __restrict__ x, float* __restrict__ y, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, int rows, int dim, float eps) {
    mn_pdl_prologue();
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * dim;
    float s = 0.f;
    for (int c = lane; c < dim; c += 32) s += xr[c];
    const float mean = mn_warp_sum(s) / (float)dim;
    float v = 0.f;
    for (int c = lane; c < dim; c += 32) { const float d = xr[c] - mean; v = fmaf(d, d, v); }
    const float rstd = rsqrtf(mn_warp_sum(v) / (float)dim + eps);
    for (int c = lane; c < dim; c += 32) y[(size_t)row * dim + c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
}

// x:[B,T,D] -> LN over T per (b,d) -> out[b,to,d] = sum_t w[to][t]*ln[t] + bias[to]
template <int TMAX>
__global__ void token_mix_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                                 int B, int T, int To, int D, float eps) {
    mn_pdl_prologue();
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (d >= D) return;
    float v[TMAX];
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < TMAX; ++t) { v[t] = t < T ? x[((size_t)b * T + t) * D + d] : 0.f; s += v[t]; }
    const float mean = s / (float)T;
    float q = 0.f;
#pragma unroll
    for (int t = 0; t < TMAX; ++t) if (t < T) { const float dd = v[t] - mean; q = fmaf(dd, dd, q); }
    const float rstd = rsqrtf(q / (float)T + eps);
#pragma unroll
    for (int t = 0; t < TMAX; ++t) if (t < T) v[t] = (v[t] - mean) * rstd * gamma[t] + beta[t];
    for (int to = 0; to < To; ++to) {
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < TMAX; ++t) if (t < T) acc = fmaf(w[(size_t)to * T + t], v[t], acc);
        out[((size_t)b * To + to) * D + d] = acc + bias[to];
    }
}

// One CTA per (batch, head).  S <= 64, dh == 64.
__global__ void __launch_bounds__(256) attention_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                        int S, int heads, float scale) {
    mn_pdl_prologue();
    constexpr int DH = 64, SM = 64;
    extern __shared__ __align__(16) float att_smem[];
    float (*Q)[DH] = reinterpret_cast<float (*)[DH]>(att_smem);
    float (*K)[DH + 1] = reinterpret_cast<float (*)[DH + 1]>(att_smem + SM * DH);
    float (*V)[DH] = reinterpret_cast<float (*)[DH]>(att_smem + SM * DH + SM * (DH + 1));
    float (*P)[SM + 1] = reinterpret_cast<float (*)[SM + 1]>(att_smem + 2 * SM * DH + SM * (DH + 1));
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int inner = heads * DH;
    const float* base = qkv + (size_t)b * S * 3 * inner;
    for (int idx = threadIdx.x; idx < S * DH; idx += blockDim.x) {
        const int i = idx / DH, d = idx % DH;
        const float* r = base + (size_t)i * 3 * inner + h * DH + d;
        Q[i][d] = r[0]; K[i][d] = r[inner]; V[i][d] = r[2 * inner];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < S * S; idx += blockDim.x) {
        const int i = idx / S, j = idx % S;
        float acc = 0.f;
#pragma unroll 16
        for (int d = 0; d < DH; ++d) acc = fmaf(Q[i][d], K[j][d], acc);
        P[i][j] = acc * scale;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = warp; i < S; i += (blockDim.x >> 5)) {
        float m = -INFINITY;
        for (int j = lane; j < S; j += 32) m = fmaxf(m, P[i][j]);
        m = mn_warp_max(m);
        float s = 0.f;
        for (int j = lane; j < S; j += 32) { const float e = expf(P[i][j] - m); P[i][j] = e; s += e; }
        s = mn_warp_sum(s);
        const float inv = 1.f / s;
        for (int j = lane; j < S; j += 32) P[i][j] *= inv;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < S * DH; idx += blockDim.x) {
        const int i = idx / DH, d = idx % DH;
        float acc = 0.f;
        for (int j = 0; j < S; ++j) acc = fmaf(P[i][j], V[j][d], acc);
        out[((size_t)b * S + i) * inner + h * DH + d] = acc;
    }
}

// Small-M linear layer (M <= 64 rows: the 64 / 16 tokens of one text line): y = act(x W + b + residual) * gain.
// One block = all M rows x 16 output columns; K is streamed in 32-wide chunks through a 4-stage cp.async ring so that
// four chunks of weights are in flight per block (the layer is pure HBM/L2 latency: 32..421 blocks, 2 KB of W per chunk).
// No split-K, no second pass: the TextViT's ~45 GEMMs per line are launch/latency bound, not FLOP bound.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__global__ void __launch_bounds__(128) linear_small_m_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ residual,
                                                             float* __restrict__ y, int M, int K, int N, int act, float gain) {
    mn_pdl_prologue();
    constexpr int ST = 4;
    __shared__ __align__(16) float Xs[ST][64][36];
    __shared__ __align__(16) float Ws[ST][32][16];
    const int tid = threadIdx.x;
    const int col = tid & 15, rg = tid >> 4;            // 16 columns x 8 row groups of 8 rows
    const int n0 = blockIdx.x * 16;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    const int nchunks = K / 32;
    auto issue = [&](int ch) {
        if (ch < nchunks) {
            const int k0 = ch * 32, buf = ch % ST;
#pragma unroll
            for (int i = 0; i < 4; ++i) {               // X chunk: 64 rows x 32 floats = 512 x 16 B
                const int idx = tid + i * 128, r = idx >> 3, c4 = (idx & 7) * 4;
                cp_async16(&Xs[buf][r][c4], x + (size_t)(r < M ? r : 0) * K + k0 + c4, r < M);
            }
            const int kr = tid >> 2, c4 = (tid & 3) * 4; // W chunk: 32 rows x 16 floats = 128 x 16 B
            cp_async16(&Ws[buf][kr][c4], w + (size_t)(k0 + kr) * N + n0 + c4, true);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) issue(s);
    for (int ch = 0; ch < nchunks; ++ch) {
        issue(ch + ST - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(ST - 1) : "memory");
        __syncthreads();
        const int buf = ch % ST;
#pragma unroll
        for (int k4 = 0; k4 < 8; ++k4) {
            const float w0 = Ws[buf][k4 * 4][col], w1 = Ws[buf][k4 * 4 + 1][col], w2 = Ws[buf][k4 * 4 + 2][col], w3 = Ws[buf][k4 * 4 + 3][col];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 xv = *reinterpret_cast<const float4*>(&Xs[buf][rg * 8 + i][k4 * 4]);
                acc[i] = fmaf(xv.x, w0, acc[i]); acc[i] = fmaf(xv.y, w1, acc[i]);
                acc[i] = fmaf(xv.z, w2, acc[i]); acc[i] = fmaf(xv.w, w3, acc[i]);
            }
        }
        __syncthreads();
    }
    const int o = n0 + col;
    const float b = bias ? bias[o] : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = rg * 8 + i;
        if (r < M) {
            float v = acc[i] + b;
            if (residual) v += residual[(size_t)r * N + o];
            y[(size_t)r * N + o] = mn_apply_act(v, act) * gain;
        }
    }
}

// 32x32 smem-tiled transposes between [C][HW] and [HW][C] per sample.
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int HW, int y_cs) {
    mn_pdl_prologue();
    __shared__ float t[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, p = p0 + threadIdx.x;
        t[r][threadIdx.x] = (c < C && p < HW) ? x[((size_t)n * C + c) * HW + p] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int p = p0 + r, c = c0 + threadIdx.x;
        if (p < HW && c < C) y[((size_t)n * HW + p) * y_cs + c] = t[threadIdx.x][r];
    }
}
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ x, int x_cs, float* __restrict__ y, int C, int HW) {
    mn_pdl_prologue();
    __shared__ float t[32][33];
    const int n = blockIdx.z;
    const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int p = p0 + r, c = c0 + threadIdx.x;
        t[r][threadIdx.x] = (c < C && p < HW) ? x[((size_t)n * HW + p) * x_cs + c] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int c = c0 + r, p = p0 + threadIdx.x;
        if (p < HW && c < C) y[((size_t)n * C + c) * HW + p] = t[threadIdx.x][r];
    }
}

}  // namespace

extern "C" int mn_layernorm(const float* x, float* y, const float* gamma, const float* beta, int rows, int dim,
                            float eps, void* stream) {
    MN_REQUIRE(x && y && gamma && beta && rows > 0 && dim > 0, "mn_layernorm: bad args");
    MN_CUDA_CHECK((mn_launch(layernorm_kernel, dim3(mn_cdiv(rows, 4)), dim3(128), 0, (cudaStream_t)stream, x, y, gamma, beta, rows, dim, eps)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_linear_small_m(const float* x, const float* w, const float* bias, const float* residual, float* y,
                                 int M, int K, int N, int act, float gain, void* stream) {
    MN_REQUIRE(x && w && y && M > 0 && M <= 64 && K > 0 && K % 32 == 0 && N > 0 && N % 16 == 0, "mn_linear_small_m: needs M<=64, K%32==0, N%16==0");
    MN_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0, "mn_linear_small_m: x, w must be 16-byte aligned");
    MN_CUDA_CHECK((mn_launch(linear_small_m_kernel, dim3(N / 16), dim3(128), 0, (cudaStream_t)stream, x, w, bias, residual, y, M, K, N, act, gain)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_token_mix(const float* x, const float* gamma, const float* beta, const float* w, const float* bias,
                            float* out, int B, int T, int To, int D, float eps, void* stream) {
    MN_REQUIRE(x && gamma && beta && w && bias && out && B > 0 && T > 0 && T <= 64 && To > 0 && D > 0, "mn_token_mix: bad args (T<=64)");
    MN_CUDA_CHECK((mn_launch(token_mix_kernel<64>, dim3(dim3(mn_cdiv(D, 128), B)), dim3(128), 0, (cudaStream_t)stream, x, gamma, beta, w, bias, out, B, T, To, D, eps)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_attention(const float* qkv, float* out, int B, int S, int heads, int dh, float scale, void* stream) {
    MN_REQUIRE(qkv && out && B > 0 && heads > 0, "mn_attention: bad args");
    MN_REQUIRE(S > 0 && S <= 64 && dh == 64, "mn_attention: needs S<=64 and dh==64 (got S=%d dh=%d)", S, dh);
    constexpr int kSmem = (64 * 64 * 2 + 64 * 65 * 2) * (int)sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        MN_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_set = true;
    }
    MN_CUDA_CHECK((mn_launch(attention_kernel, dim3(B * heads), dim3(256), kSmem, (cudaStream_t)stream, qkv, out, S, heads, scale)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_nchw_to_nhwc(const float* x, float* y, int N, int C, int H, int W, int y_cs, void* stream) {
    MN_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0 && y_cs >= C, "mn_nchw_to_nhwc: bad args");
    dim3 grid(mn_cdiv(H * W, 32), mn_cdiv(C, 32), N);
    MN_CUDA_CHECK((mn_launch(nchw_to_nhwc_kernel, dim3(grid), dim3(dim3(32, 8)), 0, (cudaStream_t)stream, x, y, C, H * W, y_cs)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_nhwc_to_nchw(const float* x, int x_cs, float* y, int N, int C, int H, int W, void* stream) {
    MN_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0 && x_cs >= C, "mn_nhwc_to_nchw: bad args");
    dim3 grid(mn_cdiv(H * W, 32), mn_cdiv(C, 32), N);
    MN_CUDA_CHECK((mn_launch(nhwc_to_nchw_kernel, dim3(grid), dim3(dim3(32, 8)), 0, (cudaStream_t)stream, x, x_cs, y, C, H * W)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}
