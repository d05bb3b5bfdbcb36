// TextViT helper operators (LayerNorm, token-axis LayerNorm+Linear, fused MHA) and NCHW<->NHWC
// boundary conversion.  All fp32; these are latency/HBM-bound (S<=64 tokens, 512 features).
#include <cstdlib>
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
    // four output tokens per trip: four independent 64-deep FMA chains instead of one (the grid is B x D/128 CTAs -- 4 for the
    // TextViT -- so the chain latency was the whole kernel: 33 us); each output keeps its own summation order (bit-identical)
    int to = 0;
    for (; to + 4 <= To; to += 4) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* w0 = w + (size_t)to * T;
#pragma unroll
        for (int t = 0; t < TMAX; ++t)
            if (t < T) {
                a0 = fmaf(w0[t], v[t], a0); a1 = fmaf(w0[T + t], v[t], a1);
                a2 = fmaf(w0[2 * T + t], v[t], a2); a3 = fmaf(w0[3 * T + t], v[t], a3);
            }
        float* o = out + ((size_t)b * To + to) * D + d;
        o[0] = a0 + bias[to]; o[D] = a1 + bias[to + 1]; o[2 * (size_t)D] = a2 + bias[to + 2]; o[3 * (size_t)D] = a3 + bias[to + 3];
    }
    for (; to < To; ++to) {
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < TMAX; ++t) if (t < T) acc = fmaf(w[(size_t)to * T + t], v[t], acc);
        out[((size_t)b * To + to) * D + d] = acc + bias[to];
    }
}

// Fused multi-head attention: one CTA per (batch, head, chunk of QC query rows).  S <= 64, dh == 64.
// Round 1 ran one CTA per (batch, head) = 8 CTAs on 148 SMs with scalar shared-memory loops (33.5 us per call, 5 calls per
// line, 3 of them on the step's critical path); splitting the queries gives B*heads*S/QC CTAs, K / V are re-read from L2
// (32 KB per CTA), and the inner products read shared memory as float4 (K rows padded to 68 floats: conflict-free).
template <int QC>
__global__ void __launch_bounds__(128) attention_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                        int S, int heads, int nch, float scale) {
    mn_pdl_prologue();
    constexpr int DH = 64, SM = 64, KP = 68;
    extern __shared__ __align__(16) float att_smem[];
    float (*Q)[DH] = reinterpret_cast<float (*)[DH]>(att_smem);                       // [QC][64]
    float (*K)[KP] = reinterpret_cast<float (*)[KP]>(att_smem + QC * DH);              // [64][68]
    float (*V)[DH] = reinterpret_cast<float (*)[DH]>(att_smem + QC * DH + SM * KP);    // [64][64]
    float (*P)[KP] = reinterpret_cast<float (*)[KP]>(att_smem + QC * DH + SM * KP + SM * DH);   // [QC][68]
    const int chunk = blockIdx.x % nch, bh = blockIdx.x / nch;
    const int b = bh / heads, h = bh % heads;
    const int inner = heads * DH;
    const int q0 = chunk * QC;
    const int nq = min(QC, S - q0);
    const float* base = qkv + (size_t)b * S * 3 * inner + h * DH;
    const int tid = threadIdx.x;
    for (int idx = tid; idx < S * (DH / 4); idx += 128) {
        const int j = idx >> 4, d4 = (idx & 15) * 4;
        const float* r = base + (size_t)j * 3 * inner + d4;
        *reinterpret_cast<float4*>(&K[j][d4]) = *reinterpret_cast<const float4*>(r + inner);
        *reinterpret_cast<float4*>(&V[j][d4]) = *reinterpret_cast<const float4*>(r + 2 * inner);
    }
    for (int idx = tid; idx < nq * (DH / 4); idx += 128) {
        const int i = idx >> 4, d4 = (idx & 15) * 4;
        *reinterpret_cast<float4*>(&Q[i][d4]) = *reinterpret_cast<const float4*>(base + (size_t)(q0 + i) * 3 * inner + d4);
    }
    __syncthreads();
    // scores: thread -> key j = tid % 64, query rows i = tid / 64 + 2k
    {
        const int j = tid & 63, i0 = tid >> 6;
        if (j < S) {
            for (int i = i0; i < nq; i += 2) {
                float acc = 0.f;
#pragma unroll
                for (int d4 = 0; d4 < DH; d4 += 4) {
                    const float4 q = *reinterpret_cast<const float4*>(&Q[i][d4]);
                    const float4 k = *reinterpret_cast<const float4*>(&K[j][d4]);
                    acc = fmaf(q.x, k.x, acc); acc = fmaf(q.y, k.y, acc); acc = fmaf(q.z, k.z, acc); acc = fmaf(q.w, k.w, acc);
                }
                P[i][j] = acc * scale;
            }
        }
    }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5;
    for (int i = warp; i < nq; i += 4) {
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
    {
        const int d = tid & 63, i0 = tid >> 6;
        for (int i = i0; i < nq; i += 2) {
            float acc = 0.f;
            for (int j = 0; j < S; ++j) acc = fmaf(P[i][j], V[j][d], acc);
            out[((size_t)b * S + q0 + i) * inner + h * DH + d] = acc;
        }
    }
}

// Small-M linear layer (M <= 64 rows: the 64 / 16 tokens of one text line, the 16 characters of the mapping network):
//   y = act(x W + b + residual) * gain.
// One CTA = MT rows (16/32/64, all of M) x NT output columns (16 or 64) x one K slice; K is streamed in 32-wide chunks through
// a 4-stage cp.async ring.  These layers are pure latency / weight-streaming problems (a 512x512 layer is 1 MB of weights and
// 0.03 GFLOP), so the work is spread over ~one CTA per SM: grid.y = ksplit K-slices that form ONE thread-block cluster; the
// slices' partial sums are reduced through distributed shared memory in rank order (deterministic), and rank 0 runs the
// epilogue.  x may be a gathered matrix: element (r, k) lives at x + r*x_rs + (k / seg_len)*seg_stride + k % seg_len, which is
// how the TextViT patch embedding (Rearrange 'b c (h p1) (w p2) -> b h w (p1 p2 c)', textvit_arch.py:33-36) reads the NHWC
// feature map in place.  grid.z = independent batches (lines).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ uint32_t lin_cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void lin_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float lin_ld_peer(const float* my_smem_addr, uint32_t rank) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(my_smem_addr);
    uint32_t pa; float v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(pa) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(pa) : "memory");
    return v;
}

struct LinArgs {
    const float* x; long long x_rs, x_bs; int seg_len; long long seg_stride;
    const float* w; const float* bias; const float* residual; long long res_bs;
    float* y; int M, K, N, act; float gain;
    int ko;            // outer K slices (grid.z = batches * ko): slice q writes its RAW partial tile to y + q*batches*M*N, a second kernel reduces
};

template <int MT, int NT>
__global__ void __launch_bounds__(128) linear_small_m_kernel(const LinArgs a) {
    mn_pdl_prologue();
    constexpr int ST = 4, RPT = MT / 8, CPT = NT / 16, XLD = 36;
    extern __shared__ __align__(16) float lin_smem[];
    float (*Xs)[MT][XLD] = reinterpret_cast<float (*)[MT][XLD]>(lin_smem);
    float (*Ws)[32][NT] = reinterpret_cast<float (*)[32][NT]>(lin_smem + ST * MT * XLD);
    float (*red)[NT + 1] = reinterpret_cast<float (*)[NT + 1]>(lin_smem);          // reuses the ring after the K loop
    const int tid = threadIdx.x;
    const int cg = tid & 15, rg = tid >> 4;              // 16 column groups x 8 row groups
    const int n0 = blockIdx.x * NT;
    const int ksplit = gridDim.y;
    const uint32_t krank = ksplit > 1 ? lin_cluster_rank() : 0;
    const int batch = blockIdx.z / a.ko, kouter = blockIdx.z % a.ko;
    const float* xb = a.x + (size_t)batch * a.x_bs;
    const int nchunks_all = a.K / 32;
    const int kslices = ksplit * a.ko, kslice = kouter * ksplit + (int)krank;
    const int ch_begin = (int)((long long)nchunks_all * kslice / kslices), ch_end = (int)((long long)nchunks_all * (kslice + 1) / kslices);
    const int nchunks = ch_end - ch_begin;
    float acc[RPT][CPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[i][c] = 0.f;
    auto issue = [&](int ch) {
        if (ch < nchunks) {
            const int k0 = (ch_begin + ch) * 32, buf = ch % ST;
            const float* xk = xb + (size_t)(k0 / a.seg_len) * a.seg_stride + (k0 % a.seg_len);
#pragma unroll
            for (int i = 0; i < MT / 16; ++i) {          // X chunk: MT rows x 32 floats
                const int idx = tid + i * 128, r = idx >> 3, c4 = (idx & 7) * 4;
                cp_async16(&Xs[buf][r][c4], xk + (size_t)(r < a.M ? r : 0) * a.x_rs + c4, r < a.M);
            }
#pragma unroll
            for (int i = 0; i < NT / 16; ++i) {          // W chunk: 32 rows x NT floats
                const int idx = tid + i * 128, kr = idx / (NT / 4), c4 = (idx % (NT / 4)) * 4;
                cp_async16(&Ws[buf][kr][c4], a.w + (size_t)(k0 + kr) * a.N + (n0 + c4 < a.N ? n0 + c4 : 0), n0 + c4 < a.N);
            }
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
            float wv[4][CPT];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if constexpr (CPT == 4) {
                    const float4 t = *reinterpret_cast<const float4*>(&Ws[buf][k4 * 4 + kk][cg * 4]);
                    wv[kk][0] = t.x; wv[kk][1] = t.y; wv[kk][2] = t.z; wv[kk][3] = t.w;
                } else {
                    wv[kk][0] = Ws[buf][k4 * 4 + kk][cg];
                }
            }
            // k-outer, accumulator-inner: the RPT x CPT FMAs of one k are independent, so they issue back to back; the round-1 order
            // (four dependent FMAs per accumulator in a row) left a single warp per scheduler waiting on its own FMA latency --
            // ~1.7k cycles per 32-deep chunk for 256 FMAs per thread.  Same summation order per accumulator: bit-identical results.
            float4 xv[RPT];
#pragma unroll
            for (int i = 0; i < RPT; ++i) xv[i] = *reinterpret_cast<const float4*>(&Xs[buf][rg * RPT + i][k4 * 4]);
#pragma unroll
            for (int i = 0; i < RPT; ++i)
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(xv[i].x, wv[0][c], acc[i][c]);
#pragma unroll
            for (int i = 0; i < RPT; ++i)
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(xv[i].y, wv[1][c], acc[i][c]);
#pragma unroll
            for (int i = 0; i < RPT; ++i)
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(xv[i].z, wv[2][c], acc[i][c]);
#pragma unroll
            for (int i = 0; i < RPT; ++i)
#pragma unroll
                for (int c = 0; c < CPT; ++c) acc[i][c] = fmaf(xv[i].w, wv[3][c], acc[i][c]);
        }
        __syncthreads();
    }
    if (ksplit > 1) {
        // K slices of one cluster: park the partials in shared memory, rank 0 adds them in rank order through DSMEM
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int i = 0; i < RPT; ++i)
#pragma unroll
            for (int c = 0; c < CPT; ++c) red[rg * RPT + i][cg * CPT + c] = acc[i][c];
        lin_cluster_sync();
        if (krank == 0) {
            for (uint32_t q = 1; q < (uint32_t)ksplit; ++q)
#pragma unroll
                for (int i = 0; i < RPT; ++i)
#pragma unroll
                    for (int c = 0; c < CPT; ++c) acc[i][c] += lin_ld_peer(&red[rg * RPT + i][cg * CPT + c], q);
        }
        lin_cluster_sync();                              // peers keep their shared memory alive until rank 0 has read it
        if (krank != 0) return;
    }
    float* yb = a.y + ((size_t)kouter * (gridDim.z / a.ko) + batch) * a.M * a.N;
    const float* rb = a.residual ? a.residual + (size_t)batch * a.res_bs : nullptr;
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
        const int r = rg * RPT + i;
        if (r >= a.M) continue;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            const int o = n0 + cg * CPT + c;
            if (o >= a.N) continue;
            if (a.ko > 1) { yb[(size_t)r * a.N + o] = acc[i][c]; continue; }       // raw partial: linear_ko_reduce_kernel finishes
            float v = acc[i][c] + (a.bias ? a.bias[o] : 0.f);
            if (rb) v += rb[(size_t)r * a.N + o];
            yb[(size_t)r * a.N + o] = mn_apply_act(v, a.act) * a.gain;
        }
    }
}

template <int MT, int NT>
static cudaError_t launch_linear(const LinArgs& a, int batches, int ksplit, cudaStream_t st) {
    constexpr int ring = 4 * (MT * 36 + 32 * NT) * (int)sizeof(float), redb = MT * (NT + 1) * (int)sizeof(float);
    constexpr int smem = ring > redb ? ring : redb;
    static unsigned long long smem_done = 0;
    {
        cudaError_t e = mn_ensure_dyn_smem(linear_small_m_kernel<MT, NT>, smem, &smem_done);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(mn_cdiv(a.N, NT), ksplit, batches * a.ko);
    cfg.blockDim = dim3(128, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (ksplit > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1; attr[na].val.clusterDim.y = ksplit; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (mn_pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    return cudaLaunchKernelEx(&cfg, linear_small_m_kernel<MT, NT>, a);
}

// y[b][r][o] = act(sum_q part[q][b][r][o] + bias[o] + residual[b][r][o]) * gain   (q in slice order: deterministic)
__global__ void linear_ko_reduce_kernel(const float* __restrict__ part, int ko, long long slice_elems, const float* __restrict__ bias,
                                        const float* __restrict__ residual, long long res_bs, float* __restrict__ y, int M, int N,
                                        int act, float gain) {
    mn_pdl_prologue();
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= slice_elems) return;
    float4 v = *reinterpret_cast<const float4*>(part + i4);
    for (int q = 1; q < ko; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(part + (size_t)q * slice_elems + i4);
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    const int o = (int)(i4 % N);
    const long long row = i4 / N;
    const int b = (int)(row / M), r = (int)(row % M);
    if (bias) { v.x += bias[o]; v.y += bias[o + 1]; v.z += bias[o + 2]; v.w += bias[o + 3]; }
    if (residual) {
        const float4 t = *reinterpret_cast<const float4*>(residual + (size_t)b * res_bs + (size_t)r * N + o);
        v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    v.x = mn_apply_act(v.x, act) * gain; v.y = mn_apply_act(v.y, act) * gain; v.z = mn_apply_act(v.z, act) * gain; v.w = mn_apply_act(v.w, act) * gain;
    *reinterpret_cast<float4*>(y + i4) = v;
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

static int linear_small_m_impl(const float* x, long long x_row_stride, long long x_batch_stride, int x_seg_len, long long x_seg_stride,
                               const float* w, const float* bias, const float* residual, long long res_batch_stride, float* y,
                               int batches, int M, int K, int N, int act, float gain, float* ws, long long ws_bytes, void* stream) {
    MN_REQUIRE(x && w && y && M > 0 && M <= 64 && K > 0 && K % 32 == 0 && N > 0 && N % 16 == 0 && batches > 0 && batches <= 8192,
               "mn_linear_small_m: needs M<=64, K%32==0, N%16==0");
    MN_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0, "mn_linear_small_m: x, w must be 16-byte aligned");
    MN_REQUIRE(x_seg_len > 0 && x_seg_len % 32 == 0 && K % x_seg_len == 0 && (x_row_stride & 3) == 0 && (x_seg_stride & 3) == 0 && (x_batch_stride & 3) == 0,
               "mn_linear_small_m: gathered x needs 32-float segments and 16-byte aligned strides");
    LinArgs a{x, x_row_stride, x_batch_stride, x_seg_len, x_seg_stride, w, bias, residual, res_batch_stride, y, M, K, N, act, gain, 1};
    // tile: all rows x 16 columns (many small CTAs: these layers are latency bound and 64-row x 64-column tiles leave too few
    // warps per SM -- measured 39 vs 14 us for 64x512x1024); 64 columns for the <= 16-row layers with many columns
    // (the generator's 17 modulation FCs as one 16x512x7168 GEMM: 18 vs 30 us) and for DEEP-K layers (the TextViT patch embedding,
    // K = 32768: with 16-column tiles every one of the N/16 column tiles re-reads the whole 8.4 MB activation matrix -- 268 MB of
    // L2->SMEM traffic next to 67 MB of weights, measured 150 us = 0.44 TB/s; 64-column tiles read it N/64 times).
    const bool deep = K >= 8192 && N % 64 == 0;
    const bool wide = (M <= 16 && N >= 1024) || deep;
    const int tiles = mn_cdiv(N, wide ? 64 : 16) * batches;
    // K slices (one cluster, <= 8 CTAs): spread a small layer over ~one CTA per SM, a long K over ~two
    const int sms = mn_num_sms(), chunks = K / 32;
    // (in-graph sweep, tools/bench_linear.py, profiles/r2_linear_ks_sweep.txt: these kernels are latency bound, up to 12 CTAs fit
    // on an SM, and a 16-column tile prefers 4-chunk K slices: 64x512x1536 16.8 -> 13.1 us, 64x512x1024 11.0 -> 9.1, 64x1024x512 12.4 -> 10.8)
    const int budget = (K >= 4096 || !wide) ? 2 * sms : sms;
    int ks = 1;
    while (ks < 8 && tiles * ks * 2 <= budget && chunks / (ks * 2) >= 2) ks *= 2;
    static int force_ks = -1;                            // developer override: MN_LIN_KS=1|2|4|8
    if (force_ks < 0) { const char* e = getenv("MN_LIN_KS"); force_ks = e ? atoi(e) : 0; }
    if (force_ks > 0) { ks = force_ks; while (ks > 1 && chunks / ks < 1) ks /= 2; }
    // outer K slices beyond the cluster limit (deep K, few tiles): raw partial tiles go to the caller's workspace, a second
    // kernel adds them in slice order and runs the epilogue (deterministic; no atomics)
    int ko = 1;
    if (deep && ws) {
        while (ko < 8 && tiles * ks * ko * 2 <= 4 * sms && chunks / (ks * ko * 2) >= 8 &&
               (long long)(ko * 2) * batches * M * N * 4 <= ws_bytes) ko *= 2;
    }
    static int force_ko = -1;
    if (force_ko < 0) { const char* e = getenv("MN_LIN_KO"); force_ko = e ? atoi(e) : 0; }
    if (force_ko > 0 && ws && (long long)force_ko * batches * M * N * 4 <= ws_bytes && chunks / (ks * force_ko) >= 1) ko = force_ko;
    a.ko = ko;
    if (ko > 1) a.y = ws;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e;
    const int mt = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
    if (wide) e = mt == 16 ? launch_linear<16, 64>(a, batches, ks, st) : (mt == 32 ? launch_linear<32, 64>(a, batches, ks, st) : launch_linear<64, 64>(a, batches, ks, st));
    else e = mt == 16 ? launch_linear<16, 16>(a, batches, ks, st) : (mt == 32 ? launch_linear<32, 16>(a, batches, ks, st) : launch_linear<64, 16>(a, batches, ks, st));
    MN_CUDA_CHECK(e);
    MN_LAUNCH_CHECK();
    if (ko > 1) {
        const long long slice = (long long)batches * M * N;
        MN_CUDA_CHECK((mn_launch(linear_ko_reduce_kernel, dim3((unsigned)mn_cdiv64(slice / 4, 256)), dim3(256), 0, st, (const float*)ws, ko, slice, bias,
                                 residual, res_batch_stride, y, M, N, act, gain)));
        MN_LAUNCH_CHECK();
    }
    return MN_OK;
}

extern "C" int mn_linear_small_m_ex(const float* x, long long x_row_stride, long long x_batch_stride, int x_seg_len, long long x_seg_stride,
                                    const float* w, const float* bias, const float* residual, long long res_batch_stride, float* y,
                                    int batches, int M, int K, int N, int act, float gain, void* stream) {
    return linear_small_m_impl(x, x_row_stride, x_batch_stride, x_seg_len, x_seg_stride, w, bias, residual, res_batch_stride, y, batches, M, K, N,
                               act, gain, nullptr, 0, stream);
}

extern "C" int mn_linear_small_m_ws(const float* x, long long x_row_stride, long long x_batch_stride, int x_seg_len, long long x_seg_stride,
                                    const float* w, const float* bias, const float* residual, long long res_batch_stride, float* y,
                                    int batches, int M, int K, int N, int act, float gain, float* workspace, long long workspace_bytes,
                                    void* stream) {
    MN_REQUIRE(!workspace || (((uintptr_t)workspace & 15) == 0 && N % 4 == 0), "mn_linear_small_m_ws: workspace alignment");
    return linear_small_m_impl(x, x_row_stride, x_batch_stride, x_seg_len, x_seg_stride, w, bias, residual, res_batch_stride, y, batches, M, K, N,
                               act, gain, workspace, workspace_bytes, stream);
}

extern "C" int mn_linear_small_m(const float* x, const float* w, const float* bias, const float* residual, float* y,
                                 int M, int K, int N, int act, float gain, void* stream) {
    return mn_linear_small_m_ex(x, K, 0, K, 0, w, bias, residual, 0, y, 1, M, K, N, act, gain, stream);
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
    MN_REQUIRE(((uintptr_t)qkv & 15) == 0, "mn_attention: qkv must be 16-byte aligned");
    constexpr int QC = 8;
    constexpr int kSmem = (QC * 64 + 64 * 68 + 64 * 64 + QC * 68) * (int)sizeof(float);
    static unsigned long long smem_done = 0;
    MN_CUDA_CHECK(mn_ensure_dyn_smem(attention_kernel<QC>, kSmem, &smem_done));
    const int nch = mn_cdiv(S, QC);
    MN_CUDA_CHECK((mn_launch(attention_kernel<QC>, dim3(B * heads * nch), dim3(128), kSmem, (cudaStream_t)stream, qkv, out, S, heads, nch, scale)));
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
