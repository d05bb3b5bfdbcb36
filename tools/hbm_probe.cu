// Micro-probe: achievable HBM bandwidth of the access patterns the HBM-bound operators use (write-only fill, copy, 1:4
// read:write like bilinear x2, read-only reduction), for "one 128-bit item per thread, huge grid" vs "grid-stride, U items
// per thread" shapes.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/hbm_probe tools/hbm_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void fill1(float4* y, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}
template <int U>
__global__ void fillU(float4* y, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * U) {
#pragma unroll
        for (int k = 0; k < U; ++k) if (i + k * stride < n) y[i + k * stride] = make_float4(1.f, 2.f, 3.f, 4.f);
    }
}
__global__ void copy1(const float4* x, float4* y, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = x[i];
}
template <int U>
__global__ void copyU(const float4* __restrict__ x, float4* __restrict__ y, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * U) {
        float4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = (i + k * stride < n) ? x[i + k * stride] : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < U; ++k) if (i + k * stride < n) y[i + k * stride] = v[k];
    }
}
// 1 read : 4 writes (each input item is written to 4 places, like an upsample without the arithmetic)
__global__ void expand4(const float4* __restrict__ x, float4* __restrict__ y, size_t n_in) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    const float4 v = x[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) y[i + k * n_in] = v;
}
__global__ void expand4_st_cs(const float4* __restrict__ x, float4* __restrict__ y, size_t n_in) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    const float4 v = x[i];
#pragma unroll
    for (int k = 0; k < 4; ++k) __stcs(&y[i + k * n_in], v);
}
template <int U>
__global__ void readU(const float4* __restrict__ x, float* out, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * U) {
        float4 v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = (i + k * stride < n) ? x[i + k * stride] : make_float4(0, 0, 0, 0);
#pragma unroll
        for (int k = 0; k < U; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
    }
    if (acc == 123.456f) *out = acc;
}

template <typename F>
static void run(const char* name, double bytes, F launch, float4* flush, size_t flush_n) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double tot = 0; const int iters = 10;
    for (int i = 0; i < iters + 2; ++i) {
        cudaMemsetAsync(flush, 0, flush_n * 16);       // leaves the L2 full of dirty lines, like a producer kernel would
        cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (i >= 2) tot += ms;
    }
    cudaError_t err = cudaGetLastError();
    printf("%-34s %8.1f us  %7.0f GB/s  %s\n", name, tot / iters * 1e3, bytes / (tot / iters * 1e-3) / 1e9, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    const size_t n_out = (size_t)16 * 128 * 128 * 256 / 4;       // float4 items of the largest bilinear x2 output (268 MB)
    const size_t n_in = n_out / 4;
    float4 *x, *y, *flush; float* o;
    cudaMalloc(&x, n_out * 16); cudaMalloc(&y, n_out * 16); cudaMalloc(&flush, (size_t)256 << 20); cudaMalloc(&o, 4);
    cudaMemset(x, 0, n_out * 16);
    const size_t fn = ((size_t)256 << 20) / 16;
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double ob = (double)n_out * 16, ib = (double)n_in * 16;
    run("fill 1/thread", ob, [&] { fill1<<<(unsigned)((n_out + 255) / 256), 256>>>(y, n_out); }, flush, fn);
    run("fill grid-stride U=4", ob, [&] { fillU<4><<<sms * 8, 256>>>(y, n_out); }, flush, fn);
    run("fill grid-stride U=8 x16 blocks", ob, [&] { fillU<8><<<sms * 16, 256>>>(y, n_out); }, flush, fn);
    run("copy 1/thread", 2 * ob, [&] { copy1<<<(unsigned)((n_out + 255) / 256), 256>>>(x, y, n_out); }, flush, fn);
    run("copy grid-stride U=4", 2 * ob, [&] { copyU<4><<<sms * 8, 256>>>(x, y, n_out); }, flush, fn);
    run("copy grid-stride U=8", 2 * ob, [&] { copyU<8><<<sms * 8, 256>>>(x, y, n_out); }, flush, fn);
    run("expand 1 read : 4 writes", ob + ib, [&] { expand4<<<(unsigned)((n_in + 255) / 256), 256>>>(x, y, n_in); }, flush, fn);
    run("expand 1:4, st.cs", ob + ib, [&] { expand4_st_cs<<<(unsigned)((n_in + 255) / 256), 256>>>(x, y, n_in); }, flush, fn);
    run("read grid-stride U=4", ob, [&] { readU<4><<<sms * 8, 256>>>(x, o, n_out); }, flush, fn);
    run("read grid-stride U=8 x16 blocks", ob, [&] { readU<8><<<sms * 16, 256>>>(x, o, n_out); }, flush, fn);
    run("cudaMemcpyAsync D2D", 2 * ob, [&] { cudaMemcpyAsync(y, x, n_out * 16, cudaMemcpyDeviceToDevice); }, flush, fn);
    run("cudaMemsetAsync", ob, [&] { cudaMemsetAsync(y, 0, n_out * 16); }, flush, fn);
    return 0;
}
