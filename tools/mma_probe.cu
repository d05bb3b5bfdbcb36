// Micro-probe: issue rate of tcgen05.mma.kind::f16 (M=128) under different operand / accumulator patterns.
// Operands are uninitialised shared memory / TMEM (values irrelevant), one CTA per SM, one issuing thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../marconet_b200/csrc/tc_ptx.cuh"
using namespace tcptx;

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

// mode bit0: A from TMEM (TS) else SS ; nacc: number of distinct accumulators cycled through
template <int N>
__global__ void __launch_bounds__(128, 1) probe(int mode, int nacc, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (warp == 1) {
        const uint64_t da = make_b_desc(smem_u32(smem));
        const uint64_t db = make_b_desc(smem_u32(smem) + 32768);
        long long t0 = 0, t1 = 0;
        if (elect_one_sync()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t d = tm + ((i * 4 + j) % nacc) * N;
                    if (mode & 1) tc_mma_ts(d, tm + 448 + j * 8, db + 2 * j, idesc, 1);
                    else mma_ss(d, da + 2 * j, db + 2 * j, idesc, 1);
                }
            }
            tc_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        t1 = clock64();
        if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int N>
void run(const char* name, int mode, int nacc, long long* dout) {
    const int iters = 2000;
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    for (int rep = 0; rep < 2; ++rep) probe<N><<<148, 128, 100 * 1024>>>(mode, nacc, iters, dout);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost);
    printf("%-34s N=%3d nacc=%d : %8.1f cycles / MMA  (ideal %d)  %s\n", name, N, nacc, (double)cyc / (iters * 4.0), N / 2,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* dout;
    cudaMalloc(&dout, 8);
    run<256>("SS same accumulator", 0, 1, dout);
    run<128>("SS same accumulator", 0, 1, dout);
    run<128>("SS 2 accumulators alternating", 0, 2, dout);
    run<128>("TS same accumulator", 1, 1, dout);
    run<128>("TS 2 accumulators alternating", 1, 2, dout);
    run<128>("TS 3 accumulators", 1, 3, dout);
    run<64>("TS same accumulator", 1, 1, dout);
    run<64>("TS 4 accumulators", 1, 4, dout);
    run<256>("TS same accumulator", 1, 1, dout);
    return 0;
}
