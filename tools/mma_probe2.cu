// Micro-probe 2: issue rate of tcgen05.mma.kind::f16 with cta_group::2 (M = 256 over a CTA pair) next to cta_group::1, for the
// operand patterns conv_tc2.cu uses (A from TMEM, two accumulators D / Dc, 3 MMAs per k-step) -- decides whether the 2-CTA form is
// worth building.  Operands are uninitialised shared memory / TMEM (values irrelevant); 74 clusters of 2 CTAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe2 tools/mma_probe2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../marconet_b200/csrc/tc_ptx.cuh"
using namespace tcptx;

__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma2_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// pattern 0: one accumulator; 1: the conv_tc2 pattern (D, Dc, Dc per k-step)
template <int N, int CG>
__global__ void __launch_bounds__(128, 1) probe(int ts, int pattern, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    const uint32_t crank = CG == 2 ? cluster_ctarank() : 0u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) { if (CG == 2) tmem_alloc2(smem_u32(&slot), 512); else tmem_alloc(smem_u32(&slot), 512); }
    tc_fence_before(); __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
    if (warp == 1 && crank == 0) {
        const uint64_t da = make_b_desc(smem_u32(smem));
        const uint64_t db = make_b_desc(smem_u32(smem) + 32768);
        long long t0 = 0, t1 = 0;
        if (elect_one_sync()) {
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t a = tm + 2 * N + j * 8;
                    if (pattern == 0) {
                        if (CG == 2) { if (ts) mma2_ts(tm, a, db + 2 * j, idesc, 1); else mma2_ss(tm, da + 2 * j, db + 2 * j, idesc, 1); }
                        else tc_mma_ts(tm, a, db + 2 * j, idesc, 1);
                    } else {
                        if (CG == 2) {
                            mma2_ts(tm, a, db + 2 * j, idesc, 1);
                            mma2_ts(tm + N, a, db + 1024 + 2 * j, idesc, 1);
                            mma2_ts(tm + N, a + 32, db + 2 * j, idesc, 1);
                        } else {
                            tc_mma_ts(tm, a, db + 2 * j, idesc, 1);
                            tc_mma_ts(tm + N, a, db + 1024 + 2 * j, idesc, 1);
                            tc_mma_ts(tm + N, a + 32, db + 2 * j, idesc, 1);
                        }
                    }
                }
            }
            if (CG == 2) commit2(smem_u32(&bar)); else tc_commit(smem_u32(&bar));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bar), 0);
        t1 = clock64();
        if (threadIdx.x == 32 && blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) { tc_fence_after(); if (CG == 2) tmem_dealloc2(tm, 512); else tmem_dealloc(tm, 512); }
}

template <int N, int CG>
void run(const char* name, int ts, int pattern, long long* dout) {
    const int iters = 1000;
    cudaFuncSetAttribute(probe<N, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaError_t e = cudaSuccess;
    for (int rep = 0; rep < 2; ++rep) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 100 * 1024;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, probe<N, CG>, ts, pattern, iters, dout);
        if (e != cudaSuccess) break;
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, dout, 8, cudaMemcpyDeviceToHost);
    const double per = (double)cyc / (iters * 4.0 * (pattern ? 3.0 : 1.0));
    printf("%-44s cta_group::%d N=%3d : %7.1f cycles / MMA (ideal %3d) = %5.1f %% of the tensor pipe  %s\n", name, CG, N, per, N / 2,
           100.0 * (N / 2) / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* dout;
    cudaMalloc(&dout, 8);
    run<128, 1>("TS one accumulator", 1, 0, dout);
    run<128, 1>("TS conv_tc2 pattern (D,Dc,Dc)", 1, 1, dout);
    run<128, 2>("TS one accumulator", 1, 0, dout);
    run<128, 2>("TS conv_tc2 pattern (D,Dc,Dc)", 1, 1, dout);
    run<128, 2>("SS one accumulator", 0, 0, dout);
    run<256, 2>("SS one accumulator", 0, 0, dout);
    run<64, 1>("TS conv_tc2 pattern (D,Dc,Dc)", 1, 1, dout);
    run<64, 2>("TS conv_tc2 pattern (D,Dc,Dc)", 1, 1, dout);
    return 0;
}
