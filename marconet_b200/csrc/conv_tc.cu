// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a -- the tensor-core path of mn_conv2d_nhwc.
//
// Problem: the reference computes every conv in fp32 and the parity budget (1e-3 end to end, bit-exact
// argmax) rules out single-pass fp16/bf16/tf32 operands.  This kernel keeps fp32
// activations in HBM and gets fp32-grade products out of the fp16 tensor pipe by operand splitting:
//     x = xh + xl,  w = wh + wl      (xh, xl, wh, wl fp16; |x - xh - xl| <= 2^-22 |x|)
//     x*w ~= xh*wh + xh*wl + xl*wh   (three kind::f16 MMAs accumulating in one fp32 TMEM accumulator)
//
// Data path per CTA (one 128-pixel x NT-channel output tile, K loop over taps x 64-channel blocks):
//   TMA   : 4-D tensor map over the NHWC fp32 activation; a box of 128 pixels x 32 channels per issue,
//           shifted by the tap offset, out-of-bounds rows/cols zero-filled by the hardware (= conv padding,
//           no im2col, no halo code), 128B-swizzled in shared memory.  Weights: 3-D map over pre-split
//           fp16 [tap][Cout][Cin] (K-major), 128B swizzle = the canonical UMMA B layout.
//   split : 4 converter warps read the fp32 tile from shared memory (conflict-free through the swizzle),
//           split each value into (hi, lo) fp16 and store them with tcgen05.st into TENSOR MEMORY as the
//           A operand (row = TMEM lane, two k-elements per 32-bit column).
//   MMA   : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 with A from TMEM and B from
//           shared memory, D (128 x NT fp32) in TMEM; tcgen05.commit frees the stage.
//   epi   : the converter warps read D with tcgen05.ld and run the shared fused epilogue (demod scale,
//           bias, residual, activation, window mask, second pre-modulated output).
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = split + epilogue.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "conv_common.cuh"
#include "mn_common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int TILE_M = 128;
constexpr int KB = 64;                 // channels per k-block
constexpr int A_BOX_BYTES = 128 * 128; // 128 pixels x 32 fp32 channels
constexpr int A_BYTES = 2 * A_BOX_BYTES;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 512;

struct TcGeom {
    int TW, TH, TN;                // pixel tile: TW*TH*TN == 128
    int tiles_w, tiles_h, tiles_n;
    int cblocks, taps;
    const float* wscale;           // device scalar: 2^-S undoing the power-of-two weight pre-scale
    int prec;
};

using namespace tcptx;

// ------------------------------------------------------------------------------------ the kernel
template <int NT, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmBlo, const ConvGeom g, const TcGeom t) {
    constexpr int B_BYTES = NT * 128;
    constexpr int STAGE_BYTES = A_BYTES + 2 * B_BYTES;
    // TMEM columns: D (hi*hi products) in [0,NT), Dc (hi*lo + lo*hi correction products) in [NT,2NT),
    // A(stage s) hi at 2NT+64s, lo at 2NT+64s+32.  The tensor core TRUNCATES when it adds into the fp32
    // accumulator, so the error of a long K loop is a bias ~ (#accumulations) * ulp(|D|)/2.  Keeping the two
    // small cross terms in their own accumulator (2^-11 the magnitude) leaves only K/16 truncations at full
    // magnitude instead of 3K/16 and the final D + Dc is one round-to-nearest fp32 add in the epilogue.
    constexpr int A_COL0 = 2 * NT;
    static_assert(2 * NT + STAGES * 64 <= TMEM_COLS, "TMEM budget");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 1);
    const uint32_t smem_base = smem_u32(smem);
    auto bar_full = [&](int s) { return smem_u32(bars + s); };
    auto bar_conv = [&](int s) { return smem_u32(bars + STAGES + s); };
    auto bar_empty = [&](int s) { return smem_u32(bars + 2 * STAGES + s); };
    const uint32_t bar_acc = smem_u32(bars + 3 * STAGES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- tile coordinates ----
    int tile = blockIdx.x;
    const int tw_i = tile % t.tiles_w; tile /= t.tiles_w;
    const int th_i = tile % t.tiles_h; tile /= t.tiles_h;
    const int n0 = tile * t.TN, oy0 = th_i * t.TH, ox0 = tw_i * t.TW;
    const int nt_i = blockIdx.y;
    const int num_kb = t.taps * t.cblocks;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBhi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_conv(s), 128);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(bar_empty(s), ph ^ 1);
                mbar_expect_tx(bar_full(s), STAGE_BYTES);
                const int tap = kb / t.cblocks, cb = kb - tap * t.cblocks;
                const int ky = tap / g.KW, kx = tap - ky * g.KW;
                const uint32_t a_dst = smem_base + s * STAGE_BYTES;
                tma_load_4d(&tmA, bar_full(s), a_dst, cb * KB, ox0 + kx - g.pw, oy0 + ky - g.ph, n0);
                tma_load_4d(&tmA, bar_full(s), a_dst + A_BOX_BYTES, cb * KB + 32, ox0 + kx - g.pw, oy0 + ky - g.ph, n0);
                tma_load_3d(&tmBhi, bar_full(s), a_dst + A_BYTES, cb * KB, nt_i * NT, tap);
                tma_load_3d(&tmBlo, bar_full(s), a_dst + A_BYTES + B_BYTES, cb * KB, nt_i * NT, tap);
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        if (lane == 0) {
            const uint32_t fmt = (t.prec == MN_PREC_BF16X3_TC) ? 1u : 0u;
            const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
            const bool three = t.prec != MN_PREC_F16X1_TC;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                mbar_wait(bar_conv(s), ph);
                mbar_wait(bar_full(s), ph);
                tc_fence_after();
                const uint32_t b_hi = smem_base + s * STAGE_BYTES + A_BYTES;
                const uint32_t b_lo = b_hi + B_BYTES;
                const uint32_t a_hi = tmem_base + A_COL0 + s * 64;
#pragma unroll
                for (int j = 0; j < KB / 16; ++j) {
                    const uint64_t dh = make_b_desc(b_hi + j * 32);
                    tc_mma_ts(tmem_base, a_hi + j * 8, dh, idesc, (kb | j) != 0);
                    if (three) {
                        const uint64_t dl = make_b_desc(b_lo + j * 32);
                        tc_mma_ts(tmem_base + NT, a_hi + j * 8, dl, idesc, (kb | j) != 0);
                        tc_mma_ts(tmem_base + NT, a_hi + 32 + j * 8, dh, idesc, 1);
                    }
                }
                tc_commit(bar_empty(s));
            }
            tc_commit(bar_acc);
        }
    } else {
        // =========================== split (fp32 -> hi/lo fp16 in TMEM) + epilogue ===========================
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;            // tile row == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool bf = t.prec == MN_PREC_BF16X3_TC;
        const uint32_t mask = bf ? 0xFFFF0000u : 0xFFFFE000u;
        const float xs = g.x_scale;
        float amax = 0.f;                      // range guard, see conv_common.cuh
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(bar_full(s), ph);
            const uint8_t* a_src = smem + s * STAGE_BYTES + r * 128;
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int box = 0; box < 2; ++box) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 v = *reinterpret_cast<const float4*>(a_src + box * A_BOX_BYTES + ((j ^ (r & 7)) << 4));
                    v.x *= xs; v.y *= xs; v.z *= xs; v.w *= xs;
                    amax = fmaxf(fmaxf(amax, fabsf(v.x)), fmaxf(fabsf(v.y), fmaxf(fabsf(v.z), fabsf(v.w))));
                    const float h0 = __uint_as_float(__float_as_uint(v.x) & mask), h1 = __uint_as_float(__float_as_uint(v.y) & mask);
                    const float h2 = __uint_as_float(__float_as_uint(v.z) & mask), h3 = __uint_as_float(__float_as_uint(v.w) & mask);
                    const int c = box * 16 + j * 2;
                    if (bf) {
                        hi[c] = pack_bf16(h0, h1); hi[c + 1] = pack_bf16(h2, h3);
                        lo[c] = pack_bf16(v.x - h0, v.y - h1); lo[c + 1] = pack_bf16(v.z - h2, v.w - h3);
                    } else {
                        hi[c] = pack_f16(h0, h1); hi[c + 1] = pack_f16(h2, h3);
                        lo[c] = pack_f16(v.x - h0, v.y - h1); lo[c + 1] = pack_f16(v.z - h2, v.w - h3);
                    }
                }
            }
            const uint32_t a_dst = lane_addr + A_COL0 + s * 64;
#pragma unroll
            for (int c = 0; c < 4; ++c) tc_st8(a_dst + c * 8, hi + c * 8);
#pragma unroll
            for (int c = 0; c < 4; ++c) tc_st8(a_dst + 32 + c * 8, lo + c * 8);
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(bar_conv(s));
        }

        conv_range_report(g, __float_as_uint(amax), t.prec == MN_PREC_F16X3_TC || t.prec == MN_PREC_F16X1_TC);
        // ---- epilogue ----
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        const int tn = r / (t.TH * t.TW);
        const int rem = r - tn * (t.TH * t.TW);
        const int th = rem / t.TW, tw = rem - th * t.TW;
        const int n = n0 + tn, oy = oy0 + th, ox = ox0 + tw;
        const bool row_ok = n < g.N && oy < g.OH && ox < g.OW;
        const int m = (n * g.OH + oy) * g.OW + ox;
        const float wscale = (t.wscale ? *t.wscale : 1.f) / g.x_scale;
        const float dfix = 1.f + 1.5e-8f * (float)(num_kb * (KB / 16));   // expected truncation shrink of the main accumulator, see conv_tc2.cu
#pragma unroll 1
        for (int chunk = 0; chunk < NT / 16; ++chunk) {
            uint32_t acc[16];
            tc_ld16(lane_addr + chunk * 16, acc);
            if (t.prec != MN_PREC_F16X1_TC) {
                uint32_t cor[16];
                tc_ld16(lane_addr + NT + chunk * 16, cor);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = __float_as_uint(fmaf(__uint_as_float(acc[i]), dfix, __uint_as_float(cor[i])));
            } else {
                tc_wait_ld();
            }
            if (row_ok) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    const int o = nt_i * NT + chunk * 16 + k4 * 4;
                    if (o < g.Cout) {
                        float v[4] = {__uint_as_float(acc[k4 * 4]) * wscale, __uint_as_float(acc[k4 * 4 + 1]) * wscale,
                                      __uint_as_float(acc[k4 * 4 + 2]) * wscale, __uint_as_float(acc[k4 * 4 + 3]) * wscale};
                        conv_epilogue4(g, m, o, v);
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------ weight packing
__global__ void absmax_kernel(const float* __restrict__ w, int64_t n, float* __restrict__ out) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(w[i]));
    m = mn_warp_max(m);
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out), __float_as_int(m));   // m >= 0: int order == float order
}

// w: [taps*Cin][Cout] fp32 (the SIMT layout) -> hi/lo 16-bit [taps][Cout][Cin] scaled by 2^S; scale[0] = absmax in,
// scale[1] = 2^-S out.
__global__ void pack_tc_kernel(const float* __restrict__ w, int taps, int Cin, int Cout, int bf, uint16_t* __restrict__ hi,
                               uint16_t* __restrict__ lo, float* __restrict__ scale) {
    const float amax = scale[0];
    int e = 0;
    if (amax > 0.f) { frexpf(amax, &e); }                 // amax = f * 2^e, f in [0.5,1)
    const int S = bf ? 0 : (14 - e);                      // |w| * 2^S < 2^14
    const float up = ldexpf(1.f, S);
    if (blockIdx.x == 0 && threadIdx.x == 0) scale[1] = ldexpf(1.f, -S);
    const int64_t total = (int64_t)taps * Cin * Cout;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % Cin);
        const int64_t rest = i / Cin;
        const int o = (int)(rest % Cout);
        const int tap = (int)(rest / Cout);
        const float v = w[((size_t)tap * Cin + c) * Cout + o] * up;
        if (bf) {
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
            hi[i] = *reinterpret_cast<const uint16_t*>(&h); lo[i] = *reinterpret_cast<const uint16_t*>(&l);
        } else {
            const __half h = __float2half_rn(v);
            const __half l = __float2half_rn(v - __half2float(h));
            hi[i] = *reinterpret_cast<const uint16_t*>(&h); lo[i] = *reinterpret_cast<const uint16_t*>(&l);
        }
    }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

struct TcPlan {
    bool ok; const char* why;
    int NT, STAGES;
    TcGeom t;
};

TcPlan plan_tc(const ConvGeom& g) {
    TcPlan p{};
    p.ok = false;
    auto fail = [&](const char* w) { p.why = w; return p; };
    if (g.sh != 1 || g.sw != 1) return fail("stride != 1");
    if (!((g.KH == 3 && g.KW == 3 && g.ph == 1 && g.pw == 1) || (g.KH == 1 && g.KW == 1 && g.ph == 0 && g.pw == 0)))
        return fail("only 3x3/pad1 and 1x1/pad0");
    if (g.Cin % KB != 0) return fail("Cin % 64 != 0");
    if (g.Cout % 64 != 0) return fail("Cout % 64 != 0");
    if (g.x_cs % 4 != 0 || (reinterpret_cast<uintptr_t>(g.x) & 15)) return fail("x alignment");
    // pixel tile
    int TW, TH, TN;
    if (g.W >= 128) { if (g.W % 128) return fail("W % 128"); TW = 128; TH = 1; TN = 1; }
    else {
        if (128 % g.W) return fail("W does not divide 128");
        TW = g.W;
        const int rows = 128 / TW;
        if (g.H >= rows) { if (g.H % rows) return fail("H % tile rows"); TH = rows; TN = 1; }
        else { if (rows % g.H) return fail("H does not divide tile rows"); TH = g.H; TN = rows / g.H; }
    }
    p.t.TW = TW; p.t.TH = TH; p.t.TN = TN;
    p.t.tiles_w = g.W / TW; p.t.tiles_h = g.H / TH; p.t.tiles_n = (g.N + TN - 1) / TN;
    p.t.cblocks = g.Cin / KB; p.t.taps = g.KH * g.KW;
    if (g.Cout % 128 == 0) { p.NT = 128; p.STAGES = 3; }
    else { p.NT = 64; p.STAGES = 4; }
    p.ok = true;
    return p;
}

template <int NT, int STAGES>
int launch_tc(const CUtensorMap& ma, const CUtensorMap& mbh, const CUtensorMap& mbl, const ConvGeom& g, const TcGeom& t, cudaStream_t st) {
    constexpr int SMEM = STAGES * (A_BYTES + 2 * NT * 128) + 1024 + 256;
    static unsigned long long smem_done = 0;
    MN_CUDA_CHECK(mn_ensure_dyn_smem(conv_tc_kernel<NT, STAGES>, SMEM, &smem_done));
    dim3 grid(t.tiles_w * t.tiles_h * t.tiles_n, g.Cout / NT);
    conv_tc_kernel<NT, STAGES><<<grid, NUM_THREADS, SMEM, st>>>(ma, mbh, mbl, g, t);
    MN_LAUNCH_CHECK();
    return MN_OK;
}

}  // namespace

int mn_conv_tc_supported(const ConvGeom& g, const char** why) {
    TcPlan p = plan_tc(g);
    if (why) *why = p.ok ? "" : p.why;
    return p.ok ? 1 : 0;
}

int mn_conv_tc_launch(const ConvGeom& g, const void* w_hi, const void* w_lo, const float* w_scale, int prec, cudaStream_t st) {
    TcPlan p = plan_tc(g);
    if (!p.ok) { mn_set_error("mn_conv2d_nhwc: tensor-core path does not support this shape (%s)", p.why); return MN_ERR_UNSUPPORTED; }
    if (!w_hi || !w_lo || !w_scale) { mn_set_error("mn_conv2d_nhwc: tensor-core precision needs packed w_tc_hi/w_tc_lo/w_tc_scale"); return MN_ERR_INVALID; }
    PFN_encodeTiled enc = get_encode();
    if (!enc) { mn_set_error("cuTensorMapEncodeTiled not available from the driver"); return MN_ERR_CUDA; }
    CUtensorMap ma, mbh, mbl;
    {
        cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N};
        cuuint64_t strides[3] = {(cuuint64_t)g.x_cs * 4, (cuuint64_t)g.W * g.x_cs * 4, (cuuint64_t)g.H * g.W * g.x_cs * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)p.t.TW, (cuuint32_t)p.t.TH, (cuuint32_t)p.t.TN};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&ma, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(g.x), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { mn_set_error("cuTensorMapEncodeTiled(A) failed: %d", (int)r); return MN_ERR_CUDA; }
    }
    const CUtensorMapDataType dt = (prec == MN_PREC_BF16X3_TC) ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    for (int which = 0; which < 2; ++which) {
        cuuint64_t dims[3] = {(cuuint64_t)g.Cin, (cuuint64_t)g.Cout, (cuuint64_t)(g.KH * g.KW)};
        cuuint64_t strides[2] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.Cin * g.Cout * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)p.NT, 1};
        cuuint32_t es[3] = {1, 1, 1};
        CUresult r = enc(which ? &mbl : &mbh, dt, 3, const_cast<void*>(which ? w_lo : w_hi), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { mn_set_error("cuTensorMapEncodeTiled(B) failed: %d", (int)r); return MN_ERR_CUDA; }
    }
    TcGeom t = p.t;
    t.wscale = w_scale + 1;
    t.prec = prec;
    if (p.NT == 128) return launch_tc<128, 3>(ma, mbh, mbl, g, t, st);
    return launch_tc<64, 4>(ma, mbh, mbl, g, t, st);
}

extern "C" int mn_conv_pack_weights_tc(const float* w, int taps, int Cin, int Cout, int precision, void* hi, void* lo,
                                       float* scale2, void* stream) {
    MN_REQUIRE(w && hi && lo && scale2 && taps > 0 && Cin > 0 && Cout > 0, "mn_conv_pack_weights_tc: bad args");
    MN_REQUIRE(precision == MN_PREC_F16X3_TC || precision == MN_PREC_BF16X3_TC || precision == MN_PREC_F16X1_TC,
               "mn_conv_pack_weights_tc: precision must be a tensor-core mode");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = (int64_t)taps * Cin * Cout;
    MN_CUDA_CHECK(cudaMemsetAsync(scale2, 0, 2 * sizeof(float), st));
    const int blocks = (int)(mn_cdiv64(total, 256) < 1184 ? mn_cdiv64(total, 256) : 1184);
    absmax_kernel<<<blocks, 256, 0, st>>>(w, total, scale2);
    MN_LAUNCH_CHECK();
    pack_tc_kernel<<<blocks, 256, 0, st>>>(w, taps, Cin, Cout, precision == MN_PREC_BF16X3_TC ? 1 : 0,
                                           reinterpret_cast<uint16_t*>(hi), reinterpret_cast<uint16_t*>(lo), scale2);
    MN_LAUNCH_CHECK();
    return MN_OK;
}
