// Host pre/post-processing of the reference scripts moved onto the device:
//   test_sr.py:98-111   cv2.resize(img, (0,0), fx=32/h, fy=32/h, INTER_CUBIC) -> zero-pad to 32x512 -> ToTensor -> Normalize(.5,.5)
//   test_sr.py:198-201  sr*0.5+0.5 -> HWC -> channel flip -> clip(.,0,1)*255 -> (cv2.imwrite :231) round to uint8
// Byte/integer work: results are bit-identical to OpenCV's own 8-bit cubic resize (imgproc/src/resize.cpp: fp32
// interpolateCubic with A=-0.75, taps rounded to 11-bit fixed point, integer horizontal pass with replicated borders, vertical
// pass in fp32 -- separate multiply and add, rows 3..0, round-half-even -- for the first floor(W*cn/8)*8 elements of a row (the
// baseline-SSE vector body) and in fixed point for the tail) and to torchvision's ToTensor/Normalize arithmetic.
// HBM-bound and tiny (one 32x512 line); every float operation uses an explicit-rounding intrinsic so that nvcc cannot contract
// a multiply-add into an FMA the CPU code does not have.
#include "mn_common.cuh"

namespace {

struct CubicTaps { int ofs; int t[4]; };

__device__ __forceinline__ CubicTaps cubic_taps(int d, double scale) {
    // fx = (float)((dx+0.5)*scale_x - 0.5); sx = cvFloor(fx); fx -= sx;  ialpha = saturate_cast<short>(coeff * 2048)
    float f = __double2float_rn(__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5));
    const int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    const float A = -0.75f;
    const float xp1 = __fadd_rn(f, 1.f), omx = __fsub_rn(1.f, f);
    float c[4];
    c[0] = __fsub_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fsub_rn(__fmul_rn(A, xp1), 5.f * A), xp1), 8.f * A), xp1), 4.f * A);
    c[1] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, f), A + 3.f), f), f), 1.f);
    c[2] = __fadd_rn(__fmul_rn(__fmul_rn(__fsub_rn(__fmul_rn(A + 2.f, omx), A + 3.f), omx), omx), 1.f);
    c[3] = __fsub_rn(__fsub_rn(__fsub_rn(1.f, c[0]), c[1]), c[2]);
    CubicTaps r;
    r.ofs = s;
#pragma unroll
    for (int k = 0; k < 4; ++k) r.t[k] = __float2int_rn(__fmul_rn(c[k], 2048.f));
    return r;
}

__global__ void preprocess_lq_kernel(const uint8_t* __restrict__ img, int h, int w, int cn, double scale_x, double scale_y,
                                     int dh, int dw, float* __restrict__ lq, uint8_t* __restrict__ lq_u8, int out_h, int out_w) {
    mn_pdl_prologue();
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= out_h * out_w * cn) return;
    const int c = idx % cn;
    const int dx = (idx / cn) % out_w;
    const int dy = idx / (cn * out_w);
    int v = 0;                                            // the canvas is zero outside the resized image (test_sr.py:104-106)
    if (dx < dw && dy < dh) {
        const CubicTaps tx = cubic_taps(dx, scale_x), ty = cubic_taps(dy, scale_y);
        int S[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int yy = min(max(ty.ofs - 1 + r, 0), h - 1);
            int acc = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int xx = min(max(tx.ofs - 1 + j, 0), w - 1);
                acc += (int)img[((size_t)yy * w + xx) * cn + c] * tx.t[j];
            }
            S[r] = acc;
        }
        const int e = dx * cn + c, nvec = (dw * cn / 8) * 8;
        if (e < nvec) {                                   // vector body of VResizeCubicVec_32s8u: fp32, mul then add, rows 3..0
            const float sc = 1.f / (2048.f * 2048.f);
            float acc = __fmul_rn((float)S[3], __fmul_rn((float)ty.t[3], sc));
            acc = __fadd_rn(__fmul_rn((float)S[2], __fmul_rn((float)ty.t[2], sc)), acc);
            acc = __fadd_rn(__fmul_rn((float)S[1], __fmul_rn((float)ty.t[1], sc)), acc);
            acc = __fadd_rn(__fmul_rn((float)S[0], __fmul_rn((float)ty.t[0], sc)), acc);
            v = __float2int_rn(acc);
        } else {                                          // scalar tail: FixedPtCast<int, uchar, 22>
            v = (S[0] * ty.t[0] + S[1] * ty.t[1] + S[2] * ty.t[2] + S[3] * ty.t[3] + (1 << 21)) >> 22;
        }
        v = min(max(v, 0), 255);
        if (lq_u8) lq_u8[((size_t)dy * dw + dx) * cn + c] = (uint8_t)v;
    }
    // ToTensor: float(u8) / 255 ; Normalize: (x - 0.5) / 0.5
    const float t = __fdiv_rn(__fsub_rn(__fdiv_rn((float)v, 255.f), 0.5f), 0.5f);
    lq[((size_t)c * out_h + dy) * out_w + dx] = t;
}

__global__ void postprocess_sr_kernel(const float* __restrict__ sr, long long sn, long long sc, long long sh, long long sw,
                                      uint8_t* __restrict__ out, int B, int C, int H, int W) {
    mn_pdl_prologue();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)B * H * W) return;
    const int x = (int)(idx % W), y = (int)((idx / W) % H), b = (int)(idx / ((long long)W * H));
    const float* p = sr + b * sn + y * sh + x * sw;
    uint8_t* o = out + idx * C;
    for (int c = 0; c < C; ++c) {
        float v = __fadd_rn(__fmul_rn(p[c * sc], 0.5f), 0.5f);
        v = fminf(fmaxf(v, 0.f), 1.f);
        const int q = __float2int_rn(__fmul_rn(v, 255.f));
        o[C - 1 - c] = (uint8_t)min(max(q, 0), 255);          // .flip(2): channel c lands in byte C-1-c
    }
}

}  // namespace

extern "C" int mn_preprocess_lq_u8(const uint8_t* img, int h, int w, int cn, double fx, double fy, int dh, int dw,
                                   float* lq, uint8_t* lq_u8, int out_h, int out_w, void* stream) {
    MN_REQUIRE(img && lq && h > 0 && w > 0 && cn > 0 && cn <= 4 && fx > 0.0 && fy > 0.0, "mn_preprocess_lq_u8: bad args");
    MN_REQUIRE(dh > 0 && dw > 0 && dh <= out_h && dw <= out_w, "mn_preprocess_lq_u8: resized image (%dx%d) does not fit the %dx%d canvas "
               "(test_sr.py:109 skips such images)", dh, dw, out_h, out_w);
    MN_REQUIRE((long long)out_h * out_w * cn < (1ll << 31), "mn_preprocess_lq_u8: canvas too large");
    const int total = out_h * out_w * cn;
    MN_CUDA_CHECK((mn_launch(preprocess_lq_kernel, dim3(mn_cdiv(total, 128)), dim3(128), 0, (cudaStream_t)stream, img, h, w, cn,
                             1.0 / fx, 1.0 / fy, dh, dw, lq, lq_u8, out_h, out_w)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}

extern "C" int mn_postprocess_sr_u8(const float* sr, long long stride_n, long long stride_c, long long stride_h, long long stride_w,
                                    uint8_t* out, int B, int C, int H, int W, void* stream) {
    MN_REQUIRE(sr && out && B > 0 && C > 0 && C <= 4 && H > 0 && W > 0, "mn_postprocess_sr_u8: bad args");
    const long long total = (long long)B * H * W;
    MN_CUDA_CHECK((mn_launch(postprocess_sr_kernel, dim3((unsigned)mn_cdiv64(total, 256)), dim3(256), 0, (cudaStream_t)stream, sr, stride_n, stride_c,
                             stride_h, stride_w, out, B, C, H, W)));
    MN_LAUNCH_CHECK();
    return MN_OK;
}
