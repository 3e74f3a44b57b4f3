// CUDA-core kernels around the tcgen05 convolutions of the UNet forward (Path B of DESIGN.md).  All are bandwidth-bound
// elementwise / small-stencil passes over NHWC bf16 activations, vectorised to 16 bytes (8 channels) per thread.
//   first conv   core/models/trunks/unet.py:20  DoubleConv(n_channels_in, 64) first 3x3 (K = 9*C_in is too small for UMMA)
//   max pool     core/models/trunks/unet_parts.py:34  nn.MaxPool2d(2)
//   upsample     core/models/trunks/unet_parts.py:50,63  bilinear x2 align_corners=True, then F.pad to the skip's size
//   head         core/models/finallayers/quantile_layer.py:15-20  three 3x3 convs 32 -> C_out stacked on dim 1
#include <cuda_bf16.h>

#include "common.cuh"

namespace im2im {
namespace {

union Bf16x8 {
    uint4 u;
    __nv_bfloat162 h[4];
};

// Eight consecutive channels of an NHWC activation as fp32 registers, for both storage types of the UNet engines:
// bf16 (16 bytes) and - reference-precision tf32 mode - fp32 (32 bytes, values rounded onto the TF32 grid on store,
// so that what is stored is exactly what the next kind::tf32 MMA reads).
template <typename T>
struct Vec8;
template <>
struct Vec8<__nv_bfloat16> {
    static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* f) {
        Bf16x8 v;
        v.u = *reinterpret_cast<const uint4*>(p);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 t = __bfloat1622float2(v.h[j]);
            f[2 * j] = t.x;
            f[2 * j + 1] = t.y;
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* p, const float* f) {
        Bf16x8 v;
#pragma unroll
        for (int j = 0; j < 4; ++j) v.h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
        *reinterpret_cast<uint4*>(p) = v.u;
    }
};
__device__ __forceinline__ float aux_round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
template <>
struct Vec8<float> {
    static __device__ __forceinline__ void load(const float* p, float* f) {
        const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* p, const float* f) {
        float4 a, b;
        a.x = aux_round_tf32(f[0]); a.y = aux_round_tf32(f[1]); a.z = aux_round_tf32(f[2]); a.w = aux_round_tf32(f[3]);
        b.x = aux_round_tf32(f[4]); b.y = aux_round_tf32(f[5]); b.z = aux_round_tf32(f[6]); b.w = aux_round_tf32(f[7]);
        *reinterpret_cast<float4*>(p) = a;
        *reinterpret_cast<float4*>(p + 4) = b;
    }
};

// ---------------------------------------------------------------------------------------------------------------
// x: fp32 NCHW [B, c_in, H, W] (c_in <= 8) -> y: NHWC [B, H, W, c_out] (bf16, or fp32 in tf32 mode), 3x3 pad 1, + bias, ReLU.
// One thread = FOUR horizontally adjacent pixels x 16 output channels: the 3x6 input window is loaded once (18 loads for
// 4 pixels) and every 128-bit weight load from shared memory feeds 16 FMAs (ncu on the 2-pixel version: 42 % of the stall
// samples were short_scoreboard = waiting for those shared-memory loads, issue slots 53 % busy).  Weights fp32
// [c_out, c_in, 3, 3] are staged transposed in shared memory.
template <typename T>
__global__ void __launch_bounds__(256, 2) conv_first_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, int B, int c_in, int H,
                                                         int W, int c_out, int relu, T* __restrict__ y) {
    extern __shared__ __align__(16) float s_w[];  // [c_in*9][c_out] (a thread's 16 channels are contiguous), bias [c_out]
    const int kk = c_in * 9;
    for (int i = threadIdx.x; i < c_out * kk; i += blockDim.x) s_w[(i % kk) * c_out + i / kk] = w[i];
    float* s_b = s_w + c_out * kk;
    for (int i = threadIdx.x; i < c_out; i += blockDim.x) s_b[i] = bias ? bias[i] : 0.f;
    __syncthreads();
    constexpr int kPx = 4;
    const int groups = c_out / 16;
    const int wq = (W + kPx - 1) / kPx;  // pixel quads per row
    const long long total = static_cast<long long>(B) * H * wq * groups;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(e % groups);
        long long pp = e / groups;
        const int xw = kPx * static_cast<int>(pp % wq);
        const int yh = static_cast<int>((pp / wq) % H);
        const int b = static_cast<int>(pp / (static_cast<long long>(wq) * H));
        float acc[kPx][16];
#pragma unroll
        for (int px = 0; px < kPx; ++px)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[px][j] = s_b[g * 16 + j];
        for (int ci = 0; ci < c_in; ++ci) {
            const float* xp = x + (static_cast<long long>(b) * c_in + ci) * H * W;
            float v[3][kPx + 2];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < kPx + 2; ++c) {
                    const int yy = yh + r - 1, xx = xw + c - 1;
                    v[r][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(xp + static_cast<long long>(yy) * W + xx) : 0.f;
                }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float* wt = s_w + (ci * 9 + t) * c_out + g * 16;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 wv = *reinterpret_cast<const float4*>(wt + 4 * q);
#pragma unroll
                    for (int px = 0; px < kPx; ++px) {
                        const float a = v[t / 3][t % 3 + px];
                        acc[px][4 * q + 0] = fmaf(a, wv.x, acc[px][4 * q + 0]);
                        acc[px][4 * q + 1] = fmaf(a, wv.y, acc[px][4 * q + 1]);
                        acc[px][4 * q + 2] = fmaf(a, wv.z, acc[px][4 * q + 2]);
                        acc[px][4 * q + 3] = fmaf(a, wv.w, acc[px][4 * q + 3]);
                    }
                }
            }
        }
        const long long pix = (static_cast<long long>(b) * H + yh) * W + xw;
#pragma unroll
        for (int px = 0; px < kPx; ++px) {
            if (xw + px >= W) break;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = relu ? fmaxf(acc[px][8 * half + j], 0.f) : acc[px][8 * half + j];
                Vec8<T>::store(y + (pix + px) * c_out + g * 16 + 8 * half, o);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 2x2 max pool, stride 2 (floor), NHWC bf16.
template <typename T>
__global__ void __launch_bounds__(256) maxpool2x2_kernel(const T* __restrict__ x, int B, int H, int W,
                                                         int C, T* __restrict__ y) {
    // grid = (ceil(Wo*groups / 256), Ho, B): row and image come from the block index, so the per-thread index math is one
    // 32-bit division (the 64-bit div/mod chain of a flat index cost more instructions than the pooling itself)
    const int Wo = W / 2, groups = C / 8;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < Wo * groups) {
        const int g = idx % groups, ox = idx / groups;
        const int oy = blockIdx.y, b = blockIdx.z;
        const long long pix = (static_cast<long long>(b) * gridDim.y + oy) * Wo + ox;
        const T* p = x + ((static_cast<long long>(b) * H + 2 * oy) * W + 2 * ox) * C + g * 8;
        float a[8], c[8], d[8], f[8], o[8];
        Vec8<T>::load(p, a);
        Vec8<T>::load(p + C, c);
        Vec8<T>::load(p + static_cast<long long>(W) * C, d);
        Vec8<T>::load(p + static_cast<long long>(W) * C + C, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaxf(a[j], c[j]), fmaxf(d[j], f[j]));   // exact in either storage type
        Vec8<T>::store(y + pix * C + g * 8, o);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Bilinear x2 upsample with align_corners=True of x [B,h,w,C], written into y [B,Ho,Wo,C] at offset (pad_top, pad_left)
// with zeros elsewhere (F.pad to the skip connection's size).  Source index and weights follow ATen's
// upsample_bilinear2d: src = dst * (in-1)/(out-1) in fp32, i0 = (int)src, frac = src - i0.
template <typename T, int CPT>   // CPT = 8-channel chunks per thread (2 when C % 16 == 0: the index / weight math is shared)
__global__ void __launch_bounds__(256) upsample2x_kernel(const T* __restrict__ x, int B, int h, int w,
                                                         int C, int Ho, int Wo, int pad_top, int pad_left,
                                                         T* __restrict__ y) {
    const int uh = 2 * h, uw = 2 * w, groups = C / (8 * CPT);
    const float sy = uh > 1 ? static_cast<float>(h - 1) / static_cast<float>(uh - 1) : 0.f;
    const float sx = uw > 1 ? static_cast<float>(w - 1) / static_cast<float>(uw - 1) : 0.f;
    // grid = (ceil(Wo*groups / 256), Ho, B): see maxpool2x2_kernel (ncu: this kernel is instruction-bound, issue 82 %)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < Wo * groups) {
        const int g = idx % groups, ox = idx / groups;
        const int oy = blockIdx.y, b = blockIdx.z;
        const long long pix = (static_cast<long long>(b) * Ho + oy) * Wo + ox;
        const int uy = oy - pad_top, ux = ox - pad_left;
        T* dst = y + pix * C + g * (8 * CPT);
        if (uy < 0 || uy >= uh || ux < 0 || ux >= uw) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
            for (int c = 0; c < CPT; ++c) Vec8<T>::store(dst + 8 * c, o);
            return;
        }
        const float fy = sy * uy, fx = sx * ux;
        const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
        const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
        const float ly = fy - y0, lx = fx - x0;
        const float hy = 1.f - ly, hx = 1.f - lx;
        const T* base = x + static_cast<long long>(b) * h * w * C + g * (8 * CPT);
        const T* p00 = base + (static_cast<long long>(y0) * w + x0) * C;
        const T* p01 = base + (static_cast<long long>(y0) * w + x1) * C;
        const T* p10 = base + (static_cast<long long>(y1) * w + x0) * C;
        const T* p11 = base + (static_cast<long long>(y1) * w + x1) * C;
        // bf16 storage: four pre-multiplied weights (4 FMAs per value instead of 7 flops); the result differs from ATen's
        // factored form by an fp32 ulp, far below the bf16 rounding of the store.  The fp32 (tf32-mode) instantiation keeps
        // ATen's exact expression.
        const float w00 = hy * hx, w01 = hy * lx, w10 = ly * hx, w11 = ly * lx;
#pragma unroll
        for (int c = 0; c < CPT; ++c) {
            float v00[8], v01[8], v10[8], v11[8], o[8];
            Vec8<T>::load(p00 + 8 * c, v00);
            Vec8<T>::load(p01 + 8 * c, v01);
            Vec8<T>::load(p10 + 8 * c, v10);
            Vec8<T>::load(p11 + 8 * c, v11);
            if (sizeof(T) == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaf(w00, v00[j], fmaf(w01, v01[j], fmaf(w10, v10[j], w11 * v11[j])));
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = hy * (hx * v00[j] + lx * v01[j]) + ly * (hx * v10[j] + lx * v11[j]);
            }
            Vec8<T>::store(dst + 8 * c, o);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Head: x bf16 NHWC (c_mid channels used, row stride c_stride) -> y fp32 [B, n_out, H, W] (n_out = 3*C_out planes:
// lower.., prediction.., upper..), 3x3 pad 1, fp32 weights [n_out, c_mid, 3, 3] + bias.  One thread = one pixel, all outputs.
// Weights sit in shared memory as [tap][c][NP] with NP = n_out padded to a multiple of 4, so one 128-bit broadcast load
// feeds 4 FMAs.  tap_bias (optional, [n_out][9]) is added once per IN-RANGE tap: it carries the bias of a 1x1 convolution
// that was folded into these weights (inference: OutConv 64->32 composed with the head, exact incl. the zero padding).
template <int N_OUT, typename T>
__global__ void __launch_bounds__(128) head_conv_kernel(const T* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ bias,
                                                        const float* __restrict__ tap_bias, int B, int H, int W,
                                                        int c_mid, int c_stride, int act_kind, int act_from,
                                                        float* __restrict__ y) {
    constexpr int NP = (N_OUT + 3) / 4 * 4;
    extern __shared__ __align__(16) float s_w[];  // [tap][c_mid][NP]
    for (int i = threadIdx.x; i < 9 * c_mid * NP; i += blockDim.x) {
        const int o = i % NP, c = (i / NP) % c_mid, t = i / (NP * c_mid);
        s_w[i] = o < N_OUT ? w[(static_cast<long long>(o) * c_mid + c) * 9 + t] : 0.f;
    }
    __syncthreads();
    const long long total = static_cast<long long>(B) * H * W;
    for (long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < total;
         pix += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int xw = static_cast<int>(pix % W);
        const int yh = static_cast<int>((pix / W) % H);
        const long long b = pix / (static_cast<long long>(W) * H);
        float acc[NP];
#pragma unroll
        for (int o = 0; o < NP; ++o) acc[o] = (bias && o < N_OUT) ? __ldg(bias + o) : 0.f;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int yy = yh + t / 3 - 1, xx = xw + t % 3 - 1;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            if (tap_bias) {
#pragma unroll
                for (int o = 0; o < N_OUT; ++o) acc[o] += __ldg(tap_bias + o * 9 + t);
            }
            const T* p = x + ((b * H + yy) * W + xx) * c_stride;
            const float* wt = s_w + t * c_mid * NP;
            for (int c = 0; c < c_mid; c += 8) {
                float v[8];
                Vec8<T>::load(p + c, v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = make_float2(v[2 * j], v[2 * j + 1]);
#pragma unroll
                    for (int q = 0; q < NP / 4; ++q) {
                        const float4 wa = *reinterpret_cast<const float4*>(wt + (c + 2 * j) * NP + 4 * q);
                        const float4 wb = *reinterpret_cast<const float4*>(wt + (c + 2 * j + 1) * NP + 4 * q);
                        acc[4 * q + 0] = fmaf(f.x, wa.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(f.x, wa.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(f.x, wa.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(f.x, wa.w, acc[4 * q + 3]);
                        acc[4 * q + 0] = fmaf(f.y, wb.x, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(f.y, wb.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(f.y, wb.z, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(f.y, wb.w, acc[4 * q + 3]);
                    }
                }
            }
        }
        const long long hw = static_cast<long long>(H) * W;
#pragma unroll
        for (int o = 0; o < N_OUT; ++o) {
            float v = acc[o];
            if (o >= act_from) {  // gaussian head: relu on the variance planes; residual heads: abs on the magnitude
                if (act_kind == 1) v = (v != v) ? v : fmaxf(v, 0.f);
                else if (act_kind == 2) v = fabsf(v);
            }
            y[(b * N_OUT + o) * hw + static_cast<long long>(yh) * W + xw] = v;
        }
    }
}

// fp32 Conv2d weight [c_out, c_in, k, k] -> bf16 [c_out, taps, c_in] (forward GEMM operand, K-major rows) and, when
// out_bwd != null, bf16 [c_in, taps, c_out] with the taps reversed (180 degree rotation) = the data-gradient operand.
__global__ void __launch_bounds__(256) pack_conv_weights_kernel(const float* __restrict__ w, int c_out, int c_in,
                                                                int taps, __nv_bfloat16* __restrict__ out_fwd,
                                                                __nv_bfloat16* __restrict__ out_bwd) {
    const long long n = static_cast<long long>(c_out) * taps * c_in;
    const long long total = out_bwd ? 2 * n : n;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        if (e < n) {  // out_fwd[co][t][ci]
            const int ci = static_cast<int>(e % c_in);
            const int t = static_cast<int>((e / c_in) % taps);
            const int co = static_cast<int>(e / (static_cast<long long>(c_in) * taps));
            out_fwd[e] = __float2bfloat16_rn(w[(static_cast<long long>(co) * c_in + ci) * taps + t]);
        } else {      // out_bwd[ci][t'][co] = w[co][ci][taps-1-t']
            const long long f = e - n;
            const int co = static_cast<int>(f % c_out);
            const int t = static_cast<int>((f / c_out) % taps);
            const int ci = static_cast<int>(f / (static_cast<long long>(c_out) * taps));
            out_bwd[f] = __float2bfloat16_rn(w[(static_cast<long long>(co) * c_in + ci) * taps + (taps - 1 - t)]);
        }
    }
}

unsigned grid_for(long long work_items, int threads) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = 16ll * sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<unsigned>(blocks);
}

}  // namespace
}  // namespace im2im

using namespace im2im;

namespace im2im { namespace {
template <typename T>
int conv_first_launch(const float* d_x, const float* d_weight, const float* d_bias, int B, int c_in, int H, int W, int c_out,
                      int relu, T* d_out, void* stream) {
    if (B <= 0 || H <= 0 || W <= 0 || c_in <= 0 || c_in > 8) return fail(IM2IM_ERANGE, "conv_first: bad shape (c_in=%d)", c_in);
    if (c_out <= 0 || c_out % 16) return fail(IM2IM_ERANGE, "conv_first: c_out must be a multiple of 16");
    if (!d_x || !d_weight || !d_out) return fail(IM2IM_EINVAL, "null tensor");
    const size_t smem = sizeof(float) * (static_cast<size_t>(c_out) * c_in * 9 + c_out);
    if (smem > 48 * 1024) return fail(IM2IM_ERANGE, "conv_first: weights do not fit shared memory");
    const long long items = static_cast<long long>(B) * H * ((W + 3) / 4) * (c_out / 16);
    conv_first_kernel<T><<<grid_for(items, 256), 256, smem, static_cast<cudaStream_t>(stream)>>>(
        d_x, d_weight, d_bias, B, c_in, H, W, c_out, relu, d_out);
    return check_launch("conv_first_kernel");
}

template <typename T>
int maxpool_launch(const T* d_x, int B, int H, int W, int C, T* d_out, void* stream) {
    if (B <= 0 || H < 2 || W < 2 || C <= 0 || C % 8) return fail(IM2IM_ERANGE, "maxpool: bad shape");
    if (!d_x || !d_out) return fail(IM2IM_EINVAL, "null tensor");
    if (B > 65535 || H / 2 > 65535) return fail(IM2IM_ERANGE, "maxpool: B and H/2 must be <= 65535");
    const dim3 pgrid(static_cast<unsigned>(((W / 2) * (C / 8) + 255) / 256), static_cast<unsigned>(H / 2), static_cast<unsigned>(B));
    maxpool2x2_kernel<T><<<pgrid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, B, H, W, C, d_out);
    return check_launch("maxpool2x2_kernel");
}

template <typename T>
int upsample_launch(const T* d_x, int B, int h, int w, int C, int H_out, int W_out, T* d_out, void* stream) {
    if (B <= 0 || h <= 0 || w <= 0 || C <= 0 || C % 8) return fail(IM2IM_ERANGE, "upsample: bad shape");
    if (H_out < 2 * h || W_out < 2 * w) return fail(IM2IM_ERANGE, "upsample: output smaller than 2x input");
    if (!d_x || !d_out) return fail(IM2IM_EINVAL, "null tensor");
    const int pad_top = (H_out - 2 * h) / 2, pad_left = (W_out - 2 * w) / 2;  // F.pad split of unet_parts.py:63-64
    if (B > 65535 || H_out > 65535) return fail(IM2IM_ERANGE, "upsample: B and H_out must be <= 65535");
    if (C % 16 == 0) {
        const dim3 ugrid(static_cast<unsigned>((W_out * (C / 16) + 255) / 256), static_cast<unsigned>(H_out), static_cast<unsigned>(B));
        upsample2x_kernel<T, 2><<<ugrid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, B, h, w, C, H_out, W_out, pad_top,
                                                                                     pad_left, d_out);
    } else {
        const dim3 ugrid(static_cast<unsigned>((W_out * (C / 8) + 255) / 256), static_cast<unsigned>(H_out), static_cast<unsigned>(B));
        upsample2x_kernel<T, 1><<<ugrid, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_x, B, h, w, C, H_out, W_out, pad_top,
                                                                                     pad_left, d_out);
    }
    return check_launch("upsample2x_kernel");
}

template <typename T>
int head_conv_launch(const T* x, const float* d_weight, const float* d_bias, const float* d_tap_bias, int B, int H, int W,
                     int c_mid, int c_stride, int n_out, int act_kind, int act_from_plane, float* d_out, void* stream) {
    if (B <= 0 || H <= 0 || W <= 0 || c_mid <= 0 || c_mid % 8 || c_stride < c_mid || c_stride % 8)
        return fail(IM2IM_ERANGE, "head: bad shape");
    if (!x || !d_weight || !d_out) return fail(IM2IM_EINVAL, "null tensor");
    if (act_kind < 0 || act_kind > 2) return fail(IM2IM_EINVAL, "head: act_kind=%d", act_kind);
    if (act_kind == 0) act_from_plane = n_out;
    const size_t smem = sizeof(float) * 9 * c_mid * ((n_out + 3) / 4 * 4);
    if (smem > 48 * 1024) return fail(IM2IM_ERANGE, "head: weights do not fit shared memory");
    const long long items = static_cast<long long>(B) * H * W;
    const unsigned grid = grid_for(items, 128);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define IM2IM_HEAD_CASE(N)                                                                                              \
    case N:                                                                                                             \
        head_conv_kernel<N, T><<<grid, 128, smem, st>>>(x, d_weight, d_bias, d_tap_bias, B, H, W, c_mid, c_stride, act_kind, \
                                                        act_from_plane, d_out);                                         \
        break
    switch (n_out) {
        IM2IM_HEAD_CASE(2); IM2IM_HEAD_CASE(3); IM2IM_HEAD_CASE(4); IM2IM_HEAD_CASE(6); IM2IM_HEAD_CASE(9);
        default: return fail(IM2IM_ENOTSUP, "head: n_out=%d (supported: 2, 3, 4, 6, 9 output planes)", n_out);
    }
#undef IM2IM_HEAD_CASE
    return check_launch("head_conv_kernel");
}
} }

extern "C" int im2im_conv_first_bf16(const float* d_x, const float* d_weight, const float* d_bias, int32_t B,
                                     int32_t c_in, int32_t H, int32_t W, int32_t c_out, int32_t relu, void* d_out,
                                     void* stream) {
    return conv_first_launch(d_x, d_weight, d_bias, B, c_in, H, W, c_out, relu, static_cast<__nv_bfloat16*>(d_out), stream);
}

extern "C" int im2im_maxpool2x2_bf16(const void* d_x, int32_t B, int32_t H, int32_t W, int32_t C, void* d_out,
                                     void* stream) {
    return maxpool_launch(static_cast<const __nv_bfloat16*>(d_x), B, H, W, C, static_cast<__nv_bfloat16*>(d_out), stream);
}

extern "C" int im2im_upsample2x_bilinear_bf16(const void* d_x, int32_t B, int32_t h, int32_t w, int32_t C,
                                              int32_t H_out, int32_t W_out, void* d_out, void* stream) {
    return upsample_launch(static_cast<const __nv_bfloat16*>(d_x), B, h, w, C, H_out, W_out,
                           static_cast<__nv_bfloat16*>(d_out), stream);
}

extern "C" int im2im_head_conv3x3_act_f32(const void* d_x, const float* d_weight, const float* d_bias,
                                          const float* d_tap_bias, int32_t B, int32_t H, int32_t W, int32_t c_mid,
                                          int32_t c_stride, int32_t n_out, int32_t act_kind, int32_t act_from_plane,
                                          float* d_out, void* stream) {
    return head_conv_launch(static_cast<const __nv_bfloat16*>(d_x), d_weight, d_bias, d_tap_bias, B, H, W, c_mid, c_stride,
                            n_out, act_kind, act_from_plane, d_out, stream);
}

extern "C" int im2im_head_conv3x3_f32(const void* d_x, const float* d_weight, const float* d_bias,
                                      const float* d_tap_bias, int32_t B, int32_t H, int32_t W, int32_t c_mid,
                                      int32_t c_stride, int32_t n_out, float* d_out, void* stream) {
    return im2im_head_conv3x3_act_f32(d_x, d_weight, d_bias, d_tap_bias, B, H, W, c_mid, c_stride, n_out, 0, 0, d_out,
                                      stream);
}

// ---- reference-precision (tf32) mode: the same kernels on fp32 NHWC activations (values rounded onto the TF32 grid)
extern "C" int im2im_conv_first_nhwc_f32(const float* d_x, const float* d_weight, const float* d_bias, int32_t B,
                                         int32_t c_in, int32_t H, int32_t W, int32_t c_out, int32_t relu, float* d_out,
                                         void* stream) {
    return conv_first_launch(d_x, d_weight, d_bias, B, c_in, H, W, c_out, relu, d_out, stream);
}

extern "C" int im2im_maxpool2x2_nhwc_f32(const float* d_x, int32_t B, int32_t H, int32_t W, int32_t C, float* d_out,
                                         void* stream) {
    return maxpool_launch(d_x, B, H, W, C, d_out, stream);
}

extern "C" int im2im_upsample2x_bilinear_nhwc_f32(const float* d_x, int32_t B, int32_t h, int32_t w, int32_t C,
                                                  int32_t H_out, int32_t W_out, float* d_out, void* stream) {
    return upsample_launch(d_x, B, h, w, C, H_out, W_out, d_out, stream);
}

extern "C" int im2im_head_conv3x3_act_nhwc_f32(const float* d_x, const float* d_weight, const float* d_bias,
                                               const float* d_tap_bias, int32_t B, int32_t H, int32_t W, int32_t c_mid,
                                               int32_t c_stride, int32_t n_out, int32_t act_kind, int32_t act_from_plane,
                                               float* d_out, void* stream) {
    return head_conv_launch(d_x, d_weight, d_bias, d_tap_bias, B, H, W, c_mid, c_stride, n_out, act_kind, act_from_plane,
                            d_out, stream);
}

extern "C" int im2im_pack_conv_weights(const float* d_weight, int32_t c_out, int32_t c_in, int32_t taps, void* d_out_fwd,
                                       void* d_out_bwd, void* stream) {
    if (c_out <= 0 || c_in <= 0 || (taps != 9 && taps != 1)) return fail(IM2IM_EINVAL, "pack_conv_weights: bad shape");
    if (!d_weight || !d_out_fwd) return fail(IM2IM_EINVAL, "pack_conv_weights: null tensor");
    const long long n = static_cast<long long>(c_out) * c_in * taps * (d_out_bwd ? 2 : 1);
    pack_conv_weights_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_weight, c_out, c_in, taps, static_cast<__nv_bfloat16*>(d_out_fwd), static_cast<__nv_bfloat16*>(d_out_bwd));
    return check_launch("pack_conv_weights_kernel");
}
