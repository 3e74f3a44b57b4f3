// Training-side CUDA-core kernels of the UNet path (Path B of DESIGN.md, rows b4/b5 of SURVEY.md §8):
// BatchNorm (batch statistics) forward/backward fused with ReLU, max-pool / bilinear-upsample backward, the head's and the
// first convolution's gradients, the fused pinball+MSE loss with its gradient, and a fused Adam step.
// All activation tensors are NHWC bf16, statistics / parameters / gradients fp32.  Everything here is bandwidth-bound.
//   BatchNorm2d + ReLU      core/models/trunks/unet_parts.py:17-18,20-21 (train mode: per-replica batch statistics)
//   MaxPool2d / Upsample    unet_parts.py:34, :50,:63
//   loss                    core/models/finallayers/quantile_layer.py:23-32, core/models/losses/pinball.py:12-24
//   Adam                    core/scripts/train.py:120,162 (torch.optim.Adam defaults: betas .9/.999, eps 1e-8, no decay)
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"

namespace im2im {
namespace {

union Bf16x8 {
    uint4 u;
    __nv_bfloat162 h[4];
};

__device__ __forceinline__ void unpack8(const Bf16x8& v, float* f) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 t = __bfloat1622float2(v.h[j]);
        f[2 * j] = t.x;
        f[2 * j + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    Bf16x8 v;
#pragma unroll
    for (int j = 0; j < 4; ++j) v.h[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    return v.u;
}

// streaming 16-byte load / store (read once, written once: keep them out of L1)
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

unsigned grid_for(long long work_items, int threads, int per_sm = 8) {
    long long blocks = (work_items + threads - 1) / threads;
    const long long cap = static_cast<long long>(per_sm) * sm_count();
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<unsigned>(blocks);
}

// ---------------------------------------------------------------------------------------------------------------
// Per-channel reductions over pixels of NHWC bf16 tensors.  Block = 256 threads = (256/groups) pixel lanes x groups,
// groups = C/8; each thread keeps 8 channels in registers, partial sums are combined in shared memory and added to the
// global accumulators with one atomicAdd per channel per block.
//   MODE 0: sums[c] += z, sums[C+c] += z*z                                  (BN statistics; also plain channel sums)
//   MODE 1: xhat = (z-mean)*rstd;  g = dy * (xhat*gamma+beta > 0);  sums[c] += g, sums[C+c] += g*xhat  (BN+ReLU backward)
// (a = z or dy; in MODE 1 `a` is dy and `zt` is the saved pre-normalisation convolution output z)
template <int MODE>
__global__ void __launch_bounds__(256, 3) channel_reduce_kernel(const __nv_bfloat16* __restrict__ a,
                                                             const __nv_bfloat16* __restrict__ zt,
                                                             const float* __restrict__ gamma,
                                                             const float* __restrict__ beta,
                                                             const float* __restrict__ mean,
                                                             const float* __restrict__ rstd, long long n_pix, int C,
                                                             float* __restrict__ sums) {
    // Block-level combination through 2*C shared-memory accumulators (<= 8 KB): the 18 KB partial-sum array this replaced
    // kept the kernel from sharing an SM with the tensor-core kernels (which leave ~13 KB of shared memory free), so it
    // could not overlap the side-stream weight gradients it is meant to run next to.
    extern __shared__ float s_acc[];   // [2][C]
    for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) s_acc[c] = 0.f;
    __syncthreads();
    const int groups = C / 8;
    const int g = threadIdx.x % groups;
    const int lane_p = threadIdx.x / groups;
    const int lanes = 256 / groups;
    // MODE 1 accumulates sum(g) and sum(g*z) per thread; sum(g*xhat) = rstd*(sum(g*z) - mean*sum(g)) is formed once per
    // block below, which keeps mean/rstd out of the loop (fewer live registers -> three resident blocks per SM)
    float acc0[8], acc1[8], sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        acc0[j] = acc1[j] = 0.f;
        if (MODE == 1) {
            const float mu = mean[g * 8 + j];
            sc[j] = gamma[g * 8 + j] * rstd[g * 8 + j];    // the forward's scale/shift, recomputed with the same ops
            sh[j] = beta[g * 8 + j] - mu * sc[j];
        }
    }
    // kUnroll pixels per thread and iteration: all their 16-byte loads are issued before the first use, so that enough
    // bytes are in flight per SM to cover the HBM latency
    constexpr int kUnroll = 4;
    const long long stride = static_cast<long long>(gridDim.x) * lanes;
    for (long long p0 = static_cast<long long>(blockIdx.x) * lanes + lane_p; p0 < n_pix; p0 += stride * kUnroll) {
        Bf16x8 va[kUnroll], vz[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long p = p0 + u * stride;
            va[u].u = make_uint4(0u, 0u, 0u, 0u);
            vz[u].u = make_uint4(0u, 0u, 0u, 0u);
            if (p < n_pix) {
                va[u].u = ld_stream_u4(a + p * C + g * 8);
                if (MODE == 1) vz[u].u = ld_stream_u4(zt + p * C + g * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (p0 + u * stride >= n_pix) break;
            float fa[8];
            unpack8(va[u], fa);
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc0[j] += fa[j]; acc1[j] = fmaf(fa[j], fa[j], acc1[j]); }
            } else {
                float fz[8];
                unpack8(vz[u], fz);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float gj = fmaf(fz[j], sc[j], sh[j]) > 0.f ? fa[j] : 0.f;   // ReLU mask exactly as the forward saw it
                    acc0[j] += gj;
                    acc1[j] = fmaf(gj, fz[j], acc1[j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) { atomicAdd(&s_acc[g * 8 + j], acc0[j]); atomicAdd(&s_acc[C + g * 8 + j], acc1[j]); }
    __syncthreads();
    // thread c < C finalises channel c
    for (int c = threadIdx.x; c < C; c += 256) {
        float s0 = s_acc[c], s1 = s_acc[C + c];
        if (MODE == 1) s1 = rstd[c] * (s1 - mean[c] * s0);
        atomicAdd(&sums[c], s0);
        atomicAdd(&sums[C + c], s1);
    }
}

// mean/var from the sums -> per-channel affine (scale, shift) for y = relu(z*scale + shift), saved mean / rstd, and the
// running-statistics update of nn.BatchNorm2d (momentum, unbiased variance).  conv_bias (may be null) only shifts the
// mean that is recorded in running_mean: z is the bias-free convolution output and the batch mean cancels the bias.
__global__ void bn_finalize_kernel(const float* __restrict__ sums, double count, const float* __restrict__ conv_bias,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float momentum, int C, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean, float* __restrict__ save_rstd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double mean = static_cast<double>(sums[c]) / count;
    double var = static_cast<double>(sums[C + c]) / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float sc = gamma[c] * rstd;
    scale[c] = sc;
    shift[c] = beta[c] - static_cast<float>(mean) * sc;
    save_mean[c] = static_cast<float>(mean);
    save_rstd[c] = rstd;
    if (running_mean) {
        const float m_full = static_cast<float>(mean) + (conv_bias ? conv_bias[c] : 0.f);
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m_full;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
    }
}

// y = relu(z*scale + shift).  gridDim.x*256 is a multiple of groups = C/8 (a power of two <= 256 or a divisor of 256,
// checked by the host), so a thread's channel group never changes: the per-channel constants live in registers.
__global__ void __launch_bounds__(256) bn_apply_relu_kernel(const __nv_bfloat16* __restrict__ z,
                                                            const float* __restrict__ scale,
                                                            const float* __restrict__ shift, long long n_pix, int C,
                                                            __nv_bfloat16* __restrict__ y) {
    constexpr int kUnroll = 2;
    const int groups = C / 8;
    const long long total = n_pix * groups;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long e0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int g = static_cast<int>(e0 % groups);
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = __ldg(scale + g * 8 + j); sh[j] = __ldg(shift + g * 8 + j); }
    for (long long e = e0; e < total; e += stride * kUnroll) {
        Bf16x8 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (e + u * stride < total) v[u].u = ld_stream_u4(z + (e + u * stride) * 8);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (e + u * stride >= total) break;
            float f[8];
            unpack8(v[u], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
            *reinterpret_cast<uint4*>(y + (e + u * stride) * 8) = pack8(f);
        }
    }
}

// dz = gamma*rstd * (g - sum(g)/M - xhat * sum(g*xhat)/M),  xhat = (z-mean)*rstd,  g = dy * (xhat*gamma+beta > 0)
__global__ void __launch_bounds__(256, 3) bn_relu_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                const __nv_bfloat16* __restrict__ z,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta,
                                                                const float* __restrict__ mean,
                                                                const float* __restrict__ rstd,
                                                                const float* __restrict__ sums, float inv_count,
                                                                long long n_pix, int C, int premasked,
                                                                __nv_bfloat16* __restrict__ dz) {
    // premasked != 0: `dy` already is g = dy * relu_mask (written by the data-gradient convolution's fused epilogue)
    // a thread's channel group is loop-invariant (see bn_apply_relu_kernel): per-channel constants in registers.
    //   dz = sc*g - k0 - (z - mu)*k1 = sc*g + c0 - z*k1,   k0 = sc*sum(g)/M,  k1 = sc*rstd*sum(g*xhat)/M,  c0 = mu*k1 - k0
    // three resident blocks per SM (<= 85 registers) x 3 pixel groups per thread: 768 threads x 96 B in flight per SM
    constexpr int kUnroll = 3;
    const int groups = C / 8;
    const long long total = n_pix * groups;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const long long e0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int g = static_cast<int>(e0 % groups);
    float sc[8], sh[8], c0[8], k1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        const float rs = __ldg(rstd + c), mu = __ldg(mean + c);
        sc[j] = __ldg(gamma + c) * rs;
        sh[j] = __ldg(beta + c) - mu * sc[j];
        k1[j] = sc[j] * (rs * (__ldg(sums + C + c) * inv_count));
        c0[j] = mu * k1[j] - sc[j] * (__ldg(sums + c) * inv_count);   // dz = sc*g + c0 - z*k1
    }
    for (long long e = e0; e < total; e += stride * kUnroll) {
        Bf16x8 vd[kUnroll], vz[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            if (e + u * stride < total) {
                vd[u].u = ld_stream_u4(dy + (e + u * stride) * 8);
                vz[u].u = ld_stream_u4(z + (e + u * stride) * 8);
            }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (e + u * stride >= total) break;
            float fd[8], fz[8], o[8];
            unpack8(vd[u], fd);
            unpack8(vz[u], fz);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float gj = (premasked || fmaf(fz[j], sc[j], sh[j]) > 0.f) ? fd[j] : 0.f;
                o[j] = fmaf(-fz[j], k1[j], fmaf(sc[j], gj, c0[j]));
            }
            *reinterpret_cast<uint4*>(dz + (e + u * stride) * 8) = pack8(o);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Skip layers (x1..x4 of the UNet: the block output feeds the skip connection AND the next block's 2x2 max-pool,
// unet.py:35-39, unet_parts.py:34): BatchNorm+ReLU and the pool in ONE pass forward, and backward the pool's gradient
// folded into the BatchNorm backward, instead of a separate scatter pass that reads y and read-modify-writes the skip
// gradient.  One thread = one 2x2 window x 8 channels; H and W even.
//   forward   y = relu(z*scale + shift) (4 pixels), p = max of the four y (as stored, bf16)
//   backward  y is recomputed from the saved z (same two ops as the forward, so the same bf16 values), the window's
//             gradient d_p goes to its FIRST maximal element (ATen's order, like maxpool2x2_bwd_kernel), is added to the
//             gradient from the up path and rounded to bf16 exactly as the two-kernel sequence stores it; then
//             g = dy * relu_mask,  PASS 0: sums[c] += g, sums[C+c] += rstd*(sum(g*z) - mean*sum(g));
//                                  PASS 1: dz = sc*g + c0 - z*k1   (bn_relu_bwd_apply_kernel's formula)
__global__ void __launch_bounds__(256) bn_apply_relu_pool_kernel(const __nv_bfloat16* __restrict__ z,
                                                                 const float* __restrict__ scale,
                                                                 const float* __restrict__ shift, int B, int H, int W,
                                                                 int C, __nv_bfloat16* __restrict__ y,
                                                                 __nv_bfloat16* __restrict__ p) {
    const int Ho = H / 2, Wo = W / 2, groups = C / 8;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= Wo * groups) return;
    const int g = idx % groups, ox = idx / groups;
    const int oy = blockIdx.y, b = blockIdx.z;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = __ldg(scale + g * 8 + j); sh[j] = __ldg(shift + g * 8 + j); }
    const long long base = ((static_cast<long long>(b) * H + 2 * oy) * W + 2 * ox) * C + g * 8;
    const long long offs[4] = {0, C, static_cast<long long>(W) * C, static_cast<long long>(W) * C + C};
    Bf16x8 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q].u = ld_stream_u4(z + base + offs[q]);
    float m[8];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float f[8];
        unpack8(v[q], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = fmaxf(fmaf(f[j], sc[j], sh[j]), 0.f);
        Bf16x8 o;
        o.u = pack8(f);
        *reinterpret_cast<uint4*>(y + base + offs[q]) = o.u;
        float r[8];
        unpack8(o, r);                       // the values as stored: max commutes with the (monotone) rounding
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = q == 0 ? r[j] : fmaxf(m[j], r[j]);
    }
    const long long pix = (static_cast<long long>(b) * Ho + oy) * Wo + ox;
    *reinterpret_cast<uint4*>(p + pix * C + g * 8) = pack8(m);
}

// g[q][j] of one window (see above); z, d_skip: the four pixels, d_p: the pooled gradient
__device__ __forceinline__ void pool_window_grad(const Bf16x8* vz, const Bf16x8* vd, const Bf16x8& vp, const float* sc,
                                                 const float* sh, float z[4][8], float g[4][8]) {
    float dsk[4][8], dp[8], yv[4][8];
    bool pos[4][8];
    unpack8(vp, dp);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        unpack8(vz[q], z[q]);
        unpack8(vd[q], dsk[q]);
        float pre[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { pre[j] = fmaf(z[q][j], sc[j], sh[j]); pos[q][j] = pre[j] > 0.f; pre[j] = fmaxf(pre[j], 0.f); }
        Bf16x8 o;
        o.u = pack8(pre);
        unpack8(o, yv[q]);                   // y exactly as the forward stored it
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        int arg = 0;
        float m = yv[0][j];
#pragma unroll
        for (int q = 1; q < 4; ++q) if (yv[q][j] > m) { m = yv[q][j]; arg = q; }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            // what maxpool2x2_bwd_kernel(accumulate = 1) leaves in the skip gradient: bf16(prev + routed), prev untouched elsewhere
            float dy = dsk[q][j];
            if (q == arg) dy = __bfloat162float(__float2bfloat16_rn(dy + dp[j]));
            g[q][j] = pos[q][j] ? dy : 0.f;
        }
    }
}

template <int PASS>
__global__ void __launch_bounds__(256, 2) bn_pool_bwd_kernel(const __nv_bfloat16* __restrict__ d_skip,
                                                             const __nv_bfloat16* __restrict__ d_p,
                                                             const __nv_bfloat16* __restrict__ zt,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             float* __restrict__ sums, float inv_count, int B, int H, int W,
                                                             int C, __nv_bfloat16* __restrict__ dz) {
    extern __shared__ float s_acc[];   // PASS 0: [2][C]
    const int Ho = H / 2, Wo = W / 2, groups = C / 8;
    const int g = threadIdx.x % groups;
    const int lane_p = threadIdx.x / groups;
    const int lanes = 256 / groups;
    if (PASS == 0) {
        for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) s_acc[c] = 0.f;
        __syncthreads();
    }
    float sc[8], sh[8], c0[8], k1[8], acc0[8], acc1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        const float rs = __ldg(rstd + c), mu = __ldg(mean + c);
        sc[j] = __ldg(gamma + c) * rs;
        sh[j] = __ldg(beta + c) - mu * sc[j];
        acc0[j] = acc1[j] = 0.f;
        if (PASS == 1) {
            k1[j] = sc[j] * (rs * (__ldg(sums + C + c) * inv_count));
            c0[j] = mu * k1[j] - sc[j] * (__ldg(sums + c) * inv_count);
        }
    }
    const long long n_win = static_cast<long long>(B) * Ho * Wo;
    const long long stride = static_cast<long long>(gridDim.x) * lanes;
    const long long offs[4] = {0, C, static_cast<long long>(W) * C, static_cast<long long>(W) * C + C};
    for (long long wi = static_cast<long long>(blockIdx.x) * lanes + lane_p; wi < n_win; wi += stride) {
        const long long b = wi / (static_cast<long long>(Ho) * Wo);
        const int r = static_cast<int>(wi - b * Ho * Wo);
        const int oy = r / Wo, ox = r - oy * Wo;
        const long long base = ((b * H + 2 * oy) * W + 2 * ox) * C + g * 8;
        Bf16x8 vz[4], vd[4], vp;
#pragma unroll
        for (int q = 0; q < 4; ++q) { vz[q].u = ld_stream_u4(zt + base + offs[q]); vd[q].u = ld_stream_u4(d_skip + base + offs[q]); }
        vp.u = ld_stream_u4(d_p + wi * C + g * 8);
        float z[4][8], gg[4][8];
        pool_window_grad(vz, vd, vp, sc, sh, z, gg);
        if (PASS == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc0[j] += gg[q][j]; acc1[j] = fmaf(gg[q][j], z[q][j], acc1[j]); }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaf(-z[q][j], k1[j], fmaf(sc[j], gg[q][j], c0[j]));
                *reinterpret_cast<uint4*>(dz + base + offs[q]) = pack8(o);
            }
        }
    }
    if (PASS == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { atomicAdd(&s_acc[g * 8 + j], acc0[j]); atomicAdd(&s_acc[C + g * 8 + j], acc1[j]); }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += 256) {
            const float s0 = s_acc[c];
            const float s1 = rstd[c] * (s_acc[C + c] - mean[c] * s0);
            atomicAdd(&sums[c], s0);
            atomicAdd(&sums[C + c], s1);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Max-pool backward: the gradient of each 2x2 window goes to its first maximal element (ATen's argmax order);
// accumulate != 0 adds into dx (skip tensors receive a second gradient from the up path).
__global__ void __launch_bounds__(256) maxpool2x2_bwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                             const __nv_bfloat16* __restrict__ dy, int B, int H, int W,
                                                             int C, int accumulate, __nv_bfloat16* __restrict__ dx) {
    const int Ho = H / 2, Wo = W / 2, groups = C / 8;
    // one thread per (input 2x2 window, channel group); grid = (ceil(Wo*groups / 256), Ho, B) so that row and image come
    // from the block index (one 32-bit division per thread instead of a 64-bit div/mod chain)
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < Wo * groups) {
        const int g = idx % groups, ox = idx / groups;
        const int oy = blockIdx.y, b = blockIdx.z;
        const long long pix = (static_cast<long long>(b) * Ho + oy) * Wo + ox;
        const long long base = ((static_cast<long long>(b) * H + 2 * oy) * W + 2 * ox) * C + g * 8;
        const long long offs[4] = {0, C, static_cast<long long>(W) * C, static_cast<long long>(W) * C + C};
        float v[4][8], d[8], out[4][8];
        Bf16x8 t;
#pragma unroll
        for (int q = 0; q < 4; ++q) { t.u = *reinterpret_cast<const uint4*>(x + base + offs[q]); unpack8(t, v[q]); }
        t.u = *reinterpret_cast<const uint4*>(dy + pix * C + g * 8);
        unpack8(t, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int arg = 0;
            float m = v[0][j];
#pragma unroll
            for (int q = 1; q < 4; ++q) if (v[q][j] > m) { m = v[q][j]; arg = q; }
#pragma unroll
            for (int q = 0; q < 4; ++q) out[q][j] = (q == arg) ? d[j] : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (accumulate) {
                float prev[8];
                t.u = *reinterpret_cast<const uint4*>(dx + base + offs[q]);
                unpack8(t, prev);
#pragma unroll
                for (int j = 0; j < 8; ++j) out[q][j] += prev[j];
            }
            *reinterpret_cast<uint4*>(dx + base + offs[q]) = pack8(out[q]);
        }
    }
}

// Bilinear x2 (align_corners=True) + pad backward, gather form: dx[b,iy,ix,:] = sum over the (<=5x5) output pixels whose
// bilinear footprint contains (iy,ix), with the forward's own fp32 weights.  The per-axis weights of the five candidate
// output rows / columns (2i-2 .. 2i+2) are computed once per thread; the 25 taps are then independent predicated 16-byte
// loads (about 16 of them live), so many loads are in flight instead of a weight computation between every two.
__device__ __forceinline__ float upsample_axis_weight(int u, int n_out, int n_in, float scale, int i) {
    // weight with which input index i contributes to output index u (0 when u is out of range or i is not a neighbour)
    if (u < 0 || u >= n_out) return 0.f;
    const float f = scale * u;
    const int i0 = static_cast<int>(f);
    const int i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
    const float l = f - i0;
    float w = 0.f;
    if (i0 == i) w += 1.f - l;
    if (i1 == i) w += l;
    return w;
}

template <int CPT>   // 8-channel chunks per thread: 2 when C % 16 == 0 (the ten axis weights are shared by both chunks)
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const __nv_bfloat16* __restrict__ du, int B, int h, int w,
                                                             int C, int Ho, int Wo, int pad_top, int pad_left,
                                                             __nv_bfloat16* __restrict__ dx) {
    const int uh = 2 * h, uw = 2 * w, groups = C / (8 * CPT);
    const float sy = uh > 1 ? static_cast<float>(h - 1) / static_cast<float>(uh - 1) : 0.f;
    const float sx = uw > 1 ? static_cast<float>(w - 1) / static_cast<float>(uw - 1) : 0.f;
    // grid = (ceil(w*groups / 256), h, B): row and image from the block index
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < w * groups) {
        const int g = idx % groups, ix = idx / groups;
        const int iy = blockIdx.y, b = blockIdx.z;
        const long long pix = (static_cast<long long>(b) * h + iy) * w + ix;
        float wy[5], wx[5];
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            wy[a] = upsample_axis_weight(2 * iy - 2 + a, uh, h, sy, iy);
            wx[a] = upsample_axis_weight(2 * ix - 2 + a, uw, w, sx, ix);
        }
        const __nv_bfloat16* base = du + ((static_cast<long long>(b) * Ho + (2 * iy - 2 + pad_top)) * Wo + (2 * ix - 2 + pad_left)) * C +
                                    g * (8 * CPT);
        float acc[CPT][8];
#pragma unroll
        for (int c = 0; c < CPT; ++c)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;
        // rows outer (wy is uniform over the block: blockIdx.y = iy, so skipping a zero-weight row does not diverge), the
        // five column taps of a row are loaded together; fp32 FMAs in the order rows outer / columns inner
#pragma unroll
        for (int a = 0; a < 5; ++a) {
            if (wy[a] == 0.f) continue;
            Bf16x8 t[CPT][5];
#pragma unroll
            for (int c5 = 0; c5 < 5; ++c5)
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    t[c][c5].u = make_uint4(0u, 0u, 0u, 0u);
                    if (wx[c5] != 0.f)
                        t[c][c5].u = *reinterpret_cast<const uint4*>(base + (static_cast<long long>(a) * Wo + c5) * C + 8 * c);
                }
#pragma unroll
            for (int c5 = 0; c5 < 5; ++c5) {
                const float wgt = wy[a] * wx[c5];
#pragma unroll
                for (int c = 0; c < CPT; ++c) {
                    float f[8];
                    unpack8(t[c][c5], f);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[c][j] = fmaf(wgt, f[j], acc[c][j]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < CPT; ++c) *reinterpret_cast<uint4*>(dx + pix * C + g * (8 * CPT) + 8 * c) = pack8(acc[c]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Loss: w_lo*pinball_qlo(pred[:,0], y) + w_hi*pinball_qhi(pred[:,2], y) + w_mse*mse(pred[:,1], y), each a mean over
// B*C*H*W.  Writes d loss / d pred and accumulates the three partial sums (double) into loss_parts[3].
// pinball(e = out - target): q*|e| for e<0, (1-q)*|e| for e>0, 0 at e == 0  ->  d/d out = -q, (1-q), 0.
__global__ void __launch_bounds__(256) quantile_loss_kernel(const float* __restrict__ pred,
                                                            const float* __restrict__ target, long long n_images,
                                                            long long px, float q_lo, float q_hi, float w_lo,
                                                            float w_hi, float w_mse, float inv_count,
                                                            float* __restrict__ dpred, double* __restrict__ loss_parts) {
    double part[3] = {0.0, 0.0, 0.0};
    const long long total = n_images * px;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long i = e / px, k = e - i * px;
        const float t = target[e];
        const long long o = i * 3 * px + k;
        const float e_lo = pred[o] - t, e_mid = pred[o + px] - t, e_hi = pred[o + 2 * px] - t;
        part[0] += e_lo < 0.f ? q_lo * fabsf(e_lo) : (e_lo > 0.f ? (1.f - q_lo) * fabsf(e_lo) : 0.f);
        part[1] += e_hi < 0.f ? q_hi * fabsf(e_hi) : (e_hi > 0.f ? (1.f - q_hi) * fabsf(e_hi) : 0.f);
        part[2] += e_mid * e_mid;
        if (dpred) {
            dpred[o] = w_lo * inv_count * (e_lo < 0.f ? -q_lo : (e_lo > 0.f ? 1.f - q_lo : 0.f));
            dpred[o + px] = w_mse * inv_count * 2.f * e_mid;
            dpred[o + 2 * px] = w_hi * inv_count * (e_hi < 0.f ? -q_hi : (e_hi > 0.f ? 1.f - q_hi : 0.f));
        }
    }
    __shared__ double s_part[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double v = part[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[q][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int wv = 0; wv < 8; ++wv) v += s_part[threadIdx.x][wv];
        atomicAdd(&loss_parts[threadIdx.x], v);
    }
}

// Training losses of the other heads, value parts + d loss / d pred in one pass (pred (B, planes, px), target (B, px)).
//   IM2IM_LOSS_QUANTILES_L1   quantile_l1_layer.py:23-32   w_lo*pinball(q_lo) + w_hi*pinball(q_hi) + w_mse*L1(pred)
//   IM2IM_LOSS_GAUSSIAN       gaussian_layer.py:20-24      nn.GaussianNLLLoss(): 0.5*(log(max(var,1e-6)) + (mean-y)^2/max(var,1e-6));
//                                                          torch clamps the variance on a detached copy, so the gradient
//                                                          uses the clamped value and flows to var unchanged
//   IM2IM_LOSS_RESIDUAL       residual_magnitude_layer.py:20-26     MSE(pred,y) + MSE(r, |y-pred|)
//   IM2IM_LOSS_RESIDUAL_L1    residual_magnitude_l1_layer.py:20-26  L1(pred,y)  + MSE(r, |y-pred|)
//   IM2IM_LOSS_INN            inn_layer.py:23-28, losses/inn.py:11-14   MSE(pred,y) + mean(relu(y-u)^2 + relu(l-y)^2 + beta*|u-l|)
// part[k] are plain sums; the host forms sum_k w_k * part[k] / count.  w[] also scales the gradients.
__device__ __forceinline__ float sgnf(float x) { return x > 0.f ? 1.f : (x < 0.f ? -1.f : 0.f); }

template <int KIND>
__global__ void __launch_bounds__(256) head_loss_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        long long n_images, long long px, float q_lo, float q_hi,
                                                        float w0, float w1, float w2, float beta, float inv_count,
                                                        float* __restrict__ dpred, double* __restrict__ loss_parts) {
    constexpr int kPlanes = (KIND == IM2IM_LOSS_QUANTILES_L1 || KIND == IM2IM_LOSS_INN) ? 3 : 2;
    double part[3] = {0.0, 0.0, 0.0};
    const long long total = n_images * px;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long i = e / px, k = e - i * px;
        const float y = target[e];
        const long long o = i * kPlanes * px + k;
        float g[3] = {0.f, 0.f, 0.f};
        if (KIND == IM2IM_LOSS_QUANTILES_L1) {
            const float e_lo = pred[o] - y, e_mid = pred[o + px] - y, e_hi = pred[o + 2 * px] - y;
            part[0] += e_lo < 0.f ? q_lo * fabsf(e_lo) : (e_lo > 0.f ? (1.f - q_lo) * fabsf(e_lo) : 0.f);
            part[1] += e_hi < 0.f ? q_hi * fabsf(e_hi) : (e_hi > 0.f ? (1.f - q_hi) * fabsf(e_hi) : 0.f);
            part[2] += fabsf(e_mid);
            g[0] = w0 * (e_lo < 0.f ? -q_lo : (e_lo > 0.f ? 1.f - q_lo : 0.f));
            g[1] = w2 * sgnf(e_mid);
            g[2] = w1 * (e_hi < 0.f ? -q_hi : (e_hi > 0.f ? 1.f - q_hi : 0.f));
        } else if (KIND == IM2IM_LOSS_GAUSSIAN) {
            const float d = pred[o] - y, v = fmaxf(pred[o + px], 1e-6f);
            part[0] += 0.5f * (logf(v) + d * d / v);
            g[0] = w0 * d / v;
            g[1] = w0 * 0.5f * (1.f / v - d * d / (v * v));
        } else if (KIND == IM2IM_LOSS_RESIDUAL || KIND == IM2IM_LOSS_RESIDUAL_L1) {
            const float p = pred[o], r = pred[o + px];
            const float d = p - y, res = y - p, rd = r - fabsf(res);
            part[0] += (KIND == IM2IM_LOSS_RESIDUAL) ? d * d : fabsf(d);
            part[1] += rd * rd;
            g[0] = w0 * ((KIND == IM2IM_LOSS_RESIDUAL) ? 2.f * d : sgnf(d)) + w1 * 2.f * rd * sgnf(res);
            g[1] = w1 * 2.f * rd;
        } else {  // INN
            const float l = pred[o], p = pred[o + px], u = pred[o + 2 * px];
            const float d = p - y, over = fmaxf(y - u, 0.f), under = fmaxf(l - y, 0.f), wd = u - l;
            part[0] += d * d;
            part[1] += over * over + under * under + beta * fabsf(wd);
            g[0] = w1 * (2.f * under - beta * sgnf(wd));
            g[1] = w0 * 2.f * d;
            g[2] = w1 * (-2.f * over + beta * sgnf(wd));
        }
        if (dpred) {
#pragma unroll
            for (int q = 0; q < kPlanes; ++q) dpred[o + q * px] = g[q] * inv_count;
        }
    }
    __shared__ double s_part[3][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        double v = part[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_part[q][warp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int wv = 0; wv < 8; ++wv) v += s_part[threadIdx.x][wv];
        atomicAdd(&loss_parts[threadIdx.x], v);
    }
}

// torch.optim.Adam (single tensor semantics, no amsgrad / weight decay / maximize):
//   m = b1*m + (1-b1)*g;  v = b2*v + (1-b2)*g*g;  p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                                   float b1, float b2, float eps, float bc1, float bc2_sqrt,
                                                   float grad_scale, const float* __restrict__ bc_dev) {
    if (bc_dev != nullptr) { bc1 = bc_dev[0]; bc2_sqrt = bc_dev[1]; }  // step counter lives on the device (CUDA graphs)
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float gr = g[e] * grad_scale;
        const float mm = b1 * m[e] + (1.f - b1) * gr;
        const float vv = b2 * v[e] + (1.f - b2) * gr * gr;
        m[e] = mm;
        v[e] = vv;
        const float denom = sqrtf(vv) / bc2_sqrt + eps;
        p[e] -= (lr / bc1) * (mm / denom);
    }
}

// state[0] = step counter (as float, exact below 2^24), state[1] = 1 - b1^step, state[2] = sqrt(1 - b2^step)
__global__ void adam_advance_kernel(float* __restrict__ state, float b1, float b2) {
    const double step = static_cast<double>(state[0]) + 1.0;
    state[0] = static_cast<float>(step);
    state[1] = static_cast<float>(1.0 - pow(static_cast<double>(b1), step));
    state[2] = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(b2), step)));
}

// ---------------------------------------------------------------------------------------------------------------
// Head backward (QuantileRegressionLayer: n_out = 3*C_out planes from c_mid channels, 3x3 pad 1).
// (a) data gradient: dm[p, c] = sum_o sum_t dOut[o, p - shift(t)] * W[o, c, t]    (bf16 NHWC, row stride c_stride;
//     channels c_mid..c_stride-1 are written as zero so the tensor can feed the 64-channel tensor-core kernels)
template <int N_OUT>
__global__ void __launch_bounds__(128) head_dgrad_kernel(const float* __restrict__ dout, const float* __restrict__ w,
                                                         int B, int H, int W, int c_mid, int c_stride,
                                                         __nv_bfloat16* __restrict__ dm) {
    extern __shared__ float s_w[];  // [tap][o][c]
    for (int i = threadIdx.x; i < 9 * N_OUT * c_mid; i += blockDim.x) {
        const int c = i % c_mid, o = (i / c_mid) % N_OUT, t = i / (c_mid * N_OUT);
        s_w[i] = w[(static_cast<long long>(o) * c_mid + c) * 9 + t];
    }
    __syncthreads();
    const long long hw = static_cast<long long>(H) * W;
    const long long total = static_cast<long long>(B) * hw;
    for (long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < total;
         pix += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int xw = static_cast<int>(pix % W);
        const int yh = static_cast<int>((pix / W) % H);
        const long long b = pix / hw;
        float g[9][N_OUT];
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            // output pixel q = p - shift(t) received x[p] through tap t
            const int yy = yh - (t / 3 - 1), xx = xw - (t % 3 - 1);
            const bool in = yy >= 0 && yy < H && xx >= 0 && xx < W;
#pragma unroll
            for (int o = 0; o < N_OUT; ++o)
                g[t][o] = in ? __ldg(dout + (b * N_OUT + o) * hw + static_cast<long long>(yy) * W + xx) : 0.f;
        }
        __nv_bfloat16* dst = dm + pix * c_stride;
        for (int c = 0; c < c_mid; c += 8) {
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
            for (int t = 0; t < 9; ++t)
#pragma unroll
                for (int o = 0; o < N_OUT; ++o)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(g[t][o], s_w[(t * N_OUT + o) * c_mid + c + j], acc[j]);
            *reinterpret_cast<uint4*>(dst + c) = pack8(acc);
        }
        for (int c = c_mid; c < c_stride; c += 8) *reinterpret_cast<uint4*>(dst + c) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// (b) weight/bias gradient: dW[o, c, t] += sum_p dOut[o, p] * m[p + shift(t), c];  db[o] += sum_p dOut[o, p].
// Scatter form: a pixel q of m contributes m[q, c] * dOut[o, q - shift(t)] to all 9*N_OUT weights of channel c.
// lane = channel c (c_mid <= 32), one warp per image row: m is read ONCE (one 64-byte load per pixel instead of nine),
// the 3x3xN_OUT window of dOut neighbours slides along the row in registers (warp-uniform broadcast loads, four new
// columns per iteration so 4 + 4*3*N_OUT independent loads are in flight); 9*N_OUT accumulators per lane.
template <int N_OUT>
__global__ void __launch_bounds__(128) head_wgrad_kernel(const float* __restrict__ dout,
                                                         const __nv_bfloat16* __restrict__ m, int B, int H, int W,
                                                         int c_mid, int c_stride, float* __restrict__ dw,
                                                         float* __restrict__ db) {
    constexpr int kUnroll = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long hw = static_cast<long long>(H) * W;
    float acc[9][N_OUT], accb[N_OUT];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int o = 0; o < N_OUT; ++o) acc[t][o] = 0.f;
#pragma unroll
    for (int o = 0; o < N_OUT; ++o) accb[o] = 0.f;
    const long long rows = static_cast<long long>(B) * H;
    const bool ch_ok = lane < c_mid;
    for (long long row = static_cast<long long>(blockIdx.x) * 4 + warp; row < rows;
         row += static_cast<long long>(gridDim.x) * 4) {
        const long long b = row / H;
        const int y = static_cast<int>(row - b * H);
        // window row r = 0..2 <-> output rows y+1, y, y-1 (tap ky = r needs output row y + 1 - ky)
        const float* drow[N_OUT][3];
        bool rok[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const int yy = y + 1 - r;
            rok[r] = yy >= 0 && yy < H;
#pragma unroll
            for (int o = 0; o < N_OUT; ++o) drow[o][r] = dout + (b * N_OUT + o) * hw + static_cast<long long>(rok[r] ? yy : 0) * W;
        }
#pragma unroll
        for (int o = 0; o < N_OUT; ++o)
            for (int x = lane; x < W; x += 32) accb[o] += __ldg(drow[o][1] + x);
        const __nv_bfloat16* mrow = m + (b * H + y) * static_cast<long long>(W) * c_stride + lane;
        // wnd[o][r][j]: output column x - 1 + j ... (j = 0..5 covers the four pixels x..x+3 of this iteration);
        // tap kx of pixel x+u needs output column x + u + 1 - kx = wnd index u + 2 - kx
        float wnd[N_OUT][3][kUnroll + 2];
#pragma unroll
        for (int o = 0; o < N_OUT; ++o)
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                wnd[o][r][0] = 0.f;                                        // column -1
                wnd[o][r][1] = rok[r] ? __ldg(drow[o][r]) : 0.f;           // column 0
            }
        for (int x = 0; x < W; x += kUnroll) {
            float mv[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                mv[u] = (ch_ok && x + u < W) ? __bfloat162float(mrow[static_cast<long long>(x + u) * c_stride]) : 0.f;
#pragma unroll
            for (int o = 0; o < N_OUT; ++o)
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int u = 0; u < kUnroll; ++u) {
                        const int col = x + 1 + u;
                        wnd[o][r][2 + u] = (rok[r] && col < W) ? __ldg(drow[o][r] + col) : 0.f;
                    }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                        for (int o = 0; o < N_OUT; ++o)
                            acc[ky * 3 + kx][o] = fmaf(wnd[o][ky][u + 2 - kx], mv[u], acc[ky * 3 + kx][o]);
#pragma unroll
            for (int o = 0; o < N_OUT; ++o)
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    wnd[o][r][0] = wnd[o][r][kUnroll];
                    wnd[o][r][1] = wnd[o][r][kUnroll + 1];
                }
        }
    }
#pragma unroll
    for (int o = 0; o < N_OUT; ++o)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) accb[o] += __shfl_xor_sync(0xffffffffu, accb[o], off);
    __shared__ float s_acc[4][9 * N_OUT][32 + 1];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int o = 0; o < N_OUT; ++o) s_acc[warp][t * N_OUT + o][lane] = acc[t][o];
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * N_OUT * 32; i += 128) {
        const int c = i % 32, to = i / 32;
        const int t = to / N_OUT, o = to % N_OUT;
        float s = 0.f;
        for (int wv = 0; wv < 4; ++wv) s += s_acc[wv][to][c];
        if (c < c_mid) atomicAdd(&dw[(static_cast<long long>(o) * c_mid + c) * 9 + t], s);
    }
    if (lane == 0)
#pragma unroll
        for (int o = 0; o < N_OUT; ++o) atomicAdd(&db[o], accb[o]);
}

// First convolution's weight gradient: dW[co, ci, t] += sum_p dz[p, co] * x[b, ci, p + shift(t)]   (x fp32 NCHW).
// thread = (pixel stream, group of 8 output channels): one 16-byte load of dz feeds 8*9*c_in FMAs; the streams of a
// block are reduced in shared memory so that each block issues one atomicAdd per weight.  c_in == 1 is specialised
// (accumulators in registers); larger c_in loops over the input channels.
__global__ void __launch_bounds__(256) conv_first_wgrad_kernel(const float* __restrict__ x,
                                                               const __nv_bfloat16* __restrict__ dz, int B, int c_in,
                                                               int H, int W, int c_out, float* __restrict__ dw) {
    extern __shared__ float s_red[];  // [streams][c_out * 9]
    const int groups = c_out / 8;
    const int g = threadIdx.x % groups;
    const int stream = threadIdx.x / groups;
    const int n_streams = blockDim.x / groups;
    const long long hw = static_cast<long long>(H) * W;
    const long long total = static_cast<long long>(B) * hw;
    for (int ci = 0; ci < c_in; ++ci) {
        float acc[8][9];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[j][t] = 0.f;
        // two horizontally adjacent pixels per iteration: one 3x4 window of x (12 loads) and two 16-byte loads of dz are
        // issued before the first FMA, and feed 2 x 72 FMAs
        const int wp = (W + 1) / 2;
        const long long pairs = static_cast<long long>(B) * H * wp;
        for (long long pp = static_cast<long long>(blockIdx.x) * n_streams + stream; pp < pairs;
             pp += static_cast<long long>(gridDim.x) * n_streams) {
            const int xw = 2 * static_cast<int>(pp % wp);
            const int yh = static_cast<int>((pp / wp) % H);
            const long long b = pp / (static_cast<long long>(wp) * H);
            const long long pix = (b * H + yh) * W + xw;
            const bool second = xw + 1 < W;
            Bf16x8 vd0, vd1;
            vd0.u = *reinterpret_cast<const uint4*>(dz + pix * c_out + g * 8);
            vd1.u = second ? *reinterpret_cast<const uint4*>(dz + (pix + 1) * c_out + g * 8) : make_uint4(0u, 0u, 0u, 0u);
            const float* xp = x + (b * c_in + ci) * hw;
            float v[3][4];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int yy = yh + r - 1, xx = xw + c - 1;
                    v[r][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(xp + static_cast<long long>(yy) * W + xx) : 0.f;
                }
            float d0[8], d1[8];
            unpack8(vd0, d0);
            unpack8(vd1, d1);
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int t = 0; t < 9; ++t)
                    acc[j][t] = fmaf(d1[j], v[t / 3][t % 3 + 1], fmaf(d0[j], v[t / 3][t % 3], acc[j][t]));
        }
        // block reduction over the pixel streams
        float* mine = s_red + static_cast<size_t>(stream) * c_out * 9;
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
            for (int t = 0; t < 9; ++t) mine[(g * 8 + j) * 9 + t] = acc[j][t];
        __syncthreads();
        for (int i = threadIdx.x; i < c_out * 9; i += blockDim.x) {
            float s = 0.f;
            for (int st = 0; st < n_streams; ++st) s += s_red[static_cast<size_t>(st) * c_out * 9 + i];
            const int co = i / 9, t = i % 9;
            atomicAdd(&dw[(static_cast<long long>(co) * c_in + ci) * 9 + t], s);
        }
        __syncthreads();
    }
}

}  // namespace
}  // namespace im2im

using namespace im2im;

#define ST(stream) static_cast<cudaStream_t>(stream)
#define BF(p) static_cast<const __nv_bfloat16*>(p)
#define BFW(p) static_cast<__nv_bfloat16*>(p)

static int check_channels(int C, const char* what) {
    if (C <= 0 || C % 8 || 256 % (C / 8) || C > 2048 / 8 * 8) return fail(IM2IM_ERANGE, "%s: C=%d unsupported", what, C);
    return IM2IM_OK;
}

// dynamic shared memory of channel_reduce_kernel: 2*C floats are used; IM2IM_REDUCE_EXCLUSIVE=1 pads the request to 18 KB, which
// keeps the kernel from sharing an SM with the tensor-core kernels (A/B switch for the co-residency experiment)
static size_t reduce_smem_bytes(int C) {
    static const bool exclusive = (getenv("IM2IM_REDUCE_EXCLUSIVE") != nullptr);
    const size_t need = sizeof(float) * 2 * C;
    return (exclusive && need < 18432) ? 18432 : need;
}

extern "C" int im2im_channel_stats_bf16(const void* d_z, int64_t n_pix, int32_t C, float* d_sums, void* stream) {
    if (int rc = check_channels(C, "channel_stats")) return rc;
    if (n_pix <= 0 || !d_z || !d_sums) return fail(IM2IM_EINVAL, "channel_stats: bad arguments");
    const int lanes = 256 / (C / 8);
    channel_reduce_kernel<0><<<grid_for((n_pix + 3) / 4, lanes, 3), 256, reduce_smem_bytes(C), ST(stream)>>>(BF(d_z), nullptr, nullptr, nullptr, nullptr,
                                                                               nullptr, n_pix, C, d_sums);
    return check_launch("channel_reduce_kernel<stats>");
}

extern "C" int im2im_bn_finalize(const float* d_sums, int64_t count, const float* d_conv_bias, const float* d_gamma,
                                 const float* d_beta, float eps, float momentum, int32_t C, float* d_running_mean,
                                 float* d_running_var, float* d_scale, float* d_shift, float* d_save_mean,
                                 float* d_save_rstd, void* stream) {
    if (C <= 0 || count <= 0 || !d_sums || !d_gamma || !d_beta || !d_scale || !d_shift || !d_save_mean || !d_save_rstd)
        return fail(IM2IM_EINVAL, "bn_finalize: bad arguments");
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(d_sums, static_cast<double>(count), d_conv_bias, d_gamma,
                                                                d_beta, eps, momentum, C, d_running_mean, d_running_var,
                                                                d_scale, d_shift, d_save_mean, d_save_rstd);
    return check_launch("bn_finalize_kernel");
}

extern "C" int im2im_bn_apply_relu_bf16(const void* d_z, const float* d_scale, const float* d_shift, int64_t n_pix,
                                        int32_t C, void* d_y, void* stream) {
    if (int rc = check_channels(C, "bn_apply")) return rc;   // also guarantees 256 % (C/8) == 0 (loop-invariant channel group)
    if (n_pix <= 0 || !d_z || !d_scale || !d_shift || !d_y) return fail(IM2IM_EINVAL, "bn_apply: bad arguments");
    bn_apply_relu_kernel<<<grid_for(n_pix * (C / 8), 256, 16), 256, 0, ST(stream)>>>(BF(d_z), d_scale, d_shift, n_pix, C,
                                                                                    BFW(d_y));
    return check_launch("bn_apply_relu_kernel");
}

extern "C" int im2im_bn_relu_bwd_bf16(const void* d_dy, const void* d_z, const float* d_gamma, const float* d_beta,
                                      const float* d_mean, const float* d_rstd, int64_t n_pix, int32_t C,
                                      float* d_sums, void* d_dz, void* stream) {
    if (int rc = check_channels(C, "bn_relu_bwd")) return rc;
    if (n_pix <= 0 || !d_dy || !d_z || !d_gamma || !d_beta || !d_mean || !d_rstd || !d_sums || !d_dz)
        return fail(IM2IM_EINVAL, "bn_relu_bwd: bad arguments");
    IM2IM_CUDA_TRY(cudaMemsetAsync(d_sums, 0, sizeof(float) * 2 * C, ST(stream)));
    const int lanes = 256 / (C / 8);
    channel_reduce_kernel<1><<<grid_for((n_pix + 3) / 4, lanes, 3), 256, reduce_smem_bytes(C), ST(stream)>>>(BF(d_dy), BF(d_z), d_gamma, d_beta, d_mean,
                                                                               d_rstd, n_pix, C, d_sums);
    if (int rc = check_launch("channel_reduce_kernel<bn_bwd>")) return rc;
    bn_relu_bwd_apply_kernel<<<grid_for(n_pix * (C / 8), 256, 12), 256, 0, ST(stream)>>>(
        BF(d_dy), BF(d_z), d_gamma, d_beta, d_mean, d_rstd, d_sums, 1.f / static_cast<float>(n_pix), n_pix, C, 0, BFW(d_dz));
    return check_launch("bn_relu_bwd_apply_kernel");
}

extern "C" int im2im_bn_relu_bwd_apply_bf16(const void* d_g, const void* d_z, const float* d_gamma, const float* d_beta,
                                            const float* d_mean, const float* d_rstd, const float* d_sums, int64_t n_pix,
                                            int32_t C, int32_t premasked, void* d_dz, void* stream) {
    if (int rc = check_channels(C, "bn_relu_bwd_apply")) return rc;
    if (n_pix <= 0 || !d_g || !d_z || !d_gamma || !d_beta || !d_mean || !d_rstd || !d_sums || !d_dz)
        return fail(IM2IM_EINVAL, "bn_relu_bwd_apply: bad arguments");
    bn_relu_bwd_apply_kernel<<<grid_for(n_pix * (C / 8), 256, 12), 256, 0, ST(stream)>>>(
        BF(d_g), BF(d_z), d_gamma, d_beta, d_mean, d_rstd, d_sums, 1.f / static_cast<float>(n_pix), n_pix, C,
        premasked ? 1 : 0, BFW(d_dz));
    return check_launch("bn_relu_bwd_apply_kernel");
}

extern "C" int im2im_bn_apply_relu_pool_bf16(const void* d_z, const float* d_scale, const float* d_shift, int32_t B, int32_t H,
                                            int32_t W, int32_t C, void* d_y, void* d_pooled, void* stream) {
    if (int rc = check_channels(C, "bn_apply_pool")) return rc;
    if (B <= 0 || H < 2 || W < 2 || (H % 2) || (W % 2)) return fail(IM2IM_ENOTSUP, "bn_apply_pool: needs even H and W (got %dx%d)", H, W);
    if (!d_z || !d_scale || !d_shift || !d_y || !d_pooled) return fail(IM2IM_EINVAL, "bn_apply_pool: bad arguments");
    if (B > 65535 || H / 2 > 65535) return fail(IM2IM_ERANGE, "bn_apply_pool: B and H/2 must be <= 65535");
    const dim3 grid(static_cast<unsigned>(((W / 2) * (C / 8) + 255) / 256), static_cast<unsigned>(H / 2), static_cast<unsigned>(B));
    bn_apply_relu_pool_kernel<<<grid, 256, 0, ST(stream)>>>(BF(d_z), d_scale, d_shift, B, H, W, C, BFW(d_y), BFW(d_pooled));
    return check_launch("bn_apply_relu_pool_kernel");
}

extern "C" int im2im_bn_relu_pool_bwd_bf16(const void* d_dskip, const void* d_dpooled, const void* d_z, const float* d_gamma,
                                           const float* d_beta, const float* d_mean, const float* d_rstd, int32_t B,
                                           int32_t H, int32_t W, int32_t C, float* d_sums, void* d_dz, void* stream) {
    if (int rc = check_channels(C, "bn_relu_pool_bwd")) return rc;
    if (B <= 0 || H < 2 || W < 2 || (H % 2) || (W % 2)) return fail(IM2IM_ENOTSUP, "bn_relu_pool_bwd: needs even H and W");
    if (!d_dskip || !d_dpooled || !d_z || !d_gamma || !d_beta || !d_mean || !d_rstd || !d_sums || !d_dz)
        return fail(IM2IM_EINVAL, "bn_relu_pool_bwd: bad arguments");
    IM2IM_CUDA_TRY(cudaMemsetAsync(d_sums, 0, sizeof(float) * 2 * C, ST(stream)));
    const long long n_win = static_cast<long long>(B) * (H / 2) * (W / 2);
    const long long n_pix = static_cast<long long>(B) * H * W;
    const int lanes = 256 / (C / 8);
    const unsigned grid = grid_for(n_win, lanes, 2);
    bn_pool_bwd_kernel<0><<<grid, 256, sizeof(float) * 2 * C, ST(stream)>>>(BF(d_dskip), BF(d_dpooled), BF(d_z), d_gamma, d_beta,
                                                                            d_mean, d_rstd, d_sums, 0.f, B, H, W, C, nullptr);
    if (int rc = check_launch("bn_pool_bwd_kernel<reduce>")) return rc;
    bn_pool_bwd_kernel<1><<<grid_for(n_win, lanes, 8), 256, 0, ST(stream)>>>(BF(d_dskip), BF(d_dpooled), BF(d_z), d_gamma, d_beta,
                                                                             d_mean, d_rstd, d_sums, 1.f / static_cast<float>(n_pix),
                                                                             B, H, W, C, BFW(d_dz));
    return check_launch("bn_pool_bwd_kernel<apply>");
}

extern "C" int im2im_maxpool2x2_bwd_bf16(const void* d_x, const void* d_dy, int32_t B, int32_t H, int32_t W, int32_t C,
                                         int32_t accumulate, void* d_dx, void* stream) {
    if (B <= 0 || H < 2 || W < 2 || C <= 0 || C % 8 || !d_x || !d_dy || !d_dx) return fail(IM2IM_EINVAL, "maxpool_bwd: bad arguments");
    if ((H % 2 || W % 2) && !accumulate)  // trailing odd row/column never reaches the pool: its gradient is zero
        IM2IM_CUDA_TRY(cudaMemsetAsync(d_dx, 0, sizeof(__nv_bfloat16) * static_cast<size_t>(B) * H * W * C, ST(stream)));
    if (B > 65535 || H / 2 > 65535) return fail(IM2IM_ERANGE, "maxpool_bwd: B and H/2 must be <= 65535");
    const dim3 pgrid(static_cast<unsigned>(((W / 2) * (C / 8) + 255) / 256), static_cast<unsigned>(H / 2), static_cast<unsigned>(B));
    maxpool2x2_bwd_kernel<<<pgrid, 256, 0, ST(stream)>>>(BF(d_x), BF(d_dy), B, H, W, C, accumulate,
                                                                           BFW(d_dx));
    return check_launch("maxpool2x2_bwd_kernel");
}

extern "C" int im2im_upsample2x_bilinear_bwd_bf16(const void* d_du, int32_t B, int32_t h, int32_t w, int32_t C,
                                                  int32_t H_out, int32_t W_out, void* d_dx, void* stream) {
    if (B <= 0 || h <= 0 || w <= 0 || C <= 0 || C % 8 || H_out < 2 * h || W_out < 2 * w || !d_du || !d_dx)
        return fail(IM2IM_EINVAL, "upsample_bwd: bad arguments");
    const int pad_top = (H_out - 2 * h) / 2, pad_left = (W_out - 2 * w) / 2;
    if (B > 65535 || h > 65535) return fail(IM2IM_ERANGE, "upsample_bwd: B and h must be <= 65535");
    if (C % 16 == 0) {
        const dim3 ugrid(static_cast<unsigned>((w * (C / 16) + 255) / 256), static_cast<unsigned>(h), static_cast<unsigned>(B));
        upsample2x_bwd_kernel<2><<<ugrid, 256, 0, ST(stream)>>>(BF(d_du), B, h, w, C, H_out, W_out, pad_top, pad_left, BFW(d_dx));
    } else {
        const dim3 ugrid(static_cast<unsigned>((w * (C / 8) + 255) / 256), static_cast<unsigned>(h), static_cast<unsigned>(B));
        upsample2x_bwd_kernel<1><<<ugrid, 256, 0, ST(stream)>>>(BF(d_du), B, h, w, C, H_out, W_out, pad_top, pad_left, BFW(d_dx));
    }
    return check_launch("upsample2x_bwd_kernel");
}

extern "C" int im2im_quantile_loss_f32(const float* d_pred, const float* d_target, int64_t n_images, int64_t px,
                                       float q_lo, float q_hi, float w_lo, float w_hi, float w_mse, float* d_dpred,
                                       double* d_loss_parts, void* stream) {
    if (n_images <= 0 || px <= 0 || !d_pred || !d_target || !d_loss_parts) return fail(IM2IM_EINVAL, "quantile_loss: bad arguments");
    IM2IM_CUDA_TRY(cudaMemsetAsync(d_loss_parts, 0, sizeof(double) * 3, ST(stream)));
    const long long n = n_images * px;
    quantile_loss_kernel<<<grid_for(n, 256, 8), 256, 0, ST(stream)>>>(d_pred, d_target, n_images, px, q_lo, q_hi, w_lo,
                                                                     w_hi, w_mse, 1.f / static_cast<float>(n), d_dpred,
                                                                     d_loss_parts);
    return check_launch("quantile_loss_kernel");
}

extern "C" int im2im_head_loss_f32(int32_t loss_kind, const float* d_pred, const float* d_target, int64_t n_images,
                                   int64_t px, float q_lo, float q_hi, float w0, float w1, float w2, float beta,
                                   float* d_dpred, double* d_loss_parts, void* stream) {
    if (n_images <= 0 || px <= 0 || !d_pred || !d_target || !d_loss_parts) return fail(IM2IM_EINVAL, "head_loss: bad arguments");
    if (loss_kind == IM2IM_LOSS_QUANTILES)
        return im2im_quantile_loss_f32(d_pred, d_target, n_images, px, q_lo, q_hi, w0, w1, w2, d_dpred, d_loss_parts, stream);
    IM2IM_CUDA_TRY(cudaMemsetAsync(d_loss_parts, 0, sizeof(double) * 3, ST(stream)));
    const long long n = n_images * px;
    const unsigned grid = grid_for(n, 256, 8);
    const float inv = 1.f / static_cast<float>(n);
#define IM2IM_LOSS_CASE(K)                                                                                          \
    case K:                                                                                                         \
        head_loss_kernel<K><<<grid, 256, 0, ST(stream)>>>(d_pred, d_target, n_images, px, q_lo, q_hi, w0, w1, w2, beta, \
                                                          inv, d_dpred, d_loss_parts);                              \
        break
    switch (loss_kind) {
        IM2IM_LOSS_CASE(IM2IM_LOSS_QUANTILES_L1); IM2IM_LOSS_CASE(IM2IM_LOSS_GAUSSIAN); IM2IM_LOSS_CASE(IM2IM_LOSS_RESIDUAL);
        IM2IM_LOSS_CASE(IM2IM_LOSS_RESIDUAL_L1); IM2IM_LOSS_CASE(IM2IM_LOSS_INN);
        default: return fail(IM2IM_ENOTSUP, "head_loss: loss_kind=%d", loss_kind);
    }
#undef IM2IM_LOSS_CASE
    return check_launch("head_loss_kernel");
}

extern "C" int im2im_adam_step_f32(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq,
                                   int64_t n, float lr, float beta1, float beta2, float eps, int32_t step,
                                   float grad_scale, void* stream) {
    if (n < 0 || step < 1) return fail(IM2IM_EINVAL, "adam: bad arguments");
    if (n == 0) return IM2IM_OK;
    if (!d_param || !d_grad || !d_exp_avg || !d_exp_avg_sq) return fail(IM2IM_EINVAL, "adam: null tensor");
    const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
    const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
    adam_kernel<<<grid_for(n, 256, 16), 256, 0, ST(stream)>>>(d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, lr, beta1,
                                                             beta2, eps, static_cast<float>(bc1),
                                                             static_cast<float>(sqrt(bc2)), grad_scale, nullptr);
    return check_launch("adam_kernel");
}

extern "C" int im2im_adam_step_dev_f32(float* d_param, const float* d_grad, float* d_exp_avg, float* d_exp_avg_sq,
                                       int64_t n, float lr, float beta1, float beta2, float eps, float* d_state,
                                       float grad_scale, void* stream) {
    if (n < 0 || !d_state) return fail(IM2IM_EINVAL, "adam_dev: bad arguments");
    if (n > 0 && (!d_param || !d_grad || !d_exp_avg || !d_exp_avg_sq)) return fail(IM2IM_EINVAL, "adam_dev: null tensor");
    adam_advance_kernel<<<1, 1, 0, ST(stream)>>>(d_state, beta1, beta2);
    if (int rc = check_launch("adam_advance_kernel")) return rc;
    if (n == 0) return IM2IM_OK;
    adam_kernel<<<grid_for(n, 256, 16), 256, 0, ST(stream)>>>(d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, lr, beta1,
                                                             beta2, eps, 1.f, 1.f, grad_scale, d_state + 1);
    return check_launch("adam_kernel");
}

extern "C" int im2im_head_bwd(const float* d_dout, const void* d_m, const float* d_weight, int32_t B, int32_t H,
                              int32_t W, int32_t c_mid, int32_t c_stride, int32_t n_out, void* d_dm, float* d_dw,
                              float* d_db, void* stream) {
    if (B <= 0 || H <= 0 || W <= 0 || c_mid <= 0 || c_mid > 32 || c_mid % 8 || c_stride < c_mid || c_stride % 8)
        return fail(IM2IM_ERANGE, "head_bwd: bad shape (c_mid=%d c_stride=%d)", c_mid, c_stride);
    if (!d_dout || !d_m || !d_weight || !d_dm || !d_dw || !d_db) return fail(IM2IM_EINVAL, "head_bwd: null tensor");
    const long long pixels = static_cast<long long>(B) * H * W;
    const size_t smem = sizeof(float) * 9 * n_out * c_mid;
    const unsigned g1 = grid_for(pixels, 128, 8);
    const long long row_blocks = (static_cast<long long>(B) * H + 3) / 4;
    const unsigned g2 = static_cast<unsigned>(row_blocks < 8ll * sm_count() ? row_blocks : 8ll * sm_count());
#define IM2IM_HEAD_BWD_CASE(N)                                                                                        \
    case N:                                                                                                          \
        head_dgrad_kernel<N><<<g1, 128, smem, ST(stream)>>>(d_dout, d_weight, B, H, W, c_mid, c_stride, BFW(d_dm));  \
        if (int rc = check_launch("head_dgrad_kernel")) return rc;                                                   \
        head_wgrad_kernel<N><<<g2, 128, 0, ST(stream)>>>(d_dout, BF(d_m), B, H, W, c_mid, c_stride, d_dw, d_db);    \
        break
    switch (n_out) {
        IM2IM_HEAD_BWD_CASE(2); IM2IM_HEAD_BWD_CASE(3); IM2IM_HEAD_BWD_CASE(4); IM2IM_HEAD_BWD_CASE(6);
        default:
            return fail(IM2IM_ENOTSUP, "head_bwd: n_out=%d (2, 3, 4 or 6 supported)", n_out);
    }
#undef IM2IM_HEAD_BWD_CASE
    return check_launch("head_wgrad_kernel");
}

extern "C" int im2im_conv_first_wgrad(const float* d_x, const void* d_dz, int32_t B, int32_t c_in, int32_t H, int32_t W,
                                      int32_t c_out, float* d_dw, void* stream) {
    if (B <= 0 || H <= 0 || W <= 0 || c_in <= 0 || c_in > 8 || c_out <= 0 || c_out % 8 || 256 % (c_out / 8))
        return fail(IM2IM_ERANGE, "conv_first_wgrad: bad shape (c_in=%d c_out=%d)", c_in, c_out);
    if (!d_x || !d_dz || !d_dw) return fail(IM2IM_EINVAL, "conv_first_wgrad: null tensor");
    const int n_streams = 256 / (c_out / 8);
    const size_t smem = sizeof(float) * n_streams * c_out * 9;
    if (smem > 200 * 1024) return fail(IM2IM_ERANGE, "conv_first_wgrad: c_out=%d too large", c_out);
    IM2IM_CUDA_TRY(cudaFuncSetAttribute(conv_first_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_first_wgrad_kernel<<<static_cast<unsigned>(2 * sm_count()), 256, smem, ST(stream)>>>(d_x, BF(d_dz), B, c_in, H,
                                                                                             W, c_out, d_dw);
    return check_launch("conv_first_wgrad_kernel");
}
