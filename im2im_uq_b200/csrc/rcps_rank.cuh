// Per-pixel rank of the RCPS miss predicate (shared by the streaming calibration kernel in rcps_kernels.cu and the
// head-fused histogram epilogue of conv_halo_kernel in conv_kernels.cu, so both book a pixel into the same bin).
#pragma once
#include "common.cuh"

namespace im2im {
namespace {

// ---------------------------------------------------------------------------------------------------------------
// Per-pixel rank.  Exactness notes (all fp32, round-to-nearest, no contraction):
//   pm = p - 1e-6f, pp = p + 1e-6f
//   reference upper miss: max(lam*(max(u,pp)-p)+p, pp) < y   <=>  (pp < y) && (lam*du + p < y)   [NaN -> false]
//   reference lower miss: min(p-lam*(p-min(l,pm)), pm) > y   <=>  (pm > y) && (p - lam*dl > y)
//   pm <= p <= pp, so only the upper side can miss when y > p and only the lower side when y < p.
//   The lower side is folded onto the upper-side form by negating (l,p,y) -> (U,P,Y) = (-l,-p,-y): negation is
//   exact and rounding is symmetric, so  P+1e-6 = -pm,  max(U,PP)-P = p-min(l,pm) = dl  and
//   p - lam*dl > y  <=>  lam*dl + P < Y  bit for bit.
//   torch.minimum/maximum propagate NaN: a NaN u (resp. l) makes that side's predicate false for every lambda,
//   a NaN p or y makes every comparison false - both are covered by `active`.
struct PixelQuery {
    float d, P, Y;
    bool active;
};

// torch.relu semantics (NaN propagates; clamp_min)
__device__ __forceinline__ float t_relu(float x) { return (x != x) ? x : fmaxf(x, 0.f); }

// Head kinds (include/im2im_uq.h).  Every head of the reference has the same shape after the outer clamp of
// add_uncertainty.py:35-36:  upper = max(fl(fl(lam*d_up) + p), p+1e-6),  lower = min(fl(p - fl(lam*d_lo)), p-1e-6)
// (fl(fl(-lam*d)+p) == fl(p - fl(lam*d)) bit for bit), only the widths differ:
//   QUANTILES      (a,p,b) = (lower, pred, upper):  d_up = max(b, p+1e-6) - p,  d_lo = p - min(a, p-1e-6)
//                  quantile_layer.py:39-42, quantile_l1_layer.py:39-42, inn_layer.py:35-38
//   RESIDUAL       (p,b) = (pred, |residual|):      d_up = d_lo = b                residual_magnitude_layer.py:33-34
//   GAUSSIAN       (p,b) = (mean, variance):        d_up = d_lo = sqrt(b)          gaussian_layer.py:31-32
//   SOFTMAX_SETS   (a,p,b) = (lower quantile, argmax, upper quantile): d_up = relu(b-p), d_lo = relu(p-a)
//                  softmax_layer.py:50-51
template <int HEAD>
__device__ __forceinline__ PixelQuery make_query(float a, float p, float b, float y) {
    const bool up = y > p;
    PixelQuery q;
    q.P = up ? p : -p;
    q.Y = up ? y : -y;
    const float PP = __fadd_rn(q.P, 1e-6f);
    if (HEAD == IM2IM_HEAD_QUANTILES) {
        const float U = up ? b : -a;
        q.active = (PP < q.Y) && (U == U);
        q.d = __fsub_rn(fmaxf(U, PP), q.P);  // du or dl, >= 0 (NaN only when inactive)
    } else if (HEAD == IM2IM_HEAD_SOFTMAX_SETS) {
        const float U = up ? b : -a;
        q.d = t_relu(__fsub_rn(U, q.P));
        q.active = (PP < q.Y) && (q.d == q.d);
    } else {
        q.d = (HEAD == IM2IM_HEAD_GAUSSIAN) ? __fsqrt_rn(b) : b;  // torch.sqrt is correctly rounded; sqrt(<0) = NaN
        q.active = (PP < q.Y) && (q.d == q.d);
    }
    return q;
}

__device__ __forceinline__ bool missed(const PixelQuery& q, float lam) {
    return __fadd_rn(__fmul_rn(lam, q.d), q.P) < q.Y;
}

// exact binary search for the first covered lambda (rare: guess off by more than the checked window)
__device__ __noinline__ int rank_bisect(float d, float P, float Y, const float* s_lam, int L) {
    int lo = 0, hi = L;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__fadd_rn(__fmul_rn(s_lam[mid], d), P) < Y) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Negative width (only reachable with the RESIDUAL head when a caller hands in a width plane that did not go through
// abs()): the predicate is non-DEcreasing in lambda, i.e. true on a suffix of the grid.  Returns the first missed index.
__device__ __noinline__ int rank_bisect_rising(float d, float P, float Y, const float* s_lam, int L) {
    int lo = 0, hi = L;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__fadd_rn(__fmul_rn(s_lam[mid], d), P) < Y) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Number of grid points at which the pixel is missed = index of the first lambda that covers it.
// guess_scale/guess_bias map the real-valued crossing lam* = (Y-P)/d onto the (uniform) grid:
// #{lam_j < lam*} = ceil((lam*-lam0)/dlam).  The guess g is then verified with the exact predicate at g-1 and g;
// s_pair[g] = (lam[g-1], lam[g]) (with -inf / +inf sentinels at the ends) so both neighbours come from one 64-bit
// shared load.
// Returns the guess and sets `ok` when the verification passed (or the pixel can never miss); branch-free so that
// the pixels of a thread interleave.  A failed verification is resolved by rank_bisect (rare).
template <int HEAD>
__device__ __forceinline__ int rank_guess(const PixelQuery& q, const float2* __restrict__ s_pair, int L,
                                          float guess_scale, float guess_bias, bool& ok) {
    float rcp;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(q.d));
    const float t = (q.Y - q.P) * rcp;
    // The conversion is done in opaque PTX on purpose: cvt.rpi.s32.f32 saturates and maps NaN to 0 by specification,
    // whereas a C++ float->int conversion of NaN (inf*0 for p = +/-inf) is undefined and NVVM was seen to fold the
    // later `g == L` test into a float compare that is true for NaN.
    int g;
    asm("cvt.rpi.s32.f32 %0, %1;" : "=r"(g) : "f"(fmaf(t, guess_scale, guess_bias)));
    g = max(0, min(g, L));
    // s_pair carries sentinels (-inf below the grid, +inf above it), so g == 0 and g == L need no special case:
    // lam = -inf is "missed" and lam = +inf is "covered" for every active pixel with a positive width.  (A zero
    // width gives NaN at the -inf sentinel, fails the check and is resolved by the bisection - still exact.)
    // Testing `g == L` on the clamped value directly is avoided on purpose: ptxas 12.9 lowers clamp+compare to a
    // VIMNMX.RELU predicate output that was observed to be true for g == 0 as well.
    const float2 nb = s_pair[g];
    const bool below_ok = missed(q, nb.x);   // missed at every grid point below g
    const bool above_ok = !missed(q, nb.y);  // covered from g upwards
    ok = !q.active || (below_ok && above_ok);
    if (HEAD == IM2IM_HEAD_RESIDUAL) ok = ok && !(q.active && q.d < 0.f);  // negative width: resolve_slow
    return q.active ? g : 0;
}

// Slow path for a pixel whose guess failed verification.  Returns the rank for the falling histogram; a pixel with a
// negative width is booked into the rising histogram instead (rise[k] = #pixels missed from index k upwards).
template <int HEAD>
__device__ __forceinline__ int resolve_slow(const PixelQuery& q, const float* s_lam, int L, unsigned* rise) {
    if (HEAD == IM2IM_HEAD_RESIDUAL && q.d < 0.f) {
        const int k2 = rank_bisect_rising(q.d, q.P, q.Y, s_lam, L);
        if (k2 < L) { atomicAdd(&rise[k2], 1u); atomicAdd(&rise[L], 1u); }  // rise[L] = number of booked pixels
        return 0;
    }
    return rank_bisect(q.d, q.P, q.Y, s_lam, L);
}

}  // namespace
}  // namespace im2im
