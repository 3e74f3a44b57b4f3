// RCPS conformal-calibration kernels for sm_100a (Path A of DESIGN.md).
//
// One pass over the scores replaces the reference's one-pass-per-lambda sweep
// (core/calibration/calibrate_model.py:134-136 in the reference tree).  Per pixel the fp32 miss predicate of
//   quantile_layer.py:39-42 -> add_uncertainty.py:35-36 -> calibrate_model.py:77-78
// is monotone non-increasing in lambda (every op is a correctly rounded monotone function of lambda for a
// non-negative width), so a pixel is summarised by k = #{j : missed at lambda_j} and
//   counts[i, j] = #{pixels of image i with k > j}.
// k is found from an arithmetic guess and then VERIFIED with the exact predicate (same op order, no FMA), so the
// counts are bit-identical to evaluating every lambda separately.
//
// HBM-bound design: persistent CTAs (one per SM); a producer warp streams 4-plane tiles into a shared-memory ring
// with cp.async.bulk (UBLKCP) + mbarrier transaction counts; 16 consumer warps read the tile with LDS.128, rank the
// pixels and bump a per-image shared-memory histogram; a block-wide suffix scan turns the histogram into the
// image's row of counts.
#include <ctime>

#include "common.cuh"
#include "rcps_rank.cuh"

namespace im2im {
namespace {

constexpr int kConsumerWarps = 16;
constexpr int kConsumerThreads = kConsumerWarps * 32;  // 512
constexpr int kThreads = kConsumerThreads + 32;        // + one producer warp
constexpr int kPxPerThread = 4;
constexpr int kTilePx = kConsumerThreads * kPxPerThread;  // 2048 pixels of each of the 4 planes per stage (32 KB)
constexpr int kStages = 4;                                // 128 KB in flight per SM
constexpr int kFlushBarrier = 1;                          // named barrier id for the consumer warps
constexpr size_t kMaxSmemOptin = 232448;                  // 227 KB per CTA on sm_100

struct RcpsParams {
    const float* plane[4];  // lower, pred, upper, label
    long long stride[4];    // elements between consecutive images of each plane
    long long n_images;
    long long px;           // C*H*W values per image
    long long tiles_per_image;
    long long total_tiles;
    const float* lambdas;   // device, sorted ascending
    int n_lambdas;
    int* counts;                  // [n_images, n_lambdas]
    unsigned long long* totals;   // [n_lambdas] or null
};

// Stop-rule screening constants (host twin: calibration/sweep.py::screening_constants)
struct ScreenParams {
    double n_px, gamma, alpha32, r_lo, r_hi, slack;
};

// Everything the fused single-launch calibration needs on top of RcpsParams (rcps_hist_kernel<true, HEAD, true>).
// Workspace invariant: totals_acc[0..L) == 0 and ticket == 0 between launches (the last CTA restores it).
struct FusedParams {
    float* table;                       // [n_images, n_lambdas] fp32 loss table or null
    unsigned long long* totals_acc;     // workspace u64[L]
    unsigned* ticket;                   // workspace: CTAs finished so far
    unsigned* go;                       // workspace u32[2]: {epoch of the published decision, first visited column}
    unsigned* err;                      // workspace u32: set when an intra-GPU wait gave up (a bug or a lost block, never a peer)
    unsigned long long* stamps;         // workspace u64[8]: globaltimer (ns) at {first CTA start, last CTA's ticket, totals read,
                                        //   pushed + fenced, all peers arrived, decision published, -, -} of the last launch
    unsigned* head_flag;                // workspace u32[grid]: epoch at which CTA b published its head-partial row
    int* head_partial;                  // workspace i32[grid][L]: miss counts of the image a CTA starts in the middle of
    unsigned* epoch;                    // device-side launch counter (incremented by the kernel: graph replays need no new args)
    ScreenParams scr;
    unsigned long long* const* peer_mailbox;   // multi-GPU (world > 1): see rcps_decide_p2p_kernel
    unsigned* const* peer_flags;
    int rank, world;
    long long timeout_cycles;           // peer wait limit in clock64 cycles, 0 = wait for ever
    unsigned long long* totals_out;     // [L] totals summed over ranks
    int* result;                        // device int32[4], as rcps_decide_kernel
    volatile int* result_host;          // mapped pinned int32[8] or null: {result[0..3], epoch tag}
};


// ---------------------------------------------------------------------------------------------------------------
// hist[1..L] -> counts row (suffix sums), accumulate per-CTA totals, zero the histogram.  Consumer threads only.
//   kFlushAtomic     row shared with another CTA, pre-zeroed: atomicAdd of the non-zero entries
//   kFlushSparse     row owned by this CTA, pre-zeroed: plain stores of the non-zero entries
//   kFlushFull       row owned by this CTA, NOT pre-zeroed: every entry is stored (fused path: no memset launch)
//   kFlushFullPlus   as kFlushFull, plus `addend[j]` (the part of the image the next CTA counted)
enum { kFlushAtomic = 0, kFlushSparse = 1, kFlushFull = 2, kFlushFullPlus = 3 };

__device__ __forceinline__ void flush_image(unsigned* hist, unsigned* rise, unsigned long long* tot,
                                            unsigned* warp_sums, int L, int* counts_row, int mode,
                                            const int* addend, int ctid, float* table_row = nullptr, float fpx = 1.f) {
    named_bar_sync(kFlushBarrier, kConsumerThreads);  // all histogram atomics of this image have landed
    const int per_thread = (L + kConsumerThreads - 1) / kConsumerThreads;
    const int r0 = ctid * per_thread;  // r = L-1-j : position counted from the top of the grid
    const int lane = ctid & 31, warp = ctid >> 5;
    const bool any_rise = rise != nullptr && rise[L] != 0u;  // uniform: read after the barrier, reset after the next
    unsigned local = 0;
    for (int e = 0; e < per_thread; ++e) {
        const int r = r0 + e;
        if (r < L) local += hist[L - r];
    }
    unsigned incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    named_bar_sync(kFlushBarrier, kConsumerThreads);
    unsigned run = incl - local;
    for (int w = 0; w < warp; ++w) run += warp_sums[w];
    for (int e = 0; e < per_thread; ++e) {
        const int r = r0 + e;
        if (r < L) {
            run += hist[L - r];
            hist[L - r] = 0;
            const int j = L - 1 - r;
            unsigned val = run;
            if (any_rise) {  // rare: pixels with a negative width are missed from rise-index k upwards
                for (int k = 0; k <= j; ++k) val += rise[k];
            }
            if (mode >= kFlushFull) {
                int out = static_cast<int>(val);
                if (mode == kFlushFullPlus) out += __ldcg(addend + j);
                counts_row[j] = out;
                // fused path: the loss-table entry while the count is in a register (columns left of the first visited one
                // are zeroed after the decision, by stores only)
                if (table_row != nullptr) table_row[j] = __fdiv_rn(static_cast<float>(out), fpx);
                tot[j] += val;
            } else if (val != 0) {
                if (mode == kFlushSparse) counts_row[j] = static_cast<int>(val);
                else atomicAdd(&counts_row[j], static_cast<int>(val));
                tot[j] += val;
            }
        }
    }
    named_bar_sync(kFlushBarrier, kConsumerThreads);  // histogram is clean, warp_sums reusable
    if (any_rise) {
        for (int k = ctid; k <= L; k += kConsumerThreads) rise[k] = 0u;
        named_bar_sync(kFlushBarrier, kConsumerThreads);
    }
}

// Verdict of the reference's stopping rule `Rhat >= alpha or HB(Rhat) > alpha` for a column with exact total t:
// +1 certainly true, -1 certainly false, 0 unsure (the host replays the reference's own fp32 expression).
__device__ __forceinline__ int screen_verdict(unsigned long long t, const ScreenParams& c) {
    const double R = static_cast<double>(t) / c.n_px;
    const double lo = R * (1.0 - c.gamma), hi = R * (1.0 + c.gamma);
    bool sure_true = lo >= c.alpha32 * (1.0 + 1e-6);
    if (isfinite(c.r_hi)) sure_true = sure_true || (lo > c.r_hi + c.slack);
    bool sure_false = hi < c.alpha32 * (1.0 - 1e-6);
    if (isfinite(c.r_lo)) sure_false = sure_false && (hi < c.r_lo - c.slack);
    int v = sure_true ? 1 : (sure_false ? -1 : 0);
    if (t == 0ull) v = 0;  // HB_mu_plus(0) takes the reference's exception path: never guessed
    return v;
}

__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// FUSED (staged only): the whole calibration step in this one launch - no memset of the outputs (every counts row is
// written in full by the CTA that owns the image's first tile; the part of an image that the NEXT CTA counted travels
// through a per-CTA "head partial" row + flag), per-CTA totals reduced with u64 atomics into a self-cleaning workspace,
// and a last-CTA-done tail: the last CTA (ticket) optionally all-reduces the totals over NVLink peer memory, screens the
// stopping rule, publishes the 16-byte result to device memory AND to mapped pinned host memory, then releases the other
// CTAs, which turn the counts rows they own into the fp32 loss table (zero left of the first visited column).
// Requires floor(total_tiles / gridDim.x) >= tiles_per_image (an image spans at most two CTAs) and all CTAs co-resident
// (cooperative launch).
template <bool STAGED, int HEAD, bool FUSED>
__global__ void __launch_bounds__(kThreads, STAGED ? 1 : 2) rcps_hist_kernel(const RcpsParams prm, const FusedParams fz) {
    constexpr int kFirstPlane = (HEAD == IM2IM_HEAD_RESIDUAL || HEAD == IM2IM_HEAD_GAUSSIAN) ? 1 : 0;  // 2-plane heads
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int L = prm.n_lambdas;
    // layout: [stage ring | STAGED only] [tot u64 x L] [mbarriers] [lambda pairs f32x2 x (L+1)] [lambda f32 x L]
    //         [hist u32 x (L+1)] [warp sums] [rise u32 x (L+1) | RESIDUAL head only]
    unsigned char* cursor = smem_raw;
    float* ring = reinterpret_cast<float*>(cursor);
    if (STAGED) cursor += sizeof(float) * kStages * 4 * kTilePx;
    unsigned long long* tot = reinterpret_cast<unsigned long long*>(cursor);
    cursor += sizeof(unsigned long long) * L;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(cursor);
    uint64_t* empty_bar = full_bar + kStages;
    cursor += sizeof(uint64_t) * 2 * kStages;
    float2* s_pair = reinterpret_cast<float2*>(cursor);  // (lam[g-1], lam[g]) for g = 0..L
    cursor += sizeof(float2) * (L + 1);
    float* s_lam = reinterpret_cast<float*>(cursor);
    cursor += sizeof(float) * L;
    unsigned* hist = reinterpret_cast<unsigned*>(cursor);
    cursor += sizeof(unsigned) * (L + 1);
    unsigned* warp_sums = reinterpret_cast<unsigned*>(cursor);
    cursor += sizeof(unsigned) * kConsumerWarps;
    unsigned* rise = (HEAD == IM2IM_HEAD_RESIDUAL) ? reinterpret_cast<unsigned*>(cursor) : nullptr;
    __shared__ int s_tail[4];   // fused tail: {is last CTA, first visited column, stop verdict, timeout}
    const unsigned epoch = FUSED ? (*fz.epoch + 1u) : 0u;   // read before anything can bump it (only the last CTA does, at the end)
    if (FUSED && blockIdx.x == 0 && threadIdx.x == 0) fz.stamps[0] = global_timer_ns();

    const int tid = threadIdx.x;
    for (int j = tid; j < L; j += kThreads) {
        s_lam[j] = prm.lambdas[j];
        tot[j] = 0ull;
    }
    for (int j = tid; j <= L; j += kThreads) {
        hist[j] = 0u;
        if (HEAD == IM2IM_HEAD_RESIDUAL) rise[j] = 0u;
        s_pair[j] = make_float2(j > 0 ? prm.lambdas[j - 1] : -INFINITY, j < L ? prm.lambdas[j] : INFINITY);
    }
    if (STAGED && tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // contiguous run of tiles for this CTA -> it touches a contiguous run of images
    const long long t_begin = prm.total_tiles * blockIdx.x / gridDim.x;
    const long long t_end = prm.total_tiles * (blockIdx.x + 1) / gridDim.x;
    const long long tpi = prm.tiles_per_image;

    if (tid >= kConsumerThreads) {
        // ------------------------------------------------------------------ producer warp
        if (STAGED && tid == kConsumerThreads) {
            long long img = t_begin / tpi;
            long long r = t_begin - img * tpi;
            int stage = 0;
            unsigned phase = 1;  // first pass over the ring: the slots are free, parity 1 of a fresh barrier succeeds
            for (long long t = t_begin; t < t_end; ++t) {
                mbar_wait(&empty_bar[stage], phase);
                const long long off = r * kTilePx;
                const long long rem = prm.px - off;
                const unsigned bytes = static_cast<unsigned>(rem < kTilePx ? rem : kTilePx) * 4u;
                mbar_arrive_expect_tx(&full_bar[stage], (4u - kFirstPlane) * bytes);
#pragma unroll
                for (int pl = kFirstPlane; pl < 4; ++pl)
                    bulk_g2s(ring + (stage * 4 + pl) * kTilePx, prm.plane[pl] + img * prm.stride[pl] + off, bytes,
                             &full_bar[stage]);
                if (++stage == kStages) { stage = 0; phase ^= 1u; }
                if (++r == tpi) { r = 0; ++img; }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int ctid = tid;
    const int lane = ctid & 31;
    const float lam0 = s_lam[0];
    const float span = s_lam[L - 1] - lam0;
    const float guess_scale = (L > 1 && span > 0.f) ? static_cast<float>(L - 1) / span : 0.f;
    const float guess_bias = -lam0 * guess_scale;
    const int tpi32 = static_cast<int>(tpi);
    const int last_npx = static_cast<int>(prm.px - (tpi - 1) * kTilePx);  // pixels in the last tile of an image

    long long img = t_begin / tpi;
    int r = static_cast<int>(t_begin - img * tpi);
    bool image_started_here = (r == 0);  // did this CTA see tile 0 of the current image?
    int stage = 0;
    unsigned phase = 0;
    const long long n_my_tiles = t_end - t_begin;
    for (long long it = 0; it < n_my_tiles; ++it) {
        const bool image_ends_here = (r == tpi32 - 1);
        const int npx = image_ends_here ? last_npx : kTilePx;
        if (STAGED) {
            mbar_wait(&full_bar[stage], phase);
            const int e0 = ctid * kPxPerThread;
            const bool have = e0 < npx;  // npx % 4 == 0 on this path: all four pixels or none
            float4 vl = make_float4(0.f, 0.f, 0.f, 0.f), vp, vu, vy;
            if (have) {
                const float* base = ring + stage * 4 * kTilePx + e0;
                if (kFirstPlane == 0) vl = *reinterpret_cast<const float4*>(base);
                vp = *reinterpret_cast<const float4*>(base + kTilePx);
                vu = *reinterpret_cast<const float4*>(base + 2 * kTilePx);
                vy = *reinterpret_cast<const float4*>(base + 3 * kTilePx);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);  // values are in registers: hand the slot back early
            if (++stage == kStages) { stage = 0; phase ^= 1u; }
            if (have) {
                const float l[4] = {vl.x, vl.y, vl.z, vl.w}, p[4] = {vp.x, vp.y, vp.z, vp.w};
                const float u[4] = {vu.x, vu.y, vu.z, vu.w}, y[4] = {vy.x, vy.y, vy.z, vy.w};
                PixelQuery q[4];
                int k[4];
                bool ok[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                    q[m] = make_query<HEAD>(l[m], p[m], u[m], y[m]);
                    k[m] = rank_guess<HEAD>(q[m], s_pair, L, guess_scale, guess_bias, ok[m]);
                }
                if (!(ok[0] && ok[1] && ok[2] && ok[3])) {  // rare: guess off by more than the checked window
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        if (!ok[m]) k[m] = resolve_slow<HEAD>(q[m], s_lam, L, rise);
                }
#pragma unroll
                for (int m = 0; m < 4; ++m)
                    if (k[m] > 0) atomicAdd(&hist[k[m]], 1u);
            }
        } else {
            const long long off = static_cast<long long>(r) * kTilePx;
            const float* gl = prm.plane[kFirstPlane] + img * prm.stride[kFirstPlane] + off;  // 2-plane heads: unused
            const float* gp = prm.plane[1] + img * prm.stride[1] + off;
            const float* gu = prm.plane[2] + img * prm.stride[2] + off;
            const float* gy = prm.plane[3] + img * prm.stride[3] + off;
            float l[kPxPerThread], p[kPxPerThread], u[kPxPerThread], y[kPxPerThread];
#pragma unroll
            for (int m = 0; m < kPxPerThread; ++m) {
                const int e = m * kConsumerThreads + ctid;  // coalesced 4-byte loads
                if (e < npx) {
                    l[m] = kFirstPlane == 0 ? ldg_stream_f32(gl + e) : 0.f; p[m] = ldg_stream_f32(gp + e);
                    u[m] = ldg_stream_f32(gu + e); y[m] = ldg_stream_f32(gy + e);
                } else {
                    l[m] = p[m] = u[m] = y[m] = 0.f;  // y == p: inactive, rank 0
                }
            }
            PixelQuery q[kPxPerThread];
            int k[kPxPerThread];
            bool ok[kPxPerThread];
#pragma unroll
            for (int m = 0; m < kPxPerThread; ++m) {
                q[m] = make_query<HEAD>(l[m], p[m], u[m], y[m]);
                k[m] = rank_guess<HEAD>(q[m], s_pair, L, guess_scale, guess_bias, ok[m]);
            }
#pragma unroll
            for (int m = 0; m < kPxPerThread; ++m) {
                if (!ok[m]) k[m] = resolve_slow<HEAD>(q[m], s_lam, L, rise);
                if (k[m] > 0) atomicAdd(&hist[k[m]], 1u);
            }
        }
        if (image_ends_here || it == n_my_tiles - 1) {
            if (!FUSED) {
                flush_image(hist, rise, tot, warp_sums, L, prm.counts + img * L,
                            (image_started_here && image_ends_here) ? kFlushSparse : kFlushAtomic, nullptr, ctid);
            } else if (image_ends_here && image_started_here) {          // whole image counted here
                flush_image(hist, rise, tot, warp_sums, L, prm.counts + img * L, kFlushFull, nullptr, ctid,
                            fz.table ? fz.table + img * L : nullptr, static_cast<float>(prm.px));
            } else if (image_ends_here) {                                 // head partial: the previous CTA owns the row
                flush_image(hist, rise, tot, warp_sums, L, fz.head_partial + static_cast<long long>(blockIdx.x) * L,
                            kFlushFull, nullptr, ctid);
                __threadfence();
                named_bar_sync(kFlushBarrier, kConsumerThreads);
                if (ctid == 0) st_release_gpu_u32(fz.head_flag + blockIdx.x, epoch);
            } else {                                                      // tail partial: add what the next CTA counted
                if (ctid == 0) {
                    // the next CTA publishes its partial row right after its first image: this wait is normally over before
                    // it starts.  Bounded (~2^33 cycles) so that a lost block surfaces as an error instead of a hung GPU.
                    const unsigned* f = fz.head_flag + blockIdx.x + 1;
                    const long long t0 = clock64();
                    while (static_cast<int>(ld_acquire_gpu_u32(f) - epoch) < 0) {
                        if (clock64() - t0 > (1ll << 33)) { atomicExch(fz.err, 1u); break; }
                    }
                }
                flush_image(hist, rise, tot, warp_sums, L, prm.counts + img * L, kFlushFullPlus,
                            fz.head_partial + static_cast<long long>(blockIdx.x + 1) * L, ctid,
                            fz.table ? fz.table + img * L : nullptr, static_cast<float>(prm.px));
            }
        }
        if (++r == tpi32) { r = 0; ++img; image_started_here = true; }
    }
    if (FUSED) {
        // ---------------------------------------------------------------- fused tail
        const int per_thread = (L + kConsumerThreads - 1) / kConsumerThreads;
        for (int e = 0; e < per_thread; ++e) {
            const int rr = ctid * per_thread + e;
            if (rr < L) {
                const int j = L - 1 - rr;
                if (tot[j] != 0ull) atomicAdd(&fz.totals_acc[j], tot[j]);
            }
        }
        __threadfence();
        named_bar_sync(kFlushBarrier, kConsumerThreads);
        if (ctid == 0) {
            const unsigned t = atomicAdd(fz.ticket, 1u);
            s_tail[0] = (t == gridDim.x - 1) ? 1 : 0;
            s_tail[1] = -1; s_tail[2] = 0; s_tail[3] = 0;
        }
        named_bar_sync(kFlushBarrier, kConsumerThreads);
        int first_visited = 0;
        if (s_tail[0]) {
            // ---- last CTA: totals are complete.  Read + clean the workspace, (multi-GPU) exchange, decide, publish.
            if (ctid == 0) fz.stamps[1] = global_timer_ns();
            __threadfence();
            unsigned long long* sum = tot;   // reuse the per-CTA totals array as the reduced totals
            for (int j = ctid; j < L; j += kConsumerThreads) {
                sum[j] = __ldcg(fz.totals_acc + j);
                fz.totals_acc[j] = 0ull;
            }
            if (ctid == 0) { *fz.ticket = 0u; fz.stamps[2] = global_timer_ns(); }
            named_bar_sync(kFlushBarrier, kConsumerThreads);
            if (fz.world > 1) {
                // one-shot push all-reduce over peer memory (protocol of rcps_decide_p2p_kernel)
                const size_t slot = (static_cast<size_t>(epoch & 1u) * fz.world + fz.rank) * L;
                for (int i = ctid; i < fz.world * L; i += kConsumerThreads) {
                    const int pr = i / L, j = i - pr * L;
                    fz.peer_mailbox[pr][slot + j] = sum[j];
                }
                __threadfence_system();
                named_bar_sync(kFlushBarrier, kConsumerThreads);
                if (ctid == 0) fz.stamps[3] = global_timer_ns();
                if (ctid < fz.world) {
                    st_release_sys_u32(fz.peer_flags[ctid] + fz.rank, epoch);
                    const unsigned* f = fz.peer_flags[fz.rank] + ctid;
                    const long long t0 = clock64();
                    while (static_cast<int>(ld_acquire_sys_u32(f) - epoch) < 0) {
                        if (fz.timeout_cycles > 0 && clock64() - t0 > fz.timeout_cycles) { s_tail[3] = 1; break; }
                    }
                }
                named_bar_sync(kFlushBarrier, kConsumerThreads);
                if (ctid == 0) fz.stamps[4] = global_timer_ns();
                if (!s_tail[3]) {
                    const unsigned long long* mine = fz.peer_mailbox[fz.rank] + static_cast<size_t>(epoch & 1u) * fz.world * L;
                    for (int j = ctid; j < L; j += kConsumerThreads) {
                        unsigned long long t = 0ull;
                        for (int pr = 0; pr < fz.world; ++pr) t += ld_relaxed_sys_u64(mine + static_cast<size_t>(pr) * L + j);
                        sum[j] = t;
                    }
                }
                named_bar_sync(kFlushBarrier, kConsumerThreads);
            }
            int best = -1, best_v = 0;
            for (int j = ctid; j < L; j += kConsumerThreads) {
                const unsigned long long t = sum[j];
                if (fz.totals_out != nullptr) fz.totals_out[j] = t;
                const int v = screen_verdict(t, fz.scr);
                if (v >= 0 && j > best) { best = j; best_v = v; }
            }
            if (best >= 0) atomicMax(&s_tail[1], best);
            named_bar_sync(kFlushBarrier, kConsumerThreads);
            if (best >= 0 && best == s_tail[1]) s_tail[2] = best_v;
            named_bar_sync(kFlushBarrier, kConsumerThreads);
            const int first = s_tail[1];
            int r0_, r1_, r2_, r3_;
            if (__ldcg(fz.err) != 0u) { r0_ = -3; r1_ = -3; r2_ = -3; r3_ = 0; }    // an intra-GPU wait gave up
            else if (s_tail[3]) { r0_ = -2; r1_ = -2; r2_ = -2; r3_ = 0; }     // a peer never arrived
            else if (first < 0) { r0_ = -1; r1_ = 1; r2_ = -1; r3_ = 0; }      // ran off the grid
            else if (s_tail[2] > 0) { r0_ = first; r1_ = 1; r2_ = -1; r3_ = first; }
            else { r0_ = -1; r1_ = 0; r2_ = first; r3_ = 0; }                  // unsure: the host replays from `first`
            first_visited = r3_;
            if (ctid == 0) {
                // release the other blocks first (they only need the first visited column), then publish to the host: the
                // system-scope fence of the host write (a PCIe flush, ~3 us) is off the other blocks' critical path
                fz.go[1] = static_cast<unsigned>(r3_);
                __threadfence();
                st_release_gpu_u32(fz.go, epoch);
                fz.result[0] = r0_; fz.result[1] = r1_; fz.result[2] = r2_; fz.result[3] = r3_;
                if (fz.result_host != nullptr) {
                    fz.result_host[0] = r0_; fz.result_host[1] = r1_; fz.result_host[2] = r2_; fz.result_host[3] = r3_;
                    __threadfence_system();
                    fz.result_host[4] = static_cast<int>(epoch);               // tag last: the host polls this word
                }
                fz.stamps[5] = global_timer_ns();
                *fz.epoch = epoch;   // a peer timeout (result -2) leaves the ranks out of step: the caller must rebuild
            }
        } else if (fz.table != nullptr) {
            if (ctid == 0) {
                const long long t0 = clock64();
                const long long limit = (1ll << 33) + 2 * fz.timeout_cycles;   // the last CTA may be waiting for peers
                bool ok = true;
                while (static_cast<int>(ld_acquire_gpu_u32(fz.go) - epoch) < 0) {
                    if (fz.timeout_cycles > 0 && clock64() - t0 > limit) { atomicExch(fz.err, 1u); ok = false; break; }
                    if (fz.world == 1 && clock64() - t0 > (1ll << 33)) { atomicExch(fz.err, 1u); ok = false; break; }
                }
                s_tail[1] = ok ? static_cast<int>(__ldcg(fz.go + 1)) : 0;
            }
            named_bar_sync(kFlushBarrier, kConsumerThreads);
            first_visited = s_tail[1];
        }
        if (fz.table != nullptr && first_visited > 0) {
            // the rows of the images whose FIRST tile lies in this CTA's range were written here (count/px in every
            // column): zero the columns the early-stopped sweep never visits (calibrate_model.py:133) - stores only
            const long long i_begin = (t_begin + tpi - 1) / tpi, i_end = (t_end + tpi - 1) / tpi;
            for (long long i = i_begin; i < i_end; ++i) {
                float* trow = fz.table + i * L;
                for (int j = ctid; j < first_visited; j += kConsumerThreads) trow[j] = 0.f;
            }
        }
        return;
    }
    if (prm.totals != nullptr) {
        const int per_thread = (L + kConsumerThreads - 1) / kConsumerThreads;
        for (int e = 0; e < per_thread; ++e) {
            const int rr = ctid * per_thread + e;
            if (rr < L) {
                const int j = L - 1 - rr;  // same ownership as flush_image: no cross-thread hazard on tot[]
                if (tot[j] != 0ull) atomicAdd(&prm.totals[j], tot[j]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Streaming calibration (the head-fused histogram of conv_halo_kernel): hist[i][k] = #pixels of image i with rank k
// (k = 1..L) -> counts[i][j] = #pixels with rank > j (suffix sums), totals[j] += counts[i][j]; the histogram row is
// zeroed on the way so that the same buffer serves the next batch.  One CTA per image.
__global__ void __launch_bounds__(256) rcps_counts_from_hist_kernel(unsigned* __restrict__ hist, int L,
                                                                    int* __restrict__ counts,
                                                                    unsigned long long* __restrict__ totals) {
    __shared__ unsigned s_warp[8];
    unsigned* row = hist + static_cast<size_t>(blockIdx.x) * (L + 1);
    int* out = counts + static_cast<size_t>(blockIdx.x) * L;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per_thread = (L + 255) / 256;
    const int r0 = tid * per_thread;   // r = L-1-j: position counted from the top of the grid
    unsigned local = 0;
    for (int e = 0; e < per_thread; ++e) {
        const int r = r0 + e;
        if (r < L) local += row[L - r];
    }
    unsigned incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned run = incl - local;
    for (int w = 0; w < warp; ++w) run += s_warp[w];
    for (int e = 0; e < per_thread; ++e) {
        const int r = r0 + e;
        if (r < L) {
            run += row[L - r];
            row[L - r] = 0u;
            const int j = L - 1 - r;
            out[j] = static_cast<int>(run);
            if (run != 0u && totals != nullptr) atomicAdd(totals + j, static_cast<unsigned long long>(run));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rcps_loss_table_kernel(const int* __restrict__ counts, long long n_elems,
                                                              int L, float px, int first_visited,
                                                              const int* __restrict__ d_first_visited,
                                                              float* __restrict__ table) {
    if (d_first_visited != nullptr) first_visited = *d_first_visited;  // decided on the device (rcps_decide_kernel)
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < n_elems; e += stride) {
        const int j = static_cast<int>(e % L);
        table[e] = (j >= first_visited) ? __fdiv_rn(static_cast<float>(counts[e]), px) : 0.f;
    }
}

// Device-side screening of the reference's stopping rule (host twin: calibration/sweep.py::_classify + the scan of
// find_stop_index).  Per column j: exact risk R = totals[j]/(N*px); the reference's fp32 Rhat lies in
// [R(1-gamma), R(1+gamma)]; the rule `Rhat >= alpha or HB(Rhat) > alpha` is certainly true / certainly false / unsure.
// Scanning from the top of the grid, the first column that is not "certainly false" decides:
//   result[0] = stop index (-1: ran off the grid)     result[1] = 1 decided, 0 the host must replay from result[2]
//   result[2] = first unsure column                   result[3] = first visited column for the loss table
__global__ void __launch_bounds__(256) rcps_decide_kernel(const unsigned long long* __restrict__ totals, int L,
                                                          double n_px, double gamma, double alpha32, double r_lo,
                                                          double r_hi, double slack, int* __restrict__ result) {
    __shared__ int s_first;  // highest column index whose verdict is not "certainly false"
    __shared__ int s_verdict;
    if (threadIdx.x == 0) { s_first = -1; s_verdict = 0; }
    __syncthreads();
    int best = -1, best_v = 0;
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        const unsigned long long t = totals[j];
        const double R = static_cast<double>(t) / n_px;
        const double lo = R * (1.0 - gamma), hi = R * (1.0 + gamma);
        bool sure_true = lo >= alpha32 * (1.0 + 1e-6);
        if (isfinite(r_hi)) sure_true = sure_true || (lo > r_hi + slack);
        bool sure_false = hi < alpha32 * (1.0 - 1e-6);
        if (isfinite(r_lo)) sure_false = sure_false && (hi < r_lo - slack);
        int v = sure_true ? 1 : (sure_false ? -1 : 0);
        if (t == 0ull) v = 0;  // HB_mu_plus(0) takes the reference's exception path: never guessed
        if (v >= 0 && j > best) { best = j; best_v = v; }
    }
    if (best >= 0) atomicMax(&s_first, best);
    __syncthreads();
    if (best >= 0 && best == s_first) s_verdict = best_v;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int first = s_first;
        if (first < 0) { result[0] = -1; result[1] = 1; result[2] = -1; result[3] = 0; }
        else if (s_verdict > 0) { result[0] = first; result[1] = 1; result[2] = -1; result[3] = first; }
        else { result[0] = -1; result[1] = 0; result[2] = first; result[3] = 0; }
    }
}

// Multi-GPU: the all-reduce of the per-lambda totals FUSED with the stop decision, over NVLink peer memory (no NCCL call
// on the data path).  One-shot "push" all-reduce: every rank stores its u64[L] totals into its slot of every peer's
// mailbox (peer-mapped symmetric memory, 8 KB per peer at L = 1000), publishes a release-flag per peer, waits for the
// flags of all peers, sums the world slots of its OWN mailbox and runs the decision of rcps_decide_kernel on the sum.
//   mailbox layout (per rank)  u64[2][world][L]   (two epochs: a slot is rewritten two steps later, after every peer has
//                                                   provably finished reading it - it has sent the flag of the step between)
//   flags (per rank)           u32[world]          flag[r] = last epoch rank r has published to this rank (monotone)
//   epoch (per rank, local)    u32                 incremented by the kernel, so a CUDA-graph replay needs no new arguments
// All ranks must call this the same number of times (collective semantics).  One CTA; the spin is on local memory.
__global__ void __launch_bounds__(1024) rcps_decide_p2p_kernel(const unsigned long long* __restrict__ local_totals,
                                                               unsigned long long* const* __restrict__ peer_mailbox,
                                                               unsigned* const* __restrict__ peer_flags,
                                                               unsigned* __restrict__ epoch_dev, int rank, int world, int L,
                                                               double n_px, double gamma, double alpha32, double r_lo,
                                                               double r_hi, double slack,
                                                               unsigned long long* __restrict__ totals_out,
                                                               int* __restrict__ result) {
    __shared__ int s_first, s_verdict;   // result[1] == -2 reports a peer timeout (see the wait below)
    const unsigned epoch = *epoch_dev + 1u;
    const size_t slot = (static_cast<size_t>(epoch & 1u) * world + rank) * L;   // my slot in every mailbox
    // 1. push my totals into every peer's mailbox (including my own)
    for (int i = threadIdx.x; i < world * L; i += blockDim.x) {
        const int r = i / L, j = i - r * L;
        peer_mailbox[r][slot + j] = local_totals[j];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish: flag[rank] on every peer = epoch (release: the stores above are visible before the flag)
    if (threadIdx.x < world) st_release_sys_u32(peer_flags[threadIdx.x] + rank, epoch);
    // 3. wait for every peer's flag in MY flag array (local memory)
    __shared__ int s_timeout;
    if (threadIdx.x == 0) { s_first = -1; s_verdict = 0; s_timeout = 0; }
    __syncthreads();
    if (threadIdx.x < world) {
        const unsigned* f = peer_flags[rank] + threadIdx.x;
        const long long t0 = clock64();
        while (static_cast<int>(ld_acquire_sys_u32(f) - epoch) < 0) {
            // a peer that never arrives (crashed process) must not hang this GPU: give up after ~2^35 cycles (~15 s)
            if (clock64() - t0 > (1ll << 35)) { s_timeout = 1; break; }
        }
    }
    __syncthreads();
    if (s_timeout) {
        if (threadIdx.x == 0) { result[0] = -2; result[1] = -2; result[2] = -2; result[3] = 0; *epoch_dev = epoch; }
        return;
    }
    // 4. sum the world slots of my own mailbox and screen the stopping rule (same logic as rcps_decide_kernel)
    const unsigned long long* mine = peer_mailbox[rank] + static_cast<size_t>(epoch & 1u) * world * L;
    int best = -1, best_v = 0;
    for (int j = threadIdx.x; j < L; j += blockDim.x) {
        unsigned long long t = 0ull;
        for (int r = 0; r < world; ++r) t += ld_relaxed_sys_u64(mine + static_cast<size_t>(r) * L + j);
        totals_out[j] = t;
        const double R = static_cast<double>(t) / n_px;
        const double lo = R * (1.0 - gamma), hi = R * (1.0 + gamma);
        bool sure_true = lo >= alpha32 * (1.0 + 1e-6);
        if (isfinite(r_hi)) sure_true = sure_true || (lo > r_hi + slack);
        bool sure_false = hi < alpha32 * (1.0 - 1e-6);
        if (isfinite(r_lo)) sure_false = sure_false && (hi < r_lo - slack);
        int v = sure_true ? 1 : (sure_false ? -1 : 0);
        if (t == 0ull) v = 0;
        if (v >= 0 && j > best) { best = j; best_v = v; }
    }
    if (best >= 0) atomicMax(&s_first, best);
    __syncthreads();
    if (best >= 0 && best == s_first) s_verdict = best_v;
    __syncthreads();
    if (threadIdx.x == 0) {
        const int first = s_first;
        if (first < 0) { result[0] = -1; result[1] = 1; result[2] = -1; result[3] = 0; }
        else if (s_verdict > 0) { result[0] = first; result[1] = 1; result[2] = -1; result[3] = first; }
        else { result[0] = -1; result[1] = 0; result[2] = first; result[3] = 0; }
        *epoch_dev = epoch;
    }
}

// torch.minimum / torch.maximum semantics (NaN propagates)
__device__ __forceinline__ float t_min(float a, float b) { return (a != a) ? a : ((b != b) ? b : fminf(a, b)); }
__device__ __forceinline__ float t_max(float a, float b) { return (a != a) ? a : ((b != b) ? b : fmaxf(a, b)); }

// Endpoints of one pixel at one lambda, op for op as the reference computes them (head set function, then the outer
// clamp of add_uncertainty.py:35-36).  l1/u1 are the in-place clamped planes of the quantile-type heads.
template <int HEAD>
__device__ __forceinline__ void nested_set(float a, float p, float b, float lam, float& lo2, float& up2, float& l1,
                                           float& u1) {
    const float pm = __fsub_rn(p, 1e-6f), pp = __fadd_rn(p, 1e-6f);
    float upper, lower;
    l1 = a;
    u1 = b;
    if (HEAD == IM2IM_HEAD_QUANTILES) {
        l1 = t_min(a, pm);
        u1 = t_max(b, pp);
        upper = __fadd_rn(__fmul_rn(lam, __fsub_rn(u1, p)), p);
        lower = __fsub_rn(p, __fmul_rn(lam, __fsub_rn(p, l1)));
    } else if (HEAD == IM2IM_HEAD_SOFTMAX_SETS) {
        lower = __fsub_rn(p, __fmul_rn(t_relu(__fsub_rn(p, a)), lam));   // softmax_layer.py:50
        upper = __fadd_rn(p, __fmul_rn(t_relu(__fsub_rn(b, p)), lam));   // softmax_layer.py:51
    } else {
        const float w = (HEAD == IM2IM_HEAD_GAUSSIAN) ? __fsqrt_rn(b) : b;
        upper = __fadd_rn(__fmul_rn(lam, w), p);                         // gaussian_layer.py:31 / residual :33
        lower = __fadd_rn(__fmul_rn(-lam, w), p);                        // gaussian_layer.py:32 / residual :34
    }
    up2 = t_max(upper, pp);
    lo2 = t_min(lower, pm);
}

template <int HEAD>
__global__ void __launch_bounds__(256) nested_sets_kernel(float* lower, const float* __restrict__ pred, float* upper,
                                                          long long n_images, long long px, long long sl,
                                                          long long sp, long long su, float lam, int write_back,
                                                          float* __restrict__ lower_out,
                                                          float* __restrict__ upper_out) {
    constexpr bool kThree = (HEAD == IM2IM_HEAD_QUANTILES || HEAD == IM2IM_HEAD_SOFTMAX_SETS);
    const long long total = n_images * px;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long i = e / px, k = e - i * px;
        float lo2, up2, l1, u1;
        nested_set<HEAD>(kThree ? lower[i * sl + k] : 0.f, ldg_stream_f32(pred + i * sp + k), upper[i * su + k], lam,
                         lo2, up2, l1, u1);
        lower_out[e] = lo2;
        upper_out[e] = up2;
        if (HEAD == IM2IM_HEAD_QUANTILES && write_back) {  // the reference's in-place clamp (quantile_layer.py:39-40)
            lower[i * sl + k] = l1;
            upper[i * su + k] = u1;
        }
    }
}

// map[k] += #{images in this block's slab whose pixel k is missed}; threads run along pixels (coalesced).
template <int HEAD>
__global__ void __launch_bounds__(256) miss_map_kernel(const float* __restrict__ lower, const float* __restrict__ pred,
                                                       const float* __restrict__ upper,
                                                       const float* __restrict__ label, long long n_images,
                                                       long long px, long long sl, long long sp, long long su,
                                                       long long sy, float lam, int images_per_slab,
                                                       int* __restrict__ map) {
    constexpr bool kThree = (HEAD == IM2IM_HEAD_QUANTILES || HEAD == IM2IM_HEAD_SOFTMAX_SETS);
    const long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k >= px) return;
    const long long i0 = static_cast<long long>(blockIdx.y) * images_per_slab;
    const long long i1 = min(i0 + images_per_slab, n_images);
    int acc = 0;
    for (long long i = i0; i < i1; ++i) {
        float lo2, up2, l1, u1;
        nested_set<HEAD>(kThree ? ldg_stream_f32(lower + i * sl + k) : 0.f, ldg_stream_f32(pred + i * sp + k),
                         ldg_stream_f32(upper + i * su + k), lam, lo2, up2, l1, u1);
        const float y = ldg_stream_f32(label + i * sy + k);
        acc += ((lo2 > y) || (up2 < y)) ? 1 : 0;
    }
    if (acc) atomicAdd(&map[k], acc);
}

// counts[i] = #{k : lower[i,k] > y[i,k] or upper[i,k] < y[i,k]} for already-computed endpoints
// (fraction_missed_loss, core/calibration/calibrate_model.py:76-80).  One CTA per (image, slab) pair.
__global__ void __launch_bounds__(256) fraction_missed_kernel(const float* __restrict__ lower,
                                                              const float* __restrict__ upper,
                                                              const float* __restrict__ label, long long px,
                                                              long long sl, long long su, long long sy,
                                                              int* __restrict__ counts) {
    const long long i = blockIdx.y;
    const float* l = lower + i * sl;
    const float* u = upper + i * su;
    const float* y = label + i * sy;
    int acc = 0;
    for (long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < px;
         k += static_cast<long long>(gridDim.x) * blockDim.x) {
        const float yy = ldg_stream_f32(y + k);
        // (lower>y)+(upper<y) clipped to 1  ==  logical or
        acc += ((ldg_stream_f32(l + k) > yy) || (ldg_stream_f32(u + k) < yy)) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&counts[i], acc);
}

// Softmax head: logits -> (lower quantile, argmax prediction, upper quantile), the lambda-independent part of
// softmax_nested_sets_from_output (softmax_layer.py:34-48); the lambda-dependent part (:50-51) is the SOFTMAX_SETS head
// kind of the sweep / nested-set kernels.  Operation order of the reference: p_k = e_k / sum with one IEEE division per
// class (:34), cumulative sum accumulated in double and rounded to fp32 at every step (torch.cumsum on the CPU, :38),
// compared with fp32 0.05 / 0.95 (:40-41), first maximal p_k as the prediction (:42).  exp is portable_expf - the same fixed
// sequence of IEEE operations as the CPU checker's C restatement (tests), so the two agree bit for bit.
// One thread per pixel, classes strided by `sk` elements (coalesced across pixels); the K <= kMaxSoftmax exponentials are
// kept in registers between the two passes, so the logits are read from HBM once.
constexpr int kMaxSoftmax = 64;

__device__ __forceinline__ float portable_expf(float x) {   // x <= 0; the test suite's C checker carries the same sequence
    const float q = rintf(__fmul_rn(x, 1.4426950408889634f));
    float s = __fmaf_rn(q, -0.693145751953125f, x);
    s = __fmaf_rn(q, -1.428606765330187045e-06f, s);
    float u = 0.000198527617612853646278381f;
    u = __fmaf_rn(u, s, 0.00139304355252534151077271f);
    u = __fmaf_rn(u, s, 0.00833336077630519866943359f);
    u = __fmaf_rn(u, s, 0.0416664853692054748535156f);
    u = __fmaf_rn(u, s, 0.166666671633720397949219f);
    u = __fmaf_rn(u, s, 0.5f);
    u = __fadd_rn(__fmaf_rn(__fmul_rn(s, s), u, s), 1.0f);
    const int qi = static_cast<int>(q), q1 = qi >> 1, q2 = qi - q1;   // 2^q as two exact factors in the normal range
    u = __fmul_rn(__fmul_rn(u, __int_as_float((q1 + 127) << 23)), __int_as_float((q2 + 127) << 23));
    return x < -104.0f ? 0.0f : u;
}

template <int KMAX>
__global__ void __launch_bounds__(256) softmax_sets_kernel(const float* __restrict__ logits, long long n_images, int K,
                                                           long long inner, long long si, long long sk,
                                                           float* __restrict__ sets) {
    const long long total = n_images * inner;
    const float fK = static_cast<float>(K);
    const float step = static_cast<float>(1.0 / static_cast<double>(K));  // python float 1/K entering an fp32 op
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long i = e / inner, k0 = e - i * inner;
        const float* src = logits + i * si + k0;
        float v[KMAX];
        float m = -INFINITY;
        bool has_nan = false;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                v[k] = ldg_stream_f32(src + k * sk);
                has_nan = has_nan || (v[k] != v[k]);
                m = fmaxf(m, v[k]);
            }
        }
        float ssum = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                v[k] = portable_expf(__fsub_rn(v[k], m));
                ssum = __fadd_rn(ssum, v[k]);
            }
        }
        // p_k = RN(e_k / s): the denominator is common to the row, so the correctly rounded quotient comes from ONE IEEE
        // reciprocal and Markstein's sequence q = RN(e*r), rem = e - q*s (exact, fma), q' = RN(q + rem*r) == RN(e/s) whenever
        // r = RN(1/s), nothing underflows and s is not all-ones in the mantissa; everything else takes the IEEE division.
        // (float)cum <= 0.05f  <=>  cum <= T_LO with T_LO the largest double that rounds to a float <= 0.05f (rounding is
        // monotone): no double->float conversion per class.
        constexpr double T_LO = 0x1.99999afffffffp-5, T_HI = 0x1.e66666fffffffp-1;
        const float rcp_s = __frcp_rn(ssum);
        const bool fast_div = (ssum >= 1.0f) && (ssum <= 128.0f) && ((__float_as_uint(ssum) & 0x7FFFFFu) != 0x7FFFFFu);
        double cum = 0.0;
        float best = -INFINITY;
        int n_lo = 0, n_hi = 0, arg = 0;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            if (k < K) {
                float pk;
                if (fast_div && v[k] >= 0x1p-60f) {
                    const float q = __fmul_rn(v[k], rcp_s);
                    const float rem = __fmaf_rn(-q, ssum, v[k]);
                    pk = __fmaf_rn(rem, rcp_s, q);
                } else {
                    pk = __fdiv_rn(v[k], ssum);
                }
                cum += static_cast<double>(pk);
                n_lo += (cum <= T_LO) ? 1 : 0;
                n_hi += (cum <= T_HI) ? 1 : 0;
                if (pk > best) { best = pk; arg = k; }  // first maximal element, like torch.argmax
            }
        }
        float lq, pr, uq;
        if (has_nan) {
            // a NaN logit makes the whole softmax row NaN: every `cumsum <= q` is false and torch.argmax returns the
            // first NaN, i.e. class 0
            lq = 0.f; uq = 0.f; pr = 0.f;
        } else {
            lq = __fdiv_rn(static_cast<float>(n_lo), fK);
            uq = __fdiv_rn(static_cast<float>(n_hi), fK);
            pr = __fdiv_rn(static_cast<float>(arg), fK);
        }
        if (pr == lq) lq = __fsub_rn(lq, step);   // softmax_layer.py:45
        if (pr == uq) uq = __fadd_rn(uq, step);   // softmax_layer.py:46
        lq = fminf(fmaxf(lq, 0.f), 1.f);          // :47-48
        uq = fminf(fmaxf(uq, 0.f), 1.f);
        float* dst = sets + i * 3 * inner + k0;
        dst[0] = lq;
        dst[inner] = pr;
        dst[2 * inner] = uq;
    }
}

size_t hist_smem_bytes(bool staged, int L, int head) {
    size_t b = staged ? sizeof(float) * kStages * 4 * kTilePx : 0;
    b += sizeof(unsigned long long) * L + sizeof(uint64_t) * 2 * kStages + sizeof(float2) * (L + 1) +
         sizeof(float) * L + sizeof(unsigned) * (L + 1) + sizeof(unsigned) * kConsumerWarps;
    if (head == IM2IM_HEAD_RESIDUAL) b += sizeof(unsigned) * (L + 1);  // rising histogram (negative widths)
    return b;
}

bool head_known(int head) { return head >= IM2IM_HEAD_QUANTILES && head <= IM2IM_HEAD_SOFTMAX_SETS; }
bool head_three_planes(int head) { return head == IM2IM_HEAD_QUANTILES || head == IM2IM_HEAD_SOFTMAX_SETS; }

template <bool STAGED, int HEAD>
int launch_hist(const RcpsParams& prm, size_t smem, long long grid, cudaStream_t st) {
    IM2IM_CUDA_TRY(cudaFuncSetAttribute(rcps_hist_kernel<STAGED, HEAD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    FusedParams none{};
    rcps_hist_kernel<STAGED, HEAD, false><<<static_cast<unsigned>(grid), kThreads, smem, st>>>(prm, none);
    return check_launch(STAGED ? "rcps_hist_kernel<staged>" : "rcps_hist_kernel<generic>");
}

// cooperative launch: every CTA is guaranteed to be resident (the tail spins on flags set by other CTAs)
template <int HEAD>
int launch_fused(const RcpsParams& prm, const FusedParams& fz, size_t smem, long long grid, cudaStream_t st) {
    IM2IM_CUDA_TRY(cudaFuncSetAttribute(rcps_hist_kernel<true, HEAD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
    void* args[2] = {const_cast<RcpsParams*>(&prm), const_cast<FusedParams*>(&fz)};
    IM2IM_CUDA_TRY(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&rcps_hist_kernel<true, HEAD, true>),
                                               dim3(static_cast<unsigned>(grid)), dim3(kThreads), args, smem, st));
    return check_launch("rcps_hist_kernel<fused>");
}

template <bool STAGED>
int launch_hist_head(int head, const RcpsParams& prm, size_t smem, long long grid, cudaStream_t st) {
    switch (head) {
        case IM2IM_HEAD_QUANTILES: return launch_hist<STAGED, IM2IM_HEAD_QUANTILES>(prm, smem, grid, st);
        case IM2IM_HEAD_RESIDUAL: return launch_hist<STAGED, IM2IM_HEAD_RESIDUAL>(prm, smem, grid, st);
        case IM2IM_HEAD_GAUSSIAN: return launch_hist<STAGED, IM2IM_HEAD_GAUSSIAN>(prm, smem, grid, st);
        default: return launch_hist<STAGED, IM2IM_HEAD_SOFTMAX_SETS>(prm, smem, grid, st);
    }
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace im2im

using namespace im2im;

extern "C" int im2im_rcps_miss_counts(const float* d_lower, const float* d_pred, const float* d_upper,
                                      const float* d_label, int64_t n_images, int64_t px, int64_t stride_lower,
                                      int64_t stride_pred, int64_t stride_upper, int64_t stride_label,
                                      const float* d_lambdas, int32_t n_lambdas, int32_t head_kind,
                                      int32_t* d_counts, unsigned long long* d_totals, uint32_t flags,
                                      void* stream) {
    if (!head_known(head_kind)) return fail(IM2IM_ENOTSUP, "head_kind %d not implemented", head_kind);
    if (!head_three_planes(head_kind)) { d_lower = d_pred; stride_lower = stride_pred; }  // 2-plane heads: unused
    if (n_images < 0 || px < 0) return fail(IM2IM_EINVAL, "negative size (n_images=%lld px=%lld)",
                                            (long long)n_images, (long long)px);
    if (n_lambdas < 1 || n_lambdas > IM2IM_RCPS_MAX_LAMBDAS)
        return fail(IM2IM_ERANGE, "n_lambdas=%d outside [1, %d]", n_lambdas, IM2IM_RCPS_MAX_LAMBDAS);
    if (px >= (1ll << 24)) return fail(IM2IM_ERANGE, "px=%lld >= 2^24: fp32 per-image mean is not exact", (long long)px);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_images == 0 || px == 0) {  // empty calibration set / empty images: nothing is missed
        if ((flags & IM2IM_RCPS_ZERO_OUTPUTS) && d_totals)
            IM2IM_CUDA_TRY(cudaMemsetAsync(d_totals, 0, sizeof(unsigned long long) * n_lambdas, st));
        if ((flags & IM2IM_RCPS_ZERO_OUTPUTS) && d_counts && n_images > 0)
            IM2IM_CUDA_TRY(cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * n_images * n_lambdas, st));
        return IM2IM_OK;
    }
    if (d_lambdas == nullptr || d_counts == nullptr) return fail(IM2IM_EINVAL, "null lambda grid or counts");
    if (flags & IM2IM_RCPS_ZERO_OUTPUTS) {
        if (n_images > 0)
            IM2IM_CUDA_TRY(cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * n_images * n_lambdas, st));
        if (d_totals) IM2IM_CUDA_TRY(cudaMemsetAsync(d_totals, 0, sizeof(unsigned long long) * n_lambdas, st));
    }
    if (!d_lower || !d_pred || !d_upper || !d_label) return fail(IM2IM_EINVAL, "null score plane");

    bool fast = !(flags & IM2IM_RCPS_FORCE_GENERIC) && (px % 4 == 0) && aligned16(d_lower) &&
                      aligned16(d_pred) && aligned16(d_upper) && aligned16(d_label) && (stride_lower % 4 == 0) &&
                      (stride_pred % 4 == 0) && (stride_upper % 4 == 0) && (stride_label % 4 == 0);
    RcpsParams prm;
    prm.plane[0] = d_lower; prm.plane[1] = d_pred; prm.plane[2] = d_upper; prm.plane[3] = d_label;
    prm.stride[0] = stride_lower; prm.stride[1] = stride_pred; prm.stride[2] = stride_upper; prm.stride[3] = stride_label;
    prm.n_images = n_images;
    prm.px = px;
    prm.tiles_per_image = (px + kTilePx - 1) / kTilePx;
    prm.total_tiles = prm.tiles_per_image * n_images;
    prm.lambdas = d_lambdas;
    prm.n_lambdas = n_lambdas;
    prm.counts = d_counts;
    prm.totals = d_totals;

    if (fast && hist_smem_bytes(true, n_lambdas, head_kind) > kMaxSmemOptin) fast = false;  // very long grids: no room for the ring
    const size_t smem = hist_smem_bytes(fast, n_lambdas, head_kind);
    const int sms = sm_count();
    if (fast) {
        const long long grid = prm.total_tiles < sms ? prm.total_tiles : sms;
        return launch_hist_head<true>(head_kind, prm, smem, grid, st);
    }
    const long long cap = 2ll * sms;
    const long long grid = prm.total_tiles < cap ? prm.total_tiles : cap;
    return launch_hist_head<false>(head_kind, prm, smem, grid, st);
}


namespace im2im { namespace {
constexpr int kFusedMaxGrid = 256;
struct FusedWorkspace {      // byte offsets inside the caller's workspace
    static size_t stamps() { return 64; }       // u64[8] %globaltimer stamps of the last launch (profiling aid)
    static size_t head_flag() { return 128; }
    static size_t totals_acc() { return 128 + sizeof(unsigned) * kFusedMaxGrid; }
    static size_t head_partial(int L) { return totals_acc() + sizeof(unsigned long long) * L; }
    static size_t bytes(int L) { return head_partial(L) + sizeof(int) * static_cast<size_t>(kFusedMaxGrid) * L; }
};
} }

namespace im2im { namespace {
bool fused_fast_path(const float* l, const float* p, const float* u, const float* y, int64_t px, int64_t sl, int64_t sp,
                     int64_t su, int64_t sy, int n_lambdas, int head_kind) {
    return (px % 4 == 0) && aligned16(l) && aligned16(p) && aligned16(u) && aligned16(y) && (sl % 4 == 0) && (sp % 4 == 0) &&
           (su % 4 == 0) && (sy % 4 == 0) && hist_smem_bytes(true, n_lambdas, head_kind) <= kMaxSmemOptin;
}
} }

extern "C" int im2im_rcps_calibrate_fused_check(const float* d_lower, const float* d_pred, const float* d_upper,
                                                const float* d_label, int64_t n_images, int64_t px, int64_t stride_lower,
                                                int64_t stride_pred, int64_t stride_upper, int64_t stride_label,
                                                int32_t n_lambdas, int32_t head_kind) {
    if (!head_known(head_kind)) return fail(IM2IM_ENOTSUP, "head_kind %d not implemented", head_kind);
    if (!head_three_planes(head_kind)) { d_lower = d_pred; stride_lower = stride_pred; }
    if (n_images <= 0 || px <= 0 || px >= (1ll << 24) || n_lambdas < 1 || n_lambdas > IM2IM_RCPS_MAX_LAMBDAS)
        return fail(IM2IM_ENOTSUP, "calibrate_fused: shape outside the fused path");
    if (!fused_fast_path(d_lower, d_pred, d_upper, d_label, px, stride_lower, stride_pred, stride_upper, stride_label,
                         n_lambdas, head_kind))
        return fail(IM2IM_ENOTSUP, "calibrate_fused: needs 16-byte aligned planes, px %% 4 == 0 and room for the staging ring");
    return IM2IM_OK;
}

extern "C" size_t im2im_rcps_fused_workspace_bytes(int32_t n_lambdas) {
    if (n_lambdas < 1 || n_lambdas > IM2IM_RCPS_MAX_LAMBDAS) return 0;
    return FusedWorkspace::bytes(n_lambdas);
}

extern "C" int im2im_rcps_calibrate_fused(const float* d_lower, const float* d_pred, const float* d_upper,
                                          const float* d_label, int64_t n_images, int64_t px, int64_t stride_lower,
                                          int64_t stride_pred, int64_t stride_upper, int64_t stride_label,
                                          const float* d_lambdas, int32_t n_lambdas, int32_t head_kind,
                                          int32_t* d_counts, float* d_table, unsigned long long* d_totals_out,
                                          double n_images_times_px, double gamma, double alpha32, double r_lo,
                                          double r_hi, double slack, void* d_workspace, size_t workspace_bytes,
                                          unsigned long long* const* d_peer_mailboxes, unsigned* const* d_peer_flags,
                                          int32_t rank, int32_t world, double peer_timeout_s, int32_t* d_result,
                                          int32_t* h_result_mapped, void* stream) {
    if (!head_known(head_kind)) return fail(IM2IM_ENOTSUP, "head_kind %d not implemented", head_kind);
    if (!head_three_planes(head_kind)) { d_lower = d_pred; stride_lower = stride_pred; }
    if (n_images <= 0 || px <= 0) return fail(IM2IM_EINVAL, "calibrate_fused: empty calibration set");
    if (n_lambdas < 1 || n_lambdas > IM2IM_RCPS_MAX_LAMBDAS)
        return fail(IM2IM_ERANGE, "n_lambdas=%d outside [1, %d]", n_lambdas, IM2IM_RCPS_MAX_LAMBDAS);
    if (px >= (1ll << 24)) return fail(IM2IM_ERANGE, "px=%lld >= 2^24: fp32 per-image mean is not exact", (long long)px);
    if (!d_lower || !d_pred || !d_upper || !d_label || !d_lambdas || !d_counts || !d_result || !d_workspace)
        return fail(IM2IM_EINVAL, "calibrate_fused: null pointer");
    if (!(n_images_times_px > 0)) return fail(IM2IM_EINVAL, "calibrate_fused: empty calibration set");
    if (world < 1 || world > 64 || rank < 0 || rank >= world) return fail(IM2IM_EINVAL, "bad rank/world (%d/%d)", rank, world);
    if (world > 1 && (!d_peer_mailboxes || !d_peer_flags)) return fail(IM2IM_EINVAL, "calibrate_fused: null peer tables");
    if (workspace_bytes < FusedWorkspace::bytes(n_lambdas))
        return fail(IM2IM_EINVAL, "calibrate_fused: workspace of %zu bytes, %zu needed", workspace_bytes,
                    FusedWorkspace::bytes(n_lambdas));
    const bool fast = fused_fast_path(d_lower, d_pred, d_upper, d_label, px, stride_lower, stride_pred, stride_upper,
                                      stride_label, n_lambdas, head_kind);
    if (!fast) return fail(IM2IM_ENOTSUP, "calibrate_fused: needs 16-byte aligned planes, px %% 4 == 0 and a lambda grid that "
                                         "leaves room for the staging ring; use the multi-launch path");
    RcpsParams prm;
    prm.plane[0] = d_lower; prm.plane[1] = d_pred; prm.plane[2] = d_upper; prm.plane[3] = d_label;
    prm.stride[0] = stride_lower; prm.stride[1] = stride_pred; prm.stride[2] = stride_upper; prm.stride[3] = stride_label;
    prm.n_images = n_images;
    prm.px = px;
    prm.tiles_per_image = (px + kTilePx - 1) / kTilePx;
    prm.total_tiles = prm.tiles_per_image * n_images;
    prm.lambdas = d_lambdas;
    prm.n_lambdas = n_lambdas;
    prm.counts = d_counts;
    prm.totals = nullptr;
    int sms = sm_count();
    if (sms > kFusedMaxGrid) sms = kFusedMaxGrid;
    long long grid = prm.total_tiles < sms ? prm.total_tiles : sms;
    // an image may span at most two CTAs: shrink the grid until every CTA owns at least one image worth of tiles
    if (prm.total_tiles / grid < prm.tiles_per_image) grid = prm.total_tiles / prm.tiles_per_image;   // = n_images
    if (grid < 1) grid = 1;
    unsigned char* ws = static_cast<unsigned char*>(d_workspace);
    FusedParams fz;
    fz.table = d_table;
    fz.epoch = reinterpret_cast<unsigned*>(ws);
    fz.ticket = reinterpret_cast<unsigned*>(ws) + 1;
    fz.go = reinterpret_cast<unsigned*>(ws) + 2;
    fz.err = reinterpret_cast<unsigned*>(ws) + 4;
    fz.stamps = reinterpret_cast<unsigned long long*>(ws + FusedWorkspace::stamps());
    fz.head_flag = reinterpret_cast<unsigned*>(ws + FusedWorkspace::head_flag());
    fz.totals_acc = reinterpret_cast<unsigned long long*>(ws + FusedWorkspace::totals_acc());
    fz.head_partial = reinterpret_cast<int*>(ws + FusedWorkspace::head_partial(n_lambdas));
    fz.scr.n_px = n_images_times_px; fz.scr.gamma = gamma; fz.scr.alpha32 = alpha32;
    fz.scr.r_lo = r_lo; fz.scr.r_hi = r_hi; fz.scr.slack = slack;
    fz.peer_mailbox = d_peer_mailboxes; fz.peer_flags = d_peer_flags; fz.rank = rank; fz.world = world;
    fz.timeout_cycles = 0;
    if (peer_timeout_s > 0) {
        int dev = 0, khz = 0;
        IM2IM_CUDA_TRY(cudaGetDevice(&dev));
        IM2IM_CUDA_TRY(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
        fz.timeout_cycles = static_cast<long long>(peer_timeout_s * 1e3 * static_cast<double>(khz));
    }
    fz.totals_out = d_totals_out;
    fz.result = d_result;
    fz.result_host = nullptr;
    if (h_result_mapped != nullptr) {   // pinned host memory as the device sees it (same address under unified addressing)
        void* dptr = nullptr;
        IM2IM_CUDA_TRY(cudaHostGetDevicePointer(&dptr, h_result_mapped, 0));
        fz.result_host = static_cast<volatile int*>(dptr);
    }
    const size_t smem = hist_smem_bytes(true, n_lambdas, head_kind);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (head_kind) {
        case IM2IM_HEAD_QUANTILES: return launch_fused<IM2IM_HEAD_QUANTILES>(prm, fz, smem, grid, st);
        case IM2IM_HEAD_RESIDUAL: return launch_fused<IM2IM_HEAD_RESIDUAL>(prm, fz, smem, grid, st);
        case IM2IM_HEAD_GAUSSIAN: return launch_fused<IM2IM_HEAD_GAUSSIAN>(prm, fz, smem, grid, st);
        default: return launch_fused<IM2IM_HEAD_SOFTMAX_SETS>(prm, fz, smem, grid, st);
    }
}

// Host side of the fused step: spin on the epoch tag the kernel writes into mapped pinned memory (h_result[4]) instead of
// a cudaStreamSynchronize + 16-byte copy.  Returns 0 when the tag reached `expected`, 1 after `spin_us` microseconds
// (the caller then falls back to synchronising the stream).
extern "C" int im2im_host_wait_flag(const volatile int32_t* h_flag, int32_t expected, int64_t spin_us) {
    if (!h_flag) return fail(IM2IM_EINVAL, "host_wait_flag: null flag");
    timespec t0;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (unsigned it = 0;; ++it) {
        if (static_cast<int32_t>(*h_flag - expected) >= 0) return 0;
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
        if ((it & 1023u) == 1023u) {
            timespec t1;
            clock_gettime(CLOCK_MONOTONIC, &t1);
            const long long us = (t1.tv_sec - t0.tv_sec) * 1000000ll + (t1.tv_nsec - t0.tv_nsec) / 1000;
            if (us > spin_us) return 1;
        }
    }
}

extern "C" int im2im_rcps_counts_from_hist(uint32_t* d_hist, int64_t n_images, int32_t n_lambdas, int32_t* d_counts,
                                           unsigned long long* d_totals_or_null, void* stream) {
    if (n_images < 0 || n_lambdas < 1) return fail(IM2IM_EINVAL, "bad histogram shape");
    if (n_images == 0) return IM2IM_OK;
    if (!d_hist || !d_counts) return fail(IM2IM_EINVAL, "null histogram/counts");
    if (n_images > 0x7fffffffLL) return fail(IM2IM_ERANGE, "too many images for one launch");
    rcps_counts_from_hist_kernel<<<static_cast<unsigned>(n_images), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_hist, n_lambdas, d_counts, d_totals_or_null);
    return check_launch("rcps_counts_from_hist_kernel");
}

extern "C" int im2im_rcps_loss_table(const int32_t* d_counts, int64_t n_images, int32_t n_lambdas, int64_t px,
                                     int32_t first_visited_col, float* d_table, void* stream) {
    if (n_images < 0 || n_lambdas < 1 || px < 1) return fail(IM2IM_EINVAL, "bad table shape");
    if (n_images == 0) return IM2IM_OK;
    if (!d_counts || !d_table) return fail(IM2IM_EINVAL, "null counts/table");
    const long long n = static_cast<long long>(n_images) * n_lambdas;
    const long long blocks = (n + 255) / 256;
    const long long cap = 8ll * sm_count();
    rcps_loss_table_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), 256, 0,
                             static_cast<cudaStream_t>(stream)>>>(d_counts, n, n_lambdas, static_cast<float>(px),
                                                                  first_visited_col, nullptr, d_table);
    return check_launch("rcps_loss_table_kernel");
}

extern "C" int im2im_rcps_loss_table_dev(const int32_t* d_counts, int64_t n_images, int32_t n_lambdas, int64_t px,
                                         const int32_t* d_first_visited_col, float* d_table, void* stream) {
    if (n_images < 0 || n_lambdas < 1 || px < 1) return fail(IM2IM_EINVAL, "bad table shape");
    if (n_images == 0) return IM2IM_OK;
    if (!d_counts || !d_table || !d_first_visited_col) return fail(IM2IM_EINVAL, "null counts/table/first column");
    const long long n = static_cast<long long>(n_images) * n_lambdas;
    const long long blocks = (n + 255) / 256;
    const long long cap = 8ll * sm_count();
    rcps_loss_table_kernel<<<static_cast<unsigned>(blocks < cap ? blocks : cap), 256, 0,
                             static_cast<cudaStream_t>(stream)>>>(d_counts, n, n_lambdas, static_cast<float>(px), 0,
                                                                  d_first_visited_col, d_table);
    return check_launch("rcps_loss_table_kernel");
}

extern "C" int im2im_rcps_decide(const unsigned long long* d_totals, int32_t n_lambdas, double n_images_times_px,
                                 double gamma, double alpha32, double r_lo, double r_hi, double slack,
                                 int32_t* d_result, void* stream) {
    if (n_lambdas < 1 || n_lambdas > IM2IM_RCPS_MAX_LAMBDAS) return fail(IM2IM_ERANGE, "n_lambdas=%d", n_lambdas);
    if (!d_totals || !d_result) return fail(IM2IM_EINVAL, "null totals/result");
    if (!(n_images_times_px > 0)) return fail(IM2IM_EINVAL, "empty calibration set");
    rcps_decide_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(d_totals, n_lambdas, n_images_times_px,
                                                                          gamma, alpha32, r_lo, r_hi, slack, d_result);
    return check_launch("rcps_decide_kernel");
}

extern "C" int im2im_nested_sets(int32_t head_kind, float* d_lower, const float* d_pred, float* d_upper,
                                int64_t n_images, int64_t px, int64_t stride_lower, int64_t stride_pred,
                                int64_t stride_upper, float lam, int32_t write_back_clamp, float* d_lower_out,
                                float* d_upper_out, void* stream) {
    if (!head_known(head_kind)) return fail(IM2IM_ENOTSUP, "head_kind %d not implemented", head_kind);
    if (!head_three_planes(head_kind)) { d_lower = d_upper; stride_lower = stride_upper; }  // unused plane
    if (n_images < 0 || px < 0) return fail(IM2IM_EINVAL, "negative size");
    if (n_images == 0 || px == 0) return IM2IM_OK;
    if (!d_lower || !d_pred || !d_upper || !d_lower_out || !d_upper_out) return fail(IM2IM_EINVAL, "null plane");
    const long long n = static_cast<long long>(n_images) * px;
    const long long blocks = (n + 255) / 256;
    const long long cap = 16ll * sm_count();
    const unsigned grid = static_cast<unsigned>(blocks < cap ? blocks : cap);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
#define IM2IM_NS_LAUNCH(H)                                                                                          \
    nested_sets_kernel<H><<<grid, 256, 0, st>>>(d_lower, d_pred, d_upper, n_images, px, stride_lower, stride_pred, \
                                                stride_upper, lam, write_back_clamp, d_lower_out, d_upper_out)
    switch (head_kind) {
        case IM2IM_HEAD_QUANTILES: IM2IM_NS_LAUNCH(IM2IM_HEAD_QUANTILES); break;
        case IM2IM_HEAD_RESIDUAL: IM2IM_NS_LAUNCH(IM2IM_HEAD_RESIDUAL); break;
        case IM2IM_HEAD_GAUSSIAN: IM2IM_NS_LAUNCH(IM2IM_HEAD_GAUSSIAN); break;
        default: IM2IM_NS_LAUNCH(IM2IM_HEAD_SOFTMAX_SETS); break;
    }
#undef IM2IM_NS_LAUNCH
    return check_launch("nested_sets_kernel");
}

extern "C" int im2im_quantile_nested_sets(float* d_lower, const float* d_pred, float* d_upper, int64_t n_images,
                                          int64_t px, int64_t stride_lower, int64_t stride_pred,
                                          int64_t stride_upper, float lam, int32_t write_back_clamp,
                                          float* d_lower_out, float* d_upper_out, void* stream) {
    return im2im_nested_sets(IM2IM_HEAD_QUANTILES, d_lower, d_pred, d_upper, n_images, px, stride_lower, stride_pred,
                             stride_upper, lam, write_back_clamp, d_lower_out, d_upper_out, stream);
}

extern "C" int im2im_rcps_miss_map(const float* d_lower, const float* d_pred, const float* d_upper,
                                   const float* d_label, int64_t n_images, int64_t px, int64_t stride_lower,
                                   int64_t stride_pred, int64_t stride_upper, int64_t stride_label, float lam,
                                   int32_t head_kind, int32_t* d_map, uint32_t flags, void* stream) {
    if (!head_known(head_kind)) return fail(IM2IM_ENOTSUP, "head_kind %d not implemented", head_kind);
    if (!head_three_planes(head_kind)) { d_lower = d_pred; stride_lower = stride_pred; }  // unused plane
    if (n_images < 0 || px < 0) return fail(IM2IM_EINVAL, "negative size");
    if (px == 0) return IM2IM_OK;
    if (!d_map) return fail(IM2IM_EINVAL, "null map");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (flags & IM2IM_RCPS_ZERO_OUTPUTS) IM2IM_CUDA_TRY(cudaMemsetAsync(d_map, 0, sizeof(int32_t) * px, st));
    if (n_images == 0) return IM2IM_OK;
    if (!d_lower || !d_pred || !d_upper || !d_label) return fail(IM2IM_EINVAL, "null score plane");
    const unsigned bx = static_cast<unsigned>((px + 255) / 256);
    // enough image slabs to fill the GPU when the image is small
    long long slabs = (4ll * sm_count() + bx - 1) / bx;
    if (slabs < 1) slabs = 1;
    if (slabs > n_images) slabs = n_images;
    if (slabs > 65535) slabs = 65535;
    const int per_slab = static_cast<int>((n_images + slabs - 1) / slabs);
    const unsigned by = static_cast<unsigned>((n_images + per_slab - 1) / per_slab);
#define IM2IM_MM_LAUNCH(H)                                                                                     \
    miss_map_kernel<H><<<dim3(bx, by), 256, 0, st>>>(d_lower, d_pred, d_upper, d_label, n_images, px, stride_lower, \
                                                     stride_pred, stride_upper, stride_label, lam, per_slab, d_map)
    switch (head_kind) {
        case IM2IM_HEAD_QUANTILES: IM2IM_MM_LAUNCH(IM2IM_HEAD_QUANTILES); break;
        case IM2IM_HEAD_RESIDUAL: IM2IM_MM_LAUNCH(IM2IM_HEAD_RESIDUAL); break;
        case IM2IM_HEAD_GAUSSIAN: IM2IM_MM_LAUNCH(IM2IM_HEAD_GAUSSIAN); break;
        default: IM2IM_MM_LAUNCH(IM2IM_HEAD_SOFTMAX_SETS); break;
    }
#undef IM2IM_MM_LAUNCH
    return check_launch("miss_map_kernel");
}

extern "C" int im2im_fraction_missed_counts(const float* d_lower_edge, const float* d_upper_edge,
                                            const float* d_label, int64_t n_images, int64_t px,
                                            int64_t stride_lower, int64_t stride_upper, int64_t stride_label,
                                            int32_t* d_counts, void* stream) {
    if (n_images < 0 || px < 0) return fail(IM2IM_EINVAL, "negative size");
    if (n_images > 65535) return fail(IM2IM_ERANGE, "n_images=%lld > 65535 per call", (long long)n_images);
    if (n_images == 0) return IM2IM_OK;
    if (!d_counts) return fail(IM2IM_EINVAL, "null counts");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    IM2IM_CUDA_TRY(cudaMemsetAsync(d_counts, 0, sizeof(int32_t) * n_images, st));
    if (px == 0) return IM2IM_OK;
    if (!d_lower_edge || !d_upper_edge || !d_label) return fail(IM2IM_EINVAL, "null plane");
    long long bx = (px + 255) / 256;
    const long long want = (8ll * sm_count() + n_images - 1) / n_images;  // fill the GPU when there are few images
    if (bx > want) bx = want;
    if (bx < 1) bx = 1;
    fraction_missed_kernel<<<dim3(static_cast<unsigned>(bx), static_cast<unsigned>(n_images)), 256, 0, st>>>(
        d_lower_edge, d_upper_edge, d_label, px, stride_lower, stride_upper, stride_label, d_counts);
    return check_launch("fraction_missed_kernel");
}

extern "C" int im2im_softmax_sets(const float* d_logits, int64_t n_images, int32_t n_classes, int64_t inner,
                                  int64_t stride_image, int64_t stride_class, float* d_sets, void* stream) {
    if (n_images < 0 || inner < 0) return fail(IM2IM_EINVAL, "negative size");
    if (n_classes < 1 || n_classes > kMaxSoftmax)
        return fail(IM2IM_ERANGE, "n_classes=%d outside [1, %d]", n_classes, kMaxSoftmax);
    if (n_images == 0 || inner == 0) return IM2IM_OK;
    if (!d_logits || !d_sets) return fail(IM2IM_EINVAL, "null logits/sets");
    const long long n = static_cast<long long>(n_images) * inner;
    const long long blocks = (n + 255) / 256;
    const long long cap = 8ll * sm_count();
    softmax_sets_kernel<kMaxSoftmax><<<static_cast<unsigned>(blocks < cap ? blocks : cap), 256, 0,
                                       static_cast<cudaStream_t>(stream)>>>(d_logits, n_images, n_classes, inner,
                                                                            stride_image, stride_class, d_sets);
    return check_launch("softmax_sets_kernel");
}

extern "C" int im2im_rcps_decide_p2p(const unsigned long long* d_local_totals, unsigned long long* const* d_peer_mailboxes,
                                     unsigned* const* d_peer_flags, unsigned* d_epoch, int32_t rank, int32_t world,
                                     int32_t n_lambdas, double n_images_times_px, double gamma, double alpha32,
                                     double r_lo, double r_hi, double slack, unsigned long long* d_totals_out,
                                     int32_t* d_result, void* stream) {
    if (n_lambdas < 1 || n_lambdas > IM2IM_RCPS_MAX_LAMBDAS) return fail(IM2IM_ERANGE, "n_lambdas=%d", n_lambdas);
    if (world < 1 || world > 64 || rank < 0 || rank >= world) return fail(IM2IM_EINVAL, "bad rank/world (%d/%d)", rank, world);
    if (!d_local_totals || !d_peer_mailboxes || !d_peer_flags || !d_epoch || !d_totals_out || !d_result)
        return fail(IM2IM_EINVAL, "decide_p2p: null pointer");
    if (!(n_images_times_px > 0)) return fail(IM2IM_EINVAL, "empty calibration set");
    rcps_decide_p2p_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(
        d_local_totals, d_peer_mailboxes, d_peer_flags, d_epoch, rank, world, n_lambdas, n_images_times_px, gamma, alpha32,
        r_lo, r_hi, slack, d_totals_out, d_result);
    return check_launch("rcps_decide_p2p_kernel");
}
