// UNet convolution path for sm_100a (Path B of DESIGN.md): 3x3 / 1x1 convolution as an implicit GEMM on tcgen05.
//
// Replaces the library convolutions behind the reference's trunk
//   core/models/trunks/unet_parts.py:16-21 (DoubleConv: conv3x3 -> BN -> ReLU, twice), :90 (OutConv 1x1)
// for inference (BatchNorm folded into weight/bias on the host, ReLU fused in the epilogue).
//
// GEMM view:  D[pixel, cout] = sum_{tap, cin} X[pixel shifted by tap, cin] * W[cout, tap, cin]
//   M tile = 128 output pixels = one TMA box (BW x BH x BB) of an NHWC bf16 activation tensor; the shifted box of a
//            tap is the same box at coordinates (w0+dx, h0+dy): out-of-range rows/columns are ZERO-FILLED by TMA,
//            which is exactly the convolution's zero padding - no im2col buffer, no halo handling.
//   N tile = BN output channels (<= 256), K step = 64 input channels of one tap (128 B rows, SWIZZLE_128B).
//   Channel concatenation (unet_parts.py:68 torch.cat([x2, x1])) = two tensor maps walked by one K loop.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected thread, tcgen05.mma kind::f16, fp32 accumulator
// in TMEM), warps 2-5 = epilogue (tcgen05.ld -> +bias -> ReLU -> bf16 -> global NHWC).
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"
#include "rcps_rank.cuh"

namespace im2im {
namespace {

constexpr int kConvThreads = 192;     // 6 warps
constexpr int kHaloThreads = 320;     // conv_halo_kernel: + 4 more epilogue warps
constexpr int kTileM = 128;           // output pixels per CTA
constexpr int kKStep = 64;            // bf16 channels per K step = 128 bytes = one swizzle row
constexpr int kUmmaK = 16;            // K of one tcgen05.mma for 16-bit inputs
constexpr int kATileBytes = kTileM * kKStep * 2;  // 16 KB

struct ConvParams {
    int taps;          // 9 (3x3, pad 1) or 1 (1x1)
    int c_in1, c_in2;  // channels of the first / second (concatenated) input; c_in2 may be 0
    int c_out;
    int B, H, W;
    int bw, bh, bb;    // TMA box extents along W, H, batch (bw*bh*bb == 128)
    int tiles_w, tiles_h, tiles_b;
    int bn;            // N tile
    int stages;
    int relu;
    int tf32;          // 1: fp32 activations / weights in shared memory, tcgen05 kind::tf32 (K step = 32 channels = 128 B)
    int kstep;         // channels per K step: 64 (bf16) or 32 (tf32) - one 128-byte swizzle row either way
    const float* bias;       // [c_out] or null
    __nv_bfloat16* out_bf16; // NHWC [B,H,W,c_out] or null
    float* out_f32;          // NHWC fp32 or null
    // persistent kernel, training forward: per-channel sum / sum of squares of the bf16 values that are stored, accumulated
    // into stat_sums[c] / stat_sums[c_out + c] (fp32 atomics) - BatchNorm's batch statistics without a pass over the output
    float* stat_sums;
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp32 data read as TF32 (10-bit mantissa; low 13 bits ignored), fp32 accumulate.
// K of one instruction = 8 elements = 32 bytes, so the descriptor arithmetic of a 128-byte K step is the bf16 one.
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// round-to-nearest (ties away) onto the TF32 grid: what is stored is exactly what the next layer's MMA will read
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// One lane of a converged warp.  The MMA-issuing warps run their loops with ALL 32 lanes (warp-uniform control flow, the
// warp index taken through a shuffle so that the compiler knows it is uniform) and predicate only the tcgen05.mma /
// tcgen05.commit instructions with this: the shared-memory descriptors then live in uniform registers.  Issued from inside
// an `if (lane == 0)` region instead, every UTCHMMA was wrapped in an ELECT + 5x R2UR.BROADCAST + branch loop (the compiler
// must move per-thread registers into uniform ones one lane at a time) - ~75 cycles per instruction, which capped the
// N = 64 (32-cycle) and N = 32 MMAs of the halo kernel at 40 % / 20 % of the tensor pipe.
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ int uniform_warp_index() { return __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0); }
// mbarrier arrives when all previously issued tcgen05 ops of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// ---- CTA pair (cluster of two CTAs on one TPC, tcgen05 cta_group::2): M = 256 MMAs whose A rows and accumulator lanes are
// split between the two CTAs (128 each) and whose B rows are split too (N/2 per CTA), so a CTA reads 4 KB of A + half the
// B bytes per K = 16 step - 160 B/cycle instead of 192 for N = 64.  The leader (cluster rank 0) issues every MMA; both CTAs
// issue their own TMA loads, whose bytes are counted on the LEADER's mbarrier.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n"
                 "barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on a barrier of the cluster.  Relaxed on purpose: what the arrival hands over is a TMEM accumulator, ordered by
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync; `.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR, i.e. every
// tile's hand-over waited for the previous tile's global stores to drain (measured: 64->64 @320^2 0.55 -> 0.86 ms).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1,
                                                 int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t cols) {   // one warp of EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once all earlier MMAs of this thread are complete) on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(static_cast<uint16_t>(3))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 32-byte global store (STG.256, sm_100): an epilogue thread owns a pixel's channel run, and the lanes of a warp are 128+
// bytes apart, so every store instruction touches 32 separate sectors - with 16-byte stores each 32-byte sector was written
// half at a time by two instructions; p must be 32-byte aligned
__device__ __forceinline__ void st_global_256(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
                 "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
__device__ __forceinline__ uint4 pack8_bf16(const float* f) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
    pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
    return pk;
}

// Shared-memory matrix descriptor for a K-major, SWIZZLE_128B tile whose rows are 128 bytes (64 bf16):
// start address (>>4), LBO = 1 (ignored for swizzled K-major), SBO = 1024 B between 8-row groups, version 1 (sm_100),
// layout type 2 (SWIZZLE_128B).  The tile base must be 1024-byte aligned (base_offset = 0).
__device__ __forceinline__ uint64_t make_sw128_desc(const void* smem_ptr) {
    const uint32_t addr = smem_u32(smem_ptr);
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
    desc |= static_cast<uint64_t>(1) << 16;            // leading byte offset (16 B units)
    desc |= static_cast<uint64_t>(1024 >> 4) << 32;    // stride byte offset
    desc |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
    desc |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
    return desc;
}
// Instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = bn
__device__ __forceinline__ uint32_t make_idesc_bf16(int bn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(bn >> 3) << 17) |
           (static_cast<uint32_t>(kTileM >> 4) << 24);
}

// Instruction descriptor for kind::tf32: D = F32, A = B = TF32 (format 2), both K-major, M = 128, N = bn
__device__ __forceinline__ uint32_t make_idesc_tf32(int bn) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(bn >> 3) << 17) |
           (static_cast<uint32_t>(kTileM >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                  const __grid_constant__ CUtensorMap map_w, const ConvParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // layout: [stages x (A tile 16 KB | B tile bn*128 B)] [full barriers] [empty barriers] [accum barrier] [tmem ptr]
    const int b_tile_bytes = p.bn * kKStep * 2;
    const int stage_bytes = kATileBytes + b_tile_bytes;
    // SWIZZLE_128B tiles need 1024-byte alignment; the runtime only guarantees 16 for dynamic shared memory
    unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + static_cast<size_t>(p.stages) * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* accum_bar = empty_bar + p.stages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = uniform_warp_index();
    const int lane = threadIdx.x & 31;
    const uint32_t tmem_cols = p.bn < 32 ? 32u : static_cast<uint32_t>(p.bn);  // power of two >= 32 (bn in {32,64,128,256})

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_x1);
        if (p.c_in2 > 0) prefetch_tmap(&map_x2);
        prefetch_tmap(&map_w);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // tile coordinates
    int t = blockIdx.x;
    const int tw = t % p.tiles_w; t /= p.tiles_w;
    const int th = t % p.tiles_h; t /= p.tiles_h;
    const int tb = t;
    const int w0 = tw * p.bw, h0 = th * p.bh, b0 = tb * p.bb;
    const int n0 = blockIdx.y * p.bn;
    const int c_in = p.c_in1 + p.c_in2;
    const int kblocks_per_tap = c_in / p.kstep;
    const int n_k = p.taps * kblocks_per_tap;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        {   // all lanes walk the loop (uniform control flow: TMA operands stay in uniform registers), one lane issues
            const bool leader = elect_one_sync();
            int stage = 0;
            unsigned phase = 1;  // fresh barriers: waiting on parity 1 passes immediately
            for (int it = 0; it < n_k; ++it) {
                const int tap = it / kblocks_per_tap;
                const int cblk = it - tap * kblocks_per_tap;
                const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
                const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
                mbar_wait(&empty_bar[stage], phase);
                unsigned char* a_dst = tiles + static_cast<size_t>(stage) * stage_bytes;
                unsigned char* b_dst = a_dst + kATileBytes;
                const int c0 = cblk * p.kstep;
                if (leader) {
                    mbar_arrive_expect_tx(&full_bar[stage], static_cast<unsigned>(stage_bytes));
                    if (c0 < p.c_in1) tma_load_4d(a_dst, &map_x1, &full_bar[stage], c0, w0 + dx, h0 + dy, b0);
                    else tma_load_4d(a_dst, &map_x2, &full_bar[stage], c0 - p.c_in1, w0 + dx, h0 + dy, b0);
                    tma_load_2d(b_dst, &map_w, &full_bar[stage], tap * c_in + c0, n0);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (all lanes walk, one issues)
        const bool leader = elect_one_sync();
        const uint32_t idesc = p.tf32 ? make_idesc_tf32(p.bn) : make_idesc_bf16(p.bn);
        int stage = 0;
        unsigned phase = 0;
        for (int it = 0; it < n_k; ++it) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const unsigned char* a_src = tiles + static_cast<size_t>(stage) * stage_bytes;
            const uint64_t desc_a = make_sw128_desc(a_src);
            const uint64_t desc_b = make_sw128_desc(a_src + kATileBytes);
#pragma unroll
            for (int k = 0; k < kKStep / kUmmaK; ++k) {
                // advance 32 bytes (16 bf16 / 8 tf32 elements) along K inside the 128-byte swizzle row: +2 in 16-byte units
                if (leader) {
                    if (p.tf32) umma_tf32(tmem_base, desc_a + static_cast<uint64_t>(2 * k), desc_b + static_cast<uint64_t>(2 * k),
                                          idesc, (it > 0 || k > 0) ? 1u : 0u);
                    else umma_bf16(tmem_base, desc_a + static_cast<uint64_t>(2 * k), desc_b + static_cast<uint64_t>(2 * k),
                                   idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
            }
            if (leader) umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        if (leader) umma_commit(accum_bar);  // accumulator complete
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue warps (2..5)
        const int quad = warp & 3;               // TMEM lane quadrant this warp may access
        const int row = quad * 32 + lane;        // row of the 128-pixel tile
        int r = row;
        const int iw = r % p.bw; r /= p.bw;
        const int ih = r % p.bh; r /= p.bh;
        const int ib = r;
        const int w = w0 + iw, h = h0 + ih, b = b0 + ib;
        const bool in_range = (w < p.W) && (h < p.H) && (b < p.B);
        const size_t pix = (static_cast<size_t>(b) * p.H + h) * p.W + w;
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        for (int c = 0; c < p.bn; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c), v);
            tmem_ld_wait();
            if (in_range) {
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float x = __uint_as_float(v[j]);
                    if (p.bias) x += __ldg(p.bias + n0 + c + j);
                    if (p.relu) x = fmaxf(x, 0.f);
                    if (p.tf32) x = round_tf32(x);
                    f[j] = x;
                }
                if (p.out_bf16) {
                    uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + pix * p.c_out + n0 + c);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 pk;
                        __nv_bfloat162 h0_ = __floats2bfloat162_rn(f[8 * q + 0], f[8 * q + 1]);
                        __nv_bfloat162 h1_ = __floats2bfloat162_rn(f[8 * q + 2], f[8 * q + 3]);
                        __nv_bfloat162 h2_ = __floats2bfloat162_rn(f[8 * q + 4], f[8 * q + 5]);
                        __nv_bfloat162 h3_ = __floats2bfloat162_rn(f[8 * q + 6], f[8 * q + 7]);
                        pk.x = *reinterpret_cast<uint32_t*>(&h0_); pk.y = *reinterpret_cast<uint32_t*>(&h1_);
                        pk.z = *reinterpret_cast<uint32_t*>(&h2_); pk.w = *reinterpret_cast<uint32_t*>(&h3_);
                        dst[q] = pk;
                    }
                }
                if (p.out_f32) {
                    float4* dst = reinterpret_cast<float4*>(p.out_f32 + pix * p.c_out + n0 + c);
#pragma unroll
                    for (int q = 0; q < 8; ++q) dst[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------ persistent kernel
// Same tile math as conv_igemm_kernel, but one CTA per SM walks many tiles and the three roles run decoupled:
// the TMA ring and the MMA issuer run ahead into the next tile while the epilogue warps drain the previous
// accumulator (two TMEM accumulator buffers of bn columns each).  This removes the per-tile launch / barrier-init /
// TMEM-alloc / pipeline-fill cost that dominates the short-K (64-channel, 320x320) layers.
__global__ void __launch_bounds__(kConvThreads, 1)
conv_igemm_persistent_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                             const __grid_constant__ CUtensorMap map_w, const ConvParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int b_tile_bytes = p.bn * kKStep * 2;
    const int stage_bytes = kATileBytes + b_tile_bytes;
    unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + static_cast<size_t>(p.stages) * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* tmem_full = empty_bar + p.stages;   // [2] accumulator ready for the epilogue
    uint64_t* tmem_empty = tmem_full + 2;         // [2] accumulator drained, MMA may overwrite
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = uniform_warp_index();
    const int lane = threadIdx.x & 31;
    const uint32_t acc_cols = p.bn < 32 ? 32u : static_cast<uint32_t>(p.bn);
    const uint32_t tmem_cols = 2 * acc_cols;  // power of two: 64..512

    if (threadIdx.x == 0) {
        prefetch_tmap(&map_x1);
        if (p.c_in2 > 0) prefetch_tmap(&map_x2);
        prefetch_tmap(&map_w);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 4);  // one arrive per epilogue warp
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int n_tiles_n = p.c_out / p.bn;
    const int n_tiles_m = p.tiles_w * p.tiles_h * p.tiles_b;
    const int n_tiles = n_tiles_m * n_tiles_n;
    const int c_in = p.c_in1 + p.c_in2;
    const int kblocks_per_tap = c_in / p.kstep;
    const int n_k = p.taps * kblocks_per_tap;

    if (warp == 0) {
        {   // all lanes walk the loops (uniform control flow: TMA operands stay in uniform registers), one lane issues
            const bool leader = elect_one_sync();
            int stage = 0;
            unsigned phase = 1;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int tn = tile % n_tiles_n;
                int t = tile / n_tiles_n;
                const int tw = t % p.tiles_w; t /= p.tiles_w;
                const int th = t % p.tiles_h; t /= p.tiles_h;
                const int w0 = tw * p.bw, h0 = th * p.bh, b0 = t * p.bb, n0 = tn * p.bn;
                for (int it = 0; it < n_k; ++it) {
                    const int tap = it / kblocks_per_tap;
                    const int cblk = it - tap * kblocks_per_tap;
                    const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
                    const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
                    mbar_wait(&empty_bar[stage], phase);
                    unsigned char* a_dst = tiles + static_cast<size_t>(stage) * stage_bytes;
                    const int c0 = cblk * p.kstep;
                    if (leader) {
                        mbar_arrive_expect_tx(&full_bar[stage], static_cast<unsigned>(stage_bytes));
                        if (c0 < p.c_in1) tma_load_4d(a_dst, &map_x1, &full_bar[stage], c0, w0 + dx, h0 + dy, b0);
                        else tma_load_4d(a_dst, &map_x2, &full_bar[stage], c0 - p.c_in1, w0 + dx, h0 + dy, b0);
                        tma_load_2d(a_dst + kATileBytes, &map_w, &full_bar[stage], tap * c_in + c0, n0);
                    }
                    __syncwarp();
                    if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // all 32 lanes walk the loops (uniform control flow); the elected lane issues - see elect_one_sync()
        const bool leader = elect_one_sync();
        const uint32_t idesc = p.tf32 ? make_idesc_tf32(p.bn) : make_idesc_bf16(p.bn);
        int stage = 0;
        unsigned phase = 0;
        unsigned acc_phase = 3u;  // bit b = parity to wait for on tmem_empty[b]; fresh barriers: parity 1 passes
        int buf = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            mbar_wait(&tmem_empty[buf], (acc_phase >> buf) & 1u);
            acc_phase ^= 1u << buf;
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf) * acc_cols;
            for (int it = 0; it < n_k; ++it) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const unsigned char* a_src = tiles + static_cast<size_t>(stage) * stage_bytes;
                const uint64_t desc_a = make_sw128_desc(a_src);
                const uint64_t desc_b = make_sw128_desc(a_src + kATileBytes);
                if (p.tf32) {
#pragma unroll
                    for (int k = 0; k < kKStep / kUmmaK; ++k)
                        if (leader)
                            umma_tf32(tmem_d, desc_a + static_cast<uint64_t>(2 * k), desc_b + static_cast<uint64_t>(2 * k),
                                      idesc, (it > 0 || k > 0) ? 1u : 0u);
                } else {
#pragma unroll
                    for (int k = 0; k < kKStep / kUmmaK; ++k)
                        if (leader)
                            umma_bf16(tmem_d, desc_a + static_cast<uint64_t>(2 * k), desc_b + static_cast<uint64_t>(2 * k),
                                      idesc, (it > 0 || k > 0) ? 1u : 0u);
                }
                if (leader) umma_commit(&empty_bar[stage]);
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            if (leader) umma_commit(&tmem_full[buf]);
            __syncwarp();
            buf ^= 1;
        }
    } else {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;
        int r = row;
        const int iw = r % p.bw; r /= p.bw;
        const int ih = r % p.bh; r /= p.bh;
        const int ib = r;
        unsigned full_phase = 0u;  // bit b = parity to wait for on tmem_full[b]
        int buf = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int tn = tile % n_tiles_n;
            int t = tile / n_tiles_n;
            const int tw = t % p.tiles_w; t /= p.tiles_w;
            const int th = t % p.tiles_h; t /= p.tiles_h;
            const int w = tw * p.bw + iw, h = th * p.bh + ih, b = t * p.bb + ib, n0 = tn * p.bn;
            const bool in_range = (w < p.W) && (h < p.H) && (b < p.B);
            const size_t pix = (static_cast<size_t>(b) * p.H + h) * p.W + w;
            mbar_wait(&tmem_full[buf], (full_phase >> buf) & 1u);
            full_phase ^= 1u << buf;
            tc_fence_after();
            const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(buf) * acc_cols + (static_cast<uint32_t>(quad * 32) << 16);
            for (int c = 0; c < p.bn; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_acc + static_cast<uint32_t>(c), v);
                tmem_ld_wait();
                if (c + 32 >= p.bn) {  // last read of this accumulator: hand it back to the MMA warp early
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                }
                if (in_range) {
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = __uint_as_float(v[j]);
                        if (p.bias) x += __ldg(p.bias + n0 + c + j);
                        if (p.relu) x = fmaxf(x, 0.f);
                        if (p.tf32) x = round_tf32(x);   // stored activations are exactly what the next kind::tf32 MMA reads
                        f[j] = x;
                    }
                    if (p.out_bf16) {
                        __nv_bfloat16* dst = p.out_bf16 + pix * p.c_out + n0 + c;   // 64-byte aligned (c_out, n0, c: x32)
#pragma unroll
                        for (int q = 0; q < 2; ++q) st_global_256(dst + 16 * q, pack8_bf16(f + 16 * q), pack8_bf16(f + 16 * q + 8));
                    }
                    if (p.out_f32) {
                        float* dst = p.out_f32 + pix * p.c_out + n0 + c;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            st_global_256(dst + 8 * q, *reinterpret_cast<const uint4*>(f + 8 * q), *reinterpret_cast<const uint4*>(f + 8 * q + 4));
                    }
                }
                if (p.stat_sums != nullptr) {
                    // Column sums over the 32 pixel rows of this warp by recursive halving: at distance o a lane keeps the
                    // half of its columns selected by bit o of its lane index and adds the partner's copy of that half -
                    // 16+8+4+2+1 = 31 shuffles per statistic instead of 5 per column; lane j ends with column c + j.
                    // Values as stored (bf16-rounded); rows outside the image contribute zero.
                    float s0[32], s1[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float x = __uint_as_float(v[j]);
                        if (p.bias) x += __ldg(p.bias + n0 + c + j);
                        if (p.relu) x = fmaxf(x, 0.f);
                        x = in_range ? __bfloat162float(__float2bfloat16_rn(x)) : 0.f;
                        s0[j] = x;
                        s1[j] = x * x;
                    }
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const bool upper = (lane & o) != 0;
#pragma unroll
                        for (int i = 0; i < o; ++i) {
                            const float send0 = upper ? s0[i] : s0[i + o], keep0 = upper ? s0[i + o] : s0[i];
                            const float send1 = upper ? s1[i] : s1[i + o], keep1 = upper ? s1[i + o] : s1[i];
                            s0[i] = keep0 + __shfl_xor_sync(0xffffffffu, send0, o);
                            s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, o);
                        }
                    }
                    atomicAdd(p.stat_sums + n0 + c + lane, s0[0]);
                    atomicAdd(p.stat_sums + p.c_out + n0 + c + lane, s1[0]);
                }
            }
            buf ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, tmem_cols);
}

// ------------------------------------------------------------------------------------------------ halo kernel
// For the wide, shallow layers (64/128 channels at 320x320 / 160x160) the persistent kernel above is bound by L2->SM
// traffic, not by the tensor pipe: every 128-pixel tile re-fetches its activation box once per tap (9 x 16 KB) plus the
// weight slab.  Here a tile is 8 x 16 output pixels and ONE halo box (16 wide x 18 tall pixel rows of 64 channels,
// 36 KB, origin (w0-1, h0-1), TMA zero fill = padding) is loaded per 64-channel block; the nine taps are nine UMMA
// descriptors into that same box: start = base + ((dy+1)*16 + (dx+1)) * 128 B, 8-row core groups = one output row
// (8 consecutive pixels), SBO = 16 rows * 128 B = 2048 B.  The hardware applies the 128B swizzle on absolute shared-memory
// address bits, so the shifted (not 1024-byte aligned) starts read the TMA-written box correctly with base_offset = 0.  The layer's weights (9 * c_in * bn * 2 B <= 144 KB)
// are loaded ONCE per CTA and stay resident in shared memory across the persistent tile loop.
struct HaloParams {
    int c_in1, c_in2, c_out;
    int B, H, W;
    int tiles_w, tiles_h;   // 8-wide, 16-tall tiles
    int bn;
    int a_stages;
    // 0: the layer's weights (this CTA's rows of the N block, all taps and channel blocks) stay resident in shared memory;
    // 1: they do not fit - every ring stage carries the nine tap slabs of its channel block behind the halo box
    int w_stream;
    // 1: fp32 activations / weights, kind::tf32 (a 128-byte K block is 32 channels), fp32 output rounded to TF32; plain epilogue only
    int tf32;
    int relu;
    const float* bias;
    __nv_bfloat16* out_bf16;
    float* out_f32;
    // head mode: only the first n_real (<= 32) output channels are real; they are written as fp32 PLANES
    // [B, n_real, H, W] (the reference's (B, planes, C, H, W) head tensor), planes >= act_from get act_kind (1 relu, 2 abs)
    float* out_planar;
    int n_real, act_kind, act_from;
    // head mode + calibration histogram (quantile head, one output channel: planes = lower, prediction, upper): every pixel
    // is ranked against the sorted lambda grid exactly as rcps_hist_kernel ranks it (rcps_rank.cuh) from the fp32 values
    // that out_planar would hold, and booked into hist_global[b][k] (u32 [B][hist_L + 1], accumulated into) - the
    // (B, 3, 1, H, W) head tensor need not exist (out_planar may be null).  Tiles are then dealt to the CTAs in contiguous
    // runs so that a CTA flushes its shared-memory histogram once per image it touches.
    // head mode, optional (a 1x1 convolution with a bias folded into this head's weights): tap_bias[cls * n_real + j] = the
    // part of bias[j] that came through the taps of the 3x3 window that lie in the zero padding for a pixel of border class
    // cls = 3 * (0 inside | 1 top row | 2 bottom row) + (0 inside | 1 left column | 2 right column) - where the folded-in
    // convolution's output, bias included, does not exist; such pixels subtract it.
    const float* tap_bias;
    // plain bf16 epilogue, optional: also write maxpool2x2 of the (bias + ReLU'd, bf16-rounded) output, NHWC
    // [B, H/2, W/2, c_out] - the 2x2 window of a pixel lives in lanes l, l^1, l^8 of its epilogue warp (tile rows are 8 wide)
    __nv_bfloat16* pool_out;
    unsigned* hist_global;
    const float* hist_labels;      // fp32 [B, 1, H, W]
    const float* hist_lambdas;     // device, ascending, hist_L entries
    int hist_L;
    // Per-channel statistics can be accumulated on the way out (training), from the bf16 values that are stored; each
    // epilogue thread keeps the partial sums of its pixel row for all 64 channels in registers until the end of the kernel:
    //   stat_mode 1  sums[c] += z, sums[C+c] += z*z                         BatchNorm batch statistics of this conv's output
    //   stat_mode 2  this conv is a data gradient producing dy for a BatchNorm+ReLU layer whose pre-normalisation output is
    //                bn_z: g = dy * (bn_z*scale + shift > 0) is stored INSTEAD of dy, and
    //                sums[c] += g, sums[C+c] += rstd*(sum(g*z) - mean*sum(g))   (what channel_reduce_kernel<1> computes)
    int debug_skip_store;          // IM2IM_HALO_SKIP_STORE=1: compute, do not store (upper bound of an ideal epilogue; dev only)
    int stat_mode;
    float* stat_sums;              // fp32 [2 * c_out], accumulated into
    const __nv_bfloat16* bn_z;     // stat_mode 2: NHWC [B,H,W,c_out]
    const float* bn_gamma; const float* bn_beta; const float* bn_mean; const float* bn_rstd;
};

constexpr int kMaxFoldedPlanes = 7;    // head mode with tap_bias: 9 border classes x planes fit the 64-float scale/shift area
constexpr int kHaloStatBn = 64;      // fused statistics need the per-thread accumulators in registers: N tile of 64

constexpr int kHaloW = 16, kHaloH = 18, kHaloTileW = 8, kHaloTileH = 16;
constexpr int kHaloBytes = kHaloW * kHaloH * kKStep * 2;  // 36864

__device__ __forceinline__ uint64_t make_sw128_desc_halo(const void* smem_ptr) {
    const uint32_t addr = smem_u32(smem_ptr);
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
    desc |= static_cast<uint64_t>(1) << 16;                       // LBO (ignored, swizzled K-major)
    desc |= static_cast<uint64_t>((kHaloW * 128) >> 4) << 32;     // SBO: next output row = 16 pixel rows further
    desc |= static_cast<uint64_t>(1) << 46;                       // descriptor version
    // base_offset stays 0: measured on B200 (both settings were run through tools/conv_check.py) - the 128B swizzle is applied on absolute shared-memory
    // address bits, so a start that is not 1024-byte aligned reads TMA-written data correctly as is; setting
    // base_offset = (addr >> 7) & 7 breaks every tap with dx != -1.
    desc |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B
    return desc;
}

// kPair: the kernel runs as clusters of two CTAs (a TPC's two SMs) that work on two neighbouring pixel tiles with M = 256
// cta_group::2 MMAs: each CTA keeps only HALF of the weight rows resident (rows [rank * bn/2, +bn/2) of the N block) and
// reads its own halo box, the leader issues the MMAs for both, every CTA drains its own 128 TMEM lanes.  Per K = 16 step a
// CTA's shared memory serves 4 KB of A + bn/2 rows of B instead of bn rows - the operand traffic that capped the N = 64
// layers at 67 % of the tensor pipe.
// kTf32: fp32 activations / weights through kind::tf32 (HaloParams::tf32).  A template parameter, not a run-time branch: with
// both instruction kinds in one MMA loop the compiler if-converts them into @UP / @!UP pairs of UTCHMMA and the nullified
// halves still cost issue time (measured: the bf16 forward's halo launches 4.23 -> 5.56 ms at batch 78).
template <bool kPair, bool kTf32 = false>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                 const __grid_constant__ CUtensorMap map_w, const HaloParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int c_in = p.c_in1 + p.c_in2;
    constexpr int kch = kTf32 ? kKStep / 2 : kKStep;            // channels per 128-byte K block
    const int cblocks = c_in / kch;
    const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
    const int bn_local = kPair ? p.bn / 2 : p.bn;              // weight rows resident in THIS CTA
    const int b_tile_bytes = bn_local * kKStep * 2;
    const bool w_stream = p.w_stream != 0;
    const int w_bytes = w_stream ? 0 : 9 * cblocks * b_tile_bytes;   // resident weights: [tap][cblock][bn rows x 128 B]
    // ring stage: one 36 KB halo box (+ streamed weights: the nine tap slabs of the stage's channel block)
    const int stage_bytes = kHaloBytes + (w_stream ? 9 * b_tile_bytes : 0);
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* w_smem = base;
    unsigned char* a_ring = base + w_bytes;                    // a_stages stages
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(a_ring + static_cast<size_t>(p.a_stages) * stage_bytes);
    uint64_t* empty_bar = full_bar + p.a_stages;
    uint64_t* w_bar = empty_bar + p.a_stages;
    uint64_t* tmem_full = w_bar + 1;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    // behind the barrier block (128 B reserved): [scale bn f32][shift bn f32] for stat_mode 2
    float* s_scale = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(full_bar) + 128);
    float* s_shift = s_scale + p.bn;
    // histogram mode: [lambda pairs f32x2 x (L+1)][lambda f32 x L][hist u32 x (L+1)] behind them
    float2* s_pair = reinterpret_cast<float2*>(s_shift + p.bn);
    float* s_lam = reinterpret_cast<float*>(s_pair + (p.hist_L + 1));
    unsigned* s_hist = reinterpret_cast<unsigned*>(s_lam + p.hist_L);

    const int warp = uniform_warp_index();
    const int lane = threadIdx.x & 31;
    const uint32_t acc_cols = p.bn < 32 ? 32u : static_cast<uint32_t>(p.bn);
    const uint32_t tmem_cols = 2 * acc_cols;
    // Eight epilogue warps: warps 2-5 drain the first half of the accumulator columns, warps 6-9 the second half (a warp can
    // only read the TMEM lanes of its quadrant, so two warps share each quadrant).  With one 36 KB halo box per tile the
    // epilogue (TMEM -> registers -> bf16 -> global), not the 36 MMAs, set the pace of the 64-channel layers.  Head mode and
    // N = 32 tiles have too few columns to split: warps 6-9 idle there.
    const bool head_mode = (p.out_planar != nullptr || p.hist_global != nullptr);
    const bool split_epilogue = !head_mode && p.bn >= 64;

    if (p.stat_mode == 2) {   // the forward's per-channel affine, recomputed with the same ops (bn_relu_bwd kernels do the same)
        for (int c = threadIdx.x; c < p.bn; c += blockDim.x) {
            const int ch = blockIdx.y * p.bn + c;
            const float sc = p.bn_gamma[ch] * p.bn_rstd[ch];
            s_scale[c] = sc;
            s_shift[c] = p.bn_beta[ch] - p.bn_mean[ch] * sc;
        }
    }
    // plain epilogue: the N block's bias staged once (read back as float4 broadcasts).  32 scalar __ldg per thread and tile
    // cost the layers whose epilogue sets the pace (ncu, 64 -> 128 @160^2: LSU wavefronts 67 %, tensor pipe 64 %)
    const bool bias_staged = p.bias != nullptr && !head_mode && p.stat_mode == 0;
    if (bias_staged)
        for (int c = threadIdx.x; c < p.bn; c += blockDim.x) s_scale[c] = p.bias[blockIdx.y * p.bn + c];
    if (p.tap_bias != nullptr)   // 9 border classes x n_real <= 2 * bn floats (checked by the launcher)
        for (int j = threadIdx.x; j < 9 * p.n_real; j += blockDim.x) s_scale[j] = p.tap_bias[j];
    if (p.hist_global != nullptr) {
        const int L = p.hist_L;
        for (int j = threadIdx.x; j <= L; j += blockDim.x) {
            s_hist[j] = 0u;
            s_pair[j] = make_float2(j > 0 ? p.hist_lambdas[j - 1] : -INFINITY, j < L ? p.hist_lambdas[j] : INFINITY);
            if (j < L) s_lam[j] = p.hist_lambdas[j];
        }
    }
    if (threadIdx.x == 0) {
        prefetch_tmap(&map_x1);
        if (p.c_in2 > 0) prefetch_tmap(&map_x2);
        prefetch_tmap(&map_w);
        for (int s = 0; s < p.a_stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(w_bar, 1);
        // pair: the leader's tmem_empty collects the epilogue warps of both CTAs (the other CTA's copy is not used)
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full[b], 1); mbar_init(&tmem_empty[b], (kPair ? 2 : 1) * (split_epilogue ? 8 : 4)); }
        mbar_fence_init();
    }
    if (kPair) { __syncthreads(); cluster_sync_all(); }   // the peer's barriers exist before anything is sent to them; both CTAs allocate together
    if (warp == 1) { if (kPair) tmem_alloc_pair(tmem_ptr_smem, tmem_cols); else tmem_alloc(tmem_ptr_smem, tmem_cols); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const int n_tiles_n = p.c_out / p.bn;
    const int n_tiles_m = p.tiles_w * p.tiles_h * p.B;
    // tiles of one N block are contiguous for a CTA so that the resident weights are loaded once: CTAs are split over
    // the N blocks first (blockIdx.y), then stride over the pixel tiles
    const int tn = blockIdx.y;
    const int n0 = tn * p.bn;
    // tile order: strided over the CTAs, or (histogram mode) one contiguous run per CTA
    // (pair: the loop variable counts tile PAIRS, the CTA's own tile is 2 * pair + rank; with an odd tile count the last
    // pair's second tile lies behind the batch: TMA zero-fills it, the epilogue skips it)
    int tile_first = blockIdx.x, tile_last = n_tiles_m, tile_step = gridDim.x;
    if (kPair) { tile_first = blockIdx.x >> 1; tile_last = (n_tiles_m + 1) >> 1; tile_step = gridDim.x >> 1; }
    if (p.hist_global != nullptr) {
        tile_first = static_cast<int>(static_cast<long long>(n_tiles_m) * blockIdx.x / gridDim.x);
        tile_last = static_cast<int>(static_cast<long long>(n_tiles_m) * (blockIdx.x + 1) / gridDim.x);
        tile_step = 1;
    }

    if (warp == 0) {
        {   // all lanes walk the loops (uniform control flow: TMA operands stay in uniform registers), one lane issues
            const bool leader = elect_one_sync();
            // resident weights, once (pair: this CTA's half of the rows; all bytes are counted on the leader's barrier)
            const uint32_t w_bar_lead = kPair ? mapa_u32(smem_u32(w_bar), 0u) : 0u;
            if (leader && cta_rank == 0 && !w_stream) mbar_arrive_expect_tx(w_bar, static_cast<unsigned>(w_bytes) * (kPair ? 2u : 1u));
            for (int tap = 0; tap < 9 && !w_stream; ++tap)
                for (int cb = 0; cb < cblocks; ++cb)
                    if (leader) {
                        unsigned char* wdst = w_smem + static_cast<size_t>(tap * cblocks + cb) * b_tile_bytes;
                        if (kPair) tma_load_2d_pair(wdst, &map_w, w_bar_lead, tap * c_in + cb * kch, n0 + static_cast<int>(cta_rank) * bn_local);
                        else tma_load_2d(wdst, &map_w, w_bar, tap * c_in + cb * kch, n0);
                    }
            __syncwarp();
            int stage = 0;
            unsigned phase = 1;
            for (int it = tile_first; it < tile_last; it += tile_step) {
                const int tile = kPair ? 2 * it + static_cast<int>(cta_rank) : it;
                int t = tile;
                const int tw = t % p.tiles_w; t /= p.tiles_w;
                const int th = t % p.tiles_h; t /= p.tiles_h;
                const int w0 = tw * kHaloTileW - 1, h0 = th * kHaloTileH - 1, b0 = t;
                for (int cb = 0; cb < cblocks; ++cb) {
                    mbar_wait(&empty_bar[stage], phase);
                    unsigned char* dst = a_ring + static_cast<size_t>(stage) * stage_bytes;
                    const int c0 = cb * kch;
                    if (leader) {
                        if (kPair) {
                            // both CTAs' bytes complete on the leader's barrier, which expects their sum
                            const uint32_t full_lead = mapa_u32(smem_u32(&full_bar[stage]), 0u);
                            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * static_cast<unsigned>(stage_bytes));
                            if (c0 < p.c_in1) tma_load_4d_pair(dst, &map_x1, full_lead, c0, w0, h0, b0);
                            else tma_load_4d_pair(dst, &map_x2, full_lead, c0 - p.c_in1, w0, h0, b0);
                        } else {
                            mbar_arrive_expect_tx(&full_bar[stage], static_cast<unsigned>(stage_bytes));
                            if (c0 < p.c_in1) tma_load_4d(dst, &map_x1, &full_bar[stage], c0, w0, h0, b0);
                            else tma_load_4d(dst, &map_x2, &full_bar[stage], c0 - p.c_in1, w0, h0, b0);
                        }
                    }
                    if (w_stream) {   // this channel block's nine tap slabs ride in the same stage
                        const uint32_t full_lead = kPair ? mapa_u32(smem_u32(&full_bar[stage]), 0u) : 0u;
                        for (int tap = 0; tap < 9; ++tap)
                            if (leader) {
                                unsigned char* wdst = dst + kHaloBytes + tap * b_tile_bytes;
                                if (kPair) tma_load_2d_pair(wdst, &map_w, full_lead, tap * c_in + c0, n0 + static_cast<int>(cta_rank) * bn_local);
                                else tma_load_2d(wdst, &map_w, &full_bar[stage], tap * c_in + c0, n0);
                            }
                    }
                    __syncwarp();
                    if (++stage == p.a_stages) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
      if (!kPair || cta_rank == 0) {     // pair: the leader CTA issues the MMAs of both
        // all 32 lanes walk the loops (uniform control flow); the elected lane issues - see elect_one_sync()
        const bool leader = elect_one_sync();
        // pair: M = 256 (bits 24-28 hold M >> 4)
        const uint32_t idesc = (kTf32 ? make_idesc_tf32(p.bn) : make_idesc_bf16(p.bn)) + (kPair ? (static_cast<uint32_t>(kTileM >> 4) << 24) : 0u);
        if (!w_stream) mbar_wait(w_bar, 0);
        int stage = 0;
        unsigned phase = 0;
        unsigned acc_phase = 3u;
        int buf = 0;
        for (int tile = tile_first; tile < tile_last; tile += tile_step) {
            mbar_wait(&tmem_empty[buf], (acc_phase >> buf) & 1u);
            acc_phase ^= 1u << buf;
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf) * acc_cols;
            for (int cb = 0; cb < cblocks; ++cb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const unsigned char* a_src = a_ring + static_cast<size_t>(stage) * stage_bytes;
                const unsigned char* w_src = w_stream ? a_src + kHaloBytes : w_smem + static_cast<size_t>(cb) * b_tile_bytes;
                const int w_tap_stride = w_stream ? b_tile_bytes : cblocks * b_tile_bytes;
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const int dy = tap / 3, dx = tap % 3;  // already offset by +1 (halo origin is (w0-1, h0-1))
                    const uint64_t desc_a = make_sw128_desc_halo(a_src + (dy * kHaloW + dx) * 128);
                    const uint64_t desc_b = make_sw128_desc(w_src + static_cast<size_t>(tap) * w_tap_stride);
#pragma unroll
                    for (int k = 0; k < kKStep / kUmmaK; ++k)
                        if (leader) {
                            // (kind::tf32: K = 8 fp32 = the same 32 bytes per instruction, so the descriptor steps are the same)
                            const uint64_t da = desc_a + static_cast<uint64_t>(2 * k), db = desc_b + static_cast<uint64_t>(2 * k);
                            const uint32_t acc = (cb > 0 || tap > 0 || k > 0) ? 1u : 0u;
                            if (kTf32) { if (kPair) umma_tf32_pair(tmem_d, da, db, idesc, acc); else umma_tf32(tmem_d, da, db, idesc, acc); }
                            else { if (kPair) umma_bf16_pair(tmem_d, da, db, idesc, acc); else umma_bf16(tmem_d, da, db, idesc, acc); }
                        }
                }
                if (leader) { if (kPair) umma_commit_pair(&empty_bar[stage]); else umma_commit(&empty_bar[stage]); }
                __syncwarp();
                if (++stage == p.a_stages) { stage = 0; phase ^= 1u; }
            }
            if (leader) { if (kPair) umma_commit_pair(&tmem_full[buf]); else umma_commit(&tmem_full[buf]); }
            __syncwarp();
            buf ^= 1;
        }
      }
    } else if (warp < 6 || split_epilogue) {
        const int quad = warp & 3;
        const int half = (warp - 2) >> 2;                    // 0: warps 2-5, 1: warps 6-9
        const int row = quad * 32 + lane;
        const int iw = row % kHaloTileW, ih = row / kHaloTileW;
        const int c_lo = split_epilogue ? half * (p.bn / 2) : 0;        // this warp's accumulator columns [c_lo, c_hi)
        const int c_hi = split_epilogue ? c_lo + p.bn / 2 : p.bn;
        unsigned full_phase = 0u;
        int buf = 0;
        // fused statistics (bn == 64 only): this thread's partial sums over its pixel row of every tile, for the 32 channels
        // [c_lo, c_lo + 32) of its warp
        float acc0[kHaloStatBn / 2], acc1[kHaloStatBn / 2];
#pragma unroll
        for (int j = 0; j < kHaloStatBn / 2; ++j) acc0[j] = acc1[j] = 0.f;
        // forward statistics of an N = 128 tile (64 columns per warp - too many for per-thread accumulators): the column
        // sums over the warp's 32 pixel rows are formed per tile by recursive halving and lane j keeps the running sums of
        // columns c_lo + j and c_lo + 32 + j
        float wide0[2] = {0.f, 0.f}, wide1[2] = {0.f, 0.f};
        // where this warp hands an accumulator back: the MMA warp's barrier (pair: the leader CTA's, through the cluster window)
        const uint32_t tmem_empty_lead0 = kPair ? mapa_u32(smem_u32(&tmem_empty[0]), 0u) : 0u;
        auto release_acc = [&](int which) {
            if (kPair) mbar_arrive_cluster(tmem_empty_lead0 + 8u * static_cast<uint32_t>(which));
            else mbar_arrive(&tmem_empty[which]);
        };
        for (int it = tile_first; it < tile_last; it += tile_step) {
            const int tile = kPair ? 2 * it + static_cast<int>(cta_rank) : it;
            int t = tile;
            const int tw = t % p.tiles_w; t /= p.tiles_w;
            const int th = t % p.tiles_h; t /= p.tiles_h;
            const int w = tw * kHaloTileW + iw, h = th * kHaloTileH + ih, b = t;
            const bool in_range = (w < p.W) && (h < p.H) && !p.debug_skip_store;
            const size_t pix = (static_cast<size_t>(b) * p.H + h) * p.W + w;
            mbar_wait(&tmem_full[buf], (full_phase >> buf) & 1u);
            full_phase ^= 1u << buf;
            tc_fence_after();
            if (kPair && tile >= n_tiles_m) {   // the odd tile out of the last pair: nothing to store, hand the accumulator back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(buf);
                buf ^= 1;
                continue;
            }
            const uint32_t tmem_acc = tmem_base + static_cast<uint32_t>(buf) * acc_cols + (static_cast<uint32_t>(quad * 32) << 16);
            if (p.out_planar != nullptr || p.hist_global != nullptr) {
                // head mode: the real outputs live in the first 32 accumulator columns; one TMEM load, then the buffer is free
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_acc, v);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release_acc(buf);
                if (p.tap_bias != nullptr && (tw == 0 || th == 0 || tw == p.tiles_w - 1 || th == p.tiles_h - 1)) {
                    // head with a 1x1 convolution folded in, tile on the image border (warp-uniform test): a border pixel takes
                    // back the share of the bias that belongs to the taps lying in the zero padding; the table (staged in shared
                    // memory, where the statistics modes keep their scale/shift) is indexed by the pixel's border class
                    const int cls = (h == 0 ? 3 : (h == p.H - 1 ? 6 : 0)) + (w == 0 ? 1 : (w == p.W - 1 ? 2 : 0));
                    if (cls != 0) {
                        const float* corr = s_scale + cls * p.n_real;
#pragma unroll
                        for (int j = 0; j < kMaxFoldedPlanes; ++j)
                            if (j < p.n_real) v[j] = __float_as_uint(__uint_as_float(v[j]) - corr[j]);
                    }
                }
                if (p.hist_global != nullptr) {
                    // planes 0..2 = lower, prediction, upper of this thread's pixel, as fp32 exactly as they would be stored
                    const int L = p.hist_L;
                    if (in_range) {
                        float o[3];
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            float x = __uint_as_float(v[j]);
                            if (p.bias) x += __ldg(p.bias + j);
                            if (j >= p.act_from) {
                                if (p.act_kind == 1) x = (x != x) ? x : fmaxf(x, 0.f);
                                else if (p.act_kind == 2) x = fabsf(x);
                            }
                            o[j] = x;
                        }
                        const float y = __ldg(p.hist_labels + pix);
                        const float lam0 = s_lam[0], span = s_lam[L - 1] - lam0;
                        const float guess_scale = (L > 1 && span > 0.f) ? static_cast<float>(L - 1) / span : 0.f;
                        const PixelQuery q = make_query<IM2IM_HEAD_QUANTILES>(o[0], o[1], o[2], y);
                        bool ok;
                        int k = rank_guess<IM2IM_HEAD_QUANTILES>(q, s_pair, L, guess_scale, -lam0 * guess_scale, ok);
                        if (!ok) k = rank_bisect(q.d, q.P, q.Y, s_lam, L);
                        if (k > 0) atomicAdd(&s_hist[k], 1u);
                    }
                    // last tile of this image in this CTA's run: move the shared histogram into the image's global row
                    const int per_image = p.tiles_w * p.tiles_h;
                    if (tile + 1 == tile_last || (tile + 1) / per_image != b) {
                        named_bar_sync(2, 128);
                        unsigned* grow = p.hist_global + static_cast<size_t>(b) * (L + 1);
                        for (int k = row; k <= L; k += 128) {
                            const unsigned c = s_hist[k];
                            if (c != 0u) { atomicAdd(grow + k, c); s_hist[k] = 0u; }
                        }
                        named_bar_sync(2, 128);
                    }
                }
                if (in_range && p.out_planar != nullptr) {
                    const size_t hw = static_cast<size_t>(p.H) * p.W;
                    float* dst = p.out_planar + static_cast<size_t>(b) * p.n_real * hw + static_cast<size_t>(h) * p.W + w;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < p.n_real) {
                            float x = __uint_as_float(v[j]);
                            if (p.bias) x += __ldg(p.bias + j);
                            if (j >= p.act_from) {
                                if (p.act_kind == 1) x = (x != x) ? x : fmaxf(x, 0.f);
                                else if (p.act_kind == 2) x = fabsf(x);
                            }
                            dst[static_cast<size_t>(j) * hw] = x;   // lanes = 8 consecutive pixels of a row: 32-byte runs
                        }
                    }
                }
                buf ^= 1;
                continue;
            }
            if (p.stat_mode == 1 && p.bn == 128) {
                // ---- bf16 output + forward statistics of an N = 128 tile (c_in >= 128: >= 4600 tensor cycles per tile pay
                // for the 2 x 62 shuffles per warp)
#pragma unroll
                for (int chunk = 0; chunk < 2; ++chunk) {
                    const int c = c_lo + 32 * chunk;
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tmem_acc + static_cast<uint32_t>(c), v);
                    tmem_ld_wait();
                    if (chunk == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) release_acc(buf);
                    }
                    float s0[32], s1[32];
                    uint4 pk[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        __nv_bfloat162 h2[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            h2[j] = __floats2bfloat162_rn(__uint_as_float(v[8 * q + 2 * j]), __uint_as_float(v[8 * q + 2 * j + 1]));
                            const float2 r = __bfloat1622float2(h2[j]);      // the values as stored
                            s0[8 * q + 2 * j] = r.x; s0[8 * q + 2 * j + 1] = r.y;
                            s1[8 * q + 2 * j] = r.x * r.x; s1[8 * q + 2 * j + 1] = r.y * r.y;
                        }
                        pk[q].x = *reinterpret_cast<uint32_t*>(&h2[0]); pk[q].y = *reinterpret_cast<uint32_t*>(&h2[1]);
                        pk[q].z = *reinterpret_cast<uint32_t*>(&h2[2]); pk[q].w = *reinterpret_cast<uint32_t*>(&h2[3]);
                    }
                    if (in_range) {
                        __nv_bfloat16* dst = p.out_bf16 + pix * p.c_out + n0 + c;
                        st_global_256(dst, pk[0], pk[1]);
                        st_global_256(dst + 16, pk[2], pk[3]);
                    }
                    // recursive halving over the 32 rows (see conv_igemm_persistent_kernel): lane j ends with column c + j
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const bool upper = (lane & o) != 0;
#pragma unroll
                        for (int i = 0; i < o; ++i) {
                            const float send0 = upper ? s0[i] : s0[i + o], keep0 = upper ? s0[i + o] : s0[i];
                            const float send1 = upper ? s1[i] : s1[i + o], keep1 = upper ? s1[i + o] : s1[i];
                            s0[i] = keep0 + __shfl_xor_sync(0xffffffffu, send0, o);
                            s1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, o);
                        }
                    }
                    wide0[chunk] += s0[0];
                    wide1[chunk] += s1[0];
                }
                buf ^= 1;
                continue;
            }
            if (p.stat_mode != 0) {
                // ---- bf16 output + fused per-channel statistics, registers only (bn == 64, 32 channels per warp): this
                // kernel is bound by shared-memory bandwidth (operand reads of N = 64 MMAs + the TMA fill), so the epilogue
                // must not touch shared memory - measured: staging the tile for coalesced / TMA stores made it slower
                const int c = c_lo;
                uint4 zreg[4];
                if (p.stat_mode == 2) {   // this thread's bn_z values (its pixel, 32 channels), requested in the shadow of the MMAs
                    const uint4* zsrc = reinterpret_cast<const uint4*>(p.bn_z + pix * p.c_out + n0 + c);
#pragma unroll
                    for (int q = 0; q < 4; ++q) zreg[q] = __ldg(zsrc + q);
                }
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_acc + static_cast<uint32_t>(c), v);
                tmem_ld_wait();
                tc_fence_before();          // only read of this accumulator by this warp: hand it back to the MMA warp
                __syncwarp();
                if (lane == 0) release_acc(buf);
                uint4* dst = reinterpret_cast<uint4*>(p.out_bf16 + pix * p.c_out + n0 + c);
                uint4 pk_even = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    __nv_bfloat162 h2[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        h2[j] = __floats2bfloat162_rn(__uint_as_float(v[8 * q + 2 * j]), __uint_as_float(v[8 * q + 2 * j + 1]));
                    if (p.stat_mode == 1) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 r = __bfloat1622float2(h2[j]);      // the values as stored
                            const int cc = 8 * q + 2 * j;
                            acc0[cc] += r.x; acc1[cc] = fmaf(r.x, r.x, acc1[cc]);
                            acc0[cc + 1] += r.y; acc1[cc + 1] = fmaf(r.y, r.y, acc1[cc + 1]);
                        }
                    } else {
                        const __nv_bfloat162* hz = reinterpret_cast<const __nv_bfloat162*>(&zreg[q]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 r = __bfloat1622float2(h2[j]);
                            const float2 z2 = __bfloat1622float2(hz[j]);
                            const int cc = 8 * q + 2 * j;
                            // ReLU mask exactly as the forward saw it; g replaces dy in the stored tensor
                            const float g0 = fmaf(z2.x, s_scale[c + cc], s_shift[c + cc]) > 0.f ? r.x : 0.f;
                            const float g1 = fmaf(z2.y, s_scale[c + cc + 1], s_shift[c + cc + 1]) > 0.f ? r.y : 0.f;
                            acc0[cc] += g0; acc1[cc] = fmaf(g0, z2.x, acc1[cc]);
                            acc0[cc + 1] += g1; acc1[cc + 1] = fmaf(g1, z2.y, acc1[cc + 1]);
                            h2[j] = __floats2bfloat162_rn(g0, g1);
                        }
                    }
                    uint4 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&h2[0]); pk.y = *reinterpret_cast<uint32_t*>(&h2[1]);
                    pk.z = *reinterpret_cast<uint32_t*>(&h2[2]); pk.w = *reinterpret_cast<uint32_t*>(&h2[3]);
                    if (q & 1) {
                        if (in_range) st_global_256(dst + (q - 1), pk_even, pk);   // 32-byte stores: whole sectors
                    } else {
                        pk_even = pk;
                    }
                }
                buf ^= 1;
                continue;
            }
            for (int c = c_lo; c < c_hi; c += 32) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_acc + static_cast<uint32_t>(c), v);
                tmem_ld_wait();
                if (c + 32 >= c_hi) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) release_acc(buf);
                }
                float f[32];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (bias_staged) bv = *reinterpret_cast<const float4*>(s_scale + c + 4 * j4);
                    const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        float x = __uint_as_float(v[4 * j4 + jj]);
                        if (bias_staged) x += bb[jj];
                        if (p.relu) x = fmaxf(x, 0.f);
                        if (kTf32) x = round_tf32(x);    // stored activations are exactly what the next kind::tf32 MMA reads
                        f[4 * j4 + jj] = x;
                    }
                }
                if (in_range) {
                    if (p.out_bf16) {
                        __nv_bfloat16* dst = p.out_bf16 + pix * p.c_out + n0 + c;   // 64-byte aligned (c_out, n0, c: x32)
#pragma unroll
                        for (int q = 0; q < 2; ++q) st_global_256(dst + 16 * q, pack8_bf16(f + 16 * q), pack8_bf16(f + 16 * q + 8));
                    }
                    if (p.out_f32) {
                        float* dst = p.out_f32 + pix * p.c_out + n0 + c;
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            st_global_256(dst + 8 * q, *reinterpret_cast<const uint4*>(f + 8 * q), *reinterpret_cast<const uint4*>(f + 8 * q + 4));
                    }
                }
                if (p.pool_out != nullptr) {
                    // 2x2 max-pool of the values as stored: max over lanes {l, l^1} (neighbour in the row) and {l, l^8} (the
                    // row below); rounding to bf16 is monotone, so this equals pooling the stored tensor (maxpool2x2_kernel)
                    uint4 pk[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) pk[q] = pack8_bf16(f + 8 * q);
                    uint32_t* r = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        uint32_t o = __shfl_xor_sync(0xffffffffu, r[i], 1);
                        __nv_bfloat162 m = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&r[i]), *reinterpret_cast<__nv_bfloat162*>(&o));
                        uint32_t mu = *reinterpret_cast<uint32_t*>(&m);
                        o = __shfl_xor_sync(0xffffffffu, mu, 8);
                        m = __hmax2(m, *reinterpret_cast<__nv_bfloat162*>(&o));
                        r[i] = *reinterpret_cast<uint32_t*>(&m);
                    }
                    if ((lane & 9) == 0 && in_range) {
                        const size_t ppix = (static_cast<size_t>(b) * (p.H >> 1) + (h >> 1)) * (p.W >> 1) + (w >> 1);
                        __nv_bfloat16* dst = p.pool_out + ppix * p.c_out + n0 + c;
                        st_global_256(dst, pk[0], pk[1]);
                        st_global_256(dst + 16, pk[2], pk[3]);
                    }
                }
            }
            buf ^= 1;
        }
        {
            if (p.stat_mode == 1 && p.bn == 128) {
#pragma unroll
                for (int chunk = 0; chunk < 2; ++chunk) {
                    atomicAdd(p.stat_sums + n0 + c_lo + 32 * chunk + lane, wide0[chunk]);
                    atomicAdd(p.stat_sums + p.c_out + n0 + c_lo + 32 * chunk + lane, wide1[chunk]);
                }
            } else if (p.stat_mode != 0) {
                // column sums over the 32 pixel rows of this warp (butterfly), then one global atomicAdd per channel and warp
#pragma unroll
                for (int j = 0; j < kHaloStatBn / 2; ++j) {
                    float a0 = acc0[j], a1 = acc1[j];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
                    }
                    if (lane == j) {
                        const int ch_g = n0 + c_lo + j;
                        if (p.stat_mode == 2) a1 = p.bn_rstd[ch_g] * (a1 - p.bn_mean[ch_g] * a0);
                        atomicAdd(p.stat_sums + ch_g, a0);
                        atomicAdd(p.stat_sums + p.c_out + ch_g, a1);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (kPair) cluster_sync_all();   // the leader's MMAs read the peer's shared memory and write its TMEM: leave together
    if (warp == 1) { if (kPair) tmem_dealloc_pair(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols); }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// dW[co, tap, ci] = sum over pixels p of dZ[p, co] * X[p + shift(tap), ci]        (autograd of the conv above)
// GEMM per CTA:  D[128 co][up to 256 (tap,ci) columns] += A * B^T with K = pixels.  Both operands are consumed
// exactly as they lie in the NHWC tensors - pixel rows of 64 contiguous channels - i.e. "MN-major" for UMMA: one TMA
// box (64 ch x 128 px, SWIZZLE_128B) is a [K=128][MN=64] slab; M = 128 uses two slabs of dZ (the second is zero-filled
// by TMA when c_out == 64), N packs up to four slabs of X, one per (tap, 64-channel block) pair, each loaded at its own
// tap-shifted origin (zero fill = padding).  Split-K over pixel tiles; partial tiles are reduced with fp32 atomics.
struct WgradParams {
    int taps;
    int c_in, c_out;
    int B, H, W;
    int bw, bh, bb;
    int tiles_w, tiles_h, tiles_b;
    int n_slabs;        // taps * c_in / 64  (tap, channel-block) column slabs of dW
    int slab_groups;    // ceil(n_slabs / 4)
    int splits;         // split-K factor
    int stages;
    float* dw;          // fp32 [c_out, taps, c_in], accumulated into
};

constexpr int kSlabBytes = kTileM * kKStep * 2;  // 16 KB: 128 pixel rows x 64 channels

__device__ __forceinline__ uint64_t make_sw128_mn_desc(const void* smem_ptr, uint32_t lbo_bytes) {
    const uint32_t addr = smem_u32(smem_ptr);
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
    desc |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;  // between 64-element MN slabs
    desc |= static_cast<uint64_t>(1024 >> 4) << 32;       // between 8-row K groups
    desc |= static_cast<uint64_t>(1) << 46;
    desc |= static_cast<uint64_t>(2) << 61;               // SWIZZLE_128B
    return desc;
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_x,
                  const WgradParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // stage = [2 dZ slabs | 4 X slabs] = 96 KB
    constexpr int kStageBytes = 6 * kSlabBytes;
    unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + static_cast<size_t>(p.stages) * kStageBytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* accum_bar = empty_bar + p.stages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = uniform_warp_index();
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        prefetch_tmap(&map_dz);
        prefetch_tmap(&map_x);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // work decomposition: blockIdx.x -> (split, slab group), blockIdx.y -> 128-row block of c_out
    const int group = blockIdx.x % p.slab_groups;
    const int split = blockIdx.x / p.slab_groups;
    const int m0 = blockIdx.y * kTileM;
    const int slab0 = group * 4;
    const int n_slab = min(4, p.n_slabs - slab0);          // slabs (64 columns each) in this group
    const int cblocks = p.c_in / kKStep;
    const int n_pix_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    const int k_begin = static_cast<int>(static_cast<long long>(n_pix_tiles) * split / p.splits);
    const int k_end = static_cast<int>(static_cast<long long>(n_pix_tiles) * (split + 1) / p.splits);
    const int bn = n_slab * 64;

    if (warp == 0) {
        {   // all lanes walk the loop (uniform control flow: TMA operands stay in uniform registers), one lane issues
            const bool leader = elect_one_sync();
            int stage = 0;
            unsigned phase = 1;
            for (int kt = k_begin; kt < k_end; ++kt) {
                int t = kt;
                const int tw = t % p.tiles_w; t /= p.tiles_w;
                const int th = t % p.tiles_h; t /= p.tiles_h;
                const int w0 = tw * p.bw, h0 = th * p.bh, b0 = t * p.bb;
                mbar_wait(&empty_bar[stage], phase);
                unsigned char* dst = tiles + static_cast<size_t>(stage) * kStageBytes;
                if (leader) {
                    mbar_arrive_expect_tx(&full_bar[stage], static_cast<unsigned>((2 + n_slab) * kSlabBytes));
                    tma_load_4d(dst, &map_dz, &full_bar[stage], m0, w0, h0, b0);
                    tma_load_4d(dst + kSlabBytes, &map_dz, &full_bar[stage], m0 + 64, w0, h0, b0);  // zero fill if >= c_out
                }
                for (int sl = 0; sl < n_slab; ++sl) {
                    const int slab = slab0 + sl;
                    const int tap = slab / cblocks, cb = slab - tap * cblocks;
                    const int dy = (p.taps == 9) ? tap / 3 - 1 : 0;
                    const int dx = (p.taps == 9) ? tap % 3 - 1 : 0;
                    if (leader)
                        tma_load_4d(dst + (2 + sl) * kSlabBytes, &map_x, &full_bar[stage], cb * kKStep, w0 + dx, h0 + dy, b0);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (k_end > k_begin) {   // all lanes walk the loop, the elected lane issues (see elect_one_sync)
            const bool leader = elect_one_sync();
            // D = F32, A = B = BF16, both MN-major (bits 15, 16), M = 128, N = bn
            const uint32_t idesc = make_idesc_bf16(bn) | (1u << 15) | (1u << 16);
            int stage = 0;
            unsigned phase = 0;
            for (int kt = k_begin; kt < k_end; ++kt) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const unsigned char* src = tiles + static_cast<size_t>(stage) * kStageBytes;
#pragma unroll
                for (int k = 0; k < kTileM / kUmmaK; ++k) {  // 8 MMAs, 16 pixel rows (2 KB) each
                    const uint64_t desc_a = make_sw128_mn_desc(src + k * 2048, kSlabBytes);
                    const uint64_t desc_b = make_sw128_mn_desc(src + 2 * kSlabBytes + k * 2048, kSlabBytes);
                    if (leader) umma_bf16(tmem_base, desc_a, desc_b, idesc, (kt > k_begin || k > 0) ? 1u : 0u);
                }
                if (leader) umma_commit(&empty_bar[stage]);
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            if (leader) umma_commit(accum_bar);
            __syncwarp();
        }
    } else if (k_end > k_begin) {
        const int quad = warp & 3;
        const int co = m0 + quad * 32 + lane;  // row of dW
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const int K_total = p.taps * p.c_in;
        for (int c = 0; c < bn; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(c), v);
            tmem_ld_wait();
            if (co < p.c_out) {
                // column c of this group -> slab (tap, channel block) -> offset tap*c_in + cb*64 + (c % 64) in a dW row
                const int slab = slab0 + c / 64;
                float* dst = p.dw + static_cast<size_t>(co) * K_total + static_cast<size_t>(slab) * 64 + (c % 64);
#pragma unroll
                for (int j = 0; j < 32; ++j) atomicAdd(dst + j, __uint_as_float(v[j]));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------ halo weight gradient
// The kernel above re-fetches X once per tap (nine tap-shifted 16 KB slabs per 64-channel block) and is bound by L2->SM
// traffic on the wide layers (tensor pipe 41-50 %).  Here a pixel tile is 8 wide x 16 tall (K = 128 pixels, k = y*8 + x, so
// an 8-row K group is one output row) and ONE halo box of X (16 x 18 pixel rows of 64 channels, origin (w0-1, h0-1), TMA
// zero fill = padding) serves all nine taps: tap (dy, dx) is the same box read from byte offset (dy*16 + dx)*128 with
// SBO = 16 rows * 128 B between K groups (the 128B swizzle is applied on absolute shared-memory address bits, as measured
// for the forward halo kernel).  The GEMM is transposed with respect to conv_wgrad_kernel:
//     D[(tap, ci) : 128 rows = TWO taps x 64 channels][c_out block of 64] += X_view^T * dZ      (both operands MN-major)
// A = two tap views of the halo box, LBO = byte distance between the two views (128 B or 1792 B); B = one dZ slab.
// Five tap pairs -> five accumulators of 64 TMEM columns; every MMA row is useful (the ninth tap wastes half of one MMA),
// where the M = c_out form wastes half of every MMA on the 64-channel layers.  Per stage 52 KB are loaded for 40 MMAs
// (1280 tensor cycles) instead of 96 KB for 8 (512 cycles): 4.6x less L2->SM traffic per FLOP.
struct WgradHaloParams {
    int c_in, c_out;
    int B, H, W;
    int tiles_w, tiles_h;   // 8-wide, 16-tall pixel tiles (partial tiles are zero-filled by TMA: they contribute nothing)
    int cblocks;            // c_in / 64
    int splits;
    int stages;
    float* dw;              // fp32 [c_out, 9, c_in], accumulated into
    // wide != 0 (c_out a multiple of 128): N = 128 - two dZ slabs per stage - so that an MMA's operand reads are 4 KB of A
    // + 4 KB of B per 64 tensor cycles (128 B/cycle) instead of 4 + 2 KB per 32 (192 B/cycle, which caps N = 64 at 67 % of the
    // tensor pipe).  Five accumulators of 128 columns do not fit the 512 TMEM columns, so the nine taps are dealt to two
    // kinds of CTA: tap pairs 0-1 (taps 0..3) and tap pairs 2-4 (taps 4..8); of the `splits` CTAs of a work item the first
    // `splits0` are of the first kind (2 : 3, the ratio of their MMA counts).
    int wide;
    int splits0;
};

constexpr int kWgHaloStageBytes = kHaloBytes + kSlabBytes;  // 36 KB X halo + 16 KB dZ slab (wide: + a second slab)

__device__ __forceinline__ uint64_t make_sw128_mn_desc_halo(const void* smem_ptr, uint32_t lbo_bytes) {
    const uint32_t addr = smem_u32(smem_ptr);
    uint64_t desc = 0;
    desc |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
    desc |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;          // between the two 64-row MN slabs (= the two taps)
    desc |= static_cast<uint64_t>((kHaloW * 128) >> 4) << 32;     // between 8-row K groups: next output row of the halo
    desc |= static_cast<uint64_t>(1) << 46;
    desc |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B
    return desc;
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_dz, const __grid_constant__ CUtensorMap map_x,
                       const WgradHaloParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* tiles = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int stage_bytes = kWgHaloStageBytes + (p.wide ? kSlabBytes : 0);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tiles + static_cast<size_t>(p.stages) * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* accum_bar = empty_bar + p.stages;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(accum_bar + 1);

    const int warp = uniform_warp_index();
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        prefetch_tmap(&map_dz);
        prefetch_tmap(&map_x);
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);   // five accumulators x 64 columns (power of two >= 320)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    // work item: blockIdx.x -> (split, channel block of X), blockIdx.y -> 64-channel (wide: 128-channel) block of c_out
    const int cb = blockIdx.x % p.cblocks;
    int split = blockIdx.x / p.cblocks;
    int n_splits = p.splits;
    const int bn = p.wide ? 128 : 64;
    const int n0 = blockIdx.y * bn;
    int pr_begin = 0, pr_end = 5;          // tap pairs of this CTA
    if (p.wide) {
        if (split < p.splits0) { pr_end = 2; n_splits = p.splits0; }
        else { pr_begin = 2; split -= p.splits0; n_splits = p.splits - p.splits0; }
    }
    const int n_pix_tiles = p.tiles_w * p.tiles_h * p.B;
    const int k_begin = static_cast<int>(static_cast<long long>(n_pix_tiles) * split / n_splits);
    const int k_end = static_cast<int>(static_cast<long long>(n_pix_tiles) * (split + 1) / n_splits);

    if (warp == 0) {
        {   // all lanes walk the loop (uniform control flow: TMA operands stay in uniform registers), one lane issues
            const bool leader = elect_one_sync();
            int stage = 0;
            unsigned phase = 1;
            for (int kt = k_begin; kt < k_end; ++kt) {
                int t = kt;
                const int tw = t % p.tiles_w; t /= p.tiles_w;
                const int th = t % p.tiles_h; t /= p.tiles_h;
                const int w0 = tw * kHaloTileW, h0 = th * kHaloTileH, b0 = t;
                mbar_wait(&empty_bar[stage], phase);
                unsigned char* dst = tiles + static_cast<size_t>(stage) * stage_bytes;
                if (leader) {
                    mbar_arrive_expect_tx(&full_bar[stage], static_cast<unsigned>(stage_bytes));
                    tma_load_4d(dst, &map_x, &full_bar[stage], cb * kKStep, w0 - 1, h0 - 1, b0);
                    tma_load_4d(dst + kHaloBytes, &map_dz, &full_bar[stage], n0, w0, h0, b0);
                    if (p.wide) tma_load_4d(dst + kHaloBytes + kSlabBytes, &map_dz, &full_bar[stage], n0 + 64, w0, h0, b0);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 1) {
        if (k_end > k_begin) {   // all lanes walk the loop, the elected lane issues (see elect_one_sync)
            const bool leader = elect_one_sync();
            // D = F32, A = B = BF16, both MN-major (bits 15, 16), M = 128, N = 64 (wide: 128 = two slabs, LBO apart)
            const uint32_t idesc = make_idesc_bf16(bn) | (1u << 15) | (1u << 16);
            int stage = 0;
            unsigned phase = 0;
            for (int kt = k_begin; kt < k_end; ++kt) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const unsigned char* halo = tiles + static_cast<size_t>(stage) * stage_bytes;
                const unsigned char* dz = halo + kHaloBytes;
#pragma unroll
                for (int k = 0; k < kTileM / kUmmaK; ++k) {   // 16 pixels = output rows 2k, 2k+1 of the tile
                    const uint64_t desc_b = make_sw128_mn_desc(dz + k * 2048, kSlabBytes);
#pragma unroll
                    for (int pr = 0; pr < 5; ++pr) {
                        if (pr < pr_begin || pr >= pr_end) continue;                  // (warp-uniform) the other kind's pairs
                        const int t0 = 2 * pr, t1 = (pr < 4) ? 2 * pr + 1 : 2 * pr;   // ninth tap: second half unused
                        const int off0 = ((t0 / 3) * kHaloW + (t0 % 3)) * 128;
                        const int off1 = ((t1 / 3) * kHaloW + (t1 % 3)) * 128;
                        const uint32_t lbo = (pr < 4) ? static_cast<uint32_t>(off1 - off0) : 128u;
                        const uint64_t desc_a = make_sw128_mn_desc_halo(halo + off0 + (2 * k) * kHaloW * 128, lbo);
                        if (leader)
                            umma_bf16(tmem_base + static_cast<uint32_t>((pr - pr_begin) * bn), desc_a, desc_b, idesc,
                                      (kt > k_begin || k > 0) ? 1u : 0u);
                    }
                }
                if (leader) umma_commit(&empty_bar[stage]);
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            if (leader) umma_commit(accum_bar);
            __syncwarp();
        }
    } else if (k_end > k_begin) {
        const int quad = warp & 3;
        const int row = quad * 32 + lane;      // accumulator row = (tap within the pair, input channel)
        const int ci = cb * kKStep + (row & 63);
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        const size_t K_total = static_cast<size_t>(9) * p.c_in;
        const int chunks_per_acc = bn / 32;
        const int n_chunks = (pr_end - pr_begin) * chunks_per_acc;
#pragma unroll 1
        for (int it = 0; it < n_chunks; ++it) {
            // every split adds into the same block of dW: start each CTA at a different 32-column chunk so that
            // concurrently finishing CTAs do not all hit the same L2 lines at once
            const int o = (it + split) % n_chunks;
            const int acc = o / chunks_per_acc, c = (o % chunks_per_acc) * 32;
            const int pr = pr_begin + acc;
            const int tap = 2 * pr + (row >> 6);
            {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * bn + c), v);
                tmem_ld_wait();
                if (tap < 9) {
                    // column j = output channel n0+c+j; the 32 lanes of a warp are 32 consecutive input channels of one tap:
                    // each atomicAdd instruction covers 128 contiguous bytes of a dW row
                    float* dst = p.dw + static_cast<size_t>(n0 + c) * K_total + static_cast<size_t>(tap) * p.c_in + ci;
#pragma unroll
                    for (int j = 0; j < 32; ++j) atomicAdd(dst + static_cast<size_t>(j) * K_total, __uint_as_float(v[j]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// NHWC activation [B,H,W,C] with a (128 bytes of channels, bw, bh, bb) box: 64 bf16 or (f32 = true) 32 fp32 channels
int make_act_map(CUtensorMap* map, const void* base, int B, int H, int W, int C, int bw, int bh, int bb, bool f32 = false) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(IM2IM_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    const cuuint64_t es = f32 ? 4 : 2;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * es, (cuuint64_t)W * C * es, (cuuint64_t)H * W * C * es};
    cuuint32_t box[4] = {(cuuint32_t)(f32 ? kKStep / 2 : kKStep), (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(IM2IM_ECUDA, "cuTensorMapEncodeTiled(activation) failed: %d", (int)r);
    return IM2IM_OK;
}

// weights [c_out, K] (K = taps*c_in contiguous) with a (128 bytes, bn) box: bf16, or fp32 when f32 = true
int make_weight_map(CUtensorMap* map, const void* base, int c_out, int K, int bn, bool f32 = false) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) return fail(IM2IM_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)c_out};
    cuuint64_t strides[1] = {(cuuint64_t)K * (f32 ? 4 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)(f32 ? kKStep / 2 : kKStep), (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(IM2IM_ECUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
    return IM2IM_OK;
}

// pick a 128-pixel box (bw, bh, bb) that tiles (W, H, B) with the least padding
void pick_box(int B, int H, int W, int& bw, int& bh, int& bb) {
    long long best = -1;
    for (int cw = 1; cw <= 128; cw <<= 1) {
        for (int ch = 1; cw * ch <= 128; ch <<= 1) {
            const int cb = 128 / (cw * ch);
            if (cw > 256 || ch > 256 || cb > 256) continue;
            const long long tiles = (long long)((W + cw - 1) / cw) * ((H + ch - 1) / ch) * ((B + cb - 1) / cb);
            // prefer fewer tiles (less padding); tie-break towards wide boxes (longer contiguous rows)
            const long long score = tiles * 1024 - cw;
            if (best < 0 || score < best) { best = score; bw = cw; bh = ch; bb = cb; }
        }
    }
}

}  // namespace
}  // namespace im2im

using namespace im2im;

namespace im2im {
struct FusedStats {          // optional per-channel statistics fused into the halo kernel's epilogue (HaloParams::stat_mode)
    int mode = 0;
    float* sums = nullptr;
    const void* bn_z = nullptr;
    const float* gamma = nullptr; const float* beta = nullptr; const float* mean = nullptr; const float* rstd = nullptr;
};
int conv_igemm_impl(const void* d_x1, int32_t c_in1, const void* d_x2, int32_t c_in2, const void* d_weight,
                    const float* d_bias, int32_t B, int32_t H, int32_t W, int32_t c_out, int32_t taps, int32_t relu,
                    void* d_out_bf16, float* d_out_f32, const FusedStats& fs, int* fused, void* stream,
                    void* d_pool_out = nullptr, int* pooled = nullptr);
int halo_dispatch(const void* d_x1, int32_t c_in1, const void* d_x2, int32_t c_in2, const void* d_weight, const float* d_bias,
                  int32_t B, int32_t H, int32_t W, int32_t c_out, int32_t taps, int32_t relu, void* d_out_bf16, float* d_out_f32,
                  const FusedStats& fs, int* fused, void* stream, void* d_pool_out, int* pooled, bool tf32, int* launched);
}

extern "C" int im2im_conv_igemm_bf16(const void* d_x1, int32_t c_in1, const void* d_x2, int32_t c_in2,
                                     const void* d_weight, const float* d_bias, int32_t B, int32_t H, int32_t W,
                                     int32_t c_out, int32_t taps, int32_t relu, void* d_out_bf16, float* d_out_f32,
                                     void* stream) {
    return conv_igemm_impl(d_x1, c_in1, d_x2, c_in2, d_weight, d_bias, B, H, W, c_out, taps, relu, d_out_bf16, d_out_f32,
                           FusedStats{}, nullptr, stream);
}

extern "C" int im2im_conv_igemm_bf16_pool(const void* d_x1, int32_t c_in1, const void* d_x2, int32_t c_in2,
                                          const void* d_weight, const float* d_bias, int32_t B, int32_t H, int32_t W,
                                          int32_t c_out, int32_t taps, int32_t relu, void* d_out_bf16, void* d_pool_out,
                                          int32_t* h_pooled, void* stream) {
    if (!d_out_bf16 || !d_pool_out || !h_pooled) return fail(IM2IM_EINVAL, "conv_igemm_pool: null pointer");
    if (reinterpret_cast<uintptr_t>(d_pool_out) & 31u) return fail(IM2IM_EINVAL, "conv_igemm_pool: the pooled tensor must be 32-byte aligned");
    int pooled = 0;
    const int rc = conv_igemm_impl(d_x1, c_in1, d_x2, c_in2, d_weight, d_bias, B, H, W, c_out, taps, relu, d_out_bf16, nullptr,
                                   FusedStats{}, nullptr, stream, d_pool_out, &pooled);
    *h_pooled = pooled;
    return rc;
}

extern "C" int im2im_conv_igemm_bf16_stats(const void* d_x1, int32_t c_in1, const void* d_x2, int32_t c_in2,
                                           const void* d_weight, int32_t B, int32_t H, int32_t W, int32_t c_out,
                                           int32_t taps, void* d_out_bf16, int32_t stat_mode, float* d_stat_sums,
                                           const void* d_bn_z, const float* d_bn_gamma, const float* d_bn_beta,
                                           const float* d_bn_mean, const float* d_bn_rstd, int32_t* h_fused,
                                           void* stream) {
    if (stat_mode != 1 && stat_mode != 2) return fail(IM2IM_EINVAL, "conv_igemm_stats: stat_mode must be 1 or 2");
    if (!d_stat_sums || !d_out_bf16 || !h_fused) return fail(IM2IM_EINVAL, "conv_igemm_stats: null pointer");
    if (stat_mode == 2 && (!d_bn_z || !d_bn_gamma || !d_bn_beta || !d_bn_mean || !d_bn_rstd))
        return fail(IM2IM_EINVAL, "conv_igemm_stats: stat_mode 2 needs bn_z / gamma / beta / mean / rstd");
    FusedStats fs;
    fs.mode = stat_mode; fs.sums = d_stat_sums; fs.bn_z = d_bn_z; fs.gamma = d_bn_gamma; fs.beta = d_bn_beta;
    fs.mean = d_bn_mean; fs.rstd = d_bn_rstd;
    int fused = 0;
    const int rc = conv_igemm_impl(d_x1, c_in1, d_x2, c_in2, d_weight, nullptr, B, H, W, c_out, taps, 0, d_out_bf16, nullptr,
                                   fs, &fused, stream);
    *h_fused = fused;
    return rc;
}

// Wide shallow layers: halo kernel (one activation box per channel block, taps via shifted descriptors, weights resident or
// streamed).  Shared by the bf16 and the tf32 (fp32 activations / weights, kind::tf32; everything is counted in 128-byte K
// blocks = 64 bf16 or 32 fp32 channels) entry points.  *launched = 1 when the layer was taken.
int im2im::halo_dispatch(const void* d_x1, int32_t c_in1, const void* d_x2, int32_t c_in2, const void* d_weight,
                         const float* d_bias, int32_t B, int32_t H, int32_t W, int32_t c_out, int32_t taps, int32_t relu,
                         void* d_out_bf16, float* d_out_f32, const FusedStats& fs, int* fused, void* stream,
                         void* d_pool_out, int* pooled, bool tf32, int* launched) {
    int rc = 0;
    *launched = 0;
    const int esz = tf32 ? 4 : 2;                 // bytes per activation / weight element
    const int kch = 128 / esz;                    // channels per 128-byte K block
    static const bool no_halo = (getenv("IM2IM_CONV_NO_HALO") != nullptr);
    if (!no_halo && taps == 9 && W % kHaloTileW == 0 && H % kHaloTileH == 0 && c_in1 % kch == 0 && c_in2 % kch == 0) {
        const int c_in = c_in1 + c_in2;
        const long long k_bytes = static_cast<long long>(c_in) * esz;   // bytes of one tap of one weight row
        const bool staged = d_out_bf16 != nullptr && d_out_f32 == nullptr;
        const bool want_stats = staged && fs.mode != 0;
        // CTA pairs (cta_group::2): each CTA keeps half of the weight rows resident.  IM2IM_HALO_PAIR=0 restores single CTAs
        // (read per call so that a test can compare the two in one process).  A pair also takes N blocks whose weights only
        // fit when halved - 128 -> 128 and 128 -> 256 run as N = 128 tiles (1.6 PFLOP/s at batch 78) instead of N = 64 -
        // unless IM2IM_HALO_PAIR_WIDE=0
        const char* pair_e = getenv("IM2IM_HALO_PAIR");
        const bool pair = !(pair_e != nullptr && pair_e[0] == '0') && sm_count() >= 2;
        const char* wide_e = getenv("IM2IM_HALO_PAIR_WIDE");
        const long long w_budget = (pair && !(wide_e != nullptr && wide_e[0] == '0')) ? 2 * 147456 : 147456;
        // forward statistics on N = 128 tiles (per-tile shuffle reduction) where the K loop is long enough to carry it
        const char* s128_e = getenv("IM2IM_HALO_STATS128");
        const bool stats128 = want_stats && fs.mode == 1 && k_bytes >= 256 && !(s128_e != nullptr && s128_e[0] == '0');
        int hbn = 0;
        for (int cand : {128, 64})
            // (the doubled budget is for N = 128 only: a layer whose N = 64 block needs it - 256 input channels - is better
            // off on the persistent kernel's N = 256 tiles; measured 256 -> 256 @80^2: 0.42 ms there, 0.47 ms here)
            // (c_out == 64 has no wider alternative, so its N = 64 block may use the doubled budget too)
            if (hbn == 0 && c_out % cand == 0 && 9ll * k_bytes * cand <= ((cand == 128 || c_out == 64) ? w_budget : 147456) &&
                !(want_stats && cand != kHaloStatBn && !(stats128 && cand == 128))) hbn = cand;
        // Weights that do not fit at all (256 input channels and more): a pair can still run N = 128 tiles with the weights
        // STREAMED - each ring stage = halo box + the nine tap slabs of its channel block, 108 KB, two stages.  Per stage
        // 36 MMAs of 64 cycles: 48 B/cycle of L2->SM traffic per SM, half of what the persistent kernel's 128 x 256 tiles
        // ask for.  IM2IM_HALO_STREAM_MAX_CIN (default 256, 0 = off) bounds the layers that take this route.
        int w_stream = 0;
        if (hbn == 0 && pair && (!want_stats || stats128) && c_out % 128 == 0) {
            const char* se = getenv("IM2IM_HALO_STREAM_MAX_CIN");
            const int max_cin = se != nullptr ? atoi(se) : 256;
            if (k_bytes <= 2ll * max_cin) { hbn = 128; w_stream = 1; }   // (the bound is stated in bf16 channels)
        }
        if (hbn != 0) {
            HaloParams h{};
            h.c_in1 = c_in1; h.c_in2 = c_in2; h.c_out = c_out; h.B = B; h.H = H; h.W = W;
            h.tiles_w = W / kHaloTileW; h.tiles_h = H / kHaloTileH; h.bn = hbn; h.relu = relu; h.bias = d_bias;
            static const bool skip_store = (getenv("IM2IM_HALO_SKIP_STORE") != nullptr);
            h.debug_skip_store = skip_store ? 1 : 0;
            h.out_bf16 = static_cast<__nv_bfloat16*>(d_out_bf16); h.out_f32 = d_out_f32;
            h.out_planar = nullptr; h.n_real = 0; h.act_kind = 0; h.act_from = 0;
            if (want_stats) {
                h.stat_mode = fs.mode; h.stat_sums = fs.sums; h.bn_z = static_cast<const __nv_bfloat16*>(fs.bn_z);
                h.bn_gamma = fs.gamma; h.bn_beta = fs.beta; h.bn_mean = fs.mean; h.bn_rstd = fs.rstd;
            }
            h.w_stream = w_stream; h.tf32 = tf32 ? 1 : 0;
            const int w_bytes = w_stream ? 0 : static_cast<int>(9 * k_bytes * hbn / (pair ? 2 : 1));
            const int stage_bytes = kHaloBytes + (w_stream ? 9 * (hbn / 2) * kKStep * 2 : 0);
            const int tail_bytes = 128 + 2 * hbn * 4;      // barriers + scale/shift behind the A ring
            h.a_stages = (232448 - 1024 - w_bytes - tail_bytes) / stage_bytes;
            if (h.a_stages > 4) h.a_stages = 4;
            if (h.a_stages >= 2) {
                CUtensorMap h1, h2, hw;
                rc = make_act_map(&h1, d_x1, B, H, W, c_in1, kHaloW, kHaloH, 1, tf32);
                if (rc) return rc;
                if (c_in2 > 0) { rc = make_act_map(&h2, d_x2, B, H, W, c_in2, kHaloW, kHaloH, 1, tf32); if (rc) return rc; }
                else h2 = h1;
                rc = make_weight_map(&hw, d_weight, c_out, taps * c_in, pair ? hbn / 2 : hbn, tf32);
                if (rc) return rc;
                const size_t hsmem = static_cast<size_t>(w_bytes) + static_cast<size_t>(h.a_stages) * stage_bytes + tail_bytes + 1024;
                void (*kern)(CUtensorMap, CUtensorMap, CUtensorMap, HaloParams) =
                    tf32 ? (pair ? conv_halo_kernel<true, true> : conv_halo_kernel<false, true>)
                         : (pair ? conv_halo_kernel<true, false> : conv_halo_kernel<false, false>);
                IM2IM_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
                if (fused) *fused = h.stat_mode != 0 ? 1 : 0;
                if (d_pool_out != nullptr && h.stat_mode == 0 && d_out_bf16 != nullptr && d_out_f32 == nullptr && H % 2 == 0 && W % 2 == 0) {
                    h.pool_out = static_cast<__nv_bfloat16*>(d_pool_out);
                    if (pooled) *pooled = 1;
                }
                const int n_blocks_n = c_out / hbn;
                const long long m_tiles = static_cast<long long>(h.tiles_w) * h.tiles_h * B;
                long long gx = sm_count() / n_blocks_n;
                if (gx < 1) gx = 1;
                if (gx > m_tiles) gx = m_tiles;
                if (pair) {
                    // clusters of two CTAs along x; a pair takes two neighbouring tiles per step
                    const long long pairs = (m_tiles + 1) / 2;
                    long long gp = sm_count() / 2 / n_blocks_n;
                    if (gp < 1) gp = 1;
                    if (gp > pairs) gp = pairs;
                    cudaLaunchConfig_t cfg{};
                    cfg.gridDim = dim3(static_cast<unsigned>(2 * gp), static_cast<unsigned>(n_blocks_n));
                    cfg.blockDim = dim3(kHaloThreads);
                    cfg.dynamicSmemBytes = hsmem;
                    cfg.stream = static_cast<cudaStream_t>(stream);
                    cudaLaunchAttribute attr[1];
                    attr[0].id = cudaLaunchAttributeClusterDimension;
                    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                    cfg.attrs = attr; cfg.numAttrs = 1;
                    IM2IM_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, h1, h2, hw, h));
                    *launched = 1;
                    return check_launch("conv_halo_kernel<pair>");
                }
                dim3 hgrid(static_cast<unsigned>(gx), static_cast<unsigned>(n_blocks_n));
                kern<<<hgrid, kHaloThreads, hsmem, static_cast<cudaStream_t>(stream)>>>(h1, h2, hw, h);
                *launched = 1;
                return check_launch("conv_halo_kernel");
            }
        }
    }
    (void)rc;
    return IM2IM_OK;
}

int im2im::conv_igemm_impl(const void* d_x1, int32_t c_in1, const void* d_x2, int32_t c_in2, const void* d_weight,
                           const float* d_bias, int32_t B, int32_t H, int32_t W, int32_t c_out, int32_t taps, int32_t relu,
                           void* d_out_bf16, float* d_out_f32, const FusedStats& fs, int* fused, void* stream,
                           void* d_pool_out, int* pooled) {
    if (taps != 9 && taps != 1) return fail(IM2IM_EINVAL, "taps must be 9 (3x3) or 1 (1x1), got %d", taps);
    if (B <= 0 || H <= 0 || W <= 0) return fail(IM2IM_EINVAL, "bad activation shape %dx%dx%d", B, H, W);
    if (c_in1 <= 0 || c_in1 % kKStep || c_in2 < 0 || c_in2 % kKStep)
        return fail(IM2IM_ERANGE, "input channels must be multiples of %d (got %d + %d)", kKStep, c_in1, c_in2);
    if (c_out < 32 || c_out % 32) return fail(IM2IM_ERANGE, "c_out must be a multiple of 32 (got %d)", c_out);
    if (!d_x1 || !d_weight || (!d_out_bf16 && !d_out_f32) || (c_in2 > 0 && !d_x2))
        return fail(IM2IM_EINVAL, "null tensor");
    if ((reinterpret_cast<uintptr_t>(d_out_bf16) | reinterpret_cast<uintptr_t>(d_out_f32)) & 31u)
        return fail(IM2IM_EINVAL, "conv_igemm: the output tensor must be 32-byte aligned (the epilogue stores 32 bytes at a time)");
    ConvParams p;
    p.taps = taps; p.c_in1 = c_in1; p.c_in2 = c_in2; p.c_out = c_out; p.B = B; p.H = H; p.W = W;
    pick_box(B, H, W, p.bw, p.bh, p.bb);
    p.tiles_w = (W + p.bw - 1) / p.bw; p.tiles_h = (H + p.bh - 1) / p.bh; p.tiles_b = (B + p.bb - 1) / p.bb;
    p.bn = (c_out % 256 == 0) ? 256 : (c_out % 128 == 0) ? 128 : (c_out % 64 == 0) ? 64 : 32;
    const int stage_bytes = kATileBytes + p.bn * kKStep * 2;
    p.stages = (200 * 1024) / stage_bytes;
    if (p.stages > 8) p.stages = 8;
    p.relu = relu; p.bias = d_bias; p.tf32 = 0; p.kstep = kKStep;
    p.out_bf16 = static_cast<__nv_bfloat16*>(d_out_bf16); p.out_f32 = d_out_f32; p.stat_sums = nullptr;
    CUtensorMap m1, m2, mw;
    int rc = make_act_map(&m1, d_x1, B, H, W, c_in1, p.bw, p.bh, p.bb);
    if (rc) return rc;
    if (c_in2 > 0) { rc = make_act_map(&m2, d_x2, B, H, W, c_in2, p.bw, p.bh, p.bb); if (rc) return rc; }
    else m2 = m1;
    rc = make_weight_map(&mw, d_weight, c_out, taps * (c_in1 + c_in2), p.bn);
    if (rc) return rc;
    {
        int launched = 0;
        rc = halo_dispatch(d_x1, c_in1, d_x2, c_in2, d_weight, d_bias, B, H, W, c_out, taps, relu, d_out_bf16, d_out_f32, fs, fused,
                           stream, d_pool_out, pooled, false, &launched);
        if (rc || launched) return rc;
    }
    const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + (2 * p.stages + 4) * sizeof(uint64_t) + 16 + 1024;
    static const bool use_v1 = (getenv("IM2IM_CONV_V1") != nullptr);  // bring-up switch: one tile per CTA
    if (use_v1) {
        IM2IM_CUDA_TRY(cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(static_cast<unsigned>(p.tiles_w * p.tiles_h * p.tiles_b), static_cast<unsigned>(c_out / p.bn));
        conv_igemm_kernel<<<grid, kConvThreads, smem, static_cast<cudaStream_t>(stream)>>>(m1, m2, mw, p);
        return check_launch("conv_igemm_kernel");
    }
    IM2IM_CUDA_TRY(cudaFuncSetAttribute(conv_igemm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    // forward statistics in the persistent kernel's epilogue when the K loop is long enough to hide the extra shuffles
    // (>= 256 input channels: >= 9 k cycles of MMAs per tile); shallower layers keep the separate reduction pass
    static const bool no_pstats = (getenv("IM2IM_NO_PERSISTENT_STATS") != nullptr);
    if (fs.mode == 1 && !no_pstats && c_in1 + c_in2 >= 256 && d_out_bf16 != nullptr && d_out_f32 == nullptr) {
        p.stat_sums = fs.sums;
        if (fused) *fused = 1;
    }
    const long long n_tiles = static_cast<long long>(p.tiles_w) * p.tiles_h * p.tiles_b * (c_out / p.bn);
    const int sms = sm_count();
    const unsigned grid = static_cast<unsigned>(n_tiles < sms ? n_tiles : sms);
    conv_igemm_persistent_kernel<<<grid, kConvThreads, smem, static_cast<cudaStream_t>(stream)>>>(m1, m2, mw, p);
    return check_launch("conv_igemm_persistent_kernel");
}

// Reference-precision mode (SURVEY.md §7 hard part 3): fp32 NHWC activations and fp32 weights [c_out, taps, c_in] staged by
// TMA as they are, tcgen05 kind::tf32 (what torch's cuDNN convolutions do by default with the reference's fp32 modules),
// fp32 accumulation, fp32 NHWC output rounded to TF32.  Same persistent kernel and tile math as the bf16 path; a K step is
// 32 channels (128 bytes).
extern "C" int im2im_conv_igemm_tf32(const float* d_x1, int32_t c_in1, const float* d_x2, int32_t c_in2,
                                     const float* d_weight, const float* d_bias, int32_t B, int32_t H, int32_t W,
                                     int32_t c_out, int32_t taps, int32_t relu, float* d_out, void* stream) {
    if (taps != 9 && taps != 1) return fail(IM2IM_EINVAL, "taps must be 9 (3x3) or 1 (1x1), got %d", taps);
    if (B <= 0 || H <= 0 || W <= 0) return fail(IM2IM_EINVAL, "bad activation shape %dx%dx%d", B, H, W);
    constexpr int ks = kKStep / 2;
    if (c_in1 <= 0 || c_in1 % ks || c_in2 < 0 || c_in2 % ks)
        return fail(IM2IM_ERANGE, "tf32: input channels must be multiples of %d (got %d + %d)", ks, c_in1, c_in2);
    if (c_out < 32 || c_out % 32) return fail(IM2IM_ERANGE, "c_out must be a multiple of 32 (got %d)", c_out);
    if (!d_x1 || !d_weight || !d_out || (c_in2 > 0 && !d_x2)) return fail(IM2IM_EINVAL, "null tensor");
    if (reinterpret_cast<uintptr_t>(d_out) & 31u)
        return fail(IM2IM_EINVAL, "conv_igemm_tf32: the output tensor must be 32-byte aligned (the epilogue stores 32 bytes at a time)");
    {   // the wide layers on the halo kernel (CTA pairs), as in the bf16 path; IM2IM_TF32_HALO=0 keeps them on the persistent kernel
        const char* e = getenv("IM2IM_TF32_HALO");
        if (!(e != nullptr && e[0] == '0')) {
            int launched = 0;
            const int hrc = halo_dispatch(d_x1, c_in1, d_x2, c_in2, d_weight, d_bias, B, H, W, c_out, taps, relu, nullptr, d_out,
                                          FusedStats{}, nullptr, stream, nullptr, nullptr, true, &launched);
            if (hrc || launched) return hrc;
        }
    }
    ConvParams p;
    p.taps = taps; p.c_in1 = c_in1; p.c_in2 = c_in2; p.c_out = c_out; p.B = B; p.H = H; p.W = W;
    pick_box(B, H, W, p.bw, p.bh, p.bb);
    p.tiles_w = (W + p.bw - 1) / p.bw; p.tiles_h = (H + p.bh - 1) / p.bh; p.tiles_b = (B + p.bb - 1) / p.bb;
    p.bn = (c_out % 256 == 0) ? 256 : (c_out % 128 == 0) ? 128 : (c_out % 64 == 0) ? 64 : 32;
    const int stage_bytes = kATileBytes + p.bn * 128;
    p.stages = (200 * 1024) / stage_bytes;
    if (p.stages > 8) p.stages = 8;
    p.relu = relu; p.bias = d_bias; p.tf32 = 1; p.kstep = ks;
    p.out_bf16 = nullptr; p.out_f32 = d_out; p.stat_sums = nullptr;
    CUtensorMap m1, m2, mw;
    int rc = make_act_map(&m1, d_x1, B, H, W, c_in1, p.bw, p.bh, p.bb, true);
    if (rc) return rc;
    if (c_in2 > 0) { rc = make_act_map(&m2, d_x2, B, H, W, c_in2, p.bw, p.bh, p.bb, true); if (rc) return rc; }
    else m2 = m1;
    rc = make_weight_map(&mw, d_weight, c_out, taps * (c_in1 + c_in2), p.bn, true);
    if (rc) return rc;
    const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + (2 * p.stages + 4) * sizeof(uint64_t) + 16 + 1024;
    IM2IM_CUDA_TRY(cudaFuncSetAttribute(conv_igemm_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    const long long n_tiles = static_cast<long long>(p.tiles_w) * p.tiles_h * p.tiles_b * (c_out / p.bn);
    const int sms = sm_count();
    const unsigned grid = static_cast<unsigned>(n_tiles < sms ? n_tiles : sms);
    conv_igemm_persistent_kernel<<<grid, kConvThreads, smem, static_cast<cudaStream_t>(stream)>>>(m1, m2, mw, p);
    return check_launch("conv_igemm_persistent_kernel<tf32>");
}

// planar fp32 [B, n, H, W] -> NHWC bf16 [B, H, W, 64] with channels >= n zero: the head's output gradient as a tensor-core
// operand (weight / data gradient of the head through the same halo kernels as every other layer)
namespace im2im { namespace {
__global__ void __launch_bounds__(256) planar_to_nhwc64_kernel(const float* __restrict__ src, int n, long long hw,
                                                               long long total_pix, __nv_bfloat16* __restrict__ dst) {
    // thread = (pixel, 8-channel group): 8 groups per pixel
    const long long total = total_pix * 8;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(e & 7);
        const long long pix = e >> 3;
        const long long b = pix / hw, k = pix - b * hw;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = g * 8 + j;
            f[j] = c < n ? __ldg(src + (b * n + c) * hw + k) : 0.f;
        }
        uint4 pk;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + pix * 64 + g * 8) = pk;
    }
}
} }

namespace im2im { namespace {
// n <= 8 planes: only the first 16 bytes (channels 0..7) of every 128-byte pixel row are written; the caller keeps the
// other 56 channels zero across calls (a persistent buffer), which cuts the bytes written 8x
__global__ void __launch_bounds__(256) planar_to_nhwc64_first8_kernel(const float* __restrict__ src, int n, long long hw,
                                                                      long long total_pix, __nv_bfloat16* __restrict__ dst) {
    for (long long pix = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pix < total_pix;
         pix += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long b = pix / hw, k = pix - b * hw;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = j < n ? __ldg(src + (b * n + j) * hw + k) : 0.f;
        uint4 pk;
        __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
        pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
        *reinterpret_cast<uint4*>(dst + pix * 64) = pk;
    }
}
} }

extern "C" int im2im_planar_to_nhwc64_first8_bf16(const float* d_src, int32_t n_planes, int32_t B, int32_t H, int32_t W,
                                                  void* d_dst, void* stream) {
    if (B <= 0 || H <= 0 || W <= 0 || n_planes < 1 || n_planes > 8) return fail(IM2IM_EINVAL, "planar_to_nhwc64_first8: bad shape");
    if (!d_src || !d_dst) return fail(IM2IM_EINVAL, "planar_to_nhwc64_first8: null tensor");
    const long long hw = static_cast<long long>(H) * W, pix = hw * B;
    long long blocks = (pix + 255) / 256;
    const long long cap = 16ll * sm_count();
    if (blocks > cap) blocks = cap;
    planar_to_nhwc64_first8_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_src, n_planes, hw, pix, static_cast<__nv_bfloat16*>(d_dst));
    return check_launch("planar_to_nhwc64_first8_kernel");
}

extern "C" int im2im_planar_to_nhwc64_bf16(const float* d_src, int32_t n_planes, int32_t B, int32_t H, int32_t W,
                                           void* d_dst, void* stream) {
    if (B <= 0 || H <= 0 || W <= 0 || n_planes < 1 || n_planes > 64) return fail(IM2IM_EINVAL, "planar_to_nhwc64: bad shape");
    if (!d_src || !d_dst) return fail(IM2IM_EINVAL, "planar_to_nhwc64: null tensor");
    const long long hw = static_cast<long long>(H) * W, pix = hw * B;
    long long blocks = (pix * 8 + 255) / 256;
    const long long cap = 16ll * sm_count();
    if (blocks > cap) blocks = cap;
    planar_to_nhwc64_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_src, n_planes, hw, pix, static_cast<__nv_bfloat16*>(d_dst));
    return check_launch("planar_to_nhwc64_kernel");
}

namespace im2im {
namespace {
int head_tc_impl(const void* d_x, const void* d_weight, const float* d_bias, const float* d_tap_bias, int32_t B, int32_t H,
                 int32_t W, int32_t n_real, int32_t act_kind, int32_t act_from_plane, float* d_out, const float* d_labels,
                 const float* d_lambdas, int32_t n_lambdas, uint32_t* d_hist, void* stream) {
    if (B <= 0 || H <= 0 || W <= 0) return fail(IM2IM_EINVAL, "head_tc: bad activation shape");
    if (n_real < 1 || n_real > 32) return fail(IM2IM_ERANGE, "head_tc: n_real=%d outside [1, 32]", n_real);
    if (act_kind < 0 || act_kind > 2) return fail(IM2IM_EINVAL, "head_tc: act_kind=%d", act_kind);
    if (W % kHaloTileW || H % kHaloTileH)
        return fail(IM2IM_ENOTSUP, "head_tc: needs W %% 8 == 0 and H %% 16 == 0 (got %dx%d); use im2im_head_conv3x3_act_f32", H, W);
    if (!d_x || !d_weight || (!d_out && !d_hist)) return fail(IM2IM_EINVAL, "head_tc: null tensor");
    HaloParams h{};
    h.c_in1 = 64; h.c_in2 = 0; h.c_out = 64; h.B = B; h.H = H; h.W = W;
    // N = 32: only the first n_real <= 32 weight rows are real, so half of the 64 packed rows are never loaded or multiplied
    h.tiles_w = W / kHaloTileW; h.tiles_h = H / kHaloTileH; h.bn = 32; h.relu = 0; h.bias = d_bias;
    h.out_bf16 = nullptr; h.out_f32 = nullptr;
    h.out_planar = d_out; h.n_real = n_real; h.act_kind = act_kind; h.act_from = act_kind ? act_from_plane : n_real;
    h.tap_bias = d_tap_bias;
    if (d_tap_bias && !d_bias) return fail(IM2IM_EINVAL, "head_tc: tap_bias without bias");
    if (d_tap_bias && n_real > kMaxFoldedPlanes)
        return fail(IM2IM_ENOTSUP, "head_tc: a folded 1x1 convolution is supported for up to %d planes (got %d)", kMaxFoldedPlanes, n_real);
    const int w_bytes = 9 * 64 * h.bn * 2;
    h.a_stages = 4;
    size_t hist_bytes = 0;
    if (d_hist != nullptr) {
        if (n_real != 3) return fail(IM2IM_ENOTSUP, "head_tc_hist: the histogram epilogue ranks (lower, prediction, upper) planes of a one-channel quantile head (n_real=%d)", n_real);
        if (!d_labels || !d_lambdas || n_lambdas < 1) return fail(IM2IM_EINVAL, "head_tc_hist: labels / lambda grid missing");
        h.hist_global = d_hist; h.hist_labels = d_labels; h.hist_lambdas = d_lambdas; h.hist_L = n_lambdas;
        hist_bytes = static_cast<size_t>(n_lambdas + 1) * 12 + static_cast<size_t>(n_lambdas) * 4 + 16;
        while (h.a_stages > 2 && w_bytes + static_cast<size_t>(h.a_stages) * kHaloBytes + 128 + 2 * h.bn * 4 + 1024 + hist_bytes > 232448)
            --h.a_stages;
        if (w_bytes + static_cast<size_t>(h.a_stages) * kHaloBytes + 128 + 2 * h.bn * 4 + 1024 + hist_bytes > 232448)
            return fail(IM2IM_ERANGE, "head_tc_hist: %d lambdas do not fit in shared memory next to the convolution", n_lambdas);
    }
    CUtensorMap h1, hw;
    int rc = make_act_map(&h1, d_x, B, H, W, 64, kHaloW, kHaloH, 1);
    if (rc) return rc;
    rc = make_weight_map(&hw, d_weight, 64, 9 * 64, h.bn);
    if (rc) return rc;
    const size_t hsmem = static_cast<size_t>(w_bytes) + static_cast<size_t>(h.a_stages) * kHaloBytes + 128 + 2 * h.bn * 4 + 1024 + hist_bytes;
    IM2IM_CUDA_TRY(cudaFuncSetAttribute(conv_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hsmem));
    const long long m_tiles = static_cast<long long>(h.tiles_w) * h.tiles_h * B;
    long long gx = sm_count();
    if (gx > m_tiles) gx = m_tiles;
    conv_halo_kernel<false><<<dim3(static_cast<unsigned>(gx), 1), kHaloThreads, hsmem, static_cast<cudaStream_t>(stream)>>>(h1, h1, hw, h);
    return check_launch(d_hist ? "conv_halo_kernel<head+hist>" : "conv_halo_kernel<head>");
}
}  // namespace
}  // namespace im2im

extern "C" int im2im_head_conv3x3_tc_f32(const void* d_x, const void* d_weight, const float* d_bias, int32_t B, int32_t H,
                                         int32_t W, int32_t n_real, int32_t act_kind, int32_t act_from_plane,
                                         float* d_out, void* stream) {
    if (!d_out) return fail(IM2IM_EINVAL, "head_tc: null tensor");
    return head_tc_impl(d_x, d_weight, d_bias, nullptr, B, H, W, n_real, act_kind, act_from_plane, d_out, nullptr, nullptr, 0,
                        nullptr, stream);
}

extern "C" int im2im_head_conv3x3_tc_folded_f32(const void* d_x, const void* d_weight, const float* d_bias,
                                                const float* d_tap_bias, int32_t B, int32_t H, int32_t W, int32_t n_real,
                                                int32_t act_kind, int32_t act_from_plane, float* d_out, void* stream) {
    if (!d_out) return fail(IM2IM_EINVAL, "head_tc: null tensor");
    return head_tc_impl(d_x, d_weight, d_bias, d_tap_bias, B, H, W, n_real, act_kind, act_from_plane, d_out, nullptr, nullptr,
                        0, nullptr, stream);
}

extern "C" int im2im_head_conv3x3_tc_hist(const void* d_x, const void* d_weight, const float* d_bias,
                                          const float* d_tap_bias_or_null, int32_t B, int32_t H, int32_t W, int32_t n_real,
                                          int32_t act_kind, int32_t act_from_plane, float* d_out_or_null,
                                          const float* d_labels, const float* d_lambdas_sorted, int32_t n_lambdas,
                                          uint32_t* d_hist, void* stream) {
    if (!d_hist) return fail(IM2IM_EINVAL, "head_tc_hist: null histogram");
    return head_tc_impl(d_x, d_weight, d_bias, d_tap_bias_or_null, B, H, W, n_real, act_kind, act_from_plane, d_out_or_null,
                        d_labels, d_lambdas_sorted, n_lambdas, d_hist, stream);
}

extern "C" int im2im_conv_wgrad_bf16(const void* d_x, const void* d_dz, int32_t B, int32_t H, int32_t W, int32_t c_in,
                                     int32_t c_out, int32_t taps, float* d_dw, void* stream) {
    if (taps != 9 && taps != 1) return fail(IM2IM_EINVAL, "taps must be 9 or 1");
    if (B <= 0 || H <= 0 || W <= 0) return fail(IM2IM_EINVAL, "bad activation shape");
    if (c_in <= 0 || c_in % kKStep) return fail(IM2IM_ERANGE, "c_in must be a multiple of %d (got %d)", kKStep, c_in);
    if (c_out <= 0 || c_out % kKStep) return fail(IM2IM_ERANGE, "c_out must be a multiple of %d (got %d)", kKStep, c_out);
    if (!d_x || !d_dz || !d_dw) return fail(IM2IM_EINVAL, "null tensor");
    // wide layers: halo kernel (one X box per channel block serves all nine taps; transposed GEMM, no wasted MMA rows)
    static const bool no_halo = (getenv("IM2IM_WGRAD_NO_HALO") != nullptr);
    static const bool halo_any = (getenv("IM2IM_WGRAD_HALO_ANY") != nullptr);   // experiment: partial tiles (TMA zero fill)
    if (!no_halo && taps == 9 && ((W % kHaloTileW == 0 && H % kHaloTileH == 0) || halo_any)) {
        WgradHaloParams h;
        h.c_in = c_in; h.c_out = c_out; h.B = B; h.H = H; h.W = W;
        h.tiles_w = (W + kHaloTileW - 1) / kHaloTileW; h.tiles_h = (H + kHaloTileH - 1) / kHaloTileH;
        h.cblocks = c_in / kKStep;
        // N = 128 (two dZ slabs per stage, taps dealt to two kinds of CTA) where c_out allows it and a work item gets at least
        // five CTAs to split 2 : 3; IM2IM_WGRAD_HALO_WIDE=0 keeps N = 64
        const char* wide_e = getenv("IM2IM_WGRAD_HALO_WIDE");
        const int n_pix_tiles = h.tiles_w * h.tiles_h * B;
        // (64 input channels: one channel block per work item, measured 3 % slower than N = 64 at 64 -> 128 @160^2)
        h.wide = (c_out % 128 == 0 && c_in >= 128 && !(wide_e != nullptr && wide_e[0] == '0')) ? 1 : 0;
        int n_blocks = 0;
        long long splits = 0;
        for (;;) {
            n_blocks = c_out / (h.wide ? 128 : 64);
            const long long items = static_cast<long long>(h.cblocks) * n_blocks;
            splits = sm_count() / items;         // at most ONE wave of CTAs (rounding up would leave a second, nearly empty
                                                 // wave); every split also costs 36.8k fp32 atomics
            if (splits > n_pix_tiles) splits = n_pix_tiles;
            if (splits < 1) splits = 1;
            if (h.wide && splits < 5) { h.wide = 0; continue; }   // too few CTAs per work item to deal 2 : 3
            break;
        }
        h.splits = static_cast<int>(splits);
        h.splits0 = 0;
        if (h.wide) {
            h.splits0 = static_cast<int>((2 * splits + 2) / 5);   // 2 : 3 = MMAs per pixel tile of the two kinds
            if (h.splits0 < 1) h.splits0 = 1;
            if (h.splits0 > h.splits - 1) h.splits0 = h.splits - 1;
        }
        h.stages = h.wide ? 3 : 4;
        h.dw = d_dw;
        CUtensorMap mdz, mx;
        int rc = make_act_map(&mdz, d_dz, B, H, W, c_out, kHaloTileW, kHaloTileH, 1);
        if (rc) return rc;
        rc = make_act_map(&mx, d_x, B, H, W, c_in, kHaloW, kHaloH, 1);
        if (rc) return rc;
        const size_t smem = static_cast<size_t>(h.stages) * (kWgHaloStageBytes + (h.wide ? kSlabBytes : 0)) +
                            (2 * h.stages + 1) * sizeof(uint64_t) + 16 + 1024;
        IM2IM_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(static_cast<unsigned>(h.cblocks * h.splits), static_cast<unsigned>(n_blocks));
        conv_wgrad_halo_kernel<<<grid, kConvThreads, smem, static_cast<cudaStream_t>(stream)>>>(mdz, mx, h);
        return check_launch("conv_wgrad_halo_kernel");
    }
    WgradParams p;
    p.taps = taps; p.c_in = c_in; p.c_out = c_out; p.B = B; p.H = H; p.W = W;
    pick_box(B, H, W, p.bw, p.bh, p.bb);
    p.tiles_w = (W + p.bw - 1) / p.bw; p.tiles_h = (H + p.bh - 1) / p.bh; p.tiles_b = (B + p.bb - 1) / p.bb;
    p.n_slabs = taps * c_in / kKStep;
    p.slab_groups = (p.n_slabs + 3) / 4;
    const int m_blocks = (c_out + kTileM - 1) / kTileM;
    const int out_tiles = p.slab_groups * m_blocks;
    const int n_pix_tiles = p.tiles_w * p.tiles_h * p.tiles_b;
    int splits = (2 * sm_count() + out_tiles - 1) / out_tiles;   // ~2 CTAs worth of work items per SM
    if (splits > n_pix_tiles) splits = n_pix_tiles;
    if (splits < 1) splits = 1;
    p.splits = splits;
    p.stages = 2;
    p.dw = d_dw;
    CUtensorMap mdz, mx;
    int rc = make_act_map(&mdz, d_dz, B, H, W, c_out, p.bw, p.bh, p.bb);
    if (rc) return rc;
    rc = make_act_map(&mx, d_x, B, H, W, c_in, p.bw, p.bh, p.bb);
    if (rc) return rc;
    const size_t smem = static_cast<size_t>(p.stages) * 6 * kSlabBytes + (2 * p.stages + 1) * sizeof(uint64_t) + 16 + 1024;
    IM2IM_CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(static_cast<unsigned>(p.slab_groups * splits), static_cast<unsigned>(m_blocks));
    conv_wgrad_kernel<<<grid, kConvThreads, smem, static_cast<cudaStream_t>(stream)>>>(mdz, mx, p);
    return check_launch("conv_wgrad_kernel");
}
