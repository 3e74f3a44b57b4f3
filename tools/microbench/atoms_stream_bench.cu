// Microbenchmarks that decided the RCPS miss-count kernel design:
//   (1) shared-memory atomic throughput under the address patterns a 1001-bin histogram produces
//   (2) 4-plane HBM streaming throughput: plain LDG.128 vs cp.async.bulk (UBLKCP) staged through smem
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o atoms_stream_bench atoms_stream_bench.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

// ---------------------------------------------------------------- (1) ATOMS patterns
// mode 0: lane-private copies (addr = bin*32+lane)  -> bank conflict free
// mode 1: single copy, random bin per lane            -> random bank conflicts
// mode 2: 8 copies (addr = bin*8 + (lane&7))
// mode 3: all lanes same address
// mode 4: non-atomic LDS/IADD/STS on lane-private copy (racy across warps; throughput probe only)
// mode 5: 4 copies
// mode 6: 16 copies
template <int MODE>
__global__ void __launch_bounds__(1024, 1) atoms_kernel(unsigned* out, int iters, int nbins) {
    extern __shared__ unsigned hist[];
    const int copies = (MODE == 0 || MODE == 4) ? 32 : (MODE == 2 ? 8 : (MODE == 5 ? 4 : (MODE == 6 ? 16 : 1)));
    for (int i = threadIdx.x; i < nbins * copies; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    const unsigned lane = threadIdx.x & 31;
    #pragma unroll 4
    for (int it = 0; it < iters; ++it) {
        x = x * 1664525u + 1013904223u;
        unsigned bin = __umulhi(x, (unsigned)nbins);
        unsigned addr;
        if (MODE == 0 || MODE == 4) addr = bin * 32 + lane;
        else if (MODE == 1) addr = bin;
        else if (MODE == 2) addr = bin * 8 + (lane & 7);
        else if (MODE == 5) addr = bin * 4 + (lane & 3);
        else if (MODE == 6) addr = bin * 16 + (lane & 15);
        else addr = 7;
        if (MODE == 4) hist[addr] = hist[addr] + 1;
        else atomicAdd(&hist[addr], 1u);
    }
    __syncthreads();
    unsigned s = 0;
    for (int i = threadIdx.x; i < nbins * copies; i += blockDim.x) s += hist[i];
    if (s == 0xffffffffu) out[0] = s;   // keep live
    if (threadIdx.x == 0) atomicAdd(&out[1], s);
}

template <int MODE>
static void run_atoms(const char* name, int threads, int nbins, unsigned* d_out) {
    const int copies = (MODE == 0 || MODE == 4) ? 32 : (MODE == 2 ? 8 : (MODE == 5 ? 4 : (MODE == 6 ? 16 : 1)));
    size_t smem = (size_t)nbins * copies * 4;
    CK(cudaFuncSetAttribute(atoms_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 4096, grid = 148 * 2;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    atoms_kernel<MODE><<<grid, threads, smem>>>(d_out, iters, nbins);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    atoms_kernel<MODE><<<grid, threads, smem>>>(d_out, iters, nbins);
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    double total = (double)grid * threads * iters;
    // 2 CTAs per SM serialised (1 CTA/SM resident) -> per-SM rate = total/148/time
    printf("ATOMS %-34s threads=%4d bins=%4d  %8.3f ms  %7.2f Gupd/s chip  %6.3f upd/ns/SM\n",
           name, threads, nbins, ms, total / ms * 1e-6, total / 148.0 / (ms * 1e6));
}

// ---------------------------------------------------------------- (2) streaming
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// plain LDG.128: grid-stride over float4 groups, 4 planes, UNROLL groups in flight
template <int UNROLL>
__global__ void __launch_bounds__(512) stream_ldg(const float4* __restrict__ a, const float4* __restrict__ b,
                                                   const float4* __restrict__ c, const float4* __restrict__ d,
                                                   size_t n4, float* out) {
    float acc = 0.f;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
        float4 va[UNROLL], vb[UNROLL], vc[UNROLL], vd[UNROLL];
        #pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            va[u] = ldg_stream(a + i + u * stride); vb[u] = ldg_stream(b + i + u * stride);
            vc[u] = ldg_stream(c + i + u * stride); vd[u] = ldg_stream(d + i + u * stride);
        }
        #pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            acc += va[u].x + vb[u].y + vc[u].z + vd[u].w + va[u].w * vb[u].x + vc[u].y * vd[u].z;
    }
    for (; i < n4; i += stride) acc += ldg_stream(a + i).x + ldg_stream(b + i).y + ldg_stream(c + i).z + ldg_stream(d + i).w;
    if (acc == 1.2345f) out[0] = acc;
}

// cp.async.bulk staged: persistent CTAs, STAGES-deep ring of 4-plane tiles of TILE floats each
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int TILE, int STAGES>
__global__ void __launch_bounds__(512) stream_bulk(const float* __restrict__ a, const float* __restrict__ b,
                                                    const float* __restrict__ c, const float* __restrict__ d,
                                                    size_t n, float* out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* buf = reinterpret_cast<float*>(smem_raw);                       // [STAGES][4][TILE]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)STAGES * 4 * TILE * 4);
    uint64_t* empty = full + STAGES;
    const size_t ntiles = n / TILE;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], blockDim.x / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = blockDim.x >> 5;
    // number of tiles this CTA handles
    size_t my = 0; for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) ++my;
    float acc = 0.f;
    // producer = thread 0 of warp 0 (also consumes); prefill
    size_t issued = 0;
    if (threadIdx.x == 0) {
        for (; issued < my && issued < STAGES; ++issued) {
            size_t t = blockIdx.x + issued * gridDim.x; int s = issued % STAGES;
            mbar_expect_tx(&full[s], 4 * TILE * 4);
            bulk_g2s(buf + (s * 4 + 0) * TILE, a + t * TILE, TILE * 4, &full[s]);
            bulk_g2s(buf + (s * 4 + 1) * TILE, b + t * TILE, TILE * 4, &full[s]);
            bulk_g2s(buf + (s * 4 + 2) * TILE, c + t * TILE, TILE * 4, &full[s]);
            bulk_g2s(buf + (s * 4 + 3) * TILE, d + t * TILE, TILE * 4, &full[s]);
        }
    }
    for (size_t k = 0; k < my; ++k) {
        int s = k % STAGES; unsigned ph = (k / STAGES) & 1;
        mbar_wait(&full[s], ph);
        const float4* pa = reinterpret_cast<const float4*>(buf + (s * 4 + 0) * TILE);
        const float4* pb = reinterpret_cast<const float4*>(buf + (s * 4 + 1) * TILE);
        const float4* pc = reinterpret_cast<const float4*>(buf + (s * 4 + 2) * TILE);
        const float4* pd = reinterpret_cast<const float4*>(buf + (s * 4 + 3) * TILE);
        for (int i = threadIdx.x; i < TILE / 4; i += blockDim.x) {
            float4 va = pa[i], vb = pb[i], vc = pc[i], vd = pd[i];
            acc += va.x + vb.y + vc.z + vd.w + va.w * vb.x + vc.y * vd.z;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (threadIdx.x == 0 && issued < my) {
            // refill the slot consumed STAGES-? ago: slot of tile `issued` is issued % STAGES == s only when issued == k+STAGES
            int s2 = issued % STAGES; unsigned ph2 = ((issued / STAGES) - 1) & 1;
            mbar_wait(&empty[s2], ph2);
            size_t t = blockIdx.x + issued * gridDim.x;
            mbar_expect_tx(&full[s2], 4 * TILE * 4);
            bulk_g2s(buf + (s2 * 4 + 0) * TILE, a + t * TILE, TILE * 4, &full[s2]);
            bulk_g2s(buf + (s2 * 4 + 1) * TILE, b + t * TILE, TILE * 4, &full[s2]);
            bulk_g2s(buf + (s2 * 4 + 2) * TILE, c + t * TILE, TILE * 4, &full[s2]);
            bulk_g2s(buf + (s2 * 4 + 3) * TILE, d + t * TILE, TILE * 4, &full[s2]);
            ++issued;
        }
    }
    (void)warp; (void)nwarps;
    if (acc == 1.2345f) out[0] = acc;
}

static float time_it(void (*launch)(void*), void* ctx) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) launch(ctx);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < 5; ++i) {
        CK(cudaEventRecord(a)); launch(ctx); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    return best;
}

struct StreamCtx { const float *a, *b, *c, *d; size_t n; float* out; int grid; int threads; };

template <int U> static void launch_ldg(void* p) {
    StreamCtx* s = (StreamCtx*)p;
    stream_ldg<U><<<s->grid, s->threads>>>((const float4*)s->a, (const float4*)s->b, (const float4*)s->c, (const float4*)s->d, s->n / 4, s->out);
}
template <int TILE, int STAGES> static void launch_bulk(void* p) {
    StreamCtx* s = (StreamCtx*)p;
    size_t smem = (size_t)STAGES * 4 * TILE * 4 + 2 * STAGES * 8;
    CK(cudaFuncSetAttribute(stream_bulk<TILE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stream_bulk<TILE, STAGES><<<s->grid, s->threads, smem>>>(s->a, s->b, s->c, s->d, s->n, s->out);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s  SMs=%d  smem/block optin=%zu\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
    unsigned* d_out; CK(cudaMalloc(&d_out, 64)); CK(cudaMemset(d_out, 0, 64));
    for (int threads : {256, 512, 1024}) {
        run_atoms<0>("lane-private x32 (conflict-free)", threads, 1001, d_out);
        run_atoms<6>("16 copies", threads, 1001, d_out);
        run_atoms<2>("8 copies", threads, 1001, d_out);
        run_atoms<5>("4 copies", threads, 1001, d_out);
        run_atoms<1>("single copy random bin", threads, 1001, d_out);
        run_atoms<3>("same address", threads, 1001, d_out);
        run_atoms<4>("non-atomic LDS/STS lane-private", threads, 1001, d_out);
    }
    // streaming: 4 planes x 1 GiB each (256 Mi floats) -> 4 GiB per pass, >> L2
    size_t n = (size_t)256 << 20;
    float *a, *b, *c, *d, *fo;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&c, n * 4)); CK(cudaMalloc(&d, n * 4)); CK(cudaMalloc(&fo, 64));
    CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4)); CK(cudaMemset(c, 0, n * 4)); CK(cudaMemset(d, 0, n * 4));
    double bytes = (double)n * 16;
    for (int threads : {256, 512}) for (int mult : {2, 4, 8}) {
        StreamCtx s{a, b, c, d, n, fo, 148 * mult, threads};
        float ms1 = time_it(launch_ldg<1>, &s), ms2 = time_it(launch_ldg<2>, &s), ms4 = time_it(launch_ldg<4>, &s);
        printf("STREAM ldg.128 threads=%d grid=148x%d  U1 %.1f GB/s  U2 %.1f GB/s  U4 %.1f GB/s\n", threads, mult,
               bytes / ms1 * 1e-6, bytes / ms2 * 1e-6, bytes / ms4 * 1e-6);
    }
    for (int threads : {256, 512}) for (int mult : {1, 2}) {
        StreamCtx s{a, b, c, d, n, fo, 148 * mult, threads};
        float m1 = time_it(launch_bulk<1024, 4>, &s), m2 = time_it(launch_bulk<2048, 4>, &s), m3 = time_it(launch_bulk<2048, 3>, &s), m4 = time_it(launch_bulk<1024, 6>, &s);
        printf("STREAM bulk    threads=%d grid=148x%d  T1024xS4 %.1f  T2048xS4 %.1f  T2048xS3 %.1f  T1024xS6 %.1f GB/s\n", threads, mult,
               bytes / m1 * 1e-6, bytes / m2 * 1e-6, bytes / m3 * 1e-6, bytes / m4 * 1e-6);
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
