// Micro-benchmark of K1 load strategies (development tool, not part of the library).
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/k1v tools/k1_variants.cu && /tmp/k1v [rows]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ float4 ld_na(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_ef(const float4 *p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float score_of(float s0, float s1) {
    const float d0 = 1.0f - s0, d1 = 1.5f * (1.0f - s1);
    return 1.0f - __fsqrt_rn((d0 * d0 + d1 * d1) * (1.0f / 3.25f));
}

// V0/V2: one row per warp iteration, target in registers.  HINT: 0 none, 1 L2 evict_first policy
template <int THREADS, int MINB, int HINT>
__global__ void __launch_bounds__(THREADS, MINB) k_ldg(const float4 *__restrict__ rows, const float4 *__restrict__ target, long long n, float *scores) {
    const int lane = threadIdx.x & 31;
    const long long w0 = (long long)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5), nw = (long long)gridDim.x * (THREADS / 32);
    uint64_t pol = 0;
    if (HINT) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    float4 t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = target[i * 32 + lane];
    for (long long r = w0; r < n; r += nw) {
        const float4 *p = rows + r * 512 + lane;
        float4 x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = HINT ? ld_ef(p + i * 32, pol) : ld_na(p + i * 32);
        float s[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                a0 = fmaf(x[h * 8 + i].x, t[h * 8 + i].x, a0); a1 = fmaf(x[h * 8 + i].y, t[h * 8 + i].y, a1);
                a2 = fmaf(x[h * 8 + i].z, t[h * 8 + i].z, a2); a3 = fmaf(x[h * 8 + i].w, t[h * 8 + i].w, a3);
            }
            s[h] = warp_sum((a0 + a1) + (a2 + a3));
        }
        if (lane == 0) scores[r] = score_of(s[0], s[1]);
    }
}

// V3: cp.async.bulk ring.  ROWS rows per stage, STAGES stages, ROWS consumer warps + 1 producer warp.
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t a, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint32_t a, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int ROWS, int STAGES>
__global__ void __launch_bounds__((ROWS + 1) * 32, 1) k_bulk(const float *__restrict__ rows, const float4 *__restrict__ target, long long n, float *scores) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr uint32_t STAGE_BYTES = ROWS * 8192;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * STAGE_BYTES);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[STAGES + s]), ROWS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long n_tiles = (n + ROWS - 1) / ROWS;
    if (warp == ROWS) {
        if (lane == 0) {
            int it = 0;
            for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                mbar_wait(smem_u32(&bars[STAGES + s]), ph ^ 1);
                const long long r0 = t * ROWS;
                const uint32_t bytes = (uint32_t)((n - r0 < ROWS ? n - r0 : ROWS) * 8192);
                mbar_expect(smem_u32(&bars[s]), bytes);
                bulk_g2s(smem_u32(smem + (size_t)s * STAGE_BYTES), rows + r0 * 2048, bytes, smem_u32(&bars[s]));
            }
        }
        return;
    }
    float4 t[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = target[i * 32 + lane];
    int it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(smem_u32(&bars[s]), ph);
        const long long r = tile * ROWS + warp;
        const float4 *p = reinterpret_cast<const float4 *>(smem + (size_t)s * STAGE_BYTES + (size_t)warp * 8192) + lane;
        float sres[2] = {0.f, 0.f};
        if (r < n) {
            float4 x[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = p[i * 32];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    a0 = fmaf(x[h * 8 + i].x, t[h * 8 + i].x, a0); a1 = fmaf(x[h * 8 + i].y, t[h * 8 + i].y, a1);
                    a2 = fmaf(x[h * 8 + i].z, t[h * 8 + i].z, a2); a3 = fmaf(x[h * 8 + i].w, t[h * 8 + i].w, a3);
                }
                sres[h] = (a0 + a1) + (a2 + a3);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[STAGES + s]));
        const float s0 = warp_sum(sres[0]), s1 = warp_sum(sres[1]);
        if (lane == 0 && r < n) scores[r] = score_of(s0, s1);
    }
}

template <class F> float time_ms(F f, int iters) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 3; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) f();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms / iters;
}

__global__ void fill(float *p, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = (float)((i * 2654435761u) & 1023) * (1.0f / 1024.0f);
}

int main(int argc, char **argv) {
    const long long n = argc > 1 ? atoll(argv[1]) : 1000000;
    const int iters = argc > 2 ? atoi(argv[2]) : 20;
    float *rows, *scores, *ref; float4 *target;
    CK(cudaMalloc(&rows, n * 8192)); CK(cudaMalloc(&scores, n * 4)); CK(cudaMalloc(&ref, n * 4)); CK(cudaMalloc(&target, 8192));
    fill<<<148 * 8, 256>>>(rows, n * 2048); fill<<<8, 256>>>((float *)target, 2048);
    CK(cudaDeviceSynchronize());
    const double gb = n * 8192.0 / 1e9;
    auto report = [&](const char *name, float ms, bool check) {
        int bad = -1;
        if (check) {
            std::vector<float> a(1000), b(1000);
            CK(cudaMemcpy(a.data(), scores + (n - 1000), 4000, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(b.data(), ref + (n - 1000), 4000, cudaMemcpyDeviceToHost));
            bad = 0; for (int i = 0; i < 1000; ++i) bad += (a[i] != b[i]);
        }
        printf("%-34s %8.4f ms  %8.1f GB/s  mismatches %d\n", name, ms, gb / (ms * 1e-3), bad);
    };
    const float4 *r4 = (const float4 *)rows;
    { float ms = time_ms([&] { k_ldg<128, 3, 0><<<148 * 3, 128>>>(r4, target, n, ref); }, iters); report("ldg 128x3 grid 444 (current)", ms, false); }
    { float ms = time_ms([&] { k_ldg<128, 3, 0><<<148 * 6, 128>>>(r4, target, n, scores); }, iters); report("ldg 128x3 grid 888", ms, true); }
    { float ms = time_ms([&] { k_ldg<128, 3, 0><<<148 * 24, 128>>>(r4, target, n, scores); }, iters); report("ldg 128x3 grid 3552", ms, true); }
    { float ms = time_ms([&] { k_ldg<256, 1, 0><<<148, 256>>>(r4, target, n, scores); }, iters); report("ldg 256x1 grid 148", ms, true); }
    { float ms = time_ms([&] { k_ldg<64, 6, 0><<<148 * 6, 64>>>(r4, target, n, scores); }, iters); report("ldg 64x6 grid 888", ms, true); }
    { float ms = time_ms([&] { k_ldg<128, 3, 1><<<148 * 3, 128>>>(r4, target, n, scores); }, iters); report("ldg 128x3 L2 evict_first", ms, true); }
    {
        constexpr int R = 8, S = 3; const size_t sm = (size_t)R * S * 8192 + 128;
        CK(cudaFuncSetAttribute(k_bulk<R, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        float ms = time_ms([&] { k_bulk<R, S><<<148, (R + 1) * 32, sm>>>(rows, target, n, scores); }, iters); report("bulk ring 8 rows x 3 stages", ms, true);
    }
    {
        constexpr int R = 4, S = 6; const size_t sm = (size_t)R * S * 8192 + 128;
        CK(cudaFuncSetAttribute(k_bulk<R, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        float ms = time_ms([&] { k_bulk<R, S><<<148, (R + 1) * 32, sm>>>(rows, target, n, scores); }, iters); report("bulk ring 4 rows x 6 stages", ms, true);
    }
    {
        constexpr int R = 4, S = 3; const size_t sm = (size_t)R * S * 8192 + 128;
        CK(cudaFuncSetAttribute(k_bulk<R, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        float ms = time_ms([&] { k_bulk<R, S><<<148 * 2, (R + 1) * 32, sm>>>(rows, target, n, scores); }, iters); report("bulk ring 4x3, 2 CTA/SM", ms, true);
    }
    {
        constexpr int R = 2, S = 4; const size_t sm = (size_t)R * S * 8192 + 128;
        CK(cudaFuncSetAttribute(k_bulk<R, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        float ms = time_ms([&] { k_bulk<R, S><<<148 * 3, (R + 1) * 32, sm>>>(rows, target, n, scores); }, iters); report("bulk ring 2x4, 3 CTA/SM", ms, true);
    }
    return 0;
}
