// K3 — batched queries on tcgen05 with a two-term bf16 split ("bf16x2"), sm_100a: the kernel.
// Included by vq_batch.cu inside its anonymous namespace (reference ticket.py:146-180, 325-327 for Q tickets at once).
//
// Arithmetic.  fp32 operands are split into two bf16 terms, round to nearest both times:
//     x = x1 + x2 (+ r_x),  t = t1 + t2 (+ r_t)       x1 = bf16(x), x2 = bf16(x - x1), same for t
//     x*t ~= x1*t1 + x2*t1 + x1*t2                    |dropped| <= 3 * 2^-18 |x*t| per product, signs random
// i.e. three kind::f16 MMAs (M128 N256 K16, bf16 in, fp32 accumulate) per 16 dims.  Measured against float64
// (profiles/): score error mean -5.3e-7, max 1.0e-6 — the same as the 3xTF32 kernel of v1-v3 at half its tensor-pipe
// cycles and half its operand bytes.  The tensor core truncates on every accumulate into TMEM, so accumulation is
// two-level: a partial accumulator takes GROUP_KB * 6 = 24 MMAs, then the epilogue drains it into fp32 running
// sums in registers (round-to-nearest adds).
//
// Tiling.  One CTA per SM, persistent over 128-clip tiles; all 256 queries of a pass share ONE read and ONE
// conversion of the clip tile.  Per K block of 32 dims:
//     TMA  -> fp32 clip tile [128 x 32] (128B swizzle, ring of 3)         -> converter warps -> x1, x2 bf16 tiles
//     TMA  -> t1, t2 bf16 query tiles [256 x 32] (64B swizzle, ring of 3)    [128 x 32] (64B swizzle, ring of 4)
//     6 x tcgen05.mma into the partial accumulator
// TMEM (512 columns) holds two partial accumulators of 256 columns: the MMA warp fills one while the epilogue drains
// the other.  The earlier streams' terms (w_s (1 - sim_s))^2 of every (clip, query) are parked in an L2-resident
// scratch (128 KB per CTA, float4 per thread, evict-last) until the last stream is done — registers hold the 128
// running sums per thread and TMEM the accumulators; neither has room.  Clip tiles are loaded evict-first, query
// tiles evict-last, so the 3 TB/s clip stream does not push the park and the query operand out of L2.
//
// Warps (512 threads; setmaxnreg moves registers from warpgroups 0-1 to the epilogue warpgroups 2-3):
//     warp 0       TMA producer, clip tiles            warp 2       TMEM allocation, TMA producer for t1 / t2
//     warp 1       MMA issuer (one elected lane)       warp 3       idle
//     warps 4-7    converters, one per SM sub-partition; the loads of K block i+1 are issued before block i is
//                  converted and stored
//     warps 8-15   epilogue: drains, running sums, scores, per-query counts (bit-mask transpose), top-k candidates
// Development history and the measurements behind these choices: profiles/r1_k3_batched_notes.md.
//
// Round 2 experiment, NOT the default (build with -DVQ_BATCH_F16F8): fp16 + fp8 operand split.
//     x' = x * bx * 2^6,  t' = t * bt * 2^6          bx (per store and stream), bt (per pass and stream): powers of two with
//                                                    max |x| bx <= 128, max |t| bt <= 128
//     x1 = fp16(x'), rx = x' - x1 (exact, |rx| <= 2^-11 |x'| <= 4);      t1, rt likewise
//     x*t * (bx bt 2^12) ~= x1*t1  +  e4m3(rx 2^6) * e4m3(t bt)  +  e4m3(x bx) * e4m3(rt 2^6)
// i.e. ONE kind::f16 MMA (K = 16) plus ONE kind::f8f6f4 MMA over a doubled K (32 e4m3 per 16 dims: [rx 2^6 | x bx] against
// [t bt | rt 2^6]) per 16 dims: two thirds of the tensor-pipe cycles of the bf16x3 split.  All three products carry the same
// power of two, so they share one accumulator and the epilogue multiplies the stream's sum by 1 / (4096 bx bt) (exact).
// The residual operand tiles have the byte layout of the old x2 / t2 tiles (64 B per row and K block, 64B swizzle, +32 B per
// K step), and the instruction descriptor is numerically the same for both kinds (format code 0 = F16 / E4M3).
// CPU model of the arithmetic (tests/probes/k3_split_schemes.py): score error max 2.2e-6, rms 4.3e-7 (bf16x3: 8.3e-7 / 1.8e-7).
// Measured on B200 (profiles/r2_k3_f16f8_notes.md): parity-green on VQSYN-1 (mean -5.1e-7, max 1.5e-6 at 1M x 256) but NOT faster
// — 3.52 ms vs 3.50 ms at 1M x 256, 34.3 ms (1537 MHz) vs 33.0 ms (1380 MHz) at 10M x 256: the converter warps now execute 48
// F2FP + 32 HADD2.F32 per thread and K block instead of 32 F2FP (the conversion pipe becomes the limiter as the tensor
// pipe stops being one), and on uniform random rows the error reaches 1.3e-5.  Kept for the record; bf16x3 stays.
#pragma once

namespace bf {

constexpr int BM = 128;                  // clips per tile (UMMA M)
constexpr int QN = 256;                  // queries per pass (max UMMA N)
constexpr int BK = 32;                   // dims per K block: 128 B of fp32, 64 B of bf16
constexpr int UK = 16;                   // UMMA K for bf16
#ifndef VQ_BF_GROUP_KB
#define VQ_BF_GROUP_KB 4
#endif
constexpr int GROUP_KB = VQ_BF_GROUP_KB; // K blocks per partial accumulator
constexpr uint32_t A32_BYTES = BM * BK * 4;      // 16 KB
constexpr uint32_t X_BYTES = BM * BK * 2;        //  8 KB
constexpr int THREADS = 512;
constexpr int CONV_WARPS = 4;
constexpr int EPI_WARPS = 8;
// Shared-memory rings, by query tier QT (rows of a t1 / t2 tile: 256, or 128 for batches of <= 128 queries).  With
// 128-row query tiles the t ring is half as large and the fp32 ring twice as deep: small batches are bound by the
// bytes in flight towards HBM (3 stages x 16 KB per SM against ~1.5 us of latency is only ~3.5 TB/s chip-wide).
template <int QT>
struct Ring {
    static constexpr int NA = QT > 128 ? 3 : 6;      // fp32 clip-tile ring: TMA -> converters (released as soon as they have read it)
    static constexpr int NXR = 4;                    // x1 / x2 ring: converters -> MMA (released by the MMA commits)
    static constexpr int NT = 3;                     // t1 / t2 ring: TMA -> MMA (released by the MMA commits)
    static constexpr uint32_t T_BYTES = QT * BK * 2; // 16 KB / 8 KB per t1 (or t2) tile
    static constexpr uint32_t RING_X = NA * A32_BYTES;                  // x1 at +0, x2 at +X_BYTES of a stage
    static constexpr uint32_t RING_T = RING_X + NXR * 2 * X_BYTES;      // t1 at +0, t2 at +T_BYTES of a stage
    static constexpr uint32_t RING_END = RING_T + NT * 2 * T_BYTES;     // 48 + 64 + 96 = 208 KB / 96 + 64 + 48 = 208 KB
    static constexpr int N_BARS = 2 * NA + 2 * NXR + 2 * NT + 4;
    static constexpr size_t SMEM = (size_t)RING_END + 1024 /*align*/ + 256 /*barriers + tmem slot*/ + QN * 4 /*cut*/ +
                                   EPI_WARPS * 128 * 2 * 4 /*per-warp counts*/;
};
// shared-memory descriptor high word: SBO = 512 B (8 rows of 64 B), descriptor version 1, SWIZZLE_64B
constexpr uint32_t DESC_HI64 = (512u >> 4) | (1u << 14) | (4u << 29);

template <bool kAcc>
__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %3, p;\n}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc),
        "n"(kAcc ? 1 : 0), "r"(DESC_HI64) : "memory");
}
template <bool kAcc>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 da, db;\nsetp.ne.b32 p, %4, 0;\nmov.b64 da, {%1, %5};\nmov.b64 db, {%2, %5};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n}" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc),
        "n"(kAcc ? 1 : 0), "r"(DESC_HI64) : "memory");
}
// L2 residency control.  The clip tiles are a read-once stream (evict-first); the query tiles and the park are
// re-read every tile (evict-last), otherwise the 3 TB/s clip stream pushes them out of L2 between two uses.
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar, uint64_t pol) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 ld_park(const float4 *p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_park(float4 *p, const float4 &v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float hi, float lo) {     // {bf16_rn(hi), bf16_rn(lo)}: lo in bits 0-15
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// 8 floats -> 8 bf16 leading parts (p) and 8 bf16 residual parts (q); x - float(x1) is exact in fp32
__device__ __forceinline__ void split8(const float4 &u, const float4 &v, uint4 &p, uint4 &q) {
#define VQ_SPLIT2(f0, f1, P, Q)                                                                  \
    {                                                                                            \
        P = pack_bf16x2(f1, f0);                                                                 \
        const float h0 = __uint_as_float(P << 16), h1 = __uint_as_float(P & 0xFFFF0000u);        \
        Q = pack_bf16x2(f1 - h1, f0 - h0);                                                       \
    }
    VQ_SPLIT2(u.x, u.y, p.x, q.x)
    VQ_SPLIT2(u.z, u.w, p.y, q.y)
    VQ_SPLIT2(v.x, v.y, p.z, q.z)
    VQ_SPLIT2(v.z, v.w, p.w, q.w)
#undef VQ_SPLIT2
}

// fp16 + fp8 split of 8 scaled floats (x' = x * sx, sx = bx * 2^6): p = 8 fp16 leading parts (16 B), rq = 8 e4m3 of the
// residuals (x' - x1) * 2^6 (8 B), xq = 8 e4m3 of x * bx = x' * 2^-6 (8 B).  Element order inside each word: lower dims in
// the lower bits (little endian), like the bf16 split above.
__device__ __forceinline__ uint32_t pack_f16x2(float hi, float lo) {
    uint32_t d;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t pack_e4m3x4(float f0, float f1, float f2, float f3) {
    uint16_t lo, hi;
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(f1), "f"(f0));
    asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(f3), "f"(f2));
    return (uint32_t)lo | ((uint32_t)hi << 16);
}
__device__ __forceinline__ void split8_f16f8(const float4 &u, const float4 &v, const float sx, uint4 &p, uint2 &rq, uint2 &xq) {
    const float f[8] = {u.x * sx, u.y * sx, u.z * sx, u.w * sx, v.x * sx, v.y * sx, v.z * sx, v.w * sx};
    float r[8];
    uint32_t pw[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pw[i] = pack_f16x2(f[2 * i + 1], f[2 * i]);
        float h0, h1;
        asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\ncvt.f32.f16 %0, lo;\ncvt.f32.f16 %1, hi;\n}" : "=f"(h0), "=f"(h1) : "r"(pw[i]));
        r[2 * i] = (f[2 * i] - h0) * 64.0f;
        r[2 * i + 1] = (f[2 * i + 1] - h1) * 64.0f;
    }
    p = make_uint4(pw[0], pw[1], pw[2], pw[3]);
    rq = make_uint2(pack_e4m3x4(r[0], r[1], r[2], r[3]), pack_e4m3x4(r[4], r[5], r[6], r[7]));
    constexpr float k = 1.0f / 64.0f;
    xq = make_uint2(pack_e4m3x4(f[0] * k, f[1] * k, f[2] * k, f[3] * k), pack_e4m3x4(f[4] * k, f[5] * k, f[6] * k, f[7] * k));
}
__device__ __forceinline__ void sts64(uint32_t addr, const uint2 &v) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(v.x), "r"(v.y) : "memory");
}

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// 32 x 32 bit-matrix transpose across a warp: lane r holds row r (bit c = column c) -> lane c holds column c
// (bit r = row r).  Five block-swap steps; used to turn "my row passes query j" masks into per-query counts.
__device__ __forceinline__ unsigned int transpose32(unsigned int x, int lane) {
    unsigned int m = 0x0000FFFFu;
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const unsigned int y = __shfl_xor_sync(0xffffffffu, x, j);
        x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y << j) & ~m));
        m ^= m << (j >> 1);
    }
    return x;
}

// End of a stream for one epilogue thread (one clip row, 128 queries): term = (w (1 - sim))^2, plus the terms of
// the earlier streams from the park; the last stream leaves the sum in `run`, the others park it.
// The park lives in global memory, L2-resident (evict-last): per CTA [64 query groups of 4][128 rows] float4, so a warp
// moves 512 contiguous bytes per instruction.
template <bool kFirst, bool kLast>
__device__ __forceinline__ void stream_final(float (&run)[128], float4 *park, uint64_t pol, float w, float ic, int nch) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        if (c < nch) {
            float4 old[8];
            if constexpr (!kFirst) {
#pragma unroll
                for (int g = 0; g < 8; ++g) old[g] = ld_park(park + (size_t)(8 * c + g) * BM, pol);
            }
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float dd = w * (1.0f - run[32 * c + 4 * g + e] * ic);
                    v[e] = dd * dd;
                }
                if constexpr (!kFirst) { v[0] += old[g].x; v[1] += old[g].y; v[2] += old[g].z; v[3] += old[g].w; }
                if constexpr (kLast) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) run[32 * c + 4 * g + e] = v[e];
                } else {
                    st_park(park + (size_t)(8 * c + g) * BM, make_float4(v[0], v[1], v[2], v[3]), pol);
                }
            }
        }
    }
}

// run[base + j] for a warp-uniform runtime j in 0..31 without local memory: 31 selects
__device__ __forceinline__ float select32(const float (&run)[128], int base, int j) {
    float v16[16], v8[8], v4[4], v2[2];
#pragma unroll
    for (int i = 0; i < 16; ++i) v16[i] = (j & 16) ? run[base + 16 + i] : run[base + i];
#pragma unroll
    for (int i = 0; i < 8; ++i) v8[i] = (j & 8) ? v16[8 + i] : v16[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) v4[i] = (j & 4) ? v8[4 + i] : v8[i];
#pragma unroll
    for (int i = 0; i < 2; ++i) v2[i] = (j & 2) ? v4[2 + i] : v4[i];
    return (j & 1) ? v2[1] : v2[0];
}

// targets fp32 [n] -> t1, t2 bf16 (round to nearest both times)
__global__ void split_targets_bf16(const float *__restrict__ t, unsigned short *t1, unsigned short *t2, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = t[i];
    const uint32_t p = pack_bf16x2(0.f, x);
    const float h = __uint_as_float(p << 16);
    const uint32_t q = pack_bf16x2(0.f, x - h);
    t1[i] = (unsigned short)(p & 0xFFFFu);
    t2[i] = (unsigned short)(q & 0xFFFFu);
}

// targets fp32 [rows][K] (K = S * stream_len) -> t1 = fp16(t * st) and, per 16 dims, 32 bytes of t2:
// [e4m3(t * bt) x 16 | e4m3((t * st - t1) * 2^6) x 16]   (st = bt * 2^6, per stream); one thread per 16 dims.
__global__ void split_targets_f16f8(const float *__restrict__ t, unsigned short *t1, unsigned char *t2, long long n16, int stream_len,
                                    int n_streams, float st0, float st1, float st2, float st3) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n16) return;
    const long long e0 = g * 16;
    const int s = (int)((e0 / stream_len) % n_streams);
    const float st = s == 0 ? st0 : (s == 1 ? st1 : (s == 2 ? st2 : st3));
    uint32_t tq[4], rq[4];
    for (int i = 0; i < 4; ++i) {
        float f[4], r[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) f[j] = t[e0 + 4 * i + j] * st;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const uint32_t pw = pack_f16x2(f[2 * j + 1], f[2 * j]);
            float h0, h1;
            asm("{\n.reg .b16 lo, hi;\nmov.b32 {lo, hi}, %2;\ncvt.f32.f16 %0, lo;\ncvt.f32.f16 %1, hi;\n}" : "=f"(h0), "=f"(h1) : "r"(pw));
            t1[e0 + 4 * i + 2 * j] = (unsigned short)(pw & 0xFFFFu);
            t1[e0 + 4 * i + 2 * j + 1] = (unsigned short)(pw >> 16);
            r[2 * j] = (f[2 * j] - h0) * 64.0f;
            r[2 * j + 1] = (f[2 * j + 1] - h1) * 64.0f;
        }
        constexpr float k = 1.0f / 64.0f;
        tq[i] = pack_e4m3x4(f[0] * k, f[1] * k, f[2] * k, f[3] * k);
        rq[i] = pack_e4m3x4(r[0], r[1], r[2], r[3]);
    }
    uint4 *out = reinterpret_cast<uint4 *>(t2 + g * 32);
    out[0] = make_uint4(tq[0], tq[1], tq[2], tq[3]);
    out[1] = make_uint4(rq[0], rq[1], rq[2], rq[3]);
}

struct ConvScale {                      // what the converter needs of BatchArgs, by value (a reference would force a stack copy of the parameters)
    float sx0, sx1, sx2, sx3;
    int kbps, n_streams;
};

// Converter role: fp32 clip tile -> x1, x2 bf16 tiles, for a group of 32 * ITEMS ... threads.
// Work item = (row r, 8 consecutive dims c8): two 16 B chunks of the fp32 row -> one 16 B chunk of x1 and of x2.
// 8 consecutive threads take rows (2p, 2p+1) x c8 = 0..3, which makes every quarter-warp phase of the 128-bit loads
// and stores hit 8 distinct 16 B bank groups under both swizzles.  The loads of K block i+1 are issued before block i
// is converted (software pipeline: shared-memory latency off the critical path).  ITEMS = items per thread and K
// block: 4 for the four dedicated converter warps (t = 0..127), 2 when the four epilogue warps of an unused query
// half join them (t = 0..255; batches of <= 128 queries are conversion-bound, not MMA-bound).
template <int ITEMS, int QT>
__device__ __forceinline__ void convert_blocks(const int t, const int lane, const int n_it, const uint32_t smem_base,
                                               const uint32_t bar_afull, const uint32_t bar_aempty, const uint32_t bar_xfull,
                                               const uint32_t bar_xempty, long long &c_wait, long long &c_wait2,
                                               const ConvScale cs) {
    constexpr int ROWS_PER_PASS = BM / ITEMS;                        // 32 or 64
    constexpr int NA = Ring<QT>::NA, NXR = Ring<QT>::NXR;
    constexpr uint32_t RING_X = Ring<QT>::RING_X;
    const int rsub = (t >> 3) * 2 + ((t & 7) >> 2);                  // 0 .. ROWS_PER_PASS-1
    const int c8 = t & 3;
    const uint32_t src_off0 = (uint32_t)rsub * 128u + (uint32_t)(((2 * c8) ^ (rsub & 7)) * 16);
    const uint32_t src_off1 = (uint32_t)rsub * 128u + (uint32_t)(((2 * c8 + 1) ^ (rsub & 7)) * 16);
    const uint32_t dst_off = (uint32_t)rsub * 64u + (uint32_t)((c8 ^ ((rsub >> 1) & 3)) * 16);
#ifdef VQ_BATCH_F16F8
    // residual tile: per 16 dims [rx x 16 | x x 16] = chunks (2g, 2g + 1) of the 64 B row, g = c8 >> 1; this item's 8 dims
    // fill bytes 8 (c8 & 1) .. +8 of both chunks
    const uint32_t swz = (uint32_t)((rsub >> 1) & 3);
    const uint32_t dst_r = (uint32_t)rsub * 64u + ((((uint32_t)(c8 >> 1) * 2u) ^ swz) * 16u) + (uint32_t)(c8 & 1) * 8u;
    const uint32_t dst_x = (uint32_t)rsub * 64u + ((((uint32_t)(c8 >> 1) * 2u + 1u) ^ swz) * 16u) + (uint32_t)(c8 & 1) * 8u;
    const int kb_total_c = cs.kbps * cs.n_streams;
#endif
    float4 u[ITEMS], v[ITEMS];
    auto load_block = [&](int it) {
        const int sa = it % NA;
        const long long t0 = VQ_CLOCK();
        mbar_wait(bar_afull + 8 * sa, (it / NA) & 1);
        c_wait += VQ_CLOCK() - t0;
        const uint32_t src = smem_base + (uint32_t)sa * A32_BYTES;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            u[j] = lds128(src + src_off0 + (uint32_t)(j * ROWS_PER_PASS * 128));
            v[j] = lds128(src + src_off1 + (uint32_t)(j * ROWS_PER_PASS * 128));
        }
    };
    if (n_it > 0) load_block(0);
    for (int it = 0; it < n_it; ++it) {
        const int sa = it % NA, sx = it % NXR;
#ifndef VQ_BATCH_F16F8
        uint4 p[ITEMS], q[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) split8(u[j], v[j], p[j], q[j]);
#else
        uint4 p[ITEMS];
        uint2 rq[ITEMS], xq[ITEMS];
        const int stream_i = (it % kb_total_c) / cs.kbps;
        const float scale = stream_i == 0 ? cs.sx0 : (stream_i == 1 ? cs.sx1 : (stream_i == 2 ? cs.sx2 : cs.sx3));
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) split8_f16f8(u[j], v[j], scale, p[j], rq[j], xq[j]);
#endif
        // the fp32 stage is free as soon as its values sit in registers (the loads above have returned: split8 used them)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_aempty + 8 * sa);
        if (it + 1 < n_it) load_block(it + 1);                       // next block's loads fly while this one is stored
        const long long t1 = VQ_CLOCK();
        mbar_wait(bar_xempty + 8 * sx, ((it / NXR) & 1) ^ 1);
        c_wait2 += VQ_CLOCK() - t1;
        const uint32_t dst = smem_base + RING_X + (uint32_t)sx * 2 * X_BYTES + dst_off;
#ifndef VQ_BATCH_F16F8
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            sts128(dst + (uint32_t)(j * ROWS_PER_PASS * 64), p[j]);
            sts128(dst + X_BYTES + (uint32_t)(j * ROWS_PER_PASS * 64), q[j]);
        }
#else
        const uint32_t res = smem_base + RING_X + (uint32_t)sx * 2 * X_BYTES + X_BYTES;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {        // ROWS_PER_PASS is a multiple of 8: the swizzle term of a row is the same in every pass
            sts128(dst + (uint32_t)(j * ROWS_PER_PASS * 64), p[j]);
            sts64(res + dst_r + (uint32_t)(j * ROWS_PER_PASS * 64), rq[j]);
            sts64(res + dst_x + (uint32_t)(j * ROWS_PER_PASS * 64), xq[j]);
        }
#endif
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (UMMA)
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_xfull + 8 * sx);
    }
}

// kTies: also report, per query, the rows within COMPUTE_EPS of the threshold or of the near-miss limit (the tie band of
// the single-query scan, vq_scan.cu classify()): four more compares per score in the scoring phase, so it is a separate
// instantiation that only runs when the caller passes eps > 0.
template <int QT, bool kTies>
__global__ void __launch_bounds__(THREADS, 1)
batch_scan_bf16(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_t1,
                const __grid_constant__ CUtensorMap map_t2, const BatchArgs a, const float *__restrict__ inv_counts,
                const float *__restrict__ cut_g, unsigned long long *counts_g /*[QN][2]*/, unsigned int *cand_cnt /*[QN]*/,
                unsigned long long *cand_keys /*[QN][cap]*/, unsigned int *tie_cnt /*[QN]*/, unsigned long long *tie_keys /*[QN][tie_cap]*/,
                float *park_g /*[grid][QN][BM], L2 park only*/, float *scores_dbg /*[Q][n_rows] or null*/, long long *prof /*[grid][8] or null*/) {
    using R = Ring<QT>;
    constexpr int NA = R::NA, NXR = R::NXR, NT = R::NT, N_BARS = R::N_BARS;
    constexpr uint32_t T_BYTES = R::T_BYTES, RING_X = R::RING_X, RING_T = R::RING_T, RING_END = R::RING_END;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + RING_END);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + N_BARS);
    float *cut_s = reinterpret_cast<float *>(smem + RING_END + 256);
    unsigned int *cnt_s = reinterpret_cast<unsigned int *>(cut_s + QN);      // [8 warps][128 queries][2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_afull = smem_u32(&bars[0]), bar_aempty = smem_u32(&bars[NA]), bar_xfull = smem_u32(&bars[2 * NA]),
                   bar_xempty = smem_u32(&bars[2 * NA + NXR]), bar_tfull = smem_u32(&bars[2 * NA + 2 * NXR]),
                   bar_tempty = smem_u32(&bars[2 * NA + 2 * NXR + NT]), bar_part_full = smem_u32(&bars[2 * NA + 2 * NXR + 2 * NT]),
                   bar_part_empty = smem_u32(&bars[2 * NA + 2 * NXR + 2 * NT + 2]);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; ++s) {
            mbar_init(bar_afull + 8 * s, 1);            // producer's expect_tx arrive + TMA bytes of the fp32 tile
            mbar_init(bar_aempty + 8 * s, a.n_mma > 128 ? CONV_WARPS : 2 * CONV_WARPS);  // one arrive per converting warp
        }
        for (int s = 0; s < NXR; ++s) {
            mbar_init(bar_xfull + 8 * s, a.n_mma > 128 ? CONV_WARPS : 2 * CONV_WARPS);   // x1, x2 written
            mbar_init(bar_xempty + 8 * s, 1);           // tcgen05.commit
        }
        for (int s = 0; s < NT; ++s) {
            mbar_init(bar_tfull + 8 * s, 1);            // expect_tx arrive + TMA bytes of t1, t2
            mbar_init(bar_tempty + 8 * s, 1);           // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_part_full + 8 * b, 1);           // tcgen05.commit
            mbar_init(bar_part_empty + 8 * b, a.n_mma > 128 ? EPI_WARPS : EPI_WARPS / 2);  // one arrive per draining epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < QN; i += blockDim.x) cut_s[i] = cut_g[i];
    for (int i = threadIdx.x; i < EPI_WARPS * 128 * 2; i += blockDim.x) cnt_s[i] = 0;
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
#ifdef VQ_BATCH_F16F8
    const ConvScale cs = {a.sx[0], a.sx[1], a.sx[2], a.sx[3], a.kb_per_stream, a.n_streams};
#else
    const ConvScale cs = {1.f, 1.f, 1.f, 1.f, 0, 0};
#endif
    const int kbps = a.kb_per_stream;
    const int kb_total = kbps * a.n_streams;
    const int n_mma = a.n_mma;                      // queries rounded up to 16: the N of every MMA
    constexpr int NPART = 2;                        // partial accumulators: TMEM columns 0-255 and 256-511
    // batches of <= 128 queries leave the epilogue warps of the upper query half without work: they convert instead
    const int conv_warps = n_mma > 128 ? CONV_WARPS : 2 * CONV_WARPS;
    const uint32_t t_rows = n_mma > 128 ? 256u : (n_mma > 64 ? 128u : 64u);   // query rows fetched per t1 / t2 tile (host: same rule)
    // this CTA's K blocks in issue order: tiles blockIdx.x, +gridDim.x, ...; kb_total blocks each
    const int my_tiles = (a.n_tiles > (int)blockIdx.x) ? (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int n_it = my_tiles * kb_total;

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0) {
            // ------------------------------------------------------------------ TMA producer, clip tiles (fp32)
            if (lane == 0) {
                int it = 0;
                const uint64_t pol = policy_evict_first();
                // (an L2 prefetch of the clip tiles 8-16 K blocks ahead was measured 4-6 % SLOWER: the kernel is bound by the
                // L2 -> SM fill rate, ~6.3 KB/clk chip-wide, and prefetches only add L2 traffic)
                for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                    const int row = (int)(a.row0 + (long long)tile * BM);
                    for (int kb = 0; kb < kb_total; ++kb, ++it) {
                        const int s = it % NA;
                        mbar_wait(bar_aempty + 8 * s, ((it / NA) & 1) ^ 1);
                        mbar_expect(bar_afull + 8 * s, A32_BYTES);
                        tma_load_2d_hint(smem_base + (uint32_t)s * A32_BYTES, &map_a, kb * BK, row, bar_afull + 8 * s, pol);
                    }
                }
            }
        } else if (warp == 2) {
            // ------------------------------------------------------------------ TMA producer, query tiles (t1, t2)
            if (lane == 0) {
                const uint64_t pol = policy_evict_last();
                for (int it = 0; it < n_it; ++it) {
                    const int s = it % NT;
                    mbar_wait(bar_tempty + 8 * s, ((it / NT) & 1) ^ 1);
                    const uint32_t base = smem_base + RING_T + (uint32_t)s * 2 * T_BYTES;
                    const int kb = it % kb_total;
                    mbar_expect(bar_tfull + 8 * s, 2 * t_rows * BK * 2);      // the tensor maps' box holds t_rows query rows
                    tma_load_2d_hint(base, &map_t1, kb * BK, 0, bar_tfull + 8 * s, pol);
                    tma_load_2d_hint(base + T_BYTES, &map_t2, kb * BK, 0, bar_tfull + 8 * s, pol);
                }
            }
        } else if (warp == 1) {
            // ------------------------------------------------------------------ MMA issuer
            uint32_t elected;
            asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(elected));
#ifndef VQ_BATCH_F16F8
            // instruction descriptor: D = f32, A = B = bf16, both K-major, N = n_mma, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
#else
            // instruction descriptor: D = f32, A / B format code 0 (= F16 for kind::f16, = E4M3 for kind::f8f6f4), both K-major
            const uint32_t idesc = (1u << 4) | ((uint32_t)(n_mma >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
#endif
            int it = 0, gcount = 0;
            long long w_acc = 0, w_data = 0, w_conv = 0;
            const long long m_t0 = VQ_CLOCK();
            unsigned long long g_t0 = 0;
            if (prof) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_t0));
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                for (int st = 0; st < a.n_streams; ++st) {
                    uint32_t d = 0;
                    for (int kb = 0; kb < kbps; ++kb, ++it) {
                        const bool group_first = (kb % GROUP_KB) == 0;
                        const bool group_last = (kb % GROUP_KB) == GROUP_KB - 1 || kb == kbps - 1;
                        long long t0 = VQ_CLOCK();
                        if (group_first) {
                            const int b = gcount % NPART;
                            mbar_wait(bar_part_empty + 8 * b, ((gcount / NPART) & 1) ^ 1);    // partial drained
                            d = tmem_base + (uint32_t)(b * QN);
                        }
                        long long t1 = VQ_CLOCK();
                        w_acc += t1 - t0;
                        const int sx = it % NXR, stg = it % NT;
                        mbar_wait(bar_tfull + 8 * stg, (it / NT) & 1);
                        t0 = VQ_CLOCK();
                        w_data += t0 - t1;
                        mbar_wait(bar_xfull + 8 * sx, (it / NXR) & 1);
                        w_conv += VQ_CLOCK() - t0;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t xbase = smem_base + RING_X + (uint32_t)sx * 2 * X_BYTES;
                        const uint32_t tbase = smem_base + RING_T + (uint32_t)stg * 2 * T_BYTES;
                        if (elected) {
                            const uint32_t x1 = desc_lo(xbase), x2 = desc_lo(xbase + X_BYTES), t1d = desc_lo(tbase), t2d = desc_lo(tbase + T_BYTES);
                            if (group_first) umma_bf16<false>(d, x1, t1d, idesc);
                            else umma_bf16<true>(d, x1, t1d, idesc);
#ifndef VQ_BATCH_F16F8
                            umma_bf16<true>(d, x2, t1d, idesc);
                            umma_bf16<true>(d, x1, t2d, idesc);
#pragma unroll
                            for (int k = 1; k < BK / UK; ++k) {                  // +32 B per K step inside the 64 B row
                                umma_bf16<true>(d, x1 + 2 * k, t1d + 2 * k, idesc);
                                umma_bf16<true>(d, x2 + 2 * k, t1d + 2 * k, idesc);
                                umma_bf16<true>(d, x1 + 2 * k, t2d + 2 * k, idesc);
                            }
#else
                            umma_f8<true>(d, x2, t2d, idesc);                    // [rx | x] . [t | rt] over 32 e4m3 = both residual terms
#pragma unroll
                            for (int k = 1; k < BK / UK; ++k) {                  // +32 B per K step inside the 64 B row (both kinds)
                                umma_bf16<true>(d, x1 + 2 * k, t1d + 2 * k, idesc);
                                umma_f8<true>(d, x2 + 2 * k, t2d + 2 * k, idesc);
                            }
#endif
                            umma_commit(bar_xempty + 8 * sx);                     // both operand stages are reusable once these MMAs retire
                            umma_commit(bar_tempty + 8 * stg);
                            if (group_last) umma_commit(bar_part_full + 8 * (gcount % NPART));
                        }
                        __syncwarp();
                        if (group_last) ++gcount;
                    }
                }
            }
            if (prof && elected) {
                prof[blockIdx.x * 16 + 0] = VQ_CLOCK() - m_t0; prof[blockIdx.x * 16 + 1] = w_conv;
                prof[blockIdx.x * 16 + 2] = w_acc; prof[blockIdx.x * 16 + 3] = w_data;
                unsigned long long g_t1;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_t1));
                prof[blockIdx.x * 16 + 12] = (long long)(g_t1 - g_t0);
            }
        }
    } else if (warp < 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        // ---------------------------------------------------------------------- converter (warps 4-7, one per SM sub-partition)
        long long c_wait = 0, c_wait2 = 0;
        const int t = threadIdx.x - 128;                             // 0..127
        if (conv_warps == 2 * CONV_WARPS) convert_blocks<2, QT>(t, lane, n_it, smem_base, bar_afull, bar_aempty, bar_xfull, bar_xempty, c_wait, c_wait2, cs);
        else convert_blocks<4, QT>(t, lane, n_it, smem_base, bar_afull, bar_aempty, bar_xfull, bar_xempty, c_wait, c_wait2, cs);
        if (prof && t == 0) { prof[blockIdx.x * 16 + 7] = c_wait; prof[blockIdx.x * 16 + 4] = c_wait2; }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 192;");
        // ------------------------------------------------------------------ epilogue (8 warps)
        if (warp >= 12 && conv_warps == 2 * CONV_WARPS) {
            // no queries in the upper half: these four warps are converter threads 128..255
            long long cw = 0, cw2 = 0;
            convert_blocks<2, QT>(threadIdx.x - 384 + 128, lane, n_it, smem_base, bar_afull, bar_aempty, bar_xfull, bar_xempty, cw, cw2, cs);
        }
        const int ew = warp - 8;                  // 0..7
        const int quarter = warp & 3;             // TMEM lanes 32*quarter .. +31 (hardware rule: warp id % 4)
        const int half = ew >> 2;                 // which 128 of the 256 queries this warp handles
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 128);
        // 32-query chunks of this warp's half that hold live MMA columns (warp-uniform)
        const int nch = max(0, min(4, (n_mma - half * 128 + 31) / 32));
        unsigned int *my_cnt = cnt_s + ew * 128 * 2;
        float4 *park = reinterpret_cast<float4 *>(park_g) + (size_t)blockIdx.x * (QN / 4) * BM + (size_t)(half * 32) * BM + quarter * 32 + lane;
        const uint64_t pol_park = policy_evict_last();
        int gcount = 0;
        long long e_wait = 0, e_busy = 0, e_score = 0, e_fin0 = 0, e_fin1 = 0, e_tiles = 0;
        const bool drains = !(warp >= 12 && conv_warps == 2 * CONV_WARPS);     // the converting warps take no part in the epilogue
        for (int tile = blockIdx.x; tile < a.n_tiles && drains; tile += gridDim.x) {
            const long long row = a.row0 + (long long)tile * BM + quarter * 32 + lane;
            const bool row_ok = row < a.row_end;
            for (int st = 0; st < a.n_streams; ++st) {
                float run[128];
#pragma unroll
                for (int j = 0; j < 128; ++j) run[j] = 0.f;
                const int n_groups = (kbps + GROUP_KB - 1) / GROUP_KB;
                for (int g = 0; g < n_groups; ++g, ++gcount) {
                    const long long t0 = VQ_CLOCK();
                    const int b = gcount % NPART;
                    mbar_wait(bar_part_full + 8 * b, (gcount / NPART) & 1);
                    const long long t1 = VQ_CLOCK();
                    e_wait += t1 - t0;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t col = tlane + (uint32_t)(b * QN);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t r0[32];
                        if (c < nch) {
                            tmem_ld32(col + 32 * c, r0);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        }
                        if (c == 3) {                                    // partial fully read: the MMA warp may refill it
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar_part_empty + 8 * b);
                        }
                        if (c < nch) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) run[32 * c + j] += __uint_as_float(r0[j]);
                        }
                    }
                    e_busy += VQ_CLOCK() - t1;
                }
                // this stream's contribution (w (1 - sim))^2; earlier streams' terms come back from the park
                {
                    const long long t1 = VQ_CLOCK();
#ifndef VQ_BATCH_F16F8
                    const float ic = (inv_counts && row_ok) ? inv_counts[row * a.n_streams + st] : a.inv_splits;
#else
                    const float ds = st == 0 ? a.descale[0] : (st == 1 ? a.descale[1] : (st == 2 ? a.descale[2] : a.descale[3]));
                    const float ic = ((inv_counts && row_ok) ? inv_counts[row * a.n_streams + st] : a.inv_splits) * ds;   // ds: a power of two
#endif
                    const float w = st == 0 ? a.w[0] : (st == 1 ? a.w[1] : (st == 2 ? a.w[2] : a.w[3]));   // no dynamic indexing: keeps `a` in the constant bank
                    const bool first = st == 0, last = st + 1 == a.n_streams;
                    if (first) {
                        if (last) stream_final<true, true>(run, park, pol_park, w, ic, nch);
                        else stream_final<true, false>(run, park, pol_park, w, ic, nch);
                    } else {
                        if (last) stream_final<false, true>(run, park, pol_park, w, ic, nch);
                        else stream_final<false, false>(run, park, pol_park, w, ic, nch);
                    }
                    if (last) e_fin1 += VQ_CLOCK() - t1;
                    else e_fin0 += VQ_CLOCK() - t1;
                }
                if (st + 1 < a.n_streams) continue;
                ++e_tiles;
                // ---- scores of this thread's clip against this warp's 128 queries.
                // Phase 1 (branch-free): all scores; per 32-query chunk three bit masks of this row: score >= threshold,
                // >= near-miss limit, > current top-k cut.  Phase 2: per-query counts by transposing the masks across
                // the warp (lane j ends up with the 32 rows' bits of query j).  Phase 3: candidate appends, only for
                // the (rare, after the first chunk) queries some row of this warp beats the cut of.
                const long long t1 = VQ_CLOCK();
                unsigned int m_th[4], m_nm[4], m_cd[4], m_tz[4];
                const unsigned int rowmask = row_ok ? 0xffffffffu : 0u;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    m_th[c] = m_nm[c] = m_cd[c] = m_tz[c] = 0u;
                    if (c < nch) {
                        unsigned int bt = 0, bl = 0, bc = 0, bz = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const int ql = 32 * c + j;
                            const float sc = 1.0f - sqrt_approx(run[ql] * a.inv_den);
                            run[ql] = sc;
                            bt |= (sc >= a.th_f ? 1u : 0u) << j;
                            bl |= (sc >= a.lo_f ? 1u : 0u) << j;
                            bc |= (sc > cut_s[half * 128 + ql] ? 1u : 0u) << j;
                            if constexpr (kTies)
                                bz |= (((sc >= a.tie_lo0) & (sc <= a.tie_hi0)) | ((sc >= a.tie_lo1) & (sc <= a.tie_hi1)) ? 1u : 0u) << j;
                        }
                        const int nlive = a.n_queries - (half * 128 + 32 * c);          // live queries of this chunk
                        const unsigned int lm = rowmask & (nlive >= 32 ? 0xffffffffu : (nlive <= 0 ? 0u : ((1u << nlive) - 1u)));
                        m_th[c] = bt & lm;
                        m_nm[c] = bl & ~bt & lm;
                        m_cd[c] = bc & lm;
                        m_tz[c] = bz & lm;
                    }
                }
                if (scores_dbg && row_ok) {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (c < nch) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int q = half * 128 + 32 * c + j;
                                if (q < a.n_queries) scores_dbg[(size_t)q * a.n_rows_total + row] = run[32 * c + j];
                            }
                        }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < nch) {                                   // lane owns query c*32+lane of this warp's half
                        my_cnt[(c * 32 + lane) * 2] += __popc(transpose32(m_th[c], lane));
                        my_cnt[(c * 32 + lane) * 2 + 1] += __popc(transpose32(m_nm[c], lane));
                    }
                }
                // rare after the seeding chunk: a real loop over the hot queries (compact code, the score comes out
                // of the register file through a select tree on the warp-uniform index); the tie band uses the same loop
                auto append_hot = [&](const unsigned int (&mask)[4], unsigned int *cnt, unsigned long long *keys, const long long cap) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        unsigned int h = __reduce_or_sync(0xffffffffu, mask[c]);
                        while (h) {
                            const int j = __ffs(h) - 1;
                            h &= h - 1;
                            const float sc = select32(run, 32 * c, j);
                            const int q = half * 128 + c * 32 + j;
                            const bool cand = (mask[c] >> j) & 1u;
                            const unsigned int bc = __ballot_sync(0xffffffffu, cand);
                            const int leader_lane = __ffs(bc) - 1;
                            unsigned int base = 0;
                            if (lane == leader_lane) base = atomicAdd(&cnt[q], (unsigned int)__popc(bc));
                            base = __shfl_sync(0xffffffffu, base, leader_lane);
                            if (cand) {
                                const long long slot = (long long)base + __popc(bc & ((1u << lane) - 1u));
                                if (slot < cap) keys[(size_t)q * cap + slot] = vq::make_key(sc, (unsigned int)row);
                            }
                        }
                    }
                };
                append_hot(m_cd, cand_cnt, cand_keys, a.cand_cap);
                if constexpr (kTies) append_hot(m_tz, tie_cnt, tie_keys, a.tie_cap);
                e_score += VQ_CLOCK() - t1;
            }
        }
        __syncwarp();
        for (int ql = lane; ql < 128; ql += 32) {
            const int q = half * 128 + ql;
            if (my_cnt[ql * 2]) atomicAdd(&counts_g[2 * q], (unsigned long long)my_cnt[ql * 2]);
            if (my_cnt[ql * 2 + 1]) atomicAdd(&counts_g[2 * q + 1], (unsigned long long)my_cnt[ql * 2 + 1]);
        }
        if (prof && threadIdx.x == 256) {
            prof[blockIdx.x * 16 + 5] = e_wait; prof[blockIdx.x * 16 + 6] = e_busy; prof[blockIdx.x * 16 + 8] = e_fin0;
            prof[blockIdx.x * 16 + 9] = e_fin1; prof[blockIdx.x * 16 + 10] = e_score; prof[blockIdx.x * 16 + 11] = e_tiles;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

}  // namespace bf
