// K3 — batched queries on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Q targets are scored against every clip of the shard in ONE pass over HBM per 128 queries: per
// stream the similarities are a dense contraction SIM_s[clip, query] = sum_d X[clip, s, d] * T[query, s, d]
// (reference ticket.py:146-160 for Q tickets at once), then the per-(clip, query) score of
// ticket.py:173-180 and the candidate tests of ticket.py:325-327 are applied in the epilogue.
// No Q x N score matrix is ever written: per query the kernel keeps the match / near-miss counts
// and a candidate list for the exact top-k (ranking rule of ticket.py:266).
//
// Precision (measured on B200, see profiles/): two effects rule out a plain TF32 GEMM for the 1e-5 bar.
//  (1) TF32 operands carry 10 mantissa bits -> each product is computed as three MMAs ("3xTF32"):
//          x*t ~= hi(x)*hi(t) + lo(x)*hi(t) + hi(x)*lo(t)
//      with lo(x) = x - trunc_tf32(x) made on the fly by converter warps (shared -> shared) and hi(t), lo(t)
//      precomputed once per call (round-to-nearest split).
//  (2) the tensor core adds each MMA into its fp32 accumulator with truncation: 384 accumulations of
//      non-negative terms gave a -1.6e-5 relative bias.  So accumulation is two-level: hi*hi products go
//      to a partial accumulator that is drained every 4 K-blocks (16 MMAs) and summed in registers with
//      round-to-nearest fp32 adds; the small lo terms use their own accumulator (their truncation is
//      relative to a 2^-11 times smaller magnitude).
//
// CTA layout (384 threads, 1 CTA per SM, persistent over 128-clip tiles), streams processed one after
// the other:
//     warp 0       TMA producer: per K block of 32 floats, A tile [128 clips] + B_hi, B_lo tiles
//                  [128 queries], 128-byte swizzle, 3-stage ring of 64 KB
//     warp 1       MMA issuer: one elected thread, tcgen05.mma.kind::tf32 M128 N128 K8
//     warps 2-3    converter: lo(x) tiles for the A operand; warp 2 also allocates TMEM
//                  (512 columns: P_hi[0], P_hi[1], P_lo, parked stream term)
//     warps 4-11   epilogue: drain partials (tcgen05.ld), running sums in registers, score, tests,
//                  warp-ballot counts, top-k candidates
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "vq_internal.cuh"
#include "vq_topk.cuh"
#include "vq_tc.cuh"

namespace {

#ifdef VQ_BATCH_PROFILE          // role cycle counters (development builds only)
#define VQ_CLOCK() clock64()
#else
#define VQ_CLOCK() 0ll
#endif

#ifndef VQ_GROUP_KB
#define VQ_GROUP_KB 4
#endif

constexpr int BM = 128;                  // clips per tile (UMMA M)
constexpr int QT = 128;                  // queries per pass (UMMA N)
constexpr int BK = 32;                   // floats per K block = one 128-byte swizzle row
constexpr int UK = 8;                    // UMMA K for tf32
constexpr int GROUP_KB = VQ_GROUP_KB;    // K blocks per partial accumulator (4 hi*hi MMAs each)
constexpr uint32_t A_BYTES = BM * BK * 4;        // 16 KB
constexpr int BATCH_THREADS = 384;      // 12 warps -> up to 168 registers per thread (the epilogue keeps 64 sums)
constexpr int CONV_THREADS = 64;        // converter = warps 2-3
constexpr int PF_KB = 8;                         // K blocks per L2 prefetch box (8 x 128 B = 1 KB per clip row)

// kPair = false: one CTA per tile of 128 clips (cta_group::1).
// kPair = true : a CTA pair (cluster of 2, cta_group::2) per tile of 256 clips: each CTA stages its own 128
//                clips and HALF of the query tile; one thread of the leader CTA issues M256 MMAs for both.
//                Halves the B bytes each SM has to fill and re-read per MMA (the kernel is shared-memory
//                bandwidth bound, profiles/r1_k3_batched_notes.md).
template <bool kPair>
struct Cfg {
    static constexpr int kStages = kPair ? 4 : 3;
    static constexpr uint32_t kBBytes = (kPair ? QT / 2 : QT) * BK * 4;          // per CTA: 8 KB or 16 KB
    static constexpr uint32_t kStageBytes = 2 * A_BYTES + 2 * kBBytes;           // A, A_lo, B_hi, B_lo
    static constexpr int kBars = 4 * kStages + 6;                                // a_full, full, conv, empty + 6
    static constexpr size_t kSmem = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ +
                                    2 * QT * 4 /*cut, gate*/ + 8 * 64 * 2 * 4 /*per-warp counts*/;
};
constexpr uint32_t COL_PHI0 = 0, COL_PHI1 = 128, COL_PLO = 256, COL_PARK = 384;

using namespace vqtc;

struct BatchArgs {
    float w[VQ_MAX_STREAMS];
    float inv_den, inv_splits;
    float th_f, lo_f;              // smallest floats >= the double thresholds: (double)s >= th  <=>  s >= th_f
    int n_queries;
    int kb_per_stream;             // stream_len / 32
    int n_streams;
    long long row0, n_rows_total;  // chunk start (local rows) and shard size
    long long row_end;             // bf16 kernel: first row past this launch's chunk
    int n_tiles;                   // tiles in this chunk
    int n_mma;                     // bf16 kernel: queries of this pass rounded up to 16 (the UMMA N)
    long long cand_cap;
};

#include "vq_batch_bf16.cuh"

// hi/lo split of the targets, round to nearest (cvt.rna): hi has a 10-bit mantissa, lo = t - hi exactly
__global__ void split_targets(const float *__restrict__ t, float *hi, float *lo, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = t[i];
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    const float hf = __uint_as_float(h);
    hi[i] = hf;
    lo[i] = x - hf;
}

template <bool kPair>
__global__ void __launch_bounds__(BATCH_THREADS, 1)
batch_scan(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bhi,
           const __grid_constant__ CUtensorMap map_blo, const __grid_constant__ CUtensorMap map_apf, const BatchArgs a, const float *__restrict__ inv_counts,
           const float *__restrict__ cut_g, unsigned long long *counts_g /*[QT][2]*/, unsigned int *cand_cnt /*[QT]*/,
           unsigned long long *cand_keys /*[QT][cap]*/, float *scores_dbg /*[Q][n_rows] or null*/,
           long long *prof /*[grid][8] cycle counters or null*/) {
    using C_ = Cfg<kPair>;
    constexpr int STAGES = C_::kStages;
    constexpr uint32_t B_BYTES = C_::kBBytes, STAGE_BYTES = C_::kStageBytes;
    extern __shared__ unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * STAGE_BYTES);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + C_::kBars);
    float *cut_s = reinterpret_cast<float *>(smem + (size_t)STAGES * STAGE_BYTES + 256);
    float *gate_s = cut_s + QT;
    unsigned int *cnt_s = reinterpret_cast<unsigned int *>(gate_s + QT);      // [8 warps][64 queries][2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    // Barriers (identical offsets in both CTAs of a pair).  Waited on locally: a_full, empty, part_full, lo_full.
    // Owned by the leader CTA (the peer arrives remotely): full, conv, part_empty, lo_empty.
    const uint32_t bar_afull = smem_u32(&bars[0]), bar_full = smem_u32(&bars[STAGES]), bar_conv = smem_u32(&bars[2 * STAGES]),
                   bar_empty = smem_u32(&bars[3 * STAGES]), bar_part_full = smem_u32(&bars[4 * STAGES]),
                   bar_part_empty = smem_u32(&bars[4 * STAGES + 2]), bar_lo_full = smem_u32(&bars[4 * STAGES + 4]),
                   bar_lo_empty = smem_u32(&bars[4 * STAGES + 5]);
    const int n_cta = kPair ? 2 : 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_afull + 8 * s, 1);          // pair mode: this CTA's A tile has landed
            mbar_init(bar_full + 8 * s, n_cta);       // producer arrive(s); bytes of all CTAs' TMA loads
            mbar_init(bar_conv + 8 * s, (CONV_THREADS / 32) * n_cta);   // one arrive per converter warp
            mbar_init(bar_empty + 8 * s, 1);          // tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_part_full + 8 * b, 1);          // tcgen05.commit
            mbar_init(bar_part_empty + 8 * b, 8 * n_cta); // one arrive per epilogue warp
        }
        mbar_init(bar_lo_full, 1);
        mbar_init(bar_lo_empty, 8 * n_cta);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < QT; i += blockDim.x) {
        const float c = cut_g[i];
        cut_s[i] = c;
        gate_s[i] = fminf(c, a.lo_f);
    }
    for (int i = threadIdx.x; i < 8 * 64 * 2; i += blockDim.x) cnt_s[i] = 0;
    if (warp == 2) {
        if constexpr (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (kPair) cluster_sync_all();          // peer barriers are initialised before any remote arrive
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const int kbps = a.kb_per_stream;
    const int kb_total = kbps * a.n_streams;
    // work distribution: a "unit" is a CTA (128 clips) or a CTA pair (256 clips)
    const int unit = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_units = kPair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int unit_rows = BM * n_cta;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int it = 0;
            long long p_wait = 0;
            const long long p_t0 = VQ_CLOCK();
            for (int tile = unit; tile < a.n_tiles; tile += n_units) {
                const int row = (int)(a.row0 + (long long)tile * unit_rows + (long long)cta_rank * BM);
                for (int kb = 0; kb < kb_total; ++kb, ++it) {
                    const int s = it % STAGES;
                    const uint32_t ph = (it / STAGES) & 1;
                    const long long t0 = VQ_CLOCK();
                    mbar_wait(bar_empty + 8 * s, ph ^ 1);
                    p_wait += VQ_CLOCK() - t0;
                    const uint32_t base = smem_u32(smem + (size_t)s * STAGE_BYTES);
                    if constexpr (kPair) {
                        // A: local barrier (the local converter waits on it).  B halves: the leader's barrier.
                        mbar_expect(bar_afull + 8 * s, A_BYTES);
                        tma_load_2d(base, &map_a, kb * BK, row, bar_afull + 8 * s);
                        const uint32_t lbar = (bar_full + 8 * s) & PEER_MASK;
                        if (leader) mbar_expect(bar_full + 8 * s, 4 * B_BYTES);      // 2 CTAs x (B_hi + B_lo halves)
                        else mbar_arrive_cluster(lbar);
                        tma_load_2d_pair(base + 2 * A_BYTES, &map_bhi, kb * BK, (int)cta_rank * (QT / 2), lbar);
                        tma_load_2d_pair(base + 2 * A_BYTES + B_BYTES, &map_blo, kb * BK, (int)cta_rank * (QT / 2), lbar);
                    } else {
                        mbar_expect(bar_full + 8 * s, A_BYTES + 2 * B_BYTES);
                        tma_load_2d(base, &map_a, kb * BK, row, bar_full + 8 * s);
                        tma_load_2d(base + 2 * A_BYTES, &map_bhi, kb * BK, 0, bar_full + 8 * s);
                        tma_load_2d(base + 2 * A_BYTES + B_BYTES, &map_blo, kb * BK, 0, bar_full + 8 * s);
                    }
                }
            }
            (void)p_wait; (void)p_t0;
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA only)
        // The whole warp runs the loop and the barrier waits, so that addresses and loop state are warp-uniform
        // (uniform registers feed UTCHMMA directly); one elected lane issues the MMAs and commits.
        if (leader) {
            uint32_t elected;
            asm volatile("{\n.reg .pred P;\nelect.sync _|P, 0xffffffff;\nselp.u32 %0, 1, 0, P;\n}" : "=r"(elected));
            // instruction descriptor: D = f32, A = B = tf32, both K-major, N = QT, M = 128 (256 for a CTA pair)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(QT >> 3) << 17) |
                                   ((uint32_t)((BM * n_cta) >> 4) << 24);
            auto commit = [&](uint32_t bar) {
                if constexpr (kPair) umma_commit_pair(bar);
                else umma_commit(bar);
            };
            auto wait = [&](uint32_t bar, uint32_t parity) {
                if constexpr (kPair) mbar_wait_cluster(bar, parity);
                else mbar_wait(bar, parity);
            };
            int it = 0, gcount = 0, lcount = 0;
            long long w_acc = 0, w_data = 0, w_lo = 0, w_conv = 0;
            const long long m_t0 = VQ_CLOCK();
            for (int tile = unit; tile < a.n_tiles; tile += n_units) {
                for (int st = 0; st < a.n_streams; ++st) {
                    uint32_t d_hi = 0;
                    for (int kb = 0; kb < kbps; ++kb, ++it) {
                        const bool group_first = (kb % GROUP_KB) == 0;
                        const bool group_last = (kb % GROUP_KB) == GROUP_KB - 1 || kb == kbps - 1;
                        long long t0 = VQ_CLOCK();
                        if (group_first) {
                            const int b = gcount & 1;
                            wait(bar_part_empty + 8 * b, ((gcount >> 1) & 1) ^ 1);        // partial drained
                            d_hi = tmem_base + (b ? COL_PHI1 : COL_PHI0);
                        }
                        long long t1 = VQ_CLOCK();
                        w_acc += t1 - t0;
                        if (kb == 0) wait(bar_lo_empty, (lcount & 1) ^ 1);
                        t0 = VQ_CLOCK();
                        w_lo += t0 - t1;
                        const int s = it % STAGES;
                        const uint32_t ph = (it / STAGES) & 1;
                        wait(bar_full + 8 * s, ph);
                        t1 = VQ_CLOCK();
                        w_data += t1 - t0;
                        wait(bar_conv + 8 * s, ph);
                        w_conv += VQ_CLOCK() - t1;
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t base = smem_u32(smem + (size_t)s * STAGE_BYTES);
                        const uint32_t d_lo = tmem_base + COL_PLO;
                        if (elected) {
                        const uint32_t xa = desc_lo(base), xl = desc_lo(base + A_BYTES), th = desc_lo(base + 2 * A_BYTES),
                                       tl = desc_lo(base + 2 * A_BYTES + B_BYTES);
                        // small terms first, into their own accumulator
                        if (kb == 0) umma_issue<kPair, false>(d_lo, xl, th, idesc);
                        else umma_issue<kPair, true>(d_lo, xl, th, idesc);
                        umma_issue<kPair, true>(d_lo, xa, tl, idesc);
#pragma unroll
                        for (int k = 1; k < BK / UK; ++k) {
                            umma_issue<kPair, true>(d_lo, xl + 2 * k, th + 2 * k, idesc);
                            umma_issue<kPair, true>(d_lo, xa + 2 * k, tl + 2 * k, idesc);
                        }
                        if (group_first) umma_issue<kPair, false>(d_hi, xa, th, idesc);
                        else umma_issue<kPair, true>(d_hi, xa, th, idesc);
#pragma unroll
                        for (int k = 1; k < BK / UK; ++k) umma_issue<kPair, true>(d_hi, xa + 2 * k, th + 2 * k, idesc);
                        commit(bar_empty + 8 * s);                       // stage reusable once these MMAs retire
                        if (group_last) commit(bar_part_full + 8 * (gcount & 1));
                        }
                        __syncwarp();
                        if (group_last) ++gcount;
                    }
                    if (elected) commit(bar_lo_full);
                    __syncwarp();
                    ++lcount;
                }
            }
            if (prof && elected) {
                prof[blockIdx.x * 8 + 2] = w_acc; prof[blockIdx.x * 8 + 3] = w_data;
                prof[blockIdx.x * 8 + 7] = w_lo; prof[blockIdx.x * 8 + 1] = w_conv; prof[blockIdx.x * 8 + 0] = VQ_CLOCK() - m_t0;
            }
        }
    } else if (warp == 2 || warp == 3) {
        // ------------------------------------------------------------------ converter: A_lo = x - trunc_tf32(x)
        const int t = threadIdx.x - 64;                              // 0..CONV_THREADS-1
        int it = 0;
        long long c_wait = 0;
        for (int tile = unit; tile < a.n_tiles; tile += n_units) {
            for (int kb = 0; kb < kb_total; ++kb, ++it) {
                const int s = it % STAGES;
                const uint32_t ph = (it / STAGES) & 1;
                const long long t0 = VQ_CLOCK();
                mbar_wait((kPair ? bar_afull : bar_full) + 8 * s, ph);
                c_wait += VQ_CLOCK() - t0;
                const float4 *src = reinterpret_cast<const float4 *>(smem + (size_t)s * STAGE_BYTES);
                float4 *dst = reinterpret_cast<float4 *>(smem + (size_t)s * STAGE_BYTES + A_BYTES);
#pragma unroll
                for (int j = 0; j < (int)(A_BYTES / 16 / CONV_THREADS); ++j) {
                    const float4 x = src[j * CONV_THREADS + t];
                    float4 l;
                    l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
                    l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
                    l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
                    l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
                    dst[j * CONV_THREADS + t] = l;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> async proxy (UMMA)
                __syncwarp();
                if (lane == 0) {
                    if constexpr (kPair) mbar_arrive_cluster((bar_conv + 8 * s) & PEER_MASK);
                    else mbar_arrive(bar_conv + 8 * s);
                }
            }
        }
        (void)c_wait;
    } else if (warp >= 4) {
        // ------------------------------------------------------------------ epilogue (8 warps)
        const int ew = warp - 4;                  // 0..7
        const int quarter = warp & 3;             // TMEM lanes 32*quarter .. +31 (hardware rule: warp id % 4)
        const int half = ew >> 2;                 // which 64 of the 128 queries this warp handles
        const uint32_t tlane = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 64);
        unsigned int *my_cnt = cnt_s + ew * 64 * 2;
        int gcount = 0, lcount = 0;
        long long e_wait = 0, e_busy = 0, e_score = 0;
        for (int tile = unit; tile < a.n_tiles; tile += n_units) {
            const long long row = a.row0 + (long long)tile * unit_rows + (long long)cta_rank * BM + quarter * 32 + lane;
            const bool row_ok = row < a.n_rows_total;
            for (int st = 0; st < a.n_streams; ++st) {
                float run[64];
#pragma unroll
                for (int j = 0; j < 64; ++j) run[j] = 0.f;
                const int n_groups = (kbps + GROUP_KB - 1) / GROUP_KB;
                for (int g = 0; g < n_groups; ++g, ++gcount) {
                    const int b = gcount & 1;
                    const long long t0 = VQ_CLOCK();
                    mbar_wait(bar_part_full + 8 * b, (gcount >> 1) & 1);
                    const long long t1 = VQ_CLOCK();
                    e_wait += t1 - t0;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    uint32_t r0[32];
                    const uint32_t col = b ? COL_PHI1 : COL_PHI0;
                    tmem_ld32(tlane + col, r0);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) run[j] += __uint_as_float(r0[j]);
                    tmem_ld32(tlane + col + 32, r0);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (kPair) mbar_arrive_cluster((bar_part_empty + 8 * b) & PEER_MASK);
                        else mbar_arrive(bar_part_empty + 8 * b);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) run[32 + j] += __uint_as_float(r0[j]);
                    e_busy += VQ_CLOCK() - t1;
                }
                // small terms of this stream, then this stream's contribution to the score
                {
                    const long long t0 = VQ_CLOCK();
                    mbar_wait(bar_lo_full, lcount & 1);
                    const long long t1 = VQ_CLOCK();
                    e_wait += t1 - t0;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const float ic = (inv_counts && row_ok) ? inv_counts[row * a.n_streams + st] : a.inv_splits;
                    const float w = a.w[st];
                    uint32_t r0[32];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {                    // two halves of 32 queries: keeps registers low
                        tmem_ld32(tlane + COL_PLO + 32 * h, r0);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (h == 1) {                                // P_lo fully read: the next stream may overwrite it
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            __syncwarp();
                            if (lane == 0) {
                                if constexpr (kPair) mbar_arrive_cluster(bar_lo_empty & PEER_MASK);
                                else mbar_arrive(bar_lo_empty);
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float d = w * (1.0f - (run[32 * h + j] + __uint_as_float(r0[j])) * ic);
                            run[32 * h + j] = d * d;
                        }
                        if (st > 0) {                                // add the terms of the earlier streams
                            tmem_ld32(tlane + COL_PARK + 32 * h, r0);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int j = 0; j < 32; ++j) run[32 * h + j] += __uint_as_float(r0[j]);
                        }
                        if (st + 1 < a.n_streams) {                  // park until the next stream is done
#pragma unroll
                            for (int j = 0; j < 32; ++j) r0[j] = __float_as_uint(run[32 * h + j]);
                            tmem_st32(tlane + COL_PARK + 32 * h, r0);
                            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                        }
                    }
                    ++lcount;
                    e_busy += VQ_CLOCK() - t1;
                }
                if (st + 1 < a.n_streams) continue;
                // ---- scores of this thread's clip against this warp's 64 queries.
                // Phase 1 (branch-free, pipelined): all 64 scores; bit ql of `hot` = this row passes query ql's gate
                // (gate = min(near-miss limit, current top-k cut)).  Phase 2: one warp-wide OR.  Phase 3: only the
                // queries some row of this warp is interesting for take the ballot / append path.
                const long long t1 = VQ_CLOCK();
                unsigned int hot_lo = 0, hot_hi = 0;
#pragma unroll
                for (int ql = 0; ql < 64; ++ql) {
                    const int q = half * 64 + ql;
                    const float sc = 1.0f - sqrt_approx(run[ql] * a.inv_den);
                    run[ql] = sc;
                    const bool live = row_ok && (q < a.n_queries);
                    if (scores_dbg && live) scores_dbg[(size_t)q * a.n_rows_total + row] = sc;
                    const unsigned int bit = (live && sc >= gate_s[q]) ? 1u : 0u;
                    if (ql < 32) hot_lo |= bit << ql;
                    else hot_hi |= bit << (ql - 32);
                }
                hot_lo = __reduce_or_sync(0xffffffffu, hot_lo);
                hot_hi = __reduce_or_sync(0xffffffffu, hot_hi);
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const unsigned int hot = c ? hot_hi : hot_lo;
                    if (hot == 0) continue;                          // warp-uniform
                    unsigned int cm = 0, cn = 0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (!(hot & (1u << j))) continue;            // warp-uniform
                        const int ql = c * 32 + j;
                        const int q = half * 64 + ql;
                        const float sc = run[ql];
                        const bool live = row_ok && (q < a.n_queries);
                        const bool m = live && (sc >= a.th_f);
                        const bool nm = live && !m && (sc >= a.lo_f);
                        const bool cand = live && (sc > cut_s[q]);
                        const unsigned int bm = __ballot_sync(0xffffffffu, m);
                        const unsigned int bn = __ballot_sync(0xffffffffu, nm);
                        const unsigned int bc = __ballot_sync(0xffffffffu, cand);
                        if (lane == j) { cm += __popc(bm); cn += __popc(bn); }
                        if (bc) {
                            const int leader_lane = __ffs(bc) - 1;
                            unsigned int base = 0;
                            if (lane == leader_lane) base = atomicAdd(&cand_cnt[q], (unsigned int)__popc(bc));
                            base = __shfl_sync(0xffffffffu, base, leader_lane);
                            if (cand) {
                                const long long slot = (long long)base + __popc(bc & ((1u << lane) - 1u));
                                if (slot < a.cand_cap)
                                    cand_keys[(size_t)q * a.cand_cap + slot] = vq::make_key(sc, (unsigned int)row);
                            }
                        }
                    }
                    my_cnt[(c * 32 + lane) * 2] += cm;               // lane owns query c*32+lane of this warp
                    my_cnt[(c * 32 + lane) * 2 + 1] += cn;
                }
                e_score += VQ_CLOCK() - t1;
            }
        }
        __syncwarp();
        for (int ql = lane; ql < 64; ql += 32) {
            const int q = half * 64 + ql;
            if (my_cnt[ql * 2]) atomicAdd(&counts_g[2 * q], (unsigned long long)my_cnt[ql * 2]);
            if (my_cnt[ql * 2 + 1]) atomicAdd(&counts_g[2 * q + 1], (unsigned long long)my_cnt[ql * 2 + 1]);
        }
        if (prof && threadIdx.x == 128) { prof[blockIdx.x * 8 + 5] = e_wait; prof[blockIdx.x * 8 + 6] = e_busy; prof[blockIdx.x * 8 + 4] = e_score; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if constexpr (kPair) cluster_sync_all();          // the peer may still be reading / being written
    if (warp == 2) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if constexpr (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

// After each chunk: keep the k best candidates per query (sorted), raise the cut to the k-th score.
__global__ void __launch_bounds__(1024)
batch_compact(unsigned int *cand_cnt, unsigned long long *cand_keys, long long cand_cap, int topk, float *cut_g) {
    __shared__ vq::TopkScratch tk;
    const int q = blockIdx.x;
    long long C = (long long)cand_cnt[q];
    if (C > cand_cap) C = cand_cap;
    unsigned long long *keys = cand_keys + (size_t)q * cand_cap;
    const int k = vq::block_topk_1024(keys, C, topk, tk);
    __syncthreads();
    for (int i = threadIdx.x; i < k; i += blockDim.x) keys[i] = tk.sel[i];
    if (threadIdx.x == 0) {
        cand_cnt[q] = (unsigned int)k;
        if (k == topk && k > 0) cut_g[q] = vq::key_score(tk.sel[k - 1]);
    }
}

__global__ void batch_output(const unsigned int *cand_cnt, const unsigned long long *cand_keys, long long cand_cap,
                             int topk, long long first_global_row, long long *rows_out, float *scores_out) {
    const int q = blockIdx.x;
    const int k = (int)cand_cnt[q];
    for (int i = threadIdx.x; i < topk; i += blockDim.x) {
        if (i < k) {
            const unsigned long long key = cand_keys[(size_t)q * cand_cap + i];
            rows_out[(size_t)q * topk + i] = first_global_row + (long long)vq::key_row(key);
            scores_out[(size_t)q * topk + i] = vq::key_score(key);
        } else {
            rows_out[(size_t)q * topk + i] = -1;
            scores_out[(size_t)q * topk + i] = __int_as_float(0xff800000);
        }
    }
}

__global__ void fill_f32(float *p, float v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int encode_map_ex(CUtensorMap *map, const void *base, CUtensorMapDataType dtype, size_t elem_bytes, uint64_t inner,
                  uint64_t rows, uint32_t box_rows, uint32_t box_inner, CUtensorMapSwizzle swizzle) {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        VQ_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr));
        VQ_REQUIRE(p && qr == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
        fn = (EncodeTiledFn)p;
    }
    const cuuint64_t dims[2] = {inner, rows};
    const cuuint64_t strides[1] = {inner * elem_bytes};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, dtype, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VQ_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return 0;
}

int encode_map(CUtensorMap *map, const float *base, uint64_t inner, uint64_t rows, uint32_t box_rows,
               uint32_t box_inner = BK, bool swizzle = true) {
    return encode_map_ex(map, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, sizeof(float), inner, rows, box_rows, box_inner,
                         swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
}

float float_ceil_of(double x) {
    // smallest float f with (double)f >= x, so that for every float s: (double)s >= x  <=>  s >= f
    float f = (float)x;
    if ((double)f < x) f = nextafterf(f, INFINITY);
    return f;
}

struct Dev {
    void *p = nullptr;
    ~Dev() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t b) { return cudaMalloc(&p, b ? b : 8); }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

int run_batch(vq_store *s, const float *targets, int n_queries, const vq_scan_params *p, int64_t *counts_out,
              int64_t *topk_rows_out, float *topk_scores_out, float *kernel_ms_out, float *scores_dbg_host) {
    VQ_REQUIRE(s && targets && p, "vq_scan_batch: null argument");
    VQ_REQUIRE(n_queries >= 1, "vq_scan_batch: need at least one query");
    VQ_REQUIRE(s->stream_len % BK == 0, "vq_scan_batch: stream length %d is not a multiple of %d", s->stream_len, BK);
    VQ_REQUIRE(p->topk >= 0 && p->topk <= VQ_MAX_TOPK, "vq_scan_batch: topk %d outside 0..%d", p->topk, VQ_MAX_TOPK);
    VQ_REQUIRE(s->n_rows < (1ll << 31), "vq_scan_batch: shard too large for 32-bit TMA coordinates");
    double den = 0.0;
    for (int i = 0; i < s->n_streams; ++i) den += p->weights[i] * p->weights[i];
    VQ_REQUIRE(den > 0.0, "vq_scan_batch: all stream weights are zero");
    VQ_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    const size_t K = s->row_floats;                       // floats per row = S * stream_len
    const int topk = p->topk;
    const long long chunk_rows = (long long)s->sm_count * BM * 6;
    const long long cap = chunk_rows + VQ_MAX_TOPK;
    // One CTA per 128-clip tile by default.  VQ_BATCH_PAIR=1 selects the CTA-pair variant (cta_group::2, M = 256):
    // correct and parity-tested, but measured 25-35 % slower on B200 (profiles/r1_k3_batched_notes.md).
    const bool pair = getenv("VQ_BATCH_PAIR") && atoi(getenv("VQ_BATCH_PAIR")) == 1 && s->sm_count >= 2;
    VQ_CUDA(cudaFuncSetAttribute(batch_scan<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<false>::kSmem));
    VQ_CUDA(cudaFuncSetAttribute(batch_scan<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg<true>::kSmem));
    Dev d_t, d_hi, d_lo, d_cut, d_counts, d_cnt, d_keys, d_rows, d_sc, d_dbg;
    VQ_CUDA(d_t.alloc((size_t)QT * K * 4));
    VQ_CUDA(d_hi.alloc((size_t)QT * K * 4));
    VQ_CUDA(d_lo.alloc((size_t)QT * K * 4));
    VQ_CUDA(d_cut.alloc(QT * 4));
    VQ_CUDA(d_counts.alloc(QT * 2 * 8));
    VQ_CUDA(d_cnt.alloc(QT * 4));
    VQ_CUDA(d_keys.alloc((size_t)QT * cap * 8));
    VQ_CUDA(d_rows.alloc((size_t)QT * (topk ? topk : 1) * 8));
    VQ_CUDA(d_sc.alloc((size_t)QT * (topk ? topk : 1) * 4));
    if (scores_dbg_host) VQ_CUDA(d_dbg.alloc((size_t)QT * s->n_rows * 4));
    Dev d_prof;
    const bool want_prof = getenv("VQ_BATCH_PROF") != nullptr;
    if (want_prof) VQ_CUDA(d_prof.alloc((size_t)s->sm_count * 8 * 8));
    cudaEvent_t e0, e1;
    VQ_CUDA(cudaEventCreate(&e0));
    VQ_CUDA(cudaEventCreate(&e1));
    float total_ms = 0.f;
    int rc = 0;
    for (int q0 = 0; q0 < n_queries && rc == 0; q0 += QT) {
        const int nq = (n_queries - q0 < QT) ? (n_queries - q0) : QT;
        VQ_CUDA(cudaMemsetAsync(d_t.p, 0, (size_t)QT * K * 4, st));
        VQ_CUDA(cudaMemcpyAsync(d_t.p, targets + (size_t)q0 * K, (size_t)nq * K * 4, cudaMemcpyHostToDevice, st));
        split_targets<<<(unsigned)(((size_t)QT * K + 255) / 256), 256, 0, st>>>(d_t.as<float>(), d_hi.as<float>(),
                                                                                d_lo.as<float>(), (long long)QT * K);
        fill_f32<<<1, QT, 0, st>>>(d_cut.as<float>(), -INFINITY, QT);
        VQ_CUDA(cudaMemsetAsync(d_counts.p, 0, QT * 2 * 8, st));
        VQ_CUDA(cudaMemsetAsync(d_cnt.p, 0, QT * 4, st));
        CUtensorMap map_a, map_bhi, map_blo, map_apf;
        if (s->n_rows > 0) {
            if ((rc = encode_map(&map_a, s->rows, K, (uint64_t)s->n_rows, BM))) break;
            if ((rc = encode_map(&map_apf, s->rows, K, (uint64_t)s->n_rows, BM, PF_KB * BK, false))) break;
            if ((rc = encode_map(&map_bhi, d_hi.as<float>(), K, QT, pair ? QT / 2 : QT))) break;
            if ((rc = encode_map(&map_blo, d_lo.as<float>(), K, QT, pair ? QT / 2 : QT))) break;
        }
        BatchArgs a;
        for (int i = 0; i < VQ_MAX_STREAMS; ++i) a.w[i] = (i < s->n_streams) ? (float)p->weights[i] : 0.f;
        a.inv_den = (float)(1.0 / den);
        a.inv_splits = (float)(1.0 / (double)s->n_splits);
        a.th_f = float_ceil_of(p->threshold);
        a.lo_f = float_ceil_of(p->lower_limit);
        a.n_queries = nq;
        a.kb_per_stream = s->stream_len / BK;
        a.n_streams = s->n_streams;
        a.n_rows_total = s->n_rows;
        a.cand_cap = cap;
        VQ_CUDA(cudaEventRecord(e0, st));
        for (long long r0 = 0; r0 < s->n_rows; r0 += chunk_rows) {
            const long long nr = (s->n_rows - r0 < chunk_rows) ? (s->n_rows - r0) : chunk_rows;
            a.row0 = r0;
            const int unit_rows = pair ? 2 * BM : BM;
            a.n_tiles = (int)((nr + unit_rows - 1) / unit_rows);
            const int max_units = pair ? s->sm_count / 2 : s->sm_count;
            const int units = a.n_tiles < max_units ? a.n_tiles : max_units;
            float *dbg = scores_dbg_host ? d_dbg.as<float>() : nullptr;
            long long *prof = want_prof ? d_prof.as<long long>() : nullptr;
            if (pair) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(2 * units);
                cfg.blockDim = dim3(BATCH_THREADS);
                cfg.dynamicSmemBytes = Cfg<true>::kSmem;
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = 2;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
                VQ_CUDA(cudaLaunchKernelEx(&cfg, batch_scan<true>, map_a, map_bhi, map_blo, map_apf, a, (const float *)s->inv_counts,
                                           (const float *)d_cut.as<float>(), d_counts.as<unsigned long long>(),
                                           d_cnt.as<unsigned int>(), d_keys.as<unsigned long long>(), dbg, prof));
            } else {
                batch_scan<false><<<units, BATCH_THREADS, Cfg<false>::kSmem, st>>>(
                    map_a, map_bhi, map_blo, map_apf, a, s->inv_counts, d_cut.as<float>(), d_counts.as<unsigned long long>(),
                    d_cnt.as<unsigned int>(), d_keys.as<unsigned long long>(), dbg, prof);
            }
            if (topk > 0)
                batch_compact<<<QT, 1024, 0, st>>>(d_cnt.as<unsigned int>(), d_keys.as<unsigned long long>(), cap, topk,
                                                   d_cut.as<float>());
            else
                VQ_CUDA(cudaMemsetAsync(d_cnt.p, 0, QT * 4, st));
        }
        VQ_CUDA(cudaEventRecord(e1, st));
        if (want_prof) {
            std::vector<long long> h((size_t)s->sm_count * 8);
            VQ_CUDA(cudaMemcpyAsync(h.data(), d_prof.p, h.size() * 8, cudaMemcpyDeviceToHost, st));
            VQ_CUDA(cudaStreamSynchronize(st));
            static const char *names[8] = {"mma thread total", "mma wait conv", "mma wait part_empty", "mma wait full",
                                           "epilogue scoring", "epilogue wait", "epilogue drains+finals", "mma wait lo_empty"};
            for (int c = 0; c < 8; ++c) fprintf(stderr, "[K3 prof, last chunk, CTA 0] %-24s %12lld cycles\n", names[c], h[c]);
        }
        if (topk > 0)
            batch_output<<<QT, 128, 0, st>>>(d_cnt.as<unsigned int>(), d_keys.as<unsigned long long>(), cap, topk,
                                             s->first_global_row, d_rows.as<long long>(), d_sc.as<float>());
        VQ_CUDA(cudaGetLastError());
        if (counts_out) {
            std::vector<unsigned long long> h(QT * 2);
            VQ_CUDA(cudaMemcpyAsync(h.data(), d_counts.p, QT * 2 * 8, cudaMemcpyDeviceToHost, st));
            VQ_CUDA(cudaStreamSynchronize(st));
            for (int q = 0; q < nq; ++q) {
                counts_out[2 * (q0 + q)] = (int64_t)h[2 * q];
                counts_out[2 * (q0 + q) + 1] = (int64_t)h[2 * q + 1];
            }
        }
        if (topk > 0 && topk_rows_out)
            VQ_CUDA(cudaMemcpyAsync(topk_rows_out + (size_t)q0 * topk, d_rows.p, (size_t)nq * topk * 8,
                                    cudaMemcpyDeviceToHost, st));
        if (topk > 0 && topk_scores_out)
            VQ_CUDA(cudaMemcpyAsync(topk_scores_out + (size_t)q0 * topk, d_sc.p, (size_t)nq * topk * 4,
                                    cudaMemcpyDeviceToHost, st));
        if (scores_dbg_host)
            VQ_CUDA(cudaMemcpyAsync(scores_dbg_host + (size_t)q0 * s->n_rows, d_dbg.p, (size_t)nq * s->n_rows * 4,
                                    cudaMemcpyDeviceToHost, st));
        VQ_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) total_ms += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (kernel_ms_out) *kernel_ms_out = total_ms;
    return rc;
}


// Default batched path: bf16x2 kernel (vq_batch_bf16.cuh), 256 queries per pass over the shard.
int run_batch_bf16(vq_store *s, const float *targets, int n_queries, const vq_scan_params *p, int64_t *counts_out,
                   int64_t *topk_rows_out, float *topk_scores_out, float *kernel_ms_out, float *scores_dbg_host) {
    VQ_REQUIRE(s && targets && p, "vq_scan_batch: null argument");
    VQ_REQUIRE(n_queries >= 1, "vq_scan_batch: need at least one query");
    VQ_REQUIRE(s->stream_len % bf::BK == 0, "vq_scan_batch: stream length %d is not a multiple of %d", s->stream_len, bf::BK);
    VQ_REQUIRE(p->topk >= 0 && p->topk <= VQ_MAX_TOPK, "vq_scan_batch: topk %d outside 0..%d", p->topk, VQ_MAX_TOPK);
    VQ_REQUIRE(s->n_rows < (1ll << 31), "vq_scan_batch: shard too large for 32-bit TMA coordinates");
    double den = 0.0;
    for (int i = 0; i < s->n_streams; ++i) den += p->weights[i] * p->weights[i];
    VQ_REQUIRE(den > 0.0, "vq_scan_batch: all stream weights are zero");
    VQ_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = s->stream;
    constexpr int QN = bf::QN;
    const size_t K = s->row_floats;                       // floats per row = S * stream_len
    const int topk = p->topk;
    // chunk schedule: a short first launch (1 tile per CTA) seeds the per-query top-k cuts, so that only the first
    // 19k clips are all candidates; then 8 tiles per CTA per launch
    const long long first_rows = (long long)s->sm_count * bf::BM;
    const long long chunk_rows = (long long)s->sm_count * bf::BM * 8;
    const long long cap = chunk_rows + VQ_MAX_TOPK;
    // CTAs per cluster sharing the query tiles by TMA multicast (VQ_BATCH_CLUSTER = 1, 2 or 4; default 2)
    int kc = getenv("VQ_BATCH_CLUSTER") ? atoi(getenv("VQ_BATCH_CLUSTER")) : 2;
    if (kc != 1 && kc != 2 && kc != 4) kc = 2;
    if (s->sm_count < kc) kc = 1;
    VQ_CUDA(cudaFuncSetAttribute(bf::batch_scan_bf16<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf::SMEM));
    VQ_CUDA(cudaFuncSetAttribute(bf::batch_scan_bf16<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf::SMEM));
    VQ_CUDA(cudaFuncSetAttribute(bf::batch_scan_bf16<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf::SMEM));
    Dev d_t, d_t1, d_t2, d_cut, d_counts, d_cnt, d_keys, d_rows, d_sc, d_dbg, d_park, d_prof;
    VQ_CUDA(d_t.alloc((size_t)QN * K * 4));
    VQ_CUDA(d_t1.alloc((size_t)QN * K * 2));
    VQ_CUDA(d_t2.alloc((size_t)QN * K * 2));
    VQ_CUDA(d_cut.alloc(QN * 4));
    VQ_CUDA(d_counts.alloc(QN * 2 * 8));
    VQ_CUDA(d_cnt.alloc(QN * 4));
    VQ_CUDA(d_keys.alloc((size_t)QN * cap * 8));
    VQ_CUDA(d_rows.alloc((size_t)QN * (topk ? topk : 1) * 8));
    VQ_CUDA(d_sc.alloc((size_t)QN * (topk ? topk : 1) * 4));
    VQ_CUDA(d_park.alloc((size_t)s->sm_count * bf::PARK_FLOATS_PER_CTA * 4));
    if (scores_dbg_host) VQ_CUDA(d_dbg.alloc((size_t)QN * s->n_rows * 4));
    const bool want_prof = getenv("VQ_BATCH_PROF") != nullptr;
    if (want_prof) VQ_CUDA(d_prof.alloc((size_t)s->sm_count * 16 * 8));
    cudaEvent_t e0, e1;
    VQ_CUDA(cudaEventCreate(&e0));
    VQ_CUDA(cudaEventCreate(&e1));
    float total_ms = 0.f;
    int rc = 0;
    for (int q0 = 0; q0 < n_queries && rc == 0; q0 += QN) {
        const int nq = (n_queries - q0 < QN) ? (n_queries - q0) : QN;
        VQ_CUDA(cudaMemsetAsync(d_t.p, 0, (size_t)QN * K * 4, st));
        VQ_CUDA(cudaMemcpyAsync(d_t.p, targets + (size_t)q0 * K, (size_t)nq * K * 4, cudaMemcpyHostToDevice, st));
        bf::split_targets_bf16<<<(unsigned)(((size_t)QN * K + 255) / 256), 256, 0, st>>>(
            d_t.as<float>(), d_t1.as<unsigned short>(), d_t2.as<unsigned short>(), (long long)QN * K);
        fill_f32<<<1, QN, 0, st>>>(d_cut.as<float>(), topk > 0 ? -INFINITY : INFINITY, QN);   // no top-k: nothing is a candidate
        VQ_CUDA(cudaMemsetAsync(d_counts.p, 0, QN * 2 * 8, st));
        VQ_CUDA(cudaMemsetAsync(d_cnt.p, 0, QN * 4, st));
        CUtensorMap map_a, map_t1, map_t2;
        if (s->n_rows > 0) {
            if ((rc = encode_map(&map_a, s->rows, K, (uint64_t)s->n_rows, bf::BM))) break;
            if ((rc = encode_map_ex(&map_t1, d_t1.p, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, QN, QN / kc, bf::BK, CU_TENSOR_MAP_SWIZZLE_64B))) break;
            if ((rc = encode_map_ex(&map_t2, d_t2.p, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, QN, QN / kc, bf::BK, CU_TENSOR_MAP_SWIZZLE_64B))) break;
        }
        BatchArgs a;
        for (int i = 0; i < VQ_MAX_STREAMS; ++i) a.w[i] = (i < s->n_streams) ? (float)p->weights[i] : 0.f;
        a.inv_den = (float)(1.0 / den);
        a.inv_splits = (float)(1.0 / (double)s->n_splits);
        a.th_f = float_ceil_of(p->threshold);
        a.lo_f = float_ceil_of(p->lower_limit);
        a.n_queries = nq;
        a.n_mma = ((nq + 15) / 16) * 16;
        a.kb_per_stream = s->stream_len / bf::BK;
        a.n_streams = s->n_streams;
        a.n_rows_total = s->n_rows;
        a.cand_cap = cap;
        VQ_CUDA(cudaEventRecord(e0, st));
        for (long long r0 = 0, step = first_rows; r0 < s->n_rows; r0 += step, step = chunk_rows) {
            const long long nr = (s->n_rows - r0 < step) ? (s->n_rows - r0) : step;
            a.row0 = r0;
            a.n_tiles = (int)((nr + bf::BM - 1) / bf::BM);
            a.row_end = r0 + nr;
            const int n_unit_tiles = (a.n_tiles + kc - 1) / kc;
            const int max_units = s->sm_count / kc;
            const int units = n_unit_tiles < max_units ? n_unit_tiles : max_units;
            float *dbg = scores_dbg_host ? d_dbg.as<float>() : nullptr;
            long long *prof = want_prof ? d_prof.as<long long>() : nullptr;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(units * kc);
            cfg.blockDim = dim3(bf::THREADS);
            cfg.dynamicSmemBytes = bf::SMEM;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = kc;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            auto kern = kc == 1 ? bf::batch_scan_bf16<1> : (kc == 2 ? bf::batch_scan_bf16<2> : bf::batch_scan_bf16<4>);
            VQ_CUDA(cudaLaunchKernelEx(&cfg, kern, map_a, map_t1, map_t2, a, (const float *)s->inv_counts,
                                       (const float *)d_cut.as<float>(), d_counts.as<unsigned long long>(),
                                       d_cnt.as<unsigned int>(), d_keys.as<unsigned long long>(), d_park.as<float>(), dbg, prof));
            if (topk > 0)
                batch_compact<<<QN, 1024, 0, st>>>(d_cnt.as<unsigned int>(), d_keys.as<unsigned long long>(), cap, topk,
                                                   d_cut.as<float>());
            else
                VQ_CUDA(cudaMemsetAsync(d_cnt.p, 0, QN * 4, st));
        }
        VQ_CUDA(cudaEventRecord(e1, st));
        if (want_prof) {
            std::vector<long long> h((size_t)s->sm_count * 16);
            VQ_CUDA(cudaMemcpyAsync(h.data(), d_prof.p, h.size() * 8, cudaMemcpyDeviceToHost, st));
            VQ_CUDA(cudaStreamSynchronize(st));
            static const char *names[12] = {"mma thread total", "mma wait x_full", "mma wait part_empty", "mma wait t_full",
                                            "converter wait xt_empty", "epilogue wait part_full", "epilogue drains", "converter wait a_full",
                                            "epilogue finals (park st)", "epilogue finals (park ld)", "epilogue scoring", "tiles"};
            for (int c = 0; c < 12; ++c) fprintf(stderr, "[K3 bf16 prof, last chunk, CTA 0] %-28s %12lld\n", names[c], h[c]);
        }
        if (topk > 0)
            batch_output<<<QN, 128, 0, st>>>(d_cnt.as<unsigned int>(), d_keys.as<unsigned long long>(), cap, topk,
                                             s->first_global_row, d_rows.as<long long>(), d_sc.as<float>());
        VQ_CUDA(cudaGetLastError());
        if (counts_out) {
            std::vector<unsigned long long> h(QN * 2);
            VQ_CUDA(cudaMemcpyAsync(h.data(), d_counts.p, QN * 2 * 8, cudaMemcpyDeviceToHost, st));
            VQ_CUDA(cudaStreamSynchronize(st));
            for (int q = 0; q < nq; ++q) {
                counts_out[2 * (q0 + q)] = (int64_t)h[2 * q];
                counts_out[2 * (q0 + q) + 1] = (int64_t)h[2 * q + 1];
            }
        }
        if (topk > 0 && topk_rows_out)
            VQ_CUDA(cudaMemcpyAsync(topk_rows_out + (size_t)q0 * topk, d_rows.p, (size_t)nq * topk * 8,
                                    cudaMemcpyDeviceToHost, st));
        if (topk > 0 && topk_scores_out)
            VQ_CUDA(cudaMemcpyAsync(topk_scores_out + (size_t)q0 * topk, d_sc.p, (size_t)nq * topk * 4,
                                    cudaMemcpyDeviceToHost, st));
        if (scores_dbg_host)
            VQ_CUDA(cudaMemcpyAsync(scores_dbg_host + (size_t)q0 * s->n_rows, d_dbg.p, (size_t)nq * s->n_rows * 4,
                                    cudaMemcpyDeviceToHost, st));
        VQ_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) total_ms += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (kernel_ms_out) *kernel_ms_out = total_ms;
    return rc;
}

bool use_tf32_path() {
    const char *e = getenv("VQ_BATCH_IMPL");
    return e && strcmp(e, "tf32") == 0;
}

}  // namespace

extern "C" int vq_scan_batch(vq_store *s, const float *targets, int32_t n_queries, const vq_scan_params *p,
                             int64_t *counts_out, int64_t *topk_rows_out, float *topk_scores_out,
                             float *kernel_ms_out) {
    if (use_tf32_path()) return run_batch(s, targets, n_queries, p, counts_out, topk_rows_out, topk_scores_out, kernel_ms_out, nullptr);
    return run_batch_bf16(s, targets, n_queries, p, counts_out, topk_rows_out, topk_scores_out, kernel_ms_out, nullptr);
}

extern "C" int vq_scan_batch_scores(vq_store *s, const float *targets, int32_t n_queries, const vq_scan_params *p,
                                    float *scores_out) {
    VQ_REQUIRE(scores_out, "vq_scan_batch_scores: null output");
    VQ_REQUIRE(s && (long long)s->n_rows * 256 <= (1ll << 28), "vq_scan_batch_scores: debug dump limited to 1M rows x 256 queries");
    vq_scan_params q = *p;
    q.topk = 0;
    if (use_tf32_path()) return run_batch(s, targets, n_queries, &q, nullptr, nullptr, nullptr, nullptr, scores_out);
    return run_batch_bf16(s, targets, n_queries, &q, nullptr, nullptr, nullptr, nullptr, scores_out);
}
