// Single-query scan: the HBM-bound hot path.
//
//   K1  scan_rows_*      one pass over the shard: per-stream dot products against the target
//                        (reference ticket.py:146-160), split mean, weighted fusion into the score
//                        (ticket.py:173-180), score histogram for top-k pruning.  8192 B read per
//                        clip (S=2, 1024-d), 4 B written.
//   K2a select_count     per 4096-row chunk: match / near-miss / tie-band counts (ticket.py:325-327)
//                        and top-k candidate collection (rows whose histogram bin can hold the k-th).
//   K2b select_finish    block 0: exclusive scan of chunk counts; block 1: exact top-k of the
//                        candidates (radix select on (score, row) keys + bitonic sort), ranking
//                        rule of ticket.py:266 (score descending, database order among equals).
//   K2c select_compact   order-preserving compaction of the three row lists.
//   publish_results      host-facing tail of vq_scan / vq_scan_select: counts, top-k and lists into a pinned,
//                        device-mapped host mirror (one synchronisation per query, zero-copy views).
//   gather_list          vq_gather_list: the entries at caller-drawn list positions (review rounds sample a few dozen
//                        clips, ticket.py:333,341; the best near miss of ticket.py:335-340 is tracked by K2a / K2c).
//   K7  rank_sort_*      vq_fetch_ranked: the whole match / near-miss list in report order (ticket.py:266).
//
// Summation order is fixed by the launch geometry, so results are deterministic run to run.
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "vq_internal.cuh"
#include "vq_topk.cuh"

namespace {

using vq::kChunkRows;
using vq::kHistBins;
using vq::ScanArgs;

constexpr int kScanThreads = 128;   // 4 warps per block, one clip row per warp iteration

__device__ __forceinline__ float4 ld_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ int score_bin(float sc) {
    // [-1, 1) in steps of 1/2048; monotone in sc; NaN -> 0
    const int b = __float2int_rd((sc + 1.0f) * 2048.0f);
    return min(max(b, 0), kHistBins - 1);
}

template <int S>
__device__ __forceinline__ float fuse_score(const float (&sim)[S], const ScanArgs &a) {
    float ssum = 0.f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const float d = a.w[s] * (1.0f - sim[s]);
        ssum = fmaf(d, d, ssum);
    }
    return 1.0f - __fsqrt_rn(ssum * a.inv_den);
}

// Tail shared by both K1 variants: flush the block histogram, and let the last block to finish
// turn the global histogram into the cut bin (largest bin b with count(bins >= b) >= k).
__device__ void finish_histogram(unsigned int *hist_s, unsigned int *hist_g, int topk) {
    __shared__ unsigned int is_last;
    __shared__ unsigned int part[32];
    __syncthreads();
    for (int b = threadIdx.x; b < kHistBins; b += blockDim.x) {
        const unsigned int c = hist_s[b];
        if (c) atomicAdd(&hist_g[b], c);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&hist_g[kHistBins], 1u) == gridDim.x - 1);
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < 4) hist_g[kHistBins + 2 + threadIdx.x] = 0;   // best near-miss key / list position of this scan (K2a, K2c)
    // suffix counts: thread t owns bins [t*per, (t+1)*per) from the TOP of the range
    const int per = kHistBins / blockDim.x;          // blockDim divides 4096
    const int top = kHistBins - 1 - threadIdx.x * per;
    unsigned int mine = 0;
    for (int j = 0; j < per; ++j) mine += ((volatile unsigned int *)hist_g)[top - j];
    // inclusive scan of `mine` across threads (thread 0 = highest bins)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned int inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
    }
    if (lane == 31) part[wid] = inc;
    __syncthreads();
    unsigned int before = 0;
    for (int w = 0; w < wid; ++w) before += part[w];
    inc += before;
    const unsigned int excl = inc - mine;
    if (threadIdx.x == 0) hist_g[kHistBins + 1] = 0;   // default: everything is a candidate
    __syncthreads();
    if (topk > 0 && excl < (unsigned int)topk && inc >= (unsigned int)topk) {
        unsigned int run = excl;
        for (int j = 0; j < per; ++j) {
            run += ((volatile unsigned int *)hist_g)[top - j];
            if (run >= (unsigned int)topk) {
                hist_g[kHistBins + 1] = (unsigned int)(top - j);
                break;
            }
        }
    }
}

// K1, specialised: S streams of V*128 floats, target held in registers.
template <int S, int V>
__global__ void __launch_bounds__(kScanThreads, 3)
scan_rows_reg(const float4 *__restrict__ rows, const float4 *__restrict__ target,
              const float *__restrict__ inv_counts, const ScanArgs a, const long long n_rows,
              float *__restrict__ scores, float *__restrict__ sims, unsigned int *hist_g) {
    __shared__ unsigned int hist_s[kHistBins];
    for (int b = threadIdx.x; b < kHistBins; b += blockDim.x) hist_s[b] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (kScanThreads / 32) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (kScanThreads / 32);
    float4 t[S][V];
#pragma unroll
    for (int s = 0; s < S; ++s)
#pragma unroll
        for (int i = 0; i < V; ++i) t[s][i] = target[(s * V + i) * 32 + lane];

    for (long long row = warp0; row < n_rows; row += n_warps) {
        const float4 *p = rows + row * (long long)(S * V * 32) + lane;
        float4 x[S][V];
#pragma unroll
        for (int s = 0; s < S; ++s)
#pragma unroll
            for (int i = 0; i < V; ++i) x[s][i] = ld_stream(p + (s * V + i) * 32);
        float sim[S];
#pragma unroll
        for (int s = 0; s < S; ++s) {
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int i = 0; i < V; ++i) {
                a0 = fmaf(x[s][i].x, t[s][i].x, a0);
                a1 = fmaf(x[s][i].y, t[s][i].y, a1);
                a2 = fmaf(x[s][i].z, t[s][i].z, a2);
                a3 = fmaf(x[s][i].w, t[s][i].w, a3);
            }
            const float d = warp_sum((a0 + a1) + (a2 + a3));
            sim[s] = d * (inv_counts ? inv_counts[row * S + s] : a.inv_splits);
        }
        if (lane == 0) {
            const float sc = fuse_score<S>(sim, a);
            scores[row] = sc;
            if (a.want_sims) {
#pragma unroll
                for (int s = 0; s < S; ++s) sims[row * S + s] = sim[s];
            }
            if (a.topk > 0 && sc == sc) atomicAdd(&hist_s[score_bin(sc)], 1u);
        }
    }
    finish_histogram(hist_s, hist_g, a.topk);
}

// K1, generic: any stream count <= 4 and any stream length (multiple of 4 floats); the target sits in shared
// memory.  Used for databases with several splits per stream (fixtures: 3 x 1024 floats per stream).
// A warp walks its rows as one linear sequence of chunks of 32 x U float4 (stream after stream, row after row);
// the 128-bit loads of chunk k+1 are issued before chunk k is consumed (register double buffer), so U loads per
// lane are always in flight whatever the row shape.  Score terms are accumulated in stream order, like fuse_score.
template <int U>
__global__ void __launch_bounds__(kScanThreads)
scan_rows_generic(const float4 *__restrict__ rows, const float4 *__restrict__ target,
                  const float *__restrict__ inv_counts, const ScanArgs a, const long long n_rows, const int n_streams,
                  const int len4, float *__restrict__ scores, float *__restrict__ sims, unsigned int *hist_g) {
    extern __shared__ float4 tgt_s[];                 // [S][len4]
    __shared__ unsigned int hist_s[kHistBins];
    for (int b = threadIdx.x; b < kHistBins; b += blockDim.x) hist_s[b] = 0;
    for (int i = threadIdx.x; i < n_streams * len4; i += blockDim.x) tgt_s[i] = target[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long warp0 = (long long)blockIdx.x * (kScanThreads / 32) + (threadIdx.x >> 5);
    const long long n_warps = (long long)gridDim.x * (kScanThreads / 32);
    const int cps = (len4 + 32 * U - 1) / (32 * U);   // chunks per stream
    const long long row_f4 = (long long)n_streams * len4;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    float4 xn[U];
    long long row = warp0;
    int st = 0, ck = 0;                               // stream and chunk-in-stream of the chunk held in xn
    auto load = [&](long long r, int s_, int c_) {
        const float4 *p = rows + r * row_f4 + (long long)s_ * len4 + c_ * (32 * U) + lane;
        const int j0 = c_ * (32 * U) + lane;
#pragma unroll
        for (int u = 0; u < U; ++u) xn[u] = (r < n_rows && j0 + u * 32 < len4) ? ld_stream(p + u * 32) : zero4;
    };
    if (row < n_rows) load(row, 0, 0);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, ssum = 0.f;
    while (row < n_rows) {
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = xn[u];
        // position of the next chunk, and its loads
        long long nrow = row;
        int nst = st, nck = ck + 1;
        if (nck == cps) { nck = 0; if (++nst == n_streams) { nst = 0; nrow += n_warps; } }
        load(nrow, nst, nck);
        const int j0 = ck * (32 * U) + lane;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + u * 32;
            const float4 tv = (j < len4) ? tgt_s[st * len4 + j] : zero4;
            a0 = fmaf(x[u].x, tv.x, a0);
            a1 = fmaf(x[u].y, tv.y, a1);
            a2 = fmaf(x[u].z, tv.z, a2);
            a3 = fmaf(x[u].w, tv.w, a3);
        }
        if (ck == cps - 1) {                          // end of a stream: similarity, score term
            const float d = warp_sum((a0 + a1) + (a2 + a3)) * (inv_counts ? inv_counts[row * n_streams + st] : a.inv_splits);
            a0 = a1 = a2 = a3 = 0.f;
            const float e = a.w[st] * (1.0f - d);
            ssum = fmaf(e, e, ssum);
            if (a.want_sims && lane == 0) sims[row * n_streams + st] = d;
            if (st == n_streams - 1) {                // end of the row
                if (lane == 0) {
                    const float sc = 1.0f - __fsqrt_rn(ssum * a.inv_den);
                    scores[row] = sc;
                    if (a.topk > 0 && sc == sc) atomicAdd(&hist_s[score_bin(sc)], 1u);
                }
                ssum = 0.f;
            }
        }
        row = nrow; st = nst; ck = nck;
    }
    finish_histogram(hist_s, hist_g, a.topk);
}

// ------------------------------------------------------------------------------ selection
constexpr int kSelThreads = 256;
constexpr int kRowsPerThread = kChunkRows / kSelThreads;   // 16 contiguous rows per thread

using vq::make_key;

struct Flags {
    bool m, nm, tie;
};
__device__ __forceinline__ Flags classify(float sc, const ScanArgs &a) {
    const double d = (double)sc;
    Flags f;
    f.m = d >= a.th;
    f.nm = (!f.m) && (d >= a.lo);
    f.tie = (fabs(d - a.th) < a.eps) || (fabs(d - a.lo) < a.eps);
    return f;
}

__device__ __forceinline__ void load_chunk_scores(const float *scores, long long n_rows, long long r0,
                                                  float (&sc)[kRowsPerThread]) {
    if (r0 + kRowsPerThread <= n_rows) {
        const float4 *p = reinterpret_cast<const float4 *>(scores + r0);
#pragma unroll
        for (int q = 0; q < kRowsPerThread / 4; ++q) {
            const float4 v = p[q];
            sc[4 * q] = v.x; sc[4 * q + 1] = v.y; sc[4 * q + 2] = v.z; sc[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < kRowsPerThread; ++j)
            sc[j] = (r0 + j < n_rows) ? scores[r0 + j] : __int_as_float(0x7fc00000);
    }
}

__global__ void __launch_bounds__(kSelThreads)
select_count(const float *__restrict__ scores, const long long n_rows, const ScanArgs a,
             const unsigned int *__restrict__ hist_g, unsigned int *chunk_counts, const long long n_chunks,
             unsigned int *cand_count, unsigned long long *cand_keys, const long long cand_cap,
             unsigned long long *near_best_key) {
    __shared__ unsigned int red[3][kSelThreads / 32];
    const long long chunk = blockIdx.x;
    const long long r0 = chunk * kChunkRows + (long long)threadIdx.x * kRowsPerThread;
    float sc[kRowsPerThread];
    load_chunk_scores(scores, n_rows, r0, sc);
    const int cut = (int)hist_g[kHistBins + 1];
    unsigned int cm = 0, cn = 0, ct = 0, cc = 0;
    unsigned int cmask = 0;
    unsigned long long best_nm = 0ull;                 // best near miss of this thread: max score, then lowest row
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
        const Flags f = classify(sc[j], a);
        cm += f.m; cn += f.nm; ct += f.tie;
        if (f.nm) {
            const unsigned long long k = make_key(sc[j], (unsigned int)(r0 + j));
            best_nm = k > best_nm ? k : best_nm;
        }
        if (a.topk > 0 && sc[j] == sc[j] && score_bin(sc[j]) >= cut) { cmask |= 1u << j; ++cc; }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // the best near miss of the shard (ticket.py:335-340 holds it out of the sampling): one atomicMax per warp that has one
    if (__any_sync(0xffffffffu, best_nm != 0ull)) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best_nm, o);
            best_nm = other > best_nm ? other : best_nm;
        }
        if (lane == 0) atomicMax(near_best_key, best_nm);
    }
    // top-k candidates: one atomic per warp
    if (a.topk > 0) {
        unsigned int inc = cc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const unsigned int tot = __shfl_sync(0xffffffffu, inc, 31);
        unsigned int base = 0;
        if (lane == 31 && tot) base = atomicAdd(cand_count, tot);
        base = __shfl_sync(0xffffffffu, base, 31);
        unsigned int at = base + inc - cc;
#pragma unroll
        for (int j = 0; j < kRowsPerThread; ++j)
            if (cmask & (1u << j)) {
                if ((long long)at < cand_cap) cand_keys[at] = make_key(sc[j], (unsigned int)(r0 + j));
                ++at;
            }
    }
    unsigned int v0 = cm, v1 = cn, v2 = ct;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
    }
    if (lane == 0) { red[0][wid] = v0; red[1][wid] = v1; red[2][wid] = v2; }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned int t = 0;
        for (int w = 0; w < kSelThreads / 32; ++w) t += red[threadIdx.x][w];
        chunk_counts[threadIdx.x * n_chunks + chunk] = t;
    }
}

constexpr int kFinThreads = 1024;

__device__ unsigned int block_excl_scan_1024(unsigned int v, unsigned int *ws /*[33]*/, unsigned int *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    __syncthreads();
    if (lane == 31) ws[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        unsigned int w = ws[lane];
        unsigned int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int u = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += u;
        }
        ws[lane] = winc - w;
        if (lane == 31) ws[32] = winc;
    }
    __syncthreads();
    *total = ws[32];
    return inc - v + ws[wid];
}

__global__ void __launch_bounds__(kFinThreads)
select_finish(const unsigned int *__restrict__ chunk_counts, unsigned int *chunk_offsets,
              const long long n_chunks, long long *counts, unsigned int *hist_g,
              unsigned int *cand_count, const unsigned long long *__restrict__ cand_keys,
              const long long cand_cap, const int topk, const long long first_global_row,
              float *topk_scores, long long *topk_rows) {
    __shared__ unsigned int ws[33];
    __shared__ vq::TopkScratch tk;
    if (blockIdx.x == 0) {
        // exclusive scan of the three chunk-count arrays; also re-arm the histogram for the next scan
        for (int which = 0; which < 3; ++which) {
            unsigned long long carry = 0;
            for (long long base = 0; base < n_chunks; base += kFinThreads) {
                const long long i = base + threadIdx.x;
                const unsigned int v = (i < n_chunks) ? chunk_counts[which * n_chunks + i] : 0u;
                unsigned int total;
                const unsigned int ex = block_excl_scan_1024(v, ws, &total);
                if (i < n_chunks) chunk_offsets[which * n_chunks + i] = (unsigned int)carry + ex;
                carry += total;
                __syncthreads();
            }
            if (threadIdx.x == 0) counts[which] = (long long)carry;
        }
        for (int b = threadIdx.x; b < kHistBins + 2; b += kFinThreads) hist_g[b] = 0;
        return;
    }
    // ---- block 1: exact top-k over the candidate keys (all keys are distinct)
    long long C = (long long)*cand_count;
    if (C > cand_cap) C = cand_cap;
    const int k = vq::block_topk_1024(cand_keys, C, topk, tk);
    for (int i = threadIdx.x; i < VQ_MAX_TOPK; i += kFinThreads) {
        if (i < k) {
            const unsigned long long key = tk.sel[i];
            topk_scores[i] = vq::key_score(key);
            topk_rows[i] = first_global_row + (long long)vq::key_row(key);
        } else {
            topk_scores[i] = __int_as_float(0xff800000);   // -inf
            topk_rows[i] = -1;
        }
    }
    if (threadIdx.x == 0) {
        counts[3] = k;
        *cand_count = 0;
    }
}

__global__ void __launch_bounds__(kSelThreads)
select_compact(const float *__restrict__ scores, const long long n_rows, const ScanArgs a,
               const unsigned int *__restrict__ chunk_offsets, const long long n_chunks,
               unsigned int *rows_m, float *sc_m, unsigned int *rows_n, float *sc_n,
               unsigned int *rows_t, float *sc_t, const long long *__restrict__ counts,
               const float *__restrict__ topk_scores, const long long *__restrict__ topk_rows,
               long long *pack, const unsigned long long *__restrict__ near_best_key, long long *near_best_pos) {
    __shared__ unsigned int wsum[3][kSelThreads / 32];
    if (blockIdx.x == 0) {
        // allgather payload for the multi-GPU merge: counts | top-k rows | top-k score bits
        const int k = a.topk;
        if (threadIdx.x < 4) pack[threadIdx.x] = counts[threadIdx.x];
        for (int i = threadIdx.x; i < k; i += kSelThreads) {
            pack[4 + i] = topk_rows[i];
            pack[4 + k + i] = (long long)__float_as_uint(topk_scores[i]);
        }
    }
    const long long chunk = blockIdx.x;
    const long long r0 = chunk * kChunkRows + (long long)threadIdx.x * kRowsPerThread;
    float sc[kRowsPerThread];
    load_chunk_scores(scores, n_rows, r0, sc);
    unsigned int mm = 0, mn = 0, mt = 0;
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
        const Flags f = classify(sc[j], a);
        mm |= (unsigned int)f.m << j; mn |= (unsigned int)f.nm << j; mt |= (unsigned int)f.tie << j;
    }
    const unsigned int c[3] = {(unsigned int)__popc(mm), (unsigned int)__popc(mn), (unsigned int)__popc(mt)};
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned int inc[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        unsigned int v = c[q];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int u = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += u;
        }
        inc[q] = v;
        if (lane == 31) wsum[q][wid] = v;
    }
    __syncthreads();
    unsigned int *const out_rows[3] = {rows_m, rows_n, rows_t};
    float *const out_sc[3] = {sc_m, sc_n, sc_t};
    const unsigned int masks[3] = {mm, mn, mt};
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        unsigned int before = 0;
        for (int w = 0; w < wid; ++w) before += wsum[q][w];
        unsigned int at = chunk_offsets[q * n_chunks + chunk] + before + inc[q] - c[q];
        if (c[q]) {
            const unsigned int best_row = (q == 1) ? vq::key_row(*near_best_key) : 0xFFFFFFFFu;
#pragma unroll
            for (int j = 0; j < kRowsPerThread; ++j)
                if (masks[q] & (1u << j)) {
                    out_rows[q][at] = (unsigned int)(r0 + j);
                    out_sc[q][at] = sc[j];
                    if (q == 1 && (unsigned int)(r0 + j) == best_row) *near_best_pos = (long long)at;
                    ++at;
                }
        }
    }
}

// Multi-GPU merge (C1): `gathered` holds n_lists payloads of (4 + 2k) int64 as laid out above,
// one per rank, after the NCCL allgather.  Counts are summed; the global top-k is found by exact
// ranking of the n_lists * k candidates under (score descending, global row ascending).
__global__ void __launch_bounds__(1024)
merge_packed(const long long *__restrict__ gathered, const int n_lists, const int k, long long *out) {
    const int per = 4 + 2 * k;
    const int n = n_lists * k;
    __shared__ unsigned int n_valid;
    if (threadIdx.x == 0) n_valid = 0;
    __syncthreads();
    if (threadIdx.x < 3) {
        long long t = 0;
        for (int l = 0; l < n_lists; ++l) t += gathered[(size_t)l * per + threadIdx.x];
        out[threadIdx.x] = t;
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        out[4 + i] = -1;
        out[4 + k + i] = (long long)0xff800000u;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int li = i / k, ii = i - li * k;
        const long long row = gathered[(size_t)li * per + 4 + ii];
        if (row < 0) continue;
        const float sc = __uint_as_float((unsigned int)gathered[(size_t)li * per + 4 + k + ii]);
        int rank = 0;
        for (int j = 0; j < n; ++j) {
            const int lj = j / k, jj = j - lj * k;
            const long long r2 = gathered[(size_t)lj * per + 4 + jj];
            if (r2 < 0) continue;
            const float s2 = __uint_as_float((unsigned int)gathered[(size_t)lj * per + 4 + k + jj]);
            rank += (s2 > sc) || (s2 == sc && r2 < row);
        }
        atomicAdd(&n_valid, 1u);
        if (rank < k) {
            out[4 + rank] = row;
            out[4 + k + rank] = (long long)__float_as_uint(sc);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[3] = (long long)min((unsigned int)k, n_valid);
}

// Host-facing tail of vq_scan: copies the counts, the top-k and the three ordered lists (rows as GLOBAL int64) into
// pinned, device-mapped host memory with coalesced stores, so that the call needs ONE stream synchronisation and
// no host-side conversion.  Lists longer than the mirror's capacity set the overflow flag (the host grows the
// mirror and publishes again).
struct PublishArgs {
    const long long *counts;             // device [4]
    const unsigned int *rows[3];
    const float *scores[3];
    long long *h_rows[3];
    float *h_scores[3];
    long long cap[3];
    const long long *topk_rows;
    const float *topk_scores;
    long long *h_topk_rows;
    float *h_topk_scores;
    long long *h_result;                 // [8]: counts[4], overflow, best near miss: global row, score bits, list position
    long long first_global_row;
    const unsigned long long *near_best_key;
    const long long *near_best_pos;
    int lists;                           // 1: all three lists; 0: only the tie band (the caller gathers what it samples)
};

__global__ void __launch_bounds__(256)
publish_results(const PublishArgs a) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    long long overflow = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        if (!a.lists && q < 2) continue;
        long long n = a.counts[q];
        if (n > a.cap[q]) { overflow = 1; n = a.cap[q]; }
        for (long long i = tid; i < n; i += nth) {
            a.h_rows[q][i] = a.first_global_row + (long long)a.rows[q][i];
            a.h_scores[q][i] = a.scores[q][i];
        }
    }
    const long long k = a.counts[3];
    for (long long i = tid; i < k; i += nth) {
        a.h_topk_rows[i] = a.topk_rows[i];
        a.h_topk_scores[i] = a.topk_scores[i];
    }
    if (tid < 4) a.h_result[tid] = a.counts[tid];
    if (tid == 4) a.h_result[4] = overflow;
    if (tid == 5) {
        const unsigned long long k = *a.near_best_key;
        a.h_result[5] = k ? a.first_global_row + (long long)vq::key_row(k) : -1;
        a.h_result[6] = (long long)__float_as_uint(vq::key_score(k));
        a.h_result[7] = k ? *a.near_best_pos : -1;
    }
}

// Second phase of vq_scan_multi: this shard's match and near-miss lists into the search set's host mirror (pinned memory
// owned by the first shard, written by every device) at the offsets the host computed from all shards' counts.
__global__ void __launch_bounds__(256)
publish_lists_at(const unsigned int *__restrict__ rows_m, const float *__restrict__ sc_m, long long n_m,
                 const unsigned int *__restrict__ rows_n, const float *__restrict__ sc_n, long long n_n, long long first_global_row,
                 long long *h_rows_m, float *h_sc_m, long long *h_rows_n, float *h_sc_n) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nth = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i < n_m; i += nth) {
        h_rows_m[i] = first_global_row + (long long)rows_m[i];
        h_sc_m[i] = sc_m[i];
    }
    for (long long i = tid; i < n_n; i += nth) {
        h_rows_n[i] = first_global_row + (long long)rows_n[i];
        h_sc_n[i] = sc_n[i];
    }
}

int fill_args(const vq_store *s, const vq_scan_params *p, ScanArgs *a) {
    VQ_REQUIRE(p, "scan: null params");
    VQ_REQUIRE(p->topk >= 0 && p->topk <= VQ_MAX_TOPK, "scan: topk %d outside 0..%d", p->topk, VQ_MAX_TOPK);
    double den = 0.0;
    for (int i = 0; i < VQ_MAX_STREAMS; ++i) {
        a->w[i] = (i < s->n_streams) ? (float)p->weights[i] : 0.f;
        if (i < s->n_streams) den += p->weights[i] * p->weights[i];
    }
    VQ_REQUIRE(den > 0.0, "scan: all stream weights are zero");
    a->inv_den = (float)(1.0 / den);
    a->inv_splits = (float)(1.0 / (double)s->n_splits);
    a->th = p->threshold;
    a->lo = p->lower_limit;
    a->eps = p->eps;
    a->topk = p->topk;
    a->want_sims = p->want_sims;
    return 0;
}

template <int U>
void launch_generic_u(vq_store *s, const float *target_dev, const ScanArgs &a, int grid, cudaStream_t st) {
    const int len4 = s->stream_len / 4;
    const size_t smem = (size_t)s->n_streams * len4 * sizeof(float4);
    cudaFuncSetAttribute(scan_rows_generic<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;                                    // persistent grid: every resident block slot of every SM
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, scan_rows_generic<U>, kScanThreads, smem) == cudaSuccess && per_sm > 0)
        grid = s->sm_count * per_sm;
    scan_rows_generic<U><<<grid, kScanThreads, smem, st>>>(
        reinterpret_cast<const float4 *>(s->rows), reinterpret_cast<const float4 *>(target_dev),
        s->inv_counts, a, s->n_rows, s->n_streams, len4, s->scores, s->sims, s->hist);
}

void launch_generic(vq_store *s, const float *target_dev, const ScanArgs &a, int grid, cudaStream_t st) {
    const int len4 = s->stream_len / 4;
    const char *force = getenv("VQ_SCAN_U");           // development override
    // measured on B200, 8 GB shards (tools/scan_shapes_probe.py), streams of 1 / 2 / 3 x 1024 floats: U = 8 -> 7.02 / 6.21 / 6.23,
    // U = 12 -> 6.79 / 6.06 / 6.13, U = 16 -> 7.05 / 6.87 / 7.19 TB/s (a partial last chunk costs less than fewer loads in
    // flight); short streams keep 8
    const int u = force ? atoi(force) : (len4 >= 512 ? 16 : 8);
    if (u == 12) launch_generic_u<12>(s, target_dev, a, grid, st);
    else if (u == 16) launch_generic_u<16>(s, target_dev, a, grid, st);
    else launch_generic_u<8>(s, target_dev, a, grid, st);
}

}  // namespace

extern "C" int vq_scan_enqueue(vq_store *s, const float *target_dev, const vq_scan_params *p, void *stream) {
    VQ_REQUIRE(s && target_dev, "vq_scan_enqueue: null argument");
    ScanArgs a;
    if (int r = fill_args(s, p, &a)) return r;
    VQ_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    if (a.want_sims && !s->sims)
        VQ_CUDA(cudaMalloc((void **)&s->sims, (size_t)(s->n_rows > 0 ? s->n_rows : 1) * s->n_streams * sizeof(float)));
    s->last_topk = a.topk;
    s->staged = false;
    s->staged_ties = false;
    const size_t smem_need = (size_t)s->n_streams * s->stream_len * sizeof(float);
    VQ_REQUIRE(smem_need <= 200 * 1024, "scan: target of %zu bytes does not fit in shared memory", smem_need);
    const int slot = s->ev_head;
    s->ev_head = (s->ev_head + 1) % vq::kTimeRing;
    if (s->ev_count < vq::kTimeRing) s->ev_count++;
    if (!s->ev_start[slot]) {
        VQ_CUDA(cudaEventCreate(&s->ev_start[slot]));
        VQ_CUDA(cudaEventCreate(&s->ev_stop[slot]));
        VQ_CUDA(cudaEventCreate(&s->ev_sel_stop[slot]));
    }
    VQ_CUDA(cudaEventRecord(s->ev_start[slot], st));
    const int grid = s->sm_count * 3;
    if (s->n_streams == 2 && s->stream_len == 1024 && !getenv("VQ_SCAN_GENERIC")) {
        scan_rows_reg<2, 8><<<grid, kScanThreads, 0, st>>>(
            reinterpret_cast<const float4 *>(s->rows), reinterpret_cast<const float4 *>(target_dev),
            s->inv_counts, a, s->n_rows, s->scores, s->sims, s->hist);
    } else {
        launch_generic(s, target_dev, a, grid, st);
    }
    VQ_CUDA(cudaEventRecord(s->ev_stop[slot], st));
    const unsigned int sel_blocks = (unsigned int)(s->n_chunks > 0 ? s->n_chunks : 1);
    const long long sel_chunks = (long long)sel_blocks;
    select_count<<<sel_blocks, kSelThreads, 0, st>>>(s->scores, s->n_rows, a, s->hist, s->chunk_counts,
                                                     sel_chunks, s->cand_count, s->cand_keys, s->cand_cap,
                                                     reinterpret_cast<unsigned long long *>(s->hist + kHistBins + 2));
    select_finish<<<2, kFinThreads, 0, st>>>(s->chunk_counts, s->chunk_offsets, sel_chunks,
                                             (long long *)s->counts, s->hist, s->cand_count, s->cand_keys,
                                             s->cand_cap, a.topk, s->first_global_row, s->topk_scores,
                                             (long long *)s->topk_rows);
    if (s->pack_reader_pending) {                      // an exchange kernel on another stream may still be reading the payload
        VQ_CUDA(cudaStreamWaitEvent(st, s->pack_reader_done, 0));
        s->pack_reader_pending = false;
    }
    select_compact<<<sel_blocks, kSelThreads, 0, st>>>(
        s->scores, s->n_rows, a, s->chunk_offsets, sel_chunks, s->list_rows[0], s->list_scores[0],
        s->list_rows[1], s->list_scores[1], s->list_rows[2], s->list_scores[2], (const long long *)s->counts,
        s->topk_scores, (const long long *)s->topk_rows, (long long *)s->pack,
        reinterpret_cast<const unsigned long long *>(s->hist + kHistBins + 2), reinterpret_cast<long long *>(s->hist + kHistBins + 4));
    VQ_CUDA(cudaGetLastError());
    VQ_CUDA(cudaEventRecord(s->ev_sel_stop[slot], st));
    return 0;
}

extern "C" int vq_scan_wait(vq_store *s, void *stream, vq_scan_counts *out) {
    VQ_REQUIRE(s, "vq_scan_wait: null store");
    VQ_CUDA(cudaSetDevice(s->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : s->stream;
    VQ_CUDA(cudaMemcpyAsync(s->counts_host, s->counts, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    VQ_CUDA(cudaStreamSynchronize(st));
    if (out) {
        out->n_match = s->counts_host[0];
        out->n_near = s->counts_host[1];
        out->n_tie = s->counts_host[2];
        out->n_topk = (int32_t)s->counts_host[3];
        out->scan_ms = 0.f;
        if (s->ev_count > 0) {
            const int slot = (s->ev_head + vq::kTimeRing - 1) % vq::kTimeRing;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, s->ev_start[slot], s->ev_stop[slot]) == cudaSuccess) out->scan_ms = ms;
        }
    }
    return 0;
}

static int grow_mirror(vq_store *s, int which, int64_t need) {
    if (need <= s->h_cap[which]) return 0;
    if (s->h_rows[which]) cudaFreeHost(s->h_rows[which]);
    if (s->h_scores[which]) cudaFreeHost(s->h_scores[which]);
    s->h_rows[which] = nullptr;
    s->h_scores[which] = nullptr;
    s->h_cap[which] = 0;
    int64_t cap = need + need / 4 + 1024;
    if (cap > s->n_rows) cap = s->n_rows > 0 ? s->n_rows : 1;
    VQ_CUDA(cudaMallocHost((void **)&s->h_rows[which], (size_t)cap * sizeof(int64_t)));
    VQ_CUDA(cudaMallocHost((void **)&s->h_scores[which], (size_t)cap * sizeof(float)));
    s->h_cap[which] = cap;
    return 0;
}

static int publish_launch(vq_store *s, int lists) {
    if (!s->h_topk_rows) {
        VQ_CUDA(cudaMallocHost((void **)&s->h_topk_rows, VQ_MAX_TOPK * sizeof(int64_t)));
        VQ_CUDA(cudaMallocHost((void **)&s->h_topk_scores, VQ_MAX_TOPK * sizeof(float)));
        VQ_CUDA(cudaMallocHost((void **)&s->h_result, 8 * sizeof(int64_t)));
    }
    PublishArgs a;
    a.counts = (const long long *)s->counts;
    for (int i = 0; i < 3; ++i) {
        a.rows[i] = s->list_rows[i];
        a.scores[i] = s->list_scores[i];
        a.h_rows[i] = (long long *)s->h_rows[i];
        a.h_scores[i] = s->h_scores[i];
        a.cap[i] = s->h_cap[i];
    }
    a.topk_rows = (const long long *)s->topk_rows;
    a.topk_scores = s->topk_scores;
    a.h_topk_rows = (long long *)s->h_topk_rows;
    a.h_topk_scores = s->h_topk_scores;
    a.h_result = (long long *)s->h_result;
    a.first_global_row = s->first_global_row;
    a.near_best_key = reinterpret_cast<const unsigned long long *>(s->hist + kHistBins + 2);
    a.near_best_pos = reinterpret_cast<const long long *>(s->hist + kHistBins + 4);
    a.lists = lists;
    publish_results<<<lists ? s->sm_count : 8, 256, 0, s->stream>>>(a);
    VQ_CUDA(cudaGetLastError());
    return 0;
}

static int publish(vq_store *s, int lists) {
    if (int r = publish_launch(s, lists)) return r;
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

// vq_scan / vq_scan_select / vq_scan_multi in two halves: everything that is enqueued (target copy, K1, K2, publish) and
// everything that waits (one stream synchronisation, then host bookkeeping), so that a caller holding several shards
// can enqueue on all of them before it waits on any.
static int scan_host_begin(vq_store *s, const float *target, const vq_scan_params *p, int lists, const char *who) {
    VQ_REQUIRE(s && target, "%s: null argument", who);
    VQ_CUDA(cudaSetDevice(s->device));
    s->group_n[0] = s->group_n[1] = -1;
    const size_t bytes = s->row_floats * sizeof(float);
    memcpy(s->pinned_stage, target, bytes);
    VQ_CUDA(cudaMemcpyAsync(s->target, s->pinned_stage, bytes, cudaMemcpyHostToDevice, s->stream));
    if (int r = vq_scan_enqueue(s, s->target, p, s->stream)) return r;
    // the mirror starts at 1/8 of the shard per list (at least 64k entries) and grows on demand
    const int64_t first_cap = s->n_rows / 8 > 65536 ? s->n_rows / 8 : 65536;
    for (int i = lists ? 0 : 2; i < 3; ++i)
        if (s->h_cap[i] == 0)
            if (int r = grow_mirror(s, i, lists ? first_cap : 65536)) return r;
    return publish_launch(s, lists);
}

static int scan_host_finish(vq_store *s, vq_scan_counts *out, int lists) {
    VQ_CUDA(cudaSetDevice(s->device));
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    if (s->h_result[4]) {                                  // a list outgrew its mirror: grow, publish again
        for (int i = lists ? 0 : 2; i < 3; ++i)
            if (int r = grow_mirror(s, i, s->h_result[i])) return r;
        if (int r = publish(s, lists)) return r;
    }
    for (int i = 0; i < 4; ++i) s->counts_host[i] = s->h_result[i];
    s->staged = lists != 0;
    s->staged_ties = true;
    if (out) {
        out->n_match = s->counts_host[0];
        out->n_near = s->counts_host[1];
        out->n_tie = s->counts_host[2];
        out->n_topk = (int32_t)s->counts_host[3];
        out->scan_ms = 0.f;
        if (s->ev_count > 0) {
            const int slot = (s->ev_head + vq::kTimeRing - 1) % vq::kTimeRing;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, s->ev_start[slot], s->ev_stop[slot]) == cudaSuccess) out->scan_ms = ms;
        }
    }
    return 0;
}

static int scan_host(vq_store *s, const float *target, const vq_scan_params *p, vq_scan_counts *out, int lists, const char *who) {
    if (int r = scan_host_begin(s, target, p, lists, who)) return r;
    return scan_host_finish(s, out, lists);
}

extern "C" int vq_scan(vq_store *s, const float *target, const vq_scan_params *p, vq_scan_counts *out) {
    return scan_host(s, target, p, out, 1, "vq_scan");
}

extern "C" int vq_scan_select(vq_store *s, const float *target, const vq_scan_params *p, vq_scan_counts *out,
                              int64_t *near_best_pos, int64_t *near_best_row, float *near_best_score) {
    if (int r = scan_host(s, target, p, out, 0, "vq_scan_select")) return r;
    if (near_best_pos) *near_best_pos = s->h_result[7];
    if (near_best_row) *near_best_row = s->h_result[5];
    if (near_best_score) {
        const uint32_t bits = (uint32_t)s->h_result[6];
        memcpy(near_best_score, &bits, sizeof(float));
    }
    return 0;
}

// The broker's arrangement (reference src/broker.py:62-92: ONE process, one job at a time): a search set sharded over the
// GPUs of the box is scanned by ONE call.  Enqueueing a shard's work costs ~25 us of host time (a copy, four kernels, a
// publish kernel, three events); done from one thread for eight shards, the last GPU would start 0.2 ms after the first.
// The call therefore hands shards 1..n-1 to a small pool of library threads (created on first use, one per shard, spinning
// for ~2 ms after a job so that a stream of queries finds them awake, then sleeping on a condition variable) and runs
// shard 0 itself; every worker enqueues AND waits for its own shard.  The per-shard top-k lists (ranked, in the pinned
// mirrors) are merged by the caller's thread; counts and the best near miss come back per shard because list positions
// are per shard (the shards' lists, in shard order, ARE the search set's lists in database order).
namespace {

class ShardPool {
public:
    // runs fn(0..n-1): fn(0) on the calling thread, the rest on the pool; returns the first non-zero status and leaves its
    // message in the caller's vq_last_error()
    int run(int n, const std::function<int(int)> &fn) {
        if (n <= 1) return n == 1 ? fn(0) : 0;
        std::lock_guard<std::mutex> call(call_mutex_);              // one multi-shard call at a time (the contract anyway)
        if (pid_ != getpid()) {                                     // a forked child inherits the object, not the threads
            for (auto &t : threads_) t.release();                   // (never joined or destroyed: they do not exist here)
            threads_.clear();
            pid_ = getpid();
        }
        try {
            grow(n - 1);
        } catch (const std::exception &e) {                         // no exception crosses the C ABI: run the shards from this thread
            int rc = 0;
            for (int i = 0; i < n && rc == 0; ++i) rc = fn(i);
            return rc;
        }
        fn_ = &fn;
        n_jobs_ = n;
        remaining_.store(n - 1, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(m_);
            gen_.fetch_add(1, std::memory_order_release);
        }
        cv_.notify_all();
        int rc = fn(0);
        while (remaining_.load(std::memory_order_acquire) != 0) cpu_relax();
        for (int i = 1; i < n && rc == 0; ++i)
            if (rc_[(size_t)i]) {
                rc = rc_[(size_t)i];
                vq::set_error("%s", err_[(size_t)i].c_str());
            }
        return rc;
    }

private:
    static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#else
        std::this_thread::yield();
#endif
    }
    void grow(int workers) {
        while ((int)threads_.size() < workers) {
            const int id = (int)threads_.size() + 1;
            rc_.resize((size_t)id + 1, 0);
            err_.resize((size_t)id + 1);
            const unsigned long long start = gen_.load(std::memory_order_acquire);   // the job of THIS run() is start + 1
            std::thread t([this, id, start] { loop(id, start); });
            t.detach();                                              // they live as long as the process
            threads_.emplace_back(new int(id));
        }
    }
    void loop(int id, unsigned long long seen) {
        for (;;) {
            // spin ~2 ms for the next job, then sleep
            const auto t0 = std::chrono::steady_clock::now();
            while (gen_.load(std::memory_order_acquire) == seen) {
                cpu_relax();
                if (std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(2)) {
                    std::unique_lock<std::mutex> lk(m_);
                    cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
                }
            }
            seen = gen_.load(std::memory_order_acquire);
            if (id < n_jobs_) {
                rc_[(size_t)id] = (*fn_)(id);
                if (rc_[(size_t)id]) err_[(size_t)id] = vq_last_error();
                remaining_.fetch_sub(1, std::memory_order_release);
            }
        }
    }
    std::mutex call_mutex_, m_;
    std::condition_variable cv_;
    std::atomic<unsigned long long> gen_{0};
    std::atomic<int> remaining_{0};
    const std::function<int(int)> *fn_ = nullptr;
    int n_jobs_ = 0;
    std::vector<std::unique_ptr<int>> threads_;                      // one entry per live worker (the threads themselves are detached)
    pid_t pid_ = getpid();
    std::vector<int> rc_;
    std::vector<std::string> err_;
};

ShardPool &shard_pool() {
    static ShardPool *pool = new ShardPool();                        // never destroyed: its threads outlive static destruction
    return *pool;
}

}  // namespace

extern "C" int vq_scan_multi(vq_store *const *shards, int32_t n_shards, const float *target, const vq_scan_params *p,
                             int32_t lists, vq_scan_counts *counts_out, int64_t *near_best_out, int32_t topk_cap,
                             int64_t *topk_rows_out, float *topk_scores_out, int32_t *n_topk_out) {
    VQ_REQUIRE(shards && n_shards >= 1 && target && p && counts_out, "vq_scan_multi: null argument");
    for (int i = 0; i < n_shards; ++i) VQ_REQUIRE(shards[i], "vq_scan_multi: shard %d is null", i);
    // phase 1, all shards in parallel: everything but the two long lists — scan, selection, counts / top-k / tie band /
    // best near miss published — enqueued and waited for per shard
    const bool serial = getenv("VQ_SCAN_MULTI_SERIAL") != nullptr;       // development: everything from the calling thread
    std::function<int(int)> phase1 = [&](int i) -> int {
        if (int r = scan_host_begin(shards[i], target, p, 0, "vq_scan_multi")) return r;
        return scan_host_finish(shards[i], &counts_out[i], 0);
    };
    int rc = 0;
    if (serial) {
        for (int i = 0; i < n_shards; ++i)
            if (int r = scan_host_begin(shards[i], target, p, 0, "vq_scan_multi")) rc = rc ? rc : r;
        for (int i = 0; i < n_shards && rc == 0; ++i)
            if (int r = scan_host_finish(shards[i], &counts_out[i], 0)) rc = r;
    } else {
        rc = shard_pool().run(n_shards, phase1);
    }
    if (rc) {
        for (int j = 0; j < n_shards; ++j) {                 // leave no stream busy behind an error
            cudaSetDevice(shards[j]->device);
            cudaStreamSynchronize(shards[j]->stream);
        }
        return rc;
    }
    vq_store *s0 = shards[0];
    s0->group_n[0] = s0->group_n[1] = -1;
    if (lists) {
        // phase 2: the match / near-miss lists of all shards back to back (shard order = database order) in ONE host
        // mirror, each device writing its segment: the caller gets the search set's lists without a host-side copy
        int64_t tot[2] = {0, 0};
        for (int i = 0; i < n_shards; ++i) { tot[0] += shards[i]->counts_host[0]; tot[1] += shards[i]->counts_host[1]; }
        for (int w = 0; w < 2; ++w)
            if (tot[w] > s0->h_cap[w]) {
                const int64_t keep_rows = s0->n_rows;        // grow_mirror caps at the shard size: the group mirror holds all shards
                int64_t all_rows = 0;
                for (int i = 0; i < n_shards; ++i) all_rows += shards[i]->n_rows;
                s0->n_rows = all_rows;
                VQ_CUDA(cudaSetDevice(s0->device));
                const int r = grow_mirror(s0, w, tot[w]);
                s0->n_rows = keep_rows;
                if (r) return r;
            }
        std::vector<int64_t> off((size_t)n_shards * 2, 0);
        for (int i = 1; i < n_shards; ++i) {
            off[(size_t)i * 2] = off[(size_t)(i - 1) * 2] + shards[i - 1]->counts_host[0];
            off[(size_t)i * 2 + 1] = off[(size_t)(i - 1) * 2 + 1] + shards[i - 1]->counts_host[1];
        }
        std::function<int(int)> phase2 = [&](int i) -> int {
            vq_store *s = shards[i];
            const int64_t nm = s->counts_host[0], nn = s->counts_host[1];
            if (nm + nn == 0) return 0;
            VQ_CUDA(cudaSetDevice(s->device));
            publish_lists_at<<<s->sm_count, 256, 0, s->stream>>>(
                s->list_rows[0], s->list_scores[0], nm, s->list_rows[1], s->list_scores[1], nn, s->first_global_row,
                (long long *)s0->h_rows[0] + off[(size_t)i * 2], s0->h_scores[0] + off[(size_t)i * 2],
                (long long *)s0->h_rows[1] + off[(size_t)i * 2 + 1], s0->h_scores[1] + off[(size_t)i * 2 + 1]);
            VQ_CUDA(cudaGetLastError());
            VQ_CUDA(cudaStreamSynchronize(s->stream));
            return 0;
        };
        if (serial) {
            for (int i = 0; i < n_shards; ++i)
                if (int r = phase2(i)) return r;
        } else if (int r = shard_pool().run(n_shards, phase2)) {
            return r;
        }
        s0->group_n[0] = tot[0];
        s0->group_n[1] = tot[1];
    }
    if (near_best_out)
        for (int i = 0; i < n_shards; ++i) {
            near_best_out[3 * i] = shards[i]->h_result[7];         // position in the shard's near-miss list
            near_best_out[3 * i + 1] = shards[i]->h_result[5];     // global row, -1 = none
            near_best_out[3 * i + 2] = shards[i]->h_result[6];     // score bits
        }
    if (n_topk_out) {
        // k-way merge of the shards' ranked lists under (score descending, global row ascending)
        VQ_REQUIRE(topk_rows_out && topk_scores_out && topk_cap >= 0, "vq_scan_multi: null top-k output");
        std::vector<int> head((size_t)n_shards, 0);
        int n = 0;
        const int want = p->topk < topk_cap ? p->topk : topk_cap;
        for (; n < want; ++n) {
            int best = -1;
            for (int i = 0; i < n_shards; ++i) {
                const int h = head[(size_t)i];
                if (h >= (int)shards[i]->counts_host[3]) continue;
                if (best < 0) { best = i; continue; }
                const int hb = head[(size_t)best];
                const float sa = shards[i]->h_topk_scores[h], sb = shards[best]->h_topk_scores[hb];
                if (sa > sb || (sa == sb && shards[i]->h_topk_rows[h] < shards[best]->h_topk_rows[hb])) best = i;
            }
            if (best < 0) break;
            const int h = head[(size_t)best]++;
            topk_rows_out[n] = shards[best]->h_topk_rows[h];
            topk_scores_out[n] = shards[best]->h_topk_scores[h];
        }
        *n_topk_out = n;
    }
    return 0;
}

// The search set's match (which = 0) or near-miss (1) list of the last vq_scan_multi(lists = 1) whose first shard is `first`:
// read-only views of the host mirror all shards published into, global rows in database order; valid until the next scan.
extern "C" int vq_scan_multi_host_list(vq_store *first, int32_t which, const int64_t **rows, const float **scores, int64_t *n) {
    VQ_REQUIRE(first && rows && scores && n, "vq_scan_multi_host_list: null argument");
    VQ_REQUIRE(which == 0 || which == 1, "vq_scan_multi_host_list: list %d (0 = matches, 1 = near misses)", which);
    VQ_REQUIRE(first->group_n[which] >= 0, "vq_scan_multi_host_list: the last scan on these shards was not vq_scan_multi with lists");
    *rows = first->h_rows[which];
    *scores = first->h_scores[which];
    *n = first->group_n[which];
    return 0;
}

namespace {
__global__ void gather_list(const unsigned int *__restrict__ rows, const float *__restrict__ scores, long long n_list,
                            const long long *__restrict__ pos, int n, long long first_global_row, long long *rows_out,
                            float *scores_out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long p = pos[i];
    const bool ok = p >= 0 && p < n_list;
    rows_out[i] = ok ? first_global_row + (long long)rows[p] : -1;
    scores_out[i] = ok ? scores[p] : __int_as_float(0x7fc00000);
}
}  // namespace

static int grow_gather(vq_store *s, int64_t n) {
    if (n <= s->h_gather_cap) return 0;
    if (s->h_gather) cudaFreeHost(s->h_gather);
    s->h_gather = nullptr;
    s->h_gather_cap = 0;
    const int64_t cap = n + n / 2 + 4096;
    VQ_CUDA(cudaMallocHost((void **)&s->h_gather, (size_t)cap * (8 + 8 + 4)));
    s->h_gather_cap = cap;
    return 0;
}

extern "C" int vq_gather_list(vq_store *s, int32_t which, int64_t n_idx, const int64_t *positions, int64_t *rows_out,
                              float *scores_out) {
    VQ_REQUIRE(s && (n_idx == 0 || (positions && rows_out && scores_out)), "vq_gather_list: null argument");
    VQ_REQUIRE(which >= 0 && which <= 2, "vq_gather_list: list %d outside 0..2 (matches, near misses, ties)", which);
    VQ_REQUIRE(n_idx >= 0 && n_idx < (1ll << 31), "vq_gather_list: %lld positions", (long long)n_idx);
    if (n_idx == 0) return 0;
    VQ_CUDA(cudaSetDevice(s->device));
    const int64_t n_list = s->counts_host[which];
    // one round trip whatever the number of positions: the pinned staging is device-mapped, the kernel reads the
    // positions from it and writes rows / scores into it; a position outside the list comes back as row -1
    if (int r = grow_gather(s, n_idx)) return r;
    long long *h_pos = (long long *)s->h_gather, *h_rows = h_pos + s->h_gather_cap;
    float *h_sc = (float *)(h_rows + s->h_gather_cap);
    memcpy(h_pos, positions, (size_t)n_idx * sizeof(int64_t));
    gather_list<<<(unsigned int)((n_idx + 255) / 256), 256, 0, s->stream>>>(s->list_rows[which], s->list_scores[which], n_list,
                                                                        h_pos, (int)n_idx, s->first_global_row, h_rows, h_sc);
    VQ_CUDA(cudaGetLastError());
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    for (int64_t i = 0; i < n_idx; ++i)
        VQ_REQUIRE(h_rows[i] >= 0, "vq_gather_list: position %lld outside the list of %lld entries",
                   (long long)positions[i], (long long)n_list);
    memcpy(rows_out, h_rows, (size_t)n_idx * sizeof(int64_t));
    memcpy(scores_out, h_sc, (size_t)n_idx * sizeof(float));
    return 0;
}

// The same for a store of several shards, both lists, one call: positions index the search set's lists (the shards' lists in
// shard order); each entry names its list (0 matches, 1 near misses, 2 ties).  The gather kernels are enqueued on every
// shard's stream before any stream is waited for — one round trip for a review round's 20 sampled clips whatever the
// number of shards (ticket.py:333,341).
extern "C" int vq_gather_list_multi(vq_store *const *shards, int32_t n_shards, int64_t n_idx, const int32_t *which,
                                    const int64_t *positions, int64_t *rows_out, float *scores_out) {
    VQ_REQUIRE(shards && n_shards >= 1 && (n_idx == 0 || (which && positions && rows_out && scores_out)),
               "vq_gather_list_multi: null argument");
    VQ_REQUIRE(n_idx >= 0 && n_idx < (1ll << 24), "vq_gather_list_multi: %lld positions", (long long)n_idx);
    for (int i = 0; i < n_shards; ++i) VQ_REQUIRE(shards[i], "vq_gather_list_multi: shard %d is null", i);
    if (n_idx == 0) return 0;
    // per shard and list: the entries it owns, in request order
    std::vector<std::vector<int64_t>> at((size_t)n_shards * 3);
    for (int64_t e = 0; e < n_idx; ++e) {
        const int w = which[e];
        VQ_REQUIRE(w >= 0 && w <= 2, "vq_gather_list_multi: list %d outside 0..2", w);
        int64_t p = positions[e];
        int owner = -1;
        for (int i = 0; i < n_shards && p >= 0; ++i) {
            if (p < shards[i]->counts_host[w]) { owner = i; break; }
            p -= shards[i]->counts_host[w];
        }
        VQ_REQUIRE(owner >= 0, "vq_gather_list_multi: position %lld outside list %d of the search set", (long long)positions[e], w);
        at[(size_t)owner * 3 + w].push_back(e);
    }
    struct Job { int shard, w; int64_t off, n; };
    std::vector<Job> jobs;
    for (int i = 0; i < n_shards; ++i) {
        vq_store *s = shards[i];
        int64_t need = 0;
        for (int w = 0; w < 3; ++w) need += (int64_t)at[(size_t)i * 3 + w].size();
        if (!need) continue;
        VQ_CUDA(cudaSetDevice(s->device));
        if (int r = grow_gather(s, need)) return r;
        long long *h_pos = (long long *)s->h_gather, *h_rows = h_pos + s->h_gather_cap;
        float *h_sc = (float *)(h_rows + s->h_gather_cap);
        int64_t base[3] = {0, 0, 0};
        for (int j = 0; j < i; ++j)
            for (int w = 0; w < 3; ++w) base[w] += shards[j]->counts_host[w];
        int64_t off = 0;
        for (int w = 0; w < 3; ++w) {
            const std::vector<int64_t> &v = at[(size_t)i * 3 + w];
            if (v.empty()) continue;
            for (size_t q = 0; q < v.size(); ++q) h_pos[off + (int64_t)q] = positions[v[q]] - base[w];
            gather_list<<<(unsigned int)((v.size() + 255) / 256), 256, 0, s->stream>>>(
                s->list_rows[w], s->list_scores[w], s->counts_host[w], h_pos + off, (int)v.size(), s->first_global_row,
                h_rows + off, h_sc + off);
            VQ_CUDA(cudaGetLastError());
            jobs.push_back({i, w, off, (int64_t)v.size()});
            off += (int64_t)v.size();
        }
    }
    int last = -1;
    for (const Job &j : jobs) {
        vq_store *s = shards[j.shard];
        if (j.shard != last) {
            VQ_CUDA(cudaSetDevice(s->device));
            VQ_CUDA(cudaStreamSynchronize(s->stream));
            last = j.shard;
        }
        const long long *h_rows = (long long *)s->h_gather + s->h_gather_cap;
        const float *h_sc = (const float *)(h_rows + s->h_gather_cap);
        const std::vector<int64_t> &v = at[(size_t)j.shard * 3 + j.w];
        for (int64_t q = 0; q < j.n; ++q) {
            rows_out[v[(size_t)q]] = h_rows[j.off + q];
            scores_out[v[(size_t)q]] = h_sc[j.off + q];
        }
    }
    return 0;
}

namespace {
__global__ void gather_scores(const float *__restrict__ scores, long long n_rows, const long long *__restrict__ rows, int n,
                              float *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long r = rows[i];
    out[i] = (r >= 0 && r < n_rows) ? scores[r] : __int_as_float(0x7fc00000);
}
}  // namespace

// Scores of the last scan at arbitrary LOCAL rows, one round trip (the forced clips of ticket.py:346-356: the reference
// clip and every user-confirmed clip — thousands on a finalize round with many labels).
extern "C" int vq_fetch_scores_at(vq_store *s, int64_t n, const int64_t *local_rows, float *scores_out) {
    VQ_REQUIRE(s && (n == 0 || (local_rows && scores_out)), "vq_fetch_scores_at: null argument");
    VQ_REQUIRE(n >= 0 && n < (1ll << 31), "vq_fetch_scores_at: %lld rows", (long long)n);
    if (n == 0) return 0;
    for (int64_t i = 0; i < n; ++i)
        VQ_REQUIRE(local_rows[i] >= 0 && local_rows[i] < s->n_rows, "vq_fetch_scores_at: row %lld outside the shard of %lld rows",
                   (long long)local_rows[i], (long long)s->n_rows);
    VQ_CUDA(cudaSetDevice(s->device));
    if (int r = grow_gather(s, n)) return r;
    long long *h_rows = (long long *)s->h_gather;
    float *h_sc = (float *)(h_rows + 2 * s->h_gather_cap);
    memcpy(h_rows, local_rows, (size_t)n * sizeof(int64_t));
    gather_scores<<<(unsigned int)((n + 255) / 256), 256, 0, s->stream>>>(s->scores, s->n_rows, h_rows, (int)n, h_sc);
    VQ_CUDA(cudaGetLastError());
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    memcpy(scores_out, h_sc, (size_t)n * sizeof(float));
    return 0;
}

extern "C" int vq_scan_host_list(vq_store *s, int32_t which, const int64_t **rows, const float **scores, int64_t *n) {
    VQ_REQUIRE(s && rows && scores && n, "vq_scan_host_list: null argument");
    VQ_REQUIRE(which >= 0 && which <= 3, "vq_scan_host_list: list %d outside 0..3 (matches, near misses, ties, top-k)", which);
    VQ_REQUIRE(s->staged || (which == 2 && s->staged_ties) || (which == 3 && s->staged_ties),
               "vq_scan_host_list: the last scan on this store did not publish this list to the host mirror");
    *n = s->counts_host[which];
    *rows = which == 3 ? s->h_topk_rows : s->h_rows[which];
    *scores = which == 3 ? s->h_topk_scores : s->h_scores[which];
    return 0;
}

static int fetch_list(vq_store *s, int which, int64_t cap, int64_t *rows_out, float *scores_out,
                      const char *who) {
    VQ_REQUIRE(s, "%s: null store", who);
    VQ_CUDA(cudaSetDevice(s->device));
    const int64_t n = s->counts_host[which];
    VQ_REQUIRE(cap >= n, "%s: capacity %lld < %lld entries", who, (long long)cap, (long long)n);
    if (n == 0) return 0;
    if (s->staged || (which == 2 && s->staged_ties)) {
        if (rows_out) memcpy(rows_out, s->h_rows[which], (size_t)n * sizeof(int64_t));
        if (scores_out) memcpy(scores_out, s->h_scores[which], (size_t)n * sizeof(float));
        return 0;
    }
    if (rows_out) {
        std::vector<uint32_t> tmp((size_t)n);
        VQ_CUDA(cudaMemcpy(tmp.data(), s->list_rows[which], (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
        for (int64_t i = 0; i < n; ++i) rows_out[i] = s->first_global_row + (int64_t)tmp[(size_t)i];
    }
    if (scores_out)
        VQ_CUDA(cudaMemcpy(scores_out, s->list_scores[which], (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int vq_fetch_matches(vq_store *s, int64_t cap, int64_t *rows_out, float *scores_out) {
    return fetch_list(s, 0, cap, rows_out, scores_out, "vq_fetch_matches");
}
extern "C" int vq_fetch_near(vq_store *s, int64_t cap, int64_t *rows_out, float *scores_out) {
    return fetch_list(s, 1, cap, rows_out, scores_out, "vq_fetch_near");
}
extern "C" int vq_fetch_ties(vq_store *s, int64_t cap, int64_t *rows_out, float *scores_out) {
    return fetch_list(s, 2, cap, rows_out, scores_out, "vq_fetch_ties");
}

extern "C" int vq_fetch_topk(vq_store *s, int32_t cap, int64_t *rows_out, float *scores_out) {
    VQ_REQUIRE(s, "vq_fetch_topk: null store");
    VQ_CUDA(cudaSetDevice(s->device));
    const int n = (int)s->counts_host[3];
    VQ_REQUIRE(cap >= n, "vq_fetch_topk: capacity %d < %d entries", cap, n);
    if (n == 0) return 0;
    if (s->staged || s->staged_ties) {
        if (rows_out) memcpy(rows_out, s->h_topk_rows, (size_t)n * sizeof(int64_t));
        if (scores_out) memcpy(scores_out, s->h_topk_scores, (size_t)n * sizeof(float));
        return 0;
    }
    if (rows_out) VQ_CUDA(cudaMemcpy(rows_out, s->topk_rows, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (scores_out) VQ_CUDA(cudaMemcpy(scores_out, s->topk_scores, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

// ------------------------------------------------------------------------------ full ranking of a list
// GPU-side ranking for the finalize round (ticket.py:266: the report lists every selected clip by score, descending):
// bitonic sort of the list's 64-bit (score, ~row) keys — the same keys as the top-k, so equal scores keep database
// order.  All compare-exchanges run in one direction (mirror step + xor steps), so a list whose length is not a
// power of two needs no padding: a missing partner is the smallest key and never moves.  Steps with partner distance
// < 2048 are fused in shared memory (one 2048-key tile per block); the cand_keys scratch of the scan holds the keys.
namespace {

constexpr int kSortTile = 2048;

__device__ __forceinline__ void cmpx(unsigned long long &a, unsigned long long &b) {
    if (a < b) { const unsigned long long t = a; a = b; b = t; }
}

__global__ void rank_make_keys(const unsigned int *__restrict__ rows, const float *__restrict__ scores, long long n,
                               unsigned long long *keys) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) keys[i] = vq::make_key(scores[i], rows[i]);
}

// tile-local steps: full = every size 2..2048 from scratch; otherwise the tail (strides 1024..1) of one larger size
__global__ void __launch_bounds__(kSortTile / 2)
rank_sort_tile(unsigned long long *keys, long long n, bool full) {
    __shared__ unsigned long long t[kSortTile];
    const long long base = (long long)blockIdx.x * kSortTile;
    for (int i = threadIdx.x; i < kSortTile; i += blockDim.x) t[i] = (base + i < n) ? keys[base + i] : 0ull;
    __syncthreads();
    const int tid = threadIdx.x;
    if (full) {
        for (int size = 2; size <= kSortTile; size <<= 1) {
            {   // mirror step: i <-> i ^ (size - 1) inside each block of `size`
                const int blk = tid / (size / 2), off = tid % (size / 2);
                const int i = blk * size + off, j = blk * size + size - 1 - off;
                cmpx(t[i], t[j]);
                __syncthreads();
            }
            for (int stride = size / 4; stride > 0; stride >>= 1) {
                const int i = 2 * tid - (tid & (stride - 1));
                cmpx(t[i], t[i + stride]);
                __syncthreads();
            }
        }
    } else {
        for (int stride = kSortTile / 2; stride > 0; stride >>= 1) {
            const int i = 2 * tid - (tid & (stride - 1));
            cmpx(t[i], t[i + stride]);
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < kSortTile; i += blockDim.x)
        if (base + i < n) keys[base + i] = t[i];
}

// one global step: mirror (j = i ^ (size - 1)) or xor (j = i ^ stride); pair index p enumerates the lower partners
__global__ void rank_sort_step(unsigned long long *keys, long long n, long long size, long long stride, bool mirror) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long i, j;
    if (mirror) {
        const long long blk = p / (size / 2), off = p % (size / 2);
        i = blk * size + off;
        j = blk * size + size - 1 - off;
    } else {
        i = 2 * p - (p & (stride - 1));
        j = i + stride;
    }
    if (j < n) {                                        // a partner past the end is the smallest key: nothing moves
        unsigned long long a = keys[i], b = keys[j];
        if (a < b) { keys[i] = b; keys[j] = a; }
    }
}

__global__ void rank_unpack(const unsigned long long *__restrict__ keys, long long n, long long first_global_row,
                            long long *rows_out, float *scores_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        rows_out[i] = first_global_row + (long long)vq::key_row(keys[i]);
        scores_out[i] = vq::key_score(keys[i]);
    }
}

}  // namespace

// keys of list `which` in s->cand_keys, sorted descending (n <= cand_cap); second key word = rows (database order among
// equal scores) or a caller-supplied tie-break per list entry
static int rank_sort_list(vq_store *s, int which, int64_t n, const unsigned int *second_dev) {
    cudaStream_t st = s->stream;
    unsigned long long *keys = s->cand_keys;           // free between scans (the top-k pass has consumed it)
    const unsigned int nb = (unsigned int)((n + 255) / 256);
    rank_make_keys<<<nb, 256, 0, st>>>(second_dev ? second_dev : s->list_rows[which], s->list_scores[which], n, keys);
    const unsigned int tiles = (unsigned int)((n + kSortTile - 1) / kSortTile);
    rank_sort_tile<<<tiles, kSortTile / 2, 0, st>>>(keys, n, true);
    long long P = kSortTile;
    while (P < n) P <<= 1;
    for (long long size = 2 * kSortTile; size <= P; size <<= 1) {
        const unsigned int pairs_blocks = (unsigned int)((P / 2 + 255) / 256);
        rank_sort_step<<<pairs_blocks, 256, 0, st>>>(keys, n, size, 0, true);
        for (long long stride = size / 4; stride >= kSortTile; stride >>= 1)
            rank_sort_step<<<pairs_blocks, 256, 0, st>>>(keys, n, size, stride, false);
        rank_sort_tile<<<tiles, kSortTile / 2, 0, st>>>(keys, n, false);
    }
    VQ_CUDA(cudaGetLastError());
    return 0;
}

static int grow_rank_staging(vq_store *s, int64_t n) {
    if (n <= s->h_rank_cap) return 0;
    if (s->h_rank_rows) cudaFreeHost(s->h_rank_rows);
    if (s->h_rank_scores) cudaFreeHost(s->h_rank_scores);
    s->h_rank_rows = nullptr;
    s->h_rank_scores = nullptr;
    s->h_rank_cap = 0;
    const int64_t c = n + n / 4 + 1024;
    VQ_CUDA(cudaMallocHost((void **)&s->h_rank_rows, (size_t)c * sizeof(int64_t)));
    VQ_CUDA(cudaMallocHost((void **)&s->h_rank_scores, (size_t)c * sizeof(float)));
    s->h_rank_cap = c;
    return 0;
}

extern "C" int vq_fetch_ranked(vq_store *s, int32_t which, int64_t cap, int64_t *rows_out, float *scores_out) {
    VQ_REQUIRE(s && rows_out && scores_out, "vq_fetch_ranked: null argument");
    VQ_REQUIRE(which == 0 || which == 1, "vq_fetch_ranked: list %d (0 = matches, 1 = near misses)", which);
    VQ_CUDA(cudaSetDevice(s->device));
    const int64_t n = s->counts_host[which];
    VQ_REQUIRE(cap >= n, "vq_fetch_ranked: capacity %lld < %lld entries", (long long)cap, (long long)n);
    if (n == 0) return 0;
    VQ_REQUIRE(n <= s->cand_cap, "vq_fetch_ranked: list longer than the key scratch");
    if (int r = rank_sort_list(s, which, n, nullptr)) return r;
    // unpack into device-visible pinned staging (its own: the scan's host mirror keeps the database-order lists)
    if (int r = grow_rank_staging(s, n)) return r;
    rank_unpack<<<(unsigned int)((n + 255) / 256), 256, 0, s->stream>>>(s->cand_keys, n, s->first_global_row,
                                                                      (long long *)s->h_rank_rows, s->h_rank_scores);
    VQ_CUDA(cudaGetLastError());
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    memcpy(rows_out, s->h_rank_rows, (size_t)n * sizeof(int64_t));
    memcpy(scores_out, s->h_rank_scores, (size_t)n * sizeof(float));
    return 0;
}

// The report order of ticket.py:266 is a STABLE descending sort of the selected clips in the order the selection put
// them into its dict — for the finalize round a seeded permutation of the lists (ticket.py:333,341) — so equal scores
// keep the selection's order, not the database's.  The caller passes each list entry's rank in that order as the
// tie-break; the list is sorted on the device by (score descending, tie-break ascending) and the tie-breaks come back
// in report order (they identify the entries), with the scores beside them.
extern "C" int vq_rank_list(vq_store *s, int32_t which, int64_t n, const uint32_t *tiebreak, uint32_t *tiebreak_out,
                            float *scores_out) {
    VQ_REQUIRE(s && (n == 0 || (tiebreak && tiebreak_out && scores_out)), "vq_rank_list: null argument");
    VQ_REQUIRE(which == 0 || which == 1, "vq_rank_list: list %d (0 = matches, 1 = near misses)", which);
    VQ_REQUIRE(n == s->counts_host[which], "vq_rank_list: %lld tie-breaks for a list of %lld entries", (long long)n,
               (long long)s->counts_host[which]);
    if (n == 0) return 0;
    VQ_REQUIRE(n <= s->cand_cap, "vq_rank_list: list longer than the key scratch");
    VQ_CUDA(cudaSetDevice(s->device));
    const size_t bytes = (size_t)n * sizeof(uint32_t);
    if (int r = vq::scratch_reserve(s, bytes, bytes)) return r;
    if (int r = grow_rank_staging(s, n)) return r;
    memcpy(s->lab_host, tiebreak, bytes);
    VQ_CUDA(cudaMemcpyAsync(s->lab_dev, s->lab_host, bytes, cudaMemcpyHostToDevice, s->stream));
    if (int r = rank_sort_list(s, which, n, reinterpret_cast<const unsigned int *>(s->lab_dev))) return r;
    // rank_unpack with first_global_row = 0 hands the second key word back as an int64
    rank_unpack<<<(unsigned int)((n + 255) / 256), 256, 0, s->stream>>>(s->cand_keys, n, 0, (long long *)s->h_rank_rows,
                                                                      s->h_rank_scores);
    VQ_CUDA(cudaGetLastError());
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    for (int64_t i = 0; i < n; ++i) tiebreak_out[i] = (uint32_t)s->h_rank_rows[i];
    memcpy(scores_out, s->h_rank_scores, (size_t)n * sizeof(float));
    return 0;
}

extern "C" int vq_fetch_scores(vq_store *s, int64_t first_row, int64_t n_rows, float *scores_out) {
    VQ_REQUIRE(s && (scores_out || n_rows == 0), "vq_fetch_scores: null argument");
    VQ_REQUIRE(first_row >= 0 && n_rows >= 0 && first_row + n_rows <= s->n_rows, "vq_fetch_scores: range outside shard");
    VQ_CUDA(cudaSetDevice(s->device));
    if (n_rows)
        VQ_CUDA(cudaMemcpy(scores_out, s->scores + first_row, (size_t)n_rows * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int vq_fetch_sims(vq_store *s, int64_t first_row, int64_t n_rows, float *sims_out) {
    VQ_REQUIRE(s && (sims_out || n_rows == 0), "vq_fetch_sims: null argument");
    VQ_REQUIRE(s->sims, "vq_fetch_sims: the last scan did not keep similarities (want_sims = 0)");
    VQ_REQUIRE(first_row >= 0 && n_rows >= 0 && first_row + n_rows <= s->n_rows, "vq_fetch_sims: range outside shard");
    VQ_CUDA(cudaSetDevice(s->device));
    if (n_rows)
        VQ_CUDA(cudaMemcpy(sims_out, s->sims + (size_t)first_row * s->n_streams,
                           (size_t)n_rows * s->n_streams * sizeof(float), cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int vq_scan_view(vq_store *s, vq_scan_device_view *out) {
    VQ_REQUIRE(s && out, "vq_scan_view: null argument");
    out->scores_dev = s->scores;
    out->sims_dev = s->sims;
    out->counts_dev = s->counts;
    out->topk_scores_dev = s->topk_scores;
    out->topk_rows_dev = s->topk_rows;
    out->match_rows_dev = s->list_rows[0];
    out->near_rows_dev = s->list_rows[1];
    out->match_scores_dev = s->list_scores[0];
    out->near_scores_dev = s->list_scores[1];
    out->tie_rows_dev = s->list_rows[2];
    out->tie_scores_dev = s->list_scores[2];
    return 0;
}

extern "C" int vq_scan_payload(vq_store *s, const int64_t **payload_dev, int32_t *n_int64) {
    VQ_REQUIRE(s && payload_dev && n_int64, "vq_scan_payload: null argument");
    *payload_dev = s->pack;
    *n_int64 = 4 + 2 * s->last_topk;
    return 0;
}

extern "C" int vq_merge_payloads_enqueue(int device, const int64_t *gathered_dev, int32_t n_lists, int32_t topk,
                                         int64_t *merged_dev, void *stream) {
    VQ_REQUIRE(gathered_dev && merged_dev && n_lists >= 1 && topk >= 0 && topk <= VQ_MAX_TOPK,
               "vq_merge_payloads_enqueue: bad argument");
    VQ_CUDA(cudaSetDevice(device));
    merge_packed<<<1, 1024, 0, (cudaStream_t)stream>>>((const long long *)gathered_dev, n_lists, topk,
                                                       (long long *)merged_dev);
    VQ_CUDA(cudaGetLastError());
    return 0;
}

static int phase_times(vq_store *s, int phase, int32_t cap, float *ms_out, int32_t *n_out, bool reset) {
    VQ_CUDA(cudaSetDevice(s->device));
    int n = s->ev_count < cap ? s->ev_count : cap;
    int got = 0;
    for (int i = 0; i < n; ++i) {
        const int slot = (s->ev_head + vq::kTimeRing - n + i) % vq::kTimeRing;
        float ms = 0.f;
        cudaError_t e = phase == 0 ? cudaEventElapsedTime(&ms, s->ev_start[slot], s->ev_stop[slot])
                                   : cudaEventElapsedTime(&ms, s->ev_stop[slot], s->ev_sel_stop[slot]);
        if (e == cudaSuccess && ms_out) ms_out[got++] = ms;
    }
    cudaGetLastError();
    *n_out = got;
    if (reset) s->ev_count = 0;
    return 0;
}

extern "C" int vq_scan_kernel_times(vq_store *s, int32_t cap, float *ms_out, int32_t *n_out) {
    VQ_REQUIRE(s && n_out, "vq_scan_kernel_times: null argument");
    return phase_times(s, 0, cap, ms_out, n_out, true);
}

extern "C" int vq_scan_phase_times(vq_store *s, int32_t cap, float *scan_ms_out, float *select_ms_out, int32_t *n_out) {
    VQ_REQUIRE(s && n_out, "vq_scan_phase_times: null argument");
    int32_t n0 = 0, n1 = 0;
    if (int r = phase_times(s, 0, cap, scan_ms_out, &n0, false)) return r;
    if (int r = phase_times(s, 1, cap, select_ms_out, &n1, true)) return r;
    *n_out = n0 < n1 ? n0 : n1;
    return 0;
}

extern "C" int vq_merge_topk(int32_t n_lists, int32_t k, const float *scores, const int64_t *rows,
                             float *scores_out, int64_t *rows_out, int32_t *n_out) {
    VQ_REQUIRE(n_lists >= 0 && k >= 0 && scores && rows && scores_out && rows_out && n_out,
               "vq_merge_topk: bad argument");
    std::vector<std::pair<float, int64_t>> all;
    all.reserve((size_t)n_lists * k);
    for (int64_t i = 0; i < (int64_t)n_lists * k; ++i)
        if (rows[i] >= 0) all.emplace_back(scores[i], rows[i]);
    std::sort(all.begin(), all.end(), [](const std::pair<float, int64_t> &x, const std::pair<float, int64_t> &y) {
        return x.first > y.first || (x.first == y.first && x.second < y.second);
    });
    const int n = (int)std::min<size_t>(all.size(), (size_t)k);
    for (int i = 0; i < n; ++i) {
        scores_out[i] = all[(size_t)i].first;
        rows_out[i] = all[(size_t)i].second;
    }
    *n_out = n;
    return 0;
}

// Q queries at once (the batched path's per-shard / per-rank results): lists [n_lists][n_queries][k].
// Each list arrives ranked (that is how batch_output and select_finish write them), so the k best are taken by a
// k-way merge of the list heads, O(k * n_lists) per query; a list that is not ranked falls back to a full sort.
extern "C" int vq_merge_topk_batch(int32_t n_lists, int32_t n_queries, int32_t k, const float *scores, const int64_t *rows,
                                   float *scores_out, int64_t *rows_out, int32_t *n_out) {
    VQ_REQUIRE(n_lists >= 0 && n_queries >= 0 && k >= 0 && scores && rows && scores_out && rows_out && n_out,
               "vq_merge_topk_batch: bad argument");
    auto before = [](float sa, int64_t ra, float sb, int64_t rb) { return sa > sb || (sa == sb && ra < rb); };
    std::vector<std::pair<float, int64_t>> all;
    std::vector<int> head((size_t)n_lists), len((size_t)n_lists);
    for (int q = 0; q < n_queries; ++q) {
        bool ranked = true;
        for (int l = 0; l < n_lists; ++l) {
            const size_t base = ((size_t)l * n_queries + q) * k;
            int n = 0;
            while (n < k && rows[base + n] >= 0) ++n;                 // valid entries come first, padding after
            for (int i = n; i < k && ranked; ++i) ranked = rows[base + i] < 0;
            for (int i = 1; i < n && ranked; ++i)
                ranked = !before(scores[base + i], rows[base + i], scores[base + i - 1], rows[base + i - 1]);
            len[(size_t)l] = n;
            head[(size_t)l] = 0;
        }
        int n = 0;
        if (ranked) {
            for (; n < k; ++n) {
                int best = -1;
                for (int l = 0; l < n_lists; ++l) {
                    if (head[(size_t)l] >= len[(size_t)l]) continue;
                    const size_t i = ((size_t)l * n_queries + q) * k + head[(size_t)l];
                    if (best < 0) {
                        best = l;
                        continue;
                    }
                    const size_t j = ((size_t)best * n_queries + q) * k + head[(size_t)best];
                    if (before(scores[i], rows[i], scores[j], rows[j])) best = l;
                }
                if (best < 0) break;
                const size_t j = ((size_t)best * n_queries + q) * k + head[(size_t)best]++;
                scores_out[(size_t)q * k + n] = scores[j];
                rows_out[(size_t)q * k + n] = rows[j];
            }
        } else {
            all.clear();
            for (int l = 0; l < n_lists; ++l) {
                const size_t base = ((size_t)l * n_queries + q) * k;
                for (int i = 0; i < k; ++i)
                    if (rows[base + i] >= 0) all.emplace_back(scores[base + i], rows[base + i]);
            }
            std::sort(all.begin(), all.end(), [&](const std::pair<float, int64_t> &x, const std::pair<float, int64_t> &y) {
                return before(x.first, x.second, y.first, y.second);
            });
            n = (int)std::min<size_t>(all.size(), (size_t)k);
            for (int i = 0; i < n; ++i) {
                scores_out[(size_t)q * k + i] = all[(size_t)i].first;
                rows_out[(size_t)q * k + i] = all[(size_t)i].second;
            }
        }
        for (int i = n; i < k; ++i) {
            scores_out[(size_t)q * k + i] = -INFINITY;
            rows_out[(size_t)q * k + i] = -1;
        }
        n_out[q] = n;
    }
    return 0;
}
