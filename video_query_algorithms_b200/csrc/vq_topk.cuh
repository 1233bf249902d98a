// Block-level exact top-k over 64-bit (score, ~row) keys; shared by the single-query selection
// (vq_scan.cu) and the batched path (vq_batch.cu).  Keys are distinct, larger = better:
// high word = order-preserving image of the fp32 score, low word = 0xFFFFFFFF - row, so sorting keys
// descending yields score descending with database order among equal scores (ticket.py:266).
#pragma once
#include "vq_internal.cuh"

namespace vq {

__device__ __forceinline__ unsigned long long make_key(float sc, unsigned int row) {
    unsigned int u = __float_as_uint(sc);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - row);
}
__device__ __forceinline__ float key_score(unsigned long long key) {
    unsigned int u = (unsigned int)(key >> 32);
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    return __uint_as_float(u);
}
__device__ __forceinline__ unsigned int key_row(unsigned long long key) {
    return 0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull);
}

struct TopkScratch {
    unsigned int digit_hist[256];
    unsigned long long sel[VQ_MAX_TOPK];
    unsigned int sel_n;
    unsigned long long prefix_s;
    unsigned int remain_s;
};

// Called by all 1024 threads of a block.  keys[0..C) in global memory (may alias nothing else that is
// written concurrently).  On return sc.sel[0..k) holds the k = min(topk, C) best keys, sorted descending;
// returns k.  Entries sel[k..P) are zero.
__device__ inline int block_topk_1024(const unsigned long long *keys, long long C, int topk, TopkScratch &sc) {
    const int k = (int)min((long long)topk, C);
    int n_sel = 0;
    if (C <= VQ_MAX_TOPK) {
        // common case (candidates = k + one histogram bin): sort them all, no selection passes
        for (int i = threadIdx.x; i < (int)C; i += blockDim.x) sc.sel[i] = keys[i];
        n_sel = (int)C;
    } else if (k > 0) {
        // radix select, 8 bits at a time from the top: exact key of the k-th best
        unsigned long long prefix = 0, mask = 0;
        unsigned int remain = (unsigned int)k;
        for (int shift = 56; shift >= 0; shift -= 8) {
            if (threadIdx.x < 256) sc.digit_hist[threadIdx.x] = 0;
            __syncthreads();
            for (long long i = threadIdx.x; i < C; i += blockDim.x) {
                const unsigned long long key = keys[i];
                if ((key & mask) == prefix) atomicAdd(&sc.digit_hist[(unsigned int)(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x < 32) {
                // lane l owns digits 255-8l .. 248-8l (descending); find the digit where the running
                // count from the top reaches `remain`
                const int lane = threadIdx.x;
                unsigned int mine = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) mine += sc.digit_hist[255 - 8 * lane - j];
                unsigned int inc = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned int v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                const unsigned int excl = inc - mine;
                if (excl < remain && inc >= remain) {
                    unsigned int run = excl;
                    int d = 255 - 8 * lane;
                    for (int j = 0; j < 8; ++j, --d) {
                        if (run + sc.digit_hist[d] >= remain) break;
                        run += sc.digit_hist[d];
                    }
                    sc.prefix_s = prefix | ((unsigned long long)d << shift);
                    sc.remain_s = remain - run;
                }
            }
            __syncthreads();
            prefix = sc.prefix_s;
            remain = sc.remain_s;
            mask |= 0xFFull << shift;
        }
        const unsigned long long kth = prefix;
        if (threadIdx.x == 0) sc.sel_n = 0;
        __syncthreads();
        for (long long i = threadIdx.x; i < C; i += blockDim.x) {
            const unsigned long long key = keys[i];
            if (key >= kth) {
                const unsigned int at = atomicAdd(&sc.sel_n, 1u);
                if (at < VQ_MAX_TOPK) sc.sel[at] = key;
            }
        }
        __syncthreads();
        n_sel = k;
    }
    int P = 2;
    while (P < n_sel) P <<= 1;                       // sort size: next power of two, <= 1024
    for (int i = threadIdx.x; i < P; i += blockDim.x)
        if (i >= n_sel) sc.sel[i] = 0ull;
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {      // bitonic sort, descending
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            const int t = threadIdx.x;
            if (t < P / 2) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = sc.sel[lo], b = sc.sel[hi];
                if (desc ? (a < b) : (a > b)) { sc.sel[lo] = b; sc.sel[hi] = a; }
            }
            __syncthreads();
        }
    }
    return k;
}

}  // namespace vq
