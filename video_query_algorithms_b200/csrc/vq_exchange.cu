// C1' — peer-memory exchange of the per-rank scan payload, fused with the merge (one process per GPU).
//
// The only cross-GPU step of the path is the merge of per-rank results (counts + top-k, 1.6 KB at
// k = 100).  Instead of an NCCL allgather followed by a merge kernel, each rank runs ONE single-block
// kernel right behind its selection kernels: it stores its payload straight into every peer's inbox over
// NVLink (peer pointers obtained through CUDA IPC), publishes a sequence flag with system-scope release
// semantics, waits for the flags of all peers in its own inbox and merges.
// Two modes.  In-step: kernel i pushes payload i and merges step i (a rank cannot get two steps ahead, because
// finishing step i+1 needs every peer's step-(i+1) payload).  Lagged (a stream of queries): kernel i pushes payload i
// and merges step i-1, whose payloads the peers pushed a whole scan ago — no rank ever waits for the slowest rank of
// the current step, ranks may drift by one step, and a flush kernel merges the last step.  Four inbox slots
// (sequence mod 4) make the lagged mode safe: push i overwrites the slot of step i-4, which a peer reads in its
// kernel i-3; before kernel i starts, this rank's kernel i-1 has seen every peer's flag i-2, i.e. every peer has
// entered its kernel i-2 and therefore finished its kernel i-3.
#include <string.h>

#include "vq_internal.cuh"

struct vq_exchange {
    int device = 0, world = 1, rank = 0;
    long long *inbox = nullptr;                 // [kSlots][world][kSlot] payloads, then [kSlots][world] flags; IPC-exported
    long long *peer_inbox[64] = {nullptr};      // mapped inboxes of all ranks (own = inbox)
    long long **peer_table_dev = nullptr;       // device copy of peer_inbox
    long long *merged = nullptr;                // [kSlot]
    long long *scratch = nullptr;               // [world][kSlot] private copy of the gathered payloads
    unsigned long long seq = 0;
    bool connected = false;
    bool unmerged = false;                      // lagged mode: the last pushed step has not been merged yet
    int topk = 0;
    cudaStream_t stream = nullptr;
};

namespace {
constexpr int kSlot = 4 + 2 * VQ_MAX_TOPK;      // int64 per payload slot
constexpr int kSlots = 4;                       // inbox slots (sequence number mod 4)
__host__ __device__ inline size_t flags_offset(int world) { return (size_t)kSlots * world * kSlot; }

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ long long ld_volatile(const long long *p) {
    long long v;
    asm volatile("ld.volatile.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(1024)
exchange_push_merge(const long long *__restrict__ payload, long long *const *__restrict__ peers, const int world,
                    const int rank, const int k, const unsigned long long push_seq /* 0: nothing to push */,
                    const unsigned long long merge_seq /* 0: nothing to merge */, long long *merged, long long *scratch) {
    const int n_pay = 4 + 2 * k;
    if (push_seq) {
        // 1. push my payload into slot [push_seq mod 4][rank] of every inbox (own included)
        const int slot = (int)(push_seq % kSlots);
        for (int i = threadIdx.x; i < n_pay * world; i += blockDim.x) {
            const int r = i / n_pay, j = i - r * n_pay;
            peers[r][((size_t)slot * world + rank) * kSlot + j] = payload[j];
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < world) {
            unsigned long long *flag =
                reinterpret_cast<unsigned long long *>(peers[threadIdx.x] + flags_offset(world)) + (size_t)slot * world + rank;
            st_release_sys(flag, push_seq);
        }
    }
    if (!merge_seq) return;
    // 2. wait until every rank's payload for the sequence number to merge has landed in my inbox
    const int slot = (int)(merge_seq % kSlots);
    long long *mine = peers[rank];
    if (threadIdx.x < world) {
        const unsigned long long *flag =
            reinterpret_cast<const unsigned long long *>(mine + flags_offset(world)) + (size_t)slot * world + threadIdx.x;
        while (ld_acquire_sys(flag) != merge_seq) __nanosleep(64);
    }
    __syncthreads();
    // 3. private copy of the gathered payloads (peer-written memory is read once, bypassing L1), then merge:
    //    counts summed, global top-k by exact rank under (score desc, global row asc)
    const long long *g = mine + (size_t)slot * world * kSlot;
    for (int i = threadIdx.x; i < n_pay * world; i += blockDim.x) {
        const int r = i / n_pay, j = i - r * n_pay;
        scratch[(size_t)r * n_pay + j] = ld_volatile(g + (size_t)r * kSlot + j);
    }
    __shared__ unsigned int n_valid;
    if (threadIdx.x == 0) n_valid = 0;
    __syncthreads();
    if (threadIdx.x < 3) {
        long long t = 0;
        for (int l = 0; l < world; ++l) t += scratch[(size_t)l * n_pay + threadIdx.x];
        merged[threadIdx.x] = t;
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        merged[4 + i] = -1;
        merged[4 + k + i] = (long long)0xff800000u;
    }
    __syncthreads();
    const int n = world * k;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int li = i / k, ii = i - li * k;
        const long long row = scratch[(size_t)li * n_pay + 4 + ii];
        if (row < 0) continue;
        const float sc = __uint_as_float((unsigned int)scratch[(size_t)li * n_pay + 4 + k + ii]);
        int better = 0;
        for (int lj = 0; lj < world; ++lj) {
            const long long *rows2 = scratch + (size_t)lj * n_pay + 4;
            const long long *sc2 = rows2 + k;
            for (int jj = 0; jj < k; ++jj) {
                const long long r2 = rows2[jj];
                const float s2 = __uint_as_float((unsigned int)sc2[jj]);
                better += (r2 >= 0) && ((s2 > sc) || (s2 == sc && r2 < row));
            }
        }
        atomicAdd(&n_valid, 1u);
        if (better < k) {
            merged[4 + better] = row;
            merged[4 + k + better] = (long long)__float_as_uint(sc);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) merged[3] = (long long)min((unsigned int)k, n_valid);
}
}  // namespace

extern "C" int vq_exchange_create(vq_exchange **out, int device, int world, int rank) {
    VQ_REQUIRE(out, "vq_exchange_create: null output");
    *out = nullptr;
    VQ_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "vq_exchange_create: rank %d of %d", rank, world);
    VQ_CUDA(cudaSetDevice(device));
    vq_exchange *x = new vq_exchange();
    x->device = device;
    x->world = world;
    x->rank = rank;
    const size_t bytes = (flags_offset(world) + (size_t)kSlots * world) * sizeof(long long);
    if (cudaMalloc((void **)&x->inbox, bytes) != cudaSuccess || cudaMalloc((void **)&x->merged, kSlot * 8) != cudaSuccess ||
        cudaMalloc((void **)&x->scratch, (size_t)world * kSlot * 8) != cudaSuccess ||
        cudaMalloc((void **)&x->peer_table_dev, 64 * sizeof(long long *)) != cudaSuccess) {
        vq::set_error("vq_exchange_create: cudaMalloc failed");
        delete x;
        return -3;
    }
    VQ_CUDA(cudaMemset(x->inbox, 0, bytes));
    VQ_CUDA(cudaMemset(x->merged, 0, kSlot * 8));
    x->peer_inbox[rank] = x->inbox;
    *out = x;
    return 0;
}

extern "C" int vq_exchange_local_handle(vq_exchange *x, void *handle_out /* 64 bytes */) {
    VQ_REQUIRE(x && handle_out, "vq_exchange_local_handle: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    VQ_CUDA(cudaSetDevice(x->device));
    cudaIpcMemHandle_t h;
    VQ_CUDA(cudaIpcGetMemHandle(&h, x->inbox));
    memcpy(handle_out, &h, sizeof(h));
    return 0;
}

extern "C" int vq_exchange_connect(vq_exchange *x, const void *all_handles /* [world][64] */) {
    VQ_REQUIRE(x && all_handles, "vq_exchange_connect: null argument");
    VQ_CUDA(cudaSetDevice(x->device));
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char *)all_handles + (size_t)r * 64, 64);
        void *p = nullptr;
        VQ_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        x->peer_inbox[r] = (long long *)p;
    }
    VQ_CUDA(cudaMemcpy(x->peer_table_dev, x->peer_inbox, 64 * sizeof(long long *), cudaMemcpyHostToDevice));
    x->connected = true;
    return 0;
}

extern "C" int vq_exchange_destroy(vq_exchange *x) {
    if (!x) return 0;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank && x->peer_inbox[r]) cudaIpcCloseMemHandle(x->peer_inbox[r]);
    cudaFree(x->inbox);
    cudaFree(x->merged);
    cudaFree(x->scratch);
    cudaFree(x->peer_table_dev);
    delete x;
    return 0;
}

static int exchange_enqueue(vq_store *s, vq_exchange *x, void *stream, bool lagged, const char *who) {
    VQ_REQUIRE(s && x, "%s: null argument", who);
    VQ_REQUIRE(x->connected || x->world == 1, "%s: exchange is not connected", who);
    VQ_REQUIRE(s->device == x->device, "%s: store and exchange live on different devices", who);
    VQ_CUDA(cudaSetDevice(x->device));
    if (x->world == 1 && !x->connected) {
        VQ_CUDA(cudaMemcpy(x->peer_table_dev, x->peer_inbox, 64 * sizeof(long long *), cudaMemcpyHostToDevice));
        x->connected = true;
    }
    x->seq += 1;
    x->topk = s->last_topk;
    x->stream = stream ? (cudaStream_t)stream : s->stream;
    exchange_push_merge<<<1, 1024, 0, x->stream>>>((const long long *)s->pack, x->peer_table_dev, x->world, x->rank, x->topk,
                                                   x->seq, lagged ? x->seq - 1 : x->seq, x->merged, x->scratch);
    VQ_CUDA(cudaGetLastError());
    x->unmerged = lagged;
    return 0;
}

extern "C" int vq_scan_exchange_enqueue(vq_store *s, vq_exchange *x, void *stream) {
    return exchange_enqueue(s, x, stream, false, "vq_scan_exchange_enqueue");
}

extern "C" int vq_scan_exchange_enqueue_lagged(vq_store *s, vq_exchange *x, void *stream) {
    return exchange_enqueue(s, x, stream, true, "vq_scan_exchange_enqueue_lagged");
}

extern "C" int vq_exchange_flush_enqueue(vq_exchange *x, void *stream) {
    VQ_REQUIRE(x, "vq_exchange_flush_enqueue: null exchange");
    if (!x->unmerged || x->seq == 0) return 0;
    VQ_CUDA(cudaSetDevice(x->device));
    exchange_push_merge<<<1, 1024, 0, stream ? (cudaStream_t)stream : x->stream>>>(
        nullptr, x->peer_table_dev, x->world, x->rank, x->topk, 0ull, x->seq, x->merged, x->scratch);
    VQ_CUDA(cudaGetLastError());
    x->unmerged = false;
    return 0;
}

extern "C" int vq_exchange_merged(vq_exchange *x, const int64_t **merged_dev) {
    VQ_REQUIRE(x && merged_dev, "vq_exchange_merged: null argument");
    *merged_dev = (const int64_t *)x->merged;
    return 0;
}
