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
#include <stdlib.h>
#include <string.h>

#include "vq_internal.cuh"

struct vq_exchange {
    int device = 0, world = 1, rank = 0;
    long long *inbox = nullptr;                 // [kSlots][world][kSlot] payloads, then [kSlots][world] flags; IPC-exported
    long long *peer_inbox[64] = {nullptr};      // mapped inboxes of all ranks (own = inbox)
    long long **peer_table_dev = nullptr;       // device copy of peer_inbox
    long long *merged = nullptr;                // [kSlot]
    long long *scratch = nullptr;               // [world][kSlot] private copy of the gathered payloads (merges too large for shared memory)
    unsigned long long seq = 0;
    bool connected = false;
    bool unmerged = false;                      // lagged mode: the last pushed step has not been merged yet
    int topk = 0;
    cudaStream_t stream = nullptr;              // the scan stream of the last enqueue
    // The exchange kernel runs on its own stream behind an event of the scan stream, so that the next step's K1 does not
    // queue behind it (and behind the slowest peer it may wait for); the scan stream waits for it only where it must:
    // before the next select_compact overwrites the payload (vq_store::pack_reader_done).
    cudaStream_t side = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
    unsigned long long *t_ring = nullptr;       // pinned, device-mapped [256][4]: start, pushes done, peers' flags seen, end of each exchange kernel (global timer, ns)
    int ev_head = 0, ev_count = 0;
    double timeout_s = 10.0;                    // a peer that never arrives ends the kernel with an error marker instead of a hang
};

namespace {
constexpr int kSlot = 4 + 2 * VQ_MAX_TOPK;      // int64 per payload slot
constexpr int kSlots = 4;                       // inbox slots (sequence number mod 4)
constexpr int kRing = 256;
constexpr int kXThreads = 128;                   // see exchange_push_merge
constexpr int kMergeShared = 4096;              // candidates (world * k) merged out of shared memory
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
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// (score, row) comes before (s2, r2) in the ranking: score descending, global row ascending (ticket.py:266)
__device__ __forceinline__ bool before(float sa, long long ra, float sb, long long rb) {
    return sa > sb || (sa == sb && ra < rb);
}

// Merge of `world` ranked lists (each sorted under `before`, entries distinct): the rank of entry i of list l in the
// merged order is i + sum over the other lists of the number of their entries that come before it — one binary search
// per other list, O(world * log k) per entry instead of the O(world * k) comparisons of a rank count.
template <class RowPtr, class ScorePtr>
__device__ __forceinline__ void merge_ranked(RowPtr rows, ScorePtr scs, const int *len, const int world, const int k,
                                             const int stride, long long *merged) {
    const int n = world * k;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        const int l = e / k, i = e - l * k;
        if (i >= len[l]) continue;
        const float sc = scs[l * stride + i];
        const long long row = rows[l * stride + i];
        int rank = i;
        for (int l2 = 0; l2 < world && rank < k; ++l2) {
            if (l2 == l) continue;
            int lo = 0, hi = len[l2];
            while (lo < hi) {                                   // first entry of list l2 that does NOT come before (sc, row)
                const int mid = (lo + hi) >> 1;
                if (before(scs[l2 * stride + mid], rows[l2 * stride + mid], sc, row)) lo = mid + 1;
                else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            merged[4 + rank] = row;
            merged[4 + k + rank] = (long long)__float_as_uint(sc);
        }
    }
}

// One block of kXThreads = 128 threads at <= 48 registers: small enough to be co-resident with the scan kernel of the NEXT
// step (K1 keeps 3 blocks of 128 threads x 150 registers on every SM, which leaves 7168 registers per SM), so that it runs
// as soon as its inputs are ready instead of queueing behind a 1.1 ms scan.  t_ns (when not null) receives the kernel's
// own start and end on the global timer: CUDA events around a kernel on a side stream would also count the time it
// waits for an SM.
// Two alternatives were measured and dropped (8 B200s, world 8, k 100, merge phase alone: 25.6 us for the version above):
// a one-warp k-way merge of the list heads (k sequential steps of a shuffle butterfly; 31 us already at world 2 — a dependent
// chain with nothing to hide its latency behind) and eight lists' searches in lockstep with a fixed step count (49.8 us — it
// gives up the early exit above, which ends most entries after one or two lists because their rank already exceeds k).
__global__ void __launch_bounds__(kXThreads, 10)
exchange_push_merge(const long long *__restrict__ payload, long long *const *__restrict__ peers, const int world,
                    const int rank, const int k, const unsigned long long push_seq /* 0: nothing to push */,
                    const unsigned long long merge_seq /* 0: nothing to merge */, long long *merged, long long *scratch,
                    const unsigned long long timeout_ns, unsigned long long *t_ns) {
    const unsigned long long t_begin = global_ns();
    extern __shared__ long long sm_rows[];          // [world * k] rows, then [world * k] fp32 scores (when they fit)
    __shared__ int len_s[64];
    __shared__ int timed_out;
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
    const unsigned long long t_pushed = global_ns();
    if (!merge_seq) {
        if (t_ns && threadIdx.x == 0) { t_ns[0] = t_begin; t_ns[1] = t_pushed; t_ns[2] = t_pushed; t_ns[3] = global_ns(); }
        return;
    }
    // 2. wait until every rank's payload for the sequence number to merge has landed in my inbox; a peer that has died
    //    or fallen out of step must not wedge the GPU: past the deadline the kernel leaves an error marker and ends
    const int slot = (int)(merge_seq % kSlots);
    long long *mine = peers[rank];
    if (threadIdx.x == 0) timed_out = 0;
    __syncthreads();
    if (threadIdx.x < world) {
        const unsigned long long *flag =
            reinterpret_cast<const unsigned long long *>(mine + flags_offset(world)) + (size_t)slot * world + threadIdx.x;
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(flag) != merge_seq) {
            __nanosleep(64);
            if (global_ns() - t0 > timeout_ns) { timed_out = 1; break; }
        }
    }
    __syncthreads();
    const unsigned long long t_seen = global_ns();
    if (timed_out) {
        if (threadIdx.x < 4) merged[threadIdx.x] = -1;         // counts < 0: vq_exchange_check / the host reader raise
        return;
    }
    // 3. counts summed; global top-k by merging the ranks' ranked lists (peer-written memory is read once, bypassing L1)
    const long long *g = mine + (size_t)slot * world * kSlot;
    if (threadIdx.x < world) len_s[threadIdx.x] = (int)ld_volatile(g + (size_t)threadIdx.x * kSlot + 3);
    if (threadIdx.x < 3) {
        long long t = 0;
        for (int l = 0; l < world; ++l) t += ld_volatile(g + (size_t)l * kSlot + threadIdx.x);
        merged[threadIdx.x] = t;
    }
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        merged[4 + i] = -1;
        merged[4 + k + i] = (long long)0xff800000u;
    }
    const int n = world * k;
    if (n <= kMergeShared) {
        float *sm_sc = reinterpret_cast<float *>(sm_rows + n);
        // all of a thread's loads are issued before the first one is used (each is a full round trip to L2 / HBM)
        constexpr int kBatch = 8;
        for (int e0 = threadIdx.x; e0 < n; e0 += blockDim.x * kBatch) {
            long long r[kBatch], sb[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int e = e0 + u * blockDim.x;
                if (e < n) {
                    const int l = e / k, i = e - l * k;
                    r[u] = ld_volatile(g + (size_t)l * kSlot + 4 + i);
                    sb[u] = ld_volatile(g + (size_t)l * kSlot + 4 + k + i);
                }
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int e = e0 + u * blockDim.x;
                if (e < n) {
                    sm_rows[e] = r[u];
                    sm_sc[e] = __uint_as_float((unsigned int)sb[u]);
                }
            }
        }
        __syncthreads();
        merge_ranked(sm_rows, sm_sc, len_s, world, k, k, merged);
    } else {
        float *g_sc = reinterpret_cast<float *>(scratch + n);
        for (int e = threadIdx.x; e < n; e += blockDim.x) {
            const int l = e / k, i = e - l * k;
            scratch[e] = ld_volatile(g + (size_t)l * kSlot + 4 + i);
            g_sc[e] = __uint_as_float((unsigned int)ld_volatile(g + (size_t)l * kSlot + 4 + k + i));
        }
        __syncthreads();
        merge_ranked(scratch, g_sc, len_s, world, k, k, merged);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long tot = 0;
        for (int l = 0; l < world; ++l) tot += len_s[l];
        merged[3] = tot < k ? tot : k;
        if (t_ns) { t_ns[0] = t_begin; t_ns[1] = t_pushed; t_ns[2] = t_seen; t_ns[3] = global_ns(); }
    }
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
    const char *ss = getenv("VQ_EXCHANGE_SIDE_STREAM");
    if (!ss || atoi(ss) != 0) {
        VQ_CUDA(cudaStreamCreateWithFlags(&x->side, cudaStreamNonBlocking));
        VQ_CUDA(cudaEventCreateWithFlags(&x->ev_ready, cudaEventDisableTiming));
        VQ_CUDA(cudaEventCreateWithFlags(&x->ev_done, cudaEventDisableTiming));
    }
    VQ_CUDA(cudaMallocHost((void **)&x->t_ring, kRing * 4 * sizeof(unsigned long long)));
    memset(x->t_ring, 0, kRing * 4 * sizeof(unsigned long long));
    if (const char *t = getenv("VQ_EXCHANGE_TIMEOUT_S")) x->timeout_s = atof(t) > 0 ? atof(t) : x->timeout_s;
    cudaFuncSetAttribute(exchange_push_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, kMergeShared * 12);
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
    if (x->t_ring) cudaFreeHost(x->t_ring);
    if (x->ev_ready) cudaEventDestroy(x->ev_ready);
    if (x->ev_done) cudaEventDestroy(x->ev_done);
    if (x->side) cudaStreamDestroy(x->side);
    cudaFree(x->inbox);
    cudaFree(x->merged);
    cudaFree(x->scratch);
    cudaFree(x->peer_table_dev);
    delete x;
    return 0;
}

static size_t merge_smem(int world, int k) {
    const size_t n = (size_t)world * k;
    return n <= (size_t)kMergeShared ? n * 12 : 0;
}

// launch on the side stream behind the scan stream (or on the scan stream itself), timed by an event pair
static int launch_exchange(vq_exchange *x, vq_store *s, cudaStream_t scan_st, const long long *payload,
                           unsigned long long push_seq, unsigned long long merge_seq) {
    cudaStream_t run = scan_st;
    if (x->side) {
        VQ_CUDA(cudaEventRecord(x->ev_ready, scan_st));
        VQ_CUDA(cudaStreamWaitEvent(x->side, x->ev_ready, 0));
        run = x->side;
    }
    const int slot = x->ev_head;
    x->ev_head = (x->ev_head + 1) % kRing;
    if (x->ev_count < kRing) x->ev_count++;
    exchange_push_merge<<<1, kXThreads, merge_smem(x->world, x->topk), run>>>(
        payload, x->peer_table_dev, x->world, x->rank, x->topk, push_seq, merge_seq, x->merged, x->scratch,
        (unsigned long long)(x->timeout_s * 1e9), x->t_ring + 4 * slot);
    VQ_CUDA(cudaGetLastError());
    if (x->side && s) {                              // the next select_compact on this store waits for this kernel before it rewrites the payload
        if (!s->pack_reader_done) VQ_CUDA(cudaEventCreateWithFlags(&s->pack_reader_done, cudaEventDisableTiming));
        VQ_CUDA(cudaEventRecord(s->pack_reader_done, x->side));
        s->pack_reader_pending = true;
    }
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
    if (int r = launch_exchange(x, s, x->stream, (const long long *)s->pack, x->seq, lagged ? x->seq - 1 : x->seq)) return r;
    x->unmerged = lagged;
    return 0;
}

extern "C" int vq_scan_exchange_enqueue(vq_store *s, vq_exchange *x, void *stream) {
    return exchange_enqueue(s, x, stream, false, "vq_scan_exchange_enqueue");
}

extern "C" int vq_scan_exchange_enqueue_lagged(vq_store *s, vq_exchange *x, void *stream) {
    return exchange_enqueue(s, x, stream, true, "vq_scan_exchange_enqueue_lagged");
}

// Merges the last pushed step (lagged mode) and makes `stream` wait for the exchange stream: after this call the merged
// buffer is ordered behind everything enqueued so far, for readers on `stream`.
extern "C" int vq_exchange_flush_enqueue(vq_exchange *x, void *stream) {
    VQ_REQUIRE(x, "vq_exchange_flush_enqueue: null exchange");
    VQ_CUDA(cudaSetDevice(x->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : x->stream;
    if (x->unmerged && x->seq != 0) {
        if (int r = launch_exchange(x, nullptr, st, nullptr, 0ull, x->seq)) return r;
        x->unmerged = false;
    }
    if (x->side && st) {
        VQ_CUDA(cudaEventRecord(x->ev_done, x->side));
        VQ_CUDA(cudaStreamWaitEvent(st, x->ev_done, 0));
    }
    return 0;
}

extern "C" int vq_exchange_merged(vq_exchange *x, const int64_t **merged_dev) {
    VQ_REQUIRE(x && merged_dev, "vq_exchange_merged: null argument");
    *merged_dev = (const int64_t *)x->merged;
    return 0;
}

// Device times (ms) of the exchange kernels launched since the last call (ring of 256), taken by the kernels themselves on
// the global timer (start of the block to its last store); parts_out splits each into pushes | wait for the peers' flags |
// merge.  Call after a synchronisation.
extern "C" int vq_exchange_kernel_times(vq_exchange *x, int32_t cap, float *ms_out, float *parts_out /* [cap][3] or NULL */, int32_t *n_out) {
    VQ_REQUIRE(x && n_out, "vq_exchange_kernel_times: null argument");
    VQ_CUDA(cudaSetDevice(x->device));
    const int n = x->ev_count < cap ? x->ev_count : cap;
    int got = 0;
    for (int i = 0; i < n; ++i) {
        const int slot = (x->ev_head + kRing - n + i) % kRing;
        const unsigned long long *t = x->t_ring + 4 * slot;
        if (t[3] >= t[0] && t[0] != 0 && ms_out) {
            ms_out[got] = (float)((double)(t[3] - t[0]) * 1e-6);
            if (parts_out) {                                 // push | wait for the peers | merge
                parts_out[3 * got] = (float)((double)(t[1] - t[0]) * 1e-6);
                parts_out[3 * got + 1] = (float)((double)(t[2] - t[1]) * 1e-6);
                parts_out[3 * got + 2] = (float)((double)(t[3] - t[2]) * 1e-6);
            }
            ++got;
        }
    }
    *n_out = got;
    x->ev_count = 0;
    return 0;
}

// After a synchronisation: < 0 with an error text when the last merge gave up waiting for a peer.
extern "C" int vq_exchange_check(vq_exchange *x) {
    VQ_REQUIRE(x, "vq_exchange_check: null exchange");
    VQ_CUDA(cudaSetDevice(x->device));
    long long c[4];
    VQ_CUDA(cudaMemcpy(c, x->merged, sizeof(c), cudaMemcpyDeviceToHost));
    VQ_REQUIRE(c[0] >= 0 && c[3] >= 0, "vq_exchange: rank %d gave up after %.1f s waiting for a peer's payload (step %llu): "
               "a rank died, skipped a step or is out of sequence", x->rank, x->timeout_s, x->seq);
    return 0;
}
