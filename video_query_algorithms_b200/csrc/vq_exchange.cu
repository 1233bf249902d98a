// C1' — peer-memory exchange of the per-rank scan payload, fused with the merge (one process per GPU).
//
// The only cross-GPU step of the path is the merge of per-rank results (counts + top-k, 1.6 KB at
// k = 100).  Instead of an NCCL allgather followed by a merge kernel, each rank runs ONE small kernel (a block
// per rank) right behind its selection kernels: it stores its payload straight into every peer's inbox over
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
    unsigned long long *sync = nullptr;         // [4 + 4 * kMaxBlocks]: the exchange kernel's completion ticket, error mark and per-block times
    unsigned long long seq = 0;
    bool connected = false;
    bool ipc_mapped = false;                    // peer_inbox entries are IPC mappings to close (vq_exchange_connect)
    bool unmerged = false;                      // lagged mode: the last pushed step has not been merged yet
    int topk = 0;
    cudaStream_t stream = nullptr;              // the scan stream of the last enqueue
    // The exchange kernel runs on its own stream behind an event of the scan stream, so that the next step's K1 does not
    // queue behind it (and behind the slowest peer it may wait for); the scan stream waits for it only where it must:
    // before the next select_compact overwrites the payload (vq_store::pack_reader_done).
    cudaStream_t side = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr;
    unsigned long long *t_ring = nullptr;       // pinned, device-mapped [256][kMaxBlocks][kStamps]: start, pushes done, peers' flags seen, merge end, flags raised of each block of each exchange kernel (global timer, ns)
    int blocks = 1;                             // blocks per exchange kernel
    int ev_head = 0, ev_count = 0;
    double timeout_s = 10.0;                    // a peer that never arrives ends the kernel with an error marker instead of a hang
};

namespace {
constexpr int kSlot = 4 + 2 * VQ_MAX_TOPK;      // int64 per payload slot
constexpr int kSlots = 4;                       // inbox slots (sequence number mod 4)
constexpr int kRing = 256;
constexpr int kXThreads = 128;                   // see exchange_push_merge
constexpr int kMergeThreads = 96;                // warps 0-2 wait for the peers and merge; warp 3 raises the flags
constexpr int kStamps = 8;                       // global-timer stamps per block: start, pushed, peers seen, merge end, flags raised
constexpr int kMaxBlocks = 16;                   // blocks of one exchange kernel (one per rank up to here)
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
__device__ __forceinline__ unsigned int ld_volatile_lo32(const long long *p) {     // low half of a little-endian int64
    unsigned int v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
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
// per other list, O(world * log k) per entry instead of the O(world * k) comparisons of a rank count.  This block takes
// the entries [e_begin, e_end) of the world * k candidates; the blocks of the grid share the work.
template <class Rows, class Scores>
__device__ __forceinline__ void merge_ranked(Rows rows, Scores scs, const int *len, const int world, const int k,
                                             const int stride, const int e_begin, const int e_end, long long *merged) {
    for (int e = e_begin + (int)threadIdx.x; e < e_end; e += kMergeThreads) {
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

// candidates read straight from the inbox (merges too large for shared memory): every access is one volatile load
struct InboxRows {
    const long long *p;
    __device__ __forceinline__ long long operator[](int i) const { return ld_volatile(p + i); }
};
struct InboxScores {
    const long long *p;
    __device__ __forceinline__ float operator[](int i) const { return __uint_as_float((unsigned int)ld_volatile(p + i)); }
};

// A few blocks (one per rank, at most kMaxBlocks) of kXThreads = 128 threads at <= 48 registers: each small enough to be
// co-resident with the scan kernel of the NEXT step (K1 keeps 3 blocks of 128 threads x 150 registers on every SM, which
// leaves 7168 registers per SM), so that the kernel runs as soon as its inputs are ready instead of queueing behind a
// 1.1 ms scan.  Block b stores the payload into the inboxes of the ranks r with r mod grid = b, every block waits for the
// peers' flags by itself and stages all lists in its own shared memory, then ranks its share of the world * k candidates
// (about one candidate per thread at world 8, k 100), pads its share of the tail and leaves its own times; block 0 also
// sums the counts.  No block waits for another.  Round 2's single block took 32 us at world 8 (6.7 us for eight peers' stores and one fence,
// 25.6 us for the merge: 6 candidates x 7 searches per thread, a dependent chain of shared-memory reads).
// The flags are raised by the block's fourth warp alone: a store-release at system scope has to wait until the payload
// stores are acknowledged across NVLink (~2 us on B200 + NVSwitch), and warps 0-2 spend that time waiting for the peers and
// merging (with the release at the end of one instruction stream: 8.6 us at world 2; with every thread fencing: 10 us).
// t_ns (when not null) receives every block's own start, end of pushes, end of waiting and end on the global timer: CUDA
// events around a kernel on a side stream would also count the time it waits for an SM.
// Alternatives measured and dropped in round 2 (8 B200s, world 8, k 100, one block): a one-warp k-way merge of the list
// heads (k sequential steps of a shuffle butterfly; 31 us already at world 2 — a dependent chain with nothing to hide its
// latency behind) and eight lists' searches in lockstep with a fixed step count (49.8 us — it gives up the early exit
// above, which ends most entries after one or two lists because their rank already exceeds k).
__global__ void __launch_bounds__(kXThreads, 9)    // <= 56 registers: 128 x 56 = the 7168 registers K1 leaves free per SM
exchange_push_merge(const long long *__restrict__ payload, long long *const *__restrict__ peers, const int world,
                    const int rank, const int k, const unsigned long long push_seq /* 0: nothing to push */,
                    const unsigned long long merge_seq /* 0: nothing to merge */, long long *merged,
                    unsigned long long *sync /* [1]: sticky mark left by a block that gave up waiting */,
                    const unsigned long long timeout_ns, unsigned long long *t_ns) {
    const unsigned long long t_begin = global_ns();
    extern __shared__ long long sm_rows[];          // [world * k] rows, then [world * k] fp32 scores (when they fit)
    __shared__ int len_s[64];
    __shared__ int timed_out;
    const int n_pay = 4 + 2 * k;
    const int nb = (int)gridDim.x, b = (int)blockIdx.x;
    const int tid = (int)threadIdx.x;

    // 1. my payload into slot [push_seq mod 4][rank] of the inboxes this block serves (own included): all four warps
    const int push_slot = (int)(push_seq % kSlots);
    if (push_seq) {
        for (int r = b; r < world; r += nb)
            for (int j = tid; j < n_pay; j += kXThreads)
                peers[r][((size_t)push_slot * world + rank) * kSlot + j] = payload[j];
    }
    if (tid >= kMergeThreads) {
        // The flag warp: a store-release at system scope waits until the payload stores are acknowledged across NVLink
        // (~2 us).  It waits here, in one warp, while the other three go on to merge.  The named barrier orders every
        // thread's payload stores before the release, which is cumulative over them (the merge warps only arrive).
        if (push_seq) {
            asm volatile("bar.sync 2, %0;" ::"n"(kXThreads) : "memory");
            for (int r = b + (tid - kMergeThreads) * nb; r < world; r += 32 * nb) {
                unsigned long long *flag =
                    reinterpret_cast<unsigned long long *>(peers[r] + flags_offset(world)) + (size_t)push_slot * world + rank;
                st_release_sys(flag, push_seq);
            }
        }
        if (t_ns && tid == kMergeThreads) t_ns[kStamps * b + 4] = global_ns();
        return;
    }
    if (push_seq) asm volatile("bar.arrive 2, %0;" ::"n"(kXThreads) : "memory");
    const unsigned long long ns_push = global_ns() - t_begin;
    unsigned long long ns_wait = 0;
    auto merge_barrier = []() { asm volatile("bar.sync 1, %0;" ::"n"(kMergeThreads) : "memory"); };

    if (tid == 0) timed_out = 0;
    merge_barrier();
    if (merge_seq) {
        // 2. wait until every rank's payload for the sequence number to merge has landed in my inbox; a peer that has
        //    died or fallen out of step must not wedge the GPU: past the deadline the kernel leaves an error mark and ends
        const int slot = (int)(merge_seq % kSlots);
        long long *mine = peers[rank];
        const unsigned long long t0 = global_ns();
        if (tid < world) {
            const unsigned long long *flag =
                reinterpret_cast<const unsigned long long *>(mine + flags_offset(world)) + (size_t)slot * world + tid;
            while (ld_acquire_sys(flag) != merge_seq) {
                __nanosleep(64);
                if (global_ns() - t0 > timeout_ns) { timed_out = 1; break; }
            }
        }
        merge_barrier();
        ns_wait = global_ns() - t0;
        if (timed_out) {
            if (tid == 0) sync[1] = merge_seq;                 // sticky: vq_exchange_check reports and clears it
            if (tid < 4) merged[tid] = -1;
        } else {
            // 3. global top-k by merging the ranks' ranked lists (peer-written memory is read once, bypassing L1).  The
            //    lists' lengths and (block 0) the counts are requested first and used after the staging loads have been
            //    issued: one round trip for all of them
            const long long *g = mine + (size_t)slot * world * kSlot;
            long long len_r = 0, c0 = 0, c1 = 0, c2 = 0;
            if (tid < world) len_r = ld_volatile(g + (size_t)tid * kSlot + 3);
            if (b == 0 && tid < 32) {
                for (int l = tid; l < world; l += 32) {
                    c0 += ld_volatile(g + (size_t)l * kSlot);
                    c1 += ld_volatile(g + (size_t)l * kSlot + 1);
                    c2 += ld_volatile(g + (size_t)l * kSlot + 2);
                }
            }
            const int n = world * k;
            const int per = (n + nb - 1) / nb;
            const int e_begin = b * per, e_end = e_begin + per < n ? e_begin + per : n;
            if (n <= kMergeShared) {
                float *sm_sc = reinterpret_cast<float *>(sm_rows + n);
                // all of a thread's loads are issued before the first one is used (each is a full round trip to L2 / HBM)
                constexpr int kBatch = 9;                      // 2 rounds of 96 threads cover 8 lists of 100
                for (int e0 = tid; e0 < n; e0 += kMergeThreads * kBatch) {
                    long long r[kBatch];
                    unsigned int sb[kBatch];
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const int e = e0 + u * kMergeThreads;
                        if (e < n) {
                            const int l = e / k, i = e - l * k;
                            r[u] = ld_volatile(g + (size_t)l * kSlot + 4 + i);
                            sb[u] = ld_volatile_lo32(g + (size_t)l * kSlot + 4 + k + i);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kBatch; ++u) {
                        const int e = e0 + u * kMergeThreads;
                        if (e < n) {
                            sm_rows[e] = r[u];
                            sm_sc[e] = __uint_as_float(sb[u]);
                        }
                    }
                }
                if (tid < world) len_s[tid] = (int)len_r;
                merge_barrier();
                merge_ranked(sm_rows, sm_sc, len_s, world, k, k, e_begin, e_end, merged);
            } else {
                if (tid < world) len_s[tid] = (int)len_r;
                merge_barrier();
                merge_ranked(InboxRows{g + 4}, InboxScores{g + 4 + k}, len_s, world, k, kSlot, e_begin, e_end, merged);
            }
            // 4. close the step: every block pads its share of the tail, block 0 writes the summed counts
            int tot = 0;
            for (int l = 0; l < world; ++l) tot += len_s[l];
            const int filled = tot < k ? tot : k;
            for (int i = filled + b * kMergeThreads + tid; i < k; i += nb * kMergeThreads) {
                merged[4 + i] = -1;
                merged[4 + k + i] = (long long)0xff800000u;
            }
            if (b == 0 && tid < 32) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    c0 += __shfl_xor_sync(0xffffffffu, c0, o);
                    c1 += __shfl_xor_sync(0xffffffffu, c1, o);
                    c2 += __shfl_xor_sync(0xffffffffu, c2, o);
                }
                if (tid == 0) {
                    merged[0] = c0;
                    merged[1] = c1;
                    merged[2] = c2;
                    merged[3] = filled;
                }
            }
        }
    }
    if (t_ns) {                                                // every block leaves its own times
        merge_barrier();                                       // the merge warps' last stores are issued
        if (tid == 0) {
            unsigned long long *t = t_ns + kStamps * b;
            t[0] = t_begin;
            t[1] = t_begin + ns_push;
            t[2] = t_begin + ns_push + ns_wait;
            t[3] = global_ns();
        }
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
        cudaMalloc((void **)&x->sync, (4 + 4 * kMaxBlocks) * sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc((void **)&x->peer_table_dev, 64 * sizeof(long long *)) != cudaSuccess) {
        vq::set_error("vq_exchange_create: cudaMalloc failed");
        delete x;
        return -3;
    }
    VQ_CUDA(cudaMemset(x->inbox, 0, bytes));
    VQ_CUDA(cudaMemset(x->merged, 0, kSlot * 8));
    VQ_CUDA(cudaMemset(x->sync, 0, (4 + 4 * kMaxBlocks) * sizeof(unsigned long long)));
    x->peer_inbox[rank] = x->inbox;
    const char *ss = getenv("VQ_EXCHANGE_SIDE_STREAM");
    if (!ss || atoi(ss) != 0) {
        VQ_CUDA(cudaStreamCreateWithFlags(&x->side, cudaStreamNonBlocking));
        VQ_CUDA(cudaEventCreateWithFlags(&x->ev_ready, cudaEventDisableTiming));
        VQ_CUDA(cudaEventCreateWithFlags(&x->ev_done, cudaEventDisableTiming));
    }
    VQ_CUDA(cudaMallocHost((void **)&x->t_ring, (size_t)kRing * kMaxBlocks * kStamps * sizeof(unsigned long long)));
    memset(x->t_ring, 0, (size_t)kRing * kMaxBlocks * kStamps * sizeof(unsigned long long));
    x->blocks = world < kMaxBlocks ? world : kMaxBlocks;
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
    x->ipc_mapped = true;
    return 0;
}

extern "C" int vq_exchange_connect_local(vq_exchange *const *all, int32_t world) {
    VQ_REQUIRE(all && world >= 1 && world <= 64, "vq_exchange_connect_local: null argument or world %d", world);
    for (int r = 0; r < world; ++r)
        VQ_REQUIRE(all[r] && all[r]->world == world && all[r]->rank == r && !all[r]->connected,
                   "vq_exchange_connect_local: entry %d is not the unconnected exchange of rank %d of %d", r, r, world);
    for (int r = 0; r < world; ++r) {
        vq_exchange *x = all[r];
        VQ_CUDA(cudaSetDevice(x->device));
        for (int p = 0; p < world; ++p) {
            if (all[p]->device != x->device) {
                int can = 0;
                VQ_CUDA(cudaDeviceCanAccessPeer(&can, x->device, all[p]->device));
                VQ_REQUIRE(can, "vq_exchange_connect_local: device %d cannot address device %d", x->device, all[p]->device);
                const cudaError_t e = cudaDeviceEnablePeerAccess(all[p]->device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else VQ_CUDA(e);
            }
            x->peer_inbox[p] = all[p]->inbox;
        }
        VQ_CUDA(cudaMemcpy(x->peer_table_dev, x->peer_inbox, 64 * sizeof(long long *), cudaMemcpyHostToDevice));
        x->connected = true;
    }
    return 0;
}

extern "C" int vq_exchange_destroy(vq_exchange *x) {
    if (!x) return 0;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world && x->ipc_mapped; ++r)
        if (r != x->rank && x->peer_inbox[r]) cudaIpcCloseMemHandle(x->peer_inbox[r]);
    if (x->t_ring) cudaFreeHost(x->t_ring);
    if (x->ev_ready) cudaEventDestroy(x->ev_ready);
    if (x->ev_done) cudaEventDestroy(x->ev_done);
    if (x->side) cudaStreamDestroy(x->side);
    cudaFree(x->inbox);
    cudaFree(x->merged);
    cudaFree(x->sync);
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
    memset(x->t_ring + (size_t)slot * kMaxBlocks * kStamps, 0, (size_t)kMaxBlocks * kStamps * sizeof(unsigned long long));
    exchange_push_merge<<<x->blocks, kXThreads, merge_smem(x->world, x->topk), run>>>(
        payload, x->peer_table_dev, x->world, x->rank, x->topk, push_seq, merge_seq, x->merged, x->sync,
        (unsigned long long)(x->timeout_s * 1e9), x->t_ring + (size_t)slot * kMaxBlocks * kStamps);
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
        const unsigned long long *t = x->t_ring + (size_t)slot * kMaxBlocks * kStamps;
        // the kernel = its blocks: first start to last end; pushes / waiting = the slowest block's
        unsigned long long t0 = 0, t3 = 0, push = 0, wait = 0;
        bool complete = true;
        for (int b = 0; b < x->blocks; ++b) {
            const unsigned long long *tb = t + kStamps * b;
            if (tb[0] == 0 || tb[3] < tb[0] || tb[4] == 0) { complete = false; break; }
            if (b == 0 || tb[0] < t0) t0 = tb[0];
            if (tb[3] > t3) t3 = tb[3];
            if (tb[4] > t3) t3 = tb[4];                      // the flag warp ends by itself
            if (tb[1] - tb[0] > push) push = tb[1] - tb[0];
            if (tb[2] - tb[1] > wait) wait = tb[2] - tb[1];
        }
        if (complete && ms_out) {
            const double total = (double)(t3 - t0);
            ms_out[got] = (float)(total * 1e-6);
            if (parts_out) {                                 // push | wait for the peers | merge and the rest
                parts_out[3 * got] = (float)((double)push * 1e-6);
                parts_out[3 * got + 1] = (float)((double)wait * 1e-6);
                const double rest = total - (double)push - (double)wait;
                parts_out[3 * got + 2] = (float)((rest > 0 ? rest : 0) * 1e-6);
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
    unsigned long long mark = 0;
    VQ_CUDA(cudaMemcpy(c, x->merged, sizeof(c), cudaMemcpyDeviceToHost));
    VQ_CUDA(cudaMemcpy(&mark, x->sync + 1, sizeof(mark), cudaMemcpyDeviceToHost));
    if (mark) VQ_CUDA(cudaMemset(x->sync + 1, 0, sizeof(mark)));           // reported once
    VQ_REQUIRE(mark == 0 && c[0] >= 0 && c[3] >= 0, "vq_exchange: rank %d gave up after %.1f s waiting for a peer's payload (step %llu): "
               "a rank died, skipped a step or is out of sequence", x->rank, x->timeout_s, mark ? mark : x->seq);
    return 0;
}
