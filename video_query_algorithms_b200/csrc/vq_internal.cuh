// Internal definitions shared by the translation units of libvq_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "vq.h"

namespace vq {

constexpr int kHistBins = 4096;         // score histogram for top-k pruning: [-1, 1) in 1/2048 steps
constexpr int kChunkRows = 4096;        // rows per block in the selection kernels
constexpr int kTimeRing = 1024;

void set_error(const char *fmt, ...);

#define VQ_CUDA(call)                                                                   \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess) {                                                       \
            vq::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                  \
                          cudaGetErrorString(e__));                                     \
            (void)cudaGetLastError(); /* a failed allocation must not fail the next launch check */ \
            return -2;                                                                  \
        }                                                                               \
    } while (0)

#define VQ_REQUIRE(cond, ...)                                                           \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            vq::set_error(__VA_ARGS__);                                                 \
            return -1;                                                                  \
        }                                                                               \
    } while (0)

// Device-side copy of the scan parameters.
struct ScanArgs {
    float w[VQ_MAX_STREAMS];       // weights
    float inv_den;                 // 1 / sum w^2
    float inv_splits;              // 1 / n_splits when every slot is present
    double th, lo, eps;
    int topk;
    int want_sims;
};

}  // namespace vq

// The opaque store.  One shard of the clip-major feature database plus the scratch the
// scan writes into.  Everything is allocated once at create time: a scan allocates nothing.
struct vq_store {
    int device = 0;
    int64_t n_rows = 0, first_global_row = 0;
    int64_t capacity = 0;            // rows the buffers are allocated for (>= n_rows; grows with vq_store_append)
    int n_streams = 0, n_splits = 0, dim = 0, stream_len = 0;   // stream_len = n_splits * dim
    size_t row_floats = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;

    float *rows = nullptr;           // [n_rows][n_streams][stream_len]
    float *inv_counts = nullptr;     // [n_rows][n_streams] or null
    float *target = nullptr;         // [n_streams][stream_len] fp32 staging for vq_scan
    float *scores = nullptr;         // [n_rows]
    float *sims = nullptr;           // [n_rows][n_streams] (allocated on first want_sims)
    unsigned int *hist = nullptr;    // [kHistBins] + [1] block ticket + [1] cut bin + u64 best near-miss key + u64 its list position
    int64_t n_chunks = 0;
    unsigned int *chunk_counts = nullptr;    // [3][n_chunks]
    unsigned int *chunk_offsets = nullptr;   // [3][n_chunks]
    int64_t *counts = nullptr;       // [4] n_match n_near n_tie n_topk   (device)
    int64_t *counts_host = nullptr;  // pinned mirror
    uint32_t *list_rows[3] = {nullptr, nullptr, nullptr};   // match / near / tie, capacity n_rows
    float *list_scores[3] = {nullptr, nullptr, nullptr};
    unsigned int *cand_count = nullptr;      // [1]
    unsigned long long *cand_keys = nullptr; // [cand_cap] composite (score, ~row) keys
    int64_t cand_cap = 0;
    float *topk_scores = nullptr;    // [VQ_MAX_TOPK]
    int64_t *topk_rows = nullptr;    // [VQ_MAX_TOPK]
    int last_topk = 0;
    int64_t *pack = nullptr;         // [4 + 2*VQ_MAX_TOPK] counts | top-k rows | top-k score bits (allgather payload)
    // an exchange kernel running on another stream reads `pack`: it records this event (created on first use, owned by the
    // store) behind itself and the next select_compact waits for it before rewriting the payload (vq_exchange.cu)
    cudaEvent_t pack_reader_done = nullptr;
    bool pack_reader_pending = false;

    // timing ring, a slot's three events created on its first use (3 x 1024 cudaEventCreate per store cost ~5 ms up front)
    cudaEvent_t ev_start[vq::kTimeRing] = {}, ev_stop[vq::kTimeRing] = {};   // around K1
    cudaEvent_t ev_sel_stop[vq::kTimeRing] = {};                              // after K2c: ev_stop .. ev_sel_stop = the selection kernels
    int ev_head = 0, ev_count = 0;
    void *pinned_stage = nullptr;    // small pinned buffer for targets / params
    // pinned, device-mapped mirror of the last scan's lists + top-k: vq_scan's publish kernel writes it over PCIe,
    // so one stream synchronisation ends the call; the fetch calls are host memcpys, vq_scan_host_list hands out views
    int64_t *h_rows[3] = {nullptr, nullptr, nullptr};        // GLOBAL rows (written by the publish kernel)
    float *h_scores[3] = {nullptr, nullptr, nullptr};
    int64_t h_cap[3] = {0, 0, 0};
    int64_t *h_topk_rows = nullptr;
    float *h_topk_scores = nullptr;
    int64_t *h_result = nullptr;     // pinned [8]: n_match n_near n_tie n_topk overflow
    int64_t *h_rank_rows = nullptr;  // pinned staging of vq_fetch_ranked
    float *h_rank_scores = nullptr;
    int64_t h_rank_cap = 0;
    bool staged = false;             // the mirror holds all three lists + top-k of the last scan (vq_scan)
    int64_t group_n[2] = {-1, -1};   // first shard of a vq_scan_multi(lists = 1): h_rows[0/1] hold ALL shards' lists, this many entries
    bool staged_ties = false;        // ... at least the tie band + top-k (vq_scan, vq_scan_select)
    void *h_gather = nullptr;        // pinned staging of vq_gather_list / vq_fetch_scores_at (grow-only)
    int64_t h_gather_cap = 0;        // entries
    // grow-only device + pinned scratch of the labelled path and of the list ranking (vq::scratch_reserve): no allocation per call once warm
    char *lab_dev = nullptr;
    size_t lab_dev_cap = 0;
    char *lab_host = nullptr;
    size_t lab_host_cap = 0;
    // scratch of the batched path (vq_batch.cu), allocated on its first call and kept: a batched scan allocates nothing
    float batch_absmax[VQ_MAX_STREAMS] = {0.f, 0.f, 0.f, 0.f};   // max |x| per stream (operand scale of the batched path), valid flag below
    bool batch_absmax_valid = false;                              // reset by every write to the rows
    void *batch_scratch = nullptr;
    void (*batch_scratch_free)(void *) = nullptr;
};

namespace vq {
// Grow-only scratch owned by the store: one device block (s->lab_dev) and one pinned host block (s->lab_host).  With a
// multi-gigabyte shard resident a cudaMalloc / cudaFree pair costs ~2 ms — more than the kernels of a revise round —
// so nothing is allocated per call once the blocks are large enough; copies go through the pinned block on the store's
// stream.  Calls on one store are serialised by contract (vq.h), so one block serves every user.
int scratch_reserve(vq_store *s, size_t dev_bytes, size_t host_bytes);
}  // namespace vq
