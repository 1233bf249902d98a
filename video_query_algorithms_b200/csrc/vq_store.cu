// Feature store: HBM-resident, clip-major fp32 shard (replaces the per-job HTTP fetch of
// reference Ticket._get_candidate_features, src/models/ticket.py:358-382), plus the VQSYN-1
// synthetic generator (CPU twin: oracle/synth.py).
#include <stdarg.h>
#include <string.h>

#include "vq_internal.cuh"

namespace vq {
static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace vq

int vq::scratch_reserve(vq_store *s, size_t dev_bytes, size_t host_bytes) {
    if (dev_bytes > s->lab_dev_cap) {
        if (s->lab_dev) cudaFree(s->lab_dev);
        s->lab_dev = nullptr;
        s->lab_dev_cap = 0;
        const size_t want = dev_bytes + dev_bytes / 2 + 4096;
        VQ_CUDA(cudaMalloc((void **)&s->lab_dev, want));
        s->lab_dev_cap = want;
    }
    if (host_bytes > s->lab_host_cap) {
        if (s->lab_host) cudaFreeHost(s->lab_host);
        s->lab_host = nullptr;
        s->lab_host_cap = 0;
        const size_t want = host_bytes + host_bytes / 2 + 4096;
        VQ_CUDA(cudaMallocHost((void **)&s->lab_host, want));
        s->lab_host_cap = want;
    }
    return 0;
}

extern "C" const char *vq_last_error(void) { return vq::g_err; }
extern "C" int vq_abi_version(void) { return VQ_ABI_VERSION; }

extern "C" int vq_device_count(int *count_out) {
    VQ_REQUIRE(count_out, "vq_device_count: null output");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        vq::set_error("cudaGetDeviceCount -> %s", cudaGetErrorString(e));
        *count_out = 0;
        return -2;
    }
    *count_out = n;
    return 0;
}

static void store_free(vq_store *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    cudaFree(s->rows);
    cudaFree(s->inv_counts);
    cudaFree(s->target);
    cudaFree(s->scores);
    cudaFree(s->sims);
    cudaFree(s->hist);
    cudaFree(s->chunk_counts);
    cudaFree(s->chunk_offsets);
    cudaFree(s->counts);
    if (s->counts_host) cudaFreeHost(s->counts_host);
    for (int i = 0; i < 3; ++i) {
        cudaFree(s->list_rows[i]);
        cudaFree(s->list_scores[i]);
    }
    cudaFree(s->cand_count);
    cudaFree(s->cand_keys);
    cudaFree(s->topk_scores);
    cudaFree(s->topk_rows);
    cudaFree(s->pack);
    if (s->pinned_stage) cudaFreeHost(s->pinned_stage);
    for (int i = 0; i < 3; ++i) {
        if (s->h_rows[i]) cudaFreeHost(s->h_rows[i]);
        if (s->h_scores[i]) cudaFreeHost(s->h_scores[i]);
    }
    if (s->batch_scratch && s->batch_scratch_free) s->batch_scratch_free(s->batch_scratch);
    s->batch_scratch = nullptr;
    if (s->h_result) cudaFreeHost(s->h_result);
    if (s->h_gather) cudaFreeHost(s->h_gather);
    if (s->lab_dev) cudaFree(s->lab_dev);
    if (s->lab_host) cudaFreeHost(s->lab_host);
    if (s->h_rank_rows) cudaFreeHost(s->h_rank_rows);
    if (s->h_rank_scores) cudaFreeHost(s->h_rank_scores);
    if (s->h_topk_rows) cudaFreeHost(s->h_topk_rows);
    if (s->h_topk_scores) cudaFreeHost(s->h_topk_scores);
    for (int i = 0; i < vq::kTimeRing; ++i)                  // created on first use of their ring slot (vq_scan_enqueue)
        if (s->ev_start[i]) {
            cudaEventDestroy(s->ev_start[i]);
            cudaEventDestroy(s->ev_stop[i]);
            cudaEventDestroy(s->ev_sel_stop[i]);
        }
    if (s->pack_reader_done) cudaEventDestroy(s->pack_reader_done);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

// Everything sized by the row capacity: the rows themselves and the per-row scratch of the scan.  Called at create
// time and by vq_store_reserve (which moves the rows over and swaps the scratch).
static int alloc_row_buffers(vq_store *s, int64_t capacity) {
    const size_t nr = (size_t)capacity;
    const size_t nc = (size_t)((capacity + vq::kChunkRows - 1) / vq::kChunkRows);
#define VQ_ALLOC(ptr, bytes)                                                              \
    do {                                                                                  \
        cudaError_t e2 = cudaMalloc((void **)&(ptr), (bytes));                            \
        if (e2 != cudaSuccess) {                                                          \
            (void)cudaGetLastError();                                                     \
            vq::set_error("vq_store: cudaMalloc(%zu bytes) for %s -> %s", (size_t)(bytes), #ptr, cudaGetErrorString(e2)); \
            return -3;                                                                    \
        }                                                                                 \
    } while (0)
    VQ_ALLOC(s->rows, nr * s->row_floats * sizeof(float));
    VQ_ALLOC(s->scores, nr * sizeof(float));
    VQ_ALLOC(s->chunk_counts, 3 * nc * sizeof(unsigned int));
    VQ_ALLOC(s->chunk_offsets, 3 * nc * sizeof(unsigned int));
    for (int i = 0; i < 3; ++i) {
        VQ_ALLOC(s->list_rows[i], nr * sizeof(uint32_t));
        VQ_ALLOC(s->list_scores[i], nr * sizeof(float));
    }
    VQ_ALLOC(s->cand_keys, nr * sizeof(unsigned long long));
#undef VQ_ALLOC
    s->cand_cap = (int64_t)nr;
    s->capacity = capacity;
    return 0;
}

extern "C" int vq_store_create(vq_store **out, int device, int64_t n_rows, int n_streams,
                               int n_splits, int dim, int64_t first_global_row) {
    VQ_REQUIRE(out, "vq_store_create: null output");
    *out = nullptr;
    VQ_REQUIRE(n_rows >= 0 && n_rows < (int64_t)0xFFFFFFF0u, "vq_store_create: n_rows %lld out of range",
               (long long)n_rows);
    VQ_REQUIRE(n_streams >= 1 && n_streams <= VQ_MAX_STREAMS, "vq_store_create: n_streams must be 1..%d",
               VQ_MAX_STREAMS);
    VQ_REQUIRE(n_splits >= 1 && dim >= 4 && dim % 4 == 0, "vq_store_create: dim must be a multiple of 4");
    int ndev = 0;
    VQ_CUDA(cudaGetDeviceCount(&ndev));
    VQ_REQUIRE(device >= 0 && device < ndev, "vq_store_create: device %d of %d", device, ndev);
    VQ_CUDA(cudaSetDevice(device));
    vq_store *s = new vq_store();
    s->device = device;
    s->n_rows = n_rows;
    s->first_global_row = first_global_row;
    s->n_streams = n_streams;
    s->n_splits = n_splits;
    s->dim = dim;
    s->stream_len = n_splits * dim;
    s->row_floats = (size_t)n_streams * s->stream_len;
    s->n_chunks = (n_rows + vq::kChunkRows - 1) / vq::kChunkRows;
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, device);
    s->sm_count = (e == cudaSuccess) ? prop.multiProcessorCount : 148;
    if (cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) != cudaSuccess) {
        vq::set_error("vq_store_create: cudaStreamCreate failed");
        store_free(s);
        return -2;
    }
    if (int r = alloc_row_buffers(s, n_rows > 0 ? n_rows : 1)) {
        store_free(s);
        return r;
    }
#define VQ_ALLOC(ptr, bytes)                                                              \
    do {                                                                                  \
        cudaError_t e2 = cudaMalloc((void **)&(ptr), (bytes));                            \
        if (e2 != cudaSuccess) {                                                          \
            (void)cudaGetLastError();                                                     \
            vq::set_error("vq_store_create: cudaMalloc(%zu bytes) for %s -> %s", (size_t)(bytes), #ptr, \
                          cudaGetErrorString(e2));                                        \
            store_free(s);                                                                \
            return -3;                                                                    \
        }                                                                                 \
    } while (0)
    VQ_ALLOC(s->target, s->row_floats * sizeof(float));
    VQ_ALLOC(s->hist, (vq::kHistBins + 8) * sizeof(unsigned int));
    VQ_ALLOC(s->counts, 8 * sizeof(int64_t));
    VQ_ALLOC(s->cand_count, 4 * sizeof(unsigned int));
    VQ_ALLOC(s->topk_scores, VQ_MAX_TOPK * sizeof(float));
    VQ_ALLOC(s->topk_rows, VQ_MAX_TOPK * sizeof(int64_t));
    VQ_ALLOC(s->pack, (4 + 2 * VQ_MAX_TOPK) * sizeof(int64_t));
#undef VQ_ALLOC
    if (cudaMallocHost((void **)&s->counts_host, 8 * sizeof(int64_t)) != cudaSuccess ||
        cudaMallocHost(&s->pinned_stage, s->row_floats * sizeof(double) + 4096) != cudaSuccess) {
        vq::set_error("vq_store_create: cudaMallocHost failed");
        store_free(s);
        return -3;
    }
    cudaMemsetAsync(s->counts, 0, 8 * sizeof(int64_t), s->stream);
    cudaMemsetAsync(s->hist, 0, (vq::kHistBins + 8) * sizeof(unsigned int), s->stream);
    cudaMemsetAsync(s->cand_count, 0, 4 * sizeof(unsigned int), s->stream);
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    *out = s;
    return 0;
}

extern "C" int vq_store_destroy(vq_store *s) {
    store_free(s);
    return 0;
}

extern "C" int vq_store_reserve(vq_store *s, int64_t capacity) {
    VQ_REQUIRE(s, "vq_store_reserve: null store");
    VQ_REQUIRE(capacity < (int64_t)0xFFFFFFF0u, "vq_store_reserve: capacity %lld out of range", (long long)capacity);
    if (capacity <= s->capacity) return 0;
    VQ_CUDA(cudaSetDevice(s->device));
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    // keep the old buffers until the new ones exist: a failed reservation leaves the store as it was
    vq_store old = *s;
    s->rows = nullptr; s->scores = nullptr; s->chunk_counts = nullptr; s->chunk_offsets = nullptr; s->cand_keys = nullptr;
    for (int i = 0; i < 3; ++i) { s->list_rows[i] = nullptr; s->list_scores[i] = nullptr; }
    if (int r = alloc_row_buffers(s, capacity)) {
        cudaFree(s->rows); cudaFree(s->scores); cudaFree(s->chunk_counts); cudaFree(s->chunk_offsets); cudaFree(s->cand_keys);
        for (int i = 0; i < 3; ++i) { cudaFree(s->list_rows[i]); cudaFree(s->list_scores[i]); }
        s->rows = old.rows; s->scores = old.scores; s->chunk_counts = old.chunk_counts; s->chunk_offsets = old.chunk_offsets;
        s->cand_keys = old.cand_keys; s->cand_cap = old.cand_cap; s->capacity = old.capacity;
        for (int i = 0; i < 3; ++i) { s->list_rows[i] = old.list_rows[i]; s->list_scores[i] = old.list_scores[i]; }
        return r;
    }
    if (s->n_rows > 0)
        VQ_CUDA(cudaMemcpyAsync(s->rows, old.rows, (size_t)s->n_rows * s->row_floats * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    cudaFree(old.rows); cudaFree(old.scores); cudaFree(old.chunk_counts); cudaFree(old.chunk_offsets); cudaFree(old.cand_keys);
    for (int i = 0; i < 3; ++i) { cudaFree(old.list_rows[i]); cudaFree(old.list_scores[i]); }
    if (s->sims) { cudaFree(s->sims); s->sims = nullptr; }      // re-allocated by the next want_sims scan
    s->staged = false;
    return 0;
}

extern "C" int vq_store_append(vq_store *s, int64_t n_new, const float *rows) {
    VQ_REQUIRE(s, "vq_store_append: null store");
    VQ_REQUIRE(n_new >= 0 && (rows || n_new == 0), "vq_store_append: bad argument");
    if (n_new == 0) return 0;
    const int64_t need = s->n_rows + n_new;
    if (need > s->capacity) {                           // grow geometrically: amortised O(1) copies per appended row
        int64_t cap = s->capacity + s->capacity / 2;
        if (cap < need) cap = need;
        if (int r = vq_store_reserve(s, cap)) return r;
    }
    const int64_t first = s->n_rows;
    s->n_rows = need;
    s->n_chunks = (s->n_rows + vq::kChunkRows - 1) / vq::kChunkRows;
    if (s->inv_counts) {                                // per-row split weights no longer cover the shard: the caller sets them again
        cudaFree(s->inv_counts);
        s->inv_counts = nullptr;
    }
    if (s->sims) { cudaFree(s->sims); s->sims = nullptr; }
    s->staged = false;
    return vq_store_upload(s, first, n_new, rows);
}

extern "C" int vq_store_describe(const vq_store *s, int64_t *n_rows, int *n_streams, int *n_splits,
                                 int *dim, int64_t *first_global_row, int *device) {
    VQ_REQUIRE(s, "vq_store_describe: null store");
    if (n_rows) *n_rows = s->n_rows;
    if (n_streams) *n_streams = s->n_streams;
    if (n_splits) *n_splits = s->n_splits;
    if (dim) *dim = s->dim;
    if (first_global_row) *first_global_row = s->first_global_row;
    if (device) *device = s->device;
    return 0;
}

static int check_range(const vq_store *s, int64_t first, int64_t n, const char *who) {
    VQ_REQUIRE(s, "%s: null store", who);
    VQ_REQUIRE(first >= 0 && n >= 0 && first + n <= s->n_rows, "%s: rows [%lld, %lld) outside shard of %lld",
               who, (long long)first, (long long)(first + n), (long long)s->n_rows);
    return 0;
}

extern "C" int vq_store_upload(vq_store *s, int64_t first_row, int64_t n_rows, const float *rows) {
    if (int r = check_range(s, first_row, n_rows, "vq_store_upload")) return r;
    VQ_REQUIRE(rows || n_rows == 0, "vq_store_upload: null rows");
    s->batch_absmax_valid = false;
    VQ_CUDA(cudaSetDevice(s->device));
    // Chunked so that pageable sources are staged in bounded pieces; pinned sources go at PCIe rate.
    const size_t row_bytes = s->row_floats * sizeof(float);
    const int64_t step = (int64_t)((size_t)(256u << 20) / row_bytes) + 1;
    for (int64_t r = 0; r < n_rows; r += step) {
        const int64_t m = (n_rows - r < step) ? (n_rows - r) : step;
        VQ_CUDA(cudaMemcpyAsync(s->rows + (size_t)(first_row + r) * s->row_floats,
                                rows + (size_t)r * s->row_floats, (size_t)m * row_bytes,
                                cudaMemcpyHostToDevice, s->stream));
    }
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

// Pipelined ingest: the caller fills a pinned chunk (vq_pinned_alloc) while the previous chunk is still in flight —
// enqueue only, one cudaMemcpyAsync per call; vq_store_sync waits for everything enqueued on the store's stream.
extern "C" int vq_store_upload_async(vq_store *s, int64_t first_row, int64_t n_rows, const float *rows_pinned) {
    if (int r = check_range(s, first_row, n_rows, "vq_store_upload_async")) return r;
    VQ_REQUIRE(rows_pinned || n_rows == 0, "vq_store_upload_async: null rows");
    if (n_rows == 0) return 0;
    s->batch_absmax_valid = false;
    VQ_CUDA(cudaSetDevice(s->device));
    VQ_CUDA(cudaMemcpyAsync(s->rows + (size_t)first_row * s->row_floats, rows_pinned, (size_t)n_rows * s->row_floats * sizeof(float),
                            cudaMemcpyHostToDevice, s->stream));
    return 0;
}

extern "C" int vq_store_sync(vq_store *s) {
    VQ_REQUIRE(s, "vq_store_sync: null store");
    VQ_CUDA(cudaSetDevice(s->device));
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int vq_pinned_alloc(void **out, int64_t bytes) {
    VQ_REQUIRE(out && bytes > 0, "vq_pinned_alloc: bad argument");
    *out = nullptr;
    VQ_CUDA(cudaMallocHost(out, (size_t)bytes));
    return 0;
}

extern "C" int vq_pinned_free(void *p) {
    if (p) cudaFreeHost(p);
    return 0;
}

extern "C" int vq_store_download(vq_store *s, int64_t first_row, int64_t n_rows, float *rows_out) {
    if (int r = check_range(s, first_row, n_rows, "vq_store_download")) return r;
    VQ_REQUIRE(rows_out || n_rows == 0, "vq_store_download: null output");
    VQ_CUDA(cudaSetDevice(s->device));
    VQ_CUDA(cudaMemcpyAsync(rows_out, s->rows + (size_t)first_row * s->row_floats,
                            (size_t)n_rows * s->row_floats * sizeof(float), cudaMemcpyDeviceToHost,
                            s->stream));
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int vq_store_set_split_weights(vq_store *s, const float *inv_counts) {
    VQ_REQUIRE(s, "vq_store_set_split_weights: null store");
    VQ_CUDA(cudaSetDevice(s->device));
    if (!inv_counts) {
        if (s->inv_counts) cudaFree(s->inv_counts);
        s->inv_counts = nullptr;
        return 0;
    }
    const size_t bytes = (size_t)(s->n_rows > 0 ? s->n_rows : 1) * s->n_streams * sizeof(float);
    if (!s->inv_counts) VQ_CUDA(cudaMalloc((void **)&s->inv_counts, bytes));
    VQ_CUDA(cudaMemcpyAsync(s->inv_counts, inv_counts, (size_t)s->n_rows * s->n_streams * sizeof(float),
                            cudaMemcpyHostToDevice, s->stream));
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    return 0;
}

extern "C" int vq_store_device_ptr(vq_store *s, void **rows_dev) {
    VQ_REQUIRE(s && rows_dev, "vq_store_device_ptr: null argument");
    *rows_dev = s->rows;
    return 0;
}

// ------------------------------------------------------------------------------------ VQSYN-1
namespace {

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(unsigned long long k, unsigned long long seed) {
    const unsigned long long z = mix64(k * 0x9E3779B97F4A7C15ull + seed);
    return __fmul_rn((float)(unsigned int)(z >> 40), 5.9604644775390625e-08f);   // * 2^-24, exact
}

__global__ void synth_base_kernel(float *base, int n_streams, int stream_len, unsigned long long seed,
                                  const float *means3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_streams * stream_len) return;
    const int s = i / stream_len, d = i - s * stream_len;
    const unsigned long long kb = ((1ull << 48) + (unsigned long long)s) * (unsigned long long)stream_len + d;
    const float xb = u01(kb, seed);
    base[i] = __fmul_rn(__fmul_rn(xb, xb), means3[s]);
}

// One thread per float4 of the shard.  All float ops single-rounded (no FMA contraction), so the
// numpy twin reproduces every bit.
__global__ void synth_fill_kernel(float4 *rows, const float *__restrict__ base, long long n_rows,
                                  int n_streams, int stream_len, long long first_global_row,
                                  unsigned long long seed, const float *__restrict__ means3) {
    const int vec_per_row = n_streams * stream_len / 4;
    const long long total = n_rows * vec_per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / vec_per_row;
        const int v = (int)(i - r * vec_per_row);
        const int s = (v * 4) / stream_len;
        const int d0 = v * 4 - s * stream_len;
        const unsigned long long c = (unsigned long long)(first_global_row + r);
        const float a = u01((1ull << 56) + c, seed);
        const float a2 = __fmul_rn(a, a), a4 = __fmul_rn(a2, a2), alpha = __fmul_rn(a4, a4);
        const float beta = __fsub_rn(1.0f, alpha);
        const float m3 = means3[s];
        const unsigned long long k0 =
            (c * (unsigned long long)n_streams + (unsigned long long)s) * (unsigned long long)stream_len + d0;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float x = u01(k0 + j, seed);
            const float noise = __fmul_rn(__fmul_rn(x, x), m3);
            o[j] = __fadd_rn(__fmul_rn(alpha, base[s * stream_len + d0 + j]), __fmul_rn(beta, noise));
        }
        rows[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
}
}  // namespace

extern "C" int vq_store_fill_synthetic(vq_store *s, uint64_t seed, const float *stream_means) {
    VQ_REQUIRE(s, "vq_store_fill_synthetic: null store");
    s->batch_absmax_valid = false;
    static const float kDefaultMeans[VQ_MAX_STREAMS] = {2.5f, 0.9f, 1.7f, 1.3f};
    VQ_CUDA(cudaSetDevice(s->device));
    float means3[VQ_MAX_STREAMS];
    for (int i = 0; i < s->n_streams; ++i)
        means3[i] = 3.0f * (stream_means ? stream_means[i] : kDefaultMeans[i]);   // one fp32 rounding
    float *d_means = nullptr, *d_base = nullptr;
    VQ_CUDA(cudaMalloc((void **)&d_means, sizeof(means3)));
    VQ_CUDA(cudaMalloc((void **)&d_base, s->row_floats * sizeof(float)));
    VQ_CUDA(cudaMemcpyAsync(d_means, means3, sizeof(means3), cudaMemcpyHostToDevice, s->stream));
    const int nb = (int)((s->row_floats + 255) / 256);
    synth_base_kernel<<<nb, 256, 0, s->stream>>>(d_base, s->n_streams, s->stream_len, seed, d_means);
    if (s->n_rows > 0)
        synth_fill_kernel<<<s->sm_count * 16, 256, 0, s->stream>>>(
            reinterpret_cast<float4 *>(s->rows), d_base, s->n_rows, s->n_streams, s->stream_len,
            s->first_global_row, seed, d_means);
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(s->stream);
    cudaFree(d_means);
    cudaFree(d_base);
    VQ_CUDA(e);
    VQ_CUDA(e2);
    return 0;
}
