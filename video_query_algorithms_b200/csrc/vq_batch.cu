// K3 — batched queries on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a: host side.
//
// Q targets are scored against every clip of the shard in ONE pass over HBM per 256 queries: per stream the
// similarities are a dense contraction SIM_s[clip, query] = sum_d X[clip, s, d] * T[query, s, d] (reference
// ticket.py:146-160 for Q tickets at once), then the per-(clip, query) score of ticket.py:173-180 and the candidate
// tests of ticket.py:325-327 are applied in the kernel's epilogue.  No Q x N score matrix is ever written: per query
// the kernel keeps the match / near-miss counts and a candidate list for the exact top-k (ranking rule of
// ticket.py:266).  The kernel itself is in vq_batch_bf16.cuh; this file holds the per-chunk top-k compaction
// kernels, the tensor maps and the launch loop.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "vq_internal.cuh"
#include "vq_topk.cuh"
#include "vq_tc.cuh"

namespace {

using namespace vqtc;

#ifdef VQ_BATCH_PROFILE          // role cycle counters (development builds only)
#define VQ_CLOCK() clock64()
#else
#define VQ_CLOCK() 0ll
#endif

struct BatchArgs {
    float w[VQ_MAX_STREAMS];
    float inv_den, inv_splits;
    float th_f, lo_f;              // smallest floats >= the double thresholds: (double)s >= th  <=>  s >= th_f
    int n_queries;
    int kb_per_stream;             // stream_len / 32
    int n_streams;
    long long row0, n_rows_total;  // chunk start (local rows) and shard size
    long long row_end;             // first row past this launch's chunk
    int n_tiles;                   // tiles in this chunk
    int n_mma;                     // queries of this pass rounded up to 16 (the UMMA N)
    long long cand_cap;
    // tie band: closed fp32 intervals holding exactly the floats s with |(double)s - th| < eps (0) and |(double)s - lo| < eps (1)
    float tie_lo0, tie_hi0, tie_lo1, tie_hi1;
    long long tie_cap;
#ifdef VQ_BATCH_F16F8
    // fp16 + fp8 operand split (vq_batch_bf16.cuh): per stream, sx = bx * 2^6 scales the clip rows into fp16 range and
    // descale = 1 / (4096 bx bt) undoes both operand scales on the stream's sum (powers of two: exact).
    // (Only in that build: 32 more bytes of kernel parameters cost the default kernel 40 bytes of extra spills.)
    float sx[VQ_MAX_STREAMS], descale[VQ_MAX_STREAMS];
#endif
};

#include "vq_batch_bf16.cuh"

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
               uint32_t box_inner = 32, bool swizzle = true) {
    return encode_map_ex(map, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, sizeof(float), inner, rows, box_rows, box_inner,
                         swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
}

float float_ceil_of(double x) {
    // smallest float f with (double)f >= x, so that for every float s: (double)s >= x  <=>  s >= f
    float f = (float)x;
    if ((double)f < x) f = nextafterf(f, INFINITY);
    return f;
}

// [lo, hi] = the floats f with fabs((double)f - c) < eps (the test of vq_scan.cu classify()); empty (lo > hi) when eps <= 0
void tie_interval(double c, double eps, float *lo, float *hi) {
    *lo = INFINITY;
    *hi = -INFINITY;
    if (!(eps > 0.0) || !(c == c) || fabs(c) > 1e30) return;
    float l = (float)(c - eps), h = (float)(c + eps);
    for (int i = 0; i < 8 && !(fabs((double)l - c) < eps); ++i) l = nextafterf(l, INFINITY);
    for (int i = 0; i < 8 && !(fabs((double)h - c) < eps); ++i) h = nextafterf(h, -INFINITY);
    if (fabs((double)l - c) < eps && fabs((double)h - c) < eps && l <= h) {
        *lo = l;
        *hi = h;
    }
}

// max |x| per stream over the shard (bit pattern of a non-negative float orders like the float); NaNs are skipped
__global__ void absmax_kernel(const float4 *__restrict__ rows, long long n_vec, int vec_per_row, int vec_per_stream, unsigned int *out) {
    __shared__ unsigned int red[VQ_MAX_STREAMS];
    if (threadIdx.x < VQ_MAX_STREAMS) red[threadIdx.x] = 0u;
    __syncthreads();
    float m[VQ_MAX_STREAMS] = {0.f, 0.f, 0.f, 0.f};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = rows[i];
        const int s = (int)(i % vec_per_row) / vec_per_stream;
        float a = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));      // fmaxf drops NaNs
        if (!(a < INFINITY)) a = 0.f;
#pragma unroll
        for (int q = 0; q < VQ_MAX_STREAMS; ++q) m[q] = (q == s) ? fmaxf(m[q], a) : m[q];
    }
#pragma unroll
    for (int q = 0; q < VQ_MAX_STREAMS; ++q) {
        float v = m[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((threadIdx.x & 31) == 0 && v > 0.f) atomicMax(&red[q], __float_as_uint(v));
    }
    __syncthreads();
    if (threadIdx.x < VQ_MAX_STREAMS && red[threadIdx.x]) atomicMax(&out[threadIdx.x], red[threadIdx.x]);
}

// largest power of two b with mx * b <= top (1 when mx is 0 or not finite)
float pow2_scale(float mx, float top) {
    if (!(mx > 0.f) || !(mx < INFINITY)) return 1.f;
    int e;
    frexpf(top / mx, &e);                       // top / mx = f * 2^e, 0.5 <= f < 1  ->  2^(e-1) <= top / mx
    return ldexpf(1.f, e - 1);
}

constexpr long long kTieCap = 4096;      // tie-band entries kept per query (the band is 2 * COMPUTE_EPS wide: ~1e-5 of the rows)

struct Dev {
    void *p = nullptr;
    ~Dev() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t b) { return cudaMalloc(&p, b ? b : 8); }
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

// Scratch of the batched path, owned by the store: allocated on the first batched scan, reused by every later one.
struct BatchScratch {
    Dev t, t1, t2, cut, counts, cnt, keys, rows, sc, park, tie_cnt, tie_keys, amax;
    void *pinned_t = nullptr;            // pinned staging for one pass of targets
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~BatchScratch() {
        if (pinned_t) cudaFreeHost(pinned_t);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
    }
};
void free_batch_scratch(void *p) { delete static_cast<BatchScratch *>(p); }

int alloc_batch_scratch(const vq_store *s, size_t K, long long cap, BatchScratch *b) {
    constexpr int QN = bf::QN;
    VQ_CUDA(b->t.alloc((size_t)QN * K * 4));
    VQ_CUDA(b->t1.alloc((size_t)QN * K * 2));
    VQ_CUDA(b->t2.alloc((size_t)QN * K * 2));
    VQ_CUDA(b->cut.alloc(QN * 4));
    VQ_CUDA(b->counts.alloc(QN * 2 * 8));
    VQ_CUDA(b->cnt.alloc(QN * 4));
    VQ_CUDA(b->keys.alloc((size_t)QN * cap * 8));
    VQ_CUDA(b->rows.alloc((size_t)QN * VQ_MAX_TOPK * 8));
    VQ_CUDA(b->sc.alloc((size_t)QN * VQ_MAX_TOPK * 4));
    VQ_CUDA(b->park.alloc((size_t)s->sm_count * QN * bf::BM * 4));
    VQ_CUDA(b->tie_cnt.alloc(QN * 4));
    VQ_CUDA(b->amax.alloc(VQ_MAX_STREAMS * 4));
    VQ_CUDA(b->tie_keys.alloc((size_t)QN * kTieCap * 8));
    VQ_CUDA(cudaMallocHost(&b->pinned_t, (size_t)QN * K * 4));
    VQ_CUDA(cudaEventCreate(&b->e0));
    VQ_CUDA(cudaEventCreate(&b->e1));
    return 0;
}

int get_batch_scratch(vq_store *s, size_t K, long long cap, BatchScratch **out) {
    if (!s->batch_scratch) {
        BatchScratch *b = new BatchScratch();
        if (int r = alloc_batch_scratch(s, K, cap, b)) {   // all or nothing: a half-built scratch is never kept
            delete b;
            return r;
        }
        s->batch_scratch = b;
        s->batch_scratch_free = free_batch_scratch;
    }
    *out = static_cast<BatchScratch *>(s->batch_scratch);
    return 0;
}

// The batched path: bf16x2 kernel (vq_batch_bf16.cuh), 256 queries per pass over the shard.
int run_batch_bf16(vq_store *s, const float *targets, int n_queries, const vq_scan_params *p, int64_t *counts_out,
                   int64_t *topk_rows_out, float *topk_scores_out, float *kernel_ms_out, float *scores_dbg_host,
                   int64_t *tie_counts_out = nullptr, int32_t tie_cap_out = 0, int64_t *tie_rows_out = nullptr,
                   float *tie_scores_out = nullptr) {
    VQ_REQUIRE(s && targets && p, "vq_scan_batch: null argument");
    VQ_REQUIRE(n_queries >= 1, "vq_scan_batch: need at least one query");
    VQ_REQUIRE(s->stream_len % bf::BK == 0, "vq_scan_batch: stream length %d is not a multiple of %d", s->stream_len, bf::BK);
    VQ_REQUIRE(s->stream_len % 16 == 0, "vq_scan_batch: stream length %d is not a multiple of 16", s->stream_len);
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
    // 19k clips are all candidates; then 16 tiles per CTA per launch
    const long long first_rows = (long long)s->sm_count * bf::BM;
    const long long chunk_rows = (long long)s->sm_count * bf::BM * 16;
    const long long cap = chunk_rows + VQ_MAX_TOPK;
    VQ_CUDA(cudaFuncSetAttribute(bf::batch_scan_bf16<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf::Ring<256>::SMEM));
    VQ_CUDA(cudaFuncSetAttribute(bf::batch_scan_bf16<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf::Ring<128>::SMEM));
    VQ_CUDA(cudaFuncSetAttribute(bf::batch_scan_bf16<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf::Ring<256>::SMEM));
    VQ_CUDA(cudaFuncSetAttribute(bf::batch_scan_bf16<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bf::Ring<128>::SMEM));
    const bool want_ties = tie_counts_out != nullptr && p->eps > 0.0;
    VQ_REQUIRE(!tie_rows_out || (tie_scores_out && tie_cap_out > 0), "vq_scan_batch_ties: tie lists need scores and a capacity");
    BatchScratch *bs = nullptr;
    if (int r = get_batch_scratch(s, K, cap, &bs)) return r;
    Dev &d_t = bs->t, &d_t1 = bs->t1, &d_t2 = bs->t2, &d_cut = bs->cut, &d_counts = bs->counts, &d_cnt = bs->cnt,
        &d_keys = bs->keys, &d_rows = bs->rows, &d_sc = bs->sc, &d_park = bs->park;
    Dev d_dbg, d_prof;
    if (scores_dbg_host) VQ_CUDA(d_dbg.alloc((size_t)QN * s->n_rows * 4));
    const bool want_prof = getenv("VQ_BATCH_PROF") != nullptr;
    if (want_prof) VQ_CUDA(d_prof.alloc((size_t)s->sm_count * 16 * 8));
    cudaEvent_t e0 = bs->e0, e1 = bs->e1;
#ifdef VQ_BATCH_WATCHDOG          // development build (vq_tc.cuh): stuck barrier waits trap and are reported here
    static unsigned long long *wd_host = nullptr;
    if (!wd_host) {
        VQ_CUDA(cudaMallocHost((void **)&wd_host, 256 * 16 * 2 * 8));
        VQ_CUDA(cudaMemcpyToSymbol(vqtc::vq_watchdog_host, &wd_host, sizeof(wd_host)));
    }
    memset(wd_host, 0, 256 * 16 * 2 * 8);
#endif
#ifdef VQ_BATCH_F16F8
    // clip-row scale per stream: max |x| of the shard, taken once per content version (upload / append / fill reset it)
    if (!s->batch_absmax_valid && s->n_rows > 0) {
        VQ_CUDA(cudaMemsetAsync(bs->amax.p, 0, VQ_MAX_STREAMS * 4, st));
        const long long n_vec = (long long)s->n_rows * (long long)(K / 4);
        absmax_kernel<<<s->sm_count * 8, 256, 0, st>>>(reinterpret_cast<const float4 *>(s->rows), n_vec, (int)(K / 4),
                                                      s->stream_len / 4, bs->amax.as<unsigned int>());
        VQ_CUDA(cudaGetLastError());
        unsigned int bits[VQ_MAX_STREAMS];
        VQ_CUDA(cudaMemcpyAsync(bits, bs->amax.p, sizeof(bits), cudaMemcpyDeviceToHost, st));
        VQ_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < VQ_MAX_STREAMS; ++i) memcpy(&s->batch_absmax[i], &bits[i], 4);
        s->batch_absmax_valid = true;
    }
#endif
    float total_ms = 0.f;
    int rc = 0;
    for (int q0 = 0; q0 < n_queries && rc == 0; q0 += QN) {
        const int nq = (n_queries - q0 < QN) ? (n_queries - q0) : QN;
        if (nq < QN) VQ_CUDA(cudaMemsetAsync(d_t.p, 0, (size_t)QN * K * 4, st));
        memcpy(bs->pinned_t, targets + (size_t)q0 * K, (size_t)nq * K * 4);       // caller memory may be pageable
        VQ_CUDA(cudaMemcpyAsync(d_t.p, bs->pinned_t, (size_t)nq * K * 4, cudaMemcpyHostToDevice, st));
#ifdef VQ_BATCH_F16F8
        float sx[VQ_MAX_STREAMS] = {1.f, 1.f, 1.f, 1.f}, stq[VQ_MAX_STREAMS] = {1.f, 1.f, 1.f, 1.f}, descale[VQ_MAX_STREAMS] = {1.f, 1.f, 1.f, 1.f};
#endif
#ifndef VQ_BATCH_F16F8
        bf::split_targets_bf16<<<(unsigned)(((size_t)QN * K + 255) / 256), 256, 0, st>>>(
            d_t.as<float>(), d_t1.as<unsigned short>(), d_t2.as<unsigned short>(), (long long)QN * K);
#else
        for (int si = 0; si < s->n_streams; ++si) {          // query scale per stream: max |t| over this pass's targets
            float mx = 0.f;
            for (int q = 0; q < nq; ++q) {
                const float *tp = targets + (size_t)(q0 + q) * K + (size_t)si * s->stream_len;
                for (int d = 0; d < s->stream_len; ++d) {
                    const float v = fabsf(tp[d]);
                    if (v > mx && v < INFINITY) mx = v;
                }
            }
            const float bx = pow2_scale(s->batch_absmax[si], 128.f), bt = pow2_scale(mx, 128.f);
            sx[si] = bx * 64.f;
            stq[si] = bt * 64.f;
            descale[si] = 1.f / (4096.f * bx * bt);
            VQ_REQUIRE(descale[si] > 0.f && descale[si] < INFINITY && sx[si] < INFINITY && stq[si] < INFINITY,
                       "vq_scan_batch: operand scales out of range (stream %d: max |x| %g, max |t| %g)", si, (double)s->batch_absmax[si], (double)mx);
        }
        bf::split_targets_f16f8<<<(unsigned)(((size_t)QN * K / 16 + 255) / 256), 256, 0, st>>>(
            d_t.as<float>(), d_t1.as<unsigned short>(), d_t2.as<unsigned char>(), (long long)QN * (long long)K / 16, s->stream_len,
            s->n_streams, stq[0], stq[1], stq[2], stq[3]);
#endif
        fill_f32<<<1, QN, 0, st>>>(d_cut.as<float>(), topk > 0 ? -INFINITY : INFINITY, QN);   // no top-k: nothing is a candidate
        VQ_CUDA(cudaMemsetAsync(d_counts.p, 0, QN * 2 * 8, st));
        VQ_CUDA(cudaMemsetAsync(d_cnt.p, 0, QN * 4, st));
        VQ_CUDA(cudaMemsetAsync(bs->tie_cnt.p, 0, QN * 4, st));
        CUtensorMap map_a, map_t1, map_t2;
        if (s->n_rows > 0) {
            if ((rc = encode_map(&map_a, s->rows, K, (uint64_t)s->n_rows, bf::BM))) break;
            // only the query rows in use are fetched per K block (the kernel is bound by the L2 -> SM fill rate)
            const int n_mma_q = ((nq + 15) / 16) * 16;
            const uint32_t t_rows = n_mma_q > 128 ? 256u : (n_mma_q > 64 ? 128u : 64u);
            if ((rc = encode_map_ex(&map_t1, d_t1.p, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, QN, t_rows, bf::BK, CU_TENSOR_MAP_SWIZZLE_64B))) break;
            if ((rc = encode_map_ex(&map_t2, d_t2.p, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, K, QN, t_rows, bf::BK, CU_TENSOR_MAP_SWIZZLE_64B))) break;
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
#ifdef VQ_BATCH_F16F8
        for (int i = 0; i < VQ_MAX_STREAMS; ++i) { a.sx[i] = sx[i]; a.descale[i] = descale[i]; }
#endif
        tie_interval(p->threshold, want_ties ? p->eps : 0.0, &a.tie_lo0, &a.tie_hi0);
        tie_interval(p->lower_limit, want_ties ? p->eps : 0.0, &a.tie_lo1, &a.tie_hi1);
        a.tie_cap = kTieCap;
        VQ_CUDA(cudaEventRecord(e0, st));
        // launch sizes 1, 4, 16, 16, ... tiles per CTA: a launch appends ~(its clips) * k / (clips seen before) candidates per
        // query, so growing the launches by 4x keeps that near 4k instead of 16k for the first full-size launch (ncu: the
        // 16-tile launch right after a 1-tile seed ran at 1143 us against 850 us in steady state)
        for (long long r0 = 0, step = first_rows, li = 0; r0 < s->n_rows; r0 += step, ++li, step = (li == 1 ? 4 * first_rows : chunk_rows)) {
            const long long nr = (s->n_rows - r0 < step) ? (s->n_rows - r0) : step;
            a.row0 = r0;
            a.n_tiles = (int)((nr + bf::BM - 1) / bf::BM);
            a.row_end = r0 + nr;
            const int units = a.n_tiles < s->sm_count ? a.n_tiles : s->sm_count;
            auto kernel = a.n_mma > 128 ? (want_ties ? bf::batch_scan_bf16<256, true> : bf::batch_scan_bf16<256, false>)
                                        : (want_ties ? bf::batch_scan_bf16<128, true> : bf::batch_scan_bf16<128, false>);
            kernel<<<units, bf::THREADS, a.n_mma > 128 ? bf::Ring<256>::SMEM : bf::Ring<128>::SMEM, st>>>(
                map_a, map_t1, map_t2, a, s->inv_counts, d_cut.as<float>(), d_counts.as<unsigned long long>(),
                d_cnt.as<unsigned int>(), d_keys.as<unsigned long long>(), bs->tie_cnt.as<unsigned int>(),
                bs->tie_keys.as<unsigned long long>(), d_park.as<float>(),
                scores_dbg_host ? d_dbg.as<float>() : nullptr, want_prof ? d_prof.as<long long>() : nullptr);
            if (topk > 0)
                batch_compact<<<QN, 1024, 0, st>>>(d_cnt.as<unsigned int>(), d_keys.as<unsigned long long>(), cap, topk,
                                                   d_cut.as<float>());
            else
                VQ_CUDA(cudaMemsetAsync(d_cnt.p, 0, QN * 4, st));
        }
        VQ_CUDA(cudaEventRecord(e1, st));
#ifdef VQ_BATCH_WATCHDOG
        if (cudaStreamSynchronize(st) != cudaSuccess) {
            for (int b = 0; b < 256; ++b)
                for (int w = 0; w < 16; ++w) {
                    const unsigned long long v = wd_host[(size_t)(b * 16 + w) * 2];
                    if (v & 1ull)
                        fprintf(stderr, "[K3 watchdog] block %d warp %d stuck on barrier 0x%llx (index %llu) parity %llu\n", b, w, v >> 32,
                                ((v >> 32) & 0x3ffull) >> 3, (v >> 1) & 1ull);
                }
            vq::set_error("vq_scan_batch: watchdog trap (a barrier wait lasted more than 3e9 cycles)");
            return -2;
        }
#endif
        if (want_prof) {
            std::vector<long long> h((size_t)s->sm_count * 16);
            VQ_CUDA(cudaMemcpyAsync(h.data(), d_prof.p, h.size() * 8, cudaMemcpyDeviceToHost, st));
            VQ_CUDA(cudaStreamSynchronize(st));
            static const char *names[13] = {"mma thread total", "mma wait x_full", "mma wait part_empty", "mma wait t_full",
                                            "converter wait xt_empty", "epilogue wait part_full", "epilogue drains", "converter wait a_full",
                                            "epilogue finals (park st)", "epilogue finals (park ld)", "epilogue scoring", "tiles", "mma thread total, ns"};
            for (int c = 0; c < 13; ++c) fprintf(stderr, "[K3 bf16 prof, last chunk, CTA 0] %-28s %12lld\n", names[c], h[c]);
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
        if (tie_counts_out) {
            // tie band: counts always; the lists (global rows ascending = database order) up to the caller's capacity
            std::vector<unsigned int> hc(QN, 0u);
            if (want_ties) VQ_CUDA(cudaMemcpy(hc.data(), bs->tie_cnt.p, QN * 4, cudaMemcpyDeviceToHost));
            std::vector<unsigned long long> hk;
            for (int q = 0; q < nq; ++q) {
                tie_counts_out[q0 + q] = (int64_t)hc[(size_t)q];
                if (!tie_rows_out) continue;
                const long long have = hc[(size_t)q] < (unsigned long long)kTieCap ? (long long)hc[(size_t)q] : kTieCap;
                hk.resize((size_t)have);
                if (have)
                    VQ_CUDA(cudaMemcpy(hk.data(), bs->tie_keys.as<unsigned long long>() + (size_t)q * kTieCap, (size_t)have * 8,
                                       cudaMemcpyDeviceToHost));
                std::sort(hk.begin(), hk.end(), [](unsigned long long x, unsigned long long y) { return (x & 0xFFFFFFFFull) > (y & 0xFFFFFFFFull); });
                for (int i = 0; i < tie_cap_out; ++i) {
                    const size_t at = (size_t)(q0 + q) * tie_cap_out + i;
                    if (i < have) {
                        const unsigned long long key = hk[(size_t)i];
                        unsigned int u = (unsigned int)(key >> 32);
                        u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
                        float f;
                        memcpy(&f, &u, 4);
                        tie_rows_out[at] = s->first_global_row + (int64_t)(0xFFFFFFFFu - (unsigned int)(key & 0xFFFFFFFFull));
                        tie_scores_out[at] = f;
                    } else {
                        tie_rows_out[at] = -1;
                        tie_scores_out[at] = -INFINITY;
                    }
                }
            }
        }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) total_ms += ms;
    }
    if (kernel_ms_out) *kernel_ms_out = total_ms;
    return rc;
}

}  // namespace

extern "C" int vq_scan_batch(vq_store *s, const float *targets, int32_t n_queries, const vq_scan_params *p,
                             int64_t *counts_out, int64_t *topk_rows_out, float *topk_scores_out,
                             float *kernel_ms_out) {
    return run_batch_bf16(s, targets, n_queries, p, counts_out, topk_rows_out, topk_scores_out, kernel_ms_out, nullptr);
}

// The same with the tie band of every query (north_star: "ties within COMPUTE_EPS of the threshold ... are reported"):
// rows whose fp32 score is within p->eps of the threshold or of the near-miss limit, compared in double like the
// single-query scan.  tie_counts_out [Q]; tie_rows_out / tie_scores_out [Q][tie_cap] in database order, -1 / -inf padded
// (may be NULL: counts only).  At most 4096 entries per query are kept on the device; the counts are exact.
extern "C" int vq_scan_batch_ties(vq_store *s, const float *targets, int32_t n_queries, const vq_scan_params *p,
                                  int64_t *counts_out, int64_t *topk_rows_out, float *topk_scores_out, int64_t *tie_counts_out,
                                  int32_t tie_cap, int64_t *tie_rows_out, float *tie_scores_out, float *kernel_ms_out) {
    VQ_REQUIRE(tie_counts_out, "vq_scan_batch_ties: null tie_counts_out");
    return run_batch_bf16(s, targets, n_queries, p, counts_out, topk_rows_out, topk_scores_out, kernel_ms_out, nullptr,
                          tie_counts_out, tie_cap, tie_rows_out, tie_scores_out);
}

extern "C" int vq_scan_batch_scores(vq_store *s, const float *targets, int32_t n_queries, const vq_scan_params *p,
                                    float *scores_out) {
    VQ_REQUIRE(scores_out, "vq_scan_batch_scores: null output");
    VQ_REQUIRE(s && (long long)s->n_rows * 256 <= (1ll << 28), "vq_scan_batch_scores: debug dump limited to 1M rows x 256 queries");
    vq_scan_params q = *p;
    q.topk = 0;
    return run_batch_bf16(s, targets, n_queries, &q, nullptr, nullptr, nullptr, nullptr, scores_out);
}
