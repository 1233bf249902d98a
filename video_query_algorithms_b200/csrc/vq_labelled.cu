// Labelled-subset path, all float64: the reference does every step on the user-labelled clips in
// float64 and the small linear solves are ill-conditioned enough (cond up to ~1e4) that fp32
// arithmetic breaks the 1e-5 bar.  Inputs are the fp32 store rows, widened.
//
//   K4  labelled_sims     fp64 similarities of a row list        (inputs of hyperparameter.py:57-65)
//   K5  loss_grid         R replicates x 40 weights x 31 thresholds hinge loss (hyperparameter.py:56-65)
//   K6  gram / solve / combine   new target from labelled rows   (target_clip.py:161-261)
#include <vector>

#include <chrono>
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "vq_internal.cuh"

namespace {

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per (labelled row, stream); per-split dots summed in split order, then / n_splits
__global__ void labelled_sims_kernel(const float *__restrict__ rows, const double *__restrict__ target,
                                     const long long *__restrict__ row_ids, const float *__restrict__ inv_counts,
                                     long long n, int n_streams, int n_splits, int dim, double *sims) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n * n_streams) return;
    const long long i = w / n_streams;
    const int s = (int)(w - i * n_streams);
    const long long r = row_ids[i];
    const size_t stream_len = (size_t)n_splits * dim;
    const float *x = rows + ((size_t)r * n_streams + s) * stream_len;
    const double *t = target + (size_t)s * stream_len;
    double total = 0.0;
    for (int p = 0; p < n_splits; ++p) {
        double acc = 0.0;
        for (int d = lane; d < dim; d += 32) acc = fma((double)x[p * dim + d], t[p * dim + d], acc);
        total += warp_sum_d(acc);
    }
    if (lane == 0) {
        const double cnt = inv_counts ? (double)__float2int_rn(1.0f / inv_counts[r * n_streams + s]) : (double)n_splits;
        sims[i * n_streams + s] = total / cnt;
    }
}

// score table [n_w][L]: 1 - sqrt(((1 - s0))^2 + (w (1 - s1))^2) / (1 + w^2)) in the reference's
// operation order (ticket.py:174-180 with weights {1.0, w}); no FMA contraction.
__global__ void score_table_kernel(const double *__restrict__ sims, long long L, const double *__restrict__ wgrid,
                                   int n_w, double *table) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L * n_w) return;
    const int iw = (int)(i / L);
    const long long l = i - (long long)iw * L;
    const double w = wgrid[iw];
    const double d0 = __dmul_rn(1.0, __dsub_rn(1.0, sims[2 * l]));
    const double d1 = __dmul_rn(w, __dsub_rn(1.0, sims[2 * l + 1]));
    const double ssum = __dadd_rn(__dadd_rn(0.0, __dmul_rn(d0, d0)), __dmul_rn(d1, d1));
    const double den = __dadd_rn(__dadd_rn(0.0, 1.0), __dmul_rn(w, w));
    table[i] = __dsub_rn(1.0, sqrt(__ddiv_rn(ssum, den)));
}

constexpr int kLossThreads = 128;
constexpr int kMaxTh = 32;

// one block per (replicate, weight, group of up to 32 thresholds): one running sum per threshold and thread (31 for the
// reference's grid, hyperparameter.py:21), block tree reduction; longer threshold grids take more groups (blockIdx.y)
__global__ void __launch_bounds__(kLossThreads)
loss_grid_kernel(const double *__restrict__ table, const unsigned char *__restrict__ labels, long long L,
                 const double *__restrict__ thgrid, int n_th_all, int n_w, double ballast,
                 const int *__restrict__ rep_offset, const int *__restrict__ rep_index, double *losses) {
    __shared__ double th_s[kMaxTh];
    __shared__ double red[kLossThreads / 32][kMaxTh];
    const int r = blockIdx.x / n_w, iw = blockIdx.x - r * n_w;
    const int th0 = blockIdx.y * kMaxTh;
    const int n_th = n_th_all - th0 < kMaxTh ? n_th_all - th0 : kMaxTh;
    if (threadIdx.x < n_th) th_s[threadIdx.x] = thgrid[th0 + threadIdx.x];
    __syncthreads();
    const int lo = rep_offset[r], hi = rep_offset[r + 1];
    double acc[kMaxTh];
#pragma unroll
    for (int j = 0; j < kMaxTh; ++j) acc[j] = 0.0;
    const double *row = table + (size_t)iw * L;
    for (int q = lo + threadIdx.x; q < hi; q += kLossThreads) {
        const int l = rep_index[q];
        const double sc = row[l];
        const double y = labels[l] ? 1.0 : 0.0;
        const double wgt = 1.0 + y * ballast;
#pragma unroll
        for (int j = 0; j < kMaxTh; ++j) {
            if (j < n_th) {
                const double d = sc - th_s[j];
                const double h = (d >= 0.0) ? 1.0 : 0.0;        // np.heaviside(d, 1)
                acc[j] += (h - y) * d * wgt;
            }
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < kMaxTh; ++j) {
        const double v = warp_sum_d(acc[j]);
        if (lane == 0) red[wid][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < n_th) {
        double v = 0.0;
        for (int w = 0; w < kLossThreads / 32; ++w) v += red[w][threadIdx.x];
        const double n = (double)(hi - lo);
        losses[((size_t)r * n_w + iw) * n_th_all + th0 + threadIdx.x] = (0.5 * th_s[threadIdx.x] + v) / n;
    }
}

// ------------------------------------------------------------------ target bootstrap (K6)
// Z = [valid rows; invalid rows] (n + m rows).  Gram matrices per (stream, split) slot:
// G[slot][i][j] = z_i . z_j over that slot's `dim` floats.  One warp per (slot, i <= j).
__global__ void gram_kernel(const float *__restrict__ rows, const long long *__restrict__ ids, int nz,
                            int n_slots, int dim, size_t row_floats, double *G) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long pairs = (long long)nz * nz;
    if (w >= pairs * n_slots) return;
    const int slot = (int)(w / pairs);
    const long long ij = w - (long long)slot * pairs;
    const int i = (int)(ij / nz), j = (int)(ij - (long long)i * nz);
    if (j < i) return;
    const float *a = rows + (size_t)ids[i] * row_floats + (size_t)slot * dim;
    const float *b = rows + (size_t)ids[j] * row_floats + (size_t)slot * dim;
    double acc = 0.0;
    for (int d = lane; d < dim; d += 32) acc = fma((double)a[d], (double)b[d], acc);
    acc = warp_sum_d(acc);
    if (lane == 0) {
        G[((size_t)slot * nz + i) * nz + j] = acc;
        G[((size_t)slot * nz + j) * nz + i] = acc;
    }
}

// Block-cooperative LU with partial pivoting on A (n x n, row-major, in global scratch), applied
// to nrhs right-hand sides B (n x nrhs, row-major).  On return B holds the solutions.
// A zero (or non-finite) pivot — a duplicated labelled clip, more independent rows than dimensions — sets *singular:
// numpy.linalg.inv raises LinAlgError for such a matrix (target_clip.py:194,248), the caller here turns the flag into an error.
__device__ void lu_solve(double *A, double *B, int n, int nrhs, int *piv_s, int *singular) {
    __shared__ double red_v[32];
    __shared__ int red_i[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int k = 0; k < n; ++k) {
        // pivot: largest |A[i][k]|, i >= k, lowest i among equals (what a serial scan finds) — block-wide arg-max
        double best = -1.0;
        int bi = k;
        for (int i = k + threadIdx.x; i < n; i += blockDim.x) {
            const double v = fabs(A[(size_t)i * n + k]);
            if (v > best) { best = v; bi = i; }                      // a thread's candidates come in ascending i
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { red_v[wid] = best; red_i[wid] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < n_warps; ++w)
                if (red_v[w] > best || (red_v[w] == best && red_i[w] < bi)) { best = red_v[w]; bi = red_i[w]; }
            *piv_s = bi;
        }
        __syncthreads();
        const int p = *piv_s;
        if (p != k) {
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                const double t = A[(size_t)k * n + j];
                A[(size_t)k * n + j] = A[(size_t)p * n + j];
                A[(size_t)p * n + j] = t;
            }
            for (int j = threadIdx.x; j < nrhs; j += blockDim.x) {
                const double t = B[(size_t)k * nrhs + j];
                B[(size_t)k * nrhs + j] = B[(size_t)p * nrhs + j];
                B[(size_t)p * nrhs + j] = t;
            }
        }
        __syncthreads();
        const double pivot = A[(size_t)k * n + k];
        if (threadIdx.x == 0 && !(fabs(pivot) > 0.0 && fabs(pivot) < 1.0e300)) *singular = 1;
        for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) A[(size_t)i * n + k] /= pivot;
        __syncthreads();
        // trailing update, one warp per row: lanes walk the row's columns (A's, then B's) — coalesced, no division
        const int rem = n - k - 1;
        for (int r = wid; r < rem; r += n_warps) {
            const int i = k + 1 + r;
            const double l = A[(size_t)i * n + k];
            double *Ai = A + (size_t)i * n + (k + 1);
            const double *Ak = A + (size_t)k * n + (k + 1);
            for (int c = lane; c < rem; c += 32) Ai[c] -= l * Ak[c];
            double *Bi = B + (size_t)i * nrhs;
            const double *Bk = B + (size_t)k * nrhs;
            for (int c = lane; c < nrhs; c += 32) Bi[c] -= l * Bk[c];
        }
        __syncthreads();
    }
    // back substitution, one column of B per thread
    for (int j = threadIdx.x; j < nrhs; j += blockDim.x) {
        for (int i = n - 1; i >= 0; --i) {
            double v = B[(size_t)i * nrhs + j];
            for (int c = i + 1; c < n; ++c) v -= A[(size_t)i * n + c] * B[(size_t)c * nrhs + j];
            B[(size_t)i * nrhs + j] = v / A[(size_t)i * n + i];
        }
    }
    __syncthreads();
}

// One block per slot.  Computes coefficient vectors a (n) and b (m) with
//     w = X^T a + Y^T b
// equal to the reference's w_final (target_clip.py:248-260), via the Woodbury identity so that only
// (n+m)-sized Gram blocks are needed instead of the reference's dim x dim inverse:
//     c = mu / tr(Gyy);  K = c (I + c Gyy)^-1;  B = Gxx - Gxy K Gyx;  beta = B^-1 1;
//     gamma = 1 - K Gyy 1;  delta = B^-1 Gxy gamma;
//     a = beta - c delta;   b = c gamma + c K Gyx delta - K Gyx beta.
// With m = 0 (or mu = 0 => c = 0, K = 0) this is a = Gxx^-1 1, the valid-only rule (:194-198).
// Scratch per slot (doubles): Bm[n*n] | R[n*2] | Ky[m*m] | Z[m*(n+1)] | tmp[n+m]
__global__ void __launch_bounds__(1024)
bootstrap_solve_kernel(const double *__restrict__ G, int n, int m, double mu, double *scratch,
                       size_t scratch_per_slot, double *coef, int *singular, const unsigned char *__restrict__ slot_mask) {
    __shared__ int piv_s;
    __shared__ double c_s;
    const int slot = blockIdx.x;
    const int nz = n + m;
    if (slot_mask && !slot_mask[slot]) {                 // a slot the caller does not want: zero coefficients, no solve
        for (int i = threadIdx.x; i < nz; i += blockDim.x) coef[(size_t)slot * nz + i] = 0.0;
        return;
    }
    const double *Gs = G + (size_t)slot * nz * nz;
    double *Bm = scratch + (size_t)slot * scratch_per_slot;
    double *R = Bm + (size_t)n * n;
    double *Ky = R + (size_t)n * 2;
    double *Z = Ky + (size_t)m * m;
    double *co = coef + (size_t)slot * nz;
#define GXX(i, j) Gs[(size_t)(i) * nz + (j)]
#define GXY(i, j) Gs[(size_t)(i) * nz + n + (j)]
#define GYY(i, j) Gs[(size_t)(n + (i)) * nz + n + (j)]
    if (threadIdx.x == 0) {
        double tr = 0.0;
        for (int j = 0; j < m; ++j) tr += GYY(j, j);
        c_s = (m > 0) ? mu / tr : 0.0;
    }
    __syncthreads();
    const double c = c_s;
    const bool use_y = (m > 0) && (c != 0.0);
    if (use_y) {
        // Z = (I + c Gyy)^-1 [ c Gyx | c Gyy 1 ]   -> Z[:, :n] = K Gyx, Z[:, n] = K Gyy 1
        for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
            const int i = e / m, j = e % m;
            Ky[e] = (i == j ? 1.0 : 0.0) + c * GYY(i, j);
        }
        for (int e = threadIdx.x; e < m * (n + 1); e += blockDim.x) {
            const int i = e / (n + 1), j = e % (n + 1);
            double v;
            if (j < n) v = c * GXY(j, i);
            else {
                v = 0.0;
                for (int q = 0; q < m; ++q) v += GYY(i, q);
                v *= c;
            }
            Z[e] = v;
        }
        __syncthreads();
        lu_solve(Ky, Z, m, n + 1, &piv_s, singular);
    }
    // B = Gxx - Gxy (K Gyx);  R = [1 | Gxy gamma],  gamma = 1 - K Gyy 1
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e % n;
        double v = GXX(i, j);
        if (use_y)
            for (int q = 0; q < m; ++q) v -= GXY(i, q) * Z[(size_t)q * (n + 1) + j];
        Bm[e] = v;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        R[2 * i] = 1.0;
        double v = 0.0;
        if (use_y)
            for (int q = 0; q < m; ++q) v += GXY(i, q) * (1.0 - Z[(size_t)q * (n + 1) + n]);
        R[2 * i + 1] = v;
    }
    __syncthreads();
    lu_solve(Bm, R, n, 2, &piv_s, singular);                 // R[:,0] = beta, R[:,1] = delta
    for (int i = threadIdx.x; i < n; i += blockDim.x) co[i] = R[2 * i] - c * R[2 * i + 1];
    for (int j = threadIdx.x; j < m; j += blockDim.x) {
        double v = 0.0;
        if (use_y) {
            const double gamma = 1.0 - Z[(size_t)j * (n + 1) + n];
            double kd = 0.0, kb = 0.0;               // (K Gyx delta)_j, (K Gyx beta)_j
            for (int i = 0; i < n; ++i) {
                const double z = Z[(size_t)j * (n + 1) + i];
                kd += z * R[2 * i + 1];
                kb += z * R[2 * i];
            }
            v = c * gamma + c * kd - kb;
        }
        co[n + j] = v;
    }
#undef GXX
#undef GXY
#undef GYY
}

// w[slot][d] = sum_i coef[slot][i] * z_i[slot][d]
__global__ void combine_kernel(const float *__restrict__ rows, const long long *__restrict__ ids, int nz,
                               int n_slots, int dim, size_t row_floats, const double *__restrict__ coef,
                               double *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots * dim) return;
    const int slot = i / dim, d = i - slot * dim;
    double acc = 0.0;
    for (int r = 0; r < nz; ++r)
        acc = fma(coef[(size_t)slot * nz + r], (double)rows[(size_t)ids[r] * row_floats + (size_t)slot * dim + d], acc);
    out[i] = acc;
}

using vq::scratch_reserve;
static inline int lab_reserve(vq_store *s, size_t dev_bytes, size_t host_bytes) { return scratch_reserve(s, dev_bytes, host_bytes); }

struct Carver {                                   // 256-byte aligned pieces of a block
    char *base;
    size_t off = 0;
    explicit Carver(char *b) : base(b) {}
    template <class T> T *take(size_t n) {
        T *p = reinterpret_cast<T *>(base + off);
        off += (n * sizeof(T) + 255) & ~(size_t)255;
        return p;
    }
};
inline size_t padded(size_t bytes) { return (bytes + 255) & ~(size_t)255; }

int local_rows(const vq_store *s, const int64_t *rows, int64_t n, std::vector<long long> &out, const char *who) {
    out.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t r = rows[i] - s->first_global_row;
        VQ_REQUIRE(r >= 0 && r < s->n_rows, "%s: global row %lld is not in this shard [%lld, %lld)", who,
                   (long long)rows[i], (long long)s->first_global_row,
                   (long long)(s->first_global_row + s->n_rows));
        out[(size_t)i] = r;
    }
    return 0;
}

}  // namespace

extern "C" int vq_labelled_sims(vq_store *s, const double *target, const int64_t *rows, int64_t n,
                                double *sims_out) {
    VQ_REQUIRE(s && target && (rows || n == 0) && (sims_out || n == 0), "vq_labelled_sims: null argument");
    if (n == 0) return 0;
    VQ_CUDA(cudaSetDevice(s->device));
    const size_t t_bytes = s->row_floats * sizeof(double), id_bytes = (size_t)n * sizeof(long long),
                 out_bytes = (size_t)n * s->n_streams * sizeof(double);
    const size_t total = padded(t_bytes) + padded(id_bytes) + padded(out_bytes);
    if (int r = lab_reserve(s, total, total)) return r;
    Carver h(s->lab_host), d(s->lab_dev);
    double *h_t = h.take<double>(s->row_floats), *d_t = d.take<double>(s->row_floats);
    long long *h_ids = h.take<long long>((size_t)n), *d_ids = d.take<long long>((size_t)n);
    double *h_out = h.take<double>((size_t)n * s->n_streams), *d_out = d.take<double>((size_t)n * s->n_streams);
    memcpy(h_t, target, t_bytes);
    for (int64_t i = 0; i < n; ++i) {
        const int64_t r = rows[i] - s->first_global_row;
        VQ_REQUIRE(r >= 0 && r < s->n_rows, "vq_labelled_sims: global row %lld is not in this shard [%lld, %lld)",
                   (long long)rows[i], (long long)s->first_global_row, (long long)(s->first_global_row + s->n_rows));
        h_ids[i] = r;
    }
    // target and ids are adjacent in both blocks: one copy in
    VQ_CUDA(cudaMemcpyAsync(d_t, h_t, padded(t_bytes) + id_bytes, cudaMemcpyHostToDevice, s->stream));
    const long long warps = n * s->n_streams;
    const int blocks = (int)((warps * 32 + 255) / 256);
    labelled_sims_kernel<<<blocks, 256, 0, s->stream>>>(s->rows, d_t, d_ids, s->inv_counts, n, s->n_streams, s->n_splits,
                                                        s->dim, d_out);
    VQ_CUDA(cudaGetLastError());
    VQ_CUDA(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, s->stream));
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    memcpy(sims_out, h_out, out_bytes);
    return 0;
}

namespace {
struct ArenaView {
    char *p;
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};
}  // namespace

extern "C" int vq_loss_grid(int device, const double *sims, const uint8_t *labels, int64_t L,
                            const double *weight_grid, int32_t n_w, const double *threshold_grid, int32_t n_th,
                            double ballast, const int32_t *rep_offset, const int32_t *rep_index, int32_t R,
                            double *losses_out) {
    VQ_REQUIRE(sims && labels && weight_grid && threshold_grid && rep_offset && rep_index && losses_out,
               "vq_loss_grid: null argument");
    VQ_REQUIRE(L > 0 && n_w > 0 && R > 0 && n_th > 0 && n_th <= 65535 * kMaxTh, "vq_loss_grid: need L, R, n_w, n_th > 0");
    VQ_REQUIRE((int64_t)R * n_w < (1ll << 31), "vq_loss_grid: %d replicates x %d weights exceed the launch grid", R, n_w);
    for (int r = 0; r < R; ++r)
        VQ_REQUIRE(rep_offset[r + 1] > rep_offset[r], "vq_loss_grid: replicate %d is empty", r);
    const int64_t n_idx = rep_offset[R];
    for (int64_t i = 0; i < n_idx; ++i)
        VQ_REQUIRE(rep_index[i] >= 0 && rep_index[i] < L, "vq_loss_grid: replicate index out of range");
    VQ_CUDA(cudaSetDevice(device));
    const bool timing = getenv("VQ_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms_since = [&](std::chrono::steady_clock::time_point t) { return std::chrono::duration<double, std::milli>(now() - t).count(); };
    auto t_start = now();
    // One grow-only arena per device, kept for the life of the process: with a multi-gigabyte store resident, every
    // cudaMalloc / cudaFree costs ~2 ms, which was most of this call (8 of each per call before).
    struct Arena { char *p = nullptr; size_t cap = 0; char *h = nullptr; size_t hcap = 0; cudaStream_t st = nullptr; };
    static Arena arenas[64];
    static std::mutex arena_mutex;
    std::lock_guard<std::mutex> lock(arena_mutex);
    VQ_REQUIRE(device >= 0 && device < 64, "vq_loss_grid: device %d", device);
    // inputs back to back (one pinned block, one copy in), then the score table and the output grid
    const size_t sizes[8] = {(size_t)L * 2 * sizeof(double), (size_t)L, (size_t)n_w * sizeof(double), (size_t)n_th * sizeof(double),
                             (size_t)(R + 1) * sizeof(int), (size_t)n_idx * sizeof(int), (size_t)L * n_w * sizeof(double),
                             (size_t)R * n_w * n_th * sizeof(double)};
    size_t offs[8], total = 0;
    for (int i = 0; i < 8; ++i) { offs[i] = total; total += (sizes[i] + 255) & ~(size_t)255; }
    const size_t in_bytes = offs[6], out_bytes = sizes[7];
    Arena &ar = arenas[device];
    if (!ar.st) VQ_CUDA(cudaStreamCreateWithFlags(&ar.st, cudaStreamNonBlocking));
    if (total > ar.cap) {
        if (ar.p) cudaFree(ar.p);
        ar.p = nullptr;
        ar.cap = 0;
        const size_t want = total + total / 2;
        VQ_CUDA(cudaMalloc((void **)&ar.p, want));
        ar.cap = want;
    }
    const size_t host_need = in_bytes + out_bytes;
    if (host_need > ar.hcap) {
        if (ar.h) cudaFreeHost(ar.h);
        ar.h = nullptr;
        ar.hcap = 0;
        const size_t want = host_need + host_need / 2;
        VQ_CUDA(cudaMallocHost((void **)&ar.h, want));
        ar.hcap = want;
    }
    ArenaView d_sims{ar.p + offs[0]}, d_lab{ar.p + offs[1]}, d_w{ar.p + offs[2]}, d_th{ar.p + offs[3]}, d_off{ar.p + offs[4]},
        d_idx{ar.p + offs[5]}, d_tab{ar.p + offs[6]}, d_out{ar.p + offs[7]};
    if (timing) fprintf(stderr, "[vq_loss_grid] alloc %.2f ms\n", ms_since(t_start));
    t_start = now();
    const void *src[6] = {sims, labels, weight_grid, threshold_grid, rep_offset, rep_index};
    for (int i = 0; i < 6; ++i) memcpy(ar.h + offs[i], src[i], sizes[i]);
    VQ_CUDA(cudaMemcpyAsync(ar.p, ar.h, in_bytes, cudaMemcpyHostToDevice, ar.st));
    score_table_kernel<<<(unsigned int)((L * n_w + 255) / 256), 256, 0, ar.st>>>(d_sims.as<double>(), L, d_w.as<double>(),
                                                                                 n_w, d_tab.as<double>());
    loss_grid_kernel<<<dim3((unsigned int)(R * n_w), (unsigned int)((n_th + kMaxTh - 1) / kMaxTh)), kLossThreads, 0, ar.st>>>(d_tab.as<double>(), d_lab.as<unsigned char>(), L,
                                                                          d_th.as<double>(), n_th, n_w, ballast,
                                                                          d_off.as<int>(), d_idx.as<int>(), d_out.as<double>());
    VQ_CUDA(cudaGetLastError());
    VQ_CUDA(cudaMemcpyAsync(ar.h + in_bytes, d_out.p, out_bytes, cudaMemcpyDeviceToHost, ar.st));
    VQ_CUDA(cudaStreamSynchronize(ar.st));
    memcpy(losses_out, ar.h + in_bytes, out_bytes);
    if (timing) fprintf(stderr, "[vq_loss_grid] copies + kernels %.2f ms\n", ms_since(t_start));
    return 0;
}

extern "C" int vq_bootstrap_target(vq_store *s, const int64_t *valid_rows, int32_t n_valid,
                                   const int64_t *invalid_rows, int32_t n_invalid, double mu, const uint8_t *slot_mask,
                                   double *target_out) {
    VQ_REQUIRE(s && valid_rows && target_out, "vq_bootstrap_target: null argument");
    VQ_REQUIRE(n_valid >= 1 && n_invalid >= 0 && (n_invalid == 0 || invalid_rows),
               "vq_bootstrap_target: need at least one valid row");
    VQ_REQUIRE(n_valid <= s->dim, "vq_bootstrap_target: %d valid rows exceed the feature dimension %d (singular system)",
               n_valid, s->dim);
    const int n = n_valid, m = n_invalid, nz = n + m;
    const int n_slots = s->n_streams * s->n_splits;
    std::vector<long long> loc, loc2;
    if (int r = local_rows(s, valid_rows, n, loc, "vq_bootstrap_target")) return r;
    if (m) {
        if (int r = local_rows(s, invalid_rows, m, loc2, "vq_bootstrap_target")) return r;
        loc.insert(loc.end(), loc2.begin(), loc2.end());
    }
    VQ_CUDA(cudaSetDevice(s->device));
    const size_t per_slot = (size_t)n * n + (size_t)n * 2 + (size_t)m * m + (size_t)m * (n + 1) + (size_t)nz + 8;
    const size_t out_doubles = (size_t)n_slots * s->dim;
    const size_t dev_total = padded((size_t)nz * 8) + padded(16) + padded((size_t)n_slots) + padded((size_t)n_slots * nz * nz * 8) +
                             padded((size_t)n_slots * per_slot * 8) + padded((size_t)n_slots * nz * 8) + padded(out_doubles * 8);
    const size_t host_total = padded((size_t)nz * 8) + padded(16) + padded((size_t)n_slots) + padded(out_doubles * 8);
    if (int r = lab_reserve(s, dev_total, host_total)) return r;
    Carver d(s->lab_dev), h(s->lab_host);
    long long *d_ids = d.take<long long>((size_t)nz), *h_ids = h.take<long long>((size_t)nz);
    int *d_flag = d.take<int>(4), *h_flag = h.take<int>(4);
    unsigned char *d_mask = d.take<unsigned char>((size_t)n_slots), *h_mask = h.take<unsigned char>((size_t)n_slots);
    for (int i = 0; i < n_slots; ++i) h_mask[i] = slot_mask ? (slot_mask[i] ? 1 : 0) : 1;
    double *d_G = d.take<double>((size_t)n_slots * nz * nz);
    double *d_scr = d.take<double>((size_t)n_slots * per_slot);
    double *d_coef = d.take<double>((size_t)n_slots * nz);
    double *d_out = d.take<double>(out_doubles), *h_out = h.take<double>(out_doubles);
    memcpy(h_ids, loc.data(), (size_t)nz * sizeof(long long));
    h_flag[0] = 0;
    VQ_CUDA(cudaMemcpyAsync(d_ids, h_ids, padded((size_t)nz * 8) + padded(16) + (size_t)n_slots, cudaMemcpyHostToDevice,
                            s->stream));                                                          // ids + cleared flag + mask
    const long long warps = (long long)nz * nz * n_slots;
    gram_kernel<<<(unsigned int)((warps * 32 + 255) / 256), 256, 0, s->stream>>>(s->rows, d_ids, nz, n_slots, s->dim,
                                                                               s->row_floats, d_G);
    bootstrap_solve_kernel<<<n_slots, 1024, 0, s->stream>>>(d_G, n, m, mu, d_scr, per_slot, d_coef, d_flag, d_mask);
    combine_kernel<<<(n_slots * s->dim + 255) / 256, 256, 0, s->stream>>>(s->rows, d_ids, nz, n_slots, s->dim, s->row_floats,
                                                                         d_coef, d_out);
    VQ_CUDA(cudaGetLastError());
    VQ_CUDA(cudaMemcpyAsync(h_out, d_out, out_doubles * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    VQ_CUDA(cudaMemcpyAsync(h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    VQ_CUDA(cudaStreamSynchronize(s->stream));
    VQ_REQUIRE(h_flag[0] == 0, "vq_bootstrap_target: singular system (a labelled clip appears twice, or the labelled rows are "
               "linearly dependent): numpy.linalg.inv raises LinAlgError here (target_clip.py:194,248)");
    for (size_t i = 0; i < out_doubles; ++i)
        VQ_REQUIRE(h_out[i] == h_out[i] && h_out[i] - h_out[i] == 0.0, "vq_bootstrap_target: the solve produced a non-finite target "
                   "(ill-conditioned labelled set)");
    memcpy(target_out, h_out, out_doubles * sizeof(double));
    return 0;
}
