/* A plain C99 consumer of include/vq.h — no Python, no ctypes: what a non-Python host (the C side of a cgo / JNI / N-API
 * binding) would do for one query.  Builds a synthetic shard, takes the target from one of its rows the way
 * TargetClip._scale_feature does (target_clip.py:311-313: f / (f . f), in double), scans with host buffers and prints
 * the counts, the top-k and a checksum of the ordered lists as one JSON line; tests/test_gpu_parity.py compares the line
 * with what the Python host gets through ctypes for the same shard.  Also exercises the error contract: a bad call
 * returns < 0 and leaves a message in vq_last_error().
 *   gcc -std=c99 -I include tests/c_abi/scan_consumer.c -L video_query_algorithms_b200/lib -lvq_b200 -o scan_consumer  */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "vq.h"

#define CHECK(call)                                                                      \
    do {                                                                                 \
        int rc__ = (call);                                                               \
        if (rc__ != 0) {                                                                 \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, vq_last_error());       \
            return 1;                                                                    \
        }                                                                                \
    } while (0)

int main(int argc, char **argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : 50000;
    const uint64_t seed = argc > 2 ? strtoull(argv[2], NULL, 10) : 20261018ull;
    const int64_t ref_row = argc > 3 ? atoll(argv[3]) : 18120;
    const int S = 2, P = 1, D = 1024, K = 100;
    if (vq_abi_version() != VQ_ABI_VERSION) {
        fprintf(stderr, "ABI %d, header %d\n", vq_abi_version(), VQ_ABI_VERSION);
        return 1;
    }
    int n_dev = 0;
    CHECK(vq_device_count(&n_dev));
    if (n_dev < 1) {
        fprintf(stderr, "no CUDA device\n");
        return 1;
    }
    /* the error contract first: a null handle is refused with a message, nothing crashes */
    if (vq_store_sync(NULL) >= 0 || strlen(vq_last_error()) == 0) {
        fprintf(stderr, "a bad call must return < 0 and leave a message\n");
        return 1;
    }
    vq_store *st = NULL;
    CHECK(vq_store_create(&st, 0, n, S, P, D, 0));
    CHECK(vq_store_fill_synthetic(st, seed, NULL));
    float *row = (float *)malloc(sizeof(float) * S * P * D);
    float *target = (float *)malloc(sizeof(float) * S * P * D);
    CHECK(vq_store_download(st, ref_row, 1, row));
    for (int s = 0; s < S; ++s) {                              /* target = f / (f . f) per stream, double arithmetic */
        double ff = 0.0;
        for (int d = 0; d < D; ++d) ff += (double)row[s * D + d] * (double)row[s * D + d];
        for (int d = 0; d < D; ++d) target[s * D + d] = (float)((double)row[s * D + d] / ff);
    }
    vq_scan_params p;
    memset(&p, 0, sizeof p);
    p.weights[0] = 1.0;
    p.weights[1] = 1.5;
    p.threshold = 0.8;
    p.lower_limit = 0.8 - 0.35 * (1.0 - 0.8);
    p.eps = 3e-6;
    p.topk = K;
    vq_scan_counts c;
    CHECK(vq_scan(st, target, &p, &c));
    int64_t *rows = (int64_t *)malloc(sizeof(int64_t) * (size_t)(c.n_match + c.n_near + K + 1));
    float *scores = (float *)malloc(sizeof(float) * (size_t)(c.n_match + c.n_near + K + 1));
    /* checksums: sum of rows and sum of score bit patterns of the match and near-miss lists, copied out of the mirror */
    uint64_t sum_rows[2] = {0, 0}, sum_bits[2] = {0, 0};
    CHECK(vq_fetch_matches(st, c.n_match, rows, scores));
    for (int64_t i = 0; i < c.n_match; ++i) {
        uint32_t b;
        memcpy(&b, &scores[i], 4);
        sum_rows[0] += (uint64_t)rows[i];
        sum_bits[0] += b;
        if (i && rows[i] <= rows[i - 1]) {
            fprintf(stderr, "match list is not in database order\n");
            return 1;
        }
    }
    CHECK(vq_fetch_near(st, c.n_near, rows, scores));
    for (int64_t i = 0; i < c.n_near; ++i) {
        uint32_t b;
        memcpy(&b, &scores[i], 4);
        sum_rows[1] += (uint64_t)rows[i];
        sum_bits[1] += b;
    }
    /* the zero-copy view must hold the same list */
    const int64_t *v_rows = NULL;
    const float *v_scores = NULL;
    int64_t v_n = 0;
    CHECK(vq_scan_host_list(st, 1, &v_rows, &v_scores, &v_n));
    if (v_n != c.n_near || (v_n && (memcmp(v_rows, rows, (size_t)v_n * 8) || memcmp(v_scores, scores, (size_t)v_n * 4)))) {
        fprintf(stderr, "host view differs from the fetched near-miss list\n");
        return 1;
    }
    CHECK(vq_fetch_topk(st, K, rows, scores));
    printf("{\"n_match\": %" PRId64 ", \"n_near\": %" PRId64 ", \"n_tie\": %" PRId64 ", \"n_topk\": %d, "
           "\"match_rows_sum\": %" PRIu64 ", \"match_bits_sum\": %" PRIu64 ", \"near_rows_sum\": %" PRIu64
           ", \"near_bits_sum\": %" PRIu64 ", \"topk_rows\": [",
           c.n_match, c.n_near, c.n_tie, c.n_topk, sum_rows[0], sum_bits[0], sum_rows[1], sum_bits[1]);
    for (int i = 0; i < c.n_topk; ++i) printf("%s%" PRId64, i ? ", " : "", rows[i]);
    printf("], \"top1_score\": %.9g}\n", c.n_topk ? (double)scores[0] : 0.0);
    CHECK(vq_store_destroy(st));
    free(row);
    free(target);
    free(rows);
    free(scores);
    return 0;
}
