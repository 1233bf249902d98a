// Batched queries (placeholder until the tcgen05 kernel lands): fails loudly, never falls back.
#include "vq_internal.cuh"

extern "C" int vq_scan_batch(vq_store *s, const float *targets, int32_t n_queries, const vq_scan_params *p,
                             int64_t *counts_out, int64_t *topk_rows_out, float *topk_scores_out,
                             float *kernel_ms_out) {
    (void)s; (void)targets; (void)n_queries; (void)p; (void)counts_out; (void)topk_rows_out;
    (void)topk_scores_out; (void)kernel_ms_out;
    vq::set_error("vq_scan_batch: the batched tcgen05 kernel is not part of this build");
    return -4;
}
