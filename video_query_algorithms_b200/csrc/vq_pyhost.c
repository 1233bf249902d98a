/*
 * vq_pyhost — host glue between CPython objects and the store's row layout (plain C, CPython API; no CUDA, no arithmetic
 * of the scoring path).  Loaded with ctypes.PyDLL, i.e. every entry point is called WITH the GIL held.
 *
 * The `search-sets/features` response (reference src/models/ticket.py:363-381) arrives as a list of dicts
 * {dnn_stream_id, dnn_stream_split, name, video_clip_id, feature_vector: [1024 Python floats]}; a million clips are two
 * million such records and two billion boxed floats.  Walking them in the interpreter costs ~13 us per vector; here the
 * record fields are read in one pass (vq_py_index_records) and the vectors are unboxed straight into a pinned staging
 * chunk in the store's row layout by several threads (vq_py_fill_chunk).  The worker threads only READ objects that
 * the calling thread keeps alive — list item pointers, type pointers and ob_fval of exact floats; they never touch a
 * reference count or call into the interpreter — and the caller holds the GIL for the whole call, so nothing they read
 * can change underneath them.  Vectors that hold anything but exact floats are converted by the calling thread.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

static char g_err[512];
const char *vq_py_last_error(void) { return g_err; }
int vq_py_abi_version(void) { return 1; }

static int fail(const char *msg, Py_ssize_t i) {
    snprintf(g_err, sizeof(g_err), "%s (record %lld)", msg, (long long)i);
    PyErr_Clear();
    return -1;
}

/* Pass 1.  records: list of dicts.  streams: tuple of str.  For record i: stream_out[i] = index of its dnn_stream_id
 * in `streams`, or -1 when the record is filtered out (other stream, other feature name: ticket.py:374-381);
 * clip_out[i] = video_clip_id, split_out[i] = int(dnn_stream_split), len_out[i] = len(feature_vector).              */
int vq_py_index_records(PyObject *records, PyObject *streams, PyObject *feature_name, int64_t *clip_out,
                        int32_t *stream_out, int32_t *split_out, int32_t *len_out) {
    if (!records || !PyList_Check(records) || !streams || !PyTuple_Check(streams) || !feature_name || !clip_out ||
        !stream_out || !split_out || !len_out) {
        snprintf(g_err, sizeof(g_err), "vq_py_index_records: need a list of records, a tuple of streams and output arrays");
        return -1;
    }
    PyObject *k_stream = PyUnicode_InternFromString("dnn_stream_id"), *k_split = PyUnicode_InternFromString("dnn_stream_split"),
             *k_name = PyUnicode_InternFromString("name"), *k_clip = PyUnicode_InternFromString("video_clip_id"),
             *k_vec = PyUnicode_InternFromString("feature_vector");
    const Py_ssize_t n = PyList_GET_SIZE(records), n_streams = PyTuple_GET_SIZE(streams);
    int rc = 0;
    for (Py_ssize_t i = 0; i < n && rc == 0; ++i) {
        PyObject *rec = PyList_GET_ITEM(records, i);
        stream_out[i] = -1;
        clip_out[i] = 0;
        split_out[i] = 0;
        len_out[i] = 0;
        if (!PyDict_Check(rec)) { rc = fail("vq_py_index_records: record is not a dict", i); break; }
        PyObject *sid = PyDict_GetItemWithError(rec, k_stream), *nm = PyDict_GetItemWithError(rec, k_name);
        if (!sid || !nm) { rc = fail("vq_py_index_records: record lacks dnn_stream_id / name", i); break; }
        int si = -1;
        for (Py_ssize_t s = 0; s < n_streams; ++s) {
            const int eq = PyObject_RichCompareBool(sid, PyTuple_GET_ITEM(streams, s), Py_EQ);
            if (eq < 0) { rc = fail("vq_py_index_records: cannot compare dnn_stream_id", i); break; }
            if (eq) { si = (int)s; break; }
        }
        if (rc || si < 0) continue;
        const int same = PyObject_RichCompareBool(nm, feature_name, Py_EQ);
        if (same < 0) { rc = fail("vq_py_index_records: cannot compare name", i); break; }
        if (!same) continue;
        PyObject *clip = PyDict_GetItemWithError(rec, k_clip), *split = PyDict_GetItemWithError(rec, k_split),
                 *vec = PyDict_GetItemWithError(rec, k_vec);
        if (!clip || !split || !vec) { rc = fail("vq_py_index_records: record lacks video_clip_id / dnn_stream_split / feature_vector", i); break; }
        const long long c = PyLong_AsLongLong(clip);
        if (c == -1 && PyErr_Occurred()) { rc = fail("vq_py_index_records: video_clip_id is not an integer", i); break; }
        PyObject *sp = PyNumber_Long(split);                      /* int(tf["dnn_stream_split"]) */
        if (!sp) { rc = fail("vq_py_index_records: dnn_stream_split is not a number", i); break; }
        const long p = PyLong_AsLong(sp);
        Py_DECREF(sp);
        if (p == -1 && PyErr_Occurred()) { rc = fail("vq_py_index_records: dnn_stream_split out of range", i); break; }
        const Py_ssize_t len = PyObject_Length(vec);
        if (len < 0) { rc = fail("vq_py_index_records: feature_vector has no length", i); break; }
        clip_out[i] = c;
        split_out[i] = (int32_t)p;
        len_out[i] = (int32_t)len;
        stream_out[i] = si;
    }
    Py_DECREF(k_stream); Py_DECREF(k_split); Py_DECREF(k_name); Py_DECREF(k_clip); Py_DECREF(k_vec);
    return rc;
}

typedef struct {
    PyObject **vecs;          /* borrowed: the feature_vector lists of the records to convert */
    const int64_t *dest;      /* float offset of each vector inside the chunk */
    int64_t lo, hi, dim;
    float *chunk;
    volatile int slow;        /* a vector needs the interpreter (not a list of exact floats) */
} fill_job;

static void *fill_worker(void *arg) {
    fill_job *j = (fill_job *)arg;
    for (int64_t i = j->lo; i < j->hi; ++i) {
        PyObject *v = j->vecs[i];
        if (!PyList_CheckExact(v) || PyList_GET_SIZE(v) != j->dim) { j->slow = 1; continue; }
        float *out = j->chunk + j->dest[i];
        PyObject **items = ((PyListObject *)v)->ob_item;
        int ok = 1;
        for (int64_t d = 0; d < j->dim; ++d) {
            PyObject *x = items[d];
            if (!PyFloat_CheckExact(x)) { ok = 0; break; }
            out[d] = (float)PyFloat_AS_DOUBLE(x);
        }
        if (!ok) j->slow = 1;
    }
    return NULL;
}

/* Pass 2.  For the m records rec_index[0..m) (ascending: a later record of the same slot overwrites an earlier one, the
 * reference's dict semantics) write float32(feature_vector) at chunk + dest[i].  Lists of exact floats are unboxed by
 * n_threads workers; anything else (tuples, ints, numpy rows) by the calling thread through the number protocol.     */
int vq_py_fill_chunk(PyObject *records, const int64_t *rec_index, const int64_t *dest, int64_t m, int64_t dim,
                     float *chunk, int32_t n_threads) {
    if (!records || !PyList_Check(records) || (m > 0 && (!rec_index || !dest || !chunk)) || dim <= 0 || m < 0) {
        snprintf(g_err, sizeof(g_err), "vq_py_fill_chunk: bad argument");
        return -1;
    }
    if (m == 0) return 0;
    PyObject *k_vec = PyUnicode_InternFromString("feature_vector");
    PyObject **vecs = (PyObject **)malloc((size_t)m * sizeof(PyObject *));
    if (!vecs) { Py_DECREF(k_vec); snprintf(g_err, sizeof(g_err), "vq_py_fill_chunk: out of memory"); return -1; }
    const Py_ssize_t n = PyList_GET_SIZE(records);
    int rc = 0;
    for (int64_t i = 0; i < m && rc == 0; ++i) {
        if (rec_index[i] < 0 || rec_index[i] >= n) { rc = fail("vq_py_fill_chunk: record index out of range", (Py_ssize_t)rec_index[i]); break; }
        PyObject *rec = PyList_GET_ITEM(records, (Py_ssize_t)rec_index[i]);
        vecs[i] = PyDict_Check(rec) ? PyDict_GetItemWithError(rec, k_vec) : NULL;
        if (!vecs[i]) rc = fail("vq_py_fill_chunk: record has no feature_vector", (Py_ssize_t)rec_index[i]);
    }
    Py_DECREF(k_vec);
    if (rc) { free(vecs); return rc; }
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 64) n_threads = 64;
    if ((int64_t)n_threads > m) n_threads = (int32_t)m;
    fill_job jobs[64];
    pthread_t tids[64];
    int started[64];
    for (int t = 0; t < n_threads; ++t) {
        jobs[t].vecs = vecs; jobs[t].dest = dest; jobs[t].dim = dim; jobs[t].chunk = chunk; jobs[t].slow = 0;
        jobs[t].lo = m * t / n_threads;
        jobs[t].hi = m * (t + 1) / n_threads;
        started[t] = 0;
    }
    for (int t = 1; t < n_threads; ++t) started[t] = pthread_create(&tids[t], NULL, fill_worker, &jobs[t]) == 0;
    fill_worker(&jobs[0]);
    for (int t = 1; t < n_threads; ++t) {
        if (started[t]) pthread_join(tids[t], NULL);
        else fill_worker(&jobs[t]);
    }
    int slow = 0;
    for (int t = 0; t < n_threads; ++t) slow |= jobs[t].slow;
    if (slow) {                                              /* second pass, in order, for what the workers skipped */
        for (int64_t i = 0; i < m && rc == 0; ++i) {
            PyObject *v = vecs[i];
            if (PyList_CheckExact(v) && PyList_GET_SIZE(v) == dim) {
                int exact = 1;
                PyObject **items = ((PyListObject *)v)->ob_item;
                for (int64_t d = 0; d < dim && exact; ++d) exact = PyFloat_CheckExact(items[d]);
                if (exact) {                                 /* may have been overwritten out of order by a slow neighbour: redo */
                    float *out = chunk + dest[i];
                    for (int64_t d = 0; d < dim; ++d) out[d] = (float)PyFloat_AS_DOUBLE(items[d]);
                    continue;
                }
            }
            PyObject *seq = PySequence_Fast(v, "feature_vector is not a sequence");
            if (!seq) { rc = fail("vq_py_fill_chunk: feature_vector is not a sequence", (Py_ssize_t)rec_index[i]); break; }
            if (PySequence_Fast_GET_SIZE(seq) != dim) {
                Py_DECREF(seq);
                rc = fail("vq_py_fill_chunk: feature_vector length differs from the store's dimension", (Py_ssize_t)rec_index[i]);
                break;
            }
            float *out = chunk + dest[i];
            for (int64_t d = 0; d < dim; ++d) {
                const double x = PyFloat_AsDouble(PySequence_Fast_GET_ITEM(seq, d));
                if (x == -1.0 && PyErr_Occurred()) { rc = fail("vq_py_fill_chunk: feature_vector holds a non-number", (Py_ssize_t)rec_index[i]); break; }
                out[d] = (float)x;
            }
            Py_DECREF(seq);
        }
    }
    free(vecs);
    return rc;
}
