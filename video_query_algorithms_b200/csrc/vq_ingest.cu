// Native reader for the feature CSVs of the TSN extractor (host code, no kernels): the file format that
// reference src/api/api_load_records.py:41-58 walks with csv.reader + float() per cell, i.e. ~1 ms per
// 1024-d row in Python.  Here the file is mapped, split at line boundaries and parsed by a pool of
// threads with std::from_chars (correctly rounded, locale-independent: the same doubles as Python's
// float()), ~100x faster per core; the caller (ingest.py) hands the arrays to the store upload.
//
// Format: line 1 = header (5 `key =value` fields, returned verbatim); every other non-empty line =
// `clip_no, f0, f1, ... f{dim-1}`; all rows must have the same number of cells.
#include <fcntl.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <charconv>
#include <string>
#include <thread>
#include <vector>

#include "vq_internal.cuh"

namespace {

struct CsvFile {
    const char *p = nullptr;          // the file, mapped read-only
    size_t n = 0;
    std::vector<size_t> line_begin;   // data rows only (header and empty lines skipped)
    std::vector<size_t> line_end;
    size_t header_end = 0;
    ~CsvFile() {
        if (p && n) munmap((void *)p, n);
    }
};

bool load_csv(const char *path, CsvFile *f) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat sb;
    if (fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) {
        close(fd);
        return false;
    }
    f->n = (size_t)sb.st_size;
    if (f->n) {
        void *m = mmap(nullptr, f->n, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m == MAP_FAILED) {
            close(fd);
            f->n = 0;
            return false;
        }
        madvise(m, f->n, MADV_SEQUENTIAL);
        f->p = (const char *)m;
    }
    close(fd);
    const char *p = f->p;
    const size_t n = f->n;
    size_t pos = 0;
    bool first = true;
    while (pos < n) {
        const char *nl = (const char *)memchr(p + pos, '\n', n - pos);
        size_t end = nl ? (size_t)(nl - p) : n;
        const size_t next = end + 1;
        while (end > pos && (p[end - 1] == '\r' || p[end - 1] == ' ')) --end;
        if (first) {
            f->header_end = end;
            first = false;
        } else if (end > pos) {
            f->line_begin.push_back(pos);
            f->line_end.push_back(end);
        }
        pos = next;
    }
    return true;
}

inline const char *skip_blank(const char *p, const char *e) {
    while (p < e && (*p == ' ' || *p == '\t')) ++p;
    return p;
}

long long count_cells(const char *b, const char *e) {
    long long c = 1;
    for (const char *p = b; p < e; ++p) c += (*p == ',');
    return c;
}

// one data row -> clip number + (cells - 1) doubles; returns 0 or the 1-based index of the offending cell
int parse_row(const char *b, const char *e, long long n_cells, int64_t *clip, double *out) {
    const char *p = skip_blank(b, e);
    long long c = 0;
    auto r = std::from_chars(p, e, c);
    if (r.ec != std::errc()) return 1;
    p = skip_blank(r.ptr, e);
    *clip = (int64_t)c;
    for (long long i = 1; i < n_cells; ++i) {
        if (p >= e || *p != ',') return (int)(i + 1);
        p = skip_blank(p + 1, e);
        if (p < e && *p == '+') ++p;                       // float() accepts a leading '+', from_chars does not
        double v = 0.0;
        auto q = std::from_chars(p, e, v);
        if (q.ec == std::errc::result_out_of_range) {      // float() gives +-inf / 0.0 here: let strtod decide
            std::string cell(p, (size_t)(q.ptr - p));
            v = strtod(cell.c_str(), nullptr);
        } else if (q.ec != std::errc()) {
            return (int)(i + 1);
        }
        out[i - 1] = v;
        p = skip_blank(q.ptr, e);
    }
    return p == e ? 0 : (int)n_cells + 1;
}

}  // namespace

extern "C" int vq_csv_shape(const char *path, int64_t *n_rows_out, int32_t *dim_out, char *header_out, int32_t header_cap) {
    VQ_REQUIRE(path && n_rows_out && dim_out, "vq_csv_shape: null argument");
    CsvFile f;
    VQ_REQUIRE(load_csv(path, &f), "vq_csv_shape: cannot read %s", path);
    *n_rows_out = (int64_t)f.line_begin.size();
    *dim_out = f.line_begin.empty()
                   ? 0
                   : (int32_t)(count_cells(f.p + f.line_begin[0], f.p + f.line_end[0]) - 1);
    if (header_out && header_cap > 0) {
        const size_t n = f.header_end < (size_t)header_cap - 1 ? f.header_end : (size_t)header_cap - 1;
        memcpy(header_out, f.p, n);
        header_out[n] = 0;
    }
    return 0;
}

extern "C" int vq_csv_read(const char *path, int32_t n_threads, int64_t cap_rows, int32_t dim, int64_t *clip_numbers_out,
                           double *features_out, int64_t *n_rows_out) {
    VQ_REQUIRE(path && clip_numbers_out && features_out && n_rows_out && dim >= 0 && cap_rows >= 0,
               "vq_csv_read: bad argument");
    CsvFile f;
    VQ_REQUIRE(load_csv(path, &f), "vq_csv_read: cannot read %s", path);
    const long long n = (long long)f.line_begin.size();
    VQ_REQUIRE(n <= cap_rows, "vq_csv_read: %s holds %lld rows, capacity %lld", path, n, (long long)cap_rows);
    *n_rows_out = n;
    if (n == 0) return 0;
    unsigned hw = std::thread::hardware_concurrency();
    long long T = n_threads > 0 ? n_threads : (hw ? (long long)hw : 1);
    if (T > 64) T = 64;
    if (T > n) T = n;
    std::atomic<long long> bad_row(-1);
    std::atomic<int> bad_cell(0);
    const char *base = f.p;
    auto work = [&](long long lo, long long hi) {
        for (long long i = lo; i < hi && bad_row.load(std::memory_order_relaxed) < 0; ++i) {
            const char *b = base + f.line_begin[(size_t)i], *e = base + f.line_end[(size_t)i];
            int rc = count_cells(b, e) == (long long)dim + 1 ? 0 : -1;
            if (rc == 0) rc = parse_row(b, e, (long long)dim + 1, clip_numbers_out + i, features_out + (size_t)i * dim);
            if (rc != 0) {
                long long expect = -1;
                if (bad_row.compare_exchange_strong(expect, i)) bad_cell.store(rc);
            }
        }
    };
    std::vector<std::thread> pool;
    const long long per = (n + T - 1) / T;
    for (long long t = 1; t < T; ++t) {
        const long long lo = t * per, hi = (t + 1) * per < n ? (t + 1) * per : n;
        if (lo < hi) pool.emplace_back(work, lo, hi);
    }
    work(0, per < n ? per : n);
    for (auto &th : pool) th.join();
    const long long br = bad_row.load();
    if (br >= 0) {
        if (bad_cell.load() < 0)
            vq::set_error("vq_csv_read: %s: data row %lld does not have %d cells", path, br + 1, dim + 1);
        else
            vq::set_error("vq_csv_read: %s: data row %lld, cell %d is not a number", path, br + 1, bad_cell.load());
        return -1;
    }
    return 0;
}
