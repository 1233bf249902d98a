// Host mailbox: an all-gather of small records between the rank processes of one box through a POSIX
// shared-memory segment (host code, no kernels).  It is the host-side twin of the peer-memory exchange kernel
// in vq_exchange.cu: the per-query records of the rank-level path (counts, top-k, tie band, best near miss;
// the sampled list entries of a review round — SURVEY.md §8(e)) are a few KB, and a collective library moves
// them host -> device -> NVLink -> device -> host with two stream synchronisations (~200 us measured per
// collective on 2 B200s); here a rank stores its record and spins on its peers' sequence numbers (~ the skew
// between the ranks).
//
// Protocol.  The segment holds, per rank, one cache line with a sequence number and two data slots.  Call
// number c (1, 2, ...) of a rank: write slot[c & 1][rank], then store-release seq[rank] = c; for every peer
// wait until load-acquire seq[peer] >= c and copy slot[c & 1][peer].  A slot of parity c & 1 is overwritten by
// its owner in call c + 2, which the owner starts only after call c + 1 returned, i.e. after every peer
// published c + 1, which a peer does only after it has copied all of call c: two slots suffice.
// Every rank must make the same sequence of calls with the same sizes (it is a collective).
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <atomic>
#include <new>
#include <string>

#include "vq_internal.cuh"

struct vq_hostx {
    std::string name;
    int world = 0, rank = 0;
    int64_t slot_bytes = 0;
    size_t map_bytes = 0;
    unsigned char *base = nullptr;
    uint64_t calls = 0;
    bool owner = false;
};

namespace {

constexpr size_t kLine = 128;
constexpr uint64_t kMagic = 0x5651484f53545831ull;   // "VQHOSTX1"

struct Header {
    std::atomic<uint64_t> magic;
    int64_t world, slot_bytes;
};

static_assert(std::atomic<uint64_t>::is_always_lock_free, "sequence numbers must be plain 64-bit words");

inline std::atomic<uint64_t> *seq_of(const vq_hostx *x, int r) {
    return reinterpret_cast<std::atomic<uint64_t> *>(x->base + kLine * (1 + r));
}

inline unsigned char *slot_of(const vq_hostx *x, int parity, int r) {
    return x->base + kLine * (1 + x->world) + ((size_t)parity * x->world + r) * (size_t)x->slot_bytes;
}

double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

}  // namespace

extern "C" {

int vq_hostx_create(vq_hostx **out, const char *name, int world, int rank, int64_t slot_bytes) {
    VQ_REQUIRE(out && name && name[0] == '/' && strlen(name) < 200, "vq_hostx_create: name must look like \"/vq-...\"");
    VQ_REQUIRE(world >= 1 && world <= 1024 && rank >= 0 && rank < world, "vq_hostx_create: rank %d of %d", rank, world);
    VQ_REQUIRE(slot_bytes > 0 && slot_bytes <= (int64_t(1) << 26), "vq_hostx_create: slot_bytes %lld outside (0, 64 MiB]",
               (long long)slot_bytes);
    const int64_t slot = (slot_bytes + (int64_t)kLine - 1) / (int64_t)kLine * (int64_t)kLine;
    const size_t bytes = kLine * (1 + (size_t)world) + 2 * (size_t)world * (size_t)slot;
    // rank 0 creates the segment (it must not exist: names are unique per job), the others attach to it.  Callers put
    // a barrier between the two; a rank that is early all the same waits up to 2 s for the segment to appear, to
    // reach its size and (below) to be initialised, instead of failing on a half-made one.
    const double give_up = now_s() + 2.0;
    int fd = -1;
    if (rank == 0) {
        fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
        VQ_REQUIRE(fd >= 0, "vq_hostx_create: shm_open(%s) by rank 0: %s", name, strerror(errno));
        if (ftruncate(fd, (off_t)bytes) != 0) {
            vq::set_error("vq_hostx_create: ftruncate(%zu): %s", bytes, strerror(errno));
            close(fd);
            shm_unlink(name);
            return -1;
        }
    } else {
        struct stat sb;
        for (;;) {
            fd = shm_open(name, O_RDWR, 0600);
            if (fd >= 0 && fstat(fd, &sb) == 0 && sb.st_size != 0) break;
            if (fd >= 0) close(fd);
            if ((fd >= 0 || errno == ENOENT) && now_s() < give_up) {
                fd = -1;
                usleep(1000);
                continue;
            }
            vq::set_error("vq_hostx_create: shm_open(%s) by rank %d: %s", name, rank,
                          fd >= 0 ? "segment stays empty" : strerror(errno));
            return -1;
        }
        if ((size_t)sb.st_size != bytes) {
            vq::set_error("vq_hostx_create: segment %s has %lld bytes, expected %zu (world / slot size differ between "
                          "ranks?)", name, (long long)sb.st_size, bytes);
            close(fd);
            return -1;
        }
    }
    void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    VQ_REQUIRE(p != MAP_FAILED, "vq_hostx_create: mmap: %s", strerror(errno));
    vq_hostx *x = new (std::nothrow) vq_hostx;
    if (!x) {
        munmap(p, bytes);
        vq::set_error("vq_hostx_create: out of memory");
        return -1;
    }
    x->name = name, x->world = world, x->rank = rank, x->slot_bytes = slot, x->map_bytes = bytes;
    x->base = static_cast<unsigned char *>(p), x->owner = rank == 0;
    Header *h = reinterpret_cast<Header *>(x->base);
    if (rank == 0) {                                   // a fresh segment is zero-filled: sequence numbers start at 0
        h->world = world, h->slot_bytes = slot;
        h->magic.store(kMagic, std::memory_order_release);
    } else {
        while (h->magic.load(std::memory_order_acquire) != kMagic && now_s() < give_up) usleep(1000);
        if (h->magic.load(std::memory_order_acquire) != kMagic || h->world != world || h->slot_bytes != slot) {
            vq::set_error("vq_hostx_create: segment %s was not initialised by rank 0 with world %d, slot %lld", name, world,
                          (long long)slot);
            munmap(p, bytes);
            delete x;
            return -1;
        }
    }
    *out = x;
    return 0;
}

// Remove the name (rank 0; the mapping lives on in every process that has it): call once all ranks are attached, so
// that nothing is left in /dev/shm whatever happens to the job later.
int vq_hostx_unlink(vq_hostx *x) {
    VQ_REQUIRE(x, "vq_hostx_unlink: null mailbox");
    if (x->owner) {
        x->owner = false;
        VQ_REQUIRE(shm_unlink(x->name.c_str()) == 0, "vq_hostx_unlink: %s: %s", x->name.c_str(), strerror(errno));
    }
    return 0;
}

int vq_hostx_destroy(vq_hostx *x) {
    if (!x) return 0;
    if (x->owner) shm_unlink(x->name.c_str());
    if (x->base) munmap(x->base, x->map_bytes);
    delete x;
    return 0;
}

int vq_hostx_allgather(vq_hostx *x, const void *mine, int64_t nbytes, void *all_out, double timeout_s) {
    VQ_REQUIRE(x && all_out && (mine || nbytes == 0), "vq_hostx_allgather: null argument");
    VQ_REQUIRE(nbytes >= 0 && nbytes <= x->slot_bytes, "vq_hostx_allgather: %lld bytes do not fit the %lld-byte slots",
               (long long)nbytes, (long long)x->slot_bytes);
    const uint64_t c = ++x->calls;
    const int parity = (int)(c & 1);
    if (nbytes) memcpy(slot_of(x, parity, x->rank), mine, (size_t)nbytes);
    seq_of(x, x->rank)->store(c, std::memory_order_release);
    unsigned char *out = static_cast<unsigned char *>(all_out);
    double deadline = 0.0;
    for (int i = 0; i < x->world; ++i) {
        const int r = (x->rank + i) % x->world;        // own record first, then the peers in ring order
        std::atomic<uint64_t> *s = seq_of(x, r);
        uint64_t spins = 0;
        while (s->load(std::memory_order_acquire) < c) {
            if ((++spins & 0x3ff) == 0) {              // every 1024 polls: look at the clock, let another thread run
                const double t = now_s();
                if (deadline == 0.0) deadline = t + (timeout_s > 0 ? timeout_s : 60.0);
                if (t > deadline) {
                    vq::set_error("vq_hostx_allgather: rank %d waited %.1f s for call %llu of rank %d (a rank died, or the "
                                  "ranks do not make the same calls)", x->rank, timeout_s > 0 ? timeout_s : 60.0,
                                  (unsigned long long)c, r);
                    return -3;
                }
                sched_yield();
            }
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
        }
        if (nbytes) memcpy(out + (size_t)r * (size_t)nbytes, slot_of(x, parity, r), (size_t)nbytes);
    }
    return 0;
}

}  // extern "C"

// ------------------------------------------------------------------------------ per-query summary record
// What one rank tells the others after its local scan (sharded.RankSummary): int64
//   [0] first global row | [1..4] n_match n_near n_tie n_topk | [5 .. 5+k) top-k global rows (-1 padded) |
//   [5+k .. 5+2k) top-k fp32 score bits | 3: best near miss (position in this rank's near-miss list, global row or -1,
//   score bits) | tie_cap tie rows (-1 padded) | tie_cap tie score bits.
// Packing and merging are on the per-query path of every rank (twice per query in Python cost 130 us at 8 ranks).
extern "C" int vq_summary_pack(int64_t first_row, const int64_t *counts3, int32_t k, int32_t n_topk, const int64_t *topk_rows,
                               const float *topk_scores, const int64_t *near_best /* pos, row, score bits; NULL = none */,
                               int32_t tie_cap, int32_t n_ties, const int64_t *tie_rows, const float *tie_scores,
                               int64_t *rec_out /* [5 + 2k + 3 + 2 tie_cap] */) {
    VQ_REQUIRE(counts3 && rec_out && k >= 0 && n_topk >= 0 && n_topk <= k && tie_cap >= 0 && n_ties >= 0,
               "vq_summary_pack: bad argument");
    VQ_REQUIRE((n_topk == 0 || (topk_rows && topk_scores)) && (n_ties == 0 || n_ties > tie_cap || (tie_rows && tie_scores)),
               "vq_summary_pack: null list");
    const uint32_t ninf = 0xff800000u;
    rec_out[0] = first_row;
    rec_out[1] = counts3[0];
    rec_out[2] = counts3[1];
    rec_out[3] = counts3[2];
    rec_out[4] = n_topk;
    for (int i = 0; i < k; ++i) {
        uint32_t bits = ninf;
        if (i < n_topk) memcpy(&bits, &topk_scores[i], 4);
        rec_out[5 + i] = i < n_topk ? topk_rows[i] : -1;
        rec_out[5 + k + i] = (int64_t)bits;
    }
    int64_t *o = rec_out + 5 + 2 * k;
    o[0] = -1; o[1] = -1; o[2] = 0;
    if (near_best && near_best[1] >= 0) { o[0] = near_best[0]; o[1] = near_best[1]; o[2] = near_best[2]; }
    o += 3;
    const bool fits = n_ties <= tie_cap;                 // a longer band rides with the lists; the count says so
    for (int i = 0; i < tie_cap; ++i) {
        uint32_t bits = ninf;
        if (fits && i < n_ties) memcpy(&bits, &tie_scores[i], 4);
        o[i] = (fits && i < n_ties) ? tie_rows[i] : -1;
        o[tie_cap + i] = (int64_t)bits;
    }
    return 0;
}

// gathered [world][rec_len] -> per-rank first rows and counts, the merged top-k (k-way merge of the ranked lists under
// score descending, global row ascending), the best near miss of the search set (highest score; equal scores: the lower
// rank = the earlier rows; position in the search set's near-miss list), the tie band in database order (n_ties_out = -1
// when some rank's band did not fit its record).
extern "C" int vq_summary_merge(const int64_t *gathered, int32_t world, int32_t k, int32_t tie_cap, int64_t *first_rows_out,
                                int64_t *counts_out /* [world][3] */, int64_t *topk_rows_out, float *topk_scores_out,
                                int32_t *n_topk_out, int64_t *best_out /* pos, row (-1 none), score bits */,
                                int64_t *tie_rows_out /* [world * tie_cap] */, float *tie_scores_out, int64_t *n_ties_out) {
    VQ_REQUIRE(gathered && world >= 1 && k >= 0 && tie_cap >= 0 && first_rows_out && counts_out && n_topk_out && best_out &&
               n_ties_out && (k == 0 || (topk_rows_out && topk_scores_out)), "vq_summary_merge: bad argument");
    const size_t len = (size_t)5 + 2 * (size_t)k + 3 + 2 * (size_t)tie_cap;
    auto score_of = [](int64_t bits) { const uint32_t u = (uint32_t)bits; float f; memcpy(&f, &u, 4); return f; };
    int head[64];
    VQ_REQUIRE(world <= 64, "vq_summary_merge: at most 64 ranks");
    bool ties_fit = true;
    for (int r = 0; r < world; ++r) {
        const int64_t *g = gathered + (size_t)r * len;
        first_rows_out[r] = g[0];
        counts_out[3 * r] = g[1];
        counts_out[3 * r + 1] = g[2];
        counts_out[3 * r + 2] = g[3];
        VQ_REQUIRE(g[4] >= 0 && g[4] <= k, "vq_summary_merge: rank %d reports %lld top-k entries of %d", r, (long long)g[4], k);
        head[r] = 0;
        if (g[3] > tie_cap) ties_fit = false;
    }
    int n = 0;
    for (; n < k; ++n) {
        int best = -1;
        float bs = 0.f;
        int64_t br = 0;
        for (int r = 0; r < world; ++r) {
            const int64_t *g = gathered + (size_t)r * len;
            if (head[r] >= (int)g[4]) continue;
            const float s = score_of(g[5 + k + head[r]]);
            const int64_t row = g[5 + head[r]];
            if (best < 0 || s > bs || (s == bs && row < br)) { best = r; bs = s; br = row; }
        }
        if (best < 0) break;
        ++head[best];
        topk_rows_out[n] = br;
        topk_scores_out[n] = bs;
    }
    *n_topk_out = n;
    best_out[0] = -1; best_out[1] = -1; best_out[2] = 0;
    int64_t base = 0;
    float best_score = 0.f;
    for (int r = 0; r < world; ++r) {
        const int64_t *o = gathered + (size_t)r * len + 5 + 2 * (size_t)k;
        if (o[1] >= 0) {
            const float s = score_of(o[2]);
            if (best_out[1] < 0 || s > best_score) {
                best_out[0] = base + o[0];
                best_out[1] = o[1];
                best_out[2] = o[2];
                best_score = s;
            }
        }
        base += gathered[(size_t)r * len + 2];
    }
    *n_ties_out = -1;
    if (ties_fit) {
        int64_t m = 0;
        for (int r = 0; r < world; ++r) {
            const int64_t *g = gathered + (size_t)r * len;
            const int64_t *o = g + 5 + 2 * (size_t)k + 3;
            for (int64_t i = 0; i < g[3]; ++i) {
                if (tie_rows_out) tie_rows_out[m] = o[i];
                if (tie_scores_out) tie_scores_out[m] = score_of(o[tie_cap + i]);
                ++m;
            }
        }
        *n_ties_out = m;
    }
    return 0;
}
