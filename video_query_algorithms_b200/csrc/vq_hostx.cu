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
