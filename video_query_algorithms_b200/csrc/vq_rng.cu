// Seeded sampling on Python's own generator state (host code, no kernels).
//
// The reference draws its review sets with `random.sample` on the process-wide Mersenne Twister that the broker seeds from
// RANDOM_SEED once per tick (reference src/broker.py:83-84, src/models/ticket.py:333,341).  A review round draws a few
// dozen positions and uses Python's `random` directly.  The finalize round has no size limit (compute_matches.py:78-84):
// `random.sample(M.items(), len(M))` is then a seeded permutation of EVERY match — one interpreter-level loop iteration
// per clip (1 s per million).  vq_mt_sample_range produces the same picks from the same generator state and leaves the
// state exactly where `random.sample(range(n), k)` would have left it, so every later draw of the tick is unchanged.
// The caller moves the state in and out with random.getstate() / random.setstate() (video_query_algorithms_b200/_rng.py).
//
// What is restated here is CPython's documented behaviour (Lib/random.py, 3.2 ... 3.13): sample() picks by repeated
// randbelow(); randbelow(n) takes getrandbits(n.bit_length()) until the value is below n; getrandbits(k <= 32) is the
// top k bits of one 32-bit output, wider values are assembled from 32-bit outputs, low word first.  For a population of
// at most 21 + 4**ceil(log4(3k)) items (k > 5; 21 otherwise) sample() swaps picks out of a pool, beyond that it keeps a
// set of taken indices and redraws on a repeat.
#include <algorithm>
#include <unordered_set>
#include <vector>

#include "vq_internal.cuh"

namespace {

struct MT {
    uint32_t *mt;      // [624]
    int pos;           // next word to hand out, 624 = regenerate first

    void refill() {
        constexpr int N = 624, M = 397;
        constexpr uint32_t kUpper = 0x80000000u, kLower = 0x7fffffffu, kMatrix = 0x9908b0dfu;
        for (int i = 0; i < N; ++i) {
            const uint32_t y = (mt[i] & kUpper) | (mt[(i + 1) % N] & kLower);
            mt[i] = mt[(i + M) % N] ^ (y >> 1) ^ ((y & 1u) ? kMatrix : 0u);
        }
        pos = 0;
    }
    uint32_t next32() {
        if (pos >= 624) refill();
        uint32_t y = mt[pos++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    uint64_t bits(int k) {                                  // 1 <= k <= 64
        if (k <= 32) return next32() >> (32 - k);
        const uint64_t lo = next32();
        const uint64_t hi = next32() >> (64 - k);
        return (hi << 32) | lo;
    }
    uint64_t below(uint64_t n) {                            // n >= 1
        int k = 0;
        for (uint64_t v = n; v; v >>= 1) ++k;
        uint64_t r = bits(k);
        while (r >= n) r = bits(k);
        return r;
    }
};

}  // namespace

extern "C" int vq_mt_sample_range(uint32_t *mt_state /* [624] in/out */, int32_t *mt_pos /* in/out */, int64_t n, int64_t k,
                                  int64_t *picks_out /* [k] */) {
    VQ_REQUIRE(mt_state && mt_pos && (picks_out || k == 0), "vq_mt_sample_range: null argument");
    VQ_REQUIRE(n >= 0 && k >= 0 && k <= n, "vq_mt_sample_range: sample of %lld from a population of %lld", (long long)k, (long long)n);
    VQ_REQUIRE(*mt_pos >= 0 && *mt_pos <= 624, "vq_mt_sample_range: generator position %d outside 0..624", (int)*mt_pos);
    MT g{mt_state, *mt_pos};
    int64_t setsize = 21;
    if (k > 5) {
        int64_t p = 1;                                      // 4 ** ceil(log4(3k)); 3k is never a power of 4
        while (p < 3 * k) p *= 4;
        setsize += p;
    }
    if (n <= setsize) {
        std::vector<int64_t> pool((size_t)n);
        for (int64_t i = 0; i < n; ++i) pool[(size_t)i] = i;
        for (int64_t i = 0; i < k; ++i) {
            const uint64_t j = g.below((uint64_t)(n - i));
            picks_out[i] = pool[(size_t)j];
            pool[(size_t)j] = pool[(size_t)(n - i - 1)];
        }
    } else if (n <= (int64_t)1 << 31) {                     // taken-set as a bitmap
        std::vector<uint64_t> taken((size_t)((n + 63) / 64), 0ull);
        for (int64_t i = 0; i < k; ++i) {
            uint64_t j = g.below((uint64_t)n);
            while (taken[(size_t)(j >> 6)] >> (j & 63) & 1ull) j = g.below((uint64_t)n);
            taken[(size_t)(j >> 6)] |= 1ull << (j & 63);
            picks_out[i] = (int64_t)j;
        }
    } else {
        std::unordered_set<uint64_t> taken;
        for (int64_t i = 0; i < k; ++i) {
            uint64_t j = g.below((uint64_t)n);
            while (taken.count(j)) j = g.below((uint64_t)n);
            taken.insert(j);
            picks_out[i] = (int64_t)j;
        }
    }
    *mt_pos = g.pos;
    return 0;
}
