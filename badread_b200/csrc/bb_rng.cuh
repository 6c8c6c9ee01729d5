// bb_rng.cuh — counter-based per-read random streams and the reference's samplers on top of them.
//
// The reference draws everything from one sequential MT19937 (`random` module, simulate.py:35); per-read
// consumption is data dependent, so that stream cannot be reproduced in parallel.  The CUDA path keeps the
// reference's SAMPLERS bit for bit (Random._randbelow_with_getrandbits, random(), choices() as used by
// misc.py:156-182, error_model.py:135-176, simulate.py:294,338, qscore_model.py:283) and feeds them from
// Philox4x32-10 streams keyed by (seed, read index, purpose, index):
//     key = (seed_lo, seed_hi);  counter = (index, purpose<<24 | block, read_lo, read_hi)
// Each block yields four 32-bit words consumed in order.  oracle/badread_oracle.c (mode "philox") uses the
// identical layout, which is what makes GPU output byte-identical to the CPU oracle.
#pragma once
#include <cstdint>

#define BB_PURPOSE_PAD 2u     // 2k pad bases (simulate.py:260)
#define BB_PURPOSE_LOOP 3u    // one stream per k-mer loop iteration (simulate.py:294-296), index = loop_count-1
#define BB_PURPOSE_WINDOW 4u  // window position (simulate.py:338), index = alignment ordinal
#define BB_PURPOSE_QSCORE 5u  // one stream per base of the untrimmed read (qscore_model.py:283), index = base

struct BBRng {
    uint32_t k0, k1;          // seed
    uint32_t r0, r1;          // read index
    uint32_t c0, c1;          // index, purpose<<24 | block
    uint32_t b0, b1, b2, b3;  // current block
    int pos;

    __device__ __forceinline__ void init(uint64_t seed, uint64_t read) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        r0 = (uint32_t)read; r1 = (uint32_t)(read >> 32);
        c0 = 0; c1 = 0; pos = 4; b0 = b1 = b2 = b3 = 0;
    }
    __device__ __forceinline__ void stream(uint32_t purpose, uint32_t index) {
        c0 = index; c1 = purpose << 24; pos = 4;
    }
    __device__ __forceinline__ void refill() {
        uint32_t x0 = c0, x1 = c1, x2 = r0, x3 = r1, ka = k0, kb = k1;
#pragma unroll
        for (int i = 0; i < 10; i++) {
            const uint32_t hi0 = __umulhi(0xD2511F53u, x0), lo0 = 0xD2511F53u * x0;
            const uint32_t hi1 = __umulhi(0xCD9E8D57u, x2), lo1 = 0xCD9E8D57u * x2;
            const uint32_t n0 = hi1 ^ x1 ^ ka, n2 = hi0 ^ x3 ^ kb;
            x0 = n0; x1 = lo1; x2 = n2; x3 = lo0;
            ka += 0x9E3779B9u; kb += 0xBB67AE85u;
        }
        b0 = x0; b1 = x1; b2 = x2; b3 = x3;
        c1++; pos = 0;
    }
    __device__ __forceinline__ uint32_t next() {
        if (pos >= 4) refill();
        const uint32_t v = pos == 0 ? b0 : pos == 1 ? b1 : pos == 2 ? b2 : b3;
        pos++;
        return v;
    }
    // Random._randbelow_with_getrandbits: k = n.bit_length(); r = getrandbits(k); while r >= n: redraw
    __device__ __forceinline__ uint32_t randbelow(uint32_t n) {
        const int k = 32 - __clz(n);
        uint32_t v = next() >> (32 - k);
        while (v >= n) v = next() >> (32 - k);
        return v;
    }
    // random.random(): (a>>5, b>>6) -> (a*2^26+b)/2^53
    __device__ __forceinline__ double random() {
        const uint32_t a = next() >> 5, b = next() >> 6;
        return __dmul_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 1.0 / 9007199254740992.0);
    }
    // misc.get_random_base: 'ACGT'[randint(0,3)]  (3-bit draws, half rejected)
    __device__ __forceinline__ uint8_t random_base() {
        const uint32_t v = randbelow(4);
        return v == 0 ? 'A' : v == 1 ? 'C' : v == 2 ? 'G' : 'T';
    }
    __device__ __forceinline__ uint8_t random_different_base(uint8_t b) {
        uint8_t x = random_base();
        while (x == b) x = random_base();
        return x;
    }
};

// random.choices(pop, weights)[0] given cum = list(accumulate(weights)): bisect_right(cum, random()*cum[-1], 0, n-1)
__device__ __forceinline__ int bb_choices(BBRng &rng, const double *__restrict__ cum, int n) {
    const double x = __dmul_rn(rng.random(), cum[n - 1]);
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (x < cum[mid]) hi = mid; else lo = mid + 1;
    }
    return lo;
}
