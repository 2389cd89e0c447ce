// wrh_rng.hpp — host random source for SCENE CONSTRUCTION only (balls / rtw_final / synthetic draw sphere positions
// and materials from `rand`, scene.zig:99-131,440,486).  The reference uses std.Random.DefaultPrng seeded from
// getrandom (rng.zig:16-26); here the seed is explicit so a scene can be rebuilt bit for bit.
// Restated from the published algorithms (Xoshiro256++ by Blackman & Vigna, SplitMix64 seeding, and the Zig
// standard library's float / uintLessThan constructions); the Zig std sources are not part of the reference tree.
#pragma once

#include <cstdint>
#include <cstring>

#include "wrh_math.hpp"

namespace wrh {

class Random {
   public:
    explicit Random(uint64_t seed) {
        uint64_t sm = seed;
        for (auto& w : s_) w = splitmix(sm);
    }

    uint64_t next() {
        const uint64_t result = rotl(s_[0] + s_[3], 23) + s_[0];
        const uint64_t t = s_[1] << 17;
        s_[2] ^= s_[0];
        s_[3] ^= s_[1];
        s_[1] ^= s_[2];
        s_[0] ^= s_[3];
        s_[2] ^= t;
        s_[3] = rotl(s_[3], 45);
        return result;
    }

    // Random.float(f64): 52 mantissa bits, exponent from the leading zeros of the remaining bits
    Real floatReal() {
        const uint64_t r = next();
        unsigned lz = r ? static_cast<unsigned>(__builtin_clzll(r)) : 64u;
        if (lz >= 12) {
            lz = 12;
            for (;;) {
                const uint64_t more = next();
                const unsigned add = more ? static_cast<unsigned>(__builtin_clzll(more)) : 64u;
                lz += add;
                if (add != 64) break;
                if (lz >= 1022) { lz = 1022; break; }
            }
        }
        const uint64_t bits = (static_cast<uint64_t>(1022 - lz) << 52) | (r & 0xFFFFFFFFFFFFFull);
        Real out;
        std::memcpy(&out, &bits, sizeof out);
        return out;
    }

    // Random.intRangeAtMost(usize, 0, n - 1): Lemire's method
    uint32_t below(uint32_t n) {
        const uint64_t bound = n;
        uint64_t x = next();
        __uint128_t m = static_cast<__uint128_t>(x) * bound;
        uint64_t l = static_cast<uint64_t>(m);
        if (l < bound) {
            const uint64_t t = (0 - bound) % bound;
            while (l < t) {
                x = next();
                m = static_cast<__uint128_t>(x) * bound;
                l = static_cast<uint64_t>(m);
            }
        }
        return static_cast<uint32_t>(m >> 64);
    }

    Vec3 sampleVec3() {  // rng.zig:35-41
        const Real a = floatReal(), b = floatReal(), c = floatReal();
        return {a, b, c};
    }
    Vec3 sampleVec3Interval(Interval range) {  // rng.zig:43-49
        const Real a = floatReal() * range.size() + range.min;
        const Real b = floatReal() * range.size() + range.min;
        const Real c = floatReal() * range.size() + range.min;
        return {a, b, c};
    }

   private:
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    static uint64_t splitmix(uint64_t& state) {
        state += 0x9e3779b97f4a7c15ull;
        uint64_t z = state;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    uint64_t s_[4];
};

}  // namespace wrh
