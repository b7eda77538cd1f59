// Philox4x32-10 counter-based generator (Salmon et al., SC'11), written out by hand.
// Key = 64-bit seed; counter = (particle id lo, particle id hi, stage, (sweep<<8)|slot), so a draw
// depends only on (seed, global particle id, stage, sweep, slot) and never on how particles are
// sharded over GPUs or threads.
#pragma once
#include <stdint.h>

#define PHILOX_M0 0xD2511F53u
#define PHILOX_M1 0xCD9E8D57u
#define PHILOX_W0 0x9E3779B9u
#define PHILOX_W1 0xBB67AE85u

#define SMCB_SLOT_UNIFORM 255u   // slot of the accept-step uniform; normals use slots 0..127

struct Philox4 {
    uint32_t x, y, z, w;
};

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                 uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    Philox4 o = {c0, c1, c2, c3};
    return o;
}

// 53-bit uniform in [0,1) from two 32-bit words
__host__ __device__ inline double u53(uint32_t hi, uint32_t lo) {
    const uint64_t v = (((uint64_t)hi << 32) | lo) >> 11;
    return (double)v * (1.0 / 9007199254740992.0);
}

__device__ inline Philox4 philox_draw(uint64_t seed, uint64_t id, uint32_t stage, uint32_t sweep, uint32_t slot) {
    return philox4x32_10((uint32_t)id, (uint32_t)(id >> 32), stage, (sweep << 8) | slot, (uint32_t)seed,
                         (uint32_t)(seed >> 32));
}

// two standard normals from one Philox block (Box-Muller, FP64)
__device__ inline void philox_normal2(uint64_t seed, uint64_t id, uint32_t stage, uint32_t sweep, uint32_t slot,
                                      double* z0, double* z1) {
    const Philox4 r = philox_draw(seed, id, stage, sweep, slot);
    const double u1 = 1.0 - u53(r.x, r.y);   // (0,1]
    const double u2 = u53(r.z, r.w);
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    *z0 = rad * c;
    *z1 = rad * s;
}

__device__ inline double philox_uniform(uint64_t seed, uint64_t id, uint32_t stage, uint32_t sweep, uint32_t slot) {
    const Philox4 r = philox_draw(seed, id, stage, sweep, slot);
    return u53(r.x, r.y);
}
