// sda_device.cuh -- device building blocks of the dual-annealing hot path (sm_100a).
//
// Mapping: ONE CHAIN PER WARP.  Coordinates are striped over the 32 lanes (coordinate k lives in
// lane k & 31), chain vectors live in shared memory, scalars (energies, counters) are replicated
// in the registers of every lane and stay warp-uniform, reductions are xor-butterflies so every
// lane ends with the same bits.
//
// Reference citations are to /root/reference/sdaopt/_sda.py unless another file is named.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "sdaopt_b200.h"   // include/sdaopt_b200.h (-I include; shipped next to csrc/ in a wheel)
#include "sda_math.cuh"

#define SDA_WARP 32
#define SDA_FULL_MASK 0xffffffffu

namespace sda {

constexpr double kBigValue = 1e16;        // BIG_VALUE            :17
constexpr double kTailLimit = 1.e8;       // tail_limit           :27
constexpr double kMinVisitBound = 1.e-10; // min_visit_bound      :28
constexpr int kMaxReinit = 1000;          // MAX_REINIT_COUNT     :116
constexpr double kTwoPi = 6.283185307179586476925286766559;  // 2*np.pi rounds to this double
constexpr double kPi = 3.141592653589793238462643383279;

enum Purpose : uint32_t { kPolar = 0, kClip = 1, kAccept = 2, kInit = 3 };

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  Counter-based: any draw is addressable, so
// the lanes of a warp sample their coordinates independently (K2).
// ---------------------------------------------------------------------------------------------
struct Philox {
    uint32_t k0, k1;
    __device__ __forceinline__ void block(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                          uint32_t (&o)[4]) const
    {
        uint32_t a0 = k0, a1 = k1;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            // one IMAD.WIDE per product: both halves come out of the same 32x32->64 multiply
            const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
            const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
            const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
            const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
            c0 = hi1 ^ c1 ^ a0;
            c1 = lo1;
            c2 = hi0 ^ c3 ^ a1;
            c3 = lo0;
            a0 += 0x9E3779B9u;
            a1 += 0xBB67AE85u;
        }
        o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
    }
};

// numpy RandomState.random_sample() construction (genrand_res53)
__device__ __forceinline__ double u53(uint32_t a, uint32_t b)
{
    // ((a>>5)*2^26 + (b>>6)) / 2^53, exactly, without a 64-bit int->double conversion (XU pipe):
    // 0x43300000:w is the double 2^52 + w, so subtracting 2^52 yields w exactly (FP64 pipe), and
    // hi*2^-27 + lo*2^-53 has at most 53 significant bits, so the fma is exact too.
    const double hi = __hiloint2double(0x43300000, (int)(a >> 5)) - 4503599627370496.0;
    const double lo = __hiloint2double(0x43300000, (int)(b >> 6)) - 4503599627370496.0;
    return fma(hi, 7.450580596923828125e-09, lo * 1.1102230246251565404e-16);
}

// ---------------------------------------------------------------------------------------------
// Chain groups.  A chain is advanced by G consecutive lanes (G = 1, 2, 4, ... 32, chosen from the
// dimension so that ceil(D/G)*G wastes few lanes); a warp carries 32/G chains side by side.  All
// synchronisation inside group code uses the group's lane mask, so groups of one warp may diverge
// (local search, re-initialisation, early exit) without deadlock; the warp-level loop re-aligns
// them with full-mask barriers between phases so that they share instruction issue.
// ---------------------------------------------------------------------------------------------
struct Group {
    int G;          // lanes per chain
    int gl;         // lane index inside the group: owns coordinates gl, gl + G, ...
    int gid;        // group index inside the warp
    unsigned mask;  // lanes of this group
};

__device__ __forceinline__ Group make_group(int G, int lane)
{
    Group g;
    g.G = G;
    g.gl = lane & (G - 1);
    g.gid = lane / G;
    g.mask = (G == 32) ? SDA_FULL_MASK : (((1u << G) - 1u) << (g.gid * G));
    return g;
}

// reductions over the lanes of a group (xor butterfly: every lane ends with identical bits)
__device__ __forceinline__ double group_sum(double v, const Group &g)
{
    for (int o = g.G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(g.mask, v, o);
    return v;
}
__device__ __forceinline__ double group_prod(double v, const Group &g)
{
    for (int o = g.G >> 1; o > 0; o >>= 1) v *= __shfl_xor_sync(g.mask, v, o);
    return v;
}
__device__ __forceinline__ void group_sync(const Group &g) { __syncwarp(g.mask); }

// ---------------------------------------------------------------------------------------------
// warp reductions: every lane receives bit-identical results
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SDA_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_prod(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v *= __shfl_xor_sync(SDA_FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(SDA_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(SDA_FULL_MASK, v, o));
    return v;
}
__device__ __forceinline__ double warp_bcast(double v, int src)
{
    return __shfl_sync(SDA_FULL_MASK, v, src);
}

// ---------------------------------------------------------------------------------------------
// K2: visiting distribution.  The arithmetic follows the reference operation by operation with
// explicit round-to-nearest intrinsics (no FMA contraction) so that the only differences from
// the reference are the last-bit differences of log/exp themselves.
// ---------------------------------------------------------------------------------------------
struct VisitConst {
    double sigmax;   // sigmax(T), K1 table                               :76-86
    double qv_m1;    // qv - 1.0
    double three_mq; // 3.0 - qv
    double c;        // (qv - 1) / (3 - qv), the exponent of |y| in the denominator
};

// gaussian_fn(1) / gaussian_fn(0) + visit_fn from one accepted polar pair   :87-91, :103-107
__device__ __forceinline__ double visit_from_polar(double xg, double yg, double s,
                                                   const VisitConst &vc)
{
    const double root = sqrt(__dmul_rn(__ddiv_rn(-2.0, s), log(s)));   // :103-104
    const double x = __dmul_rn(vc.sigmax, __dmul_rn(root, yg));        // :87, :105
    const double y = __dmul_rn(root, xg);                              // :88, :107
    const double den = exp(__ddiv_rn(__dmul_rn(vc.qv_m1, log(fabs(y))), vc.three_mq)); // :89-90
    return __ddiv_rn(x, den);                                          // :91
}

// Production (Philox) form of the same sample: visit = sigmax * g1 * |g0|^-c with
// c = (qv-1)/(3-qv), i.e. x / exp(c ln|y|) = x * exp(-c ln|y|).  Algebraically identical to
// visit_from_polar, one IEEE division instead of three; the results differ from the reference
// order by a few ulp (the same size as the exp/log differences between math libraries), so the
// tape-replay path keeps visit_from_polar and this form is validated against the oracle on the
// same Philox keys (tests/test_gpu_parity.py::test_philox_chains_match_oracle).
__device__ __forceinline__ double visit_from_polar_fast(double xg, double yg, double s,
                                                        const VisitConst &vc)
{
    const double root = sqrt_normal(div_normal(-2.0 * log_pos(s), s));     // 0 < s < 1, s >= 2^-106
    const double y = root * xg;
    return (vc.sigmax * (root * yg)) * exp_any(-vc.c * log_pos(fabs(y)));
}

// C fmod(a, R) for R > 0, bit-exact, in ~8 FP64 instructions instead of libdevice's long-division
// loop.  q = rint(|a|/R) (magic-number rounding), t = fma(-q, R, |a|) is the exact remainder as
// long as q < 2^50: q carries a relative error <= 2^-52, so |t| < R and t is a multiple of
// ulp(R) -> representable, the fma does not round.  fmod takes the sign of a.
__device__ __forceinline__ double fmod_exact(double a, double R, double invR)
{
    const double aa = fabs(a);
    if (aa < R) return a;
    const double q = aa * invR;
    if (!(q < 1125899906842624.0)) return cold_fmod(a, R);      // 2^50: fall back (also inf/nan)
    const double qi = __dadd_rn(__dadd_rn(q, 6755399441055744.0), -6755399441055744.0);
    double t = __fma_rn(-qi, R, aa);
    if (t < 0.0) t = __dadd_rn(t, R);
    else if (t >= R) t = __dsub_rn(t, R);
    return copysign(t, a);
}

// a = x - lower; b = fmod(a, R) + R; x = fmod(b, R) + lower; nudge      :50-55 / :65-72
// invR = 1/R.  Every rounding of the reference's sequence is kept: fmod is exact by definition,
// b = fmod(a,R) + R rounds once, and b lies in (0, 2R] so the second fmod is a conditional subtract.
__device__ __forceinline__ double wrap_coord(double xv, double lower, double range, double inv_range)
{
    const double a = __dsub_rn(xv, lower);
    const double b = __dadd_rn(fmod_exact(a, range, inv_range), range);
    double m = b;
    if (m >= range) m = __dsub_rn(m, range);
    if (m >= range) m = __dsub_rn(m, range);
    if (!(b <= 2.0 * range && b >= 0.0)) m = cold_fmod(b, range);   // not reachable for finite inputs
    double r = __dadd_rn(m, lower);
    if (fabs(__dsub_rn(r, lower)) < kMinVisitBound) r = __dadd_rn(r, 1.e-10);
    return r;
}

// Uniform sources ------------------------------------------------------------------------------
// Tape: sequential replay of recorded RandomState draws; every lane walks the same tape.
struct TapeRng {
    const double *tape;
    int64_t pos, len;
    bool exhausted;
    __device__ __forceinline__ double next()
    {
        if (pos >= len) { exhausted = true; return 0.25; }
        return tape[pos++];
    }
};

// Philox addressing context (see include/sdaopt_b200.h "RNG protocol")
struct PhiloxRng {
    Philox ph;
    uint32_t iter, j;
    __device__ __forceinline__ void pair(uint32_t purpose, uint32_t k, uint32_t attempt, double &u1,
                                         double &u2) const
    {
        uint32_t o[4];
        ph.block(k, attempt, (purpose << 24) | j, iter, o);
        u1 = u53(o[0], o[1]);
        u2 = u53(o[2], o[3]);
    }
};

// accepted polar pair for coordinate k, Philox source                                :94-102
__device__ __forceinline__ void polar_pair_philox(const PhiloxRng &rng, uint32_t k, double &xg,
                                                  double &yg, double &s)
{
    uint32_t attempt = 0;
    do {
        double u1, u2;
        rng.pair(kPolar, k, attempt++, u1, u2);
        xg = __dsub_rn(__dmul_rn(u1, 2.0), 1.0);
        yg = __dsub_rn(__dmul_rn(u2, 2.0), 1.0);
        s = __dadd_rn(__dmul_rn(xg, xg), __dmul_rn(yg, yg));
    } while (s <= 0.0 || s >= 1.0);
}

// Accepted polar pairs for all coordinates of one lane (k = first, first + stride, ... < dim), Philox
// source.  The acceptance loop is flattened over the lane's coordinates: every trip makes ONE
// attempt for the lane's current coordinate; an accepted pair is stored (xs[k] = X, ys[k] = Y) and
// the lane moves on, a rejected one bumps the attempt counter.  A warp therefore runs
// max-over-lanes(total attempts) trips (~1.7 per coordinate at 7 coordinates per lane) instead of
// the sum over coordinates of max-over-lanes(attempts) (~3.1 per coordinate) of a per-coordinate
// loop.  Same counters, same "first accepted attempt" rule -> same samples as polar_pair_philox.
// per_coord_j: the one-coordinate phase addresses draw k with j = dim + k.
template <bool PER_COORD_J>
__device__ __forceinline__ void polar_fill_philox(PhiloxRng rng, int first, int stride, int dim,
                                                  double *xs, double *ys)
{
    int k = first;
    uint32_t attempt = 0;
#pragma unroll 1
    while (k < dim) {
        if (PER_COORD_J) rng.j = (uint32_t)(dim + k);
        double u1, u2;
        rng.pair(kPolar, (uint32_t)k, attempt, u1, u2);
        const double xg = __dsub_rn(__dmul_rn(u1, 2.0), 1.0);
        const double yg = __dsub_rn(__dmul_rn(u2, 2.0), 1.0);
        const double s = __dadd_rn(__dmul_rn(xg, xg), __dmul_rn(yg, yg));
        if (s > 0.0 && s < 1.0) {
            xs[k] = xg;
            ys[k] = yg;
            k += stride;
            attempt = 0;
        } else {
            attempt++;
        }
    }
}

// one visit_fn(T) draw for coordinate k, Philox source: polar rejection loop        :94-102
__device__ __forceinline__ double visit_sample_philox(const PhiloxRng &rng, uint32_t k,
                                                      const VisitConst &vc)
{
    double xg, yg, s;
    uint32_t attempt = 0;
    do {
        double u1, u2;
        rng.pair(kPolar, k, attempt++, u1, u2);
        xg = __dsub_rn(__dmul_rn(u1, 2.0), 1.0);
        yg = __dsub_rn(__dmul_rn(u2, 2.0), 1.0);
        s = __dadd_rn(__dmul_rn(xg, xg), __dmul_rn(yg, yg));
    } while (s <= 0.0 || s >= 1.0);
    return visit_from_polar_fast(xg, yg, s, vc);
}

// same from the tape (uniform across the warp)
__device__ __forceinline__ double visit_sample_tape(TapeRng &rng, const VisitConst &vc)
{
    double xg, yg, s;
    do {
        const double u1 = rng.next();
        const double u2 = rng.next();
        if (rng.exhausted) return 0.0;
        xg = __dsub_rn(__dmul_rn(u1, 2.0), 1.0);
        yg = __dsub_rn(__dmul_rn(u2, 2.0), 1.0);
        s = __dadd_rn(__dmul_rn(xg, xg), __dmul_rn(yg, yg));
    } while (s <= 0.0 || s >= 1.0);
    return visit_from_polar(xg, yg, s, vc);
}

}  // namespace sda
