// Counter-based normal stream: Philox4x32-10 + Box-Muller, entirely in registers.
//
// Stream addressing (shared with oracle/philox.py, bit for bit on the uniforms):
//   key     = (low 32 bits of seed, low 32 bits of stream)
//   counter = (path_lo, path_hi, block, kind)      kind 0: normals, 1: uniforms
//   normal  #n of a path = element (n & 1) of BoxMuller(block n >> 1):
//             sqrt(-2 log u1) (cos, sin)(2 pi u2), u1 = (k1 + 1/2) 2^-52, u2 = k2 2^-52,
//             k = (low 20 bits of the first word):(second word) of a word pair
//   uniform #m of a path = element (m & 1) of the two 52-bit uniforms of block m >> 1
// so a draw depends only on (seed, stream, global path id, draw index): sharding paths
// over GPUs or replaying a path in a second pass reproduces it exactly.  Takes the place
// of torch.manual_seed(42|43) + torch.randn(N, d) per sub-step (src/engine/engine.py:25,
// src/models/model.py:47).
#pragma once
#include <cstdint>
#ifdef MCRE_FAST_MATH
#include "fastmath.cuh"
#endif

namespace mcre {

struct Philox {
  uint32_t k0, k1;
  __host__ __device__ static inline void round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
                                               uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    // one 32x32 -> 64 multiply per product (IMAD.WIDE.U32) instead of separate high / low halves
    const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
  }
  __host__ __device__ inline void operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                             uint32_t out[4]) const {
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      round(c0, c1, c2, c3, a, b);
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

// 52-bit uniform in (0,1): (k + 0.5) * 2^-52 with k = (low 20 bits of hi):(lo).  Built in the
// mantissa of a double in [1,2) with ONE logic instruction (the exponent is OR-ed over the 12 unused
// bits of hi, lo is the low word as it is) and shifted down with one exact subtraction - no
// int -> double conversion, no funnel shifts.  (Round 1 used the top 52 bits of hi:lo, 4 instructions.)
__host__ __device__ inline double u52(uint32_t hi, uint32_t lo) {
#ifdef __CUDA_ARCH__
  const double d = __hiloint2double((int)(0x3ff00000u | (hi & 0x000fffffu)), (int)lo);
  return d - 0.99999999999999988898;   // 1 - 2^-53: exact, result (2k + 1) * 2^-53
#else
  uint64_t k = (((uint64_t)(hi & 0x000fffffu)) << 32) | lo;
  return ((double)k + 0.5) * 2.220446049250313e-16;
#endif
}

// mantissa double in [1, 2) from the same 52 bits: 1 + k 2^-52
__device__ inline double mant52(uint32_t hi, uint32_t lo) {
  return __hiloint2double((int)(0x3ff00000u | (hi & 0x000fffffu)), (int)lo);
}

struct RngDev {
  int mode;             // MCRE_RNG_*
  uint32_t k0, k1;
  const double *z;      // injected normals [n_sub][n_total][d]
  const double *u;      // injected uniforms [n_sub][n_total]
  long long n_total;
};

// Sequential normal stream of one path.  All threads of a warp advance in lockstep, so
// the spare-normal branch is uniform.
struct NormalStream {
  Philox ph;
  uint32_t p_lo, p_hi;
  uint32_t n;       // next normal index
  double spare;
  __device__ inline void init(const RngDev &r, unsigned long long gpath, uint32_t first = 0) {
    ph.k0 = r.k0; ph.k1 = r.k1;
    p_lo = (uint32_t)gpath; p_hi = (uint32_t)(gpath >> 32);
    n = first; spare = 0.0;
  }
  __device__ inline void pair(uint32_t block, double &z0, double &z1) const {
    uint32_t o[4];
    ph(p_lo, p_hi, block, 0u, o);
    // radius from u1 = (k1 + 1/2) 2^-52 in (0,1); angle from u2 = k2 2^-52 in [0,1)
    const double u1 = u52(o[0], o[1]);
#if defined(MCRE_FAST_MATH) && MCRE_FAST_MATH >= 2
    // same arithmetic as NormalStreamV::next2 (the angle reduction works on the mantissa double)
    const double dd[1] = {mant52(o[2], o[3])};
    const double uu[1] = {u1};
    double lg[1];
    fm_neg2log_tv<1>(uu, lg);
    const double rr[1] = {fm_sqrt_pos(lg[0])};   // u1 < 1: the argument is strictly positive
    double zc[1], zs[1];
    fm_polar_tv<1>(dd, rr, zc, zs);
    z0 = zc[0]; z1 = zs[0];
#else
    const double u2 = mant52(o[2], o[3]) - 1.0;
    double s, c;
#if defined(MCRE_FAST_MATH)
    const double rad = fm_sqrt(-2.0 * fm_log(u1));
    fm_sincos2pi(u2, s, c);
#else
    const double rad = sqrt(-2.0 * log(u1));
    sincospi(2.0 * u2, &s, &c);
#endif
    z0 = rad * c; z1 = rad * s;
#endif
  }
  __device__ inline double next() {
    double z;
    if ((n & 1u) == 0u) { pair(n >> 1, z, spare); } else { z = spare; }
    ++n;
    return z;
  }
  // two consecutive normals starting at an even index (noise_dim 2 fast path)
  __device__ inline void next2(double &z0, double &z1) { pair(n >> 1, z0, z1); n += 2; }
  // both uniforms of block `block` (uniform #2 block and #2 block + 1)
  __device__ inline void uniform_pair(uint32_t block, double &ua, double &ub) const {
    uint32_t o[4];
    ph(p_lo, p_hi, block, 1u, o);
    ua = u52(o[0], o[1]); ub = u52(o[2], o[3]);
  }
  // both uniforms of block `block` of stream kind `kind` (kind 2: Brownian-bridge draws of barrier options)
  __device__ inline void uniform_pair_kind(uint32_t block, uint32_t kind, double &ua, double &ub) const {
    uint32_t o[4];
    ph(p_lo, p_hi, block, kind, o);
    ua = u52(o[0], o[1]); ub = u52(o[2], o[3]);
  }
  __device__ inline double uniform(uint32_t m) const {
    uint32_t o[4];
    ph(p_lo, p_hi, m >> 1, 1u, o);
    return (m & 1u) ? u52(o[2], o[3]) : u52(o[0], o[1]);
  }
};


#if defined(MCRE_FAST_MATH) && MCRE_FAST_MATH >= 2
// PP paths of one thread advanced in lock-step: the ten Philox rounds and every Box-Muller
// operation are interleaved over the paths in source order, so a warp always has PP independent
// instructions in flight (see the note on MCRE_VP in fastmath.cuh).  Same streams as NormalStream.
template <int PP>
struct NormalStreamV {
  uint32_t k0, k1;
  uint32_t p_lo[PP], p_hi[PP];
  uint32_t n;   // next normal index (even)
  __device__ inline void init(const RngDev &r, const long long (&gpath)[PP]) {
    k0 = r.k0; k1 = r.k1; n = 0;
    MCRE_VP { p_lo[p] = (uint32_t)gpath[p]; p_hi[p] = (uint32_t)((unsigned long long)gpath[p] >> 32); }
  }
  // the integer half of next2: ten Philox rounds for the block of the next normal pair of every path.
  // Split from the floating-point half so that a kernel can issue the rounds of step s+1 next to the FP64
  // work of step s (the FP64 pipe and the integer pipe then overlap inside one warp).
  __device__ inline void raw(uint32_t (&c0)[PP], uint32_t (&c1)[PP], uint32_t (&c2)[PP], uint32_t (&c3)[PP]) {
    const uint32_t block = n >> 1;
    MCRE_VP { c0[p] = p_lo[p]; c1[p] = p_hi[p]; c2[p] = block; c3[p] = 0u; }
    uint32_t a = k0, b = k1;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
      MCRE_VP Philox::round(c0[p], c1[p], c2[p], c3[p], a, b);
      a += 0x9E3779B9u; b += 0xBB67AE85u;
    }
    n += 2;
  }
  // the floating-point half: Box-Muller on the raw words
  __device__ static inline void box_muller(const uint32_t (&c0)[PP], const uint32_t (&c1)[PP], const uint32_t (&c2)[PP],
                                           const uint32_t (&c3)[PP], double (&z0)[PP], double (&z1)[PP]) {
    double u1[PP], d2[PP], lg[PP], rad[PP];
    MCRE_VP u1[p] = u52(c0[p], c1[p]);
    MCRE_VP d2[p] = mant52(c2[p], c3[p]);
    fm_neg2log_tv<PP>(u1, lg);
    fm_sqrt_posv<PP>(lg, rad);          // u1 < 1: the argument is strictly positive
    fm_polar_tv<PP>(d2, rad, z0, z1);
  }
  // two consecutive normals of every path (noise_dim 2)
  __device__ inline void next2(double (&z0)[PP], double (&z1)[PP]) {
    uint32_t c0[PP], c1[PP], c2[PP], c3[PP];
    raw(c0, c1, c2, c3);
    box_muller(c0, c1, c2, c3, z0, z1);
  }
};
#endif

}  // namespace mcre
