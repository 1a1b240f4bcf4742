// Counter-based Philox4x32-10 (Salmon et al., SC'11) and the uniform/normal
// helpers every sampler in this library draws from.
//
// Replaces, for the CUDA path, the reference's *unseeded global* numpy
// RandomState (e_step_gibbs.py:137,144,216,340,376) and the per-thread GSL
// Mersenne twisters inside pypolyagamma (e_step_gibbs.py:200-209).  A stream is
// addressed by (seed, em_iter) as the key and (gene, cell, sweep|purpose, block)
// as the counter, with cell/sample indices GLOBAL -- so a draw does not depend
// on how cells are sharded over GPUs (check #4).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define URSM_HD __host__ __device__ __forceinline__
#else
#define URSM_HD inline
#endif

namespace ursm {

// uniform on (0,1) from the top 23 bits of a word (see Philox::uniform)
URSM_HD float u01(uint32_t w) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(0x3f800000u | (w >> 9)) - 0.99999994f;
#else
  union { uint32_t i; float f; } c;
  c.i = 0x3f800000u | (w >> 9);
  return c.f - 0.99999994f;
#endif
}

struct Philox {
  uint32_t c0, c1, c2, c3;  // counter; c3 is the running block index
  uint32_t k0, k1;          // key
  uint32_t b0, b1, b2, b3;  // buffered outputs (shift register, no local memory)
  int n;                    // outputs left in the buffer

  URSM_HD Philox() {}
  URSM_HD Philox(uint64_t key, uint32_t x, uint32_t y, uint32_t z) {
    k0 = (uint32_t)key;
    k1 = (uint32_t)(key >> 32);
    c0 = x;
    c1 = y;
    c2 = z;
    c3 = 0;
    n = 0;
  }

  static URSM_HD void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
    lo = a * b;
    hi = __umulhi(a, b);
#else
    uint64_t p = (uint64_t)a * b;
    lo = (uint32_t)p;
    hi = (uint32_t)(p >> 32);
#endif
  }

  URSM_HD void refill() {
    uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3, ka = k0, kb = k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      uint32_t h0, l0, h1, l1;
      mulhilo(0xD2511F53u, x0, h0, l0);
      mulhilo(0xCD9E8D57u, x2, h1, l1);
      x0 = h1 ^ x1 ^ ka;
      x1 = l1;
      x2 = h0 ^ x3 ^ kb;
      x3 = l0;
      ka += 0x9E3779B9u;
      kb += 0xBB67AE85u;
    }
    b0 = x0;
    b1 = x1;
    b2 = x2;
    b3 = x3;
    n = 4;
    ++c3;
  }

  URSM_HD uint32_t next() {
    if (n == 0) refill();
    uint32_t r = b0;
    b0 = b1;
    b1 = b2;
    b2 = b3;
    --n;
    return r;
  }

  // uniform on the open interval (0,1): 23 random mantissa bits + 2^-24, i.e. the
  // centres k 2^-23 + 2^-24 of 2^23 equal cells -- never 0 or 1, no int->float convert
  URSM_HD float uniform() { return u01(next()); }
  // 53-bit-ish double uniform in (0,1) from two words
  URSM_HD double uniform_d() {
    uint32_t a = next(), b = next();
    return (((uint64_t)(a >> 5) << 26) + (b >> 6) + 0.5) * (1.0 / 9007199254740992.0);
  }
};

// One Philox4x32-10 block as a pure function: counter (c0,c1,c2,c3) -> 4 words.
// Bit-identical to the first refill of Philox(key, c0, c1, c2) when c3 == 0 and to
// its (c3+1)-th refill in general.  The Gibbs kernel calls this once per sampling
// ATTEMPT, with c3 = attempt index, so that every lane of a warp executes the
// generator at full width (no data-dependent refill).
//
// The round keys are an argument: kernels take a PhiloxKeys by value inside their
// parameter struct, so every round key is a constant-bank operand of the round's
// three-input XOR (LOP3 ..., c[0x0][..]) and a round is exactly two IMAD.WIDE.U32 and
// two LOP3 -- no key-schedule arithmetic in the loop.
struct PhiloxKeys {
  uint32_t ka[10], kb[10];
};

inline PhiloxKeys philox_keys(uint64_t key) {
  PhiloxKeys k;
  uint32_t a = (uint32_t)key, b = (uint32_t)(key >> 32);
  for (int r = 0; r < 10; ++r) {
    k.ka[r] = a;
    k.kb[r] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  return k;
}

URSM_HD void mul_wide(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#if defined(__CUDA_ARCH__)
  asm("{\n\t.reg .u64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}"
      : "=r"(lo), "=r"(hi)
      : "r"(a), "r"(b));
#else
  const uint64_t p = (uint64_t)a * b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

URSM_HD void philox_block(const PhiloxKeys& k, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t h0, l0, h1, l1;
    mul_wide(0xD2511F53u, c0, h0, l0);
    mul_wide(0xCD9E8D57u, c2, h1, l1);
    c0 = h1 ^ c1 ^ k.ka[r];
    c2 = h0 ^ c3 ^ k.kb[r];
    c1 = l1;
    c3 = l0;
  }
  r0 = c0;
  r1 = c1;
  r2 = c2;
  r3 = c3;
}

URSM_HD void philox_block(uint64_t key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                          uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  PhiloxKeys k;
  uint32_t a = (uint32_t)key, b = (uint32_t)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    k.ka[r] = a;
    k.kb[r] = b;
    a += 0x9E3779B9u;
    b += 0xBB67AE85u;
  }
  philox_block(k, c0, c1, c2, c3, r0, r1, r2, r3);
}

// Stream "purposes" folded into the third counter word (top 4 bits) so the
// different consumers of one (cell, sweep) never share a counter.
enum : uint32_t {
  kPurposeGene = 0u << 28,      // per (cell, gene, sweep): S uniform first, then PG draw
  kPurposeCell = 1u << 28,      // per (cell, sweep): the (kappa, tau) normal pair
  kPurposeInit = 2u << 28,      // per (cell, gene): initial S ~ Bern(1/2)
  kPurposeBulkZ = 3u << 28,     // per (sample, gene, sweep): multinomial split
  kPurposeBulkW = 4u << 28,     // per (sample, sweep): Dirichlet
};

}  // namespace ursm
