/*
  smax_swar.h -- the bit-parallel core of K1 (plateau detection) and K2
  (left-distinctness) for ONE 16-byte chunk of the lcp / bwt tables, written
  as host+device inline code so that tests/test_swar_host.py can check it on
  the CPU against a scalar restatement, byte for byte, without a GPU.

  Notation: a "mask" is a 32-bit word whose bit 7 of byte j says something
  about byte j of the operand word (bits 0..6 of every byte are zero).

  For chunk bytes L[0..15] (lcp) with L[-4..-1] (prev word) and L[16..19]
  (next word), and B[-4..15] (bwt):

    END(i)   = L[i] >= mb  and  L[i] != 255  and  L[i] > L[i+1]
    width 2  = END(i) and L[i-1] <  L[i]
    width 3  = END(i) and L[i-1] == L[i] and L[i-2] <  L[i]
    width 4  = END(i) and L[i-1] == L[i-2] == L[i] and L[i-3] < L[i]
    LONG     = END(i) and L[i-1] == L[i-2] == L[i-3] == L[i]   (walked elsewhere)

  which are exactly the local-maximum plateaus [i - width + 2, i] of SA width
  2..4 of SURVEY.md section E (an lcp-interval without a child interval,
  /root/reference/src/match/esa-bottomup.c:160-199).  255 bytes are overflow
  markers (/root/reference/src/match/lcpoverflow.h:24): they compare larger
  than every small value, which is what the true value (>= 255) would do, and
  they never END a small plateau; large values are handled in .llv space.

  K2 for width w: the left characters B[i-w+1..i] must be pairwise distinct,
  where under the GenomeTools policy a special (>= 254) never collides
  (/root/reference/src/match/esa-maxpairs.c:24-31).
*/
#ifndef SMAX_SWAR_H
#define SMAX_SWAR_H

#include <stdint.h>

#if defined(__CUDACC__)
#define SMAX_HD __host__ __device__ __forceinline__
#else
#define SMAX_HD static inline
#endif

#define SMAX_H7 0x80808080u
#define SMAX_L7 0x7f7f7f7fu

/* bytes (hi:lo) >> 8*k, k = 1: word of the NEXT bytes (i+1) */
SMAX_HD uint32_t smax_shr_bytes(uint32_t lo, uint32_t hi, int k)
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, 8 * k);
#else
  return (uint32_t) ((((uint64_t) hi << 32) | lo) >> (8 * k));
#endif
}

/* word of the bytes k positions EARLIER (i-k): prev = word before cur */
SMAX_HD uint32_t smax_shl_bytes(uint32_t prev, uint32_t cur, int k)
{
#if defined(__CUDA_ARCH__)
  return __funnelshift_l(prev, cur, 8 * k);
#else
  return (uint32_t) ((((uint64_t) cur << 32) | prev) >> (32 - 8 * k));
#endif
}

/* mask: byte >= mb, with kadd/himode precomputed by smax_ge_consts */
SMAX_HD void smax_ge_consts(uint32_t mb, uint32_t *kadd, int *himode)
{
  *himode = mb > 128;
  *kadd = ((*himode) ? (0x100u - mb) : (0x80u - mb)) * 0x01010101u;
}

SMAX_HD uint32_t smax_ge(uint32_t w, uint32_t kadd, int himode)
{
  const uint32_t t = (w & SMAX_L7) + kadd;
  return (himode ? (t & w) : (t | w)) & SMAX_H7;
}

/* mask: byte of x > byte of y (unsigned) */
SMAX_HD uint32_t smax_gt(uint32_t x, uint32_t y)
{
  const uint32_t t = (y | SMAX_H7) - (x & SMAX_L7);          /* bit7: low7(y) >= low7(x) */
  const uint32_t ge_yx = (y & ~x) | (~(y ^ x) & t);          /* bit7: y >= x */
  return ~ge_yx & SMAX_H7;
}

/* mask: byte of x == 0 */
SMAX_HD uint32_t smax_zero(uint32_t x)
{
  return ~(((x & SMAX_L7) + SMAX_L7) | x) & SMAX_H7;
}

/* mask: byte of x != 0 */
SMAX_HD uint32_t smax_nonzero(uint32_t x)
{
  return (((x & SMAX_L7) + SMAX_L7) | x) & SMAX_H7;
}

/* mask: byte == 255 */
SMAX_HD uint32_t smax_is255(uint32_t w)
{
  return ((w & SMAX_L7) + 0x01010101u) & w & SMAX_H7;
}

/* mask: byte >= 254 (special left character) */
SMAX_HD uint32_t smax_special(uint32_t b)
{
  return ((b & SMAX_L7) + 0x02020202u) & b & SMAX_H7;
}

/* mask: the pair (a, b) of left characters does NOT collide */
SMAX_HD uint32_t smax_pair_ok(uint32_t a, uint32_t b, uint32_t special_a)
{
  return smax_nonzero(a ^ b) | special_a;
}

typedef struct
{
  uint32_t c2[4], c3[4], c4[4];   /* candidate plateau ENDS by SA width (K1)      */
  uint32_t lng[4];                /* ends of runs of >= 4 equal values: walk them */
  uint32_t any_cand, any_long;
} smax_chunk_k1;

/* K1 of one chunk.  w[0] = bytes -4..-1, w[1..4] = the 16 chunk bytes,
   w[5] = bytes 16..19.  Returns 0 when the chunk holds no plateau end.
   Branches only between stages, never per word: the four words of a stage
   are independent instruction streams. */
SMAX_HD int smax_chunk_detect(const uint32_t w[6], uint32_t kadd, int himode, smax_chunk_k1 *o)
{
  uint32_t end[4], e1[4];
  uint32_t any = 0, anyc = 0, anye1 = 0, anyl = 0;
  int k;
  for (k = 0; k < 4; k++)
  {
    end[k] = smax_ge(w[k + 1], kadd, himode);
    any |= end[k];
  }
  if (any == 0)
    return 0;
  any = 0;
  for (k = 0; k < 4; k++)
  {
    end[k] &= smax_gt(w[k + 1], smax_shr_bytes(w[k + 1], w[k + 2], 1)) & ~smax_is255(w[k + 1]);
    any |= end[k];
  }
  if (any == 0)
    return 0;
  for (k = 0; k < 4; k++)
  {
    const uint32_t p1 = smax_shl_bytes(w[k], w[k + 1], 1);
    o->c2[k] = end[k] & smax_gt(w[k + 1], p1);
    e1[k] = end[k] & smax_zero(w[k + 1] ^ p1);
    o->c3[k] = o->c4[k] = o->lng[k] = 0;
    anye1 |= e1[k];
    anyc |= o->c2[k];
  }
  if (anye1)
  {
    for (k = 0; k < 4; k++)
    {
      const uint32_t p1 = smax_shl_bytes(w[k], w[k + 1], 1);
      const uint32_t p2 = smax_shl_bytes(w[k], w[k + 1], 2);
      const uint32_t p3 = smax_shl_bytes(w[k], w[k + 1], 3);
      const uint32_t e2 = e1[k] & smax_zero(p1 ^ p2);
      o->c3[k] = e1[k] & smax_gt(p1, p2);
      o->c4[k] = e2 & smax_gt(p2, p3);
      o->lng[k] = e2 & smax_zero(p2 ^ p3);
      anyc |= o->c3[k] | o->c4[k];
      anyl |= o->lng[k];
    }
  }
  o->any_cand = anyc;
  o->any_long = anyl;
  return (anyc | anyl) != 0;
}

/* K2 of one chunk: b[0] = bwt bytes -4..-1, b[1..4] = the 16 chunk bytes.
   Narrows c2/c3/c4 to the survivors; returns their union.
   gt_policy != 0: specials (>= 254) never collide. */
SMAX_HD uint32_t smax_chunk_distinct(const uint32_t b[5], int gt_policy, smax_chunk_k1 *o)
{
  uint32_t any = 0, wide = 0, wide4 = 0;
  uint32_t ok01[4];
  int k;
  for (k = 0; k < 4; k++)
  {
    const uint32_t b0 = b[k + 1];
    const uint32_t q1 = smax_shl_bytes(b[k], b0, 1);
    ok01[k] = smax_pair_ok(b0, q1, gt_policy ? smax_special(b0) : 0u);
    o->c2[k] &= ok01[k];
    any |= o->c2[k];
    wide |= o->c3[k] | o->c4[k];
    wide4 |= o->c4[k];
  }
  if (wide)
  {
    for (k = 0; k < 4; k++)
    {
      const uint32_t b0 = b[k + 1];
      const uint32_t q1 = smax_shl_bytes(b[k], b0, 1);
      const uint32_t q2 = smax_shl_bytes(b[k], b0, 2);
      const uint32_t s0 = gt_policy ? smax_special(b0) : 0u;
      const uint32_t s1 = gt_policy ? smax_special(q1) : 0u;
      const uint32_t ok3 = ok01[k] & smax_pair_ok(q1, q2, s1) & smax_pair_ok(b0, q2, s0);
      o->c3[k] &= ok3;
      if (wide4)
      {
        const uint32_t q3 = smax_shl_bytes(b[k], b0, 3);
        const uint32_t s2 = gt_policy ? smax_special(q2) : 0u;
        o->c4[k] &= ok3 & smax_pair_ok(q2, q3, s2) & smax_pair_ok(q1, q3, s1) &
                    smax_pair_ok(b0, q3, s0);
      }
      any |= o->c3[k] | o->c4[k];
    }
  }
  return any;
}

/* The two-stage form the scan kernel uses: smax_chunk_ends yields the ends of SA width 2
   (c2) and the ends of longer runs (e1: END(i) and L[i-1] == L[i]; those are walked entry
   by entry together with their left characters); smax_chunk_distinct2 is K2 for width 2. */
SMAX_HD int smax_chunk_ends(const uint32_t w[6], uint32_t kadd, int himode, uint32_t c2[4],
                            uint32_t e1[4])
{
  uint32_t end[4];
  uint32_t any = 0;
  int k;
  for (k = 0; k < 4; k++)
  {
    end[k] = smax_ge(w[k + 1], kadd, himode);
    any |= end[k];
  }
  if (any == 0)
    return 0;
  any = 0;
  for (k = 0; k < 4; k++)
  {
    end[k] &= smax_gt(w[k + 1], smax_shr_bytes(w[k + 1], w[k + 2], 1)) & ~smax_is255(w[k + 1]);
    any |= end[k];
  }
  if (any == 0)
    return 0;
  for (k = 0; k < 4; k++)
  {
    const uint32_t p1 = smax_shl_bytes(w[k], w[k + 1], 1);
    c2[k] = end[k] & smax_gt(w[k + 1], p1);
    e1[k] = end[k] & smax_zero(w[k + 1] ^ p1);
  }
  return 1;
}

SMAX_HD uint32_t smax_chunk_distinct2(const uint32_t b[5], int gt_policy, uint32_t c2[4])
{
  uint32_t any = 0;
  int k;
  for (k = 0; k < 4; k++)
  {
    const uint32_t b0 = b[k + 1];
    const uint32_t q1 = smax_shl_bytes(b[k], b0, 1);
    c2[k] &= smax_pair_ok(b0, q1, gt_policy ? smax_special(b0) : 0u);
    any |= c2[k];
  }
  return any;
}

#endif /* SMAX_SWAR_H */
