/*
  smax_dec.h -- decimal rendering helpers shared by the device formatter
  (smax_format.cu) and its host-side check (tests/dec_check.c).

  The text conventions are the reference's: unsigned decimals as printed by
  "%lu" (GT_WU, /root/reference/src/core/types_api.h:53-54), separated by
  single blanks, one result per line
  (/root/reference/src/match/esa-lcpintervals.c:183-189).
*/
#ifndef SMAX_DEC_H
#define SMAX_DEC_H

#include <stdint.h>

#ifdef __CUDACC__
#define SMAX_DEC_FN __host__ __device__ __forceinline__
#else
#define SMAX_DEC_FN static inline
#endif

/* 10^k for k = 0..19 without a table (constant folded; usable on both sides) */
SMAX_DEC_FN uint64_t smax_pow10(unsigned k)
{
  uint64_t p = 1;
  /* two-level multiply keeps the dependent chain short */
  if (k & 16) p *= 10000000000000000ull;
  if (k & 8) p *= 100000000ull;
  if (k & 4) p *= 10000ull;
  if (k & 2) p *= 100ull;
  if (k & 1) p *= 10ull;
  return p;
}

SMAX_DEC_FN unsigned smax_bitlen64(uint64_t v)
{
#ifdef __CUDA_ARCH__
  return 64u - (unsigned) __clzll((long long) v);
#else
  return v == 0 ? 0u : 64u - (unsigned) __builtin_clzll(v);
#endif
}

/* number of decimal digits of v ("0" has one) */
SMAX_DEC_FN unsigned smax_dec_digits(uint64_t v)
{
  /* t = floor(log10(2^(bits-1))) is the digit count minus one or two */
  const unsigned bits = smax_bitlen64(v | 1);
  const unsigned t = (bits * 1233u) >> 12;
  return t + (v >= smax_pow10(t) ? 1u : 0u) + (t == 0 && v == 0 ? 1u : 0u);
}

/* writes the decimal digits of v into dst[0..digits) (no terminator) */
SMAX_DEC_FN void smax_dec_write(char *dst, uint64_t v, unsigned digits)
{
  while (digits > 0)
  {
    const uint64_t q = v / 10;
    dst[--digits] = (char) ('0' + (unsigned) (v - q * 10));
    v = q;
  }
}

#endif
