/* Host-side check of genometools_smax_b200/csrc/smax_dec.h (the decimal
   rendering the device formatter uses) against printf("%lu"), the reference's
   convention (GT_WU, /root/reference/src/core/types_api.h:53-54). */
#include <inttypes.h>
#include <stdio.h>
#include <string.h>
#include "smax_dec.h"

static int check(uint64_t v)
{
  char want[32], got[32];
  const int n = snprintf(want, sizeof want, "%" PRIu64, v);
  const unsigned d = smax_dec_digits(v);
  if ((int) d != n)
  {
    fprintf(stderr, "digits(%" PRIu64 ") = %u, want %d\n", v, d, n);
    return 1;
  }
  memset(got, 0, sizeof got);
  smax_dec_write(got, v, d);
  if (memcmp(got, want, (size_t) n) != 0)
  {
    fprintf(stderr, "write(%" PRIu64 ") = %s\n", v, got);
    return 1;
  }
  return 0;
}

int main(void)
{
  uint64_t p = 1, x = 88172645463325252ull;
  int k, bad = 0;
  long i;
  for (k = 0; k < 20; k++)
  {
    if (smax_pow10((unsigned) k) != p)
    {
      fprintf(stderr, "pow10(%d)\n", k);
      return 1;
    }
    bad += check(p) + check(p - 1) + check(p + 1) + check(p * 9) + check(p * 2 - 1);
    if (k < 19) p *= 10;
  }
  for (k = 0; k < 64; k++)
    bad += check(1ull << k) + check((1ull << k) - 1) + check((1ull << k) + 1);
  bad += check(0) + check(~0ull) + check(~0ull - 1);
  for (i = 0; i < 200000; i++)
    bad += check((uint64_t) i);
  for (i = 0; i < 3000000; i++)
  {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;       /* xorshift64 */
    bad += check(x >> (x & 63));
  }
  if (bad)
    return 1;
  puts("dec ok");
  return 0;
}
