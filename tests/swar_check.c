/* Host check of genometools_smax_b200/csrc/smax_swar.h against a scalar
   restatement (compiled and run by tests/test_swar_host.py). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "smax_swar.h"

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static uint32_t rnd(void)
{
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return (uint32_t) (rng_state >> 16);
}

static int bit_of(const uint32_t m[4], int i) { return (m[i >> 2] >> (8 * (i & 3) + 7)) & 1; }

static int check_primitives(void)
{
  int trial, j;
  for (trial = 0; trial < 2000000; trial++)
  {
    uint32_t x = rnd(), y = rnd(), mb = 1 + rnd() % 255, kadd;
    int himode;
    if (trial & 1) y = (y & 0xffff0000u) | (x & 0x0000ffffu);   /* force equal bytes */
    if ((trial & 6) == 2) x |= 0xfe00ff00u;
    smax_ge_consts(mb, &kadd, &himode);
    for (j = 0; j < 4; j++)
    {
      const uint32_t a = (x >> (8 * j)) & 255, b = (y >> (8 * j)) & 255, bit = 0x80u << (8 * j);
      if (!!(smax_ge(x, kadd, himode) & bit) != (a >= mb)) return 1;
      if (!!(smax_gt(x, y) & bit) != (a > b)) return 2;
      if (!!(smax_zero(x ^ y) & bit) != (a == b)) return 3;
      if (!!(smax_nonzero(x ^ y) & bit) != (a != b)) return 4;
      if (!!(smax_is255(x) & bit) != (a == 255)) return 5;
      if (!!(smax_special(x) & bit) != (a >= 254)) return 6;
    }
    if ((smax_ge(x, kadd, himode) | smax_gt(x, y) | smax_zero(x) | smax_nonzero(x) | smax_is255(x) |
         smax_special(x)) & ~SMAX_H7)
      return 7;
  }
  return 0;
}

/* scalar: classify end i of the chunk; returns width 2..4, 5 = LONG, 0 = none */
static int scalar_k1(const uint8_t *L /* L[-4..19] at L+4 */, int i, unsigned mb)
{
  const uint8_t *p = L + 4;
  unsigned v = p[i];
  if (!(v >= mb && v != 255 && v > p[i + 1])) return 0;
  if (p[i - 1] < v) return 2;
  if (p[i - 1] > v) return 0;
  if (p[i - 2] < v) return 3;
  if (p[i - 2] > v) return 0;
  if (p[i - 3] < v) return 4;
  if (p[i - 3] > v) return 0;
  return 5;
}

static int scalar_k2(const uint8_t *B /* B[-4..15] at B+4 */, int i, int width, int gt_policy)
{
  const uint8_t *p = B + 4;
  int a, b;
  for (a = i - width + 1; a <= i; a++)
    for (b = a + 1; b <= i; b++)
      if (p[a] == p[b] && !(gt_policy && p[a] >= 254))
        return 0;
  return 1;
}

static int check_chunks(void)
{
  int trial, i;
  for (trial = 0; trial < 3000000; trial++)
  {
    uint8_t L[24], B[20];
    uint32_t w[6], b[5], kadd;
    int himode, gt_policy = trial & 1, mode = (trial >> 1) % 5;
    unsigned mb = (mode == 4) ? 1 + rnd() % 255 : 1 + rnd() % 6;
    smax_chunk_k1 o;
    for (i = 0; i < 24; i++)
    {
      const uint32_t r = rnd();
      switch (mode)
      {
        case 0: L[i] = r % 4; break;                         /* many ties            */
        case 1: L[i] = r % 12; break;
        case 2: L[i] = (r % 8 == 0) ? 255 : 250 + r % 5; break;   /* around the overflow mark */
        case 3: L[i] = (i > 0 && r % 3) ? L[i - 1] : r % 7; break; /* long runs       */
        default: L[i] = r & 255; break;
      }
    }
    for (i = 0; i < 20; i++)
    {
      const uint32_t r = rnd();
      B[i] = (r % 11 == 0) ? 254 + (r >> 8) % 2 : (mode == 4 ? (r >> 4) % 20 : (r >> 4) % 4);
    }
    memcpy(w, L, 24);
    memcpy(b, B, 20);
    smax_ge_consts(mb, &kadd, &himode);
    memset(&o, 0, sizeof o);
    {
      const int found = smax_chunk_detect(w, kadd, himode, &o);
      if (!found) memset(&o, 0, sizeof o);
      for (i = 0; i < 16; i++)
      {
        const int want = scalar_k1(L, i, mb);
        const int got = bit_of(o.c2, i) ? 2 : bit_of(o.c3, i) ? 3 : bit_of(o.c4, i) ? 4 :
                        bit_of(o.lng, i) ? 5 : 0;
        if (want != got || bit_of(o.c2, i) + bit_of(o.c3, i) + bit_of(o.c4, i) + bit_of(o.lng, i) > 1)
        {
          fprintf(stderr, "K1 mismatch trial %d i %d want %d got %d\n", trial, i, want, got);
          return 1;
        }
      }
      if (found && (o.any_cand != 0) !=
          ((o.c2[0] | o.c2[1] | o.c2[2] | o.c2[3] | o.c3[0] | o.c3[1] | o.c3[2] | o.c3[3] |
            o.c4[0] | o.c4[1] | o.c4[2] | o.c4[3]) != 0))
        return 2;
      if (found && o.any_cand)
      {
        smax_chunk_k1 s = o;
        const uint32_t any = smax_chunk_distinct(b, gt_policy, &s);
        uint32_t un = 0;
        for (i = 0; i < 16; i++)
        {
          const int width = bit_of(o.c2, i) ? 2 : bit_of(o.c3, i) ? 3 : bit_of(o.c4, i) ? 4 : 0;
          const int got = bit_of(s.c2, i) | bit_of(s.c3, i) | bit_of(s.c4, i);
          const int want = width ? scalar_k2(B, i, width, gt_policy) : 0;
          if (want != got)
          {
            fprintf(stderr, "K2 mismatch trial %d i %d width %d want %d got %d\n", trial, i, width,
                    want, got);
            return 3;
          }
        }
        for (i = 0; i < 4; i++) un |= s.c2[i] | s.c3[i] | s.c4[i];
        if (un != any) return 4;
      }
    }
  }
  return 0;
}

int main(void)
{
  int rc = check_primitives();
  if (rc) { fprintf(stderr, "primitive check failed: %d\n", rc); return 1; }
  rc = check_chunks();
  if (rc) { fprintf(stderr, "chunk check failed: %d\n", rc); return 1; }
  printf("swar ok\n");
  return 0;
}
