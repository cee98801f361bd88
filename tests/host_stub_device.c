/*
  host_stub_device.c -- TEST INFRASTRUCTURE, never linked into libsmax.so.

  A CPU stand-in for the device half of the C ABI (smax_device_* / smax_scan_*
  of include/smax.h) so that the HOST side of the path -- loader, option
  parser, shard / chunk drivers (smax_run.c, smax_stream.c), emitter -- can be
  exercised end to end without a GPU, under AddressSanitizer and UBSan
  (tests/test_host_asan.py).  It keeps the contract of the real device
  manager: a shard owns the plateaus that END in [lo, hi), holds the window
  [lo - 256, hi + 16) of the tables, walks left into the views of its left
  neighbours and reports a plateau that leaves the resident range as an error.

  The scan itself is the plain linear statement of the definition (SURVEY.md
  section E; interval semantics /root/reference/src/match/esa-bottomup.c:116-273,
  left characters /root/reference/src/match/esa-maxpairs.c:24-31): a checker,
  like oracle/, not a product path.
*/
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "smax_host.h"

struct smax_device
{
  uint8_t *lcp, *bwt;
  smax_llv *llv;
  uint64_t nllv, a_lo, a_hi, lo, hi, n;
  const struct smax_device *left[8];
  int nleft;
  smax_record *recs;
  uint64_t nrecs, cap;
  int error, scanned;
};

static int fail(char *err, size_t errlen, const char *msg)
{
  if (err != NULL && errlen > 0)
    snprintf(err, errlen, "%s", msg);
  return -1;
}

/* SMAX_STUB_DEVICES=N makes the stub show N devices (multi-GPU shard driver on CPU) */
int smax_device_count(char *err, size_t errlen)
{
  const char *e = getenv("SMAX_STUB_DEVICES");
  (void) err; (void) errlen;
  return e != NULL && atoi(e) > 0 ? atoi(e) : 1;
}

int smax_device_create(int ordinal, smax_device **out, char *err, size_t errlen)
{
  if (ordinal < 0 || ordinal >= smax_device_count(NULL, 0))
    return fail(err, errlen, "CUDA device does not exist (stub)");
  *out = calloc(1, sizeof **out);
  return *out != NULL ? 0 : fail(err, errlen, "out of memory");
}

static void drop_tables(smax_device *d)
{
  free(d->lcp); free(d->bwt); free(d->llv);
  d->lcp = d->bwt = NULL;
  d->llv = NULL;
}

void *smax_device_own_stream(smax_device *d)
{
  (void) d;
  return NULL;
}

int smax_device_synchronize(smax_device *d)
{
  (void) d;
  return 0;
}

void smax_device_destroy(smax_device *d)
{
  if (d == NULL)
    return;
  drop_tables(d);
  free(d->recs);
  free(d);
}

int smax_device_upload(smax_device *d, const smax_index *idx, uint64_t lo, uint64_t hi,
                       int with_suf, uint64_t *h2d_bytes, char *err, size_t errlen)
{
  return smax_device_upload_halo(d, idx, lo, hi, 256, with_suf, h2d_bytes, err, errlen);
}

int smax_device_upload_halo(smax_device *d, const smax_index *idx, uint64_t lo, uint64_t hi,
                            uint64_t halo, int with_suf, uint64_t *h2d_bytes, char *err,
                            size_t errlen)
{
  const uint64_t n = idx->info.numberofallsortedsuffixes, base = idx->base,
                 wend = idx->base + idx->len;
  uint64_t a_lo, a_hi, k, L = idx->info.largelcpvalues, k0 = 0, k1;
  (void) with_suf; (void) h2d_bytes;
  if (hi > n) hi = n;
  if (lo > hi || (lo & 15) != 0)
    return fail(err, errlen, "shard range must start at a multiple of 16");
  halo = halo < 256 ? 256 : (halo + 15) & ~(uint64_t) 15;
  a_lo = lo >= halo ? lo - halo : 0;
  if (a_lo < base) a_lo = (base + 15) & ~(uint64_t) 15;
  a_hi = hi + 16 < n ? hi + 16 : n;
  if (a_hi > wend) a_hi = wend;
  if (a_lo > lo || a_hi < hi || (hi < n && a_hi < hi + 1))
    return fail(err, errlen, "host tables do not cover the shard plus one entry");
  drop_tables(d);
  d->lcp = malloc(a_hi - a_lo + 1);
  d->bwt = malloc(a_hi - a_lo + 1);
  if (d->lcp == NULL || d->bwt == NULL)
    return fail(err, errlen, "out of memory");
  memcpy(d->lcp, idx->lcp + (a_lo - base), a_hi - a_lo);
  memcpy(d->bwt, idx->bwt + (a_lo - base), a_hi - a_lo);
  while (k0 < L && idx->llv[k0].position < a_lo) k0++;
  k1 = k0;
  while (k1 < L && idx->llv[k1].position < a_hi) k1++;
  d->nllv = k1 - k0;
  d->llv = malloc((d->nllv + 1) * sizeof *d->llv);
  if (d->llv == NULL)
    return fail(err, errlen, "out of memory");
  for (k = 0; k < d->nllv; k++)
    d->llv[k] = idx->llv[k0 + k];
  d->a_lo = a_lo; d->a_hi = a_hi; d->lo = lo; d->hi = hi; d->n = n;
  d->nleft = 0;
  d->scanned = 0;
  return 0;
}

int smax_device_view(const smax_device *d, smax_shard_view *v)
{
  memset(v, 0, sizeof *v);
  v->a_lo = d->a_lo; v->a_hi = d->a_hi;
  v->d_lcp = (uint64_t) (uintptr_t) d;       /* the stub's "device address" is the handle */
  v->nllv = d->nllv;
  return 0;
}

int smax_device_set_left_views(smax_device *d, const smax_shard_view *views, int nviews,
                               char *err, size_t errlen)
{
  int k;
  if (nviews < 0 || nviews > 8)
    return fail(err, errlen, "at most 8 left neighbours are supported");
  for (k = 0; k < nviews; k++)
  {
    if (views[k].a_lo > d->a_lo)
      return fail(err, errlen, "left view does not lie left of the shard");
    if (k > 0 && views[k - 1].a_hi > views[k].a_hi)
      return fail(err, errlen, "left views must be in shard order (nearest neighbour last)");
    d->left[k] = (const smax_device *) (uintptr_t) views[k].d_lcp;
  }
  d->nleft = nviews;
  return 0;
}

/* table that holds lcp index i: the shard itself, else the nearest left view */
static const smax_device *holder(const smax_device *d, uint64_t i)
{
  int k;
  if (i >= d->a_lo && i < d->a_hi)
    return d;
  for (k = d->nleft - 1; k >= 0; k--)
    if (i >= d->left[k]->a_lo && i < d->left[k]->a_hi)
      return d->left[k];
  return NULL;
}

static int value_at(smax_device *d, uint64_t i, uint64_t *v)
{
  const smax_device *t = holder(d, i);
  uint64_t lo = 0, hi;
  if (t == NULL) { d->error = 6; return -1; }
  *v = t->lcp[i - t->a_lo];
  if (*v != 255)
    return 0;
  hi = t->nllv;
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) / 2;
    if (t->llv[mid].position < i) lo = mid + 1; else hi = mid;
  }
  if (lo >= t->nllv || t->llv[lo].position != i) { d->error = 1; return -1; }
  *v = t->llv[lo].value;
  return 0;
}

static int left_at(smax_device *d, uint64_t i, unsigned *c)
{
  const smax_device *t = holder(d, i);
  if (t == NULL) { d->error = 6; return -1; }
  *c = t->bwt[i - t->a_lo];
  return 0;
}

int smax_scan_launch(smax_device *d, uint64_t minlength, int policy, int gather, void *stream,
                     char *err, size_t errlen)
{
  uint64_t e;
  (void) stream;
  if (d->lcp == NULL)
    return fail(err, errlen, "no tables resident");
  if (gather)
    return fail(err, errlen, "the stub has no suffix table");
  d->nrecs = 0;
  d->error = 0;
  d->scanned = 1;
  for (e = d->lo; e < d->hi && !d->error; e++)
  {
    uint64_t v, next = 0, s, pv = 0, i;
    int distinct = 1;
    if (value_at(d, e, &v) != 0) break;
    if (v < minlength || v == 0) continue;
    if (e + 1 < d->n && value_at(d, e + 1, &next) != 0) break;
    if (next >= v) continue;                    /* the run goes on, or rises */
    for (s = e; s > 0; s--)
    {
      if (value_at(d, s - 1, &pv) != 0) break;
      if (pv != v) break;
    }
    if (d->error) break;
    if (s == 0 || pv > v) continue;             /* entered from a larger value: no local maximum */
    /* SA interval [s - 1, e]: pairwise different left characters? */
    {
      unsigned char seen[256];
      memset(seen, 0, sizeof seen);
      for (i = s - 1; i <= e && distinct; i++)
      {
        unsigned c;
        if (left_at(d, i, &c) != 0) break;
        if (policy == SMAX_POLICY_GT && c >= 254)
          continue;                             /* specials never collide */
        if (seen[c]) distinct = 0;
        seen[c] = 1;
      }
    }
    if (d->error) break;
    if (!distinct) continue;
    if (d->nrecs == d->cap)
    {
      d->cap = d->cap ? 2 * d->cap : 1024;
      d->recs = realloc(d->recs, d->cap * sizeof *d->recs);
      if (d->recs == NULL)
        return fail(err, errlen, "out of memory");
    }
    d->recs[d->nrecs].len = v;
    d->recs[d->nrecs].lb = s - 1;
    d->recs[d->nrecs].width = e - s + 2;
    d->nrecs++;
  }
  return 0;
}

int smax_scan_counts(smax_device *d, uint64_t *nrecs, uint64_t *npositions, char *err,
                     size_t errlen)
{
  if (!d->scanned)
    return fail(err, errlen, "no scan has been launched");
  if (d->error == 6)   /* the codes and messages of the real device manager */
  {
    fail(err, errlen, "a plateau leaves the resident range of the tables (the shard's own "
                      "arrays and its left neighbour views)");
    return SMAX_E_RANGE;
  }
  if (d->error)
    return fail(err, errlen, "inconsistent ESA tables (code 1): a 255 entry of the lcp table "
                             "has no .llv record, or a repeat is wider than a shard");
  if (nrecs) *nrecs = d->nrecs;
  if (npositions) *npositions = 0;
  return 0;
}

int smax_scan_fetch(smax_device *d, smax_record *recs, uint64_t *positions, char *err,
                    size_t errlen)
{
  uint64_t n;
  (void) positions;
  if (smax_scan_counts(d, &n, NULL, err, errlen) != 0)
    return -1;
  if (recs != NULL && n > 0)
    memcpy(recs, d->recs, n * sizeof *recs);
  return 0;
}

/* the device-side emit path has no CPU stand-in */
int smax_device_set_separators(smax_device *d, const uint64_t *seps, uint64_t nseps, char *err,
                               size_t errlen)
{
  (void) d; (void) seps; (void) nseps;
  return fail(err, errlen, "stub: no device formatter");
}

int smax_scan_format(smax_device *d, int format, int relative, uint64_t *nbytes, char *err,
                     size_t errlen)
{
  (void) d; (void) format; (void) relative; (void) nbytes;
  return fail(err, errlen, "stub: no device formatter");
}

int smax_scan_fetch_text(smax_device *d, char *dst, char *err, size_t errlen)
{
  (void) d; (void) dst;
  return fail(err, errlen, "stub: no device formatter");
}
