/*
  smax_run.c -- the algorithm-level entry of libsmax (host side, C): shard the
  suffix-array index range over the requested GPUs, make each shard resident,
  scan, collect the records in SA order and hand every repeat to the caller.

  Shape of gt_callenummaxpairs(indexname, minlength, scan, callback, info,
  logger, err) (/root/reference/src/match/esa-maxpairs.c:476-513): load ->
  enumerate -> callback per result -> release; errors are returned as -1 with
  a message, results are delivered in ascending left boundary, which is the
  order in which the reference's sweep pops intervals
  (/root/reference/src/match/esa-bottomup.c:160-170).

  This single-process driver addresses all GPUs of the box itself (peer
  access over NVLink for plateaus that cross a cut); the one-process-per-GPU
  variant with NCCL lives in genometools_smax_b200/shard.py on the same
  smax_device_* calls.
*/
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "smax_host.h"

#define SMAX_MAX_GPUS 16
#define SMAX_MAX_SHARDS 256
#define SMAX_MAX_LEFT 8          /* peer shards a plateau may walk into (kMaxLeft) */

void smax_free(void *p)
{
  free(p);
}

/* largest SA range one shard may own: the scan kernel keeps 32-bit tile offsets
   (smax_device_upload refuses more than SMAX_MAX_SHARD_LEN suffixes).  SMAX_MAX_SHARD is a
   test hook that lowers it, so that the several-shards-per-device path can be
   exercised on small indexes. */
static uint64_t max_shard_len(void)
{
  const char *e = getenv("SMAX_MAX_SHARD");
  uint64_t v = e != NULL ? strtoull(e, NULL, 10) : 0;
  if (v < 1024)
    v = SMAX_MAX_SHARD_LEN;
  return v & ~(uint64_t) 15;
}

/* number of shards for n suffixes on ngpus devices: one per device unless a
   shard would exceed the kernel's range, then k per device */
static int shard_count(uint64_t n, int ngpus)
{
  const uint64_t per = max_shard_len();
  uint64_t k = (n + per * (uint64_t) ngpus - 1) / (per * (uint64_t) ngpus);
  if (k < 1) k = 1;
  return k * (uint64_t) ngpus > SMAX_MAX_SHARDS ? -1 : (int) (k * (uint64_t) ngpus);
}

/* shards of [0, n): contiguous ranges of the lcp index space, cut at multiples
   of 16; every shard made resident (with its suffix table if with_suf) on its
   device -- consecutive shards share a device when there are more shards than
   devices --, the nearest left neighbours set as views, one scan launched per
   shard */
static int scan_all_shards(const smax_index *idx, const smax_opts *opts, int with_suf,
                           smax_device **dev, int ngpus, int nshards, char *err, size_t errlen)
{
  smax_shard_view views[SMAX_MAX_SHARDS];
  uint64_t cut[SMAX_MAX_SHARDS + 1];
  const uint64_t n = idx->info.numberofallsortedsuffixes;
  const uint64_t minlength = opts->minlength ? opts->minlength : 1;
  const int per_device = nshards / ngpus;
  int g;
  for (g = 0; g <= nshards; g++)
    cut[g] = g == nshards ? n : ((n / (uint64_t) nshards) * (uint64_t) g) & ~(uint64_t) 15;
  for (g = 0; g < nshards; g++)
  {
    const int nleft = g < SMAX_MAX_LEFT ? g : SMAX_MAX_LEFT;
    if (smax_device_create(opts->first_device + g / per_device, &dev[g], err, errlen) != 0)
      return -1;
    if (smax_device_upload(dev[g], idx, cut[g], cut[g + 1], with_suf, NULL, err, errlen) != 0)
      return -1;
    smax_device_view(dev[g], &views[g]);
    /* the nearest neighbours, sorted by a_lo */
    if (g > 0 && smax_device_set_left_views(dev[g], views + (g - nleft), nleft, err, errlen) != 0)
      return -1;
  }
  for (g = 0; g < nshards; g++)
    if (smax_scan_launch(dev[g], minlength, opts->policy, with_suf, NULL, err, errlen) != 0)
      return -1;
  return 0;
}

static int check_run_args(const smax_index *idx, const smax_opts *opts, int *ngpus_out,
                          int *nshards_out, int *empty, char *err, size_t errlen)
{
  int ngpus, navail;
  uint64_t minlength;
  if (idx->lcp == NULL || idx->bwt == NULL)
    return smax_fail(err, errlen, "smax_run: the lcp and bwt tables are required");
  minlength = opts->minlength ? opts->minlength : 1;
  ngpus = opts->ngpus > 1 ? opts->ngpus : 1;
  if (ngpus > SMAX_MAX_GPUS)
    return smax_fail(err, errlen, "at most %d GPUs are supported", SMAX_MAX_GPUS);
  /* maxbranchdepth is the largest lcp value (.prj, sfx-outprj.c:53-82): a
     larger minimum length has an empty answer without touching a table */
  *empty = idx->map_lcp != NULL && idx->info.maxbranchdepth > 0 &&
           minlength > idx->info.maxbranchdepth;
  *ngpus_out = ngpus;
  *nshards_out = shard_count(idx->info.numberofallsortedsuffixes, ngpus);
  if (*nshards_out < 0)
    return smax_fail(err, errlen, "%lu suffixes need more than %d shards on %d GPU(s); use more "
                     "GPUs or the -scan mode", (unsigned long) idx->info.numberofallsortedsuffixes,
                     SMAX_MAX_SHARDS, ngpus);
  if (*empty)
    return 0;
  navail = smax_device_count(err, errlen);
  if (navail < 0)
    return -1;
  if (opts->first_device < 0 || opts->first_device + ngpus > navail)
    return smax_fail(err, errlen, "%d GPU(s) requested starting at device %d, but only %d "
                     "visible", ngpus, opts->first_device, navail);
  return 0;
}

int smax_run_records(const smax_index *idx, const smax_opts *opts, smax_record **recs_out,
                     uint64_t *nrecs_out, char *err, size_t errlen)
{
  smax_device *dev[SMAX_MAX_SHARDS];
  uint64_t cnt[SMAX_MAX_SHARDS], total = 0, off = 0;
  smax_record *recs = NULL;
  int g, ngpus = 1, nshards = 1, rc = -1, empty = 0;

  if (idx == NULL || opts == NULL || recs_out == NULL || nrecs_out == NULL)
    return smax_fail(err, errlen, "smax_run: null argument");
  *recs_out = NULL;
  *nrecs_out = 0;
  if (check_run_args(idx, opts, &ngpus, &nshards, &empty, err, errlen) != 0)
    return -1;
  if (empty)
    return 0;
  memset(dev, 0, sizeof dev);
  if (scan_all_shards(idx, opts, 0, dev, ngpus, nshards, err, errlen) != 0)
    goto done;
  for (g = 0; g < nshards; g++)
  {
    if (smax_scan_counts(dev[g], &cnt[g], NULL, err, errlen) != 0)
      goto done;
    total += cnt[g];
  }
  if (total > 0)
  {
    recs = malloc(total * sizeof *recs);
    if (recs == NULL)
    {
      smax_fail(err, errlen, "out of memory for %lu records", (unsigned long) total);
      goto done;
    }
    for (g = 0; g < nshards; g++)
    {
      if (smax_scan_fetch(dev[g], recs + off, NULL, err, errlen) != 0)
        goto done;
      off += cnt[g];
    }
  }
  *recs_out = recs;
  *nrecs_out = total;
  recs = NULL;
  rc = 0;
done:
  free(recs);
  for (g = 0; g < nshards; g++)
    smax_device_destroy(dev[g]);
  return rc;
}

/* the emit path rendered on the devices: what the
   reference does with one printf per result
   (/root/reference/src/match/esa-lcpintervals.c:183-189,
   /root/reference/src/match/querymatch.c:169-187) happens in HBM; the host
   writes the bytes of every shard in shard order = ascending left boundary */
int smax_run_text(const smax_index *idx, const smax_opts *opts, void *file, uint64_t *nbytes,
                  char *err, size_t errlen)
{
  smax_device *dev[SMAX_MAX_SHARDS];
  FILE *fp = file != NULL ? (FILE *) file : stdout;
  const uint64_t *seps = NULL;
  uint64_t nseps = 0, total = 0, bytes[SMAX_MAX_SHARDS];
  char *buf = NULL;
  size_t bufcap = 0;
  int g, ngpus = 1, nshards = 1, rc = -1, empty = 0, with_suf;

  if (idx == NULL || opts == NULL)
    return smax_fail(err, errlen, "smax_run_text: null argument");
  if (nbytes != NULL) *nbytes = 0;
  if (opts->format != SMAX_FORMAT_SMAX && opts->format != SMAX_FORMAT_ITV)
    return smax_fail(err, errlen, "the device formatter renders the smax and itv formats; "
                     "format %d is rendered by the host emitter", (int) opts->format);
  with_suf = opts->format == SMAX_FORMAT_SMAX;
  if (with_suf && idx->suf == NULL)
    return smax_fail(err, errlen, "the index was opened without the suffix table");
  if (check_run_args(idx, opts, &ngpus, &nshards, &empty, err, errlen) != 0)
    return -1;
  if (empty)
    return 0;
  if (with_suf && opts->relative &&
      smax_index_separators((smax_index *) idx, &seps, &nseps, err, errlen) != 0)
    return -1;
  memset(dev, 0, sizeof dev);
  if (scan_all_shards(idx, opts, with_suf, dev, ngpus, nshards, err, errlen) != 0)
    goto done;
  for (g = 0; g < nshards; g++)
  {
    if (with_suf && opts->relative &&
        smax_device_set_separators(dev[g], seps, nseps, err, errlen) != 0)
      goto done;
    if (smax_scan_format(dev[g], opts->format, opts->relative, &bytes[g], err, errlen) != 0)
      goto done;
  }
  for (g = 0; g < nshards; g++)
  {
    if (bytes[g] == 0)
      continue;
    if (bytes[g] > bufcap)
    {
      char *p = realloc(buf, bytes[g]);
      if (p == NULL)
      {
        smax_fail(err, errlen, "out of memory for %lu bytes of text", (unsigned long) bytes[g]);
        goto done;
      }
      buf = p;
      bufcap = bytes[g];
    }
    if (smax_scan_fetch_text(dev[g], buf, err, errlen) != 0)
      goto done;
    if (fwrite(buf, 1, bytes[g], fp) != bytes[g])
    {
      smax_fail(err, errlen, "cannot write results");
      goto done;
    }
    total += bytes[g];
  }
  if (nbytes != NULL) *nbytes = total;
  rc = 0;
done:
  free(buf);
  for (g = 0; g < nshards; g++)
    smax_device_destroy(dev[g]);
  return rc;
}

int smax_index_gather_positions(const smax_index *idx, const smax_record *recs,
                                uint64_t nrecs, uint64_t *out, char *err, size_t errlen)
{
  uint64_t r, k, o = 0;
  if (idx == NULL || idx->suf == NULL)
    return smax_fail(err, errlen, "the index was opened without the suffix table");
  for (r = 0; r < nrecs; r++)
  {
    const uint64_t lb = recs[r].lb, w = recs[r].width;
    if (lb < idx->base || lb + w > idx->base + idx->len)
      return smax_fail(err, errlen, "record %lu lies outside the host suffix table",
                       (unsigned long) r);
    if (idx->info.sufbytes == 8)
      memcpy(out + o, (const uint64_t *) idx->suf + (lb - idx->base), w * sizeof *out);
    else
      for (k = 0; k < w; k++)
        out[o + k] = ((const uint32_t *) idx->suf)[lb - idx->base + k];
    o += w;
  }
  return 0;
}

int smax_run(const smax_index *idx, const smax_opts *opts, smax_emit_cb cb, void *info,
             char *err, size_t errlen)
{
  smax_record *recs = NULL;
  uint64_t nrecs = 0, r, k, poscap = 0, *pos = NULL;
  int rc = 0;

  if (cb == NULL)
    return smax_fail(err, errlen, "smax_run: null callback");
  if (smax_run_records(idx, opts, &recs, &nrecs, err, errlen) != 0)
    return -1;
  for (r = 0; r < nrecs && rc == 0; r++)
  {
    const uint64_t w = recs[r].width, lb = recs[r].lb;
    if (idx->suf != NULL)
    {
      if (w > poscap)
      {
        uint64_t *p = realloc(pos, w * sizeof *pos);
        if (p == NULL)
        {
          rc = smax_fail(err, errlen, "out of memory");
          break;
        }
        pos = p;
        poscap = w;
      }
      /* occurrence positions in suffix-array order */
      if (idx->info.sufbytes == 8)
        memcpy(pos, (const uint64_t *) idx->suf + lb, w * sizeof *pos);
      else
        for (k = 0; k < w; k++)
          pos[k] = ((const uint32_t *) idx->suf)[lb + k];
    }
    if (cb(info, recs[r].len, lb, w, idx->suf != NULL ? pos : NULL) != 0)
      rc = smax_fail(err, errlen, "result callback failed");
  }
  free(pos);
  free(recs);
  return rc;
}
