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

void smax_free(void *p)
{
  free(p);
}

int smax_run_records(const smax_index *idx, const smax_opts *opts, smax_record **recs_out,
                     uint64_t *nrecs_out, char *err, size_t errlen)
{
  smax_device *dev[SMAX_MAX_GPUS];
  smax_shard_view views[SMAX_MAX_GPUS];
  uint64_t cut[SMAX_MAX_GPUS + 1], cnt[SMAX_MAX_GPUS], total = 0, off = 0;
  smax_record *recs = NULL;
  int g, ngpus, rc = -1, navail;
  uint64_t n, minlength;

  if (idx == NULL || opts == NULL || recs_out == NULL || nrecs_out == NULL)
    return smax_fail(err, errlen, "smax_run: null argument");
  if (idx->lcp == NULL || idx->bwt == NULL)
    return smax_fail(err, errlen, "smax_run: the lcp and bwt tables are required");
  n = idx->info.numberofallsortedsuffixes;
  minlength = opts->minlength ? opts->minlength : 1;
  ngpus = opts->ngpus > 1 ? opts->ngpus : 1;
  if (ngpus > SMAX_MAX_GPUS)
    return smax_fail(err, errlen, "at most %d GPUs are supported", SMAX_MAX_GPUS);
  *recs_out = NULL;
  *nrecs_out = 0;
  /* maxbranchdepth is the largest lcp value (.prj, sfx-outprj.c:53-82): a
     larger minimum length has an empty answer without touching a table */
  if (idx->map_lcp != NULL && idx->info.maxbranchdepth > 0 &&
      minlength > idx->info.maxbranchdepth)
    return 0;
  navail = smax_device_count(err, errlen);
  if (navail < 0)
    return -1;
  if (opts->first_device < 0 || opts->first_device + ngpus > navail)
    return smax_fail(err, errlen, "%d GPU(s) requested starting at device %d, but only %d "
                     "visible", ngpus, opts->first_device, navail);
  memset(dev, 0, sizeof dev);
  /* contiguous ranges of the lcp index space, cut at multiples of 16 */
  for (g = 0; g <= ngpus; g++)
    cut[g] = g == ngpus ? n : ((n / (uint64_t) ngpus) * (uint64_t) g) & ~(uint64_t) 15;
  for (g = 0; g < ngpus; g++)
  {
    if (smax_device_create(opts->first_device + g, &dev[g], err, errlen) != 0)
      goto done;
    if (smax_device_upload(dev[g], idx, cut[g], cut[g + 1], 0, NULL, err, errlen) != 0)
      goto done;
    smax_device_view(dev[g], &views[g]);
    if (g > 0 && smax_device_set_left_views(dev[g], views, g < 8 ? g : 8, err, errlen) != 0)
      goto done;
  }
  for (g = 0; g < ngpus; g++)
    if (smax_scan_launch(dev[g], minlength, opts->policy, 0, NULL, err, errlen) != 0)
      goto done;
  for (g = 0; g < ngpus; g++)
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
    for (g = 0; g < ngpus; g++)
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
  for (g = 0; g < ngpus; g++)
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
