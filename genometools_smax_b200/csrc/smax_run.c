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
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "smax_host.h"

#define SMAX_MAX_GPUS 16
#define SMAX_MAX_SHARDS 256
#define SMAX_MAX_LEFT 8          /* peer shards a plateau may walk into (kMaxLeft) */
#define SMAX_PREFETCH_AHEAD 24   /* records the host gather of positions runs ahead with its prefetches */

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

/* ---- device handles are kept between calls -------------------------------------
   Creating a device handle (context, streams, events) and allocating its tables costs
   far more than a scan; a process that calls smax_run repeatedly (the tool once, a
   server or the benchmark many times) gets the handles of its previous call back --
   their allocations are reused by smax_device_upload when they are large enough.
   Slot = shard number of the call; smax_release_devices() (also at exit) frees them. */
static smax_device *cache_dev[SMAX_MAX_SHARDS];
static int cache_ordinal[SMAX_MAX_SHARDS];
static pthread_mutex_t cache_lock = PTHREAD_MUTEX_INITIALIZER;
static int cache_busy, cache_atexit;

void smax_release_devices(void)
{
  int g;
  pthread_mutex_lock(&cache_lock);
  if (!cache_busy)
    for (g = 0; g < SMAX_MAX_SHARDS; g++)
    {
      smax_device_destroy(cache_dev[g]);
      cache_dev[g] = NULL;
    }
  pthread_mutex_unlock(&cache_lock);
}

/* the handle for shard g on CUDA device `ordinal`: the cached one or a new one.  Only one
   run at a time uses the cache; a concurrent second run creates (and destroys) its own. */
static int acquire_device(int g, int ordinal, int cached, smax_device **out, char *err, size_t errlen)
{
  if (cached && cache_dev[g] != NULL && cache_ordinal[g] == ordinal)
  {
    *out = cache_dev[g];
    cache_dev[g] = NULL;
    return 0;
  }
  if (cached && cache_dev[g] != NULL)
  {
    smax_device_destroy(cache_dev[g]);
    cache_dev[g] = NULL;
  }
  return smax_device_create(ordinal, out, err, errlen);
}

static int begin_run(void)
{
  int cached;
  pthread_mutex_lock(&cache_lock);
  cached = !cache_busy;
  if (cached)
  {
    cache_busy = 1;
    if (!cache_atexit)
    {
      cache_atexit = 1;
      atexit(smax_release_devices);
    }
  }
  pthread_mutex_unlock(&cache_lock);
  return cached;
}

/* every device is through with its work before any table is given up: a right neighbour
   may still be reading a left neighbour's tables through its peer views */
static void end_run(smax_device **dev, const int *ordinal, int nshards, int cached, int failed)
{
  int g;
  for (g = 0; g < nshards; g++)
    if (dev[g] != NULL)
      smax_device_synchronize(dev[g]);
  for (g = 0; g < nshards; g++)
  {
    if (dev[g] == NULL)
      continue;
    if (cached && !failed)
    {
      cache_dev[g] = dev[g];
      cache_ordinal[g] = ordinal[g];
    } else
      smax_device_destroy(dev[g]);
    dev[g] = NULL;
  }
  if (cached)
  {
    pthread_mutex_lock(&cache_lock);
    cache_busy = 0;
    pthread_mutex_unlock(&cache_lock);
  }
}

/* Cuts of [0, n) into nshards contiguous ranges (multiples of 16) of equal COST rather than
   equal length: a shard's scan reads one byte per entry and 8 + 8 bytes per large value
   (compact record + its share of the directory), so cost(x) = x + 16 * #{.llv records
   before x}; with equal-length cuts the shard that holds the repeat-rich part of the
   suffix array finishes last (round 1: 12 % at 8 GPUs). */
static void balanced_cuts(const smax_index *idx, int nshards, uint64_t *cut)
{
  const uint64_t n = idx->info.numberofallsortedsuffixes, base = idx->base;
  const smax_llv *llv = idx->llv;
  const uint64_t L = llv != NULL ? idx->info.largelcpvalues : 0;
  const uint64_t total = n + 16 * L, limit = max_shard_len();
  int g;
  cut[0] = 0;
  cut[nshards] = n;
  for (g = 1; g < nshards; g++)
  {
    /* smallest x with x + 16 * rank(x) >= total * g / nshards */
    const uint64_t want = (uint64_t) ((unsigned __int128) total * (unsigned) g / (unsigned) nshards);
    uint64_t lo = 0, hi = n;
    while (lo < hi)
    {
      const uint64_t x = lo + (hi - lo) / 2;
      uint64_t a = 0, b = L;             /* rank(x): records with position < x */
      while (a < b)
      {
        const uint64_t m = a + (b - a) / 2;
        if (llv[m].position < x) a = m + 1; else b = m;
      }
      if (x + 16 * a < want) lo = x + 1; else hi = x;
    }
    cut[g] = lo & ~(uint64_t) 15;
    if (cut[g] < cut[g - 1]) cut[g] = cut[g - 1];
  }
  (void) base;
  /* no shard longer than a device shard may be: fall back to equal lengths */
  for (g = 0; g < nshards; g++)
    if (cut[g + 1] - cut[g] > limit)
    {
      for (g = 0; g <= nshards; g++)
        cut[g] = g == nshards ? n : ((n / (uint64_t) nshards) * (uint64_t) g) & ~(uint64_t) 15;
      break;
    }
}

typedef struct
{
  smax_device *dev;
  const smax_index *idx;
  uint64_t lo, hi;
  int with_suf, rc;
  char err[512];
} UploadJob;

static void *upload_thread(void *arg)
{
  UploadJob *j = arg;
  j->rc = smax_device_upload(j->dev, j->idx, j->lo, j->hi, j->with_suf, NULL, j->err, sizeof j->err);
  return NULL;
}

/* shards of [0, n): contiguous ranges of the lcp index space, cut at multiples
   of 16; every shard made resident (with its suffix table if with_suf) on its
   device -- consecutive shards share a device when there are more shards than
   devices; the uploads of shards on different devices run concurrently, one host
   thread each --, the nearest left neighbours set as views, one scan launched per
   shard */
static int scan_all_shards(const smax_index *idx, const smax_opts *opts, int with_suf,
                           smax_device **dev, int *ordinal, int ngpus, int nshards, int cached,
                           char *err, size_t errlen)
{
  smax_shard_view views[SMAX_MAX_SHARDS];
  uint64_t cut[SMAX_MAX_SHARDS + 1];
  UploadJob *job;
  pthread_t thr[SMAX_MAX_SHARDS];
  const uint64_t minlength = opts->minlength ? opts->minlength : 1;
  const int per_device = nshards / ngpus;
  int g, r, rc = 0;
  balanced_cuts(idx, nshards, cut);
  for (g = 0; g < nshards; g++)
  {
    ordinal[g] = opts->first_device + g / per_device;
    if (acquire_device(g, ordinal[g], cached, &dev[g], err, errlen) != 0)
      return -1;
  }
  job = calloc((size_t) nshards, sizeof *job);
  if (job == NULL)
    return smax_fail(err, errlen, "out of memory");
  /* round r: the r-th shard of every device */
  for (r = 0; r < per_device && rc == 0; r++)
  {
    int started = 0;
    for (g = r; g < nshards; g += per_device)
    {
      UploadJob *j = &job[g];
      j->dev = dev[g]; j->idx = idx; j->lo = cut[g]; j->hi = cut[g + 1]; j->with_suf = with_suf;
      j->rc = 0; j->err[0] = '\0';
      if (ngpus == 1 || pthread_create(&thr[g], NULL, upload_thread, j) != 0)
      {
        upload_thread(j);
        thr[g] = 0;
      } else
        started++;
    }
    for (g = r; g < nshards; g += per_device)
    {
      if (thr[g] != 0)
        pthread_join(thr[g], NULL);
      if (job[g].rc != 0 && rc == 0)
      {
        smax_fail(err, errlen, "%s", job[g].err);
        rc = -1;
      }
    }
    (void) started;
  }
  free(job);
  if (rc != 0)
    return -1;
  for (g = 0; g < nshards; g++)
  {
    const int nleft = g < SMAX_MAX_LEFT ? g : SMAX_MAX_LEFT;
    smax_device_view(dev[g], &views[g]);
    /* the nearest neighbours, sorted by a_lo */
    if (smax_device_set_left_views(dev[g], views + (g - nleft), nleft, err, errlen) != 0)
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
  int ordinal[SMAX_MAX_SHARDS];
  uint64_t cnt[SMAX_MAX_SHARDS], total = 0, off = 0;
  smax_record *recs = NULL;
  int g, ngpus = 1, nshards = 1, rc = -1, empty = 0, cached;

  if (idx == NULL || opts == NULL || recs_out == NULL || nrecs_out == NULL)
    return smax_fail(err, errlen, "smax_run: null argument");
  *recs_out = NULL;
  *nrecs_out = 0;
  if (check_run_args(idx, opts, &ngpus, &nshards, &empty, err, errlen) != 0)
    return -1;
  if (empty)
    return 0;
  memset(dev, 0, sizeof dev);
  cached = begin_run();
  if (scan_all_shards(idx, opts, 0, dev, ordinal, ngpus, nshards, cached, err, errlen) != 0)
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
  end_run(dev, ordinal, nshards, cached, rc != 0);
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
  int ordinal[SMAX_MAX_SHARDS], cached = 0;
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
  cached = begin_run();
  if (scan_all_shards(idx, opts, with_suf, dev, ordinal, ngpus, nshards, cached, err, errlen) != 0)
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
  end_run(dev, ordinal, nshards, cached, rc != 0);
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
    if (r + SMAX_PREFETCH_AHEAD < nrecs && recs[r + SMAX_PREFETCH_AHEAD].lb >= idx->base)
      __builtin_prefetch((const char *) idx->suf +
                         (recs[r + SMAX_PREFETCH_AHEAD].lb - idx->base) * idx->info.sufbytes, 0, 0);
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
      /* the suffix table entries of a repeat are a random access into a table of 8 n bytes:
         ask for those of a later record now (a cache miss per record otherwise costs more
         than everything else the host does for it) */
      if (r + SMAX_PREFETCH_AHEAD < nrecs)
      {
        const uint64_t plb = recs[r + SMAX_PREFETCH_AHEAD].lb;
        const char *pa = (const char *) idx->suf + plb * idx->info.sufbytes;
        __builtin_prefetch(pa, 0, 0);
        __builtin_prefetch(pa + 64, 0, 0);
      }
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
