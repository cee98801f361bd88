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
#include <time.h>
#include "smax_host.h"

#define SMAX_MAX_GPUS 16
#define SMAX_MAX_SHARDS 256
#define SMAX_MAX_LEFT 8          /* peer shards a plateau may walk into (kMaxLeft) */
#define SMAX_WORKERS 4           /* threads per device that scan / render its shards (parallel pipeline) */
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
int smax_cache_acquire(int g, int ordinal, int cached, smax_device **out, char *err, size_t errlen)
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

int smax_cache_begin(void)
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
void smax_cache_end(smax_device **dev, const int *ordinal, int nshards, int cached, int failed)
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

/* ---- the shard pipeline ----------------------------------------------------------
   One uploader thread per device makes that device's shards resident one after the other; the
   calling thread launches the scan of a shard as soon as the shard and everything left of it
   is resident (the scan may walk into its left neighbours), on the shard handle's own stream,
   and then consumes its result (fetch, callbacks, text) while the next shard is uploaded:
   host work, scan and upload overlap, results are consumed in suffix-array order. */
typedef int (*ShardConsumer)(void *ctx, int g, smax_device *dev, char *err, size_t errlen);

typedef struct
{
  const smax_index *idx;
  smax_device **dev;
  const uint64_t *cut;
  int with_suf, first, step, count;      /* shards first, first + step, ... (count of them) */
  pthread_mutex_t *lock;
  pthread_cond_t *cond;
  int *uploaded;                         /* per shard: 0 pending, 1 resident, -1 failed */
  volatile int *abort;
  char err[512];
} UploadJob;

/* SMAX_TRACE: milliseconds since the first call, for the per-shard time line on stderr */
static double trace_now(void)
{
  static struct timespec t0;
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  if (t0.tv_sec == 0 && t0.tv_nsec == 0)
    t0 = t;
  return (double) (t.tv_sec - t0.tv_sec) * 1e3 + (double) (t.tv_nsec - t0.tv_nsec) * 1e-6;
}

static void *upload_thread(void *arg)
{
  UploadJob *j = arg;
  int k;
  for (k = 0; k < j->count; k++)
  {
    const int g = j->first + k * j->step;
    int rc = -1;
    if (!*j->abort)
      rc = smax_device_upload(j->dev[g], j->idx, j->cut[g], j->cut[g + 1], j->with_suf, NULL,
                              j->err, sizeof j->err);
    if (getenv("SMAX_TRACE") != NULL)
      fprintf(stderr, "# %9.3f ms  shard %d resident\n", trace_now(), g);
    pthread_mutex_lock(j->lock);
    j->uploaded[g] = rc == 0 ? 1 : -1;
    pthread_cond_broadcast(j->cond);
    pthread_mutex_unlock(j->lock);
    if (rc != 0)
      break;
  }
  return NULL;
}

/* a plateau of shard g reaches further left than its own arrays and the resident neighbour
   views (SMAX_E_RANGE): the shard alone is made resident again with a wider window to the
   left, doubled until the plateau fits or the window starts with the table */
static int redo_with_wider_halo(const smax_index *idx, const smax_opts *opts, int with_suf,
                                smax_device *dev, uint64_t lo, uint64_t hi, char *err, size_t errlen)
{
  const uint64_t minlength = opts->minlength ? opts->minlength : 1;
  uint64_t halo = hi - lo < 4096 ? 4096 : hi - lo;
  for (;;)
  {
    uint64_t nrecs;
    int rc;
    if (halo > lo) halo = lo;
    if (hi - lo + halo > SMAX_MAX_SHARD_LEN)       /* (what a device holds, not the test hook) */
      return smax_fail(err, errlen, "a plateau that ends in [%lu, %lu) is wider than one device shard "
                       "may be; use more GPUs or the -scan mode", (unsigned long) lo, (unsigned long) hi);
    if (smax_device_upload_halo(dev, idx, lo, hi, halo, with_suf, NULL, err, errlen) != 0)
      return -1;
    if (smax_scan_launch(dev, minlength, opts->policy, with_suf, smax_device_own_stream(dev), err, errlen) != 0)
      return -1;
    rc = smax_scan_counts(dev, &nrecs, NULL, err, errlen);
    if (rc != SMAX_E_RANGE)
      return rc;
    if (halo == lo)
      return -1;                       /* the message of smax_scan_counts stands */
    halo *= 2;
  }
}

/* ---- the parallel variant of the pipeline (run_shards with a producer) -----------
   With a producer, every device gets a second thread next to its uploader: it launches the scan
   of each of the device's shards as soon as the shard is resident, waits for the counts and
   PRODUCES the shard's result (records fetched, positions gathered, text rendered) into a blob,
   while the uploader is at the next shard; the calling thread only CONSUMES the blobs in shard
   order (writes them).  The devices work independently: a shard sees the shards of its own
   device that lie left of it through views, and a plateau that crosses the cut between two
   devices by more than the halo is caught by the kernel (SMAX_E_RANGE) and that shard redone
   with a wider halo -- so N devices render N shards at a time, where the serial consumer did
   one (the host emitter was what kept `-gpus 8` at the speed of two). */
typedef int (*ShardProducer)(void *ctx, int g, smax_device *dev, void **blob, char *err, size_t errlen);

typedef struct
{
  const smax_index *idx;
  const smax_opts *opts;
  smax_device **dev;
  const uint64_t *cut;
  smax_shard_view *views;
  int with_suf, first, count, stride, dev_first;   /* shards first, first + stride, ... (count of them) of the
                                                       device whose shards begin at dev_first */
  pthread_mutex_t *lock;
  pthread_cond_t *cond;
  int *uploaded, *scanned, *done;        /* per shard: 0 pending, 1 ok, -1 failed */
  void **blobs;
  volatile int *abort;
  ShardProducer produce;
  void *ctx;
  char err[512];
} WorkJob;

static void *work_thread(void *arg)
{
  WorkJob *j = arg;
  const uint64_t minlength = j->opts->minlength ? j->opts->minlength : 1;
  int k;
  for (k = 0; k < j->count; k++)
  {
    const int g = j->first + k * j->stride;
    const int nleft = g - j->dev_first < SMAX_MAX_LEFT ? g - j->dev_first : SMAX_MAX_LEFT;
    int state, rc = 0;
    pthread_mutex_lock(j->lock);
    while ((state = j->uploaded[g]) == 0 && !*j->abort)
      pthread_cond_wait(j->cond, j->lock);
    pthread_mutex_unlock(j->lock);
    if (state <= 0)
      rc = -1;                            /* (the uploader, or the worker that failed, has the message) */
    /* The scans of a device go in shard order, one at a time (they take a fraction of what the
       host does with a shard's result): shard g may walk into the shards left of it, so their
       tables must be through with being made resident -- including a redo with a wider halo. */
    if (rc == 0 && g > j->dev_first)
    {
      pthread_mutex_lock(j->lock);
      while ((state = j->scanned[g - 1]) == 0 && !*j->abort)
        pthread_cond_wait(j->cond, j->lock);
      pthread_mutex_unlock(j->lock);
      if (state <= 0)
        rc = -1;
    }
    if (rc == 0)
    {
      if (smax_device_set_left_views(j->dev[g], j->views + (g - nleft), nleft, j->err, sizeof j->err) != 0 ||
          smax_scan_launch(j->dev[g], minlength, j->opts->policy, j->with_suf,
                           smax_device_own_stream(j->dev[g]), j->err, sizeof j->err) != 0)
        rc = -1;
    }
    if (rc == 0)
    {
      uint64_t nrecs;
      int src = smax_scan_counts(j->dev[g], &nrecs, NULL, j->err, sizeof j->err);
      if (src == SMAX_E_RANGE)
        src = redo_with_wider_halo(j->idx, j->opts, j->with_suf, j->dev[g], j->cut[g], j->cut[g + 1],
                                   j->err, sizeof j->err);
      if (src != 0)
        rc = -1;
    }
    if (rc == 0)
      smax_device_view(j->dev[g], &j->views[g]);      /* what the shards right of g see of it from now on */
    pthread_mutex_lock(j->lock);
    j->scanned[g] = rc == 0 ? 1 : -1;
    if (rc != 0)
      *j->abort = 1;
    pthread_cond_broadcast(j->cond);
    pthread_mutex_unlock(j->lock);
    if (getenv("SMAX_TRACE") != NULL)
      fprintf(stderr, "# %9.3f ms  shard %d scanned\n", trace_now(), g);
    if (rc == 0 && j->produce(j->ctx, g, j->dev[g], &j->blobs[g], j->err, sizeof j->err) != 0)
      rc = -1;
    if (getenv("SMAX_TRACE") != NULL)
      fprintf(stderr, "# %9.3f ms  shard %d %s%s\n", trace_now(), g, rc == 0 ? "produced" : "FAILED: ", rc == 0 ? "" : j->err);
    pthread_mutex_lock(j->lock);
    j->done[g] = rc == 0 ? 1 : -1;
    if (rc != 0)
      *j->abort = 1;
    pthread_cond_broadcast(j->cond);
    pthread_mutex_unlock(j->lock);
    if (rc != 0)
      break;
  }
  return NULL;
}

/* shards of [0, n): contiguous ranges of the lcp index space, cut at multiples of 16 by
   cost; consecutive shards share a device when there are more shards than devices */
static int run_shards(const smax_index *idx, const smax_opts *opts, int with_suf,
                      smax_device **dev, int *ordinal, int ngpus, int nshards, int cached,
                      ShardProducer produce, void **blobs,
                      ShardConsumer consume, void *ctx, char *err, size_t errlen)
{
  smax_shard_view views[SMAX_MAX_SHARDS];
  uint64_t cut[SMAX_MAX_SHARDS + 1];
  int uploaded[SMAX_MAX_SHARDS];
  UploadJob job[SMAX_MAX_GPUS];
  pthread_t thr[SMAX_MAX_GPUS];
  int started[SMAX_MAX_GPUS];
  WorkJob wjob[SMAX_MAX_GPUS * SMAX_WORKERS];
  pthread_t wthr[SMAX_MAX_GPUS * SMAX_WORKERS];
  int wstarted[SMAX_MAX_GPUS * SMAX_WORKERS];
  int done[SMAX_MAX_SHARDS], scanned[SMAX_MAX_SHARDS];
  pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
  pthread_cond_t cond = PTHREAD_COND_INITIALIZER;
  volatile int abort_flag = 0;
  const uint64_t minlength = opts->minlength ? opts->minlength : 1;
  const int per_device = nshards / ngpus;
  int g, dv, rc = 0, launched = 0, nworkers = 1;

  if (getenv("SMAX_TRACE") != NULL)
    fprintf(stderr, "# %9.3f ms  run of %d shards on %d device(s) begins\n", trace_now(), nshards, ngpus);
  balanced_cuts(idx, nshards, cut);
  for (g = 0; g < nshards; g++)
  {
    uploaded[g] = 0;
    ordinal[g] = opts->first_device + g / per_device;
    if (smax_cache_acquire(g, ordinal[g], cached, &dev[g], err, errlen) != 0)
      return -1;
  }
  for (dv = 0; dv < ngpus; dv++)
  {
    UploadJob *j = &job[dv];
    j->idx = idx; j->dev = dev; j->cut = cut; j->with_suf = with_suf;
    j->first = dv * per_device; j->step = 1; j->count = per_device;
    j->lock = &lock; j->cond = &cond; j->uploaded = uploaded; j->abort = &abort_flag;
    j->err[0] = '\0';
    started[dv] = pthread_create(&thr[dv], NULL, upload_thread, j) == 0;
    if (!started[dv])
      upload_thread(j);                /* no thread: this device's shards now, one after the other */
  }
  if (produce != NULL)
  {
    /* parallel variant: a worker per device scans and produces, this thread consumes in order */
    for (g = 0; g < nshards; g++)
    {
      done[g] = 0;
      scanned[g] = 0;
      blobs[g] = NULL;
    }
    /* up to SMAX_WORKERS workers per device take the device's shards in turn: the shards of a
       device are produced side by side (a shard with a million repeats takes the host longer
       than its upload) */
    nworkers = per_device < SMAX_WORKERS ? per_device : SMAX_WORKERS;
    for (dv = 0; dv < ngpus * nworkers; dv++)
    {
      WorkJob *w = &wjob[dv];
      const int device = dv / nworkers, turn = dv % nworkers;
      w->idx = idx; w->opts = opts; w->dev = dev; w->cut = cut; w->views = views;
      w->with_suf = with_suf; w->dev_first = device * per_device; w->first = w->dev_first + turn;
      w->stride = nworkers; w->count = (per_device - turn + nworkers - 1) / nworkers;
      w->lock = &lock; w->cond = &cond; w->uploaded = uploaded; w->scanned = scanned; w->done = done; w->blobs = blobs;
      w->abort = &abort_flag; w->produce = produce; w->ctx = ctx;
      w->err[0] = '\0';
      wstarted[dv] = pthread_create(&wthr[dv], NULL, work_thread, w) == 0;
    }
    for (dv = 0; dv < ngpus * nworkers; dv++)
      if (!wstarted[dv])
        work_thread(&wjob[dv]);         /* no thread: these shards now (in shard order per worker) */
    for (g = 0; g < nshards && rc == 0; g++)
    {
      int state;
      pthread_mutex_lock(&lock);
      while ((state = done[g]) == 0 && !abort_flag)
        pthread_cond_wait(&cond, &lock);
      pthread_mutex_unlock(&lock);
      if (state <= 0)
      {
        const char *m = job[g / per_device].err;
        int q;
        for (q = 0; q < ngpus * nworkers && m[0] == '\0'; q++)
          m = wjob[q].err;
        for (q = 0; q < ngpus && m[0] == '\0'; q++)
          m = job[q].err;
        smax_fail(err, errlen, "%s", m[0] ? m : "a shard of the run failed");
        rc = -1;
        break;
      }
      if (consume(ctx, g, dev[g], err, errlen) != 0)
        rc = -1;
      if (getenv("SMAX_TRACE") != NULL)
        fprintf(stderr, "# %9.3f ms  shard %d consumed\n", trace_now(), g);
    }
    if (rc != 0)
    {
      pthread_mutex_lock(&lock);
      abort_flag = 1;
      pthread_cond_broadcast(&cond);
      pthread_mutex_unlock(&lock);
    }
    for (dv = 0; dv < ngpus * nworkers; dv++)
      if (wstarted[dv])
        pthread_join(wthr[dv], NULL);
    for (dv = 0; dv < ngpus; dv++)
      if (started[dv])
        pthread_join(thr[dv], NULL);
    pthread_mutex_destroy(&lock);
    pthread_cond_destroy(&cond);
    return rc;
  }
  for (g = 0; g < nshards && rc == 0; g++)
  {
    int state;
    pthread_mutex_lock(&lock);
    while ((state = uploaded[g]) == 0)
      pthread_cond_wait(&cond, &lock);
    pthread_mutex_unlock(&lock);
    if (state < 0)
    {
      smax_fail(err, errlen, "%s", job[g / per_device].err);
      rc = -1;
      break;
    }
    /* everything up to g is resident: wire and launch what has not been launched yet */
    for (; launched <= g && rc == 0; launched++)
    {
      const int nleft = launched < SMAX_MAX_LEFT ? launched : SMAX_MAX_LEFT;
      smax_device_view(dev[launched], &views[launched]);
      if (smax_device_set_left_views(dev[launched], views + (launched - nleft), nleft, err, errlen) != 0 ||
          smax_scan_launch(dev[launched], minlength, opts->policy, with_suf,
                           smax_device_own_stream(dev[launched]), err, errlen) != 0)
        rc = -1;
    }
    if (rc != 0)
      break;
    {
      uint64_t nrecs;
      int src = smax_scan_counts(dev[g], &nrecs, NULL, err, errlen);
      if (src == SMAX_E_RANGE)
      {
        /* the uploads still running belong to other handles; this shard is redone alone */
        src = redo_with_wider_halo(idx, opts, with_suf, dev[g], cut[g], cut[g + 1], err, errlen);
        if (src == 0)
          smax_device_view(dev[g], &views[g]);     /* (new arrays: what the shards right of g see of it) */
      }
      if (src != 0)
      {
        rc = -1;
        break;
      }
    }
    if (getenv("SMAX_TRACE") != NULL)
      fprintf(stderr, "# %9.3f ms  shard %d scanned\n", trace_now(), g);
    if (consume(ctx, g, dev[g], err, errlen) != 0)
      rc = -1;
    if (getenv("SMAX_TRACE") != NULL)
      fprintf(stderr, "# %9.3f ms  shard %d consumed\n", trace_now(), g);
  }
  if (rc != 0)
    abort_flag = 1;
  for (dv = 0; dv < ngpus; dv++)
    if (started[dv])
      pthread_join(thr[dv], NULL);
  pthread_mutex_destroy(&lock);
  pthread_cond_destroy(&cond);
  return rc;
}

/* shards per device of one run: what the kernel's range demands, and at least SMAX_PIPELINE
   (default 4) of them when a device's share is large enough for the overlap to pay */
static int shards_per_device(uint64_t n, int ngpus, int needed_total)
{
  const char *e = getenv("SMAX_PIPELINE");
  int want = e != NULL ? atoi(e) : 4, per = needed_total / ngpus;
  if (want < 1) want = 1;
  if (n / (uint64_t) ngpus < ((uint64_t) 1 << 24) && e == NULL)
    want = 1;
  if (per < want) per = want;
  while ((uint64_t) per * (uint64_t) ngpus > SMAX_MAX_SHARDS && per > 1)
    per--;
  return per;
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
  if (*nshards_out > 0)
    *nshards_out = ngpus * shards_per_device(idx->info.numberofallsortedsuffixes, ngpus, *nshards_out);
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

/* consumer of smax_run_records: the shard's records are appended */
typedef struct
{
  smax_record *recs;
  uint64_t n, cap;
} RecordSink;

static int sink_records(void *ctx, int g, smax_device *dev, char *err, size_t errlen)
{
  RecordSink *k = ctx;
  uint64_t cnt;
  (void) g;
  if (smax_scan_counts(dev, &cnt, NULL, err, errlen) != 0)
    return -1;
  if (cnt == 0)
    return 0;
  if (k->n + cnt > k->cap)
  {
    const uint64_t cap = k->n + cnt > 2 * k->cap ? k->n + cnt : 2 * k->cap;
    smax_record *p = realloc(k->recs, cap * sizeof *p);
    if (p == NULL)
      return smax_fail(err, errlen, "out of memory for %lu records", (unsigned long) cap);
    k->recs = p;
    k->cap = cap;
  }
  if (smax_scan_fetch(dev, k->recs + k->n, NULL, err, errlen) != 0)
    return -1;
  k->n += cnt;
  return 0;
}

int smax_run_records(const smax_index *idx, const smax_opts *opts, smax_record **recs_out,
                     uint64_t *nrecs_out, char *err, size_t errlen)
{
  smax_device *dev[SMAX_MAX_SHARDS];
  int ordinal[SMAX_MAX_SHARDS];
  RecordSink sink = {NULL, 0, 0};
  int ngpus = 1, nshards = 1, rc, empty = 0, cached;

  if (idx == NULL || opts == NULL || recs_out == NULL || nrecs_out == NULL)
    return smax_fail(err, errlen, "smax_run: null argument");
  *recs_out = NULL;
  *nrecs_out = 0;
  if (check_run_args(idx, opts, &ngpus, &nshards, &empty, err, errlen) != 0)
    return -1;
  if (empty)
    return 0;
  memset(dev, 0, sizeof dev);
  cached = smax_cache_begin();
  rc = run_shards(idx, opts, 0, dev, ordinal, ngpus, nshards, cached, NULL, NULL, sink_records, &sink, err, errlen);
  if (rc == 0)
  {
    *recs_out = sink.recs;
    *nrecs_out = sink.n;
  } else
    free(sink.recs);
  smax_cache_end(dev, ordinal, nshards, cached, rc != 0);
  return rc;
}

/* consumer of smax_run_text: the shard's records rendered on its device, the bytes written */
typedef struct
{
  const smax_opts *opts;
  FILE *fp;
  const uint64_t *seps;
  uint64_t nseps, total;
  int with_suf;
  char *buf;
  size_t bufcap;
} TextSink;

static int sink_text(void *ctx, int g, smax_device *dev, char *err, size_t errlen)
{
  TextSink *k = ctx;
  uint64_t bytes = 0;
  (void) g;
  if (k->with_suf && k->opts->relative &&
      smax_device_set_separators(dev, k->seps, k->nseps, err, errlen) != 0)
    return -1;
  if (smax_scan_format(dev, k->opts->format, k->opts->relative, &bytes, err, errlen) != 0)
    return -1;
  if (bytes == 0)
    return 0;
  if (bytes > k->bufcap)
  {
    char *p = realloc(k->buf, bytes);
    if (p == NULL)
      return smax_fail(err, errlen, "out of memory for %lu bytes of text", (unsigned long) bytes);
    k->buf = p;
    k->bufcap = bytes;
  }
  if (smax_scan_fetch_text(dev, k->buf, err, errlen) != 0)
    return -1;
  if (fwrite(k->buf, 1, bytes, k->fp) != bytes)
    return smax_fail(err, errlen, "cannot write results");
  k->total += bytes;
  return 0;
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
  TextSink sink;
  int ngpus = 1, nshards = 1, rc, empty = 0;

  if (idx == NULL || opts == NULL)
    return smax_fail(err, errlen, "smax_run_text: null argument");
  if (nbytes != NULL) *nbytes = 0;
  if (opts->format != SMAX_FORMAT_SMAX && opts->format != SMAX_FORMAT_ITV)
    return smax_fail(err, errlen, "the device formatter renders the smax and itv formats; "
                     "format %d is rendered by the host emitter", (int) opts->format);
  memset(&sink, 0, sizeof sink);
  sink.opts = opts;
  sink.fp = file != NULL ? (FILE *) file : stdout;
  sink.with_suf = opts->format == SMAX_FORMAT_SMAX;
  if (sink.with_suf && idx->suf == NULL)
    return smax_fail(err, errlen, "the index was opened without the suffix table");
  if (check_run_args(idx, opts, &ngpus, &nshards, &empty, err, errlen) != 0)
    return -1;
  if (empty)
    return 0;
  if (sink.with_suf && opts->relative &&
      smax_index_separators((smax_index *) idx, &sink.seps, &sink.nseps, err, errlen) != 0)
    return -1;
  memset(dev, 0, sizeof dev);
  cached = smax_cache_begin();
  rc = run_shards(idx, opts, sink.with_suf, dev, ordinal, ngpus, nshards, cached, NULL, NULL, sink_text, &sink,
                  err, errlen);
  if (rc == 0 && nbytes != NULL) *nbytes = sink.total;
  free(sink.buf);
  smax_cache_end(dev, ordinal, nshards, cached, rc != 0);
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

/* consumer of smax_run: every repeat of the shard goes to the caller's callback, positions from
   the host suffix table, while the next shard is uploaded and scanned */
typedef struct
{
  const smax_index *idx;
  smax_emit_cb cb;
  void *info;
  smax_record *recs;
  uint64_t cap, poscap, *pos;
} CallbackSink;

static int sink_callback(void *ctx, int g, smax_device *dev, char *err, size_t errlen)
{
  CallbackSink *k = ctx;
  const smax_index *idx = k->idx;
  uint64_t nrecs, r, j;
  (void) g;
  if (smax_scan_counts(dev, &nrecs, NULL, err, errlen) != 0)
    return -1;
  if (nrecs == 0)
    return 0;
  if (nrecs > k->cap)
  {
    smax_record *p = realloc(k->recs, nrecs * sizeof *p);
    if (p == NULL)
      return smax_fail(err, errlen, "out of memory for %lu records", (unsigned long) nrecs);
    k->recs = p;
    k->cap = nrecs;
  }
  if (smax_scan_fetch(dev, k->recs, NULL, err, errlen) != 0)
    return -1;
  for (r = 0; r < nrecs; r++)
  {
    const uint64_t w = k->recs[r].width, lb = k->recs[r].lb;
    if (idx->suf != NULL)
    {
      /* the suffix table entries of a repeat are a random access into a table of 8 n bytes:
         ask for those of a later record now (a cache miss per record otherwise costs more
         than everything else the host does for it) */
      if (r + SMAX_PREFETCH_AHEAD < nrecs)
      {
        const uint64_t plb = k->recs[r + SMAX_PREFETCH_AHEAD].lb;
        const char *pa = (const char *) idx->suf + plb * idx->info.sufbytes;
        __builtin_prefetch(pa, 0, 0);
        __builtin_prefetch(pa + 64, 0, 0);
      }
      if (w > k->poscap)
      {
        uint64_t *p = realloc(k->pos, w * sizeof *p);
        if (p == NULL)
          return smax_fail(err, errlen, "out of memory");
        k->pos = p;
        k->poscap = w;
      }
      /* occurrence positions in suffix-array order */
      if (idx->info.sufbytes == 8)
        memcpy(k->pos, (const uint64_t *) idx->suf + lb, w * sizeof *k->pos);
      else
        for (j = 0; j < w; j++)
          k->pos[j] = ((const uint32_t *) idx->suf)[lb + j];
    }
    if (k->cb(k->info, k->recs[r].len, lb, w, idx->suf != NULL ? k->pos : NULL) != 0)
      return smax_fail(err, errlen, "result callback failed");
  }
  return 0;
}

/* producer / consumer of smax_run when the callback is the library's own emitter: the text of a
   shard is rendered by the shard's worker thread through a clone of the emitter (same format,
   same code), the calling thread writes the shards' texts in order */
typedef struct
{
  const smax_index *idx;
  smax_emitter *em;
  void *blobs[SMAX_MAX_SHARDS];
} EmitSink;

typedef struct
{
  char *text;
  size_t len;
} TextBlob;

static int produce_emit_text(void *ctx, int g, smax_device *dev, void **blob, char *err, size_t errlen)
{
  EmitSink *k = ctx;
  const smax_index *idx = k->idx;
  const int want_pos = smax_emitter_wants_positions(k->em) && idx->suf != NULL;
  smax_emitter *clone = NULL;
  smax_record *recs = NULL;
  uint64_t *pos = NULL, poscap = 0, nrecs, r, j;
  TextBlob *b;
  int rc = 0;
  (void) g;
  if (smax_scan_counts(dev, &nrecs, NULL, err, errlen) != 0)
    return -1;
  b = calloc(1, sizeof *b);
  if (b == NULL)
    return smax_fail(err, errlen, "out of memory");
  *blob = b;
  if (nrecs == 0)
    return 0;
  recs = malloc(nrecs * sizeof *recs);
  if (recs == NULL)
    return smax_fail(err, errlen, "out of memory for %lu records", (unsigned long) nrecs);
  if (smax_scan_fetch(dev, recs, NULL, err, errlen) != 0 ||
      smax_emitter_clone_mem(k->em, &clone, err, errlen) != 0)
  {
    free(recs);
    return -1;
  }
  for (r = 0; r < nrecs && rc == 0; r++)
  {
    const uint64_t w = recs[r].width, lb = recs[r].lb;
    if (want_pos)
    {
      if (r + SMAX_PREFETCH_AHEAD < nrecs)
      {
        const char *pa = (const char *) idx->suf + recs[r + SMAX_PREFETCH_AHEAD].lb * idx->info.sufbytes;
        __builtin_prefetch(pa, 0, 0);
        __builtin_prefetch(pa + 64, 0, 0);
      }
      if (w > poscap)
      {
        uint64_t *p = realloc(pos, w * sizeof *p);
        if (p == NULL)
        {
          rc = smax_fail(err, errlen, "out of memory");
          break;
        }
        pos = p;
        poscap = w;
      }
      if (idx->info.sufbytes == 8)
        memcpy(pos, (const uint64_t *) idx->suf + lb, w * sizeof *pos);
      else
        for (j = 0; j < w; j++)
          pos[j] = ((const uint32_t *) idx->suf)[lb + j];
    }
    if (smax_emitter_emit(clone, recs[r].len, lb, w, want_pos ? pos : NULL) != 0)
      rc = smax_fail(err, errlen, "the result emitter failed");
  }
  if (smax_emitter_finish_mem(clone, &b->text, &b->len) != 0 && rc == 0)
    rc = smax_fail(err, errlen, "the result emitter failed");
  free(recs);
  free(pos);
  return rc;
}

static int consume_emit_text(void *ctx, int g, smax_device *dev, char *err, size_t errlen)
{
  EmitSink *k = ctx;
  TextBlob *b = k->blobs[g];
  int rc = 0;
  (void) dev;
  if (b != NULL)
  {
    if (b->len > 0 && smax_emitter_write_raw(k->em, b->text, b->len) != 0)
      rc = smax_fail(err, errlen, "cannot write results");
    free(b->text);
    free(b);
    k->blobs[g] = NULL;
  }
  return rc;
}

int smax_run(const smax_index *idx, const smax_opts *opts, smax_emit_cb cb, void *info,
             char *err, size_t errlen)
{
  smax_device *dev[SMAX_MAX_SHARDS];
  int ordinal[SMAX_MAX_SHARDS];
  CallbackSink sink;
  int ngpus = 1, nshards = 1, rc, empty = 0, cached;

  if (cb == NULL)
    return smax_fail(err, errlen, "smax_run: null callback");
  if (idx == NULL || opts == NULL)
    return smax_fail(err, errlen, "smax_run: null argument");
  if (check_run_args(idx, opts, &ngpus, &nshards, &empty, err, errlen) != 0)
    return -1;
  if (empty)
    return 0;
  memset(dev, 0, sizeof dev);
  if (cb == smax_emitter_emit && info != NULL && idx->base == 0 && getenv("SMAX_SERIAL_EMIT") == NULL)
  {
    /* the library's emitter: the shards' texts are rendered in parallel, written in order */
    EmitSink *es = calloc(1, sizeof *es);
    int g;
    if (es == NULL)
      return smax_fail(err, errlen, "out of memory");
    es->idx = idx;
    es->em = info;
    if (smax_emitter_is_relative(es->em))
    {
      /* (the separator table is built on first use: before the threads start) */
      uint64_t sq, rp;
      if (smax_index_seqnum_relpos((smax_index *) idx, 0, &sq, &rp, err, errlen) != 0)
      {
        free(es);
        return -1;
      }
    }
    cached = smax_cache_begin();
    rc = run_shards(idx, opts, 0, dev, ordinal, ngpus, nshards, cached, produce_emit_text, es->blobs,
                    consume_emit_text, es, err, errlen);
    for (g = 0; g < nshards; g++)
      if (es->blobs[g] != NULL)
      {
        free(((TextBlob *) es->blobs[g])->text);
        free(es->blobs[g]);
      }
    free(es);
    smax_cache_end(dev, ordinal, nshards, cached, rc != 0);
    return rc;
  }
  memset(&sink, 0, sizeof sink);
  sink.idx = idx; sink.cb = cb; sink.info = info;
  cached = smax_cache_begin();
  rc = run_shards(idx, opts, 0, dev, ordinal, ngpus, nshards, cached, NULL, NULL, sink_callback, &sink, err, errlen);
  free(sink.recs);
  free(sink.pos);
  smax_cache_end(dev, ordinal, nshards, cached, rc != 0);
  return rc;
}
