/*
  smax_stream.c -- the "-scan" mode of the tool: the
  tables are never mapped; the suffix-array range is cut into chunks whose
  table bytes are read straight from the index files, made resident on ONE
  GPU, scanned and replaced by the next chunk, so that neither host memory
  nor HBM has to hold the index.

  Reference analogue: streamsuffixarray + the GtBufferedfile_* readers
  (/root/reference/src/match/esa-map.c:488-501,
  /root/reference/src/match/sarr-def.h:49-95,
  /root/reference/src/match/esa-seqread.h:50-94): tables are consumed front
  to back through bounded buffers, .llv through a sequential cursor
  (sarr-def.h:128-178).

  A chunk is a shard of the multi-GPU driver that happens to live on the same
  device as its neighbours: chunk c owns the plateaus that END in its range
  and walks left into the tables of chunks c-1 and c-2 through the same peer
  views (three device handles take turns).  A plateau that reaches further
  back is rare; its chunk is redone with a window extended to the left until
  the plateau fits.  Results are delivered chunk by chunk, i.e. in ascending
  left boundary as in smax_run.
*/
#define _FILE_OFFSET_BITS 64
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>
#include "smax_host.h"

#define STREAM_DEVS 3
#define STREAM_HALO 256         /* left coverage smax_device_upload asks for */
#define STREAM_DEFAULT_CHUNK ((uint64_t) 1 << 28)

typedef struct
{
  int fd;
  uint64_t bytes;
  char path[4096];
} Tabfile;

static int tab_open(Tabfile *t, const char *indexname, const char *suffix, char *err,
                    size_t errlen)
{
  struct stat st;
  snprintf(t->path, sizeof t->path, "%s%s", indexname, suffix);
  t->fd = open(t->path, O_RDONLY);
  if (t->fd < 0)   /* wording of gt_fa_fopen's failure, as in smax_index.c */
    return smax_fail(err, errlen, "fopen(): cannot open file '%s': %s", t->path,
                     strerror(errno));
  if (fstat(t->fd, &st) != 0)
    return smax_fail(err, errlen, "cannot stat file '%s': %s", t->path, strerror(errno));
  t->bytes = (uint64_t) st.st_size;
  return 0;
}

static void tab_close(Tabfile *t)
{
  if (t->fd >= 0)
    close(t->fd);
  t->fd = -1;
}

static int tab_read_part(const Tabfile *t, void *dst, uint64_t offset, uint64_t bytes, char *err,
                         size_t errlen)
{
  uint64_t done = 0;
  while (done < bytes)
  {
    const ssize_t got = pread(t->fd, (char *) dst + done, bytes - done, (off_t) (offset + done));
    if (got < 0 && errno == EINTR)
      continue;
    if (got <= 0)
      return smax_fail(err, errlen, "cannot read %lu bytes at offset %lu of file '%s': %s",
                       (unsigned long) (bytes - done), (unsigned long) (offset + done), t->path,
                       got == 0 ? "unexpected end of file" : strerror(errno));
    done += (uint64_t) got;
  }
  return 0;
}

/* A chunk of a table is hundreds of megabytes, and one pread of a page-cache-warm file is one
   thread copying: large reads are cut into slices read by several threads (pread takes its own
   offset, so the descriptor is shared). */
#ifndef TAB_READ_SLICE                       /* (tests/test_host_asan.py builds with 4096: every read is sliced) */
#define TAB_READ_SLICE ((uint64_t) 8 << 20)
#endif
#define TAB_READ_THREADS 8

typedef struct
{
  const Tabfile *t;
  void *dst;
  uint64_t offset, bytes;
  int rc;
  char err[256];
} ReadJob;

static void *tab_read_thread(void *arg)
{
  ReadJob *j = arg;
  j->rc = tab_read_part(j->t, j->dst, j->offset, j->bytes, j->err, sizeof j->err);
  return NULL;
}

static int tab_read(const Tabfile *t, void *dst, uint64_t offset, uint64_t bytes, char *err,
                    size_t errlen)
{
  ReadJob job[TAB_READ_THREADS];
  pthread_t thr[TAB_READ_THREADS];
  int started[TAB_READ_THREADS];
  uint64_t per;
  int n, k, rc = 0;
  if (bytes < 2 * TAB_READ_SLICE)
    return tab_read_part(t, dst, offset, bytes, err, errlen);
  n = (int) (bytes / TAB_READ_SLICE);
  if (n > TAB_READ_THREADS)
    n = TAB_READ_THREADS;
  per = ((bytes / (uint64_t) n) + 4095) & ~(uint64_t) 4095;
  for (k = 0; k < n; k++)
  {
    const uint64_t lo = (uint64_t) k * per;
    job[k].t = t;
    job[k].dst = (char *) dst + lo;
    job[k].offset = offset + lo;
    job[k].bytes = lo >= bytes ? 0 : (bytes - lo < per || k == n - 1 ? bytes - lo : per);
    job[k].rc = 0;
    job[k].err[0] = '\0';
    started[k] = job[k].bytes > 0 && k > 0 &&
                 pthread_create(&thr[k], NULL, tab_read_thread, &job[k]) == 0;
  }
  for (k = 0; k < n; k++)
    if (!started[k] && job[k].bytes > 0)
      tab_read_thread(&job[k]);              /* the first slice, and what got no thread */
  for (k = 0; k < n; k++)
  {
    if (started[k])
      pthread_join(thr[k], NULL);
    if (job[k].rc != 0 && rc == 0)
      rc = smax_fail(err, errlen, "%s", job[k].err);
  }
  return rc;
}

static int units_check(const Tabfile *t, uint64_t expected, unsigned unit, char *err,
                       size_t errlen)
{
  if (t->bytes / unit != expected || t->bytes % unit != 0)
    return smax_fail(err, errlen, "file %s: number of units (of size %u) = %lu != %lu = "
                     "expected number of units", t->path, unit,
                     (unsigned long) (t->bytes / unit), (unsigned long) expected);
  return 0;
}

static int suf_read(const Tabfile *suf, unsigned sufbytes, uint64_t first, uint64_t count,
                    uint64_t *out, char *err, size_t errlen)
{
  if (sufbytes == 8)
    return tab_read(suf, out, first * 8, count * 8, err, errlen);
  /* 4-byte entries (-suftabuint): widen in place, back to front */
  if (tab_read(suf, out, first * 4, count * 4, err, errlen) != 0)
    return -1;
  while (count-- > 0)
    out[count] = ((const uint32_t *) out)[count];
  return 0;
}

static int cmp_u64(const void *a, const void *b)
{
  const uint64_t x = *(const uint64_t *) a, y = *(const uint64_t *) b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

/* separator table for relative output without mapped tables: one pass over
   the bwt file, one suffix-table read per separator (smax_index.c build_seps,
   gt_encseq_seqnum semantics /root/reference/src/core/encseq.c:3815-3900) */
static int stream_seps(smax_index *idx, const Tabfile *bwt, const Tabfile *suf,
                       unsigned sufbytes, char *err, size_t errlen)
{
  const uint64_t n = idx->info.numberofallsortedsuffixes, block = 1u << 22;
  uint64_t cap = idx->info.numofsequences + 1, cnt = 0, lo, i;
  uint8_t *buf;
  if (idx->seps_ready)
    return 0;
  if (idx->info.numofsequences <= 1)
  {
    idx->seps_ready = 1;
    return 0;
  }
  buf = malloc(block);
  idx->seps = malloc(cap * sizeof (uint64_t));
  if (buf == NULL || idx->seps == NULL)
  {
    free(buf);
    return smax_fail(err, errlen, "out of memory");
  }
  for (lo = 0; lo < n; lo += block)
  {
    const uint64_t len = n - lo < block ? n - lo : block;
    if (tab_read(bwt, buf, lo, len, err, errlen) != 0)
    {
      free(buf);
      return -1;
    }
    for (i = 0; i < len; i++)
    {
      uint64_t pos;
      if (buf[i] != 255)
        continue;
      if (cnt == cap)
      {
        uint64_t *p = realloc(idx->seps, 2 * cap * sizeof (uint64_t));
        if (p == NULL)
        {
          free(buf);
          return smax_fail(err, errlen, "out of memory");
        }
        idx->seps = p;
        cap *= 2;
      }
      if (suf_read(suf, sufbytes, lo + i, 1, &pos, err, errlen) != 0)
      {
        free(buf);
        return -1;
      }
      idx->seps[cnt++] = pos - 1;
    }
  }
  free(buf);
  qsort(idx->seps, cnt, sizeof (uint64_t), cmp_u64);
  idx->nseps = cnt;
  idx->seps_ready = 1;
  return 0;
}

/* table bytes of one window of the SA range, read from the files */
typedef struct
{
  uint8_t *lcp, *bwt;
  smax_llv *llv;
  uint64_t cap, llvcap, nllv;
} Window;

static void window_free(Window *w)
{
  free(w->lcp); free(w->bwt); free(w->llv);
  memset(w, 0, sizeof *w);
}

/* index of the first .llv record with position >= pos (records ascend by
   position, lcpoverflow.h:24-30): binary search with one small read per step */
static int llv_lower_bound(const Tabfile *llvf, uint64_t L, uint64_t pos, uint64_t *k,
                           char *err, size_t errlen)
{
  uint64_t lo = 0, hi = L;
  while (lo < hi)
  {
    const uint64_t mid = lo + (hi - lo) / 2;
    smax_llv one;
    if (tab_read(llvf, &one, mid * sizeof one, sizeof one, err, errlen) != 0)
      return -1;
    if (one.position < pos) lo = mid + 1; else hi = mid;
  }
  *k = lo;
  return 0;
}

/* lcp / bwt bytes of [w_lo, w_hi) and the .llv records that belong to its 255
   entries (their number = the number of 255 bytes just read: what the
   sequential reader of sarr-def.h:128-178 relies on) */
static int window_load(Window *w, const Tabfile *lcpf, const Tabfile *bwtf, const Tabfile *llvf,
                       uint64_t L, uint64_t w_lo, uint64_t w_hi, char *err, size_t errlen)
{
  const uint64_t len = w_hi - w_lo;
  uint64_t cnt255 = 0, i, k0 = 0;
  if (len + 64 > w->cap)
  {
    uint8_t *a = realloc(w->lcp, len + 64), *b;
    if (a != NULL) w->lcp = a;
    b = realloc(w->bwt, len + 64);
    if (b != NULL) w->bwt = b;
    if (a == NULL || b == NULL)
      return smax_fail(err, errlen, "out of memory for a window of %lu suffixes",
                       (unsigned long) len);
    w->cap = len + 64;
  }
  if (tab_read(lcpf, w->lcp, w_lo, len, err, errlen) != 0 ||
      tab_read(bwtf, w->bwt, w_lo, len, err, errlen) != 0)
    return -1;
  /* the number of 255 entries, eight bytes per step: a byte is 255 iff all its bits are set */
  for (i = 0; i + 8 <= len; i += 8)
  {
    uint64_t x;
    memcpy(&x, w->lcp + i, 8);
    x &= x >> 4; x &= x >> 2; x &= x >> 1;
    x &= 0x0101010101010101ull;
    cnt255 += (x * 0x0101010101010101ull) >> 56;
  }
  for (; i < len; i++)
    cnt255 += w->lcp[i] == 255;
  w->nllv = cnt255;
  if (cnt255 == 0)
    return 0;
  if (cnt255 > w->llvcap)
  {
    smax_llv *p = realloc(w->llv, (cnt255 + cnt255 / 4 + 1) * sizeof *p);
    if (p == NULL)
      return smax_fail(err, errlen, "out of memory");
    w->llv = p;
    w->llvcap = cnt255 + cnt255 / 4 + 1;
  }
  if (llv_lower_bound(llvf, L, w_lo, &k0, err, errlen) != 0)
    return -1;
  if (k0 + cnt255 > L)
    return smax_fail(err, errlen, "inconsistent ESA tables: the lcp table holds more 255 "
                     "entries than the .llv file has records");
  if (tab_read(llvf, w->llv, k0 * sizeof (smax_llv), cnt255 * sizeof (smax_llv), err, errlen) != 0)
    return -1;
  if (w->llv[0].position < w_lo || w->llv[cnt255 - 1].position >= w_hi)
    return smax_fail(err, errlen, "inconsistent ESA tables: .llv records do not match the 255 "
                     "entries of the lcp table");
  return 0;
}

/* the table bytes of the NEXT chunk are read by a second thread while this one is uploaded,
   scanned and its results are handed out */
typedef struct
{
  Window *w;
  const Tabfile *lcpf, *bwtf, *llvf;
  uint64_t L, w_lo, w_hi;
  int rc;
  char err[256];
} Prefetch;

static void *prefetch_thread(void *arg)
{
  Prefetch *p = arg;
  p->rc = window_load(p->w, p->lcpf, p->bwtf, p->llvf, p->L, p->w_lo, p->w_hi, p->err, sizeof p->err);
  return NULL;
}

/* suffix table entries of the records r0 .. r1-1 (ascending left boundaries) that lie close
   together, with ONE read: out receives suf[lb(r0) .. lb(r1-1) + width(r1-1)) */
#define STREAM_SUF_GAP 1024       /* entries between two records that are still read together */
#define STREAM_SUF_BLOCK 32768    /* entries of one read */

int smax_run_stream(smax_index *idx, const smax_opts *opts, uint64_t chunk, smax_emit_cb cb,
                    void *info, char *err, size_t errlen)
{
  Tabfile lcpf = { -1, 0, "" }, bwtf = { -1, 0, "" }, llvf = { -1, 0, "" }, suff = { -1, 0, "" };
  smax_device *dev[STREAM_DEVS];
  smax_shard_view views[STREAM_DEVS], left[STREAM_DEVS];
  Window w[2];
  Prefetch pre;
  pthread_t pre_thr;
  int pre_running = 0, pre_for = -1;
  smax_record *recs = NULL;
  uint64_t *pos = NULL;
  uint64_t n, L, reccap = 0, poscap = 0, lo, minlength, c = 0;
  unsigned sufbytes = 0;
  int want_pos, rc = -1, g, navail, cached = -1, ordinal[STREAM_DEVS];

  if (idx == NULL || opts == NULL || cb == NULL)
    return smax_fail(err, errlen, "smax_run_stream: null argument");
  if (idx->indexname == NULL)
    return smax_fail(err, errlen, "smax_run_stream: the index has no files behind it");
  memset(dev, 0, sizeof dev);
  memset(w, 0, sizeof w);
  memset(&pre, 0, sizeof pre);
  n = idx->info.numberofallsortedsuffixes;
  L = idx->info.largelcpvalues;
  minlength = opts->minlength ? opts->minlength : 1;
  want_pos = opts->format != SMAX_FORMAT_ITV;
  if (chunk == 0)
    chunk = STREAM_DEFAULT_CHUNK;
  chunk = (chunk + 15) & ~(uint64_t) 15;
  if (chunk < 1024)
    chunk = 1024;           /* the left halo of a chunk must lie in its neighbour */
  if (chunk > (SMAX_MAX_SHARD_LEN & ~(uint64_t) 15))
    chunk = SMAX_MAX_SHARD_LEN & ~(uint64_t) 15;     /* what one device shard may hold */
  if (idx->info.maxbranchdepth > 0 && minlength > idx->info.maxbranchdepth)
    return 0;
  if (tab_open(&lcpf, idx->indexname, ".lcp", err, errlen) != 0 ||
      units_check(&lcpf, n, 1, err, errlen) != 0 ||
      tab_open(&bwtf, idx->indexname, ".bwt", err, errlen) != 0 ||
      units_check(&bwtf, idx->info.totallength + 1, 1, err, errlen) != 0)
    goto done;
  if (L > 0 && (tab_open(&llvf, idx->indexname, ".llv", err, errlen) != 0 ||
                units_check(&llvf, L, sizeof (smax_llv), err, errlen) != 0))
    goto done;
  if (want_pos)
  {
    if (tab_open(&suff, idx->indexname, ".suf", err, errlen) != 0)
      goto done;
    if (n > 0 && suff.bytes == n * 4)
      sufbytes = 4;
    else if (units_check(&suff, n, 8, err, errlen) != 0)
      goto done;
    else
      sufbytes = 8;
    if (opts->relative && stream_seps(idx, &bwtf, &suff, sufbytes, err, errlen) != 0)
      goto done;
  }
  navail = smax_device_count(err, errlen);
  if (navail < 0)
    goto done;
  if (opts->first_device < 0 || opts->first_device >= navail)
  {
    smax_fail(err, errlen, "device %d requested, but only %d visible", opts->first_device,
              navail);
    goto done;
  }
  cached = smax_cache_begin();
  for (g = 0; g < STREAM_DEVS; g++)
  {
    ordinal[g] = opts->first_device;
    if (smax_cache_acquire(g, ordinal[g], cached, &dev[g], err, errlen) != 0)
      goto done;
  }

  for (lo = 0; lo < n; lo += chunk, c++)
  {
    const uint64_t hi = n - lo < chunk ? n : lo + chunk;
    const uint64_t w_hi = hi + 16 < n ? hi + 16 : n;
    uint64_t nrecs = 0, r, npos = 0, halo = STREAM_HALO;
    smax_device *d = dev[c % STREAM_DEVS];
    Window *cw = &w[c & 1];
    int loaded = 0;

    /* this chunk's window: read ahead by the second thread, or read now */
    if (pre_running)
    {
      pthread_join(pre_thr, NULL);
      pre_running = 0;
      if (pre_for == (int) (c & 0x3fffffff))
      {
        if (pre.rc != 0)
        {
          smax_fail(err, errlen, "%s", pre.err);
          goto done;
        }
        loaded = 1;
      }
    }
    if (hi < n)
    {
      /* the next chunk's bytes are read while this chunk is worked on */
      const uint64_t nlo = hi, nhi = n - nlo < chunk ? n : nlo + chunk;
      pre.w = &w[(c + 1) & 1];
      pre.lcpf = &lcpf; pre.bwtf = &bwtf; pre.llvf = &llvf; pre.L = L;
      pre.w_lo = nlo >= STREAM_HALO ? nlo - STREAM_HALO : 0;
      pre.w_hi = nhi + 16 < n ? nhi + 16 : n;
      pre.rc = 0; pre.err[0] = '\0';
      pre_for = (int) ((c + 1) & 0x3fffffff);
      pre_running = pthread_create(&pre_thr, NULL, prefetch_thread, &pre) == 0;
      if (!pre_running)
        pre_for = -1;
    }

    /* usual case: the chunk with the standard halo, the two chunks before it as
       left views.  A plateau that reaches further back makes the scan report
       that it left the resident range: the chunk is then redone with its own
       window extended to the left (doubling) until the plateau fits -- at worst
       from the start of the table, where any error left is a real inconsistency */
    for (;;)
    {
      const uint64_t w_lo = lo >= halo ? lo - halo : 0;
      smax_index *win = NULL;
      int nleft = 0, failed = 1;
      char scanerr[256] = "";
      if (!(loaded && halo == STREAM_HALO) &&
          window_load(cw, &lcpf, &bwtf, &llvf, L, w_lo, w_hi, err, errlen) != 0)
        goto done;
      if (smax_index_from_memory_window(cw->lcp, cw->bwt, cw->llv, cw->nllv, NULL, 8, w_lo, w_hi - w_lo, n,
                                        &win, err, errlen) != 0)
        goto done;
      if (smax_device_upload_halo(d, win, lo, hi, halo, 0, NULL, err, errlen) != 0)
      {
        smax_index_close(win);
        goto done;
      }
      smax_index_close(win);
      smax_device_view(d, &views[c % STREAM_DEVS]);
      if (halo == STREAM_HALO)
      {
        /* the two chunks before this one stay resident: left views, nearest last */
        /* (a neighbour that was redone with a wide window covers the one before it) */
        if (c >= 2 && views[(c - 2) % STREAM_DEVS].a_lo < views[(c - 1) % STREAM_DEVS].a_lo)
          left[nleft++] = views[(c - 2) % STREAM_DEVS];
        if (c >= 1) left[nleft++] = views[(c - 1) % STREAM_DEVS];
      }
      if (smax_device_set_left_views(d, left, nleft, err, errlen) != 0 ||
          smax_scan_launch(d, minlength, opts->policy, 0, NULL, err, errlen) != 0)
        goto done;
      {
        /* only "a plateau leaves the resident range" is answered with a wider window;
           every other failure (CUDA, inconsistent tables) is final */
        const int src = smax_scan_counts(d, &nrecs, NULL, scanerr, sizeof scanerr);
        if (src == 0)
          failed = 0;
        if (!failed)
          break;
        if (src != SMAX_E_RANGE || w_lo == 0)
        {
          smax_fail(err, errlen, "%s", scanerr);
          goto done;
        }
      }
      halo = halo == STREAM_HALO ? 2 * chunk + STREAM_HALO : 2 * halo;
    }
    if (nrecs > reccap)
    {
      smax_record *p = realloc(recs, (nrecs + nrecs / 4) * sizeof *recs);
      if (p == NULL)
      {
        smax_fail(err, errlen, "out of memory for %lu records", (unsigned long) nrecs);
        goto done;
      }
      recs = p;
      reccap = nrecs + nrecs / 4;
    }
    if (nrecs > 0 && smax_scan_fetch(d, recs, NULL, err, errlen) != 0)
      goto done;
    /* occurrence positions: the suffix table entries of records that lie close together are
       read with one call (a read per record costs more than the bytes between them) */
    for (r = 0; r < nrecs; )
    {
      uint64_t r1 = r + 1, first = recs[r].lb, last = recs[r].lb + recs[r].width, q;
      if (want_pos)
      {
        while (r1 < nrecs && recs[r1].lb <= last + STREAM_SUF_GAP &&
               recs[r1].lb + recs[r1].width - first <= STREAM_SUF_BLOCK)
        {
          last = recs[r1].lb + recs[r1].width;
          r1++;
        }
        npos = last - first;
        if (npos > poscap)
        {
          uint64_t *p = realloc(pos, (npos + npos / 4) * sizeof *pos);
          if (p == NULL)
          {
            smax_fail(err, errlen, "out of memory");
            goto done;
          }
          pos = p;
          poscap = npos + npos / 4;
        }
        if (suf_read(&suff, sufbytes, first, npos, pos, err, errlen) != 0)
          goto done;
      }
      for (q = r; q < r1; q++)
        if (cb(info, recs[q].len, recs[q].lb, recs[q].width,
               want_pos ? pos + (recs[q].lb - first) : NULL) != 0)
        {
          smax_fail(err, errlen, "result callback failed");
          goto done;
        }
      r = r1;
    }
  }
  rc = 0;
done:
  if (pre_running)
    pthread_join(pre_thr, NULL);
  if (cached >= 0)
    smax_cache_end(dev, ordinal, STREAM_DEVS, cached, rc != 0);
  window_free(&w[0]);
  window_free(&w[1]);
  free(recs); free(pos);
  tab_close(&lcpf); tab_close(&bwtf); tab_close(&llvf); tab_close(&suff);
  return rc;
}
