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
  views (three device handles take turns).  Results are delivered chunk by
  chunk, i.e. in ascending left boundary as in smax_run.
*/
#define _FILE_OFFSET_BITS 64
#include <errno.h>
#include <fcntl.h>
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

static int tab_read(const Tabfile *t, void *dst, uint64_t offset, uint64_t bytes, char *err,
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

int smax_run_stream(smax_index *idx, const smax_opts *opts, uint64_t chunk, smax_emit_cb cb,
                    void *info, char *err, size_t errlen)
{
  Tabfile lcpf = { -1, 0, "" }, bwtf = { -1, 0, "" }, llvf = { -1, 0, "" }, suff = { -1, 0, "" };
  smax_device *dev[STREAM_DEVS];
  smax_shard_view views[STREAM_DEVS], left[STREAM_DEVS];
  uint8_t *lcp = NULL, *bwt = NULL;
  smax_llv *llv = NULL;
  smax_record *recs = NULL;
  uint64_t *pos = NULL;
  uint64_t n, L, llvcap = 0, reccap = 0, poscap = 0, lo, minlength, c = 0;
  uint64_t kcur = 0;      /* .llv cursor: first record with position >= the window start */
  unsigned sufbytes = 0;
  int want_pos, rc = -1, g, navail;

  if (idx == NULL || opts == NULL || cb == NULL)
    return smax_fail(err, errlen, "smax_run_stream: null argument");
  if (idx->indexname == NULL)
    return smax_fail(err, errlen, "smax_run_stream: the index has no files behind it");
  memset(dev, 0, sizeof dev);
  n = idx->info.numberofallsortedsuffixes;
  L = idx->info.largelcpvalues;
  minlength = opts->minlength ? opts->minlength : 1;
  want_pos = opts->format != SMAX_FORMAT_ITV;
  if (chunk == 0)
    chunk = STREAM_DEFAULT_CHUNK;
  chunk = (chunk + 15) & ~(uint64_t) 15;
  if (chunk < 1024)
    chunk = 1024;           /* the left halo of a chunk must lie in its neighbour */
  if (chunk > ((uint64_t) 1 << 32))
    chunk = (uint64_t) 1 << 32;
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
  {
    const uint64_t span = (chunk < n ? chunk : n) + STREAM_HALO + 64;
    lcp = malloc(span);
    bwt = malloc(span);
    if (lcp == NULL || bwt == NULL)
    {
      smax_fail(err, errlen, "out of memory for a chunk of %lu suffixes", (unsigned long) chunk);
      goto done;
    }
  }
  for (g = 0; g < STREAM_DEVS; g++)
    if (smax_device_create(opts->first_device, &dev[g], err, errlen) != 0)
      goto done;

  for (lo = 0; lo < n; lo += chunk, c++)
  {
    const uint64_t hi = n - lo < chunk ? n : lo + chunk;
    const uint64_t w_lo = lo >= STREAM_HALO ? lo - STREAM_HALO : 0;
    const uint64_t w_hi = hi + 16 < n ? hi + 16 : n;
    uint64_t k0, k1, nrecs = 0, r, npos = 0;
    smax_index *win = NULL;
    smax_device *d = dev[c % STREAM_DEVS];
    int nleft = 0, failed = 1;

    if (tab_read(&lcpf, lcp, w_lo, w_hi - w_lo, err, errlen) != 0 ||
        tab_read(&bwtf, bwt, w_lo, w_hi - w_lo, err, errlen) != 0)
      goto done;
    /* .llv records of the window: positions ascend, the cursor only moves
       forward (sequential reader of sarr-def.h:128-178); the number of
       records equals the number of 255 entries of the window's lcp bytes */
    {
      uint64_t cnt255 = 0, i;
      for (i = 0; i < w_hi - w_lo; i++)
        cnt255 += lcp[i] == 255;
      if (cnt255 + 1 > llvcap)
      {
        smax_llv *p = realloc(llv, (cnt255 + 1 + cnt255 / 4) * sizeof *llv);
        if (p == NULL)
        {
          smax_fail(err, errlen, "out of memory");
          goto done;
        }
        llv = p;
        llvcap = cnt255 + 1 + cnt255 / 4;
      }
      /* advance the cursor to the first record at or right of the window */
      k0 = kcur;
      while (k0 < L)
      {
        smax_llv one;
        if (tab_read(&llvf, &one, k0 * sizeof one, sizeof one, err, errlen) != 0)
          goto done;
        if (one.position >= w_lo)
          break;
        k0++;
      }
      k1 = k0 + cnt255;
      if (k1 > L)
      {
        smax_fail(err, errlen, "inconsistent ESA tables: the lcp table holds more 255 entries "
                  "than the .llv file has records");
        goto done;
      }
      if (cnt255 > 0 && tab_read(&llvf, llv, k0 * sizeof *llv, cnt255 * sizeof *llv, err,
                                 errlen) != 0)
        goto done;
      if (cnt255 > 0 && (llv[0].position < w_lo || llv[cnt255 - 1].position >= w_hi))
      {
        smax_fail(err, errlen, "inconsistent ESA tables: .llv records do not match the 255 "
                  "entries of the lcp table");
        goto done;
      }
      /* where the next window starts in record space: skip the records left
         of its first entry (the windows overlap by the halo) */
      {
        const uint64_t next_lo = lo + chunk >= STREAM_HALO ? lo + chunk - STREAM_HALO : 0;
        uint64_t skip = 0;
        for (i = 0; i < w_hi - w_lo && w_lo + i < next_lo; i++)
          skip += lcp[i] == 255;
        kcur = k0 + skip;
      }
    }
    if (smax_index_from_memory_window(lcp, bwt, llv, k1 - k0, NULL, 8, w_lo, w_hi - w_lo, n,
                                      &win, err, errlen) != 0)
      goto done;
    if (smax_device_upload(d, win, lo, hi, 0, NULL, err, errlen) == 0)
    {
      smax_device_view(d, &views[c % STREAM_DEVS]);
      /* the two chunks before this one stay resident: left views, nearest last */
      if (c >= 2) left[nleft++] = views[(c - 2) % STREAM_DEVS];
      if (c >= 1) left[nleft++] = views[(c - 1) % STREAM_DEVS];
      if ((c == 0 || smax_device_set_left_views(d, left, nleft, err, errlen) == 0) &&
          smax_scan_launch(d, minlength, opts->policy, 0, NULL, err, errlen) == 0 &&
          smax_scan_counts(d, &nrecs, NULL, err, errlen) == 0)
        failed = 0;
    }
    smax_index_close(win);
    if (failed)
      goto done;
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
    for (r = 0; r < nrecs; r++)
      if (recs[r].width > npos)
        npos = recs[r].width;
    if (want_pos && npos > poscap)
    {
      uint64_t *p = realloc(pos, npos * sizeof *pos);
      if (p == NULL)
      {
        smax_fail(err, errlen, "out of memory");
        goto done;
      }
      pos = p;
      poscap = npos;
    }
    for (r = 0; r < nrecs; r++)
    {
      if (want_pos && suf_read(&suff, sufbytes, recs[r].lb, recs[r].width, pos, err, errlen) != 0)
        goto done;
      if (cb(info, recs[r].len, recs[r].lb, recs[r].width, want_pos ? pos : NULL) != 0)
      {
        smax_fail(err, errlen, "result callback failed");
        goto done;
      }
    }
  }
  rc = 0;
done:
  for (g = 0; g < STREAM_DEVS; g++)
    smax_device_destroy(dev[g]);
  free(lcp); free(bwt); free(llv); free(recs); free(pos);
  tab_close(&lcpf); tab_close(&bwtf); tab_close(&llvf); tab_close(&suff);
  return rc;
}
