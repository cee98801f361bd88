/*
  smax_emit.c -- result emitter of libsmax (host side, C): turns one
  supermaximal repeat (len, lb, width, positions) into text.

  Conventions taken from the reference's emitters: space separated unsigned
  decimals ("%lu", /root/reference/src/core/types_api.h:53-54,75), one record
  per line on stdout (/root/reference/src/match/esa-lcpintervals.c:183-189,
  /root/reference/src/match/querymatch.c:169-187).  The grammar of the absent
  reference smax tool is unpinned, therefore the format
  is ONE switch (opts->format / opts->relative) and nothing else in the
  library depends on it:

    SMAX_FORMAT_SMAX   abs:  <len> <count> <pos_1> ... <pos_count>
                       rel:  <len> <count> <seq_1> <rel_1> ... <seq_c> <rel_c>
    SMAX_FORMAT_ITV          <len> <lb> <rb>
    SMAX_FORMAT_PAIRS  abs:  <len> <pos_i> F <len> <pos_j>           (all i < j)
                       rel:  <len> <seq_i> <rel_i> F <len> <seq_j> <rel_j>
  Positions are printed in suffix-array order.
*/
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "smax_host.h"

#define EMIT_BUF (1u << 20)

struct smax_emitter
{
  smax_index *idx;     /* for relative positions */
  FILE *fp;
  char *mem;           /* clones that render into memory (smax_emitter_clone_mem): the stream's buffer */
  size_t memlen;
  int format, relative;
  char *buf;
  size_t fill;
  int failed;
  char err[256];
};

int smax_emitter_new(const smax_index *idx, const smax_opts *opts, void *file,
                     smax_emitter **out, char *err, size_t errlen)
{
  smax_emitter *em;
  if (opts->format < SMAX_FORMAT_SMAX || opts->format > SMAX_FORMAT_PAIRS)
    return smax_fail(err, errlen, "unknown output format %d", opts->format);
  em = calloc(1, sizeof *em);
  if (em == NULL || (em->buf = malloc(EMIT_BUF)) == NULL)
  {
    free(em);
    return smax_fail(err, errlen, "out of memory");
  }
  em->idx = (smax_index *) idx;
  em->fp = file != NULL ? (FILE *) file : stdout;
  em->format = opts->format;
  em->relative = opts->relative;
  *out = em;
  return 0;
}

static void em_flush(smax_emitter *em)
{
  if (em->fill > 0 && fwrite(em->buf, 1, em->fill, em->fp) != em->fill)
    em->failed = 1;
  em->fill = 0;
}

static inline void em_reserve(smax_emitter *em, size_t need)
{
  if (em->fill + need > EMIT_BUF)
    em_flush(em);
}

/* decimal rendering, two digits per division (what printf("%lu") prints, types_api.h:53-54) */
static const char em_pairs[201] =
  "00010203040506070809101112131415161718192021222324252627282930313233343536373839"
  "40414243444546474849505152535455565758596061626364656667686970717273747576777879"
  "8081828384858687888990919293949596979899";

static inline void em_u64(smax_emitter *em, uint64_t v)
{
  char tmp[24];
  int k = 24;
  while (v >= 100)
  {
    const unsigned d = (unsigned) (v % 100);
    v /= 100;
    tmp[--k] = em_pairs[2 * d + 1];
    tmp[--k] = em_pairs[2 * d];
  }
  if (v >= 10)
  {
    tmp[--k] = em_pairs[2 * v + 1];
    tmp[--k] = em_pairs[2 * v];
  } else
    tmp[--k] = (char) ('0' + v);
  memcpy(em->buf + em->fill, tmp + k, (size_t) (24 - k));
  em->fill += (size_t) (24 - k);
}

static inline void em_ch(smax_emitter *em, char c)
{
  em->buf[em->fill++] = c;
}

static int em_pos(smax_emitter *em, uint64_t pos)
{
  em_reserve(em, 64);
  if (!em->relative)
  {
    em_u64(em, pos);
    return 0;
  } else
  {
    uint64_t seqnum, relpos;
    if (smax_index_seqnum_relpos(em->idx, pos, &seqnum, &relpos, em->err, sizeof em->err) != 0)
      return -1;
    em_u64(em, seqnum);
    em_ch(em, ' ');
    em_u64(em, relpos);
    return 0;
  }
}

int smax_emitter_emit(void *emitter, uint64_t len, uint64_t lb, uint64_t width,
                      const uint64_t *positions)
{
  smax_emitter *em = emitter;
  uint64_t i, j;

  em_reserve(em, 128);
  switch (em->format)
  {
    case SMAX_FORMAT_ITV:
      em_u64(em, len); em_ch(em, ' ');
      em_u64(em, lb); em_ch(em, ' ');
      em_u64(em, lb + width - 1); em_ch(em, '\n');
      break;
    case SMAX_FORMAT_SMAX:
      if (positions == NULL)
        return -1;
      em_u64(em, len); em_ch(em, ' ');
      em_u64(em, width);
      for (i = 0; i < width; i++)
      {
        em_reserve(em, 8);
        em_ch(em, ' ');
        if (em_pos(em, positions[i]) != 0)
          return -1;
      }
      em_reserve(em, 8);
      em_ch(em, '\n');
      break;
    case SMAX_FORMAT_PAIRS:
      if (positions == NULL)
        return -1;
      for (i = 0; i < width; i++)
      {
        for (j = i + 1; j < width; j++)
        {
          em_reserve(em, 128);
          em_u64(em, len); em_ch(em, ' ');
          if (em_pos(em, positions[i]) != 0) return -1;
          em_ch(em, ' '); em_ch(em, 'F'); em_ch(em, ' ');
          em_u64(em, len); em_ch(em, ' ');
          if (em_pos(em, positions[j]) != 0) return -1;
          em_ch(em, '\n');
        }
      }
      break;
    default:
      return -1;
  }
  return em->failed ? -1 : 0;
}

int smax_emitter_emit_records(smax_emitter *em, const smax_record *recs, uint64_t nrecs,
                              const uint64_t *positions)
{
  uint64_t r, o = 0;
  for (r = 0; r < nrecs; r++)
  {
    if (smax_emitter_emit(em, recs[r].len, recs[r].lb, recs[r].width,
                          positions != NULL ? positions + o : NULL) != 0)
      return -1;
    o += recs[r].width;
  }
  return 0;
}

/* ---- internal (smax_host.h): the shards of a run are rendered by several threads at once,
   each into a memory buffer through a clone of the caller's emitter (same format, same code path),
   and the buffers are written through the caller's emitter in shard order */
int smax_emitter_is_relative(const smax_emitter *em)
{
  return em->relative;
}

int smax_emitter_wants_positions(const smax_emitter *em)
{
  return em->format != SMAX_FORMAT_ITV;
}

int smax_emitter_clone_mem(const smax_emitter *em, smax_emitter **out, char *err, size_t errlen)
{
  smax_emitter *c = calloc(1, sizeof *c);
  if (c == NULL || (c->buf = malloc(EMIT_BUF)) == NULL)
  {
    free(c);
    return smax_fail(err, errlen, "out of memory");
  }
  c->idx = em->idx;
  c->format = em->format;
  c->relative = em->relative;
  c->fp = open_memstream(&c->mem, &c->memlen);
  if (c->fp == NULL)
  {
    free(c->buf);
    free(c);
    return smax_fail(err, errlen, "cannot open a memory stream");
  }
  *out = c;
  return 0;
}

/* ends a clone: its text (malloc'ed, the caller frees it) and the number of bytes */
int smax_emitter_finish_mem(smax_emitter *c, char **text, size_t *len)
{
  int rc;
  em_flush(c);
  if (fclose(c->fp) != 0)
    c->failed = 1;
  rc = c->failed ? -1 : 0;
  *text = c->mem;
  *len = c->memlen;
  free(c->buf);
  free(c);
  return rc;
}

int smax_emitter_write_raw(smax_emitter *em, const char *text, size_t len)
{
  em_flush(em);
  if (len > 0 && fwrite(text, 1, len, em->fp) != len)
    em->failed = 1;
  return em->failed ? -1 : 0;
}

int smax_emitter_delete(smax_emitter *em)
{
  int rc;
  if (em == NULL)
    return 0;
  em_flush(em);
  if (fflush(em->fp) != 0)
    em->failed = 1;
  rc = em->failed ? -1 : 0;
  free(em->buf);
  free(em);
  return rc;
}
