/*
  smax_host.h -- internal declarations shared by the host-side C sources of
  libsmax (not part of the ABI; include/smax.h is).
*/
#ifndef SMAX_HOST_H
#define SMAX_HOST_H

#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include "smax.h"

#ifdef __cplusplus
extern "C" {
#endif

struct smax_index
{
  smax_index_info info;
  const uint8_t *lcp, *bwt;
  const smax_llv *llv;
  const void *suf;
  uint64_t base, len;   /* the arrays cover lcp indices [base, base+len) */
  /* mmap bookkeeping (NULL / 0 for smax_index_from_memory) */
  void *map_lcp, *map_bwt, *map_llv, *map_suf;
  size_t len_lcp, len_bwt, len_llv, len_suf;
  /* sequence separators (absolute positions, ascending), built lazily for
     relative output */
  uint64_t *seps;
  uint64_t nseps;
  int seps_ready;
  char *indexname;
};

/* writes a printf-style message into (err, errlen); always returns -1 */
int smax_fail(char *err, size_t errlen, const char *fmt, ...);

/* the handle's own (non-blocking) CUDA stream as a void*, NULL in the CPU stand-in: what
   smax_run launches its scans on, so that the upload of the next shard overlaps them */
void *smax_device_own_stream(smax_device *dev);

/* device handles kept between the runs of a process (smax_run.c): begin tells whether this run
   may use the cache (one run at a time does), acquire hands out the cached handle of slot g on
   CUDA device `ordinal` or a new one, end waits for all handles and puts them back (or destroys
   them after a failure / when the run did not own the cache) */
int smax_cache_begin(void);
int smax_cache_acquire(int g, int ordinal, int cached, smax_device **out, char *err, size_t errlen);
void smax_cache_end(smax_device **dev, const int *ordinal, int nshards, int cached, int failed);

/* the library's emitter rendered by several threads (smax_emit.c; smax_run.c uses them when the
   callback of smax_run is smax_emitter_emit) */
int smax_emitter_is_relative(const smax_emitter *em);
int smax_emitter_wants_positions(const smax_emitter *em);
int smax_emitter_clone_mem(const smax_emitter *em, smax_emitter **out, char *err, size_t errlen);
int smax_emitter_finish_mem(smax_emitter *clone, char **text, size_t *len);
int smax_emitter_write_raw(smax_emitter *em, const char *text, size_t len);

/* position -> (seqnum, relpos); builds the separator table on first use */
int smax_index_seqnum_relpos(smax_index *idx, uint64_t pos, uint64_t *seqnum,
                             uint64_t *relpos, char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif
