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

/* position -> (seqnum, relpos); builds the separator table on first use */
int smax_index_seqnum_relpos(smax_index *idx, uint64_t pos, uint64_t *seqnum,
                             uint64_t *relpos, char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif
