/*
  smax_index.c -- ESA table loader of libsmax (host side, C).

  Reads an unmodified `gt suffixerator` index: parses <idx>.prj, validates it
  the way the reference's loader does and maps <idx>.lcp/.llv/.bwt/.suf with
  an exact size check.  Mirrors, in behaviour and error text,
    inputsuffixarray / gt_mapsuffixarray   /root/reference/src/match/esa-map.c:296-517
    scanprjfileuintkeysviafileptr          /root/reference/src/match/esa-map.c:55-211
    gt_fa_mmap_check_size_with_suffix      /root/reference/src/core/fa.c:700-741
  Differences, on purpose: a 4-byte suffix table (-suftabuint,
  /root/reference/src/match/sfx-suffixgetset.c:48-55) is accepted in mapped
  mode too (the reference only reads it with -scan, esa-map.c:351-377), and
  of the .esq only the header is decoded (alphabet type/size; field order
  /root/reference/src/core/encseq.c:1222-1241) because the left characters
  come from the .bwt table.
*/
#define _GNU_SOURCE
#include <errno.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "smax_host.h"

int smax_fail(char *err, size_t errlen, const char *fmt, ...)
{
  if (err != NULL && errlen > 0)
  {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, errlen, fmt, ap);
    va_end(ap);
  }
  return -1;
}

/* ------------------------------- .prj -------------------------------- */
typedef struct
{
  const char *key;
  int found, optional;
  uint64_t value;
} Prjkey;

enum
{
  K_TOTALLENGTH, K_SPECIALCHARACTERS, K_SPECIALRANGES, K_REALSPECIALRANGES,
  K_LENGTHOFSPECIALPREFIX, K_LENGTHOFSPECIALSUFFIX, K_WILDCARDS,
  K_WILDCARDRANGES, K_REALWILDCARDRANGES, K_LENGTHOFWILDCARDPREFIX,
  K_LENGTHOFWILDCARDSUFFIX, K_NUMOFSEQUENCES, K_NUMOFDBSEQUENCES,
  K_NUMOFQUERYSEQUENCES, K_NUMBEROFALLSORTEDSUFFIXES, K_LONGEST,
  K_PREFIXLENGTH, K_LARGELCPVALUES, K_AVERAGELCP, K_MAXBRANCHDEPTH,
  K_INTEGERSIZE, K_LITTLEENDIAN, K_READMODE, K_MIRRORED, K_NUMKEYS
};

static int parse_prj(const char *indexname, smax_index_info *info,
                     int *has_longest, int *has_llv, char *err, size_t errlen)
{
  /* key set and optionality as registered in esa-map.c:78-122: longest,
     largelcpvalues, averagelcp and maxbranchdepth carry a "defined" flag */
  Prjkey keys[K_NUMKEYS] = {
    {"totallength", 0, 0, 0}, {"specialcharacters", 0, 0, 0},
    {"specialranges", 0, 0, 0}, {"realspecialranges", 0, 0, 0},
    {"lengthofspecialprefix", 0, 0, 0}, {"lengthofspecialsuffix", 0, 0, 0},
    {"wildcards", 0, 0, 0}, {"wildcardranges", 0, 0, 0},
    {"realwildcardranges", 0, 0, 0}, {"lengthofwildcardprefix", 0, 0, 0},
    {"lengthofwildcardsuffix", 0, 0, 0}, {"numofsequences", 0, 0, 0},
    {"numofdbsequences", 0, 0, 0}, {"numofquerysequences", 0, 0, 0},
    {"numberofallsortedsuffixes", 0, 0, 0}, {"longest", 0, 1, 0},
    {"prefixlength", 0, 0, 0}, {"largelcpvalues", 0, 1, 0},
    {"averagelcp", 0, 1, 0}, {"maxbranchdepth", 0, 1, 0},
    {"integersize", 0, 0, 0}, {"littleendian", 0, 0, 0},
    {"readmode", 0, 0, 0}, {"mirrored", 0, 0, 0}
  };
  char path[4096], line[8192];
  unsigned int linenum = 0;
  FILE *fp;
  int k;

  snprintf(path, sizeof path, "%s.prj", indexname);
  fp = fopen(path, "rb");
  if (fp == NULL)
    return smax_fail(err, errlen, "fopen(): cannot open file '%s': %s", path,
                     strerror(errno));
  for (; fgets(line, (int) sizeof line, fp) != NULL; linenum++)
  {
    char *eq;
    size_t len = strlen(line);
    while (len > 0 && (line[len - 1] == '\n' || line[len - 1] == '\r'))
      line[--len] = '\0';
    if (strncmp(line, "dbfile=", 7) == 0)
      continue;
    eq = strchr(line, '=');
    if (eq == NULL)
    {
      fclose(fp);
      return smax_fail(err, errlen, "missing equality symbol in \"%s\"", line);
    }
    for (k = 0; k < K_NUMKEYS; k++)
      if (strlen(keys[k].key) == (size_t) (eq - line) &&
          strncmp(keys[k].key, line, (size_t) (eq - line)) == 0)
        break;
    if (k == K_NUMKEYS)
    {
      fclose(fp);
      return smax_fail(err, errlen, "file %s.prj, line %u: cannot find key for \"%s\"",
                       indexname, linenum, line);
    }
    if (k == K_AVERAGELCP)
    {
      double dv;
      if (sscanf(eq + 1, "%lf", &dv) != 1)
      {
        fclose(fp);
        return smax_fail(err, errlen, "cannot find floating point number in \"%s\"", eq + 1);
      }
    } else
    {
      char *end;
      errno = 0;
      keys[k].value = strtoull(eq + 1, &end, 10);
      if (end == eq + 1 || errno != 0 || eq[1] == '-')
      {
        fclose(fp);
        return smax_fail(err, errlen, "cannot find non-negative integer in \"%s\"", eq + 1);
      }
    }
    keys[k].found = 1;
  }
  fclose(fp);
  for (k = 0; k < K_NUMKEYS; k++)
    if (!keys[k].found && !keys[k].optional)
      return smax_fail(err, errlen, "file %s.prj: missing line beginning with \"%s=\"",
                       indexname, keys[k].key);
  if (keys[K_INTEGERSIZE].value != 32 && keys[K_INTEGERSIZE].value != 64)
    return smax_fail(err, errlen, "%s.prj contains illegal line defining the integer size",
                     indexname);
  if (keys[K_INTEGERSIZE].value != 64)
    return smax_fail(err, errlen, "index was generated for %u-bit integers while "
                     "this program uses %u-bit integers",
                     (unsigned) keys[K_INTEGERSIZE].value, 64u);
  if (keys[K_LITTLEENDIAN].value != 1)
    return smax_fail(err, errlen, "computer has little endian byte order, while index "
                     "was built on computer with big endian byte order");
  if (keys[K_READMODE].value > 3)
    return smax_fail(err, errlen, "illegal readmode %u", (unsigned) keys[K_READMODE].value);
  if (keys[K_MIRRORED].value > 1)
    return smax_fail(err, errlen, "illegal mirroring flag: only 0(=no mirroring) and "
                     "1 (=mirroring) is supported, but read %u",
                     (unsigned) keys[K_MIRRORED].value);
  memset(info, 0, sizeof *info);
  info->totallength = keys[K_TOTALLENGTH].value;
  info->specialcharacters = keys[K_SPECIALCHARACTERS].value;
  info->numofsequences = keys[K_NUMOFSEQUENCES].value;
  info->numberofallsortedsuffixes = keys[K_NUMBEROFALLSORTEDSUFFIXES].value;
  /* esa-seqread.c:56-57 */
  info->nonspecials = info->totallength - info->specialcharacters;
  info->largelcpvalues = keys[K_LARGELCPVALUES].value;
  info->maxbranchdepth = keys[K_MAXBRANCHDEPTH].value;
  info->longest = keys[K_LONGEST].value;
  info->integersize = (uint32_t) keys[K_INTEGERSIZE].value;
  info->littleendian = (uint32_t) keys[K_LITTLEENDIAN].value;
  info->readmode = (uint32_t) keys[K_READMODE].value;
  info->mirrored = (uint32_t) keys[K_MIRRORED].value;
  *has_longest = keys[K_LONGEST].found;
  *has_llv = keys[K_LARGELCPVALUES].found;
  return 0;
}

/* ------------------------------- mmap -------------------------------- */
static void *map_file(const char *indexname, const char *suffix, size_t *len,
                      char *err, size_t errlen)
{
  char path[4096];
  struct stat sb;
  void *p;
  int fd;

  snprintf(path, sizeof path, "%s%s", indexname, suffix);
  fd = open(path, O_RDONLY);
  if (fd < 0)
  {
    smax_fail(err, errlen, "fopen(): cannot open file '%s': %s", path, strerror(errno));
    return NULL;
  }
  if (fstat(fd, &sb) != 0)
  {
    smax_fail(err, errlen, "cannot fstat() file '%s': %s", path, strerror(errno));
    close(fd);
    return NULL;
  }
  *len = (size_t) sb.st_size;
  if (sb.st_size == 0)
  {
    close(fd);
    return (void *) 1;   /* empty table: nothing to map (e.g. .llv with L = 0) */
  }
  p = mmap(NULL, (size_t) sb.st_size, PROT_READ, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED)
  {
    smax_fail(err, errlen, "cannot mmap() file '%s': %s", path, strerror(errno));
    return NULL;
  }
  (void) madvise(p, (size_t) sb.st_size, MADV_SEQUENTIAL);
  return p;
}

static int check_units(const char *indexname, const char *suffix, size_t numofbytes,
                       uint64_t expectedunits, size_t sizeofunit, char *err, size_t errlen)
{
  /* text of check_mapped_file_size, fa.c:703-722 */
  if (expectedunits != (uint64_t) (numofbytes / sizeofunit))
    return smax_fail(err, errlen, "mapping file %s%s: number of mapped units (of size %u) "
                     " = %lu != %lu = expected number of mapped units", indexname, suffix,
                     (unsigned) sizeofunit, (unsigned long) (numofbytes / sizeofunit),
                     (unsigned long) expectedunits);
  return 0;
}

/* ----------------------------- .esq header ---------------------------- */
static int read_esq_header(const char *indexname, smax_index_info *info,
                           char *err, size_t errlen)
{
  /* mapspec records, each padded to 8 bytes (mapspec.c:208-215): word 0
     is64bit, 1 version, 2 sat, 3 totallength, 4 numofdbsequences,
     5 numofdbfiles, 6 lengthofdbfilenames, 7..20 GtSpecialcharinfo,
     21 minseqlen, 22 maxseqlen, 23 alphatype, 24 lengthofalphadef,
     then alphadef[] (encseq.c:1222-1241) */
  char path[4096];
  uint64_t w[25];
  FILE *fp;

  snprintf(path, sizeof path, "%s.esq", indexname);
  fp = fopen(path, "rb");
  if (fp == NULL)
    return smax_fail(err, errlen, "fopen(): cannot open file '%s': %s", path,
                     strerror(errno));
  if (fread(w, sizeof w[0], 25, fp) != 25)
  {
    fclose(fp);
    return smax_fail(err, errlen, "file %s is too short to hold an encoded sequence header",
                     path);
  }
  info->alphatype = (uint32_t) w[23];
  if (w[23] == 0)
    info->numofchars = 4;          /* alphabet.c:63-70: a c g t */
  else if (w[23] == 1)
    info->numofchars = 20;         /* LVIFKREDAGSTNQYWPHMC */
  else
  {
    /* custom alphabet: one line per symbol class, the last one is the wildcard */
    uint64_t k, lines = 0, lena = w[24];
    char *def = malloc(lena + 1);
    if (def != NULL && fread(def, 1, lena, fp) == lena)
      for (k = 0; k < lena; k++)
        if (def[k] == '\n') lines++;
    free(def);
    info->numofchars = lines > 0 ? (uint32_t) (lines - 1) : 0;
  }
  fclose(fp);
  return 0;
}

/* ------------------------------ open/close ---------------------------- */
int smax_index_open(const char *indexname, unsigned demand, smax_index **out,
                    char *err, size_t errlen)
{
  smax_index *idx;
  int has_longest = 0, has_llv = 0;
  uint64_t n;

  if (indexname == NULL || out == NULL)
    return smax_fail(err, errlen, "smax_index_open: null argument");
  idx = calloc(1, sizeof *idx);
  if (idx == NULL)
    return smax_fail(err, errlen, "out of memory");
  idx->indexname = strdup(indexname);
  /* order of esa-map.c:296-466: encoded sequence first, then the project file */
  if ((demand & SMAX_TAB_ESQ) && read_esq_header(indexname, &idx->info, err, errlen) != 0)
    goto fail;
  {
    uint32_t alphatype = idx->info.alphatype, numofchars = idx->info.numofchars;
    if (parse_prj(indexname, &idx->info, &has_longest, &has_llv, err, errlen) != 0)
      goto fail;
    idx->info.alphatype = alphatype;
    idx->info.numofchars = numofchars;
  }
  n = idx->info.numberofallsortedsuffixes;
  if (demand & SMAX_TAB_SUF)
  {
    idx->map_suf = map_file(indexname, ".suf", &idx->len_suf, err, errlen);
    if (idx->map_suf == NULL)
      goto fail;
    if (n > 0 && idx->len_suf == n * 4)
      idx->info.sufbytes = 4;
    else
    {
      if (check_units(indexname, ".suf", idx->len_suf, n, 8, err, errlen) != 0)
        goto fail;
      idx->info.sufbytes = 8;
    }
    idx->suf = idx->len_suf ? idx->map_suf : NULL;
    if (!has_longest)   /* esa-map.c:384-388 */
    {
      smax_fail(err, errlen, "longest not defined");
      goto fail;
    }
  }
  if (demand & SMAX_TAB_LCP)
  {
    idx->map_lcp = map_file(indexname, ".lcp", &idx->len_lcp, err, errlen);
    if (idx->map_lcp == NULL)
      goto fail;
    if (check_units(indexname, ".lcp", idx->len_lcp, n, 1, err, errlen) != 0)
      goto fail;
    idx->lcp = idx->len_lcp ? idx->map_lcp : NULL;
    if (!has_llv)       /* esa-map.c:419-423 */
    {
      smax_fail(err, errlen, "numoflargelcpvalues not defined");
      goto fail;
    }
    if (idx->info.largelcpvalues > 0)
    {
      idx->map_llv = map_file(indexname, ".llv", &idx->len_llv, err, errlen);
      if (idx->map_llv == NULL)
        goto fail;
      if (check_units(indexname, ".llv", idx->len_llv, idx->info.largelcpvalues,
                      sizeof (smax_llv), err, errlen) != 0)
        goto fail;
      idx->llv = idx->map_llv;
    }
  }
  if (demand & SMAX_TAB_BWT)
  {
    idx->map_bwt = map_file(indexname, ".bwt", &idx->len_bwt, err, errlen);
    if (idx->map_bwt == NULL)
      goto fail;
    /* esa-map.c:451-455 expects totallength+1 units */
    if (check_units(indexname, ".bwt", idx->len_bwt, idx->info.totallength + 1, 1,
                    err, errlen) != 0)
      goto fail;
    idx->bwt = idx->len_bwt ? idx->map_bwt : NULL;
  }
  idx->base = 0;
  idx->len = n;
  *out = idx;
  return 0;
fail:
  smax_index_close(idx);
  return -1;
}

int smax_index_from_memory(const uint8_t *lcp, const uint8_t *bwt, const smax_llv *llv,
                           uint64_t nllv, const void *suf, unsigned sufbytes, uint64_t n,
                           smax_index **out, char *err, size_t errlen)
{
  return smax_index_from_memory_window(lcp, bwt, llv, nllv, suf, sufbytes, 0, n, n, out,
                                       err, errlen);
}

int smax_index_from_memory_window(const uint8_t *lcp, const uint8_t *bwt, const smax_llv *llv,
                                  uint64_t nllv, const void *suf, unsigned sufbytes,
                                  uint64_t base, uint64_t len, uint64_t n,
                                  smax_index **out, char *err, size_t errlen)
{
  smax_index *idx;
  if (base + len > n)
    return smax_fail(err, errlen, "table window [%lu, %lu) exceeds the table size %lu",
                     (unsigned long) base, (unsigned long) (base + len), (unsigned long) n);
  if (lcp == NULL || bwt == NULL || out == NULL || (nllv > 0 && llv == NULL))
    return smax_fail(err, errlen, "smax_index_from_memory: null table");
  if (suf != NULL && sufbytes != 8 && sufbytes != 4)
    return smax_fail(err, errlen, "suffix table entries must be 8 or 4 bytes");
  idx = calloc(1, sizeof *idx);
  if (idx == NULL)
    return smax_fail(err, errlen, "out of memory");
  idx->lcp = lcp; idx->bwt = bwt; idx->llv = llv; idx->suf = suf;
  idx->info.numberofallsortedsuffixes = n;
  idx->info.totallength = n ? n - 1 : 0;
  idx->info.largelcpvalues = nllv;
  idx->info.integersize = 64; idx->info.littleendian = 1;
  idx->info.sufbytes = suf ? sufbytes : 0;
  idx->info.numofsequences = 1;
  idx->base = base;
  idx->len = len;
  *out = idx;
  return 0;
}

static void unmap(void *p, size_t len)
{
  if (p != NULL && p != (void *) 1 && len > 0)
    munmap(p, len);
}

void smax_index_close(smax_index *idx)
{
  if (idx == NULL)
    return;
  unmap(idx->map_lcp, idx->len_lcp);
  unmap(idx->map_bwt, idx->len_bwt);
  unmap(idx->map_llv, idx->len_llv);
  unmap(idx->map_suf, idx->len_suf);
  free(idx->seps);
  free(idx->indexname);
  free(idx);
}

int smax_index_info_get(const smax_index *idx, smax_index_info *info)
{
  if (idx == NULL || info == NULL)
    return -1;
  *info = idx->info;
  return 0;
}

const uint8_t *smax_index_lcptab(const smax_index *idx) { return idx->lcp; }
const uint8_t *smax_index_bwttab(const smax_index *idx) { return idx->bwt; }
const smax_llv *smax_index_llvtab(const smax_index *idx) { return idx->llv; }
const void *smax_index_suftab(const smax_index *idx) { return idx->suf; }

/* ----------------------- position -> (seqnum, relpos) ------------------ */
static int cmp_u64(const void *a, const void *b)
{
  const uint64_t x = *(const uint64_t *) a, y = *(const uint64_t *) b;
  return x < y ? -1 : (x > y ? 1 : 0);
}

/* Separator positions are recovered from the tables themselves:
   { suf[i] - 1 : bwt[i] == 255 } (bwt semantics
   /root/reference/src/match/sfx-run.c:188-207).  This also yields the virtual
   separator of a -mirrored index.  Semantics follow gt_encseq_seqnum /
   gt_encseq_seqstartpos (/root/reference/src/core/encseq.c:3815-3900):
   seqnum = number of separators left of pos, relpos = pos - seqstart. */
static int build_seps(smax_index *idx, char *err, size_t errlen)
{
  const uint64_t n = idx->info.numberofallsortedsuffixes;
  uint64_t i, cnt = 0, cap;
  if (idx->seps_ready)
    return 0;
  if (idx->info.numofsequences <= 1)
  {
    idx->seps_ready = 1;
    return 0;
  }
  if (idx->suf == NULL || idx->bwt == NULL)
    return smax_fail(err, errlen, "relative positions need the bwt and suffix tables");
  cap = idx->info.numofsequences;
  idx->seps = malloc(cap * sizeof (uint64_t));
  if (idx->seps == NULL)
    return smax_fail(err, errlen, "out of memory");
  for (i = 0; i < n; i++)
  {
    if (idx->bwt[i] != 255)
      continue;
    if (cnt == cap)
    {
      uint64_t *p;
      cap *= 2;
      p = realloc(idx->seps, cap * sizeof (uint64_t));
      if (p == NULL)
        return smax_fail(err, errlen, "out of memory");
      idx->seps = p;
    }
    idx->seps[cnt++] = (idx->info.sufbytes == 8 ? ((const uint64_t *) idx->suf)[i]
                                                : ((const uint32_t *) idx->suf)[i]) - 1;
  }
  qsort(idx->seps, cnt, sizeof (uint64_t), cmp_u64);
  idx->nseps = cnt;
  idx->seps_ready = 1;
  return 0;
}

int smax_index_separators(smax_index *idx, const uint64_t **seps, uint64_t *nseps,
                          char *err, size_t errlen)
{
  if (build_seps(idx, err, errlen) != 0)
    return -1;
  if (seps != NULL) *seps = idx->seps;
  if (nseps != NULL) *nseps = idx->nseps;
  return 0;
}

int smax_index_seqnum_relpos(smax_index *idx, uint64_t pos, uint64_t *seqnum,
                             uint64_t *relpos, char *err, size_t errlen)
{
  uint64_t lo = 0, hi;
  if (build_seps(idx, err, errlen) != 0)
    return -1;
  hi = idx->nseps;
  while (lo < hi)      /* number of separators < pos */
  {
    const uint64_t mid = (lo + hi) / 2;
    if (idx->seps[mid] < pos) lo = mid + 1; else hi = mid;
  }
  *seqnum = lo;
  *relpos = lo == 0 ? pos : pos - (idx->seps[lo - 1] + 1);
  return 0;
}
