/*
  smax_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the
  supermaximal-repeat path over a GenomeTools enhanced suffix array.

  Nothing in the product (genometools_smax_b200/, include/) may include, link
  or call this file.  Only tests/, __graft_entry__.smoke() and bench.py's
  cpu_baseline / --impl reference legs use it, and only as the checker or the
  reported CPU baseline.

  PARITY STATUS: the smax sources themselves (esa-smax.c, esa_linsmax.c,
  gt_smax.c) are NOT in /root/reference, so the
  tool's text format and option spellings are "parity unpinned".  What IS
  pinned, and what this file restates, is everything the path is made of in
  the mounted reference:
    * table semantics: lcp[i] byte with 255 = overflow resolved through the
      position-sorted .llv records   (src/match/lcpoverflow.h:24-30,
      src/match/esa-seqread.h:106-159, src/match/sfx-lcpvalues.c:371-433)
    * bwt[i] = left character of suffix suf[i], 254 for suf[i]==0
      (src/match/sfx-run.c:174-211), specials >= 254 (src/core/chardef.h:34-65)
    * lcp-interval enumeration = the bottom-up stack sweep
      (src/match/esa-bottomup.c:116-273); a reported repeat is a popped
      interval that never received a branching edge
    * left-diversity convention: two occurrences differ on the left if their
      left characters differ OR either is special
      (src/match/esa-maxpairs.c:24-31, :205-220)
    * threshold: depth >= minlength (src/match/esa-maxpairs.c:200-204)
  The restatement is pinned against reference CODE executed here:
  oracle/_ref/gtref smax-bu / smax-lin run gt_esa_bottomup and the reader
  macros of the reference itself (tests/test_oracle_vs_ref.py), against the
  committed golden vectors in tests/golden/, and against the known-answer
  vector of SURVEY.md section B.

  Two independent algorithms are given so they can be checked against each
  other:  smax_oracle_linear (run/plateau scan, the shape of esa_linsmax) and
  smax_oracle_stack (explicit lcp-interval stack, the shape of esa-smax).
*/
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct
{
  uint64_t len, lb, width;
} SmaxOracleRecord;

typedef struct
{
  uint64_t position, value;   /* src/match/lcpoverflow.h:26-30 */
} OracleLlv;

typedef struct
{
  SmaxOracleRecord *recs;
  uint64_t n, allocated;
} Recvec;

static int recvec_push(Recvec *rv, uint64_t len, uint64_t lb, uint64_t width)
{
  if (rv->n == rv->allocated)
  {
    uint64_t na = rv->allocated ? 2 * rv->allocated : 1024;
    SmaxOracleRecord *p = realloc(rv->recs, na * sizeof *p);
    if (p == NULL)
      return -1;
    rv->recs = p;
    rv->allocated = na;
  }
  rv->recs[rv->n].len = len;
  rv->recs[rv->n].lb = lb;
  rv->recs[rv->n].width = width;
  rv->n++;
  return 0;
}

/* policy 0: GenomeTools convention -- every special left character (>= 254,
             which includes the virtual character left of text position 0) is
             different from everything (esa-maxpairs.c:27-31).
   policy 1: plain one-bit-per-code mask: 254 and 255 are ordinary codes. */
static int left_distinct(const uint8_t *bwt, uint64_t lb, uint64_t rb,
                         int policy)
{
  uint8_t seen[256];
  uint64_t k;
  memset(seen, 0, sizeof seen);
  for (k = lb; k <= rb; k++)
  {
    uint8_t c = bwt[k];
    if (policy == 0 && c >= 254)
      continue;
    if (seen[c])
      return 0;
    seen[c] = 1;
  }
  return 1;
}

/* Sequential cursor over the resolved lcp values, as the reference's
   SSAR_NEXTSEQUENTIALLCPTABVALUE does (esa-seqread.h:106-159): the k-th
   255 byte takes the value of the k-th .llv record. */
typedef struct
{
  const uint8_t *lcp;
  const OracleLlv *llv;
  uint64_t nllv, nextllv;
  int bad;
} LcpCursor;

static uint64_t cursor_value(LcpCursor *c, uint64_t i)
{
  uint8_t b = c->lcp[i];
  if (b < 255)
    return b;
  if (c->nextllv >= c->nllv || c->llv[c->nextllv].position != i)
  {
    c->bad = 1;
    return 255;
  }
  return c->llv[c->nextllv++].value;
}

/*
  Linear plateau scan.  L[0..n-1] resolved lcp values, virtual L[n] = 0.
  A repeat is a maximal run L[s..e] == v with L[s-1] < v > L[e+1], v >=
  minlength, SA interval [s-1, e], whose bwt[s-1..e] are pairwise distinct.
  Returns the number of repeats (>= 0) or -1 (allocation) / -2 (corrupt llv).
*/
int64_t smax_oracle_linear(const uint8_t *lcp, uint64_t n,
                           const OracleLlv *llv, uint64_t nllv,
                           const uint8_t *bwt, uint64_t minlength, int policy,
                           SmaxOracleRecord **out)
{
  Recvec rv = {NULL, 0, 0};
  LcpCursor cur = {lcp, llv, nllv, 0, 0};
  uint64_t i, prev = 0, runstart = 0;
  int rise = 0;

  if (minlength == 0)
    minlength = 1;
  for (i = 1; i <= n; i++)
  {
    uint64_t v = (i < n) ? cursor_value(&cur, i) : 0;
    if (v != prev)
    {
      /* the run of value prev covers lcp indices [runstart, i-1] */
      if (rise && v < prev && prev >= minlength &&
          left_distinct(bwt, runstart - 1, i - 1, policy))
      {
        if (recvec_push(&rv, prev, runstart - 1, i - runstart + 1) != 0)
        {
          free(rv.recs);
          return -1;
        }
      }
      rise = (v > prev);
      runstart = i;
      prev = v;
    }
  }
  if (cur.bad)
  {
    free(rv.recs);
    return -2;
  }
  *out = rv.recs;
  return (int64_t) rv.n;
}

/*
  Bottom-up sweep with an explicit stack of open lcp-intervals, restating the
  control flow of gt_esa_bottomup (esa-bottomup.c:130-245): iteration idx
  reads lcpvalue = L[idx+1]; intervals with lcp > lcpvalue are popped with
  rb = idx; an interval is a leaf of the lcp-interval tree iff it never became
  the father of a popped interval.  The sweep runs over all n-1 boundaries
  plus the sentinel, which is equivalent to the reference's idx < nonspecials
  bound because lcp[i] == 0 for i >= nonspecials (sfx-lcpvalues.c:435-451).
*/
typedef struct
{
  uint64_t lcp, lb;
  int haschild;
} StackItv;

int64_t smax_oracle_stack(const uint8_t *lcp, uint64_t n,
                          const OracleLlv *llv, uint64_t nllv,
                          const uint8_t *bwt, uint64_t minlength, int policy,
                          SmaxOracleRecord **out)
{
  Recvec rv = {NULL, 0, 0};
  LcpCursor cur = {lcp, llv, nllv, 0, 0};
  StackItv *stack;
  uint64_t top = 0, allocated = 64, idx;

  if (minlength == 0)
    minlength = 1;
  stack = malloc(allocated * sizeof *stack);
  if (stack == NULL)
    return -1;
  stack[0].lcp = 0; stack[0].lb = 0; stack[0].haschild = 0;
  for (idx = 0; idx < n; idx++)
  {
    uint64_t lcpvalue = (idx + 1 < n) ? cursor_value(&cur, idx + 1) : 0;
    uint64_t lastlb = idx;
    int popped = 0;
    while (lcpvalue < stack[top].lcp)
    {
      StackItv itv = stack[top--];
      uint64_t rb = idx;
      if (!itv.haschild && itv.lcp >= minlength &&
          left_distinct(bwt, itv.lb, rb, policy))
      {
        if (recvec_push(&rv, itv.lcp, itv.lb, rb - itv.lb + 1) != 0)
        {
          free(stack); free(rv.recs);
          return -1;
        }
      }
      lastlb = itv.lb;
      popped = 1;
      if (lcpvalue <= stack[top].lcp)
      {
        stack[top].haschild = 1;   /* branching edge father <- popped */
        popped = 0;
      }
    }
    if (lcpvalue > stack[top].lcp)
    {
      if (top + 1 == allocated)
      {
        StackItv *p;
        allocated *= 2;
        p = realloc(stack, allocated * sizeof *stack);
        if (p == NULL)
        {
          free(stack); free(rv.recs);
          return -1;
        }
        stack = p;
      }
      top++;
      stack[top].lcp = lcpvalue;
      stack[top].lb = popped ? lastlb : idx;
      stack[top].haschild = popped;  /* pushed on top of a popped child */
    }
  }
  free(stack);
  if (cur.bad)
  {
    free(rv.recs);
    return -2;
  }
  *out = rv.recs;
  return (int64_t) rv.n;
}

void smax_oracle_free(void *p)
{
  free(p);
}

/* Ragged gather of the occurrence positions suf[lb..lb+width) in SA order;
   sufbytes is 8 (default suffixerator output) or 4 (-suftabuint,
   src/match/sfx-suffixgetset.c:48-55, :467-482). */
void smax_oracle_positions(const void *suftab, int sufbytes,
                           const SmaxOracleRecord *recs, uint64_t nrecs,
                           uint64_t *positions)
{
  uint64_t r, k, o = 0;
  for (r = 0; r < nrecs; r++)
  {
    for (k = 0; k < recs[r].width; k++)
    {
      uint64_t i = recs[r].lb + k;
      positions[o++] = (sufbytes == 8) ? ((const uint64_t*) suftab)[i]
                                       : ((const uint32_t*) suftab)[i];
    }
  }
}
