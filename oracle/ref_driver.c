/*
  ref_driver.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).

  A tiny main() that is compiled TOGETHER WITH the unmodified reference sources
  where they lie under /root/reference (recipe: oracle/Makefile.ref, outputs
  only into oracle/_ref/).  It lets the parity tests and the bench's CPU
  baseline execute *reference code* for everything the mounted reference
  contains on the smax path:

    gtref suffixerator <args>       reference index builder
                                    (src/match/sfx-run.c:720, format authority)
    gtref smax-bu  <idx> <minlen> [scan] [rel] [quiet]
                                    supermaximal repeats as a GtESAVisitor
                                    plug-in on the reference's own sweep
                                    gt_esa_bottomup (src/match/esa-bottomup.c:116-273);
                                    stands in for the absent esa-smax.c
    gtref smax-lin <idx> <minlen> [scan] [rel] [quiet]
                                    stack-free linear scan on the reference's
                                    reader macros (src/match/esa-seqread.h:96-215);
                                    stands in for the absent esa_linsmax.c
    gtref repfind <idx> <minlen>    maximal pairs via gt_callenummaxpairs
                                    (src/match/esa-maxpairs.c:476-513), abs positions
    gtref scanesa <idx> <mode>      sequential table read floor
                                    (src/match/esa-lcpintervals.c:228-300)
    gtref lcpitvs <idx>             all lcp-intervals "N l lb rb"
                                    (src/match/esa-lcpintervals.c, gt_runenumlcpvalues)

  In both smax stand-ins the table decoding (.suf/.lcp/.llv), the sequential
  reader, the interval enumeration and the left characters
  (gt_encseq_get_encoded_char on the .esq, NOT the .bwt) are executed by
  reference code; only "leaf interval && lcp >= minlength && left characters
  pairwise distinct (specials distinct from everything,
  src/match/esa-maxpairs.c:24-31)" is ours.

  Output of smax-bu / smax-lin, one line per repeat, ascending left boundary:
      <length> <count> <pos_1> ... <pos_count>        (absolute, SA order)
  with `rel` every position is printed as "<seqnum> <relpos>" computed by the
  reference's own gt_encseq_seqnum / gt_encseq_seqstartpos
  (src/core/encseq.c:3815-3900, incl. the -mirrored arithmetic; the index must
  have been built with -ssp when it holds several sequences); `quiet` counts the
  repeats without printing them (the bench's timing of the scan alone).
  A trailing "# t_scan_s=<seconds>" line goes to stderr for the bench.
*/
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "core/init_api.h"
#include "core/error_api.h"
#include "core/logger.h"
#include "core/ma_api.h"
#include "core/encseq.h"
#include "core/chardef.h"
#include "core/class_alloc_lock.h"
#include "core/unused_api.h"
#include "match/sfx-run.h"
#include "match/esa-seqread.h"
#include "match/esa-bottomup.h"
#include "match/esa_visitor_rep.h"
#include "match/esa-maxpairs.h"
#include "match/esa-lcpintervals.h"

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

/* ---- growable position buffer + left-char distinctness bookkeeping ---- */
typedef struct
{
  GtUword *pos;
  GtUword npos, allocated;
  unsigned char seen[256];   /* non-special left chars already met */
  bool dup;
} Leftset;

static void leftset_reset(Leftset *ls)
{
  ls->npos = 0;
  ls->dup = false;
  memset(ls->seen, 0, sizeof ls->seen);
}

static void leftset_add(Leftset *ls, const GtEncseq *encseq, GtReadmode rm,
                        GtUword pos)
{
  GtUchar cc;
  if (ls->npos == ls->allocated)
  {
    ls->allocated = ls->allocated ? 2 * ls->allocated : 64;
    ls->pos = gt_realloc(ls->pos, sizeof (*ls->pos) * ls->allocated);
  }
  ls->pos[ls->npos++] = pos;
  /* left context convention of esa-maxpairs.c:205-220: position 0 has the
     "initial char", every special is distinct from everything */
  if (pos == 0)
    return;
  cc = gt_encseq_get_encoded_char(encseq, pos - 1, rm);
  if (ISSPECIAL(cc))
    return;
  if (ls->seen[cc])
    ls->dup = true;
  ls->seen[cc] = 1;
}

static bool g_rel = false, g_quiet = false;   /* command line: rel, quiet */
static const GtEncseq *g_encseq = NULL;       /* for rel */

static void emit_repeat(GtUword len, const Leftset *ls, GtUword *nout)
{
  GtUword i;
  (*nout)++;
  if (g_quiet)
    return;
  printf(GT_WU " " GT_WU, len, ls->npos);
  for (i = 0; i < ls->npos; i++)
  {
    if (g_rel)
    {
      /* the reference's own position -> (sequence number, relative position) */
      const GtUword seqnum = gt_encseq_seqnum(g_encseq, ls->pos[i]);
      printf(" " GT_WU " " GT_WU, seqnum,
             ls->pos[i] - gt_encseq_seqstartpos(g_encseq, seqnum));
    } else
      printf(" " GT_WU, ls->pos[i]);
  }
  printf("\n");
}

/* ---------------------- smax-bu: visitor plug-in ---------------------- */
typedef struct
{
  bool haschild;
} SmaxNodeinfo;

typedef struct
{
  const GtESAVisitor parent_instance;
  const GtEncseq *encseq;
  GtReadmode readmode;
  GtUword minlength, nout;
  /* leaves arrive in SA order.  A node without child interval is on top of
     the stack from its push to its pop, so its leaves are exactly the last
     (rb-lb+1) leaf events before the pop; after ANY pop every later
     childless node starts further right, so the log can be cleared. */
  GtUword *leaflog;
  GtUword nleaf, leafalloc;
  Leftset ls;
} SmaxVisitor;

static const GtESAVisitorClass *smax_visitor_class(void);
#define smax_visitor_cast(GV) gt_esa_visitor_cast(smax_visitor_class(), GV)

static GtESAVisitorInfo *smax_info_new(GT_UNUSED GtESAVisitor *ev)
{
  SmaxNodeinfo *ni = gt_malloc(sizeof *ni);
  ni->haschild = false;
  return (GtESAVisitorInfo*) ni;
}

static void smax_info_delete(GtESAVisitorInfo *info, GT_UNUSED GtESAVisitor *ev)
{
  gt_free(info);
}

static int smax_leafedge(GtESAVisitor *ev, bool firstsucc, GT_UNUSED GtUword fd,
                         GT_UNUSED GtUword flb, GtESAVisitorInfo *info,
                         GtUword leafnumber, GT_UNUSED GtError *err)
{
  SmaxVisitor *sv = smax_visitor_cast(ev);
  SmaxNodeinfo *ni = (SmaxNodeinfo*) info;
  if (sv->nleaf == sv->leafalloc)
  {
    sv->leafalloc = sv->leafalloc ? 2 * sv->leafalloc : 1024;
    sv->leaflog = gt_realloc(sv->leaflog, sizeof (GtUword) * sv->leafalloc);
  }
  if (firstsucc)   /* stack slots are recycled: (re)initialise on first edge */
    ni->haschild = false;
  sv->leaflog[sv->nleaf++] = leafnumber;
  return 0;
}

static int smax_branchingedge(GT_UNUSED GtESAVisitor *ev, GT_UNUSED bool firstsucc,
                              GT_UNUSED GtUword fd, GT_UNUSED GtUword flb,
                              GtESAVisitorInfo *finfo, GT_UNUSED GtUword sd,
                              GT_UNUSED GtUword slb, GT_UNUSED GtUword srb,
                              GT_UNUSED GtESAVisitorInfo *sinfo,
                              GT_UNUSED GtError *err)
{
  ((SmaxNodeinfo*) finfo)->haschild = true;
  return 0;
}

static int smax_lcpinterval(GtESAVisitor *ev, GtUword lcp, GtUword lb, GtUword rb,
                            GtESAVisitorInfo *info, GT_UNUSED GtError *err)
{
  SmaxVisitor *sv = smax_visitor_cast(ev);
  SmaxNodeinfo *ni = (SmaxNodeinfo*) info;
  if (!ni->haschild && lcp >= sv->minlength)
  {
    GtUword i, width = rb - lb + 1;
    if (sv->nleaf < width)
    {
      fprintf(stderr, "smax-bu: leaf bookkeeping mismatch at lb=" GT_WU "\n", lb);
      exit(3);
    }
    leftset_reset(&sv->ls);
    for (i = 0; i < width; i++)
      leftset_add(&sv->ls, sv->encseq, sv->readmode,
                  sv->leaflog[sv->nleaf - width + i]);
    if (!sv->ls.dup)
      emit_repeat(lcp, &sv->ls, &sv->nout);
  }
  sv->nleaf = 0;
  return 0;
}

static void smax_visitor_free(GtESAVisitor *ev)
{
  SmaxVisitor *sv = smax_visitor_cast(ev);
  gt_free(sv->leaflog);
  gt_free(sv->ls.pos);
}

static const GtESAVisitorClass *smax_visitor_class(void)
{
  static const GtESAVisitorClass *esc = NULL;
  gt_class_alloc_lock_enter();
  if (!esc)
    esc = gt_esa_visitor_class_new(sizeof (SmaxVisitor), smax_visitor_free,
                                   smax_leafedge, smax_branchingedge,
                                   smax_lcpinterval, smax_info_new,
                                   smax_info_delete);
  gt_class_alloc_lock_leave();
  return esc;
}

static int run_smax_bu(const char *indexname, GtUword minlength, bool scan,
                       GtError *err)
{
  Sequentialsuffixarrayreader *ssar;
  GtESAVisitor *ev;
  SmaxVisitor *sv;
  int had_err;
  double t0, t1;

  ssar = gt_newSequentialsuffixarrayreaderfromfile(indexname,
                                                   SARR_LCPTAB | SARR_SUFTAB |
                                                   SARR_ESQTAB |
                                                   (g_rel ? SARR_SSPTAB : 0),
                                                   scan, NULL, err);
  if (ssar == NULL)
    return -1;
  ev = gt_esa_visitor_create(smax_visitor_class());
  sv = smax_visitor_cast(ev);
  sv->encseq = g_encseq = gt_encseqSequentialsuffixarrayreader(ssar);
  sv->readmode = gt_readmodeSequentialsuffixarrayreader(ssar);
  sv->minlength = minlength;
  sv->nout = 0;
  sv->leaflog = NULL; sv->nleaf = sv->leafalloc = 0;
  memset(&sv->ls, 0, sizeof sv->ls);
  t0 = now_s();
  had_err = gt_esa_bottomup(ssar, ev, err);
  t1 = now_s();
  fflush(stdout);
  fprintf(stderr, "# t_scan_s=%.6f repeats=" GT_WU "\n", t1 - t0, sv->nout);
  gt_esa_visitor_delete(ev);
  gt_freeSequentialsuffixarrayreader(&ssar);
  return had_err;
}

/* ------------- smax-lin: linear scan on the reader macros ------------- */
static int run_smax_lin(const char *indexname, GtUword minlength, bool scan,
                        GtError *err)
{
  Sequentialsuffixarrayreader *ssar;
  const GtEncseq *encseq;
  GtReadmode readmode;
  GtUword idx, nonspecials, lcpvalue, suffix = 0, nout = 0;
  /* state of the currently open run of equal lcp values */
  GtUword runvalue = 0, prevlcp = 0;
  bool rise = false;
  Leftset ls;
  bool haserr = false;    /* set by the reader macro on a truncated .llv */
  double t0, t1;

  ssar = gt_newSequentialsuffixarrayreaderfromfile(indexname,
                                                   SARR_LCPTAB | SARR_SUFTAB |
                                                   SARR_ESQTAB |
                                                   (g_rel ? SARR_SSPTAB : 0),
                                                   scan, NULL, err);
  if (ssar == NULL)
    return -1;
  encseq = g_encseq = gt_encseqSequentialsuffixarrayreader(ssar);
  readmode = gt_readmodeSequentialsuffixarrayreader(ssar);
  nonspecials = gt_Sequentialsuffixarrayreader_nonspecials(ssar);
  memset(&ls, 0, sizeof ls);
  leftset_reset(&ls);
  t0 = now_s();
  /* iteration idx delivers lcp[idx+1] and suf[idx] (esa-seqread.c:48:
     nextlcptabindex starts at 1) */
  for (idx = 0; idx < nonspecials; idx++)
  {
    SSAR_NEXTSEQUENTIALLCPTABVALUE(lcpvalue, ssar);
    SSAR_NEXTSEQUENTIALSUFTABVALUE(suffix, ssar);
    /* suffix = suf[idx] closes the run that ended at lcp index idx */
    if (rise)
      leftset_add(&ls, encseq, readmode, suffix);
    if (lcpvalue != prevlcp)
    {
      /* the run of value prevlcp over lcp indices [..idx] ends here */
      if (rise && lcpvalue < prevlcp && runvalue >= minlength && !ls.dup)
        emit_repeat(runvalue, &ls, &nout);
      if (lcpvalue > prevlcp)
      {
        /* a new run starts at lcp index idx+1 with a rise:
           SA interval starts at idx */
        rise = true;
        runvalue = lcpvalue;
        leftset_reset(&ls);
        leftset_add(&ls, encseq, readmode, suffix);
      } else
      {
        rise = false;
      }
    }
    prevlcp = lcpvalue;
  }
  t1 = now_s();
  fflush(stdout);
  fprintf(stderr, "# t_scan_s=%.6f repeats=" GT_WU "\n", t1 - t0, nout);
  gt_free(ls.pos);
  gt_freeSequentialsuffixarrayreader(&ssar);
  return haserr ? -1 : 0;
}

/* ------------------------------ repfind ------------------------------ */
static int print_maxpair(GT_UNUSED void *info, GT_UNUSED const GtGenericEncseq *ge,
                         GtUword len, GtUword pos1, GtUword pos2,
                         GT_UNUSED GtError *err)
{
  if (pos1 > pos2) { GtUword t = pos1; pos1 = pos2; pos2 = t; }
  printf(GT_WU " " GT_WU " " GT_WU "\n", len, pos1, pos2);
  return 0;
}

int main(int argc, char **argv)
{
  GtError *err;
  int rc = 0;

  if (argc < 2)
  {
    fprintf(stderr, "usage: %s suffixerator|smax-bu|smax-lin|repfind|scanesa|"
                    "lcpitvs ...\n", argv[0]);
    return 2;
  }
  gt_lib_init();
  err = gt_error_new();
  gt_error_set_progname(err, "gtref");
  if (strcmp(argv[1], "suffixerator") == 0)
  {
    rc = gt_parseargsandcallsuffixerator(true, argc - 1,
                                         (const char**) (argv + 1), err);
  } else if ((strcmp(argv[1], "smax-bu") == 0 ||
              strcmp(argv[1], "smax-lin") == 0) && argc >= 4)
  {
    GtUword minlength = strtoul(argv[3], NULL, 10);
    bool scan = false;
    int a;
    for (a = 4; a < argc; a++)
    {
      if (strcmp(argv[a], "scan") == 0) scan = true;
      else if (strcmp(argv[a], "rel") == 0) g_rel = true;
      else if (strcmp(argv[a], "quiet") == 0) g_quiet = true;
    }
    rc = (argv[1][5] == 'b' ? run_smax_bu : run_smax_lin)(argv[2], minlength,
                                                         scan, err);
  } else if (strcmp(argv[1], "repfind") == 0 && argc >= 4)
  {
    double t0 = now_s();
    rc = gt_callenummaxpairs(argv[2], (unsigned int) atoi(argv[3]), false,
                             print_maxpair, NULL, NULL, err);
    fflush(stdout);
    fprintf(stderr, "# t_total_s=%.6f\n", now_s() - t0);
  } else if (strcmp(argv[1], "scanesa") == 0 && argc >= 4)
  {
    double t0 = now_s();
    rc = gt_runscanesa(argv[2], (unsigned int) atoi(argv[3]), NULL, err);
    fflush(stdout);
    fprintf(stderr, "# t_total_s=%.6f\n", now_s() - t0);
  } else if (strcmp(argv[1], "lcpitvs") == 0 && argc >= 3)
  {
    rc = gt_runenumlcpvalues(argv[2], false, false, NULL, err);
  } else
  {
    fprintf(stderr, "gtref: bad arguments\n");
    rc = 2;
  }
  if (rc != 0 && gt_error_is_set(err))
  {
    fprintf(stderr, "gtref %s: error: %s\n", argv[1], gt_error_get(err));
    rc = 1;
  }
  gt_error_delete(err);
  (void) gt_lib_clean();
  return rc;
}
