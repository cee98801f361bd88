/*
  smax_tool.c -- the `smax` tool of libsmax (host side, C).

  Keeps the GenomeTools tool plug-in shape: five callbacks
      arguments_new / arguments_delete / option_parser_new /
      arguments_check / runner
  (/root/reference/src/core/tool_api.h:30-70) driven in the order of
  gt_tool_run (/root/reference/src/core/tool.c:62-114); the option surface
  follows the sibling tool `gt repfind`
  (/root/reference/src/tools/gt_repfind.c:405-495: -l default 20 min 1,
  -ii mandatory, -scan, -v) plus the abs/rel switch the task names.  Parser
  behaviour and error texts restate /root/reference/src/core/option.c
  (:755 missing argument, :784 mandatory, :874 exclude, :1028 already set,
  :1158-1277 integer checks, :1427 unknown option, boolean yes/no arguments
  :1036-1056, "--opt" accepted :1021-1023); errors are reported as
  "<prog> <tool>: error: <msg>" on stderr with exit code 1
  (/root/reference/src/gt.c:48-50), -help / -version exit 0
  (/root/reference/src/core/tool.c:88-94).

  INTEGRATION.md shows the ~10-line gt_smax.c shim that registers these
  callbacks with gt_tool_new() inside a GenomeTools build.
*/
#include <errno.h>
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "smax_host.h"

#define ERRLEN 1024

/* ----------------------- a tiny GtOptionParser ------------------------ */
typedef enum { OPT_BOOL, OPT_UINT, OPT_STRING, OPT_CHOICE, OPT_HELP, OPT_VERSION } Opttype;

typedef struct
{
  const char *name, *description;
  Opttype type;
  void *value;
  unsigned int default_uint, min_uint;
  int default_bool;
  const char *default_str;
  const char **domain;      /* OPT_CHOICE, NULL terminated */
  int mandatory, is_set, hide_default;
} Option;

typedef struct
{
  const char *progname, *synopsis, *one_liner, *mail;
  Option *options;
  int noptions;
  int excl[4][2];           /* pairs of option indices that exclude each other */
  int nexcl;
} OptionParser;

enum { PARSE_OK = 0, PARSE_ERROR = -1, PARSE_REQUESTS_EXIT = 1 };

static void show_help(const OptionParser *op)
{
  size_t maxlen = 0;
  int i;
  for (i = 0; i < op->noptions; i++)
    if (strlen(op->options[i].name) > maxlen)
      maxlen = strlen(op->options[i].name);
  /* layout of show_help, option.c:383-460 */
  printf("Usage: %s %s\n", op->progname, op->synopsis);
  printf("%s\n\n", op->one_liner);
  for (i = 0; i < op->noptions; i++)
  {
    const Option *o = &op->options[i];
    {
      /* continuation lines of a description start in the description column
         (show_description, src/core/option.c) */
      const char *d = o->description;
      printf("-%s%*s ", o->name, (int) (maxlen - strlen(o->name)), "");
      for (; *d != '\0'; d++)
      {
        putchar(*d);
        if (*d == '\n')
          printf("%*s  ", (int) maxlen, "");
      }
      putchar('\n');
    }
    if (o->hide_default)
      continue;
    if (o->type == OPT_BOOL)
      printf("%*s  default: %s\n", (int) maxlen, "", o->default_bool ? "yes" : "no");
    else if (o->type == OPT_UINT)
      printf("%*s  default: %u\n", (int) maxlen, "", o->default_uint);
    else if (o->type == OPT_STRING || o->type == OPT_CHOICE)
      printf("%*s  default: %s\n", (int) maxlen, "",
             (o->default_str && o->default_str[0]) ? o->default_str : "undefined");
  }
  printf("\nReport bugs to %s.\n", op->mail);
}

static int parse_uint(unsigned int *out, const char *s)
{
  char *end;
  long v;
  errno = 0;
  v = strtol(s, &end, 10);
  if (end == s || *end != '\0' || errno != 0 || v < 0 || v > (long) UINT_MAX)
    return -1;
  *out = (unsigned int) v;
  return 0;
}

static int option_parser_parse(OptionParser *op, int *parsed_args, int argc,
                               const char **argv, char *err)
{
  int argnum, i, k;
  for (i = 0; i < op->noptions; i++)
  {
    Option *o = &op->options[i];
    o->is_set = 0;
    if (o->type == OPT_BOOL) *(int *) o->value = o->default_bool;
    else if (o->type == OPT_UINT) *(unsigned int *) o->value = o->default_uint;
    else if (o->type == OPT_STRING || o->type == OPT_CHOICE)
      *(const char **) o->value = o->default_str;
  }
  for (argnum = 1; argnum < argc; argnum++)
  {
    const char *a = argv[argnum];
    Option *o = NULL;
    if (!(a && a[0] == '-' && strlen(a) > 1) || !strcmp(a, "--"))
      break;
    for (i = 0; i < op->noptions; i++)
      if (!strcmp(a + 1 + (a[1] == '-' ? 1 : 0), op->options[i].name))
      {
        o = &op->options[i];
        break;
      }
    if (o == NULL)
    {
      snprintf(err, ERRLEN, "unknown option: %s (-help shows possible options)", a);
      return PARSE_ERROR;
    }
    if (o->is_set)
    {
      snprintf(err, ERRLEN, "option \"%s\" already set", o->name);
      return PARSE_ERROR;
    }
    o->is_set = 1;
    switch (o->type)
    {
      case OPT_HELP:
        show_help(op);
        return PARSE_REQUESTS_EXIT;
      case OPT_VERSION:
        printf("%s (libsmax, B200-native GenomeTools smax path) %s\n", op->progname,
               SMAX_VERSION);
        return PARSE_REQUESTS_EXIT;
      case OPT_BOOL:
        if (argnum + 1 < argc && argv[argnum + 1][0] != '-')
        {
          if (!strcmp(argv[argnum + 1], "yes") || !strcmp(argv[argnum + 1], "true"))
          {
            argnum++;
            *(int *) o->value = 1;
            break;
          }
          if (!strcmp(argv[argnum + 1], "no") || !strcmp(argv[argnum + 1], "false"))
          {
            argnum++;
            *(int *) o->value = 0;
            break;
          }
        }
        *(int *) o->value = 1;
        break;
      case OPT_UINT:
        if (argnum + 1 >= argc)
        {
          snprintf(err, ERRLEN, "missing argument to option \"-%s\"", o->name);
          return PARSE_ERROR;
        }
        argnum++;
        if (parse_uint((unsigned int *) o->value, argv[argnum]) != 0)
        {
          snprintf(err, ERRLEN, "argument to option \"-%s\" must be a non-negative integer <= %u",
                   o->name, UINT_MAX);
          return PARSE_ERROR;
        }
        if (*(unsigned int *) o->value < o->min_uint)
        {
          snprintf(err, ERRLEN, "argument to option \"-%s\" must be an integer >= %u",
                   o->name, o->min_uint);
          return PARSE_ERROR;
        }
        break;
      case OPT_STRING:
      case OPT_CHOICE:
        if (argnum + 1 >= argc || (argv[argnum + 1][0] == '-' && argv[argnum + 1][1] != '\0'))
        {
          snprintf(err, ERRLEN, "missing argument to option \"-%s\"", o->name);
          return PARSE_ERROR;
        }
        argnum++;
        if (o->type == OPT_CHOICE)
        {
          int ok = 0;
          for (k = 0; o->domain[k] != NULL; k++)
            if (!strcmp(argv[argnum], o->domain[k])) ok = 1;
          if (!ok)
          {
            size_t l = (size_t) snprintf(err, ERRLEN, "argument to option \"-%s\" must be one of: ",
                                         o->name);
            for (k = 0; o->domain[k] != NULL && l < ERRLEN; k++)
              l += (size_t) snprintf(err + l, ERRLEN - l, "%s%s", k ? ", " : "", o->domain[k]);
            return PARSE_ERROR;
          }
        }
        *(const char **) o->value = argv[argnum];
        break;
    }
  }
  if (argnum < argc && !strcmp(argv[argnum], "--"))
    argnum++;
  for (i = 0; i < op->noptions; i++)
    if (op->options[i].mandatory && !op->options[i].is_set)
    {
      snprintf(err, ERRLEN, "option \"-%s\" is mandatory", op->options[i].name);
      return PARSE_ERROR;
    }
  for (k = 0; k < op->nexcl; k++)
    if (op->options[op->excl[k][0]].is_set && op->options[op->excl[k][1]].is_set)
    {
      snprintf(err, ERRLEN, "option \"-%s\" and option \"-%s\" exclude each other",
               op->options[op->excl[k][0]].name, op->options[op->excl[k][1]].name);
      return PARSE_ERROR;
    }
  *parsed_args = argnum;
  return PARSE_OK;
}

/* ------------------------- the five callbacks ------------------------- */
typedef struct
{
  unsigned int minlength, gpus;
  int absolute, relative, scanfile, beverbose;
  const char *indexname, *policy, *format, *emit;
} Smaxoptions;

static void *gt_smax_arguments_new(void)
{
  return calloc(1, sizeof (Smaxoptions));
}

static void gt_smax_arguments_delete(void *tool_arguments)
{
  free(tool_arguments);
}

static const char *policy_domain[] = {"gt", "plain", NULL};
static const char *format_domain[] = {"smax", "itv", "pairs", NULL};
static const char *emit_domain[] = {"host", "device", NULL};

enum { O_L, O_ABS, O_REL, O_SCAN, O_II, O_GPUS, O_POLICY, O_FORMAT, O_EMIT, O_V, O_HELP, O_VERSION,
       O_NUM };

static OptionParser *gt_smax_option_parser_new(void *tool_arguments)
{
  Smaxoptions *a = tool_arguments;
  OptionParser *op = calloc(1, sizeof *op);
  Option *o = calloc(O_NUM, sizeof *o);
  op->progname = "gt smax";
  op->synopsis = "[options] -ii indexname";
  op->one_liner = "Compute supermaximal repeats.";
  op->mail = "<gt-users@genometools.org>";
  op->options = o;
  op->noptions = O_NUM;
  o[O_L] = (Option) {"l", "Specify minimum length of repeats", OPT_UINT, &a->minlength,
                     20U, 1U, 0, NULL, NULL, 0, 0, 0};
  o[O_ABS] = (Option) {"abs", "Report absolute positions", OPT_BOOL, &a->absolute,
                       0, 0, 1, NULL, NULL, 0, 0, 0};
  o[O_REL] = (Option) {"rel", "Report positions as sequence number and relative position",
                       OPT_BOOL, &a->relative, 0, 0, 0, NULL, NULL, 0, 0, 0};
  o[O_SCAN] = (Option) {"scan", "scan index rather than mapping it to main memory",
                        OPT_BOOL, &a->scanfile, 0, 0, 0, NULL, NULL, 0, 0, 0};
  o[O_II] = (Option) {"ii", "Specify input index", OPT_STRING, &a->indexname,
                      0, 0, 0, NULL, NULL, 1, 0, 0};
  o[O_GPUS] = (Option) {"gpus", "Specify number of GPUs the suffix array range is sharded over",
                        OPT_UINT, &a->gpus, 1U, 1U, 0, NULL, NULL, 0, 0, 0};
  o[O_POLICY] = (Option) {"policy", "Left context policy for special characters:\n"
                          "gt (specials are pairwise different) or plain",
                          OPT_CHOICE, &a->policy, 0, 0, 0, "gt", policy_domain, 0, 0, 0};
  o[O_FORMAT] = (Option) {"format", "Output format: smax, itv or pairs", OPT_CHOICE,
                          &a->format, 0, 0, 0, "smax", format_domain, 0, 0, 0};
  o[O_EMIT] = (Option) {"emit", "Where the result lines are rendered: host (positions are\n"
                        "gathered from the mapped suffix table) or device (the suffix\n"
                        "table is made resident and the text is rendered on the GPUs)",
                        OPT_CHOICE, &a->emit, 0, 0, 0, "host", emit_domain, 0, 0, 0};
  o[O_V] = (Option) {"v", "be verbose ", OPT_BOOL, &a->beverbose, 0, 0, 0, NULL, NULL, 0, 0, 0};
  o[O_HELP] = (Option) {"help", "display help and exit", OPT_HELP, NULL,
                        0, 0, 0, NULL, NULL, 0, 0, 1};
  o[O_VERSION] = (Option) {"version", "display version information and exit", OPT_VERSION,
                           NULL, 0, 0, 0, NULL, NULL, 0, 0, 1};
  op->excl[0][0] = O_ABS; op->excl[0][1] = O_REL;
  op->nexcl = 1;
  return op;
}

static void option_parser_delete(OptionParser *op)
{
  if (op != NULL)
  {
    free(op->options);
    free(op);
  }
}

static int gt_smax_arguments_check(int rest_argc, void *tool_arguments, char *err)
{
  (void) rest_argc; (void) tool_arguments; (void) err;
  return 0;
}

static int gt_smax_runner(int argc, const char **argv, int parsed_args,
                          void *tool_arguments, char *err)
{
  Smaxoptions *a = tool_arguments;
  smax_index *idx = NULL;
  smax_emitter *em = NULL;
  smax_opts opts;
  smax_index_info info;
  unsigned demand = SMAX_TAB_LCP | SMAX_TAB_BWT | SMAX_TAB_ESQ;
  int rc = 0;

  /* as gt_repfind.c:521-525 */
  if (parsed_args < argc)
  {
    snprintf(err, ERRLEN, "superfluous arguments: \"%s\"", argv[argc - 1]);
    return -1;
  }
  memset(&opts, 0, sizeof opts);
  opts.minlength = a->minlength;
  opts.relative = a->relative ? 1 : 0;
  opts.ngpus = (int) a->gpus;
  opts.policy = strcmp(a->policy, "plain") == 0 ? SMAX_POLICY_PLAIN : SMAX_POLICY_GT;
  opts.format = strcmp(a->format, "itv") == 0 ? SMAX_FORMAT_ITV
              : strcmp(a->format, "pairs") == 0 ? SMAX_FORMAT_PAIRS : SMAX_FORMAT_SMAX;
  opts.verbose = a->beverbose;
  if (opts.format != SMAX_FORMAT_ITV)
    demand |= SMAX_TAB_SUF;
  if (a->scanfile)        /* tables are streamed from the files, not mapped */
    demand = SMAX_TAB_ESQ;
  if (a->scanfile && strcmp(a->emit, "device") == 0)
  {
    snprintf(err, ERRLEN, "option \"-scan\" and option \"-emit device\" exclude each other");
    return -1;
  }
  if (strcmp(a->emit, "device") == 0 && opts.format == SMAX_FORMAT_PAIRS)
  {
    snprintf(err, ERRLEN, "option \"-emit device\" renders the formats smax and itv; "
             "use \"-emit host\" for pairs");
    return -1;
  }
  if (smax_index_open(a->indexname, demand, &idx, err, ERRLEN) != 0)
    return -1;
  smax_index_info_get(idx, &info);
  if (a->beverbose)   /* GtLogger lines: "# " prefix on stdout, gt_repfind.c:520 */
  {
    printf("# indexname=%s\n", a->indexname);
    printf("# numberofallsortedsuffixes=%lu\n", (unsigned long) info.numberofallsortedsuffixes);
    printf("# largelcpvalues=%lu\n", (unsigned long) info.largelcpvalues);
    printf("# suftab uses %ubit values\n", info.sufbytes * 8);
    printf("# minlength=%u gpus=%u policy=%s\n", a->minlength, a->gpus, a->policy);
  }
  if (strcmp(a->emit, "device") == 0)
  {
    if (smax_run_text(idx, &opts, stdout, NULL, err, ERRLEN) != 0)
      rc = -1;
    smax_index_close(idx);
    return rc;
  }
  if (smax_emitter_new(idx, &opts, stdout, &em, err, ERRLEN) != 0)
    rc = -1;
  if (rc == 0 && (a->scanfile
                  ? smax_run_stream(idx, &opts, 0, smax_emitter_emit, em, err, ERRLEN)
                  : smax_run(idx, &opts, smax_emitter_emit, em, err, ERRLEN)) != 0)
    rc = -1;
  if (smax_emitter_delete(em) != 0 && rc == 0)
  {
    snprintf(err, ERRLEN, "cannot write results");
    rc = -1;
  }
  smax_index_close(idx);
  return rc;
}

/* gt_tool_run (tool.c:62-114) over the five callbacks + main()'s error print
   (gt.c:48-50) */
int smax_tool_main(int argc, const char **argv)
{
  char err[ERRLEN];
  void *arguments;
  OptionParser *op;
  int parsed_args = 0, oprval, had_err = 0;

  err[0] = '\0';
  arguments = gt_smax_arguments_new();
  op = gt_smax_option_parser_new(arguments);
  oprval = option_parser_parse(op, &parsed_args, argc, argv, err);
  if (oprval == PARSE_ERROR)
    had_err = -1;
  else if (oprval == PARSE_REQUESTS_EXIT)
  {
    option_parser_delete(op);
    gt_smax_arguments_delete(arguments);
    return 0;
  }
  if (!had_err)
    had_err = gt_smax_arguments_check(argc - parsed_args, arguments, err);
  if (!had_err)
    had_err = gt_smax_runner(argc, argv, parsed_args, arguments, err);
  option_parser_delete(op);
  gt_smax_arguments_delete(arguments);
  if (had_err)
  {
    fprintf(stderr, "%s: error: %s\n", "gt smax", err);
    return 1;
  }
  return 0;
}
