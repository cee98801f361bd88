/*
  gt_smax.c -- `gt smax` as a GenomeTools tool whose work is done by libsmax.so.

  This is the file a GenomeTools maintainer adds as src/tools/gt_smax.c (with
  libsmax.so on the link line and include/smax.h on the include path): the five
  GtTool callbacks (src/core/tool_api.h:30-70, driver gt_tool_run
  src/core/tool.c:62-114) on the reference's OWN GtOptionParser, written after
  the sibling tool src/tools/gt_repfind.c:405-495, :625-632.  The runner calls
  the C ABI of include/smax.h and hands its message to gt_error_set, so the
  caller's error print (src/gt.c:48-50) and exit codes are the reference's.

  tests/test_gt_shim.py compiles this file against the reference library that
  oracle/Makefile.ref builds (oracle/_ref/libgtref.a), registers the tool in a
  GtToolbox, runs gt_tool_run on the golden indexes and diffs stdout with the
  reference-run golden text.
*/
#include <stdio.h>
#include <string.h>
#include "core/error_api.h"
#include "core/ma_api.h"
#include "core/option_api.h"
#include "core/str_api.h"
#include "core/unused_api.h"
#include "core/tool_api.h"
#include "tools/gt_smax.h"
#include "smax.h"

typedef struct
{
  unsigned int minlength, gpus;
  bool absolute, relative, scanfile, beverbose;
  GtStr *indexname, *policy, *format, *emit;
} GtSmaxArguments;

static void* gt_smax_arguments_new(void)
{
  GtSmaxArguments *arguments = gt_malloc(sizeof *arguments);
  arguments->indexname = gt_str_new();
  arguments->policy = gt_str_new();
  arguments->format = gt_str_new();
  arguments->emit = gt_str_new();
  return arguments;
}

static void gt_smax_arguments_delete(void *tool_arguments)
{
  GtSmaxArguments *arguments = tool_arguments;
  if (!arguments)
    return;
  gt_str_delete(arguments->indexname);
  gt_str_delete(arguments->policy);
  gt_str_delete(arguments->format);
  gt_str_delete(arguments->emit);
  gt_free(arguments);
}

static GtOptionParser* gt_smax_option_parser_new(void *tool_arguments)
{
  static const char *policies[] = { "gt", "plain", NULL },
                    *formats[] = { "smax", "itv", "pairs", NULL },
                    *emitters[] = { "host", "device", NULL };
  GtSmaxArguments *arguments = tool_arguments;
  GtOptionParser *op;
  GtOption *option, *absoption, *reloption;

  op = gt_option_parser_new("[options] -ii indexname",
                            "Compute supermaximal repeats.");
  gt_option_parser_set_mail_address(op, "<gt-users@genometools.org>");

  /* as the sibling tool: gt_repfind.c:415-419 */
  option = gt_option_new_uint_min("l", "Specify minimum length of repeats",
                                  &arguments->minlength, 20U, 1U);
  gt_option_parser_add_option(op, option);

  absoption = gt_option_new_bool("abs", "Report absolute positions",
                                 &arguments->absolute, true);
  gt_option_parser_add_option(op, absoption);

  reloption = gt_option_new_bool("rel", "Report positions as sequence number "
                                 "and relative position",
                                 &arguments->relative, false);
  gt_option_parser_add_option(op, reloption);
  gt_option_exclude(absoption, reloption);

  /* gt_repfind.c:451-455 */
  option = gt_option_new_bool("scan", "scan index rather than mapping "
                                      "it to main memory",
                              &arguments->scanfile, false);
  gt_option_parser_add_option(op, option);

  /* gt_repfind.c:457-461 */
  option = gt_option_new_string("ii", "Specify input index",
                                arguments->indexname, NULL);
  gt_option_parser_add_option(op, option);
  gt_option_is_mandatory(option);

  option = gt_option_new_uint_min("gpus", "Specify number of GPUs the suffix "
                                  "array range is sharded over",
                                  &arguments->gpus, 1U, 1U);
  gt_option_parser_add_option(op, option);

  option = gt_option_new_choice("policy", "Left context policy for special "
                                "characters:\ngt (specials are pairwise "
                                "different) or plain",
                                arguments->policy, policies[0], policies);
  gt_option_parser_add_option(op, option);

  option = gt_option_new_choice("format", "Output format: smax, itv or pairs",
                                arguments->format, formats[0], formats);
  gt_option_parser_add_option(op, option);

  option = gt_option_new_choice("emit", "Where the result lines are rendered: "
                                "host (positions are\ngathered from the mapped "
                                "suffix table) or device (the suffix\ntable is "
                                "made resident and the text is rendered on the "
                                "GPUs)",
                                arguments->emit, emitters[0], emitters);
  gt_option_parser_add_option(op, option);

  /* gt_repfind.c:469-473 */
  option = gt_option_new_bool("v", "be verbose ", &arguments->beverbose, false);
  gt_option_parser_add_option(op, option);
  return op;
}

static int gt_smax_arguments_check(GT_UNUSED int rest_argc, void *tool_arguments,
                                   GtError *err)
{
  GtSmaxArguments *arguments = tool_arguments;
  gt_error_check(err);
  if (arguments->scanfile && strcmp(gt_str_get(arguments->emit), "device") == 0)
  {
    gt_error_set(err, "option \"-scan\" and option \"-emit device\" exclude "
                      "each other");
    return -1;
  }
  if (strcmp(gt_str_get(arguments->emit), "device") == 0 &&
      strcmp(gt_str_get(arguments->format), "pairs") == 0)
  {
    gt_error_set(err, "option \"-emit device\" renders the formats smax and "
                      "itv; use \"-emit host\" for pairs");
    return -1;
  }
  return 0;
}

static int gt_smax_runner(int argc, const char **argv, int parsed_args,
                          void *tool_arguments, GtError *err)
{
  GtSmaxArguments *arguments = tool_arguments;
  char msg[1024];
  smax_index *idx = NULL;
  smax_emitter *em = NULL;
  smax_opts opts;
  unsigned demand = SMAX_TAB_LCP | SMAX_TAB_BWT | SMAX_TAB_ESQ;
  const bool on_device = strcmp(gt_str_get(arguments->emit), "device") == 0;
  int had_err = 0;

  gt_error_check(err);
  /* as gt_repfind.c:521-525 */
  if (parsed_args < argc)
  {
    gt_error_set(err, "superfluous arguments: \"%s\"", argv[argc - 1]);
    return -1;
  }
  memset(&opts, 0, sizeof opts);
  opts.minlength = arguments->minlength;
  opts.relative = arguments->relative ? 1 : 0;
  opts.ngpus = (int) arguments->gpus;
  opts.policy = strcmp(gt_str_get(arguments->policy), "plain") == 0
                  ? SMAX_POLICY_PLAIN : SMAX_POLICY_GT;
  opts.format = strcmp(gt_str_get(arguments->format), "itv") == 0
                  ? SMAX_FORMAT_ITV
                  : strcmp(gt_str_get(arguments->format), "pairs") == 0
                      ? SMAX_FORMAT_PAIRS : SMAX_FORMAT_SMAX;
  opts.verbose = arguments->beverbose;
  if (opts.format != SMAX_FORMAT_ITV)
    demand |= SMAX_TAB_SUF;
  if (arguments->scanfile)   /* tables are streamed from the files, not mapped */
    demand = SMAX_TAB_ESQ;
  msg[0] = '\0';
  if (smax_index_open(gt_str_get(arguments->indexname), demand, &idx, msg,
                      sizeof msg) != 0)
  {
    gt_error_set(err, "%s", msg);
    return -1;
  }
  if (arguments->beverbose)  /* GtLogger lines: "# " on stdout, gt_repfind.c:520 */
  {
    smax_index_info info;
    smax_index_info_get(idx, &info);
    printf("# indexname=%s\n", gt_str_get(arguments->indexname));
    printf("# numberofallsortedsuffixes=%lu\n",
           (unsigned long) info.numberofallsortedsuffixes);
    printf("# largelcpvalues=%lu\n", (unsigned long) info.largelcpvalues);
    printf("# suftab uses %ubit values\n", info.sufbytes * 8);
    printf("# minlength=%u gpus=%u policy=%s\n", arguments->minlength,
           arguments->gpus, gt_str_get(arguments->policy));
  }
  if (on_device)
  {
    if (smax_run_text(idx, &opts, stdout, NULL, msg, sizeof msg) != 0)
      had_err = -1;
  } else
  {
    if (smax_emitter_new(idx, &opts, stdout, &em, msg, sizeof msg) != 0)
      had_err = -1;
    if (!had_err &&
        (arguments->scanfile
           ? smax_run_stream(idx, &opts, 0, smax_emitter_emit, em, msg, sizeof msg)
           : smax_run(idx, &opts, smax_emitter_emit, em, msg, sizeof msg)) != 0)
      had_err = -1;
    if (em != NULL && smax_emitter_delete(em) != 0 && !had_err)
    {
      snprintf(msg, sizeof msg, "cannot write results");
      had_err = -1;
    }
  }
  smax_index_close(idx);
  if (had_err)
    gt_error_set(err, "%s", msg[0] != '\0' ? msg : "libsmax call failed");
  return had_err;
}

GtTool* gt_smax(void)
{
  return gt_tool_new(gt_smax_arguments_new,
                     gt_smax_arguments_delete,
                     gt_smax_option_parser_new,
                     gt_smax_arguments_check,
                     gt_smax_runner);
}
