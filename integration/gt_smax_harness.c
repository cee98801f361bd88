/*
  gt_smax_harness.c -- TEST INFRASTRUCTURE: what src/gt.c + src/gtr.c do for one tool.
  Registers gt_smax() in a GtToolbox, looks it up by argv[1], runs it through the
  reference's own gt_tool_run (src/core/tool.c:62-114) and prints an error the way
  src/gt.c:48-50 does.  Linked against the reference library built by
  oracle/Makefile.ref (oracle/_ref/libgtref.a) and against libsmax.so.

      gt_smax_harness smax -l 20 -ii <index>
*/
#include <stdio.h>
#include <string.h>
#include "core/init_api.h"
#include "core/error_api.h"
#include "core/error.h"
#include "core/toolbox_api.h"
#include "core/tool_api.h"
#include "core/cstr_array.h"
#include "tools/gt_smax.h"

int main(int argc, char **argv)
{
  GtToolbox *tools;
  GtTool *tool;
  GtError *err;
  char **nargv;
  int rc = 0;

  if (argc < 2)
  {
    fprintf(stderr, "usage: %s smax [options]\n", argv[0]);
    return 2;
  }
  gt_lib_init();
  err = gt_error_new();
  tools = gt_toolbox_new();
  gt_toolbox_add_tool(tools, "smax", gt_smax());      /* the line next to src/gtt.c:234 */
  tool = gt_toolbox_get_tool(tools, argv[1]);
  if (tool == NULL)
  {
    fprintf(stderr, "gt: error: neither tool nor script '%s' found\n", argv[1]);
    rc = 1;
  } else
  {
    /* gtr_run prefixes argv[0] with the program name: "gt smax" (src/gtr.c:419-497) */
    nargv = gt_cstr_array_prefix_first((const char**) argv + 1, "gt");
    gt_error_set_progname(err, nargv[0]);
    if (gt_tool_run(tool, argc - 1, (const char**) nargv, err) != 0)
    {
      fprintf(stderr, "%s: error: %s\n", gt_error_get_progname(err), gt_error_get(err));
      rc = 1;
    }
    gt_cstr_array_delete(nargv);
  }
  gt_toolbox_delete(tools);
  gt_error_delete(err);
  if (gt_lib_clean())
    return 2;           /* memory / file pointer / mmap leak (GT_MEM_BOOKKEEPING) */
  return rc;
}
