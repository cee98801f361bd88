/*
  gt_smax.h -- the GenomeTools side of the drop-in: `gt smax` as a GtTool.
  Goes to src/tools/gt_smax.h of a GenomeTools tree.
*/
#ifndef GT_SMAX_H
#define GT_SMAX_H

#include "core/tool_api.h"

/* the smax tool: supermaximal repeats of an enhanced suffix array, computed by libsmax on
   B200 GPUs.  Register with gt_toolbox_add_tool(tools, "smax", gt_smax()) next to
   src/gtt.c:234. */
GtTool* gt_smax(void);

#endif
