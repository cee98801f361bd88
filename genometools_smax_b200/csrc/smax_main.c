/*
  smax_main.c -- `smax` executable: the tool of smax_tool.c as a stand-alone
  program (inside GenomeTools it would be reached as `gt smax`, dispatched by
  /root/reference/src/gtr.c:419-497).
*/
#include "smax.h"

int main(int argc, char **argv)
{
  return smax_tool_main(argc, (const char **) argv);
}
