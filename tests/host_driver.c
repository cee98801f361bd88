/*
  host_driver.c -- TEST INFRASTRUCTURE: drives the host side of libsmax (loader,
  smax_run / smax_run_stream, emitter) over tests/host_stub_device.c.

    host_driver <indexname> <minlength> map|stream <chunk> smax|itv|pairs <relative 0|1> [policy [ngpus]]
*/
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "smax.h"

int main(int argc, char **argv)
{
  char err[1024] = "";
  smax_index *idx = NULL;
  smax_emitter *em = NULL;
  smax_opts opts;
  int stream, rc = 1;
  if (argc < 7)
  {
    fprintf(stderr, "usage: %s indexname minlength map|stream chunk smax|itv|pairs relative [policy]\n",
            argv[0]);
    return 2;
  }
  memset(&opts, 0, sizeof opts);
  opts.minlength = strtoull(argv[2], NULL, 10);
  stream = strcmp(argv[3], "stream") == 0;
  opts.format = strcmp(argv[5], "itv") == 0 ? SMAX_FORMAT_ITV
              : strcmp(argv[5], "pairs") == 0 ? SMAX_FORMAT_PAIRS : SMAX_FORMAT_SMAX;
  opts.relative = atoi(argv[6]);
  opts.policy = argc > 7 && strcmp(argv[7], "plain") == 0 ? SMAX_POLICY_PLAIN : SMAX_POLICY_GT;
  opts.ngpus = argc > 8 ? atoi(argv[8]) : 1;
  if (smax_index_open(argv[1], stream ? 0u : (SMAX_TAB_SUF | SMAX_TAB_LCP | SMAX_TAB_BWT), &idx, err, sizeof err) != 0)
    goto done;
  if (smax_emitter_new(idx, &opts, stdout, &em, err, sizeof err) != 0)
    goto done;
  if (stream)
    rc = smax_run_stream(idx, &opts, strtoull(argv[4], NULL, 10), smax_emitter_emit, em, err,
                         sizeof err) != 0;
  else
    rc = smax_run(idx, &opts, smax_emitter_emit, em, err, sizeof err) != 0;
done:
  if (em != NULL && smax_emitter_delete(em) != 0 && rc == 0)
  {
    snprintf(err, sizeof err, "cannot write results");
    rc = 1;
  }
  smax_index_close(idx);
  if (rc != 0)
    fprintf(stderr, "error: %s\n", err);
  return rc;
}
