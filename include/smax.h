/*
  smax.h -- C ABI of libsmax.so, the B200-native supermaximal-repeat scan over
  a GenomeTools enhanced suffix array (.prj/.suf/.lcp/.llv/.bwt/.esq).

  Plain C types only (pointers, sizes, fixed-width integers); no CUDA or torch
  types cross this boundary.  Every fallible function returns 0 on success and
  a negative value on error and writes a message into (err, errlen) -- the same
  "int + message" convention as GenomeTools' trailing GtError*
  (/root/reference/src/core/error_api.h:24-58).

  Reference interfaces each entry point replaces:

    smax_index_open/close   gt_mapsuffixarray + gt_freesuffixarray
                            (src/match/esa-map.c:503-517, :260-294; the
                            demand bits mirror SARR_* in src/match/sarr-def.h:33-40)
    smax_index_info         fields of Suffixarray / the .prj keys
                            (src/match/sarr-def.h:101-126, src/match/esa-map.c:78-122)
    smax_run                the algorithm-level call of the smax tool; shape of
                            gt_callenummaxpairs(indexname, minlength, scan, cb,
                            cbinfo, logger, err) (src/match/esa-maxpairs.h:57-63)
                            with a per-repeat callback instead of GtProcessmaxpairs
                            (src/match/esa-maxpairs.h:38-43)
    smax_tool_main          gt_tool_run over the five GtTool callbacks
                            (src/core/tool.c:62-114, src/core/tool_api.h:30-70;
                            template src/tools/gt_repfind.c:625-632)
    smax_device_* / smax_scan_*
                            no reference counterpart: the device-resident half
                            of the path (table upload, the fused scan kernel,
                            record fetch), exposed so a host that already owns
                            device memory and streams (torch, the bench, the
                            multi-GPU driver) can drive the same kernels.
*/
#ifndef SMAX_H
#define SMAX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMAX_VERSION "0.2.0"

/* error codes (every other failure is -1) */
#define SMAX_E_RANGE (-2)   /* smax_scan_counts / smax_scan_fetch: a plateau reaches further left
                               than the shard's own arrays and its left neighbour views; the
                               tables are valid -- make more of them resident (a wider halo) */
/* suffixes one device shard may hold (tile offsets within a shard are 32-bit) */
#define SMAX_MAX_SHARD_LEN ((1ull << 32) - (1ull << 20))

/* demand bits for smax_index_open (cf. SARR_ESQTAB.. in sarr-def.h:33-40) */
#define SMAX_TAB_ESQ 1u
#define SMAX_TAB_SUF (1u << 1)
#define SMAX_TAB_LCP (1u << 2)
#define SMAX_TAB_BWT (1u << 3)
#define SMAX_TAB_ALL (SMAX_TAB_ESQ | SMAX_TAB_SUF | SMAX_TAB_LCP | SMAX_TAB_BWT)

/* left-character policy of the distinctness test */
#define SMAX_POLICY_GT    0  /* specials (>=254) differ from everything: esa-maxpairs.c:24-31 */
#define SMAX_POLICY_PLAIN 1  /* 254 / 255 are ordinary codes */

/* text formats of the emitter (the reference tool's grammar is unpinned,
   SURVEY.md section D; the format is a single switch point) */
#define SMAX_FORMAT_SMAX 0   /* "<len> <count> <pos>..."                       */
#define SMAX_FORMAT_ITV  1   /* "<len> <lb> <rb>"   (cf. esa-lcpintervals.c:183-189) */
#define SMAX_FORMAT_PAIRS 2  /* one "<len> <p1> F <len> <p2>" line per pair
                                (cf. src/match/querymatch.c:169-187) */

typedef struct smax_index smax_index;     /* host: mmapped ESA tables   */
typedef struct smax_device smax_device;   /* one GPU: resident shard    */

/* one .llv record, src/match/lcpoverflow.h:26-30 */
typedef struct
{
  uint64_t position, value;
} smax_llv;

/* one supermaximal repeat: SA interval [lb, lb+width), repeat length len */
typedef struct
{
  uint64_t len, lb, width;
} smax_record;

typedef struct
{
  uint64_t totallength, specialcharacters, numofsequences,
           numberofallsortedsuffixes, /* n */
           nonspecials, largelcpvalues, maxbranchdepth, longest;
  uint32_t integersize, littleendian, readmode, mirrored;
  uint32_t sufbytes;      /* 8, 4 (-suftabuint) or 0 if not demanded */
  uint32_t alphatype;     /* 0 DNA, 1 protein, 2 other (from .esq, encseq.c:1014-1042) */
  uint32_t numofchars;    /* alphabet size sigma (4, 20, ...) or 0 if .esq not demanded */
  uint32_t reserved;
} smax_index_info;

typedef struct
{
  uint64_t minlength;     /* >= 1                                        */
  int32_t relative;       /* 0: absolute positions, 1: seqnum/relpos     */
  int32_t ngpus;          /* 0/1: one GPU; N: SA range sharded over N    */
  int32_t policy;         /* SMAX_POLICY_*                               */
  int32_t format;         /* SMAX_FORMAT_* (used by the text emitter)    */
  int32_t first_device;   /* CUDA ordinal of shard 0                     */
  int32_t verbose;
} smax_opts;

/* per-repeat callback; positions = suf[lb .. lb+width) in SA order.
   Return non-zero to stop (smax_run then fails with that message-less rc). */
typedef int (*smax_emit_cb)(void *info, uint64_t len, uint64_t lb,
                            uint64_t width, const uint64_t *positions);

/* ------------------------------ host side ------------------------------ */
int smax_index_open(const char *indexname, unsigned demand, smax_index **out,
                    char *err, size_t errlen);
/* wrap caller-owned host tables (no files); suf may be NULL, sufbytes 8|4 */
int smax_index_from_memory(const uint8_t *lcp, const uint8_t *bwt,
                           const smax_llv *llv, uint64_t nllv,
                           const void *suf, unsigned sufbytes, uint64_t n,
                           smax_index **out, char *err, size_t errlen);
/* as above for a WINDOW of a larger table: the arrays hold lcp indices
   [base, base+len) of a table with n_total entries (llv positions stay
   global).  Used by one-process-per-GPU hosts that only keep their shard. */
int smax_index_from_memory_window(const uint8_t *lcp, const uint8_t *bwt,
                                  const smax_llv *llv, uint64_t nllv,
                                  const void *suf, unsigned sufbytes,
                                  uint64_t base, uint64_t len, uint64_t n_total,
                                  smax_index **out, char *err, size_t errlen);
void smax_index_close(smax_index *idx);
int smax_index_info_get(const smax_index *idx, smax_index_info *info);
const uint8_t *smax_index_lcptab(const smax_index *idx);
const uint8_t *smax_index_bwttab(const smax_index *idx);
const smax_llv *smax_index_llvtab(const smax_index *idx);
const void *smax_index_suftab(const smax_index *idx);

/* whole path: upload -> scan on ngpus devices -> ordered records -> callback */
int smax_run(const smax_index *idx, const smax_opts *opts, smax_emit_cb cb,
             void *info, char *err, size_t errlen);
/* as smax_run but returns the records (malloc'd; free with smax_free) */
int smax_run_records(const smax_index *idx, const smax_opts *opts,
                     smax_record **recs, uint64_t *nrecs,
                     char *err, size_t errlen);
void smax_free(void *p);
/* smax_run / smax_run_records / smax_run_text keep their device handles (contexts, streams,
   table allocations) for the next call of the process; this frees them (also done at exit). */
void smax_release_devices(void);
/* bytes this process has copied host -> device through libsmax so far (measurement: the .llv
   records cross the link as 4-byte values -- their positions are redundant with the lcp table,
   /root/reference/src/match/sarr-def.h:128-178 reads them with a cursor for the same reason --
   so the bytes of an upload are not simply the sizes of the tables) */
uint64_t smax_h2d_bytes_total(void);
/* occurrence positions of the records, in record order: out[] receives
   suf[lb .. lb+width) of every record (sum of widths entries) from the host
   suffix table of idx -- what smax_run hands to its callback */
int smax_index_gather_positions(const smax_index *idx, const smax_record *recs,
                                uint64_t nrecs, uint64_t *out, char *err, size_t errlen);

/* text emitter: formats one repeat into buf (returns bytes written) */
typedef struct smax_emitter smax_emitter;
int smax_emitter_new(const smax_index *idx, const smax_opts *opts, void *file,
                     smax_emitter **out, char *err, size_t errlen);
int smax_emitter_emit(void *emitter, uint64_t len, uint64_t lb, uint64_t width,
                      const uint64_t *positions);   /* an smax_emit_cb */
/* the same for a batch: positions holds suf[lb..lb+width) of every record
   back to back (NULL for SMAX_FORMAT_ITV) */
int smax_emitter_emit_records(smax_emitter *em, const smax_record *recs, uint64_t nrecs,
                              const uint64_t *positions);
int smax_emitter_delete(smax_emitter *em);           /* flushes */

/* as smax_run with an smax_emitter behind it, but the text is rendered ON THE
   DEVICES (smax_scan_format): every shard is made resident with its suffix
   table, scanned with the position gather, formatted in HBM, and the bytes
   are written to `file` (a FILE*, NULL = stdout) in shard order.  Same bytes
   as smax_run + smax_emitter_emit for SMAX_FORMAT_SMAX / SMAX_FORMAT_ITV
   (absolute or relative); SMAX_FORMAT_PAIRS is host-only and fails here.
   *nbytes (may be NULL) receives the number of bytes written. */
int smax_run_text(const smax_index *idx, const smax_opts *opts, void *file,
                  uint64_t *nbytes, char *err, size_t errlen);
/* the "-scan" mode (streamsuffixarray, src/match/esa-map.c:488-501): idx is
   opened WITHOUT tables (demand 0 or SMAX_TAB_ESQ); the SA range is processed
   in chunks of `chunk` suffixes (0 = 2^28) whose table bytes are read from the
   index files, scanned on ONE GPU (opts->first_device) and dropped again; the
   two previous chunks stay resident for plateaus that cross a cut, and a chunk
   whose plateau reaches even further back is redone with a wider window.  Results reach the
   callback as in smax_run (positions = NULL for SMAX_FORMAT_ITV). */
int smax_run_stream(smax_index *idx, const smax_opts *opts, uint64_t chunk,
                    smax_emit_cb cb, void *info, char *err, size_t errlen);
/* ascending absolute positions of the sequence separators, recovered from the
   tables ({ suf[i] - 1 : bwt[i] == 255 }); what gt_encseq_seqnum /
   gt_encseq_seqstartpos (src/core/encseq.c:3815-3900) answer from.  The
   pointer stays valid until smax_index_close. */
int smax_index_separators(smax_index *idx, const uint64_t **seps, uint64_t *nseps,
                          char *err, size_t errlen);

/* CLI entry: GtTool-shaped option parsing + runner; returns the exit code */
int smax_tool_main(int argc, const char **argv);

/* ----------------------------- device side ----------------------------- */
/* number of visible CUDA devices, or a negative value (no driver / no GPU) */
int smax_device_count(char *err, size_t errlen);

int smax_device_create(int ordinal, smax_device **out, char *err, size_t errlen);
void smax_device_destroy(smax_device *dev);
/* waits until the device has finished everything issued through this handle */
int smax_device_synchronize(smax_device *dev);

/* Make the SA range [lo, hi) of idx resident on the device (lcp, bwt, llv;
   suf too when with_suf != 0).  lo/hi are lcp indices; the shard owns every
   plateau whose END lies in [lo, hi).  Copies go through pinned staging
   buffers.  Total bytes copied host->device are added to *h2d_bytes. */
int smax_device_upload(smax_device *dev, const smax_index *idx, uint64_t lo,
                       uint64_t hi, int with_suf, uint64_t *h2d_bytes,
                       char *err, size_t errlen);

/* As smax_device_upload with a caller-chosen left halo (entries of the tables
   left of lo that are made resident too; at least 256, rounded up to 16):
   the -scan driver widens it for a chunk whose plateau reaches further back
   than its neighbours cover. */
int smax_device_upload_halo(smax_device *dev, const smax_index *idx, uint64_t lo,
                            uint64_t hi, uint64_t halo, int with_suf,
                            uint64_t *h2d_bytes, char *err, size_t errlen);

/* Adopt tables that are ALREADY in device memory (e.g. torch tensors):
   d_lcp/d_bwt cover lcp indices [a_lo, a_hi) and must be readable up to
   SMAX_PAD bytes past a_hi - a_lo (zero padded); a_lo must be a multiple of
   16 and the pointers 16-byte aligned; d_llv holds the nllv records whose
   position lies in [a_lo, a_hi).  The shard owns [lo, hi). */
#define SMAX_PAD 64
int smax_device_adopt(smax_device *dev, const void *d_lcp, const void *d_bwt,
                      const void *d_llv, uint64_t nllv, const void *d_suf,
                      unsigned sufbytes, uint64_t a_lo, uint64_t a_hi,
                      uint64_t lo, uint64_t hi, uint64_t n_total,
                      char *err, size_t errlen);

/* Peer shards for plateaus that cross a cut: the scan walks left out of its
   own arrays into the left neighbours' through these views (P2P loads over
   NVLink when they live on another GPU).  A view is a POD that can be sent
   to another process together with CUDA IPC handles. */
typedef struct
{
  uint64_t a_lo, a_hi;        /* coverage of the arrays (lcp index space) */
  uint64_t d_lcp, d_bwt, d_llv, d_llvdir, d_suf;   /* device addresses (0 = absent) */
  uint64_t nllv;
  int32_t device;             /* CUDA ordinal that owns the memory */
  uint32_t sufbytes;
} smax_shard_view;

int smax_device_view(const smax_device *dev, smax_shard_view *view);
/* views[0..nviews) in shard order (ascending a_hi, the nearest neighbour last), none beginning
   right of this shard's a_lo (the scan only ever walks left; a neighbour that was made resident
   again with a wider halo may begin left of the neighbours before it); enables peer access
   when in-process. */
int smax_device_set_left_views(smax_device *dev, const smax_shard_view *views,
                               int nviews, char *err, size_t errlen);
/* CUDA IPC plumbing for one-process-per-GPU launches (torchrun): export the
   5 table allocations of this shard (lcp, bwt, llv, llvdir, suf) / map a
   neighbour's into this process (fills the device addresses of the view). */
#define SMAX_IPC_BYTES 64
#define SMAX_IPC_TABLES 5
int smax_device_ipc_export(const smax_device *dev,
                           uint8_t handles[SMAX_IPC_TABLES][SMAX_IPC_BYTES],
                           smax_shard_view *view, char *err, size_t errlen);
int smax_device_ipc_import(smax_device *dev,
                           const uint8_t handles[SMAX_IPC_TABLES][SMAX_IPC_BYTES],
                           smax_shard_view *view_inout, char *err, size_t errlen);

/* One-sided exchange of the per-shard record counts of a multi-GPU scan (what
   turns local output positions into global ones).  Every shard owns a small
   device array of `world` slots; at the end of its scan the kernel itself
   stores its record count, tagged with the caller's exchange tag, into slot
   [rank] of every shard's array (P2P stores over NVLink for the peers) -- no
   collective launch on the step's critical path.
     smax_device_counts_export   allocate the array, return its device address
                                 and a CUDA IPC handle for other processes
     smax_device_counts_connect  addresses of all `world` arrays in rank order
                                 (entry [rank] = own); handles != NULL: entries
                                 of other processes are opened through IPC
     smax_device_set_exchange_tag  tag the following scans (same value on all
                                 ranks of a step, > 0, different from the last)
     smax_scan_peer_counts       wait for this shard's own scan, then until all
                                 slots of the own array carry `tag`; return them */
#define SMAX_MAX_PEERS 16
int smax_device_counts_export(smax_device *dev, int world, uint8_t handle[SMAX_IPC_BYTES],
                              uint64_t *d_ptr, char *err, size_t errlen);
int smax_device_counts_connect(smax_device *dev, int rank, int world,
                               const uint8_t (*handles)[SMAX_IPC_BYTES],
                               const uint64_t *d_ptrs, char *err, size_t errlen);
int smax_device_set_exchange_tag(smax_device *dev, uint64_t tag);
int smax_scan_peer_counts(smax_device *dev, uint64_t tag, uint64_t *counts,
                          char *err, size_t errlen);

/* One scan of the resident shard on `stream` (a cudaStream_t passed as
   void*, NULL = default stream).  Asynchronous: results stay on the device
   until smax_scan_fetch.  gather != 0 additionally gathers the occurrence
   positions suf[lb..lb+width) on the device (needs a resident suf). */
int smax_scan_launch(smax_device *dev, uint64_t minlength, int policy,
                     int gather, void *stream, char *err, size_t errlen);
/* Waits for the scan; returns counts. */
int smax_scan_counts(smax_device *dev, uint64_t *nrecs, uint64_t *npositions,
                     char *err, size_t errlen);
/* Copies records (and positions, if gathered; may be NULL) to host buffers
   with capacity for the counts smax_scan_counts reported. */
int smax_scan_fetch(smax_device *dev, smax_record *recs, uint64_t *positions,
                    char *err, size_t errlen);
/* Device time of the last smax_scan_launch in milliseconds (CUDA events on
   the launching stream; waits for completion): *ms = whole launch sequence,
   *ms_scan = the scan kernel alone (event between scan and gather); and the
   number of kernel launches it issued. */
int smax_scan_elapsed_ms(smax_device *dev, float *ms, float *ms_scan, int *launches,
                         char *err, size_t errlen);
/* Asynchronously copies the record count of the last scan (one uint64) to
   d_dst on `stream` -- feeds the NCCL count exchange without a host trip. */
int smax_scan_copy_count(smax_device *dev, void *d_dst, void *stream,
                         char *err, size_t errlen);
/* Device pointers of the last scan's outputs (for device-side consumers). */
int smax_scan_device_buffers(smax_device *dev, uint64_t *d_records,
                             uint64_t *d_positions, uint64_t *d_count);

/* ---- emit on the device ----
   Separator table for relative positions: either uploaded from the host
   (smax_index_separators; any shard) or built on the device from the resident
   bwt + suffix tables (only when the device holds the whole index). */
int smax_device_set_separators(smax_device *dev, const uint64_t *seps, uint64_t nseps,
                               char *err, size_t errlen);
int smax_device_build_separators(smax_device *dev, uint64_t *nseps,
                                 char *err, size_t errlen);
int smax_device_fetch_separators(smax_device *dev, uint64_t *seps, uint64_t *nseps,
                                 char *err, size_t errlen);
/* Render the records of the last scan as text in HBM, on the scan's stream:
   SMAX_FORMAT_SMAX ("<len> <count> <pos>...", needs a scan with gather;
   relative != 0: "<len> <count> <seq> <rel>...") or SMAX_FORMAT_ITV
   ("<len> <lb> <rb>").  Byte-identical to smax_emitter_emit over the same
   records.  *nbytes = size of the text. */
int smax_scan_format(smax_device *dev, int format, int relative, uint64_t *nbytes,
                     char *err, size_t errlen);
/* Copies the text (nbytes of smax_scan_format, no terminator) to dst. */
int smax_scan_fetch_text(smax_device *dev, char *dst, char *err, size_t errlen);
/* Device time of the last smax_scan_format (CUDA events; includes the one
   host round trip that reads the text size). */
int smax_scan_format_elapsed_ms(smax_device *dev, float *ms, char *err, size_t errlen);

/* Algorithmic-byte accounting of the last scan:
   stats[0]=n scanned, [1]=candidate plateaus, [2]=sum of candidate widths,
   [3]=llv records inspected, [4]=survivors, [5]=sum of survivor widths.
   Only filled when the scan was launched after smax_device_set_stats(dev,1). */
int smax_device_set_stats(smax_device *dev, int on);
/* Limit the number of CTAs of the scan kernel (0 = as many as are resident).
   Results do not depend on it; the tests use it to make a few CTAs walk
   through many tiles each (ring reuse, mid-scan flushes of the survivor log). */
int smax_device_set_grid_limit(smax_device *dev, int max_ctas);
/* Tuning probes (tools/probe_scan.py): results are NOT valid while flags != 0. */
int smax_device_set_debug(smax_device *dev, int flags);
int smax_scan_stats(smax_device *dev, uint64_t stats[8], char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* SMAX_H */
