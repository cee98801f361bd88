/*
  smax_kernels.cuh -- device-side data structures shared by the kernels
  (smax_scan.cu) and the device manager (smax_device.cu).  sm_100a only.
*/
#ifndef SMAX_KERNELS_CUH
#define SMAX_KERNELS_CUH

#include <cstdint>
#include <cuda_runtime.h>
#include "smax.h"

#ifndef SMAX_PROBE
#define SMAX_PROBE 0        // 1: tuning build with per-phase timers in k_scan (never the shipped library)
#endif

namespace smax {

// ---- geometry of the scan kernel -----------------------------------------
constexpr int kThreads   = 128;               // threads per CTA: 4 warps that work independently
constexpr int kWarps     = kThreads / 32;
#ifndef SMAX_MINBLOCKS
#define SMAX_MINBLOCKS 7
#endif
constexpr int kMinBlocks = SMAX_MINBLOCKS;    // resident CTAs per SM the kernel is compiled for
constexpr int kChunk     = 16;                // bytes per 128-bit shared-memory load
constexpr int kUnitBytes = 4096;              // lcptab entries per unit (what a warp takes at a time)
constexpr int kUnitChunks = kUnitBytes / kChunk;   // 256: a chunk number within a unit fits a byte
constexpr int kHalo      = 16;                // table bytes staged either side of a unit
constexpr int kBitWords  = kUnitBytes / 32;   // words of a per-unit position bitmap
constexpr int kLlvRow    = 128;               // .llv records a warp classifies per step (four per lane)
constexpr int kLlvPad    = 8;                 // "no record" entries behind the compact .llv tables
constexpr int kSlowList  = 128;               // records a warp collects for the general (slow) path
constexpr int kCandList  = 224;               // large-value candidates of SA width 2 a warp collects
constexpr int kEndList   = 448;               // ENDs of small values a warp collects
constexpr int kArenaChunk = 256;              // survivor arena entries a warp takes per allocation
constexpr int kMaxLeft   = 8;                 // peer shards a plateau may walk into
constexpr int kLlvBucketShift = 12;           // .llv directory: one entry per 4096 lcp entries
// compact .llv records, two parallel arrays built at upload (k_llvpack):
//   llvv[k]  bits 0..29 value (kLlvEscape: does not fit, read the 16-byte record);
//            bit 30 PEAK: the value rises from the entry before it and falls to the entry behind it
//            (neighbours that are no large values are smaller by definition) -- the record ends a
//            plateau of SA width 2 whenever its value reaches the minimum length;
//            bit 31 GENERAL: the record may end a wider plateau (it closes a run of equal large
//            values), or its neighbourhood holds a value that does not fit -- decided by the general
//            path (large_plateau).  Both bits are properties of the table, not of a scan: K1 for
//            large values is one compare per record.
//   llvp[k]  position - a_lo
constexpr uint32_t kLlvValueMask = 0x3fffffffu;
constexpr uint32_t kLlvEscape = kLlvValueMask;
constexpr uint32_t kLlvPeak    = 0x40000000u;
constexpr uint32_t kLlvGeneral = 0x80000000u;
constexpr uint32_t kNoRecord  = 0xfffffffeu;  // compact .llv position of "no record" (never adjacent to one)

// tile status of the decoupled look-back: per tile two 16-byte pairs (record
// count, position count) -- the tile's own aggregate and its inclusive prefix,
// each written once per scan -- every word [63:35] epoch, [34:33] state,
// [32:0] value.  The epoch makes a memset between scans unnecessary.
constexpr uint64_t kStateInvalid = 0, kStateAggregate = 1, kStatePrefix = 2;
constexpr int kStatusWords = 4;               // 64-bit words per tile
constexpr int kValueBits = 33;
constexpr uint64_t kValueMask = (1ull << kValueBits) - 1;
constexpr uint32_t kEpochMask = (1u << 29) - 1;

// One shard's tables as the kernel sees them.  Element i (global lcp index)
// of a table lives at table[i - a_lo].
struct TableView
{
  const uint8_t  *lcp;
  const uint8_t  *bwt;
  const smax_llv *llv;      // records with position in [a_lo, a_hi)
  const uint32_t *llvdir;   // lower_bound(llv.position, a_lo + b*4096), b = 0..nbuckets
  const uint32_t *llvv;     // compact records: values + run flags (own shard only, see kLlvFirst)
  const uint32_t *llvp;     // compact records: position - a_lo
  const void     *suf;      // may be null
  uint64_t nllv;
  uint64_t a_lo, a_hi;
};

// indices into the result block (device, uint64 each)
enum ResultSlot
{
  kResCount = 0,        // number of records (exact even when capacity overflowed)
  kResOverflow = 1,     // != 0: record or position capacity was too small
  kResError = 2,        // != 0: inconsistent tables (missing .llv record, ...)
  kResPositions = 3,    // number of gathered positions
  kResStatCand = 4,     // candidate plateaus (local maxima with value >= minlength)
  kResStatCandWidth = 5,
  kResStatLlv = 6,      // .llv records inspected
  kResStatSurvWidth = 7,
  kResWalks = 8,        // runs of >= 4 equal values walked entry by entry
  kResArena = 9,        // survivor arena entries handed out (> capacity: overflow)
#if SMAX_PROBE
  kResProbe = 12,       // tuning build (tools/probe_units.py): nanoseconds per phase of the unit kernel
  kResSlots = 20
#else
  kResSlots = 12
#endif
};

// K3 works in two steps.  The detection kernel leaves, per "unit" (what a warp takes at a
// time), the number of supermaximal repeats and of their occurrences and where the unit's
// entries sit in the survivor arena, and adds both numbers to the sums of the unit's block of
// kEmitBlock units; k_emit (one CTA per block) takes the sums of the blocks before its own and
// the aggregates of its own units, and writes every arena entry to its place in suffix-array order.
struct UnitMeta
{
  uint32_t count;        // repeats that end in the unit
  uint32_t pad;
  uint64_t wsum;         // occurrences of those repeats
  uint64_t base;         // arena index of the unit's first entry (its entries are consecutive;
                         //   ~0: they did not fit the arena region of the warp that found them)
};
struct ArenaEntry
{
  uint32_t unit;
  uint32_t end_off;      // the repeat's last suffix-array index, relative to a_lo
  uint32_t width;        // SA width
  uint32_t pad;
  uint32_t len, len_hi;  // repeat length
};
constexpr int kEmitBlock     = 32;            // k_emit: units per CTA,
constexpr int kEmitThreads   = 256;           //   its threads,


struct ScanParams
{
  TableView own;
  TableView left[kMaxLeft];   // in shard order; left[nleft-1] is the nearest neighbour
  int nleft;
  int policy;
  int sufbytes;               // 8 or 4
  int debug;                  // tuning probes only (tools/probe_scan.py; results are not valid while != 0):
                              //   1 no look-back, 2 no small values, 4 no large values, 8 filter only, 16 no write
  uint32_t epoch;
  uint64_t g_lo, g_hi;        // plateau ENDS in [g_lo, g_hi) belong to this shard
  uint64_t minlength;
  uint32_t mb;                // min(minlength, 255): byte threshold of the filter
  uint32_t nunits;            // units of the shard's own range [g_lo, g_hi)
  smax_record *recs;
  uint64_t rec_capacity;
  uint64_t *positions;        // null: do not gather positions
  uint64_t pos_capacity;
  uint64_t *blocksum;         // per block of kEmitBlock units: {repeats, occurrences}, summed by the scan
  uint64_t *blocksum_next;    //   the other set, zeroed by k_emit for the next scan
  UnitMeta *meta;             // nunits unit aggregates
  ArenaEntry *arena;          // the scan's survivors, unit by unit (handed out in chunks of
  uint64_t arena_capacity;    //   kArenaChunk entries per warp)
  const uint32_t *unitdir;    // nunits + 1: first .llv record at or behind the start of each unit
  const uint32_t *unitorder;  // nunits: the units, heaviest first (built at upload, k_unitorder_*); null: as they come
  int has_escape;             // != 0: some compact .llv record holds kLlvEscape
  int edge_rec0;              // != 0: record 0 sits on the first entry of the shard's arrays and
                              //   the table goes on to the left (what precedes it is in a neighbour shard)
  uint32_t *ctrl;             // [0] ticket, [1] finished CTAs
  uint64_t *peer_counts[SMAX_MAX_PEERS];   // count arrays of all shards (one-sided exchange), or none
  int npeers, my_rank;
  uint64_t exchange_tag;      // < 2^24; stored above the count
  uint64_t *result;           // kResSlots words of this scan
  uint64_t *result_next;      // the other block, zeroed by the last CTA for the next scan
};

// launchers (smax_scan.cu)
cudaError_t launch_llvdir(const smax_llv *llv, uint64_t nllv, uint64_t a_lo,
                          uint32_t *dir, uint64_t nentries, cudaStream_t st);
cudaError_t launch_llvpack(const smax_llv *llv, uint64_t nllv, uint64_t a_lo, uint32_t *vals,
                           uint32_t *poss, uint32_t *has_escape, cudaStream_t st);
cudaError_t launch_lcphist(const uint8_t *lcp, uint64_t len, unsigned long long *hist, int sm_count,
                           cudaStream_t st);
cudaError_t launch_unitdir(const smax_llv *llv, uint64_t nllv, uint64_t g_lo, uint64_t g_hi,
                           uint32_t *dir, uint64_t ntiles, cudaStream_t st);
// the order the scan takes the units in: heaviest first (weight = entries that reach a threshold
// taken from the lcp histogram + .llv records), so that what is in flight at the end of the
// scan are the light units.  order: nunits words + kOrderScratch words of scratch behind them.
constexpr int kOrderBuckets = 1024;
constexpr int kOrderScratch = kOrderBuckets + 8;
cudaError_t launch_unitorder(const uint8_t *lcp_own, uint64_t own_len, const uint32_t *unitdir,
                             const unsigned long long *hist, uint32_t *order, uint64_t nunits,
                             cudaStream_t st);
// the shard's 16-byte .llv records from its lcp bytes (the k-th 255 byte is the k-th record) and
// the values the host uploaded (4 bytes each); scratch[len / 65536 (rounded up)] = 255 bytes found
cudaError_t launch_llv_rebuild(const uint8_t *lcp, uint64_t len, uint64_t a_lo, const uint32_t *vals32,
                               uint64_t nllv, smax_llv *llv, uint32_t *scratch, cudaStream_t st);
uint64_t llv_rebuild_scratch_words(uint64_t len);
// the two launches of one scan: k_scan (detection), k_emit
cudaError_t launch_scan(const ScanParams &p, bool stats, int grid, int sm_count, cudaStream_t st);
constexpr int kScanLaunches = 2;
int scan_blocks_per_sm(bool stats);

// ---- device-side text formatting of a scan's records (smax_format.cu) ---------
struct FormatJob
{
  const smax_record *recs;
  uint64_t nrecs;
  const uint64_t *pos;        // gathered positions (SMAX_FORMAT_SMAX), record order
  uint64_t npos;
  const uint64_t *seps;       // ascending separator positions (relative output)
  uint64_t nseps;
  int format, relative;
  uint64_t *sums;             // format_sums_words(max(nrecs, npos)) words
  uint64_t *hoff;             // nrecs + 1: bytes of the own text of records < r
  uint64_t *pfirst;           // nrecs + 1: index of the first position of record r
  uint64_t *poff;             // npos + 1: bytes of the position items < j
  char *text;
};
uint64_t format_sums_words(uint64_t n);
cudaError_t launch_format_measure(const FormatJob &j, cudaStream_t st);
cudaError_t launch_format_write(const FormatJob &j, cudaStream_t st);
cudaError_t launch_sep_mark(const uint8_t *bwt, const void *suf, int sufbytes, uint64_t len,
                            uint64_t *bitmap, uint64_t nbits, uint64_t *bad, int sm_count,
                            cudaStream_t st);
cudaError_t launch_sep_rank(const uint64_t *bitmap, uint64_t nwords, uint64_t *sums, uint64_t *rank,
                            cudaStream_t st);
cudaError_t launch_sep_fill(const uint64_t *bitmap, uint64_t nwords, const uint64_t *rank,
                            uint64_t *seps, cudaStream_t st);

}  // namespace smax
#endif
