/*
  smax_ring.cuh -- device-side data structures of the ring kernel (smax_ring.cu), the scan
  kernel for indexes with sparse survivors, as the device manager (smax_device.cu) sees them.
  The unit kernel for dense indexes lives in smax_scan.cu / smax_kernels.cuh.  sm_100a only.
*/
#ifndef SMAX_RING_CUH
#define SMAX_RING_CUH

#include <cstdint>
#include <cuda_runtime.h>
#include "smax.h"

namespace smax_ring {

// ---- geometry of the scan kernel -----------------------------------------
constexpr int kThreads   = 256;               // consumer threads per CTA (+ one producer warp)
constexpr int kMinBlocks = 2;                 // resident CTAs per SM
constexpr int kItems     = 4;                 // 16-byte chunks per thread per tile
constexpr int kChunk     = 16;                // bytes per 128-bit shared-memory load
constexpr int kTileBytes = kThreads * kItems * kChunk;   // 16 KiB of lcptab per tile
constexpr int kHalo      = 16;                // table bytes staged either side of a tile
constexpr int kStages    = 2;                 // ring depth in (lcp + bwt) pairs: 2 * kStages buffers
constexpr int kLlvSlot   = 2048;              // .llv records of a tile staged in shared memory
constexpr int kWarpList  = 128;               // filter hits a warp collects before it works on them
constexpr int kLogCap    = 896;               // survivors a CTA collects before it writes them out
constexpr int kMaxGen    = 32;                // generations resolved per batch
constexpr int kMaxLeft   = 8;                 // peer shards a plateau may walk into
constexpr int kLlvBucketShift = 12;           // .llv directory: one entry per 4096 lcp entries

// tile status of the decoupled look-back: a 16-byte pair (record count,
// position count), each word [63:35] epoch, [34:33] state, [32:0] value.  The
// epoch makes a memset between scans unnecessary.
constexpr uint64_t kStateInvalid = 0, kStateAggregate = 1, kStatePrefix = 2;
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
  kResSlowTiles = 8,    // tiles redone by the slow path (their survivors did not fit the log)
  kResFlushes = 9,      // log flushes of all CTAs
  kResSlots = 12
};

struct ScanParams
{
  TableView own;
  TableView left[kMaxLeft];   // in shard order; left[nleft-1] is the nearest neighbour
  int nleft;
  int policy;
  int sufbytes;               // 8 or 4
  int debug;                  // tuning probes only (tools/probe_scan.py): 1 = skip look-back, 2 = skip K1 tail
  uint32_t epoch;
  uint64_t g_lo, g_hi;        // plateau ENDS in [g_lo, g_hi) belong to this shard
  uint64_t minlength;
  uint32_t mb;                // min(minlength, 255): byte threshold of the filter
  uint32_t ntiles;
  smax_record *recs;
  uint64_t rec_capacity;
  uint64_t *positions;        // null: do not gather positions
  uint64_t pos_capacity;
  uint64_t *status;           // 2 * ntiles look-back words (16-byte pairs)
  uint32_t *ctrl;             // [0] ticket, [1] finished CTAs
  uint64_t *peer_counts[SMAX_MAX_PEERS];   // count arrays of all shards (one-sided exchange), or none
  int npeers, my_rank;
  uint64_t exchange_tag;      // < 2^24; stored above the count
  uint64_t *result;           // kResSlots words of this scan
  uint64_t *result_next;      // the other block, zeroed by the last CTA for the next scan
};

// launchers (smax_ring.cu)
cudaError_t launch_scan(const ScanParams &p, bool stats, int grid, cudaStream_t st);
int scan_blocks_per_sm(bool stats);

}  // namespace smax_ring
#endif
