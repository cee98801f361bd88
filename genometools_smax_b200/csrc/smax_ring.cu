/*
  smax_ring.cu -- the ring kernel: the scan kernel for indexes with SPARSE survivors
  (few supermaximal repeats per tile: config C2).  The unit kernel of smax_scan.cu takes the
  dense ones; the device manager picks one per scan (smax_device.cu: pick_kernel).  ONE fused pass over the lcptab replaces the reference's stack sweep
  (/root/reference/src/match/esa-bottomup.c:116-273) and its per-node
  left-character bookkeeping (/root/reference/src/match/esa-maxpairs.c:181-360).

  k_scan is a persistent, warp-specialised kernel (cooperative launch, 2 CTAs
  per SM, each 8 consumer warps + 1 producer warp).  CTA b works on the 16 KiB
  lcptab tiles b, b + grid, b + 2 grid, ...

    feed  The producer warp describes each tile, arms its `ready` mbarrier and
          starts TMA bulk copies (cp.async.bulk.shared::cluster.global,
          complete_tx) into a ring of four buffers -- one buffer (lcp) per tile
          in sparse regions of the index, two (lcp + bwt) where plateau ends are
          frequent -- each with a 16-byte halo either side, plus one slot for the
          tile's .llv records that is refilled behind the small-value pass of
          the tile before.  No register staging, no per-thread loads of table
          bytes, and no CTA-wide barrier per tile: consumer warps arrive on the
          tile's `done` mbarrier and run ahead into the tiles in flight.
    K1    plateau detection, flat and bit-parallel (smax_swar.h).  Large values
          (byte 255) first, resolved in place in .llv RECORD space out of the
          staged slot (the tile's records are found through a per-4096-entry
          directory, no global rank): a record ends a plateau iff its right
          neighbour is no consecutive record with a value >= its own and its run
          is entered from a smaller value.  Small values: every thread filters
          its four 16-byte chunks for a byte >= minlength; the hits, compacted
          per warp with a ballot, are classified with SWAR byte arithmetic: ends
          of runs that fall to a smaller value, entered from a smaller value 1, 2
          or 3 entries back (SA width 2, 3, 4 -- 99.9 % of all plateaus).  Only
          runs of >= 4 equal values are walked (shared memory first, then global
          memory / the left neighbour shard).
    K2    left-distinctness, bit-parallel on the staged bwt words for widths
          <= 4 (pairwise byte compares; specials (>= 254) never collide under
          the GenomeTools convention, esa-maxpairs.c:24-31), a 256-bit
          alphabet mask otherwise.  Tiles of a sparse region read the few
          bwt bytes they need straight from global memory instead.
    K3    order-preserving compaction + emit.  Survivors join a log in shared
          memory; tile totals (record count, position count) are published as
          epoch-tagged 16-byte status pairs (no memset between scans) and
          exchanged generation-wise (no chain of dependent look-backs) when the
          log is written: records in suffix-array order, the occurrence
          positions suf[lb..lb+width) gathered right behind them.  A tile with
          more survivors than the log could take is redone by slow_tile.

*/
#include "smax_ring.cuh"
#include "smax_swar.h"

namespace smax_ring {

// ------------------------------------------------------------------ utils
__device__ __forceinline__ void ld_pair(const uint64_t *p, uint64_t &a, uint64_t &b)
{
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

__device__ __forceinline__ void st_pair(uint64_t *p, uint64_t a, uint64_t b)
{
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}

// streaming 128-bit load of .llv records: read-only path, do not keep in L1
__device__ __forceinline__ smax_llv ld_llv(const smax_llv *p)
{
  smax_llv r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];"
               : "=l"(r.position), "=l"(r.value) : "l"(p));
  return r;
}

__device__ __forceinline__ uint64_t pack_status(uint32_t epoch, uint64_t state, uint64_t value)
{
  return ((uint64_t) epoch << (kValueBits + 2)) | (state << kValueBits) | (value & kValueMask);
}

// ------------------------------------------------ TMA bulk copy + mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
  return (uint32_t) __cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do
  {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                 "selp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity), "r"(0x989680u) : "memory");   // sleep in hardware, not in a loop
  } while (!done);
}

// global -> shared bulk copy of `bytes` (multiple of 16, both sides 16-byte
// aligned), completion counted on `bar`
__device__ __forceinline__ void tma_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ------------------------------------------------------- table accessors
__device__ __forceinline__ const TableView *view_for(const ScanParams &P, uint64_t q)
{
  if (q >= P.own.a_lo)
    return &P.own;
  for (int k = P.nleft - 1; k >= 0; k--)
    if (q >= P.left[k].a_lo && q < P.left[k].a_hi)
      return &P.left[k];
  return nullptr;
}

// index of the .llv record with position i (the reference finds it with a
// binary search over the whole table, sarr-def.h:128-160; here the directory
// narrows it to one 4096-entry bucket)
__device__ __forceinline__ bool llv_find(const TableView &tv, uint64_t i, uint64_t &k)
{
  const uint64_t b = (i - tv.a_lo) >> kLlvBucketShift;
  uint64_t lo = tv.llvdir[b], hi = tv.llvdir[b + 1];
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (tv.llv[mid].position < i) lo = mid + 1; else hi = mid;
  }
  k = lo;
  return lo < tv.nllv && tv.llv[lo].position == i;
}

// resolved lcp value at an arbitrary index (slow, fully general; used only
// when a run leaves the shard's own arrays)
// Inconsistent tables (a plateau leaves every resident view, a 255 byte has no
// .llv record) are reported through the result block; the scan then fails on
// the host.  Such a value reads as "larger than everything" so that walks stop.
constexpr uint64_t kBadValue = ~0ull;

__device__ __noinline__ uint64_t value_at(const ScanParams &P, uint64_t q)
{
  const TableView *tv = view_for(P, q);
  if (tv == nullptr) { P.result[kResError] = 6; return kBadValue; }
  const uint32_t b = tv->lcp[q - tv->a_lo];
  if (b < 255)
    return b;
  uint64_t k;
  if (!llv_find(*tv, q, k)) { P.result[kResError] = 1; return kBadValue; }
  return tv->llv[k].value;
}

__device__ __noinline__ uint32_t byte_at_left(const ScanParams &P, uint64_t q, bool want_bwt)
{
  const TableView *tv = view_for(P, q);
  if (tv == nullptr) { P.result[kResError] = 6; return 255; }
  return want_bwt ? tv->bwt[q - tv->a_lo] : tv->lcp[q - tv->a_lo];
}

// K2 in full generality: are the left characters bwt[lb..e] pairwise distinct?
__device__ __noinline__ bool left_distinct(const ScanParams &P, uint64_t lb, uint64_t e)
{
  const uint64_t a_lo = P.own.a_lo;
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  uint64_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
  for (uint64_t q = lb; q <= e; q++)
  {
    const uint32_t c = (q >= a_lo) ? (uint32_t) P.own.bwt[q - a_lo] : byte_at_left(P, q, true);
    if (gt_policy && c >= 254)
      continue;
    const uint64_t bit = 1ull << (c & 63);
    uint64_t hit;
    switch (c >> 6)
    {
      case 0: hit = m0 & bit; m0 |= bit; break;
      case 1: hit = m1 & bit; m1 |= bit; break;
      case 2: hit = m2 & bit; m2 |= bit; break;
      default: hit = m3 & bit; m3 |= bit; break;
    }
    if (hit)
      return false;
  }
  return true;
}

// --------------------------------------------------- shared memory layout
constexpr int kStageBytes = kHalo + kTileBytes + kHalo;   // table bytes of one ring slot

// one survivor waiting for the prefix of its tile (12 bytes, structure of arrays):
// repeat length (kLongValue: does not fit, read it again from the tables),
// SA width, end offset in its tile | which of this CTA's tiles << 16
constexpr uint32_t kLongValue = 0xffffffffu;

// the .llv records of a tile: [k0, k1) lie in the tile, the slot holds
// [kfirst, kfirst + nrec) (one neighbour either side, capacity permitting)
struct LlvMeta
{
  uint32_t k0, k1, kfirst, nrec;
};

// what the producer warp tells the consumer warps about the tile in a ring slot
// (written before the slot's mbarrier is armed, read after it has completed)
struct TileDesc
{
  LlvMeta llv;
  uint32_t flags;                // kDesc*
  uint8_t lbuf, bbuf, pad[2];    // ring buffers that hold its lcp / bwt bytes
};
constexpr uint32_t kDescBwt = 1;         // the bwt slot is being filled too
constexpr uint32_t kDescFlush = 2;       // write the survivor log out before this tile

constexpr int kConsumers = kThreads;             // threads that scan (8 warps)
constexpr int kBlockThreads = kThreads + 32;     // + the producer warp
constexpr int kMaxDrop = 32;                     // tiles waiting for their redo at any time
constexpr int kSegs = 64;                        // see ScanSmem::seg_lo
constexpr int kBufs = 2 * kStages;               // ring buffers: a tile takes one (lcp) or two (lcp + bwt)
constexpr int kInFlight = kBufs;                 // tiles described at any time

struct ScanSmem
{
  // ring slots: slot[kHalo + i] = table[tile_lo + i], i in [-kHalo, kTileBytes + kHalo)
  alignas(128) uint8_t buf[kBufs][kStageBytes];
  alignas(16) smax_llv llv[kLlvSlot + 2];    // one slot: filled while the next tile's small values are scanned
  uint32_t log_v[kLogCap], log_w[kLogCap], log_t[kLogCap];
  uint16_t wlist[kThreads / 32][kWarpList];   // per warp: chunks that passed the filter
  // per generation of the batch being resolved: totals, then prefix of this CTA's tile
  unsigned long long gtot_c[kMaxGen], gtot_w[kMaxGen], gexc_c[kMaxGen], gexc_w[kMaxGen];
  TileDesc desc[kInFlight];
  unsigned long long tile_w[kInFlight];  // per tile in flight: position count,
  uint32_t tile_c[kInFlight];            //   survivors,
  uint32_t tile_met[kInFlight];          //   warps that met a candidate plateau,
  uint32_t tile_drop[kInFlight];         //   != 0: the tile lost survivors (tag + 1; listed below)
  uint32_t drop_tag[kMaxDrop];           // tiles that lost survivors and wait for their redo (tag + 1)
  uint32_t ndrop;
  unsigned long long run_c, run_w;   // records / positions of all resolved generations
  uint32_t log_n;                // survivors appended (> kLogCap: dropped, their tile is redone)
  uint32_t seg_lo[kSegs], seg_hi[kSegs];   // log index range of the entries of tile tag % kSegs
  alignas(8) uint64_t ready[kInFlight];  // mbarriers: the table bytes of the tile have landed
  alignas(8) uint64_t vfull;             //   the .llv records have landed
  alignas(8) uint64_t vdone[kInFlight];  //   all consumer warps are through with the tile's .llv records
  alignas(8) uint64_t done[kInFlight];   //   all consumer warps are through with the tile
};

// barrier of the consumer warps only (the producer warp never joins)
__device__ __forceinline__ void consumer_sync()
{
  asm volatile("bar.sync 1, %0;" :: "n"(kConsumers) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// K3, first half: a survivor joins the log of its CTA.
__device__ __noinline__ void emit_survivor(const ScanParams &P, ScanSmem &sm, uint32_t it16, int par,
                                           uint32_t o, uint64_t v, uint64_t width)
{
  const uint32_t slot = atomicAdd(&sm.log_n, 1u);
  atomicAdd(&sm.tile_c[par], 1u);
  atomicAdd(&sm.tile_w[par], (unsigned long long) width);
  if (width >> 32)
    P.result[kResError] = 4;              // wider than a shard can be
  if (slot < (uint32_t) kLogCap)
  {
    sm.log_v[slot] = v < (uint64_t) kLongValue ? (uint32_t) v : kLongValue;
    sm.log_w[slot] = (uint32_t) width;
    sm.log_t[slot] = o | (it16 << 16);
    atomicMin(&sm.seg_lo[it16 % kSegs], slot);     // where the tile's entries sit in the log
    atomicMax(&sm.seg_hi[it16 % kSegs], slot);
  } else if (atomicExch(&sm.tile_drop[par], it16 + 1) != it16 + 1)
  {
    // the first survivor of this tile that did not fit: list the tile for its redo
    const uint32_t k = atomicAdd(&sm.ndrop, 1u);
    if (k < (uint32_t) kMaxDrop)
      sm.drop_tag[k] = it16 + 1;
    else
      P.result[kResError] = 5;
  }
}

// K2 for one candidate plateau [lb, e] from global memory: left characters
// pairwise distinct?  Short plateaus inside the shard's own arrays load all
// their bwt bytes at once.
__device__ __noinline__ bool candidate_survives(const ScanParams &P, uint64_t lb, uint64_t e,
                                                   uint64_t width)
{
  const uint64_t a_lo = P.own.a_lo;
  if (width <= 4 && lb >= a_lo)
  {
    const uint8_t *bp = P.own.bwt + (lb - a_lo);
    uint32_t c[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
      c[j] = (uint64_t) j < width ? (uint32_t) bp[j] : 0x100u + j;   // absent: unique
    const bool gt_policy = (P.policy == SMAX_POLICY_GT);
    bool dup = false;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = i + 1; j < 4; j++)
        dup |= (c[i] == c[j]) && !(gt_policy && c[i] >= 254);
    return !dup;
  }
  return left_distinct(P, lb, e);
}

// The run of small value b that ends at e is known to reach back to s: walk
// further left in global memory with 128-bit compares.  Returns the SA width of
// the local-maximum plateau ending at e, or 0 if the run is entered from a
// larger value.
__device__ __noinline__ uint64_t small_plateau_width_global(const ScanParams &P, uint64_t e,
                                                         uint64_t s, uint32_t b)
{
  const uint8_t *lcp = P.own.lcp;
  const uint64_t a_lo = P.own.a_lo;
  const uint32_t v4 = b * 0x01010101u;
  for (;;)
  {
    if (s == 0)
      break;                              // start of the table
    const uint64_t q = s - 1;
    uint32_t pb;
    if (q >= a_lo)
    {
      const uint64_t so = s - a_lo;
      if ((so & 15) == 0 && so >= 16)
      {
        const uint4 w = *reinterpret_cast<const uint4 *>(lcp + so - 16);
        if (((w.x ^ v4) | (w.y ^ v4) | (w.z ^ v4) | (w.w ^ v4)) == 0)
        {
          s -= 16;
          continue;
        }
      }
      pb = lcp[q - a_lo];
    } else
      pb = byte_at_left(P, q, false);     // 255 on error: stops the walk
    if (pb == b) { s = q; continue; }
    if (pb > b)
      return 0;
    break;
  }
  return e - s + 2;
}

// The same walk inside the resident slot: st[kHalo + i] = lcp[tile_lo + i] for
// i in [-kHalo, kTileBytes + kHalo); the run is known to cover [o - known, o].
__device__ __forceinline__ uint64_t small_plateau_width_staged(const ScanParams &P,
                                                               const uint8_t *st, uint64_t tile_lo,
                                                               uint32_t o, int known, uint32_t b)
{
  int i = (int) o - known;               // run start candidate, tile offset (may go down to -kHalo)
  for (;;)
  {
    if (i == -kHalo)                     // staged range exhausted: continue in global memory
      return small_plateau_width_global(P, tile_lo + o, tile_lo + i, b);
    if ((i & 15) == 0)                   // 16 entries per step while they all equal b
    {
      const uint4 w = *reinterpret_cast<const uint4 *>(st + kHalo + i - 16);
      const uint32_t v4 = b * 0x01010101u;
      if (((w.x ^ v4) | (w.y ^ v4) | (w.z ^ v4) | (w.w ^ v4)) == 0)
      {
        i -= 16;
        continue;
      }
    }
    const uint32_t pb = st[kHalo + i - 1];
    if (pb == b) { i--; continue; }
    if (pb > b)
      return 0;
    break;
  }
  return (uint64_t) ((int) o - i) + 2;
}

// K2 from the staged bwt bytes when the plateau lies inside the staged range
__device__ __forceinline__ bool candidate_survives_staged(const ScanParams &P, const uint8_t *sb,
                                                          uint64_t tile_lo, uint32_t o,
                                                          uint64_t width)
{
  if (width == 2)                        // by far the most common: two left characters
  {
    const uint32_t c0 = sb[kHalo + (int) o - 1], c1 = sb[kHalo + o];
    return c0 != c1 || (P.policy == SMAX_POLICY_GT && c0 >= 254);
  }
  if (width <= 4 && (int) o + 1 - (int) width >= -kHalo)
  {
    const uint8_t *bp = sb + kHalo + (int) o + 1 - (int) width;
    uint32_t c[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
      c[j] = (uint64_t) j < width ? (uint32_t) bp[j] : 0x100u + j;   // absent: unique
    const bool gt_policy = (P.policy == SMAX_POLICY_GT);
    bool dup = false;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = i + 1; j < 4; j++)
        dup |= (c[i] == c[j]) && !(gt_policy && c[i] >= 254);
    return !dup;
  }
  const uint64_t e = tile_lo + o;
  return candidate_survives(P, e + 1 - width, e, width);
}

// SA width of the local-maximum plateau of large values ending at record k, or
// 0 if the run is entered from a larger value.  Runs are walked in record space;
// rec(k) reads record k (from the staged slot where it holds it).  The walk of
// the usual tile (llv_walk below is the interruptible one of the wide-repeat path).
template <typename R>
__device__ __forceinline__ uint64_t llv_plateau_width(const ScanParams &P, R rec, uint64_t k,
                                                      uint64_t p, uint64_t v)
{
  uint64_t s = p, kk = k;
  for (;;)
  {
    if (s == 0)
      break;
    const uint64_t q = s - 1;
    uint64_t pv;
    if (q >= P.own.a_lo)
    {
      if (kk == 0)
        break;                            // no record at q: a small value, rise
      const smax_llv pr = rec((uint32_t) (kk - 1));
      if (pr.position != q)
        break;
      pv = pr.value;
      kk--;
    } else
      pv = value_at(P, q);                // kBadValue on error: stops the walk
    if (pv == v) { s = q; continue; }
    if (pv > v)
      return 0;
    break;
  }
  return p - s + 2;
}

// Walk left over the run of equal large values that ends at position p (value
// v): the run is known to cover [s, p], s being the position of record kk.
// Returns true when the walk is over -- width = SA width of the local-maximum
// plateau, or 0 if the run is entered from a larger value -- and false when the
// run is still going after `limit` steps inside the shard's own records (kk, s
// then describe how far it got).  rec(k) reads record k (from the staged slot
// where it holds it).
template <typename R>
__device__ __forceinline__ bool llv_walk(const ScanParams &P, R rec, uint32_t &kk, uint64_t &s,
                                         uint64_t p, uint64_t v, int limit, uint64_t &width)
{
  for (int step = 0;; step++)
  {
    if (s == 0)
      break;
    const uint64_t q = s - 1;
    uint64_t pv;
    const bool own = q >= P.own.a_lo;
    if (own)
    {
      if (kk == 0)
        break;                            // no record at q: a small value, rise
      const smax_llv pr = rec(kk - 1);
      if (pr.position != q)
        break;
      pv = pr.value;
    } else
      pv = value_at(P, q);                // kBadValue on error: stops the walk
    if (pv == v)
    {
      if (own && step >= limit)
        return false;
      s = q;
      if (own) kk--;
      continue;
    }
    if (pv > v) { width = 0; return true; }
    break;
  }
  width = p - s + 2;
  return true;
}

// The lanes in `todo` each follow a run of equal large values that is still
// going after their own few steps: the whole warp walks those runs, one after
// the other, 32 records per step (lane j looks at record wk - 1 - j).  Stops
// where the run ends or leaves the shard's own records; the lane finishes the
// walk from there.
__device__ __noinline__ void walk_runs_together(const ScanParams &P, const smax_llv *sv,
                                                const LlvMeta &M, uint32_t todo, uint32_t &wk,
                                                uint64_t &ws, uint64_t value)
{
  const int lane = threadIdx.x & 31;
  const uint64_t a_lo = P.own.a_lo;
  while (todo)
  {
    const int src = __ffs(todo) - 1;
    todo &= todo - 1;
    uint32_t bk = __shfl_sync(0xffffffffu, wk, src);
    unsigned long long bs = __shfl_sync(0xffffffffu, (unsigned long long) ws, src);
    const unsigned long long bv = __shfl_sync(0xffffffffu, (unsigned long long) value, src);
    for (;;)
    {
      bool same = false;
      if (bk > (uint32_t) lane && bs > (uint64_t) lane && bs - 1 - lane >= a_lo)
      {
        const uint32_t k = bk - 1 - lane, i = k - M.kfirst;
        const smax_llv q = i < M.nrec ? sv[i] : ld_llv(&P.own.llv[k]);
        same = q.position == bs - 1 - lane && q.value == bv;
      }
      const uint32_t votes = __ballot_sync(0xffffffffu, same);
      const uint32_t run = votes == 0xffffffffu ? 32u : (uint32_t) (__ffs(~votes) - 1);
      bk -= run; bs -= run;
      if (run < 32)
        break;
    }
    if (lane == src) { wk = bk; ws = bs; }
  }
}

// per-pass context of one tile
struct PassCtx
{
  const uint8_t *sl;       // resident lcp slot
  const uint8_t *sb;       // resident bwt slot, or nullptr (sparse region)
  const smax_llv *sv;      // resident .llv slot
  uint64_t tile_lo;        // global lcp index of tile offset 0
  uint32_t it16;           // tag of the tile in the survivor log
  int par;                 // ring slot of the tile (its counters)
  uint32_t vparity;        // parity of the .llv slot barrier to wait for
  LlvMeta llv;             // the tile's .llv records
};

// K2 + emit for one local-maximum plateau [e + 1 - width, e] of value v.
template <bool STATS>
__device__ __forceinline__ void test_and_emit(const ScanParams &P, ScanSmem &sm, const PassCtx &C,
                                              uint32_t o, uint64_t v, uint64_t width, uint64_t *stat)
{
  if (STATS) { stat[0]++; stat[1] += width; }
  if (P.debug & 16)
    return;
  const uint64_t e = C.tile_lo + o;
  const bool ok = C.sb != nullptr ? candidate_survives_staged(P, C.sb, C.tile_lo, o, width)
                                  : candidate_survives(P, e + 1 - width, e, width);
  if (ok)
  {
    if (STATS) stat[3] += width;
    if ((P.debug & 32) == 0)
      emit_survivor(P, sm, C.it16, C.par, o, v, width);
  }
}

// bit 7 of byte j of m[i] -> bit 4 i + j
__device__ __forceinline__ uint32_t pack_ends16(const uint32_t m[4])
{
  uint32_t r = 0;
#pragma unroll
  for (int i = 0; i < 4; i++)
    r |= ((((m[i] >> 7) & 0x01010101u) * 0x01020408u) >> 24) << (4 * i);
  return r;
}

// the set bits of a mask word are the plateau ends at tile offsets o0 + byte
template <typename F>
__device__ __forceinline__ void for_each_end(uint32_t m, uint32_t o0, F f)
{
  while (m)
  {
    const uint32_t bit = __ffs(m) - 1;    // 7, 15, 23 or 31
    m &= m - 1;
    f(o0 + (bit >> 3));
  }
}

// The large-value half of a pass in the kernel variant for indexes with a large
// share of large values -- wide repeats (many exact copies of a long element).  As in the usual
// tile a record ends a plateau iff its right neighbour is no consecutive record
// with a value >= its own and its run is entered from a smaller value, but runs
// of EQUAL values are long here: a lane walks its run for a few steps; what is
// still going then is walked by the whole warp, 32 records per step.  kThreads
// records per round, all lanes in step.  Returns whether this thread met a plateau.
template <bool STATS>
__device__ __forceinline__ int llv_pass_wide(const ScanParams &P, ScanSmem &sm, const PassCtx &C,
                                             uint64_t *stat_out)
{
  uint64_t stat[4] = {0, 0, 0, 0};
  const int tid = threadIdx.x;
  const LlvMeta M = C.llv;
  const smax_llv *llv = P.own.llv;
  const uint64_t nllv = P.own.nllv, a_lo = P.own.a_lo, tile_lo = C.tile_lo;
  const uint64_t lo = max(tile_lo, P.g_lo);
  const uint64_t hi = min(tile_lo + (uint64_t) kTileBytes, P.g_hi);
  auto rec = [&](uint32_t k) -> smax_llv
  {
    const uint32_t i = k - M.kfirst;
    return i < M.nrec ? C.sv[i] : ld_llv(&llv[k]);
  };
  int met = 0;
#pragma unroll 1
  for (uint32_t kb = M.k0; kb < M.k1; kb += kThreads)
  {
    const uint32_t k = kb + tid;
    smax_llv r;
    r.position = 0; r.value = 0;
    uint64_t width = 0;                    // 0: no plateau ends here
    bool walking = false;
    uint64_t ws = 0;                       // the run is known to cover [ws, r.position],
    uint32_t wk = 0;                       //   ws being the position of record wk
    if (k < M.k1)
    {
      r = rec(k);
      if (STATS) stat[2]++;
      bool end = r.position >= lo && r.position < hi && r.value >= P.minlength;
      if (end && (uint64_t) k + 1 < nllv)
      {
        const smax_llv nx = rec(k + 1);
        end = !(nx.position == r.position + 1 && nx.value >= r.value);   // else the run goes on
      }
      if (end)
      {
        width = 2;                         // previous entry is a smaller value
        ws = r.position; wk = k;
        bool walk = r.position == a_lo && a_lo > 0;   // shard edge
        if (!walk && k > 0)
        {
          const smax_llv pr = rec(k - 1);
          if (pr.position == r.position - 1)
          {
            if (pr.value > r.value)
              width = 0;                   // entered from a larger value
            walk = pr.value == r.value;    // run of equal values
          }
        }
        if (walk)
          walking = !llv_walk(P, rec, wk, ws, r.position, r.value, 6, width);
      }
    }
    const uint32_t todo = __ballot_sync(0xffffffffu, walking);
    if (todo)
      walk_runs_together(P, C.sv, M, todo, wk, ws, r.value);
    if (walking)                           // what comes before the run: a step or two more
      llv_walk(P, rec, wk, ws, r.position, r.value, 1 << 30, width);
    if (width != 0)
    {
      met = 1;
      test_and_emit<STATS>(P, sm, C, (uint32_t) (r.position - tile_lo), r.value, width, stat);
    }
  }
  if (STATS)
    for (int k = 0; k < 4; k++)
      stat_out[k] += stat[k];
  return met;
}

// One detection pass over the resident tile (K1 + K2 + logging of survivors).
// Both halves work in two phases per warp so that the expensive part runs on
// full warps: phase A is a cheap filter over everything (does the chunk hold a
// byte >= the threshold / does the .llv record end a run), whose hits are
// compacted into a small per-warp list with a ballot; phase B takes the list
// entries lane by lane.
// Executed by each consumer warp on its own (no CTA barrier): the warp counts
// itself into tile_met if it met a candidate plateau (the density signal that
// decides whether the next tiles prefetch their bwt) and arrives on the slot's
// `done` barrier.
template <bool STATS, bool WIDE>
__device__ __forceinline__ void tile_pass(const ScanParams &P, ScanSmem &sm, const PassCtx &C)
{
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint16_t *list = sm.wlist[tid >> 5];
  const uint64_t tile_lo = C.tile_lo;
  uint32_t kadd; int himode;
  smax_ge_consts(P.mb, &kadd, &himode);
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  // ends at or beyond g_hi belong to the next shard
  const uint32_t valid = (uint32_t) min((uint64_t) kTileBytes, P.g_hi - tile_lo);
  uint64_t stat[4] = {0, 0, 0, 0};
  int met = 0;

  // ---- large values: the tile's .llv records out of the staged slot (the few
  // beyond its capacity come from global memory).  A record ends a plateau iff
  // its right neighbour is no consecutive record with a value >= its own.
  const LlvMeta M = C.llv;
  if (M.k0 < M.k1)
  {
    mbar_wait(&sm.vfull, C.vparity);       // issued when the previous tile was done
    const smax_llv *llv = P.own.llv;
    const uint64_t nllv = P.own.nllv;
    const uint64_t a_lo = P.own.a_lo;
    const uint64_t lo = max(tile_lo, P.g_lo);
    const uint64_t hi = min(tile_lo + (uint64_t) kTileBytes, P.g_hi);
    auto rec = [&](uint32_t k) -> smax_llv
    {
      const uint32_t i = k - M.kfirst;
      return i < M.nrec ? C.sv[i] : ld_llv(&llv[k]);
    };
    if (!WIDE)
    {
      // the usual tile: every lane on its own (runs of equal large values are short)
      // the record k ends a run of large values >= minlength: plateau? K2, log
      auto process_end = [&](const uint32_t k)
      {
        const smax_llv r = rec(k);
        uint64_t width = 2;                    // previous entry is a smaller value
        bool walk = r.position == a_lo && a_lo > 0;   // shard edge
        if (!walk && k > 0)
        {
          const smax_llv pr = rec(k - 1);
          if (pr.position == r.position - 1)
          {
            if (pr.value > r.value)
              width = 0;                       // entered from a larger value
            walk = pr.value == r.value;        // run of equal values
          }
        }
        if (walk)
          width = llv_plateau_width(P, rec, k, r.position, r.value);
        if (width != 0)
        {
          met = 1;
          test_and_emit<STATS>(P, sm, C, (uint32_t) (r.position - tile_lo), r.value, width, stat);
        }
      };
      auto is_end = [&](const uint32_t k) -> bool
      {
        const smax_llv r = rec(k);
        if (STATS) stat[2]++;
        if (r.position < lo || r.position >= hi || r.value < P.minlength)
          return false;
        if ((uint64_t) k + 1 < nllv)
        {
          const smax_llv nx = rec(k + 1);
          if (nx.position == r.position + 1 && nx.value >= r.value)
            return false;                      // the run goes on
        }
        return true;
      };
#pragma unroll 1
      for (uint32_t k = M.k0 + tid; k < M.k1; k += kThreads)
        if (is_end(k))
          process_end(k);
    } else
      met |= llv_pass_wide<STATS>(P, sm, C, STATS ? stat : nullptr);
  }
  // the .llv slot may be refilled as soon as every warp is past this point
  __syncwarp();
  if (lane == 0)
    mbar_arrive(&sm.vdone[C.par]);

  // ---- small values
  if (!(P.debug & 2))
  {
    // K1 + K2 + logging for the chunk at tile offset o0
    auto process_chunk = [&](const uint32_t o0)
    {
      const uint8_t *lp = C.sl + kHalo + o0;
      uint32_t w[6];
      {
        const uint4 x = *reinterpret_cast<const uint4 *>(lp);
        w[1] = x.x; w[2] = x.y; w[3] = x.z; w[4] = x.w;
      }
      w[0] = *reinterpret_cast<const uint32_t *>(lp - 4);
      w[5] = *reinterpret_cast<const uint32_t *>(lp + 16);
      smax_chunk_k1 k;
      if (!smax_chunk_detect(w, kadd, himode, &k))
        return;
      if (o0 + kChunk > valid)              // the shard (or the piece) ends inside this chunk
      {
        const uint32_t keep = valid - o0;   // 1..15 bytes
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
          const uint32_t m = (uint32_t) (4 * j + 4) <= keep ? 0xffffffffu
                             : ((uint32_t) (4 * j) >= keep ? 0u : (0xffffffffu >> (8 * (4 * j + 4 - keep))));
          k.c2[j] &= m; k.c3[j] &= m; k.c4[j] &= m; k.lng[j] &= m;
        }
        k.any_cand = k.c2[0] | k.c2[1] | k.c2[2] | k.c2[3] | k.c3[0] | k.c3[1] | k.c3[2] | k.c3[3] |
                     k.c4[0] | k.c4[1] | k.c4[2] | k.c4[3];
        k.any_long = k.lng[0] | k.lng[1] | k.lng[2] | k.lng[3];
      }
      met |= (k.any_cand | k.any_long) != 0;
      if (k.any_long)
      {
        // runs of >= 4 equal values: walk them
        uint32_t u = pack_ends16(k.lng);
        while (u)
        {
          const uint32_t o = o0 + (__ffs(u) - 1);
          u &= u - 1;
          const uint32_t b = lp[o - o0];
          const uint64_t width = small_plateau_width_staged(P, C.sl, tile_lo, o, 3, b);
          if (width != 0)
            test_and_emit<STATS>(P, sm, C, o, b, width, stat);
        }
      }
      if (k.any_cand)
      {
        if (STATS)
        {
#pragma unroll
          for (int j = 0; j < 4; j++)
          {
            const uint32_t n2 = __popc(k.c2[j]), n3 = __popc(k.c3[j]), n4 = __popc(k.c4[j]);
            stat[0] += n2 + n3 + n4;
            stat[1] += 2 * n2 + 3 * n3 + 4 * n4;
          }
        }
        if (P.debug & 16)
          return;
        // K2: bit-parallel on the staged bwt words, or -- in a sparse region --
        // per candidate with the few left characters straight from global memory
        const bool staged = C.sb != nullptr;
        if (staged)
        {
          const uint8_t *bp = C.sb + kHalo + o0;
          uint32_t b[5];
          const uint4 y = *reinterpret_cast<const uint4 *>(bp);
          b[0] = *reinterpret_cast<const uint32_t *>(bp - 4);
          b[1] = y.x; b[2] = y.y; b[3] = y.z; b[4] = y.w;
          if (!smax_chunk_distinct(b, gt_policy, &k))
            return;
        }
        // the rest is rare in the staged case: one loop over the ends, width from the masks
        const uint32_t p3 = pack_ends16(k.c3), p4 = pack_ends16(k.c4);
        uint32_t u = pack_ends16(k.c2) | p3 | p4;
        while (u)
        {
          const uint32_t bit = __ffs(u) - 1;
          u &= u - 1;
          const uint32_t o = o0 + bit;
          const uint64_t width = 2 + ((p3 >> bit) & 1u) + 2 * ((p4 >> bit) & 1u);
          if (!staged && !candidate_survives(P, tile_lo + o + 1 - width, tile_lo + o, width))
            continue;
          if (STATS) stat[3] += width;
          if (!(P.debug & 32))
            emit_survivor(P, sm, C.it16, C.par, o, lp[bit], width);
        }
      }
    };
    // phase A: which of the warp's 128 chunks hold a byte >= the threshold?
    uint32_t n = 0;
#pragma unroll
    for (int c = 0; c < kItems; c++)
    {
      const uint32_t ch = (uint32_t) (c * kThreads + tid), o0 = ch * kChunk;
      bool hit = false;
      if (o0 < valid)
      {
        const uint4 x = *reinterpret_cast<const uint4 *>(C.sl + kHalo + o0);
        hit = (smax_ge(x.x, kadd, himode) | smax_ge(x.y, kadd, himode) | smax_ge(x.z, kadd, himode) |
               smax_ge(x.w, kadd, himode)) != 0;
      }
      const uint32_t votes = __ballot_sync(0xffffffffu, hit);
      if (hit)
        list[n + __popc(votes & lt_mask)] = (uint16_t) ch;
      n += __popc(votes);
    }
    __syncwarp();
    // phase B: K1 + K2 on the listed chunks
#pragma unroll 1
    for (uint32_t i = lane; i < n; i += 32)
      process_chunk((uint32_t) list[i] * kChunk);
    __syncwarp();                          // the list is reused below
  }

  if (STATS)
  {
    if (stat[0]) atomicAdd((unsigned long long *) &P.result[kResStatCand], (unsigned long long) stat[0]);
    if (stat[1]) atomicAdd((unsigned long long *) &P.result[kResStatCandWidth], (unsigned long long) stat[1]);
    if (stat[2]) atomicAdd((unsigned long long *) &P.result[kResStatLlv], (unsigned long long) stat[2]);
    if (stat[3]) atomicAdd((unsigned long long *) &P.result[kResStatSurvWidth], (unsigned long long) stat[3]);
  }
  // this warp is through with the tile
  if (__any_sync(0xffffffffu, met) && lane == 0)
    atomicAdd(&sm.tile_met[C.par], 1u);
  __syncwarp();
  if (lane == 0)
    mbar_arrive(&sm.done[C.par]);
}

__device__ __forceinline__ uint64_t suf_at(const ScanParams &P, uint64_t i)
{
  const TableView *tv = view_for(P, i);
  if (tv == nullptr || tv->suf == nullptr) { P.result[kResError] = 6; return 0; }
  const uint64_t o = i - tv->a_lo;
  if (i >= tv->a_hi) { P.result[kResError] = 3; return 0; }
  return P.sufbytes == 8 ? reinterpret_cast<const uint64_t *>(tv->suf)[o]
                         : (uint64_t) reinterpret_cast<const uint32_t *>(tv->suf)[o];
}

// ------------------------------------------------- ordered prefix exchange
// Tiles are assigned round-robin: CTA b owns tiles b, b+grid, b+2*grid, ...
// ("generation" g = tiles [g*grid, (g+1)*grid)), and the whole grid is
// resident, so every generation is worked on by all CTAs at the same time.
// Each tile publishes its aggregate pair (records, positions) as soon as its
// detection pass is done.  Survivors wait in the log of their CTA; when the log
// is written out, the CTA reads the aggregates of all generations it has not
// resolved yet in one sweep (every thread a different tile: no chain of
// dependent look-backs), turns them into the prefix of its own tile per
// generation, and keeps the running totals.  A CTA without survivors never
// reads a single aggregate.
__device__ __forceinline__ void publish_aggregate(uint64_t *status, uint32_t tile, uint64_t agg_a,
                                                  uint64_t agg_b, uint32_t epoch)
{
  st_pair(&status[2 * (uint64_t) tile], pack_status(epoch, kStateAggregate, agg_a),
          pack_status(epoch, kStateAggregate, agg_b));
}

// K3, second half: write the log entries tagged [t0, t0 + nt) in suffix-array
// order; sm.gexc_c/w[t - t0] is the global prefix of the tile tagged t.  Rank and
// position offset of an entry within its tile come from a look at the log range
// that holds the tile's entries.  Entries of the tiles that lost survivors are
// skipped; those tiles are redone by slow_tile.
__device__ __forceinline__ void write_log(const ScanParams &P, ScanSmem &sm, uint32_t n, uint32_t t0,
                                          uint32_t nt, uint32_t it_of_t0, uint32_t me, uint32_t grid,
                                          uint32_t ndrop)
{
  const uint64_t base_off = P.g_lo - P.own.a_lo;
  for (uint32_t e = threadIdx.x; e < n; e += kConsumers)
  {
    const uint32_t tag = sm.log_t[e], off = tag & 0xffffu, mine = tag >> 16;
    const uint32_t t = mine - t0;
    bool skip = t >= nt;
    for (uint32_t k = 0; k < ndrop; k++)
      skip |= mine + 1 == sm.drop_tag[k];
    if (skip)
      continue;
    // same-tile entries with a smaller end offset: the tile's entries sit in
    // [seg_lo, seg_hi] of the log (interleaved with those of the tiles that were
    // in work at the same time); four entries per step, the loads are independent
    uint32_t rank = 0;
    uint64_t posoff = 0;
    auto look = [&](uint32_t i)
    {
      const uint32_t other = sm.log_t[i];
      if ((other >> 16) == mine && other < tag) { rank++; posoff += sm.log_w[i]; }
    };
    {
      uint32_t i = sm.seg_lo[mine % kSegs];
      const uint32_t last = min(sm.seg_hi[mine % kSegs], n - 1);
      for (; i + 3 <= last; i += 4)
      {
        look(i); look(i + 1); look(i + 2); look(i + 3);
      }
      for (; i <= last; i++)
        look(i);
    }
    const uint64_t dst = sm.gexc_c[t] + rank;
    const uint64_t po = sm.gexc_w[t] + posoff;
    const uint64_t tile = (uint64_t) me + (uint64_t) (it_of_t0 + t) * grid;
    const uint64_t end = P.own.a_lo + base_off + tile * kTileBytes + off;
    const uint64_t wd = sm.log_w[e];
    if (wd < 2 || wd > end + 1) { P.result[kResError] = 2; continue; }
    const uint64_t lb = end + 1 - wd;
    if (dst < P.rec_capacity)
    {
      smax_record r;
      r.len = sm.log_v[e] != kLongValue ? (uint64_t) sm.log_v[e] : value_at(P, end);
      r.lb = lb; r.width = wd;
      P.recs[dst] = r;
    } else
      P.result[kResOverflow] = 1;
    if (P.positions != nullptr)
    {
      if (po + wd > P.pos_capacity)
        P.result[kResOverflow] = 1;
      else if (wd <= 4 && lb >= P.own.a_lo && P.own.suf != nullptr)
      {
        // the common case: all (<= 4) scattered suftab reads in flight together
        const uint64_t o = lb - P.own.a_lo;
        uint64_t v[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
          v[k] = (uint64_t) k >= wd ? 0
                 : P.sufbytes == 8 ? reinterpret_cast<const uint64_t *>(P.own.suf)[o + k]
                                   : (uint64_t) reinterpret_cast<const uint32_t *>(P.own.suf)[o + k];
#pragma unroll
        for (int k = 0; k < 4; k++)
          if ((uint64_t) k < wd)
            P.positions[po + k] = v[k];
      } else
        for (uint64_t k = 0; k < wd; k++)
          P.positions[po + k] = suf_at(P, lb + k);
    }
  }
}

// A tile with more survivors than the log could take (only in indexes where a
// large share of all suffixes ends a supermaximal repeat) is redone here, by all
// consumer warps together, straight from the tables in global memory: one
// thread per entry, kConsumers entries per round, the survivors of a round
// written in order behind those of the rounds before.  No shared-memory staging,
// no log, no bit tricks: simple, and exercised by the parity tests (dense and
// alternating fuzz tables, minlength 1) next to the fast path.
__device__ __noinline__ void slow_tile(const ScanParams &P, ScanSmem &sm, uint64_t tile,
                                       uint64_t rec_base, uint64_t pos_base)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t tile_lo = P.g_lo + tile * kTileBytes;
  if (tid == 0)
    atomicAdd((unsigned long long *) &P.result[kResSlowTiles], 1ull);
  unsigned long long *scratch = reinterpret_cast<unsigned long long *>(&sm.wlist[0][0]);
  for (uint32_t r0 = 0; r0 < (uint32_t) kTileBytes; r0 += kConsumers)
  {
    const uint64_t e = tile_lo + r0 + tid;
    uint64_t v = 0, width = 0;
    if (e < P.g_hi)
    {
      const uint8_t *lcp = P.own.lcp;
      const uint32_t b = lcp[e - P.own.a_lo];
      if (b < 255)
      {
        // a small value ends a run iff the next byte is smaller (255 stands for a
        // larger value); the run is walked 16 entries per step
        v = b;
        if (v >= P.minlength && b > (uint32_t) lcp[e + 1 - P.own.a_lo])
          width = small_plateau_width_global(P, e, e, b);
      } else
      {
        // a large value: find its record once, then stay in record space
        uint64_t k;
        if (!llv_find(P.own, e, k))
          P.result[kResError] = 1;
        else
        {
          const smax_llv *llv = P.own.llv;
          auto rec = [&](uint32_t i) -> smax_llv { return llv[i]; };
          v = llv[k].value;
          bool end = v >= P.minlength;
          if (end && k + 1 < P.own.nllv)
          {
            const smax_llv nx = llv[k + 1];
            end = !(nx.position == e + 1 && nx.value >= v);
          }
          if (end)
          {
            uint32_t wk = (uint32_t) k;
            uint64_t ws = e;
            llv_walk(P, rec, wk, ws, e, v, 1 << 30, width);
          }
        }
      }
      if (width != 0 && !candidate_survives(P, e + 1 - width, e, width))
        width = 0;
    }
    // ordered write of the round: rank / position offset by ballot + scans
    const uint32_t votes = __ballot_sync(0xffffffffu, width != 0);
    uint64_t x = width;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    consumer_sync();                         // scratch is free again
    if (lane == 31)
    {
      scratch[warp] = (unsigned long long) __popc(votes);
      scratch[kConsumers / 32 + warp] = x;
    }
    consumer_sync();
    uint64_t cbase = 0, wbase = 0, ctot = 0, wtot = 0;
#pragma unroll
    for (int k = 0; k < kConsumers / 32; k++)
    {
      if (k < warp) { cbase += scratch[k]; wbase += scratch[kConsumers / 32 + k]; }
      ctot += scratch[k]; wtot += scratch[kConsumers / 32 + k];
    }
    if (width != 0)
    {
      const uint64_t dst = rec_base + cbase + __popc(votes & ((1u << lane) - 1u));
      const uint64_t po = pos_base + wbase + x - width;
      const uint64_t lb = e + 1 - width;
      if (width >> 32)
        P.result[kResError] = 4;
      if (dst < P.rec_capacity)
      {
        smax_record r;
        r.len = v; r.lb = lb; r.width = width;
        P.recs[dst] = r;
      } else
        P.result[kResOverflow] = 1;
      if (P.positions != nullptr)
      {
        if (po + width <= P.pos_capacity)
          for (uint64_t k = 0; k < width; k++)
            P.positions[po + k] = suf_at(P, lb + k);
        else
          P.result[kResOverflow] = 1;
      }
    }
    rec_base += ctot; pos_base += wtot;
  }
  consumer_sync();
}

constexpr int kFlushLag = 3;         // a mid-scan flush waits only for generations this far behind
constexpr int kStagedTiles = 512;    // tiles of a generation a warp can stage (resolve_generation_staged)
static_assert((size_t) (kThreads / 32) * kStagedTiles * 16 <= sizeof(uint8_t) * kBufs * kStageBytes, "staging regions exceed the ring");
constexpr int kResolveBatch = 4;     // aggregates a lane requests before it looks at the first


// End of the scan: one warp sums one generation with all of its aggregates in flight at
// once -- they are copied into the (now idle) table ring by cp.async, which holds no
// registers, instead of kResolveBatch register loads per round trip to L2.  Tiles that
// have not published yet are polled as in the register path.
__device__ __noinline__ void resolve_generation_staged(const ScanParams &P, ScanSmem &sm,
                                                       uint64_t first, uint32_t ng, uint32_t g,
                                                       uint32_t me, uint32_t *scratch)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // 16 bytes per tile; a warp's region holds kStagedTiles tiles (8 regions fit the ring)
  uint64_t *region = reinterpret_cast<uint64_t *>(&sm.buf[0][0]) + (size_t) warp * 2 * kStagedTiles;
  uint64_t ea = 0, eb = 0, ta = 0, tb = 0;
  for (uint32_t j = lane; j < ng; j += 32)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                 :: "r"(smem_u32(region + 2 * j)), "l"(&P.status[2 * (first + j)]) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  for (uint32_t j = lane; j < ng; j += 32)
  {
    uint64_t wa = region[2 * j], wb = region[2 * j + 1];
    unsigned backoff = 32;
    while ((uint32_t) (wa >> (kValueBits + 2)) != P.epoch ||
           (uint32_t) (wb >> (kValueBits + 2)) != P.epoch)
    {
      __nanosleep(backoff);                  // a straggler has not published yet
      backoff = min(backoff * 2u, 1024u);
      ld_pair(&P.status[2 * (first + j)], wa, wb);
    }
    const uint64_t va = wa & kValueMask, vb = wb & kValueMask;
    ta += va; tb += vb;
    if (j < me) { ea += va; eb += vb; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
  {
    ea += __shfl_xor_sync(0xffffffffu, ea, o);
    eb += __shfl_xor_sync(0xffffffffu, eb, o);
    ta += __shfl_xor_sync(0xffffffffu, ta, o);
    tb += __shfl_xor_sync(0xffffffffu, tb, o);
  }
  if (lane == 0)
  {
    sm.gtot_c[g] = ta; sm.gtot_w[g] = tb; sm.gexc_c[g] = ea; sm.gexc_w[g] = eb;
    scratch[g] = 1;
  }
}


// Executed by the consumer warps together: resolve generations of this CTA from
// base_it on, write their log entries, redo their tiles that lost survivors, and
// keep the rest of the log.  Returns the first generation NOT resolved.
//   final   wait for every generation < upto (end of the scan);
//   else    wait only for generations at least kFlushLag behind (every CTA has
//           all but certainly published those); of the recent ones stop at the
//           first that some CTA has not published yet -- a mid-scan flush must
//           not turn into a grid-wide barrier.
__device__ __noinline__ uint32_t flush_log(const ScanParams &P, ScanSmem &sm, uint32_t base_it,
                                           uint32_t upto, uint32_t me, uint32_t grid, bool final
                                           )
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t *scratch = reinterpret_cast<uint32_t *>(&sm.wlist[0][0]);   // free between passes
  if (tid == 0)
    atomicAdd((unsigned long long *) &P.result[kResFlushes], 1ull);
  consumer_sync();                               // the log is complete
  const uint32_t n = min(sm.log_n, (uint32_t) kLogCap);
  const bool must = final || n > (uint32_t) kLogCap * 3 / 4 || upto - base_it > 40000u ||
                    sm.ndrop > (uint32_t) kMaxDrop / 2;
  const uint32_t ndrop = min(sm.ndrop, (uint32_t) kMaxDrop);   // tiles that lost survivors
  uint32_t resolved = base_it;
  for (uint32_t g0 = base_it; g0 < upto; g0 += kMaxGen)
  {
    const uint32_t gn = min(upto - g0, (uint32_t) kMaxGen);
    // warp w sums the aggregates of generations g0 + w, g0 + w + 8, ...: the lanes
    // read different tiles, kResolveBatch loads in flight each, no atomics (more
    // in flight costs registers that spill in the scan loop: measured slower)
    for (uint32_t g = warp; g < gn; g += kConsumers / 32)
    {
      const uint64_t first = (uint64_t) (g0 + g) * grid;
      const uint32_t ng = (uint32_t) min((uint64_t) grid, (uint64_t) P.ntiles - first);
      // generations well behind this CTA have (all but certainly) been published by
      // everybody: wait for those; the recent ones are taken only if they are
      // complete, unless room has to be made
      const bool wait = final || g0 + g + kFlushLag <= upto || (must && g0 + g == base_it);
      if (final && grid <= (uint32_t) kStagedTiles && !(P.debug & (1 | 256)))
      {
        resolve_generation_staged(P, sm, first, ng, g, me, scratch);
        continue;
      }
      uint64_t ea = 0, eb = 0, ta = 0, tb = 0;
      bool ok = true;
      if (!(P.debug & 1))
        for (uint32_t j0 = lane; j0 < ng; j0 += kResolveBatch * 32)
        {
          uint64_t wa[kResolveBatch], wb[kResolveBatch];
#pragma unroll
          for (int r = 0; r < kResolveBatch; r++)           // all loads first, then the checks
          {
            const uint32_t j = j0 + r * 32;
            wa[r] = wb[r] = 0;
            if (j < ng)
              ld_pair(&P.status[2 * (first + j)], wa[r], wb[r]);
          }
#pragma unroll
          for (int r = 0; r < kResolveBatch; r++)
          {
            const uint32_t j = j0 + r * 32;
            if (j < ng)
            {
              unsigned backoff = 32;
              while ((uint32_t) (wa[r] >> (kValueBits + 2)) != P.epoch ||
                     (uint32_t) (wb[r] >> (kValueBits + 2)) != P.epoch)
              {
                if (!wait) { ok = false; break; }
                __nanosleep(backoff);            // a straggler has not published yet
                backoff = min(backoff * 2u, 1024u);
                ld_pair(&P.status[2 * (first + j)], wa[r], wb[r]);
              }
              const uint64_t va = wa[r] & kValueMask, vb = wb[r] & kValueMask;
              ta += va; tb += vb;
              if (j < me) { ea += va; eb += vb; }
            }
          }
        }
      ok = __all_sync(0xffffffffu, ok);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        ea += __shfl_xor_sync(0xffffffffu, ea, o);
        eb += __shfl_xor_sync(0xffffffffu, eb, o);
        ta += __shfl_xor_sync(0xffffffffu, ta, o);
        tb += __shfl_xor_sync(0xffffffffu, tb, o);
      }
      if (lane == 0)
      {
        sm.gtot_c[g] = ta; sm.gtot_w[g] = tb; sm.gexc_c[g] = ea; sm.gexc_w[g] = eb;
        scratch[g] = ok;
      }
    }
    consumer_sync();
    if (tid == 0)
    {
      // the leading generations whose tiles have all published
      unsigned long long rc = sm.run_c, rw = sm.run_w;
      uint32_t good = 0;
      while (good < gn && scratch[good])
      {
        const unsigned long long ec = sm.gexc_c[good], ew = sm.gexc_w[good];
        sm.gexc_c[good] = rc + ec; sm.gexc_w[good] = rw + ew;
        rc += sm.gtot_c[good]; rw += sm.gtot_w[good];
        good++;
      }
      sm.run_c = rc; sm.run_w = rw;
      scratch[kMaxGen] = good;
    }
    consumer_sync();
    const uint32_t good = scratch[kMaxGen];
    if (good != 0)
    {
      if (!(P.debug & 64))
        write_log(P, sm, n, g0 - base_it, good, g0, me, grid, ndrop);
      for (uint32_t k = 0; k < ndrop; k++)
      {
        const uint32_t d = sm.drop_tag[k];
        if (d != 0 && d - 1 >= g0 - base_it && d - 1 < g0 - base_it + good)
        {
          const uint32_t t = d - 1 - (g0 - base_it);
          slow_tile(P, sm, (uint64_t) me + (uint64_t) (g0 + t) * grid, sm.gexc_c[t], sm.gexc_w[t]);
        }
      }
    }
    resolved = g0 + good;
    consumer_sync();                             // scratch / tables are reused by the next batch
    if (good < gn)
      break;
  }
  // keep the entries of the generations not resolved: order-preserving compaction,
  // tags (they count from the first unresolved generation) rebased
  const uint32_t delta = resolved - base_it;
  if (tid == 0 && delta != 0)
  {
    // the redo list: drop what has been redone, rebase the rest
    uint32_t kept = 0;
    for (uint32_t k = 0; k < ndrop; k++)
      if (sm.drop_tag[k] > delta)
        sm.drop_tag[kept++] = sm.drop_tag[k] - delta;
    sm.ndrop = kept;
  }
  if (final)
  {
    if (tid == 0) sm.log_n = 0;              // the scan is over
  } else if (delta != 0)
  {
    constexpr int kPer = (kLogCap + kConsumers - 1) / kConsumers;
    uint32_t ev[kPer], ew[kPer], et[kPer];
    uint32_t keep = 0;                           // bit r: this thread's r-th entry stays
#pragma unroll
    for (int r = 0; r < kPer; r++)
    {
      const uint32_t e = (uint32_t) r * kConsumers + tid;
      ev[r] = ew[r] = et[r] = 0;
      if (e < n)
      {
        ev[r] = sm.log_v[e]; ew[r] = sm.log_w[e]; et[r] = sm.log_t[e];
        if ((et[r] >> 16) >= delta)
          keep |= 1u << r;
      }
    }
    // entry e = r * kConsumers + tid: count the keepers per (round, warp)
#pragma unroll
    for (int r = 0; r < kPer; r++)
    {
      const uint32_t votes = __ballot_sync(0xffffffffu, (keep >> r) & 1u);
      if (lane == 0)
        scratch[r * (kConsumers / 32) + warp] = __popc(votes);
    }
    if (tid < kSegs) { sm.seg_lo[tid] = ~0u; sm.seg_hi[tid] = 0; }
    consumer_sync();                             // all entries are in registers, counts are in
    uint32_t total = 0;
#pragma unroll
    for (int r = 0; r < kPer; r++)
    {
      uint32_t before = 0;
      for (int q = 0; q < r * (kConsumers / 32) + warp; q++)
        before += scratch[q];
      const uint32_t votes = __ballot_sync(0xffffffffu, (keep >> r) & 1u);
      if ((keep >> r) & 1u)
      {
        const uint32_t dst = before + __popc(votes & ((1u << lane) - 1u));
        sm.log_v[dst] = ev[r]; sm.log_w[dst] = ew[r];
        sm.log_t[dst] = et[r] - (delta << 16);
        atomicMin(&sm.seg_lo[((et[r] >> 16) - delta) % kSegs], dst);
        atomicMax(&sm.seg_hi[((et[r] >> 16) - delta) % kSegs], dst);
      }
    }
    for (int q = 0; q < kPer * (kConsumers / 32); q++)
      total += scratch[q];
    consumer_sync();
    if (tid == 0)
    {
      sm.log_n = total;
    }
  } else if (tid == 0 && sm.log_n > (uint32_t) kLogCap)
    sm.log_n = kLogCap;                          // (entries beyond the capacity were dropped)
  consumer_sync();
  return resolved;
}

// ------------------------------------------------------------ scan kernel
struct Feed            // geometry of the TMA copies of one tile
{
  uint64_t src;        // first table offset copied
  uint32_t dst;        // slot offset it lands at
  uint32_t bytes;      // multiple of 16, > 0
};

__device__ __forceinline__ Feed feed_of(uint64_t toff, uint64_t readable)
{
  Feed f;
  f.src = toff >= (uint64_t) kHalo ? toff - kHalo : 0;
  const uint64_t end = min(toff + kTileBytes + kHalo, readable);
  f.dst = (uint32_t) (f.src + kHalo - toff);
  f.bytes = (uint32_t) (end - f.src);
  return f;
}

// Warp-specialised: 8 consumer warps scan, one producer warp feeds them and
// does the bookkeeping.  Per tile the producer describes it, arms its `ready`
// mbarrier and starts the TMA copies into ring buffers; every consumer warp waits
// for the bytes, runs the pass on its share of the tile and arrives on the
// tile's `done` barrier -- there is no CTA-wide barrier per tile, so a warp that
// is held up in one tile does not hold up the others (they run ahead into the
// tiles already in flight).  When all warps are through, the producer publishes
// the tile's aggregate, hands its buffers to the next tiles and, if the survivor
// log is filling up, asks the consumers (through the descriptor of the next tile
// it starts) to write the log out before they scan that tile.
// WIDE selects the large-value half of the pass: the host picks the variant per
// scan from the share of .llv records in the shard (launch_scan).
template <bool STATS, bool WIDE>
__global__ void __launch_bounds__(kBlockThreads, kMinBlocks)
k_scan(const __grid_constant__ ScanParams P)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  ScanSmem &sm = *reinterpret_cast<ScanSmem *>(smem_raw);
  const int tid = threadIdx.x;
  const uint32_t grid = gridDim.x, me = blockIdx.x;
  const uint64_t base_off = P.g_lo - P.own.a_lo;                 // multiple of 16
  // table bytes that may be read: the arrays are zero padded (SMAX_PAD)
  const uint64_t readable = ((P.own.a_hi - P.own.a_lo + 15) & ~15ull) + 48;

  if (tid == 0)
  {
    for (int q = 0; q < kInFlight; q++)
    {
      mbar_init(&sm.ready[q], 1);
      mbar_init(&sm.done[q], kConsumers / 32);
      mbar_init(&sm.vdone[q], kConsumers / 32);
      sm.tile_c[q] = 0; sm.tile_w[q] = 0; sm.tile_met[q] = 0; sm.tile_drop[q] = 0;
    }
    mbar_init(&sm.vfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sm.log_n = 0; sm.run_c = 0; sm.run_w = 0; sm.ndrop = 0;
    for (int k = 0; k < kSegs; k++) { sm.seg_lo[k] = ~0u; sm.seg_hi[k] = 0; }
  }
  __syncthreads();

  if (tid >= kConsumers)
  {
    // ================================================== producer warp
    if (tid != kConsumers)
      return;
    // a region is "dense" when its tiles keep meeting candidate plateaus: then the
    // bwt tiles are prefetched with the lcp tiles.  First guess from the .llv share.
    bool dense_mode = P.own.nllv * 64 > (P.g_hi - P.g_lo) && !(P.debug & 8);
    // .llv directory entries of a tile
    auto dir_of = [&](uint64_t t, uint32_t &d0, uint32_t &d1)
    {
      d0 = d1 = 0;
      if (P.own.nllv != 0 && !(P.debug & 4))
      {
        const uint64_t toff = base_off + t * kTileBytes;
        d0 = P.own.llvdir[toff >> kLlvBucketShift];
        d1 = P.own.llvdir[((toff + kTileBytes - 1) >> kLlvBucketShift) + 1];
      }
    };
    // The ring: kBufs buffers handed out in order; a tile of a sparse region
    // takes one (lcp), a tile of a dense region two (lcp + bwt), so a sparse
    // region has up to four tiles in flight and a dense one two.
    uint32_t bhead = 0, nfree = kBufs;      // next buffer to hand out / free buffers
    uint32_t needs = 0;                     // 2 bits per tile in flight: buffers it holds
    uint32_t issue_it = 0;                  // next of this CTA's tiles to start
    uint32_t flush_at = 0;                  // the consumers flush before this iteration
    bool flush_pending = false;             // ... once its tile gets described
    // describe the next tile and start its lcp (+ bwt) copies, if buffers are free
    auto issue_next = [&]() -> bool
    {
      const uint64_t t = (uint64_t) me + (uint64_t) issue_it * grid;
      const uint32_t need = dense_mode ? 2u : 1u;
      if (t >= P.ntiles || nfree < need)
        return false;
      const int q = issue_it % kInFlight;
      uint32_t d0, d1;
      dir_of(t, d0, d1);
      TileDesc d;
      d.llv.k0 = d0; d.llv.k1 = d1; d.llv.kfirst = d0 > 0 ? d0 - 1 : 0; d.llv.nrec = 0;
      if (d0 < d1)
        d.llv.nrec = (uint32_t) min((uint64_t) min((uint64_t) d1 + 1, P.own.nllv) - d.llv.kfirst,
                                    (uint64_t) (kLlvSlot + 2));
      d.flags = dense_mode ? kDescBwt : 0u;
      if (flush_pending)
      {
        d.flags |= kDescFlush;
        flush_pending = false;
        flush_at = issue_it;
      }
      d.lbuf = (uint8_t) bhead;
      d.bbuf = (uint8_t) ((bhead + 1) % kBufs);
      d.pad[0] = d.pad[1] = 0;
      sm.desc[q] = d;
      const Feed f = feed_of(base_off + t * kTileBytes, readable);
      mbar_expect_tx(&sm.ready[q], f.bytes * need);     // release: the descriptor is visible
      tma_load(sm.buf[d.lbuf] + f.dst, P.own.lcp + f.src, f.bytes, &sm.ready[q]);
      if (need == 2)
        tma_load(sm.buf[d.bbuf] + f.dst, P.own.bwt + f.src, f.bytes, &sm.ready[q]);
      bhead = (bhead + need) % kBufs;
      nfree -= need;
      needs = (needs & ~(3u << (2 * q))) | (need << (2 * q));
      issue_it++;
      return true;
    };
    // start the copy of a tile's .llv records into the .llv slot
    auto issue_llv = [&](const LlvMeta &m)
    {
      if (m.nrec != 0)
      {
        mbar_expect_tx(&sm.vfull, m.nrec * (uint32_t) sizeof(smax_llv));
        tma_load(sm.llv, P.own.llv + m.kfirst, m.nrec * (uint32_t) sizeof(smax_llv), &sm.vfull);
      }
    };
    while (issue_next()) { }
    if (me < P.ntiles)
      issue_llv(sm.desc[0].llv);
    uint32_t it = 0;
    for (uint64_t tile = me; tile < P.ntiles; tile += grid, it++)
    {
      const int q = it % kInFlight;
      // the large values come first in a pass: when all warps are through with
      // them, the next tile's .llv records are fetched behind this tile's small values
      mbar_wait(&sm.vdone[q], (it / kInFlight) & 1);
      if (tile + grid < P.ntiles)
        issue_llv(sm.desc[(it + 1) % kInFlight].llv);
      mbar_wait(&sm.done[q], (it / kInFlight) & 1);
      const uint32_t c = sm.tile_c[q], met = sm.tile_met[q], drop = sm.tile_drop[q];
      const unsigned long long w = sm.tile_w[q];
      sm.tile_c[q] = 0; sm.tile_w[q] = 0; sm.tile_met[q] = 0; sm.tile_drop[q] = 0;
      publish_aggregate(P.status, (uint32_t) tile, c, w, P.epoch);
      dense_mode = (met >= 4 || sm.desc[q].llv.k1 - sm.desc[q].llv.k0 >= 64) && !(P.debug & 8);
      nfree += (needs >> (2 * q)) & 3u;
      // ask for a flush when the log is half full (a flush keeps what it cannot
      // resolve without waiting, so look at the log itself), not more often than
      // every other tile
      if (!flush_pending && issue_it >= flush_at + 2 &&
          (sm.log_n > (uint32_t) kLogCap / 2 || drop != 0 || issue_it - flush_at >= 30000u))
        flush_pending = true;              // the next tile described carries the request
      while (issue_next()) { }
    }
    return;
  }

  // ==================================================== consumer warps
  uint32_t vphase = 0;                    // parity of the .llv slot to wait for
  uint32_t base_it = 0;                   // first generation this CTA has not resolved yet
  uint32_t it = 0;
  for (uint64_t tile = me; tile < P.ntiles; tile += grid, it++)
  {
    const int q = it % kInFlight;
    const uint64_t toff = base_off + tile * kTileBytes;
    mbar_wait(&sm.ready[q], (it / kInFlight) & 1);
    const TileDesc D = sm.desc[q];
    if (D.flags & kDescFlush)
    {
      // everything logged so far belongs to generations < it, which every CTA has
      // published or is about to
      base_it = flush_log(P, sm, base_it, it, me, grid, false);
    }
    PassCtx C;
    C.tile_lo = P.own.a_lo + toff;
    C.it16 = it - base_it;
    C.par = q;
    C.sl = sm.buf[D.lbuf];
    C.sb = (D.flags & kDescBwt) ? sm.buf[D.bbuf] : nullptr;
    C.sv = sm.llv;
    C.vparity = vphase;
    C.llv = D.llv;
    if (D.llv.nrec != 0)
      vphase ^= 1u;                        // the pass waits for this phase of the .llv slot
    // edges of the table: the left halo of the first tile comes from the left
    // neighbour shard (or repeats the first entry, which sends every plateau
    // that touches the edge into the walk that reports the missing range);
    // bytes past the zero pad are zero
    {
      const Feed f = feed_of(toff, readable);
      if (f.dst != 0 || f.dst + f.bytes != (uint32_t) kStageBytes)
      {
        uint8_t *wl = sm.buf[D.lbuf], *wb = (D.flags & kDescBwt) ? sm.buf[D.bbuf] : nullptr;
        consumer_sync();                   // (rare) the buffers are written by hand: all warps here
        if (f.dst != 0 && tid < kHalo)
        {
          const uint64_t a_lo = P.own.a_lo;
          uint32_t lv = 0, bv = 0;
          if (a_lo >= (uint64_t) kHalo)
          {
            const uint64_t qq = a_lo - kHalo + tid;
            const TableView *tv = view_for(P, qq);
            if (tv != nullptr) { lv = tv->lcp[qq - tv->a_lo]; bv = tv->bwt[qq - tv->a_lo]; }
            else { lv = P.own.lcp[0]; bv = 0; }
          }
          wl[tid] = (uint8_t) lv;
          if (wb != nullptr) wb[tid] = (uint8_t) bv;
        }
        for (uint32_t i = f.dst + f.bytes + tid; i < (uint32_t) kStageBytes; i += kConsumers)
        {
          wl[i] = 0;
          if (wb != nullptr) wb[i] = 0;
        }
        consumer_sync();
      }
    }
    tile_pass<STATS, WIDE>(P, sm, C);
  }
  // ---- the survivors still in the log; the owner of the last tile also resolves
  // every generation to report the totals
  consumer_sync();
  const bool owns_last = P.ntiles != 0 && (P.ntiles - 1) % grid == me;
  if (it > base_it && (sm.log_n != 0 || sm.ndrop != 0 || owns_last) && !(P.debug & 128))
    flush_log(P, sm, base_it, it, me, grid, true);
  if (owns_last && tid == 0)
  {
    P.result[kResCount] = sm.run_c;
    P.result[kResPositions] = sm.run_w;
    // one-sided count exchange: the shard's record count goes straight into every
    // shard's count array (P2P stores over NVLink), tagged with the step
    const uint64_t word = (P.exchange_tag << 40) | (sm.run_c & ((1ull << 40) - 1));
    for (int k = 0; k < P.npeers; k++)
      asm volatile("st.release.sys.global.u64 [%0], %1;"
                   :: "l"(P.peer_counts[k] + P.my_rank), "l"(word) : "memory");
  }
  // the last CTA to leave clears the other result block for the next scan
  if (tid == 0)
  {
    __threadfence();
    const uint32_t done = atomicAdd(&P.ctrl[1], 1u);
    if (done == gridDim.x - 1)
    {
      P.ctrl[1] = 0;
      if (P.ntiles == 0)
      {
        P.result[kResCount] = 0;
        P.result[kResPositions] = 0;
        for (int k = 0; k < P.npeers; k++)
          asm volatile("st.release.sys.global.u64 [%0], %1;"
                       :: "l"(P.peer_counts[k] + P.my_rank), "l"(P.exchange_tag << 40) : "memory");
      }
      for (int k = 0; k < kResSlots; k++)   // result blocks ping-pong: no memset per scan
        P.result_next[k] = 0;
    }
  }
}

// --------------------------------------------------------------- launchers
// Cooperative launch: the ordered prefix exchange needs every CTA of the grid
// resident at the same time (grid <= SMs x resident CTAs per SM, computed by the
// caller); the runtime then guarantees co-residency instead of assuming it.
static const void *scan_kernel(bool stats, bool wide)
{
  return stats ? (wide ? (const void *) k_scan<true, true> : (const void *) k_scan<true, false>)
               : (wide ? (const void *) k_scan<false, true> : (const void *) k_scan<false, false>);
}

cudaError_t launch_scan(const ScanParams &p, bool stats, int grid, cudaStream_t st)
{
  // a shard where more than one entry in eight is a large value has wide repeats:
  // take the variant whose warps walk runs of equal large values together
  const bool wide = p.own.nllv * 8 > p.own.a_hi - p.own.a_lo;
  void *args[] = {(void *) &p};
  return cudaLaunchCooperativeKernel(scan_kernel(stats, wide), dim3(grid), dim3(kBlockThreads), args,
                                     sizeof(ScanSmem), st);
}

int scan_blocks_per_sm(bool stats)
{
  int least = 1 << 30;
  for (int wide = 0; wide < 2; wide++)
  {
    const void *fn = scan_kernel(stats, wide != 0);
    if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int) sizeof(ScanSmem)) != cudaSuccess)
      return 0;
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, kBlockThreads, sizeof(ScanSmem)) !=
            cudaSuccess || n <= 0)
      return 0;
    least = n < least ? n : least;
  }
  return least;
}

}  // namespace smax_ring
