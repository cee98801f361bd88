/*
  smax_scan.cu -- hand-written sm_100a kernels of the supermaximal-repeat
  scan.  ONE fused pass over the lcptab replaces the reference's stack sweep
  (/root/reference/src/match/esa-bottomup.c:116-273) and its per-node
  left-character bookkeeping (/root/reference/src/match/esa-maxpairs.c:181-360).

  k_scan: CTAs of 4 warps take 16 KiB lcptab tiles from a ticket (tiles are
  handed out in suffix-array order, so a tile only ever waits for tiles that
  are already being worked on: no cooperative launch, no round-robin tail).
  Per tile:

    feed  one TMA bulk copy (cp.async.bulk.shared::cluster.global, mbarrier
          complete_tx) of the tile's lcp bytes + a 16-byte halo either side.
          While it is in flight the CTA works on the tile's large values.
    K1a   large values (byte 255) in .llv RECORD space, out of the compact
          8-byte records {position - a_lo, value} built at upload (k_llvpack;
          the tile's records are found through a per-4096-entry directory): a
          record ends a plateau iff its right neighbour is no consecutive
          record with a value >= its own and its run is entered from a smaller
          value.  Run ends are compacted per warp (ballot) before K2.
    K1b   small values, flat and bit-parallel (smax_swar.h): every lane filters
          its 16-byte chunks for a byte >= minlength; the hits, compacted per
          warp, are classified with SWAR byte arithmetic: ends of runs that
          fall to a smaller value, entered from a smaller value 1, 2 or 3
          entries back (SA width 2, 3, 4).  Runs of >= 4 equal values are
          walked together with their left characters (K2 ends the walk at the
          first repeated character, so a wide plateau costs O(alphabet)).
    K2    left-distinctness, bit-parallel on the chunk's bwt words (fetched
          from global memory only for chunks that passed the filter) for widths
          <= 4; a 256-bit alphabet mask otherwise.  Specials (>= 254) never
          collide under the GenomeTools convention (esa-maxpairs.c:24-31).
    K3    order-preserving compaction + emit.  A survivor [lb, e] sets bit e of
          the tile's END bitmap and bit lb of its START bitmap in shared memory
          (supermaximal repeats are disjoint SA intervals, so the two bitmaps
          describe them completely, in order, at any density).  The tile's
          (records, positions) aggregate goes through a decoupled look-back
          over epoch-tagged status words (no memset between scans); then every
          thread writes the records of its bitmap words in suffix-array order,
          the occurrence positions suf[lb..lb+width) gathered right behind them.

  k_llvdir / k_llvpack build the .llv bucket directory and the compact records
  at upload time.
*/
#include <cstring>
#include "smax_kernels.cuh"
#include "smax_swar.h"

namespace smax {

// ------------------------------------------------------------------ utils
__device__ __forceinline__ void ld_pair(const uint64_t *p, uint64_t &a, uint64_t &b)
{
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}

__device__ __forceinline__ void st_pair(uint64_t *p, uint64_t a, uint64_t b)
{
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1,%2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}

__device__ __forceinline__ uint64_t pack_status(uint32_t epoch, uint64_t state, uint64_t value)
{
  return ((uint64_t) epoch << (kValueBits + 2)) | (state << kValueBits) | (value & kValueMask);
}

__device__ __forceinline__ bool status_is(uint64_t w, uint32_t epoch, uint64_t state)
{
  return (w >> kValueBits) == (((uint64_t) epoch << 2) | state);
}

// read-only, streaming loads of table bytes that are used once
__device__ __forceinline__ uint2 ldg_rec(const uint2 *p)
{
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

__device__ __forceinline__ uint4 ldg_chunk(const uint8_t *p)
{
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
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

// Inconsistent tables (a 255 byte has no .llv record) and plateaus that leave
// every resident view are reported through the result block; the scan then
// fails on the host.  Such a value reads as "larger than everything" so that
// walks stop.
constexpr uint64_t kBadValue = ~0ull;
constexpr uint64_t kErrTables = 1, kErrRange = 6;   // codes in result[kResError]

// resolved lcp value at an arbitrary index (slow, fully general; used when a
// run leaves the shard's own arrays and for the length of a large survivor)
__device__ __noinline__ uint64_t value_at(const ScanParams &P, uint64_t q)
{
  const TableView *tv = view_for(P, q);
  if (tv == nullptr) { P.result[kResError] = kErrRange; return kBadValue; }
  const uint32_t b = tv->lcp[q - tv->a_lo];
  if (b < 255)
    return b;
  uint64_t k;
  if (!llv_find(*tv, q, k)) { P.result[kResError] = kErrTables; return kBadValue; }
  return tv->llv[k].value;
}

__device__ __noinline__ uint32_t byte_at_left(const ScanParams &P, uint64_t q, bool want_bwt)
{
  const TableView *tv = view_for(P, q);
  if (tv == nullptr) { P.result[kResError] = kErrRange; return 255; }
  return want_bwt ? tv->bwt[q - tv->a_lo] : tv->lcp[q - tv->a_lo];
}

__device__ __forceinline__ uint32_t bwt_at(const ScanParams &P, uint64_t q)
{
  return q >= P.own.a_lo ? (uint32_t) P.own.bwt[q - P.own.a_lo] : byte_at_left(P, q, true);
}

__device__ __forceinline__ uint64_t suf_at(const ScanParams &P, uint64_t i)
{
  const TableView *tv = view_for(P, i);
  if (tv == nullptr || tv->suf == nullptr) { P.result[kResError] = kErrRange; return 0; }
  const uint64_t o = i - tv->a_lo;
  if (i >= tv->a_hi) { P.result[kResError] = 3; return 0; }
  return P.sufbytes == 8 ? reinterpret_cast<const uint64_t *>(tv->suf)[o]
                         : (uint64_t) reinterpret_cast<const uint32_t *>(tv->suf)[o];
}

// value of the own record k given its compact form
__device__ __forceinline__ uint64_t rec_value(const ScanParams &P, uint32_t k, uint32_t v32)
{
  return v32 != kLlvEscape ? (uint64_t) v32 : P.own.llv[k].value;
}

// K2 bookkeeping: the set of left characters met so far (256-bit alphabet mask)
struct CharSet
{
  uint64_t m0, m1, m2, m3;
  __device__ __forceinline__ CharSet() : m0(0), m1(0), m2(0), m3(0) {}
  // returns true when c repeats a character of the set (specials never do under the GT policy)
  __device__ __forceinline__ bool add(uint32_t c, bool gt_policy)
  {
    if (gt_policy && c >= 254)
      return false;
    const uint64_t bit = 1ull << (c & 63);
    uint64_t hit;
    switch (c >> 6)
    {
      case 0: hit = m0 & bit; m0 |= bit; break;
      case 1: hit = m1 & bit; m1 |= bit; break;
      case 2: hit = m2 & bit; m2 |= bit; break;
      default: hit = m3 & bit; m3 |= bit; break;
    }
    return hit != 0;
  }
};

// ------------------------------------------------------ run walks (K1 + K2)
// The run of small value b that ends at e is known to cover [e - 1, e].  Walk
// left, left characters first: the plateau is [s - 1, e] (s = start of the run)
// if it is entered from a smaller value, and bwt[s - 1 .. e] must be pairwise
// distinct -- so the walk ends at the first repeated character (FULL = false),
// which bounds it by the alphabet size however wide the run is.  FULL walks to
// the start of the run regardless (statistics build: the plateau's true width).
// Returns the SA width of the local-maximum plateau (0: entered from a larger
// value, or -- !FULL -- a left character repeats); *distinct says whether K2 holds.
template <bool FULL>
__device__ __noinline__ uint64_t small_run_plateau(const ScanParams &P, uint64_t e, uint32_t b,
                                                   bool *distinct)
{
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  const uint64_t a_lo = P.own.a_lo;
  CharSet cs;
  bool dup = cs.add(bwt_at(P, e), gt_policy);
  dup |= cs.add(bwt_at(P, e - 1), gt_policy);
  uint64_t s = e - 1;
  *distinct = false;
  if (dup && !FULL)
    return 0;
  for (;;)
  {
    if (s == 0)
      break;                              // start of the table
    const uint64_t q = s - 1;
    dup |= cs.add(bwt_at(P, q), gt_policy);
    if (dup && !FULL)
      return 0;
    const uint32_t pb = q >= a_lo ? (uint32_t) P.own.lcp[q - a_lo] : byte_at_left(P, q, false);
    if (pb == b) { s = q; continue; }
    if (pb > b)
      return 0;                           // entered from a larger value (255 stands for one)
    break;
  }
  *distinct = !dup;
  return e - s + 2;
}

// The same for a run of equal large values v that ends at position p = record k
// of the own shard: walked in record space (compact records), into the left
// neighbours through value_at when it leaves the own arrays.
template <bool FULL>
__device__ __noinline__ uint64_t large_run_plateau(const ScanParams &P, uint32_t k, uint64_t p,
                                                   uint64_t v, bool *distinct)
{
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  const uint64_t a_lo = P.own.a_lo;
  CharSet cs;
  bool dup = cs.add(bwt_at(P, p), gt_policy);
  uint64_t s = p;
  uint32_t kk = k;
  *distinct = false;
  for (;;)
  {
    if (s == 0)
      break;
    const uint64_t q = s - 1;
    dup |= cs.add(bwt_at(P, q), gt_policy);
    if (dup && !FULL)
      return 0;
    uint64_t pv;
    if (q >= a_lo)
    {
      if (kk == 0)
        break;                            // no record at q: a small value, rise
      const uint2 pr = ldg_rec(&P.own.llvc[kk - 1]);
      if ((uint64_t) pr.x != q - a_lo)
        break;
      pv = rec_value(P, kk - 1, pr.y);
      if (pv == v) { s = q; kk--; continue; }
    } else
    {
      pv = value_at(P, q);                // kBadValue on error: stops the walk
      if (pv == v) { s = q; continue; }
    }
    if (pv > v)
      return 0;
    break;
  }
  *distinct = !dup;
  return p - s + 2;
}

// K1 for the large value of record k (compact form c) between its neighbour records pv / nx
// (x = kNoRecord: no such record): does a plateau candidate end here, and does its run of
// equal values have to be walked?  The rare case of a value that does not fit the compact
// record compares the 16-byte records.
__device__ __noinline__ void classify_escaped(const ScanParams &P, uint32_t k, uint2 c, uint2 pv, uint2 nx,
                                              bool &cand, bool &walk)
{
  const bool adj_n = nx.x == c.x + 1, adj_p = pv.x + 1 == c.x;
  const uint64_t v = rec_value(P, k, c.y);
  const uint64_t nv = adj_n ? rec_value(P, k + 1, nx.y) : 0;
  const uint64_t qv = adj_p ? rec_value(P, k - 1, pv.y) : 0;
  cand = v >= P.minlength && !(adj_n && nv >= v) && !(adj_p && qv > v);
  walk = adj_p && qv == v;
}

__device__ __forceinline__ void classify_large(const ScanParams &P, uint32_t k, uint2 c, uint2 pv, uint2 nx,
                                               uint32_t m32, bool &cand, bool &walk)
{
  const bool adj_n = nx.x == c.x + 1, adj_p = pv.x + 1 == c.x;
  if (!P.has_escape ||
      (c.y != kLlvEscape && !(adj_n && nx.y == kLlvEscape) && !(adj_p && pv.y == kLlvEscape)))
  {
    // a record ends a plateau iff its right neighbour is no consecutive record with a value
    // >= its own and it is not entered from a larger one
    cand = c.y >= m32 && !(adj_n && nx.y >= c.y) && !(adj_p && pv.y > c.y);
    walk = adj_p && pv.y == c.y;
  } else
    classify_escaped(P, k, c, pv, nx, cand, walk);
  walk |= c.x == 0 && P.own.a_lo > 0;      // the shard's edge: what lies left of it?
  walk &= cand;
}

// --------------------------------------------------- shared memory layout
constexpr int kStageBytes = kHalo + kUnitBytes + kHalo;

// Everything a warp needs for its unit (kUnitBytes lcptab entries): the warps of a CTA share
// nothing but the allocation, so no warp ever waits for another one.
struct WarpSmem
{
  // lcp[kHalo + i] = lcptab[unit_lo + i], i in [-kHalo, kUnitBytes + kHalo)
  alignas(128) uint8_t lcp[kStageBytes];
  alignas(16) uint32_t endbits[kBitWords];    // bit e: a supermaximal repeat ends at unit offset e
  alignas(16) uint32_t startbits[kBitWords];  // bit lb: ... starts at unit offset lb
  alignas(16) uint16_t bigk[kBitWords];       // per END word: record (+1, from the unit's first) of a
                                              //   large survivor that ends there (saves the .llv search)
  union                                   // (the large-value pass is over when the small one starts)
  {
    uint16_t candlist[kEndList];          // large-value candidates of SA width 2 (record - first)
    uint16_t endlist[kSmallEnds];         // END candidates of the chunks that passed the filter (unit offsets)
  };
  union
  {
    uint16_t walklist[kLlvBatch * 64];    // records that end a run of EQUAL large values
    uint8_t chunklist[kUnitChunks];       // chunks that passed the filter
  };
  unsigned long long open_width;          // width of the survivor that starts left of the unit
  uint32_t any;                           // != 0: the unit has survivors
  alignas(8) uint64_t ready;              // mbarrier: the unit's lcp bytes have landed
};

// K3, first half: the survivor [e + 1 - width, e] (e at unit offset o) is marked
// in the unit's bitmaps.
__device__ __forceinline__ void mark_survivor(WarpSmem &ws, uint32_t o, uint64_t width)
{
  atomicOr(&ws.endbits[o >> 5], 1u << (o & 31));
  if (width <= (uint64_t) o + 1)
  {
    const uint32_t lb = o + 1 - (uint32_t) width;
    atomicOr(&ws.startbits[lb >> 5], 1u << (lb & 31));
  } else
    ws.open_width = width;                // starts left of the unit: at most one per unit
  ws.any = 1;
}

// SA width of the survivor that ends at unit offset o: its start is the nearest
// START bit below o (survivors are disjoint and each has width >= 2)
__device__ __forceinline__ uint64_t survivor_width(const WarpSmem &ws, uint32_t o)
{
  int w = (int) (o >> 5);
  uint32_t m = ws.startbits[w] & ((1u << (o & 31)) - 1u);
  while (m == 0 && w > 0)
    m = ws.startbits[--w];
  if (m == 0)
    return ws.open_width;
  return (uint64_t) (o - ((uint32_t) w * 32u + 31u - (uint32_t) __clz(m)) + 1u);
}

// repeat length of the large survivor that ends at shard offset off: its record among the
// unit's records [k0, k1) of the compact table (the rare miss of the per-word hint)
__device__ __noinline__ uint64_t large_value_of(const ScanParams &P, uint32_t k0, uint32_t k1, uint32_t off)
{
  uint32_t lo = k0, hi = k1;
  while (lo < hi)
  {
    const uint32_t mid = (lo + hi) >> 1;
    if (ldg_rec(&P.own.llvc[mid]).x < off) lo = mid + 1; else hi = mid;
  }
  if (lo >= k1)
  {
    P.result[kResError] = kErrTables;
    return kBadValue;
  }
  const uint2 r = ldg_rec(&P.own.llvc[lo]);
  if (r.x != off)
  {
    P.result[kResError] = kErrTables;
    return kBadValue;
  }
  return rec_value(P, lo, r.y);
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

struct Feed            // geometry of the TMA copy of one unit
{
  uint64_t src;        // first table offset copied
  uint32_t dst;        // stage offset it lands at
  uint32_t bytes;      // multiple of 16, > 0
};

__device__ __forceinline__ Feed feed_of(uint64_t toff, uint64_t readable)
{
  Feed f;
  f.src = toff >= (uint64_t) kHalo ? toff - kHalo : 0;
  const uint64_t end = min(toff + kUnitBytes + kHalo, readable);
  f.dst = (uint32_t) (f.src + kHalo - toff);
  f.bytes = (uint32_t) (end - f.src);
  return f;
}

// ------------------------------------------------------------ scan kernel
template <bool STATS>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_scan(const __grid_constant__ ScanParams P)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  WarpSmem &ws = reinterpret_cast<WarpSmem *>(smem_raw)[warp];
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint64_t a_lo = P.own.a_lo;
  const uint64_t base_off = P.g_lo - a_lo;                         // multiple of 16
  // table bytes that may be read: the arrays are zero padded (SMAX_PAD)
  const uint64_t readable = ((P.own.a_hi - a_lo + 15) & ~15ull) + 48;
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  const uint32_t m32 = (uint32_t) min(P.minlength, (uint64_t) kLlvEscape);
  uint32_t kadd; int himode;
  smax_ge_consts(P.mb, &kadd, &himode);
  uint64_t stat[4] = {0, 0, 0, 0};    // candidates, their widths, .llv records inspected, survivor widths

  if (lane == 0)
  {
    mbar_init(&ws.ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t parity = 0;
  // the survivor arena is handed out in chunks of kArenaChunk entries (one atomic per chunk and
  // warp: a unit's entries are consecutive, a warp fills its chunk unit by unit)
  uint64_t chunk_next = 0;            // next free entry of the warp's chunk,
  uint32_t chunk_left = 0;            //   entries left in it
  // units are taken from a ticket, kTicketUnits consecutive ones at a time (the next ticket is
  // requested while the last unit of this one is worked on)
  uint32_t unit = 0, unit_end = 0;
  if (lane == 0)
    unit = atomicAdd(&P.ctrl[0], (uint32_t) kTicketUnits);
  unit = __shfl_sync(0xffffffffu, unit, 0);
  unit_end = unit + kTicketUnits;
  __syncwarp();

  while (unit < P.nunits)
  {
    const uint64_t toff = base_off + (uint64_t) unit * kUnitBytes;
    const uint64_t unit_lo = a_lo + toff;
    uint32_t ticket = 0;
    {
      // clear the unit's state: END / START bitmaps and the record hints are adjacent
      uint4 *z = reinterpret_cast<uint4 *>(ws.endbits);
      const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int i = lane; i < (int) ((2 * kBitWords * 4 + kBitWords * 2) / 16); i += 32)
        z[i] = zero;
    }
    if (lane == 0)
    {
      ws.any = 0;
      ws.open_width = 0;
      if (unit + 1 == unit_end)
        ticket = atomicAdd(&P.ctrl[0], (uint32_t) kTicketUnits);
      const Feed f = feed_of(toff, readable);
      if (f.dst != 0)
      {
        // left edge of the shard's arrays: the halo comes from the left neighbour shard, or
        // repeats the first entry (which sends every plateau that touches the edge into the
        // walk that reports the missing range); the table itself starts with lcp[0] = 0
        for (uint32_t i = 0; i < f.dst; i++)
        {
          uint32_t lv = 0;
          if (a_lo >= (uint64_t) kHalo)
          {
            const uint64_t qq = a_lo - kHalo + i;
            const TableView *tv = view_for(P, qq);
            lv = tv != nullptr ? tv->lcp[qq - tv->a_lo] : P.own.lcp[0];
          }
          ws.lcp[i] = (uint8_t) lv;
        }
      }
      // (the stage was read through the generic proxy: order those reads before the bulk copy)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&ws.ready, f.bytes);
      tma_load(ws.lcp + f.dst, P.own.lcp + f.src, f.bytes, &ws.ready);
    }
    __syncwarp();
    // ends at or beyond g_hi belong to the next shard
    const uint32_t valid = (uint32_t) min((uint64_t) kUnitBytes, P.g_hi - unit_lo);
    // the unit's .llv records [kt0, k1) (exact: the per-unit directory built at upload)
    uint32_t kt0 = 0, k1 = 0;
    if (P.own.nllv != 0)
    {
      kt0 = P.unitdir[unit];
      k1 = P.unitdir[unit + 1];
    }

    // ---------------- K1a: large values, in record space (the lcp bytes are in flight)
    if (kt0 < k1 && !(P.debug & 4))
    {
      const uint32_t k0 = kt0 & ~1u;       // rows start at an even record (16-byte loads)
      const uint32_t nllv = (uint32_t) P.own.nllv;
      const uint2 *llvc = P.own.llvc;
      uint16_t *cands = ws.candlist, *walks = ws.walklist;
      uint32_t nc = 0, nw = 0;
      // rows of 64 records (two per lane, one 16-byte load), kLlvBatch rows in flight; neighbour
      // records come from the adjacent lanes
#pragma unroll 1
      for (uint32_t kb = k0;; kb += kLlvBatch * 64)
      {
        const bool last = kb >= k1;
        if (nw != 0)
        {
          // records that end a run of EQUAL large values (or sit on the shard's edge): walk
          __syncwarp();
#pragma unroll 1
          for (uint32_t i = lane; i < nw; i += 32)
          {
            const uint32_t k = k0 + walks[i];
            const uint2 r = ldg_rec(&llvc[k]);
            const uint64_t p = a_lo + r.x, v = rec_value(P, k, r.y);
            bool ok = false;
            const uint64_t width = large_run_plateau<STATS>(P, k, p, v, &ok);
            if (STATS) atomicAdd((unsigned long long *) &P.result[kResWalks], 1ull);
            if (width != 0)
            {
              if (STATS) { stat[0]++; stat[1] += width; }
              if (ok)
              {
                if (STATS) stat[3] += width;
                const uint32_t o = r.x - (uint32_t) toff;
                mark_survivor(ws, o, width);
                ws.bigk[o >> 5] = (uint16_t) (k - kt0 + 1);
              }
            }
          }
          nw = 0;
          __syncwarp();
        }
        if (last ? nc != 0 : nc > (uint32_t) (kEndList - kLlvBatch * 64))
        {
          // K2 + marking of the listed candidates of SA width 2: two entries per lane and
          // round, their four left characters in flight together
          __syncwarp();
#pragma unroll 1
          for (uint32_t i = lane; i < nc; i += 64)
          {
            const bool two = i + 32 < nc;
            const uint32_t ka = k0 + cands[i], kb2 = two ? k0 + cands[i + 32] : ka;
            const uint2 ra = ldg_rec(&llvc[ka]), rb = ldg_rec(&llvc[kb2]);
            const uint32_t a0 = P.own.bwt[ra.x - 1], a1 = P.own.bwt[ra.x];
            const uint32_t b0 = P.own.bwt[rb.x - 1], b1 = P.own.bwt[rb.x];
            if (a0 != a1 || (gt_policy && a0 >= 254))
            {
              if (STATS) stat[3] += 2;
              const uint32_t o = ra.x - (uint32_t) toff;
              mark_survivor(ws, o, 2);
              ws.bigk[o >> 5] = (uint16_t) (ka - kt0 + 1);
            }
            if (two && (b0 != b1 || (gt_policy && b0 >= 254)))
            {
              if (STATS) stat[3] += 2;
              const uint32_t o = rb.x - (uint32_t) toff;
              mark_survivor(ws, o, 2);
              ws.bigk[o >> 5] = (uint16_t) (kb2 - kt0 + 1);
            }
          }
          nc = 0;
          __syncwarp();
        }
        if (last)
          break;
        uint4 q[kLlvBatch];
        uint2 edge[kLlvBatch];
        const uint4 none = make_uint4(kNoRecord, 0, kNoRecord, 0);     // (no record: never adjacent)
#pragma unroll
        for (int u = 0; u < kLlvBatch; u++)
        {
          const uint32_t k = kb + u * 64 + 2 * lane;
          q[u] = k < nllv ? __ldg(reinterpret_cast<const uint4 *>(&llvc[k])) : none;    // (padded behind nllv)
          // the record before the row (lane 0) and the one after it (lane 31)
          edge[u] = make_uint2(kNoRecord, 0);
          if (lane == 0 && k > 0 && k < nllv)
            edge[u] = ldg_rec(&llvc[k - 1]);
          if (lane == 31 && k + 2 < nllv)
            edge[u] = ldg_rec(&llvc[k + 2]);
        }
#pragma unroll
        for (int u = 0; u < kLlvBatch; u++)
        {
          const uint32_t k = kb + u * 64 + 2 * lane;
          const uint2 ca = make_uint2(q[u].x, q[u].y), cb = make_uint2(q[u].z, q[u].w);
          uint2 pv, nx;
          pv.x = __shfl_up_sync(0xffffffffu, cb.x, 1);
          pv.y = __shfl_up_sync(0xffffffffu, cb.y, 1);
          nx.x = __shfl_down_sync(0xffffffffu, ca.x, 1);
          nx.y = __shfl_down_sync(0xffffffffu, ca.y, 1);
          if (lane == 0) pv = edge[u];
          if (lane == 31) nx = edge[u];
          bool in_a = k >= kt0 && k < k1;
          bool in_b = k + 1 < k1;
          if (STATS) stat[2] += (in_a ? 1 : 0) + (in_b ? 1 : 0);
          bool walk_a = false, walk_b = false;
          if (in_a) classify_large(P, k, ca, pv, cb, m32, in_a, walk_a);
          if (in_b) classify_large(P, k + 1, cb, ca, nx, m32, in_b, walk_b);
          const bool sa = in_a && !walk_a, sb = in_b && !walk_b;    // entered from a smaller value: SA width 2
          if (STATS) { const uint32_t c2 = (sa ? 1 : 0) + (sb ? 1 : 0); stat[0] += c2; stat[1] += 2 * c2; }
          const uint32_t va = __ballot_sync(0xffffffffu, sa), vb = __ballot_sync(0xffffffffu, sb);
          if (sa) cands[nc + __popc(va & lt_mask)] = (uint16_t) (k - k0);
          nc += __popc(va);
          if (sb) cands[nc + __popc(vb & lt_mask)] = (uint16_t) (k + 1 - k0);
          nc += __popc(vb);
          if (__any_sync(0xffffffffu, walk_a | walk_b))
          {
            const uint32_t wa = __ballot_sync(0xffffffffu, walk_a), wb = __ballot_sync(0xffffffffu, walk_b);
            if (walk_a) walks[nw + __popc(wa & lt_mask)] = (uint16_t) (k - k0);
            nw += __popc(wa);
            if (walk_b) walks[nw + __popc(wb & lt_mask)] = (uint16_t) (k + 1 - k0);
            nw += __popc(wb);
          }
        }
      }
    }

    // ---------------- K1b: small values out of the staged unit
    mbar_wait(&ws.ready, parity);
    parity ^= 1u;
    if (P.minlength < 255 && !(P.debug & 2))
    {
      uint8_t *list = ws.chunklist;
      // phase A: which of the unit's 256 chunks hold a byte >= the threshold?
      uint32_t n = 0;
#pragma unroll
      for (int j = 0; j < kUnitChunks / 32; j++)
      {
        const uint32_t c = (uint32_t) (j * 32 + lane), o0 = c * kChunk;
        bool hit = false;
        if (o0 < valid)
        {
          const uint4 x = *reinterpret_cast<const uint4 *>(ws.lcp + kHalo + o0);
          hit = (smax_ge(x.x, kadd, himode) | smax_ge(x.y, kadd, himode) | smax_ge(x.z, kadd, himode) |
                 smax_ge(x.w, kadd, himode)) != 0;
        }
        const uint32_t votes = __ballot_sync(0xffffffffu, hit);
        if (hit)
        {
          list[n + __popc(votes & lt_mask)] = (uint8_t) c;
          // the chunk's left characters will be wanted: start them on their way to L2
          asm volatile("prefetch.global.L2 [%0];" :: "l"(P.own.bwt + toff + o0));
        }
        n += __popc(votes);
      }
      __syncwarp();
      if (P.debug & 8)
        n = 0;
      // phase B, second level: one END candidate (a byte >= minlength that is followed by a
      // smaller one) per lane: K1 (is its run entered from a smaller value? SA width 2, 3, 4 out of
      // the staged bytes; longer runs are walked) + K2 on the left characters + marking
      uint16_t *ends = ws.endlist;
      uint32_t ne = 0;
      auto process_ends = [&]()
      {
        __syncwarp();
#pragma unroll 1
        for (uint32_t i = lane; i < ne; i += 32)
        {
          const uint32_t o = ends[i];
          const uint8_t *lp = ws.lcp + kHalo + o;
          // the left characters bwt[o - 3 .. o] as one word (byte 3 = bwt[o]): requested first
          const uint64_t g = toff + o;
          const uint8_t *bp = P.own.bwt + (g & ~3ull);
          const uint32_t bhi = *reinterpret_cast<const uint32_t *>(bp);
          const uint32_t blo = g >= 4 ? *reinterpret_cast<const uint32_t *>(bp - 4) : 0u;
          const uint32_t v = lp[0], l1 = lp[-1];
          if (l1 > v)
            continue;                        // entered from a larger value (255 stands for one)
          const uint32_t l2 = lp[-2], l3 = lp[-3];
          const uint32_t sh = 8 * (((uint32_t) g & 3u) + 1u);
          const uint32_t cw = sh == 32 ? bhi : __funnelshift_r(blo, bhi, sh);
          const uint32_t c0 = cw >> 24, c1 = (cw >> 16) & 255u, c2 = (cw >> 8) & 255u, c3 = cw & 255u;
          const uint32_t lim = gt_policy ? 254u : 256u;      // specials never collide (GT policy)
          uint64_t width = 2;
          bool ok = !(c0 == c1 && c0 < lim);
          if (l1 == v)
          {
            const bool dup3 = !ok || (c0 == c2 && c0 < lim) || (c1 == c2 && c1 < lim);
            if (l2 < v)
            {
              width = 3;
              ok = !dup3;
            } else if (l2 == v && l3 < v)
            {
              width = 4;
              ok = !(dup3 || (c3 == c0 && c3 < lim) || (c3 == c1 && c3 < lim) || (c3 == c2 && c3 < lim));
            } else if (l2 == v && l3 == v)
            {
              width = small_run_plateau<STATS>(P, unit_lo + o, v, &ok);
              if (STATS) atomicAdd((unsigned long long *) &P.result[kResWalks], 1ull);
            } else
              width = 0;                     // entered from a larger value further left
          }
          if (width != 0)
          {
            if (STATS) { stat[0]++; stat[1] += width; }
            if (ok)
            {
              if (STATS) stat[3] += width;
              mark_survivor(ws, o, width);
            }
          }
        }
        ne = 0;
        __syncwarp();
      };
      // phase B, first level: the END candidates of the listed chunks
#pragma unroll 1
      for (uint32_t i0 = 0; i0 < n; i0 += 32)
      {
        const uint32_t i = i0 + lane;
        uint32_t em = 0;                     // bit j: an END candidate at byte j of the lane's chunk
        uint32_t o0 = 0;
        if (i < n)
        {
          o0 = (uint32_t) list[i] * kChunk;
          const uint8_t *lp = ws.lcp + kHalo + o0;
          const uint4 x = *reinterpret_cast<const uint4 *>(lp);
          const uint32_t w[5] = {x.x, x.y, x.z, x.w, *reinterpret_cast<const uint32_t *>(lp + 16)};
          uint32_t end[4];
#pragma unroll
          for (int k = 0; k < 4; k++)
            end[k] = smax_ge(w[k], kadd, himode) & smax_gt(w[k], smax_shr_bytes(w[k], w[k + 1], 1)) &
                     ~smax_is255(w[k]);
          em = pack_ends16(end);
          if (o0 + kChunk > valid)           // the shard (or the piece) ends inside this chunk
            em &= (1u << (valid - o0)) - 1u;
        }
        // append the lanes' candidates (order does not matter)
        uint32_t cnt = __popc(em), inc = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
          const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d) inc += y;
        }
        uint32_t slot = ne + inc - cnt;
        while (em)
        {
          ends[slot++] = (uint16_t) (o0 + (__ffs(em) - 1));
          em &= em - 1;
        }
        ne += __shfl_sync(0xffffffffu, inc, 31);
        if (ne > (uint32_t) (kSmallEnds - 32 * 8))
          process_ends();
      }
      process_ends();
    }
    __syncwarp();                          // every survivor of the unit is marked

    // ---------------- K3, first half: the warp turns the unit's bitmaps into entries of the
    // survivor arena, in suffix-array order, and leaves the unit's (records, positions)
    // aggregate for the offset scan (k_offsets).  No unit waits for another one here.
    {
      uint32_t cnt = 0;
      uint64_t wsum = 0;
      const bool any = ws.any != 0 && !(P.debug & 16);
      if (any)
      {
#pragma unroll
        for (int j = 0; j < kBitWords / 32; j++)
        {
          const uint32_t wi = (uint32_t) lane * (kBitWords / 32) + j;
          uint32_t ew = ws.endbits[wi];
          cnt += __popc(ew);
          while (ew)
          {
            const uint32_t o = wi * 32 + (__ffs(ew) - 1);
            ew &= ew - 1;
            wsum += survivor_width(ws, o);
          }
        }
      }
      const uint32_t votes = __ballot_sync(0xffffffffu, cnt != 0);
      uint32_t inc_c = cnt;
      uint64_t inc_w = wsum, unit_base = 0;
      bool fits = true;
      if (votes)
      {
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
          const uint32_t yc = __shfl_up_sync(0xffffffffu, inc_c, d);
          const uint64_t yw = __shfl_up_sync(0xffffffffu, inc_w, d);
          if (lane >= d) { inc_c += yc; inc_w += yw; }
        }
        const uint32_t tot = __shfl_sync(0xffffffffu, inc_c, 31);
        if (tot > chunk_left)
        {
          // a new chunk (a unit with more entries than a chunk holds gets exactly its own)
          const uint32_t take = tot > (uint32_t) kArenaChunk ? tot : (uint32_t) kArenaChunk;
          unsigned long long got = 0;
          if (lane == 31)
            got = atomicAdd((unsigned long long *) &P.result[kResArena], (unsigned long long) take);
          chunk_next = __shfl_sync(0xffffffffu, got, 31);
          chunk_left = take;
        }
        unit_base = chunk_next;
        fits = unit_base + tot <= P.arena_capacity;
        chunk_next += tot;
        chunk_left -= tot;
        if (lane == 31)
        {
          if (!fits)
            P.result[kResOverflow] = 1;         // the host enlarges the arena and scans again
          if (inc_w >> 32)
            P.result[kResError] = 4;            // wider than a shard can be
        }
      }
      if (lane == 31)
      {
        UnitMeta m;
        m.count = inc_c; m.pad = 0; m.wsum = inc_w;
        m.base = fits ? unit_base : ~0ull;
        P.meta[unit] = m;
      }
      if (cnt != 0 && fits)
      {
        uint64_t slot = unit_base + inc_c - cnt;
        uint64_t wpre = inc_w - wsum;
#pragma unroll 1
        for (int j = 0; j < kBitWords / 32; j++)
        {
          const uint32_t wi = (uint32_t) lane * (kBitWords / 32) + j;
          uint32_t ew = ws.endbits[wi];
          while (ew)
          {
            const uint32_t o = wi * 32 + (__ffs(ew) - 1);
            ew &= ew - 1;
            const uint64_t wd = survivor_width(ws, o);
            ArenaEntry e;
            e.unit = unit;
            e.end_off = (uint32_t) toff + o;
            e.width = (uint32_t) wd;
            e.wpre = (uint32_t) wpre;           // (a unit's positions: < 2^32, checked above)
            const uint32_t b = ws.lcp[kHalo + o];
            e.len = b;
            e.len_hi = 0;
            if (b == 255)
            {
              // a large survivor: its record is hinted at per END word
              const uint32_t kk = ws.bigk[o >> 5];
              uint64_t v = kBadValue;
              if (kk != 0)
              {
                const uint2 r = ldg_rec(&P.own.llvc[kt0 + kk - 1]);
                if (r.x == e.end_off)
                  v = rec_value(P, kt0 + kk - 1, r.y);
              }
              if (v == kBadValue)
                v = large_value_of(P, kt0, k1, e.end_off);
              e.len = (uint32_t) v;
              e.len_hi = (uint32_t) (v >> 32);
            }
            P.arena[slot] = e;
            slot++; wpre += wd;
          }
        }
      }
    }
    // the next unit
    if (++unit == unit_end)
    {
      unit = __shfl_sync(0xffffffffu, ticket, 0);
      unit_end = unit + kTicketUnits;
    }
  }

  if (STATS)
  {
    if (stat[0]) atomicAdd((unsigned long long *) &P.result[kResStatCand], (unsigned long long) stat[0]);
    if (stat[1]) atomicAdd((unsigned long long *) &P.result[kResStatCandWidth], (unsigned long long) stat[1]);
    if (stat[2]) atomicAdd((unsigned long long *) &P.result[kResStatLlv], (unsigned long long) stat[2]);
    if (stat[3]) atomicAdd((unsigned long long *) &P.result[kResStatSurvWidth], (unsigned long long) stat[3]);
  }
  // k_offsets may be put in place (it waits for this grid to complete before it reads anything)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // the last warp to leave resets the ticket for the next scan
  if (lane == 0)
  {
    __threadfence();
    const uint32_t done = atomicAdd(&P.ctrl[1], 1u);
    if (done == gridDim.x * kWarps - 1)
    {
      P.ctrl[1] = 0;
      P.ctrl[0] = 0;
    }
  }
}

// ------------------------------------------------------------ K3, second half
// Exclusive prefix of the unit aggregates: one pass, decoupled look-back over blocks of
// kOffsetBlock units (epoch-tagged status words, as cheap as the blocks are uniform).  The
// last block reports the totals and -- multi-GPU -- stores the shard's record count into
// every shard's count array.
__global__ void __launch_bounds__(kOffsetThreads)
k_offsets(const __grid_constant__ ScanParams P)
{
  __shared__ unsigned long long warp_c[kOffsetThreads / 32], warp_w[kOffsetThreads / 32];
  __shared__ unsigned long long blk_c, blk_w;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");       // the detection grid has completed
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t nunits = P.nunits;
  const uint64_t first = (uint64_t) blockIdx.x * kOffsetBlock + (uint64_t) tid * kOffsetItems;
  uint32_t c[kOffsetItems];
  uint64_t w[kOffsetItems];
  uint64_t sc = 0, sw = 0;
#pragma unroll
  for (int k = 0; k < kOffsetItems; k++)
  {
    c[k] = 0; w[k] = 0;
    if (first + k < nunits)
    {
      const UnitMeta m = P.meta[first + k];
      c[k] = m.count; w[k] = m.wsum;
    }
    sc += c[k]; sw += w[k];
  }
  uint64_t ic = sc, iw = sw;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const uint64_t yc = __shfl_up_sync(0xffffffffu, ic, d), yw = __shfl_up_sync(0xffffffffu, iw, d);
    if (lane >= d) { ic += yc; iw += yw; }
  }
  if (lane == 31) { warp_c[warp] = ic; warp_w[warp] = iw; }
  __syncthreads();
  uint64_t ec = ic - sc, ew = iw - sw, tot_c = 0, tot_w = 0;
#pragma unroll
  for (int q = 0; q < kOffsetThreads / 32; q++)
  {
    if (q < warp) { ec += warp_c[q]; ew += warp_w[q]; }
    tot_c += warp_c[q]; tot_w += warp_w[q];
  }
  if (warp == 0)
  {
    const uint32_t blk = blockIdx.x;
    uint64_t *mine = P.status + (uint64_t) kStatusWords * blk;
    if (lane == 0)
      st_pair(mine, pack_status(P.epoch, kStateAggregate, tot_c),
              pack_status(P.epoch, kStateAggregate, tot_w));
    // lane l looks at block base - 1 - l; the nearest block that has published its inclusive
    // prefix ends the walk
    uint64_t exc_c = 0, exc_w = 0;
    int64_t base = (int64_t) blk;
    for (;;)
    {
      const int64_t j = base - 1 - lane;
      uint64_t vc = 0, vw = 0;
      bool is_prefix = true;            // the virtual block -1: prefix 0
      if (j >= 0)
      {
        const uint64_t *st = P.status + (uint64_t) kStatusWords * (uint64_t) j;
        unsigned backoff = 20;
        for (;;)
        {
          uint64_t pa, pb, aa, ab;
          ld_pair(st + 2, pa, pb);
          ld_pair(st, aa, ab);
          if (status_is(pa, P.epoch, kStatePrefix) && status_is(pb, P.epoch, kStatePrefix))
          {
            vc = pa & kValueMask; vw = pb & kValueMask;
            break;
          }
          if (status_is(aa, P.epoch, kStateAggregate) && status_is(ab, P.epoch, kStateAggregate))
          {
            vc = aa & kValueMask; vw = ab & kValueMask; is_prefix = false;
            break;
          }
          __nanosleep(backoff);
          backoff = min(backoff * 2u, 256u);
        }
      }
      const uint32_t pvotes = __ballot_sync(0xffffffffu, is_prefix);
      const int firstp = pvotes ? __ffs(pvotes) - 1 : 32;
      if (lane > firstp) { vc = 0; vw = 0; }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1)
      {
        vc += __shfl_xor_sync(0xffffffffu, vc, d);
        vw += __shfl_xor_sync(0xffffffffu, vw, d);
      }
      exc_c += vc; exc_w += vw;
      if (pvotes)
        break;
      base -= 32;
    }
    if (lane == 0)
    {
      st_pair(mine + 2, pack_status(P.epoch, kStatePrefix, exc_c + tot_c),
              pack_status(P.epoch, kStatePrefix, exc_w + tot_w));
      blk_c = exc_c; blk_w = exc_w;
      if (blk + 1 == gridDim.x)
      {
        P.result[kResCount] = exc_c + tot_c;
        P.result[kResPositions] = exc_w + tot_w;
        // one-sided count exchange: the shard's record count goes straight into every
        // shard's count array (P2P stores over NVLink), tagged with the step
        const uint64_t word = (P.exchange_tag << 40) | ((exc_c + tot_c) & ((1ull << 40) - 1));
        for (int q = 0; q < P.npeers; q++)
          asm volatile("st.release.sys.global.u64 [%0], %1;"
                       :: "l"(P.peer_counts[q] + P.my_rank), "l"(word) : "memory");
        for (int k = 0; k < kResSlots; k++)   // result blocks ping-pong: no memset per scan
          P.result_next[k] = 0;
      }
    }
  }
  __syncthreads();
  ec += blk_c; ew += blk_w;
#pragma unroll
  for (int k = 0; k < kOffsetItems; k++)
    if (first + k < nunits)
    {
      UnitOffset o;
      o.c = ec; o.w = ew;
      P.unitoff[first + k] = o;
      ec += c[k]; ew += w[k];
    }
}

// Every arena entry goes to its place: the record, and the occurrence positions
// suf[lb..lb+width) gathered right behind those of the records before it.  kEmitLanes threads
// per unit (a unit holds a handful of repeats).
__global__ void __launch_bounds__(256)
k_emit(const __grid_constant__ ScanParams P)
{
  asm volatile("griddepcontrol.wait;" ::: "memory");       // the offsets are complete
  const uint64_t nunits = (P.debug & 128) ? 0 : (uint64_t) P.nunits;
  const uint64_t a_lo = P.own.a_lo;
  const uint32_t sub = threadIdx.x % kEmitLanes;
  for (uint64_t u = ((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) / kEmitLanes; u < nunits;
       u += (uint64_t) gridDim.x * blockDim.x / kEmitLanes)
  {
    const UnitMeta m = P.meta[u];
    if (m.count <= sub || m.base == ~0ull)
      continue;
    const UnitOffset uo = P.unitoff[u];
    for (uint32_t i = sub; i < m.count; i += kEmitLanes)
    {
      const ArenaEntry e = P.arena[m.base + i];
      const uint64_t dst = uo.c + i, po = uo.w + e.wpre;
      const uint64_t end = a_lo + e.end_off, wd = e.width;
      if (wd < 2 || wd > end + 1) { P.result[kResError] = 2; continue; }
      const uint64_t lb = end + 1 - wd;
      if (dst < P.rec_capacity)
      {
        smax_record r;
        r.len = ((uint64_t) e.len_hi << 32) | e.len;
        r.lb = lb; r.width = wd;
        P.recs[dst] = r;
      } else
        P.result[kResOverflow] = 1;
      if (P.positions != nullptr)
      {
        if (po + wd > P.pos_capacity)
          P.result[kResOverflow] = 1;
        else if (wd <= 4 && lb >= a_lo && P.own.suf != nullptr)
        {
          // the common case: all (<= 4) scattered suftab reads in flight together
          const uint64_t so = lb - a_lo;
          uint64_t v[4];
#pragma unroll
          for (int q = 0; q < 4; q++)
            v[q] = (uint64_t) q >= wd ? 0
                   : P.sufbytes == 8 ? reinterpret_cast<const uint64_t *>(P.own.suf)[so + q]
                                     : (uint64_t) reinterpret_cast<const uint32_t *>(P.own.suf)[so + q];
#pragma unroll
          for (int q = 0; q < 4; q++)
            if ((uint64_t) q < wd)
              P.positions[po + q] = v[q];
        } else
          for (uint64_t q = 0; q < wd; q++)
            P.positions[po + q] = suf_at(P, lb + q);
      }
    }
  }
}

// ------------------------------------------------ upload-time .llv tables
__global__ void k_llvdir(const smax_llv *llv, uint64_t nllv, uint64_t a_lo,
                         uint32_t *dir, uint64_t nentries)
{
  const uint64_t b = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nentries)
    return;
  const uint64_t target = a_lo + (b << kLlvBucketShift);
  uint64_t lo = 0, hi = nllv;
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (llv[mid].position < target) lo = mid + 1; else hi = mid;
  }
  dir[b] = (uint32_t) lo;
}

// compact form of the shard's .llv records: {position - a_lo, value} in 8 bytes
// (a shard holds < 2^32 entries; a value that does not fit reads kLlvEscape and is
// taken from the 16-byte record; *has_escape tells the scan whether there is one)
__global__ void k_llvpack(const smax_llv *llv, uint64_t nllv, uint64_t a_lo, uint2 *out,
                          uint32_t *has_escape)
{
  const uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nllv)
  {
    if (k < nllv + kLlvPad)
      out[k] = make_uint2(kNoRecord, 0);            // "no record" behind the last one (16-byte loads)
    return;
  }
  const smax_llv r = llv[k];
  if (r.value >= (uint64_t) kLlvEscape)
    *has_escape = 1;
  out[k] = make_uint2((uint32_t) (r.position - a_lo),
                      r.value < (uint64_t) kLlvEscape ? (uint32_t) r.value : kLlvEscape);
}

// per-unit directory of the scan: dir[u] = first record at or behind the start of unit u of
// the shard's own range [g_lo, g_hi) (unit = kUnitBytes entries), dir[nunits] = first record
// at or behind g_hi
__global__ void k_unitdir(const smax_llv *llv, uint64_t nllv, uint64_t g_lo, uint64_t g_hi,
                          uint32_t *dir, uint64_t ntiles)
{
  const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t > ntiles)
    return;
  const uint64_t target = t < ntiles ? g_lo + t * kUnitBytes : g_hi;
  uint64_t lo = 0, hi = nllv;
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (llv[mid].position < target) lo = mid + 1; else hi = mid;
  }
  dir[t] = (uint32_t) lo;
}

// histogram of the lcp bytes of the shard's own range: what share of the entries reaches a
// minimum length tells the host which scan kernel suits the index (smax_device.cu: pick_kernel)
__global__ void __launch_bounds__(256)
k_lcphist(const uint8_t *lcp, uint64_t len, unsigned long long *hist)
{
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint64_t nchunks = len / 16;
  for (uint64_t c = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; c < nchunks;
       c += (uint64_t) gridDim.x * blockDim.x)
  {
    const uint4 x = __ldg(reinterpret_cast<const uint4 *>(lcp) + c);
    const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int k = 0; k < 4; k++)
#pragma unroll
      for (int b = 0; b < 4; b++)
        atomicAdd(&h[(w[k] >> (8 * b)) & 255u], 1u);
  }
  if (blockIdx.x == 0 && threadIdx.x < (len & 15))      // the ragged tail
    atomicAdd(&h[lcp[nchunks * 16 + threadIdx.x]], 1u);
  __syncthreads();
  if (h[threadIdx.x] != 0)
    atomicAdd(&hist[threadIdx.x], (unsigned long long) h[threadIdx.x]);
}

// --------------------------------------------------------------- launchers
cudaError_t launch_lcphist(const uint8_t *lcp, uint64_t len, unsigned long long *hist, int sm_count,
                           cudaStream_t st)
{
  k_lcphist<<<(unsigned) (sm_count > 0 ? sm_count : 1) * 4, 256, 0, st>>>(lcp, len, hist);
  return cudaGetLastError();
}

cudaError_t launch_llvdir(const smax_llv *llv, uint64_t nllv, uint64_t a_lo,
                          uint32_t *dir, uint64_t nentries, cudaStream_t st)
{
  const int threads = 256;
  const uint64_t blocks = (nentries + threads - 1) / threads;
  k_llvdir<<<(unsigned) blocks, threads, 0, st>>>(llv, nllv, a_lo, dir, nentries);
  return cudaGetLastError();
}

cudaError_t launch_llvpack(const smax_llv *llv, uint64_t nllv, uint64_t a_lo, uint2 *out,
                           uint32_t *has_escape, cudaStream_t st)
{
  const int threads = 256;
  const uint64_t blocks = (nllv + kLlvPad + threads - 1) / threads;
  k_llvpack<<<(unsigned) blocks, threads, 0, st>>>(llv, nllv, a_lo, out, has_escape);
  return cudaGetLastError();
}

cudaError_t launch_unitdir(const smax_llv *llv, uint64_t nllv, uint64_t g_lo, uint64_t g_hi,
                           uint32_t *dir, uint64_t ntiles, cudaStream_t st)
{
  const int threads = 256;
  const uint64_t blocks = (ntiles + 1 + threads - 1) / threads;
  k_unitdir<<<(unsigned) blocks, threads, 0, st>>>(llv, nllv, g_lo, g_hi, dir, ntiles);
  return cudaGetLastError();
}

static const void *scan_kernel(bool stats)
{
  return stats ? (const void *) k_scan<true> : (const void *) k_scan<false>;
}

// One scan = three launches on the stream: detection (tiles are taken from a ticket; no tile
// depends on another one, so any grid size makes progress), the offset scan over the unit
// aggregates, and the emit of the arena entries.
cudaError_t launch_scan(const ScanParams &p, bool stats, int grid, int sm_count, cudaStream_t st)
{
  void *args[] = {(void *) &p};
  cudaError_t e = cudaLaunchKernel(scan_kernel(stats), dim3(grid), dim3(kThreads), args,
                                            kWarps * sizeof(WarpSmem), st);
  if (e != cudaSuccess)
    return e;
  // the two small kernels are launched programmatically dependent: they are put in place while
  // the grid before them drains and wait (griddepcontrol.wait) for its completion
  const uint64_t nunits = p.nunits;
  const unsigned blocks = (unsigned) ((nunits + kOffsetBlock - 1) / kOffsetBlock);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(blocks > 0 ? blocks : 1);
  cfg.blockDim = dim3(kOffsetThreads);
  e = cudaLaunchKernelEx(&cfg, k_offsets, p);
  if (e != cudaSuccess)
    return e;
  {
    const uint64_t want = (nunits * kEmitLanes + 255) / 256;
    const uint64_t most = (uint64_t) (sm_count > 0 ? sm_count : 1) * 8;
    cfg.gridDim = dim3((unsigned) (want < 1 ? 1 : (want > most ? most : want)));
  }
  cfg.blockDim = dim3(256);
  return cudaLaunchKernelEx(&cfg, k_emit, p);
}

int scan_blocks_per_sm(bool stats)
{
  const void *fn = scan_kernel(stats);
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int) (kWarps * sizeof(WarpSmem))) != cudaSuccess)
    return 0;
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, kThreads, kWarps * sizeof(WarpSmem)) !=
          cudaSuccess || n <= 0)
    return 0;
  return n;
}

}  // namespace smax
