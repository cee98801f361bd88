/*
  smax_scan.cu -- hand-written sm_100a kernels of the supermaximal-repeat
  scan ("unit kernel").  ONE fused pass over the lcptab replaces the
  reference's stack sweep (/root/reference/src/match/esa-bottomup.c:116-273)
  and its per-node left-character bookkeeping
  (/root/reference/src/match/esa-maxpairs.c:181-360).

  k_scan: CTAs of 4 INDEPENDENT warps; a warp takes a unit of 4096 lcptab
  entries at a time -- heaviest units first (k_unitorder_*, built at upload),
  the first 5/8 of that order dealt round-robin, the rest from a ticket -- and
  never waits for another warp: no cooperative launch, any grid size is
  correct.  Per unit:

    feed  one TMA bulk copy (cp.async.bulk.shared::cluster.global, mbarrier
          complete_tx) of the unit's lcp bytes + a 16-byte halo either side.
          While it is in flight the warp works on the unit's large values.
          The next unit's directory words / the ticket after it arrive by
          cp.async in shared memory, its first records are asked into L2.
    K1a   large values (byte 255) in .llv RECORD space, out of the compact
          records built at upload (k_llvpack: value + PEAK / GENERAL flags,
          position - a_lo; the unit's records are found through k_unitdir's
          per-unit directory): K1 is one compare per record.  PEAK records
          (SA width 2) and GENERAL ones (runs of equal values, values that do
          not fit, the shard's edge) are compacted per warp (ballot) into two
          lists that are worked off one record per lane.
    K1b   small values, flat and bit-parallel (smax_swar.h): every lane filters
          its 16-byte chunks for a byte >= minlength; the hits, compacted per
          warp, are classified with SWAR byte arithmetic: ENDs of runs that
          fall to a smaller value and whose last two left characters differ;
          one END per lane then: entered from a smaller value 1, 2 or 3
          entries back (SA width 2, 3, 4)?  Runs of >= 4 equal values are
          walked together with their left characters (K2 ends the walk at the
          first repeated character, so a wide plateau costs O(alphabet)).
    K2    left-distinctness, bit-parallel on the chunk's bwt words (fetched
          from global memory only for chunks that passed the filter) for widths
          <= 4; a 256-bit alphabet mask otherwise.  Specials (>= 254) never
          collide under the GenomeTools convention (esa-maxpairs.c:24-31).
    K3    first half.  A survivor [lb, e] sets bit e of the unit's END bitmap
          and bit lb of its START bitmap in shared memory (supermaximal repeats
          are disjoint SA intervals, so the two bitmaps describe them
          completely, in order, at any density).  At the end of the unit the
          warp lists the ENDs in order, one entry per lane goes to the survivor
          arena, and the unit's (repeats, occurrences) aggregate is left for
          k_emit and added to the sums of the unit's block of 32 units.

  k_emit (second launch, programmatically dependent): offsets from the block
  sums + a scan over the block's units; one thread per arena entry writes the
  record in suffix-array order and gathers suf[lb..lb+width) behind it.

  k_llvdir / k_llvpack / k_unitdir / k_lcphist / k_unitorder_* build the
  directories, the compact records and the unit order at upload time.
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
#if defined(SMAX_L1_LEFT) && SMAX_L1_LEFT
  asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#else
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
#endif
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

// 4-byte asynchronous copy global -> shared (no register waits for the word)
__device__ __forceinline__ void cp_async4(void *dst, const void *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(smem_u32(dst)), "l"(src) : "memory");
}

__device__ __forceinline__ void cp_async_wait_all()
{
  asm volatile("cp.async.wait_all;" ::: "memory");
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

// compact record k of the own shard: .x = position - a_lo, .y = value (kLlvEscape: does not fit)
__device__ __forceinline__ uint2 rec_at(const ScanParams &P, uint32_t k)
{
  return make_uint2(__ldg(P.own.llvp + k), __ldg(P.own.llvv + k) & kLlvValueMask);
}

// value of the own record k given its compact form
__device__ __forceinline__ uint64_t rec_value(const ScanParams &P, uint32_t k, uint32_t v30)
{
  return v30 != kLlvEscape ? (uint64_t) v30 : P.own.llv[k].value;
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
// value, or -- !FULL -- a left character repeats), with kDistinct set when K2 holds
// (results come back in a register: an out-parameter would live in local memory).
constexpr uint64_t kDistinct = 1ull << 63;
template <bool FULL>
__device__ __noinline__ uint64_t small_run_plateau(const ScanParams &P, uint64_t e, uint32_t b)
{
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  const uint64_t a_lo = P.own.a_lo;
  CharSet cs;
  bool dup = cs.add(bwt_at(P, e), gt_policy);
  dup |= cs.add(bwt_at(P, e - 1), gt_policy);
  uint64_t s = e - 1;
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
  return (e - s + 2) | (dup ? 0 : kDistinct);
}

// The general path for the large value of the own record k (a value that does not fit the
// compact record, the end of a run of EQUAL large values, the record on the shard's edge):
// does a local-maximum plateau end at it?  The run is walked in record space, into the left
// neighbours through value_at when it leaves the own arrays.  Returns the SA width (0: none,
// or -- !FULL -- a left character repeats), with kDistinct set when K2 holds.
template <bool FULL>
__device__ __noinline__ uint64_t large_plateau(const ScanParams &P, uint32_t k)
{
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  const uint64_t a_lo = P.own.a_lo;
  const uint32_t nllv = (uint32_t) P.own.nllv;
  const uint2 r = rec_at(P, k);
  const uint64_t p = a_lo + r.x, v = rec_value(P, k, r.y);
  if (v < P.minlength)
    return 0;
  if (k + 1 < nllv)
  {
    // not the end of a plateau when the next entry holds a value >= v
    const uint2 nx = rec_at(P, k + 1);
    if (nx.x == r.x + 1 && rec_value(P, k + 1, nx.y) >= v)
      return 0;
  }
  CharSet cs;
  bool dup = cs.add(bwt_at(P, p), gt_policy);
  uint64_t s = p;
  uint32_t kk = k;
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
      const uint2 pr = rec_at(P, kk - 1);
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
  return (p - s + 2) | (dup ? 0 : kDistinct);
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
  union
  {
    struct
    {
      alignas(16) uint8_t chunklist[kUnitChunks]; // chunks that passed the filter
      uint16_t slowlist[kSlowList];               // records of the large pass that go the general way
    };
    // K3 (the lists above are through by then): bit o of bigbits = entry o of the unit is a large
    // value (byte 255), bigrank[l] = large values before entry 128 l: the r-th large value of the
    // unit is the r-th .llv record of the unit
    struct
    {
      alignas(16) uint32_t bigbits[kBitWords];
      uint16_t bigrank[32];
    };
  };
  union
  {
    uint32_t candlist[kCandList];             // large pass: records that end a plateau of SA width 2 (unit offsets)
    uint16_t endlist[kEndList];               // small pass: ENDs whose last two left characters differ;
                                              //   K3: the unit's ENDs in order
  };
  uint32_t anybig;                        // a large survivor ends in the unit
  unsigned long long chunk_next;          // next free entry of the warp's arena chunk,
  uint32_t chunk_left;                    //   entries left in it
  uint32_t pipe_dir[2];                   // the next unit's .llv records [first, behind the last)
  uint32_t pipe_next;                     // the unit after the next one (>= nunits: none)
#if SMAX_PROBE
  unsigned long long probe[8];            // tuning build: nanoseconds per phase, summed over the warp's units
  uint32_t uphase[8];                     //   ... and of the unit at hand
#endif
  unsigned long long open_width;          // width of the survivor that starts left of the unit
  uint32_t count;                         // survivors of the unit,
  uint32_t wsum;                          //   sum of their widths (without the open one)
  alignas(8) uint64_t ready;              // mbarrier: the unit's lcp bytes have landed
};

static_assert(kMinBlocks * (kWarps * sizeof(WarpSmem) + 1024) <= 227 * 1024 || kMinBlocks > 8,
              "the warps' shared memory must leave room for kMinBlocks CTAs per SM");

// K3, first half: the survivor [e + 1 - width, e] (e at unit offset o) is marked
// in the unit's bitmaps and counted.
__device__ __forceinline__ void mark_survivor(WarpSmem &ws, uint32_t o, uint64_t width)
{
  atomicOr(&ws.endbits[o >> 5], 1u << (o & 31));
  if (width <= (uint64_t) o + 1)
  {
    const uint32_t lb = o + 1 - (uint32_t) width;
    atomicOr(&ws.startbits[lb >> 5], 1u << (lb & 31));
    atomicAdd(&ws.wsum, (uint32_t) width);
  } else
    ws.open_width = width;                // starts left of the unit: at most one per unit
  atomicAdd(&ws.count, 1u);
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


// ---- second-level passes over a warp's lists (inlined: as real functions -- one copy of each
// in the kernel -- they measured 6 % slower on C2).  stat: [0] candidate plateaus, [1] their
// widths (statistics build only).

// the general path for the listed .llv records (numbers relative to kt0)
template <bool STATS>
__device__ __forceinline__ void run_slow_list(const ScanParams &P, WarpSmem &ws, uint32_t ns, uint32_t kt0,
                                           uint32_t toff32, int lane, uint64_t *stat)
{
  __syncwarp();
#pragma unroll 1
  for (uint32_t i = lane; i < ns; i += 32)
  {
    const uint32_t k = kt0 + ws.slowlist[i];
    const uint64_t wd = large_plateau<STATS>(P, k);
    const uint64_t width = wd & ~kDistinct;
    if (STATS) atomicAdd((unsigned long long *) &P.result[kResWalks], 1ull);
    if (width != 0)
    {
      if (STATS) { stat[0]++; stat[1] += width; }
      if (wd & kDistinct)
      {
        mark_survivor(ws, __ldg(P.own.llvp + k) - toff32, width);
        ws.anybig = 1;
      }
    }
  }
  __syncwarp();
}

// K2 of the listed large-value candidates of SA width 2 (unit offsets): one per lane
__device__ __forceinline__ void run_cand_list(const ScanParams &P, WarpSmem &ws, uint32_t nc, uint64_t toff,
                                           int lane)
{
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  __syncwarp();
#pragma unroll 1
  for (uint32_t i = lane; i < nc; i += 32)
  {
    const uint32_t o = ws.candlist[i];
    const uint8_t *bp = P.own.bwt + toff + o;
    const uint32_t b0 = bp[-1], b1 = bp[0];
    if (b0 != b1 || (gt_policy && b0 >= 254))
    {
      mark_survivor(ws, o, 2);
      ws.anybig = 1;
    }
  }
  __syncwarp();
}

// phase B, second level: one listed END of small values per lane.  K1: is its run entered from a
// smaller value?  SA width 2, 3, 4 out of the staged bytes, longer runs are walked; K2 on the
// left characters.
template <bool STATS>
__device__ __forceinline__ void run_end_list(const ScanParams &P, WarpSmem &ws, uint32_t ne, uint64_t toff,
                                          int lane, uint64_t *stat)
{
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  const uint32_t lim = gt_policy ? 254u : 256u;      // specials never collide (GT policy)
  __syncwarp();
#pragma unroll 1
  for (uint32_t i = lane; i < ne; i += 32)
  {
    const uint32_t o = ws.endlist[i];
    const uint8_t *lp = ws.lcp + kHalo + o;
    // the left characters bwt[o - 3 .. o] as one word (byte 3 = bwt[o]): requested first
    const uint64_t g = toff + o;
    const uint8_t *bp = P.own.bwt + (g & ~3ull);
    const uint32_t bhi = __ldg(reinterpret_cast<const uint32_t *>(bp));
    const uint32_t blo = g >= 4 ? __ldg(reinterpret_cast<const uint32_t *>(bp - 4)) : 0u;
    const uint32_t v = lp[0], l1 = lp[-1];
    if (l1 > v)
      continue;                        // entered from a larger value (255 stands for one)
    const uint32_t l2 = lp[-2], l3 = lp[-3];
    const uint32_t sh = 8 * (((uint32_t) g & 3u) + 1u);
    const uint32_t cw = sh == 32 ? bhi : __funnelshift_r(blo, bhi, sh);
    const uint32_t c0 = cw >> 24, c1 = (cw >> 16) & 255u, c2 = (cw >> 8) & 255u, c3 = cw & 255u;
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
        const uint64_t wd = small_run_plateau<STATS>(P, P.own.a_lo + g, v);
        width = wd & ~kDistinct;
        ok = (wd & kDistinct) != 0;
        if (STATS) atomicAdd((unsigned long long *) &P.result[kResWalks], 1ull);
      } else
        width = 0;                     // entered from a larger value further left
    }
    if (width != 0)
    {
      if (STATS) { stat[0]++; stat[1] += width; }
      if (ok)
        mark_survivor(ws, o, width);
    }
  }
  __syncwarp();
}

// left edge of the shard's arrays: the halo of the first unit comes from the left neighbour
// shard, or repeats the first entry (which sends every plateau that touches the edge into the
// walk that reports the missing range); the table itself starts with lcp[0] = 0
__device__ __noinline__ void fill_left_halo(const ScanParams &P, WarpSmem &ws, uint32_t nbytes)
{
  const uint64_t a_lo = P.own.a_lo;
  for (uint32_t i = 0; i < nbytes; i++)
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

#ifndef SMAX_UNROLL_A
#define SMAX_UNROLL_A 1
#endif
constexpr int kUnrollA = SMAX_UNROLL_A;     // unroll factor of the filter loop (1: the rolled loop measured fastest, 93 vs 95 / 99 us on C2 for 2-4 / 8)

// ------------------------------------------------------------ scan kernel
template <bool STATS>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_scan(const __grid_constant__ ScanParams P)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // (volatile reads: the compiler keeps these in registers instead of re-reading the special
  // registers -- tens of cycles each -- wherever a register is short)
  int lane, warp;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(warp));
  warp >>= 5;
  WarpSmem &ws = reinterpret_cast<WarpSmem *>(smem_raw)[warp];
  uint32_t lt_mask;
  asm volatile("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
  const uint64_t a_lo = P.own.a_lo;
  const uint64_t base_off = P.g_lo - a_lo;                         // multiple of 16
  // table bytes that may be read: the arrays are zero padded (SMAX_PAD)
  const uint64_t readable = ((P.own.a_hi - a_lo + 15) & ~15ull) + 48;
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  const uint32_t m30 = (uint32_t) min(P.minlength, (uint64_t) kLlvEscape);
  uint32_t kadd; int himode;
  smax_ge_consts(P.mb, &kadd, &himode);
  uint64_t stat0 = 0, stat1 = 0, stat2 = 0, stat3 = 0;    // candidates, their widths, .llv records inspected, survivor widths

  // the unit's bitmaps, hints and counters start out clean and are cleaned after a unit that used them
  {
    uint4 *z = reinterpret_cast<uint4 *>(ws.endbits);
    const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int i = lane; i < (int) ((2 * kBitWords * 4) / 16); i += 32)
      z[i] = zero;
  }
  if (lane == 0)
  {
    ws.count = 0; ws.wsum = 0; ws.open_width = 0; ws.anybig = 0;
    ws.chunk_next = 0; ws.chunk_left = 0;
#if SMAX_PROBE
    for (int i = 0; i < 8; i++) ws.probe[i] = 0;
#endif
    mbar_init(&ws.ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // the survivor arena is handed out in chunks of kArenaChunk entries (one atomic per chunk and
  // warp: a unit's entries are consecutive, a warp fills its chunk unit by unit; the chunk's
  // state lives in shared memory, where it costs no register)
  // The first 5/8 of the units are dealt out round-robin (warp g takes g, g + W, g + 2 W, ...:
  // nobody waits for anything), the rest is taken from a ticket one unit at a time, which evens
  // out what the warps' units differed by.  (The round trip of a ticket cannot be hidden: the
  // compiler aggregates the atomics of a warp and shuffles the result out at once.)
  const uint32_t nwarps = gridDim.x * kWarps, gwarp = blockIdx.x * kWarps + warp;
#ifndef SMAX_STATIC_EIGHTHS
#define SMAX_STATIC_EIGHTHS 5
#endif
  const uint32_t rounds = (uint32_t) (((uint64_t) P.nunits * SMAX_STATIC_EIGHTHS / 8) / nwarps);   // static units per warp
  const uint32_t first_dynamic = rounds * nwarps;
  // The units of a warp come through a small pipeline, driven by lane 0: while unit u is worked
  // on, the directory words of unit u + 1 and the number of unit u + 2 (a ticket, looked up in
  // P.unitorder: tickets number the units heaviest first) are fetched with cp.async straight
  // into shared memory, and the first .llv records of unit u + 1 are asked into L2 -- the chain
  // ticket -> unit -> directory -> records is four dependent round trips that nothing waits for.
  const uint32_t *uorder = P.unitorder;
  const uint32_t nunits = P.nunits;
  const bool have_llv = P.own.nllv != 0;
  uint32_t taken = 2;                   // tickets the warp has taken so far
  uint32_t unit, next;
  {
    uint32_t t0 = nunits, t1 = nunits;
    if (lane == 0)
    {
      t0 = 0 < rounds ? gwarp : first_dynamic + atomicAdd(&P.ctrl[0], 1u);
      t1 = 1 < rounds ? nwarps + gwarp : first_dynamic + atomicAdd(&P.ctrl[0], 1u);
      if (uorder != nullptr)
      {
        if (t0 < nunits) t0 = __ldg(uorder + t0);
        if (t1 < nunits) t1 = __ldg(uorder + t1);
      }
    }
    unit = __shfl_sync(0xffffffffu, t0, 0);
    next = __shfl_sync(0xffffffffu, t1, 0);
  }
  // the unit's .llv records [kt0, k1) (exact: the per-unit directory built at upload)
  uint32_t kt0 = 0, k1 = 0;
  if (have_llv && unit < nunits)
  {
    kt0 = P.unitdir[unit];
    k1 = P.unitdir[unit + 1];
  }
  __syncwarp();

  while (unit < nunits)
  {
    const uint64_t toff = base_off + (uint64_t) unit * kUnitBytes;
    const uint32_t toff32 = (uint32_t) toff;         // (a shard holds < 2^32 entries)
#if SMAX_PROBE
    uint64_t t_begin, t_last;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
    t_last = t_begin;
#define SMAX_PHASE(I) { __syncwarp(); uint64_t t_now; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now)); \
                        if (lane == 0) { ws.probe[I] += t_now - t_last; ws.uphase[I] = (uint32_t) (t_now - t_last); } t_last = t_now; }
#else
#define SMAX_PHASE(I)
#endif
    const uint64_t unit_lo = a_lo + toff;
    uint32_t ticket = nunits;
    if (lane == 0)
    {
      const Feed f = feed_of(toff, readable);
      if (f.dst != 0)
        fill_left_halo(P, ws, f.dst);
      // (the stage was read through the generic proxy: order those reads before the bulk copy)
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&ws.ready, f.bytes);
      tma_load(ws.lcp + f.dst, P.own.lcp + f.src, f.bytes, &ws.ready);
      // pipeline: the directory words of the next unit, the ticket of the one after it
      if (next < nunits)
      {
        if (have_llv)
        {
          cp_async4(&ws.pipe_dir[0], P.unitdir + next);
          cp_async4(&ws.pipe_dir[1], P.unitdir + next + 1);
        }
        ticket = taken < rounds ? taken * nwarps + gwarp : first_dynamic + atomicAdd(&P.ctrl[0], 1u);
      }
    }
    taken++;
    __syncwarp();
    // ends at or beyond g_hi belong to the next shard
    const uint32_t valid = (uint32_t) min((uint64_t) kUnitBytes, P.g_hi - unit_lo);

    // ---------------- K1a: large values, in record space (the lcp bytes are in flight).
    // Four records per lane and step (the loads of the next step are issued before this one is
    // worked on); the run flags of the compact records make the neighbour tests value
    // compares.  A record whose value rises from its predecessor and falls to its successor
    // ends a plateau of SA width 2: such records are collected (their left characters are
    // asked for and compared further down, K2).  Everything else that may end a plateau (a run
    // of equal values, a value that does not fit, the shard's edge) goes through the general path.
    uint32_t *cands = ws.candlist;
    uint16_t *slow = ws.slowlist;
    uint32_t nc = 0, ns = 0;
    const uint32_t *pp = P.own.llvp;
    uint64_t stat01[2] = {0, 0};            // (statistics build: what the list passes count)
    auto run_slow = [&]() { run_slow_list<STATS>(P, ws, ns, kt0, toff32, lane, STATS ? stat01 : nullptr); ns = 0; };
    auto run_cands = [&]() { run_cand_list(P, ws, nc, toff, lane); nc = 0; };
    const bool large = kt0 < k1 && !(P.debug & 4);
    const uint32_t nllv = (uint32_t) P.own.nllv;
    const uint32_t *vv = P.own.llvv;
    const uint4 none4 = make_uint4(0, 0, 0, 0);
    uint32_t kb = kt0 & ~3u;
    uint4 qn = none4, pn = none4;
    if (large)
    {
      const uint32_t k = kb + 4 * lane;
      if (k < nllv)
      {
        qn = __ldg(reinterpret_cast<const uint4 *>(vv + k));               // (padded behind nllv)
        pn = __ldg(reinterpret_cast<const uint4 *>(pp + k));
      }
    }
    if (large)
    {
#pragma unroll 1
      for (; kb < k1; kb += kLlvRow)
      {
        const uint32_t k = kb + 4 * lane;
        const uint4 q = qn, pz = pn;
        if (kb + kLlvRow < k1)
        {
          const uint32_t k2 = k + kLlvRow;
          qn = none4;
          if (k2 < nllv)
          {
            qn = __ldg(reinterpret_cast<const uint4 *>(vv + k2));
            pn = __ldg(reinterpret_cast<const uint4 *>(pp + k2));
          }
        }
        const uint32_t c[4] = {q.x, q.y, q.z, q.w};
        const uint32_t ps[4] = {pz.x, pz.y, pz.z, pz.w};
        uint32_t c2 = 0, sl = 0;          // bit j: record k + j ends a width-2 plateau / goes the general way
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
          const bool inr = k + j >= kt0 && k + j < k1;
          const bool big = inr && (c[j] & kLlvValueMask) >= m30;
          c2 |= big && (c[j] & kLlvPeak) ? 1u << j : 0u;
          sl |= big && (c[j] & kLlvGeneral) ? 1u << j : 0u;
          if (STATS) { stat2 += inr ? 1 : 0; if (big && (c[j] & kLlvPeak)) { stat0++; stat1 += 2; } }
        }
        // append the lanes' candidates to the lists (order does not matter)
        if (__any_sync(0xffffffffu, c2 != 0))
        {
#pragma unroll
          for (int j = 0; j < 4; j++)
          {
            const uint32_t votes = __ballot_sync(0xffffffffu, (c2 >> j & 1u) != 0);
            if (c2 >> j & 1u)
            {
              const uint32_t o = ps[j] - toff32;
              cands[nc + __popc(votes & lt_mask)] = o;
              asm volatile("prefetch.global.L2 [%0];" :: "l"(P.own.bwt + toff + o - 1));
            }
            nc += __popc(votes);
          }
          if (nc > (uint32_t) (kCandList - kLlvRow))
            run_cands();
        }
        if (__any_sync(0xffffffffu, sl != 0))
        {
#pragma unroll
          for (int j = 0; j < 4; j++)
          {
            const uint32_t votes = __ballot_sync(0xffffffffu, (sl >> j & 1u) != 0);
            if (sl >> j & 1u)
              slow[ns + __popc(votes & lt_mask)] = (uint16_t) (k + j - kt0);
            ns += __popc(votes);
          }
          if (ns > (uint32_t) (kSlowList - kLlvRow))
            run_slow();
        }
      }
    }

    SMAX_PHASE(0)
    // ---------------- K1b + K2: small values out of the staged unit
    mbar_wait(&ws.ready, (taken + 1u) & 1u);   // (one bulk copy per unit: the barrier's phase follows the unit count; taken is 3 in the first unit)
    SMAX_PHASE(1)
    const bool small = P.minlength < 255 && !(P.debug & 2);
    uint8_t *list = ws.chunklist;
    uint32_t n = 0;
    if (small)
    {
      // phase A: which of the unit's 256 chunks hold a byte >= the threshold?
#pragma unroll kUnrollA
      for (int j = 0; j < kUnitChunks / 32; j++)
      {
        const uint32_t c = (uint32_t) (j * 32 + lane), o0 = c * kChunk;
        const uint4 x = *reinterpret_cast<const uint4 *>(ws.lcp + kHalo + o0);
        const bool hit = (smax_ge(x.x, kadd, himode) | smax_ge(x.y, kadd, himode) | smax_ge(x.z, kadd, himode) |
                          smax_ge(x.w, kadd, himode)) != 0 && o0 < valid;
        const uint32_t votes = __ballot_sync(0xffffffffu, hit);
        if (hit)
        {
          list[n + __popc(votes & lt_mask)] = (uint8_t) c;
          // the chunk's left characters will be wanted: start them on their way
          asm volatile("prefetch.global.L2 [%0];" :: "l"(P.own.bwt + toff + o0));
        }
        n += __popc(votes);
      }
    }
    // pipeline: the ticket has arrived -> the number of the unit after the next one is fetched; the
    // next unit's directory words have landed -> its first records are asked into L2
    if (lane == 0)
    {
      cp_async_wait_all();
      if (ticket < nunits && uorder != nullptr)
        cp_async4(&ws.pipe_next, uorder + ticket);
      else
        ws.pipe_next = ticket;
    }
    __syncwarp();
    if (have_llv && next < nunits && lane < 16)
    {
      const uint32_t kn0 = ws.pipe_dir[0] & ~3u, kn1 = ws.pipe_dir[1];
      const uint32_t kq = kn0 + (uint32_t) (lane & 7) * 32u;              // 128 bytes of records per lane
      if (kq < kn1)
        asm volatile("prefetch.global.L2 [%0];" :: "l"((lane < 8 ? P.own.llvv : P.own.llvp) + kq));
    }
    SMAX_PHASE(2)
    // K2 of the large values (their left characters have been on their way since K1a)
    if (nc != 0)
      run_cands();
    if (ns != 0)
      run_slow();
    SMAX_PHASE(3)
    if (small)
    {
      __syncwarp();
      if (P.debug & 8)
        n = 0;
      // phase B, first level: one listed chunk per lane, bit-parallel (smax_swar.h).  Which entries
      // >= minlength fall to a smaller value (the END of a run) AND differ from their predecessor in
      // the left character?  Every supermaximal repeat ends at such an entry, and in repeat-rich
      // regions -- where the entries >= minlength are -- neighbouring suffixes mostly share their left
      // character, so few ENDs remain.  (The statistics build keeps all ENDs: it counts the
      // candidate plateaus.)
      uint16_t *ends = ws.endlist;
      uint32_t ne = 0;
      auto run_ends = [&]() { run_end_list<STATS>(P, ws, ne, toff, lane, STATS ? stat01 : nullptr); ne = 0; };
      // (the left characters of the next round's chunks are asked for before this round is worked on:
      // bwt[o0 - 4 .. o0 + 16) of the lane's chunk, in registers)
      uint4 bnx = make_uint4(0, 0, 0, 0);
      uint32_t bnp = 0;
#define SMAX_LOAD_LEFT(I)                                                                      \
      if (!STATS && (I) < n)                                                                   \
      {                                                                                        \
        const uint32_t lo0 = (uint32_t) list[(I)] * kChunk;                                    \
        const uint8_t *lbp = P.own.bwt + toff + lo0;                                           \
        bnx = ldg_chunk(lbp);                                                                  \
        bnp = toff + lo0 >= 4 ? __ldg(reinterpret_cast<const uint32_t *>(lbp - 4)) : 0u;       \
      }
      SMAX_LOAD_LEFT(lane)
#pragma unroll 1
      for (uint32_t i0 = 0; i0 < n; i0 += 32)
      {
        const uint32_t i = i0 + lane;
        uint32_t em = 0;                     // bit j: a remaining END at byte j of the lane's chunk
        uint32_t o0 = 0;
        const uint32_t b[5] = {bnp, bnx.x, bnx.y, bnx.z, bnx.w};
        SMAX_LOAD_LEFT(i + 32)
        if (i < n)
        {
          o0 = (uint32_t) list[i] * kChunk;
          const uint8_t *lp = ws.lcp + kHalo + o0;
          const uint4 x = *reinterpret_cast<const uint4 *>(lp);
          const uint32_t w[5] = {x.x, x.y, x.z, x.w, *reinterpret_cast<const uint32_t *>(lp + 16)};
          uint32_t end[4];
#pragma unroll
          for (int k = 0; k < 4; k++)
          {
            end[k] = smax_ge(w[k], kadd, himode) & smax_gt(w[k], smax_shr_bytes(w[k], w[k + 1], 1)) &
                     ~smax_is255(w[k]);
            if (!STATS)
            {
              const uint32_t b0 = b[k + 1];
              end[k] &= smax_pair_ok(b0, smax_shl_bytes(b[k], b0, 1), gt_policy ? smax_special(b0) : 0u);
            }
          }
          em = pack_ends16(end);
          if (o0 + kChunk > valid)           // the shard ends inside this chunk
            em &= (1u << (valid - o0)) - 1u;
        }
        if (__any_sync(0xffffffffu, em != 0))
        {
          // append the lanes' ENDs (order does not matter)
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
          if (ne > (uint32_t) (kEndList - 32 * 8))
            run_ends();
        }
      }
      if (ne != 0)
        run_ends();
#undef SMAX_LOAD_LEFT
    }
    __syncwarp();                          // every survivor of the unit is marked
    if (STATS) { stat0 += stat01[0]; stat1 += stat01[1]; }
    SMAX_PHASE(4)

    // ---------------- K3, first half: the warp turns the unit's bitmaps into entries of the
    // survivor arena, in suffix-array order, and leaves the unit's (records, positions)
    // aggregate for k_emit.  No unit waits for another one here.
    {
      const uint32_t total = (P.debug & 16) ? 0u : ws.count;
      if (total == 0)
      {
        if (lane == 0)
        {
          UnitMeta m;
          m.count = 0; m.pad = 0; m.wsum = 0; m.base = 0;
          P.meta[unit] = m;
        }
      } else
      {
        const uint64_t wtotal = (uint64_t) ws.wsum + ws.open_width;
        const uint4 ew4 = *reinterpret_cast<const uint4 *>(&ws.endbits[4 * lane]);
        const uint32_t ew[4] = {ew4.x, ew4.y, ew4.z, ew4.w};
        const uint32_t cnt = __popc(ew4.x) + __popc(ew4.y) + __popc(ew4.z) + __popc(ew4.w);
        uint32_t inc_c = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
          const uint32_t yc = __shfl_up_sync(0xffffffffu, inc_c, d);
          if (lane >= d) inc_c += yc;
        }
        uint64_t unit_base = ws.chunk_next;
        uint32_t chunk_left = ws.chunk_left;
        __syncwarp();
        if (total > chunk_left)
        {
          // a new chunk (a unit with more entries than a chunk holds gets exactly its own)
          const uint32_t take = total > (uint32_t) kArenaChunk ? total : (uint32_t) kArenaChunk;
          unsigned long long got = 0;
          if (lane == 0)
            got = atomicAdd((unsigned long long *) &P.result[kResArena], (unsigned long long) take);
          unit_base = __shfl_sync(0xffffffffu, got, 0);
          chunk_left = take;
        }
        const bool fits = unit_base + total <= P.arena_capacity;
        if (lane == 0)
        {
          ws.chunk_next = unit_base + total;
          ws.chunk_left = chunk_left - total;
        }
        if (STATS) { if (lane == 0) stat3 += wtotal; }
        if (lane == 0)
        {
          if (!fits)
            P.result[kResOverflow] = 1;         // the host enlarges the arena and scans again
          if (wtotal >> 32)
            P.result[kResError] = 4;            // wider than a shard can be
          UnitMeta m;
          m.count = total; m.pad = 0; m.wsum = wtotal;
          m.base = fits ? unit_base : ~0ull;
          P.meta[unit] = m;
          // the sums of the unit's block (k_emit turns them into offsets)
          const uint64_t *bs = P.blocksum + 2 * (size_t) (unit / kEmitBlock);
          asm volatile("red.global.add.u64 [%0], %1;" :: "l"(bs), "l"((uint64_t) total) : "memory");
          asm volatile("red.global.add.u64 [%0], %1;" :: "l"(bs + 1), "l"(wtotal) : "memory");
        }
        SMAX_PHASE(5)
        if (fits)
        {
          // Large survivors take their length from the .llv record: the r-th large value of the
          // unit is the unit's r-th record, and r comes from a bitmap of the unit's 255 bytes
          // (every lane turns its 128 entries into four words; taken only when needed).
          const bool anybig = ws.anybig != 0;
          __syncwarp();                       // (the lists the bitmap shares its place with are through)
          if (anybig)
          {
            uint32_t c255 = 0;
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
              const uint4 x = *reinterpret_cast<const uint4 *>(ws.lcp + kHalo + lane * 128 + q * 32);
              const uint4 y = *reinterpret_cast<const uint4 *>(ws.lcp + kHalo + lane * 128 + q * 32 + 16);
              const uint32_t mx[4] = {smax_is255(x.x), smax_is255(x.y), smax_is255(x.z), smax_is255(x.w)};
              const uint32_t my[4] = {smax_is255(y.x), smax_is255(y.y), smax_is255(y.z), smax_is255(y.w)};
              const uint32_t wbits = pack_ends16(mx) | pack_ends16(my) << 16;
              ws.bigbits[4 * lane + q] = wbits;
              c255 += __popc(wbits);
            }
            uint32_t inc = c255;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
              const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
              if (lane >= d) inc += y;
            }
            ws.bigrank[lane] = (uint16_t) (inc - c255);
          }
          // the unit's ENDs are listed in suffix-array order (kEndList at a time), then every lane
          // takes one entry at a time: a unit full of repeats does not hang on a few lanes
          uint16_t *ends = ws.endlist;
#pragma unroll 1
          for (uint32_t lbase = 0; lbase < total; lbase += kEndList)
          {
            uint32_t slot = inc_c - cnt;
            if (slot < lbase + kEndList && slot + cnt > lbase)
            {
#pragma unroll
              for (int j = 0; j < 4; j++)
              {
                uint32_t e32 = ew[j];
                while (e32)
                {
                  if (slot >= lbase && slot < lbase + kEndList)
                    ends[slot - lbase] = (uint16_t) (((uint32_t) lane * 4 + j) * 32 + (__ffs(e32) - 1));
                  e32 &= e32 - 1;
                  slot++;
                }
              }
            }
            __syncwarp();
            const uint32_t m = min((uint32_t) kEndList, total - lbase);
#pragma unroll 1
            for (uint32_t i = lane; i < m; i += 32)
            {
              const uint32_t o = ends[i];
              ArenaEntry e;
              e.unit = unit;
              e.end_off = toff32 + o;
              e.width = (uint32_t) survivor_width(ws, o);
              e.pad = 0;
              const uint32_t b = ws.lcp[kHalo + o];
              e.len = b;
              e.len_hi = 0;
              if (b == 255)
              {
                const uint32_t w = o >> 5;
                uint32_t r = ws.bigrank[o >> 7];
                for (uint32_t q = w & ~3u; q < w; q++)
                  r += __popc(ws.bigbits[q]);
                r += __popc(ws.bigbits[w] & ((1u << (o & 31)) - 1u));
                // (the record itself is read by k_emit: nothing is waited for here)
                e.len = kt0 + r;
                e.pad = 1;
              }
              P.arena[unit_base + lbase + i] = e;
            }
            __syncwarp();
          }
        }
        __syncwarp();
        // clean up for the next unit
        {
          uint4 *z = reinterpret_cast<uint4 *>(ws.endbits);
          const uint4 zero = make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int i = lane; i < (int) ((2 * kBitWords * 4) / 16); i += 32)
            z[i] = zero;
          if (lane == 0) { ws.count = 0; ws.wsum = 0; ws.open_width = 0; ws.anybig = 0; }
        }
      }
    }
    SMAX_PHASE(6)
#if SMAX_PROBE
    if (P.debug & 1024)
    {
      // tuning probe: nanoseconds the warp spent on the unit, and when it began
      if (lane == 0)
      {
        P.meta[unit].pad = (uint32_t) (t_last - t_begin);
        P.meta[unit].base = t_begin;
        // phases of the unit in microseconds (8 bits each): large pass, filter, K2 of large values,
        // phase B, K3 head, K3 entries
        unsigned long long ph = 0;
        const int which[6] = {0, 2, 3, 4, 5, 6};
        for (int i = 0; i < 6; i++)
        {
          unsigned long long us = ws.uphase[which[i]] / 1000u;
          ph |= (us > 255 ? 255ull : us) << (8 * i);
          ws.uphase[which[i]] = 0;
        }
        P.meta[unit].wsum = ph;
      }
    }
#endif
    // the next unit
    if (lane == 0)
      cp_async_wait_all();
    __syncwarp();
    unit = next;
    if (have_llv && unit < nunits)
    {
      kt0 = ws.pipe_dir[0];
      k1 = ws.pipe_dir[1];
    }
    next = ws.pipe_next;
    __syncwarp();                           // (lane 0 writes these words again right away)
  }

  if (STATS)
  {
    if (stat0) atomicAdd((unsigned long long *) &P.result[kResStatCand], (unsigned long long) stat0);
    if (stat1) atomicAdd((unsigned long long *) &P.result[kResStatCandWidth], (unsigned long long) stat1);
    if (stat2) atomicAdd((unsigned long long *) &P.result[kResStatLlv], (unsigned long long) stat2);
    if (stat3) atomicAdd((unsigned long long *) &P.result[kResStatSurvWidth], (unsigned long long) stat3);
  }
#if SMAX_PROBE
  if (lane == 0)
    for (int i = 0; i < 8; i++)
      atomicAdd((unsigned long long *) &P.result[kResProbe + i], ws.probe[i]);
#endif
  // k_emit may be put in place (it waits for this grid to complete before it reads anything)
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
// Every arena entry goes to its place: the record, and the occurrence positions
// suf[lb..lb+width) gathered right behind those of the records before it.  One CTA per block
// of kEmitBlock units: the repeats / occurrences before the block are the sums of the blocks
// before it (added up by the scan), those before a unit within the block come from a scan over
// the block's unit aggregates; then one thread per entry of the block (kEmitThreads at a time: the
// block's entries are consecutive in the output, a scan over their widths gives the positions'
// places; the record value of a large survivor is read here).  The last block reports the totals
// and -- multi-GPU -- stores the shard's record count into every shard's count array.
__global__ void __launch_bounds__(kEmitThreads)
k_emit(const __grid_constant__ ScanParams P)
{
  __shared__ unsigned long long uc[kEmitBlock];                     // records before each of the block's units
  __shared__ unsigned long long ubase[kEmitBlock];                  // where their entries sit in the arena
  __shared__ unsigned long long red_c[kEmitThreads / 32], red_w[kEmitThreads / 32];
  __shared__ unsigned long long tot_c[kEmitThreads / 32], tot_w[kEmitThreads / 32];
  asm volatile("griddepcontrol.wait;" ::: "memory");       // the detection grid has completed
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t blk = blockIdx.x;
  const uint64_t nunits = P.nunits;
  const uint64_t a_lo = P.own.a_lo;

  // ---- repeats / occurrences before the block
  unsigned long long sc = 0, sw = 0;
  for (uint32_t j = tid; j < blk; j += kEmitThreads)
  {
    sc += P.blocksum[2 * (size_t) j];
    sw += P.blocksum[2 * (size_t) j + 1];
  }
  // ---- the block's own units
  const uint64_t u = (uint64_t) blk * kEmitBlock + tid;
  UnitMeta m;
  m.count = 0; m.pad = 0; m.wsum = 0; m.base = 0;
  if (tid < kEmitBlock && u < nunits)
    m = P.meta[u];
  unsigned long long ic = m.count, iw = m.wsum;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const unsigned long long yc = __shfl_up_sync(0xffffffffu, ic, d), yw = __shfl_up_sync(0xffffffffu, iw, d);
    if (lane >= d) { ic += yc; iw += yw; }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
  {
    sc += __shfl_xor_sync(0xffffffffu, sc, d);
    sw += __shfl_xor_sync(0xffffffffu, sw, d);
  }
  if (lane == 31) { tot_c[warp] = ic; tot_w[warp] = iw; }
  if (lane == 0) { red_c[warp] = sc; red_w[warp] = sw; }
  __syncthreads();
  unsigned long long ec = ic - m.count, ew = iw - m.wsum, bc = 0, bw = 0, all_c = 0, all_w = 0;
#pragma unroll
  for (int q = 0; q < kEmitThreads / 32; q++)
  {
    if (q < warp) { ec += tot_c[q]; ew += tot_w[q]; }
    all_c += tot_c[q]; all_w += tot_w[q];
    bc += red_c[q]; bw += red_w[q];
  }
  if (tid < kEmitBlock)
  {
    uc[tid] = bc + ec;
    ubase[tid] = m.base;
  }
  if (tid == 0)
  {
    // the other set of block sums is made clean for the next scan
    P.blocksum_next[2 * (size_t) blk] = 0;
    P.blocksum_next[2 * (size_t) blk + 1] = 0;
    if (blk + 1 == gridDim.x)
    {
      P.result[kResCount] = bc + all_c;
      P.result[kResPositions] = bw + all_w;
      // one-sided count exchange: the shard's record count goes straight into every
      // shard's count array (P2P stores over NVLink), tagged with the step
      const uint64_t word = (P.exchange_tag << 40) | ((bc + all_c) & ((1ull << 40) - 1));
      for (int q = 0; q < P.npeers; q++)
        asm volatile("st.release.sys.global.u64 [%0], %1;"
                     :: "l"(P.peer_counts[q] + P.my_rank), "l"(word) : "memory");
      for (int k = 0; k < kResSlots; k++)   // result blocks ping-pong: no memset per scan
        P.result_next[k] = 0;
    }
  }
  __syncthreads();
  if (P.debug & 128)
    return;

  // ---- the entries: one thread each.  The block's entries are consecutive in the output, and so
  // are their occurrence positions: the record index of an entry is its number, the index of its
  // first position the sum of the widths of the entries before it (a scan over the block's
  // entries, kEmitThreads at a time).  The unit of an entry (for its arena slot) is found in the
  // block's prefix of entry counts.
  const uint64_t first_c = uc[0];                          // records before the block
  const uint32_t nent = (uint32_t) all_c;
  uint64_t carry_w = bw;                                   // positions before the block (+ those handled so far)
  for (uint32_t base = 0; base < nent; base += kEmitThreads)
  {
    const uint32_t en = base + tid;
    bool have = en < nent;
    ArenaEntry e;
    e.end_off = 0; e.width = 0; e.pad = 0; e.len = 0; e.len_hi = 0;
    if (have)
    {
      // the last unit whose exclusive count is <= en (units without entries are skipped by the search)
      uint32_t lo = 0, hi = kEmitBlock - 1;
      while (lo < hi)
      {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (uc[mid] - first_c <= en) lo = mid; else hi = mid - 1;
      }
      if (ubase[lo] == ~0ull)
        have = false;                                      // (the arena was too small: the scan is repeated)
      else
        e = P.arena[ubase[lo] + (en - (uint32_t) (uc[lo] - first_c))];
      if (have && e.pad != 0)
      {
        // a large survivor: e.len is the number of its .llv record (the rank of its 255 byte)
        const uint32_t k = e.len;
        uint64_t v = kBadValue;
        if (k < (uint32_t) P.own.nllv)
        {
          const uint2 rc = rec_at(P, k);
          if (rc.x == e.end_off)
            v = rec_value(P, k, rc.y);
        }
        if (v == kBadValue)
          P.result[kResError] = kErrTables;                // a 255 entry without its .llv record
        e.len = (uint32_t) v;
        e.len_hi = (uint32_t) (v >> 32);
      }
    }
    unsigned long long iwd = e.width;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, iwd, d);
      if (lane >= d) iwd += y;
    }
    __syncthreads();                                       // (red_w of the round before has been read)
    if (lane == 31) red_w[warp] = iwd;
    __syncthreads();
    uint64_t po = carry_w + iwd - e.width, round_w = 0;
#pragma unroll
    for (int q = 0; q < kEmitThreads / 32; q++)
    {
      if (q < warp) po += red_w[q];
      round_w += red_w[q];
    }
    carry_w += round_w;
    if (!have)
      continue;
    const uint64_t dst = first_c + en;
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

// compact form of the shard's .llv records (smax_kernels.cuh: kLlvFirst): value + run flags, and
// position - a_lo, in two parallel arrays (a shard holds < 2^32 entries; a value that does not fit
// reads kLlvEscape and is taken from the 16-byte record).  flags[0] tells the scan whether there
// is such a value, flags[1] whether record 0 sits on the very first entry of the arrays.
__global__ void k_llvpack(const smax_llv *llv, uint64_t nllv, uint64_t a_lo, uint32_t *vals,
                          uint32_t *poss, uint32_t *flags)
{
  const uint64_t k = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nllv)
  {
    if (k < nllv + kLlvPad)
    {
      vals[k] = 0;                                  // "no record" behind the last one (16-byte loads)
      poss[k] = kNoRecord;
    }
    return;
  }
  const smax_llv r = llv[k];
  uint32_t v = r.value < (uint64_t) kLlvEscape ? (uint32_t) r.value : kLlvEscape;
  const bool escape = v == kLlvEscape;
  if (escape)
    flags[0] = 1;
  if (k == 0 && r.position == a_lo)
    flags[1] = 1;
  // the neighbour entries: large values when the neighbour records sit right next to this one
  // (compared in full here, so a value that does not fit only matters for the minimum length)
  bool adj_p = false, adj_n = false;
  uint64_t vp = 0, vn = 0;
  if (k > 0)
  {
    const smax_llv q = llv[k - 1];
    adj_p = q.position + 1 == r.position;
    vp = q.value;
  }
  if (k + 1 < nllv)
  {
    const smax_llv q = llv[k + 1];
    adj_n = q.position == r.position + 1;
    vn = q.value;
  }
  const bool falls = !adj_n || vn < r.value;            // the END of a run
  const bool rises = !adj_p || vp < r.value;
  const bool runend = falls && adj_p && vp == r.value;
  const bool edge = k == 0 && r.position == a_lo && a_lo > 0;     // what lies left of it is in a neighbour shard
  if (falls && rises && !escape && !edge)
    v |= kLlvPeak;
  if (runend || (falls && rises && escape) || (falls && edge))
    v |= kLlvGeneral;
  vals[k] = v;
  poss[k] = (uint32_t) (r.position - a_lo);
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



// ------------------------------------------------ .llv records rebuilt on the device
// The positions of the .llv records are redundant with the lcp table: the k-th 255 byte IS
// the k-th record.  The host therefore uploads only the values (4 bytes each,
// 16 -> 4 bytes per record on the PCIe link), and the 16-byte records the rest of the library
// works on are put together here: count the 255 bytes per block of kPosBlock entries, scan
// the counts, then every block writes {a_lo + offset, value} for its 255 bytes in order.
constexpr int kPosBlock = 1 << 16;              // entries per block (8 warps x 16 rounds x 32 chunks)

__device__ __forceinline__ uint32_t mask255_16(const uint4 &x)
{
  const uint32_t m[4] = {smax_is255(x.x), smax_is255(x.y), smax_is255(x.z), smax_is255(x.w)};
  return pack_ends16(m);
}

__global__ void __launch_bounds__(256)
k_llv_count255(const uint8_t *lcp, uint64_t len, uint32_t *counts)
{
  __shared__ uint32_t red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t base = (uint64_t) blockIdx.x * kPosBlock + (uint64_t) warp * (kPosBlock / 8);
  uint32_t cnt = 0;
#pragma unroll 4
  for (int i = 0; i < kPosBlock / 8 / 512; i++)
  {
    const uint64_t o = base + (uint64_t) (i * 32 + lane) * 16;
    if (o < len)                                  // (zero padded: whole chunks may be read)
      cnt += __popc(mask255_16(__ldg(reinterpret_cast<const uint4 *>(lcp + o))));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if (lane == 0) red[warp] = cnt;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    uint32_t t = 0;
    for (int w = 0; w < 8; w++) t += red[w];
    counts[blockIdx.x] = t;
  }
}

// exclusive scan of the block counts in place (one CTA); counts[nblocks] receives the total
__global__ void __launch_bounds__(1024)
k_llv_scan255(uint32_t *counts, uint32_t nblocks)
{
  __shared__ unsigned long long part[1024];
  const uint32_t per = max(2u, (nblocks + 1023u) / 1024u);   // (>= 2: small shards walk the same loop as large ones)
  const uint32_t lo = threadIdx.x * per, hi = min(nblocks, lo + per);
  unsigned long long sum = 0;
  for (uint32_t i = lo; i < hi; i++)
    sum += counts[i];
  part[threadIdx.x] = sum;
  __syncthreads();
  for (int d = 1; d < 1024; d <<= 1)
  {
    const unsigned long long y = threadIdx.x >= (unsigned) d ? part[threadIdx.x - d] : 0ull;
    __syncthreads();
    part[threadIdx.x] += y;
    __syncthreads();
  }
  unsigned long long run = part[threadIdx.x] - sum;
  for (uint32_t i = lo; i < hi; i++)
  {
    const uint32_t c = counts[i];
    counts[i] = (uint32_t) min(run, 0xffffffffull);
    run += c;
  }
  if (threadIdx.x == 1023)
    counts[nblocks] = (uint32_t) min(part[1023], 0xffffffffull);
}

__global__ void __launch_bounds__(256)
k_llv_fill(const uint8_t *lcp, uint64_t len, uint64_t a_lo, const uint32_t *offs, const uint32_t *vals32,
           uint64_t nllv, smax_llv *llv)
{
  __shared__ uint32_t wsum[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint64_t base = (uint64_t) blockIdx.x * kPosBlock + (uint64_t) warp * (kPosBlock / 8);
  // the warp's 255 bytes, for its place within the block
  uint32_t cnt = 0;
#pragma unroll 4
  for (int i = 0; i < kPosBlock / 8 / 512; i++)
  {
    const uint64_t o = base + (uint64_t) (i * 32 + lane) * 16;
    if (o < len)
      cnt += __popc(mask255_16(__ldg(reinterpret_cast<const uint4 *>(lcp + o))));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if (lane == 0) wsum[warp] = cnt;
  __syncthreads();
  uint64_t k = offs[blockIdx.x];
  for (int w = 0; w < warp; w++) k += wsum[w];
  // in order: round by round, chunk by chunk (lane), byte by byte
#pragma unroll 1
  for (int i = 0; i < kPosBlock / 8 / 512; i++)
  {
    const uint64_t o = base + (uint64_t) (i * 32 + lane) * 16;
    uint32_t m = 0;
    if (o < len)
      m = mask255_16(__ldg(reinterpret_cast<const uint4 *>(lcp + o)));
    uint32_t c = __popc(m), inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const uint32_t y = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += y;
    }
    uint64_t kk = k + inc - c;
    while (m)
    {
      const uint32_t j = __ffs(m) - 1;
      m &= m - 1;
      if (kk < nllv)                              // (more 255 bytes than values: the host notices the total)
      {
        smax_llv r;
        r.position = a_lo + o + j;
        r.value = vals32[kk];
        llv[kk] = r;
      }
      kk++;
    }
    k += __shfl_sync(0xffffffffu, inc, 31);
  }
}

// ------------------------------------------------------------ upload-time unit order
// The scan's warps take whole units, a handful each: what is in flight when the last unit is
// handed out decides how long the end of the scan drags on.  So the units are taken heaviest
// first.  The weight of a unit is a property of the tables alone: the number of its entries
// that reach a threshold (the smallest value that at most one entry in eight reaches, from the
// histogram of the shard's lcp bytes) plus twice the number of its .llv records.  A counting
// sort over kOrderBuckets weight classes puts the units in order (within a class as they come).
__device__ __forceinline__ uint32_t order_bucket(uint32_t weight)
{
  const uint32_t b = weight / 12u;                       // weight <= 3 * kUnitBytes
  return (uint32_t) (kOrderBuckets - 1) - min(b, (uint32_t) (kOrderBuckets - 1));
}

__global__ void __launch_bounds__(256)
k_unitorder_weight(const uint8_t *lcp_own, uint64_t own_len, const uint32_t *unitdir,
                   const unsigned long long *hist, uint32_t *weight, uint32_t *bins, uint64_t nunits)
{
  __shared__ uint32_t s_thr;
  if (threadIdx.x < 32)
  {
    // threshold: the smallest v >= 2 such that at most 1/8 of the entries are >= v
    const int lane = threadIdx.x;
    unsigned long long part = 0;
    for (int v = lane * 8; v < lane * 8 + 8; v++)
      part += hist[v];
    unsigned long long incl = part;                      // suffix sums over the lanes
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const unsigned long long y = __shfl_down_sync(0xffffffffu, incl, d);
      if (lane + d < 32) incl += y;
    }
    const unsigned long long total = __shfl_sync(0xffffffffu, incl, 0);
    // entries >= v for the v of this lane's group, from the top
    unsigned long long ge = incl - part;                 // entries >= (lane + 1) * 8
    uint32_t best = 256;
    for (int v = lane * 8 + 7; v >= lane * 8; v--)
    {
      ge += hist[v];
      if (ge * 8 <= total && v >= 2)
        best = (uint32_t) v;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
      best = min(best, __shfl_xor_sync(0xffffffffu, best, d));
    if (lane == 0)
      s_thr = min(best, 255u);
  }
  __syncthreads();
  uint32_t kadd; int himode;
  smax_ge_consts(s_thr, &kadd, &himode);
  const int lane = threadIdx.x & 31;
  const uint64_t u = (uint64_t) blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (u >= nunits)
    return;
  const uint64_t off = u * kUnitBytes;
  uint32_t cnt = 0;
#pragma unroll
  for (int j = 0; j < kUnitBytes / (32 * 16); j++)
  {
    const uint64_t o = off + (uint64_t) (j * 32 + lane) * 16;
    if (o < own_len)                                      // (the arrays are zero padded: whole chunks may be read)
    {
      const uint4 x = __ldg(reinterpret_cast<const uint4 *>(lcp_own + o));
      cnt += __popc(smax_ge(x.x, kadd, himode)) + __popc(smax_ge(x.y, kadd, himode)) +
             __popc(smax_ge(x.z, kadd, himode)) + __popc(smax_ge(x.w, kadd, himode));
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  if (lane == 0)
  {
    const uint32_t w = cnt + 2u * min(unitdir[u + 1] - unitdir[u], (uint32_t) kUnitBytes);
    weight[u] = w;
    atomicAdd(&bins[order_bucket(w)], 1u);
  }
}

__global__ void __launch_bounds__(kOrderBuckets)
k_unitorder_offsets(uint32_t *bins)
{
  __shared__ uint32_t s[kOrderBuckets];
  const int t = threadIdx.x;
  s[t] = bins[t];
  __syncthreads();
  for (int d = 1; d < kOrderBuckets; d <<= 1)
  {
    const uint32_t y = t >= d ? s[t - d] : 0u;
    __syncthreads();
    s[t] += y;
    __syncthreads();
  }
  bins[t] = s[t] - bins[t];                              // exclusive
}

__global__ void k_unitorder_fill(const uint32_t *weight, uint32_t *bins, uint32_t *order, uint64_t nunits)
{
  const uint64_t u = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= nunits)
    return;
  order[atomicAdd(&bins[order_bucket(weight[u])], 1u)] = (uint32_t) u;
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

cudaError_t launch_llvpack(const smax_llv *llv, uint64_t nllv, uint64_t a_lo, uint32_t *vals,
                           uint32_t *poss, uint32_t *flags, cudaStream_t st)
{
  const int threads = 256;
  const uint64_t blocks = (nllv + kLlvPad + threads - 1) / threads;
  k_llvpack<<<(unsigned) blocks, threads, 0, st>>>(llv, nllv, a_lo, vals, poss, flags);
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


// order: nunits words + kOrderScratch scratch words; the weights live in the words the order
// ends up in only while they are needed (a second array of nunits words behind the scratch)
cudaError_t launch_unitorder(const uint8_t *lcp_own, uint64_t own_len, const uint32_t *unitdir,
                             const unsigned long long *hist, uint32_t *order, uint64_t nunits,
                             cudaStream_t st)
{
  if (nunits == 0)
    return cudaSuccess;
  uint32_t *bins = order + nunits, *weight = order + nunits + kOrderScratch;
  cudaError_t e = cudaMemsetAsync(bins, 0, kOrderScratch * sizeof(uint32_t), st);
  if (e != cudaSuccess)
    return e;
  k_unitorder_weight<<<(unsigned) ((nunits + 7) / 8), 256, 0, st>>>(lcp_own, own_len, unitdir, hist, weight,
                                                                  bins, nunits);
  k_unitorder_offsets<<<1, kOrderBuckets, 0, st>>>(bins);
  k_unitorder_fill<<<(unsigned) ((nunits + 255) / 256), 256, 0, st>>>(weight, bins, order, nunits);
  return cudaGetLastError();
}


// rebuilds the 16-byte .llv records of a shard from its lcp bytes and the uploaded values;
// scratch: len / kPosBlock + 2 words, scratch[nblocks] receives the number of 255 bytes found
cudaError_t launch_llv_rebuild(const uint8_t *lcp, uint64_t len, uint64_t a_lo, const uint32_t *vals32,
                               uint64_t nllv, smax_llv *llv, uint32_t *scratch, cudaStream_t st)
{
  const uint32_t nblocks = (uint32_t) ((len + kPosBlock - 1) / kPosBlock);
  if (nblocks == 0)
    return cudaMemsetAsync(scratch, 0, sizeof(uint32_t), st);
  k_llv_count255<<<nblocks, 256, 0, st>>>(lcp, len, scratch);
  k_llv_scan255<<<1, 1024, 0, st>>>(scratch, nblocks);
  k_llv_fill<<<nblocks, 256, 0, st>>>(lcp, len, a_lo, scratch, vals32, nllv, llv);
  return cudaGetLastError();
}

uint64_t llv_rebuild_scratch_words(uint64_t len)
{
  return (len + kPosBlock - 1) / kPosBlock + 2;
}

static const void *scan_kernel(bool stats)
{
  return stats ? (const void *) k_scan<true> : (const void *) k_scan<false>;
}

// One scan = two launches on the stream: detection (units are taken from a ticket; no unit
// depends on another one, so any grid size makes progress) and the emit of the arena entries.
cudaError_t launch_scan(const ScanParams &p, bool stats, int grid, int sm_count, cudaStream_t st)
{
  void *args[] = {(void *) &p};
  cudaError_t e = cudaLaunchKernel(scan_kernel(stats), dim3(grid), dim3(kThreads), args,
                                            kWarps * sizeof(WarpSmem), st);
  if (e != cudaSuccess)
    return e;
  // the second kernel is launched programmatically dependent: it is put in place while the
  // grid before it drains and waits (griddepcontrol.wait) for its completion
  (void) sm_count;
  const uint64_t nunits = p.nunits;
  const unsigned blocks = (unsigned) ((nunits + kEmitBlock - 1) / kEmitBlock);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.stream = st;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(blocks > 0 ? blocks : 1);
  cfg.blockDim = dim3(kEmitThreads);
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
