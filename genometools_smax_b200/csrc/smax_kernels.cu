/*
  smax_kernels.cu -- hand-written sm_100a kernels of the supermaximal-repeat
  scan.  ONE fused pass over the lcptab replaces the reference's stack sweep
  (/root/reference/src/match/esa-bottomup.c:116-273) and its per-node
  left-character bookkeeping (/root/reference/src/match/esa-maxpairs.c:181-360).

  k_scan is a persistent kernel (2 CTAs of 512 threads per SM, cooperative
  launch).  CTA b works on the 16 KiB lcptab tiles b, b + grid, b + 2 grid, ...

    feed  One elected thread keeps a ring of TMA bulk copies
          (cp.async.bulk.shared::cluster.global + mbarrier complete_tx) in
          flight: three lcp tiles, and -- in regions of the index where plateau
          ends are frequent -- two bwt tiles, each with a 16-byte halo either
          side.  No register staging, no per-thread loads of table bytes.
    K1    plateau detection, flat and bit-parallel (smax_swar.h).  Each thread
          classifies two 16-byte chunks out of shared memory with SWAR byte
          arithmetic: ends of runs with a value >= minlength that fall to a
          smaller value, entered from a smaller value 1, 2 or 3 entries back
          (SA width 2, 3, 4 -- 99.9 % of all plateaus).  Only runs of >= 4 equal
          values are walked (shared memory first, then global memory / the left
          neighbour shard).  Large values (byte 255) are resolved in place in
          .llv RECORD space: each tile streams the .llv records that fall into
          it (found through a per-4096-entry directory, no global rank) with
          coalesced 16-byte loads; a record's neighbours come from the adjacent
          lanes.
    K2    left-distinctness, bit-parallel on the staged bwt words for widths
          <= 4 (pairwise byte compares; specials (>= 254) never collide under
          the GenomeTools convention, esa-maxpairs.c:24-31), a 256-bit
          alphabet mask otherwise.  Tiles of a sparse region read the few
          bwt bytes they need straight from global memory instead.
    K3    order-preserving compaction + emit.  A survivor sets the bit of its
          end offset in a per-tile bitmap (rank = popcount prefix) and is staged
          in shared memory; tile totals (record count, position count) are
          exchanged generation-wise through epoch-tagged 16-byte status pairs
          (no memset between scans, no chain of dependent look-backs); records
          are then written in suffix-array order and the occurrence positions
          suf[lb..lb+width) are gathered right behind them.  The write of a
          tile is deferred by one tile so that nobody waits for a straggler; a
          tile with more survivors than the stage holds is written at once in
          rank windows, re-running K1/K2 on the still resident stage.

  k_llvdir builds the .llv bucket directory at upload time.
*/
#include "smax_kernels.cuh"
#include "smax_swar.h"

// tuning switches (tools/build_variants.py): the defaults are the measured best
#ifdef SMAX_OUTLINE_PASS
#define SMAX_PASS_INLINE __noinline__
#else
#define SMAX_PASS_INLINE __forceinline__
#endif

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
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.b32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
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
  if (tv == nullptr) { P.result[kResError] = 1; return kBadValue; }
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
  if (tv == nullptr) { P.result[kResError] = 1; return 255; }
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

struct ScanSmem
{
  // ring slots: slot[kHalo + i] = table[tile_lo + i], i in [-kHalo, kTileBytes + kHalo)
  alignas(128) uint8_t lcp[kStages][kStageBytes];
  alignas(128) uint8_t bwt[kStages][kStageBytes];
  alignas(16) smax_llv llv[kLlvSlot + 2];    // one slot: filled while the next tile's small values are scanned
  uint32_t log_v[kLogCap], log_w[kLogCap], log_t[kLogCap];
  uint16_t wlist[kThreads / 32][kWarpList];   // per warp: chunks / records that passed the filter
  // per generation of the batch being resolved: totals, then prefix of this CTA's tile
  unsigned long long gtot_c[kMaxGen], gtot_w[kMaxGen], gexc_c[kMaxGen], gexc_w[kMaxGen];
  LlvMeta meta;                  // of the tile in work (then of the next one)
  unsigned long long tile_w;     // position count of the tile in work
  unsigned long long run_c, run_w;   // records / positions of all resolved generations
  unsigned long long last_c, last_w; // prefix of this CTA's tile in the generation resolved last
  uint32_t log_n;                // survivors appended (> kLogCap: the tile did not fit)
  alignas(8) uint64_t lfull[kStages];   // mbarriers: the bytes of the slot have landed
  alignas(8) uint64_t bfull[kStages];
  alignas(8) uint64_t vfull;
};

// K3, first half: a survivor joins the log of its CTA.  [plo, phi) restricts a
// replay to a piece of the tile.
__device__ __noinline__ void emit_survivor(const ScanParams &P, ScanSmem &sm, uint32_t it16,
                                              uint32_t o, uint64_t v, uint64_t width, uint32_t plo,
                                              uint32_t phi)
{
  if (o < plo || o >= phi)
    return;
  const uint32_t slot = atomicAdd(&sm.log_n, 1u);
  atomicAdd(&sm.tile_w, (unsigned long long) width);
  if (width >> 32)
    P.result[kResError] = 4;              // wider than a shard can be
  if (slot < (uint32_t) kLogCap)
  {
    sm.log_v[slot] = v < (uint64_t) kLongValue ? (uint32_t) v : kLongValue;
    sm.log_w[slot] = (uint32_t) width;
    sm.log_t[slot] = o | (it16 << 16);
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
// rec(k) reads record k (from the staged slot where it holds it).
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

// per-pass context of one tile
struct PassCtx
{
  const uint8_t *sl;       // resident lcp slot
  const uint8_t *sb;       // resident bwt slot, or nullptr (sparse region)
  const smax_llv *sv;      // resident .llv slot
  uint64_t tile_lo;        // global lcp index of tile offset 0
  uint32_t it16;           // tag of the tile in the survivor log
  uint32_t plo, phi;       // tile offsets whose ends are wanted (a replay takes pieces)
  uint32_t vparity;        // parity of the .llv slot barrier to wait for
};

// K2 + emit for one local-maximum plateau [e + 1 - width, e] of value v.
template <bool STATS>
__device__ SMAX_PASS_INLINE void test_and_emit(const ScanParams &P, ScanSmem &sm, const PassCtx &C,
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
      emit_survivor(P, sm, C.it16, o, v, width, C.plo, C.phi);
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

// One detection pass over the resident tile (K1 + K2 + logging of survivors).
// Both halves work in two phases per warp so that the expensive part runs on
// full warps: phase A is a cheap filter over everything (does the chunk hold a
// byte >= the threshold / does the .llv record end a run), whose hits are
// compacted into a small per-warp list with a ballot; phase B takes the list
// entries lane by lane.
// Returns, to every thread, the number of threads that met a candidate plateau
// (the density signal that decides whether the next tiles prefetch their bwt).
template <bool STATS>
__device__ SMAX_PASS_INLINE int tile_pass(const ScanParams &P, ScanSmem &sm, const PassCtx &C)
{
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  uint16_t *list = sm.wlist[tid >> 5];
  const uint64_t tile_lo = C.tile_lo;
  uint32_t kadd; int himode;
  smax_ge_consts(P.mb, &kadd, &himode);
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  // ends at or beyond g_hi belong to the next shard
  const uint32_t valid = min((uint32_t) min((uint64_t) kTileBytes, P.g_hi - tile_lo), C.phi);
  uint64_t stat[4] = {0, 0, 0, 0};
  int met = 0;

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
            emit_survivor(P, sm, C.it16, o, lp[bit], width, C.plo, C.phi);
        }
      }
    };
#ifndef SMAX_NO_COMPACT_SMALL
    // phase A: which of the warp's 128 chunks hold a byte >= the threshold?
    uint32_t n = 0;
#pragma unroll
    for (int c = 0; c < kItems; c++)
    {
      const uint32_t ch = (uint32_t) (c * kThreads + tid), o0 = ch * kChunk;
      bool hit = false;
      if (o0 < valid && o0 + kChunk > C.plo)
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
#else
#pragma unroll 1
    for (int c = 0; c < kItems; c++)
    {
      const uint32_t o0 = (uint32_t) (c * kThreads + tid) * kChunk;
      if (o0 < valid && o0 + kChunk > C.plo)
      {
        const uint4 x = *reinterpret_cast<const uint4 *>(C.sl + kHalo + o0);
        if ((smax_ge(x.x, kadd, himode) | smax_ge(x.y, kadd, himode) | smax_ge(x.z, kadd, himode) |
             smax_ge(x.w, kadd, himode)) != 0)
          process_chunk(o0);
      }
    }
#endif
  }

  // ---- large values: the tile's .llv records out of the staged slot (the few
  // beyond its capacity come from global memory).  A record ends a plateau iff
  // its right neighbour is no consecutive record with a value >= its own.
  const LlvMeta M = sm.meta;
  if (M.k0 < M.k1)
  {
    mbar_wait(&sm.vfull, C.vparity);       // issued when the previous tile was done
    const smax_llv *llv = P.own.llv;
    const uint64_t nllv = P.own.nllv;
    const uint64_t a_lo = P.own.a_lo;
    const uint64_t lo = max(tile_lo + C.plo, P.g_lo);
    const uint64_t hi = min(tile_lo + (uint64_t) min((uint32_t) kTileBytes, C.phi), P.g_hi);
    auto rec = [&](uint32_t k) -> smax_llv
    {
      const uint32_t i = k - M.kfirst;
      return i < M.nrec ? C.sv[i] : ld_llv(&llv[k]);
    };
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
      if (STATS && C.plo == 0) stat[2]++;
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
#ifdef SMAX_COMPACT_LLV
    uint32_t n = 0;
    auto drain = [&]()
    {
      __syncwarp();
#pragma unroll 1
      for (uint32_t i = lane; i < n; i += 32)
        process_end(M.k0 + list[i]);
      __syncwarp();
      n = 0;
    };
    // phase A over the records, kThreads at a time
#pragma unroll 1
    for (uint32_t kb = M.k0; kb < M.k1; kb += kThreads)
    {
      const uint32_t k = kb + tid;
      const bool hit = k < M.k1 && is_end(k);
      const uint32_t votes = __ballot_sync(0xffffffffu, hit);
      if (hit)
        list[n + __popc(votes & lt_mask)] = (uint16_t) (k - M.k0);
      n += __popc(votes);
      if (n > (uint32_t) kWarpList - 32)
        drain();
    }
    drain();
#else
#pragma unroll 1
    for (uint32_t k = M.k0 + tid; k < M.k1; k += kThreads)
      if (is_end(k))
        process_end(k);
#endif
  }
  if (STATS && C.plo == 0 && C.phi >= (uint32_t) kTileBytes)
  {
    if (stat[0]) atomicAdd((unsigned long long *) &P.result[kResStatCand], (unsigned long long) stat[0]);
    if (stat[1]) atomicAdd((unsigned long long *) &P.result[kResStatCandWidth], (unsigned long long) stat[1]);
    if (stat[2]) atomicAdd((unsigned long long *) &P.result[kResStatLlv], (unsigned long long) stat[2]);
    if (stat[3]) atomicAdd((unsigned long long *) &P.result[kResStatSurvWidth], (unsigned long long) stat[3]);
  }
  return __syncthreads_count(met);
}

// the replay of a tile that did not fit into the log: same pass, kept out of line
// so that the hot loop of the kernel stays small
__device__ __noinline__ int tile_pass_replay(const ScanParams &P, ScanSmem &sm, const PassCtx &C)
{
  return tile_pass<false>(P, sm, C);
}

__device__ __forceinline__ uint64_t suf_at(const ScanParams &P, uint64_t i)
{
  const TableView *tv = view_for(P, i);
  if (tv == nullptr || tv->suf == nullptr) { P.result[kResError] = 1; return 0; }
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
// order; sm.gexc_c/w[t - t0] is the global prefix of the tile tagged t.  The
// entries of one tile are adjacent in the log, so rank and position offset of
// an entry within its tile come from a look at its neighbours.
__device__ __forceinline__ void write_log(const ScanParams &P, ScanSmem &sm, uint32_t n, uint32_t t0,
                                          uint32_t nt, uint32_t it_of_t0, uint32_t me, uint32_t grid)
{
  const uint64_t base_off = P.g_lo - P.own.a_lo;
  for (uint32_t e = threadIdx.x; e < n; e += kThreads)
  {
    const uint32_t tag = sm.log_t[e], off = tag & 0xffffu;
    const uint32_t t = (tag >> 16) - t0;
    if (t >= nt)
      continue;
    uint32_t rank = 0;
    uint64_t posoff = 0;
    for (int i = (int) e - 1; i >= 0 && (sm.log_t[i] >> 16) == (tag >> 16); i--)
      if (sm.log_t[i] < tag) { rank++; posoff += sm.log_w[i]; }
    for (uint32_t i = e + 1; i < n && (sm.log_t[i] >> 16) == (tag >> 16); i++)
      if (sm.log_t[i] < tag) { rank++; posoff += sm.log_w[i]; }
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
      if (po + wd <= P.pos_capacity)
        for (uint64_t k = 0; k < wd; k++)
          P.positions[po + k] = suf_at(P, lb + k);
      else
        P.result[kResOverflow] = 1;
    }
  }
}

// Resolve the generations [base_it, upto) of this CTA and write the log.
// Leaves sm.last_c/w = prefix of this CTA's tile in generation upto - 1.
__device__ __noinline__ void flush_log(const ScanParams &P, ScanSmem &sm, uint32_t base_it,
                                       uint32_t upto, uint32_t me, uint32_t grid)
{
  const int tid = threadIdx.x;
  __syncthreads();                               // the log is complete
  const uint32_t n = min(sm.log_n, (uint32_t) kLogCap);
  for (uint32_t g0 = base_it; g0 < upto; g0 += kMaxGen)
  {
    const uint32_t gn = min(upto - g0, (uint32_t) kMaxGen);
    // warp w sums the aggregates of generations g0 + w, g0 + w + 8, ...: the lanes
    // read different tiles, four loads in flight each, no atomics
    for (uint32_t g = tid >> 5; g < gn; g += kThreads / 32)
    {
      const uint64_t first = (uint64_t) (g0 + g) * grid;
      const uint32_t ng = (uint32_t) min((uint64_t) grid, (uint64_t) P.ntiles - first);
      uint64_t ea = 0, eb = 0, ta = 0, tb = 0;
      if (!(P.debug & 1))
        for (uint32_t j0 = tid & 31; j0 < ng; j0 += 4 * 32)
        {
          uint64_t wa[4], wb[4];
#pragma unroll
          for (int r = 0; r < 4; r++)           // all loads first, then the checks
          {
            const uint32_t j = j0 + r * 32;
            wa[r] = wb[r] = 0;
            if (j < ng)
              ld_pair(&P.status[2 * (first + j)], wa[r], wb[r]);
          }
#pragma unroll
          for (int r = 0; r < 4; r++)
          {
            const uint32_t j = j0 + r * 32;
            if (j < ng)
            {
              unsigned backoff = 32;
              while ((uint32_t) (wa[r] >> (kValueBits + 2)) != P.epoch ||
                     (uint32_t) (wb[r] >> (kValueBits + 2)) != P.epoch)
              {
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
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        ea += __shfl_xor_sync(0xffffffffu, ea, o);
        eb += __shfl_xor_sync(0xffffffffu, eb, o);
        ta += __shfl_xor_sync(0xffffffffu, ta, o);
        tb += __shfl_xor_sync(0xffffffffu, tb, o);
      }
      if ((tid & 31) == 0)
      {
        sm.gtot_c[g] = ta; sm.gtot_w[g] = tb; sm.gexc_c[g] = ea; sm.gexc_w[g] = eb;
      }
    }
    __syncthreads();
    if (tid == 0)
    {
      unsigned long long rc = sm.run_c, rw = sm.run_w;
      for (uint32_t g = 0; g < gn; g++)
      {
        const unsigned long long ec = sm.gexc_c[g], ew = sm.gexc_w[g];
        sm.gexc_c[g] = rc + ec; sm.gexc_w[g] = rw + ew;
        rc += sm.gtot_c[g]; rw += sm.gtot_w[g];
      }
      sm.run_c = rc; sm.run_w = rw;
      sm.last_c = sm.gexc_c[gn - 1]; sm.last_w = sm.gexc_w[gn - 1];
    }
    __syncthreads();
    write_log(P, sm, n, g0 - base_it, gn, g0, me, grid);
    __syncthreads();
  }
  if (tid == 0)
    sm.log_n = 0;
  __syncthreads();
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

template <bool STATS>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
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
    for (int s = 0; s < kStages; s++)
    {
      mbar_init(&sm.lfull[s], 1); mbar_init(&sm.bfull[s], 1);
    }
    mbar_init(&sm.vfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sm.log_n = 0; sm.tile_w = 0; sm.run_c = 0; sm.run_w = 0; sm.last_c = 0; sm.last_w = 0;
  }
  __syncthreads();

  // a region is "dense" when its tiles keep meeting candidate plateaus: then the
  // bwt tiles are prefetched with the lcp tiles.  First guess from the .llv share.
  bool dense_mode = P.own.nllv * 64 > (P.g_hi - P.g_lo) && !(P.debug & 8);
  uint32_t bissued = 0, bphase = 0, vphase = 0;   // per slot: bwt copy under way / parities to wait for

  // .llv directory entries of a tile (thread 0 reads them one tile early)
  auto dir_of = [&](uint32_t t, uint32_t &d0, uint32_t &d1)
  {
    d0 = d1 = 0;
    if (P.own.nllv != 0 && !(P.debug & 4))
    {
      const uint64_t toff = base_off + (uint64_t) t * kTileBytes;
      d0 = P.own.llvdir[toff >> kLlvBucketShift];
      d1 = P.own.llvdir[((toff + kTileBytes - 1) >> kLlvBucketShift) + 1];
    }
  };
  // thread 0: start the copies of tile t into ring slot `slot`
  auto issue_tile = [&](uint32_t t, int slot, bool with_bwt)
  {
    const Feed f = feed_of(base_off + (uint64_t) t * kTileBytes, readable);
    mbar_expect_tx(&sm.lfull[slot], f.bytes);
    tma_load(sm.lcp[slot] + f.dst, P.own.lcp + f.src, f.bytes, &sm.lfull[slot]);
    if (with_bwt)
    {
      mbar_expect_tx(&sm.bfull[slot], f.bytes);
      tma_load(sm.bwt[slot] + f.dst, P.own.bwt + f.src, f.bytes, &sm.bfull[slot]);
    }
  };
  // thread 0: start the copy of the .llv records [d0, d1) of the next tile (plus one
  // neighbour either side) into the .llv slot
  auto issue_llv = [&](uint32_t d0, uint32_t d1)
  {
    LlvMeta m;
    m.k0 = d0; m.k1 = d1; m.kfirst = d0 > 0 ? d0 - 1 : 0; m.nrec = 0;
    if (d0 < d1)
    {
      m.nrec = (uint32_t) min((uint64_t) min((uint64_t) d1 + 1, P.own.nllv) - m.kfirst,
                              (uint64_t) (kLlvSlot + 2));
      mbar_expect_tx(&sm.vfull, m.nrec * (uint32_t) sizeof(smax_llv));
      tma_load(sm.llv, P.own.llv + m.kfirst, m.nrec * (uint32_t) sizeof(smax_llv), &sm.vfull);
    }
    sm.meta = m;
  };

  for (uint32_t j = 0; j < (uint32_t) kStages; j++)
    if ((uint64_t) me + (uint64_t) j * grid < P.ntiles)
    {
      if (tid == 0)
      {
        issue_tile(me + j * grid, (int) j, dense_mode);
        if (j == 0)
        {
          uint32_t d0, d1;
          dir_of(me, d0, d1);
          issue_llv(d0, d1);
        }
      }
      if (dense_mode)
        bissued |= 1u << j;
    }
  __syncthreads();                        // meta of the first tile

  uint32_t base_it = 0;                   // first generation this CTA has not resolved yet
  uint32_t it = 0;
  for (uint32_t tile = me; tile < P.ntiles; tile += grid, it++)
  {
    const int slot = it & 1;
    const uint64_t toff = base_off + (uint64_t) tile * kTileBytes;
    const bool more = (uint64_t) tile + (uint64_t) kStages * grid < P.ntiles;
    // write the log out before a tile could overflow it; the generations before
    // this one were published a tile ago, so nobody is waited for
    uint32_t log_before = sm.log_n;
    if (log_before > (uint32_t) kLogCap * 3 / 4 || it - base_it >= 60000u)
    {
      flush_log(P, sm, base_it, it, me, grid);
      base_it = it;
      log_before = 0;
    }
    const bool next = (uint64_t) tile + grid < P.ntiles;
    uint32_t d0 = 0, d1 = 0;
    if (tid == 0 && next)
      dir_of(tile + grid, d0, d1);               // consumed after the pass
    PassCtx C;
    C.tile_lo = P.own.a_lo + toff;
    C.it16 = it - base_it;
    C.plo = 0; C.phi = kTileBytes;
    C.sl = sm.lcp[slot];
    C.sb = nullptr;
    C.sv = sm.llv;
    C.vparity = vphase;
    mbar_wait(&sm.lfull[slot], (it >> 1) & 1);
    if ((bissued >> slot) & 1u)
    {
      mbar_wait(&sm.bfull[slot], (bphase >> slot) & 1u);
      bphase ^= 1u << slot;
      bissued &= ~(1u << slot);
      C.sb = sm.bwt[slot];
    }
    const uint32_t nllv_tile = sm.meta.k1 - sm.meta.k0;
    if (nllv_tile != 0)
      vphase ^= 1u;                        // the pass waits for this phase of the .llv slot
    // edges of the table: the left halo of the first tile comes from the left
    // neighbour shard (or repeats the first entry, which sends every plateau
    // that touches the edge into the walk that reports the missing range);
    // bytes past the zero pad are zero
    {
      const Feed f = feed_of(toff, readable);
      if (f.dst != 0 || f.dst + f.bytes != (uint32_t) kStageBytes)
      {
        if (f.dst != 0 && tid < kHalo)
        {
          const uint64_t a_lo = P.own.a_lo;
          uint32_t lv = 0, bv = 0;
          if (a_lo >= (uint64_t) kHalo)
          {
            const uint64_t q = a_lo - kHalo + tid;
            const TableView *tv = view_for(P, q);
            if (tv != nullptr) { lv = tv->lcp[q - tv->a_lo]; bv = tv->bwt[q - tv->a_lo]; }
            else { lv = P.own.lcp[0]; bv = 0; }
          }
          sm.lcp[slot][tid] = (uint8_t) lv;
          if (C.sb != nullptr) sm.bwt[slot][tid] = (uint8_t) bv;
        }
        for (uint32_t i = f.dst + f.bytes + tid; i < (uint32_t) kStageBytes; i += kThreads)
        {
          sm.lcp[slot][i] = 0;
          if (C.sb != nullptr) sm.bwt[slot][i] = 0;
        }
        __syncthreads();
      }
    }
    const int met = tile_pass<STATS>(P, sm, C);
    dense_mode = (met >= 8 || nllv_tile >= 64) && !(P.debug & 8);
    const uint32_t log_after = sm.log_n;
    if (tid == 0)
    {
      publish_aggregate(P.status, tile, log_after - log_before, sm.tile_w, P.epoch);
      sm.tile_w = 0;
    }
    if (log_after > (uint32_t) kLogCap)
    {
      // The tile did not fit into the log: drop its partial entries, write the
      // earlier tiles and resolve this generation at once, then redo the tile in
      // pieces of its offset range (a piece of 1024 entries holds at most 512
      // plateau ends, so halving always ends).
      const uint32_t count = log_after - log_before;
      __syncthreads();
      if (tid == 0)
        sm.log_n = log_before;
      flush_log(P, sm, base_it, it + 1, me, grid);
      base_it = it + 1;
      uint64_t bc = sm.last_c, bw = sm.last_w;
      uint32_t piece = kTileBytes;
      while (piece > 1024 && (uint64_t) count * piece > (uint64_t) (kLogCap / 2) * kTileBytes)
        piece >>= 1;
      C.it16 = 0;
      for (uint32_t lo = 0; lo < (uint32_t) kTileBytes;)
      {
        C.plo = lo; C.phi = lo + piece;
        tile_pass_replay(P, sm, C);
        const uint32_t n = sm.log_n;
        const uint64_t w = sm.tile_w;
        __syncthreads();
        if (n > (uint32_t) kLogCap)
        {
          piece >>= 1;                           // retry the same range in halves
          if (tid == 0) { sm.log_n = 0; sm.tile_w = 0; }
          __syncthreads();
          continue;
        }
        if (tid == 0) { sm.gexc_c[0] = bc; sm.gexc_w[0] = bw; }
        __syncthreads();
        write_log(P, sm, n, 0, 1, it, me, grid);
        __syncthreads();
        if (tid == 0) { sm.log_n = 0; sm.tile_w = 0; }
        __syncthreads();
        bc += n; bw += w;
        lo += piece;
      }
    }
    // the slots of this tile are free: start the copies of the tile after next
    if (tid == 0 && next)
      issue_llv(d0, d1);
    if (more)
    {
      if (tid == 0)
        issue_tile(tile + kStages * grid, slot, dense_mode);
      if (dense_mode)
        bissued |= 1u << slot;
    }
    __syncthreads();                 // slot contents, meta[] and log counters settle
  }
  // ---- the survivors still in the log; the owner of the last tile also resolves
  // every generation to report the totals
  const bool owns_last = P.ntiles != 0 && (P.ntiles - 1) % grid == me;
  if (it > base_it && (sm.log_n != 0 || owns_last))
    flush_log(P, sm, base_it, it, me, grid);
  if (owns_last && tid == 0)
  {
    P.result[kResCount] = sm.run_c;
    P.result[kResPositions] = sm.run_w;
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
      }
      for (int k = 0; k < kResSlots; k++)   // result blocks ping-pong: no memset per scan
        P.result_next[k] = 0;
    }
  }
}

// ------------------------------------------------------- .llv directory
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

// --------------------------------------------------------------- launchers
cudaError_t launch_llvdir(const smax_llv *llv, uint64_t nllv, uint64_t a_lo,
                          uint32_t *dir, uint64_t nentries, cudaStream_t st)
{
  const int threads = 256;
  const uint64_t blocks = (nentries + threads - 1) / threads;
  k_llvdir<<<(unsigned) blocks, threads, 0, st>>>(llv, nllv, a_lo, dir, nentries);
  return cudaGetLastError();
}

// Cooperative launch: the ordered prefix exchange needs every CTA of the grid
// resident at the same time (grid <= SMs x resident CTAs per SM, computed by the
// caller); the runtime then guarantees co-residency instead of assuming it.
cudaError_t launch_scan(const ScanParams &p, bool stats, int grid, cudaStream_t st)
{
  void *args[] = {(void *) &p};
  const void *fn = stats ? (const void *) k_scan<true> : (const void *) k_scan<false>;
  return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(kThreads), args, sizeof(ScanSmem), st);
}

int scan_blocks_per_sm(bool stats)
{
  const void *fn = stats ? (const void *) k_scan<true> : (const void *) k_scan<false>;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int) sizeof(ScanSmem)) != cudaSuccess)
    return 0;
  int n = 0;
  cudaError_t e = stats
    ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_scan<true>, kThreads, sizeof(ScanSmem))
    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_scan<false>, kThreads, sizeof(ScanSmem));
  return (e == cudaSuccess && n > 0) ? n : 0;
}

}  // namespace smax
