/*
  smax_kernels.cu -- hand-written sm_100a kernels of the supermaximal-repeat
  scan.  ONE fused pass over the lcptab replaces the reference's stack sweep
  (/root/reference/src/match/esa-bottomup.c:116-273) and its per-node
  left-character bookkeeping (/root/reference/src/match/esa-maxpairs.c:181-360).

  k_scan, per 16 KiB tile of lcptab bytes (tiles handed out by an atomic
  ticket so that a tile's predecessors are always running or finished):

    K1  plateau detection.  Every lcp byte is read once with 128-bit streaming
        loads; neighbours across lanes come from warp shuffles.  SWAR byte
        arithmetic on the 32-bit words yields, per chunk, the exact mask of
        plateau ENDS with a small value:  byte >= min(minlength,255)  AND
        byte > next byte  (8 integer ops per 4 bytes, the second half only
        where the first fires).  The owner of an end looks at the previous
        byte: smaller -> the common width-2 plateau, equal -> walk left over
        the run with 128-bit compares, larger -> no local maximum.
        Large values (byte 255) are resolved in place in .llv RECORD space:
        each tile streams the .llv records that fall into it (found through a
        per-4096-entry directory, no global rank) with coalesced 16-byte
        loads; a record ends a plateau iff its right neighbour is no
        consecutive record with a value >= its own; runs are walked record by
        record.
    K2  left-distinctness over bwt[lb..e]: two bytes straight from the
        prefetched bwt chunk for width 2, a 256-bit alphabet mask otherwise;
        specials (>= 254) never collide under the GenomeTools convention
        (esa-maxpairs.c:24-31).
    K3  order-preserving compaction + emit.  A survivor sets the bit of its end
        offset in a per-tile bitmap (rank = popcount prefix) and is staged in
        shared memory; tile totals (record count, position count) are chained
        with a CTA-wide decoupled look-back over epoch-tagged 16-byte status
        pairs (no memset between scans); records are then written in
        suffix-array order and the occurrence positions suf[lb..lb+width) are
        gathered right behind them.  A tile with more survivors than the stage
        holds is replayed in rank windows.

  k_llvdir builds the .llv bucket directory at upload time.
*/
#include "smax_kernels.cuh"

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

// streaming 128-bit load of table bytes: read-only path, do not keep in L1
__device__ __forceinline__ uint4 ld_stream(const uint4 *p)
{
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ uint64_t pack_status(uint32_t epoch, uint64_t state, uint64_t value)
{
  return ((uint64_t) epoch << (kValueBits + 2)) | (state << kValueBits) | (value & kValueMask);
}

// SWAR: bit 7 of every byte of w that is >= mb (kadd / himode derived from mb)
__device__ __forceinline__ uint32_t swar_ge(uint32_t w, uint32_t kadd, bool himode)
{
  const uint32_t t = (w & 0x7f7f7f7fu) + kadd;
  return (himode ? (t & w) : (t | w)) & 0x80808080u;
}

// SWAR: h & (bit 7 of every byte where x > y), h being a subset of 0x80808080
__device__ __forceinline__ uint32_t swar_and_gt(uint32_t h, uint32_t x, uint32_t y)
{
  const uint32_t t = (y | 0x80808080u) - (x & 0x7f7f7f7fu);   // bit7: low7(y) >= low7(x)
  const uint32_t ge_yx = (y & ~x) | (~(y ^ x) & t);           // bit7: y >= x
  return h & ~ge_yx;
}

// SWAR: h & (bit 7 of every byte where x >= y)
__device__ __forceinline__ uint32_t swar_and_ge(uint32_t h, uint32_t x, uint32_t y)
{
  const uint32_t t = (x | 0x80808080u) - (y & 0x7f7f7f7fu);   // bit7: low7(x) >= low7(y)
  return h & ((x & ~y) | (~(x ^ y) & t));
}

// ------------------------------------------------- ordered prefix exchange
// Tiles are assigned round-robin: CTA b owns tiles b, b+grid, b+2*grid, ...
// ("generation" g = tiles [g*grid, (g+1)*grid)), and the whole grid is
// resident, so every generation is worked on by all CTAs at the same time.
// Each tile publishes its aggregate pair (records, positions) as soon as its
// detection pass is done; a CTA obtains the exclusive prefix of its tile by
// reading the <= grid aggregates of its generation in one round (no chain of
// dependent prefixes as in a tile-by-tile look-back) and carries the totals of
// all earlier generations in registers.  The read is deferred by one tile, so
// the aggregates have normally all arrived and nobody spins.
__device__ __forceinline__ void publish_aggregate(uint64_t *status, uint32_t tile, uint64_t agg_a,
                                                  uint64_t agg_b, uint32_t epoch)
{
  st_pair(&status[2 * (uint64_t) tile], pack_status(epoch, kStateAggregate, agg_a),
          pack_status(epoch, kStateAggregate, agg_b));
}

// Every thread of the CTA calls this; warp 0 reads the <= grid aggregates of
// the generation (loads issued four at a time), the other warps wait at the one
// barrier without issuing.  first = first tile of the generation, ng = tiles in
// it, mine = index of the caller's tile within it.
__device__ __forceinline__ void resolve_generation(const uint64_t *status, uint32_t first,
                                                   uint32_t ng, uint32_t mine, uint32_t epoch,
                                                   uint64_t *red, uint64_t &excl_a,
                                                   uint64_t &excl_b, uint64_t &tot_a,
                                                   uint64_t &tot_b)
{
  const int tid = threadIdx.x;
  if (tid < 32)
  {
    uint64_t ea = 0, eb = 0, ta = 0, tb = 0;
    for (uint32_t j0 = tid; j0 < ng; j0 += 4 * 32)
    {
      uint64_t wa[4], wb[4];
#pragma unroll
      for (int r = 0; r < 4; r++)             // all loads first, then the checks
      {
        const uint32_t j = j0 + r * 32;
        wa[r] = wb[r] = 0;
        if (j < ng)
          ld_pair(&status[2 * (uint64_t) (first + j)], wa[r], wb[r]);
      }
#pragma unroll
      for (int r = 0; r < 4; r++)
      {
        const uint32_t j = j0 + r * 32;
        if (j < ng)
        {
          unsigned backoff = 32;
          while ((uint32_t) (wa[r] >> (kValueBits + 2)) != epoch ||
                 (uint32_t) (wb[r] >> (kValueBits + 2)) != epoch)
          {
            __nanosleep(backoff);              // a straggler has not published yet
            backoff = min(backoff * 2u, 1024u);
            ld_pair(&status[2 * (uint64_t) (first + j)], wa[r], wb[r]);
          }
          const uint64_t va = wa[r] & kValueMask, vb = wb[r] & kValueMask;
          ta += va; tb += vb;
          if (j < mine) { ea += va; eb += vb; }
        }
      }
    }
    if (__any_sync(0xffffffffu, (ta | tb) != 0))   // sparse index: mostly all zero
    {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1)
      {
        ea += __shfl_xor_sync(0xffffffffu, ea, o);
        eb += __shfl_xor_sync(0xffffffffu, eb, o);
        ta += __shfl_xor_sync(0xffffffffu, ta, o);
        tb += __shfl_xor_sync(0xffffffffu, tb, o);
      }
    }
    if (tid == 0)
    {
      red[0] = ea; red[1] = eb; red[2] = ta; red[3] = tb;
    }
  }
  __syncthreads();
  excl_a = red[0]; excl_b = red[1]; tot_a = red[2]; tot_b = red[3];
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
// Per-tile state is double buffered: while the survivors of tile t wait for
// their prefix, the CTA already detects in tile t + grid.
struct TileState
{
  uint64_t stage_v[kStageCap];     // staged survivors: value,
  uint64_t stage_w[kStageCap];     //   SA width,
  uint32_t bitmap[kTileWords];     // one bit per lcp entry of the tile: survivor ends here
  uint16_t stage_off[kStageCap];   //   end offset in the tile
  uint16_t wprefix[kTileWords];    // exclusive popcount prefix of bitmap words
  uint16_t order[kStageCap];       // rank -> stage slot
  unsigned long long wsum;         // widths of survivors that did not fit the stage
  uint32_t count;                  // survivors of the tile
};

constexpr int kHalo = 16;               // staged lcp bytes either side of the tile

struct ScanSmem
{
  // a tile with many plateau ends keeps its lcp and bwt bytes (+ 16 either side)
  // here, so that the per-candidate work reads neighbours and left characters
  // at shared-memory latency instead of gathering sectors from L2 / HBM
  alignas(16) uint8_t lcp_tile[kHalo + kTileBytes + kHalo];
  alignas(16) uint8_t bwt_tile[kHalo + kTileBytes + kHalo];
  TileState ts[2];
  uint64_t red[kThreads / 32 * 4];
  uint64_t warp_tot[kThreads / 32];
  uint16_t endmap[kTileBytes / kChunk];   // small-value plateau ends: 16 bits per 16-byte chunk
};

__device__ __forceinline__ uint32_t rank_in_tile(const TileState &T, uint32_t o)
{
  return T.wprefix[o >> 5] + __popc(T.bitmap[o >> 5] & ((1u << (o & 31)) - 1u));
}

// win < 0: first pass over the tile -- mark the end offset, count, stage in
// arrival order.  win >= 0: replay -- ranks are known, stage rank window win
// in rank order.
__device__ __forceinline__ void emit_survivor(TileState &T, uint32_t o, uint64_t v, uint64_t width,
                                              int win)
{
  uint32_t slot;
  if (win < 0)
  {
    atomicOr(&T.bitmap[o >> 5], 1u << (o & 31));
    slot = atomicAdd(&T.count, 1u);
    if (slot >= (uint32_t) kStageCap)     // staged widths are summed later; only the
      atomicAdd(&T.wsum, (unsigned long long) width);   // overflow needs the (slow) 64-bit atomic
  } else
    slot = rank_in_tile(T, o) - (uint32_t) win * kStageCap;   // wraps for other windows
  if (slot < (uint32_t) kStageCap)
  {
    T.stage_v[slot] = v;
    T.stage_w[slot] = width;
    T.stage_off[slot] = (uint16_t) o;
  }
}

// K2 for one candidate plateau [lb, e]: left characters pairwise distinct?
// Short plateaus inside the shard's own arrays load all their bwt bytes at once.
__device__ __forceinline__ bool candidate_survives(const ScanParams &P, uint64_t lb, uint64_t e,
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

// A plateau end e with small value b: look at the previous entries.  Returns
// the SA width of the local-maximum plateau ending at e, or 0 if the run that
// ends at e is entered from a larger value.  Runs are walked with 128-bit
// compares in global memory (used once a run leaves the staged tile).
__device__ __forceinline__ uint64_t small_plateau_width_global(const ScanParams &P, uint64_t e,
                                                            uint64_t s, uint32_t b)
{
  const uint8_t *lcp = P.own.lcp;
  const uint64_t a_lo = P.own.a_lo;
  const uint32_t v4 = b * 0x01010101u;
  for (;;)
  {
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

// The same walk inside a staged tile: o = offset of the end in the tile,
// st[kHalo + i] = lcp[tile_lo + i] for i in [low, kTileBytes + kHalo).
__device__ __forceinline__ uint64_t small_plateau_width_staged(const ScanParams &P,
                                                               const uint8_t *st, int low,
                                                               uint64_t tile_lo, uint32_t o,
                                                               uint32_t b)
{
  int i = (int) o;                       // run start candidate, tile offset (may go down to low)
  for (;;)
  {
    if (i == low)                        // staged range exhausted: continue in global memory
      return small_plateau_width_global(P, tile_lo + o, tile_lo + i, b);
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
                                                          int low, uint64_t tile_lo, uint32_t o,
                                                          uint64_t width)
{
  if (width <= 4 && (int) o + 1 - (int) width >= low)
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

// A .llv record k (position p, value v): does a run of large values end at p,
// i.e. is the next entry no consecutive record with a value >= v?
__device__ __forceinline__ bool llv_is_end(const ScanParams &P, uint64_t k, uint64_t p, uint64_t v)
{
  if (k + 1 < P.own.nllv)
  {
    const smax_llv nx = P.own.llv[k + 1];
    if (nx.position == p + 1 && nx.value >= v)
      return false;
  }
  return true;
}

// SA width of the local-maximum plateau of large values ending at record k, or
// 0 if the run is entered from a larger value.  Runs are walked in record space.
__device__ __forceinline__ uint64_t llv_plateau_width(const ScanParams &P, uint64_t k, uint64_t p,
                                                      uint64_t v)
{
  const smax_llv *llv = P.own.llv;
  uint64_t s = p, kk = k;
  for (;;)
  {
    const uint64_t q = s - 1;
    uint64_t pv;
    if (q >= P.own.a_lo)
    {
      if (kk == 0)
        break;                            // no record at q: a small value, rise
      const smax_llv pr = llv[kk - 1];
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

// K2 + emit for one local-maximum plateau [e + 1 - width, e] of value v.
template <bool STATS>
__device__ __forceinline__ void test_and_emit(const ScanParams &P, ScanSmem &sm, TileState &T,
                                              uint32_t o, uint64_t v, uint64_t width, bool dense,
                                              int low, uint64_t tile_lo, int win, uint64_t *stat)
{
  if (STATS) { stat[0]++; stat[1] += width; }
  if (P.debug & 16)
    return;
  const uint64_t e = tile_lo + o;
  const bool ok = dense ? candidate_survives_staged(P, sm.bwt_tile, low, tile_lo, o, width)
                        : candidate_survives(P, e + 1 - width, e, width);
  if (ok)
  {
    if (STATS) stat[3] += width;
    if ((P.debug & 32) == 0)
      emit_survivor(T, o, v, width, win);
  }
}

// bits 7,15,23,31 of c -> bits 0..3
__device__ __forceinline__ uint32_t pack_ends(uint32_t c)
{
  return (((c >> 7) & 0x01010101u) * 0x01020408u) >> 24;
}

__device__ __forceinline__ void load_tile(const ScanParams &P, uint64_t toff, uint4 (&w)[kItems])
{
  const uint64_t len16 = (P.own.a_hi - P.own.a_lo + 15) & ~15ull;      // loadable bytes
#pragma unroll
  for (int c = 0; c < kItems; c++)
  {
    const uint64_t off = toff + (uint64_t) (c * kThreads + threadIdx.x) * kChunk;
    w[c] = (off < len16) ? ld_stream(reinterpret_cast<const uint4 *>(P.own.lcp + off))
                         : make_uint4(0, 0, 0, 0);
  }
}

// One detection pass over a tile.  K1 on the lcp bytes in w[]: SWAR yields the
// exact 16-bit mask of small-value plateau ends of every 16-byte chunk, which is
// parked in a shared end bitmap (no atomics, no capacity limit).  One barrier
// later the CTA knows whether the tile has any end at all (most tiles of a
// sparse index have none); if so, every thread finishes the ends of its own two
// bitmap words, lanes side by side.  Large values are found in .llv record
// space and finished through a small queue.
template <bool STATS>
__device__ __forceinline__ void tile_pass(const ScanParams &P, ScanSmem &sm, TileState &T,
                                          uint64_t toff, int win, uint4 (&w)[kItems])
{
  const int tid = threadIdx.x, lane = tid & 31;
  const uint8_t *lcp = P.own.lcp;
  const uint64_t a_lo = P.own.a_lo;
  const uint64_t tile_lo = a_lo + toff;
  const bool himode = P.mb > 128;
  const uint32_t kadd = (himode ? (0x100u - P.mb) : (0x80u - P.mb)) * 0x01010101u;
  uint64_t stat[4] = {0, 0, 0, 0};
  uint32_t any_end = 0;

#pragma unroll
  for (int c = 0; c < kItems; c++)
  {
    const uint32_t h0 = swar_ge(w[c].x, kadd, himode), h1 = swar_ge(w[c].y, kadd, himode),
                   h2 = swar_ge(w[c].z, kadd, himode), h3 = swar_ge(w[c].w, kadd, himode);
    // neighbours across lanes (the warp covers 512 contiguous bytes)
    uint32_t nxtw = __shfl_down_sync(0xffffffffu, w[c].x, 1);
    uint32_t prvw = __shfl_up_sync(0xffffffffu, w[c].w, 1);
    const uint32_t coff = (uint32_t) (c * kThreads + tid) * kChunk;
    uint32_t m16 = 0;
    if ((h0 | h1 | h2 | h3) && !(P.debug & 2))
    {
      if (lane == 31)
        nxtw = *reinterpret_cast<const uint32_t *>(lcp + toff + coff + 16);   // inside the zero pad
      const uint32_t n0 = __funnelshift_r(w[c].x, w[c].y, 8), n1 = __funnelshift_r(w[c].y, w[c].z, 8),
                     n2 = __funnelshift_r(w[c].z, w[c].w, 8), n3 = __funnelshift_r(w[c].w, nxtw, 8);
      // plateau ends: >= threshold, > next, not an overflow byte (those live in .llv)
      uint32_t c0 = swar_and_gt(h0, w[c].x, n0) & ~(((w[c].x & 0x7f7f7f7fu) + 0x01010101u) & w[c].x);
      uint32_t c1 = swar_and_gt(h1, w[c].y, n1) & ~(((w[c].y & 0x7f7f7f7fu) + 0x01010101u) & w[c].y);
      uint32_t c2 = swar_and_gt(h2, w[c].z, n2) & ~(((w[c].z & 0x7f7f7f7fu) + 0x01010101u) & w[c].z);
      uint32_t c3 = swar_and_gt(h3, w[c].w, n3) & ~(((w[c].w & 0x7f7f7f7fu) + 0x01010101u) & w[c].w);
      if (c0 | c1 | c2 | c3)
      {
        // ... and not entered from a larger value (>= previous byte)
        if (lane == 0)
        {
          const uint64_t off = toff + coff;
          prvw = off >= 4 ? *reinterpret_cast<const uint32_t *>(lcp + off - 4)
                          : (a_lo > 0 ? byte_at_left(P, a_lo - 1, false) << 24 : 0u);
        }
        c0 = swar_and_ge(c0, w[c].x, __funnelshift_l(prvw, w[c].x, 8));
        c1 = swar_and_ge(c1, w[c].y, __funnelshift_l(w[c].x, w[c].y, 8));
        c2 = swar_and_ge(c2, w[c].z, __funnelshift_l(w[c].y, w[c].z, 8));
        c3 = swar_and_ge(c3, w[c].w, __funnelshift_l(w[c].z, w[c].w, 8));
        m16 = pack_ends(c0) | (pack_ends(c1) << 4) | (pack_ends(c2) << 8) | (pack_ends(c3) << 12);
      }
    }
    sm.endmap[c * kThreads + tid] = (uint16_t) m16;
    any_end |= m16;
  }
  // ---- finish the small-value ends: thread t owns bitmap words t and t + 256
  const int nthr_with_ends = __syncthreads_count((int) any_end);
  // a tile where many threads found ends (repeat-rich region of the index), or
  // one that holds large values, stages its lcp and bwt bytes in shared memory
  uint64_t k0 = 0, k1 = 0;
  if (P.own.nllv != 0 && !(P.debug & (2 | 4)))
  {
    k0 = P.own.llvdir[toff >> kLlvBucketShift];
    k1 = P.own.llvdir[((toff + kTileBytes - 1) >> kLlvBucketShift) + 1];
  }
  const bool dense = nthr_with_ends >= 32 || k1 - k0 >= 64;
  const int low = toff >= (uint64_t) kHalo ? -kHalo : 0;       // staged range starts here
  if (dense)
  {
    const uint8_t *bwt = P.own.bwt;
    const uint64_t len16 = (P.own.a_hi - a_lo + 15) & ~15ull;
    if (tid < 4)
    {
      // halos: 16 bytes left and right of the tile, both tables
      const bool left = (tid & 1) == 0;
      const uint8_t *src = (tid < 2 ? lcp : bwt) + (left ? toff - kHalo : toff + kTileBytes);
      uint4 hv = make_uint4(0, 0, 0, 0);
      if (left ? low < 0 : toff + kTileBytes < len16 + 48)
        hv = *reinterpret_cast<const uint4 *>(src);
      *reinterpret_cast<uint4 *>((tid < 2 ? sm.lcp_tile : sm.bwt_tile) +
                                 (left ? 0 : kHalo + kTileBytes)) = hv;
    }
#pragma unroll
    for (int c = 0; c < kItems; c++)
    {
      const uint32_t coff = (uint32_t) (c * kThreads + tid) * kChunk;
      *reinterpret_cast<uint4 *>(sm.lcp_tile + kHalo + coff) = w[c];
      *reinterpret_cast<uint4 *>(sm.bwt_tile + kHalo + coff) =
        toff + coff < len16 ? *reinterpret_cast<const uint4 *>(bwt + toff + coff)
                            : make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
  }
  if (nthr_with_ends)
  {
    const uint32_t *endwords = reinterpret_cast<const uint32_t *>(sm.endmap);
#pragma unroll 1
    for (int half = 0; half < 2; half++)
    {
      const uint32_t wi = tid + half * kThreads;
      uint32_t bits = endwords[wi];
      while (bits)
      {
        const uint32_t o = wi * 32 + (__ffs(bits) - 1);
        bits &= bits - 1;
        const uint64_t e = tile_lo + o;
        if (e < P.g_hi)                       // later ends belong to the next shard
        {
          const uint32_t b = dense ? sm.lcp_tile[kHalo + o] : lcp[toff + o];
          const uint64_t width = dense ? small_plateau_width_staged(P, sm.lcp_tile, low, tile_lo, o, b)
                                       : small_plateau_width_global(P, e, e, b);
          if (width != 0)
            test_and_emit<STATS>(P, sm, T, o, b, width, dense, low, tile_lo, win, stat);
        }
      }
    }
  }

  // ---- large values: the tile's slice of the .llv records.  Four rounds of
  // 256 records are requested together; a record's neighbours come from the
  // adjacent lanes, so the common case (run of length 1) needs no further load.
  if (k0 < k1)
  {
    const smax_llv *llv = P.own.llv;
    const uint64_t tile_hi = tile_lo + kTileBytes;
    const uint64_t lo = tile_lo > P.g_lo ? tile_lo : P.g_lo;
    const uint64_t hi = tile_hi < P.g_hi ? tile_hi : P.g_hi;
    const uint64_t none = ~0ull;
    for (uint64_t kb = k0; kb < k1; kb += 4 * kThreads)
    {
      uint64_t rp[4], rv[4], ep[4], ev[4];       // own record; edge lanes: outer neighbour
#pragma unroll
      for (int j = 0; j < 4; j++)
      {
        const uint64_t k = kb + (uint64_t) j * kThreads + tid;
        rp[j] = none; rv[j] = 0; ep[j] = none; ev[j] = 0;
        if (k <= k1 && k < P.own.nllv)          // k1 itself: right neighbour of the last record
        {
          const smax_llv r = llv[k];
          rp[j] = r.position; rv[j] = r.value;
          if (lane == 31 && k + 1 < P.own.nllv) { const smax_llv x = llv[k + 1]; ep[j] = x.position; ev[j] = x.value; }
          if (lane == 0 && k > 0) { const smax_llv x = llv[k - 1]; ep[j] = x.position; ev[j] = x.value; }
        }
      }
#pragma unroll
      for (int j = 0; j < 4; j++)
      {
        const uint64_t k = kb + (uint64_t) j * kThreads + tid;
        uint64_t np = __shfl_down_sync(0xffffffffu, rp[j], 1), nv = __shfl_down_sync(0xffffffffu, rv[j], 1);
        uint64_t pp = __shfl_up_sync(0xffffffffu, rp[j], 1), pv = __shfl_up_sync(0xffffffffu, rv[j], 1);
        if (lane == 31) { np = ep[j]; nv = ev[j]; }
        if (lane == 0) { pp = ep[j]; pv = ev[j]; }
        const uint64_t pos = rp[j], val = rv[j];
        if (STATS && k < k1) stat[2]++;
        if (k < k1 && pos >= lo && pos < hi && val >= P.minlength &&
            !(np == pos + 1 && nv >= val))       // the run of large values ends here
        {
          uint64_t width = 2;                    // previous entry is a smaller value
          if (pos == a_lo || (pp == pos - 1 && pv == val))
            width = llv_plateau_width(P, k, pos, val);   // run of equal values / shard edge
          else if (pp == pos - 1 && pv > val)
            width = 0;                           // entered from a larger value
          if (width != 0)
            test_and_emit<STATS>(P, sm, T, (uint32_t) (pos - tile_lo), val, width, dense, low, tile_lo,
                                 win, stat);
        }
      }
    }
  }
  if (STATS && win < 0)
  {
    if (stat[0]) atomicAdd((unsigned long long *) &P.result[kResStatCand], (unsigned long long) stat[0]);
    if (stat[1]) atomicAdd((unsigned long long *) &P.result[kResStatCandWidth], (unsigned long long) stat[1]);
    if (stat[2]) atomicAdd((unsigned long long *) &P.result[kResStatLlv], (unsigned long long) stat[2]);
    if (stat[3]) atomicAdd((unsigned long long *) &P.result[kResStatSurvWidth], (unsigned long long) stat[3]);
  }
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

// K3 tail: write the staged survivors of one rank window in SA order and
// gather their positions.  Returns the number of positions of the window.
__device__ __forceinline__ uint64_t write_window(const ScanParams &P, ScanSmem &sm, TileState &T,
                                                 uint32_t cnt, uint64_t rec_base,
                                                 uint64_t pos_base, uint64_t tile_lo)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kPer = kStageCap / kThreads;       // consecutive ranks per thread
  uint64_t wd[kPer], vv[kPer];
  uint32_t oo[kPer];
  uint64_t tsum = 0;
#pragma unroll
  for (int j = 0; j < kPer; j++)
  {
    const uint32_t i = tid * kPer + j;
    wd[j] = 0; vv[j] = 0; oo[j] = 0;
    if (i < cnt)
    {
      const uint32_t slot = T.order[i];
      wd[j] = T.stage_w[slot];
      vv[j] = T.stage_v[slot];
      oo[j] = T.stage_off[slot];
    }
    tsum += wd[j];
  }
  uint64_t x = tsum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  __syncthreads();                                 // warp_tot may still be in use
  if (lane == 31)
    sm.warp_tot[warp] = x;
  __syncthreads();
  uint64_t wbase = 0, total = 0;
#pragma unroll
  for (int k = 0; k < kThreads / 32; k++)
  {
    if (k < warp) wbase += sm.warp_tot[k];
    total += sm.warp_tot[k];
  }
  uint64_t po = pos_base + wbase + x - tsum;
  const bool gather = P.positions != nullptr;
#pragma unroll
  for (int j = 0; j < kPer; j++)
  {
    const uint32_t i = tid * kPer + j;
    if (i < cnt)
    {
      const uint64_t lb = tile_lo + oo[j] + 1 - wd[j];
      const uint64_t dst = rec_base + i;
      if (wd[j] < 2 || wd[j] > tile_lo + oo[j] + 1)
      {
        P.result[kResError] = 2;           // a stage slot that no survivor filled
        continue;
      }
      if (dst < P.rec_capacity)
      {
        smax_record r;
        r.len = vv[j]; r.lb = lb; r.width = wd[j];
        P.recs[dst] = r;
      } else
        P.result[kResOverflow] = 1;
      if (gather)
      {
        if (po + wd[j] <= P.pos_capacity)
          for (uint64_t k = 0; k < wd[j]; k++)
            P.positions[po + k] = suf_at(P, lb + k);
        else
          P.result[kResOverflow] = 1;
      }
      po += wd[j];
    }
  }
  return total;
}

// aggregates of a tile after its detection pass: survivor count is T.count;
// returns the number of positions and prepares the rank table
__device__ __forceinline__ uint64_t tile_aggregate(ScanSmem &sm, TileState &T)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t count = T.count;
  if (count == 0)
    return 0;
  // position count of the tile: staged widths (block reduction) + overflow
  const uint32_t staged = min(count, (uint32_t) kStageCap);
  uint64_t part = 0;
  for (uint32_t slot = tid; slot < staged; slot += kThreads)
    part += T.stage_w[slot];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    part += __shfl_xor_sync(0xffffffffu, part, o);
  const uint32_t p0 = __popc(T.bitmap[2 * tid]), p1 = __popc(T.bitmap[2 * tid + 1]);
  uint32_t x = p0 + p1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
  {
    const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  __syncthreads();
  if (lane == 0)
    sm.red[warp] = part;
  if (lane == 31)
    sm.warp_tot[warp] = x;
  __syncthreads();
  uint64_t wsum = T.wsum;
  uint32_t wbase = 0;
#pragma unroll
  for (int k = 0; k < kThreads / 32; k++)
  {
    wsum += sm.red[k];
    if (k < warp) wbase += (uint32_t) sm.warp_tot[k];
  }
  const uint32_t ex = wbase + x - (p0 + p1);
  T.wprefix[2 * tid] = (uint16_t) ex;
  T.wprefix[2 * tid + 1] = (uint16_t) (ex + p0);
  return wsum;
}

// ------------------------------------------------------------ scan kernel
template <bool STATS>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
k_scan(const __grid_constant__ ScanParams P)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ScanSmem &sm = *reinterpret_cast<ScanSmem *>(smem_raw);
  const int tid = threadIdx.x;
  const uint32_t grid = gridDim.x, me = blockIdx.x;
  const uint64_t base_off = P.g_lo - P.own.a_lo;                 // multiple of 16

  // clean state
  for (int b = 0; b < 2; b++)
  {
    sm.ts[b].bitmap[tid] = 0;
    sm.ts[b].bitmap[tid + kThreads] = 0;
    if (tid == 0) { sm.ts[b].count = 0; sm.ts[b].wsum = 0; }
  }

  uint64_t gen_c = 0, gen_w = 0;          // records / positions of all finished generations
  uint32_t pend_tile = 0, pend_count = 0; // tile whose survivors still wait for their prefix
  bool pending = false;
  uint4 w[kItems], wn[kItems];     // chunks of the current tile / of this CTA's next tile
  uint32_t tile = me;
  uint32_t gen = 0;
  if (tile < P.ntiles)
    load_tile(P, base_off + (uint64_t) tile * kTileBytes, w);
  __syncthreads();

  for (;; tile += grid, gen++)
  {
    const bool have_tile = tile < P.ntiles;
    TileState &T = sm.ts[gen & 1];
    if (have_tile)
    {
      const uint64_t toff = base_off + (uint64_t) tile * kTileBytes;
      // two tiles in flight per CTA: the chunks of the next tile are requested
      // before this one is touched
      if (tile + grid < P.ntiles)
        load_tile(P, toff + (uint64_t) grid * kTileBytes, wn);
      tile_pass<STATS>(P, sm, T, toff, -1, w);
      __syncthreads();
      const uint64_t wsum = tile_aggregate(sm, T);
      if (tid == 0)
        publish_aggregate(P.status, tile, T.count, wsum, P.epoch);
    }
    // ---- deferred K3 for the previous tile of this CTA (generation gen - 1)
    if (pending)
    {
      TileState &Tp = sm.ts[(gen & 1) ^ 1];
      const uint32_t g = gen - 1;
      const uint32_t first = g * grid;
      const uint32_t ng = min(grid, P.ntiles - first);
      uint64_t excl_c = 0, excl_w = 0, tot_c = 0, tot_w = 0;
      if (!(P.debug & 1))
        resolve_generation(P.status, first, ng, me, P.epoch, sm.red, excl_c, excl_w, tot_c, tot_w);
      excl_c += gen_c; excl_w += gen_w;
      gen_c += tot_c; gen_w += tot_w;
      if (pend_tile == P.ntiles - 1 && tid == 0)
      {
        P.result[kResCount] = gen_c;
        P.result[kResPositions] = gen_w;
      }
      const uint32_t count = pend_count;
      if (count)
      {
        const uint64_t ptoff = base_off + (uint64_t) pend_tile * kTileBytes;
        const uint64_t tile_lo = P.own.a_lo + ptoff;
        __syncthreads();
        if (count <= (uint32_t) kStageCap)
        {
          for (uint32_t slot = tid; slot < count; slot += kThreads)
            Tp.order[rank_in_tile(Tp, Tp.stage_off[slot])] = (uint16_t) slot;
          __syncthreads();
          write_window(P, sm, Tp, count, excl_c, excl_w, tile_lo);
        } else
        {
          // more survivors than the stage holds: replay the tile one rank window
          // at a time (the prefetched chunks of the next tile are re-loaded after)
          uint64_t pos_base = excl_w;
          const uint32_t nwin = (count + kStageCap - 1) / kStageCap;
          uint4 wr[kItems];
          for (uint32_t win = 0; win < nwin; win++)
          {
            __syncthreads();
            load_tile(P, ptoff, wr);
            tile_pass<false>(P, sm, Tp, ptoff, (int) win, wr);
            for (uint32_t slot = tid; slot < (uint32_t) kStageCap; slot += kThreads)
              Tp.order[slot] = (uint16_t) slot;
            __syncthreads();
            const uint32_t cnt = min((uint32_t) kStageCap, count - win * kStageCap);
            pos_base += write_window(P, sm, Tp, cnt, excl_c + (uint64_t) win * kStageCap, pos_base,
                                     tile_lo);
          }
        }
        __syncthreads();
        // leave the buffer clean for the tile after next
        Tp.bitmap[tid] = 0;
        Tp.bitmap[tid + kThreads] = 0;
        if (tid == 0) { Tp.count = 0; Tp.wsum = 0; }
      }
      pending = false;
    }
    if (!have_tile)
      break;
    pending = true;
    pend_tile = tile;
    pend_count = T.count;
#pragma unroll
    for (int c = 0; c < kItems; c++)
      w[c] = wn[c];
    __syncthreads();                 // buffers / counters cleaned above are visible
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
