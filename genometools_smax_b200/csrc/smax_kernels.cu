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

// CTA-wide decoupled look-back over (count, positions) pairs.  Every thread
// calls it with the tile's aggregates; returns the exclusive prefixes over
// tiles [0, tile).  Window = 256 predecessors per round, so even when all
// resident tiles finish at the same moment the chain resolves in ntiles/256
// rounds.
__device__ void lookback_exclusive(uint64_t *status, uint32_t tile, uint64_t agg_a,
                                   uint64_t agg_b, uint32_t epoch, uint64_t &excl_a,
                                   uint64_t &excl_b)
{
  __shared__ uint64_t s_wa[kThreads / 32], s_wb[kThreads / 32];
  __shared__ int s_wflag[kThreads / 32], s_wcut[kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  excl_a = 0; excl_b = 0;
  if (tile == 0)
  {
    if (tid == 0)
      st_pair(&status[0], pack_status(epoch, kStatePrefix, agg_a),
              pack_status(epoch, kStatePrefix, agg_b));
    return;
  }
  if (tid == 0)
    st_pair(&status[2 * (uint64_t) tile], pack_status(epoch, kStateAggregate, agg_a),
            pack_status(epoch, kStateAggregate, agg_b));
  int64_t hi = tile;                 // window = tiles [hi - 256, hi), nearest first
  for (;;)
  {
    const int64_t idx = hi - 1 - tid;
    uint64_t st = kStatePrefix, va = 0, vb = 0;   // virtual tiles < 0: prefix 0
    if (idx >= 0)
    {
      uint64_t wa, wb;
      ld_pair(&status[2 * idx], wa, wb);
      const uint64_t sa = (wa >> kValueBits) & 3, sb = (wb >> kValueBits) & 3;
      if ((uint32_t) (wa >> (kValueBits + 2)) == epoch &&
          (uint32_t) (wb >> (kValueBits + 2)) == epoch && sa == sb)
      {
        st = sa;
        va = wa & kValueMask;
        vb = wb & kValueMask;
      } else
        st = kStateInvalid;          // not published yet (or a torn pair: retry)
    }
    const unsigned inv = __ballot_sync(0xffffffffu, st == kStateInvalid);
    const unsigned pm = __ballot_sync(0xffffffffu, st == kStatePrefix);
    const int first_inv = inv ? __ffs(inv) - 1 : 32;
    const int first_p = pm ? __ffs(pm) - 1 : 32;
    int flag, cut;
    if (first_p < first_inv) { flag = 1; cut = first_p + 1; }   // reached a prefix
    else if (first_inv < 32) { flag = 2; cut = first_inv; }     // not published yet
    else { flag = 0; cut = 32; }                                // 32 aggregates
    uint64_t xa = lane < cut ? va : 0, xb = lane < cut ? vb : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
      xa += __shfl_xor_sync(0xffffffffu, xa, o);
      xb += __shfl_xor_sync(0xffffffffu, xb, o);
    }
    if (lane == 0)
    {
      s_wa[warp] = xa; s_wb[warp] = xb;
      s_wflag[warp] = flag; s_wcut[warp] = cut;
    }
    __syncthreads();
    uint64_t acc_a = 0, acc_b = 0;
    int outcome = 0, consumed = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++)
    {
      if (outcome == 0)
      {
        acc_a += s_wa[w]; acc_b += s_wb[w];
        consumed += s_wcut[w];
        outcome = s_wflag[w];
      }
    }
    __syncthreads();
    excl_a += acc_a; excl_b += acc_b;
    if (outcome == 1)
      break;
    hi -= consumed;
    if (outcome == 2 && consumed == 0)
      __nanosleep(40);
  }
  if (tid == 0)
    st_pair(&status[2 * (uint64_t) tile], pack_status(epoch, kStatePrefix, excl_a + agg_a),
            pack_status(epoch, kStatePrefix, excl_b + agg_b));
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
__device__ __noinline__ uint64_t value_at(const ScanParams &P, uint64_t q, bool &bad)
{
  const TableView *tv = view_for(P, q);
  if (tv == nullptr) { bad = true; return 0; }
  const uint32_t b = tv->lcp[q - tv->a_lo];
  if (b < 255)
    return b;
  uint64_t k;
  if (!llv_find(*tv, q, k)) { bad = true; return 255; }
  return tv->llv[k].value;
}

__device__ __noinline__ uint32_t byte_at_left(const ScanParams &P, uint64_t q, bool want_bwt,
                                              bool &bad)
{
  const TableView *tv = view_for(P, q);
  if (tv == nullptr) { bad = true; return 0; }
  return want_bwt ? tv->bwt[q - tv->a_lo] : tv->lcp[q - tv->a_lo];
}

// K2 in full generality: are the left characters bwt[lb..e] pairwise distinct?
__device__ __noinline__ bool left_distinct(const ScanParams &P, uint64_t lb, uint64_t e, bool &bad)
{
  const uint64_t a_lo = P.own.a_lo;
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  uint64_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
  for (uint64_t q = lb; q <= e; q++)
  {
    const uint32_t c = (q >= a_lo) ? (uint32_t) P.own.bwt[q - a_lo] : byte_at_left(P, q, true, bad);
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
  return !bad;
}

// shared memory of the scan kernel
struct ScanSmem
{
  uint64_t stage_v[kStageCap];
  uint64_t stage_w[kStageCap];
  uint64_t warp_tot[kThreads / 32];
  unsigned long long wsum;
  uint32_t bitmap[kTileWords];
  uint16_t wprefix[kTileWords];
  uint16_t stage_off[kStageCap];
  uint16_t order[kStageCap];
  uint32_t count;
  uint32_t tile;
};

__device__ __forceinline__ uint32_t rank_in_tile(const ScanSmem &sm, uint32_t o)
{
  return sm.wprefix[o >> 5] + __popc(sm.bitmap[o >> 5] & ((1u << (o & 31)) - 1u));
}

// win < 0: first pass over the tile -- mark the end offset, count, stage in
// arrival order.  win >= 0: replay -- ranks are known, stage rank window win
// in rank order.
__device__ __forceinline__ void emit_survivor(ScanSmem &sm, uint32_t o, uint64_t v, uint64_t width,
                                              int win)
{
  uint32_t slot;
  if (win < 0)
  {
    atomicOr(&sm.bitmap[o >> 5], 1u << (o & 31));
    atomicAdd(&sm.wsum, (unsigned long long) width);
    slot = atomicAdd(&sm.count, 1u);
  } else
    slot = rank_in_tile(sm, o) - (uint32_t) win * kStageCap;   // wraps for other windows
  if (slot < (uint32_t) kStageCap)
  {
    sm.stage_v[slot] = v;
    sm.stage_w[slot] = width;
    sm.stage_off[slot] = (uint16_t) o;
  }
}

// A plateau end e with small value b whose previous entry equals b: walk left
// over the run (128-bit compares), check the rise, then the left characters.
template <bool STATS>
__device__ __noinline__ bool examine_run(const ScanParams &P, uint64_t e, uint32_t b,
                                         uint64_t &width, bool &bad, uint64_t *stat)
{
  const uint8_t *lcp = P.own.lcp;
  const uint64_t a_lo = P.own.a_lo;
  const uint32_t v4 = b * 0x01010101u;
  uint64_t s = e;
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
    {
      pb = byte_at_left(P, q, false, bad);
      if (bad) return false;
    }
    if (pb == b) { s = q; continue; }
    if (pb > b)
      return false;
    break;
  }
  width = e - s + 2;
  if (STATS) { stat[0]++; stat[1] += width; }
  return left_distinct(P, s - 1, e, bad);
}

// A .llv record k (position p, value v >= minlength) inside the tile: is p the
// end of a local-maximum plateau of large values, and is it supermaximal?
template <bool STATS>
__device__ __forceinline__ bool examine_llv(const ScanParams &P, uint64_t k, uint64_t p, uint64_t v,
                                            uint64_t &width, bool &bad, uint64_t *stat)
{
  const TableView &own = P.own;
  const smax_llv *llv = own.llv;
  if (k + 1 < own.nllv)
  {
    const smax_llv nx = llv[k + 1];
    if (STATS) stat[2]++;
    if (nx.position == p + 1 && nx.value >= v)
      return false;                       // run continues or rises: not an end
  }
  uint64_t s = p, kk = k;
  for (;;)
  {
    const uint64_t q = s - 1;
    uint64_t pv;
    if (q >= own.a_lo)
    {
      if (kk == 0)
        break;                            // no record at q: a small value, rise
      const smax_llv pr = llv[kk - 1];
      if (STATS) stat[2]++;
      if (pr.position != q)
        break;
      pv = pr.value;
      kk--;
    } else
    {
      pv = value_at(P, q, bad);
      if (bad) return false;
    }
    if (pv == v) { s = q; continue; }
    if (pv > v)
      return false;
    break;
  }
  width = p - s + 2;
  if (STATS) { stat[0]++; stat[1] += width; }
  return left_distinct(P, s - 1, p, bad);
}

// candidates of one 16-byte chunk (small values), K1 tail + K2
template <bool STATS>
__device__ __noinline__ void process_chunk(const ScanParams &P, ScanSmem &sm, uint4 w,
                                           uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t prvb, uint64_t off, uint32_t coff, int win)
{
  const uint64_t a_lo = P.own.a_lo;
  const uint4 bw = *reinterpret_cast<const uint4 *>(P.own.bwt + off);
  const uint32_t words[4] = {w.x, w.y, w.z, w.w};
  const uint32_t bwords[4] = {bw.x, bw.y, bw.z, bw.w};
  const uint32_t masks[4] = {c0, c1, c2, c3};
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  uint64_t stat[4] = {0, 0, 0, 0};
  bool bad = false;
#pragma unroll
  for (int q = 0; q < 4; q++)
  {
    uint32_t m = masks[q];
    while (m)
    {
      const int bit = __ffs(m) - 1;          // 7, 15, 23 or 31
      m &= m - 1;
      const int sh = bit - 7;
      const uint32_t b = (words[q] >> sh) & 0xffu;
      if (b == 255)
        continue;                            // large values live in the .llv pass
      const uint32_t j = q * 4 + (sh >> 3);
      const uint64_t e = a_lo + off + j;
      if (e < P.g_lo || e >= P.g_hi)
        continue;
      const uint32_t pb = sh ? (words[q] >> (sh - 8)) & 0xffu
                             : (q ? words[q ? q - 1 : 0] >> 24 : prvb);
      if (pb > b)
        continue;
      uint64_t width = 2;
      bool ok;
      if (pb < b)
      {
        // width-2 plateau: the two left characters sit in the bwt chunk
        const uint32_t ch1 = (bwords[q] >> sh) & 0xffu;
        uint32_t ch0;
        if (sh) ch0 = (bwords[q] >> (sh - 8)) & 0xffu;
        else if (q) ch0 = bwords[q ? q - 1 : 0] >> 24;
        else ch0 = off > 0 ? (uint32_t) P.own.bwt[off - 1] : byte_at_left(P, a_lo - 1, true, bad);
        ok = gt_policy ? (ch0 != ch1 || ch0 >= 254) : (ch0 != ch1);
        if (STATS) { stat[0]++; stat[1] += 2; }
      } else
        ok = examine_run<STATS>(P, e, b, width, bad, stat);
      if (ok && !bad)
      {
        if (STATS) stat[3] += width;
        emit_survivor(sm, coff + j, b, width, win);
      }
    }
  }
  if (bad)
    P.result[kResError] = 1;
  if (STATS && win < 0)
  {
    if (stat[0]) atomicAdd((unsigned long long *) &P.result[kResStatCand], (unsigned long long) stat[0]);
    if (stat[1]) atomicAdd((unsigned long long *) &P.result[kResStatCandWidth], (unsigned long long) stat[1]);
    if (stat[3]) atomicAdd((unsigned long long *) &P.result[kResStatSurvWidth], (unsigned long long) stat[3]);
  }
}

// One detection pass over a tile (K1 + K2): small values from the lcp bytes,
// large values from the tile's .llv records.
template <bool STATS>
__device__ __noinline__ void tile_pass(const ScanParams &P, ScanSmem &sm, uint64_t toff, int win)
{
  const int tid = threadIdx.x, lane = tid & 31;
  const uint8_t *lcp = P.own.lcp;
  const uint64_t a_lo = P.own.a_lo;
  const uint64_t len16 = (P.own.a_hi - a_lo + 15) & ~15ull;      // loadable bytes
  const bool himode = P.mb > 128;
  const uint32_t kadd = (himode ? (0x100u - P.mb) : (0x80u - P.mb)) * 0x01010101u;

  // ---- small values: all loads first, then SWAR
  uint4 w[kItems];
#pragma unroll
  for (int c = 0; c < kItems; c++)
  {
    const uint64_t off = toff + (uint64_t) (c * kThreads + tid) * kChunk;
    w[c] = (off < len16) ? ld_stream(reinterpret_cast<const uint4 *>(lcp + off))
                         : make_uint4(0, 0, 0, 0);
  }
#pragma unroll
  for (int c = 0; c < kItems; c++)
  {
    const uint32_t h0 = swar_ge(w[c].x, kadd, himode), h1 = swar_ge(w[c].y, kadd, himode),
                   h2 = swar_ge(w[c].z, kadd, himode), h3 = swar_ge(w[c].w, kadd, himode);
    // neighbours across lanes (the warp covers 512 contiguous bytes)
    uint32_t nxtw = __shfl_down_sync(0xffffffffu, w[c].x, 1);
    const uint32_t prvw = __shfl_up_sync(0xffffffffu, w[c].w, 1);
    if (h0 | h1 | h2 | h3)
    {
      const uint32_t coff = (uint32_t) (c * kThreads + tid) * kChunk;
      const uint64_t off = toff + coff;
      if (lane == 31)
        nxtw = *reinterpret_cast<const uint32_t *>(lcp + off + 16);   // inside the zero pad
      const uint32_t n0 = __funnelshift_r(w[c].x, w[c].y, 8), n1 = __funnelshift_r(w[c].y, w[c].z, 8),
                     n2 = __funnelshift_r(w[c].z, w[c].w, 8), n3 = __funnelshift_r(w[c].w, nxtw, 8);
      const uint32_t c0 = swar_and_gt(h0, w[c].x, n0), c1 = swar_and_gt(h1, w[c].y, n1),
                     c2 = swar_and_gt(h2, w[c].z, n2), c3 = swar_and_gt(h3, w[c].w, n3);
      if (c0 | c1 | c2 | c3)
      {
        uint32_t prvb = prvw >> 24;
        if (lane == 0)
        {
          bool bad = false;
          prvb = off > 0 ? (uint32_t) lcp[off - 1]
                         : (a_lo > 0 ? byte_at_left(P, a_lo - 1, false, bad) : 0u);
          if (bad) P.result[kResError] = 1;
        }
        process_chunk<STATS>(P, sm, w[c], c0, c1, c2, c3, prvb, off, coff, win);
      }
    }
  }

  // ---- large values: the tile's slice of the .llv records
  if (P.own.nllv != 0)
  {
    const uint64_t tile_lo = a_lo + toff, tile_hi = tile_lo + kTileBytes;
    const uint64_t lo = tile_lo > P.g_lo ? tile_lo : P.g_lo;
    const uint64_t hi = tile_hi < P.g_hi ? tile_hi : P.g_hi;
    const uint64_t k0 = P.own.llvdir[toff >> kLlvBucketShift];
    const uint64_t k1 = P.own.llvdir[((toff + kTileBytes - 1) >> kLlvBucketShift) + 1];
    uint64_t stat[4] = {0, 0, 0, 0};
    bool bad = false;
    for (uint64_t k = k0 + tid; k < k1; k += kThreads)
    {
      const smax_llv r = P.own.llv[k];
      if (STATS) stat[2]++;
      if (r.position < lo || r.position >= hi || r.value < P.minlength)
        continue;
      uint64_t width;
      if (examine_llv<STATS>(P, k, r.position, r.value, width, bad, stat) && !bad)
      {
        if (STATS) stat[3] += width;
        emit_survivor(sm, (uint32_t) (r.position - tile_lo), r.value, width, win);
      }
    }
    if (bad)
      P.result[kResError] = 1;
    if (STATS && win < 0)
    {
      if (stat[0]) atomicAdd((unsigned long long *) &P.result[kResStatCand], (unsigned long long) stat[0]);
      if (stat[1]) atomicAdd((unsigned long long *) &P.result[kResStatCandWidth], (unsigned long long) stat[1]);
      if (stat[2]) atomicAdd((unsigned long long *) &P.result[kResStatLlv], (unsigned long long) stat[2]);
      if (stat[3]) atomicAdd((unsigned long long *) &P.result[kResStatSurvWidth], (unsigned long long) stat[3]);
    }
  }
}

__device__ __forceinline__ uint64_t suf_at(const ScanParams &P, uint64_t i, bool &bad)
{
  const TableView *tv = view_for(P, i);
  if (tv == nullptr || tv->suf == nullptr) { bad = true; return 0; }
  const uint64_t o = i - tv->a_lo;
  return P.sufbytes == 8 ? reinterpret_cast<const uint64_t *>(tv->suf)[o]
                         : (uint64_t) reinterpret_cast<const uint32_t *>(tv->suf)[o];
}

// K3 tail: write the staged survivors of one rank window in SA order and
// gather their positions.  Returns the number of positions of the window.
__device__ __forceinline__ uint64_t write_window(const ScanParams &P, ScanSmem &sm, uint32_t cnt,
                                                 uint64_t rec_base, uint64_t pos_base,
                                                 uint64_t tile_lo)
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
      const uint32_t slot = sm.order[i];
      wd[j] = sm.stage_w[slot];
      vv[j] = sm.stage_v[slot];
      oo[j] = sm.stage_off[slot];
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
  bool bad = false;
#pragma unroll
  for (int j = 0; j < kPer; j++)
  {
    const uint32_t i = tid * kPer + j;
    if (i < cnt)
    {
      const uint64_t lb = tile_lo + oo[j] + 1 - wd[j];
      const uint64_t dst = rec_base + i;
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
            P.positions[po + k] = suf_at(P, lb + k, bad);
        else
          P.result[kResOverflow] = 1;
      }
      po += wd[j];
    }
  }
  if (bad)
    P.result[kResError] = 1;
  return total;
}

// ------------------------------------------------------------ scan kernel
template <bool STATS>
__global__ void __launch_bounds__(kThreads, 4)
k_scan(const __grid_constant__ ScanParams P)
{
  __shared__ ScanSmem sm;
  const int tid = threadIdx.x;
  const uint64_t base_off = P.g_lo - P.own.a_lo;                 // multiple of 16
  bool dirty = true;

  for (;;)
  {
    __syncthreads();
    if (tid == 0)
    {
      sm.tile = atomicAdd(&P.ctrl[0], 1u);
      sm.count = 0;
      sm.wsum = 0;
    }
    if (dirty)
    {
      sm.bitmap[tid] = 0;
      sm.bitmap[tid + kThreads] = 0;
    }
    __syncthreads();
    const uint32_t tile = sm.tile;
    if (tile >= P.ntiles)
      break;
    const uint64_t toff = base_off + (uint64_t) tile * kTileBytes;

    tile_pass<STATS>(P, sm, toff, -1);
    __syncthreads();

    // ---- K3: ranks from the bitmap, look-back, ordered write + gather
    const uint32_t count = sm.count;
    const uint64_t wsum = sm.wsum;
    dirty = count != 0;
    if (count)
    {
      const uint32_t p0 = __popc(sm.bitmap[2 * tid]), p1 = __popc(sm.bitmap[2 * tid + 1]);
      uint32_t x = p0 + p1;
      const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
      }
      if (lane == 31)
        sm.warp_tot[warp] = x;
      __syncthreads();
      uint32_t wbase = 0;
#pragma unroll
      for (int k = 0; k < kThreads / 32; k++)
        if (k < warp) wbase += (uint32_t) sm.warp_tot[k];
      const uint32_t ex = wbase + x - (p0 + p1);
      sm.wprefix[2 * tid] = (uint16_t) ex;
      sm.wprefix[2 * tid + 1] = (uint16_t) (ex + p0);
    }
    uint64_t excl_c, excl_w;
    lookback_exclusive(P.status, tile, count, wsum, P.epoch, excl_c, excl_w);
    if (tile == P.ntiles - 1 && tid == 0)
    {
      P.result[kResCount] = excl_c + count;
      P.result[kResPositions] = excl_w + wsum;
    }
    if (count == 0)
      continue;
    __syncthreads();                     // wprefix visible
    const uint64_t tile_lo = P.own.a_lo + toff;
    if (count <= (uint32_t) kStageCap)
    {
      for (uint32_t slot = tid; slot < count; slot += kThreads)
        sm.order[rank_in_tile(sm, sm.stage_off[slot])] = (uint16_t) slot;
      __syncthreads();
      write_window(P, sm, count, excl_c, excl_w, tile_lo);
    } else
    {
      // more survivors than the stage holds: replay the tile window by window
      uint64_t pos_base = excl_w;
      const uint32_t nwin = (count + kStageCap - 1) / kStageCap;
      for (uint32_t win = 0; win < nwin; win++)
      {
        const uint32_t cnt = min((uint32_t) kStageCap, count - win * kStageCap);
        __syncthreads();
        tile_pass<false>(P, sm, toff, (int) win);
        for (uint32_t slot = tid; slot < (uint32_t) kStageCap; slot += kThreads)
          sm.order[slot] = (uint16_t) slot;
        __syncthreads();
        pos_base += write_window(P, sm, cnt, excl_c + (uint64_t) win * kStageCap, pos_base,
                                 tile_lo);
      }
    }
  }
  // the last CTA to leave re-arms the ticket for the next scan
  if (tid == 0)
  {
    __threadfence();
    const uint32_t done = atomicAdd(&P.ctrl[1], 1u);
    if (done == gridDim.x - 1)
    {
      P.ctrl[0] = 0;
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

cudaError_t launch_scan(const ScanParams &p, bool stats, int grid, cudaStream_t st)
{
  if (stats)
    k_scan<true><<<grid, kThreads, 0, st>>>(p);
  else
    k_scan<false><<<grid, kThreads, 0, st>>>(p);
  return cudaGetLastError();
}

int scan_blocks_per_sm(bool stats)
{
  int n = 0;
  cudaError_t e = stats
    ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_scan<true>, kThreads, 0)
    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_scan<false>, kThreads, 0);
  return (e == cudaSuccess && n > 0) ? n : 1;
}

}  // namespace smax
