/*
  smax_kernels.cu -- hand-written sm_100a kernels of the supermaximal-repeat
  scan.  One fused pass over the lcptab replaces the reference's stack sweep
  (/root/reference/src/match/esa-bottomup.c:116-273) and its per-node
  left-character bookkeeping (/root/reference/src/match/esa-maxpairs.c:181-360).

  k_scan   (K1+K2+K3)  per 16 KiB tile of lcptab bytes:
      K1  plateau detection.  Every byte is read once with 128-bit loads.  A
          SWAR filter (3 integer ops per 4 bytes) keeps bytes >= min(minlength,
          255); only those are examined: byte e ends a local-maximum plateau
          iff L[e] > L[e+1]; its owner walks left over the run L[s..e] == v and
          needs L[s-1] < v.  255-bytes are resolved in place through the
          position-sorted .llv records, located with a per-4096-entry
          directory (no global rank/scan needed); runs of large values are
          walked in .llv record space.
      K2  left-distinctness over bwt[s-1..e] with an alphabet bitmask
          (256 bits; specials >= 254 never collide under the GenomeTools
          convention).
      K3  order-preserving compaction: survivors set a bit at their end offset
          in a per-tile bitmap (rank = popcount prefix) and are staged in
          shared memory; tile totals are chained with a CTA-wide decoupled
          look-back over epoch-tagged status words, so records land in
          suffix-array order in one pass and no memset is needed per scan.
      Tiles are handed out by an atomic ticket so that a tile's predecessors
      are always running or finished (forward progress of the look-back).
  k_gather (K4)  exclusive scan of record widths (same look-back) + gather of
          suf[lb..lb+width) into the ragged positions array.
  k_llvdir        builds the .llv bucket directory at upload time.
*/
#include "smax_kernels.cuh"

namespace smax {

// ------------------------------------------------------------------ utils
__device__ __forceinline__ uint64_t ld_relaxed(const uint64_t *p)
{
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void st_relaxed(uint64_t *p, uint64_t v)
{
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
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

// per-byte "byte >= mb" for the four bytes of w; result has bit 7 of each
// qualifying byte set.  kadd / himode are derived from mb on the host side of
// the kernel (see filter_consts).
__device__ __forceinline__ uint32_t swar_ge(uint32_t w, uint32_t kadd, bool himode)
{
  const uint32_t t = (w & 0x7f7f7f7fu) + kadd;
  return (himode ? (t & w) : (t | w)) & 0x80808080u;
}

// CTA-wide decoupled look-back.  Every thread of the CTA calls it with the
// tile's aggregate; returns the exclusive prefix over tiles [0, tile).
// Window = blockDim.x predecessors per round, so even when all resident tiles
// finish at the same moment the chain resolves in ntiles/256 rounds.
__device__ uint64_t lookback_exclusive(uint64_t *status, uint32_t tile, uint64_t agg,
                                       uint32_t epoch)
{
  __shared__ uint64_t s_wsum[kThreads / 32];
  __shared__ int s_wflag[kThreads / 32];
  __shared__ int s_wcut[kThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tile == 0)
  {
    if (tid == 0)
      st_relaxed(&status[0], pack_status(epoch, kStatePrefix, agg));
    return 0;
  }
  if (tid == 0)
    st_relaxed(&status[tile], pack_status(epoch, kStateAggregate, agg));
  uint64_t excl = 0;
  int64_t hi = tile;                 // window = tiles [hi - 256, hi), nearest first
  for (;;)
  {
    const int64_t idx = hi - 1 - tid;
    uint64_t st = kStatePrefix, val = 0;     // virtual tiles < 0: prefix 0
    if (idx >= 0)
    {
      const uint64_t w = ld_relaxed(&status[idx]);
      if ((uint32_t) (w >> (kValueBits + 2)) == epoch)
      {
        st = (w >> kValueBits) & 3;
        val = w & kValueMask;
      } else
        st = kStateInvalid;
    }
    const unsigned inv = __ballot_sync(0xffffffffu, st == kStateInvalid);
    const unsigned pm = __ballot_sync(0xffffffffu, st == kStatePrefix);
    const int first_inv = inv ? __ffs(inv) - 1 : 32;
    const int first_p = pm ? __ffs(pm) - 1 : 32;
    int flag, cut;
    if (first_p < first_inv) { flag = 1; cut = first_p + 1; }   // reached a prefix
    else if (first_inv < 32) { flag = 2; cut = first_inv; }     // not published yet
    else { flag = 0; cut = 32; }                                // 32 aggregates
    uint64_t x = lane < cut ? val : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0)
    {
      s_wsum[warp] = x;
      s_wflag[warp] = flag;
      s_wcut[warp] = cut;
    }
    __syncthreads();
    uint64_t acc = 0;
    int outcome = 0, consumed = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++)
    {
      if (outcome == 0)
      {
        acc += s_wsum[w];
        consumed += s_wcut[w];
        outcome = s_wflag[w];
      }
    }
    __syncthreads();
    excl += acc;
    if (outcome == 1)
      break;
    hi -= consumed;
    if (outcome == 2 && consumed == 0)
      __nanosleep(64);
  }
  if (tid == 0)
    st_relaxed(&status[tile], pack_status(epoch, kStatePrefix, excl + agg));
  return excl;
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

// resolved lcp value at an arbitrary index (slow, fully general)
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

__device__ __noinline__ uint32_t bwt_at_left(const ScanParams &P, uint64_t q, bool &bad)
{
  const TableView *tv = view_for(P, q);
  if (tv == nullptr) { bad = true; return 0; }
  return tv->bwt[q - tv->a_lo];
}

struct Survivor
{
  uint64_t v, lb, width;
};

// Examine lcp index e (byte b >= mb).  Returns true iff [.., e] is a
// supermaximal repeat and fills sv.  All the rare work lives here.
template <bool STATS>
__device__ __forceinline__ bool examine(const ScanParams &P, uint64_t e, uint32_t b,
                                        Survivor &sv, bool &bad, uint64_t *stat)
{
  const TableView &own = P.own;
  const uint8_t *lcp = own.lcp;
  const uint64_t a_lo = own.a_lo;
  uint64_t v = b, k = 0;

  if (e < P.g_lo || e >= P.g_hi)
    return false;
  const uint32_t nb = lcp[e + 1 - a_lo];
  if (b < 255)
  {
    if (nb >= b)                      // no fall (nb == 255 is larger)
      return false;
  } else
  {
    if (!llv_find(own, e, k)) { bad = true; return false; }
    v = own.llv[k].value;
    if (STATS) stat[2]++;
    if (v < P.minlength)
      return false;
    if (nb == 255)
    {
      if (k + 1 >= own.nllv || own.llv[k + 1].position != e + 1) { bad = true; return false; }
      if (STATS) stat[2]++;
      if (own.llv[k + 1].value >= v)
        return false;
    }
  }
  // ---- walk left over the run of value v
  uint64_t s = e;
  if (b < 255)
  {
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
      {
        const TableView *tv = view_for(P, q);
        if (tv == nullptr) { bad = true; return false; }
        pb = tv->lcp[q - tv->a_lo];
      }
      if (pb == b) { s = q; continue; }
      if (pb > b)
        return false;
      break;
    }
  } else
  {
    uint64_t kk = k;
    for (;;)
    {
      const uint64_t q = s - 1;
      uint64_t pv;
      if (q >= a_lo)
      {
        if (lcp[q - a_lo] < 255)
          break;
        if (kk == 0 || own.llv[kk - 1].position != q) { bad = true; return false; }
        pv = own.llv[--kk].value;
        if (STATS) stat[2]++;
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
  }
  const uint64_t lb = s - 1, width = e - lb + 1;
  if (STATS) { stat[0]++; stat[1] += width; }
  // ---- left characters pairwise distinct?
  uint64_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
  const bool gt_policy = (P.policy == SMAX_POLICY_GT);
  for (uint64_t q = lb; q <= e; q++)
  {
    const uint32_t c = (q >= a_lo) ? (uint32_t) own.bwt[q - a_lo] : bwt_at_left(P, q, bad);
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
  if (bad)
    return false;
  if (STATS) stat[3] += width;
  sv.v = v; sv.lb = lb; sv.width = width;
  return true;
}

// shared memory of the scan kernel
struct ScanSmem
{
  uint64_t stage[kStageCap * 3];
  uint32_t bitmap[kTileWords];
  uint16_t wprefix[kTileWords];
  uint16_t stage_off[kStageCap];
  uint32_t warp_tot[kThreads / 32];
  uint32_t count;
  uint32_t tile;
};

__device__ __forceinline__ uint32_t rank_in_tile(const ScanSmem &sm, uint32_t o)
{
  return sm.wprefix[o >> 5] + __popc(sm.bitmap[o >> 5] & ((1u << (o & 31)) - 1u));
}

// Slow path for one 16-byte chunk that passed the filter.  DIRECT == false:
// mark + stage survivors.  DIRECT == true (tile had more survivors than the
// stage holds): ranks are known, write the records straight to the output.
template <bool STATS, bool DIRECT>
__device__ __noinline__ void process_chunk(const ScanParams &P, ScanSmem &sm, uint4 w,
                                           uint32_t h0, uint32_t h1, uint32_t h2, uint32_t h3,
                                           uint64_t cbase, uint32_t coff, uint64_t excl)
{
  const uint32_t words[4] = {w.x, w.y, w.z, w.w};
  const uint32_t masks[4] = {h0, h1, h2, h3};
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
      const int j = q * 4 + (bit >> 3);
      const uint32_t b = (words[q] >> (bit - 7)) & 0xffu;
      Survivor sv;
      if (!examine<STATS>(P, cbase + j, b, sv, bad, stat))
        continue;
      const uint32_t o = coff + j;           // offset of the plateau end in the tile
      if (!DIRECT)
      {
        atomicOr(&sm.bitmap[o >> 5], 1u << (o & 31));
        const uint32_t slot = atomicAdd(&sm.count, 1u);
        if (slot < (uint32_t) kStageCap)
        {
          sm.stage[slot * 3 + 0] = sv.v;
          sm.stage[slot * 3 + 1] = sv.lb;
          sm.stage[slot * 3 + 2] = sv.width;
          sm.stage_off[slot] = (uint16_t) o;
        }
      } else
      {
        const uint64_t dst = excl + rank_in_tile(sm, o);
        if (dst < P.rec_capacity)
        {
          smax_record r;
          r.len = sv.v; r.lb = sv.lb; r.width = sv.width;
          P.recs[dst] = r;
        } else
          P.result[kResOverflow] = 1;
      }
    }
  }
  if (bad)
    P.result[kResError] = 1;
  if (STATS && !DIRECT)
  {
    if (stat[0]) atomicAdd((unsigned long long *) &P.result[kResStatCand], (unsigned long long) stat[0]);
    if (stat[1]) atomicAdd((unsigned long long *) &P.result[kResStatCandWidth], (unsigned long long) stat[1]);
    if (stat[2]) atomicAdd((unsigned long long *) &P.result[kResStatLlv], (unsigned long long) stat[2]);
    if (stat[3]) atomicAdd((unsigned long long *) &P.result[kResStatSurvWidth], (unsigned long long) stat[3]);
  }
}

// ------------------------------------------------------------ scan kernel
template <bool STATS>
__global__ void __launch_bounds__(kThreads, 4)
k_scan(const __grid_constant__ ScanParams P)
{
  __shared__ ScanSmem sm;
  const int tid = threadIdx.x;
  const uint8_t *lcp = P.own.lcp;
  const uint64_t a_lo = P.own.a_lo;
  const uint64_t base_off = P.g_lo - a_lo;                       // multiple of 16
  const uint64_t len16 = (P.own.a_hi - a_lo + 15) & ~15ull;      // loadable bytes
  const bool himode = P.mb > 128;
  const uint32_t kadd = (himode ? (0x100u - P.mb) : (0x80u - P.mb)) * 0x01010101u;
  bool dirty = true;

  for (;;)
  {
    __syncthreads();
    if (tid == 0)
    {
      sm.tile = atomicAdd(&P.ctrl[0], 1u);
      sm.count = 0;
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

    // ---- K1: all loads first, then filter
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
      if (h0 | h1 | h2 | h3)
      {
        const uint32_t coff = (uint32_t) (c * kThreads + tid) * kChunk;
        process_chunk<STATS, false>(P, sm, w[c], h0, h1, h2, h3, a_lo + toff + coff, coff, 0);
      }
    }
    __syncthreads();

    // ---- K3: ranks from the bitmap, look-back, ordered write
    const uint32_t count = sm.count;
    dirty = count != 0;
    if (count)
    {
      const uint32_t c0 = __popc(sm.bitmap[2 * tid]), c1 = __popc(sm.bitmap[2 * tid + 1]);
      uint32_t x = c0 + c1;
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
        if (k < warp) wbase += sm.warp_tot[k];
      const uint32_t ex = wbase + x - (c0 + c1);
      sm.wprefix[2 * tid] = (uint16_t) ex;
      sm.wprefix[2 * tid + 1] = (uint16_t) (ex + c0);
    }
    const uint64_t excl = lookback_exclusive(P.status, tile, count, P.epoch);
    if (tile == P.ntiles - 1 && tid == 0)
      P.result[kResCount] = excl + count;
    if (count == 0)
      continue;
    __syncthreads();                     // wprefix visible
    if (count <= (uint32_t) kStageCap)
    {
      for (uint32_t slot = tid; slot < count; slot += kThreads)
      {
        const uint64_t dst = excl + rank_in_tile(sm, sm.stage_off[slot]);
        if (dst < P.rec_capacity)
        {
          smax_record r;
          r.len = sm.stage[slot * 3 + 0];
          r.lb = sm.stage[slot * 3 + 1];
          r.width = sm.stage[slot * 3 + 2];
          P.recs[dst] = r;
        } else
          P.result[kResOverflow] = 1;
      }
    } else
    {
      // more survivors than the stage holds: redo the tile, writing directly
#pragma unroll 1
      for (int c = 0; c < kItems; c++)
      {
        const uint32_t coff = (uint32_t) (c * kThreads + tid) * kChunk;
        const uint64_t off = toff + coff;
        if (off >= len16)
          continue;
        const uint4 ww = *reinterpret_cast<const uint4 *>(lcp + off);
        const uint32_t h0 = swar_ge(ww.x, kadd, himode), h1 = swar_ge(ww.y, kadd, himode),
                       h2 = swar_ge(ww.z, kadd, himode), h3 = swar_ge(ww.w, kadd, himode);
        if (h0 | h1 | h2 | h3)
          process_chunk<false, true>(P, sm, ww, h0, h1, h2, h3, a_lo + off, coff, excl);
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
        P.result[kResCount] = 0;
      for (int k = 0; k < kResSlots; k++)   // result blocks ping-pong: no memset per scan
        P.result_next[k] = 0;
    }
  }
}

// ---------------------------------------------------------- gather kernel
__device__ __forceinline__ uint64_t suf_at(const ScanParams &P, uint64_t i, bool &bad)
{
  const TableView *tv = view_for(P, i);
  if (tv == nullptr || tv->suf == nullptr) { bad = true; return 0; }
  const uint64_t o = i - tv->a_lo;
  return P.sufbytes == 8 ? reinterpret_cast<const uint64_t *>(tv->suf)[o]
                         : (uint64_t) reinterpret_cast<const uint32_t *>(tv->suf)[o];
}

__global__ void __launch_bounds__(kThreads, 4)
k_gather(const __grid_constant__ ScanParams P)
{
  __shared__ uint64_t s_warp_tot[kThreads / 32];
  __shared__ uint32_t s_tile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool overflow = P.result[kResOverflow] != 0;
  const uint64_t count = overflow ? 0 : P.result[kResCount];
  const uint64_t ntiles = (count + kGatherTile - 1) / kGatherTile;
  bool bad = false;

  for (;;)
  {
    __syncthreads();
    if (tid == 0)
      s_tile = atomicAdd(&P.ctrl[2], 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= ntiles)
      break;
    const uint64_t r0 = (uint64_t) tile * kGatherTile + (uint64_t) tid * kGatherItems;
    uint64_t wdt[kGatherItems], lbs[kGatherItems], tsum = 0;
#pragma unroll
    for (int j = 0; j < kGatherItems; j++)
    {
      wdt[j] = 0; lbs[j] = 0;
      if (r0 + j < count)
      {
        wdt[j] = P.recs[r0 + j].width;
        lbs[j] = P.recs[r0 + j].lb;
      }
      tsum += wdt[j];
    }
    uint64_t x = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const uint64_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31)
      s_warp_tot[warp] = x;
    __syncthreads();
    uint64_t wbase = 0, total = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; k++)
    {
      if (k < warp) wbase += s_warp_tot[k];
      total += s_warp_tot[k];
    }
    const uint64_t excl = lookback_exclusive(P.status2, tile, total, P.epoch);
    if (tile == ntiles - 1 && tid == 0)
      P.result[kResPositions] = excl + total;
    uint64_t o = excl + wbase + x - tsum;
#pragma unroll
    for (int j = 0; j < kGatherItems; j++)
    {
      for (uint64_t k = 0; k < wdt[j]; k++)
      {
        if (o + k < P.pos_capacity)
          P.positions[o + k] = suf_at(P, lbs[j] + k, bad);
        else
          P.result[kResOverflow] = 1;
      }
      o += wdt[j];
    }
  }
  if (bad)
    P.result[kResError] = 1;
  if (tid == 0)
  {
    __threadfence();
    const uint32_t done = atomicAdd(&P.ctrl[3], 1u);
    if (done == gridDim.x - 1)
    {
      P.ctrl[2] = 0;
      P.ctrl[3] = 0;
      if (ntiles == 0)
        P.result[kResPositions] = 0;
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

cudaError_t launch_gather(const ScanParams &p, int grid, cudaStream_t st)
{
  k_gather<<<grid, kThreads, 0, st>>>(p);
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

int gather_blocks_per_sm()
{
  int n = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_gather, kThreads, 0);
  return (e == cudaSuccess && n > 0) ? n : 1;
}

}  // namespace smax
