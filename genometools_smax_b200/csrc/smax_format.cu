/*
  smax_format.cu -- the emit half of the path on the device (SURVEY.md 8f,
  ranks 1 and 2): the records and gathered positions a scan left in HBM are
  rendered as the tool's text there, so that the host only writes bytes.

    separator table   { suf[i] - 1 : bwt[i] == 255 } in ascending order
                      (what gt_encseq_seqnum / gt_encseq_seqstartpos answer
                      from .ssp, /root/reference/src/core/encseq.c:3815-3900;
                      bwt semantics /root/reference/src/match/sfx-run.c:188-207)
                      by a bitmap over text positions + rank: a counting sort
    k_fmt_*           size of every header / position item, exclusive scans,
                      then every item writes its decimals at its byte offset;
                      lines stay in suffix-array order
                      (/root/reference/src/match/esa-bottomup.c:160-170)

  Text conventions: smax_dec.h, csrc/smax_emit.c (the host emitter is the
  byte-exact reference of these kernels in tests/).  sm_100a.
*/
#include "smax_kernels.cuh"
#include "smax_dec.h"

namespace smax {

namespace {

constexpr int kFmtThreads = 256;
constexpr int kFmtItems = 8;
constexpr int kFmtBlockItems = kFmtThreads * kFmtItems;

__device__ __forceinline__ uint64_t warp_sum(uint64_t v)
{
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// inclusive scan of one value per thread over the CTA; *total = sum of all
__device__ __forceinline__ uint64_t block_scan_incl(uint64_t v, uint64_t *total)
{
  __shared__ uint64_t wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const uint64_t o = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += o;
  }
  __syncthreads();   // wsum may still be read by a previous call
  if (lane == 31) wsum[warp] = v;
  __syncthreads();
  uint64_t before = 0, all = 0;
  for (int w = 0; w < nwarps; w++)
  {
    const uint64_t s = wsum[w];
    if (w < warp) before += s;
    all += s;
  }
  *total = all;
  return v + before;
}

// ---- item sizes -------------------------------------------------------
struct HeaderSize      // bytes of a record's own text, its newline included
{
  const smax_record *recs;
  int format;
  __device__ __forceinline__ uint64_t operator()(uint64_t r) const
  {
    const smax_record rec = recs[r];
    if (format == SMAX_FORMAT_ITV)
      return smax_dec_digits(rec.len) + 1 + smax_dec_digits(rec.lb) + 1 +
             smax_dec_digits(rec.lb + rec.width - 1) + 1;
    return smax_dec_digits(rec.len) + 1 + smax_dec_digits(rec.width) + 1;
  }
};

struct RecordWidth
{
  const smax_record *recs;
  __device__ __forceinline__ uint64_t operator()(uint64_t r) const { return recs[r].width; }
};

// position -> (sequence number, offset in the sequence): number of separators
// left of pos, pos - start of that sequence (csrc/smax_index.c, same answer)
__device__ __forceinline__ void seq_rel(const uint64_t *seps, uint64_t nseps, uint64_t pos,
                                        uint64_t *seqnum, uint64_t *rel)
{
  uint64_t lo = 0, hi = nseps;
  while (lo < hi)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (__ldg(seps + mid) < pos) lo = mid + 1; else hi = mid;
  }
  *seqnum = lo;
  *rel = lo == 0 ? pos : pos - (__ldg(seps + lo - 1) + 1);
}

struct PositionSize    // " <pos>" or " <seqnum> <relpos>"
{
  const uint64_t *pos;
  const uint64_t *seps;
  uint64_t nseps;
  int relative;
  __device__ __forceinline__ uint64_t operator()(uint64_t j) const
  {
    const uint64_t p = pos[j];
    if (!relative)
      return 1 + smax_dec_digits(p);
    uint64_t s, r;
    seq_rel(seps, nseps, p, &s, &r);
    return 1 + smax_dec_digits(s) + 1 + smax_dec_digits(r);
  }
};

struct WordPopc        // separators per 64 text positions
{
  const uint64_t *bitmap;
  __device__ __forceinline__ uint64_t operator()(uint64_t w) const
  {
    return (uint64_t) __popcll(bitmap[w]);
  }
};

// ---- exclusive scan of f(0..n) into out[0..n], out[n] = total ------------
template <class F>
__global__ void __launch_bounds__(kFmtThreads) k_fmt_reduce(F f, uint64_t n, uint64_t *sums)
{
  __shared__ uint64_t part[kFmtThreads / 32];
  const uint64_t base = (uint64_t) blockIdx.x * kFmtBlockItems;
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < kFmtItems; k++)
  {
    const uint64_t i = base + (uint64_t) k * kFmtThreads + threadIdx.x;
    if (i < n) s += f(i);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    uint64_t t = 0;
    for (int w = 0; w < kFmtThreads / 32; w++) t += part[w];
    sums[blockIdx.x] = t;
  }
}

// one CTA: sums[0..nb) -> exclusive prefixes, sums[nb] = total
__global__ void __launch_bounds__(1024) k_fmt_scan_sums(uint64_t *sums, uint64_t nb)
{
  uint64_t carry = 0;
  for (uint64_t base = 0; base < nb; base += blockDim.x)
  {
    const uint64_t i = base + threadIdx.x;
    const uint64_t v = i < nb ? sums[i] : 0;
    uint64_t total;
    const uint64_t incl = block_scan_incl(v, &total);
    if (i < nb) sums[i] = carry + incl - v;
    carry += total;
  }
  if (threadIdx.x == 0) sums[nb] = carry;
}

template <class F>
__global__ void __launch_bounds__(kFmtThreads) k_fmt_apply(F f, uint64_t n, const uint64_t *sums,
                                                          uint64_t nb, uint64_t *out)
{
  const uint64_t first = (uint64_t) blockIdx.x * kFmtBlockItems + (uint64_t) threadIdx.x * kFmtItems;
  uint64_t v[kFmtItems], mine = 0;
#pragma unroll
  for (int k = 0; k < kFmtItems; k++)
  {
    v[k] = first + k < n ? f(first + k) : 0;
    mine += v[k];
  }
  uint64_t total;
  uint64_t run = sums[blockIdx.x] + block_scan_incl(mine, &total) - mine;
#pragma unroll
  for (int k = 0; k < kFmtItems; k++)
  {
    if (first + k < n) out[first + k] = run;
    run += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[nb];
}

template <class F>
cudaError_t exclusive_scan(F f, uint64_t n, uint64_t *sums, uint64_t *out, cudaStream_t st)
{
  const uint64_t nb = (n + kFmtBlockItems - 1) / kFmtBlockItems;
  if (nb > 0)
    k_fmt_reduce<<<(unsigned) nb, kFmtThreads, 0, st>>>(f, n, sums);
  k_fmt_scan_sums<<<1, 1024, 0, st>>>(sums, nb);
  if (nb > 0)
    k_fmt_apply<<<(unsigned) nb, kFmtThreads, 0, st>>>(f, n, sums, nb, out);
  else
    k_fmt_apply<<<1, kFmtThreads, 0, st>>>(f, n, sums, nb, out);
  return cudaGetLastError();
}

// ---- the three scans of a format call, fused: 3 launches instead of 9 ------
// blocks [0, nbr) work on records (two values per record: own text bytes, width),
// blocks [nbr, nbr + nbp) on positions (item bytes); sums: [hdr | width | pos] block sums,
// each part followed by its total.
struct FmtScanPlan
{
  HeaderSize hdr;
  RecordWidth wid;
  PositionSize psz;
  uint64_t nrecs, npos, nbr, nbp;
  uint64_t *sums;                 // (nbr + 1) + (nbr + 1) + (nbp + 1) words
  uint64_t *hoff, *pfirst, *poff;
  __device__ __forceinline__ uint64_t *sums_h() const { return sums; }
  __device__ __forceinline__ uint64_t *sums_w() const { return sums + nbr + 1; }
  __device__ __forceinline__ uint64_t *sums_p() const { return sums + 2 * (nbr + 1); }
};

__device__ __forceinline__ uint64_t block_sum(uint64_t s, uint64_t *part)
{
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  uint64_t t = 0;
  for (int w = 0; w < kFmtThreads / 32; w++) t += part[w];
  return t;
}

__global__ void __launch_bounds__(kFmtThreads) k_fmt_reduce_all(FmtScanPlan pl)
{
  __shared__ uint64_t part[kFmtThreads / 32];
  if (blockIdx.x < pl.nbr)
  {
    const uint64_t base = (uint64_t) blockIdx.x * kFmtBlockItems;
    uint64_t sh = 0, sw = 0;
#pragma unroll
    for (int k = 0; k < kFmtItems; k++)
    {
      const uint64_t i = base + (uint64_t) k * kFmtThreads + threadIdx.x;
      if (i < pl.nrecs) { sh += pl.hdr(i); sw += pl.wid(i); }
    }
    sh = block_sum(sh, part);
    sw = block_sum(sw, part);
    if (threadIdx.x == 0) { pl.sums_h()[blockIdx.x] = sh; pl.sums_w()[blockIdx.x] = sw; }
  } else
  {
    const uint64_t b = blockIdx.x - pl.nbr, base = b * kFmtBlockItems;
    uint64_t sp = 0;
#pragma unroll
    for (int k = 0; k < kFmtItems; k++)
    {
      const uint64_t i = base + (uint64_t) k * kFmtThreads + threadIdx.x;
      if (i < pl.npos) sp += pl.psz(i);
    }
    sp = block_sum(sp, part);
    if (threadIdx.x == 0) pl.sums_p()[b] = sp;
  }
}

// one CTA: the three block-sum arrays -> exclusive prefixes, each followed by its total
__global__ void __launch_bounds__(1024) k_fmt_scan_sums_all(FmtScanPlan pl)
{
  for (int part = 0; part < 3; part++)
  {
    uint64_t *sums = part == 0 ? pl.sums_h() : part == 1 ? pl.sums_w() : pl.sums_p();
    const uint64_t nb = part == 2 ? pl.nbp : pl.nbr;
    uint64_t carry = 0;
    for (uint64_t base = 0; base < nb; base += blockDim.x)
    {
      const uint64_t i = base + threadIdx.x;
      const uint64_t v = i < nb ? sums[i] : 0;
      uint64_t total;
      const uint64_t incl = block_scan_incl(v, &total);
      if (i < nb) sums[i] = carry + incl - v;
      carry += total;
    }
    if (threadIdx.x == 0) sums[nb] = carry;
    __syncthreads();
  }
}

template <class F>
__device__ __forceinline__ void apply_block(const F &f, uint64_t n, uint64_t block, uint64_t offset,
                                            uint64_t *out)
{
  const uint64_t first = block * kFmtBlockItems + (uint64_t) threadIdx.x * kFmtItems;
  uint64_t v[kFmtItems], mine = 0;
#pragma unroll
  for (int k = 0; k < kFmtItems; k++)
  {
    v[k] = first + k < n ? f(first + k) : 0;
    mine += v[k];
  }
  uint64_t total;
  uint64_t run = offset + block_scan_incl(mine, &total) - mine;
#pragma unroll
  for (int k = 0; k < kFmtItems; k++)
  {
    if (first + k < n) out[first + k] = run;
    run += v[k];
  }
}

__global__ void __launch_bounds__(kFmtThreads) k_fmt_apply_all(FmtScanPlan pl)
{
  if (blockIdx.x < pl.nbr)
  {
    apply_block(pl.hdr, pl.nrecs, blockIdx.x, pl.sums_h()[blockIdx.x], pl.hoff);
    apply_block(pl.wid, pl.nrecs, blockIdx.x, pl.sums_w()[blockIdx.x], pl.pfirst);
  } else
  {
    const uint64_t b = blockIdx.x - pl.nbr;
    apply_block(pl.psz, pl.npos, b, pl.sums_p()[b], pl.poff);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
  {
    pl.hoff[pl.nrecs] = pl.sums_h()[pl.nbr];
    pl.pfirst[pl.nrecs] = pl.sums_w()[pl.nbr];
    pl.poff[pl.npos] = pl.sums_p()[pl.nbp];
  }
}

// ---- writers ------------------------------------------------------------
__device__ __forceinline__ char *put(char *p, uint64_t v)
{
  const unsigned d = smax_dec_digits(v);
  smax_dec_write(p, v, d);
  return p + d;
}

// header of every record and its newline.  hoff[r]: bytes of the own text of
// records < r; poff[j]: bytes of position items < j; pfirst[r]: index of the
// record's first position
__device__ __forceinline__ void write_record(
    uint64_t r, const smax_record *recs, uint64_t nrecs, int format, const uint64_t *hoff,
    const uint64_t *pfirst, const uint64_t *poff, char *text)
{
  if (r >= nrecs) return;
  const smax_record rec = recs[r];
  if (format == SMAX_FORMAT_ITV)
  {
    char *p = text + hoff[r];
    p = put(p, rec.len); *p++ = ' ';
    p = put(p, rec.lb); *p++ = ' ';
    p = put(p, rec.lb + rec.width - 1); *p = '\n';
    return;
  }
  char *p = text + hoff[r] + poff[pfirst[r]];
  p = put(p, rec.len); *p++ = ' ';
  put(p, rec.width);
  text[hoff[r + 1] + poff[pfirst[r + 1]] - 1] = '\n';
}

__device__ __forceinline__ void write_position(
    uint64_t j, const uint64_t *pos, uint64_t npos, uint64_t nrecs, const uint64_t *hoff,
    const uint64_t *pfirst, const uint64_t *poff, const uint64_t *seps, uint64_t nseps,
    int relative, char *text)
{
  if (j >= npos) return;
  // the record that holds position j: last r with pfirst[r] <= j
  uint64_t lo = 0, hi = nrecs;
  while (hi - lo > 1)
  {
    const uint64_t mid = (lo + hi) >> 1;
    if (__ldg(pfirst + mid) <= j) lo = mid; else hi = mid;
  }
  // own texts of records <= lo without the newline of lo, position items < j
  char *p = text + hoff[lo + 1] - 1 + poff[j];
  const uint64_t v = pos[j];
  *p++ = ' ';
  if (!relative)
  {
    put(p, v);
    return;
  }
  uint64_t s, rel;
  seq_rel(seps, nseps, v, &s, &rel);
  p = put(p, s); *p++ = ' ';
  put(p, rel);
}

// blocks [0, rec_blocks): one thread per record; the rest: one thread per position
__global__ void __launch_bounds__(kFmtThreads) k_fmt_write(
    const smax_record *recs, uint64_t nrecs, int format, const uint64_t *pos, uint64_t npos,
    const uint64_t *hoff, const uint64_t *pfirst, const uint64_t *poff, const uint64_t *seps,
    uint64_t nseps, int relative, char *text, unsigned rec_blocks)
{
  if (blockIdx.x < rec_blocks)
    write_record((uint64_t) blockIdx.x * kFmtThreads + threadIdx.x, recs, nrecs, format, hoff, pfirst,
                 poff, text);
  else
    write_position((uint64_t) (blockIdx.x - rec_blocks) * kFmtThreads + threadIdx.x, pos, npos, nrecs,
                   hoff, pfirst, poff, seps, nseps, relative, text);
}

// ---- separator table ----------------------------------------------------
__global__ void __launch_bounds__(kFmtThreads) k_sep_mark(
    const uint8_t *bwt, const void *suf, int sufbytes, uint64_t len,
    unsigned long long *bitmap, uint64_t nbits, unsigned long long *bad)
{
  const uint64_t stride = (uint64_t) gridDim.x * kFmtThreads * 4;
  const bool aligned = (reinterpret_cast<uintptr_t>(bwt) & 3) == 0;
  for (uint64_t i = ((uint64_t) blockIdx.x * kFmtThreads + threadIdx.x) * 4; i < len; i += stride)
  {
    uint32_t w;
    if (aligned && i + 4 <= len)
      w = *reinterpret_cast<const uint32_t *>(bwt + i);
    else
    {
      w = 0;
      for (int k = 0; k < 4 && i + k < len; k++)
        w |= (uint32_t) bwt[i + k] << (8 * k);
    }
    const uint32_t inv = ~w;   // a byte 255 is a zero byte of the complement
    if (((inv - 0x01010101u) & ~inv & 0x80808080u) == 0)
      continue;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
      if (((w >> (8 * k)) & 0xffu) != 0xffu || i + k >= len) continue;
      const uint64_t s = sufbytes == 8 ? reinterpret_cast<const uint64_t *>(suf)[i + k]
                                       : reinterpret_cast<const uint32_t *>(suf)[i + k];
      if (s == 0 || s - 1 >= nbits) { atomicAdd(bad, 1ull); continue; }
      atomicOr(bitmap + ((s - 1) >> 6), 1ull << ((s - 1) & 63));
    }
  }
}

__global__ void __launch_bounds__(kFmtThreads) k_sep_fill(
    const uint64_t *bitmap, uint64_t nwords, const uint64_t *rank, uint64_t *seps)
{
  const uint64_t w = (uint64_t) blockIdx.x * kFmtThreads + threadIdx.x;
  if (w >= nwords) return;
  uint64_t bits = bitmap[w], k = rank[w];
  while (bits != 0)
  {
    const int b = __ffsll((long long) bits) - 1;
    seps[k++] = w * 64 + (uint64_t) b;
    bits &= bits - 1;
  }
}

}  // namespace

// scratch words the scans of a format call need for n items
uint64_t format_sums_words(uint64_t n)
{
  return 3 * ((n + kFmtBlockItems - 1) / kFmtBlockItems + 2);
}

cudaError_t launch_format_measure(const FormatJob &j, cudaStream_t st)
{
  if (j.format == SMAX_FORMAT_ITV)
    return exclusive_scan(HeaderSize{j.recs, j.format}, j.nrecs, j.sums, j.hoff, st);
  FmtScanPlan pl = { HeaderSize{j.recs, j.format}, RecordWidth{j.recs},
                     PositionSize{j.pos, j.seps, j.nseps, j.relative},
                     j.nrecs, j.npos, (j.nrecs + kFmtBlockItems - 1) / kFmtBlockItems,
                     (j.npos + kFmtBlockItems - 1) / kFmtBlockItems, j.sums, j.hoff, j.pfirst,
                     j.poff };
  const unsigned blocks = (unsigned) (pl.nbr + pl.nbp);
  if (blocks > 0)
    k_fmt_reduce_all<<<blocks, kFmtThreads, 0, st>>>(pl);
  k_fmt_scan_sums_all<<<1, 1024, 0, st>>>(pl);
  k_fmt_apply_all<<<blocks > 0 ? blocks : 1, kFmtThreads, 0, st>>>(pl);
  return cudaGetLastError();
}

cudaError_t launch_format_write(const FormatJob &j, cudaStream_t st)
{
  const uint64_t npos = j.format == SMAX_FORMAT_ITV ? 0 : j.npos;
  const uint64_t br = (j.nrecs + kFmtThreads - 1) / kFmtThreads, bp = (npos + kFmtThreads - 1) / kFmtThreads;
  if (br + bp > 0)
    k_fmt_write<<<(unsigned) (br + bp), kFmtThreads, 0, st>>>(j.recs, j.nrecs, j.format, j.pos, npos, j.hoff,
                                                           j.pfirst, j.poff, j.seps, j.nseps, j.relative,
                                                           j.text, (unsigned) br);
  return cudaGetLastError();
}

cudaError_t launch_sep_mark(const uint8_t *bwt, const void *suf, int sufbytes, uint64_t len,
                            uint64_t *bitmap, uint64_t nbits, uint64_t *bad, int sm_count,
                            cudaStream_t st)
{
  const uint64_t want = (len / 4 + kFmtThreads - 1) / kFmtThreads + 1;
  const unsigned grid = (unsigned) (want < (uint64_t) sm_count * 8 ? want : (uint64_t) sm_count * 8);
  k_sep_mark<<<grid, kFmtThreads, 0, st>>>(bwt, suf, sufbytes, len,
                                          reinterpret_cast<unsigned long long *>(bitmap), nbits,
                                          reinterpret_cast<unsigned long long *>(bad));
  return cudaGetLastError();
}

cudaError_t launch_sep_rank(const uint64_t *bitmap, uint64_t nwords, uint64_t *sums, uint64_t *rank,
                            cudaStream_t st)
{
  return exclusive_scan(WordPopc{bitmap}, nwords, sums, rank, st);
}

cudaError_t launch_sep_fill(const uint64_t *bitmap, uint64_t nwords, const uint64_t *rank,
                            uint64_t *seps, cudaStream_t st)
{
  if (nwords > 0)
    k_sep_fill<<<(unsigned) ((nwords + kFmtThreads - 1) / kFmtThreads), kFmtThreads, 0, st>>>(
        bitmap, nwords, rank, seps);
  return cudaGetLastError();
}

}  // namespace smax
