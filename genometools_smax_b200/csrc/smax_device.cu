/*
  smax_device.cu -- the device-resident half of the path behind the C ABI of
  include/smax.h: table upload through pinned staging buffers, the .llv
  directory, scan/gather launches, record fetch, peer views and CUDA IPC.

  Replaces, on the device, what Suffixarray + Sequentialsuffixarrayreader hold
  on the host in the reference (/root/reference/src/match/sarr-def.h:101-126,
  /root/reference/src/match/esa-seqread.h:27-39).  There is no CPU fallback:
  every entry point fails with a CUDA error message when no B200 is present.
*/
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdarg>
#include <thread>
#include <atomic>
#include <vector>
#include <algorithm>
#include "smax_kernels.cuh"
#include "smax_ring.cuh"
#include "smax_host.h"

using namespace smax;

struct smax_device
{
  int ordinal;
  int sm_count;
  int bps_scan, bps_scan_stats;         // resident CTAs per SM: the unit kernel,
  int bps_ring, bps_ring_stats;         //   the ring kernel
  unsigned long long *d_hist;           // 256 bins: lcp bytes of the shard's own range (upload time)
  uint64_t h_hist[256];
  int last_kernel;                      // kernel of the last scan: 0 ring, 1 units
  bool seen_valid;                      // what an earlier scan of these tables found:
  uint64_t seen_minlength, seen_policy, seen_records;
  cudaStream_t stream;          // uploads / internal work
  // resident shard
  TableView tv;                 // device pointers + coverage
  bool owns_tables;
  size_t cap_lcp, cap_llv, cap_suf, cap_dir, cap_llvv, cap_llvp;   // bytes allocated (owned tables only; dir / compact records always)
  uint64_t g_lo, g_hi, n_total;
  unsigned sufbytes;
  size_t llvdir_entries;
  uint32_t *d_unitdir;          // per unit of [g_lo, g_hi): first .llv record at or behind its start
  size_t cap_unitdir;
  uint32_t *d_unitorder;        // the units, heaviest first (+ the scratch of the sort behind them)
  size_t cap_unitorder;
  uint32_t *d_vals32;           // stripped upload of the .llv records: the values, 4 bytes each,
  size_t cap_vals32;            //   and the scratch of the rebuild behind them (k_llv_*)
  uint32_t *d_rebuild;
  size_t cap_rebuild;
  uint32_t h_rebuild_total;     // 255 bytes the rebuild found (must equal nllv)
  int has_escape;               // some .llv value does not fit the compact record
  int edge_rec0;                // record 0 sits on the first entry of the arrays and the table goes on to the left
  // left neighbours
  TableView left[kMaxLeft];
  int nleft;
  void *ipc_mapped[kMaxLeft * SMAX_IPC_TABLES];
  int n_ipc_mapped;
  // scan scratch
  uint64_t *d_status;
  size_t status_cap;
  uint32_t *d_ctrl;
  uint64_t *d_result;           // 2 * kResSlots (ping-pong)
  UnitMeta *d_meta;             // per unit (a warp's quarter of a tile): aggregate, arena base
  uint64_t *d_blocksum;         // two sets of per-block {repeats, occurrences} sums (ping-pong)
  size_t blocks_cap;            //   blocks per set
  uint64_t unit_scan_no;        //   scans of the unit kernel so far (which set is the clean one)
  size_t unit_cap;
  ArenaEntry *d_arena;          // the scan's survivors before they are put in order (rec_cap entries)
  smax_record *d_recs;
  uint64_t rec_cap;
  uint64_t *d_pos;
  uint64_t pos_cap;
  uint32_t epoch;
  uint32_t scan_no;
  bool stats;
  int debug;
  int grid_limit;
  // last scan
  cudaEvent_t ev0, ev_mid, ev1;
  cudaStream_t last_stream;
  uint64_t last_minlength;
  int last_policy, last_gather, last_launches;
  bool scanned;
  uint64_t h_result[kResSlots];
  bool result_valid;
  // one-sided count exchange
  uint64_t *d_counts;           // own array, `world` slots
  uint64_t *peer_counts[SMAX_MAX_PEERS];
  void *counts_mapped[SMAX_MAX_PEERS];
  int npeers, my_rank;
  uint64_t exchange_tag;
  // device-side text formatting (smax_format.cu)
  uint64_t *d_seps; size_t cap_seps; uint64_t nseps;
  uint64_t *d_fsums, *d_hoff, *d_pfirst, *d_poff;
  size_t cap_fsums, cap_hoff, cap_pfirst, cap_poff;
  char *d_text; size_t cap_text;
  uint64_t text_bytes;
  bool text_valid;
  cudaEvent_t ev_f0, ev_f1;
  // pinned staging ring
  void *pinned[2];
  cudaEvent_t pinned_ev[2];
  size_t pinned_bytes;
};

static int fail(char *err, size_t errlen, const char *fmt, ...)
{
  if (err != NULL && errlen > 0)
  {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, errlen, fmt, ap);
    va_end(ap);
  }
  return -1;
}

#define CU(call)                                                              \
  do {                                                                        \
    cudaError_t e_ = (call);                                                  \
    if (e_ != cudaSuccess)                                                    \
      return fail(err, errlen, "CUDA error: %s (%s) at %s:%d", cudaGetErrorString(e_), \
                  #call, __FILE__, __LINE__);                                 \
  } while (0)

extern "C" int smax_device_count(char *err, size_t errlen)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess)
  {
    fail(err, errlen, "no CUDA device available: %s", cudaGetErrorString(e));
    (void) cudaGetLastError();
    return -1;
  }
  if (n == 0)
    return fail(err, errlen, "no CUDA device available");
  return n;
}

extern "C" int smax_device_create(int ordinal, smax_device **out, char *err, size_t errlen)
{
  int n = smax_device_count(err, errlen);
  if (n < 0)
    return -1;
  if (ordinal < 0 || ordinal >= n)
    return fail(err, errlen, "CUDA device %d does not exist (%d visible)", ordinal, n);
  CU(cudaSetDevice(ordinal));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, ordinal));
  if (prop.major != 10)
    return fail(err, errlen, "device %d is sm_%d%d; libsmax is built for sm_100a (B200) only",
                ordinal, prop.major, prop.minor);
  smax_device *d = (smax_device *) calloc(1, sizeof *d);
  if (d == NULL)
    return fail(err, errlen, "out of memory");
  d->ordinal = ordinal;
  d->sm_count = prop.multiProcessorCount;
  d->bps_scan = scan_blocks_per_sm(false);
  d->bps_scan_stats = scan_blocks_per_sm(true);
  d->bps_ring = smax_ring::scan_blocks_per_sm(false);
  d->bps_ring_stats = smax_ring::scan_blocks_per_sm(true);
  if (d->bps_scan <= 0 || d->bps_scan_stats <= 0 || d->bps_ring <= 0 || d->bps_ring_stats <= 0)
  {
    free(d);
    return fail(err, errlen, "the scan kernel cannot be made resident on device %d", ordinal);
  }
  CU(cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&d->ev0));
  CU(cudaEventCreate(&d->ev_mid));
  CU(cudaEventCreate(&d->ev1));
  CU(cudaEventCreate(&d->ev_f0));
  CU(cudaEventCreate(&d->ev_f1));
  CU(cudaMalloc(&d->d_hist, 256 * sizeof(unsigned long long)));
  CU(cudaMalloc(&d->d_ctrl, 4 * sizeof(uint32_t)));
  CU(cudaMemset(d->d_ctrl, 0, 4 * sizeof(uint32_t)));
  CU(cudaMalloc(&d->d_result, 2 * kResSlots * sizeof(uint64_t)));
  CU(cudaMemset(d->d_result, 0, 2 * kResSlots * sizeof(uint64_t)));
  d->epoch = 0;
  *out = d;
  return 0;
}

static void free_tables(smax_device *d)
{
  if (d->owns_tables)
  {
    cudaFree((void *) d->tv.lcp);
    cudaFree((void *) d->tv.bwt);
    cudaFree((void *) d->tv.llv);
    cudaFree((void *) d->tv.suf);
  }
  cudaFree((void *) d->tv.llvdir);
  cudaFree((void *) d->tv.llvv);
  cudaFree((void *) d->tv.llvp);
  cudaFree(d->d_unitdir);
  d->d_unitdir = NULL;
  d->cap_unitdir = 0;
  cudaFree(d->d_unitorder);
  d->d_unitorder = NULL;
  d->cap_unitorder = 0;
  cudaFree(d->d_vals32);
  d->d_vals32 = NULL;
  d->cap_vals32 = 0;
  cudaFree(d->d_rebuild);
  d->d_rebuild = NULL;
  d->cap_rebuild = 0;
  memset(&d->tv, 0, sizeof d->tv);
  d->owns_tables = false;
  d->cap_lcp = d->cap_llv = d->cap_suf = d->cap_dir = d->cap_llvv = d->cap_llvp = 0;
}

// (re)allocate an owned device table only when it has to grow, so that
// repeated uploads (one per end-to-end step) do not pay cudaMalloc/cudaFree
static cudaError_t ensure_alloc(const void **ptr, size_t *cap, size_t need)
{
  if (*ptr != NULL && *cap >= need)
    return cudaSuccess;
  if (*ptr != NULL)
    cudaFree((void *) *ptr);
  *ptr = NULL;
  *cap = 0;
  void *p = NULL;
  cudaError_t e = cudaMalloc(&p, need);
  if (e == cudaSuccess)
  {
    *ptr = p;
    *cap = need;
  }
  return e;
}

extern "C" void smax_device_destroy(smax_device *d)
{
  if (d == NULL)
    return;
  cudaSetDevice(d->ordinal);
  cudaDeviceSynchronize();
  for (int k = 0; k < d->n_ipc_mapped; k++)
    if (d->ipc_mapped[k] != NULL)
      cudaIpcCloseMemHandle(d->ipc_mapped[k]);
  free_tables(d);
  for (int k = 0; k < SMAX_MAX_PEERS; k++)
    if (d->counts_mapped[k] != NULL)
      cudaIpcCloseMemHandle(d->counts_mapped[k]);
  cudaFree(d->d_counts);
  cudaFree(d->d_meta); cudaFree(d->d_blocksum); cudaFree(d->d_arena);
  cudaFree(d->d_status); cudaFree(d->d_ctrl); cudaFree(d->d_hist);
  cudaFree(d->d_result); cudaFree(d->d_recs); cudaFree(d->d_pos);
  cudaFree(d->d_seps); cudaFree(d->d_fsums); cudaFree(d->d_hoff); cudaFree(d->d_pfirst);
  cudaFree(d->d_poff); cudaFree(d->d_text);
  if (d->ev_f0) cudaEventDestroy(d->ev_f0);
  if (d->ev_f1) cudaEventDestroy(d->ev_f1);
  for (int k = 0; k < 2; k++)
  {
    if (d->pinned[k]) cudaFreeHost(d->pinned[k]);
    if (d->pinned_ev[k]) cudaEventDestroy(d->pinned_ev[k]);
  }
  cudaEventDestroy(d->ev0); cudaEventDestroy(d->ev_mid); cudaEventDestroy(d->ev1);
  cudaStreamDestroy(d->stream);
  free(d);
}

extern "C" void *smax_device_own_stream(smax_device *d)
{
  return d != NULL ? (void *) d->stream : NULL;
}

extern "C" int smax_device_synchronize(smax_device *d)
{
  if (d == NULL)
    return 0;
  if (cudaSetDevice(d->ordinal) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
  {
    (void) cudaGetLastError();
    return -1;
  }
  return 0;
}

// ------------------------------------------------------------- upload
static std::atomic<unsigned long long> g_h2d_total{0};   // bytes copied host -> device by this process

static void parallel_copy(void *dst, const void *src, size_t bytes)
{
  const size_t min_per_thread = 4u << 20;
  unsigned hw = std::thread::hardware_concurrency();
  size_t nthreads = std::min<size_t>(hw ? hw : 4, 16);
  nthreads = std::min(nthreads, std::max<size_t>(1, bytes / min_per_thread));
  if (nthreads <= 1)
  {
    memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> th;
  const size_t per = ((bytes / nthreads) + 4095) & ~(size_t) 4095;
  for (size_t t = 0; t < nthreads; t++)
  {
    const size_t lo = t * per;
    if (lo >= bytes) break;
    const size_t len = std::min(per, bytes - lo);
    th.emplace_back([=] { memcpy((char *) dst + lo, (const char *) src + lo, len); });
  }
  for (auto &t : th) t.join();
}

// host table bytes -> pinned ring -> device, overlapping the host memcpy of
// chunk k+1 with the DMA of chunk k
static int staged_h2d(smax_device *d, void *dst, const void *src, size_t bytes,
                      uint64_t *h2d_bytes, char *err, size_t errlen)
{
  if (bytes == 0)
    return 0;
  cudaPointerAttributes attr;
  const bool src_pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess &&
                          attr.type == cudaMemoryTypeHost;
  (void) cudaGetLastError();
  // the staging ring is sized by the largest table seen so far (1 MiB .. 32 MiB per half):
  // page-locking 64 MiB costs tens of milliseconds, which a small index need not pay
  const size_t want = std::min<size_t>(32u << 20, std::max<size_t>(1u << 20, (bytes + 4095) & ~(size_t) 4095));
  if (!src_pinned && d->pinned_bytes < want)
  {
    for (int k = 0; k < 2; k++)
    {
      if (d->pinned[k] != NULL)
      {
        CU(cudaEventSynchronize(d->pinned_ev[k]));
        CU(cudaFreeHost(d->pinned[k]));
        d->pinned[k] = NULL;
      } else
        CU(cudaEventCreateWithFlags(&d->pinned_ev[k], cudaEventDisableTiming));
      CU(cudaHostAlloc(&d->pinned[k], want, cudaHostAllocDefault));
    }
    d->pinned_bytes = want;
  }
  if (src_pinned)   // caller's buffer is already page-locked: DMA straight from it
  {
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, d->stream));
  } else
  {
    size_t done = 0;
    int k = 0;
    while (done < bytes)
    {
      const size_t len = std::min(d->pinned_bytes, bytes - done);
      CU(cudaEventSynchronize(d->pinned_ev[k]));
      parallel_copy(d->pinned[k], (const char *) src + done, len);
      CU(cudaMemcpyAsync((char *) dst + done, d->pinned[k], len, cudaMemcpyHostToDevice,
                         d->stream));
      CU(cudaEventRecord(d->pinned_ev[k], d->stream));
      done += len;
      k ^= 1;
    }
  }
  if (h2d_bytes) *h2d_bytes += bytes;
  g_h2d_total += bytes;
  return 0;
}

// ---- stripped upload of the .llv records -----------------------------------------------
// process-wide count of the bytes libsmax copied host -> device (measurement: bench.py reports
// the bytes that really crossed the link)
extern "C" uint64_t smax_h2d_bytes_total(void)
{
  return g_h2d_total.load();
}

// values of the records [0, n) as 32-bit words; false when one does not fit
static bool strip_values(uint32_t *dst, const smax_llv *src, size_t n)
{
  const size_t min_per_thread = 1u << 18;
  unsigned hw = std::thread::hardware_concurrency();
  size_t nthreads = std::min<size_t>(hw ? hw : 4, 8);
  nthreads = std::min(nthreads, std::max<size_t>(1, n / min_per_thread));
  std::atomic<bool> fits{true};
  auto work = [&](size_t lo, size_t hi)
  {
    bool ok = true;
    for (size_t i = lo; i < hi; i++)
    {
      const uint64_t v = src[i].value;
      ok &= v < 0xffffffffull;
      dst[i] = (uint32_t) v;
    }
    if (!ok) fits = false;
  };
  if (nthreads <= 1)
  {
    work(0, n);
    return fits;
  }
  std::vector<std::thread> th;
  const size_t per = (n + nthreads - 1) / nthreads;
  for (size_t t = 0; t < nthreads; t++)
  {
    const size_t lo = t * per, hi = std::min(n, lo + per);
    if (lo >= hi) break;
    th.emplace_back(work, lo, hi);
  }
  for (auto &t : th) t.join();
  return fits;
}

// The positions of the .llv records are redundant with the lcp table (the k-th 255 byte is the
// k-th record): only the values cross the link, 4 instead of 16 bytes per record (C2: 22 of
// 90 MB, a fifth of everything the plug-in call uploads), and the device puts the records
// together (k_llv_*).  Returns 1 when the records are on their way, 0 when the caller has to
// upload them whole (a value >= 2^32 - 1, few records, SMAX_LLV_STRIP=0), -1 on error.
static int upload_llv_stripped(smax_device *d, const smax_llv *h_llv, uint64_t nllv, uint64_t len,
                               uint64_t *h2d_bytes, char *err, size_t errlen)
{
  // By default only for a pageable source (mapped index files) and from 65536 records on: those
  // bytes go through the staging ring anyway, and stripping is less work than copying them.  From
  // page-locked tables the records are DMAed as they are: measured on C2 (6.6 ms per call), the
  // CPU pass over the records took as long as the DMA of lcp + bwt it was meant to hide behind
  // (7.2 ms).  SMAX_LLV_STRIP=0 switches it off, =N takes it from N records on for any source (tests: 1).
  const char *env = getenv("SMAX_LLV_STRIP");
  uint64_t from = 1u << 16;
  if (env != NULL)
    from = strtoull(env, NULL, 10);
  else
  {
    cudaPointerAttributes attr;
    const bool src_pinned = cudaPointerGetAttributes(&attr, h_llv) == cudaSuccess &&
                            attr.type == cudaMemoryTypeHost;
    (void) cudaGetLastError();
    if (src_pinned)
      from = 0;
  }
  if (from == 0 || nllv < from)
    return 0;
  const size_t ring = 8u << 20;                  // bytes per half of the staging ring used here
  if (d->pinned_bytes < ring)
  {
    for (int k = 0; k < 2; k++)
    {
      if (d->pinned[k] != NULL)
      {
        CU(cudaEventSynchronize(d->pinned_ev[k]));
        CU(cudaFreeHost(d->pinned[k]));
        d->pinned[k] = NULL;
      } else
        CU(cudaEventCreateWithFlags(&d->pinned_ev[k], cudaEventDisableTiming));
      CU(cudaHostAlloc(&d->pinned[k], ring, cudaHostAllocDefault));
    }
    d->pinned_bytes = ring;
  }
  CU(ensure_alloc((const void **) &d->d_vals32, &d->cap_vals32, (nllv + 4) * sizeof(uint32_t)));
  CU(ensure_alloc((const void **) &d->d_rebuild, &d->cap_rebuild,
                  llv_rebuild_scratch_words(len) * sizeof(uint32_t)));
  const size_t per = d->pinned_bytes / sizeof(uint32_t);
  size_t done = 0;
  int k = 0;
  while (done < nllv)
  {
    const size_t cnt = std::min<size_t>(per, nllv - done);
    CU(cudaEventSynchronize(d->pinned_ev[k]));
    if (!strip_values((uint32_t *) d->pinned[k], h_llv + done, cnt))
    {
      CU(cudaStreamSynchronize(d->stream));      // (what is in flight reads the ring)
      return 0;
    }
    CU(cudaMemcpyAsync(d->d_vals32 + done, d->pinned[k], cnt * sizeof(uint32_t), cudaMemcpyHostToDevice,
                       d->stream));
    CU(cudaEventRecord(d->pinned_ev[k], d->stream));
    done += cnt;
    k ^= 1;
  }
  if (h2d_bytes) *h2d_bytes += nllv * sizeof(uint32_t);
  g_h2d_total += nllv * sizeof(uint32_t);
  CU(launch_llv_rebuild(d->tv.lcp, len, d->tv.a_lo, d->d_vals32, nllv, (smax_llv *) d->tv.llv,
                        d->d_rebuild, d->stream));
  const uint64_t nblocks = llv_rebuild_scratch_words(len) - 2;
  CU(cudaMemcpyAsync(&d->h_rebuild_total, d->d_rebuild + nblocks, sizeof(uint32_t), cudaMemcpyDeviceToHost,
                     d->stream));
  return 1;
}

static int build_llvdir(smax_device *d, char *err, size_t errlen)
{
  // the last tile may reach up to one tile past the covered range
  const uint64_t len = d->tv.a_hi - d->tv.a_lo + SMAX_PAD + smax_ring::kTileBytes;
  d->llvdir_entries = (size_t) ((len + (1u << kLlvBucketShift) - 1) >> kLlvBucketShift) + 3;
  CU(ensure_alloc((const void **) &d->tv.llvdir, &d->cap_dir,
                  d->llvdir_entries * sizeof(uint32_t)));
  CU(launch_llvdir(d->tv.llv, d->tv.nllv, d->tv.a_lo, (uint32_t *) d->tv.llvdir,
                   d->llvdir_entries, d->stream));
  // the compact records the scan streams instead of the 16-byte ones
  CU(ensure_alloc((const void **) &d->tv.llvv, &d->cap_llvv,
                  (d->tv.nllv + kLlvPad) * sizeof(uint32_t)));
  CU(ensure_alloc((const void **) &d->tv.llvp, &d->cap_llvp,
                  (d->tv.nllv + kLlvPad) * sizeof(uint32_t)));
  CU(cudaMemsetAsync(d->d_ctrl + 2, 0, 2 * sizeof(uint32_t), d->stream));
  CU(launch_llvpack(d->tv.llv, d->tv.nllv, d->tv.a_lo, (uint32_t *) d->tv.llvv, (uint32_t *) d->tv.llvp,
                    d->d_ctrl + 2, d->stream));
  {
    const uint64_t nunits = (d->g_hi - d->g_lo + kUnitBytes - 1) / kUnitBytes;
    CU(ensure_alloc((const void **) &d->d_unitdir, &d->cap_unitdir, (nunits + 2) * sizeof(uint32_t)));
    CU(launch_unitdir(d->tv.llv, d->tv.nllv, d->g_lo, d->g_hi, d->d_unitdir, nunits, d->stream));
  }
  // lcp value histogram of the own range (pick_kernel)
  CU(cudaMemsetAsync(d->d_hist, 0, 256 * sizeof(unsigned long long), d->stream));
  CU(launch_lcphist(d->tv.lcp + (d->g_lo - d->tv.a_lo), d->g_hi - d->g_lo, d->d_hist, d->sm_count,
                    d->stream));
  {
    // the order the unit kernel takes the units in: heaviest first
    const uint64_t nunits = (d->g_hi - d->g_lo + kUnitBytes - 1) / kUnitBytes;
    CU(ensure_alloc((const void **) &d->d_unitorder, &d->cap_unitorder,
                    (2 * nunits + kOrderScratch) * sizeof(uint32_t)));
    CU(launch_unitorder(d->tv.lcp + (d->g_lo - d->tv.a_lo), d->g_hi - d->g_lo, d->d_unitdir, d->d_hist,
                        d->d_unitorder, nunits, d->stream));
  }
  uint32_t esc[2] = {0, 0};        // [0] some value does not fit, [1] record 0 sits on the arrays' first entry
  CU(cudaMemcpyAsync(esc, d->d_ctrl + 2, sizeof esc, cudaMemcpyDeviceToHost, d->stream));
  CU(cudaMemcpyAsync(d->h_hist, d->d_hist, sizeof d->h_hist, cudaMemcpyDeviceToHost, d->stream));
  CU(cudaStreamSynchronize(d->stream));
  d->has_escape = esc[0] != 0;
  d->edge_rec0 = esc[1] != 0 && d->tv.a_lo > 0;
  return 0;
}

extern "C" int smax_device_upload(smax_device *d, const smax_index *idx, uint64_t lo,
                                  uint64_t hi, int with_suf, uint64_t *h2d_bytes,
                                  char *err, size_t errlen)
{
  return smax_device_upload_halo(d, idx, lo, hi, 256, with_suf, h2d_bytes, err, errlen);
}

extern "C" int smax_device_upload_halo(smax_device *d, const smax_index *idx, uint64_t lo,
                                       uint64_t hi, uint64_t halo, int with_suf,
                                       uint64_t *h2d_bytes, char *err, size_t errlen)
{
  smax_index_info info;
  smax_index_info_get(idx, &info);
  const uint64_t n = info.numberofallsortedsuffixes;
  const uint64_t base = idx->base, wend = idx->base + idx->len;   // host window
  const uint8_t *h_lcp = smax_index_lcptab(idx), *h_bwt = smax_index_bwttab(idx);
  const smax_llv *h_llv = smax_index_llvtab(idx);
  const void *h_suf = smax_index_suftab(idx);
  if (h_lcp == NULL || h_bwt == NULL)
    return fail(err, errlen, "index was opened without the lcp/bwt tables");
  if (with_suf && h_suf == NULL)
    return fail(err, errlen, "index was opened without the suffix table");
  if (hi > n) hi = n;
  if (lo > hi || (lo & 15) != 0)
    return fail(err, errlen, "shard range [%llu, %llu) must start at a multiple of 16",
                (unsigned long long) lo, (unsigned long long) hi);
  if (hi - lo > SMAX_MAX_SHARD_LEN)
    return fail(err, errlen, "a shard may hold at most 2^32 - 2^20 suffixes; use more shards");
  // coverage: a left halo (256 entries by default) so that almost every plateau
  // crossing the cut is resolved locally, 16 entries to the right for L[e+1]
  halo = std::max<uint64_t>(256, (halo + 15) & ~15ull);
  uint64_t a_lo = lo >= halo ? lo - halo : 0;
  if (a_lo < base) a_lo = (base + 15) & ~15ull;
  const uint64_t a_hi = std::min(std::min(n, hi + 16), wend);
  if (a_lo > lo || a_hi < hi || (hi < n && a_hi < hi + 1))
    return fail(err, errlen, "host tables [%llu, %llu) do not cover the shard [%llu, %llu) "
                "plus one entry", (unsigned long long) base, (unsigned long long) wend,
                (unsigned long long) lo, (unsigned long long) hi);
  CU(cudaSetDevice(d->ordinal));
  if (!d->owns_tables)
    free_tables(d);
  d->owns_tables = true;
  d->result_valid = false;
  d->scanned = false;
  d->seen_valid = false;
  const uint64_t len = a_hi - a_lo;
  if (len > SMAX_MAX_SHARD_LEN + (1ull << 19))
    return fail(err, errlen, "the window [%llu, %llu) is wider than one device shard may be",
                (unsigned long long) a_lo, (unsigned long long) a_hi);
  const size_t alloc = (size_t) ((len + 15) & ~15ull) + SMAX_PAD;
  {
    // lcp and bwt share one capacity figure
    size_t cap_bwt = d->cap_lcp;
    CU(ensure_alloc((const void **) &d->tv.lcp, &d->cap_lcp, alloc));
    CU(ensure_alloc((const void **) &d->tv.bwt, &cap_bwt, alloc));
  }
  uint8_t *d_lcp = (uint8_t *) d->tv.lcp, *d_bwt = (uint8_t *) d->tv.bwt;
  CU(cudaMemsetAsync(d_lcp + len, 0, alloc - len, d->stream));
  CU(cudaMemsetAsync(d_bwt + len, 0, alloc - len, d->stream));
  if (staged_h2d(d, d_lcp, h_lcp + (a_lo - base), len, h2d_bytes, err, errlen) != 0) return -1;
  if (staged_h2d(d, d_bwt, h_bwt + (a_lo - base), len, h2d_bytes, err, errlen) != 0) return -1;
  // .llv slice: records with position in [a_lo, a_hi)
  const uint64_t L = info.largelcpvalues;
  uint64_t k0 = 0, k1 = L;
  if (L > 0)
  {
    k0 = std::lower_bound(h_llv, h_llv + L, a_lo,
                          [](const smax_llv &r, uint64_t p) { return r.position < p; }) - h_llv;
    k1 = std::lower_bound(h_llv, h_llv + L, a_hi,
                          [](const smax_llv &r, uint64_t p) { return r.position < p; }) - h_llv;
  }
  d->tv.nllv = k1 - k0;
  if (d->tv.nllv >= (1ull << 32))
    return fail(err, errlen, "too many large lcp values in one shard");
  CU(ensure_alloc((const void **) &d->tv.llv, &d->cap_llv,
                  std::max<size_t>(16, d->tv.nllv * sizeof(smax_llv))));
  d->tv.a_lo = a_lo; d->tv.a_hi = a_hi;
  int stripped = d->tv.nllv != 0 ? upload_llv_stripped(d, h_llv + k0, d->tv.nllv, len, h2d_bytes, err, errlen) : 0;
  if (stripped < 0)
    return -1;
  if (stripped == 0 &&
      staged_h2d(d, (void *) d->tv.llv, h_llv + k0, d->tv.nllv * sizeof(smax_llv), h2d_bytes,
                 err, errlen) != 0)
    return -1;
  d->sufbytes = info.sufbytes ? info.sufbytes : 8;
  if (with_suf)
  {
    CU(ensure_alloc(&d->tv.suf, &d->cap_suf, len * info.sufbytes + SMAX_PAD));
    if (staged_h2d(d, (void *) d->tv.suf, (const char *) h_suf + (a_lo - base) * info.sufbytes,
                   len * info.sufbytes, h2d_bytes, err, errlen) != 0)
      return -1;
  } else if (d->tv.suf != NULL)
  {
    cudaFree((void *) d->tv.suf);
    d->tv.suf = NULL;
    d->cap_suf = 0;
  }
  d->tv.a_lo = a_lo; d->tv.a_hi = a_hi;
  d->g_lo = lo; d->g_hi = hi; d->n_total = n;
  if (build_llvdir(d, err, errlen) != 0) return -1;
  CU(cudaStreamSynchronize(d->stream));
  if (stripped == 1 && d->h_rebuild_total != d->tv.nllv)
  {
    // the lcp table and the .llv table disagree about the number of large values: the records
    // as they are, so that the scan reports what it always reported for such tables
    if (staged_h2d(d, (void *) d->tv.llv, h_llv + k0, d->tv.nllv * sizeof(smax_llv), h2d_bytes,
                   err, errlen) != 0)
      return -1;
    if (build_llvdir(d, err, errlen) != 0) return -1;
    CU(cudaStreamSynchronize(d->stream));
  }
  return 0;
}

extern "C" int smax_device_adopt(smax_device *d, const void *d_lcp, const void *d_bwt,
                                 const void *d_llv, uint64_t nllv, const void *d_suf,
                                 unsigned sufbytes, uint64_t a_lo, uint64_t a_hi,
                                 uint64_t lo, uint64_t hi, uint64_t n_total,
                                 char *err, size_t errlen)
{
  if ((a_lo & 15) || (lo & 15) || ((uintptr_t) d_lcp & 15) || ((uintptr_t) d_bwt & 15))
    return fail(err, errlen, "adopted tables must be 16-byte aligned and start at a multiple of 16");
  if (lo < a_lo || hi > a_hi || lo > hi || (hi < n_total && a_hi < hi + 1))
    return fail(err, errlen, "adopted coverage [%llu,%llu) does not contain the shard [%llu,%llu) plus one entry",
                (unsigned long long) a_lo, (unsigned long long) a_hi,
                (unsigned long long) lo, (unsigned long long) hi);
  if (a_hi - a_lo > SMAX_MAX_SHARD_LEN + 4096 || nllv >= (1ull << 32))
    return fail(err, errlen, "a shard may hold at most 2^32 - 2^20 suffixes; use more shards");
  if (d_suf != NULL && sufbytes != 8 && sufbytes != 4)
    return fail(err, errlen, "suffix table entries must be 8 or 4 bytes");
  CU(cudaSetDevice(d->ordinal));
  free_tables(d);
  d->result_valid = false;
  d->scanned = false;
  d->seen_valid = false;
  d->tv.lcp = (const uint8_t *) d_lcp; d->tv.bwt = (const uint8_t *) d_bwt;
  d->tv.llv = (const smax_llv *) d_llv; d->tv.nllv = nllv;
  d->tv.suf = d_suf; d->sufbytes = sufbytes ? sufbytes : 8;
  d->tv.a_lo = a_lo; d->tv.a_hi = a_hi;
  d->g_lo = lo; d->g_hi = hi; d->n_total = n_total;
  d->owns_tables = false;
  if (build_llvdir(d, err, errlen) != 0) return -1;
  CU(cudaStreamSynchronize(d->stream));
  return 0;
}

// ------------------------------------------------------------ peer views
extern "C" int smax_device_view(const smax_device *d, smax_shard_view *v)
{
  memset(v, 0, sizeof *v);
  v->a_lo = d->tv.a_lo; v->a_hi = d->tv.a_hi;
  v->d_lcp = (uint64_t) (uintptr_t) d->tv.lcp; v->d_bwt = (uint64_t) (uintptr_t) d->tv.bwt;
  v->d_llv = (uint64_t) (uintptr_t) d->tv.llv; v->d_llvdir = (uint64_t) (uintptr_t) d->tv.llvdir;
  v->d_suf = (uint64_t) (uintptr_t) d->tv.suf;
  v->nllv = d->tv.nllv; v->device = d->ordinal; v->sufbytes = d->sufbytes;
  return 0;
}

extern "C" int smax_device_set_left_views(smax_device *d, const smax_shard_view *views,
                                          int nviews, char *err, size_t errlen)
{
  if (nviews < 0 || nviews > kMaxLeft)
    return fail(err, errlen, "at most %d left neighbours are supported", kMaxLeft);
  CU(cudaSetDevice(d->ordinal));
  for (int k = 0; k < nviews; k++)
  {
    const smax_shard_view &v = views[k];
    if (v.a_lo > d->tv.a_lo)
      return fail(err, errlen, "left view %d does not lie left of the shard", k);
    // (by the END of their coverage: a neighbour that was made resident again with a wider halo
    // begins further left than the neighbours before it)
    if (k > 0 && views[k - 1].a_hi > v.a_hi)
      return fail(err, errlen, "left views must be in shard order (nearest neighbour last)");
    if (v.device != d->ordinal && v.device >= 0)
    {
      int can = 0;
      CU(cudaDeviceCanAccessPeer(&can, d->ordinal, v.device));
      if (!can)
        return fail(err, errlen, "device %d cannot access peer %d", d->ordinal, v.device);
      cudaError_t e = cudaDeviceEnablePeerAccess(v.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(err, errlen, "cudaDeviceEnablePeerAccess(%d): %s", v.device,
                    cudaGetErrorString(e));
      (void) cudaGetLastError();
    }
    TableView &t = d->left[k];
    t.lcp = (const uint8_t *) (uintptr_t) v.d_lcp; t.bwt = (const uint8_t *) (uintptr_t) v.d_bwt;
    t.llv = (const smax_llv *) (uintptr_t) v.d_llv; t.llvdir = (const uint32_t *) (uintptr_t) v.d_llvdir;
    t.suf = (const void *) (uintptr_t) v.d_suf;
    t.nllv = v.nllv; t.a_lo = v.a_lo; t.a_hi = v.a_hi;
  }
  d->nleft = nviews;
  return 0;
}

extern "C" int smax_device_ipc_export(const smax_device *d,
                                      uint8_t handles[SMAX_IPC_TABLES][SMAX_IPC_BYTES],
                                      smax_shard_view *view, char *err, size_t errlen)
{
  static_assert(sizeof(cudaIpcMemHandle_t) <= SMAX_IPC_BYTES, "ipc handle size");
  if (!d->owns_tables)
    return fail(err, errlen, "only tables uploaded by smax_device_upload can be exported");
  CU(cudaSetDevice(d->ordinal));
  const void *ptrs[SMAX_IPC_TABLES] = {d->tv.lcp, d->tv.bwt, d->tv.llv, d->tv.llvdir, d->tv.suf};
  memset(handles, 0, SMAX_IPC_TABLES * SMAX_IPC_BYTES);
  for (int k = 0; k < SMAX_IPC_TABLES; k++)
  {
    if (ptrs[k] == NULL) continue;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, (void *) ptrs[k]));
    memcpy(handles[k], &h, sizeof h);
  }
  return smax_device_view(d, view);
}

extern "C" int smax_device_ipc_import(smax_device *d,
                                      const uint8_t handles[SMAX_IPC_TABLES][SMAX_IPC_BYTES],
                                      smax_shard_view *v, char *err, size_t errlen)
{
  CU(cudaSetDevice(d->ordinal));
  uint64_t *slots[SMAX_IPC_TABLES] = {&v->d_lcp, &v->d_bwt, &v->d_llv, &v->d_llvdir, &v->d_suf};
  for (int k = 0; k < SMAX_IPC_TABLES; k++)
  {
    if (*slots[k] == 0) continue;
    if (d->n_ipc_mapped >= kMaxLeft * SMAX_IPC_TABLES)
      return fail(err, errlen, "too many imported IPC tables");
    cudaIpcMemHandle_t h;
    memcpy(&h, handles[k], sizeof h);
    void *p = NULL;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    d->ipc_mapped[d->n_ipc_mapped++] = p;
    *slots[k] = (uint64_t) (uintptr_t) p;
  }
  v->device = -1;   // already mapped into this process: no peer-enable needed
  return 0;
}

// ----------------------------------------------------------------- scan
// Which scan kernel runs?  Two kernels compute the same function (tests/test_gpu_parity.py runs
// every case through both):
//   units  (smax_scan.cu)  independent warps over 4 KiB units taken heaviest first, bitmaps +
//          arena + offset scan, walks that end at the first repeated left character: the default.
//          It is the faster one on every workload measured on B200 (uniform DNA 48 vs 54 us per
//          1e8 suffixes, C2 103 vs 144 us, C3 / C4 2-3x).
//   ring   (smax_ring.cu)  persistent CTAs over a TMA ring, survivors collected in a log and
//          written through a generation-wise prefix exchange: kept behind SMAX_KERNEL=ring as the
//          second, independently written implementation the parity tests compare with.
static int pick_kernel(const smax_device *d, uint64_t minlength, int policy)
{
  (void) d; (void) minlength; (void) policy;
  const char *env = getenv("SMAX_KERNEL");
  if (env != NULL && strcmp(env, "ring") == 0) return 0;
  return 1;
}

// entries of the survivor arena: one per record plus the chunk every warp of the grid may leave
// partly filled
static uint64_t arena_entries(const smax_device *d)
{
  return d->rec_cap + (uint64_t) d->sm_count * 16 * kWarps * kArenaChunk;
}

static int ensure_scratch(smax_device *d, uint64_t nunits, char *err, size_t errlen)
{
  if (d->rec_cap == 0)
  {
    const uint64_t len = d->g_hi - d->g_lo;
    d->rec_cap = std::max<uint64_t>(1u << 16, len / 16);
    CU(cudaMalloc(&d->d_recs, d->rec_cap * sizeof(smax_record)));
    CU(cudaMalloc(&d->d_arena, arena_entries(d) * sizeof(ArenaEntry)));
  }
  if (d->unit_cap < (size_t) (nunits + 1))
  {
    cudaFree(d->d_meta); cudaFree(d->d_blocksum);
    d->d_meta = NULL; d->d_blocksum = NULL;
    d->unit_cap = (size_t) (nunits + 1);
    d->blocks_cap = d->unit_cap / kEmitBlock + 2;
    CU(cudaMalloc(&d->d_meta, d->unit_cap * sizeof(UnitMeta)));
    CU(cudaMalloc(&d->d_blocksum, 2 * d->blocks_cap * 2 * sizeof(uint64_t)));
    CU(cudaMemsetAsync(d->d_blocksum, 0, 2 * d->blocks_cap * 2 * sizeof(uint64_t), d->stream));
    CU(cudaStreamSynchronize(d->stream));
  }
  if (d->pos_cap == 0)
  {
    d->pos_cap = d->rec_cap * 3;
    CU(cudaMalloc(&d->d_pos, d->pos_cap * sizeof(uint64_t)));
  }
  bool fresh = false;
  // look-back words: the unit kernel's offset scan (per block of units), the ring kernel (per tile)
  const size_t status_need = std::max<size_t>(
      (size_t) 8,
      2 * ((d->g_hi - d->g_lo) / smax_ring::kTileBytes + 2));
  if (d->status_cap < status_need)
  {
    cudaFree(d->d_status);
    d->status_cap = status_need;
    CU(cudaMalloc(&d->d_status, d->status_cap * sizeof(uint64_t)));
    fresh = true;
  }
  if (fresh || d->epoch >= kEpochMask)
  {
    // epoch 0 marks "never written"; only needed after (re)allocation or wrap
    CU(cudaMemsetAsync(d->d_status, 0, d->status_cap * sizeof(uint64_t), d->stream));
    CU(cudaStreamSynchronize(d->stream));
    d->epoch = 0;
  }
  return 0;
}

extern "C" int smax_scan_launch(smax_device *d, uint64_t minlength, int policy, int gather,
                                void *stream, char *err, size_t errlen)
{
  if (d->tv.lcp == NULL)
    return fail(err, errlen, "no tables resident on device %d", d->ordinal);
  if (gather && d->tv.suf == NULL)
    return fail(err, errlen, "position gather needs a resident suffix table");
  if (policy != SMAX_POLICY_GT && policy != SMAX_POLICY_PLAIN)
    return fail(err, errlen, "unknown left-character policy %d", policy);
  CU(cudaSetDevice(d->ordinal));
  cudaStream_t st = (cudaStream_t) stream;
  if (minlength == 0) minlength = 1;
  const uint64_t len = d->g_hi - d->g_lo;
  const uint64_t nunits = (len + kUnitBytes - 1) / kUnitBytes;
  if (ensure_scratch(d, nunits, err, errlen) != 0) return -1;

  const int kernel = pick_kernel(d, minlength, policy);
  d->last_kernel = kernel;
  if (kernel == 0)
  {
    // ---- the ring kernel
    smax_ring::ScanParams q;
    memset(&q, 0, sizeof q);
    auto view = [](const TableView &t)
    {
      smax_ring::TableView r;
      r.lcp = t.lcp; r.bwt = t.bwt; r.llv = t.llv; r.llvdir = t.llvdir; r.suf = t.suf;
      r.nllv = t.nllv; r.a_lo = t.a_lo; r.a_hi = t.a_hi;
      return r;
    };
    q.own = view(d->tv);
    for (int k = 0; k < d->nleft; k++) q.left[k] = view(d->left[k]);
    q.nleft = d->nleft;
    q.policy = policy;
    q.sufbytes = (int) d->sufbytes;
    q.debug = 0;
    q.epoch = ++d->epoch;
    q.g_lo = d->g_lo; q.g_hi = d->g_hi;
    q.minlength = minlength;
    q.mb = (uint32_t) std::min<uint64_t>(minlength, 255);
    const uint64_t ntiles = (len + smax_ring::kTileBytes - 1) / smax_ring::kTileBytes;
    q.ntiles = (uint32_t) ntiles;
    q.recs = d->d_recs; q.rec_capacity = d->rec_cap;
    q.positions = gather ? d->d_pos : NULL; q.pos_capacity = d->pos_cap;
    q.status = d->d_status;
    q.ctrl = d->d_ctrl;
    for (int k = 0; k < d->npeers; k++) q.peer_counts[k] = d->peer_counts[k];
    q.npeers = d->npeers; q.my_rank = d->my_rank;
    q.exchange_tag = d->exchange_tag & 0xffffffu;
    q.result = d->d_result + (d->scan_no & 1) * kResSlots;
    q.result_next = d->d_result + ((d->scan_no + 1) & 1) * kResSlots;
    const int bps = d->stats ? d->bps_ring_stats : d->bps_ring;
    int grid = (int) std::min<uint64_t>(std::max<uint64_t>(ntiles, 1), (uint64_t) d->sm_count * bps);
    if (d->grid_limit > 0)
      grid = std::min(grid, d->grid_limit);
    if (getenv("SMAX_TRACE") != NULL)
      fprintf(stderr, "# smax scan (ring kernel): %llu tiles, grid %d (%d CTAs/SM x %d SMs), minlength %llu\n",
              (unsigned long long) ntiles, grid, bps, d->sm_count, (unsigned long long) minlength);
    CU(cudaEventRecord(d->ev0, st));
    CU(smax_ring::launch_scan(q, d->stats, grid, st));
    d->last_launches = 1;
    CU(cudaEventRecord(d->ev1, st));
    d->last_stream = st;
    d->last_minlength = minlength; d->last_policy = policy; d->last_gather = gather;
    d->scan_no++;
    d->scanned = true;
    d->result_valid = false;
    d->text_valid = false;
    return 0;
  }

  // ---- the unit kernel
  ScanParams p;
  memset(&p, 0, sizeof p);
  p.own = d->tv;
  for (int k = 0; k < d->nleft; k++) p.left[k] = d->left[k];
  p.nleft = d->nleft;
  p.policy = policy;
  p.sufbytes = (int) d->sufbytes;
  p.debug = d->debug;
  p.epoch = ++d->epoch;
  p.g_lo = d->g_lo; p.g_hi = d->g_hi;
  p.minlength = minlength;
  p.mb = (uint32_t) std::min<uint64_t>(minlength, 255);
  p.nunits = (uint32_t) nunits;
  p.recs = d->d_recs; p.rec_capacity = d->rec_cap;
  p.positions = gather ? d->d_pos : NULL; p.pos_capacity = d->pos_cap;
  p.meta = d->d_meta;
  p.blocksum = d->d_blocksum + (d->unit_scan_no & 1) * d->blocks_cap * 2;
  p.blocksum_next = d->d_blocksum + ((d->unit_scan_no + 1) & 1) * d->blocks_cap * 2;
  d->unit_scan_no++;
  p.arena = d->d_arena; p.arena_capacity = arena_entries(d);
  {
    static const char *order_env = getenv("SMAX_ORDER");      // tuning: 0 = units as they come
    p.unitorder = order_env != NULL && atoi(order_env) == 0 ? NULL : d->d_unitorder;
  }
  p.unitdir = d->d_unitdir; p.has_escape = d->has_escape; p.edge_rec0 = d->edge_rec0;
  p.ctrl = d->d_ctrl;
  for (int k = 0; k < d->npeers; k++) p.peer_counts[k] = d->peer_counts[k];
  p.npeers = d->npeers; p.my_rank = d->my_rank;
  p.exchange_tag = d->exchange_tag & 0xffffffu;
  p.result = d->d_result + (d->scan_no & 1) * kResSlots;
  p.result_next = d->d_result + ((d->scan_no + 1) & 1) * kResSlots;

  // units are handed out by a ticket to the warps of a resident grid (any grid size is correct)
  const int bps = d->stats ? d->bps_scan_stats : d->bps_scan;
  static const char *grid_env = getenv("SMAX_GRID");
  uint64_t want = (uint64_t) d->sm_count * bps;
  if (grid_env != NULL && atoi(grid_env) > 0)
    want = (uint64_t) d->sm_count * atoi(grid_env);
  int grid = (int) std::min<uint64_t>(std::max<uint64_t>((nunits + kWarps - 1) / kWarps, 1),
                                      std::max<uint64_t>(want, 1));
  if (d->grid_limit > 0)
    grid = std::min(grid, d->grid_limit);
  if (getenv("SMAX_TRACE") != NULL)
    fprintf(stderr, "# smax scan (unit kernel): %llu units, grid %d (%d CTAs/SM x %d SMs), minlength %llu\n",
            (unsigned long long) nunits, grid, bps, d->sm_count, (unsigned long long) minlength);
  CU(cudaEventRecord(d->ev0, st));
  CU(launch_scan(p, d->stats, grid, d->sm_count, st));
  d->last_launches = kScanLaunches;
  CU(cudaEventRecord(d->ev1, st));
  d->last_stream = st;
  d->last_minlength = minlength; d->last_policy = policy; d->last_gather = gather;
  d->scan_no++;
  d->scanned = true;
  d->result_valid = false;
  d->text_valid = false;
  return 0;
}

static int read_result(smax_device *d, char *err, size_t errlen)
{
  if (!d->scanned)
    return fail(err, errlen, "no scan has been launched");
  if (d->result_valid)
    return 0;
  CU(cudaSetDevice(d->ordinal));
  const uint64_t *src = d->d_result + ((d->scan_no - 1) & 1) * kResSlots;
  CU(cudaMemcpyAsync(d->h_result, src, sizeof d->h_result, cudaMemcpyDeviceToHost,
                     d->last_stream));
  CU(cudaStreamSynchronize(d->last_stream));
  d->result_valid = true;
  if (!d->h_result[kResError] && !d->h_result[kResOverflow])
  {
    d->seen_valid = true;
    d->seen_minlength = d->last_minlength;
    d->seen_policy = (uint64_t) d->last_policy;
    d->seen_records = d->h_result[kResCount];
  }
  return 0;
}

extern "C" int smax_scan_counts(smax_device *d, uint64_t *nrecs, uint64_t *npositions,
                                char *err, size_t errlen)
{
  for (int attempt = 0; attempt < 4; attempt++)
  {
    if (read_result(d, err, errlen) != 0) return -1;
    if (d->h_result[kResError] == 6)
    {
      // a valid index whose plateau reaches further left than the resident left views do
      fail(err, errlen, "a plateau leaves the resident range of the tables (the shard's own "
                        "arrays and its left neighbour views)");
      return SMAX_E_RANGE;
    }
    if (d->h_result[kResError])
      return fail(err, errlen, "inconsistent ESA tables (code %llu): a 255 entry of the lcp table "
                               "has no .llv record, or a repeat is wider than a shard",
                  (unsigned long long) d->h_result[kResError]);
    if (!d->h_result[kResOverflow])
    {
      if (nrecs) *nrecs = d->h_result[kResCount];
      if (npositions) *npositions = d->last_gather ? d->h_result[kResPositions] : 0;
      return 0;
    }
    // output capacity was too small: counts are exact, grow and rescan
    // (the arena is cut into one region per warp of the grid: a region can overflow while the
    // total still fits -- then the capacity doubles)
    uint64_t need_recs = d->h_result[kResCount];
    const uint64_t need_pos0 = std::max<uint64_t>(d->h_result[kResPositions], 2 * need_recs);
    if (need_recs <= d->rec_cap && !(d->last_gather && need_pos0 > d->pos_cap))
      need_recs = 2 * d->rec_cap;
    if (need_recs > d->rec_cap)
    {
      cudaFree(d->d_recs); d->d_recs = NULL;
      cudaFree(d->d_arena); d->d_arena = NULL;
      d->rec_cap = need_recs + need_recs / 8 + 1024;
      CU(cudaMalloc(&d->d_recs, d->rec_cap * sizeof(smax_record)));
      CU(cudaMalloc(&d->d_arena, arena_entries(d) * sizeof(ArenaEntry)));
    }
    const uint64_t need_pos = std::max<uint64_t>(d->h_result[kResPositions], 2 * need_recs);
    if (d->last_gather && need_pos > d->pos_cap)
    {
      cudaFree(d->d_pos); d->d_pos = NULL;
      d->pos_cap = need_pos + need_pos / 8 + 1024;
      CU(cudaMalloc(&d->d_pos, d->pos_cap * sizeof(uint64_t)));
    }
    if (smax_scan_launch(d, d->last_minlength, d->last_policy, d->last_gather,
                         (void *) d->last_stream, err, errlen) != 0)
      return -1;
  }
  return fail(err, errlen, "output buffers kept overflowing");
}

extern "C" int smax_scan_fetch(smax_device *d, smax_record *recs, uint64_t *positions,
                               char *err, size_t errlen)
{
  uint64_t nrecs = 0, npos = 0;
  if (smax_scan_counts(d, &nrecs, &npos, err, errlen) != 0) return -1;
  CU(cudaSetDevice(d->ordinal));
  if (recs != NULL && nrecs > 0)
    CU(cudaMemcpyAsync(recs, d->d_recs, nrecs * sizeof(smax_record), cudaMemcpyDeviceToHost,
                       d->last_stream));
  if (positions != NULL && npos > 0)
    CU(cudaMemcpyAsync(positions, d->d_pos, npos * sizeof(uint64_t), cudaMemcpyDeviceToHost,
                       d->last_stream));
  CU(cudaStreamSynchronize(d->last_stream));
  return 0;
}

extern "C" int smax_scan_elapsed_ms(smax_device *d, float *ms, float *ms_scan, int *launches,
                                    char *err, size_t errlen)
{
  if (!d->scanned)
    return fail(err, errlen, "no scan has been launched");
  CU(cudaSetDevice(d->ordinal));
  CU(cudaEventSynchronize(d->ev1));
  if (ms) CU(cudaEventElapsedTime(ms, d->ev0, d->ev1));
  if (ms_scan) CU(cudaEventElapsedTime(ms_scan, d->ev0, d->ev1));   // one fused kernel
  if (launches) *launches = d->last_launches;
  return 0;
}

// ------------------------------------------------ one-sided count exchange
extern "C" int smax_device_counts_export(smax_device *d, int world, uint8_t handle[SMAX_IPC_BYTES],
                                         uint64_t *d_ptr, char *err, size_t errlen)
{
  if (world < 1 || world > SMAX_MAX_PEERS)
    return fail(err, errlen, "the count exchange supports 1..%d shards", SMAX_MAX_PEERS);
  CU(cudaSetDevice(d->ordinal));
  if (d->d_counts == NULL)
  {
    CU(cudaMalloc(&d->d_counts, SMAX_MAX_PEERS * sizeof(uint64_t)));
    CU(cudaMemset(d->d_counts, 0, SMAX_MAX_PEERS * sizeof(uint64_t)));
  }
  if (handle != NULL)
  {
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, d->d_counts));
    memset(handle, 0, SMAX_IPC_BYTES);
    memcpy(handle, &h, sizeof h);
  }
  if (d_ptr) *d_ptr = (uint64_t) (uintptr_t) d->d_counts;
  return 0;
}

extern "C" int smax_device_counts_connect(smax_device *d, int rank, int world,
                                          const uint8_t (*handles)[SMAX_IPC_BYTES],
                                          const uint64_t *d_ptrs, char *err, size_t errlen)
{
  if (world < 1 || world > SMAX_MAX_PEERS || rank < 0 || rank >= world)
    return fail(err, errlen, "bad rank %d / world %d for the count exchange", rank, world);
  if (d->d_counts == NULL)
    return fail(err, errlen, "smax_device_counts_export must be called first");
  CU(cudaSetDevice(d->ordinal));
  for (int k = 0; k < world; k++)
  {
    if (k == rank)
    {
      d->peer_counts[k] = d->d_counts;
      continue;
    }
    if (handles != NULL)
    {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles[k], sizeof h);
      void *p = NULL;
      if (d->counts_mapped[k] != NULL)
      {
        cudaIpcCloseMemHandle(d->counts_mapped[k]);
        d->counts_mapped[k] = NULL;
      }
      CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      d->counts_mapped[k] = p;
      d->peer_counts[k] = (uint64_t *) p;
    } else
    {
      if (d_ptrs == NULL || d_ptrs[k] == 0)
        return fail(err, errlen, "no address for the count array of shard %d", k);
      d->peer_counts[k] = (uint64_t *) (uintptr_t) d_ptrs[k];
      // an array of this process on another device: the kernel stores into it over NVLink
      cudaPointerAttributes attr;
      if (cudaPointerGetAttributes(&attr, d->peer_counts[k]) == cudaSuccess &&
          attr.type == cudaMemoryTypeDevice && attr.device != d->ordinal)
      {
        int can = 0;
        CU(cudaDeviceCanAccessPeer(&can, d->ordinal, attr.device));
        if (!can)
          return fail(err, errlen, "device %d cannot access peer %d", d->ordinal, attr.device);
        cudaError_t e = cudaDeviceEnablePeerAccess(attr.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(err, errlen, "cudaDeviceEnablePeerAccess(%d): %s", attr.device, cudaGetErrorString(e));
      }
      (void) cudaGetLastError();
    }
  }
  d->npeers = world;
  d->my_rank = rank;
  return 0;
}

extern "C" int smax_device_set_exchange_tag(smax_device *d, uint64_t tag)
{
  d->exchange_tag = tag;
  return 0;
}

extern "C" int smax_scan_peer_counts(smax_device *d, uint64_t tag, uint64_t *counts,
                                     char *err, size_t errlen)
{
  if (d->npeers == 0)
    return fail(err, errlen, "the count exchange is not connected");
  if (!d->scanned)
    return fail(err, errlen, "no scan has been launched");
  CU(cudaSetDevice(d->ordinal));
  CU(cudaStreamSynchronize(d->last_stream));
  uint64_t h[SMAX_MAX_PEERS];
  for (int spin = 0; spin < 2000000; spin++)
  {
    CU(cudaMemcpy(h, d->d_counts, d->npeers * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    bool all = true;
    for (int k = 0; k < d->npeers; k++)
      all = all && (h[k] >> 40) == (tag & 0xffffffu);
    if (all)
    {
      for (int k = 0; k < d->npeers; k++)
        counts[k] = h[k] & ((1ull << 40) - 1);
      return 0;
    }
  }
  return fail(err, errlen, "the peers' record counts for step %llu did not arrive",
              (unsigned long long) tag);
}

extern "C" int smax_scan_copy_count(smax_device *d, void *d_dst, void *stream,
                                    char *err, size_t errlen)
{
  if (!d->scanned)
    return fail(err, errlen, "no scan has been launched");
  CU(cudaSetDevice(d->ordinal));
  const uint64_t *src = d->d_result + ((d->scan_no - 1) & 1) * kResSlots + kResCount;
  CU(cudaMemcpyAsync(d_dst, src, sizeof(uint64_t), cudaMemcpyDeviceToDevice,
                     (cudaStream_t) stream));
  return 0;
}

extern "C" int smax_scan_device_buffers(smax_device *d, uint64_t *d_records,
                                        uint64_t *d_positions, uint64_t *d_count)
{
  if (d_records) *d_records = (uint64_t) (uintptr_t) d->d_recs;
  if (d_positions) *d_positions = (uint64_t) (uintptr_t) d->d_pos;
  if (d_count)
    *d_count = (uint64_t) (uintptr_t) (d->d_result + ((d->scan_no - 1) & 1) * kResSlots);
  return 0;
}

extern "C" int smax_device_set_grid_limit(smax_device *d, int max_ctas)
{
  d->grid_limit = max_ctas > 0 ? max_ctas : 0;
  return 0;
}

extern "C" int smax_device_set_debug(smax_device *d, int flags)
{
  d->debug = flags;
  return 0;
}

// tuning probe (tools/probe_units.py): the unit aggregates of the last scan of the unit kernel,
// 24 bytes per unit {count, pad, wsum, base}
extern "C" int smax_device_debug_meta(smax_device *d, void *out, uint64_t nunits)
{
  if (d->d_meta == NULL || nunits > d->unit_cap)
    return -1;
  if (cudaSetDevice(d->ordinal) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess)
    return -1;
  return cudaMemcpy(out, d->d_meta, nunits * sizeof(UnitMeta), cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : -1;
}

// tuning probe: the raw result block of the last scan (kResSlots words)
extern "C" int smax_device_debug_result(smax_device *d, uint64_t *out, int n)
{
  if (read_result(d, NULL, 0) != 0)
    return -1;
  for (int k = 0; k < n && k < (int) kResSlots; k++)
    out[k] = d->h_result[k];
  return (int) kResSlots;
}

extern "C" int smax_device_set_stats(smax_device *d, int on)
{
  d->stats = on != 0;
  return 0;
}

extern "C" int smax_scan_stats(smax_device *d, uint64_t stats[8], char *err, size_t errlen)
{
  if (read_result(d, err, errlen) != 0) return -1;
  stats[0] = d->g_hi - d->g_lo;
  stats[1] = d->h_result[kResStatCand];
  stats[2] = d->h_result[kResStatCandWidth];
  stats[3] = d->h_result[kResStatLlv];
  stats[4] = d->h_result[kResCount];
  stats[5] = d->h_result[kResStatSurvWidth];
  stats[6] = d->h_result[kResPositions];
  stats[7] = d->h_result[kResWalks] | ((uint64_t) d->last_kernel << 63);
  return 0;
}

// ---------------------------------------------------------------------------
// Device-side emit: separator table and text formatting
// ---------------------------------------------------------------------------
extern "C" int smax_device_set_separators(smax_device *d, const uint64_t *seps, uint64_t nseps,
                                          char *err, size_t errlen)
{
  if (nseps > 0 && seps == NULL)
    return fail(err, errlen, "smax_device_set_separators: null table");
  for (uint64_t k = 1; k < nseps; k++)
    if (seps[k - 1] >= seps[k])
      return fail(err, errlen, "separator positions must be strictly ascending");
  CU(cudaSetDevice(d->ordinal));
  CU(ensure_alloc((const void **) &d->d_seps, &d->cap_seps,
                  std::max<size_t>(16, nseps * sizeof(uint64_t))));
  if (nseps > 0)
    CU(cudaMemcpyAsync(d->d_seps, seps, nseps * sizeof(uint64_t), cudaMemcpyHostToDevice,
                       d->stream));
  CU(cudaStreamSynchronize(d->stream));
  d->nseps = nseps;
  d->text_valid = false;
  return 0;
}

extern "C" int smax_device_build_separators(smax_device *d, uint64_t *nseps_out,
                                            char *err, size_t errlen)
{
  if (d->tv.bwt == NULL || d->tv.suf == NULL)
    return fail(err, errlen, "the separator table is built from resident bwt and suffix tables");
  if (d->g_lo != 0 || d->g_hi != d->n_total || d->tv.a_lo != 0)
    return fail(err, errlen, "the separator table can only be built on a device that holds the "
                             "whole index; upload it with smax_device_set_separators instead");
  CU(cudaSetDevice(d->ordinal));
  const uint64_t n = d->n_total;
  const uint64_t nwords = (n + 63) / 64;
  uint64_t *bitmap = NULL, *rank = NULL, *bad = NULL;
  int rc = -1;
  cudaError_t e = cudaSuccess;
  uint64_t h_tail[1] = { 0 }, h_bad = 0;
  // transient: n/8 bytes of bitmap + n/8 bytes of ranks
  if ((e = cudaMalloc(&bitmap, (nwords + 1) * sizeof(uint64_t))) != cudaSuccess ||
      (e = cudaMalloc(&rank, (nwords + 1) * sizeof(uint64_t))) != cudaSuccess ||
      (e = cudaMalloc(&bad, sizeof(uint64_t))) != cudaSuccess ||
      (e = ensure_alloc((const void **) &d->d_fsums, &d->cap_fsums,
                        format_sums_words(nwords) * sizeof(uint64_t))) != cudaSuccess ||
      (e = cudaMemsetAsync(bitmap, 0, (nwords + 1) * sizeof(uint64_t), d->stream)) != cudaSuccess ||
      (e = cudaMemsetAsync(bad, 0, sizeof(uint64_t), d->stream)) != cudaSuccess ||
      (e = launch_sep_mark(d->tv.bwt, d->tv.suf, (int) d->sufbytes, n, bitmap, n, bad,
                           d->sm_count, d->stream)) != cudaSuccess ||
      (e = launch_sep_rank(bitmap, nwords, d->d_fsums, rank, d->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(h_tail, rank + nwords, sizeof(uint64_t), cudaMemcpyDeviceToHost,
                           d->stream)) != cudaSuccess ||
      (e = cudaMemcpyAsync(&h_bad, bad, sizeof(uint64_t), cudaMemcpyDeviceToHost,
                           d->stream)) != cudaSuccess ||
      (e = cudaStreamSynchronize(d->stream)) != cudaSuccess)
    goto done;
  if (h_bad != 0)
  {
    fail(err, errlen, "inconsistent ESA tables: %llu separator entries of the bwt table have "
                      "no valid suffix position", (unsigned long long) h_bad);
    goto done_noerr;
  }
  if ((e = ensure_alloc((const void **) &d->d_seps, &d->cap_seps,
                        std::max<size_t>(16, h_tail[0] * sizeof(uint64_t)))) != cudaSuccess ||
      (e = launch_sep_fill(bitmap, nwords, rank, d->d_seps, d->stream)) != cudaSuccess ||
      (e = cudaStreamSynchronize(d->stream)) != cudaSuccess)
    goto done;
  d->nseps = h_tail[0];
  d->text_valid = false;
  if (nseps_out) *nseps_out = d->nseps;
  rc = 0;
done:
  if (e != cudaSuccess)
    fail(err, errlen, "CUDA error: %s (separator table)", cudaGetErrorString(e));
done_noerr:
  cudaFree(bitmap); cudaFree(rank); cudaFree(bad);
  return rc;
}

extern "C" int smax_device_fetch_separators(smax_device *d, uint64_t *seps, uint64_t *nseps,
                                            char *err, size_t errlen)
{
  CU(cudaSetDevice(d->ordinal));
  if (nseps) *nseps = d->nseps;
  if (seps != NULL && d->nseps > 0)
    CU(cudaMemcpy(seps, d->d_seps, d->nseps * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return 0;
}

extern "C" int smax_scan_format(smax_device *d, int format, int relative, uint64_t *nbytes,
                                char *err, size_t errlen)
{
  if (format != SMAX_FORMAT_SMAX && format != SMAX_FORMAT_ITV)
    return fail(err, errlen, "the device formatter renders the smax and itv formats; format %d "
                             "is rendered by the host emitter", format);
  uint64_t nrecs = 0, npos = 0;
  if (smax_scan_counts(d, &nrecs, &npos, err, errlen) != 0) return -1;
  if (format == SMAX_FORMAT_SMAX && !d->last_gather)
    return fail(err, errlen, "the smax format needs the positions: launch the scan with gather");
  if (format == SMAX_FORMAT_ITV) npos = 0;
  CU(cudaSetDevice(d->ordinal));
  cudaStream_t st = d->last_stream;
  CU(ensure_alloc((const void **) &d->d_fsums, &d->cap_fsums,
                  format_sums_words(std::max(nrecs, npos)) * sizeof(uint64_t)));
  CU(ensure_alloc((const void **) &d->d_hoff, &d->cap_hoff, (nrecs + 1) * sizeof(uint64_t)));
  CU(ensure_alloc((const void **) &d->d_pfirst, &d->cap_pfirst, (nrecs + 1) * sizeof(uint64_t)));
  CU(ensure_alloc((const void **) &d->d_poff, &d->cap_poff, (npos + 1) * sizeof(uint64_t)));
  FormatJob j;
  memset(&j, 0, sizeof j);
  j.recs = d->d_recs; j.nrecs = nrecs;
  j.pos = d->d_pos; j.npos = npos;
  j.seps = d->d_seps; j.nseps = d->nseps;
  j.format = format; j.relative = relative != 0 && format == SMAX_FORMAT_SMAX;
  j.sums = d->d_fsums; j.hoff = d->d_hoff; j.pfirst = d->d_pfirst; j.poff = d->d_poff;
  // an upper bound of the text size (20 digits per number) lets measure and write run
  // back to back; only a huge result takes the exact size first (one host round trip)
  const uint64_t bound = format == SMAX_FORMAT_ITV ? nrecs * 63
                         : nrecs * 42 + npos * (j.relative ? 42 : 21);
  const bool one_go = bound <= std::max<uint64_t>(d->cap_text, 256ull << 20);
  uint64_t h_sizes[2] = { 0, 0 };
  if (one_go)
    CU(ensure_alloc((const void **) &d->d_text, &d->cap_text, std::max<size_t>(16, bound)));
  j.text = d->d_text;
  CU(cudaEventRecord(d->ev_f0, st));
  CU(launch_format_measure(j, st));
  if (one_go)
    CU(launch_format_write(j, st));
  CU(cudaMemcpyAsync(&h_sizes[0], d->d_hoff + nrecs, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  if (format == SMAX_FORMAT_SMAX)
    CU(cudaMemcpyAsync(&h_sizes[1], d->d_poff + npos, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
  if (one_go)
    CU(cudaEventRecord(d->ev_f1, st));
  CU(cudaStreamSynchronize(st));
  const uint64_t total = h_sizes[0] + h_sizes[1];
  if (!one_go)
  {
    CU(ensure_alloc((const void **) &d->d_text, &d->cap_text, std::max<size_t>(16, total)));
    j.text = d->d_text;
    CU(launch_format_write(j, st));
    CU(cudaEventRecord(d->ev_f1, st));
  }
  d->text_bytes = total;
  d->text_valid = true;
  if (nbytes) *nbytes = total;
  return 0;
}

extern "C" int smax_scan_fetch_text(smax_device *d, char *dst, char *err, size_t errlen)
{
  if (!d->text_valid)
    return fail(err, errlen, "no formatted text: call smax_scan_format after the scan");
  CU(cudaSetDevice(d->ordinal));
  if (d->text_bytes > 0)
    CU(cudaMemcpyAsync(dst, d->d_text, d->text_bytes, cudaMemcpyDeviceToHost, d->last_stream));
  CU(cudaStreamSynchronize(d->last_stream));
  return 0;
}

extern "C" int smax_scan_format_elapsed_ms(smax_device *d, float *ms, char *err, size_t errlen)
{
  if (!d->text_valid)
    return fail(err, errlen, "no formatted text: call smax_scan_format after the scan");
  CU(cudaSetDevice(d->ordinal));
  CU(cudaEventSynchronize(d->ev_f1));
  if (ms) CU(cudaEventElapsedTime(ms, d->ev_f0, d->ev_f1));
  return 0;
}
