// align_abi.cu -- host side of the C ABI declared in include/ambivert_align.h.
//
// Everything computes on the GPU (kernels.cuh).  There is deliberately no CPU implementation of
// any alignment in this library: if CUDA is unavailable the calls fail (NULL / ABV_ERR_NO_DEVICE).
#include <algorithm>
#include <array>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <chrono>
#include <memory>
#include <vector>

#include "../../include/ambivert_align.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "kernels.cuh"
#include <cuda.h>      // CUstream / CUdeviceptr of cuStreamWaitValue32, which is looked up at run time (cudaGetDriverEntryPoint)

namespace {

using namespace abv;

thread_local std::string t_err;
thread_local double t_timing[8] = {0, 0, 0, 0, 0, 0, 0, 0};

int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  t_err = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(ABV_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

// Device allocations go through a small per-device cache: cudaMalloc/cudaFree cost up to ~100 ms
// per job on a busy box (cudaFree synchronises the device), which showed up as end-to-end time.  Freed
// blocks are kept (up to a byte budget) and handed out again to requests of a similar size.  All callers
// hold the device context's mutex, and every user of a block has synchronised its stream before the block is
// released (abv_job_free waits for the job's last launch stream).
struct DevCache {
  struct Block { void *p; size_t bytes; };
  std::vector<Block> free_blocks;
  size_t cached_bytes = 0;
  size_t budget = (size_t)8 << 30;
  cudaError_t get(size_t n, void **out, size_t *got) {
    int best = -1;
    for (int i = 0; i < (int)free_blocks.size(); ++i)
      if (free_blocks[i].bytes >= n && free_blocks[i].bytes <= 2 * n + 4096 && (best < 0 || free_blocks[i].bytes < free_blocks[best].bytes)) best = i;
    if (best >= 0) {
      *out = free_blocks[best].p; *got = free_blocks[best].bytes;
      cached_bytes -= free_blocks[best].bytes;
      free_blocks.erase(free_blocks.begin() + best);
      return cudaSuccess;
    }
    cudaError_t e = cudaMalloc(out, n);
    if (e != cudaSuccess && !free_blocks.empty()) {   // out of memory: drop the cache and retry once
      cudaGetLastError();
      trim(0);
      e = cudaMalloc(out, n);
    }
    *got = n;
    return e;
  }
  void put(void *p, size_t bytes) {
    if (bytes > budget / 2) { cudaFree(p); return; }
    free_blocks.push_back({p, bytes});
    cached_bytes += bytes;
    if (cached_bytes > budget) trim(budget / 2);
  }
  void trim(size_t keep) {
    while (!free_blocks.empty() && cached_bytes > keep) {
      cudaFree(free_blocks.front().p);
      cached_bytes -= free_blocks.front().bytes;
      free_blocks.erase(free_blocks.begin());
    }
  }
};

// Page-locked host staging blocks, cached like the device blocks.  A copy from pageable memory holds the calling thread until the
// driver has staged it (0.3-0.5 ms per 1.6 MB offsets array); from a pinned block it is queued and the host goes on.
struct PinCache {
  struct Block { void *p; size_t bytes; };
  std::vector<Block> free_blocks;
  size_t cached_bytes = 0;
  cudaError_t get(size_t n, void **out, size_t *got) {
    int best = -1;
    for (int i = 0; i < (int)free_blocks.size(); ++i)
      if (free_blocks[i].bytes >= n && free_blocks[i].bytes <= 2 * n + 4096 && (best < 0 || free_blocks[i].bytes < free_blocks[best].bytes)) best = i;
    if (best >= 0) {
      *out = free_blocks[best].p; *got = free_blocks[best].bytes;
      cached_bytes -= free_blocks[best].bytes;
      free_blocks.erase(free_blocks.begin() + best);
      return cudaSuccess;
    }
    *got = n;
    return cudaHostAlloc(out, n, cudaHostAllocDefault);
  }
  void put(void *p, size_t bytes) {
    free_blocks.push_back({p, bytes});
    cached_bytes += bytes;
    while (cached_bytes > ((size_t)256 << 20) && !free_blocks.empty()) {      // at most 256 MB stay locked
      cudaFreeHost(free_blocks.front().p);
      cached_bytes -= free_blocks.front().bytes;
      free_blocks.erase(free_blocks.begin());
    }
  }
};

// ---- device contexts -----------------------------------------------------------------------------------------------
// One context per CUDA device (streams, events, pinned words, allocation cache, a mutex).  A call works on the context of
// the CALLING THREAD's device: the process default (abv_set_device; one process per GPU under torchrun) unless the thread
// chose another one (abv_set_thread_device; the worker threads of abv_best_hits_multi do, one per device of the box).
// Calls on different devices run concurrently; calls on the same device are serialised by its mutex.
struct Options {
  int path = ABV_PATH_AUTO;
  int wf32_blocks_per_sm = 4;
  int mm16_blocks_per_sm = 0;   // 0 = what the occupancy query returns
  int split = 0;                // 0 = automatic; otherwise slices of the target pairs per query
  int fine_tail = 1;            // 1 = the last work items of a packed launch are cut finer (short tail of the persistent grid)
  int split_min_items = 4;      // a launch with fewer work items per resident CTA than this cuts every query into slices of the target pairs
                                // (measured on a rank-sized shard of 12.5k queries: 24 -> 8617, 12 -> 8683, 6 -> 8696, 3 -> 8698, 1 -> 8661 GCUPS:
                                // every slice rebuilds the query profile, and the finer tail items already keep the tail short)
  int global_v1 = 0;            // 1 = use the first-generation packed kernel for global mode too
  int no_gop_imm = 0;           // 1 = never take the streaming kernel's instantiation with the gap addend as an immediate
  int no_levels = 0;            // 1 = streaming global kernel resets every pair explicitly (no level staircase, mm16g_levels)
  int no_border = 0;            // 1 = streaming global never takes the border-lane instantiation (every query keeps all 32 lanes)
  int concurrent_buckets = 1;   // 1 = the packed launches of a job (one per strip width) go to streams of their own, largest first:
                                //     the CTAs of the next launch fill the tail of the previous persistent grid
  int tb16 = 1;                 // 0 = one-to-one jobs always take the 32-bit kernel
  int tb16_extra_smem = 0;      // experiment knob: unused dynamic shared memory per CTA of the packed traceback kernel (lowers its occupancy)
  int pairs_chunk = 200000;     // one-to-one jobs of at least 2x this many pairs run as a pipeline of chunks (0 = never)
};
Options g_opt;                  // process-wide

constexpr int ABV_MAX_DEVICES = 64;
struct Ctx {
  std::recursive_mutex mu;
  bool inited = false;
  int device = 0;
  int sms = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;      // second lane of the chunk pipeline of the one-to-one jobs
  cudaStream_t stream3 = nullptr;      // its downloads
  unsigned long long *pin64 = nullptr; // 64 pinned words: counters read back per chunk (a copy into pageable memory would block the host until the chunk has finished)
  cudaEvent_t ev[6] = {};
  DevCache cache;
  PinCache pin_cache;
  bool md5_ready = false;              // constant memory is per device: the MD5 table has been uploaded here
};
Ctx g_ctx[ABV_MAX_DEVICES];
int g_default_device = 0;
thread_local int t_device = -1;        // -1: the process default
inline Ctx &cur_ctx() { return g_ctx[t_device >= 0 ? t_device : g_default_device]; }
#define g (cur_ctx())

// runs the enclosed calls on another device's context (jobs remember the device they were prepared on)
struct DeviceScope {
  int saved;
  explicit DeviceScope(int device) : saved(t_device) { t_device = device; }
  ~DeviceScope() { t_device = saved; if (cur_ctx().inited) cudaSetDevice(cur_ctx().device); }
};

int ctx_init() {
  Ctx &c = cur_ctx();
  const int want = t_device >= 0 ? t_device : g_default_device;
  if (c.inited) {
    // the CUDA "current device" is per host thread: worker threads (multiprocessing.dummy, abv_best_hits_multi) arrive unset
    if (cudaSetDevice(c.device) != cudaSuccess) return fail(ABV_ERR_CUDA, "cudaSetDevice(%d): %s", c.device, cudaGetErrorString(cudaGetLastError()));
    return ABV_OK;
  }
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return fail(ABV_ERR_NO_DEVICE, "no usable CUDA device (%s); this library has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (want >= n) return fail(ABV_ERR_ARGS, "device %d out of range (%d devices)", want, n);
  c.device = want;
  CU(cudaSetDevice(c.device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, c.device));
  if (prop.major < 10)
    return fail(ABV_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", c.device, prop.major, prop.minor);
  c.sms = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c.stream2, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c.stream3, cudaStreamNonBlocking));
  CU(cudaHostAlloc((void **)&c.pin64, 64 * sizeof(unsigned long long), cudaHostAllocDefault));
  for (auto &ev : c.ev) CU(cudaEventCreate(&ev));
  c.inited = true;
  return ABV_OK;
}

// cuStreamWaitValue32: a stream waits until a 32-bit word in device memory reaches a value.  The packed launches of a job use it
// to start the grid of the next strip width exactly when the previous persistent grid begins to drain (Mm16Params::tail_flag).
typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWaitValue32Fn stream_wait_value32() {
  static StreamWaitValue32Fn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &f, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess) { cudaGetLastError(); f = nullptr; }
    return (StreamWaitValue32Fn)f;
  }();
  return fn;
}

// RAII device allocation
struct DevMem {
  void *p = nullptr;
  size_t bytes = 0;
  Ctx *owner = nullptr;         // the device context whose cache the block came from (and goes back to)
  DevMem() = default;
  DevMem(const DevMem &) = delete;
  DevMem &operator=(const DevMem &) = delete;
  ~DevMem() { release(); }
  void release() { if (p) owner->cache.put(p, bytes); p = nullptr; bytes = 0; }
  cudaError_t alloc(size_t n) {
    release();
    if (n == 0) n = 16;
    n = (n + 255) & ~(size_t)255;
    owner = &cur_ctx();
    return owner->cache.get(n, &p, &bytes);
  }
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// RAII page-locked host block from the context's cache
struct PinMem {
  void *p = nullptr;
  size_t bytes = 0;
  Ctx *owner = nullptr;
  PinMem() = default;
  PinMem(const PinMem &) = delete;
  PinMem &operator=(const PinMem &) = delete;
  ~PinMem() { release(); }
  void release() { if (p) owner->pin_cache.put(p, bytes); p = nullptr; bytes = 0; }
  cudaError_t alloc(size_t n) {
    if (p && bytes >= n) return cudaSuccess;
    release();
    owner = &cur_ctx();
    return owner->pin_cache.get(std::max<size_t>(n, 256), &p, &bytes);
  }
  template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

struct Scoring {
  int alpha, go, ge;
  const unsigned char *map256;   // may be NULL (inputs are codes)
  const int *matrix;
};

int check_scoring(const Scoring &sc) {
  if (!sc.matrix) return fail(ABV_ERR_ARGS, "score_matrix is NULL");
  if (sc.alpha <= 0 || sc.alpha > 256) return fail(ABV_ERR_ARGS, "alpha_len %d out of range", sc.alpha);
  if (sc.go > 0 || sc.ge > 0) return fail(ABV_ERR_ARGS, "gap penalties must be <= 0 (global_align.c:121)");
  return ABV_OK;
}

// lengths of a concatenated sequence set; rejects what the reference rejects
int check_seqs(const char *buf, const long long *off, int n, const char *what, int &maxlen, int &minlen, int base = 0) {
  maxlen = 0; minlen = INT_MAX;
  if (n < 0) return fail(ABV_ERR_ARGS, "%s: negative count", what);
  if (n == 0) { minlen = 0; return ABV_OK; }
  if (!buf || !off) return fail(ABV_ERR_ARGS, "%s: NULL buffer", what);
  for (int i = 0; i < n; ++i) {
    long long l = off[i + 1] - off[i];
    if (l <= 0) return fail(ABV_ERR_ARGS, "%s %d has length %lld (the reference returns NULL for length <= 0)", what, base + i, l);
    if (l > (1 << 24)) return fail(ABV_ERR_ARGS, "%s %d is longer than 2^24", what, base + i);
    maxlen = std::max(maxlen, (int)l);
    minlen = std::min(minlen, (int)l);
  }
  return ABV_OK;
}

// Symbols outside the alphabet: the reference indexes score_matrix[a*alpha_len + b] with whatever the map or the caller
// supplies (helpers.h:26-32, global_align.c:148), i.e. reads past the matrix.  The batched calls reject such input.  A map
// whose entries are all < alpha_len (DNA_MAP, PROT_MAP) needs no look at the data; otherwise (make_map(s, None): unknown ==
// alpha_len, or raw codes) the bytes are scanned.
int check_symbols(const char *buf, const long long *off, int n, const Scoring &sc, const char *what) {
  if (n <= 0) return ABV_OK;
  bool bad[256];
  bool any = false;
  for (int i = 0; i < 256; ++i) { bad[i] = (sc.map256 ? sc.map256[i] : i) >= sc.alpha; any = any || bad[i]; }
  if (!any || (sc.map256 && sc.alpha > 255)) return ABV_OK;
  const unsigned char *p = (const unsigned char *)buf + off[0], *e = (const unsigned char *)buf + off[n];
  unsigned hit = 0;
  for (; p < e; ++p) hit |= bad[*p];
  if (hit) return fail(ABV_ERR_ARGS, "%s: a symbol maps outside the alphabet (code >= alpha_len %d; the reference reads past its score matrix there)", what, sc.alpha);
  return ABV_OK;
}

int check_mode_lengths(int mode, int min_la, int min_lb) {
  if (mode < 0 || mode > 3) return fail(ABV_ERR_ARGS, "unknown mode %d", mode);
  if (mode == ABV_GLOCAL && min_lb < 2) return fail(ABV_ERR_UNDEFINED, "glocal with a target of length 1 is undefined in the reference (glocal_align.c:138,172)");
  if (mode == ABV_OVERLAP && min_la < 2) return fail(ABV_ERR_UNDEFINED, "overlap with a first sequence of length 1 is undefined in the reference (align.c:183)");
  return ABV_OK;
}

// upload a concatenated sequence set and encode it on the device
// Host buffers an upload reads asynchronously; they must outlive the stream work (the caller syncs before dropping them).
struct UploadKeep {
  PinMem rel[4];               // offsets relative to the first sequence, page-locked: their copies do not hold the host
  unsigned char clamped[256];
  int used = 0;
};

// Sequences -> device: raw bytes + offsets, then to_raw on the device (helpers.h:26-32).  With `keep` nothing here waits
// for the stream (pinned caller buffers then copy while the host goes on planning); without it the call returns synchronised.
int upload_encoded(const char *buf, const long long *off, int n, const Scoring &sc, cudaStream_t st,
                   DevMem &d_codes, DevMem &d_off, DevMem &d_map, double &h2d_bytes, DevMem *keep_raw = nullptr, UploadKeep *keep = nullptr,
                   int phase = 0) {
  // phase 0: everything; 1: the small pageable pieces only (offsets, map); 2: the sequence bytes + encode only.  A pageable
  // copy holds the host until the copy engine has served it, so a pipelined caller sends all small pieces of a chunk before
  // the first large (pinned, truly asynchronous) one.
  UploadKeep local;
  UploadKeep &K = keep ? *keep : local;
  const long long total = n ? off[n] - off[0] : 0;
  if (phase != 2) {
    PinMem &relm = K.rel[K.used++ & 3];
    CU(relm.alloc(sizeof(long long) * ((size_t)n + 1)));
    long long *rel = relm.as<long long>();
    for (int i = 0; i <= n; ++i) rel[i] = n ? off[i] - off[0] : 0;
    CU(d_off.alloc(sizeof(long long) * (n + 1)));
    CU(cudaMemcpyAsync(d_off.p, rel, sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, st));
    unsigned char ident[256];
    const unsigned char *m = sc.map256;
    if (!m) { for (int i = 0; i < 256; ++i) ident[i] = (unsigned char)i; m = ident; }
    for (int i = 0; i < 256; ++i) K.clamped[i] = (unsigned char)std::min<int>(m[i], sc.alpha - 1);
    if (!d_map.p) CU(d_map.alloc(256));
    CU(cudaMemcpyAsync(d_map.p, K.clamped, 256, cudaMemcpyHostToDevice, st));
    h2d_bytes += sizeof(long long) * (n + 1) + 256;
  }
  if (phase != 1) {
    CU(d_codes.alloc((size_t)total + 16));
    DevMem raw;
    if (total > 0) {
      CU(raw.alloc((size_t)total));
      CU(cudaMemcpyAsync(raw.p, buf + off[0], (size_t)total, cudaMemcpyHostToDevice, st));
      int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
      encode_kernel<<<blocks, 256, 0, st>>>(raw.as<uint8_t>(), d_codes.as<uint8_t>(), total, d_map.as<uint8_t>());
      CU(cudaGetLastError());
      if (keep_raw) { keep_raw->release(); keep_raw->p = raw.p; keep_raw->bytes = raw.bytes; keep_raw->owner = raw.owner; raw.p = nullptr; raw.bytes = 0; }
    }
    h2d_bytes += (double)total;
  }
  if (!keep) CU(cudaStreamSynchronize(st));     // the staging vectors are locals
  return ABV_OK;                                // `raw` returns to the allocation cache; its reuse is ordered on the same stream
}

// ---- wf32 launcher ----------------------------------------------------------------------------
// shared-memory staging of the target + border rows: 9 bytes per target position per warp
int wf32_stage_len(int max_lb, int alpha) {
  if (alpha > WF32_MAX_SMEM_ALPHA) return 0;
  const int len = (max_lb + 15) & ~15;
  return (size_t)len * 9 * WF32_WARPS <= 64 * 1024 ? len : 0;      // 0: the batch is too long, use the global-memory variant
}

template <int MODE, bool TRACE>
int wf32_launch_mode(int grid, cudaStream_t st, Wf32Params p, int max_lb) {
  p.stage_len = wf32_stage_len(max_lb, p.alpha);
  if (p.stage_len > 0) {
    const int smem = p.stage_len * 9 * WF32_WARPS;
    auto kern = wf32_kernel<MODE, TRACE, true>;
    if (smem > 48 * 1024 - (int)sizeof(int) * WF32_MAX_SMEM_ALPHA * WF32_MAX_SMEM_ALPHA)
      CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kern<<<grid, WF32_WARPS * 32, smem, st>>>(p);
  } else {
    wf32_kernel<MODE, TRACE, false><<<grid, WF32_WARPS * 32, 0, st>>>(p);
  }
  CU(cudaGetLastError());
  return ABV_OK;
}

template <bool TRACE>
int wf32_dispatch(int mode, int grid, cudaStream_t st, const Wf32Params &p, int max_lb) {
  switch (mode) {
    case ABV_OVERLAP: return wf32_launch_mode<OVERLAP, TRACE>(grid, st, p, max_lb);
    case ABV_LOCAL:   return wf32_launch_mode<LOCAL, TRACE>(grid, st, p, max_lb);
    case ABV_GLOBAL:  return wf32_launch_mode<GLOBAL, TRACE>(grid, st, p, max_lb);
    default:          return wf32_launch_mode<GLOCAL, TRACE>(grid, st, p, max_lb);
  }
}

int wf32_grid(long long n_pairs, size_t scratch_per_warp) {
  long long want = (n_pairs + WF32_WARPS - 1) / WF32_WARPS;
  long long cap = (long long)g.sms * g_opt.wf32_blocks_per_sm;
  const size_t budget = (size_t)3 << 30;   // bound the per-warp scratch to 3 GiB
  if (scratch_per_warp > 0) {
    long long by_mem = (long long)(budget / (scratch_per_warp * WF32_WARPS));
    cap = std::max<long long>(1, std::min(cap, by_mem));
  }
  return (int)std::max<long long>(1, std::min(want, cap));
}

// ---- mm16 launcher ----------------------------------------------------------------------------
const int MM16_KS[] = {2, 3, 4, 5, 6, 7, 8, 9, 10, 12, 16};   // odd K: streaming kernel (global, glocal) only
constexpr int MM16_NK = 11;

// Work items per launch: a query is cut into `split` slices of the target pairs when there are too few
// queries to keep the tail of the persistent grid short (each slice keeps >= 4 pairs per warp).
int pick_split(int n_list, int n_pairs, int resident_ctas) {
  if (g_opt.split > 0) return std::max(1, std::min(g_opt.split, std::max(1, n_pairs / MM16_WARPS)));
  int split = 1;
  while (split < 16 && (long long)n_list * split < (long long)g_opt.split_min_items * resident_ctas && n_pairs / (split * 2) >= 4 * MM16_WARPS) split *= 2;
  return split;
}
// Work items of one packed launch, built on the host and kept on the device: item = (query, slice of the target pairs).
// A query is cut into `split` slices when there are too few queries to keep the persistent grid busy (multi-GPU shards).  The
// tail of a persistent grid is as long as its last work item, and the CTAs finish their last coarse item spread over one item's
// duration: so about one coarse item per CTA of work is cut as fine as the pipeline allows (>= 4 target pairs per warp and
// slice) and handed out LAST -- it fills that span and leaves a tail of one fine item (1/16 of a query instead of a whole one).
struct WorkItems {
  std::vector<int4> host;          // stays alive while the upload may be in flight (the job owns it)
  DevMem dev;
  int n = 0;
  int key[5] = {-1, -1, -1, -1, -1};   // resident CTAs, n_pairs, n_list, split option, fine_tail option
};
int build_work_items(WorkItems &W, const int *h_list, int n_list, int n_pairs, int resident_ctas, cudaStream_t st) {
  const int key[5] = {resident_ctas, n_pairs, n_list, g_opt.split * 1000 + g_opt.split_min_items, g_opt.fine_tail};
  if (W.dev.p && !memcmp(key, W.key, sizeof key)) return ABV_OK;
  const int split = pick_split(n_list, n_pairs, resident_ctas);
  int fine = split, n_fine = 0;
  if (g_opt.fine_tail) {
    while (fine < 16 && n_pairs / (fine * 2) >= 4 * MM16_WARPS) fine *= 2;
    if (fine > split) n_fine = (int)std::min<long long>(n_list, (resident_ctas + split - 1) / split);
  }
  W.host.clear();
  W.host.reserve((size_t)(n_list - n_fine) * split + (size_t)n_fine * fine);
  auto cut = [&](int li, int parts) {
    for (int part = 0; part < parts; ++part) {
      const int lo = (int)((long long)n_pairs * part / parts), hi = (int)((long long)n_pairs * (part + 1) / parts);
      if (hi > lo) W.host.push_back(make_int4(h_list[li], lo, hi, 0));
    }
  };
  for (int li = 0; li < n_list - n_fine; ++li) cut(li, split);
  for (int li = n_list - n_fine; li < n_list; ++li) cut(li, fine);
  W.n = (int)W.host.size();
  CU(W.dev.alloc(sizeof(int4) * std::max<size_t>(W.host.size(), 1)));
  CU(cudaMemcpyAsync(W.dev.p, W.host.data(), sizeof(int4) * W.host.size(), cudaMemcpyHostToDevice, st));
  memcpy(W.key, key, sizeof key);
  return ABV_OK;
}

template <int MODE, int K>
int mm16_launch_one(int grid_cap, cudaStream_t st, const Mm16Params &p, int &launched, WorkItems &W, const int *h_list) {
  const Mm16Smem L = mm16_smem_layout(K, p.a1, p.tp_stride);
  auto kern = mm16_kernel<MODE, K>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, MM16_WARPS * 32, L.total));
  if (per_sm < 1) return fail(ABV_ERR_CUDA, "mm16 kernel K=%d does not fit (smem %d)", K, L.total);
  if (g_opt.mm16_blocks_per_sm > 0) per_sm = std::min(per_sm, g_opt.mm16_blocks_per_sm);
  Mm16Params q = p;
  int rc;
  if ((rc = build_work_items(W, h_list, p.n_list, p.n_pairs, g.sms * per_sm, st))) return rc;
  q.items = W.dev.as<int4>(); q.n_items = W.n;
  if (W.n == 0) return ABV_OK;
  int grid = (int)std::min<long long>(W.n, g.sms * per_sm);
  if (grid_cap > 0) grid = std::min(grid, grid_cap);
  kern<<<grid, MM16_WARPS * 32, L.total, st>>>(q);
  CU(cudaGetLastError());
  ++launched;
  return ABV_OK;
}

template <int MODE, int K, int GOPC, bool BORDER>
int mm16g_launch_gop(cudaStream_t st, const Mm16Params &p, int &launched, WorkItems &W, const int *h_list) {
  const Mm16gSmem L = mm16g_smem_layout(K, p.a1, p.tp_stride, mm16g_defer(MODE == GLOCAL, K));
  auto kern = mm16g_kernel<MODE, K, GOPC, BORDER>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, MM16_WARPS * 32, L.total));
  if (per_sm < 1) return fail(ABV_ERR_CUDA, "mm16g kernel K=%d does not fit (smem %d)", K, L.total);
  if (g_opt.mm16_blocks_per_sm > 0) per_sm = std::min(per_sm, g_opt.mm16_blocks_per_sm);
  Mm16Params q = p;
  int rc;
  if ((rc = build_work_items(W, h_list, p.n_list, p.n_pairs, g.sms * per_sm, st))) return rc;
  q.items = W.dev.as<int4>(); q.n_items = W.n;
  if (W.n == 0) return ABV_OK;
  const int grid = (int)std::min<long long>(W.n, g.sms * per_sm);
  kern<<<grid, MM16_WARPS * 32, L.total, st>>>(q);
  CU(cudaGetLastError());
  ++launched;
  return ABV_OK;
}

template <int MODE, int K, bool BORDER>
int mm16g_launch_var(cudaStream_t st, const Mm16Params &p, int &launched, WorkItems &W, const int *h_list) {
  // the pipeline's gap penalties (-7, -1) take the instantiation with the gap addend as an immediate
  if (p.go - p.ge == MM16G_GOP_DNA && !g_opt.no_gop_imm) return mm16g_launch_gop<MODE, K, MM16G_GOP_DNA, BORDER>(st, p, launched, W, h_list);
  return mm16g_launch_gop<MODE, K, MM16G_GOP_RUNTIME, BORDER>(st, p, launched, W, h_list);
}
template <int MODE, int K>
int mm16g_launch_one(bool border, cudaStream_t st, const Mm16Params &p, int &launched, WorkItems &W, const int *h_list) {
  if constexpr (mm16g_border(K, MODE == GLOCAL)) {
    if (border) return mm16g_launch_var<MODE, K, true>(st, p, launched, W, h_list);
  }
  return mm16g_launch_var<MODE, K, false>(st, p, launched, W, h_list);
}

template <int MODE>
int mm16g_launch_mode(int K, bool border, cudaStream_t st, const Mm16Params &p, int &launched, WorkItems &W, const int *h_list) {
  switch (K) {
#ifndef ABV_TUNE_FAST   // tuning builds (tools/sweep_builds.sh) carry the strip widths of the 200..260 bp panel only
    case 2: return mm16g_launch_one<MODE, 2>(border, st, p, launched, W, h_list);
    case 3: return mm16g_launch_one<MODE, 3>(border, st, p, launched, W, h_list);
    case 4: return mm16g_launch_one<MODE, 4>(border, st, p, launched, W, h_list);
    case 5: return mm16g_launch_one<MODE, 5>(border, st, p, launched, W, h_list);
    case 6: return mm16g_launch_one<MODE, 6>(border, st, p, launched, W, h_list);
    case 10: return mm16g_launch_one<MODE, 10>(border, st, p, launched, W, h_list);
    case 12: return mm16g_launch_one<MODE, 12>(border, st, p, launched, W, h_list);
    case 16: return mm16g_launch_one<MODE, 16>(border, st, p, launched, W, h_list);
#endif
    case 7: return mm16g_launch_one<MODE, 7>(border, st, p, launched, W, h_list);
    case 8: return mm16g_launch_one<MODE, 8>(border, st, p, launched, W, h_list);
    case 9: return mm16g_launch_one<MODE, 9>(border, st, p, launched, W, h_list);
  }
  return fail(ABV_ERR_ARGS, "no packed kernel for K=%d", K);
}

int mm16g_launch(int mode, int K, bool border, cudaStream_t st, const Mm16Params &p, int &launched, WorkItems &W, const int *h_list) {
  return mode == ABV_GLOCAL ? mm16g_launch_mode<GLOCAL>(K, border, st, p, launched, W, h_list) : mm16g_launch_mode<GLOBAL>(K, border, st, p, launched, W, h_list);
}

template <int MODE>
int mm16_launch_mode(int K, cudaStream_t st, const Mm16Params &p, int &launched, WorkItems &W, const int *h_list) {
  switch (K) {
    case 2: return mm16_launch_one<MODE, 2>(0, st, p, launched, W, h_list);
    case 4: return mm16_launch_one<MODE, 4>(0, st, p, launched, W, h_list);
    case 6: return mm16_launch_one<MODE, 6>(0, st, p, launched, W, h_list);
    case 8: return mm16_launch_one<MODE, 8>(0, st, p, launched, W, h_list);
    case 10: return mm16_launch_one<MODE, 10>(0, st, p, launched, W, h_list);
    case 12: return mm16_launch_one<MODE, 12>(0, st, p, launched, W, h_list);
    case 16: return mm16_launch_one<MODE, 16>(0, st, p, launched, W, h_list);
  }
  return fail(ABV_ERR_ARGS, "no packed kernel for K=%d", K);
}

int mm16_launch(int mode, int K, bool streaming, bool border, cudaStream_t st, const Mm16Params &p, int &launched, WorkItems &W, const int *h_list) {
  if ((mode == ABV_GLOBAL || mode == ABV_GLOCAL) && streaming) return mm16g_launch(mode, K, border, st, p, launched, W, h_list);
  if (mode == ABV_GLOBAL) return mm16_launch_mode<GLOBAL>(K, st, p, launched, W, h_list);
  if (mode == ABV_LOCAL) return mm16_launch_mode<LOCAL>(K, st, p, launched, W, h_list);
  if (mode == ABV_GLOCAL) return mm16_launch_mode<GLOCAL>(K, st, p, launched, W, h_list);
  if (mode == ABV_OVERLAP) return mm16_launch_mode<OVERLAP>(K, st, p, launched, W, h_list);
  return fail(ABV_ERR_ARGS, "packed kernel does not implement mode %d", mode);
}

}  // namespace

// =================================================================================================
// the many-to-many job
struct abv_job {
  int device = 0;                  // the device context the job lives on
  int mode = 0, n_q = 0, n_t = 0, seed = 0;
  int alpha = 0, go = 0, ge = 0;
  double cells = 0;
  double h2d_bytes = 0;
  bool want_all_scores = false;
  // device data
  DevMem q_codes, q_off, t_codes, t_off, map, matrix, matrix_ext, tp_codes, tp_meta, lists, counters;
  DevMem best_score, best_idx, all_scores, chunk_scores, border, best_key;
  int tp_stride = 0, n_pairs = 0, a1 = 0, max_lt = 0, max_lq = 0, maxabs = 1, maxpos = 0;
  int n_pairs_stream = 0;          // leading target pairs (longest first) the streaming kernel takes; the rest goes to mm16_kernel
  // per-launch timing: event pairs of the last PROFILE_RING launches of every bucket, so that a caller can time many
  // launches back to back and read them afterwards (no host synchronisation per launch)
  static constexpr int PROFILE_RING = 32;
  struct Bucket {
    int K, n, offset; double cells; std::vector<cudaEvent_t> ev; bool streaming;
    bool border;                    // streaming global: the border-lane instantiation (queries of up to 31K columns)
    std::shared_ptr<WorkItems> items, items_short;      // work-item tables of its launch(es), built at the first launch
  };
  std::vector<int> h_lists;        // host copy of `lists`
  std::vector<Bucket> buckets;     // packed-kernel launches (one per K class), lists at `offset`; largest first
  // concurrent launches: bucket j runs on side[j] between the fork event (recorded on the caller's stream after the counter
  // memsets) and its join event (which the caller's stream waits for before the finalize kernel).  The launch of bucket j > 0
  // additionally waits (cuStreamWaitValue32) for the tail flag of bucket j-1: the word its first CTA without a work item
  // sets.  So the grids run one after the other, largest first, each starting exactly when the one before begins to drain
  // and filling the SM slots that grid's CTAs give up; every launch is still timed alone by its event pair.  (Stream
  // priorities alone did this on one GPU but not under torchrun with NCCL initialised: the grids then shared the SMs from
  // the start -- same throughput, no clean per-launch time.)
  std::vector<cudaStream_t> side;
  cudaEvent_t ev_fork = nullptr;
  std::vector<cudaEvent_t> ev_join;
  long long n_launches = 0;        // abv_best_hits_launch calls so far
  ~abv_job() {
    for (auto &b : buckets) for (auto e : b.ev) if (e) cudaEventDestroy(e);
    for (auto s : side) if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }   // (a failed launch may have left forked work unjoined)
    for (auto e : ev_join) if (e) cudaEventDestroy(e);
    if (ev_fork) cudaEventDestroy(ev_fork);
  }
  int n_wf32 = 0, wf32_offset = 0; // queries served by the 32-bit kernel, list at `wf32_offset`
  int wf32_blocks = 0;             // its grid, fixed when `border` was sized (the option may change before the launch)
  int n_counters = 0;
  int launches_per_run = 0;
  cudaStream_t last_stream = nullptr;
};

namespace {

// is the packed kernel exact for queries padded to P columns against these targets?
bool packed_exact(int mode, int P, int max_lt, int maxabs) {
  const long long bound = (long long)(P + max_lt + 8) * maxabs;
  return bound <= ((mode == ABV_GLOCAL || mode == ABV_OVERLAP) ? p16::P16_LIMIT : 30000);      // these two keep half the range for their minus-infinity sentinel
}

int mm_validate(int mode, const char *q, const long long *q_off, int n_q, const char *t, const long long *t_off, int n_t,
                const Scoring &sc, int &max_lq, int &max_lt, int &min_lt) {
  int rc, min_lq;
  if ((rc = check_scoring(sc))) return rc;
  if (mode < 0 || mode > 3) return fail(ABV_ERR_ARGS, "unknown mode %d", mode);
  if ((rc = check_seqs(q, q_off, n_q, "query", max_lq, min_lq))) return rc;
  if ((rc = check_seqs(t, t_off, n_t, "target", max_lt, min_lt))) return rc;
  if (n_q > 0 && n_t > 0 && (rc = check_mode_lengths(mode, min_lq, min_lt))) return rc;
  if ((rc = check_symbols(q, q_off, n_q, sc, "query"))) return rc;
  if ((rc = check_symbols(t, t_off, n_t, sc, "target"))) return rc;
  return ABV_OK;
}

// drift-normalised values carry an extra |ge| per column and row (dp_core.cuh, mm16d_row); glocal
// keeps half the range free for its minus-infinity sentinel
bool packed_exact_streaming(int mode, int P, int max_lt, int maxabs, int ge) {
  if (mode != ABV_GLOBAL && mode != ABV_GLOCAL) return false;
  return (long long)(P + max_lt + 12) * (maxabs + std::abs(ge)) <= (mode == ABV_GLOCAL ? p16::P16_LIMIT : b16::LIMIT);
}

int job_prepare(abv_job &J, int mode, const char *q, const long long *q_off, int n_q,
                const char *t, const long long *t_off, int n_t, const Scoring &sc, int seed, bool all_scores) {
  int rc, max_lq, max_lt, min_lt = 0;
  if ((rc = mm_validate(mode, q, q_off, n_q, t, t_off, n_t, sc, max_lq, max_lt, min_lt))) return rc;
  if ((rc = ctx_init())) return rc;
  cudaStream_t st = g.stream;
  J.device = g.device;
  J.mode = mode; J.n_q = n_q; J.n_t = n_t; J.seed = seed; J.alpha = sc.alpha; J.go = sc.go; J.ge = sc.ge;
  J.max_lq = max_lq; J.max_lt = max_lt; J.want_all_scores = all_scores;
  {
    double tq = n_q ? (double)(q_off[n_q] - q_off[0]) : 0, tt = n_t ? (double)(t_off[n_t] - t_off[0]) : 0;
    J.cells = tq * tt;
  }
  CU(J.best_score.alloc(sizeof(int) * (size_t)std::max(n_q, 1)));
  CU(J.best_idx.alloc(sizeof(int) * (size_t)std::max(n_q, 1)));
  CU(J.best_key.alloc(sizeof(unsigned long long) * (size_t)std::max(n_q, 1)));
  if (all_scores) {
    if ((long long)n_q * n_t > (1ll << 28)) return fail(ABV_ERR_ARGS, "score matrix of %d x %d is too large", n_q, n_t);
    CU(J.all_scores.alloc(sizeof(int) * (size_t)std::max<long long>((long long)n_q * n_t, 1)));
  }
  if (n_q == 0) return ABV_OK;
  if ((rc = upload_encoded(q, q_off, n_q, sc, st, J.q_codes, J.q_off, J.map, J.h2d_bytes))) return rc;
  if ((rc = upload_encoded(t, t_off, n_t, sc, st, J.t_codes, J.t_off, J.map, J.h2d_bytes))) return rc;
  CU(J.matrix.alloc(sizeof(int) * sc.alpha * sc.alpha));
  CU(cudaMemcpyAsync(J.matrix.p, sc.matrix, sizeof(int) * sc.alpha * sc.alpha, cudaMemcpyHostToDevice, st));
  J.h2d_bytes += sizeof(int) * sc.alpha * sc.alpha;

  // ---- which queries can take the packed kernel ------------------------------------------------
  int maxabs = std::max(std::abs(sc.go), std::abs(sc.ge));
  for (int i = 0; i < sc.alpha * sc.alpha; ++i) maxabs = std::max(maxabs, std::abs(sc.matrix[i]));
  maxabs = std::max(maxabs, 1);
  int maxpos = 0;
  for (int i = 0; i < sc.alpha * sc.alpha; ++i) maxpos = std::max(maxpos, sc.matrix[i]);
  J.maxabs = maxabs; J.maxpos = maxpos;
  const int a1 = sc.alpha + 1;
  // overlap: the packed form is exact where re-opening a gap state from itself cannot win, go <= ge (dp_core.cuh, mm16_overlap_cands)
  bool packed_ok = g_opt.path != ABV_PATH_WF32 && n_t > 0 && a1 * a1 <= 64 &&
                   (mode == ABV_GLOBAL || mode == ABV_LOCAL || mode == ABV_GLOCAL || (mode == ABV_OVERLAP && sc.go <= sc.ge));
  // the streaming kernel sweeps at least 32 rows per pair, plus one pad row after every pair (its separator row, mm16g_levels)
  const int tp_stride = (std::max(max_lt, MM16G_MIN_ROWS) + 1 + 15) & ~15;
  // glocal streams only target pairs whose members have >= 32 rows (deferred last rows, kernels.cuh): the targets are paired
  // longest first, so those are the leading pairs; the few short ones at the end take the first-generation kernel with the
  // next even strip width, and both kernels fold into the same per-query best key
  const int n_pairs_all = (n_t + 1) / 2;
  int n_pairs_stream = n_pairs_all;
  if (mode == ABV_GLOCAL) {
    int n_long = 0;
    for (int j = 0; j < n_t; ++j) n_long += (t_off[j + 1] - t_off[j]) >= MM16G_MIN_ROWS;
    n_pairs_stream = n_long / 2 + (((n_long & 1) && n_long == n_t) ? 1 : 0);
  }
  J.n_pairs_stream = n_pairs_stream;
  auto stream_ok = [&](int K) {
    if (g_opt.global_v1 || n_pairs_stream == 0) return false;
    if (!packed_exact_streaming(mode, 32 * K, tp_stride, maxabs, sc.ge)) return false;
    if (n_pairs_stream < n_pairs_all) {
      const int K2 = K + (K & 1);
      if (!packed_exact(mode, 32 * K2, tp_stride, maxabs) || mm16_smem_layout(K2, a1, tp_stride).total > 100 * 1024) return false;
    }
    return true;
  };
  // lists[2k]: the bucket of strip width K; lists[2k + 1]: streaming global, its queries that fit the border-lane instantiation
  // (up to 31K columns; the border lane's fixed point needs Ec <= Hc, i.e. go <= ge: dp_core.cuh, mm16g_border)
  std::vector<std::vector<int>> lists(2 * MM16_NK);
  std::vector<int> wf32_list;
  bool st_okv[MM16_NK], border_okv[MM16_NK];
  for (int k = 0; k < MM16_NK; ++k) {
    st_okv[k] = stream_ok(MM16_KS[k]);
    border_okv[k] = st_okv[k] && mm16g_border(MM16_KS[k], mode == ABV_GLOCAL) && sc.go <= sc.ge && !g_opt.no_border;
  }
  for (int i = 0; i < n_q; ++i) {
    const int lq = (int)(q_off[i + 1] - q_off[i]);
    int slot = -1;
    if (packed_ok) {
      for (int k = 0; k < MM16_NK; ++k) {
        const int K = MM16_KS[k];
        if (K % 2 && !st_okv[k]) continue;
        if (32 * K >= lq) { slot = k; break; }
      }
      if (slot >= 0) {
        const int K = MM16_KS[slot];
        if (!packed_exact(mode, 32 * K, tp_stride, maxabs)) slot = -1;
        else if (std::max(mm16_smem_layout(K, a1, tp_stride).total, mm16g_smem_layout(K, a1, tp_stride, mm16g_defer(mode == ABV_GLOCAL, K)).total) > 100 * 1024) slot = -1;
      }
    }
    if (slot >= 0) lists[2 * slot + ((border_okv[slot] && 31 * MM16_KS[slot] >= lq) ? 1 : 0)].push_back(i); else wf32_list.push_back(i);
  }
  if (g_opt.path == ABV_PATH_PACKED && !wf32_list.empty())
    return fail(ABV_ERR_ARGS, "packed path forced but %zu queries are not eligible for it", wf32_list.size());

  std::vector<int> flat;
  for (int k2 = 0; k2 < 2 * MM16_NK; ++k2)
    if (!lists[k2].empty()) {
      const int k = k2 / 2;
      double qlen = 0;
      for (int i : lists[k2]) qlen += (double)(q_off[i + 1] - q_off[i]);
      abv_job::Bucket b{MM16_KS[k], (int)lists[k2].size(), (int)flat.size(), qlen * (double)(t_off[n_t] - t_off[0]), {},
                        st_okv[k], (k2 & 1) != 0, std::make_shared<WorkItems>(), std::make_shared<WorkItems>()};
      b.ev.assign(2 * abv_job::PROFILE_RING, nullptr);
      for (auto &e : b.ev) CU(cudaEventCreate(&e));
      J.buckets.push_back(b);
      flat.insert(flat.end(), lists[k2].begin(), lists[k2].end());
    }
  // largest launch first: it is the one whose tail the others fill, and the one bench.py reports the roofline of
  std::stable_sort(J.buckets.begin(), J.buckets.end(), [](const abv_job::Bucket &a, const abv_job::Bucket &b) { return a.cells > b.cells; });
  if (J.buckets.size() > 1) {
    CU(cudaEventCreateWithFlags(&J.ev_fork, cudaEventDisableTiming));
    J.side.assign(J.buckets.size(), nullptr);
    J.ev_join.assign(J.buckets.size(), nullptr);
    int prio_least = 0, prio_greatest = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    for (size_t j = 0; j < J.side.size(); ++j)
      CU(cudaStreamCreateWithPriority(&J.side[j], cudaStreamNonBlocking, j == 0 ? prio_greatest : prio_least));
    for (auto &e : J.ev_join) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  }
  J.wf32_offset = (int)flat.size();
  J.n_wf32 = (int)wf32_list.size();
  flat.insert(flat.end(), wf32_list.begin(), wf32_list.end());
  J.h_lists = flat;
  CU(J.lists.alloc(sizeof(int) * std::max<size_t>(flat.size(), 1)));
  CU(cudaMemcpyAsync(J.lists.p, flat.data(), sizeof(int) * flat.size(), cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  J.h2d_bytes += sizeof(int) * flat.size();
  J.n_counters = 2 * (int)J.buckets.size() + 64 + 16;      // the last 16 words (32 x 32 bits): tail flags of the packed launches
  CU(J.counters.alloc(sizeof(unsigned long long) * J.n_counters));

  // ---- target pairs for the packed kernel: sort by length, pair neighbours ----------------------
  if (!J.buckets.empty()) {
    std::vector<int> order(n_t);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
      return (t_off[x + 1] - t_off[x]) > (t_off[y + 1] - t_off[y]);
    });
    const int n_pairs = (n_t + 1) / 2;
    unsigned char ident[256];
    const unsigned char *m = sc.map256;
    if (!m) { for (int i = 0; i < 256; ++i) ident[i] = (unsigned char)i; m = ident; }
    std::vector<unsigned char> codes((size_t)n_pairs * tp_stride);
    std::vector<int> meta((size_t)n_pairs * 4);
    const int pad = sc.alpha;
    for (int pi = 0; pi < n_pairs; ++pi) {
      const int i1 = order[2 * pi], i2 = (2 * pi + 1 < n_t) ? order[2 * pi + 1] : -1;
      const int l1 = (int)(t_off[i1 + 1] - t_off[i1]), l2 = i2 >= 0 ? (int)(t_off[i2 + 1] - t_off[i2]) : 0;
      const unsigned char *s1 = (const unsigned char *)t + t_off[i1];
      const unsigned char *s2 = i2 >= 0 ? (const unsigned char *)t + t_off[i2] : nullptr;
      unsigned char *row = codes.data() + (size_t)pi * tp_stride;
      for (int r = 0; r < tp_stride; ++r) {
        const int a = r < l1 ? std::min<int>(m[s1[r]], sc.alpha - 1) : pad;
        const int b = r < l2 ? std::min<int>(m[s2[r]], sc.alpha - 1) : pad;
        row[r] = (unsigned char)(a * a1 + b);
      }
      meta[4 * pi] = l1; meta[4 * pi + 1] = l2; meta[4 * pi + 2] = i1; meta[4 * pi + 3] = i2;
    }
    std::vector<int> mext((size_t)a1 * a1, 0);
    for (int x = 0; x < sc.alpha; ++x)
      for (int y = 0; y < sc.alpha; ++y) mext[x * a1 + y] = sc.matrix[x * sc.alpha + y];
    CU(J.tp_codes.alloc(codes.size()));
    CU(J.tp_meta.alloc(sizeof(int) * meta.size()));
    CU(J.matrix_ext.alloc(sizeof(int) * mext.size()));
    CU(cudaMemcpyAsync(J.tp_codes.p, codes.data(), codes.size(), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(J.tp_meta.p, meta.data(), sizeof(int) * meta.size(), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(J.matrix_ext.p, mext.data(), sizeof(int) * mext.size(), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    J.h2d_bytes += codes.size() + sizeof(int) * (meta.size() + mext.size());
    J.tp_stride = tp_stride; J.n_pairs = n_pairs; J.a1 = a1;
  }

  // ---- scratch of the 32-bit path ----------------------------------------------------------------
  if (J.n_wf32 > 0 && n_t > 0) {
    const long long max_chunk_scores = 32ll << 20;
    long long rows = std::max<long long>(1, std::min<long long>(J.n_wf32, max_chunk_scores / n_t));
    if (!all_scores) CU(J.chunk_scores.alloc(sizeof(int) * (size_t)(rows * n_t)));
    J.wf32_blocks = wf32_grid(rows * n_t, 0);
    CU(J.border.alloc(sizeof(int) * 2 * (size_t)max_lt * J.wf32_blocks * WF32_WARPS));
  }
  CU(cudaStreamSynchronize(st));
  // launches per run: counter memset is not a kernel
  J.launches_per_run = (int)J.buckets.size();
  if (J.n_pairs_stream < J.n_pairs)
    for (const auto &b : J.buckets) J.launches_per_run += b.streaming ? 1 : 0;      // the short target pairs: a second launch
  if (J.n_wf32 > 0 && n_t > 0) {
    const long long max_chunk_scores = 32ll << 20;
    long long rows = std::max<long long>(1, std::min<long long>(J.n_wf32, max_chunk_scores / n_t));
    long long chunks = (J.n_wf32 + rows - 1) / rows;
    J.launches_per_run += (int)chunks * 2;
  }
  return ABV_OK;
}

int job_launch(abv_job &J, cudaStream_t st, int *d_best_score, int *d_best_idx) {
  int rc;
  if ((rc = ctx_init())) return rc;
  if (!st) st = g.stream;
  J.last_stream = st;
  if (!d_best_score) d_best_score = J.best_score.as<int>();
  if (!d_best_idx) d_best_idx = J.best_idx.as<int>();
  if (J.n_q == 0) return ABV_OK;
  if (J.n_t == 0) {   // nothing can beat the seed
    std::vector<int> s(J.n_q, J.seed), i(J.n_q, -1);
    CU(cudaMemcpyAsync(d_best_score, s.data(), sizeof(int) * J.n_q, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_best_idx, i.data(), sizeof(int) * J.n_q, cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));
    return ABV_OK;
  }
  CU(cudaMemsetAsync(J.counters.p, 0, sizeof(unsigned long long) * J.n_counters, st));
  if (!J.buckets.empty()) CU(cudaMemsetAsync(J.best_key.p, 0, sizeof(unsigned long long) * (size_t)J.n_q, st));
  int launched = 0;
  int ci = 0;
  const bool fork = g_opt.concurrent_buckets && J.buckets.size() > 1 && stream_wait_value32() != nullptr;
  unsigned int *flags = reinterpret_cast<unsigned int *>(J.counters.as<unsigned long long>() + J.n_counters - 16);   // zeroed with the counters
  if (fork) CU(cudaEventRecord(J.ev_fork, st));
  cudaStream_t caller = st;
  for (size_t bj = 0; bj < J.buckets.size(); ++bj) {
    const auto &b = J.buckets[bj];
    st = fork ? J.side[bj] : caller;
    if (fork) {
      CU(cudaStreamWaitEvent(st, J.ev_fork, 0));
      if (bj > 0 && stream_wait_value32()((CUstream)st, (CUdeviceptr)(uintptr_t)(flags + (bj - 1)), 1u, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
        return fail(ABV_ERR_CUDA, "cuStreamWaitValue32 failed");
    }
    Mm16Params p{};
    p.q = J.q_codes.as<uint8_t>(); p.q_off = J.q_off.as<long long>();
    p.q_list = J.lists.as<int>() + b.offset; p.n_list = b.n;
    p.tp_codes = J.tp_codes.as<uint8_t>(); p.tp_stride = J.tp_stride;
    p.tp_meta = J.tp_meta.as<int4>(); p.n_pairs = J.n_pairs;
    p.matrix_ext = J.matrix_ext.as<int>(); p.a1 = J.a1;
    p.go = J.go; p.ge = J.ge; p.seed = J.seed;
    p.best_score = d_best_score; p.best_idx = d_best_idx;
    p.all_scores = J.want_all_scores ? J.all_scores.as<int>() : nullptr; p.n_t = J.n_t;
    p.q_counter = reinterpret_cast<unsigned int *>(J.counters.as<unsigned long long>() + ci++);
    p.best_key = J.best_key.as<unsigned long long>();
    p.tail_flag = (fork && bj + 1 < J.buckets.size() && bj < 32) ? flags + bj : nullptr;
    {
      const Mm16gLevels lv = (b.streaming && J.mode == ABV_GLOBAL && !g_opt.no_levels)
                                 ? mm16g_levels(32 * b.K, J.tp_stride, J.go, J.ge, J.maxabs, J.maxpos) : Mm16gLevels{1, 0, 0};
      p.lv_n = lv.n; p.lv0 = lv.lvl0; p.lv_step = lv.step;
    }
    const int slot = (int)(J.n_launches % abv_job::PROFILE_RING);
    CU(cudaEventRecord(b.ev[2 * slot], st));
    if (b.streaming && J.n_pairs_stream < J.n_pairs) {
      Mm16Params r = p;                                  // the short target pairs at the end of the list
      p.tail_flag = nullptr;                             // ... whose launch is the bucket's last one
      r.tp_codes += (size_t)J.n_pairs_stream * J.tp_stride; r.tp_meta += J.n_pairs_stream; r.n_pairs = J.n_pairs - J.n_pairs_stream;
      r.q_counter = reinterpret_cast<unsigned int *>(J.counters.as<unsigned long long>() + ci++);
      p.n_pairs = J.n_pairs_stream;
      if ((rc = mm16_launch(J.mode, b.K, true, b.border, st, p, launched, *b.items, J.h_lists.data() + b.offset))) return rc;
      if ((rc = mm16_launch(J.mode, b.K + (b.K & 1), false, false, st, r, launched, *b.items_short, J.h_lists.data() + b.offset))) return rc;
    } else if ((rc = mm16_launch(J.mode, b.K, b.streaming, b.border, st, p, launched, *b.items, J.h_lists.data() + b.offset))) return rc;
    CU(cudaEventRecord(b.ev[2 * slot + 1], st));
    if (fork) {
      CU(cudaEventRecord(J.ev_join[bj], st));
      CU(cudaStreamWaitEvent(caller, J.ev_join[bj], 0));
    }
  }
  st = caller;
  if (!J.buckets.empty()) {   // seed rule + unpack, over the queries the packed kernels served
    const int n_packed = J.wf32_offset;
    best_key_finalize_kernel<<<(n_packed + 255) / 256, 256, 0, st>>>(J.best_key.as<unsigned long long>(), J.lists.as<int>(), n_packed,
                                                                     J.seed, d_best_score, d_best_idx);
    CU(cudaGetLastError());
    ++launched;
  }
  if (J.n_wf32 > 0) {
    const long long max_chunk_scores = 32ll << 20;
    const long long rows_per = std::max<long long>(1, std::min<long long>(J.n_wf32, max_chunk_scores / J.n_t));
    const int grid = J.wf32_blocks;        // the grid `border` was sized for in job_prepare
    int chunk = 0;
    for (long long base = 0; base < J.n_wf32; base += rows_per, ++chunk) {
      const long long rows = std::min<long long>(rows_per, J.n_wf32 - base);
      Wf32Params p{};
      p.a = J.q_codes.as<uint8_t>(); p.a_off = J.q_off.as<long long>();
      p.b = J.t_codes.as<uint8_t>(); p.b_off = J.t_off.as<long long>();
      p.n_pairs = rows * J.n_t; p.n_t = J.n_t; p.q_base = base;
      p.q_list = J.lists.as<int>() + J.wf32_offset;
      p.matrix = J.matrix.as<int>(); p.alpha = J.alpha; p.go = J.go; p.ge = J.ge;
      p.border = J.border.as<int>(); p.border_len = J.max_lt;
      if (ci >= J.n_counters - 16) {   // counters exhausted: reset and reuse (launches are stream-ordered); the last 16 words are the tail flags
        CU(cudaMemsetAsync(J.counters.p, 0, sizeof(unsigned long long) * (J.n_counters - 16), st));
        ci = 0;
      }
      p.pair_counter = J.counters.as<unsigned long long>() + ci++;
      // full-matrix requests write scores[query*n_t + target]; otherwise a dense chunk in row order
      p.scores = J.want_all_scores ? J.all_scores.as<int>() : J.chunk_scores.as<int>();
      p.score_rows_by_query = J.want_all_scores ? 1 : 0;
      if ((rc = wf32_dispatch<false>(J.mode, (int)std::min<long long>(grid, (p.n_pairs + WF32_WARPS - 1) / WF32_WARPS), st, p, J.max_lt))) return rc;
      ++launched;
      const int wpb = 8;
      argmax_rows_kernel<<<(int)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(
          p.scores, (int)rows, J.n_t, J.seed, d_best_score, d_best_idx, base,
          J.lists.as<int>() + J.wf32_offset, J.want_all_scores ? 1 : 0);
      CU(cudaGetLastError());
      ++launched;
    }
  }
  J.launches_per_run = launched;
  ++J.n_launches;
  return ABV_OK;
}

// ---- tb16 launcher: packed one-to-one kernel with traceback (local, global) ------------------------
const int TB16_KS[] = {2, 4, 5, 6, 7, 8, 10};
constexpr int TB16_NK = 7;
constexpr int TB16_MAX_ALPHA = 7;

template <int MODE, int K>
int tb16_launch_one(cudaStream_t st, Tb16Params p, size_t dir_bytes_total, size_t tmp_entries_total) {
  auto kern = tb16_kernel<MODE, K>;
  const int smem = ((p.a1 * p.a1 * 4 + 15) & ~15) + TB16_WARPS * tb16_smem_per_warp<K>(p.a1, p.rows_cap) + g_opt.tb16_extra_smem;
  if (smem > 200 * 1024) return fail(ABV_ERR_ARGS, "tb16: %d bytes of shared memory", smem);
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int per_sm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TB16_WARPS * 32, smem));
  if (per_sm < 1) return fail(ABV_ERR_CUDA, "tb16 kernel does not fit an SM");
  p.dir_words_per_warp = (long long)K * TB16_PLANES * tb16_steps16(p.rows_cap) * 32;
  long long grid = std::min<long long>(((long long)p.n_items + TB16_WARPS - 1) / TB16_WARPS, (long long)g.sms * per_sm);
  grid = std::min<long long>(grid, (long long)(dir_bytes_total / ((size_t)p.dir_words_per_warp * 4 * TB16_WARPS)));
  grid = std::min<long long>(grid, (long long)(tmp_entries_total / ((size_t)p.frag_tmp_cap * TB16_WARPS)));
  if (grid < 1) return fail(ABV_ERR_NOMEM, "tb16 scratch too small");
  kern<<<(int)grid, TB16_WARPS * 32, smem, st>>>(p);
  CU(cudaGetLastError());
  return ABV_OK;
}
template <int MODE>
int tb16_launch_mode(int K, cudaStream_t st, const Tb16Params &p, size_t dir_bytes, size_t tmp_entries) {
  switch (K) {
    case 2: return tb16_launch_one<MODE, 2>(st, p, dir_bytes, tmp_entries);
    case 4: return tb16_launch_one<MODE, 4>(st, p, dir_bytes, tmp_entries);
    case 5: return tb16_launch_one<MODE, 5>(st, p, dir_bytes, tmp_entries);
    case 6: return tb16_launch_one<MODE, 6>(st, p, dir_bytes, tmp_entries);
    case 7: return tb16_launch_one<MODE, 7>(st, p, dir_bytes, tmp_entries);
    case 8: return tb16_launch_one<MODE, 8>(st, p, dir_bytes, tmp_entries);
    default: return tb16_launch_one<MODE, 10>(st, p, dir_bytes, tmp_entries);
  }
}

// Pairs the packed kernel can take, bucketed by columns per lane and sorted by target length so that the two
// pairs of a warp have similar shapes.  list = [bucket 0 items: (i0, i1) ...][bucket 1 ...]...[pairs for wf32]
struct Tb16Plan {
  std::vector<int> list, key, mext;
  struct Bucket { int K, offset, n_items, rows_cap, max_la, max_lb; };
  std::vector<Bucket> buckets;
  int rest_offset = 0, n_rest = 0, rest_max_la = 0, rest_max_lb = 0;
  int pad_score = -1;
  double cells = 0;          // sum of la * lb over the pairs (the pass over the offsets is made anyway)
};
bool tb16_eligible(int mode, const Scoring &sc) {
  if (!g_opt.tb16 || g_opt.path == ABV_PATH_WF32) return false;
  if (mode != ABV_LOCAL && mode != ABV_GLOBAL) return false;
  if (sc.alpha > TB16_MAX_ALPHA) return false;
  if (mode == ABV_LOCAL && sc.go >= 0) return false;      // the padding argument needs a strictly negative gap_open
  return true;
}
void tb16_plan(Tb16Plan &P, int mode, const long long *a_off, const long long *b_off, int n, const Scoring &sc) {
  int maxabs = 1;
  for (int i = 0; i < sc.alpha * sc.alpha; ++i) maxabs = std::max(maxabs, std::abs(sc.matrix[i]));
  P.pad_score = -(maxabs + 1);
  const int unit = std::max(std::max(-sc.go, -sc.ge), maxabs + 1);
  const bool ok = tb16_eligible(mode, sc);
  // counting sort by (bucket, lb), two passes over the pairs; the sorted order is written straight into the list
  constexpr int NKEY = (TB16_NK + 1) * (TB16_MAX_ROWS + 1);
  P.key.resize((size_t)n);
  std::vector<int> hist((size_t)NKEY + 1, 0);
  int max_la[TB16_NK + 1] = {0};
  int la_limit[TB16_NK], lb_limit[TB16_NK];          // per bucket: longest query, longest target that keeps values in range
  for (int b = 0; b < TB16_NK; ++b) {
    la_limit[b] = tb16_max_la(mode == ABV_LOCAL, TB16_KS[b]);      // local keeps lane 31 free (dp_core.cuh)
    lb_limit[b] = (int)std::min<long long>(TB16_MAX_ROWS, (b16::LIMIT + 2ll * sc.go) / unit - 32ll * TB16_KS[b] - 2);
  }
  P.rest_max_la = P.rest_max_lb = 0;
  double cells = 0;
  for (int i = 0; i < n; ++i) {
    const int la = (int)std::min<long long>(a_off[i + 1] - a_off[i], INT_MAX), lb = (int)std::min<long long>(b_off[i + 1] - b_off[i], INT_MAX);
    cells += (double)la * (double)lb;
    int b = TB16_NK;
    if (ok && la <= la_limit[TB16_NK - 1]) {
      b = 0;
      while (la_limit[b] < la) ++b;
      if (lb > lb_limit[b]) b = TB16_NK;
    }
    if (b == TB16_NK) { P.rest_max_la = std::max(P.rest_max_la, la); P.rest_max_lb = std::max(P.rest_max_lb, lb); }
    else max_la[b] = std::max(max_la[b], la);
    const int k = b * (TB16_MAX_ROWS + 1) + (b == TB16_NK ? 0 : lb);
    P.key[i] = k;
    ++hist[k + 1];
  }
  P.cells = cells;
  for (int k = 1; k <= NKEY; ++k) hist[k] += hist[k - 1];
  // list position of the first pair of every bucket (an odd bucket gets one -1 partner at its end)
  P.buckets.clear();
  int shift[TB16_NK + 1];
  int extra = 0;
  for (int b = 0; b < TB16_NK; ++b) {
    shift[b] = extra;
    const int lo = hist[b * (TB16_MAX_ROWS + 1)], hi = hist[(b + 1) * (TB16_MAX_ROWS + 1)];
    if (hi == lo) continue;
    int max_lb = 0;
    for (int lb = TB16_MAX_ROWS; lb >= 0; --lb)
      if (hist[b * (TB16_MAX_ROWS + 1) + lb + 1] > hist[b * (TB16_MAX_ROWS + 1) + lb]) { max_lb = lb; break; }
    P.buckets.push_back(Tb16Plan::Bucket{TB16_KS[b], lo + extra, (hi - lo + 1) / 2, (max_lb + 15) & ~15, max_la[b], max_lb});
    extra += (hi - lo) & 1;
  }
  shift[TB16_NK] = extra;
  const int lo_rest = hist[TB16_NK * (TB16_MAX_ROWS + 1)];
  P.rest_offset = lo_rest + extra;
  P.n_rest = n - lo_rest;
  P.list.assign((size_t)n + extra, -1);
  for (int i = 0; i < n; ++i) {
    const int k = P.key[i];
    P.list[(size_t)(hist[k]++) + shift[k / (TB16_MAX_ROWS + 1)]] = i;
  }
}

}  // namespace

// =================================================================================================
extern "C" {

int abv_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

int abv_set_device(int ordinal) {
  if (ordinal < 0 || ordinal >= ABV_MAX_DEVICES) return fail(ABV_ERR_ARGS, "device ordinal %d out of range", ordinal);
  g_default_device = ordinal;          // the process default (threads that never chose a device use it) ...
  t_device = -1;                       // ... and this thread follows it
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  return ctx_init();
}

int abv_set_thread_device(int ordinal) {
  if (ordinal < -1 || ordinal >= ABV_MAX_DEVICES) return fail(ABV_ERR_ARGS, "device ordinal %d out of range", ordinal);
  t_device = ordinal;                  // -1: back to the process default
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  return ctx_init();
}

int abv_get_device(void) { return t_device >= 0 ? t_device : g_default_device; }
const char *abv_last_error(void) { return t_err.c_str(); }
const char *abv_version(void) { return "ambivert-b200 0.1 (sm_100a)"; }

int abv_set_option(const char *key, int value) {
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  if (!key) return fail(ABV_ERR_ARGS, "NULL key");
  if (!strcmp(key, "path")) { if (value < 0 || value > 2) return fail(ABV_ERR_ARGS, "bad path"); g_opt.path = value; return ABV_OK; }
  if (!strcmp(key, "wf32_blocks_per_sm")) { g_opt.wf32_blocks_per_sm = std::max(1, std::min(value, 8)); return ABV_OK; }
  if (!strcmp(key, "cache_mb")) {
    for (Ctx &c : g_ctx) {                      // per device
      std::lock_guard<std::recursive_mutex> lc(c.mu);
      c.cache.budget = (size_t)std::max(0, value) << 20;
      if (c.inited && cudaSetDevice(c.device) == cudaSuccess) c.cache.trim(c.cache.budget);
    }
    if (g.inited) cudaSetDevice(g.device);
    return ABV_OK;
  }
  if (!strcmp(key, "split")) { g_opt.split = std::max(0, std::min(value, 64)); return ABV_OK; }
  if (!strcmp(key, "fine_tail")) { g_opt.fine_tail = value ? 1 : 0; return ABV_OK; }
  if (!strcmp(key, "split_min_items")) { g_opt.split_min_items = value < 1 ? 1 : value; return ABV_OK; }
  if (!strcmp(key, "no_border")) { g_opt.no_border = value ? 1 : 0; return ABV_OK; }
  if (!strcmp(key, "concurrent_buckets")) { g_opt.concurrent_buckets = value ? 1 : 0; return ABV_OK; }
  if (!strcmp(key, "tb16")) { g_opt.tb16 = value ? 1 : 0; return ABV_OK; }
  if (!strcmp(key, "tb16_extra_smem")) { g_opt.tb16_extra_smem = value < 0 ? 0 : value; return ABV_OK; }
  if (!strcmp(key, "global_v1")) { g_opt.global_v1 = value ? 1 : 0; return ABV_OK; }
  if (!strcmp(key, "no_gop_imm")) { g_opt.no_gop_imm = value ? 1 : 0; return ABV_OK; }
  if (!strcmp(key, "no_levels")) { g_opt.no_levels = value ? 1 : 0; return ABV_OK; }
  if (!strcmp(key, "pairs_chunk")) { g_opt.pairs_chunk = std::max(0, value); return ABV_OK; }
  if (!strcmp(key, "mm16_blocks_per_sm")) { g_opt.mm16_blocks_per_sm = std::max(0, std::min(value, 8)); return ABV_OK; }
  return fail(ABV_ERR_ARGS, "unknown option %s", key);
}

abv_job *abv_best_hits_prepare(int mode, const char *q, const long long *q_off, int n_q,
                               const char *t, const long long *t_off, int n_t,
                               int alpha_len, const unsigned char *map256, const int *score_matrix,
                               int gap_open, int gap_extend, int seed_score) {
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  abv_job *J = new abv_job();
  Scoring sc{alpha_len, gap_open, gap_extend, map256, score_matrix};
  if (job_prepare(*J, mode, q, q_off, n_q, t, t_off, n_t, sc, seed_score, false) != ABV_OK) { delete J; return nullptr; }
  return J;
}

int abv_best_hits_launch(abv_job *job, void *stream, int *d_best_score, int *d_best_idx) {
  if (!job) return fail(ABV_ERR_ARGS, "NULL job");
  DeviceScope on(job->device);
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  return job_launch(*job, (cudaStream_t)stream, d_best_score, d_best_idx);
}

int abv_best_hits_fetch(abv_job *job, int *best_score, int *best_idx) {
  if (!job) return fail(ABV_ERR_ARGS, "NULL job");
  DeviceScope on(job->device);
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  int rc;
  if ((rc = ctx_init())) return rc;
  cudaStream_t st = job->last_stream ? job->last_stream : g.stream;
  if (job->n_q == 0) return ABV_OK;
  if (best_score) CU(cudaMemcpyAsync(best_score, job->best_score.p, sizeof(int) * job->n_q, cudaMemcpyDeviceToHost, st));
  if (best_idx) CU(cudaMemcpyAsync(best_idx, job->best_idx.p, sizeof(int) * job->n_q, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return ABV_OK;
}

double abv_job_cells(const abv_job *job) { return job ? job->cells : 0; }

int abv_job_profile(abv_job *job, double *out, int max_rows) {
  if (!job || !out) return 0;
  DeviceScope on(job->device);
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  if (ctx_init() != ABV_OK) return 0;
  int n = 0;
  if (job->n_launches == 0) return 0;
  const int slot = (int)((job->n_launches - 1) % abv_job::PROFILE_RING);
  for (const auto &b : job->buckets) {
    if (n >= max_rows) break;
    float ms = 0;
    if (cudaEventElapsedTime(&ms, b.ev[2 * slot], b.ev[2 * slot + 1]) != cudaSuccess) { cudaGetLastError(); ms = -1; }
    out[4 * n] = b.K + (b.border ? 0.5 : 0.0); out[4 * n + 1] = b.n; out[4 * n + 2] = b.cells; out[4 * n + 3] = ms;
    ++n;
  }
  return n;
}
int abv_job_profile_history(abv_job *job, int last_launches, double *out, int max_rows) {
  if (!job || !out || last_launches <= 0) return 0;
  DeviceScope on(job->device);
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  if (ctx_init() != ABV_OK) return 0;
  const long long have = std::min<long long>(std::min<long long>(job->n_launches, abv_job::PROFILE_RING), last_launches);
  int n = 0;
  for (long long l = job->n_launches - have; l < job->n_launches; ++l) {
    const int slot = (int)(l % abv_job::PROFILE_RING);
    for (const auto &b : job->buckets) {
      if (n >= max_rows) return n;
      float ms = 0;
      if (cudaEventElapsedTime(&ms, b.ev[2 * slot], b.ev[2 * slot + 1]) != cudaSuccess) { cudaGetLastError(); ms = -1; }
      out[5 * n] = b.K + (b.border ? 0.5 : 0.0); out[5 * n + 1] = b.n; out[5 * n + 2] = b.cells; out[5 * n + 3] = ms; out[5 * n + 4] = (double)l;
      ++n;
    }
  }
  return n;
}
int abv_job_kernel_launches(const abv_job *job) { return job ? job->launches_per_run : 0; }
void abv_job_free(abv_job *job) {
  if (!job) return;
  DeviceScope on(job->device);
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  ctx_init();
  // abv_best_hits_launch is asynchronous and may have run on a caller's stream: the job's blocks go back to the
  // allocation cache (and from there to the next job's uploads on the library's stream) only once those kernels are done
  if (job->last_stream && cudaStreamSynchronize(job->last_stream) != cudaSuccess) cudaGetLastError();
  delete job;
}

static int best_hits_impl(int mode, const char *q, const long long *q_off, int n_q,
                          const char *t, const long long *t_off, int n_t,
                          int alpha_len, const unsigned char *map256, const int *score_matrix,
                          int gap_open, int gap_extend, int seed_score,
                          int *best_score, int *best_idx, int *scores) {
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  int rc;
  Scoring sc{alpha_len, gap_open, gap_extend, map256, score_matrix};
  { int a, b, c; if ((rc = mm_validate(mode, q, q_off, n_q, t, t_off, n_t, sc, a, b, c))) return rc; }
  if ((rc = ctx_init())) return rc;
  for (double &v : t_timing) v = 0;
  abv_job J;
  CU(cudaEventRecord(g.ev[0], g.stream));
  if ((rc = job_prepare(J, mode, q, q_off, n_q, t, t_off, n_t, sc, seed_score, scores != nullptr))) return rc;
  CU(cudaEventRecord(g.ev[1], g.stream));
  if ((rc = job_launch(J, g.stream, nullptr, nullptr))) return rc;
  CU(cudaEventRecord(g.ev[2], g.stream));
  double d2h = 0;
  if (n_q > 0) {
    if (best_score) { CU(cudaMemcpyAsync(best_score, J.best_score.p, sizeof(int) * n_q, cudaMemcpyDeviceToHost, g.stream)); d2h += sizeof(int) * (double)n_q; }
    if (best_idx) { CU(cudaMemcpyAsync(best_idx, J.best_idx.p, sizeof(int) * n_q, cudaMemcpyDeviceToHost, g.stream)); d2h += sizeof(int) * (double)n_q; }
    if (scores && n_t > 0) {
      CU(cudaMemcpyAsync(scores, J.all_scores.p, sizeof(int) * (size_t)n_q * n_t, cudaMemcpyDeviceToHost, g.stream));
      d2h += sizeof(int) * (double)n_q * n_t;
    }
  }
  CU(cudaEventRecord(g.ev[3], g.stream));
  CU(cudaStreamSynchronize(g.stream));
  float a = 0, b = 0, c = 0;
  cudaEventElapsedTime(&a, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&b, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&c, g.ev[2], g.ev[3]);
  t_timing[0] = a; t_timing[1] = b; t_timing[2] = c; t_timing[3] = J.cells; t_timing[4] = J.launches_per_run;
  t_timing[5] = J.h2d_bytes; t_timing[6] = d2h; t_timing[7] = b;
  return ABV_OK;
}

int abv_best_hits(int mode, const char *q, const long long *q_off, int n_q,
                  const char *t, const long long *t_off, int n_t,
                  int alpha_len, const unsigned char *map256, const int *score_matrix,
                  int gap_open, int gap_extend, int seed_score, int *best_score, int *best_idx) {
  if (n_q > 0 && (!best_score || !best_idx)) return fail(ABV_ERR_ARGS, "NULL output");
  return best_hits_impl(mode, q, q_off, n_q, t, t_off, n_t, alpha_len, map256, score_matrix, gap_open, gap_extend,
                        seed_score, best_score, best_idx, nullptr);
}

int abv_score_matrix(int mode, const char *q, const long long *q_off, int n_q,
                     const char *t, const long long *t_off, int n_t,
                     int alpha_len, const unsigned char *map256, const int *score_matrix,
                     int gap_open, int gap_extend, int *scores) {
  if (n_q > 0 && n_t > 0 && !scores) return fail(ABV_ERR_ARGS, "NULL output");
  return best_hits_impl(mode, q, q_off, n_q, t, t_off, n_t, alpha_len, map256, score_matrix, gap_open, gap_extend,
                        INT_MIN, nullptr, nullptr, scores);
}

// -------------------------------------------------------------------------------------------------
// The hashtable build on several GPUs of one box from ONE process (the successor of the reference's --threads,
// ambivert.py:474-476, :1092-1095): the queries are cut into contiguous shards balanced by their summed length (cells are
// proportional to it), every device gets the whole target set, one host thread per device runs the single-device job on
// its shard, and each shard's 8 bytes per query go straight from its device into the caller's arrays.  The arg-max is
// per query, so the result does not depend on the sharding.
void abv_shard_bounds(const long long *q_off, int n_q, int n_shards, int *bounds) {
  bounds[0] = 0;
  const long long base = n_q > 0 ? q_off[0] : 0, total = n_q > 0 ? q_off[n_q] - base : 0;
  for (int r = 1; r < n_shards; ++r) {
    // first query index whose cumulative length reaches r/n of the total (numpy.searchsorted(csum, total*r/n, 'left'))
    const double want = (double)total * r / n_shards;
    int lo = 0, hi = n_q;
    while (lo < hi) {
      const int mid = (lo + hi) / 2;
      if ((double)(q_off[mid + 1] - base) < want) lo = mid + 1; else hi = mid;
    }
    bounds[r] = std::max(bounds[r - 1], std::min(lo, n_q));
  }
  bounds[n_shards] = n_q;
}

int abv_best_hits_multi(int mode, const char *q, const long long *q_off, int n_q,
                        const char *t, const long long *t_off, int n_t,
                        int alpha_len, const unsigned char *map256, const int *score_matrix,
                        int gap_open, int gap_extend, int seed_score,
                        const int *devices, int n_devices, int *best_score, int *best_idx) {
  if (n_devices < 1 || n_devices > ABV_MAX_DEVICES) return fail(ABV_ERR_ARGS, "n_devices %d out of range", n_devices);
  if (n_q > 0 && (!best_score || !best_idx)) return fail(ABV_ERR_ARGS, "NULL output");
  int rc;
  Scoring sc{alpha_len, gap_open, gap_extend, map256, score_matrix};
  { int a, b, c; if ((rc = mm_validate(mode, q, q_off, n_q, t, t_off, n_t, sc, a, b, c))) return rc; }
  std::vector<int> dev(n_devices), bounds(n_devices + 1);
  for (int i = 0; i < n_devices; ++i) {
    dev[i] = devices ? devices[i] : i;
    if (dev[i] < 0 || dev[i] >= ABV_MAX_DEVICES) return fail(ABV_ERR_ARGS, "device ordinal %d out of range", dev[i]);
  }
  if (n_q == 0) return ABV_OK;
  abv_shard_bounds(q_off, n_q, n_devices, bounds.data());
  std::vector<int> rcs(n_devices, ABV_OK);
  std::vector<std::string> errs(n_devices);
  std::vector<std::array<double, 8>> tms(n_devices);
  auto work = [&](int i) {
    t_device = dev[i];
    const int lo = bounds[i], hi = bounds[i + 1];
    rcs[i] = best_hits_impl(mode, q, q_off + lo, hi - lo, t, t_off, n_t, alpha_len, map256, score_matrix, gap_open, gap_extend,
                            seed_score, best_score + lo, best_idx + lo, nullptr);
    if (rcs[i] != ABV_OK) errs[i] = t_err;
    for (int k = 0; k < 8; ++k) tms[i][k] = t_timing[k];
  };
  std::vector<std::thread> th;
  for (int i = 1; i < n_devices; ++i) th.emplace_back(work, i);
  const int saved = t_device;
  work(0);                                   // the calling thread takes the first shard
  t_device = saved;
  for (auto &x : th) x.join();
  if (cur_ctx().inited) cudaSetDevice(cur_ctx().device);
  for (double &v : t_timing) v = 0;
  for (int i = 0; i < n_devices; ++i) {
    for (int k : {0, 1, 2, 7}) t_timing[k] = std::max(t_timing[k], tms[i][k]);      // the slowest device
    for (int k : {3, 4, 5, 6}) t_timing[k] += tms[i][k];                            // cells, launches, bytes: all devices
  }
  for (int i = 0; i < n_devices; ++i)
    if (rcs[i] != ABV_OK) return fail(rcs[i], "device %d: %s", dev[i], errs[i].c_str());
  return ABV_OK;
}

// -------------------------------------------------------------------------------------------------
// One-to-one alignments with traceback, results left on the device.
struct PairsRun {
  DevMem a_codes, a_off, b_codes, b_off, a_raw, b_raw, map, matrix;
  DevMem dir, tmp, border, scores, start, count, frags, counters, lists, matrix_ext;
  UploadKeep keep;
  long long pool = 0;
  unsigned long long total = 0;
  int max_la = 0, max_lb = 0;
  double h2d = 0, cells = 0;
  // a chunk of a pipelined job: its own stream, events and host-side plan (they live until the whole job has finished)
  cudaStream_t st = nullptr;
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  Tb16Plan *own_plan = nullptr;
  unsigned long long *pin_total = nullptr;   // pinned landing place of `total`
  ~PairsRun() {
    for (auto &e : ev) if (e) cudaEventDestroy(e);
    delete own_plan;
  }
};

static int pairs_validate(int mode, const char *a, const long long *a_off, const char *b, const long long *b_off, int n,
                          const Scoring &sc, int &max_la, int &max_lb, int base = 0) {
  int rc, min_la, min_lb;
  if ((rc = check_scoring(sc))) return rc;
  if ((rc = check_seqs(a, a_off, n, "sequence a", max_la, min_la, base))) return rc;
  if ((rc = check_seqs(b, b_off, n, "sequence b", max_lb, min_lb, base))) return rc;
  if (mode < 0 || mode > 3) return fail(ABV_ERR_ARGS, "unknown mode %d", mode);
  if (n > 0 && (rc = check_mode_lengths(mode, min_la, min_lb))) return rc;
  if ((rc = check_symbols(a, a_off, n, sc, "sequence a"))) return rc;
  if ((rc = check_symbols(b, b_off, n, sc, "sequence b"))) return rc;
  return ABV_OK;
}

static double trace_now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static bool trace_on() { static const bool on = getenv("ABV_TRACE") != nullptr; return on; }
#define ABV_TRACE(...) do { if (trace_on()) { fprintf(stderr, "[abv %.3f] ", trace_now()); fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); } } while (0)

// upload + encode + fill + traceback for n > 0 validated pairs; g.ev[0..2] bracket upload and kernel.
// A chunk of a pipelined job (R.st set) runs on its own stream with its own events and plan and returns WITHOUT waiting:
// the caller synchronises R.st before it reads R.total.
static int pairs_run(PairsRun &R, int mode, const char *a, const long long *a_off, const char *b, const long long *b_off, int n,
                     const Scoring &sc, long long pool_cap, bool keep_raw) {
  int rc;
  const bool chunk = R.st != nullptr;
  cudaStream_t st = chunk ? R.st : g.stream;
  cudaEvent_t *ev = chunk ? R.ev : g.ev;
  CU(cudaEventRecord(ev[0], st));
  // One piece: the sequences go first and copy while the host plans.  A chunk of a pipeline sends its small pageable pieces
  // first (offsets here, the plan below) and the sequence bytes last, see upload_encoded.
  auto upload = [&](int phase) -> int {
    int r2;
    if ((r2 = upload_encoded(a, a_off, n, sc, st, R.a_codes, R.a_off, R.map, R.h2d, keep_raw ? &R.a_raw : nullptr, &R.keep, phase))) return r2;
    if ((r2 = upload_encoded(b, b_off, n, sc, st, R.b_codes, R.b_off, R.map, R.h2d, keep_raw ? &R.b_raw : nullptr, &R.keep, phase))) return r2;
    return ABV_OK;
  };
  const bool first = !R.a_codes.p && !R.a_off.p;      // first attempt: the inputs are not on the device yet
  if (first) {
    if ((rc = upload(chunk ? 1 : 0))) return rc;
    CU(R.matrix.alloc(sizeof(int) * sc.alpha * sc.alpha));
    CU(cudaMemcpyAsync(R.matrix.p, sc.matrix, sizeof(int) * sc.alpha * sc.alpha, cudaMemcpyHostToDevice, st));
  }
  // plan: pairs the packed traceback kernel takes (bucketed, partnered) and the rest for the 32-bit kernel
  static thread_local Tb16Plan tl_plan;      // keeps its vectors' capacity from call to call
  Tb16Plan &plan = chunk ? *R.own_plan : tl_plan;
  ABV_TRACE("pairs_run n=%d: uploads issued", n);
  tb16_plan(plan, mode, a_off, b_off, n, sc);
  R.cells = plan.cells;
  ABV_TRACE("pairs_run n=%d: planned", n);
  const int tmp_cap = R.max_la + R.max_lb + 2;
  const long long dwords = plan.n_rest ? dir_words(plan.rest_max_la, plan.rest_max_lb) : 0;
  const size_t per_warp = (size_t)dwords * 4 + (size_t)tmp_cap * sizeof(FragOut) + (size_t)R.max_lb * 8;
  const int grid = plan.n_rest ? wf32_grid(plan.n_rest, per_warp) : 0;
  const long long warps = (long long)grid * WF32_WARPS;
  size_t dir_bytes = (size_t)dwords * 4 * warps, tmp_entries = (size_t)tmp_cap * warps;
  for (const auto &B : plan.buckets) {       // scratch for a full persistent grid of the packed kernel (32 warps per SM at most)
    const long long w = std::min<long long>(((long long)B.n_items + TB16_WARPS - 1) / TB16_WARPS * TB16_WARPS, (long long)g.sms * 32);
    dir_bytes = std::max(dir_bytes, (size_t)((long long)B.K * TB16_PLANES * tb16_steps16(B.rows_cap) * 32 * 4 * w));
    tmp_entries = std::max(tmp_entries, (size_t)((long long)tmp_cap * w));
  }
  R.pool = std::min<long long>(pool_cap, (long long)n * (R.max_la + R.max_lb + 1));
  CU(R.dir.alloc(dir_bytes));
  CU(R.tmp.alloc(tmp_entries * sizeof(FragOut)));
  CU(R.border.alloc((size_t)R.max_lb * 8 * std::max<long long>(warps, 1)));
  CU(R.scores.alloc(sizeof(int) * (size_t)n));
  CU(R.start.alloc(sizeof(long long) * (size_t)n));
  CU(R.count.alloc(sizeof(int) * (size_t)n));
  CU(R.frags.alloc(sizeof(FragOut) * (size_t)std::max<long long>(R.pool, 1)));
  CU(R.counters.alloc(8 * 16));
  CU(cudaMemsetAsync(R.counters.p, 0, 8 * 16, st));
  cudaStream_t sp = st;
  CU(R.lists.alloc(sizeof(int) * std::max<size_t>(plan.list.size(), 1)));
  CU(cudaMemcpyAsync(R.lists.p, plan.list.data(), sizeof(int) * plan.list.size(), cudaMemcpyHostToDevice, sp));
  if (!plan.buckets.empty()) {
    const int a1 = sc.alpha + 1;
    plan.mext.assign((size_t)a1 * a1, plan.pad_score);
    for (int i = 0; i < sc.alpha; ++i)
      for (int j = 0; j < sc.alpha; ++j) plan.mext[(size_t)i * a1 + j] = sc.matrix[i * sc.alpha + j];
    CU(R.matrix_ext.alloc(sizeof(int) * plan.mext.size()));
    CU(cudaMemcpyAsync(R.matrix_ext.p, plan.mext.data(), sizeof(int) * plan.mext.size(), cudaMemcpyHostToDevice, sp));
  }
  if (chunk && first && (rc = upload(2))) return rc;     // the sequence bytes, behind every small piece of this chunk
  // no wait here: plan is thread-local and this function synchronises the stream before it returns
  CU(cudaEventRecord(ev[1], st));

  unsigned long long *counters = R.counters.as<unsigned long long>();   // [0] wf32 pairs, [1] fragment total, [2..] tb16 items per bucket
  int bi = 0;
  for (const auto &B : plan.buckets) {
    if (B.max_la > tb16_max_la(mode == ABV_LOCAL, B.K)) return fail(ABV_ERR_ARGS, "tb16 plan: a query of %d columns in the K = %d bucket", B.max_la, B.K);
    Tb16Params q{};
    q.a = R.a_codes.as<uint8_t>(); q.a_off = R.a_off.as<long long>();
    q.b = R.b_codes.as<uint8_t>(); q.b_off = R.b_off.as<long long>();
    q.pair_list = R.lists.as<int>() + B.offset; q.n_items = B.n_items;
    q.matrix_ext = R.matrix_ext.as<int>(); q.a1 = sc.alpha + 1; q.go = sc.go; q.ge = sc.ge;
    q.rows_cap = B.rows_cap;
    q.scores = R.scores.as<int>();
    q.dir = R.dir.as<uint32_t>();
    q.frag_tmp = R.tmp.as<FragOut>(); q.frag_tmp_cap = tmp_cap;
    q.frags = R.frags.as<FragOut>(); q.frag_cap = R.pool;
    q.frag_total = counters + 1;
    q.frag_start = R.start.as<long long>(); q.frag_count = R.count.as<int>();
    q.item_counter = counters + 2 + bi;
    ++bi;
    rc = mode == ABV_LOCAL ? tb16_launch_mode<LOCAL>(B.K, st, q, dir_bytes, tmp_entries)
                           : tb16_launch_mode<GLOBAL>(B.K, st, q, dir_bytes, tmp_entries);
    if (rc) return rc;
  }
  Wf32Params p{};
  p.frag_total = counters + 1;
  if (plan.n_rest) {
    p.a = R.a_codes.as<uint8_t>(); p.a_off = R.a_off.as<long long>();
    p.b = R.b_codes.as<uint8_t>(); p.b_off = R.b_off.as<long long>();
    p.n_pairs = plan.n_rest; p.n_t = 0; p.q_base = 0; p.q_list = nullptr;
    p.pair_list = R.lists.as<int>() + plan.rest_offset;
    p.matrix = R.matrix.as<int>(); p.alpha = sc.alpha; p.go = sc.go; p.ge = sc.ge;
    p.scores = R.scores.as<int>(); p.score_rows_by_query = 0;
    p.dir = R.dir.as<uint32_t>(); p.dir_words_per_warp = dwords;
    p.frag_tmp = R.tmp.as<FragOut>(); p.frag_tmp_cap = tmp_cap;
    p.frags = R.frags.as<FragOut>(); p.frag_cap = R.pool;
    p.frag_start = R.start.as<long long>(); p.frag_count = R.count.as<int>();
    p.border = R.border.as<int>(); p.border_len = R.max_lb;
    p.pair_counter = counters;
    if ((rc = wf32_dispatch<true>(mode, grid, st, p, plan.rest_max_lb))) return rc;
  }
  CU(cudaEventRecord(ev[2], st));
  ABV_TRACE("pairs_run n=%d: kernels issued", n);
  CU(cudaMemcpyAsync(chunk ? R.pin_total : &R.total, p.frag_total, sizeof R.total, cudaMemcpyDeviceToHost, st));
  if (!chunk) CU(cudaStreamSynchronize(st));
  return ABV_OK;
}

// ---- pipeline of chunks for large one-to-one jobs -------------------------------------------------------------------
// The stage is upload (PCIe) -> fill + traceback (SMs) -> download (PCIe).  Run in one piece the three phases follow one
// another (1M read pairs: 14 + 23 + 5 ms); cut into chunks on two alternating streams, the upload of chunk c+1 and the
// download of chunk c-1 overlap the kernels of chunk c.  Every chunk is a complete job of its own (plan, scratch,
// fragment pool); the caller stitches the outputs together.  Blocks go back to the allocation cache only after every
// stream has been synchronised (the chunks' PairsRun objects outlive the pipeline).
// Chunk boundaries: nothing overlaps the first chunk's upload and the last chunk's download, so the pipeline starts with a
// quarter-size chunk, grows by 1.6x per chunk (an upload takes ~0.6 of the kernel time of the same pairs, so the next chunk has
// arrived when the current one finishes) up to the nominal size, and ends with a quarter-size chunk.  At most 16 chunks.
static std::vector<int> pairs_chunk_bounds(int n) {
  const long long per = g_opt.pairs_chunk;
  std::vector<int> b{0};
  if (per <= 0 || n < 2 * per) { b.push_back(n); return b; }
  long long s = std::max<long long>(per / 4, 1), pos = 0;
  while (pos < n) {
    long long take = std::min<long long>(s, n - pos);
    if (n - pos - take < per / 8) take = n - pos;          // no crumb at the end
    pos += take;
    b.push_back((int)pos);
    s = std::min<long long>(per, s * 8 / 5 + 1);
  }
  const int last = b[b.size() - 1] - b[b.size() - 2];
  if (last > per / 2) b.insert(b.end() - 1, n - (int)std::max<long long>(per / 4, 1));
  b.erase(std::unique(b.begin(), b.end()), b.end());          // no empty chunk (tiny chunk sizes)
  if (b.size() - 1 > 16) {                                   // pinned counter slots: 16 chunks
    b.assign(1, 0);
    for (int c = 1; c <= 16; ++c) b.push_back((int)((long long)n * c / 16));
  }
  return b;
}
// Validation of a pipelined job in two stages.  Walking the offsets of a million pairs takes 2 ms in which the device would wait
// for its first chunk: the pairs of the first two chunks are checked before anything is sent, the others once those two chunks are
// on their way.  No result leaves the device before the second stage has passed, so a rejected job still writes nothing.
struct StagedValidation {
  int mode; const char *a; const long long *a_off; const char *b; const long long *b_off; int n; const Scoring *sc;
  int first = 0;          // pairs covered by stage 1 (all of them if the job is not pipelined)
  int stage1(const std::vector<int> &bounds, int &max_la, int &max_lb) {
    first = bounds.size() > 2 ? bounds[2] : n;
    return pairs_validate(mode, a, a_off, b, b_off, first, *sc, max_la, max_lb);
  }
  int stage2(int &max_la, int &max_lb) const {
    if (first >= n) return ABV_OK;
    int la = 0, lb = 0;
    const int rc = pairs_validate(mode, a, a_off + first, b, b_off + first, n - first, *sc, la, lb, first);
    max_la = std::max(max_la, la); max_lb = std::max(max_lb, lb);
    return rc;
  }
};
static int pairs_chunk_begin(PairsRun &R, int c) {
  R.st = (c & 1) ? g.stream2 : g.stream;
  for (auto &e : R.ev) CU(cudaEventCreate(&e));
  R.own_plan = new Tb16Plan();
  R.pin_total = g.pin64 + c;
  return ABV_OK;
}

int abv_align_pairs(int mode, const char *a, const long long *a_off, const char *b, const long long *b_off, int n,
                    int alpha_len, const unsigned char *map256, const int *score_matrix,
                    int gap_open, int gap_extend,
                    int *scores, long long *frag_start, int *frag_count,
                    abv_frag *frags, long long frag_cap, long long *frag_total) {
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  int rc;
  Scoring sc{alpha_len, gap_open, gap_extend, map256, score_matrix};
  PairsRun R;
  const std::vector<int> bounds = pairs_chunk_bounds(n);
  StagedValidation val{mode, a, a_off, b, b_off, n, &sc};
  if ((rc = val.stage1(bounds, R.max_la, R.max_lb))) return rc;
  if (frag_cap < 0 || (frag_cap > 0 && !frags)) return fail(ABV_ERR_ARGS, "bad fragment pool");
  if (frag_total) *frag_total = 0;
  if (n == 0) return ABV_OK;
  if (!scores || !frag_start || !frag_count) return fail(ABV_ERR_ARGS, "NULL output");
  if ((rc = ctx_init())) return rc;
  for (double &v : t_timing) v = 0;
  cudaStream_t st = g.stream;
  if (bounds.size() > 2) {
    // pipeline of chunks: enqueue two ahead, then per chunk wait, download, enqueue the next one
    const int C = (int)bounds.size() - 1;
    std::vector<std::unique_ptr<PairsRun>> Rs;
    for (int c = 0; c < C; ++c) Rs.emplace_back(new PairsRun());
    auto lo_of = [&](int c) { return bounds[c]; };
    auto enqueue = [&](int c) -> int {
      PairsRun &Rc = *Rs[c];
      Rc.max_la = R.max_la; Rc.max_lb = R.max_lb;
      int r2;
      if ((r2 = pairs_chunk_begin(Rc, c))) return r2;
      const int lo = lo_of(c), hi = lo_of(c + 1);
      // a chunk's pool: twice its share of the caller's pool (fragments per pair vary little over >= 50k pairs), not the whole of it
      const long long share = std::min<long long>(frag_cap, 2 * (frag_cap / n + 1) * (long long)(hi - lo) + 4096);
      return pairs_run(Rc, mode, a, a_off + lo, b, b_off + lo, hi - lo, sc, share, true);
    };
    auto drain = [&]() { cudaStreamSynchronize(g.stream); cudaStreamSynchronize(g.stream2); cudaStreamSynchronize(g.stream3); };
    for (int c = 0; c < std::min(2, C); ++c)
      if ((rc = enqueue(c))) { drain(); return rc; }
    if ((rc = val.stage2(R.max_la, R.max_lb))) { drain(); return rc; }      // the other chunks' pairs, beside the first kernels
    long long base = 0;
    bool overflow = false, chunk_overflow = false;
    double d2h = 0, h2d = 0, cells = 0, ms_up = 0, ms_k = 0, ms_down = 0;
    ABV_TRACE("align_pairs: %d chunks, first two enqueued", C);
    for (int c = 0; c < C; ++c) {
      PairsRun &Rc = *Rs[c];
      const int lo = lo_of(c), nc = lo_of(c + 1) - lo;
      ABV_TRACE("chunk %d: wait", c);
      if (cudaStreamSynchronize(Rc.st) != cudaSuccess) { drain(); return fail(ABV_ERR_CUDA, "chunk %d failed: %s", c, cudaGetErrorString(cudaGetLastError())); }
      ABV_TRACE("chunk %d: done on the device", c);
      if (c + 2 < C && (rc = enqueue(c + 2))) { drain(); return rc; }      // keeps the other stream busy while this chunk downloads
      ABV_TRACE("chunk %d: next enqueued", c);
      Rc.total = *Rc.pin_total;
      const long long total = (long long)Rc.total;
      if (total > Rc.pool) chunk_overflow = true;
      if (total > Rc.pool || base + total > frag_cap) overflow = true;
      {
        const auto t0 = std::chrono::steady_clock::now();
        // the chunk has finished (synchronised above): its results leave on the download stream, beside the other chunks' kernels.
        // Scores, starts and counts always (they are valid on ABV_ERR_FRAG_CAP, include/ambivert_align.h); the fragments while they fit.
        CU(cudaMemcpyAsync(scores + lo, Rc.scores.p, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, g.stream3));
        CU(cudaMemcpyAsync(frag_start + lo, Rc.start.p, sizeof(long long) * (size_t)nc, cudaMemcpyDeviceToHost, g.stream3));
        CU(cudaMemcpyAsync(frag_count + lo, Rc.count.p, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, g.stream3));
        if (!overflow && total > 0) CU(cudaMemcpyAsync(frags + base, Rc.frags.p, sizeof(FragOut) * (size_t)total, cudaMemcpyDeviceToHost, g.stream3));
        CU(cudaStreamSynchronize(g.stream3));
        if (base)
          for (int i = 0; i < nc; ++i) frag_start[lo + i] += base;         // the chunk's pool starts at `base` of the caller's
        ABV_TRACE("chunk %d: downloaded", c);
        ms_down += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        d2h += 16.0 * nc + 8 + (overflow ? 0.0 : 16.0 * total);
      }
      // the chunk is complete on the device and on the host: its scratch goes back to the allocation cache for the chunks to come
      Rc.dir.release(); Rc.tmp.release(); Rc.border.release(); Rc.frags.release(); Rc.a_raw.release(); Rc.b_raw.release();
      base += total;
      h2d += Rc.h2d; cells += Rc.cells;
      float t0 = 0, t1 = 0;
      cudaEventElapsedTime(&t0, Rc.ev[0], Rc.ev[1]);
      cudaEventElapsedTime(&t1, Rc.ev[1], Rc.ev[2]);
      ms_up += t0; ms_k += t1;
    }
    drain();
    if (frag_total) *frag_total = base;
    t_timing[0] = ms_up; t_timing[1] = ms_k; t_timing[2] = ms_down; t_timing[3] = cells; t_timing[4] = 3 * C;
    t_timing[5] = h2d; t_timing[6] = d2h; t_timing[7] = ms_k;
    if (base > frag_cap) return fail(ABV_ERR_FRAG_CAP, "fragment pool of %lld entries is too small, %lld needed", frag_cap, base);
    if (!chunk_overflow) return ABV_OK;
    // the caller's pool is large enough but one chunk's share was not (very uneven fragment counts): run the job in one piece
    Rs.clear();
    for (double &v : t_timing) v = 0;
  }
  if ((rc = pairs_run(R, mode, a, a_off, b, b_off, n, sc, frag_cap, false))) return rc;
  CU(cudaMemcpyAsync(scores, R.scores.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(frag_start, R.start.p, sizeof(long long) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(frag_count, R.count.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  double d2h = 16.0 * n + 8;
  const long long ncopy = std::min<long long>((long long)R.total, R.pool);
  if (ncopy > 0) {
    CU(cudaMemcpyAsync(frags, R.frags.p, sizeof(FragOut) * (size_t)ncopy, cudaMemcpyDeviceToHost, st));
    d2h += 16.0 * ncopy;
  }
  CU(cudaEventRecord(g.ev[3], st));
  CU(cudaStreamSynchronize(st));
  if (frag_total) *frag_total = (long long)R.total;
  float t0 = 0, t1 = 0, t2 = 0;
  cudaEventElapsedTime(&t0, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&t1, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&t2, g.ev[2], g.ev[3]);
  t_timing[0] = t0; t_timing[1] = t1; t_timing[2] = t2; t_timing[3] = R.cells; t_timing[4] = 3;
  t_timing[5] = R.h2d; t_timing[6] = d2h; t_timing[7] = t1;
  if ((long long)R.total > frag_cap)
    return fail(ABV_ERR_FRAG_CAP, "fragment pool of %lld entries is too small, %llu needed", frag_cap, R.total);
  return ABV_OK;
}

// -------------------------------------------------------------------------------------------------
// The two host-side string builders that follow the aligner in the pipeline, on the device:
//   gapped rows  (ambivert.py:95-111 / :136-152)  and  merged read (ambivert.py:324-327)
// The pipelined form of assemble_impl for large jobs (see pairs_chunk_bounds): per chunk alignment + traceback, sizing, scan, assembly
// on the chunk's stream; downloads on the download stream.  Returns ABV_RETRY_WHOLE when a chunk's fragment pool or the
// caller's output buffer is too small: the caller then runs the job in one piece, which sizes the pool exactly and reports errors.
constexpr int ABV_RETRY_WHOLE = -1000;
struct AsmChunk {
  PairsRun R;
  DevMem lens, offs, dstatus, dsa, dsb, dout1, dout2, scan_tmp;
  long long bytes = 0;
  cudaEvent_t done = nullptr;
  ~AsmChunk() { if (done) cudaEventDestroy(done); }
};
static int assemble_chunked(const std::vector<int> &bounds, int what, int mode, const char *a, const long long *a_off, const char *b, const long long *b_off, int n,
                            const Scoring &sc, int &max_la, int &max_lb, const StagedValidation &val, int minimum_overlap, int *scores, int *status,
                            int *start_a, int *start_b, long long *out_off, char *out1, char *out2, long long out_cap, long long *out_total) {
  int rc;
  const int C = (int)bounds.size() - 1;
  std::vector<std::unique_ptr<AsmChunk>> Cs;
  for (int c = 0; c < C; ++c) Cs.emplace_back(new AsmChunk());
  auto lo_of = [&](int c) { return bounds[c]; };
  auto drain = [&]() { cudaStreamSynchronize(g.stream); cudaStreamSynchronize(g.stream2); cudaStreamSynchronize(g.stream3); };
  auto params = [&](AsmChunk &K, int nc) {
    AssembleParams q{};
    q.what = what; q.n = nc; q.min_overlap = minimum_overlap;
    q.a = K.R.a_raw.as<uint8_t>(); q.a_off = K.R.a_off.as<long long>();
    q.b = K.R.b_raw.as<uint8_t>(); q.b_off = K.R.b_off.as<long long>();
    q.frags = K.R.frags.as<FragOut>(); q.frag_start = K.R.start.as<long long>(); q.frag_count = K.R.count.as<int>();
    q.lens = K.lens.as<long long>(); q.offs = K.offs.as<long long>(); q.status = K.dstatus.as<int>();
    q.start_a = K.dsa.as<int>(); q.start_b = K.dsb.as<int>();
    return q;
  };
  // phase A: everything that does not need a size from the device
  auto phase_a = [&](int c) -> int {
    AsmChunk &K = *Cs[c];
    const int lo = lo_of(c), nc = lo_of(c + 1) - lo;
    K.R.max_la = max_la; K.R.max_lb = max_lb;
    int r2;
    if ((r2 = pairs_chunk_begin(K.R, c))) return r2;
    CU(cudaEventCreate(&K.done));
    if ((r2 = pairs_run(K.R, mode, a, a_off + lo, b, b_off + lo, nc, sc, std::max<long long>(1024, 8ll * nc), true))) return r2;
    cudaStream_t st = K.R.st;
    CU(K.lens.alloc(sizeof(long long) * ((size_t)nc + 1)));
    CU(K.offs.alloc(sizeof(long long) * ((size_t)nc + 1)));
    CU(K.dstatus.alloc(sizeof(int) * (size_t)nc));
    CU(K.dsa.alloc(sizeof(int) * (size_t)nc));
    CU(K.dsb.alloc(sizeof(int) * (size_t)nc));
    AssembleParams q = params(K, nc);
    assemble_len_kernel<<<(nc + 255) / 256, 256, 0, st>>>(q);
    CU(cudaGetLastError());
    size_t tmp_bytes = 0;
    CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, K.lens.as<long long>(), K.offs.as<long long>(), nc + 1, st));
    CU(K.scan_tmp.alloc(tmp_bytes));
    CU(cub::DeviceScan::ExclusiveSum(K.scan_tmp.p, tmp_bytes, K.lens.as<long long>(), K.offs.as<long long>(), nc + 1, st));
    CU(cudaMemcpyAsync(g.pin64 + 16 + c, K.offs.as<long long>() + nc, sizeof K.bytes, cudaMemcpyDeviceToHost, st));
    return ABV_OK;
  };
  for (int c = 0; c < std::min(2, C); ++c)
    if ((rc = phase_a(c))) { drain(); return rc; }
  if ((rc = val.stage2(max_la, max_lb))) { drain(); return rc; }      // the other chunks' pairs, beside the first kernels
  long long byte_base = 0;
  double d2h = 0, h2d = 0, cells = 0, ms_up = 0, ms_k = 0, ms_down = 0;
  for (int c = 0; c < C; ++c) {
    AsmChunk &K = *Cs[c];
    const int lo = lo_of(c), nc = lo_of(c + 1) - lo;
    cudaStream_t st = K.R.st;
    if (cudaStreamSynchronize(st) != cudaSuccess) { drain(); return fail(ABV_ERR_CUDA, "chunk %d failed: %s", c, cudaGetErrorString(cudaGetLastError())); }
    K.R.total = *K.R.pin_total;
    K.bytes = (long long)g.pin64[16 + c];
    if ((long long)K.R.total > K.R.pool || byte_base + K.bytes > out_cap) { drain(); return ABV_RETRY_WHOLE; }
    // phase B: the strings of this chunk, then the next chunk's phase A behind them on the same stream
    if (K.bytes > 0) {
      CU(K.dout1.alloc((size_t)K.bytes));
      if (what == ASSEMBLE_GAPPED) CU(K.dout2.alloc((size_t)K.bytes));
      AssembleParams q = params(K, nc);
      q.out1 = K.dout1.as<uint8_t>(); q.out2 = K.dout2.as<uint8_t>();
      const int wpb = 8;
      assemble_kernel<<<(nc + wpb - 1) / wpb, wpb * 32, 0, st>>>(q);
      CU(cudaGetLastError());
    }
    CU(cudaEventRecord(K.done, st));
    if (c + 2 < C && (rc = phase_a(c + 2))) { drain(); return rc; }
    const auto t0 = std::chrono::steady_clock::now();
    CU(cudaStreamWaitEvent(g.stream3, K.done, 0));
    CU(cudaMemcpyAsync(scores + lo, K.R.scores.p, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, g.stream3));
    CU(cudaMemcpyAsync(out_off + lo, K.offs.p, sizeof(long long) * (size_t)nc, cudaMemcpyDeviceToHost, g.stream3));
    if (start_a) CU(cudaMemcpyAsync(start_a + lo, K.dsa.p, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, g.stream3));
    if (start_b) CU(cudaMemcpyAsync(start_b + lo, K.dsb.p, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, g.stream3));
    if (K.bytes > 0) {
      if (out1) CU(cudaMemcpyAsync(out1 + byte_base, K.dout1.p, (size_t)K.bytes, cudaMemcpyDeviceToHost, g.stream3));
      if (what == ASSEMBLE_GAPPED && out2) CU(cudaMemcpyAsync(out2 + byte_base, K.dout2.p, (size_t)K.bytes, cudaMemcpyDeviceToHost, g.stream3));
    }
    CU(cudaMemcpyAsync(status + lo, K.dstatus.p, sizeof(int) * (size_t)nc, cudaMemcpyDeviceToHost, g.stream3));   // after assembly: it may flag -2
    CU(cudaStreamSynchronize(g.stream3));
    if (byte_base)
      for (int i = 0; i < nc; ++i) out_off[lo + i] += byte_base;
    ms_down += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    d2h += 24.0 * nc + (what == ASSEMBLE_GAPPED ? 2.0 : 1.0) * K.bytes;
    byte_base += K.bytes;
    h2d += K.R.h2d; cells += K.R.cells;
    float u = 0, k = 0;
    cudaEventElapsedTime(&u, K.R.ev[0], K.R.ev[1]);
    cudaEventElapsedTime(&k, K.R.ev[1], K.R.ev[2]);
    ms_up += u; ms_k += k;
  }
  drain();
  out_off[n] = byte_base;
  if (out_total) *out_total = byte_base;
  t_timing[0] = ms_up; t_timing[1] = ms_k; t_timing[2] = ms_down; t_timing[3] = cells; t_timing[4] = 6 * C;
  t_timing[5] = h2d; t_timing[6] = d2h; t_timing[7] = ms_k;
  return ABV_OK;
}

static int assemble_impl(int what, int mode, const char *a, const long long *a_off, const char *b, const long long *b_off, int n,
                         int alpha_len, const unsigned char *map256, const int *score_matrix, int gap_open, int gap_extend,
                         int minimum_overlap, int *scores, int *status, int *start_a, int *start_b,
                         long long *out_off, char *out1, char *out2, long long out_cap, long long *out_total) {
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  int rc;
  Scoring sc{alpha_len, gap_open, gap_extend, map256, score_matrix};
  PairsRun R;
  const std::vector<int> bounds = pairs_chunk_bounds(n);
  StagedValidation val{mode, a, a_off, b, b_off, n, &sc};
  if ((rc = val.stage1(bounds, R.max_la, R.max_lb))) return rc;
  if (out_total) *out_total = 0;
  if (n == 0) { if (out_off) out_off[0] = 0; return ABV_OK; }
  if (!scores || !status || !out_off || out_cap < 0) return fail(ABV_ERR_ARGS, "NULL output");
  if ((rc = ctx_init())) return rc;
  for (double &v : t_timing) v = 0;
  cudaStream_t st = g.stream;
  if (bounds.size() > 2) {
    rc = assemble_chunked(bounds, what, mode, a, a_off, b, b_off, n, sc, R.max_la, R.max_lb, val, minimum_overlap, scores, status, start_a, start_b,
                          out_off, out1, out2, out_cap, out_total);
    if (rc != ABV_RETRY_WHOLE) return rc;
    for (double &v : t_timing) v = 0;
  }
  long long cap = std::max<long long>(1024, 8ll * n);
  for (int attempt = 0; attempt < 2; ++attempt) {          // the fragment pool lives on the device only
    if ((rc = pairs_run(R, mode, a, a_off, b, b_off, n, sc, cap, true))) return rc;
    if ((long long)R.total <= R.pool) break;
    cap = (long long)R.total;
  }
  DevMem lens, offs, dstatus, dsa, dsb, dout1, dout2, scan_tmp;
  CU(lens.alloc(sizeof(long long) * ((size_t)n + 1)));
  CU(offs.alloc(sizeof(long long) * ((size_t)n + 1)));
  CU(dstatus.alloc(sizeof(int) * (size_t)n));
  CU(dsa.alloc(sizeof(int) * (size_t)n));
  CU(dsb.alloc(sizeof(int) * (size_t)n));
  AssembleParams q{};
  q.what = what; q.n = n; q.min_overlap = minimum_overlap;
  q.a = R.a_raw.as<uint8_t>(); q.a_off = R.a_off.as<long long>();
  q.b = R.b_raw.as<uint8_t>(); q.b_off = R.b_off.as<long long>();
  q.frags = R.frags.as<FragOut>(); q.frag_start = R.start.as<long long>(); q.frag_count = R.count.as<int>();
  q.lens = lens.as<long long>(); q.offs = offs.as<long long>(); q.status = dstatus.as<int>();
  q.start_a = dsa.as<int>(); q.start_b = dsb.as<int>();
  assemble_len_kernel<<<(n + 255) / 256, 256, 0, st>>>(q);
  CU(cudaGetLastError());
  size_t tmp_bytes = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, lens.as<long long>(), offs.as<long long>(), n + 1, st));
  CU(scan_tmp.alloc(tmp_bytes));
  CU(cub::DeviceScan::ExclusiveSum(scan_tmp.p, tmp_bytes, lens.as<long long>(), offs.as<long long>(), n + 1, st));
  long long total = 0;
  CU(cudaMemcpyAsync(&total, offs.as<long long>() + n, sizeof total, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (out_total) *out_total = total;
  CU(cudaMemcpyAsync(scores, R.scores.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(out_off, offs.p, sizeof(long long) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st));
  if (start_a) CU(cudaMemcpyAsync(start_a, dsa.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  if (start_b) CU(cudaMemcpyAsync(start_b, dsb.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  double d2h = 24.0 * n;
  if (total > out_cap) {
    CU(cudaMemcpyAsync(status, dstatus.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return fail(ABV_ERR_FRAG_CAP, "output buffer of %lld bytes is too small, %lld needed", out_cap, total);
  }
  if (total > 0) {
    CU(dout1.alloc((size_t)total));
    if (what == ASSEMBLE_GAPPED) CU(dout2.alloc((size_t)total));
    q.out1 = dout1.as<uint8_t>(); q.out2 = dout2.as<uint8_t>();
    const int wpb = 8;
    assemble_kernel<<<(n + wpb - 1) / wpb, wpb * 32, 0, st>>>(q);
    CU(cudaGetLastError());
    if (out1) CU(cudaMemcpyAsync(out1, dout1.p, (size_t)total, cudaMemcpyDeviceToHost, st));
    if (what == ASSEMBLE_GAPPED && out2) CU(cudaMemcpyAsync(out2, dout2.p, (size_t)total, cudaMemcpyDeviceToHost, st));
    d2h += (what == ASSEMBLE_GAPPED ? 2.0 : 1.0) * total;
  }
  CU(cudaMemcpyAsync(status, dstatus.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));   // after assembly: it may flag -2
  CU(cudaEventRecord(g.ev[3], st));
  CU(cudaStreamSynchronize(st));
  float t0 = 0, t1 = 0, t2 = 0;
  cudaEventElapsedTime(&t0, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&t1, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&t2, g.ev[2], g.ev[3]);
  t_timing[0] = t0; t_timing[1] = t1; t_timing[2] = t2; t_timing[3] = R.cells; t_timing[4] = 6;
  t_timing[5] = R.h2d; t_timing[6] = d2h; t_timing[7] = t1;
  return ABV_OK;
}

int abv_merge_pairs(const char *fwd, const long long *fwd_off, const char *rev, const long long *rev_off, int n,
                    int alpha_len, const unsigned char *map256, const int *score_matrix, int gap_open, int gap_extend,
                    int minimum_overlap, int *scores, int *status, long long *merged_off, char *merged, long long merged_cap,
                    long long *merged_total) {
  if (!map256) return fail(ABV_ERR_ARGS, "abv_merge_pairs needs character input and a map");
  return assemble_impl(ASSEMBLE_MERGE, ABV_LOCAL, fwd, fwd_off, rev, rev_off, n, alpha_len, map256, score_matrix, gap_open, gap_extend,
                       minimum_overlap, scores, status, nullptr, nullptr, merged_off, merged, nullptr, merged_cap, merged_total);
}

int abv_gapped_pairs(int mode, const char *a, const long long *a_off, const char *b, const long long *b_off, int n,
                     int alpha_len, const unsigned char *map256, const int *score_matrix, int gap_open, int gap_extend,
                     int *scores, int *status, int *start_a, int *start_b, long long *out_off, char *gapped_a, char *gapped_b,
                     long long out_cap, long long *out_total) {
  if (!map256) return fail(ABV_ERR_ARGS, "abv_gapped_pairs needs character input and a map");
  return assemble_impl(ASSEMBLE_GAPPED, mode, a, a_off, b, b_off, n, alpha_len, map256, score_matrix, gap_open, gap_extend,
                       0, scores, status, start_a, start_b, out_off, gapped_a, gapped_b, out_cap, out_total);
}

// -------------------------------------------------------------------------------------------------
// Clustering front end: md5 keys of read pairs + grouping, on the device.
__global__ void gather_u64_kernel(const unsigned long long *in, const unsigned int *idx, int n, unsigned long long *out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[idx[i]];
}

int abv_cluster_pairs(const char *fwd, const long long *fwd_off, const char *rev, const long long *rev_off, int n,
                      int has_trim5, int trim5, int has_trim3, int trim3,
                      int *cluster_of, int *n_clusters, unsigned char *digest, int *count, int *first) {
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  if (n < 0) return fail(ABV_ERR_ARGS, "negative count");
  if (n_clusters) *n_clusters = 0;
  if (n == 0) return ABV_OK;
  if (!fwd_off || !rev_off || !cluster_of || !n_clusters || !digest || !count || !first) return fail(ABV_ERR_ARGS, "NULL argument");
  const long long fbytes = fwd_off[n] - fwd_off[0], rbytes = rev_off[n] - rev_off[0];
  if ((fbytes > 0 && !fwd) || (rbytes > 0 && !rev)) return fail(ABV_ERR_ARGS, "NULL read buffer");
  for (int i = 0; i < n; ++i)
    if (fwd_off[i + 1] < fwd_off[i] || rev_off[i + 1] < rev_off[i]) return fail(ABV_ERR_ARGS, "offsets must not decrease");
  int rc;
  if ((rc = ctx_init())) return rc;
  for (double &v : t_timing) v = 0;
  cudaStream_t st = g.stream;
  if (!g.md5_ready) {  // K[i] = floor(2^32 * |sin(i + 1)|)  (RFC 1321, section 3.4)
    uint32_t k[64];
    for (int i = 0; i < 64; ++i) k[i] = (uint32_t)(long long)floor(fabs(sin((double)(i + 1))) * 4294967296.0);
    CU(cudaMemcpyToSymbol(c_md5_k, k, sizeof k));
    g.md5_ready = true;
  }
  CU(cudaEventRecord(g.ev[0], st));
  DevMem df, dfo, dr, dro, lo, hi, lo2, hi2, idx, idx2, flag, seg, tmp, run_start, run_first, run_first_s, run_id, order_run, rank_of_run,
      d_cluster, d_digest, d_count, d_first;
  std::vector<long long> fo(n + 1), ro(n + 1);
  for (int i = 0; i <= n; ++i) { fo[i] = fwd_off[i] - fwd_off[0]; ro[i] = rev_off[i] - rev_off[0]; }
  CU(df.alloc((size_t)fbytes + 16)); CU(dr.alloc((size_t)rbytes + 16));
  CU(dfo.alloc(sizeof(long long) * (n + 1))); CU(dro.alloc(sizeof(long long) * (n + 1)));
  if (fbytes) CU(cudaMemcpyAsync(df.p, fwd + fwd_off[0], (size_t)fbytes, cudaMemcpyHostToDevice, st));
  if (rbytes) CU(cudaMemcpyAsync(dr.p, rev + rev_off[0], (size_t)rbytes, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dfo.p, fo.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dro.p, ro.data(), sizeof(long long) * (n + 1), cudaMemcpyHostToDevice, st));
  const size_t n8 = sizeof(unsigned long long) * (size_t)n, n4 = sizeof(int) * (size_t)n;
  CU(lo.alloc(n8)); CU(hi.alloc(n8)); CU(lo2.alloc(n8)); CU(hi2.alloc(n8)); CU(idx.alloc(n4)); CU(idx2.alloc(n4));
  CU(flag.alloc(n4)); CU(seg.alloc(n4));
  CU(cudaEventRecord(g.ev[1], st));
  const int tb = 256, nb = (n + tb - 1) / tb;
  Md5Params mp{};
  mp.f = df.as<uint8_t>(); mp.f_off = dfo.as<long long>(); mp.r = dr.as<uint8_t>(); mp.r_off = dro.as<long long>(); mp.n = n;
  mp.has5 = has_trim5; mp.trim5 = trim5; mp.has3 = has_trim3; mp.trim3 = trim3;
  mp.lo = lo.as<unsigned long long>(); mp.hi = hi.as<unsigned long long>(); mp.index = idx.as<unsigned int>();
  md5_pairs_kernel<<<(n + 127) / 128, 128, 0, st>>>(mp);
  CU(cudaGetLastError());
  // stable LSD radix sort of the 128-bit digests: by the low word, then by the high word
  size_t t1 = 0, t2 = 0, t3 = 0, t4 = 0;
  CU(cub::DeviceRadixSort::SortPairs(nullptr, t1, lo.as<unsigned long long>(), lo2.as<unsigned long long>(), idx.as<unsigned int>(), idx2.as<unsigned int>(), n, 0, 64, st));
  CU(cub::DeviceScan::InclusiveSum(nullptr, t2, flag.as<int>(), seg.as<int>(), n, st));
  CU(cub::DeviceRadixSort::SortPairs(nullptr, t3, run_first.as<unsigned int>(), run_first_s.as<unsigned int>(), run_id.as<int>(), order_run.as<int>(), n, 0, 32, st));
  t4 = std::max(t1, std::max(t2, t3));
  CU(tmp.alloc(t4));
  size_t tt = t4;
  CU(cub::DeviceRadixSort::SortPairs(tmp.p, tt, lo.as<unsigned long long>(), lo2.as<unsigned long long>(), idx.as<unsigned int>(), idx2.as<unsigned int>(), n, 0, 64, st));
  gather_u64_kernel<<<nb, tb, 0, st>>>(hi.as<unsigned long long>(), idx2.as<unsigned int>(), n, hi2.as<unsigned long long>());
  tt = t4;
  CU(cub::DeviceRadixSort::SortPairs(tmp.p, tt, hi2.as<unsigned long long>(), hi.as<unsigned long long>(), idx2.as<unsigned int>(), idx.as<unsigned int>(), n, 0, 64, st));
  // idx = input index of the i-th pair in sorted order; hi = sorted high words; fetch the matching low words
  CU(cudaMemcpyAsync(lo2.p, lo.p, n8, cudaMemcpyDeviceToDevice, st));       // lo2 = low words in INPUT order
  gather_u64_kernel<<<nb, tb, 0, st>>>(lo2.as<unsigned long long>(), idx.as<unsigned int>(), n, lo.as<unsigned long long>());
  cluster_flag_kernel<<<nb, tb, 0, st>>>(lo.as<unsigned long long>(), hi.as<unsigned long long>(), n, flag.as<int>());
  tt = t4;
  CU(cub::DeviceScan::InclusiveSum(tmp.p, tt, flag.as<int>(), seg.as<int>(), n, st));
  int n_runs = 0;
  CU(cudaMemcpyAsync(&n_runs, seg.as<int>() + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  const size_t r4 = sizeof(int) * (size_t)n_runs;
  CU(run_start.alloc(r4)); CU(run_first.alloc(r4)); CU(run_first_s.alloc(r4)); CU(run_id.alloc(r4)); CU(order_run.alloc(r4)); CU(rank_of_run.alloc(r4));
  CU(d_cluster.alloc(n4)); CU(d_digest.alloc((size_t)16 * n_runs)); CU(d_count.alloc(r4)); CU(d_first.alloc(r4));
  cluster_runs_kernel<<<nb, tb, 0, st>>>(flag.as<int>(), seg.as<int>(), idx.as<unsigned int>(), n, run_start.as<int>(), run_first.as<unsigned int>(), run_id.as<int>());
  tt = t4;
  CU(cub::DeviceRadixSort::SortPairs(tmp.p, tt, run_first.as<unsigned int>(), run_first_s.as<unsigned int>(), run_id.as<int>(), order_run.as<int>(), n_runs, 0, 32, st));
  cluster_emit_kernel<<<(n_runs + tb - 1) / tb, tb, 0, st>>>(order_run.as<int>(), n_runs, run_start.as<int>(), run_first_s.as<unsigned int>(), n,
                                                            lo.as<unsigned long long>(), hi.as<unsigned long long>(), d_digest.as<unsigned char>(),
                                                            d_count.as<int>(), d_first.as<int>(), rank_of_run.as<int>());
  cluster_assign_kernel<<<nb, tb, 0, st>>>(seg.as<int>(), idx.as<unsigned int>(), n, rank_of_run.as<int>(), d_cluster.as<int>(), d_count.as<int>());
  CU(cudaGetLastError());
  CU(cudaEventRecord(g.ev[2], st));
  CU(cudaMemcpyAsync(cluster_of, d_cluster.p, n4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(digest, d_digest.p, (size_t)16 * n_runs, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(count, d_count.p, r4, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(first, d_first.p, r4, cudaMemcpyDeviceToHost, st));
  CU(cudaEventRecord(g.ev[3], st));
  CU(cudaStreamSynchronize(st));
  *n_clusters = n_runs;
  float a = 0, b = 0, c = 0;
  cudaEventElapsedTime(&a, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&b, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&c, g.ev[2], g.ev[3]);
  t_timing[0] = a; t_timing[1] = b; t_timing[2] = c; t_timing[3] = 0; t_timing[4] = 11;
  t_timing[5] = (double)fbytes + rbytes + 16.0 * (n + 1); t_timing[6] = 4.0 * n + 24.0 * n_runs; t_timing[7] = b;
  return ABV_OK;
}

// -------------------------------------------------------------------------------------------------
// Variant scan of aligned row pairs (call_variants.py:32-128) on rows that are already
// on the device: COUNT pass, exclusive scan, EMIT pass.  d_status [n], d_count/d_start [n+1].
static int variants_on_device(const uint8_t *d_sample, const long long *d_sample_off, const uint8_t *d_ref, const long long *d_ref_off,
                              int n, int softmask, int gap, cudaStream_t st, DevMem &d_status, DevMem &d_count, DevMem &d_start,
                              DevMem &d_recs, long long &total) {
  static_assert(sizeof(VariantRec) == sizeof(abv_variant), "record layout");
  DevMem scan_tmp;
  CU(d_status.alloc(sizeof(int) * (size_t)n));
  CU(d_count.alloc(sizeof(long long) * ((size_t)n + 1)));
  CU(d_start.alloc(sizeof(long long) * ((size_t)n + 1)));
  VariantParams p{};
  p.sample = d_sample; p.sample_off = d_sample_off; p.ref = d_ref; p.ref_off = d_ref_off;
  p.n = n; p.softmask = softmask; p.gap = gap;
  p.status = d_status.as<int>(); p.count = d_count.as<long long>(); p.start = d_start.as<long long>();
  const int tb = 128, nb = (n + tb - 1) / tb;
  variants_kernel<false><<<nb, tb, 0, st>>>(p);
  CU(cudaGetLastError());
  size_t tmp_bytes = 0;
  CU(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_count.as<long long>(), d_start.as<long long>(), n + 1, st));
  CU(scan_tmp.alloc(tmp_bytes));
  CU(cub::DeviceScan::ExclusiveSum(scan_tmp.p, tmp_bytes, d_count.as<long long>(), d_start.as<long long>(), n + 1, st));
  total = 0;
  CU(cudaMemcpyAsync(&total, d_start.as<long long>() + n, sizeof total, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (total > 0) {
    CU(d_recs.alloc(sizeof(VariantRec) * (size_t)total));
    p.recs = d_recs.as<VariantRec>();
    variants_kernel<true><<<nb, tb, 0, st>>>(p);
    CU(cudaGetLastError());
  }
  return ABV_OK;
}

// copies status / starts / counts / records of a finished variant scan to the caller
static int variants_fetch(int n, cudaStream_t st, const DevMem &d_status, const DevMem &d_start, const DevMem &d_recs, long long total,
                          int *status, long long *rec_start, int *rec_count, abv_variant *recs, long long rec_cap) {
  std::vector<long long> starts((size_t)n + 1);
  CU(cudaMemcpyAsync(status, d_status.p, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
  CU(cudaMemcpyAsync(starts.data(), d_start.p, sizeof(long long) * ((size_t)n + 1), cudaMemcpyDeviceToHost, st));
  const bool fits = total <= rec_cap;
  if (fits && total > 0) CU(cudaMemcpyAsync(recs, d_recs.p, sizeof(abv_variant) * (size_t)total, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  for (int i = 0; i < n; ++i) { rec_start[i] = starts[i]; rec_count[i] = (int)(starts[i + 1] - starts[i]); }
  if (!fits) return fail(ABV_ERR_FRAG_CAP, "record buffer of %lld entries is too small, %lld needed", rec_cap, total);
  return ABV_OK;
}

int abv_call_variants(const char *sample, const long long *sample_off, const char *ref, const long long *ref_off, int n,
                      int softmask, int gap, int *status, long long *rec_start, int *rec_count,
                      abv_variant *recs, long long rec_cap, long long *rec_total) {
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  if (n < 0) return fail(ABV_ERR_ARGS, "negative count");
  if (rec_total) *rec_total = 0;
  if (n == 0) return ABV_OK;
  if (!sample_off || !ref_off || !status || !rec_start || !rec_count || rec_cap < 0 || (rec_cap > 0 && !recs))
    return fail(ABV_ERR_ARGS, "NULL argument");
  if (gap < 0 || gap > 255) return fail(ABV_ERR_ARGS, "gap must be one byte");
  const long long sbytes = sample_off[n] - sample_off[0], rbytes = ref_off[n] - ref_off[0];
  if ((sbytes > 0 && !sample) || (rbytes > 0 && !ref)) return fail(ABV_ERR_ARGS, "NULL row buffer");
  for (int i = 0; i < n; ++i) {
    if (sample_off[i + 1] < sample_off[i] || ref_off[i + 1] < ref_off[i]) return fail(ABV_ERR_ARGS, "offsets must not decrease");
    if (sample_off[i + 1] - sample_off[i] > INT_MAX / 2) return fail(ABV_ERR_ARGS, "row %d is too long", i);
  }
  int rc;
  if ((rc = ctx_init())) return rc;
  for (double &v : t_timing) v = 0;
  cudaStream_t st = g.stream;
  CU(cudaEventRecord(g.ev[0], st));
  DevMem ds, dso, dr, dro, d_status, d_count, d_start, d_recs;
  std::vector<long long> so((size_t)n + 1), ro((size_t)n + 1);
  for (int i = 0; i <= n; ++i) { so[i] = sample_off[i] - sample_off[0]; ro[i] = ref_off[i] - ref_off[0]; }
  CU(ds.alloc((size_t)sbytes + 16)); CU(dr.alloc((size_t)rbytes + 16));
  CU(dso.alloc(sizeof(long long) * ((size_t)n + 1))); CU(dro.alloc(sizeof(long long) * ((size_t)n + 1)));
  if (sbytes) CU(cudaMemcpyAsync(ds.p, sample + sample_off[0], (size_t)sbytes, cudaMemcpyHostToDevice, st));
  if (rbytes) CU(cudaMemcpyAsync(dr.p, ref + ref_off[0], (size_t)rbytes, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dso.p, so.data(), sizeof(long long) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(dro.p, ro.data(), sizeof(long long) * ((size_t)n + 1), cudaMemcpyHostToDevice, st));
  CU(cudaEventRecord(g.ev[1], st));
  long long total = 0;
  if ((rc = variants_on_device(ds.as<uint8_t>(), dso.as<long long>(), dr.as<uint8_t>(), dro.as<long long>(), n, softmask, gap, st,
                               d_status, d_count, d_start, d_recs, total))) return rc;
  CU(cudaEventRecord(g.ev[2], st));
  if (rec_total) *rec_total = total;
  rc = variants_fetch(n, st, d_status, d_start, d_recs, total, status, rec_start, rec_count, recs, rec_cap);
  CU(cudaEventRecord(g.ev[3], st));
  CU(cudaStreamSynchronize(st));
  float a = 0, b = 0, c = 0;
  cudaEventElapsedTime(&a, g.ev[0], g.ev[1]);
  cudaEventElapsedTime(&b, g.ev[1], g.ev[2]);
  cudaEventElapsedTime(&c, g.ev[2], g.ev[3]);
  t_timing[0] = a; t_timing[1] = b; t_timing[2] = c; t_timing[3] = 0; t_timing[4] = total > 0 ? 4 : 3;
  t_timing[5] = (double)sbytes + rbytes + 16.0 * (n + 1); t_timing[6] = 12.0 * n + 8.0 + 20.0 * (double)total; t_timing[7] = b;
  return rc;
}

void *abv_pinned_alloc(unsigned long long bytes) {
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  if (ctx_init() != ABV_OK) return nullptr;
  void *p = nullptr;
  if (cudaHostAlloc(&p, (size_t)std::max<unsigned long long>(bytes, 16), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void abv_pinned_free(void *p) {
  if (p && cudaFreeHost(p) != cudaSuccess) cudaGetLastError();
}

int abv_get_timing(double *out, int n) {
  if (!out || n <= 0) return 0;
  n = std::min(n, 8);
  for (int i = 0; i < n; ++i) out[i] = t_timing[i];
  return n;
}

int abv_pairs_chunk_bounds(int n, int *bounds, int cap) {
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  const std::vector<int> b = pairs_chunk_bounds(std::max(n, 0));
  for (int i = 0; i < (int)b.size() && i < cap && bounds; ++i) bounds[i] = b[i];
  return (int)b.size();
}

int abv_int_peak(int which, int iters, double *lane_ops_per_s, double *sm_clock_mhz) {
  std::lock_guard<std::recursive_mutex> lk(g.mu);
  int rc;
  if ((rc = ctx_init())) return rc;
  if (which < 0 || which > 11 || iters <= 0) return fail(ABV_ERR_ARGS, "bad microbenchmark arguments");
  const int threads = 256, grid = g.sms * 8;
  DevMem out;
  CU(out.alloc(sizeof(uint32_t) * (size_t)grid * threads));
  auto run = [&](int it) {
    switch (which) {
      case 0: int_peak_kernel<0><<<grid, threads, 0, g.stream>>>(out.as<uint32_t>(), it, 12345u); break;
      case 1: int_peak_kernel<1><<<grid, threads, 0, g.stream>>>(out.as<uint32_t>(), it, 12345u); break;
      case 2: int_peak_kernel<2><<<grid, threads, 0, g.stream>>>(out.as<uint32_t>(), it, 12345u); break;
      case 3: int_peak_kernel<3><<<grid, threads, 0, g.stream>>>(out.as<uint32_t>(), it, 12345u); break;
      case 4: int_peak_kernel<4><<<grid, threads, 0, g.stream>>>(out.as<uint32_t>(), it, 12345u); break;
      case 5: int_peak_kernel<5><<<grid, threads, 0, g.stream>>>(out.as<uint32_t>(), it, 12345u); break;
      case 6: int_peak_kernel<6><<<grid, threads, 0, g.stream>>>(out.as<uint32_t>(), it, 12345u); break;
      case 7: int_peak_kernel<7><<<grid, threads, 0, g.stream>>>(out.as<uint32_t>(), it, 12345u); break;
      case 9: int_peak_kernel<9><<<grid, threads, 0, g.stream>>>(out.as<uint32_t>(), it, 12345u); break;
      case 10: int_peak_kernel<10><<<grid, threads, 0, g.stream>>>(out.as<uint32_t>(), it, 12345u); break;
      case 11: int_peak_kernel<11><<<grid, threads, 0, g.stream>>>(out.as<uint32_t>(), it, 12345u); break;
      default: int_peak_kernel<8><<<grid, threads, 0, g.stream>>>(out.as<uint32_t>(), it, 12345u); break;
    }
  };
  run(std::max(1, iters / 10));   // warm-up
  CU(cudaGetLastError());
  CU(cudaEventRecord(g.ev[4], g.stream));
  run(iters);
  CU(cudaEventRecord(g.ev[5], g.stream));
  CU(cudaStreamSynchronize(g.stream));
  CU(cudaGetLastError());
  float ms = 0;
  cudaEventElapsedTime(&ms, g.ev[4], g.ev[5]);
  const double ops = (double)grid * threads * INT_PEAK_CHAINS * INT_PEAK_OPS[which] * (double)iters;
  if (lane_ops_per_s) *lane_ops_per_s = ops / (ms * 1e-3);
  if (sm_clock_mhz) {
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, g.device);
    *sm_clock_mhz = khz / 1000.0;
  }
  return ABV_OK;
}

// =================================================================================================
// Part 1: the reference's own ABI (lib/align/align.h:43-125)

int align_frag_count(AlignFrag *f) {
  int i = 0;
  for (; f; f = f->next) ++i;
  return i;
}

void align_frag_free(AlignFrag *f) {
  while (f) {
    AlignFrag *nx = f->next;
    free(f);
    f = nx;
  }
}

void alignment_free(Alignment *alignment) {
  if (!alignment) return;     // the reference would crash here (alignment.c:36-39); harmless superset
  align_frag_free(alignment->align_frag);
  free(alignment);
}

Alignment *alignment_new(AlignFrag *align_frag, int score) {
  Alignment *al = (Alignment *)malloc(sizeof(Alignment));
  if (al) {
    al->align_frag = align_frag;
    al->frag_count = align_frag_count(align_frag);
    al->score = score;
  }
  return al;
}

}  // extern "C" (the helpers below are C++)

// ---- the lean single-pair path of the legacy symbols -------------------------------------------------
// The reference calls these from multiprocessing.dummy worker threads, one pair per call
// (ambivert.py:474-476).  A call costs one host->device copy, one launch of wf32_one_kernel and one
// device->host copy on a stream of its own, so calls from different host threads overlap on the GPU.  The
// per-call resources (stream, pinned staging, device scratch) are pooled: worker threads come and go (the
// reference builds a new Pool per amplicon), the contexts stay.
struct OneCtx {
  int device = 0;                                   // streams and allocations belong to one device
  cudaStream_t st = nullptr;
  unsigned char *h = nullptr; size_t h_cap = 0;     // pinned: [input blob][result blob]
  unsigned char *d = nullptr; size_t d_cap = 0;     // device: input blob + scratch + result blob
};
std::mutex g_one_mu;
std::vector<OneCtx *> g_one_free;

static OneCtx *one_acquire() {
  const int device = g.device;
  {
    std::lock_guard<std::mutex> lk(g_one_mu);
    for (size_t i = g_one_free.size(); i-- > 0;)
      if (g_one_free[i]->device == device) { OneCtx *c = g_one_free[i]; g_one_free.erase(g_one_free.begin() + i); return c; }
  }
  OneCtx *c = new OneCtx();
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); delete c; return nullptr; }
  return c;
}
static void one_release(OneCtx *c) {
  std::lock_guard<std::mutex> lk(g_one_mu);
  g_one_free.push_back(c);
}
static bool one_reserve(OneCtx *c, size_t h_bytes, size_t d_bytes) {
  if (h_bytes > c->h_cap) {
    if (c->h) cudaFreeHost(c->h);
    c->h = nullptr; c->h_cap = 0;
    const size_t want = std::max<size_t>(h_bytes * 2, 64 << 10);
    if (cudaMallocHost((void **)&c->h, want) != cudaSuccess) { cudaGetLastError(); return false; }
    c->h_cap = want;
  }
  if (d_bytes > c->d_cap) {
    if (c->d) cudaFree(c->d);
    c->d = nullptr; c->d_cap = 0;
    const size_t want = std::max<size_t>(d_bytes * 2, 256 << 10);
    if (cudaMalloc((void **)&c->d, want) != cudaSuccess) { cudaGetLastError(); return false; }
    c->d_cap = want;
  }
  return true;
}

template <int MODE>
static void one_launch(cudaStream_t st, const Wf32OneParams &p) {
  const size_t smem = (size_t)((p.la + 31) / 32) * 2 * p.lb * 4 + p.la + p.lb;
  if (smem <= 150 * 1024) {
    auto kern = wf32_one_kernel<MODE, true>;
    if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 150 * 1024);   // rare (long pairs); the attribute is per device
    kern<<<1, WF32_ONE_WARPS * 32, smem, st>>>(p);
  } else {
    wf32_one_kernel<MODE, false><<<1, WF32_ONE_WARPS * 32, 0, st>>>(p);
  }
}

// returns 1 = done (result in *out), 0 = not eligible for the lean path, -1 = CUDA failure
static int legacy_align_lean(int mode, const char *sa, int la, const char *sb, int lb, int alpha, const unsigned char *map,
                             const int *matrix, int go, int ge, Alignment **out) {
  if (alpha > WF32_MAX_SMEM_ALPHA || la > (1 << 15) || lb > (1 << 15)) return 0;
  const size_t dir_bytes = (size_t)dir_words(la, lb) * 4;
  if (dir_bytes > ((size_t)256 << 20)) return 0;
  {
    std::lock_guard<std::recursive_mutex> lk(g.mu);
    if (ctx_init() != ABV_OK) return -1;
  }
  auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const int frag_cap = la + lb + 2;
  const size_t in_bytes = 256 + (size_t)alpha * alpha * 4 + la + lb;
  const size_t out_bytes = 16 + (size_t)frag_cap * sizeof(FragOut);
  const size_t o_in = 0, o_codes = up(in_bytes), o_dir = o_codes + up((size_t)la + lb), o_border = o_dir + up(dir_bytes),
               o_tmp = o_border + up((size_t)lb * 8 * ((la + 31) / 32)), o_out = o_tmp + up((size_t)frag_cap * sizeof(FragOut)), d_total = o_out + up(out_bytes);
  OneCtx *c = one_acquire();
  if (!c) return -1;
  int rc = -1;
  do {
    if (cudaSetDevice(g.device) != cudaSuccess) break;
    if (!one_reserve(c, up(in_bytes) + up(out_bytes), d_total)) break;
    unsigned char *hin = c->h;
    int *hout = reinterpret_cast<int *>(c->h + up(in_bytes));
    if (map) memcpy(hin, map, 256); else memset(hin, 0, 256);
    memcpy(hin + 256, matrix, (size_t)alpha * alpha * 4);
    memcpy(hin + 256 + (size_t)alpha * alpha * 4, sa, la);
    memcpy(hin + 256 + (size_t)alpha * alpha * 4 + la, sb, lb);
    if (cudaMemcpyAsync(c->d + o_in, hin, in_bytes, cudaMemcpyHostToDevice, c->st) != cudaSuccess) break;
    Wf32OneParams p{};
    p.in = c->d + o_in; p.la = la; p.lb = lb; p.alpha = alpha; p.go = go; p.ge = ge; p.use_map = map ? 1 : 0;
    p.codes = c->d + o_codes; p.dir = reinterpret_cast<uint32_t *>(c->d + o_dir);
    p.border = reinterpret_cast<int *>(c->d + o_border);
    p.tmp = reinterpret_cast<FragOut *>(c->d + o_tmp); p.tmp_cap = frag_cap;
    p.out = reinterpret_cast<int *>(c->d + o_out); p.out_frag_cap = frag_cap;
    switch (mode) {
      case ABV_OVERLAP: one_launch<OVERLAP>(c->st, p); break;
      case ABV_LOCAL: one_launch<LOCAL>(c->st, p); break;
      case ABV_GLOBAL: one_launch<GLOBAL>(c->st, p); break;
      default: one_launch<GLOCAL>(c->st, p); break;
    }
    if (cudaGetLastError() != cudaSuccess) break;
    const int first = std::min(frag_cap, 62);           // 16 + 62*16 = 1008 bytes: most alignments fit the first copy
    if (cudaMemcpyAsync(hout, c->d + o_out, 16 + (size_t)first * sizeof(FragOut), cudaMemcpyDeviceToHost, c->st) != cudaSuccess) break;
    if (cudaStreamSynchronize(c->st) != cudaSuccess) break;
    const int score = hout[0], n = hout[1];
    if (n > first) {
      if (cudaMemcpyAsync(hout + 4 + first * 4, c->d + o_out + 16 + (size_t)first * sizeof(FragOut), (size_t)(n - first) * sizeof(FragOut),
                          cudaMemcpyDeviceToHost, c->st) != cudaSuccess) break;
      if (cudaStreamSynchronize(c->st) != cudaSuccess) break;
    }
    const FragOut *fr = reinterpret_cast<const FragOut *>(hout + 4);
    AlignFrag *head = nullptr;
    bool oom = false;
    for (int i = n - 1; i >= 0; --i) {
      AlignFrag *f = (AlignFrag *)malloc(sizeof(AlignFrag));
      if (!f) { oom = true; break; }
      f->next = head; f->type = (FragType)fr[i].type; f->sa_start = fr[i].sa_start; f->sb_start = fr[i].sb_start; f->hsp_len = fr[i].hsp_len;
      head = f;
    }
    if (oom) { align_frag_free(head); *out = nullptr; rc = 1; break; }
    *out = alignment_new(head, score);
    if (!*out) align_frag_free(head);
    rc = 1;
  } while (0);
  if (rc < 0) cudaGetLastError();
  one_release(c);
  return rc;
}

static Alignment *legacy_align(int mode, const char *sa, int la, const char *sb, int lb, int alpha_len,
                               const unsigned char *map, int *matrix, int go, int ge, bool need_map) {
  // global_align.c:118-122, :241-243
  if (la <= 0 || lb <= 0 || !sa || !sb || !matrix || go > 0 || ge > 0) return nullptr;
  if (need_map && !map) return nullptr;
  if (alpha_len <= 0 || alpha_len > 256) return nullptr;
  if ((mode == ABV_GLOCAL && lb < 2) || (mode == ABV_OVERLAP && la < 2)) return nullptr;   // undefined in the reference
  Alignment *lean = nullptr;
  const int how = legacy_align_lean(mode, sa, la, sb, lb, alpha_len, need_map ? map : nullptr, matrix, go, ge, &lean);
  if (how == 1) return lean;
  if (how < 0) return nullptr;
  const long long a_off[2] = {0, la}, b_off[2] = {0, lb};
  const long long cap = (long long)la + lb + 2;
  std::vector<abv_frag> frags((size_t)cap);
  int score = 0, count = 0;
  long long start = 0, total = 0;
  int rc = abv_align_pairs(mode, sa, a_off, sb, b_off, 1, alpha_len, need_map ? map : nullptr, matrix, go, ge,
                           &score, &start, &count, frags.data(), cap, &total);
  if (rc != ABV_OK) return nullptr;
  AlignFrag *head = nullptr;
  for (int i = count - 1; i >= 0; --i) {
    AlignFrag *f = (AlignFrag *)malloc(sizeof(AlignFrag));
    if (!f) { align_frag_free(head); return nullptr; }
    const abv_frag &s = frags[(size_t)start + i];
    f->next = head; f->type = (FragType)s.type; f->sa_start = s.sa_start; f->sb_start = s.sb_start; f->hsp_len = s.hsp_len;
    head = f;
  }
  Alignment *al = alignment_new(head, score);
  if (!al) align_frag_free(head);
  return al;
}

extern "C" {

#define ABV_LEGACY(NAME, MODE)                                                                                     \
  Alignment *NAME##_raw(const unsigned char *sa, int sa_len, const unsigned char *sb, int sb_len, int alpha_len,   \
                        int *score_matrix, int gap_open, int gap_extend) {                                         \
    return legacy_align(MODE, (const char *)sa, sa_len, (const char *)sb, sb_len, alpha_len, nullptr, score_matrix, \
                        gap_open, gap_extend, false);                                                              \
  }                                                                                                                \
  Alignment *NAME(const char *seqa, int sa_len, const char *seqb, int sb_len, int alpha_len,                       \
                  const unsigned char *map, int *score_matrix, int gap_open, int gap_extend) {                     \
    return legacy_align(MODE, seqa, sa_len, seqb, sb_len, alpha_len, map, score_matrix, gap_open, gap_extend, true); \
  }

ABV_LEGACY(align, ABV_OVERLAP)
ABV_LEGACY(local_align, ABV_LOCAL)
ABV_LEGACY(global_align, ABV_GLOBAL)
ABV_LEGACY(glocal_align, ABV_GLOCAL)

}  // extern "C"
