/* lzb_host.cu -- host side of the C-ABI (include/lzb200.h): context, staging, batch entry points and
 * the one-call SDK functions.  Host code is plain C-style C++ over the CUDA runtime; no torch types.
 *
 * Data path of Lzb_DecodeUnitsHost (the call behind LzmaUncompressBatch and the parallel LZMA2 / xz
 * front ends): units are cut into slices; slice k goes H2D -> fused decode kernel -> D2H on stream
 * k % 3, so the copies of one slice overlap the kernels of its neighbours.
 */
#include "lzb_host.h"

#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <new>
#include <string>
#include <thread>

/* ------------------------------------------------------------------ errors ----------------- */
static thread_local char g_err[512] = "";

void lzb_set_error(const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

extern "C" const char *Lzb_LastError(void) { return g_err; }

/* No C++ exception leaves the C-ABI: an allocation that fails inside a batch entry point is SZ_ERROR_MEM for the caller. */
template <class F> static SRes lzb_guarded(F &&body)
{
  try { return body(); }
  catch (const std::bad_alloc &) { lzb_set_error("out of host memory"); return SZ_ERROR_MEM; }
  catch (...) { lzb_set_error("unexpected internal error"); return SZ_ERROR_FAIL; }
}

bool lzb_cuda_ok(cudaError_t e, const char *what)
{
  if (e == cudaSuccess) return true;
  lzb_set_error("%s: %s", what, cudaGetErrorString(e));
  return false;
}

/* ------------------------------------------------------------------ context ---------------- */
extern "C" LzbContext *Lzb_Create(int device)
{
  int count = 0;
  if (!lzb_cuda_ok(cudaGetDeviceCount(&count), "cudaGetDeviceCount") || count == 0)
  {
    if (count == 0) lzb_set_error("no CUDA device: liblzb200 has no CPU fallback");
    return nullptr;
  }
  if (device < 0 || device >= count) { lzb_set_error("device %d out of range (%d present)", device, count); return nullptr; }
  if (!lzb_cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return nullptr;
  LzbContext *c = new LzbContext();
  c->device = device;
  cudaDeviceGetAttribute(&c->sms, cudaDevAttrMultiProcessorCount, device);
  for (int i = 0; i < LZB_SLOTS; i++)
  {
    if (!lzb_cuda_ok(cudaStreamCreateWithFlags(&c->slot[i].stream, cudaStreamNonBlocking), "cudaStreamCreate") ||
        !lzb_cuda_ok(cudaMalloc(&c->slot[i].d_ticket, 256), "cudaMalloc ticket"))
    {
      Lzb_Destroy(c);
      return nullptr;
    }
  }
  if (!lzb_cuda_ok(cudaMalloc(&c->d_dev_tickets, 64 * sizeof(uint32_t)), "cudaMalloc tickets")) { Lzb_Destroy(c); return nullptr; }
  return c;
}

extern "C" void Lzb_Destroy(LzbContext *c)
{
  if (!c) return;
  cudaSetDevice(c->device);
  for (int i = 0; i < LZB_SLOTS; i++)
  {
    LzbSlot &s = c->slot[i];
    if (s.stream) { cudaStreamSynchronize(s.stream); cudaStreamDestroy(s.stream); }
    cudaFree(s.d_packed); cudaFree(s.d_out); cudaFree(s.d_units); cudaFree(s.d_results);
    cudaFree(s.d_ticket); cudaFree(s.d_slab);
    if (s.h_units) cudaFreeHost(s.h_units);
    if (s.h_results) cudaFreeHost(s.h_results);
  }
  cudaFree(c->d_dev_tickets); cudaFree(c->d_dev_slab);
  cudaFree(c->d_flow);
  if (c->h_flow) cudaFreeHost(c->h_flow);
  if (c->h_one) cudaFreeHost(c->h_one);
  if (c->flow_event) cudaEventDestroy(c->flow_event);
  if (c->dev_event) { cudaEventSynchronize(c->dev_event); cudaEventDestroy(c->dev_event); }
  delete c;
}

extern "C" int Lzb_DeviceSMs(LzbContext *c) { return c ? c->sms : 0; }

extern "C" void *Lzb_HostAlloc(size_t bytes)
{
  void *p = nullptr;
  if (!lzb_cuda_ok(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault), "cudaHostAlloc")) return nullptr;
  return p;
}
extern "C" void Lzb_HostFree(void *p) { if (p) cudaFreeHost(p); }

static std::mutex g_default_mu;
static LzbContext *g_default = nullptr;

LzbContext *lzb_default_context()
{
  std::lock_guard<std::mutex> g(g_default_mu);
  if (!g_default)
  {
    const char *e = getenv("LZB_DEVICE");
    g_default = Lzb_Create(e ? atoi(e) : 0);
  }
  return g_default;
}

template <class T> static bool grow(T *&p, size_t &cap, size_t need, const char *what)
{
  if (need <= cap) return true;
  cudaFree(p);
  p = nullptr; cap = 0;
  size_t want = need + need / 8 + 256;
  if (!lzb_cuda_ok(cudaMalloc((void **)&p, want * sizeof(T)), what)) return false;
  cap = want;
  return true;
}

bool lzb_slot_reserve(LzbContext *c, LzbSlot &s, size_t packed, size_t out, size_t units, size_t slab)
{
  (void)c;
  if (!grow(s.d_packed, s.packed_cap, packed + 64, "cudaMalloc packed arena")) return false;
  if (!grow(s.d_out, s.out_cap, out + 64, "cudaMalloc output arena")) return false;
  if (units > s.units_cap)
  {
    cudaFree(s.d_units); cudaFree(s.d_results);
    s.d_units = nullptr; s.d_results = nullptr; s.units_cap = 0;
    size_t want = units + units / 8 + 64;
    if (!lzb_cuda_ok(cudaMalloc((void **)&s.d_units, want * sizeof(LzbUnit)), "cudaMalloc units")) return false;
    if (!lzb_cuda_ok(cudaMalloc((void **)&s.d_results, want * sizeof(LzbResult)), "cudaMalloc results")) return false;
    s.units_cap = want;
  }
  if (units > s.h_units_cap)
  {
    if (s.h_units) cudaFreeHost(s.h_units);
    if (s.h_results) cudaFreeHost(s.h_results);
    s.h_units = nullptr; s.h_results = nullptr; s.h_units_cap = 0;
    size_t want = units + units / 8 + 64;
    if (!lzb_cuda_ok(cudaHostAlloc((void **)&s.h_units, want * sizeof(LzbUnit), cudaHostAllocDefault), "cudaHostAlloc units")) return false;
    if (!lzb_cuda_ok(cudaHostAlloc((void **)&s.h_results, want * sizeof(LzbResult), cudaHostAllocDefault), "cudaHostAlloc results")) return false;
    s.h_units_cap = want;
  }
  if (slab && !grow(s.d_slab, s.slab_cap, slab, "cudaMalloc model slab")) return false;
  return true;
}

/* probability-model bytes needed by a set of units (upper bound: monotone in lc+lp and pb).
 * LZMA2 units carry the walker's maxima in lc (= max lc+lp) and pb. */
struct ModelBytes { uint32_t smem, gmem; };

static ModelBytes model_bytes_of(unsigned lclp, unsigned pb)
{
  ModelBytes m;
  m.smem = lzb_smem_probs(lclp, 0, pb) * 2u;     /* fixed + pos-state tables + plain-literal trees */
  m.gmem = lzb_gmem_probs(lclp, 0) * 2u;         /* matched-literal trees                          */
  return m;
}

static ModelBytes model_bytes_for(const LzbUnit *u, size_t n)
{
  unsigned lclp = 0, pb = 0;
  for (size_t i = 0; i < n; i++)
  {
    unsigned a = (unsigned)u[i].lc + u[i].lp;
    if (a > lclp) lclp = a;
    if (u[i].pb > pb) pb = u[i].pb;
  }
  return model_bytes_of(lclp, pb);
}

static size_t slab_bytes_for(ModelBytes m)
{
  LzbGeometry g;
  if (lzb_launch_geometry(m.smem, m.gmem, &g) != 0) return 0;
  return g.slab_bytes;
}

/* ------------------------------------------------------------------ device-resident batch ---- */
extern "C" SRes Lzb_DecodeUnitsDevice(LzbContext *c, const Byte *d_packed, Byte *d_out, const LzbUnit *d_units,
                                      LzbResult *d_results, size_t n, unsigned maxLcLp, unsigned maxPb, void *stream)
{
  if (!c) { lzb_set_error("null context"); return SZ_ERROR_PARAM; }
  if (n > 0xFFFFFFFFull) return SZ_ERROR_PARAM;
  std::lock_guard<std::mutex> g(c->mu);
  if (!lzb_cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return SZ_ERROR_FAIL;
  const ModelBytes model = model_bytes_of(maxLcLp, maxPb);
  const size_t slab = slab_bytes_for(model);
  if (slab > c->dev_slab_cap && c->dev_event) cudaEventSynchronize(c->dev_event);     /* never free a slab a launch still uses */
  if (slab && !grow(c->d_dev_slab, c->dev_slab_cap, slab, "cudaMalloc model slab")) return SZ_ERROR_MEM;
  cudaStream_t st = (cudaStream_t)stream;        /* NULL = the CUDA default stream, as everywhere in CUDA */
  /* every launch of a context works in the context's ONE model slab: launches are ordered one behind the other, also
   * across the caller's streams (an event after each launch, waited for by the next) */
  if (!c->dev_event && !lzb_cuda_ok(cudaEventCreateWithFlags(&c->dev_event, cudaEventDisableTiming), "cudaEventCreate")) return SZ_ERROR_FAIL;
  if (c->dev_seq != 0 && c->dev_last_stream != st && !lzb_cuda_ok(cudaStreamWaitEvent(st, c->dev_event, 0), "order launches")) return SZ_ERROR_FAIL;
  uint32_t *ticket = c->d_dev_tickets + (c->dev_seq++ % 64);
  int rc = lzb_launch_decode(d_packed, d_out, d_units, d_results, (uint32_t)n, ticket, model.smem, model.gmem, c->d_dev_slab, st);
  if (rc != 0) { lzb_set_error("decode launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return SZ_ERROR_FAIL; }
  c->dev_last_stream = st;
  if (!lzb_cuda_ok(cudaEventRecord(c->dev_event, st), "cudaEventRecord")) return SZ_ERROR_FAIL;
  return SZ_OK;
}

/* ------------------------------------------------------------------ host-arena batch --------- */
struct Slice { size_t first, count; uint64_t src_lo, src_hi, dst_lo, dst_hi; };
static const size_t kExactTailUnits = 16;

/* A slice is the granule of H2D / kernel / D2H overlap.  It must still hold at least one unit per
 * resident warp (min_units), otherwise its kernel leaves the GPU partly idle. */
static void make_slices(const LzbUnit *u, size_t n, size_t min_units, std::vector<Slice> &out)
{
  uint64_t total = 0;
  for (size_t i = 0; i < n; i++) total += u[i].dst_cap;
  uint64_t target = total / 8;
  if (target < (64ull << 20)) target = 64ull << 20;
  if (target > (1ull << 30)) target = 1ull << 30;
  size_t i = 0;
  while (i < n)
  {
    Slice s;
    s.first = i; s.count = 0;
    s.src_lo = s.dst_lo = ~0ull; s.src_hi = s.dst_hi = 0;
    uint64_t acc = 0;
    while (i < n && (s.count < min_units || acc < target))
    {
      const LzbUnit &x = u[i];
      /* keep the slice's arena windows tight: stop when a unit jumps far away from the window */
      uint64_t nlo = std::min<uint64_t>(s.src_lo, x.src_off), nhi = std::max<uint64_t>(s.src_hi, x.src_off + x.src_len);
      uint64_t dlo = std::min<uint64_t>(s.dst_lo, x.dst_off), dhi = std::max<uint64_t>(s.dst_hi, x.dst_off + x.dst_cap);
      if (s.count != 0 && ((nhi - nlo) > 4 * target + (64ull << 20) || (dhi - dlo) > 4 * target + (64ull << 20))) break;
      s.src_lo = nlo; s.src_hi = nhi; s.dst_lo = dlo; s.dst_hi = dhi;
      acc += x.dst_cap;
      s.count++; i++;
    }
    out.push_back(s);
  }
}

/* ------------------------------------------------------------------ streamed launch ---------- */
/* ONE persistent launch for the whole batch, fed and drained while it runs: the packed arena goes up in chunks
 * on a copy stream (each chunk followed by its arrival flag), the kernel takes units in index order and waits
 * for a unit's chunk, and the group that completes a chunk raises a flag in mapped host memory, on which this
 * thread starts the chunk's download on a third stream.  Compared with sliced launches there is no per-slice
 * drain (a unit takes tens of milliseconds) and the copies of the first and last slice are not exposed. */
#define LZB_FLOW_STALLED (-1000)
static const uint64_t kFlowChunkBytes = 32ull << 20;
static const size_t kFlowMaxChunks = 8192;

static bool flow_reserve(LzbContext *c, size_t n_chunks)
{
  if (!c->h_one)
  {
    if (!lzb_cuda_ok(cudaHostAlloc((void **)&c->h_one, 64, cudaHostAllocDefault), "cudaHostAlloc flow constant")) return false;
    c->h_one[0] = 1;
  }
  if (!c->flow_event && !lzb_cuda_ok(cudaEventCreateWithFlags(&c->flow_event, cudaEventDisableTiming), "cudaEventCreate")) return false;
  if (n_chunks + 16 <= c->flow_cap) return true;
  cudaFree(c->d_flow); c->d_flow = nullptr;
  if (c->h_flow) { cudaFreeHost(c->h_flow); c->h_flow = nullptr; }
  c->flow_cap = 0;
  const size_t want = n_chunks + n_chunks / 4 + 64;
  if (!lzb_cuda_ok(cudaMalloc((void **)&c->d_flow, 2 * want * sizeof(uint32_t)), "cudaMalloc flow flags")) return false;
  if (!lzb_cuda_ok(cudaHostAlloc((void **)&c->h_flow, (want + 16) * sizeof(uint32_t), cudaHostAllocMapped), "cudaHostAlloc flow flags")) return false;
  if (!lzb_cuda_ok(cudaHostGetDevicePointer((void **)&c->h_flow_dev, c->h_flow, 0), "cudaHostGetDevicePointer")) return false;
  c->flow_cap = want;
  return true;
}

/* units in index order must walk both arenas forwards without overlap (what every front end here produces) */
static bool flow_layout_ok(const LzbUnit *u, size_t n)
{
  for (size_t i = 1; i < n; i++)
    if (u[i].src_off < u[i - 1].src_off + u[i - 1].src_len || u[i].dst_off < u[i - 1].dst_off + u[i - 1].dst_cap) return false;
  return true;
}

static SRes decode_units_flow(LzbContext *c, const Byte *packed, size_t packedBytes, Byte *out, size_t outBytes,
                              const LzbUnit *units, LzbResult *results, size_t n, const ModelBytes &model, size_t slab,
                              uint64_t total_out)
{
  const auto t_entry = std::chrono::steady_clock::now();
  size_t chunk_units = (size_t)std::max<uint64_t>(1, (uint64_t)n * kFlowChunkBytes / std::max<uint64_t>(total_out, 1));
  if (chunk_units > n) chunk_units = n;
  while ((n + chunk_units - 1) / chunk_units > kFlowMaxChunks) chunk_units *= 2;
  const size_t n_chunks = (n + chunk_units - 1) / chunk_units;
  LzbSlot &s = c->slot[0];
  cudaStream_t s_run = c->slot[0].stream, s_up = c->slot[1].stream, s_down = c->slot[2].stream;
  if (!lzb_slot_reserve(c, s, packedBytes, outBytes, n, slab) || !flow_reserve(c, n_chunks)) return SZ_ERROR_MEM;
  uint32_t *d_arrived = c->d_flow, *d_done = c->d_flow + c->flow_cap;
  volatile uint32_t *h_ready = c->h_flow;
  volatile uint32_t *h_abort = c->h_flow + c->flow_cap;
  for (size_t k = 0; k < n_chunks; k++) h_ready[k] = 0;
  *h_abort = 0;
  LzbFlow flow;
  flow.arrived = d_arrived; flow.done = d_done; flow.ready = c->h_flow_dev; flow.abort = c->h_flow_dev + c->flow_cap;
  flow.chunk_units = (uint32_t)chunk_units; flow.n_chunks = (uint32_t)n_chunks;

  bool ok = lzb_cuda_ok(cudaMemsetAsync(c->d_flow, 0, 2 * c->flow_cap * sizeof(uint32_t), s_run), "flow flags") &&
            lzb_cuda_ok(cudaEventRecord(c->flow_event, s_run), "flow event") &&
            lzb_cuda_ok(cudaStreamWaitEvent(s_up, c->flow_event, 0), "flow wait") &&
            lzb_cuda_ok(cudaMemcpyAsync(s.d_units, units, n * sizeof(LzbUnit), cudaMemcpyHostToDevice, s_run), "H2D units");
  if (!ok) return SZ_ERROR_FAIL;
  /* feed: chunk k's packed bytes, then its flag (stream order = arrival order).  Enqueued BEFORE the launch: a
   * launch call that blocks until the kernel ends (a profiler does that) must not leave the kernel waiting
   * for copies nobody has asked for yet */
  for (size_t k = 0; k < n_chunks && ok; k++)
  {
    const size_t a = k * chunk_units, b = std::min(n, a + chunk_units);
    const uint64_t lo = units[a].src_off, hi = units[b - 1].src_off + units[b - 1].src_len;
    if (hi > lo) ok = lzb_cuda_ok(cudaMemcpyAsync(s.d_packed + lo, packed + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, s_up), "H2D packed chunk");
    ok = ok && lzb_cuda_ok(cudaMemcpyAsync(d_arrived + k, c->h_one, sizeof(uint32_t), cudaMemcpyHostToDevice, s_up), "H2D arrival flag");
  }
  int rc = lzb_launch_decode_flow(s.d_packed, s.d_out, s.d_units, s.d_results, (uint32_t)n, s.d_ticket, model.smem, model.gmem,
                                  s.d_slab, &flow, s_run);
  if (!ok) { *h_abort = 1; }
  if (rc != 0) { lzb_set_error("decode launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return SZ_ERROR_FAIL; }
  /* drain: as chunks complete */
  static const bool trace = getenv("LZB_FLOW_TRACE") != nullptr;
  const auto t_start = std::chrono::steady_clock::now();
  auto ms_since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
  if (trace) fprintf(stderr, "[flow] %zu units, %zu chunks of %zu units; fed at %.2f ms\n", n, n_chunks, chunk_units, ms_since());
  SRes ret = ok ? SZ_OK : SZ_ERROR_FAIL;
  if (!ok) *h_abort = 1;
  bool up_done = false, stalled = false;
  if (ok) cudaEventRecord(c->flow_event, s_up);         /* (its first use, ordering the flag reset, is over) */
  for (size_t k = 0; k < n_chunks && ret == SZ_OK; k++)
  {
    unsigned idle = 0;
    while (h_ready[k] == 0)
    {
      if (!up_done && cudaEventQuery(c->flow_event) == cudaSuccess)
      {
        up_done = true;
        if (trace) fprintf(stderr, "[flow] uploads complete at %.2f ms\n", ms_since());
      }
      /* watchdog: copies that do not progress next to the running kernel (a profiler serialising the device,
       * a platform without concurrent copy) would leave every group waiting for its input: give up early, the
       * caller repeats the call with sliced launches and stops using streamed ones */
      if (k == 0 && !up_done && ms_since() > 2000.0) { stalled = true; ret = SZ_ERROR_FAIL; break; }
      if ((++idle & 1023u) == 0)
      {
        /* the kernel ending (or failing) without the flag means something went wrong on the device */
        const cudaError_t q = cudaStreamQuery(s_run);
        if (q != cudaErrorNotReady && h_ready[k] == 0)
        {
          if (q != cudaSuccess) lzb_cuda_ok(q, "streamed decode launch"); else lzb_set_error("streamed launch ended before chunk %zu was complete", k);
          ret = SZ_ERROR_FAIL;
          break;
        }
      }
      std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    if (ret != SZ_OK) break;
    if (trace && (k % 16 == 0 || k + 1 == n_chunks)) fprintf(stderr, "[flow] chunk %zu ready at %.2f ms\n", k, ms_since());
    const size_t a = k * chunk_units, b = std::min(n, a + chunk_units);
    size_t j = a;
    while (j < b && ok)
    {
      uint64_t lo = units[j].dst_off, hi = lo + units[j].dst_cap;
      size_t e = j + 1;
      while (e < b && units[e].dst_off == hi) { hi += units[e].dst_cap; e++; }
      if (hi > lo) ok = lzb_cuda_ok(cudaMemcpyAsync(out + lo, s.d_out + lo, (size_t)(hi - lo), cudaMemcpyDeviceToHost, s_down), "D2H output chunk");
      j = e;
    }
    if (!ok) ret = SZ_ERROR_FAIL;
  }
  if (ret != SZ_OK) *h_abort = 1;
  ok = lzb_cuda_ok(cudaMemcpyAsync(s.h_results, s.d_results, n * sizeof(LzbResult), cudaMemcpyDeviceToHost, s_run), "D2H results");
  ok = lzb_cuda_ok(cudaStreamSynchronize(s_up), "upload stream") && ok;
  if (trace) fprintf(stderr, "[flow] upload stream idle at %.2f ms\n", ms_since());
  ok = lzb_cuda_ok(cudaStreamSynchronize(s_run), "streamed decode launch") && ok;
  if (trace) fprintf(stderr, "[flow] kernel done at %.2f ms\n", ms_since());
  ok = lzb_cuda_ok(cudaStreamSynchronize(s_down), "download stream") && ok;
  if (trace) fprintf(stderr, "[flow] download stream idle at %.2f ms\n", ms_since());
  if (!ok) return SZ_ERROR_FAIL;
  if (stalled) { lzb_set_error("streamed launch: uploads do not overlap the kernel here"); return LZB_FLOW_STALLED; }
  memcpy(results, s.h_results, n * sizeof(LzbResult));
  if (trace) fprintf(stderr, "[flow] call took %.2f ms (%.2f ms before the first poll)\n",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_entry).count(),
                     std::chrono::duration<double, std::milli>(t_start - t_entry).count());
  return ret;
}

/* may this batch take the streamed launch?  (caller holds the context lock) */
static bool flow_eligible(LzbContext *c, const Byte *packed, size_t packedBytes, const Byte *out, size_t outBytes,
                          const LzbUnit *units, size_t n, uint64_t total_out)
{
  size_t free_b = 0, total_b = 0;
  static const char *const mode = getenv("LZB_FLOW");
  const bool want = !c->flow_off && (mode ? atoi(mode) != 0 : total_out >= (256ull << 20));
  /* page-locked arenas only: copies from pageable memory block the calling thread, which must stay free to
   * feed and drain the launch */
  auto pinned = [](const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
  };
  if (!(want && n != 0 && n < 0xFFFFFFFFull && pinned(packed) && pinned(out) && flow_layout_ok(units, n) &&
        cudaMemGetInfo(&free_b, &total_b) == cudaSuccess))
    return false;
  const LzbSlot &s0 = c->slot[0];
  const uint64_t have = (uint64_t)free_b + s0.packed_cap + s0.out_cap;
  return (uint64_t)packedBytes + outBytes + (1ull << 30) < have;
}

/* 7z: the decode pass as a streamed launch (outputs go down while later folders still decode), then a
 * check-only pass over the output that is still resident.  *used = false: not eligible, nothing was done. */
SRes lzb_flow_then_checks(LzbContext *c, const Byte *packed, size_t packedBytes, Byte *out, size_t outBytes,
                          const LzbUnit *units, LzbResult *results, size_t n,
                          const LzbUnit *checks, LzbResult *check_results, size_t n_checks, bool *used)
{
  *used = false;
  std::lock_guard<std::mutex> g(c->mu);
  if (!lzb_cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return SZ_ERROR_FAIL;
  uint64_t total_out = 0;
  for (size_t i = 0; i < n; i++) total_out += units[i].dst_cap;
  if (!flow_eligible(c, packed, packedBytes, out, outBytes, units, n, total_out)) return SZ_OK;
  const ModelBytes model = model_bytes_for(units, n);
  const SRes r = decode_units_flow(c, packed, packedBytes, out, outBytes, units, results, n, model, slab_bytes_for(model), total_out);
  if (r == LZB_FLOW_STALLED) { c->flow_off = true; return SZ_OK; }
  *used = true;
  if (r != SZ_OK || n_checks == 0) return r;
  LzbSlot &s = c->slot[0];
  if (!lzb_slot_reserve(c, s, 0, 0, std::max(n, n_checks), 0)) return SZ_ERROR_MEM;
  if (!lzb_cuda_ok(cudaMemcpyAsync(s.d_units, checks, n_checks * sizeof(LzbUnit), cudaMemcpyHostToDevice, s.stream), "H2D check units")) return SZ_ERROR_FAIL;
  const int rc = lzb_launch_post_only(s.d_out, s.d_units, s.d_results, (uint32_t)n_checks, s.stream);
  if (rc != 0) { lzb_set_error("post launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return SZ_ERROR_FAIL; }
  if (!lzb_cuda_ok(cudaMemcpyAsync(check_results, s.d_results, n_checks * sizeof(LzbResult), cudaMemcpyDeviceToHost, s.stream), "D2H check results") ||
      !lzb_cuda_ok(cudaStreamSynchronize(s.stream), "check pass")) return SZ_ERROR_FAIL;
  return SZ_OK;
}

static SRes Lzb_DecodeUnitsHost_body(LzbContext *c, const Byte *packed, size_t packedBytes, Byte *out, size_t outBytes,
                                    const LzbUnit *units, LzbResult *results, size_t n)
{
  if (!c) { lzb_set_error("null context"); return SZ_ERROR_PARAM; }
  if (n == 0) return SZ_OK;
  uint64_t total_out = 0;
  for (size_t i = 0; i < n; i++)
  {
    if (units[i].src_len > packedBytes || units[i].src_off > packedBytes - units[i].src_len ||
        units[i].dst_cap > outBytes || units[i].dst_off > outBytes - units[i].dst_cap)
    { lzb_set_error("unit %zu exceeds its arena", i); return SZ_ERROR_PARAM; }
    total_out += units[i].dst_cap;
  }
  std::lock_guard<std::mutex> g(c->mu);
  if (!lzb_cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return SZ_ERROR_FAIL;

  const ModelBytes model = model_bytes_for(units, n);
  const size_t slab = slab_bytes_for(model);
  /* big batches whose arenas fit in device memory next to each other: one streamed launch */
  if (flow_eligible(c, packed, packedBytes, out, outBytes, units, n, total_out))
  {
    const SRes r = decode_units_flow(c, packed, packedBytes, out, outBytes, units, results, n, model, slab, total_out);
    if (r != LZB_FLOW_STALLED) return r;
    c->flow_off = true;                                  /* sticky: sliced launches from here on */
  }
  std::vector<Slice> slices;
  {
    LzbGeometry geo;
    geo.warps_per_cta = 1; geo.units_per_cta = 1; geo.ctas = 1;
    lzb_launch_geometry(model.smem, model.gmem, &geo);
    /* every slice kernel ends with a tail of about one unit time during which the SMs drain (CTAs are
     * persistent); four waves per slice keep that under ~10 % while copies still overlap the kernels */
    const uint32_t w = geo.units_per_cta * 4, ctas = geo.ctas;
    make_slices(units, n, (size_t)w * ctas, slices);
    /* a short last slice would run alone on an almost empty GPU: merge it into its predecessor */
    if (slices.size() >= 2 && slices.back().count < (size_t)w * ctas)
    {
      Slice tail = slices.back();
      slices.pop_back();
      Slice &p = slices.back();
      p.count += tail.count;
      p.src_lo = std::min(p.src_lo, tail.src_lo); p.src_hi = std::max(p.src_hi, tail.src_hi);
      p.dst_lo = std::min(p.dst_lo, tail.dst_lo); p.dst_hi = std::max(p.dst_hi, tail.dst_hi);
    }
  }

  SRes ret = SZ_OK;
  std::vector<int> pending_slot(LZB_SLOTS, -1);          /* slice index in flight on each slot */

  auto drain = [&](int si) -> bool
  {
    int k = pending_slot[si];
    if (k < 0) return true;
    LzbSlot &s = c->slot[si];
    if (!lzb_cuda_ok(cudaStreamSynchronize(s.stream), "decode slice")) return false;
    memcpy(results + slices[k].first, s.h_results, slices[k].count * sizeof(LzbResult));
    pending_slot[si] = -1;
    return true;
  };

  for (size_t k = 0; k < slices.size() && ret == SZ_OK; k++)
  {
    const int si = (int)(k % LZB_SLOTS);
    LzbSlot &s = c->slot[si];
    const Slice &sl = slices[k];
    if (!drain(si)) { ret = SZ_ERROR_FAIL; break; }
    const size_t pbytes = (size_t)(sl.src_hi - sl.src_lo), obytes = (size_t)(sl.dst_hi - sl.dst_lo);
    if (!lzb_slot_reserve(c, s, pbytes, obytes, sl.count, slab)) { ret = SZ_ERROR_MEM; break; }
    for (size_t j = 0; j < sl.count; j++)
    {
      s.h_units[j] = units[sl.first + j];
      s.h_units[j].src_off -= sl.src_lo;
      s.h_units[j].dst_off -= sl.dst_lo;
    }
    bool ok = lzb_cuda_ok(cudaMemcpyAsync(s.d_packed, packed + sl.src_lo, pbytes, cudaMemcpyHostToDevice, s.stream), "H2D packed") &&
              lzb_cuda_ok(cudaMemcpyAsync(s.d_units, s.h_units, sl.count * sizeof(LzbUnit), cudaMemcpyHostToDevice, s.stream), "H2D units");
    if (!ok) { ret = SZ_ERROR_FAIL; break; }
    int rc = lzb_launch_decode(s.d_packed, s.d_out, s.d_units, s.d_results, (uint32_t)sl.count, s.d_ticket, model.smem, model.gmem, s.d_slab, s.stream);
    if (rc != 0) { lzb_set_error("decode launch failed: %s", cudaGetErrorString((cudaError_t)rc)); ret = SZ_ERROR_FAIL; break; }
    if (sl.count <= kExactTailUnits)
    {
      /* a handful of units (the one-call SDK functions): fetch the results first and bring back exactly the bytes
       * each unit produced -- what lies behind *destLen in the caller's buffer is not touched, and nothing that an
       * earlier call left in the device arena can travel to this caller */
      ok = lzb_cuda_ok(cudaMemcpyAsync(s.h_results, s.d_results, sl.count * sizeof(LzbResult), cudaMemcpyDeviceToHost, s.stream), "D2H results") &&
           lzb_cuda_ok(cudaStreamSynchronize(s.stream), "decode slice");
      for (size_t j = 0; j < sl.count && ok; j++)
      {
        const LzbUnit &u = units[sl.first + j];
        const size_t made = std::min<size_t>(s.h_results[j].dst_done, u.dst_cap);
        if (made) ok = lzb_cuda_ok(cudaMemcpyAsync(out + u.dst_off, s.d_out + (u.dst_off - sl.dst_lo), made, cudaMemcpyDeviceToHost, s.stream), "D2H output");
      }
      if (!ok) { ret = SZ_ERROR_FAIL; break; }
      pending_slot[si] = (int)k;
      continue;
    }
    /* D2H: one copy per maximal run of units whose outputs tile the arena (never touch the gaps).  Batches copy whole
     * capacities: waiting for the results first would serialise the download behind the kernel */
    size_t j = 0;
    while (j < sl.count && ok)
    {
      uint64_t lo = units[sl.first + j].dst_off, hi = lo + units[sl.first + j].dst_cap;
      size_t e = j + 1;
      while (e < sl.count && units[sl.first + e].dst_off == hi) { hi += units[sl.first + e].dst_cap; e++; }
      if (hi > lo)
        ok = lzb_cuda_ok(cudaMemcpyAsync(out + lo, s.d_out + (lo - sl.dst_lo), (size_t)(hi - lo), cudaMemcpyDeviceToHost, s.stream), "D2H output");
      j = e;
    }
    ok = ok && lzb_cuda_ok(cudaMemcpyAsync(s.h_results, s.d_results, sl.count * sizeof(LzbResult), cudaMemcpyDeviceToHost, s.stream), "D2H results");
    if (!ok) { ret = SZ_ERROR_FAIL; break; }
    pending_slot[si] = (int)k;
  }
  for (int si = 0; si < LZB_SLOTS; si++)
    if (!drain(si)) ret = SZ_ERROR_FAIL;
  return ret;
}

/* ------------------------------------------------------------------ several GPUs of one box --- */
extern "C" SRes Lzb_DecodeUnitsHost(LzbContext *c, const Byte *packed, size_t packedBytes, Byte *out, size_t outBytes,
                                    const LzbUnit *units, LzbResult *results, size_t n)
{
  return lzb_guarded([&]() { return Lzb_DecodeUnitsHost_body(c, packed, packedBytes, out, outBytes, units, results, n); });
}

extern "C" void Lzb_ShardBounds(const LzbUnit *units, size_t n, size_t world, size_t *bounds)
{
  if (world == 0) return;
  uint64_t total = 0;
  for (size_t i = 0; i < n; i++) total += units[i].dst_cap;
  bounds[0] = 0;
  size_t i = 0;
  uint64_t acc = 0;
  for (size_t r = 1; r < world; r++)
  {
    /* the first unit whose cumulative size reaches r / world of the total closes range r - 1 */
    const double target = (double)total * (double)r / (double)world;
    while (i < n && (double)acc < target) acc += units[i++].dst_cap;
    bounds[r] = i;
  }
  bounds[world] = n;
}

static SRes Lzb_DecodeUnitsHostMulti_body(LzbContext *const *ctxs, size_t nCtx, const Byte *packed, size_t packedBytes, Byte *out,
                                         size_t outBytes, const LzbUnit *units, LzbResult *results, size_t n)
{
  if (nCtx == 0 || !ctxs) { lzb_set_error("no context"); return SZ_ERROR_PARAM; }
  if (nCtx == 1) return Lzb_DecodeUnitsHost(ctxs[0], packed, packedBytes, out, outBytes, units, results, n);
  std::vector<size_t> bounds(nCtx + 1);
  Lzb_ShardBounds(units, n, nCtx, bounds.data());
  std::vector<SRes> verdict(nCtx, SZ_OK);
  std::vector<std::string> errors(nCtx);
  std::vector<std::thread> workers;
  workers.reserve(nCtx);
  bool spawned = true;
  /* a range is failed until its worker says otherwise (a worker that never started leaves it that way) */
  for (size_t r = 0; r < nCtx; r++)
    if (bounds[r] != bounds[r + 1]) { verdict[r] = SZ_ERROR_MEM; errors[r] = "the range's worker did not run (out of memory or threads)"; }
  for (size_t r = 0; r < nCtx && spawned; r++)
  {
    auto work = [&, r]()
    {
      const size_t lo = bounds[r], hi = bounds[r + 1];
      if (lo == hi) return;
      /* this range's windows into the two arenas */
      uint64_t s_lo = ~0ull, s_hi = 0, d_lo = ~0ull, d_hi = 0;
      for (size_t i = lo; i < hi; i++)
      {
        s_lo = std::min<uint64_t>(s_lo, units[i].src_off); s_hi = std::max<uint64_t>(s_hi, units[i].src_off + units[i].src_len);
        d_lo = std::min<uint64_t>(d_lo, units[i].dst_off); d_hi = std::max<uint64_t>(d_hi, units[i].dst_off + units[i].dst_cap);
      }
      if (s_hi > packedBytes || d_hi > outBytes) { verdict[r] = SZ_ERROR_PARAM; errors[r] = "a unit exceeds its arena"; return; }
      std::vector<LzbUnit> mine(units + lo, units + hi);
      for (LzbUnit &u : mine) { u.src_off -= s_lo; u.dst_off -= d_lo; }
      verdict[r] = Lzb_DecodeUnitsHost(ctxs[r], packed + s_lo, (size_t)(s_hi - s_lo), out + d_lo, (size_t)(d_hi - d_lo),
                                       mine.data(), results + lo, hi - lo);
      if (verdict[r] != SZ_OK) errors[r] = Lzb_LastError();          /* (the error text is per thread) */
    };
    /* a range that cannot get its thread or its copy of the table fails alone; the started ones are always joined */
    try { workers.emplace_back([&verdict, r, work]() { try { work(); } catch (...) { verdict[r] = SZ_ERROR_MEM; } }); }
    catch (...) { verdict[r] = SZ_ERROR_MEM; spawned = false; }
  }
  for (std::thread &t : workers) t.join();
  SRes worst = SZ_OK;
  for (size_t r = 0; r < nCtx; r++)
    if (verdict[r] != SZ_OK && worst == SZ_OK) { worst = verdict[r]; lzb_set_error("GPU %zu of %zu: %s", r, nCtx, errors[r].c_str()); }
  return worst;
}

/* Run a few units whose arenas the caller already placed in device memory; unit table and results are
 * host arrays.  Synchronous.  Used by the stateful SDK objects (lzb_state.cu). */
SRes lzb_run_resident(LzbContext *c, const uint8_t *d_packed, uint8_t *d_out, const LzbUnit *units, LzbResult *results, size_t n)
{
  std::lock_guard<std::mutex> g(c->mu);
  if (!lzb_cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return SZ_ERROR_FAIL;
  LzbSlot &s = c->slot[0];
  const ModelBytes model = model_bytes_for(units, n);
  if (!lzb_slot_reserve(c, s, 0, 0, n, slab_bytes_for(model))) return SZ_ERROR_MEM;
  if (!lzb_cuda_ok(cudaMemcpyAsync(s.d_units, units, n * sizeof(LzbUnit), cudaMemcpyHostToDevice, s.stream), "H2D units")) return SZ_ERROR_FAIL;
  int rc = lzb_launch_decode(d_packed, d_out, s.d_units, s.d_results, (uint32_t)n, s.d_ticket, model.smem, model.gmem, s.d_slab, s.stream);
  if (rc != 0) { lzb_set_error("decode launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return SZ_ERROR_FAIL; }
  if (!lzb_cuda_ok(cudaMemcpyAsync(results, s.d_results, n * sizeof(LzbResult), cudaMemcpyDeviceToHost, s.stream), "D2H results") ||
      !lzb_cuda_ok(cudaStreamSynchronize(s.stream), "decode kernel")) return SZ_ERROR_FAIL;
  return SZ_OK;
}

/* One call of a resumable decoder object (lzb_state.cu).  Everything the call needs travels in two small uploads
 * from page-locked memory -- the scalars of the object (-> the header of its device-resident state) and one block
 * holding [LzbUnit | LzbResult | parked bytes + new input] -- then the launch, then the scalars, the result and the
 * newly written part of the dictionary mirror come back, all queued on one stream with a single wait at the end.
 * The size of the new dictionary part is only known afterwards: `guess` bytes are fetched with the rest of the
 * queue, the caller fetches what lies beyond (rare) once it has the scalars. */
SRes lzb_run_stateful(LzbContext *c, uint8_t *d_block, const uint8_t *h_block, size_t block_bytes, const LzbUnit *unit,
                      uint8_t *d_state, LzbStreamState *h_state, LzbResult *h_result,
                      const uint8_t *d_dic, Byte *dic, size_t dic_from, size_t guess)
{
  std::lock_guard<std::mutex> g(c->mu);
  if (!lzb_cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return SZ_ERROR_FAIL;
  LzbSlot &s = c->slot[0];
  const ModelBytes model = model_bytes_for(unit, 1);
  if (!lzb_slot_reserve(c, s, 0, 0, 1, slab_bytes_for(model))) return SZ_ERROR_MEM;
  const LzbUnit *d_unit = reinterpret_cast<const LzbUnit *>(d_block);
  LzbResult *d_result = reinterpret_cast<LzbResult *>(d_block + LZB_STATEFUL_RESULT_AT);
  bool ok = lzb_cuda_ok(cudaMemcpyAsync(d_state, h_state, sizeof(LzbStreamState), cudaMemcpyHostToDevice, s.stream), "H2D decoder scalars") &&
            lzb_cuda_ok(cudaMemcpyAsync(d_block, h_block, block_bytes, cudaMemcpyHostToDevice, s.stream), "H2D unit + input");
  if (!ok) return SZ_ERROR_FAIL;
  const int rc = lzb_launch_decode(d_block + LZB_STATEFUL_INPUT_AT, const_cast<uint8_t *>(d_dic), d_unit, d_result, 1, s.d_ticket,
                                   model.smem, model.gmem, s.d_slab, s.stream);
  if (rc != 0) { lzb_set_error("decode launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return SZ_ERROR_FAIL; }
  ok = lzb_cuda_ok(cudaMemcpyAsync(h_state, d_state, sizeof(LzbStreamState), cudaMemcpyDeviceToHost, s.stream), "D2H decoder scalars") &&
       lzb_cuda_ok(cudaMemcpyAsync(h_result, d_result, sizeof(LzbResult), cudaMemcpyDeviceToHost, s.stream), "D2H result");
  if (ok && guess) ok = lzb_cuda_ok(cudaMemcpyAsync(dic + dic_from, d_dic + dic_from, guess, cudaMemcpyDeviceToHost, s.stream), "D2H dictionary");
  return (ok && lzb_cuda_ok(cudaStreamSynchronize(s.stream), "stateful decode")) ? SZ_OK : SZ_ERROR_FAIL;
}

/* One archive, several dependent passes over the same device-resident arenas (7z folder path): upload the
 * packed bytes once, device-to-device copies for stored ranges, the fused decode pass, then up to two
 * filter/check-only passes (folder level, then file level), one download of the output at the end. */
SRes lzb_run_stages(LzbContext *c, const Byte *packed, size_t packedBytes, Byte *out, size_t outBytes, size_t devOutBytes,
                    const LzbCopyRange *copies, size_t n_copies,
                    const LzbUnit *units, LzbResult *results, size_t n,
                    LzbBcj2Job *bcj2, size_t n_bcj2,
                    const LzbUnit *post_a, LzbResult *res_a, size_t n_a,
                    const LzbUnit *post_b, LzbResult *res_b, size_t n_b)
{
  if (devOutBytes < outBytes) devOutBytes = outBytes;
  /* everything below is indexed by offsets that came out of an archive: nothing may leave the two device arenas */
  {
    auto inside = [](uint64_t off, uint64_t len, uint64_t arena) { return len <= arena && off <= arena - len; };
    bool ok = true;
    for (size_t i = 0; ok && i < n_copies; i++) ok = inside(copies[i].src_off, copies[i].len, packedBytes) && inside(copies[i].dst_off, copies[i].len, devOutBytes);
    for (size_t i = 0; ok && i < n; i++) ok = inside(units[i].src_off, units[i].src_len, packedBytes) && inside(units[i].dst_off, units[i].dst_cap, devOutBytes);
    for (size_t i = 0; ok && i < n_a; i++) ok = inside(post_a[i].dst_off, post_a[i].dst_cap, devOutBytes);
    for (size_t i = 0; ok && i < n_b; i++) ok = inside(post_b[i].dst_off, post_b[i].dst_cap, devOutBytes);
    for (size_t i = 0; ok && i < n_bcj2; i++)
      ok = inside(bcj2[i].main_off, bcj2[i].main_size, devOutBytes) && inside(bcj2[i].call_off, bcj2[i].call_size, devOutBytes) &&
           inside(bcj2[i].jump_off, bcj2[i].jump_size, devOutBytes) && inside(bcj2[i].out_off, bcj2[i].out_size, devOutBytes) &&
           inside(bcj2[i].rc_off, bcj2[i].rc_size, packedBytes);
    if (!ok) { lzb_set_error("a range of the job lies outside its arena"); return SZ_ERROR_FAIL; }
  }
  std::lock_guard<std::mutex> g(c->mu);
  if (!lzb_cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return SZ_ERROR_FAIL;
  LzbSlot &s = c->slot[0];
  const ModelBytes model = model_bytes_for(units, n);
  const size_t most = std::max(n, std::max(n_a, n_b));
  const size_t jobs_at = (packedBytes + 16 + 15) & ~(size_t)15;             /* the BCJ2 job table rides behind the archive */
  if (!lzb_slot_reserve(c, s, jobs_at + n_bcj2 * sizeof(LzbBcj2Job) + 16, devOutBytes + 16, most ? most : 1, n ? slab_bytes_for(model) : 0)) return SZ_ERROR_MEM;
  bool ok = lzb_cuda_ok(cudaMemcpyAsync(s.d_packed, packed, packedBytes, cudaMemcpyHostToDevice, s.stream), "H2D archive");
  for (size_t i = 0; ok && i < n_copies; i++)
    ok = lzb_cuda_ok(cudaMemcpyAsync(s.d_out + copies[i].dst_off, s.d_packed + copies[i].src_off, copies[i].len,
                                     cudaMemcpyDeviceToDevice, s.stream), "stored folder copy");
  if (!ok) return SZ_ERROR_FAIL;
  if (n)
  {
    if (!lzb_cuda_ok(cudaMemcpyAsync(s.d_units, units, n * sizeof(LzbUnit), cudaMemcpyHostToDevice, s.stream), "H2D units")) return SZ_ERROR_FAIL;
    int rc = lzb_launch_decode(s.d_packed, s.d_out, s.d_units, s.d_results, (uint32_t)n, s.d_ticket, model.smem, model.gmem, s.d_slab, s.stream);
    if (rc != 0) { lzb_set_error("decode launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return SZ_ERROR_FAIL; }
    if (!lzb_cuda_ok(cudaMemcpyAsync(results, s.d_results, n * sizeof(LzbResult), cudaMemcpyDeviceToHost, s.stream), "D2H results")) return SZ_ERROR_FAIL;
  }
  if (n_bcj2)
  {
    LzbBcj2Job *d_jobs = reinterpret_cast<LzbBcj2Job *>(s.d_packed + jobs_at);
    if (!lzb_cuda_ok(cudaMemcpyAsync(d_jobs, bcj2, n_bcj2 * sizeof(LzbBcj2Job), cudaMemcpyHostToDevice, s.stream), "H2D bcj2 jobs")) return SZ_ERROR_FAIL;
    int rc = lzb_launch_bcj2(s.d_packed, s.d_out, d_jobs, (uint32_t)n_bcj2, s.stream);
    if (rc != 0) { lzb_set_error("bcj2 launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return SZ_ERROR_FAIL; }
    if (!lzb_cuda_ok(cudaMemcpyAsync(bcj2, d_jobs, n_bcj2 * sizeof(LzbBcj2Job), cudaMemcpyDeviceToHost, s.stream), "D2H bcj2 jobs")) return SZ_ERROR_FAIL;
  }
  const LzbUnit *pu[2] = { post_a, post_b };
  LzbResult *pr[2] = { res_a, res_b };
  const size_t pn[2] = { n_a, n_b };
  for (int k = 0; k < 2; k++)
  {
    if (!pn[k]) continue;
    if (!lzb_cuda_ok(cudaMemcpyAsync(s.d_units, pu[k], pn[k] * sizeof(LzbUnit), cudaMemcpyHostToDevice, s.stream), "H2D post units")) return SZ_ERROR_FAIL;
    int rc = lzb_launch_post_only(s.d_out, s.d_units, s.d_results, (uint32_t)pn[k], s.stream);
    if (rc != 0) { lzb_set_error("post launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return SZ_ERROR_FAIL; }
    if (!lzb_cuda_ok(cudaMemcpyAsync(pr[k], s.d_results, pn[k] * sizeof(LzbResult), cudaMemcpyDeviceToHost, s.stream), "D2H post results")) return SZ_ERROR_FAIL;
  }
  if (outBytes && !lzb_cuda_ok(cudaMemcpyAsync(out, s.d_out, outBytes, cudaMemcpyDeviceToHost, s.stream), "D2H output")) return SZ_ERROR_FAIL;
  if (!lzb_cuda_ok(cudaStreamSynchronize(s.stream), "7z passes")) return SZ_ERROR_FAIL;
  return SZ_OK;
}

/* ------------------------------------------------------------------ props ------------------- */
extern "C" SRes Lzb_DecodeUnitsHostMulti(LzbContext *const *ctxs, size_t nCtx, const Byte *packed, size_t packedBytes, Byte *out,
                                         size_t outBytes, const LzbUnit *units, LzbResult *results, size_t n)
{
  return lzb_guarded([&]() { return Lzb_DecodeUnitsHostMulti_body(ctxs, nCtx, packed, packedBytes, out, outBytes, units, results, n); });
}

extern "C" SRes LzmaProps_Decode(CLzmaProps *p, const Byte *data, unsigned size)
{
  /* reference LzmaDec.c:898-922: dictionary clamped up to 4 KiB, d < 225 */
  if (size < LZMA_PROPS_SIZE) return SZ_ERROR_UNSUPPORTED;
  UInt32 dict = lzb_le32(data + 1);
  p->dicSize = dict < (1u << 12) ? (1u << 12) : dict;
  unsigned d = data[0];
  if (d >= 9 * 5 * 5) return SZ_ERROR_UNSUPPORTED;
  p->lc = d % 9; d /= 9;
  p->pb = d / 5;
  p->lp = d % 5;
  return SZ_OK;
}

/* ------------------------------------------------------------------ callers that do not batch -- */
/* LzmaUncompress / LzmaDecode decode ONE stream per call; a stream alone occupies one warp of one SM for its whole
 * decode (a 64 KiB stream: ~15 ms), so a host that calls from many threads -- the GPU analogue of the reference's
 * one-unit-per-thread use -- would queue up behind the context lock, one launch each.  Instead concurrent small
 * calls are COMBINED: a caller leaves its request in a list; whoever finds no launch in flight becomes the leader,
 * takes everything in the list (its own request included), runs it as one batch through Lzb_DecodeUnitsHost and
 * hands the results out; callers that arrive meanwhile form the next batch.  One caller alone pays two extra host
 * copies of its bytes; N concurrent callers finish in about the time of one.  Still no CPU fallback. */
namespace {

const size_t kCombineMaxBytes = (size_t)8 << 20;      /* larger calls go straight to Lzb_DecodeUnitsHost */

struct OneCall
{
  const Byte *src = nullptr;
  Byte *dest = nullptr;
  LzbUnit unit;
  LzbResult res;
  SRes call_res = SZ_OK;
  bool done = false;
  std::condition_variable cv;
};

struct Combiner
{
  std::mutex mu;
  std::vector<OneCall *> waiting;
  bool busy = false;                                  /* a leader is running a batch */
  Byte *h_in = nullptr, *h_out = nullptr;             /* page-locked staging arenas, grown on demand */
  size_t in_cap = 0, out_cap = 0;
  std::vector<LzbUnit> units;
  std::vector<LzbResult> results;
} g_combiner;

bool stage_reserve(Byte *&p, size_t &cap, size_t need)
{
  if (need <= cap) return true;
  Lzb_HostFree(p);
  p = (Byte *)Lzb_HostAlloc(need + need / 2 + 4096);
  cap = p ? need + need / 2 + 4096 : 0;
  return p != nullptr;
}

void run_batch(LzbContext *c, std::vector<OneCall *> &batch)
{
  Combiner &q = g_combiner;                           /* staging is the leader's while `busy` is set */
  const size_t n = batch.size();
  q.units.resize(n); q.results.resize(n);
  size_t in_bytes = 0, out_bytes = 0;
  for (size_t i = 0; i < n; i++)
  {
    LzbUnit &u = q.units[i];
    u = batch[i]->unit;
    u.src_off = in_bytes; u.dst_off = out_bytes;
    in_bytes += ((size_t)u.src_len + 15) & ~(size_t)15;
    out_bytes += ((size_t)u.dst_cap + 15) & ~(size_t)15;
  }
  SRes r = SZ_OK;
  if (!stage_reserve(q.h_in, q.in_cap, in_bytes + 64) || !stage_reserve(q.h_out, q.out_cap, out_bytes + 64)) r = SZ_ERROR_MEM;
  if (r == SZ_OK)
  {
    for (size_t i = 0; i < n; i++) memcpy(q.h_in + q.units[i].src_off, batch[i]->src, q.units[i].src_len);
    r = Lzb_DecodeUnitsHost(c, q.h_in, in_bytes + 64, q.h_out, out_bytes + 64, q.units.data(), q.results.data(), n);
  }
  for (size_t i = 0; i < n; i++)
  {
    OneCall &o = *batch[i];
    o.call_res = r;
    if (r != SZ_OK) continue;
    o.res = q.results[i];
    const size_t made = std::min<size_t>(o.res.dst_done, q.units[i].dst_cap);
    if (made) memcpy(o.dest, q.h_out + q.units[i].dst_off, made);
  }
}

SRes combined_decode(LzbContext *c, OneCall &me)
{
  Combiner &q = g_combiner;
  std::unique_lock<std::mutex> lk(q.mu);
  try { q.waiting.push_back(&me); }
  catch (...) { return SZ_ERROR_MEM; }
  while (!me.done && q.busy) me.cv.wait(lk);
  if (me.done) return me.call_res;                    /* a leader served this call */
  q.busy = true;                                      /* lead: everything that is waiting now, this call included */
  std::vector<OneCall *> batch;
  batch.swap(q.waiting);
  lk.unlock();
  SRes failed = SZ_OK;                                /* nobody of the batch may be left waiting, whatever happens */
  try { run_batch(c, batch); }
  catch (...) { failed = SZ_ERROR_MEM; }
  lk.lock();
  q.busy = false;
  if (failed != SZ_OK) for (OneCall *o : batch) o->call_res = failed;
  for (OneCall *o : batch) if (o != &me) { o->done = true; o->cv.notify_one(); }
  if (!q.waiting.empty()) q.waiting.front()->cv.notify_one();      /* the next batch gets its leader */
  return me.call_res;
}

}  // namespace

/* ------------------------------------------------------------------ one-call raw LZMA -------- */
extern "C" SRes LzmaDecode(Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, const Byte *propData,
                           unsigned propSize, ELzmaFinishMode finishMode, ELzmaStatus *status, ISzAlloc *alloc)
{
  /* reference LzmaDec.c:972-1002.  `alloc` is accepted for signature compatibility; the model lives in
   * device memory, so nothing is requested from it. */
  (void)alloc;
  const SizeT inSize = *srcLen, outSize = *destLen;
  *srcLen = *destLen = 0;
  if (inSize < 5) return SZ_ERROR_INPUT_EOF;
  CLzmaProps pr;
  SRes r = LzmaProps_Decode(&pr, propData, propSize);
  if (r != SZ_OK) return r;
  if (inSize > 0xFFFFFFFFull || outSize > 0xFFFFFFFFull) { lzb_set_error("single stream larger than 4 GiB"); return SZ_ERROR_UNSUPPORTED; }
  LzbContext *c = lzb_default_context();
  if (!c) return SZ_ERROR_FAIL;
  OneCall me;
  memset(&me.unit, 0, sizeof me.unit);
  LzbUnit &u = me.unit;
  u.src_len = (uint32_t)inSize; u.dst_cap = (uint32_t)outSize; u.dict_size = pr.dicSize;
  u.codec = LZB_CODEC_LZMA; u.lc = (uint8_t)pr.lc; u.lp = (uint8_t)pr.lp; u.pb = (uint8_t)pr.pb;
  u.finish = (uint8_t)finishMode;
  me.src = src; me.dest = dest;
  if (inSize + outSize <= kCombineMaxBytes)
    r = combined_decode(c, me);                /* small call: shares a launch with whoever else is calling right now */
  else
    r = Lzb_DecodeUnitsHost(c, src, inSize, dest, outSize, &u, &me.res, 1);
  if (r != SZ_OK) return r;
  *srcLen = me.res.src_used;
  *destLen = me.res.dst_done;
  *status = (ELzmaStatus)me.res.status;
  return me.res.res;
}

extern "C" int LzmaUncompress(unsigned char *dest, size_t *destLen, const unsigned char *src, SizeT *srcLen,
                              const unsigned char *props, size_t propsSize)
{
  /* reference LzmaLib.c:41-46 */
  ELzmaStatus status;
  return LzmaDecode(dest, destLen, src, srcLen, props, (unsigned)propsSize, LZMA_FINISH_ANY, &status, nullptr);
}

static SRes LzmaUncompressBatch_body(LzbContext *c, size_t n, unsigned char *const *dest, size_t *destLen,
                                    const unsigned char *const *src, size_t *srcLen, const unsigned char *props,
                                    size_t propsStride, SRes *res)
{
  if (!c) c = lzb_default_context();
  if (!c) return SZ_ERROR_FAIL;
  if (n == 0) return SZ_OK;
  /* gather the scattered streams into pinned arenas (callers that already hold arenas use
   * Lzb_DecodeUnitsHost directly and skip this copy) */
  std::vector<LzbUnit> units(n);
  std::vector<LzbResult> results(n);
  size_t pin = 0, pout = 0;
  for (size_t i = 0; i < n; i++)
  {
    res[i] = SZ_OK;
    CLzmaProps pr;
    SRes r = LzmaProps_Decode(&pr, props + i * propsStride, LZMA_PROPS_SIZE);
    LzbUnit &u = units[i];
    memset(&u, 0, sizeof u);
    if (srcLen[i] < 5) r = SZ_ERROR_INPUT_EOF;              /* LzmaDec.c:981-982 precedes the props check */
    else if (r == SZ_OK && (srcLen[i] > 0xFFFFFFFFull || destLen[i] > 0xFFFFFFFFull)) r = SZ_ERROR_UNSUPPORTED;
    if (r != SZ_OK) { res[i] = r; u.flags = 0x80; continue; }   /* 0x80: skipped on the host */
    u.src_off = pin; u.src_len = (uint32_t)srcLen[i];
    u.dst_off = pout; u.dst_cap = (uint32_t)destLen[i];
    u.dict_size = pr.dicSize; u.codec = LZB_CODEC_LZMA;
    u.lc = (uint8_t)pr.lc; u.lp = (uint8_t)pr.lp; u.pb = (uint8_t)pr.pb; u.finish = LZMA_FINISH_ANY;
    pin += (srcLen[i] + 15) & ~(size_t)15;
    pout += (destLen[i] + 15) & ~(size_t)15;
  }
  Byte *hin = (Byte *)Lzb_HostAlloc(pin + 16), *hout = (Byte *)Lzb_HostAlloc(pout + 16);
  if (!hin || !hout) { Lzb_HostFree(hin); Lzb_HostFree(hout); return SZ_ERROR_MEM; }
  std::vector<LzbUnit> run;
  std::vector<size_t> map;
  run.reserve(n); map.reserve(n);
  for (size_t i = 0; i < n; i++)
    if (!(units[i].flags & 0x80)) { memcpy(hin + units[i].src_off, src[i], srcLen[i]); run.push_back(units[i]); map.push_back(i); }
  SRes r = Lzb_DecodeUnitsHost(c, hin, pin + 16, hout, pout + 16, run.data(), results.data(), run.size());
  if (r == SZ_OK)
  {
    for (size_t k = 0; k < run.size(); k++)
    {
      size_t i = map[k];
      res[i] = results[k].res;
      srcLen[i] = results[k].src_used;
      destLen[i] = results[k].dst_done;
      memcpy(dest[i], hout + run[k].dst_off, results[k].dst_done);
    }
    for (size_t i = 0; i < n; i++)
      if (units[i].flags & 0x80) { srcLen[i] = 0; destLen[i] = 0; }
  }
  Lzb_HostFree(hin); Lzb_HostFree(hout);
  return r;
}

/* ------------------------------------------------------------------ checks (GPU) ------------- */
/* GF(2)[x]/P arithmetic on the host, only to stitch per-segment registers (reflected bit order) */
static uint32_t mul32(uint32_t a, uint32_t b)
{
  uint32_t p = 0;
  for (int k = 0; k < 32; k++) { if ((a >> (31 - k)) & 1) p ^= b; b = (b >> 1) ^ (0xEDB88320u & (0u - (b & 1))); }
  return p;
}
static uint64_t mul64(uint64_t a, uint64_t b)
{
  uint64_t p = 0;
  for (int k = 0; k < 64; k++) { if ((a >> (63 - k)) & 1) p ^= b; b = (b >> 1) ^ (0xC96C5795D7870F42ull & (0ull - (b & 1))); }
  return p;
}
uint32_t lzb_crc32_shift(uint32_t reg, uint64_t n)
{
  uint32_t r = 1u << 31, sq = 1u << 23;
  for (; n; n >>= 1) { if (n & 1) r = mul32(r, sq); sq = mul32(sq, sq); }
  return mul32(reg, r);
}
uint64_t lzb_crc64_shift(uint64_t reg, uint64_t n)
{
  uint64_t r = 1ull << 63, sq = 1ull << 55;
  for (; n; n >>= 1) { if (n & 1) r = mul64(r, sq); sq = mul64(sq, sq); }
  return mul64(reg, r);
}

UInt32 g_CrcTable[256 * 8];
UInt64 g_Crc64Table[256];

extern "C" SRes LzmaUncompressBatch(LzbContext *c, size_t n, unsigned char *const *dest, size_t *destLen,
                                    const unsigned char *const *src, size_t *srcLen, const unsigned char *props,
                                    size_t propsStride, SRes *res)
{
  return lzb_guarded([&]() { return LzmaUncompressBatch_body(c, n, dest, destLen, src, srcLen, props, propsStride, res); });
}

extern "C" void CrcGenerateTable(void)
{
  /* reference 7zCrc.c:49-74: the tables are part of the exported data, so they are still produced;
   * the GPU path does not read them (each warp builds its own copy in shared memory). */
  for (UInt32 i = 0; i < 256; i++)
  {
    UInt32 r = i;
    for (int j = 0; j < 8; j++) r = (r >> 1) ^ (0xEDB88320u & ~((r & 1) - 1));
    g_CrcTable[i] = r;
  }
  for (UInt32 i = 256; i < 256 * 8; i++)
  {
    UInt32 r = g_CrcTable[i - 256];
    g_CrcTable[i] = g_CrcTable[r & 0xFF] ^ (r >> 8);
  }
}

extern "C" void Crc64GenerateTable(void)
{
  for (UInt32 i = 0; i < 256; i++)
  {
    UInt64 r = i;
    for (int j = 0; j < 8; j++) r = (r >> 1) ^ (0xC96C5795D7870F42ull & ~((r & 1) - 1));
    g_Crc64Table[i] = r;
  }
}

/* finalized digest of a host buffer, computed on the GPU in 1 MiB segments (one warp each) */
static bool gpu_digest(const void *data, size_t size, int kind, uint64_t *digest)
{
  LzbContext *c = lzb_default_context();
  if (!c) return false;
  const size_t seg = 1u << 20;
  const size_t n = size ? (size + seg - 1) / seg : 1;
  std::vector<LzbUnit> units(n);
  std::vector<LzbResult> results(n);
  for (size_t i = 0; i < n; i++)
  {
    memset(&units[i], 0, sizeof(LzbUnit));
    units[i].dst_off = i * seg;
    units[i].dst_cap = (uint32_t)std::min(seg, size - i * seg);
    units[i].check = (uint8_t)kind;
  }
  std::lock_guard<std::mutex> g(c->mu);
  if (!lzb_cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return false;
  LzbSlot &s = c->slot[0];
  if (!lzb_slot_reserve(c, s, 0, size, n, 0)) return false;
  bool ok = lzb_cuda_ok(cudaMemcpyAsync(s.d_out, data, size, cudaMemcpyHostToDevice, s.stream), "H2D data") &&
            lzb_cuda_ok(cudaMemcpyAsync(s.d_units, units.data(), n * sizeof(LzbUnit), cudaMemcpyHostToDevice, s.stream), "H2D units");
  if (!ok) return false;
  int rc = lzb_launch_post_only(s.d_out, s.d_units, s.d_results, (uint32_t)n, s.stream);
  if (rc != 0) { lzb_set_error("post launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return false; }
  ok = lzb_cuda_ok(cudaMemcpyAsync(results.data(), s.d_results, n * sizeof(LzbResult), cudaMemcpyDeviceToHost, s.stream), "D2H results") &&
       lzb_cuda_ok(cudaStreamSynchronize(s.stream), "digest kernel");
  if (!ok) return false;
  /* crc(A||B) = shift(crc(A), |B|) ^ crc(B) for finalized values (preset/inversion cancel out) */
  uint64_t acc = 0;
  for (size_t i = 0; i < n; i++)
  {
    const uint64_t len = units[i].dst_cap;
    acc = (kind == LZB_CHECK_CRC32 ? (uint64_t)lzb_crc32_shift((uint32_t)acc, len) : lzb_crc64_shift(acc, len)) ^ results[i].check;
  }
  *digest = acc;
  return true;
}

extern "C" UInt32 CrcCalc(const void *data, size_t size)
{
  uint64_t d = 0;
  if (!gpu_digest(data, size, LZB_CHECK_CRC32, &d)) return 0;
  return (UInt32)d;
}
extern "C" UInt32 CrcUpdate(UInt32 v, const void *data, size_t size)
{
  /* register after `data` starting from register v:  shift(v ^ ~0, n) ^ crc(data) ^ ~0 */
  uint64_t d = 0;
  if (!gpu_digest(data, size, LZB_CHECK_CRC32, &d)) return v;
  return lzb_crc32_shift(v ^ 0xFFFFFFFFu, size) ^ (UInt32)d ^ 0xFFFFFFFFu;
}
extern "C" UInt64 Crc64Calc(const void *data, size_t size)
{
  uint64_t d = 0;
  if (!gpu_digest(data, size, LZB_CHECK_CRC64, &d)) return 0;
  return d;
}
extern "C" UInt64 Crc64Update(UInt64 v, const void *data, size_t size)
{
  uint64_t d = 0;
  if (!gpu_digest(data, size, LZB_CHECK_CRC64, &d)) return v;
  return lzb_crc64_shift(~v, size) ^ d ^ ~0ull;
}

/* ------------------------------------------------------------------ filters (GPU) ------------ */
static bool gpu_filter(Byte *data, size_t size, int filter, uint32_t param, uint8_t flags = 0, uint64_t aux = 0,
                       LzbResult *res_out = nullptr)
{
  LzbContext *c = lzb_default_context();
  if (!c) return false;
  if (size > 0xFFFFFFFFull) { lzb_set_error("filter buffer larger than 4 GiB"); return false; }
  LzbUnit u;
  LzbResult r;
  memset(&u, 0, sizeof u);
  u.dst_cap = (uint32_t)size; u.filter = (uint8_t)filter; u.filter_param = param;
  u.flags = flags; u.expect = aux;
  std::lock_guard<std::mutex> g(c->mu);
  if (!lzb_cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return false;
  LzbSlot &s = c->slot[0];
  if (!lzb_slot_reserve(c, s, 0, size, 1, 0)) return false;
  bool ok = lzb_cuda_ok(cudaMemcpyAsync(s.d_out, data, size, cudaMemcpyHostToDevice, s.stream), "H2D data") &&
            lzb_cuda_ok(cudaMemcpyAsync(s.d_units, &u, sizeof u, cudaMemcpyHostToDevice, s.stream), "H2D unit");
  if (!ok) return false;
  int rc = lzb_launch_post_only(s.d_out, s.d_units, s.d_results, 1, s.stream);
  if (rc != 0) { lzb_set_error("post launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return false; }
  ok = lzb_cuda_ok(cudaMemcpyAsync(data, s.d_out, size, cudaMemcpyDeviceToHost, s.stream), "D2H data") &&
       lzb_cuda_ok(cudaMemcpyAsync(&r, s.d_results, sizeof r, cudaMemcpyDeviceToHost, s.stream), "D2H result") &&
       lzb_cuda_ok(cudaStreamSynchronize(s.stream), "filter kernel");
  if (ok && res_out) *res_out = r;
  return ok;
}

/* A large buffer is cut into SEGMENTS that other warps can convert at the same time.  A segment may start at p when
 * none of the four bytes in front of p is an E8 / E9 opcode candidate: the converter's 3-bit history is provably zero
 * there and no conversion in front of p writes at or behind p (the same argument that lets the 32 lanes of a warp split
 * a buffer, lzb_post.cuh) -- so segment by segment gives byte for byte what one sequential pass gives.  The cut points
 * are found here on the host copy (a few bytes around each nominal boundary).  Every segment but the last is handed
 * over with four extra bytes so that its candidates run right up to the next segment's start. */
static const size_t kX86SegmentBytes = (size_t)1 << 16;

static bool x86_segments(const Byte *data, size_t size, std::vector<size_t> &starts)
{
  starts.assign(1, 0);
  if (size < 4 * kX86SegmentBytes) return false;
  for (size_t nominal = kX86SegmentBytes; nominal + kX86SegmentBytes <= size; nominal += kX86SegmentBytes)
  {
    size_t p = std::max(nominal, starts.back() + 8), clean = 0;
    for (size_t q = p - 4; q < p; q++) clean = (data[q] & 0xFE) == 0xE8 ? 0 : clean + 1;
    while (clean < 4 && p + 8 < size) { clean = (data[p] & 0xFE) == 0xE8 ? 0 : clean + 1; p++; }
    if (clean < 4 || p + 8 >= size) break;                 /* (a buffer that is nothing but candidates: fewer segments) */
    starts.push_back(p);
  }
  return starts.size() > 1;
}

extern "C" SizeT x86_Convert(Byte *data, SizeT size, UInt32 ip, UInt32 *state, int encoding)
{
  /* reference Bra86.c:11-85: both directions, history carried in *state, returns the bytes processed.
   * One warp scans a segment in 32 pieces that start at history-free positions (lzb_post.cuh). */
  if (size < 5) return 0;
  const uint8_t flags = (uint8_t)(4 | (encoding ? 2 : 0));
  std::vector<size_t> starts;
  if (size > 0xFFFFFFFFull || !x86_segments(data, size, starts))
  {
    LzbResult r;
    if (!gpu_filter(data, size, LZB_FILTER_X86, ip, flags, (uint64_t)(*state & 7) << 32, &r)) return 0;
    *state = (UInt32)r.status;
    return r.src_used;
  }
  LzbContext *c = lzb_default_context();
  if (!c) return 0;
  const size_t n = starts.size();
  std::vector<LzbUnit> units(n);
  std::vector<LzbResult> results(n);
  for (size_t i = 0; i < n; i++)
  {
    LzbUnit &u = units[i];
    memset(&u, 0, sizeof u);
    const size_t end = i + 1 < n ? starts[i + 1] + 4 : size;
    u.dst_off = starts[i]; u.dst_cap = (uint32_t)(end - starts[i]);
    u.filter = LZB_FILTER_X86; u.filter_param = ip + (UInt32)starts[i];
    u.flags = flags;
    u.expect = i == 0 ? (uint64_t)(*state & 7) << 32 : 0;
  }
  std::lock_guard<std::mutex> g(c->mu);
  if (!lzb_cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return 0;
  LzbSlot &s = c->slot[0];
  if (!lzb_slot_reserve(c, s, 0, size, n, 0)) return 0;
  bool ok = lzb_cuda_ok(cudaMemcpyAsync(s.d_out, data, size, cudaMemcpyHostToDevice, s.stream), "H2D data") &&
            lzb_cuda_ok(cudaMemcpyAsync(s.d_units, units.data(), n * sizeof(LzbUnit), cudaMemcpyHostToDevice, s.stream), "H2D units");
  if (!ok) return 0;
  const int rc = lzb_launch_post_only(s.d_out, s.d_units, s.d_results, (uint32_t)n, s.stream);
  if (rc != 0) { lzb_set_error("post launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return 0; }
  ok = lzb_cuda_ok(cudaMemcpyAsync(data, s.d_out, size, cudaMemcpyDeviceToHost, s.stream), "D2H data") &&
       lzb_cuda_ok(cudaMemcpyAsync(results.data(), s.d_results, n * sizeof(LzbResult), cudaMemcpyDeviceToHost, s.stream), "D2H results") &&
       lzb_cuda_ok(cudaStreamSynchronize(s.stream), "filter kernel");
  if (!ok) return 0;
  *state = (UInt32)results[n - 1].status;
  return starts[n - 1] + results[n - 1].src_used;
}

/* ------------------------------------------------------------------ LZMA + x86 one-call form -------- */
/* reference Lzma86Dec.c:13-56: 1 filter byte, 5 props bytes, 8 size bytes, then a raw LZMA stream */
extern "C" SRes Lzma86_GetUnpackSize(const Byte *src, SizeT srcLen, UInt64 *unpackSize)
{
  if (srcLen < LZMA86_HEADER_SIZE) return SZ_ERROR_INPUT_EOF;
  *unpackSize = lzb_le64(src + LZMA86_SIZE_OFFSET);
  return SZ_OK;
}

extern "C" SRes Lzma86_Decode(Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen)
{
  if (*srcLen < LZMA86_HEADER_SIZE) return SZ_ERROR_INPUT_EOF;
  const int use_filter = src[0];
  if (use_filter > 1) { *destLen = 0; return SZ_ERROR_UNSUPPORTED; }
  SizeT pure = *srcLen - LZMA86_HEADER_SIZE;
  ELzmaStatus status;
  const SRes res = LzmaDecode(dest, destLen, src + LZMA86_HEADER_SIZE, &pure, src + 1, LZMA_PROPS_SIZE, LZMA_FINISH_ANY, &status, nullptr);
  *srcLen = pure + LZMA86_HEADER_SIZE;
  if (res != SZ_OK) return res;
  if (use_filter == 1)
  {
    UInt32 st = 0;
    x86_Convert(dest, *destLen, 0, &st, 0);
  }
  return SZ_OK;
}

/* ------------------------------------------------------------------ BCJ2 merge, stand-alone ---------- */
/* reference Bcj2.c:28-127 (Bcj2.h:24-30): the four streams go to the device, one warp merges them */
extern "C" int Bcj2_Decode(const Byte *buf0, SizeT size0, const Byte *buf1, SizeT size1, const Byte *buf2, SizeT size2,
                           const Byte *buf3, SizeT size3, Byte *outBuf, SizeT outSize)
{
  if ((size0 | size1 | size2 | size3 | outSize) > 0xFFFFFFFFull) { lzb_set_error("Bcj2_Decode: stream larger than 4 GiB"); return SZ_ERROR_UNSUPPORTED; }
  LzbContext *c = lzb_default_context();
  if (!c) return SZ_ERROR_FAIL;
  std::lock_guard<std::mutex> g(c->mu);
  if (!lzb_cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return SZ_ERROR_FAIL;
  LzbSlot &s = c->slot[0];
  auto up16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
  /* output arena: [out | main | call | jump]; packed arena: [range-coder stream | job] */
  LzbBcj2Job j;
  memset(&j, 0, sizeof j);
  j.out_off = 0; j.out_size = (uint32_t)outSize;
  j.main_off = up16(outSize); j.main_size = (uint32_t)size0;
  j.call_off = j.main_off + up16(size0); j.call_size = (uint32_t)size1;
  j.jump_off = j.call_off + up16(size1); j.jump_size = (uint32_t)size2;
  j.rc_off = 0; j.rc_size = (uint32_t)size3;
  const size_t job_at = up16(size3 + 16);
  if (!lzb_slot_reserve(c, s, job_at + sizeof j + 16, (size_t)(j.jump_off + up16(size2)) + 16, 1, 0)) return SZ_ERROR_MEM;
  LzbBcj2Job *d_job = reinterpret_cast<LzbBcj2Job *>(s.d_packed + job_at);
  bool ok = true;
  if (size0) ok = ok && lzb_cuda_ok(cudaMemcpyAsync(s.d_out + j.main_off, buf0, size0, cudaMemcpyHostToDevice, s.stream), "H2D bcj2 main");
  if (size1) ok = ok && lzb_cuda_ok(cudaMemcpyAsync(s.d_out + j.call_off, buf1, size1, cudaMemcpyHostToDevice, s.stream), "H2D bcj2 call");
  if (size2) ok = ok && lzb_cuda_ok(cudaMemcpyAsync(s.d_out + j.jump_off, buf2, size2, cudaMemcpyHostToDevice, s.stream), "H2D bcj2 jump");
  if (size3) ok = ok && lzb_cuda_ok(cudaMemcpyAsync(s.d_packed, buf3, size3, cudaMemcpyHostToDevice, s.stream), "H2D bcj2 rc");
  ok = ok && lzb_cuda_ok(cudaMemcpyAsync(d_job, &j, sizeof j, cudaMemcpyHostToDevice, s.stream), "H2D bcj2 job");
  if (!ok) return SZ_ERROR_FAIL;
  const int rc = lzb_launch_bcj2(s.d_packed, s.d_out, d_job, 1, s.stream);
  if (rc != 0) { lzb_set_error("bcj2 launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return SZ_ERROR_FAIL; }
  if (outSize) ok = ok && lzb_cuda_ok(cudaMemcpyAsync(outBuf, s.d_out, outSize, cudaMemcpyDeviceToHost, s.stream), "D2H bcj2 out");
  ok = ok && lzb_cuda_ok(cudaMemcpyAsync(&j, d_job, sizeof j, cudaMemcpyDeviceToHost, s.stream), "D2H bcj2 job") &&
       lzb_cuda_ok(cudaStreamSynchronize(s.stream), "bcj2 kernel");
  return ok ? j.res : SZ_ERROR_FAIL;
}

/* reference Bra.c:6-133, BraIA64.c:16-69: stateless, one aligned instruction slot per GPU thread */
static SizeT risc_convert(int filter, Byte *data, SizeT size, UInt32 ip, int encoding, SizeT min_size)
{
  if (size < min_size) return 0;
  LzbResult r;
  if (!gpu_filter(data, size, filter, ip, (uint8_t)(4 | (encoding ? 2 : 0)), 0, &r)) return 0;
  return r.src_used;
}
extern "C" SizeT ARM_Convert(Byte *data, SizeT size, UInt32 ip, int encoding) { return risc_convert(LZB_FILTER_ARM, data, size, ip, encoding, 4); }
extern "C" SizeT ARMT_Convert(Byte *data, SizeT size, UInt32 ip, int encoding) { return risc_convert(LZB_FILTER_ARMT, data, size, ip, encoding, 4); }
extern "C" SizeT PPC_Convert(Byte *data, SizeT size, UInt32 ip, int encoding) { return risc_convert(LZB_FILTER_PPC, data, size, ip, encoding, 4); }
extern "C" SizeT SPARC_Convert(Byte *data, SizeT size, UInt32 ip, int encoding) { return risc_convert(LZB_FILTER_SPARC, data, size, ip, encoding, 4); }
extern "C" SizeT IA64_Convert(Byte *data, SizeT size, UInt32 ip, int encoding) { return risc_convert(LZB_FILTER_IA64, data, size, ip, encoding, 16); }

extern "C" void Delta_Init(Byte *state) { memset(state, 0, DELTA_STATE_SIZE); }

extern "C" void Delta_Encode(Byte *state, unsigned delta, Byte *data, SizeT size)
{
  /* reference Delta.c:20-41.  The GPU pass takes a zero history; the incoming one only touches the first
   * `delta` bytes (out[i] = in[i] - hist[i]); the outgoing history is the last `delta` INPUT bytes. */
  if (size == 0) return;
  Byte next[DELTA_STATE_SIZE];
  if (size >= delta) memcpy(next, data + size - delta, delta);
  else { memcpy(next, state + size, delta - size); memcpy(next + delta - size, data, size); }
  if (!gpu_filter(data, size, LZB_FILTER_DELTA, delta, 2)) return;
  for (unsigned i = 0; i < delta && i < size; i++) data[i] = (Byte)(data[i] - state[i]);
  memcpy(state, next, delta);
}

extern "C" void Delta_Decode(Byte *state, unsigned delta, Byte *data, SizeT size)
{
  /* reference Delta.c:43-62.  A non-zero incoming history is folded in on the host (it only affects
   * the first `delta` bytes: out[i] = in[i] + hist[i]) before the GPU prefix-sum, and the outgoing
   * history is the last `delta` outputs. */
  if (size == 0) return;
  for (unsigned i = 0; i < delta && i < size; i++) data[i] = (Byte)(data[i] + state[i]);
  if (!gpu_filter(data, size, LZB_FILTER_DELTA, delta)) return;
  Byte next[DELTA_STATE_SIZE];
  if (size >= delta) memcpy(next, data + size - delta, delta);
  else { memcpy(next, state + size, delta - size); memcpy(next + delta - size, data, size); }
  memcpy(state, next, delta);
}
