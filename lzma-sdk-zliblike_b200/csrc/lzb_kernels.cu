/* lzb_kernels.cu -- sm_100a kernels and their launchers.
 *
 *   lzb_decode_units_kernel<shared model, G> : persistent grid (one CTA per SM).  Every warp carries
 *       K = 32 / G units (G lanes each) and runs ONE loop in which each unit decodes one symbol per trip
 *       (lzb_core.cuh); a group that finishes its unit pulls the next one from a global ticket, applies the
 *       unit's post-filter and check and writes one LzbResult.
 *
 * Residency.  The bit loop is a serial dependency chain: throughput comes from resident units, and a warp
 * instruction that serves K units at once costs one issue slot (profiles/r01_c_residency.md).  Two things
 * cap the units per SM: shared memory for the models and the register file.  Only the part of the model a
 * plain symbol touches lives in shared memory (6,448 B at lc3/lp0/pb2 -> 36 fit); the matched-literal trees
 * (8 KiB, touched once per match) live in an L2-resident HBM slab.  With K units
 * per warp only units/K warps need registers: the default is G = 16, 18 warps, 36 units per SM.
 * Launches with at most 29 units per SM run one unit per warp (G = 32) in the look-ahead form of the decoder
 * (lzb_plan_few_units; up to 16 per SM one wave spread over all SMs, beyond that 16 persistent warps per SM).
 * Models whose shared part exceeds 56 KiB (raw LZMA with lc+lp >= 6) live entirely in the slab.
 *
 * Streaming.  The <..., kFlow = true> instantiation is the same loop for a launch that is fed and drained while
 * it runs (LzbFlow, lzb_kernels.h; host side: decode_units_flow in lzb_host.cu): a group waits for the arrival
 * flag of its unit's chunk before it starts the unit, and the group that completes a chunk raises a flag in
 * mapped host memory.
 *
 *   lzb_post_only_kernel : filter / check over bytes already in device memory (stand-alone x86_Convert,
 *       ARM/ARMT/PPC/SPARC/IA64_Convert, Delta_*, Crc* calls; block- and file-level passes of xz and 7z).
 *   lzb_sha256_kernel    : xz block check 10, one thread per digest.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "lzb_core.cuh"
#include "lzb_post.cuh"
#include "lzb_sha256.cuh"
#include "lzb_kernels.h"

namespace lzb {

/* warps per CTA a variant is compiled for (register budget = 65536 / (32 * warps), warps rounded up to a
 * multiple of four): G = 32: 16 warps (128 registers), G = 16: 18 warps (96; a bound of 20 gives the same
 * budget but a slower schedule: 7.17 vs 7.68 GB/s), G <= 8: 8 warps */
constexpr uint32_t max_warps_for(uint32_t G) { return G == 32 ? 16u : (G == 16 ? 18u : 8u); }

/* ---- a launch that is fed and drained while it runs (LzbFlow, lzb_kernels.h) ---- */
constexpr long long kFlowTimeoutCycles = 8000000000ll;         /* ~4 s at the B200's 1.97 GHz */

/* wait until unit i's packed bytes (and the chunk after them: a 32-byte sector may straddle the boundary, and
 * L1 must never see a sector before all of it has landed) are in device memory; false = give up */
template <uint32_t G> __device__ __forceinline__ bool flow_wait(const LzbFlow &f, uint32_t i)
{
  typedef Group<G> Grp;
  const uint32_t c = i / f.chunk_units;
  const uint32_t need = c + 1 < f.n_chunks ? c + 1 : f.n_chunks - 1;
  uint32_t ok = 1;
  if (Grp::lane() == 0)
  {
    const long long t0 = clock64();
    uint32_t v, spins = 0;
    for (;;)
    {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f.arrived + need) : "memory");
      if (v != 0) break;
      __nanosleep(1000);
      /* the abort word lives in host memory: every group of the GPU polling it over PCIe would starve the
       * very upload it is waiting for (measured: 0.8 GB/s), so it is looked at every ~16 ms only */
      if ((++spins & 16383u) == 0)
      {
        uint32_t a;
        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(a) : "l"(f.abort) : "memory");
        if (a != 0 || clock64() - t0 > kFlowTimeoutCycles) { ok = 0; break; }
      }
    }
  }
  ok = Grp::bcast(ok, 0);
  return ok != 0;
}

/* unit i is finished and its bytes are visible system-wide; whoever completes a chunk tells the host */
template <uint32_t G> __device__ __forceinline__ void flow_done(const LzbFlow &f, uint32_t i, uint32_t n_units)
{
  typedef Group<G> Grp;
  __threadfence_system();
  Grp::sync();
  if (Grp::lane() == 0)
  {
    const uint32_t c = i / f.chunk_units;
    const uint32_t left = n_units - c * f.chunk_units;
    const uint32_t count = left < f.chunk_units ? left : f.chunk_units;
    if (atomicAdd(f.done + c, 1u) + 1u == count)
    {
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" : : "l"(f.ready + c), "r"(1u) : "memory");
    }
  }
  Grp::sync();
}

template <bool kSharedModel, uint32_t G, bool kFlow, bool kAhead>
__global__ void __launch_bounds__(max_warps_for(G) * 32, 1)
lzb_decode_units_kernel(const uint8_t *__restrict__ packed, uint8_t *__restrict__ out,
                        const LzbUnit *__restrict__ units, LzbResult *__restrict__ results, uint32_t n_units,
                        uint32_t *ticket, uint32_t region_bytes, uint32_t smem_stride, uint32_t slab_stride, uint8_t *model_slab,
                        const LzbFlow flow)
{
  extern __shared__ __align__(16) uint8_t lzb_smem[];
  typedef Group<G> Grp;
  constexpr uint32_t K = 32 / G;
  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t grp = warp * K + ((threadIdx.x & 31u) / G);               /* group index inside the CTA */
  const uint32_t groups_per_cta = (blockDim.x >> 5) * K;
  uint8_t *slab = model_slab + ((size_t)blockIdx.x * groups_per_cta + grp) * slab_stride;
  /* shared model: region in shared memory, matched-literal trees at the start of the group's slab slice --
   * or, in a one-unit-per-warp launch whose units leave room for it (smem_stride > region_bytes), behind the
   * region in shared memory as well: there the unit is bound by its own dependency chain and shared memory is
   * mostly empty, so the eight dependent L2 round trips of a matched literal are saved for free (matched
   * literals are 2-4 % of the symbols of the bench corpora, profiles/r01_l_symbol_stats.md);
   * global model: both in the slab slice (region first) */
  uint8_t *region = kSharedModel ? lzb_smem + (size_t)grp * smem_stride : slab;
  uint8_t *matched_at = kSharedModel ? slab : slab + region_bytes;
  if (kSharedModel && G == 32 && smem_stride > region_bytes) matched_at = region + region_bytes;
  uint16_t *matched = reinterpret_cast<uint16_t *>(matched_at);

  UnitDecoderT<G, kSharedModel, kAhead> d;
  LzbUnit u;
  uint32_t ui = 0;
  uint8_t *dst = out;
  d.st = S_FETCH;
  for (;;)
  {
    if (d.st == S_FETCH)
    {
      uint32_t i = 0;
      /* (tried: once fewer units than warps are left, only the first group of a warp takes units -- at most one unit per
       * warp in the tail.  Slower: 7.46 vs 7.52 GB/s at 8,192 units, 7.85 vs 8.25 at two full waves: an empty slot costs
       * more than two units sharing a warp; profiles/r02_kernel_log.md K) */
      if (Grp::lane() == 0) i = atomicAdd(ticket, 1u);
      i = Grp::bcast(i, 0);
      if (i >= n_units) d.st = S_EXIT;
      else if (kFlow && flow.arrived && !flow_wait<G>(flow, i))
      {
        /* the input never came (host gave up or timed out): answer and move on, never hang */
        if (Grp::lane() == 0)
        {
          LzbResult r;
          r.res = LZB_SZ_ERROR_FAIL; r.status = LZB_ST_NOT_SPECIFIED; r.src_used = 0; r.dst_done = 0; r.check = 0;
          results[i] = r;
        }
        flow_done<G>(flow, i, n_units);
      }
      else
      {
        ui = i;
        u = units[i];
        dst = out + u.dst_off;
        const bool stateful = (u.flags & LZB_UNIT_STATEFUL) != 0;
        d.begin(packed + u.src_off, u.src_len, dst, stateful ? u.filter_param : u.dst_cap, u.dict_size,
                reinterpret_cast<uint16_t *>(region), matched);
        d.start_unit(u, stateful ? reinterpret_cast<LzbStreamState *>(u.expect) : nullptr);
      }
    }
    if (d.st != S_RUN && d.st != S_EXIT)
    {
      d.advance();
      if (d.st == S_UNIT_END)
      {
        LzbResult r;
        d.finish_unit(r);
        /* (inlined on purpose: out of line, with the unit and result records forced into local memory, the dense
         * instantiation ran 2 % slower -- profiles/r02_kernel_log.md) */
        if (!(u.flags & LZB_UNIT_STATEFUL)) post_process<G, false, kAhead>(u, r, dst, region, reinterpret_cast<uint8_t *>(matched), kSharedModel ? region_bytes : kPostSmemBytes);
        if (Grp::lane() == 0) results[ui] = r;
        Grp::sync();
        if (kFlow && flow.done) flow_done<G>(flow, ui, n_units);
        d.st = S_FETCH;
      }
    }
    if (__all_sync(0xffffffffu, d.st == S_EXIT)) break;
    /* hot loop: one symbol per unit and trip, until some unit of the warp needs the driver again */
    if (kAhead)
    {
      /* one unit per warp: every lane holds the same state, the loop test is uniform without asking the warp */
      while (d.st == S_RUN) d.step();
    }
    else
    {
      do
      {
        if (d.st == S_RUN) d.step();
      } while (!__any_sync(0xffffffffu, d.st != S_RUN && d.st != S_EXIT));
    }
  }
}

/* Stand-alone post-filter / check pass over already decoded bytes (x86_Convert, Delta_Decode, CrcCalc,
 * Crc64Calc entry points of the C-ABI): one warp per unit, same device routines as the fused path. */
__global__ void __launch_bounds__(128)
lzb_post_only_kernel(uint8_t *__restrict__ data, const LzbUnit *__restrict__ units, LzbResult *__restrict__ results,
                     uint32_t n_units)
{
  __shared__ __align__(16) uint8_t scratch[4][kPostScratchBytes];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t i = blockIdx.x * (blockDim.x >> 5) + warp;
  if (i >= n_units) return;
  const LzbUnit u = units[i];
  LzbResult r;
  r.res = LZB_SZ_OK; r.status = LZB_ST_NOT_SPECIFIED; r.src_used = 0; r.dst_done = u.dst_cap; r.check = 0;
  post_process<32, true>(u, r, data + u.dst_off, scratch[warp], scratch[warp], kPostScratchBytes);
  if (lane == 0) results[i] = r;
}

/* BCJ2 merge (what the reference's Bcj2_Decode computes, Bcj2.c:28-127).  The encoder moved the absolute targets of
 * x86 CALLs (E8) and of JMP / Jcc (E9, 0F 8x) out of the main stream into two side streams, big-endian, and coded one
 * adaptive bit per opcode candidate -- 256 models for E8 (chosen by the byte before it), one for E9, one for Jcc --
 * with a range coder into a fourth stream: bit 1 = "the four bytes after this opcode come from the side stream".
 *
 * One WARP per folder.  The merge is a serial chain only from candidate to candidate; between two candidates the main
 * stream is plain bytes, and that is where the time goes, so the warp looks at 32 bytes at once: every lane tests
 * its byte (with its left neighbour's byte for the 0F 8x form), a ballot finds the first candidate, and the bytes up
 * to and including it are copied by the lanes in one coalesced step.  The bit models live in shared memory; every
 * lane carries the coder's scalars redundantly.  Main stream and output share memory (the main stream sits at the
 * end of the output range, 7zDec.c:408-430): a batch is loaded completely before it is stored and the output never
 * overtakes the input, so the in-place merge is safe front to back. */
namespace {
constexpr uint32_t kBcj2Warps = 4;
constexpr uint32_t kBcj2Models = 2 + 256;
}

__global__ void __launch_bounds__(kBcj2Warps * 32)
lzb_bcj2_kernel(const uint8_t *__restrict__ packed, uint8_t *out_arena, LzbBcj2Job *jobs, uint32_t n_jobs)
{
  __shared__ uint16_t models[kBcj2Warps][kBcj2Models + 2];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t j = blockIdx.x * kBcj2Warps + warp;
  if (j >= n_jobs) return;
  const LzbBcj2Job job = jobs[j];
  const uint8_t *main_s = out_arena + job.main_off;
  const uint8_t *side[2] = { out_arena + job.call_off, out_arena + job.jump_off };
  uint32_t side_left[2] = { job.call_size, job.jump_size };
  const uint8_t *rc = packed + job.rc_off;
  uint8_t *out = out_arena + job.out_off;
  uint16_t *model = models[warp];
  for (uint32_t i = lane; i < kBcj2Models; i += 32) model[i] = 1024;
  __syncwarp();

  /* range coder start: five bytes, the first ends up shifted out */
  uint32_t range = 0xFFFFFFFFu, code = 0, rc_at = 0;
  bool broken = job.rc_size < 5;
  if (!broken) for (; rc_at < 5; rc_at++) code = (code << 8) | rc[rc_at];

  uint32_t in_at = 0, out_at = 0, prev = 0;
  while (!broken && out_at < job.out_size)
  {
    /* plain run: copy up to and including the next opcode candidate */
    uint32_t budget = min(job.main_size - in_at, job.out_size - out_at);
    bool at_candidate = false;
    while (budget != 0 && !at_candidate)
    {
      /* (a damaged folder can make the output overtake the main stream it shares memory with: then the reference's
       * byte-by-byte order is what defines the bytes, one at a time) */
      const uint32_t n = (out + out_at > main_s + in_at) ? 1u : min(budget, 32u);
      const uint32_t b = lane < n ? main_s[in_at + lane] : 0u;
      uint32_t left = __shfl_up_sync(0xffffffffu, b, 1);
      if (lane == 0) left = prev;
      const bool cand = lane < n && ((b & 0xFEu) == 0xE8u || (left == 0x0Fu && (b & 0xF0u) == 0x80u));
      const uint32_t hits = __ballot_sync(0xffffffffu, cand);
      const uint32_t plain = hits ? (uint32_t)__ffs((int)hits) - 1u : n;      /* bytes before the candidate */
      const uint32_t copy = hits ? plain + 1u : n;
      __syncwarp();                                       /* the whole batch is in registers before any byte of it is stored */
      if (lane < copy) out[out_at + lane] = (uint8_t)b;
      if (plain != 0) prev = __shfl_sync(0xffffffffu, b, (int)plain - 1);
      in_at += plain; out_at += copy; budget -= plain;
      at_candidate = hits != 0;
      __syncwarp();
    }
    if (!at_candidate || out_at == job.out_size) break;

    /* the candidate's bit */
    const uint32_t op = main_s[in_at++];
    uint16_t *m = model + (op == 0xE8u ? prev : (op == 0xE9u ? 256u : 257u));
    const uint32_t pr = *m;
    const uint32_t bound = (range >> 11) * pr;
    const bool moved = code >= bound;
    __syncwarp();
    if (moved) { range -= bound; code -= bound; *m = (uint16_t)(pr - (pr >> 5)); }
    else { range = bound; *m = (uint16_t)(pr + ((2048u - pr) >> 5)); }
    __syncwarp();
    if (range < (1u << 24))
    {
      if (rc_at == job.rc_size) { broken = true; break; }
      range <<= 8;
      code = (code << 8) | rc[rc_at++];
    }
    if (!moved) { prev = op; continue; }

    /* four bytes of absolute target from the side stream -> relative displacement */
    const uint32_t which = op == 0xE8u ? 0u : 1u;
    if (side_left[which] < 4) { broken = true; break; }
    const uint8_t *v = side[which];
    side[which] += 4; side_left[which] -= 4;
    const uint32_t target = ((uint32_t)v[0] << 24) | ((uint32_t)v[1] << 16) | ((uint32_t)v[2] << 8) | v[3];
    const uint32_t rel = target - (out_at + 4);
    const uint32_t room = min(4u, job.out_size - out_at);
    if (lane < room) out[out_at + lane] = (uint8_t)(rel >> (8 * lane));
    out_at += room;
    prev = rel >> 24;
    __syncwarp();
  }
  if (lane == 0) jobs[j].res = (!broken && out_at == job.out_size) ? LZB_SZ_OK : LZB_SZ_ERROR_DATA;
}

/* xz block check 10 (SHA-256, Xz.c:45-88): a digest is one serial chain, so one thread per job; blocks of
 * an xz file are independent jobs.  Not fused into the decode kernel: it would grow the hot kernel's code
 * and registers for a check the reference's own encoder never selects by default. */
__global__ void __launch_bounds__(32)
lzb_sha256_kernel(const uint8_t *__restrict__ data, LzbShaJob *jobs, uint32_t n_jobs)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_jobs) return;
  LzbShaJob j = jobs[i];
  const uint8_t *p = data + j.off;
  if (j.finalize) sha256_finish(j.state, p, j.len, j.before, j.digest);
  else for (uint64_t k = 0; k + 64 <= j.len; k += 64) sha256_block(j.state, p + k);
  jobs[i] = j;
}

}  // namespace lzb

/* ------------------------------------------------------------------ launchers (C linkage) ------ */
static uint32_t env_u32(const char *name, uint32_t dflt)
{
  const char *v = getenv(name);
  return (v && atoi(v) > 0) ? (uint32_t)atoi(v) : dflt;
}

static int g_dev_sms = 0;
static uint32_t g_ahead_max_warps = 16;     /* LZB_AHEAD_WARPS overrides (0 = never), read once */
static size_t g_dev_smem_optin = 0;

static cudaError_t query_device()
{
  if (g_dev_sms) return cudaSuccess;
  int dev = 0, sms = 0, smem = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if (e != cudaSuccess) return e;
  g_dev_sms = sms;
  g_dev_smem_optin = (size_t)smem;
  if (const char *v = getenv("LZB_AHEAD_WARPS")) g_ahead_max_warps = (uint32_t)atoi(v);
  return cudaSuccess;
}


extern "C" int lzb_launch_geometry(uint32_t smem_model_bytes, uint32_t gmem_model_bytes, LzbGeometry *g)
{
  cudaError_t e = query_device();
  if (e != cudaSuccess) return (int)e;
  /* the dead model doubles as post-pass scratch: keep the minima lzb_post.cuh asks for */
  uint32_t region = smem_model_bytes < lzb::kPostSmemBytes ? lzb::kPostSmemBytes : smem_model_bytes;
  region = (region + 15u) & ~15u;
  uint32_t gpart = gmem_model_bytes < lzb::kPostScratchBytes ? lzb::kPostScratchBytes : gmem_model_bytes;
  gpart = (gpart + 15u) & ~15u;
  g->region_bytes = region;
  g->smem_stride = region;
  g->ctas = (uint32_t)g_dev_sms;
  /* lanes per unit (LZB_GROUP = 32, 16, 8 or 4) and units per SM (LZB_UNITS), both overridable for
   * measurements; defaults from profiles/r01_c_residency.md */
  static const uint32_t env_group = env_u32("LZB_GROUP", LZB_DEFAULT_GROUP), env_units = env_u32("LZB_UNITS", LZB_DEFAULT_UNITS_PER_SM);
  uint32_t G = env_group;
  if (G != 32 && G != 16 && G != 8 && G != 4) G = LZB_DEFAULT_GROUP;
  const uint32_t K = 32 / G;
  uint32_t want_units = env_units;
  if (region <= LZB_MAX_SHARED_MODEL_BYTES)
  {
    uint32_t units = (uint32_t)(g_dev_smem_optin / region);
    if (units > want_units) units = want_units;
    uint32_t warps = units / K;
    if (warps > lzb::max_warps_for(G)) warps = lzb::max_warps_for(G);
    /* a model so large that fewer than K fit: fall back to one unit per warp */
    if (warps < 1)
    {
      G = 32;
      warps = units < 1 ? 1 : (units > lzb::max_warps_for(32) ? lzb::max_warps_for(32) : units);
    }
    g->warps_per_cta = warps;
    g->shared_model = 1;
    g->slab_stride = gpart;
  }
  else
  {
    G = 32;
    g->warps_per_cta = 8;
    g->shared_model = 0;
    g->slab_stride = region + gpart;
  }
  g->group = G;
  g->ahead = 0;
  g->units_per_cta = g->warps_per_cta * (32 / G);
  g->slab_bytes = (size_t)g->slab_stride * g->units_per_cta * g->ctas;
  return 0;
}

/* Launches with fewer units than warps: pure arithmetic on the default geometry (no device access, tested on the
 * CPU in tests/test_abi_and_host.py).  regroup_allowed = LZB_GROUP not forced; dense = LZB_SMEM_MATCHED=0. */
extern "C" void lzb_plan_few_units(LzbGeometry *gp, uint32_t n_units, size_t smem_optin, int regroup_allowed, int dense)
{
  LzbGeometry &g = *gp;
  /* Few units (large LZMA2 groups, xz blocks): when every unit can have a warp of its own in one wave,
   * sharing warps only serialises them.  The slab is sized for the default geometry, which is the larger. */
  if (g.shared_model && g.group < 32 && regroup_allowed)
  {
    const uint32_t w32 = g.units_per_cta < lzb::max_warps_for(32) ? g.units_per_cta : lzb::max_warps_for(32);
    /* ... and a little beyond one wave: up to ~29 units per SM the look-ahead form with 16 persistent warps (units
     * taken from the ticket as warps come free) still beats two units per warp at partial residency -- 5.15 vs 4.74
     * GB/s at 3,072 units of 1 MiB, 6.41 vs 6.19 at 4,096; at 5,328 (36 per SM) the dense form wins 8.08 to 6.19
     * (profiles/r02_kernel_log.md J) */
    const uint32_t dense_from = (g.units_per_cta * 29u) / 36u;          /* units per SM where the dense form takes over */
    if (n_units <= g.ctas * w32 || (w32 == lzb::max_warps_for(32) && n_units <= g.ctas * dense_from))
    {
      g.group = 32;
      g.warps_per_cta = w32;
      g.units_per_cta = w32;
    }
  }
  /* One unit per warp: spread the units over all SMs (a unit alone on an SM decodes 3.3 MB/s, one of sixteen
   * 2.5 MB/s), and when the units of a CTA leave room for it in shared memory the matched-literal trees move in
   * as well (the slab slice of the group stays unused): measured 2.62 -> 3.04 GB/s at 1,024 units of 1 MiB,
   * 0.39 -> 0.48 GB/s at 148 (profiles/r01_j_few_units.md).  dense keeps the trees in the slab and the old
   * packing (for measurements). */
  if (g.shared_model && g.group == 32 && !dense && n_units != 0)
  {
    const uint32_t need = (n_units + g.ctas - 1) / g.ctas;
    if (need < g.warps_per_cta)
    {
      g.warps_per_cta = need;
      g.units_per_cta = need;
    }
    const uint32_t full = g.region_bytes + g.slab_stride;
    if (g.warps_per_cta <= (uint32_t)(smem_optin / full)) g.smem_stride = full;
  }
  /* One unit per warp (at most 16 warps on an SM): a unit is bound by its own dependency chain, so the look-ahead form
   * of the decoder (predicated normalisation, both children of a tree node loaded while the bit is decided;
   * lzb_core.cuh) trades ~15 % more issued instructions for a shorter chain: +18 % at 7 warps per SM, +21 % at 1,
   * still +6..8 % at 15-16 (profiles/r02_d_*, r02_e_*).  With two units per warp the issue slots are the bound and the
   * lean form wins by 18 % (profiles/r02_b_*). */
  g.ahead = (g.shared_model && g.group == 32 && g.warps_per_cta <= g_ahead_max_warps) ? 1 : 0;
}

template <bool kShared, uint32_t G, bool kFlow, bool kAhead>
static cudaError_t launch_variant3(const LzbGeometry &g, uint32_t ctas, const uint8_t *d_packed, uint8_t *d_out,
                                  const LzbUnit *d_units, LzbResult *d_results, uint32_t n_units, uint32_t *d_ticket,
                                  uint8_t *d_slab, const LzbFlow &flow, cudaStream_t stream)
{
  const size_t smem = kShared ? (size_t)g.units_per_cta * g.smem_stride : 0;
  if (kShared)
  {
    cudaError_t e = cudaFuncSetAttribute(lzb::lzb_decode_units_kernel<kShared, G, kFlow, kAhead>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  lzb::lzb_decode_units_kernel<kShared, G, kFlow, kAhead><<<ctas, g.warps_per_cta * 32, smem, stream>>>(
      d_packed, d_out, d_units, d_results, n_units, d_ticket, g.region_bytes, g.smem_stride, g.slab_stride, d_slab, flow);
  return cudaGetLastError();
}

/* the look-ahead form of the decoder exists for one-unit-per-warp launches with the model in shared memory only */
template <bool kShared, uint32_t G, bool kFlow>
static cudaError_t launch_variant2(const LzbGeometry &g, uint32_t ctas, const uint8_t *d_packed, uint8_t *d_out,
                                  const LzbUnit *d_units, LzbResult *d_results, uint32_t n_units, uint32_t *d_ticket,
                                  uint8_t *d_slab, const LzbFlow &flow, cudaStream_t stream)
{
  if (kShared && G == 32 && g.ahead)
    return launch_variant3<kShared, G, kFlow, kShared && G == 32>(g, ctas, d_packed, d_out, d_units, d_results, n_units, d_ticket, d_slab, flow, stream);
  return launch_variant3<kShared, G, kFlow, false>(g, ctas, d_packed, d_out, d_units, d_results, n_units, d_ticket, d_slab, flow, stream);
}

/* the streamed variant is a separate instantiation: the plain launch keeps its register allocation (the
 * schedule ptxas finds is sensitive: with the two flow helpers out of line the same loop ran 3 % slower) */
template <bool kShared, uint32_t G>
static cudaError_t launch_variant(const LzbGeometry &g, uint32_t ctas, const uint8_t *d_packed, uint8_t *d_out,
                                  const LzbUnit *d_units, LzbResult *d_results, uint32_t n_units, uint32_t *d_ticket,
                                  uint8_t *d_slab, const LzbFlow &flow, cudaStream_t stream)
{
  if (flow.arrived) return launch_variant2<kShared, G, true>(g, ctas, d_packed, d_out, d_units, d_results, n_units, d_ticket, d_slab, flow, stream);
  return launch_variant2<kShared, G, false>(g, ctas, d_packed, d_out, d_units, d_results, n_units, d_ticket, d_slab, flow, stream);
}

extern "C" int lzb_launch_decode(const uint8_t *d_packed, uint8_t *d_out, const LzbUnit *d_units,
                                 LzbResult *d_results, uint32_t n_units, uint32_t *d_ticket,
                                 uint32_t smem_model_bytes, uint32_t gmem_model_bytes, uint8_t *d_model_slab,
                                 cudaStream_t stream)
{
  return lzb_launch_decode_flow(d_packed, d_out, d_units, d_results, n_units, d_ticket, smem_model_bytes, gmem_model_bytes,
                                d_model_slab, nullptr, stream);
}

extern "C" int lzb_launch_decode_flow(const uint8_t *d_packed, uint8_t *d_out, const LzbUnit *d_units,
                                      LzbResult *d_results, uint32_t n_units, uint32_t *d_ticket,
                                      uint32_t smem_model_bytes, uint32_t gmem_model_bytes, uint8_t *d_model_slab,
                                      const LzbFlow *flow_in, cudaStream_t stream)
{
  LzbFlow flow;
  memset(&flow, 0, sizeof flow);
  if (flow_in) flow = *flow_in;
  LzbGeometry g;
  int rc = lzb_launch_geometry(smem_model_bytes, gmem_model_bytes, &g);
  if (rc) return rc;
  if (n_units == 0) return 0;
  if (!d_model_slab) return (int)cudaErrorInvalidValue;
  static const bool group_forced = getenv("LZB_GROUP") != nullptr;
  static const bool dense = getenv("LZB_SMEM_MATCHED") != nullptr && atoi(getenv("LZB_SMEM_MATCHED")) == 0;
  lzb_plan_few_units(&g, n_units, g_dev_smem_optin, !group_forced, dense);
  uint32_t ctas = g.ctas;
  if ((uint64_t)ctas * g.units_per_cta > n_units) ctas = (n_units + g.units_per_cta - 1) / g.units_per_cta;
  cudaError_t e = cudaMemsetAsync(d_ticket, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return (int)e;
#define LZB_LAUNCH(S, GG) launch_variant<S, GG>(g, ctas, d_packed, d_out, d_units, d_results, n_units, d_ticket, d_model_slab, flow, stream)
  if (!g.shared_model) e = LZB_LAUNCH(false, 32);
  else if (g.group == 32) e = LZB_LAUNCH(true, 32);
  else if (g.group == 16) e = LZB_LAUNCH(true, 16);
  else if (g.group == 8) e = LZB_LAUNCH(true, 8);
  else e = LZB_LAUNCH(true, 4);
#undef LZB_LAUNCH
  return (int)e;
}

extern "C" int lzb_launch_post_only(uint8_t *d_data, const LzbUnit *d_units, LzbResult *d_results, uint32_t n_units,
                                    cudaStream_t stream)
{
  if (n_units == 0) return 0;
  const uint32_t warps = 4;
  lzb::lzb_post_only_kernel<<<(n_units + warps - 1) / warps, warps * 32, 0, stream>>>(d_data, d_units, d_results, n_units);
  return (int)cudaGetLastError();
}

extern "C" int lzb_launch_sha256(const uint8_t *d_data, LzbShaJob *d_jobs, uint32_t n_jobs, cudaStream_t stream)
{
  if (n_jobs == 0) return 0;
  lzb::lzb_sha256_kernel<<<(n_jobs + 31) / 32, 32, 0, stream>>>(d_data, d_jobs, n_jobs);
  return (int)cudaGetLastError();
}

extern "C" int lzb_launch_bcj2(const uint8_t *d_packed, uint8_t *d_out, LzbBcj2Job *d_jobs, uint32_t n_jobs, cudaStream_t stream)
{
  if (n_jobs == 0) return 0;
  lzb::lzb_bcj2_kernel<<<(n_jobs + 3) / 4, 128, 0, stream>>>(d_packed, d_out, d_jobs, n_jobs);
  return (int)cudaGetLastError();
}
