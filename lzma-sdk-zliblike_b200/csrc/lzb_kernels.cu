/* lzb_kernels.cu -- sm_100a kernels and their launchers.
 *
 *   lzb_decode_units_kernel : persistent grid (one CTA per SM), every warp pulls the next unit from a
 *                             global ticket, decodes it (LZMA / LZMA2), applies the unit's post-filter
 *                             and check in the same launch and writes one LzbResult.
 *
 * Residency.  The bit loop is bound by issue slots and dependent-issue latency (profiles/), so what it
 * needs is resident warps per scheduler, and two things cap them: shared memory for the models and the
 * 64 K-entry register file (all 32 lanes of a warp carry the same decoder state).  Hence:
 *   - only the part of the model a plain symbol touches lives in shared memory (6,448 B at lc3/lp0/pb2);
 *     the matched-literal trees (8 KiB, touched once per match) live in an L2-resident HBM slab;
 *   - the kernel is built for kMaxWarps in {16, 24, 32} resident warps per SM with __launch_bounds__
 *     (kMaxWarps*32, 1), i.e. <= 128 / 80 / 64 registers per thread; the launcher picks the largest
 *     variant the shared-memory model allows (LZB_WARPS overrides, for measurements).
 * Models whose shared part exceeds 56 KiB (raw LZMA with lc+lp >= 6) live entirely in the slab.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "lzb_core.cuh"
#include "lzb_post.cuh"
#include "lzb_kernels.h"

namespace lzb {

template <bool kSharedModel, int kMaxWarps>
__global__ void __launch_bounds__(kMaxWarps * 32, 1)
lzb_decode_units_kernel(const uint8_t *__restrict__ packed, uint8_t *__restrict__ out,
                        const LzbUnit *__restrict__ units, LzbResult *__restrict__ results, uint32_t n_units,
                        uint32_t *ticket, uint32_t region_bytes, uint32_t slab_stride, uint8_t *model_slab)
{
  extern __shared__ __align__(16) uint8_t lzb_smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t warps_per_cta = blockDim.x >> 5;
  uint8_t *slab = model_slab + ((size_t)blockIdx.x * warps_per_cta + warp) * slab_stride;
  /* shared model: region in shared memory, matched-literal trees at the start of the warp's slab slice;
   * global model: both in the slab slice (region first) */
  uint8_t *region = kSharedModel ? lzb_smem + (size_t)warp * region_bytes : slab;
  uint16_t *matched = reinterpret_cast<uint16_t *>(kSharedModel ? slab : slab + region_bytes);
  for (;;)
  {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(ticket, 1u);
    i = __shfl_sync(0xffffffffu, i, 0);
    if (i >= n_units) break;
    const LzbUnit u = units[i];
    UnitDecoder d;
    LzbResult r;
    uint8_t *dst = out + u.dst_off;
    const bool stateful = (u.flags & LZB_UNIT_STATEFUL) != 0;
    d.begin(packed + u.src_off, u.src_len, dst, stateful ? u.filter_param : u.dst_cap, u.dict_size,
            reinterpret_cast<uint16_t *>(region), matched, kSharedModel);
    decode_unit(d, u, r, stateful ? reinterpret_cast<LzbStreamState *>(u.expect) : nullptr);
    if (!stateful) post_process(u, r, dst, region, reinterpret_cast<uint8_t *>(matched));
    if (lane == 0) results[i] = r;
    __syncwarp();
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
  post_process(u, r, data + u.dst_off, scratch[warp], scratch[warp]);
  if (lane == 0) results[i] = r;
}

}  // namespace lzb

/* ------------------------------------------------------------------ launchers (C linkage) ------ */
static int g_dev_sms = 0;
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
  g->ctas = (uint32_t)g_dev_sms;
  /* measured on B200 (profiles/r01_c_residency.md): 16 / 24 / 32 warps per SM = 5.27 / 5.82 / 5.68 GB/s on the
   * 64 KiB raw-LZMA batch; 24 (80 registers, no hot-loop spills) is the default, LZB_WARPS overrides */
  uint32_t want = LZB_DEFAULT_WARPS_PER_CTA;
  const char *ov = getenv("LZB_WARPS");
  if (ov && atoi(ov) > 0) want = (uint32_t)atoi(ov);
  if (region <= LZB_MAX_SHARED_MODEL_BYTES)
  {
    uint32_t w = (uint32_t)(g_dev_smem_optin / region);
    if (w > want) w = want;
    if (w > LZB_MAX_WARPS_PER_CTA) w = LZB_MAX_WARPS_PER_CTA;
    if (w < 1) w = 1;
    g->warps_per_cta = w;
    g->shared_model = 1;
    g->slab_stride = gpart;
  }
  else
  {
    g->warps_per_cta = 8;
    g->shared_model = 0;
    g->slab_stride = region + gpart;
  }
  g->slab_bytes = (size_t)g->slab_stride * g->warps_per_cta * g->ctas;
  return 0;
}

template <bool kShared, int kMaxWarps>
static cudaError_t launch_variant(const LzbGeometry &g, uint32_t ctas, const uint8_t *d_packed, uint8_t *d_out,
                                  const LzbUnit *d_units, LzbResult *d_results, uint32_t n_units, uint32_t *d_ticket,
                                  uint8_t *d_slab, cudaStream_t stream)
{
  const size_t smem = kShared ? (size_t)g.warps_per_cta * g.region_bytes : 0;
  if (kShared)
  {
    cudaError_t e = cudaFuncSetAttribute(lzb::lzb_decode_units_kernel<kShared, kMaxWarps>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  lzb::lzb_decode_units_kernel<kShared, kMaxWarps><<<ctas, g.warps_per_cta * 32, smem, stream>>>(
      d_packed, d_out, d_units, d_results, n_units, d_ticket, g.region_bytes, g.slab_stride, d_slab);
  return cudaGetLastError();
}

extern "C" int lzb_launch_decode(const uint8_t *d_packed, uint8_t *d_out, const LzbUnit *d_units,
                                 LzbResult *d_results, uint32_t n_units, uint32_t *d_ticket,
                                 uint32_t smem_model_bytes, uint32_t gmem_model_bytes, uint8_t *d_model_slab,
                                 cudaStream_t stream)
{
  LzbGeometry g;
  int rc = lzb_launch_geometry(smem_model_bytes, gmem_model_bytes, &g);
  if (rc) return rc;
  if (n_units == 0) return 0;
  if (!d_model_slab) return (int)cudaErrorInvalidValue;
  uint32_t ctas = g.ctas;
  if ((uint64_t)ctas * g.warps_per_cta > n_units) ctas = (n_units + g.warps_per_cta - 1) / g.warps_per_cta;
  cudaError_t e = cudaMemsetAsync(d_ticket, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return (int)e;
  const uint32_t w = g.warps_per_cta;
  if (!g.shared_model)
    e = launch_variant<false, 16>(g, ctas, d_packed, d_out, d_units, d_results, n_units, d_ticket, d_model_slab, stream);
  else if (w <= 16)
    e = launch_variant<true, 16>(g, ctas, d_packed, d_out, d_units, d_results, n_units, d_ticket, d_model_slab, stream);
  else if (w <= 24)
    e = launch_variant<true, 24>(g, ctas, d_packed, d_out, d_units, d_results, n_units, d_ticket, d_model_slab, stream);
  else
    e = launch_variant<true, 32>(g, ctas, d_packed, d_out, d_units, d_results, n_units, d_ticket, d_model_slab, stream);
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
