/* lzb_kernels.cu -- sm_100a kernels and their launchers.
 *
 *   lzb_decode_units_kernel : persistent grid (one CTA per SM), every warp pulls the next unit from a
 *                             global ticket, decodes it (LZMA / LZMA2), applies the unit's post-filter
 *                             and check in the same launch and writes one LzbResult.
 *
 * Residency: one CTA per SM holding W warps, W = floor((227 KiB - 1 KiB) / model bytes) capped at 16
 * (15 at lc3/lp0/pb2, 8 at the LZMA2 worst case lc+lp=4/pb=4).  Models larger than 56 KiB (raw LZMA
 * with lc+lp >= 6, legal in the reference but unused by its encoder defaults) live in an HBM scratch
 * slab instead -- same code path, template flag only.
 */
#include <cuda_runtime.h>
#include <stdint.h>

#include "lzb_core.cuh"
#include "lzb_post.cuh"
#include "lzb_kernels.h"

namespace lzb {

template <bool kSharedModel>
__global__ void __launch_bounds__(LZB_MAX_WARPS_PER_CTA * 32, 1)
lzb_decode_units_kernel(const uint8_t *__restrict__ packed, uint8_t *__restrict__ out,
                        const LzbUnit *__restrict__ units, LzbResult *__restrict__ results, uint32_t n_units,
                        uint32_t *ticket, uint32_t region_bytes, uint8_t *model_slab)
{
  extern __shared__ __align__(16) uint8_t lzb_smem[];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const uint32_t warps_per_cta = blockDim.x >> 5;
  uint8_t *region = kSharedModel ? lzb_smem + (size_t)warp * region_bytes
                                 : model_slab + ((size_t)blockIdx.x * warps_per_cta + warp) * region_bytes;
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
            reinterpret_cast<uint16_t *>(region));
    decode_unit(d, u, r, stateful ? reinterpret_cast<LzbStreamState *>(u.expect) : nullptr);
    if (!stateful) post_process(u, r, dst, region);
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
  (void)scratch;
  post_process(u, r, data + u.dst_off, scratch[warp]);
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

extern "C" int lzb_launch_geometry(uint32_t model_bytes, uint32_t *warps_per_cta, uint32_t *ctas,
                                   uint32_t *region_bytes, int *shared_model)
{
  cudaError_t e = query_device();
  if (e != cudaSuccess) return (int)e;
  uint32_t region = model_bytes < lzb::kPostScratchBytes ? lzb::kPostScratchBytes : model_bytes;
  region = (region + 15u) & ~15u;
  *region_bytes = region;
  *ctas = (uint32_t)g_dev_sms;
  if (region <= LZB_MAX_SHARED_MODEL_BYTES)
  {
    uint32_t w = (uint32_t)(g_dev_smem_optin / region);
    if (w > LZB_MAX_WARPS_PER_CTA) w = LZB_MAX_WARPS_PER_CTA;
    if (w < 1) w = 1;
    *warps_per_cta = w;
    *shared_model = 1;
  }
  else
  {
    *warps_per_cta = 8;
    *shared_model = 0;
  }
  return 0;
}

extern "C" int lzb_launch_decode(const uint8_t *d_packed, uint8_t *d_out, const LzbUnit *d_units,
                                 LzbResult *d_results, uint32_t n_units, uint32_t *d_ticket,
                                 uint32_t model_bytes, uint8_t *d_model_slab, cudaStream_t stream)
{
  uint32_t w, ctas, region; int shared_model;
  int rc = lzb_launch_geometry(model_bytes, &w, &ctas, &region, &shared_model);
  if (rc) return rc;
  if (n_units == 0) return 0;
  /* no point in launching more warps than units */
  if ((uint64_t)ctas * w > n_units)
  {
    ctas = (n_units + w - 1) / w;
  }
  cudaError_t e = cudaMemsetAsync(d_ticket, 0, sizeof(uint32_t), stream);
  if (e != cudaSuccess) return (int)e;
  if (shared_model)
  {
    const size_t smem = (size_t)w * region;
    e = cudaFuncSetAttribute(lzb::lzb_decode_units_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    lzb::lzb_decode_units_kernel<true><<<ctas, w * 32, smem, stream>>>(d_packed, d_out, d_units, d_results, n_units,
                                                                      d_ticket, region, nullptr);
  }
  else
  {
    if (!d_model_slab) return (int)cudaErrorInvalidValue;
    lzb::lzb_decode_units_kernel<false><<<ctas, w * 32, 0, stream>>>(d_packed, d_out, d_units, d_results, n_units,
                                                                    d_ticket, region, d_model_slab);
  }
  return (int)cudaGetLastError();
}

extern "C" int lzb_launch_post_only(uint8_t *d_data, const LzbUnit *d_units, LzbResult *d_results, uint32_t n_units,
                                    cudaStream_t stream)
{
  if (n_units == 0) return 0;
  const uint32_t warps = 4;
  lzb::lzb_post_only_kernel<<<(n_units + warps - 1) / warps, warps * 32, 0, stream>>>(d_data, d_units, d_results, n_units);
  return (int)cudaGetLastError();
}
