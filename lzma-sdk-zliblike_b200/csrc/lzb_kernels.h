/* lzb_kernels.h -- launcher prototypes between lzb_kernels.cu and the host layer (lzb_host.cu). */
#ifndef LZB_KERNELS_H
#define LZB_KERNELS_H

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "lzb_types.h"

#define LZB_DEFAULT_GROUP 16            /* lanes per unit: 16 = two units per warp (measured best, profiles/r01_c_residency.md) */
#define LZB_DEFAULT_UNITS_PER_SM 36
#define LZB_MAX_SHARED_MODEL_BYTES (56u * 1024u)

#ifdef __cplusplus
extern "C" {
#endif

typedef struct LzbGeometry
{
  uint32_t group;           /* lanes per unit (G)                                        */
  uint32_t warps_per_cta;
  uint32_t units_per_cta;   /* resident units per SM = warps_per_cta * 32 / group        */
  uint32_t ctas;            /* one per SM                                                */
  uint32_t region_bytes;    /* shared-memory bytes per unit (model / post-pass scratch)  */
  uint32_t smem_stride;     /* shared-memory bytes between units: region_bytes, or region_bytes + slab_stride when a
                               one-unit-per-warp launch keeps the matched-literal trees in shared memory too */
  uint32_t slab_stride;     /* HBM slab bytes per unit (matched-literal trees, or the whole model) */
  int shared_model;
  int ahead;                /* one-unit-per-warp launch with few enough warps per SM: look-ahead form of the decoder */
  size_t slab_bytes;        /* slab the caller must provide: ctas * units_per_cta * slab_stride */
} LzbGeometry;

/* Residency the decode kernel will use for a model of the given shared / HBM part sizes (bytes). */
int lzb_launch_geometry(uint32_t smem_model_bytes, uint32_t gmem_model_bytes, LzbGeometry *g);
/* What a launch of n_units does with that geometry when the units are fewer than the warps: one unit per warp,
 * spread over all SMs, matched-literal trees in shared memory when they fit (pure arithmetic, no device access). */
void lzb_plan_few_units(LzbGeometry *g, uint32_t n_units, size_t smem_optin, int regroup_allowed, int dense);

/* A launch that is fed and drained while it runs (Lzb_DecodeUnitsHost): units are taken in index order, in
 * chunks of chunk_units; a unit of chunk c starts once arrived[min(c+1, n_chunks-1)] is non-zero (the packed
 * bytes are copied in chunk order by another stream, each copy followed by its flag), and the group that
 * finishes the last unit of a chunk sets ready[c] in mapped host memory so the host can start that chunk's
 * download.  All-null = an ordinary launch. */
typedef struct LzbFlow
{
  const uint32_t *arrived;   /* device memory, n_chunks words                         */
  uint32_t *done;            /* device memory, n_chunks counters (zeroed by the host) */
  uint32_t *ready;           /* mapped host memory, n_chunks words                    */
  const uint32_t *abort;     /* mapped host memory, one word: non-zero = give up      */
  uint32_t chunk_units, n_chunks;
} LzbFlow;

/* Decode n_units units.  d_ticket: one uint32 of device scratch.  d_model_slab: g.slab_bytes bytes. */
int lzb_launch_decode(const uint8_t *d_packed, uint8_t *d_out, const LzbUnit *d_units, LzbResult *d_results,
                      uint32_t n_units, uint32_t *d_ticket, uint32_t smem_model_bytes, uint32_t gmem_model_bytes,
                      uint8_t *d_model_slab, cudaStream_t stream);

int lzb_launch_decode_flow(const uint8_t *d_packed, uint8_t *d_out, const LzbUnit *d_units, LzbResult *d_results,
                           uint32_t n_units, uint32_t *d_ticket, uint32_t smem_model_bytes, uint32_t gmem_model_bytes,
                           uint8_t *d_model_slab, const LzbFlow *flow, cudaStream_t stream);

/* Filter/check only, over bytes already in device memory (dst_off/dst_cap/filter/check of each unit). */
int lzb_launch_post_only(uint8_t *d_data, const LzbUnit *d_units, LzbResult *d_results, uint32_t n_units,
                         cudaStream_t stream);

/* BCJ2 merges over streams in device memory, one thread per job (jobs: device array, res written back). */
int lzb_launch_bcj2(const uint8_t *d_packed, uint8_t *d_out, LzbBcj2Job *d_jobs, uint32_t n_jobs, cudaStream_t stream);

/* SHA-256 over bytes in device memory, one thread per job (jobs: device array, updated in place). */
int lzb_launch_sha256(const uint8_t *d_data, LzbShaJob *d_jobs, uint32_t n_jobs, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif
