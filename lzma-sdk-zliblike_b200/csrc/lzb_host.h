/* lzb_host.h -- internals shared by the host translation units (context, staging, error text). */
#ifndef LZB_HOST_H
#define LZB_HOST_H

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "../../include/lzb200.h"
#include "lzb_kernels.h"

#define LZB_SLOTS 3

struct LzbSlot
{
  cudaStream_t stream = nullptr;
  uint8_t *d_packed = nullptr; size_t packed_cap = 0;
  uint8_t *d_out = nullptr;    size_t out_cap = 0;
  LzbUnit *d_units = nullptr;  LzbResult *d_results = nullptr; size_t units_cap = 0;
  LzbUnit *h_units = nullptr;  LzbResult *h_results = nullptr; size_t h_units_cap = 0;   /* pinned */
  uint32_t *d_ticket = nullptr;
  uint8_t *d_slab = nullptr;   size_t slab_cap = 0;
};

struct LzbContext
{
  int device = 0;
  int sms = 0;
  LzbSlot slot[LZB_SLOTS];
  /* device-resident API: its own ticket ring (one counter per launch in flight) and model slab */
  uint32_t *d_dev_tickets = nullptr;
  unsigned dev_seq = 0;
  uint8_t *d_dev_slab = nullptr; size_t dev_slab_cap = 0;
  cudaEvent_t dev_event = nullptr;  /* recorded behind the last device-resident launch (they are ordered, lzb_host.cu) */
  cudaStream_t dev_last_stream = nullptr;
  /* streamed launch (LzbFlow): device flags/counters, mapped host flags, a pinned constant, an event */
  uint32_t *d_flow = nullptr; uint32_t *h_flow = nullptr; uint32_t *h_flow_dev = nullptr; size_t flow_cap = 0;
  uint32_t *h_one = nullptr;
  cudaEvent_t flow_event = nullptr;
  bool flow_off = false;            /* uploads did not overlap a running kernel once: use sliced launches */
  std::mutex mu;
};

void lzb_set_error(const char *fmt, ...);
bool lzb_cuda_ok(cudaError_t e, const char *what);
LzbContext *lzb_default_context();                 /* lazily created; NULL (+ error text) without a GPU */
bool lzb_slot_reserve(LzbContext *c, LzbSlot &s, size_t packed, size_t out, size_t units, size_t slab);

struct LzbCopyRange { uint64_t src_off, dst_off, len; };
/* devOutBytes >= outBytes: size of the device output arena (scratch streams may live behind the part that is
 * downloaded); bcj2: merges that run between the decode pass and the folder pass */
SRes lzb_run_stages(LzbContext *c, const Byte *packed, size_t packedBytes, Byte *out, size_t outBytes, size_t devOutBytes,
                    const LzbCopyRange *copies, size_t n_copies,
                    const LzbUnit *units, LzbResult *results, size_t n,
                    LzbBcj2Job *bcj2, size_t n_bcj2,
                    const LzbUnit *post_a, LzbResult *res_a, size_t n_a,
                    const LzbUnit *post_b, LzbResult *res_b, size_t n_b);

SRes lzb_flow_then_checks(LzbContext *c, const Byte *packed, size_t packedBytes, Byte *out, size_t outBytes,
                          const LzbUnit *units, LzbResult *results, size_t n,
                          const LzbUnit *checks, LzbResult *check_results, size_t n_checks, bool *used);

SRes lzb_run_resident(LzbContext *c, const uint8_t *d_packed, uint8_t *d_out, const LzbUnit *units, LzbResult *results, size_t n);

/* one call of a resumable decoder object: block = [LzbUnit | LzbResult | input bytes] at these offsets */
#define LZB_STATEFUL_RESULT_AT 64
#define LZB_STATEFUL_INPUT_AT 128
SRes lzb_run_stateful(LzbContext *c, uint8_t *d_block, const uint8_t *h_block, size_t block_bytes, const LzbUnit *unit,
                      uint8_t *d_state, LzbStreamState *h_state, LzbResult *h_result,
                      const uint8_t *d_dic, Byte *dic, size_t dic_from, size_t guess);

/* little helpers */
static inline uint32_t lzb_le32(const Byte *p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint64_t lzb_le64(const Byte *p) { return (uint64_t)lzb_le32(p) | ((uint64_t)lzb_le32(p + 4) << 32); }

/* host-side GF(2) helpers for stitching CRCs of independently processed segments */
uint32_t lzb_crc32_shift(uint32_t reg, uint64_t nbytes);   /* reg * x^(8n) mod P32 */
uint64_t lzb_crc64_shift(uint64_t reg, uint64_t nbytes);

#endif
