/* lzb_xz.h -- xz pieces shared by lzb_containers.cu (walk, batched block decode) and lzb_xzstream.cu
 * (the resumable XzUnpacker_Code machine and its filter chain). */
#ifndef LZB_XZ_H
#define LZB_XZ_H

#include <vector>

#include "lzb_host.h"

uint32_t lzb_small_crc32(const Byte *p, size_t n);        /* host CRC-32 of container framing bytes */
uint32_t lzb_small_crc32_update(uint32_t reg, const Byte *p, size_t n);   /* raw register form */
SRes lzb_xz_classify_chain(const CXzBlock *blk, LzbXzBlock *b);

struct XzWalk
{
  std::vector<LzbXzBlock> blocks;
  std::vector<LzbUnit> units;          /* src_off absolute in the buffer, dst_off absolute in the output */
  std::vector<SRes> after_block;       /* structural error met right after block k's payload (padding) */
  SRes end_res = SZ_OK;                /* verdict if every walked block decodes and verifies */
  int end_status = CODER_STATUS_NEEDS_MORE_INPUT;
  size_t end_src = 0;
  bool finished = false;
  uint64_t out_total = 0;
  int partial = -1;                    /* index of a block whose payload is cut short / malformed */
};
void lzb_xz_walk(const Byte *src, size_t n, XzWalk &w, bool in_blocks, unsigned in_flags);

/* a block decoded ahead of the state machine */
struct XzReady
{
  uint64_t data_off;                   /* first payload byte, relative to the src of the prefetch call */
  uint64_t pack, unpack, dst_off;
  uint64_t digest;                     /* CRC-32 / CRC-64 of the block's bytes */
  uint8_t sha[32];                     /* SHA-256 of the block's bytes */
  uint8_t check_type, n_filters;
};
SRes lzb_xz_prefetch(LzbContext *c, const Byte *src, size_t n, bool in_blocks, unsigned flags, Byte *dest, size_t cap,
                     std::vector<XzReady> &ready);

#endif
