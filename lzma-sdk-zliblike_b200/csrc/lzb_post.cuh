/* lzb_post.cuh -- what runs on a unit's bytes right after its decode, by the same warp, in the same
 * launch ("fused after decode"): the x86 BCJ / Delta post-filter in place, then CRC-32 / CRC-64.
 *
 * The warp's probability region in shared memory is dead once the unit is decoded; it is reused as
 * scratch here (CRC tables, or the Delta filter's per-lane column sums) -- kPostScratchBytes of it.
 *
 *  x86 BCJ  (Bra86.c:11-85)   sequential state (3-bit history of recent E8/E9 candidates) resets after
 *           four bytes without an opcode candidate, and a conversion never reaches across such a gap;
 *           every lane therefore starts at the first such "safe point" at or after its nominal
 *           boundary and runs the scalar converter up to the next lane's start.
 *  Delta    (Delta.c:43-62)   out[i] = in[i] + out[i-d] is a prefix sum per residue class mod d:
 *           per-lane column sums -> exclusive scan across lanes -> second sweep.
 *  CRC      (7zCrc.c, 7zCrcOpt.c, XzCrc64.c)  per-lane register over a contiguous segment, then the
 *           32 partial registers are aligned with x^(8*bytes_after) mod P and XOR-reduced.
 */
#pragma once

#include "lzb_core.cuh"

namespace lzb {

constexpr uint32_t kPostScratchBytes = 8192;                 /* Delta column sums: 256 x 32 bytes   */
constexpr uint32_t kPostSmemBytes = 2048;                    /* CRC-64 table: 256 x 8 bytes         */
constexpr uint32_t kPoly32 = 0xEDB88320u;                    /* 7zCrc.c:7  */
constexpr uint64_t kPoly64 = 0xC96C5795D7870F42ull;          /* XzCrc64.c:6 */

#ifdef LZB_EMU
template <class T> static inline T lzb_shfl_xor(T v, int) { return v; }
template <class T> static inline T lzb_shfl_down(T v, int) { return v; }
#else
template <class T> __device__ __forceinline__ T lzb_shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <class T> __device__ __forceinline__ T lzb_shfl_down(T v, int d) { return __shfl_down_sync(0xffffffffu, v, d); }
#endif

/* ------------------------------------------------------------------ x86 BCJ, decode direction --- */
LZB_FN bool is_opcode(uint8_t b) { return (b & 0xFE) == 0xE8; }
LZB_FN bool is_fill(uint8_t b) { return b == 0x00 || b == 0xFF; }

/* scalar converter over candidate positions [from, to) (Bra86.c:18-81).  hist/last carry the
 * 3-bit neighbour history and the previous candidate position in and out. */
LZB_FN uint32_t x86_convert_range(uint8_t *data, uint32_t from, uint32_t to, uint32_t ip, bool encoding,
                                  uint32_t &hist, uint32_t &last)
{
  /* kMaskToAllowedStatus / kMaskToBitNumber (Bra86.c:8-9) packed into immediates */
  const uint32_t allow = 0x17;          /* bit h set  <=> history h allows a conversion: {1,1,1,0,1,0,0,0} */
  const uint32_t topbit = 0x6DB488u;    /* 3 bits per h: {0,1,2,2,3,3,3,3} */
  uint32_t pos = from;
  ip += 5;
  for (;;)
  {
    while (pos < to && !is_opcode(data[pos])) pos++;
    if (pos >= to) break;
    const uint32_t gap = pos - last;
    last = pos;
    if (gap > 3) hist = 0;
    else
    {
      hist = (hist << (gap - 1)) & 7;
      if (hist != 0)
      {
        const uint8_t b = data[pos + 4 - ((topbit >> (3 * hist)) & 7)];
        if (!((allow >> hist) & 1) || is_fill(b)) { hist = ((hist << 1) & 7) | 1; pos++; continue; }
      }
    }
    if (is_fill(data[pos + 4]))
    {
      uint32_t v = ((uint32_t)data[pos + 4] << 24) | ((uint32_t)data[pos + 3] << 16) |
                   ((uint32_t)data[pos + 2] << 8) | data[pos + 1];
      const uint32_t here = ip + pos;
      uint32_t out;
      for (;;)
      {
        out = encoding ? here + v : v - here;
        if (hist == 0) break;
        const uint32_t sh = ((topbit >> (3 * hist)) & 7) * 8;
        if (!is_fill((uint8_t)(out >> (24 - sh)))) break;
        v = out ^ ((1u << (32 - sh)) - 1);
      }
      data[pos + 4] = (uint8_t)(0u - ((out >> 24) & 1));
      data[pos + 3] = (uint8_t)(out >> 16);
      data[pos + 2] = (uint8_t)(out >> 8);
      data[pos + 1] = (uint8_t)out;
      pos += 5;
    }
    else { hist = ((hist << 1) & 7) | 1; pos++; }
  }
  return pos;
}

/* Whole-buffer x86_Convert.  Returns the processed byte count; *state in/out as in the reference. */
LZB_FN uint32_t x86_convert_warp(uint8_t *data, uint32_t size, uint32_t ip, bool encoding, uint32_t &state)
{
  if (size < 5) return 0;                               /* Bra86.c:15-16 (state untouched) */
  const uint32_t stop = size - 4;                       /* candidates live in [0, stop) */
  const uint32_t lane = lzb_lane();
  const uint32_t seg = (stop + LZB_LANES - 1) / LZB_LANES;
  uint32_t start;
  lzb_syncwarp();
  if (lane == 0) start = 0;
  else
  {
    /* first p >= lane*seg with no opcode candidate in [p-4, p-1] (bytes before 0 count as clean) */
    const uint32_t nominal = lane * seg < stop ? lane * seg : stop;
    uint32_t q = nominal >= 4 ? nominal - 4 : 0;
    uint32_t clean = 4 - (nominal - q);
    start = stop;
    for (; q < stop + 3; q++)
    {
      if (clean >= 4) { start = q; break; }
      clean = is_opcode(data[q]) ? 0 : clean + 1;
    }
    if (start > stop) start = stop;
  }
  uint32_t next = lzb_shfl_down(start, 1);
  if (lane == LZB_LANES - 1) next = stop;
  lzb_syncwarp();                                        /* all safe points found on unmodified bytes */
  uint32_t hist = lane == 0 ? (state & 7) : 0;
  uint32_t last = lane == 0 ? 0xFFFFFFFFu : start - 5;   /* lane 0: prevPosT = -1 (Bra86.c:18); others: far */
  uint32_t pos = start;
  const bool busy = start < next;
  if (busy) pos = x86_convert_range(data, start, next, ip, encoding, hist, last);
  lzb_syncwarp();
  /* the scan's final (pos, last, hist) are those of the last lane that owned a range */
#ifndef LZB_EMU
  const uint32_t owners = __ballot_sync(0xffffffffu, busy);
  const int top = 31 - __clz((int)owners);
  pos = __shfl_sync(0xffffffffu, pos, top);
  last = __shfl_sync(0xffffffffu, last, top);
  hist = __shfl_sync(0xffffffffu, hist, top);
#endif
  const uint32_t gap = pos - last;
  state = gap > 3 ? 0 : ((hist << (gap - 1)) & 7);      /* Bra86.c:82-83 */
  return pos;
}

/* ------------------------------------------------------------------ Delta ---------------------- */
LZB_FN void delta_decode_warp(uint8_t *data, uint32_t size, uint32_t dist, uint8_t *acc /* dist*LANES bytes */)
{
  const uint32_t lane = lzb_lane();
  const uint32_t rows = (size + dist - 1) / dist;
  const uint32_t seg = ((rows + LZB_LANES - 1) / LZB_LANES) * dist;
  const uint32_t begin = (uint64_t)lane * seg < size ? lane * seg : size;
  const uint32_t end = size - begin < seg ? size : begin + seg;
  lzb_syncwarp();
  for (uint32_t c = 0; c < dist; c++) acc[c * LZB_LANES + lane] = 0;
  for (uint32_t i = begin, c = 0; i < end; i++)
  {
    acc[c * LZB_LANES + lane] = (uint8_t)(acc[c * LZB_LANES + lane] + data[i]);
    if (++c == dist) c = 0;
  }
  lzb_syncwarp();
  for (uint32_t c = lane; c < dist; c += LZB_LANES)     /* exclusive scan across lanes, per column */
  {
    uint8_t run = 0;
    for (uint32_t l = 0; l < LZB_LANES; l++)
    {
      const uint8_t t = acc[c * LZB_LANES + l];
      acc[c * LZB_LANES + l] = run;
      run = (uint8_t)(run + t);
    }
  }
  lzb_syncwarp();
  for (uint32_t i = begin, c = 0; i < end; i++)
  {
    const uint8_t v = (uint8_t)(acc[c * LZB_LANES + lane] + data[i]);
    acc[c * LZB_LANES + lane] = v;
    data[i] = v;
    if (++c == dist) c = 0;
  }
  lzb_syncwarp();
}

/* ------------------------------------------------------------------ CRC ------------------------ */
/* products in GF(2)[x]/P, reflected bit order: the top bit is x^0 */
LZB_FN uint32_t mulmod32(uint32_t a, uint32_t b)
{
  uint32_t p = 0;
  for (int k = 0; k < 32; k++)
  {
    p ^= b & (0u - ((a >> (31 - k)) & 1));
    b = (b >> 1) ^ (kPoly32 & (0u - (b & 1)));
  }
  return p;
}
LZB_FN uint64_t mulmod64(uint64_t a, uint64_t b)
{
  uint64_t p = 0;
  for (int k = 0; k < 64; k++)
  {
    p ^= b & (0ull - ((a >> (63 - k)) & 1));
    b = (b >> 1) ^ (kPoly64 & (0ull - (b & 1)));
  }
  return p;
}
/* x^(8n) mod P */
LZB_FN uint32_t xpow32(uint32_t n)
{
  uint32_t r = 1u << 31, sq = 1u << 23;
  for (; n != 0; n >>= 1) { if (n & 1) r = mulmod32(r, sq); sq = mulmod32(sq, sq); }
  return r;
}
LZB_FN uint64_t xpow64(uint32_t n)
{
  uint64_t r = 1ull << 63, sq = 1ull << 55;
  for (; n != 0; n >>= 1) { if (n & 1) r = mulmod64(r, sq); sq = mulmod64(sq, sq); }
  return r;
}

LZB_FN void lane_segment(uint32_t size, uint32_t &begin, uint32_t &end)
{
  const uint32_t lane = lzb_lane();
  uint32_t seg = (size + LZB_LANES - 1) / LZB_LANES;
  seg = (seg + 15u) & ~15u;
  begin = (uint64_t)lane * seg < size ? lane * seg : size;
  end = size - begin < seg ? size : begin + seg;
}

/* CRC-32 as CrcCalc (7zCrc.c:44): preset all-ones, final inversion */
LZB_FN uint32_t crc32_warp(const uint8_t *data, uint32_t size, uint32_t *table /* 256 entries */)
{
  const uint32_t lane = lzb_lane();
  lzb_syncwarp();
  for (uint32_t i = lane; i < 256; i += LZB_LANES)          /* CrcGenerateTable, 7zCrc.c:49-60 */
  {
    uint32_t r = i;
    for (int k = 0; k < 8; k++) r = (r >> 1) ^ (kPoly32 & (0u - (r & 1)));
    table[i] = r;
  }
  lzb_syncwarp();
  uint32_t begin, end;
  lane_segment(size, begin, end);
  uint32_t v = lane == 0 ? 0xFFFFFFFFu : 0u;
  for (uint32_t i = begin; i < end; i++) v = table[(v ^ data[i]) & 0xFF] ^ (v >> 8);
  v = mulmod32(v, xpow32(size - end));
  for (uint32_t m = LZB_LANES >> 1; m != 0; m >>= 1) v ^= lzb_shfl_xor(v, (int)m);
  lzb_syncwarp();
  return v ^ 0xFFFFFFFFu;
}

/* CRC-64/XZ as Crc64Calc (XzCrc64.c:30) */
LZB_FN uint64_t crc64_warp(const uint8_t *data, uint32_t size, uint64_t *table /* 256 entries */)
{
  const uint32_t lane = lzb_lane();
  lzb_syncwarp();
  for (uint32_t i = lane; i < 256; i += LZB_LANES)          /* Crc64GenerateTable, XzCrc64.c:9-20 */
  {
    uint64_t r = i;
    for (int k = 0; k < 8; k++) r = (r >> 1) ^ (kPoly64 & (0ull - (r & 1)));
    table[i] = r;
  }
  lzb_syncwarp();
  uint32_t begin, end;
  lane_segment(size, begin, end);
  uint64_t v = lane == 0 ? ~0ull : 0ull;
  for (uint32_t i = begin; i < end; i++) v = table[(v ^ data[i]) & 0xFF] ^ (v >> 8);
  v = mulmod64(v, xpow64(size - end));
  for (uint32_t m = LZB_LANES >> 1; m != 0; m >>= 1) v ^= lzb_shfl_xor(v, (int)m);
  lzb_syncwarp();
  return ~v;
}

/* filter -> check -> verdict, for one decoded unit.  `scratch` = the warp's dead probability region. */
/* `scratch`     : >= kPostSmemBytes of the warp's (dead) shared-memory model region -- CRC table.
 * `big_scratch` : >= kPostScratchBytes, the warp's (dead) matched-literal slab slice in HBM/L2 -- Delta
 *                 column sums (up to 256 x 32 bytes; Delta blocks are rare, L2 latency is acceptable). */
LZB_FN void post_process(const LzbUnit &u, LzbResult &r, uint8_t *data, uint8_t *scratch, uint8_t *big_scratch)
{
  if (r.res != LZB_SZ_OK) return;
  const uint32_t n = r.dst_done;
  if (u.filter == LZB_FILTER_X86)
  {
    /* stand-alone x86_Convert calls pass the direction in flags bit 1 and the history in `expect` */
    uint32_t st = (uint32_t)(u.expect >> 32) & 7;
    const uint32_t done = x86_convert_warp(data, n, u.filter_param, (u.flags & 2) != 0, st);
    if (u.flags & 4) { r.src_used = done; r.status = (int32_t)st; }   /* stand-alone call: report them */
  }
  else if (u.filter == LZB_FILTER_DELTA) delta_decode_warp(data, n, u.filter_param, big_scratch);
  if (u.check == LZB_CHECK_CRC32) r.check = crc32_warp(data, n, reinterpret_cast<uint32_t *>(scratch));
  else if (u.check == LZB_CHECK_CRC64) r.check = crc64_warp(data, n, reinterpret_cast<uint64_t *>(scratch));
  if ((u.flags & LZB_UNIT_VERIFY) && u.check != LZB_CHECK_NONE && r.check != u.expect) r.res = LZB_SZ_ERROR_CRC;
}

}  // namespace lzb
