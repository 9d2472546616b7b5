/* lzb_post.cuh -- what runs on a unit's bytes right after its decode, by the same G-lane group, in the
 * same launch ("fused after decode"): the x86 BCJ / Delta post-filter in place, then CRC-32 / CRC-64.
 *
 * The group's model storage is dead once the unit is decoded and is reused as scratch here: the CRC
 * table (kPostSmemBytes of the shared-memory region) and the Delta filter's per-lane column sums
 * (kPostScratchBytes of the matched-literal slab slice in HBM/L2).
 *
 *  x86 BCJ  (Bra86.c:11-85)   the opcode candidates (E8/E9) have to be decided in order (3-bit history), but they are
 *           sparse: the group finds them with coalesced 16-byte loads per lane, a tile of G x 16 bytes at a time, and
 *           walks the marked positions in order (ballot + broadcast), every lane running the same decision.
 *  Delta    (Delta.c:43-62)   out[i] = in[i] + out[i-d] is a prefix sum per residue class mod d:
 *           per-lane column sums -> exclusive scan across lanes -> second sweep.
 *  CRC      (7zCrc.c, 7zCrcOpt.c, XzCrc64.c)  slicing tables in the group's scratch (four bytes per dependent step
 *           when they fit), per-lane register over a contiguous segment, then the G partial registers are aligned
 *           with x^(8*bytes_after) mod P and XOR-reduced.
 */
#pragma once

#include "lzb_core.cuh"

namespace lzb {

constexpr uint32_t kPostScratchBytes = 8192;                 /* Delta column sums: 256 x 32 bytes   */
constexpr uint32_t kPostSmemBytes = 2048;                    /* CRC-64 table: 256 x 8 bytes         */
constexpr uint32_t kPoly32 = 0xEDB88320u;                    /* 7zCrc.c:7  */
constexpr uint64_t kPoly64 = 0xC96C5795D7870F42ull;          /* XzCrc64.c:6 */

/* ------------------------------------------------------------------ x86 BCJ ---------------------- */
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

/* Whole-buffer x86_Convert, one contiguous segment per lane (round 1).  Still what the LEAN decode kernels carry: their
 * hot loop's register allocation depends on the cold code inlined next to it, and with the tile scan below in its place
 * the dense launches ran 3 % (inlined) to 8 % (out of line) slower on workloads that have no filter at all
 * (profiles/r02_z_x86_tiles.md); their SASS is byte-identical to the library of the sessions before. */
template <uint32_t G> LZB_FN uint32_t x86_convert_lanes(uint8_t *data, uint32_t size, uint32_t ip, bool encoding, uint32_t &state)
{
  typedef Group<G> Grp;
  if (size < 5) return 0;                               /* Bra86.c:15-16 (state untouched) */
  const uint32_t stop = size - 4;                       /* candidates live in [0, stop) */
  const uint32_t lane = Grp::lane();
  const uint32_t seg = (stop + G - 1) / G;
  uint32_t start;
  Grp::sync();
  if (lane == 0) start = 0;
  else
  {
    /* first p >= lane*seg with no opcode candidate in [p-4, p-1].  Bytes before 0 count as clean only when no
     * history is carried in: with (state & 7) != 0 the reference starts from prevPosT = -1 and the carried mask
     * (Bra86.c:18-19), which still reaches candidates at positions 0..2 -- those belong to lane 0 alone. */
    const uint32_t nominal = lane * seg < stop ? lane * seg : stop;
    uint32_t q = nominal >= 4 ? nominal - 4 : 0;
    uint32_t clean = (state & 7) != 0 ? 0u : 4 - (nominal - q);
    start = stop;
    for (; q < stop + 3; q++)
    {
      if (clean >= 4) { start = q; break; }
      clean = is_opcode(data[q]) ? 0 : clean + 1;
    }
    if (start > stop) start = stop;
  }
  uint32_t next = Grp::shfl_down(start, 1);
  if (lane == G - 1) next = stop;
  Grp::sync();                                           /* all safe points found on unmodified bytes */
  uint32_t hist = lane == 0 ? (state & 7) : 0;
  uint32_t last = lane == 0 ? 0xFFFFFFFFu : start - 5;   /* lane 0: prevPosT = -1 (Bra86.c:18); others: far */
  uint32_t pos = start;
  const bool busy = start < next;
  if (busy) pos = x86_convert_range(data, start, next, ip, encoding, hist, last);
  Grp::sync();
  /* the scan's final (pos, last, hist) are those of the last lane that owned a range */
  const uint32_t owners = Grp::ballot(busy);
#ifdef LZB_EMU
  const uint32_t top = 31u - (uint32_t)__builtin_clz(owners | 1u);
#else
  const uint32_t top = 31u - (uint32_t)__clz((int)(owners | 1u));
#endif
  pos = Grp::bcast(pos, top);
  last = Grp::bcast(last, top);
  hist = Grp::bcast(hist, top);
  const uint32_t gap = pos - last;
  state = gap > 3 ? 0 : ((hist << (gap - 1)) & 7);      /* Bra86.c:82-83 */
  return pos;
}

/* One opcode candidate at `pos` (Bra86.c:26-80): updates the 3-bit neighbour history `hist` and the previous candidate
 * position `last`, converts the operand when the rules allow it, and returns where the scan goes on (pos + 1, or pos + 5
 * behind a conversion).  Everything it READS is original data: a conversion rewrites pos+1 .. pos+4 and resumes at
 * pos+5, and no later candidate looks further back than its own pos+1.  `store`: this lane writes the result (the
 * lanes of a group run the decision redundantly on the same bytes). */
LZB_FN uint32_t x86_candidate(uint8_t *data, uint32_t pos, uint32_t ip5, bool encoding, uint32_t &hist, uint32_t &last, bool store)
{
  /* kMaskToAllowedStatus / kMaskToBitNumber (Bra86.c:8-9) packed into immediates */
  const uint32_t allow = 0x17;          /* bit h set  <=> history h allows a conversion: {1,1,1,0,1,0,0,0} */
  const uint32_t topbit = 0x6DB488u;    /* 3 bits per h: {0,1,2,2,3,3,3,3} */
  const uint32_t gap = pos - last;
  last = pos;
  if (gap > 3) hist = 0;
  else
  {
    hist = (hist << (gap - 1)) & 7;
    if (hist != 0)
    {
      const uint8_t b = data[pos + 4 - ((topbit >> (3 * hist)) & 7)];
      if (!((allow >> hist) & 1) || is_fill(b)) { hist = ((hist << 1) & 7) | 1; return pos + 1; }
    }
  }
  if (!is_fill(data[pos + 4])) { hist = ((hist << 1) & 7) | 1; return pos + 1; }
  uint32_t v = ((uint32_t)data[pos + 4] << 24) | ((uint32_t)data[pos + 3] << 16) | ((uint32_t)data[pos + 2] << 8) | data[pos + 1];
  const uint32_t here = ip5 + pos;
  uint32_t out;
  for (;;)
  {
    out = encoding ? here + v : v - here;
    if (hist == 0) break;
    const uint32_t sh = ((topbit >> (3 * hist)) & 7) * 8;
    if (!is_fill((uint8_t)(out >> (24 - sh)))) break;
    v = out ^ ((1u << (32 - sh)) - 1);
  }
  if (store)
  {
    data[pos + 4] = (uint8_t)(0u - ((out >> 24) & 1));
    data[pos + 3] = (uint8_t)(out >> 16);
    data[pos + 2] = (uint8_t)(out >> 8);
    data[pos + 1] = (uint8_t)out;
  }
  return pos + 5;
}

/* the 16 bytes at data + at (a 16-byte-aligned ADDRESS) as one vector load when all of them are the unit's, byte by byte
 * at the two ends of the buffer; bytes outside [0, size) read as 0 */
LZB_FN void x86_load16(const uint8_t *data, int64_t at, uint32_t size, uint32_t w[4])
{
  if (at >= 0 && at + 16 <= (int64_t)size)
  {
#ifdef LZB_EMU
    memcpy(w, data + at, 16);
#else
    const uint4 v = *reinterpret_cast<const uint4 *>(data + at);
    w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
#endif
    return;
  }
  w[0] = w[1] = w[2] = w[3] = 0;
  if (at + 16 <= 0 || at >= (int64_t)size) return;
#pragma unroll
  for (int k = 0; k < 4; k++)                            /* w[] stays in registers: static index, rolled inner loop */
  {
    uint32_t acc = 0;
#pragma unroll 1
    for (int j = 0; j < 4; j++)
    {
      const int64_t q = at + 4 * k + j;
      if (q >= 0 && q < (int64_t)size) acc |= (uint32_t)data[q] << (8 * j);
    }
    w[k] = acc;
  }
}

/* bit j set <=> byte j of the 16 is an opcode candidate (E8 / E9) */
LZB_FN uint32_t x86_opcode_mask16(const uint32_t w[4])
{
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 16; j++) m |= (((w[j >> 2] >> (8 * (j & 3))) & 0xFEu) == 0xE8u ? 1u : 0u) << j;
  return m;
}

/* Whole-buffer x86_Convert by one group.  Returns the processed byte count; *state in/out as in the reference.
 *
 * The candidates have to be taken in order (each decision depends on the history the earlier ones left), but they
 * are sparse, and finding them is what the bytes cost: the group sweeps the buffer in tiles of G x 16 bytes, every
 * lane loads one aligned 16-byte vector (a coalesced tile, the next one already in flight) and marks its opcode
 * bytes; the group then walks the marked positions in ascending order -- ballot for the first lane that has one at
 * or behind the scan position, broadcast of its lowest bit -- and every lane runs the (uniform) decision of
 * `x86_candidate` on it, lane 0 storing.  Marks are taken from original bytes and positions a conversion jumps over
 * are simply behind the scan position, so nothing needs to be re-read.  (Before: one contiguous segment per lane
 * scanned byte by byte from "safe points" -- G uncoalesced byte streams, 52 ms for a 4 MiB block, 3.6 % of the
 * `xz4m` launch: profiles/r02_z_x86_tiles.md.) */
#ifdef LZB_EMU
#define LZB_COLD_FN inline
#else
#define LZB_COLD_FN __device__ __noinline__
#endif
template <uint32_t G> LZB_COLD_FN uint32_t x86_convert_tiles(uint8_t *data, uint32_t size, uint32_t ip, bool encoding, uint32_t &state)
{
  typedef Group<G> Grp;
  if (size < 5) return 0;                               /* Bra86.c:15-16 (state untouched) */
  const uint32_t stop = size - 4;                       /* candidates live in [0, stop) */
  const uint32_t lane = Grp::lane();
  Grp::sync();                                          /* the decode's stores of the other lanes are visible */
  uint32_t hist = state & 7;
  uint32_t last = 0xFFFFFFFFu;                          /* prevPosT = -1 (Bra86.c:18) */
  uint32_t pos = 0;
  const uint32_t ip5 = ip + 5;
  const int64_t tile = (int64_t)G * 16;
  int64_t at = -(int64_t)((uintptr_t)data & 15) + (int64_t)lane * 16;    /* this lane's 16 bytes of the current tile */
  uint32_t cur[4], nxt[4];
  x86_load16(data, at, size, cur);
  for (int64_t tb = at - (int64_t)lane * 16; tb < (int64_t)stop; tb += tile, at += tile)
  {
    x86_load16(data, at + tile, size, nxt);
    uint32_t m = x86_opcode_mask16(cur);
    if (at + 16 > (int64_t)stop) m = at >= (int64_t)stop ? 0u : m & ((1u << (uint32_t)((int64_t)stop - at)) - 1u);
    for (;;)
    {
      const int64_t rel = (int64_t)pos - at;             /* marks below the scan position are behind us */
      const uint32_t mm = rel <= 0 ? m : (rel >= 16 ? 0u : m & (0xFFFFu << (uint32_t)rel));
      const uint32_t have = Grp::ballot(mm != 0);
      if (have == 0) break;
#ifdef LZB_EMU
      const uint32_t owner = (uint32_t)__builtin_ctz(have);
      const uint32_t mine = (uint32_t)at + (uint32_t)__builtin_ctz(mm | 0x10000u);
#else
      const uint32_t owner = (uint32_t)__ffs((int)have) - 1u;
      const uint32_t mine = (uint32_t)at + (uint32_t)__ffs((int)(mm | 0x10000u)) - 1u;
#endif
      const uint32_t c = Grp::bcast(mine, owner);
      pos = x86_candidate(data, c, ip5, encoding, hist, last, lane == 0);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) cur[j] = nxt[j];
  }
  Grp::sync();                                          /* lane 0's stores are visible to the check that follows */
  if (pos < stop) pos = stop;
  const uint32_t gap = pos - last;
  state = gap > 3 ? 0 : ((hist << (gap - 1)) & 7);      /* Bra86.c:82-83 */
  return pos;
}

/* ------------------------------------------------------------------ Delta ---------------------- */
template <uint32_t G> LZB_FN void delta_decode_group(uint8_t *data, uint32_t size, uint32_t dist, uint8_t *acc /* dist*G bytes */)
{
  typedef Group<G> Grp;
  const uint32_t lane = Grp::lane();
  const uint32_t rows = (size + dist - 1) / dist;
  const uint32_t seg = ((rows + G - 1) / G) * dist;
  const uint32_t begin = (uint64_t)lane * seg < size ? lane * seg : size;
  const uint32_t end = size - begin < seg ? size : begin + seg;
  Grp::sync();
  for (uint32_t c = 0; c < dist; c++) acc[c * G + lane] = 0;
  for (uint32_t i = begin, c = 0; i < end; i++)
  {
    acc[c * G + lane] = (uint8_t)(acc[c * G + lane] + data[i]);
    if (++c == dist) c = 0;
  }
  Grp::sync();
  for (uint32_t c = lane; c < dist; c += G)              /* exclusive scan across lanes, per column */
  {
    uint8_t run = 0;
    for (uint32_t l = 0; l < G; l++)
    {
      const uint8_t t = acc[c * G + l];
      acc[c * G + l] = run;
      run = (uint8_t)(run + t);
    }
  }
  Grp::sync();
  for (uint32_t i = begin, c = 0; i < end; i++)
  {
    const uint8_t v = (uint8_t)(acc[c * G + lane] + data[i]);
    acc[c * G + lane] = v;
    data[i] = v;
    if (++c == dist) c = 0;
  }
  Grp::sync();
}

/* ------------------------------------------------------------------ CRC ------------------------ */
/* products in GF(2)[x]/P, reflected bit order: the top bit is x^0 */
LZB_FN uint32_t mulmod32(uint32_t a, uint32_t b)
{
  uint32_t p = 0;
#pragma unroll 1                       /* cold (once per unit and lane): keep it small, the kernel it is inlined into is hot */
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
#pragma unroll 1
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
#pragma unroll 1
  for (; n != 0; n >>= 1) { if (n & 1) r = mulmod32(r, sq); sq = mulmod32(sq, sq); }
  return r;
}
LZB_FN uint64_t xpow64(uint32_t n)
{
  uint64_t r = 1ull << 63, sq = 1ull << 55;
#pragma unroll 1
  for (; n != 0; n >>= 1) { if (n & 1) r = mulmod64(r, sq); sq = mulmod64(sq, sq); }
  return r;
}

template <uint32_t G> LZB_FN void lane_segment(uint32_t size, uint32_t &begin, uint32_t &end)
{
  const uint32_t lane = Group<G>::lane();
  uint32_t seg = (size + G - 1) / G;
  seg = (seg + 15u) & ~15u;
  begin = (uint64_t)lane * seg < size ? lane * seg : size;
  end = size - begin < seg ? size : begin + seg;
}

/* Slicing tables in the group's scratch: table k answers "what does this byte contribute once k more bytes have
 * gone through the register" (T0 = the classic byte table; Tk[i] = T0[Tk-1[i] & 0xFF] ^ (Tk-1[i] >> 8); the
 * reference's CrcGenerateTable builds the same family, 7zCrc.c:49-74).  With four tables a lane folds an aligned
 * 32-bit word per step -- one dependent shared-memory round trip per four bytes instead of one per byte. */
template <uint32_t G, class T> LZB_FN void crc_tables(T *table, T poly, uint32_t slices)
{
  typedef Group<G> Grp;
  const uint32_t lane = Grp::lane();
  Grp::sync();
#pragma unroll 1
  for (uint32_t i = lane; i < 256; i += G)
  {
    T r = (T)i;
#pragma unroll 1
    for (int k = 0; k < 8; k++) r = (r >> 1) ^ (poly & (T)(0 - (r & 1)));
    table[i] = r;
  }
  Grp::sync();
  for (uint32_t k = 1; k < slices; k++)
  {
    for (uint32_t i = lane; i < 256; i += G)
    {
      const T r = table[(k - 1) * 256 + i];
      table[k * 256 + i] = table[(uint32_t)(r & 0xFF)] ^ (r >> 8);
    }
    Grp::sync();
  }
}

/* the register of one lane over its segment [begin, end) of data, starting from v */
template <class T> LZB_FN T crc_segment(const uint8_t *data, uint32_t begin, uint32_t end, T v, const T *table, uint32_t slices)
{
  uint32_t i = begin;
  if (slices >= 4)
  {
    while (i < end && (((uintptr_t)(data + i)) & 3u) != 0) { v = table[(uint32_t)((v ^ data[i]) & 0xFF)] ^ (v >> 8); i++; }
    for (; i + 4 <= end; i += 4)
    {
      const uint32_t w = *reinterpret_cast<const uint32_t *>(data + i) ^ (uint32_t)v;
      v = (sizeof(T) == 8 ? (T)((uint64_t)v >> 32) : (T)0) ^
          table[768 + (w & 0xFF)] ^ table[512 + ((w >> 8) & 0xFF)] ^ table[256 + ((w >> 16) & 0xFF)] ^ table[w >> 24];
    }
  }
  for (; i < end; i++) v = table[(uint32_t)((v ^ data[i]) & 0xFF)] ^ (v >> 8);
  return v;
}

/* CRC-32 as CrcCalc (7zCrc.c:44): preset all-ones, final inversion.  scratch_bytes: what `table` may use. */
template <uint32_t G> LZB_FN uint32_t crc32_group(const uint8_t *data, uint32_t size, uint32_t *table, uint32_t scratch_bytes)
{
  typedef Group<G> Grp;
  const uint32_t lane = Grp::lane();
  const uint32_t slices = scratch_bytes >= 4096 ? 4u : 1u;
  crc_tables<G, uint32_t>(table, kPoly32, slices);
  uint32_t begin, end;
  lane_segment<G>(size, begin, end);
  uint32_t v = crc_segment<uint32_t>(data, begin, end, lane == 0 ? 0xFFFFFFFFu : 0u, table, slices);
  v = mulmod32(v, xpow32(size - end));
  for (uint32_t m = G >> 1; m != 0; m >>= 1) v ^= Grp::shfl_xor(v, m);
  Grp::sync();
  return v ^ 0xFFFFFFFFu;
}

/* CRC-64/XZ as Crc64Calc (XzCrc64.c:30) */
template <uint32_t G> LZB_FN uint64_t crc64_group(const uint8_t *data, uint32_t size, uint64_t *table, uint32_t scratch_bytes)
{
  typedef Group<G> Grp;
  const uint32_t lane = Grp::lane();
  const uint32_t slices = scratch_bytes >= 8192 ? 4u : 1u;
  crc_tables<G, uint64_t>(table, kPoly64, slices);
  uint32_t begin, end;
  lane_segment<G>(size, begin, end);
  uint64_t v = crc_segment<uint64_t>(data, begin, end, lane == 0 ? ~0ull : 0ull, table, slices);
  v = mulmod64(v, xpow64(size - end));
  for (uint32_t m = G >> 1; m != 0; m >>= 1) v ^= Grp::shfl_xor(v, m);
  Grp::sync();
  return ~v;
}

/* ------------------------------------------------------------------ RISC branch converters ----- */
/* ARM, ARMT, PPC, SPARC (Bra.c:6-133) and IA64 (BraIA64.c:16-69) rewrite one aligned instruction slot at a
 * time and carry no state, so every slot is an independent job: lanes take slots round-robin.  (ARMT looks
 * sequential -- a converted pair skips the next halfword -- but a BL pair at i and one at i+2 would need
 * byte i+3 to be both 11110xxx and 11111xxx, and conversion keeps those bits: candidates never overlap.)
 * Each returns what the reference's loop counter returns: the bytes it is done with. */
LZB_FN uint32_t bra_rel(uint32_t v, uint32_t pc, bool encoding) { return encoding ? pc + v : v - pc; }

template <uint32_t G> LZB_FN uint32_t risc_convert_group(uint8_t *d, uint32_t size, uint32_t filter, uint32_t ip, bool enc)
{
  typedef Group<G> Grp;
  const uint32_t lane = Grp::lane();
  Grp::sync();
  if (filter == LZB_FILTER_IA64)
  {
    if (size < 16) return 0;
    const uint32_t n = size >> 4;
    for (uint32_t b = lane; b < n; b += G)
    {
      uint8_t *q = d + ((size_t)b << 4);
      const uint32_t tmpl = q[0] & 0x1F;
      /* kBranchTable (BraIA64.c:6-12): templates 16,17,24,25,28,29 -> slot 2; 18,19 -> slots 1,2; 22,23 -> all */
      const uint32_t mask = (tmpl == 16 || tmpl == 17 || tmpl == 24 || tmpl == 25 || tmpl == 28 || tmpl == 29) ? 4u
                          : ((tmpl == 18 || tmpl == 19) ? 6u : ((tmpl == 22 || tmpl == 23) ? 7u : 0u));
      for (uint32_t slot = 0, bit = 5; slot < 3; slot++, bit += 41)
      {
        if (((mask >> slot) & 1) == 0) continue;
        const uint32_t at = bit >> 3, sh = bit & 7;
        uint64_t ins = 0;
        for (int j = 0; j < 6; j++) ins += (uint64_t)q[at + j] << (8 * j);
        uint64_t norm = ins >> sh;
        if (((norm >> 37) & 0xF) != 0x5 || ((norm >> 9) & 0x7) != 0) continue;
        uint32_t v = (uint32_t)((norm >> 13) & 0xFFFFF) | (((uint32_t)(norm >> 36) & 1) << 20);
        v = bra_rel(v << 4, ip + (b << 4), enc) >> 4;
        norm &= ~((uint64_t)0x8FFFFF << 13);
        norm |= (uint64_t)(v & 0xFFFFF) << 13;
        norm |= (uint64_t)(v & 0x100000) << (36 - 20);
        ins = (ins & ((1u << sh) - 1)) | (norm << sh);
        for (int j = 0; j < 6; j++) q[at + j] = (uint8_t)(ins >> (8 * j));
      }
    }
    Grp::sync();
    return n << 4;
  }
  if (size < 4) return 0;
  if (filter == LZB_FILTER_ARMT)
  {
    const uint32_t last = (size - 4) & ~1u;               /* last halfword position the loop can look at */
    bool hit_last = false;
    for (uint32_t i = lane * 2; i <= last; i += G * 2)
    {
      uint8_t *q = d + i;
      if ((q[1] & 0xF8) != 0xF0 || (q[3] & 0xF8) != 0xF8) continue;
      uint32_t v = (((uint32_t)q[1] & 7) << 19) | ((uint32_t)q[0] << 11) | (((uint32_t)q[3] & 7) << 8) | q[2];
      v = bra_rel(v << 1, ip + 4 + i, enc) >> 1;
      q[1] = (uint8_t)(0xF0 | ((v >> 19) & 7)); q[0] = (uint8_t)(v >> 11);
      q[3] = (uint8_t)(0xF8 | ((v >> 8) & 7));  q[2] = (uint8_t)v;
      if (i == last) hit_last = true;
    }
    const bool any = Grp::ballot(hit_last) != 0;
    Grp::sync();
    return last + (any ? 4u : 2u);
  }
  const uint32_t n = size >> 2;
  for (uint32_t w = lane; w < n; w += G)
  {
    uint8_t *q = d + ((size_t)w << 2);
    const uint32_t i = w << 2;
    if (filter == LZB_FILTER_ARM)
    {
      if (q[3] != 0xEB) continue;
      uint32_t v = ((uint32_t)q[2] << 16) | ((uint32_t)q[1] << 8) | q[0];
      v = bra_rel(v << 2, ip + 8 + i, enc) >> 2;
      q[2] = (uint8_t)(v >> 16); q[1] = (uint8_t)(v >> 8); q[0] = (uint8_t)v;
    }
    else if (filter == LZB_FILTER_PPC)
    {
      if ((q[0] >> 2) != 0x12 || (q[3] & 3) != 1) continue;
      uint32_t v = ((uint32_t)(q[0] & 3) << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (q[3] & ~3u);
      v = bra_rel(v, ip + i, enc);
      q[0] = (uint8_t)(0x48 | ((v >> 24) & 3)); q[1] = (uint8_t)(v >> 16); q[2] = (uint8_t)(v >> 8);
      q[3] = (uint8_t)((q[3] & 3) | v);
    }
    else /* SPARC */
    {
      if (!((q[0] == 0x40 && (q[1] & 0xC0) == 0x00) || (q[0] == 0x7F && (q[1] & 0xC0) == 0xC0))) continue;
      uint32_t v = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
      v = bra_rel(v << 2, ip + i, enc) >> 2;
      v = (((0u - ((v >> 22) & 1)) << 22) & 0x3FFFFFFFu) | (v & 0x3FFFFFu) | 0x40000000u;
      q[0] = (uint8_t)(v >> 24); q[1] = (uint8_t)(v >> 16); q[2] = (uint8_t)(v >> 8); q[3] = (uint8_t)v;
    }
  }
  Grp::sync();
  return n << 2;
}

/* Delta_Encode (Delta.c:20-41), zero incoming history: out[i] = in[i] - in[i-d].  In place: chunks of G bytes
 * from the top down, every chunk read completely before it is written (out[i] needs in[i-d], below it). */
template <uint32_t G> LZB_FN void delta_encode_group(uint8_t *data, uint32_t n, uint32_t d)
{
  typedef Group<G> Grp;
  uint32_t top = n;
  while (top > 0)
  {
    const uint32_t base = top > G ? top - G : 0;
    const uint32_t i = base + Grp::lane();
    const bool on = i < top;
    uint8_t v = 0;
    if (on)
    {
      v = data[i];
      if (i >= d) v = (uint8_t)(v - data[i - d]);
    }
    Grp::sync();
    if (on) data[i] = v;
    Grp::sync();
    top = base;
  }
}

/* filter -> check -> verdict, for one decoded unit.
 * `scratch`     : >= kPostSmemBytes of the group's (dead) shared-memory model region -- CRC table.
 * `big_scratch` : >= kPostScratchBytes, the group's (dead) matched-literal slab slice in HBM/L2 -- Delta
 *                 column sums (up to 256 x G bytes; Delta blocks are rare, L2 latency is acceptable). */
/* kRisc: also the stateless RISC converters (only the stand-alone / second-pass kernel carries them, the
 * fused decode kernel stays lean: BCJ x86 and Delta are what the reference's own writers emit). */
/* kTiles: x86 BCJ as the tile scan (look-ahead decode kernels, the stand-alone / second-pass kernel) or as per-lane
 * segments (lean decode kernels, see x86_convert_lanes) -- same results, tests run both. */
template <uint32_t G, bool kRisc = false, bool kTiles = true>
LZB_FN void post_process(const LzbUnit &u, LzbResult &r, uint8_t *data, uint8_t *scratch, uint8_t *big_scratch,
                         uint32_t scratch_bytes = kPostSmemBytes)
{
  if (r.res != LZB_SZ_OK) return;
  const uint32_t n = r.dst_done;
  if (u.filter == LZB_FILTER_X86)
  {
    /* stand-alone x86_Convert calls pass the direction in flags bit 1 and the history in `expect`; a filter fused
     * behind a decode starts a block / folder with no history (`expect` is the stored digest there) */
    uint32_t st = (u.flags & LZB_UNIT_POST_REPORT) ? ((uint32_t)(u.expect >> 32) & 7) : 0u;
    const uint32_t done = kTiles ? x86_convert_tiles<G>(data, n, u.filter_param, (u.flags & 2) != 0, st)
                                 : x86_convert_lanes<G>(data, n, u.filter_param, (u.flags & 2) != 0, st);
    if (u.flags & 4) { r.src_used = done; r.status = (int32_t)st; }   /* stand-alone call: report them */
  }
  else if (u.filter == LZB_FILTER_DELTA)
  {
    if (u.flags & 2) delta_encode_group<G>(data, n, u.filter_param);   /* stand-alone Delta_Encode call */
    else delta_decode_group<G>(data, n, u.filter_param, big_scratch);
  }
  else if (kRisc && u.filter >= LZB_FILTER_PPC && u.filter <= LZB_FILTER_SPARC)
  {
    const uint32_t done = risc_convert_group<G>(data, n, u.filter, u.filter_param, (u.flags & 2) != 0);
    if (u.flags & 4) r.src_used = done;
  }
  if (u.check == LZB_CHECK_CRC32) r.check = crc32_group<G>(data, n, reinterpret_cast<uint32_t *>(scratch), scratch_bytes);
  else if (u.check == LZB_CHECK_CRC64) r.check = crc64_group<G>(data, n, reinterpret_cast<uint64_t *>(scratch), scratch_bytes);
  if ((u.flags & LZB_UNIT_VERIFY) && u.check != LZB_CHECK_NONE && r.check != u.expect) r.res = LZB_SZ_ERROR_CRC;
}

}  // namespace lzb
