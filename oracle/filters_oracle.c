/* filters_oracle.c -- CPU restatement of the x86 BCJ and Delta post-filters.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 */
#include "oracle.h"

#include <string.h>

/* Bra86.c:6-9.  `hist` is a 3-bit record of which of the previous three byte positions
 * held an E8/E9 opcode candidate that was NOT converted; such neighbours veto or adjust
 * a conversion so that encode and decode stay inverse to each other. */
static const uint8_t k_hist_allows[8]  = { 1, 1, 1, 0, 1, 0, 0, 0 };
static const uint8_t k_hist_top_bit[8] = { 0, 1, 2, 2, 3, 3, 3, 3 };

static int is_sign_fill(uint8_t b) { return b == 0x00 || b == 0xFF; }

/* Bra86.c:11-85 */
size_t ora_x86_convert(uint8_t *data, size_t size, uint32_t ip, uint32_t *state, int encoding)
{
  size_t pos = 0, last_hit = (size_t)0 - 1, stop, gap;
  uint32_t hist = *state & 7;
  if (size < 5) return 0;
  stop = size - 4;
  ip += 5;
  for (;;)
  {
    while (pos < stop && (data[pos] & 0xFE) != 0xE8) pos++;
    if (pos >= stop) break;
    gap = pos - last_hit;
    last_hit = pos;
    if (gap > 3)
      hist = 0;
    else
    {
      hist = (hist << (gap - 1)) & 7;
      if (hist != 0)
      {
        uint8_t b = data[pos + 4 - k_hist_top_bit[hist]];
        if (!k_hist_allows[hist] || is_sign_fill(b))
        {
          hist = ((hist << 1) & 7) | 1;
          pos++;
          continue;
        }
      }
    }
    if (is_sign_fill(data[pos + 4]))
    {
      uint32_t v = ((uint32_t)data[pos + 4] << 24) | ((uint32_t)data[pos + 3] << 16) |
                   ((uint32_t)data[pos + 2] << 8) | data[pos + 1];
      uint32_t here = ip + (uint32_t)pos, out;
      for (;;)
      {
        unsigned sh;
        out = encoding ? here + v : v - here;
        if (hist == 0) break;
        sh = k_hist_top_bit[hist] * 8u;
        if (!is_sign_fill((uint8_t)(out >> (24 - sh)))) break;
        v = out ^ ((1u << (32 - sh)) - 1);
      }
      data[pos + 4] = (uint8_t)(0u - ((out >> 24) & 1));
      data[pos + 3] = (uint8_t)(out >> 16);
      data[pos + 2] = (uint8_t)(out >> 8);
      data[pos + 1] = (uint8_t)out;
      pos += 5;
    }
    else
    {
      hist = ((hist << 1) & 7) | 1;
      pos++;
    }
  }
  gap = pos - last_hit;
  *state = gap > 3 ? 0 : ((hist << (gap - 1)) & 7);
  return pos;
}

/* Delta.c:6-11 */
void ora_delta_init(uint8_t *state256) { memset(state256, 0, 256); }

/* Delta.c:43-62: out[i] = in[i] + out[i - delta]; the 256-byte state is the last `delta`
 * outputs, oldest first. */
void ora_delta_decode(uint8_t *state256, unsigned delta, uint8_t *data, size_t size)
{
  uint8_t ring[256];
  unsigned k = 0;
  size_t i;
  memcpy(ring, state256, delta);
  for (i = 0; i < size; i++)
  {
    ring[k] = data[i] = (uint8_t)(ring[k] + data[i]);
    if (++k == delta) k = 0;
  }
  memcpy(state256, ring + k, delta - k);
  memcpy(state256 + delta - k, ring, k);
}

/* Delta.c:21-41 */
void ora_delta_encode(uint8_t *state256, unsigned delta, uint8_t *data, size_t size)
{
  uint8_t ring[256];
  unsigned k = 0;
  size_t i;
  memcpy(ring, state256, delta);
  for (i = 0; i < size; i++)
  {
    uint8_t b = data[i];
    data[i] = (uint8_t)(b - ring[k]);
    ring[k] = b;
    if (++k == delta) k = 0;
  }
  memcpy(state256, ring + k, delta - k);
  memcpy(state256 + delta - k, ring, k);
}

/* ---- stateless RISC branch converters: Bra.c:6-133 (ARM, ARMT, PPC, SPARC), BraIA64.c:6-69 (IA64) ----
 * id = xz filter id (Xz.h:14-18): 5 PPC, 6 IA64, 7 ARM, 8 ARMT, 9 SPARC.  Each walks aligned instruction
 * slots, rewrites the displacement field of the branch-and-link form between "relative" and "absolute"
 * (encoding: absolute = pc + rel; decoding: rel = absolute - pc) and returns the bytes it is done with. */
static uint32_t ora_rel(uint32_t v, uint32_t pc, int encoding) { return encoding ? pc + v : v - pc; }

static size_t ora_ia64(uint8_t *d, size_t size, uint32_t ip, int enc)
{
  /* which of the three 41-bit slots of a bundle can hold a branch, by template number (BraIA64.c:6-12) */
  size_t base;
  if (size < 16) return 0;
  for (base = 0; base + 16 <= size; base += 16)
  {
    unsigned t = d[base] & 0x1F, slots, s;
    switch (t)
    {
      case 16: case 17: case 24: case 25: case 28: case 29: slots = 4; break;
      case 18: case 19: slots = 6; break;
      case 22: case 23: slots = 7; break;
      default: slots = 0;
    }
    for (s = 0; s < 3; s++)
    {
      unsigned bit = 5 + 41 * s, byte = bit >> 3, sh = bit & 7, j;
      uint64_t raw = 0, ins;
      uint32_t v;
      if (!((slots >> s) & 1)) continue;
      for (j = 0; j < 6; j++) raw |= (uint64_t)d[base + byte + j] << (8 * j);
      ins = raw >> sh;
      if (((ins >> 37) & 0xF) != 5 || ((ins >> 9) & 7) != 0) continue;      /* br.call-type opcode, btype 0 */
      v = (uint32_t)((ins >> 13) & 0xFFFFF) | ((uint32_t)((ins >> 36) & 1) << 20);
      v = ora_rel(v << 4, ip + (uint32_t)base, enc) >> 4;
      ins &= ~((uint64_t)0x8FFFFF << 13);
      ins |= (uint64_t)(v & 0xFFFFF) << 13;
      ins |= (uint64_t)(v & 0x100000) << 16;
      raw = (raw & (((uint64_t)1 << sh) - 1)) | (ins << sh);
      for (j = 0; j < 6; j++) d[base + byte + j] = (uint8_t)(raw >> (8 * j));
    }
  }
  return base;
}

size_t ora_bra_convert(int id, uint8_t *d, size_t size, uint32_t ip, int enc)
{
  size_t i;
  if (id == 6) return ora_ia64(d, size, ip, enc);
  if (size < 4) return 0;
  if (id == 8)
  {
    /* Thumb BL: two halfwords 11110 imm11(hi) | 11111 imm11(lo); a converted pair is stepped over whole */
    i = 0;
    while (i + 4 <= size)
    {
      if ((d[i + 1] & 0xF8) == 0xF0 && (d[i + 3] & 0xF8) == 0xF8)
      {
        uint32_t v = ((uint32_t)(d[i + 1] & 7) << 19) | ((uint32_t)d[i] << 11) | ((uint32_t)(d[i + 3] & 7) << 8) | d[i + 2];
        v = ora_rel(v << 1, ip + 4 + (uint32_t)i, enc) >> 1;
        d[i + 1] = (uint8_t)(0xF0 | ((v >> 19) & 7)); d[i] = (uint8_t)(v >> 11);
        d[i + 3] = (uint8_t)(0xF8 | ((v >> 8) & 7));  d[i + 2] = (uint8_t)v;
        i += 4;
      }
      else i += 2;
    }
    return i;
  }
  for (i = 0; i + 4 <= size; i += 4)
  {
    uint8_t *q = d + i;
    uint32_t v;
    if (id == 7)                                   /* ARM BL: cond=AL, 24-bit word offset, little endian */
    {
      if (q[3] != 0xEB) continue;
      v = ((uint32_t)q[2] << 16) | ((uint32_t)q[1] << 8) | q[0];
      v = ora_rel(v << 2, ip + 8 + (uint32_t)i, enc) >> 2;
      q[2] = (uint8_t)(v >> 16); q[1] = (uint8_t)(v >> 8); q[0] = (uint8_t)v;
    }
    else if (id == 5)                              /* PowerPC bl: opcode 18, AA=0, LK=1, big endian */
    {
      if ((q[0] >> 2) != 0x12 || (q[3] & 3) != 1) continue;
      v = ((uint32_t)(q[0] & 3) << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (q[3] & 0xFCu);
      v = ora_rel(v, ip + (uint32_t)i, enc);
      q[0] = (uint8_t)(0x48 | ((v >> 24) & 3)); q[1] = (uint8_t)(v >> 16); q[2] = (uint8_t)(v >> 8);
      q[3] = (uint8_t)((q[3] & 3) | (v & 0xFF));
    }
    else if (id == 9)                              /* SPARC call: 01 + disp30, only small +/- displacements */
    {
      if (!((q[0] == 0x40 && (q[1] & 0xC0) == 0) || (q[0] == 0x7F && (q[1] & 0xC0) == 0xC0))) continue;
      v = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
      v = ora_rel(v << 2, ip + (uint32_t)i, enc) >> 2;
      v = (((0u - ((v >> 22) & 1)) << 22) & 0x3FFFFFFFu) | (v & 0x3FFFFFu) | 0x40000000u;
      q[0] = (uint8_t)(v >> 24); q[1] = (uint8_t)(v >> 16); q[2] = (uint8_t)(v >> 8); q[3] = (uint8_t)v;
    }
    else return 0;
  }
  return i;
}
