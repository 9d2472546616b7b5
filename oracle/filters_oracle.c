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
