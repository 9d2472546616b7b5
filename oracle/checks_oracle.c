/* checks_oracle.c -- CPU restatement of CRC-32 (7zCrc.c:49-74, 7zCrcOpt.c:10-27) and
 * CRC-64/XZ (XzCrc64.c:9-33).  TEST INFRASTRUCTURE ONLY (see oracle.h).
 * Both are reflected CRCs with all-ones preset and final inversion; the reference's
 * slicing-by-4 is an evaluation order only, the value equals the bytewise definition.
 */
#include "oracle.h"

#define POLY32 0xEDB88320u
#define POLY64 0xC96C5795D7870F42ull

static uint32_t t32[4][256];
static uint64_t t64[256];
static volatile int ready;

static void make_tables(void)
{
  unsigned i, k;
  if (ready) return;
  for (i = 0; i < 256; i++)
  {
    uint32_t r = i;
    uint64_t q = i;
    for (k = 0; k < 8; k++)
    {
      r = (r >> 1) ^ (POLY32 & (0u - (r & 1)));
      q = (q >> 1) ^ (POLY64 & (0ull - (q & 1)));
    }
    t32[0][i] = r;
    t64[i] = q;
  }
  for (k = 1; k < 4; k++)
    for (i = 0; i < 256; i++)
      t32[k][i] = t32[0][t32[k - 1][i] & 0xFF] ^ (t32[k - 1][i] >> 8);
  ready = 1;
}

uint32_t ora_crc32_update(uint32_t v, const void *data, size_t size)
{
  const uint8_t *p = (const uint8_t *)data;
  make_tables();
  for (; size > 0 && ((uintptr_t)p & 3) != 0; size--, p++) v = t32[0][(v ^ *p) & 0xFF] ^ (v >> 8);
  for (; size >= 4; size -= 4, p += 4)
  {
    v ^= (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    v = t32[3][v & 0xFF] ^ t32[2][(v >> 8) & 0xFF] ^ t32[1][(v >> 16) & 0xFF] ^ t32[0][v >> 24];
  }
  for (; size > 0; size--, p++) v = t32[0][(v ^ *p) & 0xFF] ^ (v >> 8);
  return v;
}

uint32_t ora_crc32(const void *data, size_t size)
{
  return ora_crc32_update(0xFFFFFFFFu, data, size) ^ 0xFFFFFFFFu;
}

uint64_t ora_crc64_update(uint64_t v, const void *data, size_t size)
{
  const uint8_t *p = (const uint8_t *)data;
  make_tables();
  for (; size > 0; size--, p++) v = t64[(v ^ *p) & 0xFF] ^ (v >> 8);
  return v;
}

uint64_t ora_crc64(const void *data, size_t size)
{
  return ora_crc64_update(~0ull, data, size) ^ ~0ull;
}
