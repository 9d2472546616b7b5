/* lzma2_oracle.c -- CPU restatement of the reference LZMA2 chunk layer.  TEST INFRASTRUCTURE ONLY.
 *
 * Chunk grammar (Lzma2Dec.c:14-26):
 *   00                     end of stream
 *   01 UU UU               stored chunk, dictionary reset
 *   02 UU UU               stored chunk
 *   1mmuuuuu UU UU PP PP [props]   LZMA chunk; mm: 0 none, 1 state reset, 2 +props, 3 +dict reset
 * unpack = ((c & 0x1F) << 16 | UU UU) + 1, pack = PP PP + 1.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

enum { PH_CONTROL, PH_UNPACK_HI, PH_UNPACK_LO, PH_PACK_HI, PH_PACK_LO, PH_PROPS, PH_DATA, PH_DATA_MORE,
       PH_FINISHED, PH_ERROR };

#define IS_STORED(c) (((c) & 0x80) == 0)
#define LZMA_MODE(c) (((c) >> 5) & 3)

int ora_lzma2_dict_from_prop(uint8_t prop, uint32_t *dict_size)
{
  if (prop > 40) return ORA_ERROR_UNSUPPORTED;
  *dict_size = prop == 40 ? 0xFFFFFFFFu : ((uint32_t)(2 | (prop & 1)) << (prop / 2 + 11));
  return ORA_OK;
}

/* Lzma2Dec.c:61-87: LZMA2 always sizes the model for lc + lp = 4 */
static void legacy_props(uint32_t dict_size, uint8_t *p5)
{
  p5[0] = 4;
  p5[1] = (uint8_t)dict_size; p5[2] = (uint8_t)(dict_size >> 8);
  p5[3] = (uint8_t)(dict_size >> 16); p5[4] = (uint8_t)(dict_size >> 24);
}

int ora_lzma2_alloc_probs(OraLzma2 *d, uint8_t prop)
{
  uint8_t p5[5];
  uint32_t ds;
  int r = ora_lzma2_dict_from_prop(prop, &ds);
  if (r != ORA_OK) return r;
  legacy_props(ds, p5);
  return ora_lzma_alloc_probs(&d->lz, p5, 5);
}

int ora_lzma2_alloc(OraLzma2 *d, uint8_t prop)
{
  uint8_t p5[5];
  uint32_t ds;
  int r = ora_lzma2_dict_from_prop(prop, &ds);
  if (r != ORA_OK) return r;
  legacy_props(ds, p5);
  return ora_lzma_alloc(&d->lz, p5, 5);
}

void ora_lzma2_free(OraLzma2 *d) { ora_lzma_free(&d->lz); }

void ora_lzma2_reset(OraLzma2 *d)
{
  d->phase = PH_CONTROL;
  d->want_dic_reset = d->want_state_reset = d->want_props = 1;
  ora_lzma_reset(&d->lz);
}

/* Lzma2Dec.c:98-157: one header byte */
static int header_byte(OraLzma2 *d, uint8_t b)
{
  switch (d->phase)
  {
    case PH_CONTROL:
      d->control = b;
      if (b == 0) return PH_FINISHED;
      if (IS_STORED(b))
      {
        if ((b & 0x7F) > 2) return PH_ERROR;
        d->unpack_left = 0;
      }
      else
        d->unpack_left = (uint32_t)(b & 0x1F) << 16;
      return PH_UNPACK_HI;
    case PH_UNPACK_HI:
      d->unpack_left |= (uint32_t)b << 8;
      return PH_UNPACK_LO;
    case PH_UNPACK_LO:
      d->unpack_left |= b;
      d->unpack_left++;
      return IS_STORED(d->control) ? PH_DATA : PH_PACK_HI;
    case PH_PACK_HI:
      d->pack_left = (uint32_t)b << 8;
      return PH_PACK_LO;
    case PH_PACK_LO:
      d->pack_left |= b;
      d->pack_left++;
      if (LZMA_MODE(d->control) >= 2) return PH_PROPS;
      return d->want_props ? PH_ERROR : PH_DATA;
    case PH_PROPS:
    {
      unsigned lc, lp;
      if (b >= 9 * 5 * 5) return PH_ERROR;
      lc = b % 9; b /= 9;
      lp = b % 5;
      if (lc + lp > 4) return PH_ERROR;
      d->lz.pb = b / 5; d->lz.lc = lc; d->lz.lp = lp;
      d->want_props = 0;
      return PH_DATA;
    }
    default:
      return PH_ERROR;
  }
}

int ora_lzma2_to_dic(OraLzma2 *d, size_t dic_limit, const uint8_t *src, size_t *src_len, int finish, int *status)
{
  size_t in_size = *src_len;
  *src_len = 0;
  *status = ORA_ST_NOT_SPECIFIED;

  while (d->phase != PH_FINISHED)
  {
    size_t at = d->lz.dic_pos;
    if (d->phase == PH_ERROR) return ORA_ERROR_DATA;
    if (at == dic_limit && finish == ORA_FINISH_ANY) { *status = ORA_ST_NOT_FINISHED; return ORA_OK; }

    if (d->phase != PH_DATA && d->phase != PH_DATA_MORE)
    {
      if (*src_len == in_size) { *status = ORA_ST_NEEDS_MORE_INPUT; return ORA_OK; }
      (*src_len)++;
      d->phase = header_byte(d, *src++);
      continue;
    }

    {
      size_t out_now = dic_limit - at;
      size_t in_now = in_size - *src_len;
      int mode = ORA_FINISH_ANY;
      if (d->unpack_left <= out_now) { out_now = d->unpack_left; mode = ORA_FINISH_END; }

      if (IS_STORED(d->control))                                  /* Lzma2Dec.c:209-240 */
      {
        if (*src_len == in_size) { *status = ORA_ST_NEEDS_MORE_INPUT; return ORA_OK; }
        if (d->phase == PH_DATA)
        {
          int reset_dic = d->control == 1;
          if (reset_dic) d->want_props = d->want_state_reset = 1;
          else if (d->want_dic_reset) return ORA_ERROR_DATA;
          d->want_dic_reset = 0;
          ora_lzma_reset_dic_state(&d->lz, reset_dic, 0);
        }
        if (in_now > out_now) in_now = out_now;
        if (in_now == 0) return ORA_ERROR_DATA;
        /* Lzma2Dec.c:159-166 */
        memcpy(d->lz.dic + d->lz.dic_pos, src, in_now);
        d->lz.dic_pos += in_now;
        if (d->lz.dict_full == 0 && d->lz.dict_size - d->lz.total_pos <= in_now) d->lz.dict_full = d->lz.dict_size;
        d->lz.total_pos += (uint32_t)in_now;
        src += in_now; *src_len += in_now;
        d->unpack_left -= (uint32_t)in_now;
        d->phase = d->unpack_left == 0 ? PH_CONTROL : PH_DATA_MORE;
      }
      else                                                        /* Lzma2Dec.c:241-284 */
      {
        size_t made;
        int res;
        if (d->phase == PH_DATA)
        {
          int m = LZMA_MODE(d->control);
          int reset_dic = m == 3, reset_state = m > 0;
          if ((!reset_dic && d->want_dic_reset) || (!reset_state && d->want_state_reset)) return ORA_ERROR_DATA;
          ora_lzma_reset_dic_state(&d->lz, reset_dic, reset_state);
          d->want_dic_reset = d->want_state_reset = 0;
          d->phase = PH_DATA_MORE;
        }
        if (in_now > d->pack_left) in_now = d->pack_left;
        res = ora_lzma_to_dic(&d->lz, at + out_now, src, &in_now, mode, status);
        src += in_now; *src_len += in_now;
        d->pack_left -= (uint32_t)in_now;
        made = d->lz.dic_pos - at;
        d->unpack_left -= (uint32_t)made;
        if (res != ORA_OK) return res;
        if (*status == ORA_ST_NEEDS_MORE_INPUT) return res;
        if (in_now == 0 && made == 0)
        {
          if (*status != ORA_ST_MAYBE_FINISHED_WITHOUT_MARK || d->unpack_left != 0 || d->pack_left != 0)
            return ORA_ERROR_DATA;
          d->phase = PH_CONTROL;
        }
        if (*status == ORA_ST_MAYBE_FINISHED_WITHOUT_MARK) *status = ORA_ST_NOT_FINISHED;
      }
    }
  }
  *status = ORA_ST_FINISHED_WITH_MARK;
  return ORA_OK;
}

int ora_lzma2_to_buf(OraLzma2 *d, uint8_t *dest, size_t *dest_len, const uint8_t *src, size_t *src_len,
                     int finish, int *status)
{
  size_t out_room = *dest_len, in_left = *src_len;
  *src_len = *dest_len = 0;
  for (;;)
  {
    size_t in_now = in_left, start, stop, made;
    int mode, res;
    if (d->lz.dic_pos == d->lz.dic_size) d->lz.dic_pos = 0;
    start = d->lz.dic_pos;
    if (out_room > d->lz.dic_size - start) { stop = d->lz.dic_size; mode = ORA_FINISH_ANY; }
    else { stop = start + out_room; mode = finish; }
    res = ora_lzma2_to_dic(d, stop, src, &in_now, mode, status);
    src += in_now; in_left -= in_now; *src_len += in_now;
    made = d->lz.dic_pos - start;
    memcpy(dest, d->lz.dic + start, made);
    dest += made; out_room -= made; *dest_len += made;
    if (res != ORA_OK) return res;
    if (made == 0 || out_room == 0) return ORA_OK;
  }
}

int ora_lzma2_decode_unit(uint8_t *dest, size_t dest_cap, size_t *dest_len, const uint8_t *src, size_t *src_len,
                          uint8_t prop, int finish, int *status)
{
  OraLzma2 d;
  int res;
  memset(&d, 0, sizeof d);
  *status = ORA_ST_NOT_SPECIFIED;
  res = ora_lzma2_alloc_probs(&d, prop);
  if (res != ORA_OK) { *dest_len = 0; *src_len = 0; return res; }
  d.lz.dic = dest;
  d.lz.dic_size = dest_cap;
  ora_lzma2_reset(&d);
  res = ora_lzma2_to_dic(&d, dest_cap, src, src_len, finish, status);
  *dest_len = d.lz.dic_pos;
  free(d.lz.probs);
  return res;
}

long ora_lzma2_index(const uint8_t *src, size_t src_len, OraUnit *units, size_t max_units,
                     size_t *end_off, int *has_eos)
{
  size_t p = 0;
  long n = 0;
  *has_eos = 0;
  while (p < src_len)
  {
    uint8_t c = src[p];
    size_t hdr, pack, unpack;
    unsigned lclp = 0;
    if (c == 0) { p++; *has_eos = 1; break; }
    if (IS_STORED(c))
    {
      if (c > 2) { *end_off = p; return -1; }
      hdr = 3;
      if (p + hdr > src_len) break;
      unpack = (((size_t)src[p + 1] << 8) | src[p + 2]) + 1;
      pack = unpack;
    }
    else
    {
      hdr = 5 + (LZMA_MODE(c) >= 2 ? 1 : 0);
      if (p + hdr > src_len) break;
      unpack = (((size_t)(c & 0x1F) << 16) | ((size_t)src[p + 1] << 8) | src[p + 2]) + 1;
      pack = (((size_t)src[p + 3] << 8) | src[p + 4]) + 1;
      if (hdr == 6)
      {
        unsigned b = src[p + 5];
        if (b >= 225 || (b % 9) + ((b / 9) % 5) > 4) { *end_off = p; return -1; }
        lclp = (b % 9) + ((b / 9) % 5);
      }
    }
    if (c == 1 || c >= 0xE0 || n == 0)
    {
      n++;
      if ((size_t)n <= max_units)
      {
        OraUnit *u = &units[n - 1];
        u->src_off = p; u->src_len = 0; u->unpack_len = 0; u->n_chunks = 0; u->max_lclp = 0;
      }
    }
    if ((size_t)n <= max_units)
    {
      OraUnit *u = &units[n - 1];
      u->src_len += hdr + pack; u->unpack_len += unpack; u->n_chunks++;
      if (lclp > u->max_lclp) u->max_lclp = lclp;
    }
    p += hdr + pack;
  }
  *end_off = p < src_len ? p : src_len;
  return n;
}
