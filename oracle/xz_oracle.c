/* xz_oracle.c -- CPU restatement of the reference's xz container handling for WHOLE buffers.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 *  ora_xz_index  : backward reader  (XzIn.c:71-114 index, :141-224 footer/locate, XzDec.c:505-556
 *                  block header) -- what the host block splitter of the product must reproduce.
 *  ora_xz_decode : forward decoder with the XzUnpacker_Code contract (XzDec.c:604-870) for an input
 *                  that is completely present and a destination that is large enough.
 *
 * Deliberate simplifications, stated so tests know what is pinned:
 *  - the unpacker proves "index == blocks seen" by comparing SHA-256 digests of the record bytes
 *    (XzDec.c:642-645,778,799-801); comparing the bytes themselves is the same predicate.
 *  - check type SHA-256 (10) is reported as ORA_ERROR_UNSUPPORTED (outside the hot-path scope).
 *  - on an error inside a block only `res` is pinned, not the partial lengths.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

static const uint8_t k_magic[6] = { 0xFD, '7', 'z', 'X', 'Z', 0 };

static uint32_t le32(const uint8_t *p) { return p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

/* XzDec.c:30-44 */
static unsigned read_varint(const uint8_t *p, size_t max, uint64_t *v)
{
  unsigned i, lim = max > 9 ? 9 : (unsigned)max;
  *v = 0;
  for (i = 0; i < lim;)
  {
    uint8_t b = p[i];
    *v |= (uint64_t)(b & 0x7F) << (7 * i++);
    if ((b & 0x80) == 0) return (b == 0 && i != 1) ? 0 : i;
  }
  return 0;
}

/* Xz.c:12-24 */
static unsigned write_varint(uint8_t *buf, uint64_t v)
{
  unsigned i = 0;
  do { buf[i++] = (uint8_t)((v & 0x7F) | 0x80); v >>= 7; } while (v != 0);
  buf[i - 1] &= 0x7F;
  return i;
}

/* Xz.c:39-43 */
static unsigned check_size(unsigned type) { return type == 0 ? 0 : (4u << ((type - 1) / 3)); }

/* XzDec.c:482-489: 12-byte stream header (magic already compared by the caller) */
static int parse_stream_header(const uint8_t *h, unsigned *flags)
{
  *flags = ((unsigned)h[6] << 8) | h[7];
  if (ora_crc32(h + 6, 2) != le32(h + 8)) return ORA_ERROR_NO_ARCHIVE;
  return *flags <= 0xF ? ORA_OK : ORA_ERROR_UNSUPPORTED;
}

typedef struct { uint64_t id; uint32_t props_size; uint8_t props[20]; } Filter;
typedef struct { int n; Filter f[4]; } FilterChain;

/* XzDec.c:505-556; `h` holds (h[0]*4 + 4) bytes */
static int parse_block_header(const uint8_t *h, FilterChain *fc)
{
  uint32_t hs = (uint32_t)h[0] << 2;
  unsigned pos = 1, flags;
  uint64_t v;
  int i;
  if (ora_crc32(h, hs) != le32(h + hs)) return ORA_ERROR_ARCHIVE;
  if (pos == hs) return ORA_ERROR_ARCHIVE;
  flags = h[pos++];
  if (flags & 0x40)
  {
    unsigned s = read_varint(h + pos, hs - pos, &v);
    if (!s) return ORA_ERROR_ARCHIVE;
    pos += s;
    if (v == 0 || v + hs >= (uint64_t)1 << 63) return ORA_ERROR_ARCHIVE;
  }
  if (flags & 0x80)
  {
    unsigned s = read_varint(h + pos, hs - pos, &v);
    if (!s) return ORA_ERROR_ARCHIVE;
    pos += s;
  }
  fc->n = (int)(flags & 3) + 1;
  for (i = 0; i < fc->n; i++)
  {
    uint64_t sz;
    unsigned s = read_varint(h + pos, hs - pos, &fc->f[i].id);
    if (!s) return ORA_ERROR_ARCHIVE;
    pos += s;
    s = read_varint(h + pos, hs - pos, &sz);
    if (!s) return ORA_ERROR_ARCHIVE;
    pos += s;
    if (sz > hs - pos || sz > 20) return ORA_ERROR_ARCHIVE;
    fc->f[i].props_size = (uint32_t)sz;
    memcpy(fc->f[i].props, h + pos, (size_t)sz);
    pos += (unsigned)sz;
  }
  while (pos < hs) if (h[pos++] != 0) return ORA_ERROR_ARCHIVE;
  return ORA_OK;
}

/* XzDec.c:558-588 + :356-371 + :63-106: which chains the decoder accepts, reduced to
 * (optional one pre-filter) + LZMA2 -- the shapes on the hot path.  Longer chains and the
 * other branch converters are legal in the reference but out of scope: UNSUPPORTED here. */
static int classify_chain(const FilterChain *fc, int *filter_id, uint32_t *param, uint8_t *lzma2_prop)
{
  const Filter *last = &fc->f[fc->n - 1];
  *filter_id = 0; *param = 0;
  if (last->id != 0x21) return ORA_ERROR_UNSUPPORTED;
  if (last->props_size != 1) return ORA_ERROR_UNSUPPORTED;
  *lzma2_prop = last->props[0];
  if (*lzma2_prop > 40) return ORA_ERROR_UNSUPPORTED;
  if (fc->n == 1) return ORA_OK;
  if (fc->n > 2) return ORA_ERROR_UNSUPPORTED;
  if (fc->f[0].id == 3)
  {
    if (fc->f[0].props_size != 1) return ORA_ERROR_UNSUPPORTED;
    *filter_id = 3; *param = (uint32_t)fc->f[0].props[0] + 1;
    return ORA_OK;
  }
  if (fc->f[0].id == 4)
  {
    if (fc->f[0].props_size == 4) *param = le32(fc->f[0].props);
    else if (fc->f[0].props_size != 0) return ORA_ERROR_UNSUPPORTED;
    *filter_id = 4;
    return ORA_OK;
  }
  return ORA_ERROR_UNSUPPORTED;
}

int ora_xz_index(const uint8_t *src, size_t src_len, OraXzBlock *blocks, size_t max_blocks,
                 size_t *n_blocks, int *check_type)
{
  size_t end = src_len, idx_size, idx_off, pos, i, nb, data_total = 0, cur;
  const uint8_t *foot, *idx;
  unsigned flags, hflags, csz;
  uint64_t v;
  *n_blocks = 0; *check_type = 0;
  /* XzIn.c:141-190: stream padding is whole zero words */
  if ((end & 3) != 0 || end < 12) return ORA_ERROR_NO_ARCHIVE;
  if (memcmp(src + end - 2, "YZ", 2) != 0)
  {
    while (end > 0 && src[end - 1] == 0) end--;
    if ((end & 3) != 0 || end < 12) return ORA_ERROR_NO_ARCHIVE;
    if (memcmp(src + end - 2, "YZ", 2) != 0) return ORA_ERROR_NO_ARCHIVE;
  }
  foot = src + end - 12;
  flags = ((unsigned)foot[8] << 8) | foot[9];
  if (flags > 0xF) return ORA_ERROR_UNSUPPORTED;
  if (le32(foot) != ora_crc32(foot + 4, 6)) return ORA_ERROR_ARCHIVE;
  idx_size = ((size_t)le32(foot + 4) + 1) << 2;
  if (idx_size + 12 > end) return ORA_ERROR_NO_ARCHIVE;
  idx_off = end - 12 - idx_size;
  idx = src + idx_off;
  /* XzIn.c:71-114 */
  if (idx_size < 5 || idx[0] != 0) return ORA_ERROR_ARCHIVE;
  if (ora_crc32(idx, idx_size - 4) != le32(idx + idx_size - 4)) return ORA_ERROR_ARCHIVE;
  pos = 1;
  {
    unsigned s = read_varint(idx + pos, idx_size - 4 - pos, &v);
    if (!s) return ORA_ERROR_ARCHIVE;
    pos += s;
    nb = (size_t)v;
    if (nb != v || nb * 2 > idx_size - 4) return ORA_ERROR_ARCHIVE;
  }
  csz = check_size(flags & 0xF);
  for (i = 0; i < nb; i++)
  {
    uint64_t total, unpack;
    unsigned s = read_varint(idx + pos, idx_size - 4 - pos, &total);
    if (!s) return ORA_ERROR_ARCHIVE;
    pos += s;
    s = read_varint(idx + pos, idx_size - 4 - pos, &unpack);
    if (!s) return ORA_ERROR_ARCHIVE;
    pos += s;
    if (total == 0) return ORA_ERROR_ARCHIVE;
    if (i < max_blocks) { blocks[i].pack_size = total; blocks[i].unpack_size = unpack; }
    data_total += (size_t)((total + 3) & ~(uint64_t)3);
  }
  while ((pos & 3) != 0) if (idx[pos++] != 0) return ORA_ERROR_ARCHIVE;
  if (pos != idx_size - 4) return ORA_ERROR_ARCHIVE;
  if (12 + data_total > idx_off) return ORA_ERROR_ARCHIVE;
  cur = idx_off - data_total - 12;
  /* XzIn.c:215-223 */
  if (memcmp(src + cur, k_magic, 6) != 0) return ORA_ERROR_NO_ARCHIVE;
  {
    int r = parse_stream_header(src + cur, &hflags);
    if (r != ORA_OK) return r;
  }
  if (hflags != flags) return ORA_ERROR_ARCHIVE;
  *check_type = (int)(flags & 0xF);
  *n_blocks = nb;
  cur += 12;
  for (i = 0; i < nb && i < max_blocks; i++)
  {
    OraXzBlock *b = &blocks[i];
    uint64_t total = b->pack_size;
    FilterChain fc;
    uint32_t hs;
    int r;
    if (src[cur] == 0) return ORA_ERROR_ARCHIVE;
    hs = ((uint32_t)src[cur] << 2) + 4;
    if (hs + csz > total) return ORA_ERROR_ARCHIVE;
    r = parse_block_header(src + cur, &fc);
    if (r != ORA_OK) return r;
    r = classify_chain(&fc, &b->filter_id, &b->filter_param, &b->lzma2_prop);
    if (r != ORA_OK) return r;
    b->n_filters = fc.n;
    b->hdr_off = cur; b->hdr_size = hs;
    b->data_off = cur + hs;
    b->pack_size = total - hs - csz;
    b->check_off = b->data_off + ((b->pack_size + 3) & ~(uint64_t)3);
    cur += (size_t)((total + 3) & ~(uint64_t)3);
  }
  return ORA_OK;
}

int ora_xz_decode(uint8_t *dest, size_t *dest_len, const uint8_t *src, size_t *src_len,
                  int *status, uint64_t *num_blocks, int *stream_finished)
{
  size_t in_size = *src_len, out_cap = *dest_len, ip = 0, op = 0;
  uint8_t *rec = NULL;            /* index records as implied by the blocks seen */
  size_t rec_len = 0, rec_cap = 0;
  int res = ORA_OK;
  *status = ORA_ST_NOT_SPECIFIED; *num_blocks = 0; *stream_finished = 0;

#define STARVED() do { ip = in_size; *status = ORA_ST_NEEDS_MORE_INPUT; goto out; } while (0)
#define FAIL(code) do { res = (code); goto out; } while (0)

  for (;;)                                           /* one iteration per stream */
  {
    unsigned flags, csz, ctype;
    uint64_t nblk = 0;
    size_t k;
    /* XzDec.c:667-686 */
    for (k = 0; k < 6; k++)
    {
      if (ip + k >= in_size) STARVED();
      if (src[ip + k] != k_magic[k]) { ip += k; FAIL(ORA_ERROR_NO_ARCHIVE); }
    }
    if (in_size - ip < 12) STARVED();
    /* the reference parses the header only when the NEXT byte arrives (:669-683) */
    if (in_size - ip == 12) { ip = in_size; *status = ORA_ST_NEEDS_MORE_INPUT; goto out; }
    res = parse_stream_header(src + ip, &flags);
    ip += 12;
    if (res != ORA_OK) goto out;
    ctype = flags & 0xF;
    csz = check_size(ctype);
    rec_len = 0;

    for (;;)                                         /* blocks, XzDec.c:688-760 */
    {
      uint32_t hs;
      FilterChain fc;
      int filter_id; uint32_t fparam; uint8_t l2prop;
      size_t blk_in, blk_out, pad;
      int st;
      if (ip >= in_size) STARVED();
      if (src[ip] == 0) break;                       /* index indicator */
      hs = ((uint32_t)src[ip] << 2) + 4;
      if (in_size - ip < hs) STARVED();
      if (in_size - ip == hs) { ip = in_size; *status = ORA_ST_NEEDS_MORE_INPUT; goto out; }
      res = parse_block_header(src + ip, &fc);
      ip += hs;
      if (res != ORA_OK) goto out;
      res = classify_chain(&fc, &filter_id, &fparam, &l2prop);
      if (res != ORA_OK) goto out;
      if (ctype == 10) FAIL(ORA_ERROR_UNSUPPORTED);

      blk_in = in_size - ip;
      res = ora_lzma2_decode_unit(dest + op, out_cap - op, &blk_out, src + ip, &blk_in, l2prop, ORA_FINISH_ANY, &st);
      ip += blk_in;
      if (res != ORA_OK) { op += blk_out; goto out; }
      if (st != ORA_ST_FINISHED_WITH_MARK)
      {
        op += blk_out;
        if (st == ORA_ST_NEEDS_MORE_INPUT) { ip = in_size; *status = ORA_ST_NEEDS_MORE_INPUT; }
        else *status = ORA_ST_NOT_FINISHED;          /* destination exhausted */
        goto out;
      }
      if (filter_id == 4)
      {
        uint32_t xs = 0;
        (void)ora_x86_convert(dest + op, blk_out, fparam, &xs, 0);
      }
      else if (filter_id == 3)
      {
        uint8_t ds[256];
        ora_delta_init(ds);
        ora_delta_decode(ds, fparam, dest + op, blk_out);
      }
      /* XzDec.c:637-651: remember the index record this block implies */
      if (rec_cap - rec_len < 32)
      {
        rec_cap = rec_cap ? rec_cap * 2 : 256;
        rec = (uint8_t *)realloc(rec, rec_cap);
      }
      rec_len += write_varint(rec + rec_len, (uint64_t)blk_in + hs + csz);
      rec_len += write_varint(rec + rec_len, blk_out);
      nblk++;
      (*num_blocks)++;
      /* XzDec.c:728-760: padding, then digest */
      pad = (4 - (blk_in & 3)) & 3;
      for (k = 0; k < pad; k++)
      {
        if (ip >= in_size) { op += blk_out; STARVED(); }
        if (src[ip++] != 0) { op += blk_out; FAIL(ORA_ERROR_CRC); }
      }
      if (in_size - ip < csz) { op += blk_out; STARVED(); }
      if (csz != 0 && in_size - ip == csz) { op += blk_out; ip = in_size; *status = ORA_ST_NEEDS_MORE_INPUT; goto out; }
      if (ctype == 1)
      {
        if (ora_crc32(dest + op, blk_out) != le32(src + ip)) { op += blk_out; ip += csz; FAIL(ORA_ERROR_CRC); }
      }
      else if (ctype == 4)
      {
        uint64_t want = (uint64_t)le32(src + ip) | ((uint64_t)le32(src + ip + 4) << 32);
        if (ora_crc64(dest + op, blk_out) != want) { op += blk_out; ip += csz; FAIL(ORA_ERROR_CRC); }
      }
      ip += csz;
      op += blk_out;
    }

    /* index, XzDec.c:694-703,762-822 */
    {
      uint8_t pre[16];
      size_t pre_len = 1 + write_varint(pre + 1, nblk), idx_start = ip, idx_len;
      uint32_t crc;
      pre[0] = 0;
      for (k = 0; k < pre_len; k++)
      {
        if (ip >= in_size) STARVED();
        if (src[ip++] != pre[k]) FAIL(ORA_ERROR_CRC);
      }
      if (in_size - ip < rec_len) STARVED();
      ip += rec_len;
      while (((ip - idx_start) & 3) != 0)
      {
        if (ip >= in_size) STARVED();
        if (src[ip++] != 0) FAIL(ORA_ERROR_CRC);
      }
      if (ip >= in_size) STARVED();
      if (rec_len != 0 && memcmp(src + idx_start + pre_len, rec, rec_len) != 0) FAIL(ORA_ERROR_CRC);
      idx_len = ip - idx_start;
      if (in_size - ip < 4) STARVED();
      crc = ora_crc32(src + idx_start, idx_len);
      ip += 4;
      if (ip >= in_size) STARVED();
      if (crc != le32(src + ip - 4)) FAIL(ORA_ERROR_CRC);
      /* footer, XzDec.c:491-498,824-842 */
      if (in_size - ip < 12) STARVED();
      {
        const uint8_t *f = src + ip;
        int ok = (idx_len + 4 == (((size_t)le32(f + 4) + 1) << 2)) && le32(f) == ora_crc32(f + 4, 6) &&
                 flags == (((unsigned)f[8] << 8) | f[9]) && f[10] == 'Y' && f[11] == 'Z';
        ip += 12;
        if (!ok) FAIL(ORA_ERROR_CRC);
      }
    }
    /* padding, XzDec.c:844-860 */
    {
      size_t padn = 0;
      while (ip < in_size && src[ip] == 0) { ip++; padn++; }
      if (ip >= in_size)
      {
        *stream_finished = (padn & 3) == 0;
        *status = ORA_ST_NEEDS_MORE_INPUT;
        goto out;
      }
      if ((padn & 3) != 0) FAIL(ORA_ERROR_NO_ARCHIVE);
    }
  }
out:
  free(rec);
  *src_len = ip; *dest_len = op;
  return res;
#undef STARVED
#undef FAIL
}
