/* lzb_containers.cu -- host front ends that turn container byte streams into units for the GPU:
 *
 *   LZMA2 : walk the chunk headers (reference Lzma2Dec.c:14-26,98-157) and cut at dictionary resets
 *           (control 0x01 or >= 0xE0) -- the only points where a chunk depends on nothing before it
 *           (SURVEY.md F2).  The decoding warp re-reads the headers of its own unit.
 *   xz    : forward walk of stream header / block headers / LZMA2 chunk headers / index / footer
 *           (reference XzDec.c:482-556,604-870, Xz.c:12-43); the reference's block headers carry no
 *           sizes (F10), the chunk walk is what finds a block's end.  Also the backward index reader
 *           (reference XzIn.c:71-114,141-224).
 *
 * Results of the per-unit decodes are stitched back in stream order so that the whole-buffer calls
 * return what the reference's sequential decoder returns: the first failing unit decides.
 */
#include "lzb_host.h"
#include "lzb_xz.h"
#include "lzb_sha256.cuh"

#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

enum { LZB_FINISH_UNIT = 2 };   /* internal: validate the unit's last chunk, then stop at its end */

/* ------------------------------------------------------------------ LZMA2 walk ---------------- */
struct L2Walk
{
  std::vector<LzbUnit> units;   /* offsets relative to the walked buffer, dst_off cumulative */
  size_t end_off = 0;           /* just past the 0x00 marker, or where the walk stopped */
  bool has_eos = false;
  bool malformed = false;       /* stopped at a header the decoder will reject */
  bool truncated = false;       /* stopped because a header or payload runs past the end */
  uint64_t unpack_total = 0;
};

static SRes lzma2_dict_size(Byte prop, UInt32 *dict)
{
  if (prop > 40) return SZ_ERROR_UNSUPPORTED;                      /* Lzma2Dec.c:64-66 */
  *dict = prop == 40 ? 0xFFFFFFFFu : ((UInt32)(2 | (prop & 1)) << (prop / 2 + 11));
  return SZ_OK;
}

static void lzma2_walk(const Byte *src, size_t n, UInt32 dict, L2Walk &w)
{
  size_t p = 0;
  uint64_t out = 0;
  while (p < n)
  {
    const Byte c = src[p];
    size_t hdr, pack, unpack;
    unsigned lclp = 0, pb = 0;
    bool has_props = false;
    if (c == 0) { p++; w.has_eos = true; break; }
    if ((c & 0x80) == 0)
    {
      if (c > 2) { w.malformed = true; break; }
      hdr = 3;
      if (n - p < hdr) { w.truncated = true; break; }
      unpack = (((size_t)src[p + 1] << 8) | src[p + 2]) + 1;
      pack = unpack;
    }
    else
    {
      hdr = 5 + (((c >> 5) & 3) >= 2 ? 1 : 0);
      if (n - p < hdr) { w.truncated = true; break; }
      unpack = (((size_t)(c & 0x1F) << 16) | ((size_t)src[p + 1] << 8) | src[p + 2]) + 1;
      pack = (((size_t)src[p + 3] << 8) | src[p + 4]) + 1;
      if (hdr == 6)
      {
        unsigned b = src[p + 5];
        if (b >= 225 || (b % 9) + ((b / 9) % 5) > 4) { w.malformed = true; break; }
        lclp = (b % 9) + ((b / 9) % 5);
        pb = b / 45;
        has_props = true;
      }
    }
    if (c == 1 || c >= 0xE0 || w.units.empty())
    {
      LzbUnit u;
      memset(&u, 0, sizeof u);
      u.src_off = p; u.dst_off = out; u.dict_size = dict;
      u.codec = LZB_CODEC_LZMA2; u.finish = LZB_FINISH_UNIT;
      w.units.push_back(u);
    }
    LzbUnit &u = w.units.back();
    if (has_props) { if (lclp > u.lc) u.lc = (uint8_t)lclp; if (pb > u.pb) u.pb = (uint8_t)pb; }
    if (n - p - hdr < pack) { w.truncated = true; u.src_len = (uint32_t)std::min<size_t>(n - u.src_off, 0xFFFFFFFFu); break; }
    if ((uint64_t)u.dst_cap + unpack > 0xFFFFFFFFull || (p + hdr + pack - u.src_off) > 0xFFFFFFFFull) { w.malformed = true; break; }
    u.dst_cap += (uint32_t)unpack;
    u.src_len = (uint32_t)(p + hdr + pack - u.src_off);
    out += unpack;
    p += hdr + pack;
  }
  w.end_off = p;
  w.unpack_total = out;
}

extern "C" SRes Lzma2_IndexUnits(const Byte *src, size_t srcLen, Byte prop, LzbUnit *units, size_t maxUnits, size_t *n,
                                 size_t *endOffset, int *hasEos)
{
  UInt32 dict;
  SRes r = lzma2_dict_size(prop, &dict);
  if (r != SZ_OK) return r;
  L2Walk w;
  lzma2_walk(src, srcLen, dict, w);
  *n = w.units.size();
  for (size_t i = 0; i < w.units.size() && i < maxUnits; i++) units[i] = w.units[i];
  *endOffset = w.end_off;
  *hasEos = w.has_eos ? 1 : 0;
  return w.malformed ? SZ_ERROR_DATA : SZ_OK;
}

/* Shape the walked units for one whole-buffer call with output limit `cap` and the caller's finish
 * mode: units past the limit are dropped; the last decoded unit sees the rest of the input and of the
 * output so that it behaves exactly like the sequential decoder from there on. */
static void shape_units(L2Walk &w, size_t in_size, size_t cap, int finish, bool dirty_tail)
{
  if (w.units.empty())
  {
    LzbUnit u;
    memset(&u, 0, sizeof u);
    u.codec = LZB_CODEC_LZMA2;
    w.units.push_back(u);
  }
  size_t keep = 0;
  uint64_t acc = 0;
  for (; keep < w.units.size(); keep++)
  {
    if (keep + 1 == w.units.size()) { keep++; break; }
    if (acc + w.units[keep].dst_cap >= cap) { keep++; break; }        /* limit falls inside or at the end of this unit */
    acc += w.units[keep].dst_cap;
  }
  w.units.resize(keep);
  LzbUnit &last = w.units.back();
  last.src_len = (uint32_t)std::min<size_t>(in_size - last.src_off, 0xFFFFFFFFu);
  last.dst_cap = (uint32_t)std::min<uint64_t>(cap - std::min<uint64_t>(cap, last.dst_off), 0xFFFFFFFFu);
  last.finish = (uint8_t)finish;
  if (dirty_tail) { last.lc = 4; last.lp = 0; last.pb = 4; }        /* headers the walk did not vet: worst-case model */
}

extern "C" SRes Lzma2DecodeParallel(LzbContext *c, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, Byte prop,
                                    ELzmaFinishMode finishMode, ELzmaStatus *status)
{
  const size_t cap = *destLen, in_size = *srcLen;
  *destLen = *srcLen = 0;
  *status = LZMA_STATUS_NOT_SPECIFIED;
  UInt32 dict;
  SRes r = lzma2_dict_size(prop, &dict);
  if (r != SZ_OK) return r;
  if (!c) c = lzb_default_context();
  if (!c) return SZ_ERROR_FAIL;
  L2Walk w;
  lzma2_walk(src, in_size, dict, w);
  /* a unit that would start exactly at the limit is never reached with FINISH_ANY */
  shape_units(w, in_size, cap, (int)finishMode, w.malformed || w.truncated || !w.has_eos);
  for (auto &u : w.units) u.dict_size = dict;
  std::vector<LzbResult> res(w.units.size());
  r = Lzb_DecodeUnitsHost(c, src, in_size, dest, cap, w.units.data(), res.data(), w.units.size());
  if (r != SZ_OK) return r;
  size_t done = 0;
  int carried = LZMA_STATUS_NOT_SPECIFIED;               /* *status is sticky across chunks in the reference */
  for (size_t k = 0; k < w.units.size(); k++)
  {
    const LzbUnit &u = w.units[k];
    const LzbResult &x = res[k];
    done += x.dst_done;
    if (x.status != LZB_ST_INHERIT) carried = x.status;
    const bool last = k + 1 == w.units.size();
    const bool clean = x.res == SZ_OK && x.dst_done == u.dst_cap && carried != LZMA_STATUS_NEEDS_MORE_INPUT &&
                       carried != LZMA_STATUS_FINISHED_WITH_MARK;
    if (last || !clean)
    {
      *destLen = done;
      *srcLen = (size_t)u.src_off + x.src_used;
      *status = (ELzmaStatus)carried;
      return x.res;
    }
  }
  return SZ_ERROR_FAIL;
}

extern "C" SRes Lzma2Decode(Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, Byte prop,
                            ELzmaFinishMode finishMode, ELzmaStatus *status, ISzAlloc *alloc)
{
  (void)alloc;
  SRes r = Lzma2DecodeParallel(nullptr, dest, destLen, src, srcLen, prop, finishMode, status);
  if (r == SZ_OK && *status == LZMA_STATUS_NEEDS_MORE_INPUT) r = SZ_ERROR_INPUT_EOF;   /* Lzma2Dec.c:351-352 */
  return r;
}

/* ------------------------------------------------------------------ xz helpers ---------------- */
static const Byte k_xz_magic[6] = { 0xFD, '7', 'z', 'X', 'Z', 0 };

/* host CRC-32 of a few header bytes (container framing only; payload digests run on the GPU) */
uint32_t lzb_small_crc32_update(uint32_t reg, const Byte *p, size_t n)
{
  for (size_t i = 0; i < n; i++)
  {
    reg ^= p[i];
    for (int k = 0; k < 8; k++) reg = (reg >> 1) ^ (0xEDB88320u & (0u - (reg & 1)));
  }
  return reg;
}

uint32_t lzb_small_crc32(const Byte *p, size_t n) { return ~lzb_small_crc32_update(0xFFFFFFFFu, p, n); }

/* xz format 1.2, "multibyte integers": seven bits per byte, least significant group first, bit 7 set on every byte
 * but the last, nine bytes at most; a final 0x00 that is not the only byte would be a padded (non-minimal) encoding
 * and is refused.  Returns the encoded length, 0 when the bytes are not a valid integer (as the reference's
 * Xz_ReadVarInt does, Xz.h:21). */
extern "C" unsigned Xz_ReadVarInt(const Byte *p, size_t maxSize, UInt64 *value)
{
  const size_t longest = maxSize < 9 ? maxSize : 9;
  UInt64 v = 0;
  unsigned len = 0;
  bool ended = false;
  while (len < longest && !ended)
  {
    const Byte b = p[len];
    v |= (UInt64)(b & 0x7Fu) << (7 * len);
    len++;
    ended = b < 0x80;
  }
  *value = v;
  if (!ended) return 0;
  if (len > 1 && p[len - 1] == 0) return 0;
  return len;
}

extern "C" unsigned Xz_WriteVarInt(Byte *buf, UInt64 v)
{
  unsigned len = 0;
  for (; v >= 0x80; v >>= 7) buf[len++] = (Byte)(v | 0x80);
  buf[len++] = (Byte)v;
  return len;
}

extern "C" unsigned XzFlags_GetCheckSize(CXzStreamFlags f)
{
  int t = f & 0xF;                                                 /* reference Xz.c:39-43 */
  return t == 0 ? 0 : (4u << ((t - 1) / 3));
}

extern "C" SRes Xz_ParseHeader(CXzStreamFlags *p, const Byte *buf)
{
  /* reference XzDec.c:482-489 */
  *p = (CXzStreamFlags)(((unsigned)buf[6] << 8) | buf[7]);
  if (lzb_small_crc32(buf + 6, 2) != lzb_le32(buf + 8)) return SZ_ERROR_NO_ARCHIVE;
  return *p <= 0xF ? SZ_OK : SZ_ERROR_UNSUPPORTED;
}

/* Block header, xz format 3.1:
 *   size byte s (real size = (s + 1) * 4, the last four bytes are the CRC32 of everything before them)
 *   block flags: bits 0-1 = number of filters - 1, bit 6 = compressed size present, bit 7 = uncompressed size present
 *   [compressed size] [uncompressed size]           multibyte integers
 *   per filter: id, size of properties, properties
 *   zero padding up to the CRC32
 * Everything wrong with a header is SZ_ERROR_ARCHIVE (what the reference's XzBlock_Parse answers, Xz.h:85). */
namespace {
struct HeaderFields
{
  const Byte *at, *end;                                   /* the bytes between the size byte and the CRC32 */
  bool take_byte(Byte &b) { if (at == end) return false; b = *at++; return true; }
  bool take_int(UInt64 &v)
  {
    const unsigned n = Xz_ReadVarInt(at, (size_t)(end - at), &v);
    at += n;
    return n != 0;
  }
  bool take_bytes(Byte *dst, UInt64 n)
  {
    if (n > (UInt64)(end - at)) return false;
    memcpy(dst, at, (size_t)n);
    at += n;
    return true;
  }
  bool rest_is_zero() const { for (const Byte *q = at; q != end; q++) if (*q != 0) return false; return true; }
};
}  // namespace

extern "C" SRes XzBlock_Parse(CXzBlock *p, const Byte *h)
{
  const UInt32 body = (UInt32)h[0] * 4;                    /* offset of the CRC32 */
  if (lzb_le32(h + body) != lzb_small_crc32(h, body)) return SZ_ERROR_ARCHIVE;
  HeaderFields f = { h + 1, h + body };
  if (!f.take_byte(p->flags)) return SZ_ERROR_ARCHIVE;
  if (XzBlock_HasPackSize(p))
  {
    /* a block holds at least one byte, and its size plus this header must stay a positive 63-bit number */
    if (!f.take_int(p->packSize) || p->packSize == 0 || p->packSize + body >= ((UInt64)1 << 63)) return SZ_ERROR_ARCHIVE;
  }
  if (XzBlock_HasUnpackSize(p) && !f.take_int(p->unpackSize)) return SZ_ERROR_ARCHIVE;
  for (int i = 0, n = XzBlock_GetNumFilters(p); i < n; i++)
  {
    CXzFilter &flt = p->filters[i];
    UInt64 props_bytes;
    if (!f.take_int(flt.id) || !f.take_int(props_bytes) || props_bytes > XZ_FILTER_PROPS_SIZE_MAX ||
        !f.take_bytes(flt.props, props_bytes))
      return SZ_ERROR_ARCHIVE;
    flt.propsSize = (UInt32)props_bytes;
  }
  return f.rest_is_zero() ? SZ_OK : SZ_ERROR_ARCHIVE;
}

/* Filter chains on the hot path: [x86 | delta] + LZMA2 (reference XzDec.c:63-106,356-371,558-588).
 * The other branch converters / longer chains are legal in the reference and out of scope here. */
SRes lzb_xz_classify_chain(const CXzBlock *blk, LzbXzBlock *b)
{
  const int nf = XzBlock_GetNumFilters(blk);
  const CXzFilter *f = blk->filters;
  const CXzFilter &last = f[nf - 1];
  b->filter = LZB_FILTER_NONE; b->filter_param = 0;
  if (last.id != XZ_ID_LZMA2 || last.propsSize != 1 || last.props[0] > 40) return SZ_ERROR_UNSUPPORTED;
  b->lzma2_prop = last.props[0];
  if (nf == 1) return SZ_OK;
  if (nf > 2) { lzb_set_error("xz filter chains longer than 2 are outside the GPU hot path"); return SZ_ERROR_UNSUPPORTED; }
  if (f[0].id == XZ_ID_Delta)
  {
    if (f[0].propsSize != 1) return SZ_ERROR_UNSUPPORTED;
    b->filter = LZB_FILTER_DELTA; b->filter_param = (UInt32)f[0].props[0] + 1;
    return SZ_OK;
  }
  if (f[0].id == XZ_ID_X86)
  {
    if (f[0].propsSize == 4) b->filter_param = lzb_le32(f[0].props);
    else if (f[0].propsSize != 0) return SZ_ERROR_UNSUPPORTED;
    b->filter = LZB_FILTER_X86;
    return SZ_OK;
  }
  if (f[0].id >= XZ_ID_PPC && f[0].id <= XZ_ID_SPARC)
  {
    /* stateless RISC converters; start-offset alignment rules of BraState_SetProps (XzDec.c:84-104) */
    UInt32 v = 0;
    if (f[0].propsSize == 4) v = lzb_le32(f[0].props);
    else if (f[0].propsSize != 0) return SZ_ERROR_UNSUPPORTED;
    const UInt32 need = f[0].id == XZ_ID_IA64 ? 15u : (f[0].id == XZ_ID_ARMT ? 1u : 3u);
    if (v & need) return SZ_ERROR_UNSUPPORTED;
    b->filter = (int32_t)f[0].id; b->filter_param = v;
    return SZ_OK;
  }
  lzb_set_error("xz filter id %llu is not a filter of the reference", (unsigned long long)f[0].id);
  return SZ_ERROR_UNSUPPORTED;
}

/* ------------------------------------------------------------------ xz forward walk ----------- */
/* Forward walk.  in_blocks: `src` starts at a block header (or the index) of a stream whose header has been
 * seen already (its flags are passed in); the index / footer verdicts are then meaningless to the caller,
 * which takes only the block table. */
void lzb_xz_walk(const Byte *src, size_t n, XzWalk &w, bool in_blocks, unsigned in_flags)
{
  size_t ip = 0;
  uint64_t op = 0;
  bool skip_header = in_blocks;
#define XZ_END(code, st, at) do { w.end_res = (code); w.end_status = (st); w.end_src = (at); return; } while (0)
#define XZ_STARVED() XZ_END(SZ_OK, CODER_STATUS_NEEDS_MORE_INPUT, n)
  for (;;)
  {
    CXzStreamFlags flags = (CXzStreamFlags)in_flags;
    SRes r = SZ_OK;
    if (!skip_header)
    {
      for (size_t k = 0; k < 6; k++)
      {
        if (ip + k >= n) XZ_STARVED();
        if (src[ip + k] != k_xz_magic[k]) XZ_END(SZ_ERROR_NO_ARCHIVE, CODER_STATUS_NOT_SPECIFIED, ip + k);
      }
      if (n - ip <= 12) XZ_STARVED();                /* the header is parsed when the next byte arrives */
      r = Xz_ParseHeader(&flags, src + ip);
      ip += 12;
      if (r != SZ_OK) XZ_END(r, CODER_STATUS_NOT_SPECIFIED, ip);
    }
    skip_header = false;
    const unsigned ctype = flags & 0xF, csz = XzFlags_GetCheckSize(flags);
    const size_t first_block = w.blocks.size();

    for (;;)
    {
      if (ip >= n) XZ_STARVED();
      if (src[ip] == 0) break;
      const UInt32 hs = ((UInt32)src[ip] << 2) + 4;
      if (n - ip <= hs) XZ_STARVED();
      CXzBlock blk;
      LzbXzBlock b;
      memset(&b, 0, sizeof b);
      memset(&blk, 0, sizeof blk);
      r = XzBlock_Parse(&blk, src + ip);
      b.hdr_off = ip; b.hdr_size = hs;
      ip += hs;
      if (r != SZ_OK) XZ_END(r, CODER_STATUS_NOT_SPECIFIED, ip);
      r = lzb_xz_classify_chain(&blk, &b);
      if (r != SZ_OK) XZ_END(r, CODER_STATUS_NOT_SPECIFIED, ip);
      b.check_type = (uint8_t)ctype;
      UInt32 dict;
      lzma2_dict_size(b.lzma2_prop, &dict);
      L2Walk lw;
      lzma2_walk(src + ip, n - ip, dict, lw);
      b.data_off = ip;
      b.first_unit = (uint32_t)w.units.size();
      const bool whole = lw.has_eos;
      if (!whole)
      {
        /* payload cut short or malformed: hand everything that is left to the decoder, it returns what
         * the sequential decoder returns from here (NEEDS_MORE_INPUT or SZ_ERROR_DATA) */
        shape_units(lw, n - ip, (size_t)0xFFFFFFFFu, LZMA_FINISH_ANY, true);
      }
      for (auto u : lw.units)
      {
        u.src_off += ip;
        u.dst_off += op;
        w.units.push_back(u);
      }
      if (whole && !lw.units.empty())
      {
        /* the 0x00 end marker belongs to the last unit: with FINISH_END it must be there */
        LzbUnit &last = w.units.back();
        last.src_len = (uint32_t)(ip + lw.end_off - last.src_off);
        last.finish = LZMA_FINISH_END;
      }
      if (whole && lw.units.empty())
      {
        LzbUnit u;                                   /* empty block: a lone end marker */
        memset(&u, 0, sizeof u);
        u.src_off = ip; u.src_len = 1; u.dst_off = op; u.dst_cap = 0; u.dict_size = dict;
        u.codec = LZB_CODEC_LZMA2; u.finish = LZMA_FINISH_END;
        w.units.push_back(u);
      }
      b.n_units = (uint32_t)w.units.size() - b.first_unit;
      b.pack_size = lw.end_off;
      b.unpack_size = lw.unpack_total;
      b.complete = whole ? 1 : 0;
      w.blocks.push_back(b);
      w.after_block.push_back(SZ_OK);
      if (!whole) { w.partial = (int)w.blocks.size() - 1; XZ_END(SZ_OK, CODER_STATUS_NEEDS_MORE_INPUT, n); }
      ip += lw.end_off;
      op += lw.unpack_total;
      LzbXzBlock &bb = w.blocks.back();
      /* padding to a multiple of four, then the stored digest (reference XzDec.c:728-760) */
      const size_t pad = (4 - (lw.end_off & 3)) & 3;
      for (size_t k = 0; k < pad; k++)
      {
        if (ip >= n) XZ_STARVED();
        if (src[ip++] != 0) { w.after_block.back() = SZ_ERROR_CRC; XZ_END(SZ_ERROR_CRC, CODER_STATUS_NOT_SPECIFIED, ip); }
      }
      if (n - ip < csz || (csz != 0 && n - ip == csz)) { bb.complete = 2; XZ_STARVED(); }   /* digest not all here */
      bb.check_off = ip;
      bb.expect = csz == 4 ? lzb_le32(src + ip) : (csz == 8 ? lzb_le64(src + ip) : 0);
      ip += csz;
      w.out_total = op;
    }

    /* index + footer (reference XzDec.c:694-703,762-842) */
    {
      const uint64_t nblk = w.blocks.size() - first_block;
      Byte pre[16];
      const size_t pre_len = 1 + Xz_WriteVarInt(pre + 1, nblk), idx_start = ip;
      pre[0] = 0;
      for (size_t k = 0; k < pre_len; k++)
      {
        if (ip >= n) XZ_STARVED();
        if (src[ip++] != pre[k]) XZ_END(SZ_ERROR_CRC, CODER_STATUS_NOT_SPECIFIED, ip);
      }
      std::vector<Byte> rec;
      for (size_t k = first_block; k < w.blocks.size(); k++)
      {
        Byte t[24];
        unsigned m = Xz_WriteVarInt(t, w.blocks[k].pack_size + w.blocks[k].hdr_size + csz);
        m += Xz_WriteVarInt(t + m, w.blocks[k].unpack_size);
        rec.insert(rec.end(), t, t + m);
      }
      if (n - ip < rec.size()) XZ_STARVED();
      const size_t rec_at = ip;
      ip += rec.size();
      while (((ip - idx_start) & 3) != 0)
      {
        if (ip >= n) XZ_STARVED();
        if (src[ip++] != 0) XZ_END(SZ_ERROR_CRC, CODER_STATUS_NOT_SPECIFIED, ip);
      }
      if (ip >= n) XZ_STARVED();
      if (!rec.empty() && memcmp(src + rec_at, rec.data(), rec.size()) != 0) XZ_END(SZ_ERROR_CRC, CODER_STATUS_NOT_SPECIFIED, ip);
      const size_t idx_len = ip - idx_start;
      if (n - ip < 4) XZ_STARVED();
      const uint32_t crc = lzb_small_crc32(src + idx_start, idx_len);
      ip += 4;
      if (ip >= n) XZ_STARVED();
      if (crc != lzb_le32(src + ip - 4)) XZ_END(SZ_ERROR_CRC, CODER_STATUS_NOT_SPECIFIED, ip);
      if (n - ip < 12) XZ_STARVED();
      const Byte *f = src + ip;
      const bool ok = (idx_len + 4 == (((size_t)lzb_le32(f + 4) + 1) << 2)) && lzb_le32(f) == lzb_small_crc32(f + 4, 6) &&
                      flags == (((unsigned)f[8] << 8) | f[9]) && f[10] == 'Y' && f[11] == 'Z';
      ip += 12;
      if (!ok) XZ_END(SZ_ERROR_CRC, CODER_STATUS_NOT_SPECIFIED, ip);
    }
    {
      size_t padn = 0;
      while (ip < n && src[ip] == 0) { ip++; padn++; }
      if (ip >= n) { w.finished = (padn & 3) == 0; XZ_END(SZ_OK, CODER_STATUS_NEEDS_MORE_INPUT, ip); }
      if ((padn & 3) != 0) XZ_END(SZ_ERROR_NO_ARCHIVE, CODER_STATUS_NOT_SPECIFIED, ip);
    }
  }
#undef XZ_END
#undef XZ_STARVED
}

/* ------------------------------------------------------------------ xz block prefetch --------- */
/* Decode, in ONE batched GPU pass, the whole blocks at the front of src[0..n) -- all of them in parallel,
 * and the dictionary-reset groups inside them -- into dest at the offsets a sequential decoder would give
 * them, with BCJ/Delta applied and the check digest computed (not compared: XzUnpacker_Code does that when
 * it reaches the block footer).  Only blocks whose units all decoded cleanly are reported; a block that is
 * cut short, does not fit in dest, or fails is left to the resumable chain, which reproduces the reference's
 * return values for it.  Advisory: returns the blocks it could do in `ready` (possibly none). */
SRes lzb_xz_prefetch(LzbContext *c, const Byte *src, size_t n, bool in_blocks, unsigned flags, Byte *dest, size_t cap,
                     std::vector<XzReady> &ready)
{
  ready.clear();
  XzWalk w;
  lzb_xz_walk(src, n, w, in_blocks, flags);
  /* prefix of blocks whose payload is all here and whose output fits */
  size_t nb = 0, nu = 0;
  for (; nb < w.blocks.size(); nb++)
  {
    const LzbXzBlock &b = w.blocks[nb];
    if (b.complete == 0) break;
    const uint64_t off = w.units[b.first_unit].dst_off;
    if (off + b.unpack_size > cap) break;
    nu = b.first_unit + b.n_units;
  }
  if (nb == 0) return SZ_OK;
  w.blocks.resize(nb);
  w.units.resize(nu);
  if (!c) c = lzb_default_context();
  if (!c) return SZ_ERROR_FAIL;

  /* single-unit blocks: filter + CRC fused into the decode launch; others (and SHA-256): second pass */
  struct Post { size_t block; LzbUnit u; };
  std::vector<Post> post;
  for (size_t k = 0; k < nb; k++)
  {
    LzbXzBlock &b = w.blocks[k];
    const bool crc = b.check_type == LZB_CHECK_CRC32 || b.check_type == LZB_CHECK_CRC64;
    const bool sha = b.check_type == LZB_CHECK_SHA256;
    if (b.filter == LZB_FILTER_NONE && !crc && !sha) continue;
    /* the decode kernel fuses x86 / Delta + CRC for one-unit blocks; everything else is a second pass */
    const bool fused = b.n_units == 1 && (b.filter == LZB_FILTER_NONE || b.filter == LZB_FILTER_X86 || b.filter == LZB_FILTER_DELTA);
    if (fused)
    {
      LzbUnit &u = w.units[b.first_unit];
      u.filter = (uint8_t)b.filter; u.filter_param = b.filter_param;
      if (crc) u.check = b.check_type;
    }
    if (!fused || sha)
    {
      Post q;
      memset(&q.u, 0, sizeof q.u);
      q.block = k;
      q.u.dst_off = w.units[b.first_unit].dst_off; q.u.dst_cap = (uint32_t)b.unpack_size;
      if (!fused) { q.u.filter = (uint8_t)b.filter; q.u.filter_param = b.filter_param; if (crc) q.u.check = b.check_type; }
      post.push_back(q);
    }
  }
  std::vector<LzbResult> res(w.units.size());
  SRes r = Lzb_DecodeUnitsHost(c, src, n, dest, cap, w.units.data(), res.data(), w.units.size());
  if (r != SZ_OK) return r;

  /* which blocks came out clean */
  std::vector<char> clean(nb, 1);
  std::vector<uint64_t> digest(nb, 0);
  for (size_t k = 0; k < nb; k++)
  {
    const LzbXzBlock &b = w.blocks[k];
    for (uint32_t j = 0; j < b.n_units; j++)
    {
      const LzbUnit &u = w.units[b.first_unit + j];
      const LzbResult &x = res[b.first_unit + j];
      const bool lastu = j + 1 == b.n_units;
      bool ok;
      if (lastu) ok = x.res == SZ_OK && x.status == LZMA_STATUS_FINISHED_WITH_MARK && x.dst_done == u.dst_cap;
      else ok = x.res == SZ_OK && x.dst_done == u.dst_cap && x.status != LZMA_STATUS_NEEDS_MORE_INPUT &&
                x.status != LZMA_STATUS_FINISHED_WITH_MARK;
      if (!ok) clean[k] = 0;
    }
    digest[k] = res[b.first_unit].check;      /* fused one-unit blocks; overwritten by the second pass otherwise */
  }
  std::vector<LzbShaJob> jobs;
  std::vector<size_t> job_block;
  if (!post.empty())
  {
    /* BCJ/Delta + digest over whole multi-unit blocks, SHA-256 over every block that asks for it */
    std::vector<Post> live;
    for (auto &q : post) if (clean[q.block]) live.push_back(q);
    if (!live.empty())
    {
      std::lock_guard<std::mutex> g(c->mu);
      LzbSlot &s = c->slot[0];
      uint64_t lo = ~0ull, hi = 0;
      for (auto &q : live) { lo = std::min<uint64_t>(lo, q.u.dst_off); hi = std::max<uint64_t>(hi, q.u.dst_off + q.u.dst_cap); }
      if (hi <= lo) hi = lo + 1;
      std::vector<LzbUnit> rel;
      for (auto &q : live)
      {
        LzbUnit u = q.u; u.dst_off -= lo; rel.push_back(u);
        if (w.blocks[q.block].check_type == LZB_CHECK_SHA256)
        {
          LzbShaJob j;
          memset(&j, 0, sizeof j);
          j.off = u.dst_off; j.len = u.dst_cap; j.finalize = 1;
          lzb::sha256_init(j.state);
          jobs.push_back(j); job_block.push_back(q.block);
        }
      }
      const size_t job_bytes = jobs.size() * sizeof(LzbShaJob);
      if (!lzb_slot_reserve(c, s, job_bytes + 16, (size_t)(hi - lo), rel.size(), 0)) return SZ_ERROR_MEM;
      std::vector<LzbResult> pres(rel.size());
      bool ok = lzb_cuda_ok(cudaMemcpyAsync(s.d_out, dest + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, s.stream), "H2D blocks") &&
                lzb_cuda_ok(cudaMemcpyAsync(s.d_units, rel.data(), rel.size() * sizeof(LzbUnit), cudaMemcpyHostToDevice, s.stream), "H2D units");
      if (!ok) return SZ_ERROR_FAIL;
      int rc = lzb_launch_post_only(s.d_out, s.d_units, s.d_results, (uint32_t)rel.size(), s.stream);
      if (rc != 0) { lzb_set_error("post launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return SZ_ERROR_FAIL; }
      if (!jobs.empty())
      {
        /* the job table rides in the (otherwise unused) packed staging area of the slot */
        LzbShaJob *d_jobs = reinterpret_cast<LzbShaJob *>(s.d_packed);
        ok = lzb_cuda_ok(cudaMemcpyAsync(d_jobs, jobs.data(), job_bytes, cudaMemcpyHostToDevice, s.stream), "H2D sha jobs");
        if (!ok) return SZ_ERROR_FAIL;
        rc = lzb_launch_sha256(s.d_out, d_jobs, (uint32_t)jobs.size(), s.stream);
        if (rc != 0) { lzb_set_error("sha256 launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return SZ_ERROR_FAIL; }
        ok = lzb_cuda_ok(cudaMemcpyAsync(jobs.data(), d_jobs, job_bytes, cudaMemcpyDeviceToHost, s.stream), "D2H sha jobs");
        if (!ok) return SZ_ERROR_FAIL;
      }
      for (size_t i = 0; i < live.size(); i++)
        if (rel[i].filter != LZB_FILTER_NONE)
          ok = ok && lzb_cuda_ok(cudaMemcpyAsync(dest + live[i].u.dst_off, s.d_out + rel[i].dst_off, rel[i].dst_cap, cudaMemcpyDeviceToHost, s.stream), "D2H blocks");
      ok = ok && lzb_cuda_ok(cudaMemcpyAsync(pres.data(), s.d_results, pres.size() * sizeof(LzbResult), cudaMemcpyDeviceToHost, s.stream), "D2H results") &&
           lzb_cuda_ok(cudaStreamSynchronize(s.stream), "post kernel");
      if (!ok) return SZ_ERROR_FAIL;
      for (size_t i = 0; i < live.size(); i++)
      {
        if (pres[i].res != SZ_OK) clean[live[i].block] = 0;
        if (rel[i].check != LZB_CHECK_NONE) digest[live[i].block] = pres[i].check;
      }
    }
  }
  for (size_t k = 0; k < nb; k++)
  {
    if (!clean[k]) continue;
    const LzbXzBlock &b = w.blocks[k];
    XzReady q;
    memset(&q, 0, sizeof q);
    q.data_off = b.data_off; q.pack = b.pack_size; q.unpack = b.unpack_size;
    q.dst_off = w.units[b.first_unit].dst_off;
    q.check_type = b.check_type; q.n_filters = b.filter == LZB_FILTER_NONE ? 1 : 2;
    q.digest = digest[k];
    for (size_t i = 0; i < jobs.size(); i++) if (job_block[i] == k) memcpy(q.sha, jobs[i].digest, 32);
    ready.push_back(q);
  }
  return SZ_OK;
}

/* ------------------------------------------------------------------ xz backward index --------- */
extern "C" SRes Xz_ReadIndexBackward(const Byte *src, size_t srcLen, LzbXzBlock *blocks, size_t maxBlocks, size_t *n,
                                     int *checkType)
{
  /* reference XzIn.c:141-224 (footer, locate) and :71-114 (index records) */
  size_t end = srcLen;
  *n = 0; *checkType = 0;
  if ((end & 3) != 0 || end < 12) return SZ_ERROR_NO_ARCHIVE;
  if (memcmp(src + end - 2, "YZ", 2) != 0)
  {
    while (end > 0 && src[end - 1] == 0) end--;
    if ((end & 3) != 0 || end < 12 || memcmp(src + end - 2, "YZ", 2) != 0) return SZ_ERROR_NO_ARCHIVE;
  }
  const Byte *foot = src + end - 12;
  const unsigned flags = ((unsigned)foot[8] << 8) | foot[9];
  if (flags > 0xF) return SZ_ERROR_UNSUPPORTED;
  if (lzb_le32(foot) != lzb_small_crc32(foot + 4, 6)) return SZ_ERROR_ARCHIVE;
  const size_t idx_size = ((size_t)lzb_le32(foot + 4) + 1) << 2;
  if (idx_size + 12 > end) return SZ_ERROR_NO_ARCHIVE;
  const size_t idx_off = end - 12 - idx_size;
  const Byte *idx = src + idx_off;
  if (idx_size < 5 || idx[0] != 0) return SZ_ERROR_ARCHIVE;
  if (lzb_small_crc32(idx, idx_size - 4) != lzb_le32(idx + idx_size - 4)) return SZ_ERROR_ARCHIVE;
  size_t pos = 1;
  UInt64 v;
  unsigned s = Xz_ReadVarInt(idx + pos, idx_size - 4 - pos, &v);
  if (!s) return SZ_ERROR_ARCHIVE;
  pos += s;
  const size_t nb = (size_t)v;
  if (nb != v || nb * 2 > idx_size - 4) return SZ_ERROR_ARCHIVE;
  const unsigned csz = XzFlags_GetCheckSize((CXzStreamFlags)flags);
  std::vector<UInt64> total, unpack;
  try { total.resize(nb); unpack.resize(nb); } catch (const std::bad_alloc &) { return SZ_ERROR_MEM; }
  uint64_t data_total = 0;
  for (size_t i = 0; i < nb; i++)
  {
    s = Xz_ReadVarInt(idx + pos, idx_size - 4 - pos, &total[i]);
    if (!s) return SZ_ERROR_ARCHIVE;
    pos += s;
    s = Xz_ReadVarInt(idx + pos, idx_size - 4 - pos, &unpack[i]);
    if (!s) return SZ_ERROR_ARCHIVE;
    pos += s;
    if (total[i] == 0) return SZ_ERROR_ARCHIVE;
    /* the blocks lie in front of the index inside this buffer: a size that does not fit there is a broken index
     * (and must not be allowed to wrap the running sum) */
    if (total[i] > idx_off || data_total > idx_off - total[i]) return SZ_ERROR_ARCHIVE;
    data_total += (total[i] + 3) & ~(UInt64)3;
  }
  while ((pos & 3) != 0) if (idx[pos++] != 0) return SZ_ERROR_ARCHIVE;
  if (pos != idx_size - 4) return SZ_ERROR_ARCHIVE;
  if (12 + data_total > idx_off) return SZ_ERROR_ARCHIVE;
  size_t cur = idx_off - (size_t)data_total - 12;
  if (memcmp(src + cur, k_xz_magic, 6) != 0) return SZ_ERROR_NO_ARCHIVE;
  CXzStreamFlags hflags;
  SRes r = Xz_ParseHeader(&hflags, src + cur);
  if (r != SZ_OK) return r;
  if (hflags != flags) return SZ_ERROR_ARCHIVE;
  *checkType = (int)(flags & 0xF);
  *n = nb;
  cur += 12;
  for (size_t i = 0; i < nb && i < maxBlocks; i++)
  {
    LzbXzBlock &b = blocks[i];
    memset(&b, 0, sizeof b);
    if (cur >= idx_off || src[cur] == 0) return SZ_ERROR_ARCHIVE;
    const UInt32 hs = ((UInt32)src[cur] << 2) + 4;
    if (hs + csz > total[i] || hs > idx_off - cur) return SZ_ERROR_ARCHIVE;
    CXzBlock blk;
    memset(&blk, 0, sizeof blk);
    r = XzBlock_Parse(&blk, src + cur);
    if (r != SZ_OK) return r;
    r = lzb_xz_classify_chain(&blk, &b);
    if (r != SZ_OK) return r;
    b.hdr_off = cur; b.hdr_size = hs; b.data_off = cur + hs;
    b.pack_size = total[i] - hs - csz; b.unpack_size = unpack[i];
    b.check_off = b.data_off + ((b.pack_size + 3) & ~(uint64_t)3);
    b.check_type = (uint8_t)(flags & 0xF);
    b.expect = csz == 4 ? lzb_le32(src + b.check_off) : (csz == 8 ? lzb_le64(src + b.check_off) : 0);
    b.complete = 1;
    cur += (size_t)((total[i] + 3) & ~(UInt64)3);
  }
  return SZ_OK;
}
