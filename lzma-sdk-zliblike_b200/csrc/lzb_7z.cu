/* lzb_7z.cu -- the 7z folder path: archive database on the host, every folder of the
 * archive decoded in ONE batched GPU pass.
 *
 *   SzArchive_Open        reference SzArEx_Open   (7zIn.c:1214-1320): signature header, next header (plain or
 *                         LZMA-packed: kEncodedHeader, decoded on the GPU), PackInfo / UnpackInfo /
 *                         SubStreamsInfo / FilesInfo (7zIn.c:346-1110), SzArEx_Fill (7zIn.c:175-243)
 *   SzArchive_ExtractAll  reference SzArEx_Extract (7zIn.c:1322-1405) for every file, over SzFolder_Decode
 *                         (7zDec.c:269-472): folders are independent units -- Copy / LZMA / LZMA2, optionally
 *                         followed by BCJ x86 or ARM -- so all of them go to the GPU at once, with the branch
 *                         filter and the folder CRC fused after the decode and the per-file CRCs in a second
 *                         pass over the resident output.
 * The result codes are the reference's: what SzArEx_Open returns for a damaged database, and per file what
 * SzArEx_Extract returns (folder decode error, folder CRC, file CRC).  BCJ2 folders (7zDec.c:308-329: three
 * coders feeding a four-stream merge) decode their three streams as ordinary units and merge on the device
 * (lzb_bcj2_kernel, one warp per folder).  PPMd (disabled in the reference) is SZ_ERROR_UNSUPPORTED.
 */
#include "lzb_host.h"
#include "lzb_xz.h"

#include <string.h>

#include <memory>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

enum
{
  kEnd, kHeader, kArchiveProperties, kAdditionalStreamsInfo, kMainStreamsInfo, kFilesInfo, kPackInfo, kUnpackInfo,
  kSubStreamsInfo, kSize, kCRC, kFolder, kCodersUnpackSize, kNumUnpackStream, kEmptyStream, kEmptyFile, kAnti, kName,
  kCTime, kATime, kMTime, kWinAttributes, kComment, kEncodedHeader
};

const uint64_t kMethodCopy = 0, kMethodLzma2 = 0x21, kMethodLzma = 0x30101, kMethodBcj = 0x03030103, kMethodArm = 0x03030501,
               kMethodBcj2 = 0x0303011B;

/* bounds-checked cursor over header bytes; the first failure sticks (7zIn.c:316-400) */
struct Cursor
{
  const Byte *p;
  size_t left;
  SRes err = SZ_OK;

  bool fail(SRes e) { if (err == SZ_OK) err = e; return false; }
  bool byte(Byte &b)
  {
    if (err != SZ_OK) return false;
    if (left == 0) return fail(SZ_ERROR_ARCHIVE);
    b = *p++; left--;
    return true;
  }
  bool u32(UInt32 &v)
  {
    v = 0;
    for (int i = 0; i < 4; i++) { Byte b; if (!byte(b)) return false; v |= (UInt32)b << (8 * i); }
    return true;
  }
  /* first byte: k leading one bits = k more bytes, its remaining low bits are the top of the value */
  bool number(UInt64 &v)
  {
    Byte first;
    v = 0;
    if (!byte(first)) return false;
    Byte mask = 0x80;
    for (int i = 0; i < 8; i++, mask >>= 1)
    {
      if ((first & mask) == 0) { v += (UInt64)(first & (mask - 1)) << (8 * i); return true; }
      Byte b;
      if (!byte(b)) return false;
      v |= (UInt64)b << (8 * i);
    }
    return true;
  }
  bool number32(UInt32 &v)
  {
    UInt64 t;
    if (!number(t)) return false;
    if (t >= 0x80000000u) return fail(SZ_ERROR_UNSUPPORTED);
    v = (UInt32)t;
    return true;
  }
  bool skip(UInt64 n)
  {
    if (err != SZ_OK) return false;
    if (n > left) return fail(SZ_ERROR_ARCHIVE);
    p += (size_t)n; left -= (size_t)n;
    return true;
  }
  bool skip_sized() { UInt64 n; return number(n) && skip(n); }
  /* read ids until `want`; an End first is a broken database */
  bool wait_for(UInt64 want)
  {
    for (;;)
    {
      UInt64 t;
      if (!number(t)) return false;
      if (t == want) return true;
      if (t == kEnd) return fail(SZ_ERROR_ARCHIVE);
      if (!skip_sized()) return false;
    }
  }
  bool no_external() { Byte b; if (!byte(b)) return false; return b == 0 ? true : fail(SZ_ERROR_UNSUPPORTED); }
  bool bit_vector(size_t n, std::vector<Byte> &v)
  {
    v.assign(n, 0);
    Byte cur = 0, mask = 0;
    for (size_t i = 0; i < n; i++)
    {
      if (mask == 0) { if (!byte(cur)) return false; mask = 0x80; }
      v[i] = (cur & mask) ? 1 : 0;
      mask >>= 1;
    }
    return true;
  }
  bool bit_vector_or_all(size_t n, std::vector<Byte> &v)
  {
    Byte all;
    if (!byte(all)) return false;
    if (all == 0) return bit_vector(n, v);
    v.assign(n, 1);
    return true;
  }
  bool digests(size_t n, std::vector<Byte> &defined, std::vector<UInt32> &crc)
  {
    if (!bit_vector_or_all(n, defined)) return false;
    crc.assign(n, 0);
    for (size_t i = 0; i < n; i++) if (defined[i] && !u32(crc[i])) return false;
    return true;
  }
};

struct Coder { UInt64 method = 0; UInt32 n_in = 1, n_out = 1; std::vector<Byte> props; };
struct BindPair { UInt32 in, out; };

struct Folder
{
  std::vector<Coder> coders;
  std::vector<BindPair> binds;
  std::vector<UInt32> pack_streams;
  std::vector<UInt64> unpack_sizes;
  bool crc_defined = false;
  UInt32 crc = 0;
  UInt32 n_streams = 1;                 /* files stored in it (NumUnpackStreams) */

  UInt32 out_streams() const { UInt32 n = 0; for (auto &c : coders) n += c.n_out; return n; }
  bool out_is_bound(UInt32 i) const { for (auto &b : binds) if (b.out == i) return true; return false; }
  bool in_is_bound(UInt32 i) const { for (auto &b : binds) if (b.in == i) return true; return false; }
  /* the one output nobody consumes is the folder's content (7zIn.c:83-93) */
  UInt64 unpack_size() const
  {
    for (int i = (int)out_streams() - 1; i >= 0; i--) if (!out_is_bound((UInt32)i)) return unpack_sizes[(size_t)i];
    return 0;
  }
};

struct Database
{
  std::vector<UInt64> pack_sizes;
  std::vector<Folder> folders;
  UInt64 data_pos = 0;
};

/* 7zIn.c:528-636 */
bool read_folder(Cursor &c, Folder &f)
{
  UInt32 n_coders;
  if (!c.number32(n_coders)) return false;
  if (n_coders > 32) return c.fail(SZ_ERROR_UNSUPPORTED);
  f.coders.resize(n_coders);
  UInt32 total_in = 0, total_out = 0;
  for (auto &cd : f.coders)
  {
    Byte head;
    if (!c.byte(head)) return false;
    const unsigned id_size = head & 0xF;
    Byte id[15];
    for (unsigned j = 0; j < id_size; j++) if (!c.byte(id[j])) return false;
    if (id_size > 8) return c.fail(SZ_ERROR_UNSUPPORTED);
    for (unsigned j = 0; j < id_size; j++) cd.method |= (UInt64)id[id_size - 1 - j] << (8 * j);
    if (head & 0x10)
    {
      if (!c.number32(cd.n_in) || !c.number32(cd.n_out)) return false;
      if (cd.n_in > 32 || cd.n_out > 32) return c.fail(SZ_ERROR_UNSUPPORTED);
    }
    if (head & 0x20)
    {
      UInt64 n;
      if (!c.number(n)) return false;
      if (n > c.left) return c.fail(SZ_ERROR_ARCHIVE);
      cd.props.assign(c.p, c.p + (size_t)n);
      c.skip(n);
    }
    while (head & 0x80)                                  /* alternative methods: skipped */
    {
      if (!c.byte(head) || !c.skip(head & 0xF)) return false;
      if (head & 0x10) { UInt32 t; if (!c.number32(t) || !c.number32(t)) return false; }
      if (head & 0x20) { if (!c.skip_sized()) return false; }
    }
    total_in += cd.n_in; total_out += cd.n_out;
  }
  if (total_out == 0) return c.fail(SZ_ERROR_UNSUPPORTED);
  f.binds.resize(total_out - 1);
  for (auto &b : f.binds) if (!c.number32(b.in) || !c.number32(b.out)) return false;
  if (total_in < f.binds.size()) return c.fail(SZ_ERROR_UNSUPPORTED);
  const UInt32 n_pack = total_in - (UInt32)f.binds.size();
  f.pack_streams.resize(n_pack);
  if (n_pack == 1)
  {
    UInt32 i = 0;
    while (i < total_in && f.in_is_bound(i)) i++;
    if (i == total_in) return c.fail(SZ_ERROR_UNSUPPORTED);
    f.pack_streams[0] = i;
  }
  else
    for (auto &s : f.pack_streams) if (!c.number32(s)) return false;
  return true;
}

/* 7zIn.c:860-898 with its three sections (:478-527, :638-712, :714-856) */
bool read_streams_info(Cursor &c, Database &db, UInt64 &data_offset, UInt32 &n_sub, std::vector<UInt64> &sub_sizes,
                       std::vector<Byte> &sub_crc_defined, std::vector<UInt32> &sub_crc)
{
  for (;;)
  {
    UInt64 type;
    if (!c.number(type)) return false;
    if (type > 0x7FFFFFFF) return c.fail(SZ_ERROR_UNSUPPORTED);
    if (type == kEnd) return true;
    if (type == kPackInfo)
    {
      UInt32 n;
      if (!c.number(data_offset) || !c.number32(n) || !c.wait_for(kSize)) return false;
      if (n > c.left) return c.fail(SZ_ERROR_ARCHIVE);          /* every size takes at least one byte */
      db.pack_sizes.resize(n);
      for (auto &s : db.pack_sizes) if (!c.number(s)) return false;
      for (;;)
      {
        UInt64 t;
        if (!c.number(t)) return false;
        if (t == kEnd) break;
        if (t == kCRC) { std::vector<Byte> d; std::vector<UInt32> v; if (!c.digests(n, d, v)) return false; continue; }
        if (!c.skip_sized()) return false;
      }
    }
    else if (type == kUnpackInfo)
    {
      UInt32 n;
      if (!c.wait_for(kFolder) || !c.number32(n) || !c.no_external()) return false;
      if (n > c.left) return c.fail(SZ_ERROR_ARCHIVE);          /* every folder record takes at least one byte */
      db.folders.assign(n, Folder());
      for (auto &f : db.folders) if (!read_folder(c, f)) return false;
      if (!c.wait_for(kCodersUnpackSize)) return false;
      for (auto &f : db.folders)
      {
        f.unpack_sizes.resize(f.out_streams());
        for (auto &s : f.unpack_sizes) if (!c.number(s)) return false;
      }
      for (;;)
      {
        UInt64 t;
        if (!c.number(t)) return false;
        if (t == kEnd) break;
        if (t == kCRC)
        {
          std::vector<Byte> d; std::vector<UInt32> v;
          if (!c.digests(n, d, v)) return false;
          for (UInt32 i = 0; i < n; i++) { db.folders[i].crc_defined = d[i] != 0; db.folders[i].crc = v[i]; }
          continue;
        }
        if (!c.skip_sized()) return false;
      }
    }
    else if (type == kSubStreamsInfo)
    {
      UInt64 t = 0;
      for (auto &f : db.folders) f.n_streams = 1;
      n_sub = (UInt32)db.folders.size();
      for (;;)
      {
        if (!c.number(t)) return false;
        if (t == kNumUnpackStream)
        {
          n_sub = 0;
          for (auto &f : db.folders)
          {
            if (!c.number32(f.n_streams)) return false;
            if (f.n_streams > 0x7FFFFFFFu - n_sub) return c.fail(SZ_ERROR_UNSUPPORTED);
            n_sub += f.n_streams;
          }
          continue;
        }
        if (t == kCRC || t == kSize || t == kEnd) break;
        if (!c.skip_sized()) return false;
      }
      sub_sizes.clear();
      for (auto &f : db.folders)
      {
        if (f.n_streams == 0) continue;
        UInt64 sum = 0;
        if (t == kSize)
          for (UInt32 j = 1; j < f.n_streams; j++)
          {
            UInt64 s;
            if (!c.number(s)) return false;
            if (s > ~sum) return c.fail(SZ_ERROR_ARCHIVE);         /* the sizes of one folder must add up without wrapping */
            sub_sizes.push_back(s); sum += s;
          }
        sub_sizes.push_back(f.unpack_size() - sum);      /* the last one is what is left */
      }
      if (t == kSize && !c.number(t)) return false;
      if (t != kSize && n_sub > db.folders.size() + c.left) return c.fail(SZ_ERROR_ARCHIVE);
      sub_crc_defined.assign(n_sub, 0);
      sub_crc.assign(n_sub, 0);
      size_t listed = 0;                                  /* streams whose digest is not the folder's own */
      for (auto &f : db.folders) if (f.n_streams != 1 || !f.crc_defined) listed += f.n_streams;
      for (;;)
      {
        if (t == kCRC)
        {
          std::vector<Byte> d; std::vector<UInt32> v;
          if (!c.digests(listed, d, v)) return false;
          size_t si = 0, di = 0;
          for (auto &f : db.folders)
          {
            if (f.n_streams == 1 && f.crc_defined) { sub_crc_defined[si] = 1; sub_crc[si] = f.crc; si++; }
            else for (UInt32 j = 0; j < f.n_streams; j++, di++, si++) { sub_crc_defined[si] = d[di]; sub_crc[si] = v[di]; }
          }
        }
        else if (t == kEnd) break;
        else if (!c.skip_sized()) return false;
        if (!c.number(t)) return false;
      }
    }
    else return c.fail(SZ_ERROR_UNSUPPORTED);
  }
}

/* 7zDec.c:269-331 */
SRes classify_folder(const Folder &f, LzbSzFolder &o)
{
  auto main_ok = [](const Coder &c) {
    return c.n_in == 1 && c.n_out == 1 && (c.method == kMethodCopy || c.method == kMethodLzma || c.method == kMethodLzma2);
  };
  const size_t n = f.coders.size();
  if (n < 1 || n > 4 || !main_ok(f.coders[0])) return SZ_ERROR_UNSUPPORTED;
  const Coder &m = f.coders[0];
  o.method = m.method == kMethodCopy ? 0u : (m.method == kMethodLzma ? 1u : 2u);
  o.props_size = (uint8_t)(m.props.size() > 255 ? 255 : m.props.size());
  memset(o.props, 0, sizeof o.props);
  if (!m.props.empty()) memcpy(o.props, m.props.data(), m.props.size() < sizeof o.props ? m.props.size() : sizeof o.props);
  o.filter = LZB_FILTER_NONE;
  if (n == 1)
    return (f.pack_streams.size() == 1 && f.pack_streams[0] == 0 && f.binds.empty()) ? SZ_OK : SZ_ERROR_UNSUPPORTED;
  if (n == 2)
  {
    const Coder &c = f.coders[1];
    if (c.method > 0xFFFFFFFFull || c.n_in != 1 || c.n_out != 1 || f.pack_streams.size() != 1 || f.pack_streams[0] != 0 ||
        f.binds.size() != 1 || f.binds[0].in != 1 || f.binds[0].out != 0)
      return SZ_ERROR_UNSUPPORTED;
    if (c.method == kMethodBcj) o.filter = LZB_FILTER_X86;
    else if (c.method == kMethodArm) o.filter = LZB_FILTER_ARM;
    else return SZ_ERROR_UNSUPPORTED;
    return SZ_OK;
  }
  if (n == 4)
  {
    const Coder &b = f.coders[3];
    if (!main_ok(f.coders[1]) || !main_ok(f.coders[2]) || b.method != kMethodBcj2 || b.n_in != 4 || b.n_out != 1) return SZ_ERROR_UNSUPPORTED;
    static const UInt32 want_pack[4] = { 2, 6, 1, 0 }, want_in[3] = { 5, 4, 3 };
    if (f.pack_streams.size() != 4 || f.binds.size() != 3) return SZ_ERROR_UNSUPPORTED;
    for (int i = 0; i < 4; i++) if (f.pack_streams[i] != want_pack[i]) return SZ_ERROR_UNSUPPORTED;
    for (UInt32 i = 0; i < 3; i++) if (f.binds[i].in != want_in[i] || f.binds[i].out != i) return SZ_ERROR_UNSUPPORTED;
    o.method = 3;                                          /* three coders + BCJ2 merge */
    o.filter = 0x1B;
    return SZ_OK;
  }
  return SZ_ERROR_UNSUPPORTED;
}

}  // namespace

/* the three coders in front of a BCJ2 merge (7zDec.c:308-329): [0] -> jump stream, [1] -> call stream,
 * [2] -> main stream; plus the raw range-coder stream */
struct SubCoder { uint32_t method = 0; Byte props[8] = { 0 }; uint8_t props_size = 0; UInt64 pack_off = 0, pack_size = 0, unpack_size = 0; };
struct Bcj2Folder { SubCoder sub[3]; UInt64 rc_off = 0, rc_size = 0; };

struct LzbSzArchive
{
  std::vector<LzbSzFolder> folders;
  std::vector<LzbSzFile> files;
  std::vector<UInt16> names;
  std::vector<Bcj2Folder> bcj2;            /* for folders with method == 3, in folder order */
  std::vector<int> bcj2_index;             /* folder -> index in bcj2, or -1 */
  UInt64 unpack_total = 0;
};

/* one folder of the database -> the unit the kernel decodes (SzDecodeLzma / SzDecodeLzma2, 7zDec.c:126-216) */
static SRes coder_unit(uint32_t method, const Byte *props, unsigned props_size, UInt64 pack_off, UInt64 pack_size, UInt64 out_off,
                       UInt64 out_size, size_t archive_size, LzbUnit &u)
{
  memset(&u, 0, sizeof u);
  if (out_size > 0xFFFFFFFFull || pack_size > 0xFFFFFFFFull) { lzb_set_error("7z folder larger than 4 GiB"); return SZ_ERROR_UNSUPPORTED; }
  const UInt64 avail = pack_off >= archive_size ? 0 : std::min<UInt64>(pack_size, archive_size - pack_off);
  u.src_off = pack_off; u.src_len = (uint32_t)avail;
  u.dst_off = out_off; u.dst_cap = (uint32_t)out_size;
  u.finish = LZMA_FINISH_END;
  if (method == 1)
  {
    CLzmaProps p;
    const SRes r = LzmaProps_Decode(&p, props, props_size);
    if (r != SZ_OK) return r;
    u.codec = LZB_CODEC_LZMA; u.lc = (uint8_t)p.lc; u.lp = (uint8_t)p.lp; u.pb = (uint8_t)p.pb; u.dict_size = p.dicSize;
  }
  else
  {
    if (props_size != 1) return SZ_ERROR_DATA;
    if (props[0] > 40) return SZ_ERROR_UNSUPPORTED;
    u.codec = LZB_CODEC_LZMA2;
    u.dict_size = props[0] == 40 ? 0xFFFFFFFFu : ((UInt32)(2 | (props[0] & 1)) << (props[0] / 2 + 11));
  }
  return SZ_OK;
}

static SRes folder_unit(const LzbSzFolder &f, size_t archive_size, LzbUnit &u)
{
  return coder_unit(f.method, f.props, f.props_size, f.pack_off, f.pack_size, f.out_off, f.unpack_size, archive_size, u);
}

/* what the reference's feed loop makes of a finished decode: it hands the decoder at most 256 KiB per call,
 * and accepts the end only if the call that completed the output also used up the window it was given */
static SRes coder_verdict(uint32_t method, UInt64 pack_size, const LzbUnit &u, const LzbResult &x);
static SRes folder_verdict(const LzbSzFolder &f, const LzbUnit &u, const LzbResult &x) { return coder_verdict(f.method, f.pack_size, u, x); }

static SRes coder_verdict(uint32_t method, UInt64 pack_size, const LzbUnit &u, const LzbResult &x)
{
  /* decode verdict first, the folder CRC (fused into the same launch) only counts for a clean decode */
  if (x.res != SZ_OK && x.res != SZ_ERROR_CRC) return SZ_ERROR_DATA;
  if (x.dst_done != u.dst_cap) return SZ_ERROR_DATA;
  const bool ended = x.status == LZMA_STATUS_FINISHED_WITH_MARK || (method == 1 && x.status == LZMA_STATUS_MAYBE_FINISHED_WITHOUT_MARK);
  if (!ended) return SZ_ERROR_DATA;
  const UInt64 used = x.src_used, window = (UInt64)1 << 18;
  const UInt64 last_start = used == 0 ? 0 : ((used - 1) / window) * window;
  const UInt64 last_end = std::min<UInt64>(last_start + window, pack_size);
  return used == last_end ? x.res : SZ_ERROR_DATA;
}

/* SzArEx_Extract, 7zIn.c:1390-1391: `offset + size > unpackSize` is SZ_ERROR_FAIL -- asked BEFORE anything is
 * computed on the file's bytes, and without wrapping */
static bool file_inside_folder(const LzbSzFile &fi, const LzbSzFolder &f)
{
  if (fi.out_off < f.out_off || fi.size > f.unpack_size) return false;
  return fi.out_off - f.out_off <= f.unpack_size - fi.size;
}

static SRes decode_folders(LzbContext *ctx, const LzbSzArchive &a, const Byte *archive, size_t size, Byte *out, size_t out_bytes,
                           SRes *folder_res, SRes *file_res)
{
  const size_t nf = a.folders.size();
  std::vector<LzbUnit> units, post_a, post_b;
  std::vector<size_t> unit_folder, a_folder, b_file;
  std::vector<int> unit_sub;                            /* -1: the folder's only coder; 0..2: coder in front of a BCJ2 merge */
  std::vector<LzbCopyRange> copies;
  std::vector<LzbBcj2Job> jobs;
  std::vector<size_t> job_folder;
  UInt64 scratch = (out_bytes + 15) & ~(UInt64)15;      /* call / jump streams of BCJ2 folders live behind the output */
  for (size_t i = 0; i < nf; i++)
  {
    const LzbSzFolder &f = a.folders[i];
    folder_res[i] = (SRes)f.supported;
    if (folder_res[i] != SZ_OK) continue;
    if (f.method == 3)
    {
      /* 7zDec.c:348-430 for the four-coder chain: coder 0 -> jump stream, 1 -> call stream (both scratch),
       * 2 -> main stream at the END of the folder's output, then the merge in place */
      if (i >= a.bcj2_index.size() || a.bcj2_index[i] < 0) { folder_res[i] = SZ_ERROR_UNSUPPORTED; continue; }
      const Bcj2Folder &b = a.bcj2[(size_t)a.bcj2_index[i]];
      if (b.sub[2].unpack_size > f.unpack_size) { folder_res[i] = SZ_ERROR_PARAM; continue; }
      UInt64 where[3];
      where[0] = scratch; scratch += (b.sub[0].unpack_size + 15) & ~(UInt64)15;
      where[1] = scratch; scratch += (b.sub[1].unpack_size + 15) & ~(UInt64)15;
      where[2] = f.out_off + (f.unpack_size - b.sub[2].unpack_size);
      SRes r = SZ_OK;
      const size_t units_before = units.size(), copies_before = copies.size();
      for (int ci = 0; ci < 3 && r == SZ_OK; ci++)
      {
        const SubCoder &sc = b.sub[ci];
        if (sc.method == 0)
        {
          if (sc.pack_size != sc.unpack_size) r = SZ_ERROR_DATA;
          else if (sc.pack_off > size || sc.pack_size > size - sc.pack_off) r = SZ_ERROR_INPUT_EOF;
          else copies.push_back({ sc.pack_off, where[ci], sc.pack_size });
        }
        else
        {
          LzbUnit u;
          r = coder_unit(sc.method, sc.props, sc.props_size, sc.pack_off, sc.pack_size, where[ci], sc.unpack_size, size, u);
          if (r == SZ_OK) { units.push_back(u); unit_folder.push_back(i); unit_sub.push_back(ci); }
        }
      }
      if (r == SZ_OK && (b.rc_off > size || b.rc_size > size - b.rc_off || b.rc_size > 0xFFFFFFFFull)) r = SZ_ERROR_INPUT_EOF;
      if (r == SZ_OK && f.unpack_size > 0xFFFFFFFFull) r = SZ_ERROR_UNSUPPORTED;
      if (r != SZ_OK)
      {
        units.resize(units_before); unit_folder.resize(units_before); unit_sub.resize(units_before); copies.resize(copies_before);
        folder_res[i] = r;
        continue;
      }
      LzbBcj2Job j;
      memset(&j, 0, sizeof j);
      j.main_off = where[2]; j.main_size = (uint32_t)b.sub[2].unpack_size;
      j.call_off = where[1]; j.call_size = (uint32_t)b.sub[1].unpack_size;
      j.jump_off = where[0]; j.jump_size = (uint32_t)b.sub[0].unpack_size;
      j.rc_off = b.rc_off; j.rc_size = (uint32_t)b.rc_size;
      j.out_off = f.out_off; j.out_size = (uint32_t)f.unpack_size;
      jobs.push_back(j); job_folder.push_back(i);
      if (f.crc_defined)
      {
        LzbUnit p;
        memset(&p, 0, sizeof p);
        p.dst_off = f.out_off; p.dst_cap = (uint32_t)f.unpack_size;
        p.check = LZB_CHECK_CRC32; p.flags |= LZB_UNIT_VERIFY; p.expect = f.crc;
        post_a.push_back(p); a_folder.push_back(i);
      }
      continue;
    }
    if (f.method == 0)
    {
      if (f.pack_size != f.unpack_size) { folder_res[i] = SZ_ERROR_DATA; continue; }
      if (f.pack_off > size || f.pack_size > size - f.pack_off) { folder_res[i] = SZ_ERROR_INPUT_EOF; continue; }
      if (f.unpack_size > 0xFFFFFFFFull) { lzb_set_error("7z folder larger than 4 GiB"); folder_res[i] = SZ_ERROR_UNSUPPORTED; continue; }
      copies.push_back({ f.pack_off, f.out_off, f.pack_size });
    }
    else
    {
      LzbUnit u;
      const SRes r = folder_unit(f, size, u);
      if (r != SZ_OK) { folder_res[i] = r; continue; }
      /* BCJ x86 and the folder CRC ride in the decode launch; ARM and stored folders take the folder pass */
      if (f.filter == LZB_FILTER_X86) u.filter = LZB_FILTER_X86;
      if (f.filter != LZB_FILTER_ARM && f.crc_defined) { u.check = LZB_CHECK_CRC32; u.flags |= LZB_UNIT_VERIFY; u.expect = f.crc; }
      units.push_back(u); unit_folder.push_back(i); unit_sub.push_back(-1);
    }
    if (f.method == 0 || f.filter == LZB_FILTER_ARM)
    {
      LzbUnit p;
      memset(&p, 0, sizeof p);
      p.dst_off = f.out_off; p.dst_cap = (uint32_t)f.unpack_size;
      p.filter = (uint8_t)f.filter;
      if (f.crc_defined) { p.check = LZB_CHECK_CRC32; p.flags |= LZB_UNIT_VERIFY; p.expect = f.crc; }
      if (p.filter != LZB_FILTER_NONE || p.check != LZB_CHECK_NONE) { post_a.push_back(p); a_folder.push_back(i); }
    }
  }
  for (size_t k = 0; k < a.files.size(); k++)
  {
    const LzbSzFile &fi = a.files[k];
    file_res[k] = SZ_OK;
    if (fi.folder == 0xFFFFFFFFu || !fi.has_stream || !fi.crc_defined || fi.size == 0) continue;
    if (folder_res[fi.folder] != SZ_OK) continue;
    if (!file_inside_folder(fi, a.folders[fi.folder])) continue;          /* SZ_ERROR_FAIL below; never reaches the GPU */
    LzbUnit p;
    memset(&p, 0, sizeof p);
    p.dst_off = fi.out_off; p.dst_cap = (uint32_t)fi.size;
    p.check = LZB_CHECK_CRC32; p.flags |= LZB_UNIT_VERIFY; p.expect = fi.crc;
    post_b.push_back(p); b_file.push_back(k);
  }
  std::vector<LzbResult> res(units.size()), res_a(post_a.size()), res_b(post_b.size());
  if (!units.empty() || !post_a.empty() || !post_b.empty() || !copies.empty() || !jobs.empty())
  {
    if (!ctx) ctx = lzb_default_context();
    if (!ctx) return SZ_ERROR_FAIL;
    /* nothing but decode units (+ fused BCJ / folder CRC) and file checks: the decode pass can be one streamed
     * launch, with the outputs going down while later folders still decode */
    bool streamed = false;
    if (copies.empty() && jobs.empty() && post_a.empty() && !units.empty())
    {
      const SRes r = lzb_flow_then_checks(ctx, archive, size, out, out_bytes, units.data(), res.data(), units.size(),
                                          post_b.data(), res_b.data(), post_b.size(), &streamed);
      if (r != SZ_OK) return r;
    }
    const SRes r = streamed ? SZ_OK :
                   lzb_run_stages(ctx, archive, size, out, out_bytes, (size_t)scratch, copies.data(), copies.size(),
                                  units.data(), res.data(), units.size(), jobs.data(), jobs.size(),
                                  post_a.data(), res_a.data(), post_a.size(), post_b.data(), res_b.data(), post_b.size());
    if (r != SZ_OK) return r;
  }
  for (size_t j = 0; j < units.size(); j++)
  {
    const size_t fi = unit_folder[j];
    if (unit_sub[j] < 0) { folder_res[fi] = folder_verdict(a.folders[fi], units[j], res[j]); continue; }
    /* coders of a BCJ2 chain run in order 0, 1, 2 in the reference: the first failure is the folder's */
    const SubCoder &sc = a.bcj2[(size_t)a.bcj2_index[fi]].sub[unit_sub[j]];
    const SRes v = coder_verdict(sc.method, sc.pack_size, units[j], res[j]);
    if (folder_res[fi] == SZ_OK && v != SZ_OK) folder_res[fi] = v;
  }
  for (size_t j = 0; j < jobs.size(); j++)
    if (folder_res[job_folder[j]] == SZ_OK && jobs[j].res != SZ_OK) folder_res[job_folder[j]] = (SRes)jobs[j].res;
  for (size_t j = 0; j < post_a.size(); j++)
    if (folder_res[a_folder[j]] == SZ_OK && res_a[j].res != SZ_OK) folder_res[a_folder[j]] = res_a[j].res;
  /* SzArEx_Extract, 7zIn.c:1383-1404: folder verdict first, then the file's place in it, then its own CRC */
  std::vector<char> crc_bad(a.files.size(), 0);
  for (size_t j = 0; j < post_b.size(); j++) if (res_b[j].res != SZ_OK) crc_bad[b_file[j]] = 1;
  for (size_t k = 0; k < a.files.size(); k++)
  {
    const LzbSzFile &fi = a.files[k];
    if (fi.folder == 0xFFFFFFFFu) { file_res[k] = SZ_OK; continue; }
    const LzbSzFolder &f = a.folders[fi.folder];
    if (folder_res[fi.folder] != SZ_OK) { file_res[k] = folder_res[fi.folder]; continue; }
    if (!file_inside_folder(fi, f)) { file_res[k] = SZ_ERROR_FAIL; continue; }
    if (fi.crc_defined && fi.size == 0 && fi.crc != 0) { file_res[k] = SZ_ERROR_CRC; continue; }     /* CrcCalc of nothing is 0 */
    file_res[k] = crc_bad[k] ? SZ_ERROR_CRC : SZ_OK;
  }
  return SZ_OK;
}

/* 7zIn.c:900-1110 (header) + 175-243 (fill) */
static SRes read_header(Cursor &c, LzbSzArchive &a, UInt64 start_after_header)
{
  Database db;
  UInt64 type;
  UInt32 n_sub = 0;
  std::vector<UInt64> sub_sizes;
  std::vector<Byte> sub_crc_defined;
  std::vector<UInt32> sub_crc;
  if (!c.number(type)) return c.err;
  if (type == kArchiveProperties)
  {
    for (;;)
    {
      UInt64 t;
      if (!c.number(t)) return c.err;
      if (t == kEnd) break;
      /* the reference ignores a failing skip here (7zIn.c:424) and runs into the error on its next read */
      Cursor probe = c;
      if (probe.skip_sized()) c = probe;
    }
    if (!c.number(type)) return c.err;
  }
  if (type == kMainStreamsInfo)
  {
    if (!read_streams_info(c, db, db.data_pos, n_sub, sub_sizes, sub_crc_defined, sub_crc)) return c.err;
    db.data_pos += start_after_header;
    if (sub_sizes.empty() && n_sub == 0)
      /* no SubStreamsInfo at all: one stream per folder (the reference dereferences a null table here) */
      for (auto &f : db.folders) { sub_sizes.push_back(f.unpack_size()); sub_crc_defined.push_back(f.crc_defined); sub_crc.push_back(f.crc); }
    if (!c.number(type)) return c.err;
  }
  if (type == kEnd) return SZ_OK;
  if (type != kFilesInfo) return SZ_ERROR_ARCHIVE;
  UInt32 n_files;
  if (!c.number32(n_files)) return c.err;
  /* a file costs the header at least one bit (its kEmptyStream flag) unless it owns one of the streams above */
  if (n_files > 8 * (UInt64)c.left + sub_sizes.size()) return SZ_ERROR_ARCHIVE;
  a.files.assign(n_files, LzbSzFile());
  for (auto &f : a.files) { memset(&f, 0, sizeof f); f.has_stream = 1; }
  std::vector<Byte> empty_stream, empty_file, defined;
  std::vector<size_t> name_off;
  UInt32 n_empty = 0;
  for (;;)
  {
    UInt64 t, sz;
    if (!c.number(t)) return c.err;
    if (t == kEnd) break;
    if (!c.number(sz)) return c.err;
    if (sz > c.left) return SZ_ERROR_ARCHIVE;
    if (t == kName)
    {
      if (!c.no_external()) return c.err;
      const size_t bytes = (size_t)sz - 1;
      if (bytes & 1) return SZ_ERROR_ARCHIVE;
      const size_t units = bytes >> 1;
      a.names.resize(units);
      for (size_t i = 0; i < units; i++) a.names[i] = (UInt16)(c.p[2 * i] | (c.p[2 * i + 1] << 8));
      name_off.assign(n_files + 1, 0);
      size_t pos = 0;
      for (UInt32 i = 0; i < n_files; i++)
      {
        name_off[i] = pos;
        for (;; pos++) { if (pos >= units) return SZ_ERROR_ARCHIVE; if (a.names[pos] == 0) break; }
        pos++;
      }
      name_off[n_files] = pos;
      if (pos != units) return SZ_ERROR_ARCHIVE;
      if (!c.skip(bytes)) return c.err;
    }
    else if (t == kEmptyStream)
    {
      if (!c.bit_vector(n_files, empty_stream)) return c.err;
      n_empty = 0;
      for (Byte b : empty_stream) n_empty += b;
    }
    else if (t == kEmptyFile) { if (!c.bit_vector(n_empty, empty_file)) return c.err; }
    else if (t == kWinAttributes || t == kMTime)
    {
      if (!c.bit_vector_or_all(n_files, defined) || !c.no_external()) return c.err;
      for (UInt32 i = 0; i < n_files; i++)
      {
        LzbSzFile &f = a.files[i];
        if (t == kWinAttributes) { f.attrib_defined = defined[i]; f.attrib = 0; if (defined[i] && !c.u32(f.attrib)) return c.err; }
        else { f.mtime_defined = defined[i]; f.mtime_lo = f.mtime_hi = 0; if (defined[i] && (!c.u32(f.mtime_lo) || !c.u32(f.mtime_hi))) return c.err; }
      }
    }
    else if (!c.skip(sz)) return c.err;
  }
  /* sizes / digests of the streams go to the files that have one, in order (7zIn.c:1078-1107) */
  size_t si = 0, ei = 0;
  for (UInt32 i = 0; i < n_files; i++)
  {
    LzbSzFile &f = a.files[i];
    f.has_stream = empty_stream.empty() ? 1 : (empty_stream[i] ? 0 : 1);
    if (f.has_stream)
    {
      if (si >= sub_sizes.size()) return SZ_ERROR_ARCHIVE;   /* (the reference reads past its arrays here) */
      f.is_dir = 0; f.size = sub_sizes[si]; f.crc = sub_crc[si]; f.crc_defined = sub_crc_defined[si];
      si++;
    }
    else
    {
      f.is_dir = ei < empty_file.size() ? (empty_file[ei] ? 0 : 1) : 1;
      ei++;
    }
    if (!name_off.empty()) { f.name_off = (uint32_t)name_off[i]; f.name_len = (uint32_t)(name_off[i + 1] - name_off[i]); }
  }
  /* folders: pack stream positions, output placement */
  a.folders.assign(db.folders.size(), LzbSzFolder());
  UInt64 pack_pos = 0, out_pos = 0;
  size_t pack_index = 0;
  for (size_t i = 0; i < db.folders.size(); i++)
  {
    const Folder &f = db.folders[i];
    LzbSzFolder &o = a.folders[i];
    memset(&o, 0, sizeof o);
    o.supported = (uint8_t)classify_folder(f, o);
    o.pack_off = db.data_pos + pack_pos;
    o.pack_size = pack_index < db.pack_sizes.size() ? db.pack_sizes[pack_index] : 0;
    a.bcj2_index.push_back(-1);
    if (o.supported == SZ_OK && o.method == 3 && pack_index + 4 <= db.pack_sizes.size())
    {
      /* pack streams of the folder, in order: [0] main (coder 2), [1] range coder (raw), [2] call (coder 1),
       * [3] jump (coder 0) -- 7zDec.c:358-362,408-411 */
      Bcj2Folder b;
      UInt64 off[4], at = o.pack_off;
      for (int k = 0; k < 4; k++) { off[k] = at; at += db.pack_sizes[pack_index + k]; }
      static const int stream_of[3] = { 3, 2, 0 };
      for (int ci = 0; ci < 3; ci++)
      {
        const Coder &cd = f.coders[(size_t)ci];
        SubCoder &sc = b.sub[ci];
        sc.method = cd.method == kMethodCopy ? 0u : (cd.method == kMethodLzma ? 1u : 2u);
        sc.props_size = (uint8_t)(cd.props.size() > 255 ? 255 : cd.props.size());
        if (!cd.props.empty()) memcpy(sc.props, cd.props.data(), cd.props.size() < sizeof sc.props ? cd.props.size() : sizeof sc.props);
        sc.pack_off = off[stream_of[ci]]; sc.pack_size = db.pack_sizes[pack_index + (size_t)stream_of[ci]];
        sc.unpack_size = f.unpack_sizes[(size_t)ci];
      }
      b.rc_off = off[1]; b.rc_size = db.pack_sizes[pack_index + 1];
      a.bcj2_index.back() = (int)a.bcj2.size();
      a.bcj2.push_back(b);
    }
    else if (o.method == 3 && o.supported == SZ_OK) o.supported = (uint8_t)SZ_ERROR_ARCHIVE;
    for (size_t k = 0; k < f.pack_streams.size() && pack_index + k < db.pack_sizes.size(); k++)
    {
      const UInt64 ps = db.pack_sizes[pack_index + k];
      if (ps > ~pack_pos || o.pack_off > ~(pack_pos + ps)) return SZ_ERROR_ARCHIVE;      /* positions must not wrap */
      pack_pos += ps;
    }
    pack_index += f.pack_streams.size();
    o.unpack_size = f.unpack_size();
    o.out_off = out_pos;
    if (o.unpack_size > ~out_pos) return SZ_ERROR_ARCHIVE;
    out_pos += o.unpack_size;
    o.crc = f.crc; o.crc_defined = f.crc_defined ? 1 : 0;
    o.num_streams = f.n_streams;
  }
  a.unpack_total = out_pos;
  /* files -> folders (SzArEx_Fill) */
  UInt32 folder = 0, in_folder = 0;
  UInt64 within = 0;
  for (UInt32 i = 0; i < n_files; i++)
  {
    LzbSzFile &f = a.files[i];
    const bool empty = !f.has_stream;
    if (empty && in_folder == 0) { f.folder = 0xFFFFFFFFu; continue; }
    if (in_folder == 0)
    {
      for (;; folder++)
      {
        if (folder >= a.folders.size()) return SZ_ERROR_ARCHIVE;
        a.folders[folder].first_file = i;
        if (a.folders[folder].num_streams != 0) break;
      }
      within = 0;
    }
    f.folder = folder;
    f.out_off = a.folders[folder].out_off + within;
    if (empty) continue;
    within += f.size;
    if (++in_folder >= a.folders[folder].num_streams) { folder++; in_folder = 0; }
  }
  return SZ_OK;
}

static SRes open_archive(LzbContext *ctx, const Byte *archive, size_t size, LzbSzArchive **out);

/* sizes in a header are the archive's word: an allocation they make fail is SZ_ERROR_MEM (as ISzAlloc returning 0
 * is in the reference), never an exception through the C ABI */
extern "C" SRes SzArchive_Open(LzbContext *ctx, const Byte *archive, size_t size, LzbSzArchive **out)
{
  try { return open_archive(ctx, archive, size, out); }
  catch (const std::bad_alloc &) { *out = nullptr; return SZ_ERROR_MEM; }
  catch (const std::length_error &) { *out = nullptr; return SZ_ERROR_MEM; }
}

static SRes open_archive(LzbContext *ctx, const Byte *archive, size_t size, LzbSzArchive **out)
{
  static const Byte k_sig[6] = { '7', 'z', 0xBC, 0xAF, 0x27, 0x1C };
  *out = nullptr;
  if (size < 32 || memcmp(archive, k_sig, 6) != 0) return SZ_ERROR_NO_ARCHIVE;
  if (archive[6] != 0) return SZ_ERROR_UNSUPPORTED;
  const UInt64 next_off = lzb_le64(archive + 12), next_size = lzb_le64(archive + 20);
  const UInt32 next_crc = lzb_le32(archive + 28);
  if (lzb_small_crc32(archive + 12, 20) != lzb_le32(archive + 8)) return SZ_ERROR_CRC;
  std::unique_ptr<LzbSzArchive> a(new LzbSzArchive());
  if (next_size == 0) { *out = a.release(); return SZ_OK; }
  SRes res = SZ_OK;
  if (next_off > next_off + next_size || next_off > next_off + next_size + 32) res = SZ_ERROR_NO_ARCHIVE;
  else if (size < next_off || size < 32 + next_off || size < 32 + next_off + next_size) res = SZ_ERROR_INPUT_EOF;
  std::vector<Byte> unpacked;
  if (res == SZ_OK)
  {
    const Byte *hdr = archive + 32 + next_off;
    res = SZ_ERROR_ARCHIVE;
    /* the database is container framing: its CRC is checked on the host like the xz headers */
    if (lzb_small_crc32(hdr, (size_t)next_size) == next_crc)
    {
      Cursor c { hdr, (size_t)next_size };
      UInt64 type = 0;
      res = c.number(type) ? SZ_OK : c.err;
      if (res == SZ_OK && type == kEncodedHeader)
      {
        /* 7zIn.c:1112-1192: one folder holds the real header */
        Database db;
        UInt32 n_sub = 0;
        std::vector<UInt64> ss; std::vector<Byte> sd; std::vector<UInt32> sc;
        UInt64 data_off = 0;
        if (!read_streams_info(c, db, data_off, n_sub, ss, sd, sc)) res = c.err;
        else if (db.folders.size() != 1) res = SZ_ERROR_ARCHIVE;
        else
        {
          LzbSzArchive tmp;
          LzbSzFolder f;
          memset(&f, 0, sizeof f);
          f.supported = (uint8_t)classify_folder(db.folders[0], f);
          f.pack_off = data_off + 32;
          f.pack_size = db.pack_sizes.empty() ? 0 : db.pack_sizes[0];
          f.unpack_size = db.folders[0].unpack_size();
          f.crc = db.folders[0].crc; f.crc_defined = db.folders[0].crc_defined;
          tmp.folders.push_back(f);
          if (f.unpack_size > 0xFFFFFFFFull) { res = SZ_ERROR_UNSUPPORTED; }
          else {
          unpacked.assign((size_t)f.unpack_size + 16, 0);
          SRes fr = SZ_OK;
          res = decode_folders(ctx, tmp, archive, size, unpacked.data(), (size_t)f.unpack_size, &fr, nullptr);
          if (res == SZ_OK) res = fr;
          if (res == SZ_OK)
          {
            c = Cursor { unpacked.data(), (size_t)f.unpack_size };
            res = c.number(type) ? SZ_OK : c.err;
          }
          }
        }
      }
      if (res == SZ_OK) res = type == kHeader ? read_header(c, *a, 32) : SZ_ERROR_UNSUPPORTED;
    }
  }
  if (res != SZ_OK) return res;
  *out = a.release();
  return SZ_OK;
}

extern "C" void SzArchive_Free(LzbSzArchive *a) { delete a; }
extern "C" size_t SzArchive_NumFolders(const LzbSzArchive *a) { return a->folders.size(); }
extern "C" size_t SzArchive_NumFiles(const LzbSzArchive *a) { return a->files.size(); }
extern "C" const LzbSzFolder *SzArchive_Folders(const LzbSzArchive *a) { return a->folders.data(); }
extern "C" const LzbSzFile *SzArchive_Files(const LzbSzArchive *a) { return a->files.data(); }
extern "C" const UInt16 *SzArchive_Names(const LzbSzArchive *a, size_t *n) { *n = a->names.size(); return a->names.data(); }
extern "C" UInt64 SzArchive_UnpackSize(const LzbSzArchive *a) { return a->unpack_total; }

extern "C" SRes SzArchive_ExtractAll(LzbContext *ctx, const LzbSzArchive *a, const Byte *archive, size_t size, Byte *out, size_t outCap,
                                     SRes *folderRes, SRes *fileRes)
{
  if (outCap < a->unpack_total) { lzb_set_error("SzArchive_ExtractAll: output smaller than SzArchive_UnpackSize"); return SZ_ERROR_PARAM; }
  try
  {
    std::vector<SRes> fr(a->folders.size() + 1), fi(a->files.size() + 1);
    const SRes r = decode_folders(ctx, *a, archive, size, out, (size_t)a->unpack_total, fr.data(), fi.data());
    if (r != SZ_OK) return r;
    if (folderRes) memcpy(folderRes, fr.data(), a->folders.size() * sizeof(SRes));
    if (fileRes) memcpy(fileRes, fi.data(), a->files.size() * sizeof(SRes));
    return SZ_OK;
  }
  catch (const std::bad_alloc &) { return SZ_ERROR_MEM; }
  catch (const std::length_error &) { return SZ_ERROR_MEM; }
}
