/* lzb_xzstream.cu -- the resumable xz layer of the C-ABI.
 *
 *   Sha256_*            Sha256.h:13-22        digest object; payload blocks are hashed by lzb_sha256_kernel
 *   XzCheck_*           Xz.c:45-88            block check object over CrcUpdate / Crc64Update / SHA-256 (GPU)
 *   MixCoder_*          XzDec.c:356-480       filter chain of a block, resumable at any byte
 *     LZMA2 stage       XzDec.c:276-316       = Lzma2Dec_DecodeToBuf (lzb_state.cu: device-resident decoder)
 *     BCJ/Delta stage   XzDec.c:48-232        16 KiB staging buffer around x86_Convert / Delta_Decode (GPU)
 *   XzDec_Init          XzDec.c:558-588
 *   XzUnpacker_*        XzDec.c:590-876       container state machine
 *
 * The container framing (headers, index, footer, padding: a few dozen bytes per block) is handled on the
 * host, one state per framing element like the reference, so a caller may cut the input and the output
 * anywhere.  Block PAYLOADS run on the GPU.  Two routes:
 *   fast  whole blocks present in `src`, with room for their output, are decoded ahead of the machine in one
 *         batched pass (lzb_xz_prefetch: every block and every dictionary-reset group in parallel, filter and
 *         digest on the GPU); when the machine reaches such a block it only does the bookkeeping;
 *   slow  a block cut by the end of `src` or `dest` goes through MixCoder_Code, call by call.
 * Both must yield the reference's (SRes, status, srcLen, destLen) at every call; the places where the
 * reference's loop structure is observable are marked "observable" below.
 */
#include "lzb_host.h"
#include "lzb_xz.h"
#include "lzb_sha256.cuh"

#include <string.h>

#include <algorithm>

#define LZB_CODER_BUF_SIZE ((size_t)1 << 17)     /* XzDec.c:28: stage-to-stage buffer */
#define LZB_BRA_BUF_SIZE ((size_t)1 << 14)       /* XzDec.c:50: filter staging buffer */
#define LZB_XZ_CHECK_SIZE_MAX 64

/* ------------------------------------------------------------------ SHA-256 object -------------- */
/* whole 64-byte blocks of payload through the GPU; state in, state out */
static bool gpu_sha_blocks(uint32_t state[8], const Byte *data, size_t bytes)
{
  LzbContext *c = lzb_default_context();
  if (!c) return false;
  std::lock_guard<std::mutex> g(c->mu);
  if (!lzb_cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return false;
  LzbSlot &s = c->slot[0];
  if (!lzb_slot_reserve(c, s, sizeof(LzbShaJob) + 16, bytes, 1, 0)) return false;
  LzbShaJob j;
  memset(&j, 0, sizeof j);
  j.len = bytes;
  memcpy(j.state, state, sizeof j.state);
  LzbShaJob *d_job = reinterpret_cast<LzbShaJob *>(s.d_packed);
  bool ok = lzb_cuda_ok(cudaMemcpyAsync(s.d_out, data, bytes, cudaMemcpyHostToDevice, s.stream), "H2D sha data") &&
            lzb_cuda_ok(cudaMemcpyAsync(d_job, &j, sizeof j, cudaMemcpyHostToDevice, s.stream), "H2D sha job");
  if (!ok) return false;
  int rc = lzb_launch_sha256(s.d_out, d_job, 1, s.stream);
  if (rc != 0) { lzb_set_error("sha256 launch failed: %s", cudaGetErrorString((cudaError_t)rc)); return false; }
  ok = lzb_cuda_ok(cudaMemcpyAsync(&j, d_job, sizeof j, cudaMemcpyDeviceToHost, s.stream), "D2H sha job") &&
       lzb_cuda_ok(cudaStreamSynchronize(s.stream), "sha256 kernel");
  if (ok) memcpy(state, j.state, sizeof j.state);
  return ok;
}

extern "C" void Sha256_Init(CSha256 *p)
{
  lzb::sha256_init(p->state);
  p->count = 0;
}

/* on_gpu: the bytes are payload (hashed on the device); otherwise container framing (a few bytes, host) */
static void sha_absorb(CSha256 *p, const Byte *data, size_t size, bool on_gpu)
{
  const unsigned fill = (unsigned)(p->count & 63);
  p->count += size;
  if (fill != 0)
  {
    const size_t take = std::min<size_t>(64 - fill, size);
    memcpy(p->buffer + fill, data, take);
    data += take; size -= take;
    if (fill + take < 64) return;
    lzb::sha256_block(p->state, p->buffer);
  }
  const size_t whole = size & ~(size_t)63;
  if (whole != 0)
  {
    if (on_gpu)
    {
      /* no host fallback for payload: without a device the state is poisoned, no digest will match */
      if (!gpu_sha_blocks(p->state, data, whole)) memset(p->state, 0, sizeof p->state);
    }
    else
      for (size_t k = 0; k < whole; k += 64) lzb::sha256_block(p->state, data + k);
    data += whole; size -= whole;
  }
  memcpy(p->buffer, data, size);
}

/* the exported object hashes its whole blocks on the device; only this library's own framing bytes (index
 * records inside XzUnpacker_Code) take the host routine */
extern "C" void Sha256_Update(CSha256 *p, const Byte *data, size_t size) { sha_absorb(p, data, size, true); }

extern "C" void Sha256_Final(CSha256 *p, Byte *digest)
{
  uint32_t s[8];
  memcpy(s, p->state, sizeof s);
  const unsigned rem = (unsigned)(p->count & 63);
  lzb::sha256_finish(s, p->buffer, rem, p->count - rem, digest);
  Sha256_Init(p);                                                    /* Sha256.c:203 */
}

/* ------------------------------------------------------------------ block check object ---------- */
#define LZB_CHECK_SHA_READY (XZ_CHECK_SHA256 | 0x100)   /* private: digest already final, kept in sha.buffer */

extern "C" void XzCheck_Init(CXzCheck *p, int mode)
{
  p->mode = mode;
  if (mode == XZ_CHECK_CRC32) p->crc = 0xFFFFFFFFu;
  else if (mode == XZ_CHECK_CRC64) p->crc64 = ~(UInt64)0;
  else if (mode == XZ_CHECK_SHA256) Sha256_Init(&p->sha);
}

extern "C" void XzCheck_Update(CXzCheck *p, const void *data, size_t size)
{
  if (size == 0) return;
  if (p->mode == XZ_CHECK_CRC32) p->crc = CrcUpdate(p->crc, data, size);
  else if (p->mode == XZ_CHECK_CRC64) p->crc64 = Crc64Update(p->crc64, data, size);
  else if (p->mode == XZ_CHECK_SHA256) sha_absorb(&p->sha, (const Byte *)data, size, true);
}

extern "C" int XzCheck_Final(CXzCheck *p, Byte *digest)
{
  if (p->mode == XZ_CHECK_CRC32)
  {
    const UInt32 v = p->crc ^ 0xFFFFFFFFu;
    for (int i = 0; i < 4; i++) digest[i] = (Byte)(v >> (8 * i));
    return 1;
  }
  if (p->mode == XZ_CHECK_CRC64)
  {
    const UInt64 v = ~p->crc64;
    for (int i = 0; i < 8; i++) digest[i] = (Byte)(v >> (8 * i));
    return 1;
  }
  if (p->mode == XZ_CHECK_SHA256) { Sha256_Final(&p->sha, digest); return 1; }
  if (p->mode == LZB_CHECK_SHA_READY) { memcpy(digest, p->sha.buffer, SHA256_DIGEST_SIZE); return 1; }
  return 0;                                                          /* unknown ids are not verified */
}

/* ------------------------------------------------------------------ BCJ / Delta stage ----------- */
/* Staging buffer: [0, handed) already given to the caller, [handed, converted) filtered and waiting,
 * [converted, held) raw bytes the filter could not finish yet (an x86 candidate needs 5 bytes). */
struct BraStage
{
  size_t handed, converted, held;
  UInt32 method;
  int encode;
  UInt32 delta, ip, x86_state;
  Byte delta_state[DELTA_STATE_SIZE];
  Byte buf[LZB_BRA_BUF_SIZE];
};

static void bra_free(void *pp, ISzAlloc *alloc) { alloc->Free(alloc, pp); }

static SRes bra_set_props(void *pp, const Byte *props, size_t propSize, ISzAlloc *alloc)
{
  /* XzDec.c:71-110 for the two methods this library runs */
  BraStage *b = (BraStage *)pp;
  (void)alloc;
  b->encode = 0;
  b->ip = 0;
  if (b->method == XZ_ID_Delta)
  {
    if (propSize != 1) return SZ_ERROR_UNSUPPORTED;
    b->delta = (UInt32)props[0] + 1;
    return SZ_OK;
  }
  if (propSize == 4)
  {
    const UInt32 v = lzb_le32(props);
    const UInt32 need = b->method == XZ_ID_X86 ? 0u : (b->method == XZ_ID_IA64 ? 15u : (b->method == XZ_ID_ARMT ? 1u : 3u));
    if (v & need) return SZ_ERROR_UNSUPPORTED;
    b->ip = v;
  }
  else if (propSize != 0) return SZ_ERROR_UNSUPPORTED;
  return SZ_OK;
}

static void bra_init(void *pp)
{
  BraStage *b = (BraStage *)pp;
  b->handed = b->converted = b->held = 0;
  b->x86_state = 0;
  if (b->method == XZ_ID_Delta) Delta_Init(b->delta_state);
}

/* XzDec.c:124-196 */
static SRes bra_code(void *pp, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, int srcWasFinished,
                     ECoderFinishMode finishMode, int *wasFinished)
{
  BraStage *b = (BraStage *)pp;
  size_t room = *destLen, avail = *srcLen;
  (void)finishMode;
  *destLen = 0; *srcLen = 0; *wasFinished = 0;
  while (room != 0)
  {
    if (b->handed != b->converted)
    {
      const size_t n = std::min(b->converted - b->handed, room);
      memcpy(dest, b->buf + b->handed, n);
      b->handed += n; dest += n; room -= n; *destLen += n;
      continue;
    }
    /* nothing converted is waiting: drop what was handed out, top the buffer up, filter again */
    b->held -= b->handed;
    memmove(b->buf, b->buf + b->handed, b->held);
    b->handed = b->converted = 0;
    {
      const size_t n = std::min(LZB_BRA_BUF_SIZE - b->held, avail);
      memcpy(b->buf + b->held, src, n);
      b->held += n; src += n; avail -= n; *srcLen += n;
    }
    if (b->held == 0) break;
    if (b->method == XZ_ID_Delta)
    {
      if (b->encode) Delta_Encode(b->delta_state, b->delta, b->buf, b->held);
      else Delta_Decode(b->delta_state, b->delta, b->buf, b->held);
      b->converted = b->held;
    }
    else if (b->method == XZ_ID_X86) b->converted = x86_Convert(b->buf, b->held, b->ip, &b->x86_state, b->encode);
    else if (b->method == XZ_ID_PPC) b->converted = PPC_Convert(b->buf, b->held, b->ip, b->encode);
    else if (b->method == XZ_ID_IA64) b->converted = IA64_Convert(b->buf, b->held, b->ip, b->encode);
    else if (b->method == XZ_ID_ARM) b->converted = ARM_Convert(b->buf, b->held, b->ip, b->encode);
    else if (b->method == XZ_ID_ARMT) b->converted = ARMT_Convert(b->buf, b->held, b->ip, b->encode);
    else if (b->method == XZ_ID_SPARC) b->converted = SPARC_Convert(b->buf, b->held, b->ip, b->encode);
    else
      return SZ_ERROR_UNSUPPORTED;
    b->ip += (UInt32)b->converted;
    if (b->converted == 0)
    {
      if (!srcWasFinished) break;          /* the tail may still grow into a convertible instruction */
      b->converted = b->held;              /* end of data: the last bytes pass unfiltered (XzDec.c:186-191) */
    }
  }
  if (b->held == b->handed && avail == 0 && srcWasFinished) *wasFinished = 1;
  return SZ_OK;
}

static SRes bra_attach(IStateCoder *sc, UInt64 id, ISzAlloc *alloc)
{
  /* XzDec.c:198-221 */
  if (id != XZ_ID_Delta && (id < XZ_ID_X86 || id > XZ_ID_SPARC)) return SZ_ERROR_UNSUPPORTED;
  sc->p = 0;
  BraStage *b = (BraStage *)alloc->Alloc(alloc, sizeof(BraStage));
  if (!b) return SZ_ERROR_MEM;
  b->method = (UInt32)id;
  sc->p = b;
  sc->Free = bra_free; sc->SetProps = bra_set_props; sc->Init = bra_init; sc->Code = bra_code;
  return SZ_OK;
}

/* ------------------------------------------------------------------ LZMA2 stage ----------------- */
static void l2_free(void *pp, ISzAlloc *alloc)
{
  Lzma2Dec_Free((CLzma2Dec *)pp, alloc);
  alloc->Free(alloc, pp);
}

static SRes l2_set_props(void *pp, const Byte *props, size_t propSize, ISzAlloc *alloc)
{
  if (propSize != 1) return SZ_ERROR_UNSUPPORTED;
  return Lzma2Dec_Allocate((CLzma2Dec *)pp, props[0], alloc);
}

static void l2_init(void *pp) { Lzma2Dec_Init((CLzma2Dec *)pp); }

static SRes l2_code(void *pp, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, int srcWasFinished,
                    ECoderFinishMode finishMode, int *wasFinished)
{
  ELzmaStatus st;
  (void)srcWasFinished;
  const SRes r = Lzma2Dec_DecodeToBuf((CLzma2Dec *)pp, dest, destLen, src, srcLen, (ELzmaFinishMode)finishMode, &st);
  *wasFinished = st == LZMA_STATUS_FINISHED_WITH_MARK;
  return r;
}

static SRes l2_attach(IStateCoder *sc, ISzAlloc *alloc)
{
  CLzma2Dec *d = (CLzma2Dec *)alloc->Alloc(alloc, sizeof(CLzma2Dec));
  sc->p = d;
  if (!d) return SZ_ERROR_MEM;
  sc->Free = l2_free; sc->SetProps = l2_set_props; sc->Init = l2_init; sc->Code = l2_code;
  Lzma2Dec_Construct(d);
  return SZ_OK;
}

/* ------------------------------------------------------------------ filter chain ---------------- */
extern "C" void MixCoder_Construct(CMixCoder *p, ISzAlloc *alloc)
{
  p->alloc = alloc;
  p->buf = 0;
  p->numCoders = 0;
  for (int i = 0; i < MIXCODER_NUM_FILTERS_MAX; i++) p->coders[i].p = NULL;
}

extern "C" void MixCoder_Free(CMixCoder *p)
{
  for (int i = 0; i < p->numCoders; i++)
  {
    IStateCoder *sc = &p->coders[i];
    if (p->alloc && sc->p) sc->Free(sc->p, p->alloc);
    sc->p = NULL;
  }
  p->numCoders = 0;
  if (p->buf) { p->alloc->Free(p->alloc, p->buf); p->buf = 0; }
}

extern "C" void MixCoder_Init(CMixCoder *p)
{
  for (int i = 0; i + 1 < p->numCoders; i++) { p->size[i] = 0; p->pos[i] = 0; p->finished[i] = 0; }
  for (int i = 0; i < p->numCoders; i++) p->coders[i].Init(p->coders[i].p);
}

extern "C" SRes MixCoder_SetFromMethod(CMixCoder *p, int coderIndex, UInt64 methodId)
{
  IStateCoder *sc = &p->coders[coderIndex];
  p->ids[coderIndex] = methodId;
  if (methodId == XZ_ID_LZMA2) return l2_attach(sc, p->alloc);
  if (coderIndex == 0) return SZ_ERROR_UNSUPPORTED;        /* a chain starts with a real decoder */
  return bra_attach(sc, methodId, p->alloc);
}

/* XzDec.c:375-480.  Stage i reads what stage i-1 left in its hand-over buffer and writes into its own
 * (or into dest for the last stage); rounds repeat until one makes no progress.
 * observable: `status` is FINISHED_WITH_MARK only when EVERY stage call made during this invocation said
 * "finished" -- a chain that finishes in a later round of the same invocation still reports NOT_FINISHED
 * and is seen finished by the next invocation. */
extern "C" SRes MixCoder_Code(CMixCoder *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, int srcWasFinished,
                              ECoderFinishMode finishMode, ECoderStatus *status)
{
  const SizeT dest_cap = *destLen, src_avail = *srcLen;
  bool all_said_finished = true;
  *destLen = 0; *srcLen = 0;
  *status = CODER_STATUS_NOT_FINISHED;
  if (!p->buf)
  {
    p->buf = (Byte *)p->alloc->Alloc(p->alloc, LZB_CODER_BUF_SIZE * (MIXCODER_NUM_FILTERS_MAX - 1));
    if (!p->buf) return SZ_ERROR_MEM;
  }
  if (p->numCoders != 1) finishMode = CODER_FINISH_ANY;
  const int last = p->numCoders - 1;
  for (bool progress = true; progress;)
  {
    progress = false;
    for (int i = 0; i <= last; i++)
    {
      const Byte *in;
      SizeT in_len, out_len;
      Byte *out;
      int in_finished, said_finished = 0;
      if (i == 0) { in = src; in_len = src_avail - *srcLen; in_finished = srcWasFinished; }
      else
      {
        in = p->buf + LZB_CODER_BUF_SIZE * (size_t)(i - 1) + p->pos[i - 1];
        in_len = p->size[i - 1] - p->pos[i - 1];
        in_finished = p->finished[i - 1];
      }
      if (i == last) { out = dest; out_len = dest_cap - *destLen; }
      else
      {
        if (p->pos[i] != p->size[i]) continue;             /* the next stage has not drained it yet */
        out = p->buf + LZB_CODER_BUF_SIZE * (size_t)i;
        out_len = LZB_CODER_BUF_SIZE;
      }
      const SRes r = p->coders[i].Code(p->coders[i].p, out, &out_len, in, &in_len, in_finished, finishMode, &said_finished);
      if (!said_finished) all_said_finished = false;
      if (i == 0) { *srcLen += in_len; src += in_len; }
      else p->pos[i - 1] += in_len;
      if (i == last) { *destLen += out_len; dest += out_len; }
      else { p->size[i] = out_len; p->pos[i] = 0; p->finished[i] = said_finished; }
      if (r != SZ_OK) return r;
      if (out_len != 0 || in_len != 0) progress = true;
    }
  }
  if (all_said_finished) *status = CODER_STATUS_FINISHED_WITH_MARK;
  return SZ_OK;
}

/* XzDec.c:558-588: filters are listed in encoding order, the chain runs them backwards */
extern "C" SRes XzDec_Init(CMixCoder *p, const CXzBlock *block)
{
  const int n = XzBlock_GetNumFilters(block);
  bool same = n == p->numCoders;
  for (int i = 0; same && i < n; i++) same = p->ids[i] == block->filters[n - 1 - i].id;
  if (!same)
  {
    MixCoder_Free(p);
    p->numCoders = n;
    for (int i = 0; i < n; i++)
    {
      const SRes r = MixCoder_SetFromMethod(p, i, block->filters[n - 1 - i].id);
      if (r != SZ_OK) return r;
    }
  }
  for (int i = 0; i < n; i++)
  {
    const CXzFilter *f = &block->filters[n - 1 - i];
    IStateCoder *sc = &p->coders[i];
    const SRes r = sc->SetProps(sc->p, f->props, f->propsSize, p->alloc);
    if (r != SZ_OK) return r;
  }
  MixCoder_Init(p);
  return SZ_OK;
}

/* ------------------------------------------------------------------ XzUnpacker ------------------ */
/* While state == XZ_STATE_BLOCK the reference does not use `pos`; this library parks one private marker
 * there: the payload of the block is complete (decoded ahead), the chain's "finished" has not been
 * delivered to the caller yet (see MixCoder_Code, observable). */
#define LZB_POS_BLOCK_DONE 0x4C5A4231u

static const Byte k_sig[6] = { 0xFD, '7', 'z', 'X', 'Z', 0 };

extern "C" SRes XzUnpacker_Create(CXzUnpacker *p, ISzAlloc *alloc)
{
  MixCoder_Construct(&p->decoder, alloc);
  p->state = XZ_STATE_STREAM_HEADER;
  p->pos = 0;
  p->numStreams = 0;
  p->numBlocks = 0;
  p->padSize = 0;
  return SZ_OK;
}

extern "C" void XzUnpacker_Free(CXzUnpacker *p) { MixCoder_Free(&p->decoder); }

extern "C" Bool XzUnpacker_IsStreamWasFinished(CXzUnpacker *p)
{
  return p->state == XZ_STATE_STREAM_PADDING && ((UInt32)p->padSize & 3) == 0;
}

extern "C" UInt64 XzUnpacker_GetNumBlocks(const CXzUnpacker *p) { return p->numBlocks; }

/* a block's payload is complete: index record, counters, on to the footer (XzDec.c:642-656) */
static void block_payload_done(CXzUnpacker *p)
{
  Byte rec[32];
  unsigned n = Xz_WriteVarInt(rec, p->packSize + p->blockHeaderSize + XzFlags_GetCheckSize(p->streamFlags));
  n += Xz_WriteVarInt(rec + n, p->unpackSize);
  sha_absorb(&p->sha, rec, n, false);
  p->indexSize += n;
  p->numBlocks++;
  p->state = XZ_STATE_BLOCK_FOOTER;
  p->pos = 0;
  p->alignPos = 0;
}

static SRes unpack(CXzUnpacker *p, LzbContext *ctx, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                   int finishMode, ECoderStatus *status)
{
  const SizeT dest_cap = *destLen, src_avail = *srcLen;
  const Byte *const src0 = src;
  Byte *const dest0 = dest;
  std::vector<XzReady> ready;          /* blocks decoded ahead in this call (offsets relative to src0/dest0) */
  size_t walked_to = 0;                /* no prefetch walk is started before this input offset */
  size_t next_ready = 0;
  *destLen = 0; *srcLen = 0;
  *status = CODER_STATUS_NOT_SPECIFIED;
  for (;;)
  {
    const size_t at = *srcLen;
    SizeT rem = src_avail - at;

    /* fast route: decode every whole block that lies ahead in one batched pass */
    if (p->pos == 0 && (p->state == XZ_STATE_BLOCK_HEADER || p->state == XZ_STATE_STREAM_HEADER) &&
        at >= walked_to && rem >= 32)
    {
      std::vector<XzReady> got;
      const SRes r = lzb_xz_prefetch(ctx, src0 + at, rem, p->state == XZ_STATE_BLOCK_HEADER, p->streamFlags,
                                     dest0 + *destLen, dest_cap - *destLen, got);
      if (r != SZ_OK) return r;
      walked_to = at + 1;
      for (auto &q : got)
      {
        q.data_off += at; q.dst_off += *destLen;
        walked_to = (size_t)(q.data_off + q.pack);
        ready.push_back(q);
      }
    }

    if (p->state == XZ_STATE_BLOCK)
    {
      SizeT out_n = dest_cap - *destLen, in_n = rem;
      if (in_n == 0 && out_n == 0) { *status = CODER_STATUS_NOT_FINISHED; return SZ_OK; }
      if (p->pos == LZB_POS_BLOCK_DONE)
      {
        /* the invocation that sees every stage already finished */
        *status = CODER_STATUS_FINISHED_WITH_MARK;
        block_payload_done(p);
        continue;
      }
      if (p->packSize == 0 && p->unpackSize == 0)
      {
        while (next_ready < ready.size() && ready[next_ready].data_off < at) next_ready++;
        if (next_ready < ready.size() && ready[next_ready].data_off == at && ready[next_ready].dst_off == *destLen)
        {
          const XzReady &q = ready[next_ready];
          /* observable: a lone LZMA2 stage stops at a full output buffer without reading the end marker
           * unless told to finish (Lzma2Dec_DecodeToBuf), so such a block goes the slow route */
          const bool usable = q.n_filters > 1 || finishMode == CODER_FINISH_END || q.dst_off + q.unpack < dest_cap;
          if (usable)
          {
            next_ready++;
            *srcLen += (SizeT)q.pack; src += q.pack;
            *destLen += (SizeT)q.unpack; dest += q.unpack;
            p->packSize = q.pack; p->unpackSize = q.unpack;
            const int ct = XzFlags_GetCheckSize(p->streamFlags) ? (p->streamFlags & 0xF) : 0;
            if (ct == XZ_CHECK_CRC32) p->check.crc = (UInt32)q.digest ^ 0xFFFFFFFFu;
            else if (ct == XZ_CHECK_CRC64) p->check.crc64 = ~q.digest;
            else if (ct == XZ_CHECK_SHA256) { p->check.mode = LZB_CHECK_SHA_READY; memcpy(p->check.sha.buffer, q.sha, 32); }
            /* observable: a chain of two stages needs a second invocation to report "finished" once its
             * first stage had to hand over a full buffer (output >= 128 KiB); that invocation only happens
             * while input or output room is left */
            const bool one_call = q.n_filters == 1 || q.unpack < LZB_CODER_BUF_SIZE;
            if (one_call) { *status = CODER_STATUS_FINISHED_WITH_MARK; block_payload_done(p); }
            else { *status = CODER_STATUS_NOT_FINISHED; p->pos = LZB_POS_BLOCK_DONE; }
            continue;
          }
        }
      }
      const SRes r = MixCoder_Code(&p->decoder, dest, &out_n, src, &in_n, 0, (ECoderFinishMode)finishMode, status);
      XzCheck_Update(&p->check, dest, out_n);
      *srcLen += in_n; src += in_n; p->packSize += in_n;
      *destLen += out_n; dest += out_n; p->unpackSize += out_n;
      if (r != SZ_OK) return r;
      if (*status == CODER_STATUS_FINISHED_WITH_MARK) block_payload_done(p);
      else if (in_n == 0 && out_n == 0) return SZ_OK;
      continue;
    }

    if (rem == 0) { *status = CODER_STATUS_NEEDS_MORE_INPUT; return SZ_OK; }

    switch (p->state)
    {
      case XZ_STATE_STREAM_HEADER:
        if (p->pos < 12)
        {
          if (p->pos < 6 && *src != k_sig[p->pos]) return SZ_ERROR_NO_ARCHIVE;
          p->buf[p->pos++] = *src++;
          (*srcLen)++;
        }
        else
        {
          /* parsed when the byte after the header arrives (XzDec.c:677-686) */
          const SRes r = Xz_ParseHeader(&p->streamFlags, p->buf);
          if (r != SZ_OK) return r;
          p->state = XZ_STATE_BLOCK_HEADER;
          Sha256_Init(&p->sha);
          p->indexSize = 0;
          p->numBlocks = 0;
          p->pos = 0;
        }
        break;

      case XZ_STATE_BLOCK_HEADER:
        if (p->pos == 0)
        {
          p->buf[p->pos++] = *src++;
          (*srcLen)++;
          if (p->buf[0] == 0)
          {
            /* index indicator: what the index must start with is known already (XzDec.c:694-703) */
            p->indexPreSize = 1 + Xz_WriteVarInt(p->buf + 1, p->numBlocks);
            p->indexPos = p->indexPreSize;
            p->indexSize += p->indexPreSize;
            Sha256_Final(&p->sha, p->shaDigest);
            Sha256_Init(&p->sha);
            p->crc = lzb_small_crc32_update(0xFFFFFFFFu, p->buf, p->indexPreSize);
            p->state = XZ_STATE_STREAM_INDEX;
          }
          p->blockHeaderSize = ((UInt32)p->buf[0] << 2) + 4;
        }
        else if (p->pos != p->blockHeaderSize)
        {
          const UInt32 n = (UInt32)std::min<size_t>(p->blockHeaderSize - p->pos, rem);
          memcpy(p->buf + p->pos, src, n);
          p->pos += n; *srcLen += n; src += n;
        }
        else
        {
          SRes r = XzBlock_Parse(&p->block, p->buf);
          if (r != SZ_OK) return r;
          p->state = XZ_STATE_BLOCK;
          p->pos = 0;                                   /* (unused by the reference in this state) */
          p->packSize = 0;
          p->unpackSize = 0;
          XzCheck_Init(&p->check, p->streamFlags & 0xF);
          r = XzDec_Init(&p->decoder, &p->block);
          if (r != SZ_OK) return r;
        }
        break;

      case XZ_STATE_BLOCK_FOOTER:
        if (((p->packSize + p->alignPos) & 3) != 0)
        {
          (*srcLen)++;
          p->alignPos++;
          if (*src++ != 0) return SZ_ERROR_CRC;
        }
        else
        {
          const UInt32 check_size = XzFlags_GetCheckSize(p->streamFlags);
          if (p->pos != check_size)
          {
            const UInt32 n = (UInt32)std::min<size_t>(check_size - p->pos, rem);
            memcpy(p->buf + p->pos, src, n);
            p->pos += n; *srcLen += n; src += n;
          }
          else
          {
            Byte digest[LZB_XZ_CHECK_SIZE_MAX];
            p->state = XZ_STATE_BLOCK_HEADER;
            p->pos = 0;
            if (XzCheck_Final(&p->check, digest) && memcmp(digest, p->buf, check_size) != 0) return SZ_ERROR_CRC;
          }
        }
        break;

      case XZ_STATE_STREAM_INDEX:
        if (p->pos < p->indexPreSize)
        {
          (*srcLen)++;
          if (*src++ != p->buf[p->pos++]) return SZ_ERROR_CRC;
        }
        else if (p->indexPos < p->indexSize)
        {
          /* the records: hashed, compared (by digest) with the blocks that were seen */
          const SizeT n = (SizeT)std::min<UInt64>(p->indexSize - p->indexPos, rem);
          p->crc = lzb_small_crc32_update(p->crc, src, n);
          sha_absorb(&p->sha, src, n, false);
          *srcLen += n; src += n; p->indexPos += n;
        }
        else if ((p->indexPos & 3) != 0)
        {
          const Byte b = *src++;
          p->crc = lzb_small_crc32_update(p->crc, &b, 1);
          (*srcLen)++;
          p->indexPos++;
          p->indexSize++;
          if (b != 0) return SZ_ERROR_CRC;
        }
        else
        {
          Byte digest[SHA256_DIGEST_SIZE];
          p->state = XZ_STATE_STREAM_INDEX_CRC;
          p->indexSize += 4;
          p->pos = 0;
          Sha256_Final(&p->sha, digest);
          if (memcmp(digest, p->shaDigest, SHA256_DIGEST_SIZE) != 0) return SZ_ERROR_CRC;
        }
        break;

      case XZ_STATE_STREAM_INDEX_CRC:
        if (p->pos < 4)
        {
          (*srcLen)++;
          p->buf[p->pos++] = *src++;
        }
        else
        {
          p->state = XZ_STATE_STREAM_FOOTER;
          p->pos = 0;
          if ((p->crc ^ 0xFFFFFFFFu) != lzb_le32(p->buf)) return SZ_ERROR_CRC;
        }
        break;

      case XZ_STATE_STREAM_FOOTER:
      {
        const UInt32 n = (UInt32)std::min<size_t>(12 - p->pos, rem);
        memcpy(p->buf + p->pos, src, n);
        p->pos += n; *srcLen += n; src += n;
        if (p->pos == 12)
        {
          p->state = XZ_STATE_STREAM_PADDING;
          p->numStreams++;
          p->padSize = 0;
          const Byte *f = p->buf;
          const bool ok = p->indexSize == (((UInt64)lzb_le32(f + 4) + 1) << 2) && lzb_le32(f) == lzb_small_crc32(f + 4, 6) &&
                          p->streamFlags == (((unsigned)f[8] << 8) | f[9]) && f[10] == 'Y' && f[11] == 'Z';
          if (!ok) return SZ_ERROR_CRC;
        }
        break;
      }

      case XZ_STATE_STREAM_PADDING:
        if (*src != 0)
        {
          if (((UInt32)p->padSize & 3) != 0) return SZ_ERROR_NO_ARCHIVE;
          p->pos = 0;
          p->state = XZ_STATE_STREAM_HEADER;
        }
        else
        {
          (*srcLen)++;
          src++;
          p->padSize++;
        }
        break;

      case XZ_STATE_BLOCK: break;
    }
  }
}

extern "C" SRes XzUnpacker_Code(CXzUnpacker *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                                int finishMode, ECoderStatus *status)
{
  return unpack(p, nullptr, dest, destLen, src, srcLen, finishMode, status);
}

/* Whole-buffer front end over the same machine (batched extension, include/lzb200.h). */
static void *std_alloc(void *, size_t n) { return n ? malloc(n) : 0; }
static void std_free(void *, void *a) { free(a); }

extern "C" SRes XzDecodeParallel(LzbContext *c, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                                 ECoderStatus *status, UInt64 *numBlocks, int *streamFinished)
{
  ISzAlloc alloc = { std_alloc, std_free };
  CXzUnpacker *u = (CXzUnpacker *)malloc(sizeof(CXzUnpacker));
  *numBlocks = 0; *streamFinished = 0;
  if (!u) { *destLen = *srcLen = 0; return SZ_ERROR_MEM; }
  XzUnpacker_Create(u, &alloc);
  /* call until a call makes no progress, as a whole-buffer caller of the reference does */
  const SizeT cap = *destLen, avail = *srcLen;
  SizeT op = 0, ip = 0;
  SRes r;
  for (;;)
  {
    SizeT dl = cap - op, sl = avail - ip;
    r = unpack(u, c, dest + op, &dl, src + ip, &sl, CODER_FINISH_ANY, status);
    op += dl; ip += sl;
    if (r != SZ_OK || (dl == 0 && sl == 0)) break;
  }
  *destLen = op; *srcLen = ip;
  *numBlocks = u->numBlocks;
  *streamFinished = XzUnpacker_IsStreamWasFinished(u) ? 1 : 0;
  XzUnpacker_Free(u);
  free(u);
  return r;
}
