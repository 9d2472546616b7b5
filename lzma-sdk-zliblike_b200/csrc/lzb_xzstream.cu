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
/* An xz branch / delta filter (xz format 5.3; reference object: BraState, XzDec.c:48-232) works in place on a
 * 16 KiB WINDOW of the stream that flows through it.  The window is three runs of bytes, front to back:
 *
 *      [ taken | ripe | raw ]        taken : already handed to the next stage
 *        0   t_end  r_end  fill      ripe  : filtered, waiting to be handed on
 *                                    raw   : not filtered yet (a converter stops short of the end of what it is
 *                                            given: an x86 candidate needs 5 bytes, ARM/PPC/SPARC 4, IA64 16)
 *
 * One call alternates "hand ripe bytes on" and "slide, refill from the input, filter" until the output is full or
 * nothing moves.  The window size and this order are what fix how many bytes each call consumes and produces
 * -- which a caller of XzUnpacker_Code can observe -- so they are kept; the converters run on the GPU. */
struct FilterStage
{
  size_t t_end, r_end, fill;
  UInt32 method;
  int encode;
  UInt32 delta, ip, x86_state;
  Byte delta_state[DELTA_STATE_SIZE];
  Byte win[LZB_BRA_BUF_SIZE];
};

/* what the filter ids of the xz format ask for: start-offset alignment and the converter behind it */
typedef size_t (*FilterRun)(FilterStage *f, Byte *data, size_t n);
static size_t run_delta(FilterStage *f, Byte *d, size_t n)
{
  if (f->encode) Delta_Encode(f->delta_state, f->delta, d, n); else Delta_Decode(f->delta_state, f->delta, d, n);
  return n;
}
static size_t run_x86(FilterStage *f, Byte *d, size_t n) { return x86_Convert(d, n, f->ip, &f->x86_state, f->encode); }
static size_t run_ppc(FilterStage *f, Byte *d, size_t n) { return PPC_Convert(d, n, f->ip, f->encode); }
static size_t run_ia64(FilterStage *f, Byte *d, size_t n) { return IA64_Convert(d, n, f->ip, f->encode); }
static size_t run_arm(FilterStage *f, Byte *d, size_t n) { return ARM_Convert(d, n, f->ip, f->encode); }
static size_t run_armt(FilterStage *f, Byte *d, size_t n) { return ARMT_Convert(d, n, f->ip, f->encode); }
static size_t run_sparc(FilterStage *f, Byte *d, size_t n) { return SPARC_Convert(d, n, f->ip, f->encode); }

static const struct FilterKind { UInt32 id; UInt32 offset_mask; FilterRun run; } k_filter_kinds[] =
{
  { XZ_ID_Delta, 0, run_delta }, { XZ_ID_X86, 0, run_x86 },   { XZ_ID_PPC, 3, run_ppc },     { XZ_ID_IA64, 15, run_ia64 },
  { XZ_ID_ARM, 3, run_arm },     { XZ_ID_ARMT, 1, run_armt }, { XZ_ID_SPARC, 3, run_sparc },
};

static const FilterKind *filter_kind(UInt64 id)
{
  for (const FilterKind &k : k_filter_kinds) if (k.id == id) return &k;
  return nullptr;
}

static void stage_release(void *pp, ISzAlloc *alloc) { alloc->Free(alloc, pp); }

/* filter properties (xz format 5.3.1 / 5.3.2): Delta = one byte "distance - 1"; a branch converter = nothing, or a
 * 32-bit start offset that must respect the instruction alignment of its architecture */
static SRes stage_props(void *pp, const Byte *props, size_t n, ISzAlloc *)
{
  FilterStage *f = (FilterStage *)pp;
  const FilterKind *kind = filter_kind(f->method);
  f->encode = 0;
  f->ip = 0;
  if (f->method == XZ_ID_Delta)
  {
    if (n != 1) return SZ_ERROR_UNSUPPORTED;
    f->delta = 1u + props[0];
    return SZ_OK;
  }
  if (n == 0) return SZ_OK;
  if (n != 4 || !kind) return SZ_ERROR_UNSUPPORTED;
  const UInt32 start = lzb_le32(props);
  if (start & kind->offset_mask) return SZ_ERROR_UNSUPPORTED;
  f->ip = start;
  return SZ_OK;
}

static void stage_reset(void *pp)
{
  FilterStage *f = (FilterStage *)pp;
  f->t_end = f->r_end = f->fill = 0;
  f->x86_state = 0;
  if (f->method == XZ_ID_Delta) Delta_Init(f->delta_state);
}

static SRes stage_code(void *pp, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, int srcWasFinished,
                       ECoderFinishMode, int *wasFinished)
{
  FilterStage *f = (FilterStage *)pp;
  const FilterKind *kind = filter_kind(f->method);
  Byte *out = dest, *const out_end = dest + *destLen;
  const Byte *in = src, *const in_end = src + *srcLen;
  *wasFinished = 0;
  while (out != out_end)
  {
    if (f->t_end != f->r_end)
    {
      /* hand ripe bytes on */
      const size_t n = std::min<size_t>(f->r_end - f->t_end, (size_t)(out_end - out));
      memcpy(out, f->win + f->t_end, n);
      out += n; f->t_end += n;
      continue;
    }
    /* slide what is still raw to the front, refill, filter */
    f->fill -= f->t_end;
    memmove(f->win, f->win + f->t_end, f->fill);
    f->t_end = f->r_end = 0;
    const size_t n = std::min<size_t>(LZB_BRA_BUF_SIZE - f->fill, (size_t)(in_end - in));
    memcpy(f->win + f->fill, in, n);
    in += n; f->fill += n;
    if (f->fill == 0) break;
    if (!kind) { *destLen = (SizeT)(out - dest); *srcLen = (SizeT)(in - src); return SZ_ERROR_UNSUPPORTED; }
    f->r_end = kind->run(f, f->win, f->fill);
    f->ip += (UInt32)f->r_end;
    if (f->r_end == 0)
    {
      if (!srcWasFinished) break;         /* the raw tail may still grow into something convertible */
      f->r_end = f->fill;                 /* end of the data: the last bytes pass as they are */
    }
  }
  *destLen = (SizeT)(out - dest);
  *srcLen = (SizeT)(in - src);
  if (srcWasFinished && in == in_end && f->t_end == f->fill) *wasFinished = 1;
  return SZ_OK;
}

static SRes stage_attach(IStateCoder *sc, UInt64 id, ISzAlloc *alloc)
{
  sc->p = 0;
  if (!filter_kind(id)) return SZ_ERROR_UNSUPPORTED;
  FilterStage *f = (FilterStage *)alloc->Alloc(alloc, sizeof(FilterStage));
  if (!f) return SZ_ERROR_MEM;
  f->method = (UInt32)id;
  sc->p = f;
  sc->Free = stage_release; sc->SetProps = stage_props; sc->Init = stage_reset; sc->Code = stage_code;
  return SZ_OK;
}

/* ------------------------------------------------------------------ LZMA2 stage ----------------- */
static CLzma2Dec *l2_of(void *stage) { return static_cast<CLzma2Dec *>(stage); }

static void l2_free(void *pp, ISzAlloc *alloc)
{
  CLzma2Dec *d = l2_of(pp);
  Lzma2Dec_Free(d, alloc);                               /* device state, control block, dictionary */
  alloc->Free(alloc, d);
}

/* LZMA2 as an xz filter has one property byte: the dictionary size code (xz format 5.3.1) */
static SRes l2_set_props(void *pp, const Byte *props, size_t propSize, ISzAlloc *alloc)
{
  return propSize == 1 ? Lzma2Dec_Allocate(l2_of(pp), props[0], alloc) : SZ_ERROR_UNSUPPORTED;
}

static void l2_init(void *pp) { Lzma2Dec_Init(l2_of(pp)); }

static SRes l2_code(void *pp, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, int srcWasFinished,
                    ECoderFinishMode finishMode, int *wasFinished)
{
  ELzmaStatus st;
  (void)srcWasFinished;
  const SRes r = Lzma2Dec_DecodeToBuf(l2_of(pp), dest, destLen, src, srcLen, (ELzmaFinishMode)finishMode, &st);
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
  memset(p, 0, sizeof *p);                               /* no stages, no hand-over buffers yet */
  p->alloc = alloc;
}

extern "C" void MixCoder_Free(CMixCoder *p)
{
  for (IStateCoder *st = p->coders, *end = p->coders + p->numCoders; st != end; st++)
  {
    if (p->alloc && st->p) st->Free(st->p, p->alloc);
    st->p = NULL;
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
  IStateCoder *const slot = p->coders + coderIndex;
  p->ids[coderIndex] = methodId;
  if (methodId == XZ_ID_LZMA2) return l2_attach(slot, p->alloc);
  if (coderIndex == 0) return SZ_ERROR_UNSUPPORTED;        /* a chain starts with a real decoder */
  return stage_attach(slot, methodId, p->alloc);
}

/* The chain of a block (xz format 3.1.2: filters are listed in the order the ENCODER applied them; the decoder runs
 * them backwards).  Reference object: CMixCoder, XzDec.c:356-480 -- its fields are ABI (include/lzb200.h).
 *
 * Neighbouring stages are joined by a hand-over buffer of 128 KiB: link j carries the bytes stage j has produced
 * (size[j]), how many of them stage j+1 has taken (pos[j]) and whether stage j said it was finished.  A stage may
 * write into its link only when the link is empty -- it always gets the whole buffer.  Rounds over the stages repeat
 * until a whole round moves no byte.
 * observable: `status` is FINISHED_WITH_MARK only when EVERY stage call made during this invocation said
 * "finished" -- a chain that finishes in a later round of the same invocation still reports NOT_FINISHED and is
 * seen finished by the next invocation. */
namespace {

struct ByteRun { const Byte *p; SizeT n; int ended; };          /* what a stage may read */
struct ByteRoom { Byte *p; SizeT n; };                          /* where a stage may write */

struct Chain
{
  CMixCoder *m;
  const Byte *src; SizeT src_left; int src_ended;
  Byte *dest; SizeT dest_left;

  Byte *link(int j) const { return m->buf + LZB_CODER_BUF_SIZE * (size_t)j; }
  int last() const { return m->numCoders - 1; }

  ByteRun input_of(int i) const
  {
    if (i == 0) return { src, src_left, src_ended };
    const int j = i - 1;
    return { link(j) + m->pos[j], (SizeT)(m->size[j] - m->pos[j]), m->finished[j] };
  }
  /* false: the stage's link still holds bytes the next stage has not taken */
  bool output_of(int i, ByteRoom &r) const
  {
    if (i == last()) { r = { dest, dest_left }; return true; }
    if (m->pos[i] != m->size[i]) return false;
    r = { link(i), (SizeT)LZB_CODER_BUF_SIZE };
    return true;
  }
  void consumed(int i, SizeT n)
  {
    if (i == 0) { src += n; src_left -= n; }
    else m->pos[i - 1] += n;
  }
  void produced(int i, SizeT n, int ended)
  {
    if (i == last()) { dest += n; dest_left -= n; }
    else { m->size[i] = n; m->pos[i] = 0; m->finished[i] = ended; }
  }
};

}  // namespace

extern "C" SRes MixCoder_Code(CMixCoder *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, int srcWasFinished,
                              ECoderFinishMode finishMode, ECoderStatus *status)
{
  Chain c = { p, src, *srcLen, srcWasFinished, dest, *destLen };
  const SizeT src_all = *srcLen, dest_all = *destLen;
  *srcLen = 0; *destLen = 0;
  *status = CODER_STATUS_NOT_FINISHED;
  if (!p->buf)
  {
    p->buf = (Byte *)p->alloc->Alloc(p->alloc, LZB_CODER_BUF_SIZE * (MIXCODER_NUM_FILTERS_MAX - 1));
    if (!p->buf) return SZ_ERROR_MEM;
  }
  /* only a lone stage sees the caller's finish mode: inside a chain "the output is full" says nothing about the end */
  const ECoderFinishMode mode = p->numCoders == 1 ? finishMode : CODER_FINISH_ANY;
  int calls = 0, calls_finished = 0;
  SRes res = SZ_OK;
  for (bool moved = true; moved && res == SZ_OK;)
  {
    moved = false;
    for (int i = 0; i <= c.last() && res == SZ_OK; i++)
    {
      ByteRoom room;
      if (!c.output_of(i, room)) continue;
      const ByteRun in = c.input_of(i);
      SizeT took = in.n, made = room.n;
      int ended = 0;
      IStateCoder &st = p->coders[i];
      res = st.Code(st.p, room.p, &made, in.p, &took, in.ended, mode, &ended);
      calls++; calls_finished += ended ? 1 : 0;
      c.consumed(i, took);
      c.produced(i, made, ended);
      if (took != 0 || made != 0) moved = true;
    }
  }
  *srcLen = src_all - c.src_left;
  *destLen = dest_all - c.dest_left;
  if (res != SZ_OK) return res;
  if (calls == calls_finished) *status = CODER_STATUS_FINISHED_WITH_MARK;
  return SZ_OK;
}

/* Bind the chain to a parsed block header (reference entry point: XzDec_Init, XzDec.c:558-588).  Stage i runs the
 * block's filter n-1-i.  The stage objects are kept when the new block asks for the same methods in the same
 * order, which is the normal case for every block of a stream after the first. */
static bool chain_matches(const CMixCoder *p, const CXzBlock *block, int n)
{
  if (p->numCoders != n) return false;
  for (int i = 0; i < n; i++)
    if (p->ids[i] != block->filters[n - 1 - i].id) return false;
  return true;
}

extern "C" SRes XzDec_Init(CMixCoder *p, const CXzBlock *block)
{
  const int n = XzBlock_GetNumFilters(block);
  if (!chain_matches(p, block, n))
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
    const CXzFilter &f = block->filters[n - 1 - i];
    const SRes r = p->coders[i].SetProps(p->coders[i].p, f.props, f.propsSize, p->alloc);
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

extern "C" SRes XzUnpacker_Create(CXzUnpacker *p, ISzAlloc *alloc)
{
  MixCoder_Construct(&p->decoder, alloc);
  p->state = XZ_STATE_STREAM_HEADER;                       /* at the first byte of the first stream's magic */
  p->pos = 0;
  p->numStreams = p->numBlocks = p->padSize = 0;
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
  p->numBlocks += 1;
  p->alignPos = 0;
  p->state = XZ_STATE_BLOCK_FOOTER;
  p->pos = 0;
}

/* ---- container framing (xz format 2.1 stream, 3.1 block header, 3.3 padding, 3.4 check, 4 index, 2.1.2 footer,
 * 2.2 stream padding).  The machine's position inside the container is (state, pos, ...) of the public struct;
 * pending_of() turns it into "which ELEMENT of the format is being read and how many of its bytes are still
 * outstanding".  frame_feed() absorbs bytes of that element, frame_close() acts on a complete one.
 *
 * observable: the reference asks for input before every step, also before acting on an element that is already
 * complete (XzDec.c:665-670) -- a header is parsed only when the byte AFTER it has arrived; unpack() keeps that by
 * returning NEEDS_MORE_INPUT before it comes here.  The stream footer is the one element acted on as soon as its
 * last byte is in.  Bytes the format fixes (magic, index head, padding) are checked as they arrive: a wrong magic byte
 * is refused before it is consumed, a wrong index-head or padding byte after. */
namespace {

const unsigned kMagicBytes = 6, kStreamHeaderBytes = 12, kStreamFooterBytes = 12;      /* xz format 2.1.1, 2.1.2 */

enum Element
{
  EL_STREAM_HEADER,      /* magic (6) + stream flags (2) + CRC32 (4)                                         */
  EL_BLOCK_SIZE,         /* first byte of a block header (size / 4 - 1), or 0x00 = the index follows          */
  EL_BLOCK_HEADER,       /* the rest of the block header                                                     */
  EL_BLOCK_PAD,          /* zero bytes up to a multiple of four after the block's payload                    */
  EL_BLOCK_CHECK,        /* the block's check value                                                          */
  EL_INDEX_HEAD,         /* index indicator + number of records: known in advance, compared                  */
  EL_INDEX_RECORDS,      /* (unpadded size, uncompressed size) pairs: hashed, the digest is compared         */
  EL_INDEX_PAD,          /* zero bytes up to a multiple of four                                              */
  EL_INDEX_CRC,          /* CRC32 of the index                                                               */
  EL_STREAM_FOOTER,      /* CRC32 (4) + backward size (4) + stream flags (2) + "YZ"                          */
  EL_STREAM_PAD          /* zero bytes (a multiple of four) before another stream or the end                 */
};

struct Pending { Element el; UInt64 bytes; };            /* bytes == 0: the element is complete */

Pending pending_of(const CXzUnpacker *p)
{
  switch (p->state)
  {
    case XZ_STATE_STREAM_HEADER: return { EL_STREAM_HEADER, (UInt64)kStreamHeaderBytes - p->pos };
    case XZ_STATE_BLOCK_HEADER:
      if (p->pos == 0) return { EL_BLOCK_SIZE, 1 };
      return { EL_BLOCK_HEADER, (UInt64)p->blockHeaderSize - p->pos };
    case XZ_STATE_BLOCK_FOOTER:
    {
      const unsigned pad = (unsigned)(0u - (unsigned)(p->packSize + p->alignPos)) & 3u;
      if (pad != 0) return { EL_BLOCK_PAD, pad };
      return { EL_BLOCK_CHECK, (UInt64)XzFlags_GetCheckSize(p->streamFlags) - p->pos };
    }
    case XZ_STATE_STREAM_INDEX:
      if (p->pos < p->indexPreSize) return { EL_INDEX_HEAD, (UInt64)p->indexPreSize - p->pos };
      if (p->indexPos < p->indexSize) return { EL_INDEX_RECORDS, p->indexSize - p->indexPos };
      return { EL_INDEX_PAD, (0u - (unsigned)p->indexPos) & 3u };
    case XZ_STATE_STREAM_INDEX_CRC: return { EL_INDEX_CRC, 4u - (UInt64)p->pos };
    case XZ_STATE_STREAM_FOOTER: return { EL_STREAM_FOOTER, (UInt64)kStreamFooterBytes - p->pos };
    default: return { EL_STREAM_PAD, 1 };
  }
}

/* move to another part of the container; every part starts with its byte counter at zero */
void enter(CXzUnpacker *p, EXzState part)
{
  p->state = part;
  p->pos = 0;
}

void keep(CXzUnpacker *p, const Byte *src, size_t n)
{
  memcpy(p->buf + p->pos, src, n);
  p->pos += (UInt32)n;
}

/* the index of a stream starts with bytes the decoder can predict: 0x00 and the number of blocks it has seen */
void index_begins(CXzUnpacker *p)
{
  const unsigned head = 1u + Xz_WriteVarInt(&p->buf[1], p->numBlocks);
  p->indexPreSize = head;
  p->indexPos = head;
  p->indexSize += head;
  Sha256_Final(&p->sha, &p->shaDigest[0]);               /* digest of the records the blocks implied */
  lzb::sha256_init(p->sha.state); p->sha.count = 0;      /* ... against the records that are stored    */
  p->crc = lzb_small_crc32_update(0xFFFFFFFFu, p->buf, head);
  p->state = XZ_STATE_STREAM_INDEX;                      /* (pos stays 1: the indicator byte is the head's first) */
}

/* absorb up to n (<= want.bytes, <= what the caller has) bytes of the element being read */
SRes frame_feed(CXzUnpacker *p, Pending want, const Byte *src, size_t n, size_t *used)
{
  static const Byte magic[kMagicBytes] = { 0xFD, '7', 'z', 'X', 'Z', 0 };
  switch (want.el)
  {
    case EL_STREAM_HEADER:
    {
      size_t ok = 0;
      while (ok < n && (p->pos + ok >= kMagicBytes || src[ok] == magic[p->pos + ok])) ok++;
      keep(p, src, ok);
      *used = ok;
      return ok == n ? SZ_OK : SZ_ERROR_NO_ARCHIVE;
    }
    case EL_BLOCK_SIZE:
      keep(p, src, 1);
      *used = 1;
      if (src[0] == 0) index_begins(p);
      p->blockHeaderSize = 4u * ((UInt32)src[0] + 1u);
      return SZ_OK;
    case EL_BLOCK_HEADER:
    case EL_BLOCK_CHECK:
    case EL_INDEX_CRC:
      keep(p, src, n);
      *used = n;
      return SZ_OK;
    case EL_BLOCK_PAD:
      for (size_t i = 0; i < n; i++)
      {
        p->alignPos++;
        *used = i + 1;
        if (src[i] != 0) return SZ_ERROR_CRC;
      }
      return SZ_OK;
    case EL_INDEX_HEAD:
      for (size_t i = 0; i < n; i++)
      {
        *used = i + 1;
        if (src[i] != p->buf[p->pos++]) return SZ_ERROR_CRC;
      }
      return SZ_OK;
    case EL_INDEX_RECORDS:
      p->crc = lzb_small_crc32_update(p->crc, src, n);
      sha_absorb(&p->sha, src, n, false);
      p->indexPos += n;
      *used = n;
      return SZ_OK;
    case EL_INDEX_PAD:
      for (size_t i = 0; i < n; i++)
      {
        p->crc = lzb_small_crc32_update(p->crc, src + i, 1);
        p->indexPos++;
        p->indexSize++;
        *used = i + 1;
        if (src[i] != 0) return SZ_ERROR_CRC;
      }
      return SZ_OK;
    case EL_STREAM_FOOTER:
    {
      keep(p, src, n);
      *used = n;
      if (p->pos != kStreamFooterBytes) return SZ_OK;
      p->state = XZ_STATE_STREAM_PADDING;                /* (pos keeps the footer's 12: padding counts in padSize) */
      p->numStreams += 1;
      p->padSize = 0;
      const Byte *f = p->buf;
      const bool good = lzb_le32(f) == lzb_small_crc32(f + 4, 6) &&
                        p->indexSize == (((UInt64)lzb_le32(f + 4) + 1) << 2) &&
                        p->streamFlags == (((unsigned)f[8] << 8) | f[9]) && f[10] == 'Y' && f[11] == 'Z';
      return good ? SZ_OK : SZ_ERROR_CRC;
    }
    case EL_STREAM_PAD:
      if (src[0] == 0) { p->padSize++; *used = 1; return SZ_OK; }
      /* something else begins here: only on a four-byte boundary, and it is not consumed by this element */
      if (((UInt32)p->padSize & 3) != 0) return SZ_ERROR_NO_ARCHIVE;
      enter(p, XZ_STATE_STREAM_HEADER);
      return SZ_OK;
  }
  return SZ_ERROR_FAIL;
}

/* act on a complete element */
SRes frame_close(CXzUnpacker *p, Element el)
{
  switch (el)
  {
    case EL_STREAM_HEADER:
    {
      const SRes r = Xz_ParseHeader(&p->streamFlags, p->buf);
      if (r != SZ_OK) return r;
      lzb::sha256_init(p->sha.state); p->sha.count = 0;
      p->indexSize = p->numBlocks = 0;
      enter(p, XZ_STATE_BLOCK_HEADER);
      return SZ_OK;
    }
    case EL_BLOCK_HEADER:
    {
      SRes r = XzBlock_Parse(&p->block, p->buf);
      if (r != SZ_OK) return r;
      enter(p, XZ_STATE_BLOCK);                          /* (this library's own marker lives in pos while in BLOCK) */
      p->packSize = p->unpackSize = 0;
      XzCheck_Init(&p->check, p->streamFlags & 0xF);
      return XzDec_Init(&p->decoder, &p->block);
    }
    case EL_BLOCK_CHECK:
    {
      Byte digest[LZB_XZ_CHECK_SIZE_MAX];
      const UInt32 n = XzFlags_GetCheckSize(p->streamFlags);
      enter(p, XZ_STATE_BLOCK_HEADER);
      return (XzCheck_Final(&p->check, digest) && memcmp(digest, p->buf, n) != 0) ? SZ_ERROR_CRC : SZ_OK;
    }
    case EL_INDEX_PAD:
    {
      Byte stored[SHA256_DIGEST_SIZE];
      enter(p, XZ_STATE_STREAM_INDEX_CRC);
      p->indexSize += 4;                                 /* the CRC32 that follows belongs to the index's size */
      Sha256_Final(&p->sha, stored);
      return memcmp(stored, p->shaDigest, SHA256_DIGEST_SIZE) != 0 ? SZ_ERROR_CRC : SZ_OK;
    }
    case EL_INDEX_CRC:
      enter(p, XZ_STATE_STREAM_FOOTER);
      return (p->crc ^ 0xFFFFFFFFu) != lzb_le32(p->buf) ? SZ_ERROR_CRC : SZ_OK;
    default:
      return SZ_ERROR_FAIL;                              /* the other elements are never "complete and pending" */
  }
}

}  // namespace

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

    /* container framing: one element of the format at a time (below) */
    const Pending want = pending_of(p);
    size_t used = 0;
    const SRes r = want.bytes != 0 ? frame_feed(p, want, src, (size_t)std::min<UInt64>(want.bytes, rem), &used)
                                   : frame_close(p, want.el);
    src += used; *srcLen += used;
    if (r != SZ_OK) return r;
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
