/* lzb_state.cu -- the SDK's stateful interfaces (dictionary / buffer interface of LzmaDec.h and the
 * Lzma2Dec.h state interface) on top of the GPU path.
 *
 * CLzmaDec.probs does not hold probabilities here (they live in device memory); it points at a small
 * host control block obtained from the caller's ISzAlloc, so Allocate/Free pair up as in the reference.
 *
 * Round-1 behaviour: a decoder object decodes ONE whole buffer per LzmaDec_Init / Lzma2Dec_Init -- the
 * pattern of LzmaDecode (LzmaDec.c:984-999) and of 7zDec.c:134-152,184-202 when the folder's packed
 * bytes are handed over in one piece.  A call that would have to RESUME a suspended stream (partial
 * input fed in pieces, LzmaUtil.c:55-106) returns SZ_ERROR_UNSUPPORTED with an explanatory
 * Lzb_LastError(); it never falls back to a CPU decoder.  Suspend/resume of device-resident state is
 * the next item of SURVEY.md §8f.
 */
#include "lzb_host.h"

#include <stdlib.h>
#include <string.h>

struct LzbCtl
{
  uint64_t magic;
  int owns_dic;
  int used;            /* a whole-buffer decode has consumed this Init */
};
#define LZB_CTL_MAGIC 0x4C5A4243544C3031ull

static LzbCtl *ctl_of(CLzmaDec *p) { return (LzbCtl *)(void *)p->probs; }

static SRes alloc_ctl(CLzmaDec *p, const CLzmaProps *pr, ISzAlloc *alloc)
{
  /* reference LzmaDec.c:924-936: re-use when the model size is unchanged */
  const UInt32 num = 1846u + (768u << (pr->lc + pr->lp));   /* reported in numProbs like the reference */
  if (p->probs == 0 || num != p->numProbs)
  {
    LzmaDec_FreeProbs(p, alloc);
    LzbCtl *c = (LzbCtl *)alloc->Alloc(alloc, sizeof(LzbCtl));
    p->numProbs = num;
    if (!c) return SZ_ERROR_MEM;
    memset(c, 0, sizeof *c);
    c->magic = LZB_CTL_MAGIC;
    p->probs = (CLzmaProb *)(void *)c;
  }
  return SZ_OK;
}

extern "C" void LzmaDec_FreeProbs(CLzmaDec *p, ISzAlloc *alloc)
{
  alloc->Free(alloc, p->probs);
  p->probs = 0;
}

extern "C" void LzmaDec_Free(CLzmaDec *p, ISzAlloc *alloc)
{
  LzmaDec_FreeProbs(p, alloc);
  alloc->Free(alloc, p->dic);
  p->dic = 0;
}

extern "C" SRes LzmaDec_AllocateProbs(CLzmaDec *p, const Byte *props, unsigned propsSize, ISzAlloc *alloc)
{
  CLzmaProps pr;
  SRes r = LzmaProps_Decode(&pr, props, propsSize);
  if (r != SZ_OK) return r;
  r = alloc_ctl(p, &pr, alloc);
  if (r != SZ_OK) return r;
  p->prop = pr;
  return SZ_OK;                       /* (the fork's debug printf, LzmaDec.c:944-946, is not reproduced) */
}

extern "C" SRes LzmaDec_Allocate(CLzmaDec *p, const Byte *props, unsigned propsSize, ISzAlloc *alloc)
{
  /* reference LzmaDec.c:950-970 */
  CLzmaProps pr;
  SRes r = LzmaProps_Decode(&pr, props, propsSize);
  if (r != SZ_OK) return r;
  r = alloc_ctl(p, &pr, alloc);
  if (r != SZ_OK) return r;
  const SizeT need = pr.dicSize;
  if (p->dic == 0 || need != p->dicBufSize)
  {
    alloc->Free(alloc, p->dic);
    p->dic = (Byte *)alloc->Alloc(alloc, need);
    if (p->dic == 0) { LzmaDec_FreeProbs(p, alloc); return SZ_ERROR_MEM; }
  }
  p->dicBufSize = need;
  p->prop = pr;
  return SZ_OK;
}

extern "C" void LzmaDec_InitDicAndState(CLzmaDec *p, Bool initDic, Bool initState)
{
  /* reference LzmaDec.c:685-699 */
  p->needFlush = 1;
  p->remainLen = 0;
  p->tempBufSize = 0;
  if (initDic) { p->processedPos = 0; p->checkDicSize = 0; p->needInitState = 1; }
  if (initState) p->needInitState = 1;
}

extern "C" void LzmaDec_Init(CLzmaDec *p)
{
  p->dicPos = 0;
  LzmaDec_InitDicAndState(p, 1, 1);
  if (p->probs && ctl_of(p)->magic == LZB_CTL_MAGIC) ctl_of(p)->used = 0;
}

static SRes need_fresh(CLzmaDec *p, const char *who)
{
  if (!p->probs || ctl_of(p)->magic != LZB_CTL_MAGIC) { lzb_set_error("%s: decoder not allocated", who); return SZ_ERROR_PARAM; }
  if (ctl_of(p)->used || !p->needFlush || p->processedPos != 0 || p->dicPos != 0)
  {
    lzb_set_error("%s: resuming a suspended stream is not supported in this round; call the Init function and "
                  "present the whole packed buffer in one call", who);
    return SZ_ERROR_UNSUPPORTED;
  }
  return SZ_OK;
}

extern "C" SRes LzmaDec_DecodeToDic(CLzmaDec *p, SizeT dicLimit, const Byte *src, SizeT *srcLen,
                                    ELzmaFinishMode finishMode, ELzmaStatus *status)
{
  const SizeT inSize = *srcLen;
  *srcLen = 0;
  *status = LZMA_STATUS_NOT_SPECIFIED;
  SRes r = need_fresh(p, "LzmaDec_DecodeToDic");
  if (r != SZ_OK) return r;
  if (dicLimit > p->dicBufSize) return SZ_ERROR_PARAM;
  if (inSize > 0xFFFFFFFFull || dicLimit > 0xFFFFFFFFull) return SZ_ERROR_UNSUPPORTED;
  LzbContext *c = lzb_default_context();
  if (!c) return SZ_ERROR_FAIL;
  LzbUnit u;
  LzbResult x;
  memset(&u, 0, sizeof u);
  u.src_len = (uint32_t)inSize; u.dst_cap = (uint32_t)dicLimit; u.dict_size = p->prop.dicSize;
  u.codec = LZB_CODEC_LZMA; u.lc = (uint8_t)p->prop.lc; u.lp = (uint8_t)p->prop.lp; u.pb = (uint8_t)p->prop.pb;
  u.finish = (uint8_t)finishMode;
  if (inSize < 5)
  {
    /* LzmaDec.c:732-740: the bytes are parked in tempBuf and more input is requested */
    memcpy(p->tempBuf, src, inSize);
    p->tempBufSize = (unsigned)inSize;
    *srcLen = inSize;
    *status = LZMA_STATUS_NEEDS_MORE_INPUT;
    ctl_of(p)->used = 1;
    return SZ_OK;
  }
  r = Lzb_DecodeUnitsHost(c, src, inSize, p->dic, dicLimit, &u, &x, 1);
  if (r != SZ_OK) return r;
  ctl_of(p)->used = 1;
  *srcLen = x.src_used;
  p->dicPos = x.dst_done;
  p->processedPos = x.dst_done;
  p->needFlush = 0;
  p->needInitState = 0;
  *status = (ELzmaStatus)x.status;
  /* the one-call driver folds NEEDS_MORE_INPUT into SZ_ERROR_INPUT_EOF (LzmaDec.c:996-997); undo that */
  if (x.res == SZ_ERROR_INPUT_EOF) return SZ_OK;
  return x.res;
}

extern "C" SRes LzmaDec_DecodeToBuf(CLzmaDec *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                                    ELzmaFinishMode finishMode, ELzmaStatus *status)
{
  /* reference LzmaDec.c:840-878, for a decoder right after LzmaDec_Init and an output that fits the
   * dictionary buffer (one slice of the reference's loop) */
  const SizeT outSize = *destLen;
  SizeT inSize = *srcLen;
  *srcLen = *destLen = 0;
  if (outSize > p->dicBufSize)
  {
    lzb_set_error("LzmaDec_DecodeToBuf: output larger than the dictionary buffer needs the resumable path (next round)");
    *status = LZMA_STATUS_NOT_SPECIFIED;
    return SZ_ERROR_UNSUPPORTED;
  }
  SRes r = LzmaDec_DecodeToDic(p, outSize, src, &inSize, finishMode, status);
  *srcLen = inSize;
  if (r == SZ_ERROR_UNSUPPORTED || r == SZ_ERROR_PARAM || r == SZ_ERROR_FAIL) return r;
  memcpy(dest, p->dic, p->dicPos);
  *destLen = p->dicPos;
  return r;
}

/* ------------------------------------------------------------------ LZMA2 state interface ----- */
static void lzma2_old_props(Byte prop, Byte *p5)
{
  /* reference Lzma2Dec.c:61-73 */
  const UInt32 dict = prop == 40 ? 0xFFFFFFFFu : ((UInt32)(2 | (prop & 1)) << (prop / 2 + 11));
  p5[0] = 4;
  p5[1] = (Byte)dict; p5[2] = (Byte)(dict >> 8); p5[3] = (Byte)(dict >> 16); p5[4] = (Byte)(dict >> 24);
}

/* the LZMA2 dictionary byte is kept in the otherwise unused CLzma2Dec.control until decoding starts */
extern "C" SRes Lzma2Dec_AllocateProbs(CLzma2Dec *p, Byte prop, ISzAlloc *alloc)
{
  Byte p5[LZMA_PROPS_SIZE];
  if (prop > 40) return SZ_ERROR_UNSUPPORTED;
  lzma2_old_props(prop, p5);
  return LzmaDec_AllocateProbs(&p->decoder, p5, LZMA_PROPS_SIZE, alloc);
}

extern "C" SRes Lzma2Dec_Allocate(CLzma2Dec *p, Byte prop, ISzAlloc *alloc)
{
  Byte p5[LZMA_PROPS_SIZE];
  if (prop > 40) return SZ_ERROR_UNSUPPORTED;
  lzma2_old_props(prop, p5);
  return LzmaDec_Allocate(&p->decoder, p5, LZMA_PROPS_SIZE, alloc);
}

extern "C" void Lzma2Dec_Init(CLzma2Dec *p)
{
  /* reference Lzma2Dec.c:89-96 */
  p->state = 0;
  p->needInitDic = 1; p->needInitState = 1; p->needInitProp = 1;
  LzmaDec_Init(&p->decoder);
}

static Byte lzma2_prop_from_dict(UInt32 dict)
{
  for (Byte q = 0; q <= 40; q++)
  {
    const UInt32 d = q == 40 ? 0xFFFFFFFFu : ((UInt32)(2 | (q & 1)) << (q / 2 + 11));
    if (d >= dict) return q;
  }
  return 40;
}

extern "C" SRes Lzma2Dec_DecodeToDic(CLzma2Dec *p, SizeT dicLimit, const Byte *src, SizeT *srcLen,
                                     ELzmaFinishMode finishMode, ELzmaStatus *status)
{
  CLzmaDec *d = &p->decoder;
  const SizeT inSize = *srcLen;
  *srcLen = 0;
  *status = LZMA_STATUS_NOT_SPECIFIED;
  SRes r = need_fresh(d, "Lzma2Dec_DecodeToDic");
  if (r != SZ_OK) return r;
  if (p->state != 0 || !p->needInitDic) { lzb_set_error("Lzma2Dec_DecodeToDic: Lzma2Dec_Init was not called"); return SZ_ERROR_UNSUPPORTED; }
  if (dicLimit > d->dicBufSize) return SZ_ERROR_PARAM;
  SizeT dl = dicLimit, sl = inSize;
  r = Lzma2DecodeParallel(nullptr, d->dic, &dl, src, &sl, lzma2_prop_from_dict(d->prop.dicSize), finishMode, status);
  if (r == SZ_ERROR_FAIL || r == SZ_ERROR_MEM) return r;
  ctl_of(d)->used = 1;
  d->needFlush = 0;
  d->dicPos = dl;
  d->processedPos = (UInt32)dl;
  *srcLen = sl;
  return r;
}

extern "C" SRes Lzma2Dec_DecodeToBuf(CLzma2Dec *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                                     ELzmaFinishMode finishMode, ELzmaStatus *status)
{
  /* reference Lzma2Dec.c:291-328, one slice (output must fit the dictionary buffer) */
  const SizeT outSize = *destLen;
  SizeT inSize = *srcLen;
  *srcLen = *destLen = 0;
  if (outSize > p->decoder.dicBufSize)
  {
    lzb_set_error("Lzma2Dec_DecodeToBuf: output larger than the dictionary buffer needs the resumable path (next round)");
    *status = LZMA_STATUS_NOT_SPECIFIED;
    return SZ_ERROR_UNSUPPORTED;
  }
  SRes r = Lzma2Dec_DecodeToDic(p, outSize, src, &inSize, finishMode, status);
  *srcLen = inSize;
  if (r == SZ_ERROR_UNSUPPORTED || r == SZ_ERROR_PARAM || r == SZ_ERROR_FAIL) return r;
  memcpy(dest, p->decoder.dic, p->decoder.dicPos);
  *destLen = p->decoder.dicPos;
  return r;
}
