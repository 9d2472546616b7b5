/* lzb_state.cu -- the SDK's stateful interfaces on the GPU path: the dictionary / buffer interfaces of
 * LzmaDec.h (LzmaDec_DecodeToDic, LzmaDec_DecodeToBuf) and the state interface of Lzma2Dec.h.
 *
 * A decoder object is resumable exactly like the reference's: any call may stop at any input or
 * output byte (NEEDS_MORE_INPUT, output limit) and the next call continues.  Division of labour:
 *
 *   host   CLzmaDec / CLzma2Dec keep every scalar field of the reference with the reference's meaning
 *          (range, code, dicPos, processedPos, checkDicSize, state, reps, remainLen, needFlush,
 *          needInitState, tempBuf/tempBufSize; packSize, unpackSize, state, control, needInit*) -- the
 *          caller may read and poke them as it does with the reference (7zDec.c:135-157).
 *   device the probability model and a mirror of the dictionary live in device memory between calls
 *          (CLzmaDec.probs points at a small host control block that owns them).
 *
 * Each call uploads the scalars + the new input, runs ONE stateful unit (csrc/lzb_core.cuh, flag
 * LZB_UNIT_STATEFUL) and downloads scalars + the newly produced dictionary range.  Instead of the
 * reference's tempBuf branch (LzmaDec.c:804-833) the bytes parked in tempBuf are re-presented in front
 * of the new input, which decodes the same symbols.  A single stream still runs on a single warp: this
 * interface exists for drop-in completeness, the throughput is in the batched calls.
 */
#include "lzb_host.h"

#include <stdlib.h>
#include <string.h>

#include <algorithm>

struct LzbCtl
{
  uint64_t magic;
  uint8_t *d_state; size_t state_bytes;
  uint8_t *d_dic;   size_t dic_cap;
  uint8_t *d_in;    size_t in_cap;      /* device block of a call: [LzbUnit | LzbResult | parked bytes + input] */
  uint8_t *h_stage; size_t stage_cap;   /* page-locked: [LzbStreamState | LzbResult | pad | the same block]   */
};
static const size_t kStageStateAt = 0, kStageResultAt = 128, kStageBlockAt = 192;
static const size_t kDicGuessBytes = (size_t)1 << 20;
#define LZB_CTL_MAGIC 0x4C5A4243544C3032ull

static LzbCtl *ctl_of(CLzmaDec *p) { return (LzbCtl *)(void *)p->probs; }

static void ctl_release_device(LzbCtl *c)
{
  if (!c || c->magic != LZB_CTL_MAGIC) return;
  cudaFree(c->d_state); cudaFree(c->d_dic); cudaFree(c->d_in);
  if (c->h_stage) cudaFreeHost(c->h_stage);
  c->d_state = c->d_dic = c->d_in = c->h_stage = nullptr;
  c->state_bytes = c->dic_cap = c->in_cap = c->stage_cap = 0;
}

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
  if (p->probs) ctl_release_device(ctl_of(p));
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
}

template <class T> static bool grow_dev(T *&ptr, size_t &cap, size_t need, const char *what)
{
  if (need <= cap && ptr) return true;
  cudaFree(ptr);
  ptr = nullptr; cap = 0;
  if (!lzb_cuda_ok(cudaMalloc((void **)&ptr, need + 64), what)) return false;
  /* zeroed: a call may fetch dictionary bytes the decoder has not written yet (the speculative part of the download),
   * and nothing an earlier user of this device memory left behind may reach the caller */
  /* (cudaMemset runs on the legacy default stream, which the context's non-blocking streams do not wait for: it
   * must have finished before anything is uploaded into the new buffer) */
  if (!lzb_cuda_ok(cudaMemset(ptr, 0, need + 64), what) || !lzb_cuda_ok(cudaDeviceSynchronize(), what)) return false;
  cap = need;
  return true;
}

/* One stateful step.  l2 == NULL: raw LZMA (LzmaDec_DecodeToDic); else the LZMA2 layer as well. */
static SRes step(CLzmaDec *p, CLzma2Dec *l2, SizeT dicLimit, const Byte *src, SizeT *srcLen,
                 ELzmaFinishMode finishMode, ELzmaStatus *status, const char *who)
{
  const SizeT inSize = *srcLen;
  *srcLen = 0;
  *status = LZMA_STATUS_NOT_SPECIFIED;
  if (!p->probs || ctl_of(p)->magic != LZB_CTL_MAGIC) { lzb_set_error("%s: decoder not allocated", who); return SZ_ERROR_PARAM; }
  if (!p->dic || dicLimit > p->dicBufSize) { lzb_set_error("%s: dicLimit beyond dicBufSize", who); return SZ_ERROR_PARAM; }
  if (p->dicBufSize > 0xFFFFFFFFull || inSize > 0xFFFFFF00ull) { lzb_set_error("%s: buffers above 4 GiB", who); return SZ_ERROR_UNSUPPORTED; }
  LzbContext *c = lzb_default_context();
  if (!c) return SZ_ERROR_FAIL;
  if (!lzb_cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return SZ_ERROR_FAIL;
  LzbCtl *ctl = ctl_of(p);

  const unsigned mlc = l2 ? 4 : p->prop.lc, mlp = l2 ? 0 : p->prop.lp, mpb = l2 ? 4 : p->prop.pb;
  const size_t state_bytes = sizeof(LzbStreamState) + (size_t)lzb_num_probs(mlc, mlp, mpb) * 2;
  const unsigned prefix = p->tempBufSize;
  const size_t total_in = (size_t)prefix + inSize;
  const size_t block_bytes = LZB_STATEFUL_INPUT_AT + total_in + 16;
  if (!grow_dev(ctl->d_state, ctl->state_bytes, state_bytes, "cudaMalloc decoder state") ||
      !grow_dev(ctl->d_dic, ctl->dic_cap, p->dicBufSize, "cudaMalloc dictionary mirror") ||
      !grow_dev(ctl->d_in, ctl->in_cap, block_bytes, "cudaMalloc input staging"))
    return SZ_ERROR_MEM;
  if (kStageBlockAt + block_bytes > ctl->stage_cap)
  {
    if (ctl->h_stage) cudaFreeHost(ctl->h_stage);
    ctl->h_stage = nullptr; ctl->stage_cap = 0;
    const size_t want = kStageBlockAt + block_bytes + block_bytes / 4 + 4096;
    if (!lzb_cuda_ok(cudaHostAlloc((void **)&ctl->h_stage, want, cudaHostAllocDefault), "cudaHostAlloc decoder staging")) return SZ_ERROR_MEM;
    ctl->stage_cap = want;
  }

  LzbStreamState &st = *reinterpret_cast<LzbStreamState *>(ctl->h_stage + kStageStateAt);
  LzbResult &x = *reinterpret_cast<LzbResult *>(ctl->h_stage + kStageResultAt);
  uint8_t *block = ctl->h_stage + kStageBlockAt;
  memset(&st, 0, sizeof st);
  st.range = p->range; st.code = p->code; st.state = p->state;
  for (int i = 0; i < 4; i++) st.rep[i] = p->reps[i];
  st.total_pos = p->processedPos; st.dict_full = p->checkDicSize; st.pending = p->remainLen;
  st.flags = (p->needFlush ? 1u : 0u) | (p->needInitState ? 2u : 0u);
  st.dic_pos = (uint32_t)p->dicPos;
  st.held = prefix;
  st.lc = p->prop.lc; st.lp = p->prop.lp; st.pb = p->prop.pb;
  if (l2)
  {
    st.l2_phase = (uint32_t)l2->state; st.l2_control = l2->control;
    st.l2_pack_left = l2->packSize; st.l2_unpack_left = l2->unpackSize;
    st.l2_flags = (l2->needInitDic ? 1u : 0u) | (l2->needInitState ? 2u : 0u) | (l2->needInitProp ? 4u : 0u);
  }
  LzbUnit &u = *reinterpret_cast<LzbUnit *>(block);
  memset(block, 0, LZB_STATEFUL_INPUT_AT);
  u.src_len = (uint32_t)total_in; u.dst_cap = (uint32_t)dicLimit; u.dict_size = p->prop.dicSize;
  u.codec = l2 ? LZB_CODEC_LZMA2 : LZB_CODEC_LZMA;
  u.lc = (uint8_t)mlc; u.lp = (uint8_t)mlp; u.pb = (uint8_t)mpb;      /* model sizing (LZMA2: worst case) */
  u.finish = (uint8_t)finishMode;
  u.flags = LZB_UNIT_STATEFUL;
  u.filter_param = (uint32_t)p->dicBufSize;
  u.expect = (uint64_t)(uintptr_t)ctl->d_state;
  if (prefix) memcpy(block + LZB_STATEFUL_INPUT_AT, p->tempBuf, prefix);
  if (inSize) memcpy(block + LZB_STATEFUL_INPUT_AT + prefix, src, inSize);
  memset(block + LZB_STATEFUL_INPUT_AT + total_in, 0, 16);

  /* one queue, one wait: scalars + block up, launch, scalars + result + the first `guess` new dictionary bytes down */
  const SizeT old_pos = p->dicPos;
  const size_t guess = old_pos < dicLimit ? std::min<size_t>(dicLimit - old_pos, kDicGuessBytes) : 0;
  SRes r = lzb_run_stateful(c, ctl->d_in, block, block_bytes, &u, ctl->d_state, &st, &x, ctl->d_dic, p->dic, old_pos, guess);
  if (r != SZ_OK) return r;
  if (st.dic_pos > old_pos + guess &&
      !lzb_cuda_ok(cudaMemcpy(p->dic + old_pos + guess, ctl->d_dic + old_pos + guess, st.dic_pos - old_pos - guess, cudaMemcpyDeviceToHost), "D2H dictionary"))
    return SZ_ERROR_FAIL;

  /* scalars back into the reference's fields */
  p->range = st.range; p->code = st.code; p->state = st.state;
  for (int i = 0; i < 4; i++) p->reps[i] = st.rep[i];
  p->processedPos = st.total_pos; p->checkDicSize = st.dict_full; p->remainLen = st.pending;
  p->needFlush = (st.flags & 1) ? 1 : 0; p->needInitState = (st.flags & 2) ? 1 : 0;
  p->dicPos = st.dic_pos;
  p->prop.lc = st.lc; p->prop.lp = st.lp; p->prop.pb = st.pb;
  if (l2)
  {
    l2->state = (int)st.l2_phase; l2->control = (Byte)st.l2_control;
    l2->packSize = st.l2_pack_left; l2->unpackSize = st.l2_unpack_left;
    l2->needInitDic = (st.l2_flags & 1) ? 1 : 0; l2->needInitState = (st.l2_flags & 2) ? 1 : 0; l2->needInitProp = (st.l2_flags & 4) ? 1 : 0;
  }

  /* input accounting: what the reference parks in tempBuf (LzmaDec.c:734-740,779-786) */
  const size_t used = x.src_used;
  if (x.status == LZMA_STATUS_NEEDS_MORE_INPUT && st.held > 0 && used == total_in)
  {
    Byte keep[LZMA_REQUIRED_INPUT_MAX];
    const unsigned h = st.held;                           /* <= 19 */
    for (unsigned i = 0; i < h; i++)
    {
      const size_t at = total_in - h + i;                 /* index in (tempBuf || src) */
      keep[i] = at < prefix ? p->tempBuf[at] : src[at - prefix];
    }
    memcpy(p->tempBuf, keep, h);
    p->tempBufSize = h;
    *srcLen = inSize;
  }
  else if (used >= prefix)
  {
    p->tempBufSize = 0;
    *srcLen = used - prefix;
  }
  else
    *srcLen = 0;                                          /* nothing consumed: parked bytes stay parked */
  *status = (ELzmaStatus)(x.status == LZB_ST_INHERIT ? LZMA_STATUS_NOT_SPECIFIED : x.status);
  return x.res;
}

extern "C" SRes LzmaDec_DecodeToDic(CLzmaDec *p, SizeT dicLimit, const Byte *src, SizeT *srcLen,
                                    ELzmaFinishMode finishMode, ELzmaStatus *status)
{
  return step(p, nullptr, dicLimit, src, srcLen, finishMode, status, "LzmaDec_DecodeToDic");
}

/* The buffer interface over the dictionary interface (reference LzmaDec.c:840-878, Lzma2Dec.c:291-328): the decoder
 * writes into its ring dictionary, every slice it produces is copied out to the caller.  A slice ends where the ring
 * ends or where the caller's buffer ends; only the second is an end the caller's finish mode applies to. */
template <class ToDic>
static SRes decode_to_buf(CLzmaDec *d, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                          ELzmaFinishMode finishMode, ELzmaStatus *status, ToDic to_dic)
{
  Byte *out = dest, *const out_end = dest + *destLen;
  const Byte *in = src, *const in_end = src + *srcLen;
  SRes res;
  for (;;)
  {
    if (d->dicPos == d->dicBufSize) d->dicPos = 0;                      /* the ring wraps */
    const SizeT from = d->dicPos;
    const SizeT wanted = (SizeT)(out_end - out), ring_left = d->dicBufSize - from;
    const bool to_callers_end = wanted <= ring_left;
    SizeT used = (SizeT)(in_end - in);
    res = to_dic(to_callers_end ? from + wanted : d->dicBufSize, in, &used, to_callers_end ? finishMode : LZMA_FINISH_ANY, status);
    in += used;
    const SizeT made = d->dicPos - from;
    memcpy(out, d->dic + from, made);
    out += made;
    if (res != SZ_OK || made == 0 || out == out_end) break;
  }
  *destLen = (SizeT)(out - dest);
  *srcLen = (SizeT)(in - src);
  return res;
}

extern "C" SRes LzmaDec_DecodeToBuf(CLzmaDec *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                                    ELzmaFinishMode finishMode, ELzmaStatus *status)
{
  return decode_to_buf(p, dest, destLen, src, srcLen, finishMode, status,
                       [p](SizeT limit, const Byte *in, SizeT *inLen, ELzmaFinishMode mode, ELzmaStatus *st)
                       { return LzmaDec_DecodeToDic(p, limit, in, inLen, mode, st); });
}

/* ------------------------------------------------------------------ LZMA2 state interface ----- */
static void lzma2_old_props(Byte prop, Byte *p5)
{
  /* reference Lzma2Dec.c:61-73 */
  const UInt32 dict = prop == 40 ? 0xFFFFFFFFu : ((UInt32)(2 | (prop & 1)) << (prop / 2 + 11));
  p5[0] = 4;
  p5[1] = (Byte)dict; p5[2] = (Byte)(dict >> 8); p5[3] = (Byte)(dict >> 16); p5[4] = (Byte)(dict >> 24);
}

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
  /* A decoder right after Lzma2Dec_Init that is handed a buffer holding many dictionary-reset groups is
   * the 7zDec.c:175-222 whole-folder case: decode the groups in parallel (lzb_containers.cu).  Anything
   * else -- partial input, continuation -- takes the resumable single-unit path. */
  if (p->state == 0 && p->needInitDic && p->needInitState && p->needInitProp && d->needFlush && d->dicPos == 0 &&
      d->tempBufSize == 0 && d->probs && d->dic && dicLimit <= d->dicBufSize && *srcLen >= (1u << 20))
  {
    size_t n = 0, end = 0;
    int eos = 0;
    LzbUnit probe[2];
    const Byte prop = lzma2_prop_from_dict(d->prop.dicSize);
    if (Lzma2_IndexUnits(src, *srcLen, prop, probe, 2, &n, &end, &eos) == SZ_OK && n >= 2 && eos)
    {
      SizeT dl = dicLimit, sl = *srcLen;
      SRes r = Lzma2DecodeParallel(nullptr, d->dic, &dl, src, &sl, prop, finishMode, status);
      if (r == SZ_OK && *status == LZMA_STATUS_FINISHED_WITH_MARK)
      {
        *srcLen = sl;
        d->dicPos = dl; d->processedPos = (UInt32)dl; d->needFlush = 0;
        p->state = 8 /* LZMA2_STATE_FINISHED */;
        p->needInitDic = p->needInitState = p->needInitProp = 0;
        return r;
      }
      /* not a clean whole-stream decode: fall through to the sequential path, which reproduces the
       * reference's exact stopping point and leaves a resumable object behind */
    }
  }
  return step(d, p, dicLimit, src, srcLen, finishMode, status, "Lzma2Dec_DecodeToDic");
}

extern "C" SRes Lzma2Dec_DecodeToBuf(CLzma2Dec *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                                     ELzmaFinishMode finishMode, ELzmaStatus *status)
{
  return decode_to_buf(&p->decoder, dest, destLen, src, srcLen, finishMode, status,
                       [p](SizeT limit, const Byte *in, SizeT *inLen, ELzmaFinishMode mode, ELzmaStatus *st)
                       { return Lzma2Dec_DecodeToDic(p, limit, in, inLen, mode, st); });
}
