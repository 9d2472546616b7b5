/* ref_harness.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Thin driver around the UNMODIFIED reference sources (compiled where they lie,
 * /root/reference/<name>.c, into oracle/_ref/libref.so by oracle/Makefile).  Nothing
 * here re-implements the algorithm; every function only sequences reference API
 * calls the way the reference's own callers do:
 *   - one-call raw LZMA          : LzmaEncode / LzmaDecode      (LzmaLib.c:15-46)
 *   - LZMA2 state interface      : 7zDec.c:175-222 pattern (never Lzma2Decode,
 *                                  which skips Lzma2Dec_Init -- Lzma2Dec.c:330-356)
 *   - zlib-like buffer interface : LzmaUtil.c:55-106 pattern
 *   - xz container               : XzEnc.c:41-160 writers + XzDec.c:604 unpacker
 *   - 7z archive                 : SzArEx_Open + SzArEx_Extract per file (7zIn.c:1314-1405)
 *   - pthread "one unit per core" decode loop = the CPU baseline of BASELINE.md §2
 *
 * Used by: tests/ (differential pinning of oracle/ and of the CUDA path),
 * tools/make_golden.py (fixture generation), bench.py (cpu_baseline leg,
 * --impl reference, synthetic-corpus compression).  The product library
 * (lzma-sdk-zliblike_b200/lib/liblzb200.so) never links or loads this.
 */
#include <pthread.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "7z.h"
#include "7zCrc.h"
#include "7zStream.h"
#include "Alloc.h"
#include "Bra.h"
#include "Delta.h"
#include "Lzma2Dec.h"
#include "Lzma2Enc.h"
#include "LzmaDec.h"
#include "LzmaEnc.h"
#include "Xz.h"
#include "XzCrc64.h"
#include "XzEnc.h"

/* the fork prints one line per LzmaDec_AllocateProbs (LzmaDec.c:944-946); the
 * Makefile renames that printf to this sink with -Dprintf=refh_quiet_printf. */
int refh_quiet_printf(const char *fmt, ...) { (void)fmt; return 0; }

/* non-static writers that XzEnc.h does not declare (XzEnc.c:41,53,79,124) */
SRes Xz_WriteHeader(CXzStreamFlags f, ISeqOutStream *s);
SRes XzBlock_WriteHeader(const CXzBlock *p, ISeqOutStream *s);
SRes Xz_WriteFooter(CXzStream *p, ISeqOutStream *s);
SRes Xz_AddIndexRecord(CXzStream *p, UInt64 unpackSize, UInt64 totalSize, ISzAlloc *alloc);

static void *h_alloc(void *p, size_t n) { (void)p; return n ? malloc(n) : 0; }
static void h_free(void *p, void *a) { (void)p; free(a); }
static ISzAlloc g_alloc = { h_alloc, h_free };

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int g_inited;
void refh_init(void)
{
  if (!g_inited) { CrcGenerateTable(); Crc64GenerateTable(); g_inited = 1; }
}

/* ---------------- in-memory streams (Types.h:137-154) ---------------- */
typedef struct { ISeqInStream s; const Byte *p; size_t left; } MemIn;
typedef struct { ISeqOutStream s; Byte *p; size_t cap, pos; int overflow; } MemOut;

static SRes memin_read(void *pp, void *buf, size_t *size)
{
  MemIn *m = (MemIn *)pp;
  size_t n = *size < m->left ? *size : m->left;
  memcpy(buf, m->p, n);
  m->p += n; m->left -= n; *size = n;
  return SZ_OK;
}
static size_t memout_write(void *pp, const void *buf, size_t size)
{
  MemOut *m = (MemOut *)pp;
  if (m->cap - m->pos < size) { m->overflow = 1; return 0; }
  memcpy(m->p + m->pos, buf, size);
  m->pos += size;
  return size;
}
static void memin_init(MemIn *m, const Byte *p, size_t n) { m->s.Read = memin_read; m->p = p; m->left = n; }
static void memout_init(MemOut *m, Byte *p, size_t cap) { m->s.Write = memout_write; m->p = p; m->cap = cap; m->pos = 0; m->overflow = 0; }

/* ---------------- raw LZMA ---------------- */
int refh_lzma_encode(const uint8_t *src, size_t n, uint8_t *dst, size_t *dstLen, uint8_t *props5,
                     int level, uint32_t dictSize, int lc, int lp, int pb, int fb, int endMark)
{
  CLzmaEncProps ep;
  SizeT psz = LZMA_PROPS_SIZE;
  LzmaEncProps_Init(&ep);
  ep.level = level; ep.dictSize = dictSize; ep.lc = lc; ep.lp = lp; ep.pb = pb;
  if (fb > 0) ep.fb = fb;
  ep.numThreads = 1;
  return LzmaEncode(dst, dstLen, src, n, &ep, props5, &psz, endMark, NULL, &g_alloc, &g_alloc);
}

int refh_lzma_decode(uint8_t *dst, size_t *dstLen, const uint8_t *src, size_t *srcLen,
                     const uint8_t *props5, int finishMode, int *status)
{
  ELzmaStatus st = LZMA_STATUS_NOT_SPECIFIED;
  SRes r = LzmaDecode(dst, dstLen, src, srcLen, props5, LZMA_PROPS_SIZE, (ELzmaFinishMode)finishMode, &st, &g_alloc);
  *status = (int)st;
  return r;
}

/* zlib-like buffer interface, fed in_chunk bytes at a time, drained out_chunk at a
 * time (LzmaUtil.c:55-106).  dict lives inside the decoder (LzmaDec_Allocate). */
int refh_lzma_decode_tobuf(uint8_t *dst, size_t dstCap, size_t *dstLen, const uint8_t *src, size_t srcLen,
                           size_t *srcUsed, const uint8_t *props5, size_t in_chunk, size_t out_chunk,
                           int finishModeAtEnd, int *status)
{
  CLzmaDec d;
  SRes r;
  size_t ip = 0, op = 0;
  ELzmaStatus st = LZMA_STATUS_NOT_SPECIFIED;
  LzmaDec_Construct(&d);
  r = LzmaDec_Allocate(&d, props5, LZMA_PROPS_SIZE, &g_alloc);
  if (r != SZ_OK) { *dstLen = 0; *srcUsed = 0; *status = 0; return r; }
  LzmaDec_Init(&d);
  for (;;)
  {
    size_t inAvail = srcLen - ip < in_chunk ? srcLen - ip : in_chunk;
    size_t inLeft = inAvail, inPos = 0;
    int progressed = 0;
    for (;;)
    {
      SizeT il = inLeft, ol = dstCap - op < out_chunk ? dstCap - op : out_chunk;
      ELzmaFinishMode fm = (ol == dstCap - op) ? (ELzmaFinishMode)finishModeAtEnd : LZMA_FINISH_ANY;
      r = LzmaDec_DecodeToBuf(&d, dst + op, &ol, src + ip + inPos, &il, fm, &st);
      inPos += il; inLeft -= il; op += ol;
      if (il || ol) progressed = 1;
      if (r != SZ_OK || st == LZMA_STATUS_FINISHED_WITH_MARK) goto done;
      if (op == dstCap) goto done;
      if (inLeft == 0 || (il == 0 && ol == 0)) break;
    }
    ip += inPos;
    inPos = 0;
    if (ip == srcLen && !progressed) break;
    if (ip == srcLen && st == LZMA_STATUS_NEEDS_MORE_INPUT) break;
    continue;
done:
    ip += inPos;
    break;
  }
  *dstLen = op; *srcUsed = ip; *status = (int)st;
  LzmaDec_Free(&d, &g_alloc);
  return r;
}

/* ---------------- LZMA2 ---------------- */
/* One Lzma2Enc_Encode call over src[0..n): first chunk resets the dictionary,
 * output ends with the 0x00 end-of-stream byte (Lzma2Enc.c:302-361). */
int refh_lzma2_encode(const uint8_t *src, size_t n, uint8_t *dst, size_t *dstLen,
                      int level, uint32_t dictSize, int lc, int lp, int pb, uint8_t *propByte)
{
  CLzma2EncProps p2;
  CLzma2EncHandle h;
  MemIn in; MemOut out;
  SRes r;
  h = Lzma2Enc_Create(&g_alloc, &g_alloc);
  if (!h) return SZ_ERROR_MEM;
  Lzma2EncProps_Init(&p2);
  p2.lzmaProps.level = level;
  p2.lzmaProps.dictSize = dictSize;
  if (lc >= 0) p2.lzmaProps.lc = lc;
  if (lp >= 0) p2.lzmaProps.lp = lp;
  if (pb >= 0) p2.lzmaProps.pb = pb;
  p2.lzmaProps.numThreads = 1;
  p2.numBlockThreads = 1; p2.numTotalThreads = 1;
  r = Lzma2Enc_SetProps(h, &p2);
  if (r == SZ_OK)
  {
    *propByte = Lzma2Enc_WriteProperties(h);
    memin_init(&in, src, n);
    memout_init(&out, dst, *dstLen);
    r = Lzma2Enc_Encode(h, &out.s, &in.s, NULL);
    if (out.overflow && r == SZ_OK) r = SZ_ERROR_OUTPUT_EOF;
    *dstLen = out.pos;
  }
  Lzma2Enc_Destroy(h);
  return r;
}

/* State interface straight into the caller's buffer (7zDec.c:175-222). */
int refh_lzma2_decode(uint8_t *dst, size_t dstCap, size_t *dstLen, const uint8_t *src, size_t *srcLen,
                      uint8_t prop, int finishMode, int *status)
{
  CLzma2Dec d;
  ELzmaStatus st = LZMA_STATUS_NOT_SPECIFIED;
  SRes r;
  Lzma2Dec_Construct(&d);
  r = Lzma2Dec_AllocateProbs(&d, prop, &g_alloc);
  if (r != SZ_OK) { *dstLen = 0; *srcLen = 0; *status = 0; return r; }
  d.decoder.dic = dst;
  d.decoder.dicBufSize = dstCap;
  Lzma2Dec_Init(&d);
  r = Lzma2Dec_DecodeToDic(&d, dstCap, src, srcLen, (ELzmaFinishMode)finishMode, &st);
  *dstLen = d.decoder.dicPos;
  *status = (int)st;
  Lzma2Dec_FreeProbs(&d, &g_alloc);
  return r;
}

/* zlib-like buffer interface of the LZMA2 layer, fed/drained in chunks (XzDec.c:285-298 pattern) */
int refh_lzma2_decode_tobuf(uint8_t *dst, size_t dstCap, size_t *dstLen, const uint8_t *src, size_t srcLen,
                            size_t *srcUsed, uint8_t prop, size_t in_chunk, size_t out_chunk,
                            int finishModeAtEnd, int *status)
{
  CLzma2Dec d;
  SRes r;
  size_t ip = 0, op = 0;
  ELzmaStatus st = LZMA_STATUS_NOT_SPECIFIED;
  Lzma2Dec_Construct(&d);
  r = Lzma2Dec_Allocate(&d, prop, &g_alloc);
  if (r != SZ_OK) { *dstLen = 0; *srcUsed = 0; *status = 0; return r; }
  Lzma2Dec_Init(&d);
  for (;;)
  {
    size_t inAvail = srcLen - ip < in_chunk ? srcLen - ip : in_chunk;
    size_t inLeft = inAvail, inPos = 0;
    int progressed = 0, done = 0;
    for (;;)
    {
      SizeT il = inLeft, ol = dstCap - op < out_chunk ? dstCap - op : out_chunk;
      ELzmaFinishMode fm = (ol == dstCap - op) ? (ELzmaFinishMode)finishModeAtEnd : LZMA_FINISH_ANY;
      r = Lzma2Dec_DecodeToBuf(&d, dst + op, &ol, src + ip + inPos, &il, fm, &st);
      inPos += il; inLeft -= il; op += ol;
      if (il || ol) progressed = 1;
      if (r != SZ_OK || st == LZMA_STATUS_FINISHED_WITH_MARK || op == dstCap) { done = 1; break; }
      if (inLeft == 0 || (il == 0 && ol == 0)) break;
    }
    ip += inPos;
    if (done) break;
    if (ip == srcLen && (!progressed || st == LZMA_STATUS_NEEDS_MORE_INPUT)) break;
  }
  *dstLen = op; *srcUsed = ip; *status = (int)st;
  Lzma2Dec_Free(&d, &g_alloc);
  return r;
}

/* ---------------- multi-threaded helpers (pthread) ---------------- */
typedef struct
{
  int kind;                 /* 0 raw LZMA one-call, 1 LZMA2 unit */
  size_t n;
  const uint8_t *src; const uint64_t *srcOff; const uint32_t *srcLen;
  uint8_t *dst; const uint64_t *dstOff; const uint32_t *dstCap;
  const uint8_t *props;     /* raw: 5 bytes shared; lzma2: 1 byte shared */
  int finishMode;
  int32_t *res; int32_t *status; uint32_t *srcUsed; uint32_t *dstDone;
  int scratch_out;          /* 1: decode into a per-thread scratch buffer (timing only) */
  size_t scratch_cap;
  volatile size_t next;
} DecJob;

static void *dec_worker(void *arg)
{
  DecJob *j = (DecJob *)arg;
  CLzmaDec d1; CLzma2Dec d2;
  uint8_t *scratch = j->scratch_out ? (uint8_t *)malloc(j->scratch_cap ? j->scratch_cap : 1) : NULL;
  /* probabilities allocated ONCE per thread */
  LzmaDec_Construct(&d1); Lzma2Dec_Construct(&d2);
  if (j->kind == 0) LzmaDec_AllocateProbs(&d1, j->props, LZMA_PROPS_SIZE, &g_alloc);
  else Lzma2Dec_AllocateProbs(&d2, j->props[0], &g_alloc);
  for (;;)
  {
    size_t u = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    SizeT sl; ELzmaStatus st = LZMA_STATUS_NOT_SPECIFIED; SRes r; size_t produced;
    uint8_t *out;
    if (u >= j->n) break;
    sl = j->srcLen[u];
    out = scratch ? scratch : j->dst + j->dstOff[u];
    if (j->kind == 0)
    {
      /* LzmaDecode body (LzmaDec.c:972-1002) minus the per-call malloc */
      if (sl < 5) { r = SZ_ERROR_INPUT_EOF; sl = 0; produced = 0; }
      else
      {
        d1.dic = out; d1.dicBufSize = j->dstCap[u];
        LzmaDec_Init(&d1);
        r = LzmaDec_DecodeToDic(&d1, j->dstCap[u], j->src + j->srcOff[u], &sl, (ELzmaFinishMode)j->finishMode, &st);
        if (r == SZ_OK && st == LZMA_STATUS_NEEDS_MORE_INPUT) r = SZ_ERROR_INPUT_EOF;
        produced = d1.dicPos;
      }
    }
    else
    {
      d2.decoder.dic = out; d2.decoder.dicBufSize = j->dstCap[u];
      Lzma2Dec_Init(&d2);
      r = Lzma2Dec_DecodeToDic(&d2, j->dstCap[u], j->src + j->srcOff[u], &sl, (ELzmaFinishMode)j->finishMode, &st);
      produced = d2.decoder.dicPos;
    }
    if (j->res) j->res[u] = r;
    if (j->status) j->status[u] = (int32_t)st;
    if (j->srcUsed) j->srcUsed[u] = (uint32_t)sl;
    if (j->dstDone) j->dstDone[u] = (uint32_t)produced;
  }
  if (j->kind == 0) LzmaDec_FreeProbs(&d1, &g_alloc); else Lzma2Dec_FreeProbs(&d2, &g_alloc);
  free(scratch);
  return NULL;
}

/* Returns wall seconds.  kind 0 = raw LZMA streams (props5 shared), 1 = LZMA2 units. */
double refh_decode_batch_mt(int kind, size_t n, const uint8_t *src, const uint64_t *srcOff, const uint32_t *srcLen,
                            uint8_t *dst, const uint64_t *dstOff, const uint32_t *dstCap, const uint8_t *props,
                            int finishMode, int32_t *res, int32_t *status, uint32_t *srcUsed, uint32_t *dstDone,
                            int threads, int scratch_out)
{
  DecJob j;
  pthread_t *th;
  int i;
  double t0, t1;
  size_t k, mx = 0;
  if (threads < 1) threads = 1;
  for (k = 0; k < n; k++) if (dstCap[k] > mx) mx = dstCap[k];
  memset(&j, 0, sizeof j);
  j.kind = kind; j.n = n; j.src = src; j.srcOff = srcOff; j.srcLen = srcLen;
  j.dst = dst; j.dstOff = dstOff; j.dstCap = dstCap; j.props = props; j.finishMode = finishMode;
  j.res = res; j.status = status; j.srcUsed = srcUsed; j.dstDone = dstDone;
  j.scratch_out = scratch_out; j.scratch_cap = mx; j.next = 0;
  th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
  t0 = now_s();
  for (i = 0; i < threads; i++) pthread_create(&th[i], NULL, dec_worker, &j);
  for (i = 0; i < threads; i++) pthread_join(th[i], NULL);
  t1 = now_s();
  free(th);
  return t1 - t0;
}

typedef struct
{
  int kind; size_t n;
  const uint8_t *src; const uint64_t *srcOff; const uint32_t *srcLen;
  uint8_t *dst; const uint64_t *dstOff; const uint32_t *dstCap; uint32_t *dstLen;
  uint8_t *propsOut;
  int level; uint32_t dictSize; int lc, lp, pb, fb, endMark;
  int32_t *res;
  volatile size_t next;
} EncJob;

static void *enc_worker(void *arg)
{
  EncJob *j = (EncJob *)arg;
  for (;;)
  {
    size_t u = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    size_t dl;
    int r;
    if (u >= j->n) break;
    dl = j->dstCap[u];
    if (j->kind == 0)
    {
      uint8_t props[5];
      r = refh_lzma_encode(j->src + j->srcOff[u], j->srcLen[u], j->dst + j->dstOff[u], &dl, props,
                           j->level, j->dictSize, j->lc, j->lp, j->pb, j->fb, j->endMark);
      if (u == 0 && j->propsOut) memcpy(j->propsOut, props, 5);
    }
    else
    {
      uint8_t pb1 = 0;
      r = refh_lzma2_encode(j->src + j->srcOff[u], j->srcLen[u], j->dst + j->dstOff[u], &dl,
                            j->level, j->dictSize, j->lc, j->lp, j->pb, &pb1);
      if (u == 0 && j->propsOut) j->propsOut[0] = pb1;
    }
    j->dstLen[u] = (uint32_t)dl;
    if (j->res) j->res[u] = r;
  }
  return NULL;
}

/* Compress n independent slices on `threads` host threads with the reference encoder. */
double refh_encode_batch_mt(int kind, size_t n, const uint8_t *src, const uint64_t *srcOff, const uint32_t *srcLen,
                            uint8_t *dst, const uint64_t *dstOff, const uint32_t *dstCap, uint32_t *dstLen,
                            uint8_t *propsOut, int level, uint32_t dictSize, int lc, int lp, int pb, int fb,
                            int endMark, int32_t *res, int threads)
{
  EncJob j;
  pthread_t *th;
  int i;
  double t0, t1;
  if (threads < 1) threads = 1;
  memset(&j, 0, sizeof j);
  j.kind = kind; j.n = n; j.src = src; j.srcOff = srcOff; j.srcLen = srcLen;
  j.dst = dst; j.dstOff = dstOff; j.dstCap = dstCap; j.dstLen = dstLen; j.propsOut = propsOut;
  j.level = level; j.dictSize = dictSize; j.lc = lc; j.lp = lp; j.pb = pb; j.fb = fb; j.endMark = endMark;
  j.res = res; j.next = 0;
  th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
  t0 = now_s();
  for (i = 0; i < threads; i++) pthread_create(&th[i], NULL, enc_worker, &j);
  for (i = 0; i < threads; i++) pthread_join(th[i], NULL);
  t1 = now_s();
  free(th);
  return t1 - t0;
}

/* ---------------- xz container ---------------- */
/* Multi-block .xz writer composed from the reference's own low-level writers
 * (Xz_Encode itself can only emit one CRC32 block -- XzEnc.c:397-465).
 * filterId: 0 none, XZ_ID_X86 (4) or XZ_ID_Delta (3); filterParam: delta distance (1..256).
 * lzma2 payloads are supplied pre-encoded per block (packed[blk] of packedLen[blk], each a
 * complete LZMA2 stream incl. 0x00) so the slow encoder can run in parallel outside. */
int refh_xz_assemble(uint8_t *out, size_t *outLen, size_t nBlocks,
                     const uint8_t *packed, const uint64_t *packedOff, const uint32_t *packedLen,
                     const uint32_t *unpackLen, const uint64_t *checkVal /* crc32 or crc64 of ORIGINAL data */,
                     int checkType, int filterId, uint32_t filterParam, uint8_t lzma2Prop)
{
  CXzStream xz;
  MemOut mo;
  size_t b;
  SRes r;
  unsigned checkSize;
  refh_init();
  Xz_Construct(&xz);
  /* pre-allocate: Xz_AddIndexRecord drops records when it grows (XzEnc.c:139 -> Xz.c:32-37) */
  xz.blocks = (CXzBlockSizes *)malloc(sizeof(CXzBlockSizes) * (nBlocks + 1));
  xz.numBlocksAllocated = nBlocks + 1;
  xz.flags = (CXzStreamFlags)checkType;
  checkSize = XzFlags_GetCheckSize(xz.flags);
  memout_init(&mo, out, *outLen);
  r = Xz_WriteHeader(xz.flags, &mo.s);
  for (b = 0; r == SZ_OK && b < nBlocks; b++)
  {
    CXzBlock blk;
    size_t hdrStart = mo.pos, hdrSize, pad, i;
    int nf = 0;
    memset(&blk, 0, sizeof blk);
    if (filterId == XZ_ID_Delta)
    {
      blk.filters[nf].id = XZ_ID_Delta; blk.filters[nf].propsSize = 1;
      blk.filters[nf].props[0] = (Byte)(filterParam - 1); nf++;
    }
    else if (filterId != 0)
    {
      blk.filters[nf].id = (UInt64)filterId; blk.filters[nf].propsSize = 0; nf++;
    }
    blk.filters[nf].id = XZ_ID_LZMA2; blk.filters[nf].propsSize = 1; blk.filters[nf].props[0] = lzma2Prop; nf++;
    blk.flags = (Byte)(nf - 1);
    r = XzBlock_WriteHeader(&blk, &mo.s);
    if (r != SZ_OK) break;
    hdrSize = mo.pos - hdrStart;
    if (mo.s.Write(&mo, packed + packedOff[b], packedLen[b]) != packedLen[b]) { r = SZ_ERROR_OUTPUT_EOF; break; }
    pad = (4 - (packedLen[b] & 3)) & 3;
    for (i = 0; i < pad; i++) { Byte z = 0; if (mo.s.Write(&mo, &z, 1) != 1) { r = SZ_ERROR_OUTPUT_EOF; break; } }
    for (i = 0; i < checkSize; i++)
    {
      Byte c = (Byte)(checkVal[b] >> (8 * i));
      if (mo.s.Write(&mo, &c, 1) != 1) { r = SZ_ERROR_OUTPUT_EOF; break; }
    }
    if (r != SZ_OK) break;
    r = Xz_AddIndexRecord(&xz, unpackLen[b], hdrSize + packedLen[b] + checkSize, &g_alloc);
  }
  if (r == SZ_OK) r = Xz_WriteFooter(&xz, &mo.s);
  if (mo.overflow && r == SZ_OK) r = SZ_ERROR_OUTPUT_EOF;
  *outLen = mo.pos;
  free(xz.blocks);
  return r;
}

/* Whole-buffer XzUnpacker_Code loop (XzDec.c:604-870). */
int refh_xz_decode(uint8_t *dst, size_t *dstLen, const uint8_t *src, size_t *srcLen, int finishMode,
                   int *status, uint64_t *numBlocks, int *streamFinished)
{
  CXzUnpacker u;
  size_t ip = 0, op = 0, inSize = *srcLen, outCap = *dstLen;
  ECoderStatus st = CODER_STATUS_NOT_SPECIFIED;
  SRes r;
  refh_init();
  XzUnpacker_Create(&u, &g_alloc);
  for (;;)
  {
    SizeT sl = inSize - ip, dl = outCap - op;
    r = XzUnpacker_Code(&u, dst + op, &dl, src + ip, &sl, finishMode, &st);
    ip += sl; op += dl;
    if (r != SZ_OK) break;
    if (sl == 0 && dl == 0) break;
  }
  *srcLen = ip; *dstLen = op; *status = (int)st;
  *numBlocks = u.numBlocks;
  *streamFinished = XzUnpacker_IsStreamWasFinished(&u) ? 1 : 0;
  XzUnpacker_Free(&u);
  return r;
}

/* pthread baseline for config 4: each worker runs a full XzUnpacker over one single-block-run
 * .xz "member" (independent complete xz streams), one member per core. */
typedef struct
{
  size_t n; const uint8_t *src; const uint64_t *srcOff; const uint32_t *srcLen;
  uint8_t *dst; const uint64_t *dstOff; const uint32_t *dstCap; int32_t *res; uint32_t *dstDone;
  int scratch_out; size_t scratch_cap; volatile size_t next;
} XzJob;

static void *xz_worker(void *arg)
{
  XzJob *j = (XzJob *)arg;
  uint8_t *scratch = j->scratch_out ? (uint8_t *)malloc(j->scratch_cap ? j->scratch_cap : 1) : NULL;
  for (;;)
  {
    size_t u = __atomic_fetch_add(&j->next, 1, __ATOMIC_RELAXED);
    size_t sl, dl; int st, fin; uint64_t nb; int r;
    if (u >= j->n) break;
    sl = j->srcLen[u]; dl = j->dstCap[u];
    r = refh_xz_decode(scratch ? scratch : j->dst + j->dstOff[u], &dl, j->src + j->srcOff[u], &sl, CODER_FINISH_ANY, &st, &nb, &fin);
    if (r == SZ_OK && !fin) r = SZ_ERROR_INPUT_EOF;
    if (j->res) j->res[u] = r;
    if (j->dstDone) j->dstDone[u] = (uint32_t)dl;
  }
  free(scratch);
  return NULL;
}

double refh_xz_decode_batch_mt(size_t n, const uint8_t *src, const uint64_t *srcOff, const uint32_t *srcLen,
                               uint8_t *dst, const uint64_t *dstOff, const uint32_t *dstCap,
                               int32_t *res, uint32_t *dstDone, int threads, int scratch_out)
{
  XzJob j; pthread_t *th; int i; double t0, t1; size_t k, mx = 0;
  refh_init();
  if (threads < 1) threads = 1;
  for (k = 0; k < n; k++) if (dstCap[k] > mx) mx = dstCap[k];
  memset(&j, 0, sizeof j);
  j.n = n; j.src = src; j.srcOff = srcOff; j.srcLen = srcLen; j.dst = dst; j.dstOff = dstOff; j.dstCap = dstCap;
  j.res = res; j.dstDone = dstDone; j.scratch_out = scratch_out; j.scratch_cap = mx; j.next = 0;
  th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
  t0 = now_s();
  for (i = 0; i < threads; i++) pthread_create(&th[i], NULL, xz_worker, &j);
  for (i = 0; i < threads; i++) pthread_join(th[i], NULL);
  t1 = now_s();
  free(th);
  return t1 - t0;
}

/* ---------------- filters / checks (thin pass-throughs for ctypes) ---------------- */
size_t refh_x86_convert(uint8_t *data, size_t size, uint32_t ip, uint32_t *state, int encoding)
{
  return x86_Convert(data, size, ip, state, encoding);
}
/* Whole-buffer BCJ as the xz BraState applies it at end of input: the bytes the converter
 * leaves unprocessed are passed through unchanged (XzDec.c:186-191). */
void refh_x86_filter_whole(uint8_t *data, size_t size, uint32_t ip, int encoding)
{
  UInt32 st;
  x86_Convert_Init(st);
  (void)x86_Convert(data, size, ip, &st, encoding);
}
size_t refh_bra_convert(int id, uint8_t *data, size_t size, uint32_t ip, int encoding)
{
  switch (id)
  {
    case 5: return PPC_Convert(data, size, ip, encoding);
    case 6: return IA64_Convert(data, size, ip, encoding);
    case 7: return ARM_Convert(data, size, ip, encoding);
    case 8: return ARMT_Convert(data, size, ip, encoding);
    case 9: return SPARC_Convert(data, size, ip, encoding);
  }
  return 0;
}
void refh_delta(uint8_t *data, size_t size, unsigned delta, int encoding)
{
  Byte st[DELTA_STATE_SIZE];
  Delta_Init(st);
  if (encoding) Delta_Encode(st, delta, data, size); else Delta_Decode(st, delta, data, size);
}
uint32_t refh_crc32(const uint8_t *p, size_t n) { refh_init(); return CrcCalc(p, n); }
uint64_t refh_crc64(const uint8_t *p, size_t n) { refh_init(); return Crc64Calc(p, n); }

/* struct layout probes (SURVEY.md §8b ABI table) */
size_t refh_sizeof(int which)
{
  switch (which)
  {
    case 0: return sizeof(CLzmaDec);
    case 1: return sizeof(CLzma2Dec);
    case 2: return sizeof(CXzUnpacker);
    case 3: return sizeof(CLzmaProps);
    default: return 0;
  }
}

/* ---------------- 7z archive: open + extract every file (7zMain.c pattern, from memory) ---------------- */
/* the fork's file helpers reference two Win32-only openers (7zFile.h:36-37); never called here */
int InFile_OpenW(void *p, const void *name, int writeMode) { (void)p; (void)name; (void)writeMode; return 1; }
int OutFile_OpenW(void *p, const void *name, int writeMode) { (void)p; (void)name; (void)writeMode; return 1; }

typedef struct { ISeekInStream s; const uint8_t *data; size_t size, pos; } MemSeekIn;

static SRes mem_read(void *pp, void *buf, size_t *size)
{
  MemSeekIn *m = (MemSeekIn *)pp;
  size_t n = m->size - m->pos;
  if (n > *size) n = *size;
  memcpy(buf, m->data + m->pos, n);
  m->pos += n;
  *size = n;
  return SZ_OK;
}

static SRes mem_seek(void *pp, Int64 *pos, ESzSeek origin)
{
  MemSeekIn *m = (MemSeekIn *)pp;
  Int64 base = origin == SZ_SEEK_SET ? 0 : (origin == SZ_SEEK_CUR ? (Int64)m->pos : (Int64)m->size);
  Int64 t = base + *pos;
  if (t < 0) return SZ_ERROR_FAIL;
  m->pos = (size_t)t;
  *pos = t;
  return SZ_OK;
}

/* Returns SzArEx_Open's SRes.  For every file i < maxFiles: fileRes = SzArEx_Extract's SRes, fileSize,
 * fileFolder (0xFFFFFFFF: no stream), fileFlags (bit0 IsDir, bit1 HasStream, bit2 CrcDefined), fileCrc, and
 * the file's bytes appended to out (fileOff = where).  Names: UTF-16 code units of all names, 0-terminated. */
int refh_7z_extract_all(const uint8_t *arc, size_t n, uint8_t *out, size_t cap, uint32_t maxFiles,
                        uint32_t *numFiles, uint32_t *numFolders, int32_t *fileRes, uint64_t *fileOff,
                        uint64_t *fileSize, uint32_t *fileFolder, uint8_t *fileFlags, uint32_t *fileCrc,
                        uint16_t *names, size_t namesCap, size_t *namesLen)
{
  MemSeekIn mem;
  CLookToRead look;
  CSzArEx db;
  SRes res;
  size_t op = 0, np = 0;
  UInt32 i, block = 0xFFFFFFFF;
  Byte *cache = 0;
  size_t cacheSize = 0;
  refh_init();
  mem.s.Read = mem_read; mem.s.Seek = mem_seek; mem.data = arc; mem.size = n; mem.pos = 0;
  LookToRead_CreateVTable(&look, False);
  look.realStream = &mem.s;
  LookToRead_Init(&look);
  SzArEx_Init(&db);
  *numFiles = *numFolders = 0;
  if (namesLen) *namesLen = 0;
  res = SzArEx_Open(&db, &look.s, &g_alloc, &g_alloc);
  if (res != SZ_OK) return res;
  *numFiles = db.db.NumFiles; *numFolders = db.db.NumFolders;
  for (i = 0; i < db.db.NumFiles && i < maxFiles; i++)
  {
    const CSzFileItem *f = db.db.Files + i;
    size_t offset = 0, done = 0;
    SRes r = SzArEx_Extract(&db, &look.s, i, &block, &cache, &cacheSize, &offset, &done, &g_alloc, &g_alloc);
    fileRes[i] = r;
    fileSize[i] = f->Size;
    fileFolder[i] = db.FileIndexToFolderIndexMap[i];
    fileFlags[i] = (uint8_t)((f->IsDir ? 1 : 0) | (f->HasStream ? 2 : 0) | (f->CrcDefined ? 4 : 0));
    fileCrc[i] = f->Crc;
    fileOff[i] = op;
    if ((r == SZ_OK || r == SZ_ERROR_CRC) && cache && done != 0 && offset + done <= cacheSize && op + done <= cap)
    {
      memcpy(out + op, cache + offset, done);
      op += done;
    }
    if (r != SZ_OK)
    {
      /* 7zMain stops at the first failure; to get a verdict for EVERY file the solid-block cache is dropped
       * after a failure, so that the next file of the same folder decodes it again instead of reading the
       * half-written cache (7zIn.c:1345) */
      g_alloc.Free(&g_alloc, cache);
      cache = 0; cacheSize = 0; block = 0xFFFFFFFF;
    }
    if (names && db.FileNameOffsets)
    {
      size_t len = SzArEx_GetFileNameUtf16(&db, i, NULL);
      if (np + len <= namesCap) { SzArEx_GetFileNameUtf16(&db, i, names + np); np += len; }
    }
  }
  if (namesLen) *namesLen = np;
  g_alloc.Free(&g_alloc, cache);
  SzArEx_Free(&db, &g_alloc);
  return SZ_OK;
}
