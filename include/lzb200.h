/* lzb200.h -- C-ABI of liblzb200.so: the B200-native decode path behind the LZMA SDK 9.20 C API.
 *
 * Part 1 re-declares the reference's decode-side interface with identical names, argument meaning,
 * return codes and PUBLIC STRUCT LAYOUTS, so a caller of the reference links against this library
 * unchanged (each item cites the reference declaration it replaces).  Part 2 is the batched
 * extension -- the reference has no notion of it -- that exposes the unit-parallel GPU path
 * directly: many independent raw LZMA streams, LZMA2 dictionary-reset groups or xz blocks per call.
 *
 * Every entry point that decodes runs on the GPU.  There is no CPU fallback: without a CUDA device
 * the calls return SZ_ERROR_FAIL and Lzb_LastError() says why.
 */
#ifndef LZB200_H
#define LZB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ====================================================================================== */
/* Part 1 -- the SDK surface                                                               */
/* ====================================================================================== */

/* ---- basic types and result codes: reference Types.h:25-41,54-104,227-234 ---- */
typedef int SRes;
typedef unsigned char Byte;
typedef unsigned short UInt16;
typedef unsigned int UInt32;
typedef unsigned long long UInt64;
typedef size_t SizeT;
typedef int Bool;

#define SZ_OK 0
#define SZ_ERROR_DATA 1
#define SZ_ERROR_MEM 2
#define SZ_ERROR_CRC 3
#define SZ_ERROR_UNSUPPORTED 4
#define SZ_ERROR_PARAM 5
#define SZ_ERROR_INPUT_EOF 6
#define SZ_ERROR_OUTPUT_EOF 7
#define SZ_ERROR_READ 8
#define SZ_ERROR_WRITE 9
#define SZ_ERROR_PROGRESS 10
#define SZ_ERROR_FAIL 11
#define SZ_ERROR_THREAD 12
#define SZ_ERROR_ARCHIVE 16
#define SZ_ERROR_NO_ARCHIVE 17

typedef struct
{
  void *(*Alloc)(void *p, size_t size);
  void (*Free)(void *p, void *address); /* address can be 0 */
} ISzAlloc;

/* ---- LZMA decoder: reference LzmaDec.h ---- */
#define LZMA_PROPS_SIZE 5            /* LzmaDec.h:26 */
#define LZMA_REQUIRED_INPUT_MAX 20   /* LzmaDec.h:48 */
typedef UInt16 CLzmaProb;            /* LzmaDec.h:17-21 (_LZMA_PROB32 not defined) */

typedef struct _CLzmaProps           /* LzmaDec.h:28-32 */
{
  unsigned lc, lp, pb;
  UInt32 dicSize;
} CLzmaProps;

/* LzmaDec.h:50-69.  sizeof == 136 on LP64; callers construct it with LzmaDec_Construct and poke
 * dic / dicBufSize / dicPos directly (7zDec.c:135-136,150,185-186), so the layout is ABI.
 * In this library `probs` points at a host control block (not at probabilities: those live in
 * device memory); every other field keeps the reference's meaning and is kept up to date. */
typedef struct
{
  CLzmaProps prop;
  CLzmaProb *probs;
  Byte *dic;
  const Byte *buf;
  UInt32 range, code;
  SizeT dicPos;
  SizeT dicBufSize;
  UInt32 processedPos;
  UInt32 checkDicSize;
  unsigned state;
  UInt32 reps[4];
  unsigned remainLen;
  int needFlush;
  int needInitState;
  UInt32 numProbs;
  unsigned tempBufSize;
  Byte tempBuf[LZMA_REQUIRED_INPUT_MAX];
} CLzmaDec;

#define LzmaDec_Construct(p) { (p)->dic = 0; (p)->probs = 0; }   /* LzmaDec.h:71 */

typedef enum { LZMA_FINISH_ANY, LZMA_FINISH_END } ELzmaFinishMode;   /* LzmaDec.h:79-83 */
typedef enum                                                         /* LzmaDec.h:100-107 */
{
  LZMA_STATUS_NOT_SPECIFIED,
  LZMA_STATUS_FINISHED_WITH_MARK,
  LZMA_STATUS_NOT_FINISHED,
  LZMA_STATUS_NEEDS_MORE_INPUT,
  LZMA_STATUS_MAYBE_FINISHED_WITHOUT_MARK
} ELzmaStatus;

SRes LzmaProps_Decode(CLzmaProps *p, const Byte *data, unsigned size);                          /* LzmaDec.h:40  */
void LzmaDec_Init(CLzmaDec *p);                                                                 /* LzmaDec.h:73  */
SRes LzmaDec_AllocateProbs(CLzmaDec *p, const Byte *props, unsigned propsSize, ISzAlloc *alloc); /* LzmaDec.h:134 */
void LzmaDec_FreeProbs(CLzmaDec *p, ISzAlloc *alloc);                                           /* LzmaDec.h:135 */
SRes LzmaDec_Allocate(CLzmaDec *state, const Byte *prop, unsigned propsSize, ISzAlloc *alloc);  /* LzmaDec.h:137 */
void LzmaDec_Free(CLzmaDec *state, ISzAlloc *alloc);                                            /* LzmaDec.h:138 */
void LzmaDec_InitDicAndState(CLzmaDec *p, Bool initDic, Bool initState);                        /* LzmaDec.c:685 */
SRes LzmaDec_DecodeToDic(CLzmaDec *p, SizeT dicLimit, const Byte *src, SizeT *srcLen,
                         ELzmaFinishMode finishMode, ELzmaStatus *status);                      /* LzmaDec.h:181 */
SRes LzmaDec_DecodeToBuf(CLzmaDec *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                         ELzmaFinishMode finishMode, ELzmaStatus *status);                      /* LzmaDec.h:198 */
SRes LzmaDecode(Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, const Byte *propData,
                unsigned propSize, ELzmaFinishMode finishMode, ELzmaStatus *status,
                ISzAlloc *alloc);                                                               /* LzmaDec.h:221 */

/* ---- one-call library API: reference LzmaLib.h:128 ---- */
int LzmaUncompress(unsigned char *dest, size_t *destLen, const unsigned char *src, SizeT *srcLen,
                   const unsigned char *props, size_t propsSize);

/* ---- LZMA2: reference Lzma2Dec.h ---- */
typedef struct                        /* Lzma2Dec.h:15-25, sizeof == 168 */
{
  CLzmaDec decoder;
  UInt32 packSize;
  UInt32 unpackSize;
  int state;
  Byte control;
  Bool needInitDic;
  Bool needInitState;
  Bool needInitProp;
} CLzma2Dec;

#define Lzma2Dec_Construct(p) LzmaDec_Construct(&(p)->decoder)            /* Lzma2Dec.h:27 */
#define Lzma2Dec_FreeProbs(p, alloc) LzmaDec_FreeProbs(&(p)->decoder, alloc);
#define Lzma2Dec_Free(p, alloc) LzmaDec_Free(&(p)->decoder, alloc);

SRes Lzma2Dec_AllocateProbs(CLzma2Dec *p, Byte prop, ISzAlloc *alloc);     /* Lzma2Dec.h:31 */
SRes Lzma2Dec_Allocate(CLzma2Dec *p, Byte prop, ISzAlloc *alloc);          /* Lzma2Dec.h:32 */
void Lzma2Dec_Init(CLzma2Dec *p);                                          /* Lzma2Dec.h:33 */
SRes Lzma2Dec_DecodeToDic(CLzma2Dec *p, SizeT dicLimit, const Byte *src, SizeT *srcLen,
                          ELzmaFinishMode finishMode, ELzmaStatus *status); /* Lzma2Dec.h:51 */
SRes Lzma2Dec_DecodeToBuf(CLzma2Dec *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                          ELzmaFinishMode finishMode, ELzmaStatus *status); /* Lzma2Dec.h:54 */
/* Lzma2Dec.h:75.  The reference's body forgets Lzma2Dec_Init (Lzma2Dec.c:330-356) and returns garbage;
 * this one performs the documented contract (the 7zDec.c:175-222 sequence). */
SRes Lzma2Decode(Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, Byte prop,
                 ELzmaFinishMode finishMode, ELzmaStatus *status, ISzAlloc *alloc);

/* ---- checks: reference 7zCrc.h:11-21, XzCrc64.h:13-22 ---- */
extern UInt32 g_CrcTable[];
extern UInt64 g_Crc64Table[];
void CrcGenerateTable(void);
UInt32 CrcUpdate(UInt32 crc, const void *data, size_t size);
UInt32 CrcCalc(const void *data, size_t size);
void Crc64GenerateTable(void);
UInt64 Crc64Update(UInt64 crc, const void *data, size_t size);
UInt64 Crc64Calc(const void *data, size_t size);

/* ---- post-filters: reference Bra.h:56-58, Delta.h:13-17 ---- */
#define x86_Convert_Init(state) { state = 0; }
SizeT x86_Convert(Byte *data, SizeT size, UInt32 ip, UInt32 *state, int encoding);
#define DELTA_STATE_SIZE 256
void Delta_Init(Byte *state);
void Delta_Decode(Byte *state, unsigned delta, Byte *data, SizeT size);

/* ---- xz: reference Xz.h:21,54,85-96,136-148,223-248 ---- */
typedef enum { CODER_STATUS_NOT_SPECIFIED, CODER_STATUS_FINISHED_WITH_MARK, CODER_STATUS_NOT_FINISHED,
               CODER_STATUS_NEEDS_MORE_INPUT } ECoderStatus;
typedef enum { CODER_FINISH_ANY, CODER_FINISH_END } ECoderFinishMode;
typedef UInt16 CXzStreamFlags;
unsigned Xz_ReadVarInt(const Byte *p, size_t maxSize, UInt64 *value);
unsigned XzFlags_GetCheckSize(CXzStreamFlags f);
SRes Xz_ParseHeader(CXzStreamFlags *p, const Byte *buf);

/* Xz.h:194-221: same size as the reference object (1792 bytes) so callers' storage fits; the content
 * is this library's own (host-side container state; block payloads are decoded on the GPU). */
typedef struct { UInt64 opaque[224]; } CXzUnpacker;
SRes XzUnpacker_Create(CXzUnpacker *p, ISzAlloc *alloc);
void XzUnpacker_Free(CXzUnpacker *p);
/* Contract of Xz.h:223-246.  Round-1 limit: each call must present whole blocks (any number of them,
 * with the headers/index/footer around them) and room for their output; a call that would have to
 * suspend inside a block returns SZ_ERROR_UNSUPPORTED instead of falling back to a CPU decoder. */
SRes XzUnpacker_Code(CXzUnpacker *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                     int finishMode, ECoderStatus *status);
Bool XzUnpacker_IsStreamWasFinished(CXzUnpacker *p);
UInt64 XzUnpacker_GetNumBlocks(const CXzUnpacker *p);   /* reads what callers take from p->numBlocks */

/* ====================================================================================== */
/* Part 2 -- batched extension (no reference analogue)                                     */
/* ====================================================================================== */

#include "../lzma-sdk-zliblike_b200/csrc/lzb_types.h"

typedef struct LzbContext LzbContext;

/* One context = one GPU + its streams, device arenas and pinned staging.  NULL on failure. */
LzbContext *Lzb_Create(int device);
void Lzb_Destroy(LzbContext *ctx);
const char *Lzb_LastError(void);
int Lzb_DeviceSMs(LzbContext *ctx);

/* Pinned host memory for the arenas handed to Lzb_DecodeUnitsHost (pageable memory also works, slower). */
void *Lzb_HostAlloc(size_t bytes);
void Lzb_HostFree(void *p);

/* Decode n units whose packed bytes and outputs live in two HOST arenas (unit offsets are relative to
 * them).  Copies are pipelined against the kernels on several streams.  results[i] mirrors what the
 * reference returns for unit i: (SRes, status, srcLen, destLen) plus the computed digest. */
SRes Lzb_DecodeUnitsHost(LzbContext *ctx, const Byte *packed, size_t packedBytes, Byte *out, size_t outBytes,
                         const LzbUnit *units, LzbResult *results, size_t n);

/* Same with both arenas already resident in DEVICE memory; asynchronous on `stream` (a cudaStream_t;
 * NULL = the CUDA default stream).  d_units / d_results are device arrays.  The packed arena must be
 * readable up to the next 4-byte boundary past its last unit. */
SRes Lzb_DecodeUnitsDevice(LzbContext *ctx, const Byte *d_packed, Byte *d_out, const LzbUnit *d_units,
                           LzbResult *d_results, size_t n, unsigned maxLcLp, unsigned maxPb, void *stream);

/* n independent LzmaUncompress calls (LzmaLib.c:41-46) in one GPU pass.  props: 5 bytes shared by all
 * streams (propsStride == 0) or one 5-byte record per stream.  destLen/srcLen are in/out exactly as in
 * LzmaUncompress; res[i] is the SRes of stream i.  Returns SZ_OK when the batch ran. */
SRes LzmaUncompressBatch(LzbContext *ctx, size_t n, unsigned char *const *dest, size_t *destLen,
                         const unsigned char *const *src, size_t *srcLen, const unsigned char *props,
                         size_t propsStride, SRes *res);

/* Host front end for LZMA2: walk the chunk headers (Lzma2Dec.c:14-26,98-157) and cut the stream at
 * dictionary resets into independently decodable units.  Fills units[0..min(*n,maxUnits)) with
 * src_off/src_len/dst_off/dst_cap relative to `src` / a contiguous output, codec = LZMA2.
 * *endOffset = offset just past the 0x00 end marker when *hasEos, else where the walk stopped. */
SRes Lzma2_IndexUnits(const Byte *src, size_t srcLen, Byte prop, LzbUnit *units, size_t maxUnits, size_t *n,
                      size_t *endOffset, int *hasEos);

/* Whole-buffer LZMA2 decode with the Lzma2Dec_DecodeToDic contract, units decoded in parallel. */
SRes Lzma2DecodeParallel(LzbContext *ctx, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, Byte prop,
                         ELzmaFinishMode finishMode, ELzmaStatus *status);

/* xz block splitter (host): forward walk of stream header / block headers / LZMA2 chunk headers. */
typedef struct LzbXzBlock
{
  uint64_t hdr_off;  uint32_t hdr_size;
  uint64_t data_off; uint64_t pack_size; uint64_t unpack_size;
  uint64_t check_off; uint64_t expect;
  int32_t filter;    uint32_t filter_param;
  uint8_t lzma2_prop; uint8_t check_type; uint8_t complete; uint8_t pad_;
  uint32_t first_unit, n_units;
} LzbXzBlock;

/* Whole-buffer xz decode with the XzUnpacker_Code contract (complete input, large enough output);
 * blocks (and dictionary-reset groups inside them) are decoded in parallel, BCJ/Delta and CRC on the GPU. */
SRes XzDecodeParallel(LzbContext *ctx, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                      ECoderStatus *status, UInt64 *numBlocks, int *streamFinished);

/* Backward index reader (reference XzIn.c:71-114,141-224): block table from the stream index. */
SRes Xz_ReadIndexBackward(const Byte *src, size_t srcLen, LzbXzBlock *blocks, size_t maxBlocks, size_t *n,
                          int *checkType);

#ifdef __cplusplus
}
#endif
#endif
