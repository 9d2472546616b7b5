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

/* ---- LZMA + x86 filter, one-call: reference Lzma86.h:11-12,87,107 ---- */
#define LZMA86_SIZE_OFFSET (1 + 5)
#define LZMA86_HEADER_SIZE (LZMA86_SIZE_OFFSET + 8)
SRes Lzma86_GetUnpackSize(const Byte *src, SizeT srcLen, UInt64 *unpackSize);
SRes Lzma86_Decode(Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen);

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
SizeT ARM_Convert(Byte *data, SizeT size, UInt32 ip, int encoding);      /* Bra.h:59-63; stateless, elementwise */
SizeT ARMT_Convert(Byte *data, SizeT size, UInt32 ip, int encoding);
SizeT PPC_Convert(Byte *data, SizeT size, UInt32 ip, int encoding);
SizeT SPARC_Convert(Byte *data, SizeT size, UInt32 ip, int encoding);
SizeT IA64_Convert(Byte *data, SizeT size, UInt32 ip, int encoding);
/* BCJ2 merge (Bcj2.h:24-30): buf0 main, buf1 CALL targets, buf2 JUMP targets, buf3 range-coder stream */
int Bcj2_Decode(const Byte *buf0, SizeT size0, const Byte *buf1, SizeT size1, const Byte *buf2, SizeT size2,
                const Byte *buf3, SizeT size3, Byte *outBuf, SizeT outSize);
#define DELTA_STATE_SIZE 256
void Delta_Init(Byte *state);
void Delta_Encode(Byte *state, unsigned delta, Byte *data, SizeT size);
void Delta_Decode(Byte *state, unsigned delta, Byte *data, SizeT size);

/* ---- xz: reference Xz.h (ids :11-19, block :33-54, check :66-85, coders :136-185, unpacker :187-248),
 *      Sha256.h:13-22.  The structs below are ABI (callers allocate them and read their fields), so they
 *      keep the reference's field order and sizes: CXzBlock 152, CXzCheck 120, IStateCoder 40,
 *      CMixCoder 272, CXzUnpacker 1792 bytes on x86-64 (asserted in tests/test_abi_and_host.py). ---- */
#define XZ_ID_Delta 3
#define XZ_ID_X86 4
#define XZ_ID_PPC 5
#define XZ_ID_IA64 6
#define XZ_ID_ARM 7
#define XZ_ID_ARMT 8
#define XZ_ID_SPARC 9
#define XZ_ID_LZMA2 0x21
#define XZ_CHECK_NO 0
#define XZ_CHECK_CRC32 1
#define XZ_CHECK_CRC64 4
#define XZ_CHECK_SHA256 10
#define XZ_BLOCK_HEADER_SIZE_MAX 1024
#define XZ_NUM_FILTERS_MAX 4
#define XZ_FILTER_PROPS_SIZE_MAX 20
#define SHA256_DIGEST_SIZE 32

typedef struct { UInt32 state[8]; UInt64 count; Byte buffer[64]; } CSha256;
void Sha256_Init(CSha256 *p);
void Sha256_Update(CSha256 *p, const Byte *data, size_t size);
void Sha256_Final(CSha256 *p, Byte *digest);

typedef enum { CODER_STATUS_NOT_SPECIFIED, CODER_STATUS_FINISHED_WITH_MARK, CODER_STATUS_NOT_FINISHED,
               CODER_STATUS_NEEDS_MORE_INPUT } ECoderStatus;
typedef enum { CODER_FINISH_ANY, CODER_FINISH_END } ECoderFinishMode;
typedef UInt16 CXzStreamFlags;
unsigned Xz_ReadVarInt(const Byte *p, size_t maxSize, UInt64 *value);
unsigned Xz_WriteVarInt(Byte *buf, UInt64 v);
unsigned XzFlags_GetCheckSize(CXzStreamFlags f);
SRes Xz_ParseHeader(CXzStreamFlags *p, const Byte *buf);

typedef struct { UInt64 id; UInt32 propsSize; Byte props[XZ_FILTER_PROPS_SIZE_MAX]; } CXzFilter;
typedef struct { UInt64 packSize; UInt64 unpackSize; Byte flags; CXzFilter filters[XZ_NUM_FILTERS_MAX]; } CXzBlock;
#define XzBlock_GetNumFilters(p) (((p)->flags & 3) + 1)
#define XzBlock_HasPackSize(p)   (((p)->flags & (1 << 6)) != 0)
#define XzBlock_HasUnpackSize(p) (((p)->flags & (1 << 7)) != 0)
SRes XzBlock_Parse(CXzBlock *p, const Byte *header);                  /* XzDec.c:505-556 */

typedef struct { int mode; UInt32 crc; UInt64 crc64; CSha256 sha; } CXzCheck;
void XzCheck_Init(CXzCheck *p, int mode);                             /* Xz.c:45-88; digests computed on the GPU */
void XzCheck_Update(CXzCheck *p, const void *data, size_t size);
int XzCheck_Final(CXzCheck *p, Byte *digest);

typedef struct _IStateCoder
{
  void *p;
  void (*Free)(void *p, ISzAlloc *alloc);
  SRes (*SetProps)(void *p, const Byte *props, size_t propSize, ISzAlloc *alloc);
  void (*Init)(void *p);
  SRes (*Code)(void *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
               int srcWasFinished, ECoderFinishMode finishMode, int *wasFinished);
} IStateCoder;

#define MIXCODER_NUM_FILTERS_MAX 4
typedef struct
{
  ISzAlloc *alloc;
  Byte *buf;
  int numCoders;
  int finished[MIXCODER_NUM_FILTERS_MAX - 1];
  size_t pos[MIXCODER_NUM_FILTERS_MAX - 1];
  size_t size[MIXCODER_NUM_FILTERS_MAX - 1];
  UInt64 ids[MIXCODER_NUM_FILTERS_MAX];
  IStateCoder coders[MIXCODER_NUM_FILTERS_MAX];
} CMixCoder;

/* The resumable filter chain of an xz block (XzDec.c:356-480): LZMA2 first (stateful, on the GPU), then a
 * branch converter (x86, PPC, IA64, ARM, ARMT, SPARC) or Delta over a 16 KiB staging buffer exactly as
 * BraState_Code does. */
void MixCoder_Construct(CMixCoder *p, ISzAlloc *alloc);
void MixCoder_Free(CMixCoder *p);
void MixCoder_Init(CMixCoder *p);
SRes MixCoder_SetFromMethod(CMixCoder *p, int coderIndex, UInt64 methodId);
SRes MixCoder_Code(CMixCoder *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen, int srcWasFinished,
                   ECoderFinishMode finishMode, ECoderStatus *status);
SRes XzDec_Init(CMixCoder *p, const CXzBlock *block);                 /* XzDec.c:558-588 */

typedef enum { XZ_STATE_STREAM_HEADER, XZ_STATE_STREAM_INDEX, XZ_STATE_STREAM_INDEX_CRC, XZ_STATE_STREAM_FOOTER,
               XZ_STATE_STREAM_PADDING, XZ_STATE_BLOCK_HEADER, XZ_STATE_BLOCK, XZ_STATE_BLOCK_FOOTER } EXzState;

typedef struct
{
  EXzState state;
  UInt32 pos;
  unsigned alignPos;
  unsigned indexPreSize;
  CXzStreamFlags streamFlags;
  UInt32 blockHeaderSize;
  UInt64 packSize;
  UInt64 unpackSize;
  UInt64 numBlocks;
  UInt64 indexSize;
  UInt64 indexPos;
  UInt64 padSize;
  UInt64 numStreams;
  UInt32 crc;
  CMixCoder decoder;
  CXzBlock block;
  CXzCheck check;
  CSha256 sha;
  Byte shaDigest[SHA256_DIGEST_SIZE];
  Byte buf[XZ_BLOCK_HEADER_SIZE_MAX];
} CXzUnpacker;

SRes XzUnpacker_Create(CXzUnpacker *p, ISzAlloc *alloc);
void XzUnpacker_Free(CXzUnpacker *p);
/* Contract of Xz.h:223-246 / XzDec.c:604-870, resumable at any byte: the container state machine runs on
 * the host; block payloads are decoded on the GPU.  Whole blocks found in `src` with room for their output
 * in `dest` are decoded in parallel in one pass (all of them, and the dictionary-reset groups inside them);
 * a block that is cut by the end of `src` or `dest` goes through the resumable chain (MixCoder_Code). */
SRes XzUnpacker_Code(CXzUnpacker *p, Byte *dest, SizeT *destLen, const Byte *src, SizeT *srcLen,
                     int finishMode, ECoderStatus *status);
Bool XzUnpacker_IsStreamWasFinished(CXzUnpacker *p);
UInt64 XzUnpacker_GetNumBlocks(const CXzUnpacker *p);   /* p->numBlocks of the current stream */

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

/* The same job on SEVERAL GPUs of one box (units are independent, the multi-GPU path is a partition,
 * not a collective).  ctxs[0..nCtx) are contexts of distinct devices.  The unit table is cut into nCtx contiguous
 * ranges balanced by unpacked bytes (Lzb_ShardBounds); one host thread per context uploads only the packed bytes of
 * its range and brings back its own slab of the output.  results[] is filled in unit order; the return value is
 * the worst per-context status. */
SRes Lzb_DecodeUnitsHostMulti(LzbContext *const *ctxs, size_t nCtx, const Byte *packed, size_t packedBytes, Byte *out,
                              size_t outBytes, const LzbUnit *units, LzbResult *results, size_t n);

/* bounds[0..world]: rank r owns units [bounds[r], bounds[r+1]).  Contiguous, balanced by cumulative dst_cap. */
void Lzb_ShardBounds(const LzbUnit *units, size_t n, size_t world, size_t *bounds);

/* Same with both arenas already resident in DEVICE memory; asynchronous on `stream` (a cudaStream_t;
 * NULL = the CUDA default stream).  Launches of one context are ordered one behind the other, also across streams.  d_units / d_results are device arrays.  The packed arena must be
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

/* ---- 7z folder path (reference 7zIn.c SzArEx_Open / SzArEx_Extract, 7zDec.c SzFolder_Decode) ----
 * The archive database is parsed on the host (the LZMA-packed header is decoded on the GPU); all folders of
 * the archive are then decoded in one batched GPU pass, with BCJ x86 / ARM, the folder CRC-32 and every
 * file's CRC-32 computed on the device. */
typedef struct LzbSzFolder
{
  uint64_t pack_off;        /* absolute offset of the folder's packed stream in the archive            */
  uint64_t pack_size;
  uint64_t unpack_size;
  uint64_t out_off;         /* where SzArchive_ExtractAll puts the folder's bytes in the output arena   */
  uint32_t method;          /* 0 Copy, 1 LZMA, 2 LZMA2                                                  */
  uint32_t filter;          /* 0 none, 4 BCJ x86, 7 ARM (xz filter ids)                                 */
  uint32_t crc;
  uint8_t crc_defined, props_size;
  uint8_t supported;        /* SRes of the reference's CheckSupportedFolder (7zDec.c:269-331)           */
  uint8_t pad_;
  uint8_t props[8];
  uint32_t first_file, num_streams;
} LzbSzFolder;

typedef struct LzbSzFile
{
  uint64_t size;
  uint64_t out_off;         /* offset of the file's bytes in the output arena (files with a stream)     */
  uint32_t folder;          /* 0xFFFFFFFF: no folder (empty file / directory)                           */
  uint32_t crc;
  uint8_t crc_defined, is_dir, has_stream, attrib_defined, mtime_defined, pad_[3];
  uint32_t name_off, name_len;   /* UTF-16 code units in SzArchive_Names(), name_len includes the 0     */
  uint32_t attrib, mtime_lo, mtime_hi;
} LzbSzFile;

typedef struct LzbSzArchive LzbSzArchive;
/* SzArEx_Open's result codes (7z.h:197-205) for an archive held in memory. */
SRes SzArchive_Open(LzbContext *ctx, const Byte *archive, size_t size, LzbSzArchive **out);
void SzArchive_Free(LzbSzArchive *a);
size_t SzArchive_NumFolders(const LzbSzArchive *a);
size_t SzArchive_NumFiles(const LzbSzArchive *a);
const LzbSzFolder *SzArchive_Folders(const LzbSzArchive *a);
const LzbSzFile *SzArchive_Files(const LzbSzArchive *a);
const UInt16 *SzArchive_Names(const LzbSzArchive *a, size_t *numUnits);
UInt64 SzArchive_UnpackSize(const LzbSzArchive *a);
/* Decode every folder into out[0 .. SzArchive_UnpackSize).  folderRes[i]: SzFolder_Decode + folder CRC verdict;
 * fileRes[k]: what SzArEx_Extract returns for file k (7zIn.c:1322-1405).  Either may be NULL. */
SRes SzArchive_ExtractAll(LzbContext *ctx, const LzbSzArchive *a, const Byte *archive, size_t size, Byte *out, size_t outCap,
                          SRes *folderRes, SRes *fileRes);

#ifdef __cplusplus
}
#endif
#endif
