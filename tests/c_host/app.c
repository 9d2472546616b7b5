/* app.c -- a C host written against the LZMA SDK 9.20 API names, compiled against include/lzb200.h and linked with
 * -llzb200 instead of the SDK's LzmaDec.c / Lzma2Dec.c / LzmaLib.c / XzDec.c / 7zCrc.c (the
 * link-line swap).  tests/test_c_host.py builds it on the CPU (the link must resolve every symbol) and runs it on a GPU.
 *
 *   app <packed.lzma-raw> <props-hex> <unpacked-size>      decodes the stream three ways and prints crc32 of each:
 *        LzmaUncompress (LzmaLib.h:128), the LzmaDec_Allocate + LzmaDec_DecodeToBuf loop of LzmaUtil.c:55-106 with
 *        64 KiB buffers, and LzmaDecode (LzmaDec.h:221); then CrcCalc of the output (7zCrc.h:20).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lzb200.h"

static void *SzAlloc(void *p, size_t size) { (void)p; return size ? malloc(size) : 0; }
static void SzFree(void *p, void *address) { (void)p; free(address); }
static ISzAlloc g_Alloc = { SzAlloc, SzFree };

#define IN_BUF_SIZE (1 << 16)
#define OUT_BUF_SIZE (1 << 16)

int main(int argc, char **argv)
{
  if (argc != 4) { fprintf(stderr, "usage: app packed props-hex size\n"); return 2; }
  FILE *f = fopen(argv[1], "rb");
  if (!f) return 2;
  fseek(f, 0, SEEK_END);
  size_t packedSize = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  Byte *packed = (Byte *)malloc(packedSize + 1);
  if (fread(packed, 1, packedSize, f) != packedSize) return 2;
  fclose(f);
  Byte props[LZMA_PROPS_SIZE];
  for (int i = 0; i < LZMA_PROPS_SIZE; i++) { unsigned v; sscanf(argv[2] + 2 * i, "%2x", &v); props[i] = (Byte)v; }
  const size_t unpackSize = (size_t)strtoull(argv[3], 0, 10);
  Byte *out = (Byte *)malloc(unpackSize + 1);
  CrcGenerateTable();

  /* 1. LzmaUncompress */
  size_t destLen = unpackSize, srcLen = packedSize;
  int res = LzmaUncompress(out, &destLen, packed, &srcLen, props, LZMA_PROPS_SIZE);
  printf("LzmaUncompress res=%d destLen=%zu srcLen=%zu crc=%08x\n", res, destLen, srcLen, (unsigned)CrcCalc(out, destLen));

  /* 2. the LzmaUtil loop over a resumable decoder object */
  CLzmaDec state;
  LzmaDec_Construct(&state);
  res = LzmaDec_Allocate(&state, props, LZMA_PROPS_SIZE, &g_Alloc);
  if (res != SZ_OK) { printf("LzmaDec_Allocate res=%d\n", res); return 1; }
  LzmaDec_Init(&state);
  memset(out, 0, unpackSize);
  size_t inPos = 0, outPos = 0;
  ELzmaStatus status = LZMA_STATUS_NOT_SPECIFIED;
  for (;;)
  {
    SizeT inProcessed = packedSize - inPos < IN_BUF_SIZE ? packedSize - inPos : IN_BUF_SIZE;
    SizeT outProcessed = unpackSize - outPos < OUT_BUF_SIZE ? unpackSize - outPos : OUT_BUF_SIZE;
    const ELzmaFinishMode mode = outProcessed == unpackSize - outPos ? LZMA_FINISH_END : LZMA_FINISH_ANY;
    res = LzmaDec_DecodeToBuf(&state, out + outPos, &outProcessed, packed + inPos, &inProcessed, mode, &status);
    inPos += inProcessed;
    outPos += outProcessed;
    if (res != SZ_OK || outPos == unpackSize || (inProcessed == 0 && outProcessed == 0)) break;
  }
  LzmaDec_Free(&state, &g_Alloc);
  printf("LzmaDec_DecodeToBuf res=%d status=%d out=%zu in=%zu crc=%08x\n", res, (int)status, outPos, inPos, (unsigned)CrcCalc(out, outPos));

  /* 3. LzmaDecode */
  memset(out, 0, unpackSize);
  destLen = unpackSize; srcLen = packedSize;
  res = LzmaDecode(out, &destLen, packed, &srcLen, props, LZMA_PROPS_SIZE, LZMA_FINISH_END, &status, &g_Alloc);
  printf("LzmaDecode res=%d status=%d destLen=%zu srcLen=%zu crc=%08x\n", res, (int)status, destLen, srcLen, (unsigned)CrcCalc(out, destLen));
  free(out);
  free(packed);
  return 0;
}
