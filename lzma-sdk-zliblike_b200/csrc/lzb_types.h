/* lzb_types.h -- records shared by the host C-ABI layer and the sm_100a kernels.
 *
 * A UNIT is the independently decodable piece of work one warp owns:
 *   raw LZMA stream (LzmaUncompress call), LZMA2 dictionary-reset group, or xz block.
 * Vocabulary follows the reference (units/chunks/blocks/dictionary), not ML terms.
 */
#ifndef LZB_TYPES_H
#define LZB_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { LZB_CODEC_LZMA = 0, LZB_CODEC_LZMA2 = 1 };
enum { LZB_FILTER_NONE = 0, LZB_FILTER_DELTA = 3, LZB_FILTER_X86 = 4, LZB_FILTER_PPC = 5, LZB_FILTER_IA64 = 6,
       LZB_FILTER_ARM = 7, LZB_FILTER_ARMT = 8, LZB_FILTER_SPARC = 9 };        /* xz filter ids, Xz.h:11-19 */
enum { LZB_CHECK_NONE = 0, LZB_CHECK_CRC32 = 1, LZB_CHECK_CRC64 = 4, LZB_CHECK_SHA256 = 10 };   /* xz check ids, Xz.h:66-69 */
enum
{
  LZB_UNIT_VERIFY = 1,       /* compare the computed digest with `expect` -> SZ_ERROR_CRC              */
  LZB_UNIT_X86_ENCODE = 2,   /* stand-alone x86_Convert: encoding direction                          */
  LZB_UNIT_POST_REPORT = 4,  /* stand-alone filter call: report processed count / history in result  */
  LZB_UNIT_STATEFUL = 8      /* resumable decoder: `expect` = device address of an LzbStreamState,
                                `filter_param` = dicBufSize (ring size), `dst_cap` = dicLimit         */
};

/* Device-resident state of a resumable decoder (the SDK's CLzmaDec / CLzma2Dec between calls).
 * Scalars mirror the reference fields; the probability model follows the 128-byte header. */
typedef struct LzbStreamState
{
  uint32_t range, code, state, rep[4];            /* CLzmaDec.range/code/state/reps          */
  uint32_t total_pos, dict_full, pending;         /* processedPos, checkDicSize, remainLen   */
  uint32_t flags;                                 /* bit0 needFlush, bit1 needInitState      */
  uint32_t dic_pos;                               /* dicPos in, dicPos out                   */
  uint32_t held;                                  /* in: bytes re-presented from tempBuf; out: bytes to keep */
  uint32_t lc, lp, pb;
  uint32_t l2_phase, l2_control, l2_pack_left, l2_unpack_left;   /* CLzma2Dec.state/control/packSize/unpackSize */
  uint32_t l2_flags;                              /* bit0 needInitDic, bit1 needInitState, bit2 needInitProp */
  uint32_t reserved[11];
} LzbStreamState;                                 /* 128 bytes */

typedef struct LzbUnit
{
  uint64_t src_off;       /* packed bytes start here in the packed arena            */
  uint64_t dst_off;       /* output goes here in the output arena (also the window) */
  uint32_t src_len;       /* packed bytes available                                 */
  uint32_t dst_cap;       /* output capacity == dicLimit == dicBufSize              */
  uint32_t dict_size;     /* LZMA dictionary size (match-distance validity)         */
  uint8_t  codec;         /* LZB_CODEC_*                                            */
  uint8_t  lc, lp, pb;    /* raw LZMA properties (LZMA2 reads them from chunks)     */
  uint8_t  finish;        /* ELzmaFinishMode: 0 ANY, 1 END                          */
  uint8_t  filter;        /* LZB_FILTER_* applied in place after decode             */
  uint8_t  check;         /* LZB_CHECK_*  computed over the final bytes             */
  uint8_t  flags;         /* LZB_UNIT_*                                             */
  uint32_t filter_param;  /* delta distance (1..256) or x86 start offset            */
  uint64_t expect;        /* stored digest, little-endian value                     */
} LzbUnit;                /* 48 bytes */

typedef struct LzbResult
{
  int32_t  res;           /* SRes (Types.h:25-41)                        */
  int32_t  status;        /* ELzmaStatus (LzmaDec.h:100-107)             */
  uint32_t src_used;      /* *srcLen on return                           */
  uint32_t dst_done;      /* *destLen / dicPos on return                 */
  uint64_t check;         /* computed digest (0 when check == NONE)      */
} LzbResult;              /* 24 bytes */

#define LZB_SZ_OK 0
#define LZB_SZ_ERROR_DATA 1
#define LZB_SZ_ERROR_MEM 2
#define LZB_SZ_ERROR_CRC 3
#define LZB_SZ_ERROR_UNSUPPORTED 4
#define LZB_SZ_ERROR_PARAM 5
#define LZB_SZ_ERROR_INPUT_EOF 6
#define LZB_SZ_ERROR_OUTPUT_EOF 7
#define LZB_SZ_ERROR_FAIL 11
#define LZB_SZ_ERROR_ARCHIVE 16
#define LZB_SZ_ERROR_NO_ARCHIVE 17

/* LZMA2 units only: the unit never assigned *status (only stored chunks / headers ran), so the caller's
 * previous value stands -- the reference's *status is sticky across chunks (Lzma2Dec.c:175,282-283) */
#define LZB_ST_INHERIT (-1)
#define LZB_ST_NOT_SPECIFIED 0
#define LZB_ST_FINISHED_WITH_MARK 1
#define LZB_ST_NOT_FINISHED 2
#define LZB_ST_NEEDS_MORE_INPUT 3
#define LZB_ST_MAYBE_FINISHED_WITHOUT_MARK 4

/* Probability-model footprint for given properties (compact layout of lzb_core.cuh). */
#ifdef __CUDACC__
__host__ __device__
#endif
static inline uint32_t lzb_num_probs(unsigned lc, unsigned lp, unsigned pb)
{
  return 950u + (56u << pb) + (768u << (lc + lp));
}
/* The model is kept in two places (lzb_core.cuh): everything a plain symbol touches -- 950 fixed +
 * pos-state tables + the 256-entry plain-literal tree of every context -- in shared memory, and the
 * 512-entry "matched literal" trees (used only for the literal right after a match) in an L2-resident
 * slab in HBM.  lzb_num_probs == lzb_smem_probs + lzb_gmem_probs. */
#ifdef __CUDACC__
__host__ __device__
#endif
static inline uint32_t lzb_smem_probs(unsigned lc, unsigned lp, unsigned pb)
{
  return 950u + (56u << pb) + (256u << (lc + lp));
}
#ifdef __CUDACC__
__host__ __device__
#endif
static inline uint32_t lzb_gmem_probs(unsigned lc, unsigned lp)
{
  return 512u << (lc + lp);
}

/* One SHA-256 job of lzb_sha256_kernel (xz block check 10, Xz.c:45-88): hash `len` bytes at data + off.
 * finalize = 0: len is a multiple of 64, `state` is updated in place (Sha256_Update of whole blocks);
 * finalize = 1: the message ends here (`before` bytes were hashed into `state` earlier), digest is written. */
typedef struct LzbShaJob
{
  uint64_t off, len, before;
  uint32_t state[8];
  uint32_t finalize, pad_;
  uint8_t digest[32];
} LzbShaJob;

/* One BCJ2 merge (reference Bcj2.c:28-127; 7z folders with the four-coder x86 chain, 7zDec.c:308-329,406-430):
 * main / call / jump streams (already decoded, in the output arena) + the raw range-coder stream (in the
 * packed arena) -> `out_size` bytes at out_off.  The main stream sits at the END of the output range and the
 * merge runs in place, front to back, exactly as the reference lays it out. */
typedef struct LzbBcj2Job
{
  uint64_t main_off, call_off, jump_off, out_off;   /* offsets in the output arena */
  uint64_t rc_off;                                  /* offset in the packed arena  */
  uint32_t main_size, call_size, jump_size, rc_size, out_size;
  int32_t res;                                      /* SZ_OK / SZ_ERROR_DATA       */
} LzbBcj2Job;

#ifdef __cplusplus
}
#endif
#endif
