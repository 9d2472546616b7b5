/* lzb_core.cuh -- the decode path as device code (sm_100a), one UNIT per warp.
 *
 * Execution model
 *   The range decoder is one serial dependency chain per stream, so a unit is owned
 *   by one warp.  All 32 lanes run the bit loop REDUNDANTLY on identical registers (shared-memory
 *   loads broadcast, stores coalesce to one), which keeps control flow warp-uniform: no divergence,
 *   no shuffles to publish len/distance, and the lanes are already converged when the warp-wide
 *   parts start -- model init, match copy, stored-chunk copy, post-filters and CRC, where each lane
 *   takes its own bytes.
 *
 * Memory
 *   probabilities : uint16 in shared memory, COMPACT layout (pos-state strides follow pb, not the
 *                   reference's fixed 16):  950 + (56 << pb) + (768 << (lc+lp)) entries
 *                   = 7,318 (14,636 B) at lc3 lp0 pb2  ->  15 units resident per SM.
 *   packed input  : consumed through a 4-byte register window refilled by aligned 32-bit loads,
 *                   the next word prefetched one refill ahead (the bit loop never waits on HBM).
 *   window/output : the unit's slice of the output arena in HBM *is* the dictionary (dic = dest,
 *                   LzmaDec.c:988-989); literals are stored by lane 0, matches copied by 32 lanes.
 *                   The two bytes the next literal can depend on (previous byte, match byte) are
 *                   fetched when the match is issued, so no literal waits on a window load.
 *
 * Semantics restated from the reference (cited per function): LzmaDec.c / Lzma2Dec.c / Bra86.c /
 * Delta.c / 7zCrc.c / XzCrc64.c.  The same source compiles for the host with -DLZB_EMU (tests/emu):
 * a TEST-ONLY single-lane emulation used to exercise this logic on machines without a GPU.  The
 * product library never contains that build.
 */
#pragma once

#include <stdint.h>
#include <stddef.h>

#include "lzb_types.h"

#ifdef LZB_EMU
#include <string.h>
#define LZB_FN inline
#define LZB_NOINLINE
#define LZB_LANES 1u
static inline uint32_t lzb_lane() { return 0; }
static inline void lzb_syncwarp() {}
static inline uint32_t lzb_bswap(uint32_t w) { return __builtin_bswap32(w); }
static inline uint32_t lzb_funnel_l8(uint32_t lo, uint32_t hi) { return (hi << 8) | (lo >> 24); }
template <class T> static inline T lzb_shfl(T v, int) { return v; }
template <class T> static inline T lzb_shfl_up(T v, unsigned) { return v; }
#else
#define LZB_FN __device__ __forceinline__
#define LZB_NOINLINE __device__ __noinline__
#define LZB_LANES 32u
__device__ __forceinline__ uint32_t lzb_lane() { return threadIdx.x & 31u; }
__device__ __forceinline__ void lzb_syncwarp() { __syncwarp(); }
__device__ __forceinline__ uint32_t lzb_bswap(uint32_t w) { return __byte_perm(w, 0, 0x0123); }
__device__ __forceinline__ uint32_t lzb_funnel_l8(uint32_t lo, uint32_t hi) { return __funnelshift_l(lo, hi, 8); }
template <class T> __device__ __forceinline__ T lzb_shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
#endif

namespace lzb {

constexpr uint32_t kTop = 1u << 24;          /* kTopValue, LzmaDec.c:8-9 */
constexpr uint32_t kProbInit = 1024;         /* kBitModelTotal >> 1 */
constexpr uint32_t kLookahead = 20;          /* LZMA_REQUIRED_INPUT_MAX, LzmaDec.h:48 */
constexpr uint32_t kEndMarkLen = 274;        /* kMatchSpecLenStart, LzmaDec.c:92 */

/* Compact probability map.  Fixed-size tables first, pos-state tables after (stride 1 << pb). */
constexpr uint32_t oIsRep = 0, oRepG0 = 12, oRepG1 = 24, oRepG2 = 36;
constexpr uint32_t oDistSlot = 48;           /* [4][64] */
constexpr uint32_t oDistSpec = 304;          /* 114 */
constexpr uint32_t oDistAlign = 418;         /* 16 */
constexpr uint32_t oLenHigh = 434;           /* [2][256]  (0: match, 1: rep) */
constexpr uint32_t oLenChoice = 946;         /* [2][2] */
constexpr uint32_t oVar = 950;               /* pb-dependent part starts here: */
/*   is_match  = oVar + ( 0 << pb)  [12 << pb]
 *   rep0_long = oVar + (12 << pb)  [12 << pb]
 *   len_low   = oVar + (24 << pb)  [2][8 << pb]
 *   len_mid   = oVar + (40 << pb)  [2][8 << pb]
 *   literal   = oVar + (56 << pb)  [768 << (lc+lp)]                                   */

enum { kSymLit = 0, kSymMatch = 1, kSymRep = 2, kSymStarved = 3 };

/* ------------------------------------------------------------------------------------------ */
/* warp-wide byte copy inside the window; reproduces the reference's strictly forward,        */
/* byte-by-byte semantics (LzmaDec.c:388-407): a source that overlaps the destination repeats. */
LZB_FN void warp_copy_fwd(uint8_t *dic, uint32_t dst, uint32_t src, uint32_t n)
{
  const uint32_t lane = lzb_lane();
  lzb_syncwarp();                                   /* earlier stores of other lanes are visible */
  if (src < dst && dst - src < n)
  {
    /* overlapping: output is the period-(dst-src) repetition of bytes already in the window */
    const uint32_t period = dst - src;
    if (period == 1)
    {
      const uint8_t b = dic[src];
      for (uint32_t i = lane; i < n; i += LZB_LANES) dic[dst + i] = b;
    }
    else
    {
      for (uint32_t i = lane; i < n; i += LZB_LANES) dic[dst + i] = dic[src + i % period];
    }
  }
  else
  {
#ifdef LZB_EMU
    /* emulate a 32-lane batch: all loads of a batch happen before its stores */
    for (uint32_t base = 0; base < n; base += 32)
    {
      uint8_t tmp[32];
      uint32_t m = n - base < 32 ? n - base : 32;
      for (uint32_t l = 0; l < m; l++) tmp[l] = dic[src + base + l];
      for (uint32_t l = 0; l < m; l++) dic[dst + base + l] = tmp[l];
    }
#else
    for (uint32_t i = lane; i < n; i += LZB_LANES) dic[dst + i] = dic[src + i];
#endif
  }
}

/* plain memcpy global -> window (LZMA2 stored chunks, Lzma2Dec.c:159-162) */
LZB_FN void warp_copy_in(uint8_t *dst, const uint8_t *src, uint32_t n)
{
  const uint32_t lane = lzb_lane();
  uint32_t i = 0;
#ifndef LZB_EMU
  if ((((uintptr_t)dst | (uintptr_t)src) & 15u) == 0)
  {
    const uint32_t n16 = n >> 4;
    for (uint32_t k = lane; k < n16; k += LZB_LANES)
      reinterpret_cast<uint4 *>(dst)[k] = reinterpret_cast<const uint4 *>(src)[k];
    i = n16 << 4;
  }
#endif
  for (uint32_t k = i + lane; k < n; k += LZB_LANES) dst[k] = src[k];
  lzb_syncwarp();
}

/* ------------------------------------------------------------------------------------------ */
struct UnitDecoder
{
  /* range coder + input window */
  uint32_t range, code;
  uint32_t win;            /* upcoming input bytes, next byte in bits 31..24 */
  uint32_t win_left;       /* bytes still in `win` (1..4) */
  uint32_t nextw;          /* raw little-endian word that follows `win` (prefetched) */
  const uint32_t *wbase;   /* `in` rounded down to a 4-byte boundary (pointer arithmetic keeps it a global pointer) */
  uint32_t in_skew;        /* in - wbase (0..3) */
  uint32_t widx;           /* word index `nextw` was loaded from */
  uint32_t wcount;         /* first word index we must not load */
  const uint8_t *in;       /* unit's packed bytes */
  uint32_t in_len;

  /* model */
  uint16_t *probs;         /* shared memory: fixed tables, pos-state tables, plain-literal trees (256/context) */
  uint16_t *mprobs;        /* HBM (L2-resident slab): matched-literal trees (512/context), 16-byte aligned    */
  uint32_t lc, lp, lp_mask, pb, pb_mask;

  /* LZ state (CLzmaDec, LzmaDec.h:50-69) */
  uint8_t *dic;
  uint32_t dic_pos, dic_size, dic_commit;
  uint32_t total_pos;      /* processedPos */
  uint32_t dict_full;      /* checkDicSize */
  uint32_t dict_size;
  uint32_t state, rep0, rep1, rep2, rep3;
  uint32_t pending;        /* remainLen */
  uint32_t prev_byte;      /* dic[dicPos-1] (0 right after a dictionary reset)          */
  uint32_t peek_byte;      /* dic[dicPos-rep0], valid while state >= 7 (matched literal) */
  bool want_rc_init, want_state_init;
  uint32_t held_out;       /* input bytes parked for the next call (CLzmaDec.tempBufSize) when starved */
  bool probs_in_smem;      /* `probs` is a shared-memory address (the usual case): enables the ld/st.shared forms */

  /* ---------------- input window ---------------- */
  /* `win` holds the next 1..4 packed bytes (next byte on top); `nextw` is the word after it, loaded one
   * refill ahead; `widx` is the word index nextw came from.  Consumed position = 4*widx - skew - win_left. */
  LZB_FN uint32_t in_pos() const { return (widx << 2) - in_skew - win_left; }

  LZB_FN uint32_t load_word(uint32_t i) const { return i < wcount ? wbase[i] : 0u; }

  LZB_FN void rc_seek(uint32_t pos)
  {
    const uint32_t a = in_skew + pos;
    const uint32_t sh = a & 3;
    widx = a >> 2;
    win = lzb_bswap(load_word(widx)) << (8 * sh);
    win_left = 4 - sh;
    widx += 1;
    nextw = load_word(widx);
  }

  /* LzmaDec.c:17.  The bit loop is issue/latency bound (profiles/), so the hot primitives are written as
   * PTX with UNIFORM branches: all 32 lanes hold identical values, `bra.uni` tells ptxas so and spares
   * the BSSY/BSYNC reconvergence pair a possibly-divergent `if` costs on every bit. */
  LZB_FN void rc_norm()
  {
#ifdef LZB_EMU
    if (range < kTop)
    {
      range <<= 8;
      code = lzb_funnel_l8(win, code);
      win <<= 8;
      if (--win_left == 0)
      {
        win = lzb_bswap(nextw);
        win_left = 4;
        widx += 1;
        nextw = load_word(widx);
      }
    }
#else
    asm volatile(
        "{\n\t"
        ".reg .pred pn, pr, pe;\n\t"
        ".reg .u64 ad;\n\t"
        "setp.ge.u32 pn, %0, 0x1000000;\n\t"
        "@pn bra.uni LZB_NORM_DONE_%=;\n\t"
        "shl.b32 %0, %0, 8;\n\t"
        "shf.l.wrap.b32 %1, %2, %1, 8;\n\t"
        "shl.b32 %2, %2, 8;\n\t"
        "add.u32 %3, %3, -1;\n\t"
        "setp.ne.u32 pr, %3, 0;\n\t"
        "@pr bra.uni LZB_NORM_DONE_%=;\n\t"
        "prmt.b32 %2, %5, 0, 0x0123;\n\t"
        "mov.u32 %3, 4;\n\t"
        "add.u32 %4, %4, 1;\n\t"
        "mov.u32 %5, 0;\n\t"
        "setp.ge.u32 pe, %4, %6;\n\t"
        "@pe bra.uni LZB_NORM_DONE_%=;\n\t"
        "mad.wide.u32 ad, %4, 4, %7;\n\t"
        "ld.global.u32 %5, [ad];\n\t"
        "LZB_NORM_DONE_%=:\n\t"
        "}"
        : "+r"(range), "+r"(code), "+r"(win), "+r"(win_left), "+r"(widx), "+r"(nextw)
        : "r"(wcount), "l"(wbase)
        : "memory");
#endif
  }

  /* LzmaDec.c:19-21.  Select/predicate form, no branch on the decoded bit.  Probability update with
   * one select:  bit 0: p + ((2048 - p) >> 5) == p - ((p - 2017) >>arith 5);  bit 1: p - ((p - 0) >>arith 5). */
  LZB_FN uint32_t rc_bit(uint16_t *p)
  {
    const uint32_t pr = *p;
    rc_norm();
#ifdef LZB_EMU
    const uint32_t bound = (range >> 11) * pr;
    const bool one = code >= bound;
    const uint32_t rest = range - bound;
    range = one ? rest : bound;
    code = one ? code - bound : code;
    const int32_t k = one ? 0 : 2017;
    *p = (uint16_t)(pr - (uint32_t)(((int32_t)pr - k) >> 5));
    return one ? 1u : 0u;
#else
    uint32_t bit, np;
    asm("{\n\t"
        ".reg .pred q;\n\t"
        ".reg .u32 b, t, k;\n\t"
        "shr.u32 b, %0, 11;\n\t"
        "mul.lo.u32 b, b, %4;\n\t"
        "setp.ge.u32 q, %1, b;\n\t"
        "sub.u32 t, %0, b;\n\t"
        "selp.u32 %0, t, b, q;\n\t"
        "@q sub.u32 %1, %1, b;\n\t"
        "selp.u32 %2, 1, 0, q;\n\t"
        "selp.u32 k, 0, 2017, q;\n\t"
        "sub.u32 k, %4, k;\n\t"
        "shr.s32 k, k, 5;\n\t"
        "sub.u32 %3, %4, k;\n\t"
        "}"
        : "+r"(range), "+r"(code), "=r"(bit), "=r"(np)
        : "r"(pr));
    *p = (uint16_t)np;
    return bit;
#endif
  }

  /* Plain literal, LzmaDec.c:174: eight steps down the context's 256-entry tree.  On the GPU the walk is
   * kept in terms of the node's 32-bit shared-memory BYTE address A (node m of a tree at L: A = L + 2m):
   * the child address is 2A + (bit ? 2 - L : -L), one select inside the bit step + one multiply-add, where
   * index arithmetic on m needs four instructions.  80 % of all decoded bits are of this kind (profiles/). */
  LZB_FN uint32_t literal_plain(uint16_t *lit)
  {
#ifdef LZB_EMU
    uint32_t sym = 1;
    for (int i = 0; i < 8; i++) sym = (sym << 1) | rc_bit(lit + sym);
    return sym;
#else
    if (!probs_in_smem)
    {
      uint32_t sym = 1;
#pragma unroll 1
      for (int i = 0; i < 8; i++) sym = (sym << 1) | rc_bit(lit + sym);
      return sym;
    }
    const uint32_t L = (uint32_t)__cvta_generic_to_shared(lit);
    const uint32_t k0 = 0u - L, k1 = 2u - L;
    uint32_t A = L + 2;
#pragma unroll
    for (int i = 0; i < 8; i++)
    {
      uint32_t pr, np, c;
      asm volatile("ld.shared.u16 %0, [%1];" : "=r"(pr) : "r"(A) : "memory");
      rc_norm();
      asm("{\n\t"
          ".reg .pred q;\n\t"
          ".reg .u32 b, t, k;\n\t"
          "shr.u32 b, %0, 11;\n\t"
          "mul.lo.u32 b, b, %4;\n\t"
          "setp.ge.u32 q, %1, b;\n\t"
          "sub.u32 t, %0, b;\n\t"
          "selp.u32 %0, t, b, q;\n\t"
          "@q sub.u32 %1, %1, b;\n\t"
          "selp.u32 %2, %6, %5, q;\n\t"
          "selp.u32 k, 0, 2017, q;\n\t"
          "sub.u32 k, %4, k;\n\t"
          "shr.s32 k, k, 5;\n\t"
          "sub.u32 %3, %4, k;\n\t"
          "}"
          : "+r"(range), "+r"(code), "=r"(c), "=r"(np)
          : "r"(pr), "r"(k0), "r"(k1));
      asm volatile("st.shared.u16 [%0], %1;" : : "r"(A), "r"(np) : "memory");
      A = A * 2 + c;
    }
    return (A - L) >> 1;
#endif
  }

  /* LzmaDec.c:325-334, exact wrap-around arithmetic */
  LZB_FN uint32_t rc_direct()
  {
    rc_norm();
    range >>= 1;
    code -= range;
    const uint32_t neg = 0u - (code >> 31);
    code += range & neg;
    return neg + 1u;
  }

  template <int BITS> LZB_FN uint32_t tree_msb(uint16_t *t)
  {
    uint32_t m = 1;
#pragma unroll
    for (int i = 0; i < BITS; i++) m = (m << 1) | rc_bit(t + m);
    return m - (1u << BITS);
  }

  LZB_FN uint32_t tree_lsb(uint16_t *t, uint32_t bits)
  {
    uint32_t m = 1, v = 0;
    for (uint32_t i = 0; i < bits; i++)
    {
      const uint32_t b = rc_bit(t + m);
      m = (m << 1) | b;
      v |= b << i;
    }
    return v;
  }

  /* ---------------- window ---------------- */
  LZB_FN uint32_t back_index(uint32_t back) const
  {
    return dic_pos - back + (dic_pos < back ? dic_size : 0u);    /* LzmaDec.c:178,220,383 */
  }

  /* bytes the NEXT literal may depend on, fetched now so that it will not wait (design note above) */
  LZB_FN void refresh_context()
  {
    prev_byte = (total_pos != 0 || dict_full != 0) ? dic[(dic_pos == 0 ? dic_size : dic_pos) - 1] : 0u;
    peek_byte = (state >= 7) ? dic[back_index(rep0)] : 0u;
  }

  /* LzmaDec.c:707-717 */
  LZB_FN void init_models()
  {
    const uint32_t ns = lzb_smem_probs(lc, lp, pb), ng = lzb_gmem_probs(lc, lp);
    const uint32_t lane = lzb_lane();
#ifdef LZB_EMU
    for (uint32_t i = 0; i < ns; i++) probs[i] = (uint16_t)kProbInit;
    for (uint32_t i = 0; i < ng; i++) mprobs[i] = (uint16_t)kProbInit;
#else
    /* both counts are even; write two models per 32-bit store (the HBM part: 16 bytes per lane) */
    uint32_t *p32 = reinterpret_cast<uint32_t *>(probs);
    for (uint32_t i = lane; i < (ns >> 1); i += LZB_LANES) p32[i] = kProbInit | (kProbInit << 16);
    uint4 *g128 = reinterpret_cast<uint4 *>(mprobs);
    const uint32_t v = kProbInit | (kProbInit << 16);
    for (uint32_t i = lane; i < (ng >> 3); i += LZB_LANES) g128[i] = make_uint4(v, v, v, v);
#endif
    (void)lane;
    lzb_syncwarp();
    rep0 = rep1 = rep2 = rep3 = 1;
    state = 0;
    want_state_init = false;
  }

  /* LzmaDec.c:685-699 */
  LZB_FN void reset_dic_state(bool reset_dic, bool reset_state)
  {
    want_rc_init = true;
    pending = 0;
    if (reset_dic) { total_pos = 0; dict_full = 0; want_state_init = true; }
    if (reset_state) want_state_init = true;
  }

  /* match / rep copy, LzmaDec.c:376-408.  Caller has checked limit != dic_pos. */
  LZB_FN void copy_from_window(uint32_t cur)
  {
    const uint32_t src = back_index(rep0);
    lzb_syncwarp();                                   /* literal stores of lane 0 are visible to all */
    /* context for the following literal comes from bytes that are already in the window */
    if (src + cur <= dic_size)
    {
      if (src < dic_pos && rep0 <= cur)
      {
        prev_byte = dic[src + (cur - 1) % rep0];
        peek_byte = dic[src + cur % rep0];
      }
      else
      {
        prev_byte = dic[src + cur - 1];
        const uint32_t nx = src + cur;                       /* may step onto the wrap point */
        peek_byte = (nx == dic_size) ? dic[0] : ((nx == dic_pos) ? dic[src] : dic[nx]);
      }
      warp_copy_fwd(dic, dic_pos, src, cur);
    }
    else
    {
      /* source crosses the end of a ring dictionary (buffer interface only) */
      const uint32_t first = dic_size - src;
      warp_copy_fwd(dic, dic_pos, src, first);
      warp_copy_fwd(dic, dic_pos + first, 0, cur - first);
      lzb_syncwarp();
      prev_byte = dic[dic_pos + cur - 1];
      peek_byte = dic[cur - first];
    }
    dic_pos += cur;
    total_pos += cur;
  }

  /* LzmaDec.c:131-426.  Returns 0, or 1 for SZ_ERROR_DATA (nothing is committed then). */
  LZB_FN int run(uint32_t limit, uint32_t soft_end)
  {
    uint32_t left_over = 0;
    uint16_t *const P = probs;
    const uint32_t lane = lzb_lane();
    do
    {
      const uint32_t ps = total_pos & pb_mask;
      if (rc_bit(P + oVar + (state << pb) + ps) == 0)
      {
        /* ---- literal, :161-196 ---- */
        const uint32_t ctx = ((total_pos & lp_mask) << lc) + (prev_byte >> (8 - lc));
        uint16_t *lit = P + oVar + (56u << pb) + (ctx << 8);       /* plain tree of this context (shared) */
        uint32_t sym = 1;
        if (state < 7)
        {
          state -= (state < 4) ? state : 3;
          sym = literal_plain(lit);
        }
        else
        {
          /* matched literal (:176-192): while the decoded bits agree with the match byte the node comes
           * from the context's matched trees (reference index 0x100 + mbit + sym -> HBM slab index
           * mbit + sym); after the first disagreement the rest of the byte uses the plain tree. */
          uint16_t *mlit = mprobs + (ctx << 9);
          uint32_t mb = peek_byte, offs = 0x100;
          state -= (state < 10) ? 3 : 6;
#pragma unroll 1
          for (int i = 0; i < 8; i++)
          {
            mb <<= 1;
            const uint32_t mbit = mb & offs;
            const uint32_t b = rc_bit(offs ? mlit + mbit + sym : lit + sym);
            sym = (sym << 1) | b;
            offs &= b ? mbit : ~mbit;
          }
        }
        prev_byte = sym & 0xFF;
        if (lane == 0) dic[dic_pos] = (uint8_t)sym;
        dic_pos++;
        total_pos++;
        continue;
      }

      uint32_t is_rep = rc_bit(P + oIsRep + state);
      if (is_rep == 0)
      {
        state += 12;                                   /* marker: "new distance follows", :204 */
      }
      else
      {
        if (dict_full == 0 && total_pos == 0) return 1;                         /* :210-211 */
        if (rc_bit(P + oRepG0 + state) == 0)
        {
          if (rc_bit(P + oVar + (12u << pb) + (state << pb) + ps) == 0)
          {
            /* ---- short rep, :219-224 ---- */
            lzb_syncwarp();
            const uint32_t s0 = back_index(rep0);
            const uint8_t b = dic[s0];
            const uint32_t s1 = (s0 + 1 == dic_size) ? 0u : s0 + 1;
            peek_byte = (s1 == dic_pos) ? b : dic[s1];
            prev_byte = b;
            if (lane == 0) dic[dic_pos] = b;
            dic_pos++;
            total_pos++;
            state = state < 7 ? 9 : 11;
            continue;
          }
        }
        else
        {
          uint32_t d;
          if (rc_bit(P + oRepG1 + state) == 0) d = rep1;
          else
          {
            if (rc_bit(P + oRepG2 + state) == 0) d = rep2;
            else { d = rep3; rep3 = rep2; }
            rep2 = rep1;
          }
          rep1 = rep0;
          rep0 = d;
        }
        state = state < 7 ? 8 : 11;
      }

      /* ---- length, :261-292 (match and rep tables side by side) ---- */
      uint32_t len;
      {
        uint16_t *ch = P + oLenChoice + 2 * is_rep;
        if (rc_bit(ch) == 0)
          len = tree_msb<3>(P + oVar + (24u << pb) + ((is_rep << 3) << pb) + (ps << 3));
        else if (rc_bit(ch + 1) == 0)
          len = 8 + tree_msb<3>(P + oVar + (40u << pb) + ((is_rep << 3) << pb) + (ps << 3));
        else
          len = 16 + tree_msb<8>(P + oLenHigh + (is_rep << 8));
      }

      if (state >= 12)
      {
        /* ---- distance, :294-374 ---- */
        uint32_t dist = tree_msb<6>(P + oDistSlot + ((len < 4 ? len : 3u) << 6));
        if (dist >= 4)
        {
          const uint32_t slot = dist;
          uint32_t nbits = (slot >> 1) - 1;
          dist = 2 | (slot & 1);
          if (slot < 14)
          {
            dist <<= nbits;
            dist |= tree_lsb(P + oDistSpec + dist - slot - 1, nbits);
          }
          else
          {
            nbits -= 4;
            do { dist = (dist << 1) + rc_direct(); } while (--nbits != 0);
            dist <<= 4;
            {
              uint16_t *al = P + oDistAlign;
              uint32_t m = 1, b;
              b = rc_bit(al + m); m = (m << 1) | b; dist |= b;
              b = rc_bit(al + m); m = (m << 1) | b; dist |= b << 1;
              b = rc_bit(al + m); m = (m << 1) | b; dist |= b << 2;
              b = rc_bit(al + m); dist |= b << 3;
            }
            if (dist == 0xFFFFFFFFu)
            {
              left_over = len + kEndMarkLen;           /* end marker, :354-359 */
              state -= 12;
              break;
            }
          }
        }
        rep3 = rep2; rep2 = rep1; rep1 = rep0;
        rep0 = dist + 1;
        if (dict_full == 0 ? dist >= total_pos : dist >= dict_full) return 1;    /* :366-372 */
        state = (state < 12 + 7) ? 7 : 10;
      }

      len += 2;
      if (limit == dic_pos) return 1;                                           /* :378-379 */
      {
        const uint32_t room = limit - dic_pos;
        const uint32_t cur = room < len ? room : len;
        left_over = len - cur;
        copy_from_window(cur);
      }
    }
    while (dic_pos < limit && in_pos() < soft_end);
    rc_norm();                                                                   /* :412 */
    pending = left_over;
    dic_commit = dic_pos;
    return 0;
  }

  /* LzmaDec.c:428-452 */
  LZB_FN void flush_pending(uint32_t limit)
  {
    if (pending != 0 && pending < kEndMarkLen)
    {
      uint32_t len = pending;
      if (limit - dic_pos < len) len = limit - dic_pos;
      if (dict_full == 0 && dict_size - total_pos <= len) dict_full = dict_size;
      pending -= len;
      if (len != 0) copy_from_window(len);
      dic_commit = dic_pos;
    }
  }

  /* LzmaDec.c:454-477 */
  LZB_FN int span(uint32_t limit, uint32_t soft_end)
  {
    do
    {
      uint32_t lim = limit;
      if (dict_full == 0)
      {
        const uint32_t room = dict_size - total_pos;
        if (limit - dic_pos > room) lim = dic_pos + room;
      }
      if (run(lim, soft_end)) return 1;
      if (total_pos >= dict_size) dict_full = dict_size;
      flush_pending(limit);
    }
    while (dic_pos < limit && in_pos() < soft_end && pending < kEndMarkLen);
    if (pending > kEndMarkLen) pending = kEndMarkLen;
    return 0;
  }

  /* ---- look-ahead: would one more whole symbol fit?  LzmaDec.c:487-675.  No side effects. ---- */
  struct Peek
  {
    uint32_t range, code, at, end;
    const uint8_t *in;
    bool starved;
    LZB_FN void norm()
    {
      if (range < kTop)
      {
        if (at >= end) { starved = true; return; }
        range <<= 8;
        code = (code << 8) | in[at++];
      }
    }
    LZB_FN uint32_t bit(const uint16_t *p)
    {
      const uint32_t pr = *p;
      norm();
      const uint32_t bound = (range >> 11) * pr;
      if (code < bound) { range = bound; return 0; }
      range -= bound; code -= bound;
      return 1;
    }
    LZB_FN uint32_t tree(const uint16_t *t, uint32_t bits)
    {
      uint32_t m = 1;
      for (uint32_t i = 0; i < bits; i++) m = (m << 1) | bit(t + m);
      return m - (1u << bits);
    }
  };

  /* a free-standing (non-member, by-value) routine so that the decoder object stays in registers */
  static LZB_NOINLINE int peek_symbol_impl(const uint16_t *P, const uint16_t *M, const uint8_t *in, uint32_t range, uint32_t code,
                                           uint32_t at, uint32_t avail, uint32_t state, uint32_t total_pos,
                                           uint32_t lc, uint32_t lp_mask, uint32_t pb, uint32_t prev_byte,
                                           uint32_t peek_byte)
  {
    Peek k;
    k.range = range; k.code = code; k.at = at; k.end = at + avail; k.in = in; k.starved = false;
    const uint32_t ps = total_pos & ((1u << pb) - 1);
    int kind;
    if (k.bit(P + oVar + (state << pb) + ps) == 0)
    {
      const uint32_t ctx = ((total_pos & lp_mask) << lc) + (prev_byte >> (8 - lc));
      const uint16_t *lit = P + oVar + (56u << pb) + (ctx << 8);
      const uint16_t *mlit = M + (ctx << 9);
      uint32_t sym = 1;
      if (state < 7)
      {
        while (sym < 0x100) sym = (sym << 1) | k.bit(lit + sym);
      }
      else
      {
        uint32_t mb = peek_byte, offs = 0x100;
        while (sym < 0x100)
        {
          mb <<= 1;
          const uint32_t mbit = mb & offs;
          const uint32_t b = k.bit(offs ? mlit + mbit + sym : lit + sym);
          sym = (sym << 1) | b;
          offs &= b ? mbit : ~mbit;
        }
      }
      kind = kSymLit;
    }
    else
    {
      uint32_t is_rep = k.bit(P + oIsRep + state);
      kind = is_rep ? kSymRep : kSymMatch;
      bool has_len = true;
      if (is_rep)
      {
        if (k.bit(P + oRepG0 + state) == 0)
        {
          if (k.bit(P + oVar + (12u << pb) + (state << pb) + ps) == 0) has_len = false;   /* short rep */
        }
        else if (k.bit(P + oRepG1 + state) != 0) (void)k.bit(P + oRepG2 + state);
      }
      if (has_len)
      {
        const uint16_t *ch = P + oLenChoice + 2 * is_rep;
        uint32_t len;
        if (k.bit(ch) == 0) len = k.tree(P + oVar + (24u << pb) + ((is_rep << 3) << pb) + (ps << 3), 3);
        else if (k.bit(ch + 1) == 0) len = 8 + k.tree(P + oVar + (40u << pb) + ((is_rep << 3) << pb) + (ps << 3), 3);
        else len = 16 + k.tree(P + oLenHigh + (is_rep << 8), 8);
        if (!is_rep)
        {
          const uint32_t slot = k.tree(P + oDistSlot + ((len < 4 ? len : 3u) << 6), 6);
          if (slot >= 4)
          {
            uint32_t nbits = (slot >> 1) - 1;
            const uint16_t *t;
            if (slot < 14) t = P + oDistSpec + ((2 | (slot & 1)) << nbits) - slot - 1;
            else
            {
              nbits -= 4;
              do
              {
                k.norm();
                k.range >>= 1;
                k.code -= k.range & (((k.code - k.range) >> 31) - 1);
              }
              while (--nbits != 0);
              t = P + oDistAlign;
              nbits = 4;
            }
            uint32_t m = 1;
            do { m = (m << 1) | k.bit(t + m); } while (--nbits != 0);
          }
        }
      }
    }
    k.norm();                                          /* :673 */
    return k.starved ? kSymStarved : kind;
  }

  LZB_FN int peek_symbol(uint32_t at, uint32_t avail) const
  {
    return peek_symbol_impl(probs, mprobs, in, range, code, at, avail, state, total_pos, lc, lp_mask, pb, prev_byte, peek_byte);
  }

  /* LzmaDec.c:719-838 for an input that is completely present (the `tempBuf` continuation branch,
   * :804-833, is only reachable across calls; the host re-feeds held bytes instead -- DESIGN.md). */
  LZB_FN int to_dic(uint32_t dic_limit, uint32_t src, uint32_t avail, int finish, int &status, uint32_t &used)
  {
    bool ctx_stale = false;
    used = 0;
    held_out = 0;
    if (!want_rc_init) { rc_seek(src); ctx_stale = true; }     /* continuing a chunk / a suspended stream */
    flush_pending(dic_limit);
    status = LZB_ST_NOT_SPECIFIED;
    while (pending != kEndMarkLen)
    {
      bool must_end = false;
      if (want_rc_init)                                                          /* :732-746 */
      {
        if (avail < 5) { used += avail; held_out = avail; status = LZB_ST_NEEDS_MORE_INPUT; return LZB_SZ_OK; }
        used += 5;
        if (in[src] != 0) return LZB_SZ_ERROR_DATA;
        code = ((uint32_t)in[src + 1] << 24) | ((uint32_t)in[src + 2] << 16) | ((uint32_t)in[src + 3] << 8) | in[src + 4];
        range = 0xFFFFFFFFu;
        want_rc_init = false;
        src += 5; avail -= 5;
        rc_seek(src);
        ctx_stale = true;
      }
      if (dic_pos >= dic_limit)                                                  /* :749-767 */
      {
        if (pending == 0 && code == 0) { status = LZB_ST_MAYBE_FINISHED_WITHOUT_MARK; return LZB_SZ_OK; }
        if (finish == 0) { status = LZB_ST_NOT_FINISHED; return LZB_SZ_OK; }
        if (pending != 0) { status = LZB_ST_NOT_FINISHED; return LZB_SZ_ERROR_DATA; }
        must_end = true;
      }
      if (want_state_init) init_models();
      if (ctx_stale) { lzb_syncwarp(); refresh_context(); ctx_stale = false; }

      uint32_t soft_end;
      if (avail < kLookahead || must_end)                                        /* :776-793 */
      {
        const int k = peek_symbol(src, avail);
        if (k == kSymStarved) { used += avail; held_out = avail; status = LZB_ST_NEEDS_MORE_INPUT; return LZB_SZ_OK; }
        if (must_end && k != kSymMatch) { status = LZB_ST_NOT_FINISHED; return LZB_SZ_ERROR_DATA; }
        soft_end = src;
      }
      else
        soft_end = src + avail - kLookahead;
      if (span(dic_limit, soft_end)) return LZB_SZ_ERROR_DATA;
      const uint32_t n = in_pos() - src;
      used += n; src += n; avail -= n;
    }
    if (code == 0) status = LZB_ST_FINISHED_WITH_MARK;                           /* :835-837 */
    return code == 0 ? LZB_SZ_OK : LZB_SZ_ERROR_DATA;
  }

  LZB_FN void set_props(uint32_t lc_, uint32_t lp_, uint32_t pb_)
  {
    lc = lc_; lp = lp_; lp_mask = (1u << lp_) - 1; pb = pb_; pb_mask = (1u << pb_) - 1;
  }

  /* resumable decoders: scalars <-> LzbStreamState, model <-> the words behind its header */
  LZB_FN void load_state(const LzbStreamState *s)
  {
    range = s->range; code = s->code; state = s->state;
    rep0 = s->rep[0]; rep1 = s->rep[1]; rep2 = s->rep[2]; rep3 = s->rep[3];
    total_pos = s->total_pos; dict_full = s->dict_full; pending = s->pending;
    want_rc_init = (s->flags & 1) != 0; want_state_init = (s->flags & 2) != 0;
    dic_pos = dic_commit = s->dic_pos;
    set_props(s->lc, s->lp, s->pb);
    if (!want_state_init)
    {
      /* blob behind the header: shared-memory part, then the matched-literal part */
      const uint32_t ns32 = lzb_smem_probs(lc, lp, pb) >> 1, ng32 = lzb_gmem_probs(lc, lp) >> 1;
      const uint32_t *g = reinterpret_cast<const uint32_t *>(s + 1);
      uint32_t *p32 = reinterpret_cast<uint32_t *>(probs);
      uint32_t *m32 = reinterpret_cast<uint32_t *>(mprobs);
      for (uint32_t i = lzb_lane(); i < ns32; i += LZB_LANES) p32[i] = g[i];
      for (uint32_t i = lzb_lane(); i < ng32; i += LZB_LANES) m32[i] = g[ns32 + i];
      lzb_syncwarp();
    }
  }

  LZB_FN void save_state(LzbStreamState *s) const
  {
    lzb_syncwarp();
    if (!want_state_init)
    {
      const uint32_t ns32 = lzb_smem_probs(lc, lp, pb) >> 1, ng32 = lzb_gmem_probs(lc, lp) >> 1;
      uint32_t *g = reinterpret_cast<uint32_t *>(s + 1);
      const uint32_t *p32 = reinterpret_cast<const uint32_t *>(probs);
      const uint32_t *m32 = reinterpret_cast<const uint32_t *>(mprobs);
      for (uint32_t i = lzb_lane(); i < ns32; i += LZB_LANES) g[i] = p32[i];
      for (uint32_t i = lzb_lane(); i < ng32; i += LZB_LANES) g[ns32 + i] = m32[i];
    }
    if (lzb_lane() == 0)
    {
      s->range = range; s->code = code; s->state = state;
      s->rep[0] = rep0; s->rep[1] = rep1; s->rep[2] = rep2; s->rep[3] = rep3;
      s->total_pos = total_pos; s->dict_full = dict_full; s->pending = pending;
      s->flags = (want_rc_init ? 1u : 0u) | (want_state_init ? 2u : 0u);
      s->dic_pos = dic_commit;
      s->held = held_out;
      s->lc = lc; s->lp = lp; s->pb = pb;
    }
  }

  LZB_FN void begin(const uint8_t *packed, uint32_t packed_len, uint8_t *out, uint32_t out_cap,
                    uint32_t dict, uint16_t *model, uint16_t *matched_model, bool model_in_smem = false)
  {
    held_out = 0;
    mprobs = matched_model;
    probs_in_smem = model_in_smem;
    in = packed; in_len = packed_len;
    in_skew = (uint32_t)((uintptr_t)packed & 3);
    wbase = reinterpret_cast<const uint32_t *>(packed - in_skew);
    wcount = (in_skew + packed_len + 3) >> 2;
    probs = model;
    dic = out; dic_size = out_cap; dic_pos = 0; dic_commit = 0;
    dict_size = dict;
    range = code = 0; win = 0; win_left = 4; nextw = 0; widx = wcount;
    state = 0; rep0 = rep1 = rep2 = rep3 = 1;
    prev_byte = peek_byte = 0;
    reset_dic_state(true, true);                                                 /* LzmaDec_Init, :701-705 */
  }
};

/* ------------------------------------------------------------------------------------------ */
/* One-call raw LZMA: LzmaDecode (LzmaDec.c:972-1002), which LzmaUncompress wraps with           */
/* LZMA_FINISH_ANY (LzmaLib.c:41-46).                                                           */
/* (the raw stream shares the loop below so that the hot symbol loop is instantiated exactly once)
 *
 * LZMA2 chunk layer over a whole unit: Lzma2Dec_DecodeToDic (Lzma2Dec.c:170-289) with the chunk
 * header machine (:98-157) folded into the loop.  The unit's chunk headers are re-read here by
 * the decoding warp (<= 6 bytes per <= 2 MiB chunk); the HOST walks the same headers once to cut
 * the stream into units (dictionary-reset groups) -- see lzb_host.cu. */
LZB_FN void decode_unit(UnitDecoder &d, const LzbUnit &u, LzbResult &r, LzbStreamState *st = nullptr)
{
  /* phase numbering == ELzma2State (Lzma2Dec.c:47-59): it is what CLzma2Dec.state holds between calls */
  enum { PH_CONTROL, PH_U0, PH_U1, PH_P0, PH_P1, PH_PROPS, PH_DATA, PH_DATA_MORE, PH_FINISHED, PH_ERROR };
  const uint8_t *in = d.in;
  const uint32_t in_size = u.src_len, dic_limit = u.dst_cap;
  const int finish = u.finish;
  const bool raw = u.codec == LZB_CODEC_LZMA;
  uint32_t at_in = 0, pack_left = 0, unpack_left = 0, control = 0;
  int phase = PH_CONTROL;
  bool want_dic = true, want_state = true, want_props = true;
  int status = LZB_ST_INHERIT, res = LZB_SZ_OK;
  r.check = 0;
  if (raw)
  {
    if (!st && in_size < 5) { r.res = LZB_SZ_ERROR_INPUT_EOF; r.status = LZB_ST_NOT_SPECIFIED; r.src_used = 0; r.dst_done = 0; return; }
    d.set_props(u.lc, u.lp, u.pb);
    control = 0x80; phase = PH_DATA_MORE; pack_left = in_size; unpack_left = 0xFFFFFFFFu;
  }
  else
    d.set_props(4, 0, 0);                              /* Lzma2Dec_GetOldProps, Lzma2Dec.c:61-73 */
  if (st)
  {
    /* resume: CLzmaDec / CLzma2Dec as the previous call left them (host keeps the scalars authoritative) */
    d.load_state(st);
    if (!raw)
    {
      phase = (int)st->l2_phase; control = st->l2_control;
      pack_left = st->l2_pack_left + st->held;         /* bytes parked in tempBuf were already charged */
      unpack_left = st->l2_unpack_left;
      want_dic = (st->l2_flags & 1) != 0; want_state = (st->l2_flags & 2) != 0; want_props = (st->l2_flags & 4) != 0;
    }
  }

  for (;;)
  {
    if (phase == PH_FINISHED) { status = LZB_ST_FINISHED_WITH_MARK; break; }
    const uint32_t at = d.dic_commit;
    if (phase == PH_ERROR) { res = LZB_SZ_ERROR_DATA; break; }
    /* finish == 2 (host-internal "unit" mode): like FINISH_END inside the unit's last chunk, so that its
     * proper termination is still verified, but stop at the unit's end instead of reading the next header */
    if (!raw && at == dic_limit && (finish == 0 || (finish == 2 && phase == PH_CONTROL)))
    { if (finish == 0) status = LZB_ST_NOT_FINISHED; break; }   /* unit mode: a boundary is not a stop of the stream */

    if (phase != PH_DATA && phase != PH_DATA_MORE)
    {
      if (at_in == in_size) { status = LZB_ST_NEEDS_MORE_INPUT; break; }
      uint32_t b = in[at_in++];
      switch (phase)                                   /* Lzma2Dec.c:98-157 */
      {
        case PH_CONTROL:
          control = b;
          if (b == 0) phase = PH_FINISHED;
          else if ((b & 0x80) == 0) { if ((b & 0x7F) > 2) phase = PH_ERROR; else { unpack_left = 0; phase = PH_U0; } }
          else { unpack_left = (b & 0x1F) << 16; phase = PH_U0; }
          break;
        case PH_U0: unpack_left |= b << 8; phase = PH_U1; break;
        case PH_U1: unpack_left |= b; unpack_left++; phase = (control & 0x80) ? PH_P0 : PH_DATA; break;
        case PH_P0: pack_left = b << 8; phase = PH_P1; break;
        case PH_P1:
          pack_left |= b; pack_left++;
          phase = (((control >> 5) & 3) >= 2) ? PH_PROPS : (want_props ? PH_ERROR : PH_DATA);
          break;
        default: /* PH_PROPS */
        {
          if (b >= 225) { phase = PH_ERROR; break; }
          const uint32_t lc = b % 9; b /= 9;
          const uint32_t lp = b % 5;
          if (lc + lp > 4) { phase = PH_ERROR; break; }
          d.set_props(lc, lp, b / 5);
          want_props = false;
          phase = PH_DATA;
        }
      }
      continue;
    }

    uint32_t out_now = dic_limit - at;
    uint32_t in_now = in_size - at_in;
    int mode = raw ? finish : 0;
    if (!raw && unpack_left <= out_now) { out_now = unpack_left; mode = 1; }

    if ((control & 0x80) == 0)                         /* stored chunk, :209-240 */
    {
      if (at_in == in_size) { status = LZB_ST_NEEDS_MORE_INPUT; break; }
      if (phase == PH_DATA)
      {
        const bool reset_dic = control == 1;
        if (reset_dic) want_props = want_state = true;
        else if (want_dic) { res = LZB_SZ_ERROR_DATA; break; }
        want_dic = false;
        d.reset_dic_state(reset_dic, false);
      }
      if (in_now > out_now) in_now = out_now;
      if (in_now == 0) { res = LZB_SZ_ERROR_DATA; break; }
      warp_copy_in(d.dic + d.dic_pos, in + at_in, in_now);     /* :159-166 */
      d.dic_pos += in_now;
      d.dic_commit = d.dic_pos;
      if (d.dict_full == 0 && d.dict_size - d.total_pos <= in_now) d.dict_full = d.dict_size;
      d.total_pos += in_now;
      at_in += in_now;
      unpack_left -= in_now;
      phase = unpack_left == 0 ? PH_CONTROL : PH_DATA_MORE;
    }
    else                                               /* LZMA chunk, :241-284 */
    {
      if (phase == PH_DATA)
      {
        const uint32_t m = (control >> 5) & 3;
        const bool reset_dic = m == 3, reset_state = m > 0;
        if ((!reset_dic && want_dic) || (!reset_state && want_state)) { res = LZB_SZ_ERROR_DATA; break; }
        d.reset_dic_state(reset_dic, reset_state);
        want_dic = want_state = false;
        phase = PH_DATA_MORE;
      }
      if (in_now > pack_left) in_now = pack_left;
      uint32_t used;
      res = d.to_dic(at + out_now, at_in, in_now, mode, status, used);
      at_in += used;
      if (raw)
      {
        /* one-call LzmaDecode (LzmaDec.c:972-1002), which LzmaUncompress wraps with FINISH_ANY */
        if (!st && res == LZB_SZ_OK && status == LZB_ST_NEEDS_MORE_INPUT) res = LZB_SZ_ERROR_INPUT_EOF;
        break;
      }
      pack_left -= used;
      const uint32_t made = d.dic_commit - at;
      unpack_left -= made;
      if (res != LZB_SZ_OK) break;
      if (status == LZB_ST_NEEDS_MORE_INPUT) break;
      if (used == 0 && made == 0)
      {
        if (status != LZB_ST_MAYBE_FINISHED_WITHOUT_MARK || unpack_left != 0 || pack_left != 0)
        { res = LZB_SZ_ERROR_DATA; break; }
        phase = PH_CONTROL;
      }
      if (status == LZB_ST_MAYBE_FINISHED_WITHOUT_MARK) status = LZB_ST_NOT_FINISHED;
    }
  }
  r.res = res; r.status = status; r.src_used = at_in; r.dst_done = d.dic_commit; r.check = 0;
  if (st)
  {
    d.save_state(st);
    if (lzb_lane() == 0)
    {
      st->l2_phase = (uint32_t)phase; st->l2_control = control;
      st->l2_pack_left = pack_left; st->l2_unpack_left = unpack_left;
      st->l2_flags = (want_dic ? 1u : 0u) | (want_state ? 2u : 0u) | (want_props ? 4u : 0u);
    }
  }
}

}  // namespace lzb
