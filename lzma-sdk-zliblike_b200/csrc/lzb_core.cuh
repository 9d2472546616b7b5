/* lzb_core.cuh -- the decode path as device code (sm_100a).
 *
 * Execution model: K = 32 / G UNITS PER WARP, G lanes each, decoded in lockstep
 *   A range-coded stream is one serial dependency chain and the kernel is bound by issue
 *   slots (profiles/): an instruction issued for a warp costs the same whether 1 or 32 lanes do useful
 *   work.  So a warp carries K independent units; every lane of a unit's G-lane GROUP holds that unit's
 *   decoder state redundantly (shared-memory loads broadcast inside the group, control flow is uniform
 *   inside the group), and the warp runs ONE loop in which each group decodes one symbol per trip.
 *   Groups that decode the same kind of symbol execute the same instructions together (a plain literal is
 *   80 % of all decoded bits on the bench corpus); groups on different paths are serialised by the
 *   hardware and reconverge at the end of the trip.  The G lanes of a group split the byte work:
 *   model init, match copy, stored-chunk copy, post-filters, CRC.
 *
 *   To let K units share one loop the SDK's nested drivers (LzmaDec_DecodeToDic -> DecodeReal2 ->
 *   DecodeReal, and Lzma2Dec_DecodeToDic around them) are written as ONE explicit state machine
 *   (`advance`): it runs until the unit needs symbols decoded (state S_RUN: the hot `step`) or is finished.
 *
 * Memory
 *   probabilities : uint16, COMPACT layout (pos-state strides follow pb, not the reference's fixed 16).
 *                   What a plain symbol touches -- 950 fixed + (56 << pb) + the 256-entry plain-literal tree
 *                   of every context -- lives in shared memory (6,448 B at lc3 lp0 pb2); the 512-entry
 *                   matched-literal trees (used only right after a match) live in an L2-resident HBM slab.
 *   packed input  : consumed through a register window refilled by aligned 32-bit loads, the next word prefetched
 *                   one refill ahead (the bit loop never waits on HBM).
 *   window/output : the unit's slice of the output arena in HBM *is* the dictionary (dic = dest,
 *                   LzmaDec.c:988-989); literals are stored by the group's lane 0, matches copied by G lanes.
 *                   The two bytes the next literal can depend on (previous byte, match byte) are fetched
 *                   when the match is issued, so no literal waits on a window load.
 *
 * Two forms of the symbol decoder (UnitDecoderT<G, kSmem, kAhead>; DESIGN.md 4a, profiles/r02_kernel_log.md):
 *   lean       (kAhead = false, lzb_core_lean.inc)  dense launches, two units per warp: fewest issued instructions
 *              (normalisation = a rarely taken branch, 4-byte window)
 *   look-ahead (kAhead = true, this file)           one unit per warp: predicated normalisation over an 8-byte window,
 *              both children of a tree node loaded while the node's bit is decided
 * Everything around the symbol (driver machine, window copies, model init, state save / load) is shared.
 *
 * Semantics restated from the reference (cited per function): LzmaDec.c / Lzma2Dec.c.  The same source
 * compiles for the host with -DLZB_EMU (tests/emu): a TEST-ONLY emulation with G = 1 used to exercise this
 * logic -- both forms -- on machines without a GPU.  The product library never contains that build.
 */
#pragma once

#include <stdint.h>
#include <stddef.h>
#include <string.h>

#include "lzb_types.h"

#ifdef LZB_EMU
#define LZB_FN inline
#define LZB_NOINLINE
static inline uint32_t lzb_bswap(uint32_t w) { return __builtin_bswap32(w); }
static inline uint32_t lzb_funnel_l8(uint32_t lo, uint32_t hi) { return (hi << 8) | (lo >> 24); }
#else
#define LZB_FN __device__ __forceinline__
#define LZB_NOINLINE __device__ __noinline__
__device__ __forceinline__ uint32_t lzb_bswap(uint32_t w) { return __byte_perm(w, 0, 0x0123); }
__device__ __forceinline__ uint32_t lzb_funnel_l8(uint32_t lo, uint32_t hi) { return __funnelshift_l(lo, hi, 8); }
#endif

namespace lzb {

constexpr uint32_t kTop = 1u << 24;          /* kTopValue, LzmaDec.c:8-9 */
constexpr uint32_t kProbInit = 1024;         /* kBitModelTotal >> 1 */
constexpr uint32_t kLookahead = 20;          /* LZMA_REQUIRED_INPUT_MAX, LzmaDec.h:48 */
constexpr uint32_t kEndMarkLen = 274;        /* kMatchSpecLenStart, LzmaDec.c:92 */

/* Compact probability map (shared-memory part).  Fixed-size tables first, pos-state tables after. */
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
 *   literal   = oVar + (56 << pb)  [256 << (lc+lp)]   plain trees; matched trees: HBM slab [512 << (lc+lp)]
 * (The high length trees stay in shared memory: with them in the slab 40 units fit instead of 36, but the
 *  eight dependent L2 round trips per long match cost more than the residency gains: 7.56 vs 7.67 GB/s.) */

enum { kSymLit = 0, kSymMatch = 1, kSymRep = 2, kSymStarved = 3 };

/* states of the per-unit driver machine */
enum { S_FETCH = 0, S_L2_TOP, S_TD_ENTER, S_TD_TOP, S_SPAN_TOP, S_RUN, S_AFTER_RUN, S_TD_RET, S_UNIT_END, S_EXIT };
/* LZMA2 header phases == ELzma2State (Lzma2Dec.c:47-59): it is what CLzma2Dec.state holds between calls */
enum { PH_CONTROL, PH_U0, PH_U1, PH_P0, PH_P1, PH_PROPS, PH_DATA, PH_DATA_MORE, PH_FINISHED, PH_ERROR };

/* ------------------------------------------------------------------------------------------ */
/* the G lanes that share one unit                                                             */
template <uint32_t G> struct Group
{
#ifdef LZB_EMU
  static LZB_FN uint32_t lane() { return 0; }
  static LZB_FN void sync() {}
  template <class T> static LZB_FN T bcast(T v, uint32_t) { return v; }
  template <class T> static LZB_FN T shfl_xor(T v, uint32_t) { return v; }
  template <class T> static LZB_FN T shfl_down(T v, uint32_t) { return v; }
  static LZB_FN uint32_t ballot(bool p) { return p ? 1u : 0u; }
#else
  static LZB_FN uint32_t lane() { return threadIdx.x & (G - 1); }
  static LZB_FN uint32_t base() { return (threadIdx.x & 31u) & ~(G - 1); }
  static LZB_FN uint32_t mask() { return G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << base()); }
  static LZB_FN void sync() { __syncwarp(mask()); }
  template <class T> static LZB_FN T bcast(T v, uint32_t src) { return __shfl_sync(mask(), v, (int)src, (int)G); }
  template <class T> static LZB_FN T shfl_xor(T v, uint32_t m) { return __shfl_xor_sync(mask(), v, (int)m, (int)G); }
  template <class T> static LZB_FN T shfl_down(T v, uint32_t d) { return __shfl_down_sync(mask(), v, d, (int)G); }
  /* bit i set <=> sub-lane i of this group voted true */
  static LZB_FN uint32_t ballot(bool p) { return (__ballot_sync(mask(), p) >> base()) & (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u)); }
#endif
};

/* group-wide byte copy inside the window; reproduces the reference's strictly forward, byte-by-byte
 * semantics (LzmaDec.c:388-407): a source that overlaps the destination repeats. */
template <uint32_t G> LZB_FN void group_copy_fwd(uint8_t *dic, uint32_t dst, uint32_t src, uint32_t n)
{
  const uint32_t lane = Group<G>::lane();
  Group<G>::sync();                                   /* earlier stores of other lanes are visible */
  if (src < dst && dst - src < n)
  {
    /* overlapping: output is the period-(dst-src) repetition of bytes already in the window */
    const uint32_t period = dst - src;
    if (period == 1)
    {
      /* a byte run (rep0 = 1).  One byte per lane and trip is enough: a 16-byte-per-lane fill was measured on data
       * that is nothing but runs and changed nothing (165 vs 169 GB/s out: the run's ~14 modelled bits, not its
       * stores, are what a full-length run costs; profiles/r02_copy_width.md) */
      const uint32_t b = dic[src];
      for (uint32_t i = lane; i < n; i += G) dic[dst + i] = (uint8_t)b;
    }
    else
    {
      for (uint32_t i = lane; i < n; i += G) dic[dst + i] = dic[src + i % period];
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
    for (uint32_t i = lane; i < n; i += G) dic[dst + i] = dic[src + i];
#endif
  }
}

/* plain memcpy global -> window (LZMA2 stored chunks, Lzma2Dec.c:159-162) */
template <uint32_t G> LZB_FN void group_copy_in(uint8_t *dst, const uint8_t *src, uint32_t n)
{
  const uint32_t lane = Group<G>::lane();
  uint32_t i = 0;
#ifndef LZB_EMU
  if ((((uintptr_t)dst | (uintptr_t)src) & 15u) == 0)
  {
    const uint32_t n16 = n >> 4;
    for (uint32_t k = lane; k < n16; k += G)
      reinterpret_cast<uint4 *>(dst)[k] = reinterpret_cast<const uint4 *>(src)[k];
    i = n16 << 4;
  }
#endif
  for (uint32_t k = i + lane; k < n; k += G) dst[k] = src[k];
  Group<G>::sync();
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

/* a free-standing, by-value routine so that the decoder object stays in registers */
static LZB_NOINLINE int peek_symbol_impl(const uint16_t *P, const uint16_t *M, const uint8_t *in, uint32_t range,
                                         uint32_t code, uint32_t at, uint32_t avail, uint32_t state,
                                         uint32_t total_pos, uint32_t lc, uint32_t lp_mask, uint32_t pb,
                                         uint32_t prev_byte, uint32_t peek_byte)
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

/* ------------------------------------------------------------------------------------------ */
template <uint32_t G, bool kSmem = false, bool kAhead = false> struct UnitDecoderT
{
  typedef Group<G> Grp;

  /* range coder + input window */
  uint32_t range, code;
  /*  kAhead = false (many units per SM: issue slots are what is scarce) */
  uint32_t win;            /* upcoming input bytes, next byte in bits 31..24 */
  uint32_t win_left;       /* bytes still in `win` (1..4) */
  /*  kAhead = true (a warp per unit and SMs mostly idle: the unit's own dependency chain is what is scarce) */
  uint32_t wlo, whi;       /* two consecutive raw (little-endian) words of the packed input               */
  uint32_t k;              /* bytes of that pair already consumed (0..8)                                  */
  uint32_t nb;             /* the next input byte (byte k of the pair) in bits 7..0; bits 31..8 undefined */
  uint32_t im_pr;          /* is-match probability of the NEXT symbol, loaded while this one is still being decoded */
  /*  both */
  uint32_t nextw;          /* raw little-endian word that follows the window (prefetched)                 */
  const uint32_t *wbase;   /* `in` rounded down to a 4-byte boundary (pointer arithmetic keeps it a global pointer) */
  uint32_t in_skew;        /* in - wbase (0..3) */
  uint32_t widx;           /* word index `nextw` was loaded from */
  uint32_t wcount;         /* first word index we must not load */
  const uint8_t *in;       /* unit's packed bytes */

  /* model */
  uint16_t *probs;         /* fixed tables, pos-state tables, plain-literal trees (256/context): shared memory when kSmem */
  uint16_t *mprobs;        /* matched-literal trees (512/context), 16-byte aligned: HBM slab (L2-resident) or shared memory */
  uint32_t pbase;          /* model ADDRESS of probs[0]: its 32-bit shared-memory address when kSmem, else 0 (byte offset
                            * from `probs`).  Every model access of the hot path is ld/st.shared on pbase + 2 * index. */
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

  /* driver machine */
  uint32_t st;             /* S_* */
  /*  frame of LzmaDec_DecodeToDic */
  uint32_t td_limit, td_src, td_avail, td_used;
  int td_finish, td_res;
  bool ctx_stale, run_err;
  uint32_t run_limit, run_soft;
  /*  frame of the unit (Lzma2Dec_DecodeToDic, or the one-call raw stream) */
  uint32_t in_size, dic_limit, at, at_in, pack_left, unpack_left, control;
  int phase, finish, status, res;
  bool raw, want_dic, want_state, want_props;
  LzbStreamState *sst;

  /* ---------------- input window ----------------
   * Two forms of the same thing (LzmaDec.c:17: "range < 2^24 => shift one input byte in"), chosen per launch:
   *
   * kAhead = false, launches with many units per SM.  Issued instructions are what is scarce, so the normalisation
   *   is a compare and a branch that is taken once in eight bits (2 instructions when it is not); behind the branch a
   *   4-byte register window (`win`, next byte on top) is shifted and, once in four bytes, refilled from `nextw`,
   *   which is loaded one refill ahead.  Consumed position = 4 * widx - skew - win_left.
   *
   * kAhead = true, launches with a warp per unit and at most a few units per SM (profiles/r02_a_*: the unit's
   *   instructions issue one every 4-5 cycles, each waiting for the one before; the branch around the normalisation
   *   alone cost 30 cycles per bit).  Straight-line predicated code: (wlo, whi) hold eight input bytes in memory
   *   order, `k` of them consumed, `nb` is kept ready as byte k of the pair, so a normalisation is "shift the range,
   *   shift the byte into the code, count, pick the next byte".  top_up() -- a compare and a rarely taken branch,
   *   placed so that at most six modelled bits (<= 5 bytes: a bit shrinks the range by less than 2^6.05) lie between
   *   two of them -- slides the pair one word along when k >= 4.  Consumed position = 4 * (widx - 2) + k - skew. */
  LZB_FN uint32_t in_pos() const
  {
    return kAhead ? ((widx - 2) << 2) + k - in_skew : (widx << 2) - in_skew - win_left;
  }

  LZB_FN uint32_t load_word(uint32_t i) const { return i < wcount ? wbase[i] : 0u; }

  LZB_FN uint32_t byte_at(uint32_t kk) const
  {
#ifdef LZB_EMU
    return (uint32_t)(((((uint64_t)whi) << 32) | wlo) >> (8 * (kk & 7))) & 0xFFu;
#else
    return __byte_perm(wlo, whi, kk);                   /* selector nibble 0 = kk (0..7): byte kk of the pair */
#endif
  }

  LZB_FN void rc_seek(uint32_t pos)
  {
    const uint32_t a = in_skew + pos;
    const uint32_t i = a >> 2;
    if (kAhead)
    {
      wlo = load_word(i);
      whi = load_word(i + 1);
      widx = i + 2;
      nextw = load_word(widx);
      k = a & 3;
      nb = byte_at(k);
    }
    else
    {
      const uint32_t sh = a & 3;
      win = lzb_bswap(load_word(i)) << (8 * sh);
      win_left = 4 - sh;
      widx = i + 1;
      nextw = load_word(widx);
    }
  }

  LZB_FN void slide_word()
  {
    wlo = whi; whi = nextw; k -= 4;
    widx += 1;
    nextw = load_word(widx);
  }

  LZB_FN void top_up()
  {
    if (kAhead && k >= 4)
    {
      slide_word();
      if (k >= 4) slide_word();                          /* only after five bytes in six bits */
      nb = byte_at(k);
    }
  }

  /* LzmaDec.c:17 */
  LZB_FN void rc_norm()
  {
    if (kAhead)
    {
#ifdef LZB_EMU
      if (range < kTop)
      {
        range <<= 8;
        code = (code << 8) | (nb & 0xFFu);
        k += 1;
        nb = byte_at(k);
      }
#else
      asm("{\n\t"
          ".reg .pred n;\n\t"
          "setp.lt.u32 n, %0, 0x1000000;\n\t"
          "@n shl.b32 %0, %0, 8;\n\t"
          "@n prmt.b32 %1, %1, %3, 0x2104;\n\t"          /* code = code << 8 | byte 0 of nb */
          "@n add.u32 %2, %2, 1;\n\t"
          "@n prmt.b32 %3, %4, %5, %2;\n\t"              /* nb = byte k of (wlo, whi) */
          "}"
          : "+r"(range), "+r"(code), "+r"(k), "+r"(nb)
          : "r"(wlo), "r"(whi));
#endif
      return;
    }
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
    /* PTX so that the common case stays at setp + bra; the groups of a warp are not in general uniform here,
     * so these are ordinary (reconverging) branches */
    asm volatile(
        "{\n\t"
        ".reg .pred pn, pr, pe;\n\t"
        ".reg .u64 ad;\n\t"
        "setp.ge.u32 pn, %0, 0x1000000;\n\t"
        "@pn bra LZB_NORM_DONE_%=;\n\t"
        "shl.b32 %0, %0, 8;\n\t"
        "shf.l.wrap.b32 %1, %2, %1, 8;\n\t"
        "shl.b32 %2, %2, 8;\n\t"
        "add.u32 %3, %3, -1;\n\t"
        "setp.ne.u32 pr, %3, 0;\n\t"
        "@pr bra LZB_NORM_DONE_%=;\n\t"
        "prmt.b32 %2, %5, 0, 0x0123;\n\t"
        "mov.u32 %3, 4;\n\t"
        "add.u32 %4, %4, 1;\n\t"
        "mov.u32 %5, 0;\n\t"
        "setp.ge.u32 pe, %4, %6;\n\t"
        "@pe bra LZB_NORM_DONE_%=;\n\t"
        "mad.wide.u32 ad, %4, 4, %7;\n\t"
        "ld.global.u32 %5, [ad];\n\t"
        "LZB_NORM_DONE_%=:\n\t"
        "}"
        : "+r"(range), "+r"(code), "+r"(win), "+r"(win_left), "+r"(widx), "+r"(nextw)
        : "r"(wcount), "l"(wbase)
        : "memory");
#endif
  }

  /* model access by ADDRESS (pbase + 2 * index) */
  LZB_FN uint32_t P(uint32_t index) const { return pbase + (index << 1); }

  LZB_FN uint32_t ldp(uint32_t a) const
  {
#ifndef LZB_EMU
    if (kSmem)
    {
      uint32_t v;
      asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
      return v;
    }
#endif
    return *reinterpret_cast<const uint16_t *>(reinterpret_cast<const uint8_t *>(probs) + a);
  }

  LZB_FN void stp(uint32_t a, uint32_t v) const
  {
#ifndef LZB_EMU
    if (kSmem)
    {
      asm volatile("st.shared.u16 [%0], %1;" : : "r"(a), "r"(v) : "memory");
      return;
    }
#endif
    *reinterpret_cast<uint16_t *>(reinterpret_cast<uint8_t *>(probs) + a) = (uint16_t)v;
  }

  /* LzmaDec.c:19-21 on a loaded probability: range / code step and the updated probability.  Select form, no
   * branch on the decoded bit; returns v1 for a 1 bit, v0 for a 0 bit.  Probability update with one select:
   * bit 0: p + ((2048 - p) >> 5) == p - ((p - 2017) >>arith 5);  bit 1: p - ((p - 0) >>arith 5). */
  LZB_FN uint32_t bit_math(uint32_t pr, uint32_t v0, uint32_t v1, uint32_t &np)
  {
#ifdef LZB_EMU
    const uint32_t bound = (range >> 11) * pr;
    const bool one = code >= bound;
    const uint32_t rest = range - bound;
    range = one ? rest : bound;
    code = one ? code - bound : code;
    const int32_t kk = one ? 0 : 2017;
    np = (uint16_t)(pr - (uint32_t)(((int32_t)pr - kk) >> 5));
    return one ? v1 : v0;
#else
    uint32_t c;
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
        : "r"(pr), "r"(v0), "r"(v1));
    return c;
#endif
  }

  LZB_FN uint32_t ldp32(uint32_t a) const               /* two adjacent probabilities, a 4-byte aligned */
  {
#ifndef LZB_EMU
    if (kSmem)
    {
      uint32_t v;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
      return v;
    }
#endif
    const uint16_t *q = reinterpret_cast<const uint16_t *>(reinterpret_cast<const uint8_t *>(probs) + a);
    return (uint32_t)q[0] | ((uint32_t)q[1] << 16);
  }

  /* the same step with two values selected by the decoded bit */
  LZB_FN void bit_math2(uint32_t pr, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, uint32_t &np, uint32_t &sa, uint32_t &sb)
  {
#ifdef LZB_EMU
    const uint32_t one = bit_math(pr, 0u, 1u, np);
    sa = one ? a1 : a0;
    sb = one ? b1 : b0;
#else
    asm("{\n\t"
        ".reg .pred q;\n\t"
        ".reg .u32 b, t, k;\n\t"
        "shr.u32 b, %0, 11;\n\t"
        "mul.lo.u32 b, b, %5;\n\t"
        "setp.ge.u32 q, %1, b;\n\t"
        "selp.u32 %2, %7, %6, q;\n\t"
        "selp.u32 %3, %9, %8, q;\n\t"
        "sub.u32 t, %0, b;\n\t"
        "selp.u32 %0, t, b, q;\n\t"
        "@q sub.u32 %1, %1, b;\n\t"
        "selp.u32 k, 0, 2017, q;\n\t"
        "sub.u32 k, %5, k;\n\t"
        "shr.s32 k, k, 5;\n\t"
        "sub.u32 %4, %5, k;\n\t"
        "}"
        : "+r"(range), "+r"(code), "=r"(sa), "=r"(sb), "=r"(np)
        : "r"(pr), "r"(a0), "r"(a1), "r"(b0), "r"(b1));
#endif
  }

  /* one modelled bit at model address a */
  LZB_FN uint32_t bit_sel(uint32_t a, uint32_t v0, uint32_t v1)
  {
    const uint32_t pr = ldp(a);
    rc_norm();
    uint32_t np;
    const uint32_t c = bit_math(pr, v0, v1, np);
    stp(a, np);
    return c;
  }
  /* the same through a generic pointer (matched-literal trees: HBM slab or shared memory) */
  LZB_FN uint32_t rc_bit(uint16_t *p)
  {
    const uint32_t pr = *p;
    rc_norm();
    uint32_t np;
    const uint32_t c = bit_math(pr, 0u, 1u, np);
    *p = (uint16_t)np;
    return c;
  }

  LZB_FN uint32_t rc_bit_at(uint32_t a) { return bit_sel(a, 0u, 1u); }

  /* Walks down a binary tree stored as t[1 .. 2^n), kept in terms of the node's model ADDRESS A (node m of the
   * tree whose element 0 is at L: A = L + 2m): the child address is 2A + (bit ? 2 - L : -L) -- the bit step
   * selects the constant, one multiply-add follows.  80 % of all decoded bits are plain-literal walks. */
  template <int BITS> LZB_FN uint32_t tree_walk(uint32_t L)
  {
    if (kAhead)
    {
      /* The two children of node m sit side by side (t[2m], t[2m+1]: one aligned 32-bit word, every tree starts
       * at an even index), so both candidates for the next step are loaded while this step's bit is still being
       * decided: the shared-memory round trip leaves the chain, what stays is select -> multiply -> compare. */
      uint32_t A = L + 2;
      uint32_t pr = ldp(A);
#pragma unroll
      for (int i = 0; i < BITS; i++)
      {
        if (BITS > 6 && i == 4) top_up();
        const uint32_t A0 = A * 2 - L;                   /* child 0; child 1 is the next entry */
        const uint32_t pair = (i + 1 < BITS) ? ldp32(A0) : 0u;
        /* (a version of this step shaped for a shorter chain -- (range >> 11) picked from two shifts that do not wait for
         * the normalisation, both probability updates computed ahead of the compare: +4 instructions -- ran 2-9 % SLOWER
         * at 1, 7, 14 and 15 warps per SM: a warp issues about one instruction per four cycles whatever their
         * dependences, so the count is what matters here too; profiles/r02_kernel_log.md G) */
        rc_norm();
        uint32_t np, nA, npr;
        bit_math2(pr, A0, A0 + 2, pair & 0xFFFFu, pair >> 16, np, nA, npr);
        stp(A, np);
        A = nA;
        pr = npr;
      }
      return (A - L) >> 1;
    }
    const uint32_t k0 = 0u - L, k1 = 2u - L;
    uint32_t A = L + 2;
#pragma unroll
    for (int i = 0; i < BITS; i++) A = A * 2 + bit_sel(A, k0, k1);
    return (A - L) >> 1;                                 /* node index reached, leading 1 bit kept */
  }

  /* n more steps (n <= 6 without a top_up) from node address A; returns the node address reached */
  LZB_FN uint32_t tree_steps(uint32_t L, uint32_t A, uint32_t n)
  {
    const uint32_t k0 = 0u - L, k1 = 2u - L;
#pragma unroll 1
    for (; n != 0; n--) A = A * 2 + bit_sel(A, k0, k1);
    return A;
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

  template <int BITS> LZB_FN uint32_t tree_msb(uint32_t index) { return tree_walk<BITS>(P(index)) - (1u << BITS); }

  /* reverse tree (LzmaDec.c:305-319,345-353): the same walk; the value is the decoded bits in reverse order */
  static LZB_FN uint32_t reversed(uint32_t m, uint32_t bits)        /* m = 1 b0 b1 ... b(bits-1)  ->  sum of b_i << i */
  {
#ifdef LZB_EMU
    uint32_t v = 0;
    for (uint32_t i = 0; i < bits; i++) v |= ((m >> (bits - 1 - i)) & 1u) << i;
    return v;
#else
    return __brev(m) >> (32 - bits);
#endif
  }
  template <int BITS> LZB_FN uint32_t tree_lsb_fixed(uint32_t index) { return reversed(tree_walk<BITS>(P(index)), BITS); }

  LZB_FN uint32_t tree_lsb(uint32_t index, uint32_t bits)
  {
    const uint32_t L = P(index);
    const uint32_t m = (tree_steps(L, L + 2, bits) - L) >> 1;     /* 1 b0 b1 ... b(bits-1) */
    return reversed(m, bits);
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
    const uint32_t lane = Grp::lane();
#ifdef LZB_EMU
    for (uint32_t i = 0; i < ns; i++) probs[i] = (uint16_t)kProbInit;
    for (uint32_t i = 0; i < ng; i++) mprobs[i] = (uint16_t)kProbInit;
#else
    /* both counts are even; write two models per 32-bit store (the HBM part: 16 bytes per lane) */
    uint32_t *p32 = reinterpret_cast<uint32_t *>(probs);
    for (uint32_t i = lane; i < (ns >> 1); i += G) p32[i] = kProbInit | (kProbInit << 16);
    uint4 *g128 = reinterpret_cast<uint4 *>(mprobs);
    const uint32_t v = kProbInit | (kProbInit << 16);
    for (uint32_t i = lane; i < (ng >> 3); i += G) g128[i] = make_uint4(v, v, v, v);
#endif
    (void)lane;
    Grp::sync();
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
    Grp::sync();                                      /* literal stores of lane 0 are visible to the group */
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
        /* dic[dicPos + cur - rep0]: wrap FIRST, then ask whether that is the first byte this copy is about to
         * write (only when the copy starts at the wrap point, dic_pos == 0, with cur == rep0) */
        uint32_t nx = src + cur;
        if (nx == dic_size) nx = 0;
        peek_byte = (nx == dic_pos) ? dic[src] : dic[nx];
      }
      group_copy_fwd<G>(dic, dic_pos, src, cur);
    }
    else
    {
      /* source crosses the end of a ring dictionary (buffer interface only) */
      const uint32_t first = dic_size - src;
      group_copy_fwd<G>(dic, dic_pos, src, first);
      group_copy_fwd<G>(dic, dic_pos + first, 0, cur - first);
      Grp::sync();
      prev_byte = dic[dic_pos + cur - 1];
      peek_byte = dic[cur - first];
    }
    dic_pos += cur;
    total_pos += cur;
  }

#include "lzb_core_lean.inc"

  /* One LZMA symbol: the body of LzmaDec_DecodeReal's loop, LzmaDec.c:153-410.
   * Returns 0, 1 for SZ_ERROR_DATA (nothing is committed then), 2 for the end marker.
   * left_over receives the part of a match the output limit cut off (remainLen).
   * Below: the look-ahead form (the lean form is decode_symbol_lean, lzb_core_lean.inc).
   * top_up() calls: at most six modelled bits (or 33 direct bits) between two of them. */
  LZB_FN int decode_symbol(uint32_t limit, uint32_t &left_over)
  {
    if (!kAhead) return decode_symbol_lean(limit, left_over);
    const uint32_t ps = total_pos & pb_mask;
    left_over = 0;
    top_up();
    if (rc_bit_at(P(oVar + (state << pb) + ps)) == 0)
    {
      /* ---- literal, :161-196 ---- */
      const uint32_t ctx = ((total_pos & lp_mask) << lc) + (prev_byte >> (8 - lc));
      const uint32_t lit = P(oVar + (56u << pb) + (ctx << 8));    /* plain tree of this context */
      uint32_t sym;
      if (state < 7)
      {
        state -= (state < 4) ? state : 3;
        sym = tree_walk<8>(lit);
      }
      else
      {
        /* matched literal (:176-192): while the decoded bits agree with the match byte the node comes
         * from the context's matched trees (reference index 0x100 + mbit + sym -> matched-tree index
         * mbit + sym); after the first disagreement the rest of the byte uses the plain tree. */
        uint16_t *mlit = mprobs + (ctx << 9);
        uint32_t mb = peek_byte;
        bool agree = true;
        state -= (state < 10) ? 3 : 6;
        sym = 1;
#pragma unroll 1
        for (int i = 0; i < 8; i++)
        {
          if (i == 4) top_up();
          uint32_t b;
          if (agree)
          {
            mb <<= 1;
            const uint32_t mbit = mb & 0x100u;
            b = rc_bit(mlit + mbit + sym);
            agree = (mbit >> 8) == b;
          }
          else
            b = rc_bit_at(lit + (sym << 1));
          sym = (sym << 1) | b;
        }
      }
      prev_byte = sym & 0xFF;
      if (Grp::lane() == 0) dic[dic_pos] = (uint8_t)sym;
      dic_pos++;
      total_pos++;
      return 0;
    }

    uint32_t is_rep = rc_bit_at(P(oIsRep + state));
    if (is_rep == 0)
    {
      state += 12;                                   /* marker: "new distance follows", :204 */
    }
    else
    {
      if (dict_full == 0 && total_pos == 0) return 1;                         /* :210-211 */
      if (rc_bit_at(P(oRepG0 + state)) == 0)
      {
        if (rc_bit_at(P(oVar + (12u << pb) + (state << pb) + ps)) == 0)
        {
          /* ---- short rep, :219-224 ---- */
          Grp::sync();
          const uint32_t s0 = back_index(rep0);
          const uint8_t b = dic[s0];
          const uint32_t s1 = (s0 + 1 == dic_size) ? 0u : s0 + 1;
          peek_byte = (s1 == dic_pos) ? b : dic[s1];
          prev_byte = b;
          if (Grp::lane() == 0) dic[dic_pos] = b;
          dic_pos++;
          total_pos++;
          state = state < 7 ? 9 : 11;
          return 0;
        }
      }
      else
      {
        uint32_t d;
        if (rc_bit_at(P(oRepG1 + state)) == 0) d = rep1;
        else
        {
          if (rc_bit_at(P(oRepG2 + state)) == 0) d = rep2;
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
    top_up();
    {
      const uint32_t ch = P(oLenChoice + 2 * is_rep);
      if (rc_bit_at(ch) == 0)
        len = tree_msb<3>(oVar + (24u << pb) + ((is_rep << 3) << pb) + (ps << 3));
      else if (rc_bit_at(ch + 2) == 0)
        len = 8 + tree_msb<3>(oVar + (40u << pb) + ((is_rep << 3) << pb) + (ps << 3));
      else
        len = 16 + tree_msb<8>(oLenHigh + (is_rep << 8));
    }

    if (state >= 12)
    {
      /* ---- distance, :294-374 ---- */
      top_up();
      uint32_t dist = tree_msb<6>(oDistSlot + ((len < 4 ? len : 3u) << 6));
      if (dist >= 4)
      {
        const uint32_t slot = dist;
        uint32_t nbits = (slot >> 1) - 1;
        dist = 2 | (slot & 1);
        top_up();
        if (slot < 14)
        {
          dist <<= nbits;
          dist |= tree_lsb(oDistSpec + dist - slot - 1, nbits);
        }
        else
        {
          nbits -= 4;
          do { dist = (dist << 1) + rc_direct(); } while (--nbits != 0);
          top_up();
          dist = (dist << 4) | tree_lsb_fixed<4>(oDistAlign);
          if (dist == 0xFFFFFFFFu)
          {
            left_over = len + kEndMarkLen;           /* end marker, :354-359 */
            state -= 12;
            return 2;
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
    return 0;
  }

  /* One trip of the shared loop for a unit in state S_RUN: a symbol, then LzmaDec_DecodeReal's loop
   * test and epilogue (LzmaDec.c:411-423). */
  LZB_FN void step()
  {
    uint32_t left_over;
    const int r = decode_symbol(run_limit, left_over);
    if (r == 1) { run_err = true; st = S_AFTER_RUN; return; }
    if (r == 2 || !(dic_pos < run_limit && in_pos() < run_soft))
    {
      rc_norm();                                                               /* :412 */
      pending = left_over;
      dic_commit = dic_pos;
      run_err = false;
      st = S_AFTER_RUN;
    }
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

  LZB_FN int peek_symbol(uint32_t at_, uint32_t avail) const
  {
    return peek_symbol_impl(probs, mprobs, in, range, code, at_, avail, state, total_pos, lc, lp_mask, pb, prev_byte, peek_byte);
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
      for (uint32_t i = Grp::lane(); i < ns32; i += G) p32[i] = g[i];
      for (uint32_t i = Grp::lane(); i < ng32; i += G) m32[i] = g[ns32 + i];
      Grp::sync();
    }
  }

  LZB_FN void save_state(LzbStreamState *s) const
  {
    Grp::sync();
    if (!want_state_init)
    {
      const uint32_t ns32 = lzb_smem_probs(lc, lp, pb) >> 1, ng32 = lzb_gmem_probs(lc, lp) >> 1;
      uint32_t *g = reinterpret_cast<uint32_t *>(s + 1);
      const uint32_t *p32 = reinterpret_cast<const uint32_t *>(probs);
      const uint32_t *m32 = reinterpret_cast<const uint32_t *>(mprobs);
      for (uint32_t i = Grp::lane(); i < ns32; i += G) g[i] = p32[i];
      for (uint32_t i = Grp::lane(); i < ng32; i += G) g[ns32 + i] = m32[i];
    }
    if (Grp::lane() == 0)
    {
      s->range = range; s->code = code; s->state = state;
      s->rep[0] = rep0; s->rep[1] = rep1; s->rep[2] = rep2; s->rep[3] = rep3;
      s->total_pos = total_pos; s->dict_full = dict_full; s->pending = pending;
      s->flags = (want_rc_init ? 1u : 0u) | (want_state_init ? 2u : 0u);
      s->dic_pos = dic_commit;
      s->held = held_out;
      s->lc = lc; s->lp = lp; s->pb = pb;
      s->l2_phase = (uint32_t)phase; s->l2_control = control;
      s->l2_pack_left = pack_left; s->l2_unpack_left = unpack_left;
      s->l2_flags = (want_dic ? 1u : 0u) | (want_state ? 2u : 0u) | (want_props ? 4u : 0u);
    }
  }

  /* ---------------- unit set-up ---------------- */
  LZB_FN void begin(const uint8_t *packed, uint32_t packed_len, uint8_t *out, uint32_t out_cap,
                    uint32_t dict, uint16_t *model, uint16_t *matched_model)
  {
    held_out = 0;
    mprobs = matched_model;
#ifdef LZB_EMU
    pbase = 0;
#else
    pbase = kSmem ? (uint32_t)__cvta_generic_to_shared(model) : 0u;
#endif
    in = packed;
    in_skew = (uint32_t)((uintptr_t)packed & 3);
    wbase = reinterpret_cast<const uint32_t *>(packed - in_skew);
    wcount = (in_skew + packed_len + 3) >> 2;
    probs = model;
    dic = out; dic_size = out_cap; dic_pos = 0; dic_commit = 0;
    dict_size = dict;
    range = code = 0; nextw = 0;
    if (kAhead) { wlo = whi = 0; nb = 0; widx = 2; k = in_skew; im_pr = 0; }
    else { win = 0; win_left = 4; widx = wcount; }
    state = 0; rep0 = rep1 = rep2 = rep3 = 1;
    prev_byte = peek_byte = 0;
    reset_dic_state(true, true);                                                 /* LzmaDec_Init, :701-705 */
  }

  /* Enter the unit: one-call raw LZMA (LzmaDecode, LzmaDec.c:972-1002 -- LzmaUncompress is that with
   * FINISH_ANY, LzmaLib.c:41-46) or the LZMA2 chunk layer over a whole unit (Lzma2Dec_DecodeToDic,
   * Lzma2Dec.c:170-289).  The unit's chunk headers are re-read here by the decoding group (<= 6 bytes per
   * <= 2 MiB chunk); the HOST walks the same headers once to cut the stream into units -- lzb_containers.cu. */
  LZB_FN void start_unit(const LzbUnit &u, LzbStreamState *stream_state)
  {
    sst = stream_state;
    in_size = u.src_len; dic_limit = u.dst_cap; finish = u.finish;
    raw = u.codec == LZB_CODEC_LZMA;
    at_in = 0; pack_left = 0; unpack_left = 0; control = 0;
    phase = PH_CONTROL;
    want_dic = want_state = want_props = true;
    status = LZB_ST_INHERIT; res = LZB_SZ_OK;
    ctx_stale = false; run_err = false;
    st = S_L2_TOP;
    if (raw)
    {
      if (!sst && in_size < 5) { res = LZB_SZ_ERROR_INPUT_EOF; status = LZB_ST_NOT_SPECIFIED; st = S_UNIT_END; return; }
      set_props(u.lc, u.lp, u.pb);
      control = 0x80; phase = PH_DATA_MORE; pack_left = in_size; unpack_left = 0xFFFFFFFFu;
    }
    else
      set_props(4, 0, 0);                              /* Lzma2Dec_GetOldProps, Lzma2Dec.c:61-73 */
    if (sst)
    {
      /* resume: CLzmaDec / CLzma2Dec as the previous call left them (host keeps the scalars authoritative) */
      load_state(sst);
      held_out = sst->held;
      if (!raw)
      {
        phase = (int)sst->l2_phase; control = sst->l2_control;
        pack_left = sst->l2_pack_left + sst->held;       /* bytes parked in tempBuf were already charged */
        unpack_left = sst->l2_unpack_left;
        want_dic = (sst->l2_flags & 1) != 0; want_state = (sst->l2_flags & 2) != 0; want_props = (sst->l2_flags & 4) != 0;
      }
    }
  }

  /* ---------------- the driver machine ----------------
   * Runs the unit until it needs symbols decoded (st == S_RUN, bounds in run_limit / run_soft) or is
   * finished (st == S_UNIT_END; res / status / at_in / dic_commit hold the reference's return values).
   *   S_L2_TOP     top of Lzma2Dec_DecodeToDic's loop (Lzma2Dec.c:177-285), header machine (:98-157) folded in
   *   S_TD_ENTER   entry of LzmaDec_DecodeToDic (LzmaDec.c:722-726)
   *   S_TD_TOP     top of its loop (:728-797), for an input that is completely present: the `tempBuf`
   *                continuation branch (:804-833) is only reachable across calls; the host re-presents the
   *                parked bytes instead (lzb_state.cu)
   *   S_SPAN_TOP   top of LzmaDec_DecodeReal2's loop (:456-464)
   *   S_AFTER_RUN  the rest of that loop (:465-476) and of DecodeToDic's (:799-802)
   *   S_TD_RET     back in Lzma2Dec_DecodeToDic after the call (Lzma2Dec.c:264-284)                      */
  LZB_FN void advance()
  {
    /* Resuming with bytes the previous call had parked (CLzmaDec.tempBuf): the reference decodes exactly ONE
     * symbol out of tempBuf and commits it before it goes back to the caller's buffer (LzmaDec.c:804-833).
     * The bytes are contiguous here, but that commit point is observable when a later symbol of the same
     * call fails (srcLen / destLen of the failing call), so the first run is cut after one symbol too. */
    bool one_symbol_run = false;
    for (;;)
    {
      switch (st)
      {
        case S_L2_TOP:
        {
          if (phase == PH_FINISHED) { status = LZB_ST_FINISHED_WITH_MARK; st = S_UNIT_END; return; }
          at = dic_commit;
          if (phase == PH_ERROR) { res = LZB_SZ_ERROR_DATA; st = S_UNIT_END; return; }
          /* finish == 2 (host-internal "unit" mode): like FINISH_END inside the unit's last chunk, so that its
           * proper termination is still verified, but stop at the unit's end instead of reading the next header */
          if (!raw && at == dic_limit && (finish == 0 || (finish == 2 && phase == PH_CONTROL)))
          {
            if (finish == 0) status = LZB_ST_NOT_FINISHED;      /* unit mode: a boundary is not a stop of the stream */
            st = S_UNIT_END;
            return;
          }
          if (phase != PH_DATA && phase != PH_DATA_MORE)
          {
            if (at_in == in_size) { status = LZB_ST_NEEDS_MORE_INPUT; st = S_UNIT_END; return; }
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
                const uint32_t lc_ = b % 9; b /= 9;
                const uint32_t lp_ = b % 5;
                if (lc_ + lp_ > 4) { phase = PH_ERROR; break; }
                set_props(lc_, lp_, b / 5);
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
            if (at_in == in_size) { status = LZB_ST_NEEDS_MORE_INPUT; st = S_UNIT_END; return; }
            if (phase == PH_DATA)
            {
              const bool reset_dic = control == 1;
              if (reset_dic) want_props = want_state = true;
              else if (want_dic) { res = LZB_SZ_ERROR_DATA; st = S_UNIT_END; return; }
              want_dic = false;
              reset_dic_state(reset_dic, false);
            }
            if (in_now > out_now) in_now = out_now;
            if (in_now == 0) { res = LZB_SZ_ERROR_DATA; st = S_UNIT_END; return; }
            group_copy_in<G>(dic + dic_pos, in + at_in, in_now);     /* :159-166 */
            dic_pos += in_now;
            dic_commit = dic_pos;
            if (dict_full == 0 && dict_size - total_pos <= in_now) dict_full = dict_size;
            total_pos += in_now;
            at_in += in_now;
            unpack_left -= in_now;
            phase = unpack_left == 0 ? PH_CONTROL : PH_DATA_MORE;
            continue;
          }
          /* LZMA chunk, :241-284 */
          if (phase == PH_DATA)
          {
            const uint32_t m = (control >> 5) & 3;
            const bool reset_dic = m == 3, reset_state = m > 0;
            if ((!reset_dic && want_dic) || (!reset_state && want_state)) { res = LZB_SZ_ERROR_DATA; st = S_UNIT_END; return; }
            reset_dic_state(reset_dic, reset_state);
            want_dic = want_state = false;
            phase = PH_DATA_MORE;
          }
          if (in_now > pack_left) in_now = pack_left;
          td_limit = at + out_now; td_src = at_in; td_avail = in_now; td_finish = mode;
          st = S_TD_ENTER;
          continue;
        }

        case S_TD_ENTER:
          ctx_stale = false;
          td_used = 0;
          one_symbol_run = held_out != 0 && !want_rc_init;      /* held_out still holds what the last call parked */
          held_out = 0;
          if (!want_rc_init) { rc_seek(td_src); ctx_stale = true; }     /* continuing a chunk / a suspended stream */
          flush_pending(td_limit);
          status = LZB_ST_NOT_SPECIFIED;
          st = S_TD_TOP;
          continue;

        case S_TD_TOP:
        {
          if (pending == kEndMarkLen)                                                /* :835-837 */
          {
            if (code == 0) status = LZB_ST_FINISHED_WITH_MARK;
            td_res = code == 0 ? LZB_SZ_OK : LZB_SZ_ERROR_DATA;
            st = S_TD_RET;
            continue;
          }
          bool must_end = false;
          if (want_rc_init)                                                          /* :732-746 */
          {
            if (td_avail < 5)
            {
              td_used += td_avail; held_out = td_avail; status = LZB_ST_NEEDS_MORE_INPUT;
              td_res = LZB_SZ_OK; st = S_TD_RET;
              continue;
            }
            td_used += 5;
            if (in[td_src] != 0) { td_res = LZB_SZ_ERROR_DATA; st = S_TD_RET; continue; }
            code = ((uint32_t)in[td_src + 1] << 24) | ((uint32_t)in[td_src + 2] << 16) |
                   ((uint32_t)in[td_src + 3] << 8) | in[td_src + 4];
            range = 0xFFFFFFFFu;
            want_rc_init = false;
            td_src += 5; td_avail -= 5;
            rc_seek(td_src);
            ctx_stale = true;
          }
          if (dic_pos >= td_limit)                                                   /* :749-767 */
          {
            td_res = LZB_SZ_OK;
            if (pending == 0 && code == 0) { status = LZB_ST_MAYBE_FINISHED_WITHOUT_MARK; st = S_TD_RET; continue; }
            if (td_finish == 0) { status = LZB_ST_NOT_FINISHED; st = S_TD_RET; continue; }
            if (pending != 0) { status = LZB_ST_NOT_FINISHED; td_res = LZB_SZ_ERROR_DATA; st = S_TD_RET; continue; }
            must_end = true;
          }
          if (want_state_init) init_models();
          if (ctx_stale) { Grp::sync(); refresh_context(); ctx_stale = false; }
          if (td_avail < kLookahead || must_end)                                     /* :776-793 */
          {
            const int k = peek_symbol(td_src, td_avail);
            if (k == kSymStarved)
            {
              td_used += td_avail; held_out = td_avail; status = LZB_ST_NEEDS_MORE_INPUT;
              td_res = LZB_SZ_OK; st = S_TD_RET;
              continue;
            }
            if (must_end && k != kSymMatch) { status = LZB_ST_NOT_FINISHED; td_res = LZB_SZ_ERROR_DATA; st = S_TD_RET; continue; }
            run_soft = td_src;
          }
          else
            run_soft = one_symbol_run ? td_src : td_src + td_avail - kLookahead;
          one_symbol_run = false;
          st = S_SPAN_TOP;
          continue;
        }

        case S_SPAN_TOP:                                                             /* LzmaDec.c:456-464 */
        {
          uint32_t lim = td_limit;
          if (dict_full == 0)
          {
            const uint32_t room = dict_size - total_pos;
            if (td_limit - dic_pos > room) lim = dic_pos + room;
          }
          run_limit = lim;
          st = S_RUN;
          return;
        }

        case S_AFTER_RUN:                                                            /* :465-476, :799-802 */
        {
          if (run_err) { td_res = LZB_SZ_ERROR_DATA; st = S_TD_RET; continue; }
          if (total_pos >= dict_size) dict_full = dict_size;
          flush_pending(td_limit);
          if (dic_pos < td_limit && in_pos() < run_soft && pending < kEndMarkLen) { st = S_SPAN_TOP; continue; }
          if (pending > kEndMarkLen) pending = kEndMarkLen;
          const uint32_t n = in_pos() - td_src;
          td_used += n; td_src += n; td_avail -= n;
          st = S_TD_TOP;
          continue;
        }

        case S_TD_RET:                                                               /* Lzma2Dec.c:264-284 */
        {
          res = td_res;
          at_in += td_used;
          if (raw)
          {
            /* one-call LzmaDecode folds NEEDS_MORE_INPUT into SZ_ERROR_INPUT_EOF (LzmaDec.c:996-997) */
            if (!sst && res == LZB_SZ_OK && status == LZB_ST_NEEDS_MORE_INPUT) res = LZB_SZ_ERROR_INPUT_EOF;
            st = S_UNIT_END;
            return;
          }
          pack_left -= td_used;
          const uint32_t made = dic_commit - at;
          unpack_left -= made;
          if (res != LZB_SZ_OK) { st = S_UNIT_END; return; }
          if (status == LZB_ST_NEEDS_MORE_INPUT) { st = S_UNIT_END; return; }
          if (td_used == 0 && made == 0)
          {
            if (status != LZB_ST_MAYBE_FINISHED_WITHOUT_MARK || unpack_left != 0 || pack_left != 0)
            { res = LZB_SZ_ERROR_DATA; st = S_UNIT_END; return; }
            phase = PH_CONTROL;
          }
          if (status == LZB_ST_MAYBE_FINISHED_WITHOUT_MARK) status = LZB_ST_NOT_FINISHED;
          st = S_L2_TOP;
          continue;
        }

        default:   /* S_RUN, S_UNIT_END, S_FETCH, S_EXIT: nothing for the driver to do */
          return;
      }
    }
  }

  /* the reference's return values for the unit that just reached S_UNIT_END */
  LZB_FN void finish_unit(LzbResult &r)
  {
    r.res = res; r.status = status; r.src_used = at_in; r.dst_done = dic_commit; r.check = 0;
    if (sst) save_state(sst);
  }
};

}  // namespace lzb
