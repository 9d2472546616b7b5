/* lzma_oracle.c -- CPU restatement of the reference LZMA decoder.  TEST INFRASTRUCTURE ONLY
 * (see oracle.h for the contract; parity is pinned against oracle/_ref/libref.so).
 *
 * Shape: every LZMA symbol is first PARSED from the range coder into a small description
 * (kind, literal byte / length, distance or rep index) and then APPLIED to the window.
 * Running the parser with adaptation switched off and bounds checks on gives the
 * side-effect-free look-ahead that the reference hand-codes as LzmaDec_TryDummy.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

#define RC_TOP        (1u << 24)        /* kTopValue,            LzmaDec.c:8-9   */
#define PROB_BITS     11                /* kNumBitModelTotalBits LzmaDec.c:11    */
#define PROB_ONE      (1u << PROB_BITS)
#define PROB_SHIFT    5                 /* kNumMoveBits          LzmaDec.c:13    */
#define RC_HEADER     5                 /* RC_INIT_SIZE          LzmaDec.c:15    */
#define N_STATES      12
#define N_LIT_STATES  7
#define MIN_DICT      (1u << 12)        /* LZMA_DIC_MIN          LzmaDec.c:116   */

/* probability table map, LzmaDec.c:60-110 (posState stride fixed at 16) */
enum
{
  T_IS_MATCH   = 0,       /* [state][posState]            12*16 */
  T_IS_REP     = 192,     /* [state]                            */
  T_IS_REP_G0  = 204,
  T_IS_REP_G1  = 216,
  T_IS_REP_G2  = 228,
  T_REP0_LONG  = 240,     /* [state][posState]            12*16 */
  T_DIST_SLOT  = 432,     /* [min(len,3)][64]                   */
  T_DIST_SPEC  = 688,     /* reverse trees for slots 4..13  114 */
  T_DIST_ALIGN = 802,     /* 16                                 */
  T_LEN_MATCH  = 818,     /* 514                                */
  T_LEN_REP    = 1332,    /* 514                                */
  T_LITERAL    = 1846     /* 768 << (lc+lp)                     */
};
enum { LEN_CHOICE = 0, LEN_CHOICE2 = 1, LEN_LOW = 2, LEN_MID = 2 + (16 << 3), LEN_HIGH = 2 + (32 << 3) };

enum { SYM_LITERAL, SYM_MATCH, SYM_REP, SYM_SHORTREP, SYM_STARVED };

typedef struct
{
  int kind;
  unsigned byte;        /* SYM_LITERAL */
  unsigned len_index;   /* length - 2 */
  uint32_t dist;        /* SYM_MATCH: distance - 1; 0xFFFFFFFF = end marker */
  unsigned rep_index;   /* SYM_REP */
} Sym;

/* ---- range coder ---------------------------------------------------------------- */
typedef struct
{
  uint32_t range, code;
  const uint8_t *p, *end;
  int adapt;            /* 1: update probabilities (real decode), 0: look-ahead */
  int starved;          /* look-ahead needed a byte past `end` */
} Rc;

/* LzmaDec.c:17 (NORMALIZE) and :47 (NORMALIZE_CHECK) */
static void rc_fill(Rc *rc)
{
  if (rc->range < RC_TOP)
  {
    if (rc->p >= rc->end) { rc->starved = 1; return; }
    rc->range <<= 8;
    rc->code = (rc->code << 8) | *rc->p++;
  }
}

/* LzmaDec.c:19-21 / :49-51 -- one adaptive binary decision */
static unsigned rc_bit(Rc *rc, uint16_t *prob)
{
  uint32_t p = *prob, bound;
  rc_fill(rc);
  bound = (rc->range >> PROB_BITS) * p;
  if (rc->code < bound)
  {
    rc->range = bound;
    if (rc->adapt) *prob = (uint16_t)(p + ((PROB_ONE - p) >> PROB_SHIFT));
    return 0;
  }
  rc->range -= bound;
  rc->code -= bound;
  if (rc->adapt) *prob = (uint16_t)(p - (p >> PROB_SHIFT));
  return 1;
}

/* LzmaDec.c:325-334 -- equiprobable bit, with the reference's exact wrap-around arithmetic */
static unsigned rc_direct(Rc *rc)
{
  uint32_t neg;
  rc_fill(rc);
  rc->range >>= 1;
  rc->code -= rc->range;
  neg = 0u - (rc->code >> 31);
  rc->code += rc->range & neg;
  return neg + 1;
}

/* LzmaDec.c:27-29 */
static unsigned tree_msb_first(Rc *rc, uint16_t *t, unsigned bits)
{
  unsigned m = 1, i;
  for (i = 0; i < bits; i++) m = (m << 1) | rc_bit(rc, t + m);
  return m - (1u << bits);
}

/* LzmaDec.c:309-318, :347-353 */
static unsigned tree_lsb_first(Rc *rc, uint16_t *t, unsigned bits)
{
  unsigned m = 1, v = 0, i;
  for (i = 0; i < bits; i++)
  {
    unsigned b = rc_bit(rc, t + m);
    m = (m << 1) | b;
    v |= b << i;
  }
  return v;
}

/* LzmaDec.c:261-292 */
static unsigned parse_len(Rc *rc, uint16_t *t, unsigned pos_state)
{
  if (!rc_bit(rc, t + LEN_CHOICE)) return tree_msb_first(rc, t + LEN_LOW + (pos_state << 3), 3);
  if (!rc_bit(rc, t + LEN_CHOICE2)) return 8 + tree_msb_first(rc, t + LEN_MID + (pos_state << 3), 3);
  return 16 + tree_msb_first(rc, t + LEN_HIGH, 8);
}

/* LzmaDec.c:294-361 */
static uint32_t parse_dist(Rc *rc, uint16_t *probs, unsigned len_index)
{
  unsigned slot = tree_msb_first(rc, probs + T_DIST_SLOT + ((len_index < 4 ? len_index : 3) << 6), 6);
  unsigned nbits;
  uint32_t d;
  if (slot < 4) return slot;
  nbits = (slot >> 1) - 1;
  d = 2 | (slot & 1);
  if (slot < 14)
  {
    d <<= nbits;
    return d | tree_lsb_first(rc, probs + T_DIST_SPEC + d - slot - 1, nbits);
  }
  for (nbits -= 4; nbits != 0; nbits--) d = (d << 1) + rc_direct(rc);
  d <<= 4;
  return d | tree_lsb_first(rc, probs + T_DIST_ALIGN, 4);
}

static uint8_t window_back(const OraLzma *d, uint32_t back)
{
  /* LzmaDec.c:178,220: (dicPos - rep0) + (dicPos < rep0 ? dicBufSize : 0) */
  return d->dic[d->dic_pos - back + (d->dic_pos < back ? d->dic_size : 0)];
}

/* LzmaDec.c:161-192 */
static unsigned parse_literal(const OraLzma *d, Rc *rc)
{
  uint16_t *t = d->probs + T_LITERAL;
  unsigned sym = 1;
  if (d->dict_full != 0 || d->total_pos != 0)
  {
    unsigned prev = d->dic[(d->dic_pos == 0 ? d->dic_size : d->dic_pos) - 1];
    t += 768u * (((d->total_pos & ((1u << d->lp) - 1)) << d->lc) + (prev >> (8 - d->lc)));
  }
  if (d->state < N_LIT_STATES)
  {
    while (sym < 0x100) sym = (sym << 1) | rc_bit(rc, t + sym);
  }
  else
  {
    unsigned match = window_back(d, d->rep[0]);
    unsigned offs = 0x100;
    while (sym < 0x100)
    {
      unsigned mbit, bit;
      match <<= 1;
      mbit = match & offs;
      bit = rc_bit(rc, t + offs + mbit + sym);
      sym = (sym << 1) | bit;
      offs &= bit ? mbit : ~mbit;
    }
  }
  return sym & 0xFF;
}

/* Symbol grammar, LzmaDec.c:153-361 (and :496-671 for the look-ahead twin). */
static void parse_symbol(const OraLzma *d, Rc *rc, Sym *s)
{
  uint16_t *probs = d->probs;
  unsigned pos_state = d->total_pos & ((1u << d->pb) - 1);
  unsigned st = d->state;
  s->byte = 0; s->len_index = 0; s->dist = 0; s->rep_index = 0;
  if (!rc_bit(rc, probs + T_IS_MATCH + (st << 4) + pos_state))
  {
    s->kind = SYM_LITERAL;
    s->byte = parse_literal(d, rc);
    return;
  }
  if (!rc_bit(rc, probs + T_IS_REP + st))
  {
    s->kind = SYM_MATCH;
    s->len_index = parse_len(rc, probs + T_LEN_MATCH, pos_state);
    s->dist = parse_dist(rc, probs, s->len_index);
    return;
  }
  if (!rc_bit(rc, probs + T_IS_REP_G0 + st))
  {
    if (!rc_bit(rc, probs + T_REP0_LONG + (st << 4) + pos_state)) { s->kind = SYM_SHORTREP; return; }
    s->rep_index = 0;
  }
  else if (!rc_bit(rc, probs + T_IS_REP_G1 + st)) s->rep_index = 1;
  else if (!rc_bit(rc, probs + T_IS_REP_G2 + st)) s->rep_index = 2;
  else s->rep_index = 3;
  s->kind = SYM_REP;
  s->len_index = parse_len(rc, probs + T_LEN_REP, pos_state);
}

/* LzmaDec.c:376-408: byte-by-byte forward copy, possibly cut by the output limit */
static int copy_match(OraLzma *d, unsigned len, size_t limit, unsigned *left_over)
{
  size_t room, pos;
  unsigned cur;
  if (limit == d->dic_pos) return ORA_ERROR_DATA;
  room = limit - d->dic_pos;
  cur = room < len ? (unsigned)room : len;
  pos = d->dic_pos - d->rep[0] + (d->dic_pos < d->rep[0] ? d->dic_size : 0);
  d->total_pos += cur;
  *left_over = len - cur;
  while (cur-- != 0)
  {
    d->dic[d->dic_pos++] = d->dic[pos];
    if (++pos == d->dic_size) pos = 0;
  }
  return ORA_OK;
}

/* State transitions: LzmaDec.c:172,180,223,258,373 */
static int apply_symbol(OraLzma *d, const Sym *s, size_t limit, unsigned *left_over, int *end_mark)
{
  unsigned st = d->state;
  *end_mark = 0;
  switch (s->kind)
  {
    case SYM_LITERAL:
      d->dic[d->dic_pos++] = (uint8_t)s->byte;
      d->total_pos++;
      d->state = st < 4 ? 0 : (st < 10 ? st - 3 : st - 6);
      return ORA_OK;
    case SYM_MATCH:
      if (s->dist == 0xFFFFFFFFu)
      {
        *left_over = s->len_index + ORA_LEN_END_MARK;     /* LzmaDec.c:354-359 */
        *end_mark = 1;
        return ORA_OK;
      }
      d->rep[3] = d->rep[2]; d->rep[2] = d->rep[1]; d->rep[1] = d->rep[0];
      d->rep[0] = s->dist + 1;
      if (d->dict_full == 0 ? s->dist >= d->total_pos : s->dist >= d->dict_full)   /* :366-372 */
        return ORA_ERROR_DATA;
      d->state = st < N_LIT_STATES ? 7 : 10;
      return copy_match(d, s->len_index + 2, limit, left_over);
    case SYM_SHORTREP:
      if (d->dict_full == 0 && d->total_pos == 0) return ORA_ERROR_DATA;            /* :210-211 */
      d->dic[d->dic_pos] = window_back(d, d->rep[0]);
      d->dic_pos++;
      d->total_pos++;
      d->state = st < N_LIT_STATES ? 9 : 11;
      return ORA_OK;
    default: /* SYM_REP */
      if (d->dict_full == 0 && d->total_pos == 0) return ORA_ERROR_DATA;
      if (s->rep_index != 0)
      {
        uint32_t pick = d->rep[s->rep_index];
        unsigned k;
        for (k = s->rep_index; k > 0; k--) d->rep[k] = d->rep[k - 1];
        d->rep[0] = pick;
      }
      d->state = st < N_LIT_STATES ? 8 : 11;
      return copy_match(d, s->len_index + 2, limit, left_over);
  }
}

/* LzmaDec.c:131-426: decode symbols until the window limit or the soft input limit.
 * All progress is committed only on success (the reference keeps it in locals). */
static int decode_run(OraLzma *d, Rc *rc, size_t limit, const uint8_t *soft_end)
{
  OraLzma keep = *d;
  Rc w = *rc;
  unsigned left_over = 0;
  w.adapt = 1;
  do
  {
    Sym s;
    int end_mark;
    parse_symbol(d, &w, &s);
    if (apply_symbol(d, &s, limit, &left_over, &end_mark) != ORA_OK)
    {
      uint16_t *probs = d->probs;
      *d = keep;
      d->probs = probs;
      return ORA_ERROR_DATA;
    }
    if (end_mark) break;
  }
  while (d->dic_pos < limit && w.p < soft_end);
  rc_fill(&w);                                  /* :412 */
  d->range = w.range;
  d->code = w.code;
  d->pending = left_over;
  *rc = w;
  return ORA_OK;
}

/* LzmaDec.c:428-452 */
static void flush_pending(OraLzma *d, size_t limit)
{
  if (d->pending != 0 && d->pending < ORA_LEN_END_MARK)
  {
    unsigned len = d->pending;
    if (limit - d->dic_pos < len) len = (unsigned)(limit - d->dic_pos);
    if (d->dict_full == 0 && d->dict_size - d->total_pos <= len) d->dict_full = d->dict_size;
    d->total_pos += len;
    d->pending -= len;
    while (len-- != 0)
    {
      d->dic[d->dic_pos] = window_back(d, d->rep[0]);
      d->dic_pos++;
    }
  }
}

/* LzmaDec.c:454-477 */
static int decode_span(OraLzma *d, Rc *rc, size_t limit, const uint8_t *soft_end)
{
  do
  {
    size_t lim = limit;
    if (d->dict_full == 0)
    {
      uint32_t room = d->dict_size - d->total_pos;
      if (limit - d->dic_pos > room) lim = d->dic_pos + room;
    }
    if (decode_run(d, rc, lim, soft_end) != ORA_OK) return ORA_ERROR_DATA;
    if (d->total_pos >= d->dict_size) d->dict_full = d->dict_size;
    flush_pending(d, limit);
  }
  while (d->dic_pos < limit && rc->p < soft_end && d->pending < ORA_LEN_END_MARK);
  if (d->pending > ORA_LEN_END_MARK) d->pending = ORA_LEN_END_MARK;
  return ORA_OK;
}

/* LzmaDec.c:487-675 -- would one more whole symbol fit in [src, src+avail)? */
static int peek_symbol(const OraLzma *d, const uint8_t *src, size_t avail)
{
  Rc rc;
  Sym s;
  rc.range = d->range; rc.code = d->code; rc.p = src; rc.end = src + avail; rc.adapt = 0; rc.starved = 0;
  parse_symbol(d, &rc, &s);
  rc_fill(&rc);                                 /* :673 */
  if (rc.starved) return SYM_STARVED;
  return s.kind == SYM_SHORTREP ? SYM_REP : s.kind;
}

/* LzmaDec.c:707-717 */
static void init_models(OraLzma *d)
{
  uint32_t n = T_LITERAL + (768u << (d->lc + d->lp)), i;
  for (i = 0; i < n; i++) d->probs[i] = PROB_ONE >> 1;
  d->rep[0] = d->rep[1] = d->rep[2] = d->rep[3] = 1;
  d->state = 0;
  d->want_state_init = 0;
}

void ora_lzma_reset_dic_state(OraLzma *d, int reset_dic, int reset_state)
{
  d->want_rc_init = 1;
  d->pending = 0;
  d->held = 0;
  if (reset_dic) { d->total_pos = 0; d->dict_full = 0; d->want_state_init = 1; }
  if (reset_state) d->want_state_init = 1;
}

void ora_lzma_reset(OraLzma *d)
{
  d->dic_pos = 0;
  ora_lzma_reset_dic_state(d, 1, 1);
}

int ora_lzma_to_dic(OraLzma *d, size_t dic_limit, const uint8_t *src, size_t *src_len, int finish, int *status)
{
  size_t avail = *src_len;
  *src_len = 0;
  flush_pending(d, dic_limit);
  *status = ORA_ST_NOT_SPECIFIED;

  while (d->pending != ORA_LEN_END_MARK)
  {
    int must_end = 0;
    Rc rc;

    if (d->want_rc_init)                                          /* :732-746 */
    {
      while (avail > 0 && d->held < RC_HEADER) { d->hold[d->held++] = *src++; (*src_len)++; avail--; }
      if (d->held < RC_HEADER) { *status = ORA_ST_NEEDS_MORE_INPUT; return ORA_OK; }
      if (d->hold[0] != 0) return ORA_ERROR_DATA;
      d->code = ((uint32_t)d->hold[1] << 24) | ((uint32_t)d->hold[2] << 16) | ((uint32_t)d->hold[3] << 8) | d->hold[4];
      d->range = 0xFFFFFFFFu;                                     /* :678-683 */
      d->want_rc_init = 0;
      d->held = 0;
    }

    if (d->dic_pos >= dic_limit)                                  /* :749-767 */
    {
      if (d->pending == 0 && d->code == 0) { *status = ORA_ST_MAYBE_FINISHED_WITHOUT_MARK; return ORA_OK; }
      if (finish == ORA_FINISH_ANY) { *status = ORA_ST_NOT_FINISHED; return ORA_OK; }
      if (d->pending != 0) { *status = ORA_ST_NOT_FINISHED; return ORA_ERROR_DATA; }
      must_end = 1;
    }

    if (d->want_state_init) init_models(d);

    rc.adapt = 1; rc.starved = 0;
    if (d->held == 0)                                             /* :772-803 */
    {
      const uint8_t *soft_end;
      size_t used;
      if (avail < ORA_LOOKAHEAD || must_end)
      {
        int k = peek_symbol(d, src, avail);
        if (k == SYM_STARVED)
        {
          memcpy(d->hold, src, avail);
          d->held = (unsigned)avail;
          *src_len += avail;
          *status = ORA_ST_NEEDS_MORE_INPUT;
          return ORA_OK;
        }
        if (must_end && k != SYM_MATCH) { *status = ORA_ST_NOT_FINISHED; return ORA_ERROR_DATA; }
        soft_end = src;
      }
      else
        soft_end = src + avail - ORA_LOOKAHEAD;
      rc.range = d->range; rc.code = d->code; rc.p = src; rc.end = src + avail;
      if (decode_span(d, &rc, dic_limit, soft_end) != ORA_OK) return ORA_ERROR_DATA;
      used = (size_t)(rc.p - src);
      *src_len += used; src += used; avail -= used;
    }
    else                                                          /* :804-833 */
    {
      unsigned have = d->held, look = 0;
      while (have < ORA_LOOKAHEAD && look < avail) d->hold[have++] = src[look++];
      d->held = have;
      if (have < ORA_LOOKAHEAD || must_end)
      {
        int k = peek_symbol(d, d->hold, have);
        if (k == SYM_STARVED) { *src_len += look; *status = ORA_ST_NEEDS_MORE_INPUT; return ORA_OK; }
        if (must_end && k != SYM_MATCH) { *status = ORA_ST_NOT_FINISHED; return ORA_ERROR_DATA; }
      }
      rc.range = d->range; rc.code = d->code; rc.p = d->hold; rc.end = d->hold + have;
      if (decode_span(d, &rc, dic_limit, d->hold) != ORA_OK) return ORA_ERROR_DATA;
      look -= have - (unsigned)(rc.p - d->hold);
      *src_len += look; src += look; avail -= look;
      d->held = 0;
    }
  }
  if (d->code == 0) *status = ORA_ST_FINISHED_WITH_MARK;          /* :835-837 */
  return d->code == 0 ? ORA_OK : ORA_ERROR_DATA;
}

int ora_lzma_to_buf(OraLzma *d, uint8_t *dest, size_t *dest_len, const uint8_t *src, size_t *src_len,
                    int finish, int *status)
{
  size_t out_room = *dest_len, in_left = *src_len;
  *src_len = *dest_len = 0;
  for (;;)
  {
    size_t in_now = in_left, start, stop, made;
    int mode, res;
    if (d->dic_pos == d->dic_size) d->dic_pos = 0;                /* ring wrap, :850 */
    start = d->dic_pos;
    if (out_room > d->dic_size - start) { stop = d->dic_size; mode = ORA_FINISH_ANY; }
    else { stop = start + out_room; mode = finish; }
    res = ora_lzma_to_dic(d, stop, src, &in_now, mode, status);
    src += in_now; in_left -= in_now; *src_len += in_now;
    made = d->dic_pos - start;
    memcpy(dest, d->dic + start, made);
    dest += made; out_room -= made; *dest_len += made;
    if (res != ORA_OK) return res;
    if (made == 0 || out_room == 0) return ORA_OK;
  }
}

int ora_lzma_props(OraLzma *d, const uint8_t *props, unsigned size)
{
  uint32_t ds;
  unsigned b;
  if (size < 5) return ORA_ERROR_UNSUPPORTED;
  ds = props[1] | ((uint32_t)props[2] << 8) | ((uint32_t)props[3] << 16) | ((uint32_t)props[4] << 24);
  d->dict_size = ds < MIN_DICT ? MIN_DICT : ds;
  b = props[0];
  if (b >= 9 * 5 * 5) return ORA_ERROR_UNSUPPORTED;
  d->lc = b % 9; b /= 9;
  d->pb = b / 5;
  d->lp = b % 5;
  return ORA_OK;
}

static int grow_probs(OraLzma *d)
{
  uint32_t n = T_LITERAL + (768u << (d->lc + d->lp));
  if (d->probs == NULL || n != d->num_probs)
  {
    free(d->probs);
    d->probs = (uint16_t *)malloc((size_t)n * sizeof(uint16_t));
    d->num_probs = n;
    if (!d->probs) return ORA_ERROR_MEM;
  }
  return ORA_OK;
}

int ora_lzma_alloc_probs(OraLzma *d, const uint8_t *props, unsigned size)
{
  OraLzma t = *d;
  int r = ora_lzma_props(&t, props, size);
  if (r != ORA_OK) return r;
  d->lc = t.lc; d->lp = t.lp; d->pb = t.pb; d->dict_size = t.dict_size;
  return grow_probs(d);
}

int ora_lzma_alloc(OraLzma *d, const uint8_t *props, unsigned size)
{
  int r = ora_lzma_alloc_probs(d, props, size);
  if (r != ORA_OK) return r;
  if (d->dic == NULL || !d->owns_dic || d->dic_size != d->dict_size)
  {
    if (d->owns_dic) free(d->dic);
    d->dic = (uint8_t *)malloc(d->dict_size);
    d->owns_dic = 1;
    if (!d->dic) { free(d->probs); d->probs = NULL; return ORA_ERROR_MEM; }
  }
  d->dic_size = d->dict_size;
  return ORA_OK;
}

void ora_lzma_free(OraLzma *d)
{
  free(d->probs); d->probs = NULL;
  if (d->owns_dic) free(d->dic);
  d->dic = NULL; d->owns_dic = 0;
}

int ora_lzma_decode(uint8_t *dest, size_t *dest_len, const uint8_t *src, size_t *src_len,
                    const uint8_t *props, unsigned props_size, int finish, int *status)
{
  OraLzma d;
  size_t in_size = *src_len, out_size = *dest_len;
  int res;
  *src_len = *dest_len = 0;
  if (in_size < RC_HEADER) return ORA_ERROR_INPUT_EOF;
  memset(&d, 0, sizeof d);
  res = ora_lzma_alloc_probs(&d, props, props_size);
  if (res != ORA_OK) return res;
  d.dic = dest;
  d.dic_size = out_size;
  ora_lzma_reset(&d);
  *src_len = in_size;
  res = ora_lzma_to_dic(&d, out_size, src, src_len, finish, status);
  if (res == ORA_OK && *status == ORA_ST_NEEDS_MORE_INPUT) res = ORA_ERROR_INPUT_EOF;
  *dest_len = d.dic_pos;
  free(d.probs);
  return res;
}

int ora_lzma_uncompress(uint8_t *dest, size_t *dest_len, const uint8_t *src, size_t *src_len,
                        const uint8_t *props, size_t props_size)
{
  int status = ORA_ST_NOT_SPECIFIED;
  return ora_lzma_decode(dest, dest_len, src, src_len, props, (unsigned)props_size, ORA_FINISH_ANY, &status);
}
