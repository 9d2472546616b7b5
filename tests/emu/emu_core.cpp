/* emu_core.cpp -- TEST-ONLY host build of the device decode logic.
 *
 * Compiles lzma-sdk-zliblike_b200/csrc/lzb_core.cuh + lzb_post.cuh with -DLZB_EMU (one emulated lane per
 * unit, G = 1) so that `pytest -m "not gpu"` can exercise the kernel's symbol grammar, the driver state
 * machine (finish modes, look-ahead, LZMA2 chunk layer, resume) and the post-filters against the oracle on
 * machines without a GPU.  It is NOT a CPU fallback: nothing in the product package loads this library,
 * and the product fails loudly when the CUDA library or a device is missing.
 */
#define LZB_EMU 1
#include <stdlib.h>
#include <vector>

#include "lzb_core.cuh"
#include "lzb_post.cuh"

/* filter / check only, like lzb_post_only_kernel */
static int g_post_lanes = 0;          /* 1: the per-lane x86 converter of the lean decode kernels instead of the tile scan */
extern "C" void emu_post_use_lanes(int on) { g_post_lanes = on; }

extern "C" int emu_post_units(uint8_t *data, const LzbUnit *units, LzbResult *results, uint32_t n)
{
  std::vector<uint8_t> scratch(lzb::kPostScratchBytes);
  for (uint32_t i = 0; i < n; i++)
  {
    const LzbUnit &u = units[i];
    LzbResult r;
    r.res = LZB_SZ_OK; r.status = LZB_ST_NOT_SPECIFIED; r.src_used = 0; r.dst_done = u.dst_cap; r.check = 0;
    if (g_post_lanes) lzb::post_process<1, true, false>(u, r, data + u.dst_off, scratch.data(), scratch.data(), lzb::kPostScratchBytes);
    else lzb::post_process<1, true, true>(u, r, data + u.dst_off, scratch.data(), scratch.data(), lzb::kPostScratchBytes);
    results[i] = r;
  }
  return 0;
}

/* the kernel's loop, for one group: fetch -> advance -> (step)* -> ... -> finish */
template <bool kAhead>
static int decode_units(const uint8_t *packed, uint8_t *out, const LzbUnit *units, LzbResult *results, uint32_t n)
{
  std::vector<uint16_t> probs, matched;
  for (uint32_t i = 0; i < n; i++)
  {
    const LzbUnit &u = units[i];
    LzbResult r;
    const bool l2 = u.codec == LZB_CODEC_LZMA2;
    uint32_t ns = l2 ? lzb_smem_probs(4, 0, 4) : lzb_smem_probs(u.lc, u.lp, u.pb);
    uint32_t ng = l2 ? lzb_gmem_probs(4, 0) : lzb_gmem_probs(u.lc, u.lp);
    if (ns * 2 < lzb::kPostSmemBytes) ns = lzb::kPostSmemBytes / 2;
    if (ng * 2 < lzb::kPostScratchBytes) ng = lzb::kPostScratchBytes / 2;
    probs.assign(ns, 0xDEAD);       /* poison: every model entry must be initialised by the decoder */
    matched.assign(ng, 0xDEAD);
    lzb::UnitDecoderT<1, false, kAhead> d;
    const bool stateful = (u.flags & LZB_UNIT_STATEFUL) != 0;
    d.begin(packed + u.src_off, u.src_len, out + u.dst_off, stateful ? u.filter_param : u.dst_cap, u.dict_size,
            probs.data(), matched.data());
    d.start_unit(u, stateful ? reinterpret_cast<LzbStreamState *>((uintptr_t)u.expect) : nullptr);
    for (;;)
    {
      if (d.st != lzb::S_RUN) d.advance();
      if (d.st == lzb::S_UNIT_END) break;
      d.step();
    }
    d.finish_unit(r);
    /* like the kernel: the lean form carries the per-lane x86 converter, the look-ahead form the tile scan */
    if (!stateful) lzb::post_process<1, false, kAhead>(u, r, out + u.dst_off, (uint8_t *)probs.data(), (uint8_t *)matched.data(), ns * 2);
    results[i] = r;
  }
  return 0;
}

/* ahead = 1: the look-ahead form of the decoder (what one-unit-per-warp launches with few units per SM run) */
extern "C" int emu_decode_units2(const uint8_t *packed, uint8_t *out, const LzbUnit *units, LzbResult *results,
                                 uint32_t n, int ahead)
{
  return ahead ? decode_units<true>(packed, out, units, results, n) : decode_units<false>(packed, out, units, results, n);
}

extern "C" int emu_decode_units(const uint8_t *packed, uint8_t *out, const LzbUnit *units, LzbResult *results,
                                uint32_t n)
{
  return decode_units<false>(packed, out, units, results, n);
}
