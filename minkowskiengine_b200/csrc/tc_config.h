// Host-side launch configuration of the tcgen05 kernels (plain C++, no CUDA): tile counts,
// TMEM accumulator packing and shared-memory pipeline depths.  Kept separate so it can be
// unit-tested on a machine without a GPU (tests/test_tc_config.py).
#pragma once
#include <stdint.h>

namespace meb200 {
namespace tc {

constexpr int kTileM = 128;
constexpr int kMaxStages = 32;
constexpr uint32_t kMaxLag = 16;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kSmemBudget = 224 * 1024;            // of the 227 KB a CTA may opt in to
constexpr uint32_t kBarBytes = (4 * kMaxStages + 4) * 8 + 16;
constexpr int kWgRows = 64;                             // wgrad: reduction rows per stage
constexpr uint32_t kBlkBytes = kWgRows * 128;           // wgrad: one 64-channel block of a stage

inline uint32_t cdiv_u(uint64_t a, uint64_t b) { return (uint32_t)((a + b - 1) / b); }

// Producers publish a stage with cp.async.mbarrier.arrive (the barrier fires when the copies
// land), so every stage of the ring may be in flight.  The B ring must hold the B slice of every
// A stage that can be in flight (`share` A stages use one B slice) plus one being filled —
// otherwise a producer would wait for a B stage whose consumers are still behind it in the ring.
struct PipeCfg { uint32_t lag, nA, nB; };   // lag = nA - 1 (kept for reporting)
inline PipeCfg pick_pipeline(uint32_t a_bytes, uint32_t b_bytes, uint32_t share, uint32_t budget) {
  PipeCfg best{0, 0, 0};
  for (uint32_t nA = kMaxStages; nA >= 3; --nA) {
    uint32_t nB = (nA + share - 1) / share + 1;
    if (nB > (uint32_t)kMaxStages) continue;
    if ((uint64_t)nA * a_bytes + (uint64_t)nB * b_bytes <= budget) { best = {nA - 1, nA, nB}; break; }
  }
  return best;
}

// ---- forward / dgrad: out[r, 0:c_cols] over n_rows rows, reduction over c_red channels ----
struct FwdCfg {
  int bk;                 // channel-chunk width (64/32/16 -> 128B/64B/32B swizzle), 0 = unsupported
  uint32_t R;             // row tiles sharing one B slice (1, 2 or 4)
  uint32_t n_super;       // super tiles of R*128 rows
  uint32_t a_stage_bytes, b_stage_bytes;
  PipeCfg pipe;
  uint32_t smem_bytes;
};
inline FwdCfg fwd_config(uint32_t c_red, uint32_t c_cols, uint32_t n_rows) {
  FwdCfg c{};
  uint32_t R = kTmemCols / (2 * c_cols);                // accumulator set is double buffered
  R = R >= 4 ? 4 : (R >= 2 ? 2 : 1);
  uint32_t tiles = cdiv_u(n_rows, kTileM);
  while (R > 1 && R > tiles) R >>= 1;
  c.R = R;
  c.n_super = cdiv_u(tiles, R);
  const int cands[3] = {64, 32, 16};
  for (int cand : cands) {
    if (c_red % cand != 0) continue;
    uint32_t a_bytes = kTileM * cand * 2;
    uint32_t b_bytes = ((c_cols * cand * 2 + 1023) / 1024) * 1024;
    PipeCfg p = pick_pipeline(a_bytes, b_bytes, R, kSmemBudget - kBarBytes - 1024);
    if (p.lag > c.pipe.lag) { c.pipe = p; c.bk = cand; c.a_stage_bytes = a_bytes; c.b_stage_bytes = b_bytes; }
    if (p.lag >= 4) break;
  }
  c.smem_bytes = 1024 + c.pipe.nA * c.a_stage_bytes + c.pipe.nB * c.b_stage_bytes + kBarBytes;
  return c;
}

// ---- wgrad: dW[K, c_in, c_out] reduced over n_out rows --------------------------------------
struct WgCfg {
  uint32_t mt_cta;        // 128-channel m-tiles of c_in per CTA (1 or 2); 0 = unsupported
  uint32_t n_mtgroups, G, n_kgroups, blkA, blkB;
  uint32_t a_stage_bytes, b_stage_bytes;
  PipeCfg pipe;
  uint32_t rows_per_split, n_splits;
  uint32_t smem_bytes;
};
inline WgCfg wgrad_config(uint32_t c_in, uint32_t c_out, uint32_t K, uint32_t n_out, uint32_t n_sms) {
  WgCfg best{};
  const uint32_t mt_total = cdiv_u(c_in, 128);
  uint32_t mt_max = mt_total < kTmemCols / c_out ? mt_total : kTmemCols / c_out;
  if (mt_max > 2) mt_max = 2;
  for (uint32_t mt = mt_max; mt >= 1; --mt) {
    WgCfg c{};
    c.mt_cta = mt;
    c.n_mtgroups = cdiv_u(mt_total, mt);
    c.G = kTmemCols / (mt * c_out);
    if (c.G > K) c.G = K;
    if (c.G > 8) c.G = 8;
    c.n_kgroups = cdiv_u(K, c.G);
    c.blkA = mt * 2;
    c.blkB = cdiv_u(c_out, 64);
    c.a_stage_bytes = c.blkA * kBlkBytes;
    c.b_stage_bytes = c.blkB * kBlkBytes;
    c.pipe = pick_pipeline(c.a_stage_bytes, c.b_stage_bytes, c.G, kSmemBudget - kBarBytes - 1024);
    if (c.pipe.lag > best.pipe.lag) best = c;
    if (c.pipe.lag >= 3) break;
  }
  if (best.pipe.lag == 0) { best.mt_cta = 0; return best; }
  // row slices: ~2 waves of CTAs, each slice a multiple of the stage height
  uint32_t base = best.n_kgroups * best.n_mtgroups;
  uint32_t want = cdiv_u(2ull * n_sms, base);
  uint32_t max_splits = cdiv_u(n_out, 4 * kWgRows);
  uint32_t splits = want < 1 ? 1 : (want > max_splits ? max_splits : want);
  best.rows_per_split = cdiv_u(cdiv_u(n_out, splits), kWgRows) * kWgRows;
  best.n_splits = cdiv_u(n_out, best.rows_per_split);
  best.smem_bytes = 1024 + best.pipe.nA * best.a_stage_bytes + best.pipe.nB * best.b_stage_bytes + kBarBytes;
  return best;
}

}  // namespace tc
}  // namespace meb200
