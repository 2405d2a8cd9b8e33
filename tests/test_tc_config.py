"""Host-side launch configuration of the tcgen05 kernels (csrc/tc_config.h), checked on CPU:
every channel shape MinkUNet14/34C/... can produce must get a pipeline (lag >= 1) that fits in
the 227 KB of shared memory a CTA may use and 512 TMEM columns."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <stdio.h>
#include <initializer_list>
#include "tc_config.h"
using namespace meb200::tc;
int main() {
  const unsigned chans[] = {16, 32, 48, 64, 96, 128, 160, 192, 256, 384, 512};
  const unsigned ncs[] = {16, 32, 48, 64, 96, 128, 192, 256};
  int bad = 0;
  for (unsigned cr : chans) for (unsigned cc : ncs) for (unsigned rows : {1u, 129u, 5000u, 800000u}) {
    FwdCfg f = fwd_config(cr, cc, rows);
    bool ok = f.bk != 0 && f.pipe.lag >= 1 && f.smem_bytes <= 227 * 1024 && 2 * f.R * cc <= kTmemCols
              && f.pipe.nA >= 3 && f.pipe.nB >= 2 && f.cps >= 1 && (cr / f.bk) % f.cps == 0
              && f.a_stage_bytes == f.cps * f.a_sub_bytes && f.a_stage_bytes <= 32 * 1024
              && 2 * f.pipe.nB * f.R >= f.pipe.nA + f.R - 1;  /* B-ring parity safety */
    if (!ok) { printf("FWD BAD cr=%u cc=%u rows=%u bk=%d lag=%u\n", cr, cc, rows, f.bk, f.pipe.lag); ++bad; }
    printf("fwd %u %u %u : bk=%d cps=%u R=%u depth=%u nA=%u nB=%u smem=%u\n", cr, cc, rows, f.bk, f.cps, f.R, f.pipe.lag, f.pipe.nA, f.pipe.nB, f.smem_bytes);
  }
  const unsigned cins[] = {16, 24, 32, 64, 96, 128, 192, 256, 384, 512};
  for (unsigned ci : cins) for (unsigned co : ncs) for (unsigned K : {1u, 8u, 27u, 125u}) for (unsigned rows : {1u, 300u, 800000u}) {
    WgCfg w = wgrad_config(ci, co, K, rows, 148);
    bool ok = w.mt_cta != 0 && w.pipe.lag >= 1 && w.smem_bytes <= 227 * 1024 &&
              w.G * w.mt_cta * co <= kTmemCols && w.G >= 1 && w.n_splits * w.rows_per_split >= rows
              && w.pipe.nA >= 2 && w.pipe.nB >= 2;
    /* one CTA per SM: never a third, nearly empty wave; k-groups cover K, sizes <= G, differ by <= 1 */
    if (ok && w.n_splits > 1) ok = w.n_splits * w.n_kgroups * w.n_mtgroups <= 2 * 148;
    if (ok) {
      unsigned lo = K, hi = 0;
      for (unsigned kg = 0; kg < w.n_kgroups; ++kg) {
        unsigned g = kgroup_begin(kg + 1, K, w.n_kgroups) - kgroup_begin(kg, K, w.n_kgroups);
        lo = g < lo ? g : lo; hi = g > hi ? g : hi;
      }
      ok = kgroup_begin(0, K, w.n_kgroups) == 0 && kgroup_begin(w.n_kgroups, K, w.n_kgroups) == K
           && hi <= w.G && hi - lo <= 1 && lo >= 1;
    }
    if (w.mt_cta == 0 && co == 256 && K == 1) ok = true;  // declared unsupported -> SIMT fallback
    if (!ok) { printf("WG BAD ci=%u co=%u K=%u rows=%u\n", ci, co, K, rows); ++bad; }
    printf("wg %u %u %u %u : mt=%u G=%u lag=%u nA=%u nB=%u splits=%u smem=%u\n", ci, co, K, rows, w.mt_cta, w.G, w.pipe.lag, w.pipe.nA, w.pipe.nB, w.n_splits, w.smem_bytes);
  }
  for (unsigned ci : cins) for (unsigned co : ncs) for (unsigned K : {1u, 8u, 27u, 125u}) for (unsigned rows : {1u, 300u, 800000u}) {
    Wg2Cfg w = wgrad2_config(ci, co, K, rows, 148);
    if (w.mt_cta == 0) continue;   // falls back to the 64-row kernel
    bool ok = w.pipe.nA >= 3 && w.pipe.nB >= 2 && w.smem_bytes <= 227 * 1024 &&
              w.G * w.mt_cta * co <= kTmemCols && w.n_splits * w.rows_per_split >= rows &&
              w.a_stage_bytes == w.blkA * kWg2BlkBytes && w.blkB * 64 >= co;
    if (!ok) { printf("WG2 BAD ci=%u co=%u K=%u rows=%u\n", ci, co, K, rows); ++bad; }
    printf("wg2 %u %u %u %u : mt=%u G=%u nA=%u nB=%u splits=%u smem=%u\n", ci, co, K, rows, w.mt_cta, w.G, w.pipe.nA, w.pipe.nB, w.n_splits, w.smem_bytes);
  }
  printf("bad=%d\n", bad);
  return bad != 0;
}
"""


def test_tc_launch_configs(tmp_path):
    src = tmp_path / "cfg.cpp"
    src.write_text(SRC)
    exe = tmp_path / "cfg"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I",
                    os.path.join(ROOT, "minkowskiengine_b200", "csrc"), str(src), "-o", str(exe)],
                   check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
    assert "bad=0" in r.stdout


if __name__ == "__main__":
    import pathlib
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        test_tc_launch_configs(pathlib.Path(d))
        print("ok")
