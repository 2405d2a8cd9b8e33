"""Text summary of one `ncu --set full --import-source on` capture (a .ncu-rep brought back in
gpurun_out/): the raw metrics the roofline discussion uses plus the SASS instructions with the most
warp-stall samples.  Run where `ncu` is installed (no GPU needed to read a report):

    python profiles/ncu_summary.py gpurun_out/r2_runN/ncu_k_conv_ts.ncu-rep "header line" > profiles/r2_ncu_....txt
"""
import csv
import io
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_sectors_srcunit_tex_op_read.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True,
                         text=True, check=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    if len(sys.argv) > 2:
        print(sys.argv[2])
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
    print(f"kernel: {name}\n")
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print(f"{m} [{units[i]}] = {vals[i]}")
    src = page(rep, "source")
    h, data = src[1], src[2:]
    ia, isrc, isamp, iex = (h.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed"))
    stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot_s = sum(int(r[isamp] or 0) for r in data) or 1
    tot_e = sum(int(r[iex] or 0) for r in data) or 1
    print(f"\nWarp-stall sampling by SASS instruction ({tot_s} samples, {tot_e} instructions executed):")
    for r in sorted(data, key=lambda r: -int(r[isamp] or 0))[:14]:
        why = sorted(((int(r[i] or 0), h[i][6:]) for i in stall), reverse=True)[:2]
        why = ", ".join(f"{n} {w}" for c, n in ((c, n) for c, n in why) for w in [c] if c > 0)
        print(f"  {int(r[isamp]) / tot_s * 100:5.1f} %  executed {int(r[iex] or 0):>10}  "
              f"{r[isrc].strip()[:70]:70s}  [{why}]")


if __name__ == "__main__":
    main()
