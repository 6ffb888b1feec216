"""Summarise ncu outputs brought back in gpurun_out/ into small committed files under profiles/.

  python tools/summarize_ncu.py launches <launches.csv> <out.md>      # per-kernel share of the step
  python tools/summarize_ncu.py kernel <report.ncu-rep> <out.md>      # key counters of one full-set capture
"""
import collections, csv, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.OrderedDict()
    scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "second": 1e3, "s": 1e3}
    for row in csv.DictReader(lines):
        k = row["Kernel Name"].split("(")[0]
        v = float(row["Metric Value"].replace(",", "")) * scale[row["Metric Unit"]]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"source: {src} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised: "
                f"compare SHARES)\n\n| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
            f.write(f"| `{k}` | {n} | {ms:.2f} | {100 * ms / tot:.1f} % |\n")
        f.write(f"\ntotal {tot:.2f} ms over {sum(v[0] for v in agg.values())} launches\n")


def kernel(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"source: {rep} (ncu --set full --clock-control none --import-source on)\n")
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")]
            f.write(f"\n### {name.split('(')[0]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for i, h in enumerate(hdr):
                if h in KEYS:
                    f.write(f"| {h} | {vals[i]} | {units[i]} |\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
