#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text/JSON summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches_bench.txt
  python tools/summarize_ncu.py full gpurun_out/prof.ncu-rep profiles/r01_bin_scan_full.txt [profiles/bin_scan_traffic.json units]
"""
import csv
import io
import json
import subprocess
import sys


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = {}
    for r in rows:
        name = r["Kernel Name"].split("(")[0]
        try:
            ns = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        if r.get("Metric Unit", "ns") in ("us", "usecond"):
            ns *= 1e3
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; {len(rows)} launches, {tot/1e3:.1f} us total\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':70s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{k[:70]:70s} {v[0]:8d} {v[1]/1e3:12.1f} {v[1]/1e3/v[0]:10.2f} {100*v[1]/tot:6.1f}%\n")
    print(open(dst).read())


def full(src, dst, traffic_json=None, units=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, unit = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
            "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum"]
    out = []
    traffic = None
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, unit))
        out.append(f"== {d.get('Kernel Name', '?')[:100]}")
        for k in want:
            if k in d:
                out.append(f"   {k:95s} {d[k]:>18s} {u.get(k, '')}")
        try:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd = float(d["dram__bytes_read.sum"].replace(",", "")) * scale[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"].replace(",", "")) * scale[u["dram__bytes_write.sum"]]
            traffic = rd + wr
            out.append(f"   dram bytes read+write per launch = {traffic:.0f}")
        except Exception:
            pass
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))
    if traffic_json and traffic is not None:
        json.dump({"dram_bytes_per_launch": traffic, "units_in_profiled_launch": int(units) if units else None,
                   "source": src}, open(traffic_json, "w"))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(*sys.argv[2:])
