"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (tracked).

    python tools/ncu_summary.py launches gpurun_out/launches_c3.csv profiles/r1_launches_c3.txt
    python tools/ncu_summary.py kernel   gpurun_out/attn_c3.ncu-rep profiles/r1_attention_c3.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "smsp__inst_executed.sum"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
        a = agg[row["Kernel Name"][:110]]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (per-launch times are cold-cache and serialised: compare SHARES)\n")
        f.write(f"# source: {src}; launches {sum(v[0] for v in agg.values())}; total {tot:.2f} ms\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:10.3f} ms {100 * v[1] / tot:6.2f}%  n={v[0]:5d}  avg {v[1] / v[0]:9.4f} ms  {k}\n")
    print(open(dst).read()[:3000])


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none; source: {src}\n")
        for vals in rows[2:]:
            d = dict(zip(hdr, vals))
            f.write(f"## {d.get('Kernel Name', '')[:100]}  grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}\n")
            for h, u, v in zip(hdr, units, vals):
                if any(h.endswith(k) or h == k for k in KEYS):
                    f.write(f"{h:95s} {u:16s} {v}\n")
    print(open(dst).read()[:4000])


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3])
