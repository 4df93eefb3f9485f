"""Turns the raw ncu outputs in gpurun_out/ into the small text/CSV/JSON summaries kept under profiles/."""
import collections
import csv
import json
import subprocess
import sys


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        agg.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write("kernel,launches,avg_us,total_ms,share_pct\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"\"{k}\",{len(v)},{sum(v) / len(v):.1f},{sum(v) / 1e3:.3f},{100 * sum(v) / tot:.2f}\n")
    return agg


def full(rep, out_txt, out_json):
    txt = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True, stderr=subprocess.DEVNULL)
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg"]
    summary = {}
    with open(out_txt, "w") as f:
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")]
            f.write(name + "\n")
            d = {}
            for w in want:
                if w in hdr:
                    i = hdr.index(w)
                    f.write(f"   {w:70s} {r[i]:>20s} {units[i]}\n")
                    d[w] = (r[i], units[i])
            st = [(float(r[i]), h) for i, h in enumerate(hdr)
                  if "smsp__average_warps_issue_stalled" in h and "_not_issued" not in h and r[i] not in ("", "n/a")]
            tot = sum(v for v, _ in st)
            for v, h in sorted(st, reverse=True)[:10]:
                f.write(f"   stall {h[34:]:55s} {v:8.3f} ({100 * v / tot:4.1f}%)\n")
            summary.setdefault(name, []).append(d)
    json.dump(summary, open(out_json, "w"), indent=1)
    return summary


if __name__ == "__main__":
    tag = sys.argv[1]
    launches(f"gpurun_out/launches_{tag}.csv", f"profiles/{tag}_launch_list.csv")
    rep = sys.argv[2]
    s = full(rep, f"profiles/{tag}_flux_kernel_ncu_full.txt", f"profiles/{tag}_flux_kernel_ncu_full.json")
    # per-launch DRAM traffic of the dominant kernel (average of the x and y instantiations)
    tr = []
    for name, lst in s.items():
        for d in lst:
            def mb(x):
                v, u = x
                v = float(v.replace(",", ""))
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
            tr.append(mb(d["dram__bytes_read.sum"]) + mb(d["dram__bytes_write.sum"]))
    json.dump({"dram_bytes_per_launch": sum(tr) / len(tr), "per_capture": tr,
               "source": f"profiles/{tag}_flux_kernel_ncu_full.txt (ncu --set full, bench.py workload)"},
              open("profiles/flux_kernel_traffic.json", "w"), indent=1)
    print(open(f"profiles/{tag}_launch_list.csv").read())
    print(open(f"profiles/{tag}_flux_kernel_ncu_full.txt").read())
