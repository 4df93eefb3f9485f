"""Turns the raw ncu outputs in gpurun_out/ into the small text/CSV/JSON summaries kept under profiles/.

    python scripts/summarize_profiles.py r2 gpurun_out/prof_r2f.ncu-rep [steps_in_the_capture]

launches_<tag>.csv   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv of the bench
                     command: per kernel launches, average duration, share of the captured time, DRAM bytes per launch.
<rep>                one ncu --set full capture of the dominant kernel (x and y sweep)."""
import collections
import csv
import json
import subprocess
import sys

UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3,
        "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(path, out, steps):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ii, ki, mi, vi, ui = (hdr.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Value", "Metric Unit"))
    per = collections.OrderedDict()               # launch id -> dict
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
        except ValueError:
            continue
        d = per.setdefault(r[ii], {"kernel": r[ki]})
        d[r[mi]] = v
    agg = collections.OrderedDict()
    for d in per.values():
        a = agg.setdefault(d["kernel"], {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1
        a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0)
        a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["us"] for a in agg.values())
    tot_b = sum(a["rd"] + a["wr"] for a in agg.values())
    with open(out, "w") as f:
        f.write("kernel,launches,avg_us,total_ms,share_pct,dram_read_MB_per_launch,dram_write_MB_per_launch,dram_GB_total\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
            f.write(f"\"{k}\",{a['n']},{a['us'] / a['n']:.1f},{a['us'] / 1e3:.3f},{100 * a['us'] / tot:.2f},"
                    f"{a['rd'] / a['n'] / 1e6:.1f},{a['wr'] / a['n'] / 1e6:.1f},{(a['rd'] + a['wr']) / 1e9:.3f}\n")
        f.write(f"\"TOTAL ({steps} steps incl. warm-up, upload excluded from nothing: every launch of the process)\",,,"
                f"{tot / 1e3:.3f},100.00,,,{tot_b / 1e9:.3f}\n")
    return agg, tot_b


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
    rep = sys.argv[2]
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    agg, tot_b = launches(f"gpurun_out/launches_{tag}.csv", f"profiles/{tag}_launch_list.csv", steps)
    if tag.endswith("3d"):           # 3-D extension (scripts/r2_profile3d.sh): summaries only, bench.py does not read them
        full(rep, f"profiles/{tag}_sweep_kernels_ncu_full.txt", f"profiles/{tag}_sweep_kernels_ncu_full.json")
        print(open(f"profiles/{tag}_launch_list.csv").read())
        print(open(f"profiles/{tag}_sweep_kernels_ncu_full.txt").read())
        sys.exit(0)
    s = full(rep, f"profiles/{tag}_flux_kernel_ncu_full.txt", f"profiles/{tag}_flux_kernel_ncu_full.json")
    tr, pipe = [], []
    for name, lst in s.items():
        for d in lst:
            def val(x):
                v, u = x
                return float(v.replace(",", "")) * UNIT.get(u, 1.0)
            tr.append(val(d["dram__bytes_read.sum"]) + val(d["dram__bytes_write.sum"]))
            pipe.append(float(d["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"][0]))
    # DRAM bytes of one step: every kernel launched per step (all but the start-up kernels), from the launch list
    flux = [a for k, a in agg.items() if "flux" in k]
    nsteps_captured = sum(a["n"] for a in flux) / 8.0
    step_b = sum((a["rd"] + a["wr"]) for k, a in agg.items() if "center2face" not in k) / max(nsteps_captured, 1)
    json.dump({"dram_bytes_per_launch": sum(tr) / len(tr), "per_capture": tr, "fp64_pipe_pct": pipe,
               "step_dram_bytes": step_b, "steps_in_launch_list": nsteps_captured,
               "source": f"profiles/{tag}_flux_kernel_ncu_full.txt (ncu --set full) and profiles/{tag}_launch_list.csv, bench.py workload"},
              open("profiles/flux_kernel_traffic.json", "w"), indent=1)
    print(open(f"profiles/{tag}_launch_list.csv").read())
    print(open(f"profiles/{tag}_flux_kernel_ncu_full.txt").read())
