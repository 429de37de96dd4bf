"""Copy the evidence produced by tools/gpu_evidence.sh (and the multi-GPU scripts) from gpurun_out/
(scratch) into profiles/ (tracked), and condense the `ncu --set full` captures into
profiles/ncu_summary.json (what bench.py reads for `roofline.traffic`).

    python tools/collect_profiles.py [tag]        # tag defaults to r02 (round 2)
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

COPY = {
    "t_gpu_all.log": "{tag}_pytest_gpu.log",
    "smoke.log": "{tag}_smoke.log",
    "bench_headline.json": "{tag}_bench_headline_38p6M_1gpu.json",
    "bench_reference.json": "{tag}_bench_reference_cpu.json",
    "bench_4p8M.json": "{tag}_bench_4p8M_rows_1gpu.json",
    "launches_headline.csv": "{tag}_launches_headline_38p6M.csv",
    "launches_4p8M.csv": "{tag}_launches_4p8M_rows.csv",
    "sweep_8p8M_k100.json": "{tag}_sweep_8p8M_k100.json",
    "sweep_8p8M_k1000.json": "{tag}_sweep_8p8M_k1000.json",
    "bench_c2.json": "{tag}_bench_c2_8p8M_k1000_1gpu.json",
}
SCALE = {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
TIME = {"s": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6}
KEEP = ("gpu__time_duration", "dram__bytes", "dram__throughput", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor",
        "sm__cycles_elapsed", "smsp__average_warps_issue_stalled", "lts__t_sector_hit_rate", "lts__throughput",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block",
        "launch__cluster", "sm__throughput", "l1tex__m_xbar2l1tex_read_bytes", "smsp__inst_executed.sum")


def ncu_raw(rep):
    res = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    rows = list(csv.reader(res.stdout.splitlines()))
    rows = [r for r in rows if len(r) > 10]
    return rows[0], rows[1], rows[2:]


def summarise(rep, rows_streamed, label, tag):
    hdr, units, vals = ncu_raw(rep)
    v = vals[0]
    d, u = dict(zip(hdr, v)), dict(zip(hdr, units))
    f = lambda k: float(d[k].replace(",", ""))
    rd = f("dram__bytes_read.sum") * SCALE[u["dram__bytes_read.sum"]]
    wr = f("dram__bytes_write.sum") * SCALE[u["dram__bytes_write.sum"]]
    ms = f("gpu__time_duration.sum") * TIME[u["gpu__time_duration.sum"]]
    slim = os.path.join(PROF, f"{tag}_ncu_full_{label}.csv")
    with open(slim, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["metric", "unit", "value"])
        for h, un, val in zip(hdr, units, v):
            if h in ("Kernel Name", "Grid Size", "Block Size") or any(h.startswith(k) or ("." + k) in h for k in KEEP):
                w.writerow([h, un, val])
    return {
        "source": os.path.relpath(slim, ROOT), "captured_launch": label, "captured_rows": rows_streamed,
        "kernel": d.get("Kernel Name", "")[:60], "duration_ms": ms, "dram_bytes_read": rd, "dram_bytes_write": wr,
        "dram_bytes_per_row": (rd + wr) / rows_streamed, "algorithmic_bytes_per_row_streamed": 1536,
        "achieved_read_tbs": rd / ms / 1e9,
        "tensor_pipe_active_pct": f("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "sm_clock_ghz": f("sm__cycles_elapsed.avg.per_second"),
        "registers_per_thread": f("launch__registers_per_thread"),
    }


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    os.makedirs(PROF, exist_ok=True)
    for src, dst in COPY.items():
        p = os.path.join(OUT, src)
        if os.path.exists(p) and os.path.getsize(p) > 0:
            shutil.copyfile(p, os.path.join(PROF, dst.format(tag=tag)))
            print("copied", src, "->", dst.format(tag=tag))
    path = os.path.join(PROF, "ncu_summary.json")
    summary = {}
    if os.path.exists(path):
        with open(path) as fh:
            summary = json.load(fh)          # earlier rounds' captures stay (the TS kernel's are round 1's)
    fresh = {}
    for rep, rows, label in (("prof_qs_headline.ncu-rep", 38636520, "umma_qs_38p6M_rows"),
                             ("prof_qs_4p8M.ncu-rep", 4829565, "umma_qs_4p8M_rows")):
        p = os.path.join(OUT, rep)
        if os.path.exists(p):
            fresh[label] = summarise(p, rows, label, tag)
            print(label, json.dumps(fresh[label]))
    summary.update(fresh)
    # bench.py reads <kernel name>.dram_bytes_per_row (the headline-size capture of each scoring kernel)
    if "umma_qs_38p6M_rows" in summary or "umma_qs_4p8M_rows" in summary:
        summary["umma_qs_score_select_kernel"] = summary.get("umma_qs_38p6M_rows") or summary["umma_qs_4p8M_rows"]
    if "umma_ts_38p6M_rows" in summary:
        summary["umma_score_select_kernel"] = summary["umma_ts_38p6M_rows"]
    with open(path, "w") as fh:
        json.dump(summary, fh, indent=1)


if __name__ == "__main__":
    main()
