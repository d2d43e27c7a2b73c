"""Summarise `ncu --page raw --csv` output of tools/prof_hbm.py into profiles/ (one line per profiled launch).

    python tools/ncu_summary.py gpurun_out/prof_hbm_raw.csv > profiles/r1_hbm_kernels_ncu.txt
"""
import csv, json, os, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
B, D, HW = 16, 512, 65536
n = B * D * HW
# algorithmic bytes per launch, by kernel name and element size
def algorithmic(name, esz, nth):
    if "tv_fwd" in name: return n * esz
    if "tv_bwd" in name: return (2 if nth == 0 else 3) * n * esz        # second launch: accumulate form (reads dX too)
    if "pool_fwd" in name or "pool_bwd" in name: return n * esz + B * HW * 8
    if "rownorm" in name: return n * esz + B * HW * 4 + (n * 2 if esz == 4 else 0)
    if "eval_hist" in name: return B * HW * 48
    if "sample_map" in name: return B * HW * (8 + 4 + 4 + 4)
    return None

def f(r, k):
    return float(r[col[k]].replace(",", ""))

print("# ncu --set full --clock-control none -k regex:pool_|tv_|rownorm|eval_hist|eval_fold|sample_|weight_sum : PROF_REPS=1 python tools/prof_hbm.py")
print("# B=16 of the B=64 batch, 256x256, D=512 (X bf16 1.07 GB / f32 2.15 GB); one launch per kernel; every ncu pass is cold-cache.")
print(f"# measured HBM peak (MEASURED_PEAKS.json): {peak} GB/s.  dram GB/s = (dram__bytes_read + dram__bytes_write) / gpu__time_duration;")
print("# alg GB/s = algorithmic bytes / gpu__time_duration (the figure bench_kernels.py reports with CUDA events, warm).")
print(f"{'kernel':28s} {'dtype':5s} {'grid':>6s} {'blk':>4s} {'regs':>4s} {'time_us':>8s} {'rd_GB':>7s} {'wr_GB':>7s} {'dram_GB/s':>9s} {'dram%pk':>7s} "
      f"{'alg_GB':>7s} {'alg_GB/s':>8s} {'alg/meas_peak':>13s} {'traffic/alg':>11s} {'warps%':>6s} {'issue%':>6s}")
seen = {}
for i, r in enumerate(data):
    name = r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("rc::", "")
    full = r[col["Kernel Name"]]
    esz = 4 if "<float>" in full else 2
    key = (name, esz)
    nth = seen.get(key, 0); seen[key] = nth + 1
    t_ms = f(r, "gpu__time_duration.sum")
    t_unit = units[col["gpu__time_duration.sum"]]
    t_us = t_ms * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}[t_unit]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    rd = f(r, "dram__bytes_read.sum") * scale[units[col["dram__bytes_read.sum"]]]
    wr = f(r, "dram__bytes_write.sum") * scale[units[col["dram__bytes_write.sum"]]]
    alg = algorithmic(name, esz, nth)
    gbs = (rd + wr) / (t_us * 1e-6) / 1e9
    line = (f"{name[:28]:28s} {('f32' if esz == 4 else 'bf16') if ('<float>' in full or 'bfloat16' in full or 'bf16' in name) else '-':5s} {int(f(r, 'launch__grid_size')):6d} {int(f(r, 'launch__block_size')):4d} "
            f"{int(f(r, 'launch__registers_per_thread')):4d} {t_us:8.1f} {rd / 1e9:7.3f} {wr / 1e9:7.3f} {gbs:9.0f} "
            f"{f(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):7.1f} ")
    if alg:
        ag = alg / (t_us * 1e-6) / 1e9
        line += f"{alg / 1e9:7.3f} {ag:8.0f} {ag / peak:13.3f} {(rd + wr) / alg:11.2f} "
    else:
        line += f"{'-':>7s} {'-':>8s} {'-':>13s} {'-':>11s} "
    line += f"{f(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f}"
    print(line)
