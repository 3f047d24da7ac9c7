"""Development tool: turns the files a `tools/gpu_session.sh <tag> final` run brought back in gpurun_out/ into the
tracked summaries under profiles/ (bench line, ncu launch list, per-kernel ncu summary, DRAM traffic per ray of the
extend kernel's first launch). usage: python tools/make_profiles.py <tag> [round-prefix, default r02]"""
import csv
import io
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rnd = sys.argv[2] if len(sys.argv) > 2 else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def run(*cmd):
    return subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT).stdout


for gz in (f"{tag}_kernels.ncu-rep.gz", f"{tag}_extend_first.ncu-rep.gz"):
    if os.path.exists(os.path.join(G, gz)):
        subprocess.run(["gunzip", "-f", os.path.join(G, gz)], check=True)
bench = os.path.join(G, f"{tag}_bench.json")
if os.path.exists(bench):
    line = open(bench).read().strip().splitlines()[-1]
    json.loads(line)
    open(os.path.join(P, f"{rnd}_bench_n1.json"), "w").write(line + "\n")
for cfg in ("c1", "c2", "c5"):
    f = os.path.join(G, f"{tag}_bench_{cfg}.json")
    if os.path.exists(f):
        line = open(f).read().strip().splitlines()[-1]
        open(os.path.join(P, f"{rnd}_bench_{cfg}.json"), "w").write(line + "\n")
launches = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(launches):
    open(os.path.join(P, f"{rnd}_c3_launch_list.txt"), "w").write(
        run(sys.executable, "tools/ncu_summary.py", "--launches", launches))
rep = os.path.join(G, f"{tag}_kernels.ncu-rep")
if os.path.exists(rep):
    open(os.path.join(P, f"{rnd}_c3_kernels_full.txt"), "w").write(run(sys.executable, "tools/ncu_summary.py", rep))
first = os.path.join(G, f"{tag}_extend_first.ncu-rep")
if os.path.exists(first):
    raw = run("ncu", "-i", first, "--page", "raw", "--csv")
    rows = list(csv.reader(io.StringIO(raw)))
    idx = {h: i for i, h in enumerate(rows[0])}
    r = rows[2]
    rd, wr = float(r[idx["dram__bytes_read.sum"]]), float(r[idx["dram__bytes_write.sum"]])
    unit = rows[1][idx["dram__bytes_read.sum"]]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    rays = 1920 * 1080
    out = {"source": f"profiles/{rnd}_c3_extend_first_launch_full.txt (ncu --set full of the FIRST k_extend launch of a C3 "
                     f"frame: {rays:,} primary rays)",
           "dram_bytes_read": rd * scale, "dram_bytes_write": wr * scale, "rays": rays,
           "dram_bytes_per_ray": (rd + wr) * scale / rays}
    json.dump(out, open(os.path.join(P, "extend_traffic.json"), "w"), indent=1)
    open(os.path.join(P, f"{rnd}_c3_extend_first_launch_full.txt"), "w").write(
        run(sys.executable, "tools/ncu_summary.py", first))
for n in (1, 2, 4, 8):
    f = os.path.join(G, f"{tag}_scale_n{n}.json")
    if os.path.exists(f):
        shutil.copyfile(f, os.path.join(P, f"{rnd}_scale_n{n}.json"))
print(sorted(os.listdir(P)))
