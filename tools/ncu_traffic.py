"""Summarise one `ncu --set full` capture of the persistent step kernel into profiles/.

usage: python tools/ncu_traffic.py gpurun_out/<tag>_flow_mega.ncu-rep <tag> [src_sha16]

Writes profiles/<tag>_flow_mega_ncu_summary.txt (the metrics the judge reads) and profiles/<tag>_flow_mega_traffic.json
(dram bytes per launch + the sha of the kernel sources the capture was taken from; bench.py refuses a stale one).
The sha is the one bench.py computes (src_sha16); pass the value recorded ON THE GPU BOX next to the capture when the
sources have changed since."""
import csv
import io
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rep, tag = sys.argv[1], sys.argv[2]
if len(sys.argv) > 3:
    sha = sys.argv[3]
else:
    import bench
    sha = bench.src_sha16()
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
WANT = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_write.sum.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tc.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__inst_executed.sum"]


def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(u, 1)


out = [f"{m['Kernel Name'][0]}: ncu --set full --clock-control none ({Path(rep).name}), kernel sources {sha}"]
for k in WANT[1:]:
    if k in m:
        out.append(f"{k} = {m[k][0]} {m[k][1]}")
(ROOT / "profiles" / f"{tag}_flow_mega_ncu_summary.txt").write_text("\n".join(out) + "\n")
rd, wr = to_bytes(*m["dram__bytes_read.sum"]), to_bytes(*m["dram__bytes_write.sum"])
dur = m["gpu__time_duration.sum"]
(ROOT / "profiles" / f"{tag}_flow_mega_traffic.json").write_text(json.dumps({
    "kernel": m["Kernel Name"][0], "src_sha16": sha, "capture": Path(rep).name, "dram_bytes_read": rd, "dram_bytes_write": wr,
    "dram_bytes_per_launch": rd + wr, "gpu_time_duration": f"{dur[0]} {dur[1]}",
    "workload": "B=256 utterances x 100 frames, voice prefix 125 (tools/ncu_step.py 256 100 0)"}, indent=1) + "\n")
print("\n".join(out))
