"""Summarise an .ncu-rep (ncu --set full) per kernel launch: duration, DRAM bytes, pipe / issue utilisation, registers, top stalls.
   python profiles/ncu_summary.py X.ncu-rep > profiles/NAME.txt"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("lts__t_bytes.sum", "L2 bytes"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "registers / thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
        ("smsp__inst_executed.avg", "warp instructions / SMSP"), ("sm__cycles_active.avg", "SM active cycles")]
STALL = "smsp__average_warps_issue_stalled_"
for r in rows[2:]:
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    print(f"== {d.get('Kernel Name', '?')[:110]}   (launch id {d.get('ID')})")
    for k, name in KEYS:
        if k in d and d[k] != "":
            print(f"   {name:28s} {d[k]} {u.get(k, '')}")
    stalls = sorted(((float(v), k[len(STALL):].replace('_per_issue_active.ratio', '')) for k, v in d.items()
                     if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)[:5]
    print("   top stall reasons (warps per issue):", ", ".join(f"{n} {v:.2f}" for v, n in stalls))
    print()
