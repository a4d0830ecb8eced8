"""Turns an .ncu-rep (ncu --set full) into the short text summary committed under profiles/."""
import csv, io, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "sm__inst_executed_pipe_tmem", "smsp__inst_executed_pipe_xu"]
lines = [f"# ncu --set full summary of {rep}"]
for h, u, v in zip(hdr, units, vals):
    if h in keys or any(h.startswith(k) for k in keys[3:]):
        if h in keys or h.endswith(".sum") or h.endswith(".pct") or "pct_of_peak_sustained_elapsed" in h:
            lines.append(f"{h:95s} {v:>20s} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
h2 = srows[1]; data = srows[2:]; ci = {h: i for i, h in enumerate(h2)}
tot = sum(int(r[ci["# Samples"]] or 0) for r in data)
stalls = [h for h in h2 if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[ci[h]] or 0) for r in data) for h in stalls}
lines.append(f"\n# warp stall sampling, {tot} samples: " + ", ".join(f"{k[6:]}={v * 100 // max(tot, 1)}%" for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
lines.append("# top instructions by samples")
for r in sorted(data, key=lambda r: -int(r[ci["# Samples"]] or 0))[:14]:
    lines.append(f"{int(r[ci['# Samples']]):7d}  {r[1].strip()[:90]}")
sass = " ".join(r[1] for r in data)
lines.append("\n# Blackwell instructions present in the SASS: " + ", ".join(m for m in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "SYNCS", "HMMA") if m in sass))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:40]))
