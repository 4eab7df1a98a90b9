"""Turns an .ncu-rep (ncu --set full --import-source on) into the two text summaries kept under profiles/:
  <out>_full_metrics.csv     : metric,unit,value of the first profiled kernel (ncu --page raw)
  <out>_source_hotspots.txt  : executed warp instructions grouped into straight-line regions of equal execution
                               count (ncu --page source): share of instructions / samples, active threads, opcode mix
Usage (here, no GPU needed): python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r01_name"""
import csv
import io
import subprocess
import sys
from collections import Counter


import re

# the metrics the summaries under profiles/ quote (the full set stays in the .ncu-rep under gpurun_out/)
KEEP = re.compile(r"^(Kernel Name|Block Size|Grid Size|launch__(registers_per_thread|shared_mem_per_block_(dynamic|static)|"
                  r"occupancy_limit_\w+|waves_per_multiprocessor)|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|"
                  r"dram__throughput\.avg\.pct_of_peak_sustained_elapsed|lts__t_bytes\.sum|lts__t_sector_hit_rate\.pct|"
                  r"l1tex__m_xbar2l1tex_read_bytes\.sum|l1tex__data_bank_conflicts_pipe_lsu\.sum|"
                  r"l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"sm__pipe_(fma|fmaheavy|fmalite|alu|fp64|xu)_cycles_active\.avg\.pct_of_peak_sustained_active|"
                  r"sm__inst_executed_pipe_(fma|alu|lsu|xu|fp64|uniform|cbu|adu)\.sum\.pct_of_peak_sustained_active|"
                  r"sm__inst_executed_pipe_\w+\.sum|smsp__inst_executed\.sum|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__warps_eligible\.avg\.per_cycle_active|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__thread_inst_executed_per_inst_executed\.ratio|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio|sm__cycles_elapsed\.avg|"
                  r"sm__cycles_active\.avg|smsp__cycles_active\.avg)$")


def ncu_csv(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = ncu_csv(rep, "raw")
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out + "_full_metrics.csv", "w") as f:
        f.write("metric,unit,value\n")
        for h, u, v in zip(hdr, units, vals):
            if not KEEP.search(h):
                continue
            f.write(f"{h},{u},{v}\n".replace("rtw::<", "<"))
    rows = ncu_csv(rep, "source")
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else ""
    hdr, data = rows[1], rows[2:]
    iA, iS, iE, iT, iN = (hdr.index(k) for k in ("Address", "Source", "Instructions Executed", "Avg. Threads Executed", "# Samples"))
    base = int(data[0][iA], 16)
    tot = sum(int(r[iE]) for r in data)
    tot_s = sum(int(r[iN]) for r in data)
    groups, cur = [], None
    for r in data:
        off, e, t, sm = int(r[iA], 16) - base, int(r[iE]), float(r[iT]), int(r[iN])
        w = r[iS].split()
        op = (w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "?"))
        if cur and cur["e"] > 0 and abs(e - cur["e"]) <= 0.03 * cur["e"]:
            cur["n"] += 1; cur["sum"] += e; cur["tsum"] += t * e; cur["ops"].append(op); cur["end"] = off; cur["s"] += sm
        else:
            cur = {"start": off, "end": off, "e": e, "n": 1, "sum": e, "tsum": t * e, "ops": [op], "s": sm}
            groups.append(cur)
    pipes = Counter()
    for r in data:
        w = r[iS].split()
        op = (w[1] if w and w[0].startswith("@") and len(w) > 1 else (w[0] if w else "?")).split(".")[0]
        pipes[op] += int(r[iE])
    with open(out + "_source_hotspots.txt", "w") as f:
        f.write(f"kernel: {kernel}\n")
        f.write(f"executed warp instructions: {tot}   stall samples: {tot_s}\n\n")
        f.write("executed warp instructions by opcode (top 24):\n")
        for op, c in pipes.most_common(24):
            f.write(f"  {op:14s} {c:14d}  {100.0 * c / tot:6.2f} %\n")
        f.write("\nstraight-line regions (SASS offsets) with >= 0.1 % of the executed instructions:\n")
        for g in groups:
            if g["sum"] / tot < 0.001:
                continue
            mix = ", ".join(f"{o}x{c}" for o, c in Counter(g["ops"]).most_common(6))
            f.write(f"{g['start']:5x}-{g['end']:5x} n={g['n']:4d} exec/instr={g['e'] / 1e6:9.3f}M inst={100 * g['sum'] / tot:5.2f}% "
                    f"samples={100 * g['s'] / max(tot_s, 1):5.2f}% thr/inst={g['tsum'] / max(g['sum'], 1):5.1f}  {mix}\n")


if __name__ == "__main__":
    main()
