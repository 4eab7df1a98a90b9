"""Sweep-only FP32 rate vs occupancy (CTAs per SM) for the packed sweep, and the warp-specialisation experiment."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import rtw_b200 as R
with R.Renderer([0]) as r:
    peak, _ = r.measure_fp32_peak(0)
    print(f"FFMA peak {peak/1e12:.2f} T")
    names = {2: "packed coop1", 3: "packed coop2", 4: "packed coop4", 5: "mixed coop2: 8 sweep + 0 other warps",
             6: "mixed coop2: 8 sweep + 4 other", 7: "mixed coop2: 8 sweep + 8 other", 8: "mixed coop4: 8 sweep + 8 other"}
    for base in (3, 4, 5, 6, 7, 8):
        row = []
        for limit in (1, 2, 3, 0):
            rate, ms = r.measure_fp32_peak(base + 10 * limit)
            row.append(f"L{limit}: {rate/1e12:5.2f}")
        print(f"variant {base} ({names[base]}):", "  ".join(row), flush=True)
