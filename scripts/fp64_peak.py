"""FP64 FMA peak of this device (the roofline denominator of the shell stiffness kernels) with the clocks it was
measured at: fsgpu_measure_peaks (DFMA micro-kernel) repeated, nvidia-smi sampled during the runs."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fsb200
from bench import ClockSampler

ctx = fsb200.Context()
ctx.measure_peaks()
s = ClockSampler(0)
s.start()
vals = [ctx.measure_peaks() for _ in range(20)]
clocks = s.stop()
fp = sorted(v[0] for v in vals)
cp = sorted(v[1] for v in vals)
print(json.dumps({"fp64_tflops_best": fp[-1], "fp64_tflops_median": fp[len(fp) // 2], "copy_gbs_best": cp[-1], "copy_gbs_median": cp[len(cp) // 2],
                  "repeats": len(vals), "clocks": clocks, "how": "fsgpu_measure_peaks: DFMA chains, all SMs; device-to-device copy (read + write bytes)",
                  "see_also": "profiles/r02_dmma_microbench.txt (DFMA 36.8 / DMMA 37.1 TFLOP/s, same pipe)"}))
