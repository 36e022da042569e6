"""Shared-memory / LSU wavefront budget of one kernel from an .ncu-rep.  usage: python scripts/ncu_lsu.py X.ncu-rep <warps per launch>"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, v = rows[0], rows[2]
d = dict(zip(hdr, v))
nw = float(sys.argv[2])
g = lambda k: float(d.get(k, "nan").replace(",", ""))
tot = g("SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg") * 148
print("--- LSU data-pipe wavefronts (the binding resource of round 1's kernels), per warp of the launch")
print(f"l1tex__data_pipe_lsu_wavefronts: {g('l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed'):.1f} % of peak; total per warp {tot / nw:.0f}")
print(f"  shared {g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / nw:.0f} (ld {g('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum') / nw:.0f}, st {g('l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum') / nw:.0f}), "
      f"global/local {g('SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg') * 148 / nw:.0f}, bank conflicts {g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / nw:.0f}")
print(f"executed warp-instructions per warp: {g('smsp__inst_executed.sum') / nw:.0f}; FP64 pipe {g('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'):.1f} % (DFMA/DMUL/DADD; DMMA not included), issue slots {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} %")
