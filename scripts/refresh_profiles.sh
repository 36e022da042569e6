#!/usr/bin/env bash
# Turn the outputs of scripts/gpu_final.sh (gpurun_out/) into the committed summaries under profiles/ (round 2 names).
set -uo pipefail
export PATH=/usr/local/cuda/bin:$PATH
cd "$(dirname "$0")/.."
R=r02
tmp=$(mktemp -d)
(cd $tmp && cuobjdump -xelf all $OLDPWD/finetoolsflexstructures.jl_b200/libfsgpu.so >/dev/null 2>&1 && nvdisasm -g -c fsgpu_elements.sm_100a.cubin > el.sass 2>/dev/null)
for k in q4 t3; do ncu -i gpurun_out/${k}_final.ncu-rep --page source --csv > $tmp/${k}src.csv 2>/dev/null; done
(echo "# ncu --set full --clock-control none --import-source on, one launch of the round-2 kernel (DMMA product, staged K_e); report gpurun_out/q4_final.ncu-rep (scratch)"; cat gpurun_out/q4_final_summary.txt; python scripts/ncu_lsu.py gpurun_out/q4_final.ncu-rep 500000; echo "--- samples / executed warp-instructions by source function (scripts/ncu_by_function.py)"; python scripts/ncu_by_function.py $tmp/q4src.csv $tmp/el.sass k_q4_stiffnessILb0ELb0E EmitRuns 2>&1 | head -34) > profiles/${R}_q4_final_ncu_summary.txt
(echo "# ncu --set full --clock-control none --import-source on, one launch of the round-2 kernel; report gpurun_out/t3_final.ncu-rep (scratch)"; cat gpurun_out/t3_final_summary.txt; python scripts/ncu_lsu.py gpurun_out/t3_final.ncu-rep 400000; echo "--- samples / executed warp-instructions by source function (scripts/ncu_by_function.py)"; python scripts/ncu_by_function.py $tmp/t3src.csv $tmp/el.sass k_t3_stiffnessILb0ELb0E EmitRuns 2>&1 | head -30) > profiles/${R}_t3_final_ncu_summary.txt
(echo "# ncu --set full --clock-control none, one launch of k_spmv_step on the 4M-element C4 panel; report gpurun_out/expl_final.ncu-rep (scratch)"; cat gpurun_out/expl_final_summary.txt) > profiles/${R}_explicit_final_ncu_summary.txt
cp gpurun_out/launches_bench.csv profiles/${R}_launches_bench.csv
(echo "# ncu --metrics gpu__time_duration.sum --clock-control none -c 600: python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --expl-steps 20"; python scripts/launch_summary.py profiles/${R}_launches_bench.csv) > profiles/${R}_launches_bench_summary.txt 2>&1
cp gpurun_out/bench_1gpu.json profiles/${R}_bench_1gpu.json
cp gpurun_out/bench_reference.json profiles/${R}_bench_reference_arm.json
cp gpurun_out/fp64_peak.json profiles/${R}_fp64_peak.json
cp gpurun_out/julia_smoke.log profiles/${R}_julia_smoke.txt 2>/dev/null
python - <<'PY'
import json, re
t = open('profiles/r02_q4_final_ncu_summary.txt').read()
rd = float(re.search(r"dram__bytes_read.sum \[Gbyte\] = ([0-9.]+)", t).group(1)) * 1e9
wr = float(re.search(r"dram__bytes_write.sum \[Gbyte\] = ([0-9.]+)", t).group(1)) * 1e9
json.dump({"nelem": 1000000, "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "source": "profiles/r02_q4_final_ncu_summary.txt (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, k_q4_stiffness<0,0,EmitRuns>, 1M elements)"},
          open('profiles/r02_q4_traffic.json', 'w'), indent=1)
PY
rm -rf $tmp
