#!/usr/bin/env bash
# Turn the outputs of scripts/gpu_final.sh (gpurun_out/) into the committed summaries under profiles/.
set -uo pipefail
export PATH=/usr/local/cuda/bin:$PATH
cd "$(dirname "$0")/.."
tmp=$(mktemp -d)
(cd $tmp && cuobjdump -xelf all $OLDPWD/finetoolsflexstructures.jl_b200/libfsgpu.so >/dev/null 2>&1 && nvdisasm -g -c fsgpu_elements.sm_100a.cubin > el.sass 2>/dev/null)
for k in q4 t3; do ncu -i gpurun_out/${k}_final.ncu-rep --page source --csv > $tmp/${k}src.csv 2>/dev/null; done
(echo "# ncu --set full --clock-control none --import-source on, one launch; report gpurun_out/q4_final.ncu-rep (scratch); final round-1 kernel"; cat gpurun_out/q4_final_summary.txt; echo "--- samples / executed warp-instructions by source function (scripts/ncu_by_function.py)"; python scripts/ncu_by_function.py $tmp/q4src.csv $tmp/el.sass k_q4_stiffnessILb0ELb0E EmitRuns 2>&1 | head -34) > profiles/r01_q4_final_ncu_summary.txt
(echo "# ncu --set full --clock-control none --import-source on, one launch; report gpurun_out/t3_final.ncu-rep (scratch); final round-1 kernel"; cat gpurun_out/t3_final_summary.txt; echo "--- samples / executed warp-instructions by source function (scripts/ncu_by_function.py)"; python scripts/ncu_by_function.py $tmp/t3src.csv $tmp/el.sass k_t3_stiffnessILb0ELb0E EmitRuns 2>&1 | head -30) > profiles/r01_t3_final_ncu_summary.txt
(echo "# ncu --set full --clock-control none, one launch of k_spmv_step on the 4M-element C4 panel; report gpurun_out/expl_final.ncu-rep (scratch)"; cat gpurun_out/expl_final_summary.txt) > profiles/r01_explicit_final_ncu_summary.txt
cp gpurun_out/launches_bench.csv profiles/r01_launches_bench.csv
(echo "# ncu --metrics gpu__time_duration.sum --clock-control none -c 600: python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 --expl-steps 20"; python scripts/launch_summary.py profiles/r01_launches_bench.csv) > profiles/r01_launches_bench_summary.txt 2>&1
cp gpurun_out/bench_1gpu.json profiles/r01_bench_1gpu.json
cp gpurun_out/bench_reference.json profiles/r01_bench_reference_arm.json
python - <<'PY'
import json, re
t = open('profiles/r01_q4_final_ncu_summary.txt').read()
rd = float(re.search(r"dram__bytes_read.sum \[Gbyte\] = ([0-9.]+)", t).group(1)) * 1e9
wr = float(re.search(r"dram__bytes_write.sum \[Gbyte\] = ([0-9.]+)", t).group(1)) * 1e9
json.dump({"nelem": 1000000, "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "source": "profiles/r01_q4_final_ncu_summary.txt (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, k_q4_stiffness<0,0,EmitRuns>, 1M elements)"},
          open('profiles/r01_q4_traffic.json', 'w'), indent=1)
d = json.load(open('profiles/r01_bench_1gpu.json')); ow = d['other_workloads']
print('C2', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['north_star']['frac'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('T3', ow['t3ff_assembly_C4']['value'], ow['t3ff_assembly_C4']['ms_per_step'], ow['t3ff_assembly_C4']['kernel_ms'], ow['t3ff_assembly_C4']['fp64']['frac'], ow['t3ff_assembly_C4']['roofline']['frac'])
print('expl', ow['explicit_C4']['steps_per_s'], ow['explicit_C4']['ms_per_step'], ow['explicit_C4']['roofline']['frac'])
print('C3', ow['t3ffcomp_C3']['stiffness_elements_per_s'], ow['t3ffcomp_C3']['stiffness_ms'], ow['t3ffcomp_C3']['stiffness_kernel_ms'], ow['t3ffcomp_C3']['mass_ms'])
print('C5', ow['corotbeam_C5']['newton_assemblies_per_s'])
print('ref', json.load(open('profiles/r01_bench_reference_arm.json'))['value'])
PY
rm -rf $tmp
