#!/bin/bash
# One gpurun call of the development loop: tests, bench, ncu launch list + full capture of the dominant kernel.
# usage: tools/gpu_round.sh TAG [big] [noncu]     (outputs under gpurun_out/TAG_*)
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
nproc >> gpurun_out/${TAG}_smi.txt
timeout 900 python -m pytest tests -m "gpu and not big" -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
if [[ " $* " == *" big "* ]]; then
  timeout 1200 python -m pytest tests -m "big" -x -q --durations=10 > gpurun_out/${TAG}_pytest_big.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest_big.log
  tail -15 gpurun_out/${TAG}_pytest_big.log
fi
timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
cat gpurun_out/${TAG}_bench.json | cut -c1-2600
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for n in ("bench",):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % n)); r = d["roofline"]
        print(n, "ms/step %.1f value %.3e e2e %.3e frac %.3f useful %.3f atomics/step %.3e evals/step %.3e dom_ms %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], r["frac"], r["useful_frac"], r["atomics_per_step_per_gpu"], r["pair_evals_after_pruning_per_step_per_gpu"], r["ms_per_launch"]))
    except Exception as e:
        print(n, "unreadable", e)
PY
if [[ " $* " != *" noncu "* ]]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pa_rdf_cells_kernel -s 2 -c 2 -o gpurun_out/${TAG}_cells -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
ls -la gpurun_out | grep ${TAG}
