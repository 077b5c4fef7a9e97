#!/bin/bash
# A/B on the GPU box: bench with the shipped build, then rebuild pa_cells.o with EXTRA flags and bench again.
# usage: tools/gpu_ab.sh TAG "EXTRA flags"
TAG=$1; EXTRA=$2
mkdir -p gpurun_out
PA_DEBUG_TIMING=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_A.json 2> gpurun_out/${TAG}_A.err; echo "A rc=$?"
tail -30 gpurun_out/${TAG}_A.err
cp prealpha_b200/libprealpha_b200.so /tmp/lib_A.so
( cd prealpha_b200/csrc && touch pa_cells.cu && make -s EXTRA="$EXTRA" 2>&1 | grep -i error )
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_B.json 2> gpurun_out/${TAG}_B.err; echo "B rc=$?"
cp /tmp/lib_A.so prealpha_b200/libprealpha_b200.so
python - <<PY
import json
for n in ("A", "B"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % n)); r = d["roofline"]
        print(n, "ms/step %.1f value %.3e e2e %.3e (%.1f ms) frac %.3f useful %.3f atomics/step %.3e dom_ms %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], r["frac"], r["useful_frac"], r["atomics_per_step_per_gpu"], r["ms_per_launch"]))
    except Exception as e:
        print(n, "unreadable", e)
PY
