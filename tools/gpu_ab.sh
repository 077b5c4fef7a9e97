#!/bin/bash
# A/B on the GPU box: bench with the shipped build, then for every further argument rebuild pa_cells.o with those EXTRA
# flags and bench again.  usage: tools/gpu_ab.sh TAG "EXTRA flags 1" ["EXTRA flags 2" ...]
TAG=$1; shift
mkdir -p gpurun_out
cp prealpha_b200/libprealpha_b200.so /tmp/lib_A.so
run() { timeout 600 python bench.py ${BENCH_ARGS:---steps 6 --warmup 3} --no-cpu-baseline > gpurun_out/${TAG}_$1.json 2> gpurun_out/${TAG}_$1.err; echo "$1 rc=$?"; }
run A
n=0
for EXTRA in "$@"; do
  n=$((n+1))
  ( cd prealpha_b200/csrc && touch pa_cells.cu && make -s EXTRA="$EXTRA" 2>&1 | grep -v "^$" | grep -v "warning\|plus1_wrap\|\^\|Remark" | head -5 )
  ls -la prealpha_b200/libprealpha_b200.so | cut -c20-80
  run V$n
done
cp /tmp/lib_A.so prealpha_b200/libprealpha_b200.so
python - "$@" <<PY
import json, sys
names = ["A"] + ["V%d" % (i + 1) for i in range(len(sys.argv) - 1)]
flags = ["(shipped)"] + sys.argv[1:]
for n, f in zip(names, flags):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % n)); r = d["roofline"]
        print(n, f, "ms/step %.1f value %.3e e2e %.3e (%.1f ms)" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"]), r.get("frac"), r.get("useful_frac"))
    except Exception as e:
        print(n, f, "unreadable", e)
PY
