#!/bin/bash
# the 8-GPU measurements of a round (charged 8x): bench c4 weak scaling, c5 strong scaling, CLI from files
TAG=$1; N=${2:-8}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt; nproc >> gpurun_out/${TAG}_smi.txt
[ -n "$SKIP_C4" ] || timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/${TAG}_c4_$N.json 2> gpurun_out/${TAG}_c4_$N.err; echo "c4 N=$N rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --config c5 --gpus $N --steps 4 --warmup 3 > gpurun_out/${TAG}_c5_$N.json 2> gpurun_out/${TAG}_c5_$N.err; echo "c5 N=$N rc=$?"
timeout 1500 python tools/cli_multi_gpu_file.py ${3:-48} $N > gpurun_out/${TAG}_cli_files.jsonl 2> gpurun_out/${TAG}_cli_files.err; echo "cli rc=$?"
cut -c1-700 gpurun_out/${TAG}_cli_files.jsonl; tail -5 gpurun_out/${TAG}_cli_files.err
python - <<PY
import json
for n in ("c4_$N", "c5_$N"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_%s.json" % n).read().strip().split("\n")[-1])
        print(n, "n_gpus %d ms/step %.1f value %.3e e2e %.3e (%.1f ms) scaling %s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["scaling"]))
    except Exception as e:
        print(n, "unreadable", e)
        try: print(open("gpurun_out/${TAG}_%s.err" % n).read()[-1500:])
        except Exception: pass
PY
