#!/bin/bash
# multi-GPU checks on an N-GPU box: in-process multi-device test, bench at N ranks (c4 weak, c5 strong)
TAG=$1; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${TAG}_smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "multi_device or tile_shards" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_c4_1.json 2> gpurun_out/${TAG}_c4_1.err; echo "c4 N=1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/${TAG}_c4_$N.json 2> gpurun_out/${TAG}_c4_$N.err; echo "c4 N=$N rc=$?"
timeout 600 python bench.py --config c5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_c5_1.json 2> gpurun_out/${TAG}_c5_1.err; echo "c5 N=1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --config c5 --gpus $N --steps 4 --warmup 3 > gpurun_out/${TAG}_c5_$N.json 2> gpurun_out/${TAG}_c5_$N.err; echo "c5 N=$N rc=$?"
python - <<PY
import json
for n in ("c4_1", "c4_$N", "c5_1", "c5_$N"):
    try:
        d = json.loads(open("gpurun_out/${TAG}_%s.json" % n).read().strip().split("\n")[-1])
        print(n, "n_gpus %d ms/step %.1f value %.3e e2e %.3e (%.1f ms) scaling %s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["scaling"]))
    except Exception as e:
        print(n, "unreadable", e)
        try: print(open("gpurun_out/${TAG}_%s.err" % n).read()[-1500:])
        except Exception: pass
PY
