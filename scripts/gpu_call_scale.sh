#!/bin/bash
# one multi-GPU bench: N=$1 TAG=$2 [extra bench args...]
N=$1; TAG=$2; shift 2
mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N "$@" > gpurun_out/${TAG}.json 2> gpurun_out/${TAG}.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}.json").read().strip().splitlines()[-1])
    print("${TAG}: N=%d ms/step %.4f value %.4g e2e ms %.4f d2h %d parity %s/%s scaling %s" % (d["n_gpus"], d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["e2e"]["d2h_bytes_per_step"], d["parity"]["ok"], d["parity"]["snapshots"], d["scaling"]))
except Exception as e:
    print("no json", e)
PY
tail -3 gpurun_out/${TAG}.err
