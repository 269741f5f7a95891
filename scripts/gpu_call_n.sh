#!/bin/bash
# multi-GPU: peer tests + bench at N = $1 (tag $2)
N=${1:-2}; TAG=${2:-r02d}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_peer.py -x -q > gpurun_out/${TAG}_peer.log 2>&1; echo "peer pytest rc=$?"; tail -5 gpurun_out/${TAG}_peer.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/${TAG}_bench_n$N.json; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "d2h", d["e2e"]["d2h_bytes_per_step"], "parity", d["parity"])
except Exception as e:
    print("no json", e)
PY
tail -5 gpurun_out/${TAG}_bench_n$N.err
AMOD_BENCH_INTERLEAVED=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n${N}_interleaved.json 2> gpurun_out/${TAG}_bench_n${N}_interleaved.err; echo "bench interleaved rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_n${N}_interleaved.json").read().strip().splitlines()[-1])
    print("interleaved N=$N ms/step", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "parity", d["parity"])
except Exception as e:
    print("no json", e)
PY
