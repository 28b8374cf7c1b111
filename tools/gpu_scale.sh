#!/bin/bash
# bench.py under torchrun exactly as the driver launches it.  usage: gpu_scale.sh <tag> <N>   (gpurun --gpus N)
T=${1:-r02}; N=${2:-2}
mkdir -p gpurun_out
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${STEPS:-2000} --warmup ${WARMUP:-50} > gpurun_out/${T}_scale_n$N.json 2> gpurun_out/${T}_scale_n$N.err
echo "own arm rc=$?"; cut -c1-600 gpurun_out/${T}_scale_n$N.json; tail -n 3 gpurun_out/${T}_scale_n$N.err
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 100 --warmup 10 > gpurun_out/${T}_scale_ref_n$N.json 2> gpurun_out/${T}_scale_ref_n$N.err
echo "reference arm rc=$?"; cut -c1-600 gpurun_out/${T}_scale_ref_n$N.json; tail -n 3 gpurun_out/${T}_scale_ref_n$N.err
python - <<P
import json
d = json.loads(open("gpurun_out/${T}_scale_n$N.json").read().splitlines()[-1])
for k in ("value", "ms_per_step_per_rank", "e2e", "multi_stream", "config4_8x4k_per_gpu", "host"):
    print(k, d.get(k))
print("roofline", {k: d["roofline"][k] for k in ("frac", "kernel_ms", "achieved")})
P
