#!/bin/bash
# One 8-GPU box visit: weak scaling at 64 frames/GPU, BASELINE.json configs[2] (256 frames over 8 GPUs, 32 per GPU) with its
# 1-GPU denominator, and configs[3] (SMPL-X arm @1024^2, 128 frames over 4 GPUs) with its 1-GPU denominator.
# usage (inside gpurun --gpus 8): bash tools/scale_runs.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29711 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/${TAG}_n8_weak.json 2> gpurun_out/${TAG}_n8_weak.err
timeout 300 $TR --nproc-per-node 8 --master-port 29712 bench.py --gpus 8 --steps 100 --warmup 5 --frames 32 --no-e2e > gpurun_out/${TAG}_c2_n8_32per.json 2> gpurun_out/${TAG}_c2_n8.err
(CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 50 --frames 256 --no-e2e --no-cpu-baseline --no-standin > gpurun_out/${TAG}_c2_n1_256.json 2> gpurun_out/${TAG}_c2_n1.err) &
(CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --steps 30 --model arm --size 1024 --frames 128 --no-e2e --no-cpu-baseline --no-standin > gpurun_out/${TAG}_c3_n1_128.json 2> gpurun_out/${TAG}_c3_n1.err) &
(CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 300 $TR --nproc-per-node 4 --master-port 29713 bench.py --gpus 4 --steps 30 --warmup 5 --model arm --size 1024 --frames 32 --no-e2e > gpurun_out/${TAG}_c3_n4_32per.json 2> gpurun_out/${TAG}_c3_n4.err) &
(CUDA_VISIBLE_DEVICES=2,3 timeout 300 $TR --nproc-per-node 2 --master-port 29714 bench.py --gpus 2 --steps 100 --warmup 5 --no-e2e > gpurun_out/${TAG}_n2_weak.json 2> gpurun_out/${TAG}_n2_weak.err) &
wait
timeout 300 $TR --nproc-per-node 4 --master-port 29715 bench.py --gpus 4 --steps 100 --warmup 5 --no-e2e > gpurun_out/${TAG}_n4_weak.json 2> gpurun_out/${TAG}_n4_weak.err
for f in gpurun_out/${TAG}_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], 'N', d['n_gpus'], 'frames/s', round(d['value']), 'ms', round(d['ms_per_step'], 4), 'per-rank', d.get('per_rank_ms_per_step'), d.get('exchange', '')[:24], 'e2e', round(d['e2e']['value']))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
