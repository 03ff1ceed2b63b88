#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/c18_n2_cfg3.json 2> gpurun_out/c18_n2_cfg3.err; tail -2 gpurun_out/c18_n2_cfg3.err; cut -c1-200 gpurun_out/c18_n2_cfg3.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --config 4 > gpurun_out/c18_n2_cfg4.json 2> gpurun_out/c18_n2_cfg4.err; tail -2 gpurun_out/c18_n2_cfg4.err; cut -c1-200 gpurun_out/c18_n2_cfg4.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/c18_n2_ref.json 2> gpurun_out/c18_n2_ref.err; cut -c1-200 gpurun_out/c18_n2_ref.json
