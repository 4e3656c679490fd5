#!/bin/bash
# Usage: bash examples/resnet/scripts/train_dp.sh [GPUS_PER_NODE] (default 8).  Multi-node: set NNODES / NODE_RANK / MASTER_ADDR as for torchrun, or use
# `epl-launch --num_workers N --gpu_per_worker G <script> <args>`.
set -e
cd "$(dirname "$0")/../../.."
GPUS=${1:-8}
LAUNCH="python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node $GPUS --master-addr ${MASTER_ADDR:-127.0.0.1} --master-port ${MASTER_PORT:-29500}"
# ResNet-50 + 10 000-class head, data parallel, batch 32 per GPU, synthetic images (reference scripts/train_dp.sh).  Toggles: --gc auto --amp O1|bf16 --zero v1
$LAUNCH examples/resnet/resnet_dp.py --batch 32 --steps ${STEPS:-100} "${@:2}"
