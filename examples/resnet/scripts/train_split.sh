#!/bin/bash
# Usage: bash examples/resnet/scripts/train_split.sh [GPUS_PER_NODE] (default 8).  Multi-node: set NNODES / NODE_RANK / MASTER_ADDR as for torchrun, or use
# `epl-launch --num_workers N --gpu_per_worker G <script> <args>`.
set -e
cd "$(dirname "$0")/../../.."
GPUS=${1:-8}
LAUNCH="python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node $GPUS --master-addr ${MASTER_ADDR:-127.0.0.1} --master-port ${MASTER_PORT:-29500}"
# ResNet-50 backbone replicated, 10 000-class head + loss under split(GPUS) (reference scripts/train_split.sh)
$LAUNCH examples/resnet/resnet_split.py --batch 32 --steps ${STEPS:-100} "${@:2}"
