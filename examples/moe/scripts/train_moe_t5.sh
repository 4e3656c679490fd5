#!/bin/bash
# Usage: bash examples/moe/scripts/train_moe_t5.sh [GPUS_PER_NODE] (default 8).  Multi-node: set NNODES / NODE_RANK / MASTER_ADDR as for torchrun, or use
# `epl-launch --num_workers N --gpu_per_worker G <script> <args>`.
set -e
cd "$(dirname "$0")/../../.."
GPUS=${1:-8}
LAUNCH="python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node $GPUS --master-addr ${MASTER_ADDR:-127.0.0.1} --master-port ${MASTER_PORT:-29500}"
# MoE-T5-small: 8 experts, top-2 gating, capacity factor 1.25, 2 x 512 tokens per GPU (reference scripts/train_moe_t5.sh)
$LAUNCH examples/moe/train_t5_moe.py --size small --experts 8 --gating top2 --capacity_factor 1.25 --batch 2 --seq 512 --steps ${STEPS:-100} "${@:2}"
