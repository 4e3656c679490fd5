#!/bin/bash
# Usage: bash examples/bert/scripts/train_bert_base_dp.sh [GPUS_PER_NODE] (default 8).  Multi-node: set NNODES / NODE_RANK / MASTER_ADDR as for torchrun, or use
# `epl-launch --num_workers N --gpu_per_worker G <script> <args>`.
set -e
cd "$(dirname "$0")/../../.."
GPUS=${1:-8}
# SQUAD_DIR (train-v1.1.json, dev-v1.1.json) and BERT_DIR (vocab.txt) select the real data; without them the script trains on the
# synthetic SQuAD-format corpus run_squad.py generates (plumbing / throughput runs)
DATA=""
[ -n "$SQUAD_DIR" ] && DATA="--train_file $SQUAD_DIR/train-v1.1.json --predict_file $SQUAD_DIR/dev-v1.1.json"
[ -n "$BERT_DIR" ] && DATA="$DATA --vocab_file $BERT_DIR/vocab.txt"
LAUNCH="python -m torch.distributed.run --nnodes=${NNODES:-1} --node-rank=${NODE_RANK:-0} --nproc-per-node $GPUS --master-addr ${MASTER_ADDR:-127.0.0.1} --master-port ${MASTER_PORT:-29500}"
# BERT-base SQuAD-1.1 fine-tune, data parallel: batch 12 per GPU, sequence length 384, 2 epochs (reference train_bert_base_dp.sh)
$LAUNCH examples/bert/run_squad.py --model base --train_batch_size 12 --num_train_steps ${STEPS:-14600} --do_train --do_predict $DATA --max_seq_length 384 --doc_stride 128 --learning_rate 3e-5 --output_dir "${OUT:-/tmp/epl_squad}"
