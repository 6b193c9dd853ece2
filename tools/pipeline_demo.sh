#!/bin/bash
# self-play -> records -> training -> in-loop evaluation, end to end on one GPU (reference: selfplay.py + trainer.py)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
cd "$W" && mkdir -p models logs
PYTHONPATH="$ROOT" python -m 2048nn_b200.selfplay --seed 0 --games 64 | tail -2
PYTHONPATH="$ROOT" python -m 2048nn_b200.trainer --path selfplay/fixed10/ --t_tuple 0 64 --epochs 3 --patience 5 --lr 0.01 --batch_size 2048 --precision "${1:-fp32}" | grep -E "Epoch|mLoss|Best|took|remaining" | head -12
ls models logs
