#!/bin/bash
# usage: tools/gpu_entrypoints.sh <outdir-name>: the command lines of train.py / predict.py on the GPU with synthetic data
# (three epochs with StepLR - step_size = int(epochs / 3) as in the reference -, resume for a fourth, InpaintNet training, then the predict path writing a csv).
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
EXP=$(mktemp -d)
timeout -k 5 300 python train.py --epochs 3 --synthetic_steps 3 --batch_size 2 --seq_len 4 --bg_mode concat --alpha 0.5 --lr_scheduler StepLR --save_dir $EXP/t > $OUT/train.log 2>&1; echo "train rc=$?" > $OUT/summary.txt
timeout -k 5 300 python train.py --resume_training --epochs 4 --save_dir $EXP/t >> $OUT/train.log 2>&1; echo "train resume rc=$?" >> $OUT/summary.txt
timeout -k 5 300 python train.py --model_name InpaintNet --epochs 2 --synthetic_steps 5 --batch_size 8 --seq_len 16 --save_dir $EXP/i > $OUT/train_inpaint.log 2>&1; echo "train inpaintnet rc=$?" >> $OUT/summary.txt
timeout -k 5 300 python predict.py --frames 40 --save_dir $EXP/p > $OUT/predict.log 2>&1; echo "predict rc=$?" >> $OUT/summary.txt
ls -la $EXP/t $EXP/i $EXP/p >> $OUT/summary.txt 2>&1
tail -3 $OUT/train.log $OUT/train_inpaint.log $OUT/predict.log >> $OUT/summary.txt
cat $OUT/summary.txt
