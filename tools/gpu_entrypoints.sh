#!/bin/bash
# usage: tools/gpu_entrypoints.sh <outdir-name>: the command lines of train.py / predict.py on the GPU with synthetic data:
# three epochs with StepLR (step_size = int(epochs / 3) as in the reference), resume for a fourth, InpaintNet training,
# the predict path on the checkpoints just written in all three eval modes, and - with 2+ GPUs - the data-parallel
# launch of train.py under torchrun.
cd "$(dirname "$0")/.."
OUT=gpurun_out/$1
mkdir -p $OUT
EXP=$(mktemp -d)
timeout -k 5 300 python train.py --epochs 3 --synthetic_steps 3 --batch_size 2 --seq_len 8 --bg_mode concat --alpha 0.5 --lr_scheduler StepLR --save_dir $EXP/t > $OUT/train.log 2>&1; echo "train rc=$?" > $OUT/summary.txt
timeout -k 5 300 python train.py --resume_training --epochs 4 --save_dir $EXP/t >> $OUT/train.log 2>&1; echo "train resume rc=$?" >> $OUT/summary.txt
timeout -k 5 300 python train.py --model_name InpaintNet --epochs 2 --synthetic_steps 5 --batch_size 8 --seq_len 16 --save_dir $EXP/i > $OUT/train_inpaint.log 2>&1; echo "train inpaintnet rc=$?" >> $OUT/summary.txt
for mode in nonoverlap average weight; do
  timeout -k 5 300 python predict.py --frames 40 --tracknet_file $EXP/t/TrackNet_cur.pt --inpaintnet_file $EXP/i/InpaintNet_cur.pt --eval_mode $mode --save_dir $EXP/p_$mode > $OUT/predict_$mode.log 2>&1; echo "predict ($mode) rc=$?" >> $OUT/summary.txt
  head -4 $EXP/p_$mode/synthetic_ball.csv | tr '\n' ' ' >> $OUT/summary.txt; echo >> $OUT/summary.txt
done
NG=$(python -c "import torch; print(torch.cuda.device_count())")
if [ "$NG" -ge 2 ]; then
  timeout -k 5 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 train.py --epochs 2 --synthetic_steps 3 --batch_size 2 --seq_len 8 --bg_mode concat --alpha 0.5 --save_dir $EXP/dp > $OUT/train_dp2.log 2>&1; echo "train data-parallel x2 rc=$?" >> $OUT/summary.txt
  tail -3 $OUT/train_dp2.log >> $OUT/summary.txt
fi
ls -la $EXP/t $EXP/i >> $OUT/summary.txt 2>&1
for f in $OUT/train.log $OUT/train_inpaint.log $OUT/predict_weight.log; do tail -n 3 $f >> $OUT/summary.txt; done
cat $OUT/summary.txt
