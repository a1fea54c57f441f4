#!/bin/bash
# One 8-GPU box (gpurun --gpus 8): N-GPU gather equivalence test, pinned-copy ceilings at N = 1, 2, 4, 8, BASELINE config 5 as
# specified (512 images over 8 GPUs + the NCCL detection all-gather as its own timed leg) and the headline at N = 8 / 4 / 2.
# Outputs under gpurun_out/${1:-r02}_*.   usage: tools/run_8gpu.sh [tag] [--no-ceiling]
tag=${1:-r02}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3 > gpurun_out/${tag}_8gpu_tests.log
if [ "$2" != "--no-ceiling" ]; then
  : > gpurun_out/${tag}_8gpu_h2d.jsonl
  python tools/h2d_bandwidth.py 2>/dev/null | tail -1 >> gpurun_out/${tag}_8gpu_h2d.jsonl
  for n in 2 4 8; do
    $TR --nproc-per-node $n --master-port $((29520+n)) tools/h2d_bandwidth.py 2>/dev/null | tail -1 >> gpurun_out/${tag}_8gpu_h2d.jsonl
  done
  $TR --nproc-per-node 8 --master-port 29531 tools/h2d_bandwidth.py --bind 2>/dev/null | tail -1 >> gpurun_out/${tag}_8gpu_h2d.jsonl
  $TR --nproc-per-node 8 --master-port 29532 tools/h2d_bandwidth.py --duplex 2>/dev/null | tail -1 >> gpurun_out/${tag}_8gpu_h2d.jsonl
fi
$TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --config crowd512 --steps 300 --warmup 10 > gpurun_out/${tag}_8gpu_crowd512.json 2> gpurun_out/${tag}_8gpu_crowd512.err
$TR --nproc-per-node 8 --master-port 29542 bench.py --gpus 8 --steps 500 --warmup 10 > gpurun_out/${tag}_8gpu_headline.json 2> gpurun_out/${tag}_8gpu_headline.err
$TR --nproc-per-node 4 --master-port 29543 bench.py --gpus 4 --steps 500 --warmup 10 > gpurun_out/${tag}_4gpu_headline.json 2> gpurun_out/${tag}_4gpu_headline.err
$TR --nproc-per-node 2 --master-port 29544 bench.py --gpus 2 --steps 500 --warmup 10 > gpurun_out/${tag}_2gpu_headline.json 2> gpurun_out/${tag}_2gpu_headline.err
cat gpurun_out/${tag}_8gpu_tests.log
