#!/bin/bash
# Round evidence on ONE B200: GPU test suite, smoke, every bench configuration, the default bench line, the reference arm, the ncu
# launch list of the default bench command and `ncu --set full` captures of the pipeline kernels.  Outputs under gpurun_out/$1_*.
tag=${1:-r02}
o=gpurun_out/$tag
python -m pytest tests -m gpu -q 2>&1 | tail -4 > ${o}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > ${o}_smoke.log 2>&1
python bench.py > ${o}_bench_1gpu.json 2> ${o}_bench_1gpu.err
tools/run_all_configs.sh ${o}_all_configs_1gpu.jsonl 500
python bench.py --impl reference --steps 3 --warmup 1 --cpu-budget-s 60 > ${o}_bench_reference.json 2> /dev/null
python tools/api_latency.py > ${o}_api_latency.txt 2>&1
# ncu: launch list of the short default command (after it exited 0 without ncu), then full captures of the four pipeline kernels
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${o}_launches.csv $CMD > /dev/null 2>&1
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'yolo_decode_filter|cluster_sort|nms_segment|yolo_emit' -s 16 -c 5 -o ${o}_full $CMD > /dev/null 2>&1
CMD2="python bench.py --config retina800 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD2 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:'prior_decode_filter' -s 4 -c 1 -o ${o}_full_retina $CMD2 > /dev/null 2>&1
CMD3="python bench.py --config crowd512 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD3 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:'yolo_decode_filter' -s 4 -c 1 -o ${o}_full_crowd $CMD3 > /dev/null 2>&1
CMD4="python bench.py --config cfg2 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1"
$CMD4 > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:'yolo_decode_filter' -s 4 -c 1 -o ${o}_full_cfg2 $CMD4 > /dev/null 2>&1
cat ${o}_gpu_tests.log ${o}_smoke.log
