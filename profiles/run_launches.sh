#!/bin/bash
# launch list of the bench command (200 cases so that the ncu pass stays short), after the same command without ncu
cd /root/repo
export AEW_BENCH_CASES=200
python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/b_small3.json 2> gpurun_out/b_small3.err; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^k[1-9]_' -c 700 --csv --log-file gpurun_out/launches3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_l3.log 2>&1; echo "ncu rc=$?"
python profiles/launch_list.py gpurun_out/launches3.csv | tee gpurun_out/launches3.txt
# only the launches of the timed steps of the headline path (NVTX range pushed by bench.py)
ncu --nvtx --nvtx-include "timed_steps/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_timed.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_l4.log 2>&1; echo "ncu (timed steps only) rc=$?"
python profiles/launch_list.py gpurun_out/launches_timed.csv | tee gpurun_out/launches_timed.txt
