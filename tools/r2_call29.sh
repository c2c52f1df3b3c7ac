#!/bin/bash
# kernel durations (ncu launch list, serialised) of rank 0's TP8 shard alone on one GPU + plain TP1 for comparison
mkdir -p gpurun_out
B200Q_EMULATE_TP=8 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r2_tp8emu_launches.csv python bench.py --steps 1 --warmup 3 --no-extra > gpurun_out/r2_ncu_tp8emu.log 2>&1
echo "rc $?"; python tools/launch_share.py gpurun_out/r2_tp8emu_launches.csv
B200Q_EMULATE_TP=8 timeout 300 python bench.py --steps 64 --warmup 8 --no-extra 2>/dev/null | tail -1 | cut -c1-200
