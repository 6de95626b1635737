#!/bin/bash
# Final captures of a round on one B200 (run under gpurun from the repo root): bench line, launch list of the bench command,
# --set full of the RX kernels on the bench mix, AFC mode timing.  Each profiled command first runs to completion without ncu.
# usage: bash profiles/capture_final.sh <tag>      (outputs under gpurun_out/<tag>_*)
set -u
tag=${1:-r02f}
out=gpurun_out
python bench.py --steps 10 --warmup 3 > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err || { echo "bench failed"; tail -5 $out/${tag}_bench_n1.err; exit 1; }
tail -c 300 $out/${tag}_bench_n1.json
python bench.py --steps 3 --warmup 3 > $out/${tag}_bench_s3.json 2> $out/${tag}_bench_s3.err || exit 1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 9000 --csv --log-file $out/${tag}_launches_all.csv \
    python bench.py --steps 3 --warmup 3 > $out/${tag}_ncu_launches.log 2>&1
tail -2 $out/${tag}_ncu_launches.log
python benchmarks/rx_step.py 3 1 > $out/${tag}_rx_step.log 2>&1 || exit 1
M17B_CHAN_GROUPS=1 ncu --set full --clock-control none --import-source on \
    -k "regex:k_frontend|k_sync_frame|k_stream_gather|k_stream_acs|k_decode_frames|k_post" --launch-skip 6 -c 6 -f -o $out/${tag}_rx_full \
    python benchmarks/rx_step.py 1 1 > $out/${tag}_ncu_rx.log 2>&1
tail -2 $out/${tag}_ncu_rx.log
python benchmarks/afc_mode.py > $out/${tag}_afc_mode.json 2> $out/${tag}_afc_mode.err
cat $out/${tag}_afc_mode.json
