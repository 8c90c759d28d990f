#!/bin/bash
# How the round's evidence under profiles/ is captured on the B200 box (run through gpurun; everything lands in
# gpurun_out/, profiles/refresh.py <tag> turns it into the tracked summaries). A number printed under ncu is never a
# bench value: the bench line comes from the first command, un-profiled.
TAG=${1:-fin3}
O=gpurun_out
python bench.py --steps 30 --warmup 3 > $O/bench_$TAG.log 2> $O/bench_$TAG.err
NCU="ncu --clock-control none"
# launch list of the same command (per-launch times, serialised, cold cache: the SHARES must agree with the bench's)
timeout 600 $NCU --metrics gpu__time_duration.sum -k "regex:^k_" -c 600 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-fastq > $O/launches_$TAG.out 2>&1
FULL="$NCU --set full --import-source on"
timeout 600 $FULL -k regex:k_filter_sa -c 1 -f -o $O/prof_k_filter_sa_$TAG python bench.py --steps 1 --warmup 1 --no-fastq > /dev/null 2>&1
timeout 600 $FULL -k regex:k_band -c 1 -f -o $O/prof_k_band_$TAG python bench.py --steps 1 --warmup 1 --no-fastq > /dev/null 2>&1
timeout 600 $FULL -k regex:k_band --launch-skip 1 -c 1 -f -o $O/prof_k_band8_$TAG python bench.py --steps 1 --warmup 1 --no-fastq > /dev/null 2>&1
timeout 600 $FULL -k regex:k_merge_warp -c 1 -f -o $O/prof_k_merge_warp_$TAG python tests/merge_probe.py --pairs 1000000 --cpu-pairs 10 > $O/merge_ncu_$TAG.log 2>&1
timeout 600 $FULL -k regex:k_fq_format -c 1 -f -o $O/prof_k_fq_format_$TAG python tools/fq_ncu_driver.py > /dev/null 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_fastq_$TAG.csv python tools/fq_ncu_driver.py --pe > /dev/null 2>&1
timeout 600 python tests/merge_probe.py > $O/merge_probe_$TAG.log 2>&1
timeout 900 python bench_extra.py > $O/bench_extra_$TAG.log 2>&1
tail -c 600 $O/bench_$TAG.log; ls -la $O/*_$TAG*
