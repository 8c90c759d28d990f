#!/bin/bash
# How the round's evidence under profiles/ is captured on the B200 box (run through gpurun). Everything lands in
# gpurun_out/ (<= 64 MiB come back): every ncu --set full capture is digested ON THE BOX into its raw-metrics CSV and a
# per-source-line instruction table (tools/ncu_by_line.py); profiles/refresh.py <tag> copies the digests into profiles/.
# A number printed under ncu is never a bench value: the bench line comes from the first command, un-profiled.
TAG=${1:-r2}
O=gpurun_out
python bench.py --steps 30 --warmup 3 > $O/bench_$TAG.log 2> $O/bench_$TAG.err
NCU="ncu --clock-control none"
# launch list of the same command (per-launch times, serialised, cold cache: the SHARES must agree with the bench's)
timeout 900 $NCU --metrics gpu__time_duration.sum -k "regex:^k_" -c 400 --csv --log-file $O/launches_$TAG.csv python bench.py --steps 2 --warmup 1 --no-fastq --no-configs --no-cpu-baseline > $O/launches_$TAG.out 2>&1
FULL="$NCU --set full --import-source on"
B="python bench.py --steps 1 --warmup 1 --no-fastq --no-configs --no-cpu-baseline"
capture() {   # name, kernel regex, substring of the mangled name, extra ncu args, command...
    local name=$1 rx=$2 sub=$3 extra=$4; shift 4
    timeout 600 $FULL -k regex:$rx $extra -c 1 -f -o $O/prof_${name}_$TAG "$@" > /dev/null 2>&1
    ncu -i $O/prof_${name}_$TAG.ncu-rep --page raw --csv > $O/prof_${name}_${TAG}_raw.csv 2>/dev/null
    python tools/ncu_by_line.py $O/prof_${name}_$TAG.ncu-rep $sub --top 80 > $O/prof_${name}_${TAG}_by_line.txt 2>&1
    rm -f $O/prof_${name}_$TAG.ncu-rep
}
capture k_filter_qg k_filter_qg k_filter_qgILi3 "" $B
capture k_refine k_refine k_refine "" $B
capture k_band16 k_band k_bandILb0ELi16 "" $B
capture k_band8 k_band k_bandILb0ELi8 "--launch-skip 1" $B
capture k_wide k_wide k_wide "" $B
capture k_insert_packed_2x150 k_insert_packed k_insert_packed "" python tools/k2_driver.py --pairs 10000000 --steps 1
capture k_insert_packed_2x300 k_insert_packed k_insert_packed "" python tools/k2_driver.py --len 300 --rate 0.15 --pairs 4000000 --steps 1
python tools/k2_driver.py --pairs 10000000 > $O/k2_150_$TAG.log 2>&1
python tools/k2_driver.py --len 300 --rate 0.15 --pairs 4000000 > $O/k2_300_$TAG.log 2>&1
python tools/panel_driver.py > $O/panel_$TAG.log 2>&1
tail -c 400 $O/bench_$TAG.err; ls -la $O/*_$TAG*
