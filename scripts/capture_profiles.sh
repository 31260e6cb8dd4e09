#!/bin/bash
# Round-2 evidence run (one B200, under gpurun): plain runs first, then the ncu passes of the same commands.
# Outputs land in gpurun_out/; scripts/summarise_profiles.py turns them into the text files under profiles/.
set -x
O=gpurun_out
B="python bench.py --steps 3 --warmup 3 --blocks none --no-cpu-baseline"
$B > $O/r02_prof_bench_plain.json 2> $O/r02_prof_bench_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 60 --csv --log-file $O/launches_r02_bench.csv $B > $O/r02_prof_bench_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:beam_search -s 1 -c 1 -f -o $O/prof_r02_raw10m $B > $O/r02_prof_raw10m.log 2>&1
[ "$1" = "headline" ] && exit 0   # `capture_profiles.sh headline`: only the bench launch (the SQ / ef 280 / brute-force captures are separate calls)
P="python scripts/prof_search.py"
$P --n 1000000 --dim 960 --latent 24 --quant sq8 --ef 180 > $O/r02_prof_sq8_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:beam_search -s 1 -c 1 -f -o $O/prof_r02_sq8_c3 $P --n 1000000 --dim 960 --latent 24 --quant sq8 --ef 180 > $O/r02_prof_sq8.log 2>&1
$P --n 2000000 --dim 768 --latent 24 --quant sq4 --metric ip --topk 100 --ef 160 > $O/r02_prof_sq4_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:beam_search -s 1 -c 1 -f -o $O/prof_r02_sq4_c4 $P --n 2000000 --dim 768 --latent 24 --quant sq4 --metric ip --topk 100 --ef 160 > $O/r02_prof_sq4.log 2>&1
python scripts/bf_time.py > $O/r02_prof_bf_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'bf_|refine|norms|thresholds|merge' -c 40 --csv --log-file $O/launches_r02_bruteforce.csv python scripts/bf_time.py > $O/r02_prof_bf.log 2>&1
tail -n 2 $O/r02_prof_*.log
