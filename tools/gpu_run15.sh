mkdir -p gpurun_out
for w in c5 c4; do
timeout 600 python tools/prof_frame.py --workload $w --frames 2 --stats gpurun_out/level_stats_r1_$w.json > gpurun_out/stats_$w.log 2>&1; tail -4 gpurun_out/stats_$w.log
done
