# one --set full capture of the four hot kernels of a C2 frame (second frame), k_secondary at the pipelined grid
mkdir -p gpurun_out
export PGRT_SECONDARY_CTAS=296
python tools/prof_frame.py --frames 3 > gpurun_out/plain_prof2.log 2>&1 || { tail -5 gpurun_out/plain_prof2.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_secondary|k_shade|k_phong" -s 4 -c 4 -f -o gpurun_out/prof_final2_r1 python tools/prof_frame.py --frames 3 > gpurun_out/ncu_prof2.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_final2_r1.ncu-rep; tail -3 gpurun_out/ncu_prof2.log | cut -c1-300
