# Round-2 evidence, part B (one GPU): captures and phase breakdown of k_run_large at BASELINE configs C2, C3, C4.
mkdir -p gpurun_out
for c in c2 c3 c4; do
  python tools/large_timing.py $c --steps 2 > gpurun_out/fin_large_$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:k_run_large -s 1 -c 1 -f -o gpurun_out/fin_prof_$c python tools/large_timing.py $c --steps 2 > gpurun_out/fin_ncu_$c.log 2>&1
done
PNMOL_B200_LIB=$PWD/tools/variants/prof.so python tools/large_timing.py c2 c3 c4 --steps 4 --profile > gpurun_out/fin_phase_large.txt 2>&1
cat gpurun_out/fin_large_c*.log; tail -1 gpurun_out/fin_ncu_c4.log
