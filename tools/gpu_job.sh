python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "one_launch or generator_path or ensemble_members or solve_trajectory or simulate_final" 2>&1 | tail -3 > gpurun_out/r3_t1.log
for n in 50 51 54; do export PNMOL_NUM=$n; for g in 296 288; do PNMOL_B200_GRID=$g python tools/time_run.py 2368 48 1 2>&1 | tail -1 | sed "s/.*: members/heat$n grid $g: members/"; done; done > gpurun_out/r3_sweep.log
unset PNMOL_NUM
python tools/time_run.py 4096 48 2 2>&1 | tail -1 >> gpurun_out/r3_sweep.log
PNMOL_B200_PACE=0 python tools/time_run.py 4096 48 2 2>&1 | tail -1 | sed "s/default/nopace/" >> gpurun_out/r3_sweep.log
PNMOL_B200_PATH=cta python tools/small_d_probe.py sir17 4096 2>&1 | tail -1 | cut -c1-100 >> gpurun_out/r3_sweep.log
PNMOL_B200_PATH=cta python tools/small_d_probe.py heat24 4096 2>&1 | tail -1 | cut -c1-100 >> gpurun_out/r3_sweep.log
cat gpurun_out/r3_t1.log gpurun_out/r3_sweep.log
