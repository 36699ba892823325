python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not full_size and not c4" 2>&1 | tail -3 > gpurun_out/r3_t1.log
python tools/time_run.py 4096 48 2 2>&1 | tail -1 > gpurun_out/r3_sweep.log
PNMOL_B200_PATH=cta python tools/small_d_probe.py sir17 4096 2>&1 | tail -1 | cut -c1-100 >> gpurun_out/r3_sweep.log
PNMOL_B200_LIB=$PWD/tools/variants/prof.so python tools/phase_profile.py 592 48 > gpurun_out/r3_phase_2cta.log 2>&1
cat gpurun_out/r3_t1.log gpurun_out/r3_sweep.log; sed -n 5,8p gpurun_out/r3_phase_2cta.log; sed -n 15,17p gpurun_out/r3_phase_2cta.log
