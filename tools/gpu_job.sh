python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "one_launch or generator_path or initialize_and_steps_from_oracle or ensemble_members or solve_trajectory or adaptive" 2>&1 | tail -4 > gpurun_out/r3_t1.log
python tools/time_run.py 4096 48 2 > gpurun_out/r3_time.log 2>&1
python tools/time_run.py 148 48 2 >> gpurun_out/r3_time.log 2>&1
PNMOL_B200_LIB=$PWD/tools/variants/teams0.so python tools/time_run.py 4096 48 2 >> gpurun_out/r3_time.log 2>&1
cat gpurun_out/r3_t1.log gpurun_out/r3_time.log
