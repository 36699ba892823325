python tools/time_run.py 148 48 2 > gpurun_out/r3_time.log 2>&1
python tools/time_run.py 4096 48 2 >> gpurun_out/r3_time.log 2>&1
PNMOL_B200_LIB=$PWD/tools/variants/cg.so python tools/time_run.py 148 48 2 >> gpurun_out/r3_time.log 2>&1
PNMOL_B200_LIB=$PWD/tools/variants/cg.so python tools/time_run.py 4096 48 2 >> gpurun_out/r3_time.log 2>&1
PNMOL_B200_LIB=$PWD/tools/variants/prof.so python tools/phase_profile.py 148 48 > gpurun_out/r3_phase_1cta.log 2>&1
PNMOL_B200_LIB=$PWD/tools/variants/cgprof.so python tools/phase_profile.py 148 48 > gpurun_out/r3_phase_1cta_cg.log 2>&1
cat gpurun_out/r3_time.log
