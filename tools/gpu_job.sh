for v in anti head; do lib=$PWD/tools/variants/$v.so; [ $v = head ] && lib=$PWD/pnmol-experiments_b200/pnmol_b200/libpnmol_b200.so
PNMOL_B200_LIB=$lib python tools/time_run.py 4096 48 2 2>&1 | tail -1 | sed "s/.*: members/$v heat50: members/"
PNMOL_B200_LIB=$lib PNMOL_B200_PATH=cta python tools/small_d_probe.py sir17 4096 2>&1 | tail -1 | cut -c1-100 | sed "s/^/$v /"
done > gpurun_out/r3_sweep.log
cat gpurun_out/r3_sweep.log
