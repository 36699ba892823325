for i in 1 2; do
PNMOL_B200_LIB=$PWD/tools/variants/v55.so python tools/large_timing.py c4 --steps 8 2>&1 | tail -1 | cut -c1-150 | sed "s/^/v55: /"
python tools/large_timing.py c4 --steps 8 2>&1 | tail -1 | cut -c1-150 | sed "s/^/head: /"
done > gpurun_out/r3_c4.log
cat gpurun_out/r3_c4.log
