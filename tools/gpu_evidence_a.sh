# Round-2 evidence, part A (one GPU): tests, bench arms, launch list, captures of k_run (C5) and k_run_small, phase breakdowns.
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/fin_pytest_gpu.log
python bench.py > gpurun_out/fin_bench.json 2> gpurun_out/fin_bench.err
python bench.py --impl reference > gpurun_out/fin_bench_reference.json 2> gpurun_out/fin_bench_reference.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/fin_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/fin_ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/fin_ncu_bench.log 2>&1
python tools/short_run.py 296 8 > gpurun_out/fin_short.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_run -s 1 -c 1 -f -o gpurun_out/fin_prof_c5 python tools/short_run.py 296 8 > gpurun_out/fin_ncu_c5.log 2>&1
python tools/small_d_probe.py heat6 16384 > gpurun_out/fin_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_run_small -s 1 -c 1 -f -o gpurun_out/fin_prof_small python tools/small_d_probe.py heat6 16384 > gpurun_out/fin_ncu_small.log 2>&1
PNMOL_B200_LIB=$PWD/tools/variants/prof.so python tools/phase_profile.py 592 48 > gpurun_out/fin_phase_c5.txt 2>&1
PNMOL_B200_LIB=$PWD/tools/variants/prof.so python tools/phase_profile.py 148 48 >> gpurun_out/fin_phase_c5.txt 2>&1
PNMOL_B200_LIB=$PWD/tools/variants/prof.so python tools/phase_profile_small.py 6 16384 > gpurun_out/fin_phase_small.txt 2>&1
python tools/parity_margins.py gpurun_out/fin_parity_margins.json > gpurun_out/fin_margins.log 2>&1
tail -3 gpurun_out/fin_pytest_gpu.log; cut -c1-400 gpurun_out/fin_bench.json; tail -n 2 gpurun_out/fin_ncu_c5.log; tail -n 2 gpurun_out/fin_ncu_small.log
