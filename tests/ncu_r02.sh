set -x
B="python bench.py --steps 2 --warmup 1 --strong-bins 0 --arm-bins 0 --large-n 0 --no-cpu-baseline"
$B > gpurun_out/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
python tests/prof_one.py 2000 2 > gpurun_out/r02_plain_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:osj_kernel|cholinv8|coniss_sweep|ch_kernel|io_gemm|ig_gram|dgemm_kernel' -s 82 -c 82 -f -o gpurun_out/r02_full_n2000 python tests/prof_one.py 2000 2 > gpurun_out/r02_ncu_one.log 2>&1
python tests/prof_one.py 8000 2 > gpurun_out/r02_plain_8k.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:io_gemm|ig_gram' -s 106 -c 10 -f -o gpurun_out/r02_full_n8000 python tests/prof_one.py 8000 2 > gpurun_out/r02_ncu_8k.log 2>&1
ls -la gpurun_out | tail -12; tail -3 gpurun_out/r02_ncu_bench.log gpurun_out/r02_ncu_one.log gpurun_out/r02_ncu_8k.log
