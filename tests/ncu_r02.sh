# Round-2 ncu evidence (run under gpurun on one B200): launch list of a short bench.py run, --set full counters of the
# kernels of one 2000-bin call and of the tensor kernels of one 8000-bin call.  Reports are turned into CSV on the box
# (gpurun brings back at most 64 MiB) -- profiles/summarize.py makes the committed tables from them.
set -x
B="python bench.py --steps 1 --warmup 1 --batch 2 --streams 2 --strong-bins 0 --arm-bins 0 --large-n 0 --no-cpu-baseline"
$B > gpurun_out/r02_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
python tests/prof_one.py 2000 2 > gpurun_out/r02_plain_one.log 2>&1 && \
ncu --set full --clock-control none -k 'regex:osj_kernel|cholinv8|coniss_sweep|ch_kernel|io_gemm|ig_gram|dgemm_kernel' -s 82 -c 82 -f -o /tmp/r02_full_n2000 python tests/prof_one.py 2000 2 > gpurun_out/r02_ncu_one.log 2>&1
ncu -i /tmp/r02_full_n2000.ncu-rep --page raw --csv > gpurun_out/r02_full_n2000_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k 'regex:osj_kernel|coniss_sweep' -s 4 -c 2 -f -o gpurun_out/r02_src_osj_sweep python tests/prof_one.py 2000 2 > gpurun_out/r02_ncu_src.log 2>&1
python tests/prof_one.py 8000 2 > gpurun_out/r02_plain_8k.log 2>&1 && \
ncu --set full --clock-control none -k 'regex:io_gemm|ig_gram' -s 106 -c 8 -f -o /tmp/r02_full_n8000 python tests/prof_one.py 8000 2 > gpurun_out/r02_ncu_8k.log 2>&1
ncu -i /tmp/r02_full_n8000.ncu-rep --page raw --csv > gpurun_out/r02_full_n8000_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out
