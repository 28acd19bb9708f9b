# ig_gram_kernel (exact int8 Gram of the correlation) and the sliced Gram of M at 8000 bins after the band-ordered tile walk
set -x
python tests/prof_one.py 8000 1 > gpurun_out/r02_plain_gram.log 2>&1 && \
ncu --set full --clock-control none -k 'regex:ig_gram|io_gemm_kernel<8, 64>' -c 2 -f -o /tmp/r02_gram python tests/prof_one.py 8000 1 > gpurun_out/r02_ncu_gram.log 2>&1
ncu -i /tmp/r02_gram.ncu-rep --page raw --csv > gpurun_out/r02_gram_n8000_raw.csv 2>/dev/null
ls -la gpurun_out
