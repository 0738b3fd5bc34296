# round-end measurement + as much of the GPU suite as the remaining budget allows (most important first)
timeout 330 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err
tail -c 600 gpurun_out/bench_r2d.json
timeout 90 python bench.py --workload c5 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c5_r2d_n1.json 2> gpurun_out/bench_c5_r2d.err
cut -c1-200 gpurun_out/bench_c5_r2d_n1.json
rm -f gpurun_out/t13.log
for f in tests/test_reference_vectors.py tests/test_golden.py tests/test_nif_harness.py tests/test_donor_pool.py tests/test_file_frontend.py tests/test_parity_gpu.py tests/test_zz_full_size_parity.py tests/test_zzz_async_gpu.py; do
  echo "== $f" >> gpurun_out/t13.log
  timeout 280 python -m pytest $f -m gpu -q -x 2>&1 | tail -4 >> gpurun_out/t13.log
done
grep -E "^==|passed|failed" gpurun_out/t13.log
