for f in tests/test_file_frontend.py tests/test_donor_pool.py tests/test_nif_harness.py tests/test_golden.py tests/test_reference_vectors.py tests/test_parity_gpu.py tests/test_document_models.py; do
  echo "== $f" >> gpurun_out/t11.log
  EB200_DUMP_FLAGS=1 timeout 300 python -m pytest $f -m gpu -q -x --durations=3 2>&1 | tail -8 >> gpurun_out/t11.log
  echo "rc=$?" >> gpurun_out/t11.log
done
