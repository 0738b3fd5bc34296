# development aid: the fuse / lexer / reference-vector subset of the GPU suite, under both builds of the general kernel
for wide in 0 1; do
  echo "== EB200_WIDE=$wide" >> gpurun_out/t12.log
  EB200_WIDE=$wide timeout 400 python -m pytest tests/test_reference_vectors.py tests/test_golden.py -m gpu -q -x 2>&1 | tail -3 >> gpurun_out/t12.log
  EB200_WIDE=$wide timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "structure_mutators or true_default or default_pattern_mix or uri or b64 or sgm_js" 2>&1 | tail -3 >> gpurun_out/t12.log
done
