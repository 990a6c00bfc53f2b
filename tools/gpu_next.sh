#!/bin/bash
# First GPU call after the CPU-only tail of round 2: the two GPU test files that
# have not seen a device yet (general-option matrix, realm_has_vof_ branch),
# then the full GPU suite and the default bench line -- the kernels of every
# other test are byte-identical to the last verified library
# (profiles/r02vof_sass_diff.txt), so a failure here is in the new files.
#   gpurun --timeout 600 -- 'bash tools/gpu_next.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests/test_zy_option_matrix_gpu.py tests/test_zz_vof_gpu.py \
  tests/test_zzz_reference_runs_gpu.py tests/test_zzzz_udiag_post_gpu.py -m gpu -q \
  > gpurun_out/next_new_gpu_tests.log 2>&1
echo "new GPU tests rc=$?"; tail -5 gpurun_out/next_new_gpu_tests.log
python -m pytest tests -m gpu -x -q > gpurun_out/next_pytest_gpu.log 2>&1
echo "GPU suite rc=$?"; tail -3 gpurun_out/next_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/next_bench_default.json 2> gpurun_out/next_bench_default.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/next_bench_default.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "sustained", d["sustained"]["value"],
      "e2e", d["e2e"]["value"], "cpu", {k: d["cpu_baseline"].get(k) for k in ("value", "cores", "openmp_atomic_value", "single_thread_value")})
PY
