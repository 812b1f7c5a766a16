# GPU check of the SAM cs tag -> MIDSV path: parity tests through the C ABI, then a timing of both passes.
set -x
mkdir -p gpurun_out
R=${ROUND:-r02cs}
timeout 600 python -m pytest tests/test_cstag.py -m gpu -q 2>&1 | tail -15 > gpurun_out/${R}_pytest_cstag.log; cat gpurun_out/${R}_pytest_cstag.log
timeout 600 python scripts/cstag_bench.py > gpurun_out/${R}_cstag_bench.json 2> gpurun_out/${R}_cstag_bench.err; tail -3 gpurun_out/${R}_cstag_bench.err; cat gpurun_out/${R}_cstag_bench.json
