#!/bin/bash
# round 2, call j (1 GPU): the whole GPU suite and the default bench line on the final code; throughput of the
# complex field on c2
set -u
mkdir -p gpurun_out
O=gpurun_out/r02j
timeout 1200 python -m pytest tests -m gpu -x -q > ${O}_pytest_1gpu.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest_1gpu.log; tail -4 ${O}_pytest_1gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${O}_smoke.log
timeout 900 python bench.py > ${O}_bench_n1_default.json 2> ${O}_bench_n1_default.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_n1_reference.json 2> ${O}_bench_n1_reference.err; echo "reference arm rc=$?"
timeout 300 python bench.py --config c2 --field complex --steps 3 > ${O}_bench_c2_complex.json 2> ${O}_bench_c2_complex.err; echo "complex bench rc=$?"
