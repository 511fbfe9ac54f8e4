#!/bin/bash
# end-of-round validation: GPU suite, smoke(), default bench, reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
P=gpurun_out/r2final
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > ${P}_pytest_all.log 2>&1; echo "exit $?" >> ${P}_pytest_all.log; tail -3 ${P}_pytest_all.log | cut -c1-200; grep -E "^FAILED|^ERROR" ${P}_pytest_all.log | head
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > ${P}_smoke.log 2>&1; tail -2 ${P}_smoke.log | cut -c1-300
timeout 900 python bench.py > ${P}_bench.json 2> ${P}_bench.err; tail -1 ${P}_bench.json | cut -c1-400
timeout 600 python bench.py --impl reference > ${P}_bench_ref.json 2> ${P}_bench_ref.err; tail -1 ${P}_bench_ref.json | cut -c1-600
