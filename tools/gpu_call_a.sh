#!/bin/bash
# round-2 GPU call A: whole GPU suite, encoder breakdown, encoder launch list, bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/a_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/a_tests.log
python tools/encode_cloud_profile.py > gpurun_out/a_encprof.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a_enc_launches.csv python tools/encode_cloud_profile.py --small > gpurun_out/a_enc_ncu.log 2>&1
python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -3 gpurun_out/a_tests.log; cat gpurun_out/a_encprof.log | tail -3; cat gpurun_out/a_bench.json | head -c 3000
