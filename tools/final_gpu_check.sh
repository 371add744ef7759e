#!/bin/bash
# one gpurun call: the GPU suite, the ME suite under the balanced schedule, the ME A/B, smoke, the default bench line
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest_gpu.log
tail -3 gpurun_out/f_pytest_gpu.log
( CCGP_ME_BALANCED=1 timeout 200 python -m pytest tests -m gpu -x -q -k "me_ or entropy or paired or stencil or subdesigns or smoke" ) > gpurun_out/f_pytest_me_balanced.log 2>&1; echo "rc=$?" >> gpurun_out/f_pytest_me_balanced.log
tail -2 gpurun_out/f_pytest_me_balanced.log
timeout 120 python tools/time_me.py 1000 5 ab > gpurun_out/f_me_ab.txt 2>&1; cat gpurun_out/f_me_ab.txt
timeout 60 python tools/time_me.py 60 5 ab > gpurun_out/f_me_ab_p60.txt 2>&1; cat gpurun_out/f_me_ab_p60.txt
timeout 120 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; tail -1 gpurun_out/f_smoke.log
( time timeout 400 python bench.py ) > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/f_bench.err; head -c 600 gpurun_out/f_bench.json
