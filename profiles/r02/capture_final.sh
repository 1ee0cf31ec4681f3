#!/bin/bash
# Evidence refresh on the final tree (one B200, `gpurun -- bash profiles/r02/capture_final.sh`): launch list, --set full of
# the tcgen05 GEMMs (3xTF32 weight gradients, transposed dX), memcheck over the tcgen05 / key-fed tests.  Numbers printed
# under ncu / the sanitizer are never bench values.
set -x
B="python bench.py --steps 2 --warmup 3 --no-side-legs --no-cpu-baseline"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_final.csv $B > gpurun_out/ncu_list_r02_final.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_umma --launch-skip 18 --launch-count 6 -o gpurun_out/umma_r02_final -f $B > gpurun_out/ncu_umma_r02_final.log 2>&1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_umma.py tests/test_gpu_round2.py -q -x -k "(umma and not large_batch and not trajectory and not raw_tile) or keys or out_of_range or table_adam" > gpurun_out/sanitizer_memcheck_umma_keys_final.log 2>&1; echo "memcheck umma/keys rc=$?" > gpurun_out/sanitizer_summary_final.log
tail -3 gpurun_out/sanitizer_memcheck_umma_keys_final.log; cat gpurun_out/sanitizer_summary_final.log
