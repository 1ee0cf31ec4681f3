#!/bin/bash
# Round-2 evidence captures (one B200, `gpurun -- bash profiles/r02/capture.sh`).  Numbers printed under ncu / the
# sanitizer are never bench values; outputs land in gpurun_out/ and are summarised into profiles/r02/ by summarize.py.
set -x
B="python bench.py --steps 2 --warmup 3 --no-side-legs --no-cpu-baseline"
# 1. launch list of the bench command (share of each kernel in the step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_list_r02.log 2>&1
# 2. --set full of the two dominant kernels on the 100M-row table (the bench's own workload)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_attn_.*_idx --launch-skip 6 --launch-count 2 -o gpurun_out/attn_idx_r02 -f $B > gpurun_out/ncu_attn_r02.log 2>&1
# 3. the tcgen05 GEMMs
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_umma --launch-skip 18 --launch-count 6 -o gpurun_out/umma_r02 -f $B > gpurun_out/ncu_umma_r02.log 2>&1
# 4. compute-sanitizer: memcheck over the index path (fused atomics, cp.async id slots), the tcgen05 / mbarrier pipelines and the key-fed path
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "gather_is_bit_exact or forward_scores_and_logits or fused_atomic or test_gradients" > gpurun_out/sanitizer_memcheck_parity.log 2>&1; echo "memcheck parity rc=$?" >> gpurun_out/sanitizer_summary.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_umma.py tests/test_gpu_round2.py -q -x -k "umma or keys or out_of_range or table_adam" > gpurun_out/sanitizer_memcheck_umma_keys.log 2>&1; echo "memcheck umma/keys rc=$?" >> gpurun_out/sanitizer_summary.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -x -k "forward_scores_and_logits and ns" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_summary.log
tail -3 gpurun_out/sanitizer_memcheck_parity.log gpurun_out/sanitizer_memcheck_umma_keys.log gpurun_out/sanitizer_racecheck.log; cat gpurun_out/sanitizer_summary.log
