CTR_UMMA_KBK=16 timeout 400 python -m pytest tests/test_gpu_umma.py tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_kbk16.log 2>&1; tail -4 gpurun_out/pytest_kbk16.log
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-side-legs --no-cpu-baseline > gpurun_out/bench_kbk_$tag.json 2> gpurun_out/bench_kbk_$tag.err
  python - "$tag" <<'P'
import json,sys
for l in open('gpurun_out/bench_kbk_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1],round(d['value']/1e6,2),round(d['ms_per_step'],4),{k.replace('umma_',''):round(v['ms_per_launch'],4) for k,v in d['kernels'].items()}, d['last_cost'])
P
}
run k16 CTR_UMMA_KBK=16
run k32 CTR_UMMA_KBK=32
rm -f gpurun_out/umma_timeline.txt
PYTHONPATH=. CTR_UMMA_KBK=16 CTR_UMMA_TIMELINE=1 timeout 120 python tests/_timeline_probe.py > /dev/null 2>&1; mv gpurun_out/umma_timeline.txt gpurun_out/umma_timeline_kbk16.txt
