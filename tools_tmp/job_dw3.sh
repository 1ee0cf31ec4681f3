timeout 500 python -m pytest tests -m gpu -x -q -s -k "not i2v" 2>&1 | grep -v "^$" | tail -12
run() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --no-side-legs --no-cpu-baseline > gpurun_out/bench_dw_$tag.json 2> gpurun_out/bench_dw_$tag.err
  python - "$tag" <<'P'
import json,sys
for l in open('gpurun_out/bench_dw_%s.json'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1],round(d['value']/1e6,2),round(d['ms_per_step'],4),{k.replace('umma_',''):round(v['ms_per_launch'],4) for k,v in d['kernels'].items()}, d['last_cost'])
P
}
run t3 A=1
run t1 CTR_DW_1XTF32=1
