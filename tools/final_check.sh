#!/bin/bash
# End-of-round visit of one GPU box: full parity suite, smoke, the bench lines of the three workloads, gather A/B + timeline.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; tail -3 gpurun_out/final_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
RENET_STREAM_TL=1 timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -2 gpurun_out/final_bench.err
timeout 600 python bench.py --workload gdelt --steps 50 --warmup 5 --no-train --no-e2e > gpurun_out/final_bench_gdelt.json 2> gpurun_out/final_bench_gdelt.err
timeout 600 python bench.py --workload synth1m --steps 6 --warmup 3 --no-train --no-e2e > gpurun_out/final_bench_synth1m.json 2> gpurun_out/final_bench_synth1m.err
bash tools/gather_ab.sh icews18 0h 0 1h 2h > gpurun_out/final_gather_ab.txt 2>&1
RENET_STREAM_WARPS=32 python tools/stream_timeline.py icews18 hot > gpurun_out/final_timeline.txt 2>&1
ncu --set full --import-source on --clock-control none -k "regex:rgcn_gather_stream_kernel|rgcn_dw_d200|sgemm_tn_splitk|adam_step" -s 4 -c 10 -f -o gpurun_out/r02_bwd2 \
  python bench.py --mode train --steps 1 --warmup 1 --pool 2 > gpurun_out/r02_bwd2.log 2>&1
python - <<'PY'
import json
for f in ('final_bench', 'final_bench_gdelt', 'final_bench_synth1m'):
    try:
        d = json.load(open('gpurun_out/%s.json' % f))
        print(f, 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], 'roofline %.3f (%.1f us)' % (d['roofline']['frac'], d['roofline']['avg_launch_us']),
              'e2e', (d.get('e2e') or {}).get('ms_per_step'), 'train', (d.get('train') or {}).get('ms_per_step'))
    except Exception as ex:
        print(f, 'failed', ex)
PY
