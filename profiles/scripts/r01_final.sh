set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_final.log 2>&1; echo tests_rc=$?
tail -4 gpurun_out/t_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo smoke_rc=$?
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref_rc=$?
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo bench_rc=$?
tail -1 gpurun_out/bench_full.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['checks'], d['cpu_baseline']['value'], d['clocks'])"
tail -1 gpurun_out/bench_ref.json | head -c 600
