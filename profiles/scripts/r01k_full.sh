set -x
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2> gpurun_out/bench_full.time; echo rc=$?
tail -c 3500 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.time
( time timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; echo rc=$?
tail -c 1200 gpurun_out/bench_ref.json
