#!/bin/bash
# Final bench lines of the round at N = 1 (the driver's command first), kept under profiles/.
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "default rc=$?"
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference_arm_n1.json 2>/dev/null
for c in c1 c3 c4 c5 reinhard; do python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r02_bench_${c}_n1.json 2> gpurun_out/r02_bench_${c}_n1.err; echo "$c rc=$?"; done
python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
