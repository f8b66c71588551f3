#!/bin/bash
# Final verification of the tree on one B200, in the order the driver runs things at round end:
# GPU tests, smoke(), the driver's bench command, the reference arm; then (time permitting) randomised parity.
# Every step is bounded by its own timeout; logs go to gpurun_out/r02_final_*.
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 240 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02_final_pytest_gpu.log 2>&1; el "pytest -m gpu rc=$? $(tail -1 gpurun_out/r02_final_pytest_gpu.log)"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; el "smoke rc=$? $(tail -1 gpurun_out/r02_final_smoke.log)"
timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err; el "bench rc=$? $(head -c 260 gpurun_out/r02_final_bench_n1.json)"
timeout 120 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_final_reference_arm_n1.json 2> /dev/null; el "reference arm rc=$? $(head -c 200 gpurun_out/r02_final_reference_arm_n1.json)"
python - <<'EOF' > gpurun_out/r02_final_loaded_so.log 2>&1
# which shared objects a process that ran the product path has mapped (the driver records the same)
import torch
from stainx_b200 import HistogramMatching
x = (torch.rand(2, 3, 64, 64) * 255).to(torch.uint8).cuda()
HistogramMatching(device="cuda", backend="torch_cuda").fit(x[:1]).transform(x)
torch.cuda.synchronize()
print([l.split()[-1] for l in open("/proc/self/maps") if "stainx" in l][:3])
EOF
el "loaded: $(tail -1 gpurun_out/r02_final_loaded_so.log)"
SEED=${1:-7}
timeout 150 python tools/fuzz_parity.py 250 $SEED > gpurun_out/r02_final_fuzz.log 2>&1; el "fuzz rc=$? $(tail -1 gpurun_out/r02_final_fuzz.log)"
timeout 60 python tools/stress_brackets.py > gpurun_out/r02_final_stress.log 2>&1; el "stress rc=$? $(tail -1 gpurun_out/r02_final_stress.log)"
