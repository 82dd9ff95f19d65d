#!/bin/bash
# Additional randomised parity sweeps on the B200 (other seeds, more cases) for profiles/.
J=$(nproc); [ $J -gt 16 ] && J=16
python tools/parity_sweep.py 2000 99 --jobs $J | tail -1 > gpurun_out/r02_parity_sweep_2000_seed99_gpu.json
python tools/parity_sweep.py 2000 2024 --jobs $J | tail -1 > gpurun_out/r02_parity_sweep_2000_seed2024_gpu.json
python tools/parity_sweep.py 1000 5 --keyword-double --jobs $J | tail -1 > gpurun_out/r02_parity_sweep_keyword_double_1000_seed5_gpu.json
for f in gpurun_out/r02_parity_sweep_2000_seed99_gpu.json gpurun_out/r02_parity_sweep_2000_seed2024_gpu.json gpurun_out/r02_parity_sweep_keyword_double_1000_seed5_gpu.json; do
python - $f <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], d["cases"], d["verdicts"], d["worst_after_protocol"], "%.0f s" % d["seconds"])
PY
done
