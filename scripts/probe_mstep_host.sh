# Host-side breakdown of the M-step on the headline workload (phase_ms_per_step: in library / between rounds /
# before the first round / after the last round).
timeout 300 python bench.py --gpus 1 --steps 6 --warmup 3 --legs "" --no-cpu-baseline > gpurun_out/r2_mstep_host.json 2> gpurun_out/r2_mstep_host.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_mstep_host.json').read().strip().splitlines()[-1])
print(d['value'], d['phase_ms_per_step'])"
