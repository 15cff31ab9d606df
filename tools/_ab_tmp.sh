mkdir -p gpurun_out/r2v
PCB_REPLAY_TIMING=1 timeout 600 python bench.py --no-cpu-baseline --frames-per-step 4096 --steps 3 --warmup 2 > gpurun_out/r2v/bench_4096.json 2> gpurun_out/r2v/bench_4096.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2v/bench_4096.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "with-events", round(d["value_with_launch_events"]["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), d["phase_ms_last_step"], d["phase_ms_last_e2e_step"], d["bank_last_step"])
PY
timeout 400 python bench.py --no-cpu-baseline --steps 10 > gpurun_out/r2v/bench_512.json 2> gpurun_out/r2v/bench_512.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2v/bench_512.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "with-events", round(d["value_with_launch_events"]["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), d["phase_ms_last_step"], d["bank_last_step"])
PY
