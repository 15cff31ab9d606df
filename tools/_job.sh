O=gpurun_out/r3j; mkdir -p $O
timeout 600 python bench.py --distractor-prob 0.4 --steps 5 > $O/bench_mixed.json 2> $O/bench_mixed.err
python - <<PY
import json
d=json.loads(open("$O/bench_mixed.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), d["phase_ms_last_step"], d["config"]["kept_spans"], d["config"]["faces_per_step_per_gpu"], d["config"]["arcface_passes_per_step_per_gpu"], d["bank_last_step"])
print(d["parity"]); print(d["cpu_baseline"])
PY
tail -3 $O/bench_mixed.err
