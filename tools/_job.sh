mkdir -p gpurun_out/r3d
N=${NG:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3d/bench_n$N.json 2> gpurun_out/r3d/bench_n$N.err
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r3d/bench_n1_on$N.json 2> gpurun_out/r3d/bench_n1_on$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 tools/bench_mainpass.py --steps 3 --warmup 1 > gpurun_out/r3d/mp_n$N.json 2> gpurun_out/r3d/mp_n$N.err
python - <<PY
import json
for f in ("bench_n$N","bench_n1_on$N"):
    d=json.loads(open("gpurun_out/r3d/%s.json"%f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"]), "with-events", round(d["value_with_launch_events"]["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), d["phase_ms_last_step"], d["phase_ms_last_e2e_step"], d["bank_last_step"], d.get("multi_gpu_check"))
d=json.loads(open("gpurun_out/r3d/mp_n$N.json").read().strip().splitlines()[-1])
print("mainpass", d["n_gpus"], round(d["value"],1), d["equals_sequential_main_pass"], d["hits"], d["fixup_rounds"])
PY
