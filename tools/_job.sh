O=gpurun_out/r3i; mkdir -p $O
for i in 1 2; do timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_$i.log 2>&1; echo "full run $i: $(tail -1 $O/pytest_$i.log)"; grep -E "^FAILED|At index" $O/pytest_$i.log | cut -c1-200; done
for i in 1 2; do timeout 600 python -m pytest tests/test_gpu_e2e.py -m gpu -q > $O/e2e_$i.log 2>&1; echo "e2e run $i: $(tail -1 $O/e2e_$i.log)"; grep -E "^FAILED|At index" $O/e2e_$i.log | cut -c1-200; done
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_20.json 2> $O/bench_20.err
python tools/bench_configs.py > $O/configs.txt 2>&1; grep -E "^C[345]" $O/configs.txt
python - <<PY
import json
for f in ("bench_default","bench_20"):
    d=json.loads(open("$O/%s.json"%f).read().strip().splitlines()[-1])
    print(f, "value", round(d["value"]), "with-events", round(d["value_with_launch_events"]["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],4), d["phase_ms_last_step"], d.get("parity",{}).get("spans_equal"))
PY
