# quick GPU check: the GPU test suite (both front ends are covered by tests/test_parity_gpu.py) and a bench summary
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python -c "
import json;d=json.load(open('gpurun_out/q_bench.json'));print(d['ms_per_step'],d['roofline']['stages_ms']);print(d['ssaa16']['ms_per_frame'],d['ssaa16']['stages_ms']);print(d['raster_info'], d['e2e']['ms_per_step'])"
tail -3 gpurun_out/q_bench.err
