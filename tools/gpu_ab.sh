# A/B on the GPU box: the GPU test suite, then the bench
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 20 --warmup 3 > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python -c "
import json;d=json.load(open('gpurun_out/q_bench.json'));print(d['ms_per_step'],d['roofline']['stages_ms']);print(d['ssaa16']['ms_per_frame'],d['ssaa16']['stages_ms']);print(d['raster_info'], d['e2e']['ms_per_step'], d['animation_batch']['ms_per_frame']); print(d['roofline']['fragment_bound'])"
tail -3 gpurun_out/q_bench.err
