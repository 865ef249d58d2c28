# bench stage times of library variants: tools/variants.sh base name1 name2 ...   (base = the real library)
for v in "$@"; do
  if [ "$v" = "base" ]; then unset FGL_LIB; else export FGL_LIB=fauxgl_b200/libfauxgl_b200.$v.so; fi
  python bench.py --steps 20 --warmup 3 --no-cpu --no-sort-last --no-batch > gpurun_out/v_$v.json 2> gpurun_out/v_$v.err
  python -c "
import json,sys;d=json.load(open('gpurun_out/v_$v.json'));print('$v', round(d['ms_per_step'],4),{k:round(x,4) for k,x in d['roofline']['stages_ms'].items()}, '8K', round(d['ssaa16']['ms_per_frame'],4),{k:round(x,4) for k,x in d['ssaa16']['stages_ms'].items()}, d['raster_info']['image_checksum'])" || tail -3 gpurun_out/v_$v.err
done
