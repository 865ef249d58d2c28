# bench stage times under environment switches: tools/ab_env.sh NAME=VALUE[,NAME=VALUE...] ...   ("-" = no switch)
for spec in "$@"; do
  tag=$(echo "$spec" | tr ',=' '__')
  if [ "$spec" = "-" ]; then envs=""; tag=default; else envs=$(echo "$spec" | tr ',' ' '); fi
  env $envs python bench.py --steps 20 --warmup 3 --no-cpu --no-sort-last > gpurun_out/e_$tag.json 2> gpurun_out/e_$tag.err
  python -c "
import json,sys;d=json.load(open('gpurun_out/e_$tag.json'));print('$tag', round(d['ms_per_step'],4),{k:round(x,4) for k,x in d['roofline']['stages_ms'].items()}, '8K', round(d['ssaa16']['ms_per_frame'],4),{k:round(x,4) for k,x in d['ssaa16']['stages_ms'].items()}, 'batch', round(d['animation_batch']['ms_per_frame'],4), 'e2e', round(d['e2e']['ms_per_step'],4), round(d['e2e']['device_transform']['ms_per_step'],4), d['raster_info']['image_checksum'])" || tail -3 gpurun_out/e_$tag.err
done
