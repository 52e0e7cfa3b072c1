bash tools/gpu_side_profiles.sh r2d "sharpen thumbnail_linear"
for v in base h500 h1000p200 h500p100 h500p100v32 h500p100e200 h2000p400; do
  if [ $v = base ]; then unset VB200_LIB; else export VB200_LIB=$PWD/libvips_b200/variants/libvb200_$v.so; fi
  python bench.py --steps 10 --warmup 3 --no-cpu --frames 592 > gpurun_out/r2d_var_$v.json 2> gpurun_out/r2d_var_$v.err
  echo "$v rc=$? $(python -c "import json;d=json.load(open('gpurun_out/r2d_var_$v.json'));print(d['ms_per_step'], d['roofline']['frac'], d['parity'], d['e2e']['value'])")"
done
unset VB200_LIB
du -sh gpurun_out
