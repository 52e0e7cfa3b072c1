bash tools/gpu_side_profiles.sh r2b "sharpen"
python bench.py --workload pipeline --steps 10 --warmup 3 --frames 512 > gpurun_out/r2b_pipeline.json 2> gpurun_out/r2b_pipeline.err; echo "pipeline rc=$?"; cut -c1-1500 gpurun_out/r2b_pipeline.json; tail -5 gpurun_out/r2b_pipeline.err
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
timeout 600 ncu --set full --clock-control none -k "regex:sharpen" -c 1 -f -o /tmp/r2b_pipe python bench.py --workload pipeline --steps 1 --warmup 0 --no-cpu --frames 148 > gpurun_out/r2b_ncu_pipe.log 2>&1
ncu -i /tmp/r2b_pipe.ncu-rep --page raw --csv > gpurun_out/r2b_pipe_sharpen.raw.csv
cp /tmp/r2b_pipe.ncu-rep gpurun_out/r2b_pipe_sharpen.ncu-rep
du -sh gpurun_out
