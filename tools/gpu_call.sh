bash tools/gpu_side_profiles.sh r2g "thumbnail_linear upsize"
du -sh gpurun_out
