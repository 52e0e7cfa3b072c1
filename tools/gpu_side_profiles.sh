#!/bin/bash
# One gpurun call: the -m gpu suite, then for every side workload a bench line and one
# `ncu --set full` capture of its kernels, reduced ON THE BOX to the raw-page CSV (gpurun_out/ is
# capped at 64 MiB; a .ncu-rep with sources is 5-30 MB per workload).
#   gpurun --timeout 1500 -- 'bash tools/gpu_side_profiles.sh r2b "colour convsep upsize sharpen icc reduce49"'
tag=${1:-r2}
workloads=${2:-"colour convsep upsize sharpen icc reduce49"}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
tail -15 $out/${tag}_pytest.log
for w in $workloads; do
	timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > $out/${tag}_side_$w.json 2> $out/${tag}_side_$w.err
	echo "$w bench rc=$?"; cut -c1-300 $out/${tag}_side_$w.json; tail -3 $out/${tag}_side_$w.err
	timeout 600 ncu --set full --clock-control none -k "regex:colour|conv|affine|sharpen|extract|icc|reduce|shrink|thumbnail|linear" -c 4 -f -o /tmp/${tag}_$w \
		python bench.py --workload $w --steps 1 --warmup 0 --no-cpu > $out/${tag}_ncu_$w.log 2>&1
	echo "$w ncu rc=$?"
	ncu -i /tmp/${tag}_$w.ncu-rep --page raw --csv > $out/${tag}_$w.raw.csv 2>/dev/null
done
du -sh $out
