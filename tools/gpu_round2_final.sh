#!/bin/bash
# One gpurun call: the -m gpu suite, the JPEG staging lines, bench + ncu (--set full, reduced on the box to the raw
# page CSV / the per-line summary) for the kernels touched late in round 2, and the launch list of the default bench.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round2_final.sh r2k'
tag=${1:-r2k}
out=gpurun_out
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 $out/${tag}_pytest.log
timeout 500 python bench.py --workload thumbnail_jpeg --steps 5 --warmup 2 > $out/${tag}_jpeg.json 2> $out/${tag}_jpeg.err
echo "jpeg rc=$?"; cut -c1-1500 $out/${tag}_jpeg.json; tail -3 $out/${tag}_jpeg.err
timeout 500 python bench.py --workload thumbnail_jpeg_norestart --steps 2 --warmup 2 --no-cpu > $out/${tag}_jpeg_nr.json 2> $out/${tag}_jpeg_nr.err
echo "jpeg_nr rc=$?"; cut -c1-1000 $out/${tag}_jpeg_nr.json; tail -3 $out/${tag}_jpeg_nr.err
for w in reduce49 icc; do
	timeout 300 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu > $out/${tag}_side_$w.json 2> $out/${tag}_side_$w.err
	echo "$w bench rc=$?"; cut -c1-300 $out/${tag}_side_$w.json; tail -3 $out/${tag}_side_$w.err
	timeout 600 ncu --set full --clock-control none -k "regex:icc|reduce" -c 4 -f -o /tmp/${tag}_$w \
		python bench.py --workload $w --steps 1 --warmup 0 --no-cpu > $out/${tag}_ncu_$w.log 2>&1
	echo "$w ncu rc=$?"
	ncu -i /tmp/${tag}_$w.ncu-rep --page raw --csv > $out/${tag}_$w.raw.csv 2>/dev/null
done
timeout 600 ncu --set full --clock-control none -k "regex:jpeg" -c 2 -f -o /tmp/${tag}_jpeg \
	python bench.py --workload thumbnail_jpeg --frames 64 --steps 1 --warmup 1 --no-cpu > $out/${tag}_ncu_jpeg.log 2>&1
echo "jpeg ncu rc=$?"
ncu -i /tmp/${tag}_jpeg.ncu-rep --page raw --csv > $out/${tag}_jpeg.raw.csv 2>/dev/null
# the headline kernel: random alpha (plain instantiation) and alpha 255 (voting instantiation, forced from the first launch)
for a in random opaque; do
	VB200_OPAQUE_PROBE=$([ $a = opaque ] && echo 2 || echo 1) timeout 600 ncu --set full --clock-control none --import-source on \
		-k "regex:thumbnail_fused_mma" -s 2 -c 1 -f -o /tmp/${tag}_head_$a \
		python bench.py --steps 2 --warmup 1 --frames 148 --no-cpu --e2e-frames 4 --alpha $a > $out/${tag}_ncu_head_$a.log 2>&1
	echo "head $a ncu rc=$?"
	python profiles/ncu_lines.py /tmp/${tag}_head_$a.ncu-rep 148 > $out/${tag}_head_${a}_ncu_summary.txt 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
	python bench.py --steps 2 --warmup 1 --no-cpu --e2e-frames 4 > $out/${tag}_launches.log 2>&1
echo "launch list rc=$?"
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"; cut -c1-400 $out/${tag}_bench.json
python bench.py --no-cpu --alpha opaque > $out/${tag}_bench_opaque.json 2>> $out/${tag}_bench.err; cut -c1-300 $out/${tag}_bench_opaque.json
du -sh $out
