set -x
# (1) launch list of one bench step (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-extras > gpurun_out/r2_bench_under_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_launches_bench.csv > gpurun_out/r2_launches_bench_summary.csv
# (2) full captures of the three kernels (small problems keep the replays short)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stream_pass_kernel -s 2 -c 1 -o gpurun_out/r2_prof_pass python scripts/microbench.py --m 1048576 --only pass > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sjlt_apply_kernel -s 4 -c 1 -o gpurun_out/r2_prof_sjlt python scripts/microbench.py --m 1048576 --only sjlt > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:qr_block_coop -s 17 -c 1 -o gpurun_out/r2_prof_qrcoop python scripts/qr_once.py 8192 2049 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gauss_sketch_kernel -s 1 -c 1 -o gpurun_out/r2_prof_gauss python scripts/microbench.py --m 262144 --only gauss > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
