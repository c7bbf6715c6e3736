ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bigvgan.csv python bench.py --leg bigvgan --steps 1 --warmup 1 > gpurun_out/r02_launch_voc.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_bigvgan.csv > gpurun_out/r02_launches_bigvgan_summary.txt; cat gpurun_out/r02_launches_bigvgan_summary.txt
ncu --set full --clock-control none --import-source on -k regex:aa_snake_pair -s 30 -c 2 -o gpurun_out/r02_aa_pair_full -f python bench.py --leg bigvgan --steps 1 --warmup 0 > gpurun_out/r02_ncu_aa.log 2>&1
ncu -i gpurun_out/r02_aa_pair_full.ncu-rep --page raw --csv > gpurun_out/r02_aa_pair_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02_aa_pair_full_raw.csv | head -50
