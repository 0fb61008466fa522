#!/bin/bash
# round 2, pass w: closing pass on one GPU -- full GPU test suite, smoke, both bench arms, launch list, ncu captures of the dominant kernels
# (the reports are summarised on the box; only the two small ones travel back: gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out
nproc
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2w_pytest_gpu.log 2>&1; echo "gpu tests exit $?"; tail -n 3 gpurun_out/r2w_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2w_smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/r2w_smoke.log
( time python bench.py --impl reference > gpurun_out/r2w_bench_reference.json 2> gpurun_out/r2w_bench_reference.log ) 2>&1 | tail -n 3
( time python bench.py > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.log ) 2>&1 | tail -n 3
echo "bench exit $?"; tail -n 3 gpurun_out/r2w_bench.log; wc -c gpurun_out/r2w_bench.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sw-long --no-strong --no-gatk --sw-len 512"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r2w_launches.csv $CMD > gpurun_out/r2w_ncu_launch.log 2>&1; echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:sw_duo_kernel --launch-skip 4 --launch-count 1 \
    -o gpurun_out/r2w_prof_sw_duo -f python bench.py --steps 1 --warmup 1 --workload sw --no-cpu-baseline > gpurun_out/r2w_ncu_sw.log 2>&1; echo "ncu sw exit $?"
ncu --set full --clock-control none --import-source on -k regex:hmm_duo_kernel --launch-skip 5 --launch-count 3 \
    -o gpurun_out/r2w_prof_hmm_duo -f python bench.py --steps 1 --warmup 1 --workload pairhmm --no-cpu-baseline > gpurun_out/r2w_ncu_hmm.log 2>&1; echo "ncu hmm exit $?"
ncu --set full --clock-control none --import-source on -k regex:"sw_duo_kernel|sw_walk_kernel" -c 3 -f \
    -o gpurun_out/r2w_prof_align python profiles/align_probe.py 250000 150 ends+align > gpurun_out/r2w_ncu_align.log 2>&1; echo "ncu align exit $?"
REPS=1 ncu --set full --clock-control none --import-source on -k regex:sw_longr_kernel -c 1 -f \
    -o gpurun_out/r2w_prof_sw_long python profiles/long_probe.py 1000000 200000 > gpurun_out/r2w_ncu_long.log 2>&1; echo "ncu long exit $?"
for r in sw_duo hmm_duo align sw_long; do python profiles/summarize_ncu.py gpurun_out/r2w_prof_$r.ncu-rep > gpurun_out/r2w_${r}_ncu.txt 2>&1; done
rm -f gpurun_out/r2w_prof_hmm_duo.ncu-rep gpurun_out/r2w_prof_sw_long.ncu-rep
python profiles/launch_shares.py gpurun_out/r2w_launches.csv "$CMD" > gpurun_out/r2w_launch_shares.txt
ls -la gpurun_out | grep r2w; du -sh gpurun_out
