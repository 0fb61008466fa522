#!/bin/bash
# round 2, pass be: closing pass on one GPU with the final code -- full GPU test suite, smoke, both bench arms, launch list,
# ncu captures of the dominant kernels (only the summaries travel back: gpurun_out is capped at 64 MiB)
mkdir -p gpurun_out
nproc
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2be_pytest_gpu.log 2>&1; echo "gpu tests exit $?"; tail -n 3 gpurun_out/r2be_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2be_smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/r2be_smoke.log
( time python bench.py --impl reference > gpurun_out/r2be_bench_reference.json 2> gpurun_out/r2be_bench_reference.log ) 2>&1 | tail -n 3
( time python bench.py > gpurun_out/r2be_bench.json 2> gpurun_out/r2be_bench.log ) 2>&1 | tail -n 3
echo "bench exit $?"; tail -n 3 gpurun_out/r2be_bench.log; wc -c gpurun_out/r2be_bench.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sw-long --no-strong --no-gatk --sw-len 512"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
    --log-file gpurun_out/r2be_launches.csv $CMD > gpurun_out/r2be_ncu_launch.log 2>&1; echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_duo_kernel --launch-skip 4 --launch-count 1 \
    -o gpurun_out/r2be_prof_sw_duo -f python bench.py --steps 1 --warmup 1 --workload sw --no-cpu-baseline --no-align > gpurun_out/r2be_ncu_sw.log 2>&1; echo "ncu sw exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hmm_duo_kernel --launch-skip 5 --launch-count 5 \
    -o gpurun_out/r2be_prof_hmm_duo -f python bench.py --steps 1 --warmup 1 --workload pairhmm --no-cpu-baseline --no-gatk > gpurun_out/r2be_ncu_hmm.log 2>&1; echo "ncu hmm exit $?"
for r in sw_duo hmm_duo; do python profiles/summarize_ncu.py gpurun_out/r2be_prof_$r.ncu-rep > gpurun_out/r2be_${r}_ncu.txt 2>&1; done
rm -f gpurun_out/r2be_prof_hmm_duo.ncu-rep gpurun_out/r2be_prof_sw_duo.ncu-rep
python profiles/launch_shares.py gpurun_out/r2be_launches.csv "$CMD" > gpurun_out/r2be_launch_shares.txt
ls -la gpurun_out | grep r2be; du -sh gpurun_out
