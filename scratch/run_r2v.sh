# round 2, call V: launch list of the default bench command's render (c5), every kernel of the timed region
C="python bench.py --workload c5 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
python profiles/source_sha.py > gpurun_out/r2v_sha.txt
$C > gpurun_out/r2v_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2v_launches.csv $C > gpurun_out/r2v_ncu.log 2>&1
tail -n 2 gpurun_out/r2v_ncu.log; wc -l gpurun_out/r2v_launches.csv
