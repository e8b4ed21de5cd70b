mkdir -p gpurun_out
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "regex:plan_rows|gather_rows|replay_rows" -c 6 --csv --log-file gpurun_out/assemble_list.csv python bench.py --steps 1 --warmup 0 --new-tokens 4 --no-e2e --no-cpu-baseline --no-profile > /dev/null 2>&1
grep -E "plan_rows|gather_rows|replay_rows" gpurun_out/assemble_list.csv | grep time_duration | awk -F'","' '{print substr($5,1,40), $(NF-1), $NF}' | head -8
