mkdir -p gpurun_out
timeout 600 python tools/bench_encode.py 2>&1 | tail -3 | tee gpurun_out/bench_encode_j.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"deflate|pack" -c 40 --csv --log-file gpurun_out/encode_launches_j.csv python tools/bench_encode.py --reps 1 > /dev/null 2>&1
grep -E "deflate|pack" gpurun_out/encode_launches_j.csv | awk -F'","' '{print $5, $NF}' | sort | uniq -c | sort -rn | head -20
