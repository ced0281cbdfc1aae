mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_codec.py -m gpu -q -x -k "png" 2>&1 | tail -4
timeout 600 python tests/perf/bench_encode.py 2>&1 | tail -1 > gpurun_out/bench_encode_v.json; python -c "
import json; e=json.load(open('gpurun_out/bench_encode_v.json'))
for k in ('png_encoder_device',): print(k, e.get(k))"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"crc|deflate|pack" -c 24 --csv --log-file gpurun_out/encode_launches_v.csv python tests/perf/bench_encode.py --reps 1 > /dev/null 2>&1
grep -E "crc|deflate|pack" gpurun_out/encode_launches_v.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*) / /; s/"$//' | awk '{n[$1]++; s[$1]+=$NF} END{for(k in n) printf "%-40s launches %3d  avg %.1f us\n", k, n[k], s[k]/n[k]/1000}' | sort | tee gpurun_out/encode_launches_summary_v.txt
