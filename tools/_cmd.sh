mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_codec.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/pytest_gpu_l.log; tail -15 gpurun_out/pytest_gpu_l.log
timeout 600 python tools/bench_encode.py 2>&1 | tail -1 > gpurun_out/bench_encode_l.json; python -c "
import json; e=json.load(open('gpurun_out/bench_encode_l.json')); print(e['png_encoder_device']); print(e['exr_encoder_device'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"deflate|pack" -c 30 --csv --log-file gpurun_out/encode_launches_l.csv python tools/bench_encode.py --reps 1 > /dev/null 2>&1
grep -E "deflate|pack" gpurun_out/encode_launches_l.csv | awk -F'","' '{print $5, $NF}' | sed 's/(.*) / /; s/"$//' | awk '{n[$1]++; s[$1]+=$NF} END{for(k in n) printf "%-40s launches %3d  avg %.1f us\n", k, n[k], s[k]/n[k]/1000}' | sort | tee gpurun_out/encode_launches_summary_l.txt
