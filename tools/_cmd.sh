mkdir -p gpurun_out
timeout 600 python tools/bench_encode.py 2>&1 | tail -1 > gpurun_out/bench_encode_n.json; python -c "
import json; e=json.load(open('gpurun_out/bench_encode_n.json'))
for k in ('exr_decoder_T1','exr_decoder_T16','png_decoder','reference_lodepng_decode'): print(k, e.get(k))"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"unpack" -c 6 --csv --log-file gpurun_out/decode_launches_n.csv python tools/bench_encode.py --reps 1 > /dev/null 2>&1
grep -E "unpack" gpurun_out/decode_launches_n.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -6
