# quick GPU sanity: stage SNRs (tiny + full) under a timeout, then a short bench with per-stage times
timeout 120 python tests/tools/debug_stages.py tiny fp16 2 9 2>&1 | grep -E "init_conv|block3|pcm|rror"
timeout 180 python tests/tools/debug_stages.py full fp16 2 20 2>&1 | grep -E "pre_transformer|upsample1|init_conv|block0|block1|block2|block3|pcm|rror"
timeout 600 python bench.py --steps ${STEPS:-3} --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('audio-s/s', round(d['value']), 'ms/step', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value']), d['clocks'])
print([(s['name'],s['ms'],s['tflops']) for s in d['roofline']['stages']])"
