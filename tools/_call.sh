set -x
for i in 1 2; do
PRAM_TWO_STREAMS=0 timeout 600 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ONE STREAM ', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['frames_within_5deg_5cm'])"
PRAM_TWO_STREAMS=1 timeout 600 python bench.py --no-cpu-baseline --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('TWO STREAMS', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['frames_within_5deg_5cm'])"
done
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5)
