python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --records 20000000 --steps 3 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']/1e6,1), 'e2e', d['e2e'])"
