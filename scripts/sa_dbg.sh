cd /root/repo
for m in 31 63 95 127; do
  PDF_SA_DEBUG=$m python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dbg=$m', 'sa1', j['stages_ms']['sa1'], 'sa2', j['stages_ms']['sa2'])"
done
