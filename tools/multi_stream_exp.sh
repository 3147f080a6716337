for C in 512 256 128; do for M in 0 1; do
  ORB_FORCE_MULTI=$M python bench.py --only-main --no-cpu --chunk $C --steps 6 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk',$C,'multi',$M,'value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']))"
done; done
