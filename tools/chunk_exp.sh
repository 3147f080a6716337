# end-to-end rate of the host-buffer path against the chunk size (bench --chunk also sets the device-resident chunk: read e2e only)
for C in ${CHUNKS:-0 96 160 192 256 0}; do
  python bench.py --only-main --no-cpu --chunk $C --steps 8 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunk',$C,'e2e',round(d['e2e']['value']),'submitted',round(d['e2e']['pipelined_submissions_frames_per_s']))"
done
