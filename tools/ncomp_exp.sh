for V in ${VARIANTS:-default ncomp1 ncomp2 default ncomp1}; do
  if [ $V = default ]; then unset ORB_B200_LIB; else export ORB_B200_LIB=$PWD/orb_slam2_chinesenotes_b200/lib/variants/liborb_b200_$V.so; fi
  python bench.py --only-main --no-cpu --steps 8 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$V','value',round(d['value']),'e2e',round(d['e2e']['value']),'submitted',round(d['e2e']['pipelined_submissions_frames_per_s']))"
done
