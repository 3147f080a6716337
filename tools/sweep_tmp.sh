python -m pytest tests/test_extract_gpu.py tests/test_dropin.py -m gpu -x -q 2>&1 | tail -2
python tools/stage_times.py 512
for v in old t256 t64; do ORB_B200_LIB=orb_slam2_chinesenotes_b200/lib/variants/liborb_b200_desc_$v.so python tools/stage_times.py 512; done
python tools/stage_times.py 64
python tools/stage_times.py 1
