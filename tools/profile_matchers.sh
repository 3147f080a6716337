#!/bin/bash
# Evidence of a state of the code whose EXTRACTOR kernels are unchanged since the last tools/profile_round.sh capture: tests, bench
# lines (ours + reference arm), smoke, and the full-set ncu capture of the matcher / map-point kernels only.
set -x
cd "$GRAFT_REPO_ROOT"
T=${1:-r02j}
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/${T}_pytest_gpu.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_n1.json 2> gpurun_out/${T}_bench_reference_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1
ORB_BENCH_PROFILE=1 ncu --set full --clock-control none --import-source on -k regex:"k_match_fixpoint|k_init_fixpoint|k_bow_fixpoint|k_window_best_free|k_distinctive|k_project_points|k_hamming_bf" -c 20 -o gpurun_out/${T}_prof_match -f python tools/match_once.py > gpurun_out/${T}_ncu_full_match.log 2>&1
ncu -i gpurun_out/${T}_prof_match.ncu-rep --page raw --csv > gpurun_out/${T}_match_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
