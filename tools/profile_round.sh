#!/bin/bash
# Evidence of a round (run on the GPU box through gpurun; outputs under gpurun_out/, summaries are then written to
# profiles/ with tools/launch_summary.py, ncu_summary.py, ncu_brief.py, ncu_traffic.py): tests, bench lines, ncu launch list, ncu full-set capture of one chunk of every kernel.
set -x
cd "$GRAFT_REPO_ROOT"
T=${1:-r01e}      # file-name tag of this capture
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${T}_pytest_gpu.txt
python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_n1.json 2> gpurun_out/${T}_bench_reference_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1
# launch list of the timed steps (after the bench has exited 0 without ncu)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -s 84 -c 56 --csv --log-file gpurun_out/${T}_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --only-main > gpurun_out/${T}_ncu_launches.log 2>&1
# full set: one chunk (512 frames) of every extractor kernel = launches 0..13 of the first warm-up step
ncu --set full --clock-control none --import-source on -k regex:^k_ -c 14 -o gpurun_out/${T}_prof_extract -f python bench.py --steps 1 --warmup 3 --no-cpu --only-main > gpurun_out/${T}_ncu_full.log 2>&1
ORB_BENCH_PROFILE=1 ncu --set full --clock-control none --import-source on -k regex:"k_match_fixpoint|k_init_fixpoint|k_bow_fixpoint|k_distinctive|k_project_points|k_hamming_bf" -c 14 -o gpurun_out/${T}_prof_match -f python tools/match_once.py > gpurun_out/${T}_ncu_full_match.log 2>&1
ls -la gpurun_out
