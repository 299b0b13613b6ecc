#!/bin/bash
# compute-sanitizer passes over one small step of the bf16 hot path (the reference has no race / init checking).
# usage (on a GPU box): bash tools/sanitize.sh > gpurun_out/sanitizer.log 2>&1
# memcheck + racecheck + initcheck + synccheck on tools/sanitize_step.py (6 experts, B = 4, bf16, train mode: fused mixer,
# fused MLP, grouped heads, router forward / backward, TMA-staged combine, CTC lattice, clip + Adam, hard-routed inference).
set -u
cd "$(dirname "$0")/.."
# the instrumented kernels run orders of magnitude slower: rebuild with the mbarrier wait guard raised from 2 s to 10 min
# (the box's copy only; restore with `python -m mrn_b200.build --force`)
MRNB_NVCC_EXTRA="-DMRNB_WAIT_TRAP_NS=600000000000ull" python -m mrn_b200.build --force > /dev/null
for tool in ${TOOLS:-memcheck racecheck initcheck synccheck}; do
  echo "=== compute-sanitizer --tool $tool"
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_step.py 2>&1 | grep -v "^$" > gpurun_out/_san_$tool.txt; head -${HEAD:-30} gpurun_out/_san_$tool.txt | cut -c1-300; echo ...; tail -4 gpurun_out/_san_$tool.txt | cut -c1-300
done
