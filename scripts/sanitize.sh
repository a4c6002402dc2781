#!/usr/bin/env bash
# compute-sanitizer sweep over the kernel tests.  Run on a GPU box:  scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck]
# Small shapes only (the sanitizer slows kernels 10-100x); results land in gpurun_out/sanitizer_<tool>.log.
set -uo pipefail
TOOL="${1:-memcheck}"
export PYTHONPATH=.
mkdir -p gpurun_out
SEL="${SANITIZE_SEL:-test_topk_merge or test_norm_embed or test_bm25 or test_passage_simhash or test_symmetric_heap or test_attention_decode or test_sim_topk_exact}"
timeout "${SANITIZE_TIMEOUT:-900}" compute-sanitizer --tool "$TOOL" --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_kernels.py -q -x -k "$SEL" > "gpurun_out/sanitizer_${TOOL}.log" 2>&1
rc=$?
tail -5 "gpurun_out/sanitizer_${TOOL}.log"
grep -c "ERROR SUMMARY: 0 errors" "gpurun_out/sanitizer_${TOOL}.log" || true
exit $rc
