#!/usr/bin/env bash
# One process per GPU on a single HGX B200 box: doc-sharded index + data-parallel reranker, peer-memory exchange.
set -euo pipefail
NGPU="${NGPU:-8}"
exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NGPU" --master-addr 127.0.0.1 --master-port "${PORT:-29500}" \
    bench.py --gpus "$NGPU" --steps "${STEPS:-50}" --warmup "${WARMUP:-5}" "$@"
