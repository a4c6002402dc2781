#!/usr/bin/env bash
# Installation smoke test: does a fresh checkout build, import, and answer a query?
#   bash tests/infra/test_install.sh            # CPU steps only
#   bash tests/infra/test_install.sh --gpu      # also the device smoke test (needs a B200)
# Each step is recorded as PASS / FAIL / SKIP; the exit code is the number of failures.
set -uo pipefail

cd "$(dirname "$0")/../.."
WANT_GPU=0
[[ "${1:-}" == "--gpu" ]] && WANT_GPU=1
RESULTS="$(mktemp)"
fails=0

step() {                      # step <name> <command...>
  local name="$1"; shift
  local log; log="$(mktemp)"
  if "$@" >"$log" 2>&1; then
    printf '%-34s PASS\n' "$name" | tee -a "$RESULTS"
  else
    printf '%-34s FAIL  (%s)\n' "$name" "$(tail -n 1 "$log")" | tee -a "$RESULTS"
    fails=$((fails + 1))
  fi
  rm -f "$log"
}

skip() { printf '%-34s SKIP  (%s)\n' "$1" "$2" | tee -a "$RESULTS"; }

echo "infomesh_b200 installation test — $(date -u +%Y-%m-%dT%H:%M:%SZ)"
echo "python: $(python --version 2>&1)   nvcc: $(nvcc --version 2>/dev/null | tail -n 1 || echo missing)"

step "python >= 3.12"              python -c 'import sys; assert sys.version_info >= (3, 12)'
step "torch importable"            python -c 'import torch'
if command -v nvcc >/dev/null; then
  step "native build (sm_100a)"    python -c 'import __graft_entry__ as g; g.build()'
  step "library loads"             python -c 'from infomesh_b200 import _native; assert _native.lib(build_if_missing=False).im_abi_version() >= 1'
else
  skip "native build (sm_100a)" "nvcc not on PATH"
fi
step "package imports"             python -c 'import infomesh_b200; from infomesh_b200.cli import cli'
step "cli --help"                  python -m infomesh_b200 --help
step "cli config show"             env INFOMESH_NODE_DATA_DIR="$(mktemp -d)" python -m infomesh_b200 config show
DATA="$(mktemp -d)"
step "index + search round trip"   env INFOMESH_NODE_DATA_DIR="$DATA" python - <<'PY'
import os
from infomesh_b200.index.local_store import LocalStore
from infomesh_b200.search.query import search_local
st = LocalStore(os.path.join(os.environ["INFOMESH_NODE_DATA_DIR"], "index.db"))
st.add_document(url="https://example.org/a", title="Install check", text="the installation smoke test indexes one page " * 5, raw_html_hash="r", text_hash="t")
hits = search_local(st, "installation smoke", limit=3).results
assert hits and hits[0].url == "https://example.org/a", hits
st.close()
PY
step "examples/basic_search.py"    env PYTHONPATH="$PWD" python examples/basic_search.py
step "cpu test subset"             python -m pytest tests/test_foundation.py tests/test_index_cpu.py -x -q -m "not gpu"
if [[ $WANT_GPU -eq 1 ]]; then
  step "device smoke test"         python -c 'import __graft_entry__ as g; g.smoke()'
  step "bench --quick"             python bench.py --quick --steps 5 --warmup 3
else
  skip "device smoke test" "run with --gpu on a B200"
fi
rm -rf "$DATA"

echo "------------------------------------------"
cat "$RESULTS"
echo "failures: $fails"
rm -f "$RESULTS"
exit "$fails"
