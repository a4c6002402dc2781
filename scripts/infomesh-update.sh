#!/usr/bin/env bash
# Unattended update for a systemd-managed node (deploy/infomesh-update.{service,timer} call `infomesh update`, which
# does the same thing from Python; this shell variant is for hosts that manage the venv themselves).
#   install -m 0755 scripts/infomesh-update.sh /usr/local/bin/ && systemctl enable --now infomesh-update.timer
set -euo pipefail
VENV="${INFOMESH_VENV:-/opt/infomesh/.venv}"
PKG="${INFOMESH_PACKAGE:-infomesh-b200}"
say() { logger -t infomesh-update -- "$*" 2>/dev/null || true; printf '[%s] %s\n' "$(date -Is)" "$*"; }

before="$("$VENV/bin/python" -c 'import infomesh_b200 as m; print(m.__version__)')"
if ! "$VENV/bin/python" -m infomesh_b200 update --check | grep -q "Update available"; then
    say "up to date (v$before)"; exit 0
fi
say "updating $PKG from v$before"
"$VENV/bin/python" -m pip install --quiet --upgrade "$PKG"
"$VENV/bin/python" -m infomesh_b200.build          # rebuild the sm_100a kernels against the new sources
after="$("$VENV/bin/python" -c 'import infomesh_b200 as m; print(m.__version__)')"
if systemctl is-active --quiet infomesh; then
    say "restarting infomesh.service"; systemctl restart infomesh
fi
say "updated v$before -> v$after"
