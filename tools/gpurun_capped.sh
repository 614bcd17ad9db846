#!/bin/bash
# gpurun with a spending cap. An N-GPU call is charged N x its wall time, INCLUDING time spent hanging until a
# timeout fires: round 1 lost 89 of its 180 GPU-minutes to one 8-GPU call whose inner `timeout 600` was a round
# number instead of a multiple of the known run time. This wrapper makes the worst case explicit and refuses it
# when it exceeds --max-minutes or what is left.
#
#   tools/gpurun_capped.sh --gpus 8 --inner 170 --max-minutes 30 -- '<command; every leg under its own timeout>'
#
# --inner = the sum of the inner `timeout` values of the command's legs (seconds). Worst case charged =
# gpus x (inner + 60 s box overhead). The gpurun --timeout is set to inner + 30 so that gpurun's own limit,
# not a hung leg, ends the call.
set -euo pipefail
GPUS=1; INNER=""; MAXMIN=""
while [ $# -gt 0 ]; do
  case "$1" in
    --gpus) GPUS=$2; shift 2;;
    --inner) INNER=$2; shift 2;;
    --max-minutes) MAXMIN=$2; shift 2;;
    --) shift; break;;
    *) echo "unknown option $1" >&2; exit 64;;
  esac
done
[ -n "$INNER" ] && [ -n "$MAXMIN" ] && [ $# -ge 1 ] || { echo "usage: $0 [--gpus N] --inner SECONDS --max-minutes M -- '<command>'" >&2; exit 64; }
WORST=$(python3 -c "print(round($GPUS * ($INNER + 60) / 60.0, 1))")
LEFT=$(/usr/local/graft/bin/gpurun --status 2>/dev/null | python3 -c "import json,sys; print(json.load(sys.stdin).get('gpu_minutes_left', 0))" 2>/dev/null || echo 0)
echo "[capped] worst case ${WORST} GPU-min (${GPUS} GPU x (${INNER}+60) s); cap ${MAXMIN}; left this round ${LEFT}" >&2
python3 - "$WORST" "$MAXMIN" "$LEFT" <<'PY'
import sys
worst, cap, left = map(float, sys.argv[1:4])
if worst > cap:
    sys.exit(f"[capped] refused: worst case {worst} > --max-minutes {cap}")
if worst > left:
    sys.exit(f"[capped] refused: worst case {worst} > {left} GPU-minutes left")
PY
ARGS=(--timeout $((INNER + 30)))
[ "$GPUS" -gt 1 ] && ARGS+=(--gpus "$GPUS")
exec /usr/local/graft/bin/gpurun "${ARGS[@]}" -- "$1"
