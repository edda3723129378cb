#!/bin/bash
# compute-sanitizer over scripts/sanitize_probe.py (every kernel family at a few thousand bodies).
#   bash scripts/sanitize.sh [outdir]       (on a GPU box; under gpurun: outdir = gpurun_out/sanitize)
# memcheck runs twice (plain launches: per-kernel attribution; captured-graph replay), racecheck / synccheck /
# initcheck once on the plain launches.  Exit code = number of tools that reported an error.
out=${1:-gpurun_out/sanitize}
mkdir -p "$out"
cs=/usr/local/cuda/bin/compute-sanitizer
fail=0
run() {   # name, probe args, tool args...
    local name=$1 pargs=$2; shift 2
    timeout 600 $cs --error-exitcode 9 --print-limit 20 "$@" python scripts/sanitize_probe.py $pargs > "$out/$name.log" 2>&1
    local rc=$?
    echo "$name rc=$rc $(grep -c 'ERROR SUMMARY' "$out/$name.log") summary: $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' "$out/$name.log" | tail -1)" | tee -a "$out/summary.txt"
    [ $rc -ne 0 ] && fail=$((fail + 1))
}
: > "$out/summary.txt"
run memcheck        ""         --tool memcheck --leak-check no
run memcheck_graphs "--graphs" --tool memcheck --leak-check no
run racecheck       ""         --tool racecheck --racecheck-report all
run synccheck       ""         --tool synccheck
run initcheck       ""         --tool initcheck
exit $fail
