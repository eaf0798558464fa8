#!/bin/bash
# compute-sanitizer over the small parity cases, all four tools; logs under gpurun_out/sanitizer/ (copy the
# ones to keep into profiles/).  Usage (on the GPU box): scripts/sanitize.sh [tag]
tag=${1:-run}
out=gpurun_out/sanitizer
mkdir -p $out
rc=0
for tool in memcheck racecheck synccheck initcheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 50 --log-file $out/${tag}_$tool.log \
        python tests/tools/sanitize_cases.py > $out/${tag}_$tool.stdout 2>&1
    echo "$tool: exit $? ; $(grep -c 'ERROR SUMMARY' $out/${tag}_$tool.log) summaries; $(grep 'ERROR SUMMARY' $out/${tag}_$tool.log | tail -1)" | tee -a $out/${tag}_summary.txt
done
