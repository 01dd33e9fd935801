# compute-sanitizer on small shapes: memcheck, racecheck, synccheck, initcheck of the hot-path kernels
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/r02
for tool in memcheck racecheck synccheck; do
  EVK_SAN_EVENTS=300000 timeout 1500 compute-sanitizer --tool $tool --print-limit 20 \
      python profiles/tools/sanitize_small.py > gpurun_out/r02/sanitizer_$tool.txt 2>&1
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_small OK|Error|hazard" gpurun_out/r02/sanitizer_$tool.txt | head -8
done
