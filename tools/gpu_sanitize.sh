mkdir -p gpurun_out/san
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_case.py > gpurun_out/san/$tool.txt 2>&1
  echo "== $tool"; tail -4 gpurun_out/san/$tool.txt
done
