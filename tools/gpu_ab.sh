# A/B of liboptistate_kf.so builds on one box: variants/<name>/ against the in-tree build
export CASE_T=${CASE_T:-1000}
for v in "$@"; do
  echo "== $v"
  if [ "$v" = "tree" ]; then python tools/rate.py f64:summary f32:summary; else LD_LIBRARY_PATH=variants/$v python tools/rate.py f64:summary f32:summary; fi
done
