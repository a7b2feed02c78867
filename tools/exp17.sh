cd $GRAFT_REPO_ROOT
for i in 1 2; do
  for cfg in "RVC_KNN_WPARTS=0" "RVC_KNN_WPARTS=148" "RVC_KNN_WPARTS=96" "RVC_KNN_WPARTS=64" "RVC_KNN_WPARTS=32"; do
    echo -n "$cfg : "; env $cfg STEPS=300 python tools/quick_ms.py 2>&1 | grep -o "ms_per_window=[0-9.]*"
  done
done
