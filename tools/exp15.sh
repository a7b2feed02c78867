cd $GRAFT_REPO_ROOT
for i in 1 2 3; do
  for cfg in "RVC_CBR=0" "RVC_CBR=1" "RVC_CBR=2"; do
    echo -n "$cfg : "; env $cfg STEPS=300 python tools/quick_ms.py 2>&1 | grep -o "ms_per_window=[0-9.]*"
  done
done
