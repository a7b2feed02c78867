cd $GRAFT_REPO_ROOT
for cfg in "RVC_CV_GATE=-1" "RVC_CV_GATE=0" "RVC_CV_GATE=1" "RVC_CV_GATE=2"; do
  echo -n "$cfg : "; env $cfg python tools/lane_stamps.py 2>&1 | grep STAMPS
  echo -n "$cfg : "; env $cfg STEPS=300 python tools/quick_ms.py 2>&1 | grep -o "ms_per_window=[0-9.]*"
done
