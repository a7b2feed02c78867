cd $GRAFT_REPO_ROOT
for cfg in "RVC_CBR=0" "RVC_CBR=2" "RVC_CBR=1" "RVC_KNN_WPARTS=148" "RVC_CV_WANT=96" "RVC_CVSTACK=0"; do
  echo -n "$cfg : "; env $cfg python tools/lane_stamps.py 2>&1 | grep STAMPS
done
