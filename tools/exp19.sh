cd $GRAFT_REPO_ROOT
for cfg in "RVC_V2_WANT=296" "RVC_V2_WANT=222" "RVC_V2_WANT=168" "RVC_V2_WANT=128" "RVC_V2_WANT=84" "RVC_V2_WANT=400"; do
  echo -n "$cfg : "; env $cfg python tools/lane_stamps.py 2>&1 | grep STAMPS
done
