cd $GRAFT_REPO_ROOT
for cfg in "RVC_PITCH_ML=0" "RVC_PITCH_ML=1" "RVC_PITCH_ML=1 RVC_SC_LANE=0" "RVC_PITCH_ML=1 RVC_CHAIN_SIDE=148"; do
  echo -n "$cfg : "; env $cfg python tools/lane_times.py 2>&1 | grep pitch
done
