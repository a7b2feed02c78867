cd $GRAFT_REPO_ROOT
for i in 1 2; do
  for cfg in "RVC_V2_DEEP=0 RVC_V2_NARROW_M=32" "RVC_V2_DEEP=16 RVC_V2_NARROW_M=32" "RVC_V2_DEEP=16" "RVC_V2_DEEP=24" "RVC_V2_DEEP=0"; do
    echo -n "$cfg : "; env $cfg STEPS=300 python tools/quick_ms.py 2>&1 | grep -o "ms_per_window=[0-9.]*"
  done
done
RVC_V2_DEEP=16 python tools/lane_times.py | tail -1
RVC_V2_DEEP=0 RVC_V2_NARROW_M=32 python tools/lane_times.py | tail -1
