cd $GRAFT_REPO_ROOT
echo -n "stack : "; python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "spin 950us, full smem : "; RVC_EXP_SPIN=950 python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "spin 950us, no smem : "; RVC_EXP_SPIN=950,0 python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "spin 950us, full smem, G=16 : "; RVC_CVSTACK_G=16 RVC_EXP_SPIN=950 python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "spin 1us : "; RVC_EXP_SPIN=1 python tools/lane_stamps.py 2>&1 | grep STAMPS
