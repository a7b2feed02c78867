cd $GRAFT_REPO_ROOT
echo -n "spin default prio : "; RVC_EXP_SPIN=950 python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "spin no prio : "; RVC_F0_PRIO=0 RVC_LANE1_PRIO=0 RVC_EXP_SPIN=950 python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "stack no prio : "; RVC_F0_PRIO=0 RVC_LANE1_PRIO=0 python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "spin, SC_LANE=0 : "; RVC_SC_LANE=0 RVC_EXP_SPIN=950 python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "spin G=1 : "; RVC_CVSTACK_G=1 RVC_EXP_SPIN=950 python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "spin PDL : "; RVC_PDL=1 RVC_EXP_SPIN=950 python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "alone PDL : "; RVC_PDL=1 RVC_PITCH_ML=1 MODE=pitch python tools/lane_stamps.py 2>&1 | grep STAMPS
