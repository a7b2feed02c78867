cd $GRAFT_REPO_ROOT
echo -n "alone : "; RVC_PITCH_ML=1 MODE=pitch python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "infer : "; python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "infer CVSTACK=0 : "; RVC_CVSTACK=0 python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "infer CBR=0 : "; RVC_CBR=0 python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "alone CBR=0 : "; RVC_CBR=0 RVC_PITCH_ML=1 MODE=pitch python tools/lane_stamps.py 2>&1 | grep STAMPS
