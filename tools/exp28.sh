cd $GRAFT_REPO_ROOT
run() { echo "== $1"; env RVC_PDL_OPS="$1" $2 python tools/lane_stamps.py 2>&1 | grep "STAMPS\|CHAIN"; }
run "none"
run "rm.enc,sy."
run "none" "RVC_EXP_SPIN=1"
echo "== alone"; RVC_PITCH_ML=1 MODE=pitch python tools/lane_stamps.py 2>&1 | grep "STAMPS\|CHAIN"
