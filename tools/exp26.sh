cd $GRAFT_REPO_ROOT
run() { echo -n "$1 : "; env RVC_PDL_OPS="$1" python tools/lane_stamps.py 2>&1 | grep STAMPS; }
run "none"
run "rm.enc1,rm.enc2,rm.enc3,rm.enc4,rm.pool"
run "rm.enc"
run "rm.dec"
run "rm.dec1,rm.dec2,rm.dec3,rm.dec4,rm.cnn,rm.gi"
run "rm.mid"
run "rm."
run "sy."
