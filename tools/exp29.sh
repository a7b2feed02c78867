cd $GRAFT_REPO_ROOT
run() { echo -n "$1 : "; env RVC_PDL_OPS="$1" python tools/lane_stamps.py 2>&1 | grep -o "pool0.*"; }
run "rm.enc,sy."
run "rm.enc,sy.,rm.dec3,rm.dec4"
run "rm.enc,sy.,rm.dec0,rm.dec1"
run "rm.enc,sy.,rm.dec*c2"
run "rm.enc,sy.,rm.dec*c1"
run "rm.enc,sy.,rm.dec*up"
run "rm.enc*c2,sy."
run "rm.enc*c1,rm.enc*c2,sy."
