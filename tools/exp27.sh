cd $GRAFT_REPO_ROOT
run() { echo -n "$1 : "; env RVC_PDL_OPS="$1" python tools/lane_stamps.py 2>&1 | grep -o "gru.*"; echo -n "    "; env RVC_PDL_OPS="$1" STEPS=200 python tools/quick_ms.py 2>&1 | grep -o "ms_per_window=[0-9.]*"; }
run "none"
run "rm.enc,sy."
run "rm.enc,sy.,cv."
run "rm.enc,sy.,knn,phone,pitch"
run "rm.enc,rm.mid,rm.gi,rm.gru,f0,sy."
run "rm.enc,sy.stage"
run "rm.enc,sy.stage0,sy.stage1"
run "rm.enc,sy.stage2,sy.stage3,sy.audio,sy.conv_post"
