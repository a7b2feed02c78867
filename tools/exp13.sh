cd $GRAFT_REPO_ROOT
M="mel,rm.pool0,rm.pool1,rm.pool2,rm.pool4,rm.mid3.3.c2,rm.dec1.3.c2,rm.dec2.3.c2,rm.dec3.3.c2,rm.dec4.3.c2,rm.cnn,rm.gru,f0,cv.conv6,knn_select,pitch,sy.enc5,sy.flow0,sy.stage0,sy.stage3,sy.audio"
for cfg in "RVC_CBR=0" "RVC_CBR=1"; do
  echo "== $cfg"
  env $cfg RVC_TL_MARKS=$M python tools/timeline.py 2>&1 | sed -n 1,30p
done
