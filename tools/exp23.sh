cd $GRAFT_REPO_ROOT
echo -n "bn48 : "; python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "bn48 : "; STEPS=300 python tools/quick_ms.py 2>&1 | grep -o "ms_per_window=[0-9.]*"
cp obs-rvc_b200/librvc_b200.so /tmp/orig.so
cp obs-rvc_b200/alt/librvc_b200_bn96.so obs-rvc_b200/librvc_b200.so
timeout 300 python -m pytest tests/test_gpu_round2.py -q -m gpu -x -k "cvstack" 2>&1 | tail -3
echo -n "bn96 : "; python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "bn96 : "; STEPS=300 python tools/quick_ms.py 2>&1 | grep -o "ms_per_window=[0-9.]*"
echo -n "bn96 G=32: "; RVC_CVSTACK_G=32 python tools/lane_stamps.py 2>&1 | grep STAMPS
echo -n "bn96 G=32: "; RVC_CVSTACK_G=32 STEPS=300 python tools/quick_ms.py 2>&1 | grep -o "ms_per_window=[0-9.]*"
