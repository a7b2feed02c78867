cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" CHAINS=1 timeout 300 python tools/quick_ms.py > gpurun_out/y_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/y_$tag.log; env "$@" timeout 300 python tools/lane_times.py >> gpurun_out/y_$tag.log 2>&1; }
run stack RVC_NOP=1
run nostack RVC_CVSTACK=0
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/y_tests.log 2>&1
grep -H -E "QUICK|LANE|rc=|Error|error" gpurun_out/y_*.log | head -40
tail -15 gpurun_out/y_tests.log
