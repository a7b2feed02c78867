cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" CHAINS=1 timeout 300 python tools/quick_ms.py > gpurun_out/v_$tag.log 2>&1; echo "rc=$?" >> gpurun_out/v_$tag.log; }
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/v_tests.log 2>&1
run base RVC_NOP=1
run nostack RVC_CVSTACK=0
run unet64 RVC_CHAIN_SIDE=64 RVC_CHAIN_SIDE_MAXM=0 RVC_SC_LANE=0
run unet84 RVC_CHAIN_SIDE=84 RVC_CHAIN_SIDE_MAXM=0 RVC_SC_LANE=0
run unet84ns RVC_CVSTACK=0 RVC_CHAIN_SIDE=84 RVC_CHAIN_SIDE_MAXM=0 RVC_SC_LANE=0
run deep64 RVC_CHAIN_SIDE=64 RVC_CHAIN_SIDE_MAXM=72 RVC_SC_LANE=0
grep -H -E "QUICK|rc=|^chain" gpurun_out/v_*.log
tail -5 gpurun_out/v_tests.log
