# cluster-mode chain experiments (round 2): ms/window + lane times under several chain configurations
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" CHAINS=1 python tools/quick_ms.py > gpurun_out/x_$tag.log 2>&1; env "$@" python tools/lane_times.py >> gpurun_out/x_$tag.log 2>&1; }
run base RVC_NOP=1
run nocl RVC_CHAIN_CLUSTER=0
run unet16 RVC_CHAIN_SIDE=16 RVC_CHAIN_SIDE_MAXM=0 RVC_SC_LANE=0
run unet8 RVC_CHAIN_SIDE=8 RVC_CHAIN_SIDE_MAXM=0 RVC_SC_LANE=0
run deep16 RVC_CHAIN_SIDE=16 RVC_CHAIN_SIDE_MAXM=72 RVC_SC_LANE=0
run main16 RVC_CHAIN_MAIN=16
run main8 RVC_CHAIN_MAIN=8
run both16 RVC_CHAIN_MAIN=16 RVC_CHAIN_SIDE=16 RVC_CHAIN_SIDE_MAXM=0 RVC_SC_LANE=0
run bothdeep RVC_CHAIN_MAIN=16 RVC_CHAIN_SIDE=16 RVC_CHAIN_SIDE_MAXM=72 RVC_SC_LANE=0
grep -H -E "QUICK|LANE|^chain" gpurun_out/x_*.log
